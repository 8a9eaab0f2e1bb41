cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "merge or bit_exact_vs_scipy" 2>&1 | tail -2
timeout 900 python tools/spmv_general_bench.py 2>&1 | grep -v "Warn\|S = torch" | grep "==\|cuSPARSE\|cfg=0 order=2\|cfg=7 order=2\|stream" | tee gpurun_out/r2x_spmv_general.txt
