cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "persistent or fused_marching" 2>&1 | tail -2
python tools/small_cg_bench.py 256 128 512 2>&1 | grep -v Warn | tee gpurun_out/r2m_small_cg.txt
