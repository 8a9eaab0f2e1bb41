cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "merge or bit_exact_vs_scipy" 2>&1 | tail -3 > gpurun_out/r2h_pytest_merge.log
cat gpurun_out/r2h_pytest_merge.log
timeout 900 python tools/spmv_general_bench.py > gpurun_out/r2h_spmv_general.txt 2>&1
grep -v Warning gpurun_out/r2h_spmv_general.txt | grep -v "S = torch"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_spmv_merge -s 4 -c 1 -o gpurun_out/r2h_merge_banded python tools/spmv_general_bench.py --one banded_100 merge 26 6 > gpurun_out/r2h_ncu.log 2>&1
tail -2 gpurun_out/r2h_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kb_spmv_merge -s 4 -c 1 -o gpurun_out/r2h_merge_powerlaw python tools/spmv_general_bench.py --one powerlaw merge 26 6 > gpurun_out/r2h_ncu2.log 2>&1
tail -2 gpurun_out/r2h_ncu2.log
