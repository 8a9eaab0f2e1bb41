cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -3
python tools/part_single.py 64 50 2>&1 | tee gpurun_out/r2_part_single4.txt
python tools/part_single.py 128 50 2>&1 | tee -a gpurun_out/r2_part_single4.txt
KRYLOV_B200_TUNE="24=100000" MC="0:0,4:4,4:5,-1:-1" python tools/part_single.py 512 30 2>&1 | tee -a gpurun_out/r2_part_single4.txt
