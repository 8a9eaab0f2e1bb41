cd /root/repo
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "marching" > gpurun_out/pytest_march.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_march.log
timeout 400 python tools/march_bench.py 512 quick > gpurun_out/march_bench_512.log 2>&1; echo "bench rc=$?"
tail -30 gpurun_out/march_bench_512.log
