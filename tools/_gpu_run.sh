cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warning | tail -4 > gpurun_out/r2o_pytest_gpu.log
tail -4 gpurun_out/r2o_pytest_gpu.log
timeout 300 python tools/bench_configs.py 2>&1 | grep "C1\|C2\|C3" | tee gpurun_out/r2o_configs.txt
