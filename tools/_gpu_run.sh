cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warning | tail -3 > gpurun_out/r2_final_pytest_gpu.log
tail -3 gpurun_out/r2_final_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2
