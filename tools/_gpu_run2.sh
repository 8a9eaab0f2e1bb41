cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "merge_schedule_vs_scipy and powerlaw" 2>&1 | grep -v Warning | tail -40
