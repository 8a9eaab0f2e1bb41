cd /root/repo
timeout 900 python -m pytest tests/test_gpu_complex.py -x -q -m gpu 2>&1 | grep -v Warn | tail -15
