cd /root/repo
timeout 900 python -m pytest tests/test_gpu_shortrec.py -x -q -m gpu -k "lookahead" 2>&1 | grep -v Warn | tail -25
