cd /root/repo
timeout 900 python -m pytest tests/test_gpu_shortrec.py -x -q -m gpu 2>&1 | grep -v Warn | tail -12
timeout 600 python tools/bench_shortrec.py 256 2>&1 | grep -v Warn | tee gpurun_out/r2v_shortrec_256.txt
