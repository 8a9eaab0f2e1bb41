cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Warning | tail -3
timeout 600 python tools/bench_spmm_var.py 2>&1 | grep -v Warn | tee gpurun_out/r2s_spmm_var.txt
