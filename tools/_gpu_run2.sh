cd /root/repo
python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-general --no-parity --e2e-trace > gpurun_out/r2q_e2e.json 2> gpurun_out/r2q_e2e_trace.txt
grep e2e-trace gpurun_out/r2q_e2e_trace.txt
