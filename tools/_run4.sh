cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_full.log
timeout 600 python bench.py > gpurun_out/bench_n1_march.json 2> gpurun_out/bench_n1_march.err; echo "bench rc=$?"
cat gpurun_out/bench_n1_march.json; tail -3 gpurun_out/bench_n1_march.err
