cd /root/repo
timeout 600 python -m pytest tests/test_gpu_shortrec.py -q -m gpu > gpurun_out/pytest_shortrec.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_shortrec.log | tail -3
grep -E "^FAILED|^ERROR|^E  " gpurun_out/pytest_shortrec.log | head -30
timeout 600 python tools/bench_shortrec.py 256 > gpurun_out/shortrec_bench.log 2>&1; echo "bench rc=$?"
tail -12 gpurun_out/shortrec_bench.log
