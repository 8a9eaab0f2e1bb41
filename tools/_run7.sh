cd /root/repo
timeout 600 python -m pytest tests/test_gpu_shortrec.py -q -m gpu > gpurun_out/pytest_shortrec.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|error" gpurun_out/pytest_shortrec.log | tail -15
grep -E "^FAILED|^ERROR" gpurun_out/pytest_shortrec.log | head -30
