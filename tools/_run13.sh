#!/bin/bash
# round 1, call 70: final full -m gpu suite on the committed tree; short-recurrence solvers at 384^3
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1e.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r1e.log
timeout 100 python tools/bench_shortrec.py 384 > gpurun_out/shortrec_bench_384.log 2>&1; echo "bench rc=$?"
cat gpurun_out/shortrec_bench_384.log | tail -12
