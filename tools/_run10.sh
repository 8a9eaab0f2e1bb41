#!/bin/bash
# round 1, call 67: line-marching SpMM -- parity tests (with a hang guard), then its bench
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -k "line_marching or spmv_bit_exact or fused_modes or windowed" > gpurun_out/pytest_lines.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_lines.log
timeout 150 python tools/bench_spmm.py > gpurun_out/spmm_bench.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/spmm_bench.txt
