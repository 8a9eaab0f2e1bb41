#!/bin/bash
# round 1, call 69: line-marching SpMM with running indices and plane-fastest work items
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -k "line_marching or spmv_bit_exact or fused_modes" > gpurun_out/pytest_lines2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_lines2.log
timeout 120 python tools/bench_spmm.py > gpurun_out/spmm_bench3.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/spmm_bench3.txt
