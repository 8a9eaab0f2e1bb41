#!/bin/bash
# round 1, call 68: full -m gpu suite, SpMM bench (chunk variants), ncu of the line-marching SpMM
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1d.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_r1d.log
timeout 120 python tools/bench_spmm.py > gpurun_out/spmm_bench2.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/spmm_bench2.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:kb_spmm_lines -c 1 -f -o gpurun_out/prof_lines python tools/bench_spmm.py --quick > gpurun_out/ncu_lines.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_lines.log
