#!/bin/bash
# round 1, call 71: ncu of the committed line-marching SpMM and of kb_block_apply
mkdir -p gpurun_out
timeout 45 ncu --set full --clock-control none --import-source on -k regex:kb_spmm_lines -c 1 -f -o gpurun_out/prof_lines2 python tools/bench_spmm.py --quick > gpurun_out/ncu_lines2.log 2>&1; echo "ncu1 rc=$?"
timeout 40 ncu --set full --clock-control none --import-source on -k regex:kb_block_apply -c 1 -f -o gpurun_out/prof_block_apply python tools/bench_block.py --quick > gpurun_out/ncu_block_apply.log 2>&1; echo "ncu2 rc=$?"
