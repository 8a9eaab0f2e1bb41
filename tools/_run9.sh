#!/bin/bash
# round 1, call 66: full -m gpu suite (incl. the new utils / ingest tests), block kernel bench, ncu of the block kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1c.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r1c.log
python tools/bench_block.py > gpurun_out/block_bench.txt 2>&1; echo "bench rc=$?"
cat gpurun_out/block_bench.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:kb_block -c 2 -f -o gpurun_out/prof_block python tools/bench_block.py --quick > gpurun_out/ncu_block.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_block.log
