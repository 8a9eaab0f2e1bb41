cd /root/repo
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 500 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/pytest_dist2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_dist2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e > gpurun_out/bench_n2_march.json 2> gpurun_out/bench_n2_march.err; echo "bench rc=$?"
cat gpurun_out/bench_n2_march.json | cut -c1-1500; tail -3 gpurun_out/bench_n2_march.err | cut -c1-300
