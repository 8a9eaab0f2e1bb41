# Row-partitioned runs on one 8-GPU box: 8-rank parity check (tools/dist_check.py), then bench.py at N = 8, 4, 2
# (what profiles/r2_bench_n{2,4,8}_b.json and r2_dist8_parity_b.json come from).  usage: gpurun --gpus 8 -- bash tools/scale_n842.sh
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py gpurun_out/r2c_dist8.json 48 > gpurun_out/r2c_dist8.log 2>&1
tail -3 gpurun_out/r2c_dist8.log | cut -c1-300
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 200 --warmup 10 --no-configs > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err
tail -2 gpurun_out/r2c_bench_n$N.err | cut -c1-300
head -c 400 gpurun_out/r2c_bench_n$N.json; echo
done
