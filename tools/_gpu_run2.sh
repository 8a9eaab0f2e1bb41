cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py gpurun_out/r2y_dist2.json 48 > gpurun_out/r2y_dist2.log 2>&1
tail -3 gpurun_out/r2y_dist2.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err
tail -2 gpurun_out/r2y_bench_n2.err | cut -c1-300
head -c 300 gpurun_out/r2y_bench_n2.json; echo
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y_dist2.json'))
for k,v in d.items():
    if isinstance(v,dict): print(k, {a:b for a,b in v.items() if not isinstance(b,(list,dict))})
d=json.loads(open('gpurun_out/r2y_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e'].get('value'), d['parity']['ok'], d['parity']['max_rel_resnorm_err'])
PY
