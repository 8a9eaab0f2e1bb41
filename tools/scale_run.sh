#!/bin/bash
# 1/2/4/8-GPU scaling run of bench.py on one box (what the driver does at round end).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py gpurun_out/dist8.json 48 > gpurun_out/dist8.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/dist8.json"))
for k, v in d.items(): print("dist8", k, v["steps"], "hist_rel %.2e sol_rel %.2e" % (v["hist_rel"], v["sol_rel"]))
PY
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 200 --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  tail -2 gpurun_out/scale_n$N.err | cut -c1-300
  python - "$N" <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/scale_n{N}.json").read().strip().splitlines()[-1])
    print(f"N={N}: value {d['value']:.1f} it/s, ms/step {d['ms_per_step']:.3f}, spmv frac {d['roofline']['frac']:.3f}, step frac {d['roofline']['whole_step']['frac_of_aggregate_peak']:.3f}, e2e {d.get('e2e',{}).get('value')}, e2e steps {d.get('e2e',{}).get('numsteps')}, e2e secs {d.get('e2e',{}).get('solve_seconds')}, clocks {d['clocks']}")
except Exception as e:
    print("N=", N, "failed:", e)
PY
done
# N=8: peer-memory (push + fused all-reduce) vs NCCL (send/recv + ncclAllReduce) for the same kernels
for mode in nccl p2p; do
  KRYLOV_B200_ALLREDUCE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 --steps 300 --warmup 10 --no-e2e > gpurun_out/n8_$mode.json 2> gpurun_out/n8_$mode.err
  python - "$mode" <<'PY'
import json, sys
m = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/n8_{m}.json").read().strip().splitlines()[-1])
    print(f"N=8 {m}: {d['value']:.1f} it/s, ms/step {d['ms_per_step']:.4f}, spmv phase ms {d['roofline']['launch_ms']:.4f}, halo {d['config']['halo']}, allreduce {d['config']['allreduce']}")
except Exception as e:
    print("N=8", m, "failed:", e)
PY
done
