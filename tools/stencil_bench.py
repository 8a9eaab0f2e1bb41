"""Constant-diagonal ("stencil") SpMV(+dot) sweep on the 7-point matrix: tile shapes (kb_tune 10)
x CTAs/SM caps (kb_tune 11), GB/s on the bytes the schedule moves, with the pattern / stream
schedules and a CG step beside it.  usage: stencil_bench.py N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.cg import FusedCG
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = device_stencil7(N, N, N)
n = A.shape[0]
ops = Ops(n, 1)
x = torch.randn(n, 1, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
out = ops.slots(1)
lines = []

def say(s):
    print(s, flush=True); lines.append(s)

def run(sched, tag, reps=20):
    A.set_schedule(sched)
    for _ in range(3): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    say(f"N={N} {sched:8s} {tag}: {ms:.4f} ms  moved {A.moved_bytes(1)/1e9:.2f} GB -> "
        f"{A.moved_bytes(1)/ms/1e6:.0f} GB/s (CSR model {A.spmv_bytes(1)/ms/1e6:.0f} GB/s)")
    return ms

say(f"schedule chosen: {A.info()['schedule']}")
best = (1e9, 0, 0)
for cfg in (0, 1, 2, 3, 4, 5, 6, 7, 8):
    for ctas in ((0, 3, 4) if cfg < 6 else (0,)):
        lib.kb_tune(10, cfg); lib.kb_tune(11, ctas)
        ms = run("stencil", f"cfg={cfg} ctas={ctas}")
        best = min(best, (ms, cfg, ctas))
say(f"best: cfg={best[1]} ctas={best[2]} {best[0]:.4f} ms")
lib.kb_tune(10, 0); lib.kb_tune(11, 0)
run("pattern", "default")
run("stream", "default")
# CG steps with each schedule
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
for sched, cfg, ctas in (("stencil", 0, 0), ("stencil", best[1], best[2]), ("pattern", 0, 0)):
    lib.kb_tune(10, cfg); lib.kb_tune(11, ctas)
    A.set_schedule(sched)
    st = FusedCG(A, b.reshape(n, 1), torch.zeros(n, 1, dtype=torch.float64, device="cuda"), 0.0, 0.0)
    st.run(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.run(30); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    say(f"CG step N={N} schedule={sched} cfg={cfg} ctas={ctas}: {ms:.4f} ms = {1e3/ms:.1f} it/s")
    del st
lib.kb_tune(10, 0); lib.kb_tune(11, 0)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/stencil_bench_{N}.txt", "w").write("\n".join(lines) + "\n")
