"""Plane-marching stencil kernel (kb_march.cuh) at N^3: SpMV(+dot) against the tiled stencil kernel,
and a CG step with the three-kernel path (kb_tune 15 = 0) against the fused two-launch path, over
tile/ring shapes (kb_tune 14) and planes per work item (kb_tune 13).  usage: march_bench.py N"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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

def spmv_ms(reps=20):
    for _ in range(3): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def tune(**kw):
    for k, v in kw.items(): lib.kb_tune({"st": 10, "ctas": 11, "l2": 12, "ch": 13, "mc": 14, "fuse": 15}[k], v)

moved = A.moved_bytes(1)
tune(st=11); ms = spmv_ms(); say(f"N={N} SpMV+dot tiled (stencil2)      : {ms:.4f} ms  {moved/ms/1e6:.0f} GB/s moved")
QUICK = len(sys.argv) > 2 and sys.argv[2] == "quick"
for mc in ((0, 2) if QUICK else (0, 1, 2, 3)):
    for ch in ((0, 64) if QUICK else (0, 16, 64)):
        for l2 in ((0,) if QUICK else (0, 1)):
            tune(st=10, mc=mc, ch=ch, l2=l2)
            ms = spmv_ms()
            say(f"N={N} SpMV+dot march mc={mc} ch={ch:2d} l2={l2}: {ms:.4f} ms  {moved/ms/1e6:.0f} GB/s moved")
tune(st=0, mc=0, ch=0, l2=0)

g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))

def cg_ms(tag):
    st = FusedCG(A, b.reshape(n, 1), torch.zeros(n, 1, dtype=torch.float64, device="cuda"), 0.0, 0.0)
    st.run(6); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h = st.run(30); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    say(f"CG step N={N} {tag}: {ms:.4f} ms = {1e3/ms:.1f} it/s  (fused={st.fused_march}, resnorm[36]={h[-1][0]:.15e})")
    del st

tune(fuse=0, st=11); cg_ms("three kernels, tiled SpMV  ")
tune(fuse=0, st=10); cg_ms("three kernels, march SpMV  ")
tune(st=0)
for mc in ((0, 2) if QUICK else (0, 2, 3, 1)):
    for ch in ((0, 64) if QUICK else (0, 16, 64)):
        for l2 in ((0, 1) if ch == 0 and not QUICK else (0,)):
            tune(fuse=1, mc=mc, ch=ch, l2=l2)
            cg_ms(f"fused march mc={mc} ch={ch:2d} l2={l2}")
tune(fuse=1, mc=0, ch=0, l2=0)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/march_bench_{N}.txt", "w").write("\n".join(lines) + "\n")
