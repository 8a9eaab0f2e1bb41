"""SpMV(+dot) micro-benchmark on the 7-point stencil: sweeps kernel
configurations (kb_tune) and prints GB/s against the CSR byte model.
usage: spmv_bench.py N [--one SCHED CFG CTAS REPS]  (the --one form is for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
A = device_stencil7(N, N, N)
n = A.shape[0]
ops = Ops(n, 1)
x = torch.randn(n, 1, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
out = ops.slots(1)

def run(sched, cfg, ctas, reps=20, vec_ctas=4):
    A.set_schedule(sched)
    lib.kb_tune(0, cfg); lib.kb_tune(1, ctas); lib.kb_tune(2, vec_ctas)
    for _ in range(3): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"N={N} {sched:8s} cfg={cfg} ctas={ctas} vec_ctas={vec_ctas}: {ms:.4f} ms  {A.spmv_bytes(1)/ms/1e6:.0f} GB/s", flush=True)

if "--one" in sys.argv:
    i = sys.argv.index("--one")
    run(sys.argv[i+1], int(sys.argv[i+2]), int(sys.argv[i+3]), int(sys.argv[i+4]))
    sys.exit(0)

for wc, pcs in ((0, (0, 3)), (1, (0, 2)), (2, (0, 1)), (3, (0, 6, 5)), (4, (0, 4)), (5, (0,)), (-1, (0,))):
    for pc in pcs:
        lib.kb_tune(4, wc); lib.kb_tune(3, pc)
        print(f"window cfg {wc} ctas {pc}: ", end="")
        run("pattern", 0, 0, vec_ctas=8)
lib.kb_tune(3, 0); lib.kb_tune(4, 0)
for tb in (0, -1, 16, 32, 64, 128, 256):
    lib.kb_tune(8, tb)
    print(f"tile block {tb}: ", end="")
    run("pattern", 0, 0, vec_ctas=8)
lib.kb_tune(8, 0)
for vc in (4, 8):
    run("rowwise", 0, 0, vec_ctas=vc)
for cfg, ctas_list in ((0, (2,)), (5, (8,))):
    for ctas in ctas_list:
        run("stream", cfg, ctas)
# copy bandwidth reference (same method as MEASURED_PEAKS.json)
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): b.copy_(a)
e1.record(); torch.cuda.synchronize()
print(f"torch copy 2 GiB: {2 * a.numel() * 8 / (e0.elapsed_time(e1) / 10) / 1e6:.0f} GB/s")
