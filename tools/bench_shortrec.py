"""Throughput of the SURVEY.md 8(f).2 solvers on one B200 (general device path: one kernel per
vector statement, scalars on the host): iterations/s for a fixed number of steps, kernel launches
per iteration and the bandwidth on the bytes those launches stream (per-launch byte models of
DESIGN.md section 3, counted by wrapping Ops).  usage: bench_shortrec.py [N]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krylov_b200 as kb
from krylov_b200.device import Ops
from krylov_b200.generate import device_stencil7
from krylov_b200 import stencils as st

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
BYTES = [0]

def wrap(name, per_elem):
    f = getattr(Ops, name)
    def g(self, *a, **kw):
        BYTES[0] += per_elem * self.n * self.k
        return f(self, *a, **kw)
    setattr(Ops, name, g)

for nm, pe in (("dot", 16), ("axpy", 24), ("xpby", 24), ("lincomb", 24), ("div_scale", 16), ("add", 24)):
    wrap(nm, pe)
_spmv = Ops.spmv
def spmv(self, A, x, y, **kw):
    BYTES[0] += A.moved_bytes(self.k)
    return _spmv(self, A, x, y, **kw)
Ops.spmv = spmv
from krylov_b200.csr import CsrMatrix
_mv = CsrMatrix.matvec_device
def mv(self, x, out=None):
    BYTES[0] += self.moved_bytes(1 if x.dim() == 1 else x.shape[1])
    return _mv(self, x, out)
CsrMatrix.matvec_device = mv

lines = []
def say(s):
    print(s, flush=True); lines.append(s)

def run(name, A, b, steps, **kw):
    fn = getattr(kb, name)
    fn(A, b, tol=0.0, atol=0.0, maxiter=3, **kw)  # warm-up (allocator, transposed matrix)
    torch.cuda.synchronize(); BYTES[0] = 0
    t0 = time.perf_counter()
    _, info = fn(A, b, tol=0.0, atol=0.0, maxiter=steps, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = BYTES[0] / dt / 1e9
    say(f"{name:10s} {N}^3: {info.numsteps} steps in {dt*1e3:8.1f} ms = {info.numsteps/dt:7.1f} it/s; "
        f"streamed {BYTES[0]/1e9/max(info.numsteps,1):6.2f} GB/step -> {gbs:5.0f} GB/s = {100*gbs/PEAK:4.1f}% of measured peak")

g = torch.Generator(device="cuda").manual_seed(0)
C = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs())
n = C.shape[0]
b = C.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
for name in ("bicgstab", "cgs", "bicg", "qmr"):
    run(name, C, b, 40)
run("gcr", C, b, 20)
del C
P = device_stencil7(N, N, N)
b = P.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
run("cgr", P, b, 40)
lam = lambda i: 4.0 * np.sin(i * np.pi / (2.0 * (N + 1))) ** 2
run("chebyshev", P, b, 40, eigenvalue_estimates=(3 * lam(1), 3 * lam(N)))
run("symmlq", P, b, 40)
os.makedirs("gpurun_out", exist_ok=True)
open(f"gpurun_out/shortrec_bench_{N}.txt", "w").write("\n".join(lines) + "\n")
