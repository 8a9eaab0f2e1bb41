"""First-contact check on a B200: SpMV schedules vs SciPy (bit-exact), a CG
parity run against the oracle, and a rough SpMV/CG timing."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops
from oracle import krylov_oracle as orc

print(torch.cuda.get_device_name(0), flush=True)
os.system("free -g | head -2; nproc")
rng = np.random.default_rng(0)

def check_spmv(Asp, k, sched):
    A = kb.CsrMatrix.from_scipy(Asp).set_schedule(sched)
    x = rng.standard_normal((Asp.shape[1], k)) if k > 1 else rng.standard_normal(Asp.shape[1])
    y = A @ x
    ref = Asp @ x
    d = np.max(np.abs(y - ref))
    print(f"spmv n={Asp.shape[0]} k={k} sched={A.info()['schedule']}: max|diff|={d:.3e} bitexact={np.array_equal(y, ref)}", flush=True)

for n in (5, 17, 40):
    Asp = st.poisson3d(n)
    for sched in ("rowwise", "stream"):
        check_spmv(Asp, 1, sched)
    check_spmv(Asp, 4, "auto"); check_spmv(Asp, 3, "auto"); check_spmv(Asp, 16, "auto")
import scipy.sparse
Ar = scipy.sparse.random(3000, 3000, density=0.02, random_state=1, format="csr")
for sched in ("rowwise", "stream"):
    check_spmv(Ar, 1, sched)
Ar2 = scipy.sparse.random(700, 700, density=0.5, random_state=2, format="csr")  # long rows: several chunks/tile
for sched in ("rowwise", "stream"):
    check_spmv(Ar2, 1, sched)

# device generator == host generator
Ad = device_stencil7(9, 7, 5)
Ah = st.to_scipy(st.stencil7_csr(9, 7, 5))
print("generator equal:", (Ad.to_scipy() != Ah).nnz == 0, flush=True)

# CG parity on 2-D Poisson 64^2
Asp = st.poisson2d(64)
xs = rng.standard_normal(Asp.shape[0]); b = Asp @ xs
sol, info = kb.cg(Asp, b, tol=1e-10, maxiter=5000)
sol_o, info_o = orc.cg(Asp, b, tol=1e-10, maxiter=5000)
print("cg steps", info.numsteps, info_o.numsteps, "success", info.success)
ro, rg = np.array(info_o.resnorms), np.array(info.resnorms)
m = min(len(ro), len(rg))
print("max rel resnorm diff", np.max(np.abs(ro[:m] - rg[:m]) / ro[:m]), "sol relerr", np.linalg.norm(sol - sol_o) / np.linalg.norm(sol_o), flush=True)
Bk = Asp @ rng.standard_normal((Asp.shape[0], 4))
sol, info = kb.cg(Asp, Bk, tol=1e-9, maxiter=5000)
sol_o, info_o = orc.cg(Asp, Bk, tol=1e-9, maxiter=5000)
print("cg k=4 steps", info.numsteps, info_o.numsteps, "relerr", np.linalg.norm(sol - sol_o) / np.linalg.norm(sol_o), flush=True)

# timing: SpMV at 256^3 both schedules, then CG iterations
def time_spmv(A, sched, reps=20):
    A.set_schedule(sched)
    n = A.shape[0]
    ops = Ops(n, 1)
    x = torch.randn(n, 1, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    out = ops.slots(1)
    for _ in range(3): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = A.spmv_bytes(1) / 1e9
    print(f"spmv+dot {sched} n={n}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s", flush=True)

for N in (256, 400):
    A = device_stencil7(N, N, N)
    print(A.info(), flush=True)
    for sched in ("rowwise", "stream"):
        time_spmv(A, sched)
    A.set_schedule("auto")
    n = A.shape[0]
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize(); t0 = time.time()
    sol, info = kb.cg(A, b, tol=0.0, atol=0.0, maxiter=200)
    torch.cuda.synchronize(); dt = time.time() - t0
    byt = 12 * A.nnz + 4 * (n + 1) + 92 * n
    print(f"cg {N}^3 200 its: {dt:.3f}s  {200/dt:.1f} it/s  {byt*200/dt/1e9:.0f} GB/s (model bytes)", flush=True)
    del A, b, sol
    torch.cuda.empty_cache()
