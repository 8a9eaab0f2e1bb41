"""One short-recurrence solver alone (fixed steps) for an ncu launch list.  usage: shortrec_one.py NAME [N] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7
name = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs())
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(A.shape[0], generator=g, dtype=torch.float64, device="cuda"))
fn = getattr(kb, name)
fn(A, b, tol=0.0, atol=0.0, maxiter=3)
torch.cuda.synchronize()
t0 = time.perf_counter()
_, info = fn(A, b, tol=0.0, atol=0.0, maxiter=steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{name} {N}^3: {info.numsteps} steps in {dt*1e3:.1f} ms = {dt/info.numsteps*1e6:.0f} us/step")
