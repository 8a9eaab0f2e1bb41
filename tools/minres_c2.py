"""BASELINE C2 alone (MINRES, shifted 3-D Laplacian 128^3): whole solve timed, for an ncu launch list."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
A = device_stencil7(N, N, N, shift=st.mild_shift(N))
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(A.shape[0], generator=g, dtype=torch.float64, device="cuda"))
kb.minres(A, b, tol=1e-8, maxiter=20000)
torch.cuda.synchronize()
t0 = time.perf_counter()
sol, info = kb.minres(A, b, tol=1e-8, maxiter=20000)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"minres {N}^3: {info.numsteps} steps in {dt*1e3:.2f} ms = {info.numsteps/dt:.0f} it/s ({dt/info.numsteps*1e6:.1f} us/step)")
