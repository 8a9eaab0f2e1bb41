import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7
N = 256; A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs()); n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
kb.gmres(A, b, tol=1e-8, maxiter=50, ortho="householder"); torch.cuda.synchronize()
del A, b; torch.cuda.empty_cache()
A = device_stencil7(N, N, N)
B = torch.randn((n, 16), generator=g, dtype=torch.float64, device="cuda")
kb.cg(A, B, tol=0.0, atol=0.0, maxiter=20); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
sol, info = kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50); torch.cuda.synchronize(); print("cg:", time.perf_counter() - t0); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(8)
import numpy as np
print("finite:", np.isfinite(np.asarray(info.resnorms)).all(), "last resnorm", np.asarray(info.resnorms)[-1][:3])
