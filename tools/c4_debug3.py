import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7
def c4(tag):
    N = 256; A = device_stencil7(N, N, N); n = A.shape[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn((n, 16), generator=g, dtype=torch.float64, device="cuda")
    kb.cg(A, B, tol=0.0, atol=0.0, maxiter=20); torch.cuda.synchronize()
    t0 = time.perf_counter(); kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50); torch.cuda.synchronize()
    print(tag, "cg k=16 50 its:", round(time.perf_counter() - t0, 3), "s", flush=True)
    del A, B; torch.cuda.empty_cache()
c4("fresh")
N = 256; A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs()); n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
kb.gmres(A, b, tol=1e-8, maxiter=50); torch.cuda.synchronize()
del A, b; torch.cuda.empty_cache()
c4("after gmres mgs")
A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs())
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
kb.gmres(A, b, tol=1e-8, maxiter=50, ortho="householder"); torch.cuda.synchronize()
print("mem after householder:", torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9, flush=True)
del A, b; torch.cuda.empty_cache()
print("mem after empty_cache:", torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9, flush=True)
c4("after gmres householder")
