import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200.generate import device_stencil7
N = 256; A = device_stencil7(N, N, N); n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
B = torch.randn((n, 16), generator=g, dtype=torch.float64, device="cuda")
kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50); torch.cuda.synchronize()
t0 = time.perf_counter(); kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50); torch.cuda.synchronize(); print("cg 50 its:", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
