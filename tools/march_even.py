"""A/B of the fused marching CG step at N^3: equal-items grid (kb_tune 20) on / off, alternating
in one process after a long warm-up (the boxes drift by >10 % with temperature / power capping, so
only interleaved medians compare).  usage: march_even.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from krylov_b200._lib import lib
from krylov_b200.cg import FusedCG
from krylov_b200.generate import device_stencil7

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
A = device_stencil7(N, N, N)
n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda")).reshape(n, 1)
x0 = torch.zeros_like(b)
st = FusedCG(A, b, x0, 0.0, 0.0)
st.run(200)  # warm-up: reach the steady thermal / power state
res = {0: [], 1: []}
for rep in range(10):
    for even in (0, 1):
        lib.kb_tune(20, even)
        ph, tot, fused = st.run_timed(20)
        res[even].append((tot / 20, ph[0], ph[1]))
for even in (0, 1):
    a = np.array(res[even])
    print(f"N={N} even={even}: step median {np.median(a[:,0]):.4f} min {a[:,0].min():.4f} ms; "
          f"KIND1 median {np.median(a[:,1]):.4f}; KIND2 median {np.median(a[:,2]):.4f}", flush=True)
