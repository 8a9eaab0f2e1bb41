"""Per-rank slab shapes of the 512^3 strong-scaling run on ONE GPU (no communication): fused
marching CG step on 512 x 512 x nz for nz = 64 / 128 / 256 (the slabs of 8 / 4 / 2 ranks), over
planes per work item (kb_tune 13) and the equal-items grid (kb_tune 20).  Interleaved medians.
usage: slab_tune.py [nz ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from krylov_b200._lib import lib
from krylov_b200.cg import FusedCG
from krylov_b200.generate import device_stencil7

PEAK = 6454.6
for nz in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    A = device_stencil7(512, 512, nz)
    n = A.shape[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda")).reshape(n, 1)
    x0 = torch.zeros_like(b)
    st = FusedCG(A, b, x0, 0.0, 0.0)
    st.run(100)
    cfgs = [(0, 0), (0, 1), (nz, 0), (nz // 2, 0), (nz // 2, 1), (nz // 3 + 1, 1), (nz // 4, 1), (8, 1)]
    res = {c: [] for c in cfgs}
    for rep in range(6):
        for c in cfgs:
            lib.kb_tune(13, c[0]); lib.kb_tune(20, c[1])
            ph, tot, fused = st.run_timed(20)
            res[c].append((tot / 20, ph[0], ph[1]))
    for c in cfgs:
        a = np.array(res[c])
        m = np.median(a, axis=0)
        print(f"nz={nz:3d} ch={c[0]:3d} even={c[1]}: step {m[0]*1e3:7.1f} us  KIND1 {m[1]*1e3:7.1f} us "
              f"({42*n/m[1]/1e6/PEAK:.2f} of peak)  KIND2 {m[2]*1e3:7.1f} us ({26*n/m[2]/1e6/PEAK:.2f})", flush=True)
    lib.kb_tune(13, 0); lib.kb_tune(20, 0)
    del st, A, b, x0
