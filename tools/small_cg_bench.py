"""Per-iteration device time of CG on small (L2-resident) problems: the persistent one-launch
kernel (csrc/kb_small.cu) against the launched path (kb_tune 28 = 0), CUDA events around
batches of 256 iterations with the stopping test disabled (tol = 0).
usage: small_cg_bench.py [n2d ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200._lib import lib
from krylov_b200.cg import FusedCG

sizes = [int(a) for a in sys.argv[1:]] or [256]
for m in sizes:
    A = st.poisson2d(m)
    n = A.shape[0]
    b = torch.from_numpy((A @ np.random.default_rng(0).standard_normal(n)).reshape(-1, 1)).cuda()
    for name, key in (("persistent", 262144), ("launched", 0)):
        lib.kb_tune(28, key)
        Ad = kb.CsrMatrix.from_scipy(A)
        s = FusedCG(Ad, b, torch.zeros_like(b), 0.0, 0.0)
        s.run(64)
        torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            s.run(256)
            e1.record()
            torch.cuda.synchronize()
            ts.append((e0.elapsed_time(e1) * 1e3 / 256, (time.perf_counter() - t0) * 1e6 / 256))
        dev_us = min(t[0] for t in ts)
        wall_us = min(t[1] for t in ts)
        print(f"2-D Poisson {m}^2 (n={n}) {name:10s}: {dev_us:6.2f} us/step on the device ({1e6 / dev_us:9.0f} it/s), "
              f"{wall_us:6.2f} us/step wall incl. the batch read-back; persistent={s.persistent}", flush=True)
    lib.kb_tune(28, 262144)
    x, info = kb.cg(kb.CsrMatrix.from_scipy(A), b.cpu().numpy().ravel(), tol=1e-10, maxiter=5000)
    t0 = time.perf_counter()
    x, info = kb.cg(kb.CsrMatrix.from_scipy(A), b.cpu().numpy().ravel(), tol=1e-10, maxiter=5000)
    dt = time.perf_counter() - t0
    print(f"   whole solve kb.cg(tol=1e-10): {info.numsteps} steps in {dt * 1e3:.2f} ms = {info.numsteps / dt:.0f} it/s (host set-up, batches of 8..256, confirmation included)")
