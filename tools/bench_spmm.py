"""SpMM for blocked right-hand sides on the 3-D Poisson stencil: row-wise kernel vs the line-marching
kernel (csrc/kb_lines.cuh), and blocked CG (BASELINE config C4: k = 16, 256^3) with either.

    python tools/bench_spmm.py [--n 256] [--quick]
Bytes are the CSR model of SURVEY.md 8d (12 nnz + 4 (n + 1) + 16 n k per product)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_b200 as kb  # noqa: E402
from krylov_b200._lib import lib  # noqa: E402
from krylov_b200.device import Ops  # noqa: E402
from krylov_b200.generate import device_stencil7  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
peak = 6454.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
N = a.n
A = device_stencil7(N, N, N)
n = A.shape[0]
print(f"3-D Poisson {N}^3, peak {peak:.1f} GB/s")
for k in ((16,) if a.quick else (16, 8, 4, 32)):
    ops = Ops(n, k)
    x = torch.randn(n, k, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    out = ops.slots(1)[0]
    res = {}
    for name, cfg, ch, chunk, order in (("row-wise", 0, 0, 0, 1), ("lines planes/32", 1, 0, 0, 1),
                                        ("lines planes/64", 1, 64, 0, 1), ("lines planes/16", 1, 16, 0, 1),
                                        ("lines natural/32", 1, 0, 0, 0), ("lines 512 pl/32", 1, 0, 1, 1)):
        if a.quick and (ch or chunk or not order):
            continue
        lib.kb_tune(16, cfg)
        lib.kb_tune(17, ch)
        lib.kb_tune(18, chunk)
        lib.kb_tune(19, order)
        for _ in range(2):
            ops.spmv(A, x, y, dot=1, w=x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2 if a.quick else 10
        e0.record()
        for _ in range(reps):
            ops.spmv(A, x, y, dot=1, w=x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = (y.clone(), out.clone())
        print(f"k={k:2d} {name:16s}: {ms:7.3f} ms  model {A.spmv_bytes(k) / ms / 1e6:6.0f} GB/s = "
              f"{A.spmv_bytes(k) / ms / 1e6 / peak:5.3f} of peak   (x in + y out only: {16.0 * n * k / ms / 1e6:6.0f} GB/s)",
              flush=True)
    ref = res["row-wise"][0]
    print("      bit-identical to row-wise:", all(torch.equal(v[0], ref) for v in res.values()))
    del x, y, res, ref
lib.kb_tune(17, 0)
lib.kb_tune(18, 0)
lib.kb_tune(19, 1)
if not a.quick:
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn((n, 16), generator=g, dtype=torch.float64, device="cuda")
    for cfg, chunk, nm in ((0, 0, "row-wise SpMM"), (1, 0, "line-marching SpMM, 1024-entry chunks"),
                           (1, 1, "line-marching SpMM, 512-entry chunks")):
        lib.kb_tune(16, cfg)
        lib.kb_tune(18, chunk)
        kb.cg(A, B, tol=0.0, atol=0.0, maxiter=5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol, info = kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        by = info.numsteps * (12 * A.nnz + 4 * (n + 1) + 92 * n * 16)
        print(f"C4 blocked cg k=16 3D Poisson {N}^3, 50 fixed iterations [{nm}]: {info.numsteps} steps in "
              f"{secs * 1e3:.2f} ms = {info.numsteps / secs:.1f} it/s; model bytes {by / 1e9:.2f} GB -> "
              f"{by / secs / 1e9:.0f} GB/s = {100 * by / secs / 1e9 / peak:.1f}% of measured peak")
lib.kb_tune(16, 1)
lib.kb_tune(18, 0)
