"""SpMM for blocked right-hand sides on a 7-point 3-D pattern with VARIABLE coefficients (schedule
"pattern": values streamed): row-wise kernel vs the line-marching kernel with the values in its
ring (kb_spmm_lines_kernel<VAR>), and blocked CG k = 16 with either.
    python tools/bench_spmm_var.py [--n 256]
Bytes: CSR model of SURVEY.md 8d (12 nnz + 4 (n+1) + 16 n k per product)."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.device import Ops
from krylov_b200.generate import device_stencil7

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
a = ap.parse_args()
peak = 6454.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
N = a.n
A0 = device_stencil7(N, N, N)
# same pattern, variable coefficients: SPD (diagonal 6 + eps_i, off-diagonals -1 scaled symmetric is
# not needed for the product bench; for CG keep the Poisson values and perturb the diagonal only)
vals = A0.vals.clone()
g = torch.Generator(device="cuda").manual_seed(0)
diag = (A0.colidx[: A0.nnz].long() == torch.repeat_interleave(
    torch.arange(A0.shape[0], device="cuda"), (A0.rowptr[1:] - A0.rowptr[:-1]).long()))
vals[: A0.nnz][diag] += 0.5 * torch.rand(int(diag.sum()), generator=g, dtype=torch.float64, device="cuda")
A = kb.CsrMatrix._from_device_arrays(A0.rowptr, A0.colidx, vals, A0.nnz, A0.shape)
n = A.shape[0]
print(f"7-point pattern {N}^3, variable diagonal, schedule {A.info()['schedule']}, peak {peak:.1f} GB/s")
for k in (16, 8, 32):
    ops = Ops(n, k)
    x = torch.randn(n, k, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    out = ops.slots(1)[0]
    res = {}
    for name, cfg, chunk in (("row-wise / windowed", 0, 0), ("lines, 1024-entry chunks", 1, 2),
                             ("lines, 512-entry chunks", 1, 1)):
        lib.kb_tune(16, cfg)
        lib.kb_tune(18, chunk)
        for _ in range(2):
            ops.spmv(A, x, y, dot=1, w=x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.spmv(A, x, y, dot=1, w=x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = y.clone()
        print(f"k={k:2d} {name:28s}: {ms:7.3f} ms  model {A.spmv_bytes(k) / ms / 1e6:6.0f} GB/s = "
              f"{A.spmv_bytes(k) / ms / 1e6 / peak:.2f} of peak", flush=True)
    v = list(res.values())
    lib.kb_tune(18, 0)
    print(f"      bit-identical: {bool(torch.equal(v[0], v[1]) and torch.equal(v[0], v[2]))}")
    del x, y
B = torch.randn(n, 16, dtype=torch.float64, device="cuda")
for name, cfg in (("row-wise / windowed", 0), ("lines (values in the ring, default chunks)", 1)):
    lib.kb_tune(16, cfg)
    kb.cg(A, B, tol=0.0, atol=0.0, maxiter=5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kb.cg(A, B, tol=0.0, atol=0.0, maxiter=30)
    e1.record()
    torch.cuda.synchronize()
    s = e0.elapsed_time(e1) / 1e3
    byt = 30 * (12 * A.nnz + 4 * (n + 1) + 92 * n * 16)
    print(f"blocked CG k=16, 30 iterations, {name:28s}: {30 / s:6.1f} it/s, model {byt / s / 1e9:6.0f} GB/s = {byt / s / 1e9 / peak:.2f} of peak")
lib.kb_tune(16, 1)
