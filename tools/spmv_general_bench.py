"""SpMV on NON-stencil matrices: the CSR-streaming schedules (merge / stream / rowwise) and
cuSPARSE (through torch) on the same device arrays.  GB/s by the CSR byte model
12 nnz + 4 (n+1) + 16 n.
usage: spmv_general_bench.py [quick] [--only NAME] [--one NAME SCHED CFG REPS]   (--one: for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import numpy as np
import torch
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.device import Ops

dev = torch.device("cuda")
quick = "quick" in sys.argv
PEAK = 6454.6
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def pad(t, dtype):
    n = t.numel()
    out = torch.zeros(((n + 3) // 4) * 4 + 4, dtype=dtype, device=dev)
    out[:n] = t
    return out


def device_csr(rowptr, cols, vals, shape):
    nnz = cols.numel()
    return kb.CsrMatrix._from_device_arrays(rowptr.to(torch.int32).contiguous(), pad(cols, torch.int32),
                                            pad(vals, torch.float64), nnz, shape)


def random_rows(n, per_row, band=None, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    if band is None:
        cols = torch.randint(0, n, (n, per_row), device=dev, generator=g, dtype=torch.int32)
    else:
        r = torch.arange(n, device=dev, dtype=torch.int32).unsqueeze(1)
        cols = (r + torch.randint(-band, band + 1, (n, per_row), device=dev, generator=g, dtype=torch.int32)).clamp_(0, n - 1)
    cols = cols.sort(dim=1).values.reshape(-1)
    vals = torch.randn(n * per_row, device=dev, generator=g, dtype=torch.float64)
    rowptr = torch.arange(0, n * per_row + 1, per_row, device=dev, dtype=torch.int64)
    return device_csr(rowptr, cols, vals, (n, n))


def powerlaw(n, scale=6.0, alpha=1.0, cap=2_000_000, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand(n, device=dev, generator=g, dtype=torch.float64)
    lens = (scale * (u ** (-1.0 / alpha) - 1.0)).to(torch.int64).clamp_(0, cap)  # Pareto II, some empty rows
    rowptr = torch.zeros(n + 1, device=dev, dtype=torch.int64)
    rowptr[1:] = lens.cumsum(0)
    nnz = int(rowptr[-1])
    assert nnz < 2 ** 31
    cols = torch.randint(0, n, (nnz,), device=dev, generator=g, dtype=torch.int32)
    vals = torch.randn(nnz, device=dev, generator=g, dtype=torch.float64)
    return device_csr(rowptr, cols, vals, (n, n)), int(lens.max())


def fem27(m, seed=0):
    """27-point pattern with variable coefficients on an m^3 grid (rows in grid order)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    n = m ** 3
    i = torch.arange(n, device=dev, dtype=torch.int64)
    ix, iy, iz = i % m, (i // m) % m, i // (m * m)
    cols, ok = [], []
    for c in (-1, 0, 1):
        for b in (-1, 0, 1):
            for a in (-1, 0, 1):
                ok.append((ix + a >= 0) & (ix + a < m) & (iy + b >= 0) & (iy + b < m) & (iz + c >= 0) & (iz + c < m))
                cols.append((i + a + b * m + c * m * m).to(torch.int32))
    ok = torch.stack(ok, 1)
    cols = torch.stack(cols, 1)
    rowptr = torch.zeros(n + 1, device=dev, dtype=torch.int64)
    rowptr[1:] = ok.sum(1).cumsum(0)
    cols = cols[ok]
    vals = -torch.rand(cols.numel(), device=dev, generator=g, dtype=torch.float64) - 0.1
    return device_csr(rowptr, cols, vals, (n, n))


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench(name, A, extra=""):
    n = A.shape[0]
    ops = Ops(n, 1)
    x = torch.randn(n, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    out = ops.slots(1)
    nbytes = A.spmv_bytes(1)
    info = A.info()
    print(f"== {name}: n={n} nnz={A.nnz} mean={A.nnz / n:.1f} max_row={info['max_row_len']} auto={info['schedule']} {extra}; "
          f"CSR bytes {nbytes / 1e9:.3f} GB", flush=True)
    reps = 5 if quick else 20
    S = torch.sparse_csr_tensor(A.rowptr.to(torch.int32), A.colidx[: A.nnz], A.vals[: A.nnz], size=A.shape)
    ref = (S @ x.unsqueeze(1)).squeeze(1)
    scale = torch.sparse_csr_tensor(A.rowptr.to(torch.int32), A.colidx[: A.nnz], A.vals[: A.nnz].abs(), size=A.shape) @ x.abs().unsqueeze(1)
    ms = timeit(lambda: S @ x.unsqueeze(1), reps)
    print(f"   cuSPARSE (torch CSR @ x)      : {ms:8.4f} ms {nbytes / ms / 1e6:7.0f} GB/s  {nbytes / ms / 1e6 / PEAK:5.2f} of peak", flush=True)
    runs = [("merge", c, o) for c in (0, 2, 4, 6, 7) for o in (1, 2)] + [("stream", 0, 1), ("rowwise", 0, 1)]
    for sched, cfg, order in runs:
        if sched != "merge" and info["max_row_len"] > 100000:
            continue  # a single thread walking a 10^5-entry row: minutes
        try:
            A.set_schedule(sched)
        except kb.KrylovB200Error as e:
            print(f"   {sched}: refused ({e})")
            continue
        lib.kb_tune(25, cfg)
        lib.kb_tune(27, order)
        ms = timeit(lambda: ops.spmv(A, x, y, dot=1, w=x, out=out[0]), reps)
        err = float(((y - ref).abs() / (scale.squeeze(1) + 1e-300)).max())
        print(f"   {sched:8s} cfg={cfg} order={order} (+ <x, Ax>): {ms:8.4f} ms {nbytes / ms / 1e6:7.0f} GB/s  {nbytes / ms / 1e6 / PEAK:5.2f} of peak   max err/|A||x| vs cuSPARSE {err:.1e}", flush=True)
    lib.kb_tune(25, 0)
    lib.kb_tune(27, 2)
    A.set_schedule("auto")


def make(name):
    big = not quick
    if name == "random_100":
        return random_rows(2_000_000 if big else 400_000, 100), ""
    if name == "banded_100":
        return random_rows(2_000_000 if big else 400_000, 100, band=4096), "columns within +-4096 of the row"
    if name == "powerlaw":
        A, mx = powerlaw(6_000_000 if big else 1_000_000)
        return A, f"Pareto row lengths, longest {mx}"
    if name == "fem27_var":
        return fem27(192 if big else 96), "27-point pattern, variable coefficients"
    raise KeyError(name)


if "--import" in sys.argv:  # bench.py uses the generators
    pass
elif "--one" in sys.argv:
    i = sys.argv.index("--one")
    name, sched, cfg, reps = sys.argv[i + 1], sys.argv[i + 2], int(sys.argv[i + 3]), int(sys.argv[i + 4])
    A, _ = make(name)
    A.set_schedule(sched)
    lib.kb_tune(25, cfg % 10)
    lib.kb_tune(27, cfg // 10)  # order * 10 + cfg
    ops = Ops(A.shape[0], 1)
    x = torch.randn(A.shape[0], dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    out = ops.slots(1)
    for _ in range(reps):
        ops.spmv(A, x, y, dot=1, w=x, out=out[0])
    torch.cuda.synchronize()
    sys.exit(0)

names = ["random_100", "banded_100", "powerlaw", "fem27_var"]
if "--only" in sys.argv:
    names = [sys.argv[sys.argv.index("--only") + 1]]
for nm in ([] if "--import" in sys.argv else names):
    A, extra = make(nm)
    bench(nm, A, extra)
    del A
    torch.cuda.empty_cache()
