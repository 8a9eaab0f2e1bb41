"""Throughput of the other BASELINE.json configurations (C1-C4) on one B200:
iterations/s and fraction of the bytes-moved roofline (SURVEY.md 8d byte models,
peak = MEASURED_PEAKS.json hbm_gbs).  Inputs are device-resident; timing with
CUDA events around the public API call after one warm-up call."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
out = []

def timed(fn, reps=1):
    fn()  # warm-up (also lets the caching allocator settle: no cudaMalloc in the timed call)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps, r

def report(name, steps, secs, bytes_total, extra=""):
    gbs = bytes_total / secs / 1e9
    line = f"{name}: {steps} steps in {secs*1e3:.2f} ms = {steps/secs:.1f} it/s; model bytes {bytes_total/1e9:.2f} GB -> {gbs:.0f} GB/s = {100*gbs/PEAK:.1f}% of measured peak {extra}"
    print(line, flush=True); out.append(line)

g = torch.Generator(device="cuda").manual_seed(0)
# C1: cg, 2-D Poisson 256^2 (L2-resident: 4 MB matrix) -- latency-bound, it/s only
A = kb.CsrMatrix.from_scipy(st.poisson2d(256)); n = A.shape[0]
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
secs, (sol, info) = timed(lambda: kb.cg(A, b, tol=1e-10, maxiter=5000), reps=3)
report("C1 cg 2D Poisson 256^2 tol 1e-10", info.numsteps, secs, info.numsteps * (12*A.nnz + 4*(n+1) + 92*n), "(L2-resident; one persistent launch per batch, two grid barriers per step)")
# C2: minres, shifted 3-D Laplacian 128^3, mild shift
N = 128; A = device_stencil7(N, N, N, shift=st.mild_shift(N)); n = A.shape[0]
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
secs, (sol, info) = timed(lambda: kb.minres(A, b, tol=1e-8, maxiter=20000))
report("C2 minres shifted Laplacian 128^3 tol 1e-8", info.numsteps, secs, info.numsteps * (12*A.nnz + 4*(n+1) + 112*n))
# C3: gmres, conv-diff 256^3, one 50-step cycle per ortho
N = 256; A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs()); n = A.shape[0]
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
for ortho, r in (("mgs", 1), ("mgs2", 2)):
    secs, (sol, info) = timed(lambda: kb.gmres(A, b, tol=1e-8, maxiter=50, ortho=ortho))
    byt = sum(12*A.nnz + 4*(n+1) + 32*n + r*32*n*(j+1) for j in range(50))
    report(f"C3 gmres conv-diff 256^3 one 50-step cycle ortho={ortho}", info.numsteps, secs, byt)
secs, (sol, info) = timed(lambda: kb.gmres(A, b, tol=1e-8, maxiter=50, ortho="householder"))
byt = sum(12*A.nnz + 4*(n+1) + 40*n + 32*n*(2*j+3) for j in range(50))
report("C3 gmres conv-diff 256^3 one 50-step cycle ortho=householder", info.numsteps, secs, byt)
del A, b, sol, info
# C4: blocked cg k=16, 3-D Poisson 256^3, 50 fixed iterations
N = 256; A = device_stencil7(N, N, N); n = A.shape[0]
B = torch.randn((n, 16), generator=g, dtype=torch.float64, device="cuda")
from krylov_b200._lib import lib
for cfg, nm in ((-1, "row-wise SpMM"), (0, "windowed SpMM RPT=2"), (1, "windowed SpMM RPT=4")):
    lib.kb_tune(7, cfg)
    secs, (sol, info) = timed(lambda: kb.cg(A, B, tol=0.0, atol=0.0, maxiter=50))
    report(f"C4 blocked cg k=16 3D Poisson 256^3, 50 fixed iterations [{nm}]", info.numsteps, secs, info.numsteps * (12*A.nnz + 4*(n+1) + 92*n*16))
lib.kb_tune(7, 0)
open(os.path.join(ROOT, "gpurun_out", "configs.txt"), "w").write("\n".join(out) + "\n")
