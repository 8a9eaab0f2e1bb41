"""C3 (GMRES(50) cycle on conv-diff 256^3): modified vs classical Gram-Schmidt (extension
ortho="cgs"/"cgs2"), with the bytes-moved model of each and the multi-dot chunk width."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs()); n = A.shape[0]
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
spmv = 12 * A.nnz + 4 * (n + 1) + 16 * n
out = []

def run(ortho, jc=8):
    lib.kb_tune(9, jc)
    f = lambda: kb.gmres(A, b, tol=1e-8, maxiter=50, ortho=ortho)
    f(); f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sol, info = f(); e1.record(); torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) / 1e3
    r = 1 if len(ortho) == 3 else int(ortho[3:])
    if ortho.startswith("mgs"):
        byt = sum(spmv + 16 * n + r * 32 * n * (j + 1) for j in range(50))
    else:
        byt = sum(spmv + 16 * n + r * 8 * n * ((j + 1) + -(-(j + 1) // jc) + (j + 1) + 2)
                  for j in range(50))
    res = float(info.resnorms[-1]) / float(info.resnorms[0])
    line = (f"gmres conv-diff {N}^3 50-step cycle ortho={ortho} jc={jc}: {secs*1e3:.1f} ms "
            f"({50/secs:.1f} steps/s); model {byt/1e9:.1f} GB -> {byt/secs/1e9:.0f} GB/s = "
            f"{100*byt/secs/1e9/PEAK:.1f}% of peak; rel resnorm after cycle {res:.6e}")
    print(line, flush=True); out.append(line)

for o, jc in (("mgs", 8), ("cgs", 8), ("cgs", 16), ("mgs2", 8), ("cgs2", 8), ("cgs2", 16)):
    run(o, jc)
lib.kb_tune(9, 8)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "cgs_bench.txt"), "w").write("\n".join(out) + "\n")
