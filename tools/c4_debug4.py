import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7
from krylov_b200.cg import FusedCG
def ev(): return torch.cuda.Event(enable_timing=True)
def per_kernel(tag, k):
    N = 256; A = device_stencil7(N, N, N); n = A.shape[0]
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn((n, k), generator=g, dtype=torch.float64, device="cuda")
    st_ = FusedCG(A, B, torch.zeros_like(B), 0.0, 0.0); ops, sl = st_.ops, st_.sl
    hp = st_.hist.data_ptr()
    for i in range(3): st_.enqueue(i, hp - (i + 1) * k * 8)
    torch.cuda.synchronize()
    tot = [0.0] * 4
    for i in range(3, 8):
        cur, nxt = sl[i % 2], sl[(i + 1) % 2]
        es = [ev() for _ in range(5)]
        es[0].record(); ops.cg_update_p(cur, nxt, st_.r, st_.p, x=st_.yk, alpha=sl[2])
        es[1].record(); ops.spmv(A, st_.p, st_.Ap, dot=1, w=st_.p, out=sl[3])
        es[2].record(); ops.cg_update_xr(cur, sl[3], None, None, st_.Ap, None, st_.r, sl[4], alpha_out=sl[2])
        es[3].record(); ops.cg_record(i + 1, sl[4], st_.crit_d, hp - (i + 1) * k * 8, st_.stop_at, rho_keep=nxt)
        es[4].record(); torch.cuda.synchronize()
        for j in range(4): tot[j] += es[j].elapsed_time(es[j + 1]) / 5
    a = torch.empty(1 << 27, dtype=torch.float64, device="cuda"); b2 = torch.empty_like(a)
    b2.copy_(a); torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record(); b2.copy_(a); e1.record(); torch.cuda.synchronize()
    print(tag, f"k={k}", {nm: round(t, 3) for nm, t in zip(["upd_p", "spmm", "upd_r", "rec"], tot)},
          "torch copy GB/s", round(2 * a.numel() * 8 / e0.elapsed_time(e1) / 1e6), flush=True)
    del A, B, st_, a, b2; torch.cuda.empty_cache()
per_kernel("fresh", 16); per_kernel("fresh", 1)
N = 256; A = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs()); n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, generator=g, dtype=torch.float64, device="cuda"))
kb.gmres(A, b, tol=1e-8, maxiter=int(sys.argv[1]) if len(sys.argv) > 1 else 50, ortho="householder"); torch.cuda.synchronize()
del A, b; torch.cuda.empty_cache()
per_kernel("after householder", 16); per_kernel("after householder", 1)
os.system("nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv,noheader")
