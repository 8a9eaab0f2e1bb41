import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7
from krylov_b200.cg import FusedCG
N = 256; A = device_stencil7(N, N, N); n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(0)
B = torch.randn((n, 16), generator=g, dtype=torch.float64, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
for spmm in (0, -1):
    lib.kb_tune(7, spmm)
    st = FusedCG(A, B, torch.zeros_like(B), 0.0, 0.0)
    ops, sl = st.ops, st.sl
    hp = st.hist.data_ptr()
    for i in range(3): st.enqueue(i, hp - (i + 1) * 16 * 8)
    torch.cuda.synchronize()
    names = ["update_p(x+p)", "spmm+dot", "update_r", "record"]
    tot = [0.0] * 4
    for i in range(3, 8):
        cur, nxt = sl[i % 2], sl[(i + 1) % 2]
        es = [ev() for _ in range(5)]
        es[0].record(); ops.cg_update_p(cur, nxt, st.r, st.p, x=st.yk, alpha=sl[2])
        es[1].record(); ops.spmv(A, st.p, st.Ap, dot=1, w=st.p, out=sl[3])
        es[2].record(); ops.cg_update_xr(cur, sl[3], None, None, st.Ap, None, st.r, sl[4], alpha_out=sl[2])
        es[3].record(); ops.cg_record(i + 1, sl[4], st.crit_d, hp - (i + 1) * 16 * 8, st.stop_at, rho_keep=nxt)
        es[4].record(); torch.cuda.synchronize()
        for j in range(4): tot[j] += es[j].elapsed_time(es[j + 1]) / 5
    print("spmm cfg", spmm, {nm: round(t, 3) for nm, t in zip(names, tot)}, flush=True)
    # whole-loop C entry
    e0, e1 = ev(), ev()
    st.kk = 8; e0.record(); st.run(20); e1.record(); torch.cuda.synchronize()
    print("   kb_cg_run 20 its:", round(e0.elapsed_time(e1) / 20, 3), "ms/it", flush=True)
    del st
