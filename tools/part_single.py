"""The row-partitioned variants of the fused CG kernels (PART = ghost-extended row space) on ONE
GPU without neighbours, next to the plain single-GPU kernels on the same 512 x 512 x planes slab:
what the PART code path itself costs, separated from communication.  usage: part_single.py [planes] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from krylov_b200._lib import lib, check, CgState
from krylov_b200.cg import FusedCG
from krylov_b200.device import ptr, cur_stream, view_device_memory
from krylov_b200.generate import device_stencil7

planes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
A = device_stencil7(512, 512, planes)
n = A.shape[0]
P = 512 * 512
g = torch.Generator(device="cuda").manual_seed(0)
b = A.matvec_device(torch.randn(n, 1, generator=g, dtype=torch.float64, device="cuda"))
x0 = torch.zeros_like(b)

MC = [tuple(int(v) for v in a.split(":")) for a in os.environ.get("MC", "0:0,4:5,4:4,0:5,-1:-1").split(",")]
st = FusedCG(A, b, x0, 0.0, 0.0)
st.run(60)
for mc in MC:
    lib.kb_tune(22, mc[0]); lib.kb_tune(23, mc[1])
    res = [st.run_timed(steps) for _ in range(5)]
    m = np.median(np.array([(t / steps, ph[0], ph[1]) for ph, t, _ in res]), axis=0) * 1e3
    print(f"plain  shapes={mc}: step {m[0]:.1f} us  KIND1 {m[1]:.1f}  KIND2 {m[2]:.1f}", flush=True)
lib.kb_tune(22, -1); lib.kb_tune(23, -1)

# the same state on the ghost-extended row space, by hand (no neighbours: ghost planes stay 0)
n_ext = n + 2 * P
si = A.stencil_info()
masks = torch.zeros(n_ext, dtype=torch.int16, device="cuda")
masks[P:P + n] = view_device_memory(si["masks_ptr"], n, torch.int16, torch.device("cuda", 0))
st2 = FusedCG(A, b, x0, 0.0, 0.0)
r_ext = torch.zeros(n_ext, dtype=torch.float64, device="cuda")
r_ext[P:P + n] = st2.r.reshape(-1)
p_ext = [torch.zeros(n_ext, dtype=torch.float64, device="cuda") for _ in (0, 1)]
sl = torch.zeros((7, 1), dtype=torch.float64, device="cuda")
sl[:6] = st2.sl[:6]
cs = CgState(A=A.handle, n=n, k=1, x=ptr(st2.yk), r=ptr(r_ext), p=ptr(p_ext[0]), Ap=ptr(st2.Ap),
             slots=ptr(sl), crit=ptr(st2.crit_d), hist=ptr(st2.hist), stop_at=ptr(st2.stop_at),
             p2=ptr(p_ext[1]), pcur=0, masks_ext=ptr(masks), n_ext=n_ext, own_lo=P,
             r_push_lo=None, r_push_hi=None)
fz = C.c_int(0)
check(lib.kb_cg_is_fused(C.byref(cs), C.byref(fz)))
assert fz.value
kk = 0
def run(nb, timed=False):
    global kk
    st2.stop_at.fill_(2**31 - 1)
    cs.pcur = kk % 2
    if timed:
        ms = (C.c_float * 3)(); tot = C.c_float(0)
        check(lib.kb_cg_run_timed(st2.ops.ws.handle, C.byref(cs), kk, nb, 1 if kk else 0, cur_stream(), ms, C.byref(tot)))
        kk += nb
        return (tot.value / nb, ms[0], ms[1])
    check(lib.kb_cg_run(st2.ops.ws.handle, C.byref(cs), kk, nb, 1 if kk else 0, cur_stream()))
    kk += nb
run(60)
for mc in MC:
    lib.kb_tune(22, mc[0]); lib.kb_tune(23, mc[1])
    m = np.median(np.array([run(steps, True) for _ in range(5)]), axis=0) * 1e3
    print(f"PART   shapes={mc}: step {m[0]:.1f} us  KIND1 {m[1]:.1f}  KIND2 {m[2]:.1f}", flush=True)
lib.kb_tune(22, -1); lib.kb_tune(23, -1)
torch.cuda.synchronize()
