"""Per-kernel timing of one distributed CG iteration (run under torchrun).
usage: dist_phase.py NX NY NZ   -- z-slabs of NZ/world planes per rank"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import krylov_b200 as kb
from krylov_b200._lib import lib, check
from krylov_b200.device import cur_stream, ptr
from krylov_b200.dist import Comm, dist_stencil7
from krylov_b200.cg import FusedCG
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
nx, ny, nz = (int(a) for a in sys.argv[1:4])
comm = Comm()
A = dist_stencil7(nx, ny, nz, comm=comm)
n = A.shape[0]
b = torch.randn((n, 1), dtype=torch.float64, device=dev)
st = FusedCG(A, b, torch.zeros_like(b), 0.0, 0.0)
ops, sl, p = st.ops, st.sl, A.plan
hp = st.hist.data_ptr()
def ev(): return torch.cuda.Event(enable_timing=True)
for i in range(5): st.enqueue(i, hp - (i + 1) * 8)
torch.cuda.synchronize(); dist.barrier()
# whole iterations, pipelined
e0, e1 = ev(), ev(); e0.record()
for i in range(5, 205): st.enqueue(i, hp - (i + 1) * 8)
e1.record(); torch.cuda.synchronize()
whole = e0.elapsed_time(e1) / 200
# per kernel (host-synchronised between kernels: no overlap, no pipelining)
names = ["update_p", "push", "spmv_local", "halo_add+allreduce", "update_r+allreduce", "record"]
tot = np.zeros(len(names)); reps = 20
halo = A._halo_for(1)
for i in range(205, 205 + reps):
    cur, nxt = sl[i % 2], sl[(i + 1) % 2]
    es = [ev() for _ in range(7)]
    dist.barrier(); torch.cuda.synchronize()
    es[0].record(); ops.cg_update_p(cur, nxt, st.r, st.p, x=st.yk, alpha=sl[2])
    es[1].record(); check(lib.kb_halo_push(halo, ops.ws.handle, 1, p.n_seg, ptr(p.segs), p.n_send, ptr(p.send_idx), ptr(st.p), cur_stream()))  # (timed alone on the compute stream)
    es[2].record(); ops.set_collective(False); check(lib.kb_spmv(A.A_loc.handle, ops.ws.handle, 1, ptr(st.p), ptr(st.Ap), 0, None, None, 1, ptr(st.p), ptr(sl[3]), cur_stream())); ops.set_collective(True)
    es[3].record(); check(lib.kb_spmv_halo_add(ops.ws.handle, 1, p.n_brows, 1.0, ptr(p.h_rows), ptr(p.h_rowptr), ptr(p.h_col), ptr(p.h_val), None, ptr(st.Ap), 1, ptr(st.p), ptr(sl[3]), halo, ptr(p.srcs), p.n_src, cur_stream()))
    es[4].record(); ops.cg_update_xr(cur, sl[3], None, None, st.Ap, None, st.r, sl[4], alpha_out=sl[2])
    es[5].record(); ops.cg_record(i + 1, sl[4], st.crit_d, hp - (i + 1) * 8, st.stop_at, rho_keep=nxt)
    es[6].record(); torch.cuda.synchronize()
    tot += np.array([es[j].elapsed_time(es[j + 1]) for j in range(6)]) / reps
t = torch.tensor(np.concatenate([[whole], tot]), device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    t = t.cpu().numpy()
    print(f"grid {nx}x{ny}x{nz} on {world} ranks ({n} rows/rank, n_brows {p.n_brows}, n_send {p.n_send}); max over ranks")
    print(f"  pipelined iteration: {t[0]*1e3:.1f} us")
    for nm, v in zip(names, t[1:]): print(f"  {nm:22s} {v*1e3:8.1f} us")
    print(f"  sum of kernels      {t[1:].sum()*1e3:8.1f} us")
A.check_p2p()
dist.barrier(); dist.destroy_process_group()
