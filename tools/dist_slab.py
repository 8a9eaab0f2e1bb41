"""Row-partitioned fused CG on slabs of `planes` 512 x 512 planes per rank (the per-rank share of
the 512^3 problem on 8 GPUs is 64 planes) under torchrun with any number of ranks: step time and
per-kernel phases (max over ranks and per rank), the latency of the stand-alone peer-memory
all-reduce, and the same slab without communication for reference.
usage: torchrun ... dist_slab.py [planes] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ctypes as C
from krylov_b200._lib import lib, check
from krylov_b200.cg import FusedCG
from krylov_b200.dist import Comm, dist_stencil7
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops, ptr, cur_stream

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
NX = 512
comm = Comm()

def say(*a):
    if rank == 0:
        print(*a, flush=True)

# stand-alone all-reduce latency
o = Ops(1024, 1, dev, comm=comm)
sl = o.slots(1)[0]
for _ in range(20):
    check(lib.kb_allreduce(o.ws.handle, 1, ptr(sl), cur_stream()))
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    sl.fill_(1.0)
    check(lib.kb_allreduce(o.ws.handle, 1, ptr(sl), cur_stream()))
e1.record(); torch.cuda.synchronize()
t_ar = e0.elapsed_time(e1) / 200
e0.record()
for _ in range(200):
    sl.fill_(1.0)
e1.record(); torch.cuda.synchronize()
t_fill = e0.elapsed_time(e1) / 200
assert float(sl) == 1.0
say(f"ranks={world}: stand-alone all-reduce launch {1e3*(t_ar - t_fill):.1f} us (fill+allreduce {1e3*t_ar:.1f}, fill {1e3*t_fill:.1f})")
comm.check_p2p()

A = dist_stencil7(NX, NX, planes * world, comm=comm)
n = A.shape[0]
g = torch.Generator(device=dev).manual_seed(rank)
xs = torch.randn(n, 1, generator=g, dtype=torch.float64, device=dev)
b = A.matvec_device(xs)
def measure(tag, dbg=0, local_r=False):
    FusedCG._debug_local_r = local_r
    lib.kb_tune(21, dbg)
    st = FusedCG(A, b, torch.zeros_like(b), 0.0, 0.0)
    assert st.gplan is not None, "fused partitioned path not taken"
    st.run(60)
    res = []
    for rep in range(5):
        dist.barrier(); torch.cuda.synchronize()
        ph, tot, fused = st.run_timed(steps)
        res.append((tot / steps, ph[0], ph[1]))
    lib.kb_tune(21, 0)
    FusedCG._debug_local_r = False
    m = np.median(np.array(res), axis=0)
    t = torch.tensor(m, dtype=torch.float64, device=dev)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    allt = torch.stack(allt).cpu().numpy() * 1e3
    say(f"{tag}: step max {allt[:,0].max():.1f} us -> {1e6/allt[:,0].max():.0f} it/s; KIND1 per rank "
        f"{np.round(allt[:,1],1).tolist()}; KIND2 per rank {np.round(allt[:,2],1).tolist()}")
    del st

measure("partitioned (full)                     ")
if len(sys.argv) > 3:
    measure("no pushes, no all-reduce               ", 3)
    measure("no pushes, no all-reduce, r in torch mem", 3, True)
    measure("pushes, no all-reduce                  ", 2)
    measure("no pushes, all-reduce                  ", 1)
    measure("partitioned (full) again               ")
A.check_p2p()

# the same slab alone (no neighbours, no all-reduce): what the kernels cost without communication
A1 = device_stencil7(NX, NX, planes)
b1 = A1.matvec_device(xs)
s1 = FusedCG(A1, b1, torch.zeros_like(b1), 0.0, 0.0)
s1.run(100)
res = []
for rep in range(5):
    dist.barrier(); torch.cuda.synchronize()
    ph, tot, fused = s1.run_timed(steps)
    res.append((tot / steps, ph[0], ph[1]))
m = np.median(np.array(res), axis=0)
t = torch.tensor(m, dtype=torch.float64, device=dev)
allt = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(allt, t)
allt = torch.stack(allt).cpu().numpy() * 1e3
say(f"slab alone (all ranks concurrently): step per rank {np.round(allt[:,0],1).tolist()}; KIND1 {np.round(allt[:,1],1).tolist()}; KIND2 {np.round(allt[:,2],1).tolist()}")
dist.barrier()
dist.destroy_process_group()
