"""DRAM traffic of the stencil SpMV under ncu (--metrics dram bytes, duration): sizes x L2 hints.
Run as: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
        -k regex:kb_spmv_stencil --csv --log-file out.csv python tools/stencil_traffic.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib.kb_tune(10, cfg)
for N in (384, 448, 480, 512):
    A = device_stencil7(N, N, N)
    n = A.shape[0]
    ops = Ops(n, 1)
    x = torch.randn(n, 1, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    out = ops.slots(1)
    for pol in (0, 1, 2):
        lib.kb_tune(12, pol)
        for _ in range(2):
            ops.spmv(A, x, y, dot=1, w=x, out=out[0])
        torch.cuda.synchronize()
        print(f"N={N} pol={pol} model_bytes={A.moved_bytes(1)}", flush=True)
    del A, x, y, ops
lib.kb_tune(12, 0)
