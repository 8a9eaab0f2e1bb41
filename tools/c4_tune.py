import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7
from krylov_b200.device import Ops
N = 256; A = device_stencil7(N, N, N); n = A.shape[0]
for k in (16, 4):
    ops = Ops(n, k)
    x = torch.randn(n, k, dtype=torch.float64, device="cuda"); y = torch.empty_like(x); out = ops.slots(1)[0]
    for contig, ctas in ((0, 0), (1, 2), (1, 3), (1, 4), (1, 6), (1, 8)):
        lib.kb_tune(5, contig); lib.kb_tune(6, ctas)
        for _ in range(3): ops.spmv(A, x, y, dot=1, w=x, out=out)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.spmv(A, x, y, dot=1, w=x, out=out)
        e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)/10
        print(f"k={k} contig={contig} ctas={ctas}: {ms:.3f} ms  {A.spmv_bytes(k)/ms/1e6:.0f} GB/s", flush=True)
lib.kb_tune(5, -1); lib.kb_tune(6, 0)
