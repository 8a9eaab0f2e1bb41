"""Marching SpMV (KIND 0) on N^3 grids: planes per work item (kb_tune 13) and the even-grid option
(kb_tune 20).  usage: march_ch.py [N ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from krylov_b200._lib import lib
from krylov_b200.device import Ops
from krylov_b200.generate import device_stencil7
for N in [int(a) for a in sys.argv[1:]] or [256]:
    A = device_stencil7(N, N, N)
    n = A.shape[0]
    ops = Ops(n, 1)
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    out = ops.slots(1)
    for even in (0, 1):
        for ch in (0, 4, 8, 16, 32, 64):
            lib.kb_tune(13, ch); lib.kb_tune(20, even)
            for dot in (0, 1):
                for _ in range(5):
                    ops.spmv(A, x, y, dot=dot, w=x, out=out[0])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    ops.spmv(A, x, y, dot=dot, w=x, out=out[0])
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 50 * 1e3
                print(f"N={N} even={even} ch={ch:2d} dot={dot}: {us:7.1f} us  {18 * n / us / 1e3:6.0f} GB/s", flush=True)
    lib.kb_tune(13, 0); lib.kb_tune(20, 0)
