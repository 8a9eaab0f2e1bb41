import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
import krylov_b200 as kb
from krylov_b200._lib import lib
from krylov_b200.generate import device_stencil7
for N in (256, 512):
    A = device_stencil7(N, N, N)
    n = A.shape[0]
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    byt = 12 * A.nnz + 4 * (n + 1) + 92 * n
    for sched, cfg, ctas, vc in (("rowwise",0,0,8),("stream",0,0,8),("pattern",0,0,8)):
        A.set_schedule(sched); lib.kb_tune(0,cfg); lib.kb_tune(1,ctas); lib.kb_tune(2,vc)
        kb.cg(A, b, tol=0.0, atol=0.0, maxiter=10)
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); kb.cg(A, b, tol=0.0, atol=0.0, maxiter=100); e1.record(); torch.cuda.synchronize()
        dt = e0.elapsed_time(e1)/1e3
        print(f"N={N} cg {sched} cfg={cfg} ctas={ctas} vec={vc}: {100/dt:.1f} it/s {byt*100/dt/1e9:.0f} GB/s", flush=True)
    del A, b; torch.cuda.empty_cache()
