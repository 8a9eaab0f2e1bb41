"""Where does the end-to-end time go?  (upload / matrix set-up / solve / download)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse
import krylov_b200 as kb
from krylov_b200.generate import device_stencil7
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
A = device_stencil7(N, N, N); n = A.shape[0]
g = torch.Generator(device="cuda").manual_seed(1234)
b = A.matvec_device(torch.randn((n, 1), generator=g, dtype=torch.float64, device="cuda"))
rp = torch.empty(n + 1, dtype=torch.int32).pin_memory(); ci = torch.empty(A.nnz, dtype=torch.int32).pin_memory()
va = torch.empty(A.nnz, dtype=torch.float64).pin_memory(); bh = torch.empty(n, dtype=torch.float64).pin_memory()
rp.copy_(A.rowptr); ci.copy_(A.colidx[:A.nnz]); va.copy_(A.vals[:A.nnz]); bh.copy_(b.reshape(-1)); torch.cuda.synchronize()
del A
def T(): torch.cuda.synchronize(); return time.perf_counter()
t0 = T()
d_va = va.to("cuda"); t1 = T(); print(f"H2D vals {va.numel()*8/1e9:.2f} GB: {t1-t0:.3f}s = {va.numel()*8/1e9/(t1-t0):.1f} GB/s")
d_va2 = torch.as_tensor(va.numpy()).to("cuda"); t2 = T(); print(f"H2D vals via numpy view: {t2-t1:.3f}s (is_pinned={torch.as_tensor(va.numpy()).is_pinned()})")
del d_va, d_va2
Ah = scipy.sparse.csr_matrix((va.numpy(), ci.numpy(), rp.numpy()), shape=(n, n), copy=False)
t3 = T(); print(f"scipy wrap: {t3-t2:.3f}s")
Ad = kb.CsrMatrix.from_scipy(Ah); t4 = T(); print(f"CsrMatrix.from_scipy (upload + pad + stats + pattern): {t4-t3:.3f}s  schedule={Ad.info()['schedule']}")
sol, info = kb.cg(Ad, bh.numpy(), tol=1e-8, maxiter=20000); t5 = T()
print(f"cg (device matrix, numpy b -> numpy x): {t5-t4:.3f}s for {info.numsteps} steps = {info.numsteps/(t5-t4):.1f} it/s")
bd = torch.as_tensor(bh.numpy()).to("cuda"); t6 = T()
sol, info = kb.cg(Ad, bd, tol=1e-8, maxiter=20000); t7 = T()
print(f"cg (all device): {t7-t6:.3f}s = {info.numsteps/(t7-t6):.1f} it/s")
x = info.xk.cpu().numpy(); t8 = T(); print(f"D2H x pageable: {t8-t7:.3f}s")
