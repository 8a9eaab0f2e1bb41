"""Small end-to-end exercise of every kernel family, meant to run under
compute-sanitizer (memcheck or racecheck, one tool per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.sparse
import krylov_b200 as kb
from krylov_b200 import stencils as st
rng = np.random.default_rng(0)
for A in (st.poisson3d(10), st.poisson2d(33), st.convection_diffusion3d(8),
          scipy.sparse.random(700, 700, density=0.02, random_state=1, format="csr") + scipy.sparse.eye(700)):
    A = A.tocsr()
    for sched in ("rowwise", "stream", "pattern"):
        Ad = kb.CsrMatrix.from_scipy(A)
        try:
            Ad.set_schedule(sched)
        except kb.KrylovB200Error:
            continue
        for k in (1, 3):
            x = rng.standard_normal((A.shape[1], k)) if k > 1 else rng.standard_normal(A.shape[1])
            assert np.array_equal(Ad @ x, A @ x), (sched, k)
A = st.poisson3d(10); b = A @ rng.standard_normal(A.shape[0])
for name, kw in (("cg", {}), ("minres", {}), ("gmres", {"maxiter": 60}), ("gmres", {"maxiter": 60, "ortho": "mgs2"}),
                 ("gmres", {"maxiter": 60, "ortho": "householder"})):
    sol, info = getattr(kb, name)(A, b, tol=1e-9, **kw)
    assert info.success, name
B = A @ rng.standard_normal((A.shape[0], 4))
for name in ("cg", "minres", "gmres"):
    sol, info = getattr(kb, name)(A, B, tol=1e-8, maxiter=200)
    assert info.success, name
M = scipy.sparse.diags(1.0 / A.diagonal())
sol, info = kb.cg(A, b, M=M, tol=1e-9); assert info.success
print("sanitize_small ok")
