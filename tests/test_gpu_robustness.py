"""-m gpu edge cases of the drop-in surface: input kinds, shapes, dtypes,
degenerate problems (the reference's behaviour on each is noted)."""
import numpy as np
import pytest
import scipy.sparse
import torch

import krylov_b200 as kb
from krylov_b200 import stencils as st
from oracle import krylov_oracle as orc

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(3)


@pytest.mark.parametrize("solver", ["cg", "minres", "gmres"])
def test_input_kinds(solver):
    fn, ofn = getattr(kb, solver), getattr(orc, solver)
    A = st.poisson2d(12)
    b = A @ rng.standard_normal(A.shape[0])
    ref, iref = ofn(A, b, tol=1e-9)
    for Ak in (A, A.tocsc(), A.tocoo(), A.toarray(), kb.CsrMatrix.from_scipy(A),
               torch.from_numpy(A.toarray()), torch.from_numpy(A.toarray()).to_sparse_csr()):
        sol, info = fn(Ak, b, tol=1e-9)
        assert info.numsteps == iref.numsteps
        assert np.linalg.norm(sol - ref) <= 1e-10 * np.linalg.norm(ref)
    # list / float32 / integer right-hand sides are converted like np.asarray would
    sol, info = fn(A, list(b), tol=1e-9)
    assert np.linalg.norm(sol - ref) <= 1e-10 * np.linalg.norm(ref)
    sol32, _ = fn(A, b.astype(np.float32), tol=1e-6)
    assert sol32.dtype == np.float64 and np.linalg.norm(sol32 - ref) <= 1e-5 * np.linalg.norm(ref)
    ones, _ = fn(A, np.ones(A.shape[0], dtype=np.int64), tol=1e-9)
    assert np.linalg.norm(A @ ones - 1.0) <= 1e-7
    # x0 as list, non-contiguous b
    B = np.asfortranarray(A @ rng.standard_normal((A.shape[0], 3)))
    sol, info = fn(A, B, x0=np.zeros_like(B).tolist(), tol=1e-9)
    so, io = ofn(A, np.ascontiguousarray(B), tol=1e-9)
    assert info.numsteps == io.numsteps and np.linalg.norm(sol - so) <= 1e-10 * np.linalg.norm(so)


@pytest.mark.parametrize("solver", ["cg", "minres", "gmres"])
def test_shapes_and_limits(solver):
    fn = getattr(kb, solver)
    # 1 x 1 system
    sol, info = fn(np.array([[4.0]]), np.array([2.0]))
    assert info.success and abs(sol[0] - 0.5) < 1e-15
    # trailing dimensions are flattened to k columns: (n, 2, 2)
    A = st.poisson2d(6)
    B = (A @ rng.standard_normal((36, 4))).reshape(36, 2, 2)
    sol, info = fn(A, B, tol=1e-9)
    assert sol.shape == (36, 2, 2) and np.asarray(info.resnorms).shape == (info.numsteps + 1, 2, 2)
    assert np.linalg.norm((A @ sol.reshape(36, 4)).reshape(36, 2, 2) - B) <= 1e-7
    # maxiter = 0: no step, not converged
    sol, info = fn(A, B, maxiter=0)
    assert sol is None and info.numsteps == 0 and len(info.resnorms) == 1
    # widest supported block and one beyond
    n = 40
    D = scipy.sparse.diags(np.linspace(1.0, 2.0, n)).tocsr()
    Bw = rng.standard_normal((n, 256))
    sol, info = fn(D, Bw, tol=1e-10)
    assert info.success and np.allclose(D @ sol, Bw, atol=1e-7)
    with pytest.raises(ValueError):
        fn(D, rng.standard_normal((n, 257)))
    with pytest.raises(ValueError):
        fn(object.__new__(type("NoMatmul", (), {"shape": (n, n)})), np.ones(n))


def test_cg_breakdown_guards_match_reference():
    """Indefinite / singular inputs run into the reference's zero guards
    (cg.py:177,185) instead of raising; same step counts and histories."""
    A = np.diag([1.0, -1.0, 2.0, 0.0])
    b = np.array([1.0, 1.0, 1.0, 0.0])
    sol, info = kb.cg(A, b, tol=1e-12, maxiter=10)
    so, io = orc.cg(A, b, tol=1e-12, maxiter=10)
    assert info.numsteps == io.numsteps and info.success == io.success
    np.testing.assert_allclose(np.asarray(info.resnorms), np.asarray(io.resnorms), rtol=1e-10, atol=1e-14)


def test_invariant_subspace_raises_like_reference():
    A = np.diag([2.0, 2.0, 2.0, 2.0]) + 0.0
    A[3, 3] = 0.0
    b = np.array([1.0, 1.0, 1.0, 1.0])  # after one step the Krylov space is exhausted, residual stays
    for name in ("minres", "gmres"):
        with pytest.raises(orc.ArgumentError):
            getattr(orc, name)(A, b, tol=1e-14, maxiter=4)
        with pytest.raises(kb.ArgumentError):
            getattr(kb, name)(A, b, tol=1e-14, maxiter=4)


def test_callback_can_override_resnorm():
    """minres/gmres hand the callback an array it may overwrite (minres.py:226-234)."""
    A = st.poisson2d(8)
    b = A @ rng.standard_normal(64)

    def cb(x, r):
        r[...] = 0.0  # pretend convergence

    # (with gmres the same callback drives the reference into a singular triangular
    # solve -- scipy raises LinAlgError -- so only minres pins this behaviour)
    _, info = kb.minres(A, b, tol=1e-6, callback=cb)
    _, io = orc.minres(A, b, tol=1e-6, callback=cb)
    assert info.numsteps == io.numsteps and bool(info.success) == bool(io.success)
    np.testing.assert_allclose(np.asarray(info.resnorms), np.asarray(io.resnorms), rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("k", [1, 3])
def test_fused_preconditioned_cg_vs_oracle(k):
    """Matrix preconditioners stay on the device-resident path (z = M r fused with
    rho = <r, z>; Ml chained behind A): Jacobi M, sparse Ml, both, vs the oracle."""
    n = 12
    A = st.poisson3d(n).tolil()
    d = 1.0 + rng.random(n ** 3) * 50.0           # badly scaled diagonal: Jacobi matters
    for i in range(n ** 3):
        A[i, i] = A[i, i] + d[i]
    A = A.tocsr()
    M = scipy.sparse.diags(1.0 / A.diagonal()).tocsr()
    shape = (A.shape[0],) if k == 1 else (A.shape[0], k)
    b = A @ rng.standard_normal(shape)
    plain_steps = orc.cg(A, b, tol=1e-10)[1].numsteps
    # commuting, symmetric Ml so that Ml A stays self-adjoint (reference test_ml uses diagonals)
    Dl = scipy.sparse.identity(A.shape[0], format="csr") * 0.5
    for kw in (dict(M=M), dict(Ml=Dl), dict(M=M, Ml=Dl)):
        sol, info = kb.cg(A, b, tol=1e-10, **kw)
        so, io = orc.cg(A, b, tol=1e-10, **kw)
        assert info.success and info.numsteps == io.numsteps
        ro, rg = np.asarray(io.resnorms, float), np.asarray(info.resnorms, float)
        live = ro / ro[0] >= 1e-6
        assert np.all(np.abs(rg - ro)[live] <= 1e-8 * ro[live])
        assert np.linalg.norm(sol - so) <= 1e-10 * np.linalg.norm(so)
    assert kb.cg(A, b, tol=1e-10, M=M)[1].numsteps < plain_steps  # the preconditioner did its job
    # the device-resident path was taken: no duck-typed operator involved
    from krylov_b200.operators import Problem
    prob = Problem(A, b)
    assert prob.operator(M).csr is not None and prob.operator(Dl).csr is not None


@pytest.mark.parametrize("k", [1, 3])
def test_fused_preconditioned_minres_vs_oracle(k):
    """MINRES with matrix M / Ml / Mr stays device-resident (chain Ml A Mr with the fusion
    on its last product, V = M P bases, beta^2 = <Av, M Av> out of M's product)."""
    n = 10
    A = st.shifted_laplace3d(n).tolil()            # symmetric indefinite
    d = rng.random(n ** 3) * 20.0
    for i in range(n ** 3):
        A[i, i] = A[i, i] + d[i]
    A = A.tocsr()
    N = A.shape[0]
    M = scipy.sparse.diags(1.0 / (1.0 + np.abs(A.diagonal()))).tocsr()   # SPD
    Dl = scipy.sparse.identity(N, format="csr") * 0.5
    Dr = scipy.sparse.identity(N, format="csr") * 2.0
    shape = (N,) if k == 1 else (N, k)
    b = A @ rng.standard_normal(shape)
    import sys
    kmin = sys.modules["krylov_b200.minres"]
    calls = []
    orig = kmin._minres_fused
    kmin._minres_fused = lambda *a, **kw: (calls.append(1), orig(*a, **kw))[1]
    try:
        for kw in (dict(M=M), dict(Ml=Dl), dict(Mr=Dr), dict(Ml=Dl, Mr=Dr),
                   dict(M=M, Ml=Dl, Mr=Dr)):
            sol, info = kb.minres(A, b, tol=1e-9, maxiter=3000, **kw)
            so, io = orc.minres(A, b, tol=1e-9, maxiter=3000, **kw)
            assert info.success and io.success, kw
            assert abs(info.numsteps - io.numsteps) <= max(2, io.numsteps // 50), kw
            m = min(info.numsteps, io.numsteps)
            ro, rg = np.asarray(io.resnorms, float)[:m], np.asarray(info.resnorms, float)[:m]
            live = ro / ro[0] >= 1e-5
            assert np.all(np.abs(rg - ro)[live] <= 1e-7 * ro[live]), kw
            assert np.linalg.norm(sol - so) <= 1e-7 * np.linalg.norm(so), kw
    finally:
        kmin._minres_fused = orig
    assert len(calls) == 5     # every variant took the device-resident path


@pytest.mark.parametrize("k", [1, 3])
def test_preconditioned_gmres_sparse_vs_oracle(k):
    """GMRES with sparse M / Ml / Mr: the chain Ml A Mr runs as device products with the first
    Gram-Schmidt dot fused into the last one; <w, M w> comes out of M's product."""
    A = st.convection_diffusion3d(8)
    N = A.shape[0]
    M = scipy.sparse.diags(1.0 / (1.0 + rng.random(N))).tocsr()          # SPD
    Jl = scipy.sparse.diags(1.0 / A.diagonal()).tocsr()
    Jr = scipy.sparse.diags(0.5 + rng.random(N)).tocsr()
    shape = (N,) if k == 1 else (N, k)
    b = A @ rng.standard_normal(shape)
    for kw in (dict(M=M), dict(Ml=Jl), dict(Mr=Jr), dict(Ml=Jl, Mr=Jr),
               dict(M=M, Ml=Jl, Mr=Jr), dict(Ml=Jl, Mr=Jr, ortho="mgs2")):
        sol, info = kb.gmres(A, b, tol=1e-9, maxiter=200, **kw)
        so, io = orc.gmres(A, b, tol=1e-9, maxiter=200, **kw)
        assert info.success and io.success, kw
        assert abs(info.numsteps - io.numsteps) <= 1, kw
        m = min(info.numsteps, io.numsteps)
        ro, rg = np.asarray(io.resnorms, float)[:m], np.asarray(info.resnorms, float)[:m]
        live = ro / ro[0] >= 1e-5
        assert np.all(np.abs(rg - ro)[live] <= 1e-7 * ro[live]), kw
        assert np.linalg.norm(sol - so) <= 1e-7 * np.linalg.norm(so), kw


def test_malformed_csr_is_refused():
    """kb_csr_create checks the structure once (rowptr[0], monotone row pointers, rowptr[n] ==
    nnz, column range): a malformed matrix raises instead of being read out of bounds."""
    import krylov_b200 as kb

    rp = np.array([0, 2, 3], dtype=np.int32)
    ci = np.array([0, 1, 1], dtype=np.int32)
    va = np.ones(3)
    kb.CsrMatrix(rp, ci, va, (2, 2))  # well-formed
    for bad_rp, bad_ci in ((np.array([1, 2, 3], dtype=np.int32), ci),
                           (np.array([0, 3, 2], dtype=np.int32), ci),
                           (np.array([0, 2, 2], dtype=np.int32), ci),
                           (rp, np.array([0, 2, 1], dtype=np.int32)),
                           (rp, np.array([0, -1, 1], dtype=np.int32))):
        with pytest.raises(kb.KrylovB200Error, match="malformed CSR"):
            kb.CsrMatrix(bad_rp, bad_ci, va, (2, 2))
