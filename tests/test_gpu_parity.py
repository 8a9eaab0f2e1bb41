"""-m gpu parity tests: the CUDA product (through the C ABI) against
  (1) outputs of the unmodified reference on the same seeded inputs
      (tests/golden/*.npz, written by make_golden.py),
  (2) the reference's own known-answer vectors and test invariants,
  (3) the oracle run live on the same inputs.

Tolerances are BASELINE.json's: final solution relative error <= 1e-10,
per-iteration residual norms within 1e-8 relative until the residual has
dropped by 1e-6, iteration counts within +-2 %.
"""
import os

import numpy as np
import pytest
import scipy.sparse
import torch

import cases
import krylov_b200 as kb
from oracle import krylov_oracle as orc

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
SOL = np.load(os.path.join(G, "solvers.npz"))
ARN = np.load(os.path.join(G, "arnoldi.npz"))
SML = np.load(os.path.join(G, "small.npz"))
CASES = cases.solver_cases()

RES_RTOL = 1e-8     # residual-norm history
RES_FLOOR = 1e-6    # ... compared while resnorm/resnorm[0] >= this
SOL_RTOL = 1e-10    # final solution


def check_history(res, ref_res, explicit_noise=0.0):
    """`explicit_noise`: absolute rounding floor of an *explicitly* computed
    residual ||b - A x|| (8 eps ||A||_inf ||x||): when a solver stops, the last
    entry is overwritten by that quantity (cg.py:156-164), which no
    implementation can reproduce more precisely than its own rounding."""
    res, ref_res = np.asarray(res, float), np.asarray(ref_res, float)
    assert res.shape == ref_res.shape
    r0 = np.where(ref_res[0] > 0, ref_res[0], 1.0)
    live = ref_res / r0 >= RES_FLOOR
    abs_err = np.abs(res - ref_res)
    abs_err[-1] = np.maximum(abs_err[-1] - explicit_noise, 0.0)
    err = abs_err / np.where(ref_res > 0, ref_res, 1.0)
    assert np.all(err[live] <= RES_RTOL), f"max rel err {err[live].max():.3e}"
    # below the floor rounding noise dominates; stay within 1e-3
    assert np.all(err[~live & (ref_res > 1e-300)] <= 1e-3)


def check_steps(n, n_ref):
    assert abs(n - n_ref) <= max(0.02 * n_ref, 0), (n, n_ref)


@pytest.mark.parametrize("name", sorted(CASES))
def test_solver_matches_reference(name):
    solver, A, b, kw = CASES[name]
    sol, info = getattr(kb, solver)(A, b, **kw)
    n_ref = int(SOL[name + "_numsteps"])
    check_steps(info.numsteps, n_ref)
    assert bool(info.success) == bool(SOL[name + "_success"])
    assert (sol is None) == bool(SOL[name + "_solnone"])
    ref_x = SOL[name + "_xk"]
    if info.numsteps == n_ref:
        Ainf = abs(A).sum(axis=1).max()
        noise = 8 * np.finfo(float).eps * Ainf * np.linalg.norm(ref_x) if info.success else 0.0
        check_history(info.resnorms, SOL[name + "_resnorms"], noise)
    scale = max(np.linalg.norm(ref_x), 1e-300)
    assert np.linalg.norm(np.asarray(info.xk) - ref_x) <= SOL_RTOL * scale
    # reference tests/helpers.py:4-23 invariants
    if sol is not None:
        assert sol.shape == b.shape
        assert np.may_share_memory(sol, info.xk)
    rn = np.asarray(info.resnorms)
    assert np.issubdtype(rn.dtype, np.floating)
    assert rn.shape == (info.numsteps + 1, *b.shape[1:])


# reference tests/test_solvers.py:123-144
@pytest.mark.parametrize(
    "method, ref",
    [
        ("cg", [1004.1873775173957, 1000.0003174916551, 999.9999999997555]),
        ("gmres", [1004.1873724888546, 1000.0003124630923, 999.999994971191]),
        ("minres", [1004.187372488912, 1000.0003124632159, 999.9999949713145]),
    ],
)
@pytest.mark.parametrize("shape", [(100,), (100, 1)])
def test_reference_known_answers(method, ref, shape):
    tol = 1.0e-11
    A = cases.kat_matrix(shape[0])
    b = np.ones(shape)
    sol, info = getattr(kb, method)(A, b)
    assert sol.shape == b.shape
    assert info.numsteps == 55
    assert abs(np.sum(np.abs(sol)) - ref[0]) < tol * ref[0]
    assert abs(np.sqrt(np.dot(sol.T, sol)) - ref[1]) < tol * ref[1]
    assert abs(np.max(np.abs(sol)) - ref[2]) < tol * ref[2]


def _consistent(A, b, info, sol, tol):
    """reference tests/helpers.py:4-23"""
    res = b - A @ info.xk
    resnorm = np.sqrt(np.einsum("i...,i...->...", res, res))
    bnorm = np.sqrt(np.einsum("i...,i...->...", b, b))
    if info.success:
        assert sol.shape == b.shape
        assert np.all(resnorm < tol * (1.0 + bnorm))
        assert np.may_share_memory(sol, info.xk)
    assert np.all(np.abs(resnorm - info.resnorms[-1]) <= 1.0e-12 * (1 + resnorm))
    assert np.asarray(info.resnorms).shape == (info.numsteps + 1, *b.shape[1:])


def _spd(shape):
    a = np.linspace(1.0, 2.0, shape[0])
    a[-1] = 1e-2
    return np.diag(a), np.ones(shape)


def _spd_rhs_0sol0():
    a = np.linspace(1.0, 2.0, 5)
    a[-1] = 1e-2
    A = np.diag(a)
    np.random.seed(0)
    b = np.column_stack([np.zeros(5), np.random.rand(5), np.random.rand(5)])
    sol = np.linalg.solve(A, b[:, 1])
    return A, np.column_stack([np.zeros(5), sol, np.zeros(5)])


def _sym_indef():
    a = np.linspace(1.0, 2.0, 5)
    a[-1] = -1.0
    return np.diag(a), np.ones(5)


def _unsym():
    a = np.arange(1, 6, dtype=float)
    a[-1] = -10.0
    A = np.diag(a)
    A[0, -1] = 10.0
    return A, np.ones(5)


REAL_PROBLEMS = [_spd((5,)), _spd((5, 1)), _spd((5, 3)), (np.diag(np.linspace(1, 2, 5)), np.zeros(5)),
                 _spd_rhs_0sol0(), _sym_indef()]


# reference tests/test_cg.py, test_minres.py, test_gmres.py (real-valued problems)
@pytest.mark.parametrize("idx", range(len(REAL_PROBLEMS)))
@pytest.mark.parametrize("solver", ["cg", "minres", "gmres", "gmres_mgs2"])
def test_reference_test_problems(solver, idx):
    A, b = REAL_PROBLEMS[idx]
    count = 0

    def callback(x, r):
        nonlocal count
        count += 1

    kw = {"ortho": "mgs2"} if solver == "gmres_mgs2" else {}
    fn = kb.gmres if solver.startswith("gmres") else getattr(kb, solver)
    sol, info = fn(A, b, tol=1.0e-7, callback=callback, **kw)
    assert count == info.numsteps + 1
    assert info.success
    _consistent(A, b, info, sol, 1.0e-7)


def test_gmres_unsymmetric_and_householder():
    A, b = _unsym()
    for ortho in ("mgs", "mgs2", "householder"):
        sol, info = kb.gmres(A, b, tol=1e-7, ortho=ortho)
        assert info.success
        _consistent(A, b, info, sol, 1e-7)
    with pytest.raises(AssertionError):
        kb.gmres(A, np.ones((5, 3)), ortho="householder")


# reference tests/test_solvers.py:80-87, 199-243
@pytest.mark.parametrize("solver", ["cg", "minres", "gmres"])
def test_operator_kinds_and_exact_x0(solver):
    import scipy.sparse
    import scipy.sparse.linalg

    fn = getattr(kb, solver)
    A = np.diag([1.0e-3] + list(range(2, 11))).astype(float)
    b = np.ones(10)
    _, info = fn(A, b, x0=np.linalg.solve(A, b))
    assert len(info.resnorms) == 1
    n = 5
    a = np.linspace(1.0, 2.0, n)
    a[-1] = 1e-2
    b = np.ones(n)
    _, info = fn(scipy.sparse.spdiags(a, [0], n, n), b, tol=1e-12)
    assert info.resnorms[-1] <= 1e-12
    _, info = fn(scipy.sparse.linalg.LinearOperator((n, n), lambda x: a * x), b, tol=1e-12)
    assert info.resnorms[-1] <= 1e-12

    class MyOp:
        shape = (n, n)
        dtype = float

        def __matmul__(self, x):
            return a * x

    _, info = fn(MyOp(), b, tol=1e-12)
    assert info.resnorms[-1] <= 1e-12
    # torch in -> torch out, device-resident
    At = kb.CsrMatrix.from_dense(np.diag(a))
    bt = torch.ones(n, dtype=torch.float64, device="cuda")
    sol, info = fn(At, bt, tol=1e-12)
    assert isinstance(sol, torch.Tensor) and sol.is_cuda and sol.data_ptr() == info.xk.data_ptr()
    np.testing.assert_allclose(sol.cpu().numpy(), 1.0 / a, rtol=1e-10)
    with pytest.raises(AssertionError):
        fn(np.eye(4), np.ones(5))
    with pytest.raises(NotImplementedError):
        fn(np.eye(3) * 1j, np.ones(3))


def test_cg_return_arnoldi():
    _, A, b, _ = CASES["p2d32_cg"]
    _, info = kb.cg(A, b, tol=1e-10, maxiter=40, return_arnoldi=True)
    V, H, P = info.arnoldi
    np.testing.assert_allclose(H, SOL["p2d32_cg_arn_H"], rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(V[5], SOL["p2d32_cg_arn_V5"], rtol=0, atol=1e-11)
    assert len(V) == 41 and len(P) == 41


def test_gmres_restart_extension():
    _, A, b, _ = CASES["cd10_gmres_mgs"]
    _, info = kb.gmres(A, b, tol=0.0, atol=0.0, maxiter=40, restart=10)
    assert info.numsteps == 40
    ref = SOL["cd10_gmres_restart10_hist"]
    flat = np.concatenate([ref[0]] + [h[1:] for h in ref[1:]])
    np.testing.assert_allclose(np.asarray(info.resnorms), flat, rtol=1e-8)
    np.testing.assert_allclose(info.xk, SOL["cd10_gmres_restart10_xk"], rtol=0, atol=1e-10)


def _self_noise_envelope(make, extract, ncols):
    """SURVEY.md 7 'oracle self-noise probe': rerun the oracle with only the
    summation order of the inner product changed (np.dot -> pairwise np.sum ->
    exact fsum).  Arnoldi amplifies such rounding differences roughly 2x per
    step, so coefficient parity is judged against 10x this envelope."""
    import math

    base = extract(make(lambda x, y: np.dot(x, y)))
    env = [np.zeros(np.asarray(a).shape[-1]) for a in base]
    for inner in (lambda x, y: np.sum(x * y), lambda x, y: math.fsum(x * y)):
        alt = extract(make(inner))
        for i, (a, b) in enumerate(zip(alt, base)):
            d = np.abs(a - b)
            env[i] = np.maximum(env[i], d.reshape(-1, d.shape[-1]).max(axis=0))
    return [10.0 * e + 1e-13 for e in env]


def _assert_within(actual, desired, col_tol, what):
    d = np.abs(np.asarray(actual) - np.asarray(desired))
    d = d.reshape(-1, d.shape[-1]).max(axis=0)
    assert np.all(d <= col_tol), f"{what}: {d} > {col_tol}"


def _check_arnoldi_relation(A, V, H, tol=1e-12):
    """reference tests/test_arnoldi.py:166-263: A V_m = V_{m+1} H, V orthonormal."""
    m = H.shape[1]
    An = np.linalg.norm(A.toarray(), 2)
    assert np.linalg.norm(A @ V[:, :m] - V @ H, 2) <= tol * An * m
    assert np.linalg.norm(np.eye(m + 1) - V.T @ V, 2) <= 1e-9  # MGS loses orthogonality slowly
    assert np.all(np.tril(H, -2) == 0.0) and np.all(np.diag(H[1:, :]) >= 0.0)


def test_arnoldi_builders_match_reference():
    A, As, v = cases.arnoldi_inputs()
    m = 20

    def drive(arn, lanczos=False):
        if lanczos:
            T, Vs = [], [np.array(arn.v)]
            for _ in range(m):
                vv, h, _p = next(arn)
                T.append(np.array(h))
                Vs.append(np.array(vv))
            return np.array(T).T, np.column_stack(Vs)  # (3, m), (n, m+1)
        H = np.zeros((m + 1, m))
        for k in range(m):
            _, h = next(arn)
            H[: k + 2, k] = h
        return H, np.column_stack(arn.V)

    for nre in (1, 2):
        envs = _self_noise_envelope(
            lambda inner: orc.ArnoldiMGS(A, v.copy(), num_reorthos=nre, inner=inner),
            lambda arn: list(drive(arn)), m + 1)
        H, V = drive(kb.ArnoldiMGS(A, v.copy(), num_reorthos=nre))
        _assert_within(H, ARN[f"mgs{nre}_H"], envs[0][:m], f"mgs{nre} H")
        _assert_within(V, ARN[f"mgs{nre}_V"], envs[1], f"mgs{nre} V")
        np.testing.assert_allclose(H[:, :8], ARN[f"mgs{nre}_H"][:, :8], rtol=0, atol=1e-12)
        _check_arnoldi_relation(A, V, H)
    # Householder: backward stable, no amplification envelope needed beyond rounding
    H, V = drive(kb.ArnoldiHouseholder(A, v.copy()))
    envs = _self_noise_envelope(
        lambda inner: orc.ArnoldiMGS(A, v.copy(), num_reorthos=2, inner=inner),
        lambda arn: list(drive(arn)), m + 1)
    _assert_within(H, ARN["house_H"], envs[0][:m], "householder H")
    _assert_within(V, ARN["house_V"], envs[1], "householder V")
    _check_arnoldi_relation(A, V, H)
    assert np.linalg.norm(np.eye(m + 1) - V.T @ V, 2) <= 1e-13
    # Lanczos
    envs = _self_noise_envelope(
        lambda inner: orc.ArnoldiLanczos(As, v.copy(), inner=inner),
        lambda arn: list(drive(arn, lanczos=True)), m + 1)
    T, Vs = drive(kb.ArnoldiLanczos(As, v.copy()), lanczos=True)
    _assert_within(T, ARN["lanczos_h"].T, envs[0][:m], "lanczos h")
    _assert_within(Vs, ARN["lanczos_V"], envs[1], "lanczos V")
    np.testing.assert_allclose(T[:, :8], ARN["lanczos_h"].T[:, :8], rtol=0, atol=1e-12)
    # invariant subspace -> ArgumentError on the next step (arnoldi.py:168-171)
    arn = kb.ArnoldiMGS(np.diag([1.0, 2.0, 3.0]), np.array([1.0, 0.0, 0.0]))
    next(arn)
    assert arn.is_invariant
    with pytest.raises(kb.ArgumentError):
        next(arn)


def test_coefficients_first_50_steps():
    """BASELINE.json: 'Hessenberg/Lanczos coefficients compared on the first 50 steps'.
    Arnoldi-MGS (x1, x2), Householder and Lanczos on 3-D stencil matrices, 50 steps, against
    the live oracle within 10x its own summation-order noise; CG's Lanczos tridiagonal
    (return_arnoldi) against the oracle's."""
    m = 50
    A = cases.st.convection_diffusion3d(12)
    As = cases.st.shifted_laplace3d(12)
    v = np.random.default_rng(7).standard_normal(A.shape[0])

    def drive(arn, lanczos=False):
        if lanczos:
            T = []
            for _ in range(m):
                _v, h, _p = next(arn)
                T.append(np.array(h))
            return [np.array(T).T]
        H = np.zeros((m + 1, m))
        for k in range(m):
            _, h = next(arn)
            H[: k + 2, k] = h
        return [H]

    for nre in (1, 2):
        env = _self_noise_envelope(
            lambda inner: orc.ArnoldiMGS(A, v.copy(), num_reorthos=nre, inner=inner), drive, m)
        ref = drive(orc.ArnoldiMGS(A, v.copy(), num_reorthos=nre, inner=lambda x, y: np.dot(x, y)))
        got = drive(kb.ArnoldiMGS(A, v.copy(), num_reorthos=nre))
        _assert_within(got[0], ref[0], env[0], f"mgs{nre} H, 50 steps")
    # Householder has no inner-product hook: its noise amplification is probed by perturbing
    # the start vector at the level of one rounding error
    ref = drive(orc.ArnoldiHouseholder(A, v.copy()))
    prng = np.random.default_rng(8)
    henv = np.zeros(m)
    for _ in range(3):
        v2 = v * (1.0 + np.finfo(float).eps * prng.standard_normal(v.shape))
        alt = drive(orc.ArnoldiHouseholder(A, v2))
        henv = np.maximum(henv, np.abs(alt[0] - ref[0]).max(axis=0))
    got = drive(kb.ArnoldiHouseholder(A, v.copy(), max_steps=m))
    _assert_within(got[0], ref[0], 10.0 * henv + 1e-13, "householder H, 50 steps")
    env = _self_noise_envelope(
        lambda inner: orc.ArnoldiLanczos(As, v.copy(), inner=inner),
        lambda arn: drive(arn, lanczos=True), m)
    ref = drive(orc.ArnoldiLanczos(As, v.copy(), inner=lambda x, y: np.dot(x, y)), lanczos=True)
    got = drive(kb.ArnoldiLanczos(As, v.copy()), lanczos=True)
    _assert_within(got[0], ref[0], env[0], "lanczos h, 50 steps")
    # CG's tridiagonal on the SPD Poisson matrix
    Ap = cases.st.poisson3d(12)
    b = Ap @ v
    _, info = kb.cg(Ap, b, tol=0.0, atol=0.0, maxiter=m, return_arnoldi=True)
    _, io = orc.cg(Ap, b, tol=0.0, atol=0.0, maxiter=m, return_arnoldi=True)
    H, Ho = info.arnoldi[1], io.arnoldi[1]
    assert H.shape == Ho.shape == (m + 1, m)
    np.testing.assert_allclose(H, Ho, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(H[:, :25], Ho[:, :25], rtol=1e-10, atol=1e-12)


def test_givens_bit_exact():
    fg = SML["givens_fg"]
    Gm, r = kb.givens(np.ascontiguousarray(fg.T))
    np.testing.assert_array_equal(Gm[0, 0], SML["givens_csr"][:, 0])
    np.testing.assert_array_equal(Gm[0, 1], SML["givens_csr"][:, 1])
    np.testing.assert_array_equal(r, SML["givens_csr"][:, 2])
    np.testing.assert_array_equal(Gm[1, 0], -SML["givens_csr"][:, 1])
    rng = np.random.default_rng(5)
    X = rng.standard_normal((2, 4000)) * 10.0 ** rng.integers(-150, 150, (2, 4000))
    Gm, r = kb.givens(X)
    ref = np.array([orc.lartg_f64(f, g) for f, g in X.T])
    np.testing.assert_array_equal(Gm[0, 0], ref[:, 0])
    np.testing.assert_array_equal(Gm[0, 1], ref[:, 1])
    np.testing.assert_array_equal(r, ref[:, 2])
    Gb, rb = kb.givens(np.array([[1.0, 0.0, 3.0], [2.0, 5.0, -4.0]]))
    np.testing.assert_array_equal(Gb, SML["givens_block_G"])
    np.testing.assert_array_equal(rb, SML["givens_block_r"])


def test_householder_matches_reference():
    for i, x in enumerate(cases.householder_inputs()):
        H = kb.Householder(x.copy())
        np.testing.assert_allclose(H.v, SML[f"house{i}_v"], rtol=0, atol=2e-15)
        np.testing.assert_allclose(
            np.array([H.alpha, H.beta, H.xnorm], dtype=float), SML[f"house{i}_abx"], rtol=4e-16)
        np.testing.assert_allclose(H @ x, SML[f"house{i}_Hx"], rtol=0, atol=4e-15)
    with pytest.raises(AssertionError):
        kb.Householder(np.ones((4, 2)))


@pytest.mark.parametrize("ortho", ["cgs", "cgs2"])
@pytest.mark.parametrize("k", [1, 4])
def test_gmres_classical_gram_schmidt_extension(ortho, k):
    """ortho="cgs"/"cgs2" (additive: not in the reference, so the oracle statement of it is
    unpinned).  Checked (1) against the oracle's classical variant, (2) against the PINNED
    modified variant with the same number of passes -- same Krylov space, so the same residual
    history up to the loss of orthogonality -- and (3) through the true residual."""
    A = cases.st.convection_diffusion3d(10)
    N = A.shape[0]
    _, b = cases.rhs(A, (N,) if k == 1 else (N, k))
    kw = dict(tol=1e-9, maxiter=120)
    sol, info = kb.gmres(A, b, ortho=ortho, **kw)
    so, io = orc.gmres(A, b, ortho=ortho, **kw)
    sm, im = orc.gmres(A, b, ortho=ortho.replace("cgs", "mgs"), **kw)
    assert info.success and io.success and im.success
    assert abs(info.numsteps - io.numsteps) <= 1 and abs(info.numsteps - im.numsteps) <= 2
    for other, tol in ((io, 1e-8), (im, 1e-6)):
        m = min(info.numsteps, other.numsteps)
        ro = np.asarray(other.resnorms, float)[:m]
        rg = np.asarray(info.resnorms, float)[:m]
        live = ro / ro[0] >= 1e-6
        assert np.all(np.abs(rg - ro)[live] <= tol * ro[live])
    assert np.linalg.norm(sol - so) <= 1e-8 * np.linalg.norm(so)
    assert np.linalg.norm(sol - sm) <= 1e-7 * np.linalg.norm(sm)
    r = b - A @ sol
    assert np.all(np.linalg.norm(r.reshape(N, -1), axis=0)
                  <= 1.0001e-9 * np.linalg.norm(b.reshape(N, -1), axis=0) + 1e-15)
    # preconditioned: two bases (dots against V, subtraction with P)
    M = scipy.sparse.diags(1.0 / (1.0 + np.arange(N) % 3)).tocsr()
    sol, info = kb.gmres(A, b, M=M, ortho=ortho, **kw)
    so, io = orc.gmres(A, b, M=M, ortho=ortho, **kw)
    assert info.success and abs(info.numsteps - io.numsteps) <= 1
    assert np.linalg.norm(sol - so) <= 1e-7 * np.linalg.norm(so)
    with pytest.raises(ValueError):
        kb.gmres(A, b, ortho=ortho, inner=lambda x, y: np.einsum("i...,i...->...", x, y))
