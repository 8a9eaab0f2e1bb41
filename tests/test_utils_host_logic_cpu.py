"""Host logic of krylov_b200.utils (qr: Gram-Schmidt and LAPACK-convention Householder; angles;
hegedus) against the outputs of the real reference (tests/golden/utils.npz), with the device layer
replaced by tests/fake_device.py.  CPU suite: the kernels behind every statement are covered by
tests/test_gpu_utils.py on the B200."""
import os

import numpy as np
import pytest

import cases_utils as cu
from fake_device import utils_host_logic

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "utils.npz"))


def _inner(ku, name, n, host_callable):
    if name is None:
        return None
    if host_callable:
        return cu.numpy_inner(name, n)
    return ku.EuclideanInner() if name == "euclid" else ku.WeightedInner(cu.weight_diag(n))


@pytest.mark.parametrize("host_callable", [False, True])
@pytest.mark.parametrize("name", sorted(cu.qr_cases()))
def test_qr(name, host_callable):
    X, inner, reorthos = cu.qr_cases()[name]
    with utils_host_logic() as ku:
        Q, R = ku.qr(X, inner=_inner(ku, inner, X.shape[0], host_callable), reorthos=reorthos)
    assert Q.shape == GOLD[name + "_Q"].shape and R.shape == GOLD[name + "_R"].shape
    # ill-conditioned blocks (Hilbert) amplify the summation-order difference of the dots:
    # reference tests/test_utils.py:31-34 asks 1e-14 * sigma_max for the residual, 1e-8 / 1e-14
    # for orthogonality; entries are compared at cond * eps
    cond = np.linalg.cond(X) if np.linalg.matrix_rank(X) == X.shape[1] else 1.0
    tol = 1e-14 * max(cond, 1.0) * 20
    np.testing.assert_allclose(R, GOLD[name + "_R"], rtol=0, atol=tol * max(np.abs(X).max(), 1.0))
    np.testing.assert_allclose(Q, GOLD[name + "_Q"], rtol=0, atol=tol)
    assert np.linalg.norm(np.tril(R, -1)) == 0
    assert np.linalg.norm(Q @ R - X, 2) <= 1e-13 * np.linalg.norm(X, 2)


@pytest.mark.parametrize("host_callable", [False, True])
@pytest.mark.parametrize("name", sorted(cu.angles_cases()))
def test_angles(name, host_callable):
    F, G, inner = cu.angles_cases()[name]
    n = F.shape[0]
    with utils_host_logic() as ku:
        fn = _inner(ku, inner, n, host_callable)
        theta, U, V = ku.angles(F, G, inner=fn, compute_vectors=True)
        theta2 = ku.angles(F, G, inner=fn)
    ref = GOLD[name + "_theta"]
    np.testing.assert_array_equal(theta, theta2)
    np.testing.assert_allclose(theta, ref, rtol=1e-6, atol=2e-15)
    assert U.shape == F.shape and V.shape == G.shape
    # the reference's own assertions (tests/test_utils.py:60-89); U, V themselves are unique only
    # up to rotations inside clusters of equal angles
    ip = cu.numpy_inner(inner, n)
    UV = ip(U, V)
    assert np.linalg.norm(UV - np.diag(np.cos(theta))[: F.shape[1], : G.shape[1]]) <= 1e-13
    assert np.linalg.norm(ip(U, U) - np.eye(F.shape[1])) <= 1e-13


@pytest.mark.parametrize("host_callable", [False, True])
@pytest.mark.parametrize("name", sorted(cu.hegedus_cases()))
def test_hegedus(name, host_callable):
    A, b, x0, M, Ml, inner = cu.hegedus_cases()[name]
    with utils_host_logic() as ku:
        got = ku.hegedus(A, b, x0, M, Ml, _inner(ku, inner, b.shape[0], host_callable))
    ref = GOLD[name + "_x0new"]
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)


def test_small_helpers_and_errors():
    import krylov_b200 as kb
    from krylov_b200 import utils as ku

    np.testing.assert_array_equal(ku.strakos(5), GOLD["strakos_5"])
    np.testing.assert_array_equal(ku.strakos(12, l_min=0.5, l_max=7.0, rho=0.8), GOLD["strakos_12"])
    p = ku.NormalizedRootsPolynomial(GOLD["nrp_roots"])
    np.testing.assert_array_equal(p(GOLD["nrp_pts"]), GOLD["nrp_vals"])
    np.testing.assert_array_equal(p(GOLD["nrp_roots"]), np.zeros(4))  # tests/test_utils.py:150
    assert p(0) == 1
    np.testing.assert_allclose(np.sort_complex(p.minmax_candidates()), GOLD["nrp_cand"])
    assert ku.gap([1, 2], [-4, 3]) == 1 and ku.gap(5, -5) == 10 and ku.gap([-5, 5], -5) == 0
    assert ku.gap(5, -5, mode="interval") == 10 and ku.gap(5, [-5, 6], mode="interval") == 1
    assert ku.gap(-5, [-5, 6], mode="interval") == 0
    assert ku.gap([-5, 5], [0], mode="interval") is None
    with pytest.raises(kb.ArgumentError):
        ku.gap([1j], [1])
    with pytest.raises(kb.ArgumentError):
        ku.NormalizedRootsPolynomial(np.ones((2, 2)))


def test_no_cpu_fallback():
    import torch

    import krylov_b200 as kb

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X = np.eye(6, 3)
    with pytest.raises(kb.KrylovB200Error):
        kb.utils.qr(X)
    with pytest.raises(kb.KrylovB200Error):
        kb.utils.angles(X, X)
    with pytest.raises(kb.KrylovB200Error):
        kb.utils.hegedus(np.eye(6), np.ones(6), np.ones(6))
    with pytest.raises(kb.KrylovB200Error):
        kb.utils.EuclideanInner()(X, X)


def test_edge_shapes():
    """empty blocks, square / wide blocks (LAPACK convention vs np.linalg.qr), a wide block under
    Gram-Schmidt, 1-D vectors, blocks where hegedus is undefined"""
    from oracle import krylov_oracle_utils as ou

    rng = np.random.default_rng(0)
    with utils_host_logic() as ku:
        e = ku.EuclideanInner()
        X0 = np.zeros((10, 0))
        for inner in (e, None):
            Q, R = ku.qr(X0, inner=inner)
            assert Q.shape == (10, 0) and R.shape == (0, 0)
        F = rng.standard_normal((10, 3))
        np.testing.assert_array_equal(ku.angles(F, X0, inner=e), np.full(3, np.pi / 2))
        th, U, V = ku.angles(X0, F, inner=e, compute_vectors=True)
        assert U.shape == (10, 0) and V.shape == (10, 3) and np.all(th == np.pi / 2)
        for shape in ((5, 5), (4, 7), (1, 3), (6, 1)):
            X = rng.standard_normal(shape)
            Q, R = ku.qr(X)
            Qn, Rn = np.linalg.qr(X, mode="reduced")
            assert Q.shape == Qn.shape and R.shape == Rn.shape
            np.testing.assert_allclose(Q, Qn, rtol=0, atol=1e-14)
            np.testing.assert_allclose(R, Rn, rtol=0, atol=1e-14)
        X = rng.standard_normal((3, 5))
        Q, R = ku.qr(X, inner=e)
        Qr, Rr = ou.qr(X, inner=cu.numpy_inner("euclid", 3))
        np.testing.assert_allclose(Q, Qr, rtol=0, atol=1e-14)
        np.testing.assert_allclose(R, Rr, rtol=0, atol=1e-14)
        A, b, x0 = np.diag(np.arange(1.0, 7.0)), np.ones(6), np.linspace(1, 2, 6)
        ref = ou.hegedus(A, b, x0, inner=cu.numpy_inner("euclid", 6))
        for inner in (None, e, lambda x, y: np.dot(x.T.conj(), y)):
            got = ku.hegedus(A, b, x0, inner=inner)
            assert got.shape == (6,)
            np.testing.assert_allclose(got, ref, rtol=1e-14)
        with pytest.raises(ValueError):
            ku.hegedus(A, np.ones((6, 2)), np.ones((6, 2)))
        with pytest.raises(TypeError):
            ku.qr(F, inner=3.0)
