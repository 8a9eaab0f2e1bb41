"""-m gpu parity of krylov_b200.utils through the C ABI: the two DMMA block kernels
against an fp64 torch reference of the same product, and qr / angles / hegedus against the outputs
of the unmodified reference (tests/golden/utils.npz) and the pinned oracle, with the device-resident
inner products and with opaque host callables."""
import os

import numpy as np
import pytest
import torch

import cases_utils as cu
import krylov_b200 as kb
from krylov_b200 import utils as ku
from krylov_b200.device import BlockOps
from oracle import krylov_oracle_utils as ou

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "utils.npz"))
EPS = np.finfo(np.float64).eps


def _rand(n, k, seed, ld=None):
    g = torch.Generator(device="cpu").manual_seed(seed)
    ld = k if ld is None else ld
    full = torch.randn((n, ld), generator=g, dtype=torch.float64).cuda()
    return full, full[:, :k]


# ------------------------------------------------------------------ kernels --
@pytest.mark.parametrize("n,k,l", [(1, 1, 1), (5, 3, 2), (127, 8, 8), (128, 9, 16), (1000, 16, 16),
                                   (4099, 1, 16), (4099, 16, 1), (30011, 7, 13), (300001, 16, 16),
                                   (2049, 20, 5), (2049, 33, 18)])
def test_block_gram_matches_fp64_reference(n, k, l):
    bo = BlockOps()
    _, X = _rand(n, k, 1, ld=k + 3)          # strided rows (a column sub-block)
    _, Y = _rand(n, l, 2)
    G = bo.gram(X, Y)
    ref = X.t() @ Y
    bound = (X.abs().t() @ Y.abs()) * (EPS * (4 + np.log2(max(n, 2)) * 4))
    assert G.shape == (k, l)
    assert torch.all((G - ref).abs() <= bound), float(((G - ref).abs() / bound).max())
    # deterministic summation order: bitwise equal run to run
    assert torch.equal(G, bo.gram(X, Y))
    # accumulate + sqrt(|.|)
    acc = torch.ones((k, l), dtype=torch.float64, device="cuda")
    S = bo.gram(X, Y, acc=acc, sqrt_abs=True)
    assert torch.equal(S, torch.sqrt(torch.abs(G)))
    assert torch.equal(acc, 1.0 + S)


def test_block_gram_single_column_views():
    """columns of a row-major (n, k) array as operands (leading dimension k)"""
    bo = BlockOps()
    full, _ = _rand(5003, 12, 3)
    for i, j in ((0, 0), (3, 11), (11, 4)):
        g = bo.gram(full[:, i:i + 1], full[:, j:j + 1])
        ref = (full[:, i] * full[:, j]).sum()
        assert abs(float(g[0, 0]) - float(ref)) <= 1e-13 * float((full[:, i] * full[:, j]).abs().sum())


@pytest.mark.parametrize("n,k,l", [(1, 1, 1), (9, 3, 2), (128, 8, 8), (1000, 16, 16), (4099, 1, 16),
                                   (4099, 16, 1), (30011, 7, 13), (300001, 16, 16), (2049, 20, 5),
                                   (2049, 33, 18)])
def test_block_apply_matches_fp64_reference(n, k, l):
    bo = BlockOps()
    _, X = _rand(n, k, 4, ld=k + 1)
    _, Y = _rand(n, l, 5, ld=l + 2)
    C = torch.randn((k, l), dtype=torch.float64, generator=torch.Generator().manual_seed(6)).cuda()
    bound = (X.abs() @ C.abs() + Y.abs()) * (EPS * (k + 4))
    for sign, ref in ((0, X @ C), (-1, Y - X @ C), (1, Y + X @ C)):
        Z = bo.apply(X, C, Y=None if sign == 0 else Y, sign=sign)
        assert Z.shape == (n, l)
        assert torch.all((Z - ref).abs() <= bound), (sign, float(((Z - ref).abs() / bound).max()))
    # in place on Y (a strided view): Y <- Y - X C
    Yfull = torch.zeros((n, l + 2), dtype=torch.float64, device="cuda")
    Yv = Yfull[:, :l]
    Yv.copy_(Y)
    bo.apply(X, C, Y=Yv, sign=-1, out=Yv)
    assert torch.all((Yv - (Y - X @ C)).abs() <= bound)
    assert torch.all(Yfull[:, l:] == 0)  # nothing written past the block's columns


def test_block_apply_in_place_on_x():
    bo = BlockOps()
    _, X = _rand(70001, 16, 7)
    C = torch.randn((16, 16), dtype=torch.float64, generator=torch.Generator().manual_seed(8)).cuda()
    ref = X @ C
    Xc = X.clone()
    bo.apply(Xc, C, out=Xc)
    assert torch.all((Xc - ref).abs() <= (X.abs() @ C.abs()) * (EPS * 20))


def test_inner_objects_are_drop_in_callables():
    r = np.random.default_rng(5)
    X, Y = r.standard_normal((3001, 5)), r.standard_normal((3001, 3))
    e = ku.EuclideanInner()
    np.testing.assert_allclose(e(X, Y), X.T @ Y, rtol=0, atol=1e-12)
    assert np.shape(e(X[:, 0], Y[:, 0])) == ()
    w = cu.weight_diag(3001)
    np.testing.assert_allclose(ku.WeightedInner(w)(X, Y), X.T @ (w[:, None] * Y), rtol=0, atol=1e-11)
    Xt = torch.from_numpy(X).cuda()
    out = e(Xt, Xt)
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.shape == (5, 5)


# ----------------------------------------------------------------------- qr --
def _inner(name, n, host_callable):
    if name is None:
        return None
    if host_callable:
        return cu.numpy_inner(name, n)
    return ku.EuclideanInner() if name == "euclid" else ku.WeightedInner(cu.weight_diag(n))


@pytest.mark.parametrize("name", sorted(cu.qr_cases()))
def test_qr_matches_reference(name):
    X, inner, reorthos = cu.qr_cases()[name]
    Q, R = ku.qr(X, inner=_inner(inner, X.shape[0], False), reorthos=reorthos)
    assert Q.shape == GOLD[name + "_Q"].shape and R.shape == GOLD[name + "_R"].shape
    cond = np.linalg.cond(X) if np.linalg.matrix_rank(X) == X.shape[1] else 1.0
    tol = 1e-14 * max(cond, 1.0) * 20  # summation order of the dots, amplified by cond(X)
    np.testing.assert_allclose(R, GOLD[name + "_R"], rtol=0, atol=tol * max(np.abs(X).max(), 1.0))
    np.testing.assert_allclose(Q, GOLD[name + "_Q"], rtol=0, atol=tol)
    # the reference's own assertions (tests/test_utils.py:25-36)
    s = np.linalg.svd(X, compute_uv=False)
    assert np.linalg.norm(Q @ R - X, 2) <= 1e-14 * max(s) * 4
    assert np.linalg.norm(np.tril(R, -1)) == 0
    if inner is not None and name.split("_")[1] not in ("zerocol",):
        ip = cu.numpy_inner(inner, X.shape[0])
        orthotol = 1e-8 if reorthos < 1 else 1e-14 * 4
        assert np.linalg.norm(ip(Q, Q) - np.eye(X.shape[1]), 2) <= orthotol


@pytest.mark.parametrize("name", ["qr_hilbert_weighted_1", "qr_rand7_euclid_1", "qr_eye_weighted_0"])
def test_qr_with_host_callable(name):
    X, inner, reorthos = cu.qr_cases()[name]
    Q, R = ku.qr(X, inner=_inner(inner, X.shape[0], True), reorthos=reorthos)
    cond = np.linalg.cond(X)
    np.testing.assert_allclose(Q, GOLD[name + "_Q"], rtol=0, atol=2e-13 * cond)
    np.testing.assert_allclose(R, GOLD[name + "_R"], rtol=0, atol=2e-13 * cond)


def test_qr_tall_block_against_oracle_and_torch_io():
    n, k = 200003, 12
    X = np.random.default_rng(21).standard_normal((n, k))
    Qr, Rr = ou.qr(X, inner=cu.numpy_inner("euclid", n), reorthos=1)
    Q, R = ku.qr(torch.from_numpy(X).cuda(), inner=ku.EuclideanInner(), reorthos=1)
    assert isinstance(Q, torch.Tensor) and Q.is_cuda
    np.testing.assert_allclose(Q.cpu().numpy(), Qr, rtol=0, atol=1e-13)
    np.testing.assert_allclose(R.cpu().numpy(), Rr, rtol=1e-12, atol=1e-11)
    Ql, Rl = ku.qr(X)  # LAPACK convention
    Qn, Rn = np.linalg.qr(X, mode="reduced")
    np.testing.assert_allclose(Rl, Rn, rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(Ql, Qn, rtol=0, atol=1e-13)


# ------------------------------------------------------------------- angles --
@pytest.mark.parametrize("host_callable", [False, True])
@pytest.mark.parametrize("name", sorted(cu.angles_cases()))
def test_angles_match_reference(name, host_callable):
    F, G, inner = cu.angles_cases()[name]
    n = F.shape[0]
    if host_callable and not name.startswith(("angles_near", "angles_e1_e4", "angles_rand_4_6")):
        pytest.skip("host-callable path sampled on three cases")
    fn = _inner(inner, n, host_callable)
    theta, U, V = ku.angles(F, G, inner=fn, compute_vectors=True)
    np.testing.assert_array_equal(ku.angles(F, G, inner=fn), theta)
    ref = GOLD[name + "_theta"]
    assert theta.shape == ref.shape
    np.testing.assert_allclose(theta, ref, rtol=1e-6, atol=4e-15)
    assert np.all(np.diff(theta) >= 0) and theta[0] >= 0 and theta[-1] <= np.pi / 2
    d = abs(F.shape[1] - G.shape[1])
    if d:
        assert np.all(theta[-d:] == np.pi / 2)
    assert U.shape == F.shape and V.shape == G.shape
    ip = cu.numpy_inner(inner, n)
    assert np.linalg.norm(ip(U, V) - np.diag(np.cos(theta))[: F.shape[1], : G.shape[1]]) <= 1e-13
    assert np.linalg.norm(ip(U, U) - np.eye(F.shape[1])) <= 1e-13
    assert np.linalg.norm(ip(V, V) - np.eye(G.shape[1])) <= 1e-13


def test_angles_default_inner_is_euclidean_and_large_blocks():
    """inner=None (additive extension) == EuclideanInner; 16-column blocks, 2^20 rows; known angles."""
    n, k = 1 << 20, 16
    r = np.random.default_rng(31)
    Qb, _ = np.linalg.qr(r.standard_normal((n, 2 * k)))
    want = np.sort(r.uniform(1e-4, np.pi / 2 - 1e-3, k))
    F = Qb[:, :k]
    G = F * np.cos(want) + Qb[:, k:] * np.sin(want)
    Ft, Gt = torch.from_numpy(F).cuda(), torch.from_numpy(G).cuda()
    theta = ku.angles(Ft, Gt)
    assert isinstance(theta, torch.Tensor)
    np.testing.assert_allclose(theta.cpu().numpy(), want, rtol=1e-8, atol=1e-14)
    theta2, U, V = ku.angles(Ft, Gt, inner=ku.EuclideanInner(), compute_vectors=True)
    assert torch.equal(theta, theta2)
    UV = (U.t() @ V).cpu().numpy()
    assert np.linalg.norm(UV - np.diag(np.cos(want))) <= 1e-12


# ------------------------------------------------------------------ hegedus --
@pytest.mark.parametrize("host_callable", [False, True])
@pytest.mark.parametrize("name", sorted(cu.hegedus_cases()))
def test_hegedus_matches_reference(name, host_callable):
    A, b, x0, M, Ml, inner = cu.hegedus_cases()[name]
    if host_callable and "_lin_" not in name and "poisson" not in name:
        pytest.skip("host-callable path sampled")
    got = ku.hegedus(A, b, x0, M, Ml, _inner(inner, b.shape[0], host_callable))
    ref = GOLD[name + "_x0new"]
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=0)


def test_hegedus_reduces_the_initial_residual_of_cg():
    """the property the reference tests (tests/test_utils.py:117-129), on the device end to end"""
    from krylov_b200 import stencils as st

    A = st.poisson3d(24)
    n = A.shape[0]
    r = np.random.default_rng(41)
    b = A @ r.standard_normal(n)
    x0 = 7.0 * r.standard_normal(n)
    x0n = ku.hegedus(A, b, x0)
    assert np.linalg.norm(b - A @ x0n) <= np.linalg.norm(b - A @ x0) * (1 + 1e-13)
    assert np.linalg.norm(b - A @ x0n) <= np.linalg.norm(b) * (1 + 1e-13)
    _, info0 = kb.cg(A, b, x0=x0, tol=1e-8, maxiter=500)
    _, info1 = kb.cg(A, b, x0=x0n, tol=1e-8, maxiter=500)
    assert info1.resnorms[0] <= info0.resnorms[0]


def test_block_ops_need_wide_workspace():
    from krylov_b200._lib import KrylovB200Error, lib
    from krylov_b200.device import Workspace, cur_stream

    ws = Workspace(16)
    x = torch.ones((64, 16), dtype=torch.float64, device="cuda")
    g = torch.zeros((16, 16), dtype=torch.float64, device="cuda")
    rc = lib.kb_block_gram(ws.handle, 64, 16, 16, x.data_ptr(), 16, x.data_ptr(), 16, g.data_ptr(),
                           16, None, 0, 0, cur_stream())
    assert rc == -1
    with pytest.raises(KrylovB200Error):
        kb._lib.check(rc)
