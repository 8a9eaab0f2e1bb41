"""-m gpu kernel-level parity: every C-ABI primitive against NumPy/SciPy on
seeded inputs.  SpMV/SpMM is bit-exact against SciPy's csr_matvec(s) (same
summation order, no FMA); reductions are checked to rounding and for bitwise
run-to-run reproducibility."""
import numpy as np
import pytest
import scipy.sparse
import torch

import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.device import Ops
from krylov_b200.generate import device_stencil7

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(11)


def _mats():
    yield "poisson3d_7", st.poisson3d(7)
    yield "poisson2d_33", st.poisson2d(33)
    yield "convdiff_12", st.convection_diffusion3d(12)
    yield "random_sparse", scipy.sparse.random(3000, 3000, density=0.01, random_state=1, format="csr")
    yield "dense_rows", scipy.sparse.random(600, 600, density=0.6, random_state=2, format="csr")
    yield "empty_rows", scipy.sparse.csr_matrix(
        (np.array([1.0, 2.0]), np.array([0, 3]), np.array([0, 0, 1, 1, 2, 2])), shape=(5, 5))
    yield "one_by_one", scipy.sparse.csr_matrix(np.array([[3.0]]))
    yield "rect", scipy.sparse.random(257, 513, density=0.05, random_state=3, format="csr")


@pytest.mark.parametrize("name,A", list(_mats()))
@pytest.mark.parametrize("k", [1, 2, 3, 16])
def test_spmv_bit_exact_vs_scipy(name, A, k):
    x = rng.standard_normal((A.shape[1], k)) if k > 1 else rng.standard_normal(A.shape[1])
    ref = A @ x
    for sched in ("rowwise", "stream", "pattern", "stencil", "auto"):
        Ad = kb.CsrMatrix.from_scipy(A)
        try:
            Ad.set_schedule(sched)
        except kb.KrylovB200Error:
            # not stencil-like / not constant-coefficient: the library refuses, CSR schedules stay
            assert sched in ("pattern", "stencil")
            continue
        if Ad.info()["schedule"] == "merge" and k == 1:
            # long / skewed rows: several lanes per row, a fixed tree instead of SciPy's
            # left-to-right order (test_merge_schedule_* below)
            np.testing.assert_allclose(Ad @ x, ref, rtol=0, atol=1e-13 * np.max(abs(A) @ abs(x)))
            continue
        np.testing.assert_array_equal(Ad @ x, ref)


@pytest.mark.parametrize("cfg", [-1, 0, 1, 2, 3, 4, 5])
def test_pattern_kernel_variants_bit_exact(cfg):
    """Windowed (TMA x windows, several tile configurations) and gather variants
    of the offset-pattern kernel; odd column counts and misaligned x fall back
    to the gather variant.  All bit-identical to SciPy."""
    from krylov_b200._lib import lib

    lib.kb_tune(4, cfg)
    try:
        mats = [st.poisson3d(8), st.poisson3d(7), st.poisson2d(64), st.poisson2d(33),
                st.convection_diffusion3d(12), st.to_scipy(st.stencil7_csr(40, 6, 4)),
                st.to_scipy(st.stencil7_csr(2, 2, 300)), st.to_scipy(st.stencil5_csr(700, 3))]
        for A in mats:
            Ad = kb.CsrMatrix.from_scipy(A)
            assert Ad.info()["schedule"] in ("pattern", "stencil")
            Ad.set_schedule("pattern")
            assert Ad.info()["schedule"] == "pattern"
            n = A.shape[0]
            x = rng.standard_normal(n)
            np.testing.assert_array_equal(Ad @ x, A @ x)
            # x only 8-byte aligned
            big = torch.from_numpy(np.concatenate([[0.0], x])).cuda()
            y = Ad.matvec_device(big[1:])
            np.testing.assert_array_equal(y.cpu().numpy(), A @ x)
            # fused epilogues through the same kernels
            ops = Ops(n, 1)
            z = rng.standard_normal(n)
            xd, zd = torch.from_numpy(x).cuda().reshape(n, 1), torch.from_numpy(z).cuda().reshape(n, 1)
            yd = torch.empty_like(xd)
            out = ops.slots(1)[0]
            ops.spmv(Ad, xd, yd, mode=2, z=zd, dot=2, out=out)
            ref = z - A @ x
            np.testing.assert_array_equal(yd.cpu().numpy().ravel(), ref)
            np.testing.assert_allclose(out.cpu().numpy()[0], ref @ ref, rtol=1e-13)
            ops.spmv(Ad, xd, yd, dot=1, w=xd, out=out)
            np.testing.assert_allclose(out.cpu().numpy()[0], x @ (A @ x), rtol=1e-12)
    finally:
        lib.kb_tune(4, 0)


@pytest.mark.parametrize("cfg", [0, 2, 3])
@pytest.mark.parametrize("block", [1, 2, 3])
def test_cache_blocked_tile_order_bit_exact(cfg, block):
    """The optional cache-blocked visiting order of the windowed kernel (kb_tune 8; off by
    default because it measured slower) is a permutation of the tiles: same bits."""
    from krylov_b200._lib import lib

    lib.kb_tune(4, cfg)
    lib.kb_tune(8, block)
    try:
        for (nx, ny, nz) in ((32, 32, 5), (64, 16, 3), (32, 16, 7)):
            A = st.to_scipy(st.stencil7_csr(nx, ny, nz, coeffs=st.convdiff_coeffs()))
            Ad = kb.CsrMatrix.from_scipy(A).set_schedule("pattern")
            assert Ad.info()["schedule"] == "pattern"
            n = A.shape[0]
            x, z = rng.standard_normal(n), rng.standard_normal(n)
            np.testing.assert_array_equal(Ad @ x, A @ x)
            ops = Ops(n, 1)
            xd, zd = (torch.from_numpy(a).cuda().reshape(n, 1) for a in (x, z))
            yd = torch.empty_like(xd)
            out = ops.slots(1)[0]
            ops.spmv(Ad, xd, yd, mode=2, z=zd, dot=2, out=out)
            ref = z - A @ x
            np.testing.assert_array_equal(yd.cpu().numpy().ravel(), ref)
            np.testing.assert_allclose(out.cpu().numpy()[0], ref @ ref, rtol=1e-13)
    finally:
        lib.kb_tune(4, 0)
        lib.kb_tune(8, 0)


@pytest.mark.parametrize("cfg", [0, 1])
@pytest.mark.parametrize("k", [2, 4, 16, 64])
def test_windowed_spmm_variants_bit_exact(cfg, k):
    """The TMA-window SpMM for blocked right-hand sides (kb_tune 7; row-wise is the
    default because it measured no slower) is bit-identical to SciPy as well."""
    from krylov_b200._lib import lib

    lib.kb_tune(7, cfg)
    try:
        for A in (st.poisson3d(9), st.convection_diffusion3d(10), st.poisson2d(31),
                  st.to_scipy(st.stencil7_csr(3, 2, 150))):
            n = A.shape[0]
            Ad = kb.CsrMatrix.from_scipy(A)
            X, Z, W = (rng.standard_normal((n, k)) for _ in range(3))
            np.testing.assert_array_equal(Ad @ X, A @ X)
            ops = Ops(n, k)
            x, z, w = (torch.from_numpy(a).cuda() for a in (X, Z, W))
            cf = torch.from_numpy(rng.standard_normal(k)).cuda()
            y = torch.empty_like(x)
            out = ops.slots(1)[0]
            ops.spmv(Ad, x, y, mode=1, z=z, coef=cf, dot=1, w=w, out=out)
            ref = A @ X - cf.cpu().numpy() * Z
            np.testing.assert_array_equal(y.cpu().numpy(), ref)
            np.testing.assert_allclose(out.cpu().numpy(), np.einsum("ij,ij->j", W, ref), rtol=1e-12,
                                       atol=1e-12)
            ops.spmv(Ad, x, y, mode=2, z=z, dot=2, out=out)
            ref = Z - A @ X
            np.testing.assert_array_equal(y.cpu().numpy(), ref)
            np.testing.assert_allclose(out.cpu().numpy(), np.einsum("ij,ij->j", ref, ref), rtol=1e-12)
    finally:
        lib.kb_tune(7, -1)


def test_pattern_schedule_selection():
    """Offset-pattern compression is chosen for stencil-like matrices only, and the value
    stream is dropped only when every diagonal is bitwise constant."""
    assert kb.CsrMatrix.from_scipy(st.poisson3d(9)).info()["schedule"] == "stencil"
    assert kb.CsrMatrix.from_scipy(st.poisson2d(40)).info()["schedule"] == "stencil"
    assert kb.CsrMatrix.from_scipy(st.convection_diffusion3d(8)).info()["schedule"] == "stencil"
    assert kb.CsrMatrix.from_scipy(st.shifted_laplace3d(8)).info()["schedule"] == "stencil"
    # one value off by one ulp, or a signed zero: variable coefficients -> values are streamed
    for pos, val in ((17, np.nextafter(-1.0, 0.0)), (401, None)):
        P = st.poisson3d(9).tocsr()
        P.data[pos] = np.nextafter(P.data[pos], 0.0) if val is None else val
        Pd = kb.CsrMatrix.from_scipy(P)
        assert Pd.info()["schedule"] == "pattern"
        with pytest.raises(kb.KrylovB200Error):
            Pd.set_schedule("stencil")
        xx = rng.standard_normal(P.shape[0])
        np.testing.assert_array_equal(Pd @ xx, P @ xx)
    R = scipy.sparse.random(3000, 3000, density=0.003, random_state=1, format="csr")
    assert kb.CsrMatrix.from_scipy(R).info()["schedule"] == "stream"
    # 27 diagonals (> 16): stays on the CSR stream kernel
    n = 2000
    B = scipy.sparse.diags([np.ones(n - abs(o)) for o in range(-13, 14)], list(range(-13, 14)), format="csr")
    assert kb.CsrMatrix.from_scipy(B).info()["schedule"] == "stream"
    # 12 diagonals: pattern kernel with the wide (16) mask path; unsorted columns -> refused
    C = scipy.sparse.diags([np.arange(1.0, n - abs(o) + 1) for o in range(-6, 6)], list(range(-6, 6)), format="csr")
    Cd = kb.CsrMatrix.from_scipy(C)
    assert Cd.info()["schedule"] == "pattern"
    x = rng.standard_normal(n)
    np.testing.assert_array_equal(Cd @ x, C @ x)
    U = st.poisson2d(12).tocsr()
    U.indices[0:2] = U.indices[0:2][::-1].copy()
    U.data[0:2] = U.data[0:2][::-1].copy()
    Ud = kb.CsrMatrix.from_scipy(U)
    assert Ud.info()["schedule"] == "stream"
    x = rng.standard_normal(U.shape[0])
    np.testing.assert_array_equal(Ud @ x, U @ x)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11])
def test_stencil_kernel_variants_bit_exact(cfg):
    """Constant-diagonal ("stencil") schedule: no index, value or row-pointer stream -- the
    coefficients are kernel parameters.  Every tile configuration, ragged grids, fused
    epilogues, misaligned x (falls back to the kernels that stream values): bit-identical to
    SciPy and to the "pattern" / "stream" schedules."""
    from krylov_b200._lib import lib

    lib.kb_tune(10, cfg)
    try:
        mats = [st.poisson3d(8), st.poisson3d(7), st.poisson2d(64), st.poisson2d(33),
                st.convection_diffusion3d(12), st.shifted_laplace3d(9),
                st.to_scipy(st.stencil7_csr(40, 6, 4)), st.to_scipy(st.stencil7_csr(2, 2, 300)),
                st.to_scipy(st.stencil5_csr(700, 3)),
                st.to_scipy(st.stencil7_csr(64, 16, 9, coeffs=st.convdiff_coeffs())),
                # more tiles than resident CTAs: every CTA loops (mask prefetch, stage ring)
                st.to_scipy(st.stencil7_csr(96, 80, 70)), st.to_scipy(st.stencil5_csr(1000, 611))]
        for A in mats:
            Ad = kb.CsrMatrix.from_scipy(A)
            assert Ad.info()["schedule"] == "stencil"
            n = A.shape[0]
            x = rng.standard_normal(n)
            ref = A @ x
            np.testing.assert_array_equal(Ad @ x, ref)
            big = torch.from_numpy(np.concatenate([[0.0], x])).cuda()
            np.testing.assert_array_equal(Ad.matvec_device(big[1:]).cpu().numpy(), ref)
            ops = Ops(n, 1)
            z, w = rng.standard_normal(n), rng.standard_normal(n)
            xd, zd, wd = (torch.from_numpy(a).cuda().reshape(n, 1) for a in (x, z, w))
            cf = torch.tensor([0.37], dtype=torch.float64, device="cuda")
            yd = torch.empty_like(xd)
            out = ops.slots(2)
            for mode, r in ((0, ref), (1, ref - 0.37 * z), (2, z - ref)):
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=1, w=wd, out=out[0])
                np.testing.assert_array_equal(yd.cpu().numpy().ravel(), r)
                np.testing.assert_allclose(out[0].cpu().numpy()[0], w @ r, rtol=1e-12, atol=1e-12)
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=2, out=out[1])
                np.testing.assert_allclose(out[1].cpu().numpy()[0], r @ r, rtol=1e-13)
                # <x, y>: the dot operand aliases x (CG's <p, Ap>; read from the centre window)
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=1, w=xd, out=out[0])
                np.testing.assert_array_equal(yd.cpu().numpy().ravel(), r)
                np.testing.assert_allclose(out[0].cpu().numpy()[0], x @ r, rtol=1e-12, atol=1e-12)
            # same bits as the schedules that stream the values
            for other in ("pattern", "stream", "rowwise"):
                Ao = kb.CsrMatrix.from_scipy(A).set_schedule(other)
                assert torch.equal(Ao.matvec_device(xd), Ad.matvec_device(xd))
    finally:
        lib.kb_tune(10, 0)


@pytest.mark.parametrize("sched", ["rowwise", "stream", "pattern", "stencil"])
@pytest.mark.parametrize("k", [1, 4])
def test_spmv_fused_modes(sched, k):
    A = st.convection_diffusion3d(9)
    n = A.shape[0]
    Ad = kb.CsrMatrix.from_scipy(A).set_schedule(sched)
    ops = Ops(n, k)
    X, Z, W = (rng.standard_normal((n, k)) for _ in range(3))
    coef = rng.standard_normal(k)
    x, z, w = (torch.from_numpy(a).cuda() for a in (X, Z, W))
    cf = torch.from_numpy(coef).cuda()
    y = torch.empty_like(x)
    out = ops.slots(1)[0]
    t = A @ X
    for mode, ref in ((0, t), (1, t - coef * Z), (2, Z - t)):
        for dot, dref in ((0, None), (1, np.einsum("ij,ij->j", W, ref)), (2, np.einsum("ij,ij->j", ref, ref))):
            ops.spmv(Ad, x, y, mode=mode, z=z, coef=cf, dot=dot, w=w, out=out)
            np.testing.assert_array_equal(y.cpu().numpy(), ref)
            if dot:
                np.testing.assert_allclose(out.cpu().numpy(), dref, rtol=1e-13)


@pytest.mark.parametrize("k", [1, 3, 16, 100])
@pytest.mark.parametrize("n", [1, 255, 4099, 300001])
def test_dot_and_determinism(n, k):
    if n * k > 40_000_000:
        pytest.skip("size")
    ops = Ops(n, k)
    X, Y = rng.standard_normal((n, k)), rng.standard_normal((n, k))
    x, y = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    out = ops.slots(2)
    ops.dot(x, y, out[0])
    ops.dot(x, y, out[1])
    o = out.cpu().numpy()
    np.testing.assert_array_equal(o[0], o[1])  # bitwise repeatable
    ref = np.einsum("ij,ij->j", X, Y)
    np.testing.assert_allclose(o[0], ref, rtol=1e-12, atol=1e-12 * np.sqrt(n))


@pytest.mark.parametrize("k", [1, 5])
def test_vector_kernels(k):
    n = 10007
    ops = Ops(n, k)
    a = {nm: rng.standard_normal((n, k)) for nm in "xprAw"}
    d = {nm: torch.from_numpy(v).cuda() for nm, v in a.items()}
    rho, pAp, c1 = rng.standard_normal(k) ** 2 + 0.1, rng.standard_normal(k) ** 2 + 0.1, rng.standard_normal(k)
    rho_d, pAp_d, c1_d = (torch.from_numpy(v).cuda() for v in (rho, pAp, c1))
    out = ops.slots(1)[0]
    # cg_update_xr (cg.py:185-209)
    alpha = rho / pAp
    x_ref = a["x"] + alpha * a["p"]
    r_ref = a["r"] - alpha * a["A"]
    ops.cg_update_xr(rho_d, pAp_d, None, d["p"], d["A"], d["x"], d["r"], out)
    np.testing.assert_array_equal(d["x"].cpu().numpy(), x_ref)
    np.testing.assert_array_equal(d["r"].cpu().numpy(), r_ref)
    np.testing.assert_allclose(out.cpu().numpy(), np.einsum("ij,ij->j", r_ref, r_ref), rtol=1e-13)
    # zero-division guard (cg.py:185): pAp == 0 -> alpha = rho
    z0 = torch.zeros(k, dtype=torch.float64, device="cuda")
    xb = d["x"].clone()
    ops.cg_update_xr(rho_d, z0, None, d["p"], d["A"], xb, d["r"].clone(), out)
    np.testing.assert_array_equal(xb.cpu().numpy(), x_ref + rho * a["p"])
    # cg_update_p (cg.py:175-178)
    p_ref = r_ref + (rho / pAp) * a["p"]
    ops.cg_update_p(rho_d, pAp_d, d["r"], d["p"])
    np.testing.assert_array_equal(d["p"].cpu().numpy(), p_ref)
    # deferred x update fused into the p update (what = 5) and the flush (what = 4)
    x2 = torch.from_numpy(a["x"]).cuda()
    p2 = torch.from_numpy(a["p"]).cuda()
    r2 = torch.from_numpy(r_ref).cuda()
    ops.cg_update_p(rho_d, pAp_d, r2, p2, x=x2, alpha=c1_d)
    np.testing.assert_array_equal(x2.cpu().numpy(), a["x"] + c1 * a["p"])
    np.testing.assert_array_equal(p2.cpu().numpy(), p_ref)
    x3 = torch.from_numpy(a["x"]).cuda()
    al_d = torch.from_numpy(alpha).cuda()
    ops.cg_flush_x(al_d, torch.from_numpy(a["p"]).cuda(), x3)
    np.testing.assert_array_equal(x3.cpu().numpy(), x_ref)
    r3 = torch.from_numpy(a["r"]).cuda()
    al_out = torch.zeros(k, dtype=torch.float64, device="cuda")
    ops.cg_update_xr(rho_d, pAp_d, None, None, d["A"], None, r3, out, alpha_out=al_out)  # r-only
    np.testing.assert_array_equal(r3.cpu().numpy(), r_ref)
    np.testing.assert_array_equal(al_out.cpu().numpy(), alpha)
    # axpy_dot (arnoldi.py:157-162)
    w_ref = a["w"] - c1 * a["A"]
    ops.axpy_dot(c1_d, d["A"], d["w"], dot=1, z=d["x"], out=out)
    np.testing.assert_array_equal(d["w"].cpu().numpy(), w_ref)
    np.testing.assert_allclose(out.cpu().numpy(), np.einsum("ij,ij->j", x_ref, w_ref), rtol=1e-12, atol=1e-10)
    # div_scale with zero guard (arnoldi.py:191)
    dd = c1.copy()
    dd[0] = 0.0
    o2 = torch.empty_like(d["w"])
    ops.div_scale(o2, d["w"], torch.from_numpy(dd).cuda())
    np.testing.assert_array_equal(o2.cpu().numpy(), w_ref / np.where(dd != 0, dd, 1.0))


def test_gate_skips_launches():
    n = 1000
    ops = Ops(n, 1)
    x = torch.ones(n, 1, dtype=torch.float64, device="cuda")
    y = torch.ones_like(x)
    one = torch.ones(1, dtype=torch.float64, device="cuda")
    stop = torch.tensor([5], dtype=torch.int32, device="cuda")
    ops.gate(stop, 4)      # 5 <= 4 false -> runs
    ops.axpy(y, one, x)
    ops.gate(stop, 5)      # 5 <= 5 -> skipped
    ops.axpy(y, one, x)
    ops.gate(None, 0)
    assert float(y.sum()) == 2.0 * n


def test_device_generator_equals_host_generator():
    for (nx, ny, nz, zl, zh) in [(9, 7, 5, 0, 5), (6, 6, 6, 2, 4), (1, 1, 3, 0, 3)]:
        Ad = device_stencil7(nx, ny, nz, coeffs=st.convdiff_coeffs(), shift=0.25, z_lo=zl, z_hi=zh)
        Ah = st.to_scipy(st.stencil7_csr(nx, ny, nz, coeffs=st.convdiff_coeffs(), shift=0.25, z_lo=zl, z_hi=zh),
                         n_cols=nx * ny * nz)
        B = Ad.to_scipy()
        np.testing.assert_array_equal(B.indptr, Ah.indptr)
        np.testing.assert_array_equal(B.indices, Ah.indices)
        np.testing.assert_array_equal(B.data, Ah.data)


@pytest.mark.parametrize("k", [1, 3, 16, 64])
@pytest.mark.parametrize("jc", [8, 16])
def test_classical_gram_schmidt_kernels(k, jc):
    """Tall-skinny V^T w (several basis vectors per pass over w) to rounding and bitwise
    reproducible; w -= P h bit-identical to the NumPy statement (rounded product, rounded
    subtraction, j ascending); the fused <w, w> to rounding."""
    from krylov_b200._lib import lib

    lib.kb_tune(9, jc)
    try:
        for n in (1, 777, 50001):
            ops = Ops(n, k)
            for cnt in (0, 1, 2, 5, 8, 9, 17, 33):
                V = rng.standard_normal((max(cnt, 1), n, k))
                w = rng.standard_normal((n, k))
                Vd, wd = torch.from_numpy(V).cuda(), torch.from_numpy(w).cuda()
                out = torch.full((max(cnt, 1) + 1, k), 7.0, dtype=torch.float64, device="cuda")
                ops.multi_dot(cnt, Vd, wd, out)
                got = out.cpu().numpy()
                ref = np.einsum("jnk,nk->jk", V[:cnt], w)
                scale = np.einsum("jnk,nk->jk", np.abs(V[:cnt]), np.abs(w))
                assert np.all(np.abs(got[:cnt] - ref) <= 1e-12 * scale + 1e-300)
                assert np.all(got[cnt:] == 7.0)                   # nothing beyond cnt rows touched
                out2 = torch.zeros_like(out)
                ops.multi_dot(cnt, Vd, wd, out2)
                assert torch.equal(out2[:cnt], out[:cnt])         # run-to-run bitwise
                # w -= sum_j h_j P_j
                h = rng.standard_normal((max(cnt, 1), k))
                hd = torch.from_numpy(h).cuda()
                ww = torch.zeros((1, k), dtype=torch.float64, device="cuda")
                wref = w.copy()
                for j in range(cnt):
                    wref -= h[j] * V[j]
                w1 = wd.clone()
                ops.multi_axpy(cnt, hd, Vd, w1)
                np.testing.assert_array_equal(w1.cpu().numpy(), wref)
                w2 = wd.clone()
                ops.multi_axpy(cnt, hd, Vd, w2, dot=2, out=ww[0])
                np.testing.assert_array_equal(w2.cpu().numpy(), wref)
                nn = np.einsum("nk,nk->k", wref, wref)
                assert np.all(np.abs(ww[0].cpu().numpy() - nn) <= 1e-13 * nn + 1e-300)
            del ops
    finally:
        lib.kb_tune(9, 8)


@pytest.mark.parametrize("mcfg", [0, 1, 2, 3])
def test_marching_stencil_kernel_bit_exact(mcfg):
    """Plane-marching kernel (kb_march.cuh, kb_tune 10 = 10): windows of x kept in a shared-memory
    ring while a CTA walks through the planes.  Partial tiles (plane size not a multiple of the
    tile), short last chunks, more work items than CTAs, every epilogue and dot variant:
    bit-identical to SciPy; matrices it does not cover fall through to the tiled kernel."""
    from krylov_b200._lib import lib

    lib.kb_tune(10, 10)
    lib.kb_tune(14, mcfg)
    try:
        cases = [(st.to_scipy(st.stencil7_csr(40, 30, 9)), 0),          # P = 1200: partial 2nd tile
                 (st.to_scipy(st.stencil7_csr(64, 64, 20)), 3),         # chunks of 3 planes, ragged end
                 (st.to_scipy(st.stencil7_csr(96, 80, 70, coeffs=st.convdiff_coeffs())), 0),
                 (st.to_scipy(st.stencil7_csr(512, 64, 5, shift=0.37)), 1),  # 512-wide lines, halo 512
                 (st.to_scipy(st.stencil7_csr(256, 256, 12)), 2),       # 64*6 = 384.. items, CTAs loop
                 (st.poisson3d(33), 0),                                  # odd plane size: tiled kernel
                 (st.poisson3d(12), 0)]                                  # plane smaller than a tile
        for A, ch in cases:
            lib.kb_tune(13, ch)
            Ad = kb.CsrMatrix.from_scipy(A)
            assert Ad.info()["schedule"] == "stencil"
            n = A.shape[0]
            x, z, w = (rng.standard_normal(n) for _ in range(3))
            ref = A @ x
            np.testing.assert_array_equal(Ad @ x, ref)
            ops = Ops(n, 1)
            xd, zd, wd = (torch.from_numpy(a).cuda().reshape(n, 1) for a in (x, z, w))
            cf = torch.tensor([-1.7], dtype=torch.float64, device="cuda")
            yd = torch.empty_like(xd)
            out = ops.slots(2)
            for mode, r in ((0, ref), (1, ref - (-1.7) * z), (2, z - ref)):
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=1, w=wd, out=out[0])
                np.testing.assert_array_equal(yd.cpu().numpy().ravel(), r)
                np.testing.assert_allclose(out[0].cpu().numpy()[0], w @ r, rtol=1e-12, atol=1e-11)
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=2, out=out[1])
                np.testing.assert_array_equal(yd.cpu().numpy().ravel(), r)
                np.testing.assert_allclose(out[1].cpu().numpy()[0], r @ r, rtol=1e-13)
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=1, w=xd, out=out[0])
                np.testing.assert_array_equal(yd.cpu().numpy().ravel(), r)
                np.testing.assert_allclose(out[0].cpu().numpy()[0], x @ r, rtol=1e-12, atol=1e-11)
                ops.spmv(Ad, xd, yd, mode=mode, z=zd, coef=cf, dot=0)
                np.testing.assert_array_equal(yd.cpu().numpy().ravel(), r)
    finally:
        lib.kb_tune(10, 0)
        lib.kb_tune(13, 0)
        lib.kb_tune(14, 0)


@pytest.mark.parametrize("mcfg,ch", [(0, 0), (1, 5), (2, 0), (3, 2), (0, 1)])
def test_fused_marching_cg_matches_three_kernel_cg(mcfg, ch):
    """kb_cg_run with a second p buffer on a 3-D constant-coefficient stencil: p/x update fused
    with A p and <p, A p>, r update with A p recomputed on chip (two launches per step).  Every
    element-wise statement is the one of the three-kernel path, only the summation order of the
    two dots differs: same step count, histories within 1e-8 while above 1e-6 of the start,
    solution to 1e-10, and both match
    the oracle.  Batches end on converged and unconverged steps (p buffer parity)."""
    from krylov_b200._lib import lib
    from oracle import krylov_oracle as orc

    for A, tol in ((st.to_scipy(st.stencil7_csr(40, 30, 9)), 1e-10),
                   (st.to_scipy(st.stencil7_csr(64, 32, 17, coeffs=st.STENCIL_POISSON, shift=-0.3)), 1e-9),
                   (st.poisson3d(32), 1e-8)):
        n = A.shape[0]
        b = A @ rng.standard_normal(n)
        res = {}
        for fuse in (0, 1):
            lib.kb_tune(15, fuse)
            lib.kb_tune(14, mcfg)
            lib.kb_tune(13, ch)
            try:
                sol, info = kb.cg(A, b, tol=tol, maxiter=3000)
            finally:
                lib.kb_tune(15, 1)
                lib.kb_tune(14, 0)
                lib.kb_tune(13, 0)
            assert info.success
            res[fuse] = (sol, np.asarray(info.resnorms), info.numsteps)
        assert res[0][2] == res[1][2]
        r0, r1 = res[0][1], res[1][1]
        live = r0 / r0[0] >= 1e-6
        assert np.all(np.abs(r1 - r0)[live] <= 1e-8 * r0[live])
        assert np.linalg.norm(res[1][0] - res[0][0]) <= 1e-10 * np.linalg.norm(res[0][0])
        sol_o, info_o = orc.cg(A, b, tol=tol, maxiter=3000)
        assert info_o.numsteps == res[1][2]
        ro = np.asarray(info_o.resnorms)
        live = ro / ro[0] >= 1e-6
        assert np.all(np.abs(res[1][1] - ro)[live] <= 1e-8 * ro[live])
        assert np.linalg.norm(res[1][0] - sol_o) <= 1e-10 * np.linalg.norm(sol_o)


def test_fused_marching_cg_is_selected():
    """The two-launch path is what runs for k = 1 on a 3-D constant stencil, and only there."""
    from krylov_b200.cg import FusedCG

    from krylov_b200._lib import lib

    def fused(A, k=1):
        n = A.shape[0]
        Ad = kb.CsrMatrix.from_scipy(A)
        b = torch.from_numpy(rng.standard_normal((n, k))).cuda()
        lib.kb_tune(28, 0)  # these sizes would otherwise take the persistent kernel
        try:
            s = FusedCG(Ad, b, torch.zeros_like(b), 1e-8, 0.0)
            s.run(4)
        finally:
            lib.kb_tune(28, 262144)
        return s.fused_march and not s.persistent

    assert fused(st.poisson3d(32))
    assert not fused(st.poisson3d(32), k=2)
    assert not fused(st.poisson2d(64))
    assert not fused(st.poisson3d(12))      # plane smaller than a tile
    P = st.poisson3d(32).tocsr()
    P.data[5] = np.nextafter(P.data[5], 0.0)  # variable coefficients
    assert not fused(P)


def test_fused_marching_cg_step_by_step_paths():
    """callback / return_arnoldi drive the fused two-launch path one iteration per call (p buffer
    parity tracked across calls, x flushed for every callback): same results as the three-kernel
    path and as the oracle."""
    from krylov_b200._lib import lib
    from oracle import krylov_oracle as orc

    A = st.poisson3d(32)
    n = A.shape[0]
    b = A @ rng.standard_normal(n)
    out = {}
    for fuse in (0, 1):
        lib.kb_tune(15, fuse)
        try:
            seen = []
            sol, info = kb.cg(A, b, tol=1e-9, maxiter=500,
                              callback=lambda x, r: seen.append((np.linalg.norm(b - A @ x), np.linalg.norm(r))))
            sol2, info2 = kb.cg(A, b, tol=1e-9, maxiter=60, return_arnoldi=True)
        finally:
            lib.kb_tune(15, 1)
        assert info.success and len(seen) == info.numsteps + 1
        # the callback's x is consistent with its r: ||b - A x|| == ||r|| up to rounding
        for true_r, rec_r in seen:
            assert abs(true_r - rec_r) <= 1e-9 * seen[0][0]
        out[fuse] = (sol, np.asarray(info.resnorms), info2.arnoldi[1], np.asarray(info2.resnorms))
    assert len(out[0][1]) == len(out[1][1])
    live = out[0][1] / out[0][1][0] >= 1e-6
    assert np.all(np.abs(out[1][1] - out[0][1])[live] <= 1e-8 * out[0][1][live])
    assert np.linalg.norm(out[1][0] - out[0][0]) <= 1e-10 * np.linalg.norm(out[0][0])
    np.testing.assert_allclose(out[1][2][:41, :40], out[0][2][:41, :40], rtol=0, atol=1e-8 * np.abs(out[0][2]).max())
    sol_o, info_o = orc.cg(A, b, tol=1e-9, maxiter=500)
    assert info_o.numsteps == len(out[1][1]) - 1
    assert np.linalg.norm(out[1][0] - sol_o) <= 1e-10 * np.linalg.norm(sol_o)


@pytest.mark.parametrize("chunk,order", [(0, 1), (1, 1), (0, 0)])
@pytest.mark.parametrize("k", [2, 4, 8, 16, 32])
def test_line_marching_spmm_bit_exact(k, chunk, order):
    """kb_spmm_lines_kernel (blocked right-hand sides on constant 3-D stencils): bit-identical to
    SciPy's csr_matvecs and to the row-wise kernel in every mode, incl. lines shorter / longer than
    a chunk, a truncated last plane, one line per item, and the <x, A x> operand taken on chip."""
    import ctypes

    from krylov_b200._lib import check, lib

    def is_lines(Ad, x):
        yes = ctypes.c_int(0)
        check(lib.kb_spmm_is_lines(Ad.handle, k, x.data_ptr(), ctypes.byref(yes)))
        return bool(yes.value)

    lib.kb_tune(16, 2)  # wherever the geometry is valid (default: only where lines fill their chunks)
    lib.kb_tune(18, chunk)  # 1024- / 512-entry chunks
    lib.kb_tune(19, order)  # work items: planes fastest (default) / natural order
    try:
        for (nx, ny, nz), coeffs, ch in (((70, 5, 4), st.STENCIL_POISSON, 0), ((33, 4, 5), st.convdiff_coeffs(), 3),
                                        ((130, 3, 3), st.convdiff_coeffs(), 1), ((16, 16, 16), st.STENCIL_POISSON, 0)):
            lib.kb_tune(17, ch)
            A = st.to_scipy(st.stencil7_csr(nx, ny, nz, coeffs=coeffs))
            if (nx, ny, nz) == (33, 4, 5):
                A = A[:-7, :-7].tocsr()  # truncated last plane
            n = A.shape[0]
            Ad = kb.CsrMatrix.from_scipy(A)
            assert Ad.info()["schedule"] == "stencil"
            Ar = kb.CsrMatrix.from_scipy(A).set_schedule("rowwise")
            X, Z, W = (rng.standard_normal((n, k)) for _ in range(3))
            coef = rng.standard_normal(k)
            x, z, w = (torch.from_numpy(a).cuda() for a in (X, Z, W))
            cf = torch.from_numpy(coef).cuda()
            assert is_lines(Ad, x) and not is_lines(Ar, x)
            ops = Ops(n, k)
            y, yr = torch.empty_like(x), torch.empty_like(x)
            out, outr = ops.slots(1)[0], ops.slots(1)[0]
            t = A @ X
            for mode, ref in ((0, t), (1, t - coef * Z), (2, Z - t)):
                for dot, ww in ((0, w), (1, w), (1, x), (2, w)):
                    y.fill_(float("nan"))
                    ops.spmv(Ad, x, y, mode=mode, z=z, coef=cf, dot=dot, w=ww, out=out)
                    ops.spmv(Ar, x, yr, mode=mode, z=z, coef=cf, dot=dot, w=ww, out=outr)
                    np.testing.assert_array_equal(y.cpu().numpy(), ref)
                    assert torch.equal(y, yr)
                    if dot:
                        Wn = X if ww is x else W
                        dref = np.einsum("ij,ij->j", Wn if dot == 1 else ref, ref)
                        np.testing.assert_allclose(out.cpu().numpy(), dref, rtol=1e-12, atol=1e-11)
                        np.testing.assert_allclose(out.cpu().numpy(), outr.cpu().numpy(), rtol=1e-12, atol=1e-11)
            # an unaligned operand falls back to the row-wise kernel, same bits
            xo = torch.empty(n * k + 1, dtype=torch.float64, device="cuda")[1:].reshape(n, k)
            xo.copy_(x)
            assert not is_lines(Ad, xo)
        # blocked CG end to end: same history as with the row-wise product
        A = st.poisson3d(20)
        B = rng.standard_normal((A.shape[0], k))
        lib.kb_tune(17, 0)
        s1, i1 = kb.cg(A, B, tol=1e-10, maxiter=300)
        lib.kb_tune(16, 0)
        s0, i0 = kb.cg(A, B, tol=1e-10, maxiter=300)
        assert i1.success and i0.success and abs(i1.numsteps - i0.numsteps) <= 1
        r1, r0 = np.asarray(i1.resnorms), np.asarray(i0.resnorms)
        m = min(len(r1), len(r0))
        live = r0[:m] / r0[0] >= 1e-6
        np.testing.assert_allclose(r1[:m][live], r0[:m][live], rtol=1e-8)
        np.testing.assert_allclose(s1, s0, rtol=0, atol=1e-8)
    finally:
        lib.kb_tune(16, 1)
        lib.kb_tune(17, 0)
        lib.kb_tune(18, 0)
        lib.kb_tune(19, 1)


def test_line_marching_spmm_default_selection():
    """default rule: the kernel runs where a line fills at least half of its chunks (C4: 256 x 16)"""
    import ctypes

    from krylov_b200._lib import check, lib

    for n1, k, want in ((64, 16, True), (64, 8, True), (16, 16, False), (64, 3, False), (40, 16, True)):
        Ad = device_stencil7(n1, 8, 8)
        x = torch.zeros((Ad.shape[0], k), dtype=torch.float64, device="cuda")
        yes = ctypes.c_int(0)
        check(lib.kb_spmm_is_lines(Ad.handle, k, x.data_ptr(), ctypes.byref(yes)))
        assert bool(yes.value) == want, (n1, k)


@pytest.mark.parametrize("shape", [(512, 512, 6), (200, 160, 9), (96, 64, 20)])
def test_fused_cg_tile_shapes_agree(shape):
    """The fused CG kernels in every tile shape (kb_tune 22 / 23: 1024-, 896-, 768-, 512-row
    tiles, ring of 4 / 5) run the same arithmetic: identical row sums, dots that differ only in
    the order of the block partials."""
    import torch

    from krylov_b200._lib import lib
    from krylov_b200.cg import FusedCG
    from krylov_b200.generate import device_stencil7

    nx, ny, nz = shape
    A = device_stencil7(nx, ny, nz)
    n = A.shape[0]
    g = torch.Generator(device="cuda").manual_seed(3)
    b = A.matvec_device(torch.randn(n, 1, generator=g, dtype=torch.float64, device="cuda"))
    x0 = torch.zeros_like(b)
    ref = None
    try:
        for c1, c2 in ((0, 0), (4, 5), (4, 4), (5, 5), (2, 2), (1, 1), (-1, -1)):
            lib.kb_tune(22, c1)
            lib.kb_tune(23, c2)
            st = FusedCG(A, b, x0, 0.0, 0.0)
            hist = np.asarray(st.run(8) + st.run(17)).reshape(-1)
            x = st.current_x()
            if not st.fused_march:
                pytest.skip("grid too small for the marching kernels")
            if ref is None:
                ref = (hist, x)
            else:
                assert np.max(np.abs(hist - ref[0]) / ref[0]) <= 1e-12, (c1, c2)
                assert float(torch.linalg.norm(x - ref[1]) / torch.linalg.norm(ref[1])) <= 1e-12
    finally:
        lib.kb_tune(22, -1)
        lib.kb_tune(23, -1)


# ------------------------------------------------------------------ merge schedule --
def _skewed_mats():
    r = np.random.default_rng(7)
    yield "random_100", scipy.sparse.random(20000, 20000, density=0.005, random_state=4, format="csr")
    lens = np.minimum((r.pareto(1.0, 30000) * 4).astype(np.int64) + 1, 25000)
    rows = np.repeat(np.arange(30000), lens)
    cols = r.integers(0, 30000, size=rows.size)
    P = scipy.sparse.csr_matrix((r.standard_normal(rows.size), (rows, cols)), shape=(30000, 30000))
    P.sum_duplicates()
    yield "powerlaw", P
    # one row of 300 000 entries (147 tiles of 2048) between short rows and empty rows
    M = scipy.sparse.lil_matrix((50, 300000))
    M[7, :] = r.standard_normal(300000)
    for i in (0, 1, 9, 30, 31):
        M[i, r.choice(300000, 5, replace=False)] = 2.5
    yield "one_very_long_row", M.tocsr()
    yield "fem27", st.fem27_var(20)
    yield "empty_rows", scipy.sparse.csr_matrix(
        (np.array([1.0, 2.0]), np.array([0, 3]), np.array([0, 0, 1, 1, 2, 2])), shape=(5, 5))
    yield "poisson3d_20", st.poisson3d(20)


@pytest.mark.parametrize("name,A", list(_skewed_mats()))
@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7])
def test_merge_schedule_vs_scipy(name, A, cfg):
    """Nonzero-balanced tiles (csrc/kb_merge.cuh): every row finished exactly once whatever
    its length, 1e-13 of |A||x| against SciPy (bit-exact where one lane sums a row), fused
    epilogues and dots, bitwise repeatable."""
    from krylov_b200._lib import lib

    A = A.tocsr()
    A.sort_indices()
    lib.kb_tune(25, cfg)
    lib.kb_tune(27, cfg % 3)  # tile -> CTA order
    try:
        n, m = A.shape
        Ad = kb.CsrMatrix.from_scipy(A).set_schedule("merge")
        assert Ad.info()["schedule"] == "merge"
        X, Z, W = r_(m), r_(n), r_(n)
        x, z, w = (torch.from_numpy(a).cuda() for a in (X, Z, W))
        ops = Ops(n, 1)
        coef = np.array([0.7])
        cf = torch.from_numpy(coef).cuda()
        y = torch.empty(n, dtype=torch.float64, device="cuda")
        out = ops.slots(1)[0]
        t = A @ X
        bound = 1e-13 * (abs(A) @ abs(X)) + 1e-300
        first = None
        for mode, ref in ((0, t), (1, t - coef[0] * Z), (2, Z - t)):
            for dot, dref in ((0, None), (1, W @ ref), (2, ref @ ref)):
                y.fill_(float("nan"))
                ops.spmv(Ad, x, y, mode=mode, z=z, coef=cf, dot=dot, w=w, out=out)
                got = y.cpu().numpy()
                assert np.all(np.abs(got - ref) <= bound + 1e-15 * np.abs(ref)), (mode, dot)
                if dot:
                    sc = (np.abs(W) @ np.abs(ref)) if dot == 1 else dref
                    assert abs(out.item() - dref) <= 1e-12 * sc + 1e-300
                if mode == 0 and dot == 0:
                    first = got.copy()
        ops.spmv(Ad, x, y, mode=0, dot=0)
        np.testing.assert_array_equal(y.cpu().numpy(), first)  # timing-independent
    finally:
        lib.kb_tune(25, 0)
        lib.kb_tune(27, 2)


def r_(n):
    return np.random.default_rng(n).standard_normal(n)


def test_merge_schedule_is_bit_exact_on_short_rows():
    """Tiles whose rows get one lane each (more than 128 rows per 2048 nonzeros) sum left to
    right: the 7-point stencil through the merge kernel equals SciPy bit for bit, except the
    rows that straddle more than two tiles (none here)."""
    A = st.poisson3d(24)
    Ad = kb.CsrMatrix.from_scipy(A).set_schedule("merge")
    x = r_(A.shape[0])
    np.testing.assert_array_equal(Ad @ x, A @ x)


def test_merge_schedule_selection_and_solvers():
    A = scipy.sparse.random(4000, 4000, density=0.02, random_state=3, format="csr")  # 80 / row
    A = (A + A.T + 200 * scipy.sparse.identity(4000)).tocsr()
    Ad = kb.CsrMatrix.from_scipy(A)
    assert Ad.info()["schedule"] == "merge"
    b = A @ r_(4000)
    from oracle import krylov_oracle as orc
    for name in ("cg", "minres", "gmres"):
        sol, info = getattr(kb, name)(Ad, b, tol=1e-10)
        sol_o, info_o = getattr(orc, name)(A, b, tol=1e-10)
        assert info.success and abs(info.numsteps - info_o.numsteps) <= 1
        assert np.linalg.norm(sol - sol_o) <= 1e-10 * np.linalg.norm(sol_o)
        ro, rr = np.asarray(info_o.resnorms), np.asarray(info.resnorms)
        m = min(len(ro), len(rr))
        live = ro[:m] / ro[0] >= 1e-6
        assert np.all(np.abs(rr[:m] - ro[:m])[live] <= 1e-8 * ro[:m][live])


# ------------------------------------------------- persistent small-problem CG kernel --
def _small_cases():
    yield "poisson2d_256", st.poisson2d(256), 1e-10          # BASELINE C1
    yield "poisson2d_37", st.poisson2d(37), 1e-12            # n = 1369: one partial CTA
    yield "poisson3d_60", st.poisson3d(60), 1e-10            # n = 216000 > 148 x 1024: 2 rows / thread
    R = scipy.sparse.random(5000, 5000, density=0.004, random_state=5, format="csr")
    yield "random_spd", (R + R.T + 30 * scipy.sparse.identity(5000)).tocsr(), 1e-12
    yield "fem27_14", st.fem27_var(14), 1e-11


@pytest.mark.parametrize("name,A,tol", list(_small_cases()), ids=[c[0] for c in _small_cases()])
def test_persistent_cg_matches_launched_cg_and_oracle(name, A, tol):
    """csrc/kb_small.cu: a whole batch of iterations in one cooperative launch (two grid
    barriers per step).  Same step count as the launched path and as the oracle, histories
    within the north-star tolerance of the oracle and 1e-11 of the launched path, same x."""
    from krylov_b200._lib import lib
    from oracle import krylov_oracle as orc

    b = A @ r_(A.shape[0])
    out = {}
    for mode, key in (("persistent", 262144), ("launched", 0)):
        lib.kb_tune(28, key)
        try:
            Ad = kb.CsrMatrix.from_scipy(A)
            out[mode] = kb.cg(Ad, b, tol=tol, maxiter=5000)
        finally:
            lib.kb_tune(28, 262144)
    sol_o, info_o = orc.cg(A, b, tol=tol, maxiter=5000)
    (xp, ip), (xl, il) = out["persistent"], out["launched"]
    assert ip.success and il.success
    assert ip.numsteps == il.numsteps == info_o.numsteps
    rp, rl, ro = (np.asarray(v.resnorms) for v in (ip, il, info_o))
    live = ro / ro[0] >= 1e-6
    assert np.all(np.abs(rp - ro)[live] <= 1e-8 * ro[live])
    assert np.all(np.abs(rp - rl)[live] <= 1e-11 * rl[live])
    assert np.linalg.norm(xp - sol_o) <= 1e-10 * np.linalg.norm(sol_o)
    assert np.linalg.norm(xp - xl) <= 1e-12 * np.linalg.norm(xl)


def test_persistent_cg_is_selected_and_repeatable():
    import ctypes as C
    from krylov_b200.cg import FusedCG

    A = st.poisson2d(256)
    b = A @ r_(A.shape[0])
    Ad = kb.CsrMatrix.from_scipy(A)
    bd = torch.from_numpy(b.reshape(-1, 1)).cuda()
    stt = FusedCG(Ad, bd, torch.zeros_like(bd), 1e-10, 0.0)
    l0 = stt.ops.launches
    h1 = np.concatenate(stt.run(64))
    assert stt.persistent and stt.ops.launches == l0 + 1  # one launch for 64 steps
    x1, i1 = kb.cg(Ad, b, tol=1e-10, maxiter=5000)
    x2, i2 = kb.cg(Ad, b, tol=1e-10, maxiter=5000)
    np.testing.assert_array_equal(np.asarray(i1.resnorms), np.asarray(i2.resnorms))
    np.testing.assert_array_equal(x1, x2)
    np.testing.assert_array_equal(np.asarray(i1.resnorms)[1:65].ravel(), h1.ravel())


@pytest.mark.parametrize("k", [1, 3])
def test_c_side_loops_equal_per_launch_loops(k):
    """kb_minres_run / kb_gmres_cycle enqueue the same kernels with the same arguments as the
    per-launch Python loops: histories and solutions are bit-identical."""
    import krylov_b200.gmres as gm
    import krylov_b200.minres as mr

    A = st.shifted_laplace3d(14)
    B = st.convection_diffusion3d(13)
    bA = A @ rng.standard_normal((A.shape[0], k) if k > 1 else A.shape[0])
    bB = B @ rng.standard_normal((B.shape[0], k) if k > 1 else B.shape[0])
    res = {}
    for flag in (True, False):
        mr.USE_C_LOOP = gm.USE_C_LOOP = flag
        try:
            res[flag] = (kb.minres(A, bA, tol=1e-9), kb.gmres(B, bB, tol=1e-9, maxiter=90),
                         kb.gmres(B, bB, tol=1e-9, maxiter=90, ortho="mgs2"))
        finally:
            mr.USE_C_LOOP = gm.USE_C_LOOP = True
    for (xc, ic), (xp, ip) in zip(res[True], res[False]):
        assert ic.numsteps == ip.numsteps and ic.success == ip.success
        np.testing.assert_array_equal(np.asarray(ic.resnorms), np.asarray(ip.resnorms))
        np.testing.assert_array_equal(ic.xk, ip.xk)


@pytest.mark.parametrize("chunk,order", [(0, 1), (1, 0)])
@pytest.mark.parametrize("k", [8, 16, 32])
def test_line_marching_spmm_variable_coefficients_bit_exact(k, chunk, order):
    """kb_spmm_lines_kernel<VAR>: the 7-point pattern with VARIABLE coefficients (schedule
    "pattern"), values streamed through the ring slots: bit-identical to SciPy's csr_matvecs and
    to the row-wise kernel in every mode; lines shorter / longer than a chunk, truncated last
    plane, missing entries (rows with fewer than 7 diagonals inside the grid)."""
    import ctypes

    from krylov_b200._lib import check, lib

    def is_lines(Ad, x):
        yes = ctypes.c_int(0)
        check(lib.kb_spmm_is_lines(Ad.handle, k, x.data_ptr(), ctypes.byref(yes)))
        return bool(yes.value)

    lib.kb_tune(16, 2)
    lib.kb_tune(18, chunk)
    lib.kb_tune(19, order)
    try:
        for (nx, ny, nz), ch in (((70, 5, 4), 0), ((33, 4, 5), 3), ((130, 3, 3), 1), ((16, 16, 16), 0)):
            lib.kb_tune(17, ch)
            A = st.to_scipy(st.stencil7_csr(nx, ny, nz))
            A.data = rng.standard_normal(A.nnz)  # same pattern, every value different
            if (nx, ny, nz) == (33, 4, 5):
                A = A[:-7, :-7].tocsr()
            if (nx, ny, nz) == (16, 16, 16):  # knock out some interior entries: masks with holes
                A = A.tolil()
                for r_ in (100, 101, 777, 2000):
                    A[r_, r_ + 1] = 0.0
                    A[r_, r_ - 16] = 0.0
                A = A.tocsr()
                A.eliminate_zeros()
            n = A.shape[0]
            Ad = kb.CsrMatrix.from_scipy(A)
            assert Ad.info()["schedule"] == "pattern"
            Ar = kb.CsrMatrix.from_scipy(A).set_schedule("rowwise")
            X, Z, W = (rng.standard_normal((n, k)) for _ in range(3))
            coef = rng.standard_normal(k)
            x, z, w = (torch.from_numpy(a).cuda() for a in (X, Z, W))
            cf = torch.from_numpy(coef).cuda()
            assert is_lines(Ad, x) and not is_lines(Ar, x)
            ops = Ops(n, k)
            y, yr = torch.empty_like(x), torch.empty_like(x)
            out, outr = ops.slots(1)[0], ops.slots(1)[0]
            t = A @ X
            for mode, ref in ((0, t), (1, t - coef * Z), (2, Z - t)):
                for dot, ww in ((0, w), (1, w), (1, x), (2, w)):
                    y.fill_(float("nan"))
                    ops.spmv(Ad, x, y, mode=mode, z=z, coef=cf, dot=dot, w=ww, out=out)
                    ops.spmv(Ar, x, yr, mode=mode, z=z, coef=cf, dot=dot, w=ww, out=outr)
                    np.testing.assert_array_equal(y.cpu().numpy(), ref)
                    assert torch.equal(y, yr)
                    if dot:
                        Wn = X if ww is x else W
                        dref = np.einsum("ij,ij->j", Wn if dot == 1 else ref, ref)
                        np.testing.assert_allclose(out.cpu().numpy(), dref, rtol=1e-12, atol=1e-10)
    finally:
        lib.kb_tune(16, 1)
        lib.kb_tune(17, 0)
        lib.kb_tune(18, 0)
        lib.kb_tune(19, 1)
