"""-m gpu tests at BASELINE.json's configurations.

Where the oracle finishes in seconds (C1, and shrunk C2-C4) the CUDA path is
compared with it on the same seeded inputs; at the full sizes the checks are
size-independent properties: the recursive residual equals the explicit one,
MINRES/GMRES residuals do not increase, exact scaling by powers of two,
a blocked column equals the single-right-hand-side solve, Arnoldi bases are
orthonormal."""
import numpy as np
import pytest
import torch

import krylov_b200 as kb
from krylov_b200 import stencils as st
from krylov_b200.generate import device_stencil7
from oracle import krylov_oracle as orc

pytestmark = pytest.mark.gpu


def _hist_close(res, ref, rtol=1e-8, floor=1e-6):
    res, ref = np.asarray(res, float), np.asarray(ref, float)
    assert res.shape == ref.shape
    live = ref / ref[0] >= floor
    assert np.all(np.abs(res - ref)[live] <= rtol * ref[live])


def _rhs(A, shape, seed=0):
    xs = np.random.default_rng(seed).standard_normal(shape)
    return xs, A @ xs


def test_c1_cg_poisson2d_256_vs_oracle():
    """configs[0]: the reference's own CPU-runnable case (SURVEY.md appendix A:
    723 steps, resnorms[0] = 1146.911897119499)."""
    A = st.poisson2d(256)
    xs, b = _rhs(A, (A.shape[0],))
    sol, info = kb.cg(A, b, tol=1e-10, maxiter=5000)
    sol_o, info_o = orc.cg(A, b, tol=1e-10, maxiter=5000)
    assert info_o.numsteps == 723
    assert abs(info.numsteps - info_o.numsteps) <= 0.02 * info_o.numsteps
    assert abs(info.resnorms[0] - 1146.911897119499) <= 1e-9
    m = min(info.numsteps, info_o.numsteps) + 1
    _hist_close(info.resnorms[:m], info_o.resnorms[:m])
    assert np.linalg.norm(sol - sol_o) <= 1e-10 * np.linalg.norm(sol_o)


def test_c2_minres_shifted_laplace():
    """configs[1]: oracle parity at 48^3, properties at 128^3 (mild shift:
    exactly one negative eigenvalue)."""
    A = st.shifted_laplace3d(48)
    xs, b = _rhs(A, (A.shape[0],))
    sol, info = kb.minres(A, b, tol=1e-8, maxiter=20000)
    sol_o, info_o = orc.minres(A, b, tol=1e-8, maxiter=20000)
    assert info.success and abs(info.numsteps - info_o.numsteps) <= 0.02 * info_o.numsteps
    m = min(info.numsteps, info_o.numsteps) + 1
    _hist_close(info.resnorms[:m], info_o.resnorms[:m])
    assert np.linalg.norm(sol - sol_o) <= 1e-9 * np.linalg.norm(sol_o)  # cond ~ 1e4 x 1e-13
    # full size, device-resident input
    N = 128
    Ad = device_stencil7(N, N, N, shift=st.mild_shift(N))
    g = torch.Generator(device="cuda").manual_seed(0)
    b = Ad.matvec_device(torch.randn(N ** 3, generator=g, dtype=torch.float64, device="cuda"))
    sol, info = kb.minres(Ad, b, tol=1e-8, maxiter=20000)
    assert info.success
    r = np.asarray(info.resnorms)
    assert np.all(np.diff(r[:-1]) <= 1e-12 * r[0])  # MINRES residuals never increase
    expl = float(torch.linalg.norm(b - Ad.matvec_device(sol)))
    assert abs(expl - r[-1]) <= 1e-10 * r[0]
    assert r[-1] <= 1e-8 * r[0]


@pytest.mark.parametrize("ortho", ["mgs", "mgs2", "householder"])
def test_c3_gmres_convdiff(ortho):
    """configs[2]: one 50-step cycle; oracle parity at 40^3, properties at 256^3."""
    A = st.convection_diffusion3d(40)
    xs, b = _rhs(A, (A.shape[0],))
    sol, info = kb.gmres(A, b, tol=1e-8, maxiter=50, ortho=ortho)
    sol_o, info_o = orc.gmres(A, b, tol=1e-8, maxiter=50, ortho=ortho)
    assert info.numsteps == info_o.numsteps and info.success == info_o.success
    _hist_close(info.resnorms, info_o.resnorms)
    assert np.linalg.norm(info.xk - info_o.xk) <= 1e-10 * np.linalg.norm(info_o.xk)
    N = 256 if ortho != "householder" else 160
    Ad = device_stencil7(N, N, N, coeffs=st.convdiff_coeffs())
    g = torch.Generator(device="cuda").manual_seed(0)
    b = Ad.matvec_device(torch.randn(N ** 3, generator=g, dtype=torch.float64, device="cuda"))
    _, info = kb.gmres(Ad, b, tol=1e-8, maxiter=50, ortho=ortho)
    r = np.asarray(info.resnorms)
    assert info.numsteps == 50 and not info.success
    assert np.all(np.diff(r) <= 1e-12 * r[0])  # GMRES residuals never increase
    expl = float(torch.linalg.norm(b - Ad.matvec_device(info.xk)))
    assert abs(expl - r[-1]) <= 1e-9 * r[0]     # projected == explicit residual
    # restarted continuation keeps decreasing from where the cycle stopped
    _, info2 = kb.gmres(Ad, b, x0=info.xk, tol=1e-8, maxiter=10, ortho=ortho)
    assert abs(info2.resnorms[0] - r[-1]) <= 1e-9 * r[0] and info2.resnorms[-1] < r[-1]


def test_c4_blocked_cg_k16():
    """configs[3]: k = 16 right-hand sides in lock-step (SpMM + column-wise dots)."""
    A = st.poisson3d(20)
    xs, B = _rhs(A, (A.shape[0], 16))
    sol, info = kb.cg(A, B, tol=1e-8, maxiter=2000)
    sol_o, info_o = orc.cg(A, B, tol=1e-8, maxiter=2000)
    assert info.success and info.numsteps == info_o.numsteps
    _hist_close(info.resnorms, info_o.resnorms)
    assert np.linalg.norm(sol - sol_o) <= 1e-10 * np.linalg.norm(sol_o)
    # full size: a blocked column equals the single-RHS solve (columns are independent)
    N = 256
    Ad = device_stencil7(N, N, N)
    g = torch.Generator(device="cuda").manual_seed(0)
    B = torch.randn((N ** 3, 16), generator=g, dtype=torch.float64, device="cuda")
    _, info = kb.cg(Ad, B, tol=0.0, atol=0.0, maxiter=25)
    _, info1 = kb.cg(Ad, B[:, 5].contiguous(), tol=0.0, atol=0.0, maxiter=25)
    rb = np.asarray(info.resnorms)[:, 5]
    r1 = np.asarray(info1.resnorms)
    assert np.all(np.abs(rb - r1) <= 1e-11 * r1)
    d = torch.linalg.norm(info.xk[:, 5] - info1.xk) / torch.linalg.norm(info1.xk)
    assert float(d) <= 1e-12
    expl = torch.linalg.norm(B - Ad.matvec_device(info.xk), dim=0).cpu().numpy()
    assert np.all(np.abs(expl - np.asarray(info.resnorms)[-1]) <= 1e-9 * np.asarray(info.resnorms)[0])


def test_c5_cg_poisson3d_512_properties():
    """configs[4] on one GPU: 134M unknowns, 938M nonzeros."""
    N = 512
    Ad = device_stencil7(N, N, N)
    assert Ad.nnz == 937951232 and Ad.info()["schedule"] == "stencil"
    g = torch.Generator(device="cuda").manual_seed(0)
    b = Ad.matvec_device(torch.randn(N ** 3, generator=g, dtype=torch.float64, device="cuda"))
    _, info = kb.cg(Ad, b, tol=0.0, atol=0.0, maxiter=60)
    r = np.asarray(info.resnorms)
    assert info.numsteps == 60 and np.all(np.isfinite(r))
    expl = float(torch.linalg.norm(b - Ad.matvec_device(info.xk)))
    assert abs(expl - r[-1]) <= 1e-9 * r[0]          # recursive == explicit residual
    # scaling by a power of two is exact in every kernel: bitwise 4x history and solution
    _, info4 = kb.cg(Ad, 4.0 * b, tol=0.0, atol=0.0, maxiter=60)
    np.testing.assert_array_equal(np.asarray(info4.resnorms), 4.0 * r)
    assert torch.equal(info4.xk, 4.0 * info.xk)
    # run-to-run bitwise reproducibility (deterministic reductions)
    _, info_b = kb.cg(Ad, b, tol=0.0, atol=0.0, maxiter=60)
    np.testing.assert_array_equal(np.asarray(info_b.resnorms), r)
    # every SpMV schedule produces the same bits
    x = torch.randn(N ** 3, generator=g, dtype=torch.float64, device="cuda")
    y1 = Ad.matvec_device(x)
    for sched in ("rowwise", "stream", "pattern"):
        Ad.set_schedule(sched)
        assert torch.equal(Ad.matvec_device(x), y1)


# ---------------------------------------------------------------------------------------------
# BASELINE-scale parity against the REAL reference: tests/golden/scale.npz holds the residual
# history and the final iterate (norms + 256 sampled entries) of the unmodified reference after a
# fixed number of steps on the seeded systems of tests/scale_cases.py
# (generated by tests/golden/make_golden_scale.py in the authoring container).
import os

import scale_cases as sc

_SCALE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale.npz")


@pytest.mark.parametrize("name", list(sc.CASES))
def test_scale_fixture_vs_reference(name):
    solver, steps, kw, N, kind, k = sc.CASES[name]
    g = np.load(_SCALE)
    coeffs, shift = sc.matrix_params(kind, N)
    Ad = device_stencil7(N, N, N, coeffs=coeffs, shift=shift)
    xs = torch.from_numpy(sc.xstar(name)).cuda()
    b = Ad.matvec_device(xs)  # bit-identical to SciPy's csr_matvec(s) (tests/test_gpu_kernels.py)
    del xs
    _, info = getattr(kb, solver)(Ad, b, tol=0.0, atol=0.0, maxiter=steps, **kw)
    assert info.numsteps == steps
    res = np.asarray(info.resnorms, dtype=float)
    ref = g[name + "_resnorms"]
    assert res.shape == ref.shape
    # north-star tolerances: residual history 1e-8 relative, solution 1e-10 relative
    assert np.max(np.abs(res - ref) / ref) <= 1e-8, np.max(np.abs(res - ref) / ref)
    x = info.xk
    xn = torch.sqrt(torch.sum(x * x, dim=0)).cpu().numpy()
    assert np.all(np.abs(xn - g[name + "_xnorm2"]) <= 1e-10 * g[name + "_xnorm2"])
    idx = torch.from_numpy(sc.sample_index(N ** 3)).cuda()
    samp = x[idx].cpu().numpy()
    sref = g[name + "_xsample"]
    assert np.max(np.abs(samp - sref)) <= 1e-10 * np.max(np.abs(sref))
