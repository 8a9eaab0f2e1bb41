"""Seeded parity cases shared by tests/golden/make_golden.py (run against the
real reference), tests/test_oracle_golden.py (oracle) and the ``-m gpu``
parity tests (CUDA product).  Pure NumPy/SciPy; no CUDA, no reference import.
"""
import importlib.util
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location(
    "kb_stencils_for_tests", os.path.join(_ROOT, "krylov_b200", "stencils.py"))
st = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(st)


def rhs(A, shape, seed=0):
    """SURVEY.md 8d recipe: x* = default_rng(seed).standard_normal, b = A x*."""
    xs = np.random.default_rng(seed).standard_normal(shape)
    return xs, A @ xs


def kat_matrix(n=100):
    # reference tests/test_solvers.py:134
    return np.diag([1.0e-3] + list(range(2, n + 1))).astype(float)


def _pre_problem(n=60):
    a = np.linspace(1.0, 2.0, n)
    A = np.diag(a)
    A[0, 0] = 1e-2
    A = A + 0.05 * (np.eye(n, k=1) + np.eye(n, k=-1))
    b = np.ones(n)
    Md = np.diag(1.0 / np.diag(A))
    return a, A, b, Md


def solver_cases():
    """name -> (solver, A, b, kwargs).  ``solver`` in {cg, minres, gmres}."""
    cases = {}
    A = kat_matrix()
    for shape in [(100,), (100, 1)]:
        b = np.ones(shape)
        tag = "kat%d" % len(shape)
        cases[tag + "_cg"] = ("cg", A, b, {})
        cases[tag + "_minres"] = ("minres", A, b, {})
        for o in ["mgs", "mgs2", "householder"]:
            cases[tag + "_gmres_" + o] = ("gmres", A, b, {"ortho": o})
    # C1 shrunk: 2-D Poisson 32^2, tol 1e-10
    A = st.poisson2d(32)
    _, b = rhs(A, (A.shape[0],))
    cases["p2d32_cg"] = ("cg", A, b, dict(tol=1e-10, maxiter=5000))
    # C2 shrunk: shifted 3-D Laplacian 12^3, mild shift
    A = st.shifted_laplace3d(12)
    _, b = rhs(A, (A.shape[0],))
    cases["sl12_minres"] = ("minres", A, b, dict(tol=1e-8, maxiter=2000))
    # C3 shrunk: convection-diffusion 10^3, one 30-step cycle per ortho
    A = st.convection_diffusion3d(10)
    _, b = rhs(A, (A.shape[0],))
    for o in ["mgs", "mgs2", "householder"]:
        cases["cd10_gmres_" + o] = ("gmres", A, b, dict(tol=1e-8, maxiter=30, ortho=o))
    # C4 shrunk: blocked k=4 on 3-D Poisson 10^3
    A = st.poisson3d(10)
    _, B = rhs(A, (A.shape[0], 4))
    cases["p3d10_cg_k4"] = ("cg", A, B, dict(tol=1e-8, maxiter=2000))
    cases["p3d10_minres_k4"] = ("minres", A, B, dict(tol=1e-8, maxiter=2000))
    cases["p3d10_gmres_k4"] = ("gmres", A, B, dict(tol=1e-8, maxiter=60))
    cases["p3d10_gmres_mgs2_k4"] = ("gmres", A, B, dict(tol=1e-8, maxiter=60, ortho="mgs2"))
    # ragged block: one all-zero column (reference tests/linear_problems.py:33-54)
    B2 = B.copy()
    B2[:, 1] = 0.0
    cases["p3d10_cg_k4_zero_col"] = ("cg", A, B2, dict(tol=1e-8, maxiter=2000))
    cases["p3d10_minres_k4_zero_col"] = ("minres", A, B2, dict(tol=1e-8, maxiter=2000))
    cases["p3d10_gmres_k4_zero_col"] = ("gmres", A, B2, dict(tol=1e-8, maxiter=60))
    # odd block width k=3 and (n,1)
    _, B3 = rhs(A, (A.shape[0], 3), seed=3)
    cases["p3d10_cg_k3"] = ("cg", A, B3, dict(tol=1e-9, maxiter=2000))
    _, B1 = rhs(A, (A.shape[0], 1), seed=4)
    cases["p3d10_cg_k1col"] = ("cg", A, B1, dict(tol=1e-9, maxiter=2000))
    # preconditioners / custom inner / x0  (reference tests/test_solvers.py:90-196)
    a, A, b, Md = _pre_problem()
    n = len(b)
    w = 10.0 / np.arange(1, n + 1)
    # Ml A must stay self-adjoint for cg/minres: diagonal A like the reference's test_ml
    Adiag = np.diag(a)
    Adiag[0, 0] = 1e-2
    Mld = np.diag(1.0 / np.sqrt(a))
    for name in ["cg", "minres", "gmres"]:
        cases["pre_M_" + name] = (name, A, b, dict(M=Md, tol=1e-10))
        cases["pre_Ml_" + name] = (name, Adiag, b, dict(Ml=Mld, tol=1e-10))
        if name != "cg":
            cases["pre_Mr_" + name] = (name, A, b, dict(Mr=Md, tol=1e-10))
        cases["inner_" + name] = (
            name, np.diag(a), b, dict(inner=lambda x, y, w=w: np.dot(x.T, w * y), tol=1e-10))
        cases["x0_" + name] = (name, A, b, dict(x0=np.linspace(-1.0, 1.0, n), tol=1e-10))
    # maxiter exhaustion -> success False, sol None
    cases["maxit_cg"] = ("cg", A, b, dict(tol=1e-14, maxiter=3))
    cases["maxit_minres"] = ("minres", A, b, dict(tol=1e-14, maxiter=3))
    cases["maxit_gmres"] = ("gmres", A, b, dict(tol=1e-14, maxiter=3))
    # zero right-hand side: zero steps (tests/linear_problems.py:24-30)
    cases["zero_rhs_cg"] = ("cg", A, np.zeros(n), dict(tol=1e-7))
    cases["zero_rhs_minres"] = ("minres", A, np.zeros(n), dict(tol=1e-7))
    cases["zero_rhs_gmres"] = ("gmres", A, np.zeros(n), dict(tol=1e-7))
    return cases


def arnoldi_inputs():
    A = st.convection_diffusion3d(6)
    As = st.shifted_laplace3d(6)
    v = np.random.default_rng(1).standard_normal(A.shape[0])
    return A, As, v


GIVENS_F = [0.0, 1.0, -1.0, 1e8, 1e-8, 3.0, -2.5, 1e-200, 1e200, 1e-310]


def householder_inputs():
    rng = np.random.default_rng(2)
    return [rng.standard_normal(10), np.array([0.0] + [1.0] * 9),
            np.array([2.0] + [0.0] * 9), np.zeros(10), np.full(10, 1e-8)]
