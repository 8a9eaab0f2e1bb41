"""Seeded cases for the "next" solvers: shared by tests/golden/make_golden_extra.py
(real reference), tests/test_oracle_extra_golden.py (oracle) and the -m gpu tests (CUDA product).
Pure NumPy/SciPy."""
import numpy as np

from cases import rhs, st


def _jacobi(A):
    import scipy.sparse

    d = np.asarray(A.diagonal()) if hasattr(A, "diagonal") else np.diag(A)
    return scipy.sparse.diags([1.0 / d], [0], format="csr")


def extra_cases():
    """name -> (solver, A, b, kwargs)"""
    c = {}
    # non-symmetric sparse: convection-diffusion 8^3 (SURVEY.md 8d, C3 shrunk)
    A = st.convection_diffusion3d(8)
    n = A.shape[0]
    _, b = rhs(A, (n,))
    _, B = rhs(A, (n, 3), seed=5)
    x0 = np.linspace(-1.0, 1.0, n)
    J = _jacobi(A)
    for name in ("bicgstab", "cgs", "bicg", "qmr"):
        c[f"cd8_{name}"] = (name, A, b, dict(tol=1e-9, maxiter=400))
        c[f"cd8_{name}_x0"] = (name, A, b, dict(tol=1e-9, maxiter=400, x0=x0))
        c[f"cd8_{name}_k3"] = (name, A, B, dict(tol=1e-8, maxiter=400))
        c[f"cd8_{name}_maxit"] = (name, A, b, dict(tol=1e-14, maxiter=4))
    c["cd8_bicgstab_Ml"] = ("bicgstab", A, b, dict(tol=1e-9, maxiter=400, Ml=J))
    c["cd8_bicgstab_Mr"] = ("bicgstab", A, b, dict(tol=1e-9, maxiter=400, Mr=J))
    c["cd8_cgs_M"] = ("cgs", A, b, dict(tol=1e-9, maxiter=400, M=J))
    c["cd8_bicg_M"] = ("bicg", A, b, dict(tol=1e-9, maxiter=400, M=J))
    c["cd8_qmr_Ml"] = ("qmr", A, b, dict(tol=1e-9, maxiter=400, Ml=J))
    c["cd8_qmr_Mr"] = ("qmr", A, b, dict(tol=1e-9, maxiter=400, Mr=J))
    c["cd8_gcr"] = ("gcr", A, b, dict(tol=1e-9, maxiter=120))
    c["cd8_gcr_k3"] = ("gcr", A, B, dict(tol=1e-8, maxiter=120))
    c["cd8_gcr_x0"] = ("gcr", A, b, dict(tol=1e-9, maxiter=120, x0=x0))
    # normal equations: small and not too ill-conditioned
    A2 = st.convection_diffusion3d(5)
    n2 = A2.shape[0]
    _, b2 = rhs(A2, (n2,), seed=2)
    c["cd5_cgne"] = ("cgne", A2, b2, dict(tol=1e-9, maxiter=600))
    c["cd5_cgnr"] = ("cgnr", A2, b2, dict(tol=1e-9, maxiter=600))
    _, B2 = rhs(A2, (n2, 2), seed=6)
    c["cd5_cgnr_k2"] = ("cgnr", A2, B2, dict(tol=1e-8, maxiter=600))
    # symmetric positive definite: conjugate residuals, Chebyshev
    P = st.poisson3d(8)
    npn = P.shape[0]
    _, bp = rhs(P, (npn,), seed=7)
    _, Bp = rhs(P, (npn, 2), seed=8)
    c["p8_cgr"] = ("cgr", P, bp, dict(tol=1e-9, maxiter=300))
    c["p8_cgr_M"] = ("cgr", P, bp, dict(tol=1e-9, maxiter=300, M=_jacobi(P)))
    c["p8_cgr_k2"] = ("cgr", P, Bp, dict(tol=1e-8, maxiter=300))
    c["p8_cgr_x0"] = ("cgr", P, bp, dict(tol=1e-9, maxiter=300, x0=np.linspace(0.0, 1.0, npn)))
    lam = lambda i: 4.0 * np.sin(i * np.pi / (2.0 * 9)) ** 2
    est = (3.0 * lam(1), 3.0 * lam(8))
    c["p8_chebyshev"] = ("chebyshev", P, bp, dict(eigenvalue_estimates=est, tol=1e-8, maxiter=400))
    c["p8_chebyshev_k2"] = ("chebyshev", P, Bp, dict(eigenvalue_estimates=est, tol=1e-7, maxiter=400))
    c["p8_chebyshev_M"] = ("chebyshev", P, bp, dict(
        eigenvalue_estimates=(est[0] / 6.0, est[1] / 6.0), M=_jacobi(P), tol=1e-8, maxiter=400))
    # symmetric problems through the non-symmetric solvers, dense operator
    D = np.diag(np.linspace(1.0, 3.0, 40)) + 0.1 * np.triu(np.ones((40, 40)), 1) / 40.0
    bd = np.ones(40)
    for name in ("bicgstab", "cgs", "bicg", "qmr", "gcr"):
        c[f"dense_{name}"] = (name, D, bd, dict(tol=1e-10, maxiter=200))
    # symmlq: the reference records the norm of the Lanczos vector, so it only "converges" when
    # the Krylov space is exhausted -- its own tests use 5 x 5 problems (tests/test_symmlq.py)
    a5 = np.linspace(1.0, 2.0, 5)
    a5i = a5.copy(); a5i[-1] = -1.0   # linear_problems.py:66-72 symmetric_indefinite
    a5p = a5.copy(); a5p[-1] = 1e-2   # linear_problems.py:5-10 spd_dense
    c["sym5_symmlq_indef"] = ("symmlq", np.diag(a5i), np.ones(5), dict(maxiter=10))
    c["sym5_symmlq_spd"] = ("symmlq", np.diag(a5p), np.ones(5), dict(tol=1e-7, maxiter=10))
    c["sym5_symmlq_k3"] = ("symmlq", np.diag(a5p), np.ones((5, 3)), dict(tol=1e-7, maxiter=10))
    S = st.shifted_laplace3d(6)
    _, bs = rhs(S, (S.shape[0],), seed=9)
    c["sl6_symmlq_25"] = ("symmlq", S, bs, dict(tol=1e-9, maxiter=25))
    c["p8_symmlq_M_30"] = ("symmlq", P, bp, dict(tol=1e-9, maxiter=30, M=_jacobi(P)))
    c["p8_symmlq_x0_30"] = ("symmlq", P, bp, dict(tol=1e-9, maxiter=30, x0=np.linspace(0.0, 1.0, npn)))
    # zero right-hand side: zero steps
    for name in ("bicgstab", "cgs", "bicg", "qmr", "cgr", "gcr"):
        c[f"zero_{name}"] = (name, P, np.zeros(npn), dict(tol=1e-7))
    return c
