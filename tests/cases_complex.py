"""Complex128 (Hermitian) problems for cg / minres: the reference's own test matrices
(tests/linear_problems.py:54-86 `hpd`, `hermitian_indefinite`) and seeded sparse ones.
Shared by tests/golden/make_golden_complex.py (runs the real reference) and the tests."""
import numpy as np
import scipy.sparse as sp


def ref_hpd():
    a = np.array(np.linspace(1.0, 2.0, 5), dtype=complex)
    a[0] = 5.0
    a[-1] = 1.0e-1
    A = np.diag(a)
    A[-1, 0] = 1.0e-1j
    A[0, -1] = -1.0e-1j
    return A, np.ones(5, dtype=complex)


def ref_hermitian_indefinite():
    a = np.array(np.linspace(1.0, 2.0, 5), dtype=complex)
    a[-1] = 1e-3
    A = np.diag(a)
    A[-1, 0] = 10.0j
    A[0, -1] = -10.0j
    return A, np.ones(5, dtype=complex)


def sparse_hermitian(n, shift, seed):
    """Hermitian sparse matrix: random complex off-diagonal band structure + real diagonal;
    shift > 0 large enough makes it positive definite, a small one leaves it indefinite."""
    rng = np.random.default_rng(seed)
    B = sp.random(n, n, density=6.0 / n, random_state=seed, format="csr")
    B.data = rng.standard_normal(B.nnz) + 1j * rng.standard_normal(B.nnz)
    H = (B + B.conj().T) * 0.5
    d = np.asarray(abs(H).sum(axis=1)).ravel()
    return (H + sp.diags(shift * d + 0.1)).tocsr()


def cases():
    out = {}
    A, b = ref_hpd()
    out["hpd_cg"] = ("cg", A, b, {"tol": 1e-12})
    out["hpd_minres"] = ("minres", A, b, {"tol": 1e-12})
    A, b = ref_hermitian_indefinite()
    out["hind_minres"] = ("minres", A, b, {"tol": 1e-12, "maxiter": 20})
    rng = np.random.default_rng(7)
    A = sparse_hermitian(300, 1.05, 1)
    b = rng.standard_normal(300) + 1j * rng.standard_normal(300)
    out["sp_hpd_cg"] = ("cg", A, b, {"tol": 1e-10})
    out["sp_hpd_minres"] = ("minres", A, b, {"tol": 1e-10})
    # (a preconditioned complex case cannot be pinned: the reference raises "inner product
    #  <x, M x> gave nonzero imaginary part" on rounding-level imaginary parts, cg.py:88-92)
    x0 = rng.standard_normal(300) + 1j * rng.standard_normal(300)
    out["sp_hpd_cg_x0"] = ("cg", A, b, {"tol": 1e-10, "x0": x0})
    B3 = rng.standard_normal((300, 3)) + 1j * rng.standard_normal((300, 3))
    out["sp_hpd_cg_k3"] = ("cg", A, B3, {"tol": 1e-10})
    out["sp_hpd_minres_k3"] = ("minres", A, B3, {"tol": 1e-10})
    Ai = sparse_hermitian(300, 0.3, 2)  # indefinite
    out["sp_hind_minres"] = ("minres", Ai, b, {"tol": 1e-9, "maxiter": 600})
    return out
