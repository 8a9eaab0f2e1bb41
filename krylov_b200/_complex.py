"""complex128 for ``cg`` and ``minres`` on HERMITIAN matrices (the reference's `hpd` and
`hermitian_indefinite` problems, tests/linear_problems.py:54-86).

The library's kernels are fp64 (north_star).  A Hermitian system ``A x = b`` is solved through its
real-equivalent embedding in the layout a complex array already has in memory (re, im interleaved):

    x  ->  (x_r[0], x_i[0], x_r[1], x_i[1], ...)         length 2 n
    A  ->  K = A.real (x) I_2 + A.imag (x) J,  J = [[0, -1], [1, 0]]          2 n x 2 n, real

``K`` acts on the interleaved vector exactly as ``A`` acts on ``x``, it is symmetric iff ``A`` is
Hermitian, and the Euclidean inner product of two interleaved vectors is ``Re(x^H y)``.  CG and
MINRES only ever form inner products that are real for Hermitian ``A`` (``<r, r>``, ``<p, A p>``,
the Lanczos coefficients: cg.py:183-209, arnoldi.py:244-267), so the real solver on ``K`` produces
the reference's iterates and residual norms up to rounding.  This does NOT hold for ``gmres`` (its
Arnoldi coefficients are genuinely complex: the real Krylov space of K is a different space) --
complex ``gmres`` stays unsupported.  Cost: K streams 4 real entries + 2 indices per complex
nonzero (2.4 x the bytes of a native complex kernel); every schedule, preconditioner path and the
multi-right-hand-side layout of the real solvers apply unchanged.
"""
from __future__ import annotations

import numpy as np
import torch


def _is_complex(v):
    if v is None:
        return False
    dt = getattr(v, "dtype", None)
    if dt is None:
        return False
    if isinstance(dt, torch.dtype):
        return dt.is_complex
    try:
        return np.issubdtype(np.dtype(dt), np.complexfloating)
    except TypeError:
        return False


def any_complex(*objs):
    return any(_is_complex(o) for o in objs)


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def embed_matrix(A, what, need_hermitian):
    """scipy CSR of the real-equivalent matrix (None stays None)."""
    import scipy.sparse as sp

    if A is None:
        return None
    if sp.issparse(A):
        Ac = A.tocsr()
    elif isinstance(A, (np.ndarray, torch.Tensor)) and getattr(A, "ndim", 0) == 2:
        Ac = sp.csr_matrix(_to_numpy(A))
    else:
        raise NotImplementedError(
            f"complex128: {what} must be given as a matrix (sparse or dense); duck-typed complex "
            "operators are not supported")
    Ac = Ac.astype(np.complex128)
    if need_hermitian:
        D = (Ac - Ac.conj().T).tocsr()
        scale = abs(Ac).max() if Ac.nnz else 0.0
        if D.nnz and abs(D).max() > 1e-13 * max(scale, 1e-300):
            raise NotImplementedError(
                "complex128: cg / minres support Hermitian matrices only (real-equivalent "
                "embedding); this matrix is not Hermitian")
    J = np.array([[0.0, -1.0], [1.0, 0.0]])
    K = sp.kron(Ac.real, np.eye(2), format="csr") + sp.kron(Ac.imag, J, format="csr")
    K = K.tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def embed_vector(v):
    """complex (n,) / (n, k) -> real float64 (2n,) / (2n, k), rows interleaved re / im."""
    a = _to_numpy(v).astype(np.complex128)
    out = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=np.float64)
    out[0::2] = a.real
    out[1::2] = a.imag
    return out


def extract_vector(r):
    r = _to_numpy(r)
    return r[0::2] + 1j * r[1::2]


def solve_hermitian(solver, A, b, x0, mats, inner, callback, kwargs):
    """solver: the real ``cg`` / ``minres``; mats: dict of preconditioner matrices (M, Ml, Mr)."""
    from .operators import Info

    if inner is not None:
        raise NotImplementedError("complex128: a user inner product is not supported")
    if kwargs.get("return_arnoldi"):
        raise NotImplementedError("complex128: return_arnoldi is not supported")
    b_is_torch = isinstance(b, torch.Tensor)
    K = embed_matrix(A, "A", need_hermitian=True)
    pre = {name: embed_matrix(M, name, need_hermitian=False) for name, M in mats.items()}
    cb = None
    if callback is not None:
        second_is_vector = solver.__name__ == "cg"  # cg: (xk, rk); minres: (xk, resnorm array)

        def cb(xk, rk):  # the user sees complex arrays, like the reference's callback
            callback(extract_vector(xk), extract_vector(rk) if second_is_vector else rk)
    sol, info = solver(K, embed_vector(b), x0=None if x0 is None else embed_vector(x0),
                       callback=cb, **pre, **kwargs)
    xk = extract_vector(info.xk)
    if b_is_torch:
        xk = torch.from_numpy(xk).to(b.device)
    out = Info(info.success, xk, info.numsteps, info.resnorms, num_operations=info.num_operations)
    return (xk if sol is not None else None), out
