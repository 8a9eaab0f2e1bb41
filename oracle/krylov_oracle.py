"""CPU oracle for the krylov iteration hot path -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the algorithms of the reference
package (ju-liu/krylov v0.0.3, pure Python) for the path named in
BASELINE.json: ``cg`` / ``minres`` / ``gmres``, the Arnoldi builders
(MGS x N, Lanczos, Householder), ``givens`` and ``Householder``.

It is the *checker* for the CUDA product in ``krylov_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``krylov_b200/`` imports
this file, and the product fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real
reference from /root/reference (NumPy-2 shim), runs it
on seeded inputs and stores its outputs in ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
fixtures and against the reference's own known-answer vectors
(reference ``tests/test_solvers.py:123-144``).

One extension has no reference counterpart and therefore no pin: ``ortho="cgs<N>"``
(classical Gram-Schmidt, ``ArnoldiMGS(classical=True)``) -- parity UNPINNED for it; the tests
tie it to the pinned MGS results (same Krylov space => same residual history up to the loss of
orthogonality, which they bound).

Each function cites the reference file:line it restates (paths relative to
/root/reference/src/krylov/).  The arithmetic order of the reference is kept
(e.g. ``yk += alpha * p`` is a rounded product followed by a rounded add) so
that per-iteration residual norms agree to rounding noise.

The third-party arithmetic the reference delegates to (SURVEY.md section 8c:
SciPy sparsetools ``csr_matvec(s)``, ``np.dot``/``einsum``, LAPACK ``lartg`` /
``trtrs``) is used here through the same libraries, plus a pure-NumPy
statement of LAPACK 3.10 ``dlartg`` (``lartg_f64``) that the device code
mirrors.
"""
from __future__ import annotations

import collections

import numpy as np

__all__ = [
    "Info",
    "ArgumentError",
    "default_inner",
    "lartg_f64",
    "givens",
    "Householder",
    "ArnoldiMGS",
    "ArnoldiLanczos",
    "ArnoldiHouseholder",
    "cg",
    "minres",
    "gmres",
    "gmres_restarted",
]

# reference _helpers.py:93-98 -- same field names, same defaults
Info = collections.namedtuple(
    "IterInfo",
    ["success", "xk", "numsteps", "resnorms", "num_operations", "arnoldi"],
    defaults=(None, None),
)


class ArgumentError(Exception):
    """reference errors.py:1-9"""


# --------------------------------------------------------------------------
# operator / inner-product protocol (reference _helpers.py)
# --------------------------------------------------------------------------
class _Eye:
    """reference _helpers.py:26-36 -- returns its argument unchanged."""

    dtype = np.dtype("u1")

    def __matmul__(self, x):
        return x


def _as_op(M):
    # reference _helpers.py:83-90 (rmatvec is not on the hot path)
    if M is None:
        return _Eye()
    if not hasattr(M, "__matmul__"):
        raise ValueError(f"Unknown linear operator {M}")
    return M


class _Chain:
    """reference _helpers.py:39-48: ops applied right-to-left on a *copy*."""

    def __init__(self, *ops):
        self.ops = ops
        self.dtype = np.result_type(*[np.dtype(o.dtype) for o in ops])

    def __matmul__(self, x):
        out = x.copy()
        for op in reversed(self.ops):
            out = op @ out
        return out


def default_inner(shape):
    """reference _helpers.py:101-110: np.dot for 1-D, column-wise einsum else."""
    if len(shape) == 1:
        return lambda x, y: np.dot(x.conj(), y)
    return lambda x, y: np.einsum("i...,i...->...", x.conj(), y)


def _nz(d):
    """The reference's zero-division guard ``np.where(d != 0, d, 1.0)``
    (cg.py:177,185; minres.py:219; arnoldi.py:147,191,229,274)."""
    return np.where(d != 0, d, 1.0)


def _real_or_raise(v, what):
    # cg.py:91-93, minres.py:102-104, gmres.py:110-112
    if np.any(np.imag(v) != 0.0):
        raise ValueError(f"inner product {what} gave nonzero imaginary part")
    return np.real(v)


def _mm2(G, v):
    """batched (2,2,...)@(2,...) -- minres.py:23-25 / gmres.py:19-21"""
    return np.einsum("ij...,j...->i...", G, v)


# --------------------------------------------------------------------------
# Givens (reference givens.py:5-47 -> LAPACK ?lartg)
# --------------------------------------------------------------------------
_SAFMIN = np.finfo(np.float64).tiny
_SAFMAX = 1.0 / _SAFMIN
_RTMIN = np.sqrt(_SAFMIN)
_RTMAX = np.sqrt(_SAFMAX / 2.0)


def lartg_f64(f, g):
    """LAPACK 3.10 ``dlartg`` (la_lartg.f90) for real doubles: returns
    ``(c, s, r)`` with ``[[c, s], [-s, c]] @ [f, g] = [r, 0]``.

    This is the published algorithm of the routine the reference calls at
    givens.py:35-38; the device function ``kb_dlartg`` mirrors it line by line.
    """
    f = float(f)
    g = float(g)
    f1, g1 = abs(f), abs(g)
    if g == 0.0:
        return 1.0, 0.0, f
    if f == 0.0:
        return 0.0, float(np.copysign(1.0, g)), g1
    if _RTMIN < f1 < _RTMAX and _RTMIN < g1 < _RTMAX:
        d = np.sqrt(f * f + g * g)
        c = f1 / d
        r = float(np.copysign(d, f))
        return c, g / r, r
    u = min(_SAFMAX, max(_SAFMIN, f1, g1))
    fs, gs = f / u, g / u
    d = np.sqrt(fs * fs + gs * gs)
    c = abs(fs) / d
    r = float(np.copysign(d, f))
    s = gs / r
    return c, s, r * u


def givens(X):
    """reference givens.py:5-47.  ``X.shape == (2, ...)``; returns ``G`` of
    shape ``(2, 2, ...)`` (``[[c, s], [-conj(s), c]]`` per trailing index) and
    ``r``.  Real input goes through :func:`lartg_f64`; complex input through
    SciPy's ``zlartg`` exactly like the reference."""
    X = np.asarray(X)
    assert X.shape[0] == 2
    tail = X.shape[1:]
    flat = X.reshape(2, -1)
    ncol = flat.shape[1]
    if np.iscomplexobj(flat):
        from scipy.linalg import lapack

        fn = lapack.get_lapack_funcs("lartg", (flat,))
        triples = [fn(flat[0, j], flat[1, j]) for j in range(ncol)]
    else:
        triples = [lartg_f64(flat[0, j], flat[1, j]) for j in range(ncol)]
    G = np.array([[[c, s], [-np.conj(s), c]] for c, s, _ in triples])
    G = np.moveaxis(G, 0, -1).reshape(2, 2, *tail)
    r = np.array([t[2] for t in triples])
    return G, r


# --------------------------------------------------------------------------
# Householder reflector (reference householder.py:6-81)
# --------------------------------------------------------------------------
class Householder:
    """``H = I - beta v v^H`` with ``H x = alpha ||x|| e_1``
    (reference householder.py:7-51).  Quasi-1-D input only (:18-21)."""

    def __init__(self, x):
        if not (x.ndim == 1 or (x.ndim == 2 and x.shape[1] == 1)):
            raise AssertionError(
                "Householder only works for quasi-1D vectors for now. "
                f"Input vector has shape {x.shape}."
            )
        self.inner = default_inner(x.shape)
        v = x.copy()
        gamma = v[0].copy()
        v[0] = 1
        sigma2 = self.inner(v[1:], v[1:])
        xnorm = np.sqrt(np.abs(gamma) ** 2 + sigma2)
        if sigma2 == 0:  # multiple of e_1 (householder.py:34-37)
            beta = 0
            xnorm = np.abs(gamma)
            alpha = 1 if gamma == 0 else gamma / xnorm
        else:  # householder.py:38-45
            beta = 2
            if gamma == 0:
                v[0] = -np.sqrt(sigma2)
                alpha = 1
            else:
                v[0] = gamma + gamma / np.abs(gamma) * xnorm
                alpha = -gamma / np.abs(gamma)
        self.xnorm = xnorm
        self.v = v / np.sqrt(np.abs(v[0]) ** 2 + sigma2)
        self.alpha = alpha
        self.beta = beta

    def __matmul__(self, x):
        # householder.py:53-62
        if x.shape != self.v.shape:
            raise ValueError(
                f"Shape mismatch! (v.shape = {self.v.shape} != {x.shape} = x.shape)"
            )
        if self.beta == 0:
            return x
        return x - self.beta * self.v * self.inner(self.v, x)

    def matrix(self):
        # householder.py:64-81 (dense; test aid)
        n = self.v.shape[0]
        eye = np.zeros([n, n] + list(self.v.shape[1:]))
        i = np.arange(n)
        eye[i, i] = 1.0
        return eye - self.beta * np.einsum("i...,j...->ij...", self.v, self.v.conj())


# --------------------------------------------------------------------------
# Arnoldi builders (reference arnoldi.py)
# --------------------------------------------------------------------------
class ArnoldiMGS:
    """reference arnoldi.py:107-200.  Keeps the two bases V (= M P) and P;
    ``next()`` returns ``(v_new | None, h)`` with ``h.shape == (k+2, ...)``.
    Unlike the reference ctor (arnoldi.py:143 uses the *argument* ``inner``,
    which may be None), the resolved inner product is used for the initial
    norm -- the solvers always pass ``inner`` so behaviour there is equal."""

    def __init__(self, A, v, num_reorthos=1, M=None, Mv=None, Mv_norm=None, inner=None,
                 classical=False):
        self.inner = default_inner(v.shape) if inner is None else inner
        self.A = A
        self.v = v
        self.num_reorthos = num_reorthos
        # classical=True is NOT in the reference: checker for the product's additive
        # ortho="cgs<N>" extension (all projections of a pass taken from the same w).
        # Parity for it is UNPINNED by construction; tests tie it to MGS instead.
        self.classical = classical
        self.M = _as_op(M)
        self.dtype = np.result_type(np.dtype(A.dtype), np.dtype(self.M.dtype), v.dtype)
        self.iter = 0
        self.is_invariant = False
        p = v
        v = self.M @ p if Mv is None else Mv
        self.vnorm = np.sqrt(self.inner(p, v)) if Mv_norm is None else Mv_norm
        d = _nz(self.vnorm)
        self.P = [p / d]
        self.V = [v / d]

    def __iter__(self):
        return self

    def __next__(self):
        if self.is_invariant:
            raise ArgumentError(
                "Krylov subspace was found to be invariant in the previous iteration."
            )
        k = self.iter
        w = self.A @ self.V[k]  # arnoldi.py:176
        h = np.zeros([k + 2] + list(self.v.shape[1:]), dtype=self.dtype)
        for _ in range(self.num_reorthos):  # arnoldi.py:181-182
            if self.classical:  # extension, see __init__
                a_all = [self.inner(self.V[j], w) for j in range(k + 1)]
                for j in range(k + 1):
                    h[j] += a_all[j]
                    w -= a_all[j] * self.P[j]
                continue
            for j in range(k + 1):  # arnoldi.py:157-162
                a = self.inner(self.V[j], w)
                h[j] += a
                w -= a * self.P[j]
        Mw = self.M @ w
        h[k + 1] = np.sqrt(self.inner(w, Mw))  # arnoldi.py:184-185
        if np.all(h[k + 1] <= 1.0e-14):
            self.is_invariant = True
            vnew = None
        else:
            d = _nz(h[k + 1])
            self.P.append(w / d)
            vnew = Mw / d
            self.V.append(vnew)
        self.h = h
        self.iter += 1
        return vnew, h


class ArnoldiLanczos:
    """reference arnoldi.py:203-281: three-term recurrence;
    ``next()`` returns ``(v, h, p)`` with ``h = [beta_{k-1}, alpha_k, beta_k]``
    (``h`` is the same array object on every call, as in the reference)."""

    def __init__(self, A, v, M=None, Mv=None, Mv_norm=None, inner=None):
        self.A = A
        self.M = _as_op(M)
        self.inner = default_inner(v.shape) if inner is None else inner
        self.dtype = np.result_type(np.dtype(A.dtype), np.dtype(self.M.dtype), v.dtype)
        self.num_iter = 0
        self.h = np.zeros([3] + list(v.shape[1:]), dtype=self.dtype)
        self.is_invariant = False
        p = v
        v = self.M @ p if Mv is None else Mv
        self.vnorm = np.sqrt(self.inner(p, v)) if Mv_norm is None else Mv_norm
        d = _nz(self.vnorm)
        self.p_old = None
        self.p = p / d
        self.v = v / d

    def __iter__(self):
        return self

    def __next__(self):
        if self.is_invariant:
            raise ArgumentError(
                "Krylov subspace was found to be invariant in the previous iteration."
            )
        w = self.A @ self.v  # arnoldi.py:244
        if self.num_iter > 0:  # arnoldi.py:246-249
            self.h[0] = self.h[2]
            w -= self.h[0] * self.p_old
        a = self.inner(self.v, w)  # arnoldi.py:252
        self.h[1] = a
        w -= a * self.p  # arnoldi.py:264
        Mw = self.M @ w
        self.h[2] = np.sqrt(self.inner(w, Mw))  # arnoldi.py:266-267
        if np.all(self.h[2] <= 1.0e-14):
            self.is_invariant = True
            self.v = None
            self.p = None
        else:
            d = _nz(self.h[2])
            self.p_old = self.p
            self.p = w / d
            self.v = Mw / d
        self.num_iter += 1
        return self.v, self.h, self.p


class ArnoldiHouseholder:
    """reference arnoldi.py:33-104 (Walker's Householder Arnoldi; Euclidean
    inner product, no M, quasi-1-D vectors)."""

    def __init__(self, A, v):
        self.inner = default_inner(v.shape)
        self.A = A
        self.v = v
        self.dtype = np.result_type(np.dtype(A.dtype), v.dtype)
        self.iter = 0
        self.is_invariant = False
        self.houses = [Householder(v)]
        self.vnorm = np.linalg.norm(v, 2)
        self.V = [v / _nz(self.vnorm)]

    def __iter__(self):
        return self

    def __next__(self):
        if self.is_invariant:
            raise ArgumentError(
                "Krylov subspace was found to be invariant in the previous iteration."
            )
        k = self.iter
        w = self.A @ self.V[k]
        for j in range(k + 1):  # arnoldi.py:75-77
            w[j:] = self.houses[j] @ w[j:]
            w[j] *= np.conj(self.houses[j].alpha)
        N = self.v.shape[0]
        if k < N - 1:
            hh = Householder(w[k + 1:])  # arnoldi.py:81-83
            self.houses.append(hh)
            w[k + 1:] = (hh @ w[k + 1:]) * np.conj(hh.alpha)
            h = w[: k + 2]
            h[-1] = np.abs(h[-1])
            if h[-1] <= 1.0e-14:
                self.is_invariant = True
                vnew = None
            else:  # arnoldi.py:91-96
                vnew = np.zeros_like(self.v)
                vnew[k + 1] = 1
                for j in range(k + 1, -1, -1):
                    vnew[j:] = self.houses[j] @ vnew[j:]
                vnew = vnew * self.houses[-1].alpha
                self.V.append(vnew)
        else:  # arnoldi.py:97-101
            h = np.zeros([len(w) + 1] + list(self.v.shape[1:]), w.dtype)
            h[:-1] = w
            self.is_invariant = True
            vnew = None
        self.iter += 1
        return vnew, h


# --------------------------------------------------------------------------
# shared set-up of the three solvers
# --------------------------------------------------------------------------
def _check_shapes(A, b):
    # cg.py:99-101 / minres.py:83-85 / gmres.py:116-118
    assert len(A.shape) == 2
    assert A.shape[0] == A.shape[1]
    assert A.shape[1] == b.shape[0]


def _residual_triple(A, b, M, Ml, inner, z):
    """(M Ml r, Ml r, <Ml r, M Ml r>) with r = b - A z.
    cg.py:72-95, gmres.py:105-114, minres.py:121-127."""
    Ml_r = Ml @ (b - A @ z)
    M_Ml_r = M @ Ml_r
    n2 = _real_or_raise(inner(Ml_r, M_Ml_r), "<x, M x>")
    return M_Ml_r, Ml_r, n2


# --------------------------------------------------------------------------
# CG (reference cg.py:16-259)
# --------------------------------------------------------------------------
def cg(A, b, M=None, Ml=None, inner=None, x0=None, tol=1e-5, atol=1.0e-15,
       maxiter=None, return_arnoldi=False, callback=None):
    b = np.asarray(b)
    _check_shapes(A, b)
    N = A.shape[0]
    inner = default_inner(b.shape) if inner is None else inner
    M, Ml = _as_op(M), _as_op(Ml)
    op = _Chain(Ml, A)  # cg.py:109
    maxiter = N if maxiter is None else maxiter
    x0 = np.zeros_like(b) if x0 is None else x0

    z0, r0, rho0 = _residual_triple(A, b, M, Ml, inner, x0)  # cg.py:116
    nrm0 = np.sqrt(rho0)
    if callback is not None:
        callback(x0, r0)

    resnorms = [nrm0]
    yk = np.zeros(x0.shape, dtype=z0.dtype)
    xk = None
    rho_prev, rho = None, rho0
    r = r0.copy()  # Ml_rk
    z = z0.copy()  # M_Ml_rk
    p = z.copy()

    if return_arnoldi:  # cg.py:141-149
        V = [z0 / np.where(nrm0 > 0.0, nrm0, 1.0)]
        P = [r0 / np.where(nrm0 > 0.0, nrm0, 1.0)]
        H = np.zeros([maxiter + 1, maxiter] + list(b.shape[1:]), dtype=float)
        alpha_old = 0

    k = 0
    success = False
    crit = np.maximum(tol * resnorms[0], atol)  # cg.py:154
    while True:
        if np.all(resnorms[-1] <= crit):  # cg.py:156-164: explicit confirmation
            xk = x0 + yk if xk is None else xk
            _, _, n2 = _residual_triple(A, b, M, Ml, inner, xk)
            resnorms[-1] = np.sqrt(n2)
            if np.all(resnorms[-1] <= crit):
                success = True
                break
        if k == maxiter:
            break
        if k > 0:  # cg.py:175-178
            omega = rho / _nz(rho_prev)
            p = z + omega * p
        Ap = op @ p  # cg.py:180
        pAp = inner(p, Ap)
        alpha = rho / _nz(pAp)  # cg.py:185
        yk += alpha * p  # cg.py:196
        xk = None
        r -= alpha * Ap  # cg.py:200
        if callback is not None:
            xk = x0 + yk
            callback(xk, r)
        z = M @ r  # cg.py:207
        rho_new = _real_or_raise(inner(r, z), "<r, M r>")
        rho_prev, rho = rho, rho_new
        nrm = np.sqrt(rho_new)
        resnorms.append(nrm)
        if return_arnoldi:  # cg.py:220-232
            sgn = (-1) ** (k + 1)
            V.append(sgn * z / nrm)
            P.append(sgn * r / nrm)
            H[k, k] = 1.0 / alpha
            if k > 0:
                H[k - 1, k] = H[k, k - 1]
                H[k, k] += omega / alpha_old
            H[k + 1, k] = np.sqrt(rho / rho_prev) / alpha
            alpha_old = alpha
        k += 1

    xk = x0 + yk if xk is None else xk
    if return_arnoldi:
        H = H[: k + 1, :k]
    nops = {"A": 1 + k, "M": 2 + k, "Ml": 2 + k, "Mr": 1 + k,
            "inner": 2 + 2 * k, "axpy": 2 + 2 * k}  # cg.py:243-250
    return (xk if success else None), Info(
        success, xk, k, resnorms, num_operations=nops,
        arnoldi=[V, H, P] if return_arnoldi else None)


# --------------------------------------------------------------------------
# MINRES (reference minres.py:28-253)
# --------------------------------------------------------------------------
def minres(A, b, M=None, Ml=None, Mr=None, inner=None, x0=None, tol=1e-5,
           atol=1.0e-15, maxiter=None, callback=None):
    b = np.asarray(b)
    _check_shapes(A, b)
    M, Ml, Mr = _as_op(M), _as_op(Ml), _as_op(Mr)
    inner = default_inner(b.shape) if inner is None else inner
    N = A.shape[0]
    maxiter = N if maxiter is None else maxiter
    if x0 is None:
        x0 = np.zeros_like(b)

    def explicit_norm(zz):  # minres.py:100-112
        Ml_r = Ml @ (b - A @ zz)
        return np.sqrt(_real_or_raise(inner(Ml_r, M @ Ml_r), "<x, x>"))

    M_Ml_r, Ml_r, n2 = _residual_triple(A, b, M, Ml, inner, x0)  # minres.py:121-127
    nrm0 = np.sqrt(n2)
    dtype = M_Ml_r.dtype
    op = _Chain(Ml, A, Mr)  # minres.py:136
    lan = ArnoldiLanczos(op, Ml_r, M=M, Mv=M_Ml_r, Mv_norm=nrm0, inner=inner)

    W = [np.zeros(b.shape, dtype=dtype), np.zeros(b.shape, dtype=dtype)]
    y = np.array([nrm0, np.zeros_like(nrm0)])
    G = [None, None]
    yk = np.zeros(b.shape, dtype=dtype)
    xk = None
    rn = np.array(nrm0)
    if callback is not None:
        callback(x0, rn)
    resnorms = [rn[()]]

    k = 0
    success = False
    crit = np.maximum(tol * resnorms[0], atol)
    while True:
        if np.all(resnorms[-1] <= crit):  # minres.py:169-175
            xk = x0 + Mr @ yk if xk is None else xk
            resnorms[-1] = explicit_norm(xk)
            if np.all(resnorms[-1] <= crit):
                success = True
                break
        if k == maxiter:
            break
        v = lan.v
        _, h, _ = next(lan)  # minres.py:187-188
        assert np.all(np.abs(np.imag(h))) < 1.0e-14
        h = np.real(h)

        # implicit QR of the tridiagonal (minres.py:195-215)
        R = np.zeros([4] + list(b.shape[1:]), dtype=float)
        R[1] = h[0]
        if G[1] is not None:
            R[:2] = _mm2(G[1], R[:2])
        R[2] = h[1]
        R[3] = h[2]
        if G[0] is not None:
            R[1:3] = _mm2(G[0], R[1:3])
        G[1] = G[0]
        G[0], rr = givens(R[2:4])
        R[2] = rr
        R[3] = 0.0
        y = _mm2(G[0], y)

        # vector update (minres.py:219-221)
        z = (v - R[0] * W[0] - R[1] * W[1]) / _nz(R[2])
        W[0], W[1] = W[1], z
        yk += y[0] * z
        xk = None
        y = np.array([y[1], np.zeros_like(y[1])])
        rn = np.array(np.abs(y[0]))
        if callback is not None:
            xk = x0 + Mr @ yk
            callback(xk, rn)
        resnorms.append(rn[()])
        k += 1

    if xk is None:
        xk = x0 + Mr @ yk
    nops = {"A": 1 + k, "M": 2 + k, "Ml": 2 + k, "Mr": 1 + k,
            "inner": 2 + 2 * k, "axpy": 4 + 8 * k}  # minres.py:242-249
    return (xk if success else None), Info(success, xk, k, resnorms, num_operations=nops)


# --------------------------------------------------------------------------
# GMRES (reference gmres.py:41-251)
# --------------------------------------------------------------------------
def _solve_upper_per_column(Rk, y):
    """gmres.py:24-38: one ``trtrs`` per right-hand-side column; an all-zero
    rhs column short-circuits to zeros."""
    import scipy.linalg

    shp = Rk.shape
    a = Rk.reshape(shp[0], shp[1], -1)
    bb = y.reshape(y.shape[0], -1)
    cols = []
    for j in range(a.shape[2]):
        if np.all(bb[:, j] == 0.0):
            cols.append(np.zeros(bb[:, j].shape))
        else:
            cols.append(scipy.linalg.solve_triangular(a[:, :, j], bb[:, j]))
    return np.array(cols).T.reshape([shp[0]] + list(shp[2:]))


def gmres(A, b, M=None, Ml=None, Mr=None, inner=None, ortho="mgs", x0=None,
          tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    b = np.asarray(b)
    _check_shapes(A, b)
    M, Ml, Mr = _as_op(M), _as_op(Ml), _as_op(Mr)
    inner_given = inner is not None
    inner = default_inner(b.shape) if inner is None else inner
    maxiter = A.shape[0] if maxiter is None else maxiter
    if x0 is None:
        x0 = np.zeros_like(b)
    x0 = np.asarray(x0)

    def explicit_norm(zz):  # gmres.py:101-114
        return np.sqrt(_residual_triple(A, b, M, Ml, inner, zz)[2])

    def solution(yv):  # gmres.py:89-99
        if yv is None:
            return x0
        kk = arn.iter
        if kk > 0:
            yy = _solve_upper_per_column(R[:kk, :kk], yv)
            comb = sum(c * v for c, v in zip(yy, arn.V))
            return x0 + Mr @ comb
        return x0

    z0, r0, n2 = _residual_triple(A, b, M, Ml, inner, x0)
    nrm0 = np.sqrt(n2)
    op = _Chain(Ml, A, Mr)
    resnorms = [nrm0]
    if callback is not None:
        callback(x0, r0)

    if ortho.startswith("mgs"):  # gmres.py:147-157
        nre = 1 if len(ortho) == 3 else int(ortho[3:])
        arn = ArnoldiMGS(op, r0, num_reorthos=nre, M=M, Mv=z0, Mv_norm=nrm0, inner=inner)
    elif ortho.startswith("cgs"):  # extension (not in the reference): classical Gram-Schmidt
        nre = 1 if len(ortho) == 3 else int(ortho[3:])
        arn = ArnoldiMGS(op, r0, num_reorthos=nre, M=M, Mv=z0, Mv_norm=nrm0, inner=inner,
                         classical=True)
    else:  # gmres.py:158-162
        assert ortho == "householder"
        assert not inner_given
        assert isinstance(M, _Eye)
        arn = ArnoldiHouseholder(op, r0)

    G = []
    dtype = z0.dtype
    R = np.zeros([maxiter + 1, maxiter] + list(b.shape[1:]), dtype=dtype)
    y = np.zeros([maxiter + 1] + list(b.shape[1:]), dtype=dtype)
    y[0] = nrm0
    yk = None
    xk = None

    k = 0
    success = False
    crit = np.maximum(tol * resnorms[0], atol)
    while True:
        if np.all(resnorms[-1] <= crit):  # gmres.py:180-187
            xk = solution(yk) if xk is None else xk
            resnorms[-1] = explicit_norm(xk)
            if np.all(resnorms[-1] <= crit):
                success = True
                break
        if k == maxiter:
            break
        _, h = next(arn)  # gmres.py:199
        R[: k + 2, k] = h[: k + 2]
        for i in range(k):  # gmres.py:209-210
            R[i: i + 2, k] = _mm2(G[i], R[i: i + 2, k])
        g, rr = givens(R[k: k + 2, k])  # gmres.py:213-217
        G.append(g)
        R[k, k] = rr
        R[k + 1, k] = 0.0
        y[k: k + 2] = _mm2(G[k], y[k: k + 2])
        yk = y[: k + 1]
        rn = np.array(np.abs(y[k + 1]))
        xk = None
        if callback is not None:
            xk = solution(yk)
            callback(xk, rn)
        resnorms.append(rn[()])
        k += 1

    if xk is None:
        xk = solution(y[: arn.iter])
    nops = {"A": 1 + k, "M": 2 + k, "Ml": 2 + k, "Mr": 1 + k,
            "inner": 2 + k + k * (k + 1) / 2,
            "axpy": 4 + 2 * k + k * (k + 1) / 2}  # gmres.py:240-247
    return (xk if success else None), Info(success, xk, k, resnorms, num_operations=nops)


def gmres_restarted(A, b, restart, max_cycles, x0=None, tol=1e-5, atol=1e-15, **kw):
    """GMRES(m) the way a reference user writes it (there is no
    restart parameter; a cycle is ``gmres(maxiter=m, x0=info.xk)``).

    The stopping target is fixed from the *first* cycle's initial residual
    (``max(tol*||r0||, atol)``); later cycles are run with ``tol=0`` and
    ``atol=target`` so every cycle tests against the same number.  Returns
    ``(x | None, Info)`` with the concatenated residual history (the first
    entry of each later cycle repeats the last one of the previous and is
    dropped)."""
    b = np.asarray(b)
    x = np.zeros_like(b) if x0 is None else np.asarray(x0)
    hist = None
    target = None
    total = 0
    ok = False
    for _ in range(max_cycles):
        if target is None:
            sol, info = gmres(A, b, x0=x, tol=tol, atol=atol, maxiter=restart, **kw)
            target = np.maximum(tol * info.resnorms[0], atol)
            hist = list(info.resnorms)
        else:
            sol, info = gmres(A, b, x0=x, tol=0.0, atol=target, maxiter=restart, **kw)
            hist.extend(info.resnorms[1:])
        total += info.numsteps
        x = info.xk
        if info.success:
            ok = True
            break
    return (x if ok else None), Info(ok, x, total, hist)
