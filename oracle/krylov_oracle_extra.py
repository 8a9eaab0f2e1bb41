"""CPU oracle for the "next" solvers of SURVEY.md 8(f).2 -- TEST INFRASTRUCTURE ONLY.

NumPy/SciPy restatement of the reference's short-recurrence solvers that share the hot path's
primitive set (sparse product, transposed product, inner product, axpy):

    bicgstab (bicgstab.py:24-144)   cgs  (cgs.py:24-117)    bicg (bicg.py:25-116)
    qmr      (qmr.py:22-160)        cgne (cgne.py:18-45)    cgnr (cgnr.py:15-21)
    cgr      (cgr.py:14-100)        gcr  (gcr.py:16-97)     chebyshev (chebyshev.py:13-99)
    symmlq   (symmlq.py:15-161)

Same rules as oracle/krylov_oracle.py: only tests/ (and bench legs) import it, nothing under
krylov_b200/ does.  Parity status: PINNED -- tests/golden/make_golden_extra.py runs the real
reference on the seeded cases of tests/cases_extra.py and stores its outputs in
tests/golden/extra.npz; tests/test_oracle_extra_golden.py checks every function below against
them.

The reference repeats the same driver in every solver (initial residual, stopping rule with the
explicit-residual confirmation, callback, Info); here it is written once (``_Loop``) and each solver
supplies its state and one ``step``.  The arithmetic statements keep the reference's order of
rounding (``x += a * p`` is a rounded product, then a rounded sum).  Quirks that are kept on
purpose, because callers see them: bicgstab evaluates ``Ml (b - A x)`` of the *old* x in every
step and leaves through that test without updating x (bicgstab.py:117-122); ``maxiter=None`` never
stops a non-converging iteration; cgr/gcr measure the plain inner product even with M.
"""
from __future__ import annotations

import numpy as np

from .krylov_oracle import Info, _Eye, _nz, default_inner

__all__ = ["bicgstab", "cgs", "bicg", "qmr", "cgne", "cgnr", "cgr", "gcr", "chebyshev", "symmlq",
           "adjoint"]


class adjoint:
    """Operator with ``rmatvec`` -- reference _helpers.py:51-90 (LinearOperatorWrapper /
    aslinearoperator): dense arrays use (A.T @ x.conj()).conj(), anything else caches A.T.conj()."""

    def __init__(self, A):
        self.A = A
        self.shape = A.shape
        self.dtype = A.dtype
        self._AH = None

    def __matmul__(self, x):
        return self.A @ x

    def rmatvec(self, x):
        if isinstance(self.A, np.ndarray):
            return (self.A.T @ x.conj()).conj()
        if self._AH is None:
            self._AH = self.A.T.conj()
        return self._AH @ x


class _EyeR(_Eye):
    def rmatvec(self, x):  # _helpers.py:26-36
        return x


def _op(M):
    if M is None:
        return _EyeR()
    if not hasattr(M, "__matmul__"):
        raise ValueError(f"Unknown linear operator {M}")
    return M if hasattr(M, "rmatvec") else adjoint(M)


def _shapes(A, b):
    assert len(A.shape) == 2
    assert A.shape[0] == A.shape[1]
    assert A.shape[1] == b.shape[0]


def _make_norm(inner, W):
    """sqrt(<x, W x>) with the reference's complaint about a complex value."""

    def norm(x):
        v = inner(x, W @ x)
        if np.any(np.imag(v) != 0.0):
            raise ValueError("inner product <x, x> gave nonzero imaginary part")
        return np.sqrt(np.real(v))

    return norm


class _Loop:
    """The driver every reference solver spells out: resnorms list, criterion
    max(tol * resnorms[0], atol), "oh really?" confirmation with the explicit residual,
    maxiter, callback after each step (e.g. cgs.py:75-117)."""

    def __init__(self, A, b, norm, tol, atol, maxiter, callback):
        self.A, self.b, self.norm = A, b, norm
        self.tol, self.atol, self.maxiter, self.callback = tol, atol, maxiter, callback

    def run(self, state, step, first_norm, cb_args, xout=None):
        """state.x is the iterate; step(k) advances it and returns the new residual norm, or the
        tuple ("leave", resnorm) to finish successfully without appending (bicgstab.py:119-122).
        xout(): the point that is checked and returned when it is not state.x itself
        (symmlq.py:84-87,95-103: the CG point)."""
        if xout is not None:
            return self._run_xout(state, step, first_norm, cb_args, xout)
        if self.callback is not None:
            self.callback(*cb_args())
        res = [first_norm]
        crit = np.maximum(self.tol * res[0], self.atol)
        k, ok = 0, False
        while True:
            if np.all(res[-1] <= crit):
                res[-1] = self.norm(self.b - self.A @ state.x)
                if np.all(res[-1] <= crit):
                    ok = True
                    break
            if k == self.maxiter:
                break
            out = step(k, crit)
            if isinstance(out, tuple):
                res[-1] = out[1]
                ok = True
                break
            if self.callback is not None:
                self.callback(*cb_args())
            res.append(out)
            k += 1
        return (state.x if ok else None), Info(ok, state.x, k, res)


    def _run_xout(self, state, step, first_norm, cb_args, xout):
        if self.callback is not None:
            self.callback(*cb_args())
        res = [first_norm]
        crit = np.maximum(self.tol * res[0], self.atol)
        k, ok, xo = 0, False, None
        while True:
            if np.all(res[-1] <= crit):
                xo = xout()
                res[-1] = self.norm(self.b - self.A @ xo)
                if np.all(res[-1] <= crit):
                    ok = True
                    break
            if k == self.maxiter:
                xo = xout()
                break
            res.append(step(k, crit))
            k += 1
        return (xo if ok else None), Info(ok, xo, k, res)


class _S:
    pass


def _start(A, b, x0, copy_x0=True):
    """x, r0 as the reference builds them: zeros / b.copy() without x0, else b - A x0."""
    if x0 is None:
        return np.zeros_like(b), b.copy()
    x = np.array(x0) if copy_x0 else np.asarray(x0)
    return x, b - A @ x


# ------------------------------------------------------------------ BiCGStab --
def bicgstab(A, b, Ml=None, Mr=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15,
             maxiter=None, callback=None):
    """bicgstab.py:24-144 (van der Vorst; netlib templates)."""
    b = np.asarray(b)
    _shapes(A, b)
    A, Ml, Mr = _op(A), _op(Ml), _op(Mr)
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, Ml)
    s = _S()
    s.x, r0 = _start(A, b, x0, copy_x0=False)  # bicgstab.py:57-62: np.asarray, no copy
    shadow = r0  # "common but arbitrary choice" :65
    s.r = r0.copy()
    s.rho = s.alpha = s.omega = 1.0
    s.p = np.zeros_like(b)
    s.v = np.zeros_like(b)

    def step(k, crit):
        rho_old, s.rho = s.rho, inner(shadow, s.r)
        beta = s.rho * s.alpha / _nz(rho_old * s.omega)
        s.p = s.r + beta * (s.p - s.omega * s.v)
        y = Mr @ (Ml @ s.p)
        s.v = A @ y
        s.alpha = s.rho / _nz(inner(shadow, s.v))
        half_r = s.r - s.alpha * s.v
        half_x = s.x + s.alpha * y
        # :117-122 -- the residual of the OLD x, measured through Ml twice
        rn = norm(Ml @ (b - A @ s.x))
        if np.all(rn <= crit):
            return ("leave", rn)
        Ml_s = Ml @ half_r
        z = Mr @ Ml_s
        t = A @ z
        Ml_t = Ml @ t
        s.omega = inner(Ml_t, Ml_s) / _nz(inner(Ml_t, Ml_t))
        s.x = half_x + s.omega * z
        s.r = half_r - s.omega * t
        return norm(s.r)

    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(
        s, step, norm(r0), lambda: (s.x, s.r))


# ----------------------------------------------------------------------- CGS --
def cgs(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """cgs.py:24-117 (Sonneveld)."""
    b = np.asarray(b)
    _shapes(A, b)
    A, M = _op(A), _op(M)
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, M)
    s = _S()
    s.x, r0 = _start(A, b, x0)
    shadow = r0
    s.r = r0.copy()
    s.rho = 1.0
    s.p = np.zeros_like(b)
    s.q = np.zeros_like(b)

    def step(k, crit):
        rho_old, s.rho = s.rho, inner(shadow, s.r)
        beta = s.rho / _nz(rho_old)
        u = s.r + beta * s.q
        s.p = u + beta * (s.q + beta * s.p)
        v = A @ (M @ s.p)
        alpha = s.rho / _nz(inner(shadow, v))
        s.q = u - alpha * v
        uq = M @ (u + s.q)
        s.x += alpha * uq
        s.r -= alpha * (A @ uq)
        return norm(s.r)

    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(
        s, step, norm(s.r), lambda: (s.x, s.r))


# ---------------------------------------------------------------------- BiCG --
def bicg(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """bicg.py:25-116: two coupled residual / direction sequences (A and A^H)."""
    b = np.asarray(b)
    _shapes(A, b)
    A, M = _op(A), _op(M)
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, M)
    s = _S()
    s.x, r0 = _start(A, b, x0)
    s.r = np.array([r0, r0.conj()])  # :60-65
    s.p = [(M @ s.r[0]).copy(), M.rmatvec(s.r[1]).copy()]
    s.rMr = inner(s.r[1], M @ s.r[0])

    def step(k, crit):
        Ap = A @ s.p[0]
        AHp = A.rmatvec(s.p[1])
        alpha = s.rMr / _nz(inner(s.p[1], Ap))
        s.x += alpha * s.p[0]
        s.r[0] -= alpha * Ap
        s.r[1] -= np.conj(alpha) * AHp
        old, s.rMr = s.rMr, inner(s.r[1], M @ s.r[0])
        beta = s.rMr / _nz(old)
        rn = norm(s.r[0])
        s.p[0] = M @ s.r[0] + beta * s.p[0]
        s.p[1] = M.rmatvec(s.r[1]) + np.conj(beta) * s.p[1]
        return rn

    # the reference calls the callback before the direction update (:101-107); the directions
    # are not visible to it, so calling it after the step is indistinguishable
    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(
        s, step, norm(s.r[0]), lambda: (s.x, s.r))


# ----------------------------------------------------------------------- QMR --
def qmr(A, b, Ml=None, Mr=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None,
        callback=None):
    """qmr.py:22-160 (Freund/Nachtigal, coupled two-term look-ahead-free variant of the netlib
    templates)."""
    b = np.asarray(b)
    _shapes(A, b)
    A, Ml, Mr = _op(A), _op(Ml), _op(Mr)
    s = _S()
    if x0 is None:
        s.x, s.r = np.zeros_like(b), b.copy()
    else:
        s.x = np.array(x0)
        s.r = b - A @ x0
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, Ml)
    first = norm(s.r)
    s.vt = s.r.copy()
    s.y = Ml @ s.vt
    s.rho = norm(s.y)
    s.wt = s.r.copy()
    s.z = Mr.rmatvec(s.wt)
    s.xi = norm(s.z)
    s.gamma, s.eta, s.theta, s.eps = 1.0, -1.0, 1.0, 1.0
    s.p = s.q = s.d = s.s = None

    def step(k, crit):
        v = s.vt / _nz(s.rho)
        s.y = s.y / _nz(s.rho)
        w = s.wt / _nz(s.xi)
        s.z = s.z / _nz(s.xi)
        delta = inner(s.z, s.y)
        yt = Mr @ s.y
        zt = Ml.rmatvec(s.z)
        if k == 0:
            s.p, s.q = yt.copy(), zt.copy()
        else:
            de = delta / _nz(s.eps)
            s.p = yt - (s.xi * de) * s.p
            s.q = zt - (s.rho * de) * s.q
        Ap = A @ s.p
        s.eps = inner(s.q, Ap)
        beta = s.eps / _nz(delta)
        s.vt = Ap - beta * v
        s.y = Ml @ s.vt
        rho_old, s.rho = s.rho, norm(s.y)
        s.wt = A.rmatvec(s.q) - beta * w
        s.z = Mr.rmatvec(s.wt)
        s.xi = norm(s.z)
        gamma_old, theta_old = s.gamma, s.theta
        s.theta = s.rho / _nz(gamma_old * np.abs(beta))
        s.gamma = 1 / np.sqrt(1 + s.theta ** 2)
        s.eta = -s.eta * rho_old * s.gamma ** 2 / _nz(beta * gamma_old ** 2)
        if k == 0:
            s.d = s.eta * s.p
            s.s = s.eta * Ap
        else:
            c2 = (theta_old * s.gamma) ** 2
            s.d = s.eta * s.p + c2 * s.d
            s.s = s.eta * Ap + c2 * s.s
        s.x += s.d
        s.r -= s.s
        return norm(s.r)

    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(s, step, first, lambda: (s.x, s.r))


# ------------------------------------------------------- CG on normal equations --
class _Normal:
    """A A^H (cgne.py:7-15) or A^H A (cgnr.py:5-12) as an operator."""

    def __init__(self, A, outer):
        self.A, self.outer = A, outer
        self.shape, self.dtype = A.shape, A.dtype

    def __matmul__(self, x):
        if self.outer:
            return self.A @ self.A.rmatvec(x)
        return self.A.rmatvec(self.A @ x)


def cgne(A, b, *args, **kwargs):
    """cgne.py:18-45: A A^H y = b, x = A^H y."""
    from .krylov_oracle import cg

    A = _op(A)
    sol, info = cg(_Normal(A, True), b, *args, **kwargs)
    xk = A.rmatvec(info.xk)
    return (xk if sol is not None else None), Info(
        info.success, xk, info.numsteps, info.resnorms, info.num_operations, info.arnoldi)


def cgnr(A, b, *args, **kwargs):
    """cgnr.py:15-21: A^H A x = A^H b."""
    from .krylov_oracle import cg

    A = _op(A)
    return cg(_Normal(A, False), A.rmatvec(b), *args, **kwargs)


# ----------------------------------------------------------------------- CGR --
def cgr(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """cgr.py:14-100: conjugate residuals; the preconditioned residual is iterated and measured
    in the plain inner product."""
    b = np.asarray(b)
    _shapes(A, b)
    A, M = _op(A), _op(M)
    s = _S()
    if x0 is None:
        s.x, r = np.zeros_like(b), b.copy()
    else:
        s.x = np.array(x0)
        r = b - A @ x0
    s.r = M @ r
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, _EyeR())
    s.Ar = A @ s.r
    s.rAr = inner(s.r, s.Ar)
    s.p = s.r.copy()
    s.Ap = s.Ar.copy()

    def step(k, crit):
        MAp = M @ s.Ap
        alpha = s.rAr / _nz(inner(s.Ap, MAp))
        s.x += alpha * s.p
        s.r -= alpha * MAp
        s.Ar = A @ s.r
        old, s.rAr = s.rAr, inner(s.r, s.Ar)
        beta = s.rAr / _nz(old)
        s.p = s.r + beta * s.p
        s.Ap = s.Ar + beta * s.Ap
        return norm(s.r)

    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(
        s, step, norm(s.r), lambda: (s.x, s.r))


# ----------------------------------------------------------------------- GCR --
def gcr(A, b, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """gcr.py:16-97: generalised conjugate residuals, full orthogonalisation (MGS) of A s_i."""
    b = np.asarray(b)
    _shapes(A, b)
    A = _op(A)
    s = _S()
    if x0 is None:
        s.x, s.r = np.zeros_like(b), b.copy()
    else:
        s.x = np.array(x0)
        s.r = b - A @ x0
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, _EyeR())
    S, V = [], []

    def step(k, crit):
        S.append(s.r.copy())
        V.append(A @ S[-1])
        for i in range(k):
            a = inner(V[-1], V[i])
            V[-1] -= a * V[i]
            S[-1] -= a * S[i]
        nb = norm(V[-1])
        V[-1] /= _nz(nb)
        S[-1] /= _nz(nb)
        g = inner(b, V[-1])  # :86 -- b, not r (equal in exact arithmetic for x0 = 0)
        s.x += g * S[-1]
        s.r -= g * V[-1]
        return norm(s.r)

    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(
        s, step, norm(s.r), lambda: (s.x, s.r))


# ----------------------------------------------------------------- Chebyshev --
def chebyshev(A, b, eigenvalue_estimates, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15,
              maxiter=None, callback=None):
    """chebyshev.py:13-99: no inner products in the recurrence (only in the residual norm)."""
    b = np.asarray(b)
    _shapes(A, b)
    A, M = _op(A), _op(M)
    s = _S()
    if x0 is None:
        s.x, s.r = np.zeros_like(b), b.copy()
    else:
        s.x = np.array(x0)
        s.r = b - A @ x0
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, M)
    assert len(eigenvalue_estimates) == 2
    assert eigenvalue_estimates[0] <= eigenvalue_estimates[1]
    lmin, lmax = eigenvalue_estimates
    d = (lmax + lmin) / 2
    c = (lmax - lmin) / 2
    s.alpha, s.p = None, None

    def step(k, crit):
        z = M @ s.r
        if k == 0:
            s.p = z.copy()
            s.alpha = 1.0 / d
        else:
            beta = 0.5 * (c * s.alpha) ** 2
            if k > 1:
                beta *= 0.5
            s.alpha = 1.0 / (d - beta / s.alpha)
            s.p = z + beta * s.p
        s.x += s.alpha * s.p
        s.r -= s.alpha * (A @ s.p)
        return norm(s.r)

    return _Loop(A, b, norm, tol, atol, maxiter, callback).run(
        s, step, norm(s.r), lambda: (s.x, s.r))


# -------------------------------------------------------------------- SYMMLQ --
def symmlq(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """symmlq.py:15-161 (Paige/Saunders): Lanczos + LQ; the iterate x is the LQ point, what is
    checked and returned is the CG point x + (zeta / c) w_bar.  ``resnorms`` holds the norm of the
    unnormalised Lanczos vector r (symmlq.py:156), not of a residual; the in-loop callback sees
    (CG point, that r).  No guard on beta = 0 (a zero right-hand side divides by zero there)."""
    b = np.asarray(b)
    _shapes(A, b)
    A, M = _op(A), _op(M)
    inner = default_inner(b.shape) if inner is None else inner
    norm = _make_norm(inner, _EyeR())
    s = _S()
    # rotation / zeta history: cur, prev, prev2  (the reference's lists [0], [-1], [-2])
    s.zeta = [None, 0.0, None]
    s.c = [1.0, 1.0, None]
    s.s = [0.0, 0.0, None]
    s.u_old = np.zeros_like(b)
    s.v_old = np.zeros_like(b)
    if x0 is None:
        s.x, s.r = np.zeros_like(b), b.copy()
    else:
        s.x = np.array(x0)
        s.r = b - A @ x0
    first = norm(s.r)
    s.z = M @ s.r
    s.beta = np.sqrt(inner(s.r, s.z))
    beta1 = s.beta
    s.v = s.r / s.beta
    s.u = s.z / s.beta
    s.w_bar = s.u.copy()

    def cg_point():
        zc = s.zeta[0] / np.where(s.c[0] != 0.0, s.c[0], 1.0e-15)
        return s.x + zc * s.w_bar

    def step(k, crit):
        if k > 0:
            s.v_old, s.u_old = s.v.copy(), s.u.copy()
            s.v = s.r * (1.0 / s.beta)
            s.u = s.z * (1.0 / s.beta)
            w = s.c[0] * s.w_bar + s.s[0] * s.u
            s.w_bar = -s.s[0] * s.w_bar + s.c[0] * s.u
            s.x += s.zeta[0] * w
            s.zeta[2], s.zeta[1] = s.zeta[1], s.zeta[0]
        s.r = A @ s.u  # Lanczos
        alpha = inner(s.u, s.r)
        s.z = M @ s.r
        s.r = s.r - alpha * s.v - s.beta * s.v_old
        s.z = s.z - alpha * s.u - s.beta * s.u_old
        beta_old = s.beta
        s.beta = np.sqrt(inner(s.r, s.z))
        s.c[2], s.c[1] = s.c[1], s.c[0]
        s.s[2], s.s[1] = s.s[1], s.s[0]
        gamma_bar = s.c[1] * alpha - s.c[2] * s.s[1] * beta_old
        gamma = np.sqrt(gamma_bar * gamma_bar + s.beta * s.beta)
        delta = s.s[1] * alpha + s.c[2] * s.c[1] * beta_old
        epsilon = s.s[2] * beta_old
        s.c[0] = gamma_bar / gamma
        s.s[0] = s.beta / gamma
        if k == 0:
            s.zeta[0] = beta1 / gamma
        else:
            s.zeta[0] = -(delta * s.zeta[1] + epsilon * s.zeta[2]) / gamma
        if callback is not None:
            callback(cg_point(), s.r)
        return norm(s.r)

    # the first callback gets (x, r) before the loop (symmlq.py:63-64); the driver's in-loop
    # callback is replaced by the one inside step (it needs the CG point)
    if callback is not None:
        callback(s.x, s.r)
    loop = _Loop(A, b, norm, tol, atol, maxiter, None)
    return loop.run(s, step, first, lambda: (), xout=cg_point)
