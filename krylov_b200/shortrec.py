"""Short-recurrence solvers next to cg / minres / gmres: ``bicgstab``, ``cgs``,
``bicg``, ``qmr``, ``cgne``, ``cgnr``, ``cgr``, ``gcr``, ``chebyshev``, ``symmlq`` with the
reference's signatures, defaults, stopping rule and ``Info`` (reference bicgstab.py, cgs.py,
bicg.py, qmr.py, cgne.py, cgnr.py, cgr.py, gcr.py, chebyshev.py, symmlq.py).

They run on the *general* device path: every vector statement of the reference loop is one kernel
launch through the C ABI (kb_spmv incl. the transposed matrix for ``rmatvec``, kb_dot, kb_axpy,
kb_xpby, kb_lincomb, kb_div_scale) on (n, k) CUDA tensors.  With the default inner product the
per-column scalars stay on the device too (``_alg.DevScalar``: every scalar statement of the
reference is one ``kb_scalar_op`` launch with the host's IEEE result); the only read-back per
iteration is the residual norm for the stopping test.  A user ``inner`` keeps host scalars as in
the reference.  No statement is evaluated on the CPU and there is no CPU fallback.

The driver all of them share in the reference (initial residual, ``max(tol * r0, atol)``,
explicit-residual confirmation, ``maxiter``, callback, Info) is written once here (``_Drive``).
"""
from __future__ import annotations

import numpy as np
import torch

from ._alg import Alg, nz, to_host
from .operators import Info, Problem

__all__ = ["bicgstab", "cgs", "bicg", "qmr", "cgne", "cgnr", "cgr", "gcr", "chebyshev", "symmlq"]

_AHEAD = 4  # iterations enqueued ahead of the host's read-back (gated on the device)
_INT_MAX = 2**31 - 1


class _Drive:
    def __init__(self, A, b, x0, inner, tol, atol, maxiter, callback):
        self.prob = prob = Problem(A, b, x0)
        self.alg = alg = Alg(prob, inner, lazy=True)
        self.A, self.b = prob.A, prob.b
        self.tol, self.atol, self.maxiter, self.callback = tol, atol, maxiter, callback
        self.x = prob.x0  # Problem clones a user x0; zeros otherwise
        # r0 = b - A x0 (b itself for the default x0, as the reference does without a product)
        self.r0 = self.b.clone() if x0 is None else alg.residual(self.A, self.b, self.x)

    def op(self, M):
        return self.prob.operator(M)

    def norm_with(self, W):
        """x -> sqrt(<x, W x>) on the host, complaining like the reference about complex values
        (cannot happen for real operators, kept for user inner products)."""
        alg = self.alg

        def norm(v):
            return np.sqrt(alg.inner(v, alg.apply(W, v)))

        return norm

    def user(self, *vecs):
        return tuple(self.prob.to_user(v) for v in vecs)

    # ---- state the look-ahead has to be able to roll back (see run)
    def track(self, *containers):
        """The dicts / lists in which a solver keeps what changes from step to step (buffer
        references, DevScalars, per-step lists).  Tensors are only ever modified by gated kernels,
        so restoring the references restores the state."""
        self._tracked = containers

    def _snapshot(self):
        snap = [self.x]
        for c in self._tracked:
            if isinstance(c, dict):
                snap.append({k: (list(v) if isinstance(v, list) else v) for k, v in c.items()})
            else:
                snap.append(list(c))
        return snap

    def _restore(self, snap):
        self.x = snap[0]
        for c, saved in zip(self._tracked, snap[1:]):
            if isinstance(c, dict):
                c.clear()
                c.update({k: (list(v) if isinstance(v, list) else v) for k, v in saved.items()})
            else:
                c[:] = saved

    def run(self, norm, step, first, cb_vecs, xout=None, ahead=True):
        """step(k, crit) advances self.x and returns the new residual norm, or ("leave", resnorm)
        to finish successfully at once (bicgstab.py:119-122).  cb_vecs() gives the callback's
        arguments in the caller's array kind.  xout(): the point that is checked and returned when
        it is not self.x itself (symmlq's CG point; the step then calls the callback on its own).

        With device-resident scalars (default inner product, no callback, ``ahead``) up to
        ``_AHEAD`` iterations are enqueued before the host reads anything back: ``kb_record``
        applies the loop condition on the device and the workspace gate turns every launch behind
        the step that met the criterion into a no-op; the host then rolls its own references back
        to that step (``track``) and applies the reference's explicit-residual confirmation."""
        prob, alg = self.prob, self.alg
        point = xout if xout is not None else (lambda: self.x)
        lookahead = (alg.lazy and ahead and self.callback is None and getattr(self, "_tracked", None)
                     is not None)
        with prob.on_device():
            if self.callback is not None:
                self.callback(*cb_vecs())
            res = [to_host(first)]
            crit = np.maximum(self.tol * res[0], self.atol)
            k, ok, xo = 0, False, None
            if lookahead:
                dev, kk = prob.device, prob.k
                crit_d = torch.from_numpy(np.array(
                    np.broadcast_to(np.asarray(crit, dtype=np.float64).reshape(-1), (kk,)))).to(dev)
                hist = torch.zeros((_AHEAD, kk), dtype=torch.float64, device=dev)
                stop_at = torch.full((1,), _INT_MAX, dtype=torch.int32, device=dev)
            while True:
                if np.all(res[-1] <= crit):
                    xo = point()
                    res[-1] = to_host(norm(alg.residual(self.A, self.b, xo)))
                    if np.all(res[-1] <= crit):
                        ok = True
                        break
                if k == self.maxiter:
                    xo = point() if xout is not None else xo
                    break
                if not lookahead:
                    out = step(k, crit)
                    if isinstance(out, tuple):
                        res[-1] = to_host(out[1])
                        ok = True
                        break
                    if self.callback is not None and xout is None:
                        self.callback(*cb_vecs())
                    res.append(to_host(out))
                    k += 1
                    continue
                nb = min(_AHEAD, self.maxiter - k)
                stop_at.fill_(_INT_MAX)
                hist_ptr = hist.data_ptr() - (k + 1) * kk * 8  # history row k + 1 == hist[0]
                snaps = []
                for j in range(nb):
                    snaps.append(self._snapshot())
                    alg.ops.gate(stop_at, k + j)
                    out = step(k + j, crit)
                    alg.ops.record(k + j + 1, alg.coef(out), crit_d, hist_ptr, stop_at)
                alg.ops.gate(None, 0)
                stop = int(stop_at.item())  # one read-back per batch
                done = min(stop, k + nb) - k
                rows = hist[:done].cpu().numpy()
                res.extend(rows[j].copy() for j in range(done))
                if done < nb:  # the launches behind step k + done were no-ops: undo the host side
                    self._restore(snaps[done])
                k += done
        prob.launches = alg.ops.launches
        if xout is not None:
            xk = prob.to_user(xo) if xo is not None else None
        else:
            xk = prob.to_user(self.x)
        return (xk if ok else None), Info(ok, xk, k, [prob.scalars_to_user(r) for r in res])


def bicgstab(A, b, Ml=None, Mr=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None,
             callback=None):
    """reference bicgstab.py:24-144."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, Aop = d.alg, d.A
    Ml, Mr = d.op(Ml), d.op(Mr)
    norm = d.norm_with(Ml)
    shadow = d.r0
    s = {"r": d.r0.clone(), "rho": 1.0, "alpha": 1.0, "omega": 1.0,
         "p": torch.zeros_like(d.b), "v": torch.zeros_like(d.b)}

    def step(k, crit):
        rho_old, s["rho"] = s["rho"], alg.inner(shadow, s["r"])
        beta = s["rho"] * s["alpha"] / nz(rho_old * s["omega"])
        t = alg.lincomb(s["p"], None, s["v"], -s["omega"], out=s["p"])  # p - omega v
        s["p"] = alg.lincomb(s["r"], None, t, beta, out=t)                          # r + beta (.)
        y = alg.apply(Mr, alg.apply(Ml, s["p"]))
        s["v"] = Aop(y)
        s["alpha"] = s["rho"] / nz(alg.inner(shadow, s["v"]))
        half_r = alg.lincomb(s["r"], None, s["v"], -s["alpha"])
        half_x = alg.lincomb(d.x, None, y, s["alpha"])
        rn = to_host(norm(alg.apply(Ml, alg.residual(Aop, d.b, d.x))))  # :117-122, the OLD x
        if np.all(rn <= crit):
            return ("leave", rn)
        Ml_s = alg.apply(Ml, half_r)
        z = alg.apply(Mr, Ml_s)
        tt = Aop(z)
        Ml_t = alg.apply(Ml, tt)
        s["omega"] = alg.inner(Ml_t, Ml_s) / nz(alg.inner(Ml_t, Ml_t))
        d.x = alg.lincomb(half_x, None, z, s["omega"], out=half_x)
        s["r"] = alg.lincomb(half_r, None, tt, -s["omega"], out=half_r)
        return norm(s["r"])

    # the mid-step test on the old x (bicgstab.py:117-122) needs the host inside every step
    return d.run(norm, step, norm(d.r0), lambda: d.user(d.x, s["r"]), ahead=False)


def cgs(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """reference cgs.py:24-117."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, Aop = d.alg, d.A
    M = d.op(M)
    norm = d.norm_with(M)
    shadow = d.r0
    s = {"r": d.r0.clone(), "rho": 1.0, "p": torch.zeros_like(d.b), "q": torch.zeros_like(d.b)}

    def step(k, crit):
        rho_old, s["rho"] = s["rho"], alg.inner(shadow, s["r"])
        beta = s["rho"] / nz(rho_old)
        u = alg.lincomb(s["r"], None, s["q"], beta)
        alg.lincomb(s["q"], None, s["p"], beta, out=s["p"])  # q + beta p
        alg.lincomb(u, None, s["p"], beta, out=s["p"])       # u + beta (.)
        v = Aop(alg.apply(M, s["p"]))
        alpha = s["rho"] / nz(alg.inner(shadow, v))
        s["q"] = alg.lincomb(u, None, v, -alpha, out=s["q"])
        uq = alg.apply(M, alg.add(u, s["q"]))
        alg.axpy(d.x, alpha, uq)
        alg.axpy(s["r"], alpha, Aop(uq), sign=-1.0)
        return norm(s["r"])

    d.track(s)
    return d.run(norm, step, norm(s["r"]), lambda: d.user(d.x, s["r"]))


def bicg(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """reference bicg.py:25-116.  The callback receives (x, [r, r~]) like the reference's
    ``np.array([r0, r1])`` (stacked on the host for NumPy callers)."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, prob, Aop = d.alg, d.prob, d.A
    AH = prob.adjoint(Aop)
    M = d.op(M)
    MH = prob.adjoint(M)
    norm = d.norm_with(M)
    r = [d.r0.clone(), d.r0.clone()]
    Mr0 = alg.apply(M, r[0])
    p = [Mr0.clone(), alg.apply(MH, r[1]).clone()]
    s = {"rMr": alg.inner(r[1], Mr0)}

    def step(k, crit):
        Ap = Aop(p[0])
        AHp = alg.apply(AH, p[1])
        alpha = s["rMr"] / nz(alg.inner(p[1], Ap))
        alg.axpy(d.x, alpha, p[0])
        alg.axpy(r[0], alpha, Ap, sign=-1.0)
        alg.axpy(r[1], alpha, AHp, sign=-1.0)
        Mr = alg.apply(M, r[0])
        old, s["rMr"] = s["rMr"], alg.inner(r[1], Mr)
        beta = s["rMr"] / nz(old)
        rn = np.sqrt(alg.inner(r[0], Mr))
        alg.xpby(p[0], Mr, beta)
        alg.xpby(p[1], alg.apply(MH, r[1]), beta)
        return rn

    def cb():  # (x, [r, r~]) like the reference's np.array([r0, r1])
        x_u, r0_u, r1_u = d.user(d.x, r[0], r[1])
        pair = torch.stack([r0_u, r1_u]) if prob.is_torch else np.array([r0_u, r1_u])
        return x_u, pair

    d.track(s, r, p)
    return d.run(norm, step, norm(r[0]), cb)


def qmr(A, b, Ml=None, Mr=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None,
        callback=None):
    """reference qmr.py:22-160."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, prob, Aop = d.alg, d.prob, d.A
    AH = prob.adjoint(Aop)
    Ml, Mr = d.op(Ml), d.op(Mr)
    MlH, MrH = prob.adjoint(Ml), prob.adjoint(Mr)
    norm = d.norm_with(Ml)
    s = {"r": d.r0.clone()}
    first = norm(s["r"])
    s["vt"] = s["r"].clone()
    s["y"] = alg.apply(Ml, s["vt"])
    if s["y"] is s["vt"]:
        s["y"] = s["y"].clone()
    s["rho"] = norm(s["y"])
    s["wt"] = s["r"].clone()
    s["z"] = alg.apply(MrH, s["wt"])
    if s["z"] is s["wt"]:
        s["z"] = s["z"].clone()
    s["xi"] = norm(s["z"])
    s.update(gamma=1.0, eta=-1.0, theta=1.0, eps=1.0, p=None, q=None, d=None, s=None)

    def step(k, crit):
        v = alg.div(s["vt"], s["rho"])
        s["y"] = alg.div(s["y"], s["rho"], out=s["y"])
        w = alg.div(s["wt"], s["xi"])
        s["z"] = alg.div(s["z"], s["xi"], out=s["z"])
        delta = alg.inner(s["z"], s["y"])
        yt = alg.apply(Mr, s["y"])
        zt = alg.apply(MlH, s["z"])
        if k == 0:
            s["p"], s["q"] = yt.clone(), zt.clone()
        else:
            de = delta / nz(s["eps"])
            alg.lincomb(yt, None, s["p"], -(s["xi"] * de), out=s["p"])
            alg.lincomb(zt, None, s["q"], -(s["rho"] * de), out=s["q"])
        Ap = Aop(s["p"])
        s["eps"] = alg.inner(s["q"], Ap)
        beta = s["eps"] / nz(delta)
        s["vt"] = alg.lincomb(Ap, None, v, -beta, out=v)
        s["y"] = alg.apply(Ml, s["vt"])
        if s["y"] is s["vt"]:
            s["y"] = s["y"].clone()
        rho_old, s["rho"] = s["rho"], norm(s["y"])
        s["wt"] = alg.lincomb(alg.apply(AH, s["q"]), None, w, -beta, out=w)
        s["z"] = alg.apply(MrH, s["wt"])
        if s["z"] is s["wt"]:
            s["z"] = s["z"].clone()
        s["xi"] = norm(s["z"])
        gamma_old, theta_old = s["gamma"], s["theta"]
        s["theta"] = s["rho"] / nz(gamma_old * np.abs(beta))
        s["gamma"] = 1 / np.sqrt(1 + s["theta"] ** 2)
        s["eta"] = -s["eta"] * rho_old * s["gamma"] ** 2 / nz(beta * gamma_old ** 2)
        if k == 0:
            s["d"] = alg.lincomb(s["p"], s["eta"])
            s["s"] = alg.lincomb(Ap, s["eta"])
        else:
            c2 = (theta_old * s["gamma"]) ** 2
            alg.lincomb(s["p"], s["eta"], s["d"], c2, out=s["d"])
            alg.lincomb(Ap, s["eta"], s["s"], c2, out=s["s"])
        alg.axpy(d.x, 1.0, s["d"])
        alg.axpy(s["r"], 1.0, s["s"], sign=-1.0)
        return norm(s["r"])

    d.track(s)
    return d.run(norm, step, first, lambda: d.user(d.x, s["r"]))


class _NormalEq:
    """A A^H (cgne.py:7-15) or A^H A (cgnr.py:5-12) applied on the device: two sparse products,
    the intermediate never leaves HBM.  Seen by ``cg`` as a duck-typed operator whose
    ``device_apply`` is called with device tensors."""

    def __init__(self, csr, outer):
        self.csr_A, self.outer = csr, outer
        self.shape = csr.shape
        self.dtype = np.dtype(np.float64)

    def device_apply(self, x):
        if self.outer:
            return self.csr_A.matvec_device(self.csr_A.T.matvec_device(x))
        return self.csr_A.T.matvec_device(self.csr_A.matvec_device(x))

    def __matmul__(self, x):  # host callers (not used by the solvers)
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)).cuda()
        return self.device_apply(t.reshape(t.shape[0], -1)).reshape(t.shape).cpu().numpy()


def _matrix_or_raise(A, b):
    from .operators import to_csr_or_none

    dev = b.device if isinstance(b, torch.Tensor) and b.is_cuda else torch.device(
        "cuda", torch.cuda.current_device())
    csr = to_csr_or_none(A, dev)
    if csr is None or getattr(csr, "is_dist_csr", False):
        raise NotImplementedError("cgne / cgnr need A as a (single-GPU) matrix")
    return csr


def cgne(A, b, *args, **kwargs):
    """reference cgne.py:18-45: A A^H y = b, x = A^H y."""
    from .cg import cg
    from .device import require_cuda

    require_cuda()
    csr = _matrix_or_raise(A, b)
    sol, info = cg(_NormalEq(csr, True), b, *args, **kwargs)
    is_torch = isinstance(info.xk, torch.Tensor)
    y = info.xk if is_torch else torch.as_tensor(np.asarray(info.xk)).to(csr.device)
    shape = tuple(y.shape)
    xk = csr.T.matvec_device(y.reshape(shape[0], -1)).reshape(shape)
    if not is_torch:
        xk = xk.cpu().numpy()
    return (xk if sol is not None else None), Info(
        info.success, xk, info.numsteps, info.resnorms, info.num_operations, info.arnoldi)


def cgnr(A, b, *args, **kwargs):
    """reference cgnr.py:15-21: A^H A x = A^H b."""
    from .cg import cg
    from .device import as_device_matrix, require_cuda

    require_cuda()
    csr = _matrix_or_raise(A, b)
    is_torch = isinstance(b, torch.Tensor)
    bd = as_device_matrix(b, csr.device)
    shape = tuple(bd.shape)
    rhs = csr.T.matvec_device(bd.reshape(shape[0], -1)).reshape(shape)
    return cg(_NormalEq(csr, False), rhs if is_torch else rhs.cpu().numpy(), *args, **kwargs)


def cgr(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """reference cgr.py:14-100."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, Aop = d.alg, d.A
    M = d.op(M)
    norm = d.norm_with(None)
    r = alg.apply(M, d.r0)
    s = {"r": r.clone() if r is d.r0 else r}
    s["Ar"] = Aop(s["r"])
    s["rAr"] = alg.inner(s["r"], s["Ar"])
    s["p"] = s["r"].clone()
    s["Ap"] = s["Ar"].clone()

    def step(k, crit):
        MAp = alg.apply(M, s["Ap"])
        alpha = s["rAr"] / nz(alg.inner(s["Ap"], MAp))
        alg.axpy(d.x, alpha, s["p"])
        alg.axpy(s["r"], alpha, MAp, sign=-1.0)
        s["Ar"] = Aop(s["r"])
        old, s["rAr"] = s["rAr"], alg.inner(s["r"], s["Ar"])
        beta = s["rAr"] / nz(old)
        alg.xpby(s["p"], s["r"], beta)
        alg.xpby(s["Ap"], s["Ar"], beta)
        return norm(s["r"])

    d.track(s)
    return d.run(norm, step, norm(s["r"]), lambda: d.user(d.x, s["r"]))


def gcr(A, b, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """reference gcr.py:16-97."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, Aop = d.alg, d.A
    norm = d.norm_with(None)
    s = {"r": d.r0.clone()}
    S, V = [], []

    def step(k, crit):
        S.append(s["r"].clone())
        V.append(Aop(S[-1]))
        for i in range(k):  # modified Gram-Schmidt, gcr.py:75-79
            a = alg.inner(V[-1], V[i])
            alg.axpy(V[-1], a, V[i], sign=-1.0)
            alg.axpy(S[-1], a, S[i], sign=-1.0)
        nb = norm(V[-1])
        alg.div(V[-1], nb, out=V[-1])
        alg.div(S[-1], nb, out=S[-1])
        g = alg.inner(d.b, V[-1])  # gcr.py:86 -- b, not r
        alg.axpy(d.x, g, S[-1])
        alg.axpy(s["r"], g, V[-1], sign=-1.0)
        return norm(s["r"])

    d.track(s, S, V)
    return d.run(norm, step, norm(s["r"]), lambda: d.user(d.x, s["r"]))


def chebyshev(A, b, eigenvalue_estimates, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15,
              maxiter=None, callback=None):
    """reference chebyshev.py:13-99."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, Aop = d.alg, d.A
    M = d.op(M)
    norm = d.norm_with(M)
    assert len(eigenvalue_estimates) == 2
    assert eigenvalue_estimates[0] <= eigenvalue_estimates[1]
    lmin, lmax = eigenvalue_estimates
    dd = (lmax + lmin) / 2
    c = (lmax - lmin) / 2
    s = {"r": d.r0.clone(), "alpha": None, "p": None}

    def step(k, crit):
        z = alg.apply(M, s["r"])
        if k == 0:
            s["p"] = z.clone()
            s["alpha"] = 1.0 / dd
        else:
            beta = 0.5 * (c * s["alpha"]) ** 2
            if k > 1:
                beta *= 0.5
            s["alpha"] = 1.0 / (dd - beta / s["alpha"])
            alg.xpby(s["p"], z, beta)
        alg.axpy(d.x, s["alpha"], s["p"])
        alg.axpy(s["r"], s["alpha"], Aop(s["p"]), sign=-1.0)
        return norm(s["r"])

    d.track(s)
    return d.run(norm, step, norm(s["r"]), lambda: d.user(d.x, s["r"]))


def symmlq(A, b, M=None, x0=None, inner=None, tol=1e-5, atol=1.0e-15, maxiter=None, callback=None):
    """reference symmlq.py:15-161.  As in the reference, ``resnorms`` holds the norm of the
    unnormalised Lanczos vector and what is checked / returned is the CG point."""
    d = _Drive(A, b, x0, inner, tol, atol, maxiter, callback)
    alg, Aop = d.alg, d.A
    M = d.op(M)
    norm = d.norm_with(None)
    s = {"zeta": [None, 0.0, None], "c": [1.0, 1.0, None], "s": [0.0, 0.0, None],
         "u_old": torch.zeros_like(d.b), "v_old": torch.zeros_like(d.b), "r": d.r0.clone()}
    first = norm(s["r"])
    s["z"] = alg.apply(M, s["r"])
    s["beta"] = np.sqrt(alg.inner(s["r"], s["z"]))
    beta1 = s["beta"]
    s["v"] = alg.div(s["r"], s["beta"])
    s["u"] = alg.div(s["z"], s["beta"])
    s["w_bar"] = s["u"].clone()

    def cg_point():
        c0 = s["c"][0]
        zc = s["zeta"][0] / nz(c0, 1.0e-15)
        return alg.lincomb(d.x, None, s["w_bar"], zc)

    def step(k, crit):
        if k > 0:
            s["v_old"], s["u_old"] = s["v"], s["u"]
            rb = 1.0 / s["beta"]
            s["v"] = alg.lincomb(s["r"], rb)
            s["u"] = alg.lincomb(s["z"], rb)
            c0, s0 = s["c"][0], s["s"][0]
            w = alg.lincomb(s["w_bar"], c0, s["u"], s0)
            alg.lincomb(s["w_bar"], -s0, s["u"], c0, out=s["w_bar"])
            alg.axpy(d.x, s["zeta"][0], w)
            s["zeta"][2], s["zeta"][1] = s["zeta"][1], s["zeta"][0]
        r = Aop(s["u"])  # Lanczos
        alpha = alg.inner(s["u"], r)
        z = alg.apply(M, r)
        if z is r:
            z = r.clone()
        alg.axpy(r, alpha, s["v"], sign=-1.0)
        alg.axpy(r, s["beta"], s["v_old"], sign=-1.0)
        alg.axpy(z, alpha, s["u"], sign=-1.0)
        alg.axpy(z, s["beta"], s["u_old"], sign=-1.0)
        s["r"], s["z"] = r, z
        beta_old = s["beta"]
        s["beta"] = np.sqrt(alg.inner(r, z))
        c, sn, zeta = s["c"], s["s"], s["zeta"]
        c[2], c[1] = c[1], c[0]
        sn[2], sn[1] = sn[1], sn[0]
        gamma_bar = c[1] * alpha - c[2] * sn[1] * beta_old
        gamma = np.sqrt(gamma_bar * gamma_bar + s["beta"] * s["beta"])
        delta = sn[1] * alpha + c[2] * c[1] * beta_old
        epsilon = sn[2] * beta_old
        c[0] = gamma_bar / gamma
        sn[0] = s["beta"] / gamma
        zeta[0] = beta1 / gamma if k == 0 else -(delta * zeta[1] + epsilon * zeta[2]) / gamma
        if callback is not None:
            callback(*d.user(cg_point(), r))
        return norm(r)

    d.track(s)
    return d.run(norm, step, first, lambda: d.user(d.x, s["r"]), xout=cg_point)
