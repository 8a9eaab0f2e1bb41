"""Arnoldi builders on the device -- drop-ins for ``krylov.ArnoldiMGS``,
``krylov.ArnoldiLanczos`` and ``krylov.ArnoldiHouseholder`` (arnoldi.py:33-281).

``_DevLanczos`` / ``_DevMGS`` are the general builders used by the
preconditioned solver paths: vectors are (n, k) CUDA tensors, every vector
statement of the reference is one kernel, the few scalars are host floats.
``_DevHouseholder`` keeps *everything* on the device (reflector parameters
included); reflectors are stored zero-padded to full length so one
dot/axpy kernel pair serves every tail, and an axpy is fused with the next
reflector's dot (32 B/element per reflector application).

The public classes wrap these for NumPy / torch callers and keep the
reference's attributes (``V``, ``P``, ``iter``, ``is_invariant``, ``dtype``).
"""
from __future__ import annotations

import numpy as np
import torch

from ._alg import Alg, nz
from .errors import ArgumentError
from .operators import Problem

_INVARIANT_MSG = "Krylov subspace was found to be invariant in the previous iteration."


# ---------------------------------------------------------------------------
class _DevLanczos:
    """arnoldi.py:203-281; ``chain`` lists the operators in application order."""

    def __init__(self, alg, chain, p, M=None, Mv=None, Mv_norm=None):
        self.alg, self.chain, self.M = alg, chain, M
        k = alg.prob.k
        self.num_iter = 0
        self.h = np.zeros((3, k))
        self.is_invariant = False
        v = alg.apply(M, p) if Mv is None else Mv
        self.vnorm = np.sqrt(alg.inner(p, v)) if Mv_norm is None else Mv_norm
        d = nz(self.vnorm)
        self.p_old = None
        self.p = alg.div(p, d)
        self.v = self.p if v is p else alg.div(v, d)

    def step(self):
        if self.is_invariant:
            raise ArgumentError(_INVARIANT_MSG)
        alg = self.alg
        w = alg.apply_chain(self.chain, self.v)  # arnoldi.py:244
        if self.num_iter > 0:
            self.h[0] = self.h[2]
            alg.axpy(w, self.h[0], self.p_old, sign=-1.0)
        a = alg.inner(self.v, w)  # arnoldi.py:252
        self.h[1] = a
        alg.axpy(w, a, self.p, sign=-1.0)
        Mw = alg.apply(self.M, w)
        self.h[2] = np.sqrt(alg.inner(w, Mw))  # arnoldi.py:266-267
        if np.all(self.h[2] <= 1.0e-14):
            self.is_invariant = True
            self.v = None
            self.p = None
        else:
            d = nz(self.h[2])
            self.p_old = self.p
            self.p = alg.div(w, d)
            self.v = self.p if Mw is w else alg.div(Mw, d)
        self.num_iter += 1
        return self.v, self.h, self.p


class _DevMGS:
    """arnoldi.py:107-200 with ``num_reorthos`` sweeps of modified Gram-Schmidt."""

    def __init__(self, alg, chain, p, num_reorthos=1, M=None, Mv=None, Mv_norm=None):
        self.alg, self.chain, self.M = alg, chain, M
        self.num_reorthos = num_reorthos
        self.iter = 0
        self.is_invariant = False
        v = alg.apply(M, p) if Mv is None else Mv
        self.vnorm = np.sqrt(alg.inner(p, v)) if Mv_norm is None else Mv_norm
        d = nz(self.vnorm)
        self.P = [alg.div(p, d)]
        self.V = [self.P[0] if v is p else alg.div(v, d)]

    def step(self):
        if self.is_invariant:
            raise ArgumentError(_INVARIANT_MSG)
        alg = self.alg
        k = self.iter
        w = alg.apply_chain(self.chain, self.V[k])  # arnoldi.py:176
        h = np.zeros((k + 2, alg.prob.k))
        for _ in range(self.num_reorthos):
            for j in range(k + 1):  # arnoldi.py:157-162
                a = alg.inner(self.V[j], w)
                h[j] += a
                alg.axpy(w, a, self.P[j], sign=-1.0)
        Mw = alg.apply(self.M, w)
        h[k + 1] = np.sqrt(alg.inner(w, Mw))
        if np.all(h[k + 1] <= 1.0e-14):
            self.is_invariant = True
            v = None
        else:
            d = nz(h[k + 1])
            self.P.append(alg.div(w, d))
            v = self.P[-1] if Mw is w else alg.div(Mw, d)
            self.V.append(v)
        self.h = h
        self.iter += 1
        return v, h


class _DevHouseholder:
    """arnoldi.py:33-104 (Walker's Householder Arnoldi), one right-hand side,
    Euclidean inner product, no M.  Device-resident: no host read per step.

    After ``step()``: ``self.h_dev[: k+2]`` holds the new Hessenberg column on
    the device; ``self.V[k+1]`` the new basis vector (meaningless if the step
    found the subspace invariant -- the caller learns that from the scalar
    kernel's flag or from ``h_dev``)."""

    def __init__(self, alg, chain, v0, max_steps):
        prob, ops = alg.prob, alg.ops
        if prob.k != 1:
            raise AssertionError(
                "Householder only works for quasi-1D vectors for now. "
                f"Input vector has shape {prob.user_shape}.")
        self.alg, self.chain = alg, chain
        self.n = prob.n
        self.iter = 0
        self.is_invariant = False
        self.hv = []      # reflector vectors, zero-padded to length n
        self.params = []  # device (5,): alpha, beta, xnorm, v0, divisor
        self.tau = ops.slots(2)
        self.scratch = ops.slots(1)
        self.h_dev = torch.zeros((max_steps + 2, 1), dtype=torch.float64, device=prob.device)
        self._make(v0, 0)
        ops.dot(v0, v0, self.scratch[0])
        self.vnorm = np.sqrt(self.scratch[0].cpu().numpy().copy())  # set-up only
        self.V = [alg.div(v0, nz(self.vnorm))]
        # the chain is a single CSR matrix -> its product can carry the first dot
        real = [c for c in chain if c is not None]
        self._csr = real[0].csr if len(real) == 1 and getattr(real[0], "csr", None) else None

    def _make(self, x, off):
        ops = self.alg.ops
        v = ops.vec(zero=False)
        pr = torch.empty((5,), dtype=torch.float64, device=v.device)
        ops.house_make(off, x, v, pr, self.scratch[0])
        self.hv.append(v)
        self.params.append(pr)

    def step(self):
        if self.is_invariant:
            raise ArgumentError(_INVARIANT_MSG)
        alg, ops = self.alg, self.alg.ops
        k, N = self.iter, self.n
        tau = self.tau
        if self._csr is not None:
            w = ops.vec(zero=False)
            ops.spmv(self._csr, self.V[k], w, dot=1, w=self.hv[0], out=tau[0])
        else:
            w = alg.apply_chain(self.chain, self.V[k])
            ops.dot(self.hv[0], w, tau[0])
        t = 0
        for j in range(k + 1):  # arnoldi.py:75-77
            nxt = self.hv[j + 1] if j < k else None
            ops.axpy_dot(tau[t], self.hv[j], w, dot=1 if nxt is not None else 0, z=nxt,
                         out=tau[1 - t], scale=self.params[j][1:2])
            ops.poke(0, w, j, s=self.params[j][0:1])  # w[j] *= conj(alpha_j)
            t = 1 - t
        if k < N - 1:
            self._make(w, k + 1)  # arnoldi.py:81-82
            hv, pr = self.hv[k + 1], self.params[k + 1]
            ops.dot(hv, w, tau[0])
            self.h_dev[: k + 1].copy_(w[: k + 1])
            ops.house_hlast(w, k + 1, hv, pr, tau[0], self.h_dev[k + 1])  # arnoldi.py:83-85
            # new basis vector H_0 ... H_{k+1} e_{k+1} * alpha   (arnoldi.py:91-96)
            e = ops.vec(zero=True)
            ops.poke(1, e, k + 1, val=1.0)
            ops.poke(2, hv, k + 1, dst=tau[0])  # <v_{k+1}, e_{k+1}>
            t = 0
            for j in range(k + 1, -1, -1):
                nxt = self.hv[j - 1] if j > 0 else None
                ops.axpy_dot(tau[t], self.hv[j], e, dot=1 if nxt is not None else 0, z=nxt,
                             out=tau[1 - t], scale=self.params[j][1:2])
                t = 1 - t
            vnew = ops.vec(zero=False)
            ops.div_scale(vnew, e, pr[0:1])  # alpha = +-1: e/alpha == e*alpha
            self.V.append(vnew)
            self.last_len = k + 2
        else:  # arnoldi.py:97-101
            self.h_dev[:N].copy_(w[:N])
            self.h_dev[N] = 0.0
            self.is_invariant = True
            self.last_len = N + 1
        self.iter += 1


# ---------------------------------------------------------------------------
# public wrappers (reference constructor signatures)
# ---------------------------------------------------------------------------
class _PublicBase:
    def _setup(self, A, v, inner):
        self._prob = Problem(A, v)
        self._alg = Alg(self._prob, inner)
        self.A = A
        self.v = v
        self.dtype = np.dtype(np.float64)
        self.is_invariant = False

    def _u(self, t):
        return None if t is None else self._prob.to_user(t)

    def _h_user(self, h):
        shp = self._prob.user_shape[1:]
        return h.reshape((h.shape[0],) + tuple(shp)) if shp else h[:, 0].copy()

    def __iter__(self):
        return self


class ArnoldiMGS(_PublicBase):
    def __init__(self, A, v, num_reorthos: int = 1, M=None, Mv=None, Mv_norm=None, inner=None):
        self._setup(A, v, inner)
        prob = self._prob
        with torch.cuda.device(prob.device):
            Mop = prob.operator(M)
            Mv_d = None if Mv is None else prob.from_user(Mv)
            nrm = None if Mv_norm is None else np.broadcast_to(
                np.asarray(Mv_norm, dtype=np.float64).reshape(-1), (prob.k,)).copy()
            self._dev = _DevMGS(self._alg, [prob.A], prob.b, num_reorthos, Mop, Mv_d, nrm)
        self.num_reorthos = num_reorthos
        self.iter = 0
        self.vnorm = prob.scalars_to_user(self._dev.vnorm)
        self.V = [self._u(self._dev.V[0])]
        self.P = [self._u(self._dev.P[0])]

    def __next__(self):
        with torch.cuda.device(self._prob.device):
            v, h = self._dev.step()
        self.iter = self._dev.iter
        self.is_invariant = self._dev.is_invariant
        if v is not None:
            self.V.append(self._u(v))
            self.P.append(self._u(self._dev.P[-1]))
        self.h = self._h_user(h)
        return (None if v is None else self.V[-1]), self.h


class ArnoldiLanczos(_PublicBase):
    def __init__(self, A, v, M=None, Mv=None, Mv_norm=None, inner=None):
        self._setup(A, v, inner)
        prob = self._prob
        with torch.cuda.device(prob.device):
            Mop = prob.operator(M)
            Mv_d = None if Mv is None else prob.from_user(Mv)
            nrm = None if Mv_norm is None else np.broadcast_to(
                np.asarray(Mv_norm, dtype=np.float64).reshape(-1), (prob.k,)).copy()
            self._dev = _DevLanczos(self._alg, [prob.A], prob.b, Mop, Mv_d, nrm)
        self.num_iter = 0
        self.vnorm = prob.scalars_to_user(self._dev.vnorm)
        self.v = self._u(self._dev.v)
        self.p = self._u(self._dev.p)
        self.h = self._h_user(self._dev.h)

    def __next__(self):
        with torch.cuda.device(self._prob.device):
            v, h, p = self._dev.step()
        self.num_iter = self._dev.num_iter
        self.is_invariant = self._dev.is_invariant
        self.v, self.p = self._u(v), self._u(p)
        self.h = self._h_user(h)
        return self.v, self.h, self.p


class ArnoldiHouseholder(_PublicBase):
    def __init__(self, A, v, max_steps=None):
        self._setup(A, v, None)
        prob = self._prob
        steps = prob.n if max_steps is None else max_steps
        with torch.cuda.device(prob.device):
            self._dev = _DevHouseholder(self._alg, [prob.A], prob.b, steps)
        self.iter = 0
        self.vnorm = prob.scalars_to_user(self._dev.vnorm)
        self.V = [self._u(self._dev.V[0])]

    def __next__(self):
        dev = self._dev
        with torch.cuda.device(self._prob.device):
            dev.step()
            h = dev.h_dev[: dev.last_len].cpu().numpy()
        self.iter = dev.iter
        v = None
        if dev.is_invariant or h[-1, 0] <= 1.0e-14:  # arnoldi.py:87-89
            dev.is_invariant = True
        else:
            v = self._u(dev.V[-1])
            self.V.append(v)
        self.is_invariant = dev.is_invariant
        return v, self._h_user(h)
