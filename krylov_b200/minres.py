"""Preconditioned MINRES on the device -- drop-in for ``krylov.minres``
(minres.py:28-253): Lanczos three-term recurrence (arnoldi.py:203-281), implicit
QR of the tridiagonal with two stored Givens rotations, two-vector ``W``
recurrence.

Fused path (default inner product; A and any of M / Ml / Mr given as matrices), per step:
  1. ``Av = A v - beta_{k-1} v_old`` fused with ``alpha = <v, Av>``   (SpMV kernel)
  2. ``Av -= alpha v`` fused with ``beta_k^2 = <Av, Av>``             (24 B/elem)
  3. one-block scalar kernel: Givens QR update on the device (dlartg semantics),
     residual norm, convergence flag                                  (no host)
  4. ``z = (v - R0 W0 - R1 W1)/R2; yk += y0 z; v_next = Av/beta_k``   (64 B/elem)
With preconditioners step 1 runs the chain ``Ml A Mr`` (fusion on its last product), step 2's
norm becomes ``<Av, M Av>`` out of M's product, step 4 also writes ``p_next`` (80 B/elem).
The general path mirrors the reference loop with host scalars.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _complex
from ._alg import Alg, nz
import ctypes as C

from ._lib import MinresRunState, MinresState, check, lib
from .arnoldi import _DevLanczos
from .device import Ops, cur_stream, ptr
from .errors import ArgumentError
from .givens import givens
from .operators import Info, Problem

INT_MAX = 2**31 - 1
_BATCH_MIN, _BATCH_MAX = 8, 256
USE_C_LOOP = True  # kb_minres_run where it applies (tests compare both ways)
_INVARIANT_MSG = "Krylov subspace was found to be invariant in the previous iteration."


def minres(A, b, M=None, Ml=None, Mr=None, inner=None, x0=None, tol=1e-5, atol=1.0e-15,
           maxiter=None, callback=None, inner_product=None):
    if inner is None and inner_product is not None:
        inner = inner_product
    if _complex.any_complex(A, b, x0, M, Ml, Mr):  # Hermitian systems: real-equivalent embedding
        return _complex.solve_hermitian(minres, A, b, x0, {"M": M, "Ml": Ml, "Mr": Mr}, inner,
                                        callback, dict(tol=tol, atol=atol, maxiter=maxiter))
    prob = Problem(A, b, x0)
    maxiter = prob.n if maxiter is None else int(maxiter)
    with torch.cuda.device(prob.device):
        if inner is None and prob.A_csr is not None:
            # device-resident path: A and the preconditioners (if any) are matrices
            pre = [prob.operator(op) for op in (M, Ml, Mr)]
            if all(op is None or op.csr is not None for op in pre):
                return _minres_fused(prob, tol, atol, maxiter, callback,
                                     *[None if op is None else op.csr for op in pre])
        return _minres_general(prob, M, Ml, Mr, inner, tol, atol, maxiter, callback)


def _num_operations(k):
    # minres.py:242-249
    return {"A": 1 + k, "M": 2 + k, "Ml": 2 + k, "Mr": 1 + k, "inner": 2 + 2 * k,
            "axpy": 4 + 8 * k}


def _finish(prob, success, xk, k, resn):
    xk_user = prob.to_user(xk)
    resnorms = [prob.scalars_to_user(r) for r in resn]
    return (xk_user if success else None), Info(
        success, xk_user, k, resnorms, num_operations=_num_operations(k))


def _callback_resnorm(prob, callback, xk, rn):
    """minres.py:226-234: the callback gets the residual norm as an array it may
    overwrite in place; what it leaves there is what gets recorded."""
    arr = np.array(prob.scalars_to_user(rn))
    callback(prob.to_user(xk), arr)
    return np.broadcast_to(np.asarray(arr[()], dtype=np.float64).reshape(-1), (prob.k,)).copy()


def _minres_fused(prob, tol, atol, maxiter, callback, M=None, Ml=None, Mr=None):
    """Device-resident MINRES; ``M``/``Ml``/``Mr`` are CSR matrices or None.  With them the
    Lanczos operator is the chain ``Ml A Mr`` (minres.py:136-141; the last product of the
    chain carries the fused ``- beta v_old`` and ``<v, .>``), ``M`` keeps the two bases
    ``V = M P`` (arnoldi.py:250-277) and ``beta^2 = <Av, M Av>`` comes out of M's own product."""
    A, b, x0 = prob.A_csr, prob.b, prob.x0
    n, k, dev = prob.n, prob.k, prob.device
    ops = Ops(n, k, dev, comm=prob.comm)
    chain = [op for op in (Mr, A, Ml) if op is not None]
    Vb = [ops.vec(zero=True), ops.vec(zero=True)]  # v_i in Vb[i % 2], v_{i-1} in the other
    Pb = Vb if M is None else [ops.vec(zero=True), ops.vec(zero=True)]  # p_i likewise
    Wb = [ops.vec(zero=True), ops.vec(zero=True)]  # W0 in Wb[i % 2], W1 in the other
    Av = ops.vec(zero=False)
    MAv = None if M is None else ops.vec(zero=False)
    tmp = [ops.vec(zero=False) for _ in range(min(len(chain) - 1, 2))]
    yk = ops.vec(zero=True)
    sl = ops.slots(5)  # alpha, ww, h2prev, y0, scratch
    g0, g1 = ops.slots(2), ops.slots(2)
    coefs = ops.slots(5)
    stop_at = torch.full((1,), INT_MAX, dtype=torch.int32, device=dev)
    flags = torch.zeros((1,), dtype=torch.int32, device=dev)
    hist = torch.zeros((_BATCH_MAX, k), dtype=torch.float64, device=dev)

    def get_x():  # minres.py:95-98
        xk = torch.empty_like(yk)
        if Mr is None:
            ops.add(xk, x0, yk)
        else:
            ops.spmv(Mr, yk, Av)
            ops.add(xk, x0, Av)
        return xk

    def explicit_norm(z, keep=False):  # minres.py:106-112 / 121-127
        """sqrt(<r, M r>), r = Ml (b - A z).  keep: leave r in Av and M r in MAv."""
        if Ml is None:
            ops.spmv(A, z, Av, mode=2, z=b, dot=(2 if M is None else 0), out=sl[4])
        else:
            ops.spmv(A, z, tmp[0], mode=2, z=b)
            ops.spmv(Ml, tmp[0], Av, dot=(2 if M is None else 0), out=sl[4])
        if M is not None:
            ops.spmv(M, Av, MAv, dot=1, w=Av, out=sl[4])
        return np.sqrt(sl[4].cpu().numpy().copy())

    # Ml r0, M Ml r0, ||.||  (minres.py:121-127); p_0, v_0 = ./nz(norm) (arnoldi.py:229-231)
    ops.gate(None, 0)
    nrm0 = explicit_norm(x0, keep=True)
    sl[3].copy_(torch.from_numpy(nrm0))  # y = [||r0||, 0]   (minres.py:149)
    ops.div_scale(Pb[0], Av, sl[3])
    if M is not None:
        ops.div_scale(Vb[0], MAv, sl[3])
    if callback is not None:
        nrm0 = _callback_resnorm(prob, callback, x0, nrm0)
    resn = [nrm0]
    crit = np.maximum(tol * resn[0], atol)
    crit_d = torch.from_numpy(np.ascontiguousarray(crit)).to(dev)

    st = MinresState(alpha=ptr(sl[0]), ww=ptr(sl[1]), h2prev=ptr(sl[2]), g0=ptr(g0), g1=ptr(g1),
                     y0=ptr(sl[3]), coefs=ptr(coefs), crit=ptr(crit_d), hist=0,
                     stop_at=ptr(stop_at), flags=ptr(flags))

    # one GPU, no preconditioner: the whole batch is enqueued by one C call (kb_minres_run)
    run_c = None
    if USE_C_LOOP and len(chain) == 1 and M is None and prob.comm is None and hasattr(A, "handle"):
        run_c = MinresRunState(A=A.handle, n=n, k=k, Av=ptr(Av), yk=ptr(yk))
        run_c.V[0], run_c.V[1] = ptr(Vb[0]), ptr(Vb[1])
        run_c.W[0], run_c.W[1] = ptr(Wb[0]), ptr(Wb[1])

    batch = 1 if callback is not None else _BATCH_MIN
    kk = 0
    success = False
    xk = None
    while True:
        if np.all(resn[-1] <= crit):  # minres.py:169-175
            if xk is None:
                xk = get_x()
            resn[-1] = explicit_norm(xk)
            if np.all(resn[-1] <= crit):
                success = True
                break
        if kk == maxiter:
            break
        if int(flags.item()) & 1:
            raise ArgumentError(_INVARIANT_MSG)  # arnoldi.py:239-242
        nb = min(batch, maxiter - kk)
        stop_at.fill_(INT_MAX)
        st.hist = hist.data_ptr() - (kk + 1) * k * 8
        if run_c is not None:
            run_c.st = st
            check(lib.kb_minres_run(ops.ws.handle, C.byref(run_c), kk, nb, cur_stream()))
            ops.launches += 3 * nb
        for i in range(kk, kk + nb) if run_c is None else ():
            v, vold = Vb[i % 2], Vb[(i + 1) % 2]
            p, pold = Pb[i % 2], Pb[(i + 1) % 2]
            ops.gate(stop_at, i)
            src = v
            for j, op in enumerate(chain[:-1]):  # Mr, A in front of the fused last product
                ops.spmv(op, src, tmp[j])
                src = tmp[j]
            if i == 0:
                ops.spmv(chain[-1], src, Av, dot=1, w=v, out=sl[0])
            else:  # Av = (Ml A Mr) v - beta_{i-1} p_old   (arnoldi.py:244-249)
                ops.spmv(chain[-1], src, Av, mode=1, z=pold, coef=sl[2], dot=1, w=v, out=sl[0])
            if M is None:
                ops.axpy_dot(sl[0], p, Av, dot=2, out=sl[1])  # arnoldi.py:264-267
                ops.minres_scalar(i, st)  # minres.py:190-228
                ops.minres_update(coefs, v, Wb[i % 2], Wb[(i + 1) % 2], Av, yk, vold)
            else:
                ops.axpy_dot(sl[0], p, Av, dot=0, out=sl[1])  # Av -= alpha p
                ops.spmv(M, Av, MAv, dot=1, w=Av, out=sl[1])  # beta^2 = <Av, M Av>
                ops.minres_scalar(i, st)
                ops.minres_update(coefs, v, Wb[i % 2], Wb[(i + 1) % 2], Av, yk, vold,
                                  MAv=MAv, pnext=pold)
        ops.gate(None, 0)
        s = int(stop_at.item())
        done = min(s, kk + nb) - kk
        rows = hist[:done].cpu().numpy()
        for j in range(done):
            resn.append(rows[j].copy())
        kk += done
        xk = None
        prob.check_peers()
        if callback is not None:
            xk = get_x()
            resn[-1] = _callback_resnorm(prob, callback, xk, resn[-1])
        else:
            batch = min(2 * batch, _BATCH_MAX)

    if xk is None:
        xk = get_x()
    prob.launches = ops.launches
    return _finish(prob, success, xk, kk, resn)


def _mm2(G, v):
    return np.einsum("ij...,j...->i...", G, v)


def _minres_general(prob, M, Ml, Mr, inner, tol, atol, maxiter, callback):
    alg = Alg(prob, inner)
    ops = alg.ops
    A, b, x0 = prob.A, prob.b, prob.x0
    M, Ml, Mr = prob.operator(M), prob.operator(Ml), prob.operator(Mr)
    k = prob.k

    def get_x(y):  # minres.py:95-98
        return alg.add(x0, alg.apply(Mr, y))

    def explicit_norm(z):  # minres.py:106-112
        r_ = alg.apply(Ml, alg.residual(A, b, z))
        return np.sqrt(alg.inner(r_, alg.apply(M, r_)))

    Ml_r = alg.apply(Ml, alg.residual(A, b, x0))  # minres.py:121-127
    M_Ml_r = alg.apply(M, Ml_r)
    nrm0 = np.sqrt(alg.inner(Ml_r, M_Ml_r))

    lan = _DevLanczos(alg, [Mr, A, Ml], Ml_r, M, M_Ml_r, nrm0)  # Product(Ml, A, Mr)
    W = [ops.vec(zero=True), ops.vec(zero=True)]
    y = np.array([nrm0, np.zeros_like(nrm0)])
    G = [None, None]
    yk = ops.vec(zero=True)
    xk = None
    if callback is not None:
        nrm0 = _callback_resnorm(prob, callback, x0, nrm0)
    resn = [nrm0]

    kk = 0
    success = False
    crit = np.maximum(tol * resn[0], atol)
    while True:
        if np.all(resn[-1] <= crit):
            xk = get_x(yk) if xk is None else xk
            resn[-1] = explicit_norm(xk)
            if np.all(resn[-1] <= crit):
                success = True
                break
        if kk == maxiter:
            break
        v = lan.v
        _, h, _ = lan.step()
        # implicit QR update of the tridiagonal (minres.py:195-215), host scalars
        R = np.zeros((4, k))
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
        # z = (v - R0 W0 - R1 W1)/nz(R2); W <- [W1, z]; yk += y0 z  (minres.py:219-221)
        z = v.clone()
        alg.axpy(z, R[0], W[0], sign=-1.0)
        alg.axpy(z, R[1], W[1], sign=-1.0)
        z = alg.div(z, nz(R[2]))
        W[0], W[1] = W[1], z
        alg.axpy(yk, y[0], z)
        xk = None
        y = np.array([y[1], np.zeros_like(y[1])])
        rn = np.abs(y[0])
        if callback is not None:
            xk = get_x(yk)
            rn = _callback_resnorm(prob, callback, xk, rn)
        resn.append(rn)
        kk += 1

    if xk is None:
        xk = get_x(yk)
    prob.launches = ops.launches
    return _finish(prob, success, xk, kk, resn)
