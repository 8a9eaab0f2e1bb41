"""Preconditioned GMRES on the device -- drop-in for ``krylov.gmres``
(gmres.py:41-251): full (non-restarted) GMRES with ``ortho`` in
``"mgs"``, ``"mgs<N>"`` (N sweeps of modified Gram-Schmidt), ``"householder"``
(plus the additive ``"cgs"`` / ``"cgs<N>"``: classical Gram-Schmidt as tall-skinny products);
Hessenberg QR by Givens rotations; the solution is only formed when needed.
``restart=`` is an additive extension: GMRES(m) as an outer loop
of ``gmres(maxiter=m, x0=xk)`` cycles.

Device layout: the Arnoldi basis is one (m+1, n, k) buffer; the Hessenberg
column, all previous rotations, the new rotation (LAPACK ``dlartg``
semantics), the projected right-hand side, the residual norm and the
convergence flag are updated by a one-block kernel -- no host round-trip per
Arnoldi step.  MGS step j is ``w -= h_j V_j`` fused with the next dot
``<V_{j+1}, w>`` (32 B/element); the last one carries ``<w, w>``.
"""
from __future__ import annotations

import numpy as np
import torch

from ._alg import Alg, nz
import ctypes as C

from ._lib import GmresCycleState, GmresState, check, lib
from .arnoldi import _DevHouseholder
from .device import Ops, cur_stream, ptr
from .errors import ArgumentError
from .operators import Identity, Info, Problem

INT_MAX = 2**31 - 1
_BATCH_MIN, _BATCH_MAX = 4, 64
_BASIS_CHUNK = 64  # Arnoldi vectors allocated up front; doubled on demand
USE_C_LOOP = True  # kb_gmres_cycle where it applies (tests compare both ways)
_INVARIANT_MSG = "Krylov subspace was found to be invariant in the previous iteration."


def gmres(A, b, M=None, Ml=None, Mr=None, inner=None, ortho="mgs", x0=None, tol=1e-5,
          atol=1.0e-15, maxiter=None, callback=None, restart=None, inner_product=None):
    if inner is None and inner_product is not None:
        inner = inner_product
    if restart is not None:
        return _gmres_restarted(A, b, M, Ml, Mr, inner, ortho, x0, tol, atol, maxiter, callback,
                                int(restart))
    prob = Problem(A, b, x0)
    maxiter = prob.n if maxiter is None else int(maxiter)
    with torch.cuda.device(prob.device):
        return _gmres(prob, M, Ml, Mr, inner, ortho, tol, atol, maxiter, callback)


def _num_operations(k):
    # gmres.py:240-247
    return {"A": 1 + k, "M": 2 + k, "Ml": 2 + k, "Mr": 1 + k,
            "inner": 2 + k + k * (k + 1) / 2, "axpy": 4 + 2 * k + k * (k + 1) / 2}


def _callback_resnorm(prob, callback, xk, rn):
    arr = np.array(prob.scalars_to_user(rn))  # gmres.py:223-233: may be overwritten
    callback(prob.to_user(xk), arr)
    return np.broadcast_to(np.asarray(arr[()], dtype=np.float64).reshape(-1), (prob.k,)).copy()


def _gmres(prob, M, Ml, Mr, inner, ortho, tol, atol, maxiter, callback):
    alg = Alg(prob, inner)
    ops = alg.ops
    A, b, x0 = prob.A, prob.b, prob.x0
    n, k, dev = prob.n, prob.k, prob.device
    M_is_identity = M is None or isinstance(M, Identity)
    M, Ml, Mr = prob.operator(M), prob.operator(Ml), prob.operator(Mr)
    chain = [Mr, A, Ml]  # Product(Ml, A, Mr)   (gmres.py:136)

    cgs = False
    if ortho.startswith("mgs"):  # gmres.py:147-157
        nre = 1 if len(ortho) == 3 else int(ortho[3:])
        householder = False
    elif ortho.startswith("cgs"):
        # additive extension: classical Gram-Schmidt, "cgs<N>" = N passes
        # ("cgs2" is the usual iterated CGS).  h = V^T w in one pass over w, w -= P h in one
        # pass over P: half the memory traffic of MGS and 2 reductions per pass instead of j+1.
        if inner is not None:
            raise ValueError('ortho="cgs" needs the default inner product')
        nre = 1 if len(ortho) == 3 else int(ortho[3:])
        householder = False
        cgs = True
    else:  # gmres.py:158-162
        assert ortho == "householder"
        assert inner is None
        assert M_is_identity
        if prob.comm is not None:
            # reflector j pivots on GLOBAL row j (arnoldi.py:75-96, householder.py:26-62): the
            # device kernels index rank-local rows, so a row-partitioned matrix would silently
            # pivot on the wrong entries -- refuse instead
            raise NotImplementedError(
                'gmres(ortho="householder") is single-GPU; on a row-partitioned matrix use '
                'ortho="mgs2" (or "cgs2"), which reach the same orthogonality')
        nre = 1
        householder = True
    # everything device-resident <=> default inner product (no host callable in the loop)
    fused_dots = inner is None
    csr_chain = None  # Ml A Mr as device matrices: the last product carries the first MGS dot
    if prob.A_csr is not None and all(op is None or op.csr is not None for op in (Ml, Mr)):
        csr_chain = [op for op in (None if Mr is None else Mr.csr, prob.A_csr,
                                   None if Ml is None else Ml.csr) if op is not None]
    M_csr = None if M is None else M.csr

    def residual_triple(z):  # gmres.py:105-114
        Ml_r = alg.apply(Ml, alg.residual(A, b, z))
        M_Ml_r = alg.apply(M, Ml_r)
        return M_Ml_r, Ml_r, np.sqrt(alg.inner(Ml_r, M_Ml_r))

    z0, r0, nrm0 = residual_triple(x0)
    resn = [nrm0]
    if callback is not None:
        callback(prob.to_user(x0), prob.to_user(r0))

    m = maxiter
    # Arnoldi bases V (and P = M^-1-dual basis when M is given; arnoldi.py:131-150), Hessenberg
    # factor R, rotations and projected right-hand side.  maxiter defaults to n like the
    # reference's (gmres.py:127), but the reference grows V as a Python list: allocate for
    # _BASIS_CHUNK steps and grow geometrically between batches (the loop synchronises there),
    # so that a default-argument call on a large system costs what its iteration count needs.
    cap = max(1, min(m, _BASIS_CHUNK))

    def _alloc(c):
        try:
            Vb = None if householder else torch.empty((c + 1, n, k), dtype=torch.float64, device=dev)
            Pb = Vb if (M is None or householder) else torch.empty((c + 1, n, k),
                                                                    dtype=torch.float64, device=dev)
        except torch.OutOfMemoryError as e:
            raise MemoryError(
                f"gmres: no device memory for an Arnoldi basis of {c + 1} vectors of length {n} "
                f"(x {k} columns); bound it with restart= or maxiter=") from e
        return (Vb, Pb, torch.zeros((c + 1, c, k), dtype=torch.float64, device=dev),
                torch.zeros((c, k), dtype=torch.float64, device=dev),
                torch.zeros((c, k), dtype=torch.float64, device=dev),
                torch.zeros((c + 1, k), dtype=torch.float64, device=dev),
                torch.zeros((c + 1, k), dtype=torch.float64, device=dev),
                torch.zeros((nre * (c + 1) + 2, k), dtype=torch.float64, device=dev))

    Vbuf, Pbuf, R, Gc, Gs, y, yy, dots = _alloc(cap)
    ww = ops.slots(1)[0]
    hlast = ops.slots(1)[0]
    ctl = torch.tensor([INT_MAX, 0], dtype=torch.int32, device=dev)  # stop_at, flags
    hist = torch.zeros((_BATCH_MAX, k), dtype=torch.float64, device=dev)
    nrm0_d = torch.from_numpy(np.ascontiguousarray(nrm0)).to(dev)
    y[0].copy_(nrm0_d)  # gmres.py:171
    crit = np.maximum(tol * resn[0], atol)
    crit_d = torch.from_numpy(np.ascontiguousarray(crit)).to(dev)

    hh = None
    if householder:
        hh = _DevHouseholder(alg, chain, r0, m)  # arnoldi.py:34-56
    else:
        ops.div_scale(Pbuf[0], r0, nrm0_d)  # arnoldi.py:147-150
        if M is not None:
            ops.div_scale(Vbuf[0], z0, nrm0_d)

    st = GmresState(dots=ptr(dots), ww=ptr(ww), num_reorthos=nre, maxiter=cap, R=ptr(R),
                    Gc=ptr(Gc), Gs=ptr(Gs), y=ptr(y), hlast=ptr(hlast), crit=ptr(crit_d), hist=0,
                    stop_at=ctl.data_ptr(), flags=ctl.data_ptr() + 4, have_h=1 if householder else 0)
    stop_at = ctl[0:1]

    def basis(j):
        return hh.V[j] if householder else Vbuf[j]

    def solution(kk):  # gmres.py:89-99
        if kk == 0:
            return x0.clone()
        ops.gmres_solve_y(kk, cap, R, y, yy)
        out = torch.empty_like(x0)
        if householder:
            comb = torch.zeros_like(x0)
            for j in range(kk):  # sum(c * v ...) one axpy per basis vector
                ops.axpy(comb, yy[j], hh.V[j])
        elif Mr is None:
            ops.basis_combine(kk, yy, Vbuf, x0, out)
            return out
        else:
            comb = torch.empty_like(x0)
            ops.basis_combine(kk, yy, Vbuf, torch.zeros_like(x0), comb)
        ops.add(out, x0, alg.apply(Mr, comb))
        return out

    def arnoldi_step(i):
        """One Arnoldi step with device scalars; enqueue only."""
        if householder:
            hh.step()
            st.dots = ptr(hh.h_dev)
            ops.gmres_scalar(i, st)
            return
        w = Wbuf
        # w = Ml A Mr V[i]  (+ first MGS dot <V[0], w> fused into the product)
        if csr_chain is not None and fused_dots:
            src = Vbuf[i]
            for j, op in enumerate(csr_chain[:-1]):
                ops.spmv(op, src, Tbuf[j])
                src = Tbuf[j]
            if cgs:
                ops.spmv(csr_chain[-1], src, w)
            else:
                ops.spmv(csr_chain[-1], src, w, dot=1, w=Vbuf[0], out=dots[0])
        else:
            w.copy_(alg.apply_chain(chain, Vbuf[i]))
            if not cgs:
                _dot(Vbuf[0], w, dots[0])
        idx = 0
        last_fused = False
        for sweep in range(nre if cgs else 0):
            rows = dots[sweep * (i + 1):]
            ops.multi_dot(i + 1, Vbuf, w, rows)  # h = V^T w
            if sweep == nre - 1 and M is None:
                ops.multi_axpy(i + 1, rows, Pbuf, w, dot=2, out=ww)  # w -= P h, <w, w>
                last_fused = True
            else:
                ops.multi_axpy(i + 1, rows, Pbuf, w)
        for sweep in range(0 if cgs else nre):
            for j in range(i + 1):  # arnoldi.py:157-162
                last = sweep == nre - 1 and j == i
                if last:
                    if fused_dots and M is None:
                        ops.axpy_dot(dots[idx], Pbuf[j], w, dot=2, out=ww)
                        last_fused = True
                    else:
                        ops.axpy_dot(dots[idx], Pbuf[j], w, dot=0)
                else:
                    nxt = Vbuf[j + 1] if j < i else Vbuf[0]
                    if fused_dots:
                        ops.axpy_dot(dots[idx], Pbuf[j], w, dot=1, z=nxt, out=dots[idx + 1])
                    else:
                        ops.axpy_dot(dots[idx], Pbuf[j], w, dot=0)
                        _dot(nxt, w, dots[idx + 1])
                idx += 1
        if not last_fused and M_csr is not None and fused_dots:
            Mw = MWbuf  # h[k+1]^2 = <w, M w> out of M's own product   (arnoldi.py:184-185)
            ops.spmv(M_csr, w, Mw, dot=1, w=w, out=ww)
        elif not last_fused:
            Mw = alg.apply(M, w)
            _dot(w, Mw, ww)
        else:
            Mw = w
        ops.gmres_scalar(i, st)  # gmres.py:206-221 on the device
        ops.div_scale(Pbuf[i + 1], w, hlast)  # arnoldi.py:191-193
        if M is not None:
            ops.div_scale(Vbuf[i + 1], Mw, hlast)

    def _dot(xv, yv, out):
        if fused_dots:
            ops.dot(xv, yv, out)
        else:  # user inner product: host value -> device slot
            out.copy_(torch.from_numpy(alg.inner(xv, yv)))

    Wbuf = None if householder else ops.vec(zero=False)
    Tbuf = [ops.vec(zero=False) for _ in range(0 if csr_chain is None else len(csr_chain) - 1)]
    MWbuf = ops.vec(zero=False) if M_csr is not None and not householder else None
    step_by_step = callback is not None or not fused_dots
    cycle_in_c = (USE_C_LOOP and not householder and not cgs and fused_dots and M is None and prob.comm is None
                  and csr_chain is not None and len(csr_chain) == 1
                  and hasattr(csr_chain[0], "handle"))
    batch = 1 if step_by_step else _BATCH_MIN
    kk = 0
    success = False
    xk = None
    while True:
        if np.all(resn[-1] <= crit):  # gmres.py:180-187
            xk = solution(kk) if xk is None else xk
            resn[-1] = residual_triple(xk)[2]
            if np.all(resn[-1] <= crit):
                success = True
                break
        if kk == maxiter:
            break
        c = ctl.cpu().numpy()
        if c[1] & 1:
            raise ArgumentError(_INVARIANT_MSG)  # arnoldi.py:67-70, 168-171
        nb = min(batch, maxiter - kk)
        if kk + nb > cap:  # grow the basis and the Hessenberg storage (contents kept)
            new = min(m, max(2 * cap, kk + nb))
            Vn, Pn, Rn, Gcn, Gsn, yn, yyn, dots = _alloc(new)
            if Vn is not None:
                Vn[: cap + 1].copy_(Vbuf)
                if Pn is not Vn:
                    Pn[: cap + 1].copy_(Pbuf)
            Rn[: cap + 1, :cap].copy_(R)
            Gcn[:cap].copy_(Gc)
            Gsn[:cap].copy_(Gs)
            yn[: cap + 1].copy_(y)
            Vbuf, Pbuf, R, Gc, Gs, y, yy, cap = Vn, Pn, Rn, Gcn, Gsn, yn, yyn, new
            st.dots, st.R, st.Gc, st.Gs, st.y, st.maxiter = (ptr(dots), ptr(R), ptr(Gc), ptr(Gs),
                                                            ptr(y), cap)
        stop_at.fill_(INT_MAX)
        st.hist = hist.data_ptr() - (kk + 1) * k * 8
        if cycle_in_c:
            # one GPU, MGS, no preconditioner: the whole batch is one C call (kb_gmres_cycle)
            cyc = GmresCycleState(A=csr_chain[0].handle, n=n, k=k, Vbuf=ptr(Vbuf), vstride=n * k,
                                  w=ptr(Wbuf), dots=ptr(dots), ww=ptr(ww), hlast=ptr(hlast), st=st)
            check(lib.kb_gmres_cycle(ops.ws.handle, C.byref(cyc), kk, nb, cur_stream()))
            ops.launches += sum(2 + nre * (i + 1) for i in range(kk, kk + nb))
        else:
            for i in range(kk, kk + nb):
                ops.gate(stop_at, i)
                arnoldi_step(i)
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
            xk = solution(kk)
            resn[-1] = _callback_resnorm(prob, callback, xk, resn[-1])
        if not step_by_step:
            batch = min(2 * batch, _BATCH_MAX)

    if xk is None:
        xk = solution(kk)
    prob.launches = ops.launches
    xk_user = prob.to_user(xk)
    resnorms = [prob.scalars_to_user(r) for r in resn]
    return (xk_user if success else None), Info(
        success, xk_user, kk, resnorms, num_operations=_num_operations(kk))


def _gmres_restarted(A, b, M, Ml, Mr, inner, ortho, x0, tol, atol, maxiter, callback, restart):
    """GMRES(restart): cycles of ``gmres(maxiter=restart, x0=xk)`` against one
    fixed target ``max(tol*||r0||, atol)`` taken from the first cycle
    (the user loop of SURVEY.md: the reference has no restart parameter).
    ``maxiter`` bounds the total number of Arnoldi steps."""
    total_cap = int(b.shape[0]) if maxiter is None else int(maxiter)
    x = x0
    hist = None
    target = None
    total = 0
    ok = False
    while True:
        m = restart if total_cap is None else min(restart, total_cap - total)
        if m <= 0 and target is not None:
            break
        if target is None:
            sol, info = gmres(A, b, M=M, Ml=Ml, Mr=Mr, inner=inner, ortho=ortho, x0=x, tol=tol,
                              atol=atol, maxiter=m, callback=callback)
            target = np.maximum(tol * np.asarray(info.resnorms[0]), atol)
            hist = list(info.resnorms)
        else:
            sol, info = gmres(A, b, M=M, Ml=Ml, Mr=Mr, inner=inner, ortho=ortho, x0=x, tol=0.0,
                              atol=target, maxiter=m, callback=callback)
            hist.extend(info.resnorms[1:])
        total += info.numsteps
        x = info.xk
        if info.success:
            ok = True
            break
        if info.numsteps == 0:
            break
    return (x if ok else None), Info(ok, x, total, hist, num_operations=_num_operations(total))
