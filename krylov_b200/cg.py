"""Preconditioned CG on the device -- drop-in for ``krylov.cg`` (cg.py:16-259).

Same signature, same return value ``(xk if success else None, Info)``, same
stopping rule (updated residual below ``max(tol*||r0||, atol)`` is confirmed
with an explicit residual that overwrites ``resnorms[-1]``, cg.py:156-164),
same zero-division guards, same ``callback(xk, Ml_rk)``.

Two paths:
  * fused (default inner product; A -- and M / Ml if given -- are matrices): three
    kernels per iteration -- x/p update, SpMV fused with ``<p, Ap>``, r update fused
    with ``<r, r>`` and the record/convergence step -- all scalars stay in HBM,
    iterations are enqueued in batches and gated on a device flag, so the host reads
    back once per batch, not per iteration.  A matrix preconditioner adds one sparse
    product ``z = M r`` fused with ``rho = <r, z>`` (e.g. Jacobi: 30 B/row).
  * general (custom ``inner``, duck-typed operators): the reference loop statement
    by statement, every vector statement one kernel, scalars on the host.
"""
from __future__ import annotations

import numpy as np
import torch

import ctypes as C

from . import _complex, _trace
from ._alg import Alg, nz
from ._lib import CgState, check, lib
from .device import Ops, cur_stream, ptr
from .operators import Info, Problem

INT_MAX = 2**31 - 1
_BATCH_MIN, _BATCH_MAX = 8, 256
_PERSISTENT_MAX_N = 262144  # kb_tune key 28: persistent one-launch batches up to this n


def cg(A, b, M=None, Ml=None, inner=None, x0=None, tol=1e-5, atol=1.0e-15, maxiter=None,
       return_arnoldi=False, callback=None, inner_product=None):
    if inner is None and inner_product is not None:  # alias used by BASELINE.json's wording
        inner = inner_product
    if _complex.any_complex(A, b, x0, M, Ml):  # Hermitian systems: real-equivalent embedding
        return _complex.solve_hermitian(cg, A, b, x0, {"M": M, "Ml": Ml}, inner, callback,
                                        dict(tol=tol, atol=atol, maxiter=maxiter,
                                             return_arnoldi=return_arnoldi))
    _trace.mark("cg: enter")
    prob = Problem(A, b, x0)
    _trace.mark("cg: Problem (b, x0, A on the device)")
    maxiter = prob.n if maxiter is None else int(maxiter)
    with torch.cuda.device(prob.device):
        if inner is None and prob.A_csr is not None:
            # device-resident path: A and the preconditioners (if any) are matrices
            Mop, Mlop = prob.operator(M), prob.operator(Ml)
            if all(op is None or op.csr is not None for op in (Mop, Mlop)):
                return _cg_fused(prob, tol, atol, maxiter, return_arnoldi, callback,
                                 None if Mop is None else Mop.csr,
                                 None if Mlop is None else Mlop.csr)
        return _cg_general(prob, M, Ml, inner, tol, atol, maxiter, return_arnoldi, callback)


def _num_operations(k):
    # cg.py:243-250
    return {"A": 1 + k, "M": 2 + k, "Ml": 2 + k, "Mr": 1 + k, "inner": 2 + 2 * k,
            "axpy": 2 + 2 * k}


def _finish(prob, success, xk, k, resn, arnoldi):
    xk_user = prob.to_user(xk)
    resnorms = [prob.scalars_to_user(r) for r in resn]
    return (xk_user if success else None), Info(
        success, xk_user, k, resnorms, num_operations=_num_operations(k), arnoldi=arnoldi)


class _LanczosLog:
    """cg.py:141-149, 220-232: Lanczos basis and tridiagonal from CG scalars."""

    def __init__(self, prob, ops, maxiter, z0, r0, nrm0):
        self.prob, self.ops = prob, ops
        self.V, self.P = [], []
        self.H = np.zeros([maxiter + 1, maxiter] + list(prob.user_shape[1:]), dtype=float)
        self.alpha_old = 0
        self._push(z0, r0, np.where(nrm0 > 0.0, nrm0, 1.0))

    def _push(self, z, r, d):
        cd = torch.from_numpy(np.array(np.broadcast_to(d, (self.prob.k,)),
                                       dtype=np.float64)).to(self.prob.device)
        for lst, vec in ((self.V, z), (self.P, r)):
            out = torch.empty_like(vec)
            self.ops.div_scale(out, vec, cd)
            lst.append(self.prob.to_user(out))

    def step(self, k, z, r, alpha, omega, rho_new, rho_old):
        shp = self.prob.user_shape[1:]
        nrm = np.sqrt(rho_new)
        sgn = (-1.0) ** (k + 1)
        self._push(z, r, sgn * nrm)  # (-1)^(k+1) * v / nrm  ==  v / ((-1)^(k+1) nrm)
        a = np.reshape(alpha, shp) if shp else alpha[0]
        self.H[k, k] = 1.0 / a
        if k > 0:
            self.H[k - 1, k] = self.H[k, k - 1]
            self.H[k, k] += (np.reshape(omega, shp) if shp else omega[0]) / self.alpha_old
        q = np.sqrt(rho_new / rho_old) / alpha
        self.H[k + 1, k] = np.reshape(q, shp) if shp else q[0]
        self.alpha_old = a

    def result(self, k):
        return [self.V, self.H[: k + 1, :k], self.P]


# ---------------------------------------------------------------------------
class FusedCG:
    """Device-resident state of the fused (P)CG iteration with the default inner
    product.  ``enqueue(i)`` launches iteration i (no host sync); ``run(nb)``
    enqueues a gated batch and reads back once.

    ``A`` is a CsrMatrix or a row-partitioned DistCsrMatrix; in the latter case
    every reduction is summed over ranks (inside the reducing kernel through peer
    memory, or by one small NCCL all-reduce).  ``M`` / ``Ml`` are optional
    preconditioner *matrices* (CsrMatrix, e.g. Jacobi): ``z = M r`` is a sparse
    product fused with ``rho = <r, z>`` and ``Ml`` is chained behind ``A`` -- all
    scalars stay on the device exactly as without preconditioner
    (cg.py:86-95, 109, 180, 207-212)."""

    def __init__(self, A, b, x0, tol, atol, M=None, Ml=None):
        n, k = b.shape
        self.A, self.b, self.x0 = A, b, x0
        self.M, self.Ml = M, Ml
        self.n, self.k, self.dev = n, k, b.device
        self.comm = getattr(A, "comm", None)
        self.ops = ops = Ops(n, k, self.dev, comm=self.comm)
        # Row-partitioned two-launch path (csrc/kb_march.cuh on the ghost-extended row space):
        # r lives in the matrix's IPC-exported area so that the neighbours' r update can store
        # their boundary planes into its ghost planes
        self.gplan = None
        if (self.comm is not None and k == 1 and M is None and Ml is None and ops.fused_allreduce
                and hasattr(A, "fused_cg_plan")):
            self.gplan = A.fused_cg_plan()
        if self.gplan is not None:
            P = self.gplan["P"]
            self.r_ext = self.gplan["r_ext"]
            if getattr(FusedCG, "_debug_local_r", False):  # measurement only (tools/dist_slab.py)
                self.r_ext = torch.zeros_like(self.r_ext)
            self.r_ext[:P].zero_()
            self.r_ext[P + n:].zero_()
            self.r = self.r_ext[P:P + n].view(n, 1)
        else:
            self.r = ops.vec(zero=False)
        self.Ap = ops.vec(zero=False)
        self.yk = ops.vec(zero=True)
        self.z = ops.vec(zero=False) if M is not None else self.r   # M_Ml_rk (== Ml_rk if M = I)
        self.t = ops.vec(zero=False) if Ml is not None else None      # A p before Ml is applied
        self.zs = ops.vec(zero=False) if M is not None else None      # scratch z of explicit checks
        # State slots (written by gated kernels only): rho ping-pong (rho_i in sl[i % 2]),
        # sl[2] = alpha of the last iteration.  Landing slots of reductions (these may see
        # un-gated NCCL all-reduces after on-device convergence, so nothing persistent
        # lives there): sl[3] = <p, Ap>, sl[4] = <r, r>, sl[5] = scratch.
        self.sl = ops.slots(7)  # sl[6] stays zero (omega of the first partitioned iteration)
        self.stop_at = torch.full((1,), INT_MAX, dtype=torch.int32, device=self.dev)
        self.hist = torch.zeros((_BATCH_MAX, k), dtype=torch.float64, device=self.dev)
        self.spmv_events = None  # bench hook: list of (start, end) CUDA events around A @ p
        # initial residual r0 = b - A x0 fused with <r0, r0>  (cg.py:116)
        ops.gate(None, 0)
        self.rho0 = self._residual_norm2(x0, self.r, self.z, self.sl[0])
        self.nrm0 = np.sqrt(self.rho0)
        self.crit = np.maximum(tol * self.nrm0, atol)  # cg.py:154
        self.crit_d = torch.from_numpy(np.ascontiguousarray(self.crit)).to(self.dev)
        # search direction: pbuf[pcur].  The fused marching kernels (single GPU, k = 1, 3-D
        # constant-coefficient stencil: kb_cg_run with a second buffer) read p with its halo
        # and write the new p to the other buffer; every other path updates pbuf[pcur] in place.
        self.pcur = 0
        self.kk = 0
        self._cstate = None
        if self.gplan is not None:
            # both search-direction buffers ghost-extended and zero: iteration 0 forms p = r + 0 p
            # on the own AND the ghost planes (the neighbours' r0 planes are exchanged once here)
            P, n_ext = self.gplan["P"], self.gplan["n_ext"]
            A.exchange_ghost_planes(self.r_ext, P)
            self.p_ext = [torch.zeros(n_ext, dtype=torch.float64, device=self.dev) for _ in (0, 1)]
            self.pbuf = [pe[P:P + n].view(n, 1) for pe in self.p_ext]
            self._cstate = CgState(A=A.A_loc.handle, n=n, k=k, x=ptr(self.yk), r=ptr(self.r_ext),
                                   p=ptr(self.p_ext[0]), Ap=ptr(self.Ap), slots=ptr(self.sl),
                                   crit=ptr(self.crit_d), hist=ptr(self.hist),
                                   stop_at=ptr(self.stop_at), p2=ptr(self.p_ext[1]), pcur=0,
                                   masks_ext=ptr(self.gplan["masks_ext"]), n_ext=n_ext, own_lo=P,
                                   r_push_lo=self.gplan["push_lo"], r_push_hi=self.gplan["push_hi"])
            fz = C.c_int(0)
            check(lib.kb_cg_is_fused(C.byref(self._cstate), C.byref(fz)))
            if not fz.value:
                raise RuntimeError("row-partitioned fused CG: the library rejected the plan")
        else:
            self.pbuf = [self.z.clone(), None]
        if self.comm is None and hasattr(A, "handle") and M is None and Ml is None:
            # a second search-direction buffer: the fused marching kernels (3-D stencils) and
            # the persistent small-problem kernel (csrc/kb_small.cu) ping-pong p
            if k == 1 and (A.info()["schedule"] == "stencil" or n <= _PERSISTENT_MAX_N):
                self.pbuf[1] = ops.vec(zero=False)
            self._cstate = CgState(A=A.handle, n=n, k=k, x=ptr(self.yk), r=ptr(self.r),
                                   p=ptr(self.pbuf[0]), Ap=ptr(self.Ap), slots=ptr(self.sl),
                                   crit=ptr(self.crit_d), hist=ptr(self.hist),
                                   stop_at=ptr(self.stop_at),
                                   p2=ptr(self.pbuf[1]) if self.pbuf[1] is not None else None,
                                   pcur=0)
        # x += alpha p of the last enqueued iteration is deferred into the next p
        # update (which streams p anyway: 64 instead of 72 B/element for the two
        # vector kernels of a step); current_x() flushes it
        self.x_pending = False
        self.fused_march = False
        self.persistent = False

    @property
    def p(self):
        return self.pbuf[self.pcur]

    def _residual_norm2(self, x, out_r, out_z, slot):
        """out_r = Ml (b - A x); out_z = M out_r; returns host <out_r, out_z> (k,)
        (cg.py:72-95).  Without preconditioners one fused kernel."""
        ops = self.ops
        if self.Ml is None:
            ops.spmv(self.A, x, out_r, mode=2, z=self.b, dot=0 if self.M is not None else 2,
                     out=slot)
        else:
            ops.spmv(self.A, x, self.t, mode=2, z=self.b)
            ops.spmv(self.Ml, self.t, out_r, dot=0 if self.M is not None else 2, out=slot)
        if self.M is not None:
            ops.spmv(self.M, out_r, out_z, dot=1, w=out_r, out=slot)
        return slot.cpu().numpy().copy()

    def explicit_resnorm(self, xk):
        return np.sqrt(self._residual_norm2(xk, self.Ap, self.zs, self.sl[5]))

    def flush_x(self):
        if self.x_pending:
            self.ops.gate(None, 0)
            self.ops.cg_flush_x(self.sl[2], self.p, self.yk)  # alpha of the last iteration run
            self.x_pending = False

    def current_x(self):
        self.flush_x()
        xk = torch.empty_like(self.yk)
        self.ops.add(xk, self.x0, self.yk)
        return xk

    def enqueue(self, i, hist_ptr):
        ops, sl = self.ops, self.sl
        cur, nxt = sl[i % 2], sl[(i + 1) % 2]
        ops.gate(self.stop_at, i)  # iteration i is a no-op once a step <= i converged
        if i > 0:
            # [x += alpha_{i-1} p;]  omega = rho_i / rho_{i-1};  p = z + omega p   (z = M r)
            if self.x_pending:
                ops.cg_update_p(cur, nxt, self.z, self.p, x=self.yk, alpha=sl[2])
            else:
                ops.cg_update_p(cur, nxt, self.z, self.p)
        if self.spmv_events is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.Ml is None:
            ops.spmv(self.A, self.p, self.Ap, dot=1, w=self.p, out=sl[3])  # Ap = A p, <p, Ap>
        else:  # Product(Ml, A) @ p   (cg.py:109,180)
            ops.spmv(self.A, self.p, self.t)
            ops.spmv(self.Ml, self.t, self.Ap, dot=1, w=self.p, out=sl[3])
        if self.spmv_events is not None:
            e1.record()
            self.spmv_events.append((e0, e1))
        # alpha -> sl[2]; r -= alpha Ap; <r, r> -> sl[4]; record step i+1, rho_{i+1} -> nxt
        if self.M is not None:
            # r -= alpha Ap;  z = M r fused with rho_{i+1} = <r, z>  (cg.py:200-212);  record
            ops.cg_update_xr(cur, sl[3], None, None, self.Ap, None, self.r, sl[5], alpha_out=sl[2])
            ops.spmv(self.M, self.r, self.z, dot=1, w=self.r, out=sl[4])
            ops.cg_record(i + 1, sl[4], self.crit_d, hist_ptr, self.stop_at, rho_keep=nxt)
        elif self.comm is None or ops.fused_allreduce:
            ops.cg_update_r_record(cur, sl[3], self.Ap, self.r, sl[4], sl[2], i + 1, self.crit_d,
                                   hist_ptr, self.stop_at, nxt)
        else:  # NCCL all-reduce lands after the kernel: record in a launch of its own
            ops.cg_update_xr(cur, sl[3], None, None, self.Ap, None, self.r, sl[4], alpha_out=sl[2])
            ops.cg_record(i + 1, sl[4], self.crit_d, hist_ptr, self.stop_at, rho_keep=nxt)
        self.x_pending = True

    def run(self, nb):
        """Enqueue iterations kk .. kk+nb-1, then one host read.  Returns the
        residual norms of the steps that actually ran (the rest were gated)."""
        kk, k = self.kk, self.k
        self.stop_at.fill_(INT_MAX)
        via_c = self.spmv_events is None and self._cstate is not None
        if via_c:
            # the whole batch is enqueued by one C call (kb_cg_run); row-partitioned: the fused
            # two-launch path with peer-memory pushes and all-reduces inside the kernels
            self._cstate.pcur = self.pcur
            fz, pz = C.c_int(0), C.c_int(0)
            check(lib.kb_cg_is_fused(C.byref(self._cstate), C.byref(fz)))
            check(lib.kb_cg_is_persistent(self.ops.ws.handle, C.byref(self._cstate), C.byref(pz)))
            self.persistent = bool(pz.value)
            # both paths move p to the other buffer with every executed iteration i > 0
            self.fused_march = bool(fz.value) or self.persistent
            check(lib.kb_cg_run(self.ops.ws.handle, C.byref(self._cstate), kk, nb,
                                1 if self.x_pending else 0, cur_stream()))
            self.ops.launches += self._launches_of(kk, nb)
            self.x_pending = True
        else:
            hist_ptr = self.hist.data_ptr() - (kk + 1) * k * 8  # history row kk+1 == hist[0]
            for i in range(kk, kk + nb):
                self.enqueue(i, hist_ptr)
            self.ops.gate(None, 0)
        s = int(self.stop_at.item())  # one host read per batch
        done = min(s, kk + nb) - kk
        if via_c and self.fused_march:
            self.pcur = self._pcur_after(kk, done)
        rows = self.hist[:done].cpu().numpy()
        self.kk += done
        self._check_peers()
        return [rows[j].copy() for j in range(done)]

    def _launches_of(self, kk, nb):
        if self.persistent:
            return 1  # the whole batch is one persistent launch
        if self.gplan is not None:
            return 2 * nb
        return (2 if self.fused_march else 3) * nb - (
            (0 if self.fused_march else 1) if kk == 0 else 0)

    def _pcur_after(self, kk, done):
        """Every executed iteration moves p to the other buffer, except iteration 0 on one
        GPU (p = r0 is used in place there)."""
        if self.gplan is not None:
            return (self.pcur + done) % 2
        return (self.pcur + done - (1 if kk == 0 and done > 0 else 0)) % 2

    def _check_peers(self):
        """A peer that never arrived (3 s device-side budget) poisons the results with NaN and
        raises the sticky error word: turn it into an exception at every batch boundary."""
        if self.comm is not None:
            chk = getattr(self.A, "check_p2p", None) or getattr(self.comm, "check_p2p", None)
            if chk is not None:
                chk()


    def run_timed(self, nb):
        """Measurement hook (bench.py): run(nb) through kb_cg_run_timed -- the same launches with
        CUDA events around each one on the launching stream.  Returns (mean ms of the step's
        phases, total ms, fused?).  Phases: fused path {p/x update + A p + <p,Ap>, r update +
        <r,r>}; three-kernel path {p/x update, A p + <p,Ap>, r update + <r,r>}."""
        assert self._cstate is not None, "C batch path only"
        kk = self.kk
        self.stop_at.fill_(INT_MAX)
        self._cstate.pcur = self.pcur
        fz = C.c_int(0)
        check(lib.kb_cg_is_fused(C.byref(self._cstate), C.byref(fz)))
        self.fused_march = bool(fz.value)
        self.persistent = False  # the timed twin always launches per phase
        ms = (C.c_float * 3)()
        tot = C.c_float(0)
        check(lib.kb_cg_run_timed(self.ops.ws.handle, C.byref(self._cstate), kk, nb,
                                  1 if self.x_pending else 0, cur_stream(), ms, C.byref(tot)))
        self.ops.launches += self._launches_of(kk, nb)
        self.x_pending = True
        s = int(self.stop_at.item())
        done = min(s, kk + nb) - kk
        if self.fused_march:
            self.pcur = self._pcur_after(kk, done)
        self.kk += done
        self._check_peers()
        n_ph = 2 if self.fused_march else 3
        return [float(ms[i]) for i in range(n_ph)], float(tot.value), self.fused_march


def _cg_fused(prob, tol, atol, maxiter, return_arnoldi, callback, M=None, Ml=None):
    st = FusedCG(prob.A_csr, prob.b, prob.x0, tol, atol, M=M, Ml=Ml)
    _trace.mark("cg: FusedCG set-up (vectors, r0 = b - A x0)")
    ops, crit = st.ops, st.crit
    if callback is not None:
        callback(prob.to_user(prob.x0), prob.to_user(st.r))
    resn = [st.nrm0]
    log = _LanczosLog(prob, ops, maxiter, st.z, st.r, st.nrm0) if return_arnoldi else None

    step_by_step = callback is not None or return_arnoldi
    batch = 1 if step_by_step else _BATCH_MIN
    success = False
    xk = None
    while True:
        if np.all(resn[-1] <= crit):
            # "oh really?" -- explicit residual of xk = x0 + yk  (cg.py:156-164)
            xk = st.current_x() if xk is None else xk
            resn[-1] = st.explicit_resnorm(xk)
            if np.all(resn[-1] <= crit):
                success = True
                break
        if st.kk == maxiter:
            break
        kk = st.kk
        resn.extend(st.run(min(batch, maxiter - kk)))
        xk = None
        if log is not None:  # batch == 1 here
            sv = st.sl.cpu().numpy()
            rho_i, rho_n, alpha = sv[kk % 2], sv[(kk + 1) % 2], sv[2]
            omega = rho_i / nz(log.rho_prev) if kk > 0 else None
            log.step(kk, st.z, st.r, alpha, omega, rho_n, rho_i)
            log.rho_prev = rho_i
        if callback is not None:
            xk = st.current_x()
            callback(prob.to_user(xk), prob.to_user(st.r))
        if not step_by_step:
            # launched paths ramp up (iterations enqueued behind the converged one are wasted
            # launches); the persistent kernel leaves its loop on the device: full batches
            batch = _BATCH_MAX if st.persistent else min(2 * batch, _BATCH_MAX)

    if xk is None:
        xk = st.current_x()
    prob.launches = ops.launches
    _trace.mark("cg: iterations + explicit residual check")
    out = _finish(prob, success, xk, st.kk, resn,
                  log.result(st.kk) if log is not None else None)
    _trace.mark("cg: result to the caller's array kind (download)")
    return out


# ---------------------------------------------------------------------------
def _cg_general(prob, M, Ml, inner, tol, atol, maxiter, return_arnoldi, callback):
    alg = Alg(prob, inner)
    ops = alg.ops
    A, b, x0 = prob.A, prob.b, prob.x0
    M = prob.operator(M)
    Ml = prob.operator(Ml)

    def residual_triple(z):  # cg.py:72-95
        r_ = alg.apply(Ml, alg.residual(A, b, z))
        z_ = alg.apply(M, r_)
        return z_, r_, alg.inner(r_, z_)

    z0, r0, rho = residual_triple(x0)
    nrm0 = np.sqrt(rho)
    if callback is not None:
        callback(prob.to_user(x0), prob.to_user(r0))
    resn = [nrm0]
    yk = ops.vec(zero=True)
    xk = None
    rho_prev = None
    r = r0.clone()
    z = z0.clone() if z0 is not r0 else r
    p = z.clone()
    log = _LanczosLog(prob, ops, maxiter, z0, r0, nrm0) if return_arnoldi else None

    kk = 0
    success = False
    crit = np.maximum(tol * nrm0, atol)
    omega = None
    while True:
        if np.all(resn[-1] <= crit):
            xk = alg.add(x0, yk) if xk is None else xk
            _, _, n2 = residual_triple(xk)
            resn[-1] = np.sqrt(n2)
            if np.all(resn[-1] <= crit):
                success = True
                break
        if kk == maxiter:
            break
        if kk > 0:
            omega = rho / nz(rho_prev)
            alg.xpby(p, z, omega)  # p = z + omega p   (cg.py:178)
        Ap = alg.apply_chain([A, Ml], p)  # Product(Ml, A) @ p  (cg.py:109,180)
        pAp = alg.inner(p, Ap)
        alpha = rho / nz(pAp)
        alg.axpy(yk, alpha, p)  # cg.py:196
        xk = None
        alg.axpy(r, alpha, Ap, sign=-1.0)  # cg.py:200
        if callback is not None:
            xk = alg.add(x0, yk)
            callback(prob.to_user(xk), prob.to_user(r))
        z = alg.apply(M, r)  # cg.py:207
        rho_new = alg.inner(r, z)
        rho_prev, rho = rho, rho_new
        resn.append(np.sqrt(rho_new))
        if log is not None:
            log.step(kk, z, r, alpha, omega, rho, rho_prev)
        kk += 1

    if xk is None:
        xk = alg.add(x0, yk)
    prob.launches = ops.launches
    return _finish(prob, success, xk, kk, resn, log.result(kk) if log is not None else None)
