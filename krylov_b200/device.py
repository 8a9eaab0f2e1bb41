"""Device plumbing: PyTorch owns memory and streams, the C ABI does the work.

`Ops` is a thin, allocation-free façade over the C entry points for vectors of
a fixed logical shape (n, k); it exists so that the solver loops read like the
reference's loops while every statement is one kernel launch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import GmresState, KrylovB200Error, MinresState, check, lib


def require_cuda():
    if not torch.cuda.is_available():
        raise KrylovB200Error(
            "krylov_b200 needs an sm_100 CUDA device; there is no CPU fallback")


def cur_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return None if t is None else t.data_ptr()


class _DeviceMemory:
    """Raw device memory owned by the library (or a peer process) as a CUDA-array-interface
    object, so that torch can view it without a copy."""

    def __init__(self, address, count, typestr, owner=None):
        self.owner = owner  # keeps the allocation alive as long as the tensor lives
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(address), False), "version": 2}


def view_device_memory(address, count, dtype, device, owner=None) -> torch.Tensor:
    """1-D tensor over `count` elements at a raw device address (no copy, no ownership)."""
    typestr = {torch.float64: "<f8", torch.int16: "<i2", torch.int32: "<i4"}[dtype]
    with torch.cuda.device(device):
        return torch.as_tensor(_DeviceMemory(address, count, typestr, owner), device=device)


def as_device_matrix(v, device=None) -> torch.Tensor:
    """Any real array-like -> contiguous fp64 CUDA tensor."""
    if isinstance(v, torch.Tensor):
        if v.is_complex():
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        return v.to(device=device or "cuda", dtype=torch.float64).contiguous()
    a = np.asarray(v)
    if np.iscomplexobj(a):
        raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
    a = np.ascontiguousarray(a, dtype=np.float64)
    return torch.from_numpy(a).to(device or "cuda")


class Workspace:
    """Owns one kb_ws handle (block-partials buffer + arrival ticket)."""

    def __init__(self, max_k: int = 1):
        require_cuda()
        self.max_k = int(max_k)
        h = C.c_void_p()
        check(lib.kb_ws_create(C.byref(h), self.max_k))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.kb_ws_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# Workspaces are pooled per (device, k): creating one is a cudaMalloc and destroying
# one a cudaFree (which synchronises the device and can stall for tens of ms), so a
# solve borrows a workspace and hands it back instead -- after the first solve of a
# given block width the library performs no device allocation at all.
_WS_POOL = {}


def _borrow_workspace(device, k):
    free = _WS_POOL.setdefault((device.index, int(k)), [])
    # A pooled workspace can be dead: when an Ops object dies as part of a reference cycle the
    # garbage collector runs the finalizers in arbitrary order, so Workspace.__del__ may have
    # destroyed the handle before or after Ops.__del__ put the object back here.
    while free:
        ws = free.pop()
        if getattr(ws, "handle", None):
            return ws
    # room for the multi-sum reductions of classical Gram-Schmidt (up to 16 sums per launch)
    return Workspace(max(int(k), min(256, 16 * int(k))))


def _return_workspace(device, k, ws):
    try:
        if not getattr(ws, "handle", None):
            return
        lib.kb_ws_set_gate(ws.handle, None, 0)
        lib.kb_ws_set_comm(ws.handle, None, 0)
        free = _WS_POOL.setdefault((device.index, int(k)), [])
        if len(free) < 16:
            free.append(ws)
    except Exception:
        pass


class Ops:
    """Kernel launches for (n, k) fp64 vectors on the current CUDA stream."""

    def __init__(self, n: int, k: int, device=None, comm=None):
        """`comm`: optional communicator of a row-partitioned problem; every
        reduction slot is then summed over ranks right after the kernel that
        produced the local part (one small all-reduce per inner product)."""
        require_cuda()
        self.comm = comm
        if k < 1 or k > 256:
            raise ValueError("blocked right-hand sides: 1 <= k <= 256 columns supported")
        self.n, self.k = int(n), int(k)
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self.ws = _borrow_workspace(self.device, k)
        self.launches = 0  # kernels enqueued through this object (bench `gpu_launches`)
        # peer-memory communicator: reductions are all-reduced inside the reducing kernel
        self.fused_allreduce = comm is not None and getattr(comm, "p2p_handle", None) is not None
        if self.fused_allreduce:
            self.set_collective(True)

    def __del__(self):
        ws = getattr(self, "ws", None)
        if ws is not None:
            self.ws = None
            _return_workspace(self.device, self.k, ws)

    def set_collective(self, on: bool):
        check(lib.kb_ws_set_comm(self.ws.handle, self.comm.p2p_handle, 1 if on else 0))

    # -- allocation helpers (set-up only, never inside an iteration)
    def vec(self, zero=True):
        f = torch.zeros if zero else torch.empty
        return f((self.n, self.k), dtype=torch.float64, device=self.device)

    def slots(self, m=1):
        return torch.zeros((m, self.k), dtype=torch.float64, device=self.device)

    # -- gating
    def gate(self, stop_at, tag):
        check(lib.kb_ws_set_gate(self.ws.handle, ptr(stop_at), int(tag)))

    # -- sparse products
    def spmv(self, A, x, y, mode=0, z=None, coef=None, dot=0, w=None, out=None):
        A._apply(self, x, y, mode, z, coef, dot, w, out)

    # -- reductions
    def reduce_over_ranks(self, slot):
        if self.comm is not None and not self.fused_allreduce:
            self.comm.allreduce(slot)

    def dot(self, x, y, out, n=None):
        self.launches += 1
        check(lib.kb_dot(self.ws.handle, self.n if n is None else n, self.k, ptr(x), ptr(y),
                         ptr(out), cur_stream()))
        self.reduce_over_ranks(out)

    # -- CG
    def cg_update_xr(self, rho, pAp, pAp2, p, Ap, x, r, rr_out, alpha_out=None):
        self.launches += 1
        check(lib.kb_cg_update_xr(self.ws.handle, self.n, self.k, ptr(rho), ptr(pAp), ptr(pAp2),
                                  ptr(p), ptr(Ap), ptr(x), ptr(r), ptr(rr_out), ptr(alpha_out),
                                  cur_stream()))
        self.reduce_over_ranks(rr_out)

    def cg_update_r_record(self, rho, pAp, Ap, r, rr_out, alpha_out, step, crit, hist_ptr, stop_at,
                           rho_keep):
        """r -= alpha Ap; <r,r> (+ fused all-reduce); record/convergence in the same launch.
        Only where the reducing kernel holds the final sum (no NCCL all-reduce after it)."""
        assert self.comm is None or self.fused_allreduce
        self.launches += 1
        check(lib.kb_cg_update_xr_record(self.ws.handle, self.n, self.k, ptr(rho), ptr(pAp), None,
                                         ptr(Ap), None, ptr(r), ptr(rr_out), ptr(alpha_out),
                                         int(step), ptr(crit), hist_ptr, ptr(stop_at), ptr(rho_keep),
                                         cur_stream()))

    def cg_update_p(self, rho_new, rho_old, r, p, x=None, alpha=None):
        """[x += alpha p;]  p = r + (rho_new / nz(rho_old)) p"""
        self.launches += 1
        what = 1 | (4 if x is not None else 0)
        check(lib.kb_cg_update_p(self.ws.handle, self.n, self.k, 0, ptr(rho_new), ptr(rho_old),
                                 ptr(alpha), None, None, None, None, ptr(r), ptr(p), ptr(x), what,
                                 cur_stream()))

    def cg_flush_x(self, alpha, p, x):
        """x += alpha p  (deferred update of the last iteration)"""
        self.launches += 1
        check(lib.kb_cg_update_p(self.ws.handle, self.n, self.k, 0, None, None, ptr(alpha), None,
                                 None, None, None, None, ptr(p), ptr(x), 4, cur_stream()))

    def cg_record(self, step, rho_new, crit, hist_ptr, stop_at, rho_keep=None):
        """hist[step] = sqrt(rho_new); rho_keep = rho_new; all columns <= crit -> stop_at = step.
        `hist_ptr` is a raw device address (row 0 of the history)."""
        self.launches += 1
        check(lib.kb_cg_update_p(self.ws.handle, self.n, self.k, int(step), ptr(rho_new), None, None,
                                 ptr(crit), hist_ptr, ptr(stop_at), ptr(rho_keep), None, None, None,
                                 2, cur_stream()))

    # -- generic vector kernels
    def axpy(self, y, coef, x, sign=1.0):
        self.launches += 1
        check(lib.kb_axpy(self.ws.handle, self.n, self.k, float(sign), ptr(coef), ptr(x), ptr(y),
                          cur_stream()))

    def lincomb(self, out, ca, x, cb=None, y=None):
        """out = ca * x + cb * y (products rounded, then the sum); ca None: x itself; cb None:
        no y term.  out may alias x or y."""
        self.launches += 1
        check(lib.kb_lincomb(self.ws.handle, self.n, self.k, ptr(ca), ptr(x), ptr(cb), ptr(y),
                             ptr(out), cur_stream()))

    def scalar_op(self, op, a, b, sa, sb, out):
        """out = A op B on (k,) device slots (kb_scalar_op); a / b None: the immediates sa / sb."""
        self.launches += 1
        check(lib.kb_scalar_op(self.ws.handle, self.k, int(op), ptr(a), ptr(b), float(sa), float(sb),
                               ptr(out), cur_stream()))

    def record(self, step, val, crit, hist_ptr, stop_at):
        """hist[step] = val; every column <= crit -> stop_at = step (kb_record); hist_ptr: raw
        device address of history row 0."""
        self.launches += 1
        check(lib.kb_record(self.ws.handle, self.k, int(step), ptr(val), ptr(crit), hist_ptr,
                            ptr(stop_at), cur_stream()))

    def xpby(self, y, x, coef):
        self.launches += 1
        check(lib.kb_xpby(self.ws.handle, self.n, self.k, ptr(x), ptr(coef), ptr(y), cur_stream()))

    def div_scale(self, out, x, coef):
        self.launches += 1
        check(lib.kb_div_scale(self.ws.handle, self.n, self.k, ptr(x), ptr(coef), ptr(out),
                               cur_stream()))

    def add(self, out, x, y):
        self.launches += 1
        check(lib.kb_add(self.ws.handle, self.n, self.k, ptr(x), ptr(y), ptr(out), cur_stream()))

    def axpy_dot(self, coef, u, w, dot=0, z=None, out=None, scale=None):
        self.launches += 1
        check(lib.kb_axpy_dot(self.ws.handle, self.n, self.k, ptr(coef), ptr(scale), ptr(u), ptr(w),
                              int(dot), ptr(z), ptr(out), cur_stream()))
        if dot:
            self.reduce_over_ranks(out)

    # -- MINRES
    def minres_scalar(self, it, state: MinresState):
        self.launches += 1
        check(lib.kb_minres_scalar(self.ws.handle, self.k, int(it), C.byref(state), cur_stream()))

    def minres_update(self, coefs, v, W0, W1, Av, yk, vnext, MAv=None, pnext=None):
        self.launches += 1
        check(lib.kb_minres_update(self.ws.handle, self.n, self.k, ptr(coefs), ptr(v), ptr(W0),
                                   ptr(W1), ptr(Av), ptr(yk), ptr(vnext), ptr(MAv), ptr(pnext),
                                   cur_stream()))

    # -- GMRES
    def gmres_scalar(self, it, state: GmresState):
        self.launches += 1
        check(lib.kb_gmres_scalar(self.ws.handle, self.k, int(it), C.byref(state), cur_stream()))

    def gmres_solve_y(self, m, maxiter, R, y, yy):
        self.launches += 1
        check(lib.kb_gmres_solve_y(self.ws.handle, self.k, int(m), int(maxiter), ptr(R), ptr(y),
                                   ptr(yy), cur_stream()))

    def multi_dot(self, cnt, V, w, out):
        """out[j] = <V[j], w>, j < cnt: V a (>= cnt, n, k) buffer, out (>= cnt, k) rows."""
        self.launches += -(-int(cnt) // 8)
        check(lib.kb_multi_dot(self.ws.handle, self.n, self.k, int(cnt), ptr(V), self.n * self.k,
                               ptr(w), ptr(out), cur_stream()))
        if cnt:
            self.reduce_over_ranks(out[:cnt])

    def multi_axpy(self, m, h, P, w, dot=0, out=None):
        """w -= sum_j h[j] P[j], j < m; dot=2 also out = <w, w>."""
        self.launches += 1
        check(lib.kb_multi_axpy(self.ws.handle, self.n, self.k, int(m), ptr(h), ptr(P),
                                self.n * self.k, ptr(w), int(dot), ptr(out), cur_stream()))
        if dot:
            self.reduce_over_ranks(out)

    def basis_combine(self, m, yy, Vbuf, x0, out):
        self.launches += 1
        check(lib.kb_basis_combine(self.ws.handle, self.n, self.k, int(m), ptr(yy), ptr(Vbuf),
                                   Vbuf.stride(0) if Vbuf is not None and Vbuf.dim() == 3
                                   else self.n * self.k,
                                   ptr(x0), ptr(out), cur_stream()))

    # -- Householder
    def house_make(self, off, x, v, params, scratch, lapack_sign=False):
        if self.comm is not None:
            raise NotImplementedError("Householder orthogonalisation is single-GPU in this round")
        self.launches += 3
        check(lib.kb_house_make2(self.ws.handle, self.n, int(off), ptr(x), ptr(v), ptr(params),
                                 ptr(scratch), 1 if lapack_sign else 0, cur_stream()))

    def house_hlast(self, w, off, v, params, tau, h_out):
        self.launches += 1
        check(lib.kb_house_hlast(self.ws.handle, ptr(w), int(off), ptr(v), ptr(params), ptr(tau),
                                 ptr(h_out), cur_stream()))

    def poke(self, op, x, idx, s=None, val=0.0, dst=None):
        self.launches += 1
        check(lib.kb_poke(self.ws.handle, int(op), ptr(x), int(idx), ptr(s), float(val), ptr(dst),
                          cur_stream()))


class BlockOps:
    """Tall-skinny block products on the FP64 tensor cores (``kb_block_gram`` /
    ``kb_block_apply``) for row-major 2-D CUDA tensors whose rows may be strided (column
    sub-blocks are valid operands).  More than 16 columns: 16-column panels."""

    _KEY = 256  # workspace width: one 16 x 16 block of sums per CTA

    def __init__(self, device=None):
        require_cuda()
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self.ws = _borrow_workspace(self.device, self._KEY)
        self.launches = 0

    def __del__(self):
        ws = getattr(self, "ws", None)
        if ws is not None:
            self.ws = None
            _return_workspace(self.device, self._KEY, ws)

    @staticmethod
    def _ld(t):
        if t.dim() != 2 or t.dtype != torch.float64 or not t.is_cuda:
            raise ValueError("block operands are 2-D fp64 CUDA tensors")
        if t.shape[1] > 1 and t.stride(1) != 1:
            raise ValueError("block operands need unit stride along the columns")
        return max(int(t.stride(0)), int(t.shape[1])) if t.shape[0] > 1 else int(t.shape[1])

    def gram(self, X, Y, out=None, acc=None, sqrt_abs=False):
        """out[i, j] = sum_r X[r, i] Y[r, j] (a (k, l) device tensor); ``acc`` (k, l): += too."""
        n, k = X.shape
        l = Y.shape[1]
        if Y.shape[0] != n:
            raise ValueError(f"row counts differ: {tuple(X.shape)} vs {tuple(Y.shape)}")
        G = out if out is not None else torch.zeros((k, l), dtype=torch.float64, device=self.device)
        if k == 0 or l == 0:
            return G
        ldx, ldy = self._ld(X), self._ld(Y)
        for i0 in range(0, k, 16):
            for j0 in range(0, l, 16):
                kk, ll = min(16, k - i0), min(16, l - j0)
                Xp, Yp, Gp = X[:, i0:], Y[:, j0:], G[i0:, j0:]
                Ap = None if acc is None else acc[i0:, j0:]
                self.launches += 1
                check(lib.kb_block_gram(self.ws.handle, n, kk, ll, ptr(Xp), ldx, ptr(Yp), ldy,
                                        ptr(Gp), G.stride(0), ptr(Ap),
                                        0 if acc is None else acc.stride(0),
                                        1 if sqrt_abs else 0, cur_stream()))
        return G

    def apply(self, X, C, Y=None, sign=0, out=None):
        """sign 0: X C;  sign -1: Y - X C;  sign +1: Y + X C   (X (n, k), C (k, l) on the device)."""
        n, k = X.shape
        l = C.shape[1]
        if C.shape[0] != k:
            raise ValueError(f"inner dimensions differ: {tuple(X.shape)} @ {tuple(C.shape)}")
        if sign != 0 and (Y is None or tuple(Y.shape) != (n, l)):
            raise ValueError("Y must have shape (n, l)")
        Z = out if out is not None else torch.empty((n, l), dtype=torch.float64, device=self.device)
        if n == 0 or l == 0:
            return Z
        if k == 0:
            if sign == 0:
                Z.zero_()
            elif Z.data_ptr() != Y.data_ptr():
                Z.copy_(Y)
            return Z
        if out is not None and out.data_ptr() == X.data_ptr() and (k > 16 or l > 16):
            raise ValueError("in-place X <- X C needs k, l <= 16")
        C = C.contiguous()
        ldx, ldz = self._ld(X), self._ld(Z)
        ldy = self._ld(Y) if Y is not None else 0
        for j0 in range(0, l, 16):
            ll = min(16, l - j0)
            for i0 in range(0, k, 16):
                kk = min(16, k - i0)
                if i0 == 0:
                    mode = {0: 0, -1: 1, 1: 2}[sign]
                    Ysrc = None if Y is None else Y[:, j0:]
                    ldsrc = ldy
                else:  # further panels of the contraction accumulate onto the result
                    mode = 1 if sign == -1 else 2
                    Ysrc, ldsrc = Z[:, j0:], ldz
                self.launches += 1
                check(lib.kb_block_apply(self.ws.handle, n, kk, ll, ptr(X[:, i0:]), ldx,
                                         ptr(C[i0:, j0:]), C.stride(0), ptr(Ysrc), ldsrc,
                                         ptr(Z[:, j0:]), ldz, mode, cur_stream()))
        return Z
