"""Device vector algebra with host-side scalars.

Used by the *general* solver paths (preconditioners M/Ml/Mr, custom ``inner``,
duck-typed operators): every vector statement of the reference loop is one
kernel launch on (n, k) CUDA tensors, while the handful of per-column scalars
lives on the host exactly as in the reference (one device->host read per inner
product).  The fused paths in cg.py / minres.py / gmres.py keep the scalars on
the device instead and never synchronise inside an iteration.
"""
from __future__ import annotations

import numpy as np
import torch

from .device import Ops


class Alg:
    def __init__(self, prob, inner=None):
        self.prob = prob
        self.ops = Ops(prob.n, prob.k, prob.device, comm=prob.comm)
        self._user_inner = None if inner is None else prob.inner(inner)
        self._slot = self.ops.slots(1)[0]
        self._cbuf = self.ops.slots(1)[0]

    # ---- scalars
    def coef(self, a):
        """host (k,) -> device (k,) coefficient tensor (a fresh one per call:
        launches are asynchronous and must not see a later overwrite)."""
        a = np.array(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1),
                                     (self.prob.k,)))  # writable copy
        return torch.from_numpy(a).to(self.prob.device)

    def inner(self, x, y):
        """<x, y> column-wise -> host float64 (k,).  Default: deterministic
        device reduction (_helpers.py:101-110); else the user's callable."""
        if self._user_inner is not None:
            return self._user_inner(x, y)
        self.ops.dot(x, y, self._slot)
        return self._slot.cpu().numpy().copy()

    # ---- vectors
    def apply(self, op, x):
        """op @ x as a new tensor (None == identity returns x itself, like
        the reference's Identity)."""
        return x if op is None else op(x)

    def apply_chain(self, ops_right_to_left, x):
        """reference Product.__matmul__ (_helpers.py:44-48); the defensive
        x.copy() of the reference is only materialised when every operator is
        the identity (otherwise each operator already returns a new tensor)."""
        out = x
        for op in ops_right_to_left:
            out = self.apply(op, out)
        return out.clone() if out is x else out

    def axpy(self, y, a, x, sign=1.0):
        self.ops.axpy(y, self.coef(a), x, sign)

    def xpby(self, y, x, a):
        self.ops.xpby(y, x, self.coef(a))

    def div(self, x, d, out=None):
        out = torch.empty_like(x) if out is None else out
        self.ops.div_scale(out, x, self.coef(d))
        return out

    def lincomb(self, x, ca=None, y=None, cb=None, out=None):
        """ca * x + cb * y with NumPy's rounding (each product, then the sum); ``ca=None``: x
        itself, ``y=None``: no second term.  ``out`` may be x or y."""
        out = torch.empty_like(x) if out is None else out
        self.ops.lincomb(out, None if ca is None else self.coef(ca), x,
                         None if y is None else self.coef(cb), y)
        return out

    def add(self, x, y):
        out = torch.empty_like(x)
        self.ops.add(out, x, y)
        return out

    def residual(self, A, b, z):
        """b - A z as a new tensor (fused into the product for CSR matrices)."""
        if getattr(A, "csr", None) is not None:
            r = torch.empty_like(b)
            self.ops.spmv(A.csr, z, r, mode=2, z=b)
            return r
        r = A(z)
        self.ops.xpby(r, b, self.coef(-1.0))  # r <- b + (-1) * r
        return r


def nz(d):
    return np.where(d != 0, d, 1.0)


def as_resnorm(prob, v):
    return prob.scalars_to_user(v)
