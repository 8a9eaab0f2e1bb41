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


_ADD, _SUB, _MUL, _DIV, _SQRT, _ABS, _NEG, _NZ, _COPY, _DIV_NZ = range(10)


class DevScalar:
    """Per-column scalars (k,) that stay on the device.  Arithmetic (+ - * /, negation, abs, sqrt,
    ** 2, with other DevScalars, Python numbers and NumPy (k,) arrays) is evaluated at once by
    ``kb_scalar_op`` -- one IEEE operation per launch, the host's bits -- and never synchronises;
    ``host()`` is the one read-back.  The short-recurrence solvers keep their NumPy-style scalar
    statements and lose the device->host round trip per inner product."""

    __slots__ = ("t", "alg")
    __array_priority__ = 1000.0

    def __init__(self, alg, t):
        self.alg, self.t = alg, t

    def host(self):
        return self.t.cpu().numpy().copy()

    # ---- evaluation
    def _operand(self, v):
        """-> (device tensor or None, immediate)"""
        if isinstance(v, DevScalar):
            return v.t, 0.0
        if isinstance(v, (int, float, np.floating, np.integer)):
            return None, float(v)
        a = np.asarray(v, dtype=np.float64)
        if a.size == 1:
            return None, float(a.reshape(-1)[0])
        return self.alg.coef(a), 0.0

    def _op(self, code, a, b=0.0):
        if code == _DIV and isinstance(b, _LazyNz) and b._t is None:
            # x / nz(y): one launch (the where() never materialises)
            ta, sa = self._operand(a)
            out = torch.empty(self.alg.prob.k, dtype=torch.float64, device=self.alg.prob.device)
            self.alg.ops.scalar_op(_DIV_NZ, ta, b.src.t, sa, b.fill, out)
            return DevScalar(self.alg, out)
        ta, sa = self._operand(a)
        tb, sb = self._operand(b)
        out = torch.empty(self.alg.prob.k, dtype=torch.float64, device=self.alg.prob.device)
        self.alg.ops.scalar_op(code, ta, tb, sa, sb, out)
        return DevScalar(self.alg, out)

    def __add__(self, o): return self._op(_ADD, self, o)
    def __radd__(self, o): return self._op(_ADD, o, self)
    def __sub__(self, o): return self._op(_SUB, self, o)
    def __rsub__(self, o): return self._op(_SUB, o, self)
    def __mul__(self, o): return self._op(_MUL, self, o)
    def __rmul__(self, o): return self._op(_MUL, o, self)
    def __truediv__(self, o): return self._op(_DIV, self, o)
    def __rtruediv__(self, o): return self._op(_DIV, o, self)
    def __neg__(self): return self._op(_NEG, self)
    def __abs__(self): return self._op(_ABS, self)

    def __pow__(self, e):
        if e == 2:
            return self._op(_MUL, self, self)
        raise TypeError("DevScalar supports ** 2 only")

    def sqrt(self): return self._op(_SQRT, self)
    def nz(self, fill=1.0): return _LazyNz(self, float(fill))

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__" or kwargs:
            return NotImplemented
        binary = {np.add: _ADD, np.subtract: _SUB, np.multiply: _MUL, np.true_divide: _DIV}
        unary = {np.sqrt: _SQRT, np.absolute: _ABS, np.negative: _NEG}
        if ufunc in binary and len(inputs) == 2:
            return self._op(binary[ufunc], inputs[0], inputs[1])
        if ufunc in unary and len(inputs) == 1:
            return self._op(unary[ufunc], inputs[0])
        if ufunc is np.square:
            return self._op(_MUL, inputs[0], inputs[0])
        raise TypeError(f"DevScalar does not implement {ufunc.__name__}")

    def __array__(self, *a, **kw):
        raise TypeError("DevScalar stays on the device: call to_host() where the host needs it")


class _LazyNz(DevScalar):
    """``nz(d)`` = ``where(d != 0, d, fill)``: almost always the denominator of the next
    statement, so it is evaluated only if something other than a division asks for its value."""

    __slots__ = ("src", "fill", "_t")

    def __init__(self, src, fill):
        self.alg, self.src, self.fill, self._t = src.alg, src, fill, None

    @property
    def t(self):
        if self._t is None:
            out = torch.empty(self.alg.prob.k, dtype=torch.float64, device=self.alg.prob.device)
            self.alg.ops.scalar_op(_NZ, self.src.t, None, 0.0, self.fill, out)
            self._t = out
        return self._t

    @t.setter
    def t(self, v):
        self._t = v


def to_host(v):
    """(k,) float64 host array of a scalar statement's value (DevScalar: the one read-back)."""
    return v.host() if isinstance(v, DevScalar) else v


class Alg:
    def __init__(self, prob, inner=None, lazy=False):
        self.prob = prob
        self.ops = Ops(prob.n, prob.k, prob.device, comm=prob.comm)
        self._user_inner = None if inner is None else prob.inner(inner)
        self._slot = self.ops.slots(1)[0]
        self._cbuf = self.ops.slots(1)[0]
        # lazy: inner products are returned as DevScalar (no read-back); only with the default
        # inner product, and only for callers written for it (the short-recurrence solvers)
        self.lazy = bool(lazy) and self._user_inner is None
        self._const = {}

    # ---- scalars
    def coef(self, a):
        """host (k,) -> device (k,) coefficient tensor (a fresh one per call:
        launches are asynchronous and must not see a later overwrite)."""
        if isinstance(a, DevScalar):
            return a.t  # immutable: every operation wrote a tensor of its own
        if isinstance(a, (int, float)):  # constants (1.0, -1.0, fixed step sizes): upload once
            t = self._const.get(float(a))
            if t is None:
                if len(self._const) > 64:
                    self._const.clear()
                t = self._const[float(a)] = torch.full((self.prob.k,), float(a), dtype=torch.float64,
                                                       device=self.prob.device)
            return t
        a = np.array(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1),
                                     (self.prob.k,)))  # writable copy
        return torch.from_numpy(a).to(self.prob.device)

    def inner(self, x, y):
        """<x, y> column-wise -> host float64 (k,).  Default: deterministic
        device reduction (_helpers.py:101-110); else the user's callable."""
        if self._user_inner is not None:
            return self._user_inner(x, y)
        if self.lazy:
            out = torch.empty(self.prob.k, dtype=torch.float64, device=self.prob.device)
            self.ops.dot(x, y, out)
            return DevScalar(self, out)
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


def nz(d, fill=1.0):
    if isinstance(d, DevScalar):
        return d.nz(fill)
    return np.where(d != 0, d, fill)


def as_resnorm(prob, v):
    return prob.scalars_to_user(v)
