"""Test double of the device layer for the HOST LOGIC of krylov_b200/shortrec.py (CPU suite only).

The product has no CPU path; this file is test infrastructure: it swaps ``Problem`` and ``Alg``
inside krylov_b200.shortrec for stand-ins that keep the (n, k) tensors in host memory and evaluate
each vector statement with torch CPU ops in the rounding order the CUDA kernels use.  What it checks
is the solver loops (state handling, scalar recurrences, stopping rule, callbacks, Info) against the
reference's golden outputs without a GPU; the kernels themselves are checked by the -m gpu tests."""
import contextlib

import numpy as np
import torch

from krylov_b200.operators import Identity


class _Apply:
    def __init__(self, M):
        self.M, self.csr, self.op = M, None, M

    def __call__(self, x):
        return torch.from_numpy(np.ascontiguousarray(self.M @ x.numpy()))


class FakeProblem:
    def __init__(self, A, b, x0=None):
        self.is_torch = False
        b = np.asarray(b, dtype=np.float64)
        assert len(A.shape) == 2 and A.shape[0] == A.shape[1] == b.shape[0]
        self.user_shape = tuple(b.shape)
        self.n = b.shape[0]
        self.k = int(np.prod(b.shape[1:])) if b.ndim > 1 else 1
        self.device = None
        self.comm = None
        self.b = torch.from_numpy(b.reshape(self.n, self.k).copy())
        self.x0 = (torch.zeros_like(self.b) if x0 is None else
                   torch.from_numpy(np.asarray(x0, dtype=np.float64).reshape(self.n, self.k).copy()))
        self.A = self.operator(A)
        self.A_csr = None
        self.launches = 0

    def on_device(self):
        return contextlib.nullcontext()

    def operator(self, op):
        if op is None or isinstance(op, Identity):
            return None
        return _Apply(op)

    def adjoint(self, applied):
        if applied is None:
            return None
        M = applied.M
        return _Apply(M.T.conj() if not isinstance(M, np.ndarray) else np.ascontiguousarray(M.T))

    def to_user(self, t):
        return t.numpy().reshape(self.user_shape)

    def scalars_to_user(self, s):
        s = np.asarray(s, dtype=np.float64)
        return np.float64(s.reshape(-1)[0]) if len(self.user_shape) == 1 else s.reshape(self.user_shape[1:])

    def inner(self, fn):
        def call(x, y):
            return np.asarray(fn(self.to_user(x), self.to_user(y)), dtype=np.float64).reshape(-1)
        return call


class _Ops:
    launches = 0


class FakeAlg:
    def __init__(self, prob, inner=None):
        self.prob, self.ops = prob, _Ops()
        self._user_inner = None if inner is None else prob.inner(inner)

    def _c(self, a):
        return torch.from_numpy(np.array(np.broadcast_to(np.asarray(a, dtype=np.float64).reshape(-1),
                                                         (self.prob.k,))))

    def inner(self, x, y):
        if self._user_inner is not None:
            return self._user_inner(x, y)
        return (x * y).sum(dim=0).numpy().copy()

    def apply(self, op, x):
        return x if op is None else op(x)

    def axpy(self, y, a, x, sign=1.0):
        y += (sign * self._c(a)) * x

    def xpby(self, y, x, a):
        y.copy_(x + self._c(a) * y)

    def add(self, x, y):
        return x + y

    def div(self, x, d, out=None):
        dd = self._c(d)
        r = x / torch.where(dd != 0, dd, torch.ones_like(dd))
        if out is None:
            return r
        out.copy_(r)
        return out

    def lincomb(self, x, ca=None, y=None, cb=None, out=None):
        t = x if ca is None else self._c(ca) * x
        if y is not None:
            t = t + self._c(cb) * y
        elif ca is None:
            t = t.clone()
        if out is None:
            return t
        out.copy_(t)
        return out

    def residual(self, A, b, z):
        return b - A(z)


@contextlib.contextmanager
def host_logic():
    import krylov_b200.shortrec as sr

    saved = sr.Problem, sr.Alg
    sr.Problem, sr.Alg = FakeProblem, FakeAlg
    try:
        yield sr
    finally:
        sr.Problem, sr.Alg = saved
