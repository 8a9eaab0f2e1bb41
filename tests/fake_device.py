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
    def __init__(self, prob, inner=None, lazy=False):  # lazy: device-resident scalars (GPU only)
        self.prob, self.ops = prob, _Ops()
        self.lazy = False  # host scalars: the per-iteration loop of _Drive.run
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


# ------------------------------------------------------------------------------------------------
# krylov_b200.utils: the same idea for qr / angles / hegedus (host logic on CPU tensors)
# ------------------------------------------------------------------------------------------------
class _FakeCsr:
    def __init__(self, M):
        import scipy.sparse

        self.M = scipy.sparse.csr_matrix(M)
        self.shape = self.M.shape

    def matvec_device(self, x):
        return torch.from_numpy(np.ascontiguousarray(self.M @ x.numpy()))


class FakeOps:
    """Stand-in for krylov_b200.device.Ops on CPU tensors: one method per kernel, evaluated in the
    kernels' rounding order (products rounded before sums; reductions as plain sums)."""

    def __init__(self, n, k, device=None, comm=None):
        self.n, self.k, self.launches = n, k, 0

    def vec(self, zero=True):
        return torch.zeros((self.n, self.k), dtype=torch.float64)

    def slots(self, m=1):
        return torch.zeros((m, self.k), dtype=torch.float64)

    def dot(self, x, y, out, n=None):
        out.copy_((x * y).sum(dim=0))

    def axpy(self, y, coef, x, sign=1.0):
        y += (sign * coef) * x

    def div_scale(self, out, x, coef):
        out.copy_(x / torch.where(coef != 0, coef, torch.ones_like(coef)))

    def lincomb(self, out, ca, x, cb=None, y=None):
        t = x if ca is None else ca * x
        if y is not None:
            t = t + cb * y
        out.copy_(t)

    def axpy_dot(self, coef, u, w, dot=0, z=None, out=None, scale=None):
        a = coef if scale is None else scale[0] * coef
        w -= a * u
        if dot == 1:
            out.copy_((z * w).sum(dim=0))
        elif dot == 2:
            out.copy_((w * w).sum(dim=0))

    def spmv(self, A, x, y, mode=0, z=None, coef=None, dot=0, w=None, out=None):
        assert mode == 0
        y.copy_(A.matvec_device(x))
        if dot == 1:
            out.copy_((w * y).sum(dim=0))

    def house_make(self, off, x, v, params, scratch, lapack_sign=False):
        """kb_house_make2: csrc/kb_scalar.cuh kb_house_params_kernel + kb_house_fill_kernel"""
        xs = x.reshape(-1)
        gamma = float(xs[off])
        sigma2 = float((xs[off + 1:] * xs[off + 1:]).sum())
        v0, xnorm = 1.0, float(np.sqrt(gamma * gamma + sigma2))
        if sigma2 == 0.0:
            beta, xnorm = 0.0, abs(gamma)
            alpha = 1.0 if gamma == 0.0 else gamma / xnorm
        else:
            beta = 2.0
            if gamma == 0.0:
                v0 = np.sqrt(sigma2) if lapack_sign else -np.sqrt(sigma2)
                alpha = -1.0 if lapack_sign else 1.0
            else:
                v0 = gamma + gamma / abs(gamma) * xnorm
                alpha = -gamma / abs(gamma)
        d = float(np.sqrt(v0 * v0 + sigma2))
        params.copy_(torch.tensor([alpha, beta, xnorm, v0, d], dtype=torch.float64))
        vv = v.reshape(-1)
        vv[:off] = 0.0
        vv[off] = v0 / d
        vv[off + 1:] = xs[off + 1:] / d


class FakeBlockOps:
    def __init__(self, device=None):
        self.launches = 0

    def gram(self, X, Y, out=None, acc=None, sqrt_abs=False):
        G = X.t() @ Y
        if sqrt_abs:
            G = torch.sqrt(torch.abs(G))
        if acc is not None:
            acc += G
        if out is not None:
            out.copy_(G)
            return out
        return G

    def apply(self, X, C, Y=None, sign=0, out=None):
        Z = X @ C
        if sign != 0:
            Z = Y + sign * Z
        if out is not None:
            out.copy_(Z)
            return out
        return Z


def _host_matrix(v, device=None):
    if isinstance(v, torch.Tensor):
        return v.to(dtype=torch.float64).contiguous()
    a = np.asarray(v)
    if np.iscomplexobj(a):
        raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())


class _UtilsProblem(FakeProblem):
    def __init__(self, A, b, x0=None):
        super().__init__(A, b, x0)
        self.device = torch.device("cpu")


@contextlib.contextmanager
def utils_host_logic():
    import krylov_b200.utils as ku

    names = ("Ops", "BlockOps", "Problem", "require_cuda", "as_device_matrix", "to_csr_or_none", "_on")
    saved = {n: getattr(ku, n) for n in names}
    ku.Ops, ku.BlockOps, ku.Problem = FakeOps, FakeBlockOps, _UtilsProblem
    ku.require_cuda = lambda: None
    ku.as_device_matrix = _host_matrix
    ku.to_csr_or_none = lambda A, device=None: _FakeCsr(A)
    ku._on = lambda device: contextlib.nullcontext()
    try:
        yield ku
    finally:
        for n, v in saved.items():
            setattr(ku, n, v)
