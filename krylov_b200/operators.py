"""Operator / inner-product protocol of the reference (``_helpers.py``) and
the adapters that let the device solvers call user objects.

Public mirrors (same names and meaning as the reference): ``Identity``,
``Product``, ``LinearOperatorWrapper``, ``aslinearoperator``, ``Info``,
``get_default_inner``.

Internal: ``Problem`` normalises ``(A, b, x0)`` into contiguous fp64 CUDA
tensors of shape (n, k) and remembers how the caller's arrays looked, so
results, callbacks, custom ``inner`` functions and duck-typed operators see
exactly the array kind and shape the reference would hand them (NumPy arrays
if ``b`` was NumPy, CUDA tensors if ``b`` was a CUDA tensor).
"""
from __future__ import annotations

import collections

import numpy as np
import torch

from .csr import CsrMatrix
from .device import as_device_matrix, require_cuda

# reference _helpers.py:93-98
Info = collections.namedtuple(
    "IterInfo",
    ["success", "xk", "numsteps", "resnorms", "num_operations", "arnoldi"],
    defaults=(None, None),
)


class Identity:
    """reference _helpers.py:26-36"""

    dtype = np.dtype("u1")

    @staticmethod
    def __matmul__(x):
        return x

    @staticmethod
    def rmatvec(x):
        return x


class Product:
    """reference _helpers.py:39-48: operators applied right to left."""

    def __init__(self, *operators):
        self.operators = operators
        self.dtype = np.result_type(*[np.dtype(op.dtype) for op in operators])

    def __matmul__(self, x):
        out = x.clone() if isinstance(x, torch.Tensor) else x.copy()
        for op in self.operators[::-1]:
            out = op @ out
        return out


class LinearOperatorWrapper:
    """reference _helpers.py:51-80: adds ``rmatvec`` to an array-like."""

    def __init__(self, array):
        self._array = array
        self._adj_array = None
        self.shape = array.shape
        self.dtype = array.dtype

    def __matmul__(self, x):
        return self._array @ x

    matvec = __matmul__

    def rmatvec(self, x):
        if isinstance(self._array, np.ndarray):
            return (self._array.T @ x.conj()).conj()
        if self._adj_array is None:
            self._adj_array = self._array.T.conj()
        return self._adj_array @ x


def aslinearoperator(A):
    """reference _helpers.py:83-90"""
    if not hasattr(A, "__matmul__"):
        raise ValueError(f"Unknown linear operator A = {A}")
    if hasattr(A, "rmatvec"):
        return A
    return LinearOperatorWrapper(A)


def get_default_inner(b_shape):
    """reference _helpers.py:101-110, for NumPy arrays and torch tensors."""

    def inner_dot(x, y):
        if isinstance(x, torch.Tensor):
            return torch.dot(x, y)
        return np.dot(x.conj(), y)

    def inner_einsum(x, y):
        if isinstance(x, torch.Tensor):
            return torch.einsum("i...,i...->...", x, y)
        return np.einsum("i...,i...->...", x.conj(), y)

    return inner_dot if len(b_shape) == 1 else inner_einsum


# ---------------------------------------------------------------------------
# internal adapters
# ---------------------------------------------------------------------------
def _is_scipy_sparse(A):
    try:
        import scipy.sparse

        return scipy.sparse.issparse(A)
    except Exception:  # pragma: no cover
        return False


def to_csr_or_none(A, device):
    """CsrMatrix for anything that *is* a matrix; None for duck-typed operators."""
    if isinstance(A, CsrMatrix) or getattr(A, "is_dist_csr", False):
        return A
    if _is_scipy_sparse(A):
        return CsrMatrix.from_scipy(A, device)
    if isinstance(A, np.ndarray) and A.ndim == 2:
        return CsrMatrix.from_dense(A, device)
    if isinstance(A, torch.Tensor) and A.dim() == 2:
        if A.is_complex():
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        return CsrMatrix.from_torch(A, device)
    return None


def _dtype_is_complex(A):
    dt = getattr(A, "dtype", None)
    if dt is None:
        return False
    if isinstance(dt, torch.dtype):
        return dt.is_complex
    try:
        return np.issubdtype(np.dtype(dt), np.complexfloating)
    except TypeError:
        return False


class Problem:
    """Normalised linear system on the device."""

    def __init__(self, A, b, x0=None):
        require_cuda()
        self.is_torch = isinstance(b, torch.Tensor)
        if not self.is_torch:
            b = np.asarray(b)
        # reference cg.py:99-101 / minres.py:83-85 / gmres.py:116-118
        assert len(A.shape) == 2
        assert A.shape[0] == A.shape[1]
        assert A.shape[1] == b.shape[0]
        if _dtype_is_complex(b) or _dtype_is_complex(A):
            raise NotImplementedError(
                "complex128: cg and minres accept HERMITIAN matrices (real-equivalent embedding, "
                "krylov_b200/_complex.py); gmres, the Arnoldi builders and the other solvers are "
                "fp64 only (north_star)")
        self.user_shape = tuple(b.shape)
        self.n = int(b.shape[0])
        self.k = int(np.prod(self.user_shape[1:])) if len(self.user_shape) > 1 else 1
        if self.is_torch and b.is_cuda:
            self.device = b.device
        else:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.b = as_device_matrix(b, self.device).reshape(self.n, self.k)
        if x0 is None:
            self.x0 = torch.zeros_like(self.b)
        else:
            self.x0 = as_device_matrix(x0, self.device).reshape(self.n, self.k).clone()
        self.A = self.operator(A)
        if self.A is None:
            raise ValueError("A must be a matrix or a linear operator")
        self.A_csr = self.A.csr  # CsrMatrix, or None for a duck-typed operator
        # row-partitioned matrix: b, x0 and every vector are this rank's rows
        self.comm = getattr(self.A_csr, "comm", None)

    def check_peers(self):
        """Row-partitioned problems: raise if a peer-memory exchange or all-reduce ever timed
        out (the kernels poison their results with NaN and set a sticky error word; called at
        every batch boundary of the solver loops)."""
        if self.comm is not None:
            chk = getattr(self.A_csr, "check_p2p", None) or getattr(self.comm, "check_p2p", None)
            if chk is not None:
                chk()

    def on_device(self):
        """Context manager: this problem's GPU is the current device."""
        return torch.cuda.device(self.device)

    # user-facing views -----------------------------------------------------
    def to_user(self, t: torch.Tensor):
        t = t.reshape(self.user_shape)
        if self.is_torch:
            return t
        if t.numel() >= (1 << 22):
            return _download(t)
        return t.cpu().numpy()

    def from_user(self, a) -> torch.Tensor:
        return as_device_matrix(a, self.device).reshape(self.n, self.k)

    def scalars_to_user(self, s):
        """(k,) host array -> what the reference's inner product returns for this
        shape of b: a NumPy scalar for 1-D b, an array of shape b.shape[1:] else."""
        s = np.asarray(s, dtype=np.float64)
        if len(self.user_shape) == 1:
            return np.float64(s.reshape(-1)[0])
        return s.reshape(self.user_shape[1:])

    # operators ---------------------------------------------------------------
    def operator(self, op):
        """callable(device (n,k) tensor) -> new device (n,k) tensor, or None
        for the identity."""
        if op is None or isinstance(op, Identity):
            return None
        csr = to_csr_or_none(op, self.device)
        if csr is not None:
            if csr.shape[0] != self.n or csr.shape[1] != self.n:
                raise ValueError("operator shape does not match the right-hand side")
            return _CsrApply(csr)
        if hasattr(op, "device_apply"):  # library-internal composite operators (A A^H, A^H A)
            return _DeviceApply(op)
        if not hasattr(op, "__matmul__"):
            raise ValueError(f"Unknown linear operator {op}")
        return _UserApply(op, self)

    def adjoint(self, applied):
        """``applied`` is what ``operator()`` returned (or ``self.A``).  Returns a callable
        (device (n,k) tensor -> new device tensor) for its adjoint -- the reference's ``rmatvec``
        (_helpers.py:51-90) -- or None for the identity.  Matrices use the cached transposed
        CsrMatrix; duck-typed operators must provide ``rmatvec``."""
        if applied is None:
            return None
        if applied.csr is not None:
            if getattr(applied.csr, "is_dist_csr", False):
                raise NotImplementedError("adjoint products of row-partitioned matrices")
            return _CsrApply(applied.csr.T)
        if not hasattr(applied.op, "rmatvec"):
            raise ValueError(f"operator {applied.op} has no rmatvec")
        return _UserApply(_Adjoint(applied.op), self)

    def inner(self, fn):
        """callable(x_dev, y_dev) -> host float64 array (k,) for a user inner product."""

        def call(x, y):
            v = fn(self.to_user(x), self.to_user(y))
            if isinstance(v, torch.Tensor):
                v = v.detach().cpu().numpy()
            v = np.asarray(v)
            if np.any(np.imag(v) != 0.0):
                raise ValueError("inner product <x, M x> gave nonzero imaginary part")
            v = np.real(v).astype(np.float64).reshape(-1)
            if v.size == 1 and self.k > 1:
                v = np.full(self.k, v[0])
            if v.size != self.k:
                raise ValueError(
                    f"inner product returned {v.size} values for {self.k} right-hand sides")
            return v

        return call


_STAGE = {}
_POOL = None


def _download(t: torch.Tensor) -> np.ndarray:
    """Large device tensor -> pageable ndarray through persistent 8 MB pinned staging buffers:
    the PCIe copy of chunk i+1 overlaps the host memcpy of earlier chunks, and the memcpys (which
    first-touch the pages of the fresh result, ~4 GB/s on one core) run on a few worker threads.
    A plain ``.cpu()`` into pageable memory runs at ~2 GB/s; pinning a fresh result-sized buffer
    per call costs more than the copy."""
    global _POOL
    from concurrent.futures import ThreadPoolExecutor

    # 8 x 8 MB pinned (pinning itself costs ~0.3 ms/MB, once); one worker per staging buffer: the
    # first-touch memcpy into the fresh result scales with threads (1.07 GB on 8 cores: 413 ms with
    # one or two threads, 120-190 ms with four, 80 ms with eight)
    import os

    nbuf = max(4, min(8, os.cpu_count() or 4))
    chunk = 1 << 20
    key = t.device.index
    if key not in _STAGE:
        _STAGE[key] = [torch.empty(chunk, dtype=torch.float64, pin_memory=True) for _ in range(nbuf)]
    nbuf = len(_STAGE[key])
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=nbuf, thread_name_prefix="kb-download")
    bufs = _STAGE[key]
    out = np.empty(tuple(t.shape), dtype=np.float64)
    flat, src = out.reshape(-1), t.reshape(-1)
    n = src.numel()

    def land(ev, buf, off, m):
        ev.synchronize()
        flat[off:off + m] = buf[:m].numpy()

    busy = [None] * nbuf
    for i, off in enumerate(range(0, n, chunk)):
        j = i % nbuf
        if busy[j] is not None:
            busy[j].result()  # the staging buffer is free again
        m = min(chunk, n - off)
        bufs[j][:m].copy_(src[off:off + m], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        busy[j] = _POOL.submit(land, ev, bufs[j], off, m)
    for f in busy:
        if f is not None:
            f.result()
    return out


class _CsrApply:
    def __init__(self, csr):
        self.csr = csr

    def __call__(self, x):
        return self.csr.matvec_device(x)


class _DeviceApply:
    """Operator that works on device tensors directly (no host round trip)."""

    csr = None

    def __init__(self, op):
        self.op = op

    def __call__(self, x):
        return self.op.device_apply(x)


class _Adjoint:
    def __init__(self, op):
        self.op = op

    def __matmul__(self, x):
        return self.op.rmatvec(x)


class _UserApply:
    """Duck-typed operator (``shape``/``dtype``/``__matmul__``, reference
    tests/test_solvers.py:212-243).  It is called with the caller's own array
    kind: host arrays make a device->host->device trip per application, which
    is the price of an opaque host operator, not a fallback of ours."""

    csr = None

    def __init__(self, op, prob):
        self.op, self.prob = op, prob

    def __call__(self, x):
        y = self.prob.from_user(self.op @ self.prob.to_user(x))
        if y.data_ptr() == x.data_ptr():  # identity-like operator: never alias the input
            y = y.clone()
        return y
