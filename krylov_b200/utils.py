"""Device-side ``krylov.utils`` (reference utils.py:11-316): block QR with a customisable inner
product, principal angles between subspaces, the Hegedues rescaling of an initial guess, and the
three small host helpers (``strakos``, ``gap``, ``NormalizedRootsPolynomial``).

This is the one place of the reference where ``inner`` is a *block* inner product
(``inner(QF, QG)`` is a k x l matrix, utils.py:100,117).  The tall operands (n x k, n in the
millions) never leave the GPU:

* block Gram matrices and the n x k by k x l updates around them run on the FP64 tensor cores
  (``kb_block_gram`` / ``kb_block_apply``, DMMA m8n8k4; csrc/kb_block.cuh);
* the Gram-Schmidt sweeps of ``qr`` work on a column-major copy with the deterministic
  dot / axpy kernels of the solver path (every statement of utils.py:30-39 is one launch, scalars
  stay on the device);
* ``qr(X)`` without an inner product is Householder QR with LAPACK's conventions
  (``kb_house_make2`` + fused reflector applications);
* the k x l (<= a few dozen entries) SVDs of ``angles`` are NumPy calls on the host exactly as in
  the reference (utils.py:100,120): O(k^3) scalar work on a matrix that is a reduction result.

Inner products.  The reference takes any callable ``inner(X, Y) -> X^H Y``-like matrix.  Here:

* ``EuclideanInner()`` (or ``inner=None`` in ``angles`` / ``hegedus``, an additive extension: the
  reference raises ``TypeError`` there) and ``WeightedInner(B)`` (``X^T B Y`` with a sparse or
  dense symmetric positive definite ``B``) are evaluated on the device;
* any other callable is called with the caller's array kind, one device -> host -> device trip per
  evaluation -- the price of an opaque host callable, as for duck-typed operators in the solvers.

Real fp64 only (north_star); complex input raises ``NotImplementedError``.
"""
from __future__ import annotations

import numpy as np
import torch

from .csr import CsrMatrix
from .device import BlockOps, Ops, as_device_matrix, require_cuda
from .errors import ArgumentError
from .operators import Problem, to_csr_or_none

__all__ = ["qr", "angles", "hegedus", "strakos", "gap", "NormalizedRootsPolynomial",
           "EuclideanInner", "WeightedInner"]


# ----------------------------------------------------------------------------------------------
# inner products
# ----------------------------------------------------------------------------------------------
def _on(device):
    """context manager: ``device`` is the current CUDA device"""
    return torch.cuda.device(device)


def _dev2(a, device=None):
    """array-like (n,) or (n, k) -> contiguous fp64 CUDA tensor (n, k)"""
    t = as_device_matrix(a, device)
    return t.reshape(t.shape[0], -1) if t.dim() != 2 else t


def _like(t, proto, shape=None):
    """device tensor -> the array kind of ``proto`` (torch stays on the device)"""
    if shape is not None:
        t = t.reshape(shape)
    return t if isinstance(proto, torch.Tensor) else t.cpu().numpy()


class EuclideanInner:
    """``inner(X, Y) = X^T Y`` on the FP64 tensor cores.  Callable with NumPy arrays or torch
    tensors of shape (n,) / (n, k) like the reference's ``lambda x, y: np.dot(x.T.conj(), y)``
    (tests/helpers.py:106); ``qr`` / ``angles`` / ``hegedus`` recognise it and keep everything on
    the device."""

    def _dev(self, bo, X, Y):
        return bo.gram(X, Y)

    def _cols(self, ops, u, w, out, tmp):
        """out[0] = <u, w> for contiguous length-n device vectors"""
        ops.dot(u, w, out)

    def __call__(self, X, Y):
        require_cuda()
        Xd = _dev2(X)
        Yd = _dev2(Y, Xd.device)
        with _on(Xd.device):
            G = self._dev(BlockOps(Xd.device), Xd, Yd)
        one_d = len(X.shape) == 1 and len(Y.shape) == 1
        return _like(G, X, () if one_d else None)


class WeightedInner(EuclideanInner):
    """``inner(X, Y) = X^T (B Y)``, B a (sparse or dense) matrix or a 1-D array of diagonal
    weights -- the second inner product of the reference's tests (tests/helpers.py:107)."""

    def __init__(self, B, device=None):
        require_cuda()
        if not isinstance(B, CsrMatrix):
            arr = B if hasattr(B, "shape") else np.asarray(B)
            if len(arr.shape) == 1:
                import scipy.sparse

                w = np.asarray(arr.cpu().numpy() if isinstance(arr, torch.Tensor) else arr,
                               dtype=np.float64)
                arr = scipy.sparse.diags(w).tocsr()
            csr = to_csr_or_none(arr, device)
            if csr is None:
                raise ValueError("B must be a matrix or a 1-D array of weights")
            B = csr
        self.B = B

    def _dev(self, bo, X, Y):
        return bo.gram(X, self.B.matvec_device(Y.contiguous()))

    def _cols(self, ops, u, w, out, tmp):
        ops.spmv(self.B, w, tmp, dot=1, w=u, out=out)  # B w and <u, B w> in one launch


class _HostInner:
    """Opaque user callable: evaluated on the caller's array kind."""

    def __init__(self, fn, proto, one_d=False):
        self.fn, self.proto, self.one_d = fn, proto, one_d

    def _call(self, X, Y, shape):
        args = [_like(t, self.proto, (t.shape[0],) if self.one_d else None) for t in (X, Y)]
        v = self.fn(*args)
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        v = np.asarray(v)
        if np.iscomplexobj(v):
            if np.any(v.imag != 0.0):
                raise NotImplementedError("complex inner products are out of scope (north_star: fp64)")
            v = v.real
        return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).reshape(shape)).to(X.device)

    def _dev(self, bo, X, Y):
        return self._call(X, Y, (X.shape[1], Y.shape[1]))

    def _cols(self, ops, u, w, out, tmp):
        out.copy_(self._call(u.reshape(-1, 1), w.reshape(-1, 1), (1,)))


def _resolve_inner(inner, proto, one_d=False):
    if inner is None or (isinstance(inner, str) and inner == "euclidean"):
        return EuclideanInner()
    if isinstance(inner, EuclideanInner):
        return inner
    if not callable(inner):
        raise TypeError("inner must be None, an EuclideanInner / WeightedInner or a callable")
    return _HostInner(inner, proto, one_d)


# ----------------------------------------------------------------------------------------------
# QR
# ----------------------------------------------------------------------------------------------
def _qr_mgs(Xd, kind, reorthos):
    """utils.py:26-40 on the device.  Columns live as contiguous vectors (a column-major copy), so
    every statement is one coalesced launch: alpha = inner(q_j, q_i) (deterministic reduction into
    a device slot), R[j, i] += alpha, q_i -= alpha q_j (product rounded, then the difference, as
    NumPy), R[i, i] = sqrt(|inner(q_i, q_i)|), q_i /= R[i, i] unless it is below 1e-15."""
    n, k = Xd.shape
    dev = Xd.device
    R = torch.zeros((k, k), dtype=torch.float64, device=dev)
    if k == 0 or n == 0:
        return Xd.clone(), R
    Xt = Xd.t().contiguous()  # (k, n): row j = column j of X
    cols = [Xt[j].reshape(n, 1) for j in range(k)]
    ops = Ops(n, 1, dev)
    slot = ops.slots(2)
    tmp = ops.vec(zero=False)
    one = torch.ones((1,), dtype=torch.float64, device=dev)
    for i in range(k):
        qi = cols[i]
        for _sweep in range(reorthos + 1):
            for j in range(i):
                kind._cols(ops, cols[j], qi, slot[0], tmp)
                R[j, i] += slot[0, 0]
                ops.axpy(qi, slot[0], cols[j], sign=-1.0)
        kind._cols(ops, qi, qi, slot[1], tmp)
        rii = torch.sqrt(torch.abs(slot[1]))  # utils.py:36 (2-norm of a 1 x 1 matrix)
        R[i, i] = rii[0]
        cols[i] = ops.vec(zero=False)  # utils.py:37-38, out of place
        ops.div_scale(cols[i], qi, torch.where(rii >= 1e-15, rii, one))
    return torch.cat(cols, dim=1), R


def _apply_reflectors(ops, hv, params, js, w, tau):
    """w <- H_{js[-1]} ... H_{js[0]} w: tau = <v_j, w>, w -= beta_j tau v_j, each update fused with
    the next reflector's dot (32 B/element per reflector)."""
    if not js:
        return
    ops.dot(hv[js[0]], w, tau[0])
    t = 0
    for idx, j in enumerate(js):
        nxt = hv[js[idx + 1]] if idx + 1 < len(js) else None
        ops.axpy_dot(tau[t], hv[j], w, dot=1 if nxt is not None else 0, z=nxt, out=tau[1 - t],
                     scale=params[j][1:2])
        t = 1 - t


def _qr_householder(Xd):
    """Householder QR with LAPACK's conventions (dgeqr2 + dorg2r: what ``np.linalg.qr`` runs for
    tall-skinny blocks, utils.py:24): R[j, j] = -sign(pivot) ||tail||, SIGN(., 0) = +, H = I for a
    zero tail.  Left-looking: column c receives H_0 ... H_{c-1} (fused dot/axpy chain), then
    yields reflector c.  Q = H_0 ... H_{r-1} [e_0 ... e_{m-1}]."""
    n, k = Xd.shape
    dev = Xd.device
    m = min(n, k)
    R = torch.zeros((m, k), dtype=torch.float64, device=dev)
    if m == 0:
        return torch.zeros((n, m), dtype=torch.float64, device=dev), R
    W = Xd.t().contiguous()  # (k, n)
    ops = Ops(n, 1, dev)
    tau = ops.slots(2)
    scratch = ops.slots(1)
    nref = min(k, n - 1)
    hv, params = [], []
    for c in range(k):
        w = W[c].reshape(n, 1)
        _apply_reflectors(ops, hv, params, list(range(min(c, nref))), w, tau)
        rows = min(c + 1, m)
        R[:rows, c].copy_(W[c][:rows])
        if c < nref:
            v = ops.vec(zero=False)
            pr = torch.empty((5,), dtype=torch.float64, device=dev)
            ops.house_make(c, w, v, pr, scratch[0], lapack_sign=True)
            hv.append(v)
            params.append(pr)
            R[c, c] = pr[0] * pr[2]  # alpha ||tail|| = LAPACK's beta
    E = torch.zeros((m, n), dtype=torch.float64, device=dev)
    E[torch.arange(m, device=dev), torch.arange(m, device=dev)] = 1.0
    for c in range(m):
        e = E[c].reshape(n, 1)
        _apply_reflectors(ops, hv, params, list(range(min(c, nref - 1), -1, -1)), e, tau)
    return E.t().contiguous(), R


def qr(X, inner=None, reorthos=1):
    """QR factorisation with a customisable inner product (utils.py:11-40).

    ``inner`` given: modified Gram-Schmidt, ``reorthos + 1`` sweeps -> ``Q (n, k)``, ``R (k, k)``
    with ``inner(Q, Q) = I``.  ``inner=None``: the reference calls
    ``np.linalg.qr(X, mode="economic")``, which current NumPy answers with one packed array (so
    ``Q, R = qr(X)`` raises there); this returns the (Q, R) its docstring promises, LAPACK's
    reduced factorisation (same reflectors, same signs: R equals the upper triangle of that packed
    array)."""
    require_cuda()
    if len(X.shape) != 2:
        raise ValueError("X must have shape (N, k)")
    Xd = _dev2(X)
    with _on(Xd.device):
        if inner is None and Xd.shape[1] > 0:
            Q, R = _qr_householder(Xd)
        else:
            Q, R = _qr_mgs(Xd, _resolve_inner(inner, X), reorthos)
    return _like(Q, X), _like(R, X)


# ----------------------------------------------------------------------------------------------
# principal angles
# ----------------------------------------------------------------------------------------------
def angles(F, G, inner=None, compute_vectors=False):
    """Principal angles between span(F) and span(G) (utils.py:43-141; Knyazev & Argentati 2002,
    algorithm 6.2: cosine branch for the large angles, sine branch for sigma^2 >= 1/2).

    Returns ``theta`` (ascending, shape ``(max(k, l),)``) or ``theta, U, V``.  The tall blocks stay
    on the device; only the k x l matrix ``inner(QF, QG)`` and the small R of the sine branch go
    to the host for their SVD, as in the reference."""
    require_cuda()
    if len(F.shape) != 2 or len(G.shape) != 2 or F.shape[0] != G.shape[0]:
        raise ValueError("F and G must have shapes (N, k) and (N, l)")
    proto = F
    Fd = _dev2(F)
    Gd = _dev2(G, Fd.device)
    dev = Fd.device
    reverse = Fd.shape[1] < Gd.shape[1]  # utils.py:86-89
    if reverse:
        Fd, Gd = Gd, Fd
    k, l = Fd.shape[1], Gd.shape[1]
    kind = _resolve_inner(inner, proto)
    U = V = None
    with _on(dev):
        bo = BlockOps(dev)
        QF, _ = _qr_mgs(Fd, kind, 1)
        QG, _ = _qr_mgs(Gd, kind, 1)

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

        if l == 0:  # utils.py:95-98
            theta = np.ones(k) * np.pi / 2
            U, V = QF, QG
        else:
            C = kind._dev(bo, QF, QG).cpu().numpy()
            Y, s, Zh = np.linalg.svd(C)  # utils.py:100
            Vcos = bo.apply(QG, up(Zh.T))
            n_large = np.flatnonzero((s ** 2) < 0.5).shape[0]
            n_small = s.shape[0] - n_large
            theta = np.hstack([np.arccos(s[n_small:]), np.ones(k - l) * np.pi / 2])
            if compute_vectors:
                Ucos = bo.apply(QF, up(Y))
                U, V = Ucos[:, n_small:], Vcos[:, n_small:]
            if n_small > 0:  # utils.py:116-135
                RG = Vcos[:, :n_small]
                S = bo.apply(QF, kind._dev(bo, QF, RG), Y=RG, sign=-1)  # RG - QF <QF, RG>
                _, Rs = _qr_mgs(S, kind, 1)
                Y2, u, Z2h = np.linalg.svd(Rs.cpu().numpy())
                theta = np.hstack([np.arcsin(u[::-1][:n_small]), theta])
                if compute_vectors:
                    RF = Ucos[:, :n_small]
                    Vsin = bo.apply(RG, up(Z2h.T))
                    T = np.dot(np.diag(1 / s[:n_small]), np.dot(Z2h.T, np.diag(s[:n_small])))
                    Usin = bo.apply(RF, up(T))
                    U = torch.cat([Usin, U], dim=1)
                    V = torch.cat([Vsin, V], dim=1)
    theta_out = torch.from_numpy(theta).to(dev) if isinstance(proto, torch.Tensor) else theta
    if not compute_vectors:
        return theta_out
    if reverse:
        U, V = V, U
    return theta_out, _like(U.contiguous(), proto), _like(V.contiguous(), proto)


# ----------------------------------------------------------------------------------------------
# Hegedues trick
# ----------------------------------------------------------------------------------------------
def hegedus(A, b, x0, M=None, Ml=None, inner=None):
    """Rescale the initial guess to ``gamma x0`` with ``gamma`` minimising
    ``||M Ml (b - gamma A x0)||_{M^-1}`` (utils.py:144-180).  Operators as for the solvers
    (matrices run as device CSR products); returns zeros when ``<z, Ml A x0> <= 1e-15``."""
    require_cuda()
    if int(np.prod(tuple(x0.shape))) != int(np.prod(tuple(b.shape))):
        raise ValueError("x0 and b differ in size")
    prob = Problem(A, b, x0)
    if prob.k != 1:
        raise ValueError("hegedus works on a single right-hand side (the reference's truth test "
                         "`znorm2 <= 1e-15` is ambiguous for blocks)")
    one_d = len(prob.user_shape) == 1
    kind = _resolve_inner(inner, b, one_d)
    with prob.on_device():
        bo = BlockOps(prob.device)
        ops = Ops(prob.n, 1, prob.device)
        Ml_op, M_op = prob.operator(Ml), prob.operator(M)
        Ax0 = prob.A(prob.x0)
        MlAx0 = Ax0 if Ml_op is None else Ml_op(Ax0)
        z = MlAx0 if M_op is None else M_op(MlAx0)
        znorm2 = float(kind._dev(bo, z, MlAx0).cpu().numpy().reshape(-1)[0])
        if znorm2 <= 1e-15:  # utils.py:176-177
            out = torch.zeros_like(prob.b)
        else:
            Mlb = prob.b if Ml_op is None else Ml_op(prob.b)
            num = float(kind._dev(bo, z, Mlb).cpu().numpy().reshape(-1)[0])
            gamma = torch.full((1,), num / znorm2, dtype=torch.float64, device=prob.device)
            out = torch.empty_like(prob.x0)
            ops.lincomb(out, gamma, prob.x0)  # gamma * x0, rounded like NumPy's product
    return _like(out, b, tuple(x0.shape))


# ----------------------------------------------------------------------------------------------
# small host helpers (O(number of eigenvalues) scalar work on host inputs -- no device work to do)
# ----------------------------------------------------------------------------------------------
def strakos(n, l_min=0.1, l_max=100, rho=0.9):
    """The Strakos matrix diag(l_min + (i-1)/(n-1) (l_max-l_min) rho^(n-i)) (utils.py:183-192)."""
    return np.diag([l_min + (i - 1) * 1.0 / (n - 1) * (l_max - l_min) * rho ** (n - i)
                    for i in range(1, n + 1)])


def gap(lamda, sigma, mode="individual"):
    """Spectral gap between two sets of reals (utils.py:195-251): ``"individual"`` the smallest
    distance, ``"interval"`` the distance of sigma to the hull of lamda (None if sigma enters it)."""
    lam = np.array([lamda] if np.isscalar(lamda) else lamda)
    sig = np.array([sigma] if np.isscalar(sigma) else sigma)
    if not np.isreal(lam).all() or not np.isreal(sig).all():
        raise ArgumentError("complex spectra not yet implemented")
    if mode == "individual":
        return np.min(np.abs(lam.reshape(-1, 1) - sig.reshape(1, -1)))
    if mode == "interval":
        lo, hi = np.min(lam), np.max(lam)
        below, above = sig <= lo, sig >= hi
        if not np.all(below + above):
            return None
        delta = np.inf
        if np.any(below):
            delta = lo - np.max(sig[below])
        if np.any(above):
            delta = np.min([delta, np.min(sig[above]) - hi])
        return delta
    return None


class NormalizedRootsPolynomial:
    """p(x) = prod_i (1 - x / theta_i), p(0) = 1 (utils.py:254-316)."""

    def __init__(self, roots):
        roots = np.asarray(roots)
        if len(roots.shape) != 1:
            raise ArgumentError("one-dimensional array of roots expected.")
        self.roots = roots

    def minmax_candidates(self):
        """Zeros of p' (extrema candidates on an interval, utils.py:274-285)."""
        from numpy.polynomial import Polynomial

        return Polynomial.fromroots(self.roots).deriv(1).roots()

    def __call__(self, points):
        p = np.asarray(points)
        if len(p.shape) > 1:
            raise ArgumentError("scalar or one-dimensional array of points expected.")
        n = self.roots.shape[0]
        vals = 1 - p / self.roots.reshape(n, 1)
        half = int(np.ceil(float(n) / 2))
        for j in range(vals.shape[1]):  # small and large factors interlaced against over/underflow
            order = np.argsort(np.abs(vals[:, j]))
            mix = np.zeros((n,), dtype=int)
            mix[::2] = order[:half]
            mix[1::2] = order[half:][::-1]
            vals[:, j] = vals[mix, j]
        vals = np.prod(vals, axis=0)
        return vals.item() if np.isscalar(points) else vals
