"""TEST INFRASTRUCTURE -- CPU restatement (NumPy) of the reference's ``krylov.utils``
(`/root/reference/src/krylov/utils.py`): block QR, principal angles, the Hegedues rescaling and the
three small host helpers.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs
may import this module; the product (``krylov_b200/utils.py``) never does.

Parity: **pinned** -- ``tests/golden/make_golden_utils.py`` runs the unmodified reference on seeded
inputs and stores its outputs in ``tests/golden/utils.npz``; ``tests/test_oracle_utils_golden.py``
checks every function below against them.

Third-party arithmetic reached through the installed NumPy (same as the reference, which does not
pin versions, setup.cfg:31-35): LAPACK ``geqrf/orgqr`` behind ``np.linalg.qr`` (utils.py:24) and
``gesdd`` behind ``np.linalg.svd`` (utils.py:100,113).
"""
import numpy as np


def qr(X, inner=None, reorthos=1):
    """utils.py:11-40.  With an inner product: modified Gram-Schmidt, column by column,
    ``reorthos + 1`` sweeps; columns whose norm falls below 1e-15 stay unnormalised.
    Without one the reference asks ``np.linalg.qr`` for ``mode="economic"`` (utils.py:24), which
    every NumPy >= 1.8 answers with a single array (geqrf's packed output), not the (Q, R) the
    docstring promises -- ``Q, R = qr(X)`` raises.  Restated as what is promised: the reduced
    LAPACK factorisation, whose R equals the upper triangle of that packed array (checked against
    the stored reference output in tests/test_oracle_utils_golden.py)."""
    n, k = X.shape
    if inner is None and k > 0:  # utils.py:23-24
        return np.linalg.qr(X, mode="reduced")
    Q = np.array(X, copy=True)
    R = np.zeros((k, k), dtype=X.dtype)
    for i in range(k):
        # list indices, not slices: the inner product then sees contiguous (n, 1) copies like the
        # reference's (utils.py:32) and BLAS takes the same summation path -- on ill-conditioned X
        # (Hilbert) a strided dot differs by 1e-12 in Q
        qi = [i]
        for _sweep in range(reorthos + 1):  # utils.py:30-34
            for j in range(i):
                qj = [j]
                a = inner(Q[:, qj], Q[:, qi])
                R[j, i] += a
                Q[:, qi] -= a * Q[:, qj]
        R[i, i] = np.sqrt(np.linalg.norm(inner(Q[:, qi], Q[:, qi]), 2))  # utils.py:36
        if R[i, i] >= 1e-15:
            Q[:, qi] /= R[i, i]
    return Q, R


def angles(F, G, inner=None, compute_vectors=False):
    """utils.py:43-141 (Knyazev & Argentati, algorithm 6.2): cosines from the SVD of
    ``inner(QF, QG)`` for the large angles, sines from the part of QG's principal vectors outside
    span(QF) for the small ones (sigma^2 >= 1/2)."""
    swapped = F.shape[1] < G.shape[1]  # utils.py:86-89
    if swapped:
        F, G = G, F
    k, l = F.shape[1], G.shape[1]
    QF, _ = qr(F, inner=inner)
    QG, _ = qr(G, inner=inner)

    if l == 0:  # utils.py:95-98
        theta = np.full(k, np.pi / 2)
        U, V = QF, QG
    else:
        Y, s, Zh = np.linalg.svd(inner(QF, QG))
        Vcos = QG @ Zh.T.conj()
        n_large = int(np.count_nonzero(s ** 2 < 0.5))
        n_small = s.shape[0] - n_large
        theta = np.concatenate([np.arccos(s[n_small:]), np.full(k - l, np.pi / 2)])
        if compute_vectors:
            Ucos = QF @ Y
            U, V = Ucos[:, n_small:], Vcos[:, n_small:]
        if n_small > 0:  # utils.py:116-135
            RG = Vcos[:, :n_small]
            S = RG - QF @ inner(QF, RG)
            _, R = qr(S, inner=inner)
            Y2, u, Z2h = np.linalg.svd(R)
            theta = np.concatenate([np.arcsin(u[::-1][:n_small]), theta])
            if compute_vectors:
                RF = Ucos[:, :n_small]
                Vsin = RG @ Z2h.T.conj()
                T = np.diag(1 / s[:n_small]) @ (Z2h.T.conj() @ np.diag(s[:n_small]))
                Usin = RF @ T
                U = np.column_stack([Usin, U])
                V = np.column_stack([Vsin, V])
    if not compute_vectors:
        return theta
    if swapped:
        U, V = V, U
    return theta, U, V


def hegedus(A, b, x0, M=None, Ml=None, inner=None):
    """utils.py:144-180: gamma x0 with gamma minimising ||M Ml (b - gamma A x0)||_{M^-1}."""
    Ax0 = A @ x0
    MlAx0 = Ax0 if Ml is None else Ml @ Ax0
    z = MlAx0 if M is None else M @ MlAx0
    znorm2 = inner(z, MlAx0)
    if znorm2 <= 1e-15:  # utils.py:176-177
        return np.zeros_like(b)
    Mlb = b if Ml is None else Ml @ b
    return (inner(z, Mlb) / znorm2) * x0


def strakos(n, l_min=0.1, l_max=100, rho=0.9):
    """utils.py:183-192: diag(l_min + (i-1)/(n-1) (l_max - l_min) rho^(n-i)), i = 1..n."""
    # Python ints and floats throughout (NumPy's integer power rounds differently in the last bit)
    return np.diag([l_min + (i - 1) * 1.0 / (n - 1) * (l_max - l_min) * rho ** (n - i)
                    for i in range(1, n + 1)])


def gap(lamda, sigma, mode="individual"):
    """utils.py:195-251."""
    lam = np.atleast_1d(np.asarray(lamda))
    sig = np.atleast_1d(np.asarray(sigma))
    if not (np.isreal(lam).all() and np.isreal(sig).all()):
        raise ValueError("complex spectra not yet implemented")
    if mode == "individual":
        return np.min(np.abs(lam[:, None] - sig[None, :]))
    if mode == "interval":
        lo, hi = lam.min(), lam.max()
        below, above = sig <= lo, sig >= hi
        if not np.all(below | above):
            return None
        delta = np.inf
        if below.any():
            delta = lo - sig[below].max()
        if above.any():
            delta = min(delta, sig[above].min() - hi)
        return delta
    return None


class NormalizedRootsPolynomial:
    """utils.py:254-316: p(x) = prod_i (1 - x / theta_i)."""

    def __init__(self, roots):
        roots = np.asarray(roots)
        if roots.ndim != 1:
            raise ValueError("one-dimensional array of roots expected.")
        self.roots = roots

    def minmax_candidates(self):
        from numpy.polynomial import Polynomial

        return Polynomial.fromroots(self.roots).deriv(1).roots()

    def __call__(self, points):
        p = np.asarray(points)
        if p.ndim > 1:
            raise ValueError("scalar or one-dimensional array of points expected.")
        n = self.roots.shape[0]
        vals = 1 - p / self.roots.reshape(n, 1)
        half = int(np.ceil(n / 2.0))
        for j in range(vals.shape[1]):  # interlace small and large factors (utils.py:303-309)
            order = np.argsort(np.abs(vals[:, j]))
            mix = np.zeros(n, dtype=int)
            mix[::2] = order[:half]
            mix[1::2] = order[half:][::-1]
            vals[:, j] = vals[mix, j]
        out = np.prod(vals, axis=0)
        return out.item() if np.isscalar(points) else out
