"""Synthetic stencil matrices of BASELINE.json / SURVEY.md section 8d.

Lexicographic ordering (x fastest), homogeneous Dirichlet boundary, unscaled
stencils, CSR with sorted int32 column indices and fp64 values.  The host
generators here are vectorised NumPy index arithmetic; the device generator
(``device_stencil7`` -> ``kb_stencil7_*`` kernels) builds the same CSR arrays
in HBM without a host copy (needed for the 512^3 case: 11.8 GB of CSR).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "stencil7_csr",
    "stencil5_csr",
    "poisson2d",
    "poisson3d",
    "shifted_laplace3d",
    "convection_diffusion3d",
    "mild_shift",
    "to_scipy",
    "STENCIL_POISSON",
    "convdiff_coeffs",
]

# (diag, lower_x, lower_y, lower_z, upper_x, upper_y, upper_z)
STENCIL_POISSON = (6.0, -1.0, -1.0, -1.0, -1.0, -1.0, -1.0)


def convdiff_coeffs(gamma=(0.5, 0.25, 0.125)):
    """C3 of SURVEY.md 8d: diag 6, lower -1-gamma_d, upper -1+gamma_d."""
    gx, gy, gz = gamma
    return (6.0, -1.0 - gx, -1.0 - gy, -1.0 - gz, -1.0 + gx, -1.0 + gy, -1.0 + gz)


def stencil7_csr(nx, ny, nz, coeffs=STENCIL_POISSON, shift=0.0, z_lo=0, z_hi=None):
    """Rows of the 7-point operator for grid planes ``z_lo <= z < z_hi``.

    Returns ``(rowptr int32 (n_loc+1,), cols int32 (nnz,), vals f64 (nnz,))``
    with *global* column indices.  Entry order per row is ascending column:
    z-1, y-1, x-1, diag, x+1, y+1, z+1.
    """
    z_hi = nz if z_hi is None else z_hi
    diag, lx, ly, lz, ux, uy, uz = coeffs
    n_loc = nx * ny * (z_hi - z_lo)
    row = np.arange(n_loc, dtype=np.int64) + np.int64(z_lo) * nx * ny
    ix = row % nx
    iy = (row // nx) % ny
    iz = row // (nx * ny)
    offs = np.array([-nx * ny, -nx, -1, 0, 1, nx, nx * ny], dtype=np.int64)
    cval = np.array([lz, ly, lx, diag - shift, ux, uy, uz], dtype=np.float64)
    mask = np.stack(
        [iz > 0, iy > 0, ix > 0, np.ones(n_loc, bool), ix < nx - 1, iy < ny - 1, iz < nz - 1],
        axis=1,
    )
    cols = (row[:, None] + offs[None, :])[mask]
    vals = np.broadcast_to(cval, (n_loc, 7))[mask]
    rowptr = np.zeros(n_loc + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    assert rowptr[-1] < 2**31
    return rowptr.astype(np.int32), cols.astype(np.int32), np.ascontiguousarray(vals)


def stencil5_csr(nx, ny, diag=4.0, off=-1.0):
    """2-D 5-point operator (C1): diag 4, off-diagonals -1."""
    n = nx * ny
    row = np.arange(n, dtype=np.int64)
    ix = row % nx
    iy = row // nx
    offs = np.array([-nx, -1, 0, 1, nx], dtype=np.int64)
    cval = np.array([off, off, diag, off, off], dtype=np.float64)
    mask = np.stack([iy > 0, ix > 0, np.ones(n, bool), ix < nx - 1, iy < ny - 1], axis=1)
    cols = (row[:, None] + offs[None, :])[mask]
    vals = np.broadcast_to(cval, (n, 5))[mask]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    return rowptr.astype(np.int32), cols.astype(np.int32), np.ascontiguousarray(vals)


def to_scipy(csr, n_cols=None):
    import scipy.sparse

    rowptr, cols, vals = csr
    n_rows = len(rowptr) - 1
    n_cols = n_rows if n_cols is None else n_cols
    return scipy.sparse.csr_matrix((vals, cols, rowptr), shape=(n_rows, n_cols))


def poisson2d(n):
    return to_scipy(stencil5_csr(n, n))


def poisson3d(n):
    return to_scipy(stencil7_csr(n, n, n))


def mild_shift(n):
    """sigma = (lambda_111 + lambda_211)/2 of the n^3 Dirichlet Laplacian:
    exactly one negative eigenvalue after the shift."""
    s = lambda i: 4.0 * np.sin(i * np.pi / (2.0 * (n + 1))) ** 2
    l111 = 3.0 * s(1)
    l211 = s(2) + 2.0 * s(1)
    return 0.5 * (l111 + l211)


def shifted_laplace3d(n, sigma=None):
    sigma = mild_shift(n) if sigma is None else sigma
    return to_scipy(stencil7_csr(n, n, n, shift=sigma))


def convection_diffusion3d(n, gamma=(0.5, 0.25, 0.125)):
    return to_scipy(stencil7_csr(n, n, n, coeffs=convdiff_coeffs(gamma)))


def fem27_var(n, seed=0):
    """27-point pattern on an n^3 grid with VARIABLE coefficients (a trilinear finite-element
    stiffness matrix of a heterogeneous medium has this shape): symmetric, strictly diagonally
    dominant, so SPD.  Not a BASELINE matrix -- a test / bench case for the CSR-streaming
    schedules (no constant diagonals, 27 entries per interior row)."""
    import scipy.sparse

    rng = np.random.default_rng(seed)
    idx = np.arange(n ** 3).reshape(n, n, n)
    rows, cols, vals = [], [], []
    offs = [(a, b, c) for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 0, 1)
            if (a, b, c) > (0, 0, 0)]
    for a, b, c in offs:
        src = idx[max(0, -a):n - max(0, a), max(0, -b):n - max(0, b), max(0, -c):n - max(0, c)]
        dst = idx[max(0, a):n - max(0, -a), max(0, b):n - max(0, -b), max(0, c):n - max(0, -c)]
        w = -rng.uniform(0.1, 1.0, size=src.size)
        rows += [src.ravel(), dst.ravel()]
        cols += [dst.ravel(), src.ravel()]
        vals += [w, w]
    rows, cols, vals = (np.concatenate(v) for v in (rows, cols, vals))
    off = scipy.sparse.csr_matrix((vals, (rows, cols)), shape=(n ** 3, n ** 3))
    diag = -np.asarray(off.sum(axis=1)).ravel() + rng.uniform(0.05, 0.5, size=n ** 3)
    A = (off + scipy.sparse.diags(diag)).tocsr()
    A.sort_indices()
    return A
