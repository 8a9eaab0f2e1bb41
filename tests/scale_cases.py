"""Seeded inputs of the BASELINE-scale parity cases (tests/golden/scale.npz).

One definition shared by the fixture generator (tests/golden/make_golden_scale.py, real
reference, authoring container), the -m gpu tests and bench.py's `parity` block, so that
every side solves the same system: SURVEY.md 8d recipe ``x* = default_rng(seed)
.standard_normal(shape)``, ``b = A x*`` (SciPy's csr_matvec; the device product is
bit-identical to it, tests/test_gpu_kernels.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name: (solver, fixed steps, solver kwargs, grid edge, matrix kind, right-hand sides)
CASES = {
    "c2_minres_128": ("minres", 30, {}, 128, "shifted", 1),
    "c3_gmres_mgs_128": ("gmres", 30, {"ortho": "mgs"}, 128, "convdiff", 1),
    "c3_gmres_mgs2_128": ("gmres", 30, {"ortho": "mgs2"}, 128, "convdiff", 1),
    "c3_gmres_householder_128": ("gmres", 30, {"ortho": "householder"}, 128, "convdiff", 1),
    "c4_cg_k16_128": ("cg", 30, {}, 128, "poisson", 16),
    "c5_cg_256": ("cg", 30, {}, 256, "poisson", 1),
    "c5_cg_512": ("cg", 30, {}, 512, "poisson", 1),
}


def _load_stencils():
    """krylov_b200/stencils.py by path: pure NumPy, and importing the package would load the
    CUDA library (the CPU arms must not map it)."""
    import importlib.util

    p = os.path.join(ROOT, "krylov_b200", "stencils.py")
    spec = importlib.util.spec_from_file_location("_kb_stencils", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def matrix_params(kind, N):
    """(coeffs, shift) of the 7-point operator of a case."""
    st = _load_stencils()
    if kind == "poisson":
        return st.STENCIL_POISSON, 0.0
    if kind == "shifted":
        return st.STENCIL_POISSON, float(st.mild_shift(N))
    if kind == "convdiff":
        return st.convdiff_coeffs(), 0.0
    raise KeyError(kind)


def xstar(name, seed=0):
    """The seeded exact solution x* (n,) or (n, k)."""
    _, _, _, N, _, k = CASES[name]
    n = N ** 3
    return np.random.default_rng(seed).standard_normal(n if k == 1 else (n, k))


def host_matrix(kind, N, slab=16, threads=None):
    """SciPy CSR of the N^3 7-point operator, assembled in z-slabs so that the 512^3 matrix
    (11.8 GB) is built without (n, 7)-shaped int64 temporaries of the whole grid; the slabs are
    independent and NumPy releases the GIL in the large array operations, so they are built by a
    small thread pool."""
    import concurrent.futures as cf

    import scipy.sparse

    st = _load_stencils()
    coeffs, shift = matrix_params(kind, N)
    n = N ** 3
    nnz = 7 * n - 6 * N * N
    rowptr = np.empty(n + 1, dtype=np.int32)
    cols = np.empty(nnz, dtype=np.int32)
    vals = np.empty(nnz, dtype=np.float64)
    rowptr[0] = 0

    # nonzeros per plane: 7 P minus the missing neighbours (x: 2 N per plane, y: 2 N per plane,
    # z: P on the first and on the last plane)
    P = N * N
    per_plane = np.full(N, 7 * P - 4 * N, dtype=np.int64)
    per_plane[0] -= P
    per_plane[-1] -= P
    start = np.concatenate([[0], np.cumsum(per_plane)])

    def work(z0):
        z1 = min(N, z0 + slab)
        rp, ci, va = st.stencil7_csr(N, N, N, coeffs, shift, z0, z1)
        pos = int(start[z0])
        assert ci.size == int(start[z1]) - pos
        rowptr[z0 * P + 1:z1 * P + 1] = rp[1:].astype(np.int64) + pos
        cols[pos:pos + ci.size] = ci
        vals[pos:pos + va.size] = va

    threads = threads or min(8, os.cpu_count() or 1)
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(0, N, slab)))
    assert int(start[-1]) == nnz and rowptr[-1] == nnz
    A = scipy.sparse.csr_matrix((vals, cols, rowptr), shape=(n, n), copy=False)
    A.has_canonical_format = True
    return A


def build(name):
    """(A as SciPy CSR, b) on the host."""
    _, _, _, N, kind, _ = CASES[name]
    A = host_matrix(kind, N)
    return A, A @ xstar(name)


def sample_index(n, count=256):
    """Fixed pseudo-random entry indices of the final iterate kept in the fixture."""
    return np.sort(np.random.default_rng(12345).choice(n, size=count, replace=False))
