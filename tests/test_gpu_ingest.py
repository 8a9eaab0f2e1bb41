"""-m gpu: input formats -- every way a matrix can reach the device gives the same
CSR arrays as SciPy's conversion and the same product, bit for bit."""
import numpy as np
import pytest
import scipy.io
import scipy.sparse
import torch

import krylov_b200 as kb
from krylov_b200 import stencils as st

pytestmark = pytest.mark.gpu


def _rand_sparse(n, m, density, seed, ints=False):
    r = np.random.default_rng(seed)
    A = scipy.sparse.random(n, m, density=density, random_state=r, format="coo")
    if ints:
        A.data = r.integers(-8, 9, A.nnz).astype(float)
    return A


def _same_csr(M, S):
    S = scipy.sparse.csr_matrix(S)
    S.sum_duplicates()
    S.sort_indices()
    got = M.to_scipy()
    np.testing.assert_array_equal(got.indptr, S.indptr)
    np.testing.assert_array_equal(got.indices, S.indices)
    np.testing.assert_array_equal(got.data, S.data)


@pytest.mark.parametrize("fmt", ["csr", "csc", "coo", "dia", "bsr", "lil", "dok"])
def test_scipy_formats(fmt):
    A = st.convection_diffusion3d(7)
    M = kb.CsrMatrix.from_scipy(A.asformat(fmt))
    _same_csr(M, A)
    x = np.random.default_rng(1).standard_normal(A.shape[1])
    np.testing.assert_array_equal(M @ x, A.tocsr() @ x)
    # and straight into a solver
    b = A @ np.ones(A.shape[0])
    sol, info = kb.gmres(A.asformat(fmt), b, tol=1e-10, maxiter=200)
    assert info.success and np.linalg.norm(sol - 1.0) <= 1e-8 * np.sqrt(A.shape[0])


def test_from_coo_on_device_with_duplicates_and_rectangular():
    A = _rand_sparse(301, 257, 0.02, 3, ints=True)
    rows = np.concatenate([A.row, A.row[:50]])  # 50 duplicated entries
    cols = np.concatenate([A.col, A.col[:50]])
    vals = np.concatenate([A.data, A.data[:50]])
    ref = scipy.sparse.coo_matrix((vals, (rows, cols)), shape=A.shape).tocsr()
    for conv in (lambda a: a, lambda a: torch.from_numpy(a).cuda()):
        M = kb.CsrMatrix.from_coo(conv(rows), conv(cols), conv(vals), A.shape)
        _same_csr(M, ref)
    x = np.random.default_rng(2).standard_normal((257, 3))
    np.testing.assert_array_equal(M @ x, ref @ x)
    with pytest.raises(ValueError):
        kb.CsrMatrix.from_coo([0, 5], [0, 0], [1.0, 1.0], (3, 3))


def test_torch_sparse_layouts():
    A = _rand_sparse(200, 200, 0.03, 4).tocsr()
    A.sort_indices()
    dense = torch.from_numpy(A.toarray())
    for T in (dense.to_sparse_csr(), dense.to_sparse_coo(), dense.to_sparse_csr().cuda(),
              dense.to_sparse_coo().cuda()):
        _same_csr(kb.CsrMatrix.from_torch(T), A)


@pytest.mark.parametrize("symmetric", [False, True])
def test_matrix_market_and_npz_files(tmp_path, symmetric):
    A = st.poisson2d(12) if symmetric else st.convection_diffusion3d(5)
    p = tmp_path / "a.mtx"
    scipy.io.mmwrite(str(p), A, symmetry="symmetric" if symmetric else "general")
    M = kb.CsrMatrix.from_file(p)
    _same_csr(M, A)
    q = tmp_path / "a.npz"
    scipy.sparse.save_npz(str(q), A.tocsc())
    _same_csr(kb.CsrMatrix.from_file(q), A)
    b = A @ np.arange(1.0, A.shape[0] + 1)
    sol, info = (kb.cg if symmetric else kb.gmres)(M, b, tol=1e-12, maxiter=300)
    assert info.success
    np.testing.assert_allclose(sol, np.arange(1.0, A.shape[0] + 1), rtol=1e-8)
