"""CPU checks of the boundary: the shared library loads without a GPU, exports
every symbol include/krylov_b200.h declares, validates arguments, and the
product refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "krylov_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(kb_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import krylov_b200._lib as L

    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L.lib, n), f"{n} declared in the header but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature"
    for n in L.SIGNATURES:
        assert n in names, f"{n} bound in Python but not declared in the header"
    assert L.lib.kb_version() == 100


def test_argument_validation_without_gpu():
    import krylov_b200._lib as L

    # pure host-side validation paths: no CUDA call is reached
    assert L.lib.kb_ws_create(None, 1) == -1
    assert "null handle" in L.last_error()
    h = ctypes.c_void_p()
    assert L.lib.kb_ws_create(ctypes.byref(h), 0) == -1
    assert "max_k" in L.last_error()
    assert L.lib.kb_csr_create(ctypes.byref(h), -1, 1, 0, None, None, None, 0, None) == -1
    assert L.lib.kb_tune(99, 0) == -1
    # entry points added for utils / blocked right-hand sides: null handles are refused, not dereferenced
    assert L.lib.kb_block_gram(None, 8, 2, 2, None, 2, None, 2, None, 2, None, 0, 0, None) == -1
    assert "null workspace" in L.last_error()
    assert L.lib.kb_block_apply(None, 8, 2, 2, None, 2, None, 2, None, 2, None, 2, 0, None) == -1
    assert L.lib.kb_house_make2(None, 8, 0, None, None, None, None, 1, None) == -1
    yes = ctypes.c_int(7)
    assert L.lib.kb_spmm_is_lines(None, 16, None, ctypes.byref(yes)) == -1
    for key in (16, 17, 18, 19):  # tunables of the line-marching SpMM exist (and are restored)
        assert L.lib.kb_tune(key, {16: 1, 17: 0, 18: 0, 19: 1}[key]) == 0
    with pytest.raises(L.KrylovB200Error):
        L.check(-1)


def test_no_cpu_fallback():
    import torch

    import krylov_b200 as kb

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = np.diag([1.0, 2.0, 3.0])
    b = np.ones(3)
    for fn in (kb.cg, kb.minres, kb.gmres):
        with pytest.raises(kb.KrylovB200Error):
            fn(A, b)
    with pytest.raises(kb.KrylovB200Error):
        kb.givens(np.array([1.0, 2.0]))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "krylov_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f


def test_stencil_generators_match_definitions():
    from krylov_b200 import stencils as st

    A = st.poisson2d(4).toarray()
    assert A.shape == (16, 16) and np.all(np.diag(A) == 4.0)
    assert A[0, 1] == -1 and A[3, 4] == 0 and A[0, 4] == -1  # no wrap across rows
    assert st.stencil5_csr(256, 256)[1].size == 326656  # C1 nnz
    B = st.convection_diffusion3d(3).toarray()
    assert B[1, 0] == -1.5 and B[0, 1] == -0.5 and B[3, 0] == -1.25 and B[9, 0] == -1.125
    # slab generator == rows of the full matrix
    full = st.to_scipy(st.stencil7_csr(4, 3, 5)).toarray()
    slab = st.to_scipy(st.stencil7_csr(4, 3, 5, z_lo=1, z_hi=3), n_cols=60).toarray()
    np.testing.assert_array_equal(slab, full[12:36])
    n = 6
    C = st.shifted_laplace3d(n).toarray()
    ev = np.linalg.eigvalsh(C)
    assert (ev < 0).sum() == 1  # mild shift: exactly one negative eigenvalue
