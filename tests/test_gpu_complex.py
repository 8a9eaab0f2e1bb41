"""-m gpu: complex128 cg / minres on Hermitian matrices through the real-equivalent embedding
(krylov_b200/_complex.py) against the outputs of the unmodified reference
(tests/golden/complex.npz: the reference's own `hpd` / `hermitian_indefinite` problems and seeded
sparse Hermitian matrices, single and blocked right-hand sides, x0)."""
import os

import numpy as np
import pytest
import torch

import cases_complex
import krylov_b200 as kb

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "complex.npz"))
CASES = cases_complex.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_complex_hermitian_matches_reference(name):
    solver, A, b, kw = CASES[name]
    sol, info = getattr(kb, solver)(A, b, **kw)
    steps = int(G[name + "_numsteps"])
    assert abs(info.numsteps - steps) <= max(1, int(0.02 * steps))
    assert bool(info.success) == bool(G[name + "_success"])
    assert sol is info.xk and np.iscomplexobj(info.xk) and info.xk.shape == G[name + "_xk"].shape
    res, ref = np.asarray(info.resnorms, dtype=float), G[name + "_resnorms"]
    m = min(len(res), len(ref))
    live = ref[:m] / ref[0] >= 1e-6
    if name == "sp_hind_minres":
        live[60:] = False  # see tests/test_complex_cpu.py
    bar = 1e-8 * np.maximum.accumulate(ref[:m], axis=0)
    assert np.all((np.abs(res[:m] - ref[:m]) <= bar)[live])
    assert np.linalg.norm(info.xk - G[name + "_xk"]) <= 1e-7 * np.linalg.norm(G[name + "_xk"])
    r = b - A @ info.xk
    assert np.all(np.linalg.norm(r, axis=0) <= 10 * kw["tol"] * np.linalg.norm(b, axis=0) + 1e-13)


def test_complex_interface():
    A, b = cases_complex.ref_hpd()
    seen = []
    sol, info = kb.cg(A, b, tol=1e-12, callback=lambda x, r: seen.append((x.copy(), r.copy())))
    assert len(seen) == info.numsteps + 1 and all(np.iscomplexobj(x) and x.shape == (5,) for x, _ in seen)
    np.testing.assert_allclose(seen[-1][0], sol, rtol=1e-12)
    # torch in -> torch out, on the caller's device
    st, it = kb.minres(torch.from_numpy(A), torch.from_numpy(b).cuda(), tol=1e-12)
    assert isinstance(st, torch.Tensor) and st.is_cuda and st.is_complex()
    np.testing.assert_allclose(st.cpu().numpy(), sol, rtol=1e-9)
    # a real matrix with a complex right-hand side
    Ar = np.diag(np.linspace(1.0, 2.0, 5))
    sr, ir = kb.cg(Ar, b * (1 + 2j), tol=1e-12)
    np.testing.assert_allclose(sr, np.linalg.solve(Ar, b * (1 + 2j)), rtol=1e-10)
    # what stays out of scope says so
    Au = A.copy()
    Au[0, 1] = 3.0j
    with pytest.raises(NotImplementedError, match="Hermitian"):
        kb.cg(Au, b)
    with pytest.raises(NotImplementedError):
        kb.gmres(A, b)
    with pytest.raises(NotImplementedError, match="inner"):
        kb.cg(A, b, inner=lambda x, y: np.vdot(x, y))
