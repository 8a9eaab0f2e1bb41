"""Pins oracle/krylov_oracle_extra.py (bicgstab, cgs, bicg, qmr, cgne, cgnr, cgr, gcr, chebyshev)
to the real reference: outputs of the unmodified reference on tests/cases_extra.py
(tests/golden/extra.npz, written by make_golden_extra.py) and the reference's own known-answer
vectors (reference tests/test_cgr.py:18-36, test_gcr.py:18-36, test_chebyshev.py:10-25)."""
import os

import numpy as np
import pytest

import cases_extra
from oracle import krylov_oracle_extra as orx

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "extra.npz"))
CASES = cases_extra.extra_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_extra_solver_matches_reference(name):
    solver, A, b, kw = CASES[name]
    sol, info = getattr(orx, solver)(A, b, **kw)
    assert info.numsteps == int(G[name + "_numsteps"])
    assert bool(info.success) == bool(G[name + "_success"])
    assert (sol is None) == bool(G[name + "_solnone"])
    ref_res = G[name + "_resnorms"]
    res = np.asarray(info.resnorms, dtype=float)
    assert res.shape == ref_res.shape
    np.testing.assert_allclose(res, ref_res, rtol=1e-9, atol=1e-300)
    ref_x = G[name + "_xk"]
    assert np.linalg.norm(np.asarray(info.xk) - ref_x) <= 1e-10 * max(np.linalg.norm(ref_x), 1e-300)
    if sol is not None:
        assert np.may_share_memory(sol, info.xk)


def _spd_dense(n=5):  # reference tests/linear_problems.py:5-10
    a = np.linspace(1.0, 2.0, n)
    a[-1] = 1e-2
    return np.diag(a), np.ones(n)


_CR_REF = [2.23606797749979, 1.06995241076096, 0.9872554076721121, 0.9709417754217194,
           0.7291872161218861, 6.377745716588144e-16]


@pytest.mark.parametrize("solver", ["cgr", "gcr"])
def test_reference_known_answers_cr(solver):
    A, b = _spd_dense()
    _, info = getattr(orx, solver)(A, b, maxiter=5)
    ref = np.array(_CR_REF)
    assert np.all(np.abs(np.asarray(info.resnorms) - ref) < 1.0e-13 * (1.0 + np.abs(ref)))


def test_reference_known_answers_chebyshev():
    A, b = _spd_dense()
    _, info = orx.chebyshev(A, b, (1.0e-2, 1.75), tol=1.0e-5, maxiter=5)
    ref = np.array([2.23606797749979, 1.626826691029081, 1.744954212067044, 1.7113839589143471,
                    1.6298632096913288, 1.4593167230617032])
    assert np.all(np.abs(np.asarray(info.resnorms) - ref) < 1.0e-12 * (1.0 + ref))


def test_callback_counts_like_reference():
    """reference tests: callback runs once before the loop and once per step."""
    A, b = CASES["cd8_bicgstab"][1:3]
    for solver in ("bicgstab", "cgs", "bicg", "qmr", "gcr"):
        cnt = [0]
        sol, info = getattr(orx, solver)(A, b, tol=1e-7, maxiter=200,
                                         callback=lambda x, r: cnt.__setitem__(0, cnt[0] + 1))
        assert info.success
        assert cnt[0] == info.numsteps + 1


def test_reference_known_answers_symmlq():
    """reference tests/test_symmlq.py:16-32"""
    a = np.linspace(1.0, 2.0, 5)
    a[-1] = -1.0
    _, info = orx.symmlq(np.diag(a), np.ones(5), maxiter=10)
    ref = np.array([2.23606797749979, 0.9823441352194251, 0.5792270481089666, 0.2307060320183781,
                    0.16833036914998076, 2.5918417478740246e-15])
    assert np.all(np.abs(np.asarray(info.resnorms) - ref) < 1.0e-13 * (1.0 + np.abs(ref)))
