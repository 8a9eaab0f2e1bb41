"""Generate tests/golden/utils.npz from the REAL reference (authoring container only).

    python tests/golden/make_golden_utils.py

Runs ju-liu/krylov's utils.qr / utils.angles / utils.hegedus (unmodified, from /root/reference/src
through the NumPy-2 shim of SURVEY.md 8c) on tests/cases_utils.py, plus the small host helpers."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.dont_write_bytecode = True
np.find_common_type = lambda a, s: np.result_type(*a, *s)  # _helpers.py:42
np.Infinity = np.inf  # utils.py:245
sys.path.insert(0, "/root/reference/src")
import krylov as ref  # noqa: E402

import cases_utils as cu  # noqa: E402

out = {}
for name, (X, inner, reorthos) in cu.qr_cases().items():
    res = ref.utils.qr(X, inner=cu.numpy_inner(inner, X.shape[0]), reorthos=reorthos)
    if inner is None:
        # utils.py:24 asks np.linalg.qr for mode="economic", which NumPy (>= 1.8, deprecated) answers
        # with ONE array: LAPACK geqrf's packed output (R in the upper triangle, reflectors below).
        # Stored as the reference's result; its upper triangle pins R of the (Q, R) the docstring
        # promises (what oracle and product return: np.linalg.qr(mode="reduced"), same geqrf).
        assert isinstance(res, np.ndarray) and res.shape == X.shape
        out[name + "_packed"] = res
        Q, R = np.linalg.qr(X, mode="reduced")
        assert np.array_equal(np.triu(res[: R.shape[0]]), R)
    else:
        Q, R = res
    out[name + "_Q"], out[name + "_R"] = Q, R
for name, (F, G, inner) in cu.angles_cases().items():
    theta, U, V = ref.utils.angles(F, G, inner=cu.numpy_inner(inner, F.shape[0]), compute_vectors=True)
    theta2 = ref.utils.angles(F, G, inner=cu.numpy_inner(inner, F.shape[0]))
    assert np.array_equal(theta, theta2)
    out[name + "_theta"], out[name + "_U"], out[name + "_V"] = theta, U, V
    print(f"{name:28s} theta {np.array2string(theta, precision=3, max_line_width=200)}")
for name, (A, b, x0, M, Ml, inner) in cu.hegedus_cases().items():
    out[name + "_x0new"] = np.asarray(ref.utils.hegedus(A, b, x0, M, Ml, cu.numpy_inner(inner, b.shape[0])))
out["strakos_5"] = ref.utils.strakos(5)
out["strakos_12"] = ref.utils.strakos(12, l_min=0.5, l_max=7.0, rho=0.8)
roots = np.array([1.0, 2.0, 1e8, 1e8 + 1e-3])
p = ref.utils.NormalizedRootsPolynomial(roots)
pts = np.linspace(0.0, 3.0, 17)
out["nrp_roots"], out["nrp_pts"], out["nrp_vals"] = roots, pts, p(pts)
out["nrp_cand"] = np.sort_complex(p.minmax_candidates())
np.savez_compressed(os.path.join(HERE, "utils.npz"), **out)
print("wrote", os.path.join(HERE, "utils.npz"), len(out), "arrays")
