"""Generate tests/golden/extra.npz from the REAL reference (authoring container only).

    python tests/golden/make_golden_extra.py

Runs ju-liu/krylov's bicgstab / cgs / bicg / qmr / cgne / cgnr / cgr / gcr / chebyshev (unmodified,
from /root/reference/src through the NumPy-2 shim of SURVEY.md 8c) on tests/cases_extra.py."""
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

import cases_extra  # noqa: E402

out = {}
for name, (solver, A, b, kw) in cases_extra.extra_cases().items():
    sol, info = getattr(ref, solver)(A, b, **kw)
    out[name + "_success"] = np.array(info.success)
    out[name + "_solnone"] = np.array(sol is None)
    out[name + "_numsteps"] = np.array(info.numsteps)
    out[name + "_resnorms"] = np.asarray(info.resnorms, dtype=float)
    out[name + "_xk"] = np.asarray(info.xk)
    print(f"{name:24s} steps {info.numsteps:4d} success {info.success} last {np.max(info.resnorms[-1]):.3e}")
np.savez_compressed(os.path.join(HERE, "extra.npz"), **out)
print("wrote", os.path.join(HERE, "extra.npz"))
