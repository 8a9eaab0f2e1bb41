"""Generate tests/golden/complex.npz: the REAL reference (ju-liu/krylov from /root/reference/src,
two-symbol NumPy-2 shim of SURVEY.md 8c) on the complex Hermitian cases of tests/cases_complex.py.

    python tests/golden/make_golden_complex.py

Run in the authoring container only (/root/reference does not exist on the GPU box)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.dont_write_bytecode = True
np.find_common_type = lambda a, s: np.result_type(*a, *s)  # _helpers.py:42
np.Infinity = np.inf  # utils.py:245
sys.path.insert(0, "/root/reference/src")
import krylov as ref  # noqa: E402

import cases_complex  # noqa: E402

out = {}
for name, (solver, A, b, kw) in cases_complex.cases().items():
    sol, info = getattr(ref, solver)(A, b, **kw)
    out[name + "_numsteps"] = np.int64(info.numsteps)
    out[name + "_success"] = np.bool_(info.success)
    out[name + "_resnorms"] = np.asarray(info.resnorms, dtype=float)
    out[name + "_xk"] = np.asarray(info.xk)
    print(name, solver, info.numsteps, info.success, np.asarray(info.resnorms)[-1])
np.savez_compressed(os.path.join(HERE, "complex.npz"), **out)
