"""Generate tests/golden/scale.npz: the REAL reference at BASELINE.json's sizes.

    python tests/golden/make_golden_scale.py [case ...]

Run in the authoring container only (/root/reference does not exist on the GPU
box).  Imports ju-liu/krylov from /root/reference/src through the two-symbol
NumPy-2 shim of SURVEY.md 8c and runs a FIXED number of steps (tol = atol = 0:
the stopping test of cg.py:156 / minres.py:171 / gmres.py:182 never fires) of

    c2_minres_128    minres, shifted 3-D Laplacian 128^3 (BASELINE configs[1])
    c3_gmres_<o>_128 gmres, convection-diffusion 128^3, o in mgs, mgs2, householder
                     (configs[2] at half the edge: 256^3 keeps 51 basis vectors of
                     134 MB in the reference's Python lists -- 30 steps at 128^3
                     pin the same arithmetic)
    c4_cg_k16_128    blocked cg, 16 right-hand sides, 3-D Poisson 128^3 (configs[3])
    c5_cg_256        cg, 3-D Poisson 256^3
    c5_cg_512        cg, 3-D Poisson 512^3 (configs[4], the benchmarked workload)

on the seeded inputs of tests/scale_cases.py and stores the residual history,
norms of the final iterate and a fixed sample of its entries.  The file is a few
kB; tests/test_gpu_fullsize.py compares the CUDA path with it at the north-star
tolerances, bench.py emits the same comparison in its JSON line.
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.dont_write_bytecode = True

np.find_common_type = lambda a, s: np.result_type(*a, *s)  # _helpers.py:42
np.Infinity = np.inf  # utils.py:245
sys.path.insert(0, "/root/reference/src")
import krylov as ref  # noqa: E402

import scale_cases  # noqa: E402

OUT = os.path.join(HERE, "scale.npz")


def main():
    want = sys.argv[1:] or list(scale_cases.CASES)
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in want:
        solver, steps, kw = scale_cases.CASES[name][:3]
        t0 = time.time()
        A, b = scale_cases.build(name)
        t1 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):  # gmres.py:201-205 prints
            sol, info = getattr(ref, solver)(A, b, tol=0.0, atol=0.0, maxiter=steps, **kw)
        t2 = time.time()
        assert sol is None and info.numsteps == steps
        x = np.asarray(info.xk)
        out[name + "_resnorms"] = np.asarray(info.resnorms, dtype=float)
        out[name + "_xnorm2"] = np.sqrt(np.sum(x * x, axis=0))
        out[name + "_xsumabs"] = np.sum(np.abs(x), axis=0)
        idx = scale_cases.sample_index(x.shape[0])
        out[name + "_xsample"] = x[idx]
        out[name + "_seconds"] = np.array([t1 - t0, t2 - t1])
        print(f"{name}: build {t1-t0:.1f} s, {steps} steps {t2-t1:.1f} s "
              f"({steps/(t2-t1):.3f} it/s), resnorm {np.ravel(info.resnorms[0])[0]:.6e} -> "
              f"{np.ravel(info.resnorms[-1])[0]:.6e}", flush=True)
        del A, b, sol, info, x
        np.savez_compressed(OUT, **out)


if __name__ == "__main__":
    main()
