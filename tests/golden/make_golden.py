"""Generate tests/golden/*.npz from the REAL reference (run in the authoring
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py

Imports ju-liu/krylov from /root/reference/src through the two-symbol NumPy-2
shim of SURVEY.md section 8c (the reference is not modified), runs it on the
seeded inputs of tests/cases.py and stores its outputs.  The oracle (oracle/)
and the CUDA product are both checked against these files.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
sys.dont_write_bytecode = True

# --- shim: symbols removed in NumPy 2 that the reference calls
np.find_common_type = lambda a, s: np.result_type(*a, *s)  # _helpers.py:42
np.Infinity = np.inf  # utils.py:245
sys.path.insert(0, "/root/reference/src")
import krylov as ref  # noqa: E402

import cases  # noqa: E402


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):  # gmres.py:201-205 prints
        return fn(*a, **kw)


def pack(prefix, out, sol, info):
    out[prefix + "_success"] = np.array(info.success)
    out[prefix + "_solnone"] = np.array(sol is None)
    out[prefix + "_numsteps"] = np.array(info.numsteps)
    out[prefix + "_resnorms"] = np.asarray(info.resnorms, dtype=float)
    out[prefix + "_xk"] = np.asarray(info.xk)


def solver_cases():
    out = {}
    for name, (solver, A, b, kw) in cases.solver_cases().items():
        pack(name, out, *quiet(getattr(ref, solver), A, b, **kw))
    # return_arnoldi (cg.py:141-149, 220-232)
    _, A, b, _ = cases.solver_cases()["p2d32_cg"]
    _, info = quiet(ref.cg, A, b, tol=1e-10, maxiter=40, return_arnoldi=True)
    out["p2d32_cg_arn_H"] = info.arnoldi[1]
    out["p2d32_cg_arn_V5"] = np.asarray(info.arnoldi[0][5])
    # restarted GMRES(10) written as the user loop of SURVEY.md (no restart kwarg)
    _, A, b, _ = cases.solver_cases()["cd10_gmres_mgs"]
    x = np.zeros_like(b)
    hist = []
    for _cyc in range(4):
        _, info = quiet(ref.gmres, A, b, x0=x, tol=0.0, atol=0.0, maxiter=10)
        hist.append(np.asarray(info.resnorms))
        x = info.xk
    out["cd10_gmres_restart10_hist"] = np.stack(hist)
    out["cd10_gmres_restart10_xk"] = x
    return out


def arnoldi_cases():
    out = {}
    A, As, v = cases.arnoldi_inputs()
    inner = lambda x, y: np.dot(x.conj(), y)
    for nre in (1, 2):
        arn = ref.ArnoldiMGS(A, v.copy(), num_reorthos=nre, inner=inner)
        H = np.zeros((21, 20))
        for k in range(20):
            _, h = next(arn)
            H[: k + 2, k] = h
        out[f"mgs{nre}_H"] = H
        out[f"mgs{nre}_V"] = np.column_stack(arn.V)
    arn = ref.ArnoldiHouseholder(A, v.copy())
    H = np.zeros((21, 20))
    for k in range(20):
        _, h = next(arn)
        H[: k + 2, k] = h
    out["house_H"] = H
    out["house_V"] = np.column_stack(arn.V)
    lan = ref.ArnoldiLanczos(As, v.copy(), inner=inner)
    T = []
    Vs = [lan.v.copy()]
    for k in range(20):
        vv, h, _ = next(lan)
        T.append(h.copy())
        Vs.append(vv.copy())
    out["lanczos_h"] = np.array(T)
    out["lanczos_V"] = np.column_stack(Vs)
    return out


def small_cases():
    out = {}
    pairs = np.array([(f, g) for f in cases.GIVENS_F for g in cases.GIVENS_F])
    cs = []
    for f, g in pairs:
        G, r = ref.givens(np.array([f, g]))
        cs.append([G[0, 0], G[0, 1], r[0]])
    out["givens_fg"] = pairs
    out["givens_csr"] = np.array(cs)
    G, r = ref.givens(np.array([[1.0, 0.0, 3.0], [2.0, 5.0, -4.0]]))
    out["givens_block_G"] = G
    out["givens_block_r"] = r
    for i, x in enumerate(cases.householder_inputs()):
        H = ref.Householder(x.copy())
        out[f"house{i}_v"] = H.v
        out[f"house{i}_abx"] = np.array([H.alpha, H.beta, H.xnorm], dtype=float)
        out[f"house{i}_Hx"] = H @ x
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "solvers.npz"), **solver_cases())
    np.savez_compressed(os.path.join(HERE, "arnoldi.npz"), **arnoldi_cases())
    np.savez_compressed(os.path.join(HERE, "small.npz"), **small_cases())
    for f in ("solvers.npz", "arnoldi.npz", "small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
