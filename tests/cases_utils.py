"""Seeded inputs of the ``krylov.utils`` parity cases (shared by the golden generator, the oracle
tests and the GPU tests).  The matrices of the reference's own tests (tests/test_utils.py:21,39-46,
tests/helpers.py:26-70, real ones only) come first, random tall-skinny blocks after them."""
import numpy as np
import scipy.linalg
import scipy.sparse


def weight_diag(n):
    """tests/helpers.py:104: B = diag(linspace(1, 5, n))."""
    return np.linspace(1.0, 5.0, n)


def numpy_inner(name, n):
    """The inner products of tests/helpers.py:103-110 as NumPy callables (reference / oracle)."""
    if name is None:
        return None
    if name == "euclid":
        return lambda x, y: np.dot(x.T.conj(), y)
    assert name == "weighted"
    B = np.diag(weight_diag(n))
    return lambda x, y: np.dot(x.T.conj(), np.dot(B, y))


def _rng(seed):
    return np.random.default_rng(seed)


def qr_cases():
    """name -> (X, inner_name, reorthos)"""
    mats = {
        "eye": np.eye(10, 5),
        "hilbert": scipy.linalg.hilbert(10)[:, :5],
        "rand7": _rng(1).standard_normal((200, 7)),
        "rand16": _rng(2).standard_normal((333, 16)),
        "rand20": _rng(3).standard_normal((257, 20)),  # two 16-column panels
        "zerocol": np.column_stack([_rng(4).standard_normal(50), np.zeros(50),
                                    _rng(5).standard_normal(50)]),
        "lastcols": np.eye(10)[:, -4:],  # zero pivots: LAPACK's sign convention
    }
    out = {}
    for mname, X in mats.items():
        out[f"qr_{mname}_lapack"] = (X, None, 1)
        for inner in ("euclid", "weighted"):
            for reorthos in (0, 1, 2):
                out[f"qr_{mname}_{inner}_{reorthos}"] = (X, inner, reorthos)
    return out


def angles_cases():
    """name -> (F, G, inner_name)"""
    E = np.eye(10)
    r = _rng(7)
    F6 = r.standard_normal((400, 6))
    G4 = r.standard_normal((400, 4))
    near = F6[:, :3] + 1e-9 * r.standard_normal((400, 3))  # tiny angles: the sine branch
    mixed = np.column_stack([F6[:, 0] + 1e-7 * r.standard_normal(400), G4[:, 0]])
    pairs = {
        "e1_e1": (E[:, :1], E[:, :1]),
        "e4_last4": (E[:, :4], E[:, -4:]),
        "e4_scaled": (E[:, :4], E[:, :4] @ np.diag([1.0, 1e1, 1e2, 1e3])),
        "e1_e4": (E[:, :1], E[:, :4]),
        "rand_6_4": (F6, G4),
        "rand_4_6": (G4, F6),
        "near": (F6, near),
        "mixed": (F6, mixed),
        "same": (F6, F6.copy()),
    }
    out = {}
    for pname, (F, G) in pairs.items():
        for inner in ("euclid", "weighted"):
            out[f"angles_{pname}_{inner}"] = (F, G, inner)
    return out


def _dense_matrices():
    a = np.linspace(1, 2, 10)
    spd = a.copy()
    spd[-1] = 1e-2
    indef = a.copy()
    indef[-1] = -1
    ns = np.arange(1, 11, dtype=float)
    ns[-1] = -1e1
    N = np.diag(ns)
    N[0, -1] = 1e1
    return {"spd": np.diag(spd), "indef": np.diag(indef), "nonsymm": N}


def hegedus_cases():
    """name -> (A, b, x0, M, Ml, inner_name)"""
    m = np.arange(1, 11, dtype=float)
    m[-1] = 1.0
    D = np.diag(m)
    out = {}
    for aname, A in _dense_matrices().items():
        b = A @ np.ones((10, 1))
        for xname, x0 in (("zero", np.zeros((10, 1))), ("lin", np.linspace(1, 5, 10).reshape(10, 1)),
                          ("ones", np.ones((10, 1)))):
            for pname, (M, Ml) in (("none", (None, None)), ("M", (D, None)), ("Ml", (None, D)),
                                   ("both", (D, D))):
                for inner in ("euclid", "weighted"):
                    out[f"heg_{aname}_{xname}_{pname}_{inner}"] = (A, b, x0, M, Ml, inner)
    # sparse operator, 1-D vectors (scalar inner products)
    n = 24
    T = scipy.sparse.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    A = (scipy.sparse.kron(scipy.sparse.eye(n), T) + scipy.sparse.kron(T, scipy.sparse.eye(n))).tocsr()
    r = _rng(11)
    xs = r.standard_normal(n * n)
    b = A @ xs
    x0 = xs + 0.3 * r.standard_normal(n * n)
    J = scipy.sparse.diags(1.0 / A.diagonal()).tocsr()
    out["heg_poisson_1d_none_euclid"] = (A, b, x0, None, None, "euclid")
    out["heg_poisson_1d_M_euclid"] = (A, b, x0, J, None, "euclid")
    out["heg_poisson_1d_Ml_weighted"] = (A, b, 3.0 * x0, None, J, "weighted")
    return out
