"""Host logic of krylov_b200._alg.DevScalar (device-resident per-column scalars of the
short-recurrence solvers) on CPU: a test double of ``ops.scalar_op`` evaluates each launch with
the kernel's semantics (csrc/kb_loops.cu: kb_scalar_op_kernel), so that operator overloading,
NumPy ufunc dispatch, immediates / arrays as operands and the lazy ``nz()`` are checked against
plain NumPy expressions bit for bit."""
import numpy as np
import pytest
import torch

from krylov_b200._alg import DevScalar, nz, to_host


class _Ops:
    def __init__(self):
        self.launches = 0
        self.codes = []

    def scalar_op(self, code, a, b, sa, sb, out):
        self.launches += 1
        self.codes.append(code)
        A = a.numpy() if a is not None else np.full(out.numel(), sa)
        B = b.numpy() if b is not None else np.full(out.numel(), sb)
        with np.errstate(all="ignore"):
            r = {0: lambda: A + B, 1: lambda: A - B, 2: lambda: A * B, 3: lambda: A / B,
                 4: lambda: np.sqrt(A), 5: lambda: np.abs(A), 6: lambda: -A,
                 7: lambda: np.where(A != 0.0, A, B), 8: lambda: A,
                 9: lambda: A / np.where(B != 0.0, B, sb)}[code]()
        out.copy_(torch.from_numpy(np.asarray(r, dtype=np.float64)))


class _Prob:
    k, device = 3, None


class _Alg:
    def __init__(self):
        self.prob, self.ops = _Prob(), _Ops()

    def coef(self, a):
        if isinstance(a, DevScalar):
            return a.t
        return torch.from_numpy(np.array(np.broadcast_to(np.asarray(a, dtype=np.float64), (3,))))


def ds(alg, v):
    return DevScalar(alg, torch.from_numpy(np.array(v, dtype=np.float64)))


def test_expressions_equal_numpy_bit_for_bit():
    alg = _Alg()
    rho, alpha, rho_old, omega = ([1.5, -2.25, 3.0], [0.1, 0.2, 0.3], [2.0, 0.0, -1.0], [0.7, 0.0, 5.0])
    R, A, Ro, O = (np.array(v) for v in (rho, alpha, rho_old, omega))
    r, a, ro, o = (ds(alg, v) for v in (rho, alpha, rho_old, omega))
    beta = r * a / nz(ro * o)  # bicgstab.py:101
    np.testing.assert_array_equal(to_host(beta), R * A / np.where(Ro * O != 0, Ro * O, 1.0))
    theta = r / nz(a * np.abs(o))  # qmr.py:136
    np.testing.assert_array_equal(to_host(theta), R / np.where(A * np.abs(O) != 0, A * np.abs(O), 1.0))
    gamma = 1 / np.sqrt(1 + theta ** 2)  # qmr.py:137
    T = to_host(theta)
    np.testing.assert_array_equal(to_host(gamma), 1 / np.sqrt(1 + T ** 2))
    eta = -ds(alg, [1.0, 2.0, 3.0]) * ro * gamma ** 2 / nz(beta * a ** 2)  # qmr.py:138
    G, Bt = to_host(gamma), to_host(beta)
    np.testing.assert_array_equal(to_host(eta), -np.array([1.0, 2.0, 3.0]) * Ro * G ** 2
                                  / np.where(Bt * A ** 2 != 0, Bt * A ** 2, 1.0))
    # immediates, NumPy scalars and (k,) arrays on either side
    arr = np.array([10.0, 20.0, 30.0])
    for got, want in ((2.0 - r, 2.0 - R), (r - 2.0, R - 2.0), (np.float64(3.0) * r, 3.0 * R),
                      (arr * r, arr * R), (r / arr, R / arr), (arr / nz(ro, 1e-15), arr / np.where(Ro != 0, Ro, 1e-15)),
                      (abs(r), np.abs(R)), (np.negative(r), -R), (np.square(r), R * R),
                      (nz(ro) + 1.0, np.where(Ro != 0, Ro, 1.0) + 1.0)):
        np.testing.assert_array_equal(to_host(got), want)


def test_division_by_nz_is_one_launch_and_nz_alone_materialises():
    alg = _Alg()
    x, y = ds(alg, [1.0, 2.0, 3.0]), ds(alg, [0.0, 4.0, 0.0])
    n0 = alg.ops.launches
    q = x / nz(y)
    assert alg.ops.launches == n0 + 1 and alg.ops.codes[-1] == 9  # kb_scalar_op 9: A / nz(B)
    np.testing.assert_array_equal(to_host(q), [1.0, 0.5, 3.0])
    lazy = nz(y, 7.0)
    assert alg.ops.launches == n0 + 1  # nothing evaluated yet
    np.testing.assert_array_equal(alg.coef(lazy).numpy(), [7.0, 4.0, 7.0])
    assert alg.ops.launches == n0 + 2
    alg.coef(lazy)
    assert alg.ops.launches == n0 + 2  # cached
    # host values pass through the same helpers unchanged (user inner products keep host scalars)
    np.testing.assert_array_equal(nz(np.array([0.0, 2.0])), [1.0, 2.0])
    assert to_host(3.5) == 3.5


def test_unsupported_operations_fail_loudly():
    alg = _Alg()
    x = ds(alg, [1.0, 2.0, 3.0])
    with pytest.raises(TypeError):
        np.exp(x)
    with pytest.raises(TypeError):
        x ** 3
    with pytest.raises(TypeError):
        np.asarray(x)
