"""ctypes binding of libkrylov_b200.so (the C ABI in include/krylov_b200.h).

There is no CPU fallback: if the shared library is missing, importing the
package raises; if no sm_100 device is present, the first compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libkrylov_b200.so")


class KrylovB200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C krylov_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "krylov_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_int64
f64 = C.c_double


class MinresState(C.Structure):
    _fields_ = [
        ("alpha", vp), ("ww", vp), ("h2prev", vp), ("g0", vp), ("g1", vp), ("y0", vp),
        ("coefs", vp), ("crit", vp), ("hist", vp), ("stop_at", vp), ("flags", vp),
    ]


class CgState(C.Structure):
    _fields_ = [
        ("A", vp), ("n", i64), ("k", i32), ("x", vp), ("r", vp), ("p", vp), ("Ap", vp),
        ("slots", vp), ("crit", vp), ("hist", vp), ("stop_at", vp), ("p2", vp), ("pcur", i32),
        ("masks_ext", vp), ("n_ext", i64), ("own_lo", i64), ("r_push_lo", vp), ("r_push_hi", vp),
    ]


class GmresState(C.Structure):
    _fields_ = [
        ("dots", vp), ("ww", vp), ("num_reorthos", i32), ("maxiter", i32), ("R", vp),
        ("Gc", vp), ("Gs", vp), ("y", vp), ("hlast", vp), ("crit", vp), ("hist", vp),
        ("stop_at", vp), ("flags", vp), ("have_h", i32),
    ]


class MinresRunState(C.Structure):
    _fields_ = [
        ("A", vp), ("n", i64), ("k", i32), ("V", vp * 2), ("W", vp * 2), ("Av", vp), ("yk", vp),
        ("st", MinresState),
    ]


class GmresCycleState(C.Structure):
    _fields_ = [
        ("A", vp), ("n", i64), ("k", i32), ("Vbuf", vp), ("vstride", i64), ("w", vp), ("dots", vp),
        ("ww", vp), ("hlast", vp), ("st", GmresState),
    ]


# name -> argtypes (restype is always int).  Kept in one table so the CPU test
# can check that every symbol declared in the header is exported.
SIGNATURES = {
    "kb_version": [],
    "kb_last_error": [C.c_char_p, C.c_size_t],
    "kb_tune": [i32, i32],
    "kb_device_info": [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
    "kb_ws_create": [C.POINTER(vp), i32],
    "kb_ws_destroy": [vp],
    "kb_ws_set_gate": [vp, vp, i32],
    "kb_comm_create": [C.POINTER(vp), i32, i32, i32],
    "kb_comm_get_handle": [vp, vp],
    "kb_comm_open": [vp, vp],
    "kb_comm_destroy": [vp],
    "kb_comm_error": [vp, C.POINTER(i32)],
    "kb_ws_set_comm": [vp, vp, i32],
    "kb_allreduce": [vp, i32, vp, vp],
    "kb_csr_create": [C.POINTER(vp), i64, i64, i64, vp, vp, vp, i32, vp],
    "kb_csr_destroy": [vp],
    "kb_csr_set_schedule": [vp, i32],
    "kb_csr_get_info": [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32),
                        C.POINTER(i32)],
    "kb_csr_get_stencil": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(f64), C.POINTER(i32),
                           C.POINTER(vp)],
    "kb_spmv": [vp, vp, i32, vp, vp, i32, vp, vp, i32, vp, vp, vp],
    "kb_spmm_is_lines": [vp, i32, vp, C.POINTER(i32)],
    "kb_spmv_halo_add": [vp, i32, i64, f64, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp],
    "kb_halo_create": [C.POINTER(vp), i32, i32, i64],
    "kb_halo_get_handle": [vp, vp],
    "kb_halo_open": [vp, vp],
    "kb_halo_destroy": [vp],
    "kb_halo_error": [vp, C.POINTER(i32)],
    "kb_halo_data_ptr": [vp, i32, C.POINTER(vp)],
    "kb_halo_push": [vp, vp, i32, i32, vp, i64, vp, vp, vp],
    "kb_pack_rows": [vp, i32, i64, vp, vp, vp, vp],
    "kb_dot": [vp, i64, i32, vp, vp, vp, vp],
    "kb_cg_update_xr": [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "kb_cg_update_xr_record": [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp],
    "kb_cg_update_p": [vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp],
    "kb_cg_run": [vp, C.POINTER(CgState), i32, i32, i32, vp],
    "kb_cg_is_fused": [C.POINTER(CgState), C.POINTER(i32)],
    "kb_cg_is_persistent": [vp, C.POINTER(CgState), C.POINTER(i32)],
    "kb_scalar_op": [vp, i32, i32, vp, vp, C.c_double, C.c_double, vp, vp],
    "kb_record": [vp, i32, i32, vp, vp, vp, vp, vp],
    "kb_axpy_dot_minres": [vp, i64, i32, vp, vp, vp, i32, C.POINTER(MinresState), vp],
    "kb_axpy_dot_gmres": [vp, i64, i32, vp, vp, vp, i32, C.POINTER(GmresState), vp],
    "kb_minres_run": [vp, C.POINTER(MinresRunState), i32, i32, vp],
    "kb_gmres_cycle": [vp, C.POINTER(GmresCycleState), i32, i32, vp],
    "kb_cg_run_timed": [vp, C.POINTER(CgState), i32, i32, i32, vp, C.POINTER(C.c_float),
                        C.POINTER(C.c_float)],
    "kb_axpy": [vp, i64, i32, f64, vp, vp, vp, vp],
    "kb_lincomb": [vp, i64, i32, vp, vp, vp, vp, vp, vp],
    "kb_xpby": [vp, i64, i32, vp, vp, vp, vp],
    "kb_div_scale": [vp, i64, i32, vp, vp, vp, vp],
    "kb_add": [vp, i64, i32, vp, vp, vp, vp],
    "kb_axpy_dot": [vp, i64, i32, vp, vp, vp, vp, i32, vp, vp, vp],
    "kb_house_hlast": [vp, vp, i64, vp, vp, vp, vp, vp],
    "kb_minres_scalar": [vp, i32, i32, C.POINTER(MinresState), vp],
    "kb_minres_update": [vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "kb_gmres_scalar": [vp, i32, i32, C.POINTER(GmresState), vp],
    "kb_gmres_solve_y": [vp, i32, i32, i32, vp, vp, vp, vp],
    "kb_basis_combine": [vp, i64, i32, i32, vp, vp, i64, vp, vp, vp],
    "kb_multi_dot": [vp, i64, i32, i32, vp, i64, vp, vp, vp],
    "kb_multi_axpy": [vp, i64, i32, i32, vp, vp, i64, vp, i32, vp, vp],
    "kb_block_gram": [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, vp],
    "kb_block_apply": [vp, i64, i32, i32, vp, i64, vp, i64, vp, i64, vp, i64, i32, vp],
    "kb_house_make": [vp, i64, i64, vp, vp, vp, vp, vp],
    "kb_house_make2": [vp, i64, i64, vp, vp, vp, vp, i32, vp],
    "kb_poke": [vp, i32, vp, i64, vp, f64, vp, vp],
    "kb_lartg": [i32, vp, vp, vp, vp],
    "kb_stencil7": [i32, i32, i32, i32, i32, C.POINTER(f64), vp, vp, vp, vp],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == symbol missing from the .so
    _fn.argtypes = _args
    _fn.restype = i32


# experiments: KRYLOV_B200_TUNE="key=value,key=value" applies kb_tune settings at import
for _kv in filter(None, os.environ.get("KRYLOV_B200_TUNE", "").split(",")):
    _k, _v = _kv.split("=")
    lib.kb_tune(int(_k), int(_v))


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib.kb_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise KrylovB200Error(f"libkrylov_b200 error {rc}: {last_error()}")
