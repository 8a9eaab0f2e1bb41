"""Device-side generators for the synthetic stencil matrices.

``device_stencil7`` builds the CSR arrays of a z-slab of the 7-point operator
directly in HBM with the ``kb_stencil7`` kernel (count pass -> scan -> fill
pass): the 512^3 case is 11.8 GB of CSR and is never materialised on the host.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import check, lib
from .csr import CsrMatrix
from .device import cur_stream, ptr, require_cuda
from .stencils import STENCIL_POISSON


def device_stencil7(nx, ny, nz, coeffs=STENCIL_POISSON, shift=0.0, z_lo=0, z_hi=None,
                    device=None, local_columns=False):
    """Rows z_lo <= z < z_hi of the 7-point operator as a CsrMatrix with
    *global* column indices (shape (n_loc, nx*ny*nz)).  Same entry order and
    values as ``stencils.stencil7_csr``."""
    require_cuda()
    z_hi = nz if z_hi is None else z_hi
    dev = torch.device(device) if device is not None else torch.device(
        "cuda", torch.cuda.current_device())
    cf = list(coeffs)
    cf[0] = cf[0] - shift
    carr = (C.c_double * 7)(*cf)
    n_loc = nx * ny * (z_hi - z_lo)
    with torch.cuda.device(dev):
        counts = torch.empty(n_loc + 1, dtype=torch.int32, device=dev)
        check(lib.kb_stencil7(nx, ny, nz, z_lo, z_hi, carr, ptr(counts), None, None, cur_stream()))
        rowptr = torch.cumsum(counts, 0, dtype=torch.int64)
        nnz = int(rowptr[-1].item())
        if nnz >= 2**31:
            raise ValueError("nnz must fit int32")
        rowptr = rowptr.to(torch.int32)
        del counts
        npad = ((nnz + 3) // 4) * 4 + 4
        colidx = torch.zeros(npad, dtype=torch.int32, device=dev)
        vals = torch.zeros(npad, dtype=torch.float64, device=dev)
        check(lib.kb_stencil7(nx, ny, nz, z_lo, z_hi, carr, ptr(rowptr), ptr(colidx), ptr(vals),
                              cur_stream()))
        return CsrMatrix._from_device_arrays(rowptr, colidx, vals, nnz, (n_loc, nx * ny * nz))
