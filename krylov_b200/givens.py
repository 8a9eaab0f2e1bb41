"""Givens rotations -- drop-in for ``krylov.givens`` (givens.py:5-47).

The rotation parameters are computed on the device by the same ``kb_dlartg``
routine (LAPACK 3.10 ``dlartg`` semantics, the function the reference calls
through SciPy) that the fused MINRES/GMRES scalar kernels use.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib
from .device import cur_stream, ptr, require_cuda


def givens(X):
    """``X.shape == (2, ...)``.  Returns ``G`` of shape ``(2, 2, ...)`` with
    ``G[..., j] = [[c, s], [-s, c]]`` such that ``G @ X[:, j] = [r, 0]``, and ``r``."""
    require_cuda()
    is_torch = isinstance(X, torch.Tensor)
    if is_torch:
        if X.is_complex():
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        Xd = X.to(device="cuda", dtype=torch.float64)
    else:
        Xn = np.asarray(X)
        if np.iscomplexobj(Xn):
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        Xd = torch.from_numpy(np.ascontiguousarray(Xn, dtype=np.float64)).cuda()
    assert Xd.shape[0] == 2
    tail = tuple(Xd.shape[1:])
    flat = Xd.reshape(2, -1).contiguous()
    m = flat.shape[1]
    out = torch.empty((m, 3), dtype=torch.float64, device=flat.device)
    with torch.cuda.device(flat.device):
        check(lib.kb_lartg(m, ptr(flat[0]), ptr(flat[1]), ptr(out), cur_stream()))
    c, s, r = out[:, 0], out[:, 1], out[:, 2]
    G = torch.stack([torch.stack([c, s]), torch.stack([-s, c])]).reshape(2, 2, *tail)
    r = r.reshape(m) if tail else r.reshape(1)
    if is_torch:
        return G, r
    return G.cpu().numpy(), r.cpu().numpy()
