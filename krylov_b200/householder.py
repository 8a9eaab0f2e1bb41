"""Householder reflector -- drop-in for ``krylov.Householder``
(householder.py:6-81): ``H = I - beta v v^T`` with ``H x = alpha ||x|| e_1``.
Quasi-1-D real vectors; built and applied by device kernels
(``kb_house_make``: fused <x[1:], x[1:]> reduction, parameter kernel, fill).
"""
from __future__ import annotations

import numpy as np
import torch

from .device import Ops, as_device_matrix, require_cuda


class Householder:
    def __init__(self, x):
        require_cuda()
        self._is_torch = isinstance(x, torch.Tensor)
        shape = tuple(x.shape)
        assert len(shape) == 1 or (len(shape) == 2 and shape[1] == 1), (
            "Householder only works for quasi-1D vectors for now. "
            f"Input vector has shape {shape}."
        )
        self._shape = shape
        xd = as_device_matrix(x).reshape(shape[0], 1)
        self._ops = ops = Ops(shape[0], 1, xd.device)
        with torch.cuda.device(xd.device):
            self._v = ops.vec(zero=False)
            self._params = torch.empty((5,), dtype=torch.float64, device=xd.device)
            self._tau = ops.slots(1)
            ops.house_make(0, xd, self._v, self._params, self._tau[0])
            pr = self._params.cpu().numpy()
        self.alpha = float(pr[0])
        self.beta = int(pr[1])
        self.xnorm = float(pr[2])
        self.v = self._out(self._v)

    def _out(self, t):
        t = t.reshape(self._shape)
        return t if self._is_torch else t.cpu().numpy()

    def __matmul__(self, x):
        """x - beta v <v, x>   (householder.py:53-62)"""
        if tuple(x.shape) != self._shape:
            raise ValueError(
                f"Shape mismatch! (v.shape = {self._shape} != {tuple(x.shape)} = x.shape)")
        if self.beta == 0:
            return x
        ops = self._ops
        xd = as_device_matrix(x, self._v.device).reshape(self._shape[0], 1).clone()
        with torch.cuda.device(xd.device):
            ops.dot(self._v, xd, self._tau[0])
            ops.axpy_dot(self._tau[0], self._v, xd, dot=0, scale=self._params[1:2])
        out = xd.reshape(self._shape)
        return out if isinstance(x, torch.Tensor) else out.cpu().numpy()

    def matrix(self):
        """Dense matrix I - beta v v^T (test aid, householder.py:64-81)."""
        v = self.v.cpu().numpy() if isinstance(self.v, torch.Tensor) else self.v
        n = v.shape[0]
        eye = np.zeros([n, n] + list(v.shape[1:]))
        i = np.arange(n)
        eye[i, i] = 1.0
        return eye - self.beta * np.einsum("i...,j...->ij...", v, v.conj())
