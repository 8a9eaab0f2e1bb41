"""Device-resident CSR matrix: the object behind ``A @ x`` on the hot path.

Replaces the SciPy CSR matrix (sparsetools ``csr_matvec`` / ``csr_matvecs``)
the reference multiplies with at ``_helpers.py:47,61``.  fp64 values, int32
indices (index traffic is a third of the matrix bytes).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _trace
from ._lib import check, lib
from .device import Ops, as_device_matrix, cur_stream, ptr, require_cuda

_SCHEDULES = {"auto": 0, "rowwise": 1, "stream": 2, "pattern": 3, "stencil": 4,
              "merge": 5}
_PAD = 4  # elements readable past nnz (TMA tiles are 4-aligned windows)


def _padded_upload(src, dtype, device) -> torch.Tensor:
    """Device copy of a 1-D array with >= 4 readable zero elements past the end, made
    with ONE allocation and one copy (host data goes straight into the padded buffer)."""
    t = torch.as_tensor(src)
    n = t.numel()
    out = torch.empty(((n + 3) // 4) * 4 + _PAD, dtype=dtype, device=device)
    out[n:].zero_()
    out[:n].copy_(t.reshape(-1), non_blocking=True)  # converts dtype / crosses PCIe as needed
    return out


class CsrMatrix:
    """CSR matrix in HBM.  ``shape``/``dtype``/``__matmul__`` make it a valid
    operator for the solvers (reference protocol: ``_helpers.py:14-17``)."""

    dtype = np.dtype(np.float64)

    def __init__(self, rowptr, colidx, vals, shape, device=None):
        require_cuda()
        dev = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        self.device = dev
        self.shape = (int(shape[0]), int(shape[1]))
        rowptr = torch.as_tensor(rowptr).to(device=dev, dtype=torch.int32).contiguous()
        if rowptr.numel() != self.shape[0] + 1:
            raise ValueError("rowptr must have n_rows + 1 entries")
        self.nnz = int(torch.as_tensor(vals).numel())
        if int(torch.as_tensor(colidx).numel()) != self.nnz:
            raise ValueError("colidx and vals differ in length")
        self.rowptr = rowptr
        _trace.mark("csr: rowptr uploaded")
        with torch.cuda.device(dev):
            self.colidx = _padded_upload(colidx, torch.int32, dev)
            _trace.mark("csr: colidx uploaded")
            self.vals = _padded_upload(vals, torch.float64, dev)
        _trace.mark("csr: vals uploaded")
        self._finish()
        _trace.mark("csr: kb_csr_create (statistics, pattern, constant diagonals)")

    @classmethod
    def _from_device_arrays(cls, rowptr, colidx_padded, vals_padded, nnz, shape):
        """Adopt already padded device arrays without a copy (stencil generator)."""
        self = cls.__new__(cls)
        self.device = vals_padded.device
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(nnz)
        self.rowptr, self.colidx, self.vals = rowptr, colidx_padded, vals_padded
        self._finish()
        return self

    def _finish(self):
        self._ops = {}
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.kb_csr_create(C.byref(h), self.shape[0], self.shape[1], self.nnz,
                                    ptr(self.rowptr), ptr(self.colidx), ptr(self.vals), 1,
                                    cur_stream()))
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.kb_csr_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------ builders --
    @classmethod
    def from_scipy(cls, A, device=None):
        A = A.tocsr()
        if np.iscomplexobj(A.data):
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        if A.nnz >= 2**31:
            raise ValueError("nnz must fit int32")
        # no host-side copies when the dtypes already match (pinned arrays stay pinned)
        return cls(np.asarray(A.indptr, dtype=np.int32), np.asarray(A.indices, dtype=np.int32),
                   np.asarray(A.data, dtype=np.float64), A.shape, device)

    @classmethod
    def from_dense(cls, A, device=None):
        import scipy.sparse

        A = np.asarray(A)
        if np.iscomplexobj(A):
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        return cls.from_scipy(scipy.sparse.csr_matrix(A.astype(np.float64)), device)

    @classmethod
    def from_torch(cls, A, device=None):
        if A.layout == torch.sparse_coo:  # converted on the device, like from_coo
            A = A.coalesce()
            idx = A.indices()
            return cls.from_coo(idx[0], idx[1], A.values(), A.shape,
                                device or (A.device if A.is_cuda else None))
        if A.layout != torch.sparse_csr:
            A = A.to_sparse_csr()
        return cls(A.crow_indices(), A.col_indices(), A.values(), A.shape,
                   device or (A.device if A.is_cuda else None))

    @classmethod
    def from_coo(cls, rows, cols, vals, shape, device=None):
        """Triplets (NumPy arrays or torch tensors, on the host or already on the device) -> CSR,
        converted ON the device: entries sorted by (row, column), duplicates summed in that order
        (SciPy's ``coo_matrix.tocsr`` convention), row pointers from the row counts."""
        require_cuda()
        dev = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        nr, nc = int(shape[0]), int(shape[1])
        with torch.cuda.device(dev):
            r = torch.as_tensor(rows).to(device=dev, dtype=torch.int64).reshape(-1)
            c = torch.as_tensor(cols).to(device=dev, dtype=torch.int64).reshape(-1)
            v = torch.as_tensor(vals)
            if v.is_complex():
                raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
            v = v.to(device=dev, dtype=torch.float64).reshape(-1)
            if not (r.numel() == c.numel() == v.numel()):
                raise ValueError("rows, cols and vals differ in length")
            if r.numel() and (int(r.min()) < 0 or int(r.max()) >= nr or int(c.min()) < 0
                              or int(c.max()) >= nc):
                raise ValueError("index out of range")
            key = r * nc + c
            key, perm = torch.sort(key, stable=True)
            v = v[perm]
            uniq, counts = torch.unique_consecutive(key, return_counts=True)
            if uniq.numel() != key.numel():  # duplicates: summed in stored order (no atomics)
                try:
                    v = torch.segment_reduce(v, "sum", lengths=counts)
                except RuntimeError:  # builds without the CUDA segment kernel: atomics (the sum of
                    # three or more duplicates may then differ in the last bit run to run)
                    inv = torch.repeat_interleave(
                        torch.arange(uniq.numel(), device=dev), counts)
                    v = torch.zeros(uniq.numel(), dtype=torch.float64,
                                    device=dev).index_add_(0, inv, v)
            if uniq.numel() >= 2**31:
                raise ValueError("nnz must fit int32")
            rp = torch.zeros(nr + 1, dtype=torch.int64, device=dev)
            rp[1:] = torch.cumsum(torch.bincount(uniq // nc, minlength=nr), 0)
            return cls(rp.to(torch.int32), (uniq % nc).to(torch.int32), v, (nr, nc), dev)

    @classmethod
    def from_file(cls, path, device=None):
        """Matrix Market (``.mtx`` / ``.mtx.gz``, coordinate or array, real / integer / pattern,
        general or symmetric -- parsed by ``scipy.io.mmread``) or SciPy's ``.npz`` sparse container.
        The triplets go to the device as they are and are converted there (``from_coo``)."""
        path = str(path)
        if path.endswith(".npz"):
            import scipy.sparse

            return cls.from_scipy(scipy.sparse.load_npz(path), device)
        import scipy.io

        M = scipy.io.mmread(path)
        if isinstance(M, np.ndarray):
            return cls.from_dense(M, device)
        if np.iscomplexobj(M.data):
            raise NotImplementedError("complex dtypes are out of scope (north_star: fp64)")
        M = M.tocoo()
        return cls.from_coo(M.row, M.col, M.data, M.shape, device)

    @property
    def T(self):
        """Transposed matrix as a CsrMatrix with ascending columns (cached): ``A.T @ x`` then
        accumulates every entry over the rows of A in ascending order -- the order of SciPy's
        ``csc_matvec`` behind the reference's ``rmatvec`` (_helpers.py:65-77), so the adjoint
        product is bit-identical too.  Built once on the device (set-up, not the hot path):
        entries sorted by (column, row), row pointers from the column counts."""
        t = getattr(self, "_T", None)
        if t is None:
            nnz, (nr, nc) = self.nnz, self.shape
            with torch.cuda.device(self.device):
                rp = self.rowptr.to(torch.int64)
                rows = torch.repeat_interleave(torch.arange(nr, device=self.device), rp[1:] - rp[:-1])
                cols = self.colidx[:nnz].to(torch.int64)
                perm = torch.argsort(cols * nr + rows)  # keys are unique: any sort gives one order
                del rows
                new_rp = torch.zeros(nc + 1, dtype=torch.int64, device=self.device)
                new_rp[1:] = torch.cumsum(torch.bincount(cols, minlength=nc), 0)
                del cols
                # row index of every sorted entry = searchsorted of its position in rowptr
                new_cols = (torch.searchsorted(rp, perm, right=True) - 1).to(torch.int32)
                new_vals = self.vals[:nnz][perm]
                t = CsrMatrix(new_rp.to(torch.int32), new_cols, new_vals, (nc, nr), self.device)
            t._T = self
            self._T = t
        return t

    # ------------------------------------------------------------- queries --
    def info(self):
        nr, nc, nz = C.c_int64(), C.c_int64(), C.c_int64()
        mx, sc = C.c_int(), C.c_int()
        check(lib.kb_csr_get_info(self.handle, C.byref(nr), C.byref(nc), C.byref(nz), C.byref(mx),
                                  C.byref(sc)))
        return {"n_rows": nr.value, "n_cols": nc.value, "nnz": nz.value,
                "max_row_len": mx.value,
                "schedule": {1: "rowwise", 2: "stream", 3: "pattern", 4: "stencil",
                             5: "merge"}[sc.value]}

    def stencil_info(self):
        """Offset pattern detected at creation: ``{"nd", "offsets", "coeffs", "constv",
        "masks_ptr"}`` (nd = 0: no pattern; coeffs valid when constv; masks_ptr: device address
        of the library-owned 16-bit row masks)."""
        nd, cv = C.c_int(), C.c_int()
        offs = (C.c_int * 16)()
        co = (C.c_double * 8)()
        mp = C.c_void_p()
        check(lib.kb_csr_get_stencil(self.handle, C.byref(nd), offs, co, C.byref(cv), C.byref(mp)))
        return {"nd": nd.value, "offsets": [int(o) for o in offs[: nd.value]],
                "coeffs": [float(c) for c in co[: min(nd.value, 8)]], "constv": bool(cv.value),
                "masks_ptr": mp.value}

    def set_schedule(self, name: str):
        check(lib.kb_csr_set_schedule(self.handle, _SCHEDULES[name]))
        return self

    def spmv_bytes(self, k=1):
        """Algorithmic bytes of one product: 12 nnz + 4(n+1) + 16 n k."""
        n = self.shape[0]
        return 12 * self.nnz + 4 * (n + 1) + 16 * n * k

    def moved_bytes(self, k=1):
        """Bytes the chosen schedule actually streams per product: the offset-pattern
        schedule replaces the 4-byte column index per nonzero by a 2-byte mask per row; the
        stencil schedule (constant diagonals) streams no matrix values or row pointers either."""
        n = self.shape[0]
        sched = self.info()["schedule"]
        if k == 1 and sched == "pattern":
            return 8 * self.nnz + 2 * n + 4 * (n + 1) + 16 * n
        if k == 1 and sched == "stencil":
            return 2 * n + 16 * n
        return self.spmv_bytes(k)

    def to_scipy(self):
        import scipy.sparse

        return scipy.sparse.csr_matrix(
            (self.vals[: self.nnz].cpu().numpy(), self.colidx[: self.nnz].cpu().numpy(),
             self.rowptr.cpu().numpy()), shape=self.shape)

    # -------------------------------------------------------------- product --
    def _apply(self, ops: Ops, x, y, mode=0, z=None, coef=None, dot=0, w=None, out=None):
        ops.launches += 1
        check(lib.kb_spmv(self.handle, ops.ws.handle, ops.k, ptr(x), ptr(y), int(mode), ptr(z),
                          ptr(coef), int(dot), ptr(w), ptr(out), cur_stream()))
        if dot:
            ops.reduce_over_ranks(out)

    def _ops_for(self, k):
        o = self._ops.get(k)
        if o is None:
            o = self._ops[k] = Ops(self.shape[0], k, self.device)
        return o

    def matvec_device(self, x: torch.Tensor, out=None) -> torch.Tensor:
        """y = A x for a CUDA fp64 tensor of shape (n,) or (n, k)."""
        k = 1 if x.dim() == 1 else x.shape[1]
        if x.shape[0] != self.shape[1]:
            raise ValueError(f"dimension mismatch: {self.shape} @ {tuple(x.shape)}")
        x = x.contiguous()
        y = out if out is not None else torch.empty(
            (self.shape[0],) + tuple(x.shape[1:]), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._apply(self._ops_for(k), x, y)
        return y

    def __matmul__(self, x):
        if isinstance(x, torch.Tensor):
            return self.matvec_device(as_device_matrix(x, self.device))
        xd = as_device_matrix(x, self.device)
        return self.matvec_device(xd).cpu().numpy()

    matvec = __matmul__
