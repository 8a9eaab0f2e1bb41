"""World-size-2 gloo tests (CPU) of the row-partition / halo-exchange logic
(SURVEY.md 8e): the same HaloPlan the GPU path uses, with the kernels
emulated by NumPy, must reproduce the global product and inner products."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, case, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        from krylov_b200 import stencils as st
        from krylov_b200.dist import Comm, HaloPlan, partition_rows

        comm = Comm()
        if case == "stencil":
            nx, ny, nz = 5, 4, 7
            zoff = partition_rows(nz, size)
            offsets = zoff * nx * ny
            rp, ci, va = st.stencil7_csr(nx, ny, nz, coeffs=st.convdiff_coeffs(),
                                         z_lo=int(zoff[rank]), z_hi=int(zoff[rank + 1]))
            Afull = st.to_scipy(st.stencil7_csr(nx, ny, nz, coeffs=st.convdiff_coeffs()))
        else:  # random sparse matrix with long-range couplings and an empty row block
            import scipy.sparse

            n = 41
            Afull = scipy.sparse.random(n, n, density=0.2, random_state=3, format="csr")
            Afull = (Afull + scipy.sparse.eye(n)).tocsr()
            Afull.sort_indices()
            offsets = partition_rows(n, size)
            sub = Afull[int(offsets[rank]):int(offsets[rank + 1])]
            rp, ci, va = sub.indptr, sub.indices, sub.data
        n = Afull.shape[0]
        r0, r1 = int(offsets[rank]), int(offsets[rank + 1])
        plan = HaloPlan(torch.from_numpy(np.asarray(rp)), torch.from_numpy(np.asarray(ci, np.int32)),
                        torch.from_numpy(np.asarray(va, np.float64)), offsets, comm)
        k = 3
        xg = np.random.default_rng(0).standard_normal((n, k))
        x = torch.from_numpy(xg[r0:r1].copy())
        # --- what DistCsrMatrix._apply does, kernels replaced by NumPy
        send = x[plan.send_idx.long()].contiguous()                # kb_pack_rows
        recv = torch.zeros((plan.n_halo, k), dtype=torch.float64)
        works = plan.exchange(send, recv)
        import scipy.sparse as sp
        Aloc = sp.csr_matrix((plan.loc_vals.numpy(), plan.loc_colidx.numpy(), plan.loc_rowptr.numpy()),
                             shape=(r1 - r0, r1 - r0))
        y = Aloc @ x.numpy()                                        # kb_spmv on A_loc
        for w in works:
            w.wait()
        hr, hp = plan.h_rows.numpy(), plan.h_rowptr.numpy()
        hc, hv = plan.h_col.numpy(), plan.h_val.numpy()
        for i, row in enumerate(hr):                                # kb_spmv_halo_add
            for j in range(hp[i], hp[i + 1]):
                y[row] += hv[j] * recv.numpy()[hc[j]]
        yref = (Afull @ xg)[r0:r1]
        ok = np.allclose(y, yref, rtol=1e-13, atol=1e-13)
        # halo columns really are the out-of-range columns, grouped by owner
        hg = plan.halo_globals.numpy()
        ok &= bool(np.all((hg < r0) | (hg >= r1))) and bool(np.all(np.diff(hg) > 0))
        ok &= int(plan.recv_counts.sum()) == plan.n_halo
        # inner product: local partial + one all-reduce
        part = torch.from_numpy(np.einsum("ij,ij->j", x.numpy(), y))
        comm.allreduce(part)
        ok &= np.allclose(part.numpy(), np.einsum("ij,ij->j", xg, Afull @ xg), rtol=1e-12)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["stencil", "random"])
@pytest.mark.parametrize("size", [2, 3])
def test_halo_plan_world(case, size):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(size, port, case, ret), nprocs=size, join=True)
    assert all(ret.get(r, False) for r in range(size)), dict(ret)


def test_partition_rows():
    from krylov_b200.dist import partition_rows

    off = partition_rows(10, 4)
    assert off.tolist() == [0, 3, 6, 8, 10]
    assert partition_rows(512, 8).tolist() == [64 * i for i in range(9)]
