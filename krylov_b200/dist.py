"""Row-partitioned sparse matrices over several GPUs (one process per GPU).

The reference has no distributed layer; this is the
B200-native addition BASELINE.json asks for.  Design:

* 1-D contiguous row partition: rank p owns rows [offsets[p], offsets[p+1]) of
  A and the matching slices of every vector, so all vector kernels are local.
* The local rows are split PETSc-style into ``A_loc`` (columns the rank owns,
  renumbered locally) and a compressed halo part (boundary rows only, columns
  renumbered into a receive buffer ordered by owner rank).
* One product = pack boundary entries -> grouped NCCL send/recv over
  NVLink/NVSwitch (``torch.distributed`` P2P, asynchronous on NCCL's stream)
  -> ``A_loc x`` on the compute stream *while the halo is in flight* -> wait ->
  ``kb_spmv_halo_add`` finishes the boundary rows and adds its share of the
  fused inner product to the same reduction slot.
* Every inner product is a local deterministic partial followed by ONE small
  all-reduce of k doubles (``Ops.reduce_over_ranks``); Givens/Hessenberg
  scalar kernels run replicated on every rank.

``HaloPlan`` is pure torch (CPU or CUDA tensors) so the partition/halo logic is
tested with the gloo backend on CPU (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition_rows(n, size):
    """Contiguous, balanced row offsets (size+1,)."""
    base, rem = divmod(n, size)
    off = np.zeros(size + 1, dtype=np.int64)
    for p in range(size):
        off[p + 1] = off[p] + base + (1 if p < rem else 0)
    return off


class Comm:
    """Process group + (on CUDA, one node) the peer-memory communicator.

    ``allreduce_mode``:
      * ``"p2p"``  -- every reduction kernel finishes with a one-shot all-reduce
        through NVLink peer memory inside its own last block (``kb_comm_*``,
        mailboxes mapped with CUDA IPC): no separate collective launch at all.
      * ``"nccl"`` -- one ``ncclAllReduce`` of k doubles after each reduction.
    Chosen by ``KRYLOV_B200_ALLREDUCE`` (default ``p2p`` on CUDA)."""

    def __init__(self, group=None, allreduce_mode=None):
        import os

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self.allreduces = 0
        self.p2p_handle = None
        mode = allreduce_mode or os.environ.get("KRYLOV_B200_ALLREDUCE", "p2p")
        if mode == "p2p" and self.size > 1 and torch.cuda.is_available() \
                and dist.get_backend(group) == "nccl":
            self._open_p2p()
        self.allreduce_mode = "p2p" if self.p2p_handle is not None else "nccl"

    def _all_ok(self, ok: bool) -> bool:
        """Collective AND: peer-memory paths are only used if every rank could set them up."""
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    def _open_p2p(self, max_k=256):
        """Maps every rank's mailbox with CUDA IPC.  If any rank cannot (no IPC in this
        container, GPUs on different nodes, ...) all ranks fall back to NCCL together."""
        import ctypes as C
        import warnings

        from ._lib import lib

        h = C.c_void_p()
        buf = C.create_string_buffer(64)
        ok = lib.kb_comm_create(C.byref(h), self.rank, self.size, max_k) == 0
        ok = ok and lib.kb_comm_get_handle(h, buf) == 0
        handles = self.allgather_object(bytes(buf.raw) if ok else b"")
        ok = ok and all(len(x) == 64 for x in handles)
        if ok:
            allh = C.create_string_buffer(b"".join(handles), 64 * self.size)
            ok = lib.kb_comm_open(h, allh) == 0
        if self._all_ok(ok):
            self.p2p_handle = h  # every mailbox is mapped before anyone writes
        else:
            if h:
                lib.kb_comm_destroy(h)
            if self.rank == 0:
                warnings.warn("krylov_b200: CUDA IPC peer mapping unavailable; using NCCL "
                              "all-reduce and send/recv")

    def close(self):
        if self.p2p_handle is not None:
            from ._lib import lib

            try:
                lib.kb_comm_destroy(self.p2p_handle)
            except Exception:
                pass
            self.p2p_handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check_p2p(self):
        """Raise if a peer ever failed to arrive in a fused all-reduce."""
        if self.p2p_handle is not None:
            import ctypes as C

            from ._lib import KrylovB200Error, check, lib

            e = C.c_int(0)
            check(lib.kb_comm_error(self.p2p_handle, C.byref(e)))
            if e.value:
                raise KrylovB200Error("peer-memory all-reduce timed out waiting for a rank")

    def allreduce(self, t):
        """In-place sum over ranks; stream-ordered for NCCL (no host sync)."""
        self.allreduces += 1
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def allgather_object(self, obj):
        """Every rank's object.  ``bytes`` of at most 256 bytes (IPC handles) on an NCCL group
        travel as ONE fixed-size device all-gather instead of torch's pickled object collective
        (two collectives + host staging: milliseconds each at 8 ranks, and a row-partitioned
        solve set-up makes several)."""
        if (isinstance(obj, bytes) and len(obj) <= 256 and torch.cuda.is_available()
                and dist.get_backend(self.group) == "nccl"):
            buf = torch.zeros(260, dtype=torch.uint8)
            buf[0] = len(obj) & 0xFF
            buf[1] = len(obj) >> 8
            if obj:
                buf[4:4 + len(obj)] = torch.frombuffer(bytearray(obj), dtype=torch.uint8)
            mine = buf.cuda()
            allb = torch.empty(260 * self.size, dtype=torch.uint8, device=mine.device)
            dist.all_gather_into_tensor(allb, mine, group=self.group)
            allb = allb.cpu().reshape(self.size, 260)
            return [bytes(allb[p, 4:4 + int(allb[p, 0]) + 256 * int(allb[p, 1])].tolist())
                    for p in range(self.size)]
        out = [None] * self.size
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def exchange_halo_ids(self, halo_globals, recv_counts):
        """halo_globals: the global column ids this rank needs, sorted (so grouped by owner);
        recv_counts[p]: how many of them rank p owns.  Returns (ids the other ranks need from this
        rank, concatenated by requesting rank; their counts; the full count matrix
        all_recv[p][q] = what p gets from q).  NCCL: one all-gather of the counts and one
        all-to-all of exactly the ids each owner has to see; otherwise object collectives."""
        size, rank = self.size, self.rank
        if (torch.cuda.is_available() and halo_globals.is_cuda
                and dist.get_backend(self.group) == "nccl"):
            dev = halo_globals.device
            mine = torch.as_tensor(np.asarray(recv_counts, dtype=np.int64)).to(dev)
            allc = torch.empty(size * size, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allc, mine, group=self.group)
            all_recv = allc.cpu().numpy().reshape(size, size)
            send_counts = all_recv[:, rank].copy()
            out = torch.empty(int(send_counts.sum()), dtype=torch.int64, device=dev)
            dist.all_to_all_single(out, halo_globals.to(torch.int64).contiguous(),
                                   output_split_sizes=[int(c) for c in send_counts],
                                   input_split_sizes=[int(c) for c in recv_counts],
                                   group=self.group)
            return out, send_counts, all_recv
        hg = halo_globals.cpu().numpy()
        bounds = np.concatenate([[0], np.cumsum(recv_counts)])
        wanted = [hg[bounds[p]:bounds[p + 1]] for p in range(size)]  # global ids I need from p
        everyone = self.allgather_object(wanted)                     # everyone[q][p]: q needs from p
        lists = [np.asarray(everyone[q][rank], dtype=np.int64) for q in range(size)]
        send_counts = np.array([len(x) for x in lists], dtype=np.int64)
        cat = np.concatenate(lists) if send_counts.sum() else np.zeros(0, np.int64)
        all_recv = np.asarray(self.allgather_object([int(c) for c in recv_counts]), dtype=np.int64)
        return torch.from_numpy(cat).to(halo_globals.device), send_counts, all_recv


class HaloPlan:
    """Splits this rank's CSR rows (global column indices) into a local part
    and a halo part, and works out who sends what to whom.

    Inputs are torch tensors on any device: rowptr (n_loc+1, int32/64),
    colidx (nnz), vals (nnz), the global row offsets of all ranks."""

    def __init__(self, rowptr, colidx, vals, offsets, comm):
        self.comm = comm
        rank = comm.rank
        self.offsets = np.asarray(offsets, dtype=np.int64)
        r0, r1 = int(self.offsets[rank]), int(self.offsets[rank + 1])
        self.row_start, self.row_end = r0, r1
        n_loc = r1 - r0
        assert rowptr.numel() == n_loc + 1
        dev = colidx.device
        nnz = int(vals.numel())
        colidx = colidx[:nnz]
        rp = rowptr.to(torch.int64)
        is_halo = (colidx < r0) | (colidx >= r1)
        # ---- local part: drop halo entries, renumber columns
        owned_prefix = torch.zeros(nnz + 1, dtype=torch.int64, device=dev)
        owned_prefix[1:] = torch.cumsum((~is_halo).to(torch.int64), 0)
        self.loc_rowptr = owned_prefix[rp].to(torch.int32)
        keep = ~is_halo
        self.loc_colidx = (colidx[keep] - r0).to(torch.int32)
        self.loc_vals = vals[:nnz][keep]
        del owned_prefix, keep
        # ---- halo part: boundary rows only
        hidx = torch.nonzero(is_halo).reshape(-1)  # ascending nnz index => ascending row
        hcols_g = colidx[hidx].to(torch.int64)
        self.halo_globals = torch.unique(hcols_g)  # sorted => grouped by owner rank
        self.n_halo = int(self.halo_globals.numel())
        self.h_col = torch.searchsorted(self.halo_globals, hcols_g).to(torch.int32)
        self.h_val = vals[:nnz][hidx]
        rows_of = torch.searchsorted(rp, hidx, right=True) - 1
        self.h_rows, counts = torch.unique_consecutive(rows_of, return_counts=True)
        self.h_rows = self.h_rows.to(torch.int32)
        self.n_brows = int(self.h_rows.numel())
        self.h_rowptr = torch.zeros(self.n_brows + 1, dtype=torch.int32, device=dev)
        if self.n_brows:
            self.h_rowptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        # ---- who owns my halo columns / who needs my rows
        hg = self.halo_globals.cpu().numpy()
        owner = np.searchsorted(self.offsets, hg, side="right") - 1
        self.recv_counts = np.bincount(owner, minlength=comm.size).astype(np.int64)
        assert self.recv_counts[rank] == 0
        # the ids every other rank needs from me (grouped by requesting rank), and the count
        # matrix all_recv[p][q] = what p gets from q
        ids, self.send_counts, all_recv = comm.exchange_halo_ids(self.halo_globals,
                                                                 self.recv_counts)
        self.n_send = int(ids.numel())
        self.send_idx = (ids - r0).to(torch.int32).to(dev)
        if self.n_send:
            lo, hi = int(self.send_idx.min()), int(self.send_idx.max())
            assert lo >= 0 and hi < n_loc
        self.peers = [p for p in range(comm.size)
                      if p != rank and (self.send_counts[p] or self.recv_counts[p])]
        # ---- peer-memory push tables: where my rows land in each destination's data area
        segs = []
        so = np.concatenate([[0], np.cumsum(self.send_counts)])
        for p in range(comm.size):
            if self.send_counts[p]:
                dst_row = int(np.sum(all_recv[p][:rank]))  # p's buffer is grouped by source rank
                segs.append([p, int(so[p]), int(self.send_counts[p]), dst_row])
        self.n_seg = len(segs)
        self.segs = torch.tensor(segs if segs else [[0, 0, 0, 0]], dtype=torch.int64, device=dev)
        srcs = [p for p in range(comm.size) if self.recv_counts[p]]
        self.n_src = len(srcs)
        self.srcs = torch.tensor(srcs if srcs else [0], dtype=torch.int32, device=dev)

    def exchange(self, send_buf, recv_buf):
        """Post the grouped P2P transfers of one product; returns the works.
        send_buf: (n_send, k) packed rows, grouped by destination rank;
        recv_buf: (n_halo, k), grouped by source rank."""
        ops = []
        so = np.concatenate([[0], np.cumsum(self.send_counts)])
        ro = np.concatenate([[0], np.cumsum(self.recv_counts)])
        for p in self.peers:
            if self.recv_counts[p]:
                ops.append(dist.P2POp(dist.irecv, recv_buf[ro[p]:ro[p + 1]], p,
                                      group=self.comm.group))
            if self.send_counts[p]:
                ops.append(dist.P2POp(dist.isend, send_buf[so[p]:so[p + 1]], p,
                                      group=self.comm.group))
        return dist.batch_isend_irecv(ops) if ops else []


class DistCsrMatrix:
    """This rank's rows of a row-partitioned CSR matrix, on its GPU.

    Behaves like a (n_loc x n_loc) operator on the rank's vector slices:
    ``cg(A_dist, b_local)`` returns the local slice of the solution, with
    residual norms / step counts identical on every rank."""

    is_dist_csr = True
    dtype = np.dtype(np.float64)

    def __init__(self, local_csr, offsets, comm=None):
        """local_csr: CsrMatrix holding rows [offsets[rank], offsets[rank+1])
        with *global* column indices."""
        from .csr import CsrMatrix

        from . import _trace

        self.comm = comm if comm is not None else Comm()
        _trace.mark("dist: communicator (peer-memory mailbox, IPC handles)")
        self.device = local_csr.device
        self.global_shape = (int(offsets[-1]), int(offsets[-1]))
        plan = HaloPlan(local_csr.rowptr, local_csr.colidx, local_csr.vals[: local_csr.nnz],
                        offsets, self.comm)
        self.plan = plan
        _trace.mark("dist: halo plan (split into local / halo parts, send lists)")
        n_loc = plan.row_end - plan.row_start
        self.shape = (n_loc, n_loc)
        self.nnz = local_csr.nnz
        self.A_loc = CsrMatrix(plan.loc_rowptr, plan.loc_colidx, plan.loc_vals, (n_loc, n_loc),
                               self.device)
        plan.loc_rowptr = plan.loc_colidx = plan.loc_vals = None
        _trace.mark("dist: local matrix (kb_csr_create)")
        self._bufs = {}
        self._ops = {}
        self._halos = {}
        self.halo_bytes_per_product = 8 * (plan.n_send + plan.n_halo)
        # peer-memory halo exchange goes with the peer-memory all-reduce (same requirement:
        # one NVLink node, CUDA IPC); otherwise grouped NCCL send/recv
        self.halo_mode = "p2p" if self.comm.p2p_handle is not None else "nccl"
        # the push kernel runs on a high-priority side stream next to the local product
        # (only the neighbours' boundary kernels wait for it, through device flags)
        self._side = None
        self._ev_ready = self._ev_pushed = None
        if self.halo_mode == "p2p":
            with torch.cuda.device(self.device):
                self._side = torch.cuda.Stream(device=self.device, priority=-1)
                self._ev_ready = torch.cuda.Event()
                self._ev_pushed = torch.cuda.Event()

    def spmv_bytes(self, k=1):
        n = self.shape[0]
        return 12 * self.nnz + 4 * (n + 1) + 16 * n * k

    def moved_bytes(self, k=1):
        return self.A_loc.moved_bytes(k) + 12 * (self.nnz - self.A_loc.nnz)

    def info(self):
        d = self.A_loc.info()
        d.update(n_halo=self.plan.n_halo, n_send=self.plan.n_send, n_brows=self.plan.n_brows,
                 peers=list(self.plan.peers))
        return d

    def set_schedule(self, name):
        self.A_loc.set_schedule(name)
        return self

    def _halo_for(self, k):
        """Receive area for k right-hand sides (collective on first use: IPC handles of
        all ranks are exchanged and mapped)."""
        import ctypes as C

        from ._lib import check, lib

        h = self._halos.get(k)
        if h is None:
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                buf = C.create_string_buffer(64)
                ok = lib.kb_halo_create(C.byref(h), self.comm.rank, self.comm.size,
                                        max(self.plan.n_halo, 1) * k * 8) == 0
                ok = ok and lib.kb_halo_get_handle(h, buf) == 0
                handles = self.comm.allgather_object(bytes(buf.raw) if ok else b"")
                ok = ok and all(len(x) == 64 for x in handles)
                if ok:
                    allh = C.create_string_buffer(b"".join(handles), 64 * self.comm.size)
                    ok = lib.kb_halo_open(h, allh) == 0
            if not self.comm._all_ok(ok):  # collective decision: everybody or nobody
                if h:
                    lib.kb_halo_destroy(h)
                self.halo_mode = "nccl"
                return None
            self._halos[k] = h
        return h

    def check_p2p(self):
        import ctypes as C

        from ._lib import KrylovB200Error, check, lib

        self.comm.check_p2p()
        for h in self._halos.values():
            e = C.c_int(0)
            check(lib.kb_halo_error(h, C.byref(e)))
            if e.value:
                raise KrylovB200Error("peer-memory halo exchange timed out waiting for a rank")

    # ------------------------------------------------ two-launch CG across ranks --
    def fused_cg_plan(self):
        """Ghost-plane layout of the fused CG path, or None if this matrix does not qualify.

        Qualifies: z-slab partition of a 3-D constant-coefficient 7-diagonal stencil whose only
        couplings across ranks are the +-P (plane) diagonals towards the two neighbouring ranks,
        every slab a whole number (>= 2) of planes, peer memory available.  Then a CG iteration
        is the two marching kernels of the single-GPU path on the row space
        ``[ghost plane | own planes | ghost plane]`` (csrc/kb_march.cuh): the r update stores its
        first / last plane straight into the neighbours' ghost planes, the ghost planes of p are
        recomputed locally, nothing else crosses NVLink.  Collective on first use (all ranks
        must agree; the extended r lives in an IPC-exported allocation that is kept and reused
        by later solves with this matrix)."""
        if hasattr(self, "_fused_plan"):
            return self._fused_plan
        self._fused_plan = None
        import ctypes as C

        from ._lib import lib
        from .device import view_device_memory

        comm, plan, dev = self.comm, self.plan, self.device
        n = self.shape[0]
        ok = self.halo_mode == "p2p" and comm.p2p_handle is not None and comm.size > 1
        P = 0
        coeffs = ()
        rows_lo = rows_hi = None
        if ok:
            si = self.A_loc.stencil_info()
            off = si["offsets"]
            ok = (si["constv"] and si["nd"] == 7 and self.A_loc.info()["schedule"] == "stencil"
                  and off[3] == 0 and off[0] == -off[6] and off[6] > 0)
        if ok:
            P = off[6]
            coeffs = tuple(si["coeffs"][:7])
            r0 = plan.row_start
            ok = (n % P == 0 and n >= 2 * P and all(int(o) % P == 0 for o in plan.offsets)
                  and set(plan.peers) <= {comm.rank - 1, comm.rank + 1})
        if ok and plan.n_brows:
            # every halo entry must be the -P diagonal of a first-plane row (value c[0]) or the
            # +P diagonal of a last-plane row (value c[6])
            cnt = (plan.h_rowptr[1:] - plan.h_rowptr[:-1]).to(torch.int64)
            rows = torch.repeat_interleave(plan.h_rows.to(torch.int64), cnt)
            diff = plan.halo_globals[plan.h_col.to(torch.int64)] - (rows + r0)
            lo, hi = diff == -P, diff == P
            c_lo = torch.tensor(coeffs[0], dtype=torch.float64, device=dev)
            c_hi = torch.tensor(coeffs[6], dtype=torch.float64, device=dev)
            ok = bool(torch.all(lo | hi)) and bool(torch.all(rows[lo] < P)) \
                and bool(torch.all(rows[hi] >= n - P)) \
                and bool(torch.all(plan.h_val[lo] == c_lo)) and bool(torch.all(plan.h_val[hi] == c_hi))
            rows_lo, rows_hi = rows[lo], rows[hi]
        # every rank must qualify with the same plane stride and coefficients (fixed-size payload:
        # the fast path of allgather_object)
        import struct

        mine = struct.pack("<?q7d", bool(ok), int(P), *(list(coeffs) + [0.0] * (7 - len(coeffs))))
        votes = [struct.unpack("<?q7d", v) for v in comm.allgather_object(mine)]
        if not all(v[0] and v[1:] == votes[0][1:] for v in votes):
            return None
        n_ext = n + 2 * P
        with torch.cuda.device(dev):
            h = C.c_void_p()
            buf = C.create_string_buffer(64)
            good = lib.kb_halo_create(C.byref(h), comm.rank, comm.size, n_ext * 8) == 0
            good = good and lib.kb_halo_get_handle(h, buf) == 0
            handles = comm.allgather_object(bytes(buf.raw) if good else b"")
            good = good and all(len(x) == 64 for x in handles)
            if good:
                allh = C.create_string_buffer(b"".join(handles), 64 * comm.size)
                good = lib.kb_halo_open(h, allh) == 0
            if not comm._all_ok(good):
                if h:
                    lib.kb_halo_destroy(h)
                return None
            self._fused_halo = h

            def data_ptr(rank):
                out = C.c_void_p()
                if lib.kb_halo_data_ptr(h, rank, C.byref(out)) != 0:
                    raise RuntimeError("kb_halo_data_ptr failed")
                return out.value

            r_ext = view_device_memory(data_ptr(comm.rank), n_ext, torch.float64, dev)
            # masks of the extended row space: global pattern on the own rows, 0 on the ghosts
            m_loc = view_device_memory(si["masks_ptr"], n, torch.int16, dev)
            masks = torch.zeros(n_ext, dtype=torch.int16, device=dev)
            masks[P:P + n] = m_loc
            if rows_lo is not None and rows_lo.numel():
                masks[P + rows_lo] |= 1
            if rows_hi is not None and rows_hi.numel():
                masks[P + rows_hi] |= 1 << 6
            sizes = [int(plan.offsets[q + 1] - plan.offsets[q]) for q in range(comm.size)]
            push_lo = push_hi = None
            if comm.rank > 0 and rows_lo is not None and rows_lo.numel():
                # the lower neighbour's UPPER ghost plane: extended rows [P + n_prev, P + n_prev + P)
                push_lo = data_ptr(comm.rank - 1) + 8 * (P + sizes[comm.rank - 1])
            if comm.rank < comm.size - 1 and rows_hi is not None and rows_hi.numel():
                push_hi = data_ptr(comm.rank + 1)  # the upper neighbour's LOWER ghost plane
        self._fused_plan = {"P": P, "n_ext": n_ext, "r_ext": r_ext, "masks_ext": masks,
                            "push_lo": push_lo, "push_hi": push_hi}
        return self._fused_plan

    def exchange_ghost_planes(self, v_ext, P):
        """One-off (set-up) exchange: first / last own plane of the extended vector -> the
        neighbours' ghost planes, by NCCL send/recv.  Inside the iteration the r update does
        this itself with peer stores."""
        n = self.shape[0]
        r, size = self.comm.rank, self.comm.size
        ops = []
        if r > 0:
            ops.append(dist.P2POp(dist.isend, v_ext[P:2 * P], r - 1, group=self.comm.group))
            ops.append(dist.P2POp(dist.irecv, v_ext[:P], r - 1, group=self.comm.group))
        if r < size - 1:
            ops.append(dist.P2POp(dist.isend, v_ext[n:n + P], r + 1, group=self.comm.group))
            ops.append(dist.P2POp(dist.irecv, v_ext[n + P:], r + 1, group=self.comm.group))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()

    def close(self):
        """Unmaps / frees the peer-memory areas of this matrix (halo receive areas, the extended
        r of the fused CG path).  Collective in spirit: call it on every rank."""
        from ._lib import lib

        self._fused_plan = None
        for h in list(self._halos.values()) + [getattr(self, "_fused_halo", None)]:
            if h:
                try:
                    lib.kb_halo_destroy(h)
                except Exception:
                    pass
        self._halos = {}
        self._fused_halo = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buffers(self, k):
        b = self._bufs.get(k)
        if b is None:
            p = self.plan
            b = self._bufs[k] = (
                torch.empty((max(p.n_send, 1), k), dtype=torch.float64, device=self.device),
                torch.empty((max(p.n_halo, 1), k), dtype=torch.float64, device=self.device))
        return b

    def _apply(self, ops, x, y, mode=0, z=None, coef=None, dot=0, w=None, out=None):
        from ._lib import check, lib
        from .device import cur_stream, ptr

        p = self.plan
        k = ops.k
        ldot = dot if dot == 1 else 0  # <y, y> cannot be split into local + halo shares
        fused = ops.fused_allreduce   # reductions end with the peer-memory all-reduce
        halo = self._halo_for(k) if self.halo_mode == "p2p" else None  # None: NCCL send/recv
        works = []
        if halo is not None:
            # boundary rows of x go straight into the neighbours' receive areas (NVLink
            # stores + flag), before the local product is launched: no NCCL kernel at all
            ops.launches += 1
            main = torch.cuda.current_stream()
            self._ev_ready.record(main)          # x is complete on the compute stream
            self._side.wait_event(self._ev_ready)
            check(lib.kb_halo_push(halo, ops.ws.handle, k, p.n_seg, ptr(p.segs), p.n_send,
                                   ptr(p.send_idx), ptr(x), self._side.cuda_stream))
            self._ev_pushed.record(self._side)
            recv_buf = None
        else:
            send_buf, recv_buf = self._buffers(k)
            if p.n_send or p.n_halo:
                if p.n_send:
                    ops.launches += 1
                    check(lib.kb_pack_rows(ops.ws.handle, k, p.n_send, ptr(p.send_idx), ptr(x),
                                           ptr(send_buf), cur_stream()))
                works = p.exchange(send_buf[: p.n_send], recv_buf[: p.n_halo])
        if fused:
            ops.set_collective(False)  # the local part of <w, y> must not be exchanged yet
        ops.launches += 1
        check(lib.kb_spmv(self.A_loc.handle, ops.ws.handle, k, ptr(x), ptr(y), int(mode), ptr(z),
                          ptr(coef), ldot, ptr(w), ptr(out), cur_stream()))
        if fused:
            ops.set_collective(True)
        for wk in works:
            wk.wait()  # compute stream waits for NCCL's stream; the host does not block
        if p.n_brows:
            ops.launches += 1
            check(lib.kb_spmv_halo_add(ops.ws.handle, k, p.n_brows, -1.0 if mode == 2 else 1.0,
                                       ptr(p.h_rows), ptr(p.h_rowptr), ptr(p.h_col), ptr(p.h_val),
                                       ptr(recv_buf), ptr(y), ldot, ptr(w), ptr(out),
                                       halo, ptr(p.srcs), p.n_src if halo is not None else 0,
                                       cur_stream()))
        elif fused and ldot:
            ops.launches += 1  # no boundary rows here, but the peers' collective needs this rank
            check(lib.kb_allreduce(ops.ws.handle, k, ptr(out), cur_stream()))
        if halo is not None:
            # nobody may overwrite x before the push has read it (it finished long ago)
            torch.cuda.current_stream().wait_event(self._ev_pushed)
        if dot == 2:
            ops.launches += 1
            check(lib.kb_dot(ops.ws.handle, ops.n, k, ptr(y), ptr(y), ptr(out), cur_stream()))
        if dot:
            ops.reduce_over_ranks(out)

    def matvec_device(self, x, out=None):
        from .device import Ops

        k = 1 if x.dim() == 1 else x.shape[1]
        o = self._ops.get(k)
        if o is None:
            o = self._ops[k] = Ops(self.shape[0], k, self.device, comm=self.comm)
        x = x.contiguous()
        y = out if out is not None else torch.empty_like(x)
        with torch.cuda.device(self.device):
            self._apply(o, x, y)
        return y

    def __matmul__(self, x):
        from .device import as_device_matrix

        if isinstance(x, torch.Tensor):
            return self.matvec_device(as_device_matrix(x, self.device))
        return self.matvec_device(as_device_matrix(x, self.device)).cpu().numpy()


def dist_stencil7(nx, ny, nz, coeffs=None, shift=0.0, comm=None, device=None):
    """z-slab partition of the 7-point operator: rank p builds planes
    [nz*p/P, nz*(p+1)/P) on its own GPU."""
    from .generate import device_stencil7
    from .stencils import STENCIL_POISSON

    comm = comm if comm is not None else Comm()
    zoff = partition_rows(nz, comm.size)
    z_lo, z_hi = int(zoff[comm.rank]), int(zoff[comm.rank + 1])
    local = device_stencil7(nx, ny, nz, coeffs or STENCIL_POISSON, shift, z_lo, z_hi, device)
    A = DistCsrMatrix(local, zoff * nx * ny, comm)
    del local
    return A
