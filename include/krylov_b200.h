/*
 * krylov_b200.h -- C ABI of libkrylov_b200.so (sm_100a).
 *
 * The reference (ju-liu/krylov, pure Python) has no FFI; its "kernels" are the
 * NumPy/SciPy/LAPACK call sites listed in SURVEY.md section 2.1.  Each entry
 * point below names the reference call site(s) it replaces
 * (paths relative to /root/reference/src/krylov/).
 *
 * Conventions
 *  - every function returns 0 on success, a negative KB_E* code otherwise;
 *    kb_last_error() returns the thread-local message of the last failure.
 *  - vectors are fp64 device pointers of logical shape (n, k), row-major,
 *    contiguous (k right-hand sides advance in lock-step);
 *    "slot" arguments are device arrays of k doubles (one scalar per column).
 *  - the caller owns every vector/matrix buffer; the library owns only the
 *    opaque handles and the reduction workspace inside kb_ws.
 *  - every launch goes to `stream` (a cudaStream_t cast to void*), is
 *    asynchronous and never synchronises the device.
 *  - reductions are deterministic: block partials in a fixed order, finished
 *    by the last-arriving block in a fixed-shape pass.
 *  - gating: kb_ws_set_gate(ws, stop_at, tag) makes every later launch through
 *    that workspace a no-op when *stop_at <= tag at kernel start (device-side
 *    convergence control without a host round-trip per iteration).
 */
#ifndef KRYLOV_B200_H
#define KRYLOV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_OK 0
#define KB_EINVAL -1
#define KB_ECUDA -2
#define KB_ENOMEM -3
#define KB_EUNSUPPORTED -4

typedef struct kb_csr_s* kb_csr_t; /* CSR matrix view + schedule */
typedef struct kb_ws_s* kb_ws_t;   /* reduction workspace + gate  */
typedef struct kb_comm_s* kb_comm_t; /* peer-memory communicator (one per process/GPU) */
typedef struct kb_halo_s* kb_halo_t; /* peer-memory halo receive area of one partitioned matrix */

/* --- library ---------------------------------------------------------- */
int kb_version(void);
int kb_last_error(char* buf, size_t len);
int kb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* developer tunables: key 0 stream-kernel configuration (0..5), key 1 stream CTAs/SM
 * (0 = default), key 2 CTAs/SM of the vector and row-wise grids, key 3 CTAs/SM of the
 * pattern kernels, key 4 windowed-kernel configuration (-1 = gather variant); keys 5-15: see
 * csrc/kb_api.cu (tile shapes and CTAs/SM of the stencil / marching / fused CG kernels);
 * key 16 line-marching SpMM for blocked right-hand sides (0 off, 1 default rule, 2 wherever valid),
 * 17 its lines per work item (0 = 32), 18 its chunk size (0 = 1024 entries, 1 = 512),
 * 19 its work-item order (1 = planes fastest, 0 = natural) */
int kb_tune(int key, int value);

/* --- workspace -------------------------------------------------------- */
int kb_ws_create(kb_ws_t* ws, int max_k);
int kb_ws_destroy(kb_ws_t ws);
/* stop_at: device int (or NULL to disable); launches are skipped when *stop_at <= tag */
int kb_ws_set_gate(kb_ws_t ws, const int* stop_at, int tag);

/* --- peer-memory communicator (row-partitioned problems) ---
 * One process per GPU on one NVLink/NVSwitch node.  Each rank creates a mailbox,
 * exports its 64-byte CUDA-IPC handle, the handles of all ranks (rank order, 64
 * bytes each) are handed to kb_comm_open.  A workspace with a communicator and
 * collective = 1 finishes every reduction with a one-shot all-reduce INSIDE the
 * reducing kernel's last block (values and flags written straight into the peers'
 * mailboxes over NVLink): one kernel does the local partial, the exchange and the
 * deterministic rank-ordered sum.  Replaces one NCCL all-reduce per inner product. */
int kb_comm_create(kb_comm_t* c, int rank, int size, int max_k);
int kb_comm_get_handle(kb_comm_t c, void* out64);
int kb_comm_open(kb_comm_t c, const void* handles);
int kb_comm_destroy(kb_comm_t c);
int kb_comm_error(kb_comm_t c, int* err); /* 1 if a peer ever failed to arrive (synchronises) */
int kb_ws_set_comm(kb_ws_t ws, kb_comm_t c, int collective);
/* stand-alone fused all-reduce of slot[0..k) (one block) */
int kb_allreduce(kb_ws_t ws, int k, double* slot, void* stream);

/* --- CSR handle ------------------------------------------------------- */
/* rowptr (n_rows+1, int32), colidx/vals (nnz) are device arrays owned by the
 * caller.  padded != 0 promises 16-byte aligned colidx/vals with >= 4 readable
 * elements past nnz (enables the TMA-staged kernel).  Replaces the scipy CSR
 * object the reference multiplies with at _helpers.py:47,61. */
int kb_csr_create(kb_csr_t* h, int64_t n_rows, int64_t n_cols, int64_t nnz,
                  const int32_t* rowptr, const int32_t* colidx, const double* vals,
                  int padded, void* stream);
int kb_csr_destroy(kb_csr_t h);
/* schedule: 0 = auto, 1 = row-wise generic kernel, 2 = TMA-staged stream kernel (k == 1),
 * 3 = TMA-staged stream kernel on offset-pattern compressed indices (k == 1; needs <= 16
 * distinct diagonals col-row and ascending columns, detected by kb_csr_create, which then
 * owns one 16-bit mask per row),
 * 4 = "stencil": schedule 3 with constant diagonals (k == 1; <= 8 diagonals whose stored values
 * are all bitwise equal, tested exactly by kb_csr_create): the <= 8 coefficients travel as kernel
 * parameters and neither indices nor values are streamed.  Bit-identical to the other schedules.
 * 5 = "merge" (k == 1): tiles of equal NONZERO count instead of rows, several lanes per row,
 * long rows finished by a second small launch (csrc/kb_merge.cuh) -- what auto picks for long
 * (mean > 32) or skewed (max > 8 (mean + 8)) rows.  Rows summed by one lane keep csr_matvec's
 * left-to-right order; rows summed by several lanes follow a fixed tree (1e-13 |A||x| of SciPy,
 * bitwise repeatable); kb_tune keys 25 (tile shape), 26 (CTAs per SM), 27 (tile -> CTA order).
 * Schedules 3 and 4 snapshot structure (3) and values (4) at creation: a caller that rewrites
 * colidx / vals in place must create a new handle; schedule 5 snapshots the row pointers at its first product.  One
 * matrix handle must not run products on two streams at once (schedule 5 owns scratch). */
int kb_csr_set_schedule(kb_csr_t h, int schedule);
int kb_csr_get_info(kb_csr_t h, int64_t* n_rows, int64_t* n_cols, int64_t* nnz,
                    int* max_row_len, int* schedule);
/* The offset pattern found by kb_csr_create: *nd distinct diagonals (0 = none), their offsets
 * col - row ascending in offsets16[0..nd), the one value of each in coeffs8 when *constv, and the
 * library-owned per-row masks (bit d set: the row stores diagonal d).  Any output may be NULL. */
int kb_csr_get_stencil(kb_csr_t h, int* nd, int* offsets16, double* coeffs8, int* constv,
                       const uint16_t** masks);

/* --- sparse products: `A @ x` (cg.py:86,180; arnoldi.py:73,176,244;
 *     minres.py:111,121; gmres.py:106) fused with what follows it ---------
 * t = A x, then
 *   mode 0: y = t
 *   mode 1: y = t - coef[c] * z          (Lanczos: arnoldi.py:244-249)
 *   mode 2: y = z - t                    (residual b - A x: cg.py:86)
 * and, in the same pass,
 *   dot 0: nothing     dot 1: out[c] = sum_i w[i,c] * y[i,c]   (cg.py:183, arnoldi.py:160,252)
 *   dot 2: out[c] = sum_i y[i,c]^2                               (cg.py:89, gmres.py:108)
 * x has n_cols rows; y, z, w have n_rows rows. */
int kb_spmv(kb_csr_t A, kb_ws_t ws, int k, const double* x, double* y,
            int mode, const double* z, const double* coef,
            int dot, const double* w, double* out, void* stream);

/* *yes = 1 if kb_spmv with this block width and operand runs the line-marching SpMM
 * (csrc/kb_lines.cuh: k a power of two in [2, 32], 3-D 7-point pattern with constant diagonals,
 * x 16-byte aligned; kb_tune key 16: 0 off, 1 default, 2 wherever valid), else the row-wise kernel. */
int kb_spmm_is_lines(kb_csr_t A, int k, const double* x, int* yes);

/* boundary rows of a row-partitioned matrix, after the halo
 * entries xh have arrived: for i < n_brows, row = rows[i]:
 *   h = sum_j hval[j] * xh[hcol[j], c];  y[row, c] += sign * h
 *   dot 1: out[c] += sum_i w[row, c] * sign * h   (accumulates onto the slot the
 *          local kb_spmv wrote, so one all-reduce covers the whole inner product)
 * halo == NULL: xh is a receive buffer filled by the caller (NCCL send/recv).
 * halo != NULL: xh is ignored; the entries were pushed into this rank's peer-memory
 *          data area by the sources srcs[0..n_src) (device array); the kernel waits for
 *          their flags and acknowledges when it is done (see kb_halo_push). */
int kb_spmv_halo_add(kb_ws_t ws, int k, int64_t n_brows, double sign, const int32_t* rows,
                     const int32_t* hrowptr, const int32_t* hcol, const double* hval,
                     const double* xh, double* y, int dot, const double* w, double* out,
                     kb_halo_t halo, const int* srcs, int n_src, void* stream);
/* Peer-memory halo exchange: one IPC-exported receive area per rank (flags, acks, product
 * counter, data).  kb_halo_push gathers the boundary rows of x listed in idx straight into
 * the destinations' data areas over NVLink and raises their flags -- no NCCL kernel in the
 * iteration.  segs: n_seg x 4 int64 on the device = (destination rank, first entry in idx,
 * number of entries, first row in the destination's data area).  Must be launched by every
 * rank for every product (n_seg may be 0) so the device-resident product counters agree. */
int kb_halo_create(kb_halo_t* h, int rank, int size, int64_t data_bytes);
int kb_halo_get_handle(kb_halo_t h, void* out64);
int kb_halo_open(kb_halo_t h, const void* handles);
int kb_halo_destroy(kb_halo_t h);
int kb_halo_error(kb_halo_t h, int* err);
/* Address of rank's data area as seen from this process (own allocation, or the CUDA-IPC mapping
 * after kb_halo_open): vectors that neighbours write into -- the ghost-extended r of the row-
 * partitioned CG path -- live there. */
int kb_halo_data_ptr(kb_halo_t h, int rank, void** out);
int kb_halo_push(kb_halo_t h, kb_ws_t ws, int k, int n_seg, const int64_t* segs, int64_t n_total,
                 const int32_t* idx, const double* x, void* stream);
/* gather x[idx[i], :] -> buf[i, :] (halo send buffer) */
int kb_pack_rows(kb_ws_t ws, int k, int64_t n_idx, const int32_t* idx, const double* x,
                 double* buf, void* stream);

/* --- reductions: inner(x, y) / norms (_helpers.py:101-110) ------------- */
int kb_dot(kb_ws_t ws, int64_t n, int k, const double* x, const double* y, double* out,
           void* stream);

/* --- CG (cg.py:155-234) ------------------------------------------------ */
/* alpha = rho / nz(pAp [+ pAp2]);  x += alpha p;  r -= alpha Ap;  rr = <r, r>
 * (cg.py:185,196,200,209).  pAp2 may be NULL.  x == NULL defers the x update to the
 * next kb_cg_update_p (what & 4), which reads p anyway: 24 instead of 48 B/element here.
 * alpha_out (nullable) receives alpha: a state slot that only this (gated) kernel writes. */
int kb_cg_update_xr(kb_ws_t ws, int64_t n, int k, const double* rho, const double* pAp,
                    const double* pAp2, const double* p, const double* Ap, double* x,
                    double* r, double* rr_out, double* alpha_out, void* stream);
/* kb_cg_update_xr + the record step (what & 2 below) in the reduction's finishing block:
 * valid where that block holds the final <r, r> (single GPU, or a workspace whose
 * reductions are all-reduced through peer memory) -- one launch less per iteration. */
int kb_cg_update_xr_record(kb_ws_t ws, int64_t n, int k, const double* rho, const double* pAp,
                           const double* p, const double* Ap, double* x, double* r, double* rr_out,
                           double* alpha_out, int step, const double* crit, double* hist,
                           int* stop_at, double* rho_keep, void* stream);
/* what & 2: record resnorm[step] = sqrt(rho_new) into hist[step*k + c], copy rho_new to
 *           rho_keep (state slot, nullable); if all columns satisfy resnorm <= crit[c]
 *           set *stop_at = step (cg.py:156,214-217)
 * what & 4: x += alpha p  -- the previous iteration's deferred update, taken before p
 *           is overwritten (cg.py:196)
 * what & 1: omega = rho_new / nz(rho_old);  p = r + omega p  (cg.py:175-178) */
int kb_cg_update_p(kb_ws_t ws, int64_t n, int k, int step, const double* rho_new,
                   const double* rho_old, const double* alpha, const double* crit, double* hist,
                   int* stop_at, double* rho_keep, const double* r, double* p, double* x, int what,
                   void* stream);

/* Whole-loop entry point (SURVEY.md 8b "kb_cg_solve"): enqueues CG iterations
 * i0 .. i0+n_iters-1 of the fused path on one GPU -- per iteration kb_cg_update_p
 * (i > 0), kb_spmv fused with <p, Ap>, kb_cg_update_xr_record (r update, <r, r>, record) --
 * each gated on *stop_at <= i, with no host involvement between iterations.
 * slots: 6*k doubles = rho ping-pong (rho_i in slot i % 2), alpha, <p,Ap>, <r,r>, scratch.
 * hist row 0 receives step i0+1.  x_pending: the x update of iteration i0-1 is still
 * owed (it is folded into the first p update).  After the call the x update of the
 * last executed iteration is owed (kb_cg_update_p what = 4 flushes it). */
typedef struct {
  kb_csr_t A;
  int64_t n;
  int k;
  double* x;
  double* r;
  double* p;
  double* Ap;
  double* slots;
  const double* crit;
  double* hist;
  int* stop_at;
  /* Optional second search-direction buffer.  When given, k == 1 and A is a 3-D constant-
   * coefficient stencil, an iteration is two launches: the p (and x) update fused with A p and
   * <p, A p> -- it reads p with its halo, so the new p goes to the *other* buffer -- and the r
   * update with A p recomputed on chip.  pcur (0: p, 1: p2) names the buffer that holds the search
   * direction on entry; it changes with every executed iteration i > 0 (the caller derives the
   * new value from the number of executed steps).  p2 == NULL: three launches, p in place. */
  double* p2;
  int pcur;
  /* Row-partitioned two-launch path (one process per GPU, z-slab of a 3-D constant-coefficient
   * stencil; NULL / 0 on one GPU).  masks_ext != NULL: the kernels run on the ghost-extended row
   * space [one plane of the lower neighbour | the n own rows | one plane of the upper neighbour],
   * n_ext = n + 2 * plane rows, own_lo = plane: r, p, p2 are then the bases of EXTENDED buffers
   * (x still of the own rows), masks_ext holds one diagonal mask per extended row (0 on the ghost
   * planes, the global pattern -- neighbour couplings included -- on the own rows), slots has
   * 7*k doubles with slot 6 permanently zero.  The r update stores the new r of the first / last
   * own plane straight into the neighbours' ghost planes (r_push_lo / r_push_hi: peer-mapped
   * addresses, NULL at the ends of the partition); the peer-memory all-reduce that ends the same
   * kernel (kb_ws_set_comm, collective = 1) orders them before the next launch of every rank.
   * The ghost planes of p are maintained locally (p' = r + omega p with the replicated omega), so
   * r is the only vector that crosses NVLink: 2 planes per iteration, no separate exchange kernel.
   * Every iteration (i = 0 too: p, p2 zero-filled on entry) moves p to the other buffer. */
  const uint16_t* masks_ext;
  int64_t n_ext;
  int64_t own_lo;
  double* r_push_lo;
  double* r_push_hi;
} kb_cg_state;
int kb_cg_run(kb_ws_t ws, const kb_cg_state* s, int i0, int n_iters, int x_pending, void* stream);
/* *fused = 1 if kb_cg_run would take the two-launch path for this state (see p2 above). */
int kb_cg_is_fused(const kb_cg_state* s, int* fused);
/* *yes = 1 if kb_cg_run runs the batch as ONE persistent launch (csrc/kb_small.cu: k == 1, one
 * GPU, p2 given, n <= kb_tune key 28 [262144]; two grid-wide barriers per iteration instead of
 * launches).  p moves to the other buffer with every executed iteration i > 0, as on the fused
 * path.  The stopping test breaks the loop on the device; stop_at / hist / slots as above. */
int kb_cg_is_persistent(kb_ws_t ws, const kb_cg_state* s, int* yes);
/* kb_cg_run with CUDA events around every launch (measurement only; synchronises the stream).
 * ms[3]: mean duration of the phases of a step -- fused path: {p/x update + A p + <p,Ap>,
 * r update + <r,r>, 0}; three-kernel path: {p/x update, A p + <p,Ap>, r update + <r,r>}. */
int kb_cg_run_timed(kb_ws_t ws, const kb_cg_state* s, int i0, int n_iters, int x_pending,
                    void* stream, float* ms, float* total_ms);

/* --- per-column scalar arithmetic on device slots ---------------------- */
/* out[c] = A op B with A = a ? a[c] : sa, B = b ? b[c] : sb, c < k.  op: 0 A+B, 1 A-B, 2 A*B,
 * 3 A/B, 4 sqrt(A), 5 |A|, 6 -A, 7 (A != 0 ? A : B), 8 A, 9 A / (B != 0 ? B : sb) (the reference's
 * `x / np.where(y != 0, y, 1)` in one launch; B a device array).  One IEEE operation per launch: the
 * scalar recurrences of the short-recurrence solvers (bicgstab.py:100-133, qmr.py:101-146,
 * symmlq.py:108-150 ...) stay on the device with the host's bits and without a read-back per
 * inner product. */
int kb_scalar_op(kb_ws_t ws, int k, int op, const double* a, const double* b, double sa, double sb,
                 double* out, void* stream);
/* hist[step*k + c] = val[c]; if val[c] <= crit[c] for every column: *stop_at = step.  The loop
 * condition of the short-recurrence drivers (`resnorms[-1] > criterion`) on the device: with the
 * workspace gate, iterations can be enqueued ahead of the host's read-back and become no-ops once
 * an earlier step met the criterion. */
int kb_record(kb_ws_t ws, int k, int step, const double* val, const double* crit, double* hist,
              int* stop_at, void* stream);

/* --- generic vector kernels (fallback path for M/Ml/Mr/custom inner) ---- */
/* y += sign * coef[c] * x   (product rounded, then sum: NumPy temporaries) */
int kb_axpy(kb_ws_t ws, int64_t n, int k, double sign, const double* coef, const double* x,
            double* y, void* stream);
/* out = ca[c] * x + cb[c] * y, every product rounded before the sum (NumPy temporaries).
 * ca == NULL: the first term is x itself; cb == NULL: out = ca * x.  out may alias x or y.
 * Vector statements of the short-recurrence solvers (bicgstab.py:100-133, cgs.py:92-104,
 * qmr.py:101-146, cgr.py:83-84, chebyshev.py:86). */
int kb_lincomb(kb_ws_t ws, int64_t n, int k, const double* ca, const double* x, const double* cb,
               const double* y, double* out, void* stream);
/* y = x + coef[c] * y */
int kb_xpby(kb_ws_t ws, int64_t n, int k, const double* x, const double* coef, double* y,
            void* stream);
/* out = x / nz(coef[c])   (arnoldi.py:147-150,191-193,274-277) */
int kb_div_scale(kb_ws_t ws, int64_t n, int k, const double* x, const double* coef,
                 double* out, void* stream);
/* out = x + y */
int kb_add(kb_ws_t ws, int64_t n, int k, const double* x, const double* y, double* out,
           void* stream);

/* --- Lanczos / MINRES (arnoldi.py:237-281, minres.py:168-236) ---------- */
/* w -= (scale[0]*coef[c]) * u, then out[c] = <z, w> (dot 1) or <w, w> (dot 2) or nothing.
 * scale may be NULL (= 1).  MGS step (arnoldi.py:157-162), Lanczos alpha-step
 * (arnoldi.py:264-267); with scale = beta a Householder reflector (householder.py:62). */
int kb_axpy_dot(kb_ws_t ws, int64_t n, int k, const double* coef, const double* scale,
                const double* u, double* w, int dot, const double* z, double* out, void* stream);
/* per-column state of the MINRES recurrences; all device arrays of k doubles
 * unless noted.  One launch of one block. (minres.py:190-228, givens.py:35-45) */
typedef struct {
  const double* alpha; /* <v, Av>                 (dot slot) */
  const double* ww;    /* <Av, M Av>              (dot slot) */
  double* h2prev;      /* beta_{k-1}; updated to beta_k */
  double* g0;          /* (2k) current rotation  (c, s) */
  double* g1;          /* (2k) previous rotation (c, s) */
  double* y0;          /* running rhs entry */
  double* coefs;       /* (5k) out: R0, R1, R2, y0*, h2 for kb_minres_update */
  const double* crit;
  double* hist;        /* ((maxiter+1) k) */
  int* stop_at;
  int* flags;          /* bit0: invariant subspace */
} kb_minres_state;
int kb_minres_scalar(kb_ws_t ws, int k, int iter, const kb_minres_state* st, void* stream);
/* kb_axpy_dot (dot 2: w -= coef u, st->ww = <w, w>) and kb_minres_scalar(iter) in ONE launch: the
 * scalar recurrences run in the finishing block of the reduction (arnoldi.py:264-267 +
 * minres.py:190-228).  Same arithmetic, one launch less per MINRES step (what kb_minres_run uses). */
int kb_axpy_dot_minres(kb_ws_t ws, int64_t n, int k, const double* coef, const double* u, double* w,
                       int iter, const kb_minres_state* st, void* stream);
/* z = (v - R0 W0 - R1 W1)/nz(R2); W0 <- z (becomes W1 by buffer rotation);
 * yk += y0 z; vnext = Av / nz(h2)   (minres.py:219-221, arnoldi.py:274-277).
 * With a preconditioner M pass MAv = M Av and pnext: vnext = MAv / nz(h2),
 * pnext = Av / nz(h2) (the two bases V = M P); otherwise both NULL. */
int kb_minres_update(kb_ws_t ws, int64_t n, int k, const double* coefs, const double* v,
                     double* W0, const double* W1, const double* Av, double* yk,
                     double* vnext, const double* MAv, double* pnext, void* stream);

/* Whole-loop entry point (SURVEY.md 8b "kb_minres_solve"): enqueues MINRES iterations
 * i0 .. i0+n_iters-1 of the unpreconditioned fused path on one GPU -- per iteration kb_spmv
 * (Lanczos product fused with alpha), kb_axpy_dot_minres (beta^2 + the scalar recurrences in the
 * reduction's finishing block), kb_minres_update: three launches --
 * each gated on *st.stop_at <= i.  V[i % 2] holds v_i (V[(i+1) % 2] = v_{i-1}, overwritten with
 * v_{i+1}), W[i % 2] / W[(i+1) % 2] the two W vectors; st.hist row 0 receives step i0+1 when the
 * caller offsets the pointer as for kb_minres_scalar.  Replaces the loop minres.py:168-236. */
typedef struct {
  kb_csr_t A;
  int64_t n;
  int k;
  double* V[2];
  double* W[2];
  double* Av;
  double* yk;
  kb_minres_state st;
} kb_minres_run_state;
int kb_minres_run(kb_ws_t ws, const kb_minres_run_state* s, int i0, int n_iters, void* stream);

/* --- Arnoldi-MGS / GMRES (arnoldi.py:167-200, gmres.py:179-234) -------- */
typedef struct {
  const double* dots;  /* (num_reorthos*(j+1)*k) MGS coefficients of this step */
  const double* ww;    /* <w, M w> */
  int num_reorthos;
  int maxiter;
  double* R;           /* ((maxiter+1) * maxiter * k) Hessenberg -> R of its QR */
  double* Gc;          /* (maxiter * k) rotation cosines */
  double* Gs;          /* (maxiter * k) rotation sines */
  double* y;           /* ((maxiter+1) * k) rhs of the projected system */
  double* hlast;       /* (k) out: h[j+1] for the normalisation */
  const double* crit;
  double* hist;
  int* stop_at;
  int* flags;
  int have_h;          /* 1: Householder path -- `dots` already holds h[0..j+1] */
} kb_gmres_state;
int kb_gmres_scalar(kb_ws_t ws, int k, int iter, const kb_gmres_state* st, void* stream);
/* kb_axpy_dot (dot 2: w -= coef u, st->ww = <w, w>) and kb_gmres_scalar(iter) in ONE launch (the
 * last projection of an Arnoldi-MGS step, arnoldi.py:157-162,184-185 + gmres.py:199-221); what
 * kb_gmres_cycle uses. */
int kb_axpy_dot_gmres(kb_ws_t ws, int64_t n, int k, const double* coef, const double* u, double* w,
                      int iter, const kb_gmres_state* st, void* stream);
/* Whole-loop entry point (SURVEY.md 8b "kb_gmres_cycle"): enqueues Arnoldi steps i0 ..
 * i0+n_iters-1 of unpreconditioned GMRES with modified Gram-Schmidt (st.num_reorthos sweeps) on
 * one GPU -- per step kb_spmv (w = A V[i] fused with <V[0], w>), (i+1) x sweeps kb_axpy_dot, the
 * last of them kb_axpy_dot_gmres (Givens update of the Hessenberg column, residual norm and stop
 * flag in its finishing block), kb_div_scale (V[i+1] = w / h[i+1]) -- each gated on *st.stop_at <= i.  V[j] = Vbuf + j*vstride
 * must hold i0+n_iters+1 vectors; dots has num_reorthos*(st.maxiter+1)+2 rows of k doubles and
 * must equal st.dots.  Replaces the loop gmres.py:179-234 with arnoldi.py:153-200. */
typedef struct {
  kb_csr_t A;
  int64_t n;
  int k;
  double* Vbuf;
  int64_t vstride;
  double* w;
  double* dots;
  double* ww;
  double* hlast;
  kb_gmres_state st;
} kb_gmres_cycle_state;
int kb_gmres_cycle(kb_ws_t ws, const kb_gmres_cycle_state* s, int i0, int n_iters, void* stream);
/* yy = R[:m,:m]^-1 y[:m] per column (all-zero column -> zeros) (gmres.py:24-38) */
int kb_gmres_solve_y(kb_ws_t ws, int k, int m, int maxiter, const double* R, const double* y,
                     double* yy, void* stream);
/* out = x0 + sum_{j<m} yy[j,c] * V[j]  with V[j] = Vbuf + j*vstride  (gmres.py:96-98) */
int kb_basis_combine(kb_ws_t ws, int64_t n, int k, int m, const double* yy, const double* Vbuf,
                     int64_t vstride, const double* x0, double* out, void* stream);

/* --- Householder (householder.py:26-62, arnoldi.py:65-104), k == 1 ------ */
/* Classical Gram-Schmidt building blocks (ortho="cgs"/"cgs<N>": an additive extension, the
 * reference's Arnoldi is MGS only, arnoldi.py:157-162).  Tall-skinny V^T w in one pass over w:
 * out[j, c] = <V[j][:, c], w[:, c]>, j < cnt; V[j] = V + j * vstride.  (cnt + 1) * 8 B/element. */
int kb_multi_dot(kb_ws_t ws, int64_t n, int k, int cnt, const double* V, int64_t vstride,
                 const double* w, double* out, void* stream);
/* w -= sum_{j < m} h[j, :] * P[j]; dot = 2 also returns out = <w, w> of the result.
 * (m + 2) * 8 B/element. */
int kb_multi_axpy(kb_ws_t ws, int64_t n, int k, int m, const double* h, const double* P,
                  int64_t pstride, double* w, int dot, double* out, void* stream);

/* --- tall-skinny block products on the FP64 tensor cores (utils.py) ------
 * The reference's utils.qr / utils.angles are the one place where `inner` is a true block inner
 * product: inner(QF, QG) -> k x l matrix (utils.py:100,117), followed by n x k times k x l
 * updates (utils.py:101,112,118).  Row-major operands with free leading dimensions (a column
 * sub-block of an (n, k) array is a valid operand); 1 <= k, l <= 16 per call (the host loops over
 * 16-column panels); mma.sync m8n8k4 f64 (DMMA); deterministic summation order.
 *
 * G[i*ldg + j] = sum_r X[r*ldx + i] * Y[r*ldy + j]   (i < k, j < l);  8 n (k + l) bytes.
 * flags bit 0: store sqrt(|.|) instead (norms, utils.py:37).  Gacc != NULL: additionally
 * Gacc[i*ldacc + j] += value (R[j, i] += alpha, utils.py:34).  The workspace must have been
 * created with max_k >= 256. */
int kb_block_gram(kb_ws_t ws, int64_t n, int k, int l, const double* X, int64_t ldx,
                  const double* Y, int64_t ldy, double* G, int64_t ldg, double* Gacc,
                  int64_t ldacc, int flags, void* stream);
/* mode 0: Z = X C;  mode 1: Z = Y - X C;  mode 2: Z = Y + X C.   X: n x k, C: k x l (device),
 * Y, Z: n x l.  Z may alias Y or X.  8 n (k + 2 l) bytes (k + l for mode 0). */
int kb_block_apply(kb_ws_t ws, int64_t n, int k, int l, const double* X, int64_t ldx,
                   const double* C, int64_t ldc, const double* Y, int64_t ldy, double* Z,
                   int64_t ldz, int mode, void* stream);

/* Builds the reflector for the tail x[off:]: v (length n, zeros before off),
 * params[0..3] = alpha, beta, xnorm, sigma2-taken-from-slot.  Two launches. */
int kb_house_make(kb_ws_t ws, int64_t n, int64_t off, const double* x, double* v,
                  double* params, double* scratch, void* stream);
/* Same, with the pivot-sign convention selectable: lapack_sign != 0 treats a zero pivot with a
 * nonzero tail as positive (LAPACK dlarfg, behind np.linalg.qr of utils.py:24): H x = -||x|| e_1. */
int kb_house_make2(kb_ws_t ws, int64_t n, int64_t off, const double* x, double* v,
                   double* params, double* scratch, int lapack_sign, void* stream);
/* h_out[0] = |(w[off] - beta v[off] tau) * alpha|, params = output of kb_house_make,
 * tau = <v, w>  (arnoldi.py:83-85) */
int kb_house_hlast(kb_ws_t ws, const double* w, int64_t off, const double* v, const double* params,
                   const double* tau, double* h_out, void* stream);
/* single-element edits used by the Householder Arnoldi step:
 *  op 0: x[idx] *= s[0];  op 1: x[idx] = val;  op 2: dst[0] = x[idx] */
int kb_poke(kb_ws_t ws, int op, double* x, int64_t idx, const double* s, double val,
            double* dst, void* stream);

/* test hook: out[3i..3i+2] = (c, s, r) of LAPACK dlartg(f[i], g[i])  (givens.py:35-38) */
int kb_lartg(int n, const double* f, const double* g, double* out, void* stream);

/* --- synthetic inputs, built in HBM ---------------------- */
/* rows of the 7-point stencil for planes z_lo <= z < z_hi, global columns.
 * coeffs = {diag, lx, ly, lz, ux, uy, uz}.  Pass 1 (vals == NULL) writes
 * per-row counts into rowptr[1..]; after an inclusive scan by the caller,
 * pass 2 fills colidx/vals. */
int kb_stencil7(int nx, int ny, int nz, int z_lo, int z_hi, const double* coeffs_host,
                int32_t* rowptr, int32_t* colidx, double* vals, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KRYLOV_B200_H */
