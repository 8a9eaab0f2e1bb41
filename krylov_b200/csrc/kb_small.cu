// Whole batches of CG iterations in ONE persistent launch, for problems that are launch-latency
// bound (BASELINE C1: 2-D Poisson 256^2, 65 536 unknowns -- its vectors and its matrix live in
// L2 / L1, a three-launch iteration spends 20 us on launches and drains).  Replaces the loop body
// cg.py:155-234 for k = 1 on one GPU:
//   phase A  x += alpha p (the previous step's deferred update), p' = r + omega p, A p',
//            <p', A p'>          -- A p' is formed from r and the OLD p of the neighbouring rows
//            (p'[c] = r[c] + omega p[c] recomputed per entry: same roundings as the stored p'),
//            so no grid-wide barrier is needed between the p update and the product; p' goes to
//            the other p buffer (ping-pong, as in the fused marching path)
//   barrier 1 = grid-wide sum of <p', A p'>
//   phase B  alpha = rho / nz(<p', A p'>), r -= alpha A p', <r, r>
//   barrier 2 = grid-wide sum of <r, r>; record the residual norm, stopping test (cg.py:156,214)
// Two grid barriers per iteration, no launch.  Rows are walked left to right (csr_matvec order):
// element-wise results equal the three-kernel path's bit for bit; only the summation order of the
// two inner products differs (fixed: block partials in block order), so runs are repeatable.
// The matrix arrays are read through the non-coherent path (L1 hits from the second iteration
// on), r and p -- written by other CTAs -- through L2 (ld.global.cg).
#include "kb_handles.cuh"

int g_small_n = 262144;  // kb_tune key 28: largest n for the persistent CG kernel (0: off)

struct KbSmallCg {
  int n, i0, n_iters, x_pending;
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  double* x;
  double* r;
  double* pb0;
  double* pb1;
  int pcur;
  double* Ap;
  double* slots;
  const double* crit;
  double* hist;
  int* stop_at;
  double* partials;  // 2 sets of gridDim.x {value, flag} entries, one 128-byte line each, cleared
                     // before the launch (entries on one line made every poll queue at one L2 slice)
};

// gpu-scope variant of kb_ll_store (kb_common.cuh uses system scope: peers).
// The entry {lo word, flag, hi word, flag} is written as two 8-byte halves, the first a RELEASE
// store: everything this CTA wrote (ordered before by the block barrier) is visible at gpu scope
// before the flag -- a release, unlike fence.acq_rel / __threadfence(), does not invalidate the
// SM's L1 (the matrix stays cached across iterations)
__device__ __forceinline__ void kb_ll_store_release(double* dst16, double v, unsigned flag) {
  const unsigned long long lo = (unsigned long long)(unsigned)__double2loint(v) |
                                ((unsigned long long)flag << 32);
  const unsigned long long hi = (unsigned long long)(unsigned)__double2hiint(v) |
                                ((unsigned long long)flag << 32);
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(dst16), "l"(lo) : "memory");
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst16 + 1), "l"(hi) : "memory");
}
// Sum over the grid + barrier in one round trip: every CTA stores its block partial as ONE
// self-validating 16-byte entry {value, sequence flag} (kb_ll_store) and a warp polls all the
// entries -- arrival and data travel together, no counter, no second read.  Every CTA adds the
// same partials in the same order (32 strided lanes + butterfly): identical bits everywhere.
// Two slot sets alternate, so a CTA one barrier ahead never overwrites what a slower one still
// reads; the flags are 1, 2, ... within a launch (the slot area is cleared before the launch).
__device__ __forceinline__ double kb_small_allsum(double v, const KbSmallCg& q, unsigned& seq,
                                                  double* sm) {
  const int t = threadIdx.x;
  const double w = kb_warp_sum(v);
  __syncthreads();  // sm free; all global stores of this phase issued
  if ((t & 31) == 0) sm[t >> 5] = w;
  __syncthreads();
  double* slot = q.partials + (size_t)(seq & 1u) * 16 * gridDim.x;  // one 128-byte line per CTA
  const unsigned flag = seq + 1u;
  if (t < 32) {
    if (t == 0) {
      double tot = 0.0;
      const int nw = (blockDim.x + 31) >> 5;
      for (int i = 0; i < nw; ++i) tot += sm[i];
      kb_ll_store_release(slot + 16 * blockIdx.x, tot, flag);
    }
    __syncwarp();
    // lane t polls entries t, t + 32, ...: all its loads are issued before the first flag is
    // tested (one L2 round trip per polling round, however many CTAs there are)
    double xs[KB_BAR_CTAS / 32];
    unsigned need = 0;
#pragma unroll
    for (int i = 0; i < KB_BAR_CTAS / 32; ++i) {
      xs[i] = 0.0;
      if (t + 32 * i < (int)gridDim.x) need |= 1u << i;
    }
    while (need != 0u) {
      unsigned w0[KB_BAR_CTAS / 32], w1[KB_BAR_CTAS / 32], w2[KB_BAR_CTAS / 32], w3[KB_BAR_CTAS / 32];
#pragma unroll
      for (int i = 0; i < KB_BAR_CTAS / 32; ++i)
        if ((need >> i) & 1u)
          asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(w0[i]), "=r"(w1[i]), "=r"(w2[i]), "=r"(w3[i])
                       : "l"(slot + 16 * (size_t)(t + 32 * i))
                       : "memory");
#pragma unroll
      for (int i = 0; i < KB_BAR_CTAS / 32; ++i)
        if (((need >> i) & 1u) && w1[i] == flag && w3[i] == flag) {
          xs[i] = __hiloint2double((int)w2[i], (int)w0[i]);
          need &= ~(1u << i);
        }
    }
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < KB_BAR_CTAS / 32; ++i) a += xs[i];  // block order within the lane
    a = kb_warp_sum(a);
    if (t == 0) sm[32] = a;
  }
  __syncthreads();
  ++seq;
  return sm[32];
}

// ONE: every thread owns at most one row for the whole launch -- x, r, p and A p of that row
// stay in registers; only p' and r go to memory (the neighbouring rows read them).
#define KB_SMALL_BLOCK 512  // 128 registers per thread: the row's entries + its gathers in flight
#define KB_SMALL_ROW 8      // entries of a row kept in registers (longer rows: loop over memory)

template <bool ONE>
__global__ void __launch_bounds__(KB_SMALL_BLOCK, 1) kb_cg_small_kernel(KbSmallCg q) {
  __shared__ double sm[40];
  const int nthreads = gridDim.x * blockDim.x;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned seq = 0;
  int pc = q.pcur;
  int x_pending = q.x_pending;
  // loop-carried scalars (identical on every thread): rho_i, rho_{i-1}, alpha_{i-1}
  double rho = q.slots[q.i0 % 2], rho_old = q.slots[(q.i0 + 1) % 2], alpha_prev = q.slots[2];
  const bool mine = gtid < q.n;
  double x_own = 0.0, r_own = 0.0, p_own = 0.0, ap_own = 0.0;
  int lo1 = 0, hi1 = 0;
  // ONE: a row of at most KB_SMALL_ROW entries lives in registers for the whole launch, so that a
  // step's gathers of r and p are issued together (one L2 round trip instead of one per entry)
  int cidx[KB_SMALL_ROW];
  double cval[KB_SMALL_ROW];
  bool inreg = false;
  if (ONE && mine) {
    x_own = q.x[gtid];
    r_own = q.r[gtid];
    p_own = (pc ? q.pb1 : q.pb0)[gtid];
    lo1 = q.rowptr[gtid];
    hi1 = q.rowptr[gtid + 1];
    inreg = hi1 - lo1 <= KB_SMALL_ROW;
#pragma unroll
    for (int j = 0; j < KB_SMALL_ROW; ++j) {
      const bool in = inreg && lo1 + j < hi1;
      cidx[j] = in ? q.colidx[lo1 + j] : 0;
      cval[j] = in ? q.vals[lo1 + j] : 0.0;
    }
  }
  const int len1 = hi1 - lo1;
  for (int i = q.i0; i < q.i0 + q.n_iters; ++i) {
    const double* p_old = pc ? q.pb1 : q.pb0;
    double* p_new = (i > 0) == (pc != 0) ? q.pb0 : q.pb1;
    const double omega = i > 0 ? rho / kb_nz(rho_old) : 0.0;
    // ---------------- phase A
    double acc = 0.0;
    if (ONE) {
      if (mine) {
        double sum = 0.0;
        if (i > 0) {
          if (x_pending) x_own = kb_mul_add(alpha_prev, p_own, x_own);
          p_own = kb_mul_add(omega, p_own, r_own);
          p_new[gtid] = p_own;
        }
        if (inreg) {
          double pg[KB_SMALL_ROW], rg[KB_SMALL_ROW];
#pragma unroll
          for (int j = 0; j < KB_SMALL_ROW; ++j) {
            pg[j] = rg[j] = 0.0;
            if (j < len1) {
              pg[j] = __ldcg(p_old + cidx[j]);
              if (i > 0) rg[j] = __ldcg(q.r + cidx[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < KB_SMALL_ROW; ++j) {
            if (j < len1) {  // same products, same order as the loop below
              const double pcn = i > 0 ? kb_mul_add(omega, pg[j], rg[j]) : pg[j];
              sum = __dadd_rn(sum, __dmul_rn(cval[j], pcn));
            }
          }
        } else if (i > 0) {
          for (int j = lo1; j < hi1; ++j) {
            const int c = __ldg(q.colidx + j);
            const double pcn = kb_mul_add(omega, __ldcg(p_old + c), __ldcg(q.r + c));
            sum = __dadd_rn(sum, __dmul_rn(__ldg(q.vals + j), pcn));
          }
        } else {
          for (int j = lo1; j < hi1; ++j)
            sum = __dadd_rn(sum, __dmul_rn(__ldg(q.vals + j), __ldcg(p_old + __ldg(q.colidx + j))));
        }
        ap_own = sum;
        acc = p_own * sum;
      }
    } else {
      for (int row = gtid; row < q.n; row += nthreads) {
        const int lo = __ldg(q.rowptr + row), hi = __ldg(q.rowptr + row + 1);
        double pn, sum = 0.0;
        if (i > 0) {
          const double po = __ldcg(p_old + row);
          if (x_pending) q.x[row] = kb_mul_add(alpha_prev, po, q.x[row]);
          pn = kb_mul_add(omega, po, __ldcg(q.r + row));
          p_new[row] = pn;
          for (int j = lo; j < hi; ++j) {
            const int c = __ldg(q.colidx + j);
            const double pcn = kb_mul_add(omega, __ldcg(p_old + c), __ldcg(q.r + c));
            sum = __dadd_rn(sum, __dmul_rn(__ldg(q.vals + j), pcn));
          }
        } else {
          pn = __ldcg(p_old + row);
          for (int j = lo; j < hi; ++j)
            sum = __dadd_rn(sum, __dmul_rn(__ldg(q.vals + j), __ldcg(p_old + __ldg(q.colidx + j))));
        }
        q.Ap[row] = sum;
        acc = fma(pn, sum, acc);
      }
    }
    if (i > 0) pc ^= 1;
    const double pAp = kb_small_allsum(acc, q, seq, sm);
    // ---------------- phase B
    const double alpha = rho / kb_nz(pAp);
    acc = 0.0;
    if (ONE) {
      if (mine) {
        r_own = kb_mul_sub(alpha, ap_own, r_own);
        __stcg(q.r + gtid, r_own);
        acc = r_own * r_own;
      }
    } else {
      for (int row = gtid; row < q.n; row += nthreads) {
        const double rn = kb_mul_sub(alpha, q.Ap[row], __ldcg(q.r + row));
        __stcg(q.r + row, rn);
        acc = fma(rn, rn, acc);
      }
    }
    const double rr = kb_small_allsum(acc, q, seq, sm);
    x_pending = 1;
    const double nrm = sqrt(rr);
    const bool stop = nrm <= q.crit[0];
    if (gtid == 0) {  // state for the host and for the next batch
      q.slots[(i + 1) % 2] = rr;  // rho_{i+1}
      q.slots[2] = alpha;
      q.slots[3] = pAp;
      q.slots[4] = rr;
      q.hist[i - q.i0] = nrm;
      if (stop) *q.stop_at = i + 1;
    }
    if (stop) break;
    rho_old = rho;
    rho = rr;
    alpha_prev = alpha;
  }
  if (ONE && mine) {  // x of the own row (the last step's alpha p is still owed, as on every path)
    q.x[gtid] = x_own;
    q.Ap[gtid] = ap_own;
  }
}

// true if kb_cg_run may take the persistent kernel for this state
bool kb_cg_small_ok(const kb_ws_s* ws, const kb_cg_state* s) {
  return g_small_n > 0 && s->k == 1 && s->p2 != nullptr && s->masks_ext == nullptr &&
         !(ws->comm != nullptr && ws->collective) && s->n >= 1 && s->n <= g_small_n &&
         s->A->n_rows == s->n && s->A->n_cols == s->n && s->A->nnz > 0;
}

int kb_cg_small_run(kb_ws_s* ws, const kb_cg_state* s, int i0, int n_iters, int x_pending,
                    cudaStream_t st) {
  if (n_iters <= 0) return KB_OK;
  static int resident[64] = {0};
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  KB_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  if (resident[dev] == 0) {
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident[dev], kb_cg_small_kernel<false>,
                                                          KB_SMALL_BLOCK, 0));
    KB_REQUIRE(resident[dev] >= 1, "persistent CG kernel does not fit an SM");
  }
  KbSmallCg q;
  q.n = (int)s->n;
  q.i0 = i0;
  q.n_iters = n_iters;
  q.x_pending = x_pending;
  q.rowptr = s->A->rowptr;
  q.colidx = s->A->colidx;
  q.vals = s->A->vals;
  q.x = s->x;
  q.r = s->r;
  q.pb0 = s->p;
  q.pb1 = s->p2;
  q.pcur = s->pcur;
  q.Ap = s->Ap;
  q.slots = s->slots;
  q.crit = s->crit;
  q.hist = s->hist;
  q.stop_at = s->stop_at;
  q.partials = ws->barbuf;
  // one row per thread while the rows fit co-resident CTAs; all CTAs must be resident (they meet
  // at grid-wide barriers): cooperative launch
  int grid = (int)((s->n + KB_SMALL_BLOCK - 1) / KB_SMALL_BLOCK);
  const int cap = ws->num_sms * resident[dev];
  if (grid > cap) grid = cap;
  if (grid > KB_BAR_CTAS) grid = KB_BAR_CTAS;
  KB_CUDA(cudaMemsetAsync(q.partials, 0, 2 * 128 * (size_t)grid, st));
  void* args[] = {&q};
  const bool one = (int64_t)grid * KB_SMALL_BLOCK >= s->n;
  KB_CUDA(cudaLaunchCooperativeKernel(
      one ? (const void*)kb_cg_small_kernel<true> : (const void*)kb_cg_small_kernel<false>,
      dim3(grid), dim3(KB_SMALL_BLOCK), args, 0, st));
  return KB_OK;
}
