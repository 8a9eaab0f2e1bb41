// Line-marching SpMM for constant-coefficient 3-D stencils and blocked right-hand sides (sm_100a).
//
// Y = A X with X of shape (n, k), row-major, is the banded product of A (x) I_k with the flattened
// vector: row e = i k + c has the diagonals e + k off[d].  For the 7-point stencil these are
//   {-Pz, -L, -kI, 0, +kI, +L, +Pz},   L = nx k (one grid line),  Pz = nx ny k (one plane).
// The row-wise SpMM gathers seven 8 k-byte rows of X per row through L1/L2: 7 x 128 B at k = 16,
// and runs into the L2 -> SM limit at 0.36 of the HBM roofline (profiles/r1_c4_tune.txt: 2.50 ms
// at 256^3, k = 16).  The plane-marching kernel of k = 1 does not carry over: its ring would hold
// windows of 2 L + tile entries, 72 KB each at k = 16.
//
// Here a CTA owns a chunk of TR = 1024 (or 512) consecutive entries of a line and *marches over the
// lines*.  One ring slot holds everything the flattened vector contributes at line r:
//     [ x[r L + c TR - H, ... + TR + H) | x[. - Pz, TR entries) | x[. + Pz, TR entries) ]
// (H >= k |off of the innermost pair|: the +-kI neighbours; three 1-D TMA bulk copies, one
// mbarrier).  Computing line r reads the -L diagonal from the previous slot, 0 / +-kI / +-Pz from
// the current one and +L from the next: (TR + 2H) / TR + 2 = 3.03 entries per row cross the
// L2 -> SM boundary instead of 7, DRAM sees X once (the +-Pz chunks were brought in by the CTAs
// one plane away and are L2 hits: 126 MB of L2 against 8 MB planes).
// Products, their order and rounding are those of every other schedule (kb_march_rows): results
// are bit-identical to the row-wise kernel.
// VAR = true: the same ring for matrices with this offset pattern and VARIABLE coefficients (schedule
// "pattern"): the stored values of a chunk's rows (one contiguous range of the CSR value array)
// ride in the slot of their line as a fourth TMA piece, the consumers read them from shared memory
// (k lanes share a row: broadcast).  Streams 8 nnz more bytes per product, same x / y traffic.  Rows next to a boundary skip absent diagonals through
// the per-row masks, so any matrix with this offset pattern and constant diagonals qualifies
// (Dirichlet gaps, truncated last plane, row slabs of a partitioned matrix).
#pragma once
#include "kb_march.cuh"

struct KbLines {
  long long L;       // line stride in flattened entries
  long long Pz;      // plane stride in flattened entries
  long long N;       // n_rows * k
  long long Nx;      // n_cols * k (length of the flattened x)
  long long nlines;  // ceil(N / L)
  long long nitems;  // ncol * ceil(nlines / ch)
  int k, kshift;     // block width (power of two), log2
  int H;             // halo entries on each side of a chunk's window (even, >= k * off[4])
  int inner;         // k * off[4]
  int slotlen;       // 3 TR + 2 H doubles
  int ncol;          // chunks per line
  int ch;            // lines per work item
  int TR;            // entries per chunk (256 x entries per thread)
  // work-item order.  lpp > 0: planes fastest -- item = ((group in plane) * ncol + chunk) * nplanes
  // + plane, so that CTAs running side by side hold the same chunk of neighbouring planes and the
  // +-Pz chunks are L2 hits; lpp == 0: natural order (line groups, then chunks).
  long long lpp;     // lines per plane (Pz / L), 0 = natural order
  long long nplanes; // ceil(nlines / lpp)
  int tail;          // entries of the last line (N - (nlines - 1) L)
  // variable coefficients (VAR): the values of a chunk's rows travel in the same ring slot
  int voff;          // offset of the value area in a slot (doubles) = 3 TR + 2 H
  int vcap;          // its capacity (doubles, even): (TR / k) rows x 7 + 2
  long long n_rows;
};

// lines [r0, r1) and chunk c of a work item; false for an empty item (plane-fastest order only)
__device__ __forceinline__ bool kb_lines_item(const KbLines& g, long long item, int& c,
                                              long long& r0, long long& r1) {
  if (g.lpp == 0) {
    c = (int)(item % g.ncol);
    r0 = (item / g.ncol) * g.ch;
    r1 = (r0 + g.ch < g.nlines) ? r0 + g.ch : g.nlines;
    return true;
  }
  const long long zp = item % g.nplanes;
  const long long gc = item / g.nplanes;
  c = (int)(gc % g.ncol);
  const long long gi = gc / g.ncol;
  r0 = zp * g.lpp + gi * g.ch;
  long long e = r0 + g.ch;
  const long long pe = (zp + 1) * g.lpp;
  e = e < pe ? e : pe;
  r1 = e < g.nlines ? e : g.nlines;
  return r1 > r0;
}

// Row sums of two entries with the coefficients read from shared memory (VAR): va[q] is the shared
// address of the first stored value of entry q's row; the values follow in stored (ascending
// column) order, so a full row uses va + 8 d and a boundary row the rank of d among its diagonals.
template <int Q0>
__device__ __forceinline__ void kb_lines_rows_var(const uint32_t (&a)[7], const unsigned* m,
                                                  const uint32_t* va, double* sum, double* ctr) {
  const bool allfull = m[Q0] == 127u && m[Q0 + 1] == 127u;
  if (__all_sync(0xffffffffu, allfull)) {
    double xa[7], xb[7], ca[7], cb[7];
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      xa[d] = kb_lds_f64<Q0 * 256 * 8>(a[d]);
      xb[d] = kb_lds_f64<(Q0 + 1) * 256 * 8>(a[d]);
    }
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ca[d]) : "r"(va[Q0] + 8u * d));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(cb[d]) : "r"(va[Q0 + 1] + 8u * d));
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      s0 = __dadd_rn(s0, __dmul_rn(ca[d], xa[d]));
      s1 = __dadd_rn(s1, __dmul_rn(cb[d], xb[d]));
    }
    sum[Q0] = s0;
    sum[Q0 + 1] = s1;
    ctr[Q0] = xa[3];
    ctr[Q0 + 1] = xb[3];
  } else {
    double s0 = 0.0, s1 = 0.0;
    ctr[Q0] = kb_lds_f64<Q0 * 256 * 8>(a[3]);
    ctr[Q0 + 1] = kb_lds_f64<(Q0 + 1) * 256 * 8>(a[3]);
    uint32_t v0 = va[Q0], v1 = va[Q0 + 1];
#pragma unroll
    for (int d = 0; d < 7; ++d) {
      if ((m[Q0] >> d) & 1u) {
        double c;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c) : "r"(v0));
        v0 += 8u;
        s0 = __dadd_rn(s0, __dmul_rn(c, kb_lds_f64<Q0 * 256 * 8>(a[d])));
      }
      if ((m[Q0 + 1] >> d) & 1u) {
        double c;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(c) : "r"(v1));
        v1 += 8u;
        s1 = __dadd_rn(s1, __dmul_rn(c, kb_lds_f64<(Q0 + 1) * 256 * 8>(a[d])));
      }
    }
    sum[Q0] = s0;
    sum[Q0 + 1] = s1;
  }
}

// first / one-past-last stored value of the rows of chunk c of line `line` (VAR); b <= a: none
__device__ __forceinline__ void kb_lines_vrange(const KbLines& g, const int32_t* __restrict__ rowptr,
                                                long long line, int c, int TR, int& a, int& b) {
  a = b = 0;
  if (line < 0 || line >= g.nlines) return;
  const long long lim = line == g.nlines - 1 ? g.tail : g.L;
  const long long p0 = (long long)c * TR;
  if (p0 >= lim) return;
  const long long p1 = p0 + TR < lim ? p0 + TR : lim;
  const long long ra = (line * g.L + p0) >> g.kshift;
  long long rb = (line * g.L + p1) >> g.kshift;
  if (rb > g.n_rows) rb = g.n_rows;
  if (ra >= rb) return;
  a = rowptr[ra];
  b = rowptr[rb];
}

template <int RPT, int NS, int MINB, int DOT, bool WX, bool VAR = false>
__global__ void __launch_bounds__(256 + 32, MINB)
kb_spmm_lines_kernel(KbLines g, const uint16_t* __restrict__ masks, KbConstVals cv,
                     const int32_t* __restrict__ rowptr, const double* __restrict__ vals,
                     const double* __restrict__ x, double* __restrict__ y, int mode,
                     const double* __restrict__ z, const double* __restrict__ coef,
                     const double* __restrict__ w, double* __restrict__ out, KbRed rd) {
  static_assert(RPT == 2 || RPT == 4, "kb_march_rows works on pairs of entries");
  static_assert(NS >= 4, "ring: previous, current, next line + one in flight");
  if (kb_gated(rd)) return;
  constexpr int TR = 256 * RPT;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  double* const s_win = reinterpret_cast<double*>(kb_dyn_smem);
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(s_win + (size_t)NS * g.slotlen);
  uint64_t* const s_empty = s_full + NS;
  __shared__ double red_sm[256 + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      kb_mbar_init(&s_full[s], 1);
      kb_mbar_init(&s_empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;
  if (warp == 8) {
    // ---------------------------------------------- producer: one slot per line ---
    // Lane 0 issues the copies.  VAR: all 32 lanes fetch the value ranges of the item's next 32
    // lines at once (one round of global latency per 32 slots instead of one per slot).
    {
      const int lane = tid & 31;
      const uint64_t pol = kb_policy_evict_last();
      const uint64_t polv = kb_policy_evict_first();
      unsigned cnt = 0;
      for (long long item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        int c;
        long long r0, r1;
        if (!kb_lines_item(g, item, c, r0, r1)) continue;  // warp-uniform
        const int nload = (int)(r1 - r0) + 2;
        int my_a = 0, my_b = 0;
        for (int l = 0; l < nload; ++l, ++cnt) {
          int va = 0, vb = 0;
          if (VAR) {
            const bool inside_l = l >= 1 && l <= nload - 2;
            if (inside_l && ((l - 1) & 31) == 0) {  // lines r0 + (l-1) ... + 31 of this item
              my_a = my_b = 0;
              if (l - 1 + lane < nload - 2)
                kb_lines_vrange(g, rowptr, r0 + (l - 1) + lane, c, TR, my_a, my_b);
            }
            va = __shfl_sync(0xffffffffu, my_a, (l - 1) & 31);
            vb = __shfl_sync(0xffffffffu, my_b, (l - 1) & 31);
            if (!inside_l) va = vb = 0;
          }
          if (lane != 0) continue;
          const int slot = (int)(cnt % NS);
          const unsigned use = cnt / NS;
          if (use > 0) kb_mbar_wait(&s_empty[slot], (use - 1u) & 1u);
          double* const sl = s_win + (size_t)slot * g.slotlen;
          const long long e0 = (r0 - 1 + l) * g.L + (long long)c * TR;  // first own entry
          // pieces: the window with its halo; the -Pz and +Pz chunks (lines inside the item only)
          long long lo[3], hi[3];
          double* dst[3];
          const bool inside = l >= 1 && l <= nload - 2;
          const long long s0[3] = {e0 - g.H, e0 - g.Pz, e0 + g.Pz};
          const long long len[3] = {(long long)TR + 2 * g.H, inside ? TR : 0, inside ? TR : 0};
          double* const base[3] = {sl, sl + TR + 2 * g.H, sl + 2 * TR + 2 * g.H};
          uint32_t total = 0;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            lo[p] = s0[p] > 0 ? s0[p] : 0;
            hi[p] = (s0[p] + len[p] < g.Nx) ? s0[p] + len[p] : g.Nx;
            dst[p] = base[p] + (lo[p] - s0[p]);
            if (hi[p] > lo[p]) total += (uint32_t)(hi[p] - lo[p]) * 8u;
          }
          // VAR: the stored values of this line's rows, from an even index (16-byte source)
          const int va0 = va & ~1;
          const uint32_t vbytes = vb > va ? (uint32_t)(((vb - va0) + 1) & ~1) * 8u : 0u;
          total += vbytes;
          if (total > 0) {
            kb_mbar_expect_tx(&s_full[slot], total);
            if (vbytes > 0)
              kb_bulk_g2s_hint(sl + g.voff, vals + va0, vbytes, &s_full[slot], polv);
#pragma unroll
            for (int p = 0; p < 3; ++p)
              if (hi[p] > lo[p])
                kb_bulk_g2s_hint(dst[p], x + lo[p], (uint32_t)(hi[p] - lo[p]) * 8u, &s_full[slot],
                                 pol);
          } else {
            kb_mbar_arrive(&s_full[slot]);  // nothing of this line exists
          }
        }
      }
    }
  } else {
    // ---------------------------------------------- consumers: RPT entries per thread ---
    const uint32_t sbase = kb_smem_u32(s_win);
    const uint32_t sbytes = (uint32_t)g.slotlen * 8u;
    const uint32_t own = (uint32_t)(g.H + tid) * 8u;           // entry q = 0 in a slot's window
    const uint32_t olo = (uint32_t)(TR + 2 * g.H + tid) * 8u;  // ... in its -Pz chunk
    const uint32_t ohi = olo + (uint32_t)TR * 8u;              // ... in its +Pz chunk
    const uint32_t oin = (uint32_t)g.inner * 8u;
    const int col = tid & (g.k - 1);  // 256 % k == 0: the same column for every q
    double cf = 0.0;
    if (mode == 1) cf = coef[col];
    unsigned cnt = 0;
    const int mq = 256 >> g.kshift;  // mask entries between consecutive q (256 % k == 0)
    for (long long item = blockIdx.x; item < g.nitems; item += gridDim.x) {
      int c;
      long long r0, r1;
      if (!kb_lines_item(g, item, c, r0, r1)) continue;
      const int nload = (int)(r1 - r0) + 2;
      const int pos0 = c * TR + tid;  // in-line position of entry q = 0
      // Running index of entry q = 0 of the line computed in the current pass; everything per
      // entry is an immediate offset (q * 256) from it -- no 64-bit arithmetic per entry and pass.
      long long e = (r0 - 2) * g.L + pos0;
      unsigned mn[RPT];
      double zn[RPT], wn[RPT];
      int rpn[RPT], a0n = 0;  // VAR: first stored value of the entry's row / of the chunk's rows
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        mn[q] = 0u;
        zn[q] = wn[q] = 0.0;
        rpn[q] = 0;
      }
      for (int l = 0; l < nload; ++l, ++cnt, e += g.L) {
        const int slot = (int)(cnt % NS);
        const long long rc = r0 + l - 2;  // line computed in this pass (l >= 2)
        // valid positions of this line and of the next one (the last line may be short)
        const int limc = l >= 2 ? (rc == g.nlines - 1 ? g.tail : (int)g.L) : 0;
        const int limn = (l >= 1 && l + 1 < nload) ? (rc + 1 == g.nlines - 1 ? g.tail : (int)g.L) : 0;
        unsigned m[RPT];
        double zv[RPT], wv[RPT];
        bool okq[RPT];
        int rp[RPT];
        const int a0 = a0n;
        const long long en = e + g.L;
        const uint16_t* const mp = masks + (en >> g.kshift);
        if (VAR) {
          a0n = 0;
          if (limn > c * TR) a0n = rowptr[(en - tid) >> g.kshift] & ~1;  // row of the chunk's entry 0
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          const int pos = pos0 + q * 256;
          okq[q] = pos < limc;
          m[q] = mn[q];
          zv[q] = zn[q];
          wv[q] = wn[q];
          if (VAR) rp[q] = rpn[q];
          // operands of the next pass (line rc + 1) are requested one pass ahead
          const bool nok = pos < limn;
          mn[q] = nok ? (unsigned)mp[q * mq] : 0u;
          zn[q] = wn[q] = 0.0;
          if (VAR) rpn[q] = nok ? rowptr[(en >> g.kshift) + q * mq] : 0;
          if (nok) {
            if (mode != 0) zn[q] = z[en + q * 256];
            if (DOT == 1 && !WX) wn[q] = w[en + q * 256];
          }
        }
        kb_mbar_wait(&s_full[slot], (cnt / NS) & 1u);
        if (l >= 2) {
          const uint32_t wprev = sbase + (uint32_t)((cnt - 2u) % NS) * sbytes;
          const uint32_t wcur = sbase + (uint32_t)((cnt - 1u) % NS) * sbytes;
          const uint32_t wnext = sbase + (uint32_t)slot * sbytes;
          uint32_t a[7];
          a[0] = wcur + olo;
          a[1] = wprev + own;
          a[2] = wcur + own - oin;
          a[3] = wcur + own;
          a[4] = wcur + own + oin;
          a[5] = wnext + own;
          a[6] = wcur + ohi;
          double sum[RPT], ctr[RPT];
          if constexpr (VAR) {
            uint32_t va[RPT];
#pragma unroll
            for (int q = 0; q < RPT; ++q)
              va[q] = wcur + (uint32_t)(g.voff + (okq[q] ? rp[q] - a0 : 0)) * 8u;
            kb_lines_rows_var<0>(a, m, va, sum, ctr);
            if constexpr (RPT == 4) kb_lines_rows_var<2>(a, m, va, sum, ctr);
          } else {
            kb_march_rows<7, 0, 2>(a, m, cv, sum, ctr);
            if constexpr (RPT == 4) kb_march_rows<7, 2, 2>(a, m, cv, sum, ctr);
          }
#pragma unroll
          for (int q = 0; q < RPT; ++q) {
            if (okq[q]) {
              double yv = sum[q];
              if (mode == 1) yv = kb_mul_sub(cf, zv[q], sum[q]);
              if (mode == 2) yv = __dsub_rn(zv[q], sum[q]);
              __stcs(&y[e + q * 256], yv);
              if (DOT == 1) acc = fma(WX ? ctr[q] : wv[q], yv, acc);
              if (DOT == 2) acc = fma(yv, yv, acc);
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) kb_mbar_arrive(&s_empty[(cnt - 2u) % NS]);
        }
      }
      // the last two slots of the item are not needed by a later pass
      __syncwarp();
      if ((tid & 31) == 0) {
        kb_mbar_arrive(&s_empty[(cnt - 2u) % NS]);
        kb_mbar_arrive(&s_empty[(cnt - 1u) % NS]);
      }
    }
  }
  if (DOT == 0) return;
  kb_grid_colsum(acc, g.k, rd, out, red_sm);
}
