// Plane-marching kernel for constant-coefficient 3-D stencils (sm_100a).
//
// kb_spmv_stencil2_kernel brings every x entry into shared memory once per window that holds
// it: five windows for the 7-point stencil, 40 B of L2->SM traffic per row.  ncu at 512^3
// (profiles/r1_stencil2_ncu.txt): 5.9 GB over the crossbar in 0.76 ms = ~80 % of the measured
// L2 throughput cap while DRAM idles at 46 % -- the kernel is L2-bound, and half of the stall
// samples sit on the "window landed?" barrier.
//
// Here a CTA owns a tile of TR consecutive in-plane positions and *marches* through the planes
// (row = plane * P + position, P = the largest diagonal offset):  the window
//     W_j = x[j P + c TR - LP,  j P + c TR + TR + LP)
// serves the diagonals |off| <= LP of plane j, the +P diagonal of plane j-1 and the -P diagonal
// of plane j+1.  A ring of NS windows stays in shared memory, one 1-D TMA bulk copy per plane
// brings the next one: (TR + 2 LP) / TR entries per row instead of 5 (2 at TR = 1024, 512-wide
// lines).  Products, their order and rounding are those of every other schedule -> bit-identical.
//
// The same skeleton carries the two halves of a fused CG iteration (KIND 1 / 2): because the
// neighbouring planes of the *updated* search direction are already in the ring,
//     p <- r + omega p,  [x <- x + alpha_old p_old,]  <p, A p>          (KIND 1)
// needs no separate pass over p, r, x, and
//     r <- r - alpha (A p),  <r, r>, record / stopping test               (KIND 2)
// recomputes A p from the ring instead of reading a stored copy: a CG step streams
// 40 n + 24 n (+ masks) bytes instead of 40 n + 18 n + 24 n  (cg.py:175-217).
#pragma once
#include "kb_spmv.cuh"
#include "kb_vec.cuh"

struct KbMarch {
  int P;        // plane stride = largest diagonal offset (even)
  int LP;       // halo entries on each side of a tile's window (multiple of 256, >= inner offsets)
  int wlen;     // TR + 2 LP
  int ncol;     // tiles per plane
  int nplanes;  // ceil(n_rows / P)
  int ch;       // planes per work item
  int nitems;   // ncol * ceil(nplanes / ch)
};

// operands of the fused CG kernels (KIND 1 / 2)
struct KbMarchCg {
  const double* rho_a;   // KIND 1: rho_i (omega = rho_a / nz(rho_b));  KIND 2: rho_i
  const double* rho_b;   // KIND 1: rho_{i-1};                           KIND 2: <p, Ap>
  const double* alpha_in;  // KIND 1: alpha of the previous iteration (x update)
  double* alpha_out;       // KIND 2: state slot of alpha
  const double* r_in;      // KIND 1: r (read with halo)
  double* r;               // KIND 2: r (updated in place, own rows)
  double* xv;              // KIND 1: x (own rows), may be null: no x update
  double* p_out;           // KIND 1: the new search direction (a buffer different from p_in)
  KbCgRecord rec;          // KIND 2
  // Row-partitioned problems run the same kernels on a ghost-extended row space: one plane of the
  // lower neighbour, the rank's own planes, one plane of the upper neighbour (masks of the ghost
  // rows are 0).  Rows [own_lo, own_hi) are this rank's: only they enter the dots and update x
  // and r; p_out is written for every row (the ghost planes of p are kept up to date locally:
  // p' = r + omega p with the neighbours' r and the replicated omega).  KIND 2 also stores the
  // new r of its first / last own plane into the neighbours' ghost planes (NVLink peer stores);
  // the all-reduce that ends the kernel is what tells the neighbours that they have landed.
  int own_lo, own_hi;
  int own_pl0, own_pl1;    // the same range in planes: own_lo / P, own_hi / P
  double* push_lo;         // KIND 2: the lower neighbour's upper ghost plane of r (peer-mapped) or null
  double* push_hi;         // KIND 2: the upper neighbour's lower ghost plane of r or null
};

// mbarrier operations on a precomputed shared-window address (no generic -> shared conversion in
// the marching loop)
__device__ __forceinline__ void kb_mbar_arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void kb_mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "KB_WAITU_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra KB_DONEU_%=;\n"
      "bra KB_WAITU_%=;\n"
      "KB_DONEU_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void kb_st_f64_shared(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double kb_ld_f64_shared(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// row sums of NQ (1 or 2) rows: q = Q0, Q0 + 1 of this thread.  a[d]: shared address of
// diagonal d for row q = 0 in the current ring position.
template <int ND, int Q0, int NQ, int CT = 256>
__device__ __forceinline__ void kb_march_rows(const uint32_t (&a)[ND], const unsigned* m,
                                              const KbConstVals& cv, double* sum, double* ctr) {
  constexpr unsigned full = (1u << ND) - 1u;
  bool allfull = m[Q0] == full;
  if constexpr (NQ == 2) allfull = allfull && m[Q0 + 1] == full;
  if (__all_sync(0xffffffffu, allfull)) {
    double xa[ND], xb[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d) xa[d] = kb_lds_f64<Q0 * CT * 8>(a[d]);
    if constexpr (NQ == 2) {
#pragma unroll
      for (int d = 0; d < ND; ++d) xb[d] = kb_lds_f64<(Q0 + 1) * CT * 8>(a[d]);
    }
    sum[Q0] = kb_st2_sum<ND>(xa, cv);
    ctr[Q0] = xa[ND / 2];
    if constexpr (NQ == 2) {
      sum[Q0 + 1] = kb_st2_sum<ND>(xb, cv);
      ctr[Q0 + 1] = xb[ND / 2];
    }
  } else {  // rows next to a boundary: the absent diagonals are skipped, never loaded
    // the middle entry is the row's own x (dot operand when w aliases x): always inside the window
    double s0 = 0.0, s1 = 0.0, c1 = 0.0;
    const double c0 = kb_lds_f64<Q0 * CT * 8>(a[ND / 2]);
    if constexpr (NQ == 2) c1 = kb_lds_f64<(Q0 + 1) * CT * 8>(a[ND / 2]);
#pragma unroll
    for (int d = 0; d < ND; ++d) {
      if ((m[Q0] >> d) & 1u) {
        const double v = kb_lds_f64<Q0 * CT * 8>(a[d]);
        s0 = __dadd_rn(s0, __dmul_rn(cv.c[d], v));
      }
      if constexpr (NQ == 2) {
        if ((m[Q0 + 1] >> d) & 1u) {
          const double v = kb_lds_f64<(Q0 + 1) * CT * 8>(a[d]);
          s1 = __dadd_rn(s1, __dmul_rn(cv.c[d], v));
        }
      }
    }
    sum[Q0] = s0;
    ctr[Q0] = c0;
    if constexpr (NQ == 2) {
      sum[Q0 + 1] = s1;
      ctr[Q0 + 1] = c1;
    }
  }
}

// KIND 0: y = epilogue(A x) (+ dot), 1: CG p/x update + <p, A p>, 2: CG r update + <r, r> + record
// WX (KIND 0): the dot operand w is x itself and the middle diagonal is the main diagonal.
// PART (KIND 1 / 2): row-partitioned, ghost-extended row space (see KbMarchCg); false compiles the
// ownership tests and the peer stores out.
// CT: consumer threads (a multiple of 32; + one producer warp).  A tile is CT * RPT rows wide, so
// the number of column tiles of a plane can be matched to the CTA slots of the machine
// (512 x 512 planes: 256 tiles of 1024 rows, 293 of 896, 342 of 768).
template <int ND, int RPT, int NS, int MINB, int KIND, int DOT, bool WX, bool PART = false,
          int CT = 256>
__global__ void __launch_bounds__(CT + 32, MINB)
kb_stencil_march_kernel(int n_rows, int n_cols, KbMarch g, const uint16_t* __restrict__ masks,
                        KbPattern pat, KbConstVals cv, const double* __restrict__ x,
                        double* __restrict__ y, int mode, const double* __restrict__ z,
                        const double* __restrict__ coef, const double* __restrict__ w,
                        KbMarchCg cg, int l2pol, double* __restrict__ out, KbRed rd) {
  static_assert(RPT >= 2 && RPT <= 4, "rows per thread");
  static_assert(CT % 32 == 0 && CT >= 128 && CT <= 256, "consumer threads");
  static_assert(NS >= 4, "ring: previous, current, next plane + one in flight");
  kb_pdl_prologue();
  if (kb_gated(rd)) return;
  constexpr int TR = CT * RPT;
  constexpr int NR = KIND == 1 ? 2 : 0;  // staging windows of r (KIND 1)
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  double* const s_win = reinterpret_cast<double*>(kb_dyn_smem);
  double* const s_rwin = s_win + (size_t)NS * g.wlen;
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(s_rwin + (size_t)NR * g.wlen);
  uint64_t* const s_empty = s_full + NS;
  uint64_t* const s_rempty = s_empty + NS;  // [2], KIND 1
  __shared__ double red_sm[CT + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      kb_mbar_init(&s_full[s], 1);
      kb_mbar_init(&s_empty[s], CT / 32);
    }
    if (KIND == 1) {
      kb_mbar_init(&s_rempty[0], CT / 32);
      kb_mbar_init(&s_rempty[1], CT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;
  if (warp == CT / 32) {
    // ------------------------------------------------ producer: one window per plane ---
    if ((tid & 31) == 0) {
      const uint64_t pol = l2pol == 2 ? kb_policy_evict_first() : kb_policy_evict_last();
      unsigned cnt = 0;
      for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
        const int c = item % g.ncol;
        const int j0 = (item / g.ncol) * g.ch;
        const int j1 = min(j0 + g.ch, g.nplanes);
        const int nload = j1 - j0 + 2;
        for (int l = 0; l < nload; ++l, ++cnt) {
          const int slot = (int)(cnt % NS);
          const unsigned use = cnt / NS;
          if (use > 0) kb_mbar_wait(&s_empty[slot], (use - 1u) & 1u);
          if (KIND == 1 && cnt >= 2) kb_mbar_wait(&s_rempty[cnt & 1u], ((cnt >> 1) - 1u) & 1u);
          const long long g0 = (long long)(j0 - 1 + l) * g.P + (long long)c * TR - g.LP;
          const long long lo = g0 > 0 ? g0 : 0;
          const long long hi = (g0 + g.wlen < (long long)n_cols) ? g0 + g.wlen : (long long)n_cols;
          if (hi > lo) {
            const uint32_t bytes = (uint32_t)(hi - lo) * 8u;
            double* dst = s_win + (size_t)slot * g.wlen + (lo - g0);
            kb_mbar_expect_tx(&s_full[slot], KIND == 1 ? 2u * bytes : bytes);
            if (l2pol == 1)
              kb_bulk_g2s(dst, x + lo, bytes, &s_full[slot]);
            else
              kb_bulk_g2s_hint(dst, x + lo, bytes, &s_full[slot], pol);
            if (KIND == 1) {
              double* rdst = s_rwin + (size_t)(cnt & 1u) * g.wlen + (lo - g0);
              if (l2pol == 1)
                kb_bulk_g2s(rdst, cg.r_in + lo, bytes, &s_full[slot]);
              else
                kb_bulk_g2s_hint(rdst, cg.r_in + lo, bytes, &s_full[slot], pol);
            }
          } else {
            kb_mbar_arrive(&s_full[slot]);  // window entirely outside the vector
          }
        }
      }
    }
  } else {
    // ------------------------------------------------ consumers: RPT rows per thread ---
    const uint32_t sbase = kb_smem_u32(s_win);
    const uint32_t rbase = kb_smem_u32(s_rwin);
    const uint32_t wbytes = (uint32_t)g.wlen * 8u;
    const uint32_t tb = (uint32_t)(g.LP + tid) * 8u;  // this thread's row q = 0 in a window
    uint32_t od[ND];  // byte offset of diagonal d relative to the window start (inner diagonals)
#pragma unroll
    for (int d = 0; d < ND; ++d) od[d] = tb + (uint32_t)(8 * ((d == 0 || d == ND - 1) ? 0 : pat.off[d]));
    double omega = 0.0, alpha = 0.0, cf = 0.0;
    if (KIND == 0 && mode == 1) cf = coef[0];
    if (KIND == 1) {
      omega = cg.rho_a[0] / kb_nz(cg.rho_b[0]);
      if (cg.xv != nullptr) alpha = cg.alpha_in[0];
    }
    if (KIND == 2) {
      alpha = cg.rho_a[0] / kb_nz(cg.rho_b[0]);
      if (cg.alpha_out != nullptr && blockIdx.x == 0 && tid == 0) cg.alpha_out[0] = alpha;
    }
    // Row bookkeeping is per pass, not per row (ncu: it was most of the instructions): the rows
    // of a thread are pos0 + q * 256 of ONE plane, the first nq of them exist (n_rows is a whole
    // number of planes: kb_march_geom), whether a plane is computed / written / owned is uniform
    // over the CTA, and row indices are 32-bit with q * 256 folded into the address offsets.
    const uint32_t fullb = kb_smem_u32(s_full), emptyb = kb_smem_u32(s_empty);
    const uint32_t rempb = kb_smem_u32(s_rempty);
    const bool have_x = KIND == 1 && cg.xv != nullptr;
    unsigned cnt = 0;
    for (int item = blockIdx.x; item < g.nitems; item += gridDim.x) {
      const int c = item % g.ncol;
      const int j0 = (item / g.ncol) * g.ch;
      const int j1 = min(j0 + g.ch, g.nplanes);
      const int nload = j1 - j0 + 2;
      const int pos0 = c * TR + tid;  // in-plane position of row q = 0
      const int nq = min(RPT, max(0, (g.P - pos0 + CT - 1) / CT));  // rows of this thread in a plane
      // Operands that come from global memory (masks, z / r, w, x of the own rows) are requested
      // one pass ahead: the windows are usually there when a pass starts, so a load issued at
      // the top of the pass that needs it would be waited for in full (ncu: the top stall).
      unsigned mn[RPT];
      double zn[RPT], wn[RPT], xn[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        mn[q] = 0u;
        zn[q] = wn[q] = xn[q] = 0.0;
      }
      int rowc = (j0 - 2) * g.P + pos0;  // row q = 0 of plane jc (meaningful from l = 2 on)
      for (int l = 0; l < nload; ++l, ++cnt, rowc += g.P) {
        const int slot = (int)(cnt % NS);
        const int jc = j0 + l - 2;  // plane whose rows are computed in this pass (l >= 2)
        // uniform flags of this pass; PART: only planes [own_pl0, own_pl1) are this rank's
        bool comp = l >= 2;                    // rows of plane jc are computed now
        const bool pre = l >= 1 && l + 1 < nload;  // rows of plane jc + 1 are computed next
        bool pre_own = pre;
        const bool arr = KIND == 1 && l >= 1 && l <= nload - 2;  // own rows of the arriving plane
        bool arr_x = arr && have_x;                              // jc + 1 are written now
        bool pre_x = have_x && l + 1 <= nload - 2;               // x of plane jc + 2, for the next pass
        if (PART) {
          comp = comp && jc >= cg.own_pl0 && jc < cg.own_pl1;
          pre_own = pre_own && jc + 1 >= cg.own_pl0 && jc + 1 < cg.own_pl1;
          arr_x = arr_x && jc + 1 >= cg.own_pl0 && jc + 1 < cg.own_pl1;
          pre_x = pre_x && jc + 2 >= cg.own_pl0 && jc + 2 < cg.own_pl1;
        }
        unsigned m[RPT];
        double zv[RPT], wv[RPT], xo[RPT];
        const int rown = rowc + g.P;  // plane jc + 1: computed next (KIND 0 / 2), arriving now (KIND 1)
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          m[q] = mn[q];
          zv[q] = zn[q];
          wv[q] = wn[q];
          xo[q] = xn[q];
          mn[q] = 0u;
          zn[q] = wn[q] = 0.0;
          if (pre && q < nq) {
            mn[q] = (unsigned)masks[rown + q * CT];
            if (KIND == 0 && mode != 0) zn[q] = z[rown + q * CT];
            if (KIND == 0 && DOT == 1 && !WX) wn[q] = w[rown + q * CT];
            if (KIND == 2 && pre_own) zn[q] = cg.r[rown + q * CT];
          }
          if (KIND == 1) xn[q] = (pre_x && q < nq) ? cg.xv[rown + g.P + q * CT] : 0.0;
        }
        kb_mbar_wait_u32(fullb + (uint32_t)slot * 8u, (cnt / NS) & 1u);
        if (KIND == 1) {
          // p <- r + omega p on the whole arriving window (halo included), in place; the own
          // rows also go to global memory together with x += alpha p_old  (cg.py:178,196)
          const uint32_t pw = sbase + (uint32_t)slot * wbytes + (uint32_t)tid * 8u;
          const uint32_t rw = rbase + (cnt & 1u) * wbytes + (uint32_t)tid * 8u;
          const uint32_t own = (uint32_t)g.LP * 8u;  // entries LP/256 ... + RPT are this thread's rows
          {
            double po[RPT], ro[RPT];
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
              po[q] = kb_ld_f64_shared(pw + own + (uint32_t)(q * CT * 8));
              ro[q] = kb_ld_f64_shared(rw + own + (uint32_t)(q * CT * 8));
            }
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
              const double pn = kb_mul_add(omega, po[q], ro[q]);
              kb_st_f64_shared(pw + own + (uint32_t)(q * CT * 8), pn);
              if (arr && q < nq) {
                if (arr_x) __stcs(&cg.xv[rown + q * CT], kb_mul_add(alpha, po[q], xo[q]));
                __stcs(&cg.p_out[rown + q * CT], pn);
              }
            }
          }
          // halo entries on both sides: LP / 256 per side and thread, two at a time
          const int nh = g.LP / CT;  // LP is a multiple of CT (kb_march_geom)
          for (int side = 0; side < 2; ++side) {
            const uint32_t hb = side == 0 ? 0u : own + (uint32_t)(RPT * CT * 8);
            int u = 0;
            for (; u + 1 < nh; u += 2) {
              const uint32_t o0 = hb + (uint32_t)u * (CT * 8u), o1 = o0 + CT * 8u;
              const double p0 = kb_ld_f64_shared(pw + o0), p1 = kb_ld_f64_shared(pw + o1);
              const double r0 = kb_ld_f64_shared(rw + o0), r1 = kb_ld_f64_shared(rw + o1);
              kb_st_f64_shared(pw + o0, kb_mul_add(omega, p0, r0));
              kb_st_f64_shared(pw + o1, kb_mul_add(omega, p1, r1));
            }
            if (u < nh) {
              const uint32_t o0 = hb + (uint32_t)u * (CT * 8u);
              const double p0 = kb_ld_f64_shared(pw + o0);
              const double r0 = kb_ld_f64_shared(rw + o0);
              kb_st_f64_shared(pw + o0, kb_mul_add(omega, p0, r0));
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) kb_mbar_arrive_u32(rempb + (cnt & 1u) * 8u);
          asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory");  // the window is complete for everyone
        }
        if (l >= 2) {
          if (comp) {
            const uint32_t wprev = sbase + (uint32_t)((cnt - 2u) % NS) * wbytes;
            const uint32_t wcur = sbase + (uint32_t)((cnt - 1u) % NS) * wbytes;
            const uint32_t wnext = sbase + (uint32_t)slot * wbytes;
            uint32_t a[ND];
            a[0] = wprev + od[0];
#pragma unroll
            for (int d = 1; d < ND - 1; ++d) a[d] = wcur + od[d];
            a[ND - 1] = wnext + od[ND - 1];
            double sum[RPT], ctr[RPT];
            kb_march_rows<ND, 0, 2, CT>(a, m, cv, sum, ctr);
            if constexpr (RPT == 4) kb_march_rows<ND, 2, 2, CT>(a, m, cv, sum, ctr);
            if constexpr (RPT == 3) kb_march_rows<ND, 2, 1, CT>(a, m, cv, sum, ctr);
            // PART, KIND 2: the first / last own plane also goes to the neighbours' ghost planes
            double* plo = nullptr;
            double* phi = nullptr;
            if (PART && KIND == 2) {
              if (cg.push_lo != nullptr && jc == cg.own_pl0) plo = cg.push_lo + pos0;
              if (cg.push_hi != nullptr && jc == cg.own_pl1 - 1) phi = cg.push_hi + pos0;
            }
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
              if (q < nq) {
                if (KIND == 0) {
                  double yv = sum[q];
                  if (mode == 1) yv = kb_mul_sub(cf, zv[q], sum[q]);
                  if (mode == 2) yv = __dsub_rn(zv[q], sum[q]);
                  __stcs(&y[rowc + q * CT], yv);
                  if (DOT == 1) acc = fma(WX ? ctr[q] : wv[q], yv, acc);
                  if (DOT == 2) acc = fma(yv, yv, acc);
                } else if (KIND == 1) {
                  acc = fma(ctr[q], sum[q], acc);  // <p, A p>
                } else {
                  const double rn = kb_mul_sub(alpha, sum[q], zv[q]);  // r - alpha (A p)
                  cg.r[rowc + q * CT] = rn;
                  if (PART) {
                    if (plo != nullptr) plo[q * CT] = rn;
                    if (phi != nullptr) phi[q * CT] = rn;
                  }
                  acc = fma(rn, rn, acc);
                }
              }
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) kb_mbar_arrive_u32(emptyb + ((cnt - 2u) % NS) * 8u);
        }
      }
      // the last two windows of the item are not needed by a later pass
      __syncwarp();
      if ((tid & 31) == 0) {
        kb_mbar_arrive_u32(emptyb + ((cnt - 2u) % NS) * 8u);
        kb_mbar_arrive_u32(emptyb + ((cnt - 1u) % NS) * 8u);
      }
    }
  }
  if (KIND == 0 && DOT == 0) return;
  // peer stores of r must be visible system-wide before this block's arrival is counted: the
  // finishing block's all-reduce flag is the neighbours' "ghost planes landed" signal
  if (PART && KIND == 2 && (cg.push_lo != nullptr || cg.push_hi != nullptr)) __threadfence_system();
  const bool last = kb_grid_colsum(acc, 1, rd, out, red_sm);
  if (KIND == 2 && last && cg.rec.step >= 0) {  // cg.py:156,214-217 in the finishing block
    __syncthreads();
    if (threadIdx.x == 0) {
      const double rn = out[0];
      if (cg.rec.rho_keep != nullptr) cg.rec.rho_keep[0] = rn;
      const double nrm = sqrt(rn);
      cg.rec.hist[(size_t)cg.rec.step] = nrm;
      if (nrm <= cg.rec.crit[0]) *cg.rec.stop_at = cg.rec.step;
    }
  }
}
