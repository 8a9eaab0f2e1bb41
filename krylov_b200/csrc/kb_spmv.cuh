// CSR sparse products fused with the operation that follows them.
//
//   t = A x ;  y = t | t - coef*z | z - t ;  out = <w,y> | <y,y> | -
//
// Two schedules:
//  * kb_spmv_rowwise_kernel  -- any k, any matrix: one thread per (row, column),
//    k lanes share a row (its vals/cols loads are broadcast, the x row is one
//    contiguous k*8-byte read).
//  * kb_spmv_stream_kernel   -- k == 1, short rows (stencils): persistent CTAs;
//    a producer lane streams each 256-row tile's colidx/vals into a shared-memory
//    ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 8
//    consumer warps take one row per thread out of shared memory, gather x
//    through L1/L2 and write y coalesced.  Matrix bytes are read from HBM exactly
//    once, fully coalesced, independent of row length.
//
// Row sums are accumulated left to right with a rounded product and a rounded
// add (no FMA): the same order and roundings as SciPy's csr_matvec(s)
// (the routine behind `A @ x` at _helpers.py:47), so y is bit-identical to the
// reference's product.
#pragma once
#include "kb_common.cuh"
#include "kb_ptx.cuh"
#include "kb_handles.cuh"

// ------------------------------------------------------------ row-wise k>=1 --
template <int DOT>
__global__ void __launch_bounds__(KB_BLOCK)
kb_spmv_rowwise_kernel(int64_t n_rows, int k, const int32_t* __restrict__ rowptr,
                       const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                       const double* __restrict__ x, double* __restrict__ y, int mode,
                       const double* __restrict__ z, const double* __restrict__ coef,
                       const double* __restrict__ w, double* __restrict__ out, int contiguous,
                       KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  const int c = threadIdx.x % k;
  const int rows_per_block = blockDim.x / k;
  const int rsub = threadIdx.x / k;
  const double cf = (mode == 1) ? coef[c] : 0.0;
  double acc = 0.0;
  // contiguous: block b walks its own contiguous slice of rows, so the x rows that
  // neighbouring matrix rows share (stencils: +-1, +-nx) are still in this SM's L1
  // when they are needed again; otherwise rows are dealt round-robin to the blocks.
  int64_t row0 = (int64_t)blockIdx.x * rows_per_block, row_end = n_rows;
  int64_t step = (int64_t)gridDim.x * rows_per_block;
  if (contiguous) {
    const int64_t per = (n_rows + gridDim.x - 1) / gridDim.x;
    row0 = (int64_t)blockIdx.x * per;
    row_end = row0 + per < n_rows ? row0 + per : n_rows;
    step = rows_per_block;
  }
  for (int64_t row = row0 + rsub; row < row_end; row += step) {
    const int lo = rowptr[row], hi = rowptr[row + 1];
    double sum = 0.0;
    int j = lo;
    // four independent index/value/gather chains in flight, then the ordered sum
    for (; j + 3 < hi; j += 4) {
      const int c0 = colidx[j], c1 = colidx[j + 1], c2 = colidx[j + 2], c3 = colidx[j + 3];
      const double v0 = vals[j], v1 = vals[j + 1], v2 = vals[j + 2], v3 = vals[j + 3];
      const double x0 = x[(size_t)c0 * k + c], x1 = x[(size_t)c1 * k + c];
      const double x2 = x[(size_t)c2 * k + c], x3 = x[(size_t)c3 * k + c];
      sum = __dadd_rn(sum, __dmul_rn(v0, x0));
      sum = __dadd_rn(sum, __dmul_rn(v1, x1));
      sum = __dadd_rn(sum, __dmul_rn(v2, x2));
      sum = __dadd_rn(sum, __dmul_rn(v3, x3));
    }
    for (; j < hi; ++j)
      sum = __dadd_rn(sum, __dmul_rn(vals[j], x[(size_t)colidx[j] * k + c]));
    const size_t idx = (size_t)row * k + c;
    const double yv = kb_spmv_epilogue(sum, mode, z, cf, idx);
    y[idx] = yv;
    if (DOT == 1) acc = fma(w[idx], yv, acc);
    if (DOT == 2) acc = fma(yv, yv, acc);
  }
  if (DOT != 0) kb_grid_colsum(acc, k, rd, out, sm);
}

// ---------------------------------------------------- TMA-staged stream k==1 --
// rows per tile == consumer threads (template parameter ROWS); + one producer warp
#define KB_ST_MAX_THREADS (512 + 32)

template <int STAGES, int CAP>
struct KbStreamSmem {
  double vals[STAGES][CAP];
  int32_t cols[STAGES][CAP];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
};

// chunk c of a tile whose nnz occupy [s, e): 4-aligned window of <= CAP entries
__device__ __forceinline__ void kb_chunk_range(int s, int e, int cidx, int cap, int& a, int& b) {
  const int a0 = s & ~3;
  const int a1 = (e + 3) & ~3;
  a = a0 + cidx * cap;
  b = min(a + cap, a1);
}

template <int ROWS, int STAGES, int CAP, int DOT>
__global__ void __launch_bounds__(ROWS + 32)
kb_spmv_stream_kernel(int n_rows, int n_tiles, const int32_t* __restrict__ rowptr,
                      const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                      const double* __restrict__ x, double* __restrict__ y, int mode,
                      const double* __restrict__ z, const double* __restrict__ coef,
                      const double* __restrict__ w, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  KbStreamSmem<STAGES, CAP>& S = *reinterpret_cast<KbStreamSmem<STAGES, CAP>*>(kb_dyn_smem);
  __shared__ double red_sm[ROWS + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int nconsumer_warps = ROWS / 32;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      kb_mbar_init(&S.full[s], 1);                  // producer's expect_tx arrive
      kb_mbar_init(&S.empty[s], nconsumer_warps);   // one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;

  if (warp == nconsumer_warps) {
    // ===================== producer warp ====================================
    // All 32 lanes fetch the nnz bounds of this CTA's next 32 tiles at once (one
    // round of latency per 32 tiles); lane 0 then issues the bulk copies.
    const int lane = tid & 31;
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t base = blockIdx.x; base < n_tiles; base += 32ll * gridDim.x) {
      const int64_t my_tile = base + (int64_t)lane * gridDim.x;
      int my_s = 0, my_e = 0;
      if (my_tile < n_tiles) {
        const int r0 = (int)my_tile * ROWS;
        const int r1 = min(r0 + ROWS, n_rows);
        my_s = rowptr[r0];
        my_e = rowptr[r1];
      }
      for (int q = 0; q < 32; ++q) {
        if (base + (int64_t)q * gridDim.x >= n_tiles) break;  // warp-uniform
        const int s = __shfl_sync(0xffffffffu, my_s, q);
        const int e = __shfl_sync(0xffffffffu, my_e, q);
        if (lane == 0 && e > s) {
          const int nch = (((e + 3) & ~3) - (s & ~3) + CAP - 1) / CAP;
          for (int ci = 0; ci < nch; ++ci) {
            int a, b;
            kb_chunk_range(s, e, ci, CAP, a, b);
            kb_mbar_wait(&S.empty[stage], phase ^ 1u);  // slot free (passes at once in round 0)
            const uint32_t cnt = (uint32_t)(b - a);
            kb_mbar_expect_tx(&S.full[stage], cnt * 12u);
            kb_bulk_g2s(&S.vals[stage][0], vals + a, cnt * 8u, &S.full[stage]);
            kb_bulk_g2s(&S.cols[stage][0], colidx + a, cnt * 4u, &S.full[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ consumers: one row per thread ==============
    int stage = 0;
    uint32_t phase = 0;
    const double cf = (mode == 1) ? coef[0] : 0.0;
    int tile = blockIdx.x;
    // software-pipelined row pointers of the next tile
    int lo_n = 0, hi_n = 0, s_n = 0, e_n = 0;
    if (tile < n_tiles) {
      const int r0 = tile * ROWS;
      const int r1 = min(r0 + ROWS, n_rows);
      const int row = r0 + tid;
      s_n = rowptr[r0];
      e_n = rowptr[r1];
      if (row < n_rows) {
        lo_n = rowptr[row];
        hi_n = rowptr[row + 1];
      }
    }
    for (; tile < n_tiles; tile += gridDim.x) {
      const int r0 = tile * ROWS;
      const int row = r0 + tid;
      const int lo = lo_n, hi = hi_n, s = s_n, e = e_n;
      {
        const int nt = tile + gridDim.x;
        if (nt < n_tiles) {
          const int q0 = nt * ROWS;
          const int q1 = min(q0 + ROWS, n_rows);
          const int qrow = q0 + tid;
          s_n = rowptr[q0];
          e_n = rowptr[q1];
          lo_n = hi_n = 0;
          if (qrow < n_rows) {
            lo_n = rowptr[qrow];
            hi_n = rowptr[qrow + 1];
          }
        }
      }
      double sum = 0.0;
      if (e > s) {
        const int nch = (((e + 3) & ~3) - (s & ~3) + CAP - 1) / CAP;
        for (int ci = 0; ci < nch; ++ci) {
          int a, b;
          kb_chunk_range(s, e, ci, CAP, a, b);
          kb_mbar_wait(&S.full[stage], phase);
          const int jb = max(lo, a), je = min(hi, b);
          const double* sv = &S.vals[stage][0];
          const int32_t* sc = &S.cols[stage][0];
          int j = jb - a;
          const int jend = je - a;
          // gathers first (independent), then the ordered sum
          for (; j + 3 < jend; j += 4) {
            const double x0 = __ldg(x + sc[j]);
            const double x1 = __ldg(x + sc[j + 1]);
            const double x2 = __ldg(x + sc[j + 2]);
            const double x3 = __ldg(x + sc[j + 3]);
            sum = __dadd_rn(sum, __dmul_rn(sv[j], x0));
            sum = __dadd_rn(sum, __dmul_rn(sv[j + 1], x1));
            sum = __dadd_rn(sum, __dmul_rn(sv[j + 2], x2));
            sum = __dadd_rn(sum, __dmul_rn(sv[j + 3], x3));
          }
          for (; j < jend; ++j) sum = __dadd_rn(sum, __dmul_rn(sv[j], __ldg(x + sc[j])));
          __syncwarp();
          if ((tid & 31) == 0) kb_mbar_arrive(&S.empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (row < n_rows) {
        const double yv = kb_spmv_epilogue(sum, mode, z, cf, (size_t)row);
        y[row] = yv;
        if (DOT == 1) acc = fma(w[row], yv, acc);
        if (DOT == 2) acc = fma(yv, yv, acc);
      }
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm);
}

// ------------------------------------------- offset-pattern compressed stream --
// Stencil-like matrices have only a handful of distinct diagonals: every column
// index is row + D[d] for d in a set of <= 16 offsets.  kb_csr_create detects
// this and stores ONE 16-bit mask per row (which diagonals are present) instead
// of one 32-bit index per nonzero; the values stay in CSR order, so the row sum
// visits the same products in the same order (bit-identical result) while the
// matrix stream shrinks from 12 to 8 bytes per nonzero (+2 per row).
// Per launch: 8 nnz + 2 n + 4 (n+1) + 16 n bytes  (7-point: 78 instead of 104 B/row).
#define KB_PAT_EMPTY (-2147483647 - 1)

__global__ void __launch_bounds__(256)
kb_pattern_collect_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr,
                          const int32_t* __restrict__ colidx, int* table, int* overflow) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    const int lo = rowptr[row], hi = rowptr[row + 1];
    for (int j = lo; j < hi; ++j) {
      const int off = colidx[j] - (int)row;
      const unsigned h = ((unsigned)off * 2654435761u) >> 26;  // 64 slots
      bool done = false;
      for (int s = 0; s < 64 && !done; ++s) {
        const int idx = (h + s) & 63;
        int v = *(volatile int*)&table[idx];
        if (v == off) {
          done = true;
        } else if (v == KB_PAT_EMPTY) {
          v = atomicCAS(&table[idx], KB_PAT_EMPTY, off);
          if (v == KB_PAT_EMPTY || v == off) done = true;
        }
      }
      if (!done) *overflow = 1;
    }
  }
}

__global__ void __launch_bounds__(256)
kb_pattern_build_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr,
                        const int32_t* __restrict__ colidx, KbPattern pat,
                        uint16_t* __restrict__ masks, int* fail) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    const int lo = rowptr[row], hi = rowptr[row + 1];
    unsigned m = 0;
    int prev = -1;
    for (int j = lo; j < hi; ++j) {
      const int off = colidx[j] - (int)row;
      int d = -1;
      for (int q = 0; q < pat.nd; ++q)
        if (pat.off[q] == off) d = q;
      if (d <= prev) *fail = 1;  // unknown offset, duplicate, or not in ascending column order
      prev = d;
      if (d >= 0) m |= 1u << d;
    }
    masks[row] = (uint16_t)m;
  }
}

// Constant diagonals ("stencil" schedule).  On constant-coefficient stencils every entry of a
// diagonal carries the same value (boundary rows only lack entries, which the mask records), so
// the 8 B/nonzero value stream can go too: <= 8 doubles travel as kernel parameters and the
// product streams 2 (mask) + 8 (x) + 8 (y) bytes per row.  Same products in the same order with
// bitwise the same values -> bit-identical result.  Detection is exact (64-bit compare of every
// stored value against its diagonal's representative) and runs once in kb_csr_create.

__global__ void __launch_bounds__(256)
kb_constdiag_fill_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr,
                         const double* __restrict__ vals, const uint16_t* __restrict__ masks,
                         unsigned long long* cbits) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    int j = rowptr[row];
    const unsigned m = masks[row];
    for (int d = 0; d < 16; ++d)
      if ((m >> d) & 1u) cbits[d] = (unsigned long long)__double_as_longlong(vals[j++]);
  }
}

__global__ void __launch_bounds__(256)
kb_constdiag_check_kernel(int64_t n_rows, const int32_t* __restrict__ rowptr,
                          const double* __restrict__ vals, const uint16_t* __restrict__ masks,
                          const unsigned long long* __restrict__ cbits, int* fail) {
  unsigned long long c[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) c[d] = cbits[d];
  bool bad = false;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows;
       row += (int64_t)gridDim.x * blockDim.x) {
    int j = rowptr[row];
    const unsigned m = masks[row];
#pragma unroll
    for (int d = 0; d < 16; ++d)
      if ((m >> d) & 1u)
        bad |= (unsigned long long)__double_as_longlong(vals[j++]) != c[d];
  }
  if (bad) *fail = 1;
}

template <int STAGES, int CAP>
struct KbPatternSmem {
  double vals[STAGES][CAP];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
};

template <int ROWS, int STAGES, int CAP, int MAXD, int MINCTAS, int DOT>
__global__ void __launch_bounds__(ROWS + 32, MINCTAS)
kb_spmv_pattern_kernel(int n_rows, int n_tiles, const int32_t* __restrict__ rowptr,
                       const uint16_t* __restrict__ masks, const double* __restrict__ vals,
                       KbPattern pat, const double* __restrict__ x, double* __restrict__ y,
                       int mode, const double* __restrict__ z, const double* __restrict__ coef,
                       const double* __restrict__ w, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  KbPatternSmem<STAGES, CAP>& S = *reinterpret_cast<KbPatternSmem<STAGES, CAP>*>(kb_dyn_smem);
  __shared__ double red_sm[ROWS + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int nconsumer_warps = ROWS / 32;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      kb_mbar_init(&S.full[s], 1);
      kb_mbar_init(&S.empty[s], nconsumer_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;

  if (warp == nconsumer_warps) {
    // producer warp: bounds of 32 tiles per round, lane 0 issues the TMA bulk copies
    const int lane = tid & 31;
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t base = blockIdx.x; base < n_tiles; base += 32ll * gridDim.x) {
      const int64_t my_tile = base + (int64_t)lane * gridDim.x;
      int my_s = 0, my_e = 0;
      if (my_tile < n_tiles) {
        const int r0 = (int)my_tile * ROWS;
        const int r1 = min(r0 + ROWS, n_rows);
        my_s = rowptr[r0];
        my_e = rowptr[r1];
      }
      for (int q = 0; q < 32; ++q) {
        if (base + (int64_t)q * gridDim.x >= n_tiles) break;
        const int s = __shfl_sync(0xffffffffu, my_s, q);
        const int e = __shfl_sync(0xffffffffu, my_e, q);
        if (lane == 0 && e > s) {
          const int nch = (((e + 3) & ~3) - (s & ~3) + CAP - 1) / CAP;
          for (int ci = 0; ci < nch; ++ci) {
            int a, b;
            kb_chunk_range(s, e, ci, CAP, a, b);
            kb_mbar_wait(&S.empty[stage], phase ^ 1u);
            const uint32_t cnt = (uint32_t)(b - a);
            kb_mbar_expect_tx(&S.full[stage], cnt * 8u);
            kb_bulk_g2s(&S.vals[stage][0], vals + a, cnt * 8u, &S.full[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        __syncwarp();
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    const double cf = (mode == 1) ? coef[0] : 0.0;
    int tile = blockIdx.x;
    int lo_n = 0, s_n = 0, e_n = 0;
    unsigned m_n = 0;
    if (tile < n_tiles) {
      const int r0 = tile * ROWS;
      const int r1 = min(r0 + ROWS, n_rows);
      const int row = r0 + tid;
      s_n = rowptr[r0];
      e_n = rowptr[r1];
      if (row < n_rows) {
        lo_n = rowptr[row];
        m_n = masks[row];
      }
    }
    for (; tile < n_tiles; tile += gridDim.x) {
      const int r0 = tile * ROWS;
      const int row = r0 + tid;
      const int lo = lo_n, s = s_n, e = e_n;
      const unsigned mask = m_n;
      {
        const int nt = tile + gridDim.x;
        if (nt < n_tiles) {
          const int q0 = nt * ROWS;
          const int q1 = min(q0 + ROWS, n_rows);
          const int qrow = q0 + tid;
          s_n = rowptr[q0];
          e_n = rowptr[q1];
          lo_n = 0;
          m_n = 0;
          if (qrow < n_rows) {
            lo_n = rowptr[qrow];
            m_n = masks[qrow];
          }
        }
      }
      double sum = 0.0;
      if (e > s) {
        const int nch = (((e + 3) & ~3) - (s & ~3) + CAP - 1) / CAP;
        for (int ci = 0; ci < nch; ++ci) {
          int a, b;
          kb_chunk_range(s, e, ci, CAP, a, b);
          kb_mbar_wait(&S.full[stage], phase);
          const double* sv = &S.vals[stage][0];
          // all gathers are independent of each other: issue them, then the ordered sum
          double xv[MAXD];
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            xv[d] = 0.0;
            if (d < pat.nd && (mask >> d) & 1u) xv[d] = __ldg(x + (row + pat.off[d]));
          }
#pragma unroll
          for (int d = 0; d < MAXD; ++d) {
            if (d < pat.nd && (mask >> d) & 1u) {
              const int j = lo + __popc(mask & ((1u << d) - 1u));  // CSR position of diagonal d
              if (j >= a && j < b) sum = __dadd_rn(sum, __dmul_rn(sv[j - a], xv[d]));
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) kb_mbar_arrive(&S.empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (row < n_rows) {
        const double yv = kb_spmv_epilogue(sum, mode, z, cf, (size_t)row);
        y[row] = yv;
        if (DOT == 1) acc = fma(w[row], yv, acc);
        if (DOT == 2) acc = fma(yv, yv, acc);
      }
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm);
}

// ------------------------------------------------ windowed pattern kernel -----
// Same compressed matrix as above, but no per-thread gathers at all: for a tile
// of ROWS consecutive rows the x entries on diagonal d are the CONTIGUOUS range
// x[r0 + off[d] .. r0 + off[d] + ROWS), so the producer lane TMA-copies a handful
// of x windows (nearby diagonals share one) into shared memory together with the
// tile's values.  Consumers touch global memory only for rowptr/mask (prefetched
// one tile ahead), the epilogue operands and the y store; everything else is
// asynchronous bulk traffic, so the kernel is bound by HBM again
// (8 nnz + 2 n + 4 (n+1) + 16 n bytes per launch).
// Requires: nd <= 8, window spans <= KB_WIN_SLACK-2, even n_cols, 16-byte aligned x.
#define KB_WIN_SLACK 40

// first (even) global index held by window `wlo` of the tile starting at r0
__device__ __forceinline__ int kb_win_start(int r0, int wlo) { return max(r0 + wlo, 0) & ~1; }

// Cache-blocked tile order.  With a large "plane" offset P (3-D stencils) an x entry is
// used when the sweep passes rows j-P, j and j+P; at 512^2-row planes two planes of matrix
// stream (40 MB) separate those uses and the x windows fall out of L2 (measured: DRAM traffic
// 1.13x at 512^3 vs 1.00x at 384^3).  So tiles are visited block-of-lines by block-of-lines:
// the i-th tile processed is  plane z, tile (yb*tpb + tt) of that plane  with
//   yb = i / (nplanes*tpb),  z = (i % (nplanes*tpb)) / tpb,  tt = i % tpb,
// which brings the +-P reuse distance down to tpb tiles (~2.5 MB).  tpb == 0: natural order.
struct KbTileOrder {
  int tpp;      // tiles per plane
  int tpb;      // tiles per block of lines (divides tpp); 0 = natural order
  int nplanes;  // n_tiles / tpp
};
__device__ __forceinline__ int kb_tile_of(int i, const KbTileOrder& o) {
  if (o.tpb <= 0) return i;
  const int per = o.nplanes * o.tpb;
  const int yb = i / per;
  const int rem = i - yb * per;
  const int z = rem / o.tpb;
  return z * o.tpp + yb * o.tpb + (rem - z * o.tpb);
}

// Shared memory (dynamic, sized by the host from the actual pattern so that as many
// CTAs as possible are resident):  vals[STAGES][cap] | win[STAGES][nw][wlen] | barriers
// cap = 4-aligned (ROWS * nd + 8), wlen = even (ROWS + max span + 6).
// CONSTV (constant diagonals): no value stream and no row pointers at all -- cap == 0, the stage
// buffers hold x windows only, the coefficients come from `cv`.
template <int ROWS, int STAGES, int MINB, int DOT, bool CONSTV>
__global__ void __launch_bounds__(ROWS + 32, MINB)
kb_spmv_window_kernel(int n_rows, int n_cols, int n_tiles, int cap, int wlen, KbTileOrder ord,
                      const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ masks,
                      const double* __restrict__ vals, KbPattern pat, KbConstVals cv,
                      const double* __restrict__ x, double* __restrict__ y, int mode,
                      const double* __restrict__ z, const double* __restrict__ coef,
                      const double* __restrict__ w, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  double* const s_vals = reinterpret_cast<double*>(kb_dyn_smem);
  double* const s_win = s_vals + (size_t)STAGES * cap;
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(s_win + (size_t)STAGES * pat.nw * wlen);
  uint64_t* const s_empty = s_full + STAGES;
  __shared__ double red_sm[ROWS + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int nconsumer_warps = ROWS / 32;
  const int nw = pat.nw;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      kb_mbar_init(&s_full[s], 1);
      kb_mbar_init(&s_empty[s], nconsumer_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;

  if (warp == nconsumer_warps) {
    const int lane = tid & 31;
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_stream = kb_policy_evict_first();  // matrix values: read once
    const uint64_t pol_keep = kb_policy_evict_last();     // x windows: re-read by later tiles
    for (int64_t base = blockIdx.x; base < n_tiles; base += 32ll * gridDim.x) {
      const int64_t my_tile = base + (int64_t)lane * gridDim.x;  // position in the visiting order
      int my_s = 0, my_e = CONSTV ? 1 : 0;
      if (!CONSTV && my_tile < n_tiles) {
        const int r0 = kb_tile_of((int)my_tile, ord) * ROWS;
        const int r1 = min(r0 + ROWS, n_rows);
        my_s = rowptr[r0];
        my_e = rowptr[r1];
      }
      for (int q = 0; q < 32; ++q) {
        const int64_t tile = base + (int64_t)q * gridDim.x;
        if (tile >= n_tiles) break;
        const int s = __shfl_sync(0xffffffffu, my_s, q);
        const int e = __shfl_sync(0xffffffffu, my_e, q);
        if (lane == 0 && e > s) {
          const int r0 = kb_tile_of((int)tile, ord) * ROWS;
          const int a0 = s & ~3, a1 = CONSTV ? a0 : ((e + 3) & ~3);  // <= ROWS*nd + 6 <= cap entries
          kb_mbar_wait(&s_empty[stage], phase ^ 1u);
          uint32_t bytes = (uint32_t)(a1 - a0) * 8u;
          for (int g = 0; g < nw; ++g) {  // pass 1: transaction size
            const int ge = min(r0 + pat.wlo[g] + ROWS + pat.wspan[g], n_cols);
            const int gn = max(((ge + 1) & ~1) - kb_win_start(r0, pat.wlo[g]), 0);
            bytes += (uint32_t)gn * 8u;  // n_cols is even: the rounded end stays in bounds
          }
          kb_mbar_expect_tx(&s_full[stage], bytes);
          if (!CONSTV)
            kb_bulk_g2s_hint(s_vals + (size_t)stage * cap, vals + a0, (uint32_t)(a1 - a0) * 8u,
                             &s_full[stage], pol_stream);
          for (int g = 0; g < nw; ++g) {  // pass 2: issue
            const int ge = min(r0 + pat.wlo[g] + ROWS + pat.wspan[g], n_cols);
            const int ga = kb_win_start(r0, pat.wlo[g]);
            const int gn = max(((ge + 1) & ~1) - ga, 0);
            if (gn > 0)
              kb_bulk_g2s_hint(s_win + ((size_t)stage * nw + g) * wlen, x + ga,
                               (uint32_t)gn * 8u, &s_full[stage], pol_keep);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        __syncwarp();
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    const double cf = (mode == 1) ? coef[0] : 0.0;
    int tile = blockIdx.x;
    int lo_n = 0, s_n = 0, e_n = CONSTV ? 1 : 0;
    unsigned m_n = 0;
    if (tile < n_tiles) {
      const int r0 = kb_tile_of(tile, ord) * ROWS;
      const int r1 = min(r0 + ROWS, n_rows);
      const int row = r0 + tid;
      if (!CONSTV) {
        s_n = rowptr[r0];
        e_n = rowptr[r1];
      }
      if (row < n_rows) {
        if (!CONSTV) lo_n = rowptr[row];
        m_n = masks[row];
      }
    }
    for (; tile < n_tiles; tile += gridDim.x) {  // `tile` = position in the visiting order
      const int r0 = kb_tile_of(tile, ord) * ROWS;
      const int row = r0 + tid;
      const int lo = lo_n, s = s_n, e = e_n;
      const unsigned mask = m_n;
      {
        const int nt = tile + gridDim.x;
        if (nt < n_tiles) {
          const int q0 = kb_tile_of(nt, ord) * ROWS;
          const int q1 = min(q0 + ROWS, n_rows);
          const int qrow = q0 + tid;
          if (!CONSTV) {
            s_n = rowptr[q0];
            e_n = rowptr[q1];
          }
          lo_n = 0;
          m_n = 0;
          if (qrow < n_rows) {
            if (!CONSTV) lo_n = rowptr[qrow];
            m_n = masks[qrow];
          }
          (void)q1;
        }
      }
      // epilogue operands do not depend on the tile data: fetch them early
      double zv = 0.0, wv = 0.0;
      if (row < n_rows) {
        if (mode != 0) zv = z[row];
        if (DOT == 1) wv = w[row];
      }
      double sum = 0.0;
      if (e > s) {
        kb_mbar_wait(&s_full[stage], phase);
        const double* sv = s_vals + (size_t)stage * cap;
        const double* sw = s_win + (size_t)stage * nw * wlen;
        int j = lo - (s & ~3);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          if (d < pat.nd && (mask >> d) & 1u) {
            const int idx = row + pat.off[d] - kb_win_start(r0, pat.dwlo[d]);
            const double xv = sw[pat.grp[d] * wlen + idx];
            sum = __dadd_rn(sum, __dmul_rn(CONSTV ? cv.c[d] : sv[j], xv));
            ++j;
          }
        }
        __syncwarp();
        if ((tid & 31) == 0) kb_mbar_arrive(&s_empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (row < n_rows) {
        double yv = sum;
        if (mode == 1) yv = kb_mul_sub(cf, zv, sum);
        if (mode == 2) yv = __dsub_rn(zv, sum);
        __stcs(&y[row], yv);  // streaming store: y is not re-read by this kernel
        if (DOT == 1) acc = fma(wv, yv, acc);
        if (DOT == 2) acc = fma(yv, yv, acc);
      }
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm);
}

// ------------------------------------------ constant-diagonal stencil kernel ---
// The windowed kernel without any matrix stream (see KbConstVals): 256 consumer threads own
// RPT rows each of a (256 * RPT)-row tile, one producer lane TMA-copies the x windows.
// Measured on the windowed kernel run with constant values (profiles/prof_stencil512_r1*):
// ~200 instructions per row-warp at 68 % issue utilisation -- instruction-bound, not
// memory-bound.  So everything that does not depend on the tile is hoisted (one shared-memory
// offset per diagonal and thread) and interior warps (all masks full: 98.8 % of the rows at
// 512^3) run a branch-free 7 x (LDS, DMUL, DADD) sequence; boundary warps keep the masked
// generic path.  Products and their order are unchanged -> bit-identical.
// With WX (w aliases x, the CG case <p, A p>) the dot operand comes from the centre window.
template <int RPT, int STAGES, int MINB, int DOT>
__global__ void __launch_bounds__(256 + 32, MINB)
kb_spmv_stencil_kernel(int n_rows, int n_cols, int n_tiles, int wlen,
                       const uint16_t* __restrict__ masks, KbPattern pat, KbConstVals cv,
                       const double* __restrict__ x, double* __restrict__ y, int mode,
                       const double* __restrict__ z, const double* __restrict__ coef,
                       const double* __restrict__ w, int d_center, int center_base, int l2pol,
                       double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  constexpr int TR = 256 * RPT;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  double* const s_win = reinterpret_cast<double*>(kb_dyn_smem);
  const int nw = pat.nw;
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(s_win + (size_t)STAGES * nw * wlen);
  uint64_t* const s_empty = s_full + STAGES;
  __shared__ double red_sm[256 + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      kb_mbar_init(&s_full[s], 1);
      kb_mbar_init(&s_empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;
  if (warp == 8) {
    if ((tid & 31) == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // x windows are re-read by later tiles: evict_last (kb_tune 12: 1 = no hint, 2 = evict_first)
      const uint64_t pol_keep = l2pol == 2 ? kb_policy_evict_first() : kb_policy_evict_last();
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r0 = tile * TR;
        kb_mbar_wait(&s_empty[stage], phase ^ 1u);
        uint32_t bytes = 0;
        for (int g = 0; g < nw; ++g) {
          const int ge = min(r0 + pat.wlo[g] + TR + pat.wspan[g], n_cols);
          bytes += (uint32_t)max(((ge + 1) & ~1) - kb_win_start(r0, pat.wlo[g]), 0) * 8u;
        }
        kb_mbar_expect_tx(&s_full[stage], bytes);
        for (int g = 0; g < nw; ++g) {
          const int ge = min(r0 + pat.wlo[g] + TR + pat.wspan[g], n_cols);
          const int ga = kb_win_start(r0, pat.wlo[g]);
          const int gn = max(((ge + 1) & ~1) - ga, 0);
          if (gn > 0) {
            if (l2pol == 1)
              kb_bulk_g2s(s_win + ((size_t)stage * nw + g) * wlen, x + ga, (uint32_t)gn * 8u,
                          &s_full[stage]);
            else
              kb_bulk_g2s_hint(s_win + ((size_t)stage * nw + g) * wlen, x + ga, (uint32_t)gn * 8u,
                               &s_full[stage], pol_keep);
          }
        }
        if (bytes == 0) kb_mbar_arrive(&s_full[stage]);  // (cannot happen: n_cols > 0)
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    const int nd = pat.nd;
    const unsigned full = (1u << nd) - 1u;
    // shared-memory index of diagonal d for this thread's row q = 0, valid while the windows
    // are not clamped at 0 (r0 + wlo >= 0; r0 is even): idx = row + off - ((r0 + wlo) & ~1)
    int so[8];
#pragma unroll
    for (int d = 0; d < 8; ++d)
      so[d] = pat.grp[d] * wlen + tid + pat.off[d] - pat.dwlo[d] + (pat.dwlo[d] & 1);
    const bool wx = DOT == 1 && d_center >= 0;
    const int sc = center_base + tid;  // so[d_center], computed by the host
    const double cf = (mode == 1) ? coef[0] : 0.0;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int r0 = tile * TR;
      const bool unclamped = r0 + pat.wlo[0] >= 0;  // wlo[0] is the lowest window
      unsigned m[RPT];
      double zv[RPT], wv[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int row = r0 + tid + q * 256;
        m[q] = 0;
        zv[q] = wv[q] = 0.0;
        if (row < n_rows) {
          m[q] = masks[row];
          if (mode != 0) zv[q] = z[row];
          if (DOT == 1 && !wx) wv[q] = w[row];
        }
      }
      kb_mbar_wait(&s_full[stage], phase);
      const double* sw = s_win + (size_t)stage * nw * wlen;
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int row = r0 + tid + q * 256;
        double sum = 0.0;
        if (unclamped && __all_sync(0xffffffffu, m[q] == full)) {
#pragma unroll
          for (int d = 0; d < 8; ++d)
            if (d < nd) sum = __dadd_rn(sum, __dmul_rn(cv.c[d], sw[so[d] + q * 256]));
          if (wx) wv[q] = sw[sc + q * 256];
        } else {
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            if (d < nd && (m[q] >> d) & 1u) {
              const int idx = row + pat.off[d] - kb_win_start(r0, pat.dwlo[d]);
              sum = __dadd_rn(sum, __dmul_rn(cv.c[d], sw[pat.grp[d] * wlen + idx]));
            }
          }
          if (wx && row < n_rows) wv[q] = w[row];
        }
        if (row < n_rows) {
          double yv = sum;
          if (mode == 1) yv = kb_mul_sub(cf, zv[q], sum);
          if (mode == 2) yv = __dsub_rn(zv[q], sum);
          __stcs(&y[row], yv);
          if (DOT == 1) acc = fma(wv[q], yv, acc);
          if (DOT == 2) acc = fma(yv, yv, acc);
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) kb_mbar_arrive(&s_empty[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm);
}

// ------------------------------ constant-diagonal stencil kernel, second version ---
// ncu of kb_spmv_stencil_kernel at 512^3 (profiles/r1_stencil_v1_ncu.txt): 169 instructions per
// row-warp, issue slots 68 % busy, DRAM 39 % -- the row sum was compiled into 11 instructions per
// diagonal (the generic->shared address conversion S2UR/UMOV/ULEA redone for every load, a
// `d < nd` branch per diagonal, each DMUL waiting on its own LDS) and the warps stalled on the
// per-tile mask load right after the barrier.  This version
//   * takes the number of diagonals as a template parameter (no branches in the row sum),
//   * keeps one 32-bit shared-memory address per diagonal and thread and issues
//     `ld.shared.f64 [addr + imm]` directly, all loads of a row before its multiply/add chain,
//   * prefetches the next tile's masks one tile ahead.
// Products, their order and the rounding of each are those of kb_spmv_stencil_kernel.
template <int IMM>
__device__ __forceinline__ double kb_lds_f64(uint32_t addr) {
  double v;
  // volatile: stays between the (volatile) barrier wait and arrive; no memory clobber, so the
  // arithmetic and the global stores of neighbouring rows may be scheduled around it
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
  return v;
}

template <int ND, int Q>
__device__ __forceinline__ void kb_st2_load(const uint32_t (&a)[ND], double (&xv)[ND]) {
#pragma unroll
  for (int d = 0; d < ND; ++d) xv[d] = kb_lds_f64<Q * 2048>(a[d]);
}

template <int ND>
__device__ __forceinline__ double kb_st2_sum(const double (&xv)[ND], const KbConstVals& cv) {
  double sum = 0.0;
#pragma unroll
  for (int d = 0; d < ND; ++d) sum = __dadd_rn(sum, __dmul_rn(cv.c[d], xv[d]));
  return sum;
}

__device__ __forceinline__ void kb_st2_finish(double sum, int row, int mode, double cf, double zv,
                                              double wv, double* __restrict__ y, int dot,
                                              double& acc) {
  double yv = sum;
  if (mode == 1) yv = kb_mul_sub(cf, zv, sum);
  if (mode == 2) yv = __dsub_rn(zv, sum);
  __stcs(&y[row], yv);
  if (dot == 1) acc = fma(wv, yv, acc);
  if (dot == 2) acc = fma(yv, yv, acc);
}

// NQ (1 or 2) rows of the tile, Q0-th and following rows of this thread (row = r0 + tid + 256 q).
// Interior warps (every mask of the group full, windows not clamped: 98 % at 512^3) take the
// branch-free path: all loads first, then the multiply/add chains.  WX: the dot operand w is x
// itself and the middle diagonal has offset 0 -- it is the value already loaded for the sum.
template <int ND, int Q0, int NQ, int DOT, bool WX>
__device__ __forceinline__ void kb_st2_rows(const uint32_t (&a)[ND], bool unclamped,
                                            const unsigned* m, const double* zv, const double* wv,
                                            int r0, int n_rows, const double* sw, int wlen,
                                            const KbPattern& pat, const KbConstVals& cv, int mode,
                                            double cf, const double* __restrict__ w,
                                            double* __restrict__ y, double& acc) {
  constexpr unsigned full = (1u << ND) - 1u;
  const int tid = threadIdx.x;
  bool allfull = unclamped && m[Q0] == full;
  if constexpr (NQ == 2) allfull = allfull && m[Q0 + 1] == full;
  if (__all_sync(0xffffffffu, allfull)) {
    double xa[ND], xb[ND];
    kb_st2_load<ND, Q0>(a, xa);
    if constexpr (NQ == 2) kb_st2_load<ND, Q0 + 1>(a, xb);
    const double sa = kb_st2_sum<ND>(xa, cv);
    kb_st2_finish(sa, r0 + tid + Q0 * 256, mode, cf, zv[Q0], WX ? xa[ND / 2] : wv[Q0], y, DOT, acc);
    if constexpr (NQ == 2) {
      const double sb = kb_st2_sum<ND>(xb, cv);
      kb_st2_finish(sb, r0 + tid + (Q0 + 1) * 256, mode, cf, zv[Q0 + 1],
                    WX ? xb[ND / 2] : wv[Q0 + 1], y, DOT, acc);
    }
  } else {
#pragma unroll
    for (int q = Q0; q < Q0 + NQ; ++q) {
      const int row = r0 + tid + q * 256;
      double sum = 0.0;
#pragma unroll
      for (int d = 0; d < ND; ++d) {
        if ((m[q] >> d) & 1u) {
          const int idx = row + pat.off[d] - kb_win_start(r0, pat.dwlo[d]);
          sum = __dadd_rn(sum, __dmul_rn(cv.c[d], sw[pat.grp[d] * wlen + idx]));
        }
      }
      if (row < n_rows) kb_st2_finish(sum, row, mode, cf, zv[q], WX ? w[row] : wv[q], y, DOT, acc);
    }
  }
}

template <int ND, int RPT, int STAGES, int MINB, int DOT, bool WX>
__global__ void __launch_bounds__(256 + 32, MINB)
kb_spmv_stencil2_kernel(int n_rows, int n_cols, int n_tiles, int wlen,
                        const uint16_t* __restrict__ masks, KbPattern pat, KbConstVals cv,
                        const double* __restrict__ x, double* __restrict__ y, int mode,
                        const double* __restrict__ z, const double* __restrict__ coef,
                        const double* __restrict__ w, int l2pol, double* __restrict__ out,
                        KbRed rd) {
  static_assert(RPT == 1 || RPT == 2 || RPT == 4, "rows per thread");
  if (kb_gated(rd)) return;
  constexpr int TR = 256 * RPT;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  double* const s_win = reinterpret_cast<double*>(kb_dyn_smem);
  const int nw = pat.nw;
  uint64_t* const s_full = reinterpret_cast<uint64_t*>(s_win + (size_t)STAGES * nw * wlen);
  uint64_t* const s_empty = s_full + STAGES;
  __shared__ double red_sm[256 + 32];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      kb_mbar_init(&s_full[s], 1);
      kb_mbar_init(&s_empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;
  if (warp == 8) {
    if ((tid & 31) == 0) {  // producer lane: the x windows of every tile of this CTA
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_keep = l2pol == 2 ? kb_policy_evict_first() : kb_policy_evict_last();
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r0 = tile * TR;
        kb_mbar_wait(&s_empty[stage], phase ^ 1u);
        uint32_t bytes = 0;
        for (int g = 0; g < nw; ++g) {
          const int ge = min(r0 + pat.wlo[g] + TR + pat.wspan[g], n_cols);
          bytes += (uint32_t)max(((ge + 1) & ~1) - kb_win_start(r0, pat.wlo[g]), 0) * 8u;
        }
        kb_mbar_expect_tx(&s_full[stage], bytes);
        for (int g = 0; g < nw; ++g) {
          const int ge = min(r0 + pat.wlo[g] + TR + pat.wspan[g], n_cols);
          const int ga = kb_win_start(r0, pat.wlo[g]);
          const int gn = max(((ge + 1) & ~1) - ga, 0);
          if (gn > 0) {
            if (l2pol == 1)
              kb_bulk_g2s(s_win + ((size_t)stage * nw + g) * wlen, x + ga, (uint32_t)gn * 8u,
                          &s_full[stage]);
            else
              kb_bulk_g2s_hint(s_win + ((size_t)stage * nw + g) * wlen, x + ga, (uint32_t)gn * 8u,
                               &s_full[stage], pol_keep);
          }
        }
        if (bytes == 0) kb_mbar_arrive(&s_full[stage]);  // (cannot happen: n_cols > 0)
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // byte address (shared window) of diagonal d for this thread's row q = 0 in stage 0, valid
    // while the windows are not clamped at 0: element  row + off - ((r0 + wlo) & ~1), r0 even
    const uint32_t sbase = kb_smem_u32(s_win);
    uint32_t a0[ND];
#pragma unroll
    for (int d = 0; d < ND; ++d)
      a0[d] = sbase + 8u * (uint32_t)(pat.grp[d] * wlen + tid + pat.off[d] - pat.dwlo[d] +
                                      (pat.dwlo[d] & 1));
    const uint32_t stage_bytes = (uint32_t)(nw * wlen) * 8u;
    const double cf = (mode == 1) ? coef[0] : 0.0;
    int stage = 0;
    uint32_t phase = 0;
    unsigned mn[RPT];  // masks of the next tile, fetched one tile ahead
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      const int row = blockIdx.x * TR + tid + q * 256;
      mn[q] = (blockIdx.x < n_tiles && row < n_rows) ? masks[row] : 0u;
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int r0 = tile * TR;
      const bool unclamped = r0 + pat.wlo[0] >= 0;  // wlo[0] is the lowest window
      unsigned m[RPT];
      double zv[RPT], wv[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int row = r0 + tid + q * 256;
        m[q] = mn[q];
        zv[q] = wv[q] = 0.0;
        if (row < n_rows) {
          if (mode != 0) zv[q] = z[row];
          if (DOT == 1 && !WX) wv[q] = w[row];
        }
        const int nrow = row + gridDim.x * TR;
        mn[q] = (tile + (int)gridDim.x < n_tiles && nrow < n_rows) ? masks[nrow] : 0u;
      }
      kb_mbar_wait(&s_full[stage], phase);
      const uint32_t so = (uint32_t)stage * stage_bytes;
      uint32_t a[ND];
#pragma unroll
      for (int d = 0; d < ND; ++d) a[d] = a0[d] + so;
      const double* sw = s_win + (size_t)stage * nw * wlen;
      if constexpr (RPT == 1)
        kb_st2_rows<ND, 0, 1, DOT, WX>(a, unclamped, m, zv, wv, r0, n_rows, sw, wlen, pat, cv, mode,
                                       cf, w, y, acc);
      if constexpr (RPT >= 2)
        kb_st2_rows<ND, 0, 2, DOT, WX>(a, unclamped, m, zv, wv, r0, n_rows, sw, wlen, pat, cv, mode,
                                       cf, w, y, acc);
      if constexpr (RPT == 4)
        kb_st2_rows<ND, 2, 2, DOT, WX>(a, unclamped, m, zv, wv, r0, n_rows, sw, wlen, pat, cv, mode,
                                       cf, w, y, acc);
      __syncwarp();
      if ((tid & 31) == 0) kb_mbar_arrive(&s_empty[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm);
}

// ------------------------------------------ windowed pattern kernel, k > 1 ----
// SpMM for k right-hand sides in lock-step (x, y are (n, k) row-major): the same
// compressed matrix and TMA x windows; a window is now rows_t + span consecutive
// x ROWS (k doubles each, contiguous).  256 consumer threads = (256/k) rows x k
// columns per pass, RPT passes per tile (rows_t = RPT * 256 / k), so the shared
// memory per stage is independent of k.  Requires 256 % k == 0 and 16-byte
// aligned x (k even makes every window start 16-byte aligned).
template <int STAGES, int RPT, int MINB, int DOT>
__global__ void __launch_bounds__(288, MINB)
kb_spmm_window_kernel(int n_rows, int n_cols, int n_tiles, int k, int cap, int wlen,
                      const int32_t* __restrict__ rowptr, const uint16_t* __restrict__ masks,
                      const double* __restrict__ vals, KbPattern pat,
                      const double* __restrict__ x, double* __restrict__ y, int mode,
                      const double* __restrict__ z, const double* __restrict__ coef,
                      const double* __restrict__ w, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  const int nw = pat.nw;
  const int rows_pass = 256 / k;
  const int rows_t = RPT * rows_pass;
  double* const s_vals = reinterpret_cast<double*>(kb_dyn_smem);
  double* const s_win = s_vals + (size_t)STAGES * cap;
  uint64_t* const s_full =
      reinterpret_cast<uint64_t*>(s_win + (size_t)STAGES * nw * wlen * k);
  uint64_t* const s_empty = s_full + STAGES;
  __shared__ double red_sm[288];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      kb_mbar_init(&s_full[s], 1);
      kb_mbar_init(&s_empty[s], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double acc = 0.0;

  if (warp == 8) {
    const int lane = tid & 31;
    int stage = 0;
    uint32_t phase = 0;
    const uint64_t pol_stream = kb_policy_evict_first();
    const uint64_t pol_keep = kb_policy_evict_last();
    for (int64_t base = blockIdx.x; base < n_tiles; base += 32ll * gridDim.x) {
      const int64_t my_tile = base + (int64_t)lane * gridDim.x;
      int my_s = 0, my_e = 0;
      if (my_tile < n_tiles) {
        const int r0 = (int)my_tile * rows_t;
        const int r1 = min(r0 + rows_t, n_rows);
        my_s = rowptr[r0];
        my_e = rowptr[r1];
      }
      for (int q = 0; q < 32; ++q) {
        const int64_t tile = base + (int64_t)q * gridDim.x;
        if (tile >= n_tiles) break;
        const int s = __shfl_sync(0xffffffffu, my_s, q);
        const int e = __shfl_sync(0xffffffffu, my_e, q);
        if (lane == 0 && e > s) {
          const int r0 = (int)tile * rows_t;
          const int a0 = s & ~3, a1 = (e + 3) & ~3;
          kb_mbar_wait(&s_empty[stage], phase ^ 1u);
          uint32_t bytes = (uint32_t)(a1 - a0) * 8u;
          for (int g = 0; g < nw; ++g) {
            const int gs = max(r0 + pat.wlo[g], 0);
            const int ge = min(r0 + pat.wlo[g] + rows_t + pat.wspan[g], n_cols);
            bytes += (uint32_t)max(ge - gs, 0) * (uint32_t)k * 8u;
          }
          kb_mbar_expect_tx(&s_full[stage], bytes);
          kb_bulk_g2s_hint(s_vals + (size_t)stage * cap, vals + a0, (uint32_t)(a1 - a0) * 8u,
                           &s_full[stage], pol_stream);
          for (int g = 0; g < nw; ++g) {
            const int gs = max(r0 + pat.wlo[g], 0);
            const int ge = min(r0 + pat.wlo[g] + rows_t + pat.wspan[g], n_cols);
            if (ge > gs)
              kb_bulk_g2s_hint(s_win + ((size_t)stage * nw + g) * wlen * k, x + (size_t)gs * k,
                               (uint32_t)(ge - gs) * (uint32_t)k * 8u, &s_full[stage], pol_keep);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        __syncwarp();
      }
    }
  } else {
    int stage = 0;
    uint32_t phase = 0;
    const int c = tid % k;
    const int rs = tid / k;
    const double cf = (mode == 1) ? coef[c] : 0.0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int r0 = tile * rows_t;
      const int r1 = min(r0 + rows_t, n_rows);
      const int s = rowptr[r0], e = rowptr[r1];
      int lo[RPT];
      unsigned mk[RPT];
      double zv[RPT], wv[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int row = r0 + rs + q * rows_pass;
        lo[q] = 0;
        mk[q] = 0;
        zv[q] = 0.0;
        wv[q] = 0.0;
        if (row < n_rows) {
          lo[q] = rowptr[row];
          mk[q] = masks[row];
          if (mode != 0) zv[q] = z[(size_t)row * k + c];
          if (DOT == 1) wv[q] = w[(size_t)row * k + c];
        }
      }
      double sum[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) sum[q] = 0.0;
      if (e > s) {
        kb_mbar_wait(&s_full[stage], phase);
        const double* sv = s_vals + (size_t)stage * cap;
        const double* sw = s_win + (size_t)stage * nw * wlen * k;
        const int a0 = s & ~3;
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          const int row = r0 + rs + q * rows_pass;
          int j = lo[q] - a0;
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            if (d < pat.nd && (mk[q] >> d) & 1u) {
              const int xr = row + pat.off[d] - max(r0 + pat.dwlo[d], 0);
              const double xv = sw[((size_t)pat.grp[d] * wlen + xr) * k + c];
              sum[q] = __dadd_rn(sum[q], __dmul_rn(sv[j], xv));
              ++j;
            }
          }
        }
        __syncwarp();
        if ((tid & 31) == 0) kb_mbar_arrive(&s_empty[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        const int row = r0 + rs + q * rows_pass;
        if (row < n_rows) {
          double yv = sum[q];
          if (mode == 1) yv = kb_mul_sub(cf, zv[q], sum[q]);
          if (mode == 2) yv = __dsub_rn(zv[q], sum[q]);
          __stcs(&y[(size_t)row * k + c], yv);
          if (DOT == 1) acc = fma(wv[q], yv, acc);
          if (DOT == 2) acc = fma(yv, yv, acc);
        }
      }
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, k, rd, out, red_sm);
}

// ------------------------------------------------- boundary rows (halo part) --
// Row-partitioned matrices: the local product runs on the
// columns a rank owns while the halo entries travel; this kernel then finishes
// the boundary rows.  for i < n_brows: row = rows[i];
//   h = sum_j hval[j] * xh[hcol[j], c];   y[row, c] += sign * h
//   DOT 1: out[c] += sum_i w[row, c] * sign * h     (added to the local product's dot)
template <int DOT>
__global__ void __launch_bounds__(KB_BLOCK)
kb_spmv_halo_add_kernel(int64_t n_brows, int k, double sign, const int32_t* __restrict__ rows,
                        const int32_t* __restrict__ hrowptr, const int32_t* __restrict__ hcol,
                        const double* __restrict__ hval, const double* xh,
                        double* __restrict__ y, const double* __restrict__ w,
                        double* __restrict__ out, KbHalo hd, const int* __restrict__ srcs,
                        int n_src, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  unsigned char* own = nullptr;
  unsigned long long q = 0;
  bool dead = false;  // a source never delivered: the receive area is stale -> poison y and the dot
  if (hd.peers != nullptr) {
    // peer-memory halo: the entries were pushed into this rank's data area; wait for the
    // flag of every source of product q (the push of this product already counted it)
    own = hd.peers[hd.rank];
    q = *kb_halo_u64(own, KB_HALO_RECV_COUNTER) + 1ull;  // bumped by the last block below
    if ((int)threadIdx.x < n_src)
      kb_halo_wait(kb_halo_u64(own, KB_HALO_FLAGS + 8 * (size_t)srcs[threadIdx.x]), q, own);
    __syncthreads();
    __threadfence_system();
    dead = *reinterpret_cast<volatile int*>(own + KB_HALO_ERROR) != 0;  // sticky
    xh = reinterpret_cast<const double*>(own + KB_HALO_DATA);
  }
  const int c = threadIdx.x % k;
  const int rows_per_block = blockDim.x / k;
  const int rsub = threadIdx.x / k;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * rows_per_block + rsub; i < n_brows;
       i += (int64_t)gridDim.x * rows_per_block) {
    const int lo = hrowptr[i], hi = hrowptr[i + 1];
    double h = 0.0;
    for (int j = lo; j < hi; ++j)
      h = __dadd_rn(h, __dmul_rn(hval[j], __ldcv(xh + (size_t)hcol[j] * k + c)));
    h *= sign;
    if (dead) h = nan("");
    const size_t idx = (size_t)rows[i] * k + c;
    y[idx] = __dadd_rn(y[idx], h);
    if (DOT == 1) acc = fma(w[idx], h, acc);
  }
  // all blocks done reading -> acknowledge product q to the sources (flow control).  With a
  // fused dot the reduction's own arrival ticket already identifies the last block.
  if (DOT != 0) {
    const bool last = kb_grid_colsum(acc, k, rd, out, sm, /*accumulate=*/true);
    if (hd.peers != nullptr && last) {
      if ((int)threadIdx.x < n_src)
        *kb_halo_u64(hd.peers[srcs[threadIdx.x]], KB_HALO_ACKS + 8 * (size_t)hd.rank) = q;
      if (threadIdx.x == 0) *kb_halo_u64(own, KB_HALO_RECV_COUNTER) = q;
    }
  } else if (hd.peers != nullptr) {
    __shared__ int s_done;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long* done = reinterpret_cast<unsigned long long*>(own + KB_HALO_DONE);
      const unsigned long long prev = atomicAdd(done, 1ull);
      s_done = (prev == (unsigned long long)gridDim.x - 1ull) ? 1 : 0;
      if (s_done) *done = 0ull;
    }
    __syncthreads();
    if (s_done) {
      if ((int)threadIdx.x < n_src)
        *kb_halo_u64(hd.peers[srcs[threadIdx.x]], KB_HALO_ACKS + 8 * (size_t)hd.rank) = q;
      if (threadIdx.x == 0) *kb_halo_u64(own, KB_HALO_RECV_COUNTER) = q;
    }
  }
}

// ----------------------------------------------------------- row statistics --
// out[0] = longest row; out[1] |= 1: rowptr[0] != 0, 2: a row pointer decreases, 4: rowptr[n] != nnz
// (structure check of kb_csr_create: a malformed matrix is refused instead of read out of bounds)
__global__ void kb_max_row_len_kernel(int64_t n_rows, int64_t nnz,
                                      const int32_t* __restrict__ rowptr, int* out) {
  int m = 0, bad = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int len = rowptr[r + 1] - rowptr[r];
    if (len < 0) bad |= 2;
    m = max(m, len);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (rowptr[0] != 0) bad |= 1;
    if ((int64_t)rowptr[n_rows] != nnz) bad |= 4;
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, m);
    if (bad) atomicOr(out + 1, bad);
  }
}

// out[1] |= 8 if a column index lies outside [0, n_cols)
__global__ void __launch_bounds__(256)
kb_colidx_check_kernel(int64_t nnz, int64_t n_cols, const int32_t* __restrict__ colidx, int* out) {
  int bad = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < nnz;
       j += (int64_t)gridDim.x * blockDim.x) {
    const int c = colidx[j];
    if (c < 0 || (int64_t)c >= n_cols) bad = 8;
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && bad) atomicOr(out + 1, bad);
}

// ------------------------------------------------------- 7-point generator --
// Entry order per row: z-1, y-1, x-1, diag, x+1, y+1, z+1 (ascending column).
struct KbStencil7 {
  int nx, ny, nz, z_lo, z_hi;
  double c[7];  // lz, ly, lx, diag, ux, uy, uz  (already in entry order)
};

__global__ void __launch_bounds__(KB_BLOCK)
kb_stencil7_kernel(KbStencil7 p, int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx,
                   double* __restrict__ vals) {
  const int64_t plane = (int64_t)p.nx * p.ny;
  const int64_t n_loc = plane * (p.z_hi - p.z_lo);
  const int64_t offs[7] = {-plane, -(int64_t)p.nx, -1, 0, 1, (int64_t)p.nx, plane};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_loc;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i + plane * p.z_lo;
    const int ix = (int)(g % p.nx);
    const int iy = (int)((g / p.nx) % p.ny);
    const int iz = (int)(g / plane);
    const bool ok[7] = {iz > 0, iy > 0, ix > 0, true, ix < p.nx - 1, iy < p.ny - 1, iz < p.nz - 1};
    if (vals == nullptr) {
      int cnt = 0;
#pragma unroll
      for (int q = 0; q < 7; ++q) cnt += ok[q] ? 1 : 0;
      rowptr[i + 1] = cnt;
      if (i == 0) rowptr[0] = 0;
    } else {
      int j = rowptr[i];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        if (ok[q]) {
          colidx[j] = (int32_t)(g + offs[q]);
          vals[j] = p.c[q];
          ++j;
        }
      }
    }
  }
}
