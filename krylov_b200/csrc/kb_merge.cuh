// CSR SpMV for long or skewed rows: nonzero-balanced tiles (schedule "merge").
//
//   t = A x ;  y = t | t - coef*z | z - t ;  out = <w,y> | <y,y> | -
//
// The stream kernel gives one thread one row, which is right for short even rows and wrong for
// rows of hundreds of entries or for a few very long ones (`_helpers.py:47` multiplies with
// whatever the user hands in).  Here the unit of work is a tile of T consecutive NONZEROS, not
// of rows, so every CTA streams the same number of matrix bytes whatever the row lengths are
// (the nonzero axis of a merge-path decomposition):
//   * a producer lane brings the tile's vals / colidx into a shared-memory ring with two 1-D
//     TMA bulk copies (the matrix is read from HBM once, in 128-byte lines);
//   * phase 1: all 256 consumer threads form the products vals[j] * x[colidx[j]] in place
//     (128-bit LDS of two values, 64-bit LDS of two indices, independent gathers in flight);
//   * phase 2: the rows of the tile are summed out of shared memory by groups of G lanes,
//     G = 1 ... 32 chosen per tile from its row count (G = 1: strictly left to right, i.e.
//     SciPy's csr_matvec order; G > 1: G strided partial sums + butterfly, a fixed tree).
// Rows do not end where tiles end:
//   * a row that runs at most TAIL entries past the end of the tile it STARTS in is finished by
//     that tile -- the tail's values / indices come straight from global memory (the next
//     tile's TMA brings the same lines, so they cost L2 traffic, not DRAM traffic), its products
//     sit behind the tile's in shared memory, and the next tile skips the row: the summation
//     order is unchanged;
//   * a longer row ("long row": it crosses a tile boundary by more than TAIL) leaves one partial
//     sum per tile in carry[] and is finished by a second, small launch
//     (kb_merge_fix_kernel: one warp per long row adds its carries in a fixed order, applies
//     the epilogue, and adds the row's share onto the fused dot).  No CTA ever waits for
//     another one, so nothing depends on co-residency or timing.
// Per-tile bookkeeping {first row, last row, tail} and the list of long rows are computed once
// per matrix (kb_merge_tiles_kernel, kb_merge_fixlist_kernel); row pointers of the next tile
// are requested while the current one is summed, so no dependent global load sits between two
// tiles.  Tiles can be dealt to CTAs so that the CTAs of one SM work on neighbouring tiles at
// the same time (their x gathers then share L1 lines).
// Bytes per launch: 12 nnz + 4 (n+1) + 16 n (+ 16 per tile), as for the stream kernel.
#pragma once
#include "kb_handles.cuh"
#include "kb_ptx.cuh"

// NG consumer groups of 256 threads per CTA, each with its own ring; the groups take
// neighbouring tiles at the same time, so their x gathers share the SM's L1.
template <int T, int TAIL, int STAGES, int NG>
struct KbMergeSmem {
  double vals[NG][STAGES][T + TAIL];
  int32_t cols[NG][STAGES][T];
  uint64_t full[NG][STAGES];
  uint64_t empty[NG][STAGES];
};

// lanes per row of a tile (log2), from a cost model of phase 2: a pass over 256 / G rows costs
// about 125 cycles of bookkeeping, a lane's serial add about 30, a butterfly step about 20 --
//   cost(G) = ceil(nr / (256 / G)) * 125 + ceil(maxlen / G) * 30 + log2(G) * 20.
// Few long rows get many lanes, many short rows one lane each (G = 1 sums left to right: SciPy's
// order), and a tile that mixes one long row with many short ones lands in between instead of
// serialising either way.
__device__ __forceinline__ int kb_merge_lg(int nr, int maxlen) {
  int best = 0;
  long long bestc = -1;
  for (int lg = 0; lg <= 5; ++lg) {
    const long long passes = (nr + (256 >> lg) - 1) / (256 >> lg);
    const long long c = passes * 125 + (long long)((maxlen + (1 << lg) - 1) >> lg) * 30 + lg * 20;
    if (bestc < 0 || c < bestc) {
      bestc = c;
      best = lg;
    }
  }
  return best;
}

// meta[t] = {first row this tile sums, last row, tail entries past the tile it finishes,
//            bit 0: the tile's first row is a long row that ENDS here; bits 8..: log2 lanes per row}
__global__ void __launch_bounds__(256)
kb_merge_tiles_kernel(int n_rows, int nnz, int n_tiles, int T, int TAIL,
                      const int32_t* __restrict__ rowptr, int4* __restrict__ meta) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  // row that owns nonzero key (rowptr[r] <= key < rowptr[r+1]); key 0 -> 0 (leading empty
  // rows belong to tile 0), key >= nnz -> n_rows
  auto owner = [&](long long key) -> int {
    if (key <= 0) return 0;
    if (key >= nnz) return n_rows;
    int lo = 0, hi = n_rows + 1;  // first index i in [0, n_rows] with rowptr[i] > key
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((long long)rowptr[mid] > key) hi = mid;
      else lo = mid + 1;
    }
    return lo - 1;
  };
  const long long a = (long long)t * T;
  const long long b = a + T < nnz ? a + T : nnz;
  const int r_lo = owner(a), r_nx = owner(b);
  int r_last = r_nx - 1;
  if (r_nx < n_rows && rowptr[r_nx] < b) r_last = r_nx;
  int r_first = r_lo;
  int ends_long = 0;
  if (rowptr[r_lo] < a) {  // began in an earlier tile
    const long long b0 = ((long long)rowptr[r_lo] / T + 1) * T;
    if (rowptr[r_lo + 1] - b0 <= TAIL) r_first = r_lo + 1;  // ... which has finished it
    else if (rowptr[r_lo + 1] <= b) ends_long = 1;
  }
  int tail = 0;
  if (r_last >= r_first && rowptr[r_last + 1] > b && rowptr[r_last] >= a &&
      rowptr[r_last + 1] - b <= TAIL)
    tail = (int)(rowptr[r_last + 1] - b);
  int maxlen = 0;
  for (int r = r_first; r <= r_last; ++r) {
    const long long lo = rowptr[r] > a ? rowptr[r] : a;
    const long long hi = rowptr[r + 1] < b + tail ? rowptr[r + 1] : b + tail;
    if (hi - lo > maxlen) maxlen = (int)(hi - lo);
  }
  meta[t] = make_int4(r_first, r_last, tail,
                      ends_long | (kb_merge_lg(r_last - r_first + 1, maxlen) << 8));
}

// Compacts the tiles with meta.w == 1 into fix[] in tile order (one block; runs once per matrix):
// fix[i] = {row, first tile of the row, last tile, 0}.
__global__ void __launch_bounds__(1024)
kb_merge_fixlist_kernel(int n_tiles, int T, const int32_t* __restrict__ rowptr,
                        const int4* __restrict__ meta, int4* __restrict__ fix,
                        int* __restrict__ n_fix) {
  __shared__ int cnt[1024];
  const int t = threadIdx.x;
  const int per = (n_tiles + 1023) / 1024;
  const int t0 = min(t * per, n_tiles), t1 = min(t0 + per, n_tiles);
  int c = 0;
  for (int q = t0; q < t1; ++q) c += meta[q].w & 1;
  cnt[t] = c;
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int i = 0; i < 1024; ++i) {
      const int v = cnt[i];
      cnt[i] = run;
      run += v;
    }
    *n_fix = run;
  }
  __syncthreads();
  if (fix == nullptr) return;  // counting pass
  int o = cnt[t];
  for (int q = t0; q < t1; ++q) {
    const int4 m = meta[q];
    if (m.w & 1) fix[o++] = make_int4(m.x, rowptr[m.x] / T, q, 0);
  }
}

// A CTA works on "super-tiles" of NG neighbouring tiles (one per group).  order: 0 super-tile
// = i * grid + block; 1 the same with the CTAs renumbered so that the `cps` CTAs an SM usually
// holds (blocks b, b + nsm, ...) take neighbouring ones; 2 contiguous ranges per CTA
struct KbMergeOrder {
  int mode, nsm, cps;
};

template <int T, int TAIL, int STAGES, int NG, int MINB, int DOT>
__global__ void __launch_bounds__(NG * 256 + 32, MINB)
kb_spmv_merge_kernel(int n_rows, int nnz, int n_tiles, KbMergeOrder ord,
                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                     const double* __restrict__ vals, const int4* __restrict__ meta,
                     double* __restrict__ carry, const double* __restrict__ x,
                     double* __restrict__ y, int mode, const double* __restrict__ z,
                     const double* __restrict__ coef, const double* __restrict__ w,
                     double* __restrict__ out, KbRed rd) {
  static_assert(TAIL <= 512 && TAIL <= T && T % 512 == 0 && NG >= 1 && NG <= 3, "tile shape");
  if (kb_gated(rd)) return;
  extern __shared__ __align__(128) unsigned char kb_dyn_smem[];
  typedef KbMergeSmem<T, TAIL, STAGES, NG> Smem;
  Smem& S = *reinterpret_cast<Smem*>(kb_dyn_smem);
  __shared__ double red_sm[NG * 256 + 32];

  const int warp = threadIdx.x >> 5;
  const int grp = threadIdx.x >> 8;    // consumer group (NG: the producer warp)
  const int tid = threadIdx.x & 255;   // thread within the group

  if (threadIdx.x == 0) {
    for (int g = 0; g < NG; ++g)
      for (int s = 0; s < STAGES; ++s) {
        kb_mbar_init(&S.full[g][s], 1);
        kb_mbar_init(&S.empty[g][s], 8);  // one arrive per consumer warp of the group
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // this CTA's super-tiles: u0, u0 + ustep, ... < u_end; group g takes tile u * NG + g
  const int n_super = (n_tiles + NG - 1) / NG;
  int u0, ustep, u_end = n_super;
  {
    int vb = blockIdx.x;
    if (ord.mode == 1 && (int)gridDim.x == ord.nsm * ord.cps)
      vb = (vb % ord.nsm) * ord.cps + vb / ord.nsm;
    if (ord.mode == 2) {
      const int per = (n_super + gridDim.x - 1) / gridDim.x;
      u0 = min(vb * per, n_super);
      u_end = min(u0 + per, n_super);
      ustep = 1;
    } else {
      u0 = vb;
      ustep = gridDim.x;
    }
  }

  double acc = 0.0;

  if (warp == NG * 8) {
    // ===================== producer warp: lane g feeds group g's ring ================
    const int g = threadIdx.x & 31;
    if (g < NG) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol = kb_policy_evict_first();
      for (int u = u0; u < u_end; u += ustep) {
        const int tile = u * NG + g;
        if (tile >= n_tiles) break;
        const int a = tile * T;
        const uint32_t cnt = (uint32_t)((min(T, nnz - a) + 3) & ~3);
        kb_mbar_wait(&S.empty[g][stage], phase ^ 1u);
        kb_mbar_expect_tx(&S.full[g][stage], cnt * 12u);
        kb_bulk_g2s_hint(&S.vals[g][stage][0], vals + a, cnt * 8u, &S.full[g][stage], pol);
        kb_bulk_g2s_hint(&S.cols[g][stage][0], colidx + a, cnt * 4u, &S.full[g][stage], pol);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== consumers ===================================================
    int stage = 0;
    uint32_t phase = 0;
    const double cf = (mode == 1) ? coef[0] : 0.0;
    const int step = ustep * NG;
    const int tile_end = min(u_end * NG, n_tiles);
    int tile = u0 * NG + grp;
    // software pipeline: bookkeeping two tiles ahead, first-pass row pointers one tile ahead
    int4 m0 = make_int4(0, -1, 0, 0), m1 = m0;
    if (tile < tile_end) m0 = meta[tile];
    if (tile + step < tile_end) m1 = meta[tile + step];
    int lo = 0, hi = 0;
    {
      const int row0 = m0.x + (tid >> (m0.w >> 8));
      if (tile < tile_end && row0 <= m0.y) {
        lo = rowptr[row0];
        hi = rowptr[row0 + 1];
      }
    }
    for (; tile < tile_end; tile += step) {
      const int a = tile * T;
      const int cnt = min(T, nnz - a);
      const int b = a + cnt;
      const int r_first = m0.x, r_last = m0.y, tail = m0.z;
      const int nr = r_last - r_first + 1;
      const int lg = m0.w >> 8;
      const int G = 1 << lg;
      const int ngroups = 256 >> lg;
      const int gl = tid & (G - 1);
      int row = r_first + (tid >> lg);
      // bookkeeping of the tile after the next one
      int4 m2 = make_int4(0, -1, 0, 0);
      if (tile + 2 * step < tile_end) m2 = meta[tile + 2 * step];
      // first-pass row pointers of the next tile
      int lo_t = 0, hi_t = 0;
      {
        const int row1 = m1.x + (tid >> (m1.w >> 8));
        if (tile + step < tile_end && row1 <= m1.y) {
          lo_t = rowptr[row1];
          hi_t = rowptr[row1 + 1];
        }
      }
      // the tail of the row this tile finishes past its end: straight from global memory
      int tc0 = 0, tc1 = 0;
      double tv0 = 0.0, tv1 = 0.0;
      if (tid < tail) {
        tc0 = colidx[b + tid];
        tv0 = vals[b + tid];
      }
      if (TAIL > 256 && tid + 256 < tail) {
        tc1 = colidx[b + tid + 256];
        tv1 = vals[b + tid + 256];
      }

      kb_mbar_wait(&S.full[grp][stage], phase);
      double* sv = &S.vals[grp][stage][0];
      const int32_t* sc = &S.cols[grp][stage][0];
      // ---- phase 1: products in place
      {
        double2 v[T / 512];
        int2 c[T / 512];
        double x0[T / 512], x1[T / 512];
#pragma unroll
        for (int i = 0; i < T / 512; ++i) {
          const int j = 2 * tid + 512 * i;
          v[i] = *reinterpret_cast<const double2*>(sv + j);
          c[i] = *reinterpret_cast<const int2*>(sc + j);
        }
#pragma unroll
        for (int i = 0; i < T / 512; ++i) {
          const int j = 2 * tid + 512 * i;
          x0[i] = (j < cnt) ? __ldg(x + c[i].x) : 0.0;
          x1[i] = (j + 1 < cnt) ? __ldg(x + c[i].y) : 0.0;
        }
        double tx0 = 0.0, tx1 = 0.0;
        if (tid < tail) tx0 = __ldg(x + tc0);
        if (TAIL > 256 && tid + 256 < tail) tx1 = __ldg(x + tc1);
#pragma unroll
        for (int i = 0; i < T / 512; ++i) {
          const int j = 2 * tid + 512 * i;
          double2 p;
          p.x = __dmul_rn(v[i].x, x0[i]);
          p.y = __dmul_rn(v[i].y, x1[i]);
          *reinterpret_cast<double2*>(sv + j) = p;
        }
        if (tid < tail) sv[T + tid] = __dmul_rn(tv0, tx0);  // tail > 0 only in full tiles
        if (TAIL > 256 && tid + 256 < tail) sv[T + tid + 256] = __dmul_rn(tv1, tx1);
      }
      asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
      // ---- phase 2: row sums
      const int npass = (nr + ngroups - 1) >> (8 - lg);
      const int bt = b + tail;
      for (int pass = 0; pass < npass; ++pass) {
        const bool valid = row <= r_last;
        const int row_n = row + ngroups;
        int lo_n = 0, hi_n = 0;
        if (pass + 1 < npass && row_n <= r_last) {
          lo_n = rowptr[row_n];
          hi_n = rowptr[row_n + 1];
        }
        double s = 0.0;
        const int jb = max(lo, a) - a, je = min(hi, bt) - a;
        for (int j = jb + gl; j < je; j += G) s = __dadd_rn(s, sv[j]);
        for (int o = G >> 1; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (valid && gl == 0) {
          if (lo < a) {
            carry[2 * (size_t)tile] = s;  // a long row's middle or last piece
          } else if (hi > bt) {
            carry[2 * (size_t)tile + 1] = s;  // a long row's first piece
          } else {
            const double yv = kb_spmv_epilogue(s, mode, z, cf, (size_t)row);
            y[row] = yv;
            if (DOT == 1) acc = fma(w[row], yv, acc);
            if (DOT == 2) acc = fma(yv, yv, acc);
          }
        }
        row = row_n;
        lo = lo_n;
        hi = hi_n;
      }
      // the products were written through the generic proxy; the refill is an async-proxy write
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if ((tid & 31) == 0) kb_mbar_arrive(&S.empty[grp][stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
      m0 = m1;
      m1 = m2;
      lo = lo_t;
      hi = hi_t;
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm);
}

// Second launch when the matrix has long rows: one warp per long row.
//   t = carry[first piece] + carry[middle pieces ...] + carry[last piece]  (32 strided partial
//   sums in tile order + butterfly), then the epilogue, y[row], and the row's share of the dot
//   ADDED onto out[] (and, on a row-partitioned matrix, the all-reduce the first launch left out).
template <int DOT>
__global__ void __launch_bounds__(256)
kb_merge_fix_kernel(int n_fix, const int4* __restrict__ fix, const double* __restrict__ carry,
                    double* __restrict__ y, int mode, const double* __restrict__ z,
                    const double* __restrict__ coef, const double* __restrict__ w,
                    double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double red_sm[256];
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const double cf = (mode == 1) ? coef[0] : 0.0;
  double acc = 0.0;
  for (int i = wid; i < n_fix; i += nw) {
    const int4 f = fix[i];
    double s = 0.0;
    if (lane == 0) s = carry[2 * (size_t)f.y + 1];
    for (int q = f.y + 1 + lane; q <= f.z; q += 32) s = __dadd_rn(s, carry[2 * (size_t)q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (lane == 0) {
      const double yv = kb_spmv_epilogue(s, mode, z, cf, (size_t)f.x);
      y[f.x] = yv;
      if (DOT == 1) acc = fma(w[f.x], yv, acc);
      if (DOT == 2) acc = fma(yv, yv, acc);
    }
  }
  if (DOT != 0) kb_grid_colsum(acc, 1, rd, out, red_sm, true);
}
