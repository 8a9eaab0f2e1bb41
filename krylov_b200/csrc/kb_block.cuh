// Tall-skinny block kernels on the FP64 tensor cores (sm_100a, mma.sync m8n8k4 DMMA).
//
// The one place of the reference where a true block inner product appears is utils.py
// (`inner(QF, QG)` -> a k x l matrix, utils.py:100,117) together with the block updates around
// it (`np.dot(QG, Z.T.conj())`, `RG - np.dot(QF, inner(QF, RG))`, utils.py:101,118,112).  Both are
// n x k by k x l products with n in the millions and k, l <= 16: every byte of the tall operands
// is read once (HBM-bound up to k ~ 32), the k*l multiply-adds per row run on the DMMA pipe.
//
//   kb_block_gram_kernel  : G = X^T Y           (rows are the contraction index)
//   kb_block_apply_kernel : Z = X C | Y -+ X C  (rows are the M index)
//
// Column counts are padded to 8 (HX/HY/HK/HL = number of 8-wide halves, 1 or 2) with
// predicated loads, leading dimensions are free, so column sub-blocks and single columns of
// a row-major (n, k) array are valid operands.
#pragma once
#include "kb_common.cuh"

#define KB_BG_WARPS 8   // warps per CTA
#define KB_BG_STEPS 4   // 4-row MMA steps a warp loads ahead (gram)
#define KB_BA_STEPS 2   // 8-row MMA steps a warp loads ahead (apply)

// D(8x8) += A(8x4) B(4x8): lane t holds A[t>>2][t&3], B[t&3][t>>2], D[t>>2][2(t&3) + {0,1}]
__device__ __forceinline__ void kb_dmma(double& d0, double& d1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ Gram --
// G[i*ldg + j] = sum_r X[r*ldx + i] Y[r*ldy + j], i < k <= 8 HX, j < l <= 8 HY.
// A = X^T tile (8 columns x 4 rows), B = Y tile (4 rows x 8 columns): both fragments are
// "element (row r0 + (t&3), column c0 + (t>>2))", i.e. a warp load covers 4 rows x 64 B.
// Deterministic: warp partials are added in warp order, block partials in block order by the
// last-arriving block (fixed shape) -> bitwise reproducible run to run.
// flags bit 0: store sqrt(|g|).   Gacc: Gacc[i*ldacc + j] += g  (same launch).
template <int HX, int HY>
__global__ void __launch_bounds__(KB_BG_WARPS * 32)
kb_block_gram_kernel(int64_t n, int k, int l, const double* __restrict__ X, int64_t ldx,
                     const double* __restrict__ Y, int64_t ldy, double* G, int64_t ldg,
                     double* Gacc, int64_t ldacc, int flags, KbRed rd) {
  if (kb_gated(rd)) return;
  constexpr int NJ = HY * 8;
  constexpr int KL = HX * 8 * NJ;
  __shared__ double sm[KB_BG_WARPS * KL];
  __shared__ int s_last_b;
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int r = lane & 3, c = lane >> 2;

  double acc[HX][HY][2];
  bool okx[HX], oky[HY];
#pragma unroll
  for (int hx = 0; hx < HX; ++hx) {
    okx[hx] = hx * 8 + c < k;
#pragma unroll
    for (int hy = 0; hy < HY; ++hy) acc[hx][hy][0] = acc[hx][hy][1] = 0.0;
  }
#pragma unroll
  for (int hy = 0; hy < HY; ++hy) oky[hy] = hy * 8 + c < l;

  constexpr int64_t TILE = (int64_t)KB_BG_WARPS * KB_BG_STEPS * 4;  // 128 rows per CTA pass
  const int64_t ntiles = (n + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TILE + (int64_t)warp * (KB_BG_STEPS * 4) + r;
    double a[KB_BG_STEPS][HX], b[KB_BG_STEPS][HY];
#pragma unroll
    for (int s = 0; s < KB_BG_STEPS; ++s) {
      const int64_t row = base + 4 * s;
      const bool ok = row < n;
#pragma unroll
      for (int hx = 0; hx < HX; ++hx)
        a[s][hx] = (ok && okx[hx]) ? X[row * ldx + hx * 8 + c] : 0.0;
#pragma unroll
      for (int hy = 0; hy < HY; ++hy)
        b[s][hy] = (ok && oky[hy]) ? Y[row * ldy + hy * 8 + c] : 0.0;
    }
#pragma unroll
    for (int s = 0; s < KB_BG_STEPS; ++s)
#pragma unroll
      for (int hx = 0; hx < HX; ++hx)
#pragma unroll
        for (int hy = 0; hy < HY; ++hy)
          kb_dmma(acc[hx][hy][0], acc[hx][hy][1], a[s][hx], b[s][hy]);
  }

  // warp partials -> shared, entry (i, j) at i * NJ + j
#pragma unroll
  for (int hx = 0; hx < HX; ++hx)
#pragma unroll
    for (int hy = 0; hy < HY; ++hy) {
      const int e = (hx * 8 + c) * NJ + hy * 8 + 2 * r;
      sm[warp * KL + e] = acc[hx][hy][0];
      sm[warp * KL + e + 1] = acc[hx][hy][1];
    }
  __syncthreads();
  if (t < KL) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < KB_BG_WARPS; ++w) tot += sm[w * KL + t];
    rd.partials[(size_t)blockIdx.x * KL + t] = tot;
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int prev = atomicAdd(rd.ticket, 1u);
    s_last_b = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last_b) return;
  __threadfence();
  // finishing block: Q threads share an entry, each adds every Q-th block partial in block
  // order, then the Q sub-sums are added in fixed order
  constexpr int Q = (KB_BG_WARPS * 32) / KL;
  {
    const int e = t % KL, q = t / KL;
    double a2 = 0.0;
    for (unsigned int bb = q; bb < gridDim.x; bb += Q)
      a2 += __ldcg(&rd.partials[(size_t)bb * KL + e]);
    sm[t] = a2;
  }
  __syncthreads();
  double fin = 0.0;
  if (t < KL) {
#pragma unroll
    for (int i = 0; i < Q; ++i) fin += sm[i * KL + t];
  }
  if (rd.collective && rd.cm.size > 1) {  // row-partitioned operands: sum over ranks, same launch
    fin = kb_p2p_allreduce(fin, KL, rd.cm);
    if (t < KL && *rd.cm.error) fin = nan("");
  }
  if (t < KL) {
    const int i = t / NJ, j = t % NJ;
    if (i < k && j < l) {
      if (flags & 1) fin = sqrt(fabs(fin));
      G[(size_t)i * ldg + j] = fin;
      if (Gacc != nullptr) Gacc[(size_t)i * ldacc + j] += fin;
    }
  }
  if (t == 0) *rd.ticket = 0u;
}

// ----------------------------------------------------------------- apply --
// MODE 0: Z = X C      MODE 1: Z = Y - X C      MODE 2: Z = Y + X C
// X is n x k (ldx), C is k x l (ldc, device memory), Y and Z are n x l.  A = X tile (8 rows x 4
// columns), B = C tile held in registers for the whole kernel (negated for MODE 1, which is
// exact), the accumulator starts from Y.  Z may alias Y, and Z may alias X (in-place X <- X C):
// a warp has loaded all operands of its 8 rows before the (warp-synchronous) MMA, stores follow it.
template <int HK, int HL, int MODE>
__global__ void __launch_bounds__(KB_BG_WARPS * 32)
kb_block_apply_kernel(int64_t n, int k, int l, const double* X, int64_t ldx,
                      const double* __restrict__ C, int64_t ldc, const double* Y, int64_t ldy,
                      double* Z, int64_t ldz, KbRed rd) {
  if (kb_gated(rd)) return;
  constexpr int KS = HK * 2;  // 4-wide contraction steps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & 3, g = lane >> 2;

  double bf[KS][HL];
#pragma unroll
  for (int s = 0; s < KS; ++s)
#pragma unroll
    for (int hl = 0; hl < HL; ++hl) {
      const int kk = s * 4 + q, j = hl * 8 + g;
      double v = (kk < k && j < l) ? C[(size_t)kk * ldc + j] : 0.0;
      bf[s][hl] = (MODE == 1) ? -v : v;
    }

  constexpr int64_t TILE = (int64_t)KB_BG_WARPS * KB_BA_STEPS * 8;  // 128 rows per CTA pass
  const int64_t ntiles = (n + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * TILE + (int64_t)warp * (KB_BA_STEPS * 8) + g;
    double a[KB_BA_STEPS][KS], d[KB_BA_STEPS][HL][2];
#pragma unroll
    for (int st = 0; st < KB_BA_STEPS; ++st) {
      const int64_t row = base + 8 * st;
      const bool ok = row < n;
#pragma unroll
      for (int s = 0; s < KS; ++s)
        a[st][s] = (ok && s * 4 + q < k) ? X[row * ldx + s * 4 + q] : 0.0;
#pragma unroll
      for (int hl = 0; hl < HL; ++hl)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = hl * 8 + 2 * q + e;
          d[st][hl][e] = (MODE != 0 && ok && col < l) ? Y[row * ldy + col] : 0.0;
        }
    }
#pragma unroll
    for (int st = 0; st < KB_BA_STEPS; ++st)
#pragma unroll
      for (int hl = 0; hl < HL; ++hl)
#pragma unroll
        for (int s = 0; s < KS; ++s) kb_dmma(d[st][hl][0], d[st][hl][1], a[st][s], bf[s][hl]);
#pragma unroll
    for (int st = 0; st < KB_BA_STEPS; ++st) {
      const int64_t row = base + 8 * st;
      if (row < n) {
#pragma unroll
        for (int hl = 0; hl < HL; ++hl)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = hl * 8 + 2 * q + e;
            if (col < l) Z[row * ldz + col] = d[st][hl][e];
          }
      }
    }
  }
}
