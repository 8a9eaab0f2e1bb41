// Common host/device helpers for libkrylov_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/krylov_b200.h"

#define KB_BLOCK 256         // nominal threads per block of the vector kernels
#define KB_MAX_K 256         // widest block of right-hand sides
#define KB_CTAS_PER_SM 8     // resident CTAs per SM the vector grids are sized for
#define KB_MAX_BLOCKS 2048   // upper bound of any reduction grid (partials buffer)

// ---------------------------------------------------------------- errors --
extern thread_local char kb_errbuf[512];
int kb_fail(int code, const char* fmt, ...);

#define KB_CUDA(call)                                                             \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess)                                                        \
      return kb_fail(KB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                         \
  } while (0)

#define KB_REQUIRE(cond, msg)                                    \
  do {                                                           \
    if (!(cond)) return kb_fail(KB_EINVAL, "%s: %s", __func__, msg); \
  } while (0)

#define KB_LAUNCH_CHECK()                                                        \
  do {                                                                           \
    cudaError_t e_ = cudaPeekAtLastError();                                      \
    if (e_ != cudaSuccess)                                                       \
      return kb_fail(KB_ECUDA, "%s: launch failed: %s", __func__, cudaGetErrorString(e_)); \
  } while (0)

// ------------------------------------------------------------- workspace --
struct kb_ws_s {
  double* partials;      // KB_MAX_BLOCKS * max_k doubles
  unsigned int* ticket;  // arrival counter of the single-launch reductions
  int max_k;
  int num_sms;
  const int* gate;       // device int or nullptr
  int gate_tag;
};

// Passed by value to every kernel.
struct KbRed {
  double* partials;
  unsigned int* ticket;
  const int* gate;
  int gate_tag;
};

static inline KbRed kb_red(const kb_ws_s* ws) {
  KbRed r;
  r.partials = ws->partials;
  r.ticket = ws->ticket;
  r.gate = ws->gate;
  r.gate_tag = ws->gate_tag;
  return r;
}

// threads per block so that blockDim % k == 0 (each thread keeps one column)
static inline int kb_block_for(int k) { return (KB_BLOCK / k) * k; }

extern int g_vec_ctas;  // kb_tune key 2

static inline int kb_grid_for(const kb_ws_s* ws, int64_t total, int block, int per_thread) {
  int64_t need = (total + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
  int64_t cap = (int64_t)ws->num_sms * g_vec_ctas;
  if (cap > KB_MAX_BLOCKS) cap = KB_MAX_BLOCKS;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

// ----------------------------------------------------------- device side --
__device__ __forceinline__ bool kb_gated(const KbRed& rd) {
  // uniform across the grid: only single-block scalar kernels of *earlier*
  // launches ever write the gate word
  return rd.gate != nullptr && (*(volatile const int*)rd.gate) <= rd.gate_tag;
}

__device__ __forceinline__ double kb_nz(double d) { return d != 0.0 ? d : 1.0; }

// Rounded product followed by rounded sum/difference: what NumPy does with
// the temporaries of `y += a * x` (no FMA contraction), kept so that
// element-wise results equal the reference's bit for bit.
__device__ __forceinline__ double kb_mul_add(double a, double x, double y) {
  return __dadd_rn(y, __dmul_rn(a, x));
}
__device__ __forceinline__ double kb_mul_sub(double a, double x, double y) {
  return __dsub_rn(y, __dmul_rn(a, x));
}

__device__ __forceinline__ double kb_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column-wise sum over the block.  Precondition: blockDim.x % k == 0 and
// thread t contributes to column t % k.  Result valid in threads t < k
// (column t).  `sm` holds >= blockDim.x doubles.  Fixed summation shape ->
// bitwise reproducible.
__device__ __forceinline__ double kb_block_colsum(double acc, int k, double* sm) {
  const int t = threadIdx.x;
  double tot = 0.0;
  if (k == 1) {
    double v = kb_warp_sum(acc);
    __syncthreads();  // protect sm reuse across consecutive calls
    if ((t & 31) == 0) sm[t >> 5] = v;
    __syncthreads();
    if (t == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      for (int i = 0; i < nw; ++i) tot += sm[i];
    }
  } else {
    __syncthreads();
    sm[t] = acc;
    __syncthreads();
    if (t < k) {
      const int q = blockDim.x / k;
      for (int i = 0; i < q; ++i) tot += sm[i * k + t];
    }
  }
  return tot;
}

// Finish a grid-wide column-wise reduction in the same launch: every block
// publishes its partial, the last block to arrive (ticket) adds the partials
// in block order and writes out[0..k).  `acc` follows the kb_block_colsum
// precondition.  All threads of all blocks must call this.
__device__ __forceinline__ void kb_grid_colsum(double acc, int k, const KbRed& rd, double* out,
                                               double* sm, bool accumulate = false) {
  __shared__ int s_last;
  const int t = threadIdx.x;
  double tot = kb_block_colsum(acc, k, sm);
  if (t < k) rd.partials[(size_t)blockIdx.x * k + t] = tot;
  __threadfence();
  __syncthreads();
  if (t == 0) {
    unsigned int prev = atomicAdd(rd.ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int usable = (blockDim.x / k) * k;  // == blockDim.x for vector kernels
    double a2 = 0.0;
    if (t < usable) {
      const int c = t % k;
      const int q = t / k;
      const int Q = usable / k;
      for (unsigned int b = q; b < gridDim.x; b += Q) a2 += __ldcg(&rd.partials[(size_t)b * k + c]);
    }
    // kb_block_colsum needs blockDim % k == 0 on the participating range;
    // threads >= usable carry 0 and (for k > 1) are never read.
    double fin;
    if (k == 1) {
      fin = kb_block_colsum(a2, 1, sm);
    } else {
      __syncthreads();
      sm[t] = a2;
      __syncthreads();
      fin = 0.0;
      if (t < k) {
        const int Q = usable / k;
        for (int i = 0; i < Q; ++i) fin += sm[i * k + t];
      }
    }
    // accumulate: add to what an earlier launch on this stream left in out[]
    if (t < k) out[t] = accumulate ? out[t] + fin : fin;
    if (t == 0) *rd.ticket = 0u;  // ready for the next launch on this stream
  }
}
