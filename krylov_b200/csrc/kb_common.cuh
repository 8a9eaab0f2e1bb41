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
#define KB_BAR_CTAS 256      // most CTAs of the persistent CG kernel
#define KB_BAR_BYTES (2 * KB_BAR_CTAS * 128)  // two sets of one 128-byte line per CTA

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

// ------------------------------------------------------------ peer comm --
// One-shot all-reduce over NVLink peer memory, fused into the last block of a
// reduction kernel ("custom one-shot P2P allreduce fused into the
// reduction kernel's last-block epilogue").  Every rank owns a mailbox
//   mbox[parity][src_rank][stride]   (k entries of 16 bytes: value + sequence flag, kb_ll_store)
// mapped into every peer with CUDA IPC.  A collective with sequence number q
// writes this rank's k partial sums into slot [q & 1][rank] of EVERY rank's
// mailbox (one self-validating 16-byte NVLink store per value and destination), then waits
// until all `size` entries of its own mailbox carry q and adds the contributions in rank
// order (deterministic, identical on all ranks).  q is a device-resident counter,
// so launches skipped by the gate do not consume a number and parities alternate:
// a rank can only overwrite slot [q & 1] after every peer has finished reading
// collective q - 2 (it needed their contribution to q - 1).
struct KbComm {
  double* const* peers;         // device array [size]: mailbox base of every rank (peer-mapped)
  unsigned long long* counter;  // this rank's collective sequence counter
  int* error;                   // set to 1 if a peer did not arrive within the spin budget
  int rank, size, stride;       // stride in doubles per (parity, src) slot
};

struct kb_comm_s {
  KbComm dev;
  double* mailbox;       // own mailbox (cudaMalloc, IPC-exported)
  double** peers_dev;    // device copy of the pointer table
  void* peer_base[64];   // opened IPC mappings (host view)
  int max_k;
  int opened;
};

// ------------------------------------------------------------ peer halo ---
// Halo exchange of row-partitioned products through NVLink peer memory, no NCCL
// kernel involved (a NCCL send/recv kernel cannot get an SM while the persistent
// SpMV fills the GPU, so the exchange used to serialise behind it).
// Every rank owns one IPC-exported allocation:
//   u64 flags[64] | u64 acks[64] | u64 counter | u64 done | int error | ... | data @ KB_HALO_DATA
// Product number q (device-resident counter, gating-safe):
//   push kernel   : waits until every destination acknowledged q-1, gathers the boundary
//                   rows of x straight into the destinations' data areas, fences, sets
//                   flags[my rank] = q there.
//   boundary kernel: waits for flags[src] >= q of its sources (q from its own consumed-
//                   products counter: the push may run on a side stream), reads its own data
//                   area, and when all its blocks are done sets acks[my rank] = q at the sources.
#define KB_HALO_DATA 2048
struct KbHalo {
  unsigned char* const* peers;  // device array [size]: allocation base of every rank
  int rank, size;
};

struct kb_halo_s {
  KbHalo dev;
  unsigned char* base;          // own allocation
  unsigned char** peers_dev;
  void* peer_base[64];
  size_t data_bytes;
  int opened;
};

__device__ __forceinline__ volatile unsigned long long* kb_halo_u64(unsigned char* base,
                                                                    size_t byte_off) {
  return reinterpret_cast<volatile unsigned long long*>(base + byte_off);
}
#define KB_HALO_FLAGS 0
#define KB_HALO_ACKS 512
#define KB_HALO_COUNTER 1024
#define KB_HALO_DONE 1032
#define KB_HALO_ERROR 1040
#define KB_HALO_RECV_COUNTER 1056  // u64: products whose halo this rank has consumed
#define KB_HALO_PUSH_TICKET 1048  // u32 arrival counter of the push kernel (it may run on a side
                                  // stream next to a reduction that uses the workspace ticket)

// spin until *p >= want (3 s budget, then raise the STICKY error word and go on: the callers
// test it, stop pushing and poison their results with NaN; hosts turn it into an exception at
// the next batch boundary, Comm.check_p2p / DistCsrMatrix.check_p2p)
__device__ __forceinline__ void kb_halo_wait(volatile unsigned long long* p,
                                             unsigned long long want, unsigned char* own) {
  const long long t0 = clock64();
  while (*p < want) {
    if (clock64() - t0 > 6000000000ll) {
      *reinterpret_cast<volatile int*>(own + KB_HALO_ERROR) = 1;
      break;
    }
  }
}

// ------------------------------------------------------------- workspace --
struct kb_ws_s {
  double* partials;      // KB_MAX_BLOCKS * max_k doubles
  unsigned int* ticket;  // arrival counter of the single-launch reductions
  double* barbuf;        // grid-barrier slots of the persistent CG kernel (KB_BAR_BYTES)
  int max_k;
  int num_sms;
  const int* gate;       // device int or nullptr
  int gate_tag;
  kb_comm_s* comm;       // optional: fused all-reduce of every finished reduction
  int collective;        // 1: reductions launched now end with the all-reduce
};

// Passed by value to every kernel.
struct KbRed {
  double* partials;
  unsigned int* ticket;
  const int* gate;
  int gate_tag;
  int collective;
  KbComm cm;
};

static inline KbRed kb_red(const kb_ws_s* ws) {
  KbRed r;
  r.partials = ws->partials;
  r.ticket = ws->ticket;
  r.gate = ws->gate;
  r.gate_tag = ws->gate_tag;
  r.collective = (ws->comm != nullptr && ws->collective) ? 1 : 0;
  if (ws->comm != nullptr) {
    r.cm = ws->comm->dev;
  } else {
    r.cm.peers = nullptr;
    r.cm.counter = nullptr;
    r.cm.error = nullptr;
    r.cm.rank = 0;
    r.cm.size = 1;
    r.cm.stride = 0;
  }
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
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-
// serialization attribute may be scheduled while its predecessor in the stream drains;
// griddepcontrol.wait blocks until that predecessor has completed and its memory is visible (a
// no-op for a normal launch), launch_dependents lets the successor's CTAs take SM slots as soon as
// every CTA of this grid has started.  Used at the very top of the short kernels of a solver step,
// so that launch latency overlaps the predecessor's tail instead of following it.
__device__ __forceinline__ void kb_pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

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

// One 16-byte store carries a double and its sequence flag twice: {lo, flag, hi, flag}.  Each
// 8-byte half validates itself (the scheme of NCCL's LL protocol), so data and "it has arrived"
// travel in ONE NVLink store -- no fence and no second round trip for a separate flag.
__device__ __forceinline__ void kb_ll_store(double* dst16, double v, unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst16), "r"(lo), "r"(flag),
               "r"(hi), "r"(flag)
               : "memory");
}
__device__ __forceinline__ bool kb_ll_load(const double* src16, unsigned flag, double* v) {
  unsigned a, b, c, d;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "l"(src16)
               : "memory");
  if (b != flag || d != flag) return false;
  *v = __hiloint2double((int)c, (int)a);
  return true;
}

// All-reduce of k values held by threads t < k of ONE block (see KbComm).  Must be
// called by every thread of that block.  Returns the global sum in threads t < k.
// Mailbox entry (parity, source rank, value t) = 16 bytes written by kb_ll_store; the flag is the
// low word of the collective's sequence number (it differs from what the entry held two
// collectives earlier).  The system fences make the all-reduce a release / acquire point: what
// the blocks of this grid stored to peer memory (and fenced) before arriving is visible to a
// peer once that peer's all-reduce returns.
__device__ __forceinline__ double kb_p2p_allreduce(double v, int k, const KbComm& cm) {
  __shared__ unsigned long long s_seq;
  const int t = threadIdx.x;
  if (t == 0) s_seq = ++(*cm.counter);
  __threadfence_system();
  __syncthreads();
  const unsigned long long q = s_seq;
  const unsigned flag = (unsigned)q;
  const size_t par = (size_t)(q & 1ull) * cm.size;
  double tot = 0.0;
  if (t < k) {
    const size_t slot = (par + cm.rank) * cm.stride + 2 * (size_t)t;
    for (int i = 1; i <= cm.size; ++i) {  // own mailbox last
      const int p = (cm.rank + i) % cm.size;
      kb_ll_store(cm.peers[p] + slot, v, flag);
    }
    const double* mine = cm.peers[cm.rank];
    const long long t0 = clock64();
    bool dead = false;
    for (int p = 0; p < cm.size; ++p) {  // rank order: identical bits on every rank
      const double* e = mine + (par + p) * cm.stride + 2 * (size_t)t;
      double x = 0.0;
      while (!dead && !kb_ll_load(e, flag, &x)) {
        if (clock64() - t0 > 6000000000ll) {  // ~3 s: a peer is gone; fail loudly, do not hang
          *cm.error = 1;
          dead = true;
        }
      }
      tot += x;
    }
    if (dead) tot = nan("");
  }
  __threadfence_system();
  __syncthreads();
  return tot;
}

// Finish a grid-wide column-wise reduction in the same launch: every block
// publishes its partial, the last block to arrive (ticket) adds the partials
// in block order and writes out[0..k).  `acc` follows the kb_block_colsum
// precondition.  All threads of all blocks must call this.
// Returns true (block-uniform) in the block that finished the reduction, i.e. after every
// block of the grid has arrived.
__device__ __forceinline__ bool kb_grid_colsum(double acc, int k, const KbRed& rd, double* out,
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
    if (accumulate && t < k) fin += out[t];
    // row-partitioned problems: sum over ranks through NVLink peer memory, same launch
    if (rd.collective && rd.cm.size > 1) {
      fin = kb_p2p_allreduce(fin, k, rd.cm);
      if (t < k && *rd.cm.error) fin = nan("");
    }
    if (t < k) out[t] = fin;
    if (t == 0) *rd.ticket = 0u;  // ready for the next launch on this stream
  }
  return s_last != 0;
}

// The same single-launch reduction for JC column-wise sums at once (tall-skinny V^T w):
// acc[jj] feeds out[jj * k + c], jj < cnt.  Needs cnt * k <= blockDim.x and <= the workspace /
// communicator max_k (checked by the host).  Fixed summation shape -> bitwise reproducible.
template <int JC>
__device__ __forceinline__ void kb_grid_multisum(const double (&acc)[JC], int cnt, int k,
                                                 const KbRed& rd, double* out, double* sm) {
  __shared__ int s_last_m;
  const int t = threadIdx.x;
  const int kk = cnt * k;
#pragma unroll
  for (int jj = 0; jj < JC; ++jj) {
    if (jj < cnt) {  // block-uniform
      const double tot = kb_block_colsum(acc[jj], k, sm);
      if (t < k) rd.partials[(size_t)blockIdx.x * kk + jj * k + t] = tot;
    }
  }
  __threadfence();
  __syncthreads();
  if (t == 0) {
    unsigned int prev = atomicAdd(rd.ticket, 1u);
    s_last_m = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last_m) {
    __threadfence();
    const int Q = blockDim.x / kk;
    double a2 = 0.0;
    if (t < Q * kk) {
      const int c = t % kk;
      for (unsigned int b = t / kk; b < gridDim.x; b += Q)
        a2 += __ldcg(&rd.partials[(size_t)b * kk + c]);
    }
    __syncthreads();
    sm[t] = a2;
    __syncthreads();
    double fin = 0.0;
    if (t < kk)
      for (int i = 0; i < Q; ++i) fin += sm[i * kk + t];
    if (rd.collective && rd.cm.size > 1) {
      fin = kb_p2p_allreduce(fin, kk, rd.cm);
      if (t < kk && *rd.cm.error) fin = nan("");
    }
    if (t < kk) out[t] = fin;
    if (t == 0) *rd.ticket = 0u;
  }
}
