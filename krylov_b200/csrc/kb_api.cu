// libkrylov_b200.so -- extern "C" entry points (see include/krylov_b200.h).
// Host code only validates arguments, picks the schedule and launches; no
// entry point synchronises the device or allocates per call.
#include <stdarg.h>

#include <new>

#include "kb_common.cuh"
#include "kb_scalar.cuh"
#include "kb_spmv.cuh"
#include "kb_vec.cuh"
#include "kb_march.cuh"
#include "kb_lines.cuh"
#include "kb_block.cuh"

thread_local char kb_errbuf[512] = {0};

int kb_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(kb_errbuf, sizeof(kb_errbuf), fmt, ap);
  va_end(ap);
  return code;
}



static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

// Launch with the programmatic-stream-serialization attribute (see kb_pdl_prologue): only for
// kernels that start with kb_pdl_prologue().
template <typename... KArgs, typename... Args>
static inline void kb_launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// runtime tunables (kb_tune): stream-kernel configuration and grid sizing
static int g_stream_cfg = 0;
static int g_stream_ctas = 0;  // 0 = configuration default
static int g_pattern_ctas = 0; // kb_tune key 3 (0 = default)
static int g_window_cfg = 0;   // kb_tune key 4
static int g_stencil_cfg = 0;   // kb_tune key 10: tile shape of the constant-diagonal kernel
static int g_stencil_l2pol = 0; // kb_tune key 12: L2 hint of the x windows (0 evict_last, 1 none, 2 evict_first)
static int g_march_ch = 0;     // kb_tune key 13: planes per work item of the marching kernel (0 auto)
static int g_march_cfg = 0;    // kb_tune key 14: tile / ring shape of the marching kernel
static int g_cg_fuse = 1;      // kb_tune key 15: fused marching CG kernels in kb_cg_run (0 off)
static int g_part_dbg = 0;     // kb_tune key 21 (measurement only, results become wrong): bit 0 no
                               // peer pushes of r, bit 1 no all-reduce in the partitioned CG kernels
static int g_cg_cfg[2] = {-1, -1};  // kb_tune keys 22 / 23: shape of the fused CG kernels KIND 1 / 2 (-1 auto)
static int g_pdl = 1;           // kb_tune key 29: programmatic dependent launch of the short kernels
                                // of a MINRES step and of the marching kernels (0 off)
static int g_march_depth = 96;  // kb_tune key 24: deepest grid marched top to bottom by one CTA per column
static int g_march_even = 0;   // kb_tune key 20: marching grids sized for equal items per CTA (measured
                               // no gain at 512^3 on one GPU: profiles/r2_march_even.txt)
static int g_spmm_lines = 1;    // kb_tune key 16: line-marching SpMM (k > 1, constant 3-D stencils):
                                // 0 off, 1 where a line fills >= half of its chunks, 2 wherever valid
static int g_lines_ch = 0;      // kb_tune key 17: lines per work item of it (0 = 32)
static int g_lines_order = 1;   // kb_tune key 19: work-item order, 0 natural, 1 planes fastest
static int g_lines_cfg = 0;      // kb_tune key 18: 0 auto (1024-entry chunks, 2 CTAs/SM; 512 where the values are streamed and k <= 16), 1 = 512, 4 CTAs/SM, 2 = 1024
static int g_stencil_ctas = 0;  // kb_tune key 11: CTAs/SM cap of it (0 = occupancy limit)
static int g_cgs_jc = 8;  // kb_tune key 9: basis vectors per multi-dot launch (8 or 16)
static int g_rowwise_contig = -1;  // kb_tune key 5: -1 auto, 0 strided, 1 contiguous rows per block
static int g_rowwise_ctas = 0;     // kb_tune key 6: CTAs/SM of the contiguous row-wise grid
static int g_tile_block = 0;       // kb_tune key 8: 0 natural tile order (default: the blocked
                                   // order cut DRAM traffic 11.80 -> 11.57 GB at 512^3 but ran
                                   // 2.2 instead of 1.73 ms, profiles/r1_tile_order.txt),
                                   // -1 auto-blocked for large planes, n > 0: blocks of <= n tiles
static int g_spmm_cfg = -1;        // kb_tune key 7: -1 = row-wise kernel (default: measured equal
                                   // or faster, profiles/r1_configs.txt), 0 = windowed RPT 2, 1 = RPT 4
int g_vec_ctas = KB_CTAS_PER_SM;

// Finds the set of distinct diagonals (col - row); if there are at most 16 and every
// row lists its entries in ascending column order, builds the per-row masks.
static int kb_detect_pattern(kb_csr_s* h, cudaStream_t st) {
  int* d_tab = nullptr;  // 64 hash slots + overflow + fail
  KB_CUDA(cudaMalloc(&d_tab, sizeof(int) * 66));
  int init[66];
  for (int i = 0; i < 64; ++i) init[i] = KB_PAT_EMPTY;
  init[64] = init[65] = 0;
  cudaError_t e = cudaMemcpyAsync(d_tab, init, sizeof(init), cudaMemcpyHostToDevice, st);
  int grid = (int)((h->n_rows + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  int res[66];
  if (e == cudaSuccess) {
    kb_pattern_collect_kernel<<<grid, 256, 0, st>>>(h->n_rows, h->rowptr, h->colidx, d_tab,
                                                    d_tab + 64);
    e = cudaMemcpyAsync(res, d_tab, sizeof(res), cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    cudaFree(d_tab);
    return kb_fail(KB_ECUDA, "kb_detect_pattern: %s", cudaGetErrorString(e));
  }
  int offs[64], nd = 0;
  for (int i = 0; i < 64; ++i)
    if (res[i] != KB_PAT_EMPTY) offs[nd++] = res[i];
  if (res[64] != 0 || nd == 0 || nd > 16) {
    cudaFree(d_tab);
    return KB_OK;  // not stencil-like: keep the CSR schedules
  }
  for (int i = 1; i < nd; ++i)  // insertion sort, ascending
    for (int j = i; j > 0 && offs[j - 1] > offs[j]; --j) {
      int t = offs[j];
      offs[j] = offs[j - 1];
      offs[j - 1] = t;
    }
  h->pat.nd = nd;
  for (int i = 0; i < 16; ++i) h->pat.off[i] = i < nd ? offs[i] : 0;
  // group nearby diagonals into contiguous x windows (windowed kernel)
  h->pat.nw = 0;
  for (int i = 0; i < 16; ++i) h->pat.grp[i] = h->pat.dwlo[i] = 0;
  for (int i = 0; i < 8; ++i) h->pat.wlo[i] = h->pat.wspan[i] = 0;
  if (nd <= 8) {
    int nw = 0;
    bool ok = true;
    for (int i = 0; i < nd; ++i) {
      if (nw > 0 && offs[i] - h->pat.wlo[nw - 1] <= KB_WIN_SLACK - 10) {
        h->pat.wspan[nw - 1] = offs[i] - h->pat.wlo[nw - 1];
      } else {
        if (nw == 8) {
          ok = false;
          break;
        }
        h->pat.wlo[nw] = offs[i];
        h->pat.wspan[nw] = 0;
        ++nw;
      }
      h->pat.grp[i] = nw - 1;
      h->pat.dwlo[i] = h->pat.wlo[nw - 1];
    }
    h->pat.nw = ok ? nw : 0;
  }
  e = cudaMalloc(&h->masks, sizeof(uint16_t) * (size_t)h->n_rows);
  if (e != cudaSuccess) {
    h->masks = nullptr;
    cudaFree(d_tab);
    cudaGetLastError();
    return KB_OK;  // no memory for the masks: keep the CSR schedules
  }
  kb_pattern_build_kernel<<<grid, 256, 0, st>>>(h->n_rows, h->rowptr, h->colidx, h->pat, h->masks,
                                                d_tab + 65);
  int fail = 1;
  e = cudaMemcpyAsync(&fail, d_tab + 65, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_tab);
  if (e != cudaSuccess) return kb_fail(KB_ECUDA, "kb_detect_pattern: %s", cudaGetErrorString(e));
  if (fail) {
    cudaFree(h->masks);
    h->masks = nullptr;
    return KB_OK;
  }
  h->pattern_ok = 1;
  return KB_OK;
}

// Constant-coefficient stencil?  Exact test: every stored value equals (64-bit compare) the
// representative of its diagonal.  Only for the windowed kernel's reach (<= 8 diagonals).
static int kb_detect_constdiag(kb_csr_s* h, cudaStream_t st) {
  h->constv = 0;
  if (!h->pattern_ok || h->pat.nd > 8 || h->pat.nw == 0 || h->nnz == 0) return KB_OK;
  unsigned long long* d_c = nullptr;  // 16 representatives + fail flag
  KB_CUDA(cudaMalloc(&d_c, sizeof(unsigned long long) * 17));
  cudaError_t e = cudaMemsetAsync(d_c, 0, sizeof(unsigned long long) * 17, st);
  int grid = (int)((h->n_rows + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  unsigned long long res[17];
  if (e == cudaSuccess) {
    kb_constdiag_fill_kernel<<<grid, 256, 0, st>>>(h->n_rows, h->rowptr, h->vals, h->masks, d_c);
    kb_constdiag_check_kernel<<<grid, 256, 0, st>>>(h->n_rows, h->rowptr, h->vals, h->masks, d_c,
                                                    reinterpret_cast<int*>(d_c + 16));
    e = cudaMemcpyAsync(res, d_c, sizeof(res), cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_c);
  if (e != cudaSuccess) return kb_fail(KB_ECUDA, "kb_detect_constdiag: %s", cudaGetErrorString(e));
  if ((int)(res[16] & 0xffffffffull) != 0) return KB_OK;
  for (int d = 0; d < 8; ++d) memcpy(&h->cv.c[d], &res[d], sizeof(double));
  h->constv = 1;
  return KB_OK;
}

extern "C" {

int kb_version(void) { return 100; }

int kb_last_error(char* buf, size_t len) {
  if (buf == nullptr || len == 0) return KB_EINVAL;
  strncpy(buf, kb_errbuf, len - 1);
  buf[len - 1] = 0;
  return KB_OK;
}

int kb_tune(int key, int value) {
  switch (key) {
    case 0: g_stream_cfg = value; return KB_OK;
    case 1: g_stream_ctas = value; return KB_OK;
    case 2: g_vec_ctas = value > 0 ? value : KB_CTAS_PER_SM; return KB_OK;
    case 3: g_pattern_ctas = value; return KB_OK;
    case 4: g_window_cfg = value; return KB_OK;  // -1: gather variant of the pattern kernel
    case 5: g_rowwise_contig = value; return KB_OK;
    case 6: g_rowwise_ctas = value; return KB_OK;
    case 7: g_spmm_cfg = value; return KB_OK;
    case 8: g_tile_block = value; return KB_OK;
    case 9: g_cgs_jc = value; return KB_OK;
    case 10: g_stencil_cfg = value; return KB_OK;
    case 11: g_stencil_ctas = value; return KB_OK;
    case 12: g_stencil_l2pol = value; return KB_OK;
    case 13: g_march_ch = value; return KB_OK;
    case 14: g_march_cfg = value; return KB_OK;
    case 15: g_cg_fuse = value; return KB_OK;
    case 16: g_spmm_lines = value; return KB_OK;
    case 17: g_lines_ch = value; return KB_OK;
    case 18: g_lines_cfg = value; return KB_OK;
    case 19: g_lines_order = value; return KB_OK;
    case 20: g_march_even = value; return KB_OK;
    case 21: g_part_dbg = value; return KB_OK;
    case 22: g_cg_cfg[0] = value; return KB_OK;
    case 23: g_cg_cfg[1] = value; return KB_OK;
    case 24: g_march_depth = value; return KB_OK;
    case 25: g_merge_cfg = value; return KB_OK;  // tile shape of the merge kernel
    case 26: g_merge_ctas = value; return KB_OK; // its CTAs per SM (0 = all that fit)
    case 27: g_merge_order = value; return KB_OK; // tile -> CTA order (kb_merge.cuh)
    case 28: g_small_n = value; return KB_OK;     // largest n of the persistent CG kernel (0 off)
    case 29: g_pdl = value; return KB_OK;
    default: return kb_fail(KB_EINVAL, "kb_tune: unknown key %d", key);
  }
}

int kb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  KB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return KB_OK;
}

// ------------------------------------------------------------ workspace --
int kb_ws_create(kb_ws_t* out, int max_k) {
  KB_REQUIRE(out != nullptr, "null handle pointer");
  KB_REQUIRE(max_k >= 1 && max_k <= KB_MAX_K, "max_k out of range [1, 256]");
  int dev = 0, sms = 0, major = 0;
  KB_CUDA(cudaGetDevice(&dev));
  KB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  KB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10)
    return kb_fail(KB_EUNSUPPORTED, "libkrylov_b200 is built for sm_100a only (device is sm_%d*)",
                   major * 10);
  kb_ws_s* ws = new (std::nothrow) kb_ws_s();
  if (!ws) return kb_fail(KB_ENOMEM, "out of host memory");
  ws->max_k = max_k;
  ws->num_sms = sms;
  ws->gate = nullptr;
  ws->gate_tag = 0;
  ws->comm = nullptr;
  ws->collective = 0;
  cudaError_t e = cudaMalloc(&ws->partials, sizeof(double) * KB_MAX_BLOCKS * (size_t)max_k);
  // ticket[0]: arrival counter of the reductions; ticket[2]: barrier counter of kb_cg_small
  if (e == cudaSuccess) e = cudaMalloc(&ws->ticket, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(ws->ticket, 0, 4 * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMalloc(&ws->barbuf, KB_BAR_BYTES);
  if (e != cudaSuccess) {
    delete ws;
    return kb_fail(KB_ECUDA, "workspace allocation failed: %s", cudaGetErrorString(e));
  }
  *out = ws;
  return KB_OK;
}

int kb_ws_destroy(kb_ws_t ws) {
  if (!ws) return KB_OK;
  cudaFree(ws->partials);
  cudaFree(ws->ticket);
  cudaFree(ws->barbuf);
  delete ws;
  return KB_OK;
}

int kb_ws_set_gate(kb_ws_t ws, const int* stop_at, int tag) {
  KB_REQUIRE(ws != nullptr, "null workspace");
  ws->gate = stop_at;
  ws->gate_tag = tag;
  return KB_OK;
}

// ------------------------------------------------------------ peer comm --
int kb_comm_create(kb_comm_t* out, int rank, int size, int max_k) {
  KB_REQUIRE(out != nullptr, "null handle pointer");
  KB_REQUIRE(size >= 1 && size <= 64 && rank >= 0 && rank < size, "bad rank/size");
  KB_REQUIRE(max_k >= 1 && max_k <= KB_MAX_K, "max_k out of range [1, 256]");
  kb_comm_s* c = new (std::nothrow) kb_comm_s();
  if (!c) return kb_fail(KB_ENOMEM, "out of host memory");
  memset(c, 0, sizeof(*c));
  c->max_k = max_k;
  c->dev.rank = rank;
  c->dev.size = size;
  c->dev.stride = ((2 * max_k + 15) / 16) * 16;  // 16 bytes per value; 128-byte multiples
  // mailbox | counter | error live in one IPC-exported allocation
  const size_t mbox_doubles = (size_t)2 * size * c->dev.stride;
  const size_t bytes = (mbox_doubles + 16) * sizeof(double);
  cudaError_t e = cudaMalloc(&c->mailbox, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&c->peers_dev, sizeof(double*) * size);
  if (e != cudaSuccess) {
    if (c->mailbox) cudaFree(c->mailbox);
    delete c;
    return kb_fail(KB_ECUDA, "kb_comm_create: %s", cudaGetErrorString(e));
  }
  c->dev.counter = reinterpret_cast<unsigned long long*>(c->mailbox + mbox_doubles);
  c->dev.error = reinterpret_cast<int*>(c->mailbox + mbox_doubles + 8);
  c->dev.peers = c->peers_dev;
  *out = c;
  return KB_OK;
}

int kb_comm_get_handle(kb_comm_t c, void* out64) {
  KB_REQUIRE(c != nullptr && out64 != nullptr, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  KB_CUDA(cudaIpcGetMemHandle(&h, c->mailbox));
  memcpy(out64, &h, 64);
  return KB_OK;
}

int kb_comm_open(kb_comm_t c, const void* handles) {
  KB_REQUIRE(c != nullptr && handles != nullptr, "null argument");
  KB_REQUIRE(!c->opened, "already opened");
  double* table[64];
  for (int p = 0; p < c->dev.size; ++p) {
    if (p == c->dev.rank) {
      table[p] = c->mailbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * (size_t)p, 64);
    void* base = nullptr;
    KB_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[p] = base;
    table[p] = (double*)base;
  }
  KB_CUDA(cudaMemcpy(c->peers_dev, table, sizeof(double*) * c->dev.size, cudaMemcpyHostToDevice));
  c->opened = 1;
  return KB_OK;
}

int kb_comm_destroy(kb_comm_t c) {
  if (!c) return KB_OK;
  for (int p = 0; p < c->dev.size; ++p)
    if (c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  if (c->peers_dev) cudaFree(c->peers_dev);
  if (c->mailbox) cudaFree(c->mailbox);
  delete c;
  return KB_OK;
}

int kb_comm_error(kb_comm_t c, int* err) {
  KB_REQUIRE(c != nullptr && err != nullptr, "null argument");
  KB_CUDA(cudaMemcpy(err, c->dev.error, sizeof(int), cudaMemcpyDeviceToHost));
  return KB_OK;
}

int kb_ws_set_comm(kb_ws_t ws, kb_comm_t c, int collective) {
  KB_REQUIRE(ws != nullptr, "null workspace");
  KB_REQUIRE(c == nullptr || c->opened || c->dev.size == 1, "communicator not opened");
  KB_REQUIRE(c == nullptr || c->max_k >= ws->max_k, "communicator max_k too small");
  ws->comm = c;
  ws->collective = collective ? 1 : 0;
  return KB_OK;
}

int kb_allreduce(kb_ws_t ws, int k, double* slot, void* stream) {
  KB_REQUIRE(ws != nullptr && slot != nullptr, "null argument");
  KB_REQUIRE(ws->comm != nullptr, "workspace has no communicator");
  KB_REQUIRE(k >= 1 && k <= ws->max_k, "k exceeds workspace max_k");
  KbRed rd = kb_red(ws);
  const int block = ((k > ws->comm->dev.size ? k : ws->comm->dev.size) + 31) / 32 * 32;
  kb_allreduce_kernel<<<1, block, 0, S(stream)>>>(k, slot, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// ------------------------------------------------------------------ CSR --
static void kb_csr_pick(kb_csr_s* h) {
  // Row-length statistics -> schedule (SURVEY.md 7 "short rows").  The stream
  // kernel gives one thread one row: right for stencil-like matrices whose
  // rows are short and even.  Long or very skewed rows go to the merge kernel
  // (tiles of equal nonzero count, several lanes per row); k > 1 and unpadded
  // arrays to the row-wise kernel.
  const double mean = h->n_rows > 0 ? (double)h->nnz / (double)h->n_rows : 0.0;
  const bool tma_ok = h->padded && h->n_rows >= 1 && h->n_rows < (1ll << 31) - 512;
  int sched = 1;
  if (tma_ok && h->nnz > 0) sched = 5;
  if (tma_ok && mean <= 32.0 && h->max_row_len <= 8 * (mean + 8.0)) sched = 2;
  if (sched == 2 && h->pattern_ok) sched = h->constv ? 4 : 3;
  if (h->forced == 1) sched = 1;
  if (h->forced == 2 && h->padded) sched = 2;
  if (h->forced == 5 && tma_ok && h->nnz > 0) sched = 5;
  if (h->forced == 3 && h->pattern_ok) sched = 3;
  if (h->forced == 4 && h->constv) sched = 4;
  h->schedule = sched;
}

int kb_csr_create(kb_csr_t* out, int64_t n_rows, int64_t n_cols, int64_t nnz,
                  const int32_t* rowptr, const int32_t* colidx, const double* vals, int padded,
                  void* stream) {
  KB_REQUIRE(out != nullptr, "null handle pointer");
  KB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "negative size");
  KB_REQUIRE(nnz < (1ll << 31), "nnz must fit int32 row pointers");
  KB_REQUIRE(rowptr != nullptr, "null rowptr");
  KB_REQUIRE(nnz == 0 || (colidx != nullptr && vals != nullptr), "null colidx/vals");
  if (padded)
    KB_REQUIRE(((uintptr_t)colidx % 16 == 0) && ((uintptr_t)vals % 16 == 0),
               "padded CSR arrays must be 16-byte aligned");
  kb_csr_s* h = new (std::nothrow) kb_csr_s();
  if (!h) return kb_fail(KB_ENOMEM, "out of host memory");
  h->n_rows = n_rows;
  h->n_cols = n_cols;
  h->nnz = nnz;
  h->rowptr = rowptr;
  h->colidx = colidx;
  h->vals = vals;
  h->padded = padded;
  h->forced = 0;
  h->max_row_len = 0;
  if (n_rows > 0) {
    // one pass over the row pointers (longest row + structure) and one over the column indices:
    // a malformed matrix is refused here instead of being read out of bounds by the kernels
    int* d_max = nullptr;
    int host[2] = {0, 0};
    cudaError_t e = cudaMalloc(&d_max, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_max, 0, 2 * sizeof(int), S(stream));
    if (e == cudaSuccess) {
      int grid = (int)((n_rows + 255) / 256);
      if (grid > 1184) grid = 1184;
      kb_max_row_len_kernel<<<grid, 256, 0, S(stream)>>>(n_rows, nnz, rowptr, d_max);
      if (nnz > 0) {
        int g2 = (int)((nnz + 255) / 256 > 2368 ? 2368 : (nnz + 255) / 256);
        kb_colidx_check_kernel<<<g2, 256, 0, S(stream)>>>(nnz, n_cols, colidx, d_max);
      }
      e = cudaMemcpyAsync(host, d_max, 2 * sizeof(int), cudaMemcpyDeviceToHost, S(stream));
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(S(stream));  // creation is not on the hot path
    if (d_max) cudaFree(d_max);
    if (e != cudaSuccess) {
      delete h;
      return kb_fail(KB_ECUDA, "kb_csr_create: %s", cudaGetErrorString(e));
    }
    h->max_row_len = host[0];
    if (host[1] != 0) {
      delete h;
      return kb_fail(KB_EINVAL, "kb_csr_create: malformed CSR arrays (%s%s%s%s)",
                     (host[1] & 1) ? "rowptr[0] != 0; " : "",
                     (host[1] & 2) ? "row pointers decrease; " : "",
                     (host[1] & 4) ? "rowptr[n_rows] != nnz; " : "",
                     (host[1] & 8) ? "column index outside [0, n_cols)" : "");
    }
  }
  h->masks = nullptr;
  h->pattern_ok = 0;
  h->constv = 0;
  h->pat.nd = 0;
  h->merge_meta = nullptr;
  h->carry = nullptr;
  h->merge_fix = nullptr;
  h->n_fix = 0;
  h->merge_T = 0;
  h->n_mtiles = 0;
  if (padded && n_rows > 0 && n_rows < (1ll << 31) - 1024 && n_cols < (1ll << 31) &&
      h->max_row_len >= 1 && h->max_row_len <= 16) {
    int rc = kb_detect_pattern(h, S(stream));
    if (rc == KB_OK) rc = kb_detect_constdiag(h, S(stream));
    if (rc != KB_OK) {
      if (h->masks) cudaFree(h->masks);
      delete h;
      return rc;
    }
  }
  kb_csr_pick(h);
  *out = h;
  return KB_OK;
}


int kb_csr_destroy(kb_csr_t h) {
  if (h && h->masks) cudaFree(h->masks);
  if (h) kb_merge_release(h);
  delete h;
  return KB_OK;
}

int kb_csr_set_schedule(kb_csr_t h, int schedule) {
  KB_REQUIRE(h != nullptr, "null matrix");
  KB_REQUIRE(schedule >= 0 && schedule <= 5, "schedule must be 0 ... 5");
  if (schedule == 5 && !(h->padded && h->nnz > 0 && h->n_rows < (1ll << 31) - 512))
    return kb_fail(KB_EUNSUPPORTED, "merge schedule needs padded, 16-byte aligned CSR arrays");
  if (schedule == 4 && !h->constv)
    return kb_fail(KB_EUNSUPPORTED,
                   "stencil schedule needs <= 8 diagonals with one constant value each");
  if (schedule == 2 && !h->padded)
    return kb_fail(KB_EUNSUPPORTED, "stream schedule needs padded, 16-byte aligned CSR arrays");
  if (schedule == 3 && !h->pattern_ok)
    return kb_fail(KB_EUNSUPPORTED,
                   "pattern schedule needs <= 16 distinct diagonals and ascending columns");
  h->forced = schedule;
  kb_csr_pick(h);
  return KB_OK;
}

int kb_csr_get_info(kb_csr_t h, int64_t* n_rows, int64_t* n_cols, int64_t* nnz, int* max_row_len,
                    int* schedule) {
  KB_REQUIRE(h != nullptr, "null matrix");
  if (n_rows) *n_rows = h->n_rows;
  if (n_cols) *n_cols = h->n_cols;
  if (nnz) *nnz = h->nnz;
  if (max_row_len) *max_row_len = h->max_row_len;
  if (schedule) *schedule = h->schedule;
  return KB_OK;
}

int kb_csr_get_stencil(kb_csr_t h, int* nd, int* offsets16, double* coeffs8, int* constv,
                       const uint16_t** masks) {
  KB_REQUIRE(h != nullptr, "null matrix");
  if (nd) *nd = h->pattern_ok ? h->pat.nd : 0;
  if (offsets16)
    for (int i = 0; i < 16; ++i) offsets16[i] = h->pattern_ok ? h->pat.off[i] : 0;
  if (coeffs8)
    for (int i = 0; i < 8; ++i) coeffs8[i] = h->constv ? h->cv.c[i] : 0.0;
  if (constv) *constv = h->constv;
  if (masks) *masks = h->pattern_ok ? h->masks : nullptr;
  return KB_OK;
}

// ----------------------------------------------------------------- SpMV --
}  // extern "C"

template <int ROWS, int STAGES, int CAP, int DOT>
static int kb_launch_stream_cfg(kb_csr_s* A, kb_ws_s* ws, int ctas_default, const double* x,
                                double* y, int mode, const double* z, const double* coef,
                                const double* w, double* out, cudaStream_t st) {
  typedef KbStreamSmem<STAGES, CAP> Smem;
  static bool configured[64] = {false};  // per device (function attributes are per device)
  auto kern = kb_spmv_stream_kernel<ROWS, STAGES, CAP, DOT>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(Smem)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int n_tiles = (int)((A->n_rows + ROWS - 1) / ROWS);
  int grid = ws->num_sms * (g_stream_ctas > 0 ? g_stream_ctas : ctas_default);
  if (grid > n_tiles) grid = n_tiles;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  kern<<<grid, ROWS + 32, sizeof(Smem), st>>>((int)A->n_rows, n_tiles, A->rowptr, A->colidx,
                                              A->vals, x, y, mode, z, coef, w, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

template <int DOT>
static int kb_launch_stream(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                            const double* z, const double* coef, const double* w, double* out,
                            cudaStream_t st) {
  // default: 512-row tiles, 2 stages x 4096 nnz (96 KB smem), 2 CTAs/SM = 1088 threads/SM:
  // best of the sweep in profiles/r1_spmv_sweep.txt (occupancy, not stage depth, was the lever)
  switch (g_stream_cfg) {
    case 1: return kb_launch_stream_cfg<256, 2, 2048, DOT>(A, ws, 4, x, y, mode, z, coef, w, out, st);
    case 2: return kb_launch_stream_cfg<256, 3, 2048, DOT>(A, ws, 3, x, y, mode, z, coef, w, out, st);
    case 3: return kb_launch_stream_cfg<128, 3, 1024, DOT>(A, ws, 4, x, y, mode, z, coef, w, out, st);
    case 4: return kb_launch_stream_cfg<256, 4, 2048, DOT>(A, ws, 2, x, y, mode, z, coef, w, out, st);
    case 5: return kb_launch_stream_cfg<128, 2, 1024, DOT>(A, ws, 8, x, y, mode, z, coef, w, out, st);
    default: return kb_launch_stream_cfg<512, 2, 4096, DOT>(A, ws, 2, x, y, mode, z, coef, w, out, st);
  }
}

template <int MAXD, int MINCTAS, int DOT>
static int kb_launch_pattern_cfg(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                                 const double* z, const double* coef, const double* w, double* out,
                                 cudaStream_t st) {
  constexpr int ROWS = 512, STAGES = 2, CAP = 4096;
  typedef KbPatternSmem<STAGES, CAP> Smem;
  static bool configured[64] = {false};
  auto kern = kb_spmv_pattern_kernel<ROWS, STAGES, CAP, MAXD, MINCTAS, DOT>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(Smem)));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int n_tiles = (int)((A->n_rows + ROWS - 1) / ROWS);
  int grid = ws->num_sms * (g_pattern_ctas > 0 ? g_pattern_ctas : MINCTAS);
  if (grid > n_tiles) grid = n_tiles;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  kern<<<grid, ROWS + 32, sizeof(Smem), st>>>((int)A->n_rows, n_tiles, A->rowptr, A->masks, A->vals,
                                              A->pat, x, y, mode, z, coef, w, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

template <int ROWS, int STAGES, int MINB, int DOT, bool CONSTV>
static int kb_launch_window_cfg(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                                const double* z, const double* coef, const double* w, double* out,
                                cudaStream_t st) {
  static int max_smem[64] = {0};  // per device: opt-in dynamic shared memory limit once set
  auto kern = kb_spmv_window_kernel<ROWS, STAGES, MINB, DOT, CONSTV>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  // stage buffers sized from the actual pattern
  int span = 0;
  for (int g = 0; g < A->pat.nw; ++g) span = A->pat.wspan[g] > span ? A->pat.wspan[g] : span;
  const int cap = CONSTV ? 0 : ((ROWS * A->pat.nd + 8 + 3) & ~3);
  const int wlen = (ROWS + span + 6 + 1) & ~1;
  const size_t smem = ((size_t)STAGES * cap + (size_t)STAGES * A->pat.nw * wlen) * 8 + 16 * STAGES;
  if (dev < 0 || dev >= 64 || max_smem[dev] == 0) {
    int lim = 0;
    KB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lim - 8 * 1024));
    if (dev >= 0 && dev < 64) max_smem[dev] = lim - 8 * 1024;
  }
  int ctas = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, ROWS + 32, smem));
  if (ctas < 1) return kb_fail(KB_EUNSUPPORTED, "windowed SpMV: stage buffers do not fit");
  if (CONSTV && g_stencil_ctas > 0 && g_stencil_ctas < ctas) ctas = g_stencil_ctas;
  if (g_pattern_ctas > 0 && g_pattern_ctas < ctas) ctas = g_pattern_ctas;
  const int n_tiles = (int)((A->n_rows + ROWS - 1) / ROWS);
  int grid = ws->num_sms * ctas;
  if (grid > n_tiles) grid = n_tiles;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  // cache-blocked visiting order (see KbTileOrder): only when a plane of matrix stream is
  // too large for its x windows to survive in L2 until the next plane needs them
  KbTileOrder ord = {0, 0, 0};
  const int64_t plane = A->pat.off[A->pat.nd - 1];  // largest offset
  const int64_t row_bytes = 8 * A->pat.nd + 14;
  if (g_tile_block != 0 && plane > 0 && plane % ROWS == 0 && A->n_rows % plane == 0 &&
      (g_tile_block > 0 || plane * row_bytes > (12ll << 20))) {
    const int tpp = (int)(plane / ROWS);
    int tpb = 0;
    for (int d = 1; d <= tpp; ++d)
      if (tpp % d == 0 && (g_tile_block > 0 ? d <= g_tile_block
                                            : (int64_t)d * ROWS * row_bytes <= (5ll << 19)))
        tpb = d;
    if (tpb > 0 && tpb < tpp) {
      ord.tpp = tpp;
      ord.tpb = tpb;
      ord.nplanes = n_tiles / tpp;
    }
  }
  kern<<<grid, ROWS + 32, smem, st>>>((int)A->n_rows, (int)A->n_cols, n_tiles, cap, wlen, ord, A->rowptr,
                                      A->masks, A->vals, A->pat, A->cv, x, y, mode, z, coef, w, out,
                                      kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// windowed kernel applicable to this product?
static inline bool kb_window_ok(const kb_csr_s* A, const double* x) {
  return A->pat.nw > 0 && (A->n_cols % 2 == 0) && ((uintptr_t)x % 16 == 0);
}

template <int DOT>
static int kb_launch_window(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                            const double* z, const double* coef, const double* w, double* out,
                            cudaStream_t st) {
  // measured (profiles/r1_window_sweep.txt): 512-row tiles win by ~4 % on the 512^3 matrix,
  // 256-row tiles (4 CTAs/SM) at 256^3 and below
  if (g_window_cfg == 0 && A->n_rows > (40ll << 20))
    return kb_launch_window_cfg<512, 2, 2, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
  switch (g_window_cfg) {
    case 1: return kb_launch_window_cfg<256, 3, 3, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
    case 2: return kb_launch_window_cfg<512, 2, 2, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
    case 3: return kb_launch_window_cfg<128, 2, 7, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
    case 4: return kb_launch_window_cfg<128, 3, 6, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
    case 5: return kb_launch_window_cfg<256, 2, 5, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
    default: return kb_launch_window_cfg<256, 2, 4, DOT, false>(A, ws, x, y, mode, z, coef, w, out, st);
  }
}

// constant diagonals: x windows only, no matrix stream (kb_spmv_stencil_kernel)
template <int RPT, int STAGES, int MINB, int DOT>
static int kb_launch_stencil_cfg(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                                 const double* z, const double* coef, const double* w,
                                 double* out, cudaStream_t st) {
  static int max_smem[64] = {0};
  auto kern = kb_spmv_stencil_kernel<RPT, STAGES, MINB, DOT>;
  constexpr int TR = 256 * RPT;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  int span = 0;
  for (int g = 0; g < A->pat.nw; ++g) span = A->pat.wspan[g] > span ? A->pat.wspan[g] : span;
  const int wlen = (TR + span + 6 + 1) & ~1;
  const size_t smem = (size_t)STAGES * A->pat.nw * wlen * 8 + 16 * STAGES;
  if (dev < 0 || dev >= 64 || max_smem[dev] == 0) {
    int lim = 0;
    KB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lim - 8 * 1024));
    if (dev >= 0 && dev < 64) max_smem[dev] = lim - 8 * 1024;
  }
  int ctas = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 288, smem));
  if (ctas < 1) return kb_fail(KB_EUNSUPPORTED, "stencil SpMV: stage buffers do not fit");
  if (g_stencil_ctas > 0 && g_stencil_ctas < ctas) ctas = g_stencil_ctas;
  const int n_tiles = (int)((A->n_rows + TR - 1) / TR);
  int grid = ws->num_sms * ctas;
  if (grid > n_tiles) grid = n_tiles;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  int d_center = -1, center_base = 0;  // <w, y> with w == x: the operand is in the centre window
  if (DOT == 1 && w == x)
    for (int d = 0; d < A->pat.nd; ++d)
      if (A->pat.off[d] == 0) {
        d_center = d;
        center_base = A->pat.grp[d] * wlen - A->pat.dwlo[d] + (A->pat.dwlo[d] & 1);
      }
  kern<<<grid, 288, smem, st>>>((int)A->n_rows, (int)A->n_cols, n_tiles, wlen, A->masks, A->pat,
                                A->cv, x, y, mode, z, coef, w, d_center, center_base,
                                g_stencil_l2pol, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// second version (kb_spmv_stencil2_kernel): the number of diagonals is a template parameter
template <int ND, int RPT, int STAGES, int MINB, int DOT, bool WX>
static int kb_launch_stencil2_wx(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                                 const double* z, const double* coef, const double* w,
                                 double* out, cudaStream_t st) {
  static int max_smem[64] = {0};
  auto kern = kb_spmv_stencil2_kernel<ND, RPT, STAGES, MINB, DOT, WX>;
  constexpr int TR = 256 * RPT;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  int span = 0;
  for (int g = 0; g < A->pat.nw; ++g) span = A->pat.wspan[g] > span ? A->pat.wspan[g] : span;
  const int wlen = (TR + span + 6 + 1) & ~1;
  const size_t smem = (size_t)STAGES * A->pat.nw * wlen * 8 + 16 * STAGES;
  if (dev < 0 || dev >= 64 || max_smem[dev] == 0) {
    int lim = 0;
    KB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lim - 8 * 1024));
    if (dev >= 0 && dev < 64) max_smem[dev] = lim - 8 * 1024;
  }
  int ctas = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 288, smem));
  if (ctas < 1) return kb_fail(KB_EUNSUPPORTED, "stencil SpMV: stage buffers do not fit");
  if (g_stencil_ctas > 0 && g_stencil_ctas < ctas) ctas = g_stencil_ctas;
  const int n_tiles = (int)((A->n_rows + TR - 1) / TR);
  int grid = ws->num_sms * ctas;
  if (grid > n_tiles) grid = n_tiles;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  kern<<<grid, 288, smem, st>>>((int)A->n_rows, (int)A->n_cols, n_tiles, wlen, A->masks, A->pat,
                                A->cv, x, y, mode, z, coef, w, g_stencil_l2pol, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

template <int ND, int RPT, int STAGES, int MINB, int DOT>
static int kb_launch_stencil2_cfg(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                                  const double* z, const double* coef, const double* w,
                                  double* out, cudaStream_t st) {
  // <w, y> with w == x and a main diagonal in the middle of the pattern: the dot operand is the
  // value the row sum loads anyway
  if (DOT == 1 && w == x && (ND & 1) && A->pat.off[ND / 2] == 0)
    return kb_launch_stencil2_wx<ND, RPT, STAGES, MINB, DOT, (DOT == 1)>(A, ws, x, y, mode, z, coef,
                                                                         w, out, st);
  return kb_launch_stencil2_wx<ND, RPT, STAGES, MINB, DOT, false>(A, ws, x, y, mode, z, coef, w,
                                                                  out, st);
}

template <int ND, int DOT>
static int kb_launch_stencil2(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                              const double* z, const double* coef, const double* w, double* out,
                              cudaStream_t st) {
  switch (g_stencil_cfg) {
    case 1: return kb_launch_stencil2_cfg<ND, 2, 3, 4, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 2: return kb_launch_stencil2_cfg<ND, 1, 3, 6, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 3: return kb_launch_stencil2_cfg<ND, 4, 2, 2, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 4: return kb_launch_stencil2_cfg<ND, 2, 2, 5, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 5: return kb_launch_stencil2_cfg<ND, 4, 2, 3, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    default: return kb_launch_stencil2_cfg<ND, 2, 2, 4, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
  }
}

// ------------------------------------------------ plane-marching stencil kernel ---
// Geometry of kb_stencil_march_kernel for this matrix and tile height; false if the pattern is
// not {-P, inner diagonals within +-1024, +P} with constant coefficients.
static bool kb_march_geom(const kb_csr_s* A, int RPT, KbMarch* g, int CT = 256, int slots = 0) {
  if (!A->constv || A->pat.nd != 7 || A->masks == nullptr) return false;
  const int* off = A->pat.off;
  const int P = off[6];
  if (P <= 0 || off[0] != -P || (P & 1) || A->n_rows % P != 0) return false;
  int L = 1;
  for (int d = 1; d < 6; ++d) {
    const int a = off[d] < 0 ? -off[d] : off[d];
    L = a > L ? a : L;
  }
  const int LP = ((L + CT - 1) / CT) * CT;
  const int TR = CT * RPT;
  if (LP > 1024 || L >= P || P < TR) return false;
  if ((A->n_cols & 1) || A->n_rows >= (1ll << 31) - 4096 || A->n_cols >= (1ll << 31) - 4096)
    return false;
  g->P = P;
  g->LP = LP;
  g->wlen = TR + 2 * LP;
  g->ncol = (P + TR - 1) / TR;
  g->nplanes = (int)((A->n_rows + P - 1) / P);
  // planes per work item: 32 / 16 measured best on deep grids (profiles/r1_march_bench_512.txt);
  // a thin slab with at least one column tile per SM (the per-rank share of a row-partitioned
  // problem) is marched top to bottom by one CTA per column: no quantisation of the work over
  // the grid, all columns in lock-step (profiles/r2_slab_tune.txt: 0.75 -> 0.80 of peak at 64 planes)
  int ch = g_march_ch > 0 ? g_march_ch : (g->nplanes >= 256 ? 32 : 16);
  if (g_march_ch == 0 && g->nplanes <= g_march_depth &&
      (slots > 0 ? (g->ncol <= slots && g->ncol * 100 >= slots * 85) : g->ncol >= 148))
    ch = g->nplanes;
  if (ch > g->nplanes) ch = g->nplanes;
  // Grids below 512 planes: the work items (ncol x ceil(nplanes / ch)) are dealt to `slots` CTAs in
  // rounds, and a last round that fills a fraction of the machine costs a whole march -- pick the
  // ch that minimises rounds x (ch + 2 planes per march).  256^3: 32 planes per item = 512 items =
  // 2 rounds on 444 slots, 83.9 us; 8 planes = 2048 items = 5 rounds, 67.2 us; 128^3: 4 planes,
  // 14.0 instead of 24.9 us (profiles/r2_march_ch.txt).  Only where the default gives fewer than four
  // rounds: with many rounds the quantisation is small and the tuned 32 / 16 stay (2-GPU slabs of
  // 512^3: 1263 it/s with 32 planes per item, 1237 with the model's 64).
  const long long sl = slots > 0 ? slots : 444;
  if (g_march_ch == 0 && ch < g->nplanes && g->nplanes < 512 &&
      (long long)g->ncol * ((g->nplanes + ch - 1) / ch) < 4 * sl) {  // fewer than 4 rounds
    long long best = -1;
    for (int c = 2; c <= 64 && c <= g->nplanes; c *= 2) {
      const long long items = (long long)g->ncol * ((g->nplanes + c - 1) / c);
      const long long cost = ((items + sl - 1) / sl) * (c + 2);
      if (best < 0 || cost <= best) {  // ties: the longer march (fewer redundant planes)
        best = cost;
        ch = c;
      }
    }
  }
  g->ch = ch;
  g->nitems = g->ncol * ((g->nplanes + ch - 1) / ch);
  return true;
}

template <int RPT, int NS, int MINB, int KIND, int DOT, bool WX, bool PART = false, int CT = 256>
static int kb_launch_march_t(kb_csr_s* A, kb_ws_s* ws, const KbMarch& g, const double* x, double* y,
                             int mode, const double* z, const double* coef, const double* w,
                             const KbMarchCg& cg, double* out, cudaStream_t st) {
  static int max_smem[64] = {0};
  auto kern = kb_stencil_march_kernel<7, RPT, NS, MINB, KIND, DOT, WX, PART, CT>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  const size_t smem = (size_t)(NS + (KIND == 1 ? 2 : 0)) * g.wlen * 8 + (2 * NS + 2) * 8;
  if (dev < 0 || dev >= 64 || max_smem[dev] == 0) {
    int lim = 0;
    KB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lim - 8 * 1024));
    if (dev >= 0 && dev < 64) max_smem[dev] = lim - 8 * 1024;
  }
  int ctas = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, CT + 32, smem));
  if (ctas < 1) return kb_fail(KB_EUNSUPPORTED, "marching stencil kernel: ring does not fit");
  if (g_stencil_ctas > 0 && g_stencil_ctas < ctas) ctas = g_stencil_ctas;
  int grid = ws->num_sms * ctas;
  if (grid > g.nitems) grid = g.nitems;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  if (g_march_even && grid > 0) {
    // the work items cost the same: the smallest grid that needs the same number of rounds ends
    // with every CTA busy instead of a last round that fills a fraction of the machine
    const int rounds = (g.nitems + grid - 1) / grid;
    grid = (g.nitems + rounds - 1) / rounds;
  }
  kb_launch_pdl(g_pdl != 0, kern, dim3(grid), dim3(CT + 32), smem, st, (int)A->n_rows,
                (int)A->n_cols, g, (const uint16_t*)A->masks, A->pat, A->cv, x, y, mode, z, coef, w, cg,
                g_stencil_l2pol, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// kb_tune 14: 0 = 1024-row tiles, ring of 4; 1 = 1024 rows, ring of 5; 2 = 512 rows, ring of 4;
// 3 = 512 rows, ring of 5; (fused CG kernels only) 4 = 896-row tiles (224 consumer threads x 4),
// 5 = 768-row tiles (256 x 3)
static int kb_march_rpt() { return (g_march_cfg == 2 || g_march_cfg == 3) ? 2 : 4; }

// Tile shapes of the fused CG kernels: {rows per thread, consumer threads, resident CTAs per SM}
struct KbShape { int cfg, rpt, ct, ctas1, ctas2; };  // ctas1 / ctas2: KIND 1 / KIND 2
static const KbShape kb_shapes[] = {{0, 4, 256, 2, 3}, {1, 4, 256, 2, 2}, {2, 2, 256, 3, 4},
                                    {3, 2, 256, 3, 3}, {4, 4, 224, 2, 3}, {5, 3, 256, 2, 3}};
static const KbShape& kb_shape_of(int cfg) {
  for (const KbShape& sh : kb_shapes)
    if (sh.cfg == cfg) return sh;
  return kb_shapes[0];
}

template <int KIND, int DOT, bool WX>
static int kb_launch_march(kb_csr_s* A, kb_ws_s* ws, const KbMarch& g, const double* x, double* y,
                           int mode, const double* z, const double* coef, const double* w,
                           const KbMarchCg& cg, double* out, cudaStream_t st) {
  // the fused p update keeps two more windows (r staging): one CTA per SM less
  switch (g_march_cfg) {
    case 1: return kb_launch_march_t<4, 5, 2, KIND, DOT, WX>(A, ws, g, x, y, mode, z, coef, w, cg, out, st);
    case 2: return kb_launch_march_t<2, 4, (KIND == 1 ? 3 : 4), KIND, DOT, WX>(A, ws, g, x, y, mode, z, coef, w, cg, out, st);
    case 3: return kb_launch_march_t<2, 5, 3, KIND, DOT, WX>(A, ws, g, x, y, mode, z, coef, w, cg, out, st);
    default: return kb_launch_march_t<4, 4, (KIND == 1 ? 2 : 3), KIND, DOT, WX>(A, ws, g, x, y, mode, z, coef, w, cg, out, st);
  }
}

// The fused CG kernels (KIND 1 / 2) in any tile shape
template <int KIND, int DOT, bool PART>
static int kb_launch_march_cg(int cfg, kb_csr_s* A, kb_ws_s* ws, const KbMarch& g, const double* x,
                              const KbMarchCg& cg, double* out, cudaStream_t st) {
#define KB_MCG(RPT, NS, MINB, CT)                                                                  \
  return kb_launch_march_t<RPT, NS, MINB, KIND, DOT, false, PART, CT>(A, ws, g, x, nullptr, 0,     \
                                                                      nullptr, nullptr, nullptr,  \
                                                                      cg, out, st)
  switch (cfg) {
    case 1: KB_MCG(4, 5, 2, 256);
    case 2: KB_MCG(2, 4, (KIND == 1 ? 3 : 4), 256);
    case 3: KB_MCG(2, 5, 3, 256);
    case 4: KB_MCG(4, 4, (KIND == 1 ? 2 : 3), 224);
    case 5: KB_MCG(3, 4, (KIND == 1 ? 2 : 3), 256);
    default: KB_MCG(4, 4, (KIND == 1 ? 2 : 3), 256);
  }
#undef KB_MCG
}

// Tile shape + geometry of a fused CG kernel.  kb_tune 14 != 0 forces one shape; otherwise the
// default 1024-row tiles unless another shape needs >= 5 % fewer tile-row-steps per CTA: the
// per-CTA critical path is what bounds these kernels on thin slabs (profiles/r2_slab_tune.txt),
// so a shape whose column count fills the machine's CTA slots in ONE round wins there
// (512 x 512 planes: 293 tiles of 896 rows on 296 slots, 342 of 768 on 444).
static bool kb_cg_pick_shape(const kb_csr_s* A, int num_sms, int kind, int* cfg, KbMarch* g) {
  const int forced =
      g_cg_cfg[kind - 1] >= 0 ? g_cg_cfg[kind - 1] : (g_march_cfg != 0 ? g_march_cfg : -1);
  if (forced >= 0) {
    const KbShape& sh = kb_shape_of(forced);
    *cfg = sh.cfg;
    return kb_march_geom(A, sh.rpt, g, sh.ct, num_sms * (kind == 1 ? sh.ctas1 : sh.ctas2));
  }
  static const int cand1[] = {0, 4}, cand2[] = {0, 5, 4};
  const int* cand = kind == 1 ? cand1 : cand2;
  const int nc = kind == 1 ? 2 : 3;
  double cost0 = 0.0, best = 0.0;
  bool found = false;
  for (int i = 0; i < nc; ++i) {
    const KbShape& sh = kb_shape_of(cand[i]);
    const int slots = num_sms * (kind == 1 ? sh.ctas1 : sh.ctas2);
    KbMarch gg;
    if (!kb_march_geom(A, sh.rpt, &gg, sh.ct, slots)) continue;
    const int rounds = (gg.nitems + slots - 1) / slots;
    const double cost = (double)rounds * (gg.ch + 2) * sh.rpt * sh.ct;
    if (!found) {  // the default shape (or the first that applies)
      cost0 = best = cost;
      *cfg = sh.cfg;
      *g = gg;
      found = true;
    } else if (cost < 0.95 * cost0 && cost < best) {
      best = cost;
      *cfg = sh.cfg;
      *g = gg;
    }
  }
  return found;
}

// A x (+ epilogue, dot) by the marching kernel; KB_EUNSUPPORTED if the matrix does not qualify
template <int DOT>
static int kb_launch_march_spmv(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                                const double* z, const double* coef, const double* w, double* out,
                                cudaStream_t st) {
  KbMarch g;
  if (!kb_march_geom(A, kb_march_rpt(), &g)) return KB_EUNSUPPORTED;
  KbMarchCg cg;
  memset(&cg, 0, sizeof(cg));
  cg.rec.step = -1;
  if (DOT == 1 && w == x && A->pat.off[3] == 0)
    return kb_launch_march<0, DOT, (DOT == 1)>(A, ws, g, x, y, mode, z, coef, w, cg, out, st);
  return kb_launch_march<0, DOT, false>(A, ws, g, x, y, mode, z, coef, w, cg, out, st);
}

// kb_tune key 10: 0 (default) / 10 the plane-marching kernel where the matrix qualifies (3-D, plane
// >= one tile), else the tiled second version; 1-5 and 11 tiled second version (5 or 7 diagonals;
// 11 = its default shape); 6 the generic windowed kernel with constant values (the first
// implementation); 7-9 kb_spmv_stencil_kernel (any <= 8 diagonals)
template <int DOT>
static int kb_launch_stencil(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                             const double* z, const double* coef, const double* w, double* out,
                             cudaStream_t st) {
  if (g_stencil_cfg == 0 || g_stencil_cfg == 10) {  // plane-marching kernel (3-D stencils); else fall through
    const int rc = kb_launch_march_spmv<DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    if (rc != KB_EUNSUPPORTED) return rc;
  }
  if ((g_stencil_cfg <= 5 || g_stencil_cfg >= 10) && A->pat.nd == 7)
    return kb_launch_stencil2<7, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
  if ((g_stencil_cfg <= 5 || g_stencil_cfg >= 10) && A->pat.nd == 5)
    return kb_launch_stencil2<5, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
  switch (g_stencil_cfg) {
    case 6: return kb_launch_window_cfg<256, 3, 4, DOT, true>(A, ws, x, y, mode, z, coef, w, out, st);
    case 8: return kb_launch_stencil_cfg<2, 3, 3, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 9: return kb_launch_stencil_cfg<1, 3, 6, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    default: return kb_launch_stencil_cfg<2, 2, 4, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
  }
}

// Geometry of kb_spmm_lines_kernel; false if the matrix / block width / operand do not qualify:
// constant diagonals {-P, -n, -i, 0, +i, +n, +P}, k a power of two in [2, 32] (a thread keeps one
// column; 288 % k == 0 for the block reduction), x 16-byte aligned (TMA).
// var: the matrix has the pattern but not constant diagonals -- its values are streamed (k >= 8:
// below that the value area would dominate the ring slot)
static bool kb_lines_geom(const kb_csr_s* A, int k, const double* x, KbLines* g, bool* var = nullptr) {
  if (g_spmm_lines <= 0 || !A->pattern_ok || A->pat.nd != 7 || A->masks == nullptr) return false;
  const bool v = !A->constv || A->schedule == 3;
  // chunk size: 1024 entries; with streamed values the value area pushes four 1024-entry slots
  // past half an SM's shared memory for k <= 16 (one CTA per SM: 2.6 ms at 256^3, k = 16, against
  // 1.87 ms with 512-entry chunks, profiles/r2_spmm_var.txt)
  const int TR = (g_lines_cfg == 1 || (g_lines_cfg == 0 && v && k <= 16)) ? 512 : 1024;
  g->TR = TR;
  if (v && (var == nullptr || k < 8 || !A->padded)) return false;
  if (var) *var = v;
  if (k < 2 || k > 32 || (k & (k - 1)) != 0) return false;
  if (((uintptr_t)x & 15u) != 0) return false;
  const int* off = A->pat.off;
  if (off[3] != 0) return false;
  for (int d = 0; d < 3; ++d)
    if (off[d] != -off[6 - d]) return false;
  if (!(off[4] > 0 && off[5] > off[4] && off[6] > off[5])) return false;
  const long long inner = (long long)off[4] * k;
  if (inner > 64) return false;
  int ks = 0;
  while ((1 << ks) < k) ++ks;
  g->k = k;
  g->kshift = ks;
  g->inner = (int)inner;
  g->H = (int)((inner + 1) & ~1ll);
  g->L = (long long)off[5] * k;
  g->Pz = (long long)off[6] * k;
  g->N = (long long)A->n_rows * k;
  g->Nx = (long long)A->n_cols * k;
  g->nlines = (g->N + g->L - 1) / g->L;
  g->ncol = (int)((g->L + TR - 1) / TR);
  if (g_spmm_lines == 1 && 2 * g->L < (long long)g->ncol * TR) return false;  // chunks mostly empty
  g->voff = 3 * TR + 2 * g->H;
  g->vcap = v ? (TR / k) * 7 + 2 : 0;
  g->n_rows = A->n_rows;
  g->slotlen = g->voff + g->vcap;
  long long ch = g_lines_ch > 0 ? g_lines_ch : 32;
  if (ch > g->nlines) ch = g->nlines;
  if (ch < 1) ch = 1;
  g->ch = (int)ch;
  g->nitems = (long long)g->ncol * ((g->nlines + ch - 1) / ch);
  if (g->L + TR >= (1ll << 31)) return false;  // in-line positions are 32-bit
  g->tail = (int)(g->N - (g->nlines - 1) * g->L);
  g->lpp = 0;
  g->nplanes = 1;
  if (g_lines_order == 1 && g->Pz % g->L == 0) {
    g->lpp = g->Pz / g->L;
    g->nplanes = (g->nlines + g->lpp - 1) / g->lpp;
    const long long gpp = (g->lpp + ch - 1) / ch;
    g->nitems = gpp * g->ncol * g->nplanes;
  }
  return true;
}

template <int RPT, int MINB, int DOT, bool WX, bool VAR = false>
static int kb_launch_lines_t(kb_csr_s* A, kb_ws_s* ws, const KbLines& g, const double* x, double* y,
                             int mode, const double* z, const double* coef, const double* w,
                             double* out, cudaStream_t st) {
  static int max_smem[64] = {0};
  constexpr int NS = 4;
  auto kern = kb_spmm_lines_kernel<RPT, NS, MINB, DOT, WX, VAR>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  const size_t smem = (size_t)NS * g.slotlen * 8 + 2 * NS * 8;
  if (dev < 0 || dev >= 64 || max_smem[dev] == 0) {
    int lim = 0;
    KB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lim - 8 * 1024));
    if (dev >= 0 && dev < 64) max_smem[dev] = lim - 8 * 1024;
  }
  int ctas = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 288, smem));
  if (ctas < 1) return KB_EUNSUPPORTED;
  long long grid = (long long)ws->num_sms * ctas;
  if (grid > g.nitems) grid = g.nitems;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  kern<<<(int)grid, 288, smem, st>>>(g, A->masks, A->cv, A->rowptr, A->vals, x, y, mode, z, coef,
                                     w, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

template <int DOT>
static int kb_launch_lines(kb_csr_s* A, kb_ws_s* ws, const KbLines& g, bool var, const double* x,
                           double* y, int mode, const double* z, const double* coef,
                           const double* w, double* out, cudaStream_t st) {
  if (var) {  // variable coefficients: values streamed through the ring
    if (g.TR == 512)
      return kb_launch_lines_t<2, 4, DOT, false, true>(A, ws, g, x, y, mode, z, coef, w, out, st);
    return kb_launch_lines_t<4, 2, DOT, false, true>(A, ws, g, x, y, mode, z, coef, w, out, st);
  }
  const bool wx = DOT == 1 && w == x;  // <x, A x>: the operand is the middle diagonal's entry, on chip
  if (g.TR == 512) {
    if (wx) return kb_launch_lines_t<2, 4, DOT, (DOT == 1)>(A, ws, g, x, y, mode, z, coef, w, out, st);
    return kb_launch_lines_t<2, 4, DOT, false>(A, ws, g, x, y, mode, z, coef, w, out, st);
  }
  if (wx) return kb_launch_lines_t<4, 2, DOT, (DOT == 1)>(A, ws, g, x, y, mode, z, coef, w, out, st);
  return kb_launch_lines_t<4, 2, DOT, false>(A, ws, g, x, y, mode, z, coef, w, out, st);
}

template <int STAGES, int RPT, int MINB, int DOT>
static int kb_launch_spmm_window_cfg(kb_csr_s* A, kb_ws_s* ws, int k, const double* x, double* y,
                                     int mode, const double* z, const double* coef,
                                     const double* w, double* out, cudaStream_t st) {
  static int max_smem[64] = {0};
  auto kern = kb_spmm_window_kernel<STAGES, RPT, MINB, DOT>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  int span = 0;
  for (int g = 0; g < A->pat.nw; ++g) span = A->pat.wspan[g] > span ? A->pat.wspan[g] : span;
  const int rows_t = RPT * (256 / k);
  const int cap = (rows_t * A->pat.nd + 8 + 3) & ~3;
  const int wlen = rows_t + span;  // x rows per window
  const size_t smem =
      ((size_t)STAGES * cap + (size_t)STAGES * A->pat.nw * wlen * k) * 8 + 16 * STAGES;
  if (dev < 0 || dev >= 64 || max_smem[dev] == 0) {
    int lim = 0;
    KB_CUDA(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 lim - 8 * 1024));
    if (dev >= 0 && dev < 64) max_smem[dev] = lim - 8 * 1024;
  }
  if (smem > (size_t)max_smem[dev < 64 && dev >= 0 ? dev : 0]) return KB_EUNSUPPORTED;
  int ctas = 0;
  KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, 288, smem));
  if (ctas < 1) return KB_EUNSUPPORTED;
  const int n_tiles = (int)((A->n_rows + rows_t - 1) / rows_t);
  int grid = ws->num_sms * ctas;
  if (grid > n_tiles) grid = n_tiles;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  kern<<<grid, 288, smem, st>>>((int)A->n_rows, (int)A->n_cols, n_tiles, k, cap, wlen, A->rowptr,
                                A->masks, A->vals, A->pat, x, y, mode, z, coef, w, out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

template <int DOT>
static int kb_launch_spmm_window(kb_csr_s* A, kb_ws_s* ws, int k, const double* x, double* y,
                                 int mode, const double* z, const double* coef, const double* w,
                                 double* out, cudaStream_t st) {
  if (g_spmm_cfg == 1)
    return kb_launch_spmm_window_cfg<2, 4, 2, DOT>(A, ws, k, x, y, mode, z, coef, w, out, st);
  return kb_launch_spmm_window_cfg<2, 2, 4, DOT>(A, ws, k, x, y, mode, z, coef, w, out, st);
}

template <int DOT>
static int kb_launch_pattern(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                             const double* z, const double* coef, const double* w, double* out,
                             cudaStream_t st) {
  if (A->pat.nd <= 8) return kb_launch_pattern_cfg<8, 2, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
  return kb_launch_pattern_cfg<16, 1, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
}

extern "C" {

int kb_spmv(kb_csr_t A, kb_ws_t ws, int k, const double* x, double* y, int mode, const double* z,
            const double* coef, int dot, const double* w, double* out, void* stream) {
  KB_REQUIRE(A != nullptr && ws != nullptr, "null handle");
  KB_REQUIRE(k >= 1 && k <= ws->max_k, "k exceeds workspace max_k");
  KB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  KB_REQUIRE(dot >= 0 && dot <= 2, "dot must be 0, 1 or 2");
  KB_REQUIRE(y != nullptr && (x != nullptr || A->n_cols == 0), "null vector");
  KB_REQUIRE(mode == 0 || z != nullptr, "mode 1/2 need z");
  KB_REQUIRE(mode != 1 || coef != nullptr, "mode 1 needs coef");
  KB_REQUIRE(dot == 0 || out != nullptr, "dot needs out");
  KB_REQUIRE(dot != 1 || w != nullptr, "dot 1 needs w");
  cudaStream_t st = S(stream);
  if (A->n_rows == 0) {
    if (dot) KB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * k, st));
    return KB_OK;
  }
  if (k == 1 && A->schedule == 4 && kb_window_ok(A, x)) {
    if (dot == 0) return kb_launch_stencil<0>(A, ws, x, y, mode, z, coef, w, out, st);
    if (dot == 1) return kb_launch_stencil<1>(A, ws, x, y, mode, z, coef, w, out, st);
    return kb_launch_stencil<2>(A, ws, x, y, mode, z, coef, w, out, st);
  }
  if (k == 1 && (A->schedule == 3 || A->schedule == 4) && g_window_cfg >= 0 && kb_window_ok(A, x)) {
    if (dot == 0) return kb_launch_window<0>(A, ws, x, y, mode, z, coef, w, out, st);
    if (dot == 1) return kb_launch_window<1>(A, ws, x, y, mode, z, coef, w, out, st);
    return kb_launch_window<2>(A, ws, x, y, mode, z, coef, w, out, st);
  }
  if (k == 1 && (A->schedule == 3 || A->schedule == 4)) {
    if (dot == 0) return kb_launch_pattern<0>(A, ws, x, y, mode, z, coef, w, out, st);
    if (dot == 1) return kb_launch_pattern<1>(A, ws, x, y, mode, z, coef, w, out, st);
    return kb_launch_pattern<2>(A, ws, x, y, mode, z, coef, w, out, st);
  }
  if (k == 1 && A->schedule == 5) {
    return kb_launch_merge(A, ws, dot, x, y, mode, z, coef, w, out, st);
  }
  if (k == 1 && A->schedule >= 2) {
    if (dot == 0) return kb_launch_stream<0>(A, ws, x, y, mode, z, coef, w, out, st);
    if (dot == 1) return kb_launch_stream<1>(A, ws, x, y, mode, z, coef, w, out, st);
    return kb_launch_stream<2>(A, ws, x, y, mode, z, coef, w, out, st);
  }
  // blocked right-hand sides on a constant-coefficient 3-D stencil: line-marching SpMM
  if (k > 1 && (A->schedule == 4 || A->schedule == 3) && A->forced != 1) {
    KbLines g;
    bool var = false;
    if (kb_lines_geom(A, k, x, &g, &var)) {
      int rc;
      if (dot == 0) rc = kb_launch_lines<0>(A, ws, g, var, x, y, mode, z, coef, w, out, st);
      else if (dot == 1) rc = kb_launch_lines<1>(A, ws, g, var, x, y, mode, z, coef, w, out, st);
      else rc = kb_launch_lines<2>(A, ws, g, var, x, y, mode, z, coef, w, out, st);
      if (rc != KB_EUNSUPPORTED) return rc;
    }
  }
  // blocked right-hand sides on a stencil-like matrix: windowed SpMM
  if (k > 1 && 256 % k == 0 && A->pattern_ok && A->pat.nw > 0 && g_spmm_cfg >= 0 &&
      A->forced != 1 && ((uintptr_t)x % 16 == 0)) {
    int rc;
    if (dot == 0) rc = kb_launch_spmm_window<0>(A, ws, k, x, y, mode, z, coef, w, out, st);
    else if (dot == 1) rc = kb_launch_spmm_window<1>(A, ws, k, x, y, mode, z, coef, w, out, st);
    else rc = kb_launch_spmm_window<2>(A, ws, k, x, y, mode, z, coef, w, out, st);
    if (rc != KB_EUNSUPPORTED) return rc;  // else: stage buffers do not fit -> row-wise kernel
  }
  const int block = kb_block_for(k);
  int grid = kb_grid_for(ws, A->n_rows * (int64_t)k, block, 1);
  // optional (kb_tune 5): contiguous row slices per block
  const int contiguous = g_rowwise_contig > 0 ? 1 : 0;  // measured slower (profiles/r1_c4_tune.txt)
  if (contiguous) {
    const int ctas = g_rowwise_ctas > 0 ? g_rowwise_ctas : 4;
    if (grid > ws->num_sms * ctas) grid = ws->num_sms * ctas;
  }
  KbRed rd = kb_red(ws);
  if (dot == 0)
    kb_spmv_rowwise_kernel<0><<<grid, block, 0, st>>>(A->n_rows, k, A->rowptr, A->colidx, A->vals,
                                                      x, y, mode, z, coef, w, out, contiguous, rd);
  else if (dot == 1)
    kb_spmv_rowwise_kernel<1><<<grid, block, 0, st>>>(A->n_rows, k, A->rowptr, A->colidx, A->vals,
                                                      x, y, mode, z, coef, w, out, contiguous, rd);
  else
    kb_spmv_rowwise_kernel<2><<<grid, block, 0, st>>>(A->n_rows, k, A->rowptr, A->colidx, A->vals,
                                                      x, y, mode, z, coef, w, out, contiguous, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_spmm_is_lines(kb_csr_t A, int k, const double* x, int* yes) {
  KB_REQUIRE(A != nullptr && yes != nullptr, "null argument");
  KbLines g;
  bool var = false;
  *yes = (k > 1 && (A->schedule == 4 || A->schedule == 3) && A->forced != 1 &&
          kb_lines_geom(A, k, x, &g, &var))
             ? 1
             : 0;
  return KB_OK;
}

int kb_spmv_halo_add(kb_ws_t ws, int k, int64_t n_brows, double sign, const int32_t* rows,
                     const int32_t* hrowptr, const int32_t* hcol, const double* hval,
                     const double* xh, double* y, int dot, const double* w, double* out,
                     kb_halo_t halo, const int* srcs, int n_src, void* stream) {
  KB_REQUIRE(ws != nullptr, "null workspace");
  KB_REQUIRE(k >= 1 && k <= ws->max_k, "k exceeds workspace max_k");
  KB_REQUIRE(dot == 0 || dot == 1, "dot must be 0 or 1");
  KB_REQUIRE(dot == 0 || (w != nullptr && out != nullptr), "dot 1 needs w and out");
  KB_REQUIRE(halo == nullptr || (halo->opened && n_src >= 0 && n_src <= KB_BLOCK / 2),
             "halo not opened / too many sources");
  KB_REQUIRE(halo != nullptr || xh != nullptr || n_brows == 0, "need xh or a peer halo");
  cudaStream_t st = S(stream);
  if (n_brows == 0) return KB_OK;  // nothing to add (the dot slot keeps the local part)
  const int block = kb_block_for(k);
  int grid = kb_grid_for(ws, n_brows * (int64_t)k, block, 1);
  // a small kernel (a few MB): keep the number of blocks, and with it the serialized
  // arrivals on the reduction ticket, low
  if (grid > 2 * ws->num_sms) grid = 2 * ws->num_sms;
  KbRed rd = kb_red(ws);
  KbHalo hd;
  hd.peers = halo ? halo->dev.peers : nullptr;
  hd.rank = halo ? halo->dev.rank : 0;
  hd.size = halo ? halo->dev.size : 1;
  if (dot == 0)
    kb_spmv_halo_add_kernel<0><<<grid, block, 0, st>>>(n_brows, k, sign, rows, hrowptr, hcol, hval,
                                                       xh, y, w, out, hd, srcs, n_src, rd);
  else
    kb_spmv_halo_add_kernel<1><<<grid, block, 0, st>>>(n_brows, k, sign, rows, hrowptr, hcol, hval,
                                                       xh, y, w, out, hd, srcs, n_src, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// ------------------------------------------------------------ peer halo ---
int kb_halo_create(kb_halo_t* out, int rank, int size, int64_t data_bytes) {
  KB_REQUIRE(out != nullptr, "null handle pointer");
  KB_REQUIRE(size >= 1 && size <= 64 && rank >= 0 && rank < size, "bad rank/size");
  KB_REQUIRE(data_bytes >= 0, "negative size");
  kb_halo_s* h = new (std::nothrow) kb_halo_s();
  if (!h) return kb_fail(KB_ENOMEM, "out of host memory");
  memset(h, 0, sizeof(*h));
  h->dev.rank = rank;
  h->dev.size = size;
  h->data_bytes = (size_t)data_bytes;
  const size_t bytes = KB_HALO_DATA + ((size_t)data_bytes + 255) / 256 * 256 + 256;
  cudaError_t e = cudaMalloc(&h->base, bytes);
  if (e == cudaSuccess) e = cudaMemset(h->base, 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->peers_dev, sizeof(unsigned char*) * size);
  if (e != cudaSuccess) {
    if (h->base) cudaFree(h->base);
    delete h;
    return kb_fail(KB_ECUDA, "kb_halo_create: %s", cudaGetErrorString(e));
  }
  h->dev.peers = h->peers_dev;
  *out = h;
  return KB_OK;
}

int kb_halo_get_handle(kb_halo_t h, void* out64) {
  KB_REQUIRE(h != nullptr && out64 != nullptr, "null argument");
  cudaIpcMemHandle_t ih;
  KB_CUDA(cudaIpcGetMemHandle(&ih, h->base));
  memcpy(out64, &ih, 64);
  return KB_OK;
}

int kb_halo_open(kb_halo_t h, const void* handles) {
  KB_REQUIRE(h != nullptr && handles != nullptr, "null argument");
  KB_REQUIRE(!h->opened, "already opened");
  unsigned char* table[64];
  for (int p = 0; p < h->dev.size; ++p) {
    if (p == h->dev.rank) {
      table[p] = h->base;
      continue;
    }
    cudaIpcMemHandle_t ih;
    memcpy(&ih, (const char*)handles + 64 * (size_t)p, 64);
    void* base = nullptr;
    KB_CUDA(cudaIpcOpenMemHandle(&base, ih, cudaIpcMemLazyEnablePeerAccess));
    h->peer_base[p] = base;
    table[p] = (unsigned char*)base;
  }
  KB_CUDA(cudaMemcpy(h->peers_dev, table, sizeof(unsigned char*) * h->dev.size,
                     cudaMemcpyHostToDevice));
  h->opened = 1;
  return KB_OK;
}

int kb_halo_destroy(kb_halo_t h) {
  if (!h) return KB_OK;
  for (int p = 0; p < h->dev.size; ++p)
    if (h->peer_base[p]) cudaIpcCloseMemHandle(h->peer_base[p]);
  if (h->peers_dev) cudaFree(h->peers_dev);
  if (h->base) cudaFree(h->base);
  delete h;
  return KB_OK;
}

int kb_halo_data_ptr(kb_halo_t h, int rank, void** out) {
  KB_REQUIRE(h != nullptr && out != nullptr, "null argument");
  KB_REQUIRE(h->opened || rank == h->dev.rank, "halo not opened");
  KB_REQUIRE(rank >= 0 && rank < h->dev.size, "bad rank");
  unsigned char* b = rank == h->dev.rank ? h->base : (unsigned char*)h->peer_base[rank];
  KB_REQUIRE(b != nullptr, "peer not mapped");
  *out = b + KB_HALO_DATA;
  return KB_OK;
}

int kb_halo_error(kb_halo_t h, int* err) {
  KB_REQUIRE(h != nullptr && err != nullptr, "null argument");
  KB_CUDA(cudaMemcpy(err, h->base + KB_HALO_ERROR, sizeof(int), cudaMemcpyDeviceToHost));
  return KB_OK;
}

int kb_halo_push(kb_halo_t h, kb_ws_t ws, int k, int n_seg, const int64_t* segs, int64_t n_total,
                 const int32_t* idx, const double* x, void* stream) {
  KB_REQUIRE(h != nullptr && ws != nullptr, "null handle");
  KB_REQUIRE(h->opened, "halo not opened");
  KB_REQUIRE(k >= 1 && n_seg >= 0 && n_seg <= KB_BLOCK / 2 && n_total >= 0, "bad sizes");
  KB_REQUIRE(n_seg == 0 || (segs && idx && x), "null argument");
  // two elements per thread: the gather -> remote store chains are latency bound
  int grid = kb_grid_for(ws, n_total * (int64_t)k, KB_BLOCK, 2);
  if (grid > 2 * ws->num_sms) grid = 2 * ws->num_sms;
  kb_halo_push_kernel<<<grid, KB_BLOCK, 0, S(stream)>>>(k, n_seg, segs, n_total, idx, x, h->dev,
                                                        kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_pack_rows(kb_ws_t ws, int k, int64_t n_idx, const int32_t* idx, const double* x,
                 double* buf, void* stream) {
  KB_REQUIRE(ws != nullptr, "null workspace");
  KB_REQUIRE(k >= 1, "k must be positive");
  if (n_idx == 0) return KB_OK;
  const int grid = kb_grid_for(ws, n_idx * (int64_t)k, KB_BLOCK, 1);
  kb_pack_rows_kernel<<<grid, KB_BLOCK, 0, S(stream)>>>(n_idx, k, idx, x, buf, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// ------------------------------------------------------- vector kernels --
#define KB_VEC_PROLOGUE()                                              \
  KB_REQUIRE(ws != nullptr, "null workspace");                         \
  KB_REQUIRE(k >= 1 && k <= ws->max_k, "k exceeds workspace max_k");   \
  KB_REQUIRE(n >= 0, "negative length");                               \
  const int64_t total = n * (int64_t)k;                                \
  const int block = kb_block_for(k);                                   \
  cudaStream_t st = S(stream);                                         \
  KbRed rd = kb_red(ws);                                               \
  (void)rd

int kb_dot(kb_ws_t ws, int64_t n, int k, const double* x, const double* y, double* out,
           void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(out != nullptr, "null out");
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_dot_kernel<<<grid, block, 0, st>>>(total, k, x, y, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

static int kb_cg_update_xr_impl(kb_ws_t ws, int64_t n, int k, const double* rho, const double* pAp,
                                const double* pAp2, const double* p, const double* Ap, double* x,
                                double* r, double* rr_out, double* alpha_out, KbCgRecord rec,
                                void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(rho && pAp && Ap && r && rr_out, "null argument");
  KB_REQUIRE((x == nullptr) || (p != nullptr), "x update needs p");
  KB_REQUIRE(rec.step < 0 || (rec.crit && rec.hist && rec.stop_at), "record needs crit, hist, stop_at");
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  if (x != nullptr)
    kb_cg_update_xr_kernel<true><<<grid, block, 0, st>>>(total, k, rho, pAp, pAp2, p, Ap, x, r,
                                                         rr_out, alpha_out, rec, rd);
  else
    kb_cg_update_xr_kernel<false><<<grid, block, 0, st>>>(total, k, rho, pAp, pAp2, p, Ap, x, r,
                                                          rr_out, alpha_out, rec, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_cg_update_xr(kb_ws_t ws, int64_t n, int k, const double* rho, const double* pAp,
                    const double* pAp2, const double* p, const double* Ap, double* x, double* r,
                    double* rr_out, double* alpha_out, void* stream) {
  KbCgRecord rec = {-1, nullptr, nullptr, nullptr, nullptr};
  return kb_cg_update_xr_impl(ws, n, k, rho, pAp, pAp2, p, Ap, x, r, rr_out, alpha_out, rec, stream);
}

int kb_cg_update_xr_record(kb_ws_t ws, int64_t n, int k, const double* rho, const double* pAp,
                           const double* p, const double* Ap, double* x, double* r, double* rr_out,
                           double* alpha_out, int step, const double* crit, double* hist,
                           int* stop_at, double* rho_keep, void* stream) {
  KB_REQUIRE(step >= 0, "negative step");
  KbCgRecord rec = {step, crit, hist, stop_at, rho_keep};
  return kb_cg_update_xr_impl(ws, n, k, rho, pAp, nullptr, p, Ap, x, r, rr_out, alpha_out, rec,
                              stream);
}

int kb_cg_update_p(kb_ws_t ws, int64_t n, int k, int step, const double* rho_new,
                   const double* rho_old, const double* alpha, const double* crit, double* hist,
                   int* stop_at, double* rho_keep, const double* r, double* p, double* x, int what,
                   void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(what >= 1 && what <= 7, "what is a mask of 1 (update p), 2 (record), 4 (update x)");
  KB_REQUIRE(!(what & 2) || (rho_new && crit && hist && stop_at),
             "record needs rho_new, crit, hist, stop_at");
  KB_REQUIRE(!(what & 1) || (r && p && rho_new && rho_old), "p update needs r, p, rho_new, rho_old");
  KB_REQUIRE(!(what & 4) || (x && p && alpha), "x update needs x, p, alpha");
  const int grid = (what & 5) ? kb_grid_for(ws, total, block, KB_UNROLL) : 1;
  kb_cg_update_p_kernel<<<grid, block, 0, st>>>(total, k, step, rho_new, rho_old, alpha, crit, hist,
                                                stop_at, rho_keep, r, p, x, what, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

// Fused CG iteration on a 3-D constant-coefficient stencil (kb_march.cuh): p/x update fused with
// A p and <p, A p> (reads p_in with halo, writes the other p buffer), r update with A p
// recomputed from the ring.  Two launches and 64 n + masks bytes per iteration.
// Row-partitioned variant (s->masks_ext != NULL): the same two kernels on the ghost-extended row
// space [ghost plane | own planes | ghost plane]; r, p, p2 are bases of extended buffers, x of the
// own rows.  *ext receives a copy of the local matrix handle re-dimensioned to that space.
static bool kb_cg_fusable(const kb_cg_state* s, kb_csr_s* ext, int num_sms, int cfg[2],
                          KbMarch geo[2]) {
  if (!g_cg_fuse || s->k != 1 || s->p2 == nullptr || s->A->schedule != 4) return false;
  *ext = *s->A;
  if (s->masks_ext != nullptr) {
    ext->n_rows = ext->n_cols = s->n_ext;
    ext->masks = const_cast<uint16_t*>(s->masks_ext);
  }
  if (s->A->pat.off[3] != 0) return false;
  if (!kb_cg_pick_shape(ext, num_sms, 1, &cfg[0], &geo[0])) return false;
  if (!kb_cg_pick_shape(ext, num_sms, 2, &cfg[1], &geo[1])) return false;
  if (s->A->n_rows != s->n || s->A->n_cols != s->n) return false;
  const int P = geo[0].P;
  if (s->masks_ext != nullptr &&
      (s->own_lo != P || s->n_ext != s->n + 2 * (int64_t)P || s->n % P != 0))
    return false;
  return ((uintptr_t)s->p % 16 == 0) && ((uintptr_t)s->p2 % 16 == 0) && ((uintptr_t)s->r % 16 == 0);
}

int kb_cg_is_fused(const kb_cg_state* s, int* fused) {
  KB_REQUIRE(s != nullptr && fused != nullptr && s->A != nullptr, "null argument");
  KbMarch geo[2];
  int cfg[2];
  kb_csr_s ext;
  *fused = kb_cg_fusable(s, &ext, 148, cfg, geo) ? 1 : 0;
  return KB_OK;
}

int kb_cg_is_persistent(kb_ws_t ws, const kb_cg_state* s, int* yes) {
  KB_REQUIRE(ws != nullptr && s != nullptr && yes != nullptr && s->A != nullptr, "null argument");
  *yes = kb_cg_small_ok(ws, s) ? 1 : 0;
  return KB_OK;
}

static int kb_cg_run_impl(kb_ws_t ws, const kb_cg_state* s, int i0, int n_iters, int x_pending,
                          void* stream, cudaEvent_t* ev) {
  KB_REQUIRE(ws != nullptr && s != nullptr, "null argument");
  KB_REQUIRE(s->A && s->x && s->r && s->p && s->Ap && s->slots && s->crit && s->hist && s->stop_at,
             "null field in kb_cg_state");
  KB_REQUIRE(i0 >= 0 && n_iters >= 0, "negative iteration range");
  KB_REQUIRE(s->pcur == 0 || (s->pcur == 1 && s->p2 != nullptr), "pcur must name an existing buffer");
  if (ev == nullptr && kb_cg_small_ok(ws, s))  // launch-latency-bound sizes: one persistent launch
    return kb_cg_small_run(ws, s, i0, n_iters, x_pending, S(stream));
  const int k = s->k;
  double* sl = s->slots;
  const int* saved_gate = ws->gate;
  const int saved_tag = ws->gate_tag;
  const int saved_coll = ws->collective;
  if (g_part_dbg & 2) ws->collective = 0;
  int rc = KB_OK;
  KbMarch geo[2];
  int cfg[2] = {0, 0};
  kb_csr_s ext;
  const bool fused = kb_cg_fusable(s, &ext, ws->num_sms, cfg, geo);
  const bool parted = s->masks_ext != nullptr;  // row-partitioned: ghost-extended row space
  KB_REQUIRE(!parted || fused, "row-partitioned kb_cg_run needs the fused marching path");
  const int own_lo = parted ? (int)s->own_lo : 0;
  const int own_hi = own_lo + (int)s->n;
  double* pb[2] = {s->p, s->p2};
  int pc = s->pcur;
  cudaStream_t st = S(stream);
  for (int i = i0; i < i0 + n_iters && rc == KB_OK; ++i) {
    double* cur = sl + (size_t)(i % 2) * k;        // rho_i
    double* nxt = sl + (size_t)((i + 1) % 2) * k;  // rho_{i-1}, then rho_{i+1}
    double* alpha = sl + 2 * (size_t)k;
    double* pAp = sl + 3 * (size_t)k;
    double* rr = sl + 4 * (size_t)k;
    ws->gate = s->stop_at;
    ws->gate_tag = i;
    if (ev) cudaEventRecord(ev[3 * (i - i0)], st);
    if (fused) {
      KbMarchCg cg;
      memset(&cg, 0, sizeof(cg));
      cg.rec.step = -1;
      cg.own_lo = own_lo;
      cg.own_hi = own_hi;
      cg.own_pl0 = own_lo / geo[0].P;
      cg.own_pl1 = own_hi / geo[0].P;
      if (i > 0 || parted) {  // [x += alpha p;] p' = r + omega p (into the other buffer); <p', A p'>
        // row-partitioned, i == 0: omega = 0 / nz(0) from the permanently zero slot 6 and a
        // zero-filled p, so that p' = r on the ghost planes as well
        cg.rho_a = i > 0 ? cur : sl + 6 * (size_t)k;
        cg.rho_b = i > 0 ? nxt : sl + 6 * (size_t)k;
        cg.alpha_in = alpha;
        cg.r_in = s->r;
        cg.xv = (x_pending && i > 0) ? s->x - own_lo : nullptr;
        cg.p_out = pb[pc ^ 1];
        rc = parted ? kb_launch_march_cg<1, 1, true>(cfg[0], &ext, ws, geo[0], pb[pc], cg, pAp, st)
                    : kb_launch_march_cg<1, 1, false>(cfg[0], &ext, ws, geo[0], pb[pc], cg, pAp, st);
        pc ^= 1;
      } else {
        rc = kb_spmv(s->A, ws, k, pb[pc], s->Ap, 0, nullptr, nullptr, 1, pb[pc], pAp, stream);
      }
      if (ev) cudaEventRecord(ev[3 * (i - i0) + 1], st);
      if (rc == KB_OK) {  // alpha; r -= alpha (A p); <r, r>; record step i+1, rho_{i+1} -> nxt
        memset(&cg, 0, sizeof(cg));
        cg.own_lo = own_lo;
        cg.own_hi = own_hi;
        cg.own_pl0 = own_lo / geo[1].P;
        cg.own_pl1 = own_hi / geo[1].P;
        cg.push_lo = (parted && !(g_part_dbg & 1)) ? s->r_push_lo : nullptr;
        cg.push_hi = (parted && !(g_part_dbg & 1)) ? s->r_push_hi : nullptr;
        cg.rho_a = cur;
        cg.rho_b = pAp;
        cg.alpha_out = alpha;
        cg.r = s->r;
        cg.rec.step = i + 1;
        cg.rec.crit = s->crit;
        cg.rec.hist = s->hist - (size_t)(i0 + 1) * k;
        cg.rec.stop_at = s->stop_at;
        cg.rec.rho_keep = nxt;
        rc = parted ? kb_launch_march_cg<2, 2, true>(cfg[1], &ext, ws, geo[1], pb[pc], cg, rr, st)
                    : kb_launch_march_cg<2, 2, false>(cfg[1], &ext, ws, geo[1], pb[pc], cg, rr, st);
      }
      if (ev) cudaEventRecord(ev[3 * (i - i0) + 2], st);
      x_pending = 1;
      continue;
    }
    double* p = pb[pc];
    if (i > 0)
      rc = kb_cg_update_p(ws, s->n, k, 0, cur, nxt, alpha, nullptr, nullptr, nullptr, nullptr, s->r,
                          p, x_pending ? s->x : nullptr, x_pending ? 5 : 1, stream);
    if (ev) cudaEventRecord(ev[3 * (i - i0) + 1], st);  // here: after the p update
    if (rc == KB_OK)
      rc = kb_spmv(s->A, ws, k, p, s->Ap, 0, nullptr, nullptr, 1, p, pAp, stream);
    if (ev) cudaEventRecord(ev[3 * (i - i0) + 2], st);  // after A p
    if (rc == KB_OK)  // r update + <r,r> + record: hist row (i - i0) <- step i+1, rho_{i+1} -> nxt
      rc = kb_cg_update_xr_record(ws, s->n, k, cur, pAp, nullptr, s->Ap, nullptr, s->r, rr, alpha,
                                  i + 1, s->crit, s->hist - (size_t)(i0 + 1) * k, s->stop_at, nxt,
                                  stream);
    x_pending = 1;
  }
  ws->gate = saved_gate;
  ws->gate_tag = saved_tag;
  ws->collective = saved_coll;
  return rc;
}

int kb_cg_run(kb_ws_t ws, const kb_cg_state* s, int i0, int n_iters, int x_pending, void* stream) {
  return kb_cg_run_impl(ws, s, i0, n_iters, x_pending, stream, nullptr);
}

// kb_cg_run with CUDA events around every launch (on the launching stream).  Synchronises the
// stream at the end and returns the mean duration in ms of the step's phases:
//   fused path:        ms[0] = p/x update + A p + <p,Ap>,  ms[1] = r update + <r,r>,  ms[2] = 0
//   three-kernel path: ms[0] = p/x update,  ms[1] = A p + <p,Ap>,  ms[2] = r update + <r,r>
// total_ms: first event of the first iteration to the end of the last.  For measurement only.
int kb_cg_run_timed(kb_ws_t ws, const kb_cg_state* s, int i0, int n_iters, int x_pending,
                    void* stream, float* ms, float* total_ms) {
  KB_REQUIRE(ms != nullptr && total_ms != nullptr && n_iters >= 1 && n_iters <= 100000,
             "bad argument");
  KbMarch geo[2];
  int cfg[2];
  kb_csr_s ext;
  const bool fused = s != nullptr && s->A != nullptr && ws != nullptr &&
                     kb_cg_fusable(s, &ext, ws->num_sms, cfg, geo);
  const int ne = 3 * n_iters + 1;
  cudaEvent_t* ev = new (std::nothrow) cudaEvent_t[ne];
  if (!ev) return kb_fail(KB_ECUDA, "kb_cg_run_timed: out of memory");
  for (int i = 0; i < ne; ++i) cudaEventCreate(&ev[i]);
  int rc = kb_cg_run_impl(ws, s, i0, n_iters, x_pending, stream, ev);
  cudaEventRecord(ev[ne - 1], S(stream));
  cudaError_t e = cudaStreamSynchronize(S(stream));
  ms[0] = ms[1] = ms[2] = 0.f;
  *total_ms = 0.f;
  if (rc == KB_OK && e == cudaSuccess) {
    double acc[3] = {0, 0, 0};
    for (int i = 0; i < n_iters; ++i) {
      float a = 0, b = 0, c = 0;
      cudaEventElapsedTime(&a, ev[3 * i], ev[3 * i + 1]);
      cudaEventElapsedTime(&b, ev[3 * i + 1], ev[3 * i + 2]);
      cudaEventElapsedTime(&c, ev[3 * i + 2], ev[3 * i + 3]);
      if (fused) {  // events: start, after the first launch, after the second (== next start)
        acc[0] += a;
        acc[1] += b;
      } else {
        acc[0] += a;
        acc[1] += b;
        acc[2] += c;
      }
    }
    for (int q = 0; q < 3; ++q) ms[q] = (float)(acc[q] / n_iters);
    cudaEventElapsedTime(total_ms, ev[0], ev[ne - 1]);
  }
  for (int i = 0; i < ne; ++i) cudaEventDestroy(ev[i]);
  delete[] ev;
  if (rc != KB_OK) return rc;
  if (e != cudaSuccess) return kb_fail(KB_ECUDA, "kb_cg_run_timed: %s", cudaGetErrorString(e));
  return KB_OK;
}

int kb_axpy(kb_ws_t ws, int64_t n, int k, double sign, const double* coef, const double* x,
            double* y, void* stream) {
  KB_VEC_PROLOGUE();
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_axpy_kernel<<<grid, block, 0, st>>>(total, k, sign, coef, x, y, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_lincomb(kb_ws_t ws, int64_t n, int k, const double* ca, const double* x, const double* cb,
               const double* y, double* out, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(x != nullptr && out != nullptr, "null argument");
  KB_REQUIRE(cb == nullptr || y != nullptr, "cb needs y");
  KB_REQUIRE(ca != nullptr || cb != nullptr, "nothing to do: give ca and/or cb");
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_lincomb_kernel<<<grid, block, 0, st>>>(total, k, ca, x, cb, y, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_xpby(kb_ws_t ws, int64_t n, int k, const double* x, const double* coef, double* y,
            void* stream) {
  KB_VEC_PROLOGUE();
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_xpby_kernel<<<grid, block, 0, st>>>(total, k, x, coef, y, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_div_scale(kb_ws_t ws, int64_t n, int k, const double* x, const double* coef, double* out,
                 void* stream) {
  KB_VEC_PROLOGUE();
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_div_scale_kernel<<<grid, block, 0, st>>>(total, k, x, coef, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_add(kb_ws_t ws, int64_t n, int k, const double* x, const double* y, double* out,
           void* stream) {
  KB_VEC_PROLOGUE();
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_add_kernel<<<grid, block, 0, st>>>(total, x, y, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_axpy_dot(kb_ws_t ws, int64_t n, int k, const double* coef, const double* scale,
                const double* u, double* w, int dot, const double* z, double* out, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(coef && u && w, "null argument");
  KB_REQUIRE(dot >= 0 && dot <= 2, "dot must be 0, 1 or 2");
  KB_REQUIRE(dot == 0 || out != nullptr, "dot needs out");
  KB_REQUIRE(dot != 1 || z != nullptr, "dot 1 needs z");
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  if (dot == 0)
    kb_axpy_dot_kernel<0><<<grid, block, 0, st>>>(total, k, coef, scale, u, w, z, out, rd);
  else if (dot == 1)
    kb_axpy_dot_kernel<1><<<grid, block, 0, st>>>(total, k, coef, scale, u, w, z, out, rd);
  else
    kb_axpy_dot_kernel<2><<<grid, block, 0, st>>>(total, k, coef, scale, u, w, z, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_axpy_dot_minres(kb_ws_t ws, int64_t n, int k, const double* coef, const double* u, double* w,
                       int iter, const kb_minres_state* stt, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(coef && u && w && stt, "null argument");
  KB_REQUIRE(stt->ww != nullptr && stt->alpha != nullptr, "null field in kb_minres_state");
  KB_REQUIRE(total > 0, "empty vectors");
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_launch_pdl(g_pdl != 0, kb_axpy_dot_minres_kernel, dim3(grid), dim3(block), 0, st, total, k,
                coef, u, w, (double*)stt->ww, iter, *stt, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_axpy_dot_gmres(kb_ws_t ws, int64_t n, int k, const double* coef, const double* u, double* w,
                      int iter, const kb_gmres_state* stt, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(coef && u && w && stt, "null argument");
  KB_REQUIRE(stt->ww != nullptr && stt->have_h == 0, "needs the Gram-Schmidt state (ww, have_h == 0)");
  KB_REQUIRE(total > 0, "empty vectors");
  KB_REQUIRE(iter >= 0 && iter < stt->maxiter, "iter exceeds the Hessenberg storage");
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_axpy_dot_gmres_kernel<<<grid, block, 0, st>>>(total, k, coef, u, w, (double*)stt->ww, iter,
                                                   *stt, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_minres_scalar(kb_ws_t ws, int k, int iter, const kb_minres_state* stt, void* stream) {
  KB_REQUIRE(ws != nullptr && stt != nullptr, "null argument");
  KB_REQUIRE(k >= 1 && k <= ws->max_k, "k exceeds workspace max_k");
  const int block = ((k + 31) / 32) * 32;
  kb_minres_scalar_kernel<<<1, block, 0, S(stream)>>>(k, iter, *stt, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_minres_update(kb_ws_t ws, int64_t n, int k, const double* coefs, const double* v,
                     double* W0, const double* W1, const double* Av, double* yk, double* vnext,
                     const double* MAv, double* pnext, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(coefs && v && W0 && W1 && Av && yk && vnext, "null argument");
  KB_REQUIRE((MAv == nullptr) == (pnext == nullptr), "MAv and pnext go together");
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_MR_UNROLL);
  if (MAv)
    kb_minres_update_kernel<true><<<grid, block, 0, st>>>(total, k, coefs, v, W0, W1, Av, yk, vnext,
                                                          MAv, pnext, rd);
  else
    kb_launch_pdl(g_pdl != 0, kb_minres_update_kernel<false>, dim3(grid), dim3(block), 0, st, total,
                  k, coefs, v, W0, W1, Av, yk, vnext, (const double*)nullptr, (double*)nullptr, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_gmres_scalar(kb_ws_t ws, int k, int iter, const kb_gmres_state* stt, void* stream) {
  KB_REQUIRE(ws != nullptr && stt != nullptr, "null argument");
  KB_REQUIRE(k >= 1 && k <= ws->max_k, "k exceeds workspace max_k");
  KB_REQUIRE(iter >= 0 && iter < stt->maxiter, "iter out of range");
  const int block = ((k + 31) / 32) * 32;
  kb_gmres_scalar_kernel<<<1, block, 0, S(stream)>>>(k, iter, *stt, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_gmres_solve_y(kb_ws_t ws, int k, int m, int maxiter, const double* R, const double* y,
                     double* yy, void* stream) {
  KB_REQUIRE(ws != nullptr && R && y && yy, "null argument");
  KB_REQUIRE(m >= 0 && m <= maxiter, "m out of range");
  if (m == 0) return KB_OK;
  kb_gmres_solve_y_kernel<<<(k + 63) / 64, 64, 0, S(stream)>>>(k, m, maxiter, R, y, yy, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_basis_combine(kb_ws_t ws, int64_t n, int k, int m, const double* yy, const double* Vbuf,
                     int64_t vstride, const double* x0, double* out, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(x0 && out && (m == 0 || (yy && Vbuf)), "null argument");
  if (total == 0) return KB_OK;
  const int grid = kb_grid_for(ws, total, block, KB_UNROLL);
  kb_basis_combine_kernel<<<grid, block, 0, st>>>(total, k, m, yy, Vbuf, vstride, x0, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_multi_dot(kb_ws_t ws, int64_t n, int k, int cnt, const double* V, int64_t vstride,
                 const double* w, double* out, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(cnt >= 0 && (cnt == 0 || (V && w && out)), "null argument");
  // cnt * k sums per launch are bounded by the partials buffer / mailbox width and the block
  int cap = ws->max_k < block ? ws->max_k : block;
  int per = cap / k;
  if (per < 1) per = 1;
  const int jcmax = g_cgs_jc == 16 ? 16 : 8;
  if (per > jcmax) per = jcmax;
  const int grid = kb_grid_for(ws, total, block, KB_MD_UNROLL);
  for (int j0 = 0; j0 < cnt; j0 += per) {
    const int c = cnt - j0 < per ? cnt - j0 : per;
    const double* Vj = V + (size_t)j0 * vstride;
    double* oj = out + (size_t)j0 * k;
    if (c > 8)
      kb_multi_dot_kernel<16><<<grid, block, 0, st>>>(total, k, c, Vj, vstride, w, oj, rd);
    else if (c > 4)
      kb_multi_dot_kernel<8><<<grid, block, 0, st>>>(total, k, c, Vj, vstride, w, oj, rd);
    else if (c > 2)
      kb_multi_dot_kernel<4><<<grid, block, 0, st>>>(total, k, c, Vj, vstride, w, oj, rd);
    else
      kb_multi_dot_kernel<2><<<grid, block, 0, st>>>(total, k, c, Vj, vstride, w, oj, rd);
    KB_LAUNCH_CHECK();
  }
  return KB_OK;
}

int kb_multi_axpy(kb_ws_t ws, int64_t n, int k, int m, const double* h, const double* P,
                  int64_t pstride, double* w, int dot, double* out, void* stream) {
  KB_VEC_PROLOGUE();
  KB_REQUIRE(m >= 0 && w && (m == 0 || (h && P)), "null argument");
  KB_REQUIRE(dot == 0 || dot == 2, "dot must be 0 or 2");
  KB_REQUIRE(dot == 0 || out, "dot needs out");
  const int grid = kb_grid_for(ws, total, block, KB_MD_UNROLL);
  if (dot == 2)
    kb_multi_axpy_kernel<2><<<grid, block, 0, st>>>(total, k, m, h, P, pstride, w, out, rd);
  else
    kb_multi_axpy_kernel<0><<<grid, block, 0, st>>>(total, k, m, h, P, pstride, w, out, rd);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

}  // extern "C"

// ---- tall-skinny block products on the DMMA pipe (kb_block.cuh; utils.py:100-118) ----
static inline int kb_block_grid(const kb_ws_s* ws, int64_t n) {
  int64_t need = (n + 127) / 128;
  int64_t cap = (int64_t)ws->num_sms * 2;  // 2 CTAs of 8 warps per SM: 64 KB of loads in flight
  if (cap > KB_MAX_BLOCKS) cap = KB_MAX_BLOCKS;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

template <int MODE>
static int kb_block_apply_mode(kb_ws_s* ws, int64_t n, int k, int l, const double* X, int64_t ldx,
                               const double* C, int64_t ldc, const double* Y, int64_t ldy,
                               double* Z, int64_t ldz, cudaStream_t st) {
  const int hk = k > 8 ? 2 : 1, hl = l > 8 ? 2 : 1;
  const int grid = kb_block_grid(ws, n);
  KbRed rd = kb_red(ws);
#define KB_APPLY(HK, HL)                                                                       \
  kb_block_apply_kernel<HK, HL, MODE><<<grid, KB_BG_WARPS * 32, 0, st>>>(n, k, l, X, ldx, C, ldc, \
                                                                         Y, ldy, Z, ldz, rd)
  if (hk == 1 && hl == 1) KB_APPLY(1, 1);
  else if (hk == 1) KB_APPLY(1, 2);
  else if (hl == 1) KB_APPLY(2, 1);
  else KB_APPLY(2, 2);
#undef KB_APPLY
  KB_LAUNCH_CHECK();
  return KB_OK;
}

extern "C" {

int kb_block_gram(kb_ws_t ws, int64_t n, int k, int l, const double* X, int64_t ldx,
                  const double* Y, int64_t ldy, double* G, int64_t ldg, double* Gacc,
                  int64_t ldacc, int flags, void* stream) {
  KB_REQUIRE(ws != nullptr, "null workspace");
  KB_REQUIRE(n >= 0, "negative length");
  KB_REQUIRE(k >= 1 && k <= 16 && l >= 1 && l <= 16, "1 <= k, l <= 16 columns per call");
  KB_REQUIRE(X && Y && G, "null argument");
  KB_REQUIRE(ldx >= k && ldy >= l && ldg >= l && (Gacc == nullptr || ldacc >= l),
             "leading dimension smaller than the column count");
  const int hx = k > 8 ? 2 : 1, hy = l > 8 ? 2 : 1;
  KB_REQUIRE(ws->max_k >= hx * hy * 64, "workspace too narrow: create it with max_k >= 256");
  KbRed rd = kb_red(ws);
  KB_REQUIRE(!rd.collective || ws->comm->max_k >= hx * hy * 64, "communicator too narrow");
  const int grid = kb_block_grid(ws, n);
  cudaStream_t st = S(stream);
#define KB_GRAM(HX, HY)                                                                      \
  kb_block_gram_kernel<HX, HY><<<grid, KB_BG_WARPS * 32, 0, st>>>(n, k, l, X, ldx, Y, ldy, G, \
                                                                  ldg, Gacc, ldacc, flags, rd)
  if (hx == 1 && hy == 1) KB_GRAM(1, 1);
  else if (hx == 1) KB_GRAM(1, 2);
  else if (hy == 1) KB_GRAM(2, 1);
  else KB_GRAM(2, 2);
#undef KB_GRAM
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_block_apply(kb_ws_t ws, int64_t n, int k, int l, const double* X, int64_t ldx,
                   const double* C, int64_t ldc, const double* Y, int64_t ldy, double* Z,
                   int64_t ldz, int mode, void* stream) {
  KB_REQUIRE(ws != nullptr, "null workspace");
  KB_REQUIRE(n >= 0, "negative length");
  KB_REQUIRE(k >= 1 && k <= 16 && l >= 1 && l <= 16, "1 <= k, l <= 16 columns per call");
  KB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  KB_REQUIRE(X && C && Z && (mode == 0 || Y), "null argument");
  KB_REQUIRE(ldx >= k && ldc >= l && ldz >= l && (mode == 0 || ldy >= l),
             "leading dimension smaller than the column count");
  if (mode == 0) return kb_block_apply_mode<0>(ws, n, k, l, X, ldx, C, ldc, Y, ldy, Z, ldz, S(stream));
  if (mode == 1) return kb_block_apply_mode<1>(ws, n, k, l, X, ldx, C, ldc, Y, ldy, Z, ldz, S(stream));
  return kb_block_apply_mode<2>(ws, n, k, l, X, ldx, C, ldc, Y, ldy, Z, ldz, S(stream));
}

int kb_house_make(kb_ws_t ws, int64_t n, int64_t off, const double* x, double* v, double* params,
                  double* scratch, void* stream) {
  return kb_house_make2(ws, n, off, x, v, params, scratch, 0, stream);
}

int kb_house_make2(kb_ws_t ws, int64_t n, int64_t off, const double* x, double* v, double* params,
                   double* scratch, int lapack_sign, void* stream) {
  KB_REQUIRE(ws && x && v && params && scratch, "null argument");
  KB_REQUIRE(off >= 0 && off < n, "offset out of range");
  int rc = kb_dot(ws, n - off - 1, 1, x + off + 1, x + off + 1, scratch, stream);
  if (rc != KB_OK) return rc;
  kb_house_params_kernel<<<1, 32, 0, S(stream)>>>(x, off, scratch, params, lapack_sign ? 1 : 0,
                                                  kb_red(ws));
  KB_LAUNCH_CHECK();
  const int grid = kb_grid_for(ws, n, KB_BLOCK, 2);
  kb_house_fill_kernel<<<grid, KB_BLOCK, 0, S(stream)>>>(n, off, x, params, v, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_house_hlast(kb_ws_t ws, const double* w, int64_t off, const double* v, const double* params,
                   const double* tau, double* h_out, void* stream) {
  KB_REQUIRE(ws && w && v && params && tau && h_out, "null argument");
  kb_house_hlast_kernel<<<1, 32, 0, S(stream)>>>(w, off, v, params, tau, h_out, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_poke(kb_ws_t ws, int op, double* x, int64_t idx, const double* s, double val, double* dst,
            void* stream) {
  KB_REQUIRE(ws != nullptr && x != nullptr, "null argument");
  KB_REQUIRE(op >= 0 && op <= 2, "op must be 0, 1 or 2");
  kb_poke_kernel<<<1, 1, 0, S(stream)>>>(op, x, idx, s, val, dst, kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_lartg(int n, const double* f, const double* g, double* out, void* stream) {
  if (n <= 0) return KB_OK;
  kb_lartg_kernel<<<(n + 127) / 128, 128, 0, S(stream)>>>(n, f, g, out);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_stencil7(int nx, int ny, int nz, int z_lo, int z_hi, const double* coeffs_host,
                int32_t* rowptr, int32_t* colidx, double* vals, void* stream) {
  KB_REQUIRE(nx > 0 && ny > 0 && nz > 0 && z_lo >= 0 && z_hi <= nz && z_lo < z_hi, "bad grid");
  KB_REQUIRE(coeffs_host != nullptr && rowptr != nullptr, "null argument");
  KB_REQUIRE((vals == nullptr) == (colidx == nullptr), "colidx and vals go together");
  KbStencil7 p;
  p.nx = nx;
  p.ny = ny;
  p.nz = nz;
  p.z_lo = z_lo;
  p.z_hi = z_hi;
  // coeffs = {diag, lx, ly, lz, ux, uy, uz} -> entry order z-1,y-1,x-1,diag,x+1,y+1,z+1
  p.c[0] = coeffs_host[3];
  p.c[1] = coeffs_host[2];
  p.c[2] = coeffs_host[1];
  p.c[3] = coeffs_host[0];
  p.c[4] = coeffs_host[4];
  p.c[5] = coeffs_host[5];
  p.c[6] = coeffs_host[6];
  const int64_t n_loc = (int64_t)nx * ny * (z_hi - z_lo);
  int64_t grid = (n_loc + KB_BLOCK - 1) / KB_BLOCK;
  if (grid > 148 * 16) grid = 148 * 16;
  kb_stencil7_kernel<<<(int)grid, KB_BLOCK, 0, S(stream)>>>(p, rowptr, colidx, vals);
  KB_LAUNCH_CHECK();
  return KB_OK;
}

}  // extern "C"
