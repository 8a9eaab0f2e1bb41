// Launch code of the merge schedule (own translation unit: the library builds in parallel).
#include <new>

#include "kb_merge.cuh"

int g_merge_cfg = 0;    // kb_tune key 25
int g_merge_ctas = 0;   // kb_tune key 26
int g_merge_order = 2;  // kb_tune key 27

void kb_merge_release(kb_csr_s* h) {
  if (h->merge_meta) cudaFree(h->merge_meta);
  if (h->carry) cudaFree(h->carry);
  if (h->merge_fix) cudaFree(h->merge_fix);
  h->merge_meta = nullptr;
  h->carry = nullptr;
  h->merge_fix = nullptr;
  h->merge_T = 0;
  h->n_fix = 0;
}

// Tiles of T nonzeros.  The per-tile table, the carry slots and the list of long rows belong
// to the matrix and are built (once per tile shape) on the stream of the first product.
static int kb_merge_prepare(kb_csr_s* A, int T, int TAIL, cudaStream_t st) {
  if (A->merge_T == T * 1024 + TAIL) return KB_OK;
  kb_merge_release(A);
  const int n_tiles = (int)((A->nnz + T - 1) / T);
  KB_CUDA(cudaMalloc(&A->merge_meta, sizeof(int4) * (size_t)n_tiles));
  KB_CUDA(cudaMalloc(&A->carry, 16 * (size_t)n_tiles));
  kb_merge_tiles_kernel<<<(n_tiles + 255) / 256, 256, 0, st>>>(
      (int)A->n_rows, (int)A->nnz, n_tiles, T, TAIL, A->rowptr, A->merge_meta);
  KB_LAUNCH_CHECK();
  int* d_n = nullptr;
  KB_CUDA(cudaMalloc(&d_n, sizeof(int)));
  kb_merge_fixlist_kernel<<<1, 1024, 0, st>>>(n_tiles, T, A->rowptr, A->merge_meta, nullptr, d_n);
  int n_fix = 0;
  cudaError_t e = cudaMemcpyAsync(&n_fix, d_n, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // once per matrix, not on the hot path
  if (e == cudaSuccess && n_fix > 0) {
    e = cudaMalloc(&A->merge_fix, sizeof(int4) * (size_t)n_fix);
    if (e == cudaSuccess)
      kb_merge_fixlist_kernel<<<1, 1024, 0, st>>>(n_tiles, T, A->rowptr, A->merge_meta,
                                                  A->merge_fix, d_n);
  }
  cudaFree(d_n);
  if (e != cudaSuccess) {
    kb_merge_release(A);
    return kb_fail(KB_ECUDA, "merge schedule set-up: %s", cudaGetErrorString(e));
  }
  KB_LAUNCH_CHECK();
  A->merge_T = T * 1024 + TAIL;
  A->n_mtiles = n_tiles;
  A->n_fix = n_fix;
  return KB_OK;
}

template <int T, int TAIL, int STAGES, int NG, int MINB, int DOT>
static int kb_launch_merge_cfg(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                               const double* z, const double* coef, const double* w, double* out,
                               cudaStream_t st) {
  typedef KbMergeSmem<T, TAIL, STAGES, NG> Smem;
  static bool configured[64] = {false};
  static int resident[64] = {0};
  auto kern = kb_spmv_merge_kernel<T, TAIL, STAGES, NG, MINB, DOT>;
  int dev = 0;
  KB_CUDA(cudaGetDevice(&dev));
  KB_REQUIRE(dev >= 0 && dev < 64, "device ordinal out of range");
  if (!configured[dev]) {
    KB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(Smem)));
    KB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident[dev], kern, NG * 256 + 32,
                                                          sizeof(Smem)));
    configured[dev] = true;
  }
  KB_REQUIRE(resident[dev] >= 1, "merge kernel does not fit an SM");
  const int rc = kb_merge_prepare(A, T, TAIL, st);
  if (rc != KB_OK) return rc;
  int per_sm = resident[dev];
  if (g_merge_ctas > 0 && g_merge_ctas < per_sm) per_sm = g_merge_ctas;
  int grid = ws->num_sms * per_sm;
  if (grid > (A->n_mtiles + NG - 1) / NG) grid = (A->n_mtiles + NG - 1) / NG;
  if (grid > KB_MAX_BLOCKS) grid = KB_MAX_BLOCKS;
  KbMergeOrder ord = {g_merge_order, ws->num_sms, per_sm};
  KbRed rd = kb_red(ws);
  KbRed rd_main = rd;
  if (A->n_fix > 0) rd_main.collective = 0;  // the second launch finishes the dot
  kern<<<grid, NG * 256 + 32, sizeof(Smem), st>>>((int)A->n_rows, (int)A->nnz, A->n_mtiles, ord,
                                             A->rowptr, A->colidx, A->vals, A->merge_meta,
                                             A->carry, x, y, mode, z, coef, w, out, rd_main);
  KB_LAUNCH_CHECK();
  if (A->n_fix > 0) {
    int fgrid = (A->n_fix + 7) / 8;
    if (fgrid > ws->num_sms * 8) fgrid = ws->num_sms * 8;
    if (fgrid > KB_MAX_BLOCKS) fgrid = KB_MAX_BLOCKS;
    kb_merge_fix_kernel<DOT><<<fgrid, 256, 0, st>>>(A->n_fix, A->merge_fix, A->carry, y, mode, z,
                                                    coef, w, out, rd);
    KB_LAUNCH_CHECK();
  }
  return KB_OK;
}

template <int DOT>
static int kb_launch_merge_dot(kb_csr_s* A, kb_ws_s* ws, const double* x, double* y, int mode,
                               const double* z, const double* coef, const double* w, double* out,
                               cudaStream_t st) {
  switch (g_merge_cfg) {
    case 1: return kb_launch_merge_cfg<4096, 512, 2, 1, 2, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 2: return kb_launch_merge_cfg<1024, 256, 3, 1, 4, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 3: return kb_launch_merge_cfg<2048, 256, 2, 1, 4, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 4: return kb_launch_merge_cfg<1024, 256, 2, 3, 1, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 5: return kb_launch_merge_cfg<1024, 256, 3, 3, 1, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 6: return kb_launch_merge_cfg<2048, 512, 2, 1, 3, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    case 7: return kb_launch_merge_cfg<1024, 256, 2, 2, 2, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
    // default: three groups of 256 threads per CTA on neighbouring tiles of 2048 nonzeros (one
    // CTA per SM, contiguous tile ranges): best of profiles/r2_spmv_general_v4.txt on banded
    // and FEM-like matrices, within 3 % of the best on uniformly random columns
    default: return kb_launch_merge_cfg<2048, 256, 2, 3, 1, DOT>(A, ws, x, y, mode, z, coef, w, out, st);
  }
}

int kb_launch_merge(kb_csr_s* A, kb_ws_s* ws, int dot, const double* x, double* y, int mode,
                    const double* z, const double* coef, const double* w, double* out,
                    cudaStream_t st) {
  if (dot == 0) return kb_launch_merge_dot<0>(A, ws, x, y, mode, z, coef, w, out, st);
  if (dot == 1) return kb_launch_merge_dot<1>(A, ws, x, y, mode, z, coef, w, out, st);
  return kb_launch_merge_dot<2>(A, ws, x, y, mode, z, coef, w, out, st);
}
