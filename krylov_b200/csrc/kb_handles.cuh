// The CSR matrix handle behind kb_csr_t and the plain structs it carries (shared by the
// translation units of the library).
#pragma once
#include "kb_common.cuh"

struct KbPattern {
  int nd;
  int off[16];   // ascending diagonal offsets col - row
  // x windows of the windowed kernel: nearby diagonals share one contiguous window
  int nw;        // number of windows (0: windowed kernel not applicable)
  int wlo[8];    // window g of the tile starting at row r0 begins at x[r0 + wlo[g]]
  int wspan[8];  // ... and holds ROWS + wspan[g] entries
  int grp[16];   // window of diagonal d
  int dwlo[16];  // wlo[grp[d]]
};


struct KbConstVals {
  double c[8];
};

struct kb_csr_s {
  int64_t n_rows, n_cols, nnz;
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  int padded;
  int max_row_len;
  int schedule;  // 1 row-wise, 2 TMA stream, 3 offset-pattern compressed TMA stream,
                 // 4 stencil (offset pattern + constant diagonals: no value stream either),
                 // 5 merge (nonzero-balanced tiles: long or skewed rows, k == 1)
  int forced;    // user override (0 = auto)
  // offset-pattern compression (library-owned): one 16-bit mask per row
  uint16_t* masks;
  KbPattern pat;
  int pattern_ok;
  KbConstVals cv;  // one value per diagonal when constv
  int constv;
  // merge schedule (library-owned, built at the first product that uses it)
  int4* merge_meta;  // per tile {first row, last row, tail, ends a long row}
  double* carry;     // per tile two partial sums of long rows
  int4* merge_fix;   // per long row {row, first tile, last tile, -}
  int merge_T, n_mtiles, n_fix;
};

// kb_merge.cu
extern int g_merge_cfg, g_merge_ctas, g_merge_order;
void kb_merge_release(kb_csr_s* h);
int kb_launch_merge(kb_csr_s* A, kb_ws_s* ws, int dot, const double* x, double* y, int mode,
                    const double* z, const double* coef, const double* w, double* out,
                    cudaStream_t st);

// kb_small.cu
extern int g_small_n;
bool kb_cg_small_ok(const kb_ws_s* ws, const kb_cg_state* s);
int kb_cg_small_run(kb_ws_s* ws, const kb_cg_state* s, int i0, int n_iters, int x_pending,
                    cudaStream_t st);
