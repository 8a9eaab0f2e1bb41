// Streaming vector kernels: every one is a single pass over its operands
// (HBM-bound; bytes per element are stated per kernel), vectors are (n, k)
// row-major and flat-indexed, per-column scalars come from device memory.
//
// Indexing rule shared by all kernels here: blockDim.x % k == 0 and every tile
// starts at a multiple of blockDim.x, so thread t only ever touches column
// c = t % k.  That makes the per-column coefficients loop-invariant and the
// column-wise reductions a plain per-thread accumulation.
//
// Work split: tiles of blockDim.x * KB_UNROLL contiguous elements, tile t goes
// to block t % gridDim.x.  At any moment the resident blocks stream one
// contiguous window of every operand (DRAM-page and TLB friendly; a plain
// grid-stride loop with megabyte strides between a thread's loads was 2x
// slower for some grid sizes at n = 512^3), and a thread's KB_UNROLL loads per
// operand are issued before the first use.
#pragma once
#include "kb_common.cuh"

#define KB_UNROLL 4

// for (tiles of this block) { full tile: FULL(i) unrolled over u ; tail: guarded }
#define KB_TILE_LOOP_BEGIN(total) KB_TILE_LOOP_BEGIN_U(total, KB_UNROLL)
#define KB_TILE_LOOP_BEGIN_U(total, U)                                                \
  const int64_t kb_tile = (int64_t)blockDim.x * (U);                                  \
  for (int64_t kb_base = (int64_t)blockIdx.x * kb_tile; kb_base < (total);            \
       kb_base += (int64_t)gridDim.x * kb_tile) {                                     \
    const int64_t kb_e0 = kb_base + threadIdx.x;                                      \
    const bool kb_full = kb_base + kb_tile <= (total);
#define KB_TILE_LOOP_END }
#define KB_IDX(u) (kb_e0 + (int64_t)(u) * blockDim.x)

// ---------------------------------------------------------------- dot ----
// out[c] = sum_i x[i,c] * y[i,c]        16 B/element (8 when x == y)
// reference: _helpers.py:101-110 (np.dot / einsum "i...,i...->...")
__global__ void __launch_bounds__(KB_BLOCK)
kb_dot_kernel(int64_t total, int k, const double* __restrict__ x, const double* __restrict__ y,
              double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  double acc = 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double a[KB_UNROLL], b[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      a[u] = x[KB_IDX(u)];
      b[u] = y[KB_IDX(u)];
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) acc = fma(a[u], b[u], acc);
  } else {
    for (int u = 0; u < KB_UNROLL; ++u)
      if (KB_IDX(u) < total) acc = fma(x[KB_IDX(u)], y[KB_IDX(u)], acc);
  }
  KB_TILE_LOOP_END
  kb_grid_colsum(acc, k, rd, out, sm);
}

// ------------------------------------------------------------ CG: x, r ---
// alpha = rho / nz(pAp [+ pAp2]);  [x += alpha p;]  r -= alpha Ap;  rr = <r,r>
// 48 B/element (reads x p r Ap, writes x r); 24 B/element with UPDX == false,
// when the x update is deferred into the next p update (kb_cg_update_p_kernel,
// what & 4), which streams p anyway.     cg.py:185,196,200,209
// Optional fused record (single GPU / peer-memory all-reduce, where the last block holds
// the final <r,r>): what kb_cg_update_p does with `what & 2`, without a launch of its own.
struct KbCgRecord {
  int step;               // < 0: no record
  const double* crit;
  double* hist;
  int* stop_at;
  double* rho_keep;
};

template <bool UPDX>
__global__ void __launch_bounds__(KB_BLOCK, 4)
kb_cg_update_xr_kernel(int64_t total, int k, const double* __restrict__ rho,
                       const double* __restrict__ pAp, const double* __restrict__ pAp2,
                       const double* __restrict__ p, const double* __restrict__ Ap,
                       double* __restrict__ x, double* __restrict__ r, double* __restrict__ out,
                       double* __restrict__ alpha_out, KbCgRecord rec, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  const int c = threadIdx.x % k;
  double d = pAp[c];
  if (pAp2 != nullptr) d += pAp2[c];
  const double alpha = rho[c] / kb_nz(d);
  // persistent copy of alpha (state slot: only gated kernels write it, so it survives
  // un-gated NCCL all-reduces of the landing slots after on-device convergence)
  if (alpha_out != nullptr && blockIdx.x == 0 && threadIdx.x < k) alpha_out[c] = alpha;
  double acc = 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double xv[KB_UNROLL], pv[KB_UNROLL], rv[KB_UNROLL], av[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      const int64_t i = KB_IDX(u);
      if (UPDX) {
        xv[u] = x[i];
        pv[u] = p[i];
      }
      rv[u] = r[i];
      av[u] = Ap[i];
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      const int64_t i = KB_IDX(u);
      if (UPDX) x[i] = kb_mul_add(alpha, pv[u], xv[u]);
      const double rn = kb_mul_sub(alpha, av[u], rv[u]);
      r[i] = rn;
      acc = fma(rn, rn, acc);
    }
  } else {
    for (int u = 0; u < KB_UNROLL; ++u) {
      const int64_t i = KB_IDX(u);
      if (i < total) {
        if (UPDX) x[i] = kb_mul_add(alpha, p[i], x[i]);
        const double rn = kb_mul_sub(alpha, Ap[i], r[i]);
        r[i] = rn;
        acc = fma(rn, rn, acc);
      }
    }
  }
  KB_TILE_LOOP_END
  const bool last = kb_grid_colsum(acc, k, rd, out, sm);
  if (last && rec.step >= 0) {  // cg.py:156,214-217 in the reduction's finishing block
    int ok = 1;
    __syncthreads();  // out[] written by threads t < k of this block
    if (threadIdx.x < k) {
      const double rn = out[threadIdx.x];
      if (rec.rho_keep != nullptr) rec.rho_keep[threadIdx.x] = rn;
      const double nrm = sqrt(rn);
      rec.hist[(size_t)rec.step * k + threadIdx.x] = nrm;
      ok = (nrm <= rec.crit[threadIdx.x]) ? 1 : 0;
    }
    const int all_ok = __syncthreads_and(ok);
    if (all_ok && threadIdx.x == 0) *rec.stop_at = rec.step;
  }
}

// ------------------------------------------------------------- CG: p -----
// what & 2 (block 0): hist[step] = sqrt(rho_new); rho_keep = rho_new (state copy);
//                     all columns <= crit -> *stop_at = step
// what & 4 (all):     x += alpha p     (the deferred update of the previous iteration,
//                     taken before p is overwritten)                       cg.py:196
// what & 1 (all):     omega = rho_new / nz(rho_old);  p = r + omega p     cg.py:175-178
// 24 B/element (what = 1), 40 B/element (what = 5), 24 B/element (what = 4)
__global__ void __launch_bounds__(KB_BLOCK)
kb_cg_update_p_kernel(int64_t total, int k, int step, const double* __restrict__ rho_new,
                      const double* __restrict__ rho_old, const double* __restrict__ alpha_in,
                      const double* __restrict__ crit, double* __restrict__ hist, int* stop_at,
                      double* __restrict__ rho_keep, const double* __restrict__ r,
                      double* __restrict__ p, double* __restrict__ x, int what, KbRed rd) {
  if (kb_gated(rd)) return;
  const int c = threadIdx.x % k;
  if ((what & 2) && blockIdx.x == 0) {
    int ok = 1;
    if (threadIdx.x < k) {
      const double rn = rho_new[c];
      if (rho_keep != nullptr) rho_keep[c] = rn;
      const double nrm = sqrt(rn);
      hist[(size_t)step * k + c] = nrm;
      ok = (nrm <= crit[c]) ? 1 : 0;
    }
    const int all_ok = __syncthreads_and(ok);
    if (all_ok && threadIdx.x == 0) *stop_at = step;
  }
  if (!(what & 5)) return;
  const bool updp = (what & 1) != 0, updx = (what & 4) != 0;
  const double omega = updp ? rho_new[c] / kb_nz(rho_old[c]) : 0.0;
  const double alpha = updx ? alpha_in[c] : 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double rv[KB_UNROLL], pv[KB_UNROLL], xv[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      pv[u] = p[KB_IDX(u)];
      if (updp) rv[u] = r[KB_IDX(u)];
      if (updx) xv[u] = x[KB_IDX(u)];
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      if (updx) x[KB_IDX(u)] = kb_mul_add(alpha, pv[u], xv[u]);
      if (updp) p[KB_IDX(u)] = kb_mul_add(omega, pv[u], rv[u]);
    }
  } else {
    for (int u = 0; u < KB_UNROLL; ++u) {
      const int64_t i = KB_IDX(u);
      if (i < total) {
        const double pv = p[i];
        if (updx) x[i] = kb_mul_add(alpha, pv, x[i]);
        if (updp) p[i] = kb_mul_add(omega, pv, r[i]);
      }
    }
  }
  KB_TILE_LOOP_END
}

// -------------------------------------------------------- generic axpy ---
// y += sign * coef[c] * x     24 B/element
__global__ void __launch_bounds__(KB_BLOCK)
kb_axpy_kernel(int64_t total, int k, double sign, const double* __restrict__ coef,
               const double* __restrict__ x, double* __restrict__ y, KbRed rd) {
  if (kb_gated(rd)) return;
  const double a = sign * coef[threadIdx.x % k];
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double xv[KB_UNROLL], yv[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      xv[u] = x[KB_IDX(u)];
      yv[u] = y[KB_IDX(u)];
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) y[KB_IDX(u)] = kb_mul_add(a, xv[u], yv[u]);
  } else {
    for (int u = 0; u < KB_UNROLL; ++u)
      if (KB_IDX(u) < total) y[KB_IDX(u)] = kb_mul_add(a, x[KB_IDX(u)], y[KB_IDX(u)]);
  }
  KB_TILE_LOOP_END
}

// out = ca[c] * x + cb[c] * y   (each product rounded, then the sum: NumPy temporaries).
// ca == nullptr: the x term is x itself; cb == nullptr: no y term (out = ca * x, a scaling).
// out may alias x or y.  Statements of the short-recurrence solvers (qmr.py:141-146,
// bicgstab.py:100,110,113,132-133, cgs.py:92-93,100).            16-24 B/element
__global__ void __launch_bounds__(KB_BLOCK)
kb_lincomb_kernel(int64_t total, int k, const double* __restrict__ ca, const double* x,
                  const double* __restrict__ cb, const double* y, double* out, KbRed rd) {
  if (kb_gated(rd)) return;
  const int c = threadIdx.x % k;
  const bool hx = ca != nullptr, hy = cb != nullptr;
  const double a = hx ? ca[c] : 1.0;
  const double b = hy ? cb[c] : 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double xv[KB_UNROLL], yv[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      xv[u] = x[KB_IDX(u)];
      yv[u] = hy ? y[KB_IDX(u)] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      const double t = hx ? __dmul_rn(a, xv[u]) : xv[u];
      out[KB_IDX(u)] = hy ? __dadd_rn(t, __dmul_rn(b, yv[u])) : t;
    }
  } else {
    for (int u = 0; u < KB_UNROLL; ++u) {
      const int64_t i = KB_IDX(u);
      if (i < total) {
        const double t = hx ? __dmul_rn(a, x[i]) : x[i];
        out[i] = hy ? __dadd_rn(t, __dmul_rn(b, y[i])) : t;
      }
    }
  }
  KB_TILE_LOOP_END
}

// y = x + coef[c] * y         24 B/element
__global__ void __launch_bounds__(KB_BLOCK)
kb_xpby_kernel(int64_t total, int k, const double* __restrict__ x,
               const double* __restrict__ coef, double* __restrict__ y, KbRed rd) {
  if (kb_gated(rd)) return;
  const double a = coef[threadIdx.x % k];
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double xv[KB_UNROLL], yv[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      xv[u] = x[KB_IDX(u)];
      yv[u] = y[KB_IDX(u)];
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) y[KB_IDX(u)] = kb_mul_add(a, yv[u], xv[u]);
  } else {
    for (int u = 0; u < KB_UNROLL; ++u)
      if (KB_IDX(u) < total) y[KB_IDX(u)] = kb_mul_add(a, y[KB_IDX(u)], x[KB_IDX(u)]);
  }
  KB_TILE_LOOP_END
}

// out = x / nz(coef[c])       16 B/element   (arnoldi.py:191-196, 274-277)
__global__ void __launch_bounds__(KB_BLOCK)
kb_div_scale_kernel(int64_t total, int k, const double* __restrict__ x,
                    const double* __restrict__ coef, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  const double d = kb_nz(coef[threadIdx.x % k]);
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double xv[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) xv[u] = x[KB_IDX(u)];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) out[KB_IDX(u)] = xv[u] / d;
  } else {
    for (int u = 0; u < KB_UNROLL; ++u)
      if (KB_IDX(u) < total) out[KB_IDX(u)] = x[KB_IDX(u)] / d;
  }
  KB_TILE_LOOP_END
}

// out = x + y                 24 B/element
__global__ void __launch_bounds__(KB_BLOCK)
kb_add_kernel(int64_t total, const double* __restrict__ x, const double* __restrict__ y,
              double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double xv[KB_UNROLL], yv[KB_UNROLL];
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) {
      xv[u] = x[KB_IDX(u)];
      yv[u] = y[KB_IDX(u)];
    }
#pragma unroll
    for (int u = 0; u < KB_UNROLL; ++u) out[KB_IDX(u)] = xv[u] + yv[u];
  } else {
    for (int u = 0; u < KB_UNROLL; ++u)
      if (KB_IDX(u) < total) out[KB_IDX(u)] = x[KB_IDX(u)] + y[KB_IDX(u)];
  }
  KB_TILE_LOOP_END
}

// ------------------------------------------------ MGS / Lanczos: axpy+dot -
// w -= (scale * coef[c]) * u; then out[c] = <z, w> (dot 1) | <w, w> (dot 2) | nothing
// 32 B/element with dot 1, 24 otherwise.   arnoldi.py:157-162, 264-267;
// with scale = beta it is a Householder reflector application (householder.py:62)
template <int DOT>
__global__ void __launch_bounds__(KB_BLOCK)
kb_axpy_dot_kernel(int64_t total, int k, const double* __restrict__ coef,
                   const double* __restrict__ scale, const double* __restrict__ u,
                   double* __restrict__ w, const double* __restrict__ z,
                   double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  double a = coef[threadIdx.x % k];
  if (scale != nullptr) a = __dmul_rn(scale[0], a);
  double acc = 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double uv[KB_UNROLL], wv[KB_UNROLL], zv[KB_UNROLL];
#pragma unroll
    for (int q = 0; q < KB_UNROLL; ++q) {
      uv[q] = u[KB_IDX(q)];
      wv[q] = w[KB_IDX(q)];
      if (DOT == 1) zv[q] = z[KB_IDX(q)];
    }
#pragma unroll
    for (int q = 0; q < KB_UNROLL; ++q) {
      const double wn = kb_mul_sub(a, uv[q], wv[q]);
      w[KB_IDX(q)] = wn;
      if (DOT == 1) acc = fma(zv[q], wn, acc);
      if (DOT == 2) acc = fma(wn, wn, acc);
    }
  } else {
    for (int q = 0; q < KB_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      if (i < total) {
        const double wn = kb_mul_sub(a, u[i], w[i]);
        w[i] = wn;
        if (DOT == 1) acc = fma(z[i], wn, acc);
        if (DOT == 2) acc = fma(wn, wn, acc);
      }
    }
  }
  KB_TILE_LOOP_END
  if (DOT != 0) kb_grid_colsum(acc, k, rd, out, sm);
}

// Lanczos alpha-step of MINRES with the scalar recurrences in the reduction's finishing block:
// Av -= alpha v, beta^2 = <Av, Av> (arnoldi.py:264-267), then -- in the block that holds the
// finished sum -- the tridiagonal QR update, residual norm and stopping test of minres.py:190-228
// (kb_minres_scalar_body).  Same arithmetic as kb_axpy_dot (dot 2) + kb_minres_scalar, one launch.
__global__ void __launch_bounds__(KB_BLOCK)
kb_axpy_dot_minres_kernel(int64_t total, int k, const double* __restrict__ coef,
                          const double* __restrict__ u, double* __restrict__ w,
                          double* __restrict__ out, int iter, kb_minres_state st, KbRed rd) {
  kb_pdl_prologue();
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  const double a = coef[threadIdx.x % k];
  double acc = 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double uv[KB_UNROLL], wv[KB_UNROLL];
#pragma unroll
    for (int q = 0; q < KB_UNROLL; ++q) {
      uv[q] = u[KB_IDX(q)];
      wv[q] = w[KB_IDX(q)];
    }
#pragma unroll
    for (int q = 0; q < KB_UNROLL; ++q) {
      const double wn = kb_mul_sub(a, uv[q], wv[q]);
      w[KB_IDX(q)] = wn;
      acc = fma(wn, wn, acc);
    }
  } else {
    for (int q = 0; q < KB_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      if (i < total) {
        const double wn = kb_mul_sub(a, u[i], w[i]);
        w[i] = wn;
        acc = fma(wn, wn, acc);
      }
    }
  }
  KB_TILE_LOOP_END
  const bool last = kb_grid_colsum(acc, k, rd, out, sm);
  if (last) {  // block-uniform
    __syncthreads();  // out[] (== st.ww) written by threads t < k of this block
    kb_minres_scalar_body(k, iter, st);
  }
}

// Last projection of an Arnoldi-MGS step with the Hessenberg update in the reduction's finishing
// block: w -= h V[j], <w, w> (arnoldi.py:157-162, 184-185), then gmres.py:199-221
// (kb_gmres_scalar_body).  Same arithmetic as kb_axpy_dot (dot 2) + kb_gmres_scalar, one launch.
__global__ void __launch_bounds__(KB_BLOCK)
kb_axpy_dot_gmres_kernel(int64_t total, int k, const double* __restrict__ coef,
                          const double* __restrict__ u, double* __restrict__ w,
                          double* __restrict__ out, int iter, kb_gmres_state st, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  const double a = coef[threadIdx.x % k];
  double acc = 0.0;
  KB_TILE_LOOP_BEGIN(total)
  if (kb_full) {
    double uv[KB_UNROLL], wv[KB_UNROLL];
#pragma unroll
    for (int q = 0; q < KB_UNROLL; ++q) {
      uv[q] = u[KB_IDX(q)];
      wv[q] = w[KB_IDX(q)];
    }
#pragma unroll
    for (int q = 0; q < KB_UNROLL; ++q) {
      const double wn = kb_mul_sub(a, uv[q], wv[q]);
      w[KB_IDX(q)] = wn;
      acc = fma(wn, wn, acc);
    }
  } else {
    for (int q = 0; q < KB_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      if (i < total) {
        const double wn = kb_mul_sub(a, u[i], w[i]);
        w[i] = wn;
        acc = fma(wn, wn, acc);
      }
    }
  }
  KB_TILE_LOOP_END
  const bool last = kb_grid_colsum(acc, k, rd, out, sm);
  if (last) {  // block-uniform
    __syncthreads();  // out[] (== st.ww) written by threads t < k of this block
    kb_gmres_scalar_body(k, iter, st);
  }
}

// --------------------------------------------------------- MINRES update -
// coefs = [R0 | R1 | R2 | y0 | h2] (k each)
// z = (v - R0 W0 - R1 W1) / nz(R2);  W0 <- z;  yk += y0 z;  vnext = Av / nz(h2)
// 64 B/element (reads v W0 W1 Av yk, writes W0 yk vnext).
// With a preconditioner M (MAv != NULL): vnext = MAv / nz(h2), pnext = Av / nz(h2)
// (the V = M P pair of bases, arnoldi.py:274-277): 80 B/element.
// minres.py:219-221 ("take the longest"), arnoldi.py:274-277
#define KB_MR_UNROLL 2
template <bool PRE>
__global__ void __launch_bounds__(KB_BLOCK, 4)
kb_minres_update_kernel(int64_t total, int k, const double* __restrict__ coefs,
                        const double* __restrict__ v, double* __restrict__ W0,
                        const double* __restrict__ W1, const double* __restrict__ Av,
                        double* __restrict__ yk, double* __restrict__ vnext,
                        const double* __restrict__ MAv, double* __restrict__ pnext, KbRed rd) {
  kb_pdl_prologue();
  if (kb_gated(rd)) return;
  const int c = threadIdx.x % k;
  const double R0 = coefs[c];
  const double R1 = coefs[k + c];
  const double R2 = kb_nz(coefs[2 * k + c]);
  const double y0 = coefs[3 * k + c];
  const double h2 = kb_nz(coefs[4 * k + c]);
  constexpr bool pre = PRE;
  KB_TILE_LOOP_BEGIN_U(total, KB_MR_UNROLL)
  if (kb_full) {
    double vv[KB_MR_UNROLL], w0[KB_MR_UNROLL], w1[KB_MR_UNROLL], av[KB_MR_UNROLL],
        yv[KB_MR_UNROLL], mv[KB_MR_UNROLL];
#pragma unroll
    for (int q = 0; q < KB_MR_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      vv[q] = v[i];
      w0[q] = W0[i];
      w1[q] = W1[i];
      av[q] = Av[i];
      yv[q] = yk[i];
      mv[q] = pre ? MAv[i] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < KB_MR_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      const double zz = kb_mul_sub(R1, w1[q], kb_mul_sub(R0, w0[q], vv[q])) / R2;
      W0[i] = zz;
      yk[i] = kb_mul_add(y0, zz, yv[q]);
      if (pre) {
        vnext[i] = mv[q] / h2;
        pnext[i] = av[q] / h2;
      } else {
        vnext[i] = av[q] / h2;
      }
    }
  } else {
    for (int q = 0; q < KB_MR_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      if (i < total) {
        const double zz = kb_mul_sub(R1, W1[i], kb_mul_sub(R0, W0[i], v[i])) / R2;
        const double a = Av[i];
        W0[i] = zz;
        yk[i] = kb_mul_add(y0, zz, yk[i]);
        if (pre) {
          vnext[i] = MAv[i] / h2;
          pnext[i] = a / h2;
        } else {
          vnext[i] = a / h2;
        }
      }
    }
  }
  KB_TILE_LOOP_END
}

// ------------------------------------------------------- basis combine ---
// out = x0 + sum_{j<m} yy[j,c] * V[j]      8(m+2) B/element
// gmres.py:96-98 (Python `sum` of scaled vectors, left to right from 0)
__global__ void __launch_bounds__(KB_BLOCK)
kb_basis_combine_kernel(int64_t total, int k, int m, const double* __restrict__ yy,
                        const double* __restrict__ Vbuf, int64_t vstride,
                        const double* __restrict__ x0, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  const int c = threadIdx.x % k;
  KB_TILE_LOOP_BEGIN(total)
  (void)kb_full;
  for (int u = 0; u < KB_UNROLL; ++u) {
    const int64_t e = KB_IDX(u);
    if (e >= total) break;
    double acc = 0.0;
    int j = 0;
    for (; j + 3 < m; j += 4) {
      const double v0 = Vbuf[(size_t)j * vstride + e];
      const double v1 = Vbuf[(size_t)(j + 1) * vstride + e];
      const double v2 = Vbuf[(size_t)(j + 2) * vstride + e];
      const double v3 = Vbuf[(size_t)(j + 3) * vstride + e];
      acc = kb_mul_add(yy[(size_t)j * k + c], v0, acc);
      acc = kb_mul_add(yy[(size_t)(j + 1) * k + c], v1, acc);
      acc = kb_mul_add(yy[(size_t)(j + 2) * k + c], v2, acc);
      acc = kb_mul_add(yy[(size_t)(j + 3) * k + c], v3, acc);
    }
    for (; j < m; ++j) acc = kb_mul_add(yy[(size_t)j * k + c], Vbuf[(size_t)j * vstride + e], acc);
    out[e] = x0[e] + acc;
  }
  KB_TILE_LOOP_END
}

// ------------------------------------- classical Gram-Schmidt (extension) -
// Tall-skinny V^T w: out[jj, c] = <V[j0 + jj][:, c], w[:, c]>, jj < cnt <= JC, in ONE pass
// over w: (cnt + 1) * 8 B/element instead of the 16-32 of cnt separate dots.  Not in the
// reference (its Arnoldi is MGS only, arnoldi.py:157-162): ortho="cgs"/"cgs<N>" is additive.
#define KB_MD_UNROLL 2
template <int JC>
__global__ void __launch_bounds__(KB_BLOCK, (JC > 8 ? 2 : 4))
kb_multi_dot_kernel(int64_t total, int k, int cnt, const double* __restrict__ V, int64_t vstride,
                    const double* __restrict__ w, double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  double acc[JC];
#pragma unroll
  for (int jj = 0; jj < JC; ++jj) acc[jj] = 0.0;
  KB_TILE_LOOP_BEGIN_U(total, KB_MD_UNROLL)
  if (kb_full) {
    double wv[KB_MD_UNROLL];
#pragma unroll
    for (int q = 0; q < KB_MD_UNROLL; ++q) wv[q] = w[KB_IDX(q)];
#pragma unroll
    for (int jj = 0; jj < JC; ++jj) {
      if (jj < cnt) {
        const double* __restrict__ vj = V + (size_t)jj * vstride;
#pragma unroll
        for (int q = 0; q < KB_MD_UNROLL; ++q) acc[jj] = fma(vj[KB_IDX(q)], wv[q], acc[jj]);
      }
    }
  } else {
    for (int q = 0; q < KB_MD_UNROLL; ++q) {
      const int64_t i = KB_IDX(q);
      if (i < total) {
        const double wi = w[i];
#pragma unroll
        for (int jj = 0; jj < JC; ++jj)
          if (jj < cnt) acc[jj] = fma(V[(size_t)jj * vstride + i], wi, acc[jj]);
      }
    }
  }
  KB_TILE_LOOP_END
  kb_grid_multisum<JC>(acc, cnt, k, rd, out, sm);
}

// w -= sum_{j < m} h[j, c] * P[j]   (+ out = <w, w> with DOT 2): one pass, (m + 2) * 8 B/element.
template <int DOT>
__global__ void __launch_bounds__(KB_BLOCK, 4)
kb_multi_axpy_kernel(int64_t total, int k, int m, const double* __restrict__ h,
                     const double* __restrict__ P, int64_t pstride, double* __restrict__ w,
                     double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  __shared__ double sm[KB_BLOCK];
  const int c = threadIdx.x % k;
  double acc = 0.0;
  KB_TILE_LOOP_BEGIN_U(total, KB_MD_UNROLL)
  double wv[KB_MD_UNROLL];
  bool ok[KB_MD_UNROLL];
#pragma unroll
  for (int q = 0; q < KB_MD_UNROLL; ++q) {
    ok[q] = kb_full || KB_IDX(q) < total;
    wv[q] = ok[q] ? w[KB_IDX(q)] : 0.0;
  }
  int j = 0;
  for (; j + 7 < m; j += 8) {
    double pv[8][KB_MD_UNROLL];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
#pragma unroll
      for (int q = 0; q < KB_MD_UNROLL; ++q)
        pv[jj][q] = ok[q] ? P[(size_t)(j + jj) * pstride + KB_IDX(q)] : 0.0;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const double hj = __ldg(&h[(size_t)(j + jj) * k + c]);
#pragma unroll
      for (int q = 0; q < KB_MD_UNROLL; ++q) wv[q] = kb_mul_sub(hj, pv[jj][q], wv[q]);
    }
  }
  for (; j < m; ++j) {
    const double hj = __ldg(&h[(size_t)j * k + c]);
#pragma unroll
    for (int q = 0; q < KB_MD_UNROLL; ++q)
      if (ok[q]) wv[q] = kb_mul_sub(hj, P[(size_t)j * pstride + KB_IDX(q)], wv[q]);
  }
#pragma unroll
  for (int q = 0; q < KB_MD_UNROLL; ++q) {
    if (ok[q]) {
      w[KB_IDX(q)] = wv[q];
      if (DOT == 2) acc = fma(wv[q], wv[q], acc);
    }
  }
  KB_TILE_LOOP_END
  if (DOT != 0) kb_grid_colsum(acc, k, rd, out, sm);
}

// gather rows: buf[i, :] = x[idx[i], :]   (halo send buffer)
__global__ void __launch_bounds__(KB_BLOCK)
kb_pack_rows_kernel(int64_t n_idx, int k, const int32_t* __restrict__ idx,
                    const double* __restrict__ x, double* __restrict__ buf, KbRed rd) {
  if (kb_gated(rd)) return;
  const int64_t total = n_idx * k;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / k;
    const int c = (int)(e - i * k);
    buf[e] = x[(size_t)idx[i] * k + c];
  }
}

// single-element edits (Householder Arnoldi bookkeeping, arnoldi.py:75-96)
__global__ void kb_poke_kernel(int op, double* x, int64_t idx, const double* s, double val,
                               double* dst, KbRed rd) {
  if (kb_gated(rd)) return;
  if (op == 0) x[idx] *= s[0];
  else if (op == 1) x[idx] = val;
  else if (op == 2) dst[0] = x[idx];
}

// stand-alone fused all-reduce of a k-slot (a rank without boundary rows still
// has to take part in the collective that its peers run inside kb_spmv_halo_add)
__global__ void kb_allreduce_kernel(int k, double* slot, KbRed rd) {
  if (kb_gated(rd)) return;
  const int t = threadIdx.x;
  double v = (t < k) ? slot[t] : 0.0;
  if (rd.cm.size > 1) {
    v = kb_p2p_allreduce(v, k, rd.cm);
    if (t < k && *rd.cm.error) v = nan("");
  }
  if (t < k) slot[t] = v;
}

// ---------------------------------------------------------- halo push -------
// segs: n_seg x 4 int64 (device): destination rank, first entry in idx, entries, first row
// in the destination's data area.  One launch per product on every rank (also with
// n_seg == 0, so that the product counters stay equal on all ranks).
__global__ void __launch_bounds__(KB_BLOCK)
kb_halo_push_kernel(int k, int n_seg, const int64_t* __restrict__ segs, int64_t n_total,
                    const int32_t* __restrict__ idx, const double* __restrict__ x, KbHalo hd,
                    KbRed rd) {
  if (kb_gated(rd)) return;
  unsigned char* own = hd.peers[hd.rank];
  const unsigned long long q = *kb_halo_u64(own, KB_HALO_COUNTER) + 1ull;
  // flow control: destinations must have consumed product q-1 before their buffer is reused
  if ((int)threadIdx.x < n_seg)
    kb_halo_wait(kb_halo_u64(own, KB_HALO_ACKS + 8 * (size_t)segs[4 * threadIdx.x]), q - 1ull, own);
  __syncthreads();
  // A destination that never acknowledged (sticky error word) may still be reading its buffer:
  // nothing is pushed and no flag is raised any more -- the destinations time out in turn and
  // poison their products with NaN, so the failure surfaces on every rank.
  const bool dead = *reinterpret_cast<volatile int*>(own + KB_HALO_ERROR) != 0;
  const int64_t total = dead ? 0 : n_total * k;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / k;
    const int c = (int)(e - i * k);
    int s = 0;
    while (s + 1 < n_seg && i >= segs[4 * (s + 1) + 1]) ++s;
    double* dst = reinterpret_cast<double*>(hd.peers[segs[4 * s]] + KB_HALO_DATA);
    dst[(size_t)(segs[4 * s + 3] + (i - segs[4 * s + 1])) * k + c] = x[(size_t)idx[i] * k + c];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(own + KB_HALO_PUSH_TICKET);
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {
    __threadfence_system();
    if ((int)threadIdx.x < n_seg && !dead)
      *kb_halo_u64(hd.peers[segs[4 * threadIdx.x]], KB_HALO_FLAGS + 8 * (size_t)hd.rank) = q;
    if (threadIdx.x == 0) {
      *kb_halo_u64(own, KB_HALO_COUNTER) = q;
      *ticket = 0u;
    }
  }
}
