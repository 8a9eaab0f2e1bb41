// Small per-column recurrences that stay on the device so an iteration never
// needs a host round-trip: Givens rotations (LAPACK dlartg semantics), the
// MINRES tridiagonal QR, the GMRES Hessenberg QR, the small triangular solve.
// One block per launch, one thread per right-hand-side column.
#pragma once
#include <float.h>

#include "kb_common.cuh"

// LAPACK 3.10 dlartg (la_lartg.f90), the routine behind givens.py:35-38.
// [ c  s ] [f]   [r]
// [-s  c ] [g] = [0]
__device__ __forceinline__ void kb_dlartg(double f, double g, double& c, double& s, double& r) {
  const double safmin = DBL_MIN;
  const double safmax = 1.0 / safmin;
  const double rtmin = sqrt(safmin);
  const double rtmax = sqrt(safmax / 2.0);
  const double f1 = fabs(f), g1 = fabs(g);
  if (g == 0.0) {
    c = 1.0;
    s = 0.0;
    r = f;
  } else if (f == 0.0) {
    c = 0.0;
    s = copysign(1.0, g);
    r = g1;
  } else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
    const double d = sqrt(__dadd_rn(__dmul_rn(f, f), __dmul_rn(g, g)));
    c = f1 / d;
    r = copysign(d, f);
    s = g / r;
  } else {
    const double u = fmin(safmax, fmax(safmin, fmax(f1, g1)));
    const double fs = f / u, gs = g / u;
    const double d = sqrt(__dadd_rn(__dmul_rn(fs, fs), __dmul_rn(gs, gs)));
    c = fabs(fs) / d;
    r = copysign(d, f);
    s = gs / r;
    r = r * u;
  }
}

// 2x2 rotation applied to (a, b): einsum("ij,j->i") of minres.py:23-25 / gmres.py:19-21
__device__ __forceinline__ void kb_rot(double c, double s, double& a, double& b) {
  const double na = __dadd_rn(__dmul_rn(c, a), __dmul_rn(s, b));
  const double nb = __dadd_rn(__dmul_rn(-s, a), __dmul_rn(c, b));
  a = na;
  b = nb;
}

// test harness entry: out[3*i..] = (c, s, r) of (f[i], g[i])
__global__ void kb_lartg_kernel(int n, const double* f, const double* g, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double c, s, r;
    kb_dlartg(f[i], g[i], c, s, r);
    out[3 * i] = c;
    out[3 * i + 1] = s;
    out[3 * i + 2] = r;
  }
}

// ---------------------------------------------------------------- MINRES --
// minres.py:190-228 for Lanczos step `iter` (0-based); thread c = column.  Block-wide: every
// thread of ONE block (>= k threads) calls it -- the scalar kernel below, or the finishing block
// of the reduction that produced st.ww (kb_axpy_dot_minres_kernel: one launch less per step).
__device__ __forceinline__ void kb_minres_scalar_body(int k, int iter, const kb_minres_state& st) {
  const int c = threadIdx.x;
  int conv = 1, inv = 1;
  if (c < k) {
    // h = [beta_{k-1}, alpha_k, beta_k]   (arnoldi.py:246-267)
    const double h0 = (iter > 0) ? st.h2prev[c] : 0.0;
    const double h1 = st.alpha[c];
    const double h2 = sqrt(st.ww[c]);
    st.h2prev[c] = h2;
    inv = (h2 <= 1.0e-14) ? 1 : 0;  // arnoldi.py:269
    double R0 = 0.0, R1 = h0, R2, R3;
    if (iter >= 2) kb_rot(st.g1[2 * c], st.g1[2 * c + 1], R0, R1);  // minres.py:197-202
    R2 = h1;
    R3 = h2;
    const double c0 = st.g0[2 * c], s0 = st.g0[2 * c + 1];
    if (iter >= 1) kb_rot(c0, s0, R1, R2);  // minres.py:207-208
    st.g1[2 * c] = c0;
    st.g1[2 * c + 1] = s0;
    double cn, sn, rn;
    kb_dlartg(R2, R3, cn, sn, rn);  // minres.py:211
    st.g0[2 * c] = cn;
    st.g0[2 * c + 1] = sn;
    R2 = rn;
    double ya = st.y0[c], yb = 0.0;
    kb_rot(cn, sn, ya, yb);  // minres.py:215
    st.coefs[c] = R0;
    st.coefs[k + c] = R1;
    st.coefs[2 * k + c] = R2;
    st.coefs[3 * k + c] = ya;
    st.coefs[4 * k + c] = h2;
    st.y0[c] = yb;  // minres.py:224
    const double rn_abs = fabs(yb);
    st.hist[(size_t)(iter + 1) * k + c] = rn_abs;
    conv = (rn_abs <= st.crit[c]) ? 1 : 0;
  }
  const int all_conv = __syncthreads_and(conv);
  const int all_inv = __syncthreads_and(inv);
  if (threadIdx.x == 0) {
    // an invariant subspace also ends the enqueued batch: the next Arnoldi/Lanczos
    // step would raise ArgumentError in the reference (arnoldi.py:168-171, 239-242)
    if (all_inv) atomicOr(st.flags, 1);
    if (all_conv || all_inv) *st.stop_at = iter + 1;
  }
}

__global__ void kb_minres_scalar_kernel(int k, int iter, kb_minres_state st, KbRed rd) {
  if (kb_gated(rd)) return;
  kb_minres_scalar_body(k, iter, st);
}

// ----------------------------------------------------------------- GMRES --
// gmres.py:199-221 for Arnoldi step `iter`; R is (maxiter+1, maxiter, k)
// row-major like the reference's array, Gc/Gs the rotation list, y the rhs.
// Block-wide like kb_minres_scalar_body: the scalar kernel below, or the finishing block of the
// reduction that produced st.ww (kb_axpy_dot_gmres_kernel).
__device__ __forceinline__ void kb_gmres_scalar_body(int k, int iter, const kb_gmres_state& st) {
  const int c = threadIdx.x;
  const int mi = st.maxiter;
  int conv = 1, inv = 1;
  if (c < k) {
    const int j = iter;
#define RIDX(row, col) (((size_t)(row) * mi + (col)) * k + c)
    double hl;
    if (st.have_h) {
      // Householder: dots already holds h[0..j+1] (arnoldi.py:84-85)
      for (int i = 0; i <= j + 1; ++i) st.R[RIDX(i, j)] = st.dots[(size_t)i * k + c];
      hl = st.dots[(size_t)(j + 1) * k + c];
    } else {
      // h[i] = sum over reorthogonalisation passes (arnoldi.py:160-161)
      for (int i = 0; i <= j; ++i) {
        double hv = 0.0;
        for (int p = 0; p < st.num_reorthos; ++p)
          hv += st.dots[((size_t)p * (j + 1) + i) * k + c];
        st.R[RIDX(i, j)] = hv;
      }
      hl = sqrt(st.ww[c]);  // arnoldi.py:185
      st.R[RIDX(j + 1, j)] = hl;
    }
    st.hlast[c] = hl;
    inv = (hl <= 1.0e-14) ? 1 : 0;  // arnoldi.py:187
    // previous rotations on the new column (gmres.py:209-210)
    for (int i = 0; i < j; ++i) {
      double a = st.R[RIDX(i, j)], b = st.R[RIDX(i + 1, j)];
      kb_rot(st.Gc[(size_t)i * k + c], st.Gs[(size_t)i * k + c], a, b);
      st.R[RIDX(i, j)] = a;
      st.R[RIDX(i + 1, j)] = b;
    }
    double cn, sn, rn;
    kb_dlartg(st.R[RIDX(j, j)], st.R[RIDX(j + 1, j)], cn, sn, rn);  // gmres.py:213
    st.Gc[(size_t)j * k + c] = cn;
    st.Gs[(size_t)j * k + c] = sn;
    st.R[RIDX(j, j)] = rn;
    st.R[RIDX(j + 1, j)] = 0.0;
    double ya = st.y[(size_t)j * k + c], yb = st.y[(size_t)(j + 1) * k + c];
    kb_rot(cn, sn, ya, yb);  // gmres.py:217
    st.y[(size_t)j * k + c] = ya;
    st.y[(size_t)(j + 1) * k + c] = yb;
    const double rn_abs = fabs(yb);
    st.hist[(size_t)(j + 1) * k + c] = rn_abs;
    conv = (rn_abs <= st.crit[c]) ? 1 : 0;
#undef RIDX
  }
  const int all_conv = __syncthreads_and(conv);
  const int all_inv = __syncthreads_and(inv);
  if (threadIdx.x == 0) {
    // an invariant subspace also ends the enqueued batch: the next Arnoldi/Lanczos
    // step would raise ArgumentError in the reference (arnoldi.py:168-171, 239-242)
    if (all_inv) atomicOr(st.flags, 1);
    if (all_conv || all_inv) *st.stop_at = iter + 1;
  }
}

__global__ void kb_gmres_scalar_kernel(int k, int iter, kb_gmres_state st, KbRed rd) {
  if (kb_gated(rd)) return;
  kb_gmres_scalar_body(k, iter, st);
}

// yy = R[:m,:m]^{-1} y[:m] per column; all-zero rhs column -> zeros
// (gmres.py:24-38; back substitution like LAPACK dtrtrs 'U','N','N')
__global__ void kb_gmres_solve_y_kernel(int k, int m, int maxiter, const double* __restrict__ R,
                                        const double* __restrict__ y, double* __restrict__ yy,
                                        KbRed rd) {
  if (kb_gated(rd)) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  bool allzero = true;
  for (int i = 0; i < m; ++i) {
    const double v = y[(size_t)i * k + c];
    yy[(size_t)i * k + c] = v;
    if (v != 0.0) allzero = false;
  }
  if (allzero) return;  // yy already zero
  for (int j = m - 1; j >= 0; --j) {
    const double bj = yy[(size_t)j * k + c] / R[((size_t)j * maxiter + j) * k + c];
    yy[(size_t)j * k + c] = bj;
    for (int i = 0; i < j; ++i)
      yy[(size_t)i * k + c] =
          kb_mul_sub(bj, R[((size_t)i * maxiter + j) * k + c], yy[(size_t)i * k + c]);
  }
}

// ------------------------------------------------------------ Householder --
// householder.py:26-51 for the tail x[off:], k == 1.
// scratch[0] = sigma2 = <x[off+1:], x[off+1:]> must already be reduced.
// params: [0] alpha, [1] beta, [2] xnorm, [3] v0 (unnormalised), [4] 1/||v||-divisor
// lapack_sign: a zero pivot with a nonzero tail counts as positive (Fortran SIGN(a, 0) = +|a| in
// dlarfg), i.e. H x = -||x|| e_1; the reference's Householder maps it to +||x|| e_1.
__global__ void kb_house_params_kernel(const double* x, int64_t off, const double* scratch,
                                       double* params, int lapack_sign, KbRed rd) {
  if (kb_gated(rd)) return;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double gamma = x[off];
  const double sigma2 = scratch[0];
  double v0 = 1.0, alpha, beta, xnorm = sqrt(__dadd_rn(__dmul_rn(fabs(gamma), fabs(gamma)), sigma2));
  if (sigma2 == 0.0) {
    beta = 0.0;
    xnorm = fabs(gamma);
    alpha = (gamma == 0.0) ? 1.0 : gamma / xnorm;
  } else {
    beta = 2.0;
    if (gamma == 0.0) {
      v0 = lapack_sign ? sqrt(sigma2) : -sqrt(sigma2);
      alpha = lapack_sign ? -1.0 : 1.0;
    } else {
      v0 = __dadd_rn(gamma, __dmul_rn(gamma / fabs(gamma), xnorm));
      alpha = -gamma / fabs(gamma);
    }
  }
  params[0] = alpha;
  params[1] = beta;
  params[2] = xnorm;
  params[3] = v0;
  params[4] = sqrt(__dadd_rn(__dmul_rn(fabs(v0), fabs(v0)), sigma2));  // householder.py:48
}

// v[i] = 0 (i < off); v[off] = v0 / d; v[i] = x[i] / d (i > off)
__global__ void __launch_bounds__(KB_BLOCK)
kb_house_fill_kernel(int64_t n, int64_t off, const double* __restrict__ x,
                     const double* __restrict__ params, double* __restrict__ v, KbRed rd) {
  if (kb_gated(rd)) return;
  const double v0 = params[3], d = params[4];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    double val;
    if (e < off) val = 0.0;
    else if (e == off) val = v0 / d;
    else val = x[e] / d;
    v[e] = val;
  }
}

// h[k+1] of the Householder Arnoldi step (arnoldi.py:83-85): first entry of
// (H w)[off:] * alpha, taken in absolute value; tau = <v, w> already reduced.
__global__ void kb_house_hlast_kernel(const double* w, int64_t off, const double* v,
                                      const double* params, const double* tau, double* h_out,
                                      KbRed rd) {
  if (kb_gated(rd)) return;
  if (threadIdx.x != 0) return;
  const double alpha = params[0], beta = params[1];
  double val = w[off];
  if (beta != 0.0) val = __dsub_rn(val, __dmul_rn(__dmul_rn(beta, v[off]), tau[0]));
  h_out[0] = fabs(__dmul_rn(val, alpha));
}
