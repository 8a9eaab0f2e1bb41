// Whole-batch entry points of MINRES and GMRES ("kb_minres_solve",
// "kb_gmres_cycle"): one C call enqueues every launch of iterations i0 .. i0+n-1, each gated on
// the device-resident stop flag, so the host language is out of the loop between two read-backs
// (minres.py:168-236, gmres.py:179-234 with arnoldi.py:153-200).  Same kernels, same order and
// same arguments as the per-launch path of krylov_b200/minres.py and gmres.py.
#include "kb_handles.cuh"

// Per-column scalar arithmetic on device slots: what the reference does with NumPy (k,) arrays
// between two vector statements of the short-recurrence solvers (bicgstab.py:100-133,
// qmr.py:101-146, symmlq.py:108-150, ...).  One IEEE operation per launch, so the values equal
// the host's bit for bit; keeping them on the device removes the read-back per inner product.
__global__ void kb_scalar_op_kernel(int k, int op, const double* __restrict__ a,
                                    const double* __restrict__ b, double sa, double sb,
                                    double* __restrict__ out, KbRed rd) {
  if (kb_gated(rd)) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  const double A = a != nullptr ? a[c] : sa;
  const double B = b != nullptr ? b[c] : sb;
  double r;
  switch (op) {
    case 0: r = __dadd_rn(A, B); break;
    case 1: r = __dsub_rn(A, B); break;
    case 2: r = __dmul_rn(A, B); break;
    case 3: r = __ddiv_rn(A, B); break;
    case 4: r = __dsqrt_rn(A); break;
    case 5: r = fabs(A); break;
    case 6: r = -A; break;
    case 7: r = A != 0.0 ? A : B; break;  // nz(A) with B as the replacement
    case 9: r = __ddiv_rn(A, B != 0.0 ? B : (b != nullptr ? sb : 1.0)); break;  // A / nz(B): the
                                                                           // fill travels in sb
    default: r = A; break;
  }
  out[c] = r;
}

// hist[step * k + c] = val[c]; all columns val <= crit -> *stop_at = step.  The stopping test of
// the short-recurrence drivers on the device (bicgstab.py:96-99, cgs.py:88-91, ...: the loop
// condition `resnorms[-1] > criterion`), so that iterations can be enqueued ahead of the read-back.
__global__ void __launch_bounds__(KB_MAX_K)
kb_record_kernel(int k, int step, const double* __restrict__ val, const double* __restrict__ crit,
                 double* __restrict__ hist, int* __restrict__ stop_at, KbRed rd) {
  if (kb_gated(rd)) return;
  int ok = 1;
  if (threadIdx.x < k) {
    const double v = val[threadIdx.x];
    hist[(size_t)step * k + threadIdx.x] = v;
    ok = (v <= crit[threadIdx.x]) ? 1 : 0;
  }
  const int all_ok = __syncthreads_and(ok);
  if (all_ok && threadIdx.x == 0) *stop_at = step;
}

extern "C" {

int kb_record(kb_ws_t ws, int k, int step, const double* val, const double* crit, double* hist,
              int* stop_at, void* stream) {
  KB_REQUIRE(ws != nullptr && val && crit && hist && stop_at, "null argument");
  KB_REQUIRE(k >= 1 && k <= KB_MAX_K, "k out of range");
  kb_record_kernel<<<1, KB_MAX_K, 0, (cudaStream_t)stream>>>(k, step, val, crit, hist, stop_at,
                                                             kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_scalar_op(kb_ws_t ws, int k, int op, const double* a, const double* b, double sa, double sb,
                 double* out, void* stream) {
  KB_REQUIRE(ws != nullptr && out != nullptr, "null argument");
  KB_REQUIRE(k >= 1 && k <= KB_MAX_K, "k out of range");
  KB_REQUIRE(op >= 0 && op <= 9, "unknown operation");
  KB_REQUIRE(op != 9 || b != nullptr, "op 9 (A / nz(B)) needs B as a device array");
  kb_scalar_op_kernel<<<(k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(k, op, a, b, sa, sb, out,
                                                                         kb_red(ws));
  KB_LAUNCH_CHECK();
  return KB_OK;
}

int kb_minres_run(kb_ws_t ws, const kb_minres_run_state* s, int i0, int n_iters, void* stream) {
  KB_REQUIRE(ws != nullptr && s != nullptr, "null argument");
  KB_REQUIRE(s->A && s->V[0] && s->V[1] && s->W[0] && s->W[1] && s->Av && s->yk,
             "null field in kb_minres_run_state");
  KB_REQUIRE(i0 >= 0 && n_iters >= 0, "negative iteration range");
  const int k = s->k;
  const kb_minres_state* st = &s->st;
  const int* saved_gate = ws->gate;
  const int saved_tag = ws->gate_tag;
  int rc = KB_OK;
  for (int i = i0; i < i0 + n_iters && rc == KB_OK; ++i) {
    double* v = s->V[i % 2];
    double* vold = s->V[(i + 1) % 2];
    ws->gate = st->stop_at;
    ws->gate_tag = i;
    // Av = A v - beta_{i-1} v_old, alpha = <v, Av>   (arnoldi.py:244-252)
    if (i == 0)
      rc = kb_spmv(s->A, ws, k, v, s->Av, 0, nullptr, nullptr, 1, v, (double*)st->alpha, stream);
    else
      rc = kb_spmv(s->A, ws, k, v, s->Av, 1, vold, st->h2prev, 1, v, (double*)st->alpha, stream);
    // Av -= alpha v, beta_i^2 = <Av, Av> (arnoldi.py:264-267); the scalar recurrences of
    // minres.py:190-228 in the finishing block of the same launch
    if (rc == KB_OK) rc = kb_axpy_dot_minres(ws, s->n, k, st->alpha, v, s->Av, i, st, stream);
    if (rc == KB_OK)  // minres.py:219-221, arnoldi.py:274-277
      rc = kb_minres_update(ws, s->n, k, st->coefs, v, s->W[i % 2], s->W[(i + 1) % 2], s->Av, s->yk,
                            vold, nullptr, nullptr, stream);
  }
  ws->gate = saved_gate;
  ws->gate_tag = saved_tag;
  return rc;
}

int kb_gmres_cycle(kb_ws_t ws, const kb_gmres_cycle_state* s, int i0, int n_iters, void* stream) {
  KB_REQUIRE(ws != nullptr && s != nullptr, "null argument");
  KB_REQUIRE(s->A && s->Vbuf && s->w && s->dots && s->ww && s->hlast,
             "null field in kb_gmres_cycle_state");
  KB_REQUIRE(s->st.ww == s->ww && s->st.have_h == 0, "st.ww must be the ww slot (MGS path)");
  KB_REQUIRE(i0 >= 0 && n_iters >= 0, "negative iteration range");
  KB_REQUIRE(i0 + n_iters <= s->st.maxiter, "basis storage too small for this range");
  const int k = s->k;
  const int nre = s->st.num_reorthos;
  const int* saved_gate = ws->gate;
  const int saved_tag = ws->gate_tag;
  int rc = KB_OK;
  for (int i = i0; i < i0 + n_iters && rc == KB_OK; ++i) {
    ws->gate = s->st.stop_at;
    ws->gate_tag = i;
    double* V0 = s->Vbuf;
    double* Vi = s->Vbuf + (size_t)i * s->vstride;
    // w = A V[i], first MGS coefficient <V[0], w> in the same pass
    rc = kb_spmv(s->A, ws, k, Vi, s->w, 0, nullptr, nullptr, 1, V0, s->dots, stream);
    int idx = 0;
    for (int sweep = 0; sweep < nre && rc == KB_OK; ++sweep)
      for (int j = 0; j <= i && rc == KB_OK; ++j, ++idx) {  // arnoldi.py:157-162
        double* Vj = s->Vbuf + (size_t)j * s->vstride;
        double* coef = s->dots + (size_t)idx * k;
        if (sweep == nre - 1 && j == i) {
          // last projection + <w, w> (arnoldi.py:184-185); the Hessenberg / Givens update of
          // gmres.py:206-221 in the finishing block of the same launch
          rc = kb_axpy_dot_gmres(ws, s->n, k, coef, Vj, s->w, i, &s->st, stream);
        } else {
          const double* nxt = j < i ? Vj + s->vstride : V0;
          rc = kb_axpy_dot(ws, s->n, k, coef, nullptr, Vj, s->w, 1, nxt, coef + k, stream);
        }
      }
    if (rc == KB_OK)  // V[i+1] = w / h[i+1]   (arnoldi.py:191-193)
      rc = kb_div_scale(ws, s->n, k, s->w, s->hlast, Vi + s->vstride, stream);
  }
  ws->gate = saved_gate;
  ws->gate_tag = saved_tag;
  return rc;
}

}  // extern "C"
