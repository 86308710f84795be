// aa.cuh -- device-resident Anderson acceleration (replaces S/src/aa.c).
//
// The reference solves the regularised least squares (A'B + rI) gamma = A'g through a
// column-pivoted QR (dgeqp3) of the tall augmented matrix [A; sqrt(r) I] (aa.c:422-652) to
// avoid squaring the condition number.  The device version keeps that property without a
// Gram matrix: a two-stage Householder TSQR over the rows of [A | Y | g] (every CTA streams
// its row tiles through shared memory and keeps a running len x C trapezoid), a final CTA
// merges the per-CTA trapezoids with the sqrt(r) I rows, and one thread then runs the
// len x len pivoted QR / rank cut / LU solve / iterative refinement on the tiny factor.
// The ring buffers S, Y, D (dim x mem, column-major), the AA iteration counter, the
// accept/reject decisions and all statistics stay in HBM; the host never sees gamma.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kAaMaxMem = 31;  // lookback supported by the shared-memory TSQR (C = 2*mem+1 <= 63)

struct AaState {  // lives in device memory
  int iter, success, len, do_solve, sg_reject, pad;
  double norm_g, aa_norm;
  double nrm_s_col[kAaMaxMem + 1], nrm_y_col[kAaMaxMem + 1];
  double gamma[kAaMaxMem + 1];
  // lifetime diagnostics, aa_stats.h:21-42
  int n_accept, n_reject_lapack, n_reject_rank0, n_reject_nonfinite, n_reject_weight_cap, n_safeguard_reject,
      last_rank, pad2;
  double last_aa_norm, last_regularization;
};

struct AaDev {
  Ctx *c = nullptr;
  int dim = 0, mem = 0, min_len = 0, type1 = 1, ir_max_steps = 5;
  double regularization = 1e-8, relaxation = 1.0, safeguard_factor = 1.0, max_weight_norm = 1e10;
  double *x = nullptr, *f = nullptr, *g = nullptr, *g_prev = nullptr;
  double *Y = nullptr, *S = nullptr, *D = nullptr, *x_work = nullptr;
  double *Rpart = nullptr;  // [nblk][mem][C] per-CTA trapezoids
  int nblk = 0;
  // row-partitioned mode (dist.cuh): rows this rank counts, and the per-rank trapezoids that are gathered
  // (sum all-reduce of a buffer in which every rank fills its own slot) before the final merge
  int cnt_lo = 0, cnt_hi = 0;
  double *Rsend = nullptr, *Rrecv = nullptr;  // [world][mem][Cmax]
  int set_counted_rows(int lo, int hi);
  AaState *st = nullptr;       // device
  AaState *st_host = nullptr;  // pinned
  size_t smem1 = 0, smem2 = 0;

  int init(Ctx *ctx, int dim, int mem, int min_len, int type1, double regularization, double relaxation,
           double safeguard_factor, double max_weight_norm, int ir_max_steps);
  void destroy();
  // f (device, dim) is overwritten with the accelerated point when the step is accepted;
  // x = previous input.  vnorm2_out: device double that receives sum f^2 when f changes
  // (may be null).  Device-side return value in st->aa_norm.
  int apply(double *f, const double *x, double *vnorm2_out);
  // AA safeguard (aa.c:856-901); rolls f_new/x_new back on the device when rejected.
  // counters: optional device ints {rejected, accepted} incremented as scs.c:1386-1394 does.
  int safeguard(double *f_new, double *x_new, double *vnorm2_out, int *rej_cnt, int *acc_cnt);
  int reset();
  int fetch_state();  // device -> st_host (synchronises the stream)
};

}  // namespace b200
