// aa_small.cuh -- the parts of the Anderson-acceleration device code shared by the streaming
// engine (aa.cu) and the one-CTA-per-problem batch engine (batch.cu): kernel parameter block,
// aa_reset, and the small dense solve (aa.c:453-652) run by one thread on the len x C trapezoid.
#pragma once
#include "aa.cuh"

namespace b200 {

struct AaParams {  // by-value kernel argument
  int dim, mem, min_len, type1, ir_max_steps;
  double regularization, relaxation, safeguard_factor, max_weight_norm;
  double *x, *f, *g, *g_prev, *Y, *S, *D, *x_work, *Rpart;
  AaState *st;
  // row-partitioned mode: inner products / norms / the TSQR run over the rows [cnt_lo, cnt_hi) this rank
  // counts (the replicated shared block and tau belong to rank 0); [0, dim) everywhere else
  int cnt_lo, cnt_hi;
};

static __device__ __forceinline__ void aa_reset_dev(AaState *st, int mem) {  // aa_reset, aa.c:934-964
  st->iter = 0;
  st->success = 0;
  st->norm_g = 0.0;
  for (int i = 0; i < mem; ++i) { st->nrm_s_col[i] = 0.0; st->nrm_y_col[i] = 0.0; }
}

// Small dense stage executed by ONE thread on the merged len x C trapezoid (aa.c:453-652):
// column-pivoted Householder QR of R11, rank cut at len*eps*|R_11|, Q' applied to
// [R12 | r13], LU (type-I) or triangular (type-II) solve with iterative refinement.
static __device__ void aa_small_solve(const AaParams &a, const double *R, int len, int C, double r_reg, double *scr) {
  AaState *st = a.st;
  const int nrhs = a.type1 ? len + 1 : 1;
  double *Tm = scr;                 // len x len, column-major
  double *Bm = Tm + len * len;      // len x nrhs, column-major
  double *W = Bm + len * (len + 1);
  double *Wo = W + len * len;
  double *gam = Wo + len * len, *ctop = gam + len, *res = ctop + len, *gamma = res + len;
  int jpvt[kAaMaxMem + 1], ipiv[kAaMaxMem + 1];
  for (int j = 0; j < len; ++j) {
    jpvt[j] = j;
    for (int i = 0; i < len; ++i) Tm[i + j * len] = (i <= j) ? R[i * C + j] : 0.0;
  }
  for (int cc = 0; cc < nrhs; ++cc)
    for (int i = 0; i < len; ++i) Bm[i + cc * len] = R[i * C + len + cc];
  // ---- QR with column pivoting
  for (int j = 0; j < len; ++j) {
    double best = -1.0;
    int bi = j;
    for (int k = j; k < len; ++k) {
      double s = 0.0;
      for (int i = j; i < len; ++i) s = fma(Tm[i + k * len], Tm[i + k * len], s);
      if (s > best) { best = s; bi = k; }
    }
    if (bi != j) {
      for (int i = 0; i < len; ++i) { const double tv = Tm[i + j * len]; Tm[i + j * len] = Tm[i + bi * len]; Tm[i + bi * len] = tv; }
      const int tp = jpvt[j]; jpvt[j] = jpvt[bi]; jpvt[bi] = tp;
    }
    double xn2 = 0.0;
    for (int i = j + 1; i < len; ++i) xn2 = fma(Tm[i + j * len], Tm[i + j * len], xn2);
    if (xn2 == 0.0) continue;
    const double aj = Tm[j + j * len];
    const double nrm = sqrt(aj * aj + xn2);
    const double alpha = aj >= 0.0 ? -nrm : nrm;
    const double v0 = aj - alpha;
    const double beta = 2.0 / (v0 * v0 + xn2);
    for (int k = j + 1; k < len; ++k) {
      double w = v0 * Tm[j + k * len];
      for (int i = j + 1; i < len; ++i) w = fma(Tm[i + j * len], Tm[i + k * len], w);
      w *= beta;
      Tm[j + k * len] -= w * v0;
      for (int i = j + 1; i < len; ++i) Tm[i + k * len] -= w * Tm[i + j * len];
    }
    for (int k = 0; k < nrhs; ++k) {
      double w = v0 * Bm[j + k * len];
      for (int i = j + 1; i < len; ++i) w = fma(Tm[i + j * len], Bm[i + k * len], w);
      w *= beta;
      Bm[j + k * len] -= w * v0;
      for (int i = j + 1; i < len; ++i) Bm[i + k * len] -= w * Tm[i + j * len];
    }
    Tm[j + j * len] = alpha;
    for (int i = j + 1; i < len; ++i) Tm[i + j * len] = 0.0;
  }
  // ---- rank (aa.c:465-482)
  int info = 0, rank = 0;
  const double r11 = fabs(Tm[0]);
  if (r11 > 0.0) {
    const double tol = r11 * (double)len * DBL_EPSILON;
    for (rank = 0; rank < len; ++rank)
      if (fabs(Tm[rank + rank * len]) < tol) break;
  }
  if (rank == 0) info = 1;
  if (info == 0) {
    const double *crhs = Bm + (size_t)(nrhs - 1) * len;  // Q' [g;0]
    for (int i = 0; i < rank; ++i) ctop[i] = crhs[i];
    if (a.type1) {
      for (int cc = 0; cc < rank; ++cc)
        for (int i = 0; i < rank; ++i) {
          const double wv = Bm[i + jpvt[cc] * len];
          W[i + cc * rank] = wv;
          Wo[i + cc * rank] = wv;
        }
      // LU with partial pivoting (dgesv)
      for (int k = 0; k < rank && info == 0; ++k) {
        int pr = k;
        double pm = fabs(W[k + k * rank]);
        for (int i = k + 1; i < rank; ++i)
          if (fabs(W[i + k * rank]) > pm) { pm = fabs(W[i + k * rank]); pr = i; }
        ipiv[k] = pr;
        if (pm == 0.0 || !(pm == pm)) { info = k + 1; break; }
        if (pr != k)
          for (int cc = 0; cc < rank; ++cc) { const double tv = W[k + cc * rank]; W[k + cc * rank] = W[pr + cc * rank]; W[pr + cc * rank] = tv; }
        const double inv = 1.0 / W[k + k * rank];
        for (int i = k + 1; i < rank; ++i) W[i + k * rank] *= inv;
        for (int cc = k + 1; cc < rank; ++cc) {
          const double wk = W[k + cc * rank];
          for (int i = k + 1; i < rank; ++i) W[i + cc * rank] -= W[i + k * rank] * wk;
        }
      }
      if (info == 0) {
        auto lu_solve = [&](double *bv) {
          for (int k = 0; k < rank; ++k) { const double tv = bv[k]; bv[k] = bv[ipiv[k]]; bv[ipiv[k]] = tv; }
          for (int k = 0; k < rank; ++k)
            for (int i = k + 1; i < rank; ++i) bv[i] -= W[i + k * rank] * bv[k];
          for (int k = rank - 1; k >= 0; --k) {
            bv[k] /= W[k + k * rank];
            for (int i = 0; i < k; ++i) bv[i] -= W[i + k * rank] * bv[k];
          }
        };
        for (int i = 0; i < rank; ++i) gam[i] = ctop[i];
        lu_solve(gam);
        double prev = 0.0;
        for (int step = 0; step < a.ir_max_steps; ++step) {  // aa.c:534-550
          for (int i = 0; i < rank; ++i) {
            double rv = ctop[i];
            for (int cc = 0; cc < rank; ++cc) rv -= Wo[i + cc * rank] * gam[cc];
            res[i] = rv;
          }
          lu_solve(res);
          double dn = 0.0;
          for (int i = 0; i < rank; ++i) { dn = fma(res[i], res[i], dn); gam[i] += res[i]; }
          dn = sqrt(dn);
          if (step > 0 && dn >= 0.5 * prev) break;
          prev = dn;
        }
      }
    } else {
      auto tri_solve = [&](double *bv) {  // dtrsv Upper/NoTrans/NonUnit on the rank x rank block
        for (int k = rank - 1; k >= 0; --k) {
          bv[k] /= Tm[k + k * len];
          for (int i = 0; i < k; ++i) bv[i] -= Tm[i + k * len] * bv[k];
        }
      };
      for (int i = 0; i < rank; ++i) gam[i] = ctop[i];
      tri_solve(gam);
      double prev = 0.0;
      for (int step = 0; step < a.ir_max_steps; ++step) {  // aa.c:564-582
        for (int i = 0; i < rank; ++i) {
          double rv = 0.0;
          for (int cc = i; cc < rank; ++cc) rv = fma(Tm[i + cc * len], gam[cc], rv);
          res[i] = ctop[i] - rv;
        }
        tri_solve(res);
        double dn = 0.0;
        for (int i = 0; i < rank; ++i) { dn = fma(res[i], res[i], dn); gam[i] += res[i]; }
        dn = sqrt(dn);
        if (step > 0 && dn >= 0.5 * prev) break;
        prev = dn;
      }
    }
  }
  double aa_norm = -1.0;
  if (info == 0) {
    for (int i = 0; i < len; ++i) gamma[i] = 0.0;
    for (int i = 0; i < rank; ++i) gamma[jpvt[i]] = gam[i];
    double s = 0.0;
    for (int i = 0; i < len; ++i) s = fma(gamma[i], gamma[i], s);
    aa_norm = sqrt(s);
  }
  const bool finite = (aa_norm == aa_norm) && fabs(aa_norm) < INFINITY;
  st->last_rank = rank;
  st->last_regularization = r_reg;
  st->last_aa_norm = (info == 0 && finite) ? aa_norm : NAN;
  if (info != 0 || !finite || aa_norm >= a.max_weight_norm) {  // aa.c:612-638
    if (rank == 0) st->n_reject_rank0++;
    else if (!finite) st->n_reject_nonfinite++;
    else st->n_reject_weight_cap++;
    aa_reset_dev(st, a.mem);
    if (!finite) aa_norm = -1.0;
    st->aa_norm = (aa_norm < 0) ? aa_norm : -aa_norm;
    st->success = 0;
    st->iter = 1;  // aa_reset() followed by the unconditional a->iter++ of aa_apply (aa.c:851)
    return;
  }
  for (int i = 0; i < len; ++i) st->gamma[i] = gamma[i];
  st->success = 1;
  st->aa_norm = aa_norm;
  if (aa_norm > 0) st->n_accept++;
  st->iter = st->iter + 1;
}

}  // namespace b200
