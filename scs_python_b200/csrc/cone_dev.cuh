// cone_dev.cuh -- cone projections that run inside another kernel: per-thread and per-warp device functions shared
// by the streaming engine (cones.cu: k_exp_cones / k_pow_cones) and the one-CTA batch engine (batch.cu).
//
//   second-order cone        one warp, shuffle-reduced norm (cones.c:1242-1271), Moreau step fused
//   PSD / complex PSD cone   one warp, two-sided Jacobi with round-robin ordering on a matrix in shared memory
//                            (replaces dsyevr + dsyrk / zheevr + zherk, cones.c:991-1148, for small orders)
//   exponential / power      one thread per three-row cone: scalar root finders whose constants and branch order
//                            follow the reference (S/src/exp_cone.c, S/src/cones.c:1276-1324) because parity with
//                            its iterates depends on them
#pragma once
#include "common.cuh"

namespace b200 {
namespace {

// =========================================================== second-order / PSD, one warp per cone ======
// one second-order cone of the Moreau step, by one warp (cones.c:1242-1271 inside cones.c:1562-1585)
__device__ __forceinline__ void soc_moreau_warp(double *uy, const double *ry, int len, int lane) {
  if (len <= 0) return;
  // x = -r s ; u = Pi_K(x) / r + s
  double nn = 0.0;
  for (int k = 1 + lane; k < len; k += 32) { const double xk = -ry[k] * uy[k]; nn = fma(xk, xk, nn); }
  nn = warp_sum(nn);
  const double s0 = uy[0], r0 = ry[0];
  const double v1 = -r0 * s0;
  __syncwarp();  // every lane holds s0 before lane 0 may overwrite uy[0]
  if (len == 1) {
    if (lane == 0) uy[0] = fmax(v1, 0.0) / r0 + s0;
    return;
  }
  const double sn = len > 2 ? sqrt(nn) : fabs(-ry[1] * uy[1]);
  if (sn <= v1) {  // x in K: Pi(x) = x, u = -s + s = x / r + s
    for (int k = lane; k < len; k += 32) { const double sk = uy[k]; uy[k] = (-ry[k] * sk) / ry[k] + sk; }
    return;
  }
  if (sn <= -v1) return;  // Pi(x) = 0: u = s
  const double alpha = (sn + v1) / 2.0, sc = alpha / sn;
  for (int k = 1 + lane; k < len; k += 32) { const double sk = uy[k]; uy[k] = ((-ry[k] * sk) * sc) / ry[k] + sk; }
  if (lane == 0) uy[0] = alpha / r0 + s0;
}

// One positive-semidefinite cone of the Moreau step, by one warp (cones.c:991-1148 inside cones.c:1562-1585; replaces
// dsyevr + dsyrk / zheevr + zherk for the small orders a batch member has).  x = -r s is unpacked into a d x d
// symmetric matrix in shared memory -- real cone: lower triangle, column-major, off-diagonals scaled by sqrt 2;
// complex cone of order n (cones.c:1087-1095: per column the real diagonal entry, then (re, im) pairs): the real
// embedding [[Re, -Im], [Im, Re]] of order d = 2n, whose projection is the embedding of the projection.  Classical
// two-sided Jacobi with the round-robin ordering -- the d/2 disjoint pairs of a round rotate together: angles from the
// current matrix, then the column updates of all pairs (matrix and eigenvector matrix), then the row updates --
// until a sweep finds no off-diagonal entry above 1e-17 ||X||_F; X+ = sum_{lambda_k > 0} lambda_k v_k v_k' is
// re-packed and u = X+ / r + s.
struct PsdIdx { int i, j, part; };  // packed index -> entry (i >= j); part: 0 diagonal / real part, 1 imaginary part
__device__ __forceinline__ PsdIdx psd_index(int k, int n, bool cplx) {
  int j = 0, rem = k;
  if (!cplx) {
    while (rem >= n - j) { rem -= n - j; ++j; }
    return PsdIdx{j + rem, j, 0};
  }
  while (rem >= 2 * (n - j) - 1) { rem -= 2 * (n - j) - 1; ++j; }
  if (rem == 0) return PsdIdx{j, j, 0};
  return PsdIdx{j + 1 + ((rem - 1) >> 1), j, (rem - 1) & 1};
}
__device__ void psd_moreau_warp(double *uy, const double *ry, int len, bool cplx, double *ws, int lane) {
  if (len <= 0) return;
  if (len == 1) {
    if (lane == 0) { const double s0 = uy[0], r0 = ry[0]; uy[0] = fmax(-r0 * s0, 0.0) / r0 + s0; }
    return;
  }
  const int n = cplx ? (int)(sqrt((double)len) + 0.5) : (int)((sqrt(8.0 * len + 1.0) - 1.0) * 0.5 + 0.5);
  const int d = cplx ? 2 * n : n;
  double *A = ws, *V = ws + d * d, *cs = V + d * d;
  const double isq2 = 0.70710678118654752440, sq2 = 1.41421356237309504880;
  for (int k = lane; k < d * d; k += 32) { A[k] = 0.0; V[k] = (k / d == k % d) ? 1.0 : 0.0; }
  __syncwarp();
  for (int k = lane; k < len; k += 32) {
    const PsdIdx e = psd_index(k, n, cplx);
    const double x = -ry[k] * uy[k];
    const double a = e.i == e.j ? x : x * isq2;
    if (!cplx) {
      A[e.i * d + e.j] = a; A[e.j * d + e.i] = a;
    } else if (e.part == 0) {
      A[e.i * d + e.j] = a; A[e.j * d + e.i] = a;
      A[(n + e.i) * d + n + e.j] = a; A[(n + e.j) * d + n + e.i] = a;
    } else {  // Im H[i][j] = a, Im H[j][i] = -a: top-right block -Im H, bottom-left block Im H
      A[e.i * d + n + e.j] = -a; A[(n + e.j) * d + e.i] = -a;
      A[e.j * d + n + e.i] = a; A[(n + e.i) * d + e.j] = a;
    }
  }
  __syncwarp();
  double fro = 0.0;
  for (int k = lane; k < d * d; k += 32) fro = fma(A[k], A[k], fro);
  fro = sqrt(warp_sum(fro));
  const int de = d + (d & 1), half = de >> 1;
  const double small = 1e-17 * fro;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double mx = 0.0;
    for (int r = 0; r < de - 1; ++r) {
      for (int k = lane; k < half; k += 32) {  // rotation of pair k (Numerical-Recipes form), identity when negligible
        const int p = k == 0 ? de - 1 : (r + k) % (de - 1), q = k == 0 ? r : (r - k + de - 1) % (de - 1);
        double c = 1.0, sn = 0.0;
        if (p < d && q < d) {
          const double apq = A[p * d + q];
          mx = fmax(mx, fabs(apq));
          if (fabs(apq) > small) {
            const double th = (A[q * d + q] - A[p * d + p]) / (2.0 * apq);
            const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            sn = t * c;
          }
        }
        cs[2 * k] = c; cs[2 * k + 1] = sn;
      }
      __syncwarp();
      for (int idx = lane; idx < half * d; idx += 32) {  // columns p, q of the matrix and of V
        const int k = idx / d, i = idx - k * d;
        const int p = k == 0 ? de - 1 : (r + k) % (de - 1), q = k == 0 ? r : (r - k + de - 1) % (de - 1);
        const double c = cs[2 * k], sn = cs[2 * k + 1];
        if (p < d && q < d && sn != 0.0) {
          const double ap = A[i * d + p], aq = A[i * d + q];
          A[i * d + p] = c * ap - sn * aq; A[i * d + q] = sn * ap + c * aq;
          const double vp = V[i * d + p], vq = V[i * d + q];
          V[i * d + p] = c * vp - sn * vq; V[i * d + q] = sn * vp + c * vq;
        }
      }
      __syncwarp();
      for (int idx = lane; idx < half * d; idx += 32) {  // rows p, q of the matrix
        const int k = idx / d, j = idx - k * d;
        const int p = k == 0 ? de - 1 : (r + k) % (de - 1), q = k == 0 ? r : (r - k + de - 1) % (de - 1);
        const double c = cs[2 * k], sn = cs[2 * k + 1];
        if (p < d && q < d && sn != 0.0) {
          const double ap = A[p * d + j], aq = A[q * d + j];
          A[p * d + j] = c * ap - sn * aq; A[q * d + j] = sn * ap + c * aq;
        }
      }
      __syncwarp();
    }
    mx = warp_max(mx);
    if (mx <= small) break;  // warp-uniform
  }
  for (int k = lane; k < len; k += 32) {
    const PsdIdx e = psd_index(k, n, cplx);
    const int ra = e.part ? n + e.i : e.i, rb = e.j;  // Re H+[i][j] = S+[i][j], Im H+[i][j] = S+[n + i][j]
    double x = 0.0;
    for (int t = 0; t < d; ++t) {
      const double lam = A[t * d + t];
      if (lam > 0.0) x = fma(lam * V[ra * d + t], V[rb * d + t], x);
    }
    if (e.i != e.j) x *= sq2;
    const double sk = uy[k];
    uy[k] = x / ry[k] + sk;
  }
  __syncwarp();  // the workspace is reused by this warp's next cone
}

// =================================================================== exponential ======
// Friberg 2021 as restated by the reference (exp_cone.c); v0 = (r0, s0, t0).
#define EXP_INF 1e15
__device__ __forceinline__ bool exp_isfinite(double x) { return fabs(x) < EXP_INF; }
__device__ __forceinline__ double clipd(double x, double l, double u) { return fmax(l, fmin(u, x)); }
__device__ __forceinline__ double safediv_pos(double x, double y) { return y < 1e-18 ? x / 1e-18 : x / y; }
__device__ __forceinline__ double nds3(const double *a, const double *b) {
  const double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
  return d0 * d0 + d1 * d1 + d2 * d2;
}
__device__ __forceinline__ double hfun_f(const double *v0, double rho) {  // exp_cone.c:41-48
  const double t0 = v0[2], s0 = v0[1], r0 = v0[0];
  const double e = exp(rho), en = 1.0 / e;
  return ((rho - 1) * r0 + s0) * e - (r0 - rho * s0) * en - (rho * (rho - 1) + 1) * t0;
}
__device__ __forceinline__ void hfun_fd(const double *v0, double rho, double *f, double *df) {  // :50-62
  const double t0 = v0[2], s0 = v0[1], r0 = v0[0];
  const double e = exp(rho), en = 1.0 / e;
  *f = ((rho - 1) * r0 + s0) * e - (r0 - rho * s0) * en - (rho * (rho - 1) + 1) * t0;
  *df = (rho * r0 + s0) * e + (r0 - (rho - 1) * s0) * en - (2 * rho - 1) * t0;
}
__device__ double root_search_binary(const double *v0, double xl, double xu, double x) {  // :65-95
  double x_plus = x;
  for (int i = 0; i < 40; ++i) {
    const double f = hfun_f(v0, x);
    if (f < 0.0) xl = x; else xu = x;
    x_plus = 0.5 * (xl + xu);
    if (fabs(x_plus - x) <= 1e-12 * fmax(1.0, fabs(x_plus)) || x_plus == xl || x_plus == xu) break;
    x = x_plus;
  }
  return x_plus;
}
__device__ double root_search_newton(const double *v0, double xl, double xu, double x) {  // :98-162
  const double EPS = 1e-15, DFTOL = 1e-13, LODAMP = 0.05, HIDAMP = 0.95;
  int i;
  for (i = 0; i < 20; ++i) {
    double f, df;
    hfun_fd(v0, x, &f, &df);
    if (fabs(f) <= EPS) break;
    if (f < 0.0) xl = x; else xu = x;
    if (xu <= xl) { xu = 0.5 * (xu + xl); xl = xu; break; }
    if (!exp_isfinite(f) || df < DFTOL) break;
    const double x_plus = x - f / df;
    if (fabs(x_plus - x) <= EPS * fmax(1.0, fabs(x_plus))) break;
    if (x_plus >= xu) x = fmin(LODAMP * x + HIDAMP * xu, xu);
    else if (x_plus <= xl) x = fmax(LODAMP * x + HIDAMP * xl, xl);
    else x = x_plus;
  }
  if (i < 20) return clipd(x, xl, xu);
  return root_search_binary(v0, xl, xu, x);
}
__device__ double exp_primal_heur(const double *v0, double *vp) {  // :165-188
  const double t0 = v0[2], s0 = v0[1], r0 = v0[0];
  vp[2] = fmax(t0, 0.0); vp[1] = 0.0; vp[0] = fmin(r0, 0.0);
  double dist = nds3(v0, vp);
  if (s0 > 0.0) {
    const double tp = fmax(t0, s0 * exp(r0 / s0));
    const double nd = (tp - t0) * (tp - t0);
    if (nd < dist) { vp[2] = tp; vp[1] = s0; vp[0] = r0; dist = nd; }
  }
  return dist;
}
__device__ double exp_polar_heur(const double *v0, double *vd) {  // :191-214
  const double t0 = v0[2], s0 = v0[1], r0 = v0[0];
  vd[2] = fmin(t0, 0.0); vd[1] = fmin(s0, 0.0); vd[0] = 0.0;
  double dist = nds3(v0, vd);
  if (r0 > 0.0) {
    const double td = fmin(t0, -r0 * exp(s0 / r0 - 1.0));
    const double nd = (t0 - td) * (t0 - td);
    if (nd < dist) { vd[2] = td; vd[1] = s0; vd[0] = r0; dist = nd; }
  }
  return dist;
}
__device__ __forceinline__ double ppsi(const double *v0) {  // :216-227
  const double s0 = v0[1], r0 = v0[0];
  double psi;
  if (r0 > s0) psi = (r0 - s0 + sqrt(r0 * r0 + s0 * s0 - r0 * s0)) / r0;
  else psi = -s0 / (r0 - s0 - sqrt(r0 * r0 + s0 * s0 - r0 * s0));
  return ((psi - 1.0) * r0 + s0) / (psi * (psi - 1.0) + 1.0);
}
__device__ __forceinline__ double pomega(double rho) {  // :229-236
  double val = exp(rho) / (rho * (rho - 1.0) + 1.0);
  if (rho < 2.0) val = fmin(val, exp(2.0) / 3.0);
  return val;
}
__device__ __forceinline__ double dpsi(const double *v0) {  // :238-249
  const double s0 = v0[1], r0 = v0[0];
  double psi;
  if (s0 > r0) psi = (r0 - sqrt(r0 * r0 + s0 * s0 - r0 * s0)) / s0;
  else psi = (r0 - s0) / (r0 + sqrt(r0 * r0 + s0 * s0 - r0 * s0));
  return (r0 - psi * s0) / (psi * (psi - 1.0) + 1.0);
}
__device__ __forceinline__ double domega(double rho) {  // :251-258
  double val = -exp(-rho) / (rho * (rho - 1.0) + 1.0);
  if (rho > -1.0) val = fmax(val, -exp(1.0) / 3.0);
  return val;
}
__device__ void exp_search_bracket(const double *v0, double pdist_sq, double ddist_sq, double *low_out,
                                   double *upr_out) {  // :261-323
  const double t0 = v0[2], s0 = v0[1], r0 = v0[0];
  double baselow = -EXP_INF, baseupr = EXP_INF, low = -EXP_INF, upr = EXP_INF;
  const double ms0 = fmin(s0, 0.0), mr0 = fmin(r0, 0.0);
  const double Dp = sqrt(fmax(pdist_sq - ms0 * ms0, 0.0));
  const double Dd = sqrt(fmax(ddist_sq - mr0 * mr0, 0.0));
  double curbnd, val, sgn;
  if (t0 > 0.0) {
    curbnd = log(t0 / ppsi(v0));
    low = fmax(low, curbnd);
  } else if (t0 < 0.0) {
    curbnd = -log(-t0 / dpsi(v0));
    upr = fmin(upr, curbnd);
  }
  if (r0 > 0.0) {
    baselow = 1.0 - s0 / r0;
    low = fmax(low, baselow);
    const double tpu = fmax(1e-12, fmin(Dd, Dp + t0));
    val = r0 * pomega(low);
    sgn = val < 0 ? -1 : 1;
    curbnd = fmax(low, baselow + safediv_pos(tpu, fabs(val)) * sgn);
    upr = fmin(upr, curbnd);
  }
  if (s0 > 0.0) {
    baseupr = r0 / s0;
    upr = fmin(upr, baseupr);
    const double tdl = -fmax(1e-12, fmin(Dp, Dd - t0));
    val = s0 * domega(upr);
    sgn = val < 0 ? -1 : 1;
    curbnd = fmin(upr, baseupr - safediv_pos(tdl, fabs(val)) * sgn);
    low = fmax(low, curbnd);
  }
  low = clipd(fmin(low, upr), baselow, baseupr);
  upr = clipd(fmax(low, upr), baselow, baseupr);
  if (low != upr) {
    const double fl = hfun_f(v0, low), fu = hfun_f(v0, upr);
    if (fl * fu > 0.0) {
      if (fabs(fl) < fabs(fu)) upr = low; else low = upr;
    }
  }
  *low_out = low;
  *upr_out = upr;
}
// SCS(proj_pd_exp_cone), exp_cone.c:373-441
__device__ void proj_pd_exp_cone(double *v0, int primal) {
  const double TOL = 1e-8;
  double vp[3], vd[3], vh[3];
  if (!primal) { v0[0] = -v0[0]; v0[1] = -v0[1]; v0[2] = -v0[2]; }
  double pdist_sq = exp_primal_heur(v0, vp);
  double ddist_sq = exp_polar_heur(v0, vd);
  double err = fabs(vp[0] + vd[0] - v0[0]);
  err = fmax(err, fabs(vp[1] + vd[1] - v0[1]));
  err = fmax(err, fabs(vp[2] + vd[2] - v0[2]));
  bool opt = (v0[1] <= 0.0 && v0[0] <= 0.0);
  opt = opt || (fmin(pdist_sq, ddist_sq) <= TOL * TOL);
  opt = opt || (err <= TOL && (vp[0] * vd[0] + vp[1] * vd[1] + vp[2] * vd[2]) <= TOL);
  if (!opt) {
    double xl, xh;
    exp_search_bracket(v0, pdist_sq, ddist_sq, &xl, &xh);
    const double rho = root_search_newton(v0, xl, xh, 0.5 * (xl + xh));
    if (primal) {  // proj_sol_primal_exp_cone, :326-345
      const double linrho = (rho - 1.0) * v0[0] + v0[1];
      const double exprho = exp(rho);
      double dh;
      if (linrho > 0.0 && exp_isfinite(exprho)) {
        const double quad = rho * (rho - 1.0) + 1.0;
        vh[2] = exprho * linrho / quad; vh[1] = linrho / quad; vh[0] = rho * linrho / quad;
        dh = nds3(vh, v0);
      } else {
        vh[2] = EXP_INF; vh[1] = 0.0; vh[0] = 0.0; dh = EXP_INF;
      }
      if (dh <= pdist_sq) { vp[0] = vh[0]; vp[1] = vh[1]; vp[2] = vh[2]; }
    } else {       // proj_sol_polar_exp_cone, :348-367
      const double linrho = v0[0] - rho * v0[1];
      const double exprho = exp(-rho);
      double dh;
      if (linrho > 0.0 && exp_isfinite(exprho)) {
        const double quad = rho * (rho - 1.0) + 1.0;
        vh[2] = -exprho * linrho / quad; vh[1] = (1.0 - rho) * linrho / quad; vh[0] = linrho / quad;
        dh = nds3(v0, vh);
      } else {
        vh[2] = -EXP_INF; vh[1] = 0.0; vh[0] = 0.0; dh = EXP_INF;
      }
      if (dh <= ddist_sq) { vd[0] = vh[0]; vd[1] = vh[1]; vd[2] = vh[2]; }
    }
  }
  if (primal) { v0[0] = vp[0]; v0[1] = vp[1]; v0[2] = vp[2]; }
  else { v0[0] = -vd[0]; v0[1] = -vd[1]; v0[2] = -vd[2]; }
}

// ========================================================================= power ======
__device__ __forceinline__ double pow_calc_x(double r, double xh, double rh, double a) {  // cones.c:1276-1280
  const double x = 0.5 * (xh + sqrt(xh * xh + 4 * a * (rh - r) * r));
  return fmax(x, 1e-12);
}
__device__ void proj_power_cone(double *v, double a) {  // cones.c:1282-1324
  const double xh = v[0], yh = v[1], rh = fabs(v[2]);
  double x = 0.0, y = 0.0, r;
  if (xh >= 0 && yh >= 0 && 1e-9 + pow(xh, a) * pow(yh, 1 - a) >= rh) return;
  if (xh <= 0 && yh <= 0 && 1e-9 + pow(-xh, a) * pow(-yh, 1 - a) >= rh * pow(a, a) * pow(1 - a, 1 - a)) {
    v[0] = v[1] = v[2] = 0;
    return;
  }
  r = rh / 2;
  for (int i = 0; i < 20; ++i) {
    x = pow_calc_x(r, xh, rh, a);
    y = pow_calc_x(r, yh, rh, 1 - a);
    const double xa = pow(x, a), y1a = pow(y, 1 - a);
    const double f = xa * y1a - r;
    if (fabs(f) < 1e-9) break;
    const double dxdr = a * (rh - 2 * r) / (2 * x - xh);
    const double dydr = (1 - a) * (rh - 2 * r) / (2 * y - yh);
    const double fp = xa * y1a * (a * dxdr / x + (1 - a) * dydr / y) - 1;
    r = fmax(r - f / fp, 0.0);
    r = fmin(r, rh);
  }
  v[0] = x; v[1] = y; v[2] = (v[2] < 0) ? -r : r;
}

// Moreau step of one three-dimensional cone (cones.c:1562-1585): x = -R s is already in v; kind > 0: primal
// exponential cone, kind == 0: dual exponential cone, kind < 0: power cone with parameter a (a < 0: its dual,
// cones.c:1423-1432)
__device__ __forceinline__ void proj_cone3(double *v, int kind, double a) {
  if (kind >= 0) {
    proj_pd_exp_cone(v, kind > 0);
  } else if (a >= 0) {
    proj_power_cone(v, a);
  } else {
    double w[3] = {-v[0], -v[1], -v[2]};
    proj_power_cone(w, -a);
    v[0] += w[0]; v[1] += w[1]; v[2] += w[2];
  }
}

}  // namespace
}  // namespace b200
