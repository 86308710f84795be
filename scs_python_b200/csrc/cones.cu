// cones.cu -- cone projection kernels for sm_100a.
//
//   box    : one CTA, Newton on t with a block-reduced gradient/Hessian (cones.c:1174-1237)
//   SOC    : one warp per cone (one CTA for cones > 2048), shuffle-reduced norm (cones.c:1242-1271)
//   exp    : one thread per triple, Friberg-2021 root search (exp_cone.c)
//   power  : one thread per triple, Newton on r (cones.c:1276-1324)
//   PSD / complex PSD : one CTA per cone, parallel-ordered one-sided Jacobi on the shifted
//            matrix, reconstruction from the positive eigenpairs (replaces dsyevr+dsyrk /
//            zheevr+zherk, cones.c:991-1148); complex cones go through the real 2s x 2s
//            embedding [[Re,-Im],[Im,Re]].
#include "cones.cuh"

namespace b200 {

// =========================================================================== box ======
constexpr int kBoxThreads = 1024;

__global__ void __launch_bounds__(kBoxThreads)
k_box_cone(double *__restrict__ tx, const double *__restrict__ s_saved, const double *__restrict__ r_box,
           const double *__restrict__ bl, const double *__restrict__ bu, int bsize, double *t_warm) {
  __shared__ double sh[2 * 32];
  __shared__ double sh_t;
  __shared__ int sh_stop;
  const int tid = threadIdx.x;
  if (bsize == 1) {
    if (tid == 0) tx[0] = fmax(tx[0], 0.0) / r_box[0] + s_saved[0];
    return;
  }
  double *x = tx + 1;
  const double *rho = r_box + 1;
  const double rho_t = 1.0 / r_box[0];
  const double tx0 = tx[0];
  double t = *t_warm;
  for (int iter = 0; iter < 25; ++iter) {  // BOX_CONE_MAX_ITERS
    double v[2] = {0.0, 0.0};              // gt, ht partial sums
    for (int j = tid; j < bsize - 1; j += kBoxThreads) {
      const double rinv = 1.0 / rho[j];
      const double xj = x[j], u = bu[j], lo = bl[j];
      if (xj > t * u) {
        v[0] += rinv * (t * u - xj) * u;
        v[1] += rinv * u * u;
      } else if (xj < t * lo) {
        v[0] += rinv * (t * lo - xj) * lo;
        v[1] += rinv * lo * lo;
      }
    }
    block_reduce<2, 0>(v, sh);
    if (tid == 0) {
      const double gt = rho_t * (t - tx0) + v[0];
      const double ht = rho_t + v[1];
      const double t_prev = t;
      t = fmax(t - gt / fmax(ht, 1e-8), 0.0);
      sh_t = t;
      sh_stop = (fabs(gt / fmax(ht, 1e-6)) < 1e-12 * fmax(t, 1.0) || fabs(t - t_prev) < 1e-11 * fmax(t, 1.0));
    }
    __syncthreads();
    t = sh_t;
    const int stop = sh_stop;
    __syncthreads();
    if (stop) break;
  }
  for (int j = tid; j < bsize - 1; j += kBoxThreads) {
    double xj = x[j];
    if (xj > t * bu[j]) xj = t * bu[j];
    else if (xj < t * bl[j]) xj = t * bl[j];
    x[j] = xj / rho[j] + s_saved[1 + j];
  }
  if (tid == 0) {
    tx[0] = t / r_box[0] + s_saved[0];
    *t_warm = t;
  }
}

// =========================================================================== SOC ======
template <int GROUP>
__device__ __forceinline__ double group_sum(double v, double *sh) {
  v = warp_sum(v);
  if (GROUP == 32) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < GROUP / 32; ++w) t += sh[w];
  return t;
}

template <int GROUP>
__device__ __forceinline__ void soc_project(double *__restrict__ x, const double *__restrict__ sv,
                                            const double *__restrict__ r, int q, int lane, double *sh) {
  if (q <= 0) return;
  if (q == 1) {
    if (lane == 0) x[0] = fmax(x[0], 0.0) / r[0] + sv[0];
    return;
  }
  const double v1 = x[0];
  double ss = 0.0;
  for (int k = 1 + lane; k < q; k += GROUP) ss = fma(x[k], x[k], ss);
  ss = group_sum<GROUP>(ss, sh);
  const double s = (q == 2) ? fabs(x[1]) : sqrt(ss);
  const double alpha = (s + v1) / 2.0;
  double scale_tail, head;
  if (s <= v1) {            // inside the cone
    scale_tail = 1.0; head = v1;
  } else if (s <= -v1) {    // inside the polar cone -> 0
    scale_tail = 0.0; head = 0.0;
  } else {
    scale_tail = alpha / s; head = alpha;
  }
  for (int k = 1 + lane; k < q; k += GROUP) x[k] = (x[k] * scale_tail) / r[k] + sv[k];
  if (lane == 0) x[0] = head / r[0] + sv[0];
}

__global__ void __launch_bounds__(kThreads)
k_soc_small(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r,
            const int *__restrict__ off, const int *__restrict__ len, int ncones) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int cidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cidx < ncones; cidx += warps_per_grid) {
    const int o = off[cidx];
    soc_project<32>(x + o, sv + o, r + o, len[cidx], lane, nullptr);
  }
}

__global__ void __launch_bounds__(kThreads)
k_soc_large(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r,
            const int *__restrict__ off, const int *__restrict__ len, int ncones) {
  __shared__ double sh[kThreads / 32];
  for (int cidx = blockIdx.x; cidx < ncones; cidx += gridDim.x) {
    const int o = off[cidx];
    soc_project<kThreads>(x + o, sv + o, r + o, len[cidx], threadIdx.x, sh);
    __syncthreads();
  }
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

__global__ void __launch_bounds__(128)
k_exp_cones(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r, int ep, int ntot) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x) {
    double v[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    proj_pd_exp_cone(v, i < ep);
#pragma unroll
    for (int k = 0; k < 3; ++k) x[3 * i + k] = v[k] / r[3 * i + k] + sv[3 * i + k];
  }
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
__global__ void __launch_bounds__(128)
k_pow_cones(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r,
            const double *__restrict__ pw, int ntot) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x) {
    double v[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    const double a = pw[i];
    if (a >= 0) {
      proj_power_cone(v, a);
    } else {  // dual power cone via Moreau, cones.c:1423-1432
      double w[3] = {-v[0], -v[1], -v[2]};
      proj_power_cone(w, -a);
      v[0] += w[0]; v[1] += w[1]; v[2] += w[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) x[3 * i + k] = v[k] / r[3 * i + k] + sv[3 * i + k];
  }
}

// =========================================================== PSD / complex PSD ========
// packed index of entry (i,j), i >= j, of the real lower-triangular column-major layout
__device__ __forceinline__ long long tri_idx(int i, int j, int n) {
  return (long long)j * n - ((long long)(j - 1) * j) / 2 + (i - j);
}

constexpr int kPsdThreads = 512;
constexpr int kPsdMaxSweeps = 40;

__global__ void __launch_bounds__(kPsdThreads)
k_psd_cones(double *__restrict__ xall, const double *__restrict__ svall, const double *__restrict__ rall,
            const PsdEntry *__restrict__ ents, double *__restrict__ Gall, double *__restrict__ Vall,
            double *__restrict__ lamall) {
  __shared__ double sh[2 * 32];
  __shared__ int sh_rot;
  __shared__ double sh_sigma;
  const PsdEntry e = ents[blockIdx.x];
  const int n = e.s, d = e.d, tid = threadIdx.x, nthr = blockDim.x;
  double *x = xall + e.off;
  const double *sv = svall + e.off, *r = rall + e.off;
  if (n == 0) return;
  if (n == 1) {
    if (tid == 0) x[0] = fmax(x[0], 0.0) / r[0] + sv[0];
    return;
  }
  double *G = Gall + e.woff, *V = Vall + e.woff, *lam = lamall + e.loff;
  const double sqrt2 = sqrt(2.0);
  // ---- unpack into the full symmetric d x d matrix (diagonal * sqrt2, cones.c:1011-1017)
  for (long long idx = tid; idx < (long long)d * d; idx += nthr) {
    const int col = (int)(idx / d), row = (int)(idx % d);
    double val;
    if (!e.is_complex) {
      const int i = row > col ? row : col, j = row > col ? col : row;
      val = x[tri_idx(i, j, n)];
      if (i == j) val *= sqrt2;
    } else {
      // H = A + iB, embedding [[A, -B], [B, A]]; column j of the packed layout starts at
      // j*(2n-j): real diagonal, then (re, im) pairs of rows j+1.. (cones.c:1088-1095)
      const int br = row / n, bc = col / n, i0 = row % n, j0 = col % n;
      const int i = i0 > j0 ? i0 : j0, j = i0 > j0 ? j0 : i0;
      const long long base = (long long)j * (2 * n - j);
      if (i == j) {
        val = (br == bc) ? x[base] * sqrt2 : 0.0;
      } else {
        const double re = x[base + 1 + 2 * (i - j - 1)], im = x[base + 2 + 2 * (i - j - 1)];
        // B is antisymmetric: B[i0][j0] = im if i0 > j0 else -im
        const double bij = (i0 > j0) ? im : -im;
        if (br == bc) val = re;
        else if (br == 1) val = bij;   // bottom-left block  = B
        else val = -bij;               // top-right block    = -B
      }
    }
    G[idx] = val;
    V[idx] = (row == col) ? 1.0 : 0.0;
  }
  __syncthreads();
  // ---- shift by sigma >= spectral radius so that W = A + sigma I is PSD (one-sided Jacobi
  //      needs distinct |eigenvalues| to separate +/- pairs)
  {
    double v[2] = {0.0, 0.0};
    for (long long idx = tid; idx < (long long)d * d; idx += nthr) v[0] = fma(G[idx], G[idx], v[0]);
    block_reduce<1, 0>(v, sh);
    if (tid == 0) sh_sigma = sqrt(v[0]);
    __syncthreads();
  }
  const double sigma = sh_sigma;
  if (sigma == 0.0) {  // zero matrix -> projection is zero
    const int len = e.is_complex ? n * n : n * (n + 1) / 2;
    for (int k = tid; k < len; k += nthr) x[k] = 0.0 / r[k] + sv[k];
    return;
  }
  for (int k = tid; k < d; k += nthr) G[(long long)k * d + k] += sigma;
  __syncthreads();
  // ---- parallel-ordered one-sided Jacobi: G <- G J, V <- V J
  const int dd = d + (d & 1);
  const int lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  const double tol = sqrt((double)d) * DBL_EPSILON;
  for (int sweep = 0; sweep < kPsdMaxSweeps; ++sweep) {
    if (tid == 0) sh_rot = 0;
    __syncthreads();
    for (int rnd = 0; rnd < dd - 1; ++rnd) {
      for (int k = wid; k < dd / 2; k += nw) {
        int p, q;
        if (k == 0) { p = dd - 1; q = rnd; }
        else { p = (rnd + k) % (dd - 1); q = (rnd - k + (dd - 1)) % (dd - 1); }
        if (p >= d || q >= d) continue;  // dummy column of an odd-sized problem
        if (p > q) { const int t = p; p = q; q = t; }
        double *gp = G + (long long)p * d, *gq = G + (long long)q * d;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int i = lane; i < d; i += 32) {
          const double u = gp[i], w = gq[i];
          a = fma(u, u, a); b = fma(w, w, b); g = fma(u, w, g);
        }
        a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
        if (fabs(g) > tol * sqrt(a * b) && g != 0.0) {
          const double zeta = (b - a) / (2.0 * g);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
          double *vp = V + (long long)p * d, *vq = V + (long long)q * d;
          for (int i = lane; i < d; i += 32) {
            const double u = gp[i], w = gq[i];
            gp[i] = cs * u - sn * w;
            gq[i] = sn * u + cs * w;
            const double vu = vp[i], vw = vq[i];
            vp[i] = cs * vu - sn * vw;
            vq[i] = sn * vu + cs * vw;
          }
          if (lane == 0) sh_rot = 1;
        }
      }
      __syncthreads();
    }
    const int rotated = sh_rot;
    __syncthreads();
    if (!rotated) break;
  }
  // ---- eigenvalues of A: lambda_i = v_i . g_i - sigma
  for (int k = wid; k < d; k += nw) {
    const double *gk = G + (long long)k * d, *vk = V + (long long)k * d;
    double acc = 0.0;
    for (int i = lane; i < d; i += 32) acc = fma(gk[i], vk[i], acc);
    acc = warp_sum(acc);
    if (lane == 0) lam[k] = acc - sigma;
  }
  __syncthreads();
  // ---- X+ = sum_{lambda>0} lambda v v' on the lower triangle (of the leading n columns
  //      for the complex embedding), re-packed with diagonal / sqrt2 and Moreau-recombined.
  const double isqrt2 = 1.0 / sqrt2;
  if (!e.is_complex) {
    const long long len = (long long)n * (n + 1) / 2;
    for (long long idx = tid; idx < (long long)n * n; idx += nthr) {
      const int col = (int)(idx / n), row = (int)(idx % n);
      if (row < col) continue;
      double acc = 0.0;
      for (int k = 0; k < d; ++k) {
        const double lk = lam[k];
        if (lk > 0.0) acc = fma(lk * V[(long long)k * d + row], V[(long long)k * d + col], acc);
      }
      if (row == col) acc *= isqrt2;
      const long long pi = tri_idx(row, col, n);
      x[pi] = acc / r[pi] + sv[pi];
    }
    (void)len;
  } else {
    for (long long idx = tid; idx < (long long)n * n; idx += nthr) {
      const int col = (int)(idx / n), row = (int)(idx % n);
      if (row < col) continue;
      double are = 0.0, aim = 0.0;  // A[row][col] (top-left) and B[row][col] (bottom-left)
      for (int k = 0; k < d; ++k) {
        const double lk = lam[k];
        if (lk > 0.0) {
          const double vc = lk * V[(long long)k * d + col];
          are = fma(V[(long long)k * d + row], vc, are);
          aim = fma(V[(long long)k * d + n + row], vc, aim);
        }
      }
      const long long base = (long long)col * (2 * n - col);
      if (row == col) {
        x[base] = (are * isqrt2) / r[base] + sv[base];
      } else {
        const long long pr = base + 1 + 2 * (row - col - 1), pi = pr + 1;
        x[pr] = are / r[pr] + sv[pr];
        x[pi] = aim / r[pi] + sv[pi];
      }
    }
  }
}

// ============================================== cone-boundary aggregation (setup) =====
__global__ void __launch_bounds__(kThreads)
k_enforce_boundaries(double *__restrict__ D, const int *__restrict__ off, const int *__restrict__ len, int ncones,
                     int use_mean) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int cidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cidx < ncones; cidx += warps_per_grid) {
    double *v = D + off[cidx];
    const int n = len[cidx];
    double acc = 0.0;
    if (use_mean) {
      for (int k = lane; k < n; k += 32) acc += v[k];
      acc = warp_sum(acc) / (double)n;   // SCS(mean), linalg.c
    } else {
      for (int k = lane; k < n; k += 32) acc = fmax(acc, fabs(v[k]));
      acc = warp_max(acc);               // SCS(norm_inf)
    }
    for (int k = lane; k < n; k += 32) v[k] = acc;
  }
}

// ============================================================================ host ====
int ConeDev::init(Ctx *ctx, const ScsCone *k, int m_) {
  c = ctx;
  m = m_;
  z = k->z; l = k->l; bsize = k->bsize; ep = k->ep; ed = k->ed;
  q.assign(k->q, k->q + (k->qsize > 0 ? k->qsize : 0));
  s.assign(k->s, k->s + (k->ssize > 0 ? k->ssize : 0));
  cs.assign(k->cs, k->cs + (k->cssize > 0 ? k->cssize : 0));
  p.assign(k->p, k->p + (k->psize > 0 ? k->psize : 0));
  if (bsize > 1) {
    bu.assign(k->bu, k->bu + bsize - 1);
    bl.assign(k->bl, k->bl + bsize - 1);
  }
  int cnt = z + l;
  off_box = cnt; cnt += bsize;
  off_q = cnt;
  std::vector<int> qo_s, ql_s, qo_l, ql_l, bo, blen;
  boundaries.clear();
  boundaries.push_back(z + l + bsize);
  for (int qi : q) {
    if (qi > 2048) { qo_l.push_back(cnt); ql_l.push_back(qi); }
    else { qo_s.push_back(cnt); ql_s.push_back(qi); }
    if (qi > 0) { bo.push_back(cnt); blen.push_back(qi); }
    boundaries.push_back(qi);
    cnt += qi;
  }
  off_s = cnt;
  std::vector<PsdEntry> ents;
  long long woff = 0;
  int loff = 0;
  psd_max_d = 0;
  for (int si : s) {
    PsdEntry e{cnt, si, si, 0, woff, loff, 0};
    ents.push_back(e);
    const int sz = si * (si + 1) / 2;
    if (sz > 0) { bo.push_back(cnt); blen.push_back(sz); }
    boundaries.push_back(sz);
    woff += (long long)si * si; loff += si; cnt += sz;
    if (si > psd_max_d) psd_max_d = si;
  }
  off_cs = cnt;
  for (int ci : cs) {
    PsdEntry e{cnt, ci, 2 * ci, 1, woff, loff, 0};
    ents.push_back(e);
    const int sz = ci * ci;
    if (sz > 0) { bo.push_back(cnt); blen.push_back(sz); }
    boundaries.push_back(sz);
    woff += 4ll * ci * ci; loff += 2 * ci; cnt += sz;
    if (2 * ci > psd_max_d) psd_max_d = 2 * ci;
  }
  off_exp = cnt;
  for (int i = 0; i < ep + ed; ++i) { bo.push_back(cnt); blen.push_back(3); boundaries.push_back(3); cnt += 3; }
  off_pow = cnt;
  for (size_t i = 0; i < p.size(); ++i) { bo.push_back(cnt); blen.push_back(3); boundaries.push_back(3); cnt += 3; }
  if (cnt != m) {
    B200_PRINTF("Error: Cone dims %li != rows in A %li\n", (long)cnt, (long)m);
    return -1;
  }
  CUDA_OK(cudaSetDevice(c->device));
  // SOC lists: small first, then large
  n_q_small = (int)qo_s.size(); n_q_large = (int)qo_l.size();
  std::vector<int> qo(qo_s), ql(ql_s);
  qo.insert(qo.end(), qo_l.begin(), qo_l.end());
  ql.insert(ql.end(), ql_l.begin(), ql_l.end());
  if (!qo.empty()) {
    if (dev_alloc(&q_off, qo.size()) || dev_alloc(&q_len, ql.size()) || h2d(*c, q_off, qo.data(), qo.size()) ||
        h2d(*c, q_len, ql.data(), ql.size()))
      return -1;
  }
  n_psd = (int)ents.size();
  if (n_psd) {
    if (dev_alloc(&psd, ents.size()) || h2d(*c, psd, ents.data(), ents.size()) ||
        dev_alloc(&psd_G, (size_t)woff) || dev_alloc(&psd_V, (size_t)woff) || dev_alloc(&psd_lam, (size_t)loff + 1))
      return -1;
  }
  if (bsize > 0) {
    const double one = 1.0;
    if (dev_alloc(&box_t, 1) || h2d(*c, box_t, &one, 1)) return -1;
    if (dev_alloc(&d_bu, (size_t)(bsize > 1 ? bsize - 1 : 1)) || dev_alloc(&d_bl, (size_t)(bsize > 1 ? bsize - 1 : 1)))
      return -1;
  }
  if (!p.empty()) {
    if (dev_alloc(&d_p, p.size()) || h2d(*c, d_p, p.data(), p.size())) return -1;
  }
  n_bnd = (int)bo.size();
  if (n_bnd) {
    if (dev_alloc(&bnd_off, bo.size()) || dev_alloc(&bnd_len, blen.size()) || h2d(*c, bnd_off, bo.data(), bo.size()) ||
        h2d(*c, bnd_len, blen.data(), blen.size()))
      return -1;
  }
  if (c->sync()) return -1;
  return 0;
}

void ConeDev::destroy() {
  if (!c) return;
  cudaSetDevice(c->device);
  dev_free(q_off); dev_free(q_len); dev_free(psd); dev_free(psd_G); dev_free(psd_V); dev_free(psd_lam);
  dev_free(d_bu); dev_free(d_bl); dev_free(box_t); dev_free(d_p); dev_free(bnd_off); dev_free(bnd_len);
}

int ConeDev::normalize_box(const double *D_host) {
  // normalize_box_cone, cones.c:1153-1169 (runs once: scaled_cones latch, cones.c:1549-1557)
  if (scaled_cones) return 0;
  scaled_cones = true;
  if (bsize <= 1) return 0;
  if (D_host) {  // only when a scaling exists (cones.c:1553-1554); otherwise bounds stay as given
    const double *Db = D_host + z + l;
    for (int j = 0; j < bsize - 1; ++j) {
      const double factor = Db[j + 1] / Db[0];
      bu[j] = (bu[j] >= 1e15) ? INFINITY : bu[j] * factor;
      bl[j] = (bl[j] <= -1e15) ? -INFINITY : bl[j] * factor;
    }
  }
  const double one = 1.0;  // box_t_warm_start = 1, cones.c:1552
  if (h2d(*c, d_bu, bu.data(), (size_t)bsize - 1) || h2d(*c, d_bl, bl.data(), (size_t)bsize - 1) ||
      h2d(*c, box_t, &one, 1) || c->sync())
    return -1;
  return 0;
}

int ConeDev::project_nonlinear(double *x, const double *sv, const double *r) {
  cudaStream_t st = c->stream;
  if (bsize > 0) {
    k_box_cone<<<1, kBoxThreads, 0, st>>>(x + off_box, sv + off_box, r + off_box, d_bl, d_bu, bsize, box_t);
    c->launches++;
  }
  if (n_q_small > 0) {
    int grid = (n_q_small + (kThreads / 32) - 1) / (kThreads / 32);
    if (grid > c->grid_ew()) grid = c->grid_ew();
    k_soc_small<<<grid, kThreads, 0, st>>>(x, sv, r, q_off, q_len, n_q_small);
    c->launches++;
  }
  if (n_q_large > 0) {
    int grid = n_q_large < c->grid_ew() ? n_q_large : c->grid_ew();
    k_soc_large<<<grid, kThreads, 0, st>>>(x, sv, r, q_off + n_q_small, q_len + n_q_small, n_q_large);
    c->launches++;
  }
  if (n_psd > 0) {
    k_psd_cones<<<n_psd, kPsdThreads, 0, st>>>(x, sv, r, psd, psd_G, psd_V, psd_lam);
    c->launches++;
  }
  if (ep + ed > 0) {
    const int nt = ep + ed;
    int grid = (nt + 127) / 128;
    if (grid > c->grid_ew()) grid = c->grid_ew();
    k_exp_cones<<<grid, 128, 0, st>>>(x + off_exp, sv + off_exp, r + off_exp, ep, nt);
    c->launches++;
  }
  if (!p.empty()) {
    const int nt = (int)p.size();
    int grid = (nt + 127) / 128;
    if (grid > c->grid_ew()) grid = c->grid_ew();
    k_pow_cones<<<grid, 128, 0, st>>>(x + off_pow, sv + off_pow, r + off_pow, d_p, nt);
    c->launches++;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ConeDev::enforce_boundaries(double *D, int use_mean) {
  if (n_bnd == 0) return 0;
  int grid = (n_bnd + (kThreads / 32) - 1) / (kThreads / 32);
  if (grid > c->grid_ew()) grid = c->grid_ew();
  k_enforce_boundaries<<<grid, kThreads, 0, c->stream>>>(D, bnd_off, bnd_len, n_bnd, use_mean);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// generic Moreau pre-pass used by the host-buffer test surface: sv = x ; rows < z+l are
// finished in place, the rest become -r*x.
__global__ void __launch_bounds__(kThreads)
k_cone_pre(double *__restrict__ x, double *__restrict__ sv, const double *__restrict__ r, int m, int z, int zl) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const double s = x[i], ri = r[i];
    sv[i] = s;
    x[i] = (i < zl) ? zl_moreau(i, z, s, ri) : -ri * s;
  }
}

int current_device();

}  // namespace b200

// ===================================================== C ABI: SCS(proj_dual_cone) =====
using namespace b200;

struct SCS_B200_CONE_WORK {
  Ctx ctx;
  ConeDev cone;
  double *x = nullptr, *sv = nullptr, *r = nullptr;
};

extern "C" ScsB200ConeWork *scs_b200_init_cone(const ScsCone *k, scs_int m) {
  if (!k || m <= 0) return nullptr;
  SCS_B200_CONE_WORK *w = new SCS_B200_CONE_WORK();
  if (w->ctx.init(current_device()) || w->cone.init(&w->ctx, k, m) || dev_alloc(&w->x, (size_t)m) ||
      dev_alloc(&w->sv, (size_t)m) || dev_alloc(&w->r, (size_t)m)) {
    scs_b200_finish_cone(w);
    return nullptr;
  }
  return w;
}

extern "C" scs_int scs_b200_proj_dual_cone(scs_float *x, ScsB200ConeWork *w, const scs_float *D,
                                           const scs_float *r_y) {
  if (!w || !x) return -1;
  Ctx &c = w->ctx;
  if (cudaSetDevice(c.device) != cudaSuccess) return -1;
  const int m = w->cone.m;
  if (w->cone.normalize_box(D)) return -1;
  std::vector<double> ones;
  if (!r_y) { ones.assign((size_t)m, 1.0); r_y = ones.data(); }
  if (h2d(c, w->x, x, (size_t)m) || h2d(c, w->r, r_y, (size_t)m)) return -1;
  int grid = (m + kThreads - 1) / kThreads;
  if (grid > c.grid_ew()) grid = c.grid_ew();
  k_cone_pre<<<grid, kThreads, 0, c.stream>>>(w->x, w->sv, w->r, m, w->cone.z, w->cone.z + w->cone.l);
  c.launches++;
  if (w->cone.project_nonlinear(w->x, w->sv, w->r)) return -1;
  if (d2h(c, x, w->x, (size_t)m) || c.sync()) return -1;
  return 0;
}

extern "C" void scs_b200_finish_cone(ScsB200ConeWork *w) {
  if (!w) return;
  cudaSetDevice(w->ctx.device);
  if (w->ctx.stream) cudaStreamSynchronize(w->ctx.stream);
  w->cone.destroy();
  dev_free(w->x); dev_free(w->sv); dev_free(w->r);
  w->ctx.destroy();
  delete w;
}
