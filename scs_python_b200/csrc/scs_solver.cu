// scs_solver.cu -- device-resident ADMM loop behind the public SCS C API.
//
// Restates the host control flow of S/src/scs.c (scs_init / scs_update / scs_solve /
// scs_finish, scs.c:1193-1496) around a loop body in which every vector lives in HBM:
//
//   k_prep      normalize_v + v_prev copy + RHS build + CG warm start + CG tolerance
//               (scs.c:1315-1324, 696-719)
//   LinSys      PCG (linsys.cu)
//   k_rootplus  the five R-weighted dot products and the tau root (scs.c:667-688)
//   k_pre       u_t -= tau g ; u = 2 u_t - v ; zero/nonneg cones fully, others pre-scaled
//               (scs.c:727, 754-768 ; cones.c:1562-1571)
//   ConeDev     box / SOC / PSD / exp / power kernels with the Moreau recombination fused
//   k_post      rsk = R (v + u - 2 u_t) ; v += alpha (u - u_t) ; sum v^2 for the next
//               normalize_v (scs.c:739-751)
//   residuals   two SpMV passes with fat epilogues -> ~20 scalars, D2H every 25 iterations
//               (scs.c:513-585, 465-509)
//   AaDev       Anderson acceleration (aa.cu)
//
// Only the residual scalar block and one CG stop flag per iteration cross to the host.
#include <algorithm>
#include <chrono>
#include <csignal>
#include <mutex>
#include <string>

#include "aa.cuh"
#include "cones.cuh"
#include "dist.cuh"
#include "linsys.cuh"

namespace b200 {

int current_device();

// ------------------------------------------------------------------ constants ---------
// S/include/glbopts.h:35-50,184-257
#define B200_SCS_VERSION "3.2.11"
constexpr int kFeasibleIters = 1, kRescalingMinIters = 100, kConvergedInterval = 25, kPrintInterval = 250;
constexpr double kDivEps = 1e-18, kTauFactor = 10.0, kInfeasNegTol = 1e-9;
constexpr double kMaxScale = 1e6, kMinScale = 1e-6, kCgBestTol = 1e-12, kCgTolFactor = 0.2, kCgRate = 1.5;
constexpr double kMinNormFactor = 1e-4, kMaxNormFactor = 1e4;
constexpr int kRuizPasses = 25, kL2Passes = 1;
constexpr double kAaSafeguard = 1.0, kAaMaxWeight = 1e10;
constexpr int kAaIrSteps = 5;

static inline double safediv_pos_h(double x, double y) { return y < kDivEps ? x / kDivEps : x / y; }

static inline int ew_grid(const Ctx &c, long long len) {
  long long g = (len + kThreads - 1) / kThreads;
  if (g < 1) g = 1;
  return (int)(g < c.grid_ew() ? g : c.grid_ew());
}

// ------------------------------------------------------------ SIGINT (ctrlc.c) --------
// The handler flag is the only process-wide state (as in the reference, whose ctrlc.c guards it with a
// mutex).  Workspaces solve concurrently from different threads (scsobject.h:984-987 releases the GIL), so
// installing / restoring the handler is reference-counted under a mutex, and the flag is only cleared by
// the solve that installs the handler -- never while another solve is listening.
static volatile sig_atomic_t g_int_detected = 0;
static struct sigaction g_old_action;
static int g_listener_depth = 0;
static std::mutex g_listener_mu;
static void b200_sigint_handler(int) { g_int_detected = 1; }
static void start_interrupt_listener() {
  std::lock_guard<std::mutex> lk(g_listener_mu);
  if (g_listener_depth++ == 0) {
    struct sigaction act;
    g_int_detected = 0;
    act.sa_flags = 0;
    sigemptyset(&act.sa_mask);
    act.sa_handler = b200_sigint_handler;
    sigaction(SIGINT, &act, &g_old_action);
  }
}
static void end_interrupt_listener() {
  std::lock_guard<std::mutex> lk(g_listener_mu);
  if (g_listener_depth > 0 && --g_listener_depth == 0) {
    struct sigaction act;
    sigaction(SIGINT, &g_old_action, &act);
  }
}

// ------------------------------------------------------------------ kernels -----------
// diag_r = [rho_x 1_n ; r_y ; TAU_FACTOR]   (scs.c:929-938, cones.c:349-363)
__global__ void __launch_bounds__(kThreads)
k_set_diag_r(double *__restrict__ R, int n, int m, int z, double rho_x, double scale) {
  const int l = n + m + 1;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) {
    double v;
    if (j < n) v = rho_x;
    else if (j < n + z) v = 1.0 / (1000.0 * scale);
    else if (j < n + m) v = 1.0 / scale;
    else v = kTauFactor;
    R[j] = v;
  }
}

__global__ void __launch_bounds__(kThreads) k_fill(double *__restrict__ x, double v, long long len) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < len; j += (long long)gridDim.x * blockDim.x)
    x[j] = v;
}

// x <- 1 / sqrt(limit(pre(x)))  (apply_limit + SQRTF + SAFEDIV_POS, scs_matrix.c:203-208,235-236)
__global__ void __launch_bounds__(kThreads) k_inv_sqrt_limit(double *__restrict__ x, int len, int pre_sqrt) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) {
    double v = x[j];
    if (pre_sqrt) v = sqrt(v);
    v = v < kMinNormFactor ? 1.0 : v;
    v = v > kMaxNormFactor ? kMaxNormFactor : v;
    v = sqrt(v);
    x[j] = v < kDivEps ? 1.0 / kDivEps : 1.0 / v;
  }
}
__global__ void __launch_bounds__(kThreads) k_sqrt(double *__restrict__ x, int len) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) x[j] = sqrt(x[j]);
}
__global__ void __launch_bounds__(kThreads)
k_mul_inplace(double *__restrict__ a, const double *__restrict__ b, int len) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) a[j] *= b[j];
}

// rescale functors (scs_matrix.c:344-364): the same expression on both copies of A
struct RescaleA {   // CSR(A): row = i, col = j
  const double *Dt, *Et;
  __device__ __forceinline__ double operator()(int row, int col, double v) const { return v * (Dt[row] * Et[col]); }
};
struct RescaleAt {  // CSR(A'): row = j, col = i
  const double *Dt, *Et;
  __device__ __forceinline__ double operator()(int row, int col, double v) const { return v * (Dt[col] * Et[row]); }
};
struct RescaleP {
  const double *Et;
  __device__ __forceinline__ double operator()(int row, int col, double v) const { return v * (Et[row] * Et[col]); }
};
// Ruiz / L2 column pass over (A', P): separate accumulators are not needed, max and sum
// of squares both combine across the two matrices
struct EpiStoreComb : EpiNoState {
  double *y;
  __device__ __forceinline__ void row(State &, int r, double acc) const { y[r] = acc; }
};

struct FinSigma {  // normalize.c:46-52
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    double sigma = fmax(o[0], o[1]);
    sigma = sigma < kMinNormFactor ? 1.0 : sigma;
    sigma = sigma > kMaxNormFactor ? kMaxNormFactor : sigma;
    S->sigma = sigma < kDivEps ? 1.0 / kDivEps : 1.0 / sigma;
  }
};
// b *= D ; c *= E ; sigma from the inf-norms (normalize.c:33-52)
__global__ void __launch_bounds__(kThreads)
k_scale_bc(double *__restrict__ b, const double *__restrict__ D, int m, double *__restrict__ c,
           const double *__restrict__ E, int n, RedWs ws, DevScalars *S) {
  double v[2] = {0.0, 0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m + n; j += gridDim.x * blockDim.x) {
    if (j < m) {
      const double t = b[j] * D[j];
      b[j] = t;
      v[0] = fmax(v[0], fabs(t));
    } else {
      const double t = c[j - m] * E[j - m];
      c[j - m] = t;
      v[1] = fmax(v[1], fabs(t));
    }
  }
  grid_reduce_fin<0, 2>(v, ws, S, FinSigma{});
}
__global__ void __launch_bounds__(kThreads)
k_scale_sigma(double *__restrict__ b, int m, double *__restrict__ c, int n, const DevScalars *S) {
  const double sg = S->sigma;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m + n; j += gridDim.x * blockDim.x) {
    if (j < m) b[j] *= sg; else c[j - m] *= sg;
  }
}

// g = [c ; -b]   (scs.c:1066-1074)
__global__ void __launch_bounds__(kThreads)
k_build_g(double *__restrict__ g, const double *__restrict__ c, const double *__restrict__ b, int n, int m) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n + m; j += gridDim.x * blockDim.x)
    g[j] = j < n ? c[j] : -b[j - n];
}

// cold start v = [0 ; 0 ; 1]  (scs.c:659-663)
__global__ void __launch_bounds__(kThreads) k_cold_start(double *__restrict__ v, int l) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) v[j] = (j == l - 1) ? 1.0 : 0.0;
}
// warm start v = [x ; y + s / R_y ; 1] after normalising (x,y,s)  (scs.c:638-657, normalize.c:64-76)
__global__ void __launch_bounds__(kThreads)
k_warm_start(double *__restrict__ v, const double *__restrict__ x, const double *__restrict__ y,
             const double *__restrict__ s, const double *__restrict__ D, const double *__restrict__ E,
             const double *__restrict__ R, int n, int m, double primal_scale, double dual_scale) {
  const int l = n + m + 1;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) {
    double val;
    if (j < n) {
      val = x[j] / (E[j] / dual_scale);
    } else if (j < n + m) {
      const int i = j - n;
      const double yn = y[i] / (D[i] / primal_scale);
      const double sn = s[i] * (D[i] * dual_scale);
      val = yn + sn / R[j];
    } else {
      val = 1.0;
    }
    v[j] = (val != val) ? 0.0 : val;
  }
}

struct FinPrep {  // CG tolerance rule, scs.c:706-719 ; zero right-hand side, private.c:288-291
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    double tol = fmin(S->nm_ax_s_btau, S->nm_px_aty_ctau);
    const double nm_ws = o[0] / pow((double)S->iter + 1.0, kCgRate);
    tol = kCgTolFactor * fmin(tol, nm_ws);
    S->cg_tol = fmax(kCgBestTol, tol);
    S->zero_rhs = (o[1] <= 1e-12) ? 1 : 0;
    S->cg_done = S->zero_rhs;
    S->cg_its = 0;
  }
};
// normalize_v, v_prev copy, RHS, warm start, CG tolerance and flags
template <bool ACCEL>
__global__ void __launch_bounds__(kThreads)
k_prep(double *__restrict__ v, double *__restrict__ v_prev, double *__restrict__ u_t, const double *__restrict__ u,
       const double *__restrict__ g, const double *__restrict__ R, double *__restrict__ ws, int n, int m,
       int l_total, RedWs red, DevScalars *S) {
  const int l = n + m + 1;
  const int it = S->iter;
  if (blockIdx.x == 0 && threadIdx.x == 0) phase_lap(S, -1);  // lin-sys interval starts
  double sc = 1.0;
  if (it >= kFeasibleIters) {
    const double vn = sqrt(S->vnorm2);
    // ITERATE_NORM == 1; l_total = length of the whole iterate (== l unless row-partitioned)
    if (vn != 0.0) sc = sqrt((double)l_total) / vn;
  }
  const double tau_u = u[l - 1];
  double mx[2] = {0.0, 0.0};  // ||ws||_inf, ||rhs||_inf
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) {
    const double vj = v[j] * sc;
    v[j] = vj;
    if (ACCEL) v_prev[j] = vj;
    if (j < n) {
      const double ut = vj * R[j];
      u_t[j] = ut;
      const double w = fma(tau_u, g[j], u[j]);
      ws[j] = w;
      mx[0] = fmax(mx[0], fabs(w));
      mx[1] = fmax(mx[1], fabs(ut));
    } else if (j < l - 1) {
      const double ut = -vj * R[j];
      u_t[j] = ut;
      mx[1] = fmax(mx[1], fabs(ut));
    } else {
      u_t[j] = vj;
    }
  }
  grid_reduce_fin<0, 2>(mx, red, S, FinPrep{});
}

struct FinRootPlus {  // the quadratic of scs.c:676-687
  const double *eta_ptr, *tau_scale_ptr;  // v[l-1], diag_r[l-1]
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    const double eta = *eta_ptr, tau_scale = *tau_scale_ptr;
    if (S->iter < kFeasibleIters) {
      S->tau = 1.0;
    } else {
      const double a = tau_scale + o[0];
      const double b = o[1] - 2 * o[2] - eta * tau_scale;
      const double c = o[3] - o[4];
      const double rad = b * b - 4 * a * c;
      S->tau = (-b + sqrt(fmax(rad, 0.0))) / (2 * a);
    }
    phase_lap(S, 0);  // lin-sys interval ends, cone interval starts
  }
};
// root_plus (scs.c:667-688): p = u_t (after the solve), mu = v, eta = v[l-1]
__global__ void __launch_bounds__(kThreads)
k_rootplus(const double *__restrict__ u_t, const double *__restrict__ v, const double *__restrict__ g,
           const double *__restrict__ R, int n, int nm, int cnt_lo, RedWs red, DevScalars *S) {
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  // cnt_lo > 0: rank > 0 of a row-partitioned solve, the replicated shared block [0, cnt_lo) is summed by rank 0
  (void)n;
  for (int j = cnt_lo + blockIdx.x * blockDim.x + threadIdx.x; j < nm; j += gridDim.x * blockDim.x) {
    const double ri = R[j], gi = g[j], pi = u_t[j], mui = v[j];
    s[0] = fma(gi * gi, ri, s[0]);
    s[1] = fma(mui * gi, ri, s[1]);
    s[2] = fma(pi * gi, ri, s[2]);
    s[3] = fma(pi * pi, ri, s[3]);
    s[4] = fma(pi * mui, ri, s[4]);
  }
  grid_reduce_fin<5, 0>(s, red, S, FinRootPlus{v + nm, R + nm});
}

// u_t -= tau g ; u = 2 u_t - v ; zero / nonneg rows done, other rows pre-scaled by -R with
// s saved in rsk (scratch until k_post overwrites it)
__global__ void __launch_bounds__(kThreads)
k_pre(double *__restrict__ u_t, double *__restrict__ u, double *__restrict__ rsk, const double *__restrict__ v,
      const double *__restrict__ g, const double *__restrict__ R, int n, int m, int z, int zl,
      const DevScalars *S) {
  const int l = n + m + 1;
  const int it = S->iter;
  const double tau = S->tau;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) {
    const double ut = (j < l - 1) ? fma(-tau, g[j], u_t[j]) : tau;
    u_t[j] = ut;
    const double s = 2 * ut - v[j];
    double out;
    if (j < n) {
      out = s;
    } else if (j < l - 1) {
      const int row = j - n;
      if (row < zl) {
        out = zl_moreau(row, z, s, R[j]);
      } else {
        out = -R[j] * s;
        rsk[j] = s;
      }
    } else {
      out = (it < kFeasibleIters) ? 1.0 : fmax(s, 0.0);
    }
    u[j] = out;
  }
}

struct FinPost {
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    S->vnorm2 = o[0];
    S->iter += 1;  // the iteration index lives on the device: the graph re-launches unchanged
  }
};
// MODE 0: rsk and dual step ; 1: rsk only ; 2: dual step only   (scs.c:739-751)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
k_post(double *__restrict__ v, double *__restrict__ rsk, const double *__restrict__ u, const double *__restrict__ u_t,
       const double *__restrict__ R, int n, int l, int cnt_lo, double alpha, RedWs red, DevScalars *S) {
  (void)n;
  if (MODE != 2 && blockIdx.x == 0 && threadIdx.x == 0) phase_lap(S, 1);  // cone interval ends
  double s[1] = {0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) {
    const double vj = v[j], uj = u[j], utj = u_t[j];
    if (MODE != 2) rsk[j] = (vj + uj - 2 * utj) * R[j];
    if (MODE != 1) {
      const double vn = fma(alpha, uj - utj, vj);
      v[j] = vn;
      // cnt_lo > 0: the replicated shared block [0, cnt_lo) and tau are summed by rank 0 only
      if (cnt_lo == 0 || (j >= cnt_lo && j < l - 1)) s[0] = fma(vn, vn, s[0]);
    }
  }
  if (MODE != 1) grid_reduce_fin<1, 0>(s, red, S, FinPost{});
}

// phase clock boundary around the acceleration kernels: slot < 0 opens, slot >= 0 closes
__global__ void k_stamp(DevScalars *S, int slot) { phase_lap(S, slot); }

// v <- rsk / R+ + 2 u_t - u after a scale update (scs.c:1180-1186)
__global__ void __launch_bounds__(kThreads)
k_remap_v(double *__restrict__ v, const double *__restrict__ rsk, const double *__restrict__ u,
          const double *__restrict__ u_t, const double *__restrict__ R, int l) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x)
    v[j] = rsk[j] / R[j] + 2 * u_t[j] - u[j];
}

// ---- residual epilogues (scs.c:513-585 and 465-509) ----
struct FinResA {
  const double *tau_ptr, *kap_ptr;
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    S->res[R_TAU] = fabs(*tau_ptr);
    S->res[R_KAP] = fabs(*kap_ptr);
    S->res[R_BTY_TAU] = o[0];
    S->res[R_NM_AX_S_BTAU] = o[1];
    S->res[R_NM_AX_S] = o[2];
    S->res[R_NM_AX] = o[3];
    S->res[R_ONM_AX_S_BTAU] = o[4];
    S->res[R_ONM_AX_S] = o[5];
    S->res[R_ONM_AX] = o[6];
    S->res[R_ONM_S] = o[7];
    S->nm_ax_s_btau = o[1];
  }
};
// primal pass over CSR(A), gather x = u[0:n]
struct EpiResA {
  static constexpr bool kSeparate = false;
  struct State { double bty, m[7]; };
  const double *u, *rsk, *b, *D;
  int n, m_;
  double inv_ds, dual_scale;
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  __device__ __forceinline__ void init(State &s) const {
    s.bty = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) s.m[k] = 0.0;
  }
  __device__ __forceinline__ void row(State &st, int i, double ax) const {
    const double tau = fabs(u[n + m_]);
    const double s = rsk[n + i], bi = b[i], yi = u[n + i], Di = D[i];
    const double axs = ax + s;
    const double axsb = axs - tau * bi;
    const double f = inv_ds / Di;
    st.bty = fma(yi, bi, st.bty);
    st.m[0] = fmax(st.m[0], fabs(axsb));
    st.m[1] = fmax(st.m[1], fabs(axs));
    st.m[2] = fmax(st.m[2], fabs(ax));
    st.m[3] = fmax(st.m[3], fabs(axsb * f));
    st.m[4] = fmax(st.m[4], fabs(axs * f));
    st.m[5] = fmax(st.m[5], fabs(ax * f));
    st.m[6] = fmax(st.m[6], fabs(s / (Di * dual_scale)));
  }
  __device__ __forceinline__ void finish(State &st, const RedWs &ws, DevScalars *S) const {
    double v[8] = {st.bty, st.m[0], st.m[1], st.m[2], st.m[3], st.m[4], st.m[5], st.m[6]};
    grid_reduce_fin<1, 7>(v, ws, S, FinResA{u + n + m_, rsk + n + m_});
  }
};
struct FinResAt {
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    S->res[R_XPX_TAU] = o[0];
    S->res[R_CTX_TAU] = o[1];
    S->res[R_NM_PX_ATY_CTAU] = o[2];
    S->res[R_NM_PX] = o[3];
    S->res[R_NM_ATY] = o[4];
    S->res[R_ONM_PX_ATY_CTAU] = o[5];
    S->res[R_ONM_PX] = o[6];
    S->res[R_ONM_ATY] = o[7];
    S->nm_px_aty_ctau = o[2];
  }
};
// dual pass over (CSR(A'), CSR(P)), gathers y = u[n:] and x = u[0:n]
struct EpiResAt {
  static constexpr bool kSeparate = true;
  struct State { double xpx, ctx, m[6]; };
  const double *u, *c, *E;
  int n, m_;
  double inv_ps;
  const double *extra;  // row-partitioned mode: all-reduced A'y (the kernel then runs over P only), else null
  int cnt_lo = 0;       // row-partitioned mode: the sums count columns >= cnt_lo (shared block on rank 0 only)
  __device__ __forceinline__ void init(State &s) const {
    s.xpx = 0.0; s.ctx = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s.m[k] = 0.0;
  }
  __device__ __forceinline__ void row(State &st, int j, double acc) const {
    if (extra) row2(st, j, extra[j], acc);  // acc = (P x)_j
    else row2(st, j, acc, 0.0);             // acc = (A'y)_j, no P
  }
  __device__ __forceinline__ void row2(State &st, int j, double aty, double px) const {
    const double tau = fabs(u[n + m_]);
    const double xj = u[j], cj = c[j];
    const double pac = px + aty + tau * cj;
    const double f = inv_ps / E[j];
    if (j >= cnt_lo) {
      st.xpx = fma(px, xj, st.xpx);
      st.ctx = fma(xj, cj, st.ctx);
    }
    st.m[0] = fmax(st.m[0], fabs(pac));
    st.m[1] = fmax(st.m[1], fabs(px));
    st.m[2] = fmax(st.m[2], fabs(aty));
    st.m[3] = fmax(st.m[3], fabs(pac * f));
    st.m[4] = fmax(st.m[4], fabs(px * f));
    st.m[5] = fmax(st.m[5], fabs(aty * f));
  }
  __device__ __forceinline__ void finish(State &st, const RedWs &ws, DevScalars *S) const {
    double v[8] = {st.xpx, st.ctx, st.m[0], st.m[1], st.m[2], st.m[3], st.m[4], st.m[5]};
    grid_reduce_fin<2, 6>(v, ws, S, FinResAt{});
  }
};

struct FinSol {
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    S->fin[2] = o[0];
    S->fin[0] = o[1];
    S->fin[1] = o[2];
  }
};
// un-normalised solution (normalize.c:78-90) + ||s||_inf, ||y||_inf, s'y (scs.c:885-887)
__global__ void __launch_bounds__(kThreads)
k_finalize_sol(double *__restrict__ xo, double *__restrict__ yo, double *__restrict__ so, const double *__restrict__ u,
               const double *__restrict__ rsk, const double *__restrict__ D, const double *__restrict__ E, int n, int m,
               double primal_scale, double dual_scale, RedWs red, DevScalars *S) {
  double v[3] = {0.0, 0.0, 0.0};  // s'y | max s, max y
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n + m; j += gridDim.x * blockDim.x) {
    if (j < n) {
      xo[j] = u[j] * (E[j] / dual_scale);
    } else {
      const int i = j - n;
      const double y = u[j] * (D[i] / primal_scale);
      const double s = rsk[j] / (D[i] * dual_scale);
      yo[i] = y;
      so[i] = s;
      v[0] = fma(s, y, v[0]);
      v[1] = fmax(v[1], fabs(s));
      v[2] = fmax(v[2], fabs(y));
    }
  }
  grid_reduce_fin<1, 2>(v, red, S, FinSol{});
}
// row-partitioned mode: full[loc2glob[j]] = x[j] for the columns this rank counts
__global__ void __launch_bounds__(kThreads)
k_scatter_cols(double *__restrict__ full, const double *__restrict__ x, const int *__restrict__ loc2glob, int lo, int n) {
  for (int j = lo + blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) full[loc2glob[j]] = x[j];
}
__global__ void __launch_bounds__(kThreads)
k_scale3(double *__restrict__ x, int n, double fx, double *__restrict__ y, double *__restrict__ s, int m, double fy,
         double fs) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n + m; j += gridDim.x * blockDim.x) {
    if (j < n) x[j] *= fx;
    else { y[j - n] *= fy; s[j - n] *= fs; }
  }
}

// ---- CSV trace (SCS(log_data_to_csv), rw.c:317-476): the vector norms of one row that the residual
// epilogues do not already deliver.  ax = A x (m), atp = A'y + P x (n), both in normalised scaling.
// out (S->part): 12 sums of squares then 8 maxima, see csv_log().
__global__ void __launch_bounds__(kThreads)
k_csv_norms(const double *__restrict__ u, const double *__restrict__ u_t, const double *__restrict__ v,
            const double *__restrict__ v_prev, const double *__restrict__ rsk, const double *__restrict__ ax,
            const double *__restrict__ atp, const double *__restrict__ b, const double *__restrict__ c,
            const double *__restrict__ D, const double *__restrict__ E, int n, int m, double primal_scale,
            double dual_scale, RedWs red, DevScalars *S) {
  double o[20];
#pragma unroll
  for (int k = 0; k < 20; ++k) o[k] = 0.0;
  const int l = n + m + 1;
  const double tau = fabs(u[n + m]);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < l; j += gridDim.x * blockDim.x) {
    const double du = u[j] - u_t[j], dv = v[j] - v_prev[j];
    o[10] = fma(du, du, o[10]); o[11] = fma(dv, dv, o[11]);
    o[18] = fmax(o[18], fabs(du)); o[19] = fmax(o[19], fabs(dv));
    if (j < n) {
      const double x = u[j], xo = x * (E[j] / dual_scale);
      const double r = atp[j] + tau * c[j], ro = r / (E[j] * primal_scale);
      o[0] = fma(x, x, o[0]); o[3] = fma(xo, xo, o[3]);
      o[12] = fmax(o[12], fabs(x)); o[15] = fmax(o[15], fabs(xo));
      o[7] = fma(r, r, o[7]); o[9] = fma(ro, ro, o[9]);
    } else if (j < n + m) {
      const int i = j - n;
      const double y = u[j], s = rsk[j];
      const double yo = y * (D[i] / primal_scale), so = s / (D[i] * dual_scale);
      const double r = ax[i] + s - tau * b[i], ro = r / (D[i] * dual_scale);
      o[1] = fma(y, y, o[1]); o[2] = fma(s, s, o[2]); o[4] = fma(yo, yo, o[4]); o[5] = fma(so, so, o[5]);
      o[13] = fmax(o[13], fabs(y)); o[14] = fmax(o[14], fabs(s));
      o[16] = fmax(o[16], fabs(yo)); o[17] = fmax(o[17], fabs(so));
      o[6] = fma(r, r, o[6]); o[8] = fma(ro, ro, o[8]);
    }
  }
  grid_reduce<12, 8>(o, red, [S](double *q) {
#pragma unroll
    for (int k = 0; k < 20; ++k) S->part[k] = q[k];
  });
}

// ---------------------------------------------------------- host-side structures -------
struct HostResid {  // ScsResiduals scalars (scs_work.h:29-50); vectors are reduced on the device
  int last_iter = -1;
  double xt_p_x = 0, xt_p_x_tau = 0, ctx = 0, ctx_tau = 0, bty = 0, bty_tau = 0, pobj = 0, dobj = 0, gap = 0;
  double tau = 0, kap = 0, res_pri = 0, res_dual = 0, res_infeas = NAN, res_unbdd_p = NAN, res_unbdd_a = NAN;
  double nm_ax_s_btau = 0, nm_ax_s = 0, nm_ax = 0, nm_px_aty_ctau = 0, nm_px = 0, nm_aty = 0, nm_s = 0;
};

// compute_residuals, scs.c:441-463
static void compute_residuals_h(HostResid &r, double pd) {
  const double tol = kInfeasNegTol / pd;
  r.res_pri = safediv_pos_h(r.nm_ax_s_btau, r.tau);
  r.res_dual = safediv_pos_h(r.nm_px_aty_ctau, r.tau);
  r.res_unbdd_a = NAN; r.res_unbdd_p = NAN; r.res_infeas = NAN;
  if (r.ctx_tau < -tol) {
    r.res_unbdd_a = safediv_pos_h(r.nm_ax_s, -r.ctx_tau);
    r.res_unbdd_p = safediv_pos_h(r.nm_px, -r.ctx_tau);
  }
  if (r.bty_tau < -tol) r.res_infeas = safediv_pos_h(r.nm_aty, -r.bty_tau);
}

}  // namespace b200

using namespace b200;

struct SCS_WORK {
  Ctx c;
  LinSys ls;
  ConeDev cone;
  AaDev aa;
  bool has_aa = false;
  int n = 0, m = 0, l = 0;
  // the graph-launched front half of one ADMM iteration (k_prep .. cones), CG loop = WHILE node
  bool use_graph = false;
  // CG iterations captured as plain kernel nodes in front of the WHILE node, and whether that number follows the
  // observed CG iterations per ADMM iteration (re-capture at residual checks).  Off by default: measured on
  // Cfg-1 / Cfg-3 the conditional node costs nothing that unrolling saves (profiles/r2g_configs_unroll*.jsonl), and the
  // re-captures cost ~30 ms on the cone LP.  SCS_B200_CG_UNROLL=k fixes k; SCS_B200_CG_UNROLL=auto turns following on.
  int cg_unroll = 0;
  bool cg_unroll_auto = false;
  long long unroll_cg0 = 0; int unroll_it0 = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t st_body = nullptr;
  long long graph_launches = 0, graph_spmv = 0;  // fixed (non-CG-loop) kernels per graph launch
  long long n_checks = 0, n_aa = 0;               // residual checks / AA applications so far
  // row-partitioned mode (dist.cuh): this rank owns rows [row0, row0 + m) of the m_total rows
  // and the columns [shared (n_sh, replicated) | private to this rank] of the n_total columns
  bool dist = false;
  bool print_rank0 = false, verbose_collective = false;  // dist mode: who prints / whether the verbose-only residual passes run
  int rank = 0, world = 1, m_total = 0, row0 = 0, n_total = 0, n_sh = 0;
  int cnt_lo = 0;  // reductions over n-space count [cnt_lo, n): 0 on rank 0, n_sh elsewhere (shared block counted once)
  std::vector<int> loc2glob;   // global index of every local column
  int *d_loc2glob = nullptr;
  double *gather = nullptr;  // max(m_total, n_total) doubles: assembling the full x / y / s on every rank
  ScsSettings stgs;
  std::string write_fn, csv_fn;
  // device state
  double *u = nullptr, *u_t = nullptr, *v = nullptr, *v_prev = nullptr, *rsk = nullptr, *g = nullptr;
  double *diag_r = nullptr, *b = nullptr, *cvec = nullptr, *D = nullptr, *E = nullptr, *ws = nullptr;
  double *sol_x = nullptr, *sol_y = nullptr, *sol_s = nullptr;
  double *csv_m = nullptr, *csv_n = nullptr;  // A x and A'y + P x of the CSV trace (allocated on first use)
  // host
  std::vector<double> b_orig, c_orig, c_loc, x_loc;
  double nm_b_orig = 0, nm_c_orig = 0, primal_scale = 1, dual_scale = 1;
  HostResid r_n, r_o;
  double setup_time = 0;
  int time_limit_reached = 0;
  double sum_log_scale_factor = 0;
  int last_scale_update_iter = 0, n_log_scale_factor = 0, scale_updates = 0;
  long long admm_iters = 0;
  // iteration marks (bench.py)
  int mark_begin = -1, mark_end = -1;
  cudaEvent_t mark_ev[2] = {nullptr, nullptr};
  ScsB200Marks marks;
  long long mk_launch0 = 0, mk_cg0 = 0;
  double mk_bytes0 = 0;
  int mk_iter0 = 0;
  bool mk_open = false, mk_done = false;
};

namespace b200 {

using Clock = std::chrono::steady_clock;
static inline double ms_since(const Clock::time_point &t0) {
  return std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
}

// ------------------------------------------------------------------ validation --------
int validate_problem(const ScsData *d, const ScsCone *k, const ScsSettings *stgs) {  // scs.c:364-429
  if (d->m <= 0 || d->n <= 0) {
    B200_PRINTF("m and n must both be greater than 0; m = %li, n = %li\n", (long)d->m, (long)d->n);
    return -1;
  }
  const ScsMatrix *A = d->A, *P = d->P;  // validate_lin_sys, scs_matrix.c:65-131
  if (!A) { B200_PRINTF("A matrix missing\n"); return -1; }
  if (!A->x || !A->i || !A->p) { B200_PRINTF("data incompletely specified\n"); return -1; }
  if (A->m != d->m || A->n != d->n) { B200_PRINTF("A dimensions inconsistent with m, n\n"); return -1; }
  const long long Anz = A->p[A->n];
  if (((double)Anz / A->m > A->n) || Anz < 0) {
    B200_PRINTF("Anz (nonzeros in A) = %li, outside of valid range\n", (long)Anz);
    return -1;
  }
  int r_max = 0;
  for (long long i = 0; i < Anz; ++i) {
    if (A->i[i] > r_max) r_max = A->i[i];
    if (A->i[i] < 0) { B200_PRINTF("negative row index in A\n"); return -1; }
  }
  if (r_max > A->m - 1) { B200_PRINTF("number of rows in A inconsistent with input dimension\n"); return -1; }
  if (P) {
    if (!P->x || !P->i || !P->p) { B200_PRINTF("P matrix incompletely specified\n"); return -1; }
    if (P->n != A->n) { B200_PRINTF("P dimension = %li, inconsistent with n = %li\n", (long)P->n, (long)A->n); return -1; }
    if (P->m != P->n) { B200_PRINTF("P is not square\n"); return -1; }
    for (int j = 0; j < P->n; ++j)
      for (int i = P->p[j]; i < P->p[j + 1]; ++i)
        if (P->i[i] > j) { B200_PRINTF("P is not upper triangular\n"); return -1; }
  }
  // validate_cones, cones.c:583-755
  if (k->z < 0) { B200_PRINTF("free cone dimension error\n"); return -1; }
  if (k->l < 0) { B200_PRINTF("lp cone dimension error\n"); return -1; }
  if (k->bsize < 0) { B200_PRINTF("box cone dimension error\n"); return -1; }
  if (k->bsize > 1) {
    if (!k->bl || !k->bu) { B200_PRINTF("box cone bounds missing\n"); return -1; }
    for (int i = 0; i < k->bsize - 1; ++i)
      if (k->bl[i] > k->bu[i]) { B200_PRINTF("infeasible: box lower bound larger than upper bound\n"); return -1; }
  }
  long long dims = (long long)k->z + k->l + k->bsize;
  if (k->qsize < 0 || (k->qsize > 0 && !k->q)) { B200_PRINTF("soc cone dimension error\n"); return -1; }
  for (int i = 0; i < k->qsize; ++i) {
    if (k->q[i] < 0) { B200_PRINTF("soc cone dimension error\n"); return -1; }
    dims += k->q[i];
  }
  if (k->ssize < 0 || (k->ssize > 0 && !k->s)) { B200_PRINTF("sd cone dimension error\n"); return -1; }
  for (int i = 0; i < k->ssize; ++i) {
    if (k->s[i] < 0) { B200_PRINTF("sd cone dimension error\n"); return -1; }
    dims += (long long)k->s[i] * (k->s[i] + 1) / 2;
  }
  if (k->cssize < 0 || (k->cssize > 0 && !k->cs)) { B200_PRINTF("complex psd cone dimension error\n"); return -1; }
  for (int i = 0; i < k->cssize; ++i) {
    if (k->cs[i] < 0) { B200_PRINTF("complex psd cone dimension error\n"); return -1; }
    dims += (long long)k->cs[i] * k->cs[i];
  }
  if (k->ed < 0) { B200_PRINTF("ed cone dimension error\n"); return -1; }
  if (k->ep < 0) { B200_PRINTF("ep cone dimension error\n"); return -1; }
  if (k->psize < 0 || (k->psize > 0 && !k->p)) { B200_PRINTF("power cone dimension error\n"); return -1; }
  for (int i = 0; i < k->psize; ++i)
    if (k->p[i] < -1 || k->p[i] > 1) { B200_PRINTF("power cone error, values must be in [-1,1]\n"); return -1; }
  dims += 3ll * (k->ed + k->ep + k->psize);
  if (dims != d->m) {
    B200_PRINTF("Error: Cone dims %li != rows in A %li\n", (long)dims, (long)d->m);
    return -1;
  }
  if (stgs->max_iters <= 0) { B200_PRINTF("max_iters must be positive\n"); return -1; }
  if (stgs->eps_abs < 0) { B200_PRINTF("eps_abs tolerance must be positive\n"); return -1; }
  if (stgs->eps_rel < 0) { B200_PRINTF("eps_rel tolerance must be positive\n"); return -1; }
  if (stgs->eps_infeas < 0) { B200_PRINTF("eps_infeas tolerance must be positive\n"); return -1; }
  if (stgs->alpha <= 0 || stgs->alpha >= 2) { B200_PRINTF("alpha must be in (0,2)\n"); return -1; }
  if (stgs->rho_x <= 0) { B200_PRINTF("rho_x must be positive (1e-3 works well).\n"); return -1; }
  if (stgs->scale <= 0) { B200_PRINTF("scale must be positive (1 works well).\n"); return -1; }
  if (stgs->acceleration_interval <= 0) { B200_PRINTF("acceleration_interval must be positive (10 works well).\n"); return -1; }
  if (stgs->acceleration_lookback < 0) {
    B200_PRINTF("acceleration_lookback must be nonnegative (use acceleration_type_1=0 for type-II AA).\n");
    return -1;
  }
  if (!std::isfinite(stgs->acceleration_regularization) || stgs->acceleration_regularization < 0) {
    B200_PRINTF("acceleration_regularization must be a nonnegative finite number.\n");
    return -1;
  }
  if (!std::isfinite(stgs->acceleration_relaxation) || stgs->acceleration_relaxation < 0 ||
      stgs->acceleration_relaxation > 2) {
    B200_PRINTF("acceleration_relaxation must be in [0, 2].\n");
    return -1;
  }
  return 0;
}

// ------------------------------------------------------------------ printing ----------
static void print_line() {
  char buf[68];
  memset(buf, '-', 66);
  buf[66] = '\n';
  buf[67] = 0;
  B200_PRINTF("%s", buf);
}

static void print_init_header(const ScsData *d, const ScsCone *k, const ScsSettings *s) {  // scs.c:123-177
  print_line();
  B200_PRINTF("\t       SCS v%s - Splitting Conic Solver\n\t(c) Brendan O'Donoghue, Stanford University, 2012\n",
              B200_SCS_VERSION);
  print_line();
  B200_PRINTF("problem:  variables n: %i, constraints m: %i\n", (int)d->n, (int)d->m);
  B200_PRINTF("cones: ");
  if (k->z) B200_PRINTF("\t  z: primal zero / dual free vars: %li\n", (long)k->z);
  if (k->l) B200_PRINTF("\t  l: linear vars: %li\n", (long)k->l);
  if (k->bsize) B200_PRINTF("\t  b: box cone vars: %li\n", (long)k->bsize);
  if (k->qsize) {
    long tot = 0;
    for (int i = 0; i < k->qsize; ++i) tot += k->q[i];
    B200_PRINTF("\t  q: soc vars: %li, qsize: %li\n", tot, (long)k->qsize);
  }
  if (k->ssize) {
    long tot = 0;
    for (int i = 0; i < k->ssize; ++i) tot += (long)k->s[i] * (k->s[i] + 1) / 2;
    B200_PRINTF("\t  s: psd vars: %li, ssize: %li\n", tot, (long)k->ssize);
  }
  if (k->cssize) {
    long tot = 0;
    for (int i = 0; i < k->cssize; ++i) tot += (long)k->cs[i] * k->cs[i];
    B200_PRINTF("\t  cs: complex psd vars: %li, ssize: %li\n", tot, (long)k->cssize);
  }
  if (k->ep || k->ed) B200_PRINTF("\t  e: exp vars: %li, dual exp vars: %li\n", 3l * k->ep, 3l * k->ed);
  if (k->psize) B200_PRINTF("\t  p: primal + dual power vars: %li\n", 3l * k->psize);
  B200_PRINTF("settings: eps_abs: %.1e, eps_rel: %.1e, eps_infeas: %.1e\n"
              "\t  alpha: %.2f, scale: %.2e, adaptive_scale: %i\n"
              "\t  max_iters: %i, normalize: %i, rho_x: %.2e\n",
              s->eps_abs, s->eps_rel, s->eps_infeas, s->alpha, s->scale, (int)s->adaptive_scale, (int)s->max_iters,
              (int)s->normalize, s->rho_x);
  if (s->acceleration_lookback != 0)
    B200_PRINTF("\t  acceleration_lookback: %i, acceleration_interval: %i\n", (int)s->acceleration_lookback,
                (int)s->acceleration_interval);
  if (s->time_limit_secs) B200_PRINTF("\t  time_limit_secs: %.2e\n", s->time_limit_secs);
  B200_PRINTF("lin-sys:  %s\n\t  nnz(A): %li, nnz(P): %li\n", scs_get_lin_sys_method(), (long)d->A->p[d->A->n],
              d->P ? (long)d->P->p[d->P->n] : 0l);
}

static void print_header() {  // scs.c:179-196
  print_line();
  B200_PRINTF(" iter | pri res | dua res |   gap   |   obj   |  scale  | time (s)\n");
  print_line();
}

static void print_summary(const SCS_WORK *w, int i, const Clock::time_point &t0) {  // scs.c:198-210
  const HostResid &r = w->r_o;
  B200_PRINTF("%*i|%*.2e %*.2e %*.2e %*.2e %*.2e %*.2e \n", 6, i, 9, r.res_pri, 9, r.res_dual, 9, r.gap, 9,
              0.5 * (r.pobj + r.dobj), 9, w->stgs.scale, 9, (ms_since(t0) + w->setup_time) / 1e3);
}

static void print_footer(const ScsInfo *info) {  // scs.c:237-274
  print_line();
  B200_PRINTF("status:  %s\n", info->status);
  B200_PRINTF("timings: total: %1.2es = setup: %1.2es + solve: %1.2es\n",
              (info->setup_time + info->solve_time) / 1e3, info->setup_time / 1e3, info->solve_time / 1e3);
  B200_PRINTF("\t lin-sys: %1.2es, cones: %1.2es, accel: %1.2es\n", info->lin_sys_time / 1e3, info->cone_time / 1e3,
              info->accel_time / 1e3);
  print_line();
  B200_PRINTF("objective = %.6f", 0.5 * (info->pobj + info->dobj));
  if (info->status_val == SCS_SOLVED_INACCURATE || info->status_val == SCS_UNBOUNDED_INACCURATE ||
      info->status_val == SCS_INFEASIBLE_INACCURATE)
    B200_PRINTF(" (inaccurate)");
  B200_PRINTF("\n");
  print_line();
}

// ------------------------------------------------------------------ setup pieces ------
static int set_diag_r(SCS_WORK *w) {
  k_set_diag_r<<<ew_grid(w->c, w->l), kThreads, 0, w->c.stream>>>(w->diag_r, w->n, w->m, w->cone.z, w->stgs.rho_x,
                                                                  w->stgs.scale);
  w->c.launches++;
  return 0;
}

// SCS(normalize_a_p), scs_matrix.c:407-470, on the device copies of A, A', P
static int normalize_a_p_dev(SCS_WORK *w) {
  Ctx &c = w->c;
  LinSys &ls = w->ls;
  const int n = w->n, m = w->m;
  cudaStream_t st = c.stream;
  double *Dt = nullptr, *Et = nullptr;
  if (dev_alloc(&Dt, (size_t)m) || dev_alloc(&Et, (size_t)n)) return -1;
  k_fill<<<ew_grid(c, m), kThreads, 0, st>>>(w->D, 1.0, m);
  k_fill<<<ew_grid(c, n), kThreads, 0, st>>>(w->E, 1.0, n);
  c.launches += 2;
  for (int pass = 0; pass < kRuizPasses + kL2Passes; ++pass) {
    const bool l2 = pass >= kRuizPasses;
    EpiStore eD; eD.y = Dt;
    EpiStoreComb eE; eE.y = Et;
    if (!l2) {
      ElemAbsMax e;
      row_kernel<ElemAbsMax, ElemAbsMax, EpiStore, false>
          <<<ls.chA.grid, kThreads, 0, st>>>(ls.A, e, ls.A, e, ls.chA.d, ls.chA.n, eD, c.red, c.S, nullptr);
      if (w->cone.enforce_boundaries(Dt, 0)) return -1;
      k_inv_sqrt_limit<<<ew_grid(c, m), kThreads, 0, st>>>(Dt, m, 0);
      if (ls.hasP)
        row_kernel<ElemAbsMax, ElemAbsMax, EpiStoreComb, true>
            <<<ls.chAt.grid, kThreads, 0, st>>>(ls.At, e, ls.P, e, ls.chAt.d, ls.chAt.n, eE, c.red, c.S, nullptr);
      else
        row_kernel<ElemAbsMax, ElemAbsMax, EpiStoreComb, false>
            <<<ls.chAt.grid, kThreads, 0, st>>>(ls.At, e, ls.P, e, ls.chAt.d, ls.chAt.n, eE, c.red, c.S, nullptr);
      if (dist_allreduce(c, Et, (size_t)c.n_sh, 1)) return -1;  // shared columns: inf-norms over all ranks' rows
      k_inv_sqrt_limit<<<ew_grid(c, n), kThreads, 0, st>>>(Et, n, 0);
      c.launches += 4;
    } else {
      ElemSumSq e;
      row_kernel<ElemSumSq, ElemSumSq, EpiStore, false>
          <<<ls.chA.grid, kThreads, 0, st>>>(ls.A, e, ls.A, e, ls.chA.d, ls.chA.n, eD, c.red, c.S, nullptr);
      k_sqrt<<<ew_grid(c, m), kThreads, 0, st>>>(Dt, m);
      if (w->cone.enforce_boundaries(Dt, 1)) return -1;
      k_inv_sqrt_limit<<<ew_grid(c, m), kThreads, 0, st>>>(Dt, m, 0);
      if (ls.hasP && !c.dist)
        row_kernel<ElemSumSq, ElemSumSq, EpiStoreComb, true>
            <<<ls.chAt.grid, kThreads, 0, st>>>(ls.At, e, ls.P, e, ls.chAt.d, ls.chAt.n, eE, c.red, c.S, nullptr);
      else
        row_kernel<ElemSumSq, ElemSumSq, EpiStoreComb, false>
            <<<ls.chAt.grid, kThreads, 0, st>>>(ls.At, e, ls.P, e, ls.chAt.d, ls.chAt.n, eE, c.red, c.S, nullptr);
      if (c.dist) {
        // shared columns: sums of squares over all ranks' rows; P's rows (replicated on the shared block, local
        // on the private one) are added afterwards so that they are counted once
        if (dist_allreduce(c, Et, (size_t)c.n_sh, 0)) return -1;
        if (ls.hasP) {
          EpiAccum ea; ea.y = Et;
          row_kernel<ElemSumSq, ElemSumSq, EpiAccum, false>
              <<<ls.chP.grid, kThreads, 0, st>>>(ls.P, e, ls.P, e, ls.chP.d, ls.chP.n, ea, c.red, c.S, nullptr);
          c.launches++;
        }
      }
      k_inv_sqrt_limit<<<ew_grid(c, n), kThreads, 0, st>>>(Et, n, 1);
      c.launches += 5;
    }
    RescaleA ra{Dt, Et};
    RescaleAt rat{Dt, Et};
    row_map_kernel<<<ls.chA.grid, kThreads, 0, st>>>(ls.A, ls.chA.d, ls.chA.n, ra);
    row_map_kernel<<<ls.chAt.grid, kThreads, 0, st>>>(ls.At, ls.chAt.d, ls.chAt.n, rat);
    if (ls.hasP) {
      RescaleP rp{Et};
      row_map_kernel<<<ls.chAt.grid, kThreads, 0, st>>>(ls.P, ls.chAt.d, ls.chAt.n, rp);
      c.launches++;
    }
    k_mul_inplace<<<ew_grid(c, m), kThreads, 0, st>>>(w->D, Dt, m);
    k_mul_inplace<<<ew_grid(c, n), kThreads, 0, st>>>(w->E, Et, n);
    c.launches += 4;
  }
  CUDA_OK(cudaGetLastError());
  if (c.sync()) return -1;
  dev_free(Dt);
  dev_free(Et);
  return 0;
}

// update_work_cache, scs.c:1066-1076: g = (R + M)^-1 [c ; -b] to CG_BEST_TOL
static int update_work_cache(SCS_WORK *w) {
  Ctx &c = w->c;
  k_build_g<<<ew_grid(c, w->n + w->m), kThreads, 0, c.stream>>>(w->g, w->cvec, w->b, w->n, w->m);
  c.launches++;
  if (w->ls.prepare_flags(w->g, kCgBestTol)) return -1;
  return w->ls.solve_dev(w->g, nullptr, 16);
}

// populate_residual_struct, scs.c:513-585 (+ unnormalize_residuals, scs.c:465-509)
static int populate_residuals(SCS_WORK *w, int iter) {
  if (w->r_n.last_iter == iter) return 0;
  Ctx &c = w->c;
  LinSys &ls = w->ls;
  const int n = w->n, m = w->m;
  {
    EpiResA epi;
    epi.u = w->u; epi.rsk = w->rsk; epi.b = w->b; epi.D = w->D; epi.n = n; epi.m_ = m;
    epi.inv_ds = 1.0 / w->dual_scale; epi.dual_scale = w->dual_scale;
    ElemMul e{w->u};
    row_kernel<ElemMul, ElemMul, EpiResA, false>
        <<<ls.chA.grid, kThreads, 0, c.stream>>>(ls.A, e, ls.A, e, ls.chA.d, ls.chA.n, epi, c.red, c.S, nullptr);
    if (dist_finish(c, 1, 7, FinResA{w->u + n + m, w->rsk + n + m})) return -1;
  }
  {
    EpiResAt epi;
    epi.u = w->u; epi.c = w->cvec; epi.E = w->E; epi.n = n; epi.m_ = m; epi.inv_ps = 1.0 / w->primal_scale;
    epi.extra = nullptr;
    ElemMul ea{w->u + n}, eb{w->u};
    if (c.dist) {  // A_g' y_g summed over the ranks, then the P pass with the same epilogue
      EpiStore es; es.y = ls.Gp;
      row_kernel<ElemMul, ElemMul, EpiStore, false>
          <<<ls.chAt.grid, kThreads, 0, c.stream>>>(ls.At, ea, ls.At, ea, ls.chAt.d, ls.chAt.n, es, c.red, c.S, nullptr);
      if (ls.dist_reduce_gp(ls.Gp, false, nullptr)) return -1;
      epi.extra = ls.Gp;
      epi.cnt_lo = c.cnt_lo;
      row_kernel<ElemMul, ElemMul, EpiResAt, false>
          <<<ls.chP.grid, kThreads, 0, c.stream>>>(ls.P, eb, ls.P, eb, ls.chP.d, ls.chP.n, epi, c.red, c.S, nullptr);
      if (dist_finish(c, 2, 6, FinResAt{})) return -1;
      c.launches++; c.spmv_calls++;
    } else if (ls.hasP)
      row_kernel<ElemMul, ElemMul, EpiResAt, true>
          <<<ls.chAt.grid, kThreads, 0, c.stream>>>(ls.At, ea, ls.P, eb, ls.chAt.d, ls.chAt.n, epi, c.red, c.S, nullptr);
    else
      row_kernel<ElemMul, ElemMul, EpiResAt, false>
          <<<ls.chAt.grid, kThreads, 0, c.stream>>>(ls.At, ea, ls.P, eb, ls.chAt.d, ls.chAt.n, epi, c.red, c.S, nullptr);
  }
  c.launches += 2; c.spmv_calls += 2;
  CUDA_OK(cudaGetLastError());
  if (c.fetch_scalars()) return -1;
  w->ls.tot_cg_its = c.S_host->cg_its_total;
  const double *R = c.S_host->res;
  HostResid &r = w->r_n;
  r.last_iter = iter;
  r.tau = R[R_TAU]; r.kap = R[R_KAP];
  r.xt_p_x_tau = ls.hasP ? R[R_XPX_TAU] : 0.0;
  r.bty_tau = R[R_BTY_TAU]; r.ctx_tau = R[R_CTX_TAU];
  r.bty = safediv_pos_h(r.bty_tau, r.tau);
  r.ctx = safediv_pos_h(r.ctx_tau, r.tau);
  r.xt_p_x = safediv_pos_h(r.xt_p_x_tau, r.tau * r.tau);
  r.gap = fabs(r.xt_p_x + r.ctx + r.bty);
  r.pobj = r.xt_p_x / 2. + r.ctx;
  r.dobj = -r.xt_p_x / 2. - r.bty;
  r.nm_ax_s_btau = R[R_NM_AX_S_BTAU]; r.nm_ax_s = R[R_NM_AX_S]; r.nm_ax = R[R_NM_AX];
  r.nm_px_aty_ctau = R[R_NM_PX_ATY_CTAU]; r.nm_px = R[R_NM_PX]; r.nm_aty = R[R_NM_ATY];
  r.nm_s = R[R_ONM_S];
  compute_residuals_h(r, 1.0);
  if (w->stgs.normalize) {
    HostResid &o = w->r_o;
    const double pd = w->primal_scale * w->dual_scale;
    o.last_iter = r.last_iter; o.tau = r.tau;
    o.kap = r.kap / pd; o.bty_tau = r.bty_tau / pd; o.ctx_tau = r.ctx_tau / pd; o.xt_p_x_tau = r.xt_p_x_tau / pd;
    o.xt_p_x = r.xt_p_x / pd; o.ctx = r.ctx / pd; o.bty = r.bty / pd; o.pobj = r.pobj / pd; o.dobj = r.dobj / pd;
    o.gap = r.gap / pd;
    o.nm_ax_s_btau = R[R_ONM_AX_S_BTAU]; o.nm_ax_s = R[R_ONM_AX_S]; o.nm_ax = R[R_ONM_AX];
    o.nm_px_aty_ctau = R[R_ONM_PX_ATY_CTAU]; o.nm_px = R[R_ONM_PX]; o.nm_aty = R[R_ONM_ATY];
    o.nm_s = R[R_ONM_S];
    compute_residuals_h(o, pd);
  } else {
    w->r_o = w->r_n;
  }
  return 0;
}

// SCS(log_data_to_csv), rw.c:317-476: one row per call, file opened "w" at iteration 0 and "a" afterwards.
// Costs two SpMV passes, one reduction and a host sync per row, like the reference's per-iteration
// populate_residual_struct (scs.c:1396-1401).
static int csv_log(SCS_WORK *w, int iter, const Clock::time_point &t0) {
  Ctx &c = w->c;
  LinSys &ls = w->ls;
  const int n = w->n, m = w->m;
  if (populate_residuals(w, iter)) return -1;
  if (!w->csv_m && (dev_alloc(&w->csv_m, (size_t)m) || dev_alloc(&w->csv_n, (size_t)n))) return -1;
  {
    EpiStore es; es.y = w->csv_m;
    ElemMul e{w->u};
    row_kernel<ElemMul, ElemMul, EpiStore, false>
        <<<ls.chA.grid, kThreads, 0, c.stream>>>(ls.A, e, ls.A, e, ls.chA.d, ls.chA.n, es, c.red, c.S, nullptr);
    EpiStore et; et.y = w->csv_n;
    ElemMul ea{w->u + n}, eb{w->u};
    if (ls.hasP)
      row_kernel<ElemMul, ElemMul, EpiStore, true>
          <<<ls.chAt.grid, kThreads, 0, c.stream>>>(ls.At, ea, ls.P, eb, ls.chAt.d, ls.chAt.n, et, c.red, c.S, nullptr);
    else
      row_kernel<ElemMul, ElemMul, EpiStore, false>
          <<<ls.chAt.grid, kThreads, 0, c.stream>>>(ls.At, ea, ls.P, eb, ls.chAt.d, ls.chAt.n, et, c.red, c.S, nullptr);
    k_csv_norms<<<ew_grid(c, w->l), kThreads, 0, c.stream>>>(w->u, w->u_t, w->v, w->v_prev, w->rsk, w->csv_m, w->csv_n,
                                                             w->b, w->cvec, w->D, w->E, n, m, w->primal_scale,
                                                             w->dual_scale, c.red, c.S);
    c.launches += 3; c.spmv_calls += 2;
  }
  if (c.fetch_scalars()) return -1;
  const double *q = c.S_host->part;
  FILE *f = fopen(w->csv_fn.c_str(), iter == 0 ? "w" : "a");
  if (!f) {
    B200_PRINTF("Error: Could not open %s for writing\n", w->csv_fn.c_str());
    return 0;
  }
  if (iter == 0) fprintf(f, "%s\n", scs_b200_csv_header());
  const HostResid &r = w->r_o, &rn = w->r_n;
  const bool nrm = w->stgs.normalize != 0;
  auto F = [f](double v) { fprintf(f, "%.16e,", v); };
  fprintf(f, "%li,", (long)iter);
  F(r.res_pri); F(r.res_dual); F(r.gap);
  // un-normalised then normalised solution norms (identical without equilibration)
  F(nrm ? q[15] : q[12]); F(nrm ? q[16] : q[13]); F(nrm ? q[17] : q[14]);
  F(sqrt(nrm ? q[3] : q[0])); F(sqrt(nrm ? q[4] : q[1])); F(sqrt(nrm ? q[5] : q[2]));
  F(q[12]); F(q[13]); F(q[14]); F(sqrt(q[0])); F(sqrt(q[1])); F(sqrt(q[2]));
  F(r.nm_ax_s_btau); F(r.nm_px_aty_ctau); F(sqrt(nrm ? q[8] : q[6])); F(sqrt(nrm ? q[9] : q[7]));
  F(r.res_infeas); F(r.res_unbdd_a); F(r.res_unbdd_p); F(r.pobj); F(r.dobj); F(r.tau); F(r.kap);
  F(rn.res_pri); F(rn.res_dual); F(rn.gap);
  F(rn.nm_ax_s_btau); F(rn.nm_px_aty_ctau); F(sqrt(q[6])); F(sqrt(q[7]));
  F(rn.res_infeas); F(rn.res_unbdd_a); F(rn.res_unbdd_p); F(rn.pobj); F(rn.dobj); F(rn.tau); F(rn.kap);
  F(r.nm_ax); F(r.nm_ax_s); F(r.nm_px); F(r.nm_aty);
  F(r.xt_p_x); F(r.xt_p_x_tau); F(r.ctx); F(r.ctx_tau); F(r.bty); F(r.bty_tau);
  F(w->nm_b_orig); F(w->nm_c_orig); F(w->stgs.scale);
  F(sqrt(q[10])); F(sqrt(q[11])); F(q[18]); F(q[19]);
  F(c.S_host->aa_norm);
  fprintf(f, "%li,", (long)c.S_host->aa_accepted);
  fprintf(f, "%li,", (long)c.S_host->aa_rejected);
  F(ms_since(t0) / 1e3);
  fprintf(f, "\n");
  fclose(f);
  return 0;
}

// has_converged, scs.c:589-627 (isless() == NaN-safe '<')
static int has_converged(const SCS_WORK *w) {
  const HostResid &r = w->r_o;
  const double eps_abs = w->stgs.eps_abs, eps_rel = w->stgs.eps_rel, eps_infeas = w->stgs.eps_infeas;
  if (r.tau > 0.) {
    const double grl = fmax(fmax(fabs(r.xt_p_x), fabs(r.ctx)), fabs(r.bty));
    const double prl = fmax(fmax(w->nm_b_orig * r.tau, r.nm_s), r.nm_ax) / r.tau;
    const double drl = fmax(fmax(w->nm_c_orig * r.tau, r.nm_px), r.nm_aty) / r.tau;
    if (std::isless(r.res_pri, eps_abs + eps_rel * prl) && std::isless(r.res_dual, eps_abs + eps_rel * drl) &&
        std::isless(r.gap, eps_abs + eps_rel * grl))
      return SCS_SOLVED;
  }
  if (std::isless(r.res_unbdd_a, eps_infeas) && std::isless(r.res_unbdd_p, eps_infeas)) return SCS_UNBOUNDED;
  if (std::isless(r.res_infeas, eps_infeas)) return SCS_INFEASIBLE;
  return 0;
}

// update_scale, scs.c:1112-1189
static int update_scale(SCS_WORK *w, int iter) {
  const HostResid &r = w->r_o;
  const int since = iter - w->last_scale_update_iter;
  double denom_pri = fmax(fmax(r.nm_ax, r.nm_s), w->nm_b_orig * r.tau);
  double rel_pri = safediv_pos_h(r.nm_ax_s_btau, denom_pri);
  double denom_dual = fmax(fmax(r.nm_px, r.nm_aty), w->nm_c_orig * r.tau);
  double rel_dual = safediv_pos_h(r.nm_px_aty_ctau, denom_dual);
  rel_pri = fmax(rel_pri, kDivEps);
  rel_dual = fmax(rel_dual, kDivEps);
  w->sum_log_scale_factor += log(rel_pri) - log(rel_dual);
  w->n_log_scale_factor++;
  const double factor = sqrt(exp(w->sum_log_scale_factor / (double)w->n_log_scale_factor));
  if (since < kRescalingMinIters) return 0;
  const double new_scale = fmin(fmax(w->stgs.scale * factor, kMinScale), kMaxScale);
  if (new_scale == w->stgs.scale) return 0;
  if (factor > sqrt(10.) || factor < 1. / sqrt(10.)) {
    w->scale_updates++;
    w->sum_log_scale_factor = 0;
    w->n_log_scale_factor = 0;
    w->last_scale_update_iter = iter;
    w->unroll_cg0 = -1;  // the g solve below must not count as CG iterations of an ADMM iteration
    w->stgs.scale = new_scale;
    if (set_diag_r(w)) return -1;
    if (w->ls.update_precond()) return -1;
    if (update_work_cache(w)) return -1;
    if (w->has_aa) w->aa.reset();
    k_remap_v<<<ew_grid(w->c, w->l), kThreads, 0, w->c.stream>>>(w->v, w->rsk, w->u, w->u_t, w->diag_r, w->l);
    w->c.launches++;
  }
  return 0;
}

// SURVEY.md 8(d) byte model, from the counters of the loop: ADMM iterations, CG iterations
// (device counter), residual checks and AA applications since the workspace was created.
static double model_bytes(const SCS_WORK *w, long long iters, long long cg_its, long long checks, long long aa_n) {
  const LinSys &ls = w->ls;
  const double n = w->n, m = w->m, l = w->l;
  const double per_iter = 2.0 * (ls.bytes_A() + ls.bytes_At()) + ls.bytes_P() + 6.0 * (n + m) * 8.0 + 14.0 * l * 8.0;
  const double per_cg = ls.bytes_A() + ls.bytes_At() + ls.bytes_P() + 10.0 * n * 8.0;
  const double per_check = ls.bytes_A() + ls.bytes_At() + ls.bytes_P() + 20.0 * l * 8.0;
  const double per_aa = w->has_aa ? (3.0 * w->aa.mem + 12.0) * l * 8.0 : 0.0;
  return iters * per_iter + cg_its * per_cg + checks * per_check + aa_n * per_aa;
}

// total kernels of this library launched so far: host-counted launches + the CG-loop kernels
// (4 per CG iteration that really executed; decided on the device)
static long long total_launches(const SCS_WORK *w, long long cg_its) {
  return w->c.launches + w->ls.cg_iter_launches() * cg_its;
}

static int set_kt(SCS_WORK *w, int on) {
  Ctx &c = w->c;
  static const unsigned long long zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  CUDA_OK(cudaMemcpyAsync(&c.S->kt_ticket[0], zeros, sizeof(unsigned int) * 4 + sizeof(unsigned long long) * 4,
                          cudaMemcpyHostToDevice, c.stream));
  CUDA_OK(cudaMemcpyAsync(&c.S->kt_on, &on, sizeof(int), cudaMemcpyHostToDevice, c.stream));
  return 0;
}

static void mark_open(SCS_WORK *w, int iter) {
  Ctx &c = w->c;
  c.fetch_scalars();  // counters at the start of the region (synchronises)
  w->mk_cg0 = c.S_host->cg_its_total;
  w->mk_launch0 = total_launches(w, w->mk_cg0);
  w->mk_bytes0 = model_bytes(w, w->admm_iters + iter, w->mk_cg0, w->n_checks, w->n_aa);
  w->mk_iter0 = iter;
  w->mk_open = true;
  set_kt(w, 1);
  cudaEventRecord(w->mark_ev[0], c.stream);
}
static void mark_close(SCS_WORK *w, int iter) {
  Ctx &c = w->c;
  cudaEventRecord(w->mark_ev[1], c.stream);
  c.fetch_scalars();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, w->mark_ev[0], w->mark_ev[1]);
  const DevScalars *h = c.S_host;
  ScsB200Marks &mk = w->marks;
  mk.ms = ms;
  mk.iters = iter - w->mk_iter0;
  mk.cg_iters = h->cg_its_total - w->mk_cg0;
  mk.kernel_launches = total_launches(w, h->cg_its_total) - w->mk_launch0;
  mk.algorithmic_bytes = model_bytes(w, w->admm_iters + iter, h->cg_its_total, w->n_checks, w->n_aa) - w->mk_bytes0;
  mk.spmv_a_ms = h->kt_ns[0] * 1e-6; mk.spmv_g_ms = h->kt_ns[1] * 1e-6;
  mk.spmv_a_launches = h->kt_cnt[0]; mk.spmv_g_launches = h->kt_cnt[1];
  mk.bytes_a = w->ls.bytes_A();
  mk.bytes_g = w->ls.bytes_At() + w->ls.bytes_P();
  set_kt(w, 0);
  w->mk_open = false;
  w->mk_done = true;
}

// enqueue the front half of one ADMM iteration: k_prep .. cone projections (scs.c:1315-1340).
// In graph mode `loop` carries the WHILE handle and the CG iterations are NOT enqueued here.
static int enqueue_front_head(SCS_WORK *w, CgCtl loop) {
  Ctx &c = w->c;
  const int n = w->n, m = w->m, gl = ew_grid(c, w->l);
  const int l_total = w->dist ? w->n_total + w->m_total + 1 : w->l;
  cudaStream_t st = c.stream;
  if (w->has_aa)
    k_prep<true><<<gl, kThreads, 0, st>>>(w->v, w->v_prev, w->u_t, w->u, w->g, w->diag_r, w->ws, n, m, l_total, c.red, c.S);
  else
    k_prep<false><<<gl, kThreads, 0, st>>>(w->v, w->v_prev, w->u_t, w->u, w->g, w->diag_r, w->ws, n, m, l_total, c.red, c.S);
  c.launches++;
  if (dist_finish(c, 0, 2, FinPrep{})) return -1;
  return w->ls.enqueue_head(w->u_t, w->ws, loop);
}
static int enqueue_front_tail(SCS_WORK *w) {
  Ctx &c = w->c;
  const int n = w->n, m = w->m, gl = ew_grid(c, w->l);
  cudaStream_t st = c.stream;
  if (w->ls.enqueue_tail(w->u_t)) return -1;
  k_rootplus<<<ew_grid(c, n + m), kThreads, 0, st>>>(w->u_t, w->v, w->g, w->diag_r, n, n + m, w->cnt_lo, c.red, c.S);
  if (dist_finish(c, 5, 0, FinRootPlus{w->v + n + m, w->diag_r + n + m})) return -1;
  k_pre<<<gl, kThreads, 0, st>>>(w->u_t, w->u, w->rsk, w->v, w->g, w->diag_r, n, m, w->cone.z, w->cone.z + w->cone.l, c.S);
  c.launches += 2;
  if (w->cone.has_nonlinear()) {
    if (w->cone.project_nonlinear(w->u + n, w->rsk + n, w->diag_r + n)) return -1;
  }
  return 0;
}

// Capture the front half into a CUDA graph whose CG loop is a WHILE conditional node: one
// graph launch per ADMM iteration, no host synchronisation until the next residual check.
static int build_iter_graph(SCS_WORK *w) {
  Ctx &c = w->c;
  const char *env = getenv("SCS_B200_NO_GRAPH");
  if (env && env[0] == '1') return 0;
  if (w->dist) return 0;  // NCCL all-reduces sit inside the CG loop: stream launches, host-read stop flag
  if (const char *eu = getenv("SCS_B200_CG_UNROLL")) {  // unrolled CG iterations (tests, comparisons)
    if (!strcmp(eu, "auto")) { if (!w->cg_unroll_auto) { w->cg_unroll_auto = true; w->cg_unroll = 3; } }
    else { w->cg_unroll = std::max(0, std::min(32, atoi(eu))); w->cg_unroll_auto = false; }
  }
  cudaStream_t st = c.stream;
  const long long l0 = c.launches, s0 = c.spmv_calls;
  bool capturing = false, ok = false;
  do {
    if (cudaStreamCreateWithFlags(&w->st_body, cudaStreamNonBlocking) != cudaSuccess) break;
    if (cudaGraphCreate(&w->graph, 0) != cudaSuccess) break;
    CgCtl loop;
    if (cudaGraphConditionalHandleCreate(&loop.h, w->graph, 0, cudaGraphCondAssignDefault) != cudaSuccess) break;
    loop.use_h = 1;
    if (cudaStreamBeginCaptureToGraph(st, w->graph, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) break;
    capturing = true;
    if (enqueue_front_head(w, loop)) break;
    // Optionally the first cg_unroll CG iterations are plain kernel nodes (kernels launched after convergence return
    // at once) and only what is left goes through the WHILE node: an experiment that showed the conditional node is
    // not what makes small problems slow (see SCS_WORK::cg_unroll).
    {
      bool bad = false;
      for (int k = 0; k < w->cg_unroll && !bad; ++k) bad = w->ls.enqueue_cg_iter(w->u_t, loop, -1) != 0;
      if (bad) break;
    }
    // WHILE node after everything captured so far
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    if (cudaStreamGetCaptureInfo_v2(st, &status, nullptr, nullptr, &deps, &ndeps) != cudaSuccess) break;
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = loop.h;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t wnode;
    if (cudaGraphAddNode(&wnode, w->graph, deps, ndeps, &np) != cudaSuccess) break;
    cudaGraph_t body = np.conditional.phGraph_out[0];
    if (cudaStreamUpdateCaptureDependencies(st, &wnode, 1, cudaStreamSetCaptureDependencies) != cudaSuccess) break;
    {  // loop body: one CG iteration, captured on a second stream into the body graph
      if (cudaStreamBeginCaptureToGraph(w->st_body, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) break;
      c.stream = w->st_body;
      const int rc = w->ls.enqueue_cg_iter(w->u_t, loop, -1);
      c.stream = st;
      cudaGraph_t tmp = nullptr;
      if (cudaStreamEndCapture(w->st_body, &tmp) != cudaSuccess || rc) break;
    }
    if (enqueue_front_tail(w)) break;
    cudaGraph_t tmp = nullptr;
    capturing = false;
    if (cudaStreamEndCapture(st, &tmp) != cudaSuccess) break;
    if (cudaGraphInstantiate(&w->gexec, w->graph, 0) != cudaSuccess) break;
    ok = true;
  } while (0);
  if (capturing) {
    cudaGraph_t tmp = nullptr;
    cudaStreamEndCapture(st, &tmp);
  }
  w->graph_launches = c.launches - l0;
  w->graph_spmv = c.spmv_calls - s0;
  c.launches = l0;
  c.spmv_calls = s0;
  if (!ok) {
    cudaGetLastError();
    B200_PRINTF("WARN: CUDA graph capture of the ADMM iteration failed; using stream launches.\n");
    if (w->gexec) { cudaGraphExecDestroy(w->gexec); w->gexec = nullptr; }
    if (w->graph) { cudaGraphDestroy(w->graph); w->graph = nullptr; }
    return 0;
  }
  w->use_graph = true;
  return 0;
}

// re-capture with the current w->cg_unroll (the stream is idle: called right after a residual check)
static int rebuild_iter_graph(SCS_WORK *w) {
  if (w->gexec) { cudaGraphExecDestroy(w->gexec); w->gexec = nullptr; }
  if (w->graph) { cudaGraphDestroy(w->graph); w->graph = nullptr; }
  if (w->st_body) { cudaStreamDestroy(w->st_body); w->st_body = nullptr; }
  w->use_graph = false;
  return build_iter_graph(w);
}

static void free_work(SCS_WORK *w) {
  if (!w) return;
  cudaSetDevice(w->c.device);
  if (w->c.stream) cudaStreamSynchronize(w->c.stream);
  if (w->mark_ev[0]) cudaEventDestroy(w->mark_ev[0]);
  if (w->mark_ev[1]) cudaEventDestroy(w->mark_ev[1]);
  w->ls.destroy();
  w->cone.destroy();
  if (w->has_aa) w->aa.destroy();
  if (w->gexec) cudaGraphExecDestroy(w->gexec);
  if (w->graph) cudaGraphDestroy(w->graph);
  if (w->st_body) cudaStreamDestroy(w->st_body);
  dev_free(w->u); dev_free(w->u_t); dev_free(w->v); dev_free(w->v_prev); dev_free(w->rsk); dev_free(w->g);
  dev_free(w->diag_r); dev_free(w->b); dev_free(w->cvec); dev_free(w->D); dev_free(w->E); dev_free(w->ws);
  dev_free(w->sol_x); dev_free(w->sol_y); dev_free(w->sol_s); dev_free(w->gather); dev_free(w->d_loc2glob);
  dev_free(w->csv_m); dev_free(w->csv_n);
  w->c.destroy();
  delete w;
}

// populate_on_failure / failure, scs.c:316-359
void populate_on_failure(int m, int n, ScsSolution *sol, ScsInfo *info, int status_val, const char *msg) {
  if (info) {
    info->gap = NAN; info->res_pri = NAN; info->res_dual = NAN; info->pobj = NAN; info->dobj = NAN;
    info->iter = -1; info->status_val = status_val; info->solve_time = NAN;
    strcpy(info->status, msg);
    memset(&info->aa_stats, 0, sizeof(info->aa_stats));
    info->aa_stats.last_aa_norm = NAN;
  }
  if (sol) {
    if (n > 0) {
      if (!sol->x) sol->x = (double *)calloc(n, sizeof(double));
      for (int i = 0; i < n; ++i) sol->x[i] = NAN;
    }
    if (m > 0) {
      if (!sol->y) sol->y = (double *)calloc(m, sizeof(double));
      if (!sol->s) sol->s = (double *)calloc(m, sizeof(double));
      for (int i = 0; i < m; ++i) { sol->y[i] = NAN; sol->s[i] = NAN; }
    }
  }
}
static int failure(SCS_WORK *w, int m, int n, ScsSolution *sol, ScsInfo *info, int stint, const char *msg,
                   const char *ststr) {
  (void)w;
  populate_on_failure(m, n, sol, info, stint, ststr);
  B200_PRINTF("Failure:%s\n", msg);
  end_interrupt_listener();
  return stint;
}

static void fill_aa_stats(SCS_WORK *w, ScsInfo *info) {
  memset(&info->aa_stats, 0, sizeof(info->aa_stats));
  info->aa_stats.last_aa_norm = NAN;
  if (w->has_aa && w->aa.fetch_state() == 0) {
    const AaState *h = w->aa.st_host;
    AaStats &s = info->aa_stats;
    s.iter = h->iter; s.n_accept = h->n_accept; s.n_reject_lapack = h->n_reject_lapack;
    s.n_reject_rank0 = h->n_reject_rank0; s.n_reject_nonfinite = h->n_reject_nonfinite;
    s.n_reject_weight_cap = h->n_reject_weight_cap; s.n_safeguard_reject = h->n_safeguard_reject;
    s.last_rank = h->last_rank; s.last_aa_norm = h->last_aa_norm; s.last_regularization = h->last_regularization;
  }
}


// ------------------------------------------------------------- row partition (dist) ---
// Rank g of a row-partitioned solve owns a contiguous block of rows of A.  Cuts may fall on
// any row of the zero / nonneg cones and otherwise only between cones, and are
// placed so that every rank holds about the same modelled work (stored entries + a per-row
// charge for the m-space vector passes).
//
// Columns: a column of A whose non-zeros all lie in one rank's rows, and whose row of P holds
// nothing but the diagonal, is PRIVATE to that rank -- no other rank ever multiplies by it, so its
// entries of x, c, E, the CG vectors ... exist on that rank only.  Every other column is SHARED
// (replicated).  The local column order is [shared, ascending | private, ascending]; the shared
// block is the only part of an n-vector that is ever all-reduced.
struct LocalProblem {
  std::vector<double> Ax, b, c, Px;
  std::vector<int> Ai, Ap, Pi, Pp;
  std::vector<int> loc2glob;  // n_loc global column indices
  ScsMatrix A, P;
  ScsData d;
  ScsCone k;
  int row0 = 0, m = 0, n_sh = 0, n_loc = 0;
  long long nnz_max_rank = 0;
};

static int dist_row_weight() {
  const char *e = getenv("SCS_B200_DIST_ROW_WEIGHT");
  return e && *e ? atoi(e) : 4;
}

// cut[g] .. cut[g+1]: rows of rank g
static int partition_cuts(const ScsData *d, const ScsCone *k, int world, std::vector<int> &cut) {
  const int m = d->m, n = d->n;
  const ScsMatrix *A = d->A;
  const long long rw = dist_row_weight();
  std::vector<long long> pref((size_t)m + 1, 0);  // work in rows [0, i)
  for (int t = 0; t < A->p[n]; ++t) pref[(size_t)A->i[t] + 1]++;
  for (int i = 0; i < m; ++i) pref[(size_t)i + 1] += pref[(size_t)i] + rw;
  // allowed cut positions beyond the z / l rows: cone boundaries
  std::vector<int> bnd;
  int off = k->z + k->l;
  auto push = [&](int sz) { off += sz; bnd.push_back(off); };
  if (k->bsize > 0) push(k->bsize);
  for (int i = 0; i < k->qsize; ++i) push(k->q[i]);
  for (int i = 0; i < k->ssize; ++i) push(k->s[i] * (k->s[i] + 1) / 2);
  for (int i = 0; i < k->cssize; ++i) push(k->cs[i] * k->cs[i]);
  for (int i = 0; i < k->ep + k->ed + k->psize; ++i) push(3);
  const int zl = k->z + k->l;
  auto next_allowed = [&](int pos) {  // smallest allowed cut >= pos
    if (pos <= zl) return pos;
    auto it = std::lower_bound(bnd.begin(), bnd.end(), pos);
    return it == bnd.end() ? m : *it;
  };
  cut.assign((size_t)world + 1, 0);
  cut[(size_t)world] = m;
  const long long total = pref[(size_t)m];
  for (int g = 1; g < world; ++g) {
    const long long target = total * g / world;
    int pos = (int)(std::lower_bound(pref.begin(), pref.end(), target) - pref.begin());
    pos = next_allowed(pos);
    if (pos <= cut[(size_t)g - 1]) pos = next_allowed(cut[(size_t)g - 1] + 1);
    cut[(size_t)g] = pos > m ? m : pos;
  }
  for (int g = 0; g < world; ++g)
    if (cut[(size_t)g + 1] <= cut[(size_t)g]) {
      B200_PRINTF("ERROR: cannot cut %d rows into %d non-empty cone-aligned blocks\n", m, world);
      return -1;
    }
  return 0;
}

// owner[j] = rank the column is private to, or -1 when shared
static void classify_columns(const ScsData *d, const std::vector<int> &cut, int world, std::vector<int> &owner) {
  const int n = d->n;
  const ScsMatrix *A = d->A, *P = d->P;
  owner.assign((size_t)n, 0);
  auto rank_of = [&](int row) { return (int)(std::upper_bound(cut.begin() + 1, cut.end(), row) - (cut.begin() + 1)); };
  for (int j = 0; j < n; ++j) {
    const int s = A->p[j], e = A->p[j + 1];
    if (e == s) { owner[(size_t)j] = j % world; continue; }  // touched by no row: any rank may own it
    const int r0 = rank_of(A->i[s]), r1 = rank_of(A->i[e - 1]);  // row indices are sorted inside a column
    owner[(size_t)j] = (r0 == r1) ? r0 : -1;
  }
  if (P)
    for (int j = 0; j < n; ++j)
      for (int t = P->p[j]; t < P->p[j + 1]; ++t)
        if (P->i[t] != j) { owner[(size_t)j] = -1; owner[(size_t)P->i[t]] = -1; }  // coupled columns stay replicated
  if (const char *e = getenv("SCS_B200_DIST_FORCE_SHARED")) {  // tests: every k-th column is treated as shared
    const int kk = atoi(e);
    if (kk > 0)
      for (int j = 0; j < n; j += kk) owner[(size_t)j] = -1;
  }
}

static int partition_rows(const ScsData *d, const ScsCone *k, int rank, int world, LocalProblem &out) {
  const int n = d->n;
  const ScsMatrix *A = d->A;
  std::vector<int> cut, owner;
  if (partition_cuts(d, k, world, cut)) return -1;
  classify_columns(d, cut, world, owner);
  const int r0 = cut[(size_t)rank], r1 = cut[(size_t)rank + 1];
  out.row0 = r0;
  out.m = r1 - r0;
  // local columns: shared first, then this rank's private ones
  out.loc2glob.clear();
  for (int j = 0; j < n; ++j) if (owner[(size_t)j] < 0) out.loc2glob.push_back(j);
  out.n_sh = (int)out.loc2glob.size();
  for (int j = 0; j < n; ++j) if (owner[(size_t)j] == rank) out.loc2glob.push_back(j);
  const int nl = out.n_loc = (int)out.loc2glob.size();
  if (nl == 0) {
    B200_PRINTF("ERROR: rank %d of %d would own no column\n", rank, world);
    return -1;
  }
  // local CSC: rows [r0, r1) of every local column (row indices are sorted inside a column)
  out.Ap.assign((size_t)nl + 1, 0);
  for (int jl = 0; jl < nl; ++jl) {
    const int j = out.loc2glob[(size_t)jl];
    const int *b0 = A->i + A->p[j], *b1 = A->i + A->p[j + 1];
    const int *lo = std::lower_bound(b0, b1, r0), *hi = std::lower_bound(b0, b1, r1);
    out.Ap[(size_t)jl + 1] = out.Ap[(size_t)jl] + (int)(hi - lo);
  }
  out.Ai.resize((size_t)out.Ap[(size_t)nl]);
  out.Ax.resize((size_t)out.Ap[(size_t)nl]);
  for (int jl = 0; jl < nl; ++jl) {
    const int j = out.loc2glob[(size_t)jl];
    const int *b0 = A->i + A->p[j], *b1 = A->i + A->p[j + 1];
    const int lo = (int)(std::lower_bound(b0, b1, r0) - A->i), cnt = out.Ap[(size_t)jl + 1] - out.Ap[(size_t)jl];
    for (int t = 0; t < cnt; ++t) {
      out.Ai[(size_t)out.Ap[(size_t)jl] + t] = A->i[lo + t] - r0;
      out.Ax[(size_t)out.Ap[(size_t)jl] + t] = A->x[lo + t];
    }
  }
  out.b.assign(d->b + r0, d->b + r1);
  out.c.resize((size_t)nl);
  for (int jl = 0; jl < nl; ++jl) out.c[(size_t)jl] = d->c[out.loc2glob[(size_t)jl]];
  out.A.x = out.Ax.data(); out.A.i = out.Ai.data(); out.A.p = out.Ap.data(); out.A.m = out.m; out.A.n = nl;
  out.d = *d;
  out.d.m = out.m; out.d.n = nl; out.d.A = &out.A; out.d.b = out.b.data(); out.d.c = out.c.data();
  out.d.P = nullptr;
  if (d->P) {  // P restricted to the local columns: a coupled (off-diagonal) pair is shared on both ends, and
               // the shared block keeps the global order, so the result is still upper triangular and sorted
    const ScsMatrix *P = d->P;
    std::vector<int> g2l((size_t)n, -1);
    for (int jl = 0; jl < nl; ++jl) g2l[(size_t)out.loc2glob[(size_t)jl]] = jl;
    out.Pp.assign((size_t)nl + 1, 0);
    for (int jl = 0; jl < nl; ++jl) {
      const int j = out.loc2glob[(size_t)jl];
      for (int t = P->p[j]; t < P->p[j + 1]; ++t) {
        const int il = g2l[(size_t)P->i[t]];
        if (il < 0) { B200_PRINTF("ERROR: internal: P couples a local column to a foreign private column\n"); return -1; }
        out.Pi.push_back(il);
        out.Px.push_back(P->x[t]);
      }
      out.Pp[(size_t)jl + 1] = (int)out.Pi.size();
    }
    if (out.Pi.empty()) { out.Pi.push_back(0); out.Px.push_back(0.0); }  // keep the pointers non-null
    out.P.x = out.Px.data(); out.P.i = out.Pi.data(); out.P.p = out.Pp.data(); out.P.m = nl; out.P.n = nl;
    out.d.P = &out.P;
  }
  // the cones inside [r0, r1)
  ScsCone &lk = out.k;
  lk = *k;
  const int zl = k->z + k->l;
  auto overlap = [&](int a0, int a1) { const int lo = a0 > r0 ? a0 : r0, hi = a1 < r1 ? a1 : r1; return hi > lo ? hi - lo : 0; };
  lk.z = overlap(0, k->z);
  lk.l = overlap(k->z, zl);
  int off = zl;
  auto inside = [&](int start) { return start >= r0 && start < r1; };
  if (k->bsize > 0) {
    if (!inside(off)) { lk.bsize = 0; lk.bu = nullptr; lk.bl = nullptr; }
    off += k->bsize;
  }
  auto take = [&](int count, auto size_of, int &first, int &num) {
    first = -1; num = 0;
    for (int i = 0; i < count; ++i) {
      if (inside(off) && size_of(i) > 0) { if (first < 0) first = i; num++; }
      else if (size_of(i) == 0 && first >= 0 && off < r1) num++;  // empty cones ride along
      off += size_of(i);
    }
    if (first < 0) first = 0;
  };
  int f = 0, c = 0;
  take(k->qsize, [&](int i) { return k->q[i]; }, f, c);
  lk.q = k->q ? k->q + f : nullptr; lk.qsize = c;
  take(k->ssize, [&](int i) { return k->s[i] * (k->s[i] + 1) / 2; }, f, c);
  lk.s = k->s ? k->s + f : nullptr; lk.ssize = c;
  take(k->cssize, [&](int i) { return k->cs[i] * k->cs[i]; }, f, c);
  lk.cs = k->cs ? k->cs + f : nullptr; lk.cssize = c;
  take(k->ep, [&](int) { return 3; }, f, c);
  lk.ep = c;
  take(k->ed, [&](int) { return 3; }, f, c);
  lk.ed = c;
  take(k->psize, [&](int) { return 3; }, f, c);
  lk.p = k->p ? k->p + f : nullptr; lk.psize = c;
  return 0;
}

}  // namespace b200

// ==================================================================== public API ======
extern "C" const char *scs_version(void) { return B200_SCS_VERSION; }

extern "C" void scs_set_default_settings(ScsSettings *stgs) {  // util.c:158-179
  if (!stgs) return;
  stgs->max_iters = 100000;
  stgs->eps_abs = 1e-4;
  stgs->eps_rel = 1e-4;
  stgs->eps_infeas = 1e-7;
  stgs->alpha = 1.5;
  stgs->rho_x = 1e-6;
  stgs->scale = 0.1;
  stgs->verbose = 1;
  stgs->normalize = 1;
  stgs->warm_start = 0;
  stgs->acceleration_lookback = 10;
  stgs->acceleration_interval = 10;
  stgs->acceleration_type_1 = 1;
  stgs->acceleration_regularization = 1e-8;
  stgs->acceleration_relaxation = 1.0;
  stgs->adaptive_scale = 1;
  stgs->write_data_filename = nullptr;
  stgs->log_csv_filename = nullptr;
  stgs->time_limit_secs = 0.;
}

extern "C" scs_int scs_update(ScsWork *w, scs_float *b, scs_float *c) {  // scs.c:1235-1273
  if (!w) return -1;
  const Clock::time_point t0 = Clock::now();
  Ctx &cx = w->c;
  if (cudaSetDevice(cx.device) != cudaSuccess) return -1;
  const int n = w->n, m = w->m;
  const int mt = w->dist ? w->m_total : m;  // b is always the caller's full-length vector
  if (b) {
    if (w->b_orig.data() != b) memcpy(w->b_orig.data(), b, sizeof(double) * mt);
    double nm = 0.0;
    for (int i = 0; i < mt; ++i) nm = fmax(nm, fabs(b[i]));
    w->nm_b_orig = nm;
  }
  const int nt = w->dist ? w->n_total : n;  // c is always the caller's full-length vector
  if (c) {
    if (w->c_orig.data() != c) memcpy(w->c_orig.data(), c, sizeof(double) * nt);
    double nm = 0.0;
    for (int i = 0; i < nt; ++i) nm = fmax(nm, fabs(c[i]));
    w->nm_c_orig = nm;
  }
  if (w->dist) {  // this rank's columns of c
    w->c_loc.resize((size_t)n);
    for (int j = 0; j < n; ++j) w->c_loc[(size_t)j] = w->c_orig[(size_t)w->loc2glob[(size_t)j]];
  }
  if (h2d(cx, w->b, w->b_orig.data() + w->row0, (size_t)m) ||
      h2d(cx, w->cvec, w->dist ? w->c_loc.data() : w->c_orig.data(), (size_t)n))
    return -1;
  if (w->stgs.normalize) {  // SCS(normalize_b_c), normalize.c:33-61
    k_scale_bc<<<ew_grid(cx, n + m), kThreads, 0, cx.stream>>>(w->b, w->D, m, w->cvec, w->E, n, cx.red, cx.S);
    if (dist_finish(cx, 0, 2, FinSigma{})) return -1;
    k_scale_sigma<<<ew_grid(cx, n + m), kThreads, 0, cx.stream>>>(w->b, m, w->cvec, n, cx.S);
    cx.launches += 2;
    if (cx.fetch_scalars()) return -1;
    w->primal_scale = w->dual_scale = cx.S_host->sigma;
  } else {
    if (cx.sync()) return -1;
  }
  w->setup_time = ms_since(t0);
  return 0;
}

extern "C" ScsWork *scs_init(const ScsData *d, const ScsCone *k, const ScsSettings *stgs) {  // scs.c:1193-1233
  if (!d || !k || !stgs) {
    B200_PRINTF("ERROR: Missing ScsData, ScsCone, or ScsSettings input\n");
    return nullptr;
  }
  if (validate_problem(d, k, stgs) < 0) {
    B200_PRINTF("ERROR: Validation returned failure\n");
    return nullptr;
  }
  const Clock::time_point t0 = Clock::now();
  SCS_WORK *w = new SCS_WORK();
  w->stgs = *stgs;
  const ScsData *d_full = d;
  const ScsCone *k_full = k;
  LocalProblem local;
  Dist *dd = dist_current();
  if (dd->on() || dd->selftest) {  // keep rows [row0, row0 + m) of A, b, the cones inside them and the columns they touch
    if (dd->world > kMaxWorld) { B200_PRINTF("ERROR: at most %d ranks\n", kMaxWorld); delete w; return nullptr; }
    if (partition_rows(d, k, dd->rank, dd->world, local)) { delete w; return nullptr; }
    w->dist = true; w->rank = dd->rank; w->world = dd->world;
    w->m_total = d->m; w->row0 = local.row0; w->n_total = d->n; w->n_sh = local.n_sh;
    w->cnt_lo = dd->rank == 0 ? 0 : local.n_sh;
    w->loc2glob = local.loc2glob;
    d = &local.d;
    k = &local.k;
    // every rank walks the same sequence of collectives: what rank 0 prints is decided by print_rank0 only
    w->print_rank0 = w->stgs.verbose && dd->rank == 0;
    w->verbose_collective = w->stgs.verbose != 0;
    w->stgs.verbose = 0;
    w->stgs.time_limit_secs = 0.;  // host clocks differ between ranks; the loop must stay collective
  }
  if (w->stgs.verbose || w->print_rank0) print_init_header(d_full, k_full, &w->stgs);
  w->n = d->n; w->m = d->m; w->l = d->n + d->m + 1;
  const int n = w->n, m = w->m, l = w->l;
  if (stgs->write_data_filename) {  // scs.c:1219-1222 (the full problem, before any partitioning)
    B200_PRINTF("Writing raw problem data to %s\n", stgs->write_data_filename);
    scs_b200_write_data(d_full, k_full, stgs);
    w->write_fn = stgs->write_data_filename;
  }
  if (stgs->log_csv_filename) {  // scs.c:1223-1226
    B200_PRINTF("Logging run data to %s\n", stgs->log_csv_filename);
    w->csv_fn = stgs->log_csv_filename;
  }
  w->stgs.write_data_filename = nullptr;  // the workspace keeps its own copies of the names
  w->stgs.log_csv_filename = nullptr;
  bool ok = false;
  do {
    if (w->c.init(current_device())) break;
    Ctx &c = w->c;
    cudaStream_t st = c.stream;
    if (w->dist) {
      c.dist = true; c.rank = w->rank; c.world = w->world; c.n_sh = w->n_sh; c.cnt_lo = w->cnt_lo;
      const int hdr[3] = {1, w->rank, w->world};  // DevScalars::dist, dist_rank, dist_world
      if (cudaMemcpyAsync(&c.S->dist, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st) != cudaSuccess) break;
      const int glen = w->m_total > w->n_total ? w->m_total : w->n_total;
      if (dev_alloc_zero(&w->gather, (size_t)glen, st)) break;
      if (dev_alloc(&w->d_loc2glob, w->loc2glob.size()) || h2d(c, w->d_loc2glob, w->loc2glob.data(), w->loc2glob.size())) break;
    }
    if (dev_alloc_zero(&w->u, (size_t)l, st) || dev_alloc_zero(&w->u_t, (size_t)l, st) ||
        dev_alloc_zero(&w->v, (size_t)l, st) || dev_alloc_zero(&w->v_prev, (size_t)l, st) ||
        dev_alloc_zero(&w->rsk, (size_t)l, st) || dev_alloc_zero(&w->g, (size_t)l, st) ||
        dev_alloc_zero(&w->diag_r, (size_t)l, st) || dev_alloc_zero(&w->b, (size_t)m, st) ||
        dev_alloc_zero(&w->cvec, (size_t)n, st) || dev_alloc(&w->D, (size_t)m) || dev_alloc(&w->E, (size_t)n) ||
        dev_alloc_zero(&w->ws, (size_t)n, st) || dev_alloc_zero(&w->sol_x, (size_t)n, st) ||
        dev_alloc_zero(&w->sol_y, (size_t)m, st) || dev_alloc_zero(&w->sol_s, (size_t)m, st)) {
      B200_PRINTF("ERROR: work memory allocation failure\n");
      break;
    }
    w->b_orig.assign((size_t)(w->dist ? w->m_total : m), 0.0);
    w->c_orig.assign((size_t)(w->dist ? w->n_total : n), 0.0);
    if (w->cone.init(&w->c, k, m)) { B200_PRINTF("ERROR: init_cone failure\n"); break; }
    if (set_diag_r(w)) break;
    if (w->ls.init(&w->c, d->A, d->P)) { B200_PRINTF("ERROR: init_lin_sys_work failure\n"); break; }
    w->ls.diag_r = w->diag_r;
    w->ls.own_diag_r = false;
    k_fill<<<ew_grid(c, m), kThreads, 0, st>>>(w->D, 1.0, m);
    k_fill<<<ew_grid(c, n), kThreads, 0, st>>>(w->E, 1.0, n);
    c.launches += 2;
    if (w->stgs.normalize) {
      if (normalize_a_p_dev(w)) { B200_PRINTF("ERROR: normalize_a_p failure\n"); break; }
    }
    {  // one-time box normalisation by D (cones.c:1549-1557)
      std::vector<double> Dh;
      if (w->cone.bsize > 1 && w->stgs.normalize) {
        Dh.resize((size_t)m);
        if (d2h(c, Dh.data(), w->D, (size_t)m) || c.sync()) break;
      }
      if (w->cone.normalize_box(Dh.empty() ? nullptr : Dh.data())) break;
    }
    if (w->ls.finalize_structure()) { B200_PRINTF("ERROR: tiled SpMV format failure\n"); break; }
    if (scs_update(w, d_full->b, d_full->c)) break;
    if (w->ls.update_precond()) break;
    if (w->stgs.acceleration_lookback) {
      if (w->aa.init(&w->c, l, w->stgs.acceleration_lookback, w->stgs.acceleration_lookback,
                     w->stgs.acceleration_type_1, w->stgs.acceleration_regularization,
                     w->stgs.acceleration_relaxation, kAaSafeguard, kAaMaxWeight, kAaIrSteps) == 0) {
        w->has_aa = w->aa.mem > 0;
        // row-partitioned mode: the shared block of x and tau are replicated; rank 0 counts them
        if (w->dist && w->has_aa && w->aa.set_counted_rows(w->cnt_lo, w->rank == 0 ? l : l - 1)) break;
      } else {
        // the reference continues without acceleration when aa_init fails (scs.c:1056-1058)
        if (w->stgs.verbose) B200_PRINTF("WARN: aa_init returned NULL, no acceleration applied.\n");
        w->aa.destroy();
        w->has_aa = false;
      }
    }
    if (c.sync()) break;
    if (cudaGetLastError() != cudaSuccess) break;
    if (build_iter_graph(w)) break;
    if (c.sync()) break;
    ok = true;
  } while (0);
  if (!ok) {
    free_work(w);
    return nullptr;
  }
  w->setup_time = ms_since(t0);
  return w;
}

extern "C" void scs_finish(ScsWork *w) { free_work(w); }

extern "C" scs_int scs_solve(ScsWork *w, ScsSolution *sol, ScsInfo *info, scs_int warm_start) {  // scs.c:1275-1430
  if (!sol || !w || !info) {
    B200_PRINTF("ERROR: missing ScsWork, ScsSolution or ScsInfo input\n");
    return SCS_FAILED;
  }
  Ctx &c = w->c;
  if (cudaSetDevice(c.device) != cudaSuccess) return SCS_FAILED;
  cudaStream_t st = c.stream;
  const int n = w->n, m = w->m, l = w->l;
  const int m_out = w->dist ? w->m_total : m;  // length of the caller's y and s
  const int n_out = w->dist ? w->n_total : n;  // ... and x
  ScsSettings *stgs = &w->stgs;
  stgs->warm_start = warm_start;
  start_interrupt_listener();
  const Clock::time_point t0 = Clock::now();
  strcpy(info->lin_sys_solver, scs_get_lin_sys_method());
  info->status_val = SCS_UNFINISHED;
  // ---- update_work: reset_tracking + warm/cold start + g (scs.c:1079-1105)
  w->last_scale_update_iter = 0; w->sum_log_scale_factor = 0.; w->n_log_scale_factor = 0; w->scale_updates = 0;
  w->time_limit_reached = 0;
  w->unroll_cg0 = -1; w->unroll_it0 = 0;
  w->r_n.last_iter = -1; w->r_o.last_iter = -1;
  cudaMemsetAsync(&c.S->aa_rejected, 0, 2 * sizeof(int), st);  // safeguard counters of this solve
  if (warm_start) {
    if (!sol->x || !sol->y || !sol->s) {
      return failure(w, m_out, n_out, sol, info, SCS_FAILED, "warm-start requested without x, y, s", "failure");
    }
    const double *xsrc = sol->x;
    if (w->dist) {  // this rank's columns of the caller's full x
      w->x_loc.resize((size_t)n);
      for (int j = 0; j < n; ++j) w->x_loc[(size_t)j] = sol->x[w->loc2glob[(size_t)j]];
      xsrc = w->x_loc.data();
    }
    if (h2d(c, w->sol_x, xsrc, (size_t)n) || h2d(c, w->sol_y, sol->y + w->row0, (size_t)m) ||
        h2d(c, w->sol_s, sol->s + w->row0, (size_t)m))
      return failure(w, m_out, n_out, sol, info, SCS_FAILED, "warm-start upload", "failure");
    k_warm_start<<<ew_grid(c, l), kThreads, 0, st>>>(w->v, w->sol_x, w->sol_y, w->sol_s, w->D, w->E, w->diag_r, n, m,
                                                     w->primal_scale, w->dual_scale);
  } else {
    k_cold_start<<<ew_grid(c, l), kThreads, 0, st>>>(w->v, l);
  }
  c.launches++;
  if (update_work_cache(w)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in update_work_cache", "failure");
  // row-partitioned mode: `vwork` (same on every rank) decides whether the verbose-only residual passes --
  // which carry collectives -- run; `vprint` (rank 0 only) whether their results are printed
  const bool vwork = stgs->verbose || w->verbose_collective, vprint = stgs->verbose || w->print_rank0;
  if (vprint) print_header();

  const int gl = ew_grid(c, l);
  const bool accel = w->has_aa;
  const bool csv = !w->csv_fn.empty() && !w->dist;  // the row-partitioned mode writes no trace
  const int interval = stgs->acceleration_interval;
  {  // device-side loop state of this solve: iteration index and phase clocks
    static const unsigned long long zeros[5] = {0, 0, 0, 0, 0};
    cudaMemsetAsync(&c.S->iter, 0, sizeof(int), st);
    cudaMemcpyAsync(&c.S->t_mark, zeros, sizeof(zeros), cudaMemcpyHostToDevice, st);
  }
  int i;
  for (i = 0; i < stgs->max_iters; ++i) {
    if (w->mark_begin >= 0) {
      if (i == w->mark_begin && !w->mk_open && !w->mk_done) mark_open(w, i);
      if (i == w->mark_end && w->mk_open) mark_close(w, i);
    }
    // ---- Anderson acceleration (scs.c:1306-1313)
    if (accel && i > 0 && i % interval == 0) {
      k_stamp<<<1, 1, 0, st>>>(c.S, -1);
      if (w->aa.apply(w->v, w->v_prev, &c.S->vnorm2))
        return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in aa_apply", "failure");
      k_stamp<<<1, 1, 0, st>>>(c.S, 2);
      c.launches += 2;
      w->n_aa++;
    }
    // ---- linear system + cones (scs.c:1326-1340): one graph launch, or the same kernels
    //      enqueued on the stream with the CG loop driven from the host
    if (w->use_graph) {
      if (cudaGraphLaunch(w->gexec, st) != cudaSuccess)
        return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error launching the iteration graph", "failure");
      c.launches += w->graph_launches;
      c.spmv_calls += w->graph_spmv;
    } else {
      if (enqueue_front_head(w, CgCtl{}) || w->ls.solve_dev_loop(w->u_t, 0) || enqueue_front_tail(w))
        return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in project_lin_sys / project_cones", "failure");
    }

    const bool check = (i % kConvergedInterval == 0);
    const bool print = vwork && (i % kPrintInterval == 0);
    if (check || print) {
      k_post<1><<<gl, kThreads, 0, st>>>(w->v, w->rsk, w->u, w->u_t, w->diag_r, n, l, w->cnt_lo, stgs->alpha, c.red, c.S);
      c.launches++;
      if (check && g_int_detected && !w->dist) return failure(w, m_out, n_out, sol, info, SCS_SIGINT, "interrupted", "interrupted");
      if (populate_residuals(w, i)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in residuals", "failure");
      w->n_checks++;
      if (w->use_graph && w->cg_unroll_auto && (w->unroll_cg0 < 0 || i <= w->unroll_it0)) {
        w->unroll_cg0 = c.S_host->cg_its_total; w->unroll_it0 = i;  // baseline (start of a solve, after a scale update)
      } else if (w->use_graph && w->cg_unroll_auto) {
        // follow the CG iteration count: enough unrolled nodes for the typical ADMM iteration, no more
        const double kavg = (double)(c.S_host->cg_its_total - w->unroll_cg0) / (double)(i - w->unroll_it0);
        int want = (int)ceil(kavg + 0.25);
        want = want < 1 ? 1 : (want > 24 ? 24 : want);
        w->unroll_cg0 = c.S_host->cg_its_total; w->unroll_it0 = i;
        if (want != w->cg_unroll) {
          w->cg_unroll = want;
          if (rebuild_iter_graph(w)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error re-capturing the iteration graph", "failure");
        }
      }
      if (check) {
        if ((info->status_val = has_converged(w)) != 0) break;
        if (stgs->time_limit_secs && ms_since(t0) > 1000. * stgs->time_limit_secs) {
          w->time_limit_reached = 1;
          break;
        }
      }
      if (print && vprint) print_summary(w, i, t0);
      if (stgs->adaptive_scale && i == w->r_o.last_iter) {
        if (update_scale(w, i) < 0) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in update_scale", "failure");
      }
      k_post<2><<<gl, kThreads, 0, st>>>(w->v, w->rsk, w->u, w->u_t, w->diag_r, n, l, w->cnt_lo, stgs->alpha, c.red, c.S);
      if (dist_finish(c, 1, 0, FinPost{})) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in the dual step", "failure");
      c.launches++;
    } else {
      k_post<0><<<gl, kThreads, 0, st>>>(w->v, w->rsk, w->u, w->u_t, w->diag_r, n, l, w->cnt_lo, stgs->alpha, c.red, c.S);
      if (dist_finish(c, 1, 0, FinPost{})) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in the dual step", "failure");
      c.launches++;
    }
    // ---- AA safeguard (scs.c:1386-1394); the aa_norm > 0 gate is evaluated on the device
    if (accel && i > 0 && i % interval == 0) {
      k_stamp<<<1, 1, 0, st>>>(c.S, -1);
      if (w->aa.safeguard(w->v, w->v_prev, &c.S->vnorm2, &c.S->aa_rejected, &c.S->aa_accepted))
        return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in aa_safeguard", "failure");
      k_stamp<<<1, 1, 0, st>>>(c.S, 2);
      c.launches += 2;
    }
    // ---- CSV trace, after the scale update so that the extra residual pass does not touch the
    //      algorithm (scs.c:1396-1401)
    if (csv && csv_log(w, i, t0)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in the CSV trace", "failure");
  }
  if (csv && csv_log(w, i, t0)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in the CSV trace", "failure");
  if (w->mk_open) mark_close(w, i);
  w->mark_begin = w->mark_end = -1;
  w->admm_iters += i;
  if (vwork) {
    if (populate_residuals(w, i)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in residuals", "failure");
    if (vprint) print_summary(w, i, t0);
  }
  // ---- finalize (scs.c:874-924)
  if (!sol->x) sol->x = (double *)calloc(n_out, sizeof(double));
  if (!sol->y) sol->y = (double *)calloc(m_out, sizeof(double));
  if (!sol->s) sol->s = (double *)calloc(m_out, sizeof(double));
  k_finalize_sol<<<ew_grid(c, n + m), kThreads, 0, st>>>(w->sol_x, w->sol_y, w->sol_s, w->u, w->rsk, w->D, w->E, n, m,
                                                         w->primal_scale, w->dual_scale, c.red, c.S);
  c.launches++;
  if (dist_finish(c, 1, 2, FinSol{})) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in finalize", "failure");
  if (populate_residuals(w, i)) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "error in residuals", "failure");
  if (c.fetch_scalars()) return failure(w, m_out, n_out, sol, info, SCS_FAILED, "fetch", "failure");
  w->ls.tot_cg_its = c.S_host->cg_its_total;
  const double nm_s = c.S_host->fin[0], nm_y = c.S_host->fin[1], sty = c.S_host->fin[2];
  const HostResid &r = w->r_o;
  info->setup_time = w->setup_time;
  info->iter = i;
  info->res_infeas = r.res_infeas;
  info->res_unbdd_a = r.res_unbdd_a;
  info->res_unbdd_p = r.res_unbdd_p;
  info->scale = stgs->scale;
  info->scale_updates = w->scale_updates;
  info->rejected_accel_steps = c.S_host->aa_rejected;
  info->accepted_accel_steps = c.S_host->aa_accepted;
  fill_aa_stats(w, info);
  info->comp_slack = fabs(sty);
  if (info->comp_slack > 1e-5 * fmax(nm_s, nm_y))
    B200_PRINTF("WARNING - large complementary slackness residual: %f\n", info->comp_slack);
  double fx = 1.0, fy = 1.0, fs = 1.0;
  auto set_solved = [&]() {  // scs.c:805-816
    fx = fy = fs = safediv_pos_h(1.0, r.tau);
    info->gap = r.gap; info->res_pri = r.res_pri; info->res_dual = r.res_dual;
    info->pobj = r.xt_p_x / 2. + r.ctx;
    info->dobj = -r.xt_p_x / 2. - r.bty;
    strcpy(info->status, "solved");
    info->status_val = SCS_SOLVED;
  };
  auto set_infeasible = [&]() {  // scs.c:818-829
    fy = -1 / r.bty_tau; fx = NAN; fs = NAN;
    info->gap = NAN; info->res_pri = NAN; info->res_dual = NAN; info->pobj = INFINITY; info->dobj = INFINITY;
    strcpy(info->status, "infeasible");
    info->status_val = SCS_INFEASIBLE;
  };
  auto set_unbounded = [&]() {  // scs.c:831-842
    fx = -1 / r.ctx_tau; fs = -1 / r.ctx_tau; fy = NAN;
    info->gap = NAN; info->res_pri = NAN; info->res_dual = NAN; info->pobj = -INFINITY; info->dobj = -INFINITY;
    strcpy(info->status, "unbounded");
    info->status_val = SCS_UNBOUNDED;
  };
  switch (info->status_val) {
    case SCS_SOLVED: set_solved(); break;
    case SCS_INFEASIBLE: set_infeasible(); break;
    case SCS_UNBOUNDED: set_unbounded(); break;
    case SCS_UNFINISHED:  // set_unfinished, scs.c:845-871
      if (r.kap > r.tau && (r.bty_tau < 0 || r.ctx_tau < 0)) {
        if (r.bty_tau < 0 && r.bty_tau < r.ctx_tau) { set_infeasible(); info->status_val = SCS_INFEASIBLE_INACCURATE; }
        else { set_unbounded(); info->status_val = SCS_UNBOUNDED_INACCURATE; }
      } else if (r.tau > 0) {
        set_solved();
        info->status_val = SCS_SOLVED_INACCURATE;
      } else {
        B200_PRINTF("ERROR: could not determine problem status.\n");
        strcpy(info->status, "failure");
        info->status_val = SCS_FAILED;
      }
      if (w->time_limit_reached) strcat(info->status, " (inaccurate - reached time_limit_secs)");
      else if (info->iter >= stgs->max_iters) strcat(info->status, " (inaccurate - reached max_iters)");
      else B200_PRINTF("ERROR: should not be in this state (1).\n");
      break;
    default: B200_PRINTF("ERROR: should not be in this state (2).\n");
  }
  k_scale3<<<ew_grid(c, n + m), kThreads, 0, st>>>(w->sol_x, n, fx, w->sol_y, w->sol_s, m, fy, fs);
  c.launches++;
  if (w->dist) {
    // every rank returns the full y and s: own slice into a zeroed buffer, summed over the ranks
    for (int which = 0; which < 2; ++which) {
      const double *src = which == 0 ? w->sol_y : w->sol_s;
      double *dst = which == 0 ? sol->y : sol->s;
      if (cudaMemsetAsync(w->gather, 0, sizeof(double) * (size_t)m_out, st) != cudaSuccess ||
          cudaMemcpyAsync(w->gather + w->row0, src, sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
          dist_allreduce(c, w->gather, (size_t)m_out, 0) || d2h(c, dst, w->gather, (size_t)m_out) || c.sync())
        return failure(w, m_out, n_out, sol, info, SCS_FAILED, "solution gather", "failure");
    }
    // x: every rank scatters the columns it counts (rank 0: shared + private, others: private) into a
    // zeroed full-length buffer; the sum over the ranks is the full x
    if (cudaMemsetAsync(w->gather, 0, sizeof(double) * (size_t)w->n_total, st) != cudaSuccess)
      return failure(w, m_out, n_out, sol, info, SCS_FAILED, "solution gather", "failure");
    k_scatter_cols<<<ew_grid(c, n), kThreads, 0, st>>>(w->gather, w->sol_x, w->d_loc2glob, w->cnt_lo, n);
    c.launches++;
    if (dist_allreduce(c, w->gather, (size_t)w->n_total, 0) || d2h(c, sol->x, w->gather, (size_t)w->n_total) || c.sync())
      return failure(w, m_out, n_out, sol, info, SCS_FAILED, "solution download", "failure");
  } else if (d2h(c, sol->x, w->sol_x, (size_t)n) || d2h(c, sol->y, w->sol_y, (size_t)m) ||
             d2h(c, sol->s, w->sol_s, (size_t)m) || c.sync())
    return failure(w, m_out, n_out, sol, info, SCS_FAILED, "solution download", "failure");
  info->solve_time = ms_since(t0);
  info->lin_sys_time = c.S_host->phase_ns[0] * 1e-6;  // %globaltimer phase clocks, see phase_lap()
  info->cone_time = c.S_host->phase_ns[1] * 1e-6;
  info->accel_time = c.S_host->phase_ns[2] * 1e-6;
  if (vprint) print_footer(info);
  end_interrupt_listener();
  return info->status_val;
}

extern "C" scs_int scs(const ScsData *d, const ScsCone *k, const ScsSettings *stgs, ScsSolution *sol,
                       ScsInfo *info) {  // scs.c:1483-1496
  scs_int status;
  ScsWork *w = scs_init(d, k, stgs);
  if (w) {
    scs_solve(w, sol, info, stgs->warm_start);
    status = info->status_val;
  } else {
    status = failure(nullptr, d ? d->m : -1, d ? d->n : -1, sol, info, SCS_FAILED, "could not initialize work",
                     "failure");
  }
  scs_finish(w);
  return status;
}

// Host-only view of the row partition a rank would own (used by the CPU test tier).
extern "C" scs_int scs_b200_dist_partition(const ScsData *d, const ScsCone *k, scs_int rank, scs_int world,
                                           scs_int out[12]) {
  if (!d || !k || !out || !d->A || world < 1 || rank < 0 || rank >= world) return -1;
  LocalProblem lp;
  if (world == 1) {
    out[0] = 0; out[1] = d->m; out[2] = d->A->p[d->n]; out[3] = k->z; out[4] = k->l; out[5] = k->bsize;
    out[6] = k->qsize; out[7] = k->ssize; out[8] = k->cssize; out[9] = k->ep; out[10] = k->ed; out[11] = k->psize;
    return 0;
  }
  if (partition_rows(d, k, rank, world, lp)) return -1;
  out[0] = lp.row0; out[1] = lp.m; out[2] = lp.Ap[(size_t)lp.n_loc]; out[3] = lp.k.z; out[4] = lp.k.l; out[5] = lp.k.bsize;
  out[6] = lp.k.qsize; out[7] = lp.k.ssize; out[8] = lp.k.cssize; out[9] = lp.k.ep; out[10] = lp.k.ed; out[11] = lp.k.psize;
  return 0;
}

// Host-only: the local problem rank `rank` of `world` would build (CPU test tier).  sizes = {row0, m_loc,
// n_shared, n_loc, nnz(A_loc), nnz(P_loc)}; every array pointer may be NULL (first call: sizes only).
extern "C" scs_int scs_b200_dist_local(const ScsData *d, const ScsCone *k, scs_int rank, scs_int world, scs_int sizes[6],
                                       scs_int *loc2glob, scs_int *Ap, scs_int *Ai, scs_float *Ax, scs_int *Pp,
                                       scs_int *Pi, scs_float *Px, scs_float *c_loc) {
  if (!d || !k || !sizes || !d->A || world < 1 || rank < 0 || rank >= world) return -1;
  LocalProblem lp;
  if (partition_rows(d, k, rank, world, lp)) return -1;
  const int nl = lp.n_loc, nzA = lp.Ap[(size_t)nl], nzP = d->P ? lp.Pp[(size_t)nl] : 0;
  sizes[0] = lp.row0; sizes[1] = lp.m; sizes[2] = lp.n_sh; sizes[3] = nl; sizes[4] = nzA; sizes[5] = nzP;
  if (loc2glob) memcpy(loc2glob, lp.loc2glob.data(), sizeof(int) * (size_t)nl);
  if (Ap) memcpy(Ap, lp.Ap.data(), sizeof(int) * ((size_t)nl + 1));
  if (Ai && nzA) memcpy(Ai, lp.Ai.data(), sizeof(int) * (size_t)nzA);
  if (Ax && nzA) memcpy(Ax, lp.Ax.data(), sizeof(double) * (size_t)nzA);
  if (d->P) {
    if (Pp) memcpy(Pp, lp.Pp.data(), sizeof(int) * ((size_t)nl + 1));
    if (Pi && nzP) memcpy(Pi, lp.Pi.data(), sizeof(int) * (size_t)nzP);
    if (Px && nzP) memcpy(Px, lp.Px.data(), sizeof(double) * (size_t)nzP);
  }
  if (c_loc) memcpy(c_loc, lp.c.data(), sizeof(double) * (size_t)nl);
  return 0;
}

// root_plus (scs.c:667-688) on host buffers: the device kernel of the ADMM loop (k_rootplus + FinRootPlus) on
// p = u_t, mu = v, g, r = diag_r (nm entries each), tau_scale = diag_r[nm], eta = v[nm].  Parity-test surface
// for the reference's known-answer test S/test/problems/test_root_plus.h.
extern "C" scs_float scs_b200_root_plus(const scs_float *g, const scs_float *p, const scs_float *mu, const scs_float *r,
                                        scs_int nm, scs_float tau_scale, scs_float eta) {
  if (!g || !p || !mu || !r || nm <= 0) return NAN;
  Ctx c;
  if (c.init(current_device())) return NAN;
  double *dg = nullptr, *dp = nullptr, *dmu = nullptr, *dr = nullptr;
  double tau = NAN;
  do {
    if (dev_alloc(&dg, (size_t)nm) || dev_alloc(&dp, (size_t)nm) || dev_alloc(&dmu, (size_t)nm + 1) ||
        dev_alloc(&dr, (size_t)nm + 1))
      break;
    if (h2d(c, dg, g, (size_t)nm) || h2d(c, dp, p, (size_t)nm) || h2d(c, dmu, mu, (size_t)nm) || h2d(c, dmu + nm, &eta, 1) ||
        h2d(c, dr, r, (size_t)nm) || h2d(c, dr + nm, &tau_scale, 1))
      break;
    const int one = kFeasibleIters;  // iteration >= FEASIBLE_ITERS: the quadratic is evaluated (scs.c:722-724)
    if (cudaMemcpyAsync(&c.S->iter, &one, sizeof(int), cudaMemcpyHostToDevice, c.stream) != cudaSuccess) break;
    k_rootplus<<<ew_grid(c, nm), kThreads, 0, c.stream>>>(dp, dmu, dg, dr, 0, nm, 0, c.red, c.S);
    if (cudaGetLastError() != cudaSuccess || c.fetch_scalars()) break;
    tau = c.S_host->tau;
  } while (0);
  dev_free(dg); dev_free(dp); dev_free(dmu); dev_free(dr);
  c.destroy();
  return tau;
}

// ------------------------------------------------------------------ measurement -------
extern "C" scs_int scs_b200_get_stats(const ScsWork *w, ScsB200Stats *out) {
  if (!w || !out) return -1;
  out->kernel_launches = total_launches(w, w->ls.tot_cg_its);
  out->cg_iters = w->ls.tot_cg_its;
  out->admm_iters = w->admm_iters;
  out->spmv_calls = w->c.spmv_calls;
  out->spmv_ms = 0.0;
  out->algorithmic_bytes = model_bytes(w, w->admm_iters, w->ls.tot_cg_its, w->n_checks, w->n_aa);
  out->h2d_bytes = w->c.h2d;
  out->d2h_bytes = w->c.d2h;
  out->collectives = w->c.collectives;
  out->collective_bytes = w->c.collective_bytes;
  out->tiled_a = w->ls.tA.ok ? 1 : 0;
  out->tiled_g = w->ls.tG.ok ? 1 : 0;
  out->tiled_slots = (w->ls.tA.ok ? w->ls.tA.slots : 0) + (w->ls.tG.ok ? w->ls.tG.slots : 0);
  out->tiled_nnz = (w->ls.tA.ok ? w->ls.tA.nnz : 0) + (w->ls.tG.ok ? w->ls.tG.nnz : 0);
  return 0;
}

extern "C" scs_int scs_b200_set_marks(ScsWork *w, scs_int begin_iter, scs_int end_iter) {
  if (!w || begin_iter < 0 || end_iter <= begin_iter) return -1;
  if (cudaSetDevice(w->c.device) != cudaSuccess) return -1;
  if (!w->mark_ev[0]) {
    if (cudaEventCreate(&w->mark_ev[0]) != cudaSuccess || cudaEventCreate(&w->mark_ev[1]) != cudaSuccess) return -1;
  }
  w->mark_begin = begin_iter;
  w->mark_end = end_iter;
  w->mk_open = false;
  w->mk_done = false;
  memset(&w->marks, 0, sizeof(w->marks));
  return 0;
}
extern "C" scs_int scs_b200_get_marks(const ScsWork *w, ScsB200Marks *out) {
  if (!w || !out || !w->mk_done) return -1;
  *out = w->marks;
  return 0;
}

extern "C" double scs_b200_bench_spmv(ScsWork *w, scs_int which, scs_int reps, double *alg_bytes) {
  if (!w || reps <= 0) return -1.0;
  Ctx &c = w->c;
  if (cudaSetDevice(c.device) != cudaSuccess) return -1.0;
  LinSys &ls = w->ls;
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
  // operate on CG scratch only: p (n) -> tmp (m) -> Gp (n); ADMM state is untouched
  k_fill<<<ew_grid(c, ls.n), kThreads, 0, c.stream>>>(ls.p, 1.0, ls.n);
  k_fill<<<ew_grid(c, ls.m), kThreads, 0, c.stream>>>(ls.tmp, 1.0, ls.m);
  for (int wu = 0; wu < 2; ++wu) {
    if (which == 0) ls.launch_A_scaled(ls.p, ls.tmp, nullptr);
    else ls.launch_G(ls.tmp, ls.p, ls.Gp, nullptr);
  }
  cudaEventRecord(e0, c.stream);
  for (int r = 0; r < reps; ++r) {
    if (which == 0) ls.launch_A_scaled(ls.p, ls.tmp, nullptr);
    else ls.launch_G(ls.tmp, ls.p, ls.Gp, nullptr);
  }
  cudaEventRecord(e1, c.stream);
  if (cudaStreamSynchronize(c.stream) != cudaSuccess) return -1.0;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (alg_bytes) *alg_bytes = which == 0 ? ls.bytes_A() : (ls.bytes_At() + ls.bytes_P());
  return (double)ms / reps;
}

// per-CTA profile of the last launch of a tiled operator (which = 0: A, 1: [A' | P]); out receives
// 4 doubles per CTA: duration us, us spent streaming non-zeros, modelled cost of its item list, items
extern "C" scs_int scs_b200_tiled_profile(ScsWork *w, scs_int which, double *out, scs_int cap) {
  if (!w || !out) return -1;
  const TiledOp &op = which == 0 ? w->ls.tA : w->ls.tG;
  if (!op.ok) return 0;
  const int ncta = op.d.ncta;
  std::vector<unsigned long long> h((size_t)3 * ncta);
  if (cudaSetDevice(w->c.device) != cudaSuccess || cudaStreamSynchronize(w->c.stream) != cudaSuccess ||
      cudaMemcpy(h.data(), op.d.prof, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  for (int b = 0; b < ncta && b < cap; ++b) {
    out[4 * b + 0] = (double)(h[3 * b + 1] - h[3 * b + 0]) * 1e-3;
    out[4 * b + 1] = (double)h[3 * b + 2] * 1e-3;
    out[4 * b + 2] = op.cta_cost[b];
    out[4 * b + 3] = (double)op.cta_items[b];
  }
  return ncta;
}
