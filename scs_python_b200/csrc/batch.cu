// batch.cu -- batch engine: one CTA solves one small problem start to finish.
//
// BASELINE.json configs[4] (8192 independent MPC QPs, n = 120, m = 360) and SURVEY.md 8(d)/(e):
// a problem of this size moves ~0.1 MB per ADMM iteration, so the streaming engine of
// scs_solver.cu (≈25 dependent kernel launches per iteration) is launch-latency bound on it.
// Here the whole of scs_init + scs_solve (S/src/scs.c:1193-1430) for one problem runs inside
// ONE CTA with every matrix and vector resident in shared memory:
//
//   equilibration        normalize_a_p / normalize_b_c   (scs_matrix.c:203-470, normalize.c:33-61)
//   linear system        warm-started diagonal-PCG on R_x + P + A' R_y^-1 A
//                        (linsys/cpu/indirect/private.c:50-316), CG tolerance rule scs.c:703-720
//   tau root             root_plus (scs.c:667-688)
//   cones                zero / nonneg inline, box cone by the CTA (Newton on t, cones.c:1174-1237), one warp per
//                        second-order cone (cones.c:1242-1271), one thread per exponential / power cone (cone_dev.cuh), one warp per PSD / complex PSD cone (Jacobi, real order <= 32),
//                        Moreau wrapper cones.c:1544-1588
//   ADMM vector updates  scs.c:739-779
//   residuals / stop     populate_residual_struct, has_converged, update_scale (scs.c:441-627, 1112-1189)
//   Anderson accel.      aa_apply / aa_safeguard (aa.c:822-901); ring buffers S, Y, D in a per-CTA
//                        global workspace (L2-resident), Householder elimination with thread-owned
//                        rows, small pivoted-QR / LU by one thread (aa_small.cuh)
//   finalisation         scs.c:805-924
//
// The grid is persistent: min(count, SMs x resident CTAs) CTAs pull problem indices from an atomic
// counter, so uneven iteration counts do not leave SMs idle.  No host synchronisation happens
// between the upload of the packed batch and the download of the solutions.
// Problems the CTA cannot hold (shared-memory footprint, PSD cones of order > 32 / complex ones of order > 16, warm
// start, time limit, AA relaxation != 1, lookback > 10) are solved by the streaming engine, one
// after another -- still on the GPU, never on the host.
#include <algorithm>
#include <chrono>
#include <numeric>

#include "aa_small.cuh"
#include "cone_dev.cuh"
#include "solver_internal.cuh"

namespace b200 {
int current_device();
namespace {

typedef unsigned short u16;
constexpr int kBT = 256;            // threads per CTA
constexpr int kBW = kBT / 32;       // warps per CTA
constexpr int kBAaMax = 10;         // lookback handled in-CTA
constexpr int kBC = 2 * kBAaMax + 1;  // columns of the AA elimination [A | Y | g]
constexpr int kBRed = 24;           // reduction outputs per call (>= kBC)
constexpr int kBCtasPerSm = 2;

// glbopts.h constants
constexpr int kFeasIters = 1, kRescaleMinIters = 100, kConvInterval = 25;
constexpr double kDivEpsB = 1e-18, kTauFactorB = 10.0, kInfeasNegTolB = 1e-9;
constexpr double kMaxScaleB = 1e6, kMinScaleB = 1e-6, kCgBestTolB = 1e-12, kCgTolFactorB = 0.2, kCgRateB = 1.5;
constexpr double kMinNormB = 1e-4, kMaxNormB = 1e4;
constexpr int kRuizB = 25, kL2B = 1;

struct BDims { int n, m, nnzA, nnzP, nq, mem, direct, nb, np, tri, ds, pslots; };  // maxima over the batch; direct: dense inverse resident; nb: box bounds
                                                                    // (bsize - 1); nq: cones with a boundary (SOC + exp + power); np: power cones;
                                                                    // tri: some member has exp / power / PSD cones; ds: largest PSD order;
                                                                    // pslots: warps that project PSD cones at a time (one workspace each)
struct BStg {
  int normalize, adaptive_scale, max_iters, aa_mem, aa_interval, aa_type1, refine, pad;
  double scale, rho_x, eps_abs, eps_rel, eps_infeas, alpha, aa_reg;
};
struct BProb {
  int n, m, nnzA, nnzP, z, l, nq, bsize;  // bsize: box cone rows [t; s] right after the nonneg rows (cones.c:1174-1237)
  int nsoc, ns, ep, ed, np, ncs;          // nq = nsoc second-order cones, ns PSD + ncs complex PSD cones (packed), ep + ed exponential and
                                          // np power cones (3 rows each), in the reference's cone order (S/include/scs.h ScsCone)
  long long d_off, i_off, sol_off;
};
struct BOut {
  int iter, status_val, scale_updates, rej, acc, cg_its;
  int n_accept, n_reject_rank0, n_reject_nonfinite, n_reject_weight_cap, n_safeguard_reject, last_rank;
  int aa_iter, pad;
  double pobj, dobj, res_pri, res_dual, gap, res_infeas, res_unbdd_a, res_unbdd_p, scale, comp_slack;
  double nm_s, nm_y, last_aa_norm, last_reg, setup_ms, solve_ms;
  long long clk[6];  // cycles: equilibrate, factor, lin-sys, AA, residual checks, whole problem
};

// PSD cones: per-warp Jacobi workspace (matrix, eigenvectors, one rotation per pair of a round), order <= kBPsdMax
constexpr int kBPsdMax = 32;
__host__ __device__ inline int psd_ws_doubles(int d) { return 2 * d * d + 2 * ((d + 1) / 2) + 2; }
// shared-memory carve-up, identical on host and device (offsets in doubles / u16 elements)
struct BLay {
  int Aval, AvalR, Pval, u, ut, v, vp, rsk, g, dr, b, c, D, E, cp, cr, cGp, cM, tmp, ws, red, aaR, aaScr, bl, bu, pw, psdw, Ginv, nd;
  int st_bytes;
  int Arow, Aperm, Acol, Acp, Arp, Pcol, Prp, qoff, qlen, ni;
  __host__ __device__ explicit BLay(const BDims &d) {
    const int l = d.n + d.m + 1;
    int o = 0;
    auto take = [&o](int cnt) { int r = o; o += cnt; return r; };
    Aval = take(d.nnzA); AvalR = take(d.nnzA); Pval = take(d.nnzP);  // AvalR: the values once more, in CSR order
    u = take(l); ut = take(l); v = take(l); vp = take(d.mem > 0 ? l : 0); rsk = take(l); g = take(l);
    dr = take(l); b = take(d.m); c = take(d.n); D = take(d.m); E = take(d.n);
    cp = take(d.n); cr = take(d.n); cGp = take(d.n); cM = take(d.n); tmp = take(d.m); ws = take(d.n);
    red = take(2 * kBW * kBRed);
    aaR = take(d.mem > 0 ? d.mem * (2 * d.mem + 1) : 0);
    aaScr = take(d.mem > 0 ? 4 * d.mem * d.mem + 6 * d.mem + 8 : 0);
    bl = take(d.nb); bu = take(d.nb); pw = take(d.np);
    psdw = take(d.ds > 0 ? d.pslots * psd_ws_doubles(d.ds) : 0);
    Ginv = take(d.direct ? d.n * (d.n | 1) : 0);
    nd = o;
    st_bytes = (int)((sizeof(AaState) + 15) / 16 * 16);
    o = 0;
    Arow = take(d.nnzA); Aperm = take(d.nnzA); Acol = take(d.nnzA); Acp = take(d.n + 1); Arp = take(d.m + 1);
    Pcol = take(d.nnzP); Prp = take(d.n + 1); qoff = take(d.nq); qlen = take(d.nq);
    ni = o;
  }
  __host__ __device__ size_t bytes() const { return (size_t)nd * 8 + st_bytes + (size_t)ni * 2 + 16; }
};
// per-problem element counts of the packed pools
static inline long long dpool_count(int n, int m, int nnzA, int nnzP, int nb, int np) { return (long long)nnzA + nnzP + m + n + 2ll * nb + np; }
static inline long long ipool_count(int n, int m, int nnzA, int nnzP, int nq) {
  return 3ll * nnzA + (n + 1) + (m + 1) + nnzP + (n + 1) + 2ll * nq;
}
// AA workspace (doubles) of one CTA slot
static inline size_t aaws_count(const BDims &d) {
  const size_t l = (size_t)d.n + d.m + 1;
  return d.mem > 0 ? 3 * l * d.mem + 4 * l + (l + d.mem) * (2 * (size_t)d.mem + 1) : 0;
}

struct BArgs {
  const BProb *probs;
  int count;
  BDims dims;
  BStg stg;
  const double *dpool;
  const u16 *ipool;
  double *sol;
  BOut *out;
  double *aaws;
  size_t aaws_stride;
  int *counter;
};

// ------------------------------------------------------------------------------ device ---
struct Red { double *buf; int phase; };

// CTA-wide reduction of KS sums followed by KM maxes; every thread gets every result.
// One barrier per call: the two halves of buf alternate, and a third call can only start
// writing the half used by the first after all threads passed the second call's barrier.
template <int KS, int KM>
__device__ __forceinline__ void breduce(double (&v)[KS + KM], Red &r) {
  constexpr int K = KS + KM;
  static_assert(K <= kBRed, "reduction too wide");
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = (k < KS) ? warp_sum(v[k]) : warp_max(v[k]);
  double *b = r.buf + r.phase * (kBW * kBRed);
  r.phase ^= 1;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) b[w * K + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double acc = b[k];
#pragma unroll
    for (int ww = 1; ww < kBW; ++ww) acc = (k < KS) ? acc + b[ww * K + k] : fmax(acc, b[ww * K + k]);
    v[k] = acc;
  }
}

__device__ __forceinline__ double safediv_b(double x, double y) { return y < kDivEpsB ? x / kDivEpsB : x / y; }
__device__ __forceinline__ double inv_sqrt_limited(double x) {  // scs_matrix.c:203-208 + the sqrt/inverse of its callers
  if (x < kMinNormB) x = 1.0;
  else if (x > kMaxNormB) x = kMaxNormB;
  x = sqrt(x);
  return x < kDivEpsB ? 1.0 / kDivEpsB : 1.0 / x;
}

struct Resid {  // ScsResiduals scalars in the ORIGINAL scaling (scs_work.h:29-50) + the two normalised norms
  int last_iter;
  double tau, kap, bty_tau, ctx_tau, xpx_tau, bty, ctx, xpx, gap, pobj, dobj;
  double res_pri, res_dual, res_infeas, res_unbdd_a, res_unbdd_p;
  double o_ax, o_s, o_px, o_aty, o_ax_s_btau, o_px_aty_ctau;
  double n_ax_s_btau, n_px_aty_ctau;
};

struct B {  // one CTA's view of its problem
  int n, m, l, nnzA, nnzP, z, nl, nq, tid;
  double *Aval, *AvalR, *Pval, *u, *ut, *v, *vp, *rsk, *g, *dr, *b, *c, *D, *E, *cp, *cr, *cGp, *cM, *tmp, *ws;
  double *aaR, *aaScr, *Ginv, *bl, *bu, *pw, *psdw;
  int bsize, nsoc, ns, ncs, ep, ed, np, psd_ws;
  int direct, refine, gld, gparts, gshift;
  u16 *Arow, *Aperm, *Acol, *Acp, *Arp, *Pcol, *Prp, *qoff, *qlen;
  AaState *st;
  // AA global workspace
  double *aS, *aY, *aD, *ax, *af, *ag, *agp, *aW;
  int ald;
  Red red;
  int cg_its;
};

#define BFOR(i, N) for (int i = s.tid; i < (N); i += kBT)

__device__ __forceinline__ double rowdot_A(const B &s, int i, const double *x) {  // (A x)_i, CSR view
  double acc = 0.0;
  for (int k = s.Arp[i]; k < s.Arp[i + 1]; ++k) acc = fma(s.AvalR[k], x[s.Acol[k]], acc);  // (valid after equilibrate())
  return acc;
}
__device__ __forceinline__ double coldot_A(const B &s, int j, const double *y) {  // (A' y)_j, CSC as given
  double acc = 0.0;
  for (int k = s.Acp[j]; k < s.Acp[j + 1]; ++k) acc = fma(s.Aval[k], y[s.Arow[k]], acc);
  return acc;
}
__device__ __forceinline__ double rowdot_P(const B &s, int j, const double *x) {  // (P x)_j, full symmetric
  double acc = 0.0;
  for (int k = s.Prp[j]; k < s.Prp[j + 1]; ++k) acc = fma(s.Pval[k], x[s.Pcol[k]], acc);
  return acc;
}

// Gp = (R_x + P + A' R_y^-1 A) p  (private.c:108-121); returns this thread's share of p'Gp
__device__ __forceinline__ double mat_vec(B &s, const double *p, double *Gp) {
  BFOR(i, s.m) s.tmp[i] = rowdot_A(s, i, p) / s.dr[s.n + i];
  __syncthreads();
  double pgp = 0.0;
  BFOR(j, s.n) {
    const double gj = coldot_A(s, j, s.tmp) + rowdot_P(s, j, p) + s.dr[j] * p[j];
    Gp[j] = gj;
    pgp = fma(p[j], gj, pgp);
  }
  return pgp;
}

// pcg, private.c:135-219.  x: right-hand side in, solution out.  warm may be null.
__device__ int pcg(B &s, double *x, const double *warm, int max_its, double tol) {
  double nr[1] = {0.0};
  if (!warm) {
    BFOR(j, s.n) { const double r = x[j]; s.cr[j] = r; x[j] = 0.0; nr[0] = fmax(nr[0], fabs(r)); }
  } else {
    mat_vec(s, warm, s.cGp);
    BFOR(j, s.n) { const double r = x[j] - s.cGp[j]; s.cr[j] = r; x[j] = warm[j]; nr[0] = fmax(nr[0], fabs(r)); }
  }
  breduce<0, 1>(nr, s.red);
  if (nr[0] < fmax(tol, 1e-12)) return 0;
  double zr[1] = {0.0};
  BFOR(j, s.n) { const double z = s.cM[j] * s.cr[j]; s.cp[j] = z; zr[0] = fma(z, s.cr[j], zr[0]); }
  breduce<1, 0>(zr, s.red);
  double ztr = zr[0];
  int i = 0;
  while (i < max_its) {
    double pg[1];
    pg[0] = mat_vec(s, s.cp, s.cGp);
    breduce<1, 0>(pg, s.red);
    const double alpha = ztr / pg[0];
    double rv[2] = {0.0, 0.0};
    BFOR(j, s.n) {
      x[j] = fma(alpha, s.cp[j], x[j]);
      const double r = fma(-alpha, s.cGp[j], s.cr[j]);
      s.cr[j] = r;
      rv[0] = fma(s.cM[j] * r, r, rv[0]);
      rv[1] = fmax(rv[1], fabs(r));
    }
    breduce<1, 1>(rv, s.red);
    if (rv[1] < tol) return i + 1;
    if (ztr == 0.0) break;
    const double beta = rv[0] / ztr;
    ztr = rv[0];
    BFOR(j, s.n) s.cp[j] = fma(beta, s.cp[j], s.cM[j] * s.cr[j]);
    __syncthreads();
    ++i;
  }
  return i;
}

// Direct mode (problems whose dense n x n reduced matrix fits in shared memory): G = R_x + P + A' R_y^-1 A
// is formed column by column (thread j owns column j, fixed summation order) and inverted in place by
// Gauss-Jordan elimination without pivoting (G is symmetric positive definite).  Stored with an odd
// leading dimension so that both row- and column-wise sweeps are bank-conflict free.  Redone on every
// scale update, exactly where the reference refactorises (scs_update_lin_sys_diag_r, linsys.h:64).
__device__ void build_ginv(B &s) {
  const int n = s.n, ld = s.gld;
  double *G = s.Ginv;
  BFOR(k, n * ld) G[k] = 0.0;
  __syncthreads();
  BFOR(j, n) {
    double *col = G + (size_t)j * ld;
    for (int k = s.Acp[j]; k < s.Acp[j + 1]; ++k) {
      const int i = s.Arow[k];
      const double a = s.Aval[k] / s.dr[n + i];
      for (int kk = s.Arp[i]; kk < s.Arp[i + 1]; ++kk) col[s.Acol[kk]] = fma(a, s.AvalR[kk], col[s.Acol[kk]]);
    }
    for (int k = s.Prp[j]; k < s.Prp[j + 1]; ++k) col[s.Pcol[k]] += s.Pval[k];
    col[j] += s.dr[j];
  }
  __syncthreads();
  // Gauss-Jordan, one elimination step per k.  Thread (j, part) owns the rows [r0, r1) of column j (kBT / n
  // parts per column), so the rank-one update is a stride-1 walk down a column of the odd-ld array: no
  // integer division, no bank conflicts, two barriers per step.
  double *f = s.cGp;   // column k before the step
  double *rowk = s.cr;  // row k before the step, scaled by the pivot
  const int parts = kBT / n > 0 ? kBT / n : 1;
  const int rows_pp = (n + parts - 1) / parts;
  for (int k = 0; k < n; ++k) {
    const double p = 1.0 / G[k + (size_t)k * ld];
    BFOR(i, n) { f[i] = G[i + (size_t)k * ld]; rowk[i] = G[k + (size_t)i * ld] * p; }
    __syncthreads();  // column k and row k are saved (and the pivot read) before anything below overwrites them
    for (int t = s.tid; t < n * parts; t += kBT) {
      const int j = t % n, part = t / n;  // (one division per step, not per element)
      const int r0 = part * rows_pp, r1 = r0 + rows_pp < n ? r0 + rows_pp : n;
      double *col = G + (size_t)j * ld;
      if (j != k) {
        const double gkj = rowk[j];
        for (int i = r0; i < r1; ++i) col[i] = (i == k) ? gkj : fma(-f[i], gkj, col[i]);
      } else {
        for (int i = r0; i < r1; ++i) col[i] = (i == k) ? p : -f[i] * p;
      }
    }
    __syncthreads();
  }
}

// out_j = sum_k Ginv[k, j] rhs_k  (Ginv symmetric: column j read down the column).  `parts` (a power of two,
// <= kBT / n) adjacent lanes share one output: lane h of the group sums the terms k = h, h + parts, ... and the
// group adds its partial sums with a butterfly -- a fixed order, and kBT / n times shorter dependent chains.
__device__ __forceinline__ void ginv_apply(const B &s, const double *rhs, double *out, bool accumulate) {
  const int n = s.n, parts = s.gparts, sh = s.gshift;
  const int total = ((n << sh) + 31) & ~31;  // whole warps: the butterfly needs every lane
  for (int t = s.tid; t < total; t += kBT) {
    const int j = t >> sh, h = t & (parts - 1);
    const bool valid = j < n;
    const double *col = s.Ginv + (size_t)(valid ? j : 0) * s.gld;
    double a0 = 0.0, a1 = 0.0;
    int k = h;
    for (; k + parts < n; k += 2 * parts) {
      a0 = fma(col[k], rhs[k], a0);
      a1 = fma(col[k + parts], rhs[k + parts], a1);
    }
    if (k < n) a0 = fma(col[k], rhs[k], a0);
    double r = a0 + a1;
    for (int o = parts >> 1; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (valid && h == 0) out[j] = accumulate ? out[j] + r : r;
  }
}

// bv = [r_x; r_y] in, [x; y] out, through the resident inverse (+ `refine` steps of iterative refinement
// against the sparse operator)
__device__ void lin_solve_direct(B &s, double *bv) {
  const int n = s.n;
  BFOR(i, s.m) s.tmp[i] = bv[n + i] / s.dr[n + i];
  __syncthreads();
  BFOR(j, n) s.cr[j] = bv[j] + coldot_A(s, j, s.tmp);  // reduced right-hand side
  __syncthreads();
  ginv_apply(s, s.cr, s.cp, false);
  __syncthreads();
  for (int step = 0; step < s.refine; ++step) {
    mat_vec(s, s.cp, s.cGp);                 // cGp_j = (G x)_j, own j
    BFOR(j, n) s.cGp[j] = s.cr[j] - s.cGp[j];
    __syncthreads();
    ginv_apply(s, s.cGp, s.cp, true);
    __syncthreads();
  }
  BFOR(j, n) bv[j] = s.cp[j];
  BFOR(i, s.m) bv[n + i] = (-bv[n + i] + rowdot_A(s, i, s.cp)) / s.dr[n + i];
  __syncthreads();
}

// scs_solve_lin_sys, private.c:276-316: bv = [r_x; r_y] in, [x; y] out
__device__ void lin_solve(B &s, double *bv, const double *warm, double tol) {
  if (s.direct) { lin_solve_direct(s, bv); return; }
  double nb[1] = {0.0};
  BFOR(i, s.n + s.m) nb[0] = fmax(nb[0], fabs(bv[i]));
  breduce<0, 1>(nb, s.red);
  if (nb[0] <= 1e-12) {
    BFOR(i, s.n + s.m) bv[i] = 0.0;
    __syncthreads();
    return;
  }
  BFOR(i, s.m) s.tmp[i] = bv[s.n + i] / s.dr[s.n + i];
  __syncthreads();
  BFOR(j, s.n) bv[j] += coldot_A(s, j, s.tmp);
  __syncthreads();  // tmp is rewritten by the first mat_vec
  s.cg_its += pcg(s, bv, warm, 10 * s.n, tol);
  __syncthreads();
  BFOR(i, s.m) bv[s.n + i] = (-bv[s.n + i] + rowdot_A(s, i, bv)) / s.dr[s.n + i];  // reads bv[0..n) only
  __syncthreads();
}

// set_preconditioner, private.c:50-84
__device__ void set_precond(B &s) {
  BFOR(j, s.n) {
    double mj = s.dr[j];
    for (int k = s.Acp[j]; k < s.Acp[j + 1]; ++k) mj += s.Aval[k] * s.Aval[k] / s.dr[s.n + s.Arow[k]];
    for (int k = s.Prp[j]; k < s.Prp[j + 1]; ++k)
      if (s.Pcol[k] == j) mj += s.Pval[k];
    s.cM[j] = 1.0 / mj;
  }
  __syncthreads();
}

// set_diag_r, scs.c:929-938 + cones.c:349-363
__device__ void set_diag_r(B &s, double scale, double rho_x) {
  BFOR(i, s.l) {
    double r;
    if (i < s.n) r = rho_x;
    else if (i < s.n + s.z) r = 1.0 / (1000.0 * scale);
    else if (i < s.l - 1) r = 1.0 / scale;
    else r = kTauFactorB;
    s.dr[i] = r;
  }
  __syncthreads();
}

// update_work_cache, scs.c:1066-1076
__device__ void update_work_cache(B &s) {
  BFOR(i, s.n + s.m) s.g[i] = i < s.n ? s.c[i] : -s.b[i - s.n];
  __syncthreads();
  lin_solve(s, s.g, nullptr, kCgBestTolB);
}

// enforce_cone_boundaries (cones.c:366-379) on a length-m vector: second-order, exponential and power cones have size > 1 here
template <bool MEAN>
__device__ void enforce_soc(B &s, double *vec) {
  const int lane = s.tid & 31, w = s.tid >> 5;
  for (int cidx = w; cidx < s.nq; cidx += kBW) {
    const int off = s.qoff[cidx], len = s.qlen[cidx];
    double a = 0.0;
    for (int k = lane; k < len; k += 32) a = MEAN ? a + vec[off + k] : fmax(a, fabs(vec[off + k]));
    a = MEAN ? warp_sum(a) : warp_max(a);
    if (MEAN && len > 0) a /= (double)len;
    for (int k = lane; k < len; k += 32) vec[off + k] = a;
  }
}

// normalize_a_p, scs_matrix.c:407-470 (Ruiz passes :210-277, L2 pass :279-342, rescale :344-381)
__device__ void equilibrate(B &s) {
  for (int pass = 0; pass < kRuizB + kL2B; ++pass) {
    const bool l2 = pass >= kRuizB;
    double *Dt = s.tmp, *Et = s.cp;
    BFOR(i, s.m) {
      double a = 0.0;
      for (int k = s.Arp[i]; k < s.Arp[i + 1]; ++k) {
        const double v = s.Aval[s.Aperm[k]];
        a = l2 ? fma(v, v, a) : fmax(a, fabs(v));
      }
      Dt[i] = l2 ? sqrt(a) : a;
    }
    BFOR(j, s.n) {
      double a = 0.0;
      for (int k = s.Prp[j]; k < s.Prp[j + 1]; ++k) { const double v = s.Pval[k]; a = l2 ? fma(v, v, a) : fmax(a, fabs(v)); }
      for (int k = s.Acp[j]; k < s.Acp[j + 1]; ++k) { const double v = s.Aval[k]; a = l2 ? fma(v, v, a) : fmax(a, fabs(v)); }
      Et[j] = inv_sqrt_limited(l2 ? sqrt(a) : a);
    }
    __syncthreads();
    if (s.nq > 0) {
      if (l2) enforce_soc<true>(s, Dt); else enforce_soc<false>(s, Dt);
      __syncthreads();
    }
    BFOR(i, s.m) { const double d = inv_sqrt_limited(Dt[i]); Dt[i] = d; s.D[i] *= d; }
    __syncthreads();
    BFOR(j, s.n) {
      const double e = Et[j];
      for (int k = s.Acp[j]; k < s.Acp[j + 1]; ++k) s.Aval[k] *= Dt[s.Arow[k]] * e;
      for (int k = s.Prp[j]; k < s.Prp[j + 1]; ++k) s.Pval[k] *= e * Et[s.Pcol[k]];
      s.E[j] *= e;
    }
    __syncthreads();
  }
}

// populate_residual_struct + unnormalize_residuals + compute_residuals (scs.c:441-585)
__device__ void populate_residuals(B &s, Resid &r, int iter, double pscale, double dscale) {
  if (r.last_iter == iter) return;
  r.last_iter = iter;
  const double tau = fabs(s.u[s.l - 1]), kap = fabs(s.rsk[s.l - 1]);
  const double *x = s.u, *y = s.u + s.n, *sv = s.rsk + s.n;
  double v[12];  // sums: y'b, x'c, x'Px ; maxes: see below
#pragma unroll
  for (int k = 0; k < 12; ++k) v[k] = 0.0;
  BFOR(i, s.m) {
    const double ax = rowdot_A(s, i, x), si = sv[i];
    const double axs = ax + si, axsb = axs - tau * s.b[i];
    const double fD = (1.0 / dscale) / s.D[i];
    v[0] = fma(y[i], s.b[i], v[0]);
    v[3] = fmax(v[3], fabs(axsb));
    v[4] = fmax(v[4], fabs(axsb * fD));
    v[5] = fmax(v[5], fabs(axs * fD));
    v[6] = fmax(v[6], fabs(ax * fD));
    v[7] = fmax(v[7], fabs(si / (s.D[i] * dscale)));
  }
  BFOR(j, s.n) {
    const double px = rowdot_P(s, j, x), aty = coldot_A(s, j, y);
    const double pac = px + aty + tau * s.c[j];
    const double fE = (1.0 / pscale) / s.E[j];
    v[1] = fma(x[j], s.c[j], v[1]);
    v[2] = fma(px, x[j], v[2]);
    v[8] = fmax(v[8], fabs(pac));
    v[9] = fmax(v[9], fabs(pac * fE));
    v[10] = fmax(v[10], fabs(px * fE));
    v[11] = fmax(v[11], fabs(aty * fE));
  }
  breduce<3, 9>(v, s.red);
  const double pd = pscale * dscale;
  r.tau = tau;
  r.n_ax_s_btau = v[3]; r.n_px_aty_ctau = v[8];
  // normalised scalars (scs.c:556-574), then / pd (scs.c:476-486)
  const double bty_n = safediv_b(v[0], tau), ctx_n = safediv_b(v[1], tau), xpx_n = safediv_b(v[2], tau * tau);
  r.kap = kap / pd;
  r.bty_tau = v[0] / pd; r.ctx_tau = v[1] / pd; r.xpx_tau = v[2] / pd;
  r.bty = bty_n / pd; r.ctx = ctx_n / pd; r.xpx = xpx_n / pd;
  r.gap = fabs(xpx_n + ctx_n + bty_n) / pd;
  r.pobj = (xpx_n / 2.0 + ctx_n) / pd;
  r.dobj = (-xpx_n / 2.0 - bty_n) / pd;
  r.o_ax_s_btau = v[4]; r.o_ax = v[6]; r.o_s = v[7];
  r.o_px_aty_ctau = v[9]; r.o_px = v[10]; r.o_aty = v[11];
  // compute_residuals, scs.c:441-463
  const double tol = kInfeasNegTolB / pd;
  r.res_pri = safediv_b(v[4], tau);
  r.res_dual = safediv_b(v[9], tau);
  r.res_unbdd_a = r.res_unbdd_p = r.res_infeas = NAN;
  if (r.ctx_tau < -tol) {
    r.res_unbdd_a = safediv_b(v[5], -r.ctx_tau);
    r.res_unbdd_p = safediv_b(v[10], -r.ctx_tau);
  }
  if (r.bty_tau < -tol) r.res_infeas = safediv_b(v[11], -r.bty_tau);
}

// has_converged, scs.c:589-627 (a < b is false when either side is NaN, like isless)
__device__ int has_converged(const Resid &r, const BStg &g, double nm_b, double nm_c) {
  if (r.tau > 0.0) {
    const double grl = fmax(fmax(fabs(r.xpx), fabs(r.ctx)), fabs(r.bty));
    const double prl = fmax(fmax(nm_b * r.tau, r.o_s), r.o_ax) / r.tau;
    const double drl = fmax(fmax(nm_c * r.tau, r.o_px), r.o_aty) / r.tau;
    if (r.res_pri < g.eps_abs + g.eps_rel * prl && r.res_dual < g.eps_abs + g.eps_rel * drl &&
        r.gap < g.eps_abs + g.eps_rel * grl)
      return SCS_SOLVED;
  }
  if (r.res_unbdd_a < g.eps_infeas && r.res_unbdd_p < g.eps_infeas) return SCS_UNBOUNDED;
  if (r.res_infeas < g.eps_infeas) return SCS_INFEASIBLE;
  return 0;
}

// ---- Anderson acceleration (aa.c) --------------------------------------------------------
__device__ void aa_reset_b(B &s) {
  __syncthreads();
  if (s.tid == 0) aa_reset_dev(s.st, kBAaMax);
  __syncthreads();
}

// solve (aa.c:422-652) for the current history of `len` columns; f (= v) updated in place on success
__device__ double aa_solve_b(B &s, const AaParams &ap, double *f, int len) {
  AaState *st = s.st;
  double r = 0.0;
  if (ap.regularization > 0) {
    auto frob = [&](const double *nc) {
      double mx = 0.0;
      for (int i = 0; i < ap.mem; ++i) mx = fmax(mx, nc[i]);
      if (mx == 0.0) return 0.0;
      double ss = 0.0;
      for (int i = 0; i < ap.mem; ++i) { const double q = nc[i] / mx; ss += q * q; }
      return mx * sqrt(ss);
    };
    const double ny = frob(st->nrm_y_col);
    const double na = ap.type1 ? frob(st->nrm_s_col) : ny;
    r = ap.regularization * na * ny;
  } else if (ap.regularization < 0) {
    r = -ap.regularization;
  }
  const double sqrt_r = r > 0 ? sqrt(r) : 0.0;
  const int C = ap.type1 ? 2 * len + 1 : len + 1;
  const int rows = s.l + len, ld = s.ald, l = s.l;
  const double *Asrc = ap.type1 ? s.aS : s.aY;
  // stacked block [A | Y | g] over [sqrt(r) I | sqrt(r) I | 0]; every thread owns rows tid, tid + 128, ...
  BFOR(i, rows) {
    for (int cidx = 0; cidx < C; ++cidx) {
      double val;
      if (i < l) val = cidx < len ? Asrc[(size_t)cidx * l + i] : (cidx < C - 1 ? s.aY[(size_t)(cidx - len) * l + i] : s.ag[i]);
      else val = (cidx == i - l || (ap.type1 && cidx == len + i - l)) ? sqrt_r : 0.0;
      s.aW[(size_t)cidx * ld + i] = val;
    }
  }
  BFOR(k, len * C) s.aaR[k] = 0.0;
  __syncthreads();
  for (int j = 0; j < len; ++j) {
    double pv[kBC];
#pragma unroll
    for (int k = 0; k < kBC; ++k) pv[k] = 0.0;
    BFOR(i, rows) {
      const double bj = s.aW[(size_t)j * ld + i];
#pragma unroll
      for (int k = 0; k < kBC; ++k)
        if (k >= j && k < C) pv[k] = fma(bj, s.aW[(size_t)k * ld + i], pv[k]);
    }
    breduce<kBC, 0>(pv, s.red);
    double sj = 0.0;
#pragma unroll
    for (int k = 0; k < kBC; ++k) if (k == j) sj = pv[k];
    if (sj == 0.0) continue;  // uniform
    const double rjj = s.aaR[j * C + j];
    const double nrm = sqrt(rjj * rjj + sj);
    const double alpha = rjj >= 0.0 ? -nrm : nrm;
    const double v0 = rjj - alpha;
    const double beta = 2.0 / (v0 * v0 + sj);
#pragma unroll
    for (int k = 0; k < kBC; ++k) pv[k] = (k > j && k < C) ? beta * (v0 * s.aaR[j * C + k] + pv[k]) : 0.0;
    __syncthreads();  // every thread has read row j of R
    if (s.tid == 0) {
#pragma unroll
      for (int k = 0; k < kBC; ++k) if (k > j && k < C) s.aaR[j * C + k] -= pv[k] * v0;
      s.aaR[j * C + j] = alpha;
    }
    BFOR(i, rows) {
      const double bj = s.aW[(size_t)j * ld + i];
#pragma unroll
      for (int k = 0; k < kBC; ++k)
        if (k > j && k < C) s.aW[(size_t)k * ld + i] -= pv[k] * bj;
    }
  }
  __syncthreads();
  if (s.tid == 0) aa_small_solve(ap, s.aaR, len, C, r, s.aaScr);
  __syncthreads();
  const double aa_norm = st->aa_norm;
  if (st->success) {
    BFOR(i, l) {
      double acc = f[i];
      for (int k = 0; k < len; ++k) acc = fma(-s.aD[(size_t)k * l + i], st->gamma[k], acc);
      f[i] = acc;
    }
  }
  __syncthreads();
  return aa_norm;
}

// aa_apply, aa.c:822-854 (f = v is overwritten, x = v_prev)
__device__ double aa_apply_b(B &s, const AaParams &ap, double *f, const double *x) {
  AaState *st = s.st;
  const int it = st->iter, l = s.l;
  __syncthreads();
  if (it == 0) {
    BFOR(i, l) { s.ax[i] = x[i]; s.af[i] = f[i]; s.agp[i] = x[i] - f[i]; }
    if (s.tid == 0) { st->success = 0; st->aa_norm = 0.0; st->iter = 1; }
    __syncthreads();
    return 0.0;
  }
  const size_t col = (size_t)((it - 1) % ap.mem) * l;
  double v[3] = {0.0, 0.0, 0.0};
  BFOR(i, l) {  // update_accel_params, aa.c:340-390
    const double xi = x[i], fi = f[i];
    const double sd = xi - s.ax[i], dd = fi - s.af[i], gi = xi - fi, yd = gi - s.agp[i];
    s.aS[col + i] = sd; s.aD[col + i] = dd; s.aY[col + i] = yd; s.ag[i] = gi;
    s.ax[i] = xi; s.af[i] = fi; s.agp[i] = gi;
    v[0] = fma(sd, sd, v[0]); v[1] = fma(yd, yd, v[1]); v[2] = fma(gi, gi, v[2]);
  }
  breduce<3, 0>(v, s.red);
  if (s.tid == 0) {
    const int idx = (it - 1) % ap.mem;
    st->nrm_s_col[idx] = sqrt(v[0]);
    st->nrm_y_col[idx] = sqrt(v[1]);
    st->norm_g = sqrt(v[2]);
    st->success = 0;
    st->aa_norm = 0.0;
    if (it < ap.min_len) st->iter = it + 1;
  }
  __syncthreads();
  if (it < ap.min_len) return 0.0;
  return aa_solve_b(s, ap, f, it < ap.mem ? it : ap.mem);  // aa_small_solve advances st->iter
}

// aa_safeguard, aa.c:856-901.  f_new = v (after the step), x_new = v_prev
__device__ int aa_safeguard_b(B &s, const AaParams &ap, double *f_new, double *x_new) {
  AaState *st = s.st;
  const int ok = st->success;
  __syncthreads();
  if (!ok) return 0;
  double v[1] = {0.0};
  BFOR(i, s.l) { const double d = x_new[i] - f_new[i]; v[0] = fma(d, d, v[0]); }
  breduce<1, 0>(v, s.red);
  const bool reject = sqrt(v[0]) > ap.safeguard_factor * st->norm_g;
  __syncthreads();
  if (reject) {
    BFOR(i, s.l) { f_new[i] = s.af[i]; x_new[i] = s.ax[i]; }
    if (s.tid == 0) { st->n_safeguard_reject++; aa_reset_dev(st, kBAaMax); }
  } else if (s.tid == 0) {
    st->success = 0;
  }
  __syncthreads();
  return reject ? -1 : 0;
}

// second-order / PSD cones by one warp, exponential / power cones by one thread: cone_dev.cuh

// The box cone {(t, s): t bl <= s <= t bu} inside the Moreau step, by the whole CTA: uy holds s_saved (= 2 u_t - v on the
// cone's rows), ry the cone's R_y; x = -r s is projected by Newton on t with CTA-reduced gradient / Hessian (<= 25
// iterations, the reference's stopping rules), result uy = Pi(x) / r + s_saved.  Returns t (warm start of the next call).
__device__ double box_moreau(B &s, double *uy, const double *ry, double t) {
  const int bs = s.bsize;
  if (bs == 1) {
    if (s.tid == 0) { const double s0 = uy[0], r0 = ry[0]; uy[0] = fmax(-r0 * s0, 0.0) / r0 + s0; }
    return t;
  }
  const double r0 = ry[0], s0 = uy[0], tx0 = -r0 * s0, rho_t = 1.0 / r0;
  const int nb = bs - 1;
  for (int iter = 0; iter < 25; ++iter) {  // BOX_CONE_MAX_ITERS
    double v[2] = {0.0, 0.0};
    BFOR(j, nb) {
      const double rj = ry[1 + j], xj = -rj * uy[1 + j], rinv = 1.0 / rj, ub = s.bu[j], lb = s.bl[j];
      if (xj > t * ub) { v[0] += rinv * (t * ub - xj) * ub; v[1] += rinv * ub * ub; }
      else if (xj < t * lb) { v[0] += rinv * (t * lb - xj) * lb; v[1] += rinv * lb * lb; }
    }
    breduce<2, 0>(v, s.red);
    const double gt = rho_t * (t - tx0) + v[0], ht = rho_t + v[1], t_prev = t;
    t = fmax(t - gt / fmax(ht, 1e-8), 0.0);
    if (fabs(gt / fmax(ht, 1e-6)) < 1e-12 * fmax(t, 1.0) || fabs(t - t_prev) < 1e-11 * fmax(t, 1.0)) break;  // uniform
  }
  __syncthreads();  // every thread has read uy[0] (s0)
  BFOR(j, nb) {
    const double rj = ry[1 + j], sj = uy[1 + j];
    double xj = -rj * sj;
    if (xj > t * s.bu[j]) xj = t * s.bu[j];
    else if (xj < t * s.bl[j]) xj = t * s.bl[j];
    uy[1 + j] = xj / rj + sj;
  }
  if (s.tid == 0) uy[0] = t / r0 + s0;
  return t;
}

// kTri: the batch holds exponential / power cones (their root finders are kept out of the plain instantiation)
template <bool kTri>
__global__ void __launch_bounds__(kBT, kBCtasPerSm) k_batch_solve(const BArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int sh_pid;
  const BLay L(a.dims);
  double *sd = reinterpret_cast<double *>(smem_raw);
  B s;
  s.tid = threadIdx.x;
  s.Aval = sd + L.Aval; s.AvalR = sd + L.AvalR; s.Pval = sd + L.Pval; s.u = sd + L.u; s.ut = sd + L.ut; s.v = sd + L.v; s.vp = sd + L.vp;
  s.rsk = sd + L.rsk; s.g = sd + L.g; s.dr = sd + L.dr; s.b = sd + L.b; s.c = sd + L.c; s.D = sd + L.D; s.E = sd + L.E;
  s.cp = sd + L.cp; s.cr = sd + L.cr; s.cGp = sd + L.cGp; s.cM = sd + L.cM; s.tmp = sd + L.tmp; s.ws = sd + L.ws;
  s.red.buf = sd + L.red; s.red.phase = 0;
  s.aaR = sd + L.aaR; s.aaScr = sd + L.aaScr; s.Ginv = sd + L.Ginv; s.bl = sd + L.bl; s.bu = sd + L.bu; s.pw = sd + L.pw; s.psdw = sd + L.psdw; s.psd_ws = psd_ws_doubles(a.dims.ds);
  s.direct = a.dims.direct; s.refine = a.stg.refine;
  s.st = reinterpret_cast<AaState *>(smem_raw + (size_t)L.nd * 8);
  u16 *si = reinterpret_cast<u16 *>(smem_raw + (size_t)L.nd * 8 + L.st_bytes);
  s.Arow = si + L.Arow; s.Aperm = si + L.Aperm; s.Acol = si + L.Acol; s.Acp = si + L.Acp; s.Arp = si + L.Arp;
  s.Pcol = si + L.Pcol; s.Prp = si + L.Prp; s.qoff = si + L.qoff; s.qlen = si + L.qlen;
  const BStg &g = a.stg;
  const int lane = s.tid & 31, warp = s.tid >> 5;

  for (;;) {
    __syncthreads();
    if (s.tid == 0) sh_pid = atomicAdd(a.counter, 1);
    __syncthreads();
    const int pid = sh_pid;
    if (pid >= a.count) break;
    const unsigned long long t_begin = gtimer();
    const BProb pb = a.probs[pid];
    s.n = pb.n; s.m = pb.m; s.l = pb.n + pb.m + 1; s.nnzA = pb.nnzA; s.nnzP = pb.nnzP;
    s.z = pb.z; s.nl = pb.l; s.nq = pb.nq; s.bsize = pb.bsize; s.nsoc = pb.nsoc; s.ns = pb.ns; s.ncs = pb.ncs; s.ep = pb.ep; s.ed = pb.ed; s.np = pb.np; s.cg_its = 0; s.gld = pb.n | 1;
    s.gshift = 0;
    while (s.gshift < 5 && (pb.n << (s.gshift + 1)) <= kBT) ++s.gshift;  // lanes per output of ginv_apply
    s.gparts = 1 << s.gshift;
    const int n = s.n, m = s.m, l = s.l;
    const int mem = g.aa_mem < l ? g.aa_mem : l;  // aa_init, aa.c:657-700
    const bool aa_on = mem > 0;
    {  // AA workspace of this CTA slot
      double *w = a.aaws + (size_t)blockIdx.x * a.aaws_stride;
      s.aS = w; w += (size_t)l * mem; s.aY = w; w += (size_t)l * mem; s.aD = w; w += (size_t)l * mem;
      s.ax = w; w += l; s.af = w; w += l; s.ag = w; w += l; s.agp = w; w += l;
      s.aW = w; s.ald = l + mem;
    }
    AaParams ap;
    ap.dim = l; ap.mem = mem; ap.min_len = mem; ap.type1 = g.aa_type1; ap.ir_max_steps = 5;
    ap.regularization = g.aa_reg; ap.relaxation = 1.0; ap.safeguard_factor = 1.0; ap.max_weight_norm = 1e10;
    ap.x = ap.f = ap.g = ap.g_prev = ap.Y = ap.S = ap.D = ap.x_work = ap.Rpart = nullptr;
    ap.st = s.st;

    // ---- load
    const double *dp = a.dpool + pb.d_off;
    const u16 *ip = a.ipool + pb.i_off;
    BFOR(k, s.nnzA) { s.Aval[k] = dp[k]; s.Arow[k] = ip[k]; s.Aperm[k] = ip[s.nnzA + k]; s.Acol[k] = ip[2 * s.nnzA + k]; }
    BFOR(k, s.nnzP) { s.Pval[k] = dp[s.nnzA + k]; s.Pcol[k] = ip[3 * s.nnzA + (n + 1) + (m + 1) + k]; }
    BFOR(k, n + 1) { s.Acp[k] = ip[3 * s.nnzA + k]; s.Prp[k] = ip[3 * s.nnzA + (n + 1) + (m + 1) + s.nnzP + k]; }
    BFOR(k, m + 1) s.Arp[k] = ip[3 * s.nnzA + (n + 1) + k];
    BFOR(k, s.nq) {
      s.qoff[k] = ip[3 * s.nnzA + (n + 1) + (m + 1) + s.nnzP + (n + 1) + k];
      s.qlen[k] = ip[3 * s.nnzA + (n + 1) + (m + 1) + s.nnzP + (n + 1) + s.nq + k];
    }
    double nbc[2] = {0.0, 0.0};
    BFOR(i, m) { const double bi = dp[s.nnzA + s.nnzP + i]; s.b[i] = bi; s.D[i] = 1.0; nbc[0] = fmax(nbc[0], fabs(bi)); }
    BFOR(j, n) { const double cj = dp[s.nnzA + s.nnzP + m + j]; s.c[j] = cj; s.E[j] = 1.0; nbc[1] = fmax(nbc[1], fabs(cj)); }
    const int nbox = s.bsize > 1 ? s.bsize - 1 : 0;
    BFOR(j, nbox) { s.bl[j] = dp[s.nnzA + s.nnzP + m + n + j]; s.bu[j] = dp[s.nnzA + s.nnzP + m + n + nbox + j]; }
    BFOR(j, s.np) s.pw[j] = dp[s.nnzA + s.nnzP + m + n + 2 * nbox + j];
    double box_t = 1.0;  // box_t_warm_start, cones.c:1552
    if (s.tid == 0) {
      AaState *st = s.st;
      memset(st, 0, sizeof(AaState));
      st->last_aa_norm = NAN;
    }
    breduce<0, 2>(nbc, s.red);
    const double nm_b_orig = nbc[0], nm_c_orig = nbc[1];

    // ---- scs_init: equilibrate, scale b and c (normalize.c:33-61)
    double pscale = 1.0, dscale = 1.0;
    long long ck[6] = {0, 0, 0, 0, 0, 0};
    const long long ck_begin = clock64();
    if (g.normalize) {
      { const long long c0 = clock64(); equilibrate(s); ck[0] += clock64() - c0; }
    }
    BFOR(k, s.nnzA) s.AvalR[k] = s.Aval[s.Aperm[k]];  // final values in CSR order (one indirection less per non-zero)
    if (g.normalize && nbox > 0) {  // normalize_box_cone, cones.c:1153-1169: bounds follow the row scaling D
      const double *Db = s.D + s.z + s.nl;
      BFOR(j, nbox) {
        const double f = Db[j + 1] / Db[0];
        s.bu[j] = s.bu[j] >= 1e15 ? INFINITY : s.bu[j] * f;
        s.bl[j] = s.bl[j] <= -1e15 ? -INFINITY : s.bl[j] * f;
      }
    }
    __syncthreads();
    if (g.normalize) {
      double mx[1] = {0.0};
      BFOR(i, m) { const double bi = s.b[i] * s.D[i]; s.b[i] = bi; mx[0] = fmax(mx[0], fabs(bi)); }
      BFOR(j, n) { const double cj = s.c[j] * s.E[j]; s.c[j] = cj; mx[0] = fmax(mx[0], fabs(cj)); }
      breduce<0, 1>(mx, s.red);
      double sigma = mx[0];
      sigma = sigma < kMinNormB ? 1.0 : sigma;
      sigma = sigma > kMaxNormB ? kMaxNormB : sigma;
      sigma = safediv_b(1.0, sigma);
      BFOR(i, m) s.b[i] *= sigma;
      BFOR(j, n) s.c[j] *= sigma;
      pscale = dscale = sigma;
    }
    double scale = g.scale;
    set_diag_r(s, scale, g.rho_x);
    { const long long c0 = clock64(); if (s.direct) build_ginv(s); else set_precond(s); ck[1] += clock64() - c0; }
    BFOR(i, l) { s.u[i] = 0.0; s.ut[i] = 0.0; s.rsk[i] = 0.0; s.v[i] = i == l - 1 ? 1.0 : 0.0; }  // cold start, scs.c:629-636
    __syncthreads();
    const unsigned long long t_setup = gtimer();
    // ---- scs_solve
    update_work_cache(s);
    Resid r;
    r.last_iter = -1;
    r.n_ax_s_btau = 0.0; r.n_px_aty_ctau = 0.0;
    int last_scale_update_iter = 0, n_log_scale = 0, scale_updates = 0, rej = 0, acc = 0;
    double sum_log_scale = 0.0, aa_norm = 0.0;
    int status = 0, it = 0;
    for (it = 0; it < g.max_iters; ++it) {
      if (aa_on && it > 0 && it % g.aa_interval == 0) { const long long c0 = clock64(); aa_norm = aa_apply_b(s, ap, s.v, s.vp); ck[3] += clock64() - c0; }
      if (it >= kFeasIters) {  // normalize_v, scs.c:771-779
        double vn[1] = {0.0};
        BFOR(i, l) vn[0] = fma(s.v[i], s.v[i], vn[0]);
        breduce<1, 0>(vn, s.red);
        const double nrm = sqrt(vn[0]);
        if (nrm != 0.0) {  // SCS(scale_array)(v, sqrt(l) * ITERATE_NORM / ||v||)
          const double sc = sqrt((double)l) * 1.0 / nrm;
          BFOR(i, l) s.v[i] *= sc;
        }
      }
      // project_lin_sys, scs.c:691-729
      double nws[1] = {0.0};
      BFOR(i, l) {
        const double vi = s.v[i];
        if (aa_on) s.vp[i] = vi;
        s.ut[i] = i < n ? vi * s.dr[i] : (i < l - 1 ? -vi * s.dr[i] : vi);
      }
      double tol = -1.0;  // direct mode: exact solve, no warm start (scs.c:694)
      if (!s.direct) {
        const double u_tau = s.u[l - 1];
        BFOR(j, n) { const double w = s.u[j] + u_tau * s.g[j]; s.ws[j] = w; nws[0] = fmax(nws[0], fabs(w)); }
        breduce<0, 1>(nws, s.red);
        tol = fmin(r.n_ax_s_btau, r.n_px_aty_ctau);
        tol = fmax(kCgBestTolB, kCgTolFactorB * fmin(tol, nws[0] / pow((double)it + 1.0, kCgRateB)));
      } else {
        __syncthreads();
      }
      { const long long c0 = clock64(); lin_solve(s, s.ut, s.ws, tol); ck[2] += clock64() - c0; }
      double tau = 1.0;
      if (it >= kFeasIters) {  // root_plus, scs.c:667-688
        double q[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        BFOR(i, l - 1) {
          const double gi = s.g[i], ri = s.dr[i], pi = s.ut[i], mi = s.v[i];
          q[0] = fma(gi * gi, ri, q[0]);
          q[1] = fma(mi * gi, ri, q[1]);
          q[2] = fma(pi * gi, ri, q[2]);
          q[3] = fma(pi * pi, ri, q[3]);
          q[4] = fma(pi * mi, ri, q[4]);
        }
        breduce<5, 0>(q, s.red);
        const double ts = s.dr[l - 1], eta = s.v[l - 1];
        const double qa = ts + q[0], qb = q[1] - 2.0 * q[2] - eta * ts, qc = q[3] - q[4];
        const double rad = qb * qb - 4.0 * qa * qc;
        tau = (-qb + sqrt(fmax(rad, 0.0))) / (2.0 * qa);
      }
      __syncthreads();  // all threads have read v[l-1], dr[l-1] before the tau entry is rewritten
      // u_t -= tau g ; u = 2 u_t - v ; zero / nonneg rows projected inline (scs.c:727,754-768)
      BFOR(i, l) {
        double uti;
        if (i < l - 1) uti = s.ut[i] - tau * s.g[i];
        else uti = tau;
        s.ut[i] = uti;
        const double s0 = 2.0 * uti - s.v[i];
        double ui = s0;
        if (i >= n + s.z && i < n + s.z + s.nl) { const double ry = s.dr[i]; ui = fmax(-ry * s0, 0.0) / ry + s0; }
        else if (i == l - 1) ui = it < kFeasIters ? 1.0 : fmax(s0, 0.0);
        s.u[i] = ui;
      }
      if (s.bsize > 0) {  // box cone rows, Moreau wrapper cones.c:1562-1585 around proj_box_cone (cones.c:1174-1237)
        __syncthreads();
        box_t = box_moreau(s, s.u + n + s.z + s.nl, s.dr + n + s.z + s.nl, box_t);
      }
      if (s.nq > 0) {
        __syncthreads();
        for (int cidx = warp; cidx < s.nsoc; cidx += kBW)
          soc_moreau_warp(s.u + n + s.qoff[cidx], s.dr + n + s.qoff[cidx], s.qlen[cidx], lane);
        if (kTri) {  // PSD cones: one warp each; exponential / power cones: one thread each (x = -R s, project, x / r + s)
          if (warp < a.dims.pslots)
            for (int cidx = s.nsoc + warp; cidx < s.nsoc + s.ns + s.ncs; cidx += a.dims.pslots)
              psd_moreau_warp(s.u + n + s.qoff[cidx], s.dr + n + s.qoff[cidx], s.qlen[cidx], cidx >= s.nsoc + s.ns,
                              s.psdw + warp * s.psd_ws, lane);
          for (int cidx = s.nsoc + s.ns + s.ncs + s.tid; cidx < s.nq; cidx += kBT) {
            const int t3 = cidx - s.nsoc - s.ns - s.ncs;
            double *uy = s.u + n + s.qoff[cidx];
            const double *ry = s.dr + n + s.qoff[cidx];
            const double s0 = uy[0], s1 = uy[1], s2 = uy[2];
            double v3[3] = {-ry[0] * s0, -ry[1] * s1, -ry[2] * s2};
            const int kind = t3 < s.ep ? 1 : (t3 < s.ep + s.ed ? 0 : -1);
            proj_cone3(v3, kind, kind < 0 ? s.pw[t3 - s.ep - s.ed] : 0.0);
            uy[0] = v3[0] / ry[0] + s0; uy[1] = v3[1] / ry[1] + s1; uy[2] = v3[2] / ry[2] + s2;
          }
        }
      }
      __syncthreads();
      BFOR(i, l) s.rsk[i] = (s.v[i] + s.u[i] - 2.0 * s.ut[i]) * s.dr[i];  // compute_rsk, scs.c:739-744
      __syncthreads();
      if (it % kConvInterval == 0) {
        { const long long c0 = clock64(); populate_residuals(s, r, it, pscale, dscale); ck[4] += clock64() - c0; }
        status = has_converged(r, g, nm_b_orig, nm_c_orig);
        if (status != 0) break;
      }
      if (g.adaptive_scale && it == r.last_iter) {  // update_scale, scs.c:1112-1189
        const double denom_pri = fmax(fmax(r.o_ax, r.o_s), nm_b_orig * r.tau);
        const double denom_dual = fmax(fmax(r.o_px, r.o_aty), nm_c_orig * r.tau);
        const double rel_pri = fmax(safediv_b(r.o_ax_s_btau, denom_pri), kDivEpsB);
        const double rel_dual = fmax(safediv_b(r.o_px_aty_ctau, denom_dual), kDivEpsB);
        sum_log_scale += log(rel_pri) - log(rel_dual);
        n_log_scale += 1;
        const double factor = sqrt(exp(sum_log_scale / (double)n_log_scale));
        if (it - last_scale_update_iter >= kRescaleMinIters) {
          const double new_scale = fmin(fmax(scale * factor, kMinScaleB), kMaxScaleB);
          if (new_scale != scale && (factor > sqrt(10.0) || factor < 1.0 / sqrt(10.0))) {
            scale_updates++;
            sum_log_scale = 0.0; n_log_scale = 0; last_scale_update_iter = it;
            scale = new_scale;
            set_diag_r(s, scale, g.rho_x);
            { const long long c0 = clock64(); if (s.direct) build_ginv(s); else set_precond(s); ck[1] += clock64() - c0; }
            update_work_cache(s);
            if (aa_on) aa_reset_b(s);
            BFOR(i, l) s.v[i] = s.rsk[i] / s.dr[i] + 2.0 * s.ut[i] - s.u[i];
            __syncthreads();
          }
        }
      }
      BFOR(i, l) s.v[i] += g.alpha * (s.u[i] - s.ut[i]);  // update_dual_vars, scs.c:746-751
      __syncthreads();
      if (aa_on && it % g.aa_interval == 0 && aa_norm > 0) {  // scs.c:1386-1394
        const long long c0 = clock64();
        if (aa_safeguard_b(s, ap, s.v, s.vp) < 0) rej++; else acc++;
        ck[3] += clock64() - c0;
      }
    }
    // ---- finalize, scs.c:874-924
    populate_residuals(s, r, it, pscale, dscale);
    double *sol = a.sol + pb.sol_off;
    double fin[3] = {0.0, 0.0, 0.0};  // s'y ; ||s||_inf, ||y||_inf of the un-normalised, un-scaled solution
    BFOR(i, m) {
      const double yi = s.u[n + i] * (s.D[i] / pscale), si_ = s.rsk[n + i] / (s.D[i] * dscale);
      fin[0] = fma(si_, yi, fin[0]);
      fin[1] = fmax(fin[1], fabs(si_));
      fin[2] = fmax(fin[2], fabs(yi));
    }
    breduce<1, 2>(fin, s.red);
    double fx = 1.0, fy = 1.0, fs = 1.0;
    double o_gap = NAN, o_pri = NAN, o_dual = NAN, o_pobj = NAN, o_dobj = NAN;
    int status_val = status;
    auto set_solved = [&]() {
      fx = fy = fs = safediv_b(1.0, r.tau);
      o_gap = r.gap; o_pri = r.res_pri; o_dual = r.res_dual;
      o_pobj = r.xpx / 2.0 + r.ctx; o_dobj = -r.xpx / 2.0 - r.bty;
      status_val = SCS_SOLVED;
    };
    auto set_infeasible = [&]() {
      fy = -1.0 / r.bty_tau; fx = NAN; fs = NAN;
      o_pobj = INFINITY; o_dobj = INFINITY;
      status_val = SCS_INFEASIBLE;
    };
    auto set_unbounded = [&]() {
      fx = -1.0 / r.ctx_tau; fs = -1.0 / r.ctx_tau; fy = NAN;
      o_pobj = -INFINITY; o_dobj = -INFINITY;
      status_val = SCS_UNBOUNDED;
    };
    if (status == SCS_SOLVED) set_solved();
    else if (status == SCS_INFEASIBLE) set_infeasible();
    else if (status == SCS_UNBOUNDED) set_unbounded();
    else {  // set_unfinished, scs.c:845-871
      if (r.kap > r.tau && (r.bty_tau < 0 || r.ctx_tau < 0)) {
        if (r.bty_tau < 0 && r.bty_tau < r.ctx_tau) { set_infeasible(); status_val = SCS_INFEASIBLE_INACCURATE; }
        else { set_unbounded(); status_val = SCS_UNBOUNDED_INACCURATE; }
      } else if (r.tau > 0) {
        set_solved();
        status_val = SCS_SOLVED_INACCURATE;
      } else {
        status_val = SCS_FAILED;
      }
    }
    BFOR(j, n) sol[j] = s.u[j] * (s.E[j] / dscale) * fx;
    BFOR(i, m) {
      sol[n + i] = s.u[n + i] * (s.D[i] / pscale) * fy;
      sol[n + m + i] = s.rsk[n + i] / (s.D[i] * dscale) * fs;
    }
    if (s.tid == 0) {
      const unsigned long long t_end = gtimer();
      const AaState *st = s.st;
      BOut o;
      o.iter = it; o.status_val = status_val; o.scale_updates = scale_updates; o.rej = rej; o.acc = acc;
      o.cg_its = s.cg_its;
      o.n_accept = st->n_accept; o.n_reject_rank0 = st->n_reject_rank0; o.n_reject_nonfinite = st->n_reject_nonfinite;
      o.n_reject_weight_cap = st->n_reject_weight_cap; o.n_safeguard_reject = st->n_safeguard_reject;
      o.last_rank = st->last_rank; o.aa_iter = st->iter; o.pad = 0;
      o.pobj = o_pobj; o.dobj = o_dobj; o.res_pri = o_pri; o.res_dual = o_dual; o.gap = o_gap;
      o.res_infeas = r.res_infeas; o.res_unbdd_a = r.res_unbdd_a; o.res_unbdd_p = r.res_unbdd_p;
      o.scale = scale; o.comp_slack = fabs(fin[0]); o.nm_s = fin[1]; o.nm_y = fin[2];
      o.last_aa_norm = st->last_aa_norm; o.last_reg = st->last_regularization;
      o.setup_ms = (double)(t_setup - t_begin) * 1e-6;
      o.solve_ms = (double)(t_end - t_setup) * 1e-6;
      ck[5] = clock64() - ck_begin;
      for (int q = 0; q < 6; ++q) o.clk[q] = ck[q];
      a.out[pid] = o;
    }
  }
}

// ------------------------------------------------------------------------------- host ----
struct Eligibility { bool ok; int nnzP_full, nq, n3, ds; };

static Eligibility fused_eligible(const ScsData *d, const ScsCone *k, const ScsSettings *stgs) {
  Eligibility e{false, 0, 0, 0, 0};
  if (stgs->warm_start || stgs->time_limit_secs > 0) return e;
  if (stgs->acceleration_lookback > kBAaMax) return e;
  if (stgs->acceleration_lookback > 0 && stgs->acceleration_relaxation != 1.0) return e;
  long long ncones = 0;
  for (int i = 0; i < k->cssize; ++i) {  // complex cones go through the real embedding of order 2 cs
    if (k->cs[i] < 1 || 2 * k->cs[i] > kBPsdMax) return e;
    e.ds = std::max(e.ds, 2 * (int)k->cs[i]);
  }
  for (int i = 0; i < k->ssize; ++i) {
    if (k->s[i] < 1 || k->s[i] > kBPsdMax) return e;
    e.ds = std::max(e.ds, (int)k->s[i]);
  }
  if (k->psize > 0 && !k->p) return e;
  if (k->bsize > 1 && (!k->bl || !k->bu)) return e;
  const long long nnzA = d->A->p[d->n];
  long long nnzP = 0;
  if (d->P) {
    for (int j = 0; j < d->n; ++j)
      for (int q = d->P->p[j]; q < d->P->p[j + 1]; ++q) nnzP += d->P->i[q] == j ? 1 : 2;
  }
  if (d->n >= 65535 || d->m >= 65535 || nnzA >= 65535 || nnzP >= 65535) return e;
  e.nnzP_full = (int)nnzP;
  e.n3 = (int)(k->ep + k->ed + k->psize);
  ncones = (long long)k->qsize + k->ssize + k->cssize + e.n3;
  if (ncones >= 65535) return e;
  e.nq = (int)ncones;
  e.ok = true;
  return e;
}

// Host plan of a batch (no device work): which members the one-CTA kernel takes, and the shared-memory carve-up
// (maxima over those members) they share.  A member that fails validation is marked nq = -1.
static void batch_classify(int count, const ScsData *const *d, const ScsCone *const *k, const ScsSettings *stgs,
                           std::vector<Eligibility> &elig, std::vector<int> &fused, BDims &dims) {
  elig.assign((size_t)count, Eligibility{false, 0, 0, 0, 0});
  fused.clear();
  dims = BDims{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < count; ++i) {
    if (!d[i] || !k[i] || validate_problem(d[i], k[i], stgs) < 0) {
      elig[i].nq = -1;
      continue;
    }
    elig[i] = fused_eligible(d[i], k[i], stgs);
    if (!elig[i].ok) continue;
    BDims t = dims;
    t.n = std::max(t.n, (int)d[i]->n); t.m = std::max(t.m, (int)d[i]->m);
    t.nnzA = std::max(t.nnzA, (int)d[i]->A->p[d[i]->n]); t.nnzP = std::max(t.nnzP, elig[i].nnzP_full);
    t.nq = std::max(t.nq, elig[i].nq);
    t.nb = std::max(t.nb, k[i]->bsize > 1 ? (int)k[i]->bsize - 1 : 0);
    t.np = std::max(t.np, (int)k[i]->psize);
    t.tri = t.tri || elig[i].n3 > 0 || elig[i].ds > 0;
    t.ds = std::max(t.ds, elig[i].ds);
    t.mem = std::min((int)stgs->acceleration_lookback, kBAaMax);
    t.pslots = t.ds > 0 ? kBW : 0;  // as many PSD workspaces (one per warp) as the footprint allows
    while (t.pslots > 1 && BLay(t).bytes() > 200 * 1024) t.pslots >>= 1;
    if (BLay(t).bytes() > 200 * 1024) { elig[i].ok = false; continue; }  // would not fit next to the others
    dims = t;
    fused.push_back(i);
  }
}
// linear-system mode of the batch kernel: resident dense inverse when it fits, PCG otherwise
static void batch_pick_linsys(BDims &dims, bool any) {
  const char *e = getenv("SCS_B200_BATCH_DIRECT");
  BDims t = dims;
  t.direct = 1;
  if (!(e && atoi(e) == 0) && any && BLay(t).bytes() <= 227 * 1024) dims.direct = 1;
}

static void status_string(int status_val, int iter, int max_iters, char *out) {
  const char *base = "failure";
  switch (status_val) {
    case SCS_SOLVED: case SCS_SOLVED_INACCURATE: base = "solved"; break;
    case SCS_INFEASIBLE: case SCS_INFEASIBLE_INACCURATE: base = "infeasible"; break;
    case SCS_UNBOUNDED: case SCS_UNBOUNDED_INACCURATE: base = "unbounded"; break;
    default: break;
  }
  strcpy(out, base);
  if (status_val == SCS_SOLVED_INACCURATE || status_val == SCS_INFEASIBLE_INACCURATE ||
      status_val == SCS_UNBOUNDED_INACCURATE || status_val == SCS_FAILED) {
    if (iter >= max_iters) strcat(out, " (inaccurate - reached max_iters)");
  }
}

// Pinned host staging, device pools, stream and events of the batch engine are kept per host thread and
// only ever grow: a pinned allocation costs milliseconds per megabyte, which dominated the wall time of a
// 1024-problem batch (profiles/r1f_configs_b200.jsonl: 228 ms of "packing" next to 121 ms of kernel).
// They live until the process exits (thread_local objects with CUDA destructors would run after the driver
// has shut down).
struct BatchArena {
  int dev = -1;
  void *h[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t hcap[4] = {0, 0, 0, 0};
  void *d[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t dcap[7] = {0, 0, 0, 0, 0, 0, 0};
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  void release() {
    for (int k = 0; k < 4; ++k) { if (h[k]) cudaFreeHost(h[k]); h[k] = nullptr; hcap[k] = 0; }
    for (int k = 0; k < 7; ++k) { if (d[k]) cudaFree(d[k]); d[k] = nullptr; dcap[k] = 0; }
    if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    e0 = e1 = nullptr; st = nullptr;
  }
  cudaError_t host(int k, size_t bytes, void **out) {
    if (hcap[k] < bytes) {
      if (h[k]) cudaFreeHost(h[k]);
      h[k] = nullptr; hcap[k] = 0;
      const size_t cap = bytes + bytes / 4 + 256;
      const cudaError_t e = cudaMallocHost(&h[k], cap);
      if (e != cudaSuccess) return e;
      hcap[k] = cap;
    }
    *out = h[k];
    return cudaSuccess;
  }
  cudaError_t device(int k, size_t bytes, void **out) {
    if (dcap[k] < bytes) {
      if (d[k]) cudaFree(d[k]);
      d[k] = nullptr; dcap[k] = 0;
      const size_t cap = bytes + bytes / 4 + 256;
      const cudaError_t e = cudaMalloc(&d[k], cap);
      if (e != cudaSuccess) return e;
      dcap[k] = cap;
    }
    *out = d[k];
    return cudaSuccess;
  }
};
static thread_local BatchArena g_arena;

struct BatchStats { long long fused = 0, streamed = 0, launches = 0, h2d = 0, d2h = 0, cg_its = 0, iters = 0, clk[6] = {0, 0, 0, 0, 0, 0}; double kernel_ms = 0, pack_ms = 0; int ctas = 0, smem = 0, direct = 0; };
static BatchStats g_last_batch;

}  // namespace
}  // namespace b200

using namespace b200;

// Host-only: the plan scs_b200_solve_batch would make for this batch.  fused_out[count]: 1 = one-CTA kernel, 0 =
// streaming engine, -1 = fails validation; dims_out[8] = {shared-memory bytes per CTA, resident dense inverse (0/1),
// kernel instantiation with exp / power / PSD cones (0/1), largest PSD order (complex: 2 cs), PSD workspaces,
// cones with a boundary, power cones, box bounds}.  No device is touched (tests of the host logic).
extern "C" scs_int scs_b200_batch_plan(scs_int count, const ScsData *const *d, const ScsCone *const *k,
                                       const ScsSettings *stgs, scs_int *fused_out, scs_int *dims_out) {
  if (count < 0 || !d || !k || !stgs || !fused_out || !dims_out) return -1;
  std::vector<int> fused;
  std::vector<Eligibility> elig;
  BDims dims;
  batch_classify((int)count, d, k, stgs, elig, fused, dims);
  batch_pick_linsys(dims, !fused.empty());
  for (int i = 0; i < count; ++i) fused_out[i] = elig[i].nq == -1 ? -1 : (elig[i].ok ? 1 : 0);
  dims_out[0] = fused.empty() ? 0 : (scs_int)BLay(dims).bytes(); dims_out[1] = dims.direct; dims_out[2] = dims.tri;
  dims_out[3] = dims.ds; dims_out[4] = dims.pslots; dims_out[5] = dims.nq; dims_out[6] = dims.np; dims_out[7] = dims.nb;
  return (scs_int)fused.size();
}

extern "C" scs_int scs_b200_solve_batch(scs_int count, const ScsData *const *d, const ScsCone *const *k,
                                        const ScsSettings *stgs, ScsSolution *const *sol, ScsInfo *info,
                                        scs_int streams) {
  (void)streams;
  if (count < 0 || !d || !k || !stgs || !sol || !info) return -1;
  typedef std::chrono::steady_clock Clock;
  const auto t0 = Clock::now();
  g_last_batch = BatchStats();
  scs_int worst = 0;
  // ---- classify
  std::vector<int> fused;
  std::vector<Eligibility> elig;
  BDims dims;
  batch_classify((int)count, d, k, stgs, elig, fused, dims);
  for (int i = 0; i < count; ++i) {
    if (elig[i].nq == -1 || !sol[i]) {  // failed validation
      populate_on_failure(d[i] ? d[i]->m : -1, d[i] ? d[i]->n : -1, sol[i], &info[i], SCS_FAILED, "failure");
      worst = SCS_FAILED;
      if (elig[i].ok) fused.erase(std::remove(fused.begin(), fused.end(), i), fused.end());
      elig[i].ok = false;
      elig[i].nq = -1;
    }
  }
  const int dev = current_device();
  if (cudaSetDevice(dev) != cudaSuccess) return -1;
  batch_pick_linsys(dims, !fused.empty());
  // ---- fused path
  if (!fused.empty()) {
    const int nf = (int)fused.size();
    std::vector<BProb> probs((size_t)nf);
    long long dtot = 0, itot = 0, stot = 0;
    for (int f = 0; f < nf; ++f) {
      const ScsData *dd = d[fused[f]];
      const ScsCone *kk = k[fused[f]];
      BProb &p = probs[f];
      p.n = dd->n; p.m = dd->m; p.nnzA = dd->A->p[dd->n]; p.nnzP = elig[fused[f]].nnzP_full;
      p.z = kk->z; p.l = kk->l; p.bsize = kk->bsize;
      p.nsoc = kk->qsize; p.ns = kk->ssize; p.ncs = kk->cssize; p.ep = kk->ep; p.ed = kk->ed; p.np = kk->psize;
      p.nq = p.nsoc + p.ns + p.ncs + p.ep + p.ed + p.np;
      p.d_off = dtot; p.i_off = itot; p.sol_off = stot;
      dtot += dpool_count(p.n, p.m, p.nnzA, p.nnzP, p.bsize > 1 ? p.bsize - 1 : 0, p.np);
      itot += ipool_count(p.n, p.m, p.nnzA, p.nnzP, p.nq);
      itot = (itot + 7) & ~7ll;
      stot += p.n + 2ll * p.m;
    }
    BatchArena &ar = g_arena;
    if (ar.dev != dev) { ar.release(); ar.dev = dev; }
    double *h_d = nullptr; u16 *h_i = nullptr; double *h_sol = nullptr; BOut *h_out = nullptr;
    if (ar.host(0, sizeof(double) * (size_t)std::max(1ll, dtot), (void **)&h_d) != cudaSuccess ||
        ar.host(1, sizeof(u16) * (size_t)std::max(1ll, itot), (void **)&h_i) != cudaSuccess ||
        ar.host(2, sizeof(double) * (size_t)std::max(1ll, stot), (void **)&h_sol) != cudaSuccess ||
        ar.host(3, sizeof(BOut) * (size_t)nf, (void **)&h_out) != cudaSuccess) {
      fprintf(stderr, "libscsb200: batch: pinned allocation failed\n");
      return -1;
    }
    // pack: values in CSC order; CSR view by counting sort (ascending columns inside a row)
    std::vector<int> rowcnt, rowpos, colfill;
    for (int f = 0; f < nf; ++f) {
      const ScsData *dd = d[fused[f]];
      const ScsCone *kk = k[fused[f]];
      const BProb &p = probs[f];
      const ScsMatrix *A = dd->A, *P = dd->P;
      double *pd = h_d + p.d_off;
      u16 *pi = h_i + p.i_off;
      u16 *Arow = pi, *Aperm = pi + p.nnzA, *Acol = pi + 2 * p.nnzA, *Acp = pi + 3 * p.nnzA, *Arp = Acp + (p.n + 1);
      u16 *Pcol = Arp + (p.m + 1), *Prp = Pcol + p.nnzP, *qoff = Prp + (p.n + 1), *qlen = qoff + p.nq;
      rowcnt.assign((size_t)p.m + 1, 0);
      for (int q = 0; q < p.nnzA; ++q) { pd[q] = A->x[q]; Arow[q] = (u16)A->i[q]; rowcnt[A->i[q] + 1]++; }
      for (int j = 0; j <= p.n; ++j) Acp[j] = (u16)A->p[j];
      for (int i = 0; i < p.m; ++i) rowcnt[i + 1] += rowcnt[i];
      for (int i = 0; i <= p.m; ++i) Arp[i] = (u16)rowcnt[i];
      rowpos.assign(rowcnt.begin(), rowcnt.end() - 1);
      for (int j = 0; j < p.n; ++j)
        for (int q = A->p[j]; q < A->p[j + 1]; ++q) {
          const int pos = rowpos[A->i[q]]++;
          Aperm[pos] = (u16)q; Acol[pos] = (u16)j;
        }
      // P: upper triangle (CSC) -> full symmetric CSR
      double *pv = pd + p.nnzA;
      if (P && p.nnzP > 0) {
        colfill.assign((size_t)p.n + 1, 0);
        for (int j = 0; j < p.n; ++j)
          for (int q = P->p[j]; q < P->p[j + 1]; ++q) {
            const int i = P->i[q];
            colfill[i + 1]++;
            if (i != j) colfill[j + 1]++;
          }
        for (int j = 0; j < p.n; ++j) colfill[j + 1] += colfill[j];
        for (int j = 0; j <= p.n; ++j) Prp[j] = (u16)colfill[j];
        rowpos.assign(colfill.begin(), colfill.end() - 1);
        // ascending columns inside each row: first the strictly-upper entries mirrored (column < row) in
        // column order, then the row's own upper part -- two sweeps over the columns keep that order
        for (int j = 0; j < p.n; ++j)       // entry (i, j), i < j, mirrored into row j at column i
          for (int q = P->p[j]; q < P->p[j + 1]; ++q)
            if (P->i[q] != j) { const int pos = rowpos[j]++; Pcol[pos] = (u16)P->i[q]; pv[pos] = P->x[q]; }
        for (int j = 0; j < p.n; ++j)       // entry (i, j), i <= j, into row i at column j
          for (int q = P->p[j]; q < P->p[j + 1]; ++q) { const int pos = rowpos[P->i[q]]++; Pcol[pos] = (u16)j; pv[pos] = P->x[q]; }
      } else {
        for (int j = 0; j <= p.n; ++j) Prp[j] = 0;
      }
      double *pb = pv + p.nnzP;
      for (int i = 0; i < p.m; ++i) pb[i] = dd->b[i];
      for (int j = 0; j < p.n; ++j) pb[p.m + j] = dd->c[j];
      {
        const int nb = p.bsize > 1 ? p.bsize - 1 : 0;
        for (int j = 0; j < nb; ++j) { pb[p.m + p.n + j] = kk->bl[j]; pb[p.m + p.n + nb + j] = kk->bu[j]; }
        for (int j = 0; j < p.np; ++j) pb[p.m + p.n + 2 * nb + j] = kk->p[j];
      }
      int off = kk->z + kk->l + kk->bsize;
      for (int c = 0; c < p.nsoc; ++c) { qoff[c] = (u16)off; qlen[c] = (u16)kk->q[c]; off += kk->q[c]; }
      for (int c = 0; c < p.ns; ++c) { const int len = kk->s[c] * (kk->s[c] + 1) / 2; qoff[p.nsoc + c] = (u16)off; qlen[p.nsoc + c] = (u16)len; off += len; }
      for (int c = 0; c < p.ncs; ++c) { const int len = kk->cs[c] * kk->cs[c]; qoff[p.nsoc + p.ns + c] = (u16)off; qlen[p.nsoc + p.ns + c] = (u16)len; off += len; }
      for (int c = p.nsoc + p.ns + p.ncs; c < p.nq; ++c) { qoff[c] = (u16)off; qlen[c] = 3; off += 3; }  // ep, ed, power cones (reference order)
    }
    const double pack_ms = std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
    // ---- device
    cudaStream_t st = nullptr;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const BLay L(dims);
    const size_t smem = L.bytes();
    int rc = 0;
    double *d_d = nullptr, *d_sol = nullptr, *d_aa = nullptr; u16 *d_i = nullptr; BProb *d_p = nullptr; BOut *d_o = nullptr;
    int *d_cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {};  // everything below belongs to the thread's arena
#define BCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fprintf(stderr, "libscsb200: batch: %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); rc = -1; } } while (0)
    if (!ar.st) BCK(cudaStreamCreateWithFlags(&ar.st, cudaStreamNonBlocking));
    st = ar.st;
    auto kern = dims.tri ? k_batch_solve<true> : k_batch_solve<false>;
    BCK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    if (!rc) BCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBT, smem));
    if (!rc && occ < 1) { fprintf(stderr, "libscsb200: batch: kernel does not fit (%zu B shared memory)\n", smem); rc = -1; }
    const int grid = rc ? 0 : std::min(nf, sms * occ);
    const size_t aa_stride = (aaws_count(dims) + 1) & ~(size_t)1;
    if (!rc) {
      BCK(ar.device(0, sizeof(double) * (size_t)std::max(1ll, dtot), (void **)&d_d));
      BCK(ar.device(1, sizeof(u16) * (size_t)std::max(1ll, itot), (void **)&d_i));
      BCK(ar.device(2, sizeof(double) * (size_t)std::max(1ll, stot), (void **)&d_sol));
      BCK(ar.device(3, sizeof(BProb) * (size_t)nf, (void **)&d_p));
      BCK(ar.device(4, sizeof(BOut) * (size_t)nf, (void **)&d_o));
      BCK(ar.device(5, sizeof(int), (void **)&d_cnt));
      BCK(ar.device(6, sizeof(double) * std::max<size_t>(1, aa_stride * (size_t)grid), (void **)&d_aa));
    }
    if (!rc) {
      BCK(cudaMemcpyAsync(d_d, h_d, sizeof(double) * (size_t)dtot, cudaMemcpyHostToDevice, st));
      BCK(cudaMemcpyAsync(d_i, h_i, sizeof(u16) * (size_t)itot, cudaMemcpyHostToDevice, st));
      BCK(cudaMemcpyAsync(d_p, probs.data(), sizeof(BProb) * (size_t)nf, cudaMemcpyHostToDevice, st));
      BCK(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
      BArgs a;
      a.probs = d_p; a.count = nf; a.dims = dims;
      a.stg.normalize = stgs->normalize; a.stg.adaptive_scale = stgs->adaptive_scale; a.stg.max_iters = stgs->max_iters;
      a.stg.aa_mem = stgs->acceleration_lookback; a.stg.aa_interval = stgs->acceleration_interval;
      a.stg.aa_type1 = stgs->acceleration_type_1; a.stg.scale = stgs->scale; a.stg.rho_x = stgs->rho_x;
      a.stg.eps_abs = stgs->eps_abs; a.stg.eps_rel = stgs->eps_rel; a.stg.eps_infeas = stgs->eps_infeas;
      a.stg.alpha = stgs->alpha; a.stg.aa_reg = stgs->acceleration_regularization;
      {
        const char *e = getenv("SCS_B200_BATCH_REFINE");
        a.stg.refine = e ? std::max(0, atoi(e)) : 1;
        a.stg.pad = 0;
      }
      a.dpool = d_d; a.ipool = d_i; a.sol = d_sol; a.out = d_o; a.aaws = d_aa; a.aaws_stride = aa_stride; a.counter = d_cnt;
      if (!ar.e0) { BCK(cudaEventCreate(&ar.e0)); BCK(cudaEventCreate(&ar.e1)); }
      e0 = ar.e0; e1 = ar.e1;
      BCK(cudaEventRecord(e0, st));
      kern<<<grid, kBT, smem, st>>>(a);
      BCK(cudaGetLastError());
      BCK(cudaEventRecord(e1, st));
      BCK(cudaMemcpyAsync(h_sol, d_sol, sizeof(double) * (size_t)stot, cudaMemcpyDeviceToHost, st));
      BCK(cudaMemcpyAsync(h_out, d_o, sizeof(BOut) * (size_t)nf, cudaMemcpyDeviceToHost, st));
      BCK(cudaStreamSynchronize(st));
    }
    if (rc) {
      cleanup();
      for (int f = 0; f < nf; ++f)
        populate_on_failure(d[fused[f]]->m, d[fused[f]]->n, sol[fused[f]], &info[fused[f]], SCS_FAILED, "failure");
      return SCS_FAILED;
    }
    float kms = 0.f;
    cudaEventElapsedTime(&kms, e0, e1);
    for (int f = 0; f < nf; ++f) {
      const int i = fused[f];
      const BProb &p = probs[f];
      const BOut &o = h_out[f];
      ScsSolution *so = sol[i];
      ScsInfo *in = &info[i];
      if (!so->x) so->x = (double *)calloc(p.n, sizeof(double));
      if (!so->y) so->y = (double *)calloc(p.m, sizeof(double));
      if (!so->s) so->s = (double *)calloc(p.m, sizeof(double));
      memcpy(so->x, h_sol + p.sol_off, sizeof(double) * p.n);
      memcpy(so->y, h_sol + p.sol_off + p.n, sizeof(double) * p.m);
      memcpy(so->s, h_sol + p.sol_off + p.n + p.m, sizeof(double) * p.m);
      memset(in, 0, sizeof(ScsInfo));
      in->iter = o.iter; in->status_val = o.status_val; in->scale_updates = o.scale_updates;
      status_string(o.status_val, o.iter, stgs->max_iters, in->status);
      snprintf(in->lin_sys_solver, sizeof(in->lin_sys_solver), "%s", scs_get_lin_sys_method());
      in->pobj = o.pobj; in->dobj = o.dobj; in->res_pri = o.res_pri; in->res_dual = o.res_dual; in->gap = o.gap;
      in->res_infeas = o.res_infeas; in->res_unbdd_a = o.res_unbdd_a; in->res_unbdd_p = o.res_unbdd_p;
      in->setup_time = o.setup_ms; in->solve_time = o.solve_ms; in->scale = o.scale; in->comp_slack = o.comp_slack;
      in->rejected_accel_steps = o.rej; in->accepted_accel_steps = o.acc;
      in->aa_stats.iter = o.aa_iter; in->aa_stats.n_accept = o.n_accept; in->aa_stats.n_reject_rank0 = o.n_reject_rank0;
      in->aa_stats.n_reject_nonfinite = o.n_reject_nonfinite; in->aa_stats.n_reject_weight_cap = o.n_reject_weight_cap;
      in->aa_stats.n_safeguard_reject = o.n_safeguard_reject; in->aa_stats.last_rank = o.last_rank;
      in->aa_stats.last_aa_norm = o.last_aa_norm; in->aa_stats.last_regularization = o.last_reg;
      in->lin_sys_time = in->cone_time = in->accel_time = 0.0;
      if (o.status_val == SCS_FAILED) worst = SCS_FAILED;
      g_last_batch.cg_its += o.cg_its; g_last_batch.iters += o.iter;
      for (int q = 0; q < 6; ++q) g_last_batch.clk[q] += o.clk[q];
    }
    g_last_batch.fused = nf; g_last_batch.launches += 1; g_last_batch.kernel_ms = kms; g_last_batch.pack_ms = pack_ms;
    g_last_batch.h2d = (long long)(sizeof(double) * dtot + sizeof(u16) * itot + sizeof(BProb) * nf);
    g_last_batch.d2h = (long long)(sizeof(double) * stot + sizeof(BOut) * nf);
    g_last_batch.ctas = grid; g_last_batch.smem = (int)smem; g_last_batch.direct = dims.direct;
    cleanup();
#undef BCK
  }
  // ---- everything else: the streaming engine, one problem after another
  for (int i = 0; i < count; ++i) {
    if (elig[i].ok || elig[i].nq == -1) continue;
    const scs_int stv = scs(d[i], k[i], stgs, sol[i], &info[i]);
    g_last_batch.streamed++;
    if (stv < 0 && stv != SCS_INFEASIBLE && stv != SCS_UNBOUNDED) worst = stv;
  }
  return worst;
}

/* measurement hook: how the last scs_b200_solve_batch call on this thread's process ran */
extern "C" scs_int scs_b200_batch_stats(double out[18]) {
  if (!out) return -1;
  out[0] = (double)g_last_batch.fused; out[1] = (double)g_last_batch.streamed; out[2] = g_last_batch.kernel_ms;
  out[3] = g_last_batch.pack_ms; out[4] = (double)g_last_batch.h2d; out[5] = (double)g_last_batch.d2h;
  out[6] = (double)g_last_batch.ctas; out[7] = (double)g_last_batch.smem;
  out[8] = (double)g_last_batch.direct; out[9] = (double)g_last_batch.cg_its; out[10] = (double)g_last_batch.iters; out[11] = 0.0;
  for (int q = 0; q < 6; ++q) out[12 + q] = (double)g_last_batch.clk[q];
  return 0;
}
