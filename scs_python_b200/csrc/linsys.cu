// linsys.cu -- PCG kernels, the LinSys driver and the ScsLinSysWork plugin ABI.
#include "dist.cuh"
#include "linsys.cuh"

namespace b200 {

// Inside the graph-launched iteration the CG loop is a WHILE conditional node: the kernels
// that decide convergence also set the loop condition.  use_h == 0 on the stream path.
__device__ __forceinline__ void cg_set_loop(const CgCtl &c, int keep_going) {
  if (c.use_h) cudaGraphSetConditional(c.h, keep_going ? 1u : 0u);
}

// ------------------------------------------------------------------------- epilogues ---
// b[j] += (A' tmp)_j                                     (private.c:297)
struct EpiRhs : EpiNoState {
  double *b;
  __device__ __forceinline__ void row(State &, int j, double acc) const { b[j] += acc; }
};
// out[i] = (A x)_i / R_y,i                               (private.c:116-117)
struct EpiScaleRy {
  static constexpr bool kSeparate = false;
  struct State {};
  double *out;
  const double *ry;  // diag_r + n
  DevScalars *S_;
  int kt_cont = 0;   // 1: the timing window was opened by the streaming kernel of tiled_launch
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  __device__ __forceinline__ void init(State &) const { if (!kt_cont) kt_begin(S_, 0); }
  __device__ __forceinline__ void row(State &, int i, double acc) const { out[i] = acc / ry[i]; }
  // two-stage form used by tiled_epilogue_kernel: all global loads of a batch of rows first, then the stores
  struct Pre { double ry; };
  __device__ __forceinline__ Pre load(int i) const { return Pre{ry[i]}; }
  __device__ __forceinline__ void apply(State &, int i, double acc, const Pre &q) const { out[i] = acc / q.ry; }
  __device__ __forceinline__ void finish(State &, const RedWs &, DevScalars *S) const { kt_end_ticket(S, 0); }
};
// Row-partitioned CG: out[i] = (A_g x)_i / R_y,i and the local part of p'Gp that lives in m-space,
// (A_g p)' R_y^-1 (A_g p) = sum_i out_i (A_g p)_i  ->  S->pGp_local (no collective: it rides on the
// all-reduce of A_g' z_g, see LinSys::launch_G)
struct EpiScaleRyDot {
  static constexpr bool kSeparate = false;
  struct State { double dot; };
  double *out;
  const double *ry;
  DevScalars *S_;
  int kt_cont = 0;
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  __device__ __forceinline__ void init(State &s) const { s.dot = 0.0; if (!kt_cont) kt_begin(S_, 0); }
  __device__ __forceinline__ void row(State &s, int i, double acc) const {
    const double o = acc / ry[i];
    out[i] = o;
    s.dot = fma(o, acc, s.dot);
  }
  struct Pre { double ry; };
  __device__ __forceinline__ Pre load(int i) const { return Pre{ry[i]}; }
  __device__ __forceinline__ void apply(State &s, int i, double acc, const Pre &q) const {
    const double o = acc / q.ry;
    out[i] = o;
    s.dot = fma(o, acc, s.dot);
  }
  __device__ __forceinline__ void finish(State &s, const RedWs &ws, DevScalars *S) const {
    double v[1] = {s.dot};
    grid_reduce<1, 0>(v, ws, [S](double *o) {
      S->pGp_local = o[0];
      kt_end_last(S, 0);
    });
  }
};
// Row-partitioned CG: pp_j = (P p)_j + R_x,j p_j over the local columns, and this rank's whole
// contribution to p'Gp: pGp_local + sum over the columns it counts of p_j pp_j, written to *slot (the
// scalar that is all-reduced together with the shared block of A_g' z_g)
struct EpiPP {
  static constexpr bool kSeparate = false;
  struct State { double dot; };
  double *pp;
  const double *p, *rx;
  double *slot;
  int cnt_lo;
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  __device__ __forceinline__ void init(State &s) const { s.dot = 0.0; }
  __device__ __forceinline__ void row(State &s, int j, double acc) const {
    const double pj = p[j];
    const double g = acc + rx[j] * pj;
    pp[j] = g;
    if (j >= cnt_lo) s.dot = fma(pj, g, s.dot);
  }
  __device__ __forceinline__ void finish(State &s, const RedWs &ws, DevScalars *S) const {
    double v[1] = {s.dot};
    double *sl = slot;
    grid_reduce<1, 0>(v, ws, [S, sl](double *o) { *sl = S->pGp_local + o[0]; });
  }
};
// y[row] = acc, usable by both engines (the tiled epilogue pass needs load / apply)
struct EpiStoreT : EpiNoState {
  double *y;
  DevScalars *S_;
  int kt_cat = -1;  // timing category closed by this product (-1: none)
  int kt_cont = 0;
  __device__ __forceinline__ void init(State &) const { if (kt_cat >= 0 && !kt_cont) kt_begin(S_, kt_cat); }
  __device__ __forceinline__ void row(State &, int r, double acc) const { y[r] = acc; }
  struct Pre {};
  __device__ __forceinline__ Pre load(int) const { return Pre{}; }
  __device__ __forceinline__ void apply(State &, int r, double acc, const Pre &) const { y[r] = acc; }
  __device__ __forceinline__ void finish(State &, const RedWs &, DevScalars *S) const { if (kt_cat >= 0) kt_end_ticket(S, kt_cat); }
};
// Gp_j = (A' z)_j + (P p)_j + R_x,j p_j ; p'Gp ; alpha = z'r / p'Gp     (private.c:181-183)
struct EpiG {
  static constexpr bool kSeparate = false;
  struct State { double pgp; };
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  double *Gp;
  const double *p, *rx;
  DevScalars *S_;
  const double *extra;  // row-partitioned mode: all-reduced A'z (acc then only holds P p), else null
  int kt_cont = 0;      // 1: the timing window was opened by the streaming kernel of tiled_launch
  __device__ __forceinline__ void init(State &s) const { s.pgp = 0.0; if (!kt_cont) kt_begin(S_, 1); }
  __device__ __forceinline__ void row(State &s, int j, double acc) const {
    const double pj = p[j];
    if (extra) acc += extra[j];
    const double g = acc + rx[j] * pj;
    Gp[j] = g;
    s.pgp = fma(pj, g, s.pgp);
  }
  struct Pre { double p, rx; };
  __device__ __forceinline__ Pre load(int j) const { return Pre{p[j], rx[j]}; }
  __device__ __forceinline__ void apply(State &s, int j, double acc, const Pre &q) const {  // == row(), extra == null
    const double g = acc + q.rx * q.p;
    Gp[j] = g;
    s.pgp = fma(q.p, g, s.pgp);
  }
  __device__ __forceinline__ void finish(State &s, const RedWs &ws, DevScalars *S) const {
    double v[1] = {s.pgp};
    grid_reduce<1, 0>(v, ws, [S](double *o) {
      S->pGp = o[0];
      S->alpha = S->ztr / o[0];
      kt_end_last(S, 1);
    });
  }
};
// start of CG: z'r and ||r||_inf are in, decide whether to iterate at all (private.c:170-174)
struct FinCgStart {
  CgCtl ctl;
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    if (S->cg_done) return;  // deferred call (row-partitioned mode) after a kernel that exited at once
    S->ztr = o[0];
    S->norm_r = o[1];
    const int done = (o[1] < fmax(S->cg_tol, 1e-12)) ? 1 : 0;
    S->cg_done = done;
    cg_set_loop(ctl, !done);
  }
};
// one CG step done: beta, the stop test (private.c:189-213)
struct FinCgUpdate {
  CgCtl ctl;
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    if (S->cg_done) return;  // deferred call (row-partitioned mode) after a kernel that exited at once
    const double ztr_prev = S->ztr;
    const int its = S->cg_its + 1;
    S->cg_its = its;
    S->cg_its_total += 1;
    S->norm_r = o[1];
    S->ztr = o[0];
    S->beta = o[0] / ztr_prev;
    const int done = (o[1] < S->cg_tol || ztr_prev == 0.0 || !(o[1] == o[1]) || its >= S->cg_max_its) ? 1 : 0;
    if (done) S->cg_done = 1;
    cg_set_loop(ctl, !done);
  }
};
// warm-started start of CG: r = b - G s ; x = s ; z = M r ; p = z ; z'r ; ||r||_inf
//                                                       (private.c:153-174)
struct EpiG0 {
  static constexpr bool kSeparate = false;
  struct State { double ztr, nr; };
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  double *b;  // in: rhs, out: x = s
  const double *s, *rx, *M;
  double *r, *z, *p;
  CgCtl ctl;
  const double *extra;  // row-partitioned mode: all-reduced A' R_y^-1 A s (may alias r), else null
  int cnt_lo = 0;       // row-partitioned mode: z'r counts the columns >= cnt_lo (shared block on rank 0 only)
  int kt_cont = 0;      // (set by tiled_launch; this epilogue carries no timing hooks)
  __device__ __forceinline__ void init(State &st) const { st.ztr = 0.0; st.nr = 0.0; }
  __device__ __forceinline__ void row(State &st, int j, double acc) const {
    const double sj = s[j];
    if (extra) acc += extra[j];
    const double rj = b[j] - (acc + rx[j] * sj);
    const double zj = rj * M[j];
    b[j] = sj;
    r[j] = rj;
    z[j] = zj;
    p[j] = zj;
    if (j >= cnt_lo) st.ztr = fma(zj, rj, st.ztr);
    st.nr = fmax(st.nr, fabs(rj));
  }
  struct Pre { double b, s, rx, M; };
  __device__ __forceinline__ Pre load(int j) const { return Pre{b[j], s[j], rx[j], M[j]}; }
  __device__ __forceinline__ void apply(State &st, int j, double acc, const Pre &q) const {  // == row(), extra == null
    const double rj = q.b - (acc + q.rx * q.s);
    const double zj = rj * q.M;
    b[j] = q.s;
    r[j] = rj;
    z[j] = zj;
    p[j] = zj;
    st.ztr = fma(zj, rj, st.ztr);
    st.nr = fmax(st.nr, fabs(rj));
  }
  __device__ __forceinline__ void finish(State &st, const RedWs &ws, DevScalars *S) const {
    double v[2] = {st.ztr, st.nr};
    grid_reduce_fin<1, 1>(v, ws, S, FinCgStart{ctl});
  }
};
// y_i = ((A x)_i - ry_i) / R_y,i   written over ry      (private.c:304-309)
struct EpiY : EpiNoState {
  double *by;        // b + n
  const double *ry;  // diag_r + n
  int kt_cont = 0;   // (set by tiled_launch; this epilogue carries no timing hooks)
  __device__ __forceinline__ void row(State &, int i, double acc) const { by[i] = (acc - by[i]) / ry[i]; }
  struct Pre { double by, ry; };
  __device__ __forceinline__ Pre load(int i) const { return Pre{by[i], ry[i]}; }
  __device__ __forceinline__ void apply(State &, int i, double acc, const Pre &q) const { by[i] = (acc - q.by) / q.ry; }
};
// M_j = 1 / (R_x,j + sum_k A_kj^2 / R_y,k + P_jj)        (private.c:60-80)
struct EpiPrecond : EpiNoState {
  double *M;
  const double *rx, *Pdiag;
  __device__ __forceinline__ void row(State &, int j, double acc) const { M[j] = 1.0 / ((rx[j] + acc) + Pdiag[j]); }
};

// --------------------------------------------------------------------------- kernels ---
// tmp = ry ./ R_y                                         (private.c:293-295)
__global__ void __launch_bounds__(kThreads)
k_scale_ry(const double *__restrict__ by, const double *__restrict__ ry, double *__restrict__ tmp, int m,
           const int *skip) {
  if (*skip) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) tmp[i] = by[i] / ry[i];
}

// cold start of CG (s == NULL): r = b ; x = 0 ; z = M r ; p = z      (private.c:147-152,170-174)
__global__ void __launch_bounds__(kThreads)
k_cg_init_cold(double *__restrict__ b, const double *__restrict__ M, double *__restrict__ r,
               double *__restrict__ z, double *__restrict__ p, int n, int cnt_lo, RedWs ws, DevScalars *S, CgCtl ctl) {
  if (S->cg_done) return;
  double v[2] = {0.0, 0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double rj = b[j];
    const double zj = rj * M[j];
    b[j] = 0.0;
    r[j] = rj;
    z[j] = zj;
    p[j] = zj;
    if (j >= cnt_lo) v[0] = fma(zj, rj, v[0]);
    v[1] = fmax(v[1], fabs(rj));
  }
  grid_reduce_fin<1, 1>(v, ws, S, FinCgStart{ctl});
}

// x += alpha p ; r -= alpha Gp ; z = M r ; z'r ; ||r||_inf ; stop test ; beta
//                                                       (private.c:184-213)
__global__ void __launch_bounds__(kThreads)
k_cg_update(double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, const double *__restrict__ p,
            const double *__restrict__ Gp, const double *__restrict__ M, int n, RedWs ws, DevScalars *S, CgCtl ctl) {
  if (S->cg_done) return;
  const double alpha = S->alpha;
  double v[2] = {0.0, 0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    x[j] = fma(alpha, p[j], x[j]);
    const double rj = fma(-alpha, Gp[j], r[j]);
    const double zj = rj * M[j];
    r[j] = rj;
    z[j] = zj;
    v[0] = fma(zj, rj, v[0]);
    v[1] = fmax(v[1], fabs(rj));
  }
  grid_reduce_fin<1, 1>(v, ws, S, FinCgUpdate{ctl});
}

// the same step of a row-partitioned solve: Gp = (all-reduced A'z) + pp with pp = P p + R_x p, and
// alpha = z'r / p'Gp with p'Gp = the scalar that was all-reduced together with the shared block (Gs[-1]);
// z'r counts the columns >= cnt_lo
__global__ void __launch_bounds__(kThreads)
k_cg_update_dist(double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, const double *__restrict__ p,
                 const double *__restrict__ Gs, const double *__restrict__ pp, const double *__restrict__ M, int n,
                 int cnt_lo, RedWs ws, DevScalars *S, CgCtl ctl) {
  if (S->cg_done) return;
  const double alpha = S->ztr / Gs[-1];
  double v[2] = {0.0, 0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    x[j] = fma(alpha, p[j], x[j]);
    const double rj = fma(-alpha, Gs[j] + pp[j], r[j]);
    const double zj = rj * M[j];
    r[j] = rj;
    z[j] = zj;
    if (j >= cnt_lo) v[0] = fma(zj, rj, v[0]);
    v[1] = fmax(v[1], fabs(rj));
  }
  grid_reduce_fin<1, 1>(v, ws, S, FinCgUpdate{ctl});
}
// r = b - (Gs + pp) ; x = s ; z = M r ; p = z   (warm-started CG start of a row-partitioned solve)
__global__ void __launch_bounds__(kThreads)
k_cg_start_dist(double *__restrict__ b, const double *__restrict__ s, const double *__restrict__ Gs,
                const double *__restrict__ pp, const double *__restrict__ M, double *__restrict__ r,
                double *__restrict__ z, double *__restrict__ p, int n, int cnt_lo, RedWs ws, DevScalars *S, CgCtl ctl) {
  if (S->cg_done) return;
  double v[2] = {0.0, 0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double rj = b[j] - (Gs[j] + pp[j]);
    const double zj = rj * M[j];
    b[j] = s[j];
    r[j] = rj;
    z[j] = zj;
    p[j] = zj;
    if (j >= cnt_lo) v[0] = fma(zj, rj, v[0]);
    v[1] = fmax(v[1], fabs(rj));
  }
  grid_reduce_fin<1, 1>(v, ws, S, FinCgStart{ctl});
}

// p = z + beta p                                         (private.c:213-216)
__global__ void __launch_bounds__(kThreads)
k_cg_pupdate(double *__restrict__ p, const double *__restrict__ z, int n, const DevScalars *S) {
  if (S->cg_done) return;
  const double beta = S->beta;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    p[j] = fma(beta, p[j], z[j]);
}

// b = 0 when the right-hand side was (numerically) zero   (private.c:288-291)
__global__ void __launch_bounds__(kThreads) k_zero_if(double *__restrict__ b, int len, const int *flag) {
  if (!*flag) return;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) b[j] = 0.0;
}

// S->cg_tol = tol ; zero_rhs = (||b||_inf <= 1e-12) ; cg_done = zero_rhs ; cg_its = 0
struct FinPrepareFlags {
  double tol;
  __device__ __forceinline__ void operator()(double *o, DevScalars *S) const {
    S->cg_tol = tol;
    S->zero_rhs = (o[0] <= 1e-12) ? 1 : 0;
    S->cg_done = S->zero_rhs;
    S->cg_its = 0;
  }
};
__global__ void __launch_bounds__(kThreads)
k_prepare_flags(const double *__restrict__ b, int len, double tol, RedWs ws, DevScalars *S) {
  double v[1] = {0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x)
    v[0] = fmax(v[0], fabs(b[j]));
  grid_reduce_fin<0, 1>(v, ws, S, FinPrepareFlags{tol});
}

// row-partitioned mode helpers: b[j] += a[j] ; M = 1 / ((rx + colsum) + Pdiag)
__global__ void __launch_bounds__(kThreads)
k_add_if(double *__restrict__ b, const double *__restrict__ a, int n, const int *skip) {
  if (*skip) return;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) b[j] += a[j];
}
__global__ void __launch_bounds__(kThreads)
k_precond_fin(double *__restrict__ M, const double *__restrict__ rx, const double *__restrict__ Pdiag, int n) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
    M[j] = 1.0 / ((rx[j] + M[j]) + Pdiag[j]);
}

// ---------------------------------------------------------------------------- driver ---
static inline int ew_grid(const Ctx &c, long long len) {
  long long g = (len + kThreads - 1) / kThreads;
  if (g < 1) g = 1;
  return (int)(g < c.grid_ew() ? g : c.grid_ew());
}

int LinSys::init(Ctx *ctx, const ScsMatrix *Ah, const ScsMatrix *Ph) {
  c = ctx;
  n = Ah->n;
  m = Ah->m;
  CUDA_OK(cudaSetDevice(c->device));
  if (csr_upload(*c, At, n, m, Ah->p, Ah->i, Ah->x)) return -1;  // CSC(A) == CSR(A')
  if (csr_transpose(*c, At, A)) return -1;                       // CSR(A) on the device
  hasP = (Ph != nullptr);
  std::vector<double> pd((size_t)n, 0.0);
  if (hasP) {
    if (csr_from_upper_csc(*c, P, n, Ph->p, Ph->i, Ph->x)) return -1;
    for (int j = 0; j < n; ++j) {  // diagonal is the last entry of an upper-tri sorted column
      const int last = Ph->p[j + 1] - 1;
      if (last >= Ph->p[j] && Ph->i[last] == j) pd[j] = Ph->x[last];
    }
  } else {
    P = CsrDev();
    P.nrows = P.ncols = n;
    if (dev_alloc_zero(&P.ptr, (size_t)n + 1, c->stream)) return -1;  // empty rows
  }
  if (dev_alloc(&Pdiag, (size_t)n) || h2d(*c, Pdiag, pd.data(), (size_t)n)) return -1;
  if (c->sync()) return -1;
  if (chunks_build(*c, chA, A, nullptr)) return -1;
  // row-partitioned mode: A_g' z_g is all-reduced before P p + R_x p is added, so the two row sets get
  // their own chunk lists (and their own launches)
  if (chunks_build(*c, chAt, At, (hasP && !c->dist) ? &P : nullptr)) return -1;
  if (c->dist && chunks_build(*c, chP, P, nullptr)) return -1;
  {
    const long long mi = 10ll * n;  // private.c:299
    const int mi32 = mi > 0x7fffffffLL ? 0x7fffffff : (int)mi;
    CUDA_OK(cudaMemcpyAsync(&c->S->cg_max_its, &mi32, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  // Gp carries kGpFront doubles in front of it: Gp[-1] is the scalar p'Gp of a row-partitioned solve,
  // all-reduced in one call with the shared block Gp[0, n_sh) (Gp[-2] pads the region to 16 bytes)
  // Row-partitioned mode: that region lives in an arena the other ranks map (CUDA IPC) so that the all-reduce is a
  // kernel over peer memory (dist.cu: dist_p2p_setup); when mapping is not possible it is an ordinary allocation
  // and the exchange goes through NCCL.
  if (c->dist) {
    double *vec = nullptr;
    if (dist_p2p_setup(*c, (size_t)n + kGpFront, &vec)) return -1;
    if (c->p2p) { Gp_base = vec; Gp_in_arena = true; }
  }
  if (dev_alloc_zero(&M, (size_t)n, c->stream) || dev_alloc_zero(&p, (size_t)n, c->stream) ||
      dev_alloc_zero(&r, (size_t)n, c->stream) ||
      (!Gp_in_arena && dev_alloc_zero(&Gp_base, (size_t)n + kGpFront, c->stream)) ||
      dev_alloc_zero(&z, (size_t)n, c->stream) || dev_alloc_zero(&tmp, (size_t)m, c->stream))
    return -1;
  Gp = Gp_base + kGpFront;
  if (c->dist && dev_alloc_zero(&pp, (size_t)n, c->stream)) return -1;
  return 0;
}

int LinSys::finalize_structure() {
  tA.destroy();
  tG.destroy();
  const int mode = tiled_env_mode();  // SCS_B200_TILED: 0 never, 1 force, unset: heuristic
  if (mode == 0) return 0;
  if (tiled_prepare()) return -1;
  if (tA.build(*c, A, nullptr, mode == 1)) return -1;
  // row-partitioned mode: A_g' alone (P p + R_x p is a separate pass after the all-reduce)
  if (tG.build(*c, At, (hasP && !c->dist) ? &P : nullptr, mode == 1)) return -1;
  return 0;
}

void LinSys::destroy() {
  if (!c) return;
  cudaSetDevice(c->device);
  csr_free(A);
  csr_free(At);
  csr_free(P);
  tA.destroy();
  tG.destroy();
  chunks_free(chA);
  chunks_free(chAt);
  chunks_free(chP);
  if (own_diag_r) dev_free(diag_r);
  diag_r = nullptr;
  dev_free(Pdiag);
  dev_free(M); dev_free(p); dev_free(r); dev_free(z); dev_free(tmp); dev_free(pp);
  if (Gp_in_arena) { dist_p2p_teardown(*c); Gp_base = nullptr; Gp_in_arena = false; }
  else dev_free(Gp_base);
  Gp = nullptr;
}

// Refresh P's diagonal from the (possibly re-scaled) device copy of P.
__global__ void __launch_bounds__(kThreads) k_extract_diag(CsrDev P, double *__restrict__ d) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.nrows; j += gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int k = P.ptr[j]; k < P.ptr[j + 1]; ++k)
      if (P.idx[k] == j) v = P.val[k];
    d[j] = v;
  }
}

int LinSys::update_precond() {
  if (hasP) {
    k_extract_diag<<<ew_grid(*c, n), kThreads, 0, c->stream>>>(P, Pdiag);
    c->launches++;
  }
  ElemSqDiv e{diag_r + n};
  if (c->dist) {  // column sums of the local rows, summed over the ranks, then the same formula
    EpiStore es; es.y = M;
    row_kernel<ElemSqDiv, ElemSqDiv, EpiStore, false>
        <<<chAt.grid, kThreads, 0, c->stream>>>(At, e, At, e, chAt.d, chAt.n, es, c->red, c->S, nullptr);
    if (dist_allreduce(*c, M, (size_t)c->n_sh, 0)) return -1;  // private columns are complete locally
    k_precond_fin<<<ew_grid(*c, n), kThreads, 0, c->stream>>>(M, diag_r, Pdiag, n);
    c->launches += 2;
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  EpiPrecond epi;
  epi.M = M; epi.rx = diag_r; epi.Pdiag = Pdiag;
  row_kernel<ElemSqDiv, ElemSqDiv, EpiPrecond, false>
      <<<chAt.grid, kThreads, 0, c->stream>>>(At, e, At, e, chAt.d, chAt.n, epi, c->red, c->S, nullptr);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// row-partitioned CG: the same product with the m-space part of p'Gp reduced on the side (EpiScaleRyDot)
int LinSys::launch_A_scaled_dot(const double *x, double *out, const int *skip) {
  EpiScaleRyDot epi;
  epi.out = out; epi.ry = diag_r + n; epi.S_ = c->S;
  ElemMul e{x};
  if (tA.ok && tiled_aligned(x))
    tiled_launch(tA, x, nullptr, epi, *c, skip, 0);
  else
    row_kernel<ElemMul, ElemMul, EpiScaleRyDot, false>
        <<<chA.grid, kThreads, 0, c->stream>>>(A, e, A, e, chA.d, chA.n, epi, c->red, c->S, skip);
  return 0;
}

// row-partitioned mode: out[0, n) = A_g' zin of the local rows, shared block [0, n_sh) summed over the
// ranks in place.  with_scalar: out[-1] (this rank's contribution to a scalar, written beforehand) rides along.
int LinSys::dist_At(const double *zin, double *out, const int *skip, bool with_scalar, int kt_cat) {
  EpiStoreT es;
  es.y = out; es.S_ = c->S; es.kt_cat = kt_cat;
  ElemMul ea{zin};
  if (tG.ok && tiled_aligned(zin))
    tiled_launch(tG, zin, nullptr, es, *c, skip, kt_cat);
  else
    row_kernel<ElemMul, ElemMul, EpiStoreT, false>
        <<<chAt.grid, kThreads, 0, c->stream>>>(At, ea, At, ea, chAt.d, chAt.n, es, c->red, c->S, skip);
  return dist_reduce_gp(out, with_scalar, skip);
}

// sum over the ranks, in place, of the shared block of `out` (== Gp) [and the two scalars in front of it]
int LinSys::dist_reduce_gp(double *out, bool with_scalar, const int *skip) {
  if (c->p2p && out == Gp)
    return dist_p2p_allreduce(*c, with_scalar ? kGpFront - 2 : kGpFront, (long long)c->n_sh + (with_scalar ? 2 : 0), skip);
  if (with_scalar) return dist_allreduce(*c, out - 2, (size_t)c->n_sh + 2, 0);
  return dist_allreduce(*c, out, (size_t)c->n_sh, 0);
}

int LinSys::launch_A_scaled(const double *x, double *out, const int *skip, int tag, bool counted) {
  EpiScaleRy epi;
  epi.out = out; epi.ry = diag_r + n; epi.S_ = c->S;
  ElemMul e{x};
  (void)tag;
  int nl = 1;
  if (tA.ok && tiled_aligned(x)) {
    nl = tiled_launch(tA, x, nullptr, epi, *c, skip, 0);
  } else
    row_kernel<ElemMul, ElemMul, EpiScaleRy, false>
        <<<chA.grid, kThreads, 0, c->stream>>>(A, e, A, e, chA.d, chA.n, epi, c->red, c->S, skip);
  if (counted) { c->launches += nl; c->spmv_calls++; }
  return 0;
}

int LinSys::launch_G(const double *zin, const double *pin, double *out, const int *skip, int tag, bool counted) {
  EpiG epi;
  epi.Gp = out; epi.p = pin; epi.rx = diag_r; epi.S_ = c->S; epi.extra = nullptr;
  ElemMul ea{zin}, eb{pin};
  (void)tag;
  if (c->dist) {
    // pp = P p + R_x p and this rank's contribution to p'Gp -> out[-1]; out = A_g' z_g of the local rows;
    // ONE all-reduce of [out[-1] | shared block of out].  The consumers (k_cg_update_dist) add pp.
    EpiPP ep;
    ep.pp = pp; ep.p = pin; ep.rx = diag_r; ep.slot = out - 1; ep.cnt_lo = c->cnt_lo;
    row_kernel<ElemMul, ElemMul, EpiPP, false>
        <<<chP.grid, kThreads, 0, c->stream>>>(P, eb, P, eb, chP.d, chP.n, ep, c->red, c->S, skip);
    if (dist_At(zin, out, skip, true, 1)) return -1;
    if (counted) { c->launches += 2; c->spmv_calls += 2; }
    return 0;
  }
  if (tG.ok && tiled_aligned(zin) && tiled_aligned(pin)) {
    const int nl = tiled_launch(tG, zin, pin, epi, *c, skip, 1);
    if (counted) { c->launches += nl; c->spmv_calls++; }
    return 0;
  } else if (hasP)
    row_kernel<ElemMul, ElemMul, EpiG, true>
        <<<chAt.grid, kThreads, 0, c->stream>>>(At, ea, P, eb, chAt.d, chAt.n, epi, c->red, c->S, skip);
  else
    row_kernel<ElemMul, ElemMul, EpiG, false>
        <<<chAt.grid, kThreads, 0, c->stream>>>(At, ea, P, eb, chAt.d, chAt.n, epi, c->red, c->S, skip);
  if (counted) { c->launches++; c->spmv_calls++; }
  return 0;
}

int LinSys::prepare_flags(const double *b, double tol) {
  k_prepare_flags<<<ew_grid(*c, n + m), kThreads, 0, c->stream>>>(b, n + m, tol, c->red, c->S);
  c->launches++;
  return dist_finish(*c, 0, 1, FinPrepareFlags{tol});
}

int LinSys::enqueue_head(double *b, const double *ws, CgCtl ctl) {
  Ctx &cx = *c;
  DevScalars *S = cx.S;
  const int *done = &S->cg_done;
  const int gn = ew_grid(cx, n), gm = ew_grid(cx, m);
  cudaStream_t st = cx.stream;
  // tmp = R_y^-1 ry ; b[:n] += A' tmp
  k_scale_ry<<<gm, kThreads, 0, st>>>(b + n, diag_r + n, tmp, m, done);
  if (cx.dist) {
    if (dist_At(tmp, Gp, done, false, -1)) return -1;  // Gp is free until the CG iterations start
    k_add_if<<<gn, kThreads, 0, st>>>(b, Gp, n, done);
    cx.launches++;
  } else {
    EpiRhs epi; epi.b = b;
    ElemMul e{tmp};
    row_kernel<ElemMul, ElemMul, EpiRhs, false>
        <<<chAt.grid, kThreads, 0, st>>>(At, e, At, e, chAt.d, chAt.n, epi, cx.red, S, done);
  }
  cx.launches += 2; cx.spmv_calls++;
  if (ws) {
    launch_A_scaled(ws, tmp, done);
    EpiG0 epi;
    epi.b = b; epi.s = ws; epi.rx = diag_r; epi.M = M; epi.r = r; epi.z = z; epi.p = p; epi.ctl = ctl;
    epi.extra = nullptr;
    ElemMul ea{tmp}, eb{ws};
    if (cx.dist) {
      // r = b - (A' R_y^-1 A s + P s + R_x s): A_g' (tmp) all-reduced into Gp, pp = P s + R_x s
      EpiPP ep;
      ep.pp = pp; ep.p = ws; ep.rx = diag_r; ep.slot = Gp - 1; ep.cnt_lo = cx.cnt_lo;
      row_kernel<ElemMul, ElemMul, EpiPP, false>
          <<<chP.grid, kThreads, 0, st>>>(P, eb, P, eb, chP.d, chP.n, ep, cx.red, S, done);
      if (dist_At(tmp, Gp, done, false, -1)) return -1;
      k_cg_start_dist<<<gn, kThreads, 0, st>>>(b, ws, Gp, pp, M, r, z, p, n, cx.cnt_lo, cx.red, S, ctl);
      if (dist_finish(cx, 1, 1, FinCgStart{ctl})) return -1;
      cx.launches += 2;
    } else if (tG.ok && tiled_aligned(tmp) && tiled_aligned(ws))
      cx.launches += tiled_launch(tG, tmp, ws, epi, cx, done) - 1;
    else if (hasP)
      row_kernel<ElemMul, ElemMul, EpiG0, true>
          <<<chAt.grid, kThreads, 0, st>>>(At, ea, P, eb, chAt.d, chAt.n, epi, cx.red, S, done);
    else
      row_kernel<ElemMul, ElemMul, EpiG0, false>
          <<<chAt.grid, kThreads, 0, st>>>(At, ea, P, eb, chAt.d, chAt.n, epi, cx.red, S, done);
    cx.launches++; cx.spmv_calls++;
  } else {
    k_cg_init_cold<<<gn, kThreads, 0, st>>>(b, M, r, z, p, n, cx.cnt_lo, cx.red, S, ctl);
    if (dist_finish(cx, 1, 1, FinCgStart{ctl})) return -1;
    cx.launches++;
  }
  return 0;
}

int LinSys::enqueue_cg_iter(double *b, CgCtl ctl, int tag) {
  Ctx &cx = *c;
  DevScalars *S = cx.S;
  const int *done = &S->cg_done;
  const int gn = ew_grid(cx, n);
  // not added to cx.launches: the number of CG iterations that really execute is decided on
  // the device (S->cg_its_total); launch totals are 4 * that counter + cx.launches
  if (cx.dist) {
    launch_A_scaled_dot(p, tmp, done);
    if (launch_G(tmp, p, Gp, done, tag, false)) return -1;
    k_cg_update_dist<<<gn, kThreads, 0, cx.stream>>>(b, r, z, p, Gp, pp, M, n, cx.cnt_lo, cx.red, S, ctl);
    if (dist_finish(cx, 1, 1, FinCgUpdate{ctl})) return -1;
    cx.launches--;  // (dist_finish counted its finaliser; the CG-loop kernels are counted per executed iteration)
    k_cg_pupdate<<<gn, kThreads, 0, cx.stream>>>(p, z, n, S);
    return 0;
  }
  launch_A_scaled(p, tmp, done, tag, false);
  launch_G(tmp, p, Gp, done, tag, false);
  k_cg_update<<<gn, kThreads, 0, cx.stream>>>(b, r, z, p, Gp, M, n, cx.red, S, ctl);
  k_cg_pupdate<<<gn, kThreads, 0, cx.stream>>>(p, z, n, S);
  return 0;
}

int LinSys::enqueue_tail(double *b) {
  // y = R_y^-1 (A x - ry), or everything zero
  Ctx &cx = *c;
  DevScalars *S = cx.S;
  EpiY epi; epi.by = b + n; epi.ry = diag_r + n;
  ElemMul e{b};
  if (tA.ok && tiled_aligned(b))
    cx.launches += tiled_launch(tA, b, nullptr, epi, cx, &S->zero_rhs) - 1;
  else
    row_kernel<ElemMul, ElemMul, EpiY, false>
        <<<chA.grid, kThreads, 0, cx.stream>>>(A, e, A, e, chA.d, chA.n, epi, cx.red, S, &S->zero_rhs);
  k_zero_if<<<ew_grid(cx, n + m), kThreads, 0, cx.stream>>>(b, n + m, &S->zero_rhs);
  cx.launches += 2; cx.spmv_calls++;
  return 0;
}

// host-driven CG loop: enqueue a batch of iterations (kernels launched after convergence
// return at once), read the stop flag, repeat
int LinSys::solve_dev_loop(double *b, int first_batch) {
  Ctx &cx = *c;
  const CgCtl none{};
  const long long max_its = 10ll * n;  // private.c:299
  int batch = first_batch > 0 ? first_batch : (last_its + 1 < 1 ? 1 : last_its + 1);
  if (batch > 64) batch = 64;
  // row-partitioned mode: an iteration enqueued after convergence still pays its collectives (NCCL kernels
  // do not read the stop flag), so the first batch undershoots by one and the top-ups are small
  if (cx.dist && first_batch <= 0) batch = last_its > 2 ? last_its - 1 : 1;
  long long enq = 0;
  int its = 0;
  for (;;) {
    for (int k = 0; k < batch && enq < max_its; ++k, ++enq) enqueue_cg_iter(b, none, (int)enq);
    if (cx.fetch_scalars()) return -1;
    its = cx.S_host->cg_its;
    if (cx.S_host->cg_done || enq >= max_its) break;
    batch = cx.dist ? 2 : (batch < 32 ? batch * 2 : 64);
  }
  if (first_batch <= 0) last_its = its;
  tot_cg_its += its;
  return 0;
}

int LinSys::solve_dev(double *b, const double *ws, int first_batch) {
  if (enqueue_head(b, ws, CgCtl{}) || solve_dev_loop(b, first_batch) || enqueue_tail(b)) return -1;
  CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace b200

// ================================================================== plugin ABI (C) ====
using namespace b200;

struct SCS_LIN_SYS_WORK {
  Ctx ctx;
  LinSys ls;
  double *b = nullptr, *s = nullptr;  // device staging: n+m, n
};

static thread_local int g_device = 0;

extern "C" scs_int scs_b200_set_device(scs_int device) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return -1;
  g_device = device;
  return 0;
}
extern "C" scs_int scs_b200_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) return 0;
  return cnt;
}
namespace b200 { int current_device() { return g_device; } }

extern "C" const char *scs_get_lin_sys_method(void) { return "sparse-indirect-b200-pcg"; }

extern "C" ScsLinSysWork *scs_init_lin_sys_work(const ScsMatrix *A, const ScsMatrix *P, const scs_float *diag_r) {
  if (!A || !diag_r) return nullptr;
  SCS_LIN_SYS_WORK *w = new SCS_LIN_SYS_WORK();
  if (w->ctx.init(g_device) || w->ls.init(&w->ctx, A, P) || w->ls.finalize_structure()) {
    scs_free_lin_sys_work(w);
    return nullptr;
  }
  const int n = A->n, m = A->m;
  w->ls.own_diag_r = true;
  if (dev_alloc(&w->ls.diag_r, (size_t)n + m + 1) || dev_alloc(&w->b, (size_t)n + m) || dev_alloc(&w->s, (size_t)n) ||
      h2d(w->ctx, w->ls.diag_r, diag_r, (size_t)n + m) || w->ls.update_precond() || w->ctx.sync()) {
    scs_free_lin_sys_work(w);
    return nullptr;
  }
  return w;
}

extern "C" void scs_free_lin_sys_work(ScsLinSysWork *w) {
  if (!w) return;
  cudaSetDevice(w->ctx.device);
  cudaStreamSynchronize(w->ctx.stream);
  w->ls.destroy();
  dev_free(w->b);
  dev_free(w->s);
  w->ctx.destroy();
  delete w;
}

extern "C" scs_int scs_solve_lin_sys(ScsLinSysWork *w, scs_float *b, const scs_float *s, scs_float tol) {
  if (!w || !b) return -1;
  Ctx &c = w->ctx;
  if (cudaSetDevice(c.device) != cudaSuccess) return -1;
  const int n = w->ls.n, m = w->ls.m;
  if (tol <= 0.) B200_PRINTF("Warning: tol = %4f <= 0, likely compiled without setting INDIRECT flag.\n", tol);
  if (h2d(c, w->b, b, (size_t)n + m)) return -1;
  if (s && h2d(c, w->s, s, (size_t)n)) return -1;
  if (w->ls.prepare_flags(w->b, tol)) return -1;
  if (w->ls.solve_dev(w->b, s ? w->s : nullptr, s ? 0 : 16)) return -1;
  if (d2h(c, b, w->b, (size_t)n + m) || c.sync()) return -1;
  return 0;
}

extern "C" scs_int scs_update_lin_sys_diag_r(ScsLinSysWork *w, const scs_float *new_diag_r) {
  if (!w || !new_diag_r) return -1;
  Ctx &c = w->ctx;
  if (cudaSetDevice(c.device) != cudaSuccess) return -1;
  if (h2d(c, w->ls.diag_r, new_diag_r, (size_t)w->ls.n + w->ls.m) || w->ls.update_precond() || c.sync()) return -1;
  return 0;
}

extern "C" scs_int scs_b200_lin_sys_cg_its(const ScsLinSysWork *w) { return w ? (scs_int)w->ls.tot_cg_its : 0; }

// ------------------------------------------------------- SCS(accum_by_*) test surface --
static int accum_generic(int rows_out, int cols_in, const ScsMatrix *Mh, bool transpose_first, bool sym_upper,
                         const scs_float *x, scs_float *y) {
  Ctx c;
  if (c.init(g_device)) return -1;
  CsrDev a, t;
  ChunkList ch;
  int rc = -1;
  double *dx = nullptr, *dy = nullptr;
  do {
    if (sym_upper) {
      if (csr_from_upper_csc(c, a, Mh->n, Mh->p, Mh->i, Mh->x)) break;
    } else {
      if (csr_upload(c, t, Mh->n, Mh->m, Mh->p, Mh->i, Mh->x)) break;  // CSR(M')
      if (transpose_first) {
        if (csr_transpose(c, t, a)) break;  // CSR(M)
      } else {
        a = t;
        t = CsrDev();
      }
    }
    if (chunks_build(c, ch, a, nullptr)) break;
    if (dev_alloc(&dx, (size_t)cols_in) || dev_alloc(&dy, (size_t)rows_out)) break;
    if (h2d(c, dx, x, (size_t)cols_in) || h2d(c, dy, y, (size_t)rows_out)) break;
    EpiAccum epi; epi.y = dy;
    ElemMul e{dx};
    row_kernel<ElemMul, ElemMul, EpiAccum, false>
        <<<ch.grid, kThreads, 0, c.stream>>>(a, e, a, e, ch.d, ch.n, epi, c.red, c.S, nullptr);
    if (cudaGetLastError() != cudaSuccess) break;
    if (d2h(c, y, dy, (size_t)rows_out) || c.sync()) break;
    rc = 0;
  } while (0);
  csr_free(a);
  csr_free(t);
  chunks_free(ch);
  dev_free(dx);
  dev_free(dy);
  c.destroy();
  return rc;
}

extern "C" scs_int scs_b200_accum_by_a(const ScsMatrix *A, const scs_float *x, scs_float *y) {
  if (!A || !x || !y) return -1;
  return accum_generic(A->m, A->n, A, true, false, x, y);
}
extern "C" scs_int scs_b200_accum_by_atrans(const ScsMatrix *A, const scs_float *x, scs_float *y) {
  if (!A || !x || !y) return -1;
  return accum_generic(A->n, A->m, A, false, false, x, y);
}
extern "C" scs_int scs_b200_accum_by_p(const ScsMatrix *P, const scs_float *x, scs_float *y) {
  if (!P || !x || !y) return -1;
  return accum_generic(P->n, P->n, P, false, true, x, y);
}
