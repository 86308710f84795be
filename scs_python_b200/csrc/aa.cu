// aa.cu -- Anderson acceleration kernels (see aa.cuh for the design).
#include "aa.cuh"
#include "aa_small.cuh"
#include "dist.cuh"

namespace b200 {

constexpr int kAaT = 256;  // rows per TSQR tile == CTA size


static inline AaParams params_of(const AaDev &a) {
  AaParams p;
  p.dim = a.dim; p.mem = a.mem; p.min_len = a.min_len; p.type1 = a.type1; p.ir_max_steps = a.ir_max_steps;
  p.regularization = a.regularization; p.relaxation = a.relaxation; p.safeguard_factor = a.safeguard_factor;
  p.max_weight_norm = a.max_weight_norm;
  p.x = a.x; p.f = a.f; p.g = a.g; p.g_prev = a.g_prev; p.Y = a.Y; p.S = a.S; p.D = a.D; p.x_work = a.x_work;
  p.Rpart = a.Rpart; p.st = a.st;
  p.cnt_lo = a.cnt_lo; p.cnt_hi = a.cnt_hi;
  return p;
}


// finalisers of the reducing kernels below (functors: in the row-partitioned mode they run after the
// ranks' raw sums have been gathered, common.cuh grid_reduce_fin / dist.cuh dist_finish)
struct FinAaUpdate {
  AaState *st;
  int mem, min_len;
  __device__ __forceinline__ void operator()(double *o, DevScalars *) const {
    const int it = st->iter;
    st->success = 0;
    st->aa_norm = 0.0;
    st->do_solve = 0;
    if (it == 0) {
      st->iter = 1;
    } else {
      const int idx = (it - 1) % mem;
      st->nrm_s_col[idx] = sqrt(o[0]);
      st->nrm_y_col[idx] = sqrt(o[1]);
      st->norm_g = sqrt(o[2]);
      st->len = it < mem ? it : mem;
      if (it >= min_len) st->do_solve = 1;  // iter++ happens after the solve
      else st->iter = it + 1;
    }
  }
};
struct FinAaApply {
  AaState *st;
  double *vnorm2_out;
  __device__ __forceinline__ void operator()(double *o, DevScalars *) const {
    if (st->do_solve && st->success && vnorm2_out) *vnorm2_out = o[0];
  }
};
struct FinAaSafeguard {
  AaState *st;
  int mem;
  double safeguard_factor;
  int *rej_cnt, *acc_cnt;
  __device__ __forceinline__ void operator()(double *o, DevScalars *) const {
    if (!(st->aa_norm > 0.0) || !st->success) return;  // the kernel did not reduce anything (see its early exits)
    st->success = 0;
    const double nd = sqrt(o[0]);
    if (nd > safeguard_factor * st->norm_g) {
      st->sg_reject = 1;
      st->n_safeguard_reject++;
      aa_reset_dev(st, mem);
      if (rej_cnt) *rej_cnt += 1;
    } else {
      st->sg_reject = 0;
      if (acc_cnt) *acc_cnt += 1;
    }
  }
};
struct FinAaRollback {
  AaState *st;
  double *vnorm2_out;
  __device__ __forceinline__ void operator()(double *o, DevScalars *) const {
    if (!st->sg_reject) return;
    st->sg_reject = 0;
    if (vnorm2_out) *vnorm2_out = o[0];
  }
};

// init_accel_params (aa.c:310-324) when iter == 0, else update_accel_params (aa.c:340-390)
__global__ void __launch_bounds__(kThreads)
k_aa_update(AaParams a, const double *__restrict__ xin, const double *__restrict__ fin, RedWs ws, DevScalars *S) {
  AaState *st = a.st;
  const int it = st->iter;
  double v[3] = {0.0, 0.0, 0.0};
  if (it == 0) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.dim; j += gridDim.x * blockDim.x) {
      const double xj = xin[j], fj = fin[j];
      a.x[j] = xj; a.f[j] = fj; a.g_prev[j] = xj - fj;
    }
  } else {
    const size_t col = (size_t)((it - 1) % a.mem) * a.dim;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.dim; j += gridDim.x * blockDim.x) {
      const double xj = xin[j], fj = fin[j];
      const double s = xj - a.x[j], d = fj - a.f[j], gj = xj - fj;
      const double y = gj - a.g_prev[j];
      a.S[col + j] = s; a.D[col + j] = d; a.Y[col + j] = y; a.g[j] = gj;
      a.x[j] = xj; a.f[j] = fj; a.g_prev[j] = gj;
      if (a.x_work) a.x_work[j] = xj;
      if (j >= a.cnt_lo && j < a.cnt_hi) { v[0] = fma(s, s, v[0]); v[1] = fma(y, y, v[1]); v[2] = fma(gj, gj, v[2]); }
    }
  }
  grid_reduce_fin<3, 0>(v, ws, S, FinAaUpdate{st, a.mem, a.min_len});
}

// One column step set of the Householder elimination of the stacked [R; B] block: B is
// C x kAaT (column c at B + c*kAaT), R is len x C (row-major, stride C).  Only the first
// `len` columns are eliminated; the others (Y | g) are carried along as right-hand sides.
__device__ void tile_eliminate(double *B, int rows, double *R, int len, int C, double *red, double *sig,
                               double *coef) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const bool valid = t < rows;
  for (int j = 0; j < len; ++j) {
    const double bj = valid ? B[j * kAaT + t] : 0.0;
    for (int k = j; k < C; ++k) {
      double pv = valid ? bj * B[k * kAaT + t] : 0.0;
      pv = warp_sum(pv);
      if (lane == 0) red[k * (kAaT / 32) + wid] = pv;
    }
    __syncthreads();
    if (t >= j && t < C) {
      double s = 0.0;
      for (int w = 0; w < kAaT / 32; ++w) s += red[t * (kAaT / 32) + w];
      sig[t] = s;
    }
    __syncthreads();
    const double sj = sig[j];
    if (sj == 0.0) continue;  // column already zero below R (uniform branch)
    const double rjj = R[j * C + j];
    const double nrm = sqrt(rjj * rjj + sj);
    const double alpha = rjj >= 0.0 ? -nrm : nrm;
    const double v0 = rjj - alpha;
    const double beta = 2.0 / (v0 * v0 + sj);
    __syncthreads();
    if (t > j && t < C) {
      const double w = v0 * R[j * C + t] + sig[t];
      const double cf = beta * w;
      coef[t] = cf;
      R[j * C + t] -= cf * v0;
    }
    if (t == j) R[j * C + j] = alpha;
    __syncthreads();
    if (valid) {
      for (int k = j + 1; k < C; ++k) B[k * kAaT + t] -= coef[k] * bj;
    }
  }
  __syncthreads();
}

// shared-memory carve-up shared by both TSQR stages
struct AaSmem {
  double *B, *R, *red, *sig, *coef;
  __device__ AaSmem(double *base, int mem, int Cmax) {
    B = base;
    R = B + (size_t)Cmax * kAaT;
    red = R + (size_t)mem * Cmax;
    sig = red + (size_t)Cmax * (kAaT / 32);
    coef = sig + Cmax;
  }
};
static inline size_t aa_smem_bytes(int mem) {
  const int Cmax = 2 * mem + 1;
  return sizeof(double) * ((size_t)Cmax * kAaT + (size_t)mem * Cmax + (size_t)Cmax * (kAaT / 32) + 2 * Cmax + 8);
}

__global__ void __launch_bounds__(kAaT) k_aa_tsqr1(AaParams a) {
  extern __shared__ double smem[];
  AaState *st = a.st;
  if (!st->do_solve) return;
  const int len = st->len;
  const int C = a.type1 ? 2 * len + 1 : len + 1;
  const int Cmax = 2 * a.mem + 1;
  AaSmem sm(smem, a.mem, Cmax);
  const int t = threadIdx.x;
  for (int k = t; k < a.mem * Cmax; k += kAaT) sm.R[k] = 0.0;
  __syncthreads();
  const double *Asrc = a.type1 ? a.S : a.Y;
  const int dimc = a.cnt_hi - a.cnt_lo;  // rows this rank counts
  const int ntiles = (dimc + kAaT - 1) / kAaT;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row = a.cnt_lo + tile * kAaT + t;
    const bool valid = row < a.cnt_hi;
    const int rows = (dimc - tile * kAaT) < kAaT ? (dimc - tile * kAaT) : kAaT;
    for (int cidx = 0; cidx < C; ++cidx) {
      double val = 0.0;
      if (valid) {
        if (cidx < len) val = Asrc[(size_t)cidx * a.dim + row];
        else if (cidx < C - 1) val = a.Y[(size_t)(cidx - len) * a.dim + row];
        else val = a.g[row];
      }
      sm.B[cidx * kAaT + t] = val;
    }
    __syncthreads();
    tile_eliminate(sm.B, rows, sm.R, len, C, sm.red, sm.sig, sm.coef);
  }
  double *out = a.Rpart + (size_t)blockIdx.x * a.mem * Cmax;
  for (int k = t; k < len * C; k += kAaT) out[k] = sm.R[k];
}


// Merge `nblk` trapezoids read from `src` ([blk][mem][Cmax] slots, rows packed with stride C).  merge_out ==
// null: add the sqrt(r) I rows and solve (the final stage).  Otherwise (row-partitioned mode): write the merged
// trapezoid of this rank to merge_out and stop; the ranks' trapezoids are gathered and merged by a second call.
__global__ void __launch_bounds__(kAaT) k_aa_tsqr2(AaParams a, const double *__restrict__ src, int nblk, double *merge_out) {
  extern __shared__ double smem[];
  __shared__ double sh_sqrt_r, sh_r;
  AaState *st = a.st;
  if (!st->do_solve) return;
  const int len = st->len;
  const int C = a.type1 ? 2 * len + 1 : len + 1;
  const int Cmax = 2 * a.mem + 1;
  AaSmem sm(smem, a.mem, Cmax);
  const int t = threadIdx.x;
  if (t == 0) {  // compute_regularization + the three modes of aa.c:437-451
    double r = 0.0;
    if (a.regularization > 0) {
      auto frob = [&](const double *nc) {
        double mx = 0.0;
        for (int i = 0; i < a.mem; ++i) mx = fmax(mx, nc[i]);
        if (mx == 0.0) return 0.0;
        double ss = 0.0;
        for (int i = 0; i < a.mem; ++i) { const double q = nc[i] / mx; ss += q * q; }
        return mx * sqrt(ss);
      };
      const double ny = frob(st->nrm_y_col);
      const double na = a.type1 ? frob(st->nrm_s_col) : ny;
      r = a.regularization * na * ny;
    } else if (a.regularization < 0) {
      r = -a.regularization;
    }
    sh_r = r;
    sh_sqrt_r = r > 0 ? sqrt(r) : 0.0;
  }
  for (int k = t; k < a.mem * Cmax; k += kAaT) sm.R[k] = 0.0;
  __syncthreads();
  const double sqrt_r = sh_sqrt_r;
  const int nrows = nblk * len + (merge_out ? 0 : len);
  for (int base = 0; base < nrows; base += kAaT) {
    const int rho = base + t;
    const bool valid = rho < nrows;
    const int rows = (nrows - base) < kAaT ? (nrows - base) : kAaT;
    for (int cidx = 0; cidx < C; ++cidx) {
      double val = 0.0;
      if (valid) {
        if (rho < nblk * len) {
          const int blk = rho / len, i = rho % len;
          val = src[(size_t)blk * a.mem * Cmax + (size_t)i * C + cidx];
        } else {
          const int ar = rho - nblk * len;  // row of [sqrt(r) I | sqrt(r) I | 0]
          if (cidx == ar || (a.type1 && cidx == len + ar)) val = sqrt_r;
        }
      }
      sm.B[cidx * kAaT + t] = val;
    }
    __syncthreads();
    tile_eliminate(sm.B, rows, sm.R, len, C, sm.red, sm.sig, sm.coef);
  }
  if (merge_out) {
    for (int k = t; k < len * C; k += kAaT) merge_out[k] = sm.R[k];
    return;
  }
  if (t == 0) aa_small_solve(a, sm.R, len, C, sh_r, sm.B);
}

// f -= D gamma (+ relaxation, aa.c:393-408,640-647); refresh sum f^2
__global__ void __launch_bounds__(kThreads)
k_aa_apply(AaParams a, double *__restrict__ f, double *vnorm2_out, RedWs ws, DevScalars *S) {
  AaState *st = a.st;
  if (!st->do_solve || !st->success) return;
  const int len = st->len;
  double v[1] = {0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.dim; j += gridDim.x * blockDim.x) {
    double fj = f[j];
    double dg = 0.0;
    for (int k = 0; k < len; ++k) dg = fma(a.D[(size_t)k * a.dim + j], st->gamma[k], dg);
    fj -= dg;
    if (a.x_work) {
      double sg = 0.0;
      for (int k = 0; k < len; ++k) sg = fma(a.S[(size_t)k * a.dim + j], st->gamma[k], sg);
      const double xw = a.x_work[j] - sg;
      a.x_work[j] = xw;
      fj = a.relaxation * fj + (1.0 - a.relaxation) * xw;
    }
    f[j] = fj;
    if (j >= a.cnt_lo && j < a.cnt_hi) v[0] = fma(fj, fj, v[0]);
  }
  grid_reduce_fin<1, 0>(v, ws, S, FinAaApply{st, vnorm2_out});
}

// aa_safeguard, aa.c:856-901 (decision part)
__global__ void __launch_bounds__(kThreads)
k_aa_safeguard(AaParams a, const double *__restrict__ f_new, const double *__restrict__ x_new, int *rej_cnt,
               int *acc_cnt, RedWs ws, DevScalars *S) {
  AaState *st = a.st;
  if (!(st->aa_norm > 0.0)) return;  // scs.c:1386 gate
  if (!st->success) {                // aa.c:867-871: nothing to check -> counts as accepted
    if (blockIdx.x == 0 && threadIdx.x == 0 && acc_cnt) *acc_cnt += 1;
    return;
  }
  double v[1] = {0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.dim; j += gridDim.x * blockDim.x) {
    const double d = x_new[j] - f_new[j];
    if (j >= a.cnt_lo && j < a.cnt_hi) v[0] = fma(d, d, v[0]);
  }
  grid_reduce_fin<1, 0>(v, ws, S, FinAaSafeguard{st, a.mem, a.safeguard_factor, rej_cnt, acc_cnt});
}

// roll back to the last un-accelerated pair when the safeguard rejected (aa.c:886-897)
__global__ void __launch_bounds__(kThreads)
k_aa_rollback(AaParams a, double *__restrict__ f_new, double *__restrict__ x_new, double *vnorm2_out, RedWs ws,
              DevScalars *S) {
  AaState *st = a.st;
  if (!st->sg_reject) return;
  double v[1] = {0.0};
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.dim; j += gridDim.x * blockDim.x) {
    const double fj = a.f[j];
    f_new[j] = fj;
    x_new[j] = a.x[j];
    if (j >= a.cnt_lo && j < a.cnt_hi) v[0] = fma(fj, fj, v[0]);
  }
  grid_reduce_fin<1, 0>(v, ws, S, FinAaRollback{st, vnorm2_out});
}

__global__ void k_aa_reset(AaState *st, int mem) {
  if (threadIdx.x == 0 && blockIdx.x == 0) aa_reset_dev(st, mem);
}

// --------------------------------------------------------------------------- host -----
int AaDev::init(Ctx *ctx, int dim_, int mem_, int min_len_, int type1_, double reg, double relax, double sgf,
                double mwn, int irs) {
  c = ctx;
  const int mem_clamped = mem_ < dim_ ? mem_ : dim_;  // aa.c:663
  if (dim_ <= 0 || mem_ < 0 || !std::isfinite(reg) || relax < 0 || relax > 2 || sgf < 0 || mwn <= 0 || irs < 0 ||
      (mem_clamped > 0 && min_len_ < 1)) {
    B200_PRINTF("Invalid AA parameters.\n");
    return -1;
  }
  if (mem_clamped > kAaMaxMem) {
    B200_PRINTF("B200 backend: acceleration_lookback %d exceeds the supported maximum %d.\n", mem_clamped, kAaMaxMem);
    return -1;
  }
  dim = dim_; mem = mem_clamped; type1 = type1_ ? 1 : 0;
  cnt_lo = 0; cnt_hi = dim_;
  min_len = mem > 0 ? (min_len_ < mem ? min_len_ : mem) : 0;
  regularization = reg; relaxation = relax; safeguard_factor = sgf; max_weight_norm = mwn; ir_max_steps = irs;
  CUDA_OK(cudaSetDevice(c->device));
  if (dev_alloc_zero(&st, 1, c->stream)) return -1;
  CUDA_OK(cudaMallocHost(&st_host, sizeof(AaState)));
  memset(st_host, 0, sizeof(AaState));
  st_host->last_aa_norm = NAN;
  if (h2d(*c, st, st_host, 1)) return -1;
  if (mem <= 0) return c->sync();
  const size_t dm = (size_t)dim * mem;
  if (dev_alloc_zero(&x, (size_t)dim, c->stream) || dev_alloc_zero(&f, (size_t)dim, c->stream) ||
      dev_alloc_zero(&g, (size_t)dim, c->stream) || dev_alloc_zero(&g_prev, (size_t)dim, c->stream) ||
      dev_alloc_zero(&Y, dm, c->stream) || dev_alloc_zero(&S, dm, c->stream) || dev_alloc_zero(&D, dm, c->stream))
    return -1;
  if (relaxation != 1.0 && dev_alloc_zero(&x_work, (size_t)dim, c->stream)) return -1;
  const int ntiles = (dim + kAaT - 1) / kAaT;
  nblk = ntiles < c->sms * 2 ? ntiles : c->sms * 2;
  if (nblk < 1) nblk = 1;
  const int Cmax = 2 * mem + 1;
  if (dev_alloc_zero(&Rpart, (size_t)nblk * mem * Cmax, c->stream)) return -1;
  if (c->dist && (dev_alloc_zero(&Rsend, (size_t)c->world * mem * Cmax, c->stream) ||
                  dev_alloc_zero(&Rrecv, (size_t)c->world * mem * Cmax, c->stream)))
    return -1;
  smem1 = smem2 = aa_smem_bytes(mem);
  const size_t scratch = sizeof(double) * (size_t)(4 * mem * mem + 5 * mem + 8);
  if ((size_t)Cmax * kAaT * sizeof(double) < scratch) {
    B200_PRINTF("B200 backend: internal AA scratch sizing error.\n");
    return -1;
  }
  CUDA_OK(cudaFuncSetAttribute(k_aa_tsqr1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  CUDA_OK(cudaFuncSetAttribute(k_aa_tsqr2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  return c->sync();
}

void AaDev::destroy() {
  if (!c) return;
  cudaSetDevice(c->device);
  dev_free(x); dev_free(f); dev_free(g); dev_free(g_prev); dev_free(Y); dev_free(S); dev_free(D);
  dev_free(x_work); dev_free(Rpart); dev_free(st); dev_free(Rsend); dev_free(Rrecv);
  if (st_host) cudaFreeHost(st_host);
  st_host = nullptr;
}

int AaDev::apply(double *fv, const double *xv, double *vnorm2_out) {
  if (mem <= 0) return 0;
  AaParams p = params_of(*this);
  const int grid = c->grid_ew();
  cudaStream_t s = c->stream;
  k_aa_update<<<grid, kThreads, 0, s>>>(p, xv, fv, c->red, c->S);
  if (dist_finish(*c, 3, 0, FinAaUpdate{st, mem, min_len})) return -1;
  k_aa_tsqr1<<<nblk, kAaT, smem1, s>>>(p);
  if (c->dist) {  // this rank's trapezoid -> gather -> final merge over the ranks, identical on every rank
    const size_t slot = (size_t)mem * (2 * mem + 1);
    k_aa_tsqr2<<<1, kAaT, smem2, s>>>(p, Rpart, nblk, Rsend + (size_t)c->rank * slot);
    if (dist_allreduce_oop(*c, Rsend, Rrecv, (size_t)c->world * slot)) return -1;
    k_aa_tsqr2<<<1, kAaT, smem2, s>>>(p, Rrecv, c->world, nullptr);
    c->launches++;
  } else {
    k_aa_tsqr2<<<1, kAaT, smem2, s>>>(p, Rpart, nblk, nullptr);
  }
  k_aa_apply<<<grid, kThreads, 0, s>>>(p, fv, vnorm2_out, c->red, c->S);
  if (dist_finish(*c, 1, 0, FinAaApply{st, vnorm2_out})) return -1;
  c->launches += 4;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int AaDev::safeguard(double *f_new, double *x_new, double *vnorm2_out, int *rej_cnt, int *acc_cnt) {
  if (mem <= 0) return 0;
  AaParams p = params_of(*this);
  const int grid = c->grid_ew();
  k_aa_safeguard<<<grid, kThreads, 0, c->stream>>>(p, f_new, x_new, rej_cnt, acc_cnt, c->red, c->S);
  if (dist_finish(*c, 1, 0, FinAaSafeguard{st, mem, safeguard_factor, rej_cnt, acc_cnt})) return -1;
  k_aa_rollback<<<grid, kThreads, 0, c->stream>>>(p, f_new, x_new, vnorm2_out, c->red, c->S);
  if (dist_finish(*c, 1, 0, FinAaRollback{st, vnorm2_out})) return -1;
  c->launches += 2;
  CUDA_OK(cudaGetLastError());
  return 0;
}

int AaDev::set_counted_rows(int lo, int hi) {
  if (lo < 0 || hi > dim || hi <= lo) return -1;
  cnt_lo = lo; cnt_hi = hi;
  return 0;
}

int AaDev::reset() {
  if (!st) return 0;
  k_aa_reset<<<1, 32, 0, c->stream>>>(st, mem);
  c->launches++;
  return 0;
}

int AaDev::fetch_state() {
  CUDA_OK(cudaMemcpyAsync(st_host, st, sizeof(AaState), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->d2h += sizeof(AaState);
  return 0;
}

int current_device();

}  // namespace b200

// ====================================================== C ABI: aa.h with host vectors =
using namespace b200;

struct SCS_B200_AA_WORK {
  Ctx ctx;
  AaDev aa;
  double *f = nullptr, *x = nullptr;
};

extern "C" ScsB200AaWork *scs_b200_aa_init(scs_int dim, scs_int mem, scs_int min_len, scs_int type1,
                                           scs_float regularization, scs_float relaxation,
                                           scs_float safeguard_factor, scs_float max_weight_norm,
                                           scs_int ir_max_steps) {
  SCS_B200_AA_WORK *w = new SCS_B200_AA_WORK();
  if (w->ctx.init(current_device()) ||
      w->aa.init(&w->ctx, dim, mem, min_len, type1, regularization, relaxation, safeguard_factor, max_weight_norm,
                 ir_max_steps) ||
      dev_alloc(&w->f, (size_t)dim) || dev_alloc(&w->x, (size_t)dim)) {
    scs_b200_aa_finish(w);
    return nullptr;
  }
  return w;
}

extern "C" scs_float scs_b200_aa_apply(scs_float *f, const scs_float *x, ScsB200AaWork *w) {
  if (!w || !f || !x) return NAN;
  Ctx &c = w->ctx;
  cudaSetDevice(c.device);
  if (w->aa.mem <= 0) return 0.0;
  const size_t dim = (size_t)w->aa.dim;
  if (h2d(c, w->f, f, dim) || h2d(c, w->x, x, dim) || w->aa.apply(w->f, w->x, nullptr) || d2h(c, f, w->f, dim) ||
      w->aa.fetch_state())
    return NAN;
  return w->aa.st_host->aa_norm;
}

extern "C" scs_int scs_b200_aa_safeguard(scs_float *f_new, scs_float *x_new, ScsB200AaWork *w) {
  if (!w || !f_new || !x_new) return 0;
  Ctx &c = w->ctx;
  cudaSetDevice(c.device);
  if (w->aa.mem <= 0) return 0;
  const size_t dim = (size_t)w->aa.dim;
  if (w->aa.fetch_state()) return 0;
  const int before = w->aa.st_host->n_safeguard_reject;
  // the C API has no aa_norm gate (that lives in scs.c); force the gate open for this call
  if (h2d(c, w->f, f_new, dim) || h2d(c, w->x, x_new, dim)) return 0;
  AaState *st = w->aa.st;
  const double one = 1.0;
  cudaMemcpyAsync(&st->aa_norm, &one, sizeof(double), cudaMemcpyHostToDevice, c.stream);
  if (w->aa.safeguard(w->f, w->x, nullptr, nullptr, nullptr) || d2h(c, f_new, w->f, dim) || d2h(c, x_new, w->x, dim) ||
      w->aa.fetch_state())
    return 0;
  return (w->aa.st_host->n_safeguard_reject > before) ? -1 : 0;
}

extern "C" void scs_b200_aa_reset(ScsB200AaWork *w) {
  if (!w) return;
  cudaSetDevice(w->ctx.device);
  w->aa.reset();
  w->ctx.sync();
}

extern "C" AaStats scs_b200_aa_get_stats(ScsB200AaWork *w) {
  AaStats s;
  memset(&s, 0, sizeof(s));
  s.last_aa_norm = NAN;
  if (!w || w->aa.fetch_state()) return s;
  const AaState *h = w->aa.st_host;
  s.iter = h->iter; s.n_accept = h->n_accept; s.n_reject_lapack = h->n_reject_lapack;
  s.n_reject_rank0 = h->n_reject_rank0; s.n_reject_nonfinite = h->n_reject_nonfinite;
  s.n_reject_weight_cap = h->n_reject_weight_cap; s.n_safeguard_reject = h->n_safeguard_reject;
  s.last_rank = h->last_rank; s.last_aa_norm = h->last_aa_norm; s.last_regularization = h->last_regularization;
  return s;
}

extern "C" void scs_b200_aa_finish(ScsB200AaWork *w) {
  if (!w) return;
  cudaSetDevice(w->ctx.device);
  if (w->ctx.stream) cudaStreamSynchronize(w->ctx.stream);
  w->aa.destroy();
  dev_free(w->f); dev_free(w->x);
  w->ctx.destroy();
  delete w;
}
