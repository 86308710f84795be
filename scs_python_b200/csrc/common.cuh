// common.cuh -- shared device/host infrastructure of libscsb200 (sm_100a only).
//
//  * Ctx        : per-workspace CUDA context (device, stream, reduction scratch, counters)
//  * grid_reduce: deterministic grid-wide multi-value sum/max with "last block finishes"
//                 finalisation, so data-dependent scalars (CG alpha/beta, tau, tolerances)
//                 are produced and consumed on the device and never cross to the host.
//  * DevScalars : the block of device-resident scalars of one workspace.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/scs_b200.h"

namespace b200 {

constexpr int kThreads = 256;    // CTA size of every streaming kernel
constexpr int kMaxRedVals = 24;  // max simultaneous reduction outputs of one kernel
constexpr int kCtasPerSm = 8;    // resident 256-thread CTAs per SM (2048 threads / SM)
constexpr int kMaxWorld = 16;    // ranks of a row-partitioned solve (one box: 8)

#define B200_PRINTF(...)          \
  do {                            \
    fprintf(stdout, __VA_ARGS__); \
    fflush(stdout);               \
  } while (0)

#define CUDA_OK(call)                                                                     \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      fprintf(stderr, "libscsb200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err__), \
              __FILE__, __LINE__, cudaGetErrorString(err__));                             \
      return -1;                                                                          \
    }                                                                                     \
  } while (0)

#define CUDA_OK_NULL(call)                                                                \
  do {                                                                                    \
    cudaError_t err__ = (call);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      fprintf(stderr, "libscsb200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err__), \
              __FILE__, __LINE__, cudaGetErrorString(err__));                             \
      return nullptr;                                                                     \
    }                                                                                     \
  } while (0)

// Workspace of grid_reduce: partials[k * stride + block], ticket counter.
struct RedWs {
  double *partials;
  unsigned int *ticket;
  int stride;
};

// Device-resident scalars of one workspace.  Written by the finalising block of the
// producing kernel, read by every thread of the consuming kernels.
struct DevScalars {
  // --- ADMM ---
  double vnorm2;       // sum_j v_j^2 of the current v (for normalize_v, scs.c:771-779)
  double tau;          // u_t[l-1] from root_plus (scs.c:667-688)
  double nm_ax_s_btau; // ||Ax+s-b tau||_inf, normalised, from the last residual check
  double nm_px_aty_ctau;
  // --- PCG (cpu/indirect/private.c:135-219) ---
  double cg_tol;
  double ztr, pGp, alpha, beta, norm_r;
  int cg_done;   // 1: converged / nothing to do; CG kernels exit immediately
  int zero_rhs;  // 1: ||rhs||_inf <= 1e-12 -> solution is 0 (private.c:288-291)
  int cg_its;    // CG iterations of the current solve
  int cg_its_total;
  // --- AA (aa.c) ---
  double aa_norm;  // return value of aa_apply for the current ADMM iteration
  int aa_rejected, aa_accepted;  // safeguard counters of the current solve
  int pad0;
  // --- residual block, filled every CONVERGED_INTERVAL iterations (scs.c:513-585) ---
  double res[32];
  // --- misc results of setup / finalisation reductions ---
  double sigma;      // primal_scale == dual_scale (normalize.c:46-60)
  double fin[4];     // ||s||_inf, ||y||_inf, s'y of the un-normalised solution (scs.c:885-887)
  // --- device-side loop control of the graph-launched ADMM iteration ---
  int iter;          // ADMM iteration index, advanced by the dual-step kernel
  int cg_max_its;    // 10 n (private.c:299)
  // phase clocks in ns of %globaltimer: [0] lin-sys, [1] cones, [2] acceleration
  unsigned long long t_mark;
  unsigned long long phase_ns[4];
  // in-region device timing of the two CG SpMV kernels (bench marks): [0] z = R_y^-1 A p,
  // [1] Gp = A'z + P p + R_x p.  Start = first CTA's first instruction, end = last CTA done.
  int kt_on;
  // row-partitioned mode: 1 => grid reductions park their raw sums / maxes in this rank's slot of
  // gsend[] and the consuming formula runs in k_apply_fin after ONE sum all-reduce gsend -> grecv
  // (every rank only ever writes its own slot, so the sum is a gather; the finaliser then combines the
  // ranks' values in rank order: bit-identical scalars on every rank, sums and maxes in one collective)
  int dist;
  int dist_rank, dist_world;
  int pad1;
  unsigned int kt_ticket[2], kt_cnt[2];
  unsigned long long kt_start[2], kt_ns[2];
  double part[kMaxRedVals];
  double pGp_local;  // row-partitioned CG: (A_g p)' R_y^-1 (A_g p) of the local rows, see linsys.cu
  double gsend[kMaxWorld * kMaxRedVals], grecv[kMaxWorld * kMaxRedVals];
};

// indices into DevScalars::res
enum ResIdx {
  R_TAU = 0, R_KAP,
  R_BTY_TAU, R_CTX_TAU, R_XPX_TAU,
  R_NM_AX_S_BTAU, R_NM_AX_S, R_NM_AX,          // normalised inf-norms (m-space)
  R_NM_PX_ATY_CTAU, R_NM_PX, R_NM_ATY,         // normalised inf-norms (n-space)
  R_ONM_AX_S_BTAU, R_ONM_AX_S, R_ONM_AX, R_ONM_S,  // un-normalised (divided by D*dual_scale)
  R_ONM_PX_ATY_CTAU, R_ONM_PX, R_ONM_ATY,      // un-normalised (divided by E*primal_scale)
  R_COUNT
};

// Peer-memory exchange of the row-partitioned mode (dist.cuh): every rank maps the other ranks' all-reduce vector
// and a small block of flags / scalar slots (CUDA IPC over NVLink / NVSwitch); the collectives of the CG loop are
// then plain kernels that read and write peer memory.
struct PeerSync {
  unsigned flag_in[kMaxWorld], flag_out[kMaxWorld], flag_sc[kMaxWorld];
  unsigned epoch_ar, epoch_sc, done_cnt, pad;
  double grecv[2][kMaxWorld * kMaxRedVals];  // scalar slots, double-buffered by epoch parity
};
struct PeerPtrs {
  double *vec[kMaxWorld];     // base of the all-reduce vector region of every rank ([rank] = local)
  PeerSync *sync[kMaxWorld];
  int world, rank;
};

struct Ctx {
  int device = 0;
  int sms = 148;
  bool dist = false;  // workspace of a row-partitioned solve (dist.cuh)
  int rank = 0, world = 1;
  int n_sh = 0;    // shared (replicated) columns: the prefix [0, n_sh) of every local n-vector
  int cnt_lo = 0;  // reductions over n-space count [cnt_lo, n): 0 on rank 0, n_sh on the others
  bool p2p = false;        // peer-memory collectives set up (dist.cu: dist_p2p_setup); else NCCL
  PeerPtrs peers{};
  void *p2p_arena = nullptr;                 // local allocation holding [PeerSync | vector region]
  void *p2p_opened[2 * kMaxWorld] = {};      // IPC mappings to close
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // side stream of a product that forks (tiled.cuh: the short-row pass runs in the shadow of the streaming
  // kernel); forked and joined with events, so it follows `stream` into a graph capture
  // One side stream per origin stream: a stream that was pulled into a capture stays in it until the capture
  // ends, and the CG loop body is captured on a second origin stream while the main capture is still open.
  static constexpr int kSides = 3;
  cudaStream_t side_origin[kSides] = {nullptr, nullptr, nullptr};
  cudaStream_t side_stream[kSides] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t side_for(cudaStream_t origin) {
    for (int i = 0; i < kSides; ++i)
      if (side_stream[i] && side_origin[i] == origin) return side_stream[i];
    for (int i = 0; i < kSides; ++i)
      if (!side_stream[i]) {
        if (cudaStreamCreateWithFlags(&side_stream[i], cudaStreamNonBlocking) != cudaSuccess) { side_stream[i] = nullptr; return nullptr; }
        side_origin[i] = origin;
        return side_stream[i];
      }
    return nullptr;
  }
  RedWs red{nullptr, nullptr, 0};
  DevScalars *S = nullptr;       // device
  DevScalars *S_host = nullptr;  // pinned mirror
  cudaEvent_t ev = nullptr;
  // counters
  long long launches = 0, spmv_calls = 0, h2d = 0, d2h = 0;
  long long collectives = 0, collective_bytes = 0;  // NCCL all-reduces issued / bytes reduced (dist mode)
  int grid_ew() const { return sms * kCtasPerSm; }

  int init(int dev) {
    device = dev;
    CUDA_OK(cudaSetDevice(device));
    int v = 0;
    CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    sms = v > 0 ? v : 148;
    CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
    red.stride = grid_ew();
    CUDA_OK(cudaMalloc(&red.partials, sizeof(double) * kMaxRedVals * red.stride));
    CUDA_OK(cudaMalloc(&red.ticket, sizeof(unsigned int)));
    CUDA_OK(cudaMemsetAsync(red.ticket, 0, sizeof(unsigned int), stream));
    CUDA_OK(cudaMalloc(&S, sizeof(DevScalars)));
    CUDA_OK(cudaMemsetAsync(S, 0, sizeof(DevScalars), stream));
    CUDA_OK(cudaMallocHost(&S_host, sizeof(DevScalars)));
    memset(S_host, 0, sizeof(DevScalars));
    CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    return 0;
  }
  void destroy() {
    cudaSetDevice(device);
    if (red.partials) cudaFree(red.partials);
    if (red.ticket) cudaFree(red.ticket);
    if (S) cudaFree(S);
    if (S_host) cudaFreeHost(S_host);
    if (ev) cudaEventDestroy(ev);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (int i = 0; i < kSides; ++i) {
      if (side_stream[i]) cudaStreamDestroy(side_stream[i]);
      side_stream[i] = nullptr; side_origin[i] = nullptr;
    }
    ev_fork = ev_join = nullptr;
    if (own_stream && stream) cudaStreamDestroy(stream);
    red = RedWs{nullptr, nullptr, 0};
    S = nullptr; S_host = nullptr; ev = nullptr; stream = nullptr;
  }
  // copy the device scalar block to the pinned mirror and wait for it
  int fetch_scalars() {
    CUDA_OK(cudaMemcpyAsync(S_host, S, sizeof(DevScalars), cudaMemcpyDeviceToHost, stream));
    CUDA_OK(cudaStreamSynchronize(stream));
    d2h += sizeof(DevScalars);
    return 0;
  }
  int sync() {
    CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
  }
};

template <class T>
inline int dev_alloc(T **p, size_t count) {
  *p = nullptr;
  if (count == 0) count = 1;
  CUDA_OK(cudaMalloc((void **)p, count * sizeof(T)));
  return 0;
}
template <class T>
inline int dev_alloc_zero(T **p, size_t count, cudaStream_t st) {
  if (dev_alloc(p, count)) return -1;
  CUDA_OK(cudaMemsetAsync(*p, 0, (count ? count : 1) * sizeof(T), st));
  return 0;
}
template <class T>
inline void dev_free(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}
template <class T>
inline int h2d(Ctx &c, T *dst, const T *src, size_t count) {
  if (!count) return 0;
  CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, c.stream));
  c.h2d += (long long)(count * sizeof(T));
  return 0;
}
template <class T>
inline int d2h(Ctx &c, T *dst, const T *src, size_t count) {
  if (!count) return 0;
  CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, c.stream));
  c.d2h += (long long)(count * sizeof(T));
  return 0;
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ device helpers ----
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void kt_begin(DevScalars *S, int cat) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && S->kt_on) S->kt_start[cat] = gtimer();
}
// by the one thread that knows the grid is done (finaliser of a grid reduction)
__device__ __forceinline__ void kt_end_last(DevScalars *S, int cat) {
  if (S->kt_on) {
    S->kt_ns[cat] += gtimer() - S->kt_start[cat];
    S->kt_cnt[cat] += 1;
  }
}
// by every thread of a kernel without a grid reduction: ticket to find the last CTA
__device__ __forceinline__ void kt_end_ticket(DevScalars *S, int cat) {
  if (!S->kt_on) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&S->kt_ticket[cat], 1u) == gridDim.x - 1) {
      S->kt_ticket[cat] = 0u;
      kt_end_last(S, cat);
    }
  }
}
// Close the running phase interval into phase_ns[slot] and start the next one.  Called by one
// thread of a kernel that sits on a phase boundary of the ADMM iteration.
__device__ __forceinline__ void phase_lap(DevScalars *S, int slot) {
  const unsigned long long now = gtimer();
  if (slot >= 0) S->phase_ns[slot] += now - S->t_mark;
  S->t_mark = now;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// NaN-propagating |.|-max used for inf-norms: fmax() drops NaNs, the reference's loop
// (linalg.c norm_inf) keeps the running max, so NaN is ignored there as well.
__device__ __forceinline__ double absmax(double a, double x) { return fmax(a, fabs(x)); }

// Block-wide reduction of NS sums followed by NM maxes held in vals[0..NS+NM).
// Result valid in thread 0.  sh must hold (NS+NM)*32 doubles.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce(double *vals, double *sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NS + NM; ++k) vals[k] = (k < NS) ? warp_sum(vals[k]) : warp_max(vals[k]);
  __syncthreads();  // protect sh from a previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NS + NM; ++k) sh[k * 32 + wid] = vals[k];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int k = 0; k < NS + NM; ++k) {
      double v = (lane < nw) ? sh[k * 32 + lane] : ((k < NS) ? 0.0 : -INFINITY);
      vals[k] = (k < NS) ? warp_sum(v) : warp_max(v);
    }
  }
}

// Deterministic grid-wide reduction.  Every thread of every block must call it exactly
// once per kernel.  vals[0..NS) are summed, vals[NS..NS+NM) are maxed.  The last block to
// arrive re-reduces the per-block partials in a fixed order and calls fin(vals) on its
// thread 0, then re-arms the ticket.  Requires gridDim.x <= ws.stride.
template <int NS, int NM, class Fin>
__device__ __forceinline__ void grid_reduce(double *vals, const RedWs &ws, Fin fin) {
  __shared__ double sh[(NS + NM) * 32];
  __shared__ int is_last;
  block_reduce<NS, NM>(vals, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NS + NM; ++k) ws.partials[k * ws.stride + blockIdx.x] = vals[k];
    __threadfence();
    unsigned int t = atomicAdd(ws.ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int k = 0; k < NS + NM; ++k) {
    double v = (k < NS) ? 0.0 : -INFINITY;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      double pv = __ldcg(&ws.partials[k * ws.stride + b]);
      v = (k < NS) ? (v + pv) : fmax(v, pv);
    }
    vals[k] = v;
  }
  block_reduce<NS, NM>(vals, sh);
  if (threadIdx.x == 0) {
    fin(vals);
    *ws.ticket = 0u;
    __threadfence();
  }
}
// grid_reduce whose finalising formula `fin` (a functor, constructible on the host as well)
// either runs in place (single GPU) or is deferred until the raw values have been all-reduced
// across the ranks of a row-partitioned solve.
template <int NS, int NM, class Fin>
__device__ __forceinline__ void grid_reduce_fin(double *vals, const RedWs &ws, DevScalars *S, Fin fin) {
  grid_reduce<NS, NM>(vals, ws, [S, fin](double *o) {
    if (S->dist) {
      double *slot = S->gsend + S->dist_rank * (NS + NM);
#pragma unroll
      for (int k = 0; k < NS + NM; ++k) slot[k] = o[k];
    } else {
      fin(o, S);
    }
  });
}
// after the all-reduce gsend -> grecv: combine the ranks' raw values in rank order, run the formula, and
// clear this rank's slot (slots are laid out with the stride of the current reduction, so a stale value
// would land in another rank's slot of a later, wider reduction)
template <class Fin>
__global__ void k_apply_fin(Fin fin, DevScalars *S, int ns, int nm) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const int nv = ns + nm, world = S->dist_world;
  double o[kMaxRedVals];
  for (int k = 0; k < nv; ++k) {
    double v = (k < ns) ? 0.0 : -INFINITY;
    for (int r = 0; r < world; ++r) {
      const double pv = S->grecv[r * nv + k];
      v = (k < ns) ? (v + pv) : fmax(v, pv);
    }
    o[k] = v;
  }
  double *slot = S->gsend + S->dist_rank * nv;
  for (int k = 0; k < nv; ++k) slot[k] = 0.0;
  fin(o, S);
}
#endif  // __CUDACC__

}  // namespace b200
