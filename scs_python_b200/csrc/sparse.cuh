// sparse.cuh -- device CSR storage and the row-parallel SpMV engine.
//
// Reference behaviour replaced: SCS(accum_by_a / accum_by_atrans / accum_by_p)
// (S/linsys/scs_matrix.c:135-199) and the CSC(A') copy of the CPU indirect backend
// (S/linsys/cpu/indirect/private.c:7-46).  The reference scatters (y[i] += ...) over CSC
// columns; here BOTH orientations are kept in HBM as CSR -- CSR(A) (m rows) and
// CSR(A') (n rows, which is the user's CSC(A) verbatim) -- plus a full symmetric CSR(P),
// so every product is a gather with no atomics and a fixed summation order.
//
// Kernel shape ("chunked adaptive CSR-vector"): rows are cut on the host, once, into
// chunks of ~equal non-zero count; every chunk carries its own lanes-per-row (1..32, or the
// whole CTA for very long rows) chosen from its mean row length, so a matrix that mixes
// 1-2 nnz rows (identity blocks of LASSO/MPC) with 50-100 nnz rows keeps coalesced
// val/idx streams in both regimes.  A persistent grid (SMs x 8 CTAs) walks the chunk list
// round-robin.  The per-row result is handed to an epilogue functor (scale by R_y^-1, add
// R_x p, dot products, inf-norms ...) so the CG / residual vector work rides on the SpMV
// pass instead of re-reading the vectors.
#pragma once
#include "common.cuh"

namespace b200 {

struct CsrDev {
  int nrows = 0, ncols = 0;
  long long nnz = 0;
  int *ptr = nullptr;     // nrows + 1
  int *idx = nullptr;     // nnz
  double *val = nullptr;  // nnz
};

// row_begin, row_end, log2(lanes per row) (8 == whole CTA), unused
struct ChunkList {
  int n = 0;
  int4 *d = nullptr;
  int grid = 1;
};

// Build the chunk list on the host from one or two row-pointer arrays (second may be null;
// used to fuse P's rows with A' rows).  Greedy: close a chunk when it holds >= target nnz,
// >= max_rows rows, or the row-length class changes by more than 2x.
// sms: SM count of the device.  A matrix with few non-zeros is cut into smaller chunks so that the kernel still
// spreads over the whole device: with 4096-entry chunks the n = 2000 cone QP of BASELINE.json configs[0] (60 k
// non-zeros) ran on 15 CTAs and every SpMV launch took ~10 us of serial dependent loads
// (profiles/r2g_configs_unroll*.jsonl); chunk size = total / (4 x SMs), clamped to [512, 4096].
inline void build_chunks_host(int nrows, const int *ptr1, const int *ptr2, std::vector<int4> &out, int sms = 148) {
  long long total_nnz = (long long)ptr1[nrows] + (ptr2 ? (long long)ptr2[nrows] : 0);
  long long target = total_nnz / (4ll * (sms > 0 ? sms : 148));
  target = target < 512 ? 512 : (target > 4096 ? 4096 : target);
  const int max_rows = target >= 4096 ? 2048 : 256;
  const long long long_row = 16384;
  out.clear();
  int r = 0;
  auto rowlen = [&](int i) -> long long {
    long long L = (long long)ptr1[i + 1] - ptr1[i];
    if (ptr2) L += (long long)ptr2[i + 1] - ptr2[i];
    return L;
  };
  auto cls = [](long long L) {
    int c = 0;
    while ((1ll << (c + 1)) <= L) ++c;
    return c;  // floor(log2(max(L,1)))
  };
  while (r < nrows) {
    long long L0 = rowlen(r);
    if (L0 >= long_row) {  // very long row: whole CTA
      out.push_back(make_int4(r, r + 1, 8, 0));
      ++r;
      continue;
    }
    int begin = r;
    long long nnz = 0;
    int c0 = cls(L0 > 0 ? L0 : 1);
    int cmin = c0, cmax = c0;
    while (r < nrows) {
      long long L = rowlen(r);
      if (L >= long_row) break;
      int c = cls(L > 0 ? L : 1);
      int nmin = c < cmin ? c : cmin, nmax = c > cmax ? c : cmax;
      if (r > begin && (nmax - nmin > 1) && (r - begin) >= 32) break;
      cmin = nmin; cmax = nmax;
      nnz += L;
      ++r;
      if (nnz >= target || r - begin >= max_rows) break;
    }
    int rows = r - begin;
    long long avg = rows > 0 ? (nnz + rows - 1) / rows : 1;
    int lg = 0;
    while ((2ll << lg) <= avg && lg < 5) ++lg;  // lanes = pow2_floor(avg) capped at 32
    out.push_back(make_int4(begin, r, lg, 0));
  }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// element functors: how one stored entry contributes to its row's accumulator
// ---------------------------------------------------------------------------------------
struct ElemMul {  // acc += val * x[col]            (SpMV)
  const double *__restrict__ x;
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double term(double a, double v, int c) const { return fma(v, __ldg(x + c), a); }
  __device__ __forceinline__ double comb(double a, double b) const { return a + b; }
};
struct ElemAbsMax {  // acc = max(acc, |val|)       (Ruiz row/col norms, scs_matrix.c:224-228,267-273)
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double term(double a, double v, int) const { return fmax(a, fabs(v)); }
  __device__ __forceinline__ double comb(double a, double b) const { return fmax(a, b); }
};
struct ElemSumSq {  // acc += val^2                 (L2 pass, scs_matrix.c:293-297,337-339)
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double term(double a, double v, int) const { return fma(v, v, a); }
  __device__ __forceinline__ double comb(double a, double b) const { return a + b; }
};
struct ElemSqDiv {  // acc += val^2 / d[col]        (preconditioner, cpu/indirect/private.c:65-68)
  const double *__restrict__ d;
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double term(double a, double v, int c) const { return a + v * v / __ldg(d + c); }
  __device__ __forceinline__ double comb(double a, double b) const { return a + b; }
};

template <class Elem>
__device__ __forceinline__ double row_partial(const CsrDev &M, const Elem &e, int row, int lane, int G,
                                              double acc) {
  const int start = M.ptr[row], end = M.ptr[row + 1];
  int k = start + lane;
  // two independent loads in flight per lane
  for (; k + G < end; k += 2 * G) {
    const double v0 = __ldcs(M.val + k), v1 = __ldcs(M.val + k + G);
    const int c0 = __ldcs(M.idx + k), c1 = __ldcs(M.idx + k + G);
    acc = e.term(acc, v0, c0);
    acc = e.term(acc, v1, c1);
  }
  if (k < end) acc = e.term(acc, __ldcs(M.val + k), __ldcs(M.idx + k));
  return acc;
}

// Generic row engine.  For every row r: acc = reduce_k Elem(A[r,k]) (+ reduce over B's
// row r with Elem2 when DUAL), then epi.row(st, r, acc) on one lane.  After the sweep every
// thread calls epi.finish(st, ws, S) (grid reduction / scalar finalisation).
// skip: optional device flag; when set the kernel exits at once (CG already converged).
template <class Elem, class Elem2, class Epi, bool DUAL>
__global__ void __launch_bounds__(kThreads)
row_kernel(CsrDev A, Elem ea, CsrDev B, Elem2 eb, const int4 *__restrict__ chunks, int nchunks, Epi epi,
           RedWs ws, DevScalars *S, const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  __shared__ double shrow[32];
  typename Epi::State st;
  epi.init(st);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int4 ch = chunks[c];
    if (ch.z <= 5) {
      const int G = 1 << ch.z;
      const int lane = threadIdx.x & (G - 1);
      const int grp = threadIdx.x >> ch.z;
      const int ngrp = kThreads >> ch.z;
      for (int base = ch.x; base < ch.y; base += ngrp) {
        const int row = base + grp;
        const bool valid = row < ch.y;
        double acc = ea.init(), acc2 = eb.init();
        if (valid) {
          acc = row_partial(A, ea, row, lane, G, acc);
          if (DUAL) acc2 = row_partial(B, eb, row, lane, G, acc2);
        }
        if (DUAL && Epi::kSeparate) {
          for (int o = G >> 1; o > 0; o >>= 1) {
            acc = ea.comb(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            acc2 = eb.comb(acc2, __shfl_xor_sync(0xffffffffu, acc2, o));
          }
          if (valid && lane == 0) epi.row2(st, row, acc, acc2);
        } else {
          if (DUAL) acc = ea.comb(acc, acc2);
          for (int o = G >> 1; o > 0; o >>= 1) acc = ea.comb(acc, __shfl_xor_sync(0xffffffffu, acc, o));
          if (valid && lane == 0) epi.row(st, row, acc);
        }
      }
    } else {  // whole CTA per row
      for (int row = ch.x; row < ch.y; ++row) {
        double acc = row_partial(A, ea, row, threadIdx.x, kThreads, ea.init());
        double acc2 = eb.init();
        if (DUAL) acc2 = row_partial(B, eb, row, threadIdx.x, kThreads, acc2);
        for (int o = 16; o > 0; o >>= 1) {
          acc = ea.comb(acc, __shfl_xor_sync(0xffffffffu, acc, o));
          acc2 = eb.comb(acc2, __shfl_xor_sync(0xffffffffu, acc2, o));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { shrow[threadIdx.x >> 5] = acc; shrow[8 + (threadIdx.x >> 5)] = acc2; }
        __syncthreads();
        if (threadIdx.x == 0) {
          double t = shrow[0], t2 = shrow[8];
          for (int w = 1; w < kThreads / 32; ++w) { t = ea.comb(t, shrow[w]); t2 = eb.comb(t2, shrow[8 + w]); }
          if (DUAL && Epi::kSeparate) epi.row2(st, row, t, t2);
          else epi.row(st, row, DUAL ? ea.comb(t, t2) : t);
        }
      }
    }
  }
  epi.finish(st, ws, S);
}

// Row map: val[k] = f(row, col, val[k]) -- used by the equilibration rescale.
template <class F>
__global__ void __launch_bounds__(kThreads)
row_map_kernel(CsrDev A, const int4 *__restrict__ chunks, int nchunks, F f) {
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int4 ch = chunks[c];
    const int lg = ch.z <= 5 ? ch.z : 8;
    const int G = 1 << lg;
    const int lane = threadIdx.x & (G - 1);
    const int grp = threadIdx.x >> lg;
    const int ngrp = kThreads >> lg;
    for (int row = ch.x + grp; row < ch.y; row += ngrp) {
      const int start = A.ptr[row], end = A.ptr[row + 1];
      for (int k = start + lane; k < end; k += G) A.val[k] = f(row, A.idx[k], A.val[k]);
    }
  }
}

// ---- epilogues that are generic enough to live here ----
struct EpiNoState {
  static constexpr bool kSeparate = false;  // DUAL kernels: hand (accA, accB) to row2() instead of their sum
  struct State {};
  __device__ __forceinline__ void row2(State &, int, double, double) const {}
  __device__ __forceinline__ void init(State &) const {}
  __device__ __forceinline__ void finish(State &, const RedWs &, DevScalars *) const {}
};
// y[row] += acc          (test surface: SCS(accum_by_*))
struct EpiAccum : EpiNoState {
  double *y;
  __device__ __forceinline__ void row(State &, int r, double acc) const { y[r] += acc; }
};
// y[row] = acc
struct EpiStore : EpiNoState {
  double *y;
  __device__ __forceinline__ void row(State &, int r, double acc) const { y[r] = acc; }
};
#endif  // __CUDACC__

// ---- host API implemented in sparse.cu ----
// Upload a host CSC matrix as CSR of its transpose (verbatim copy).
int csr_upload(Ctx &c, CsrDev &out, int nrows, int ncols, const int *ptr, const int *idx, const double *val);
// Device transpose: given CSR(T) of shape (nrows x ncols) build CSR(T') (ncols x nrows);
// entries of every output row keep ascending column order.  perm (optional, nnz ints,
// device, may be null) receives source positions.
int csr_transpose(Ctx &c, const CsrDev &in, CsrDev &out);
// Expand upper-triangular CSC P (host) to the full symmetric CSR on the device.
int csr_from_upper_csc(Ctx &c, CsrDev &out, int n, const int *Pp, const int *Pi, const double *Px);
void csr_free(CsrDev &m);
int chunks_build(Ctx &c, ChunkList &out, const CsrDev &a, const CsrDev *b);
void chunks_free(ChunkList &cl);

}  // namespace b200
