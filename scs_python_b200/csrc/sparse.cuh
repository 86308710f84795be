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

// One unit of work of the row engine: rows [row_begin, row_end) whose non-zeros (of A, and of
// B for the fused two-matrix kernels) fit one shared-memory tile.  The nnz offsets are kept in
// the descriptor so the coalesced val/idx streams can be issued straight after the
// descriptor load, without first chasing the row pointers.
struct __align__(16) Chunk {
  int row_begin, row_end;
  int lg;      // log2(lanes per row) for the tile reduction (0..5), or kLongRow
  int pad;
  int a0, na;  // first non-zero / count inside A for these rows
  int b0, nb;  // same for B (0, 0 when the list was built for one matrix)
};
constexpr int kLongRow = 8;     // Chunk::lg marker: a single row longer than a tile
constexpr int kTile = 2048;     // non-zeros staged per chunk (16 KB of FP64 products)
constexpr int kIpt = kTile / kThreads;
constexpr int kMaxRows = 1024;  // rows per chunk

struct ChunkList {
  int n = 0;
  Chunk *d = nullptr;
  int grid = 1;
};

// Build the chunk list on the host from one or two row-pointer arrays (second may be null;
// used to fuse P's rows with A' rows).  Greedy: a chunk is closed when the next row would
// overflow the tile, when it holds kMaxRows rows, or when the row-length class drifts by
// more than 2x (so one lanes-per-row setting fits all its rows).
inline void build_chunks_host(int nrows, const int *ptr1, const int *ptr2, std::vector<Chunk> &out) {
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
  auto emit = [&](int begin, int end, int lg) {
    Chunk ch;
    ch.row_begin = begin; ch.row_end = end; ch.lg = lg; ch.pad = 0;
    ch.a0 = ptr1[begin]; ch.na = ptr1[end] - ptr1[begin];
    ch.b0 = ptr2 ? ptr2[begin] : 0; ch.nb = ptr2 ? ptr2[end] - ptr2[begin] : 0;
    out.push_back(ch);
  };
  while (r < nrows) {
    const long long L0 = rowlen(r);
    if (L0 > kTile) {  // very long row: whole CTA loops over it
      emit(r, r + 1, kLongRow);
      ++r;
      continue;
    }
    const int begin = r;
    long long nnz = 0;
    const int c0 = cls(L0 > 0 ? L0 : 1);
    int cmin = c0, cmax = c0;
    while (r < nrows && r - begin < kMaxRows) {
      const long long L = rowlen(r);
      if (nnz + L > kTile) break;
      const int c = cls(L > 0 ? L : 1);
      const int nmin = c < cmin ? c : cmin, nmax = c > cmax ? c : cmax;
      if (r > begin && (nmax - nmin > 1) && (r - begin) >= 32) break;
      cmin = nmin; cmax = nmax;
      nnz += L;
      ++r;
    }
    const int rows = r - begin;
    const long long avg = rows > 0 ? (nnz + rows - 1) / rows : 1;
    int lg = 0;
    while ((2ll << lg) <= avg && lg < 5) ++lg;  // lanes = pow2_floor(avg) capped at 32
    emit(begin, r, lg);
  }
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// element functors: the value one stored entry contributes (map) and how contributions of
// one row combine (comb); init() is comb's neutral element
// ---------------------------------------------------------------------------------------
struct ElemMul {  // sum_k val * x[col]              (SpMV)
  const double *__restrict__ x;
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double map(double v, int c) const { return v * __ldg(x + c); }
  __device__ __forceinline__ double comb(double a, double b) const { return a + b; }
};
struct ElemAbsMax {  // max_k |val|                  (Ruiz row/col norms, scs_matrix.c:224-228,267-273)
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double map(double v, int) const { return fabs(v); }
  __device__ __forceinline__ double comb(double a, double b) const { return fmax(a, b); }
};
struct ElemSumSq {  // sum_k val^2                   (L2 pass, scs_matrix.c:293-297,337-339)
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double map(double v, int) const { return v * v; }
  __device__ __forceinline__ double comb(double a, double b) const { return a + b; }
};
struct ElemSqDiv {  // sum_k val^2 / d[col]          (preconditioner, cpu/indirect/private.c:65-68)
  const double *__restrict__ d;
  __device__ __forceinline__ double init() const { return 0.0; }
  __device__ __forceinline__ double map(double v, int c) const { return v * v / __ldg(d + c); }
  __device__ __forceinline__ double comb(double a, double b) const { return a + b; }
};

// Stage the mapped entries [k0, k0+cnt) of M into the shared tile: every load of the
// (val, idx) streams is independent and coalesced, kIpt of each in flight per thread.
template <class Elem>
__device__ __forceinline__ void stage_tile(const CsrDev &M, const Elem &e, int k0, int cnt, double *tile) {
  const double *__restrict__ val = M.val + k0;
  const int *__restrict__ idx = M.idx + k0;
  double v[kIpt];
  int c[kIpt];
#pragma unroll
  for (int i = 0; i < kIpt; ++i) {
    const int k = threadIdx.x + i * kThreads;
    if (k < cnt) {
      v[i] = __ldcs(val + k);
      c[i] = __ldcs(idx + k);
    }
  }
#pragma unroll
  for (int i = 0; i < kIpt; ++i) {
    const int k = threadIdx.x + i * kThreads;
    if (k < cnt) tile[k] = e.map(v[i], c[i]);
  }
}

// whole-CTA reduction of one long row; result valid in thread 0
template <class Elem>
__device__ __forceinline__ double long_row(const CsrDev &M, const Elem &e, int k0, int cnt, double *sh8) {
  double acc = e.init();
  for (int k = threadIdx.x; k < cnt; k += kThreads) acc = e.comb(acc, e.map(__ldcs(M.val + k0 + k), __ldcs(M.idx + k0 + k)));
  for (int o = 16; o > 0; o >>= 1) acc = e.comb(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = acc;
  __syncthreads();
  double t = sh8[0];
  for (int w = 1; w < kThreads / 32; ++w) t = e.comb(t, sh8[w]);
  return t;
}

// Generic row engine ("CSR-stream").  For every row r: acc = reduce_k Elem(A[r,k]) (and the
// same over B's row r with Elem2 when DUAL), then epi.row(st, r, acc) -- or
// epi.row2(st, r, accA, accB) when Epi::kSeparate -- with consecutive threads on
// consecutive rows.  After the sweep every thread calls epi.finish(st, ws, S) (grid
// reduction / scalar finalisation).  A persistent grid walks the chunk list round-robin.
// skip: optional device flag; when set the kernel exits at once (CG already converged).
template <class Elem, class Elem2, class Epi, bool DUAL>
__global__ void __launch_bounds__(kThreads, 4)
row_kernel(CsrDev A, Elem ea, CsrDev B, Elem2 eb, const Chunk *__restrict__ chunks, int nchunks, Epi epi,
           RedWs ws, DevScalars *S, const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  constexpr bool SEP = DUAL && Epi::kSeparate;
  __shared__ double tile[kTile];
  __shared__ int sptr[(DUAL ? 2 : 1) * (kMaxRows + 1)];
  __shared__ double racc[(SEP ? 2 : 1) * kMaxRows];
  __shared__ double sh8[kThreads / 32];
  typename Epi::State st;
  epi.init(st);
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int4 c0 = __ldg(reinterpret_cast<const int4 *>(chunks + c));
    const int4 c1 = __ldg(reinterpret_cast<const int4 *>(chunks + c) + 1);
    const int r0 = c0.x, nrows = c0.y - c0.x, lg = c0.z;
    const int a0 = c1.x, na = c1.y, b0 = c1.z, nb = c1.w;
    if (lg == kLongRow) {
      double t = long_row(A, ea, a0, na, sh8);
      double t2 = eb.init();
      if (DUAL) t2 = long_row(B, eb, b0, nb, sh8);
      if (threadIdx.x == 0) {
        if (SEP) epi.row2(st, r0, t, t2);
        else epi.row(st, r0, DUAL ? ea.comb(t, t2) : t);
      }
      continue;
    }
    // ---- phase 1: stream the chunk's entries, gather, stage products; row pointers too
    for (int i = threadIdx.x; i <= nrows; i += kThreads) {
      sptr[i] = __ldg(A.ptr + r0 + i) - a0;
      if (DUAL) sptr[kMaxRows + 1 + i] = __ldg(B.ptr + r0 + i) - b0 + na;
    }
    stage_tile(A, ea, a0, na, tile);
    if (DUAL) stage_tile(B, eb, b0, nb, tile + na);
    __syncthreads();
    // ---- phase 2: segmented reduction, G lanes per row
    if (lg == 0) {
      for (int i = threadIdx.x; i < nrows; i += kThreads) {
        double acc = ea.init();
        for (int k = sptr[i]; k < sptr[i + 1]; ++k) acc = ea.comb(acc, tile[k]);
        if (DUAL) {
          double acc2 = eb.init();
          for (int k = sptr[kMaxRows + 1 + i]; k < sptr[kMaxRows + 2 + i]; ++k) acc2 = eb.comb(acc2, tile[k]);
          if (SEP) epi.row2(st, r0 + i, acc, acc2);
          else epi.row(st, r0 + i, ea.comb(acc, acc2));
        } else {
          epi.row(st, r0 + i, acc);
        }
      }
    } else {
      const int G = 1 << lg;
      const int lane = threadIdx.x & (G - 1);
      const int grp = threadIdx.x >> lg;
      const int ngrp = kThreads >> lg;
      for (int base = 0; base < nrows; base += ngrp) {
        const int i = base + grp;
        const bool valid = i < nrows;
        double acc = ea.init(), acc2 = eb.init();
        if (valid) {
          for (int k = sptr[i] + lane; k < sptr[i + 1]; k += G) acc = ea.comb(acc, tile[k]);
          if (DUAL)
            for (int k = sptr[kMaxRows + 1 + i] + lane; k < sptr[kMaxRows + 2 + i]; k += G) acc2 = eb.comb(acc2, tile[k]);
        }
        if (DUAL && !SEP) acc = ea.comb(acc, acc2);
        for (int o = G >> 1; o > 0; o >>= 1) {
          acc = ea.comb(acc, __shfl_xor_sync(0xffffffffu, acc, o));
          if (SEP) acc2 = eb.comb(acc2, __shfl_xor_sync(0xffffffffu, acc2, o));
        }
        if (valid && lane == 0) {
          racc[i] = acc;
          if (SEP) racc[kMaxRows + i] = acc2;
        }
      }
      __syncthreads();
      // ---- phase 3: epilogue, consecutive threads on consecutive rows
      for (int i = threadIdx.x; i < nrows; i += kThreads) {
        if (SEP) epi.row2(st, r0 + i, racc[i], racc[kMaxRows + i]);
        else epi.row(st, r0 + i, racc[i]);
      }
    }
    __syncthreads();  // tile / sptr / racc are reused by the next chunk
  }
  epi.finish(st, ws, S);
}

// Row map: val[k] = f(row, col, val[k]) -- used by the equilibration rescale.
template <class F>
__global__ void __launch_bounds__(kThreads)
row_map_kernel(CsrDev A, const Chunk *__restrict__ chunks, int nchunks, F f) {
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const Chunk ch = chunks[c];
    const int lg = ch.lg <= 5 ? ch.lg : 8;
    const int G = 1 << lg;
    const int lane = threadIdx.x & (G - 1);
    const int grp = threadIdx.x >> lg;
    const int ngrp = kThreads >> lg;
    for (int row = ch.row_begin + grp; row < ch.row_end; row += ngrp) {
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
