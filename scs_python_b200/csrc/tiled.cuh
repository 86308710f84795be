// tiled.cuh -- 2-D tiled SpMV engine: both operands of every multiply-add live in shared memory.
//
// Why (profiles/r1b_gather_probe_*.txt, profiles/r1f_tiled_spmv_probe.txt): on a
// uniformly random sparse matrix the row engine of sparse.cuh pays one 32-byte L2 sector for every
// 8-byte gathered operand and sits on the L2 gather ceiling (272 G gathers/s = 2.5 TB/s of
// algorithmic traffic).  Here the matrix is cut into (row bin x column bin) tiles:
//
//   * a CTA owns (a column range of) a row bin of kTR rows: the accumulators y_s[kTR] stay in
//     shared memory;
//   * it walks the column bins that hold non-zeros of that row bin; the kTC-entry slice of the
//     gather vector is staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier,
//     kTStages deep, L2 evict_last), so the vector is read from L2 in coalesced 32 KB pieces
//     instead of sectors;
//   * each of the kTW warps owns kTR / kTW rows of the bin and streams ITS non-zeros -- value
//     (f64) + packed (local row, flag, local column) (u32) = 12 B per stored entry, contiguous per
//     warp, in groups of 32.  The builder (tiled.cu) deals the entries of a (warp, column bin)
//     segment so that the rows inside a group are distinct: the update is a plain shared-memory
//     read-modify-write in a fixed order (deterministic, no atomics).
//
// One product y = M1 x1 + M2 x2 + epilogue is two launches on the workspace stream:
//   1. tiled_kernel over the row bins that are worth an x-slice per column bin (row bins made of short
//      rows only -- identity / bound blocks, a couple of entries per row -- are left to step 2).  The (row bin, column bin) cells of the whole matrix
//      form one sequence that is cut into one contiguous range of equal modelled cost per CTA, so a
//      tall row bin is shared by the two or three CTAs whose ranges meet in it; every piece writes
//      its partial sums to its own slot of an L2-resident scratch;
//   2. tiled_epilogue_kernel<Epi> (high occupancy): per row, the pieces are added in piece order
//      (deterministic) -- short rows are multiplied out right here, one thread per row -- and the sum is
//      handed to the same epilogue functors the row engine uses (scale by R_y^-1, p'Gp, ...), with
//      their usual grid reduction.  Keeping the epilogue out of the streaming kernel costs one
//      extra pass over the raw sums (~2 x 32 MB next to 1.4 GB of matrix) and keeps the streaming
//      loop at a register count that does not spill (an inlined epilogue pushed it into local
//      memory: 2x slower, profiles/r1n_*).
//
// Replaces, for large matrices, the same reference routines as sparse.cuh: SCS(accum_by_a /
// accum_by_atrans / accum_by_p) inside mat_vec (S/linsys/cpu/indirect/private.c:108-121).
#pragma once
#include <cstdint>
#include <vector>

#include "sparse.cuh"

namespace b200 {

// Tile geometry (compile time, -DB200_TILED_CFG=k):
//   0: 16384 rows x 4096 columns, 16 warps (1024 rows each), 3 x-slice stages of 32 KB, 2 x 8 groups in flight
//      per warp  -- round 1
//   1: 8192 rows x 8192 columns, 32 warps (256 rows each), 2 x-slice stages of 64 KB, 2 x 4 groups in flight per
//      warp.  Same tile area (same x-slice traffic per non-zero), same (warp, column bin) segment size (same
//      padding), same bytes in flight per SM, but twice the warps: the per-group chain (shared-memory load ->
//      DFMA -> store, ~45 % of the stall cycles in profiles/r2d_tiled_stream_ncu.txt) is hidden behind eight
//      warps per scheduler instead of four.
// Measured on the bench workload (profiles/r2e_spmv_tile*.txt): Gp product 0.437 ms (cfg 0) -> 0.389 ms (cfg 1),
// A p 0.406 -> 0.366 ms, although cfg 1 stages twice the x-slice bytes (8 kTC per cell): the kernel is bound by
// latency, not by L2 -> SM bytes.
//   4: cfg 0's tile (16384 x 4096) with cfg 1's 32 warps (512 rows each), 2 stages of 32 KB: half the x-slice bytes per
//      cell of cfg 1 at the same shared-memory footprint (192 KB).  Gp product 0.390 -> 0.376 ms, A p unchanged
//      (profiles/r2o_spmv_tile_geometries.txt).  DEFAULT.
//   2: as 4 with 3 stages (224 KB of shared memory): slower than either (0.418 ms) -- nothing is left of the L1 for
//      the entry stream.
// -DB200_TILED_CFG=0 / 1 rebuild the earlier geometries (python -m scs_python_b200.build --alt k).
#ifndef B200_TILED_CFG
#define B200_TILED_CFG 4
#endif
#if B200_TILED_CFG == 1
constexpr int kTR = 8192;              // rows per row bin
constexpr int kTC = 8192;              // columns per column bin
constexpr int kTW = 32;                // warps per CTA
constexpr int kTStages = 2;            // x-slice stages
constexpr int kTU0 = 4, kTU1 = 3;      // groups per register buffer of the two streaming-kernel variants
#elif B200_TILED_CFG == 2 || B200_TILED_CFG == 4  // cfg 0's tile with 32 warps (512 rows each): half the x-slice bytes of cfg 1
constexpr int kTR = 16384;
constexpr int kTC = 4096;
constexpr int kTW = 32;
constexpr int kTStages = B200_TILED_CFG == 2 ? 3 : 2;
constexpr int kTU0 = 4, kTU1 = 3;
#else
constexpr int kTR = 16384;             // rows per row bin
constexpr int kTC = 4096;              // columns per column bin
constexpr int kTW = 16;                // warps per CTA
constexpr int kTStages = 3;            // x-slice stages
constexpr int kTU0 = 8, kTU1 = 6;
#endif
constexpr int kTThreads = kTW * 32;
constexpr int kTRW = kTR / kTW;        // rows one warp owns inside a row bin
constexpr unsigned kTFlag = 1u << 13;  // set in the first group of a (warp, column bin) segment
constexpr int kTRowShift = 14;         // packed entry: local row << 14 | flag << 13 | local column (<= 8192 columns,
                                       // <= 16384 + 32 accumulators incl. the lanes' dummies: 15 + 1 + 13 bits)
constexpr size_t kTSmem = (size_t)(kTR + 32) * 8 + (size_t)kTStages * kTC * 8 + 64;

// one work item: a row bin restricted to a contiguous range of its active column bins
struct TItem {
  int rb;      // row bin
  int a0, a1;  // range in TiledDev::seq (the owning CTA's flat column-bin sequence)
  int slot;    // scratch slot (units of kTR doubles) receiving this piece's partial sums
};

struct TiledDev {
  int nrows = 0, nrb = 0, ncb = 0, ncbA = 0;
  int len1 = 0, len2 = 0;  // lengths of the two gather vectors (column bins [0, ncbA) read the first)
  int nitems = 0, ncta = 0;
  TItem *items = nullptr;      // grouped by CTA
  int *cta_off = nullptr;      // ncta + 1, into items
  int *seq = nullptr;          // per CTA: the column bins its items visit, in visiting order
  int *cta_seq_off = nullptr;  // ncta + 1, into seq
  int *gbase = nullptr;        // nseg + 1 group offsets, segment = (rb * kTW + warp) * ncb + cb
  unsigned *pk = nullptr;      // packed entries, [group][lane]
  double *val = nullptr;       // values, [group][lane]
  double *partial = nullptr;   // scratch: one slot of kTR doubles per piece
  int2 *binfo = nullptr;       // per row bin: {first slot, pieces}; pieces == 0: short-row bin (epilogue pass)
  // short-row bins multiplied out by tiled_direct_kernel on a side stream while the streaming kernel runs:
  // the list of those bins and their raw products (indexed by row); ydir == null: the epilogue pass does it
  int *dbins = nullptr;
  int ndbins = 0;
  double *ydir = nullptr;
  unsigned long long *prof = nullptr;  // per CTA of the last launch: %globaltimer at start, at end, ns spent streaming
};

// host handle
struct TiledOp {
  TiledDev d;
  bool ok = false;
  long long nnz = 0, slots = 0;
  int pieces_max = 1;
  std::vector<double> cta_cost;  // modelled cost of every CTA's item list (host plan)
  std::vector<int> cta_items;
  bool has_tiled = false;  // false: every row bin is a short-row bin (the epilogue pass does it all)
  int variant = 0;         // streaming-kernel variant (tiled_kernel<variant>: groups in flight per warp)
  int epi_eb = 1;          // epilogue pass: 1 / 0 = chunked kernel (plain / batched entry loop); 4 / 8 = row-parallel kernel
  CsrDev m1, m2;
  bool has2 = false;
  // Build from CSR(M1) [and CSR(M2) with the same row count, acting on a second vector]:
  // y = M1 x1 + M2 x2.  force: skip the size / padding heuristics.  Returns < 0 on a CUDA error;
  // ok stays false (and 0 is returned) when the matrix does not suit the format.
  int build(Ctx &c, const CsrDev &m1, const CsrDev *m2, bool force);
  void destroy();
};
// host plan of a tiled operator (tiled.cu: tiled_plan_host), separated from the device work for the CPU tests
struct TiledPlan {
  int ncta = 1, pslots = 0, pieces_max = 1, ndirect = 0;
  std::vector<TItem> items;               // grouped by CTA
  std::vector<int> cta_off, seq, seq_off; // as in TiledDev
  std::vector<int2> binfo;
  std::vector<double> cta_cost;
  std::vector<int> cta_items;
};
void tiled_plan_host(int nrows, int ncb, const int *hg, const int *p1, const int *p2, int sms, bool tiled_only,
                     TiledPlan &P);
// SCS_B200_TILED: "0" never, "1" whenever the structure allows, unset: size / padding heuristic
int tiled_env_mode();

#ifdef __CUDACC__
namespace tl {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
// x-slices are re-read by every row bin while 1.4 GB of matrix stream passes through L2 between two
// uses: mark them evict_last so that the 32 MB gather vector stays resident
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned *p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
  return old;
}
}  // namespace tl

// Streaming kernel: partial sums of every work item into its scratch slot.  kt_cat >= 0: this launch
// opens the in-region timing window of that SpMV category (common.cuh kt_begin).
// kVariant: groups of 32 entries a warp keeps in flight = 2 x U with U = 8 (variant 0) or 6 (variant 1)
template <int kVariant>
__global__ void __launch_bounds__(kTThreads, 1)
tiled_kernel(TiledDev T, const double *__restrict__ x1, const double *x2, DevScalars *S, int kt_cat,
             const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  if (kt_cat >= 0) kt_begin(S, kt_cat);
  extern __shared__ __align__(128) unsigned char tl_smem[];
  double *ys = reinterpret_cast<double *>(tl_smem);
  double *xs = ys + (kTR + 32);
  uint64_t *full = reinterpret_cast<uint64_t *>(xs + (size_t)kTStages * kTC);
  unsigned *cnt = reinterpret_cast<unsigned *>(full + kTStages);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = T.cta_off[blockIdx.x], i1 = T.cta_off[blockIdx.x + 1];
  const int q0 = T.cta_seq_off[blockIdx.x], nq = T.cta_seq_off[blockIdx.x + 1] - q0;
  unsigned long long t_start = 0, t_stream = 0;
  if (threadIdx.x == 0) {
    t_start = gtimer();
    for (int s = 0; s < kTStages; ++s) { tl::mbar_init(full + s, 1); cnt[s] = 0u; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kTR + 32; i += kTThreads) ys[i] = 0.0;
  __syncthreads();
  const uint64_t xpol = tl::l2_evict_last_policy();
  // TMA bulk copy of the q-th x-slice of this CTA's sequence into its stage
  auto issue = [&](int q) {
    const int cb = T.seq[q0 + q];
    const int s = q % kTStages;
    const double *src;
    int left;
    if (cb < T.ncbA) { src = x1 + (size_t)cb * kTC; left = T.len1 - cb * kTC; }
    else { src = x2 + (size_t)(cb - T.ncbA) * kTC; left = T.len2 - (cb - T.ncbA) * kTC; }
    left = left < kTC ? ((left + 1) & ~1) : kTC;  // 16-byte granules; the pad entry is never referenced
    const uint32_t bytes = (uint32_t)left * 8u;
    tl::mbar_expect_tx(full + s, bytes);
    tl::bulk_g2s(xs + (size_t)s * kTC, src, bytes, full + s, xpol);
  };
  if (threadIdx.x == 0)
    for (int q = 0; q < kTStages && q < nq; ++q) issue(q);

  int it = 0;  // x-slices of the CTA's sequence this warp has consumed
  // all lanes: this warp is done with slice q; the last warp to say so refills the stage
  auto release = [&](int q) {
    __syncwarp();
    if (lane == 0) {
      const int s = q % kTStages;
      if (tl::atom_add_acq_rel(cnt + s, 1u) == (unsigned)(kTW - 1)) {
        *reinterpret_cast<volatile unsigned *>(cnt + s) = 0u;
        if (q + kTStages < nq) issue(q + kTStages);
      }
    }
  };
  constexpr int U = kVariant == 0 ? kTU0 : kTU1;
  for (int ii = i0; ii < i1; ++ii) {
    const TItem im = T.items[ii];
    const unsigned long long t_item = threadIdx.x == 0 ? gtimer() : 0ull;
    if (im.a1 > im.a0) {
      const long long segb = ((long long)im.rb * kTW + w) * T.ncb;
      const int gbeg = T.gbase[segb + T.seq[im.a0]];
      const int ng = T.gbase[segb + T.seq[im.a1 - 1] + 1] - gbeg;
      const double *vp = T.val + (size_t)gbeg * 32 + lane;
      const unsigned *pp = T.pk + (size_t)gbeg * 32 + lane;
      double va[U], vb[U];
      unsigned pa[U], pb[U];
      const double *xv = xs;
      bool have = false;
      auto load = [&](double (&vv)[U], unsigned (&pq)[U], int q) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (q + u < ng) { vv[u] = __ldcs(vp + (size_t)(q + u) * 32); pq[u] = __ldcs(pp + (size_t)(q + u) * 32); }
      };
      auto process = [&](const double (&vv)[U], const unsigned (&pq)[U], int q) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (q + u < ng) {  // warp-uniform
            const unsigned pkd = pq[u];
            if (__any_sync(0xffffffffu, pkd & kTFlag)) {  // next column bin of this item
              if (have) release(it - 1);
              have = true;
              const int s = it % kTStages;
              tl::mbar_wait(full + s, (uint32_t)((it / kTStages) & 1));
              xv = xs + (size_t)s * kTC;
              ++it;
            }
            const int r = (int)(pkd >> kTRowShift), c = (int)(pkd & (kTC - 1));
            ys[r] = fma(vv[u], xv[c], ys[r]);
            __syncwarp();
          }
        }
      };
      load(va, pa, 0);
      for (int q = 0; q < ng; q += 2 * U) {
        load(vb, pb, q + U);
        process(va, pa, q);
        load(va, pa, q + 2 * U);
        process(vb, pb, q + U);
      }
      if (have) release(it - 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) t_stream += gtimer() - t_item;
    {  // park the partial sums (all kTR rows of the slot: the epilogue reads only the valid ones)
      double *dst = T.partial + (size_t)im.slot * kTR;
      for (int i = threadIdx.x; i < kTR; i += kTThreads) {
        __stcs(dst + i, ys[i]);
        ys[i] = 0.0;
      }
    }
    __syncthreads();  // y_s is zero again before the next item's updates
  }
  if (threadIdx.x == 0) {
    T.prof[3 * blockIdx.x + 0] = t_start;
    T.prof[3 * blockIdx.x + 1] = gtimer();
    T.prof[3 * blockIdx.x + 2] = t_stream;
  }
}

// Epilogue pass (high occupancy): y_r = sum of the pieces of r's row bin in piece order, or -- for a
// short-row bin -- the row's product straight from the CSR arrays (one thread per row: these rows hold a
// couple of entries); then q = epi.load(r) ; epi.apply(st, r, y_r, q) ; finally epi.finish as in row_kernel.
template <class Epi, int EB>
__global__ void __launch_bounds__(kThreads)
tiled_epilogue_kernel(TiledDev T, CsrDev m1, CsrDev m2, int has2, const double *__restrict__ x1,
                      const double *x2, Epi epi, RedWs ws, DevScalars *S, const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  typename Epi::State st;
  epi.init(st);
  const int stride = gridDim.x * blockDim.x;
  for (int base = blockIdx.x * blockDim.x + threadIdx.x; base < T.nrows; base += stride * EB) {
    typename Epi::Pre pre[EB];
    double acc[EB];
    int2 bi[EB];
    int sA[EB], eA[EB], sB[EB], eB[EB];
#pragma unroll
    for (int k = 0; k < EB; ++k) {
      const int row = base + k * stride;
      sA[k] = eA[k] = sB[k] = eB[k] = 0;
      bi[k] = make_int2(0, 0);
      if (row < T.nrows) {
        bi[k] = T.binfo[row / kTR];
        pre[k] = epi.load(row);
        if (bi[k].y == 0 && T.ydir == nullptr) {
          sA[k] = m1.ptr[row]; eA[k] = m1.ptr[row + 1];
          if (has2) { sB[k] = m2.ptr[row]; eB[k] = m2.ptr[row + 1]; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < EB; ++k) {
      const int row = base + k * stride;
      double a = 0.0;
      if (row < T.nrows) {
        if (bi[k].y == 0) {
          if (T.ydir != nullptr) {
            a = __ldcs(T.ydir + row);  // multiplied out by tiled_direct_kernel while the streaming kernel ran
          } else {
            for (int j = sA[k]; j < eA[k]; ++j) a = fma(__ldcs(m1.val + j), __ldg(x1 + __ldcs(m1.idx + j)), a);
            for (int j = sB[k]; j < eB[k]; ++j) a = fma(__ldcs(m2.val + j), __ldg(x2 + __ldcs(m2.idx + j)), a);
          }
        } else {
          const double *src = T.partial + (size_t)bi[k].x * kTR + (row % kTR);
          for (int p = 0; p < bi[k].y; ++p) a += __ldcs(src + (size_t)p * kTR);
        }
      }
      acc[k] = a;
    }
#pragma unroll
    for (int k = 0; k < EB; ++k) {
      const int row = base + k * stride;
      if (row < T.nrows) epi.apply(st, row, acc[k], pre[k]);
    }
  }
  epi.finish(st, ws, S);
}

// Epilogue pass, chunked form (default).  A CTA takes kEpChunk = 4 x 256 consecutive rows at a time; kTR is a
// multiple of the chunk, so the chunk lies in ONE row bin and the bin's kind is CTA-uniform.  Tiled bins: as
// above (pieces added in piece order).  Short-row bins: the chunk's CSR entries are one contiguous range of
// each source matrix, so the products val * x[idx] are formed ENTRY-parallel -- every thread issues independent,
// coalesced idx / val loads followed by one gather each -- into shared memory, and each row then adds its own
// products in entry order (deterministic).  The row-parallel form above walks ptr -> idx/val -> x row after
// row: nine dependent DRAM latencies per batch of four rows (profiles/r2n_tiled_g_ncu_full.txt: 43 % DRAM, 78 %
// of the cycles no eligible warp); here the chain is ptr -> idx -> x once per chunk.  Measured
// (profiles/r2q_spmv_epilogue_forms.txt, r2p_epilogue_ncu.txt): 67 -> 63 us for the Gp epilogue (two sources),
// 45 -> 52 us for the A p one (single source: the two barriers cost more than the chain), so tiled.cu picks the
// form per operator; batching eight entries per thread ahead of the gathers (kEpUnr = 8) changes nothing -- the
// pass is no longer bound by its dependent chain but by 12 interleaved DRAM streams at ~3.7 TB/s.
constexpr int kEpRows = 4;                      // rows per thread
constexpr int kEpChunk = kThreads * kEpRows;   // rows per CTA step
constexpr int kEpCap = 4096;                    // staged products per chunk (32 KB); larger chunks: row-parallel path
static_assert(kTR % kEpChunk == 0, "a chunk must not straddle row bins");
template <class Epi, int kEpUnr>
__global__ void __launch_bounds__(kThreads, 4)
tiled_epilogue_chunk_kernel(TiledDev T, CsrDev m1, CsrDev m2, int has2, const double *__restrict__ x1,
                            const double *x2, Epi epi, RedWs ws, DevScalars *S, const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  __shared__ double prod[kEpCap];
  typename Epi::State st;
  epi.init(st);
  const int nchunks = (T.nrows + kEpChunk - 1) / kEpChunk;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int r0 = ch * kEpChunk, r1 = min(T.nrows, r0 + kEpChunk);
    const int2 bi = T.binfo[r0 / kTR];
    const bool direct = bi.y == 0 && T.ydir == nullptr;  // CTA-uniform, like everything derived from it
    int a0 = 0, nA = 0, b0 = 0, nB = 0;
    if (direct) {
      a0 = m1.ptr[r0]; nA = m1.ptr[r1] - a0;
      if (has2) { b0 = m2.ptr[r0]; nB = m2.ptr[r1] - b0; }
    }
    const int nT = nA + nB;
    const bool staged = direct && nT <= kEpCap;
    if (staged) {
      // both sources as one index space, kEpUnr entries per thread in flight: all idx / val loads of a batch are
      // issued before the first gather (r2p ncu source page: an un-batched loop serialises idx -> x per entry)
      for (int j0 = threadIdx.x; j0 < nT; j0 += kEpUnr * kThreads) {
        int cc[kEpUnr];
        double vv[kEpUnr];
#pragma unroll
        for (int u = 0; u < kEpUnr; ++u) {
          const int j = j0 + u * kThreads;
          cc[u] = 0; vv[u] = 0.0;
          if (j < nA) { cc[u] = __ldcs(m1.idx + a0 + j); vv[u] = __ldcs(m1.val + a0 + j); }
          else if (j < nT) { cc[u] = __ldcs(m2.idx + b0 + (j - nA)); vv[u] = __ldcs(m2.val + b0 + (j - nA)); }
        }
#pragma unroll
        for (int u = 0; u < kEpUnr; ++u) {
          const int j = j0 + u * kThreads;
          if (j < nA) prod[j] = vv[u] * __ldg(x1 + cc[u]);
          else if (j < nT) prod[j] = vv[u] * __ldg(x2 + cc[u]);
        }
      }
    }
    typename Epi::Pre pre[kEpRows];
    double acc[kEpRows];
#pragma unroll
    for (int k = 0; k < kEpRows; ++k) {
      const int row = r0 + threadIdx.x + k * kThreads;
      acc[k] = 0.0;
      if (row < r1) pre[k] = epi.load(row);
    }
    if (bi.y != 0) {
#pragma unroll
      for (int k = 0; k < kEpRows; ++k) {
        const int row = r0 + threadIdx.x + k * kThreads;
        if (row < r1) {
          const double *src = T.partial + (size_t)bi.x * kTR + (row % kTR);
          double a = 0.0;
          for (int p = 0; p < bi.y; ++p) a += __ldcs(src + (size_t)p * kTR);
          acc[k] = a;
        }
      }
    } else if (!direct) {
#pragma unroll
      for (int k = 0; k < kEpRows; ++k) {
        const int row = r0 + threadIdx.x + k * kThreads;
        if (row < r1) acc[k] = __ldcs(T.ydir + row);
      }
    } else {
      int sA[kEpRows], eA[kEpRows], sB[kEpRows], eB[kEpRows];
#pragma unroll
      for (int k = 0; k < kEpRows; ++k) {
        const int row = r0 + threadIdx.x + k * kThreads;
        sA[k] = eA[k] = sB[k] = eB[k] = 0;
        if (row < r1) {
          sA[k] = m1.ptr[row]; eA[k] = m1.ptr[row + 1];
          if (has2) { sB[k] = m2.ptr[row]; eB[k] = m2.ptr[row + 1]; }
        }
      }
      if (staged) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kEpRows; ++k) {
          double a = 0.0;
          for (int j = sA[k]; j < eA[k]; ++j) a += prod[j - a0];
          for (int j = sB[k]; j < eB[k]; ++j) a += prod[nA + j - b0];
          acc[k] = a;
        }
        __syncthreads();  // prod is free for the next chunk
      } else {
#pragma unroll
        for (int k = 0; k < kEpRows; ++k) {
          double a = 0.0;
          for (int j = sA[k]; j < eA[k]; ++j) a = fma(__ldcs(m1.val + j), __ldg(x1 + __ldcs(m1.idx + j)), a);
          for (int j = sB[k]; j < eB[k]; ++j) a = fma(__ldcs(m2.val + j), __ldg(x2 + __ldcs(m2.idx + j)), a);
          acc[k] = a;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kEpRows; ++k) {
      const int row = r0 + threadIdx.x + k * kThreads;
      if (row < r1) epi.apply(st, row, acc[k], pre[k]);
    }
  }
  epi.finish(st, ws, S);
}

// Short-row bins (identity / bound blocks) multiplied out row by row straight from the CSR arrays: raw
// products into T.ydir.  Launched on the workspace's side stream so that it runs in the shadow of the
// streaming kernel (which leaves 3/4 of every SM's thread slots idle by construction).
template <int kDummy>
__global__ void __launch_bounds__(kThreads, 8)  // <= 32 registers: two of these CTAs fit next to a streaming CTA
tiled_direct_kernel(TiledDev T, CsrDev m1, CsrDev m2, int has2, const double *__restrict__ x1, const double *x2,
                    const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  const long long total = (long long)T.ndbins * kTR;
  constexpr int EB = 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * EB) {
    int row[EB], sA[EB], eA[EB], sB[EB], eB[EB];
#pragma unroll
    for (int k = 0; k < EB; ++k) {
      const long long t = base + k * stride;
      row[k] = -1; sA[k] = eA[k] = sB[k] = eB[k] = 0;
      if (t < total) {
        const int r = T.dbins[t / kTR] * kTR + (int)(t % kTR);
        if (r < T.nrows) {
          row[k] = r;
          sA[k] = m1.ptr[r]; eA[k] = m1.ptr[r + 1];
          if (has2) { sB[k] = m2.ptr[r]; eB[k] = m2.ptr[r + 1]; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < EB; ++k) {
      if (row[k] < 0) continue;
      double a = 0.0;
      for (int j = sA[k]; j < eA[k]; ++j) a = fma(__ldcs(m1.val + j), __ldg(x1 + __ldcs(m1.idx + j)), a);
      for (int j = sB[k]; j < eB[k]; ++j) a = fma(__ldcs(m2.val + j), __ldg(x2 + __ldcs(m2.idx + j)), a);
      __stcs(T.ydir + row[k], a);
    }
  }
}

inline int tiled_prepare() {
  CUDA_OK(cudaFuncSetAttribute(tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem));
  CUDA_OK(cudaFuncSetAttribute(tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem));
  return 0;
}
// kt_cat: timing category whose window the streaming kernel opens (-1: none); the epilogue functor must
// then not open it again (its kt_cont member).  Returns the number of kernels launched.
template <class Epi>
inline int tiled_launch(const TiledOp &op, const double *x1, const double *x2, Epi epi, Ctx &c, const int *skip,
                        int kt_cat = -1) {
  int launched = 1;
  TiledDev T = op.d;
  if (op.has_tiled) {
    cudaStream_t sd = T.ydir != nullptr ? c.side_for(c.stream) : nullptr;
    const bool side = sd != nullptr;
    if (side) cudaEventRecord(c.ev_fork, c.stream);
    if (op.variant == 1) tiled_kernel<1><<<T.ncta, kTThreads, kTSmem, c.stream>>>(T, x1, x2 ? x2 : x1, c.S, kt_cat, skip);
    else tiled_kernel<0><<<T.ncta, kTThreads, kTSmem, c.stream>>>(T, x1, x2 ? x2 : x1, c.S, kt_cat, skip);
    if (side) {  // fork: the short-row bins on the side stream, joined before the epilogue pass
      cudaStreamWaitEvent(sd, c.ev_fork, 0);
      tiled_direct_kernel<0><<<c.sms * 3, kThreads, 0, sd>>>(T, op.m1, op.m2, op.has2 ? 1 : 0, x1, x2 ? x2 : x1, skip);
      cudaEventRecord(c.ev_join, sd);
      cudaStreamWaitEvent(c.stream, c.ev_join, 0);
      ++launched;
    } else {
      T.ydir = nullptr;  // the epilogue pass multiplies the short rows out itself
    }
    ++launched;
    epi.kt_cont = 1;
  } else {
    T.ydir = nullptr;
  }
  const long long blocks = ((long long)op.d.nrows + kThreads - 1) / kThreads;
  const int grid = (int)(blocks < c.grid_ew() ? (blocks > 0 ? blocks : 1) : c.grid_ew());
  if (op.epi_eb == 0) {
    const long long chunks = ((long long)op.d.nrows + kEpChunk - 1) / kEpChunk;
    const int gridc = (int)(chunks < c.grid_ew() ? (chunks > 0 ? chunks : 1) : c.grid_ew());
    tiled_epilogue_chunk_kernel<Epi, 8><<<gridc, kThreads, 0, c.stream>>>(T, op.m1, op.m2, op.has2 ? 1 : 0, x1, x2 ? x2 : x1,
                                                                         epi, c.red, c.S, skip);
  } else if (op.epi_eb == 1) {
    const long long chunks = ((long long)op.d.nrows + kEpChunk - 1) / kEpChunk;
    const int gridc = (int)(chunks < c.grid_ew() ? (chunks > 0 ? chunks : 1) : c.grid_ew());
    tiled_epilogue_chunk_kernel<Epi, 1><<<gridc, kThreads, 0, c.stream>>>(T, op.m1, op.m2, op.has2 ? 1 : 0, x1, x2 ? x2 : x1,
                                                                         epi, c.red, c.S, skip);
  } else if (op.epi_eb == 8)
    tiled_epilogue_kernel<Epi, 8><<<grid, kThreads, 0, c.stream>>>(T, op.m1, op.m2, op.has2 ? 1 : 0, x1, x2 ? x2 : x1, epi,
                                                                   c.red, c.S, skip);
  else
    tiled_epilogue_kernel<Epi, 4><<<grid, kThreads, 0, c.stream>>>(T, op.m1, op.m2, op.has2 ? 1 : 0, x1, x2 ? x2 : x1, epi,
                                                                   c.red, c.S, skip);
  return launched;
}
inline bool tiled_aligned(const double *x) { return (reinterpret_cast<uintptr_t>(x) & 15u) == 0; }
#endif  // __CUDACC__

}  // namespace b200
