// tiled_spmv_probe.cu -- can a 2-D tiled SpMV with BOTH operands in shared memory beat the L2 gather
// ceiling (272 G gathers/s, profiles/r1b_gather_probe_*.txt) on a uniformly random sparse matrix?
// (dev probe, not product)
//
//   CTA (persistent) owns a row bin of R rows: accumulators y_s[R] live in shared memory.
//   It walks the column bins (C columns each); the x-slice of a bin is staged into shared memory by a
//   producer warp with TMA bulk copies (cp.async.bulk + mbarrier, NST stages).
//   Each of the W consumer warps owns R/W rows of the bin and streams ITS non-zeros of the row bin
//   (val f64 + packed (row_local << 18 | col_local) u32 = 12 B per non-zero, contiguous per warp) in
//   groups of 32; the format builder guarantees distinct rows inside a group, so the update is a plain
//   shared-memory read-modify-write (deterministic order, no atomics).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tiled_probe tools/probes/tiled_spmv_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned hash32(unsigned a, unsigned b) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  h ^= h >> 15; h *= 0xC2B2AE3Du; h ^= h >> 13; h *= 0x27D4EB2Fu; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// layout of the stream: [row bin][warp][col bin][group][lane]
__global__ void fill(unsigned *pk, double *val, long long ngroups, int R, int C, int W, int ncb, int G) {
  const int rows_per_warp = R / W;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < ngroups * 32; k += (long long)gridDim.x * blockDim.x) {
    const long long grp = k >> 5;
    const int lane = (int)(k & 31);
    const int w = (int)((grp / ((long long)ncb * G)) % W);
    // distinct rows inside a group: lane-th residue class of the warp's row range
    const unsigned h = hash32((unsigned)grp, (unsigned)lane);
    const int per_lane = rows_per_warp / 32;
    const int r = w * rows_per_warp + lane * per_lane + (int)(h % (unsigned)per_lane);
    const int c = (int)(hash32(h, 77u) % (unsigned)C);
    pk[k] = ((unsigned)r << 18) | (unsigned)c;
    val[k] = 1.0 + (k & 7) * 0.125;
  }
}
__global__ void fillx(double *x, long long n) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) x[k] = 1.0 + (k % 13) * 0.01;
}

// MODE 0: full (x gather + y RMW + __syncwarp), 1: no __syncwarp, 2: x gather only (register acc), 3: y RMW only
template <int MODE, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_tiled(const double *__restrict__ val, const unsigned *__restrict__ pk,
                                                  const double *__restrict__ x, double *__restrict__ y,
                                                  int R, int C, int NST, int W, int ncb, int nrb, int G) {
  extern __shared__ __align__(128) unsigned char smem[];
  double *ys = (double *)smem;
  double *xs = ys + R;
  uint64_t *full = (uint64_t *)(xs + (size_t)NST * C);
  uint64_t *empty = full + NST;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nct = W * 32;  // consumer threads
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, W); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < R; i += blockDim.x) ys[i] = 0.0;
  __syncthreads();
  if (w == W) {  // producer warp
    if (lane == 0) {
      long long it = 0;
      for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x)
        for (int cb = 0; cb < ncb; ++cb, ++it) {
          const int s = (int)(it % NST);
          const long long k = it / NST;
          if (k > 0) mbar_wait(empty + s, (uint32_t)((k - 1) & 1));
          mbar_expect_tx(full + s, (uint32_t)(C * 8));
          bulk_g2s(xs + (size_t)s * C, x + (size_t)cb * C, (uint32_t)(C * 8), full + s);
        }
    }
    return;
  }
  long long it = 0;
  double sink = 0.0;
  for (int rb = blockIdx.x; rb < nrb; rb += gridDim.x) {
    const long long gbase = ((long long)rb * W + w) * (long long)ncb * G;  // first group of this warp's stream
    const int ng = ncb * G;
    const double *vp = val + gbase * 32 + lane;
    const unsigned *pp = pk + gbase * 32 + lane;
    double va[U], vb[U];
    unsigned pa[U], pb[U];
    const double *xv = xs;
    int next_boundary = 0;
    bool first = true;
    auto load = [&](double (&vv)[U], unsigned (&pq)[U], int q) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (q + u < ng) { vv[u] = __ldcs(vp + (size_t)(q + u) * 32); pq[u] = __ldcs(pp + (size_t)(q + u) * 32); }
    };
    auto process = [&](const double (&vv)[U], const unsigned (&pq)[U], int q) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (q + u < ng) {
          if (q + u == next_boundary) {  // warp-uniform: next column bin
            if (!first) { __syncwarp(); if (lane == 0) mbar_arrive(empty + (int)((it - 1) % NST)); }
            first = false;
            const int s = (int)(it % NST);
            mbar_wait(full + s, (uint32_t)((it / NST) & 1));
            xv = xs + (size_t)s * C;
            ++it;
            next_boundary += G;
          }
          const unsigned pkd = pq[u];
          const int r = (int)(pkd >> 18), c = (int)(pkd & 0x3ffffu);
          if (MODE == 2) { sink = fma(vv[u], xv[c], sink); }
          else if (MODE == 3) { ys[r] = fma(vv[u], 1.5, ys[r]); }
          else { ys[r] = fma(vv[u], xv[c], ys[r]); if (MODE == 0) __syncwarp(); }
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
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + (int)((it - 1) % NST));
    consumer_sync(nct);
    for (int i = threadIdx.x; i < R; i += nct) { y[(size_t)rb * R + i] = ys[i] + (MODE == 2 ? sink : 0.0); ys[i] = 0.0; }
    consumer_sync(nct);
  }
}

template <class L>
float timeit(L launch, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); launch();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms / reps;
}

struct Cfg { int R, C, NST, W; };

int main(int argc, char **argv) {
  const double dens = argc > 1 ? atof(argv[1]) : 5.1e-5;
  const long long ncols_target = argc > 2 ? atoll(argv[2]) : 1000000;
  const long long nrows_target = argc > 3 ? atoll(argv[3]) : 2000000;
  int dev = 0; CK(cudaSetDevice(dev));
  int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const Cfg cfgs[] = {{8192, 8192, 2, 16},  {16384, 4096, 3, 16}, {16384, 4096, 3, 32}, {20480, 2048, 4, 16},
                      {20480, 2048, 4, 32}, {24576, 1024, 4, 16}, {12288, 4096, 4, 16}, {16384, 2048, 6, 16},
                      {16384, 2048, 6, 8},  {8192, 4096, 5, 16}};
  for (const Cfg &cf : cfgs) {
    const int R = cf.R, C = cf.C, NST = cf.NST, W = cf.W;
    const int ncb = (int)((ncols_target + C - 1) / C);
    int nrb = (int)((nrows_target + R - 1) / R);
    nrb = ((nrb + nsm - 1) / nsm) * nsm;  // whole waves (the real builder balances bins by nnz instead)
    int G = (int)(R * (double)C * dens / (W * 32.0) + 0.5);
    if (G < 1) G = 1;
    const long long ngroups = (long long)nrb * W * ncb * G;
    const long long nnz = ngroups * 32;
    const size_t smem = (size_t)R * 8 + (size_t)NST * C * 8 + 2 * NST * 8;
    if (smem > 227 * 1024) { printf("cfg R%d C%d NST%d: smem %zu too large\n", R, C, NST, smem); continue; }
    unsigned *pk; double *val, *x, *y;
    CK(cudaMalloc(&pk, nnz * 4)); CK(cudaMalloc(&val, nnz * 8));
    CK(cudaMalloc(&x, (size_t)ncb * C * 8)); CK(cudaMalloc(&y, (size_t)nrb * R * 8));
    fill<<<nsm * 8, 256>>>(pk, val, ngroups, R, C, W, ncb, G);
    fillx<<<nsm * 8, 256>>>(x, (long long)ncb * C);
    CK(cudaDeviceSynchronize());
    printf("R %d C %d NST %d W %d | rows %lld cols %lld nnz %lld (%.1f/row, G=%d) smem %zu B | x-slice L2 traffic %.1f B/nnz\n", R, C, NST, W,
           (long long)nrb * R, (long long)ncb * C, nnz, (double)nnz / ((double)nrb * R), G, smem, (double)C * 8 / (W * G * 32.0));
    const int threads = W * 32 + 32;
#define RUN1(MODE, U, MAXT) { \
      CK(cudaFuncSetAttribute(k_tiled<MODE, U, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      float ms = timeit([&] { k_tiled<MODE, U, MAXT><<<nsm, threads, smem>>>(val, pk, x, y, R, C, NST, W, ncb, nrb, G); }, 10); \
      printf("   mode %d U %d : %.3f ms  %.1f Gnnz/s  %.0f GB/s (12 B/nnz)\n", MODE, U, ms, nnz / ms / 1e6, 12.0 * nnz / ms / 1e6); }
#define RUN(MODE, U) { if (W <= 16) RUN1(MODE, U, 544) else RUN1(MODE, U, 1056) }
    RUN(0, 2) RUN(0, 4) RUN(1, 4) RUN(1, 8) RUN(2, 4) RUN(3, 4)
    double h[2]; CK(cudaMemcpy(h, y, 16, cudaMemcpyDeviceToHost));
    printf("   y[0..1] = %f %f\n", h[0], h[1]);
    cudaFree(pk); cudaFree(val); cudaFree(x); cudaFree(y);
  }
  return 0;
}
