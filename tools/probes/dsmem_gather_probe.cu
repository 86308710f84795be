// dsmem_gather_probe.cu -- can a cluster's distributed shared memory serve the random FP64 gather of a
// sparse mat-vec faster than L2 (272 G elements/s, profiles/r1b_gather_probe_*.txt)?  (dev probe, not product)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dsmem_probe tools/probes/dsmem_gather_probe.cu
// Each CTA of a cluster of CS holds SLAB doubles of x; every thread streams column indices from global
// memory and gathers x[idx] from whichever CTA of the cluster owns it (mapa + ld.shared::cluster).
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned hash32(unsigned a, unsigned b) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  h ^= h >> 15; h *= 0xC2B2AE3Du; h ^= h >> 13; h *= 0x27D4EB2Fu; h ^= h >> 16;
  return h;
}
__global__ void fill(int *idx, double *val, long long nnz, int ncols) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
    idx[k] = hash32((unsigned)(k >> 5), (unsigned)(k & 31)) % (unsigned)ncols;
    val[k] = 1.0 + (k & 7) * 0.125;
  }
}
__global__ void fillx(double *x, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) x[k] = 1.0 + (k % 13) * 0.01;
}

__device__ __forceinline__ double ld_cluster(unsigned local_addr, unsigned cta) {
  unsigned ra; double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra));
  return v;
}

// MODE 0: gather over the whole cluster; MODE 1: gather only from the own CTA (local smem ceiling);
// MODE 2: full mat-vec shape: val + idx stream, products, one partial written per 32 entries.
template <int U, int MODE>
__global__ void k_dsmem(const int *__restrict__ idx, const double *__restrict__ val, const double *__restrict__ x,
                        long long nnz, int slab, int cs, double *__restrict__ out) {
  extern __shared__ double sx[];
  cg::cluster_group cl = cg::this_cluster();
  const unsigned rank = cl.block_rank();
  for (int i = threadIdx.x; i < slab; i += blockDim.x) sx[i] = x[rank * slab + i];
  cl.sync();
  const unsigned base = (unsigned)__cvta_generic_to_shared(sx);
  double acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += stride * U) {
    int c[U]; double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long kk = k + u * stride;
      c[u] = kk < nnz ? __ldcs(idx + kk) : 0;
      if (MODE == 2) v[u] = kk < nnz ? __ldcs(val + kk) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned cta = MODE == 1 ? rank : (unsigned)c[u] / (unsigned)slab;
      const unsigned off = (unsigned)c[u] % (unsigned)slab;
      const double g = ld_cluster(base + off * 8u, cta);
      if (MODE == 2) acc = fma(v[u], g, acc); else acc += g;
    }
    if (MODE == 2) {
      double r = acc;
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if ((threadIdx.x & 31) == 0) out[(k >> 5) % (1 << 20)] = r;
      acc = 0;
    }
  }
  if (acc == 123.456) out[0] = acc;
  cl.sync();
}

template <int U, int MODE>
float run(int cs, int nclusters, int threads, int slab, const int *idx, const double *val, const double *x, long long nnz, double *out, int reps) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs * nclusters); cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = (size_t)slab * 8;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  auto kern = k_dsmem<U, MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, slab * 8));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int maxc = 0;
  cudaError_t qe = cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg);
  if (qe != cudaSuccess) { printf("  (occupancy query: %s)\n", cudaGetErrorString(qe)); cudaGetLastError(); }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) { cudaError_t e = cudaLaunchKernelEx(&cfg, kern, idx, val, x, nnz, slab, cs, out); if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); return -1.f; } }
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&cfg, kern, idx, val, x, nnz, slab, cs, out);
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= reps;
  printf("cs %2d clusters %3d (max active %3d) thr %4d slab %5d U %2d mode %d : %.3f ms  %.1f Gelem/s\n", cs, nclusters, maxc, threads, slab, U, MODE, ms,
         nnz / ms / 1e6);
  return ms;
}

int main(int argc, char **argv) {
  const long long nnz = argc > 1 ? atoll(argv[1]) : 100000000ll;
  int *idx; double *val, *x, *out;
  CK(cudaMalloc(&idx, nnz * 4)); CK(cudaMalloc(&val, nnz * 8)); CK(cudaMalloc(&x, (size_t)(1 << 20) * 8)); CK(cudaMalloc(&out, (size_t)(1 << 20) * 8));
  fillx<<<148 * 8, 256>>>(x, 1 << 20);
  const int reps = 5;
  for (int cs : {16, 8, 4, 2}) {
    const int slab = 25600;  // 200 KB per CTA
    fill<<<148 * 8, 256>>>(idx, val, nnz, slab * cs);
    CK(cudaDeviceSynchronize());
    const int ncl = 148 / cs;
    run<8, 0>(cs, ncl, 1024, slab, idx, val, x, nnz, out, reps);
    run<4, 0>(cs, ncl, 1024, slab, idx, val, x, nnz, out, reps);
    run<16, 0>(cs, ncl, 512, slab, idx, val, x, nnz, out, reps);
    run<8, 0>(cs, ncl, 512, slab, idx, val, x, nnz, out, reps);
    run<8, 2>(cs, ncl, 1024, slab, idx, val, x, nnz, out, reps);
    run<4, 2>(cs, ncl, 1024, slab, idx, val, x, nnz, out, reps);
    run<8, 2>(cs, ncl, 512, slab, idx, val, x, nnz, out, reps);
    if (cs == 8) {  // two CTAs per SM, 100 KB slabs
      fill<<<148 * 8, 256>>>(idx, val, nnz, 12800 * cs);
      CK(cudaDeviceSynchronize());
      run<8, 0>(cs, 2 * ncl, 1024, 12800, idx, val, x, nnz, out, reps);
      run<8, 2>(cs, 2 * ncl, 1024, 12800, idx, val, x, nnz, out, reps);
      fill<<<148 * 8, 256>>>(idx, val, nnz, slab * cs);
      CK(cudaDeviceSynchronize());
    }
  }
  // local shared-memory ceiling
  fill<<<148 * 8, 256>>>(idx, val, nnz, 25600);
  CK(cudaDeviceSynchronize());
  run<8, 1>(1, 148, 1024, 25600, idx, val, x, nnz, out, reps);
  run<8, 2>(1, 148, 1024, 25600, idx, val, x, nnz, out, reps);
  return 0;
}
