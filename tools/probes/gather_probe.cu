// gather_probe.cu -- what bounds a random-column FP64 SpMV on B200?  (dev probe, not product)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather_probe tools/probes/gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned hash32(unsigned a, unsigned b) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
  h ^= h >> 15; h *= 0xC2B2AE3Du; h ^= h >> 13; h *= 0x27D4EB2Fu; h ^= h >> 16;
  return h;
}
__global__ void fill(int *idx, double *val, long long nnz, int per_row, int ncols) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
    idx[k] = hash32((unsigned)(k / per_row), (unsigned)(k % per_row)) % (unsigned)ncols;
    val[k] = 1.0 + (k & 7) * 0.125;
  }
}
__global__ void fillx(double *x, int n) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) x[k] = 1.0 + (k % 13) * 0.01;
}

template <int F> __device__ __forceinline__ double gld(const double *p) {
  if (F == 0) return __ldg(p);
  if (F == 1) return __ldcg(p);
  if (F == 2) { double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
  if (F == 3) return __ldcv(p);
  double v; asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}

// pure stream: sum of val*idx
__global__ void __launch_bounds__(256) k_stream(const double *__restrict__ val, const int *__restrict__ idx, long long nnz, double *out) {
  double acc = 0;
  for (long long k = blockIdx.x * 256ll + threadIdx.x; k < nnz; k += (long long)gridDim.x * 256 * 4) {
    double v[4]; int c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { long long kk = k + (long long)u * gridDim.x * 256; if (kk < nnz) { v[u] = __ldcs(val + kk); c[u] = __ldcs(idx + kk);} else { v[u] = 0; c[u] = 0; } }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += v[u] * c[u];
  }
  if (acc == 123.456) out[0] = acc;
}
// pure gather: acc += x[idx[k]] with U loads in flight per thread
template <int F, int U>
__global__ void __launch_bounds__(256) k_gather(const int *__restrict__ idx, const double *__restrict__ x, long long nnz, double *out) {
  double acc = 0;
  const long long stride = (long long)gridDim.x * 256;
  for (long long k = blockIdx.x * 256ll + threadIdx.x; k < nnz; k += stride * U) {
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { long long kk = k + u * stride; c[u] = kk < nnz ? __ldcs(idx + kk) : 0; }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += gld<F>(x + c[u]);
  }
  if (acc == 123.456) out[0] = acc;
}
// warp-per-row CSR SpMV, fixed row length
template <int F>
__global__ void __launch_bounds__(256) k_rowvec(const double *__restrict__ val, const int *__restrict__ idx, int per_row, int nrows,
                                                const double *__restrict__ x, double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int wpg = (gridDim.x * 256) >> 5;
  for (int row = (blockIdx.x * 256 + threadIdx.x) >> 5; row < nrows; row += wpg) {
    const long long s = (long long)row * per_row, e = s + per_row;
    double acc = 0;
    long long k = s + lane;
    for (; k + 32 < e; k += 64) {
      double v0 = __ldcs(val + k), v1 = __ldcs(val + k + 32);
      int c0 = __ldcs(idx + k), c1 = __ldcs(idx + k + 32);
      acc = fma(v0, gld<F>(x + c0), acc);
      acc = fma(v1, gld<F>(x + c1), acc);
    }
    if (k < e) acc = fma(__ldcs(val + k), gld<F>(x + __ldcs(idx + k)), acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = acc;
  }
}
// tile-staged: each CTA stages TILE products into smem then reduces rows (fixed row length)
template <int F, int IPT>
__global__ void __launch_bounds__(256) k_tile(const double *__restrict__ val, const int *__restrict__ idx, int per_row, int nrows,
                                              const double *__restrict__ x, double *__restrict__ y) {
  constexpr int TILE = 256 * IPT;
  __shared__ double tile[TILE];
  const int rows_per_tile = TILE / per_row;
  const int ntiles = (nrows + rows_per_tile - 1) / rows_per_tile;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int r0 = t * rows_per_tile;
    const int nr = min(rows_per_tile, nrows - r0);
    const long long k0 = (long long)r0 * per_row;
    const int cnt = nr * per_row;
    double v[IPT]; int c[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) { int k = threadIdx.x + i * 256; if (k < cnt) { v[i] = __ldcs(val + k0 + k); c[i] = __ldcs(idx + k0 + k); } }
#pragma unroll
    for (int i = 0; i < IPT; ++i) { int k = threadIdx.x + i * 256; if (k < cnt) tile[k] = v[i] * gld<F>(x + c[i]); }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int r = w; r < nr; r += 8) {
      double acc = 0;
      for (int k = r * per_row + lane; k < (r + 1) * per_row; k += 32) acc += tile[k];
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) y[r0 + r] = acc;
    }
    __syncthreads();
  }
}

template <class L>
float timeit(L launch, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); launch();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b); CK(cudaEventSynchronize(b));
  float ms; cudaEventElapsedTime(&ms, a, b);
  CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char **argv) {
  const int nrows = argc > 1 ? atoi(argv[1]) : 2000000;
  const int per_row = argc > 2 ? atoi(argv[2]) : 51;
  const int ncols = argc > 3 ? atoi(argv[3]) : 1000000;
  const long long nnz = (long long)nrows * per_row;
  int *idx; double *val, *x, *y, *out;
  CK(cudaMalloc(&idx, nnz * 4)); CK(cudaMalloc(&val, nnz * 8)); CK(cudaMalloc(&x, (size_t)ncols * 8)); CK(cudaMalloc(&y, (size_t)nrows * 8));
  CK(cudaMalloc(&out, 8));
  fill<<<148 * 8, 256>>>(idx, val, nnz, per_row, ncols);
  fillx<<<148 * 8, 256>>>(x, ncols);
  CK(cudaDeviceSynchronize());
  const double bytes = 12.0 * nnz + 8.0 * nrows + 8.0 * ncols;
  printf("nrows %d per_row %d ncols %d nnz %lld  alg bytes %.1f MB\n", nrows, per_row, ncols, nnz, bytes / 1e6);
  const int reps = 10;
  for (int g : {148 * 4, 148 * 8, 148 * 16}) {
    float ms = timeit([&] { k_stream<<<g, 256>>>(val, idx, nnz, out); }, reps);
    printf("stream val+idx        grid %5d : %.3f ms  %.0f GB/s\n", g, ms, 12.0 * nnz / ms / 1e6);
  }
#define GATHER(F, U, g) { float ms = timeit([&] { k_gather<F, U><<<g, 256>>>(idx, x, nnz, out); }, reps); \
    printf("gather F%d U%d          grid %5d : %.3f ms  %.1f Gelem/s (idx stream %.0f GB/s)\n", F, U, g, ms, nnz / ms / 1e6, 4.0 * nnz / ms / 1e6); }
  GATHER(0, 4, 148 * 8) GATHER(0, 8, 148 * 8) GATHER(0, 16, 148 * 8) GATHER(0, 8, 148 * 4) GATHER(0, 8, 148 * 16)
  GATHER(1, 8, 148 * 8) GATHER(2, 8, 148 * 8) GATHER(3, 8, 148 * 8) GATHER(4, 8, 148 * 8)
  GATHER(1, 16, 148 * 8) GATHER(2, 16, 148 * 8)
#define ROWVEC(F, g) { float ms = timeit([&] { k_rowvec<F><<<g, 256>>>(val, idx, per_row, nrows, x, y); }, reps); \
    printf("rowvec F%d             grid %5d : %.3f ms  %.0f GB/s\n", F, g, ms, bytes / ms / 1e6); }
  ROWVEC(0, 148 * 8) ROWVEC(1, 148 * 8) ROWVEC(2, 148 * 8) ROWVEC(4, 148 * 8) ROWVEC(0, 148 * 16) ROWVEC(0, 148 * 32)
#define TILEK(F, I, g) { float ms = timeit([&] { k_tile<F, I><<<g, 256>>>(val, idx, per_row, nrows, x, y); }, reps); \
    printf("tile F%d IPT%d          grid %5d : %.3f ms  %.0f GB/s\n", F, I, g, ms, bytes / ms / 1e6); }
  TILEK(0, 8, 148 * 4) TILEK(0, 8, 148 * 8) TILEK(0, 4, 148 * 8) TILEK(0, 16, 148 * 4) TILEK(1, 8, 148 * 8) TILEK(2, 8, 148 * 8) TILEK(0, 8, 148 * 16)
  // checksum
  std::vector<double> hy(8); CK(cudaMemcpy(hy.data(), y, 64, cudaMemcpyDeviceToHost));
  printf("y[0..2] = %.6f %.6f %.6f\n", hy[0], hy[1], hy[2]);
  return 0;
}
