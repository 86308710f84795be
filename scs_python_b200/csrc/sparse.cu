// sparse.cu -- building the device CSR copies (setup path, runs once per workspace).
//
// CSR(A') is the caller's CSC(A) verbatim.  CSR(A) is produced ON THE DEVICE by a stable
// LSD radix sort of (row index -> source position) -- the device counterpart of the CPU
// indirect backend's transpose() (S/linsys/cpu/indirect/private.c:7-46).  Stability keeps
// every output row in ascending column order, so summation order is a function of the
// matrix only.  CUB (part of the CUDA toolkit) is used for the sort; it is setup code, not
// the per-iteration hot path.
#include <cub/device/device_radix_sort.cuh>

#include "sparse.cuh"

namespace b200 {

int csr_upload(Ctx &c, CsrDev &out, int nrows, int ncols, const int *ptr, const int *idx, const double *val) {
  out.nrows = nrows;
  out.ncols = ncols;
  out.nnz = ptr[nrows];
  if (dev_alloc(&out.ptr, (size_t)nrows + 1) || dev_alloc(&out.idx, (size_t)out.nnz) ||
      dev_alloc(&out.val, (size_t)out.nnz))
    return -1;
  if (h2d(c, out.ptr, ptr, (size_t)nrows + 1) || h2d(c, out.idx, idx, (size_t)out.nnz) ||
      h2d(c, out.val, val, (size_t)out.nnz))
    return -1;
  return 0;
}

void csr_free(CsrDev &m) {
  dev_free(m.ptr);
  dev_free(m.idx);
  dev_free(m.val);
  m.nnz = 0;
}

// position k of the source -> its row (binary search in ptr), and identity permutation
__global__ void expand_rows_kernel(const int *__restrict__ ptr, int nrows, long long nnz, int *__restrict__ rowof,
                                   int *__restrict__ pos) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz;
       k += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = nrows;  // find r with ptr[r] <= k < ptr[r+1]
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (ptr[mid] <= k) lo = mid; else hi = mid;
    }
    rowof[k] = lo;
    pos[k] = (int)k;
  }
}

// out.ptr[r] = lower_bound(sorted_keys, r)
__global__ void rowptr_from_sorted_kernel(const int *__restrict__ keys, long long nnz, int nrows_out,
                                          int *__restrict__ ptr) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= nrows_out; r += gridDim.x * blockDim.x) {
    long long lo = 0, hi = nnz;  // first k with keys[k] >= r
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (keys[mid] < r) lo = mid + 1; else hi = mid;
    }
    ptr[r] = (int)lo;
  }
}

__global__ void gather_transposed_kernel(const int *__restrict__ perm, const int *__restrict__ rowof,
                                         const double *__restrict__ val, long long nnz, int *__restrict__ oidx,
                                         double *__restrict__ oval) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < nnz;
       k += (long long)gridDim.x * blockDim.x) {
    const int src = perm[k];
    oidx[k] = rowof[src];
    oval[k] = val[src];
  }
}

int csr_transpose(Ctx &c, const CsrDev &in, CsrDev &out) {
  out.nrows = in.ncols;
  out.ncols = in.nrows;
  out.nnz = in.nnz;
  const long long nnz = in.nnz;
  if (dev_alloc(&out.ptr, (size_t)out.nrows + 1) || dev_alloc(&out.idx, (size_t)nnz) ||
      dev_alloc(&out.val, (size_t)nnz))
    return -1;
  if (nnz == 0) {
    CUDA_OK(cudaMemsetAsync(out.ptr, 0, sizeof(int) * ((size_t)out.nrows + 1), c.stream));
    return 0;
  }
  if (nnz > 0x7fffffffLL) {
    fprintf(stderr, "libscsb200: nnz exceeds 32-bit index range\n");
    return -1;
  }
  int *rowof = nullptr, *pos = nullptr, *keys_out = nullptr, *perm = nullptr;
  if (dev_alloc(&rowof, (size_t)nnz) || dev_alloc(&pos, (size_t)nnz) || dev_alloc(&keys_out, (size_t)nnz) ||
      dev_alloc(&perm, (size_t)nnz))
    return -1;
  const int grid = c.grid_ew();
  expand_rows_kernel<<<grid, kThreads, 0, c.stream>>>(in.ptr, in.nrows, nnz, rowof, pos);
  c.launches++;
  int bits = 1;
  while ((1ll << bits) < (long long)in.ncols && bits < 31) ++bits;
  size_t tmp_bytes = 0;
  CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, in.idx, keys_out, pos, perm, (int)nnz, 0, bits,
                                          c.stream));
  void *tmp = nullptr;
  CUDA_OK(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, in.idx, keys_out, pos, perm, (int)nnz, 0, bits,
                                          c.stream));
  rowptr_from_sorted_kernel<<<grid, kThreads, 0, c.stream>>>(keys_out, nnz, out.nrows, out.ptr);
  gather_transposed_kernel<<<grid, kThreads, 0, c.stream>>>(perm, rowof, in.val, nnz, out.idx, out.val);
  c.launches += 2;
  CUDA_OK(cudaStreamSynchronize(c.stream));
  cudaFree(tmp);
  dev_free(rowof);
  dev_free(pos);
  dev_free(keys_out);
  dev_free(perm);
  CUDA_OK(cudaGetLastError());
  return 0;
}

int csr_from_upper_csc(Ctx &c, CsrDev &out, int n, const int *Pp, const int *Pi, const double *Px) {
  // Host expansion (P is small next to A): row r of the full matrix holds the strictly-lower
  // mirror entries (ascending col), then the upper-triangular column... built by counting.
  std::vector<int> cnt((size_t)n + 1, 0);
  for (int j = 0; j < n; ++j)
    for (int k = Pp[j]; k < Pp[j + 1]; ++k) {
      const int i = Pi[k];
      cnt[(size_t)i + 1]++;            // entry (i,j) in row i
      if (i != j) cnt[(size_t)j + 1]++;  // mirror (j,i) in row j
    }
  for (int r = 0; r < n; ++r) cnt[(size_t)r + 1] += cnt[r];
  const long long nnz = cnt[n];
  std::vector<int> idx((size_t)nnz);
  std::vector<double> val((size_t)nnz);
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  // pass 1: mirrored (strictly lower) entries: row j gets col i < j; iterate columns j ascending,
  // rows i ascending inside -> for fixed row j the cols i arrive ascending.
  for (int j = 0; j < n; ++j)
    for (int k = Pp[j]; k < Pp[j + 1]; ++k) {
      const int i = Pi[k];
      if (i != j) {
        idx[(size_t)fill[j]] = i;
        val[(size_t)fill[j]++] = Px[k];
      }
    }
  // pass 2: upper entries (i,j), j >= i: row i gets col j, columns ascending.
  for (int j = 0; j < n; ++j)
    for (int k = Pp[j]; k < Pp[j + 1]; ++k) {
      const int i = Pi[k];
      idx[(size_t)fill[i]] = j;
      val[(size_t)fill[i]++] = Px[k];
    }
  int rc = csr_upload(c, out, n, n, cnt.data(), idx.data(), val.data());
  if (rc == 0) rc = c.sync();  // host vectors die at scope exit
  return rc;
}

int chunks_build(Ctx &c, ChunkList &out, const CsrDev &a, const CsrDev *b) {
  std::vector<int> pa((size_t)a.nrows + 1), pb;
  CUDA_OK(cudaMemcpyAsync(pa.data(), a.ptr, sizeof(int) * pa.size(), cudaMemcpyDeviceToHost, c.stream));
  if (b) {
    pb.resize((size_t)b->nrows + 1);
    CUDA_OK(cudaMemcpyAsync(pb.data(), b->ptr, sizeof(int) * pb.size(), cudaMemcpyDeviceToHost, c.stream));
  }
  CUDA_OK(cudaStreamSynchronize(c.stream));
  std::vector<int4> ch;
  build_chunks_host(a.nrows, pa.data(), b ? pb.data() : nullptr, ch, c.sms);
  out.n = (int)ch.size();
  if (dev_alloc(&out.d, ch.size())) return -1;
  CUDA_OK(cudaMemcpyAsync(out.d, ch.data(), sizeof(int4) * ch.size(), cudaMemcpyHostToDevice, c.stream));
  CUDA_OK(cudaStreamSynchronize(c.stream));
  out.grid = out.n < c.grid_ew() ? (out.n > 0 ? out.n : 1) : c.grid_ew();
  return 0;
}

void chunks_free(ChunkList &cl) {
  dev_free(cl.d);
  cl.n = 0;
}

}  // namespace b200
