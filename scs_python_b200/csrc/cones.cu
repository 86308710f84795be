// cones.cu -- cone projection kernels for sm_100a.
//
//   box    : one CTA, Newton on t with a block-reduced gradient/Hessian (cones.c:1174-1237)
//   SOC    : one warp per cone (one CTA for cones > 2048), shuffle-reduced norm (cones.c:1242-1271)
//   exp    : one thread per triple, Friberg-2021 root search (exp_cone.c)
//   power  : one thread per triple, Newton on r (cones.c:1276-1324)
//   PSD / complex PSD : one CTA per cone, parallel-ordered one-sided Jacobi on the shifted
//            matrix, reconstruction from the positive eigenpairs (replaces dsyevr+dsyrk /
//            zheevr+zherk, cones.c:991-1148); complex cones go through the real 2s x 2s
//            embedding [[Re,-Im],[Im,Re]].
#include <cooperative_groups.h>

#include "cones.cuh"
#include "cone_dev.cuh"

namespace b200 {

// =========================================================================== box ======
constexpr int kBoxThreads = 1024;

__global__ void __launch_bounds__(kBoxThreads)
k_box_cone(double *__restrict__ tx, const double *__restrict__ s_saved, const double *__restrict__ r_box,
           const double *__restrict__ bl, const double *__restrict__ bu, int bsize, double *t_warm) {
  __shared__ double sh[2 * 32];
  __shared__ double sh_t;
  __shared__ int sh_stop;
  const int tid = threadIdx.x;
  if (bsize == 1) {
    if (tid == 0) tx[0] = fmax(tx[0], 0.0) / r_box[0] + s_saved[0];
    return;
  }
  double *x = tx + 1;
  const double *rho = r_box + 1;
  const double rho_t = 1.0 / r_box[0];
  const double tx0 = tx[0];
  double t = *t_warm;
  for (int iter = 0; iter < 25; ++iter) {  // BOX_CONE_MAX_ITERS
    double v[2] = {0.0, 0.0};              // gt, ht partial sums
    for (int j = tid; j < bsize - 1; j += kBoxThreads) {
      const double rinv = 1.0 / rho[j];
      const double xj = x[j], u = bu[j], lo = bl[j];
      if (xj > t * u) {
        v[0] += rinv * (t * u - xj) * u;
        v[1] += rinv * u * u;
      } else if (xj < t * lo) {
        v[0] += rinv * (t * lo - xj) * lo;
        v[1] += rinv * lo * lo;
      }
    }
    block_reduce<2, 0>(v, sh);
    if (tid == 0) {
      const double gt = rho_t * (t - tx0) + v[0];
      const double ht = rho_t + v[1];
      const double t_prev = t;
      t = fmax(t - gt / fmax(ht, 1e-8), 0.0);
      sh_t = t;
      sh_stop = (fabs(gt / fmax(ht, 1e-6)) < 1e-12 * fmax(t, 1.0) || fabs(t - t_prev) < 1e-11 * fmax(t, 1.0));
    }
    __syncthreads();
    t = sh_t;
    const int stop = sh_stop;
    __syncthreads();
    if (stop) break;
  }
  for (int j = tid; j < bsize - 1; j += kBoxThreads) {
    double xj = x[j];
    if (xj > t * bu[j]) xj = t * bu[j];
    else if (xj < t * bl[j]) xj = t * bl[j];
    x[j] = xj / rho[j] + s_saved[1 + j];
  }
  if (tid == 0) {
    tx[0] = t / r_box[0] + s_saved[0];
    *t_warm = t;
  }
}

// =========================================================================== SOC ======
template <int GROUP>
__device__ __forceinline__ double group_sum(double v, double *sh) {
  v = warp_sum(v);
  if (GROUP == 32) return v;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < GROUP / 32; ++w) t += sh[w];
  return t;
}

template <int GROUP>
__device__ __forceinline__ void soc_project(double *__restrict__ x, const double *__restrict__ sv,
                                            const double *__restrict__ r, int q, int lane, double *sh) {
  if (q <= 0) return;
  if (q == 1) {
    if (lane == 0) x[0] = fmax(x[0], 0.0) / r[0] + sv[0];
    return;
  }
  const double v1 = x[0];
  double ss = 0.0;
  for (int k = 1 + lane; k < q; k += GROUP) ss = fma(x[k], x[k], ss);
  ss = group_sum<GROUP>(ss, sh);
  const double s = (q == 2) ? fabs(x[1]) : sqrt(ss);
  const double alpha = (s + v1) / 2.0;
  double scale_tail, head;
  if (s <= v1) {            // inside the cone
    scale_tail = 1.0; head = v1;
  } else if (s <= -v1) {    // inside the polar cone -> 0
    scale_tail = 0.0; head = 0.0;
  } else {
    scale_tail = alpha / s; head = alpha;
  }
  for (int k = 1 + lane; k < q; k += GROUP) x[k] = (x[k] * scale_tail) / r[k] + sv[k];
  if (lane == 0) x[0] = head / r[0] + sv[0];
}

__global__ void __launch_bounds__(kThreads)
k_soc_small(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r,
            const int *__restrict__ off, const int *__restrict__ len, int ncones) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int cidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cidx < ncones; cidx += warps_per_grid) {
    const int o = off[cidx];
    soc_project<32>(x + o, sv + o, r + o, len[cidx], lane, nullptr);
  }
}

__global__ void __launch_bounds__(kThreads)
k_soc_large(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r,
            const int *__restrict__ off, const int *__restrict__ len, int ncones) {
  __shared__ double sh[kThreads / 32];
  for (int cidx = blockIdx.x; cidx < ncones; cidx += gridDim.x) {
    const int o = off[cidx];
    soc_project<kThreads>(x + o, sv + o, r + o, len[cidx], threadIdx.x, sh);
    __syncthreads();
  }
}

// exponential and power cones: device functions in cone_dev.cuh (shared with the batch engine)
__global__ void __launch_bounds__(128)
k_exp_cones(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r, int ep, int ntot) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x) {
    double v[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    proj_pd_exp_cone(v, i < ep);
#pragma unroll
    for (int k = 0; k < 3; ++k) x[3 * i + k] = v[k] / r[3 * i + k] + sv[3 * i + k];
  }
}

__global__ void __launch_bounds__(128)
k_pow_cones(double *__restrict__ x, const double *__restrict__ sv, const double *__restrict__ r,
            const double *__restrict__ pw, int ntot) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ntot; i += gridDim.x * blockDim.x) {
    double v[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]};
    const double a = pw[i];
    if (a >= 0) {
      proj_power_cone(v, a);
    } else {  // dual power cone via Moreau, cones.c:1423-1432
      double w[3] = {-v[0], -v[1], -v[2]};
      proj_power_cone(w, -a);
      v[0] += w[0]; v[1] += w[1]; v[2] += w[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) x[3 * i + k] = v[k] / r[3 * i + k] + sv[3 * i + k];
  }
}

// =========================================================== PSD / complex PSD ========
// Eigen-projection pipeline (replaces dsyevr + dsyrk / zheevr + zherk, cones.c:991-1148):
//
//   k_psd_prep     unpack x into the full symmetric W = mat(x) + sigma I (sigma = ||mat(x)||_F, so W
//                  is PSD and its SVD is its eigendecomposition); decide cold / warm start
//   k_psd_gemm_wv  warm start: G = W V_prev (V_prev = eigenvectors of the previous ADMM iteration,
//                  an orthogonal basis in which W is already nearly diagonal)
//   k_psd_jacobi   parallel-ordered one-sided Jacobi on (G, V) until no pair rotates; cones
//                  with d > 64 get a 4-CTA thread-block cluster (one warp per column pair,
//                  cluster barrier per round), small cones one CTA
//   k_psd_lambda   lambda_k = v_k . g_k - sigma
//   k_psd_recon    X+ = sum_{lambda_k > 0} lambda_k v_k v_k' as a tiled FP64 GEMM, re-packed with
//                  the Moreau recombination fused
//
// Complex cones go through the real 2s x 2s embedding [[Re,-Im],[Im,Re]].  The basis is rebuilt
// from the identity every kPsdRefresh projections so rounding drift of V cannot accumulate.

// packed index of entry (i,j), i >= j, of the real lower-triangular column-major layout
__device__ __forceinline__ long long tri_idx(int i, int j, int n) {
  return (long long)j * n - ((long long)(j - 1) * j) / 2 + (i - j);
}

constexpr int kPsdRefresh = 64;
constexpr int kPsdMaxSweeps = 40;
constexpr int kPsdSmallDim = 96;   // d <= 96: (G, V) fit one CTA's shared memory; larger: one cluster per cone
constexpr int kPsdCluster = 4;
constexpr int kPsdJacMaxWarps = 26;  // 832 threads: 78 registers per thread
constexpr int kPsdEpl = 8;         // column elements per lane held in registers (d <= 256)
constexpr int kPsdPrepThreads = 512;
constexpr int kGT = 64, kGK = 16;  // GEMM tile: 64 x 64 outputs, K chunks of 16, 256 threads (4 x 4 each)

// entry (row, col) of the full symmetric working matrix (without the shift)
__device__ __forceinline__ double psd_unpack_entry(const double *__restrict__ x, const PsdEntry &e, int row, int col) {
  const int n = e.s;
  const double sqrt2 = 1.4142135623730951;
  if (!e.is_complex) {
    const int i = row > col ? row : col, j = row > col ? col : row;
    const double val = x[tri_idx(i, j, n)];
    return (i == j) ? val * sqrt2 : val;  // diagonal * sqrt2, cones.c:1011-1017
  }
  // H = A + iB, embedding [[A, -B], [B, A]]; column j of the packed layout starts at
  // j*(2n-j): real diagonal, then (re, im) pairs of rows j+1.. (cones.c:1088-1095)
  const int br = row / n, bc = col / n, i0 = row % n, j0 = col % n;
  const int i = i0 > j0 ? i0 : j0, j = i0 > j0 ? j0 : i0;
  const long long base = (long long)j * (2 * n - j);
  if (i == j) return (br == bc) ? x[base] * sqrt2 : 0.0;
  const double re = x[base + 1 + 2 * (i - j - 1)], im = x[base + 2 + 2 * (i - j - 1)];
  const double bij = (i0 > j0) ? im : -im;  // B is antisymmetric
  if (br == bc) return re;
  return (br == 1) ? bij : -bij;  // bottom-left block = B, top-right = -B
}

__global__ void __launch_bounds__(kPsdPrepThreads)
k_psd_prep(double *__restrict__ xall, const double *__restrict__ svall, const double *__restrict__ rall,
           const PsdEntry *__restrict__ ents, PsdState *__restrict__ state, double *__restrict__ Wall,
           double *__restrict__ Gall, double *__restrict__ Vall) {
  __shared__ double sh[32];
  __shared__ double sh_sigma;
  const PsdEntry e = ents[blockIdx.x];
  const int n = e.s, d = e.d, tid = threadIdx.x, nthr = blockDim.x;
  double *x = xall + e.off;
  if (n == 0) return;
  if (n == 1) {
    if (tid == 0) x[0] = fmax(x[0], 0.0) / rall[e.off] + svall[e.off];
    return;
  }
  const int len = e.is_complex ? n * n : n * (n + 1) / 2;
  double v[1] = {0.0};
  for (int k = tid; k < len; k += nthr) v[0] = fma(x[k], x[k], v[0]);
  block_reduce<1, 0>(v, sh);
  if (tid == 0) sh_sigma = sqrt((e.is_complex ? 4.0 : 2.0) * v[0]);  // == Frobenius norm of the full matrix
  __syncthreads();
  const double sigma = sh_sigma;
  const int cold = (state[blockIdx.x].age % kPsdRefresh) == 0;
  double *W = Wall + e.woff, *G = Gall + e.woff, *V = Vall + e.woff;
  for (long long idx = tid; idx < (long long)d * d; idx += nthr) {
    const int col = (int)(idx / d), row = (int)(idx % d);
    const double val = psd_unpack_entry(x, e, row, col) + (row == col ? sigma : 0.0);
    W[idx] = val;
    if (cold) {
      G[idx] = val;
      V[idx] = (row == col) ? 1.0 : 0.0;
    }
  }
  if (tid == 0) {
    state[blockIdx.x].cold = cold;
    state[blockIdx.x].sigma = sigma;
  }
}

// acc[a][b] += sum_k A(k, i0 + tx + 16 a) * B(k, j0 + ty + 16 b): 64 x 64 tile, 256 threads.
// loadA(k, i) / loadB(k, j) return 0 outside the matrix.
template <class LA, class LB>
__device__ __forceinline__ void tile_gemm(double (&acc)[4][4], int K, LA loadA, LB loadB, bool b_k_contiguous) {
  __shared__ double As[kGK][kGT + 1];
  __shared__ double Bs[kGK][kGT + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int k0 = 0; k0 < K; k0 += kGK) {
#pragma unroll
    for (int t = 0; t < (kGK * kGT) / 256; ++t) {
      const int idx = tid + 256 * t;
      const int kk = idx / kGT, i = idx % kGT;
      As[kk][i] = loadA(k0 + kk, i);
      if (b_k_contiguous) {
        const int j = idx / kGK, k2 = idx % kGK;
        Bs[k2][j] = loadB(k0 + k2, j);
      } else {
        Bs[kk][i] = loadB(k0 + kk, i);
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kGK; ++kk) {
      double av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
}

// warm start: G = W V   (grid: tiles*tiles x cones)
__global__ void __launch_bounds__(256)
k_psd_gemm_wv(const PsdEntry *__restrict__ ents, const PsdState *__restrict__ state, const double *__restrict__ Wall,
              const double *__restrict__ Vall, double *__restrict__ Gall, int tiles) {
  const PsdEntry e = ents[blockIdx.y];
  const int d = e.d;
  if (e.s < 2 || state[blockIdx.y].cold) return;
  const int ti = blockIdx.x % tiles, tj = blockIdx.x / tiles;
  const int i0 = ti * kGT, j0 = tj * kGT;
  if (i0 >= d || j0 >= d) return;
  const double *W = Wall + e.woff, *V = Vall + e.woff;
  double *G = Gall + e.woff;
  double acc[4][4] = {};
  tile_gemm(acc, d,
            [&](int k, int i) { return (k < d && i0 + i < d) ? W[(long long)k * d + i0 + i] : 0.0; },
            [&](int k, int j) { return (k < d && j0 + j < d) ? V[(long long)(j0 + j) * d + k] : 0.0; }, true);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
      if (i < d && j < d) G[(long long)j * d + i] = acc[a][b];
    }
}

// Rotation (cs, sn) that orthogonalises two columns with squared norms a, b and inner product g:
// t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 g).  The angle only has to be
// accurate enough to keep the convergence quadratic-ish; what must hold to rounding is
// cs^2 + sn^2 = 1 (V stays orthogonal), which cs = rsqrt(1 + t^2), sn = cs t gives.
__device__ __forceinline__ void jacobi_angle(double a, double b, double g, double &cs, double &sn) {
  const double diff = b - a;
  const double q = fma(diff, diff, 4.0 * g * g);
  const double rad = q * rsqrt(q);                       // sqrt(q), q > 0 here
  const double t = (diff >= 0 ? 2.0 * g : -2.0 * g) * __drcp_rn(fabs(diff) + rad);
  cs = rsqrt(fma(t, t, 1.0));
  sn = cs * t;
}

// A sweep in which every rotated pair was already orthogonal to kPsdSmallCos (|cos| of the angle between the two
// columns) is the last one: one-sided Jacobi converges quadratically, so after those rotations the largest
// cosine is of the order kPsdSmallCos^2 * d ~ 1e-12 relative to the column norms -- below what the projection
// needs (tests: 1e-9 against LAPACK) -- and the sweep that would only verify it (a third to a half of the cost of
// a rotating sweep: all the dot products, no rotation) is not run.
constexpr double kPsdSmallCos2 = 1e-14;  // (1e-7)^2

// One column pair of the one-sided Jacobi iteration on shared-memory columns: rotate
// (g_p, g_q) and (v_p, v_q) so that g_p . g_q = 0.  Returns 0: no rotation, 1: rotated a pair that was
// orthogonal to kPsdSmallCos, 2: rotated a pair that was not.
__device__ __forceinline__ int jacobi_pair_smem(double *gp, double *gq, double *vp, double *vq, int d, int lane,
                                                 double tol2) {
  double a = 0.0, b = 0.0, g = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double u = gp[i], w = gq[i];
    a = fma(u, u, a); b = fma(w, w, b); g = fma(u, w, g);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {  // three interleaved butterflies: one latency chain, not three
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
  }
  if (!(g * g > tol2 * (a * b))) return 0;  // |g| <= tol sqrt(a b): already orthogonal (also NaN / g == 0)
  const int level = (g * g > kPsdSmallCos2 * (a * b)) ? 2 : 1;
  // tan of the rotation angle: t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), zeta = (b - a) / (2 g),
  // written with one sqrt, one division and one rsqrt
  double cs, sn;
  jacobi_angle(a, b, g, cs, sn);
  for (int i = lane; i < d; i += 32) {
    const double u = gp[i], w = gq[i];
    gp[i] = cs * u - sn * w;
    gq[i] = sn * u + cs * w;
    const double vu = vp[i], vw = vq[i];
    vp[i] = cs * vu - sn * vw;
    vq[i] = sn * vu + cs * vw;
  }
  return level;
}

// round `rnd` of the round-robin tournament over `cnt` (even) players: the k-th pairing
__device__ __forceinline__ void tournament_pair(int cnt, int rnd, int k, int &p, int &q) {
  if (k == 0) { p = cnt - 1; q = rnd; }
  else { p = (rnd + k) % (cnt - 1); q = (rnd - k + (cnt - 1)) % (cnt - 1); }
}

// number of column blocks for a d x d problem solved by C CTAs with `budget` bytes of shared
// memory each: even, >= 2C, and two blocks of (G, V) columns fit the budget
__host__ __device__ inline int psd_num_blocks(int d, int C, size_t budget) {
  int nb_blocks = 2 * C;
  while (true) {
    const int nb = (d + nb_blocks - 1) / nb_blocks;
    if ((size_t)32 * nb * d <= budget || nb == 1) return nb_blocks;
    nb_blocks += 2;
  }
}

// Blocked one-sided Jacobi, shared-memory resident.  The d columns of (G, V) are cut into NB
// blocks.  A block round pairs the blocks up (tournament over blocks); each CTA of the cluster
// stages its block pair (2 nb columns of G and of V) in shared memory, orthogonalises ALL
// column pairs among them with a local tournament (one warp per pair, __syncthreads per local
// round), and writes the block pair back; the CTAs of a cluster meet at a cluster barrier once
// per block round.  C = 1: one CTA per cone (NB = 2 when the whole problem fits).
template <int C>
__global__ void __launch_bounds__(kPsdJacMaxWarps * 32)
k_psd_jacobi(const PsdEntry *__restrict__ ents, PsdState *__restrict__ state, int first, double *__restrict__ Gall,
             double *__restrict__ Vall, size_t budget) {
  namespace cg = cooperative_groups;
  extern __shared__ double psd_smem[];
  __shared__ int sh_rot;
  const int cone = first + blockIdx.x / C;
  const PsdEntry e = ents[cone];
  const int d = e.d;
  if (e.s < 2) return;  // uniform over the cluster
  int rank = 0;
  if (C > 1) rank = (int)cg::this_cluster().block_rank();
  auto sync_all = [&]() {
    if (C > 1) cg::this_cluster().sync();
    else __syncthreads();
  };
  double *G = Gall + e.woff, *V = Vall + e.woff;
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nthr >> 5;
  const int NB = psd_num_blocks(d, C, budget);
  const int nb = (d + NB - 1) / NB;        // columns per block (last blocks may be short or empty)
  double *Gs = psd_smem, *Vs = psd_smem + (size_t)2 * nb * d;
  const double tol2 = (double)d * DBL_EPSILON * DBL_EPSILON;  // (sqrt(d) eps)^2
  int sweep = 0;
  for (; sweep < kPsdMaxSweeps; ++sweep) {
    if (tid == 0) sh_rot = 0;
    sync_all();
    for (int rb = 0; rb < NB - 1; ++rb) {
      for (int kb = rank; kb < NB / 2; kb += C) {
        int ba, bb;
        tournament_pair(NB, rb, kb, ba, bb);
        if (ba > bb) { const int t = ba; ba = bb; bb = t; }
        const int a0 = min(ba * nb, d), a1 = min(a0 + nb, d), b0 = min(bb * nb, d), b1 = min(b0 + nb, d);
        const int na = a1 - a0, nloc = na + (b1 - b0);
        if (nloc < 2) continue;
        // stage the block pair (L1 bypassed: another CTA wrote these columns last round)
        for (int idx = tid; idx < nloc * d; idx += nthr) {
          const int c = idx / d, i = idx - c * d;
          const long long src = (long long)(c < na ? a0 + c : b0 + (c - na)) * d + i;
          Gs[idx] = __ldcg(G + src);
          Vs[idx] = __ldcg(V + src);
        }
        __syncthreads();
        const int nbb = nloc - na;
        int rot = 0;  // 0 / 1 / 2, see jacobi_pair_smem
        if (rb == 0) {
          // pairs inside each of the two blocks, once per sweep (every block appears exactly once in
          // a block round): two independent tournaments side by side
          const int ea = na + (na & 1), eb = nbb + (nbb & 1);
          const int pa = ea / 2, pb = eb / 2;
          const int rounds = (ea > eb ? ea : eb) - 1;
          for (int rnd = 0; rnd < rounds; ++rnd) {
            for (int k = wid; k < pa + pb; k += nw) {
              int p, q, base, cnt, cap;
              if (k < pa) { base = 0; cnt = ea; cap = na; tournament_pair(ea, rnd % (ea > 1 ? ea - 1 : 1), k, p, q); }
              else { base = na; cnt = eb; cap = nbb; tournament_pair(eb, rnd % (eb > 1 ? eb - 1 : 1), k - pa, p, q); }
              if (rnd >= cnt - 1 || p >= cap || q >= cap) continue;
              if (p > q) { const int t = p; p = q; q = t; }
              p += base; q += base;
              rot = max(rot, jacobi_pair_smem(Gs + (size_t)p * d, Gs + (size_t)q * d, Vs + (size_t)p * d, Vs + (size_t)q * d, d, lane, tol2));
            }
            __syncthreads();
          }
        }
        // pairs across the two blocks: bipartite round robin, column k of A meets column
        // (k + rnd) mod nmax of B.  Column k of A stays with warp k for the whole phase, so it is
        // kept in registers (d <= 256) and only B's columns go through shared memory.
        const int nmax = na > nbb ? na : nbb;
        if (d <= 32 * kPsdEpl) {
          for (int k0 = 0; k0 < nmax; k0 += nw) {  // same trip count for every warp (barriers inside)
            const int k = k0 + wid;
            const bool have_p = k < na;
            double gpr[kPsdEpl], vpr[kPsdEpl];
            double a = 0.0;
#pragma unroll
            for (int t = 0; t < kPsdEpl; ++t) {
              const int i = lane + 32 * t;
              gpr[t] = (have_p && i < d) ? Gs[(size_t)k * d + i] : 0.0;
              vpr[t] = (have_p && i < d) ? Vs[(size_t)k * d + i] : 0.0;
              a = fma(gpr[t], gpr[t], a);
            }
            a = warp_sum(a);
            // all warps of this pass walk the rounds together (block barrier per round)
            for (int rnd = 0; rnd < nmax; ++rnd) {
              const int q = (k + rnd) % nmax;
              if (have_p && q < nbb) {
                double *gq = Gs + (size_t)(na + q) * d, *vq = Vs + (size_t)(na + q) * d;
                double w[kPsdEpl];
                double b = 0.0, g = 0.0;
#pragma unroll
                for (int t = 0; t < kPsdEpl; ++t) {
                  const int i = lane + 32 * t;
                  w[t] = (i < d) ? gq[i] : 0.0;
                  b = fma(w[t], w[t], b);
                  g = fma(gpr[t], w[t], g);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  b += __shfl_xor_sync(0xffffffffu, b, o);
                  g += __shfl_xor_sync(0xffffffffu, g, o);
                }
                if (g * g > tol2 * (a * b)) {
                  rot = max(rot, (g * g > kPsdSmallCos2 * (a * b)) ? 2 : 1);
                  double cs, sn;
                  jacobi_angle(a, b, g, cs, sn);
                  double a_new = 0.0;
#pragma unroll
                  for (int t = 0; t < kPsdEpl; ++t) {
                    const int i = lane + 32 * t;
                    const double u = gpr[t];
                    gpr[t] = cs * u - sn * w[t];
                    a_new = fma(gpr[t], gpr[t], a_new);
                    if (i < d) {
                      gq[i] = sn * u + cs * w[t];
                      const double vu = vpr[t], vw = vq[i];
                      vpr[t] = cs * vu - sn * vw;
                      vq[i] = sn * vu + cs * vw;
                    }
                  }
                  a = warp_sum(a_new);
                }
              }
              __syncthreads();
            }
            if (have_p) {
#pragma unroll
              for (int t = 0; t < kPsdEpl; ++t) {
                const int i = lane + 32 * t;
                if (i < d) {
                  Gs[(size_t)k * d + i] = gpr[t];
                  Vs[(size_t)k * d + i] = vpr[t];
                }
              }
            }
          }
          __syncthreads();
        } else {
          for (int rnd = 0; rnd < nmax; ++rnd) {
            for (int k = wid; k < nmax; k += nw) {
              const int p = k, q = (k + rnd) % nmax;
              if (p >= na || q >= nbb) continue;
              rot = max(rot, jacobi_pair_smem(Gs + (size_t)p * d, Gs + (size_t)(na + q) * d, Vs + (size_t)p * d,
                                              Vs + (size_t)(na + q) * d, d, lane, tol2));
            }
            __syncthreads();
          }
        }
        if (rot && lane == 0) atomicMax(&sh_rot, rot);
        for (int idx = tid; idx < nloc * d; idx += nthr) {
          const int c = idx / d, i = idx - c * d;
          const long long dst = (long long)(c < na ? a0 + c : b0 + (c - na)) * d + i;
          __stcg(G + dst, Gs[idx]);
          __stcg(V + dst, Vs[idx]);
        }
        __syncthreads();
      }
      sync_all();
    }
    int rotated = sh_rot;
    if (C > 1) {
      cg::cluster_group cl = cg::this_cluster();
      for (int r = 0; r < C; ++r) rotated = max(rotated, *cl.map_shared_rank(&sh_rot, r));
    }
    sync_all();  // everyone has read the flags before they are cleared
    if (rotated < 2) break;  // no rotation at all, or only pairs that were already orthogonal to kPsdSmallCos
  }
  if (tid == 0 && rank == 0) state[cone].sweeps = sweep + 1;
}

// lambda_k = v_k . g_k - sigma   (grid: ceil(max_d / 8) x cones, 256 threads, one warp per k)
__global__ void __launch_bounds__(256)
k_psd_lambda(const PsdEntry *__restrict__ ents, const PsdState *__restrict__ state, const double *__restrict__ Gall,
             const double *__restrict__ Vall, double *__restrict__ lamall) {
  const PsdEntry e = ents[blockIdx.y];
  const int d = e.d;
  if (e.s < 2) return;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= d) return;
  const double *gk = Gall + e.woff + (long long)k * d, *vk = Vall + e.woff + (long long)k * d;
  double acc = 0.0;
  for (int i = lane; i < d; i += 32) acc = fma(gk[i], vk[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) lamall[e.loff + k] = acc - state[blockIdx.y].sigma;
}

// X+ = sum_{lambda>0} lambda v v' on the needed triangle, re-packed (diagonal / sqrt2) with the Moreau
// recombination x <- X+ / r + s fused   (grid: tiles*tiles x cones)
__global__ void __launch_bounds__(256)
k_psd_recon(double *__restrict__ xall, const double *__restrict__ svall, const double *__restrict__ rall,
            const PsdEntry *__restrict__ ents, PsdState *__restrict__ state, const double *__restrict__ Vall,
            const double *__restrict__ lamall, int tiles) {
  const PsdEntry e = ents[blockIdx.y];
  const int n = e.s, d = e.d;
  if (n < 2) return;
  const int ti = blockIdx.x % tiles, tj = blockIdx.x / tiles;
  const int i0 = ti * kGT, j0 = tj * kGT;
  if (blockIdx.x == 0 && threadIdx.x == 0) state[blockIdx.y].age += 1;
  if (i0 >= d || j0 >= n) return;          // columns: only the leading n are ever needed
  if (!e.is_complex && ti < tj) return;    // strictly-upper tiles of the real case
  const double *V = Vall + e.woff, *lam = lamall + e.loff;
  double acc[4][4] = {};
  tile_gemm(acc, d,
            [&](int k, int i) { return (k < d && i0 + i < d) ? V[(long long)k * d + i0 + i] : 0.0; },
            [&](int k, int j) {
              if (k >= d || j0 + j >= n) return 0.0;
              const double lk = lam[k];
              return lk > 0.0 ? lk * V[(long long)k * d + j0 + j] : 0.0;
            }, false);
  double *x = xall + e.off;
  const double *sv = svall + e.off, *r = rall + e.off;
  const double isqrt2 = 0.7071067811865476;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int b = 0; b < 4; ++b)
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
      if (i >= d || j >= n) continue;
      if (!e.is_complex) {
        if (i < j) continue;
        const long long pi = tri_idx(i, j, n);
        const double val = (i == j) ? acc[a][b] * isqrt2 : acc[a][b];
        x[pi] = val / r[pi] + sv[pi];
      } else {
        const long long base = (long long)j * (2 * n - j);
        if (i < n) {          // A block: real parts
          if (i < j) continue;
          if (i == j) x[base] = (acc[a][b] * isqrt2) / r[base] + sv[base];
          else { const long long pr = base + 1 + 2 * (i - j - 1); x[pr] = acc[a][b] / r[pr] + sv[pr]; }
        } else {              // B block (rows n..2n-1): imaginary parts, strictly below the diagonal
          const int ii = i - n;
          if (ii <= j) continue;
          const long long pim = base + 2 + 2 * (ii - j - 1);
          x[pim] = acc[a][b] / r[pim] + sv[pim];
        }
      }
    }
}

// ============================================== cone-boundary aggregation (setup) =====
__global__ void __launch_bounds__(kThreads)
k_enforce_boundaries(double *__restrict__ D, const int *__restrict__ off, const int *__restrict__ len, int ncones,
                     int use_mean) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int cidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; cidx < ncones; cidx += warps_per_grid) {
    double *v = D + off[cidx];
    const int n = len[cidx];
    double acc = 0.0;
    if (use_mean) {
      for (int k = lane; k < n; k += 32) acc += v[k];
      acc = warp_sum(acc) / (double)n;   // SCS(mean), linalg.c
    } else {
      for (int k = lane; k < n; k += 32) acc = fmax(acc, fabs(v[k]));
      acc = warp_max(acc);               // SCS(norm_inf)
    }
    for (int k = lane; k < n; k += 32) v[k] = acc;
  }
}

// ============================================================================ host ====
int ConeDev::init(Ctx *ctx, const ScsCone *k, int m_) {
  c = ctx;
  m = m_;
  z = k->z; l = k->l; bsize = k->bsize; ep = k->ep; ed = k->ed;
  q.assign(k->q, k->q + (k->qsize > 0 ? k->qsize : 0));
  s.assign(k->s, k->s + (k->ssize > 0 ? k->ssize : 0));
  cs.assign(k->cs, k->cs + (k->cssize > 0 ? k->cssize : 0));
  p.assign(k->p, k->p + (k->psize > 0 ? k->psize : 0));
  if (bsize > 1) {
    bu.assign(k->bu, k->bu + bsize - 1);
    bl.assign(k->bl, k->bl + bsize - 1);
  }
  int cnt = z + l;
  off_box = cnt; cnt += bsize;
  off_q = cnt;
  std::vector<int> qo_s, ql_s, qo_l, ql_l, bo, blen;
  boundaries.clear();
  boundaries.push_back(z + l + bsize);
  for (int qi : q) {
    if (qi > 2048) { qo_l.push_back(cnt); ql_l.push_back(qi); }
    else { qo_s.push_back(cnt); ql_s.push_back(qi); }
    if (qi > 0) { bo.push_back(cnt); blen.push_back(qi); }
    boundaries.push_back(qi);
    cnt += qi;
  }
  off_s = cnt;
  std::vector<PsdEntry> ents;
  long long woff = 0;
  int loff = 0;
  psd_max_d = 0;
  for (int si : s) {
    PsdEntry e{cnt, si, si, 0, woff, loff, 0};
    ents.push_back(e);
    const int sz = si * (si + 1) / 2;
    if (sz > 0) { bo.push_back(cnt); blen.push_back(sz); }
    boundaries.push_back(sz);
    woff += (long long)si * si; loff += si; cnt += sz;
    if (si > psd_max_d) psd_max_d = si;
  }
  off_cs = cnt;
  for (int ci : cs) {
    PsdEntry e{cnt, ci, 2 * ci, 1, woff, loff, 0};
    ents.push_back(e);
    const int sz = ci * ci;
    if (sz > 0) { bo.push_back(cnt); blen.push_back(sz); }
    boundaries.push_back(sz);
    woff += 4ll * ci * ci; loff += 2 * ci; cnt += sz;
    if (2 * ci > psd_max_d) psd_max_d = 2 * ci;
  }
  off_exp = cnt;
  for (int i = 0; i < ep + ed; ++i) { bo.push_back(cnt); blen.push_back(3); boundaries.push_back(3); cnt += 3; }
  off_pow = cnt;
  for (size_t i = 0; i < p.size(); ++i) { bo.push_back(cnt); blen.push_back(3); boundaries.push_back(3); cnt += 3; }
  if (cnt != m) {
    B200_PRINTF("Error: Cone dims %li != rows in A %li\n", (long)cnt, (long)m);
    return -1;
  }
  CUDA_OK(cudaSetDevice(c->device));
  // SOC lists: small first, then large
  n_q_small = (int)qo_s.size(); n_q_large = (int)qo_l.size();
  std::vector<int> qo(qo_s), ql(ql_s);
  qo.insert(qo.end(), qo_l.begin(), qo_l.end());
  ql.insert(ql.end(), ql_l.begin(), ql_l.end());
  if (!qo.empty()) {
    if (dev_alloc(&q_off, qo.size()) || dev_alloc(&q_len, ql.size()) || h2d(*c, q_off, qo.data(), qo.size()) ||
        h2d(*c, q_len, ql.data(), ql.size()))
      return -1;
  }
  n_psd = (int)ents.size();
  if (n_psd) {
    // small cones (one CTA each) first, then the cluster-per-cone ones
    std::vector<PsdEntry> small, large;
    for (const PsdEntry &e : ents) (e.d <= kPsdSmallDim ? small : large).push_back(e);
    n_psd_small = (int)small.size();
    psd_small_max_d = psd_large_max_d = 0;
    for (const PsdEntry &e : small) psd_small_max_d = e.d > psd_small_max_d ? e.d : psd_small_max_d;
    for (const PsdEntry &e : large) psd_large_max_d = e.d > psd_large_max_d ? e.d : psd_large_max_d;
    {  // launch geometry of the two Jacobi launches: maxima over the cones each one covers
      const size_t budget = (size_t)200 * 1024;
      auto geom = [&](const std::vector<PsdEntry> &v, int ctas, size_t &smem, int &warps) {
        smem = 0; warps = 1;
        for (const PsdEntry &e : v) {
          if (e.s < 2) continue;
          const int NB = psd_num_blocks(e.d, ctas, budget);
          const int nb = (e.d + NB - 1) / NB;
          const size_t sm = (size_t)32 * nb * e.d;
          smem = sm > smem ? sm : smem;
          const int w = nb < 1 ? 1 : (nb > kPsdJacMaxWarps ? kPsdJacMaxWarps : nb);  // nloc / 2 pairs per local round
          warps = w > warps ? w : warps;
        }
      };
      geom(small, 1, psd_small_smem, psd_small_warps);
      geom(large, kPsdCluster, psd_large_smem, psd_large_warps);
    }
    ents = small;
    ents.insert(ents.end(), large.begin(), large.end());
    psd_tiles = (psd_max_d + kGT - 1) / kGT;
    CUDA_OK(cudaFuncSetAttribute(k_psd_jacobi<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_OK(cudaFuncSetAttribute(k_psd_jacobi<kPsdCluster>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev_alloc(&psd, ents.size()) || h2d(*c, psd, ents.data(), ents.size()) ||
        dev_alloc(&psd_W, (size_t)woff) || dev_alloc(&psd_G, (size_t)woff) || dev_alloc(&psd_V, (size_t)woff) ||
        dev_alloc(&psd_lam, (size_t)loff + 1) || dev_alloc_zero(&psd_state, ents.size(), c->stream))
      return -1;
  }
  if (bsize > 0) {
    const double one = 1.0;
    if (dev_alloc(&box_t, 1) || h2d(*c, box_t, &one, 1)) return -1;
    if (dev_alloc(&d_bu, (size_t)(bsize > 1 ? bsize - 1 : 1)) || dev_alloc(&d_bl, (size_t)(bsize > 1 ? bsize - 1 : 1)))
      return -1;
  }
  if (!p.empty()) {
    if (dev_alloc(&d_p, p.size()) || h2d(*c, d_p, p.data(), p.size())) return -1;
  }
  n_bnd = (int)bo.size();
  if (n_bnd) {
    if (dev_alloc(&bnd_off, bo.size()) || dev_alloc(&bnd_len, blen.size()) || h2d(*c, bnd_off, bo.data(), bo.size()) ||
        h2d(*c, bnd_len, blen.data(), blen.size()))
      return -1;
  }
  if (c->sync()) return -1;
  return 0;
}

void ConeDev::destroy() {
  if (!c) return;
  cudaSetDevice(c->device);
  dev_free(q_off); dev_free(q_len); dev_free(psd); dev_free(psd_W); dev_free(psd_G); dev_free(psd_V); dev_free(psd_lam);
  dev_free(psd_state);
  dev_free(d_bu); dev_free(d_bl); dev_free(box_t); dev_free(d_p); dev_free(bnd_off); dev_free(bnd_len);
}

int ConeDev::normalize_box(const double *D_host) {
  // normalize_box_cone, cones.c:1153-1169 (runs once: scaled_cones latch, cones.c:1549-1557)
  if (scaled_cones) return 0;
  scaled_cones = true;
  if (bsize <= 1) return 0;
  if (D_host) {  // only when a scaling exists (cones.c:1553-1554); otherwise bounds stay as given
    const double *Db = D_host + z + l;
    for (int j = 0; j < bsize - 1; ++j) {
      const double factor = Db[j + 1] / Db[0];
      bu[j] = (bu[j] >= 1e15) ? INFINITY : bu[j] * factor;
      bl[j] = (bl[j] <= -1e15) ? -INFINITY : bl[j] * factor;
    }
  }
  const double one = 1.0;  // box_t_warm_start = 1, cones.c:1552
  if (h2d(*c, d_bu, bu.data(), (size_t)bsize - 1) || h2d(*c, d_bl, bl.data(), (size_t)bsize - 1) ||
      h2d(*c, box_t, &one, 1) || c->sync())
    return -1;
  return 0;
}

int ConeDev::project_nonlinear(double *x, const double *sv, const double *r) {
  cudaStream_t st = c->stream;
  if (bsize > 0) {
    k_box_cone<<<1, kBoxThreads, 0, st>>>(x + off_box, sv + off_box, r + off_box, d_bl, d_bu, bsize, box_t);
    c->launches++;
  }
  if (n_q_small > 0) {
    int grid = (n_q_small + (kThreads / 32) - 1) / (kThreads / 32);
    if (grid > c->grid_ew()) grid = c->grid_ew();
    k_soc_small<<<grid, kThreads, 0, st>>>(x, sv, r, q_off, q_len, n_q_small);
    c->launches++;
  }
  if (n_q_large > 0) {
    int grid = n_q_large < c->grid_ew() ? n_q_large : c->grid_ew();
    k_soc_large<<<grid, kThreads, 0, st>>>(x, sv, r, q_off + n_q_small, q_len + n_q_small, n_q_large);
    c->launches++;
  }
  if (n_psd > 0) {
    const dim3 gtiles((unsigned)(psd_tiles * psd_tiles), (unsigned)n_psd);
    k_psd_prep<<<n_psd, kPsdPrepThreads, 0, st>>>(x, sv, r, psd, psd_state, psd_W, psd_G, psd_V);
    k_psd_gemm_wv<<<gtiles, 256, 0, st>>>(psd, psd_state, psd_W, psd_V, psd_G, psd_tiles);
    c->launches += 2;
    // shared memory: two blocks of (G, V) columns per CTA.  Every cone of a launch recomputes its own
    // block count from its own d inside the kernel, and the footprint 32 nb(d) d is not monotone in d
    // (s = [225, 201]: 165600 vs 167232 bytes), so the launch is sized by the maximum over the cones it
    // covers, not by the largest d.
    const size_t budget = (size_t)200 * 1024;
    auto smem_for = [&](int, int ctas) { return ctas == 1 ? psd_small_smem : psd_large_smem; };
    auto threads_for = [&](int, int ctas) { return 32 * (ctas == 1 ? psd_small_warps : psd_large_warps); };
    if (n_psd_small > 0) {
      const size_t sm = smem_for(psd_small_max_d, 1);
      k_psd_jacobi<1><<<n_psd_small, threads_for(psd_small_max_d, 1), sm, st>>>(psd, psd_state, 0, psd_G, psd_V, budget);
      c->launches++;
    }
    if (n_psd > n_psd_small) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((n_psd - n_psd_small) * kPsdCluster));
      cfg.blockDim = dim3((unsigned)threads_for(psd_large_max_d, kPsdCluster));
      cfg.dynamicSmemBytes = smem_for(psd_large_max_d, kPsdCluster);
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kPsdCluster;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CUDA_OK(cudaLaunchKernelEx(&cfg, k_psd_jacobi<kPsdCluster>, (const PsdEntry *)psd, psd_state, n_psd_small, psd_G, psd_V,
                                 budget));
      c->launches++;
    }
    k_psd_lambda<<<dim3((unsigned)((psd_max_d + 7) / 8), (unsigned)n_psd), 256, 0, st>>>(psd, psd_state, psd_G, psd_V, psd_lam);
    k_psd_recon<<<gtiles, 256, 0, st>>>(x, sv, r, psd, psd_state, psd_V, psd_lam, psd_tiles);
    c->launches += 2;
  }
  if (ep + ed > 0) {
    const int nt = ep + ed;
    int grid = (nt + 127) / 128;
    if (grid > c->grid_ew()) grid = c->grid_ew();
    k_exp_cones<<<grid, 128, 0, st>>>(x + off_exp, sv + off_exp, r + off_exp, ep, nt);
    c->launches++;
  }
  if (!p.empty()) {
    const int nt = (int)p.size();
    int grid = (nt + 127) / 128;
    if (grid > c->grid_ew()) grid = c->grid_ew();
    k_pow_cones<<<grid, 128, 0, st>>>(x + off_pow, sv + off_pow, r + off_pow, d_p, nt);
    c->launches++;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}

int ConeDev::enforce_boundaries(double *D, int use_mean) {
  if (n_bnd == 0) return 0;
  int grid = (n_bnd + (kThreads / 32) - 1) / (kThreads / 32);
  if (grid > c->grid_ew()) grid = c->grid_ew();
  k_enforce_boundaries<<<grid, kThreads, 0, c->stream>>>(D, bnd_off, bnd_len, n_bnd, use_mean);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  return 0;
}

// generic Moreau pre-pass used by the host-buffer test surface: sv = x ; rows < z+l are
// finished in place, the rest become -r*x.
__global__ void __launch_bounds__(kThreads)
k_cone_pre(double *__restrict__ x, double *__restrict__ sv, const double *__restrict__ r, int m, int z, int zl) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
    const double s = x[i], ri = r[i];
    sv[i] = s;
    x[i] = (i < zl) ? zl_moreau(i, z, s, ri) : -ri * s;
  }
}

int current_device();

}  // namespace b200

// ===================================================== C ABI: SCS(proj_dual_cone) =====
using namespace b200;

struct SCS_B200_CONE_WORK {
  Ctx ctx;
  ConeDev cone;
  double *x = nullptr, *sv = nullptr, *r = nullptr;
};

extern "C" ScsB200ConeWork *scs_b200_init_cone(const ScsCone *k, scs_int m) {
  if (!k || m <= 0) return nullptr;
  SCS_B200_CONE_WORK *w = new SCS_B200_CONE_WORK();
  if (w->ctx.init(current_device()) || w->cone.init(&w->ctx, k, m) || dev_alloc(&w->x, (size_t)m) ||
      dev_alloc(&w->sv, (size_t)m) || dev_alloc(&w->r, (size_t)m)) {
    scs_b200_finish_cone(w);
    return nullptr;
  }
  return w;
}

extern "C" scs_int scs_b200_proj_dual_cone(scs_float *x, ScsB200ConeWork *w, const scs_float *D,
                                           const scs_float *r_y) {
  if (!w || !x) return -1;
  Ctx &c = w->ctx;
  if (cudaSetDevice(c.device) != cudaSuccess) return -1;
  const int m = w->cone.m;
  if (w->cone.normalize_box(D)) return -1;
  std::vector<double> ones;
  if (!r_y) { ones.assign((size_t)m, 1.0); r_y = ones.data(); }
  if (h2d(c, w->x, x, (size_t)m) || h2d(c, w->r, r_y, (size_t)m)) return -1;
  int grid = (m + kThreads - 1) / kThreads;
  if (grid > c.grid_ew()) grid = c.grid_ew();
  k_cone_pre<<<grid, kThreads, 0, c.stream>>>(w->x, w->sv, w->r, m, w->cone.z, w->cone.z + w->cone.l);
  c.launches++;
  if (w->cone.project_nonlinear(w->x, w->sv, w->r)) return -1;
  if (d2h(c, x, w->x, (size_t)m) || c.sync()) return -1;
  return 0;
}

__global__ void __launch_bounds__(kThreads)
k_axpy_set(double *__restrict__ x, const double *__restrict__ x0, const double *__restrict__ x1, double t, int m) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) x[i] = fma(t, x1[i], x0[i]);
}

// Measurement hook: `reps` projections of x_r = x0 + r*step*x1 (a slowly drifting input, as the
// ADMM iterates are), all on the device; returns the mean ms per projection (CUDA events on the
// workspace stream) and the mean number of Jacobi sweeps of the PSD cones in *sweeps_out.
extern "C" double scs_b200_bench_proj_cone(ScsB200ConeWork *w, const scs_float *x0, const scs_float *x1,
                                           scs_float step, scs_int reps, scs_int warmup, double *sweeps_out) {
  if (!w || !x0 || !x1 || reps <= 0) return -1.0;
  Ctx &c = w->ctx;
  if (cudaSetDevice(c.device) != cudaSuccess) return -1.0;
  const int m = w->cone.m;
  double *d0 = nullptr, *d1 = nullptr;
  if (dev_alloc(&d0, (size_t)m) || dev_alloc(&d1, (size_t)m) || h2d(c, d0, x0, (size_t)m) || h2d(c, d1, x1, (size_t)m))
    return -1.0;
  if (w->cone.normalize_box(nullptr)) return -1.0;
  std::vector<double> ones((size_t)m, 1.0);
  if (h2d(c, w->r, ones.data(), (size_t)m) || c.sync()) return -1.0;
  int grid = (m + kThreads - 1) / kThreads;
  if (grid > c.grid_ew()) grid = c.grid_ew();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double sw_sum = 0.0; long long sw_cnt = 0;
  std::vector<PsdState> hs((size_t)(w->cone.n_psd > 0 ? w->cone.n_psd : 1));
  float total = 0.f;
  for (int r = 0; r < warmup + reps; ++r) {
    k_axpy_set<<<grid, kThreads, 0, c.stream>>>(w->x, d0, d1, r * step, m);
    if (r >= warmup) cudaEventRecord(e0, c.stream);
    k_cone_pre<<<grid, kThreads, 0, c.stream>>>(w->x, w->sv, w->r, m, w->cone.z, w->cone.z + w->cone.l);
    if (w->cone.project_nonlinear(w->x, w->sv, w->r)) return -1.0;
    if (r >= warmup) {
      cudaEventRecord(e1, c.stream);
      if (cudaStreamSynchronize(c.stream) != cudaSuccess) return -1.0;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      total += ms;
      if (w->cone.n_psd > 0) {
        cudaMemcpy(hs.data(), w->cone.psd_state, sizeof(PsdState) * (size_t)w->cone.n_psd, cudaMemcpyDeviceToHost);
        for (int k = 0; k < w->cone.n_psd; ++k) { sw_sum += hs[(size_t)k].sweeps; sw_cnt++; }
      }
    }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  dev_free(d0); dev_free(d1);
  if (sweeps_out) *sweeps_out = sw_cnt ? sw_sum / (double)sw_cnt : 0.0;
  return (double)total / reps;
}

extern "C" void scs_b200_finish_cone(ScsB200ConeWork *w) {
  if (!w) return;
  cudaSetDevice(w->ctx.device);
  if (w->ctx.stream) cudaStreamSynchronize(w->ctx.stream);
  w->cone.destroy();
  dev_free(w->x); dev_free(w->sv); dev_free(w->r);
  w->ctx.destroy();
  delete w;
}
