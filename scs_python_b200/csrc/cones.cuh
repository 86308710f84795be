// cones.cuh -- device-resident cone projections (replaces SCS(proj_dual_cone) and the
// per-cone projections of S/src/cones.c:986-1588 and S/src/exp_cone.c).
//
// Data layout: the m rows of the dual block follow the reference's fixed cone order
// (scs.h:122-173): zero | nonneg | box | SOC... | PSD... | complex PSD... | exp primal |
// exp dual | power.  One kernel per cone family works in place on a device vector x that
// already holds -R_y * s (the Moreau pre-image, cones.c:1562-1571) and finishes with the
// Moreau recombination x <- Pi_K(x) / R_y + s (cones.c:1576-1585), s being read from a
// saved copy.  Zero and nonneg rows are pure elementwise work and are folded into the
// caller's elementwise pass (zl_moreau below).
#pragma once
#include "common.cuh"

namespace b200 {

struct PsdEntry {
  int off;        // offset of the cone inside the y block
  int s;          // matrix dimension as given (s or cs)
  int d;          // working dimension of the real symmetric problem (s, or 2*cs)
  int is_complex;
  long long woff; // offset (in doubles) of this cone's d*d blocks inside W, G and V
  int loff;       // offset into the eigenvalue scratch
  int pad;
};
// per-cone state of the warm-started eigen-solver (device)
struct PsdState {
  int age;        // projections since the eigenvector basis was last rebuilt from the identity
  int cold;       // this projection starts from V = I (set by the prep kernel)
  int sweeps;     // Jacobi sweeps of the last projection (diagnostic)
  int pad;
  double sigma;   // spectral shift: W = mat(x) + sigma I is PSD
};

struct ConeDev {
  Ctx *c = nullptr;
  int m = 0;
  // host description (deep copy of ScsCone)
  int z = 0, l = 0, bsize = 0, ep = 0, ed = 0;
  std::vector<int> q, s, cs;
  std::vector<double> p, bu, bl;
  int off_box = 0, off_q = 0, off_s = 0, off_cs = 0, off_exp = 0, off_pow = 0;
  bool scaled_cones = false;  // box bounds normalised by D once (cones.c:1549-1557)
  std::vector<int> boundaries;  // set_cone_boundaries, cones.c:386-424
  // device
  int *q_off = nullptr, *q_len = nullptr;  // small SOCs (warp each), then large (CTA each)
  int n_q_small = 0, n_q_large = 0;
  PsdEntry *psd = nullptr;
  int n_psd = 0, psd_max_d = 0;
  int n_psd_small = 0;  // entries [0, n_psd_small): one CTA each; the rest: one 4-CTA cluster each
  double *psd_W = nullptr, *psd_G = nullptr, *psd_V = nullptr, *psd_lam = nullptr;
  PsdState *psd_state = nullptr;
  int psd_tiles = 0;    // ceil(psd_max_d / 64)
  int psd_small_max_d = 0, psd_large_max_d = 0;
  size_t psd_small_smem = 0, psd_large_smem = 0;  // dynamic shared memory of the two Jacobi launches (max over their cones)
  int psd_small_warps = 1, psd_large_warps = 1;
  double *d_bu = nullptr, *d_bl = nullptr, *box_t = nullptr;
  double *d_p = nullptr;
  int *bnd_off = nullptr, *bnd_len = nullptr;  // cones of size > 1 for enforce_cone_boundaries
  int n_bnd = 0;
  int *status = nullptr;  // device int, set <0 by a failing kernel

  int init(Ctx *ctx, const ScsCone *k, int m);
  void destroy();
  // Upload the box bounds, normalised by D (host array of length m, or null).
  int normalize_box(const double *D_host);
  // x: device, m entries holding -R_y*s for all rows >= z+l (rows < z+l are left alone);
  // s_saved, r_y: device, m entries.  In place: x <- Pi_K(x)/r_y + s_saved on those rows.
  int project_nonlinear(double *x, const double *s_saved, const double *r_y);
  // D[cone] <- max or mean over every cone of size > 1 (cones.c:366-379)
  int enforce_boundaries(double *D, int use_mean);
  bool has_nonlinear() const { return m > z + l; }
};

#ifdef __CUDACC__
// zero / nonneg rows of the dual-cone projection, fully fused (cones.c:1341-1351 wrapped
// by the Moreau steps of cones.c:1562-1585):  returns Pi_{K*}^R(s) for row `row`.
__device__ __forceinline__ double zl_moreau(int row, int z, double s, double r) {
  if (row < z) return s;          // Pi_{0}(.) = 0  ->  0 / r + s
  const double t = fmax(-r * s, 0.0);
  return t / r + s;
}
#endif

}  // namespace b200
