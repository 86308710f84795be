// linsys.cuh -- device-resident quasi-definite KKT solve by preconditioned CG.
//
// Replaces S/linsys/cpu/indirect/private.c (the algorithmic template) and is NOT derived
// from S/linsys/gpu/ (cuSPARSE/cuBLAS calls with a host sync per dot product).
//
//   [R_x + P   A' ] [x]   [rx]        x = (R_x + P + A' R_y^-1 A)^-1 (rx + A' R_y^-1 ry)
//   [  A     -R_y ] [y] = [ry]   =>   y = R_y^-1 (A x - ry)                (private.c:266-275)
//
// All CG scalars (alpha, beta, z'r, ||r||_inf, the stop flag) live in DevScalars; kernels
// launched after convergence return immediately, so the host only reads one flag per batch
// of enqueued CG iterations.
#pragma once
#include "sparse.cuh"
#include "tiled.cuh"

namespace b200 {

// loop control of the CG WHILE node (zero-initialised == stream path, no graph)
struct CgCtl {
  cudaGraphConditionalHandle h = 0;
  int use_h = 0;
};

struct LinSys {
  Ctx *c = nullptr;
  int n = 0, m = 0;
  CsrDev A, At, P;  // CSR(A): m rows; CSR(A'): n rows; full symmetric CSR(P): n rows
  bool hasP = false;
  ChunkList chA, chAt;   // chAt is built over the fused (A', P) rows
  // row-partitioned mode: chunk list over the n rows of P alone (P p + R_x p is added after the
  // all-reduce of A_g' z_g).  Every data-dependent scalar is produced by dist_finish from the ranks' raw
  // values in rank order, so the CG scalars and hence the ranks' control flow are bit-identical.
  ChunkList chP;
  // tiled shared-memory format (tiled.cuh) of CSR(A) and of [CSR(A') | CSR(P)], built by
  // finalize_structure() for large matrices; the CG-loop products go through it when present
  TiledOp tA, tG;
  double *diag_r = nullptr;  // n+m(+1) on device; owned iff own_diag_r
  bool own_diag_r = false;
  double *Pdiag = nullptr;  // n (zeros when !hasP)
  double *M = nullptr, *p = nullptr, *r = nullptr, *Gp = nullptr, *z = nullptr;  // n
  static constexpr int kGpFront = 8;  // doubles allocated in front of Gp (Gp[-1]: all-reduced scalar, dist mode)
  double *Gp_base = nullptr;
  bool Gp_in_arena = false;  // Gp_base lives in the peer-mapped arena of the row-partitioned mode (dist.cu)
  double *pp = nullptr;  // n, row-partitioned mode only: P p + R_x p
  double *tmp = nullptr;                                                          // m
  long long tot_cg_its = 0;
  int last_its = 2;

  // A (m x n) and P (n x n upper, may be null) are host CSC matrices.
  int init(Ctx *ctx, const ScsMatrix *Ah, const ScsMatrix *Ph);
  // build the tiled SpMV format; call after the matrices have their final (equilibrated) values
  int finalize_structure();
  void destroy();
  int update_precond();  // M = 1 / diag(R_x + P + A' R_y^-1 A)   (private.c:50-84)
  // b = [rx; ry] (device, n+m) -> [x; y] in place.  ws: warm start for x (device, n) or
  // null.  Before the call the caller must have set S->cg_tol, S->zero_rhs and
  // S->cg_done (= zero_rhs).  first_batch <= 0 picks the adaptive default.
  int solve_dev(double *b, const double *ws, int first_batch);
  // the three pieces of solve_dev, enqueue-only (no host synchronisation): used directly when
  // the ADMM iteration is captured into a CUDA graph whose CG loop is a WHILE node
  int enqueue_head(double *b, const double *ws, CgCtl ctl);  // tmp = ry/R_y ; b_x += A' tmp ; CG start
  int enqueue_cg_iter(double *b, CgCtl ctl, int tag);         // one CG iteration (4 kernels)
  int enqueue_tail(double *b);                                // y recovery / zero right-hand side
  int solve_dev_loop(double *b, int first_batch);             // host-driven CG loop between head and tail
  // sets S->cg_tol = tol and the zero-rhs flags from ||b||_inf (device reduction)
  int prepare_flags(const double *b, double tol);
  // algorithmic bytes
  double bytes_A() const { return 12.0 * A.nnz + 4.0 * (m + 1) + 8.0 * n + 8.0 * m; }
  double bytes_At() const { return 12.0 * At.nnz + 4.0 * (n + 1) + 8.0 * m + 8.0 * n; }
  // kernels of one CG iteration (enqueue_cg_iter): two products + two vector updates; a tiled product is
  // two launches (streaming kernel, epilogue pass)
  static int tiled_launches(const TiledOp &t) { return t.ok ? (t.has_tiled ? 2 : 1) : 1; }
  int cg_iter_launches() const { return (c && c->dist ? 4 : 2) + tiled_launches(tA) + tiled_launches(tG); }
  double bytes_P() const { return hasP ? 12.0 * P.nnz + 4.0 * (n + 1) + 16.0 * n : 0.0; }
  // row-partitioned mode (dist.cuh)
  int launch_A_scaled_dot(const double *x, double *out, const int *skip);
  int dist_At(const double *zin, double *out, const int *skip, bool with_scalar, int kt_cat);
  int dist_reduce_gp(double *out, bool with_scalar, const int *skip);
  // launchers shared with the ADMM driver
  int launch_A_scaled(const double *x, double *out, const int *skip, int tag = -1, bool counted = true);  // out = R_y^-1 A x
  int launch_G(const double *zin, const double *pin, double *out, const int *skip, int tag = -1,
               bool counted = true);  // out = A'z + P p + R_x p ; S->alpha
};

}  // namespace b200
