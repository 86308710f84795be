// dist.cuh -- NCCL plumbing for the row-partitioned single-problem mode (one process per GPU).
//
// No reference counterpart (the reference is single-device).  Rank g owns a
// contiguous block of rows of A cut on cone boundaries, the matching slices of every m-space
// vector and the cones inside it.  Columns of A whose non-zeros all lie in one rank's rows (and
// whose row of P is diagonal) are PRIVATE to that rank: their entries of every n-space vector
// exist only there.  The other columns are SHARED: replicated on every rank with bit-identical
// values.  The data-path exchanges are one sum all-reduce of the shared block of A_g' z_g (+ one
// piggy-backed scalar) per CG iteration and one gather of a few reduction scalars after every
// reducing kernel (dist_finish).  NCCL is resolved with dlopen at first use so that single-GPU
// users never need the library.
#pragma once
#include "common.cuh"

namespace b200 {

struct Dist {
  int rank = 0, world = 1;
  void *comm = nullptr;  // ncclComm_t
  // world == 1 with SCS_B200_DIST_SELFTEST=1 in the environment at scs_b200_dist_init: workspaces still take the
  // row-partitioned code path (column classification, deferred finalisers, split products) with the
  // collectives degenerated to copies -- the single-GPU test of that path
  bool selftest = false;
  bool on() const { return world > 1 && comm != nullptr; }
};

Dist *dist_current();  // process-wide communicator state (world == 1 when not initialised)
// op: 0 = sum, 1 = max.  In place, FP64, on stream c.stream.  No-op when !dist.on().
int dist_allreduce(Ctx &c, double *buf, size_t count, int op);
// out-of-place sum all-reduce (send != recv); with world == 1 (self-test mode) a device copy
int dist_allreduce_oop(Ctx &c, const double *send, double *recv, size_t count);

#ifdef __CUDACC__
// Deferred finaliser of a grid_reduce_fin (common.cuh): all-reduce the ns sums and nm maxes the
// kernel parked in S->part, then run the formula on every rank.  No-op on a single GPU.
template <class Fin>
inline int dist_finish(Ctx &c, int ns, int nm, Fin fin) {
  if (!c.dist) return 0;
  if (dist_allreduce_oop(c, c.S->gsend, c.S->grecv, (size_t)c.world * (size_t)(ns + nm))) return -1;
  k_apply_fin<<<1, 32, 0, c.stream>>>(fin, c.S, ns, nm);
  c.launches++;
  return 0;
}
#endif

}  // namespace b200
