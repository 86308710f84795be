// dist.cuh -- NCCL plumbing for the row-partitioned single-problem mode (one process per GPU).
//
// No reference counterpart (the reference is single-device).  Rank g owns a
// contiguous block of rows of A cut on cone boundaries, the matching slices of every m-space
// vector and the cones inside it; n-space vectors and all scalars are replicated.  The only
// data-path exchanges are sum all-reduces of n-vectors (the A_g' z_g partial products, once
// per CG iteration) and of a handful of reduction scalars per ADMM iteration.  NCCL is resolved
// with dlopen at first use so that single-GPU users never need the library.
#pragma once
#include "common.cuh"

namespace b200 {

struct Dist {
  int rank = 0, world = 1;
  void *comm = nullptr;  // ncclComm_t
  bool on() const { return world > 1 && comm != nullptr; }
};

Dist *dist_current();  // process-wide communicator state (world == 1 when not initialised)
// op: 0 = sum, 1 = max.  In place, FP64, on stream c.stream.  No-op when !dist.on().
int dist_allreduce(Ctx &c, double *buf, size_t count, int op);

#ifdef __CUDACC__
// Deferred finaliser of a grid_reduce_fin (common.cuh): all-reduce the ns sums and nm maxes the
// kernel parked in S->part, then run the formula on every rank.  No-op on a single GPU.
template <class Fin>
inline int dist_finish(Ctx &c, int ns, int nm, Fin fin) {
  if (!c.dist) return 0;
  if (ns && dist_allreduce(c, c.S->part, (size_t)ns, 0)) return -1;
  if (nm && dist_allreduce(c, c.S->part + ns, (size_t)nm, 1)) return -1;
  k_apply_fin<<<1, 32, 0, c.stream>>>(fin, c.S);
  c.launches++;
  return 0;
}
#endif

}  // namespace b200
