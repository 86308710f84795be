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

// Peer-memory path.  dist_p2p_setup (collective: every rank, same order) allocates ONE local arena
// [PeerSync | vec_doubles doubles], exchanges its CUDA IPC handle over the NCCL communicator and maps the other
// ranks' arenas; *vec_out = the local vector region.  Returns 0 with c.p2p == false when peer mapping is not
// possible (the caller then allocates the vector itself and every collective stays on NCCL).
// SCS_B200_DIST_P2P=0 disables it.
int dist_p2p_setup(Ctx &c, size_t vec_doubles, double **vec_out);
void dist_p2p_teardown(Ctx &c);
// sum all-reduce, in place, of `count` doubles at offset `off` (doubles, even; count even) of the vector region
int dist_p2p_allreduce(Ctx &c, long long off, long long count, const int *skip);

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void wait_flag(const unsigned *p, unsigned epoch) {
  while ((int)(ld_acquire_sys(p) - epoch) < 0) {}
}
// Scalar gather + finaliser in one kernel (one warp): push this rank's raw values into every rank's slot table,
// flag, wait for everybody's flag, combine in rank order, run the formula.  Slot tables alternate with the epoch's
// parity: a fast rank may already be pushing the next gather while a slow one still combines this one.
template <class Fin>
__global__ void k_p2p_gather_fin(Fin fin, DevScalars *S, PeerPtrs P, int ns, int nm) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  PeerSync *me = P.sync[P.rank];
  const unsigned epoch = me->epoch_sc + 1u;
  const int nv = ns + nm, par = (int)(epoch & 1u), t = threadIdx.x;
  if (t < P.world) {
    volatile double *dst = P.sync[t]->grecv[par] + P.rank * nv;
    const double *src = S->gsend + P.rank * nv;
    for (int k = 0; k < nv; ++k) dst[k] = src[k];
    __threadfence_system();
    st_release_sys(&P.sync[t]->flag_sc[P.rank], epoch);
    wait_flag(&me->flag_sc[t], epoch);
  }
  __syncwarp();
  if (t != 0) return;
  double o[kMaxRedVals];
  const volatile double *tab = me->grecv[par];
  for (int k = 0; k < nv; ++k) {
    double v = (k < ns) ? 0.0 : -INFINITY;
    for (int r = 0; r < P.world; ++r) {
      const double pv = tab[r * nv + k];
      v = (k < ns) ? (v + pv) : fmax(v, pv);
    }
    o[k] = v;
  }
  double *slot = S->gsend + P.rank * nv;
  for (int k = 0; k < nv; ++k) slot[k] = 0.0;
  me->epoch_sc = epoch;
  fin(o, S);
}

// Deferred finaliser of a grid_reduce_fin (common.cuh): all-reduce the ns sums and nm maxes the
// kernel parked in S->part, then run the formula on every rank.  No-op on a single GPU.
template <class Fin>
inline int dist_finish(Ctx &c, int ns, int nm, Fin fin) {
  if (!c.dist) return 0;
  if (c.p2p) {
    k_p2p_gather_fin<<<1, 32, 0, c.stream>>>(fin, c.S, c.peers, ns, nm);
    c.launches++;
    c.collectives++;
    c.collective_bytes += (long long)c.world * (ns + nm) * 8;
    return 0;
  }
  if (dist_allreduce_oop(c, c.S->gsend, c.S->grecv, (size_t)c.world * (size_t)(ns + nm))) return -1;
  k_apply_fin<<<1, 32, 0, c.stream>>>(fin, c.S, ns, nm);
  c.launches++;
  return 0;
}
#endif

}  // namespace b200
