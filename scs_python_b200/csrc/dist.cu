// dist.cu -- see dist.cuh
#include <dlfcn.h>
#include <cstdint>

#include "dist.cuh"

namespace b200 {

namespace {
// the few NCCL entry points used, with the library's own ABI (nccl.h 2.2x)
typedef struct { char internal[128]; } NcclUniqueId;
typedef int (*fn_get_unique_id)(NcclUniqueId *);
typedef int (*fn_comm_init_rank)(void **, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_error_string)(int);
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;

struct NcclApi {
  void *lib = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_error_string error_string = nullptr;
  bool load() {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      fprintf(stderr, "libscsb200: cannot load libnccl.so.2 (%s)\n", dlerror());
      return false;
    }
    get_unique_id = (fn_get_unique_id)dlsym(lib, "ncclGetUniqueId");
    comm_init_rank = (fn_comm_init_rank)dlsym(lib, "ncclCommInitRank");
    comm_destroy = (fn_comm_destroy)dlsym(lib, "ncclCommDestroy");
    all_reduce = (fn_all_reduce)dlsym(lib, "ncclAllReduce");
    error_string = (fn_error_string)dlsym(lib, "ncclGetErrorString");
    if (!get_unique_id || !comm_init_rank || !comm_destroy || !all_reduce) {
      fprintf(stderr, "libscsb200: libnccl is missing required symbols\n");
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
Dist g_dist;
}  // namespace

Dist *dist_current() { return &g_dist; }

int dist_allreduce(Ctx &c, double *buf, size_t count, int op) {
  if (!g_dist.on() || count == 0) return 0;
  const int rc = g_nccl.all_reduce(buf, buf, count, kNcclFloat64, op == 1 ? kNcclMax : kNcclSum, g_dist.comm, c.stream);
  if (rc != 0) {
    fprintf(stderr, "libscsb200: ncclAllReduce failed: %s\n", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return -1;
  }
  c.collectives++;
  c.collective_bytes += (long long)(count * sizeof(double));
  return 0;
}

int dist_allreduce_oop(Ctx &c, const double *send, double *recv, size_t count) {
  if (count == 0) return 0;
  if (!g_dist.on()) {  // world == 1 (self-test of the partitioned code path on one GPU)
    if (cudaMemcpyAsync(recv, send, count * sizeof(double), cudaMemcpyDeviceToDevice, c.stream) != cudaSuccess) return -1;
    return 0;
  }
  const int rc = g_nccl.all_reduce(send, recv, count, kNcclFloat64, kNcclSum, g_dist.comm, c.stream);
  if (rc != 0) {
    fprintf(stderr, "libscsb200: ncclAllReduce failed: %s\n", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return -1;
  }
  c.collectives++;
  c.collective_bytes += (long long)(count * sizeof(double));
  return 0;
}

// ------------------------------------------------------------------ peer-memory path ---
namespace {
typedef int (*fn_cu_range)(unsigned long long *, size_t *, unsigned long long);
fn_cu_range g_cu_range = nullptr;
bool load_cu_range() {
  if (g_cu_range) return true;
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_cu_range = (fn_cu_range)dlsym(h, "cuMemGetAddressRange_v2");
  return g_cu_range != nullptr;
}
struct PeerSlot {  // what a rank publishes: the IPC handle of its arena and where the arena starts inside it
  cudaIpcMemHandle_t handle;
  unsigned long long offset;
  unsigned long long bytes;
};
constexpr size_t kSyncBytes = (sizeof(PeerSync) + 255) / 256 * 256;
constexpr int kNcclUint8 = 1;
}  // namespace

int dist_p2p_setup(Ctx &c, size_t vec_doubles, double **vec_out) {
  *vec_out = nullptr;
  c.p2p = false;
  const char *e = getenv("SCS_B200_DIST_P2P");
  if (!g_dist.on() || (e && atoi(e) == 0) || !load_cu_range() || g_dist.world > kMaxWorld) return 0;
  const int W = g_dist.world, R = g_dist.rank;
  const size_t bytes = kSyncBytes + vec_doubles * sizeof(double);
  void *arena = nullptr;
  PeerSlot *slots_d = nullptr;
  std::vector<PeerSlot> slots((size_t)W);
  int ok_local = 1;
  if (cudaMalloc(&arena, bytes) != cudaSuccess || cudaMemsetAsync(arena, 0, bytes, c.stream) != cudaSuccess) ok_local = 0;
  memset(slots.data(), 0, sizeof(PeerSlot) * (size_t)W);
  if (ok_local) {
    unsigned long long base = 0;
    size_t sz = 0;
    if (g_cu_range(&base, &sz, (unsigned long long)(uintptr_t)arena) != 0 ||
        cudaIpcGetMemHandle(&slots[(size_t)R].handle, arena) != cudaSuccess)
      ok_local = 0;
    else {
      slots[(size_t)R].offset = (unsigned long long)(uintptr_t)arena - base;
      slots[(size_t)R].bytes = bytes;
    }
  }
  if (!ok_local) memset(&slots[(size_t)R], 0, sizeof(PeerSlot));  // bytes == 0 tells the others
  // gather the slots: every rank fills its own, the byte-wise sum over the ranks is the table
  bool comm_ok = dev_alloc(&slots_d, (size_t)W) == 0 &&
                 cudaMemcpyAsync(slots_d, slots.data(), sizeof(PeerSlot) * (size_t)W, cudaMemcpyHostToDevice, c.stream) == cudaSuccess &&
                 g_nccl.all_reduce(slots_d, slots_d, sizeof(PeerSlot) * (size_t)W, kNcclUint8, kNcclSum, g_dist.comm, c.stream) == 0 &&
                 cudaMemcpyAsync(slots.data(), slots_d, sizeof(PeerSlot) * (size_t)W, cudaMemcpyDeviceToHost, c.stream) == cudaSuccess &&
                 cudaStreamSynchronize(c.stream) == cudaSuccess;
  dev_free(slots_d);
  bool all_ok = comm_ok;
  for (int r = 0; r < W && all_ok; ++r) all_ok = slots[(size_t)r].bytes != 0;
  PeerPtrs P{};
  P.world = W; P.rank = R;
  int opened = 0;
  for (int r = 0; r < W && all_ok; ++r) {
    char *base = nullptr;
    if (r == R) base = (char *)arena;
    else {
      void *m = nullptr;
      if (cudaIpcOpenMemHandle(&m, slots[(size_t)r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; break; }
      c.p2p_opened[opened++] = m;
      base = (char *)m + slots[(size_t)r].offset;
    }
    P.sync[r] = (PeerSync *)base;
    P.vec[r] = (double *)(base + kSyncBytes);
  }
  // every rank must reach the same verdict (a rank that failed to map would fall back to NCCL while the others spin)
  {
    int *flag_d = nullptr;
    int v = all_ok ? 0 : 1;
    bool fine = dev_alloc(&flag_d, 1) == 0 && cudaMemcpyAsync(flag_d, &v, sizeof(int), cudaMemcpyHostToDevice, c.stream) == cudaSuccess &&
                g_nccl.all_reduce(flag_d, flag_d, 1, 2 /* ncclInt32 */, kNcclSum, g_dist.comm, c.stream) == 0 &&
                cudaMemcpyAsync(&v, flag_d, sizeof(int), cudaMemcpyDeviceToHost, c.stream) == cudaSuccess &&
                cudaStreamSynchronize(c.stream) == cudaSuccess;
    dev_free(flag_d);
    if (!fine || v != 0) all_ok = false;
  }
  if (!all_ok) {
    cudaGetLastError();
    for (int k = 0; k < opened; ++k) { cudaIpcCloseMemHandle(c.p2p_opened[k]); c.p2p_opened[k] = nullptr; }
    if (arena) cudaFree(arena);
    if (R == 0) fprintf(stderr, "libscsb200: peer-memory collectives unavailable (CUDA IPC), using NCCL for every exchange\n");
    return 0;
  }
  c.p2p = true;
  c.peers = P;
  c.p2p_arena = arena;
  *vec_out = P.vec[R];
  return 0;
}

void dist_p2p_teardown(Ctx &c) {
  if (!c.p2p_arena) return;
  cudaStreamSynchronize(c.stream);
  for (int k = 0; k < 2 * kMaxWorld; ++k)
    if (c.p2p_opened[k]) { cudaIpcCloseMemHandle(c.p2p_opened[k]); c.p2p_opened[k] = nullptr; }
  cudaFree(c.p2p_arena);
  c.p2p_arena = nullptr;
  c.p2p = false;
}

// Two-shot all-reduce over peer memory: rank g sums slice g of every rank's vector in rank order (one owner per
// element: the result is bit-identical everywhere) and writes the sums back into every rank's vector.  Flags:
// flag_in = "my vector is ready", flag_out = "I am done with everybody's slice"; epochs only grow.
__global__ void __launch_bounds__(512)
k_p2p_allreduce(PeerPtrs P, long long off, long long count, const int *skip) {
  if (skip != nullptr && *skip != 0) return;
  PeerSync *me = P.sync[P.rank];
  const unsigned epoch = me->epoch_ar + 1u;
  const int W = P.world;
  if (blockIdx.x == 0 && (int)threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(&P.sync[threadIdx.x]->flag_in[P.rank], epoch);
  }
  if ((int)threadIdx.x < W) wait_flag(&me->flag_in[threadIdx.x], epoch);
  __syncthreads();
  const long long npairs = count / 2, per = (npairs + W - 1) / W;
  const long long lo = (long long)P.rank * per, hi = lo + per < npairs ? lo + per : npairs;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
    double2 acc = make_double2(0.0, 0.0);
    for (int r = 0; r < W; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2 *>(P.vec[r] + off) + i);
      acc.x += v.x; acc.y += v.y;
    }
    for (int r = 0; r < W; ++r) __stcg(reinterpret_cast<double2 *>(P.vec[r] + off) + i, acc);
  }
  if ((count & 1) && P.rank == 0 && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail element
    double acc = 0.0;
    for (int r = 0; r < W; ++r) acc += __ldcg(P.vec[r] + off + count - 1);
    for (int r = 0; r < W; ++r) __stcg(P.vec[r] + off + count - 1, acc);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(&me->done_cnt, 1u) == gridDim.x - 1) {  // last CTA of this rank
      me->done_cnt = 0u;
      for (int r = 0; r < W; ++r) st_release_sys(&P.sync[r]->flag_out[P.rank], epoch);
      for (int r = 0; r < W; ++r) wait_flag(&me->flag_out[r], epoch);
      me->epoch_ar = epoch;
      __threadfence();
    }
  }
}

int dist_p2p_allreduce(Ctx &c, long long off, long long count, const int *skip) {
  if (!c.p2p || count <= 0) return 0;
  // at most one CTA per SM: every CTA of the grid must be resident while it waits for the other ranks
  const long long blocks = (count / 2 / c.world + 511) / 512;
  const int grid = (int)(blocks < 1 ? 1 : (blocks > 64 ? 64 : blocks));
  k_p2p_allreduce<<<grid, 512, 0, c.stream>>>(c.peers, off, count, skip);
  c.launches++;
  c.collectives++;
  c.collective_bytes += count * 8;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int current_device();

}  // namespace b200

using namespace b200;

extern "C" scs_int scs_b200_dist_unique_id(void *out128) {
  if (!out128 || !g_nccl.load()) return -1;
  NcclUniqueId id;
  if (g_nccl.get_unique_id(&id) != 0) return -1;
  memcpy(out128, &id, sizeof(id));
  return 0;
}

extern "C" scs_int scs_b200_dist_init(scs_int rank, scs_int world, const void *id128) {
  if (world < 1 || rank < 0 || rank >= world) return -1;
  if (g_dist.comm) return -1;  // already initialised
  if (world == 1) {
    g_dist.rank = 0; g_dist.world = 1;
    const char *e = getenv("SCS_B200_DIST_SELFTEST");
    g_dist.selftest = e && *e && atoi(e) != 0;
    return 0;
  }
  if (!id128 || !g_nccl.load()) return -1;
  if (cudaSetDevice(current_device()) != cudaSuccess) return -1;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  void *comm = nullptr;
  const int rc = g_nccl.comm_init_rank(&comm, world, id, rank);
  if (rc != 0) {
    fprintf(stderr, "libscsb200: ncclCommInitRank failed: %s\n", g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return -1;
  }
  g_dist.comm = comm;
  g_dist.rank = rank;
  g_dist.world = world;
  return 0;
}

extern "C" void scs_b200_dist_finalize(void) {
  if (g_dist.comm) g_nccl.comm_destroy(g_dist.comm);
  g_dist = Dist();
}

extern "C" scs_int scs_b200_dist_rank(void) { return g_dist.rank; }
extern "C" scs_int scs_b200_dist_world(void) { return g_dist.world; }
