// dist.cu -- see dist.cuh
#include <dlfcn.h>

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
