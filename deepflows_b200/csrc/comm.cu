// Data-parallel communication: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The reference has no distributed layer at all (README mentions dist/ but the tree has none,
// SURVEY §0.4), so this is new: gradient buckets are sum-all-reduced in place on a dedicated
// communication stream that is ordered after the compute stream by an event, which lets the
// reduction of early buckets overlap the rest of backward.
//
// NCCL is loaded with dlopen so that libdfb200.so itself has no link-time dependency on it
// (single-GPU users never touch it).
#include "common.cuh"
#include "peer.cuh"

#include <dlfcn.h>

namespace dfb {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt32 = 2, ncclFloat32 = 7 };
enum { ncclSum = 0, ncclMin = 3 };

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*CommAbort)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;
cudaEvent_t g_ev_compute = nullptr, g_ev_comm = nullptr;

dfb_status load_nccl() {
  if (g_nccl.lib) return DFB_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  const char* env = getenv("DFB_NCCL_LIB");
  void* lib = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
  for (int i = 0; !lib && i < 2; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) DFB_FAIL(DFB_ERR_RUNTIME, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define SYM(field, name)                                                   \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                              \
  if (!g_nccl.field) DFB_FAIL(DFB_ERR_RUNTIME, "NCCL symbol %s missing", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(CommAbort, "ncclCommAbort");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Broadcast, "ncclBroadcast");
  SYM(AllGather, "ncclAllGather");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetVersion, "ncclGetVersion");
#undef SYM
  g_nccl.lib = lib;
  return DFB_OK;
}

#define DFB_NCCL(expr)                                                                          \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != ncclSuccess)                                                                      \
      DFB_FAIL(DFB_ERR_RUNTIME, "%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
  } while (0)

}  // namespace

// Small blocking exchanges of HOST values for the peer-memory set-up (peer.cu): IPC handles and verdicts travel through
// the communicator that already exists instead of a second rendezvous.
dfb_status comm_allgather_bytes(const void* mine, size_t bytes, void* all) {
  DFB_REQUIRE(g_comm != nullptr, DFB_ERR_RUNTIME, "comm_allgather: communicator not initialised");
  unsigned char* dev = nullptr;
  DFB_CUDA(cudaMalloc((void**)&dev, bytes * (size_t)(g_world + 1)));
  cudaStream_t s = comm_stream();
  DFB_CUDA(cudaMemcpyAsync(dev, mine, bytes, cudaMemcpyHostToDevice, s));
  DFB_NCCL(g_nccl.AllGather(dev, dev + bytes, bytes, ncclInt8, g_comm, s));
  DFB_CUDA(cudaMemcpyAsync(all, dev + bytes, bytes * (size_t)g_world, cudaMemcpyDeviceToHost, s));
  DFB_CUDA(cudaStreamSynchronize(s));
  DFB_CUDA(cudaFree(dev));
  return DFB_OK;
}
dfb_status comm_allreduce_min_int(int* value) {
  DFB_REQUIRE(g_comm != nullptr, DFB_ERR_RUNTIME, "comm_allreduce: communicator not initialised");
  int* dev = nullptr;
  DFB_CUDA(cudaMalloc((void**)&dev, sizeof(int)));
  cudaStream_t s = comm_stream();
  DFB_CUDA(cudaMemcpyAsync(dev, value, sizeof(int), cudaMemcpyHostToDevice, s));
  DFB_NCCL(g_nccl.AllReduce(dev, dev, 1, ncclInt32, ncclMin, g_comm, s));
  DFB_CUDA(cudaMemcpyAsync(value, dev, sizeof(int), cudaMemcpyDeviceToHost, s));
  DFB_CUDA(cudaStreamSynchronize(s));
  DFB_CUDA(cudaFree(dev));
  return DFB_OK;
}
}  // namespace dfb

using namespace dfb;

extern "C" {

dfb_status dfb_comm_unique_id(unsigned char* id128) {
  dfb_status st = load_nccl();
  if (st != DFB_OK) return st;
  ncclUniqueId id;
  DFB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return DFB_OK;
}

dfb_status dfb_comm_init(const unsigned char* id128, int rank, int world_size) {
  DFB_INIT();
  DFB_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, DFB_ERR_INVALID,
              "comm_init: bad rank %d / world %d", rank, world_size);
  DFB_REQUIRE(g_comm == nullptr, DFB_ERR_RUNTIME, "comm_init: communicator already initialised");
  dfb_status st = load_nccl();
  if (st != DFB_OK) return st;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  DFB_NCCL(g_nccl.CommInitRank(&g_comm, world_size, id, rank));
  g_rank = rank;
  g_world = world_size;
  DFB_CUDA(cudaEventCreateWithFlags(&g_ev_compute, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&g_ev_comm, cudaEventDisableTiming));
  return DFB_OK;
}

dfb_status dfb_comm_destroy(void) {
  if (!g_comm) return DFB_OK;
  dfb_peer_destroy();   // the peers' arenas are unmapped while the communicator can still synchronise the ranks
  // All collectives this rank enqueued have completed once the streams are idle. The communicator object
  // itself is then dropped, not destroyed: ncclCommDestroy / ncclCommAbort block for as long as a captured
  // CUDA graph still references the communicator (NCCL 2.27.3: both were observed to hang here with the
  // training-step graph alive), and a process that is about to exit must not wait on that. The driver
  // reclaims NCCL's resources with the process (set DFB_NCCL_DESTROY=1 to call ncclCommDestroy anyway).
  cudaStreamSynchronize(compute_stream());
  cudaStreamSynchronize(comm_stream());
  const char* env = getenv("DFB_NCCL_DESTROY");
  if (env && env[0] == '1') DFB_NCCL(g_nccl.CommDestroy(g_comm));
  g_comm = nullptr;
  g_rank = 0;
  g_world = 1;
  if (g_ev_compute) cudaEventDestroy(g_ev_compute);
  if (g_ev_comm) cudaEventDestroy(g_ev_comm);
  g_ev_compute = g_ev_comm = nullptr;
  return DFB_OK;
}

dfb_status dfb_comm_rank(int* rank, int* world_size) {
  if (rank) *rank = g_rank;
  if (world_size) *world_size = g_world;
  return DFB_OK;
}

dfb_status dfb_comm_allreduce_async(float* buf, size_t n) {
  DFB_INIT();
  DFB_REQUIRE(g_comm != nullptr, DFB_ERR_RUNTIME, "comm_allreduce: communicator not initialised");
  if (n == 0) return DFB_OK;
  // comm stream waits for the producers of this bucket (everything enqueued so far)
  DFB_CUDA(cudaEventRecord(g_ev_compute, compute_stream()));
  DFB_CUDA(cudaStreamWaitEvent(comm_stream(), g_ev_compute, 0));
  DFB_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat32, ncclSum, g_comm, comm_stream()));
  return DFB_OK;
}

dfb_status dfb_comm_broadcast_async(float* buf, size_t n, int root) {
  DFB_INIT();
  DFB_REQUIRE(g_comm != nullptr, DFB_ERR_RUNTIME, "comm_broadcast: communicator not initialised");
  if (n == 0) return DFB_OK;
  DFB_CUDA(cudaEventRecord(g_ev_compute, compute_stream()));
  DFB_CUDA(cudaStreamWaitEvent(comm_stream(), g_ev_compute, 0));
  DFB_NCCL(g_nccl.Broadcast(buf, buf, n, ncclFloat32, root, g_comm, comm_stream()));
  return DFB_OK;
}

dfb_status dfb_comm_wait(void) {
  DFB_INIT();
  if (!g_comm) return DFB_OK;
  DFB_CUDA(cudaEventRecord(g_ev_comm, comm_stream()));
  DFB_CUDA(cudaStreamWaitEvent(compute_stream(), g_ev_comm, 0));
  return DFB_OK;
}

}  // extern "C"
