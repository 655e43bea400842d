// Gradient exchange over NVLink 5 / NVSwitch peer memory: the data-parallel all-reduce as ONE kernel per bucket, and
// its completion awaited INSIDE the fused optimizer kernel (optim.cu) instead of by a stream dependency.
//
// The reference has no distributed layer; round 1 reduced the buckets with ncclAllReduce on a
// communication stream (comm.cu, still the checked fallback). For buckets of a few MB that path is all fixed cost: an
// event hop to the communication stream, NCCL's own launch + protocol latency, a second event hop back in front of the
// optimizer. Here every rank owns one symmetric arena (flags + gradient buckets, one cudaMalloc) that every other rank
// maps through CUDA IPC, and a bucket is reduced by `peer_allreduce_kernel` running on the stream that packed it:
//
//   phase 0  (peer_handshake_kernel, one warp) tells every peer "rank r's copy of this bucket is packed" and waits until
//            every peer has said so
//   reduce   rank r owns slice r of the bucket: its CTAs read that slice from ALL ranks over NVLink (128-bit loads, all
//            peers' loads in flight together), add them in rank order, and store the sum into slice r of EVERY rank's
//            arena (reduce-scatter + all-gather, "two-shot", in one pass; every element is computed by exactly one
//            rank, so replicas receive identical bits)
//   phase 1  each CTA tells every peer "my part of slice r has landed" and EXITS - nobody waits here
//
// The arrival of all slices (phase-1 count == world x CTAs x epoch) is awaited by whoever consumes the gradients:
// multi_adam_kernel / multi_sgd_kernel take the pending buckets as a bit mask and spin on the counters before their first
// gradient load (peer_wait_slots), so the last bucket's latency overlaps the optimizer kernel's launch and table loads;
// dfb_peer_wait() is the stand-alone form for everything else (non-fused optimizers, reading .grad).
//
// Epochs instead of flag resets: every launch of a slot bumps the slot's counter in local memory, signals carry the
// epoch (phase 0) or are atomic adds (phase 1), and waits compare against the epoch (x contributors) - a captured CUDA graph replays the same kernels with frozen
// arguments and stays correct. A spin that sees no progress for kSpinTimeoutNs sets a sticky error word (pinned host
// memory, dfb_peer_status) and falls through: a lost peer must not hang the GPU.
// dfb_peer_init() finishes with a self-test (a known pattern all-reduced through the kernel) and an NCCL all-reduce of its
// verdict, so either every rank uses this path or every rank stays on NCCL.
#include "common.cuh"
#include "peer.cuh"

#include <algorithm>
#include <vector>

namespace dfb {

namespace {
PeerDev g_peer;                 // host copy
PeerDev* g_peer_dev = nullptr;  // device copy (what the kernels dereference)
void* g_block = nullptr;        // this rank's [flags | arena]
void* g_mapped[kPeerMaxWorld] = {nullptr};
unsigned* g_epoch_dev = nullptr;
volatile unsigned* g_error_host = nullptr;
size_t g_arena_floats = 0;
unsigned long long g_pending = 0;   // slots reduced since the last wait (host view, in stream order of the compute stream)
bool g_ready = false;
cudaEvent_t g_ev_fill = nullptr, g_ev_done = nullptr;

constexpr size_t kFlagBytes = (size_t)kPeerSlots * kPeerFlagWords * sizeof(unsigned);
}  // namespace

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Phase 0 as a kernel of its own (one warp): a bucket's reduction waits for the slowest rank, and a grid that spins
// while it waits keeps its SMs' registers from the convolution kernels running beside it (measured on 2 GPUs: 32 CTAs x
// 512 threads x 100 registers spinning in front of the reduction cost the step 45 us). This warp costs nothing; the
// reduction grid behind it starts when every rank's bucket is packed and never waits.
__global__ void __launch_bounds__(32) peer_handshake_kernel(const PeerDev* __restrict__ pd, int slot, int ctas) {
  pdl_sync();
  const int world = pd->world, rank = pd->rank, q = threadIdx.x;
  unsigned E = 0;
  if (q == 0) {
    E = pd->epoch[slot] + 1;
    pd->epoch[slot] = E;
    pd->epoch[kPeerSlots + slot] += (unsigned)(world * ctas);   // what a consumer of this launch waits for (peer_wait_slots)
  }
  E = __shfl_sync(0xffffffffu, E, 0);
  if (q < world) {
    // my bucket is packed (stream order: the pack kernel precedes this one) ...
    __threadfence_system();
    *(volatile unsigned*)(pd->flags[q] + peer_flag_index(slot, 0, rank)) = E;
    // ... and so is rank q's
    peer_spin(pd->flags[rank] + peer_flag_index(slot, 0, q), E, pd->error);
  }
}

// The reduction proper: CTA c of rank r, see the file header. block = kPeerThreads.
// WORLD is a template parameter so that the loads of U = 8 / WORLD positions x WORLD ranks (8 x 128 bit per thread, 2 MB
// per launch) are all issued before the first add: an NVLink round trip is ~3 us, and with one position per iteration
// the first version moved 85 GB/s.
template <int WORLD>
__global__ void __launch_bounds__(kPeerThreads)
peer_allreduce_kernel(const PeerDev* __restrict__ pdp, unsigned long long off, unsigned long long n, int slot) {
  pdl_sync();
  __shared__ PeerDev pd;
  if (threadIdx.x == 0) pd = *pdp;
  __syncthreads();
  const int rank = pd.rank, world = pd.world, c = blockIdx.x;   // world <= WORLD
  // ---- reduce slice `rank`, write it to everyone ----
  float* arena[WORLD];
#pragma unroll
  for (int q = 0; q < WORLD; ++q) arena[q] = pd.arena[q < world ? q : 0] + off;
  const unsigned long long n4 = n >> 2;
  const unsigned long long per_rank = (n4 + world - 1) / world;
  const unsigned long long s_begin = min(n4, per_rank * rank), s_end = min(n4, s_begin + per_rank);
  const unsigned long long per_cta = (s_end - s_begin + gridDim.x - 1) / gridDim.x;
  const unsigned long long c_begin = min(s_end, s_begin + per_cta * c), c_end = min(s_end, c_begin + per_cta);
  constexpr int U = 8 / WORLD;
  for (unsigned long long i0 = c_begin + threadIdx.x; i0 < c_end; i0 += (unsigned long long)U * kPeerThreads) {
    float4 v[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long i = i0 + (unsigned long long)u * kPeerThreads;
      if (i < c_end) {
#pragma unroll
        for (int q = 0; q < WORLD; ++q)
          if (q < world) v[u][q] = ld_peer_f4(arena[q] + (i << 2));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long i = i0 + (unsigned long long)u * kPeerThreads;
      if (i < c_end) {
        float4 s = v[u][0];   // rank order: every element is summed by one rank only, in this order
#pragma unroll
        for (int q = 1; q < WORLD; ++q)
          if (q < world) { s.x += v[u][q].x; s.y += v[u][q].y; s.z += v[u][q].z; s.w += v[u][q].w; }
#pragma unroll
        for (int q = 0; q < WORLD; ++q)
          if (q < world) *reinterpret_cast<float4*>(arena[q] + (i << 2)) = s;
      }
    }
  }
  // ---- phase 1: my part of slice `rank` has landed everywhere ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    atomicAdd_system(pd.flags[threadIdx.x] + peer_flag_index(slot, 1, 0), 1u);
  }
}

__global__ void __launch_bounds__(256) peer_wait_kernel(const PeerDev* __restrict__ pd, unsigned long long slots) {
  pdl_sync();
  peer_wait_slots(pd, slots);
}

// self-test helpers: arena[i] = (rank + 1) * (i % 251 + 1); the sum over ranks is exact in fp32
__global__ void peer_test_fill_kernel(float* a, size_t n, int rank) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a[i] = (float)((rank + 1) * (int)(i % 251 + 1));
}
__global__ void peer_test_check_kernel(const float* a, size_t n, int world, unsigned* bad) {
  const float tri = (float)(world * (world + 1) / 2);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (a[i] != tri * (float)(int)(i % 251 + 1)) atomicAdd(bad, 1u);
}

// The reductions run on the communication stream (a bucket's reduction waits for the slowest rank: on the stream that
// packed it - the side stream - it would hold up the weight gradients queued behind it). Whoever consumes the buckets is
// ordered behind THIS rank's reduction kernels by an event (their epoch counters are what the consumer's wait reads);
// the other ranks' slices are awaited inside the consumer kernel.
unsigned long long peer_pending_take() {
  const unsigned long long m = g_pending;
  if (m) {
    cudaEventRecord(g_ev_done, comm_stream());
    cudaStreamWaitEvent(compute_stream(), g_ev_done, 0);
  }
  g_pending = 0;
  return m;
}
const PeerDev* peer_dev() { return g_peer_dev; }

static dfb_status peer_release() {
  for (int q = 0; q < kPeerMaxWorld; ++q) {
    if (g_mapped[q]) cudaIpcCloseMemHandle(g_mapped[q]);
    g_mapped[q] = nullptr;
  }
  if (g_block) cudaFree(g_block);
  if (g_peer_dev) cudaFree(g_peer_dev);
  if (g_epoch_dev) cudaFree(g_epoch_dev);
  if (g_error_host) cudaFreeHost((void*)g_error_host);
  if (g_ev_fill) cudaEventDestroy(g_ev_fill);
  if (g_ev_done) cudaEventDestroy(g_ev_done);
  g_ev_fill = g_ev_done = nullptr;
  g_block = nullptr; g_peer_dev = nullptr; g_epoch_dev = nullptr; g_error_host = nullptr;
  g_ready = false; g_pending = 0; g_arena_floats = 0;
  return DFB_OK;
}

static dfb_status peer_launch(size_t offset, size_t n, int slot) {
  // a slot is reused only after its previous launch has been awaited (on every rank: all ranks issue the same sequence)
  if (g_pending & (1ull << slot)) {
    launch_k(peer_wait_kernel, dim3(1), dim3(256), 0, comm_stream(), (const PeerDev*)g_peer_dev, 1ull << slot);
    DFB_LAUNCH_CHECK("peer_wait");
    g_pending &= ~(1ull << slot);
  }
  // behind whatever filled the range (everything enqueued on the current compute / side stream so far)
  DFB_CUDA(cudaEventRecord(g_ev_fill, compute_stream()));
  DFB_CUDA(cudaStreamWaitEvent(comm_stream(), g_ev_fill, 0));
  // Few CTAs: every reduction but the last overlaps backward, where its latency is free and its footprint is not (64 CTAs
  // per bucket cost the convolutions beside them ~20 us of the step on 2 GPUs); the last bucket is small (dist.py).
  const size_t slice_bytes = (n / (size_t)g_peer.world) * sizeof(float);
  const int ctas = (int)std::max<size_t>(1, std::min<size_t>(kPeerMaxCtas, (slice_bytes + kPeerBytesPerCta - 1) / kPeerBytesPerCta));
  launch_k(peer_handshake_kernel, dim3(1), dim3(32), 0, comm_stream(), (const PeerDev*)g_peer_dev, slot, ctas);
  DFB_LAUNCH_CHECK("peer_handshake");
  auto go = [&](auto kernel) {
    launch_k(kernel, dim3(ctas), dim3(kPeerThreads), 0, comm_stream(), (const PeerDev*)g_peer_dev, (unsigned long long)offset,
             (unsigned long long)n, slot);
  };
  switch (g_peer.world) {   // (ranks beyond the world size of a variant would only cost registers)
    case 2: go(peer_allreduce_kernel<2>); break;
    case 3: case 4: go(peer_allreduce_kernel<4>); break;
    default: go(peer_allreduce_kernel<8>); break;
  }
  DFB_LAUNCH_CHECK("peer_allreduce");
  g_pending |= 1ull << slot;
  return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

dfb_status dfb_peer_init(size_t arena_floats, float** arena) {
  DFB_INIT();
  DFB_REQUIRE(arena != nullptr, DFB_ERR_INVALID, "peer_init: null result pointer");
  DFB_REQUIRE(!g_ready, DFB_ERR_RUNTIME, "peer_init: already initialised");
  int rank = 0, world = 1;
  dfb_comm_rank(&rank, &world);
  DFB_REQUIRE(world >= 2, DFB_ERR_RUNTIME, "peer_init: needs an initialised communicator of at least two ranks (dfb_comm_init)");
  DFB_REQUIRE(world <= kPeerMaxWorld, DFB_ERR_RUNTIME, "peer_init: at most %d ranks (one NVLink domain)", kPeerMaxWorld);
  arena_floats = (arena_floats + 3) & ~size_t(3);
  const size_t test_floats = 1u << 16;   // the self-test's bucket
  if (arena_floats < test_floats) arena_floats = test_floats;
  const size_t bytes = kFlagBytes + arena_floats * sizeof(float);

  // Any failure below must leave every rank on the same side: the verdict is all-reduced (min) through NCCL before
  // anybody returns, so the steps up to there record their status instead of returning.
  int ok = 1;
  std::string why;
  auto fail = [&](const char* what, cudaError_t e) { if (ok) { ok = 0; why = std::string(what) + ": " + cudaGetErrorString(e); } cudaGetLastError(); };
  cudaError_t e;
  if ((e = cudaMalloc(&g_block, bytes)) != cudaSuccess) fail("cudaMalloc of the arena", e);
  if (ok && (e = cudaMemset(g_block, 0, bytes)) != cudaSuccess) fail("cudaMemset", e);
  if (ok && (e = cudaMalloc((void**)&g_epoch_dev, 2 * kPeerSlots * sizeof(unsigned))) != cudaSuccess) fail("cudaMalloc", e);
  if (ok && (e = cudaMemset(g_epoch_dev, 0, 2 * kPeerSlots * sizeof(unsigned))) != cudaSuccess) fail("cudaMemset", e);
  if (ok && (e = cudaHostAlloc((void**)&g_error_host, 64, cudaHostAllocMapped)) != cudaSuccess) fail("cudaHostAlloc", e);
  if (ok) *g_error_host = 0;
  if (ok && (e = cudaMalloc((void**)&g_peer_dev, sizeof(PeerDev))) != cudaSuccess) fail("cudaMalloc", e);
  if (ok && (e = cudaEventCreateWithFlags(&g_ev_fill, cudaEventDisableTiming)) != cudaSuccess) fail("cudaEventCreate", e);
  if (ok && (e = cudaEventCreateWithFlags(&g_ev_done, cudaEventDisableTiming)) != cudaSuccess) fail("cudaEventCreate", e);
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && (e = cudaIpcGetMemHandle(&mine, g_block)) != cudaSuccess) fail("cudaIpcGetMemHandle", e);

  // exchange [ok byte | handle] with every rank (NCCL all-gather on a small device buffer)
  constexpr size_t kRec = 128;
  static_assert(sizeof(cudaIpcMemHandle_t) + 1 <= kRec, "record size");
  std::vector<unsigned char> all(kRec * world, 0);
  unsigned char rec[kRec] = {0};
  rec[0] = (unsigned char)ok;
  memcpy(rec + 1, &mine, sizeof(mine));
  dfb_status st = comm_allgather_bytes(rec, kRec, all.data());
  if (st != DFB_OK) { peer_release(); return st; }
  for (int q = 0; q < world; ++q)
    if (!all[q * kRec]) { if (ok) { ok = 0; why = "rank " + std::to_string(q) + " could not allocate / export its arena"; } }
  memset(&g_peer, 0, sizeof(g_peer));
  g_peer.world = world;
  g_peer.rank = rank;
  if (ok) {
    for (int q = 0; q < world && ok; ++q) {
      void* base = g_block;
      if (q != rank) {
        cudaIpcMemHandle_t h;
        memcpy(&h, all.data() + q * kRec + 1, sizeof(h));
        if ((e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess) { fail("cudaIpcOpenMemHandle", e); break; }
        g_mapped[q] = base;
      }
      g_peer.flags[q] = (unsigned*)base;
      g_peer.arena[q] = (float*)((char*)base + kFlagBytes);
    }
  }
  if (ok) {
    g_peer.epoch = g_epoch_dev;
    void* err_dev = nullptr;
    if ((e = cudaHostGetDevicePointer(&err_dev, (void*)g_error_host, 0)) != cudaSuccess) fail("cudaHostGetDevicePointer", e);
    g_peer.error = (unsigned*)err_dev;
    if (ok && (e = cudaMemcpy(g_peer_dev, &g_peer, sizeof(g_peer), cudaMemcpyHostToDevice)) != cudaSuccess) fail("cudaMemcpy", e);
  }
  // ---- self-test: two launches of the last slot (the second proves the epoch arithmetic), pattern checked on the device ----
  unsigned bad_host = 0;
  if (ok) {
    g_ready = true;
    g_arena_floats = arena_floats;
    unsigned* bad = nullptr;
    if ((e = cudaMalloc((void**)&bad, sizeof(unsigned))) != cudaSuccess) fail("cudaMalloc", e);
    if (ok) cudaMemset(bad, 0, sizeof(unsigned));
    cudaStream_t s = compute_stream();
    for (int round = 0; round < 2 && ok; ++round) {
      peer_test_fill_kernel<<<64, 256, 0, s>>>(g_peer.arena[rank], test_floats, rank);
      peer_launch(0, test_floats, kPeerSlots - 1);
      launch_k(peer_wait_kernel, dim3(1), dim3(256), 0, s, (const PeerDev*)g_peer_dev, peer_pending_take());
      peer_test_check_kernel<<<64, 256, 0, s>>>(g_peer.arena[rank], test_floats, world, bad);
      if ((e = cudaStreamSynchronize(s)) != cudaSuccess) fail("self-test", e);
      // nobody may refill its bucket while a peer still reads it: the next round's fill is ordered behind every rank's
      // check by the all-gather below (round 0) / the verdict all-reduce (round 1)
      if (round == 0) {
        unsigned char z[kRec] = {0};
        st = comm_allgather_bytes(z, kRec, all.data());
        if (st != DFB_OK) { ok = 0; why = "all-gather between the self-test rounds failed"; }
      }
    }
    if (ok) cudaMemcpy(&bad_host, bad, sizeof(unsigned), cudaMemcpyDeviceToHost);
    if (bad) cudaFree(bad);
    if (ok && (bad_host != 0 || *g_error_host != 0)) {
      ok = 0;
      why = "self-test: " + std::to_string(bad_host) + " wrong elements, error word " + std::to_string(*g_error_host);
    }
    if (ok) cudaMemset(g_peer.arena[rank], 0, test_floats * sizeof(float));
  }
  int all_ok = ok;
  st = comm_allreduce_min_int(&all_ok);
  if (st != DFB_OK) { peer_release(); return st; }
  if (!all_ok) {
    peer_release();
    DFB_FAIL(DFB_ERR_RUNTIME, "peer_init: peer-memory exchange unavailable (%s)", ok ? "another rank failed" : why.c_str());
  }
  *arena = g_peer.arena[rank];
  return DFB_OK;
}

dfb_status dfb_peer_allreduce_async(size_t offset, size_t n, int slot) {
  DFB_INIT();
  DFB_REQUIRE(g_ready, DFB_ERR_RUNTIME, "peer_allreduce: dfb_peer_init has not succeeded");
  DFB_REQUIRE(slot >= 0 && slot < kPeerSlots - 1, DFB_ERR_INVALID, "peer_allreduce: slot %d outside [0, %d)", slot, kPeerSlots - 1);
  DFB_REQUIRE((offset & 3) == 0 && (n & 3) == 0, DFB_ERR_INVALID, "peer_allreduce: offset and count must be multiples of 4 floats");
  DFB_REQUIRE(offset + n <= g_arena_floats, DFB_ERR_OUT_OF_RANGE, "peer_allreduce: [%zu, %zu) outside the arena of %zu floats", offset,
              offset + n, g_arena_floats);
  if (n == 0) return DFB_OK;
  return peer_launch(offset, n, slot);
}

dfb_status dfb_peer_wait(void) {
  DFB_INIT();
  if (!g_ready || g_pending == 0) return DFB_OK;
  launch_k(peer_wait_kernel, dim3(1), dim3(256), 0, compute_stream(), (const PeerDev*)g_peer_dev, peer_pending_take());
  DFB_LAUNCH_CHECK("peer_wait");
  return DFB_OK;
}

dfb_status dfb_peer_status(unsigned* error_word) {
  if (error_word) *error_word = g_error_host ? *g_error_host : 0u;
  return DFB_OK;
}

dfb_status dfb_peer_destroy(void) {
  if (!g_ready) return DFB_OK;
  cudaDeviceSynchronize();
  int one = 1;
  comm_allreduce_min_int(&one);   // every rank is past its last use of every arena
  return peer_release();
}

}  // extern "C"
