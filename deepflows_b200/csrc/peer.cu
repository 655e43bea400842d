// Gradient exchange over NVLink 5 / NVSwitch peer memory: the data-parallel all-reduce as kernels of this library.
//
// The reference has no distributed layer; round 1 reduced the buckets with ncclAllReduce on a
// communication stream (comm.cu, still the checked fallback). For buckets of 0.2 - 4 MB that path is all fixed cost.
// Here every rank owns one symmetric block (flag words, a receive area, the gradient arena: one cudaMalloc) that every
// other rank maps through CUDA IPC, the buckets are windows of the arena, and a bucket is reduced in one of two forms:
//
//   one-shot push (peer_push_allreduce_kernel) - the bucket nothing overlaps (the last one of a step, <= 256 KB), on the
//     compute stream between its pack and the optimizer: every rank stores its copy to every peer as {value, launch
//     number} pairs (self-validating 8-byte stores: no fence, no flag), polls the peers' pairs locally and sums all
//     copies in rank order. One NVLink crossing.
//   two-shot pull (peer_handshake_kernel + peer_allreduce_kernel) - the buckets that overlap backward, on the
//     communication stream:
//       phase 0  one warp tells every peer "rank r's copy of this bucket is packed" and waits until every peer has
//       reduce   rank r owns slice r of the bucket: a few CTAs read that slice from ALL ranks over NVLink (128-bit loads,
//                all peers' loads in flight together), add them in rank order, and store the sum into slice r of EVERY
//                rank's arena (reduce-scatter + all-gather in one pass; every element is computed by exactly one rank)
//       phase 1  each CTA adds 1 to every rank's arrival counter and EXITS - nobody waits here: multi_adam_kernel /
//                multi_sgd_kernel take the pending buckets as a bit mask and spin on the counters before their first
//                gradient load (peer_wait_slots); dfb_peer_wait() is the stand-alone form.
//
// In both forms every rank receives identical bits. Launch numbers instead of flag resets: they live in device memory,
// signals carry them or are monotonic adds, waits compare against them - a captured CUDA graph replays the same kernels
// with frozen arguments and stays correct. A spin that sees no progress for kSpinTimeoutNs sets a sticky error word
// (pinned host memory, dfb_peer_status) and falls through: a lost peer must not hang the GPU.
// dfb_peer_init() finishes with a self-test (a known pattern reduced twice through each form) and an NCCL all-reduce of
// its verdict, so either every rank uses this path or every rank stays on NCCL.
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
constexpr size_t kHeadBytes = kFlagBytes + kPushFlagBytes + kPushRecvBytes;   // in front of the arena in every rank's block
unsigned* g_push_epoch_dev = nullptr;
bool g_comm_dirty = false;   // reductions enqueued on the communication stream since the compute stream last joined it
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

// One-shot form for the bucket whose reduction nothing overlaps (the first-registered layers: <= 256 KB): ONE kernel, ONE
// NVLink crossing, no fence and no separate flag. Every 8 bytes that cross the link are {value, launch number} (the
// "LL" idea of NCCL's low-latency protocol: an 8-byte store becomes visible as a whole, so the value validates itself):
// a CTA stores its part of the local bucket into the receive area [parity][rank] of every peer as such pairs (128-bit
// stores of two pairs), then polls the pairs the peers stored here until they carry this launch's number, and adds the W
// copies in rank order into the local bucket - every rank computes every element, in the same order, so replicas hold
// identical bits. No handshake either: the receive area has two halves used by alternate launches, and a rank cannot be
// two launches ahead of a peer (it completes launch E - 1 only after the peer's launch E - 1 stores arrived, which the
// peer issued after consuming launch E - 2). Measured on 2 GPUs (225 KB): handshake kernel + pull + push + arrival
// counters awaited in the optimizer 30 us from the pack to Adam, data + fence + flag in one kernel 25 us (a system-scope
// store takes ~4 us to become visible on the peer, the fence waits for the round trip), NCCL 22 us.
__device__ __forceinline__ void st_ll(float* dst, float a, float b, unsigned E) {
  asm volatile("st.volatile.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__float_as_uint(a)), "r"(E), "r"(__float_as_uint(b)), "r"(E)
               : "memory");
}
__device__ __forceinline__ uint4 ld_ll_raw(const float* src) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(src) : "memory");
  return r;
}
// two values once both pairs carry launch number E
__device__ __forceinline__ float2 ld_ll(const float* src, unsigned E, unsigned* error) {
  unsigned a, fa, b, fb, spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile("ld.volatile.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(src) : "memory");
    if (fa == E && fb == E) break;
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > kSpinTimeoutNs) { *(volatile unsigned*)error = 1u; break; }
    }
  }
  return make_float2(__uint_as_float(a), __uint_as_float(b));
}
template <int WORLD>
__global__ void __launch_bounds__(kPeerThreads)
peer_push_allreduce_kernel(const PeerDev* __restrict__ pdp, unsigned long long off, unsigned long long n) {
  pdl_sync();
  __shared__ PeerDev pd;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) {
    pd = *pdp;
    // launch number: push_epoch[0], advanced by the last CTA of every launch (ticket in push_epoch[1]) - all CTAs of a
    // launch see the same value, so a location of the receive area changes halves with EVERY launch whatever the grid
    s_epoch = *(volatile const unsigned*)pdp->push_epoch + 1;
  }
  __syncthreads();
  const int rank = pd.rank, world = pd.world, c = blockIdx.x;   // world <= WORLD
  const unsigned E = s_epoch;
  // receive area: [half][source rank][2 * kPushCapFloats] (pairs: twice the floats)
  const size_t half = (size_t)(E & 1u) * kPeerMaxWorld * 2 * kPushCapFloats;
  float* bucket = pd.arena[rank] + off;
  const unsigned long long n4 = n >> 2;
  const unsigned long long per_cta = (n4 + gridDim.x - 1) / gridDim.x;
  const unsigned long long c_begin = min(n4, per_cta * c), c_end = min(n4, c_begin + per_cta);
  // ---- my copy -> every peer ----
  for (unsigned long long i = c_begin + threadIdx.x; i < c_end; i += kPeerThreads) {
    const float4 v = *reinterpret_cast<const float4*>(bucket + (i << 2));
#pragma unroll
    for (int q = 0; q < WORLD; ++q)
      if (q < world && q != rank) {
        float* dst = pd.push_recv[q] + half + (size_t)rank * 2 * kPushCapFloats + (i << 3);
        st_ll(dst, v.x, v.y, E);
        st_ll(dst + 4, v.z, v.w, E);
      }
  }
  // ---- sum of all copies, rank order (the local copy from the bucket, the others as they arrive) ----
  const float* recv = pd.push_recv[rank] + half;
  for (unsigned long long i = c_begin + threadIdx.x; i < c_end; i += kPeerThreads) {
    float4 v[WORLD];
    uint4 l0[WORLD], l1[WORLD];
    // first try: every peer's pairs in flight together; only what has not arrived yet is polled again
#pragma unroll
    for (int q = 0; q < WORLD; ++q)
      if (q < world && q != rank) {
        const float* src = recv + (size_t)q * 2 * kPushCapFloats + (i << 3);
        l0[q] = ld_ll_raw(src);
        l1[q] = ld_ll_raw(src + 4);
      }
#pragma unroll
    for (int q = 0; q < WORLD; ++q)
      if (q < world) {
        if (q == rank) {
          v[q] = *reinterpret_cast<const float4*>(bucket + (i << 2));
        } else {
          const float* src = recv + (size_t)q * 2 * kPushCapFloats + (i << 3);
          float2 lo = make_float2(__uint_as_float(l0[q].x), __uint_as_float(l0[q].z)), hi = make_float2(__uint_as_float(l1[q].x), __uint_as_float(l1[q].z));
          if (l0[q].y != E || l0[q].w != E) lo = ld_ll(src, E, pd.error);
          if (l1[q].y != E || l1[q].w != E) hi = ld_ll(src + 4, E, pd.error);
          v[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
        }
      }
    float4 s = v[0];
#pragma unroll
    for (int q = 1; q < WORLD; ++q)
      if (q < world) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
    *reinterpret_cast<float4*>(bucket + (i << 2)) = s;
  }
  if (threadIdx.x == 0) {
    if (atomicAdd(pd.push_epoch + 1, 1u) == gridDim.x - 1) {
      pd.push_epoch[1] = 0u;
      __threadfence();
      *(volatile unsigned*)pd.push_epoch = E;
    }
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
  if (m || g_comm_dirty) {
    cudaEventRecord(g_ev_done, comm_stream());
    cudaStreamWaitEvent(compute_stream(), g_ev_done, 0);
  }
  g_pending = 0;
  g_comm_dirty = false;
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
  if (g_push_epoch_dev) cudaFree(g_push_epoch_dev);
  g_push_epoch_dev = nullptr;
  g_comm_dirty = false;
  if (g_error_host) cudaFreeHost((void*)g_error_host);
  if (g_ev_fill) cudaEventDestroy(g_ev_fill);
  if (g_ev_done) cudaEventDestroy(g_ev_done);
  g_ev_fill = g_ev_done = nullptr;
  g_block = nullptr; g_peer_dev = nullptr; g_epoch_dev = nullptr; g_error_host = nullptr;
  g_ready = false; g_pending = 0; g_arena_floats = 0;
  return DFB_OK;
}

// the one-shot form (see peer_push_allreduce_kernel): complete when the kernel is, nothing for the consumer to await
// Runs on the CURRENT compute stream, right behind the pack and right in front of the optimizer: what it reduces is
// exposed anyway, and the two event hops to the communication stream and back cost more than the kernel (7 us each in
// the captured step).
static dfb_status peer_launch_push(size_t offset, size_t n) {
  const size_t pass_bytes = (size_t)kPeerThreads * 16 * 2;   // two positions per thread
  const int ctas = (int)std::max<size_t>(1, std::min<size_t>(kPushMaxCtas, (n * sizeof(float) + pass_bytes - 1) / pass_bytes));
  auto go = [&](auto kernel) {
    launch_k(kernel, dim3(ctas), dim3(kPeerThreads), 0, compute_stream(), (const PeerDev*)g_peer_dev, (unsigned long long)offset,
             (unsigned long long)n);
  };
  switch (g_peer.world) {
    case 2: go(peer_push_allreduce_kernel<2>); break;
    case 3: case 4: go(peer_push_allreduce_kernel<4>); break;
    default: go(peer_push_allreduce_kernel<8>); break;
  }
  DFB_LAUNCH_CHECK("peer_push_allreduce");
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
  // Few CTAs (one per 128 KB of the rank's slice, at most 16): these reductions overlap backward, where their latency is
  // free and their footprint is not - measured on 2 GPUs, backward ends 15 - 20 us EARLIER than beside NCCL's kernels with
  // 10 - 16 CTAs per bucket and 14 us LATER with 32. The bucket nothing overlaps takes the one-shot form instead.
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
  g_comm_dirty = true;
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
  const size_t bytes = kHeadBytes + arena_floats * sizeof(float);

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
  if (ok && (e = cudaMalloc((void**)&g_push_epoch_dev, kPushMaxCtas * sizeof(unsigned))) != cudaSuccess) fail("cudaMalloc", e);
  if (ok && (e = cudaMemset(g_push_epoch_dev, 0, kPushMaxCtas * sizeof(unsigned))) != cudaSuccess) fail("cudaMemset", e);
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
      g_peer.push_flags[q] = (unsigned*)((char*)base + kFlagBytes);
      g_peer.push_recv[q] = (float*)((char*)base + kFlagBytes + kPushFlagBytes);
      g_peer.arena[q] = (float*)((char*)base + kHeadBytes);
    }
  }
  if (ok) {
    g_peer.epoch = g_epoch_dev;
    g_peer.push_epoch = g_push_epoch_dev;
    void* err_dev = nullptr;
    if ((e = cudaHostGetDevicePointer(&err_dev, (void*)g_error_host, 0)) != cudaSuccess) fail("cudaHostGetDevicePointer", e);
    g_peer.error = (unsigned*)err_dev;
    if (ok && (e = cudaMemcpy(g_peer_dev, &g_peer, sizeof(g_peer), cudaMemcpyHostToDevice)) != cudaSuccess) fail("cudaMemcpy", e);
  }
  // Every rank must have mapped every peer before any kernel touches a peer: agree on that first (one rank that could not
  // open a handle would otherwise leave the others' self-test waiting for it, and the collectives below out of step).
  {
    int mapped = ok;
    st = comm_allreduce_min_int(&mapped);
    if (st != DFB_OK) { peer_release(); return st; }
    if (!mapped) {
      peer_release();
      DFB_FAIL(DFB_ERR_RUNTIME, "peer_init: peer-memory exchange unavailable (%s)", ok ? "another rank could not map its peers" : why.c_str());
    }
  }
  // ---- self-test: two launches of each form (the second proves the epoch / parity arithmetic), pattern checked on the device.
  // Every rank runs every round's collective whatever its own outcome, so the ranks stay in step ----
  unsigned bad_host = 0;
  if (ok) {
    g_ready = true;
    g_arena_floats = arena_floats;
    unsigned* bad = nullptr;
    if ((e = cudaMalloc((void**)&bad, sizeof(unsigned))) != cudaSuccess) fail("cudaMalloc", e);
    if (ok) cudaMemset(bad, 0, sizeof(unsigned));
    cudaStream_t s = compute_stream();
    for (int round = 0; round < 4; ++round) {
      if (ok) {
        peer_test_fill_kernel<<<64, 256, 0, s>>>(g_peer.arena[rank], test_floats, rank);
        if (round < 2) peer_launch(0, test_floats, kPeerSlots - 1);
        else peer_launch_push(0, test_floats);
        launch_k(peer_wait_kernel, dim3(1), dim3(256), 0, s, (const PeerDev*)g_peer_dev, peer_pending_take());
        peer_test_check_kernel<<<64, 256, 0, s>>>(g_peer.arena[rank], test_floats, world, bad);
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) fail("self-test", e);
      }
      // nobody may refill its bucket while a peer still reads it: the next round's fill is ordered behind every rank's
      // check by the all-gather below / the verdict all-reduce (last round)
      if (round < 3) {
        unsigned char z[kRec] = {0};
        st = comm_allgather_bytes(z, kRec, all.data());
        if (st != DFB_OK && ok) { ok = 0; why = "all-gather between the self-test rounds failed"; }
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

dfb_status dfb_peer_allreduce_async(size_t offset, size_t n, int slot, int exposed) {
  DFB_INIT();
  DFB_REQUIRE(g_ready, DFB_ERR_RUNTIME, "peer_allreduce: dfb_peer_init has not succeeded");
  DFB_REQUIRE(slot >= 0 && slot < kPeerSlots - 1, DFB_ERR_INVALID, "peer_allreduce: slot %d outside [0, %d)", slot, kPeerSlots - 1);
  DFB_REQUIRE((offset & 3) == 0 && (n & 3) == 0, DFB_ERR_INVALID, "peer_allreduce: offset and count must be multiples of 4 floats");
  DFB_REQUIRE(offset + n <= g_arena_floats, DFB_ERR_OUT_OF_RANGE, "peer_allreduce: [%zu, %zu) outside the arena of %zu floats", offset,
              offset + n, g_arena_floats);
  if (n == 0) return DFB_OK;
  if (exposed && n <= kPushCapFloats) return peer_launch_push(offset, n);
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
