// Peer-memory gradient exchange (peer.cu): what the optimizer kernels (optim.cu) and the communicator (comm.cu) share.
#pragma once
#include "common.cuh"

namespace dfb {

constexpr int kPeerMaxWorld = 8;      // one NVLink / NVSwitch domain
constexpr int kPeerSlots = 64;        // buckets in flight (the last one belongs to the self-test)
constexpr int kPeerMaxCtas = 16;      // CTAs of one two-shot reduction launch: one per kPeerBytesPerCta of the rank's slice
constexpr size_t kPeerBytesPerCta = 128u << 10;
// one-shot "push" form for the bucket nothing overlaps (the last one of a step): every rank stores its copy into a
// receive area on every peer and sums all copies itself
constexpr int kPushMaxCtas = 32;
constexpr size_t kPushCapFloats = 64u << 10;      // 256 KB per copy
constexpr size_t kPushFlagBytes = (size_t)kPeerMaxWorld * kPushMaxCtas * sizeof(unsigned);
constexpr size_t kPushRecvBytes = 2 * (size_t)kPeerMaxWorld * 2 * kPushCapFloats * sizeof(float);   // [parity][source rank][copy as {value, launch number} pairs]
constexpr int kPeerThreads = 256;
constexpr int kPeerFlagWords = 16;    // per slot: [0, 8) phase-0 arrivals per source rank, [8] phase-1 arrivals
constexpr unsigned long long kSpinTimeoutNs = 20ull * 1000 * 1000 * 1000;

struct PeerDev {
  unsigned* flags[kPeerMaxWorld];   // every rank's flag words (flags[rank] = the local ones)
  float* arena[kPeerMaxWorld];      // every rank's gradient arena
  unsigned* epoch;                  // local: [slot] launches so far, then [kPeerSlots + slot] phase-1 arrivals expected so far
  unsigned* error;                  // mapped pinned host word, sticky
  unsigned* push_flags[kPeerMaxWorld];   // every rank's [source rank][cta] arrival words of the push form
  float* push_recv[kPeerMaxWorld];       // every rank's receive area
  unsigned* push_epoch;                  // local: [cta] launches so far
  int world, rank;
};

__host__ __device__ __forceinline__ int peer_flag_index(int slot, int phase, int src) {
  return slot * kPeerFlagWords + (phase ? 8 : src);
}

// host side
unsigned long long peer_pending_take();   // slots reduced since the last wait; the caller's kernel must wait for them
const PeerDev* peer_dev();
dfb_status comm_allgather_bytes(const void* mine, size_t bytes, void* all);   // host buffers, blocking
dfb_status comm_allreduce_min_int(int* value);                                 // host value, blocking

#ifdef __CUDACC__
// Spin until *p has reached `target` (wrap-safe), then acquire at system scope: what the signalling rank wrote before
// its release is visible to every thread that synchronises with the caller afterwards.
__device__ __forceinline__ void peer_spin(const unsigned* p, unsigned target, unsigned* error) {
  unsigned v, spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if ((int)(v - target) >= 0) break;
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > kSpinTimeoutNs) { *(volatile unsigned*)error = 1u; break; }
    }
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
}
// All threads of the CTA call this (blockDim.x >= kPeerSlots) before their first load of a reduced gradient: returns
// when every rank's slice of every bucket in `slots` has landed in the local arena.
__device__ __forceinline__ void peer_wait_slots(const PeerDev* __restrict__ pd, unsigned long long slots) {
  if (slots == 0) return;
  if (threadIdx.x < kPeerSlots && ((slots >> threadIdx.x) & 1ull)) {
    const int slot = threadIdx.x;
    const unsigned target = *(volatile const unsigned*)(pd->epoch + kPeerSlots + slot);
    peer_spin(pd->flags[pd->rank] + peer_flag_index(slot, 1, 0), target, pd->error);
  }
  __syncthreads();
}
#endif

}  // namespace dfb
