// Runtime of libdfb200.so: device/stream ownership, caching device allocator, host<->device
// staging, events, CUDA-graph capture.
//
// Replaces what the reference does implicitly through the CUDA runtime on device 0 / stream 0
// with one cudaMalloc + cudaFree per temporary and blocking cudaMemcpy
// (reference: DeepFlows/backend/backend_src/ndarray_backend_cuda.cu:48-83, 667-716).
#include "kernels.cuh"

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace dfb {

static thread_local char tl_error[1024] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_tc_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

namespace {
struct Runtime {
  std::mutex mu;
  bool ready = false;
  int device = 0;
  int sms = 148;
  cudaStream_t compute = nullptr;
  cudaStream_t comm = nullptr;
  unsigned* tickets = nullptr;
  void* stat_slots = nullptr;   // kStatSlots x kStatSlotBytes (kernels.cuh)
  // side stream: independent work of one op (the wgrad of a conv backward) runs beside the main stream
  // between dfb_side_begin() and dfb_side_end(); dfb_side_join() orders the main stream after it.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // copy stream of the input pipeline: host -> device prefetch of the next batch beside the running step
  cudaStream_t copy = nullptr;
  // capture only: a branch that depends on nothing but the graph's root (graph_early_h2d)
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_root = nullptr, ev_aux = nullptr;
  cudaEvent_t ev_copy_ready = nullptr, ev_copy_done = nullptr;
  bool prefetch_pending = false;
  bool on_side = false;
  // Side tasks (side_begin .. side_end) are numbered; main-stream work is ordered after task j once it has waited for
  // ev_side[j % kSideEvents] (side_joined >= j). A block freed while some task is not joined yet may still be read by it:
  // it is parked with the number of the last task launched and recycled when the main stream has joined that task.
  static constexpr int kSideEvents = 64;
  cudaEvent_t ev_side[kSideEvents] = {};
  unsigned long long side_seq = 0, side_joined = 0;
  std::vector<std::pair<void*, unsigned long long>> side_frees;

  // caching allocator: exact (rounded) size classes; blocks are never returned to the driver
  // unless dfb_empty_cache() is called or cudaMalloc fails.
  std::unordered_map<size_t, std::vector<void*>> free_blocks;
  std::unordered_map<void*, size_t> live;  // ptr -> rounded bytes
  size_t bytes_in_use = 0, bytes_reserved = 0, n_cuda_malloc = 0;

  // CUDA-graph capture. Every block handed out while a capture is active becomes the property of
  // that graph: its address is baked into the captured kernel nodes, so when it is freed it goes to
  // the graph's private free lists (re-usable by later allocations of the same capture, never by
  // anybody else) until dfb_graph_destroy() returns the pool to the general one.
  struct GraphPool {
    std::unordered_map<size_t, std::vector<void*>> free_blocks;
    std::vector<void*> host_allocs;   // pinned staging owned by the graph (optimizer tables)
    std::vector<void*> dev_allocs;
    std::vector<void*> hyper_host;    // one per captured optimizer step, in capture order
    std::vector<int> hyper_kind;      // 0 = Adam, 1 = SGD
    // A replay reads its hyper-parameter block from pinned memory when it STARTS, which can be long after the host
    // launched it: the host must not write the next replay's values before. Every block carries a sequence number in its
    // last word; the replay echoes it into `hyper_ack` (a 4-byte device -> host copy behind the upload) and
    // graph_hyper_slot waits for the echo of the previous values before it hands the block out again.
    std::vector<volatile unsigned*> hyper_ack;
    std::vector<unsigned> hyper_seq;
    int n_nodes = 0, n_kernel_nodes = 0;
  };
  std::unordered_map<void*, GraphPool*> owner;          // block -> pool, for graph-owned blocks
  std::unordered_map<void*, GraphPool*> pools_by_exec;  // cudaGraphExec_t -> pool
  GraphPool* active_pool = nullptr;

  // pinned staging ring for dfb_from_host (H2D returns without a device sync)
  static constexpr int kSlots = 4;
  struct Slot {
    void* host = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool busy = false;
  } slots[kSlots];
  int next_slot = 0;
  bool capturing = false;
};
Runtime& rt() {
  static Runtime r;
  return r;
}

size_t round_block(size_t bytes) {
  if (bytes == 0) bytes = 1;
  if (bytes <= (1u << 20)) return (bytes + 511) & ~size_t(511);
  return (bytes + (1u << 20) - 1) & ~size_t((1u << 20) - 1);
}
}  // namespace

dfb_status ensure_init() {
  Runtime& r = rt();
  if (r.ready) return DFB_OK;
  std::lock_guard<std::mutex> lk(r.mu);
  if (r.ready) return DFB_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    DFB_FAIL(DFB_ERR_RUNTIME,
             "libdfb200: no usable CUDA device (%s). This backend has no CPU fallback.",
             e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  DFB_REQUIRE(r.device < n, DFB_ERR_RUNTIME, "libdfb200: device %d requested, %d present", r.device, n);
  DFB_CUDA(cudaSetDevice(r.device));
  cudaDeviceProp prop;
  DFB_CUDA(cudaGetDeviceProperties(&prop, r.device));
  DFB_REQUIRE(prop.major == 10, DFB_ERR_RUNTIME,
              "libdfb200 is built for sm_100a only; device %d is sm_%d%d (%s)", r.device, prop.major,
              prop.minor, prop.name);
  r.sms = prop.multiProcessorCount;
  DFB_CUDA(cudaStreamCreateWithFlags(&r.compute, cudaStreamNonBlocking));
  DFB_CUDA(cudaStreamCreateWithFlags(&r.comm, cudaStreamNonBlocking));
  DFB_CUDA(cudaStreamCreateWithFlags(&r.side, cudaStreamNonBlocking));
  DFB_CUDA(cudaEventCreateWithFlags(&r.ev_fork, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&r.ev_join, cudaEventDisableTiming));
  for (int i = 0; i < Runtime::kSideEvents; ++i) DFB_CUDA(cudaEventCreateWithFlags(&r.ev_side[i], cudaEventDisableTiming));
  DFB_CUDA(cudaStreamCreateWithFlags(&r.copy, cudaStreamNonBlocking));
  DFB_CUDA(cudaStreamCreateWithFlags(&r.aux, cudaStreamNonBlocking));
  DFB_CUDA(cudaEventCreateWithFlags(&r.ev_root, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&r.ev_aux, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&r.ev_copy_ready, cudaEventDisableTiming));
  DFB_CUDA(cudaEventCreateWithFlags(&r.ev_copy_done, cudaEventDisableTiming));
  DFB_CUDA(cudaMalloc(&r.tickets, kTicketWords * sizeof(unsigned)));
  DFB_CUDA(cudaMemset(r.tickets, 0, kTicketWords * sizeof(unsigned)));
  DFB_CUDA(cudaMalloc(&r.stat_slots, (size_t)kStatSlots * kStatSlotBytes));
  DFB_CUDA(cudaMemset(r.stat_slots, 0, (size_t)kStatSlots * kStatSlotBytes));
  DFB_CUDA(cudaDeviceSynchronize());
  r.ready = true;
  return DFB_OK;
}

cudaStream_t compute_stream() { return rt().on_side ? rt().side : rt().compute; }
cudaStream_t comm_stream() { return rt().comm; }
int sm_count() { return rt().sms; }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// ---- step timeline (common.cuh) ------------------------------------------------------------------------------------------
namespace {
struct TraceState {
  std::vector<void (*)(unsigned long long*)> setters;   // one per translation unit with kernels
  unsigned long long* dev = nullptr;
  size_t capacity = 0;
  bool armed = false;
  std::vector<std::string> host;   // "stream grid.x grid.y grid.z block name" per launch since dfb_trace_begin
};
TraceState& trace_state() {
  static TraceState t;
  return t;
}
}  // namespace
void trace_register_symbol(void (*setter)(unsigned long long*)) { trace_state().setters.push_back(setter); }
bool trace_host_armed() { return trace_state().armed; }
void trace_host_launch(const void* func, dim3 grid, dim3 block, cudaStream_t stream) {
  const char* name = nullptr;
  if (cudaFuncGetName(&name, func) != cudaSuccess || !name) name = "?";
  char line[512];
  snprintf(line, sizeof(line), "%s %u %u %u %u %s", stream == rt().side ? "side" : (stream == rt().compute ? "main" : "other"), grid.x, grid.y,
           grid.z, block.x, name);
  trace_state().host.emplace_back(line);
}

// ---- statistic slots (kernels.cuh) -----------------------------------------------------------------------------------
// Host-side bookkeeping only; the device memory of a free slot is all zeros (cleared at start-up, then by the last CTA of
// the slot's last consumer kernel, or - for a slot whose consumer never came - by a memset when it is reclaimed).
namespace {
struct StatSlotState {
  const float* key = nullptr;   // the statistics buffer of the producer call (what the consumer is given)
  int remaining = 0;            // consumer kernels still to come; 0 = free
  unsigned long long stamp = 0;
};
StatSlotState g_slot[kStatSlots];
unsigned long long g_slot_clock = 0;
bool stat_slots_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DFB_STAT_SLOTS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
double* slot_ptr(int i) { return reinterpret_cast<double*>(reinterpret_cast<char*>(rt().stat_slots) + (size_t)i * kStatSlotBytes); }
}  // namespace

void stat_slot_drop(const float* key) {
  if (!key) return;
  for (int i = 0; i < kStatSlots; ++i)
    if (g_slot[i].remaining > 0 && g_slot[i].key == key) {   // never consumed: its sums are still in the slot
      cudaMemsetAsync(slot_ptr(i), 0, kStatSlotBytes, compute_stream());
      g_slot[i] = StatSlotState{};
    }
}
double* stat_slot_acquire(const float* key, int consumers, int channels) {
  if (!stat_slots_enabled() || !key || consumers < 1 || channels > kStatSlotChannels || rt().on_side) return nullptr;
  stat_slot_drop(key);
  int pick = -1, oldest = -1;
  for (int i = 0; i < kStatSlots; ++i) {
    if (g_slot[i].remaining == 0) { pick = i; break; }
    if (oldest < 0 || g_slot[i].stamp < g_slot[oldest].stamp) oldest = i;
  }
  if (pick < 0) {   // every slot waits for a consumer that never came: take the oldest back
    pick = oldest;
    cudaMemsetAsync(slot_ptr(pick), 0, kStatSlotBytes, compute_stream());
  }
  g_slot[pick].key = key;
  g_slot[pick].remaining = consumers;
  g_slot[pick].stamp = ++g_slot_clock;
  return slot_ptr(pick);
}
double* stat_slot_take(const float* key, int* last) {
  *last = 0;
  if (!key || rt().on_side) return nullptr;
  for (int i = 0; i < kStatSlots; ++i)
    if (g_slot[i].remaining > 0 && g_slot[i].key == key) {
      if (--g_slot[i].remaining == 0) {
        *last = 1;
        g_slot[i].key = nullptr;
      }
      return slot_ptr(i);
    }
  return nullptr;
}

// work on the side stream may overlap main-stream kernels of the same family: it gets its own counters (32..63)
unsigned* ticket_counter(int slot) { return rt().tickets + slot + ((rt().on_side && slot < 32) ? 32 : 0); }

}  // namespace dfb

using namespace dfb;

extern "C" {

const char* dfb_last_error(void) { return tl_error; }
const char* dfb_version(void) { return "dfb200 0.1 (sm_100a)"; }

dfb_status dfb_set_device(int device) {
  Runtime& r = rt();
  DFB_REQUIRE(!r.ready || r.device == device, DFB_ERR_RUNTIME,
              "dfb_set_device(%d): runtime already initialised on device %d", device, r.device);
  DFB_REQUIRE(device >= 0, DFB_ERR_INVALID, "dfb_set_device: negative device");
  r.device = device;
  return DFB_OK;
}
dfb_status dfb_get_device(int* device) {
  *device = rt().device;
  return DFB_OK;
}
dfb_status dfb_device_count(int* count) {
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    DFB_FAIL(DFB_ERR_RUNTIME, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  return DFB_OK;
}
dfb_status dfb_device_info(char* name, size_t name_cap, int* sm_count_out, int* cc_major,
                           int* cc_minor, size_t* total_mem_bytes) {
  DFB_INIT();
  cudaDeviceProp prop;
  DFB_CUDA(cudaGetDeviceProperties(&prop, rt().device));
  if (name && name_cap) snprintf(name, name_cap, "%s", prop.name);
  if (sm_count_out) *sm_count_out = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem_bytes) *total_mem_bytes = prop.totalGlobalMem;
  return DFB_OK;
}
dfb_status dfb_synchronize(void) {
  DFB_INIT();
  DFB_CUDA(cudaStreamSynchronize(rt().compute));
  DFB_CUDA(cudaStreamSynchronize(rt().side));
  DFB_CUDA(cudaStreamSynchronize(rt().copy));
  DFB_CUDA(cudaStreamSynchronize(rt().comm));
  return DFB_OK;
}
void* dfb_stream(void) {
  if (ensure_init() != DFB_OK) return nullptr;
  return (void*)rt().compute;
}

// ---- events ---------------------------------------------------------------------------------
dfb_status dfb_event_create(void** ev) {
  DFB_INIT();
  cudaEvent_t e;
  DFB_CUDA(cudaEventCreate(&e));
  *ev = (void*)e;
  return DFB_OK;
}
dfb_status dfb_event_destroy(void* ev) {
  DFB_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return DFB_OK;
}
dfb_status dfb_event_record(void* ev) {
  DFB_INIT();
  DFB_CUDA(cudaEventRecord((cudaEvent_t)ev, rt().compute));
  return DFB_OK;
}
dfb_status dfb_event_synchronize(void* ev) {
  DFB_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return DFB_OK;
}
dfb_status dfb_event_elapsed_ms(void* start, void* stop, float* ms) {
  DFB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return DFB_OK;
}

// ---- allocator ------------------------------------------------------------------------------
dfb_status dfb_malloc(size_t n_floats, float** out_ptr) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(n_floats < (size_t(1) << 40), DFB_ERR_NOMEM, "dfb_malloc: absurd size %zu", n_floats);
  size_t bytes = round_block(n_floats * sizeof(float));
  std::lock_guard<std::mutex> lk(r.mu);
  void* p = nullptr;
  if (r.active_pool) {
    auto git = r.active_pool->free_blocks.find(bytes);
    if (git != r.active_pool->free_blocks.end() && !git->second.empty()) {
      p = git->second.back();
      git->second.pop_back();
    }
  }
  auto it = r.free_blocks.find(bytes);
  if (p) {
  } else if (it != r.free_blocks.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
  } else {
    // (the capture runs in relaxed mode, so cudaMalloc is legal while capturing)
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      // release everything cached and retry once
      cudaGetLastError();
      cudaStreamSynchronize(r.compute);
      for (auto& kv : r.free_blocks) {
        for (void* q : kv.second) {
          cudaFree(q);
          r.bytes_reserved -= kv.first;
        }
        kv.second.clear();
      }
      e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      DFB_FAIL(DFB_ERR_NOMEM,
               "CUDA device memory allocation failed: %s (requested size: %zu bytes, reserved %zu)",
               cudaGetErrorString(e), bytes, r.bytes_reserved);
    }
    r.bytes_reserved += bytes;
    r.n_cuda_malloc++;
  }
  r.live[p] = bytes;
  r.bytes_in_use += bytes;
  if (r.active_pool) r.owner[p] = r.active_pool;
  *out_ptr = (float*)p;
  return DFB_OK;
}

dfb_status dfb_free(float* ptr) {
  if (!ptr) return DFB_OK;
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.live.find((void*)ptr);
  DFB_REQUIRE(it != r.live.end(), DFB_ERR_INVALID, "dfb_free: pointer %p not owned by the pool", (void*)ptr);
  size_t bytes = it->second;
  if (r.on_side || r.side_joined < r.side_seq) {  // possibly still in use by side-stream work the main stream is not ordered after yet
    r.side_frees.emplace_back((void*)ptr, r.on_side ? r.side_seq + 1 : r.side_seq);
    return DFB_OK;
  }
  r.live.erase(it);
  r.bytes_in_use -= bytes;
  // Single compute stream: every consumer of this block was enqueued before any later producer
  // that re-uses it, so the block can be recycled immediately without an event.
  auto ow = r.owner.find((void*)ptr);
  if (ow != r.owner.end()) ow->second->free_blocks[bytes].push_back((void*)ptr);
  else r.free_blocks[bytes].push_back((void*)ptr);
  return DFB_OK;
}

dfb_status dfb_empty_cache(void) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_CUDA(cudaStreamSynchronize(r.compute));
  DFB_CUDA(cudaStreamSynchronize(r.comm));
  std::lock_guard<std::mutex> lk(r.mu);
  for (auto& kv : r.free_blocks) {
    for (void* q : kv.second) {
      cudaFree(q);
      r.bytes_reserved -= kv.first;
    }
    kv.second.clear();
  }
  return DFB_OK;
}

dfb_status dfb_mem_stats(size_t* bytes_in_use, size_t* bytes_reserved, size_t* n_cuda_malloc) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mu);
  if (bytes_in_use) *bytes_in_use = r.bytes_in_use;
  if (bytes_reserved) *bytes_reserved = r.bytes_reserved;
  if (n_cuda_malloc) *n_cuda_malloc = r.n_cuda_malloc;
  return DFB_OK;
}

uint64_t dfb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
uint64_t dfb_tc_launch_count(void) { return g_tc_launches.load(std::memory_order_relaxed); }

dfb_status dfb_trace_begin(size_t capacity) {
  DFB_INIT();
  TraceState& t = trace_state();
  DFB_REQUIRE(capacity > 0 && capacity <= (1u << 22), DFB_ERR_INVALID, "trace_begin: capacity out of range");
  DFB_CUDA(cudaDeviceSynchronize());
  if (t.dev) cudaFree(t.dev);
  t.dev = nullptr;
  DFB_CUDA(cudaMalloc(&t.dev, (2 + 2 * capacity) * sizeof(unsigned long long)));
  DFB_CUDA(cudaMemset(t.dev, 0, (2 + 2 * capacity) * sizeof(unsigned long long)));
  const unsigned long long cap = capacity;
  DFB_CUDA(cudaMemcpy(t.dev + 1, &cap, sizeof(cap), cudaMemcpyHostToDevice));
  t.capacity = capacity;
  t.host.clear();
  for (auto set : t.setters) set(t.dev);
  DFB_CUDA(cudaDeviceSynchronize());
  t.armed = true;
  return DFB_OK;
}
dfb_status dfb_trace_reset(void) {   // forget what was recorded so far (e.g. the capture pass), stay armed
  TraceState& t = trace_state();
  DFB_REQUIRE(t.dev != nullptr, DFB_ERR_RUNTIME, "trace_reset: no trace armed");
  DFB_CUDA(cudaDeviceSynchronize());
  DFB_CUDA(cudaMemset(t.dev, 0, sizeof(unsigned long long)));
  return DFB_OK;
}
dfb_status dfb_trace_end(unsigned long long* records, size_t capacity, size_t* count) {
  TraceState& t = trace_state();
  DFB_REQUIRE(t.dev != nullptr && count != nullptr, DFB_ERR_RUNTIME, "trace_end: no trace armed");
  DFB_CUDA(cudaDeviceSynchronize());
  for (auto set : t.setters) set(nullptr);
  t.armed = false;
  unsigned long long n = 0;
  DFB_CUDA(cudaMemcpy(&n, t.dev, sizeof(n), cudaMemcpyDeviceToHost));
  if (n > t.capacity) n = t.capacity;
  if (n > capacity) n = capacity;
  if (records && n) DFB_CUDA(cudaMemcpy(records, t.dev + 2, 2 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  *count = (size_t)n;
  cudaFree(t.dev);
  t.dev = nullptr;
  return DFB_OK;
}
size_t dfb_trace_host_count(void) { return trace_state().host.size(); }
const char* dfb_trace_host_line(size_t i) { return i < trace_state().host.size() ? trace_state().host[i].c_str() : ""; }

// ---- host <-> device ------------------------------------------------------------------------
dfb_status dfb_from_host(const float* host_src, float* dst, size_t n) {
  DFB_INIT();
  if (n == 0) return DFB_OK;
  Runtime& r = rt();
  DFB_REQUIRE(!r.capturing, DFB_ERR_RUNTIME,
              "from_numpy during CUDA-graph capture: a captured step must not copy host data (keep inputs in "
              "device buffers that are refreshed before each replay)");
  size_t bytes = n * sizeof(float);
  const size_t kMaxStage = size_t(64) << 20;
  if (bytes > kMaxStage) {
    DFB_CUDA(cudaMemcpyAsync(dst, host_src, bytes, cudaMemcpyHostToDevice, r.compute));
    DFB_CUDA(cudaStreamSynchronize(r.compute));
    return DFB_OK;
  }
  // stage through a library-owned pinned slot: the caller's buffer is free on return and the
  // DMA overlaps with whatever the host does next.
  Runtime::Slot& s = r.slots[r.next_slot];
  r.next_slot = (r.next_slot + 1) % Runtime::kSlots;
  if (s.busy) {
    DFB_CUDA(cudaEventSynchronize(s.done));
    s.busy = false;
  }
  if (s.cap < bytes) {
    if (s.host) cudaFreeHost(s.host);
    size_t cap = bytes < (size_t(1) << 20) ? (size_t(1) << 20) : round_block(bytes);
    DFB_CUDA(cudaHostAlloc(&s.host, cap, cudaHostAllocDefault));
    s.cap = cap;
  }
  if (!s.done) DFB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  memcpy(s.host, host_src, bytes);
  DFB_CUDA(cudaMemcpyAsync(dst, s.host, bytes, cudaMemcpyHostToDevice, r.compute));
  DFB_CUDA(cudaEventRecord(s.done, r.compute));
  s.busy = true;
  return DFB_OK;
}

dfb_status dfb_to_host(const float* src, float* host_dst, size_t n) {
  DFB_INIT();
  if (n == 0) return DFB_OK;
  Runtime& r = rt();
  DFB_REQUIRE(!r.capturing, DFB_ERR_RUNTIME,
              "to_numpy during CUDA-graph capture: a captured step must not read results back (read them after "
              "the replay)");
  DFB_CUDA(cudaMemcpyAsync(host_dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, r.compute));
  DFB_CUDA(cudaStreamSynchronize(r.compute));
  return DFB_OK;
}

dfb_status dfb_host_alloc_pinned(size_t n_floats, float** host_ptr) {
  DFB_INIT();
  void* p = nullptr;
  DFB_CUDA(cudaHostAlloc(&p, n_floats * sizeof(float) + 16, cudaHostAllocDefault));
  *host_ptr = (float*)p;
  return DFB_OK;
}
dfb_status dfb_host_free_pinned(float* host_ptr) {
  DFB_CUDA(cudaFreeHost(host_ptr));
  return DFB_OK;
}
dfb_status dfb_from_host_async(const float* pinned_src, float* dst, size_t n) {
  DFB_INIT();
  DFB_CUDA(cudaMemcpyAsync(dst, pinned_src, n * sizeof(float), cudaMemcpyHostToDevice, rt().compute));
  return DFB_OK;
}
dfb_status dfb_to_host_async(const float* src, float* pinned_dst, size_t n) {
  DFB_INIT();
  DFB_CUDA(cudaMemcpyAsync(pinned_dst, src, n * sizeof(float), cudaMemcpyDeviceToHost, rt().compute));
  return DFB_OK;
}
dfb_status dfb_copy(const float* src, float* dst, size_t n) {
  DFB_INIT();
  if (n == 0) return DFB_OK;
  DFB_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, rt().compute));
  return DFB_OK;
}

// ---- input prefetch ---------------------------------------------------------------------------
// dfb_prefetch_from_host: the copy stream first waits for everything enqueued on the compute stream so far
// (the consumers of the destination's previous contents), then copies; dfb_prefetch_wait orders the compute
// stream after all prefetches enqueued so far. Between the two calls the copy runs beside the compute stream.
dfb_status dfb_prefetch_from_host(const float* pinned_src, float* dst, size_t n) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(!r.capturing, DFB_ERR_RUNTIME, "prefetch_from_host during CUDA-graph capture");
  if (n == 0) return DFB_OK;
  DFB_CUDA(cudaEventRecord(r.ev_copy_ready, r.compute));
  DFB_CUDA(cudaStreamWaitEvent(r.copy, r.ev_copy_ready, 0));
  DFB_CUDA(cudaMemcpyAsync(dst, pinned_src, n * sizeof(float), cudaMemcpyHostToDevice, r.copy));
  r.prefetch_pending = true;
  return DFB_OK;
}
dfb_status dfb_prefetch_wait(void) {
  DFB_INIT();
  Runtime& r = rt();
  if (!r.prefetch_pending) return DFB_OK;
  DFB_CUDA(cudaEventRecord(r.ev_copy_done, r.copy));
  DFB_CUDA(cudaStreamWaitEvent(r.compute, r.ev_copy_done, 0));
  r.prefetch_pending = false;
  return DFB_OK;
}

// ---- side stream ------------------------------------------------------------------------------
dfb_status dfb_side_begin(void) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(!r.on_side, DFB_ERR_RUNTIME, "side_begin: already on the side stream");
  DFB_CUDA(cudaEventRecord(r.ev_fork, r.compute));
  DFB_CUDA(cudaStreamWaitEvent(r.side, r.ev_fork, 0));
  r.on_side = true;
  return DFB_OK;
}
// the main stream waits for side task `upto` (and with it every earlier one); blocks parked for those tasks are recycled
static dfb_status side_join_upto(unsigned long long upto) {
  Runtime& r = rt();
  if (upto > r.side_seq) upto = r.side_seq;
  if (upto <= r.side_joined) return DFB_OK;
  DFB_CUDA(cudaStreamWaitEvent(r.compute, r.ev_side[upto % Runtime::kSideEvents], 0));
  std::vector<void*> frees;
  {
    std::lock_guard<std::mutex> lk(r.mu);
    r.side_joined = upto;
    size_t keep = 0;
    for (size_t i = 0; i < r.side_frees.size(); ++i) {
      if (r.side_frees[i].second <= upto) frees.push_back(r.side_frees[i].first);
      else r.side_frees[keep++] = r.side_frees[i];
    }
    r.side_frees.resize(keep);
  }
  if (r.side_joined == r.side_seq) {
    for (void* p : frees) dfb_free((float*)p);
  } else {   // later tasks are still out: dfb_free would park the blocks again, recycle them directly
    std::lock_guard<std::mutex> lk(r.mu);
    for (void* p : frees) {
      auto it = r.live.find(p);
      if (it == r.live.end()) continue;
      const size_t bytes = it->second;
      r.live.erase(it);
      r.bytes_in_use -= bytes;
      auto ow = r.owner.find(p);
      if (ow != r.owner.end()) ow->second->free_blocks[bytes].push_back(p);
      else r.free_blocks[bytes].push_back(p);
    }
  }
  return DFB_OK;
}
dfb_status dfb_side_end(void) {
  Runtime& r = rt();
  DFB_REQUIRE(r.on_side, DFB_ERR_RUNTIME, "side_end: not on the side stream");
  r.on_side = false;
  ++r.side_seq;
  DFB_CUDA(cudaEventRecord(r.ev_side[r.side_seq % Runtime::kSideEvents], r.side));
  if (r.side_seq - r.side_joined >= Runtime::kSideEvents / 2) return side_join_upto(r.side_seq - Runtime::kSideEvents / 4);   // events are a ring
  return DFB_OK;
}
dfb_status dfb_side_join(void) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(!r.on_side, DFB_ERR_RUNTIME, "side_join: still on the side stream");
  return side_join_upto(r.side_seq);
}
dfb_status dfb_side_join_lag(int lag) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(!r.on_side, DFB_ERR_RUNTIME, "side_join_lag: still on the side stream");
  DFB_REQUIRE(lag >= 0, DFB_ERR_INVALID, "side_join_lag: negative lag");
  if (r.side_seq <= (unsigned long long)lag) return DFB_OK;
  return side_join_upto(r.side_seq - (unsigned long long)lag);
}

// ---- CUDA graph capture ---------------------------------------------------------------------
}  // extern "C"

namespace dfb {
bool graph_capturing() { return rt().capturing; }
// A host -> device copy of a captured step that depends only on the graph's root: it runs while the step's first kernels
// do, and the stream that needs the data (the optimizer's table, a bucket's pointer list) only joins it. As a node on the
// compute stream itself the copy sat between the last backward kernel and the optimizer (~25 us of DMA latency per step).
dfb_status graph_early_h2d(void* dev, const void* host, size_t bytes, void* ack_host, size_t ack_offset) {
  Runtime& r = rt();
  DFB_REQUIRE(r.capturing, DFB_ERR_RUNTIME, "graph_early_h2d outside a capture");
  DFB_CUDA(cudaStreamWaitEvent(r.aux, r.ev_root, 0));
  DFB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, r.aux));
  if (ack_host)   // echo of the block's sequence number: "these values have been read" (GraphPool::hyper_ack)
    DFB_CUDA(cudaMemcpyAsync(ack_host, (const char*)dev + ack_offset, sizeof(unsigned), cudaMemcpyDeviceToHost, r.aux));
  DFB_CUDA(cudaEventRecord(r.ev_aux, r.aux));
  DFB_CUDA(cudaStreamWaitEvent(compute_stream(), r.ev_aux, 0));
  return DFB_OK;
}
dfb_status graph_staging(size_t bytes, int kind, void** host, void** dev) {
  Runtime& r = rt();
  DFB_REQUIRE(r.active_pool != nullptr, DFB_ERR_RUNTIME, "graph_staging outside a capture");
  DFB_CUDA(cudaHostAlloc(host, bytes, cudaHostAllocDefault));
  DFB_CUDA(cudaMalloc(dev, bytes));
  std::lock_guard<std::mutex> lk(r.mu);
  r.active_pool->host_allocs.push_back(*host);
  r.active_pool->dev_allocs.push_back(*dev);
  if (kind == 0 || kind == 1) {  // optimizer steps are addressable afterwards (dfb_graph_set_adam / _sgd)
    void* ack = nullptr;
    if (cudaHostAlloc(&ack, 64, cudaHostAllocDefault) != cudaSuccess) DFB_FAIL(DFB_ERR_NOMEM, "graph_staging: pinned allocation failed");
    *reinterpret_cast<volatile unsigned*>(ack) = 0u;
    r.active_pool->host_allocs.push_back(ack);
    r.active_pool->hyper_host.push_back(*host);
    r.active_pool->hyper_kind.push_back(kind);
    r.active_pool->hyper_ack.push_back(reinterpret_cast<volatile unsigned*>(ack));
    r.active_pool->hyper_seq.push_back(1u);   // the values written during the capture are sequence 1
  }
  return DFB_OK;
}
void* graph_last_hyper_ack() {
  Runtime& r = rt();
  return (r.active_pool && !r.active_pool->hyper_ack.empty()) ? (void*)r.active_pool->hyper_ack.back() : nullptr;
}
dfb_status graph_hyper_slot(void* graph_exec, int index, int kind, void** host) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.pools_by_exec.find(graph_exec);
  DFB_REQUIRE(it != r.pools_by_exec.end(), DFB_ERR_INVALID, "unknown graph handle");
  DFB_REQUIRE(index >= 0 && index < (int)it->second->hyper_host.size(), DFB_ERR_OUT_OF_RANGE,
              "graph has %zu captured optimizer steps, index %d requested", it->second->hyper_host.size(), index);
  DFB_REQUIRE(it->second->hyper_kind[index] == kind, DFB_ERR_INVALID, "captured optimizer step %d is of another kind", index);
  *host = it->second->hyper_host[index];
  // the replay that reads the current values must have read them (it echoes their sequence number when it starts);
  // a graph that was never launched with them (or an idle stream) has nothing in flight
  volatile unsigned* ack = it->second->hyper_ack[index];
  const unsigned want = it->second->hyper_seq[index];
  for (unsigned long long spins = 0; *ack != want; ++spins) {
    if ((spins & 0xfff) == 0xfff && cudaStreamQuery(r.compute) == cudaSuccess) break;
  }
  const unsigned next = want + 1u;
  it->second->hyper_seq[index] = next;
  memcpy((char*)*host + kGraphHyperSeqOffset, &next, sizeof(next));
  return DFB_OK;
}
}  // namespace dfb

extern "C" {

dfb_status dfb_graph_begin_capture(void) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(!r.capturing, DFB_ERR_RUNTIME, "graph capture already active");
  DFB_REQUIRE(!r.on_side, DFB_ERR_RUNTIME, "graph capture cannot begin on the side stream");
  {   // side tasks of eager work must not leak into the capture (a join inside it would wait for an uncaptured event)
    dfb_status st = side_join_upto(r.side_seq);
    if (st != DFB_OK) return st;
  }
  DFB_CUDA(cudaStreamSynchronize(r.compute));
  DFB_CUDA(cudaStreamSynchronize(r.comm));
  DFB_CUDA(cudaStreamBeginCapture(r.compute, cudaStreamCaptureModeRelaxed));
  DFB_CUDA(cudaEventRecord(r.ev_root, r.compute));
  std::lock_guard<std::mutex> lk(r.mu);
  r.active_pool = new Runtime::GraphPool();
  r.capturing = true;
  return DFB_OK;
}
dfb_status dfb_graph_end_capture(void** graph_exec) {
  Runtime& r = rt();
  DFB_REQUIRE(r.capturing, DFB_ERR_RUNTIME, "no graph capture active");
  if (!r.on_side) side_join_upto(r.side_seq);   // every forked branch has to be back on the capturing stream
  cudaGraph_t g = nullptr;
  Runtime::GraphPool* pool;
  {
    std::lock_guard<std::mutex> lk(r.mu);
    r.capturing = false;
    pool = r.active_pool;
    r.active_pool = nullptr;
  }
  cudaError_t e = cudaStreamEndCapture(r.compute, &g);
  cudaGraphExec_t ge = nullptr;
  if (e == cudaSuccess) {
    size_t n = 0;
    if (cudaGraphGetNodes(g, nullptr, &n) == cudaSuccess && n > 0) {
      std::vector<cudaGraphNode_t> nodes(n);
      cudaGraphGetNodes(g, nodes.data(), &n);
      pool->n_nodes = (int)n;
      for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel) pool->n_kernel_nodes++;
      }
    }
    e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    // the pool's blocks stay owned by a graph that does not exist: hand them back
    std::lock_guard<std::mutex> lk(r.mu);
    for (auto it = r.owner.begin(); it != r.owner.end();) it = it->second == pool ? r.owner.erase(it) : std::next(it);
    for (auto& kv : pool->free_blocks)
      for (void* q : kv.second) r.free_blocks[kv.first].push_back(q);
    for (void* h : pool->host_allocs) cudaFreeHost(h);
    for (void* d : pool->dev_allocs) cudaFree(d);
    delete pool;
    DFB_FAIL(DFB_ERR_RUNTIME, "CUDA-graph capture failed: %s", cudaGetErrorString(e));
  }
  std::lock_guard<std::mutex> lk(r.mu);
  r.pools_by_exec[(void*)ge] = pool;
  *graph_exec = (void*)ge;
  return DFB_OK;
}
dfb_status dfb_graph_launch(void* graph_exec) {
  DFB_INIT();
  DFB_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, rt().compute));
  return DFB_OK;
}
dfb_status dfb_graph_destroy(void* graph_exec) {
  Runtime& r = rt();
  DFB_CUDA(cudaStreamSynchronize(r.compute));
  DFB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.pools_by_exec.find(graph_exec);
  if (it == r.pools_by_exec.end()) return DFB_OK;
  Runtime::GraphPool* pool = it->second;
  r.pools_by_exec.erase(it);
  for (auto ow = r.owner.begin(); ow != r.owner.end();) ow = ow->second == pool ? r.owner.erase(ow) : std::next(ow);
  for (auto& kv : pool->free_blocks)
    for (void* q : kv.second) r.free_blocks[kv.first].push_back(q);
  for (void* h : pool->host_allocs) cudaFreeHost(h);
  for (void* d : pool->dev_allocs) cudaFree(d);
  delete pool;
  return DFB_OK;
}
dfb_status dfb_graph_node_counts(void* graph_exec, int* kernel_nodes, int* all_nodes) {
  Runtime& r = rt();
  std::lock_guard<std::mutex> lk(r.mu);
  auto it = r.pools_by_exec.find(graph_exec);
  DFB_REQUIRE(it != r.pools_by_exec.end(), DFB_ERR_INVALID, "unknown graph handle");
  if (kernel_nodes) *kernel_nodes = it->second->n_kernel_nodes;
  if (all_nodes) *all_nodes = it->second->n_nodes;
  return DFB_OK;
}
dfb_status dfb_graph_capturing(int* capturing) {
  *capturing = rt().capturing ? 1 : 0;
  return DFB_OK;
}

}  // extern "C"
