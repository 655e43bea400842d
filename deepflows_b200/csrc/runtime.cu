// Runtime of libdfb200.so: device/stream ownership, caching device allocator, host<->device
// staging, events, CUDA-graph capture.
//
// Replaces what the reference does implicitly through the CUDA runtime on device 0 / stream 0
// with one cudaMalloc + cudaFree per temporary and blocking cudaMemcpy
// (reference: DeepFlows/backend/backend_src/ndarray_backend_cuda.cu:48-83, 667-716).
#include "common.cuh"

#include <mutex>
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

  // caching allocator: exact (rounded) size classes; blocks are never returned to the driver
  // unless dfb_empty_cache() is called or cudaMalloc fails.
  std::unordered_map<size_t, std::vector<void*>> free_blocks;
  std::unordered_map<void*, size_t> live;  // ptr -> rounded bytes
  size_t bytes_in_use = 0, bytes_reserved = 0, n_cuda_malloc = 0;

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
  r.ready = true;
  return DFB_OK;
}

cudaStream_t compute_stream() { return rt().compute; }
cudaStream_t comm_stream() { return rt().comm; }
int sm_count() { return rt().sms; }

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
  auto it = r.free_blocks.find(bytes);
  if (it != r.free_blocks.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
  } else {
    DFB_REQUIRE(!r.capturing, DFB_ERR_RUNTIME,
                "dfb_malloc: pool miss (%zu bytes) during CUDA-graph capture; run one warm-up "
                "step before capturing", bytes);
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
  r.live.erase(it);
  r.bytes_in_use -= bytes;
  // Single compute stream: every consumer of this block was enqueued before any later producer
  // that re-uses it, so the block can be recycled immediately without an event.
  r.free_blocks[bytes].push_back((void*)ptr);
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

// ---- host <-> device ------------------------------------------------------------------------
dfb_status dfb_from_host(const float* host_src, float* dst, size_t n) {
  DFB_INIT();
  if (n == 0) return DFB_OK;
  Runtime& r = rt();
  size_t bytes = n * sizeof(float);
  const size_t kMaxStage = size_t(64) << 20;
  if (bytes > kMaxStage || r.capturing) {
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

// ---- CUDA graph capture ---------------------------------------------------------------------
dfb_status dfb_graph_begin_capture(void) {
  DFB_INIT();
  Runtime& r = rt();
  DFB_REQUIRE(!r.capturing, DFB_ERR_RUNTIME, "graph capture already active");
  DFB_CUDA(cudaStreamSynchronize(r.compute));
  DFB_CUDA(cudaStreamBeginCapture(r.compute, cudaStreamCaptureModeThreadLocal));
  r.capturing = true;
  return DFB_OK;
}
dfb_status dfb_graph_end_capture(void** graph_exec) {
  Runtime& r = rt();
  DFB_REQUIRE(r.capturing, DFB_ERR_RUNTIME, "no graph capture active");
  cudaGraph_t g = nullptr;
  r.capturing = false;
  DFB_CUDA(cudaStreamEndCapture(r.compute, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) DFB_FAIL(DFB_ERR_RUNTIME, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  *graph_exec = (void*)ge;
  return DFB_OK;
}
dfb_status dfb_graph_launch(void* graph_exec) {
  DFB_INIT();
  DFB_CUDA(cudaGraphLaunch((cudaGraphExec_t)graph_exec, rt().compute));
  return DFB_OK;
}
dfb_status dfb_graph_destroy(void* graph_exec) {
  DFB_CUDA(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return DFB_OK;
}

}  // extern "C"
