// Multi-tensor optimizer steps: one launch updates every parameter tensor.
//
// The reference walks the parameter list in Python and issues ~14 (Adam) / ~6 (SGD) out-of-place
// BackendTensor kernels plus as many allocations per parameter
// (DeepFlows/optim/adam.py:28-63, DeepFlows/optim/sgd.py:16-24). Here the host packs a table of
// device pointers, uploads it once per step through a pinned staging buffer, and a single
// grid-stride kernel walks fixed-size chunks of all tensors. The arithmetic follows the
// reference's operation order in fp32 (no FMA contraction) so a step is reproducible against it.
#include "common.cuh"
#include "peer.cuh"

#include <vector>

namespace dfb {

constexpr int kChunk = 4096;  // elements per work item (16 per thread at 256 threads, as float4)

struct TensorRef {
  float* p;
  const float* g;
  float* m;   // Adam first moment / SGD velocity
  float* v;   // Adam second moment
  unsigned long long size;
  unsigned long long first_chunk;  // prefix sum of chunk counts
};

struct AdamHyper {
  float lr, beta1, beta2, one_minus_beta1, one_minus_beta2, bias1, bias2, eps, weight_decay, grad_scale;
  int use_wd, use_scale;
};
struct SgdHyper {
  float lr, momentum, weight_decay, grad_scale;
  int use_momentum, nesterov, use_scale;
};

__device__ __forceinline__ int find_tensor(const TensorRef* __restrict__ t, int count, unsigned long long chunk) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (t[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, const AdamHyper& h) {
  if (h.use_scale) g = __fmul_rn(g, h.grad_scale);
  if (h.use_wd) g = __fadd_rn(g, __fmul_rn(p, h.weight_decay));           // grad + p * wd
  m = __fadd_rn(__fmul_rn(m, h.beta1), __fmul_rn(g, h.one_minus_beta1));   // v*b1 + g*(1-b1)
  v = __fadd_rn(__fmul_rn(v, h.beta2), __fmul_rn(__fmul_rn(g, g), h.one_minus_beta2));
  float mh = __fdiv_rn(m, h.bias1);                                         // v / (1 - b1**t)
  float vh = __fdiv_rn(v, h.bias2);
  float upd = __fmul_rn(__fdiv_rn(mh, __fadd_rn(sqrtf(vh), h.eps)), h.lr);  // v_hat/(s_hat**.5+eps)*lr
  p = __fadd_rn(p, -upd);
  return p;
}

__global__ void __launch_bounds__(256)
multi_adam_kernel(const TensorRef* __restrict__ tensors, int count, unsigned long long total_chunks,
                  const AdamHyper* __restrict__ hp, const PeerDev* __restrict__ peer, unsigned long long peer_slots) {
  pdl_sync();
  const AdamHyper h = *hp;
  // data parallel over peer memory (peer.cu): the gradients are bucket views, and the buckets named by `peer_slots` are
  // complete once every rank's slice has landed - awaited here, behind the launch and the hyper-parameter load
  peer_wait_slots(peer, peer_slots);
  for (unsigned long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int ti = find_tensor(tensors, count, chunk);
    TensorRef t = tensors[ti];
    unsigned long long begin = (chunk - t.first_chunk) * kChunk;
    unsigned long long end = begin + kChunk < t.size ? begin + kChunk : t.size;
    bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                 reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
    if (vec) {
      for (unsigned long long i = begin + threadIdx.x * 4ull; i + 4 <= end; i += blockDim.x * 4ull) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        adam_one(p.x, g.x, m.x, v.x, h); adam_one(p.y, g.y, m.y, v.y, h);
        adam_one(p.z, g.z, m.z, v.z, h); adam_one(p.w, g.w, m.w, v.w, h);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
      }
    }
    // scalar part: the <4-element tail of an aligned tensor, or everything when unaligned
    unsigned long long sbeg = vec ? begin + ((end - begin) & ~3ull) : begin;
    for (unsigned long long q = sbeg + threadIdx.x; q < end; q += blockDim.x)
      adam_one(t.p[q], t.g[q], t.m[q], t.v[q], h);
  }
}

__device__ __forceinline__ void sgd_one(float& p, float g, float* vel, const SgdHyper& h) {
  if (h.use_scale) g = __fmul_rn(g, h.grad_scale);
  g = __fadd_rn(g, __fmul_rn(p, h.weight_decay));  // always, even for wd == 0 (sgd.py:18)
  float upd = g;
  if (h.use_momentum) {
    float v = __fadd_rn(__fmul_rn(*vel, h.momentum), g);
    *vel = v;
    upd = h.nesterov ? __fadd_rn(g, __fmul_rn(v, h.momentum)) : v;
  }
  p = __fadd_rn(p, -__fmul_rn(upd, h.lr));
}

__global__ void __launch_bounds__(256)
multi_sgd_kernel(const TensorRef* __restrict__ tensors, int count, unsigned long long total_chunks,
                 const SgdHyper* __restrict__ hp, const PeerDev* __restrict__ peer, unsigned long long peer_slots) {
  pdl_sync();
  const SgdHyper h = *hp;
  peer_wait_slots(peer, peer_slots);
  for (unsigned long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int ti = find_tensor(tensors, count, chunk);
    TensorRef t = tensors[ti];
    unsigned long long begin = (chunk - t.first_chunk) * kChunk;
    unsigned long long end = begin + kChunk < t.size ? begin + kChunk : t.size;
    for (unsigned long long i = begin + threadIdx.x; i < end; i += blockDim.x) {
      float p = t.p[i];
      float vel = h.use_momentum ? t.m[i] : 0.f;
      sgd_one(p, t.g[i], &vel, h);
      t.p[i] = p;
      if (h.use_momentum) t.m[i] = vel;
    }
  }
}

// Many flat device-to-device copies in one launch (gradient buckets of the data-parallel layer: one launch per
// bucket instead of one strided-copy kernel per parameter). TensorRef::p = destination, ::g = source.
__global__ void __launch_bounds__(256)
multi_copy_kernel(const TensorRef* __restrict__ tensors, int count, unsigned long long total_chunks) {
  pdl_sync();
  for (unsigned long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int ti = find_tensor(tensors, count, chunk);
    TensorRef t = tensors[ti];
    unsigned long long begin = (chunk - t.first_chunk) * kChunk;
    unsigned long long end = begin + kChunk < t.size ? begin + kChunk : t.size;
    bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g)) & 15) == 0;
    if (vec) {
      for (unsigned long long i = begin + threadIdx.x * 4ull; i + 4 <= end; i += blockDim.x * 4ull)
        *reinterpret_cast<float4*>(t.p + i) = *reinterpret_cast<const float4*>(t.g + i);
    }
    unsigned long long sbeg = vec ? begin + ((end - begin) & ~3ull) : begin;
    for (unsigned long long q = sbeg + threadIdx.x; q < end; q += blockDim.x) t.p[q] = t.g[q];
  }
}

// Staging of one optimizer step: [hyper-parameters, 256 bytes][pointer table]. The kernel reads both from
// device memory. Eager steps go through a small pinned ring so consecutive steps do not wait on each other;
// a step issued during CUDA-graph capture gets staging that the graph owns, and the captured memcpy node
// re-reads the pinned copy at every replay - dfb_graph_set_adam / dfb_graph_set_sgd rewrite the hyper-
// parameters there between replays (learning-rate schedules, Adam's bias corrections).
constexpr size_t kHyperBytes = 256;
struct TableStage {
  static constexpr int kSlots = 4;
  void* host[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  void* dev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap[kSlots] = {0, 0, 0, 0};
  cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  int next = 0;
};
static TableStage g_stage;

static dfb_status upload_table(const std::vector<TensorRef>& tab, const void* hyper, size_t hyper_bytes, int kind,
                               const TensorRef** dev_table, const void** dev_hyper) {
  size_t bytes = kHyperBytes + tab.size() * sizeof(TensorRef);
  void *host = nullptr, *dev = nullptr;
  if (graph_capturing()) {
    dfb_status st = graph_staging(bytes, kind, &host, &dev);
    if (st != DFB_OK) return st;
  } else {
    TableStage& st = g_stage;
    int s = st.next;
    st.next = (st.next + 1) % TableStage::kSlots;
    if (st.done[s]) DFB_CUDA(cudaEventSynchronize(st.done[s]));
    if (st.cap[s] < bytes) {
      if (st.host[s]) cudaFreeHost(st.host[s]);
      if (st.dev[s]) cudaFree(st.dev[s]);
      size_t cap = bytes < 16384 ? 16384 : bytes * 2;
      DFB_CUDA(cudaHostAlloc(&st.host[s], cap, cudaHostAllocDefault));
      DFB_CUDA(cudaMalloc(&st.dev[s], cap));
      st.cap[s] = cap;
    }
    if (!st.done[s]) DFB_CUDA(cudaEventCreateWithFlags(&st.done[s], cudaEventDisableTiming));
    host = st.host[s];
    dev = st.dev[s];
  }
  memcpy(host, hyper, hyper_bytes);
  memcpy((char*)host + kHyperBytes, tab.data(), tab.size() * sizeof(TensorRef));
  if (graph_capturing()) {
    static_assert(kGraphHyperSeqOffset + 4 <= kHyperBytes && sizeof(AdamHyper) <= kGraphHyperSeqOffset && sizeof(SgdHyper) <= kGraphHyperSeqOffset,
                  "the sequence word sits behind the hyper-parameters");
    void* ack = (kind == 0 || kind == 1) ? graph_last_hyper_ack() : nullptr;
    if (ack) {
      const unsigned seq = 1u;
      memcpy((char*)host + kGraphHyperSeqOffset, &seq, sizeof(seq));
    }
    dfb_status st = graph_early_h2d(dev, host, bytes, ack, kGraphHyperSeqOffset);   // a branch off the graph's root, not a node in front of the kernel
    if (st != DFB_OK) return st;
  } else {
    DFB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, compute_stream()));
  }
  if (!graph_capturing()) {
    TableStage& st = g_stage;
    int s = (st.next + TableStage::kSlots - 1) % TableStage::kSlots;
    DFB_CUDA(cudaEventRecord(st.done[s], compute_stream()));
  }
  *dev_hyper = dev;
  *dev_table = (const TensorRef*)((char*)dev + kHyperBytes);
  return DFB_OK;
}

static AdamHyper make_adam_hyper(double lr, double beta1, double beta2, double eps, double weight_decay, int step_t,
                                 double grad_scale) {
  AdamHyper h;
  // every scalar is a Python double rounded to float32 when it reaches a scalar_* kernel
  h.lr = (float)lr;
  h.beta1 = (float)beta1;
  h.beta2 = (float)beta2;
  h.one_minus_beta1 = (float)(1.0 - beta1);
  h.one_minus_beta2 = (float)(1.0 - beta2);
  h.bias1 = (float)(1.0 - pow(beta1, (double)step_t));
  h.bias2 = (float)(1.0 - pow(beta2, (double)step_t));
  h.eps = (float)eps;
  h.weight_decay = (float)weight_decay;
  h.grad_scale = (float)grad_scale;
  h.use_wd = weight_decay > 0.0;
  h.use_scale = grad_scale != 1.0;
  return h;
}
static SgdHyper make_sgd_hyper(double lr, double momentum, double weight_decay, int nesterov, double grad_scale) {
  SgdHyper h;
  h.lr = (float)lr;
  h.momentum = (float)momentum;
  h.weight_decay = (float)weight_decay;
  h.grad_scale = (float)grad_scale;
  h.use_momentum = momentum > 0.0;
  h.nesterov = nesterov != 0;
  h.use_scale = grad_scale != 1.0;
  return h;
}

static dfb_status build_table(const char* name, float* const* params, const float* const* grads,
                              float* const* m, float* const* v, const size_t* sizes, int count,
                              bool need_m, bool need_v, std::vector<TensorRef>* tab,
                              unsigned long long* total_chunks) {
  DFB_REQUIRE(count >= 0, DFB_ERR_INVALID, "%s: negative tensor count", name);
  tab->clear();
  unsigned long long chunks = 0;
  for (int i = 0; i < count; ++i) {
    if (sizes[i] == 0) continue;
    DFB_REQUIRE(params[i] && grads[i], DFB_ERR_INVALID, "%s: null param/grad pointer for tensor %d", name, i);
    DFB_REQUIRE(!need_m || (m && m[i]), DFB_ERR_INVALID, "%s: null state pointer for tensor %d", name, i);
    DFB_REQUIRE(!need_v || (v && v[i]), DFB_ERR_INVALID, "%s: null state pointer for tensor %d", name, i);
    TensorRef t;
    t.p = params[i];
    t.g = grads[i];
    t.m = m ? m[i] : nullptr;
    t.v = v ? v[i] : nullptr;
    t.size = sizes[i];
    t.first_chunk = chunks;
    chunks += (sizes[i] + kChunk - 1) / kChunk;
    tab->push_back(t);
  }
  *total_chunks = chunks;
  return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

dfb_status dfb_multi_adam_step(float* const* params, const float* const* grads, float* const* exp_avg,
                               float* const* exp_avg_sq, const size_t* sizes, int count, double lr,
                               double beta1, double beta2, double eps, double weight_decay, int step_t,
                               double grad_scale) {
  DFB_INIT();
  DFB_REQUIRE(step_t >= 1, DFB_ERR_INVALID, "multi_adam_step: step_t starts at 1 (adam.py:26), got %d", step_t);
  std::vector<TensorRef> tab;
  unsigned long long chunks = 0;
  dfb_status st = build_table("multi_adam_step", params, grads, exp_avg, exp_avg_sq, sizes, count, true, true, &tab, &chunks);
  if (st != DFB_OK) return st;
  if (chunks == 0) return DFB_OK;
  static_assert(sizeof(AdamHyper) <= kHyperBytes && sizeof(SgdHyper) <= kHyperBytes, "hyper block too small");
  AdamHyper h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step_t, grad_scale);
  const TensorRef* dev = nullptr;
  const void* dev_h = nullptr;
  st = upload_table(tab, &h, sizeof(h), 0, &dev, &dev_h);
  if (st != DFB_OK) return st;
  unsigned grid = (unsigned)std::min<unsigned long long>(chunks, (unsigned long long)sm_count() * 8);
  launch_k(multi_adam_kernel, grid, 256, 0, compute_stream(), dev, (int)tab.size(), chunks, (const AdamHyper*)dev_h, peer_dev(),
           peer_pending_take());
  DFB_LAUNCH_CHECK("multi_adam_step");
  return DFB_OK;
}

dfb_status dfb_multi_sgd_step(float* const* params, const float* const* grads, float* const* velocity,
                              const size_t* sizes, int count, double lr, double momentum,
                              double weight_decay, int nesterov, double grad_scale) {
  DFB_INIT();
  std::vector<TensorRef> tab;
  unsigned long long chunks = 0;
  bool use_m = momentum > 0.0;
  dfb_status st = build_table("multi_sgd_step", params, grads, velocity, nullptr, sizes, count, use_m, false, &tab, &chunks);
  if (st != DFB_OK) return st;
  if (chunks == 0) return DFB_OK;
  SgdHyper h = make_sgd_hyper(lr, momentum, weight_decay, nesterov, grad_scale);
  const TensorRef* dev = nullptr;
  const void* dev_h = nullptr;
  st = upload_table(tab, &h, sizeof(h), 1, &dev, &dev_h);
  if (st != DFB_OK) return st;
  unsigned grid = (unsigned)std::min<unsigned long long>(chunks, (unsigned long long)sm_count() * 8);
  launch_k(multi_sgd_kernel, grid, 256, 0, compute_stream(), dev, (int)tab.size(), chunks, (const SgdHyper*)dev_h, peer_dev(),
           peer_pending_take());
  DFB_LAUNCH_CHECK("multi_sgd_step");
  return DFB_OK;
}

dfb_status dfb_multi_copy(const float* const* srcs, float* const* dsts, const size_t* sizes, int count) {
  DFB_INIT();
  std::vector<TensorRef> tab;
  unsigned long long chunks = 0;
  dfb_status st = build_table("multi_copy", dsts, srcs, nullptr, nullptr, sizes, count, false, false, &tab, &chunks);
  if (st != DFB_OK) return st;
  if (chunks == 0) return DFB_OK;
  const TensorRef* dev = nullptr;
  const void* dev_h = nullptr;
  char none[16] = {0};
  st = upload_table(tab, none, sizeof(none), 2, &dev, &dev_h);
  if (st != DFB_OK) return st;
  unsigned grid = (unsigned)std::min<unsigned long long>(chunks, (unsigned long long)sm_count() * 8);
  launch_k(multi_copy_kernel, grid, 256, 0, compute_stream(), dev, (int)tab.size(), chunks);
  DFB_LAUNCH_CHECK("multi_copy");
  return DFB_OK;
}

// Hyper-parameters of the index-th optimizer step captured in `graph_exec`, for its next replay.
dfb_status dfb_graph_set_adam(void* graph_exec, int index, double lr, double beta1, double beta2, double eps,
                              double weight_decay, int step_t, double grad_scale) {
  DFB_REQUIRE(step_t >= 1, DFB_ERR_INVALID, "graph_set_adam: step_t starts at 1, got %d", step_t);
  void* host = nullptr;
  dfb_status st = graph_hyper_slot(graph_exec, index, 0, &host);
  if (st != DFB_OK) return st;
  AdamHyper h = make_adam_hyper(lr, beta1, beta2, eps, weight_decay, step_t, grad_scale);
  memcpy(host, &h, sizeof(h));
  return DFB_OK;
}
dfb_status dfb_graph_set_sgd(void* graph_exec, int index, double lr, double momentum, double weight_decay,
                             int nesterov, double grad_scale) {
  void* host = nullptr;
  dfb_status st = graph_hyper_slot(graph_exec, index, 1, &host);
  if (st != DFB_OK) return st;
  SgdHyper h = make_sgd_hyper(lr, momentum, weight_decay, nesterov, grad_scale);
  memcpy(host, &h, sizeof(h));
  return DFB_OK;
}

}  // extern "C"
