// Bandwidth-bound L1 kernels on channels-last data: BatchNorm forward/backward, ReLU, row-vector
// add, column sums, max/avg pooling, fused softmax cross-entropy.
//
// The reference composes each of these from dozens of out-of-place L0 kernels with the
// broadcast operands materialised and the broadcast gradients reduced on the host
// (DeepFlows/nn/modules/batchnorm.py:30-55, DeepFlows/nn/functional.py:104-115,347-374,
// DeepFlows/tensor.py:462-483). Here each op is one or two passes over HBM with 128-bit accesses.
//
// Data layout: activations are (rows = N*H*W, C) row-major, i.e. NHWC. A "column group" is 4
// adjacent channels (one float4) when C % 4 == 0, otherwise one channel.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace dfb {

constexpr int kT = 256;

// -------------------------------------------------------------------------------------------------
// Column statistics. A CTA owns a contiguous chunk of rows; thread t handles column group
// (t % G) on row lanes (t / G), so a warp reads whole 128-byte lines.
// -------------------------------------------------------------------------------------------------
template <int V>
struct Vec;
template <>
struct Vec<4> {
  using T = float4;
  static __device__ __forceinline__ void get(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void put(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<1> {
  static __device__ __forceinline__ void get(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void put(float* p, const float (&v)[1]) { *p = v[0]; }
};

struct ColPlan {
  int V;            // 4 or 1
  int G;            // column groups = C / V
  int lanes;        // row lanes per CTA = max(1, kT / G)
  int gpass;        // column-group passes per CTA = ceil(G / kT)
  unsigned ctas;    // CTAs (row chunks)
  size_t rows_per_cta;
};
static ColPlan plan_cols(size_t rows, int C, const void* p0, const void* p1 = nullptr) {
  ColPlan p;
  bool al = ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15) == 0;
  p.V = (C % 4 == 0 && al) ? 4 : 1;
  p.G = C / p.V;
  p.lanes = std::max(1, kT / p.G);
  p.gpass = (p.G + kT - 1) / kT;
  size_t want = (rows + (size_t)p.lanes * 8 - 1) / ((size_t)p.lanes * 8);  // >= 8 rows per lane
  size_t cap = (size_t)sm_count() * 4;
  p.ctas = (unsigned)std::max<size_t>(1, std::min(want, cap));
  p.rows_per_cta = (rows + p.ctas - 1) / p.ctas;
  return p;
}

// Partial moments of one CTA's row chunk, per channel: count, mean, M2 (sum of squared
// deviations). Threads accumulate shifted sums (shift = first sample seen, which keeps
// sum(d^2) - sum(d)^2/n well conditioned), lanes are merged with Chan's formula.
template <int V>
__global__ void __launch_bounds__(kT)
col_moments_kernel(const float* __restrict__ x, size_t rows, int C, ColPlan p,
                   float* __restrict__ part_cnt, float* __restrict__ part_mean, float* __restrict__ part_m2) {
  extern __shared__ float sm[];  // [lanes][3][G*V] when lanes > 1
  const int G = p.G;
  size_t r0 = (size_t)blockIdx.x * p.rows_per_cta;
  size_t r1 = r0 + p.rows_per_cta < rows ? r0 + p.rows_per_cta : rows;
  for (int gp = 0; gp < p.gpass; ++gp) {
    int g = gp * kT + (threadIdx.x % (G < kT ? G : kT));
    int lane = G < kT ? threadIdx.x / G : 0;
    bool active = g < G && lane < p.lanes;
    float n = 0.f, shift[V], s1[V], s2[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { shift[v] = 0.f; s1[v] = 0.f; s2[v] = 0.f; }
    if (active) {
      for (size_t r = r0 + lane; r < r1; r += p.lanes) {
        float xv[V];
        Vec<V>::get(x + r * C + (size_t)g * V, xv);
        if (n == 0.f) {
#pragma unroll
          for (int v = 0; v < V; ++v) shift[v] = xv[v];
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float d = xv[v] - shift[v];
          s1[v] += d;
          s2[v] = fmaf(d, d, s2[v]);
        }
        n += 1.f;
      }
    }
    float mean[V], m2[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float dm = n > 0.f ? s1[v] / n : 0.f;
      mean[v] = shift[v] + dm;
      m2[v] = n > 0.f ? fmaxf(s2[v] - s1[v] * dm, 0.f) : 0.f;
    }
    if (p.lanes > 1) {
      // merge lanes through shared memory (lane 0 of each column group does the merge)
      int CV = G * V;
      __syncthreads();
      if (active) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          sm[(lane * 3 + 0) * CV + g * V + v] = n;
          sm[(lane * 3 + 1) * CV + g * V + v] = mean[v];
          sm[(lane * 3 + 2) * CV + g * V + v] = m2[v];
        }
      }
      __syncthreads();
      if (active && lane == 0) {
        const float n_own = n;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float na = n_own, ma = mean[v], qa = m2[v];
          for (int l = 1; l < p.lanes; ++l) {
            float nb = sm[(l * 3 + 0) * CV + g * V + v];
            if (nb == 0.f) continue;
            float mb = sm[(l * 3 + 1) * CV + g * V + v], qb = sm[(l * 3 + 2) * CV + g * V + v];
            float nt = na + nb, d = mb - ma;
            ma += d * (nb / nt);
            qa += qb + d * d * (na * nb / nt);
            na = nt;
          }
          n = na; mean[v] = ma; m2[v] = qa;
        }
      }
    }
    if (active && lane == 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        size_t o = (size_t)blockIdx.x * C + (size_t)g * V + v;
        part_cnt[o] = n;
        part_mean[o] = mean[v];
        part_m2[o] = m2[v];
      }
    }
  }
}

// Merge the per-CTA partials (one warp per channel: lanes merge strided subsets with Chan's formula, then a
// shuffle tree; the order is fixed => deterministic), produce mean / invstd, update the running statistics
// with the *biased* variance like batchnorm.py:44-46.
__device__ __forceinline__ void chan_merge(float& na, float& ma, float& qa, float nb, float mb, float qb) {
  if (nb == 0.f) return;
  float nt = na + nb, d = mb - ma;
  ma += d * (nb / nt);
  qa += qb + d * d * (na * nb / nt);
  na = nt;
}
__global__ void __launch_bounds__(kT)
bn_finalize_kernel(const float* __restrict__ part_cnt, const float* __restrict__ part_mean,
                   const float* __restrict__ part_m2, int parts, int C, float eps, float momentum,
                   float* __restrict__ save_mean, float* __restrict__ save_invstd,
                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (c >= C) return;
  float na = 0.f, ma = 0.f, qa = 0.f;
  for (int i = lane; i < parts; i += 32)
    chan_merge(na, ma, qa, part_cnt[(size_t)i * C + c], part_mean[(size_t)i * C + c], part_m2[(size_t)i * C + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float nb = __shfl_down_sync(0xffffffffu, na, o), mb = __shfl_down_sync(0xffffffffu, ma, o),
          qb = __shfl_down_sync(0xffffffffu, qa, o);
    chan_merge(na, ma, qa, nb, mb, qb);
  }
  if (lane != 0) return;
  float mean = ma;
  float var = na > 0.f ? qa / na : 0.f;
  save_mean[c] = mean;
  save_invstd[c] = 1.0f / sqrtf(var + eps);
  if (running_mean) running_mean[c] = running_mean[c] * (1.0f - momentum) + mean * momentum;
  if (running_var) running_var[c] = running_var[c] * (1.0f - momentum) + var * momentum;
}

// y = (x - mean) * (invstd * gamma) + beta, optionally followed by max(.,0)
template <int V, bool RELU>
__global__ void __launch_bounds__(kT)
bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, size_t rows, int C,
                const float* __restrict__ mean, const float* __restrict__ invstd,
                const float* __restrict__ gamma, const float* __restrict__ beta) {
  extern __shared__ float sm[];  // mean[C], scale[C], shift[C]
  float* s_mean = sm;
  float* s_scale = sm + C;
  float* s_shift = sm + 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_mean[c] = mean[c];
    s_scale[c] = invstd[c] * (gamma ? gamma[c] : 1.0f);
    s_shift[c] = beta ? beta[c] : 0.0f;
  }
  __syncthreads();
  const int G = C / V;
  size_t total = rows * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    float xv[V], yv[V];
    Vec<V>::get(x + i * V, xv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      int c = g * V + v;
      float t = fmaf(xv[v] - s_mean[c], s_scale[c], s_shift[c]);
      yv[v] = RELU ? fmaxf(t, 0.f) : t;
    }
    Vec<V>::put(y + i * V, yv);
  }
}

// BatchNorm backward, pass 1: per-CTA partial sums of dy and dy * x_hat per channel.
template <int V>
__global__ void __launch_bounds__(kT)
bn_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, size_t rows, int C,
                     ColPlan p, const float* __restrict__ mean, const float* __restrict__ invstd,
                     float* __restrict__ part_dbeta, float* __restrict__ part_dgamma) {
  extern __shared__ float sm[];  // [lanes][2][C]
  const int G = p.G;
  size_t r0 = (size_t)blockIdx.x * p.rows_per_cta;
  size_t r1 = r0 + p.rows_per_cta < rows ? r0 + p.rows_per_cta : rows;
  for (int gp = 0; gp < p.gpass; ++gp) {
    int g = gp * kT + (threadIdx.x % (G < kT ? G : kT));
    int lane = G < kT ? threadIdx.x / G : 0;
    bool active = g < G && lane < p.lanes;
    float sb[V], sg[V], mu[V], is[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      sb[v] = 0.f; sg[v] = 0.f;
      mu[v] = active ? mean[g * V + v] : 0.f;
      is[v] = active ? invstd[g * V + v] : 0.f;
    }
    if (active) {
      for (size_t r = r0 + lane; r < r1; r += p.lanes) {
        float xv[V], dv[V];
        Vec<V>::get(x + r * C + (size_t)g * V, xv);
        Vec<V>::get(dy + r * C + (size_t)g * V, dv);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          sb[v] += dv[v];
          sg[v] = fmaf(dv[v], (xv[v] - mu[v]) * is[v], sg[v]);
        }
      }
    }
    if (p.lanes > 1) {
      __syncthreads();
      if (active) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          sm[(lane * 2 + 0) * C + g * V + v] = sb[v];
          sm[(lane * 2 + 1) * C + g * V + v] = sg[v];
        }
      }
      __syncthreads();
      if (active && lane == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          for (int l = 1; l < p.lanes; ++l) {
            sb[v] += sm[(l * 2 + 0) * C + g * V + v];
            sg[v] += sm[(l * 2 + 1) * C + g * V + v];
          }
        }
      }
    }
    if (active && lane == 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        part_dbeta[(size_t)blockIdx.x * C + g * V + v] = sb[v];
        part_dgamma[(size_t)blockIdx.x * C + g * V + v] = sg[v];
      }
    }
  }
}

__global__ void __launch_bounds__(kT)
colsum_finalize_kernel(const float* __restrict__ part_a, const float* __restrict__ part_b, int parts,
                       int C, float* __restrict__ out_a, float* __restrict__ out_b) {
  // one warp per channel; lanes take strided partials, shuffle tree at the end (fixed order)
  int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int i = lane; i < parts; i += 32) {
    a += part_a[(size_t)i * C + c];
    if (part_b) b += part_b[(size_t)i * C + c];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) {
    if (out_a) out_a[c] = a;
    if (out_b) out_b[c] = b;
  }
}

// pass 2: dx = gamma * invstd * (dy - dbeta/n - x_hat * dgamma/n)
template <int V>
__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                    size_t rows, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ dbeta,
                    const float* __restrict__ dgamma) {
  extern __shared__ float sm[];  // mean, invstd, k1 = gamma*invstd, mb = dbeta/n, mg = dgamma/n
  float* s_mean = sm;
  float* s_is = sm + C;
  float* s_k1 = sm + 2 * C;
  float* s_mb = sm + 3 * C;
  float* s_mg = sm + 4 * C;
  float inv_n = 1.0f / (float)rows;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_mean[c] = mean[c];
    s_is[c] = invstd[c];
    s_k1[c] = invstd[c] * (gamma ? gamma[c] : 1.0f);
    s_mb[c] = dbeta[c] * inv_n;
    s_mg[c] = dgamma[c] * inv_n;
  }
  __syncthreads();
  const int G = C / V;
  size_t total = rows * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    float xv[V], dv[V], ov[V];
    Vec<V>::get(x + i * V, xv);
    Vec<V>::get(dy + i * V, dv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      int c = g * V + v;
      float xh = (xv[v] - s_mean[c]) * s_is[c];
      ov[v] = s_k1[c] * (dv[v] - s_mb[c] - xh * s_mg[c]);
    }
    Vec<V>::put(dx + i * V, ov);
  }
}

// plain column sum partials (bias gradients)
template <int V>
__global__ void __launch_bounds__(kT)
colsum_partial_kernel(const float* __restrict__ x, size_t rows, int C, ColPlan p, float* __restrict__ part) {
  extern __shared__ float sm[];  // [lanes][C]
  const int G = p.G;
  size_t r0 = (size_t)blockIdx.x * p.rows_per_cta;
  size_t r1 = r0 + p.rows_per_cta < rows ? r0 + p.rows_per_cta : rows;
  for (int gp = 0; gp < p.gpass; ++gp) {
    int g = gp * kT + (threadIdx.x % (G < kT ? G : kT));
    int lane = G < kT ? threadIdx.x / G : 0;
    bool active = g < G && lane < p.lanes;
    float s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) s[v] = 0.f;
    if (active) {
      for (size_t r = r0 + lane; r < r1; r += p.lanes) {
        float xv[V];
        Vec<V>::get(x + r * C + (size_t)g * V, xv);
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] += xv[v];
      }
    }
    if (p.lanes > 1) {
      __syncthreads();
      if (active) {
#pragma unroll
        for (int v = 0; v < V; ++v) sm[lane * C + g * V + v] = s[v];
      }
      __syncthreads();
      if (active && lane == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v)
          for (int l = 1; l < p.lanes; ++l) s[v] += sm[l * C + g * V + v];
      }
    }
    if (active && lane == 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) part[(size_t)blockIdx.x * C + g * V + v] = s[v];
    }
  }
}

// -------------------------------------------------------------------------------------------------
// elementwise
// -------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kT)
add_rowvec_kernel(const float* __restrict__ x, const float* __restrict__ vec, float* __restrict__ y,
                  size_t rows, int C) {
  const int G = C / V;
  size_t total = rows * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    float xv[V], bv[V];
    Vec<V>::get(x + i * V, xv);
    Vec<V>::get(vec + (size_t)g * V, bv);
#pragma unroll
    for (int v = 0; v < V; ++v) xv[v] += bv[v];
    Vec<V>::put(y + i * V, xv);
  }
}

__global__ void __launch_bounds__(kT)
relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n4 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) |
                reinterpret_cast<uintptr_t>(dx)) & 15) == 0 ? n / 4 : 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 a = ld_stream(reinterpret_cast<const float4*>(x) + i);
    float4 d = ld_stream(reinterpret_cast<const float4*>(dy) + i);
    float4 o;
    // maximum.grad_fn: (max(x,0) == x) * dy  <=>  x >= 0 ? dy : 0   (tensor.py:872-877)
    o.x = a.x >= 0.f ? d.x : 0.f; o.y = a.y >= 0.f ? d.y : 0.f;
    o.z = a.z >= 0.f ? d.z : 0.f; o.w = a.w >= 0.f ? d.w : 0.f;
    st_stream(reinterpret_cast<float4*>(dx) + i, o);
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dx[i] = x[i] >= 0.f ? dy[i] : 0.f;
}

// -------------------------------------------------------------------------------------------------
// pooling (window k, stride k, no padding) on NHWC
// -------------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(kT)
maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int32_t* __restrict__ idx, int N,
                   int H, int W, int C, int k, int OH, int OW) {
  const int G = C / V;
  size_t total = (size_t)N * OH * OW * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    size_t t = i / G;
    int ow = (int)(t % OW);
    t /= OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    float best[V];
    int bi[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { best[v] = -INFINITY; bi[v] = 0; }
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        float xv[V];
        Vec<V>::get(x + (((size_t)n * H + oh * k + r) * W + ow * k + s) * C + (size_t)g * V, xv);
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (xv[v] > best[v] || (r == 0 && s == 0)) { best[v] = xv[v]; bi[v] = r * k + s; }
      }
    Vec<V>::put(y + i * V, best);
    if (idx) {
#pragma unroll
      for (int v = 0; v < V; ++v) idx[i * V + v] = bi[v];
    }
  }
}

// MODE 0: every tied maximum receives dy (reference, tensor.py:779-791); MODE 1: arg-max routed
template <int V, int MODE>
__global__ void __launch_bounds__(kT)
maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const int32_t* __restrict__ idx,
                   const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int k,
                   int OH, int OW) {
  const int G = C / V;
  size_t total = (size_t)N * H * W * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    size_t t = i / G;
    int w = (int)(t % W);
    t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    int oh = h / k, ow = w / k;
    float o[V];
#pragma unroll
    for (int v = 0; v < V; ++v) o[v] = 0.f;
    if (oh < OH && ow < OW) {
      size_t oo = (((size_t)n * OH + oh) * OW + ow) * C + (size_t)g * V;
      float dv[V];
      Vec<V>::get(dy + oo, dv);
      if (MODE == 0) {
        float xv[V], yv[V];
        Vec<V>::get(x + i * V, xv);
        Vec<V>::get(y + oo, yv);
#pragma unroll
        for (int v = 0; v < V; ++v) o[v] = xv[v] == yv[v] ? dv[v] : 0.f;
      } else {
        int pos = (h - oh * k) * k + (w - ow * k);
#pragma unroll
        for (int v = 0; v < V; ++v) o[v] = idx[oo + v] == pos ? dv[v] : 0.f;
      }
    }
    Vec<V>::put(dx + i * V, o);
  }
}

template <int V>
__global__ void __launch_bounds__(kT)
avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int k,
                   int OH, int OW) {
  const int G = C / V;
  size_t total = (size_t)N * OH * OW * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float inv = 1.0f / (float)(k * k);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    size_t t = i / G;
    int ow = (int)(t % OW);
    t /= OW;
    int oh = (int)(t % OH);
    int n = (int)(t / OH);
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        float xv[V];
        Vec<V>::get(x + (((size_t)n * H + oh * k + r) * W + ow * k + s) * C + (size_t)g * V, xv);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += xv[v];
      }
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] *= inv;
    Vec<V>::put(y + i * V, acc);
  }
}
template <int V>
__global__ void __launch_bounds__(kT)
avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int H, int W, int C, int k,
                   int OH, int OW) {
  const int G = C / V;
  size_t total = (size_t)N * H * W * G;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  float inv = 1.0f / (float)(k * k);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int g = (int)(i % G);
    size_t t = i / G;
    int w = (int)(t % W);
    t /= W;
    int h = (int)(t % H);
    int n = (int)(t / H);
    int oh = h / k, ow = w / k;
    float o[V];
#pragma unroll
    for (int v = 0; v < V; ++v) o[v] = 0.f;
    if (oh < OH && ow < OW) {
      Vec<V>::get(dy + (((size_t)n * OH + oh) * OW + ow) * C + (size_t)g * V, o);
#pragma unroll
      for (int v = 0; v < V; ++v) o[v] *= inv;
    }
    Vec<V>::put(dx + i * V, o);
  }
}

// -------------------------------------------------------------------------------------------------
// softmax cross-entropy with dense targets. One CTA: the batch dimension of every config is
// <= a few thousand rows, and a single CTA keeps the sum order fixed.
// -------------------------------------------------------------------------------------------------
constexpr int kCeThreads = 1024;
__global__ void __launch_bounds__(kCeThreads)
softmax_ce_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                      float* __restrict__ loss, size_t rows, int cols, float scale) {
  __shared__ float warp_part[kCeThreads / 32];
  float acc = 0.f;
  for (size_t r = threadIdx.x; r < rows; r += blockDim.x) {
    const float* xr = logits + r * cols;
    const float* tr = target + r * cols;
    float m = -INFINITY;
    for (int j = 0; j < cols; ++j) m = fmaxf(m, xr[j]);
    float se = 0.f;
    for (int j = 0; j < cols; ++j) se += expf(xr[j] - m);
    float lse = logf(se);
    float row = 0.f;
    for (int j = 0; j < cols; ++j) row += -((xr[j] - m) - lse) * tr[j];
    acc += row;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kCeThreads / 32 ? warp_part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) loss[0] = v * scale;
  }
}
__global__ void __launch_bounds__(kT)
softmax_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                      const float* __restrict__ upstream, float* __restrict__ dlogits, size_t rows,
                      int cols, float scale) {
  float k = scale * (upstream ? upstream[0] : 1.0f);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const float* xr = logits + r * cols;
    const float* tr = target + r * cols;
    float m = -INFINITY, ts = 0.f;
    for (int j = 0; j < cols; ++j) { m = fmaxf(m, xr[j]); ts += tr[j]; }
    float se = 0.f;
    for (int j = 0; j < cols; ++j) se += expf(xr[j] - m);
    float inv = 1.0f / se;
    for (int j = 0; j < cols; ++j)
      dlogits[r * cols + j] = k * (expf(xr[j] - m) * inv * ts - tr[j]);
  }
}

static unsigned ew_grid(size_t items) { return bw_grid(items, kT, 8); }

}  // namespace dfb

using namespace dfb;

extern "C" {

dfb_status dfb_add_rowvec(const float* x, const float* v, float* y, size_t rows, int cols) {
  DFB_INIT();
  DFB_REQUIRE(x && v && y, DFB_ERR_INVALID, "add_rowvec: null pointer");
  DFB_REQUIRE(cols > 0, DFB_ERR_INVALID, "add_rowvec: cols must be positive");
  if (rows == 0) return DFB_OK;
  bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (cols % 4 == 0 && al)
    add_rowvec_kernel<4><<<ew_grid(rows * (cols / 4)), kT, 0, compute_stream()>>>(x, v, y, rows, cols);
  else
    add_rowvec_kernel<1><<<ew_grid(rows * cols), kT, 0, compute_stream()>>>(x, v, y, rows, cols);
  DFB_LAUNCH_CHECK("add_rowvec");
  return DFB_OK;
}

dfb_status dfb_colsum(const float* x, float* out, size_t rows, int cols) {
  DFB_INIT();
  DFB_REQUIRE(x && out, DFB_ERR_INVALID, "colsum: null pointer");
  DFB_REQUIRE(cols > 0, DFB_ERR_INVALID, "colsum: cols must be positive");
  if (rows == 0) return dfb_fill(out, 0.f, cols);
  ColPlan p = plan_cols(rows, cols, x);
  float* part = nullptr;
  dfb_status st = dfb_malloc((size_t)p.ctas * cols, &part);
  if (st != DFB_OK) return st;
  size_t smem = p.lanes > 1 ? (size_t)p.lanes * cols * sizeof(float) : 0;
  cudaStream_t s = compute_stream();
  if (p.V == 4) colsum_partial_kernel<4><<<p.ctas, kT, smem, s>>>(x, rows, cols, p, part);
  else colsum_partial_kernel<1><<<p.ctas, kT, smem, s>>>(x, rows, cols, p, part);
  DFB_LAUNCH_CHECK("colsum");
  colsum_finalize_kernel<<<cdiv((size_t)cols * 32, kT), kT, 0, s>>>(part, nullptr, (int)p.ctas, cols, out, nullptr);
  DFB_LAUNCH_CHECK("colsum");
  dfb_free(part);
  return DFB_OK;
}

static dfb_status bn_stats(const float* x, size_t rows, int C, float eps, float momentum, float* save_mean,
                           float* save_invstd, float* running_mean, float* running_var) {
  ColPlan p = plan_cols(rows, C, x);
  float* part = nullptr;
  dfb_status st = dfb_malloc((size_t)p.ctas * C * 3, &part);
  if (st != DFB_OK) return st;
  float* pc = part;
  float* pm = part + (size_t)p.ctas * C;
  float* pq = part + (size_t)p.ctas * C * 2;
  size_t smem = p.lanes > 1 ? (size_t)p.lanes * 3 * C * sizeof(float) : 0;
  cudaStream_t s = compute_stream();
  DFB_REQUIRE(smem <= 48 * 1024, DFB_ERR_INVALID, "batchnorm: %d channels need %zu B of shared memory", C, smem);
  if (p.V == 4) col_moments_kernel<4><<<p.ctas, kT, smem, s>>>(x, rows, C, p, pc, pm, pq);
  else col_moments_kernel<1><<<p.ctas, kT, smem, s>>>(x, rows, C, p, pc, pm, pq);
  DFB_LAUNCH_CHECK("bn_stats");
  bn_finalize_kernel<<<cdiv((size_t)C * 32, kT), kT, 0, s>>>(pc, pm, pq, (int)p.ctas, C, eps, momentum, save_mean, save_invstd,
                                               running_mean, running_var);
  DFB_LAUNCH_CHECK("bn_finalize");
  dfb_free(part);
  return DFB_OK;
}

static dfb_status bn_apply(const float* x, float* y, size_t rows, int C, const float* mean, const float* invstd,
                           const float* gamma, const float* beta, bool relu) {
  bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  size_t smem = (size_t)3 * C * sizeof(float);
  DFB_REQUIRE(smem <= 48 * 1024, DFB_ERR_INVALID, "batchnorm: too many channels (%d)", C);
  cudaStream_t s = compute_stream();
  if (C % 4 == 0 && al) {
    unsigned grid = ew_grid(rows * (C / 4));
    if (relu) bn_apply_kernel<4, true><<<grid, kT, smem, s>>>(x, y, rows, C, mean, invstd, gamma, beta);
    else bn_apply_kernel<4, false><<<grid, kT, smem, s>>>(x, y, rows, C, mean, invstd, gamma, beta);
  } else {
    unsigned grid = ew_grid(rows * C);
    if (relu) bn_apply_kernel<1, true><<<grid, kT, smem, s>>>(x, y, rows, C, mean, invstd, gamma, beta);
    else bn_apply_kernel<1, false><<<grid, kT, smem, s>>>(x, y, rows, C, mean, invstd, gamma, beta);
  }
  DFB_LAUNCH_CHECK("bn_apply");
  return DFB_OK;
}

dfb_status dfb_bn_fwd_train(const float* x, const float* gamma, const float* beta, float* y, float* save_mean,
                            float* save_invstd, float* running_mean, float* running_var, float momentum,
                            float eps, size_t rows, int C) {
  DFB_INIT();
  DFB_REQUIRE(x && y && save_mean && save_invstd, DFB_ERR_INVALID, "bn_fwd_train: null pointer");
  DFB_REQUIRE(rows > 0 && C > 0, DFB_ERR_INVALID, "bn_fwd_train: empty input");
  dfb_status st = bn_stats(x, rows, C, eps, momentum, save_mean, save_invstd, running_mean, running_var);
  if (st != DFB_OK) return st;
  return bn_apply(x, y, rows, C, save_mean, save_invstd, gamma, beta, false);
}

// eval: x_hat = (x - running_mean) / (running_var + eps)**0.5 (batchnorm.py:49-50)
__global__ void bn_eval_prep_kernel(const float* __restrict__ rv, float eps, int C, float* __restrict__ invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) invstd[c] = 1.0f / sqrtf(rv[c] + eps);
}
dfb_status dfb_bn_fwd_eval(const float* x, const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float* y, float eps, size_t rows, int C) {
  DFB_INIT();
  DFB_REQUIRE(x && y && running_mean && running_var, DFB_ERR_INVALID, "bn_fwd_eval: null pointer");
  if (rows == 0) return DFB_OK;
  float* invstd = nullptr;
  dfb_status st = dfb_malloc(C, &invstd);
  if (st != DFB_OK) return st;
  bn_eval_prep_kernel<<<cdiv(C, kT), kT, 0, compute_stream()>>>(running_var, eps, C, invstd);
  DFB_LAUNCH_CHECK("bn_eval_prep");
  st = bn_apply(x, y, rows, C, running_mean, invstd, gamma, beta, false);
  dfb_free(invstd);
  return st;
}

dfb_status dfb_bn_bwd(const float* x, const float* dy, const float* gamma, const float* save_mean,
                      const float* save_invstd, float* dx, float* dgamma, float* dbeta, size_t rows, int C) {
  DFB_INIT();
  DFB_REQUIRE(x && dy && save_mean && save_invstd, DFB_ERR_INVALID, "bn_bwd: null pointer");
  DFB_REQUIRE(rows > 0 && C > 0, DFB_ERR_INVALID, "bn_bwd: empty input");
  ColPlan p = plan_cols(rows, C, x, dy);
  float* scratch = nullptr;
  dfb_status st = dfb_malloc((size_t)p.ctas * C * 2 + 2 * (size_t)C, &scratch);
  if (st != DFB_OK) return st;
  float* pb = scratch;
  float* pg = scratch + (size_t)p.ctas * C;
  float* db = dbeta ? dbeta : scratch + (size_t)p.ctas * C * 2;
  float* dg = dgamma ? dgamma : scratch + (size_t)p.ctas * C * 2 + C;
  size_t smem = p.lanes > 1 ? (size_t)p.lanes * 2 * C * sizeof(float) : 0;
  cudaStream_t s = compute_stream();
  if (p.V == 4) bn_bwd_reduce_kernel<4><<<p.ctas, kT, smem, s>>>(x, dy, rows, C, p, save_mean, save_invstd, pb, pg);
  else bn_bwd_reduce_kernel<1><<<p.ctas, kT, smem, s>>>(x, dy, rows, C, p, save_mean, save_invstd, pb, pg);
  DFB_LAUNCH_CHECK("bn_bwd_reduce");
  colsum_finalize_kernel<<<cdiv((size_t)C * 32, kT), kT, 0, s>>>(pb, pg, (int)p.ctas, C, db, dg);
  DFB_LAUNCH_CHECK("bn_bwd_finalize");
  if (dx) {
    bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    size_t sm2 = (size_t)5 * C * sizeof(float);
    if (C % 4 == 0 && al)
      bn_bwd_apply_kernel<4><<<ew_grid(rows * (C / 4)), kT, sm2, s>>>(x, dy, dx, rows, C, save_mean, save_invstd, gamma, db, dg);
    else
      bn_bwd_apply_kernel<1><<<ew_grid(rows * C), kT, sm2, s>>>(x, dy, dx, rows, C, save_mean, save_invstd, gamma, db, dg);
    DFB_LAUNCH_CHECK("bn_bwd_apply");
  }
  dfb_free(scratch);
  return DFB_OK;
}

dfb_status dfb_relu_fwd(const float* x, float* y, size_t n) { return dfb_scalar_maximum(x, 0.0f, y, n); }

dfb_status dfb_relu_bwd(const float* x, const float* dy, float* dx, size_t n) {
  DFB_INIT();
  DFB_REQUIRE(x && dy && dx, DFB_ERR_INVALID, "relu_bwd: null pointer");
  if (n == 0) return DFB_OK;
  relu_bwd_kernel<<<ew_grid(n / 4 + 1), kT, 0, compute_stream()>>>(x, dy, dx, n);
  DFB_LAUNCH_CHECK("relu_bwd");
  return DFB_OK;
}

static dfb_status pool_geom(const char* name, int N, int H, int W, int C, int k, int* OH, int* OW) {
  DFB_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && k > 0 && H >= k && W >= k, DFB_ERR_INVALID,
              "%s: bad geometry N=%d H=%d W=%d C=%d k=%d", name, N, H, W, C, k);
  *OH = (H - k) / k + 1;
  *OW = (W - k) / k + 1;
  return DFB_OK;
}
#define POOL_DISPATCH(KERNEL, items_scalar, ...)                                               \
  do {                                                                                         \
    if (vec) KERNEL<4><<<ew_grid((items_scalar) / 4), kT, 0, compute_stream()>>>(__VA_ARGS__); \
    else KERNEL<1><<<ew_grid(items_scalar), kT, 0, compute_stream()>>>(__VA_ARGS__);           \
  } while (0)

static bool all_aligned(const void* a, const void* b = nullptr, const void* c = nullptr, const void* d = nullptr) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0;
}

dfb_status dfb_maxpool2d_fwd(const float* x, float* y, int32_t* idx, int N, int H, int W, int C, int k) {
  DFB_INIT();
  DFB_REQUIRE(x && y, DFB_ERR_INVALID, "maxpool2d_fwd: null pointer");
  int OH, OW;
  dfb_status st = pool_geom("maxpool2d_fwd", N, H, W, C, k, &OH, &OW);
  if (st != DFB_OK) return st;
  bool vec = C % 4 == 0 && all_aligned(x, y, idx);
  POOL_DISPATCH(maxpool_fwd_kernel, (size_t)N * OH * OW * C, x, y, idx, N, H, W, C, k, OH, OW);
  DFB_LAUNCH_CHECK("maxpool2d_fwd");
  return DFB_OK;
}
dfb_status dfb_maxpool2d_bwd(const float* x, const float* y, const float* dy, float* dx, int N, int H, int W,
                             int C, int k) {
  DFB_INIT();
  DFB_REQUIRE(x && y && dy && dx, DFB_ERR_INVALID, "maxpool2d_bwd: null pointer");
  int OH, OW;
  dfb_status st = pool_geom("maxpool2d_bwd", N, H, W, C, k, &OH, &OW);
  if (st != DFB_OK) return st;
  bool vec = C % 4 == 0 && all_aligned(x, y, dy, dx);
  size_t items = (size_t)N * H * W * C;
  if (vec) maxpool_bwd_kernel<4, 0><<<ew_grid(items / 4), kT, 0, compute_stream()>>>(x, y, nullptr, dy, dx, N, H, W, C, k, OH, OW);
  else maxpool_bwd_kernel<1, 0><<<ew_grid(items), kT, 0, compute_stream()>>>(x, y, nullptr, dy, dx, N, H, W, C, k, OH, OW);
  DFB_LAUNCH_CHECK("maxpool2d_bwd");
  return DFB_OK;
}
dfb_status dfb_maxpool2d_bwd_idx(const int32_t* idx, const float* dy, float* dx, int N, int H, int W, int C, int k) {
  DFB_INIT();
  DFB_REQUIRE(idx && dy && dx, DFB_ERR_INVALID, "maxpool2d_bwd_idx: null pointer");
  int OH, OW;
  dfb_status st = pool_geom("maxpool2d_bwd_idx", N, H, W, C, k, &OH, &OW);
  if (st != DFB_OK) return st;
  bool vec = C % 4 == 0 && all_aligned(idx, dy, dx);
  size_t items = (size_t)N * H * W * C;
  if (vec) maxpool_bwd_kernel<4, 1><<<ew_grid(items / 4), kT, 0, compute_stream()>>>(nullptr, nullptr, idx, dy, dx, N, H, W, C, k, OH, OW);
  else maxpool_bwd_kernel<1, 1><<<ew_grid(items), kT, 0, compute_stream()>>>(nullptr, nullptr, idx, dy, dx, N, H, W, C, k, OH, OW);
  DFB_LAUNCH_CHECK("maxpool2d_bwd_idx");
  return DFB_OK;
}
dfb_status dfb_avgpool2d_fwd(const float* x, float* y, int N, int H, int W, int C, int k) {
  DFB_INIT();
  DFB_REQUIRE(x && y, DFB_ERR_INVALID, "avgpool2d_fwd: null pointer");
  int OH, OW;
  dfb_status st = pool_geom("avgpool2d_fwd", N, H, W, C, k, &OH, &OW);
  if (st != DFB_OK) return st;
  bool vec = C % 4 == 0 && all_aligned(x, y);
  POOL_DISPATCH(avgpool_fwd_kernel, (size_t)N * OH * OW * C, x, y, N, H, W, C, k, OH, OW);
  DFB_LAUNCH_CHECK("avgpool2d_fwd");
  return DFB_OK;
}
dfb_status dfb_avgpool2d_bwd(const float* dy, float* dx, int N, int H, int W, int C, int k) {
  DFB_INIT();
  DFB_REQUIRE(dy && dx, DFB_ERR_INVALID, "avgpool2d_bwd: null pointer");
  int OH, OW;
  dfb_status st = pool_geom("avgpool2d_bwd", N, H, W, C, k, &OH, &OW);
  if (st != DFB_OK) return st;
  bool vec = C % 4 == 0 && all_aligned(dy, dx);
  POOL_DISPATCH(avgpool_bwd_kernel, (size_t)N * H * W * C, dy, dx, N, H, W, C, k, OH, OW);
  DFB_LAUNCH_CHECK("avgpool2d_bwd");
  return DFB_OK;
}

dfb_status dfb_softmax_ce_fwd(const float* logits, const float* target, float* loss, size_t rows, int cols,
                              float scale) {
  DFB_INIT();
  DFB_REQUIRE(logits && target && loss, DFB_ERR_INVALID, "softmax_ce_fwd: null pointer");
  DFB_REQUIRE(cols > 0, DFB_ERR_INVALID, "softmax_ce_fwd: cols must be positive");
  softmax_ce_fwd_kernel<<<1, kCeThreads, 0, compute_stream()>>>(logits, target, loss, rows, cols, scale);
  DFB_LAUNCH_CHECK("softmax_ce_fwd");
  return DFB_OK;
}
dfb_status dfb_softmax_ce_bwd(const float* logits, const float* target, const float* upstream, float* dlogits,
                              size_t rows, int cols, float scale) {
  DFB_INIT();
  DFB_REQUIRE(logits && target && dlogits, DFB_ERR_INVALID, "softmax_ce_bwd: null pointer");
  if (rows == 0) return DFB_OK;
  softmax_ce_bwd_kernel<<<bw_grid(rows, kT), kT, 0, compute_stream()>>>(logits, target, upstream, dlogits, rows, cols, scale);
  DFB_LAUNCH_CHECK("softmax_ce_bwd");
  return DFB_OK;
}

}  // extern "C"
