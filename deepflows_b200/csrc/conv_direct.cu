// Direct convolution kernels for the first layer of a network: very few input channels (C <= 4: grey-scale or
// RGB images), any square kernel with C*R*R <= 80 taps, K <= 64 output channels.
//
// With C = 3 the implicit-GEMM reduction is 27 (or 75) long: far too short for a tensor-core tile, and the
// channels-last rows are 12 bytes, which no TMA descriptor can address. These layers are pure bandwidth
// (the stem of ResNet-18/CIFAR writes 33.5 MB and computes 0.45 GFLOP), so they get FFMA kernels
// whose job is to touch x, y / dy exactly once:
//   fprop : thread = one output pixel x 8 output channels; weights in shared memory; the 4 (K = 32) threads
//           of a pixel write one 128-byte line of the channels-last output.
//   wgrad : a CTA walks its share of the pixels in tiles of 64; per tile the dy rows and the gathered taps
//           (im2col of 64 pixels, never in global memory) are staged in shared memory and every thread keeps
//           a 4 (k) x 2 (tap) block of dW in registers; CTA partials are added by a second small kernel in a
//           fixed order (deterministic).
// Replaces, for these layers, the reference's pad + k*k setitems + compact + naive matmul
// (DeepFlows/nn/functional.py:249-344). Exact fp32 (FFMA), so they serve every precision mode.
#include "kernels.cuh"

#include <algorithm>

namespace dfb {
namespace direct {

constexpr int kMaxTaps = 80;   // C * R * R
constexpr int kMaxK = 64;
constexpr int kThreads = 256;

struct Geom {
  int N, C, H, W, K, R, pad, stride, OH, OW;
  int nchw;   // x layout
  int krsc;   // weight / dW layout
};

__device__ __forceinline__ size_t x_index(const Geom& g, int n, int c, int ih, int iw) {
  return g.nchw ? (((size_t)n * g.C + c) * g.H + ih) * g.W + iw : (((size_t)n * g.H + ih) * g.W + iw) * g.C + c;
}
// tap t = (c * R + r) * R + s  (the (K,C,R,R) order of one output channel's weights)
__device__ __forceinline__ size_t w_index(const Geom& g, int k, int t) {
  if (!g.krsc) return (size_t)k * g.C * g.R * g.R + t;
  const int rr = g.R * g.R;
  const int c = t / rr, rs = t - c * rr;
  return ((size_t)k * rr + rs) * g.C + c;
}

// ---- fprop --------------------------------------------------------------------------------------------
// shared: w_s[T][K] (tap-major so that the 16 output channels of a thread are four float4 reads)
// R is a template parameter for the common kernel sizes (0 = run-time loop bounds): the tap loops unroll and
// the address arithmetic leaves the inner loop, which is what this instruction-bound kernel needs.
template <int RT>
__global__ void __launch_bounds__(kThreads)
direct_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, Geom g) {
  pdl_sync();
  extern __shared__ float w_s[];
  const int R = RT ? RT : g.R;
  const int T = g.C * R * R, K = g.K;
  for (int i = threadIdx.x; i < T * K; i += blockDim.x) {
    const int t = i / K, k = i - t * K;
    w_s[i] = __ldg(w + w_index(g, k, t));
  }
  __syncthreads();
  const int kgroups = K / 16;                      // K % 16 == 0 (host checks)
  const size_t pixels = (size_t)g.N * g.OH * g.OW;
  const size_t items = pixels * kgroups;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t sc = g.nchw ? (size_t)g.H * g.W : 1, sh = g.nchw ? g.W : (size_t)g.W * g.C, sw = g.nchw ? 1 : g.C;
  const size_t sn = (size_t)g.C * g.H * g.W;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += stride) {
    const int kg = (int)(it % kgroups);
    const size_t pix = it / kgroups;
    const int ow = (int)(pix % g.OW);
    const size_t t2 = pix / g.OW;
    const int oh = (int)(t2 % g.OH), n = (int)(t2 / g.OH);
    const int ih0 = oh * g.stride - g.pad, iw0 = ow * g.stride - g.pad;
    const float* xn = x + (size_t)n * sn;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    const float* wk = w_s + kg * 16;
    for (int c = 0; c < g.C; ++c) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ih = ih0 + r;
        const bool okh = (unsigned)ih < (unsigned)g.H;
#pragma unroll
        for (int s2 = 0; s2 < R; ++s2) {
          const int iw = iw0 + s2;
          float xv = 0.f;
          if (okh && (unsigned)iw < (unsigned)g.W) xv = __ldg(xn + c * sc + ih * sh + iw * sw);
          const float* wt = wk + ((c * R + r) * R + s2) * K;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + q * 4);
            acc[q * 4 + 0] = fmaf(xv, wv.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(xv, wv.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(xv, wv.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(xv, wv.w, acc[q * 4 + 3]);
          }
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(y + pix * K + kg * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) st_stream(dst + q, make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]));
  }
}

// ---- wgrad --------------------------------------------------------------------------------------------
constexpr int kPixTile = 64;
constexpr int kMaxBlocksPerThread = 4;   // (K/4) * ceil(T/2) <= 16 * 40 = 640 register blocks over 256 threads... see host

// partial[cta][K*T] in the memory order of dW (so the final sum over CTAs is a plain column sum)
__global__ void __launch_bounds__(kThreads)
direct_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial, Geom g,
                    size_t pix_per_cta) {
  pdl_sync();
  extern __shared__ float sm[];
  const int T = g.C * g.R * g.R, K = g.K;
  const int Tp = (T + 1) & ~1;                    // taps padded to an even count
  float* dy_s = sm;                               // [kPixTile][K]
  float* col_s = sm + kPixTile * K;               // [kPixTile][Tp]
  __shared__ signed char tap_c[kMaxTaps + 2], tap_r[kMaxTaps + 2], tap_s[kMaxTaps + 2];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int rr = g.R * g.R;
    const int c = t / rr, rs = t - c * rr;
    tap_c[t] = (signed char)c;
    tap_r[t] = (signed char)(rs / g.R);
    tap_s[t] = (signed char)(rs - (rs / g.R) * g.R);
  }
  const int kq_n = K / 4, tq_n = Tp / 2;
  const int nblocks = kq_n * tq_n;                // register blocks of 4 (k) x 2 (taps)
  float acc[kMaxBlocksPerThread][8];
#pragma unroll
  for (int b = 0; b < kMaxBlocksPerThread; ++b)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[b][j] = 0.f;

  const size_t pixels = (size_t)g.N * g.OH * g.OW;
  const size_t p_begin = (size_t)blockIdx.x * pix_per_cta;
  const size_t p_end = p_begin + pix_per_cta < pixels ? p_begin + pix_per_cta : pixels;
  for (size_t p0 = p_begin; p0 < p_end; p0 += kPixTile) {
    const int np = (int)(p_end - p0 < (size_t)kPixTile ? p_end - p0 : (size_t)kPixTile);
    __syncthreads();
    // dy rows of the tile: contiguous in channels-last memory
    for (int i = threadIdx.x; i < kPixTile * K / 4; i += blockDim.x) {
      const int p = i / (K / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < np) v = ld_stream(reinterpret_cast<const float4*>(dy + (p0 + p) * K) + (i - p * (K / 4)));
      reinterpret_cast<float4*>(dy_s)[i] = v;
    }
    // gathered taps of the tile (zero outside the image and for the padding tap): thread = (pixel, every
    // 4th tap); the pixel is decoded once and the loads of a thread are independent (all in flight together)
    {
      const int p = threadIdx.x % kPixTile, grp = threadIdx.x / kPixTile;  // 256 threads = 64 pixels x 4 tap groups
      const size_t pix = p0 + p;
      const int ow = (int)(pix % g.OW);
      const size_t t2 = pix / g.OW;
      const int oh = (int)(t2 % g.OH), n = (int)(t2 / g.OH);
      const int ih0 = oh * g.stride - g.pad, iw0 = ow * g.stride - g.pad;
      const float* xn = x + (size_t)n * g.C * g.H * g.W;
#pragma unroll 4
      for (int t = grp; t < Tp; t += kThreads / kPixTile) {
        float v = 0.f;
        if (p < np && t < T) {
          const int c = tap_c[t], ih = ih0 + tap_r[t], iw = iw0 + tap_s[t];
          if ((unsigned)ih < (unsigned)g.H && (unsigned)iw < (unsigned)g.W)
            v = __ldg(xn + (g.nchw ? ((size_t)c * g.H + ih) * g.W + iw : ((size_t)ih * g.W + iw) * g.C + c));
        }
        col_s[p * Tp + t] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < kMaxBlocksPerThread; ++b) {
      const int blk = threadIdx.x + b * kThreads;
      if (blk < nblocks) {
        const int kq = blk % kq_n, tq = blk / kq_n;
        for (int p = 0; p < np; ++p) {
          const float4 d = *reinterpret_cast<const float4*>(dy_s + p * K + kq * 4);
          const float2 c2 = *reinterpret_cast<const float2*>(col_s + p * Tp + tq * 2);
          acc[b][0] = fmaf(d.x, c2.x, acc[b][0]); acc[b][1] = fmaf(d.x, c2.y, acc[b][1]);
          acc[b][2] = fmaf(d.y, c2.x, acc[b][2]); acc[b][3] = fmaf(d.y, c2.y, acc[b][3]);
          acc[b][4] = fmaf(d.z, c2.x, acc[b][4]); acc[b][5] = fmaf(d.z, c2.y, acc[b][5]);
          acc[b][6] = fmaf(d.w, c2.x, acc[b][6]); acc[b][7] = fmaf(d.w, c2.y, acc[b][7]);
        }
      }
    }
  }
  float* out = partial + (size_t)blockIdx.x * K * T;
#pragma unroll
  for (int b = 0; b < kMaxBlocksPerThread; ++b) {
    const int blk = threadIdx.x + b * kThreads;
    if (blk < nblocks) {
      const int kq = blk % kq_n, tq = blk / kq_n;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int t = tq * 2 + j;
          if (t < T) out[w_index(g, kq * 4 + i, t)] = acc[b][i * 2 + j];  // already in dW's memory order
        }
    }
  }
}

static bool eligible(int C, int K, int R, int stride) {
  const int T = C * R * R;
  if (C > 4 || T > kMaxTaps || K > kMaxK || (K % 16) != 0 || stride < 1) return false;
  const int nblocks = (K / 4) * ((T + 1) / 2);
  return nblocks <= kMaxBlocksPerThread * kThreads;
}

static Geom make_geom(int N, int C, int H, int W, int K, int R, int pad, int stride, int x_layout, int w_layout) {
  Geom g;
  g.N = N; g.C = C; g.H = H; g.W = W; g.K = K; g.R = R; g.pad = pad; g.stride = stride;
  g.OH = (H + 2 * pad - R) / stride + 1;
  g.OW = (W + 2 * pad - R) / stride + 1;
  g.nchw = x_layout == DFB_LAYOUT_NCHW;
  g.krsc = w_layout == DFB_WLAYOUT_KRSC;
  return g;
}

}  // namespace direct

dfb_status direct_conv_fprop(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H, int W,
                             int K, int R, int pad, int stride, bool* handled) {
  using namespace direct;
  *handled = false;
  if (!eligible(C, K, R, stride) || (reinterpret_cast<uintptr_t>(y) & 15)) return DFB_OK;
  if (H + 2 * pad < R || W + 2 * pad < R) return DFB_OK;
  Geom g = make_geom(N, C, H, W, K, R, pad, stride, x_layout, w_layout);
  const size_t items = (size_t)N * g.OH * g.OW * (K / 16);
  if (items == 0) return DFB_OK;
  const size_t smem = (size_t)C * R * R * K * sizeof(float);
  *handled = true;
  const unsigned grid = bw_grid(items, kThreads, 8);
  if (R == 3) launch_k(direct_fprop_kernel<3>, grid, kThreads, smem, compute_stream(), x, w, y, g);
  else if (R == 5) launch_k(direct_fprop_kernel<5>, grid, kThreads, smem, compute_stream(), x, w, y, g);
  else launch_k(direct_fprop_kernel<0>, grid, kThreads, smem, compute_stream(), x, w, y, g);
  DFB_LAUNCH_CHECK("conv2d_fprop(direct)");
  return DFB_OK;
}

dfb_status direct_conv_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H, int W,
                             int K, int R, int pad, int stride, bool* handled) {
  using namespace direct;
  *handled = false;
  if (!eligible(C, K, R, stride) || (reinterpret_cast<uintptr_t>(dy) & 15)) return DFB_OK;
  if (H + 2 * pad < R || W + 2 * pad < R) return DFB_OK;
  Geom g = make_geom(N, C, H, W, K, R, pad, stride, x_layout, w_layout);
  const size_t pixels = (size_t)N * g.OH * g.OW;
  if (pixels == 0) return DFB_OK;
  const int T = C * R * R, Tp = (T + 1) & ~1;
  size_t ctas = std::min<size_t>((size_t)sm_count() * 2, (pixels + kPixTile - 1) / kPixTile);
  size_t pix_per_cta = (pixels + ctas - 1) / ctas;
  pix_per_cta = (pix_per_cta + kPixTile - 1) / kPixTile * kPixTile;
  ctas = (pixels + pix_per_cta - 1) / pix_per_cta;
  float* partial = nullptr;
  dfb_status st = dfb_malloc(ctas * K * T, &partial);
  if (st != DFB_OK) return st;
  *handled = true;
  const size_t smem = (size_t)kPixTile * (K + Tp) * sizeof(float);
  launch_k(direct_wgrad_kernel, (unsigned)ctas, kThreads, smem, compute_stream(), x, dy, partial, g, pix_per_cta);
  DFB_LAUNCH_CHECK("conv2d_wgrad(direct)");
  st = dfb_colsum(partial, dw, ctas, K * T);  // fixed-order sum of the CTA partials
  dfb_free(partial);  // stream-ordered
  return st;
}

}  // namespace dfb
