// Input-pipeline kernels: what the reference's training scripts do on the host, in numpy, to every
// batch before `Tensor(batch, device=cuda)` - and what bounds the step once the step itself takes a millisecond:
//   * augment_batch (test/ResNet_CIFAR10_cuda.py:129-148): reflect-pad by `pad`, per-sample random crop, per-sample
//     horizontal flip, optional per-sample erased rectangle, clip;
//   * one-hot + label smoothing (test/ResNet_CIFAR10_cuda.py:181-183): onehot * (1 - eps) + eps / classes.
// The random draws stay on the host (numpy's generator, so a seeded run sees the reference's numbers); they travel as
// a small float table with the batch, and the per-pixel work is one pass over the batch here. Results are bit-identical
// to the numpy code: the kernels only move values, compare, and (one-hot) do one multiply and one add in the
// reference's order.
#include "common.cuh"

namespace dfb {
namespace {

constexpr int kT = 256;
constexpr int kAugFields = DFB_AUGMENT_FIELDS;

// index into the unpadded image of position p of the reflect-padded one (numpy 'reflect': the edge is not repeated)
__device__ __forceinline__ int reflect(int p, int n) {
  if (p < 0) p = -p;
  if (p >= n) p = 2 * (n - 1) - p;
  return p;
}

// x, y: (N, C, H, W). table: N rows of {crop_y, crop_x, flip, erase_y, erase_x, erase_h, erase_w, unused} as floats.
__global__ void __launch_bounds__(kT) augment_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                     const float* __restrict__ table, int N, int C, int H, int W, int pad,
                                                     float lo, float hi, int clip) {
  pdl_sync();
  const size_t total = (size_t)N * C * H * W;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int w = (int)(i % W);
    size_t t = i / W;
    const int h = (int)(t % H);
    t /= H;
    const int n = (int)(t / C);
    const float* row = table + (size_t)n * kAugFields;
    const int cy = __float2int_rn(row[0]), cx = __float2int_rn(row[1]), flip = __float2int_rn(row[2]);
    const int ey = __float2int_rn(row[3]), ex = __float2int_rn(row[4]), eh = __float2int_rn(row[5]), ew = __float2int_rn(row[6]);
    float v = 0.f;
    if (!(h >= ey && h < ey + eh && w >= ex && w < ex + ew)) {   // the erased rectangle is in output coordinates
      const int wc = flip ? W - 1 - w : w;                       // the flip mirrors the cropped image
      const int sh = reflect(cy + h - pad, H), sw = reflect(cx + wc - pad, W);
      v = __ldg(x + (t * H + sh) * W + sw);                       // t == n * C + c
    }
    if (clip && v == v) v = fminf(fmaxf(v, lo), hi);  // numpy's clip propagates NaN
    y[i] = v;
  }
}

// y[n][j] = (j == label[n] ? 1 : 0) * on + off, the multiply and the add rounded separately like numpy does
__global__ void __launch_bounds__(kT) onehot_kernel(const float* __restrict__ labels, float* __restrict__ y, size_t n, int classes,
                                                    float on, float off) {
  pdl_sync();
  const size_t total = n * (size_t)classes;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int j = (int)(i % classes);
    const float hot = __float2int_rn(labels[i / classes]) == j ? 1.f : 0.f;
    y[i] = __fadd_rn(__fmul_rn(hot, on), off);
  }
}

// Dropout mask on the device (opt-in; the reference draws it on the host with numpy, dropout.py:27-29, which costs
// milliseconds per step for a 256 x 2048 activation - more than the rest of the step). Philox4x32-10 (Salmon et al., SC'11;
// the counter-based generator cuRAND and numpy's Philox implement), key = (seed, 0xCAFEF00D), counter = (element / 4, step,
// 0, 0): mask[i] = u32 * 2^-32 < keep_prob ? 1 : 0. oracle/numpy_ops.py::philox_dropout_mask restates it; bit-exact.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], const uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__global__ void __launch_bounds__(kT) dropout_mask_kernel(float* __restrict__ mask, size_t n, float keep_prob, const float* __restrict__ state) {
  pdl_sync();
  const uint32_t seed = (uint32_t)__float2uint_rn(state[0]), step = (uint32_t)__float2uint_rn(state[1]);
  const size_t groups = (n + 3) / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
    uint32_t c[4] = {(uint32_t)g, step, (uint32_t)(g >> 32), 0u};
    uint32_t k[2] = {seed, 0xCAFEF00Du};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      philox_round(c, k);
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const size_t i = g * 4 + j;
      if (i < n) mask[i] = (float)c[j] * 2.3283064365386963e-10f < keep_prob ? 1.f : 0.f;
    }
  }
}

}  // namespace
}  // namespace dfb

using namespace dfb;

extern "C" {

dfb_status dfb_augment_batch(const float* x, float* y, const float* table, int N, int C, int H, int W, int pad, int clip,
                             float clip_lo, float clip_hi) {
  DFB_INIT();
  DFB_REQUIRE(x && y && table, DFB_ERR_INVALID, "augment_batch: null pointer");
  DFB_REQUIRE(x != y, DFB_ERR_INVALID, "augment_batch: the output must not alias the input");
  DFB_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, DFB_ERR_INVALID, "augment_batch: bad shape (%d,%d,%d,%d)", N, C, H, W);
  // numpy's reflect padding needs pad <= size - 1 (one reflection)
  DFB_REQUIRE(pad >= 0 && pad < H && pad < W, DFB_ERR_INVALID, "augment_batch: pad %d must be smaller than the image (%d x %d)", pad, H, W);
  if (N == 0) return DFB_OK;
  const size_t total = (size_t)N * C * H * W;
  launch_k(augment_kernel, bw_grid(total, kT), kT, 0, compute_stream(), x, y, table, N, C, H, W, pad, clip_lo, clip_hi, clip);
  DFB_LAUNCH_CHECK("augment_batch");
  return DFB_OK;
}

dfb_status dfb_onehot_smooth(const float* labels, float* y, size_t n, int classes, float on_value, float off_value) {
  DFB_INIT();
  DFB_REQUIRE(labels && y, DFB_ERR_INVALID, "onehot_smooth: null pointer");
  DFB_REQUIRE(classes > 0, DFB_ERR_INVALID, "onehot_smooth: classes must be positive");
  if (n == 0) return DFB_OK;
  launch_k(onehot_kernel, bw_grid(n * (size_t)classes, kT), kT, 0, compute_stream(), labels, y, n, classes, on_value, off_value);
  DFB_LAUNCH_CHECK("onehot_smooth");
  return DFB_OK;
}

dfb_status dfb_dropout_mask(float* mask, size_t n, float keep_prob, const float* state) {
  DFB_INIT();
  DFB_REQUIRE(mask && state, DFB_ERR_INVALID, "dropout_mask: null pointer");
  if (n == 0) return DFB_OK;
  launch_k(dropout_mask_kernel, bw_grid((n + 3) / 4, kT), kT, 0, compute_stream(), mask, n, keep_prob, state);
  DFB_LAUNCH_CHECK("dropout_mask");
  return DFB_OK;
}

}  // extern "C"
