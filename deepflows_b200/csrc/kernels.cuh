// Internal entry points shared between the dispatchers (gemm.cu) and the kernel files.
#pragma once
#include "common.cuh"

namespace dfb {

// gemm_simt.cu — exact fp32 FFMA path
dfb_status simt_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_a,
                     int trans_b, int lda, int ldb, int ldc, int accumulate, const float* bias);
dfb_status simt_conv_fprop(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H,
                           int W, int K, int R, int pad, int stride);
dfb_status simt_conv_dgrad(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                           int R, int pad, int stride, int dgrad_mode);
dfb_status simt_conv_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H,
                           int W, int K, int R, int pad, int stride);

// conv_direct.cu — first-layer convolutions (C <= 4 input channels): bandwidth-bound FFMA kernels. *handled as below.
dfb_status direct_conv_fprop(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H, int W,
                             int K, int R, int pad, int stride, bool* handled);
dfb_status direct_conv_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H, int W,
                             int K, int R, int pad, int stride, bool* handled);

// Optional work fused into the epilogue of a convolution (gemm_tc.cu: RowEpi; gemm.cu runs the same work as separate
// kernels when the tensor-core path does not take the problem):
//   addend    : out += addend (tensor of the output's shape)
//   stats     : FUSE_STATS -> stats_out[2][n_out] = per-channel mean / biased variance of the output
//               FUSE_BNBWD -> stats_out[3][n_out] = sum(out), sum(out * x_hat_0), sum(out * x_hat_1) with
//                             x_hat_i = (bn_x[i] - bn_mean[i]) * bn_invstd[i]  (n_sets of them, 1 or 2)
enum { FUSE_NONE = 0, FUSE_STATS = 1, FUSE_BNBWD = 2 };
struct ConvFuse {
  const float* addend;
  int kind, n_sets;
  float* stats_out;
  const float* bn_x[2];
  const float* bn_mean[2];
  const float* bn_invstd[2];
  // relu != 0 (with FUSE_BNBWD): the gradient belongs to relu(bn_0(x_0) [+ bn_1(x_1)] [+ relu_res]); it is masked with
  // the recomputed pre-activation (>= 0 passes) before it is written and summed
  int relu;
  const float* bn_gamma[2];
  const float* bn_beta[2];
  const float* relu_res;
  int lazy;   // the statistics may stay in a statistic slot for dfb_bn_fwd_apply / dfb_bn_bwd_apply (see below)
};

// Statistic slots (runtime.cu): where a convolution's epilogue hands per-channel sums to the BatchNorm kernel that
// consumes them WITHOUT a reduction kernel in between. A slot is double[3][kStatSlotChannels] + an arrival counter in device
// memory, all zeros while free. The producer's CTAs add their (fp32) partial sums with fp64 atomics - the sum of a few
// hundred fp32 values in fp64 is exact or off by one fp64 ulp, so the fp32 results do not depend on the arrival order
// in practice; the consumer kernel (bn_apply_fused_kernel / bn_bwd_apply_kernel) reads the sums in its prologue, CTA 0
// publishes them as floats where the eager path would have put them, and the last CTA to have read clears the slot.
// Keyed by the statistics buffer the two calls share; only calls that announce themselves as lazy (ConvFuse::lazy,
// dfb_conv2d_fprop_stats_lazy / dfb_conv2d_dgrad_fused_lazy) use them. DFB_STAT_SLOTS=0 switches them off.
constexpr int kStatSlots = 96, kStatSlotChannels = 512;
constexpr size_t kStatSlotBytes = (size_t)3 * kStatSlotChannels * sizeof(double) + 64;   // + counter, padded
double* stat_slot_acquire(const float* key, int consumers, int channels);   // null: none (the caller reduces eagerly)
double* stat_slot_take(const float* key, int* last);                        // consumer side; null: the buffer holds floats
void stat_slot_drop(const float* key);                                      // an eager producer is about to fill `key`
__device__ __forceinline__ unsigned* stat_slot_counter(double* slot) { return reinterpret_cast<unsigned*>(slot + 3 * kStatSlotChannels); }

// gemm_tc.cu — TMA + tcgen05/TMEM path. Each returns DFB_OK and sets *handled = true when it ran
// the problem, leaves *handled = false when the shape is outside what the tensor-core kernels
// take (the dispatcher then uses the SIMT path), or returns an error status.
dfb_status tc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int trans_b,
                   int lda, int ldb, int ldc, int accumulate, const float* bias, int mode, bool* handled);
dfb_status tc_conv_fprop(const float* x, const float* w, int w_layout, float* y, int N, int C, int H, int W, int K, int R,
                         int pad, int stride, int mode, float* workspace, size_t workspace_floats,
                         bool* handled, const ConvFuse* fuse = nullptr);
dfb_status tc_conv_dgrad(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                         int R, int pad, int stride, int mode, float* workspace, size_t workspace_floats,
                         bool* handled, const ConvFuse* fuse = nullptr);
dfb_status tc_conv_wgrad(const float* x, const float* dy, float* dw, int w_layout, int N, int C, int H, int W, int K,
                         int R, int pad, int stride, int mode, float* workspace, size_t workspace_floats,
                         bool* handled);
// first-layer weight gradient (C <= 4, C*R*R <= 32) as a column matrix + the tcgen05 1x1 wgrad (TF32 mode, large batches)
dfb_status tc_stem_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H, int W,
                         int K, int R, int pad, int stride, int mode, bool* handled);
size_t tc_conv_workspace_floats(int N, int C, int H, int W, int K, int R, int pad, int stride);
// the first layer's column matrix as a separate step (see gemm_tc.cu)
dfb_status tc_stem_cols(const float* x, int x_layout, float* col, int N, int C, int H, int W, int R, int pad, int stride, int w_layout);
dfb_status tc_stem_pad_weights(const float* w, float* wp, int K, int cols);
dfb_status tc_wgrad_cols(const float* col, const float* dy, float* dw, int w_layout, int N, int OH, int OW, int K, int cols, int mode,
                         bool* handled);

}  // namespace dfb
