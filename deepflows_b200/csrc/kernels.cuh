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
};

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
