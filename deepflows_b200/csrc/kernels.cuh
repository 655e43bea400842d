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

// gemm_tc.cu — TMA + tcgen05/TMEM path. Each returns DFB_OK and sets *handled = true when it ran
// the problem, leaves *handled = false when the shape is outside what the tensor-core kernels
// take (the dispatcher then uses the SIMT path), or returns an error status.
dfb_status tc_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int trans_b,
                   int lda, int ldb, int ldc, int accumulate, const float* bias, int mode, bool* handled);
dfb_status tc_conv_fprop(const float* x, const float* w, int w_layout, float* y, int N, int C, int H, int W, int K, int R,
                         int pad, int stride, int mode, float* workspace, size_t workspace_floats,
                         bool* handled);
dfb_status tc_conv_dgrad(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                         int R, int pad, int stride, int mode, float* workspace, size_t workspace_floats,
                         bool* handled);
dfb_status tc_conv_wgrad(const float* x, const float* dy, float* dw, int w_layout, int N, int C, int H, int W, int K,
                         int R, int pad, int stride, int mode, float* workspace, size_t workspace_floats,
                         bool* handled);
// first-layer weight gradient (C <= 4, C*R*R <= 32) as a column matrix + the tcgen05 1x1 wgrad (TF32 mode, large batches)
dfb_status tc_stem_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H, int W,
                         int K, int R, int pad, int stride, int mode, bool* handled);
size_t tc_conv_workspace_floats(int N, int C, int H, int W, int K, int R, int pad, int stride);

}  // namespace dfb
