// Dispatchers for the dense contractions: matmul / gemm / conv2d fprop, dgrad, wgrad.
//
// DFB_MODE_TF32 runs the TMA + tcgen05 kernels (gemm_tc.cu) with one kind::tf32 MMA per k-step; DFB_MODE_FP32 runs the
// same kernels in their fp32-accurate form (operands split into TF32 high and low parts in shared memory, three MMAs per
// k-step into the fp32 accumulator; DFB_FP32_TC=0 keeps it on the FFMA kernels). Shapes the tensor-core path does not
// take, DFB_MODE_BF16 (no separate bf16 operand format exists: it is served like SIMT, never less accurate than asked
// for) and DFB_MODE_SIMT run the exact-fp32 FFMA kernels (gemm_simt.cu).
#include "kernels.cuh"

using namespace dfb;

static bool want_tc(int mode) { return mode == DFB_MODE_TF32 || mode == DFB_MODE_FP32; }   // gemm_tc.cu: select_mode

static dfb_status check_mode(const char* name, int mode) {
  DFB_REQUIRE(mode >= DFB_MODE_FP32 && mode <= DFB_MODE_SIMT, DFB_ERR_INVALID, "%s: unknown mode %d", name, mode);
  return DFB_OK;
}

extern "C" {

dfb_status dfb_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_a, int trans_b,
                    int lda, int ldb, int ldc, int accumulate, const float* bias, int mode) {
  DFB_INIT();
  DFB_REQUIRE(A && B && C, DFB_ERR_INVALID, "gemm: null pointer");
  DFB_REQUIRE(M >= 0 && N >= 0 && K >= 0, DFB_ERR_INVALID, "gemm: negative dimension");
  dfb_status st = check_mode("gemm", mode);
  if (st != DFB_OK) return st;
  DFB_REQUIRE(lda >= (trans_a ? M : K) && ldb >= (trans_b ? K : N) && ldc >= N, DFB_ERR_INVALID,
              "gemm: leading dimension too small (lda=%d ldb=%d ldc=%d for M=%d N=%d K=%d ta=%d tb=%d)", lda,
              ldb, ldc, M, N, K, trans_a, trans_b);
  if (M == 0 || N == 0) return DFB_OK;
  if (want_tc(mode)) {
    bool handled = false;
    st = tc_gemm(A, B, C, M, N, K, trans_a, trans_b, lda, ldb, ldc, accumulate, bias, mode, &handled);
    if (st != DFB_OK || handled) return st;
  }
  return simt_gemm(A, B, C, M, N, K, trans_a, trans_b, lda, ldb, ldc, accumulate, bias);
}

// matmul: ndarray_backend_cuda.cu:443-466 — out[M,P] = a[M,N] . b[N,P]
dfb_status dfb_matmul(const float* a, const float* b, float* out, uint32_t M, uint32_t N, uint32_t P, int mode) {
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "Matmul: out array cannot be null");
  DFB_REQUIRE(M < (1u << 31) && N < (1u << 31) && P < (1u << 31), DFB_ERR_INVALID, "Matmul: dimension too large");
  return dfb_gemm(a, b, out, (int)M, (int)P, (int)N, 0, 0, (int)N, (int)P, (int)P, 0, nullptr, mode);
}

dfb_status dfb_conv2d_workspace_floats(int N, int C, int H, int W, int K, int R, int pad, int stride,
                                       size_t* n_floats) {
  DFB_REQUIRE(n_floats != nullptr, DFB_ERR_INVALID, "conv2d_workspace_floats: null output");
  *n_floats = tc_conv_workspace_floats(N, C, H, W, K, R, pad, stride);
  return DFB_OK;
}

dfb_status dfb_conv2d_fprop(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H,
                            int W, int K, int R, int pad, int stride, int mode, float* workspace,
                            size_t workspace_floats) {
  DFB_INIT();
  DFB_REQUIRE(x && w && y, DFB_ERR_INVALID, "conv2d_fprop: null pointer");
  DFB_REQUIRE(x_layout == DFB_LAYOUT_NCHW || x_layout == DFB_LAYOUT_NHWC, DFB_ERR_INVALID, "conv2d_fprop: bad layout");
  dfb_status st = check_mode("conv2d_fprop", mode);
  if (st != DFB_OK) return st;
  if (mode != DFB_MODE_SIMT) {  // image-like inputs (C <= 4): exact-fp32 direct kernel in every precision mode
    bool handled = false;
    st = direct_conv_fprop(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride, &handled);
    if (st != DFB_OK || handled) return st;
  }
  if (want_tc(mode) && x_layout == DFB_LAYOUT_NHWC) {
    bool handled = false;
    st = tc_conv_fprop(x, w, w_layout, y, N, C, H, W, K, R, pad, stride, mode, workspace, workspace_floats, &handled);
    if (st != DFB_OK || handled) return st;
  }
  return simt_conv_fprop(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride);
}

dfb_status dfb_conv2d_dgrad(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                            int R, int pad, int stride, int mode, int dgrad_mode, float* workspace,
                            size_t workspace_floats) {
  DFB_INIT();
  DFB_REQUIRE(dy && w && dx, DFB_ERR_INVALID, "conv2d_dgrad: null pointer");
  DFB_REQUIRE(dgrad_mode == DFB_DGRAD_REFERENCE || dgrad_mode == DFB_DGRAD_EXACT, DFB_ERR_INVALID,
              "conv2d_dgrad: bad dgrad_mode %d", dgrad_mode);
  dfb_status st = check_mode("conv2d_dgrad", mode);
  if (st != DFB_OK) return st;
  if (want_tc(mode) && dgrad_mode == DFB_DGRAD_EXACT) {
    bool handled = false;
    st = tc_conv_dgrad(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, mode, workspace, workspace_floats, &handled);
    if (st != DFB_OK || handled) return st;
  }
  return simt_conv_dgrad(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, dgrad_mode);
}

dfb_status dfb_conv2d_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H,
                            int W, int K, int R, int pad, int stride, int mode, float* workspace,
                            size_t workspace_floats) {
  DFB_INIT();
  DFB_REQUIRE(x && dy && dw, DFB_ERR_INVALID, "conv2d_wgrad: null pointer");
  DFB_REQUIRE(x_layout == DFB_LAYOUT_NCHW || x_layout == DFB_LAYOUT_NHWC, DFB_ERR_INVALID, "conv2d_wgrad: bad layout");
  dfb_status st = check_mode("conv2d_wgrad", mode);
  if (st != DFB_OK) return st;
  if (want_tc(mode)) {  // first layer at training batch sizes: column matrix + tcgen05 (gemm_tc.cu: tc_stem_wgrad)
    bool handled = false;
    st = tc_stem_wgrad(x, x_layout, dy, dw, w_layout, N, C, H, W, K, R, pad, stride, mode, &handled);
    if (st != DFB_OK || handled) return st;
  }
  if (mode != DFB_MODE_SIMT) {
    bool handled = false;
    st = direct_conv_wgrad(x, x_layout, dy, dw, w_layout, N, C, H, W, K, R, pad, stride, &handled);
    if (st != DFB_OK || handled) return st;
  }
  if (want_tc(mode) && x_layout == DFB_LAYOUT_NHWC) {
    bool handled = false;
    st = tc_conv_wgrad(x, dy, dw, w_layout, N, C, H, W, K, R, pad, stride, mode, workspace, workspace_floats, &handled);
    if (st != DFB_OK || handled) return st;
  }
  return simt_conv_wgrad(x, x_layout, dy, dw, w_layout, N, C, H, W, K, R, pad, stride);
}

// ---- convolutions with fused epilogue work ---------------------------------------------------------------------------
// Same contractions as above; the tensor-core kernels do the extra work in their epilogue (gemm_tc.cu: RowEpi), every
// other path (exact-fp32 FFMA kernels, the first-layer kernels) runs it as separate passes with the same results.

// y = conv(x, w) and mean_var[2][K] = per-channel mean / biased variance of y over (N, OH, OW): what the BatchNorm that
// follows needs (DeepFlows/nn/modules/batchnorm.py:33-42), without a pass of its own over y
static dfb_status fprop_stats_impl(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H,
                                   int W, int K, int R, int pad, int stride, int mode, float* mean_var, int lazy) {
  DFB_INIT();
  DFB_REQUIRE(x && w && y && mean_var, DFB_ERR_INVALID, "conv2d_fprop_stats: null pointer");
  DFB_REQUIRE(x_layout == DFB_LAYOUT_NCHW || x_layout == DFB_LAYOUT_NHWC, DFB_ERR_INVALID, "conv2d_fprop_stats: bad layout");
  dfb_status st = check_mode("conv2d_fprop_stats", mode);
  if (st != DFB_OK) return st;
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - R) / stride + 1;
  DFB_REQUIRE(OH > 0 && OW > 0 && N > 0, DFB_ERR_INVALID, "conv2d_fprop_stats: empty output");
  stat_slot_drop(mean_var);
  bool handled = false;
  if (mode != DFB_MODE_SIMT) {
    st = direct_conv_fprop(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride, &handled);
    if (st != DFB_OK) return st;
  }
  if (!handled && want_tc(mode) && x_layout == DFB_LAYOUT_NHWC) {
    ConvFuse f{};
    f.kind = FUSE_STATS;
    f.stats_out = mean_var;
    f.lazy = lazy;
    st = tc_conv_fprop(x, w, w_layout, y, N, C, H, W, K, R, pad, stride, mode, nullptr, 0, &handled, &f);
    if (st != DFB_OK || handled) return st;
  }
  if (!handled) {
    st = simt_conv_fprop(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride);
    if (st != DFB_OK) return st;
  }
  return dfb_colstats_mean_var(y, (size_t)N * OH * OW, K, mean_var);
}
dfb_status dfb_conv2d_fprop_stats(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H,
                                  int W, int K, int R, int pad, int stride, int mode, float* mean_var) {
  return fprop_stats_impl(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride, mode, mean_var, 0);
}
dfb_status dfb_conv2d_fprop_stats_lazy(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H,
                                       int W, int K, int R, int pad, int stride, int mode, float* mean_var) {
  return fprop_stats_impl(x, x_layout, w, w_layout, y, N, C, H, W, K, R, pad, stride, mode, mean_var, 1);
}

// dx = dgrad(dy, w) [+ addend], and for n_bn (0..2) BatchNorms whose OUTPUT gradient dx is: sums[0][C] = sum(dx),
// sums[1 + i][C] = sum(dx * x_hat_i) with x_hat_i = (bn_x_i - bn_mean_i) * bn_invstd_i - the two reductions of the
// BatchNorm backward (dbeta, dgamma), so that it only needs its elementwise pass (dfb_bn_bwd_apply)
static dfb_status dgrad_fused_impl(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                                   int R, int pad, int stride, int mode, int dgrad_mode, const float* addend, int n_bn,
                                   const float* bn_x0, const float* bn_mean0, const float* bn_invstd0, const float* bn_x1,
                                   const float* bn_mean1, const float* bn_invstd1, float* sums, int relu, const float* gamma0,
                                   const float* beta0, const float* gamma1, const float* beta1, const float* relu_res, int lazy) {
  DFB_INIT();
  DFB_REQUIRE(dy && w && dx, DFB_ERR_INVALID, "conv2d_dgrad_fused: null pointer");
  DFB_REQUIRE(n_bn >= 0 && n_bn <= 2, DFB_ERR_INVALID, "conv2d_dgrad_fused: n_bn must be 0, 1 or 2");
  DFB_REQUIRE(n_bn == 0 || (sums && bn_x0 && bn_mean0 && bn_invstd0), DFB_ERR_INVALID, "conv2d_dgrad_fused: BatchNorm 0 incomplete");
  DFB_REQUIRE(n_bn < 2 || (bn_x1 && bn_mean1 && bn_invstd1), DFB_ERR_INVALID, "conv2d_dgrad_fused: BatchNorm 1 incomplete");
  DFB_REQUIRE(!relu || n_bn >= 1, DFB_ERR_INVALID, "conv2d_dgrad_fused: the ReLU mask needs the BatchNorm(s) behind it");
  DFB_REQUIRE(dgrad_mode == DFB_DGRAD_REFERENCE || dgrad_mode == DFB_DGRAD_EXACT, DFB_ERR_INVALID,
              "conv2d_dgrad_fused: bad dgrad_mode %d", dgrad_mode);
  dfb_status st = check_mode("conv2d_dgrad_fused", mode);
  if (st != DFB_OK) return st;
  stat_slot_drop(sums);
  if (want_tc(mode) && dgrad_mode == DFB_DGRAD_EXACT) {
    ConvFuse f{};
    f.lazy = lazy;
    f.addend = addend;
    f.kind = n_bn ? FUSE_BNBWD : FUSE_NONE;
    f.n_sets = n_bn;
    f.stats_out = sums;
    f.bn_x[0] = bn_x0; f.bn_mean[0] = bn_mean0; f.bn_invstd[0] = bn_invstd0;
    f.bn_x[1] = bn_x1; f.bn_mean[1] = bn_mean1; f.bn_invstd[1] = bn_invstd1;
    f.relu = relu;
    f.bn_gamma[0] = gamma0; f.bn_beta[0] = beta0; f.bn_gamma[1] = gamma1; f.bn_beta[1] = beta1;
    f.relu_res = relu_res;
    bool handled = false;
    st = tc_conv_dgrad(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, mode, nullptr, 0, &handled, &f);
    if (st != DFB_OK || handled) return st;
  }
  st = simt_conv_dgrad(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, dgrad_mode);
  if (st != DFB_OK) return st;
  const size_t n = (size_t)N * C * H * W;
  if (addend) {
    st = dfb_ewise_add(dx, addend, dx, n);
    if (st != DFB_OK) return st;
  }
  if (relu) {
    st = dfb_relu_bwd_bn(bn_x0, bn_mean0, bn_invstd0, gamma0, beta0, n_bn > 1 ? bn_x1 : nullptr, bn_mean1, bn_invstd1, gamma1, beta1,
                         relu_res, dx, dx, (size_t)N * H * W, C);
    if (st != DFB_OK) return st;
  }
  if (n_bn >= 1) {
    st = dfb_bn_bwd_sums(bn_x0, dx, bn_mean0, bn_invstd0, sums, sums + C, (size_t)N * H * W, C);
    if (st != DFB_OK) return st;
  }
  if (n_bn >= 2) {  // the first row (sum dx) is the same for both; it is written twice with the same values
    float* tmp = nullptr;
    st = dfb_malloc((size_t)C, &tmp);
    if (st != DFB_OK) return st;
    st = dfb_bn_bwd_sums(bn_x1, dx, bn_mean1, bn_invstd1, tmp, sums + 2 * (size_t)C, (size_t)N * H * W, C);
    dfb_free(tmp);
    if (st != DFB_OK) return st;
  }
  return DFB_OK;
}

dfb_status dfb_conv2d_dgrad_fused(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                                  int R, int pad, int stride, int mode, int dgrad_mode, const float* addend, int n_bn,
                                  const float* bn_x0, const float* bn_mean0, const float* bn_invstd0, const float* bn_x1,
                                  const float* bn_mean1, const float* bn_invstd1, float* sums, int relu, const float* gamma0,
                                  const float* beta0, const float* gamma1, const float* beta1, const float* relu_res) {
  return dgrad_fused_impl(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, mode, dgrad_mode, addend, n_bn, bn_x0, bn_mean0, bn_invstd0,
                          bn_x1, bn_mean1, bn_invstd1, sums, relu, gamma0, beta0, gamma1, beta1, relu_res, 0);
}
dfb_status dfb_conv2d_dgrad_fused_lazy(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                                       int R, int pad, int stride, int mode, int dgrad_mode, const float* addend, int n_bn,
                                       const float* bn_x0, const float* bn_mean0, const float* bn_invstd0, const float* bn_x1,
                                       const float* bn_mean1, const float* bn_invstd1, float* sums, int relu, const float* gamma0,
                                       const float* beta0, const float* gamma1, const float* beta1, const float* relu_res) {
  return dgrad_fused_impl(dy, w, w_layout, dx, N, C, H, W, K, R, pad, stride, mode, dgrad_mode, addend, n_bn, bn_x0, bn_mean0, bn_invstd0,
                          bn_x1, bn_mean1, bn_invstd1, sums, relu, gamma0, beta0, gamma1, beta1, relu_res, 1);
}

// ---- first layer (image input) through its column matrix ---------------------------------------------------------------
dfb_status dfb_stem_cols(const float* x, int x_layout, float* col, int N, int C, int H, int W, int R, int pad, int stride,
                         int w_layout) {
  DFB_INIT();
  DFB_REQUIRE(x && col, DFB_ERR_INVALID, "stem_cols: null pointer");
  DFB_REQUIRE(C >= 1 && R >= 1 && C * R * R <= 32 && stride >= 1 && H + 2 * pad >= R && W + 2 * pad >= R, DFB_ERR_INVALID,
              "stem_cols: needs C*R*R <= 32 (got C=%d R=%d)", C, R);
  DFB_REQUIRE((reinterpret_cast<uintptr_t>(col) & 15) == 0, DFB_ERR_INVALID, "stem_cols: col must be 16-byte aligned");
  return tc_stem_cols(x, x_layout, col, N, C, H, W, R, pad, stride, w_layout);
}
dfb_status dfb_stem_pad_weights(const float* w, float* wp, int K, int cols) {
  DFB_INIT();
  DFB_REQUIRE(w && wp && K > 0 && cols > 0 && cols <= 32, DFB_ERR_INVALID, "stem_pad_weights: bad arguments");
  return tc_stem_pad_weights(w, wp, K, cols);
}
dfb_status dfb_conv2d_wgrad_cols(const float* col, const float* dy, float* dw, int w_layout, int N, int OH, int OW, int K, int cols,
                                 int mode) {
  DFB_INIT();
  DFB_REQUIRE(col && dy && dw, DFB_ERR_INVALID, "conv2d_wgrad_cols: null pointer");
  DFB_REQUIRE(cols > 0 && cols <= 32, DFB_ERR_INVALID, "conv2d_wgrad_cols: cols must be in 1..32");
  bool handled = false;
  dfb_status st = DFB_OK;
  if (want_tc(mode)) {
    st = tc_wgrad_cols(col, dy, dw, w_layout, N, OH, OW, K, cols, mode, &handled);
    if (st != DFB_OK || handled) return st;
  }
  // exact path: dW[k][j] = sum_pix dy[pix][k] * col[pix][j], a GEMM with both operands read transposed
  float* full = nullptr;
  st = dfb_malloc((size_t)K * 32, &full);
  if (st != DFB_OK) return st;
  st = simt_gemm(dy, col, full, K, 32, N * OH * OW, 1, 0, K, 32, 32, 0, nullptr);
  if (st == DFB_OK) {  // keep the first `cols` of every 32-wide row
    const int32_t shape[2] = {K, cols}, strides[2] = {32, 1};
    st = dfb_compact(full, dw, (size_t)K * cols, 2, shape, strides, 0);
  }
  dfb_free(full);
  return st;
}

}  // extern "C"
