// Exact-fp32 (FFMA) GEMM and implicit-GEMM convolution kernels.
//
// This is the DFB_MODE_FP32 / DFB_MODE_SIMT path: fp32 operands, fp32 FMA accumulation, like the
// reference's MatmulKernel (DeepFlows/backend/backend_src/ndarray_backend_cuda.cu:443-466) but
// shared-memory tiled (64x64x16 tiles, 4x4 register blocking) instead of one thread per output
// with no reuse. It is also the fallback for shapes the tcgen05 kernels (gemm_tc.cu) do not take
// (channel counts that are not a multiple of 4, tiny problems).
//
// The same kernel body serves plain GEMMs and the three convolution passes; what differs is the
// pair of *loaders* that map a (row, k) / (k, col) coordinate to a global-memory element:
//   conv fprop : A = im2col(x) gathered on the fly (never materialised, cf. the reference's
//                __im2col2d, DeepFlows/nn/functional.py:249-283), B = Conv2d.weight read in its
//                native (K,C,R,R) layout
//   conv dgrad : A = gathered dy (transposed convolution), B = weight
//   conv wgrad : A = dy^T, B = im2col(x); output written directly in (K,C,R,R) order
#include "common.cuh"

namespace dfb {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kGemmThreads = (BM / TM) * (BN / TN);  // 256

// ---- loaders ----------------------------------------------------------------------------------
// Interface: struct MN / struct KK hold whatever can be pre-decoded for a row (or column) index
// and for a reduction index; load(mn, kk) returns the element or 0 when out of range.

template <bool TRANS>  // TRANS=false: element (i, k) at p[i*ld + k]; true: p[k*ld + i]
struct PlainLoader {
  static constexpr bool kContigK = !TRANS;
  const float* p;
  int ld, n_mn, n_k;
  struct MN { int i; bool ok; };
  struct KK { int k; bool ok; };
  __device__ MN mn(int i) const { return {i, i < n_mn}; }
  __device__ KK kk(int k) const { return {k, k < n_k}; }
  __device__ float load(const MN& a, const KK& b) const {
    if (!(a.ok && b.ok)) return 0.f;
    return TRANS ? __ldg(p + (size_t)b.k * ld + a.i) : __ldg(p + (size_t)a.i * ld + b.k);
  }
};

struct ConvGeom {
  int N, C, H, W, K, R, pad, stride, OH, OW;
  int krsc;  // weights (and dW) stored channels-last (K,R,R,C) instead of (K,C,R,R)
};
__device__ __forceinline__ size_t weight_index(const ConvGeom& g, int kout, int c, int tap) {
  const int RR = g.R * g.R;
  return g.krsc ? ((size_t)kout * RR + tap) * g.C + c : ((size_t)kout * g.C + c) * RR + tap;
}

// A operand of fprop: rows = output pixels (n, oh, ow), k = (r, s, c) with c fastest.
template <bool NCHW>
struct FpropALoader {
  static constexpr bool kContigK = true;
  const float* x;
  ConvGeom g;
  int M, Kred;
  struct MN { int n, ih0, iw0; bool ok; };
  struct KK { int r, s, c; bool ok; };
  __device__ MN mn(int m) const {
    MN o;
    o.ok = m < M;
    int hw = g.OH * g.OW;
    o.n = m / hw;
    int rem = m - o.n * hw;
    int oh = rem / g.OW, ow = rem - oh * g.OW;
    o.ih0 = oh * g.stride - g.pad;
    o.iw0 = ow * g.stride - g.pad;
    return o;
  }
  __device__ KK kk(int k) const {
    KK o;
    o.ok = k < Kred;
    int tap = k / g.C;
    o.c = k - tap * g.C;
    o.r = tap / g.R;
    o.s = tap - o.r * g.R;
    return o;
  }
  __device__ float load(const MN& a, const KK& b) const {
    if (!(a.ok && b.ok)) return 0.f;
    int ih = a.ih0 + b.r, iw = a.iw0 + b.s;
    if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
    size_t off = NCHW ? (((size_t)a.n * g.C + b.c) * g.H + ih) * g.W + iw
                      : (((size_t)a.n * g.H + ih) * g.W + iw) * g.C + b.c;
    return __ldg(x + off);
  }
};

// B operand of fprop / dgrad: weight (K, C, R, R).
//   fprop: B(k=(r,s,c), n=kout)  = w[kout][c][r][s]
//   dgrad: B(k=(r,s,kout), n=c)  = w[kout][c][r][s]
template <bool DGRAD>
struct WeightLoader {
  static constexpr bool kContigK = true;
  const float* w;
  ConvGeom g;
  int Ncols, Kred;
  struct MN { int i; bool ok; };
  struct KK { int tap, j; bool ok; };
  __device__ MN mn(int i) const { return {i, i < Ncols}; }
  __device__ KK kk(int k) const {
    KK o;
    o.ok = k < Kred;
    int inner = DGRAD ? g.K : g.C;
    o.tap = k / inner;
    o.j = k - o.tap * inner;
    return o;
  }
  __device__ float load(const MN& a, const KK& b) const {
    if (!(a.ok && b.ok)) return 0.f;
    int kout = DGRAD ? b.j : a.i, c = DGRAD ? a.i : b.j;
    return __ldg(w + weight_index(g, kout, c, b.tap));
  }
};

// A operand of the exact dgrad: rows = input pixels (n, h, w), k = (r, s, kout) with kout fastest.
// dx[n,h,w,c] = sum_{r,s,k} dy[n, (h+pad-r)/stride, (w+pad-s)/stride, k] * w[k][c][r][s]
struct DgradALoader {
  static constexpr bool kContigK = true;
  const float* dy;
  ConvGeom g;
  int M, Kred;
  struct MN { int n, hp, wp; bool ok; };
  struct KK { int r, s, k; bool ok; };
  __device__ MN mn(int m) const {
    MN o;
    o.ok = m < M;
    int hw = g.H * g.W;
    o.n = m / hw;
    int rem = m - o.n * hw;
    int h = rem / g.W;
    o.hp = h + g.pad;
    o.wp = rem - h * g.W + g.pad;
    return o;
  }
  __device__ KK kk(int k) const {
    KK o;
    o.ok = k < Kred;
    int tap = k / g.K;
    o.k = k - tap * g.K;
    o.r = tap / g.R;
    o.s = tap - o.r * g.R;
    return o;
  }
  __device__ float load(const MN& a, const KK& b) const {
    if (!(a.ok && b.ok)) return 0.f;
    int th = a.hp - b.r, tw = a.wp - b.s;
    if (th < 0 || tw < 0) return 0.f;
    int oh = th / g.stride, ow = tw / g.stride;
    if (oh * g.stride != th || ow * g.stride != tw || oh >= g.OH || ow >= g.OW) return 0.f;
    return __ldg(dy + (((size_t)a.n * g.OH + oh) * g.OW + ow) * g.K + b.k);
  }
};

// wgrad: dW[kout][(c,r,s)] = sum_p dy[p][kout] * x[n, oh*stride+r-pad, ow*stride+s-pad, c]
struct WgradALoader {  // A(m=kout, k=p) = dy[p*K + kout]
  static constexpr bool kContigK = false;
  const float* dy;
  int K, P;
  struct MN { int i; bool ok; };
  struct KK { int p; bool ok; };
  __device__ MN mn(int i) const { return {i, i < K}; }
  __device__ KK kk(int p) const { return {p, p < P}; }
  __device__ float load(const MN& a, const KK& b) const {
    if (!(a.ok && b.ok)) return 0.f;
    return __ldg(dy + (size_t)b.p * K + a.i);
  }
};
template <bool NCHW>
struct WgradBLoader {  // B(k=p, n=(c,r,s))
  static constexpr bool kContigK = false;
  const float* x;
  ConvGeom g;
  int Ncols, P;
  struct MN { int c, r, s; bool ok; };
  struct KK { int n, ih0, iw0; bool ok; };
  __device__ MN mn(int j) const {  // column j of dW's row: (c, tap) for KCRS, (tap, c) for KRSC
    MN o;
    o.ok = j < Ncols;
    int rr = g.R * g.R;
    int tap;
    if (g.krsc) {
      tap = j / g.C;
      o.c = j - tap * g.C;
    } else {
      o.c = j / rr;
      tap = j - o.c * rr;
    }
    o.r = tap / g.R;
    o.s = tap - o.r * g.R;
    return o;
  }
  __device__ KK kk(int p) const {
    KK o;
    o.ok = p < P;
    int hw = g.OH * g.OW;
    o.n = p / hw;
    int rem = p - o.n * hw;
    int oh = rem / g.OW, ow = rem - oh * g.OW;
    o.ih0 = oh * g.stride - g.pad;
    o.iw0 = ow * g.stride - g.pad;
    return o;
  }
  __device__ float load(const MN& a, const KK& b) const {
    if (!(a.ok && b.ok)) return 0.f;
    int ih = b.ih0 + a.r, iw = b.iw0 + a.s;
    if ((unsigned)ih >= (unsigned)g.H || (unsigned)iw >= (unsigned)g.W) return 0.f;
    size_t off = NCHW ? (((size_t)b.n * g.C + a.c) * g.H + ih) * g.W + iw
                      : (((size_t)b.n * g.H + ih) * g.W + iw) * g.C + a.c;
    return __ldg(x + off);
  }
};

// ---- kernel -----------------------------------------------------------------------------------
struct Epilogue {
  float* C;
  int ldc;
  int accumulate;     // C += acc
  const float* bias;  // per column, may be null
  float* partial;     // split-K: slab z at partial + z*M*N (row pitch N); else null
};

template <class AL, class BL>
__global__ void __launch_bounds__(kGemmThreads)
simt_gemm_kernel(AL A, BL B, Epilogue ep, int M, int N, int K, int k_per_split) {
  pdl_sync();
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);

  // load-slot mapping: 4 elements per thread per operand per k-tile
  //   kContigK : slot i -> (mn = tid/16 + 16*i, k = tid%16)
  //   otherwise: slot i -> (mn = tid%64,        k = tid/64 + 4*i)
  typename AL::MN a_mn[4];
  typename BL::MN b_mn[4];
  if (AL::kContigK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a_mn[i] = A.mn(m0 + tid / 16 + 16 * i);
  } else {
    a_mn[0] = A.mn(m0 + tid % 64);
  }
  if (BL::kContigK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) b_mn[i] = B.mn(n0 + tid / 16 + 16 * i);
  } else {
    b_mn[0] = B.mn(n0 + tid % 64);
  }

  float ra[4], rb[4];
  auto fetch = [&](int k0) {
    if (AL::kContigK) {
      typename AL::KK kk = A.kk(k0 + tid % 16);
      kk.ok = kk.ok && (k0 + tid % 16 < kend);
#pragma unroll
      for (int i = 0; i < 4; ++i) ra[i] = A.load(a_mn[i], kk);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        typename AL::KK kk = A.kk(k0 + tid / 64 + 4 * i);
        kk.ok = kk.ok && (k0 + tid / 64 + 4 * i < kend);
        ra[i] = A.load(a_mn[0], kk);
      }
    }
    if (BL::kContigK) {
      typename BL::KK kk = B.kk(k0 + tid % 16);
      kk.ok = kk.ok && (k0 + tid % 16 < kend);
#pragma unroll
      for (int i = 0; i < 4; ++i) rb[i] = B.load(b_mn[i], kk);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        typename BL::KK kk = B.kk(k0 + tid / 64 + 4 * i);
        kk.ok = kk.ok && (k0 + tid / 64 + 4 * i < kend);
        rb[i] = B.load(b_mn[0], kk);
      }
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (AL::kContigK) As[buf][tid % 16][tid / 16 + 16 * i] = ra[i];
      else As[buf][tid / 64 + 4 * i][tid % 64] = ra[i];
      if (BL::kContigK) Bs[buf][tid % 16][tid / 16 + 16 * i] = rb[i];
      else Bs[buf][tid / 64 + 4 * i][tid % 64] = rb[i];
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (kbeg < kend) {
    fetch(kbeg);
    stage(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) fetch(k0 + BK);  // global loads in flight while this tile is multiplied
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stage(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= N) continue;
      if (ep.partial) {
        ep.partial[((size_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        float v = acc[i][j];
        if (ep.bias) v += __ldg(ep.bias + n);
        float* c = ep.C + (size_t)m * ep.ldc + n;
        *c = ep.accumulate ? *c + v : v;
      }
    }
  }
}

// deterministic split-K reduction (fixed summation order over the slabs)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, Epilogue ep) {
  pdl_sync();
  size_t total = (size_t)M * N;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(size_t)z * total + i];
    int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    if (ep.bias) v += __ldg(ep.bias + n);
    float* c = ep.C + (size_t)m * ep.ldc + n;
    *c = ep.accumulate ? *c + v : v;
  }
}

template <class AL, class BL>
dfb_status launch_simt(const char* name, AL A, BL B, float* C, int ldc, int accumulate,
                       const float* bias, int M, int N, int K) {
  if (M <= 0 || N <= 0) return DFB_OK;
  cudaStream_t s = compute_stream();
  unsigned gx = cdiv(M, BM), gy = cdiv(N, BN);
  DFB_REQUIRE(gy <= 65535, DFB_ERR_INVALID, "%s: N=%d too large for the SIMT kernel", name, N);
  // split the reduction when the output grid cannot fill the machine and K is long
  int splits = 1;
  size_t tiles = (size_t)gx * gy;
  size_t target = (size_t)sm_count() * 2;
  if (tiles < target && K >= 4 * BK * 2) {
    splits = (int)std::min<size_t>((target + tiles - 1) / tiles, (size_t)K / (4 * BK));
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
  }
  int k_per_split = (int)(((size_t)K + splits - 1) / splits);
  k_per_split = (k_per_split + BK - 1) / BK * BK;
  splits = (K + k_per_split - 1) / k_per_split;
  if (splits < 1) splits = 1;
  Epilogue ep{C, ldc, accumulate, bias, nullptr};
  if (K == 0) {  // empty reduction: C = bias / unchanged
    k_per_split = BK;
    splits = 1;
  }
  float* partial = nullptr;
  if (splits > 1) {
    dfb_status st = dfb_malloc((size_t)splits * M * N, &partial);
    if (st != DFB_OK) return st;
    ep.partial = partial;
  }
  launch_k(simt_gemm_kernel<AL, BL>, dim3(gx, gy, splits), kGemmThreads, 0, s, A, B, ep, M, N, K, k_per_split);
  DFB_LAUNCH_CHECK(name);
  if (splits > 1) {
    ep.partial = nullptr;
    launch_k(splitk_reduce_kernel, bw_grid((size_t)M * N, 256), 256, 0, s, partial, splits, M, N, ep);
    DFB_LAUNCH_CHECK(name);
    dfb_free(partial);
  }
  return DFB_OK;
}

// ---- last-writer-wins dgrad (reference semantics) ------------------------------------
// The reference scatters d(col) back with plain assignment in (i, j) loop order
// (DeepFlows/nn/functional.py:285-294), so each padded input pixel keeps only the tap with the
// largest valid i and the largest valid j. One thread per (input pixel, channel).
__global__ void __launch_bounds__(256)
dgrad_lastwriter_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                        float* __restrict__ dx, ConvGeom g) {
  pdl_sync();
  size_t total = (size_t)g.N * g.H * g.W * g.C;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int c = (int)(i % g.C);
    size_t pix = i / g.C;
    int wq = (int)(pix % g.W);
    size_t t = pix / g.W;
    int h = (int)(t % g.H);
    int n = (int)(t / g.H);
    int ph = h + g.pad, pw = wq + g.pad;
    // largest tap index r <= min(R-1, ph) with (ph - r) % stride == 0
    int rmax = min(g.R - 1, ph), smax = min(g.R - 1, pw);
    int r = rmax - (g.stride - (ph - rmax) % g.stride) % g.stride;
    int sidx = smax - (g.stride - (pw - smax) % g.stride) % g.stride;
    float acc = 0.f;
    if (r >= 0 && sidx >= 0) {
      int oh = (ph - r) / g.stride, ow = (pw - sidx) / g.stride;
      if (oh < g.OH && ow < g.OW) {
        const float* dyp = dy + (((size_t)n * g.OH + oh) * g.OW + ow) * g.K;
        const int tap = r * g.R + sidx;
        for (int k = 0; k < g.K; ++k) acc = fmaf(__ldg(dyp + k), __ldg(w + weight_index(g, k, c, tap)), acc);
      }
    }
    dx[i] = acc;
  }
}

static dfb_status make_geom(const char* name, int N, int C, int H, int W, int K, int R, int pad,
                            int stride, int w_layout, ConvGeom* g) {
  DFB_REQUIRE(w_layout == DFB_WLAYOUT_KCRS || w_layout == DFB_WLAYOUT_KRSC, DFB_ERR_INVALID, "%s: bad weight layout %d", name,
              w_layout);
  g->krsc = w_layout == DFB_WLAYOUT_KRSC;
  DFB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && K > 0 && R > 0 && pad >= 0 && stride > 0,
              DFB_ERR_INVALID, "%s: bad geometry N=%d C=%d H=%d W=%d K=%d R=%d pad=%d stride=%d", name, N,
              C, H, W, K, R, pad, stride);
  DFB_REQUIRE(H + 2 * pad >= R && W + 2 * pad >= R, DFB_ERR_INVALID,
              "%s: kernel %d larger than padded input %dx%d", name, R, H + 2 * pad, W + 2 * pad);
  g->N = N; g->C = C; g->H = H; g->W = W; g->K = K; g->R = R; g->pad = pad; g->stride = stride;
  g->OH = (H + 2 * pad - R) / stride + 1;
  g->OW = (W + 2 * pad - R) / stride + 1;
  DFB_REQUIRE((size_t)N * g->OH * g->OW < ((size_t)1 << 31) && (size_t)N * H * W < ((size_t)1 << 31),
              DFB_ERR_INVALID, "%s: more than 2^31 pixels", name);
  return DFB_OK;
}

// entry points used by the dispatchers in conv.cu / gemm.cu -----------------------------------------
dfb_status simt_gemm(const float* A, const float* B, float* C, int M, int N, int K, int trans_a,
                     int trans_b, int lda, int ldb, int ldc, int accumulate, const float* bias) {
  // A(m,k): trans_a ? A[k*lda+m] : A[m*lda+k];  B(k,n): trans_b ? B[n*ldb+k] : B[k*ldb+n]
  // The B loader indexes (col, k): "TRANS" for it means element (n, k) at p[k*ld + n].
  if (!trans_a && !trans_b)
    return launch_simt("Matmul", PlainLoader<false>{A, lda, M, K}, PlainLoader<true>{B, ldb, N, K}, C, ldc, accumulate, bias, M, N, K);
  if (!trans_a && trans_b)
    return launch_simt("Matmul", PlainLoader<false>{A, lda, M, K}, PlainLoader<false>{B, ldb, N, K}, C, ldc, accumulate, bias, M, N, K);
  if (trans_a && !trans_b)
    return launch_simt("Matmul", PlainLoader<true>{A, lda, M, K}, PlainLoader<true>{B, ldb, N, K}, C, ldc, accumulate, bias, M, N, K);
  return launch_simt("Matmul", PlainLoader<true>{A, lda, M, K}, PlainLoader<false>{B, ldb, N, K}, C, ldc, accumulate, bias, M, N, K);
}

dfb_status simt_conv_fprop(const float* x, int x_layout, const float* w, int w_layout, float* y, int N, int C, int H,
                           int W, int K, int R, int pad, int stride) {
  ConvGeom g;
  dfb_status st = make_geom("conv2d_fprop", N, C, H, W, K, R, pad, stride, w_layout, &g);
  if (st != DFB_OK) return st;
  int M = N * g.OH * g.OW, Kred = C * R * R;
  WeightLoader<false> B{w, g, K, Kred};
  if (x_layout == DFB_LAYOUT_NCHW)
    return launch_simt("conv2d_fprop", FpropALoader<true>{x, g, M, Kred}, B, y, K, 0, nullptr, M, K, Kred);
  return launch_simt("conv2d_fprop", FpropALoader<false>{x, g, M, Kred}, B, y, K, 0, nullptr, M, K, Kred);
}

dfb_status simt_conv_dgrad(const float* dy, const float* w, int w_layout, float* dx, int N, int C, int H, int W, int K,
                           int R, int pad, int stride, int dgrad_mode) {
  ConvGeom g;
  dfb_status st = make_geom("conv2d_dgrad", N, C, H, W, K, R, pad, stride, w_layout, &g);
  if (st != DFB_OK) return st;
  if (dgrad_mode == DFB_DGRAD_REFERENCE) {
    size_t total = (size_t)N * H * W * C;
    launch_k(dgrad_lastwriter_kernel, bw_grid(total, 256), 256, 0, compute_stream(), dy, w, dx, g);
    DFB_LAUNCH_CHECK("conv2d_dgrad(reference)");
    return DFB_OK;
  }
  int M = N * H * W, Kred = K * R * R;
  return launch_simt("conv2d_dgrad", DgradALoader{dy, g, M, Kred}, WeightLoader<true>{w, g, C, Kred}, dx, C, 0,
                     nullptr, M, C, Kred);
}

dfb_status simt_conv_wgrad(const float* x, int x_layout, const float* dy, float* dw, int w_layout, int N, int C, int H,
                           int W, int K, int R, int pad, int stride) {
  ConvGeom g;
  dfb_status st = make_geom("conv2d_wgrad", N, C, H, W, K, R, pad, stride, w_layout, &g);
  if (st != DFB_OK) return st;
  int P = N * g.OH * g.OW, Ncols = C * R * R;
  WgradALoader A{dy, K, P};
  if (x_layout == DFB_LAYOUT_NCHW)
    return launch_simt("conv2d_wgrad", A, WgradBLoader<true>{x, g, Ncols, P}, dw, Ncols, 0, nullptr, K, Ncols, P);
  return launch_simt("conv2d_wgrad", A, WgradBLoader<false>{x, g, Ncols, P}, dw, Ncols, 0, nullptr, K, Ncols, P);
}

}  // namespace dfb
