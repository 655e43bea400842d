// L0 bandwidth-bound kernels: fill, the elementwise/scalar family, strided gather/scatter
// (compact / setitem) and trailing-axis reductions.
//
// Reference semantics: DeepFlows/backend/backend_src/ndarray_backend_cuda.cu:127-509 (one thread
// per element, scalar accesses, one thread per reduced row). Here: 128-bit vectorised grid-stride
// kernels sized to the SM count, dimension-collapsing + tiled transposes for the strided copies,
// warp-shuffle / block reductions.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace dfb {

// ---------------------------------------------------------------------------------------------
// elementwise
// ---------------------------------------------------------------------------------------------
struct OpAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMul { __device__ float operator()(float a, float b) const { return a * b; } };
struct OpDiv { __device__ float operator()(float a, float b) const { return a / b; } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpEq  { __device__ float operator()(float a, float b) const { return a == b ? 1.0f : 0.0f; } };
struct OpGe  { __device__ float operator()(float a, float b) const { return a >= b ? 1.0f : 0.0f; } };
struct OpPow { __device__ float operator()(float a, float b) const { return powf(a, b); } };
// cu:405 — log of a non-positive number is -inf, never NaN
struct OpLog { __device__ float operator()(float a) const { return a > 0.0f ? logf(a) : -INFINITY; } };
struct OpExp { __device__ float operator()(float a) const { return expf(a); } };
struct OpTanh { __device__ float operator()(float a) const { return tanhf(a); } };

constexpr int kThreads = 256;

template <class Op>
__global__ void __launch_bounds__(kThreads) binary_vec_kernel(const float4* __restrict__ a,
                                                              const float4* __restrict__ b,
                                                              float4* __restrict__ out, size_t n4,
                                                              Op op) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = ld_stream(a + i), y = ld_stream(b + i), r;
    r.x = op(x.x, y.x); r.y = op(x.y, y.y); r.z = op(x.z, y.z); r.w = op(x.w, y.w);
    st_stream(out + i, r);
  }
}
template <class Op>
__global__ void __launch_bounds__(kThreads) binary_tail_kernel(const float* __restrict__ a,
                                                               const float* __restrict__ b,
                                                               float* __restrict__ out, size_t begin,
                                                               size_t n, Op op) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = op(a[i], b[i]);
}
template <class Op>
__global__ void __launch_bounds__(kThreads) scalar_vec_kernel(const float4* __restrict__ a, float v,
                                                              float4* __restrict__ out, size_t n4,
                                                              Op op) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = ld_stream(a + i), r;
    r.x = op(x.x, v); r.y = op(x.y, v); r.z = op(x.z, v); r.w = op(x.w, v);
    st_stream(out + i, r);
  }
}
template <class Op>
__global__ void __launch_bounds__(kThreads) scalar_tail_kernel(const float* __restrict__ a, float v,
                                                               float* __restrict__ out, size_t begin,
                                                               size_t n, Op op) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = op(a[i], v);
}
template <class Op>
__global__ void __launch_bounds__(kThreads) unary_vec_kernel(const float4* __restrict__ a,
                                                             float4* __restrict__ out, size_t n4,
                                                             Op op) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 x = ld_stream(a + i), r;
    r.x = op(x.x); r.y = op(x.y); r.z = op(x.z); r.w = op(x.w);
    st_stream(out + i, r);
  }
}
template <class Op>
__global__ void __launch_bounds__(kThreads) unary_tail_kernel(const float* __restrict__ a,
                                                              float* __restrict__ out, size_t begin,
                                                              size_t n, Op op) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = op(a[i]);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class Op>
dfb_status run_binary(const char* name, const float* a, const float* b, float* out, size_t n, Op op) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "%s: out array cannot be null", name);
  if (n == 0) return DFB_OK;
  cudaStream_t s = compute_stream();
  size_t n4 = (aligned16(a) && aligned16(b) && aligned16(out)) ? n / 4 : 0;
  if (n4) {
    launch_k(binary_vec_kernel<Op>, bw_grid(n4, kThreads), kThreads, 0, s, (const float4*)a, (const float4*)b,
                                                                 (float4*)out, n4, op);
    DFB_LAUNCH_CHECK(name);
  }
  if (n4 * 4 < n) {
    launch_k(binary_tail_kernel<Op>, bw_grid(n - n4 * 4, kThreads), kThreads, 0, s, a, b, out, n4 * 4, n, op);
    DFB_LAUNCH_CHECK(name);
  }
  return DFB_OK;
}
template <class Op>
dfb_status run_scalar(const char* name, const float* a, float v, float* out, size_t n, Op op) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "%s: out array cannot be null", name);
  if (n == 0) return DFB_OK;
  cudaStream_t s = compute_stream();
  size_t n4 = (aligned16(a) && aligned16(out)) ? n / 4 : 0;
  if (n4) {
    launch_k(scalar_vec_kernel<Op>, bw_grid(n4, kThreads), kThreads, 0, s, (const float4*)a, v, (float4*)out, n4, op);
    DFB_LAUNCH_CHECK(name);
  }
  if (n4 * 4 < n) {
    launch_k(scalar_tail_kernel<Op>, bw_grid(n - n4 * 4, kThreads), kThreads, 0, s, a, v, out, n4 * 4, n, op);
    DFB_LAUNCH_CHECK(name);
  }
  return DFB_OK;
}
template <class Op>
dfb_status run_unary(const char* name, const float* a, float* out, size_t n, Op op) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "%s: out array cannot be null", name);
  if (n == 0) return DFB_OK;
  cudaStream_t s = compute_stream();
  size_t n4 = (aligned16(a) && aligned16(out)) ? n / 4 : 0;
  if (n4) {
    launch_k(unary_vec_kernel<Op>, bw_grid(n4, kThreads), kThreads, 0, s, (const float4*)a, (float4*)out, n4, op);
    DFB_LAUNCH_CHECK(name);
  }
  if (n4 * 4 < n) {
    launch_k(unary_tail_kernel<Op>, bw_grid(n - n4 * 4, kThreads), kThreads, 0, s, a, out, n4 * 4, n, op);
    DFB_LAUNCH_CHECK(name);
  }
  return DFB_OK;
}

// ---------------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) fill_kernel(float* __restrict__ out, float v, size_t n) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // head (to 16-byte alignment) and tail scalars are handled by the first threads
  size_t head = ((16 - (reinterpret_cast<uintptr_t>(out) & 15)) & 15) / 4;
  if (head > n) head = n;
  size_t n4 = (n - head) / 4;
  float4* o4 = reinterpret_cast<float4*>(out + head);
  float4 v4 = make_float4(v, v, v, v);
  for (size_t j = i; j < n4; j += stride) o4[j] = v4;
  if (i < head) out[i] = v;
  size_t tail_begin = head + n4 * 4;
  if (tail_begin + i < n && i < 4) out[tail_begin + i] = v;
}

// ---------------------------------------------------------------------------------------------
// strided gather / scatter (compact, ewise_setitem, scalar_setitem)
// ---------------------------------------------------------------------------------------------
struct StridedView {
  int ndim;
  uint32_t shape[DFB_MAX_DIMS];
  int64_t stride[DFB_MAX_DIMS];
};

// Merge adjacent dimensions that are contiguous with respect to each other and drop size-1
// dimensions: (N,C,H,W) views produced by permute / slicing usually collapse to 2-3 dims, which
// removes most of the div/mod chain of the reference's gid_to_idx (cu:147-155).
static StridedView collapse(int ndim, const int32_t* shape, const int32_t* strides) {
  StridedView v;
  v.ndim = 0;
  for (int d = 0; d < ndim; ++d) {
    if (shape[d] == 1) continue;
    if (v.ndim > 0 && v.stride[v.ndim - 1] == (int64_t)strides[d] * shape[d]) {
      v.shape[v.ndim - 1] *= (uint32_t)shape[d];
      v.stride[v.ndim - 1] = strides[d];
    } else {
      v.shape[v.ndim] = (uint32_t)shape[d];
      v.stride[v.ndim] = strides[d];
      v.ndim++;
    }
  }
  if (v.ndim == 0) {
    v.ndim = 1;
    v.shape[0] = 1;
    v.stride[0] = 1;
  }
  return v;
}

__device__ __forceinline__ int64_t view_index(size_t gid, const StridedView& v) {
  int64_t idx = 0;
#pragma unroll
  for (int d = DFB_MAX_DIMS - 1; d >= 0; --d) {
    if (d < v.ndim) {
      uint32_t s = v.shape[d];
      size_t q = gid / s;
      uint32_t r = (uint32_t)(gid - q * s);
      idx += (int64_t)r * v.stride[d];
      gid = q;
    }
  }
  return idx;
}

// mode 0: out[gid] = a[view(gid)]   (compact)
// mode 1: out[view(gid)] = a[gid]   (ewise_setitem)
// mode 2: out[view(gid)] = value    (scalar_setitem)
template <int MODE>
__global__ void __launch_bounds__(kThreads) strided_kernel(const float* __restrict__ a,
                                                           float* __restrict__ out, float value,
                                                           size_t n, StridedView v, int64_t offset) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += stride) {
    int64_t idx = offset + view_index(gid, v);
    if (MODE == 0) out[gid] = a[idx];
    else if (MODE == 1) out[idx] = a[gid];
    else out[idx] = value;
  }
}

// out[gid] = a[view(gid)] * scale: a broadcast / permuted view made compact and scaled on the way (the backward of
// `mean`: ones * grad * (out.size / x.size), tensor.py:763-766 - one pass instead of scalar_mul + compact)
__global__ void __launch_bounds__(kThreads) strided_scale_kernel(const float* __restrict__ a, float* __restrict__ out, float scale,
                                                                 size_t n, StridedView v, int64_t offset) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; gid < n; gid += stride) out[gid] = a[offset + view_index(gid, v)] * scale;
}
// out[gid] = (sum over j < len of a[view(gid) + j * rstride], ascending j, like reduce_thread_kernel) / divisor: the
// reduction of the LAST axis of a strided view without making it compact first, the division of `mean` folded in
// (sum: divisor 1). One thread per output; len <= 32.
__global__ void __launch_bounds__(kThreads) reduce_view_div_kernel(const float* __restrict__ a, float* __restrict__ out, size_t rows,
                                                                   uint32_t len, int64_t rstride, float divisor, StridedView v,
                                                                   int64_t offset) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const float* p = a + offset + view_index(r, v);
    float acc = p[0];
    for (uint32_t i = 1; i < len; ++i) acc = acc + p[(int64_t)i * rstride];
    out[r] = acc / divisor;
  }
}

// Batched 2-D transpose through shared memory for the permute case, i.e. a collapsed view of
// the form (outer..., R, C) where the *input* is contiguous along R (stride 1) and the output is
// contiguous along C. in index = offset + outer_off + r*1 + c*sC ; out index = o*R*C + r*C + c.
// mode 0 gathers (compact); mode 1 scatters (setitem) with the roles of a/out swapped.
struct TransposeView {
  int n_outer;
  uint32_t outer_shape[DFB_MAX_DIMS];
  int64_t outer_stride[DFB_MAX_DIMS];
  uint32_t R, C;          // compact side is (outer, R, C) row-major
  int64_t sR, sC;         // strided side element strides for r and c; sR == 1
};

template <int MODE>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ a,
                                                        float* __restrict__ out, TransposeView v,
                                                        int64_t offset, uint32_t tiles_r,
                                                        uint32_t tiles_c) {
  pdl_sync();
  __shared__ float tile[32][33];
  size_t n_tiles = (size_t)tiles_r * tiles_c;
  size_t n_outer = 1;
  for (int d = 0; d < v.n_outer; ++d) n_outer *= v.outer_shape[d];
  size_t total = n_tiles * n_outer;
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (size_t t = blockIdx.x; t < total; t += gridDim.x) {
    size_t o = t / n_tiles;
    uint32_t tt = (uint32_t)(t - o * n_tiles);
    uint32_t tr = tt / tiles_c, tc = tt - tr * tiles_c;
    int64_t obase = 0;
    size_t oo = o;
    for (int d = v.n_outer - 1; d >= 0; --d) {
      size_t q = oo / v.outer_shape[d];
      obase += (int64_t)(oo - q * v.outer_shape[d]) * v.outer_stride[d];
      oo = q;
    }
    const size_t cbase = o * (size_t)v.R * v.C;
    uint32_t r0 = tr * 32, c0 = tc * 32;
    if (MODE == 0) {
      // read strided side with r fastest (coalesced, stride 1), write compact side with c fastest
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t r = r0 + tx, c = c0 + ty + j;
        if (r < v.R && c < v.C) tile[ty + j][tx] = a[offset + obase + (int64_t)r * v.sR + (int64_t)c * v.sC];
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t r = r0 + ty + j, c = c0 + tx;
        if (r < v.R && c < v.C) out[cbase + (size_t)r * v.C + c] = tile[tx][ty + j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t r = r0 + ty + j, c = c0 + tx;
        if (r < v.R && c < v.C) tile[tx][ty + j] = a[cbase + (size_t)r * v.C + c];
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t r = r0 + tx, c = c0 + ty + j;
        if (r < v.R && c < v.C) out[offset + obase + (int64_t)r * v.sR + (int64_t)c * v.sC] = tile[ty + j][tx];
      }
    }
    __syncthreads();
  }
}

static dfb_status check_view(const char* name, int ndim, const int32_t* shape) {
  DFB_REQUIRE(ndim >= 0 && ndim <= DFB_MAX_DIMS, DFB_ERR_INVALID,
              "%s: CUDA dimension limit exceeded: max supported dimensions = %d, requested = %d", name,
              DFB_MAX_DIMS, ndim);
  for (int d = 0; d < ndim; ++d)
    DFB_REQUIRE(shape[d] > 0, DFB_ERR_INVALID, "%s: non-positive extent %d in dim %d", name, shape[d], d);
  return DFB_OK;
}

// Try the tiled-transpose fast path. Returns true when it handled the copy.
template <int MODE>
static bool try_transpose(const float* a, float* out, const StridedView& v, int64_t offset, cudaStream_t s) {
  if (v.ndim < 2) return false;
  // the last collapsed dim is the compact side's fastest dim (C); look for a dim with stride 1
  int last = v.ndim - 1;
  if (v.stride[last] == 1 || v.stride[last] <= 0) return false;
  int rdim = -1;
  for (int d = 0; d < last; ++d)
    if (v.stride[d] == 1) rdim = d;
  if (rdim < 0) return false;
  // need R to be the second-to-last compact dim for a plain 2-D tile; otherwise dims between
  // rdim and last sit inside the tile's row pitch. Handle the common case rdim == last-1, and
  // the case where everything between is folded into "outer" by treating compact pitch properly
  // only when rdim == last-1.
  if (rdim != last - 1) return false;
  if (v.shape[rdim] < 8 || v.shape[last] < 8) return false;
  TransposeView t;
  t.n_outer = 0;
  for (int d = 0; d < rdim; ++d) {
    t.outer_shape[t.n_outer] = v.shape[d];
    t.outer_stride[t.n_outer] = v.stride[d];
    t.n_outer++;
  }
  t.R = v.shape[rdim];
  t.C = v.shape[last];
  t.sR = 1;
  t.sC = v.stride[last];
  uint32_t tiles_r = cdiv(t.R, 32), tiles_c = cdiv(t.C, 32);
  size_t total = (size_t)tiles_r * tiles_c;
  for (int d = 0; d < t.n_outer; ++d) total *= t.outer_shape[d];
  unsigned grid = (unsigned)std::min<size_t>(total, (size_t)sm_count() * 16);
  launch_k(transpose_kernel<MODE>, grid, 256, 0, s, a, out, t, offset, tiles_r, tiles_c);
  return true;
}

}  // namespace dfb

using namespace dfb;

extern "C" {

dfb_status dfb_fill(float* out, float value, size_t n) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "Fill: out array cannot be null");
  if (n == 0) return DFB_OK;
  launch_k(fill_kernel, bw_grid(n / 4 + 1, kThreads), kThreads, 0, compute_stream(), out, value, n);
  DFB_LAUNCH_CHECK("Fill");
  return DFB_OK;
}

dfb_status dfb_compact(const float* a, float* out, size_t out_size, int ndim, const int32_t* shape,
                       const int32_t* strides, size_t offset) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "Compact: out array cannot be null");
  DFB_REQUIRE(out_size != 0, DFB_ERR_INVALID, "Compact: out array size cannot be zero");
  dfb_status st = check_view("Compact", ndim, shape);
  if (st != DFB_OK) return st;
  StridedView v = collapse(ndim, shape, strides);
  cudaStream_t s = compute_stream();
  if (v.ndim == 1 && v.stride[0] == 1) {  // contiguous slice
    DFB_CUDA(cudaMemcpyAsync(out, a + offset, out_size * sizeof(float), cudaMemcpyDeviceToDevice, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DFB_OK;
  }
  if (try_transpose<0>(a, out, v, (int64_t)offset, s)) {
    DFB_LAUNCH_CHECK("Compact(transpose)");
    return DFB_OK;
  }
  launch_k(strided_kernel<0>, bw_grid(out_size, kThreads), kThreads, 0, s, a, out, 0.f, out_size, v, (int64_t)offset);
  DFB_LAUNCH_CHECK("Compact");
  return DFB_OK;
}

dfb_status dfb_compact_scale(const float* a, float* out, size_t out_size, int ndim, const int32_t* shape, const int32_t* strides,
                             size_t offset, float scale) {
  DFB_INIT();
  DFB_REQUIRE(a != nullptr && out != nullptr, DFB_ERR_INVALID, "CompactScale: null array");
  DFB_REQUIRE(out_size != 0, DFB_ERR_INVALID, "CompactScale: out array size cannot be zero");
  dfb_status st = check_view("CompactScale", ndim, shape);
  if (st != DFB_OK) return st;
  StridedView v = collapse(ndim, shape, strides);
  launch_k(strided_scale_kernel, bw_grid(out_size, kThreads), kThreads, 0, compute_stream(), a, out, scale, out_size, v, (int64_t)offset);
  DFB_LAUNCH_CHECK("CompactScale");
  return DFB_OK;
}

dfb_status dfb_reduce_sum_view_div(const float* a, float* out, size_t out_size, int ndim, const int32_t* shape, const int32_t* strides,
                                   size_t offset, float divisor) {
  DFB_INIT();
  DFB_REQUIRE(a != nullptr && out != nullptr, DFB_ERR_INVALID, "ReduceSumView: null array");
  DFB_REQUIRE(out_size != 0 && ndim >= 1, DFB_ERR_INVALID, "ReduceSumView: empty output");
  dfb_status st = check_view("ReduceSumView", ndim, shape);
  if (st != DFB_OK) return st;
  const int32_t len = shape[ndim - 1];
  DFB_REQUIRE(len >= 1 && len <= 32, DFB_ERR_INVALID, "ReduceSumView: the reduced axis has %d elements (1..32 supported; longer rows: compact + reduce_sum)", len);
  DFB_REQUIRE(divisor != 0.f, DFB_ERR_DOMAIN, "ReduceSumView: division by zero");
  size_t rows = 1;
  for (int d = 0; d + 1 < ndim; ++d) rows *= (size_t)shape[d];
  DFB_REQUIRE(rows == out_size, DFB_ERR_INVALID, "ReduceSumView: out.size != product of the outer shape");
  StridedView v = collapse(ndim - 1, shape, strides);   // (ndim == 1: one output, collapse() returns the unit view)
  launch_k(reduce_view_div_kernel, bw_grid(out_size, kThreads), kThreads, 0, compute_stream(), a, out, out_size, (uint32_t)len,
           (int64_t)strides[ndim - 1], divisor, v, (int64_t)offset);
  DFB_LAUNCH_CHECK("ReduceSumView");
  return DFB_OK;
}

dfb_status dfb_ewise_setitem(const float* a, size_t a_size, float* out, int ndim, const int32_t* shape,
                             const int32_t* strides, size_t offset) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "EwiseSetitem: out array cannot be null");
  DFB_REQUIRE(a_size != 0, DFB_ERR_INVALID, "EwiseSetitem: a array size cannot be zero");
  dfb_status st = check_view("EwiseSetitem", ndim, shape);
  if (st != DFB_OK) return st;
  StridedView v = collapse(ndim, shape, strides);
  cudaStream_t s = compute_stream();
  if (v.ndim == 1 && v.stride[0] == 1) {
    DFB_CUDA(cudaMemcpyAsync(out + offset, a, a_size * sizeof(float), cudaMemcpyDeviceToDevice, s));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DFB_OK;
  }
  size_t view_size = 1;
  for (int d = 0; d < v.ndim; ++d) view_size *= v.shape[d];
  if (view_size == a_size && try_transpose<1>(a, out, v, (int64_t)offset, s)) {
    DFB_LAUNCH_CHECK("EwiseSetitem(transpose)");
    return DFB_OK;
  }
  launch_k(strided_kernel<1>, bw_grid(a_size, kThreads), kThreads, 0, s, a, out, 0.f, a_size, v, (int64_t)offset);
  DFB_LAUNCH_CHECK("EwiseSetitem");
  return DFB_OK;
}

dfb_status dfb_scalar_setitem(size_t size, float value, float* out, size_t out_size, int ndim,
                              const int32_t* shape, const int32_t* strides, size_t offset) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "ScalarSetitem: out array cannot be null");
  DFB_REQUIRE(size != 0, DFB_ERR_INVALID, "ScalarSetitem: size cannot be zero");
  DFB_REQUIRE(size <= out_size, DFB_ERR_OUT_OF_RANGE, "ScalarSetitem: size exceeds out array size");
  dfb_status st = check_view("ScalarSetitem", ndim, shape);
  if (st != DFB_OK) return st;
  StridedView v = collapse(ndim, shape, strides);
  launch_k(strided_kernel<2>, bw_grid(size, kThreads), kThreads, 0, compute_stream(), nullptr, out, value, size, v,
                                                                               (int64_t)offset);
  DFB_LAUNCH_CHECK("ScalarSetitem");
  return DFB_OK;
}

dfb_status dfb_ewise_add(const float* a, const float* b, float* out, size_t n) { return run_binary("EwiseAdd", a, b, out, n, OpAdd()); }
dfb_status dfb_ewise_mul(const float* a, const float* b, float* out, size_t n) { return run_binary("EwiseMul", a, b, out, n, OpMul()); }
dfb_status dfb_ewise_div(const float* a, const float* b, float* out, size_t n) { return run_binary("EwiseDiv", a, b, out, n, OpDiv()); }
dfb_status dfb_ewise_maximum(const float* a, const float* b, float* out, size_t n) { return run_binary("EwiseMaximum", a, b, out, n, OpMax()); }
dfb_status dfb_ewise_eq(const float* a, const float* b, float* out, size_t n) { return run_binary("EwiseEq", a, b, out, n, OpEq()); }
dfb_status dfb_ewise_ge(const float* a, const float* b, float* out, size_t n) { return run_binary("EwiseGe", a, b, out, n, OpGe()); }

dfb_status dfb_scalar_add(const float* a, float v, float* out, size_t n) { return run_scalar("ScalarAdd", a, v, out, n, OpAdd()); }
dfb_status dfb_scalar_mul(const float* a, float v, float* out, size_t n) { return run_scalar("ScalarMul", a, v, out, n, OpMul()); }
dfb_status dfb_scalar_div(const float* a, float v, float* out, size_t n) {
  DFB_REQUIRE(v != 0.0f, DFB_ERR_DOMAIN, "ScalarDiv: division by zero");
  return run_scalar("ScalarDiv", a, v, out, n, OpDiv());
}
dfb_status dfb_scalar_power(const float* a, float v, float* out, size_t n) { return run_scalar("ScalarPower", a, v, out, n, OpPow()); }
dfb_status dfb_scalar_maximum(const float* a, float v, float* out, size_t n) { return run_scalar("ScalarMaximum", a, v, out, n, OpMax()); }
dfb_status dfb_scalar_eq(const float* a, float v, float* out, size_t n) { return run_scalar("ScalarEq", a, v, out, n, OpEq()); }
dfb_status dfb_scalar_ge(const float* a, float v, float* out, size_t n) { return run_scalar("ScalarGe", a, v, out, n, OpGe()); }

dfb_status dfb_ewise_log(const float* a, float* out, size_t n) { return run_unary("EwiseLog", a, out, n, OpLog()); }
dfb_status dfb_ewise_exp(const float* a, float* out, size_t n) { return run_unary("EwiseExp", a, out, n, OpExp()); }
dfb_status dfb_ewise_tanh(const float* a, float* out, size_t n) { return run_unary("EwiseTanh", a, out, n, OpTanh()); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// reductions over the trailing contiguous axis
// ---------------------------------------------------------------------------------------------
namespace dfb {

struct RSum {
  static __device__ float init() { return 0.0f; }
  static __device__ float op(float a, float b) { return a + b; }
  static __device__ float warp(float v) { return warp_sum(v); }
};
struct RMax {
  static __device__ float init() { return -INFINITY; }
  static __device__ float op(float a, float b) { return fmaxf(a, b); }
  static __device__ float warp(float v) { return warp_max(v); }
};

// short rows: one thread per row (rows are <= 32 floats: pooling windows, class scores)
template <class R>
__global__ void __launch_bounds__(kThreads) reduce_thread_kernel(const float* __restrict__ a,
                                                                 float* __restrict__ out, size_t rows,
                                                                 uint32_t len) {
  pdl_sync();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const float* p = a + r * len;
    float acc = p[0];
    for (uint32_t i = 1; i < len; ++i) acc = R::op(acc, p[i]);
    out[r] = acc;
  }
}
// medium rows: one warp per row
template <class R>
__global__ void __launch_bounds__(kThreads) reduce_warp_kernel(const float* __restrict__ a,
                                                               float* __restrict__ out, size_t rows,
                                                               size_t len) {
  pdl_sync();
  int lane = threadIdx.x & 31;
  size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp; r < rows; r += nwarps) {
    const float* p = a + r * len;
    float acc = R::init();
    for (size_t i = lane; i < len; i += 32) acc = R::op(acc, p[i]);
    acc = R::warp(acc);
    if (lane == 0) out[r] = acc;
  }
}
// long rows: one CTA per row (grid-stride over rows), float4 loads when aligned
template <class R>
__global__ void __launch_bounds__(kThreads) reduce_block_kernel(const float* __restrict__ a,
                                                                float* __restrict__ out, size_t rows,
                                                                size_t len) {
  pdl_sync();
  __shared__ float partial[kThreads / 32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (size_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* p = a + r * len;
    float acc = R::init();
    bool vec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
    size_t n4 = vec ? len / 4 : 0;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (size_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 v = p4[i];
      acc = R::op(acc, R::op(R::op(v.x, v.y), R::op(v.z, v.w)));
    }
    for (size_t i = n4 * 4 + threadIdx.x; i < len; i += blockDim.x) acc = R::op(acc, p[i]);
    acc = R::warp(acc);
    if (lane == 0) partial[wid] = acc;
    __syncthreads();
    if (wid == 0) {
      float v = lane < kThreads / 32 ? partial[lane] : R::init();
      v = R::warp(v);
      if (lane == 0) out[r] = v;
    }
    __syncthreads();
  }
}
// one huge row (sum/max over a whole tensor): two-stage, deterministic
template <class R>
__global__ void __launch_bounds__(kThreads) reduce_split_kernel(const float* __restrict__ a,
                                                                float* __restrict__ partials, size_t len) {
  pdl_sync();
  __shared__ float partial[kThreads / 32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float acc = R::init();
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) acc = R::op(acc, a[i]);
  acc = R::warp(acc);
  if (lane == 0) partial[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    float v = lane < kThreads / 32 ? partial[lane] : R::init();
    v = R::warp(v);
    if (lane == 0) partials[blockIdx.x] = v;
  }
}

template <class R>
dfb_status run_reduce(const char* name, const float* a, float* out, size_t out_size, size_t reduce_size) {
  DFB_INIT();
  DFB_REQUIRE(out != nullptr, DFB_ERR_INVALID, "%s: out array cannot be null", name);
  DFB_REQUIRE(reduce_size != 0, DFB_ERR_INVALID, "%s: reduce_size cannot be zero", name);
  if (out_size == 0) return DFB_OK;
  cudaStream_t s = compute_stream();
  if (reduce_size <= 32) {
    launch_k(reduce_thread_kernel<R>, bw_grid(out_size, kThreads), kThreads, 0, s, a, out, out_size, (uint32_t)reduce_size);
  } else if (out_size == 1 && reduce_size >= (size_t)1 << 16) {
    unsigned blocks = bw_grid(reduce_size, kThreads, 4);
    float* partials = nullptr;
    dfb_status st = dfb_malloc(blocks, &partials);
    if (st != DFB_OK) return st;
    launch_k(reduce_split_kernel<R>, blocks, kThreads, 0, s, a, partials, reduce_size);
    DFB_LAUNCH_CHECK(name);
    launch_k(reduce_block_kernel<R>, 1, kThreads, 0, s, partials, out, 1, blocks);
    dfb_free(partials);
  } else if (reduce_size < 1024) {
    launch_k(reduce_warp_kernel<R>, bw_grid(out_size * 32, kThreads), kThreads, 0, s, a, out, out_size, reduce_size);
  } else {
    unsigned grid = (unsigned)std::min<size_t>(out_size, (size_t)sm_count() * 8);
    launch_k(reduce_block_kernel<R>, grid, kThreads, 0, s, a, out, out_size, reduce_size);
  }
  DFB_LAUNCH_CHECK(name);
  return DFB_OK;
}
}  // namespace dfb

extern "C" {
dfb_status dfb_reduce_sum(const float* a, float* out, size_t out_size, size_t reduce_size) {
  return dfb::run_reduce<dfb::RSum>("ReduceSum", a, out, out_size, reduce_size);
}
dfb_status dfb_reduce_max(const float* a, float* out, size_t out_size, size_t reduce_size) {
  return dfb::run_reduce<dfb::RMax>("ReduceMax", a, out, out_size, reduce_size);
}
}
