// Fused elementwise tails of the decoder's convolutions, channels-last (NHWC) tensors, bf16 or fp32.
//
// The STCN decoder (prop_net.py:13-30, modules.py: ResBlock / UpsampleBlock) has no normalisation layers, so between
// its convolutions sit only bias adds, residual adds, ReLUs and two bilinear x2 upsamplings - on 5 frames of a 480p
// video each is a full pass over 33 M elements at stride 4, and PyTorch runs them one kernel per op (broadcast bias
// add 70 us, add 170 us, upsample 275 us ... ~1 ms per segment, 15 % of a video's device time once the encoders are
// folded).  Two kernels cover them:
//   bias_residual:   y = [relu](y + bias[c] (+ r))                       - the tail of a ResBlock's second convolution
//   upsample2x_add:  y = y + bias[c] + bilinear_up2x(x)                  - UpsampleBlock: skip_conv(skip) + up
// Both are HBM-bound: 16-byte accesses along the channel axis, one read and one write of y (+ one read of r, a quarter
// of y for x).  Bilinear weights as ATen's upsample_bilinear2d with align_corners = False, scale 2: source coordinate
// (d + 0.5) / 2 - 0.5 clamped at 0, neighbours clamped at the edge; fp32 arithmetic, rounded once.
// Plain launches (no PDL): they follow cuDNN convolutions, mostly inside captured graphs - nothing to overlap with.
#include "common.cuh"

namespace evavos {

namespace {

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float* v) { const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
  __device__ static void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float* v) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
  __device__ static void store(__nv_bfloat16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&t); }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// y: (rows, C) channel-contiguous; bias fp32 (C); r: like y or NULL.
template <typename T>
__global__ void __launch_bounds__(256) bias_residual_kernel(T* __restrict__ y, const float* __restrict__ bias,
                                                            const T* __restrict__ r, int64_t n_vec, int c_vecs, int relu) {
  constexpr int N = Vec<T>::N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c_vecs) * N;
    float v[N], b[N];
    Vec<T>::load(y + i * N, v);
#pragma unroll
    for (int j = 0; j < N; j += 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
      b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
    }
    if (r != nullptr) {
      float rv[N];
      Vec<T>::load(r + i * N, rv);
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = (v[j] + b[j]) + rv[j];
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = v[j] + b[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    Vec<T>::store(y + i * N, v);
  }
}

// y: (n, H, W, C); x: (n, H/2, W/2, C); y += bias + up2x(x).  grid (n * H, splits): a CTA owns a slice of ONE output row,
// so the row's two source rows and vertical weights are CTA constants and a vector index needs one 32-bit division.
// (A flat grid-stride loop that takes a 64-bit vector index apart - five emulated divisions - ran at 0.61 of the HBM
// roofline with the issue slots 65 % busy; 32-bit indices: 0.70.)
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_add_kernel(T* __restrict__ y, const float* __restrict__ bias,
                                                             const T* __restrict__ x, int c_vecs, int H, int W) {
  constexpr int N = Vec<T>::N;
  const int h2 = H >> 1, w2 = W >> 1;
  const int row = blockIdx.x, n = row / H, ho = row - n * H;
  // ATen: src = max((dst + 0.5) * 0.5 - 0.5, 0); i0 = floor(src); i1 = min(i0 + 1, size - 1); l1 = src - i0; l0 = 1 - l1
  const float sh = fmaxf((ho + 0.5f) * 0.5f - 0.5f, 0.f);
  const int h0 = (int)sh, h1 = min(h0 + 1, h2 - 1);
  const float lh1 = sh - (float)h0, lh0 = 1.f - lh1;
  const int wa = (int)(((int64_t)W * blockIdx.y) / gridDim.y), wb = (int)(((int64_t)W * (blockIdx.y + 1)) / gridDim.y);
  const int64_t src_row = (int64_t)w2 * c_vecs * N;
  const T* x0 = x + ((int64_t)n * h2 + h0) * src_row;
  const T* x1 = x + ((int64_t)n * h2 + h1) * src_row;
  T* yr = y + (int64_t)row * W * c_vecs * N;
  const int n_items = (wb - wa) * c_vecs;
  for (int e = threadIdx.x; e < n_items; e += 256) {
    const int wq = e / c_vecs, co = (e - wq * c_vecs) * N, wo = wa + wq;
    const float sw = fmaxf((wo + 0.5f) * 0.5f - 0.5f, 0.f);
    const int w0 = (int)sw, w1 = min(w0 + 1, w2 - 1);
    const float lw1 = sw - (float)w0, lw0 = 1.f - lw1;
    const int o0 = w0 * c_vecs * N + co, o1 = w1 * c_vecs * N + co;
    T* yp = yr + (int64_t)wo * c_vecs * N + co;
    float a00[N], a01[N], a10[N], a11[N], v[N], b[N];
    Vec<T>::load(x0 + o0, a00);
    Vec<T>::load(x0 + o1, a01);
    Vec<T>::load(x1 + o0, a10);
    Vec<T>::load(x1 + o1, a11);
    Vec<T>::load(yp, v);
#pragma unroll
    for (int j = 0; j < N; j += 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(bias + co + j));
      b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float up = lh0 * (lw0 * a00[j] + lw1 * a01[j]) + lh1 * (lw0 * a10[j] + lw1 * a11[j]);
      v[j] = (v[j] + b[j]) + up;
    }
    Vec<T>::store(yp, v);
  }
}

int grid_for(int64_t n_vec) {
  int64_t g = ceil_div(n_vec, 256);
  if (g > 148 * 16) g = 148 * 16;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace

int launch_bias_residual(void* y, const float* bias, const void* r, int64_t rows, int C, int bf16, int relu, cudaStream_t st) {
  if (rows <= 0) return EVAVOS_OK;
  if (bf16) {
    const int cv = C / 8;
    bias_residual_kernel<__nv_bfloat16><<<grid_for(rows * cv), 256, 0, st>>>(
        reinterpret_cast<__nv_bfloat16*>(y), bias, reinterpret_cast<const __nv_bfloat16*>(r), rows * cv, cv, relu);
  } else {
    const int cv = C / 4;
    bias_residual_kernel<float><<<grid_for(rows * cv), 256, 0, st>>>(reinterpret_cast<float*>(y), bias,
                                                                     reinterpret_cast<const float*>(r), rows * cv, cv, relu);
  }
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

template <typename T>
static void launch_up(T* y, const float* bias, const T* x, int64_t n, int cv, int H, int W, cudaStream_t st) {
  const int64_t rows = n * H;
  int64_t splits = ceil_div(148 * 32, rows);    // ~4 waves of 8 CTAs per SM (two CTAs more than one wave cost a whole second one)
  if (splits > W) splits = W;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  upsample2x_add_kernel<T><<<dim3((unsigned)rows, (unsigned)splits), 256, 0, st>>>(y, bias, x, cv, H, W);
}

int launch_upsample2x_add(void* y, const float* bias, const void* x, int64_t n, int H, int W, int C, int bf16, cudaStream_t st) {
  if (n <= 0) return EVAVOS_OK;
  if (n * H > 0x7fffffffll || (int64_t)W * C > 0x3fffffffll) return EVAVOS_ERR_INVALID;
  if (bf16) launch_up(reinterpret_cast<__nv_bfloat16*>(y), bias, reinterpret_cast<const __nv_bfloat16*>(x), n, C / 8, H, W, st);
  else launch_up(reinterpret_cast<float*>(y), bias, reinterpret_cast<const float*>(x), n, C / 4, H, W, st);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

}  // namespace evavos
