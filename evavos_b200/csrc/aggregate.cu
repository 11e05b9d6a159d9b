// Soft aggregation across objects: one fused, vectorised pass (aggregate_wbg, aggregate.py:22-37).
//
// The reference issues prod / cat / clamp / div / log / (mul) / softmax as separate full-tensor
// passes; here every pixel is read once and written once.  HBM-bound: algorithmic bytes
// = (2K + 1) * npix * 4 (K inputs, K+1 outputs when keep_bg).
#include "common.cuh"

namespace evavos {

namespace {

__device__ __forceinline__ float logit_of(float p, float scale) {
  const float lo = (float)1e-7;
  const float hi = (float)(1.0 - 1e-7);
  p = fminf(fmaxf(p, lo), hi);                 // .clamp(1e-7, 1-1e-7)
  return logf(p / (1.0f - p)) * scale;         // log(p / (1 - p)), * 1000 when hard
}

// softmax_k(log(p_k / (1 - p_k))) = odds_k / sum_j odds_j with odds = p / (1 - p) of the clamped probability: the soft
// (not `hard`) aggregation needs no log and no exp.  odds <= (1 - 1e-7) / 1e-7 ~ 1e7, so the sum of <= 129 of them is
// far from overflow; against the reference's log -> max-subtract -> exp -> normalise in fp32 the result differs by
// <= 1.8e-7 (its own rounding; checked on the golden vectors), inside the 1e-6 parity gate.
__device__ __forceinline__ float odds_of(float p) {
  const float lo = (float)1e-7;
  const float hi = (float)(1.0 - 1e-7);
  p = fminf(fmaxf(p, lo), hi);
  return p / (1.0f - p);
}

// K known at compile time: probabilities of one pixel group live in registers.
template <int K, int VEC>
__global__ void __launch_bounds__(256) aggregate_kernel(const float* __restrict__ prob, float* __restrict__ out,
                                                        int64_t npix, int keep_bg, float scale) {
  const int64_t nvec = npix / VEC;
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float p[K][VEC];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)k * npix) + i);
        p[k][0] = v.x; p[k][1] = v.y; p[k][2] = v.z; p[k][3] = v.w;
      } else {
        p[k][0] = __ldg(prob + (int64_t)k * npix + i);
      }
    }
    float res[K + 1][VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      float bg = 1.0f - p[0][e];
#pragma unroll
      for (int k = 1; k < K; ++k) bg *= (1.0f - p[k][e]);   // torch.prod(1 - prob, dim=0)
      float l[K + 1];
      if (scale == 1.0f) {   // soft aggregation (every caller on the propagation path): odds / sum of odds
        l[0] = odds_of(bg);
        float sum = l[0];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          l[k + 1] = odds_of(p[k][e]);
          sum += l[k + 1];
        }
#pragma unroll
        for (int k = 0; k <= K; ++k) res[k][e] = l[k] / sum;
        continue;
      }
      l[0] = logit_of(bg, scale);
      float m = l[0];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        l[k + 1] = logit_of(p[k][e], scale);
        m = fmaxf(m, l[k + 1]);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k <= K; ++k) {
        l[k] = expf(l[k] - m);
        sum += l[k];
      }
#pragma unroll
      for (int k = 0; k <= K; ++k) res[k][e] = l[k] / sum;
    }
    const int first = keep_bg ? 0 : 1;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
      if (k < first) continue;
      float* dst = out + (int64_t)(k - first) * npix;
      if constexpr (VEC == 4)
        reinterpret_cast<float4*>(dst)[i] = make_float4(res[k][0], res[k][1], res[k][2], res[k][3]);
      else
        dst[i] = res[k][0];
    }
  }
}

// Any K: three passes over the K inputs of a pixel (they stay in L1).
__global__ void __launch_bounds__(256) aggregate_generic_kernel(const float* __restrict__ prob,
                                                                float* __restrict__ out, int K, int64_t npix,
                                                                int keep_bg, float scale) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix;
       i += (int64_t)gridDim.x * blockDim.x) {
    float bg = 1.0f;
    for (int k = 0; k < K; ++k) {
      const float c = 1.0f - prob[(int64_t)k * npix + i];
      bg = (k == 0) ? c : bg * c;
    }
    const float l0 = logit_of(bg, scale);
    float m = l0;
    for (int k = 0; k < K; ++k) m = fmaxf(m, logit_of(prob[(int64_t)k * npix + i], scale));
    float sum = expf(l0 - m);
    for (int k = 0; k < K; ++k) sum += expf(logit_of(prob[(int64_t)k * npix + i], scale) - m);
    const int first = keep_bg ? 0 : 1;
    if (keep_bg) out[i] = expf(l0 - m) / sum;
    for (int k = 0; k < K; ++k)
      out[(int64_t)(k + 1 - first) * npix + i] = expf(logit_of(prob[(int64_t)k * npix + i], scale) - m) / sum;
  }
}

template <int K>
int launch_k(const float* prob, float* out, int64_t npix, int keep_bg, float scale, cudaStream_t st) {
  const bool vec = (npix % 4 == 0) && (reinterpret_cast<uintptr_t>(prob) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const int64_t items = vec ? npix / 4 : npix;
  int64_t grid = ceil_div(items, 256);
  if (grid > 148 * 16) grid = 148 * 16;  // grid-stride; a multiple of the SM count
  if (grid < 1) grid = 1;
  if (vec)
    EVAVOS_CUDA_OK(launch_pdl(aggregate_kernel<K, 4>, dim3((unsigned)grid), dim3(256), 0, st, prob, out, npix, keep_bg, scale));
  else
    EVAVOS_CUDA_OK(launch_pdl(aggregate_kernel<K, 1>, dim3((unsigned)grid), dim3(256), 0, st, prob, out, npix, keep_bg, scale));
  return EVAVOS_OK;
}

}  // namespace

int launch_aggregate(const float* prob, float* out, int K, int64_t npix, int keep_bg, int hard,
                     cudaStream_t st) {
  if (npix <= 0) return EVAVOS_OK;
  const float scale = hard ? 1000.0f : 1.0f;
  switch (K) {
    case 1: return launch_k<1>(prob, out, npix, keep_bg, scale, st);
    case 2: return launch_k<2>(prob, out, npix, keep_bg, scale, st);
    case 3: return launch_k<3>(prob, out, npix, keep_bg, scale, st);
    case 4: return launch_k<4>(prob, out, npix, keep_bg, scale, st);
    case 5: return launch_k<5>(prob, out, npix, keep_bg, scale, st);
    case 6: return launch_k<6>(prob, out, npix, keep_bg, scale, st);
    case 7: return launch_k<7>(prob, out, npix, keep_bg, scale, st);
    case 8: return launch_k<8>(prob, out, npix, keep_bg, scale, st);
    default: break;
  }
  int64_t grid = ceil_div(npix, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  EVAVOS_CUDA_OK(launch_pdl(aggregate_generic_kernel, dim3((unsigned)grid), dim3(256), 0, st, prob, out, K, npix, keep_bg, scale));
  return EVAVOS_OK;
}

}  // namespace evavos
