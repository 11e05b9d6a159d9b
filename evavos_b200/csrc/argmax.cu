// Hard masks of all frames in one pass.
//
// InferenceCore.interact ends with one torch.argmax launch per frame, a strided un-padding slice and a D2H copy
// (mivos/inference_core.py:247-257).  Here every padded pixel of every frame is read once: the channel argmax
// (first maximal channel, NaN counting as the largest value, like torch.argmax) goes to `masks` (T, nh, nw) and, for pixels inside the original
// frame, to the contiguous un-padded `out` (T, h, w) that is copied to the host.  HBM-bound:
// algorithmic bytes = C * T * nh * nw * 4 read + T * (nh * nw + h * w) written.
#include "common.cuh"

namespace evavos {

namespace {

template <int VEC>
__global__ void __launch_bounds__(256) argmax_unpad_kernel(const float* __restrict__ prob, int C, int64_t T, int nh,
                                                           int nw, uint8_t* __restrict__ masks,
                                                           uint8_t* __restrict__ out, int pad_top, int pad_left, int h,
                                                           int w) {
  const int64_t frame_px = (int64_t)nh * nw;
  const int64_t total = T * frame_px / VEC;
  const int64_t ch_stride = T * frame_px;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = i * VEC;
    float best[VEC];
    uint8_t arg[VEC];
    if constexpr (VEC == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(prob + e0));
      best[0] = v.x; best[1] = v.y; best[2] = v.z; best[3] = v.w;
    } else {
      best[0] = __ldg(prob + e0);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) arg[j] = 0;
    for (int c = 1; c < C; ++c) {
      float cur[VEC];
      if constexpr (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * ch_stride + e0));
        cur[0] = v.x; cur[1] = v.y; cur[2] = v.z; cur[3] = v.w;
      } else {
        cur[0] = __ldg(prob + (int64_t)c * ch_stride + e0);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j)
        // torch.argmax orders NaN above everything and keeps the first one
        if (cur[j] > best[j] || (cur[j] != cur[j] && best[j] == best[j])) { best[j] = cur[j]; arg[j] = (uint8_t)c; }
    }
    if (masks) {
      if constexpr (VEC == 4) *reinterpret_cast<uchar4*>(masks + e0) = make_uchar4(arg[0], arg[1], arg[2], arg[3]);
      else masks[e0] = arg[0];
    }
    if (out) {
      const int64_t t = e0 / frame_px;
      const int64_t r = e0 - t * frame_px;
      const int y = (int)(r / nw) - pad_top, x0 = (int)(r % nw) - pad_left;  // VEC pixels share a row (nw % VEC == 0)
      if (y >= 0 && y < h) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const int x = x0 + j;
          if (x >= 0 && x < w) out[(t * h + y) * (int64_t)w + x] = arg[j];
        }
      }
    }
  }
}

}  // namespace

int launch_argmax_unpad(const float* prob, int C, int64_t T, int nh, int nw, uint8_t* masks, uint8_t* out,
                        int pad_top, int pad_left, int h, int w, cudaStream_t st) {
  const int64_t px = T * (int64_t)nh * nw;
  if (px <= 0) return EVAVOS_OK;
  const bool vec = (nw % 4 == 0) && (reinterpret_cast<uintptr_t>(prob) % 16 == 0) &&
                   (masks == nullptr || reinterpret_cast<uintptr_t>(masks) % 4 == 0);
  const int64_t items = vec ? px / 4 : px;
  int64_t grid = ceil_div(items, 256);
  if (grid > 148 * 16) grid = 148 * 16;  // grid-stride; a multiple of the SM count
  if (vec)
    argmax_unpad_kernel<4><<<(unsigned)grid, 256, 0, st>>>(prob, C, T, nh, nw, masks, out, pad_top, pad_left, h, w);
  else
    argmax_unpad_kernel<1><<<(unsigned)grid, 256, 0, st>>>(prob, C, T, nh, nw, masks, out, pad_top, pad_left, h, w);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

}  // namespace evavos
