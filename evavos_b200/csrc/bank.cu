// Memory-bank writes: reference layout + engine-private position-major shadow.
//
// Replaces the strided slice-assign append of InferenceCore.do_pass
// (mivos/inference_core.py:174-177), the certain-memory seed copy (:154-155) and the
// torch.cat growth in interact (:235-240).  One pass over the new frame writes
//   - the reference-layout bank (1,CK,T,H,W)/(K,CV,T,H,W) (optional),
//   - key_pm  [pos][CK] fp32 rows for exact rescoring,
//   - key tile images (bf16, 128B-swizzled K-major + a bf16-split -|k|^2/2 K slice) for tcgen05.mma,
//   - val_pm  [K][pos][CV] rows for the coalesced sparse readout.
// Both kernels are HBM-bound transposes: algorithmic bytes = 2 * (CK + K*CV) * n_pos * 4
// (read once, write reference layout + shadow).
#include "common.cuh"

namespace evavos {

namespace {

constexpr int kKeyBlockPos = 64;

// (hi, mid, lo, 0, 0, 0, 0, 0) bf16: hi + mid + lo == x to 24 bits.
__device__ __forceinline__ uint4 nh_slice(float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  uint4 w;
  w.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
  w.y = (uint32_t)__bfloat16_as_ushort(lo);
  w.z = 0;
  w.w = 0;
  return w;
}

// 256 threads, 64 positions per CTA, 4 threads per position.
__global__ void __launch_bounds__(256) write_keys_kernel(
    const float* __restrict__ src, int64_t src_ch_stride, int64_t pos0, int64_t n_pos, int CK,
    float* __restrict__ dst_ref, int64_t dst_ref_ch_stride, float* __restrict__ key_pm,
    uint8_t* __restrict__ tiles, float* __restrict__ maxnorm) {
  pdl_wait();  // the bank may still be read by the kernel before this one
  __shared__ float sm[64][kKeyBlockPos + 1];
  __shared__ float warp_max[8];
  const int tid = threadIdx.x;
  const int64_t blk0 = (int64_t)blockIdx.x * kKeyBlockPos;  // relative to pos0
  const int nb = (int)min((int64_t)kKeyBlockPos, n_pos - blk0);

  for (int e = tid; e < CK * kKeyBlockPos; e += 256) {
    const int c = e / kKeyBlockPos, i = e % kKeyBlockPos;
    float v = 0.f;
    if (i < nb) {
      v = src[(int64_t)c * src_ch_stride + blk0 + i];
      if (dst_ref) dst_ref[(int64_t)c * dst_ref_ch_stride + pos0 + blk0 + i] = v;
    }
    sm[c][i] = v;
  }
  __syncthreads();

  const int i = tid >> 2, part = tid & 3;
  const int cpp = CK >> 2;  // channels per thread (CK % 8 == 0)
  const int c0 = part * cpp;
  const int64_t P = pos0 + blk0 + i;
  float ss = 0.f;
  if (i < nb) {
    float* row = key_pm + P * CK + c0;
    for (int c = 0; c < cpp; ++c) {
      const float v = sm[c0 + c][i];
      ss = fmaf(v, v, ss);
      row[c] = v;
    }
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);

  if (tiles != nullptr && i < nb) {  // CK == 64: cpp == 16 -> two 16-byte chunks per thread
    const int64_t tile = P / kTilePos;
    const int r = (int)(P % kTilePos);
    uint8_t* tb = tiles + tile * (int64_t)kTileBytes;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int chunk = part * 2 + h;
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 p2 = __floats2bfloat162_rn(sm[chunk * 8 + 2 * j][i], sm[chunk * 8 + 2 * j + 1][i]);
        w[j] = *reinterpret_cast<uint32_t*>(&p2);
      }
      *reinterpret_cast<uint4*>(tb + swizzle128_offset(r, chunk)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (part < 2) {
      // extra K slice: -|k|^2/2 split into three bf16 terms (24 mantissa bits), then zeros
      uint4 aug = make_uint4(0, 0, 0, 0);
      if (part == 0) aug = nh_slice(-0.5f * ss);
      *reinterpret_cast<uint4*>(tb + kTileKeyBytes + swizzle32_offset(r, part)) = aug;
    }
  }

  // Rows of a tile beyond the written range are NOT touched: a frame may be rewritten in the middle of a bank
  // (MemoryBank.write_frames accepts any slot), and the filter masks columns >= n_pos itself, so stale or
  // never-written rows of the last tile can not reach a candidate list.

  if (maxnorm != nullptr) {
    float nrm = (i < nb) ? sqrtf(ss) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm = fmaxf(nrm, __shfl_xor_sync(0xffffffffu, nrm, o));
    if ((tid & 31) == 0) warp_max[tid >> 5] = nrm;
    __syncthreads();
    if (tid == 0) {
      float m = warp_max[0];
      for (int w = 1; w < 8; ++w) m = fmaxf(m, warp_max[w]);
      atomicMax(reinterpret_cast<unsigned int*>(maxnorm), __float_as_uint(m));  // m >= 0
    }
  }
}

// 32x32 transpose tiles; grid (pos blocks, channel blocks, K), block (32, 8).
template <typename OutT>
__global__ void __launch_bounds__(256) write_values_kernel(
    const float* __restrict__ src, int64_t src_obj_stride, int64_t src_ch_stride, int64_t pos0, int64_t n_pos,
    int CV, float* __restrict__ dst_ref, int64_t dst_ref_obj_stride, int64_t dst_ref_ch_stride,
    OutT* __restrict__ val_pm, int64_t capacity_pos) {
  pdl_wait();  // the bank may still be read by the kernel before this one
  __shared__ float t[32][33];
  const int o = blockIdx.z;
  const int64_t p_blk = (int64_t)blockIdx.x * 32;
  const int c_blk = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int cy = ty; cy < 32; cy += 8) {
    const int c = c_blk + cy;
    const int64_t p = p_blk + tx;
    float v = 0.f;
    if (c < CV && p < n_pos) {
      v = src[(int64_t)o * src_obj_stride + (int64_t)c * src_ch_stride + p];
      if (dst_ref) dst_ref[(int64_t)o * dst_ref_obj_stride + (int64_t)c * dst_ref_ch_stride + pos0 + p] = v;
    }
    t[cy][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int py = ty; py < 32; py += 8) {
    const int64_t p = p_blk + py;
    const int c = c_blk + tx;
    if (c < CV && p < n_pos) {
      const float v = t[tx][py];
      OutT* dst = val_pm + ((int64_t)o * capacity_pos + pos0 + p) * CV + c;
      if constexpr (sizeof(OutT) == 2) *dst = __float2bfloat16_rn(v);
      else *dst = v;
    }
  }
}

// CV % 128 == 0: a CTA moves 32 positions x 128 channels of one object.  Each thread has 16 coalesced 4-byte loads in
// flight (a warp reads 32 consecutive positions of one channel = one 128-byte line, and writes the same line of the
// reference layout), the transposed rows leave as 16-byte stores: one warp instruction writes the 128 channels of one
// position (512 contiguous bytes; 256 for bf16).  r1's 32 x 32 tiles with 4-byte stores moved 30 MB in 13 us.
template <typename OutT>
__global__ void __launch_bounds__(256) write_values_wide_kernel(
    const float* __restrict__ src, int64_t src_obj_stride, int64_t src_ch_stride, int64_t pos0, int64_t n_pos,
    int CV, float* __restrict__ dst_ref, int64_t dst_ref_obj_stride, int64_t dst_ref_ch_stride,
    OutT* __restrict__ val_pm, int64_t capacity_pos) {
  pdl_wait();  // the bank may still be read by the kernel before this one
  __shared__ __align__(16) float t[32][132];   // [position][channel], rows padded to keep 16-byte alignment
  const int o = blockIdx.z;
  const int64_t p_blk = (int64_t)blockIdx.x * 32;
  const int c_blk = blockIdx.y * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = p_blk + lane;
  float v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = c_blk + warp + 8 * u;
    v[u] = (p < n_pos) ? __ldg(src + (int64_t)o * src_obj_stride + (int64_t)c * src_ch_stride + p) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int c = c_blk + warp + 8 * u;
    if (dst_ref != nullptr && p < n_pos)
      dst_ref[(int64_t)o * dst_ref_obj_stride + (int64_t)c * dst_ref_ch_stride + pos0 + p] = v[u];
    t[lane][warp + 8 * u] = v[u];
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int pr = warp + 8 * u;                       // position row of the tile
    const int64_t pp = p_blk + pr;
    if (pp >= n_pos) continue;
    const float4 q = *reinterpret_cast<const float4*>(&t[pr][4 * lane]);
    OutT* dst = val_pm + ((int64_t)o * capacity_pos + pos0 + pp) * CV + c_blk + 4 * lane;
    if constexpr (sizeof(OutT) == 2) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(q.x, q.y), b = __floats2bfloat162_rn(q.z, q.w);
      *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    } else {
      *reinterpret_cast<float4*>(dst) = q;
    }
  }
}

}  // namespace

int launch_write_keys(const EvavosBankShadow& b, const float* src, int64_t src_ch_stride, int64_t pos0,
                      int64_t n_pos, float* dst_ref, int64_t dst_ref_ch_stride, cudaStream_t st) {
  if (n_pos <= 0) return EVAVOS_OK;
  const int grid = (int)ceil_div(n_pos, kKeyBlockPos);
  uint8_t* tiles = (b.CK == 64) ? reinterpret_cast<uint8_t*>(b.key_tiles) : nullptr;
  EVAVOS_CUDA_OK(launch_pdl(write_keys_kernel, dim3((unsigned)grid), dim3(256), 0, st, src, src_ch_stride, pos0, n_pos,
                            b.CK, dst_ref, dst_ref_ch_stride, b.key_pm, tiles, b.key_maxnorm));
  return EVAVOS_OK;
}

int launch_write_values(const EvavosBankShadow& b, const float* src, int64_t src_obj_stride,
                        int64_t src_ch_stride, int64_t pos0, int64_t n_pos, float* dst_ref,
                        int64_t dst_ref_obj_stride, int64_t dst_ref_ch_stride, cudaStream_t st) {
  if (n_pos <= 0) return EVAVOS_OK;
  if (b.CV % 128 == 0 && reinterpret_cast<uintptr_t>(b.val_pm) % 16 == 0) {
    dim3 wgrid((unsigned)ceil_div(n_pos, 32), (unsigned)(b.CV / 128), (unsigned)b.K);
    if (b.val_dtype == EVAVOS_BF16)
      EVAVOS_CUDA_OK(launch_pdl(write_values_wide_kernel<__nv_bfloat16>, wgrid, dim3(256), 0, st, src, src_obj_stride,
                                src_ch_stride, pos0, n_pos, b.CV, dst_ref, dst_ref_obj_stride, dst_ref_ch_stride,
                                reinterpret_cast<__nv_bfloat16*>(b.val_pm), b.capacity_pos));
    else
      EVAVOS_CUDA_OK(launch_pdl(write_values_wide_kernel<float>, wgrid, dim3(256), 0, st, src, src_obj_stride,
                                src_ch_stride, pos0, n_pos, b.CV, dst_ref, dst_ref_obj_stride, dst_ref_ch_stride,
                                reinterpret_cast<float*>(b.val_pm), b.capacity_pos));
    return EVAVOS_OK;
  }
  dim3 grid((unsigned)ceil_div(n_pos, 32), (unsigned)ceil_div(b.CV, 32), (unsigned)b.K);
  dim3 block(32, 8);
  if (b.val_dtype == EVAVOS_BF16) {
    EVAVOS_CUDA_OK(launch_pdl(write_values_kernel<__nv_bfloat16>, grid, block, 0, st, src, src_obj_stride, src_ch_stride,
                              pos0, n_pos, b.CV, dst_ref, dst_ref_obj_stride, dst_ref_ch_stride,
                              reinterpret_cast<__nv_bfloat16*>(b.val_pm), b.capacity_pos));
  } else {
    EVAVOS_CUDA_OK(launch_pdl(write_values_kernel<float>, grid, block, 0, st, src, src_obj_stride, src_ch_stride, pos0,
                              n_pos, b.CV, dst_ref, dst_ref_obj_stride, dst_ref_ch_stride,
                              reinterpret_cast<float*>(b.val_pm), b.capacity_pos));
  }
  return EVAVOS_OK;
}

}  // namespace evavos
