// Sparse value readout and on-demand dense affinity.
//
// readout: out[o][c][q] = sum_j w[q][j] * V[o][idx[q][j]][c]   (prop_net.py:108-115 evaluates this
// as a dense bmm against a matrix with top_k non-zeros per column).  V is the position-major
// value shadow, so each (query, position) pair is one contiguous CV-row: a warp reads it with
// 16-byte loads, fully coalesced.  The kernel is L2/HBM-bound: algorithmic bytes
// = K * CV * min(N, k*HW) * s (every needed row once) + K * CV * HW * 4 (output).
#include "common.cuh"

namespace evavos {

namespace {

#ifndef EVAVOS_READOUT_Q
#define EVAVOS_READOUT_Q 8
#endif
constexpr int kQPerCta = EVAVOS_READOUT_Q;  // one warp per query
constexpr int kRoThreads = 32 * kQPerCta;

// Value rows of one query in flight per lane group.  Measured on B200 (cfg2, cold L2): 2, 5 and 10 give the same
// 37 us, and ld.global.nc.L1::no_allocate is slower (45 us): the kernel sits on the L2/HBM path, not on latency.
constexpr int kRowUnroll = 5;
__device__ __forceinline__ float4 ld_row(const float4* p) { return __ldg(p); }

// Where query q lands inside a (object, channel) row: plain q, or - when several query frames are read in one launch
// and every frame has its own destination block (the decoder input (F, K, 2*CV, H, W)) - frame * frame_stride + position.
__device__ __forceinline__ int64_t query_offset(int64_t q, int q_per_frame, int64_t frame_stride) {
  if (q_per_frame <= 0) return q;
  const uint32_t f = (uint32_t)q / (uint32_t)q_per_frame;
  return (int64_t)f * frame_stride + (int64_t)((uint32_t)q - f * (uint32_t)q_per_frame);
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x);
  a.y = fmaf(w, v.y, a.y);
  a.z = fmaf(w, v.z, a.z);
  a.w = fmaf(w, v.w, a.w);
}

// fp32 rows; each CTA covers 128 * NV channels starting at blockIdx.z * 128 * NV of rows that are CV floats long.
// grid (ceil(nq/8), K, CV / (128 * NV)), 256 threads.
// MODE 0: out[o][c][q] with object / channel strides (the reference layouts), transposed through shared memory.
// MODE 1 (query-major): the output is (n_query, K, CV) - one contiguous K*CV row per query, written straight from the
//         registers (the layout the sharded read reduces over: a query slice is a contiguous chunk); strides ignored.
// MODE 2 (channels-last): out[frame][o][position][c] - an NHWC destination such as the decoder input of a
//         channels_last engine; `out_ch_stride` carries the POSITION stride; also straight from the registers.
template <int NV, int MODE>
__global__ void __launch_bounds__(kRoThreads) readout_f32_kernel(
    const float* __restrict__ val_pm_all, int64_t capacity_pos, int CVfull, const int32_t* __restrict__ idx,
    const float* __restrict__ weight, int64_t n_query, int top_k, float* __restrict__ out_all,
    int64_t out_obj_stride, int64_t out_ch_stride, int q_per_frame, int64_t frame_stride) {
  pdl_wait();  // idx / weight come from the kernel before (programmatic dependent launch)
  pdl_launch_dependents();
  constexpr int CV = 128 * NV;
  __shared__ float st[CV][kQPerCta + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.y;
  const float* val_pm = val_pm_all + (int64_t)blockIdx.z * CV;
  float* out = out_all + (int64_t)blockIdx.z * CV * out_ch_stride;
  const int64_t q0 = (int64_t)blockIdx.x * kQPerCta;
  const int64_t q = q0 + warp;
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q < n_query) {
    const float* vbase = val_pm + (int64_t)o * capacity_pos * CVfull;
    for (int jb = 0; jb < top_k; jb += 32) {
      const int jj = jb + lane;
      const int32_t my_n = jj < top_k ? idx[q * top_k + jj] : -1;
      const float my_w = jj < top_k ? weight[q * top_k + jj] : 0.f;
      const int lim = min(32, top_k - jb);
#pragma unroll kRowUnroll
      for (int j = 0; j < lim; ++j) {
        const int32_t n = __shfl_sync(0xffffffffu, my_n, j);
        const float w = __shfl_sync(0xffffffffu, my_w, j);
        if (n < 0) continue;
        const float4* row = reinterpret_cast<const float4*>(vbase + (int64_t)n * CVfull) + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) fma4(acc[i], w, ld_row(row + 32 * i));
      }
    }
  }
  if constexpr (MODE != 0) {
    if (q < n_query) {
      float* row;
      if constexpr (MODE == 1) {
        row = out_all + ((int64_t)q * gridDim.y + o) * CVfull + (int64_t)blockIdx.z * CV;
      } else {
        const uint32_t f = q_per_frame > 0 ? (uint32_t)q / (uint32_t)q_per_frame : 0u;
        const uint32_t pos = (uint32_t)q - f * (uint32_t)q_per_frame;
        row = out_all + (int64_t)f * frame_stride + (int64_t)o * out_obj_stride + (int64_t)pos * out_ch_stride +
              (int64_t)blockIdx.z * CV;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(row + i * 128 + lane * 4) = acc[i];
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = i * 128 + lane * 4;
    st[c + 0][warp] = acc[i].x;
    st[c + 1][warp] = acc[i].y;
    st[c + 2][warp] = acc[i].z;
    st[c + 3][warp] = acc[i].w;
  }
  __syncthreads();
  const int nq_here = (int)min((int64_t)kQPerCta, n_query - q0);
  for (int e = threadIdx.x; e < CV * kQPerCta; e += kRoThreads) {
    const int c = e / kQPerCta, w = e % kQPerCta;
    if (w < nq_here)
      out[(int64_t)o * out_obj_stride + (int64_t)c * out_ch_stride + query_offset(q0 + w, q_per_frame, frame_stride)] = st[c][w];
  }
}

// bf16 rows, CV = 256 * NV (8 bf16 per 16-byte load), fp32 accumulation.
template <int NV, int MODE>
__global__ void __launch_bounds__(kRoThreads) readout_bf16_kernel(
    const __nv_bfloat16* __restrict__ val_pm, int64_t capacity_pos, const int32_t* __restrict__ idx,
    const float* __restrict__ weight, int64_t n_query, int top_k, float* __restrict__ out,
    int64_t out_obj_stride, int64_t out_ch_stride, int q_per_frame, int64_t frame_stride) {
  pdl_wait();  // idx / weight come from the kernel before (programmatic dependent launch)
  pdl_launch_dependents();
  constexpr int CV = 256 * NV;
  __shared__ float st[CV][kQPerCta + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * kQPerCta;
  const int64_t q = q0 + warp;
  float acc[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
  if (q < n_query) {
    const __nv_bfloat16* vbase = val_pm + (int64_t)o * capacity_pos * CV;
    for (int jb = 0; jb < top_k; jb += 32) {
      const int jj = jb + lane;
      const int32_t my_n = jj < top_k ? idx[q * top_k + jj] : -1;
      const float my_w = jj < top_k ? weight[q * top_k + jj] : 0.f;
      const int lim = min(32, top_k - jb);
#pragma unroll 5
      for (int j = 0; j < lim; ++j) {
        const int32_t n = __shfl_sync(0xffffffffu, my_n, j);
        const float w = __shfl_sync(0xffffffffu, my_w, j);
        if (n < 0) continue;
        const uint4* row = reinterpret_cast<const uint4*>(vbase + (int64_t)n * CV) + lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const uint4 u = __ldg(row + 32 * i);
          const uint32_t ws[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            acc[i][2 * h] = fmaf(w, __uint_as_float(ws[h] << 16), acc[i][2 * h]);
            acc[i][2 * h + 1] = fmaf(w, __uint_as_float(ws[h] & 0xffff0000u), acc[i][2 * h + 1]);
          }
        }
      }
    }
  }
  if constexpr (MODE != 0) {
    if (q < n_query) {
      float* row;
      if constexpr (MODE == 1) {
        row = out + ((int64_t)q * gridDim.y + o) * CV;
      } else {
        const uint32_t f = q_per_frame > 0 ? (uint32_t)q / (uint32_t)q_per_frame : 0u;
        const uint32_t pos = (uint32_t)q - f * (uint32_t)q_per_frame;
        row = out + (int64_t)f * frame_stride + (int64_t)o * out_obj_stride + (int64_t)pos * out_ch_stride;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        *reinterpret_cast<float4*>(row + i * 256 + lane * 8) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(row + i * 256 + lane * 8 + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) st[i * 256 + lane * 8 + e][warp] = acc[i][e];
  __syncthreads();
  const int nq_here = (int)min((int64_t)kQPerCta, n_query - q0);
  for (int e = threadIdx.x; e < CV * kQPerCta; e += kRoThreads) {
    const int c = e / kQPerCta, w = e % kQPerCta;
    if (w < nq_here)
      out[(int64_t)o * out_obj_stride + (int64_t)c * out_ch_stride + query_offset(q0 + w, q_per_frame, frame_stride)] = st[c][w];
  }
}

// Any CV / dtype: one CTA per (query, object), threads stride over channels.
template <typename VT>
__global__ void __launch_bounds__(128) readout_generic_kernel(
    const VT* __restrict__ val_pm, int64_t capacity_pos, int CV, const int32_t* __restrict__ idx,
    const float* __restrict__ weight, int64_t n_query, int top_k, float* __restrict__ out,
    int64_t out_obj_stride, int64_t out_ch_stride, int q_per_frame, int64_t frame_stride) {
  pdl_wait();  // idx / weight come from the kernel before (programmatic dependent launch)
  pdl_launch_dependents();
  __shared__ int32_t s_n[EVAVOS_MAX_TOPK];
  __shared__ float s_w[EVAVOS_MAX_TOPK];
  const int64_t q = blockIdx.x;
  const int o = blockIdx.y;
  for (int j = threadIdx.x; j < top_k; j += blockDim.x) {
    s_n[j] = idx[q * top_k + j];
    s_w[j] = weight[q * top_k + j];
  }
  __syncthreads();
  const VT* vbase = val_pm + (int64_t)o * capacity_pos * CV;
  for (int c = threadIdx.x; c < CV; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < top_k; ++j) {
      const int32_t n = s_n[j];
      if (n < 0) continue;
      float v;
      if constexpr (sizeof(VT) == 2) v = __bfloat162float(vbase[(int64_t)n * CV + c]);
      else v = vbase[(int64_t)n * CV + c];
      acc = fmaf(s_w[j], v, acc);
    }
    if (out_ch_stride < 0) out[((int64_t)q * gridDim.y + o) * CV + c] = acc;   // query-major (see launch_readout)
    else out[(int64_t)o * out_obj_stride + (int64_t)c * out_ch_stride + query_offset(q, q_per_frame, frame_stride)] = acc;
  }
}

__global__ void scatter_dense_kernel(const int32_t* __restrict__ idx, const float* __restrict__ weight,
                                     int64_t n_query, int top_k, int64_t n_pos, float* __restrict__ dense) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_query * top_k) return;
  const int64_t q = e / top_k;
  const int32_t n = idx[e];
  if (n >= 0 && n < n_pos) dense[(int64_t)n * n_query + q] = weight[e];
}

}  // namespace

// out_ch_stride < 0 selects the query-major output (n_query, K, CV); out_ch_stride == 1 a channels-last destination
// whose position stride is out_obj_stride / (q_per_frame or n_query) (see include/evavos.h, readout_ch_stride).
int launch_readout(const EvavosBankShadow& b, const int32_t* idx, const float* weight, int64_t n_query,
                   int top_k, float* out, int64_t out_obj_stride, int64_t out_ch_stride, int q_per_frame,
                   int64_t frame_stride, cudaStream_t st) {
  if (n_query <= 0) return EVAVOS_OK;
  const bool qmajor = out_ch_stride < 0;
  const bool chlast = out_ch_stride == 1 && n_query > 1;   // (one query: both readings address the same elements)
  if (chlast) {
    const int64_t per_obj = q_per_frame > 0 ? q_per_frame : n_query;
    if (out_obj_stride <= 0 || out_obj_stride % per_obj != 0 || (out_obj_stride / per_obj) % 4 != 0 ||
        out_obj_stride / per_obj < b.CV || reinterpret_cast<uintptr_t>(out) % 16 != 0 || frame_stride % 4 != 0) {
      set_error("readout: channels-last destination needs obj_stride = positions x (16-byte aligned row of >= CV floats)");
      return EVAVOS_ERR_INVALID;
    }
    out_ch_stride = out_obj_stride / per_obj;     // from here on: the POSITION stride of the row kernels' MODE 2
  }
  const int mode = qmajor ? 1 : (chlast ? 2 : 0);
  if (out_ch_stride == 0) out_ch_stride = n_query;
  if (out_obj_stride == 0) out_obj_stride = (int64_t)b.CV * n_query;
  const dim3 grid((unsigned)ceil_div(n_query, kQPerCta), (unsigned)b.K);
  const bool row16 = (reinterpret_cast<uintptr_t>(b.val_pm) % 16) == 0;
#define EVAVOS_RO_F32_(NV, SPLIT, QM)                                                                         \
  EVAVOS_CUDA_OK(launch_pdl(readout_f32_kernel<NV, QM>, dim3(grid.x, grid.y, SPLIT), dim3(kRoThreads), 0, st,        \
                            reinterpret_cast<const float*>(b.val_pm), b.capacity_pos, b.CV, idx, weight,      \
                            n_query, top_k, out, out_obj_stride, out_ch_stride, q_per_frame, frame_stride))
#define EVAVOS_RO_F32(NV, SPLIT) do { if (mode == 1) EVAVOS_RO_F32_(NV, SPLIT, 1); else if (mode == 2) EVAVOS_RO_F32_(NV, SPLIT, 2); else EVAVOS_RO_F32_(NV, SPLIT, 0); } while (0)
#define EVAVOS_RO_BF16_(NV, QM)                                                                               \
  EVAVOS_CUDA_OK(launch_pdl(readout_bf16_kernel<NV, QM>, grid, dim3(kRoThreads), 0, st,                              \
                            reinterpret_cast<const __nv_bfloat16*>(b.val_pm), b.capacity_pos, idx, weight,    \
                            n_query, top_k, out, out_obj_stride, out_ch_stride, q_per_frame, frame_stride))
#define EVAVOS_RO_BF16(NV) do { if (mode == 1) EVAVOS_RO_BF16_(NV, 1); else if (mode == 2) EVAVOS_RO_BF16_(NV, 2); else EVAVOS_RO_BF16_(NV, 0); } while (0)
  if (b.val_dtype == EVAVOS_F32 && row16 && b.CV % 128 == 0 && b.CV <= 512) {
    // Splitting a row's channels over two CTAs (more resident warps) was measured SLOWER on B200 (55 vs 42 us at
    // cfg2: twice the L2 requests at half the size), so one warp keeps a whole value row.
    const bool split = false;
    switch (b.CV / 128) {
      case 1: EVAVOS_RO_F32(1, 1); break;
      case 2: if (split) EVAVOS_RO_F32(1, 2); else EVAVOS_RO_F32(2, 1); break;
      case 3: EVAVOS_RO_F32(3, 1); break;
      default: if (split) EVAVOS_RO_F32(2, 2); else EVAVOS_RO_F32(4, 1); break;
    }
  } else if (b.val_dtype == EVAVOS_BF16 && row16 && b.CV % 256 == 0 && b.CV <= 512) {
    if (b.CV == 256) EVAVOS_RO_BF16(1);
    else EVAVOS_RO_BF16(2);
  } else if (chlast) {
    set_error("readout: a channels-last destination needs CV %% 128 == 0 (fp32) / CV %% 256 == 0 (bf16), CV <= 512");
    return EVAVOS_ERR_UNSUPPORTED;
  } else {
    const dim3 g2((unsigned)n_query, (unsigned)b.K);
    if (b.val_dtype == EVAVOS_BF16)
      readout_generic_kernel<__nv_bfloat16><<<g2, 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(b.val_pm),
                                                                b.capacity_pos, b.CV, idx, weight, n_query, top_k,
                                                                out, out_obj_stride, out_ch_stride, q_per_frame, frame_stride);
    else
      readout_generic_kernel<float><<<g2, 128, 0, st>>>(reinterpret_cast<const float*>(b.val_pm), b.capacity_pos,
                                                        b.CV, idx, weight, n_query, top_k, out, out_obj_stride,
                                                        out_ch_stride, q_per_frame, frame_stride);
  }
#undef EVAVOS_RO_F32
#undef EVAVOS_RO_F32_
#undef EVAVOS_RO_BF16
#undef EVAVOS_RO_BF16_
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

int launch_affinity_dense(const int32_t* idx, const float* weight, int64_t n_query, int top_k, int64_t n_pos,
                          float* dense, cudaStream_t st) {
  EVAVOS_CUDA_OK(cudaMemsetAsync(dense, 0, sizeof(float) * (size_t)n_pos * (size_t)n_query, st));
  const int64_t total = n_query * top_k;
  scatter_dense_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(idx, weight, n_query, top_k, n_pos, dense);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

}  // namespace evavos
