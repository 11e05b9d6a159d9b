// Exact fp32 scoring and the per-query finalizer, shared by finalize_kernel and the fused filter kernel.
#pragma once

#include "common.cuh"

namespace evavos {

#ifdef EVAVOS_TRACE
static __device__ int g_fin_skip = 0;   // timing experiments only: 1 = no row loads, 2 = no rank loop, 4 = no |q|^2 loop
#define EVAVOS_FIN_SKIP(bit) (g_fin_skip & (bit))
#else
#define EVAVOS_FIN_SKIP(bit) 0
#endif

// kk += |k|^2, kq += k.q over CK channels, channel order 0..CK-1, one FMA per term.
__device__ __forceinline__ void dot_row(const float4* __restrict__ krow, const float* __restrict__ q, int CK,
                                        float& kk, float& kq) {
  kk = 0.f;
  kq = 0.f;
  if (CK == 64) {  // the network's key width: all 16 row loads in flight before the first FMA
    float4 kv[16];
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) kv[c4] = __ldg(krow + c4);
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 qv = *reinterpret_cast<const float4*>(q + 4 * c4);
      kk = fmaf(kv[c4].x, kv[c4].x, kk); kq = fmaf(kv[c4].x, qv.x, kq);
      kk = fmaf(kv[c4].y, kv[c4].y, kk); kq = fmaf(kv[c4].y, qv.y, kq);
      kk = fmaf(kv[c4].z, kv[c4].z, kk); kq = fmaf(kv[c4].z, qv.z, kq);
      kk = fmaf(kv[c4].w, kv[c4].w, kk); kq = fmaf(kv[c4].w, qv.w, kq);
    }
    return;
  }
  for (int c4 = 0; c4 < (CK >> 2); ++c4) {
    const float4 kv = __ldg(krow + c4);
    const float4 qv = *reinterpret_cast<const float4*>(q + 4 * c4);
    kk = fmaf(kv.x, kv.x, kk); kq = fmaf(kv.x, qv.x, kq);
    kk = fmaf(kv.y, kv.y, kk); kq = fmaf(kv.y, qv.y, kq);
    kk = fmaf(kv.z, kv.z, kk); kq = fmaf(kv.z, qv.z, kq);
    kk = fmaf(kv.w, kv.w, kk); kq = fmaf(kv.w, qv.w, kq);
  }
}

// The same arithmetic on a row that already sits in shared memory (CK == 64): identical FMA order, plain loads.
__device__ __forceinline__ void dot_row_smem64(const float4* krow, const float* q, float& kk, float& kq) {
  kk = 0.f;
  kq = 0.f;
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4) {
    const float4 kv = krow[c4];
    const float4 qv = *reinterpret_cast<const float4*>(q + 4 * c4);
    kk = fmaf(kv.x, kv.x, kk); kq = fmaf(kv.x, qv.x, kq);
    kk = fmaf(kv.y, kv.y, kk); kq = fmaf(kv.y, qv.y, kq);
    kk = fmaf(kv.z, kv.z, kk); kq = fmaf(kv.z, qv.z, kq);
    kk = fmaf(kv.w, kv.w, kk); kq = fmaf(kv.w, qv.w, kq);
  }
}

__device__ __forceinline__ float sumsq(const float* __restrict__ q, int CK) {
  float s = 0.f;
  for (int c = 0; c < CK; ++c) s = fmaf(q[c], q[c], s);
  return s;
}

constexpr int kRowChunk = 48;    // candidate key rows staged per round (finalize_query)
constexpr int kRowStride = 68;   // floats: 16-byte aligned rows, conflict-free LDS.128 when every lane owns a row

// |q|^2 of a 64-channel query held in shared memory, by one converged warp: two channels per lane, butterfly sum.
// Every kernel that scores against an exact key uses THIS order for CK == 64 (a different rounding of |q|^2 would
// let two paths break a near-tie at the top-k boundary differently).
__device__ __forceinline__ float sumsq64_warp(const float* q, int lane) {
  float s = fmaf(q[lane], q[lane], q[lane + 32] * q[lane + 32]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

struct FinalizeSmem {
  float qs[64];
  float rows[kRowChunk][kRowStride];
  unsigned long long keys[kCandCap];
  unsigned long long sel[EVAVOS_MAX_TOPK];
  float warp_sum[4];
};

// Finalize one query with a group of 128 threads (tid 0..127; `sync` is the group's barrier).
// Every thread rescoring one (or two) candidates exactly, keys go to shared memory, each thread ranks its
// candidates all-pairs (rank = number of candidates with a larger (score, -position) key), and the first top_k
// ranks are written best-first with their softmax weights exp(s - s0) / sum (prop_net.py:54-57).
template <typename Sync>
__device__ __forceinline__ void finalize_query(FinalizeSmem& sm, int tid, int64_t q, const float* __restrict__ key_pm,
                                               const float* __restrict__ query, int64_t query_ch_stride, int CK,
                                               int top_k, const int32_t* cand, int cnt_raw,
                                               int32_t* __restrict__ out_idx, float* __restrict__ out_weight,
                                               float* __restrict__ out_score, Sync sync) {
  const int lane = tid & 31, warp = tid >> 5;
  const int cnt = min(cnt_raw, kCandCap);
  if (tid < 64) sm.qs[tid] = (tid < CK) ? __ldg(query + (int64_t)tid * query_ch_stride + q) : 0.f;
  sync();
  const float inv_sqrt_ck = 1.0f / sqrtf((float)CK);
  if (CK == 64) {
    // |q|^2 once per warp (two channels per lane) instead of a 64-step chain in every thread
    const float qq = sumsq64_warp(sm.qs, lane);
    // Candidate rows go through shared memory: half a warp fetches one 256-byte row (two full lines per row and
    // instruction instead of 32 partial ones when every thread walks its own row), then thread t rescoring row t
    // reads it back with the FMA order of dot_row.
    for (int base = 0; base < cnt; base += kRowChunk) {
      const int nrows = min(kRowChunk, cnt - base);
      if (!EVAVOS_FIN_SKIP(1)) {
        for (int r = tid >> 4; r < nrows; r += 8) {
          const int32_t n = __ldcg(cand + q * kCandCap + base + r);
          const float4 v = __ldg(reinterpret_cast<const float4*>(key_pm + (int64_t)n * 64) + (tid & 15));
          *reinterpret_cast<float4*>(&sm.rows[r][4 * (tid & 15)]) = v;
        }
      }
      sync();
      if (tid < nrows) {
        const int32_t n = __ldcg(cand + q * kCandCap + base + tid);
        float kk, kq;
        dot_row_smem64(reinterpret_cast<const float4*>(sm.rows[tid]), sm.qs, kk, kq);
        const float s = affinity_from_parts(kk, kq, qq, inv_sqrt_ck);
        sm.keys[base + tid] =
            ((unsigned long long)float_to_ordered(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
      }
      sync();
    }
  } else {
    const float qq = sumsq(sm.qs, CK);
#pragma unroll
    for (int t = 0; t < kCandCap / 128; ++t) {
      const int ci = tid + 128 * t;
      if (ci < cnt) {
        const int32_t n = __ldcg(cand + q * kCandCap + ci);
        float kk, kq;
        dot_row(reinterpret_cast<const float4*>(key_pm + (int64_t)n * CK), sm.qs, CK, kk, kq);
        const float s = affinity_from_parts(kk, kq, qq, inv_sqrt_ck);
        sm.keys[ci] = ((unsigned long long)float_to_ordered(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
      }
    }
    sync();
  }
  const int take = min(top_k, cnt);
#pragma unroll
  for (int t = 0; t < kCandCap / 128; ++t) {
    const int ci = tid + 128 * t;
    if (ci < cnt) {
      const unsigned long long mine = sm.keys[ci];
      int rank = 0;
      if (EVAVOS_FIN_SKIP(2)) rank = ci;
      else
        for (int j = 0; j < cnt; ++j) rank += sm.keys[j] > mine ? 1 : 0;  // broadcast reads; keys are unique
      if (rank < take) sm.sel[rank] = mine;
    }
  }
  sync();
  const float s0 = take > 0 ? ordered_to_float((uint32_t)(sm.sel[0] >> 32)) : 0.f;
  float e = 0.f;
  if (tid < take) e = expf(ordered_to_float((uint32_t)(sm.sel[tid] >> 32)) - s0);  // exp(values - values[:,0])
  float part = e;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) sm.warp_sum[warp] = part;
  sync();
  const float total = (sm.warp_sum[0] + sm.warp_sum[1]) + (sm.warp_sum[2] + sm.warp_sum[3]);
  if (tid < top_k) {
    const bool live = tid < take;
    const int64_t o = q * top_k + tid;
    if (out_idx) out_idx[o] = live ? (int32_t)(0xffffffffu - (uint32_t)(sm.sel[tid] & 0xffffffffull)) : -1;
    if (out_weight) out_weight[o] = live ? e / total : 0.f;
    if (out_score) out_score[o] = live ? ordered_to_float((uint32_t)(sm.sel[tid] >> 32)) : -INFINITY;
  }
}

}  // namespace evavos
