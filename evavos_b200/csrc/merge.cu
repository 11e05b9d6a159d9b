// Merge step of the memory-axis (THW) sharded read (no reference counterpart).
//
// After the all-gather every rank holds, per query, n_shards * top_k (score, global position)
// candidates.  One warp per query selects the global top_k (score descending, position
// ascending on ties - the same order finalize_kernel uses), computes softmax weights with the
// GLOBAL maximum and denominator, and maps the winners this shard owns back to local positions
// so the ordinary sparse readout produces this shard's partial sum.  A sum all-reduce of the
// partial readouts then equals the single-device readout.
#include "common.cuh"

namespace evavos {

namespace {

// GATHERED = false: cand_idx / cand_score are (n_query, n_cand) with GLOBAL positions.
// GATHERED = true : cand_idx is the raw all-gather output [n_shards][n_query][per_shard][2] of packed
//                   (LOCAL position, score bits) pairs, n_cand = n_shards * per_shard; the position of entry c
//                   comes from shard c / per_shard and is mapped to its global position here.
template <bool GATHERED>
__global__ void __launch_bounds__(128) topk_merge_kernel(
    const int32_t* __restrict__ cand_idx, const float* __restrict__ cand_score, int64_t n_query, int n_cand,
    int top_k, int shard, int n_shards, int64_t pos_per_frame, int32_t* __restrict__ out_idx,
    float* __restrict__ out_weight, float* __restrict__ out_score, int32_t* __restrict__ local_idx) {
  extern __shared__ unsigned long long smem_keys[];  // [4 warps][n_cand] + [4][EVAVOS_MAX_TOPK]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 4 + warp;
  if (q >= n_query) return;
  unsigned long long* keys = smem_keys + (size_t)warp * n_cand;
  unsigned long long* sel = smem_keys + (size_t)4 * n_cand + (size_t)warp * EVAVOS_MAX_TOPK;
  int live = 0;
  const int per_shard = n_cand / n_shards;
  const uint32_t ppf = (uint32_t)pos_per_frame;   // positions fit int32 (the merged indices are int32): 32-bit division
  for (int c = lane; c < n_cand; c += 32) {
    int64_t n;
    float sc;
    if constexpr (GATHERED) {
      const int src = c / per_shard, j = c - src * per_shard;
      const int2 e = *reinterpret_cast<const int2*>(cand_idx + (((int64_t)src * n_query + q) * per_shard + j) * 2);
      const int32_t loc = e.x;
      sc = __int_as_float(e.y);
      n = -1;
      if (loc >= 0) {
        const uint32_t frame = (uint32_t)loc / ppf, r = (uint32_t)loc - frame * ppf;
        n = (int64_t)((frame * (uint32_t)n_shards + (uint32_t)src) * ppf + r);
      }
    } else {
      n = cand_idx[q * n_cand + c];
      sc = cand_score[q * n_cand + c];
    }
    unsigned long long k = 0ull;
    if (n >= 0) {
      k = ((unsigned long long)float_to_ordered(sc) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
      ++live;
    }
    keys[c] = k;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
  __syncwarp();
  const int take = min(top_k, live);
  if (n_shards <= 32) {
    // Every shard's list arrives best-first (finalize_kernel's order), so the global order is a tournament of the
    // list heads: lane s owns list s, one warp-wide max and one pointer advance per output.
    int ptr = 0;
    unsigned long long head = lane < n_shards ? keys[lane * per_shard] : 0ull;
    // (64-bit maximum as two 32-bit warp reductions - score bits, then position bits among the lanes that hold the
    //  best score: 2 REDUX instead of 5 x (2 SHFL + a 64-bit compare); the kernel is instruction-bound)
    for (int j = 0; j < take; ++j) {
      const uint32_t hi = (uint32_t)(head >> 32), lo = (uint32_t)head;
      const uint32_t best_hi = __reduce_max_sync(0xffffffffu, hi);
      const uint32_t best_lo = __reduce_max_sync(0xffffffffu, hi == best_hi ? lo : 0u);
      if (hi == best_hi && lo == best_lo && head != 0ull) {  // keys are unique: exactly one lane advances
        ++ptr;
        head = ptr < per_shard ? keys[lane * per_shard + ptr] : 0ull;
      }
      if (lane == 0) sel[j] = ((unsigned long long)best_hi << 32) | (unsigned long long)best_lo;
    }
    __syncwarp();
  } else {
    for (int j = 0; j < take; ++j) {
      unsigned long long best = 0ull;
      for (int c = lane; c < n_cand; c += 32) best = keys[c] > best ? keys[c] : best;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      for (int c = lane; c < n_cand; c += 32)
        if (keys[c] == best) keys[c] = 0ull;  // positions are unique across shards
      if (lane == 0) sel[j] = best;
      __syncwarp();
    }
  }
  const float s0 = take > 0 ? ordered_to_float((uint32_t)(sel[0] >> 32)) : 0.f;
  float e[EVAVOS_MAX_TOPK / 32];
  float part = 0.f;
#pragma unroll
  for (int t = 0; t < EVAVOS_MAX_TOPK / 32; ++t) {
    const int j = lane + 32 * t;
    e[t] = 0.f;
    if (j < take) {
      e[t] = expf(ordered_to_float((uint32_t)(sel[j] >> 32)) - s0);
      part += e[t];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
#pragma unroll
  for (int t = 0; t < EVAVOS_MAX_TOPK / 32; ++t) {
    const int j = lane + 32 * t;
    if (j >= top_k) continue;
    const bool ok = j < take;
    const int64_t o = q * top_k + j;
    const int64_t pos = ok ? (int64_t)(0xffffffffu - (uint32_t)(sel[j] & 0xffffffffull)) : -1;
    if (out_idx) out_idx[o] = (int32_t)pos;
    if (out_weight) out_weight[o] = ok ? e[t] / part : 0.f;
    if (out_score) out_score[o] = ok ? ordered_to_float((uint32_t)(sel[j] >> 32)) : -INFINITY;
    if (local_idx) {
      int32_t loc = -1;
      if (ok) {
        const uint32_t frame = (uint32_t)pos / ppf, r = (uint32_t)pos - frame * ppf;
        const uint32_t lf = frame / (uint32_t)n_shards;
        if (frame - lf * (uint32_t)n_shards == (uint32_t)shard) loc = (int32_t)(lf * ppf + r);
      }
      local_idx[o] = loc;
    }
  }
}

}  // namespace

int launch_topk_merge(const int32_t* cand_idx, const float* cand_score, int64_t n_query, int n_cand, int top_k,
                      int shard, int n_shards, int64_t pos_per_frame, int32_t* out_idx, float* out_weight,
                      float* out_score, int32_t* local_idx, int gathered, cudaStream_t st) {
  if (n_query <= 0) return EVAVOS_OK;
  const size_t smem = sizeof(unsigned long long) * ((size_t)4 * n_cand + 4 * EVAVOS_MAX_TOPK);
  if (smem > 200 * 1024) {
    set_error("topk_merge: n_cand=%d too large", n_cand);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024) {
    EVAVOS_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EVAVOS_CUDA_OK(cudaFuncSetAttribute(topk_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const unsigned grid = (unsigned)ceil_div(n_query, 4);
  if (gathered)
    topk_merge_kernel<true><<<grid, 128, smem, st>>>(cand_idx, cand_score, n_query, n_cand, top_k, shard, n_shards,
                                                     pos_per_frame, out_idx, out_weight, out_score, local_idx);
  else
    topk_merge_kernel<false><<<grid, 128, smem, st>>>(cand_idx, cand_score, n_query, n_cand, top_k, shard, n_shards,
                                                      pos_per_frame, out_idx, out_weight, out_score, local_idx);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

}  // namespace evavos
