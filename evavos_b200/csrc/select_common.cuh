// Exact fp32 scoring and the per-query finalizer (one warp per query).
#pragma once

#include "common.cuh"

namespace evavos {

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

__device__ __forceinline__ unsigned long long score_key(float s, int32_t n) {
  // larger = better: score descending, then position ascending
  return ((unsigned long long)float_to_ordered(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
}

// Sharded read: the gather regions of all ranks (see EvavosMemReadArgs.peers); n_ranks == 0 when not sharded.
struct PeerPush {
  int n_ranks;
  int rank;
  int2* dst[EVAVOS_MAX_RANKS];   // rank g's region [n_ranks][n_query][top_k] of (local position, score bits)
};

// Per-warp scratch of the finalizer (one warp finalizes one query).
struct FinalizeWarpSmem {
  float qs[64];
  float rows[32][kRowStride];
  unsigned long long keys[kMaxSurvivors];   // survivor positions first, their exact (score, position) keys afterwards
  unsigned long long sel[EVAVOS_MAX_TOPK];
};

// A lower bound (20 leading bits) of the k-th largest filter score among list[0..n): radix descent over the
// order-preserving keys, NJ entries per lane in registers.  n <= 32 * NJ, k <= n.
// low_bit = 0 resolves all 32 bits: the k-th largest key itself.
template <int NJ>
__device__ __forceinline__ uint32_t kth_largest_bound(const uint32_t* key, int k, int low_bit = 12) {
  uint32_t prefix = 0;
  int need = k;
  for (int bit = 31; bit >= low_bit; --bit) {
    const uint32_t want = (prefix >> bit) | 1u;
    int c = 0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) c += ((key[j] >> bit) == want) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= need) prefix |= 1u << bit;   // the k-th largest has this bit set
    else need -= c;
  }
  return prefix;
}

// Cut the scored candidate list down to the entries that can still be in the exact top-k and leave their positions
// in sm.keys[0..ns).  theta = (a lower bound of) the k-th largest filter score in the list: k listed positions reach
// it, so the k-th best exact score is >= theta - eps and every member of the exact top-k has a filter score
// >= theta - 2 eps.  Returns ns.  The survivors' positions are also compacted IN PLACE into list[0..ns).x (a slot
// never lies beyond the entries already read), which is where a query with more than kMaxSurvivors of them - a big
// cluster of near-ties, e.g. a static background seen in hundreds of memory frames - is finished from, in batches.
// EXACT: the listed scores are exact (overflow_exact_kernel; the arithmetic of rescore_rows): the cut is the k-th
// largest score itself - with near-constant keys a 20-bit bound lies below EVERY listed score and nothing would be
// cut - and the survivors' (score, position) keys are complete as they are: sm.keys receives them, nothing is rescored.
template <int NJ, bool EXACT = false>
__device__ __forceinline__ int prefilter_candidates(FinalizeWarpSmem& sm, int2* list, int n, int k, float two_eps,
                                                    int lane) {
  uint32_t key[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = lane + 32 * j;
    key[j] = e < n ? float_to_ordered(__int_as_float(__ldcg(&list[e].y))) : 0u;
  }
  uint32_t cut;
  if constexpr (EXACT) {
    cut = kth_largest_bound<NJ>(key, k, 0);
  } else {
    const float theta = ordered_to_float(kth_largest_bound<NJ>(key, k));
    cut = float_to_ordered(theta - two_eps);
  }
  int ns = 0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = lane + 32 * j;
    const bool pass = e < n && key[j] >= cut;
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (pass) {
      const int slot = ns + __popc(m & ((1u << lane) - 1u));
      const int32_t pos = __ldcg(&list[e].x);
      if (slot < kMaxSurvivors) {
        if constexpr (EXACT) sm.keys[slot] = ((unsigned long long)key[j] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)pos);
        else sm.keys[slot] = (unsigned long long)(uint32_t)pos;
      }
      list[slot].x = pos;
    }
    ns += __popc(m);
  }
  return ns;
}

// sm.keys[0..n) hold positions: replace them by the exact (score, position) keys.  32 rows per round.
__device__ __forceinline__ void rescore_rows(FinalizeWarpSmem& sm, int n, const float* __restrict__ key_pm, int CK,
                                             float qq, float inv_sqrt_ck, int lane) {
  for (int base = 0; base < n; base += 32) {
    const int nrows = min(32, n - base);
    float s = 0.f;
    int32_t pos = 0;
    if (CK == 64) {
      // Rows go through shared memory: half a warp fetches one 256-byte row (two full lines per row and
      // instruction instead of 32 partial ones when every lane walks its own row), then lane t rescoring row t
      // reads it back with the FMA order of dot_row.
      // (all 16 loads of a lane in flight before the first store: the rows come from HBM, one latency per round)
      float4 buf[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int r = (lane >> 4) + 2 * u;
        if (r < nrows) {
          const int32_t nr = (int32_t)(uint32_t)sm.keys[base + r];
          buf[u] = __ldg(reinterpret_cast<const float4*>(key_pm + (int64_t)nr * 64) + (lane & 15));
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int r = (lane >> 4) + 2 * u;
        if (r < nrows) *reinterpret_cast<float4*>(&sm.rows[r][4 * (lane & 15)]) = buf[u];
      }
      __syncwarp();
      if (lane < nrows) {
        pos = (int32_t)(uint32_t)sm.keys[base + lane];
        float kk, kq;
        dot_row_smem64(reinterpret_cast<const float4*>(sm.rows[lane]), sm.qs, kk, kq);
        s = affinity_from_parts(kk, kq, qq, inv_sqrt_ck);
      }
    } else if (lane < nrows) {
      pos = (int32_t)(uint32_t)sm.keys[base + lane];
      float kk, kq;
      dot_row(reinterpret_cast<const float4*>(key_pm + (int64_t)pos * CK), sm.qs, CK, kk, kq);
      s = affinity_from_parts(kk, kq, qq, inv_sqrt_ck);
    }
    __syncwarp();
    if (lane < nrows) sm.keys[base + lane] = score_key(s, pos);
  }
  __syncwarp();
}

// All-pairs rank of the unique keys sm.keys[0..n): the `take` largest go to sm.sel, best first.
__device__ __forceinline__ void rank_into_sel(FinalizeWarpSmem& sm, int n, int take, int lane) {
  for (int c = lane; c < n; c += 32) {
    const unsigned long long mine = sm.keys[c];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += sm.keys[j] > mine ? 1 : 0;  // broadcast reads
    if (rank < take) sm.sel[rank] = mine;
  }
  __syncwarp();
}

// Exact top-k of one query over ALL positions by one warp: a sorted list in sm.sel, 32 exact scores per step, an
// insertion only for scores that beat the current k-th (rare once the list has warmed up).  Same (score, position)
// order as everything else.  This is the slow path of queries whose candidate list overflowed (thousands of
// tied or nearly tied keys); returns the number of entries (min(top_k, n_pos)).
static __device__ __noinline__ int exact_topk_warp(FinalizeWarpSmem& sm, const float* __restrict__ key_pm, int CK,
                                            int64_t n_pos, int top_k, float qq, float inv_sqrt_ck, int lane) {
  int filled = 0;
  for (int64_t base = 0; base < n_pos; base += 32) {
    const int64_t n = base + lane;
    unsigned long long key = 0ull;
    if (n < n_pos) {
      float kk, kq;
      dot_row(reinterpret_cast<const float4*>(key_pm + n * CK), sm.qs, CK, kk, kq);
      key = score_key(affinity_from_parts(kk, kq, qq, inv_sqrt_ck), (int32_t)n);
    }
    unsigned long long kth = filled == top_k ? sm.sel[top_k - 1] : 0ull;
    unsigned m = __ballot_sync(0xffffffffu, key > kth);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const unsigned long long kk64 = __shfl_sync(0xffffffffu, key, src);
      kth = filled == top_k ? sm.sel[top_k - 1] : 0ull;
      if (kk64 <= kth) continue;   // the bar rose while this step's hits were being inserted (warp-uniform)
      int pos = 0;
      unsigned long long prev[EVAVOS_MAX_TOPK / 32];
#pragma unroll
      for (int g = 0; g < EVAVOS_MAX_TOPK / 32; ++g) {
        const int i = lane + 32 * g;
        pos += __popc(__ballot_sync(0xffffffffu, i < filled && sm.sel[i] > kk64));
        prev[g] = (i >= 1 && i - 1 < filled) ? sm.sel[i - 1] : 0ull;
      }
      const int new_filled = min(filled + 1, top_k);
      __syncwarp();
#pragma unroll
      for (int g = 0; g < EVAVOS_MAX_TOPK / 32; ++g) {
        const int i = lane + 32 * g;
        if (i == pos) sm.sel[i] = kk64;
        else if (i > pos && i < new_filled) sm.sel[i] = prev[g];
      }
      filled = new_filled;
      __syncwarp();
    }
  }
  return filled;
}

// Finalize one query with one warp: cut the candidate list (scored lists only), rescore the survivors exactly,
// rank them all-pairs (rank = number of survivors with a larger (score, -position) key), and write the first top_k
// best-first with their softmax weights exp(s - s0) / sum (prop_net.py:54-57).
__device__ __forceinline__ void finalize_query_warp(FinalizeWarpSmem& sm, int lane, int64_t q,
                                                    const float* __restrict__ key_pm, const float* __restrict__ query,
                                                    int64_t query_ch_stride, int CK, int64_t n_pos, int top_k,
                                                    const int2* __restrict__ cand, int cnt_raw, int scored,
                                                    const float* __restrict__ key_maxnorm,
                                                    int32_t* __restrict__ out_idx, float* __restrict__ out_weight,
                                                    float* __restrict__ out_score, const PeerPush& push,
                                                    int64_t n_query) {
  sm.qs[lane] = (lane < CK) ? __ldg(query + (int64_t)lane * query_ch_stride + q) : 0.f;
  sm.qs[lane + 32] = (lane + 32 < CK) ? __ldg(query + (int64_t)(lane + 32) * query_ch_stride + q) : 0.f;
  __syncwarp();
  const float inv_sqrt_ck = 1.0f / sqrtf((float)CK);
  const float qq = (CK == 64) ? sumsq64_warp(sm.qs, lane) : sumsq(sm.qs, CK);
  const int2* list = cand + q * kCandCap;

  int ns = -1;   // survivors (positions in sm.keys when <= kMaxSurvivors, else in list[].x), or -1: the exact path
  bool keys_ready = false;   // sm.keys already holds exact (score, position) keys: nothing to rescore
  if (cnt_raw <= kCandCap) {
    int2* wlist = const_cast<int2*>(list);   // the lists are workspace: the cut compacts them in place
    if (!scored) {
      ns = min(cnt_raw, kMaxSurvivors);
      for (int e = lane; e < ns; e += 32) sm.keys[e] = (unsigned long long)(uint32_t)__ldcg(&list[e].x);
    } else {
      // scored == 2: the list carries EXACT scores (overflow_exact_kernel): the cut needs no error margin
      const float two_eps = scored == 2 ? 0.f : 2.0f * filter_eps(sqrtf(qq), __ldg(key_maxnorm));
      const int k = min(top_k, cnt_raw);
      if (scored == 2) {
        if (cnt_raw <= 256) ns = prefilter_candidates<8, true>(sm, wlist, cnt_raw, k, 0.f, lane);
        else if (cnt_raw <= 512) ns = prefilter_candidates<16, true>(sm, wlist, cnt_raw, k, 0.f, lane);
        else ns = prefilter_candidates<32, true>(sm, wlist, cnt_raw, k, 0.f, lane);
        keys_ready = ns <= kMaxSurvivors;
      } else if (cnt_raw <= top_k + 32) {   // a list this short is not worth cutting: rescore all of it
        ns = cnt_raw;
        for (int e = lane; e < ns; e += 32) sm.keys[e] = (unsigned long long)(uint32_t)__ldcg(&list[e].x);
      } else if (cnt_raw <= 256) ns = prefilter_candidates<8>(sm, wlist, cnt_raw, k, two_eps, lane);
      else if (cnt_raw <= 512) ns = prefilter_candidates<16>(sm, wlist, cnt_raw, k, two_eps, lane);
      else ns = prefilter_candidates<32>(sm, wlist, cnt_raw, k, two_eps, lane);
    }
  }
  int take;
  __syncwarp();
  if (ns < 0) {
    take = exact_topk_warp(sm, key_pm, CK, n_pos, top_k, qq, inv_sqrt_ck, lane);
  } else if (ns <= kMaxSurvivors) {
    if (!keys_ready) rescore_rows(sm, ns, key_pm, CK, qq, inv_sqrt_ck, lane);
    take = min(top_k, ns);
    rank_into_sel(sm, ns, take, lane);
  } else {
    // more survivors than the buffer holds: batches of 128, each ranked together with the best-k so far
    // (cost grows linearly with the size of the near-tie cluster instead of falling off a cliff)
    __threadfence_block();
    take = 0;
    for (int base = 0; base < ns; base += 128) {
      const int nb = min(128, ns - base);
      for (int u = lane; u < nb; u += 32) sm.keys[u] = (unsigned long long)(uint32_t)__ldcg(&list[base + u].x);
      __syncwarp();
      rescore_rows(sm, nb, key_pm, CK, qq, inv_sqrt_ck, lane);
      for (int u = lane; u < take; u += 32) sm.keys[nb + u] = sm.sel[u];
      __syncwarp();
      const int total = nb + take;
      take = min(top_k, total);
      rank_into_sel(sm, total, take, lane);
    }
  }
  __syncwarp();
  const float s0 = take > 0 ? ordered_to_float((uint32_t)(sm.sel[0] >> 32)) : 0.f;
  float e[EVAVOS_MAX_TOPK / 32], part[EVAVOS_MAX_TOPK / 32];
#pragma unroll
  for (int g = 0; g < EVAVOS_MAX_TOPK / 32; ++g) {
    const int j = lane + 32 * g;
    e[g] = j < take ? expf(ordered_to_float((uint32_t)(sm.sel[j] >> 32)) - s0) : 0.f;  // exp(values - values[:,0])
    part[g] = e[g];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part[g] += __shfl_xor_sync(0xffffffffu, part[g], o);
  }
  const float total = (part[0] + part[1]) + (part[2] + part[3]);
#pragma unroll
  for (int g = 0; g < EVAVOS_MAX_TOPK / 32; ++g) {
    const int j = lane + 32 * g;
    if (j < top_k) {
      const bool live = j < take;
      const int64_t o = q * top_k + j;
      if (out_idx) out_idx[o] = live ? (int32_t)(0xffffffffu - (uint32_t)(sm.sel[j] & 0xffffffffull)) : -1;
      if (out_weight) out_weight[o] = live ? e[g] / total : 0.f;
      if (out_score) out_score[o] = live ? ordered_to_float((uint32_t)(sm.sel[j] >> 32)) : -INFINITY;
    }
  }
  // Sharded read: this query's list goes straight into every rank's gather region (stores over NVLink peer
  // memory) - the all-gather happens here, query by query, while the other queries are still being finalized.
  if (push.n_ranks > 0) {
    for (int g = 0; g < push.n_ranks; ++g) {
      int2* row = push.dst[g] + ((int64_t)push.rank * n_query + q) * top_k;
      for (int j = lane; j < top_k; j += 32) {
        const bool live = j < take;
        row[j] = make_int2(live ? (int32_t)(0xffffffffu - (uint32_t)(sm.sel[j] & 0xffffffffull)) : -1,
                           live ? (int32_t)ordered_to_float_bits((uint32_t)(sm.sel[j] >> 32)) : (int32_t)0xff800000u);
      }
    }
  }
  __syncwarp();
}

}  // namespace evavos
