// Exact fp32 selection kernels on CUDA cores.
//
//  brute_select_kernel  exact top-k of one query over ALL memory positions by a 4-round
//                       8-bit radix select on order-preserving score keys (scores are
//                       recomputed each round; no N-sized scratch).  It is the whole
//                       selection in EVAVOS_PATH_SIMT (and for CK != 64).
//  finalize_kernel      one warp per query: cuts a scored candidate list of the tcgen05 filter down with the
//                       bound its own scores give, rescoring of the survivors in exact fp32, top-k by
//                       (score desc, position asc), softmax over the survivors
//                       (softmax_w_g_top, prop_net.py:53-57).  A query whose list overflowed (thousands of
//                       tied keys) is redone exactly by the same warp (exact_topk_warp).
//
// All of them agree on one arithmetic for a score: see dot_row / affinity_from_parts.
#include <cstring>

#include "common.cuh"
#include "select_common.cuh"

namespace evavos {

namespace {

constexpr int kBruteQ = 4;  // queries per CTA pass

// Scores of position n against the kBruteQ queries staged in shared memory.
__device__ __forceinline__ void score_group(const float* __restrict__ key_pm, int64_t n, int CK,
                                            const float (*qs)[64], const float* qq, float inv_sqrt_ck,
                                            float* s) {
  const float4* krow = reinterpret_cast<const float4*>(key_pm + n * CK);
  float kk = 0.f, kq[kBruteQ];
#pragma unroll
  for (int j = 0; j < kBruteQ; ++j) kq[j] = 0.f;
  for (int c4 = 0; c4 < (CK >> 2); ++c4) {
    const float4 kv = __ldg(krow + c4);
    kk = fmaf(kv.x, kv.x, kk);
    kk = fmaf(kv.y, kv.y, kk);
    kk = fmaf(kv.z, kv.z, kk);
    kk = fmaf(kv.w, kv.w, kk);
#pragma unroll
    for (int j = 0; j < kBruteQ; ++j) {
      const float4 qv = *reinterpret_cast<const float4*>(&qs[j][4 * c4]);
      kq[j] = fmaf(kv.x, qv.x, kq[j]);
      kq[j] = fmaf(kv.y, qv.y, kq[j]);
      kq[j] = fmaf(kv.z, qv.z, kq[j]);
      kq[j] = fmaf(kv.w, qv.w, kq[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < kBruteQ; ++j) s[j] = affinity_from_parts(kk, kq[j], qq[j], inv_sqrt_ck);
}

__global__ void __launch_bounds__(256) brute_select_kernel(
    const float* __restrict__ key_pm, const float* __restrict__ query, int64_t query_ch_stride, int CK,
    int64_t n_pos, int64_t n_query, int top_k, int2* __restrict__ cand, int32_t* __restrict__ cand_cnt) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float qs[kBruteQ][64];
  __shared__ float qq[kBruteQ];
  __shared__ int qid[kBruteQ];
  __shared__ unsigned hist[kBruteQ][256];
  __shared__ uint32_t prefix[kBruteQ];
  __shared__ int remaining[kBruteQ];
  __shared__ int gt_cnt[kBruteQ];
  __shared__ int eq_base[kBruteQ];
  __shared__ int warp_eq[8][kBruteQ];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t groups = (n_query + kBruteQ - 1) / kBruteQ;
  const float inv_sqrt_ck = 1.0f / sqrtf((float)CK);

  for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x) {
    __syncthreads();
    if (tid < kBruteQ) {
      const int64_t w = grp * kBruteQ + tid;
      int q = -1;
      if (w < n_query) q = (int)w;
      qid[tid] = q;
      remaining[tid] = top_k;
      prefix[tid] = 0;
      gt_cnt[tid] = 0;
      eq_base[tid] = 0;
    }
    __syncthreads();
    for (int e = tid; e < kBruteQ * 64; e += 256) {
      const int j = e >> 6, c = e & 63;
      const int q = qid[j];
      qs[j][c] = (q >= 0 && c < CK) ? query[(int64_t)c * query_ch_stride + q] : 0.f;
    }
    __syncthreads();
    if (CK == 64) {   // the same |q|^2 arithmetic as finalize_query
      if ((tid >> 5) < kBruteQ) {
        const float v = sumsq64_warp(qs[tid >> 5], tid & 31);
        if ((tid & 31) == 0) qq[tid >> 5] = v;
      }
    } else if (tid < kBruteQ) {
      qq[tid] = sumsq(qs[tid], CK);
    }
    __syncthreads();

    // ---- 4 radix rounds, most significant byte first --------------------------------------
    for (int round = 0; round < 4; ++round) {
      const int shift = 24 - 8 * round;
      for (int e = tid; e < kBruteQ * 256; e += 256) (&hist[0][0])[e] = 0;
      __syncthreads();
      uint32_t pfx[kBruteQ];
#pragma unroll
      for (int j = 0; j < kBruteQ; ++j) pfx[j] = prefix[j];
      for (int64_t n = tid; n < n_pos; n += 256) {
        float s[kBruteQ];
        score_group(key_pm, n, CK, qs, qq, inv_sqrt_ck, s);
#pragma unroll
        for (int j = 0; j < kBruteQ; ++j) {
          const uint32_t key = float_to_ordered(s[j]);
          const bool match = (round == 0) || ((key >> (shift + 8)) == pfx[j]);
          if (match && qid[j] >= 0) atomicAdd(&hist[j][(key >> shift) & 255u], 1u);
        }
      }
      __syncthreads();
      if (tid < kBruteQ) {
        const int rem = remaining[tid];
        int d = 255;
        int cum = 0;
        for (; d > 0; --d) {
          const int c = (int)hist[tid][d];
          if (cum + c >= rem) break;
          cum += c;
        }
        prefix[tid] = (round == 0 ? 0u : (prefix[tid] << 8)) | (uint32_t)d;
        remaining[tid] = rem - cum;
      }
      __syncthreads();
    }

    // ---- collect: everything above the k-th key, then the lowest positions among ties ------
    uint32_t T[kBruteQ];
    int need[kBruteQ];
#pragma unroll
    for (int j = 0; j < kBruteQ; ++j) { T[j] = prefix[j]; need[j] = remaining[j]; }
    for (int64_t base = 0; base < n_pos; base += 256) {
      const int64_t n = base + tid;
      const bool in = n < n_pos;
      float s[kBruteQ];
      if (in) score_group(key_pm, n, CK, qs, qq, inv_sqrt_ck, s);
      bool eq[kBruteQ];
      int rank[kBruteQ];
      bool any_eq = false;
#pragma unroll
      for (int j = 0; j < kBruteQ; ++j) {
        const bool live = in && qid[j] >= 0;
        const uint32_t key = live ? float_to_ordered(s[j]) : 0u;
        if (live && key > T[j]) {
          const int pos = atomicAdd(&gt_cnt[j], 1);
          cand[(int64_t)qid[j] * kCandCap + pos] = make_int2((int32_t)n, 0);
        }
        eq[j] = live && key == T[j];
        const unsigned m = __ballot_sync(0xffffffffu, eq[j]);
        rank[j] = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) warp_eq[warp][j] = __popc(m);
        any_eq |= (m != 0u);
      }
      if (__syncthreads_or(any_eq)) {
#pragma unroll
        for (int j = 0; j < kBruteQ; ++j) {
          if (eq[j]) {
            int r = eq_base[j] + rank[j];
            for (int w = 0; w < warp; ++w) r += warp_eq[w][j];
            if (r < need[j]) cand[(int64_t)qid[j] * kCandCap + (top_k - need[j]) + r] = make_int2((int32_t)n, 0);
          }
        }
        __syncthreads();
        if (tid < kBruteQ) {
          int tot = 0;
          for (int w = 0; w < 8; ++w) tot += warp_eq[w][tid];
          eq_base[tid] += tot;
        }
        __syncthreads();
      }
    }
    if (tid < kBruteQ && qid[tid] >= 0) cand_cnt[qid[tid]] = top_k;
  }
}

constexpr int kFinWarps = 4;   // queries per CTA

// One warp per query (finalize_query_warp).  Round 1 ran one 128-thread CTA per query with an all-pairs rank over
// ~70 candidates (17-19 us for 1 620 queries: a chain of dependent L2 round trips per CTA); a warp per query keeps
// four independent chains per CTA in flight and needs no CTA barrier.
__global__ void __launch_bounds__(32 * kFinWarps) finalize_kernel(
    const float* __restrict__ key_pm, const float* __restrict__ query, int64_t query_ch_stride, int CK, int64_t n_pos,
    int64_t n_query, int top_k, const int2* __restrict__ cand, const int32_t* __restrict__ cand_cnt, int scored,
    const float* __restrict__ key_maxnorm, int32_t* __restrict__ out_idx, float* __restrict__ out_weight,
    float* __restrict__ out_score, const PeerPush push, int32_t* __restrict__ overflow_list,
    unsigned int* __restrict__ overflow_cnt, uint32_t* __restrict__ overflow_hint) {
  __shared__ FinalizeWarpSmem sm[kFinWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * kFinWarps + warp;
  pdl_wait();
  pdl_launch_dependents();
  if (q >= n_query) return;
  const int cnt = __ldcg(cand_cnt + q);
  if (cnt > kCandCap && overflow_cnt != nullptr) {
    // more positions inside the filter's margin than a list holds.  Count it, tell the host (the next read will
    // launch the tiled pass, api.cu) and, when that pass follows this kernel, leave the query to it:
    // overflow_exact_kernel scores it exactly against the whole bank together with the other queries listed here.
    // Otherwise this warp redoes it alone (finalize_query_warp's exact path).
    unsigned int slot = 0;
    if (lane == 0) {
      slot = atomicAdd(overflow_cnt, 1u);
      if (overflow_hint != nullptr) *reinterpret_cast<volatile uint32_t*>(overflow_hint) = 1u;
    }
    if (overflow_list != nullptr) {
      if (lane == 0) overflow_list[slot] = (int32_t)q;
      return;
    }
  }
  finalize_query_warp(sm[warp], lane, q, key_pm, query, query_ch_stride, CK, n_pos, top_k, cand, cnt, scored,
                      key_maxnorm, out_idx, out_weight, out_score, push, n_query);
}

}  // namespace

int launch_brute_select(const float* key_pm, const float* query, int64_t query_ch_stride, int CK, int64_t n_pos,
                        int64_t n_query, int top_k, int2* cand, int32_t* cand_cnt, int n_sm, cudaStream_t st) {
  (void)n_sm;
  int64_t grid = ceil_div(n_query, kBruteQ);
  if (grid > 0x7fffffff) grid = 0x7fffffff;
  EVAVOS_CUDA_OK(launch_pdl(brute_select_kernel, dim3((unsigned)grid), dim3(256), 0, st, key_pm, query, query_ch_stride,
                            CK, n_pos, n_query, top_k, cand, cand_cnt));
  return EVAVOS_OK;
}

int launch_finalize(const float* key_pm, const float* query, int64_t query_ch_stride, int CK, int64_t n_pos,
                    int64_t n_query, int top_k, const int2* cand, const int32_t* cand_cnt, int scored,
                    const float* key_maxnorm, int32_t* out_idx, float* out_weight, float* out_score,
                    const EvavosPeers* peers, int64_t peer_gather_offset, int32_t* overflow_list,
                    unsigned int* overflow_cnt, uint32_t* overflow_hint, cudaStream_t st) {
  PeerPush push;
  memset(&push, 0, sizeof(push));
  if (peers != nullptr) {
    push.n_ranks = peers->n_ranks;
    push.rank = peers->rank;
    for (int g = 0; g < peers->n_ranks; ++g)
      push.dst[g] = reinterpret_cast<int2*>(reinterpret_cast<uint8_t*>(peers->base[g]) + peer_gather_offset);
  }
  EVAVOS_CUDA_OK(launch_pdl(finalize_kernel, dim3((unsigned)ceil_div(n_query, kFinWarps)), dim3(32 * kFinWarps), 0, st,
                            key_pm, query, query_ch_stride, CK, n_pos, n_query, top_k, cand, cand_cnt, scored,
                            key_maxnorm, out_idx, out_weight, out_score, push, overflow_list, overflow_cnt, overflow_hint));
  return EVAVOS_OK;
}

}  // namespace evavos
