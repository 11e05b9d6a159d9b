// Exact selection for the queries the bf16 filter cannot separate.
//
// A query whose candidate list overflowed has more than kCandCap positions inside the filter's error margin of its
// k-th best score: a bank of near-identical keys (static scenes, repeated frames; networks with random weights make
// EVERY query look like this).  Round 1/2 redid such a query with one warp streaming the whole bank (~0.3 ms per
// query, 2.7 ms per read of the cfg3 workload).  Here the overflowed queries - the finalizer lists them - are scored
// exactly in tiles of 32 queries x 128 positions with an fp32 register-tiled contraction whose per-element arithmetic
// is the finalizer's own (channel order 0..63, one FMA per term, affinity_from_parts), so both paths break near-ties
// identically:
//   pass 1: exact class maxima (class = position mod 128) over every R-th block of 128 positions; the k-th largest
//           of the 128 maxima is a lower bound of the k-th best score - exact scores, no error margin;
//   pass 2: all positions, every score reaching the bound goes into the query's list (~1.3 k R entries whatever the
//           score distribution: the bound is a rank statistic);
//   pass 3: the finalizer itself (finalize_query_warp) cuts, ranks and writes the top-k, one warp per query.
// Only an exact tie of more than kCandCap scores at the boundary still takes the single-warp path.
// Replaces nothing in the reference (torch.topk over the dense affinity, prop_net.py:53): it is the tail of
// candidate selection, DESIGN.md section 3.
#include "select_common.cuh"

namespace evavos {

namespace {

// Queries per warp (each lane holds kOvQW x 4 accumulators) and resident CTAs per SM.  Measured on B200 (8 100 queries x
// 8 100 positions, whole read incl. filter, scripts/dense_time.py): 4 x 2 CTAs 844 us | 8 x 1 CTA 871 | 4 x 1 947 |
// 8 x 2 (spills) 1 367.  With 4 the tile loop moves 20 LSU cycles per 16 FMA cycles, with 8 it is FMA-bound on paper -
// but then only 8 warps fit an SM, and the kernel is latency-bound either way (FMA pipe ~30 % busy under ncu).
#ifndef EVAVOS_OVQW
#define EVAVOS_OVQW 4
#endif
#ifndef EVAVOS_OVMINB
#define EVAVOS_OVMINB 2
#endif
constexpr int kOvQW = EVAVOS_OVQW;   // queries per warp
constexpr int kOvQ = 8 * kOvQW;   // queries per CTA tile
constexpr int kOvP = 128;      // positions per step (4 per lane); also the number of classes
constexpr int kOvThreads = 256;
constexpr int kOvWarps = kOvThreads / 32;

struct OverflowTiles {
  float q[kOvQ][kRowStride];
  float k[2][kOvP][kRowStride];
  float kk[2][kOvP];
};

union OverflowSmem {
  OverflowTiles t;
  FinalizeWarpSmem fin[kOvWarps];
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 128 key rows (64 fp32 channels each) of block `blk` into k[buf]; rows past the end of the bank are zero-filled.
__device__ __forceinline__ void load_block(OverflowTiles& t, int buf, const float* __restrict__ key_pm, int64_t n_pos,
                                           int64_t blk, int tid) {
  const int64_t p0 = blk * kOvP;
#pragma unroll
  for (int i = 0; i < (kOvP * 16) / kOvThreads; ++i) {
    const int e = tid + i * kOvThreads;
    const int row = e >> 4, c4 = e & 15;
    float* dst = &t.k[buf][row][4 * c4];
    if (p0 + row < n_pos) cp_async16(dst, key_pm + (p0 + row) * 64 + 4 * c4);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_commit();
}

__global__ void __launch_bounds__(kOvThreads, EVAVOS_OVMINB) overflow_exact_kernel(
    const float* __restrict__ key_pm, const float* __restrict__ query, int64_t query_ch_stride, int64_t n_pos,
    int64_t n_query, int top_k, int2* __restrict__ cand, const int32_t* __restrict__ overflow_list,
    const unsigned int* __restrict__ overflow_cnt, const float* __restrict__ key_maxnorm,
    int32_t* __restrict__ out_idx, float* __restrict__ out_weight, float* __restrict__ out_score,
    const PeerPush push) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  OverflowSmem& sm = *reinterpret_cast<OverflowSmem*>(smem_raw);
  __shared__ int32_t s_qid[kOvQ];
  __shared__ int s_cnt[kOvQ];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();   // the list of overflowed queries comes from the finalizer
  pdl_launch_dependents();
  const int n_over = (int)__ldcg(overflow_cnt);
  if (n_over <= 0) return;

  const float inv_sqrt_ck = 1.0f / sqrtf(64.0f);
  const int64_t n_blocks = (n_pos + kOvP - 1) / kOvP;
  // pass 1 samples every 2nd / 4th block once the sample still fills every class >= 4 times: the bound is the
  // ~1.3 k-th best of the sample, so the list gets ~1.3 k sample entries whatever the bank length (k <= 64: <= 333,
  // far below kCandCap)
  const int sample = top_k > 64 ? 1 : (n_blocks >= 16 ? 4 : (n_blocks >= 8 ? 2 : 1));

  for (int tile = blockIdx.x; tile * kOvQ < n_over; tile += gridDim.x) {
    __syncthreads();   // the previous tile's finalizer scratch aliases the tiles
    if (tid < kOvQ) {
      const int e = tile * kOvQ + tid;
      s_qid[tid] = e < n_over ? __ldcg(overflow_list + e) : -1;
      s_cnt[tid] = 0;
    }
    __syncthreads();
    for (int e = tid; e < kOvQ * 64; e += kOvThreads) {
      const int qi = e % kOvQ, c = e / kOvQ;
      const int32_t qid = s_qid[qi];
      sm.t.q[qi][c] = qid >= 0 ? __ldg(query + (int64_t)c * query_ch_stride + qid) : 0.f;
    }
    __syncthreads();
    float qq[kOvQW];
#pragma unroll
    for (int j = 0; j < kOvQW; ++j) qq[j] = sumsq64_warp(sm.t.q[warp * kOvQW + j], lane);

    float cmax[kOvQW][4];
#pragma unroll
    for (int j = 0; j < kOvQW; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) cmax[j][i] = -INFINITY;
    uint32_t bound[kOvQW];   // order-preserving keys; pass 2 lists every score whose key reaches them
#pragma unroll
    for (int j = 0; j < kOvQW; ++j) bound[j] = 0u;

    for (int pass = 1; pass <= 2; ++pass) {
      const int64_t step_blocks = pass == 1 ? sample : 1;
      const int64_t n_steps = (n_blocks + step_blocks - 1) / step_blocks;
      load_block(sm.t, 0, key_pm, n_pos, 0, tid);
      for (int64_t s = 0; s < n_steps; ++s) {
        const int buf = (int)(s & 1);
        const int64_t blk = s * step_blocks;
        if (s + 1 < n_steps) {
          load_block(sm.t, buf ^ 1, key_pm, n_pos, (s + 1) * step_blocks, tid);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
        if (tid < kOvP) {   // |k|^2 of every row, once per block, in the finalizer's order
          float kk = 0.f;
          const float4* row = reinterpret_cast<const float4*>(sm.t.k[buf][tid]);
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4) {
            const float4 v = row[c4];
            kk = fmaf(v.x, v.x, kk);
            kk = fmaf(v.y, v.y, kk);
            kk = fmaf(v.z, v.z, kk);
            kk = fmaf(v.w, v.w, kk);
          }
          sm.t.kk[buf][tid] = kk;
        }
        __syncthreads();

        float acc[kOvQW][4];   // [query j][position i]: k.q, channels in order, one accumulator per pair
#pragma unroll
        for (int j = 0; j < kOvQW; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll 2
        for (int c4 = 0; c4 < 16; ++c4) {
          float4 kv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) kv[i] = *reinterpret_cast<const float4*>(&sm.t.k[buf][lane + 32 * i][4 * c4]);
#pragma unroll
          for (int j = 0; j < kOvQW; ++j) {
            const float4 qv = *reinterpret_cast<const float4*>(&sm.t.q[warp * kOvQW + j][4 * c4]);   // broadcast
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[j][i] = fmaf(kv[i].x, qv.x, acc[j][i]);
              acc[j][i] = fmaf(kv[i].y, qv.y, acc[j][i]);
              acc[j][i] = fmaf(kv[i].z, qv.z, acc[j][i]);
              acc[j][i] = fmaf(kv[i].w, qv.w, acc[j][i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t n = blk * kOvP + lane + 32 * i;
          const bool live = n < n_pos;
          const float kk = sm.t.kk[buf][lane + 32 * i];
#pragma unroll
          for (int j = 0; j < kOvQW; ++j) {
            const float sc = affinity_from_parts(kk, acc[j][i], qq[j], inv_sqrt_ck);
            if (pass == 1) {
              if (live) cmax[j][i] = fmaxf(cmax[j][i], sc);
            } else if (live && float_to_ordered(sc) >= bound[j]) {
              const int qi = warp * kOvQW + j;
              const int32_t qid = s_qid[qi];
              if (qid >= 0) {
                const int slot = atomicAdd(&s_cnt[qi], 1);
                if (slot < kCandCap) cand[(int64_t)qid * kCandCap + slot] = make_int2((int32_t)n, __float_as_int(sc));
              }
            }
          }
        }
        __syncthreads();   // k[buf] is refilled two steps from now, kk[buf] likewise
      }
      if (pass == 1) {
        const int k = top_k < kOvP ? top_k : kOvP;
#pragma unroll
        for (int j = 0; j < kOvQW; ++j) {
          uint32_t key[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) key[i] = float_to_ordered(cmax[j][i]);
          bound[j] = kth_largest_bound<4>(key, k);
        }
      }
    }

    // pass 3: the finalizer, one warp per query, on lists that carry exact scores (cut with a zero margin)
    __syncthreads();   // every listed entry is written; the tiles are dead
    for (int j = 0; j < kOvQW; ++j) {
      const int qi = warp * kOvQW + j;
      const int32_t qid = s_qid[qi];
      if (qid < 0) continue;
      finalize_query_warp(sm.fin[warp], lane, qid, key_pm, query, query_ch_stride, 64, n_pos, top_k, cand, s_cnt[qi],
                          /*scored=*/2, key_maxnorm, out_idx, out_weight, out_score, push, n_query);
    }
  }
}

}  // namespace

int launch_overflow_exact(const float* key_pm, const float* query, int64_t query_ch_stride, int64_t n_pos,
                          int64_t n_query, int top_k, int2* cand, const int32_t* overflow_list,
                          const unsigned int* overflow_cnt, const float* key_maxnorm, int32_t* out_idx,
                          float* out_weight, float* out_score, const EvavosPeers* peers, int64_t peer_gather_offset,
                          int n_sm, cudaStream_t st) {
  static bool attr_set[64] = {};
  int dev = 0;
  EVAVOS_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    EVAVOS_CUDA_OK(cudaFuncSetAttribute(overflow_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(OverflowSmem)));
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  PeerPush push;
  memset(&push, 0, sizeof(push));
  if (peers != nullptr) {
    push.n_ranks = peers->n_ranks;
    push.rank = peers->rank;
    for (int g = 0; g < peers->n_ranks; ++g)
      push.dst[g] = reinterpret_cast<int2*>(reinterpret_cast<uint8_t*>(peers->base[g]) + peer_gather_offset);
  }
  int64_t grid = ceil_div(n_query, kOvQ);
  if (grid > 2 * (int64_t)n_sm) grid = 2 * (int64_t)n_sm;
  EVAVOS_CUDA_OK(launch_pdl(overflow_exact_kernel, dim3((unsigned)grid), dim3(kOvThreads), sizeof(OverflowSmem), st,
                            key_pm, query, query_ch_stride, n_pos, n_query, top_k, cand, overflow_list, overflow_cnt,
                            key_maxnorm, out_idx, out_weight, out_score, push));
  return EVAVOS_OK;
}

}  // namespace evavos
