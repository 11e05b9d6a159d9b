// Full-softmax attention read of ONE memory frame (AttentionMemory + the two vector-matrix products of
// get_attention, mivos/model/propagation/prop_net.py:117-138, 198-211), flash style: the (HW x HW) softmax matrix W
// is never materialised.
//
//   out[c][q] = sum_n vec[c][n] * softmax_n( (-|m_n|^2 + 2 m_n.q_q - |q_q|^2) / sqrt(CK) )
//
// The -|q|^2 term is constant along n and cancels in the softmax.  Two forms of the partial pass:
//  * CUDA cores, fp32 (attention_partial_kernel): the contraction is 2*HW*HW*CK = 0.34 GF at 480p - launch-latency
//    territory;
//  * tensor cores (attention_partial_tc_kernel) for large maps (8.5 GF at 1080p, 68 x 120): warp-level
//    mma.sync.m16n8k16 on fp16 halves with the error-compensated three-product split
//    a.b ~ a_hi.b_hi + (a_lo.b_hi + a_hi.b_lo)  (hi = fp16(x), lo = fp16(x - hi), every vector first normalised
//    into fp16 range by an exact power of two: what is dropped is ~2^-22 relative, fp32-level), fp32
//    accumulation, the online softmax on the accumulator fragments.  This is the legacy tensor path (HMMA), not
//    tcgen05: the op is a few hundred microseconds a few times per interaction, and a register-resident
//    flash-style softmax maps directly onto the mma.sync fragment layout.
//
// attention_partial_kernel: grid (ceil(n_query / 128), n_splits); thread = one query with its 64-channel key in
// registers; the memory keys of the split go through shared memory in 32-position tiles (every thread reads the
// same key element: a broadcast, conflict-free); online softmax (running max, denominator, C accumulators).
// attention_merge_kernel: combines the n_splits partial (max, denominator, accumulators) triples per query.
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"

namespace evavos {

namespace {

constexpr int kAttThreads = 128;
constexpr int kAttTile = 32;   // memory positions per shared-memory tile
constexpr int64_t kAttTensorMinScores = (int64_t)1 << 22;   // n_mem * n_query from which the tensor form runs (~2048^2;
                                                            // measured: equal at 1620^2, 1.6x faster at 2880^2, 2.3x at 8160^2)

template <int CP>   // accumulators held per thread (>= n_vec)
__global__ void __launch_bounds__(kAttThreads) attention_partial_kernel(
    const float* __restrict__ mk, int64_t mk_ch_stride, const float* __restrict__ qk, int64_t qk_ch_stride,
    const float* __restrict__ vec, int64_t vec_row_stride, int n_vec, int64_t n_mem, int64_t n_query, float scale,
    float* __restrict__ part) {
  pdl_wait();
  __shared__ __align__(16) float key_s[kAttTile][68];   // row stride 68: 16-byte rows, 4-way instead of 32-way store conflicts
  __shared__ float nrm_s[kAttTile];
  __shared__ float vec_s[CP][kAttTile];

  const int tid = threadIdx.x;
  const int64_t q = (int64_t)blockIdx.x * kAttThreads + tid;
  const int n_splits = gridDim.y, split = blockIdx.y;
  const int64_t n0 = (n_mem * split) / n_splits, n1 = (n_mem * (split + 1)) / n_splits;

  float qv[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) qv[c] = (q < n_query) ? __ldg(qk + (int64_t)c * qk_ch_stride + q) : 0.f;

  float run_max = -INFINITY, denom = 0.f;
  float acc[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) acc[c] = 0.f;

  for (int64_t t0 = n0; t0 < n1; t0 += kAttTile) {
    const int nt = (int)min((int64_t)kAttTile, n1 - t0);
    __syncthreads();
    // keys: channel-major in global (coalesced along positions), position-major in shared memory
    for (int e = tid; e < 64 * kAttTile; e += kAttThreads) {
      const int c = e / kAttTile, i = e % kAttTile;
      key_s[i][c] = (i < nt) ? __ldg(mk + (int64_t)c * mk_ch_stride + t0 + i) : 0.f;
    }
    for (int e = tid; e < CP * kAttTile; e += kAttThreads) {
      const int c = e / kAttTile, i = e % kAttTile;
      vec_s[c][i] = (c < n_vec && i < nt) ? __ldg(vec + (int64_t)c * vec_row_stride + t0 + i) : 0.f;
    }
    __syncthreads();
    if (tid < kAttTile) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) s = fmaf(key_s[tid][c], key_s[tid][c], s);
      nrm_s[tid] = s;
    }
    __syncthreads();

    for (int i = 0; i < nt; ++i) {
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      const float4* kr = reinterpret_cast<const float4*>(key_s[i]);
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 k4 = kr[c4];
        d0 = fmaf(k4.x, qv[4 * c4], d0);
        d1 = fmaf(k4.y, qv[4 * c4 + 1], d1);
        d2 = fmaf(k4.z, qv[4 * c4 + 2], d2);
        d3 = fmaf(k4.w, qv[4 * c4 + 3], d3);
      }
      const float s = (2.0f * ((d0 + d1) + (d2 + d3)) - nrm_s[i]) * scale;
      if (s > run_max) {   // rescale the running sums to the new maximum
        const float r = expf(run_max - s);
        denom *= r;
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] *= r;
        run_max = s;
      }
      const float w = expf(s - run_max);
      denom += w;
#pragma unroll
      for (int c = 0; c < CP; ++c) acc[c] = fmaf(w, vec_s[c][i], acc[c]);
    }
  }

  if (q < n_query) {
    float* dst = part + ((int64_t)split * n_query + q) * (CP + 2);
    dst[0] = run_max;
    dst[1] = denom;
#pragma unroll
    for (int c = 0; c < CP; ++c) dst[2 + c] = acc[c];
  }
}

// ---- tensor-core form ---------------------------------------------------------------------------------------------
constexpr int kTcThreads = 128;   // 4 warps x 16 queries
constexpr int kTcQueries = 64;
constexpr int kTcTile = 64;       // memory positions per shared-memory tile
constexpr int kTcLd = 72;         // row stride (words): 72 = 8 mod 32 -> a B-fragment load (4 pair-rows x 8 positions) hits 32 banks

// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 significant bits for elements near the vector's largest (every
// vector is first normalised so that its largest magnitude sits at 2^14, see range_factors); smaller elements keep an
// ABSOLUTE error below 2^-25 of that, far inside what an fp32 dot product of the same vectors carries.
__device__ __forceinline__ void split_pair(float xa, float xb, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(xa, xb);
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(xa - f.x, xb - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Exact power-of-two factors that move a vector whose largest magnitude is mx to [2^14, 2^15) - inside fp16 range with
// the low halves well above its subnormals - and back.
__device__ __forceinline__ void range_factors(float mx, float& down, float& up) {
  int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 127 - 14;
  e = e < -100 ? -100 : e;   // (zero / denormal vectors; the exponent field of mx is at most 254, so e <= 113)
  down = __uint_as_float((uint32_t)(127 - e) << 23);
  up = __uint_as_float((uint32_t)(127 + e) << 23);
}

// D (16x8, fp32) += A (16x16, row) * B (16x8, col), fp16 operands.  lane = 4 g + t; a register holds two k-slots:
//   a0 (g; 2t, 2t+1)  a1 (g+8; 2t, 2t+1)  a2 (g; 2t+8, 2t+9)  a3 (g+8; 2t+8, 2t+9);  b0 (2t, 2t+1; n=g)  b1 (2t+8, 2t+9; n=g)
//   d0 (g, 2t)  d1 (g, 2t+1)  d2 (g+8, 2t)  d3 (g+8, 2t+1)
// The contraction runs over channels, so which channel sits in which k-slot is free as long as A and B agree: step s
// puts channels (16s + t, 16s + t + 4) in slots (2t, 2t+1) and (16s + t + 8, 16s + t + 12) in (2t+8, 2t+9) - "pair-row"
// p = 8s + 4 half + t of the shared-memory tile - which makes the B loads bank-conflict-free.
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float ex2(float x) {   // 2^x, flush-to-zero: the weights of far-away positions are exactly 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const float* src, bool valid) {   // zero-fills when !valid
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// grid (ceil(n_query / 64), n_splits); a split is a run of whole 64-position tiles.  Rows of the MMA are queries (A =
// the warp's 16 query keys, hi and lo halves in registers for the whole kernel), columns are memory positions (B = the
// key tile in shared memory).  The raw fp32 tile of step k+1 arrives by cp.async in a staging buffer while step k is
// computed; between the steps the CTA splits it once into packed fp16 hi / lo pair-rows.  Products: hi.hi in one
// accumulator, lo.hi + hi.lo in another - a.b to ~2^-22, fp32-level.  A lane owns 2 queries x 2
// positions of every 16 x 8 score tile and keeps its own online-softmax state for them; the four lanes of a quad are
// merged once at the end.  Scores are kept in log2 units (scale2 = log2(e) / sqrt(CK)): the exponentials are bare ex2.
// part: as attention_partial_kernel, the max in log2 units.
constexpr int tc_ctas_per_sm(int cp) { return cp <= 4 ? 4 : 3; }   // (registers: 128 / 152 / 168 per thread)

template <int CP>
__global__ void __launch_bounds__(kTcThreads, tc_ctas_per_sm(CP)) attention_partial_tc_kernel(
    const float* __restrict__ mk, int64_t mk_ch_stride, const float* __restrict__ qk, int64_t qk_ch_stride,
    const float* __restrict__ vec, int64_t vec_row_stride, int n_vec, int64_t n_mem, int64_t n_query, float scale2,
    float* __restrict__ part) {
  pdl_wait();
  __shared__ __align__(16) float stage[64][kTcTile];          // raw keys of the NEXT tile, [channel][position]
  __shared__ __align__(16) float vec_stage[CP][kTcTile];
  __shared__ __align__(16) uint32_t key_hi[32][kTcLd];        // [pair-row][position]: two channels per word
  __shared__ __align__(16) uint32_t key_lo[32][kTcLd];
  __shared__ __align__(16) float vec_s[CP][kTcTile];
  __shared__ __align__(16) float nrm_s[2][kTcTile];           // |m|^2 * scale2, one half of the channels each
  __shared__ uint32_t tile_max[2];                            // bits of max |key| of a tile (non-negative floats order as uints)

  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t q0 = (int64_t)blockIdx.x * kTcQueries + (tid >> 5) * 16 + g, q1 = q0 + 8;
  const int n_splits = gridDim.y, split = blockIdx.y;
  const int64_t n_tiles = (n_mem + kTcTile - 1) / kTcTile;
  const int64_t tile0 = (n_tiles * split) / n_splits, tile1 = (n_tiles * (split + 1)) / n_splits;

  auto fetch = [&](int64_t tile) {   // this thread's share of a tile: one position, every other channel; its vec entries
    const int64_t t0 = tile * kTcTile;
    const int i = tid & (kTcTile - 1), ch0 = tid >> 6;
    const bool ok = t0 + i < n_mem;
    const float* src = mk + (ok ? t0 + i : 0);
#pragma unroll 8
    for (int r = 0; r < 32; ++r) cp_async4(&stage[ch0 + 2 * r][i], src + (int64_t)(ch0 + 2 * r) * mk_ch_stride, ok);
    for (int e = tid; e < CP * kTcTile; e += kTcThreads) {
      const int c = e / kTcTile, j = e % kTcTile;
      const bool okv = c < n_vec && t0 + j < n_mem;
      cp_async4(&vec_stage[c][j], vec + (okv ? (int64_t)c * vec_row_stride + t0 + j : 0), okv);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (tid < 2) tile_max[tid] = 0u;
  if (tile0 < tile1) fetch(tile0);

  // query fragments, each query scaled into fp16 range by its own power of two
  float xq[2][16];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t off = (int64_t)(16 * s + 4 * j + t) * qk_ch_stride;
      xq[0][4 * s + j] = (q0 < n_query) ? __ldg(qk + off + q0) : 0.f;
      xq[1][4 * s + j] = (q1 < n_query) ? __ldg(qk + off + q1) : 0.f;
    }
  uint32_t a_hi[4][4], a_lo[4][4];
  float two_scale_q[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fabsf(xq[h][j]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float down, up;
    range_factors(mx, down, up);
    two_scale_q[h] = 2.0f * scale2 * up;
#pragma unroll
    for (int s = 0; s < 4; ++s) {   // a0 / a2 (h = 0) or a1 / a3 (h = 1): channels (16s + t, + 4) and (16s + t + 8, + 12)
      split_pair(xq[h][4 * s] * down, xq[h][4 * s + 1] * down, a_hi[s][h], a_lo[s][h]);
      split_pair(xq[h][4 * s + 2] * down, xq[h][4 * s + 3] * down, a_hi[s][2 + h], a_lo[s][2 + h]);
    }
  }

  float run_max[2] = {-3.0e38f, -3.0e38f}, denom[2] = {0.f, 0.f};
  float acc[2][CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) acc[0][c] = acc[1][c] = 0.f;

  for (int64_t tile = tile0; tile < tile1; ++tile) {
    const int par = (int)(tile - tile0) & 1;
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // the staged tile is complete; every warp is done with key_hi / key_lo / vec_s of the previous tile
    // this thread's 16 channel pairs of one position (pair-rows of its parity), out of the staging buffer
    const int i = tid & (kTcTile - 1), ph = tid >> 6;
    float xa[16], xb[16], vreg[(CP * kTcTile + kTcThreads - 1) / kTcThreads];
    float mx = 0.f, ss = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int p = ph + 2 * r, c = 16 * (p >> 3) + 8 * ((p >> 2) & 1) + (p & 3);
      xa[r] = stage[c][i];
      xb[r] = stage[c + 4][i];
      mx = fmaxf(mx, fmaxf(fabsf(xa[r]), fabsf(xb[r])));
      ss = fmaf(xa[r], xa[r], fmaf(xb[r], xb[r], ss));
    }
#pragma unroll
    for (int r = 0; r < (CP * kTcTile + kTcThreads - 1) / kTcThreads; ++r)
      vreg[r] = (&vec_stage[0][0])[tid + r * kTcThreads];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) atomicMax(&tile_max[par], __float_as_uint(mx));
    if (tid == 0) tile_max[par ^ 1] = 0u;
    __syncthreads();   // staging buffer is in registers everywhere; the tile's maximum is known
    if (tile + 1 < tile1) fetch(tile + 1);
    float down, up;
    range_factors(__uint_as_float(tile_max[par]), down, up);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      uint32_t hi, lo;
      split_pair(xa[r] * down, xb[r] * down, hi, lo);
      key_hi[ph + 2 * r][i] = hi;
      key_lo[ph + 2 * r][i] = lo;
    }
    nrm_s[ph][i] = (tile * kTcTile + i < n_mem) ? ss * scale2 : 1.0e30f;   // past the end: score ~ -1e30, weight exactly 0
#pragma unroll
    for (int r = 0; r < (CP * kTcTile + kTcThreads - 1) / kTcThreads; ++r)
      (&vec_s[0][0])[tid + r * kTcThreads] = vreg[r];
    __syncthreads();
    const float ts0 = two_scale_q[0] * up, ts1 = two_scale_q[1] * up;

    for (int jp = 0; jp < kTcTile / 8; jp += 2) {   // two 16 x 8 score tiles at a time: four independent MMA chains
      float chh[2][4], cco[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int r = 0; r < 4; ++r) chh[u][r] = cco[u][r] = 0.f;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int n = (jp + u) * 8 + g;
          const uint32_t bh0 = key_hi[8 * s + t][n], bh1 = key_hi[8 * s + 4 + t][n];
          const uint32_t bl0 = key_lo[8 * s + t][n], bl1 = key_lo[8 * s + 4 + t][n];
          mma_f16(cco[u], a_lo[s], bh0, bh1);
          mma_f16(cco[u], a_hi[s], bl0, bl1);
          mma_f16(chh[u], a_hi[s], bh0, bh1);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int nb = (jp + u) * 8 + 2 * t;
        const float2 na = *reinterpret_cast<const float2*>(&nrm_s[0][nb]), nc = *reinterpret_cast<const float2*>(&nrm_s[1][nb]);
        const float nr0 = na.x + nc.x, nr1 = na.y + nc.y;
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // h = 0: query g (d0, d1), h = 1: query g + 8 (d2, d3)
          const float ts = h ? ts1 : ts0;
          const float sa = fmaf(chh[u][2 * h] + cco[u][2 * h], ts, -nr0);
          const float sb = fmaf(chh[u][2 * h + 1] + cco[u][2 * h + 1], ts, -nr1);
          const float mxs = fmaxf(sa, sb);
          if (mxs > run_max[h]) {
            const float r = ex2(run_max[h] - mxs);
            denom[h] *= r;
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[h][c] *= r;
            run_max[h] = mxs;
          }
          const float wa = ex2(sa - run_max[h]), wb = ex2(sb - run_max[h]);
          denom[h] += wa + wb;
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            const float2 v = *reinterpret_cast<const float2*>(&vec_s[c][nb]);
            acc[h][c] = fmaf(wb, v.y, fmaf(wa, v.x, acc[h][c]));
          }
        }
      }
    }
  }

  // merge the quad (the four lanes that share g hold disjoint positions of the same two queries)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float m = run_max[h];
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    const float r = ex2(run_max[h] - m);
    float d = denom[h] * r;
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      float a = acc[h][c] * r;
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      acc[h][c] = a;
    }
    const int64_t q = h ? q1 : q0;
    if (t == 0 && q < n_query) {
      float* dst = part + ((int64_t)split * n_query + q) * (CP + 2);
      dst[0] = m;
      dst[1] = d;
#pragma unroll
      for (int c = 0; c < CP; ++c) dst[2 + c] = acc[h][c];
    }
  }
}

template <int CP, bool LOG2 = false>
__global__ void __launch_bounds__(128) attention_merge_kernel(const float* __restrict__ part, int n_splits, int n_vec,
                                                              int64_t n_query, float* __restrict__ out,
                                                              int64_t out_row_stride) {
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_query) return;
  float m = -INFINITY;
  for (int s = 0; s < n_splits; ++s) m = fmaxf(m, part[((int64_t)s * n_query + q) * (CP + 2)]);
  float denom = 0.f;
  float acc[CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) acc[c] = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float* src = part + ((int64_t)s * n_query + q) * (CP + 2);
    const float r = LOG2 ? exp2f(src[0] - m) : expf(src[0] - m);   // exp(-inf) = 0 for an empty split
    denom = fmaf(src[1], r, denom);
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = fmaf(src[2 + c], r, acc[c]);
  }
  const float inv = 1.0f / denom;
#pragma unroll
  for (int c = 0; c < CP; ++c)
    if (c < n_vec) out[(int64_t)c * out_row_stride + q] = acc[c] * inv;
}

int pick_splits(int64_t n_mem, int64_t n_query, int n_sm) {
  const int64_t q_blocks = ceil_div(n_query, kAttThreads);
  int64_t s = ceil_div((int64_t)n_sm * 4, q_blocks);     // ~4 CTAs of 128 threads per SM
  const int64_t max_s = ceil_div(n_mem, kAttTile * 2);   // at least two tiles per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

int padded_vecs(int n_vec) { return n_vec <= 4 ? 4 : n_vec <= 8 ? 8 : n_vec <= 16 ? 16 : 32; }

template <int CP>
int launch_cp(const float* mk, int64_t mk_ch_stride, const float* qk, int64_t qk_ch_stride, const float* vec,
              int64_t vec_row_stride, int n_vec, int64_t n_mem, int64_t n_query, float scale, float* out,
              int64_t out_row_stride, float* part, int n_splits, cudaStream_t st) {
  const dim3 grid((unsigned)ceil_div(n_query, kAttThreads), (unsigned)n_splits);
  EVAVOS_CUDA_OK(launch_pdl(attention_partial_kernel<CP>, grid, dim3(kAttThreads), 0, st, mk, mk_ch_stride, qk,
                            qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, part));
  EVAVOS_CUDA_OK(launch_pdl(attention_merge_kernel<CP>, dim3((unsigned)ceil_div(n_query, 128)), dim3(128), 0, st,
                            (const float*)part, n_splits, n_vec, n_query, out, out_row_stride));
  return EVAVOS_OK;
}

int pick_splits_tc(int64_t n_mem, int64_t n_query, int n_sm, int ctas_per_sm) {
  // Splits are runs of whole tiles.  Take the split count that leaves the fewest CTA slots idle in the last wave, lightly
  // preferring fewer splits (each CTA sets up its query fragments once).
  const int64_t q_blocks = ceil_div(n_query, kTcQueries), slots = (int64_t)n_sm * ctas_per_sm;
  int64_t max_s = ceil_div(n_mem, kTcTile) / 2;          // at least two tiles per split
  if (max_s > 32) max_s = 32;
  int best = 1;
  double best_score = -1.0;
  for (int64_t s = 1; s <= (max_s < 1 ? 1 : max_s); ++s) {
    const int64_t ctas = q_blocks * s;
    const double score = (double)ctas / (double)(ceil_div(ctas, slots) * slots) - 0.005 * (double)s;
    if (score > best_score) { best_score = score; best = (int)s; }
  }
  return best;
}

// The tensor form pays off once the contraction dwarfs the launch and the per-CTA query-fragment set-up; below that the
// CUDA-core form stays.  EVAVOS_ATTENTION_PATH = tensor | simt overrides (tests run both forms on every shape).
bool use_tensor_form(int64_t n_mem, int64_t n_query) {
  const char* env = getenv("EVAVOS_ATTENTION_PATH");
  if (env && env[0] == 't') return true;
  if (env && env[0] == 's') return false;
  return n_mem * n_query >= kAttTensorMinScores;
}

template <int CP>
int launch_tc(const float* mk, int64_t mk_ch_stride, const float* qk, int64_t qk_ch_stride, const float* vec,
              int64_t vec_row_stride, int n_vec, int64_t n_mem, int64_t n_query, float scale2, float* out,
              int64_t out_row_stride, float* part, int n_splits, cudaStream_t st) {
  const dim3 grid((unsigned)ceil_div(n_query, kTcQueries), (unsigned)n_splits);
  EVAVOS_CUDA_OK(launch_pdl(attention_partial_tc_kernel<CP>, grid, dim3(kTcThreads), 0, st, mk, mk_ch_stride, qk,
                            qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale2, part));
  EVAVOS_CUDA_OK(launch_pdl(attention_merge_kernel<CP, true>, dim3((unsigned)ceil_div(n_query, 128)), dim3(128), 0, st,
                            (const float*)part, n_splits, n_vec, n_query, out, out_row_stride));
  return EVAVOS_OK;
}

}  // namespace

size_t attention_workspace_bytes(int n_vec, int64_t n_mem, int64_t n_query, int n_sm) {
  // (whichever form runs: the tensor form carries at most 16 rows per pass)
  const size_t simt = (size_t)pick_splits(n_mem, n_query, n_sm) * (size_t)(padded_vecs(n_vec) + 2);
  const int s3 = pick_splits_tc(n_mem, n_query, n_sm, 3), s4 = pick_splits_tc(n_mem, n_query, n_sm, 4);
  const size_t tc = (size_t)(s3 > s4 ? s3 : s4) * (size_t)(padded_vecs(n_vec < 16 ? n_vec : 16) + 2);
  return (simt > tc ? simt : tc) * (size_t)n_query * sizeof(float);
}

int launch_attention_readout(const float* mk, int64_t mk_ch_stride, const float* qk, int64_t qk_ch_stride,
                             const float* vec, int64_t vec_row_stride, int n_vec, int CK, int64_t n_mem,
                             int64_t n_query, float* out, int64_t out_row_stride, void* workspace, int n_sm,
                             cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)CK);
  float* part = reinterpret_cast<float*>(workspace);
  if (use_tensor_form(n_mem, n_query)) {
    const float scale2 = 1.4426950408889634f * scale;
    for (int r0 = 0; r0 < n_vec; r0 += 16) {   // 16 mask rows per pass (the softmax state lives in registers)
      const int rows = n_vec - r0 < 16 ? n_vec - r0 : 16;
      const int n_splits = pick_splits_tc(n_mem, n_query, n_sm, tc_ctas_per_sm(rows));
      const float* v = vec + (int64_t)r0 * vec_row_stride;
      float* o = out + (int64_t)r0 * out_row_stride;
      const int rc = rows <= 4
          ? launch_tc<4>(mk, mk_ch_stride, qk, qk_ch_stride, v, vec_row_stride, rows, n_mem, n_query, scale2, o, out_row_stride, part, n_splits, st)
          : rows <= 8
          ? launch_tc<8>(mk, mk_ch_stride, qk, qk_ch_stride, v, vec_row_stride, rows, n_mem, n_query, scale2, o, out_row_stride, part, n_splits, st)
          : launch_tc<16>(mk, mk_ch_stride, qk, qk_ch_stride, v, vec_row_stride, rows, n_mem, n_query, scale2, o, out_row_stride, part, n_splits, st);
      if (rc != EVAVOS_OK) return rc;
    }
    return EVAVOS_OK;
  }
  const int n_splits = pick_splits(n_mem, n_query, n_sm);
  switch (padded_vecs(n_vec)) {
    case 4: return launch_cp<4>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
    case 8: return launch_cp<8>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
    case 16: return launch_cp<16>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
    default: return launch_cp<32>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
  }
}

}  // namespace evavos
