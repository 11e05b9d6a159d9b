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
//    mma.sync.m16n8k8 TF32 with the error-compensated three-product split  a.b ~ a_hi.b_hi + a_lo.b_hi + a_hi.b_lo
//    (hi = tf32(x), lo = tf32(x - hi): what is dropped is ~2^-22 relative, fp32-level), fp32 accumulation, the online
//    softmax on the accumulator fragments.  This is the legacy tensor path (HMMA), not tcgen05: the op is a few
//    hundred microseconds a few times per interaction, and a register-resident flash-style softmax maps directly
//    onto the mma.sync fragment layout.
//
// attention_partial_kernel: grid (ceil(n_query / 128), n_splits); thread = one query with its 64-channel key in
// registers; the memory keys of the split go through shared memory in 32-position tiles (every thread reads the
// same key element: a broadcast, conflict-free); online softmax (running max, denominator, C accumulators).
// attention_merge_kernel: combines the n_splits partial (max, denominator, accumulators) triples per query.
#include <cstdlib>

#include "common.cuh"

namespace evavos {

namespace {

constexpr int kAttThreads = 128;
constexpr int kAttTile = 32;   // memory positions per shared-memory tile
constexpr int64_t kAttTensorMinScores = (int64_t)1 << 24;   // n_mem * n_query from which the tensor form runs (~4096^2)

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
constexpr int kTcLd = 72;         // row stride (words): 72 = 8 mod 32 -> the B-fragment loads (4 channels x 8 positions) hit 32 banks

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// D (16x8, fp32) += A (16x8, row) * B (8x8, col).  lane = 4 g + t:
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);  b0 (k=t, n=g)  b1 (k=t+4, n=g);  d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (ceil(n_query / 64), n_splits).  Rows of the MMA are queries (A = the warp's 16 query keys, hi and lo parts in
// registers for the whole kernel), columns are memory positions (B = the key tile in shared memory, split into hi / lo
// once by the loading threads).  A lane owns 2 queries x 2 positions of every 16 x 8 score tile and keeps its own
// online-softmax state for them; the four lanes of a quad are merged once at the end.  Scores are kept in log2 units
// (scale2 = log2(e) / sqrt(CK)), so the exponentials are bare ex2.  part: like attention_partial_kernel, max in log2 units.
template <int CP>
__global__ void __launch_bounds__(kTcThreads, 3) attention_partial_tc_kernel(
    const float* __restrict__ mk, int64_t mk_ch_stride, const float* __restrict__ qk, int64_t qk_ch_stride,
    const float* __restrict__ vec, int64_t vec_row_stride, int n_vec, int64_t n_mem, int64_t n_query, float scale2,
    float* __restrict__ part) {
  pdl_wait();
  __shared__ __align__(16) uint32_t key_hi[64][kTcLd];
  __shared__ __align__(16) uint32_t key_lo[64][kTcLd];
  __shared__ __align__(16) float vec_s[CP][kTcTile];
  __shared__ __align__(16) float nrm_s[2][kTcTile];     // |m|^2 * scale2, summed over the even / odd channels

  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t q0 = (int64_t)blockIdx.x * kTcQueries + (tid >> 5) * 16 + g, q1 = q0 + 8;
  const int n_splits = gridDim.y, split = blockIdx.y;
  const int64_t n0 = (n_mem * split) / n_splits, n1 = (n_mem * (split + 1)) / n_splits;

  uint32_t a_hi[8][4], a_lo[8][4];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = 8 * s + t + ((r & 2) ? 4 : 0);
      const int64_t q = (r & 1) ? q1 : q0;
      const float x = (q < n_query) ? __ldg(qk + (int64_t)c * qk_ch_stride + q) : 0.f;
      a_hi[s][r] = to_tf32(x);
      a_lo[s][r] = to_tf32(x - __uint_as_float(a_hi[s][r]));
    }
  }

  float run_max[2] = {-3.0e38f, -3.0e38f}, denom[2] = {0.f, 0.f};
  float acc[2][CP];
#pragma unroll
  for (int c = 0; c < CP; ++c) acc[0][c] = acc[1][c] = 0.f;
  const float two_scale = 2.0f * scale2;

  for (int64_t t0 = n0; t0 < n1; t0 += kTcTile) {
    const int nt = (int)min((int64_t)kTcTile, n1 - t0);
    __syncthreads();
    {  // a thread fills one position of the tile: its even or its odd channels (coalesced along positions)
      const int i = tid & (kTcTile - 1), ch0 = tid >> 6;
      float ss = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int c = ch0 + 2 * r;
        const float x = (i < nt) ? __ldg(mk + (int64_t)c * mk_ch_stride + t0 + i) : 0.f;
        const uint32_t hi = to_tf32(x);
        key_hi[c][i] = hi;
        key_lo[c][i] = to_tf32(x - __uint_as_float(hi));
        ss = fmaf(x, x, ss);
      }
      nrm_s[ch0][i] = (i < nt) ? ss * scale2 : 1.0e30f;   // positions past the split: score ~ -1e30, weight exactly 0
    }
    for (int e = tid; e < CP * kTcTile; e += kTcThreads) {
      const int c = e / kTcTile, i = e % kTcTile;
      vec_s[c][i] = (c < n_vec && i < nt) ? __ldg(vec + (int64_t)c * vec_row_stride + t0 + i) : 0.f;
    }
    __syncthreads();

    for (int jp = 0; jp < kTcTile / 8; jp += 2) {   // two 16 x 8 score tiles at a time: four independent MMA chains
      float chh[2][4], cco[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int r = 0; r < 4; ++r) chh[u][r] = cco[u][r] = 0.f;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int n = (jp + u) * 8 + g;
          const uint32_t bh0 = key_hi[8 * s + t][n], bh1 = key_hi[8 * s + t + 4][n];
          const uint32_t bl0 = key_lo[8 * s + t][n], bl1 = key_lo[8 * s + t + 4][n];
          mma_tf32(cco[u], a_lo[s], bh0, bh1);
          mma_tf32(cco[u], a_hi[s], bl0, bl1);
          mma_tf32(chh[u], a_hi[s], bh0, bh1);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int nb = (jp + u) * 8 + 2 * t;
        const float2 na = *reinterpret_cast<const float2*>(&nrm_s[0][nb]), nc = *reinterpret_cast<const float2*>(&nrm_s[1][nb]);
        const float nr0 = na.x + nc.x, nr1 = na.y + nc.y;
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // h = 0: query g (d0, d1), h = 1: query g + 8 (d2, d3)
          const float sa = fmaf(chh[u][2 * h] + cco[u][2 * h], two_scale, -nr0);
          const float sb = fmaf(chh[u][2 * h + 1] + cco[u][2 * h + 1], two_scale, -nr1);
          const float mx = fmaxf(sa, sb);
          if (mx > run_max[h]) {
            const float r = exp2f(run_max[h] - mx);
            denom[h] *= r;
#pragma unroll
            for (int c = 0; c < CP; ++c) acc[h][c] *= r;
            run_max[h] = mx;
          }
          const float wa = exp2f(sa - run_max[h]), wb = exp2f(sb - run_max[h]);
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
    const float r = exp2f(run_max[h] - m);
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

int pick_splits_tc(int64_t n_mem, int64_t n_query, int n_sm) {
  const int64_t q_blocks = ceil_div(n_query, kTcQueries);
  int64_t s = ceil_div((int64_t)n_sm * 3, q_blocks);     // 3 CTAs of 128 threads per SM (registers)
  const int64_t max_s = ceil_div(n_mem, kTcTile * 2);    // at least two tiles per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
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
  // (whichever form runs: the tensor form carries at most 8 rows per pass)
  const size_t simt = (size_t)pick_splits(n_mem, n_query, n_sm) * (size_t)(padded_vecs(n_vec) + 2);
  const size_t tc = (size_t)pick_splits_tc(n_mem, n_query, n_sm) * (size_t)(padded_vecs(n_vec < 8 ? n_vec : 8) + 2);
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
    const int n_splits = pick_splits_tc(n_mem, n_query, n_sm);
    for (int r0 = 0; r0 < n_vec; r0 += 8) {   // 8 mask rows per pass (the softmax state lives in registers)
      const int rows = n_vec - r0 < 8 ? n_vec - r0 : 8;
      const float* v = vec + (int64_t)r0 * vec_row_stride;
      float* o = out + (int64_t)r0 * out_row_stride;
      const int rc = rows <= 4
          ? launch_tc<4>(mk, mk_ch_stride, qk, qk_ch_stride, v, vec_row_stride, rows, n_mem, n_query, scale2, o, out_row_stride, part, n_splits, st)
          : launch_tc<8>(mk, mk_ch_stride, qk, qk_ch_stride, v, vec_row_stride, rows, n_mem, n_query, scale2, o, out_row_stride, part, n_splits, st);
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
