// Full-softmax attention read of ONE memory frame (AttentionMemory + the two vector-matrix products of
// get_attention, mivos/model/propagation/prop_net.py:117-138, 198-211), flash style: the (HW x HW) softmax matrix W
// is never materialised.
//
//   out[c][q] = sum_n vec[c][n] * softmax_n( (-|m_n|^2 + 2 m_n.q_q - |q_q|^2) / sqrt(CK) )
//
// The -|q|^2 term is constant along n and cancels in the softmax.  Everything is fp32 on CUDA cores (the
// contraction is 2*HW*HW*CK = 0.34 GF at 480p: launch-latency territory, not tensor-core territory).
//
// attention_partial_kernel: grid (ceil(n_query / 128), n_splits); thread = one query with its 64-channel key in
// registers; the memory keys of the split go through shared memory in 32-position tiles (every thread reads the
// same key element: a broadcast, conflict-free); online softmax (running max, denominator, C accumulators).
// attention_merge_kernel: combines the n_splits partial (max, denominator, accumulators) triples per query.
#include "common.cuh"

namespace evavos {

namespace {

constexpr int kAttThreads = 128;
constexpr int kAttTile = 32;   // memory positions per shared-memory tile

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

template <int CP>
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
    const float r = expf(src[0] - m);   // exp(-inf) = 0 for an empty split
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

}  // namespace

size_t attention_workspace_bytes(int n_vec, int64_t n_mem, int64_t n_query, int n_sm) {
  return (size_t)pick_splits(n_mem, n_query, n_sm) * (size_t)n_query * (padded_vecs(n_vec) + 2) * sizeof(float);
}

int launch_attention_readout(const float* mk, int64_t mk_ch_stride, const float* qk, int64_t qk_ch_stride,
                             const float* vec, int64_t vec_row_stride, int n_vec, int CK, int64_t n_mem,
                             int64_t n_query, float* out, int64_t out_row_stride, void* workspace, int n_sm,
                             cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)CK);
  const int n_splits = pick_splits(n_mem, n_query, n_sm);
  float* part = reinterpret_cast<float*>(workspace);
  switch (padded_vecs(n_vec)) {
    case 4: return launch_cp<4>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
    case 8: return launch_cp<8>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
    case 16: return launch_cp<16>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
    default: return launch_cp<32>(mk, mk_ch_stride, qk, qk_ch_stride, vec, vec_row_stride, n_vec, n_mem, n_query, scale, out, out_row_stride, part, n_splits, st);
  }
}

}  // namespace evavos
