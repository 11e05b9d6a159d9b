// Shared device/host helpers for libevavos_sm100 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/evavos.h"

namespace evavos {

constexpr int kTilePos = EVAVOS_TILE_POS;        // 128 positions per key tile image
constexpr int kTileKeyBytes = 128 * 128;         // 128 rows x 128 B (64 bf16)
constexpr int kTileBytes = EVAVOS_TILE_BYTES;    // + 128 rows x 32 B: bf16 (hi, mid, lo) split of -|k|^2/2, zero padded
constexpr float kEmptyNh = -1.0e30f;             // "-|k|^2/2" of an empty row: can never pass a threshold
constexpr int kCandCap = 1024;                   // (position, filter score) candidate slots per query handed to the finalizer
constexpr int kMaxSurvivors = 256;               // candidates that may survive the finalizer's own bound and get rescored

// ---- error plumbing (thread-local message behind evavos_last_error) -------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define EVAVOS_CUDA_OK(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::evavos::cuda_fail(_e, #expr); \
  } while (0)

// Programmatic dependent launch: a kernel launched through launch_pdl() may be scheduled while the previous kernel
// of the stream drains (its launch latency overlaps that kernel's tail); it must call pdl_wait() before touching
// anything the previous kernel wrote - the wait returns once that grid has completed and its writes are visible.
// pdl_launch_dependents() lets the NEXT kernel start being scheduled early.  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- order-preserving float <-> uint key ------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

__device__ __forceinline__ uint32_t ordered_to_float_bits(uint32_t k) {
  return (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
}

// Byte offset of (row r, 16-byte chunk c) inside a 128B-swizzled K-major tile (Swizzle<3,4,3>).
__host__ __device__ __forceinline__ int swizzle128_offset(int r, int c) {
  return r * 128 + (((c ^ (r & 7)) & 7) << 4);
}

// Byte offset of (row r, 16-byte chunk c in {0,1}) inside a 32B-swizzled K-major tile (Swizzle<1,4,3>).
__host__ __device__ __forceinline__ int swizzle32_offset(int r, int c) {
  return r * 32 + (((c ^ (r >> 2)) & 1) << 4);
}

// Exact fp32 affinity of one (query, key) pair, the arithmetic every selection path agrees on:
// (-|k|^2 + 2 k.q - |q|^2) / sqrt(CK), accumulated channel by channel with FMAs
// (prop_net.py:86-90 evaluates the same expression with an SGEMM).
__device__ __forceinline__ float affinity_from_parts(float kk, float kq, float qq, float inv_sqrt_ck) {
  float t = -kk + 2.0f * kq;
  t = t - qq;
  return t * inv_sqrt_ck;
}

// Error bound of the bf16 filter score S' = q^.k^ - |k|^2/2 against its exact value:
// |q^.k^ - q.k| <= 2^-8 (1 + 2^-10) |q||k| for bf16 round-to-nearest operands, plus slack for the tensor-core fp32
// accumulation and the rounding of -|k|^2/2.  The filter and the finalizer both subtract 2 eps from their bounds.
__device__ __forceinline__ float filter_eps(float q_norm, float key_maxnorm) {
  const float qn = q_norm * 1.0001f;
  return 0.004f * qn * key_maxnorm + 2.0e-6f * key_maxnorm * key_maxnorm + 1.0e-30f;
}

// ---- launchers implemented in the .cu files -------------------------------------------------
int launch_write_keys(const EvavosBankShadow& b, const float* src, int64_t src_ch_stride, int64_t pos0,
                      int64_t n_pos, float* dst_ref, int64_t dst_ref_ch_stride, cudaStream_t st);
int launch_write_values(const EvavosBankShadow& b, const float* src, int64_t src_obj_stride,
                        int64_t src_ch_stride, int64_t pos0, int64_t n_pos, float* dst_ref,
                        int64_t dst_ref_obj_stride, int64_t dst_ref_ch_stride, cudaStream_t st);

struct SelectBuffers {
  float* class_max;   // [G][MT*128][128]
  float* tau;         // [nq_pad]
  int32_t* cand_cnt;  // [nq_pad]
  int2* cand;         // [nq_pad][kCandCap] (position, filter score bits; the exact SIMT selection leaves the score 0)
  void* strip;        // phase-B staging strips of the filter's epilogue threads (score_pass_strip_bytes)
  unsigned int* grid_counter;   // [MT + 1], zeroed by launch_score_select; the last word is overflow_cnt
  int32_t* overflow_list;       // [nq_pad] queries whose list overflowed (written by the finalizer)
  unsigned int* overflow_cnt;   // their number
};

// The query is always addressed in the caller's layout: element (c, q) at query[c * query_ch_stride + q].
int launch_brute_select(const float* key_pm, const float* query, int64_t query_ch_stride, int CK, int64_t n_pos,
                        int64_t n_query, int top_k, int2* cand, int32_t* cand_cnt, int n_sm, cudaStream_t st);
// scored != 0: the entries carry filter scores (tcgen05 path) and the list is first cut down with the bound the
// scores themselves give; key_maxnorm is only read then.
int launch_finalize(const float* key_pm, const float* query, int64_t query_ch_stride, int CK, int64_t n_pos,
                    int64_t n_query, int top_k, const int2* cand, const int32_t* cand_cnt, int scored,
                    const float* key_maxnorm, int32_t* out_idx, float* out_weight, float* out_score,
                    const EvavosPeers* peers, int64_t peer_gather_offset, int32_t* overflow_list,
                    unsigned int* overflow_cnt, uint32_t* overflow_hint, cudaStream_t st);
// Exact selection + finalization of the queries the finalizer listed as overflowed (select_dense.cu); a no-op
// (one empty launch) when there are none.
int launch_overflow_exact(const float* key_pm, const float* query, int64_t query_ch_stride, int64_t n_pos,
                          int64_t n_query, int top_k, int2* cand, const int32_t* overflow_list,
                          const unsigned int* overflow_cnt, const float* key_maxnorm, int32_t* out_idx,
                          float* out_weight, float* out_score, const EvavosPeers* peers, int64_t peer_gather_offset,
                          int n_sm, cudaStream_t st);
int launch_bias_residual(void* y, const float* bias, const void* r, int64_t rows, int C, int bf16, int relu, cudaStream_t st);
int launch_upsample2x_add(void* y, const float* bias, const void* x, int64_t n, int H, int W, int C, int bf16, cudaStream_t st);
size_t jf_workspace_bytes(int64_t T, int h, int w);
int launch_jf_metrics(const uint8_t* pred, const uint8_t* gt, int64_t T, int h, int w, int radius, void* workspace,
                      double* out, int32_t* gt_empty, cudaStream_t st);
int launch_peer_barrier(const EvavosPeers& peers, int64_t flag_offset, uint32_t epoch, cudaStream_t st);
int launch_peer_reduce_scatter(const EvavosPeers& peers, int64_t partial_offset, int rows, int64_t q0, int64_t q1,
                               float* out, int64_t out_row_stride, cudaStream_t st);
int launch_score_select(const float* query, int64_t query_ch_stride, const void* key_tiles, const float* key_maxnorm,
                        int64_t n_pos, int64_t n_query, int top_k, int n_chunks, int sample_stride, int n_sm,
                        float* class_max, float* tau, int2* cand, int32_t* cand_cnt, void* strip,
                        unsigned int* grid_counter, cudaStream_t st);
size_t score_pass_strip_bytes(int64_t n_query, int n_chunks, int n_sm);
int score_pass_chunks(int64_t n_pos, int64_t n_query, int n_sm);
int64_t score_pass_mtiles(int64_t n_query);
int score_pass_sample_stride(int64_t n_pos, int n_chunks, int requested);

int launch_readout(const EvavosBankShadow& b, const int32_t* idx, const float* weight, int64_t n_query,
                   int top_k, float* out, int64_t out_obj_stride, int64_t out_ch_stride, int q_per_frame,
                   int64_t frame_stride, cudaStream_t st);
int launch_affinity_dense(const int32_t* idx, const float* weight, int64_t n_query, int top_k, int64_t n_pos,
                          float* dense, cudaStream_t st);
int launch_aggregate(const float* prob, float* out, int K, int64_t npix, int keep_bg, int hard,
                     cudaStream_t st);
int launch_argmax_unpad(const float* prob, int C, int64_t T, int nh, int nw, uint8_t* masks, uint8_t* out,
                        int pad_top, int pad_left, int h, int w, cudaStream_t st);
size_t attention_workspace_bytes(int n_vec, int64_t n_mem, int64_t n_query, int n_sm);
int launch_attention_readout(const float* mk, int64_t mk_ch_stride, const float* qk, int64_t qk_ch_stride,
                             const float* vec, int64_t vec_row_stride, int n_vec, int CK, int64_t n_mem,
                             int64_t n_query, float* out, int64_t out_row_stride, void* workspace, int n_sm,
                             cudaStream_t st);
int launch_topk_merge(const int32_t* cand_idx, const float* cand_score, int64_t n_query, int n_cand, int top_k,
                      int shard, int n_shards, int64_t pos_per_frame, int32_t* out_idx, float* out_weight,
                      float* out_score, int32_t* local_idx, int gathered, cudaStream_t st);

}  // namespace evavos
