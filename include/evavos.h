/*
 * evavos.h - C ABI of libevavos_sm100.so: the B200 (sm_100a) space-time memory read.
 *
 * This is the drop-in boundary for the one hot path of EVA-VOS / MiVOS / STCN
 * (paths below are relative to the reference repository root):
 *
 *   evavos_memread          replaces EvalMemoryReader.get_affinity + readout
 *                           (mivos/model/propagation/prop_net.py:80-115, called from
 *                           PropagationNetwork.segment_with_query, prop_net.py:179-187)
 *                           and softmax_w_g_top (prop_net.py:46-72).
 *   evavos_readout          replaces EvalMemoryReader.readout alone (prop_net.py:108-115).
 *   evavos_affinity_dense   materialises the dense (N,HW) matrix that get_affinity
 *                           returns in the reference (prop_net.py:60), on demand only.
 *   evavos_aggregate_wbg    replaces aggregate_wbg (mivos/model/aggregate.py:22-37).
 *   evavos_bank_write_keys / evavos_bank_write_values
 *                           replace the in-place bank append and the certain-memory
 *                           copy of InferenceCore.do_pass (mivos/inference_core.py:150-155,
 *                           174-177) and the torch.cat growth in interact (:235-240).
 *   evavos_argmax_unpad     replaces the per-frame argmax + un-padding at the end of interact
 *                           (mivos/inference_core.py:247-257).
 *   evavos_attention_readout replaces AttentionMemory.forward + the two vector-matrix products of
 *                           get_attention (mivos/model/propagation/prop_net.py:117-138, 204-207) without
 *                           materialising the (HW, HW) softmax matrix.
 *   evavos_jf_metrics       replaces the per-frame J / J&F scoring of interactions/eval.py:27-81 and
 *                           interactions/metrics.py:9-160.
 *   evavos_topk_merge       the exchange step of the memory-axis sharded read
 *                           (no reference counterpart); evavos_peer_barrier,
 *                           evavos_peer_reduce_scatter and the `peers` field of the read move its data over
 *                           NVLink peer memory from inside the kernels.
 *   evavos_memread_host     the same read with HOST buffers in the reference layout
 *                           (what a ctypes/cgo caller without device memory would bind).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - all device pointers are owned by the caller (PyTorch's caching allocator in the
 *     Python host); the library never allocates or frees device memory except inside
 *     evavos_memread_host, which owns its scratch for the duration of the call.
 *   - every device entry point is asynchronous and ordered on the cudaStream_t passed
 *     as `stream` (an opaque pointer here so the header needs no CUDA include).
 *   - return value: 0 on success, a negative EVAVOS_ERR_* code otherwise; a
 *     thread-local message is available from evavos_last_error().
 *   - there is no CPU fallback: on a machine without an sm_100 device the entry
 *     points return EVAVOS_ERR_CUDA.
 *
 * Memory-bank layouts (the API contract, mivos/inference_core.py:150-151)
 *   keys   (1, CK, T, H, W) fp32     values (K, CV, T, H, W) fp32
 * described here by (pointer, channel stride, object stride) because do_pass reads
 * T-slices `[:, :, :m_front]` of a pre-allocated bank (strided views, :167-168).
 *
 * Engine-private shadow of a bank ("position-major"), maintained by the two
 * evavos_bank_write_* calls and consumed by the read:
 *   key_pm     [n_pos][CK] fp32            exact keys, one 4*CK-byte row per position
 *   key_tiles  ceil(n_pos/128) tile images of EVAVOS_TILE_BYTES each: 128 bf16 key rows
 *              in the 128B-swizzled K-major shared-memory layout tcgen05.mma consumes,
 *              followed by a 16-wide extra K slice per row (bf16 hi/mid/lo split of -|k|^2/2,
 *              32B-swizzled) so the norm term is part of the contraction; one tile image is
 *              loaded verbatim by one bulk TMA copy
 *   key_maxnorm  1 fp32: max |k| over the bank (rigorous bf16 error bound for the filter)
 *   val_pm     [K][n_pos][CV] fp32 or bf16 value rows
 */
#ifndef EVAVOS_H_
#define EVAVOS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVAVOS_ABI_VERSION 2

#define EVAVOS_OK 0
#define EVAVOS_ERR_INVALID (-1)     /* bad argument (null pointer, non-positive size, ...) */
#define EVAVOS_ERR_UNSUPPORTED (-2) /* shape / dtype outside what the kernels implement   */
#define EVAVOS_ERR_CUDA (-3)        /* a CUDA runtime call or kernel launch failed        */
#define EVAVOS_ERR_WORKSPACE (-4)   /* caller-provided workspace too small                */
#define EVAVOS_ERR_TOPK_RANGE (-5)  /* n_pos < top_k: "selected index k out of range" (prop_net.py:53) */

#define EVAVOS_F32 0
#define EVAVOS_BF16 1

#define EVAVOS_PATH_AUTO 0   /* tcgen05 filter when CK == 64, else SIMT                      */
#define EVAVOS_PATH_TENSOR 1 /* tcgen05/TMEM candidate filter + exact fp32 rescoring          */
#define EVAVOS_PATH_SIMT 2   /* exact fp32 CUDA-core radix select                             */
#define EVAVOS_PATH_TENSOR_DENSE 3 /* TENSOR, and the exact tiled pass for overflowed lists is always launched.
                                      AUTO / TENSOR launch it only when the previous read on this device saw an
                                      overflow (a host-visible hint the kernels set; until then an overflowed query
                                      is redone by one warp inside the finalizer: same result, slower)          */

#define EVAVOS_TILE_POS 128      /* memory positions per key tile image  */
#define EVAVOS_TILE_BYTES 20480  /* 128 rows x 128 B (bf16 keys, CK=64) + 128 rows x 32 B (-|k|^2/2) */
#define EVAVOS_MAX_TOPK 128

typedef void* evavos_stream_t; /* cudaStream_t */

/* Engine-private shadow of a memory bank (all device pointers). */
typedef struct EvavosBankShadow {
  float* key_pm;       /* [capacity_pos][CK] fp32                                   */
  void* key_tiles;     /* ceil(capacity_pos/128) * EVAVOS_TILE_BYTES, 16B aligned; NULL when CK != 64 */
  float* key_maxnorm;  /* 1 fp32, must be zero-initialised by the caller             */
  void* val_pm;        /* [K][capacity_pos][CV] of val_dtype                         */
  int64_t capacity_pos; /* positions the buffers were sized for                      */
  int32_t K;
  int32_t CK;
  int32_t CV;
  int32_t val_dtype;   /* EVAVOS_F32 | EVAVOS_BF16                                   */
} EvavosBankShadow;

#define EVAVOS_MAX_RANKS 16

/*
 * Exchange buffers of the ranks of a memory-axis sharded bank (one process per GPU): base[g] is rank g's buffer as
 * mapped into THIS process (CUDA IPC / peer access over NVLink; base[rank] is the local one).  The kernels below
 * store to and load from these pointers directly - the exchange steps of the sharded read are device-initiated,
 * no host-side collective is on the data path.
 */
typedef struct EvavosPeers {
  int32_t n_ranks;
  int32_t rank;
  void* base[EVAVOS_MAX_RANKS];
} EvavosPeers;

/* Arguments of the fused read. */
typedef struct EvavosMemReadArgs {
  EvavosBankShadow bank;
  const float* query;      /* (CK, n_query) fp32, row stride query_ch_stride; n_query = HW * frames */
  float* readout;          /* (K, CV, n_query) fp32 or NULL to skip the readout       */
  int32_t* topk_idx;       /* (n_query, top_k) memory positions, best first; or NULL  */
  float* topk_weight;      /* (n_query, top_k) softmax weights; or NULL               */
  float* topk_score;       /* (n_query, top_k) affinities (-a+b-c)/sqrt(CK); or NULL  */
  void* workspace;
  int64_t workspace_bytes;
  int64_t n_pos;           /* N = m_front * H * W positions to read (<= capacity_pos) */
  int64_t n_query;
  int64_t query_ch_stride;
  int64_t readout_obj_stride; /* elements between objects in `readout` (0 -> CV*n_query) */
  int64_t readout_ch_stride;  /* elements between channels in `readout` (0 -> n_query).  1 = channels-last (NHWC)
                                 destination: element (object o, channel c, position p) of a frame at
                                 o * readout_obj_stride + p * P + c with the position stride
                                 P = readout_obj_stride / (queries_per_frame or n_query) >= CV, P % 4 == 0
                                 (fp32 CV % 128 == 0 / bf16 CV % 256 == 0 banks only)              */
  int32_t top_k;
  int32_t path;            /* EVAVOS_PATH_*                                           */
  int32_t n_sm;            /* SM count to size grids for (0 -> query the device)      */
  int32_t sample_stride;   /* tensor path: the threshold pass contracts every sample_stride-th key tile
                              (0 -> the library's choice; 1 = two full sweeps; clamped for short banks) */
  /* Sharded read (optional, NULL otherwise): the finalizer also stores this rank's per-query list as packed
     (LOCAL position or -1, score bits) int32 pairs into EVERY rank's buffer at byte offset peer_gather_offset,
     laid out [n_ranks][n_query][top_k][2] with this rank's rows in slot `rank` - the all-gather of the sharded
     read, done by the epilogue of the kernel that produces the lists (evavos_topk_merge_gathered consumes it). */
  const EvavosPeers* peers;
  int64_t peer_gather_offset;
  /* Several query frames in one launch (n_query = frames * queries_per_frame), each frame with its own destination
     block: query q = f * queries_per_frame + p of (object o, channel c) is written to
     readout[f * readout_frame_stride + o * readout_obj_stride + c * readout_ch_stride + p] - e.g. straight into the
     decoder's (F, K, 2*CV, H, W) input (prop_net.py:189-190 without the cat).  0 -> one block, offset q. */
  int64_t queries_per_frame;
  int64_t readout_frame_stride;
} EvavosMemReadArgs;

int evavos_abi_version(void);
const char* evavos_last_error(void);
size_t evavos_sizeof_bank_shadow(void);
size_t evavos_sizeof_memread_args(void);

/* Bytes of key_tiles needed for `capacity_pos` positions. */
size_t evavos_key_tiles_bytes(int64_t capacity_pos);

/*
 * Write `n_pos` key positions starting at bank position `pos0`.
 *   src: (CK, n_pos) fp32 with row stride src_ch_stride (a new frame k16 of
 *        inference_core.py:175, or a T-slice of an existing bank).
 *   dst_ref: optional reference-layout bank keys (1,CK,T,H,W); element
 *        [c*dst_ref_ch_stride + pos0 + i] receives src[c][i].  NULL to skip.
 * Updates key_pm, key_tiles (if non-NULL) and key_maxnorm of `bank`.  Only the rows of the written range
 * change: any slot may be rewritten in place (the read ignores tile rows at or beyond its n_pos).
 */
int evavos_bank_write_keys(const EvavosBankShadow* bank, const float* src, int64_t src_ch_stride,
                           int64_t pos0, int64_t n_pos, float* dst_ref, int64_t dst_ref_ch_stride,
                           evavos_stream_t stream);

/*
 * Write `n_pos` value positions starting at `pos0` for all K objects.
 *   src: (K, CV, n_pos) fp32, strides src_obj_stride / src_ch_stride (inference_core.py:176).
 *   dst_ref: optional reference-layout bank values (K,CV,T,H,W).
 */
int evavos_bank_write_values(const EvavosBankShadow* bank, const float* src, int64_t src_obj_stride,
                             int64_t src_ch_stride, int64_t pos0, int64_t n_pos, float* dst_ref,
                             int64_t dst_ref_obj_stride, int64_t dst_ref_ch_stride, evavos_stream_t stream);

size_t evavos_memread_workspace_bytes(const EvavosMemReadArgs* args);
int evavos_memread(const EvavosMemReadArgs* args, evavos_stream_t stream);
/*
 * Diagnostics: how many queries of the LAST evavos_memread that ran with these arguments (same workspace, sizes and
 * path) had more candidates inside the filter's error margin than a list holds and were selected by the exact tiled
 * pass instead (select_dense.cu).  Synchronises `stream`.  0 on the SIMT path.  No reference counterpart.
 */
int evavos_memread_overflow_count(const EvavosMemReadArgs* args, uint32_t* count, evavos_stream_t stream);

/* Sparse readout: out[o][c][q] = sum_j weight[q][j] * val_pm[o][idx[q][j]][c]; idx < 0 entries are skipped. */
int evavos_readout(const EvavosBankShadow* bank, const int32_t* idx, const float* weight, int64_t n_query,
                   int32_t top_k, float* out, int64_t out_obj_stride, int64_t out_ch_stride,
                   evavos_stream_t stream);

/* dense[n][q] = weight of position n for query q, zeros elsewhere: (n_pos, n_query) fp32 (prop_net.py:60). */
int evavos_affinity_dense(const int32_t* idx, const float* weight, int64_t n_query, int32_t top_k,
                          int64_t n_pos, float* dense, evavos_stream_t stream);

/* Soft aggregation: prob (K, npix) fp32 -> out (K+1, npix) if keep_bg else (K, npix). */
int evavos_aggregate_wbg(const float* prob, float* out, int32_t K, int64_t npix, int32_t keep_bg,
                         int32_t hard, evavos_stream_t stream);

/*
 * Fused elementwise tails of the decoder's convolutions (prop_net.py:13-30; modules.py ResBlock / UpsampleBlock), on
 * channels-last tensors of `dtype` EVAVOS_F32 or EVAVOS_BF16, in place on `y`:
 *   evavos_bias_residual_nhwc   y[row][c] = [relu](y + bias[c] (+ residual[row][c]))
 *   evavos_upsample2x_add_nhwc  y[n][h][w][c] = y + bias[c] + bilinear_x2(x)[n][h][w][c], x of shape (n, H/2, W/2, C),
 *                               F.interpolate(scale_factor=2, mode="bilinear", align_corners=False)
 * bias is fp32 (C); all pointers 16-byte aligned; C % 4 == 0 (fp32) / C % 8 == 0 (bf16).  They replace the separate
 * broadcast-add / add / upsample kernels PyTorch launches after F.conv2d (no reference counterpart beyond those ops).
 */
int evavos_bias_residual_nhwc(void* y, const float* bias, const void* residual, int64_t rows, int32_t C, int32_t dtype,
                              int32_t relu, evavos_stream_t stream);
int evavos_upsample2x_add_nhwc(void* y, const float* bias, const void* x, int64_t n, int32_t H, int32_t W, int32_t C,
                               int32_t dtype, evavos_stream_t stream);

/*
 * Hard masks of all frames in one pass (replaces the per-frame torch.argmax loop, the un-padding slices and the
 * contiguous copy of mivos/inference_core.py:247-257).  prob: (C, T, nh, nw) fp32, C <= 255.
 *   masks: (T, nh, nw) uint8 channel argmax (first maximal channel), or NULL
 *   out:   (T, h, w)   uint8 the same without the padding (rows pad_top.., columns pad_left..), or NULL
 */
int evavos_argmax_unpad(const float* prob, int32_t C, int64_t T, int32_t nh, int32_t nw, uint8_t* masks, uint8_t* out,
                        int32_t pad_top, int32_t pad_left, int32_t h, int32_t w, evavos_stream_t stream);

/*
 * Per-frame segmentation quality of a whole video on the device (replaces the per-frame numpy + cv2 loop of
 * interactions/eval.py:27-81 -> interactions/metrics.py:9-36, 40-160).  pred, gt: (T, h, w) uint8, non-zero =
 * foreground.  bound_pix = ceil(0.008 * sqrt(h^2 + w^2)) in the reference (metrics.py:120-121), at most 24.
 *   out (T, 4) fp64: smoothed IoU `compute_iou`, binary Jaccard, boundary F-measure, 0.5 * Jaccard + 0.5 * F
 *   gt_empty (T) int32, optional: 1 where the ground truth has no foreground pixel (eval.py:60-63 skips those)
 * workspace: evavos_jf_workspace_bytes() bytes.
 */
size_t evavos_jf_workspace_bytes(int64_t T, int32_t h, int32_t w);
int evavos_jf_metrics(const uint8_t* pred, const uint8_t* gt, int64_t T, int32_t h, int32_t w, int32_t bound_pix,
                      void* workspace, int64_t workspace_bytes, double* out, int32_t* gt_empty,
                      evavos_stream_t stream);

/*
 * Full-softmax attention read of ONE memory frame (the fusion path, prop_net.py:117-138 and :204-207):
 *   out[c][q] = sum_n vec[c][n] * softmax_n((-|m_n|^2 + 2 m_n.q_q - |q_q|^2) / sqrt(CK))
 * mem_key (CK, n_mem) and query_key (CK, n_query) fp32 with the given channel strides, unit position stride;
 * vec (n_vec, n_mem) fp32 rows (area-interpolated positive / negative interaction masks), n_vec <= 32;
 * out (n_vec, n_query) fp32.  CK must be 64.  workspace: evavos_attention_workspace_bytes() bytes.
 */
size_t evavos_attention_workspace_bytes(int32_t n_vec, int64_t n_mem, int64_t n_query, int32_t n_sm);
int evavos_attention_readout(const float* mem_key, int64_t mem_ch_stride, const float* query_key,
                             int64_t query_ch_stride, const float* vec, int64_t vec_row_stride, int32_t n_vec,
                             int32_t CK, int64_t n_mem, int64_t n_query, float* out, int64_t out_row_stride,
                             void* workspace, int64_t workspace_bytes, int32_t n_sm, evavos_stream_t stream);

/*
 * Merge step of the memory-axis sharded read.  cand_idx/cand_score: (n_query, n_cand)
 * all-gathered per-shard top-k (GLOBAL positions, -1 = empty), shard-major: n_cand / n_shards entries per
 * shard, each shard's entries best-first as evavos_memread emits them.  Selects the global top_k
 * per query (score descending, position ascending on ties), computes softmax weights with
 * the global max and denominator, and emits
 *   out_idx/out_weight/out_score (n_query, top_k): the merged result (any may be NULL)
 *   local_idx (n_query, top_k): LOCAL position of winners owned by this shard, -1 otherwise,
 *     where a global position p = frame*pos_per_frame + r is owned iff frame % n_shards == shard
 *     and its local position is (frame / n_shards)*pos_per_frame + r.
 */
int evavos_topk_merge(const int32_t* cand_idx, const float* cand_score, int64_t n_query, int32_t n_cand,
                      int32_t top_k, int32_t shard, int32_t n_shards, int64_t pos_per_frame,
                      int32_t* out_idx, float* out_weight, float* out_score, int32_t* local_idx,
                      evavos_stream_t stream);

/*
 * Same merge, fed directly with the all-gather output: gathered is [n_shards][n_query][per_shard][2] int32
 * pairs of (LOCAL position on the source shard or -1, score bits).  The local -> global mapping under the
 * round-robin frame distribution is done inside the kernel.
 */
int evavos_topk_merge_gathered(const int32_t* gathered, int64_t n_query, int32_t per_shard, int32_t top_k,
                               int32_t shard, int32_t n_shards, int64_t pos_per_frame, int32_t* out_idx,
                               float* out_weight, float* out_score, int32_t* local_idx, evavos_stream_t stream);

/*
 * Sparse readout in query-major form: out (n_query, K, CV) fp32, one contiguous row per query (the layout the
 * sharded read sums over ranks: a slice of queries is a contiguous chunk).
 */
int evavos_readout_qmajor(const EvavosBankShadow* bank, const int32_t* idx, const float* weight, int64_t n_query,
                          int32_t top_k, float* out, evavos_stream_t stream);

/*
 * Exchange buffers of the sharded read.  They are the one kind of device memory besides evavos_memread_host's
 * scratch that the library allocates itself: the buffer has to be a whole cudaMalloc allocation to be exported, and
 * it has to be opened with the importing GPU current so that the mapping belongs to THAT device's address space
 * (a framework's own IPC import maps it under the exporting device, where other GPUs' kernels cannot reach it).
 *   alloc: zero-filled buffer on the current device + its 64-byte cudaIpcMemHandle_t (pass it to the other ranks)
 *   open : map another rank's buffer into the current device (same or different GPU, another process)
 *   close / free: undo open / alloc
 */
int evavos_peer_buffer_alloc(int64_t bytes, void** ptr, uint8_t* handle64);
int evavos_peer_buffer_open(const uint8_t* handle64, void** ptr);
int evavos_peer_buffer_close(void* ptr);
int evavos_peer_buffer_free(void* ptr);

/*
 * Let kernels of the CURRENT device load from / store to memory of `peer_device` (cudaDeviceEnablePeerAccess; a no-op
 * when it is the current device or already enabled).  Call once per peer before its mapped buffer goes into an
 * EvavosPeers; mapping a buffer (CUDA IPC) does not by itself make it addressable from another device's kernels.
 */
int evavos_peer_enable(int32_t peer_device);

/*
 * Barrier among the ranks of `peers`, on the device: every rank stores `epoch` into its slot of every peer's flag
 * array (32 x uint32 at byte offset flag_offset of each buffer, zero-initialised; epochs must increase by one per
 * call on every rank) and waits until all peers' epochs have arrived in its own.  Orders everything the stream did
 * before the call (stores to peer buffers included) before everything after it on all ranks.  A peer that does not
 * show up within ~2 s makes the kernel give up and poison slot 31 (checked by evavos_peer_barrier_ok).
 */
int evavos_peer_barrier(const EvavosPeers* peers, int64_t flag_offset, uint32_t epoch, evavos_stream_t stream);

/*
 * Sum-reduce-scatter over peer memory: out[row][q - q0] = sum_g partial_g[q][row] for q in [q0, q1), where
 * partial_g = (n_query, rows) fp32 query-major at byte offset partial_offset of rank g's buffer (rows = K * CV,
 * written by evavos_readout_qmajor).  out is (rows, q1 - q0) fp32 with row stride out_row_stride - the slice of
 * the reference-layout readout (K, CV, HW) this rank owns.  Loads come straight from the peers' memory.
 */
int evavos_peer_reduce_scatter(const EvavosPeers* peers, int64_t partial_offset, int32_t rows, int64_t q0,
                               int64_t q1, float* out, int64_t out_row_stride, evavos_stream_t stream);

/*
 * Host-buffer form of the read, reference layouts, synchronous:
 *   mem_key (CK, n_pos) fp32 contiguous, query (CK, n_query), mem_value (K, CV, n_pos),
 *   readout (K, CV, n_query); topk_idx/topk_weight optional (n_query, top_k).
 * Copies the inputs to the device, builds the shadow, reads, copies the result back.
 * h2d_bytes/d2h_bytes (optional) receive the bytes moved.
 */
int evavos_memread_host(const float* mem_key, const float* query, const float* mem_value, int32_t K,
                        int32_t CK, int32_t CV, int64_t n_pos, int64_t n_query, int32_t top_k, int32_t path,
                        float* readout, int32_t* topk_idx, float* topk_weight, int64_t* h2d_bytes,
                        int64_t* d2h_bytes);

/*
 * Diagnostics: when enabled, evavos_memread records CUDA events between its stages on the caller's stream (a ring
 * of 256 sets: no synchronisation between calls); evavos_stage_timing_read waits for the newest call and returns
 * the mean ms of {candidate selection, 0 (unused), finalize, readout} over the calls since the last read.
 * The event records sit between the kernels and defeat their programmatic-dependent-launch overlap, so the stage
 * times add up to more than an un-instrumented call.  Process-wide, not thread-safe; off by default.
 */
int evavos_stage_timing(int32_t enable);
int evavos_stage_timing_read(float* ms4);

/* Frees the device scratch evavos_memread_host keeps between calls. */
int evavos_release_host_scratch(void);

#ifdef __cplusplus
}
#endif
#endif /* EVAVOS_H_ */
