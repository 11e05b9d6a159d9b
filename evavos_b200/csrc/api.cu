// extern "C" entry points of libevavos_sm100.so (see include/evavos.h for the contract).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace evavos {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return EVAVOS_ERR_CUDA;
}

namespace {

// SM count of the CURRENT device, cached per device ordinal (a process may drive several GPUs).
int device_sm_count(int* n_sm) {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cached[kMaxDev];
  int dev = 0;
  EVAVOS_CUDA_OK(cudaGetDevice(&dev));
  int have = (dev >= 0 && dev < kMaxDev) ? cached[dev].load(std::memory_order_relaxed) : 0;
  if (have == 0) {
    int major = 0;
    EVAVOS_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
      set_error("libevavos_sm100 needs an sm_100 (B200) device, found compute capability major %d", major);
      return EVAVOS_ERR_UNSUPPORTED;
    }
    EVAVOS_CUDA_OK(cudaDeviceGetAttribute(&have, cudaDevAttrMultiProcessorCount, dev));
    if (dev >= 0 && dev < kMaxDev) cached[dev].store(have, std::memory_order_relaxed);
  }
  *n_sm = have;
  return EVAVOS_OK;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

bool use_tensor_path(const EvavosMemReadArgs& a) {
  if (a.path == EVAVOS_PATH_SIMT) return false;
  return a.bank.CK == 64 && a.bank.key_tiles != nullptr && a.bank.key_maxnorm != nullptr;
}

struct Carve {
  SelectBuffers sb;
  int32_t* idx;
  float* weight;
  size_t total;
};

// Lays the workspace out; with base == nullptr only the size is computed.
Carve carve_workspace(const EvavosMemReadArgs& a, int n_chunks, int n_sm, uint8_t* base) {
  Carve c;
  memset(&c, 0, sizeof(c));
  const int64_t mt = score_pass_mtiles(a.n_query);   // whole clusters of query tiles (select buffers cover the padding)
  const int64_t nq_pad = mt * 128;
  size_t off = 0;
  auto take = [&](size_t bytes) -> uint8_t* {
    uint8_t* p = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return p;
  };
  // (per launch [chunks][query rows of the launch][128]; multi-wave reads - more query tiles than SMs - use at most
  //  n_sm tile rows per launch, fewer than the nq_pad rows reserved here)
  c.sb.class_max = reinterpret_cast<float*>(take(n_chunks > 0 ? sizeof(float) * (size_t)n_chunks * nq_pad * 128 : 0));
  c.sb.tau = reinterpret_cast<float*>(take(sizeof(float) * nq_pad));
  c.sb.cand_cnt = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * nq_pad));
  c.sb.cand = reinterpret_cast<int2*>(take(sizeof(int2) * nq_pad * kCandCap));
  c.sb.strip = take(n_chunks > 0 ? score_pass_strip_bytes(a.n_query, n_chunks, n_sm) : 0);
  c.sb.grid_counter = reinterpret_cast<unsigned int*>(take(sizeof(unsigned int) * (size_t)(mt + 1)));
  c.sb.overflow_cnt = c.sb.grid_counter ? c.sb.grid_counter + mt : nullptr;
  c.sb.overflow_list = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * nq_pad));
  c.idx = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * a.n_query * a.top_k));
  c.weight = reinterpret_cast<float*>(take(sizeof(float) * a.n_query * a.top_k));
  c.total = off + 1024;  // slack for aligning the caller's pointer
  return c;
}

int validate_bank(const EvavosBankShadow* b) {
  if (!b) { set_error("bank is NULL"); return EVAVOS_ERR_INVALID; }
  if (b->CK <= 0 || b->CK > 64 || (b->CK % 8) != 0) {
    set_error("CK=%d unsupported (need a multiple of 8, <= 64)", b->CK);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  if (b->K < 0 || b->CV < 0 || b->capacity_pos <= 0) { set_error("bad bank sizes"); return EVAVOS_ERR_INVALID; }
  if (b->val_dtype != EVAVOS_F32 && b->val_dtype != EVAVOS_BF16) {
    set_error("val_dtype=%d unsupported", b->val_dtype);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  return EVAVOS_OK;
}

int validate_peers(const EvavosPeers* p) {
  if (!p) { set_error("peers is NULL"); return EVAVOS_ERR_INVALID; }
  if (p->n_ranks < 1 || p->n_ranks > EVAVOS_MAX_RANKS || p->rank < 0 || p->rank >= p->n_ranks) {
    set_error("peers: n_ranks=%d rank=%d out of range (1..%d)", p->n_ranks, p->rank, EVAVOS_MAX_RANKS);
    return EVAVOS_ERR_INVALID;
  }
  for (int g = 0; g < p->n_ranks; ++g)
    if (!p->base[g]) { set_error("peers: base[%d] is NULL", g); return EVAVOS_ERR_INVALID; }
  return EVAVOS_OK;
}

int validate_read(const EvavosMemReadArgs* a) {
  if (!a) { set_error("args is NULL"); return EVAVOS_ERR_INVALID; }
  int rc = validate_bank(&a->bank);
  if (rc) return rc;
  if (!a->bank.key_pm) { set_error("bank.key_pm is NULL"); return EVAVOS_ERR_INVALID; }
  if (!a->query || a->n_query <= 0 || a->n_pos <= 0) { set_error("empty query or bank"); return EVAVOS_ERR_INVALID; }
  if (a->n_pos > a->bank.capacity_pos) { set_error("n_pos exceeds bank capacity"); return EVAVOS_ERR_INVALID; }
  if (a->n_pos >= (int64_t)1 << 31) { set_error("n_pos must fit int32"); return EVAVOS_ERR_UNSUPPORTED; }
  if (a->top_k <= 0 || a->top_k > EVAVOS_MAX_TOPK) {
    set_error("top_k=%d unsupported (1..%d)", a->top_k, EVAVOS_MAX_TOPK);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  if (a->n_pos < a->top_k) {
    set_error("selected index k out of range (THW=%lld < top_k=%d)", (long long)a->n_pos, a->top_k);
    return EVAVOS_ERR_TOPK_RANGE;
  }
  if ((a->path == EVAVOS_PATH_TENSOR || a->path == EVAVOS_PATH_TENSOR_DENSE) &&
      !(a->bank.CK == 64 && a->bank.key_tiles && a->bank.key_maxnorm)) {
    set_error("tensor path needs CK == 64, key_tiles and key_maxnorm");
    return EVAVOS_ERR_UNSUPPORTED;
  }
  if (a->readout && (!a->bank.val_pm || a->bank.K <= 0)) { set_error("readout requested without values"); return EVAVOS_ERR_INVALID; }
  if (a->queries_per_frame < 0 || a->queries_per_frame > 0x7fffffff ||
      (a->queries_per_frame > 0 && (a->n_query % a->queries_per_frame) != 0)) {
    set_error("queries_per_frame must divide n_query");
    return EVAVOS_ERR_INVALID;
  }
  if (a->peers != nullptr) {
    int rc2 = validate_peers(a->peers);
    if (rc2) return rc2;
    if (a->peer_gather_offset < 0 || (a->peer_gather_offset % 8) != 0) { set_error("peer_gather_offset must be a non-negative multiple of 8"); return EVAVOS_ERR_INVALID; }
  }
  return EVAVOS_OK;
}

int resolve_sm(const EvavosMemReadArgs* a, int* n_sm) {
  if (a->n_sm > 0) { *n_sm = a->n_sm; return EVAVOS_OK; }
  return device_sm_count(n_sm);
}

}  // namespace
}  // namespace evavos

using namespace evavos;

extern "C" {

int evavos_abi_version(void) { return EVAVOS_ABI_VERSION; }
const char* evavos_last_error(void) { return g_err; }
size_t evavos_sizeof_bank_shadow(void) { return sizeof(EvavosBankShadow); }
size_t evavos_sizeof_memread_args(void) { return sizeof(EvavosMemReadArgs); }

size_t evavos_key_tiles_bytes(int64_t capacity_pos) {
  if (capacity_pos <= 0) return 0;
  return (size_t)ceil_div(capacity_pos, kTilePos) * kTileBytes;
}

int evavos_bank_write_keys(const EvavosBankShadow* bank, const float* src, int64_t src_ch_stride, int64_t pos0,
                           int64_t n_pos, float* dst_ref, int64_t dst_ref_ch_stride, evavos_stream_t stream) {
  int rc = validate_bank(bank);
  if (rc) return rc;
  if (!src || !bank->key_pm || pos0 < 0 || n_pos < 0 || pos0 + n_pos > bank->capacity_pos) {
    set_error("bank_write_keys: bad range [%lld, +%lld) for capacity %lld", (long long)pos0, (long long)n_pos,
              (long long)bank->capacity_pos);
    return EVAVOS_ERR_INVALID;
  }
  return launch_write_keys(*bank, src, src_ch_stride, pos0, n_pos, dst_ref, dst_ref_ch_stride, (cudaStream_t)stream);
}

int evavos_bank_write_values(const EvavosBankShadow* bank, const float* src, int64_t src_obj_stride,
                             int64_t src_ch_stride, int64_t pos0, int64_t n_pos, float* dst_ref,
                             int64_t dst_ref_obj_stride, int64_t dst_ref_ch_stride, evavos_stream_t stream) {
  int rc = validate_bank(bank);
  if (rc) return rc;
  if (!src || !bank->val_pm || bank->K <= 0 || bank->CV <= 0 || pos0 < 0 || n_pos < 0 ||
      pos0 + n_pos > bank->capacity_pos) {
    set_error("bank_write_values: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_write_values(*bank, src, src_obj_stride, src_ch_stride, pos0, n_pos, dst_ref, dst_ref_obj_stride,
                             dst_ref_ch_stride, (cudaStream_t)stream);
}

size_t evavos_memread_workspace_bytes(const EvavosMemReadArgs* args) {
  if (validate_read(args)) return 0;
  int n_sm = 148;
  if (args->n_sm > 0) n_sm = args->n_sm;
  else if (device_sm_count(&n_sm)) n_sm = 148;
  const int chunks = use_tensor_path(*args) ? score_pass_chunks(args->n_pos, args->n_query, n_sm) : 0;
  return carve_workspace(*args, chunks, n_sm, nullptr).total;
}

// ---- overflow hint: does the exact tiled pass (select_dense.cu) have anything to do? ---------------------------------
// Its launch costs ~2 us of every read even when no list overflowed - the normal case.  The finalizer therefore sets
// a word of mapped host memory whenever it meets an overflowed query; a read launches the tiled pass only if one of
// the previous 64 reads on this device left the word set (or the caller asks with EVAVOS_PATH_TENSOR_DENSE).
// Without the pass the finalizer redoes an overflowed query itself (one warp, whole bank): the hint only ever trades
// speed, never the result.  Races between host threads on the word are harmless for the same reason.
namespace {
struct OverflowHint {
  volatile uint32_t* host;
  uint32_t* dev;
  bool tried;
  int sticky;   // reads left for which the pass is launched after the word was last seen set: a wrong launch costs
                // 2 us, a wrong skip a query redone by one warp (~0.3 ms), so one overflow buys the pass 64 reads
};
OverflowHint g_hint[64];

OverflowHint* overflow_hint() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  OverflowHint& h = g_hint[dev];
  if (!h.tried) {
    h.tried = true;
    void* hp = nullptr;
    void* dp = nullptr;
    if (cudaHostAlloc(&hp, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      if (cudaHostGetDevicePointer(&dp, hp, 0) == cudaSuccess) {
        h.host = reinterpret_cast<volatile uint32_t*>(hp);
        h.dev = reinterpret_cast<uint32_t*>(dp);
        *h.host = 0;
      } else {
        cudaFreeHost(hp);
      }
    }
    (void)cudaGetLastError();
  }
  return h.host ? &h : nullptr;
}
}  // namespace

// ---- optional per-stage timing of evavos_memread (diagnostics; CUDA events on the caller's stream) ----------
// A ring of event sets, one per evavos_memread call, so that a timed loop needs no synchronisation between calls;
// evavos_stage_timing_read waits for the newest set and averages every set recorded since the last read.
namespace {
constexpr int kTimingRing = 256;
bool g_timing = false;
cudaEvent_t g_ev[kTimingRing][5];
bool g_ev_made = false;
int g_ev_next = 0;    // set the next call records into
int g_ev_count = 0;   // sets recorded since the last read (<= kTimingRing)
inline void stage_mark(int i, cudaStream_t st) {
  if (g_timing) cudaEventRecord(g_ev[g_ev_next][i], st);
}
inline void stage_done() {
  if (!g_timing) return;
  g_ev_next = (g_ev_next + 1) % kTimingRing;
  if (g_ev_count < kTimingRing) ++g_ev_count;
}
}  // namespace

int evavos_stage_timing(int32_t enable) {
  if (enable && !g_ev_made) {
    for (int r = 0; r < kTimingRing; ++r)
      for (int i = 0; i < 5; ++i) EVAVOS_CUDA_OK(cudaEventCreate(&g_ev[r][i]));
    g_ev_made = true;
  }
  g_timing = enable != 0;
  g_ev_count = 0;
  return EVAVOS_OK;
}

int evavos_stage_timing_read(float* ms4) {
  if (!g_ev_made || !ms4 || g_ev_count == 0) { set_error("stage timing: nothing recorded"); return EVAVOS_ERR_INVALID; }
  const int newest = (g_ev_next + kTimingRing - 1) % kTimingRing;
  EVAVOS_CUDA_OK(cudaEventSynchronize(g_ev[newest][4]));
  double acc[4] = {0, 0, 0, 0};
  for (int c = 0; c < g_ev_count; ++c) {
    const int r = (g_ev_next + kTimingRing - 1 - c) % kTimingRing;
    for (int i = 0; i < 4; ++i) {
      float ms = 0.f;
      EVAVOS_CUDA_OK(cudaEventElapsedTime(&ms, g_ev[r][i], g_ev[r][i + 1]));
      acc[i] += ms;
    }
  }
  for (int i = 0; i < 4; ++i) ms4[i] = (float)(acc[i] / g_ev_count);
  g_ev_count = 0;
  return EVAVOS_OK;
}

int evavos_memread(const EvavosMemReadArgs* a, evavos_stream_t stream) {
  int rc = validate_read(a);
  if (rc) return rc;
  int n_sm = 0;
  rc = resolve_sm(a, &n_sm);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tensor = use_tensor_path(*a);
  const int chunks = tensor ? score_pass_chunks(a->n_pos, a->n_query, n_sm) : 0;
  if (!a->workspace) { set_error("workspace is NULL"); return EVAVOS_ERR_WORKSPACE; }
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(a->workspace), 1024));
  const Carve probe = carve_workspace(*a, chunks, n_sm, nullptr);
  if ((int64_t)probe.total > a->workspace_bytes) {
    set_error("workspace too small: need %zu bytes, got %lld", probe.total, (long long)a->workspace_bytes);
    return EVAVOS_ERR_WORKSPACE;
  }
  const Carve c = carve_workspace(*a, chunks, n_sm, base);
  const int CK = a->bank.CK;

  int32_t* idx = a->topk_idx ? a->topk_idx : c.idx;
  float* weight = a->topk_weight ? a->topk_weight : c.weight;

  // 1. candidate generation (the query is consumed in the caller's layout; no query shadow)
  stage_mark(0, st);
  if (tensor) {
    const int stride = score_pass_sample_stride(a->n_pos, chunks, a->sample_stride);
    rc = launch_score_select(a->query, a->query_ch_stride, a->bank.key_tiles, a->bank.key_maxnorm, a->n_pos, a->n_query,
                             a->top_k, chunks, stride, n_sm, c.sb.class_max, c.sb.tau, c.sb.cand, c.sb.cand_cnt,
                             c.sb.strip, c.sb.grid_counter, st);
  } else {
    rc = launch_brute_select(a->bank.key_pm, a->query, a->query_ch_stride, CK, a->n_pos, a->n_query, a->top_k,
                             c.sb.cand, c.sb.cand_cnt, n_sm, st);
  }
  if (rc) return rc;
  stage_mark(1, st);   // (stage 1, the round-1 exact fallback launch, no longer exists: always ~0)
  stage_mark(2, st);

  // 2. tightening of the list (tensor path), exact rescoring, top-k, softmax; a query whose list overflowed is
  //    only recorded (tensor path) and finished by step 2b
  bool dense_pass = false;
  uint32_t* hint_dev = nullptr;
  if (tensor) {
    OverflowHint* h = overflow_hint();
    if (h) {
      if (*h->host != 0) {
        h->sticky = 64;
        *h->host = 0;   // the kernels of this and the following reads set it again if they meet an overflowed list
      }
      hint_dev = h->dev;
    }
    dense_pass = a->path == EVAVOS_PATH_TENSOR_DENSE || h == nullptr || h->sticky > 0;
    if (h && h->sticky > 0) --h->sticky;
  }
  rc = launch_finalize(a->bank.key_pm, a->query, a->query_ch_stride, CK, a->n_pos, a->n_query, a->top_k, c.sb.cand,
                       c.sb.cand_cnt, tensor ? 1 : 0, a->bank.key_maxnorm, idx, weight, a->topk_score, a->peers,
                       a->peer_gather_offset, dense_pass ? c.sb.overflow_list : nullptr,
                       tensor ? c.sb.overflow_cnt : nullptr, hint_dev, st);
  if (rc) return rc;
  // 2b. queries whose list overflowed (the filter cannot separate their candidates): exact tiled selection + the same
  //     finalizer (see the overflow hint above for when it is launched)
  if (dense_pass) {
    rc = launch_overflow_exact(a->bank.key_pm, a->query, a->query_ch_stride, a->n_pos, a->n_query, a->top_k, c.sb.cand,
                               c.sb.overflow_list, c.sb.overflow_cnt, a->bank.key_maxnorm, idx, weight, a->topk_score,
                               a->peers, a->peer_gather_offset, n_sm, st);
    if (rc) return rc;
  }
  stage_mark(3, st);

  // 3. sparse readout for all objects
  if (a->readout) {
    rc = launch_readout(a->bank, idx, weight, a->n_query, a->top_k, a->readout, a->readout_obj_stride,
                        a->readout_ch_stride, (int)a->queries_per_frame, a->readout_frame_stride, st);
    if (rc) return rc;
  }
  stage_mark(4, st);
  stage_done();
  return EVAVOS_OK;
}

int evavos_memread_overflow_count(const EvavosMemReadArgs* a, uint32_t* count, evavos_stream_t stream) {
  int rc = validate_read(a);
  if (rc) return rc;
  if (!count || !a->workspace) { set_error("overflow_count: NULL argument"); return EVAVOS_ERR_INVALID; }
  *count = 0;
  if (!use_tensor_path(*a)) return EVAVOS_OK;
  int n_sm = 0;
  if (a->n_sm > 0) n_sm = a->n_sm;
  else if (device_sm_count(&n_sm)) n_sm = 148;
  const int chunks = score_pass_chunks(a->n_pos, a->n_query, n_sm);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(a->workspace), 1024));
  const Carve c = carve_workspace(*a, chunks, n_sm, base);
  cudaStream_t st = (cudaStream_t)stream;
  EVAVOS_CUDA_OK(cudaMemcpyAsync(count, c.sb.overflow_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  EVAVOS_CUDA_OK(cudaStreamSynchronize(st));
  return EVAVOS_OK;
}

int evavos_readout(const EvavosBankShadow* bank, const int32_t* idx, const float* weight, int64_t n_query,
                   int32_t top_k, float* out, int64_t out_obj_stride, int64_t out_ch_stride,
                   evavos_stream_t stream) {
  int rc = validate_bank(bank);
  if (rc) return rc;
  if (!bank->val_pm || !idx || !weight || !out || n_query <= 0 || top_k <= 0 || top_k > EVAVOS_MAX_TOPK) {
    set_error("readout: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_readout(*bank, idx, weight, n_query, top_k, out, out_obj_stride, out_ch_stride, 0, 0, (cudaStream_t)stream);
}

int evavos_readout_qmajor(const EvavosBankShadow* bank, const int32_t* idx, const float* weight, int64_t n_query,
                          int32_t top_k, float* out, evavos_stream_t stream) {
  int rc = validate_bank(bank);
  if (rc) return rc;
  if (!bank->val_pm || !idx || !weight || !out || n_query <= 0 || top_k <= 0 || top_k > EVAVOS_MAX_TOPK ||
      (reinterpret_cast<uintptr_t>(out) % 16) != 0) {
    set_error("readout_qmajor: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_readout(*bank, idx, weight, n_query, top_k, out, 0, -1, 0, 0, (cudaStream_t)stream);
}

int evavos_peer_enable(int32_t peer_device) {
  int cur = 0;
  EVAVOS_CUDA_OK(cudaGetDevice(&cur));
  if (peer_device == cur) return EVAVOS_OK;
  int can = 0;
  EVAVOS_CUDA_OK(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) {
    set_error("device %d cannot access device %d (no NVLink / PCIe peer path)", cur, (int)peer_device);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();   // clear the sticky-less error state
    return EVAVOS_OK;
  }
  EVAVOS_CUDA_OK(e);
  return EVAVOS_OK;
}

int evavos_peer_buffer_alloc(int64_t bytes, void** ptr, uint8_t* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (bytes <= 0 || !ptr || !handle64) { set_error("peer_buffer_alloc: bad arguments"); return EVAVOS_ERR_INVALID; }
  void* p = nullptr;
  EVAVOS_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaMemset / cudaIpcGetMemHandle");
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return EVAVOS_OK;
}

int evavos_peer_buffer_open(const uint8_t* handle64, void** ptr) {
  if (!handle64 || !ptr) { set_error("peer_buffer_open: bad arguments"); return EVAVOS_ERR_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  // opened on the CURRENT device: the mapping is addressable by this device's kernels (peer access over NVLink is
  // enabled lazily by the driver when the memory lives on another GPU)
  EVAVOS_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return EVAVOS_OK;
}

int evavos_peer_buffer_close(void* ptr) {
  if (ptr) EVAVOS_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return EVAVOS_OK;
}

int evavos_peer_buffer_free(void* ptr) {
  if (ptr) EVAVOS_CUDA_OK(cudaFree(ptr));
  return EVAVOS_OK;
}

int evavos_peer_barrier(const EvavosPeers* peers, int64_t flag_offset, uint32_t epoch, evavos_stream_t stream) {
  int rc = validate_peers(peers);
  if (rc) return rc;
  if (flag_offset < 0 || (flag_offset % 4) != 0) { set_error("peer_barrier: bad flag_offset"); return EVAVOS_ERR_INVALID; }
  return launch_peer_barrier(*peers, flag_offset, epoch, (cudaStream_t)stream);
}

int evavos_peer_reduce_scatter(const EvavosPeers* peers, int64_t partial_offset, int32_t rows, int64_t q0, int64_t q1,
                               float* out, int64_t out_row_stride, evavos_stream_t stream) {
  int rc = validate_peers(peers);
  if (rc) return rc;
  if (!out || rows <= 0 || (rows % 4) != 0 || q0 < 0 || q1 < q0 || partial_offset < 0 || (partial_offset % 16) != 0) {
    set_error("peer_reduce_scatter: bad arguments (rows must be a multiple of 4, offsets 16-byte aligned)");
    return EVAVOS_ERR_INVALID;
  }
  return launch_peer_reduce_scatter(*peers, partial_offset, rows, q0, q1, out, out_row_stride, (cudaStream_t)stream);
}

size_t evavos_jf_workspace_bytes(int64_t T, int32_t h, int32_t w) {
  if (T <= 0 || h <= 0 || w <= 0) return 0;
  return jf_workspace_bytes(T, h, w);
}

int evavos_jf_metrics(const uint8_t* pred, const uint8_t* gt, int64_t T, int32_t h, int32_t w, int32_t bound_pix,
                      void* workspace, int64_t workspace_bytes, double* out, int32_t* gt_empty, evavos_stream_t stream) {
  if (!pred || !gt || !out || !workspace || T <= 0 || h <= 0 || w <= 0) {
    set_error("jf_metrics: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  if (workspace_bytes < (int64_t)jf_workspace_bytes(T, h, w)) {
    set_error("jf_metrics: workspace of %lld bytes, %zu needed", (long long)workspace_bytes, jf_workspace_bytes(T, h, w));
    return EVAVOS_ERR_WORKSPACE;
  }
  return launch_jf_metrics(pred, gt, T, h, w, bound_pix, workspace, out, gt_empty, (cudaStream_t)stream);
}

int evavos_affinity_dense(const int32_t* idx, const float* weight, int64_t n_query, int32_t top_k, int64_t n_pos,
                          float* dense, evavos_stream_t stream) {
  if (!idx || !weight || !dense || n_query <= 0 || top_k <= 0 || n_pos <= 0) {
    set_error("affinity_dense: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_affinity_dense(idx, weight, n_query, top_k, n_pos, dense, (cudaStream_t)stream);
}

int evavos_aggregate_wbg(const float* prob, float* out, int32_t K, int64_t npix, int32_t keep_bg, int32_t hard,
                         evavos_stream_t stream) {
  if (!prob || !out || K <= 0 || npix < 0) {
    set_error("aggregate_wbg: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_aggregate(prob, out, K, npix, keep_bg, hard, (cudaStream_t)stream);
}

int evavos_bias_residual_nhwc(void* y, const float* bias, const void* residual, int64_t rows, int32_t C, int32_t dtype,
                              int32_t relu, evavos_stream_t stream) {
  const int vec = dtype == EVAVOS_BF16 ? 8 : 4;
  if (!y || !bias || rows < 0 || C <= 0 || C % vec != 0 || (dtype != EVAVOS_F32 && dtype != EVAVOS_BF16) ||
      reinterpret_cast<uintptr_t>(y) % 16 != 0 || (residual && reinterpret_cast<uintptr_t>(residual) % 16 != 0) ||
      reinterpret_cast<uintptr_t>(bias) % 16 != 0) {
    set_error("bias_residual_nhwc: bad arguments (16-byte aligned channel-contiguous rows, C %% %d == 0)", vec);
    return EVAVOS_ERR_INVALID;
  }
  return launch_bias_residual(y, bias, residual, rows, C, dtype == EVAVOS_BF16, relu, (cudaStream_t)stream);
}

int evavos_upsample2x_add_nhwc(void* y, const float* bias, const void* x, int64_t n, int32_t H, int32_t W, int32_t C,
                               int32_t dtype, evavos_stream_t stream) {
  const int vec = dtype == EVAVOS_BF16 ? 8 : 4;
  if (!y || !bias || !x || n < 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || C % vec != 0 ||
      (dtype != EVAVOS_F32 && dtype != EVAVOS_BF16) || reinterpret_cast<uintptr_t>(y) % 16 != 0 ||
      reinterpret_cast<uintptr_t>(x) % 16 != 0 || reinterpret_cast<uintptr_t>(bias) % 16 != 0) {
    set_error("upsample2x_add_nhwc: bad arguments (even H, W; 16-byte aligned NHWC tensors, C %% %d == 0)", vec);
    return EVAVOS_ERR_INVALID;
  }
  return launch_upsample2x_add(y, bias, x, n, H, W, C, dtype == EVAVOS_BF16, (cudaStream_t)stream);
}

int evavos_argmax_unpad(const float* prob, int32_t C, int64_t T, int32_t nh, int32_t nw, uint8_t* masks, uint8_t* out,
                        int32_t pad_top, int32_t pad_left, int32_t h, int32_t w, evavos_stream_t stream) {
  if (!prob || C <= 0 || C > 255 || T < 0 || nh <= 0 || nw <= 0 || (!masks && !out) ||
      (out && (pad_top < 0 || pad_left < 0 || h <= 0 || w <= 0 || pad_top + h > nh || pad_left + w > nw))) {
    set_error("argmax_unpad: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_argmax_unpad(prob, C, T, nh, nw, masks, out, pad_top, pad_left, h, w, (cudaStream_t)stream);
}

size_t evavos_attention_workspace_bytes(int32_t n_vec, int64_t n_mem, int64_t n_query, int32_t n_sm) {
  if (n_vec <= 0 || n_vec > 32 || n_mem <= 0 || n_query <= 0) return 0;
  return attention_workspace_bytes(n_vec, n_mem, n_query, n_sm > 0 ? n_sm : 148);
}

int evavos_attention_readout(const float* mem_key, int64_t mem_ch_stride, const float* query_key,
                             int64_t query_ch_stride, const float* vec, int64_t vec_row_stride, int32_t n_vec,
                             int32_t CK, int64_t n_mem, int64_t n_query, float* out, int64_t out_row_stride,
                             void* workspace, int64_t workspace_bytes, int32_t n_sm, evavos_stream_t stream) {
  if (!mem_key || !query_key || !vec || !out || n_vec <= 0 || n_mem <= 0 || n_query <= 0) {
    set_error("attention_readout: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  if (CK != 64 || n_vec > 32) {
    set_error("attention_readout: CK=%d, n_vec=%d unsupported (CK must be 64, n_vec <= 32)", (int)CK, (int)n_vec);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  if (n_sm <= 0) n_sm = 148;
  if (!workspace || workspace_bytes < (int64_t)attention_workspace_bytes(n_vec, n_mem, n_query, n_sm)) {
    set_error("attention_readout: workspace of %lld bytes, %lld needed", (long long)workspace_bytes,
              (long long)attention_workspace_bytes(n_vec, n_mem, n_query, n_sm));
    return EVAVOS_ERR_WORKSPACE;
  }
  return launch_attention_readout(mem_key, mem_ch_stride, query_key, query_ch_stride, vec, vec_row_stride, n_vec, CK,
                                  n_mem, n_query, out, out_row_stride, workspace, n_sm, (cudaStream_t)stream);
}

int evavos_topk_merge(const int32_t* cand_idx, const float* cand_score, int64_t n_query, int32_t n_cand,
                      int32_t top_k, int32_t shard, int32_t n_shards, int64_t pos_per_frame, int32_t* out_idx,
                      float* out_weight, float* out_score, int32_t* local_idx, evavos_stream_t stream) {
  if (!cand_idx || !cand_score || n_query <= 0 || n_cand <= 0 || top_k <= 0 || top_k > EVAVOS_MAX_TOPK ||
      n_shards <= 0 || shard < 0 || shard >= n_shards || pos_per_frame <= 0 || pos_per_frame > 0x7fffffff) {
    set_error("topk_merge: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_topk_merge(cand_idx, cand_score, n_query, n_cand, top_k, shard, n_shards, pos_per_frame, out_idx,
                           out_weight, out_score, local_idx, 0, (cudaStream_t)stream);
}

int evavos_topk_merge_gathered(const int32_t* gathered, int64_t n_query, int32_t per_shard, int32_t top_k,
                               int32_t shard, int32_t n_shards, int64_t pos_per_frame, int32_t* out_idx,
                               float* out_weight, float* out_score, int32_t* local_idx, evavos_stream_t stream) {
  if (!gathered || n_query <= 0 || per_shard <= 0 || top_k <= 0 || top_k > EVAVOS_MAX_TOPK || n_shards <= 0 ||
      shard < 0 || shard >= n_shards || pos_per_frame <= 0 || pos_per_frame > 0x7fffffff) {
    set_error("topk_merge_gathered: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  return launch_topk_merge(gathered, nullptr, n_query, per_shard * n_shards, top_k, shard, n_shards, pos_per_frame,
                           out_idx, out_weight, out_score, local_idx, 1, (cudaStream_t)stream);
}

// ---- host-buffer form ---------------------------------------------------------------------
namespace {
struct HostScratch {
  uint8_t* dev = nullptr;
  size_t bytes = 0;
  int device = -1;   // ordinal the scratch was allocated on
};
HostScratch g_scratch;   // evavos_memread_host is synchronous and documented as not re-entrant
}  // namespace

int evavos_release_host_scratch(void) {
  if (g_scratch.dev) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (g_scratch.device >= 0 && g_scratch.device != cur) cudaSetDevice(g_scratch.device);
    cudaFree(g_scratch.dev);
    if (g_scratch.device >= 0 && g_scratch.device != cur) cudaSetDevice(cur);
    g_scratch.dev = nullptr;
    g_scratch.bytes = 0;
    g_scratch.device = -1;
  }
  return EVAVOS_OK;
}

int evavos_memread_host(const float* mem_key, const float* query, const float* mem_value, int32_t K, int32_t CK,
                        int32_t CV, int64_t n_pos, int64_t n_query, int32_t top_k, int32_t path, float* readout,
                        int32_t* topk_idx, float* topk_weight, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (!mem_key || !query || (readout && !mem_value) || n_pos <= 0 || n_query <= 0) {
    set_error("memread_host: bad arguments");
    return EVAVOS_ERR_INVALID;
  }
  EvavosMemReadArgs a;
  memset(&a, 0, sizeof(a));
  a.bank.capacity_pos = n_pos;
  a.bank.K = readout ? K : 0;
  a.bank.CK = CK;
  a.bank.CV = CV;
  a.bank.val_dtype = EVAVOS_F32;
  a.n_pos = n_pos;
  a.n_query = n_query;
  a.query_ch_stride = n_query;
  a.top_k = top_k;
  a.path = path;
  // placeholders so validate_read / workspace sizing see a complete description
  a.bank.key_pm = reinterpret_cast<float*>(0x1000);
  a.bank.key_tiles = (CK == 64) ? reinterpret_cast<void*>(0x1000) : nullptr;
  a.bank.key_maxnorm = reinterpret_cast<float*>(0x1000);
  a.bank.val_pm = readout ? reinterpret_cast<void*>(0x1000) : nullptr;
  a.query = reinterpret_cast<const float*>(0x1000);
  a.readout = readout ? reinterpret_cast<float*>(0x1000) : nullptr;
  int rc = validate_read(&a);
  if (rc) return rc;
  const size_t ws = evavos_memread_workspace_bytes(&a);

  const size_t b_key = sizeof(float) * (size_t)CK * n_pos;
  const size_t b_q = sizeof(float) * (size_t)CK * n_query;
  const size_t b_val = readout ? sizeof(float) * (size_t)K * CV * n_pos : 0;
  const size_t b_out = readout ? sizeof(float) * (size_t)K * CV * n_query : 0;
  const size_t b_tk = sizeof(float) * (size_t)n_query * top_k;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t o_key = take(b_key), o_q = take(b_q), o_val = take(b_val), o_out = take(b_out);
  const size_t o_idx = take(b_tk), o_w = take(b_tk);
  const size_t o_kpm = take(b_key), o_tiles = take(evavos_key_tiles_bytes(n_pos)), o_max = take(4);
  const size_t o_vpm = take(b_val), o_ws = take(ws);
  const size_t total = off + 1024;
  int cur_dev = 0;
  EVAVOS_CUDA_OK(cudaGetDevice(&cur_dev));
  if (g_scratch.bytes < total || g_scratch.device != cur_dev) {
    evavos_release_host_scratch();
    EVAVOS_CUDA_OK(cudaMalloc(&g_scratch.dev, total));
    g_scratch.bytes = total;
    g_scratch.device = cur_dev;
  }
  uint8_t* d = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(g_scratch.dev), 1024));
  cudaStream_t st = 0;
  EVAVOS_CUDA_OK(cudaMemcpyAsync(d + o_key, mem_key, b_key, cudaMemcpyHostToDevice, st));
  EVAVOS_CUDA_OK(cudaMemcpyAsync(d + o_q, query, b_q, cudaMemcpyHostToDevice, st));
  if (readout) EVAVOS_CUDA_OK(cudaMemcpyAsync(d + o_val, mem_value, b_val, cudaMemcpyHostToDevice, st));
  EVAVOS_CUDA_OK(cudaMemsetAsync(d + o_max, 0, 4, st));

  a.bank.key_pm = reinterpret_cast<float*>(d + o_kpm);
  a.bank.key_tiles = (CK == 64) ? (void*)(d + o_tiles) : nullptr;
  a.bank.key_maxnorm = reinterpret_cast<float*>(d + o_max);
  a.bank.val_pm = readout ? (void*)(d + o_vpm) : nullptr;
  a.query = reinterpret_cast<const float*>(d + o_q);
  a.readout = readout ? reinterpret_cast<float*>(d + o_out) : nullptr;
  a.topk_idx = reinterpret_cast<int32_t*>(d + o_idx);
  a.topk_weight = reinterpret_cast<float*>(d + o_w);
  a.workspace = d + o_ws;
  a.workspace_bytes = (int64_t)ws;

  rc = evavos_bank_write_keys(&a.bank, reinterpret_cast<const float*>(d + o_key), n_pos, 0, n_pos, nullptr, 0, st);
  if (rc) return rc;
  if (readout) {
    rc = evavos_bank_write_values(&a.bank, reinterpret_cast<const float*>(d + o_val), (int64_t)CV * n_pos, n_pos, 0,
                                  n_pos, nullptr, 0, 0, st);
    if (rc) return rc;
  }
  rc = evavos_memread(&a, st);
  if (rc) return rc;
  int64_t d2h = 0;
  if (readout) { EVAVOS_CUDA_OK(cudaMemcpyAsync(readout, d + o_out, b_out, cudaMemcpyDeviceToHost, st)); d2h += b_out; }
  if (topk_idx) { EVAVOS_CUDA_OK(cudaMemcpyAsync(topk_idx, d + o_idx, b_tk, cudaMemcpyDeviceToHost, st)); d2h += b_tk; }
  if (topk_weight) { EVAVOS_CUDA_OK(cudaMemcpyAsync(topk_weight, d + o_w, b_tk, cudaMemcpyDeviceToHost, st)); d2h += b_tk; }
  EVAVOS_CUDA_OK(cudaStreamSynchronize(st));
  if (h2d_bytes) *h2d_bytes = (int64_t)(b_key + b_q + b_val);
  if (d2h_bytes) *d2h_bytes = d2h;
  return EVAVOS_OK;
}

}  // extern "C"
