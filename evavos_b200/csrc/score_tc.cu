// tcgen05 / TMEM / TMA candidate filter for the key affinity (sm_100a).
//
// S'[q][n] = q^.k^_n - |k_n|^2/2  (bf16 operands, fp32 accumulate in TMEM; the true affinity is
// (2 S' - |q|^2)/sqrt(CK), a per-query monotone map, prop_net.py:86-90).  The -|k|^2/2 term is
// part of the contraction: every key row carries a 16-wide extra K slice (hi, mid, lo bf16 split
// of -|k|^2/2, zeros) that meets (1, 1, 1, 0...) on the query side, so the accumulator tile IS the
// score tile and the epilogue spends one instruction per score.  The THW x HW matrix never leaves
// the SM: each 128x128 accumulator tile is consumed out of TMEM and only O(k) numbers per query
// reach HBM.
//
//   sweep 1: running maximum of S' per (query, column class n mod 128) -> class_max.
//            The k-th largest of a query's 128 class maxima is a lower bound on its k-th best
//            score (k distinct positions reach it); the CTAs agree on it across the grid.
//   sweep 2: same contraction; every position with S' >= tau_q (bound minus a rigorous bf16
//            error margin) is appended to the query's candidate list, which therefore
//            contains the exact fp32 top-k.  finalize_kernel's exact rescoring picks it.
// Both sweeps run inside ONE cooperative launch (score_select_kernel): the TMA and MMA warps simply stream the
// chunk's tiles twice and run ahead into sweep 2 while the epilogue warps exchange thresholds.
// (Running the exact finalizer inside this kernel as well was measured slower: a CTA finalizes its dozen queries
// in sequence, while the separate finalize_kernel runs all queries at once.)
//
// Roles per CTA (640 threads, 1 CTA/SM, one wave): warp 0 = TMA producers (4 lanes issuing cp.async.bulk of
// pre-swizzled 20 KB key tile images; one issuing thread keeps a single copy in flight, ~50 B/clk),
// warps 1-2 = MMA issuers (one elected lane each, alternating tiles, 5 x tcgen05.mma 128x128x16 per tile),
// warp 3 = TMEM allocator, warps 4-19 = epilogue (four warpgroups, each thread owns one query row and 32
// accumulator columns; branch-free inner loops).  Rings: 6 shared-memory key stages, 4 TMEM
// accumulator stages (4 x 128 columns = all 512).  The query tile is converted to bf16 and
// swizzled into shared memory by the CTA itself.
#include "common.cuh"

namespace evavos {

#ifdef EVAVOS_TRACE
// Timeline of CTA 0 (clock64): rows = producer issue, MMA waits done, MMA issued, epilogue acc_full seen,
// epilogue TMEM load done, epilogue math done; columns = tile index (first 64 tiles).
__device__ long long g_trace[6][64];
#define EVAVOS_TR(row, i) do { if (blockIdx.x == 0 && (i) < 64) g_trace[row][i] = clock64(); } while (0)
#else
#define EVAVOS_TR(row, i) do { } while (0)
#endif

// Compile-time timing experiments (-DEVAVOS_EXP=bits; results are wrong while a bit is set):
// 1 = the epilogue skips the TMEM load, 2 = skips its math.
#ifndef EVAVOS_EXP
#define EVAVOS_EXP 0
#endif

namespace {

constexpr int kStages = 8;
constexpr int kAccStages = 3;   // 3 x 128 accumulator columns; the query operand lives in columns [384, 424)
constexpr int kQueryCol = 384;
// Epilogue organisation (EVAVOS_GROUPS):
//   1: 16 warps visit every tile in lock-step, 32 accumulator columns each; two MMA issuer warps.
//   2: two groups of 8 warps serve even / odd tiles, 64 columns per warp as 2 x 32; one in-order MMA issuer.
//   3: three groups of 8 warps, group g owns accumulator stage g (iterations i = g mod 3), 64 columns per warp as
//      4 x 16 so that a thread fits in 72 registers and 28 warps are resident; one in-order MMA issuer.
#ifndef EVAVOS_RARE
#define EVAVOS_RARE (-1)   // -1: decide per launch from the bank size; 0 / 1: force (timing experiments)
#endif
#ifndef EVAVOS_GROUPS
#define EVAVOS_GROUPS 2
#endif
constexpr int kNumGroups = EVAVOS_GROUPS;
constexpr int kEpiWarps = kNumGroups == 3 ? 24 : 16;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kThreads = 128 + kEpiThreads;
constexpr int kProducers = 4;   // TMA-issuing lanes of warp 0 (kStages % kProducers == 0)
constexpr int kCols = kNumGroups == 3 ? 16 : 32;   // accumulator columns per tcgen05.ld and epilogue thread
constexpr int kClasses = kNumGroups == 3 ? 96 : 128;   // column classes per query and chunk (sweep 1)
// Scores per staged hit group (sweep 2 tests one maximum per group).  Measured, filter time in us for 4 | 8:
// cfg2 (32 k positions, a hit in 23 % | 40 % of the warp-groups) 43.8 | 52.7, cfg4 (324 k positions) 168.8 | 162.9.
#ifndef EVAVOS_GROUP
#define EVAVOS_GROUP 4
#endif
constexpr int kGroup = EVAVOS_GROUP;
constexpr int kPend = 8 + kCols / kGroup;  // staged hit groups per epilogue thread (flushed when more than 8 wait)
constexpr bool kSplit = kNumGroups > 1;   // epilogue warp groups on different tiles (see the epilogue)
constexpr int kBarBytes = 256;
constexpr int kSmemBytes = kTileBytes * kStages + kBarBytes + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// One lane of a converged warp (the tcgen05 issue idiom: the branch stays warp-uniform for the compiler).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T, bf16 x bf16 -> fp32, one CTA: A from TMEM (128 lanes x 8 columns of packed bf16 pairs per K = 16 step).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major operand descriptors (16-byte units; version 1 = Blackwell).
//   SWIZZLE_128B: rows of 64 bf16 = 128 B, 8-row groups 1024 B apart.
//   SWIZZLE_32B : rows of 16 bf16 =  32 B, 8-row groups  256 B apart (the -|k|^2/2 slice).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}
constexpr uint64_t kLayoutSw128 = 2, kLayoutSw32 = 6;

// kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, both K-major, M=128, N=128.
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v) {
  if constexpr (kCols == 16) tmem_ld16(taddr, v);
  else tmem_ld32(taddr, v);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct PassParams {
  const float* query;       // (64, n_query) fp32, row stride query_ch_stride
  int64_t query_ch_stride;
  const uint8_t* key_tiles;
  int64_t n_pos;
  int64_t n_query;
  int64_t nq_pad;
  int n_mtiles;
  int n_ktiles;
  int n_chunks;
  float* class_max;
  float* tau;
  int32_t* cand;
  int32_t* cand_cnt;
  float4* pend_score;   // [grid][kPend * kGroup / 4][kEpiThreads] scores of staged hit groups (sweep 2)
  int32_t* pend_pos;    // [grid][kPend][kEpiThreads] first position of each staged group
  const float* key_maxnorm;
  unsigned int* grid_counter;  // one counter per query tile of the launch, zeroed before every launch
  int m_tile0;          // first query tile of this launch
  int top_k;
  int rare_hits;        // sweep 2 tests a whole block before its groups (see the epilogue)
};

// Sweep 2 stages every group of kGroup adjacent scores whose maximum reaches the threshold (scores + first
// position) in a private strip of the workspace; the strip is resolved into the query's candidate list out of
// line.  Hits are rare: about top_k + margin per query over the whole bank.
__device__ __noinline__ void flush_pending(const float4* ps, const int32_t* pp, int n, float thr, int32_t* cand,
                                           int32_t* cand_cnt, int64_t q) {
  int hits = 0;
  for (int e = 0; e < n * (kGroup / 4); ++e) {
    const float4 s = ps[e * kEpiThreads];
    hits += (s.x >= thr) + (s.y >= thr) + (s.z >= thr) + (s.w >= thr);
  }
  int at = atomicAdd(cand_cnt + q, hits);
  for (int e = 0; e < n * (kGroup / 4); ++e) {
    const float4 s = ps[e * kEpiThreads];
    const int32_t n0 = pp[(e / (kGroup / 4)) * kEpiThreads] + 4 * (e % (kGroup / 4));
    const float v[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (v[k] >= thr) {
        if (at < kCandCap) cand[q * kCandCap + at] = n0 + k;
        ++at;
      }
    }
  }
}

// 32 accumulator columns of one query row, held as registers.
template <int PASS, bool PARTIAL>
__device__ __forceinline__ void consume_tile(const float* v, float* cmax, float thr, int32_t n_first, int valid,
                                             float4* ps, int32_t* pp, int& pending) {
  // valid: number of in-range columns among this thread's kCols (only read when PARTIAL)
  if constexpr (PASS == 1) {
    // kCols / 2 classes per call, two columns each: one 3-input max per two scores
#pragma unroll
    for (int j = 0; j < kCols / 2; ++j) {
      const float s0 = (PARTIAL && j >= valid) ? kEmptyNh : v[j];
      const float s1 = (PARTIAL && j + kCols / 2 >= valid) ? kEmptyNh : v[j + kCols / 2];
      cmax[j] = fmaxf(fmaxf(cmax[j], s0), s1);
    }
  } else {
#pragma unroll
    for (int g = 0; g < kCols / kGroup; ++g) {
      float sg[kGroup];
#pragma unroll
      for (int e = 0; e < kGroup; ++e) sg[e] = (PARTIAL && g * kGroup + e >= valid) ? -INFINITY : v[g * kGroup + e];
      float m = sg[0];
#pragma unroll
      for (int e = 1; e < kGroup; ++e) m = fmaxf(m, sg[e]);
      if (m >= thr) {
#pragma unroll
        for (int h = 0; h < kGroup / 4; ++h)
          ps[(pending * (kGroup / 4) + h) * kEpiThreads] = make_float4(sg[4 * h], sg[4 * h + 1], sg[4 * h + 2], sg[4 * h + 3]);
        pp[pending * kEpiThreads] = n_first + g * kGroup;
        ++pending;
      }
    }
  }
}

// Barrier among the epilogue threads of the CTAs that share a query tile (the grid is launched cooperatively,
// so every CTA is resident).  `counter` only grows: the n-th barrier waits for n * n_chunks arrivals.
__device__ __forceinline__ void epilogue_grid_barrier(unsigned int* counter, unsigned int target) {
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  if (threadIdx.x == 128) {
    // release: cumulative over the CTA's writes ordered before the bar.sync above
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}

// Admission threshold of one query (one warp): k-th largest of its 128 class maxima (4 per lane, maximum over the
// memory-axis chunks) by an in-warp bitonic sort, minus the bf16 error margin.
__device__ __forceinline__ void warp_threshold(const PassParams& p, int64_t q, int lane) {
  const float qa = __ldg(p.query + (int64_t)lane * p.query_ch_stride + q);
  const float qb = __ldg(p.query + (int64_t)(lane + 32) * p.query_ch_stride + q);
  float v[4] = {kEmptyNh, kEmptyNh, kEmptyNh, kEmptyNh};
  {
    // 8 chunks x 4 classes = 32 independent L2 loads per round (the rows were written by other CTAs: bypass L1)
    const float* row0 = p.class_max + q * 128 + lane;
    const int64_t chunk_stride = p.nq_pad * 128;
    for (int g = 0; g < p.n_chunks; g += 8) {
      float w[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t)
          w[u][t] = (g + u < p.n_chunks && 32 * t < kClasses) ? __ldcg(row0 + (g + u) * chunk_stride + 32 * t) : kEmptyNh;
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) v[t] = fmaxf(v[t], w[u][t]);
    }
  }
  float qsq = fmaf(qa, qa, qb * qb);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) qsq += __shfl_xor_sync(0xffffffffu, qsq, o);
  // k-th largest of the warp's 128 values: bitonic sort, descending, element e = 4 * lane + t.
  // Partners e ^ j with j < 4 sit in the same lane (register swap), j >= 4 in lane ^ (j / 4) (shuffle).
#pragma unroll
  for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= 1; j >>= 1) {
      if (j >= 4) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int e = 4 * lane + t;
          const float other = __shfl_xor_sync(0xffffffffu, v[t], j >> 2);
          const bool keep_max = ((e & k) == 0) == ((e & j) == 0);   // descending blocks keep the larger in front
          v[t] = keep_max ? fmaxf(v[t], other) : fminf(v[t], other);
        }
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if ((t & j) == 0) {
            const int e = 4 * lane + t;
            const float a = v[t], b = v[t | j];
            const bool desc = (e & k) == 0;
            v[t] = desc ? fmaxf(a, b) : fminf(a, b);
            v[t | j] = desc ? fminf(a, b) : fmaxf(a, b);
          }
        }
      }
    }
  }
  const int kth = p.top_k - 1;
  float sel = v[0];
#pragma unroll
  for (int t = 1; t < 4; ++t) sel = ((kth & 3) == t) ? v[t] : sel;
  const float kth_value = __shfl_sync(0xffffffffu, sel, kth >> 2);
  if (lane == 0) {
    const float qn = sqrtf(qsq) * 1.0001f;
    const float kn = *p.key_maxnorm;
    // |q^.k^ - q.k| <= 2^-8 (1 + 2^-10) |q||k| for bf16 round-to-nearest operands, plus slack for the
    // tensor-core fp32 accumulation and the rounding of -|k|^2/2.
    const float eps = 0.004f * qn * kn + 2.0e-6f * kn * kn + 1.0e-30f;
    p.tau[q] = kth_value - 2.0f * eps;
    p.cand_cnt[q] = 0;
  }
}

// One persistent CTA per (query tile, memory chunk).  The producer and MMA warps stream the chunk's key tiles
// twice; the epilogue warps take running class maxima on the first sweep, agree on per-query thresholds across
// the grid, and collect candidates on the second sweep.
__global__ void __launch_bounds__(kThreads, 1) score_select_kernel(const PassParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t stage0 = base;
  const uint32_t bars = base + kTileBytes * kStages;
  const uint32_t bar_full = bars;                                // [kStages]
  const uint32_t bar_empty = bars + 8 * kStages;                 // [kStages]
  const uint32_t bar_acc_full = bars + 16 * kStages;             // [kAccStages]
  const uint32_t bar_acc_empty = bar_acc_full + 8 * kAccStages;  // [kAccStages]
  const uint32_t tmem_slot = bar_acc_empty + 8 * kAccStages;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kTileBytes * kStages + 16 * kStages + 16 * kAccStages);

  if (threadIdx.x == 0) EVAVOS_TR(0, 56);
  // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps everything derived from it
  // (loop counters, stage addresses, UMMA descriptors) in uniform registers and issues the five tcgen05.mma of a
  // tile back to back instead of wrapping each in an R2UR broadcast loop (~95 -> ~40 clk of issue per MMA)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m_tile = p.m_tile0 + blockIdx.x % p.n_mtiles;
  const int chunk = blockIdx.x / p.n_mtiles;
  const int t0 = (int)(((int64_t)chunk * p.n_ktiles) / p.n_chunks);
  const int t1 = (int)(((int64_t)(chunk + 1) * p.n_ktiles) / p.n_chunks);
  const int n_tiles = t1 - t0;
  const int n_iter = 2 * n_tiles;  // every key tile is contracted twice
  auto tile_of = [&](int i) { return t0 + (i >= n_tiles ? i - n_tiles : i); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, kSplit ? 8 : kEpiWarps);   // warps that visit one tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // The first ring of key tiles does not depend on anything below: get it in flight now.
  if (threadIdx.x == 0) {
    const int pre = n_iter < kStages ? n_iter : kStages;
    for (int i = 0; i < pre; ++i) {
      mbar_arrive_expect_tx(bar_full + 8 * i, kTileBytes);
      bulk_g2s(stage0 + i * kTileBytes, p.key_tiles + (int64_t)tile_of(i) * kTileBytes, kTileBytes, bar_full + 8 * i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  // Query operand: thread r of the first epilogue warpgroup converts query row q0 + r (64 channels, read in the
  // caller's layout, coalesced across the warp) to bf16 pairs and stores them, followed by the (1, 1, 1, 0...)
  // slice that meets the keys' -|k|^2/2 slice, into TMEM lane r.
  if (warp >= 4 && warp < 8) {
    const int r = (warp - 4) * 32 + lane;
    const int64_t qrow = (int64_t)m_tile * 128 + r;
    const bool live = qrow < p.n_query;
    float f[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) f[c] = live ? __ldg(p.query + (int64_t)c * p.query_ch_stride + qrow) : 0.f;
    const uint32_t a_addr = tmem_base + ((uint32_t)((warp - 4) * 32) << 16) + kQueryCol;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 v2 = __floats2bfloat162_rn(f[k * 16 + 2 * j], f[k * 16 + 2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&v2);
      }
      tmem_st8(a_addr + 8 * k, w);
    }
    const uint32_t aug[8] = {live ? 0x3f803f80u : 0u, live ? 0x00003f80u : 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    tmem_st8(a_addr + 32, aug);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) EVAVOS_TR(0, 57);

  if (warp == 0) {
    // ===== TMA producers: kProducers lanes, lane l streams iterations i = l (mod kProducers) =====
    // (one thread keeps only one bulk copy in flight; several lanes keep several)
    if (lane < kProducers) {
      // iterations below kStages were issued in the prologue
      for (int i = kStages + lane; i < n_iter; i += kProducers) {
        const int s = i % kStages;
        const uint32_t ph = (uint32_t)((i / kStages) & 1);
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        EVAVOS_TR(0, i);
        mbar_arrive_expect_tx(bar_full + 8 * s, kTileBytes);
        bulk_g2s(stage0 + s * kTileBytes, p.key_tiles + (int64_t)tile_of(i) * kTileBytes, kTileBytes, bar_full + 8 * s);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===== MMA issuers: warp 1 takes the even iterations, warp 2 the odd ones =====
    // (measured: one thread spends ~350 clk per tile in its two barrier waits; two threads overlap them with the
    //  other's MMAs: 830 -> 600 clk per tile.  Three threads, or one thread interleaving two tiles, were slower.
    //  Each commit tracks the MMAs of its own thread, which is exactly one tile.)
    const uint32_t a_tmem = tmem_base + kQueryCol;
    // kSplit: ONE issuer, every iteration in order.  The epilogue groups then see an accumulator stage only at every
    // other use, and an mbarrier parity wait is only sound for a waiter that cannot fall two phases behind: with
    // two issuers tile i + 1 may complete before tile i, a group runs ahead onto a stage whose previous phase it
    // never observed, takes the stale parity for "ready" and the pipeline deadlocks (seen on B200).  In-order
    // commits from a single thread rule that out.
    for (int i = kSplit ? (warp == 1 ? 0 : n_iter) : warp - 1; i < n_iter; i += kSplit ? 1 : 2) {
      const int s = i % kStages, a = i % kAccStages;
      mbar_wait(bar_full + 8 * s, (uint32_t)((i / kStages) & 1));
      mbar_wait(bar_acc_empty + 8 * a, (uint32_t)(((i / kAccStages) & 1) ^ 1));
      tc_fence_after();
      if (elect_one()) {
        EVAVOS_TR(1, i);
        const uint32_t st = stage0 + s * kTileBytes;
        const uint64_t bdesc0 = make_desc(st, 1024, kLayoutSw128);
        const uint64_t bdesc_aug = make_desc(st + kTileKeyBytes, 256, kLayoutSw32);
        const uint32_t d = tmem_base + a * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // K = 16 bf16 = 32 B per MMA: advance 2 x 16-byte units inside the swizzle atom
          umma_bf16_ts(d, a_tmem + 8 * k, bdesc0 + 2 * k, kInstrDesc, k > 0 ? 1u : 0u);
        umma_bf16_ts(d, a_tmem + 32, bdesc_aug, kInstrDesc, 1u);  // += -|k|^2/2
        umma_commit(bar_empty + 8 * s);      // smem stage free once these MMAs have read it
        umma_commit(bar_acc_full + 8 * a);   // accumulator tile complete
        EVAVOS_TR(2, i);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> running class max (sweep 1) / staged hit groups (sweep 2) =====
    // kSplit: the 16 warps form two groups of 8 that serve alternate tiles (even / odd tile index), each warp 64
    // accumulator columns as two 32-column loads, so that one group's math overlaps the other group's loads and
    // the MMAs of the next tile.  Otherwise all 16 warps visit every tile in lock-step, 32 columns each.
    const int ew = warp - 4;
    const int quarter = ew & 3;           // TMEM lane quarter this warp may access
    const int grp = kSplit ? (ew >> 3) : 0;
    const int colbase = kSplit ? ((ew >> 2) & 1) * 64 : (ew >> 2) * kCols;
    constexpr int kBlocks = kSplit ? 64 / kCols : 1;   // tcgen05.ld blocks per visited tile
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 128;
    const int64_t q = (int64_t)m_tile * 128 + row;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // iterations this warp visits: 2 groups - tiles of the group's parity; 3 groups - i = grp (mod 3), i.e. always
    // accumulator stage `grp`, in both sweeps; lock-step - all
    constexpr int i_step = kSplit ? kNumGroups : 1;
    const int i_first = kNumGroups == 3 ? grp : (kNumGroups == 2 ? ((((t0 & 1) == grp) ? 0 : 1)) : 0);
    const int i_first2 = kNumGroups == 3 ? n_tiles + (grp + 3 - n_tiles % 3) % 3 : n_tiles + i_first;

    // visit(i, math): wait for accumulator tile i, read this warp's columns block by block (the stage goes back to
    // the MMA warps as soon as the last block is in registers) and call math(block, values, first position).
    auto visit = [&](int i, auto&& math) {
      const int a = i % kAccStages;
      mbar_wait(bar_acc_full + 8 * a, (uint32_t)((i / kAccStages) & 1));
      tc_fence_after();
      if (threadIdx.x == 128) EVAVOS_TR(3, i);
#pragma unroll
      for (int blk = 0; blk < kBlocks; ++blk) {
        float v[kCols];
        if constexpr (!(EVAVOS_EXP & 1)) {
          tmem_ld_cols(lane_addr + (uint32_t)(a * 128 + colbase + blk * kCols), v);
          tmem_ld_wait();
        } else {
          for (int j = 0; j < kCols; ++j) v[j] = kEmptyNh;
        }
        if (blk == kBlocks - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * a);  // registers hold the tile: release the TMEM stage
          if (threadIdx.x == 128) EVAVOS_TR(4, i);
        }
        math(blk, v, (int64_t)tile_of(i) * kTilePos + colbase + blk * kCols);
      }
      if (threadIdx.x == 128) EVAVOS_TR(5, i);
    };

    // ---- sweep 1: class maxima ----
    {
      // Column classes per thread: 16 x (columns j and j + 16 of a 32-column block), one set per block (kSplit: of
      // the tiles of this group's parity) or per tile parity (lock-step) - 128 per query and chunk either way.  Any
      // partition of the positions into classes gives a valid bound; this one costs one FMNMX3 per two scores.
      constexpr int kOwn = 32;   // class maxima held per thread
      float cmax[kOwn];
#pragma unroll
      for (int j = 0; j < kOwn; ++j) cmax[j] = kEmptyNh;
      auto class_max = [&](float* cm, const float* v, int64_t n0) {
        int pending = 0;
        if constexpr ((EVAVOS_EXP & 2) != 0) {
        } else if (n0 + kCols > p.n_pos) {
          const int valid = (int)max((int64_t)0, p.n_pos - n0);
          consume_tile<1, true>(v, cm, 0.f, (int32_t)n0, valid, nullptr, nullptr, pending);
        } else {
          consume_tile<1, false>(v, cm, 0.f, (int32_t)n0, kCols, nullptr, nullptr, pending);
        }
      };
      if constexpr (kSplit) {
        for (int i = i_first; i < n_tiles; i += i_step)
          visit(i, [&](int blk, const float* v, int64_t n0) { class_max(cmax + blk * (kCols / 2), v, n0); });
      } else {
        int i = 0;   // parity of the tile's index in the bank, not in the chunk
        if (t0 & 1) visit(i++, [&](int, const float* v, int64_t n0) { class_max(cmax + kCols / 2, v, n0); });
        for (; i < n_tiles; i += 2) {
          visit(i, [&](int, const float* v, int64_t n0) { class_max(cmax, v, n0); });
          if (i + 1 < n_tiles) visit(i + 1, [&](int, const float* v, int64_t n0) { class_max(cmax + kCols / 2, v, n0); });
        }
      }
      if constexpr (kNumGroups == 3) {
        // 6 (group, half) slots of 16 classes: neighbouring classes of a thread are merged pairwise (96 per query)
        float4* dst = reinterpret_cast<float4*>(p.class_max + ((int64_t)chunk * p.nq_pad + q) * 128 + (ew >> 2) * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
          dst[j4] = make_float4(fmaxf(cmax[8 * j4], cmax[8 * j4 + 1]), fmaxf(cmax[8 * j4 + 2], cmax[8 * j4 + 3]),
                                fmaxf(cmax[8 * j4 + 4], cmax[8 * j4 + 5]), fmaxf(cmax[8 * j4 + 6], cmax[8 * j4 + 7]));
      } else {
        float4* dst = reinterpret_cast<float4*>(p.class_max + ((int64_t)chunk * p.nq_pad + q) * 128 + (ew >> 2) * 32);
#pragma unroll
        for (int j4 = 0; j4 < kOwn / 4; ++j4)
          dst[j4] = make_float4(cmax[j4 * 4], cmax[j4 * 4 + 1], cmax[j4 * 4 + 2], cmax[j4 * 4 + 3]);
      }
    }

    // ---- thresholds: every CTA of a query tile takes a slice of its 128 rows ----
    if (threadIdx.x == 128) EVAVOS_TR(0, 60);
    epilogue_grid_barrier(p.grid_counter + (blockIdx.x % p.n_mtiles), (unsigned)p.n_chunks);
    if (threadIdx.x == 128) EVAVOS_TR(0, 61);
    {
      const int r0 = (chunk * 128) / p.n_chunks, r1 = ((chunk + 1) * 128) / p.n_chunks;
      for (int r = r0 + ew; r < r1; r += kEpiThreads / 32) {
        const int64_t qq = (int64_t)m_tile * 128 + r;
        if (qq < p.n_query) warp_threshold(p, qq, lane);
      }
    }
    if (threadIdx.x == 128) EVAVOS_TR(0, 62);
    epilogue_grid_barrier(p.grid_counter + (blockIdx.x % p.n_mtiles), 2u * (unsigned)p.n_chunks);
    if (threadIdx.x == 128) EVAVOS_TR(0, 63);

    // ---- sweep 2: candidates ----
    {
      float thr = INFINITY;
      if (q < p.n_query) thr = __ldcg(p.tau + q);
      float4* ps = p.pend_score + (int64_t)blockIdx.x * (kPend * kGroup / 4) * kEpiThreads + et;
      int32_t* pp = p.pend_pos + (int64_t)blockIdx.x * kPend * kEpiThreads + et;
      int pending = 0;
      float unused[1];
      for (int i = i_first2; i < n_iter; i += i_step) {
        visit(i, [&](int, const float* v, int64_t n0) {
          if constexpr ((EVAVOS_EXP & 2) != 0) {
          } else if (n0 + kCols > p.n_pos) {
            const int valid = (int)max((int64_t)0, p.n_pos - n0);
            consume_tile<2, true>(v, unused, thr, (int32_t)n0, valid, ps, pp, pending);
          } else {
            // banks where a hit is rare per block: one maximum over the whole block first, the per-group test only
            // when it reaches the threshold (the branch is warp-divergent, but most warps skip the block)
            bool look = true;
            if (p.rare_hits) {
              float m = v[0];
#pragma unroll
              for (int j = 1; j < kCols; ++j) m = fmaxf(m, v[j]);
              look = m >= thr;
            }
            if (look) consume_tile<2, false>(v, unused, thr, (int32_t)n0, kCols, ps, pp, pending);
          }
          if (pending > kPend - kCols / kGroup) {  // the next block stages at most kCols / kGroup groups
            flush_pending(ps, pp, pending, thr, p.cand, p.cand_cnt, q);
            pending = 0;
          }
        });
      }
      if (pending > 0) flush_pending(ps, pp, pending, thr, p.cand, p.cand_cnt, q);
      if (threadIdx.x == 128) EVAVOS_TR(0, 58);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 3) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
  if (threadIdx.x == 96) EVAVOS_TR(0, 59);
}

}  // namespace

// Memory-axis chunks per query tile: one wave of CTAs (m_tiles * chunks <= n_sm) whenever possible.
int score_pass_chunks(int64_t n_pos, int64_t n_query, int n_sm) {
  const int64_t mt = ceil_div(n_query, 128), nt = ceil_div(n_pos, kTilePos);
  int64_t g = n_sm / mt;
  if (g < 1) g = 1;
  if (g > nt) g = nt;
  return (int)g;
}

size_t score_pass_pending_bytes(int64_t n_query, int n_chunks) {
  return (size_t)ceil_div(n_query, 128) * n_chunks * kPend * kEpiThreads * (kGroup / 4 * sizeof(float4) + sizeof(int32_t));
}

#ifdef EVAVOS_TRACE
extern "C" int evavos_debug_trace(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 6 * 64);
}
#endif

// Candidate generation for all queries: class maxima, thresholds and candidate lists in one cooperative launch
// per wave of query tiles (one wave whenever n_query <= 128 * n_sm).
int launch_score_select(const float* query, int64_t query_ch_stride, const void* key_tiles, const float* key_maxnorm,
                        int64_t n_pos, int64_t n_query, int top_k, int n_chunks, int n_sm, float* class_max, float* tau,
                        int32_t* cand, int32_t* cand_cnt, void* pending, unsigned int* grid_counter, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    EVAVOS_CUDA_OK(cudaFuncSetAttribute(score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int mt_total = (int)ceil_div(n_query, 128);
  const int mt_per_launch = n_chunks > 1 ? mt_total : (mt_total < n_sm ? mt_total : n_sm);
  for (int m0 = 0; m0 < mt_total; m0 += mt_per_launch) {
    PassParams p;
    p.query = query;
    p.query_ch_stride = query_ch_stride;
    p.key_tiles = reinterpret_cast<const uint8_t*>(key_tiles);
    p.n_pos = n_pos;
    p.n_query = n_query;
    p.n_mtiles = mt_total - m0 < mt_per_launch ? mt_total - m0 : mt_per_launch;
    p.nq_pad = (int64_t)mt_total * 128;
    p.n_ktiles = (int)ceil_div(n_pos, kTilePos);
    p.n_chunks = n_chunks;
    p.class_max = class_max;
    p.tau = tau;
    p.cand = cand;
    p.cand_cnt = cand_cnt;
    p.key_maxnorm = key_maxnorm;
    p.grid_counter = grid_counter;
    p.m_tile0 = m0;
    p.top_k = top_k;
    // expected candidates per query ~ 1.4 top_k, a block is 32 lanes x kCols scores: rare = a block holds a hit
    // with probability below ~1/2
    p.rare_hits = (EVAVOS_RARE >= 0) ? EVAVOS_RARE : ((double)n_pos > 2.0 * 1.4 * top_k * 32.0 * kCols ? 1 : 0);
    const unsigned grid = (unsigned)(p.n_mtiles * n_chunks);
    p.pend_score = reinterpret_cast<float4*>(pending);
    p.pend_pos = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(pending) +
                                            (size_t)mt_per_launch * n_chunks * (kPend * kGroup / 4) * kEpiThreads * sizeof(float4));
    EVAVOS_CUDA_OK(cudaMemsetAsync(grid_counter, 0, sizeof(unsigned int) * (size_t)p.n_mtiles, st));
    void* args[] = {&p};
    EVAVOS_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(score_select_kernel), dim3(grid),
                                               dim3(kThreads), args, kSmemBytes, st));
  }
  return EVAVOS_OK;
}

}  // namespace evavos
