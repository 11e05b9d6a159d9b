// Microbenchmark: tcgen05.mma issue/execute rate for the shapes the score filter uses (sm_100a).
// Back-to-back 128x128x16 (or 128x256x16) bf16 MMAs from one or two issuing threads per CTA; variants:
//   form     : SS (A, B from shared memory) or TS (A from TMEM)
//   chain    : all MMAs accumulate into ONE accumulator, or rotate over 2 / 3 accumulators
//   flags    : 1 commit after every 5 MMAs, 2 SWIZZLE_32B descriptor for the 5th, 4 first MMA overwrites,
//              8 two issuing threads, 16 random operands
//   beside   : reader warps streaming another accumulator with tcgen05.ld, lanes keeping 20 KB bulk copies in
//              flight, warps polling an mbarrier; one CTA or one per SM (grid 143)
// Prints cycles per MMA (issue to completion of the whole batch); the floor is 64 (N=128) / 128 (N=256).
// B200 results (r1): 64.1-64.3 at the floor in every isolated variant; 82-99 with 16 reader warps when ONE thread
// issues (it shares its scheduler with them), 64.4 with two issuing threads; TMEM reads beside full-rate MMAs
// ~300 B/clk (16 warps, one tcgen05.ld.x32 in flight each).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t sbo, uint64_t layout) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}
constexpr uint32_t kIdesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
template <int form, int n_acc, int n256>
__global__ void __launch_bounds__(640, 1) k(int n_mma, int n_readers, int flags, int n_copy_lanes, int n_pollers, const uint8_t* gbuf, long long* cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint64_t pbar;
  __shared__ uint64_t bar3;
  __shared__ uint64_t cbar[8];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) done = 0;
  for (int i = threadIdx.x; i < 144 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (flags & 16) ? ((h & 0x807f807fu) | 0x3f003f00u) : 0x3f803f80u;  // random +-[0.5,1) bf16 pairs
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar2)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&pbar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar3)) : "memory");
    for (int c = 0; c < 8; ++c) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&cbar[c])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  long long t0 = 0, t1 = 0;
  const int n_issuers = (flags & 8) ? 2 : 1;
  if (threadIdx.x == 32 || (n_issuers == 2 && threadIdx.x == 64)) {
    const int issuer = threadIdx.x == 64;
    const uint32_t mybar = s32(issuer ? &bar3 : &bar);
    const uint64_t adesc = desc(s32(smem), 1024, 2), bdesc = desc(s32(smem) + 32768, 1024, 2);
    const uint32_t idesc = n256 ? kIdesc256 : kIdesc128;
    const int width = n256 ? 256 : 128;
    t0 = clock64();
    const uint64_t augdesc = desc(s32(smem) + 32768 + 16384, 256, 6);  // SWIZZLE_32B slice like the filter's 5th MMA
    for (int t = 0; t < n_mma / 5 / n_issuers; ++t) {
      const uint32_t d = tmem + (uint32_t)(((t % n_acc) + issuer * n_acc) * width);
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const uint64_t bd = (kk == 4 && (flags & 2)) ? augdesc : bdesc + 2 * (kk & 3);
        const uint32_t acc = (kk == 0 && (flags & 4)) ? 0u : 1u;
        if (form == 0) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(adesc + 2 * (kk & 3)), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                       ::"r"(d), "r"(tmem + 448 + 8 * (kk & 3)), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
      }
      if (flags & 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2)) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar2)) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mybar) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(mybar) : "memory");
    t1 = clock64();
    atomicMax(reinterpret_cast<unsigned long long*>(cycles), (unsigned long long)(t1 - t0));
    atomicAdd(const_cast<int*>(&done), 1);
  } else if (warp == 3 && (threadIdx.x & 31) < n_copy_lanes) {
    // TMA traffic beside the MMAs: each lane keeps one 20 KB bulk copy in flight into its own stage (above the operands)
    const int l = threadIdx.x & 31;
    const uint32_t dst = s32(smem) + 57344 + (uint32_t)l * 20480u;   // needs 57344 + 4 * 20480 <= dynamic smem
    uint32_t ph = 0;
    long long copies = 0;
    while (done < n_issuers) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&cbar[l])), "r"(20480u) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(gbuf + (size_t)((copies * 4 + l) % 512) * 20480), "r"(20480u), "r"(s32(&cbar[l])) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&cbar[l])), "r"(ph) : "memory");
      ph ^= 1u;
      ++copies;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(cycles + 2), (unsigned long long)copies);
  } else if (warp >= 20 - n_pollers) {
    // whole warps spinning on an mbarrier that never completes, like roles waiting for their turn
    long long polls = 0;
    while (done < n_issuers) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&pbar)) : "memory");
      polls += 1 + ok;
    }
    if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(cycles + 1), 0ull * polls);
  } else if (warp >= 4 && warp < 4 + n_readers) {
    // epilogue-like readers: stream a 128x128 fp32 accumulator that the MMAs are NOT writing (columns 384..511 when
    // n_acc*width <= 384), 32 columns per tcgen05.ld, until the issuer is done
    float acc = 0.f;
    float alu[8] = {1.f, 2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f};
    long long loads = 0;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    while (done < n_issuers) {
      uint32_t r[32], r2[32];
      const uint32_t col = 384u + 32u * (uint32_t)(((warp - 4) >> 2) & 3);
#define LDTM32(R, ADDR)                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]), "=r"(R[9]),       \
                 "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), \
                 "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), \
                 "=r"(R[30]), "=r"(R[31])                                                                                                         \
               : "r"(ADDR) : "memory")
      if (!(flags & 128)) LDTM32(r, tmem + lane_base + col);
      if (flags & 32) LDTM32(r2, tmem + lane_base + (col ^ 32u));
      if (flags & (64 | 128)) {
        // ~130 dependent-free ALU instructions per lane that do not touch the load's registers
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          alu[j & 7] = fmaf(alu[j & 7], 1.0001f, 0.5f);
          // yield experiments: every 16 instructions give the scheduler a reason to switch warps
          if ((j & 15) == 15) {
            if (flags & 256) alu[0] = __shfl_sync(0xffffffffu, alu[0], threadIdx.x & 31);   // dependent shuffle
            if (flags & 512) __syncwarp();
            if (flags & 1024) asm volatile("nanosleep.u32 0;" ::: "memory");
          }
        }
      }
      if (!(flags & 128)) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r[j]));
        if (flags & 32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = fmaxf(acc, __uint_as_float(r2[j]));
        }
      }
      ++loads;
    }
    acc += alu[0] + alu[1] + alu[2] + alu[3] + alu[4] + alu[5] + alu[6] + alu[7];
    if (acc == 123.f) sink[threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(cycles + 1), (unsigned long long)loads);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
static uint8_t* g_buf;
template <int form, int n_acc, int n256>
void run(long long* d, int n_readers = 0, int flags = 0, int n_copy = 0, int n_poll = 0, int grid = 1) {
  const int n = 2400;
  cudaFuncSetAttribute(k<form, n_acc, n256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d, 0, 24);
    k<form, n_acc, n256><<<grid, 640, 144 * 1024>>>(n, n_readers, flags, n_copy, n_poll, g_buf, d, reinterpret_cast<float*>(d + 3));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  long long c2[3]; cudaMemcpy(c2, d, 24, cudaMemcpyDeviceToHost); const long long c = c2[0];
  printf("N=%d %s accumulators=%d: %.1f cycles per MMA (floor %d)\n", n256 ? 256 : 128, form ? "TS" : "SS", n_acc, (double)c / n, n256 ? 128 : 64);
  if (flags || n_copy) printf("    flags: commit-per-tile=%d sw32-5th=%d overwrite-1st=%d, %d copy lanes: %.0f B/clk of bulk copies\n", flags & 1, (flags >> 1) & 1, (flags >> 2) & 1, n_copy, (double)c2[2] * 20480 / (double)c);
  if (grid > 1 || (flags & 24)) printf("    grid %d, %d issuing threads, %s operands\n", grid, (flags & 8) ? 2 : 1, (flags & 16) ? "random" : "constant");
  if (n_poll) printf("    %d warps polling an mbarrier\n", n_poll);
  if (n_readers) printf("    with %d reader warps: TMEM read %.0f B/clk beside the MMAs; %.0f clk per reader iteration (%s%s)\n", n_readers,
                        (double)c2[1] * 32 * 32 * 4 * ((flags & 32) ? 2 : 1) * ((flags & 128) ? 0 : 1) / (double)c / (grid > 1 ? grid : 1),
                        (double)c * n_readers * (grid > 1 ? grid : 1) / (double)c2[1],
                        (flags & 128) ? "ALU only" : (flags & 32) ? "two loads in flight" : "one load", (flags & 64) ? " + 128 FMA before the wait" : "");
}
int main() {
  long long* d; cudaMalloc(&d, 24 + 640 * 4);
  cudaMalloc(&g_buf, 512 * 20480); cudaMemset(g_buf, 0, 512 * 20480);
  run<0, 1, 0>(d); run<0, 2, 0>(d); run<0, 3, 0>(d);
  run<1, 1, 0>(d); run<1, 2, 0>(d); run<1, 3, 0>(d);
  run<0, 1, 1>(d); run<1, 1, 1>(d);
  printf("--- filter-like tiles: 4 MMAs + 1 (norm slice), TS form, 3 accumulators ---\n");
  run<1, 3, 0>(d, 0, 1, 0); run<1, 3, 0>(d, 0, 2, 0); run<1, 3, 0>(d, 0, 4, 0); run<1, 3, 0>(d, 0, 7, 0);
  run<1, 3, 0>(d, 0, 0, 1); run<1, 3, 0>(d, 0, 0, 2); run<1, 3, 0>(d, 0, 0, 4); run<0, 3, 0>(d, 0, 0, 4);
  run<1, 3, 0>(d, 16, 7, 4); run<0, 3, 0>(d, 16, 7, 4);
  run<1, 3, 0>(d, 0, 0, 0, 4); run<1, 3, 0>(d, 0, 0, 0, 16); run<1, 3, 0>(d, 0, 7, 4, 16); run<1, 3, 0>(d, 8, 7, 4, 8);
  run<1, 1, 0>(d, 0, 8); run<1, 1, 0>(d, 0, 15); run<1, 1, 0>(d, 16, 15, 4);
  run<1, 3, 0>(d, 0, 0, 0, 0, 143); run<1, 3, 0>(d, 16, 7, 4, 0, 143); run<1, 1, 0>(d, 16, 15, 4, 0, 143);
  printf("--- does a tcgen05.ld block its warp? reader iteration time: ld+wait | 2 ld+wait | ld+ALU+wait | ALU only ---\n");
  run<1, 3, 0>(d, 16, 0); run<1, 3, 0>(d, 16, 32); run<1, 3, 0>(d, 16, 64); run<1, 3, 0>(d, 16, 128);
  run<1, 3, 0>(d, 4, 0); run<1, 3, 0>(d, 4, 32); run<1, 3, 0>(d, 4, 64); run<1, 3, 0>(d, 4, 128);
  printf("--- ALU-only warps (16) with a yield point every 16 instructions: none | shuffle | syncwarp | nanosleep 0 ---\n");
  run<1, 3, 0>(d, 16, 128); run<1, 3, 0>(d, 16, 128 + 256); run<1, 3, 0>(d, 16, 128 + 512); run<1, 3, 0>(d, 16, 128 + 1024);
  printf("--- same with two issuing threads ---\n");
  run<1, 1, 0>(d, 16, 8 + 128); run<1, 1, 0>(d, 16, 8 + 128 + 256); run<1, 1, 0>(d, 16, 8 + 128 + 1024);
  printf("--- random operands (flag 16) ---\n");
  run<0, 3, 0>(d, 0, 16, 0, 0, 1); run<0, 3, 0>(d, 0, 16, 0, 0, 143); run<0, 1, 0>(d, 16, 16 + 15, 4, 0, 143); run<0, 1, 0>(d, 16, 15, 4, 0, 143);
  for (int r : {16}) { run<1, 1, 0>(d, r); run<1, 3, 0>(d, r); run<0, 3, 0>(d, r); }
  return 0;
}
