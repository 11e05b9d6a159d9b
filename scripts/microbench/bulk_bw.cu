// Microbenchmark: cp.async.bulk (UBLKCP) global->shared throughput per SM with a 6-stage ring of 20 KB tiles.
// mode 0: every CTA streams its own distinct tiles (L2-resident buffer)
// mode 1: groups of 13 CTAs stream the SAME tiles in lock-step (what the score filter did)
// mode 2: groups of 13 CTAs stream the same tiles, rotated starts
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kTile = 20480;
#ifndef KSTAGES
#define KSTAGES 6
#endif
#ifndef KPROD
#define KPROD 1
#endif
#ifndef KSPLIT
#define KSPLIT 1
#endif
constexpr int kStages = KSTAGES;
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__global__ void __launch_bounds__(128, 1) k(const uint8_t* buf, int n_buf_tiles, int tiles_per_cta, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
#ifdef KLANES
  if (threadIdx.x < KPROD) {
    const int pw = threadIdx.x;
#else
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < KPROD) {
    const int pw = threadIdx.x >> 5;
#endif
    const int grp = blockIdx.x / 13, mem = blockIdx.x % 13;
    auto tile_index = [&](int i) -> int {
      if (mode == 0) return (blockIdx.x * tiles_per_cta + i) % n_buf_tiles;
      int t = i;
      if (mode == 2) t = (i + mem * tiles_per_cta / 13) % tiles_per_cta;
      return (grp * tiles_per_cta + t) % n_buf_tiles;
    };
    for (int i = pw; i < tiles_per_cta + kStages; i += KPROD) {
      if (i >= kStages) {  // consume tile i - kStages (just wait for it)
        const int j = i - kStages, s = j % kStages;
        while (!try_wait(s32(&full[s]), (j / kStages) & 1)) {}
      }
      if (i < tiles_per_cta) {
        const int s = i % kStages;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(kTile) : "memory");
        for (int part = 0; part < KSPLIT; ++part)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + s * kTile + part * (kTile / KSPLIT))),
                       "l"(buf + (size_t)tile_index(i) * kTile + part * (kTile / KSPLIT)), "r"(kTile / KSPLIT), "r"(s32(&full[s])) : "memory");
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
  const int n_buf_tiles = 2048;  // 42 MB, L2 resident
  uint8_t* buf; cudaMalloc(&buf, (size_t)n_buf_tiles * kTile); cudaMemset(buf, 1, (size_t)n_buf_tiles * kTile);
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kTile);
  for (int grid : {1, 143}) for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
    const int tiles = 230;
    k<<<grid, 128, kStages * kTile>>>(buf, n_buf_tiles, tiles, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long c[148]; cudaMemcpy(c, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
    if (rep == 1) printf("prod=%d stages=%d split=%d grid=%3d mode=%d: %lld cycles, %.1f cycles/tile, %.1f B/clk/SM, %.2f KB/clk chip\n", KPROD, KSTAGES, KSPLIT, grid, mode, mx, (double)mx / tiles,
           (double)tiles * kTile / mx, (double)tiles * kTile * grid / mx / 1024);
  }
  return 0;
}
