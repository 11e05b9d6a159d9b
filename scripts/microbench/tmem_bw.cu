// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}

template <int WAIT_EVERY>
__global__ void __launch_bounds__(512, 1) bw_kernel(int iters, int active_warps, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < active_warps) {
    uint32_t r[32];
    for (int i = 0; i < iters; ++i) {
      ld32(base + ((i * 32) & 511 & ~31), r);
      if ((i % WAIT_EVERY) == WAIT_EVERY - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 8) acc ^= r[j];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  long long* d_cycles; uint32_t* d_sink;
  cudaMalloc(&d_cycles, 148 * sizeof(long long)); cudaMalloc(&d_sink, 4);
  const int iters = 4096;
  for (int grid : {1, 148}) {
    for (int warps : {4, 8, 16}) {
      for (int mode = 0; mode < 2; ++mode) {
        if (mode == 0) bw_kernel<1><<<grid, 512>>>(iters, warps, d_cycles, d_sink);
        else bw_kernel<4><<<grid, 512>>>(iters, warps, d_cycles, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long c[148]; cudaMemcpy(c, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
        double bytes = (double)iters * warps * 32 * 32 * 4;
        printf("grid=%3d warps=%2d wait_every=%d: %lld cycles, %.1f B/clk/SM, %.1f cyc per LDTM.x32 per warp\n", grid, warps,
               mode == 0 ? 1 : 4, mx, bytes / mx, (double)mx / iters);
      }
    }
  }
  return 0;
}
