// Device-initiated exchange steps of the memory-axis sharded read over NVLink peer memory.
//
// One process per GPU; every rank maps the exchange buffers of all ranks (CUDA IPC + peer access), and the kernels
// store to / load from those mappings directly:
//   finalize_kernel (select_simt.cu)  pushes each query's local top-k list into every rank's gather region,
//   peer_barrier_kernel               orders the ranks (release / acquire at system scope on per-rank flag words),
//   peer_reduce_scatter_kernel        sums the ranks' partial readouts for the query slice this rank owns, reading
//                                     the partials straight out of the peers' memory, and writes the slice in the
//                                     reference layout (K, CV, q).
// No host-side collective is on the data path; NCCL (or gloo) only carries the IPC handles once.
#include "common.cuh"

namespace evavos {

namespace {

struct PeerPtrs {
  int n_ranks;
  int rank;
  uint8_t* base[EVAVOS_MAX_RANKS];
};

// <<<1, 32>>>: lane g < n_ranks talks to rank g.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerPtrs p, int64_t flag_offset, uint32_t epoch) {
  const int g = threadIdx.x;
  // everything this stream did before (stores to peer buffers by earlier kernels included) is ordered before the
  // flag stores below at system scope
  __threadfence_system();
  uint32_t* mine = reinterpret_cast<uint32_t*>(p.base[p.rank] + flag_offset);
  if (g < p.n_ranks && g != p.rank) {
    uint32_t* theirs = reinterpret_cast<uint32_t*>(p.base[g] + flag_offset) + p.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const long long t0 = clock64();
    uint32_t seen;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine + g) : "memory");
      if ((int32_t)(seen - epoch) >= 0) break;
      if (clock64() - t0 > 4000000000ll) {   // ~2 s: a rank died; do not hang the GPU - poison the flag word
        mine[31] = 0xdeadbeefu;
        break;
      }
      __nanosleep(64);
    } while (true);
  }
  __syncwarp();
  __threadfence_system();
}

// out[row][q - q0] = sum_g partial_g[q][row].  One CTA per (32 queries, 128 rows): float4 loads along the rows of
// every rank's query-major partial (coalesced 512-byte pieces, peers over NVLink), transposed through shared memory.
__global__ void __launch_bounds__(256) peer_reduce_scatter_kernel(const PeerPtrs p, int64_t partial_offset, int rows,
                                                                  int64_t q0, int64_t q1, float* __restrict__ out,
                                                                  int64_t out_row_stride) {
  __shared__ float tile[32][129];
  const int64_t qb = q0 + (int64_t)blockIdx.x * 32;
  const int rb = blockIdx.y * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp w sums queries qb + w, qb + w + 8, ... ; lane covers rows rb + 4 * lane .. + 3
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t q = qb + warp + 8 * i;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < q1 && rb + 4 * lane < rows) {
      for (int g = 0; g < p.n_ranks; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(
            reinterpret_cast<const float*>(p.base[g] + partial_offset) + q * rows + rb + 4 * lane);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    tile[warp + 8 * i][4 * lane + 0] = acc.x;
    tile[warp + 8 * i][4 * lane + 1] = acc.y;
    tile[warp + 8 * i][4 * lane + 2] = acc.z;
    tile[warp + 8 * i][4 * lane + 3] = acc.w;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 128 * 32; e += 256) {
    const int r = e >> 5, qi = e & 31;
    if (rb + r < rows && qb + qi < q1) out[(int64_t)(rb + r) * out_row_stride + (qb - q0) + qi] = tile[qi][r];
  }
}

PeerPtrs to_ptrs(const EvavosPeers& peers) {
  PeerPtrs p;
  p.n_ranks = peers.n_ranks;
  p.rank = peers.rank;
  for (int g = 0; g < EVAVOS_MAX_RANKS; ++g)
    p.base[g] = g < peers.n_ranks ? reinterpret_cast<uint8_t*>(peers.base[g]) : nullptr;
  return p;
}

}  // namespace

int launch_peer_barrier(const EvavosPeers& peers, int64_t flag_offset, uint32_t epoch, cudaStream_t st) {
  peer_barrier_kernel<<<1, 32, 0, st>>>(to_ptrs(peers), flag_offset, epoch);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

int launch_peer_reduce_scatter(const EvavosPeers& peers, int64_t partial_offset, int rows, int64_t q0, int64_t q1,
                               float* out, int64_t out_row_stride, cudaStream_t st) {
  if (q1 <= q0) return EVAVOS_OK;
  const dim3 grid((unsigned)ceil_div(q1 - q0, 32), (unsigned)ceil_div(rows, 128));
  peer_reduce_scatter_kernel<<<grid, 256, 0, st>>>(to_ptrs(peers), partial_offset, rows, q0, q1, out, out_row_stride);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

}  // namespace evavos
