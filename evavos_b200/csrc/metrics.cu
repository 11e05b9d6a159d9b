// Per-frame J and J&F quality of all frames of a video on the GPU.
//
// The reference scores every annotation round on the CPU, frame by frame: un-pad, argmax, D2H, then numpy + cv2 per
// frame (interactions/eval.py:27-81 -> interactions/metrics.py:9-36, 40-160: IoU, one-pixel boundary maps, two
// cv2.dilate calls with a disk of ceil(0.008 * |shape|) pixels, precision / recall).  Here the predicted and the
// ground-truth masks of ALL frames stay on the device and three launches produce the per-frame numbers:
//   jf_boundary_kernel  boundary maps of both masks (metrics.py:40-97, full resolution) + the IoU counts,
//   jf_match_kernel     for every boundary pixel: is there a boundary pixel of the other mask within the disk?
//                       (= boundary * dilate(other boundary, disk), metrics.py:127-136; out-of-image pixels do not
//                       count, like cv2.dilate's default border) - exact, tile + halo in shared memory,
//   jf_finalize_kernel  precision, recall, F, Jaccard, J&F per frame in the reference's arithmetic
//                       (fp32 for the IoU / Jaccard quotients, fp64 for F and the average).
// HBM-bound: 2 * T*h*w bytes read + 2 * T*h*w written + read again (plus halos).
#include "common.cuh"

namespace evavos {

namespace {

enum { kInter = 0, kUnion, kNFg, kNGt, kFgMatch, kGtMatch, kGtPixels, kCounters = 8 };

__device__ __forceinline__ int block_sum(int v, int* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  int t = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;   // valid on thread 0
}

// one-pixel boundary of a binary segmentation, offset by half a pixel towards the origin (metrics.py:76-86)
__device__ __forceinline__ bool boundary_at(const uint8_t* __restrict__ seg, int y, int x, int h, int w) {
  const bool c = seg[(int64_t)y * w + x] != 0;
  const bool e = (x + 1 < w) && seg[(int64_t)y * w + x + 1] != 0;
  const bool s = (y + 1 < h) && seg[(int64_t)(y + 1) * w + x] != 0;
  const bool se = (x + 1 < w) && (y + 1 < h) && seg[(int64_t)(y + 1) * w + x + 1] != 0;
  if (y == h - 1 && x == w - 1) return false;
  if (x == w - 1) return c != s;     // b[:, -1] = seg[:, -1] ^ s[:, -1] (assigned after the last row)
  if (y == h - 1) return c != e;     // b[-1, :] = seg[-1, :] ^ e[-1, :]
  return (c != e) || (c != s) || (c != se);
}

// grid (pixel blocks of one frame, T), 256 threads
__global__ void __launch_bounds__(256) jf_boundary_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                                                          int h, int w, uint8_t* __restrict__ pred_b,
                                                          uint8_t* __restrict__ gt_b, int32_t* __restrict__ counters) {
  __shared__ int sh[8];
  const int64_t frame = blockIdx.y, px = (int64_t)h * w;
  const uint8_t* p = pred + frame * px;
  const uint8_t* g = gt + frame * px;
  int inter = 0, uni = 0, nfg = 0, ngt = 0, gpx = 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < px; i += (int64_t)gridDim.x * 256) {
    const int y = (int)(i / w), x = (int)(i % w);
    const bool pv = p[i] != 0, gv = g[i] != 0;
    inter += pv && gv;
    uni += pv || gv;
    gpx += gv;
    const bool pb = boundary_at(p, y, x, h, w), gb = boundary_at(g, y, x, h, w);
    pred_b[frame * px + i] = pb;
    gt_b[frame * px + i] = gb;
    nfg += pb;
    ngt += gb;
  }
  int32_t* c = counters + frame * kCounters;
  int t;
  t = block_sum(inter, sh); if (threadIdx.x == 0 && t) atomicAdd(c + kInter, t);
  t = block_sum(uni, sh);   if (threadIdx.x == 0 && t) atomicAdd(c + kUnion, t);
  t = block_sum(nfg, sh);   if (threadIdx.x == 0 && t) atomicAdd(c + kNFg, t);
  t = block_sum(ngt, sh);   if (threadIdx.x == 0 && t) atomicAdd(c + kNGt, t);
  t = block_sum(gpx, sh);   if (threadIdx.x == 0 && t) atomicAdd(c + kGtPixels, t);
}

constexpr int kTile = 32;
constexpr int kMaxRadius = 24;   // 0.008 * |(2160, 3840)| = 35.2 would need more; 1080p gives 18

// grid (tiles_x, tiles_y, T), 256 threads = 32 x 8, four rows per thread
__global__ void __launch_bounds__(256) jf_match_kernel(const uint8_t* __restrict__ pred_b, const uint8_t* __restrict__ gt_b,
                                                       int h, int w, int radius, int32_t* __restrict__ counters) {
  extern __shared__ uint8_t smem[];
  const int span = kTile + 2 * radius;
  uint8_t* sp = smem;                 // predicted boundary, tile + halo
  uint8_t* sg = smem + span * span;   // ground-truth boundary
  __shared__ int sh[8];
  __shared__ int dxmax[2 * kMaxRadius + 1];
  const int64_t frame = blockIdx.z, px = (int64_t)h * w;
  const int x0 = blockIdx.x * kTile - radius, y0 = blockIdx.y * kTile - radius;
  for (int e = threadIdx.x; e < span * span; e += 256) {
    const int yy = y0 + e / span, xx = x0 + e % span;
    const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
    sp[e] = in ? pred_b[frame * px + (int64_t)yy * w + xx] : 0;
    sg[e] = in ? gt_b[frame * px + (int64_t)yy * w + xx] : 0;
  }
  if (threadIdx.x <= 2 * radius) {
    const int dy = (int)threadIdx.x - radius;
    int d = 0;
    while ((d + 1) * (d + 1) + dy * dy <= radius * radius) ++d;   // disk(r): dx^2 + dy^2 <= r^2
    dxmax[threadIdx.x] = d;
  }
  __syncthreads();
  int fg_match = 0, gt_match = 0;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = ty + 8 * k;
    const int cy = ly + radius, cx = tx + radius;
    const bool pb = sp[cy * span + cx] != 0, gb = sg[cy * span + cx] != 0;
    if (pb || gb) {     // boundary pixels are a few percent of a frame: most threads skip the search
      bool near_g = false, near_p = false;
      for (int dy = -radius; dy <= radius; ++dy) {
        const int d = dxmax[dy + radius];
        const uint8_t* rp = sp + (cy + dy) * span + cx;
        const uint8_t* rg = sg + (cy + dy) * span + cx;
        for (int dx = -d; dx <= d; ++dx) {
          near_p |= rp[dx] != 0;
          near_g |= rg[dx] != 0;
        }
        if ((near_g || !pb) && (near_p || !gb)) break;
      }
      fg_match += pb && near_g;   // fg_boundary * dilate(gt_boundary)
      gt_match += gb && near_p;   // gt_boundary * dilate(fg_boundary)
    }
  }
  int32_t* c = counters + frame * kCounters;
  int t;
  t = block_sum(fg_match, sh); if (threadIdx.x == 0 && t) atomicAdd(c + kFgMatch, t);
  t = block_sum(gt_match, sh); if (threadIdx.x == 0 && t) atomicAdd(c + kGtMatch, t);
}

// out (T, 4) fp64: smoothed IoU (metrics.py:9-20), binary Jaccard, F (metrics.py:138-158), 0.5 J + 0.5 F (:36)
__global__ void jf_finalize_kernel(const int32_t* __restrict__ counters, int64_t T, double* __restrict__ out,
                                   int32_t* __restrict__ gt_empty) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= T) return;
  const int32_t* c = counters + f * kCounters;
  const float inter = (float)c[kInter], uni = (float)c[kUnion];
  const float iou = (inter + 1e-6f) / (uni + 1e-6f);
  const float jac = c[kUnion] > 0 ? inter / uni : 0.f;
  const int n_fg = c[kNFg], n_gt = c[kNGt];
  double precision, recall;
  if (n_fg == 0 && n_gt > 0) { precision = 1; recall = 0; }
  else if (n_fg > 0 && n_gt == 0) { precision = 0; recall = 1; }
  else if (n_fg == 0 && n_gt == 0) { precision = 1; recall = 1; }
  else { precision = (double)c[kFgMatch] / (double)n_fg; recall = (double)c[kGtMatch] / (double)n_gt; }
  const double F = (precision + recall == 0) ? 0.0 : 2 * precision * recall / (precision + recall);
  out[4 * f + 0] = (double)iou;
  out[4 * f + 1] = (double)jac;
  out[4 * f + 2] = F;
  out[4 * f + 3] = (double)jac * 0.5 + F * 0.5;
  if (gt_empty) gt_empty[f] = c[kGtPixels] == 0;
}

}  // namespace

size_t jf_workspace_bytes(int64_t T, int h, int w) {
  return (size_t)2 * T * h * w + 256 + sizeof(int32_t) * kCounters * (size_t)T;
}

int launch_jf_metrics(const uint8_t* pred, const uint8_t* gt, int64_t T, int h, int w, int radius, void* workspace,
                      double* out, int32_t* gt_empty, cudaStream_t st) {
  if (T <= 0) return EVAVOS_OK;
  if (radius < 0 || radius > kMaxRadius) {
    set_error("jf_metrics: boundary radius %d unsupported (0..%d)", radius, kMaxRadius);
    return EVAVOS_ERR_UNSUPPORTED;
  }
  const int64_t px = (int64_t)h * w;
  uint8_t* pred_b = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* gt_b = pred_b + T * px;
  int32_t* counters = reinterpret_cast<int32_t*>((reinterpret_cast<uintptr_t>(gt_b + T * px) + 255) & ~(uintptr_t)255);
  EVAVOS_CUDA_OK(cudaMemsetAsync(counters, 0, sizeof(int32_t) * kCounters * (size_t)T, st));
  int64_t bx = ceil_div(px, 256 * 4);
  if (bx > 1024) bx = 1024;
  jf_boundary_kernel<<<dim3((unsigned)bx, (unsigned)T), 256, 0, st>>>(pred, gt, h, w, pred_b, gt_b, counters);
  const int span = kTile + 2 * radius;
  jf_match_kernel<<<dim3((unsigned)ceil_div(w, kTile), (unsigned)ceil_div(h, kTile), (unsigned)T), 256,
                    (size_t)2 * span * span, st>>>(pred_b, gt_b, h, w, radius, counters);
  jf_finalize_kernel<<<(unsigned)ceil_div(T, 128), 128, 0, st>>>(counters, T, out, gt_empty);
  EVAVOS_CUDA_OK(cudaGetLastError());
  return EVAVOS_OK;
}

}  // namespace evavos
