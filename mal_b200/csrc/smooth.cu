// smooth.cu - edge-aware smoothness term, forward + backward, for sm_100a.
//
// Replaces get_smooth_loss (manydepth/layers.py:210-223) and the mean-normalisation the trainers
// put in front of it (manydepth/loss_utils.py:119-121, manydepth/trainer.py:1440-1442):
//     dn    = disp / (mean_hw(disp) + 1e-7)
//     loss  = mean_x( |dn[x] - dn[x+1]| * exp(-mean_c |I[x] - I[x+1]|) ) + the same along y
//
// Kernels (all deterministic, no atomics):
//   1  smooth_mean_kernel      per-sample 1 / (mean(disp) + 1e-7)               (normalise only)
//   2  smooth_main_kernel      one thread per pixel: it owns its right and down edges (loss) and
//                              re-derives its left and up edges for the gradient, so the backward
//                              needs no scatter, no shared tile and no barrier; per-CTA loss partials
//   3  smooth_finalize_kernel  fixed-order reduction -> loss, per-sample loss share
//   4  smooth_fix_kernel       chain through the normalisation:  the loss is homogeneous of degree 1
//                              in dn, so sum_j g'_j dn_j = L_b and
//                              d loss / d disp_i = (g'_i - L_b / HW) / (mean_b + 1e-7)
// HBM traffic: disp 4 + img 12 B/px read, grad 4 B/px written (+ 8 B/px for the fix pass).
#include "mal_math.cuh"

namespace mal {

constexpr int SM_TW = 32, SM_TH = 8, SM_NT = 256;
constexpr int SM_SPLIT = 8;    // CTAs per sample in the mean pass

struct SmoothWs {   // offsets into the float workspace
  size_t mean_part, partials, lb, total;
};
__host__ __device__ inline SmoothWs smooth_ws(int batch, int tiles) {
  SmoothWs w;
  w.mean_part = 0;
  w.partials = w.mean_part + (size_t)batch * SM_SPLIT;
  w.lb = w.partials + (size_t)batch * tiles;
  w.total = w.lb + (size_t)batch * 2;   // [L_b, 1 / (mean_b + eps)]
  return w;
}

// partial sums of disp: SM_SPLIT CTAs per sample, 4 independent loads in flight per thread
__global__ void __launch_bounds__(1024) smooth_mean_kernel(const float* __restrict__ disp, int hw,
                                                          float* __restrict__ part) {
  __shared__ float red[32];
  const int b = blockIdx.y, sp = blockIdx.x;
  const int per = (hw + SM_SPLIT - 1) / SM_SPLIT;
  const int i1 = min(hw, (sp + 1) * per);
  const float* d = disp + (size_t)b * hw;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = sp * per + threadIdx.x;
  for (; i + 3 * 1024 < i1; i += 4 * 1024) {
    a0 += __ldg(d + i); a1 += __ldg(d + i + 1024); a2 += __ldg(d + i + 2048); a3 += __ldg(d + i + 3072);
  }
  for (; i < i1; i += 1024) a0 += __ldg(d + i);
  float acc = warp_sum((a0 + a1) + (a2 + a3));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int wv = 0; wv < 32; wv++) t += red[wv];
    part[b * SM_SPLIT + sp] = t;
  }
}

// 1 / (mean(disp) + 1e-7) from the partial sums, in a fixed order
__device__ __forceinline__ float smooth_scale(const float* __restrict__ part, int b, int hw) {
  float t = 0.0f;
#pragma unroll
  for (int k = 0; k < SM_SPLIT; k++) t += __ldg(part + b * SM_SPLIT + k);
  return 1.0f / (t / (float)hw + 1e-7f);
}

struct SmPix { float d, r, g, b; };   // scaled disparity and colour of one pixel

__device__ __forceinline__ SmPix sm_load(const float* __restrict__ d, const float* __restrict__ im, int hw, int o,
                                         float sc) {
  SmPix p;
  p.d = __ldg(d + o) * sc; p.r = __ldg(im + o); p.g = __ldg(im + hw + o); p.b = __ldg(im + 2 * hw + o);
  return p;
}

// signed, weighted edge term between pixels p and q (q = right or lower neighbour of p):
// w = exp(-mean_c |dI|) * inv_n carries the sign of (dn[p] - dn[q]); *absval gets |ddn| * w.
__device__ __forceinline__ float edge_term(const SmPix& p, const SmPix& q, float inv_n, float* absval) {
  const float g = p.d - q.d;
  const float wgt = expf(-(fabsf(p.r - q.r) + fabsf(p.g - q.g) + fabsf(p.b - q.b)) * (1.0f / 3.0f)) * inv_n;
  *absval = fabsf(g) * wgt;
  return g > 0.f ? wgt : (g < 0.f ? -wgt : 0.f);
}

// One thread per pixel, no shared tile: the pixel owns its right and down edges (loss) and
// re-derives its left and up edges for the gradient; every neighbour is loaded once (L1).
__global__ void __launch_bounds__(SM_NT) smooth_main_kernel(const mal_smooth_args a, const float inv_nx,
                                                           const float inv_ny, const int tiles_x,
                                                           const int tiles) {
  __shared__ float red[SM_NT / 32];
  const int H = a.height, W = a.width, hw = H * W;
  const int b = blockIdx.z;
  const int x = blockIdx.x * SM_TW + (threadIdx.x % SM_TW), y = blockIdx.y * SM_TH + threadIdx.x / SM_TW;
  const SmoothWs ws = smooth_ws(a.batch, tiles);
  const float sc = a.normalise ? smooth_scale(a.workspace + ws.mean_part, b, hw) : 1.0f;
  const float* d = a.disp + (size_t)b * hw;
  const float* im = a.img + (size_t)b * 3 * hw;
  float acc = 0.0f;
  if (x < W && y < H) {
    const int o = y * W + x;
    const SmPix c = sm_load(d, im, hw, o, sc);
    float g = 0.0f, av;
    if (x + 1 < W) { g += edge_term(c, sm_load(d, im, hw, o + 1, sc), inv_nx, &av); acc += av; }
    if (y + 1 < H) { g += edge_term(c, sm_load(d, im, hw, o + W, sc), inv_ny, &av); acc += av; }
    if (a.with_grad) {
      if (x > 0) g -= edge_term(sm_load(d, im, hw, o - 1, sc), c, inv_nx, &av);
      if (y > 0) g -= edge_term(sm_load(d, im, hw, o - W, sc), c, inv_ny, &av);
      a.grad_disp[(size_t)b * hw + o] = g;   // d loss / d dn; the fix pass chains through the normalisation
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int wv = 0; wv < SM_NT / 32; wv++) t += red[wv];
    a.workspace[ws.partials + (size_t)b * tiles + blockIdx.y * tiles_x + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) smooth_finalize_kernel(const mal_smooth_args a, const int tiles) {
  double* per_sample = reinterpret_cast<double*>(dyn_smem());   // [batch]
  const SmoothWs ws = smooth_ws(a.batch, tiles);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int b = warp; b < a.batch; b += nwarp) {
    double s = 0.0;
    for (int t = lane; t < tiles; t += 32) s += (double)a.workspace[ws.partials + (size_t)b * tiles + t];
    s = warp_sum(s);
    if (lane == 0) {
      per_sample[b] = s;
      a.workspace[ws.lb + b * 2] = (float)s;
      if (a.normalise) a.workspace[ws.lb + b * 2 + 1] = smooth_scale(a.workspace + ws.mean_part, b, a.height * a.width);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < a.batch; b++) s += per_sample[b];
    a.loss[0] = (float)s;
  }
}

__global__ void __launch_bounds__(SM_NT) smooth_fix_kernel(const mal_smooth_args a, const int tiles) {
  const SmoothWs ws = smooth_ws(a.batch, tiles);
  const size_t hw = (size_t)a.height * a.width, total = (size_t)a.batch * hw;
  for (size_t i = (size_t)blockIdx.x * SM_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * SM_NT) {
    const int b = (int)(i / hw);
    const float lb = a.workspace[ws.lb + b * 2], sc = a.workspace[ws.lb + b * 2 + 1];   // sc = 1 / (mean + 1e-7)
    a.grad_disp[i] = (a.grad_disp[i] - lb / (float)hw) * sc;
  }
}

inline int smooth_tiles(int height, int width, int* tx) {
  int x = (width + SM_TW - 1) / SM_TW, y = (height + SM_TH - 1) / SM_TH;
  if (tx) *tx = x;
  return x * y;
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_smooth_workspace_floats(int batch, int height, int width) {
  return smooth_ws(batch, smooth_tiles(height, width, nullptr)).total;
}

extern "C" int mal_smooth_forward(const mal_smooth_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_smooth_forward: args is NULL");
  const mal_smooth_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height >= 2 && a.width >= 2, "mal_smooth_forward: bad shape %dx%dx%d", a.batch,
              a.height, a.width);
  MAL_REQUIRE(a.batch <= 65535, "mal_smooth_forward: batch %d exceeds gridDim.z", a.batch);
  MAL_REQUIRE(a.disp && a.img && a.workspace && a.loss, "mal_smooth_forward: disp/img/workspace/loss are required");
  if (a.with_grad) MAL_REQUIRE(a.grad_disp, "mal_smooth_forward: with_grad needs grad_disp");
  cudaStream_t st = (cudaStream_t)stream;
  int tiles_x;
  const int tiles = smooth_tiles(a.height, a.width, &tiles_x);
  const int hw = a.height * a.width;
  if (a.normalise) {
    launch(smooth_mean_kernel, dim3(SM_SPLIT, a.batch), dim3(1024), 0, st, a.disp, hw,
           a.workspace + smooth_ws(a.batch, tiles).mean_part);
    int rc = check_launch("smooth_mean_kernel");
    if (rc) return rc;
  }
  // gdx.mean() over (B,1,H,W-1) and gdy.mean() over (B,1,H-1,W)
  const float inv_nx = (float)(1.0 / ((double)a.batch * a.height * (a.width - 1)));
  const float inv_ny = (float)(1.0 / ((double)a.batch * (a.height - 1) * a.width));
  dim3 grid(tiles_x, tiles / tiles_x, a.batch);
  launch(smooth_main_kernel, grid, dim3(SM_NT), 0, st, a, inv_nx, inv_ny, tiles_x, tiles);
  int rc = check_launch("smooth_main_kernel");
  if (rc) return rc;
  launch(smooth_finalize_kernel, dim3(1), dim3(1024), (size_t)a.batch * 8 + 16, st, a, tiles);
  rc = check_launch("smooth_finalize_kernel");
  if (rc) return rc;
  if (a.with_grad && a.normalise) {
    size_t total = (size_t)a.batch * hw;
    size_t blk = (total + SM_NT - 1) / SM_NT;
    if (blk > 148 * 8) blk = 148 * 8;
    launch(smooth_fix_kernel, dim3((unsigned)blk), dim3(SM_NT), 0, st, a, tiles);
    rc = check_launch("smooth_fix_kernel");
  }
  return rc;
}
