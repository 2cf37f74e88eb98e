// smooth.cu - edge-aware smoothness term, forward + backward, for sm_100a.
//
// Replaces get_smooth_loss (manydepth/layers.py:210-223) and the mean-normalisation the trainers
// put in front of it (manydepth/loss_utils.py:119-121, manydepth/trainer.py:1440-1442):
//     dn    = disp / (mean_hw(disp) + 1e-7)
//     loss  = mean_x( |dn[x] - dn[x+1]| * exp(-mean_c |I[x] - I[x+1]|) ) + the same along y
//
// ONE launch scores one or two disparities (teacher and student share the image, so the edge weights -
// the exp() and three of the five planes read per pixel - are computed once for both):
//   smooth_kernel<VEC, DUAL>  a thread owns VEC (4 when rows are 16-byte multiples) horizontally adjacent
//       pixels: 128-bit loads of its row and of the rows above / below, its right and down edges for the
//       loss, its left and up edges re-derived for the gradient - no scatter, no shared tile, no barrier in
//       the main part.  The loss is homogeneous of degree 1 in dn, so the kernel works on the UN-normalised
//       disparity and only accumulates sum(disp) beside the raw loss: no mean pass has to run first.
//       The last CTA of a sample (ticket) reduces the per-CTA partials in a fixed order:
//           s_b = 1 / (mean_b + 1e-7),  L_b = s_b * raw_b,   loss = sum_b L_b (last sample, sample order)
//   smooth_fix_kernel         chains the gradient through the normalisation,
//           d loss / d disp_i = (g'_i - L_b / HW) * s_b          (sum_j g'_j dn_j = L_b by homogeneity)
//       unless the caller asked to defer it (`defer_fix`): mal_step_combine applies it while it adds the
//       gradient planes up anyway.
// HBM traffic: disp 4 (8) + img 12 B/px read, grad 4 (8) B/px written (+ 8 B/px per term for the fix pass).
#include "mal_math.cuh"

namespace mal {

constexpr int SM_TH = 8, SM_NT = 256;   // a CTA covers 8 rows x 32 threads x VEC pixels

struct SmoothWs {   // offsets into the float workspace
  size_t partials, stats, tickets, total;
};
__host__ __device__ inline SmoothWs smooth_ws(int batch, int tiles) {
  SmoothWs w;
  w.partials = 0;                                        // [batch][tiles][4]: sum(disp), raw loss  x 2 terms
  w.stats = w.partials + (size_t)batch * tiles * 4;      // [batch][2 terms][2]: L_b, 1 / (mean_b + eps)
  w.tickets = w.stats + (size_t)batch * 4;               // [batch] + [1] unsigned
  w.total = w.tickets + (size_t)batch + 1 + 3;
  return w;
}

template <int VEC>
struct SmVec { float v[VEC]; };
template <int VEC>
__device__ __forceinline__ SmVec<VEC> sm_ld(const float* __restrict__ p) {
  SmVec<VEC> r;
  if (VEC == 4) {
#ifdef MAL_EMU
    const float4 t = *reinterpret_cast<const float4*>(p);
#else
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
#endif
    r.v[0] = t.x; r.v[1 % VEC] = t.y; r.v[2 % VEC] = t.z; r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}
__device__ __forceinline__ float sm_weight(float r0, float g0, float b0, float r1, float g1, float b1, float inv_n) {
  return expf(-(fabsf(r0 - r1) + fabsf(g0 - g1) + fabsf(b0 - b1)) * (1.0f / 3.0f)) * inv_n;
}
__device__ __forceinline__ float sm_sign(float g, float w) { return g > 0.f ? w : (g < 0.f ? -w : 0.f); }

template <int VEC, bool DUAL>
__global__ void __launch_bounds__(SM_NT) smooth_kernel(const mal_smooth_args a, const float inv_nx, const float inv_ny,
                                                      const int tiles_x, const int tiles) {
  constexpr int ND = DUAL ? 2 : 1;
  __shared__ float red[SM_NT / 32][4];
  __shared__ double tot[4];
  __shared__ int s_last;
  const int H = a.height, W = a.width, hw = H * W;
  const int b = blockIdx.z;
  const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * VEC, y = blockIdx.y * SM_TH + (threadIdx.x >> 5);
  const SmoothWs ws = smooth_ws(a.batch, tiles);
  const float* im = a.img + (size_t)b * 3 * hw;
  const float* dp[2] = {a.disp + (size_t)b * hw, DUAL ? a.disp_b + (size_t)b * hw : nullptr};
  float* gp[2] = {a.grad_disp ? a.grad_disp + (size_t)b * hw : nullptr,
                  (DUAL && a.grad_disp_b) ? a.grad_disp_b + (size_t)b * hw : nullptr};
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // sum(disp), raw loss per term
  if (x < W && y < H) {
    const int o = y * W + x;
    const bool has_r = x + VEC < W, has_l = x > 0, has_d = y + 1 < H, has_u = y > 0;
    const bool grad = a.with_grad != 0;
    // edge weights: right edge of every own pixel, the left edge of the first, down and up edges
    float wr[VEC], wd[VEC], wu[VEC], wl = 0.0f;
    {
      SmVec<VEC> c[3], dn[3], up[3];
      float rn[3] = {0.f, 0.f, 0.f}, ln[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        c[ch] = sm_ld<VEC>(im + (size_t)ch * hw + o);
        if (has_d) dn[ch] = sm_ld<VEC>(im + (size_t)ch * hw + o + W);
        if (grad && has_u) up[ch] = sm_ld<VEC>(im + (size_t)ch * hw + o - W);
        if (has_r) rn[ch] = __ldg(im + (size_t)ch * hw + o + VEC);
        if (grad && has_l) ln[ch] = __ldg(im + (size_t)ch * hw + o - 1);
      }
#pragma unroll
      for (int v = 0; v < VEC; v++) {
        const bool last = v == VEC - 1;
        const float r1 = last ? rn[0] : c[0].v[(v + 1) % VEC], g1 = last ? rn[1] : c[1].v[(v + 1) % VEC],
                    b1 = last ? rn[2] : c[2].v[(v + 1) % VEC];
        wr[v] = (!last || has_r) ? sm_weight(c[0].v[v], c[1].v[v], c[2].v[v], r1, g1, b1, inv_nx) : 0.0f;
        wd[v] = has_d ? sm_weight(c[0].v[v], c[1].v[v], c[2].v[v], dn[0].v[v], dn[1].v[v], dn[2].v[v], inv_ny) : 0.0f;
        wu[v] = (grad && has_u) ? sm_weight(up[0].v[v], up[1].v[v], up[2].v[v], c[0].v[v], c[1].v[v], c[2].v[v], inv_ny) : 0.0f;
      }
      if (grad && has_l) wl = sm_weight(ln[0], ln[1], ln[2], c[0].v[0], c[1].v[0], c[2].v[0], inv_nx);
    }
#pragma unroll
    for (int t = 0; t < ND; t++) {
      const float* d = dp[t];
      const SmVec<VEC> c = sm_ld<VEC>(d + o);
      SmVec<VEC> dn, up;
      if (has_d) dn = sm_ld<VEC>(d + o + W);
      if (grad && has_u) up = sm_ld<VEC>(d + o - W);
      const float rn = has_r ? __ldg(d + o + VEC) : 0.0f, ln = (grad && has_l) ? __ldg(d + o - 1) : 0.0f;
      float g[VEC];
#pragma unroll
      for (int v = 0; v < VEC; v++) g[v] = 0.0f;
      float sum_d = 0.0f, loss = 0.0f;
#pragma unroll
      for (int v = 0; v < VEC; v++) {
        sum_d += c.v[v];
        const bool last = v == VEC - 1;
        if (!last || has_r) {
          const float e = c.v[v] - (last ? rn : c.v[(v + 1) % VEC]);
          loss += fabsf(e) * wr[v];
          const float sg = sm_sign(e, wr[v]);
          g[v] += sg;
          if (!last) g[(v + 1) % VEC] -= sg;   // the same edge is the next pixel's left edge
        }
        if (has_d) {
          const float e = c.v[v] - dn.v[v];
          loss += fabsf(e) * wd[v];
          g[v] += sm_sign(e, wd[v]);
        }
        if (grad && has_u) g[v] -= sm_sign(up.v[v] - c.v[v], wu[v]);
      }
      if (grad && has_l) g[0] -= sm_sign(ln - c.v[0], wl);
      acc[t * 2] = sum_d;
      acc[t * 2 + 1] = loss;
      if (grad && gp[t]) {   // d loss / d dn up to the per-sample scale; chained through the normalisation later
        if (VEC == 4) *reinterpret_cast<float4*>(gp[t] + o) = make_float4(g[0], g[1 % VEC], g[2 % VEC], g[3 % VEC]);
        else gp[t][o] = g[0];
      }
    }
  }
  // ---- per-CTA partials ----------------------------------------------------------------------------------
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float v = warp_sum(acc[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  const int tile = blockIdx.y * tiles_x + blockIdx.x;
  if (threadIdx.x < 4) {
    float t = 0.0f;
    for (int wv = 0; wv < SM_NT / 32; wv++) t += red[wv][threadIdx.x];
    a.workspace[ws.partials + ((size_t)b * tiles + tile) * 4 + threadIdx.x] = t;
  }
  // ---- ticketed reduction: last CTA of the sample, then last sample ----------------------------------------
  unsigned* tickets = reinterpret_cast<unsigned*>(a.workspace + ws.tickets);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + b, 1u) == (unsigned)(tiles - 1)) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const volatile float* part = a.workspace + ws.partials + (size_t)b * tiles * 4;
  if (warp < 2 * ND) {
    double s = 0.0;
    for (int t = lane; t < tiles; t += 32) s += (double)part[(size_t)t * 4 + warp];
    s = warp_sum(s);
    if (lane == 0) tot[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* stats = a.workspace + ws.stats + (size_t)b * 4;
    for (int t = 0; t < ND; t++) {
      const float sc = a.normalise ? 1.0f / ((float)tot[t * 2] / (float)hw + 1e-7f) : 1.0f;
      stats[t * 2] = (float)((double)sc * tot[t * 2 + 1]);   // L_b
      stats[t * 2 + 1] = sc;
      if (a.stats) { a.stats[(size_t)b * 4 + t * 2] = stats[t * 2]; a.stats[(size_t)b * 4 + t * 2 + 1] = sc; }
    }
    __threadfence();
    if (atomicAdd(tickets + a.batch, 1u) == (unsigned)(a.batch - 1)) {
      __threadfence();
      const volatile float* st = a.workspace + ws.stats;
      for (int t = 0; t < ND; t++) {
        double s = 0.0;
        for (int i = 0; i < a.batch; i++) s += (double)st[(size_t)i * 4 + t * 2];
        (t == 0 ? a.loss : a.loss_b)[0] = (float)s;
      }
    }
  }
}

__global__ void __launch_bounds__(SM_NT) smooth_fix_kernel(const mal_smooth_args a, const int tiles, const int term) {
  const SmoothWs ws = smooth_ws(a.batch, tiles);
  const size_t hw = (size_t)a.height * a.width, total = (size_t)a.batch * hw;
  float* g = term == 0 ? a.grad_disp : a.grad_disp_b;
  for (size_t i = (size_t)blockIdx.x * SM_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * SM_NT) {
    const int b = (int)(i / hw);
    const float lb = a.workspace[ws.stats + (size_t)b * 4 + term * 2], sc = a.workspace[ws.stats + (size_t)b * 4 + term * 2 + 1];
    g[i] = (g[i] - lb / (float)hw) * sc;
  }
}

inline int smooth_tiles(int height, int width, int vec, int* tx) {
  int x = (width + 32 * vec - 1) / (32 * vec), y = (height + SM_TH - 1) / SM_TH;
  if (tx) *tx = x;
  return x * y;
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_smooth_workspace_floats(int batch, int height, int width) {
  return smooth_ws(batch, smooth_tiles(height, width, 1, nullptr)).total;   // the scalar tiling is the larger one
}

extern "C" int mal_smooth_forward(const mal_smooth_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_smooth_forward: args is NULL");
  const mal_smooth_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height >= 2 && a.width >= 2, "mal_smooth_forward: bad shape %dx%dx%d", a.batch,
              a.height, a.width);
  MAL_REQUIRE(a.batch <= 65535, "mal_smooth_forward: batch %d exceeds gridDim.z", a.batch);
  MAL_REQUIRE(a.disp && a.img && a.workspace && a.loss, "mal_smooth_forward: disp/img/workspace/loss are required");
  if (a.with_grad) MAL_REQUIRE(a.grad_disp, "mal_smooth_forward: with_grad needs grad_disp");
  const bool dual = a.disp_b != nullptr;
  if (dual) MAL_REQUIRE(a.loss_b && (!a.with_grad || a.grad_disp_b), "mal_smooth_forward: disp_b needs loss_b (and grad_disp_b)");
  if (a.defer_fix) MAL_REQUIRE(a.stats, "mal_smooth_forward: defer_fix needs stats");
  cudaStream_t st = (cudaStream_t)stream;
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = a.width % 4 == 0 && al(a.disp) && al(a.img) && al(a.grad_disp) && al(a.disp_b) && al(a.grad_disp_b);
  int tiles_x;
  const int tiles = smooth_tiles(a.height, a.width, vec ? 4 : 1, &tiles_x);
  const SmoothWs ws = smooth_ws(a.batch, tiles);
  cudaMemsetAsync(a.workspace + ws.tickets, 0, (size_t)(a.batch + 1) * sizeof(unsigned), st);
  // gdx.mean() over (B,1,H,W-1) and gdy.mean() over (B,1,H-1,W)
  const float inv_nx = (float)(1.0 / ((double)a.batch * a.height * (a.width - 1)));
  const float inv_ny = (float)(1.0 / ((double)a.batch * (a.height - 1) * a.width));
  dim3 grid(tiles_x, tiles / tiles_x, a.batch);
  if (vec) {
    if (dual) launch(smooth_kernel<4, true>, grid, dim3(SM_NT), 0, st, a, inv_nx, inv_ny, tiles_x, tiles);
    else launch(smooth_kernel<4, false>, grid, dim3(SM_NT), 0, st, a, inv_nx, inv_ny, tiles_x, tiles);
  } else {
    if (dual) launch(smooth_kernel<1, true>, grid, dim3(SM_NT), 0, st, a, inv_nx, inv_ny, tiles_x, tiles);
    else launch(smooth_kernel<1, false>, grid, dim3(SM_NT), 0, st, a, inv_nx, inv_ny, tiles_x, tiles);
  }
  int rc = check_launch("smooth_kernel");
  if (rc) return rc;
  if (a.with_grad && a.normalise && !a.defer_fix) {
    size_t total = (size_t)a.batch * a.height * a.width;
    size_t blk = (total + SM_NT - 1) / SM_NT;
    if (blk > 148 * 8) blk = 148 * 8;
    for (int term = 0; term < (dual ? 2 : 1); term++)
      launch(smooth_fix_kernel, dim3((unsigned)blk), dim3(SM_NT), 0, st, a, tiles, term);
    rc = check_launch("smooth_fix_kernel");
  }
  return rc;
}
