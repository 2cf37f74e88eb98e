// temporal.cu - MAL's temporal hint inside one training step, batched, for sm_100a.
//
// The reference's `--temporal` teacher pass (manydepth/trainer.py:1078-1165) materialises the two warped
// source images, runs image_synthesis on them (dyn_utils.py:121-170: per sample, generate_dynamic_instance
// :38-119 on the matched instance masks) and feeds the synthesised images to compute_mono_losses as two more
// candidates of the per-pixel min; autograd then carries d loss / d syn back through the copies into the
// warped images and on into depth and pose.  Here that is four launches for the whole batch:
//
//   tw_warp_kernel      outputs[("color", f, 0)] for f = -1, +1 from the disparity: disp -> depth ->
//                       backproject -> project -> bilinear gather (border), arithmetic as in photo.cu
//   ts_extents_kernel   bounding extents of every instance of every sample in ONE pass over the masks
//                       (dyn_utils.py:52-78; index-weighted sums: row 0 / column 0 never count)
//   ts_compose_kernel   half displacement with round-half-even (:80-100), background swap (:102-112),
//                       shifted copies summed in instance order (fill_dynamic_obj :5-36, torch's cascade
//                       summation), final where(mask_or, synthesised, original) (:114-118)
//   tb_backward_kernel  gathers d loss / d syn back through the composition (which pixels of the warped
//                       images each synthesised pixel copied) and, in the same thread, chains through the
//                       bilinear sampler, the projection and the backprojection into d/d disparity
//                       (accumulated onto the photometric kernel's plane) and per-CTA d/d(K@T) partials,
//                       which photo_finalize_kernel adds to the photometric pass's own.
//
// Instance masks are PACKED: one 32-bit word per pixel and frame, bit n = instance n (N <= 32).  A pixel's
// membership tests (mask_or, the background swaps, a shifted instance's coverage) are one or two word loads
// instead of N byte loads, and the masks of a batch are 1 MB per frame instead of N MB.  ts_pack_kernel
// converts Mask2Former-shaped (N,H,W) bool masks.
#include "mal_math.cuh"

namespace mal {

constexpr int TW_NT = 256;

// ---------------------------------------------------------------------------------------------------------
// warped source images
// ---------------------------------------------------------------------------------------------------------
template <int CONV>
__global__ void __launch_bounds__(TW_NT) tw_warp_kernel(const mal_temporal_args a, const float min_disp,
                                                       const float disp_range, const SizeDiv sdiv) {
  __shared__ Geom geom;
  const int b = blockIdx.z, H = a.height, W = a.width;
  const size_t HW = (size_t)H * W;
  if (threadIdx.x < 24) {
    const int f = threadIdx.x / 12, e = threadIdx.x % 12;
    geom.P[f][e] = kt_entry(a.K + b * 16, a.T[f] + b * 16, e / 4, e % 4);
  } else if (threadIdx.x >= 32 && threadIdx.x < 41) {
    const int e = threadIdx.x - 32;
    geom.iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t p = (size_t)y * W + x;
  float dv = __ldg(a.depth + (size_t)b * HW + p);
  if (a.depth_is_disp) dv = xdiv(1.0f, xadd(min_disp, xmul(disp_range, dv)));
  const Ray ray = pixel_ray(geom.iK, (float)x, (float)y);
#pragma unroll
  for (int f = 0; f < 2; f++) {
    const Sample s = project_pixel<CONV>(geom.P[f], ray, dv, a.eps, H, W, &sdiv);
    const Taps t = make_taps(s.ix, s.iy, H, W);
    const float* src = a.src[f] + (size_t)b * 3 * HW;
    float* out = a.warped[f] + (size_t)b * 3 * HW;
#pragma unroll
    for (int c = 0; c < 3; c++) out[c * HW + p] = bilinear(src + c * HW, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// packed instance masks
// ---------------------------------------------------------------------------------------------------------
// (N,H,W) bytes -> (H,W) words for one sample and frame; grid.y = sample * 2 + frame
__global__ void __launch_bounds__(256) ts_pack_kernel(const uint8_t* __restrict__ masks_last,
                                                     const uint8_t* __restrict__ masks_next, int nmax,
                                                     const int* __restrict__ counts, int hw,
                                                     unsigned* __restrict__ packed_last, unsigned* __restrict__ packed_next) {
  const int b = blockIdx.y >> 1, which = blockIdx.y & 1;
  const uint8_t* m = (which ? masks_next : masks_last) + (size_t)b * nmax * hw;
  unsigned* out = (which ? packed_next : packed_last) + (size_t)b * hw;
  const int n = counts[b];
  for (int p = blockIdx.x * 256 + threadIdx.x; p < hw; p += gridDim.x * 256) {
    unsigned w = 0;
    for (int i = 0; i < n; i++)
      if (m[(size_t)i * hw + p]) w |= 1u << i;
    out[p] = w;
  }
}

// Extents, one pass: every set bit of a word raises four running maxima of its instance in shared memory
// (low = max row, top = min row as 2^30 - row, right / left likewise; rows / columns 0 never count), then the
// CTA's maxima go to global memory with atomicMax.  Integer max is order-independent: deterministic.
// ext[b][frame][4][32], zero-initialised by a memset node ahead of the launch (0 = "no such pixel").
constexpr int TS_ENC = 1 << 30;
__global__ void __launch_bounds__(256) ts_extents_kernel(const unsigned* __restrict__ packed_last,
                                                        const unsigned* __restrict__ packed_next, int H, int W,
                                                        int rows_per_cta, int* __restrict__ ext) {
  __shared__ int sm[4][32];
  const int b = blockIdx.z, which = blockIdx.y;
  const unsigned* m = (which ? packed_next : packed_last) + (size_t)b * H * W;
  if (threadIdx.x < 128) sm[threadIdx.x >> 5][threadIdx.x & 31] = 0;
  __syncthreads();
  const int y0 = blockIdx.x * rows_per_cta, y1 = min(H, y0 + rows_per_cta);
  for (int i = y0 * W + threadIdx.x; i < y1 * W; i += 256) {
    unsigned w = m[i];
    if (w == 0u) continue;
    const int h = i / W, x = i - h * W;
    while (w) {
      const int n = __ffs((int)w) - 1;
      w &= w - 1;
      if (h >= 1) { atomicMax(&sm[0][n], h); atomicMax(&sm[1][n], TS_ENC - h); }
      if (x >= 1) { atomicMax(&sm[2][n], x); atomicMax(&sm[3][n], TS_ENC - x); }
    }
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int v = sm[threadIdx.x >> 5][threadIdx.x & 31];
    if (v) atomicMax(ext + (((size_t)b * 2 + which) * 4 + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31), v);
  }
}

struct TsDelta { int dxl[32], dyl[32]; };   // displacement of the "last" copies; the "next" copies use the negation

// dyn_utils.py:80-100 from the encoded extents of one sample
__device__ __forceinline__ void ts_delta(const int* __restrict__ ext_b, int n, int replace, int* dx, int* dy) {
  auto dec = [](int v) { return v ? TS_ENC - v : 0; };   // argmin over an all-"inf" row is index 0
  const int* l = ext_b;          // [4][32] of frame last
  const int* x = ext_b + 128;    // frame next
  const int dx0 = x[n] - l[n], dx1 = dec(x[32 + n]) - dec(l[32 + n]);        // [low_next-low_last, top_next-top_last]
  const int dy0 = x[64 + n] - l[64 + n], dy1 = dec(x[96 + n]) - dec(l[96 + n]);   // [right..., left...]
  const int sx = (abs(dx1) > abs(dx0)) ? dx1 : dx0;   // abs().argmax(): first index wins ties
  const int sy = (abs(dy1) > abs(dy0)) ? dy1 : dy0;
  int px = (int)rintf((float)sx / 2.0f), py = (int)rintf((float)sy / 2.0f);   // torch.round: half to even
  if (replace) {
    if (abs(px) < 3) px = 0;
    if (abs(py) < 3) py = 0;
  }
  *dx = px; *dy = py;
}

// fill_dynamic_obj at one pixel from a packed plane: sum over instances (cascade_sum: 16-element chunks) of
// source[:, y - dx, x - dy] where instance n covers the shifted location
__device__ __forceinline__ bool ts_fill(const unsigned* __restrict__ m, const int* dx, const int* dy, int sign,
                                        const float* __restrict__ source, int n_inst, int H, int W, size_t hw, int y,
                                        int x, float* acc) {
  float a0[3] = {0.f, 0.f, 0.f}, a1[3] = {0.f, 0.f, 0.f};
  bool hit = false;
  for (int n = 0; n < n_inst; n++) {
    const int sy = y - sign * dx[n], sx = x - sign * dy[n];
    float v[3] = {0.f, 0.f, 0.f};
    if (sy >= 0 && sy < H && sx >= 0 && sx < W && ((m[(size_t)sy * W + sx] >> n) & 1u)) {
      hit = true;
#pragma unroll
      for (int c = 0; c < 3; c++) v[c] = __ldg(source + c * hw + (size_t)sy * W + sx);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) a0[c] = xadd(a0[c], v[c]);
    if ((n & 15) == 15) {
#pragma unroll
      for (int c = 0; c < 3; c++) { a1[c] = xadd(a1[c], a0[c]); a0[c] = 0.0f; }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; c++) acc[c] = xadd(a0[c], a1[c]);
  return hit;
}

__global__ void __launch_bounds__(TW_NT) ts_compose_kernel(const mal_temporal_args a) {
  __shared__ TsDelta d;
  const int b = blockIdx.z, H = a.height, W = a.width;
  const size_t hw = (size_t)H * W;
  const int n_inst = a.counts[b];
  if (threadIdx.x < 32 && threadIdx.x < n_inst) {
    ts_delta(a.ext + (size_t)b * 256, threadIdx.x, a.replace, &d.dxl[threadIdx.x], &d.dyl[threadIdx.x]);
    if (blockIdx.x == 0 && blockIdx.y == 0) {   // kept for the backward pass
      a.deltas[(size_t)b * 64 + threadIdx.x] = d.dxl[threadIdx.x];
      a.deltas[(size_t)b * 64 + 32 + threadIdx.x] = d.dyl[threadIdx.x];
    }
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t p = (size_t)y * W + x;
  const unsigned* ml = a.packed_last + (size_t)b * hw;
  const unsigned* mn = a.packed_next + (size_t)b * hw;
  const float* il = a.warped[0] + (size_t)b * 3 * hw;
  const float* in_ = a.warped[1] + (size_t)b * 3 * hw;
  float* ol = a.syn[0] + (size_t)b * 3 * hw;
  float* on = a.syn[1] + (size_t)b * 3 * hw;
  const unsigned wl = n_inst ? ml[p] : 0u, wn = n_inst ? mn[p] : 0u;
  const bool m_or = (wl | wn) != 0u, bg = (wl & ~wn) != 0u, bg2 = (wn & ~wl) != 0u;
  float vl[3], vn[3];
#pragma unroll
  for (int c = 0; c < 3; c++) { vl[c] = __ldg(il + c * hw + p); vn[c] = __ldg(in_ + c * hw + p); }
  if (!m_or) {   // samples without instances, and every pixel no instance touches, keep the warped images
#pragma unroll
    for (int c = 0; c < 3; c++) { ol[c * hw + p] = vl[c]; on[c * hw + p] = vn[c]; }
    return;
  }
  float acc[3];
  bool any = ts_fill(ml, d.dxl, d.dyl, 1, il, n_inst, H, W, hw, y, x, acc);
#pragma unroll
  for (int c = 0; c < 3; c++) ol[c * hw + p] = any ? acc[c] : (bg ? vn[c] : vl[c]);
  any = ts_fill(mn, d.dxl, d.dyl, -1, in_, n_inst, H, W, hw, y, x, acc);
#pragma unroll
  for (int c = 0; c < 3; c++) on[c * hw + p] = any ? acc[c] : (bg2 ? vl[c] : vn[c]);
}

// ---------------------------------------------------------------------------------------------------------
// backward: d loss / d syn -> d loss / d warped images -> d/d disparity, d/d(K@T)
//   ori_last(q) = !or ? last(q) : any_l ? sum_n [mask_last_n(q-d)] last(q-d) : bg ? next(q) : last(q)
//   ori_next(q) = !or ? next(q) : any_n ? sum_n [mask_next_n(q+d)] next(q+d) : bg2 ? last(q) : next(q)
// ---------------------------------------------------------------------------------------------------------
constexpr int TB_NPART = 24;

template <int CONV>
__global__ void __launch_bounds__(TW_NT) tb_backward_kernel(const mal_temporal_args a, const float min_disp,
                                                           const float disp_range, const SizeDiv sdiv) {
  __shared__ Geom geom;
  __shared__ TsDelta d;
  __shared__ float red[TW_NT / 32][TB_NPART];
  __shared__ int any_grad;   // some pixel of this CTA carried a gradient into the poses
  const int b = blockIdx.z, H = a.height, W = a.width;
  const size_t hw = (size_t)H * W;
  const int n_inst = a.counts[b];
  if (threadIdx.x == 0) any_grad = 0;
  if (threadIdx.x < 24) {
    const int f = threadIdx.x / 12, e = threadIdx.x % 12;
    geom.P[f][e] = kt_entry(a.K + b * 16, a.T[f] + b * 16, e / 4, e % 4);
  } else if (threadIdx.x >= 32 && threadIdx.x < 41) {
    const int e = threadIdx.x - 32;
    geom.iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
  } else if (threadIdx.x >= 64 && threadIdx.x < 96) {
    const int n = threadIdx.x - 64;
    d.dxl[n] = n < n_inst ? a.deltas[(size_t)b * 64 + n] : 0;
    d.dyl[n] = n < n_inst ? a.deltas[(size_t)b * 64 + 32 + n] : 0;
  }
  __syncthreads();
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  float gP[24];
#pragma unroll
  for (int j = 0; j < 24; j++) gP[j] = 0.0f;
  if (x < W && y < H) {
    const size_t p = (size_t)y * W + x;
    const unsigned* ml = a.packed_last + (size_t)b * hw;
    const unsigned* mn = a.packed_next + (size_t)b * hw;
    const float* g_ol = a.grad_syn[0] + (size_t)b * 3 * hw;
    const float* g_on = a.grad_syn[1] + (size_t)b * 3 * hw;
    const unsigned wl = n_inst ? ml[p] : 0u, wn = n_inst ? mn[p] : 0u;
    float gl[3] = {0.f, 0.f, 0.f}, gn[3] = {0.f, 0.f, 0.f};
    float a_l[3], a_n[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { a_l[c] = __ldg(g_ol + c * hw + p); a_n[c] = __ldg(g_on + c * hw + p); }
    if ((wl | wn) == 0u) {
      // outside every mask: mask_or is false, ori == img (and no shifted copy reads a pixel without a mask bit)
#pragma unroll
      for (int c = 0; c < 3; c++) { gl[c] = a_l[c]; gn[c] = a_n[c]; }
    } else {
      // which branch did the forward take at p?
      bool any_l = false, any_n = false;
      for (int n = 0; n < n_inst; n++) {
        int sy = y - d.dxl[n], sx = x - d.dyl[n];
        if (sy >= 0 && sy < H && sx >= 0 && sx < W && ((ml[(size_t)sy * W + sx] >> n) & 1u)) any_l = true;
        sy = y + d.dxl[n]; sx = x + d.dyl[n];
        if (sy >= 0 && sy < H && sx >= 0 && sx < W && ((mn[(size_t)sy * W + sx] >> n) & 1u)) any_n = true;
      }
      const bool bg = (wl & ~wn) != 0u, bg2 = (wn & ~wl) != 0u;
#pragma unroll
      for (int c = 0; c < 3; c++) {
        if (!any_l) { if (bg) gn[c] += a_l[c]; else gl[c] += a_l[c]; }
        if (!any_n) { if (bg2) gl[c] += a_n[c]; else gn[c] += a_n[c]; }
      }
    }
    // shifted copies that read this pixel: instance n copies last(p) to q = p + d_n (next(p) to q = p - d_n)
    // wherever mask_or holds at q
    for (unsigned w = wl; w;) {
      const int n = __ffs((int)w) - 1;
      w &= w - 1;
      const int qy = y + d.dxl[n], qx = x + d.dyl[n];
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        const size_t q = (size_t)qy * W + qx;
        if ((ml[q] | mn[q]) != 0u) {
#pragma unroll
          for (int c = 0; c < 3; c++) gl[c] += __ldg(g_ol + c * hw + q);
        }
      }
    }
    for (unsigned w = wn; w;) {
      const int n = __ffs((int)w) - 1;
      w &= w - 1;
      const int qy = y - d.dxl[n], qx = x - d.dyl[n];
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        const size_t q = (size_t)qy * W + qx;
        if ((ml[q] | mn[q]) != 0u) {
#pragma unroll
          for (int c = 0; c < 3; c++) gn[c] += __ldg(g_on + c * hw + q);
        }
      }
    }
    if (a.grad_warped[0]) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        a.grad_warped[0][((size_t)b * 3 + c) * hw + p] = gl[c];
        a.grad_warped[1][((size_t)b * 3 + c) * hw + p] = gn[c];
      }
    }
    // chain through the sampler, the projection and the backprojection (as photo.cu phase C)
    if (a.grad_depth) {
      const float dv_in = __ldg(a.depth + (size_t)b * hw + p);
      const float dv = a.depth_is_disp ? xdiv(1.0f, xadd(min_disp, xmul(disp_range, dv_in))) : dv_in;
      const Ray ray = pixel_ray(geom.iK, (float)x, (float)y);
      const float cam[3] = {dv * ray.x, dv * ray.y, dv * ray.z};
      float gdepth = 0.0f;
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const float* g = f == 0 ? gl : gn;
        if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f) continue;
        any_grad = 1;   // (every writer stores 1)
        const float* P = geom.P[f];
        const Sample s = project_pixel<CONV>(P, ray, dv, a.eps, H, W, &sdiv);
        const Taps t = make_taps(s.ix, s.iy, H, W);
        const float* src = a.src[f] + (size_t)b * 3 * hw;
        float gix = 0.f, giy = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          float v00, v01, v10, v11;
          bilinear(src + c * hw, t, &v00, &v01, &v10, &v11);
          gix += g[c] * ((v01 - v00) * (1.0f - t.ty) + (v11 - v10) * t.ty);
          giy += g[c] * ((v10 - v00) * (1.0f - t.tx) + (v11 - v01) * t.tx);
        }
        gix *= s.gmx;
        giy *= s.gmy;
        const float iz = 1.0f / s.Zp;
        const float gX = gix * iz, gY = giy * iz;
        const float gZ = -(gX * s.X + gY * s.Y) * iz;
        gdepth += gX * (P[0] * ray.x + P[1] * ray.y + P[2] * ray.z) + gY * (P[4] * ray.x + P[5] * ray.y + P[6] * ray.z) +
                  gZ * (P[8] * ray.x + P[9] * ray.y + P[10] * ray.z);
        float* gp = gP + f * 12;
        gp[0] += gX * cam[0]; gp[1] += gX * cam[1]; gp[2] += gX * cam[2]; gp[3] += gX;
        gp[4] += gY * cam[0]; gp[5] += gY * cam[1]; gp[6] += gY * cam[2]; gp[7] += gY;
        gp[8] += gZ * cam[0]; gp[9] += gZ * cam[1]; gp[10] += gZ * cam[2]; gp[11] += gZ;
      }
      if (a.depth_is_disp) gdepth *= -disp_range * dv * dv;
      if (gdepth != 0.0f) a.grad_depth[(size_t)b * hw + p] += gdepth;   // onto the photometric kernel's plane
    }
  }
  // per-CTA d/d(K@T) partials.  The gradient lives inside and next to the instances: most CTAs carry none and
  // skip the 24 shuffle trees.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  const bool live_cta = any_grad != 0;
  if (live_cta) {
    const float v = warp_sum_transposed<24>(gP, lane);
    if (lane < 24) red[warp][lane] = v;
    __syncthreads();
  }
  if (threadIdx.x < TB_NPART && a.partials) {
    float s = 0.0f;
    if (live_cta) {
#pragma unroll
      for (int wv = 0; wv < TW_NT / 32; wv++) s += red[wv][threadIdx.x];
    }
    const size_t tiles = (size_t)gridDim.x * gridDim.y;
    a.partials[((size_t)b * tiles + blockIdx.y * gridDim.x + blockIdx.x) * TB_NPART + threadIdx.x] = s;
  }
}

// grad_P[b][v] += sum over the sample's tiles (fixed order: lanes stride the tiles, fp64 shuffle tree)
__global__ void __launch_bounds__(TB_NPART * 32) tb_reduce_kernel(const float* __restrict__ partials, int tiles,
                                                                 float* __restrict__ grad_P) {
  const int b = blockIdx.x, v = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int t = lane; t < tiles; t += 32) s += (double)partials[((size_t)b * tiles + t) * TB_NPART + v];
  s = warp_sum(s);
  if (lane == 0) grad_P[b * 24 + v] += (float)s;
}

}  // namespace mal

using namespace mal;

static int temporal_check(const mal_temporal_args& a, const char* who) {
  MAL_REQUIRE(a.batch > 0 && a.batch <= 65535 && a.height >= 2 && a.width >= 2, "%s: bad shape %dx%dx%d", who, a.batch,
              a.height, a.width);
  MAL_REQUIRE(a.convention == MAL_CONV_MANYDEPTH || a.convention == MAL_CONV_DUALREFINE, "%s: bad convention %d", who,
              a.convention);
  return MAL_OK;
}

extern "C" size_t mal_temporal_partials_floats(int batch, int height, int width) {
  return (size_t)batch * ((width + 31) / 32) * ((height + 7) / 8) * TB_NPART;
}

extern "C" int mal_temporal_warp(const mal_temporal_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_temporal_warp: args is NULL");
  const mal_temporal_args& a = *args;
  if (int rc = temporal_check(a, "mal_temporal_warp")) return rc;
  MAL_REQUIRE(a.src[0] && a.src[1] && a.depth && a.K && a.inv_K && a.T[0] && a.T[1] && a.warped[0] && a.warped[1],
              "mal_temporal_warp: src/depth/K/inv_K/T/warped are required");
  if (a.depth_is_disp) MAL_REQUIRE(a.min_depth > 0 && a.max_depth > a.min_depth, "mal_temporal_warp: bad depth range");
  const double lo = 1.0 / a.max_depth, hi = 1.0 / a.min_depth;
  const SizeDiv sdiv = size_div(a.height, a.width, a.convention);
  dim3 grid((a.width + 31) / 32, (a.height + 7) / 8, a.batch);
  if (a.convention == MAL_CONV_MANYDEPTH)
    launch(tw_warp_kernel<MAL_CONV_MANYDEPTH>, grid, dim3(TW_NT), 0, (cudaStream_t)stream, a, (float)lo, (float)(hi - lo), sdiv);
  else
    launch(tw_warp_kernel<MAL_CONV_DUALREFINE>, grid, dim3(TW_NT), 0, (cudaStream_t)stream, a, (float)lo, (float)(hi - lo), sdiv);
  return check_launch("tw_warp_kernel");
}

extern "C" int mal_temporal_pack_masks(const uint8_t* masks_last, const uint8_t* masks_next, const int32_t* counts,
                                       int batch, int nmax, int height, int width, uint32_t* packed_last,
                                       uint32_t* packed_next, mal_stream_t stream) {
  MAL_REQUIRE(masks_last && masks_next && counts && packed_last && packed_next, "mal_temporal_pack_masks: a required pointer is NULL");
  MAL_REQUIRE(batch > 0 && nmax > 0 && nmax <= 32 && height > 0 && width > 0,
              "mal_temporal_pack_masks: bad shape B=%d N=%d %dx%d (1 <= N <= 32)", batch, nmax, height, width);
  const int hw = height * width;
  launch(ts_pack_kernel, dim3(min((hw + 255) / 256, 148 * 4), batch * 2), dim3(256), 0, (cudaStream_t)stream, masks_last,
         masks_next, nmax, counts, hw, packed_last, packed_next);
  return check_launch("ts_pack_kernel");
}

extern "C" int mal_temporal_synthesis(const mal_temporal_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_temporal_synthesis: args is NULL");
  const mal_temporal_args& a = *args;
  if (int rc = temporal_check(a, "mal_temporal_synthesis")) return rc;
  MAL_REQUIRE(a.packed_last && a.packed_next && a.counts && a.warped[0] && a.warped[1] && a.syn[0] && a.syn[1] && a.ext &&
                  a.deltas,
              "mal_temporal_synthesis: packed masks/counts/warped/syn/ext/deltas are required");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(a.ext, 0, (size_t)a.batch * 256 * sizeof(int), st);
  const int rows = 8, ctas = (a.height + rows - 1) / rows;
  launch(ts_extents_kernel, dim3(ctas, 2, a.batch), dim3(256), 0, st, a.packed_last, a.packed_next, a.height, a.width, rows,
         a.ext);
  int rc = check_launch("ts_extents_kernel");
  if (rc) return rc;
  launch(ts_compose_kernel, dim3((a.width + 31) / 32, (a.height + 7) / 8, a.batch), dim3(TW_NT), 0, st, a);
  return check_launch("ts_compose_kernel");
}

extern "C" int mal_temporal_backward(const mal_temporal_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_temporal_backward: args is NULL");
  const mal_temporal_args& a = *args;
  if (int rc = temporal_check(a, "mal_temporal_backward")) return rc;
  MAL_REQUIRE(a.packed_last && a.packed_next && a.counts && a.deltas && a.grad_syn[0] && a.grad_syn[1],
              "mal_temporal_backward: packed masks/counts/deltas/grad_syn are required");
  MAL_REQUIRE((a.grad_warped[0] == nullptr) == (a.grad_warped[1] == nullptr), "mal_temporal_backward: give both grad_warped or none");
  if (a.grad_depth) {
    MAL_REQUIRE(a.src[0] && a.src[1] && a.depth && a.K && a.inv_K && a.T[0] && a.T[1] && a.partials && a.grad_P,
                "mal_temporal_backward: the chain into depth / pose needs src/depth/K/inv_K/T/partials/grad_P");
    if (a.depth_is_disp) MAL_REQUIRE(a.min_depth > 0 && a.max_depth > a.min_depth, "mal_temporal_backward: bad depth range");
  } else {
    MAL_REQUIRE(a.grad_warped[0], "mal_temporal_backward: nothing to compute (no grad_warped, no grad_depth)");
    MAL_REQUIRE(a.K && a.inv_K && a.T[0] && a.T[1], "mal_temporal_backward: K/inv_K/T are required");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const double lo = 1.0 / a.max_depth, hi = 1.0 / a.min_depth;
  const SizeDiv sdiv = size_div(a.height, a.width, a.convention);
  dim3 grid((a.width + 31) / 32, (a.height + 7) / 8, a.batch);
  if (a.convention == MAL_CONV_MANYDEPTH)
    launch(tb_backward_kernel<MAL_CONV_MANYDEPTH>, grid, dim3(TW_NT), 0, st, a, (float)lo, (float)(hi - lo), sdiv);
  else
    launch(tb_backward_kernel<MAL_CONV_DUALREFINE>, grid, dim3(TW_NT), 0, st, a, (float)lo, (float)(hi - lo), sdiv);
  int rc = check_launch("tb_backward_kernel");
  if (rc || !a.grad_depth) return rc;
  launch(tb_reduce_kernel, dim3(a.batch), dim3(TB_NPART * 32), 0, st, (const float*)a.partials, (int)(grid.x * grid.y), a.grad_P);
  return check_launch("tb_reduce_kernel");
}
