// layers.cu - the reference's geometry / SSIM layers as stand-alone kernels (forward + backward).
//
//   mal_backproject[_backward]  BackprojectDepth.forward      manydepth/layers.py:163-168
//   mal_project3d[_backward]    Project3D.forward             manydepth/layers.py:184-199,
//                                                             dualrefine/layers.py:216-226
//   mal_ssim[_backward]         SSIM.forward                  manydepth/layers.py:243-257
//
// The fused photometric kernel (photo.cu) never calls these: they exist so that each layer class
// of the reference keeps working on its own with the same bit-exact
// arithmetic (mal_math.cuh).  All are streaming kernels, one thread per pixel, coalesced planes.
#include "mal_math.cuh"

namespace mal {

constexpr int LY_NT = 256;

inline unsigned ly_blocks(size_t n) {
  size_t b = (n + LY_NT - 1) / LY_NT;
  return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// ---- BackprojectDepth ---------------------------------------------------------------------------
__global__ void __launch_bounds__(LY_NT) backproject_kernel(const float* __restrict__ depth,
                                                           const float* __restrict__ inv_K, int B, int H, int W,
                                                           float* __restrict__ out) {
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const int b = (int)(i / hw);
    const int p = (int)(i - (size_t)b * hw);
    const int y = p / W, x = p - y * W;
    const float* k = inv_K + b * 16;
    float iK[9] = {k[0], k[1], k[2], k[4], k[5], k[6], k[8], k[9], k[10]};
    Ray r = pixel_ray(iK, (float)x, (float)y);
    const float d = __ldg(depth + i);
    float* o = out + (size_t)b * 4 * hw + p;
    o[0] = xmul(d, r.x); o[hw] = xmul(d, r.y); o[2 * hw] = xmul(d, r.z); o[3 * hw] = 1.0f;
  }
}

__global__ void __launch_bounds__(LY_NT) backproject_bwd_kernel(const float* __restrict__ g_out,
                                                               const float* __restrict__ inv_K, int B, int H, int W,
                                                               float* __restrict__ g_depth) {
  const size_t hw = (size_t)H * W, total = (size_t)B * hw;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const int b = (int)(i / hw);
    const int p = (int)(i - (size_t)b * hw);
    const int y = p / W, x = p - y * W;
    const float* k = inv_K + b * 16;
    float iK[9] = {k[0], k[1], k[2], k[4], k[5], k[6], k[8], k[9], k[10]};
    Ray r = pixel_ray(iK, (float)x, (float)y);
    const float* g = g_out + (size_t)b * 4 * hw + p;
    g_depth[i] = __ldg(g) * r.x + __ldg(g + hw) * r.y + __ldg(g + 2 * hw) * r.z;
  }
}

// ---- Project3D ------------------------------------------------------------------------------------
struct ProjOut { float gx, gy, X, Y, Z, Zp; };

template <int CONV>
__device__ __forceinline__ ProjOut project_point(const float* P, float p0, float p1, float p2, float p3, float eps,
                                                 int H, int W) {
  ProjOut o;
  o.X = xfma(P[3], p3, xfma(P[2], p2, xfma(P[1], p1, xmul(P[0], p0))));
  o.Y = xfma(P[7], p3, xfma(P[6], p2, xfma(P[5], p1, xmul(P[4], p0))));
  o.Z = xfma(P[11], p3, xfma(P[10], p2, xfma(P[9], p1, xmul(P[8], p0))));
  o.Zp = xadd(o.Z, eps);
  float px = xdiv(o.X, o.Zp), py = xdiv(o.Y, o.Zp);
  if (CONV == MAL_CONV_MANYDEPTH) {
    o.gx = xmul(xsub(xdiv(px, (float)(W - 1)), 0.5f), 2.0f);
    o.gy = xmul(xsub(xdiv(py, (float)(H - 1)), 0.5f), 2.0f);
  } else {
    o.gx = xsub(xdiv(xmul(2.0f, xadd(px, 0.5f)), (float)W), 1.0f);
    o.gy = xsub(xdiv(xmul(2.0f, xadd(py, 0.5f)), (float)H), 1.0f);
  }
  return o;
}

template <int CONV>
__global__ void __launch_bounds__(LY_NT) project3d_kernel(const float* __restrict__ points,
                                                         const float* __restrict__ K, const float* __restrict__ T,
                                                         int B, int H, int W, float eps, float2* __restrict__ pix,
                                                         float* __restrict__ zout) {
  __shared__ float sP[12];
  const int b = blockIdx.y;
  const size_t hw = (size_t)H * W;
  if (threadIdx.x < 12) sP[threadIdx.x] = kt_entry(K + b * 16, T + b * 16, threadIdx.x / 4, threadIdx.x % 4);
  __syncthreads();
  const float* pt = points + (size_t)b * 4 * hw;
  for (size_t p = (size_t)blockIdx.x * LY_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * LY_NT) {
    ProjOut o = project_point<CONV>(sP, __ldg(pt + p), __ldg(pt + hw + p), __ldg(pt + 2 * hw + p),
                                    __ldg(pt + 3 * hw + p), eps, H, W);
    pix[(size_t)b * hw + p] = make_float2(o.gx, o.gy);
    if (zout) zout[(size_t)b * hw + p] = o.Z;
  }
}

// d(gx,gy[,z]) -> d points (B,4,HW) and per-CTA partials of d P (12)
template <int CONV>
__global__ void __launch_bounds__(LY_NT) project3d_bwd_kernel(const float* __restrict__ points,
                                                             const float* __restrict__ K,
                                                             const float* __restrict__ T,
                                                             const float2* __restrict__ g_pix,
                                                             const float* __restrict__ g_z, int B, int H, int W,
                                                             float eps, float* __restrict__ g_points,
                                                             float* __restrict__ partials) {
  __shared__ float sP[12];
  __shared__ float red[LY_NT / 32][12];
  const int b = blockIdx.y;
  const size_t hw = (size_t)H * W;
  if (threadIdx.x < 12) sP[threadIdx.x] = kt_entry(K + b * 16, T + b * 16, threadIdx.x / 4, threadIdx.x % 4);
  __syncthreads();
  const float* pt = points + (size_t)b * 4 * hw;
  const float sx = CONV == MAL_CONV_MANYDEPTH ? 2.0f / (float)(W - 1) : 2.0f / (float)W;
  const float sy = CONV == MAL_CONV_MANYDEPTH ? 2.0f / (float)(H - 1) : 2.0f / (float)H;
  float gP[12];
#pragma unroll
  for (int j = 0; j < 12; j++) gP[j] = 0.0f;
  for (size_t p = (size_t)blockIdx.x * LY_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * LY_NT) {
    float q[4] = {__ldg(pt + p), __ldg(pt + hw + p), __ldg(pt + 2 * hw + p), __ldg(pt + 3 * hw + p)};
    ProjOut o = project_point<CONV>(sP, q[0], q[1], q[2], q[3], eps, H, W);
    float2 g = g_pix[(size_t)b * hw + p];
    float iz = 1.0f / o.Zp;
    float gX = g.x * sx * iz, gY = g.y * sy * iz;
    float gZ = -(gX * o.X + gY * o.Y) * iz + (g_z ? g_z[(size_t)b * hw + p] : 0.0f);
#pragma unroll
    for (int j = 0; j < 4; j++) {
      g_points[((size_t)b * 4 + j) * hw + p] = gX * sP[j] + gY * sP[4 + j] + gZ * sP[8 + j];
      gP[j] += gX * q[j]; gP[4 + j] += gY * q[j]; gP[8 + j] += gZ * q[j];
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    float v = warp_sum(gP[j]);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float s = 0.0f;
    for (int wv = 0; wv < LY_NT / 32; wv++) s += red[wv][threadIdx.x];
    partials[((size_t)b * gridDim.x + blockIdx.x) * 12 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(32) project3d_bwd_finalize_kernel(const float* __restrict__ partials, int nblk,
                                                                   float* __restrict__ g_P) {
  const int b = blockIdx.x;
  for (int j = 0; j < 12; j++) {
    double s = 0.0;
    for (int t = threadIdx.x; t < nblk; t += 32) s += (double)partials[((size_t)b * nblk + t) * 12 + j];
    s = warp_sum(s);
    if (threadIdx.x == 0) g_P[b * 12 + j] = (float)s;
  }
}

// ---- SSIM -----------------------------------------------------------------------------------------
__device__ __forceinline__ void load_window(const float* __restrict__ plane, int y, int x, int H, int W, float* w) {
#pragma unroll
  for (int dy = -1; dy <= 1; dy++) {
    int ry = reflect_index(y + dy, H);
#pragma unroll
    for (int dx = -1; dx <= 1; dx++) w[(dy + 1) * 3 + dx + 1] = __ldg(plane + (size_t)ry * W + reflect_index(x + dx, W));
  }
}

__device__ __forceinline__ SsimTerms window_terms(const float* xw, const float* yw) {
  float mu_x = xdivc<9>(sum9(xw)), mu_y = xdivc<9>(sum9(yw));
  float exx = xdivc<9>(sum9_prod(xw, xw)), eyy = xdivc<9>(sum9_prod(yw, yw)), exy = xdivc<9>(sum9_prod(xw, yw));
  return ssim_terms(mu_x, mu_y, exx, eyy, exy);
}

__global__ void __launch_bounds__(LY_NT) ssim_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                    int planes, int H, int W, float* __restrict__ out) {
  const size_t hw = (size_t)H * W, total = (size_t)planes * hw;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const size_t pl = i / hw;
    const int p = (int)(i - pl * hw);
    const int py = p / W, px = p - py * W;
    float xw[9], yw[9];
    load_window(x + pl * hw, py, px, H, W, xw);
    load_window(y + pl * hw, py, px, H, W, yw);
    out[i] = clamp01(window_terms(xw, yw).v);
  }
}

// pass 1: g * d v / d{mu_x, mu_y, E[xx]|E[yy], E[xy]} per window centre -> 4 planes
__global__ void __launch_bounds__(LY_NT) ssim_bwd_coef_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ y,
                                                             const float* __restrict__ g, int planes, int H, int W,
                                                             float* __restrict__ coef) {
  const size_t hw = (size_t)H * W, total = (size_t)planes * hw;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const size_t pl = i / hw;
    const int p = (int)(i - pl * hw);
    const int py = p / W, px = p - py * W;
    float xw[9], yw[9];
    load_window(x + pl * hw, py, px, H, W, xw);
    load_window(y + pl * hw, py, px, H, W, yw);
    SsimTerms t = window_terms(xw, yw);
    float ax, bx, gm, ay = 0.0f;
    ssim_coefs(t, ax, bx, gm);
    if (t.v >= 0.0f && t.v <= 1.0f) {
      float inv_d = 1.0f / t.d, r = t.n * inv_d;
      ay = -(t.mu_x * (t.Bq - t.A) - r * t.mu_y * (t.D - t.Cq)) * inv_d;
    }
    const float s = __ldg(g + i) * (1.0f / 9.0f);
    coef[i] = ax * s; coef[total + i] = ay * s; coef[2 * total + i] = bx * s; coef[3 * total + i] = gm * s;
  }
}

// pass 2: gather the 3x3 neighbourhood of window centres (ReflectionPad2d multiplicities)
__global__ void __launch_bounds__(LY_NT) ssim_bwd_gather_kernel(const float* __restrict__ x,
                                                               const float* __restrict__ y,
                                                               const float* __restrict__ coef, int planes, int H,
                                                               int W, float* __restrict__ gx, float* __restrict__ gy) {
  const size_t hw = (size_t)H * W, total = (size_t)planes * hw;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const size_t pl = i / hw;
    const int p = (int)(i - pl * hw);
    const int qy = p / W, qx = p - qy * W;
    const float xq = __ldg(x + i), yq = __ldg(y + i);
    float ax = 0.f, ay = 0.f;
    for (int dy = -1; dy <= 1; dy++) {
      int cy = qy + dy;
      if (cy < 0 || cy >= H) continue;
      float my = ((cy == 0 && dy == -1) || (cy == H - 1 && dy == 1)) ? 2.0f : 1.0f;
      for (int dx = -1; dx <= 1; dx++) {
        int cx = qx + dx;
        if (cx < 0 || cx >= W) continue;
        float m = my * (((cx == 0 && dx == -1) || (cx == W - 1 && dx == 1)) ? 2.0f : 1.0f);
        size_t c = pl * hw + (size_t)cy * W + cx;
        float a_x = __ldg(coef + c), a_y = __ldg(coef + total + c), be = __ldg(coef + 2 * total + c),
              ga = __ldg(coef + 3 * total + c);
        ax += m * (a_x + 2.0f * xq * be + yq * ga);
        ay += m * (a_y + 2.0f * yq * be + xq * ga);
      }
    }
    gx[i] = ax;
    if (gy) gy[i] = ay;
  }
}

// ---- F.grid_sample (bilinear; border or zeros padding) ------------------------------------------------
// Forward with the arithmetic of ATen's vectorised CPU kernel (so a materialised warp is bit-identical
// to the reference's), backward w.r.t. the grid (source images are data at every reference call site).
template <int CONV>
__device__ __forceinline__ Taps grid_taps(float gx, float gy, int H, int W, int border, float* gmx, float* gmy) {
  float ux = unnormalize<CONV>(gx, W), uy = unnormalize<CONV>(gy, H);
  *gmx = *gmy = 1.0f;
  if (border) {   // clip_coordinates_set_grad: the coordinate gradient vanishes where clamped
    *gmx = (ux <= 0.0f || ux >= (float)(W - 1)) ? 0.0f : 1.0f;
    *gmy = (uy <= 0.0f || uy >= (float)(H - 1)) ? 0.0f : 1.0f;
    ux = fminf((float)(W - 1), fmaxf(ux, 0.0f));
    uy = fminf((float)(H - 1), fmaxf(uy, 0.0f));
  }
  Taps t = make_taps(ux, uy, H, W);
  if (!(ux > -2.0f && ux < (float)W + 1.0f && uy > -2.0f && uy < (float)H + 1.0f)) t.v00 = t.v01 = t.v10 = t.v11 = false;
  return t;
}

template <int CONV>
__global__ void __launch_bounds__(LY_NT) grid_sample_kernel(const float* __restrict__ img,
                                                           const float2* __restrict__ grid, int B, int C, int H, int W,
                                                           int Ho, int Wo, int border, float* __restrict__ out) {
  const size_t hwo = (size_t)Ho * Wo, hw = (size_t)H * W, total = (size_t)B * hwo;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const size_t b = i / hwo, p = i - b * hwo;
    const float2 g = grid[i];
    float mx, my;
    const Taps t = grid_taps<CONV>(g.x, g.y, H, W, border, &mx, &my);
    for (int c = 0; c < C; c++) out[(b * C + c) * hwo + p] = bilinear(img + (b * C + c) * hw, t);
  }
}

template <int CONV>
__global__ void __launch_bounds__(LY_NT) grid_sample_bwd_kernel(const float* __restrict__ img,
                                                               const float2* __restrict__ grid,
                                                               const float* __restrict__ g_out, int B, int C, int H,
                                                               int W, int Ho, int Wo, int border,
                                                               float2* __restrict__ g_grid) {
  const size_t hwo = (size_t)Ho * Wo, hw = (size_t)H * W, total = (size_t)B * hwo;
  // d(unnormalised)/d(grid): (size-1)/2 with align_corners, size/2 without
  const float sx = CONV == MAL_CONV_MANYDEPTH ? (float)(W - 1) * 0.5f : (float)W * 0.5f;
  const float sy = CONV == MAL_CONV_MANYDEPTH ? (float)(H - 1) * 0.5f : (float)H * 0.5f;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * LY_NT) {
    const size_t b = i / hwo, p = i - b * hwo;
    const float2 g = grid[i];
    float mx, my;
    const Taps t = grid_taps<CONV>(g.x, g.y, H, W, border, &mx, &my);
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < C; c++) {
      float v00, v01, v10, v11;
      bilinear(img + (b * C + c) * hw, t, &v00, &v01, &v10, &v11);
      const float go = __ldg(g_out + (b * C + c) * hwo + p);
      gix += go * ((v01 - v00) * (1.0f - t.ty) + (v11 - v10) * t.ty);
      giy += go * ((v10 - v00) * (1.0f - t.tx) + (v11 - v01) * t.tx);
    }
    g_grid[i] = make_float2(gix * mx * sx, giy * my * sy);
  }
}


// ---- F.interpolate(x, [H, W], mode="bilinear", align_corners=False) ---------------------------------
// (manydepth/trainer.py:1093-1094, dualrefine/trainer.py:412-413, dynamicdepth/trainer.py:915-916:
// every trainer brings the low-resolution disparities to full resolution before the warp).
// Forward: one thread per output pixel, ATen's CPU arithmetic (mal_math.cuh).  Backward: the adjoint as a
// gather - one thread per INPUT pixel visits the few output pixels whose 2x2 footprint contains it -
// so the gradient is deterministic and needs no atomics.
__global__ void __launch_bounds__(LY_NT) upsample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                  int planes, int ih, int iw, int oh, int ow) {
  const size_t n = (size_t)planes * oh * ow;
  const float sy = up_scale(ih, oh), sx = up_scale(iw, ow);
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < n; i += (size_t)gridDim.x * LY_NT) {   // grid-stride
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
    const size_t pl = i / ((size_t)ow * oh);
    if (ih == oh && iw == ow) { out[i] = __ldg(in + i); continue; }
    const UpAxis ay = up_axis(oy, ih, sy), ax = up_axis(ox, iw, sx);
    out[i] = upsample_at(in + pl * ih * iw, iw, ay, ax);
  }
}

__global__ void __launch_bounds__(LY_NT) upsample_bilinear_bwd_kernel(const float* __restrict__ gout,
                                                                      float* __restrict__ gin, int planes, int ih,
                                                                      int iw, int oh, int ow) {
  const size_t n = (size_t)planes * ih * iw;
  const float sy = up_scale(ih, oh), sx = up_scale(iw, ow);
  const float ry = (float)oh / (float)ih, rx = (float)ow / (float)iw;
  for (size_t i = (size_t)blockIdx.x * LY_NT + threadIdx.x; i < n; i += (size_t)gridDim.x * LY_NT) {   // grid-stride
    const int ix = (int)(i % iw), iy = (int)((i / iw) % ih);
    const size_t pl = i / ((size_t)iw * ih);
    if (ih == oh && iw == ow) { gin[i] = __ldg(gout + i); continue; }
    // output rows / columns whose source index can fall in [i-1, i+1]
    const int y_lo = max(0, (int)floorf((iy - 1) * ry) - 1), y_hi = min(oh - 1, (int)ceilf((iy + 2) * ry) + 1);
    const int x_lo = max(0, (int)floorf((ix - 1) * rx) - 1), x_hi = min(ow - 1, (int)ceilf((ix + 2) * rx) + 1);
    const float* g = gout + pl * oh * ow;
    float acc = 0.0f;
    for (int oy = y_lo; oy <= y_hi; oy++) {
      const UpAxis ay = up_axis(oy, ih, sy);
      const float wy = (ay.i0 == iy ? ay.l0 : 0.0f) + (ay.i1 == iy ? ay.l1 : 0.0f);
      if (wy == 0.0f) continue;
      float row = 0.0f;
      for (int ox = x_lo; ox <= x_hi; ox++) {
        const UpAxis ax = up_axis(ox, iw, sx);
        const float wx = (ax.i0 == ix ? ax.l0 : 0.0f) + (ax.i1 == ix ? ax.l1 : 0.0f);
        if (wx != 0.0f) row += wx * __ldg(g + (size_t)oy * ow + ox);
      }
      acc += wy * row;
    }
    gin[i] = acc;
  }
}

}  // namespace mal

using namespace mal;

extern "C" int mal_backproject(const float* depth, const float* inv_K, int batch, int height, int width, float* out,
                               mal_stream_t stream) {
  MAL_REQUIRE(depth && inv_K && out && batch > 0 && height > 0 && width > 0, "mal_backproject: bad arguments");
  launch(backproject_kernel, dim3(ly_blocks((size_t)batch * height * width)), dim3(LY_NT), 0, (cudaStream_t)stream,
         depth, inv_K, batch, height, width, out);
  return check_launch("backproject_kernel");
}

extern "C" int mal_backproject_backward(const float* grad_out, const float* inv_K, int batch, int height, int width,
                                        float* grad_depth, mal_stream_t stream) {
  MAL_REQUIRE(grad_out && inv_K && grad_depth && batch > 0 && height > 0 && width > 0,
              "mal_backproject_backward: bad arguments");
  launch(backproject_bwd_kernel, dim3(ly_blocks((size_t)batch * height * width)), dim3(LY_NT), 0,
         (cudaStream_t)stream, grad_out, inv_K, batch, height, width, grad_depth);
  return check_launch("backproject_bwd_kernel");
}

static unsigned proj_blocks(int height, int width) {
  size_t b = ((size_t)height * width + LY_NT - 1) / LY_NT;
  return (unsigned)(b > 256 ? 256 : b);
}

extern "C" size_t mal_project3d_partials_floats(int batch, int height, int width) {
  return (size_t)batch * proj_blocks(height, width) * 12;
}

extern "C" int mal_project3d(const float* points, const float* K, const float* T, int batch, int height, int width,
                             int convention, float eps, float* pix, float* z, mal_stream_t stream) {
  MAL_REQUIRE(points && K && T && pix && batch > 0 && batch <= 65535 && height > 1 && width > 1,
              "mal_project3d: bad arguments");
  MAL_REQUIRE(convention == MAL_CONV_MANYDEPTH || convention == MAL_CONV_DUALREFINE, "mal_project3d: bad convention");
  dim3 grid(proj_blocks(height, width), batch);
  if (convention == MAL_CONV_MANYDEPTH)
    launch(project3d_kernel<MAL_CONV_MANYDEPTH>, grid, dim3(LY_NT), 0, (cudaStream_t)stream, points, K, T, batch,
           height, width, eps, reinterpret_cast<float2*>(pix), z);
  else
    launch(project3d_kernel<MAL_CONV_DUALREFINE>, grid, dim3(LY_NT), 0, (cudaStream_t)stream, points, K, T, batch,
           height, width, eps, reinterpret_cast<float2*>(pix), z);
  return check_launch("project3d_kernel");
}

extern "C" int mal_project3d_backward(const float* points, const float* K, const float* T, const float* grad_pix,
                                      const float* grad_z, int batch, int height, int width, int convention,
                                      float eps, float* grad_points, float* grad_P, float* partials,
                                      mal_stream_t stream) {
  MAL_REQUIRE(points && K && T && grad_pix && grad_points && grad_P && partials && batch > 0 && batch <= 65535 &&
                  height > 1 && width > 1,
              "mal_project3d_backward: bad arguments");
  MAL_REQUIRE(convention == MAL_CONV_MANYDEPTH || convention == MAL_CONV_DUALREFINE,
              "mal_project3d_backward: bad convention");
  dim3 grid(proj_blocks(height, width), batch);
  cudaStream_t st = (cudaStream_t)stream;
  if (convention == MAL_CONV_MANYDEPTH)
    launch(project3d_bwd_kernel<MAL_CONV_MANYDEPTH>, grid, dim3(LY_NT), 0, st, points, K, T,
           reinterpret_cast<const float2*>(grad_pix), grad_z, batch, height, width, eps, grad_points, partials);
  else
    launch(project3d_bwd_kernel<MAL_CONV_DUALREFINE>, grid, dim3(LY_NT), 0, st, points, K, T,
           reinterpret_cast<const float2*>(grad_pix), grad_z, batch, height, width, eps, grad_points, partials);
  int rc = check_launch("project3d_bwd_kernel");
  if (rc) return rc;
  launch(project3d_bwd_finalize_kernel, dim3(batch), dim3(32), 0, st, (const float*)partials, (int)grid.x, grad_P);
  return check_launch("project3d_bwd_finalize_kernel");
}

extern "C" int mal_ssim(const float* x, const float* y, int planes, int height, int width, float* out,
                        mal_stream_t stream) {
  MAL_REQUIRE(x && y && out && planes > 0 && height >= 2 && width >= 2, "mal_ssim: bad arguments");
  launch(ssim_kernel, dim3(ly_blocks((size_t)planes * height * width)), dim3(LY_NT), 0, (cudaStream_t)stream, x, y,
         planes, height, width, out);
  return check_launch("ssim_kernel");
}

extern "C" int mal_ssim_backward(const float* x, const float* y, const float* grad_out, int planes, int height,
                                 int width, float* grad_x, float* grad_y, float* workspace, mal_stream_t stream) {
  MAL_REQUIRE(x && y && grad_out && grad_x && workspace && planes > 0 && height >= 2 && width >= 2,
              "mal_ssim_backward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = ly_blocks((size_t)planes * height * width);
  launch(ssim_bwd_coef_kernel, dim3(nb), dim3(LY_NT), 0, st, x, y, grad_out, planes, height, width, workspace);
  int rc = check_launch("ssim_bwd_coef_kernel");
  if (rc) return rc;
  launch(ssim_bwd_gather_kernel, dim3(nb), dim3(LY_NT), 0, st, x, y, (const float*)workspace, planes, height, width,
         grad_x, grad_y);
  return check_launch("ssim_bwd_gather_kernel");
}

extern "C" int mal_grid_sample(const float* img, const float* grid, int batch, int channels, int height, int width,
                               int out_height, int out_width, int align_corners, int border, float* out,
                               mal_stream_t stream) {
  MAL_REQUIRE(img && grid && out && batch > 0 && channels > 0 && height > 1 && width > 1 && out_height > 0 &&
                  out_width > 0,
              "mal_grid_sample: bad arguments");
  const unsigned nb = ly_blocks((size_t)batch * out_height * out_width);
  if (align_corners)
    launch(grid_sample_kernel<MAL_CONV_MANYDEPTH>, dim3(nb), dim3(LY_NT), 0, (cudaStream_t)stream, img,
           reinterpret_cast<const float2*>(grid), batch, channels, height, width, out_height, out_width, border, out);
  else
    launch(grid_sample_kernel<MAL_CONV_DUALREFINE>, dim3(nb), dim3(LY_NT), 0, (cudaStream_t)stream, img,
           reinterpret_cast<const float2*>(grid), batch, channels, height, width, out_height, out_width, border, out);
  return check_launch("grid_sample_kernel");
}

extern "C" int mal_grid_sample_backward(const float* img, const float* grid, const float* grad_out, int batch,
                                        int channels, int height, int width, int out_height, int out_width,
                                        int align_corners, int border, float* grad_grid, mal_stream_t stream) {
  MAL_REQUIRE(img && grid && grad_out && grad_grid && batch > 0 && channels > 0 && height > 1 && width > 1 &&
                  out_height > 0 && out_width > 0,
              "mal_grid_sample_backward: bad arguments");
  const unsigned nb = ly_blocks((size_t)batch * out_height * out_width);
  if (align_corners)
    launch(grid_sample_bwd_kernel<MAL_CONV_MANYDEPTH>, dim3(nb), dim3(LY_NT), 0, (cudaStream_t)stream, img,
           reinterpret_cast<const float2*>(grid), grad_out, batch, channels, height, width, out_height, out_width,
           border, reinterpret_cast<float2*>(grad_grid));
  else
    launch(grid_sample_bwd_kernel<MAL_CONV_DUALREFINE>, dim3(nb), dim3(LY_NT), 0, (cudaStream_t)stream, img,
           reinterpret_cast<const float2*>(grid), grad_out, batch, channels, height, width, out_height, out_width,
           border, reinterpret_cast<float2*>(grad_grid));
  return check_launch("grid_sample_bwd_kernel");
}

extern "C" int mal_upsample_bilinear(const float* in, int planes, int in_height, int in_width, int out_height,
                                     int out_width, float* out, mal_stream_t stream) {
  MAL_REQUIRE(in && out && planes > 0 && in_height > 0 && in_width > 0 && out_height > 0 && out_width > 0,
              "mal_upsample_bilinear: bad arguments");
  launch(upsample_bilinear_kernel, dim3(ly_blocks((size_t)planes * out_height * out_width)), dim3(LY_NT), 0,
         (cudaStream_t)stream, in, out, planes, in_height, in_width, out_height, out_width);
  return check_launch("upsample_bilinear_kernel");
}

extern "C" int mal_upsample_bilinear_backward(const float* grad_out, int planes, int in_height, int in_width,
                                              int out_height, int out_width, float* grad_in, mal_stream_t stream) {
  MAL_REQUIRE(grad_out && grad_in && planes > 0 && in_height > 0 && in_width > 0 && out_height > 0 && out_width > 0,
              "mal_upsample_bilinear_backward: bad arguments");
  launch(upsample_bilinear_bwd_kernel, dim3(ly_blocks((size_t)planes * in_height * in_width)), dim3(LY_NT), 0,
         (cudaStream_t)stream, grad_out, grad_in, planes, in_height, in_width, out_height, out_width);
  return check_launch("upsample_bilinear_bwd_kernel");
}
