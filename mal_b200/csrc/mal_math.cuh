// mal_math.cuh - per-pixel device math of the MAL photometric path.
//
// Arithmetic contract: every value that can decide a comparison in the
// reference (min over candidates, automask, distillation argmin) is computed with the SAME
// sequence of IEEE-754 single operations the reference's torch CPU kernels execute, using
// the never-contracted x* helpers:
//   * (K@T)[:3,:]            sequential, unfused      (ATen small-bmm path; layers.py:185)
//   * invK[:3,:3]@[x,y,1]    k-sequential FMA chain   (BLAS sgemm;          layers.py:164)
//   * P@[cam;1]              k-sequential FMA chain   (BLAS sgemm;          layers.py:187)
//   * grid_sample bilinear   nw*a, then 3 FMAs        (GridSamplerKernel.cpp, contracted)
//   * avg_pool2d 3x3         row-major running sum, then /9   (AvgPoolKernel.cpp)
//   * SSIM formula, mean over channels ((c0+c1)+c2)/3, 0.85*s+0.15*l: one rounding per op
// The contract is replayed on the host against torch by the CPU twin (tests/emu): the golden and
// oracle tests demand torch.equal on every selection, min-reprojection and cost-volume value.
#pragma once
#include <cstdlib>

#include "mal_common.cuh"

namespace mal {

struct Geom {        // per-sample camera constants, staged in shared memory
  float P[2][12];    // (K@T_f)[:3,:] row-major, frames -1,+1
  float iK[9];       // inv_K[:3,:3] row-major
};

// (K@T)[:3,:] entry (i,j): ((K_i0*T_0j + K_i1*T_1j) + K_i2*T_2j) + K_i3*T_3j
__device__ __forceinline__ float kt_entry(const float* K, const float* T, int i, int j) {
  float acc = xmul(K[i * 4 + 0], T[0 * 4 + j]);
  acc = xadd(acc, xmul(K[i * 4 + 1], T[1 * 4 + j]));
  acc = xadd(acc, xmul(K[i * 4 + 2], T[2 * 4 + j]));
  acc = xadd(acc, xmul(K[i * 4 + 3], T[3 * 4 + j]));
  return acc;
}

struct Ray { float x, y, z; };

// inv_K[:3,:3] @ [px,py,1]
__device__ __forceinline__ Ray pixel_ray(const float* iK, float px, float py) {
  Ray r;
  r.x = xfma(iK[2], 1.0f, xfma(iK[1], py, xmul(iK[0], px)));
  r.y = xfma(iK[5], 1.0f, xfma(iK[4], py, xmul(iK[3], px)));
  r.z = xfma(iK[8], 1.0f, xfma(iK[7], py, xmul(iK[6], px)));
  return r;
}

struct Sample {
  float ix, iy;    // clipped source pixel coordinates
  float gmx, gmy;  // d(clipped)/d(unclipped): 0 where clamped (ATen clip_coordinates_set_grad)
  float X, Y, Zp;  // P@[cam;1] numerators and (z + eps)
};

struct GridPoint { float gx, gy, X, Y, Zp; };

// Project3D divides every coordinate by an image-size constant ((W-1), (H-1) or W, H).  For the constants below
// the 3-instruction sequence of xdivc (q = RN(x*r), rem = fma(-C,q,x), RN(q + rem*r), r = RN(1/C)) returns the
// correctly rounded x / C for EVERY finite float x - checked exhaustively over all 2^32 inputs on the host
// (tests/test_exact_arithmetic.py, MAL_EXHAUSTIVE=1; the only difference is the sign of a zero result for
// x = -0, which the following "- 0.5" / "- 1" absorbs).  Any other size keeps the IEEE division (48, 160, 192, 640
// - DualRefine's W, H - fail the check for subnormal quotients and are therefore not listed).
struct SizeDiv { float cw, rw, ch, rh; int fast; };
inline bool size_div_verified(int c) {
  static const int ok[] = {3, 9, 47, 95, 127, 159, 191, 255, 383, 511, 639, 1023};
  if (c > 0 && (c & (c - 1)) == 0) return true;   // a power of two divides exactly by multiplying
  for (int v : ok)
    if (v == c) return true;
  return false;
}
inline SizeDiv size_div(int H, int W, int convention) {
  const int dw = convention == MAL_CONV_MANYDEPTH ? W - 1 : W, dh = convention == MAL_CONV_MANYDEPTH ? H - 1 : H;
  SizeDiv d;
  d.cw = (float)dw; d.ch = (float)dh;
  d.rw = 1.0f / d.cw; d.rh = 1.0f / d.ch;
  d.fast = (size_div_verified(dw) && size_div_verified(dh)) ? 1 : 0;
#ifndef MAL_EMU
  if (getenv("MAL_NO_FASTDIV")) d.fast = 0;   // tuning runs only
#endif
  return d;
}
__device__ __forceinline__ float xdiv_size(float x, float c, float r, bool fast) {
  if (fast) {
    const float q = xmul(x, r);
    return xfma(xfma(-c, q, x), r, q);
  }
  return xdiv(x, c);
}

// depth * ray -> P@[cam;1] -> /(z+eps) -> Project3D normalisation to [-1,1]
template <int CONV>
__device__ __forceinline__ GridPoint project_grid(const float* P, Ray ray, float depth, float eps, int H, int W,
                                                  const SizeDiv* sd = nullptr) {
  float cx = xmul(depth, ray.x), cy = xmul(depth, ray.y), cz = xmul(depth, ray.z);
  GridPoint g;
  g.X = xfma(P[3], 1.0f, xfma(P[2], cz, xfma(P[1], cy, xmul(P[0], cx))));
  g.Y = xfma(P[7], 1.0f, xfma(P[6], cz, xfma(P[5], cy, xmul(P[4], cx))));
  float Z = xfma(P[11], 1.0f, xfma(P[10], cz, xfma(P[9], cy, xmul(P[8], cx))));
  g.Zp = xadd(Z, eps);
  float px = xdiv(g.X, g.Zp), py = xdiv(g.Y, g.Zp);
  const bool fast = sd != nullptr && sd->fast;
  const float cw = sd ? sd->cw : (float)(CONV == MAL_CONV_MANYDEPTH ? W - 1 : W);
  const float ch = sd ? sd->ch : (float)(CONV == MAL_CONV_MANYDEPTH ? H - 1 : H);
  const float rw = sd ? sd->rw : 0.0f, rh = sd ? sd->rh : 0.0f;
  if (CONV == MAL_CONV_MANYDEPTH) {
    g.gx = xmul(xsub(xdiv_size(px, cw, rw, fast), 0.5f), 2.0f);
    g.gy = xmul(xsub(xdiv_size(py, ch, rh, fast), 0.5f), 2.0f);
  } else {
    g.gx = xsub(xdiv_size(xmul(2.0f, xadd(px, 0.5f)), cw, rw, fast), 1.0f);
    g.gy = xsub(xdiv_size(xmul(2.0f, xadd(py, 0.5f)), ch, rh, fast), 1.0f);
  }
  return g;
}

// grid_sample's unnormalisation of a [-1,1] coordinate (align_corners = CONV==MANYDEPTH).
// ATen's vectorised CPU kernel computes (g + 1) * (size / 2) - 0.5 for align_corners=False and the
// compiler contracts the multiply-subtract into one FMA (measured against F.grid_sample for both
// padding modes); the align_corners=True form has nothing to contract.
template <int CONV>
__device__ __forceinline__ float unnormalize(float g, int size) {
  if (CONV == MAL_CONV_MANYDEPTH) return xmul(xadd(g, 1.0f), (float)(size - 1) / 2.0f);
  return xfma(xadd(g, 1.0f), (float)size / 2.0f, -0.5f);
}

// ... -> grid_sample unnormalise -> border clip
template <int CONV>
__device__ __forceinline__ Sample project_pixel(const float* P, Ray ray, float depth, float eps,
                                                int H, int W, const SizeDiv* sd = nullptr) {
  GridPoint g = project_grid<CONV>(P, ray, depth, eps, H, W, sd);
  Sample s;
  s.X = g.X; s.Y = g.Y; s.Zp = g.Zp;
  float ux = unnormalize<CONV>(g.gx, W), uy = unnormalize<CONV>(g.gy, H);
  // padding_mode="border": clamp, and the coordinate gradient vanishes where clamped
  s.gmx = (ux <= 0.0f || ux >= (float)(W - 1)) ? 0.0f : 1.0f;
  s.gmy = (uy <= 0.0f || uy >= (float)(H - 1)) ? 0.0f : 1.0f;
  s.ix = fminf((float)(W - 1), fmaxf(ux, 0.0f));
  s.iy = fminf((float)(H - 1), fmaxf(uy, 0.0f));
  return s;
}

struct Taps {
  int o00, o01, o10, o11;      // plane offsets (valid where the flag is set)
  bool v00, v01, v10, v11;     // in-bounds flags (out-of-bounds taps read as 0)
  float nw, ne, sw, se;        // bilinear weights
  float tx, ty;                // fractional parts
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W) {
  Taps t;
  float x0 = floorf(ix), y0 = floorf(iy);
  t.tx = xsub(ix, x0);
  t.ty = xsub(iy, y0);
  float e = xsub(1.0f, t.tx), s = xsub(1.0f, t.ty);
  t.nw = xmul(s, e); t.ne = xmul(s, t.tx); t.sw = xmul(t.ty, e); t.se = xmul(t.ty, t.tx);
  int xi = (int)x0, yi = (int)y0;
  bool xl = xi >= 0 && xi < W, xr = xi + 1 >= 0 && xi + 1 < W;
  bool yt = yi >= 0 && yi < H, yb = yi + 1 >= 0 && yi + 1 < H;
  t.v00 = xl && yt; t.v01 = xr && yt; t.v10 = xl && yb; t.v11 = xr && yb;
  t.o00 = yi * W + xi; t.o01 = t.o00 + 1; t.o10 = t.o00 + W; t.o11 = t.o10 + 1;
  return t;
}

__device__ __forceinline__ float bilinear(const float* __restrict__ plane, const Taps& t,
                                          float* a = nullptr, float* b = nullptr, float* c = nullptr,
                                          float* d = nullptr) {
  float v00 = t.v00 ? __ldg(plane + t.o00) : 0.0f;
  float v01 = t.v01 ? __ldg(plane + t.o01) : 0.0f;
  float v10 = t.v10 ? __ldg(plane + t.o10) : 0.0f;
  float v11 = t.v11 ? __ldg(plane + t.o11) : 0.0f;
  if (a) { *a = v00; *b = v01; *c = v10; *d = v11; }
  return xfma(v11, t.se, xfma(v10, t.sw, xfma(v01, t.ne, xmul(v00, t.nw))));
}

// ---- bilinear up-sampling of a plane (F.interpolate(mode="bilinear", align_corners=False)) ------
// ATen's vectorised CPU kernel (UpSampleKernel.cpp, cpu_upsample_linear), measured against torch 2.11:
//   source index  s = fma(in/out, o + 0.5, -0.5), clamped at 0;  i0 = min(int(s), in-1), i1 = i0 + (i0 < in-1)
//   l1 = clamp(s - i0, 0, 1), l0 = 1 - l1
//   value         fma(ly0, fma(lx0, v00, lx1*v01), ly1 * fma(lx0, v10, lx1*v11))
// ATen uses this kernel when out_h + out_w > 128 (_use_vectorized_kernel_cond_2d); every disparity on the
// reference's path is up-sampled to the full image (192 + 640), smaller outputs are only within ~1 ulp.
struct UpAxis { int i0, i1; float l0, l1; };
__device__ __forceinline__ UpAxis up_axis(int o, int in_size, float scale) {
  float s = xfma(scale, xadd((float)o, 0.5f), -0.5f);
  if (s < 0.0f) s = 0.0f;
  UpAxis a;
  a.i0 = min((int)s, in_size - 1);
  a.i1 = a.i0 + (a.i0 < in_size - 1 ? 1 : 0);
  a.l1 = fminf(fmaxf(xsub(s, (float)a.i0), 0.0f), 1.0f);
  a.l0 = xsub(1.0f, a.l1);
  return a;
}
__device__ __forceinline__ float up_scale(int in_size, int out_size) { return xdiv((float)in_size, (float)out_size); }
__device__ __forceinline__ float upsample_at(const float* __restrict__ plane, int w, const UpAxis& ay, const UpAxis& ax) {
  const float v00 = __ldg(plane + ay.i0 * w + ax.i0), v01 = __ldg(plane + ay.i0 * w + ax.i1);
  const float v10 = __ldg(plane + ay.i1 * w + ax.i0), v11 = __ldg(plane + ay.i1 * w + ax.i1);
  const float t0 = xfma(ax.l0, v00, xmul(ax.l1, v01)), t1 = xfma(ax.l0, v10, xmul(ax.l1, v11));
  return xfma(ay.l0, t0, xmul(ay.l1, t1));
}

// ---- SSIM ------------------------------------------------------------------------------
// running row-major sum of a 3x3 window, as avg_pool2d accumulates it
__device__ __forceinline__ float sum9(const float* w) {
  float s = w[0];
#pragma unroll
  for (int i = 1; i < 9; i++) s = xadd(s, w[i]);
  return s;
}
__device__ __forceinline__ float sum9_prod(const float* a, const float* b) {
  float s = xmul(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < 9; i++) s = xadd(s, xmul(a[i], b[i]));
  return s;
}

// packed forms: two predictions (the halves of each pair) against the same target window
__device__ __forceinline__ pk2 sum9(const pk2* w) {
  pk2 s = w[0];
#pragma unroll
  for (int i = 1; i < 9; i++) s = x2add(s, w[i]);
  return s;
}
__device__ __forceinline__ pk2 sum9_prod(const pk2* a, const pk2* b) {
  pk2 s = x2mul_nf(a[0], b[0]);   // also feeds an add: see the contraction hazard in mal_common.cuh
#pragma unroll
  for (int i = 1; i < 9; i++) s = x2add(s, x2mul_nf(a[i], b[i]));
  return s;
}

struct SsimTerms { float mu_x, mu_y, A, Bq, Cq, D, n, d, v; };

#define MAL_C1 0.0001f   /* 0.01**2 */
#define MAL_C2 0.0009f   /* 0.03**2 */

// SSIM loss value for one channel from the 5 pooled moments (layers.py:247-257)
__device__ __forceinline__ SsimTerms ssim_terms(float mu_x, float mu_y, float exx, float eyy, float exy) {
  SsimTerms t;
  t.mu_x = mu_x; t.mu_y = mu_y;
  float sig_x = xsub(exx, xmul(mu_x, mu_x));
  float sig_y = xsub(eyy, xmul(mu_y, mu_y));
  float sig_xy = xsub(exy, xmul(mu_x, mu_y));
  t.A = xadd(xmul(xmul(2.0f, mu_x), mu_y), MAL_C1);
  t.Bq = xadd(xmul(2.0f, sig_xy), MAL_C2);
  t.Cq = xadd(xadd(xmul(mu_x, mu_x), xmul(mu_y, mu_y)), MAL_C1);
  t.D = xadd(xadd(sig_x, sig_y), MAL_C2);
  t.n = xmul(t.A, t.Bq);
  t.d = xmul(t.Cq, t.D);
  t.v = xmul(xsub(1.0f, xdiv(t.n, t.d)), 0.5f);   // (1 - n/d) / 2, /2 is exact as *0.5
  return t;
}
// Two predictions at once (halves of each pair) against one target: same operations as ssim_terms,
// the target-only products are passed in (mu_y*mu_y and sig_y are shared by every candidate).
__device__ __forceinline__ void ssim_terms2(pk2 mu_x, float mu_y, float mu_yy, float sig_y, pk2 exx, pk2 exy,
                                            SsimTerms& t0, SsimTerms& t1) {
  const pk2 my = dup2(mu_y), two = dup2(2.0f), c1 = dup2(MAL_C1), c2 = dup2(MAL_C2);
  const pk2 mxx = x2mul_nf(mu_x, mu_x);
  const pk2 sig_x = x2sub(exx, mxx);
  const pk2 sig_xy = x2sub(exy, x2mul_nf(mu_x, my));
  const pk2 A = x2add(x2mul_nf(x2mul(two, mu_x), my), c1);
  const pk2 Bq = x2add(x2mul_nf(two, sig_xy), c2);
  const pk2 Cq = x2add(x2add(mxx, dup2(mu_yy)), c1);
  const pk2 D = x2add(x2add(sig_x, dup2(sig_y)), c2);
  const pk2 n = x2mul(A, Bq), d = x2mul(Cq, D);
  unpack2(mu_x, t0.mu_x, t1.mu_x);
  t0.mu_y = t1.mu_y = mu_y;
  unpack2(A, t0.A, t1.A); unpack2(Bq, t0.Bq, t1.Bq); unpack2(Cq, t0.Cq, t1.Cq); unpack2(D, t0.D, t1.D);
  unpack2(n, t0.n, t1.n); unpack2(d, t0.d, t1.d);
  t0.v = xmul(xsub(1.0f, xdiv(t0.n, t0.d)), 0.5f);
  t1.v = xmul(xsub(1.0f, xdiv(t1.n, t1.d)), 0.5f);
}
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }

// d clamp(v)/d{mu_x, E[xx], E[xy]} (the y-side moments carry no gradient: the target is data)
__device__ __forceinline__ void ssim_coefs(const SsimTerms& t, float& alpha, float& beta, float& gamma) {
  if (!(t.v >= 0.0f && t.v <= 1.0f)) { alpha = beta = gamma = 0.0f; return; }
  float inv_d = 1.0f / t.d;
  float r = t.n * inv_d;
  alpha = -(t.mu_y * (t.Bq - t.A) - r * t.mu_x * (t.D - t.Cq)) * inv_d;
  beta = 0.5f * r * t.Cq * inv_d;
  gamma = -t.A * inv_d;
}

}  // namespace mal
