// pointwise.cu - the per-pixel MAL student terms and the matching mask for sm_100a.
//
//   mal_main_terms_forward : compute_main_losses, manydepth/loss_utils.py:192-254 (per-pixel part)
//   mal_matching_mask      : Trainer.compute_matching_mask, manydepth/trainer.py:1066-1076, fused with
//                            the nearest up-sampling at manydepth/networks/repdepth.py:331-336 and the
//                            product at trainer.py:592-593
//
// Both are streaming kernels: every plane is read once with 128-bit loads when the row length
// allows it, nothing is re-read, results are written once.  Selection arithmetic uses the exact
// x* helpers (mal_common.cuh) because the arg-min / boolean outputs must match bit for bit.
#include "mal_math.cuh"

namespace mal {

constexpr int PW_NT = 256;

__device__ __forceinline__ float disp_to_depth_exact(float disp, float min_disp, float range) {
  return xdiv(1.0f, xadd(min_disp, xmul(range, disp)));   // manydepth/layers.py:19-22
}

// deterministic CTA reduction of two values into partials[blk*2 + {0,1}]
__device__ __forceinline__ void block_reduce2(float v0, float v1, float* red /*[2*NT/32]*/, float* out2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v0 = warp_sum(v0);
  v1 = warp_sum(v1);
  if (lane == 0) { red[warp * 2] = v0; red[warp * 2 + 1] = v1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.0f;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) s += red[wv * 2 + threadIdx.x];
    out2[threadIdx.x] = s;
  }
}

struct MainPix { float cons, distil, g_cons, g_distil, g_mono, target; int idx; };

__device__ __forceinline__ MainPix main_terms_pixel(const mal_main_terms_args& a, float multi_in, float mono_in,
                                                    float pm, float sm, float r_mono, float r_ens, float r_multi,
                                                    bool has_ens, float min_disp, float range, float inv_n) {
  MainPix o;
  float multi = multi_in, mono = mono_in;
  if (a.inputs_are_disp) {
    multi = disp_to_depth_exact(multi_in, min_disp, range);
    mono = disp_to_depth_exact(mono_in, min_disp, range);
  }
  // mask = ones * consistency_mask * (1 - augmentation_mask);  consistency_mask = 1 - mask  (:192-196)
  float m = xmul(xmul(1.0f, pm), xsub(1.0f, sm));
  float cm = xsub(1.0f, m);
  float dc = xsub(multi, mono);
  o.cons = xmul(fabsf(dc), cm);                                              // :205-206
  o.target = xdiv(1.0f, xadd(xmul(mono, cm), xmul(multi, xsub(1.0f, cm))));  // :211-213
  // arg-min, first index wins ties (:228-229 / :237-238)
  int idx;
  float distil;
  if (!has_ens) {
    idx = (r_multi < r_mono) ? 1 : 0;
    distil = idx == 0 ? mono : multi;
  } else {
    idx = 0;
    float best = r_mono;
    if (r_ens < best) { best = r_ens; idx = 1; }
    if (r_multi < best) { idx = 2; }
    float ens = xmul(xadd(mono, multi), 0.5f);   // (mono + multi) / 2.0
    distil = idx == 0 ? mono : (idx == 2 ? multi : ens);
  }
  o.idx = idx;
  float wd = xsub(1.0f, cm);   // (1 - consistency_mask), :253
  float dd = xsub(distil, multi);
  o.distil = xmul(fabsf(dd), wd);
  o.g_cons = o.g_distil = o.g_mono = 0.0f;
  if (a.with_grad) {
    float sc = dc > 0.f ? 1.f : (dc < 0.f ? -1.f : 0.f);
    o.g_cons = sc * cm * inv_n;
    float sd = dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f);
    // d(distil - multi)/d multi: mono -> -1, ensemble -> -1/2, multi -> 0
    float k = has_ens ? (idx == 0 ? -1.f : (idx == 1 ? -0.5f : 0.f)) : (idx == 0 ? -1.f : 0.f);
    o.g_distil = sd * k * wd * inv_n;
    if (a.dual_distil && !has_ens && idx == 0) o.g_mono = sd * wd * inv_n;
    if (a.inputs_are_disp) {
      float jm = -range * multi * multi, jo = -range * mono * mono;   // d depth / d disp
      o.g_cons *= jm; o.g_distil *= jm; o.g_mono *= jo;
    }
  }
  return o;
}

template <int VEC>
__global__ void __launch_bounds__(PW_NT) main_terms_kernel(const mal_main_terms_args a, const float min_disp,
                                                          const float range, const float inv_n) {
  __shared__ float red[2 * PW_NT / 32];
  const size_t HW = (size_t)a.height * a.width;
  const size_t total = (size_t)a.batch * HW / VEC;
  const bool has_ens = a.ens_reproj != nullptr;
  float acc_c = 0.0f, acc_d = 0.0f;
  for (size_t i = (size_t)blockIdx.x * PW_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * PW_NT) {
    const size_t e0 = i * VEC;
    const int b = (int)(e0 / HW);
    const float sm = a.sample_mask ? __ldg(a.sample_mask + b) : 0.0f;
    float multi[VEC], mono[VEC], pm[VEC], r0[VEC], r1[VEC], r2[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(multi) = *reinterpret_cast<const float4*>(a.multi + e0);
      *reinterpret_cast<float4*>(mono) = *reinterpret_cast<const float4*>(a.mono + e0);
      *reinterpret_cast<float4*>(pm) = *reinterpret_cast<const float4*>(a.pixel_mask + e0);
      *reinterpret_cast<float4*>(r0) = *reinterpret_cast<const float4*>(a.mono_reproj + e0);
      *reinterpret_cast<float4*>(r2) = *reinterpret_cast<const float4*>(a.multi_reproj + e0);
      if (has_ens) *reinterpret_cast<float4*>(r1) = *reinterpret_cast<const float4*>(a.ens_reproj + e0);
    } else {
      multi[0] = a.multi[e0]; mono[0] = a.mono[e0]; pm[0] = a.pixel_mask[e0];
      r0[0] = a.mono_reproj[e0]; r2[0] = a.multi_reproj[e0];
      if (has_ens) r1[0] = a.ens_reproj[e0];
    }
    float gc[VEC], gd[VEC], gm[VEC], tg[VEC];
    unsigned char ix[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
      MainPix o = main_terms_pixel(a, multi[v], mono[v], pm[v], sm, r0[v], has_ens ? r1[v] : 0.0f, r2[v], has_ens,
                                   min_disp, range, inv_n);
      acc_c += o.cons; acc_d += o.distil;
      gc[v] = o.g_cons; gd[v] = o.g_distil; gm[v] = o.g_mono; tg[v] = o.target; ix[v] = (unsigned char)o.idx;
    }
    if (VEC == 4) {
      if (a.distil_index) *reinterpret_cast<uchar4*>(a.distil_index + e0) = *reinterpret_cast<uchar4*>(ix);
      if (a.consistency_target) *reinterpret_cast<float4*>(a.consistency_target + e0) = *reinterpret_cast<float4*>(tg);
      if (a.with_grad) {
        *reinterpret_cast<float4*>(a.grad_cons + e0) = *reinterpret_cast<float4*>(gc);
        *reinterpret_cast<float4*>(a.grad_distil + e0) = *reinterpret_cast<float4*>(gd);
        if (a.grad_distil_mono) *reinterpret_cast<float4*>(a.grad_distil_mono + e0) = *reinterpret_cast<float4*>(gm);
      }
    } else {
      if (a.distil_index) a.distil_index[e0] = ix[0];
      if (a.consistency_target) a.consistency_target[e0] = tg[0];
      if (a.with_grad) {
        a.grad_cons[e0] = gc[0]; a.grad_distil[e0] = gd[0];
        if (a.grad_distil_mono) a.grad_distil_mono[e0] = gm[0];
      }
    }
  }
  block_reduce2(acc_c, acc_d, red, a.partials + (size_t)blockIdx.x * 2);
  // the last CTA to finish (ticket, zeroed by a memset node ahead of the launch) adds the per-CTA partials
  // in a fixed order: sums[j] = (sum over CTAs of partials[.][j]) * inv_n   (the two .mean() calls)
  __shared__ int s_last;
  __shared__ double dred[2 * PW_NT / 32];
  const int nblk = gridDim.x;
  unsigned* ticket = reinterpret_cast<unsigned*>(a.partials + (size_t)nblk * 2);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == (unsigned)(nblk - 1)) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const volatile float* part = a.partials;
  double s0 = 0.0, s1 = 0.0;
  for (int i = threadIdx.x; i < nblk; i += PW_NT) { s0 += (double)part[i * 2]; s1 += (double)part[i * 2 + 1]; }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { dred[warp * 2] = s0; dred[warp * 2 + 1] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int wv = 0; wv < PW_NT / 32; wv++) s += dred[wv * 2 + threadIdx.x];
    a.sums[threadIdx.x] = (float)(s * (double)inv_n);
  }
}

inline int main_terms_blocks(size_t work) {
  size_t blk = (work + PW_NT - 1) / PW_NT;
  const size_t cap = 148 * 8;   // a few CTAs per SM, grid-stride beyond
  return (int)(blk < cap ? (blk ? blk : 1) : cap);
}

// ---- matching mask ----------------------------------------------------------------------------
__global__ void __launch_bounds__(PW_NT) matching_mask_kernel(const mal_matching_mask_args a, const float min_disp,
                                                             const float range) {
  const int H = a.height, W = a.width, h = a.low_height, w = a.low_width;
  const size_t HW = (size_t)H * W, total = (size_t)a.batch * HW;
  // F.interpolate(mode="nearest"): src = floor(dst * (in / out)), computed in fp32 like ATen
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (size_t i = (size_t)blockIdx.x * PW_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * PW_NT) {
    const int b = (int)(i / HW);
    const int p = (int)(i - (size_t)b * HW);
    const int y = p / W, x = p - y * W;
    const int ly = min((int)floorf((float)y * sy), h - 1), lx = min((int)floorf((float)x * sx), w - 1);
    const size_t lo = ((size_t)b * h + ly) * w + lx;
    float t = __ldg(a.mono + i);
    if (a.mono_is_disp) t = disp_to_depth_exact(t, min_disp, range);
    const float m = xdiv(1.0f, __ldg(a.lowest_cost + lo));
    const bool ok = (xdiv(xsub(m, t), t) < 1.0f) && (xdiv(xsub(t, m), m) < 1.0f);
    float v = ok ? 1.0f : 0.0f;
    if (a.confidence) v = xmul(__ldg(a.confidence + lo), v);
    a.out_mask[i] = v;
  }
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_main_terms_partials_floats(int batch, int height, int width) {
  return (size_t)main_terms_blocks((size_t)batch * height * width) * 2 + 4;   // + the ticket
}

extern "C" int mal_main_terms_forward(const mal_main_terms_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_main_terms_forward: args is NULL");
  const mal_main_terms_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height > 0 && a.width > 0, "mal_main_terms_forward: bad shape");
  MAL_REQUIRE(a.multi && a.mono && a.pixel_mask && a.mono_reproj && a.multi_reproj && a.partials && a.sums,
              "mal_main_terms_forward: multi/mono/pixel_mask/mono_reproj/multi_reproj/partials/sums are required");
  if (a.with_grad) MAL_REQUIRE(a.grad_cons && a.grad_distil, "mal_main_terms_forward: with_grad needs grad_cons/grad_distil");
  if (a.with_grad && a.dual_distil)
    MAL_REQUIRE(a.grad_distil_mono, "mal_main_terms_forward: dual_distil needs grad_distil_mono");
  if (a.inputs_are_disp) MAL_REQUIRE(a.min_depth > 0 && a.max_depth > a.min_depth, "mal_main_terms_forward: bad depth range");
  const double lo = 1.0 / a.max_depth, hi = 1.0 / a.min_depth;
  const float min_disp = (float)lo, range = (float)(hi - lo);
  const size_t n = (size_t)a.batch * a.height * a.width;
  const float inv_n = (float)(1.0 / (double)n);
  cudaStream_t st = (cudaStream_t)stream;
  auto aligned = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = (((size_t)a.height * a.width) % 4 == 0) && aligned(a.multi) && aligned(a.mono) &&
                   aligned(a.pixel_mask) && aligned(a.mono_reproj) && aligned(a.multi_reproj) &&
                   aligned(a.ens_reproj) && aligned(a.consistency_target) && aligned(a.grad_cons) &&
                   aligned(a.grad_distil) && aligned(a.grad_distil_mono) && (((uintptr_t)a.distil_index & 3) == 0);
  const int nblk = main_terms_blocks(n);
  cudaMemsetAsync(a.partials + (size_t)nblk * 2, 0, sizeof(unsigned), st);
  if (vec) launch(main_terms_kernel<4>, dim3(nblk), dim3(PW_NT), 0, st, a, min_disp, range, inv_n);
  else launch(main_terms_kernel<1>, dim3(nblk), dim3(PW_NT), 0, st, a, min_disp, range, inv_n);
  return check_launch("main_terms_kernel");
}

extern "C" int mal_matching_mask(const mal_matching_mask_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_matching_mask: args is NULL");
  const mal_matching_mask_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height > 0 && a.width > 0 && a.low_height > 0 && a.low_width > 0,
              "mal_matching_mask: bad shape");
  MAL_REQUIRE(a.lowest_cost && a.mono && a.out_mask, "mal_matching_mask: lowest_cost/mono/out_mask are required");
  if (a.mono_is_disp) MAL_REQUIRE(a.min_depth > 0 && a.max_depth > a.min_depth, "mal_matching_mask: bad depth range");
  const double lo = 1.0 / a.max_depth, hi = 1.0 / a.min_depth;
  const size_t n = (size_t)a.batch * a.height * a.width;
  launch(matching_mask_kernel, dim3(main_terms_blocks(n)), dim3(PW_NT), 0, (cudaStream_t)stream, a, (float)lo,
         (float)(hi - lo));
  return check_launch("matching_mask_kernel");
}
