// corr.cu - DualRefine's epipolar correlation lookup, forward and backward, sm_100a.
//
// Replaces CoordSampler.register / __call__ / __corr__ (dualrefine/networks/corr.py:11-76): inside every
// DEQ iteration the reference samples the matching features of the source frame (and their 2x2
// average-pooled pyramid) at L x D epipolar candidates per pixel with F.grid_sample(align_corners=False,
// zeros padding), forms |fmap1 - sample| as a (B, C, h, w, D) tensor per level and averages it over the
// channels of each head.  Here one thread owns one (level, candidate, pixel) and walks the channels, so
// nothing larger than the inputs and the (B, L*heads*D, h, w) result exists.
//
// Arithmetic contract (bit-exact against torch CPU, pinned by tests/golden/corr_lookup.npz):
//   grid             ((x + 0.5) * 2) / w1 - 1                       corr.py:34-35
//   sampler          unnormalize<DUALREFINE> + nw*a then 3 FMAs     (mal_math.cuh; ATen GridSamplerKernel)
//   pyramid          ((a + b) + c + d) / 4 row-major                ATen AvgPoolKernel, corr.py:19-21
//   mean over Cg     ATen SumKernel.cpp: for an output whose flat inner index j (over h*w*D) lies in the
//                    32-column vectorised body, cascade_sum: 16 channels sequentially, chunk sums added
//                    sequentially; for the tail columns (j >= floor(N/32)*32) row_sum: four interleaved
//                    partial sums (channel % 4), each cascaded in 16s, then ((p0+p1)+p2)+p3; then / Cg.
//
// Backward (torch.abs -> sign, grid_sample's zero-padded bilinear adjoint): d/d coords is gathered per
// thread; d/d fmap1 and d/d pyramid are scatter-adds (red.global.add.f32, fire-and-forget in L2) - the
// same atomics ATen's grid_sampler_2d_backward issues, without the (B, C, h, w, D) intermediates.
#include "mal_math.cuh"

namespace mal {

constexpr int CR_NT = 128;

__host__ __device__ inline size_t corr_level_offset(int batch, int channels, int h, int w, int level) {
  size_t off = 0;
  for (int l = 0; l < level; l++) { off += (size_t)batch * channels * h * w; h /= 2; w /= 2; }
  return off;
}

// level l+1 = F.avg_pool2d(level l, 2, stride=2); one thread per output element
__global__ void __launch_bounds__(256) corr_pool_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        size_t planes, int ih, int iw) {
  const int oh = ih / 2, ow = iw / 2;
  const size_t n = planes * oh * ow;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
  const size_t pl = i / ((size_t)ow * oh);
  const float* p = in + pl * ih * iw + (size_t)(2 * oy) * iw + 2 * ox;
  const float s = xadd(xadd(xadd(__ldg(p), __ldg(p + 1)), __ldg(p + iw)), __ldg(p + iw + 1));
  out[i] = xdiv(s, 4.0f);
}

struct CorrSite {
  int b, l, d, y, x;
  int lh, lw;            // level resolution
  size_t out_index;      // into (B, L*heads*D, h, w) for head 0
  bool tail;             // ATen's remainder columns: interleaved partial sums
};

__device__ __forceinline__ bool corr_site(const mal_corr_args& a, size_t i, CorrSite& s) {
  const int h = a.height, w = a.width, D = a.num_samples, L = a.num_levels;
  const size_t total = (size_t)a.batch * L * D * h * w;
  if (i >= total) return false;
  s.x = (int)(i % w);
  size_t r = i / w;
  s.y = (int)(r % h); r /= h;
  s.d = (int)(r % D); r /= D;
  s.l = (int)(r % L);
  s.b = (int)(r / L);
  s.lh = h >> s.l; s.lw = w >> s.l;
  s.out_index = (((size_t)s.b * L + s.l) * a.num_head * D + s.d) * h * w + (size_t)s.y * w + s.x;
  // the reference reduces a (B, heads, Cg, h, w, D) view over Cg: inner index over (h, w, D)
  const size_t N = (size_t)h * w * D, j = ((size_t)s.y * w + s.x) * D + s.d;
  s.tail = j >= N / 32 * 32;
  return true;
}

__device__ __forceinline__ Taps corr_taps(const mal_corr_args& a, const CorrSite& s, float cx, float cy) {
  // corr.py:34-35, then grid_sample(align_corners=False) at the level's own resolution
  const float gx = xsub(xdiv(xmul(2.0f, xadd(cx, 0.5f)), (float)a.width), 1.0f);
  const float gy = xsub(xdiv(xmul(2.0f, xadd(cy, 0.5f)), (float)a.height), 1.0f);
  const float ix = unnormalize<MAL_CONV_DUALREFINE>(gx, s.lw), iy = unnormalize<MAL_CONV_DUALREFINE>(gy, s.lh);
  // keep float -> int conversion defined for wild coordinates: every tap is out of range anyway
  const float lim = 1.0e6f;
  return make_taps(fminf(fmaxf(ix, -lim), lim), fminf(fmaxf(iy, -lim), lim), s.lh, s.lw);
}

__global__ void __launch_bounds__(CR_NT) corr_lookup_kernel(const mal_corr_args a) {
  CorrSite s;
  if (!corr_site(a, (size_t)blockIdx.x * CR_NT + threadIdx.x, s)) return;
  const int h = a.height, w = a.width, C = a.channels, Cg = C / a.num_head;
  const size_t hw = (size_t)h * w, lhw = (size_t)s.lh * s.lw;
  const size_t cbase = (((size_t)s.b * 2) * a.num_levels + s.l) * a.num_samples + s.d;
  const float cx = __ldg(a.coords + cbase * hw + (size_t)s.y * w + s.x);
  const float cy = __ldg(a.coords + (cbase + (size_t)a.num_levels * a.num_samples) * hw + (size_t)s.y * w + s.x);
  const Taps t = corr_taps(a, s, cx, cy);
  const float* f1 = a.fmap1 + (size_t)s.b * C * hw + (size_t)s.y * w + s.x;
  const float* f2 = a.pyramid + corr_level_offset(a.batch, C, h, w, s.l) + (size_t)s.b * C * lhw;
  for (int hd = 0; hd < a.num_head; hd++) {
    float total = 0.0f;   // acc[1] of the cascade
    if (!s.tail) {
      for (int c0 = 0; c0 < Cg; c0 += 16) {
        float acc = 0.0f;
        const int c1 = min(Cg, c0 + 16);
        for (int c = c0; c < c1; c++) {
          const int ch = hd * Cg + c;
          acc = xadd(acc, fabsf(xsub(__ldg(f1 + ch * hw), bilinear(f2 + ch * lhw, t))));
        }
        // a trailing partial chunk stays in acc[0] and is added last: same sequence
        total = xadd(total, acc);
      }
    } else {
      float part[4] = {0.f, 0.f, 0.f, 0.f}, acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int rows = Cg / 4;
      for (int r = 0; r < rows; r++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int ch = hd * Cg + r * 4 + k;
          acc[k] = xadd(acc[k], fabsf(xsub(__ldg(f1 + ch * hw), bilinear(f2 + ch * lhw, t))));
        }
        if ((r & 15) == 15) {
#pragma unroll
          for (int k = 0; k < 4; k++) { part[k] = xadd(part[k], acc[k]); acc[k] = 0.0f; }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) part[k] = xadd(acc[k], part[k]);   // acc[0] += acc[1]
      for (int c = rows * 4; c < Cg; c++) {                          // leftover channels go to partial 0
        const int ch = hd * Cg + c;
        part[0] = xadd(part[0], fabsf(xsub(__ldg(f1 + ch * hw), bilinear(f2 + ch * lhw, t))));
      }
      total = xadd(xadd(xadd(part[0], part[1]), part[2]), part[3]);
    }
    a.out[s.out_index + (size_t)hd * a.num_samples * hw] = xdiv(total, (float)Cg);
  }
}

__device__ __forceinline__ void red_add(float* p, float v) {
#ifdef MAL_EMU
  *p += v;
#else
  atomicAdd(p, v);
#endif
}

__global__ void __launch_bounds__(CR_NT) corr_lookup_bwd_kernel(const mal_corr_args a) {
  CorrSite s;
  if (!corr_site(a, (size_t)blockIdx.x * CR_NT + threadIdx.x, s)) return;
  const int h = a.height, w = a.width, C = a.channels, Cg = C / a.num_head;
  const size_t hw = (size_t)h * w, lhw = (size_t)s.lh * s.lw;
  const size_t cbase = (((size_t)s.b * 2) * a.num_levels + s.l) * a.num_samples + s.d;
  const size_t pix = (size_t)s.y * w + s.x;
  const float cx = __ldg(a.coords + cbase * hw + pix);
  const float cy = __ldg(a.coords + (cbase + (size_t)a.num_levels * a.num_samples) * hw + pix);
  const Taps t = corr_taps(a, s, cx, cy);
  const size_t loff = corr_level_offset(a.batch, C, h, w, s.l) + (size_t)s.b * C * lhw;
  const float* f1 = a.fmap1 + (size_t)s.b * C * hw + pix;
  const float* f2 = a.pyramid + loff;
  float gix = 0.0f, giy = 0.0f;
  for (int hd = 0; hd < a.num_head; hd++) {
    const float g = __ldg(a.grad_out + s.out_index + (size_t)hd * a.num_samples * hw) / (float)Cg;
    if (g == 0.0f) continue;
    for (int c = 0; c < Cg; c++) {
      const int ch = hd * Cg + c;
      float v00, v01, v10, v11;
      const float sv = bilinear(f2 + ch * lhw, t, &v00, &v01, &v10, &v11);
      const float df = __ldg(f1 + ch * hw) - sv;
      const float sg = df > 0.0f ? g : (df < 0.0f ? -g : 0.0f);   // d|f1 - s| / d f1
      if (sg == 0.0f) continue;
      if (a.grad_fmap1) red_add(a.grad_fmap1 + ((size_t)s.b * C + ch) * hw + pix, sg);
      if (a.grad_pyramid) {
        float* gp = a.grad_pyramid + loff + ch * lhw;
        if (t.v00) red_add(gp + t.o00, -sg * t.nw);
        if (t.v01) red_add(gp + t.o01, -sg * t.ne);
        if (t.v10) red_add(gp + t.o10, -sg * t.sw);
        if (t.v11) red_add(gp + t.o11, -sg * t.se);
      }
      gix -= sg * ((v01 - v00) * (1.0f - t.ty) + (v11 - v10) * t.ty);
      giy -= sg * ((v10 - v00) * (1.0f - t.tx) + (v11 - v01) * t.tx);
    }
  }
  if (a.grad_coords) {
    // ix = ((2 (x + 0.5) / w1 - 1) + 1) * lw / 2 - 0.5  =>  d ix / d x = lw / w1
    a.grad_coords[cbase * hw + pix] = gix * ((float)s.lw / (float)w);
    a.grad_coords[(cbase + (size_t)a.num_levels * a.num_samples) * hw + pix] = giy * ((float)s.lh / (float)h);
  }
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_corr_pyramid_floats(int batch, int channels, int height, int width, int num_levels) {
  return corr_level_offset(batch, channels, height, width, num_levels);
}

extern "C" int mal_corr_pyramid(const float* fmap2, int batch, int channels, int height, int width, int num_levels,
                                float* pyramid, mal_stream_t stream) {
  MAL_REQUIRE(fmap2 && pyramid && batch > 0 && channels > 0 && height > 0 && width > 0 && num_levels > 0,
              "mal_corr_pyramid: bad arguments");
  MAL_REQUIRE((height >> (num_levels - 1)) > 0 && (width >> (num_levels - 1)) > 0,
              "mal_corr_pyramid: %d levels do not fit a %dx%d map", num_levels, height, width);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t planes = (size_t)batch * channels;
#ifdef MAL_EMU
  memcpy(pyramid, fmap2, planes * height * width * sizeof(float));
#else
  cudaError_t e = cudaMemcpyAsync(pyramid, fmap2, planes * height * width * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return fail(MAL_ERR_LAUNCH, "mal_corr_pyramid: %s", cudaGetErrorString(e));
#endif
  int h = height, w = width;
  for (int l = 1; l < num_levels; l++) {
    const float* in = pyramid + corr_level_offset(batch, channels, height, width, l - 1);
    float* out = pyramid + corr_level_offset(batch, channels, height, width, l);
    const size_t n = planes * (h / 2) * (w / 2);
    launch(corr_pool_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, in, out, planes, h, w);
    h /= 2; w /= 2;
  }
  return check_launch("corr_pool_kernel");
}

static int corr_check(const mal_corr_args& a, const char* who) {
  MAL_REQUIRE(a.batch > 0 && a.channels > 0 && a.height > 0 && a.width > 0 && a.num_levels > 0 && a.num_samples > 0 &&
                  a.num_head > 0,
              "%s: bad shape", who);
  MAL_REQUIRE(a.channels % a.num_head == 0, "%s: %d channels do not split into %d heads", who, a.channels, a.num_head);
  MAL_REQUIRE((a.height >> (a.num_levels - 1)) > 0 && (a.width >> (a.num_levels - 1)) > 0, "%s: too many levels", who);
  MAL_REQUIRE(a.fmap1 && a.pyramid && a.coords, "%s: fmap1 / pyramid / coords are required", who);
  return MAL_OK;
}

extern "C" int mal_corr_lookup(const mal_corr_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_corr_lookup: args is NULL");
  const mal_corr_args& a = *args;
  int rc = corr_check(a, "mal_corr_lookup");
  if (rc) return rc;
  MAL_REQUIRE(a.out, "mal_corr_lookup: out is required");
  const size_t total = (size_t)a.batch * a.num_levels * a.num_samples * a.height * a.width;
  launch(corr_lookup_kernel, dim3((unsigned)((total + CR_NT - 1) / CR_NT)), dim3(CR_NT), 0, (cudaStream_t)stream, a);
  return check_launch("corr_lookup_kernel");
}

extern "C" int mal_corr_lookup_backward(const mal_corr_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_corr_lookup_backward: args is NULL");
  const mal_corr_args& a = *args;
  int rc = corr_check(a, "mal_corr_lookup_backward");
  if (rc) return rc;
  MAL_REQUIRE(a.grad_out && (a.grad_coords || a.grad_fmap1 || a.grad_pyramid),
              "mal_corr_lookup_backward: grad_out and at least one gradient output are required");
  const size_t total = (size_t)a.batch * a.num_levels * a.num_samples * a.height * a.width;
  launch(corr_lookup_bwd_kernel, dim3((unsigned)((total + CR_NT - 1) / CR_NT)), dim3(CR_NT), 0, (cudaStream_t)stream, a);
  return check_launch("corr_lookup_bwd_kernel");
}
