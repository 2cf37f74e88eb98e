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
// Layout: the pyramid buffer holds every level channel-quad interleaved, [B][C/4][h>>l][w>>l] float4, so
// one 128-bit load fetches four channels of a bilinear tap and the blend runs on packed fp32 pairs
// (FFMA2, mal_common.cuh); fmap1 stays NCHW (its loads are coalesced across the pixels of a warp).
//
// Backward (torch.abs -> sign, grid_sample's zero-padded bilinear adjoint): d/d coords is gathered per
// thread; d/d fmap1 and d/d pyramid are scatter-adds (red.global.add, fire-and-forget in L2; 128-bit
// vector reductions into the quad-interleaved pyramid gradient: one per tap and channel quad) - the same
// atomics ATen's grid_sampler_2d_backward issues, a quarter as many, without the (B, C, h, w, D)
// intermediates.
#include "mal_math.cuh"

namespace mal {

constexpr int CR_NT = 128;

__device__ __forceinline__ float4 ldg4c(const float4* p) {
#ifdef MAL_EMU
  return *p;
#else
  return __ldg(p);
#endif
}

__host__ __device__ inline size_t corr_level_offset(int batch, int channels, int h, int w, int level) {
  size_t off = 0;
  for (int l = 0; l < level; l++) { off += (size_t)batch * channels * h * w; h /= 2; w /= 2; }
  return off;
}

// level 0: NCHW -> [C/4][h][w] float4; one thread per (quad, pixel)
__global__ void __launch_bounds__(256) corr_pack_kernel(const float* __restrict__ src, float4* __restrict__ dst,
                                                        int C, int hw, size_t total) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int p = (int)(i % hw);
  const size_t r = i / hw;
  const int q = (int)(r % (C / 4));
  const size_t img = r / (C / 4);
  const float* sp = src + (img * C + (size_t)q * 4) * hw + p;
  dst[i] = make_float4(__ldg(sp), __ldg(sp + hw), __ldg(sp + 2 * (size_t)hw), __ldg(sp + 3 * (size_t)hw));
}

__device__ __forceinline__ float pool4(float a, float b, float c, float d) {
  return xdiv(xadd(xadd(xadd(a, b), c), d), 4.0f);   // ATen avg_pool2d: row-major running sum / 4
}

// level l+1 = F.avg_pool2d(level l, 2, stride=2) on the packed planes; one thread per output float4
__global__ void __launch_bounds__(256) corr_pool_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                        size_t planes, int ih, int iw) {
  const int oh = ih / 2, ow = iw / 2;
  const size_t n = planes * oh * ow;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
  const size_t pl = i / ((size_t)ow * oh);
  const float4* p = in + pl * ih * iw + (size_t)(2 * oy) * iw + 2 * ox;
  const float4 a = p[0], b = p[1], c = p[iw], d = p[iw + 1];
  out[i] = make_float4(pool4(a.x, b.x, c.x, d.x), pool4(a.y, b.y, c.y, d.y), pool4(a.z, b.z, c.z, d.z),
                       pool4(a.w, b.w, c.w, d.w));
}

struct CorrSite {
  int b, l, d, y, x;
  int lh, lw;            // level resolution
  size_t out_index;      // into (B, L*heads*D, h, w) for head 0
  bool tail;             // ATen's remainder columns: interleaved partial sums
};

// Thread -> work item: x fastest (coalesced coordinate reads and result writes), then y, candidate, level,
// sample.  (Putting the candidate index next to the 32-pixel segment, so that the warps of a CTA share
// texel lines, was measured slower: 365 -> 440 us forward.)
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

__host__ __device__ inline size_t corr_threads(const mal_corr_args& a) {
  return (size_t)a.batch * a.num_levels * a.num_samples * a.height * a.width;
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

// the four channels of one quad at a sampling point: zeros padding, nw*a then three FMAs per channel,
// two channels per instruction
struct Quad { pk2 lo, hi; };
__device__ __forceinline__ float4 tap4(const float4* __restrict__ plane, bool ok, int o) {
  return ok ? ldg4c(plane + o) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ Quad blend4(const float4* __restrict__ plane, const Taps& t, pk2 nw, pk2 ne, pk2 sw, pk2 se) {
  const float4 a = tap4(plane, t.v00, t.o00), b = tap4(plane, t.v01, t.o01);
  const float4 c = tap4(plane, t.v10, t.o10), d = tap4(plane, t.v11, t.o11);
  Quad q;
  q.lo = x2fma(pack2(d.x, d.y), se, x2fma(pack2(c.x, c.y), sw, x2fma(pack2(b.x, b.y), ne, x2mul(pack2(a.x, a.y), nw))));
  q.hi = x2fma(pack2(d.z, d.w), se, x2fma(pack2(c.z, c.w), sw, x2fma(pack2(b.z, b.w), ne, x2mul(pack2(a.z, a.w), nw))));
  return q;
}

__global__ void __launch_bounds__(CR_NT) corr_lookup_kernel(const mal_corr_args a) {
  CorrSite s;
  if (!corr_site(a, (size_t)blockIdx.x * CR_NT + threadIdx.x, s)) return;
  const int h = a.height, w = a.width, C = a.channels, Cg = C / a.num_head, Qg = Cg / 4;
  const size_t hw = (size_t)h * w, lhw = (size_t)s.lh * s.lw;
  const size_t cbase = (((size_t)s.b * 2) * a.num_levels + s.l) * a.num_samples + s.d;
  const float cx = __ldg(a.coords + cbase * hw + (size_t)s.y * w + s.x);
  const float cy = __ldg(a.coords + (cbase + (size_t)a.num_levels * a.num_samples) * hw + (size_t)s.y * w + s.x);
  const Taps t = corr_taps(a, s, cx, cy);
  const pk2 nw = dup2(t.nw), ne = dup2(t.ne), sw = dup2(t.sw), se = dup2(t.se);
  const float* f1 = a.fmap1 + (size_t)s.b * C * hw + (size_t)s.y * w + s.x;
  const float4* f2 = reinterpret_cast<const float4*>(a.pyramid + corr_level_offset(a.batch, C, h, w, s.l)) +
                     (size_t)s.b * (C / 4) * lhw;
  for (int hd = 0; hd < a.num_head; hd++) {
    float total = 0.0f;   // acc[1] of ATen's cascade
    if (!s.tail) {
      float acc = 0.0f;
      int q = 0;
      // one 16-channel chunk at a time: all 16 tap loads and 16 feature loads are issued before the first
      // is consumed (the kernel is bound by load latency, not by issue slots)
      for (; q + 4 <= Qg; q += 4) {
        float4 ta[4], tb[4], tc[4], td[4];
        float fv[4][4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int qq = hd * Qg + q + j;
          const float4* plane = f2 + (size_t)qq * lhw;
          ta[j] = tap4(plane, t.v00, t.o00); tb[j] = tap4(plane, t.v01, t.o01);
          tc[j] = tap4(plane, t.v10, t.o10); td[j] = tap4(plane, t.v11, t.o11);
          const float* f = f1 + (size_t)qq * 4 * hw;
#pragma unroll
          for (int e = 0; e < 4; e++) fv[j][e] = __ldg(f + e * hw);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const pk2 lo = x2fma(pack2(td[j].x, td[j].y), se, x2fma(pack2(tc[j].x, tc[j].y), sw, x2fma(pack2(tb[j].x, tb[j].y), ne, x2mul(pack2(ta[j].x, ta[j].y), nw))));
          const pk2 hi = x2fma(pack2(td[j].z, td[j].w), se, x2fma(pack2(tc[j].z, tc[j].w), sw, x2fma(pack2(tb[j].z, tb[j].w), ne, x2mul(pack2(ta[j].z, ta[j].w), nw))));
          const pk2 dlo = x2sub(pack2(fv[j][0], fv[j][1]), lo), dhi = x2sub(pack2(fv[j][2], fv[j][3]), hi);
          acc = xadd(acc, fabsf(lo2(dlo)));
          acc = xadd(acc, fabsf(hi2(dlo)));
          acc = xadd(acc, fabsf(lo2(dhi)));
          acc = xadd(acc, fabsf(hi2(dhi)));
        }
        total = xadd(total, acc);   // every 16 channels
        acc = 0.0f;
      }
      for (; q < Qg; q++) {   // channels beyond the last full chunk
        const int qq = hd * Qg + q;
        const Quad v = blend4(f2 + (size_t)qq * lhw, t, nw, ne, sw, se);
        const float* f = f1 + (size_t)qq * 4 * hw;
        const pk2 dlo = x2sub(pack2(__ldg(f), __ldg(f + hw)), v.lo);
        const pk2 dhi = x2sub(pack2(__ldg(f + 2 * hw), __ldg(f + 3 * hw)), v.hi);
        acc = xadd(acc, fabsf(lo2(dlo)));
        acc = xadd(acc, fabsf(hi2(dlo)));
        acc = xadd(acc, fabsf(lo2(dhi)));
        acc = xadd(acc, fabsf(hi2(dhi)));
      }
      total = xadd(acc, total);   // a trailing partial chunk (acc[0]) joins last
    } else {
      // four interleaved partial sums (channel % 4 == component), cascaded every 16 rows
      pk2 plo = dup2(0.0f), phi = dup2(0.0f), alo = dup2(0.0f), ahi = dup2(0.0f);
#pragma unroll 4
      for (int q = 0; q < Qg; q++) {
        const int qq = hd * Qg + q;
        const Quad v = blend4(f2 + (size_t)qq * lhw, t, nw, ne, sw, se);
        const float* f = f1 + (size_t)qq * 4 * hw;
        alo = x2add(alo, abs2(x2sub(pack2(__ldg(f), __ldg(f + hw)), v.lo)));
        ahi = x2add(ahi, abs2(x2sub(pack2(__ldg(f + 2 * hw), __ldg(f + 3 * hw)), v.hi)));
        if ((q & 15) == 15) { plo = x2add(plo, alo); phi = x2add(phi, ahi); alo = dup2(0.0f); ahi = dup2(0.0f); }
      }
      plo = x2add(alo, plo);
      phi = x2add(ahi, phi);
      total = xadd(xadd(xadd(lo2(plo), hi2(plo)), lo2(phi)), hi2(phi));
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

__device__ __forceinline__ void red_add4(float4* p, float x, float y, float z, float w) {
#ifdef MAL_EMU
  p->x += x; p->y += y; p->z += z; p->w += w;
#else
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
#endif
}

// (channel quads per load batch x resident CTAs were swept on the B200 - 1x6, 1x8, 2x6, 4x3, 4x4 -: all within 10 % of
// 2 quads per batch at the natural register count, profiles/r2_notes.md)
__global__ void __launch_bounds__(CR_NT) corr_lookup_bwd_kernel(const mal_corr_args a) {
  CorrSite s;
  if (!corr_site(a, (size_t)blockIdx.x * CR_NT + threadIdx.x, s)) return;
  const int h = a.height, w = a.width, C = a.channels, Cg = C / a.num_head, Qg = Cg / 4;
  const size_t hw = (size_t)h * w, lhw = (size_t)s.lh * s.lw;
  const size_t cbase = (((size_t)s.b * 2) * a.num_levels + s.l) * a.num_samples + s.d;
  const size_t pix = (size_t)s.y * w + s.x;
  const float cx = __ldg(a.coords + cbase * hw + pix);
  const float cy = __ldg(a.coords + (cbase + (size_t)a.num_levels * a.num_samples) * hw + pix);
  const Taps t = corr_taps(a, s, cx, cy);
  const size_t loff = corr_level_offset(a.batch, C, h, w, s.l) + (size_t)s.b * C * lhw;
  const float* f1 = a.fmap1 + (size_t)s.b * C * hw + pix;
  const float4* f2 = reinterpret_cast<const float4*>(a.pyramid + loff);
  float4* gp = a.grad_pyramid ? reinterpret_cast<float4*>(a.grad_pyramid + loff) : nullptr;
  // d/d coords: with s_k = d|f1 - sample| / d sample per channel,
  //   d/d ix = sum_k s_k ((b_k - a_k) (1 - ty) + (d_k - c_k) ty),  d/d iy = sum_k s_k ((c_k - a_k) (1 - tx) + (d_k - b_k) tx)
  // are formed from the four tap sums  sum_k s_k a_k ... sum_k s_k d_k  (one packed FMA per tap and channel pair
  // instead of eight scalar operations per channel: the coordinate-only backward went from 591 to ... us)
  const pk2 zero2 = pack2(0.0f, 0.0f);
  pk2 sa2 = zero2, sb2 = zero2, sc2 = zero2, sd2 = zero2;
  const pk2 nw2 = dup2(t.nw), ne2 = dup2(t.ne), sw2 = dup2(t.sw), se2 = dup2(t.se);
  for (int hd = 0; hd < a.num_head; hd++) {
    const float g = __ldg(a.grad_out + s.out_index + (size_t)hd * a.num_samples * hw) / (float)Cg;
    if (g == 0.0f) continue;
    // QB channel quads at a time: their tap and feature loads are all issued before the first is consumed
    constexpr int QB = 2;
    for (int q0 = 0; q0 < Qg; q0 += QB) {
      float4 ta[QB], tb[QB], tc[QB], td[QB];
      float fv[QB][4];
#pragma unroll
      for (int j = 0; j < QB; j++) {
        if (q0 + j < Qg) {
          const int qq = hd * Qg + q0 + j;
          const float4* plane = f2 + (size_t)qq * lhw;
          ta[j] = tap4(plane, t.v00, t.o00); tb[j] = tap4(plane, t.v01, t.o01);
          tc[j] = tap4(plane, t.v10, t.o10); td[j] = tap4(plane, t.v11, t.o11);
          const float* f = f1 + (size_t)qq * 4 * hw;
#pragma unroll
          for (int e = 0; e < 4; e++) fv[j][e] = __ldg(f + e * hw);
        }
      }
#pragma unroll
      for (int j = 0; j < QB; j++) {
        if (q0 + j < Qg) {
          const int qq = hd * Qg + q0 + j;
          float sg[4];
          const pk2 a2[2] = {pack2(ta[j].x, ta[j].y), pack2(ta[j].z, ta[j].w)};
          const pk2 b2[2] = {pack2(tb[j].x, tb[j].y), pack2(tb[j].z, tb[j].w)};
          const pk2 c2[2] = {pack2(tc[j].x, tc[j].y), pack2(tc[j].z, tc[j].w)};
          const pk2 d2[2] = {pack2(td[j].x, td[j].y), pack2(td[j].z, td[j].w)};
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            // the sample exactly as the forward rounds it: the sign below is the reference's
            const pk2 sv = x2fma(d2[hh], se2, x2fma(c2[hh], sw2, x2fma(b2[hh], ne2, x2mul(a2[hh], nw2))));
            const pk2 df = x2sub(pack2(fv[j][2 * hh], fv[j][2 * hh + 1]), sv);
            const float d0 = lo2(df), d1 = hi2(df);
            sg[2 * hh] = d0 > 0.0f ? g : (d0 < 0.0f ? -g : 0.0f);   // d|f1 - s| / d f1 = -d|f1 - s| / d s
            sg[2 * hh + 1] = d1 > 0.0f ? g : (d1 < 0.0f ? -g : 0.0f);
            const pk2 sg2 = pack2(sg[2 * hh], sg[2 * hh + 1]);
            sa2 = x2fma(sg2, a2[hh], sa2); sb2 = x2fma(sg2, b2[hh], sb2);
            sc2 = x2fma(sg2, c2[hh], sc2); sd2 = x2fma(sg2, d2[hh], sd2);
          }
          if (a.workspace) {   // channel-quad interleaved like the pyramid: one 128-bit reduction instead of four
            if (sg[0] != 0.0f || sg[1] != 0.0f || sg[2] != 0.0f || sg[3] != 0.0f)
              red_add4(reinterpret_cast<float4*>(a.workspace) + ((size_t)s.b * (C / 4) + qq) * hw + pix, sg[0], sg[1], sg[2], sg[3]);
          } else if (a.grad_fmap1) {
            float* g1 = a.grad_fmap1 + ((size_t)s.b * C + (size_t)qq * 4) * hw + pix;
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (sg[k] != 0.0f) red_add(g1 + k * hw, sg[k]);
          }
          if (gp && (sg[0] != 0.0f || sg[1] != 0.0f || sg[2] != 0.0f || sg[3] != 0.0f)) {
            float4* gq = gp + (size_t)qq * lhw;
            if (t.v00) red_add4(gq + t.o00, -sg[0] * t.nw, -sg[1] * t.nw, -sg[2] * t.nw, -sg[3] * t.nw);
            if (t.v01) red_add4(gq + t.o01, -sg[0] * t.ne, -sg[1] * t.ne, -sg[2] * t.ne, -sg[3] * t.ne);
            if (t.v10) red_add4(gq + t.o10, -sg[0] * t.sw, -sg[1] * t.sw, -sg[2] * t.sw, -sg[3] * t.sw);
            if (t.v11) red_add4(gq + t.o11, -sg[0] * t.se, -sg[1] * t.se, -sg[2] * t.se, -sg[3] * t.se);
          }
        }
      }
    }
  }
  const float sa = lo2(sa2) + hi2(sa2), sb = lo2(sb2) + hi2(sb2), sc = lo2(sc2) + hi2(sc2), sd = lo2(sd2) + hi2(sd2);
  const float gix = -((sb - sa) * (1.0f - t.ty) + (sd - sc) * t.ty);
  const float giy = -((sc - sa) * (1.0f - t.tx) + (sd - sb) * t.tx);
  if (a.grad_coords) {
    // ix = ((2 (x + 0.5) / w1 - 1) + 1) * lw / 2 - 0.5  =>  d ix / d x = lw / w1
    a.grad_coords[cbase * hw + pix] = gix * ((float)s.lw / (float)w);
    a.grad_coords[(cbase + (size_t)a.num_levels * a.num_samples) * hw + pix] = giy * ((float)s.lh / (float)h);
  }
}

// grad_fmap1 (NCHW) += the channel-quad interleaved workspace; one thread per (quad, pixel)
__global__ void __launch_bounds__(256) corr_unpack_kernel(const float4* __restrict__ ws, float* __restrict__ dst, int C,
                                                          int hw, size_t total) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const int p = (int)(i % hw);
  const size_t r = i / hw;
  const int q = (int)(r % (C / 4));
  const size_t img = r / (C / 4);
  const float4 v = ws[i];
  float* d = dst + (img * C + (size_t)q * 4) * hw + p;
  d[0] += v.x; d[hw] += v.y; d[2 * (size_t)hw] += v.z; d[3 * (size_t)hw] += v.w;
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
  MAL_REQUIRE(channels % 4 == 0, "mal_corr_pyramid: %d channels (the packed layout holds channel quads)", channels);
  MAL_REQUIRE((height >> (num_levels - 1)) > 0 && (width >> (num_levels - 1)) > 0,
              "mal_corr_pyramid: %d levels do not fit a %dx%d map", num_levels, height, width);
  MAL_REQUIRE(((uintptr_t)pyramid & 15) == 0, "mal_corr_pyramid: the pyramid buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t planes = (size_t)batch * (channels / 4);
  {
    const size_t total = planes * height * width;
    launch(corr_pack_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, fmap2,
           reinterpret_cast<float4*>(pyramid), channels, height * width, total);
  }
  int h = height, w = width;
  for (int l = 1; l < num_levels; l++) {
    const float4* in = reinterpret_cast<const float4*>(pyramid + corr_level_offset(batch, channels, height, width, l - 1));
    float4* out = reinterpret_cast<float4*>(pyramid + corr_level_offset(batch, channels, height, width, l));
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
  MAL_REQUIRE(a.channels % a.num_head == 0 && (a.channels / a.num_head) % 4 == 0,
              "%s: %d channels must split into %d heads of a multiple of 4 channels", who, a.channels, a.num_head);
  MAL_REQUIRE(a.channels / a.num_head <= 256, "%s: more than 256 channels per head changes ATen's summation tree", who);
  MAL_REQUIRE((a.height >> (a.num_levels - 1)) > 0 && (a.width >> (a.num_levels - 1)) > 0, "%s: too many levels", who);
  MAL_REQUIRE(a.fmap1 && a.pyramid && a.coords, "%s: fmap1 / pyramid / coords are required", who);
  MAL_REQUIRE(((uintptr_t)a.pyramid & 15) == 0 && ((uintptr_t)a.grad_pyramid & 15) == 0,
              "%s: pyramid buffers must be 16-byte aligned (128-bit loads and reductions)", who);
  return MAL_OK;
}

extern "C" int mal_corr_lookup(const mal_corr_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_corr_lookup: args is NULL");
  const mal_corr_args& a = *args;
  int rc = corr_check(a, "mal_corr_lookup");
  if (rc) return rc;
  MAL_REQUIRE(a.out, "mal_corr_lookup: out is required");
  const size_t total = corr_threads(a);
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
  const size_t total = corr_threads(a);
  cudaStream_t st = (cudaStream_t)stream;
  mal_corr_args k = a;
  const size_t nws = (size_t)a.batch * a.channels * a.height * a.width;
  if (!a.grad_fmap1) k.workspace = nullptr;
  if (k.workspace) {
    MAL_REQUIRE(((uintptr_t)k.workspace & 15) == 0, "mal_corr_lookup_backward: the workspace must be 16-byte aligned");
    cudaMemsetAsync(k.workspace, 0, nws * sizeof(float), st);
  }
  launch(corr_lookup_bwd_kernel, dim3((unsigned)((total + CR_NT - 1) / CR_NT)), dim3(CR_NT), 0, st, k);
  if (k.workspace)
    launch(corr_unpack_kernel, dim3((unsigned)((nws / 4 + 255) / 256)), dim3(256), 0, st,
           reinterpret_cast<const float4*>(k.workspace), a.grad_fmap1, a.channels, a.height * a.width, nws / 4);
  return check_launch("corr_lookup_bwd_kernel");
}
