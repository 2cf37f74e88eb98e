// dynsyn.cu - MAL's temporal hint: synthesising the two "motion-compensated" source images from
// matched instance masks, for sm_100a.
//
// Replaces manydepth/dyn_utils.py: generate_dynamic_instance (:38-119) and fill_dynamic_obj (:5-36),
// which the reference runs as TorchScript with Python-level loops over the instances (one slice
// assignment per instance, (N,3,H,W) temporaries).  Integer / byte work:
//
//   kernel 1  dyn_extents_kernel   per (instance, frame): bounding extents of the mask with the
//                                  reference's index-weighted sums, i.e. row 0 / column 0 never
//                                  count (their weight is 0) and an empty mask gives 0 (:52-78)
//   kernel 2  dyn_delta_kernel     per instance: the larger of the two edge displacements per axis,
//                                  halved and rounded half-to-even, optional dead zone (:80-100)
//   kernel 3  dyn_compose_kernel   per pixel: background swap where an instance left / entered
//                                  (:102-112), the shifted instances summed in instance order
//                                  (fill_dynamic_obj, torch's cascade summation over dim 0), and the
//                                  final where(mask_or, synthesised, original) (:114-118)
//
// mal_fill_dynamic_obj exposes kernel 3's inner function with caller-given displacements.
#include "mal_math.cuh"

namespace mal {

constexpr int DS_NT = 256;

// fill_dynamic_obj at one pixel: sum over instances of source[:, y-dx, x-dy] where the shifted
// mask is set; *any tells whether at least one instance covers the pixel.
__device__ __forceinline__ void fill_pixel(const uint8_t* __restrict__ mask, const int* __restrict__ dx,
                                           const int* __restrict__ dy, const float* __restrict__ source, int N, int C,
                                           int H, int W, int y, int x, float* acc /*[C<=4]*/, bool* any) {
  const size_t hw = (size_t)H * W;
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};   // cascade_sum: 16-element chunks
  bool hit = false;
#pragma unroll 4   // independent mask loads of several instances in flight
  for (int n = 0; n < N; n++) {
    const int sy = y - dx[n], sx = x - dy[n];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (sy >= 0 && sy < H && sx >= 0 && sx < W && mask[(size_t)n * hw + (size_t)sy * W + sx]) {
      hit = true;
      for (int c = 0; c < C; c++) v[c] = __ldg(source + c * hw + (size_t)sy * W + sx);
    }
    for (int c = 0; c < C; c++) a0[c] = xadd(a0[c], v[c]);
    if ((n & 15) == 15) {
      for (int c = 0; c < C; c++) { a1[c] = xadd(a1[c], a0[c]); a0[c] = 0.0f; }
    }
  }
  for (int c = 0; c < C; c++) acc[c] = xadd(a0[c], a1[c]);
  *any = hit;
}

constexpr int DX_NT = 1024;   // one CTA per (instance, frame); 16 mask bytes per load

struct __align__(16) MaskWords { unsigned w[4]; };

__global__ void __launch_bounds__(DX_NT) dyn_extents_kernel(const uint8_t* __restrict__ mask_last,
                                                           const uint8_t* __restrict__ mask_next, int H, int W,
                                                           int* __restrict__ ext /*[N][2][4]*/) {
  __shared__ int red[4][DX_NT / 32];
  const int n = blockIdx.x, which = blockIdx.y;
  const int total = H * W;
  const uint8_t* m = (which == 0 ? mask_last : mask_next) + (size_t)n * total;
  int low = 0, top = 0x7fffffff, right = 0, left = 0x7fffffff;
  auto visit = [&](int i) {
    const int h = i / W, w = i - h * W;
    if (h >= 1) { low = max(low, h); top = min(top, h); }     // (mask * grid_h).sum(2) is 0 on row 0
    if (w >= 1) { right = max(right, w); left = min(left, w); }
  };
  // the bulk of a mask is zero: read it 16 bytes at a time and only decode the words that are not
  const int nvec = (((uintptr_t)m & 15) == 0) ? total / 16 : 0;
  const MaskWords* mv = reinterpret_cast<const MaskWords*>(m);
  for (int v = threadIdx.x; v < nvec; v += DX_NT) {
    const MaskWords q = mv[v];
    if ((q.w[0] | q.w[1] | q.w[2] | q.w[3]) == 0u) continue;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (q.w[k] == 0u) continue;
#pragma unroll
      for (int bt = 0; bt < 4; bt++)
        if ((q.w[k] >> (8 * bt)) & 0xffu) visit(v * 16 + k * 4 + bt);
    }
  }
  for (int i = nvec * 16 + threadIdx.x; i < total; i += DX_NT)
    if (m[i]) visit(i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    low = max(low, __shfl_xor_sync(0xffffffffu, low, o));
    top = min(top, __shfl_xor_sync(0xffffffffu, top, o));
    right = max(right, __shfl_xor_sync(0xffffffffu, right, o));
    left = min(left, __shfl_xor_sync(0xffffffffu, left, o));
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][warp] = low; red[1][warp] = top; red[2][warp] = right; red[3][warp] = left; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < DX_NT / 32; wv++) {
      low = max(low, red[0][wv]); top = min(top, red[1][wv]); right = max(right, red[2][wv]); left = min(left, red[3][wv]);
    }
    int* e = ext + (n * 2 + which) * 4;
    e[0] = low;
    e[1] = top == 0x7fffffff ? 0 : top;       // argmin over an all-"inf" row is index 0
    e[2] = right;
    e[3] = left == 0x7fffffff ? 0 : left;
  }
}

__global__ void dyn_delta_kernel(const int* __restrict__ ext, int N, int replace, int* __restrict__ delta /*[4][N]*/) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int* l = ext + (n * 2 + 0) * 4;
  const int* x = ext + (n * 2 + 1) * 4;
  const int dx0 = x[0] - l[0], dx1 = x[1] - l[1];     // [low_next-low_last, top_next-top_last]
  const int dy0 = x[2] - l[2], dy1 = x[3] - l[3];     // [right..., left...]
  const int sx = (abs(dx1) > abs(dx0)) ? dx1 : dx0;   // abs().argmax(): first index wins ties
  const int sy = (abs(dy1) > abs(dy0)) ? dy1 : dy0;
  int px = (int)rintf((float)sx / 2.0f), py = (int)rintf((float)sy / 2.0f);   // torch.round: half to even
  if (replace) {
    if (abs(px) < 3) px = 0;
    if (abs(py) < 3) py = 0;
  }
  delta[0 * N + n] = px; delta[1 * N + n] = py;       // delta_x_last, delta_y_last
  delta[2 * N + n] = -px; delta[3 * N + n] = -py;     // delta_x_next, delta_y_next
}

__global__ void __launch_bounds__(DS_NT) dyn_compose_kernel(const mal_dynamic_instance_args a,
                                                           const int* __restrict__ delta) {
  const int N = a.num, C = a.channels, H = a.height, W = a.width;
  const size_t hw = (size_t)H * W;
  for (size_t p = (size_t)blockIdx.x * DS_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * DS_NT) {
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    bool m_or = false, bg = false, bg2 = false;
#pragma unroll 4
    for (int n = 0; n < N; n++) {
      const bool ml = a.mask_last[(size_t)n * hw + p] != 0, mn = a.mask_next[(size_t)n * hw + p] != 0;
      m_or |= ml | mn;
      bg |= ml & !mn;      // the instance left this pixel: show the other frame's background
      bg2 |= mn & !ml;
    }
    float il[4], in_[4];
    for (int c = 0; c < C; c++) { il[c] = __ldg(a.img_last + c * hw + p); in_[c] = __ldg(a.img_next + c * hw + p); }
    float acc[4];
    bool any;
    fill_pixel(a.mask_last, delta, delta + N, a.img_last, N, C, H, W, y, x, acc, &any);
    for (int c = 0; c < C; c++) {
      const float syn = any ? acc[c] : (bg ? in_[c] : il[c]);
      a.ori_last[c * hw + p] = m_or ? syn : il[c];
    }
    fill_pixel(a.mask_next, delta + 2 * N, delta + 3 * N, a.img_next, N, C, H, W, y, x, acc, &any);
    for (int c = 0; c < C; c++) {
      const float syn = any ? acc[c] : (bg2 ? il[c] : in_[c]);
      a.ori_next[c * hw + p] = m_or ? syn : in_[c];
    }
  }
}

__global__ void __launch_bounds__(DS_NT) dyn_fill_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ dx,
                                                        const int* __restrict__ dy, const float* __restrict__ source,
                                                        const float* __restrict__ img, int N, int C, int H, int W,
                                                        float* __restrict__ out) {
  const size_t hw = (size_t)H * W;
  for (size_t p = (size_t)blockIdx.x * DS_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * DS_NT) {
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    float acc[4];
    bool any;
    fill_pixel(mask, dx, dy, source, N, C, H, W, y, x, acc, &any);
    for (int c = 0; c < C; c++) out[c * hw + p] = any ? acc[c] : __ldg(img + c * hw + p);
  }
}

// ---- backward of the composition (the reference's copies keep autograd history) --------------------
// pass 1: per-pixel selection flags of the forward: bit0 mask_or, bit1 any_last, bit2 any_next, bit3 bg, bit4 bg2
__global__ void __launch_bounds__(DS_NT) dyn_flags_kernel(const mal_dynamic_instance_args a,
                                                         const int* __restrict__ delta, uint8_t* __restrict__ flags) {
  const int N = a.num, H = a.height, W = a.width;
  const size_t hw = (size_t)H * W;
  for (size_t p = (size_t)blockIdx.x * DS_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * DS_NT) {
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    unsigned f = 0;
    for (int n = 0; n < N; n++) {
      const bool ml = a.mask_last[(size_t)n * hw + p] != 0, mn = a.mask_next[(size_t)n * hw + p] != 0;
      if (ml | mn) f |= 1u;
      if (ml & !mn) f |= 8u;
      if (mn & !ml) f |= 16u;
      int sy = y - delta[n], sx = x - delta[N + n];
      if (sy >= 0 && sy < H && sx >= 0 && sx < W && a.mask_last[(size_t)n * hw + (size_t)sy * W + sx]) f |= 2u;
      sy = y - delta[2 * N + n]; sx = x - delta[3 * N + n];
      if (sy >= 0 && sy < H && sx >= 0 && sx < W && a.mask_next[(size_t)n * hw + (size_t)sy * W + sx]) f |= 4u;
    }
    flags[p] = (uint8_t)f;
  }
}

// pass 2: gather.  ori_last(q) = !or ? last(q) : any_l ? sum_n [mask_last_n(q-d)] last(q-d) : bg ? next(q) : last(q)
//                  ori_next(q) = !or ? next(q) : any_n ? sum_n [mask_next_n(q-d)] next(q-d) : bg2 ? last(q) : next(q)
__global__ void __launch_bounds__(DS_NT) dyn_backward_kernel(const mal_dynamic_instance_args a,
                                                            const int* __restrict__ delta,
                                                            const uint8_t* __restrict__ flags,
                                                            const float* __restrict__ g_ol,
                                                            const float* __restrict__ g_on, float* __restrict__ g_last,
                                                            float* __restrict__ g_next) {
  const int N = a.num, C = a.channels, H = a.height, W = a.width;
  const size_t hw = (size_t)H * W;
  for (size_t p = (size_t)blockIdx.x * DS_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * DS_NT) {
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const unsigned f = flags[p];
    float gl[4] = {0.f, 0.f, 0.f, 0.f}, gn[4] = {0.f, 0.f, 0.f, 0.f};
    // pass-through terms at this pixel
    const bool l_self = !(f & 1u) || (!(f & 2u) && !(f & 8u));    // ori_last(p) reads last(p)
    const bool l_from_next = (f & 1u) && !(f & 2u) && (f & 8u);   // ori_last(p) reads next(p)
    const bool n_self = !(f & 1u) || (!(f & 4u) && !(f & 16u));
    const bool n_from_last = (f & 1u) && !(f & 4u) && (f & 16u);
    for (int c = 0; c < C; c++) {
      const float a_l = g_ol[c * hw + p], a_n = g_on[c * hw + p];
      if (l_self) gl[c] += a_l;
      if (l_from_next) gn[c] += a_l;
      if (n_self) gn[c] += a_n;
      if (n_from_last) gl[c] += a_n;
    }
    // shifted copies that read this pixel
    for (int n = 0; n < N; n++) {
      if (a.mask_last[(size_t)n * hw + p]) {
        const int qy = y + delta[n], qx = x + delta[N + n];
        if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
          const size_t q = (size_t)qy * W + qx;
          if (flags[q] & 1u)
            for (int c = 0; c < C; c++) gl[c] += g_ol[c * hw + q];
        }
      }
      if (a.mask_next[(size_t)n * hw + p]) {
        const int qy = y + delta[2 * N + n], qx = x + delta[3 * N + n];
        if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
          const size_t q = (size_t)qy * W + qx;
          if (flags[q] & 1u)
            for (int c = 0; c < C; c++) gn[c] += g_on[c * hw + q];
        }
      }
    }
    for (int c = 0; c < C; c++) { g_last[c * hw + p] = gl[c]; g_next[c * hw + p] = gn[c]; }
  }
}

inline unsigned ds_blocks(size_t n) {
  size_t b = (n + DS_NT - 1) / DS_NT;
  return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace mal

using namespace mal;

extern "C" int mal_dynamic_instance(const mal_dynamic_instance_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_dynamic_instance: args is NULL");
  const mal_dynamic_instance_args& a = *args;
  MAL_REQUIRE(a.num > 0 && a.num < 256 && a.channels > 0 && a.channels <= 4 && a.height > 0 && a.width > 0,
              "mal_dynamic_instance: bad shape N=%d C=%d %dx%d (1 <= N < 256, C <= 4)", a.num, a.channels, a.height,
              a.width);
  MAL_REQUIRE(a.mask_last && a.mask_next && a.img_last && a.img_next && a.ori_last && a.ori_next && a.workspace,
              "mal_dynamic_instance: a required pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int* ext = a.workspace;
  int* delta = a.workspace + (size_t)a.num * 8;
  launch(dyn_extents_kernel, dim3(a.num, 2), dim3(DX_NT), 0, st, a.mask_last, a.mask_next, a.height, a.width, ext);
  int rc = check_launch("dyn_extents_kernel");
  if (rc) return rc;
  launch(dyn_delta_kernel, dim3((a.num + 63) / 64), dim3(64), 0, st, (const int*)ext, a.num, a.replace, delta);
  rc = check_launch("dyn_delta_kernel");
  if (rc) return rc;
  launch(dyn_compose_kernel, dim3(ds_blocks((size_t)a.height * a.width)), dim3(DS_NT), 0, st, a, (const int*)delta);
  return check_launch("dyn_compose_kernel");
}

extern "C" int mal_fill_dynamic_obj(const uint8_t* mask, const int32_t* delta_x, const int32_t* delta_y,
                                    const float* source, const float* img, int num, int channels, int height,
                                    int width, float* out, mal_stream_t stream) {
  MAL_REQUIRE(mask && delta_x && delta_y && source && img && out, "mal_fill_dynamic_obj: a required pointer is NULL");
  MAL_REQUIRE(num > 0 && num < 256 && channels > 0 && channels <= 4 && height > 0 && width > 0,
              "mal_fill_dynamic_obj: bad shape N=%d C=%d %dx%d", num, channels, height, width);
  launch(dyn_fill_kernel, dim3(ds_blocks((size_t)height * width)), dim3(DS_NT), 0, (cudaStream_t)stream, mask, delta_x,
         delta_y, source, img, num, channels, height, width, out);
  return check_launch("dyn_fill_kernel");
}

extern "C" int mal_dynamic_instance_backward(const mal_dynamic_instance_args* args, const int32_t* deltas,
                                             const float* grad_ori_last, const float* grad_ori_next,
                                             float* grad_img_last, float* grad_img_next, uint8_t* flags,
                                             mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_dynamic_instance_backward: args is NULL");
  const mal_dynamic_instance_args& a = *args;
  MAL_REQUIRE(a.num > 0 && a.num < 256 && a.channels > 0 && a.channels <= 4 && a.height > 0 && a.width > 0,
              "mal_dynamic_instance_backward: bad shape");
  MAL_REQUIRE(a.mask_last && a.mask_next && deltas && grad_ori_last && grad_ori_next && grad_img_last &&
                  grad_img_next && flags,
              "mal_dynamic_instance_backward: a required pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned nb = ds_blocks((size_t)a.height * a.width);
  launch(dyn_flags_kernel, dim3(nb), dim3(DS_NT), 0, st, a, deltas, flags);
  int rc = check_launch("dyn_flags_kernel");
  if (rc) return rc;
  launch(dyn_backward_kernel, dim3(nb), dim3(DS_NT), 0, st, a, deltas, (const uint8_t*)flags, grad_ori_last,
         grad_ori_next, grad_img_last, grad_img_next);
  return check_launch("dyn_backward_kernel");
}
