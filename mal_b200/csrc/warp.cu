// warp.cu - DynamicDepth's forward warp: z-buffered splat + inverse warp, for sm_100a.
//
// Replaces dynamicdepth/rigid_warp.forward_warp (:534-597) with its helpers pixel2cam (:34-50),
// cam2pix_trans (:513-530), torch_sparse.coalesce(op='max') (:577), inverse_warp (:337-373) and
// cam2pixel (:54-83).  The reference runs it under torch.no_grad() (dynamicdepth/trainer.py:494),
// so it is forward only.
//
//   kernel 1  fw_splat_kernel    one thread per pixel of the `upscale`-times nearest-up-sampled depth
//                                map: back-project, move by `pose`, project, truncate to the integer
//                                target pixel and atomicMax the inverse depth there.  Positive floats
//                                order like their bit patterns, so the z-buffer is a uint32 plane and
//                                the max is exact and order-independent (deterministic).  The
//                                reference's per-sample Python loop over sparse tensors disappears.
//   kernel 2  fw_gather_kernel   one thread per target pixel: depth_w = 1 / zbuf, back-project with it,
//                                project with K @ [R|t] of the inverse pose, bilinear sample of the
//                                image with zeros padding (align_corners=True), validity masks.
//
// The 3x3 / 3x4 constants (inverse intrinsics, inverse pose through the reference's euler round
// trip) are a few scalars per sample and come from the host side of the boundary
// (mal_b200/rigid_warp.py), exactly as the reference derives them with torch.
#include "mal_math.cuh"

namespace mal {

constexpr int FW_NT = 256;

// row-major 3x3 @ [x,y,1] as the k-sequential FMA chain of the reference's bmm
__device__ __forceinline__ void mat3_pixel(const float* M, float x, float y, float* o) {
  o[0] = xfma(M[2], 1.0f, xfma(M[1], y, xmul(M[0], x)));
  o[1] = xfma(M[5], 1.0f, xfma(M[4], y, xmul(M[3], x)));
  o[2] = xfma(M[8], 1.0f, xfma(M[7], y, xmul(M[6], x)));
}
// 3x3 (row stride `ld`) @ v
__device__ __forceinline__ void mat3_vec(const float* M, int ld, const float* v, float* o) {
#pragma unroll
  for (int r = 0; r < 3; r++) o[r] = xfma(M[r * ld + 2], v[2], xfma(M[r * ld + 1], v[1], xmul(M[r * ld], v[0])));
}

__global__ void __launch_bounds__(FW_NT) fw_splat_kernel(const mal_forward_warp_args a, unsigned* __restrict__ zbuf) {
  __shared__ float sK[9], sKu[9], sP[12];
  const int b = blockIdx.y;
  if (threadIdx.x < 9) { sK[threadIdx.x] = a.K[b * 9 + threadIdx.x]; sKu[threadIdx.x] = a.Ku_inv[b * 9 + threadIdx.x]; }
  if (threadIdx.x >= 32 && threadIdx.x < 44) sP[threadIdx.x - 32] = a.pose[b * 12 + threadIdx.x - 32];
  __syncthreads();
  const int H = a.height, W = a.width, U = a.upscale, uW = W * U;
  const size_t n = (size_t)H * U * uW;
  for (size_t i = (size_t)blockIdx.x * FW_NT + threadIdx.x; i < n; i += (size_t)gridDim.x * FW_NT) {
    const int uy = (int)(i / uW), ux = (int)(i - (size_t)uy * uW);
    // F.interpolate(depth, scale_factor=upscale), nearest
    const float d = __ldg(a.depth + ((size_t)b * H + uy / U) * W + ux / U);
    float ray[3], cam[3], tr[3];
    mat3_pixel(sKu, (float)ux, (float)uy, ray);                       // pixel2cam
    cam[0] = xmul(ray[0], d); cam[1] = xmul(ray[1], d); cam[2] = xmul(ray[2], d);
    mat3_vec(sP, 4, cam, tr);                                         // cam2pix_trans: rot @ cam + tr
    const float X = xadd(tr[0], sP[3]), Y = xadd(tr[1], sP[7]);
    const float Z = fmaxf(xadd(tr[2], sP[11]), 1e-3f);                 // .clamp(min=1e-3)
    float pn[3] = {xdiv(X, Z), xdiv(Y, Z), xdiv(Z, Z)};
    float pix[3];
    mat3_vec(sK, 3, pn, pix);                                         // intrinsics @ P_norm
    // .long(): truncation toward zero, then out-of-range -> the discarded pad row / column
    const float fx = truncf(pix[0]), fy = truncf(pix[1]);
    if (!(fx >= 0.0f && fx <= (float)(W - 1) && fy >= 0.0f && fy <= (float)(H - 1))) continue;
    const float inv = xdiv(1.0f, Z);
    atomicMax(zbuf + ((size_t)b * H + (int)fy) * W + (int)fx, __float_as_uint(inv));
  }
}

__global__ void __launch_bounds__(FW_NT) fw_gather_kernel(const mal_forward_warp_args a,
                                                         const unsigned* __restrict__ zbuf) {
  __shared__ float sKi[9], sPr[12];
  const int b = blockIdx.y;
  if (threadIdx.x < 9) sKi[threadIdx.x] = a.K_inv[b * 9 + threadIdx.x];
  if (threadIdx.x >= 32 && threadIdx.x < 44) sPr[threadIdx.x - 32] = a.proj[b * 12 + threadIdx.x - 32];
  __syncthreads();
  const int H = a.height, W = a.width, C = a.channels;
  const size_t hw = (size_t)H * W;
  for (size_t p = (size_t)blockIdx.x * FW_NT + threadIdx.x; p < hw; p += (size_t)gridDim.x * FW_NT) {
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const float dense = __uint_as_float(zbuf[(size_t)b * hw + p]);
    const float fw = (dense == 0.0f) ? 0.0f : 1.0f;                    // 1 - (dense == 0)
    const float depth_w = fw != 0.0f ? xdiv(1.0f, dense) : 0.0f;      // depth_w[fw_val == 0] = 0
    float ray[3], cam[3], pc[3];
    mat3_pixel(sKi, (float)x, (float)y, ray);
    cam[0] = xmul(ray[0], depth_w); cam[1] = xmul(ray[1], depth_w); cam[2] = xmul(ray[2], depth_w);
    mat3_vec(sPr, 4, cam, pc);
    const float X = xadd(pc[0], sPr[3]), Y = xadd(pc[1], sPr[7]);
    const float Z = fmaxf(xadd(pc[2], sPr[11]), 1e-3f);
    // cam2pixel: 2*(X/Z)/(w-1) - 1
    const float gx = xsub(xdiv(xmul(2.0f, xdiv(X, Z)), (float)(W - 1)), 1.0f);
    const float gy = xsub(xdiv(xmul(2.0f, xdiv(Y, Z)), (float)(H - 1)), 1.0f);
    const float iw = (fmaxf(fabsf(gx), fabsf(gy)) <= 1.0f) ? 1.0f : 0.0f;
    const float valid = xmul(fw, iw);
    // grid_sample(zeros, align_corners=True)
    const float ux = unnormalize<MAL_CONV_MANYDEPTH>(gx, W), uy = unnormalize<MAL_CONV_MANYDEPTH>(gy, H);
    Taps t = make_taps(ux, uy, H, W);
    // floor of a coordinate far outside the image must not wrap the int conversion
    if (!(ux > -2.0f && ux < (float)W + 1.0f && uy > -2.0f && uy < (float)H + 1.0f)) t.v00 = t.v01 = t.v10 = t.v11 = false;
    for (int c = 0; c < C; c++) {
      const float v = bilinear(a.img + ((size_t)b * C + c) * hw, t);
      a.img_w[((size_t)b * C + c) * hw + p] = xmul(v, valid);
    }
    a.depth_w[(size_t)b * hw + p] = xmul(depth_w, valid);
    a.valid[(size_t)b * hw + p] = valid;
  }
}

}  // namespace mal

using namespace mal;

extern "C" int mal_forward_warp(const mal_forward_warp_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_forward_warp: args is NULL");
  const mal_forward_warp_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.batch <= 65535 && a.channels > 0 && a.height > 1 && a.width > 1 && a.upscale >= 1,
              "mal_forward_warp: bad shape B=%d C=%d %dx%d upscale=%d", a.batch, a.channels, a.height, a.width,
              a.upscale);
  MAL_REQUIRE(a.img && a.depth && a.pose && a.K && a.Ku_inv && a.K_inv && a.proj && a.img_w && a.depth_w && a.valid &&
                  a.zbuf,
              "mal_forward_warp: a required pointer is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t hw = (size_t)a.height * a.width;
  cudaError_t e = cudaMemsetAsync(a.zbuf, 0, (size_t)a.batch * hw * sizeof(unsigned), st);
  if (e != cudaSuccess) return fail(MAL_ERR_LAUNCH, "mal_forward_warp: memset: %s", cudaGetErrorString(e));
  size_t nu = hw * a.upscale * a.upscale;
  unsigned gx = (unsigned)((nu + FW_NT - 1) / FW_NT);
  if (gx > 148 * 16) gx = 148 * 16;
  launch(fw_splat_kernel, dim3(gx, a.batch), dim3(FW_NT), 0, st, a, reinterpret_cast<unsigned*>(a.zbuf));
  int rc = check_launch("fw_splat_kernel");
  if (rc) return rc;
  unsigned gy = (unsigned)((hw + FW_NT - 1) / FW_NT);
  if (gy > 148 * 16) gy = 148 * 16;
  launch(fw_gather_kernel, dim3(gy, a.batch), dim3(FW_NT), 0, st, a, reinterpret_cast<const unsigned*>(a.zbuf));
  return check_launch("fw_gather_kernel");
}
