// step.cu - the scalar tail of one MAL training step and the gradient hand-over, in one launch.
//
// Replaces, for `--distil` runs: the loss arithmetic at the end of compute_mono_losses
// (manydepth/loss_utils.py:115-127) and compute_main_losses (:215-281), the teacher->student
// accumulation in Trainer.process_batch (manydepth/trainer.py:624-629), LossBalancing.compute_loss
// (loss_utils.py:303-318) and the autograd chain from the total loss back to the two disparity
// maps and the two poses - about 60 one-element torch kernels per step in the op-by-op path.
//
//   L_teacher = R_t + smoothness * S_t             R = masked mean reprojection, S = smoothness
//   L_student = R_s + C + smoothness * S_s         C = consistency, D = distillation
//   loss_list = [L_student + L_teacher, D]
//   total     = batch * (w0 * loss_list[0] + w1 * loss_list[1])      (loss_blc; the reference adds the
//               weighted sum once per batch element) or loss_list[0] + loss_list[1] (no balancing)
//
// The fused kernels left un-normalised gradient planes behind; this kernel scales and adds them:
//   d total / d disp_teacher = k0 * ( gd_t / (W_t + 1e-7) + smoothness * gs_t ) [+ k1 * gdm]
//   d total / d disp_student = k0 * ( gd_s / (W_s + 1e-7) + gc + smoothness * gs_s ) + k1 * gdd
//   d total / d T_f          = K[:3]^T @ ( k0 / (W_t + 1e-7) * gP_t[f] )   (the student pass uses detached poses)
#include "mal_math.cuh"

namespace mal {

constexpr int ST_NT = 256;

struct StepCoefs { float k0, k1, inv_wt, inv_ws; };

__device__ __forceinline__ StepCoefs step_coefs(const mal_step_combine_args& a) {
  StepCoefs c;
  if (a.weights) {
    c.k0 = (float)a.batch * __ldg(a.weights);
    c.k1 = (float)a.batch * __ldg(a.weights + 1);
  } else {
    c.k0 = c.k1 = 1.0f;
  }
  c.inv_wt = 1.0f / (__ldg(a.sums_teacher + 1) + 1e-7f);
  c.inv_ws = 1.0f / (__ldg(a.sums_student + 1) + 1e-7f);
  return c;
}

template <int VEC>
__global__ void __launch_bounds__(ST_NT) step_combine_kernel(const mal_step_combine_args a) {
  const StepCoefs c = step_coefs(a);
  const size_t total = (size_t)a.batch * a.height * a.width / VEC;
  const float sm = a.smoothness;
  for (size_t i = (size_t)blockIdx.x * ST_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * ST_NT) {
    const size_t e = i * VEC;
    float gt[VEC], st[VEC], gs[VEC], ss[VEC], gc[VEC], gd[VEC], gm[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(gt) = *reinterpret_cast<const float4*>(a.gd_teacher + e);
      *reinterpret_cast<float4*>(st) = *reinterpret_cast<const float4*>(a.gs_teacher + e);
      *reinterpret_cast<float4*>(gs) = *reinterpret_cast<const float4*>(a.gd_student + e);
      *reinterpret_cast<float4*>(ss) = *reinterpret_cast<const float4*>(a.gs_student + e);
      *reinterpret_cast<float4*>(gc) = *reinterpret_cast<const float4*>(a.g_cons + e);
      *reinterpret_cast<float4*>(gd) = *reinterpret_cast<const float4*>(a.g_distil + e);
      if (a.g_distil_mono) *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(a.g_distil_mono + e);
    } else {
      gt[0] = a.gd_teacher[e]; st[0] = a.gs_teacher[e]; gs[0] = a.gd_student[e]; ss[0] = a.gs_student[e];
      gc[0] = a.g_cons[e]; gd[0] = a.g_distil[e];
      if (a.g_distil_mono) gm[0] = a.g_distil_mono[e];
    }
    if (a.smooth_stats) {   // chain the smoothness planes through disp / (mean + 1e-7): (g - L_b / HW) * s_b
      const size_t hw = (size_t)a.height * a.width;
      const float* sv = a.smooth_stats + (e / hw) * 4;   // a vector never straddles two samples (HW % VEC == 0)
      const float lt = __ldg(sv) / (float)hw, sct = __ldg(sv + 1), ls = __ldg(sv + 2) / (float)hw, scs = __ldg(sv + 3);
#pragma unroll
      for (int v = 0; v < VEC; v++) { st[v] = (st[v] - lt) * sct; ss[v] = (ss[v] - ls) * scs; }
    }
    float ot[VEC], os[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) {
      ot[v] = c.k0 * (gt[v] * c.inv_wt + sm * st[v]);
      if (a.g_distil_mono) ot[v] += c.k1 * gm[v];
      os[v] = c.k0 * (gs[v] * c.inv_ws + gc[v] + sm * ss[v]) + c.k1 * gd[v];
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(a.grad_disp_teacher + e) = *reinterpret_cast<float4*>(ot);
      *reinterpret_cast<float4*>(a.grad_disp_student + e) = *reinterpret_cast<float4*>(os);
    } else {
      a.grad_disp_teacher[e] = ot[0];
      a.grad_disp_student[e] = os[0];
    }
  }
  // one CTA also does the scalars and the pose gradients
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      const float Rt = __ldg(a.sums_teacher + 2), Rs = __ldg(a.sums_student + 2);
      const float Lt = Rt + sm * __ldg(a.smooth_teacher);
      const float Ls = Rs + __ldg(a.main_sums) + sm * __ldg(a.smooth_student);
      const float l0 = Ls + Lt, l1 = __ldg(a.main_sums + 1);
      a.scalars[0] = a.weights ? (float)a.batch * (__ldg(a.weights) * l0 + __ldg(a.weights + 1) * l1) : l0 + l1;
      a.scalars[1] = l0;
      a.scalars[2] = l1;
      a.scalars[3] = Rt;
      a.scalars[4] = Rs;
      a.scalars[5] = __ldg(a.main_sums);
      a.scalars[6] = Lt;
      a.scalars[7] = Ls;
    }
    // d/dT_f (4x4) = K[:3,:]^T (4x3) @ dP_f (3x4), scaled
    for (int i = threadIdx.x; i < a.batch * 2 * 16; i += ST_NT) {
      const int b = i / 32, f = (i >> 4) & 1, r = (i >> 2) & 3, col = i & 3;
      const float* K = a.K + b * 16;
      const float* gP = a.gP_teacher + (b * 2 + f) * 12;
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; k++) s += K[k * 4 + r] * gP[k * 4 + col];
      a.grad_T[f][b * 16 + r * 4 + col] = c.k0 * c.inv_wt * s;
    }
  }
}

}  // namespace mal

using namespace mal;

extern "C" int mal_step_combine(const mal_step_combine_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_step_combine: args is NULL");
  const mal_step_combine_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height > 0 && a.width > 0, "mal_step_combine: bad shape");
  MAL_REQUIRE(a.sums_teacher && a.sums_student && a.smooth_teacher && a.smooth_student && a.main_sums && a.K &&
                  a.gd_teacher && a.gs_teacher && a.gP_teacher && a.gd_student && a.gs_student && a.g_cons &&
                  a.g_distil && a.scalars && a.grad_disp_teacher && a.grad_disp_student && a.grad_T[0] && a.grad_T[1],
              "mal_step_combine: a required pointer is NULL");
  const size_t n = (size_t)a.batch * a.height * a.width;
  auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = ((size_t)a.height * a.width) % 4 == 0 && al(a.gd_teacher) && al(a.gs_teacher) && al(a.gd_student) && al(a.gs_student) &&
                   al(a.g_cons) && al(a.g_distil) && al(a.g_distil_mono) && al(a.grad_disp_teacher) &&
                   al(a.grad_disp_student);
  size_t blk = (n / (vec ? 4 : 1) + ST_NT - 1) / ST_NT;
  if (blk > 148 * 8) blk = 148 * 8;
  if (blk < 1) blk = 1;
  if (vec) launch(step_combine_kernel<4>, dim3((unsigned)blk), dim3(ST_NT), 0, (cudaStream_t)stream, a);
  else launch(step_combine_kernel<1>, dim3((unsigned)blk), dim3(ST_NT), 0, (cudaStream_t)stream, a);
  return check_launch("step_combine_kernel");
}
