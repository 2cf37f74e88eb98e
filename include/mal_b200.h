/* mal_b200.h - C ABI of libmal_b200.so, the sm_100a implementation of the MAL photometric
 * hot path.
 *
 * The reference (YuejiangDong/MAL) has no FFI: the hot path is reached through Python
 * signatures.  Each entry point below names the reference code it replaces
 * (paths relative to the reference root).  mal_b200/_capi.py binds these with ctypes and
 * mal_b200/ops.py registers them as torch custom ops with autograd; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 NCHW data unless stated; NULL
 *     means "absent / not requested";
 *   - the caller owns every buffer (inputs, outputs and workspaces); the library never
 *     allocates, frees or retains a pointer;
 *   - work is enqueued on `stream` (a cudaStream_t); no call synchronises;
 *   - re-entrant, no global mutable state except the thread-local error string;
 *   - return 0 on success, a MAL_ERR_* code otherwise; mal_last_error() describes it.
 */
#ifndef MAL_B200_H_
#define MAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAL_ABI_VERSION 5

enum {
  MAL_OK = 0,
  MAL_ERR_ARGUMENT = 1, /* bad size / NULL / unsupported flag combination */
  MAL_ERR_LAUNCH = 2,   /* CUDA launch or runtime error                   */
  MAL_ERR_ARCH = 3      /* device is not sm_100                           */
};

/* Project3D / grid_sample conventions */
enum {
  MAL_CONV_MANYDEPTH = 0, /* x/(W-1), align_corners=True : manydepth/layers.py:192-194, trainer.py:1122 */
  MAL_CONV_DUALREFINE = 1 /* 2(x+.5)/W-1, align_corners=False : dualrefine/layers.py:224-225, trainer.py:444 */
};

typedef void* mal_stream_t; /* cudaStream_t */

int mal_abi_version(void);
const char* mal_last_error(void);
/* 0 if device `device` is compute capability 10.x, MAL_ERR_ARCH otherwise. */
int mal_check_device(int device);

/* ------------------------------------------------------------------------------------------
 * 1/2. Fused photometric loss, forward + (optionally) backward in one pass.
 *
 * Replaces, per call: disp_to_depth (manydepth/layers.py:14-23), BackprojectDepth.forward
 * (:163-168), Project3D.forward (:184-199), F.grid_sample(border) (manydepth/trainer.py:1122-1125),
 * SSIM.forward (layers.py:243-257), compute_reprojection_loss (loss_utils.py:46-55), the
 * per-pixel min over candidates (:103), the tie-break noise and automask (:105-109 /
 * compute_loss_masks :27-44), the multi-frame mask (:192-194) and the masked sum (:112-113),
 * plus their autograd backward.
 *
 * Candidates, in the reference's order: [pred(-1), pred(+1), syn(-1), syn(+1)] where pred(f)
 * is either warped in-kernel from src[f] (mode WARP) or given (mode PRED).
 * ------------------------------------------------------------------------------------------ */
enum { MAL_PHOTO_WARP = 0, MAL_PHOTO_PRED = 1 };

typedef struct mal_photo_args {
  int32_t batch, height, width;
  int32_t mode;        /* MAL_PHOTO_WARP | MAL_PHOTO_PRED                                   */
  int32_t convention;  /* MAL_CONV_*                                                        */
  int32_t depth_is_disp; /* 1: `depth` holds sigmoid disparity, converted with min/max_depth */
  int32_t no_ssim;     /* opt.no_ssim: L1 only (manydepth/trainer.py:1217)                  */
  int32_t with_grad;   /* 1: also emit the un-normalised gradients                          */
  double min_depth, max_depth; /* opt.min_depth / opt.max_depth as python floats            */
  float eps;           /* Project3D eps (1e-7)                                              */

  const float* target;      /* (B,3,H,W) inputs[("color",0,0)]                              */
  const float* src[2];      /* (B,3,H,W) frames -1,+1: WARP: sampled; PRED: pre-warped preds;
                               PRED with src[1]==NULL: one candidate (compute_reprojection_loss) */
  const float* syn[2];      /* (B,3,H,W) optional MAL temporal-hint candidates, both or none */
  const float* depth;       /* (B,1,H,W) depth or disparity (WARP mode)                     */
  const float* K;           /* (B,4,4)                                                      */
  const float* inv_K;       /* (B,4,4)                                                      */
  const float* T[2];        /* (B,4,4) cam_T_cam for frames -1,+1                           */
  const float* identity_min;/* (B,1,H,W) min identity reprojection; with `noise` => automask */
  const float* noise;       /* (B,1,H,W) raw randn draw; the kernel applies the 1e-5 scale  */
  const float* pixel_mask;  /* (B,H,W)   optional outputs["consistency_mask"]               */
  const float* sample_mask; /* (B)       optional outputs["augmentation_mask"]: w *= 1-m[b] */

  float* min_reproj;        /* (B,1,H,W) optional: min over candidates                      */
  uint8_t* selection;       /* (B,1,H,W) optional: argmin candidate | (automask bit << 7)   */
  float* weight;            /* (B,1,H,W) optional: final per-pixel loss weight              */
  float* grad_depth;        /* (B,1,H,W) WARP+grad: d(sum w*reproj)/d depth (or /d disp)    */
  float* grad_pred[2];      /* (B,3,H,W) PRED+grad: d(sum w*reproj)/d pred(f)               */
  float* partials;          /* workspace, mal_photo_partials_floats() floats                */
  float* sums;              /* (4): [sum w*reproj, sum w, sum w*reproj/(sum w+1e-7), 0]      */
  float* grad_P;            /* (B,2,12) WARP+grad: d(sum w*reproj)/d (K@T)[:3,:] per frame   */
  const float* depth_b;     /* (B,1,H,W) optional: the kernel uses (depth + depth_b) / 2, the
                               ensemble disparity of manydepth/trainer.py:598; no gradient      */
  float* grad_syn[2];       /* (B,3,H,W) optional, with_grad and syn given (either mode):
                               d(sum w*reproj)/d syn(f), non-zero where a temporal-hint candidate is
                               the per-pixel minimum; NULL: the candidates are treated as data     */
  int32_t depth_height, depth_width; /* 0,0: `depth`/`depth_b` are (B,1,H,W).  Otherwise they are
                               (B,1,depth_height,depth_width) and the kernel reads them through
                               F.interpolate(.., [H,W], mode="bilinear", align_corners=False)
                               (manydepth/trainer.py:1093-1094): the full-resolution disparity never
                               exists.  grad_depth stays (B,1,H,W) = d/d(up-sampled value); take it to
                               the low resolution with mal_upsample_bilinear_backward              */
  float* min_reproj_b;      /* (B,1,H,W) optional, forward-only with four candidates: `min_reproj` is then the min over
                               candidates 0, 1 and this plane the min over candidates 2, 3 - two independent
                               2-candidate passes against one target in one launch (the step scores the ensemble
                               warps and the un-warped sources of the automask, trainer.py:1172-1207 and
                               loss_utils.py:92-101, together: pass the sources as `syn`)                       */
  /* DynamicDepth's compute_losses as one pass per scale (dynamicdepth/trainer.py:958-975, :1006-1128): WARP mode,
     MAL_CONV_MANYDEPTH, the identity candidates of the automask (inputs[("color", f, 0)]) given as `syn`, `noise`
     for its tie-break, no `identity_min`.                                                                         */
  int32_t zero_img;         /* opt.zero_img: a prediction's dark pixels (RGB sum < 0.1) are zeroed in the prediction and
                               in the target as the reference's calls c0, c1, i0, i1 do one after the other           */
  int32_t selec_reproj;     /* opt.selec_reproj: where one warp is dark take the other one's loss, where both are, 0      */
  int32_t ignore_automask;  /* is_multi: the automask is computed (its calls still zero the target) but not applied     */
  int32_t identity_in_pass; /* selects this mode when neither zero_img nor selec_reproj is set: `syn` holds the identity
                               candidates, their min (+ noise) is the automask's other side                             */
  float* target_out;        /* (B,3,H,W) optional: the target as this pass's calls leave it = the next scale's target  */
  const float* warped[2];   /* (B,3,H,W) optional, WARP mode: the warped sources outputs[("color", f, s)] when the caller has
                               materialised them anyway (trainer.py:1122-1125 feeds them to image_synthesis): the pass
                               stages them instead of re-warping; gradients still chain through depth / T / src      */
  int32_t avg_reprojection; /* opt.avg_reprojection (dualrefine/trainer.py:575-586, dynamicdepth/trainer.py:1044-1056):
                               mean instead of min over the two candidates (no syn); selection index is 0; with
                               gradients (WARP mode) both warps carry half of it                             */
  int32_t skip_finalize;    /* 1: only the tile kernel runs; `sums` / `grad_P` are produced later by
                               mal_photo_finalize(args, stream) - lets a scheduler keep the tiny
                               reduction off the critical path between two heavy kernels           */
} mal_photo_args;

size_t mal_photo_partials_floats(int batch, int height, int width);
int mal_photo_forward(const mal_photo_args* args, mal_stream_t stream);
/* The deterministic reduction of the per-tile partials that mal_photo_forward normally ends with;
 * same args as the forward call it completes (only batch/height/width/mode/with_grad/partials/sums/grad_P
 * are read). */
int mal_photo_finalize(const mal_photo_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 5. Plane-sweep matching cost volume, forward only (the reference builds it under
 *    torch.no_grad(): manydepth/networks/resnet_encoder.py:292-307).
 *
 * Replaces ResnetEncoderMatching.match_features (manydepth/networks/resnet_encoder.py:151-233;
 * dualrefine/networks/resnet_encoder.py:163-245 with MAL_CONV_DUALREFINE;
 * dynamicdepth/networks/resnet_encoder.py:148-249 with the cv_min / occ fields) and, when the optional
 * outputs are given, the head of forward(): compute_confidence_mask (:255-262), the "viz"
 * arg-min and indices_to_disparity (:247-253, :309-313) and cost_volume *= confidence (:317).
 * A lookup frame whose 4x4 pose sums to 0 is skipped on the device (:183-185, no host sync).
 * ------------------------------------------------------------------------------------------ */
enum { MAL_CV_OCC_NONE = 0, MAL_CV_OCC_SET_1 = 1, MAL_CV_OCC_POOL = 2 };

typedef struct mal_cost_volume_args {
  int32_t batch, channels, height, width; /* matching resolution (H/4, W/4)                  */
  int32_t num_lookup, num_bins;
  int32_t convention;          /* MAL_CONV_*                                                 */
  int32_t set_missing_to_max;  /* self.set_missing_to_max (:223)                              */
  int32_t apply_confidence;    /* 1: cost_volume is multiplied by the confidence mask (:317)  */
  int32_t num_bins_threshold;  /* compute_confidence_mask threshold; <=0: num_bins            */
  float eps;                   /* Project3D eps                                               */

  const float* current;        /* (B,C,h,w)   current_feats                                   */
  const float* lookup;         /* (B,F,C,h,w) lookup_feats                                    */
  const float* poses;          /* (B,F,4,4)   relative_poses                                  */
  const float* K;              /* (B,4,4)     K at the matching scale                         */
  const float* inv_K;          /* (B,4,4)                                                     */
  const float* bins;           /* (num_bins)  depth_bins (compute_depth_bins :121-141)        */

  float* cost_volume;          /* (B,num_bins,h,w)                                            */
  float* missing_mask;         /* (B,num_bins,h,w) optional                                   */
  float* confidence;           /* (B,h,w) optional                                            */
  int32_t* argmin;             /* (B,h,w) optional: arg-min bin of the 0->100 "viz" volume    */
  float* lowest_cost;          /* (B,h,w) optional: 1 / bins[argmin]                          */
  float* packed;               /* workspace, mal_cost_volume_workspace_floats() floats, 16-B aligned */

  /* DynamicDepth variant (dynamicdepth/networks/resnet_encoder.py:148-249); all zero/NULL = off */
  int32_t cv_min;              /* min over lookup frames instead of the mean (:163-166, :220-227)  */
  int32_t occ_mode;            /* MAL_CV_OCC_*: what to do where the projected occlusion mask > pool_th */
  int32_t pool_radius;         /* pool_r: (2r+1)^3 max-pool window over (bin, y, x) (:199)         */
  float pool_th;               /* pool_th (:195)                                                   */
  const float* occ;            /* (B,h,w) {0,1}: occ_batch > 0 at the matching resolution (:160, :194) */
  const float* aug_mask;       /* (B) occlusion handling only where aug_mask == 0 (:192); NULL: all */
  float* desc;                 /* workspace of mal_cost_volume_desc_floats() floats, required with MAL_CV_OCC_POOL:
                                  every (lookup frame, bin, pixel) is projected once into it, the rim of every
                                  occluded blob is listed there, the warped vectors its pool windows reach are
                                  cached there and its pooled chunk sums are formed there                        */
} mal_cost_volume_args;

size_t mal_cost_volume_workspace_floats(int batch, int channels, int height, int width, int num_lookup);
size_t mal_cost_volume_desc_floats(int batch, int channels, int num_lookup, int num_bins, int height, int width);
int mal_cost_volume_forward(const mal_cost_volume_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3. Edge-aware smoothness, forward + backward in one call.
 *
 * Replaces get_smooth_loss (manydepth/layers.py:210-223) and, with `normalise`, the
 * mean-normalisation in front of it (manydepth/loss_utils.py:119-121, trainer.py:1440-1442).
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_smooth_args {
  int32_t batch, height, width;
  int32_t normalise;          /* 1: disp / (disp.mean(2,True).mean(3,True) + 1e-7) first      */
  int32_t with_grad;
  const float* disp;          /* (B,1,h,w)                                                    */
  const float* img;           /* (B,3,h,w)                                                    */
  float* grad_disp;           /* (B,1,h,w) with_grad: d loss / d disp                         */
  float* workspace;           /* mal_smooth_workspace_floats() floats                         */
  float* loss;                /* (1)                                                          */
  /* a second disparity scored against the same image in the same launch (teacher + student of one step:
     the edge weights are computed once); NULL: one term                                              */
  const float* disp_b;        /* (B,1,h,w) optional                                           */
  float* grad_disp_b;         /* (B,1,h,w) with_grad and disp_b                               */
  float* loss_b;              /* (1)       with disp_b                                        */
  int32_t defer_fix;          /* 1 (normalise + with_grad): grad_disp[_b] are left as d loss / d (normalised
                                 disp) and `stats` tells the consumer how to chain through the normalisation:
                                 d loss / d disp_i = (g_i - stats[b][t][0] / (h*w)) * stats[b][t][1]
                                 (mal_step_combine does this while it adds the gradient planes)   */
  float* stats;               /* (B,2,2) optional: per sample and term [L_b, 1 / (mean_b + 1e-7)]    */
} mal_smooth_args;

size_t mal_smooth_workspace_floats(int batch, int height, int width);
int mal_smooth_forward(const mal_smooth_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 4. MAL student ("main") terms at scale 0: consistency L1 and the distillation selection.
 *
 * Replaces the per-pixel part of compute_main_losses (manydepth/loss_utils.py:192-254):
 * mask / consistency mask (:192-196), consistency loss + target (:205-213), the arg-min over
 * [mono_reproj, (ensemble_reproj), multi_reproj] (:228-229 / :237-238), torch.where depth
 * selection (:231-245) and the distillation loss (:253-254), with their backward.
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_main_terms_args {
  int32_t batch, height, width;
  int32_t inputs_are_disp;    /* 1: multi/mono hold sigmoid disparity (disp_to_depth applied)  */
  int32_t dual_distil;        /* opt.dual_distil: the mono teacher also receives a gradient    */
  int32_t with_grad;
  double min_depth, max_depth;
  const float* multi;         /* (B,1,H,W) outputs[("depth",0,0)] or ("disp",0)                */
  const float* mono;          /* (B,1,H,W) outputs[("mono_depth",0,0)] or ("mono_disp",0)      */
  const float* pixel_mask;    /* (B,H,W)   outputs["consistency_mask"]                         */
  const float* sample_mask;   /* (B)       outputs["augmentation_mask"], optional              */
  const float* mono_reproj;   /* (B,1,H,W) teacher min reprojection                            */
  const float* ens_reproj;    /* (B,1,H,W) optional ensemble min reprojection                  */
  const float* multi_reproj;  /* (B,1,H,W) student min reprojection                            */
  uint8_t* distil_index;      /* (B,1,H,W) optional: arg-min                                   */
  float* consistency_target;  /* (B,1,H,W) optional: outputs["consistency_target/0"] (:211-213)*/
  float* grad_cons;           /* (B,1,H,W) with_grad: d consistency_loss / d multi             */
  float* grad_distil;         /* (B,1,H,W) with_grad: d distil_loss / d multi                  */
  float* grad_distil_mono;    /* (B,1,H,W) with_grad && dual_distil: d distil_loss / d mono    */
  float* partials;            /* workspace, mal_main_terms_partials_floats() floats            */
  float* sums;                /* (2): [consistency_loss, distil_loss]                          */
} mal_main_terms_args;

size_t mal_main_terms_partials_floats(int batch, int height, int width);
int mal_main_terms_forward(const mal_main_terms_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Trainer.compute_matching_mask (manydepth/trainer.py:1066-1076) fused with the nearest
 * up-sampling of lowest_cost / confidence (manydepth/networks/repdepth.py:331-336) and the
 * product at trainer.py:592-593:
 *   out = nearest(confidence) * [ (m-t)/t < 1  and  (t-m)/m < 1 ],  m = 1/nearest(lowest_cost).
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_matching_mask_args {
  int32_t batch, height, width;      /* full resolution                                       */
  int32_t low_height, low_width;     /* resolution of lowest_cost / confidence (may equal H,W)*/
  int32_t mono_is_disp;
  double min_depth, max_depth;
  const float* lowest_cost;          /* (B,h,w)                                               */
  const float* confidence;           /* (B,h,w) optional                                      */
  const float* mono;                 /* (B,1,H,W) mono depth (or disparity)                   */
  float* out_mask;                   /* (B,H,W) float {0,1} (x confidence)                    */
} mal_matching_mask_args;

int mal_matching_mask(const mal_matching_mask_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 7. The reference's layer classes on their own (forward + backward).
 *
 *   mal_backproject        BackprojectDepth.forward  manydepth/layers.py:163-168
 *                          depth (B,1,H,W), inv_K (B,4,4) -> cam points (B,4,H*W)
 *   mal_project3d          Project3D.forward         manydepth/layers.py:184-199 (MAL_CONV_MANYDEPTH),
 *                          dualrefine/layers.py:216-226 (MAL_CONV_DUALREFINE)
 *                          points (B,4,H*W), K, T (B,4,4) -> pix (B,H,W,2) [+ z (B,1,H,W) when `z`
 *                          is given: the `dc=True` second return value]
 *   mal_ssim               SSIM.forward              manydepth/layers.py:243-257
 *                          x, y (planes = B*C, H, W) -> clamp((1 - SSIM) / 2, 0, 1), same shape
 * Backward entry points take the upstream gradient and return the input gradients; grad_P is
 * d/d (K@T)[:3,:] as (B,12) (the caller chains it to T and K with two 4x4 products).
 * ------------------------------------------------------------------------------------------ */
int mal_backproject(const float* depth, const float* inv_K, int batch, int height, int width, float* out,
                    mal_stream_t stream);
int mal_backproject_backward(const float* grad_out, const float* inv_K, int batch, int height, int width,
                             float* grad_depth, mal_stream_t stream);
size_t mal_project3d_partials_floats(int batch, int height, int width);
int mal_project3d(const float* points, const float* K, const float* T, int batch, int height, int width,
                  int convention, float eps, float* pix, float* z, mal_stream_t stream);
int mal_project3d_backward(const float* points, const float* K, const float* T, const float* grad_pix,
                           const float* grad_z, int batch, int height, int width, int convention, float eps,
                           float* grad_points, float* grad_P, float* partials, mal_stream_t stream);
/* F.grid_sample(img, grid, mode="bilinear", padding_mode=border?"border":"zeros", align_corners) as the
 * reference calls it (manydepth/trainer.py:1122-1125, dualrefine/trainer.py:444-447,
 * networks/resnet_encoder.py:189, dynamicdepth/rigid_warp.py:367), with the rounding of ATen's CPU
 * kernel; the backward is w.r.t. the grid (B,Ho,Wo,2) only. */
int mal_grid_sample(const float* img, const float* grid, int batch, int channels, int height, int width,
                    int out_height, int out_width, int align_corners, int border, float* out, mal_stream_t stream);
int mal_grid_sample_backward(const float* img, const float* grid, const float* grad_out, int batch, int channels,
                             int height, int width, int out_height, int out_width, int align_corners, int border,
                             float* grad_grid, mal_stream_t stream);

/* F.interpolate(x, [out_height, out_width], mode="bilinear", align_corners=False) of `planes` planes with
 * the arithmetic of ATen's CPU kernel (manydepth/trainer.py:1093-1094 and :1176-1177, dualrefine/trainer.py:412-413,
 * dynamicdepth/trainer.py:915-916: every low-resolution disparity goes through it before disp_to_depth),
 * and its adjoint (deterministic gather, no atomics). */
int mal_upsample_bilinear(const float* in, int planes, int in_height, int in_width, int out_height, int out_width,
                          float* out, mal_stream_t stream);
int mal_upsample_bilinear_backward(const float* grad_out, int planes, int in_height, int in_width, int out_height,
                                   int out_width, float* grad_in, mal_stream_t stream);
int mal_ssim(const float* x, const float* y, int planes, int height, int width, float* out, mal_stream_t stream);
/* workspace: 4 * planes * height * width floats; grad_y may be NULL */
int mal_ssim_backward(const float* x, const float* y, const float* grad_out, int planes, int height, int width,
                      float* grad_x, float* grad_y, float* workspace, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The scalar tail of one `--distil` training step and the gradient hand-over, in one launch:
 * compute_mono_losses' / compute_main_losses' final sums (manydepth/loss_utils.py:115-127,
 * :215-281), the teacher->student accumulation (manydepth/trainer.py:624-629),
 * LossBalancing.compute_loss (loss_utils.py:303-318) and the backward of all of it down to the
 * disparity maps and poses the networks produced.  Inputs are the `sums` / loss scalars and the
 * un-normalised gradient planes the other entry points left on the device.
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_step_combine_args {
  int32_t batch, height, width;
  float smoothness;             /* opt.disparity_smoothness (1e-3)                            */
  const float* weights;         /* (2) LossBalancing.w_list on the device; NULL: no loss_blc  */
  const float* sums_teacher;    /* (4) mal_photo_forward sums of the teacher pass             */
  const float* sums_student;    /* (4) ... of the student pass                                */
  const float* smooth_teacher;  /* (1) mal_smooth_forward loss, teacher disparity             */
  const float* smooth_student;  /* (1) ... student disparity                                  */
  const float* main_sums;       /* (2) mal_main_terms_forward sums                            */
  const float* K;               /* (B,4,4)                                                    */
  const float* gd_teacher;      /* (B,1,H,W) grad_depth of the teacher pass                   */
  const float* gs_teacher;      /* (B,1,H,W) smoothness grad_disp, teacher                    */
  const float* gP_teacher;      /* (B,2,12)  grad_P of the teacher pass                       */
  const float* gd_student;      /* (B,1,H,W) grad_depth of the student pass                   */
  const float* gs_student;      /* (B,1,H,W) smoothness grad_disp, student                    */
  const float* g_cons;          /* (B,1,H,W) mal_main_terms_forward grad_cons                 */
  const float* g_distil;        /* (B,1,H,W) ... grad_distil                                  */
  const float* g_distil_mono;   /* (B,1,H,W) ... grad_distil_mono (dual_distil) or NULL       */
  float* scalars;               /* (8) [total, loss_list[0], loss_list[1], R_t, R_s, C, L_t, L_s] */
  float* grad_disp_teacher;     /* (B,1,H,W) d total / d teacher disparity                    */
  float* grad_disp_student;     /* (B,1,H,W) d total / d student disparity                    */
  float* grad_T[2];             /* (B,4,4)   d total / d cam_T_cam for frames -1,+1           */
  const float* smooth_stats;    /* (B,2,2) optional: mal_smooth_forward(defer_fix=1).stats of ONE dual call
                                   (term 0 = teacher, 1 = student); gs_teacher / gs_student are then the
                                   deferred planes and the chain through the mean-normalisation is applied here */
} mal_step_combine_args;

int mal_step_combine(const mal_step_combine_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 6. DynamicDepth forward warp: z-buffered splat of the (up-sampled) source depth into the target
 *    view, then inverse warp of the image with the splatted depth.
 *
 * Replaces dynamicdepth/rigid_warp.forward_warp (:534-597) incl. pixel2cam (:34-50),
 * cam2pix_trans (:513-530), torch_sparse.coalesce(op='max') (:577), inverse_warp (:337-373),
 * cam2pixel (:54-83).  Forward only (the reference calls it under no_grad,
 * dynamicdepth/trainer.py:494).  The small per-sample matrices are what the reference derives
 * with torch on the way: Ku_inv = inverse([K[0:2]*upscale; K[2]]) (:562-564), K_inv =
 * inverse(K) (:355), proj = K @ pose_vec2mat([t, mat2euler(R)] of inverse(pose)) (:585-589, :357-360).
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_forward_warp_args {
  int32_t batch, channels, height, width, upscale;
  const float* img;      /* (B,C,H,W)                                                        */
  const float* depth;    /* (B,1,H,W) depth of the source view                               */
  const float* pose;     /* (B,3,4)   source -> target                                       */
  const float* K;        /* (B,3,3)                                                          */
  const float* Ku_inv;   /* (B,3,3)                                                          */
  const float* K_inv;    /* (B,3,3)                                                          */
  const float* proj;     /* (B,3,4)                                                          */
  float* img_w;          /* (B,C,H,W) img warped, x valid                                    */
  float* depth_w;        /* (B,1,H,W) splatted depth, x valid                                */
  float* valid;          /* (B,1,H,W) fw_val * iw_val                                        */
  void* zbuf;            /* workspace: B*H*W uint32                                          */
} mal_forward_warp_args;

int mal_forward_warp(const mal_forward_warp_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MAL temporal hint: synthesis of the motion-compensated source images from matched instance
 * masks.
 *
 * mal_dynamic_instance replaces generate_dynamic_instance (manydepth/dyn_utils.py:38-119):
 * mask extents (:52-78), half displacement with round-half-even and the `replace` dead zone
 * (:80-100), background swap (:102-112), fill_dynamic_obj for both frames (:114-118).
 * mal_fill_dynamic_obj replaces fill_dynamic_obj (:5-36) with caller-given displacements.
 * Masks are bool tensors (1 byte per element).
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_dynamic_instance_args {
  int32_t num, channels, height, width; /* N matched instances, image (C,H,W), C <= 4          */
  int32_t replace;
  const uint8_t* mask_last;  /* (N,H,W) instance masks in the warped frame -1                  */
  const uint8_t* mask_next;  /* (N,H,W) ... frame +1, same instance order                      */
  const float* img_last;     /* (C,H,W) outputs[("color", -1, scale)][b]                       */
  const float* img_next;     /* (C,H,W) outputs[("color", +1, scale)][b]                       */
  float* ori_last;           /* (C,H,W) -> outputs[("syn", -1, scale)][b]                      */
  float* ori_next;           /* (C,H,W) -> outputs[("syn", +1, scale)][b]                      */
  int32_t* workspace;        /* 12 * N ints: extents [N][2][4] = (low, top, right, left), then
                                deltas [4][N] = (dx_last, dy_last, dx_next, dy_next)            */
} mal_dynamic_instance_args;

int mal_dynamic_instance(const mal_dynamic_instance_args* args, mal_stream_t stream);
/* Backward of mal_dynamic_instance (the reference's copies keep autograd history to the warped
 * images): `deltas` is the (4,N) block the forward left in its workspace (+ 8*N ints);
 * `flags` is an H*W byte workspace.  Uses args->mask_*, num, channels, height, width only. */
int mal_dynamic_instance_backward(const mal_dynamic_instance_args* args, const int32_t* deltas,
                                  const float* grad_ori_last, const float* grad_ori_next, float* grad_img_last,
                                  float* grad_img_next, uint8_t* flags, mal_stream_t stream);
int mal_fill_dynamic_obj(const uint8_t* mask, const int32_t* delta_x, const int32_t* delta_y, const float* source,
                         const float* img, int num, int channels, int height, int width, float* out,
                         mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MAL temporal hint inside one training step, batched over the samples (SURVEY.md section 8f item 1 as the
 * reference's Trainer runs it: manydepth/trainer.py:1078-1165 materialises outputs[("color", f, 0)], :1161-1162
 * calls dyn_utils.image_synthesis (:121-170), compute_mono_losses takes the synthesised images as candidates
 * 2 and 3 and autograd carries d loss / d syn back into the warped images, the depth and the poses).
 *
 *   mal_temporal_warp        outputs[("color", f, 0)], f = -1, +1 (BackprojectDepth, Project3D, grid_sample border)
 *   mal_temporal_pack_masks  Mask2Former-shaped (B,N,H,W) bool masks -> one 32-bit word per pixel, bit n = instance n
 *   mal_temporal_synthesis   generate_dynamic_instance (:38-119) for every sample: warped -> syn
 *   mal_temporal_backward    d loss / d syn -> d loss / d warped (optional output) -> accumulated onto `grad_depth`
 *                            and added to `grad_P` of the photometric pass that produced grad_syn
 * All four take the same argument block and read the fields they need.  Instance masks: <= 32 matched instances per
 * sample, same instance order in both frames; counts[b] = 0 leaves the sample's warped images untouched (:133-134).
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_temporal_args {
  int32_t batch, height, width;
  int32_t convention;        /* MAL_CONV_*                                                       */
  int32_t depth_is_disp;     /* 1: `depth` holds sigmoid disparity                               */
  int32_t replace;           /* generate_dynamic_instance(replace=...) dead zone (:88-96)        */
  double min_depth, max_depth;
  float eps;                 /* Project3D eps                                                    */
  const float* src[2];       /* (B,3,H,W) inputs[("color", f, 0)], f = -1, +1                    */
  const float* depth;        /* (B,1,H,W)                                                        */
  const float* K;            /* (B,4,4)                                                          */
  const float* inv_K;        /* (B,4,4)                                                          */
  const float* T[2];         /* (B,4,4)                                                          */
  float* warped[2];          /* (B,3,H,W) warp: out; synthesis: in                               */
  const uint32_t* packed_last; /* (B,H,W) bit n: instance n covers the pixel in warped frame -1  */
  const uint32_t* packed_next; /* (B,H,W) ... frame +1                                           */
  const int32_t* counts;     /* (B) matched instances per sample                                 */
  float* syn[2];             /* (B,3,H,W) synthesis: out                                         */
  int32_t* ext;              /* workspace, B*256 ints: encoded extents [B][2][4][32]             */
  int32_t* deltas;           /* (B,2,32) synthesis: out (displacement of the frame -1 copies along H, W;
                                frame +1 uses the negation); backward: in                        */
  const float* grad_syn[2];  /* (B,3,H,W) backward: d loss / d syn                               */
  float* grad_warped[2];     /* (B,3,H,W) backward, optional: d loss / d warped                  */
  float* grad_depth;         /* (B,1,H,W) backward, optional: += d loss / d depth (or disp)      */
  float* partials;           /* workspace with grad_depth: mal_temporal_partials_floats() floats */
  float* grad_P;             /* (B,2,12)  with grad_depth: += d loss / d (K@T)[:3,:]             */
} mal_temporal_args;

size_t mal_temporal_partials_floats(int batch, int height, int width);
int mal_temporal_warp(const mal_temporal_args* args, mal_stream_t stream);
int mal_temporal_pack_masks(const uint8_t* masks_last, const uint8_t* masks_next, const int32_t* counts, int batch,
                            int nmax, int height, int width, uint32_t* packed_last, uint32_t* packed_next,
                            mal_stream_t stream);
int mal_temporal_synthesis(const mal_temporal_args* args, mal_stream_t stream);
int mal_temporal_backward(const mal_temporal_args* args, mal_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * DualRefine epipolar correlation lookup, forward and backward.
 *
 * Replaces dualrefine/networks/corr.py CoordSampler: register (:11-22, the 2x2 average-pooled pyramid
 * of the source-frame features), __call__ (:24-50, used inside every DEQ iteration,
 * dualrefine/networks/depth_pose.py:433-435) and __corr__ (:52-76, num_head == 1).  For every pixel,
 * level l and epipolar candidate d the pyramid level is sampled at coords[b, :, l, d, y, x] (pixel
 * units of the level-0 map) with F.grid_sample(align_corners=False, zeros padding) and
 * out[b, (l*heads + head)*D + d, y, x] = mean over the head's channels of |fmap1 - sample|.
 * The (B, C, h, w, D) intermediates of the reference never exist.
 * ------------------------------------------------------------------------------------------ */
typedef struct mal_corr_args {
  int32_t batch, channels, height, width;  /* fmap1 / level-0 resolution                          */
  int32_t num_levels, num_samples, num_head;
  const float* fmap1;     /* (B,C,h,w)                                                            */
  const float* pyramid;   /* mal_corr_pyramid() output: level l is (B,C,h>>l,w>>l), levels back to back */
  const float* coords;    /* (B,2,L,D,h,w): x then y                                              */
  float* out;             /* (B, L*heads*D, h, w)                     [mal_corr_lookup]           */
  const float* grad_out;  /* (B, L*heads*D, h, w)                     [mal_corr_lookup_backward]  */
  float* grad_coords;     /* (B,2,L,D,h,w) optional, written                                       */
  float* grad_fmap1;      /* (B,C,h,w) optional, ACCUMULATED into: zero it first                   */
  float* grad_pyramid;    /* pyramid layout, optional, ACCUMULATED into: zero it first             */
  float* workspace;       /* optional, B*C*h*w floats, 16-byte aligned (the call zeroes it): grad_fmap1 is then
                             gathered channel-quad interleaved with 128-bit reductions (a quarter of the
                             atomics) and added into grad_fmap1 by a second small kernel                 */
} mal_corr_args;

size_t mal_corr_pyramid_floats(int batch, int channels, int height, int width, int num_levels);
int mal_corr_pyramid(const float* fmap2, int batch, int channels, int height, int width, int num_levels,
                     float* pyramid, mal_stream_t stream);
int mal_corr_lookup(const mal_corr_args* args, mal_stream_t stream);
int mal_corr_lookup_backward(const mal_corr_args* args, mal_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAL_B200_H_ */
