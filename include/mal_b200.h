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

#define MAL_ABI_VERSION 1

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
  const float* src[2];      /* (B,3,H,W) frames -1,+1: WARP: sampled; PRED: pre-warped preds */
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
} mal_photo_args;

size_t mal_photo_partials_floats(int batch, int height, int width);
int mal_photo_forward(const mal_photo_args* args, mal_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAL_B200_H_ */
