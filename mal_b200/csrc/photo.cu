// photo.cu - fused photometric loss (forward + un-normalised backward) for sm_100a.
//
// One CTA owns a 32 x 16 tile of one sample.  Phases (separated by __syncthreads):
//   0  camera constants; ONE thread asks the TMA unit (cp.async.bulk.tensor, SASS UTMALDG) for every tile
//      that comes straight from HBM - target, temporal-hint candidates, predictions (PRED mode), the
//      disparity - and everybody waits on the mbarrier.  No thread moves tile data through registers; CTAs
//      on the image border then patch the one reflected row / column ReflectionPad2d(1) needs (TMA zero-fills
//      out-of-bounds elements).  Images whose rows are not 16-byte multiples take a plain loader instead.
//   A  (WARP mode) the two warped candidates are produced in place from the staged disparity:
//      disp->depth->backproject->project->bilinear gather (border)
//      (manydepth/layers.py:14-23,163-199, trainer.py:1122-1125).  Skipped when the caller hands in the warps it has
//      materialised anyway (mal_photo_args.warped: staged by TMA like PRED mode's predictions).
//   B  per pixel of the loss region: 3x3 SSIM + L1 per candidate (layers.py:243-257,
//      loss_utils.py:46-55), min/argmin over candidates (:103), tie-break noise + automask
//      (:105-109, :27-44), multi-frame mask (:192-194); masked partial sums (:112-113).
//      With gradients, the SSIM derivative coefficients of the selected candidate are left in
//      shared memory for the ring of pixels around the tile.  (A two-pixels-per-thread variant with 64-bit
//      shared loads and shared products was measured: same instruction count, bank conflicts and spills -
//      profiles/r2_notes.md.)
//   C  (grad) per tile pixel: gather the 3x3 neighbourhood's coefficients (atomics-free SSIM
//      backward), add the L1 term, then chain through the bilinear sampler, the projection and
//      backprojection to d/d depth and d/d(K@T); per-CTA partials (one transposed warp reduction
//      of all 26 sums) go to a workspace.  d/d syn (the temporal-hint candidates) is formed in a
//      second, compact loop and only in tiles where phase B saw such a candidate win.
//   Forward-only passes score their candidates two at a time on packed fp32 pairs in phase B.
// A second tiny kernel reduces the per-CTA partials in a fixed order (deterministic).
//
// HBM traffic per pixel (WARP mode, 2+2 candidates, grad): reads 9 (target+2 src... gathers hit
// L2/L1) x 4 B x 3 ch + depth/noise/identity/mask 16 B, writes min_reproj 4 + sel 1 + grad 4 B.
#include <cstdlib>

#include "mal_math.cuh"
#include "mal_tma.cuh"

namespace mal {

constexpr int PH_TW = 32;
constexpr int PH_TH = 16;
constexpr int PH_NT = 256;
constexpr int PH_NPART = 26;  // 2 x 12 dL/dP + sum(w*reproj) + sum(w)
constexpr int PH_VW = 40;     // value-tile row pitch: the tile starts 4 columns left of the CTA's first pixel - a TMA box
                              // must start on a 16-byte boundary of the row (measured: tools/ubench/tma_check.cu)
constexpr int PH_OX = 4;

struct PhotoTile {
  int HL, HV;          // loss halo (gradient ring), value halo = HL + 1
  int VH, VN, TS;      // value tile rows, floats per plane, floats per 3-plane tile (128-byte multiple)
  int LW, LH, LN;      // loss tile dims
};
__host__ __device__ inline PhotoTile photo_tile(bool grad) {
  PhotoTile t;
  t.HL = grad ? 1 : 0;
  t.HV = t.HL + 1;
  t.VH = PH_TH + 2 * t.HV; t.VN = PH_VW * t.VH; t.TS = (3 * t.VN + 31) / 32 * 32;
  t.LW = PH_TW + 2 * t.HL; t.LH = PH_TH + 2 * t.HL; t.LN = t.LW * t.LH;
  return t;
}
constexpr int PH_SMALL = 40 + 8 * PH_NPART + 8 + 4;   // Geom (padded), reduction scratch, mbarrier + flag
inline size_t photo_smem_bytes(bool grad, bool warp, int ncand, bool avg, bool dd = false) {
  PhotoTile t = photo_tile(grad);
  size_t fl = (size_t)(1 + ncand) * t.TS + PH_SMALL + (dd ? (t.VN + 3) / 4 : 0);
  if (grad) fl += (size_t)(avg ? 19 : 10) * t.LN + (t.LN + 7) / 8 * 2;   // 9 coefficient planes, weights, selection bytes; the staged
                                                        // disparity aliases the coefficient planes
  else if (warp) fl += (size_t)2 * ((t.VN + 31) / 32 * 32);
  return fl * 4 + 128;
}

struct PhotoMaps {   // TMA descriptors of the tensors a CTA stages, one __grid_constant__ kernel parameter
  TileMap tgt, src[2], syn[2], depth, depth_b;
};

// AVG: opt.avg_reprojection (dualrefine/trainer.py:575-586, dynamicdepth/trainer.py:1044-1056): the mean instead of
// the min over the two warped candidates; both then carry half the gradient (a second set of coefficient planes).
// DD: DynamicDepth's compute_losses (dynamicdepth/trainer.py:958-975, :1006-1128) as one pass per scale.  The
// candidates are [warp(-1), warp(+1), source(-1), source(+1)] (the last two are the identity candidates of the
// automask, staged like temporal-hint images).  zero_img: a prediction's dark pixels (RGB sum < 0.1: DOMD warping
// holes) are zeroed in the prediction AND, in place, in the shared target - so call k (in the reference's order
// c0, c1, i0, i1) compares against the target zeroed wherever any of the predictions 0..k is dark: a byte per tile
// element carries those cumulative masks, the target-side window moments are formed per candidate, and the
// target as the calls leave it goes to `target_out` for the next scale.  selec_reproj (:1058-1064): where one warp
// is dark the other one's loss is taken, where both are the loss is 0.
template <bool WARP, bool GRAD, int CONV, bool LOWRES, bool SYNG, int NC, bool AVG, bool DD = false>
__global__ void __launch_bounds__(PH_NT, GRAD ? (AVG ? 2 : 3) : 4)
photo_kernel(const mal_photo_args a, const __grid_constant__ PhotoMaps maps, const int ncand, const float min_disp,
             const float disp_range, const int use_tma, const SizeDiv sdiv) {
  const PhotoTile tl = photo_tile(GRAD);
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * PH_TW, y0 = blockIdx.y * PH_TH;
  const int ox = x0 - PH_OX, oy = y0 - tl.HV;       // image coordinates of value-tile element (0, 0)
  constexpr int BS = PH_OX - 1 - (GRAD ? 1 : 0);    // loss column lx's 3x3 window starts at value column lx + BS
  const int H = a.height, W = a.width;
  const size_t HW = (size_t)H * W;
  // low-resolution disparity, up-sampled on the fly (mal_photo_args.depth_height / depth_width)
  // (a template parameter: the extra addressing costs registers and with them a resident CTA)
  constexpr bool lowres = WARP && LOWRES;
  constexpr bool dep_tile = WARP && !LOWRES;        // the disparity tile is staged in shared memory
  const size_t dhw = (size_t)a.depth_height * a.depth_width;
  const float up_sy = lowres ? up_scale(a.depth_height, H) : 1.0f, up_sx = lowres ? up_scale(a.depth_width, W) : 1.0f;

  // ---- shared memory carve-up (tiles first: TMA destinations are 128-byte aligned) ------------------------
  float* smem;
  {
    unsigned char* raw = dyn_smem();
    smem = reinterpret_cast<float*>(raw + ((128 - ((uintptr_t)raw & 127)) & 127));
  }
  float* sy = smem;                                   // [3][VN] target
  float* sx = sy + tl.TS;                             // [ncand][TS]: [3][VN] per candidate
  float* after = sx + (size_t)ncand * tl.TS;
  constexpr int NCOEF = AVG ? 18 : 9;
  float* coef = after;                                // [9][LN] (GRAD; AVG: a second set for candidate 1)
  float* lw = coef + NCOEF * tl.LN;                   // [LN] weights (GRAD)
  signed char* lsel = reinterpret_cast<signed char*>(lw + tl.LN);   // [LN] selected candidate or -1
  float* dep = after;                                 // [VN] disparity tile (aliases coef: dead before phase B)
  float* dep_b = dep + (tl.VN + 31) / 32 * 32;        // [VN] second disparity (ensemble pass, never with GRAD)
  float* small = after + (GRAD ? (NCOEF + 1) * tl.LN + (tl.LN + 7) / 8 * 2 : (WARP ? 2 * ((tl.VN + 31) / 32 * 32) : 0));
  unsigned char* dm = reinterpret_cast<unsigned char*>(small + PH_SMALL);   // [VN] (DD) bits 0-3: target zeroed for
                                                                            // call k; bits 4, 5: warp -1 / +1 dark
  Geom* geom = reinterpret_cast<Geom*>(small);
  float* red = small + 40;                            // [8][PH_NPART]
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(red + 8 * PH_NPART + 8);
  int* syn_flag = reinterpret_cast<int*>(mbar + 1);   // (SYNG) some pixel of the tile selected a temporal-hint candidate

  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) {
    // the finalize kernel's ticket (see photo_finalize_kernel); it launches after this kernel ends
    float* tail = a.partials + (size_t)gridDim.x * gridDim.y * gridDim.z * PH_NPART;
    reinterpret_cast<unsigned*>(tail + (size_t)gridDim.z * 2)[gridDim.z] = 0u;
  }
  // ---- phase 0: camera constants, TMA requests ------------------------------------------------------------
  if (tid == 0 && use_tma) mbar_init(mbar, 1);
  if (SYNG && tid == 0) *syn_flag = 0;
  if (WARP) {
    if (tid >= 32 && tid < 56) {
      int f = (tid - 32) / 12, e = (tid - 32) % 12;
      geom->P[f][e] = kt_entry(a.K + b * 16, a.T[f] + b * 16, e / 4, e % 4);
    } else if (tid >= 64 && tid < 73) {
      int e = tid - 64;
      geom->iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
    }
  }
  __syncthreads();
  const float* tgt = a.target + (size_t)b * 3 * HW;
  const bool with_syn = NC > 2 && ncand > 2;
  // WARP mode with ready-made warped sources: staged like PRED mode's predictions, phase A is skipped
  const bool pre = WARP && !DD && a.warped[0] != nullptr;
  if (use_tma) {
    if (tid == 0) {
      const unsigned tile_bytes = (unsigned)(3 * tl.VN * 4), plane_bytes = (unsigned)(tl.VN * 4);
      unsigned bytes = tile_bytes;
      if (!WARP || pre) bytes += tile_bytes * (unsigned)min(ncand, 2);
      if (with_syn) bytes += 2 * tile_bytes;
      if (dep_tile && !pre) bytes += plane_bytes * (a.depth_b ? 2u : 1u);
      mbar_expect_tx(mbar, bytes);
      tma_load_3d(sy, &maps.tgt, ox, oy, b * 3, mbar);
      if (!WARP || pre)
        for (int f = 0; f < min(ncand, 2); f++) tma_load_3d(sx + (size_t)f * tl.TS, &maps.src[f], ox, oy, b * 3, mbar);
      if (with_syn)
        for (int k = 0; k < 2; k++) tma_load_3d(sx + (size_t)(2 + k) * tl.TS, &maps.syn[k], ox, oy, b * 3, mbar);
      if (dep_tile && !pre) {
        tma_load_3d(dep, &maps.depth, ox, oy, b, mbar);
        if (a.depth_b) tma_load_3d(dep_b, &maps.depth_b, ox, oy, b, mbar);
      }
    }
    mbar_wait(mbar, 0);
#ifdef MAL_EMU
    __syncthreads();   // the emulated copies ran inside thread 0
#endif
  }
  // Elements outside the image: TMA delivered zeros; ReflectionPad2d(1) wants the mirrored pixel (and tiles
  // hanging over the right / bottom edge want any in-range value).  Only border CTAs have such elements.
  // Without TMA the same loop stages every element.
  const bool over = ox < 0 || oy < 0 || ox + PH_VW > W || oy + tl.VH > H;
  if (!use_tma || over) {
    for (int i = tid; i < tl.VN; i += PH_NT) {
      const int ty = i / PH_VW, tx = i - ty * PH_VW;
      const int gy = oy + ty, gx = ox + tx;
      if (use_tma && gy >= 0 && gy < H && gx >= 0 && gx < W) continue;
      const int ry = min(max(reflect_index(gy, H), 0), H - 1), rx = min(max(reflect_index(gx, W), 0), W - 1);
      const size_t o = (size_t)ry * W + rx;
#pragma unroll
      for (int c = 0; c < 3; c++) sy[c * tl.VN + i] = __ldg(tgt + c * HW + o);
      if (with_syn) {
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
          for (int c = 0; c < 3; c++)
            sx[(size_t)(2 + k) * tl.TS + c * tl.VN + i] = __ldg(a.syn[k] + ((size_t)b * 3 + c) * HW + o);
      }
      if (!WARP || pre) {
#pragma unroll
        for (int f = 0; f < 2; f++)
          if (f < ncand) {
            const float* cand = pre ? a.warped[f] : a.src[f];
#pragma unroll
            for (int c = 0; c < 3; c++) sx[(size_t)f * tl.TS + c * tl.VN + i] = __ldg(cand + ((size_t)b * 3 + c) * HW + o);
          }
      }
      if (dep_tile && !pre) {
        dep[i] = __ldg(a.depth + (size_t)b * HW + o);
        if (a.depth_b) dep_b[i] = __ldg(a.depth_b + (size_t)b * HW + o);
      }
    }
    if (WARP) __syncthreads();   // phase A reads the disparity tile through a different index map (CTA-uniform branch)
  }

  // ---- phase A: warped candidates -------------------------------------------------------------------------
  if (WARP && !pre) {
    constexpr int NEED = PH_TW + 2 * (GRAD ? 1 : 0) + 2;   // only the columns some 3x3 window reaches
    for (int j = tid; j < NEED * tl.VH; j += PH_NT) {
      const int ty = j / NEED, tx = BS + j - ty * NEED;
      const int i = ty * PH_VW + tx;
      const int ry = min(max(reflect_index(oy + ty, H), 0), H - 1), rx = min(max(reflect_index(ox + tx, W), 0), W - 1);
      float dv;
      if (lowres) {   // F.interpolate(disp, [H, W], bilinear) read straight from the low-resolution plane
        const UpAxis ay = up_axis(ry, a.depth_height, up_sy), ax = up_axis(rx, a.depth_width, up_sx);
        dv = upsample_at(a.depth + (size_t)b * dhw, a.depth_width, ay, ax);
        if (a.depth_b) dv = xmul(xadd(dv, upsample_at(a.depth_b + (size_t)b * dhw, a.depth_width, ay, ax)), 0.5f);
      } else {
        dv = dep[i];
        if (a.depth_b) dv = xmul(xadd(dv, dep_b[i]), 0.5f);   // (a + b) / 2.0
      }
      if (a.depth_is_disp) dv = xdiv(1.0f, xadd(min_disp, xmul(disp_range, dv)));
      Ray ray = pixel_ray(geom->iK, (float)rx, (float)ry);
      unsigned dark = 0;
#pragma unroll
      for (int f = 0; f < 2; f++) {
        Sample s = project_pixel<CONV>(geom->P[f], ray, dv, a.eps, H, W, &sdiv);
        Taps t = make_taps(s.ix, s.iy, H, W);
        const float* src = a.src[f] + (size_t)b * 3 * HW;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; c++) v[c] = bilinear(src + c * HW, t);
        if (DD && (a.zero_img || a.selec_reproj) && xadd(xadd(v[0], v[1]), v[2]) < 0.1f) {   // pred.sum(1) < 0.1
          dark |= 1u << f;
          if (a.zero_img) v[0] = v[1] = v[2] = 0.0f;                              // pred[mask] = 0
        }
#pragma unroll
        for (int c = 0; c < 3; c++) sx[(size_t)f * tl.TS + c * tl.VN + i] = v[c];
      }
      if (DD) {
        // the identity candidates (un-warped sources, staged by TMA / the plain loader) take part in the zeroing too
        unsigned cum = a.zero_img ? dark : 0u;            // bit k: the target is zero for call k (c0, c1, i0, i1)
        cum |= (cum & 1u) << 1;
        if (NC > 2 && ncand > 2) {
#pragma unroll
          for (int k = 0; k < 2; k++) {
            float* X = sx + (size_t)(2 + k) * tl.TS + i;
            bool dk = false;
            if (a.zero_img && xadd(xadd(X[0], X[tl.VN]), X[2 * tl.VN]) < 0.1f) {
              dk = true;
              X[0] = X[tl.VN] = X[2 * tl.VN] = 0.0f;
            }
            if (dk || (cum >> (1 + k)) & 1u) cum |= 1u << (2 + k);
          }
        }
        dm[i] = (unsigned char)(cum | (dark << 4));
      }
    }
  }
  __syncthreads();

  // ---- phase B: losses, selection, weights -----------------------------------------------
  const bool automask = a.identity_min != nullptr;
  float acc_loss = 0.0f, acc_w = 0.0f;
  // Work is handed out in warp-sized tasks so that the 32 lanes of a task read 32 consecutive
  // words of a value-tile row (conflict-free): LH row tasks cover columns 0..31 of the loss
  // region; the LW-32 leftover columns (the gradient halo) are packed column-major into extra tasks.
  const int extra_items = (tl.LW - 32) * tl.LH;
  const int ntasks = tl.LH + (extra_items + 31) / 32;
  for (int task = tid >> 5; task < ntasks; task += PH_NT / 32) {
    int ly, lx;
    if (task < tl.LH) {
      ly = task; lx = tid & 31;
    } else {
      const int j = (task - tl.LH) * 32 + (tid & 31);
      if (j >= extra_items) continue;
      ly = j % tl.LH; lx = 32 + j / tl.LH;
    }
    const int i = ly * tl.LW + lx;
    int gy = y0 - tl.HL + ly, gx = x0 - tl.HL + lx;
    bool in_img = gy >= 0 && gy < H && gx >= 0 && gx < W;
    if (!in_img) {
      if (GRAD) { lsel[i] = -1; lw[i] = 0.0f; }
      continue;
    }
    const int vc = (ly + 1) * PH_VW + lx + BS + 1;   // window centre in a value plane
    // per-pixel planes are fetched now so their latency hides behind the SSIM arithmetic
    const size_t po = (size_t)b * HW + (size_t)gy * W + gx;
    float p_ident = 0.0f, p_noise = 0.0f, p_mask = 1.0f;
    if (automask) { p_ident = __ldg(a.identity_min + po); p_noise = __ldg(a.noise + po); }
    if (DD && NC > 2 && ncand > 2 && a.noise) p_noise = __ldg(a.noise + po);
    if (a.pixel_mask) p_mask = __ldg(a.pixel_mask + po);
    float ssum[NC], lsum[NC];   // NC: compiled-in candidate capacity (2 or 4); ncand <= NC are live
    float cf0[9], cf1[9];
    unsigned dmw[DD ? 9 : 1];   // (DD) the window's zeroing masks
    if (DD) {
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) dmw[DD ? dy * 3 + dx : 0] = dm[vc + (dy - 1) * PH_VW + dx - 1];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float yw0[9];
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) yw0[dy * 3 + dx] = sy[c * tl.VN + vc + (dy - 1) * PH_VW + dx - 1];
      float mu_y = 0.f, eyy = 0.f;
      if (!a.no_ssim && !DD) {
        mu_y = xdivc<9>(sum9(yw0));
        eyy = xdivc<9>(sum9_prod(yw0, yw0));
      }
      // Forward-only passes score the candidates two at a time on packed fp32 pairs (FFMA2 / FADD2 / FMUL2: the same
      // IEEE operations per half, mal_common.cuh): these passes are the most issue-bound ones (78 % of the issue
      // slots in the split-min pass, 71 % of its instructions in this loop).
      constexpr bool PAIRS = !GRAD && !DD;
      if (PAIRS) {
        const float mu_yy = xmul(mu_y, mu_y), sig_y = xsub(eyy, mu_yy);
#pragma unroll
        for (int k = 0; k + 1 < NC; k += 2) {
          if (k + 1 < ncand) {
            const float* X0 = sx + (size_t)k * tl.TS + c * tl.VN + vc;
            const float* X1 = X0 + tl.TS;
            pk2 xw2[9];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
              for (int dx = 0; dx < 3; dx++)
                xw2[dy * 3 + dx] = pack2(X0[(dy - 1) * PH_VW + dx - 1], X1[(dy - 1) * PH_VW + dx - 1]);
            const pk2 dl = x2sub(dup2(yw0[4]), xw2[4]);
            const float l1a = fabsf(lo2(dl)), l1b = fabsf(hi2(dl));
            lsum[k] = (c == 0) ? l1a : xadd(lsum[k], l1a);
            lsum[k + 1] = (c == 0) ? l1b : xadd(lsum[k + 1], l1b);
            if (!a.no_ssim) {
              pk2 yw2[9];
#pragma unroll
              for (int q = 0; q < 9; q++) yw2[q] = dup2(yw0[q]);
              const pk2 mu_x = x2divc<9>(sum9(xw2));
              const pk2 exx = x2divc<9>(sum9_prod(xw2, xw2));
              const pk2 exy = x2divc<9>(sum9_prod(xw2, yw2));
              SsimTerms t0, t1;
              ssim_terms2(mu_x, mu_y, mu_yy, sig_y, exx, exy, t0, t1);
              const float sa = clamp01(t0.v), sb = clamp01(t1.v);
              ssum[k] = (c == 0) ? sa : xadd(ssum[k], sa);
              ssum[k + 1] = (c == 0) ? sb : xadd(ssum[k + 1], sb);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NC; k++) {
        if (PAIRS && (k | 1) < ncand) continue;   // scored above as half of a pair
        if (k < ncand) {
          float yw[9];   // the target as call k sees it (DD: zeroed where a prediction up to k is dark)
#pragma unroll
          for (int q = 0; q < 9; q++) yw[q] = (DD && ((dmw[DD ? q : 0] >> k) & 1u)) ? 0.0f : yw0[q];
          if (DD && !a.no_ssim) {
            mu_y = xdivc<9>(sum9(yw));
            eyy = xdivc<9>(sum9_prod(yw, yw));
          }
          const float* X = sx + (size_t)k * tl.TS + c * tl.VN + vc;
          float xw[9];
#pragma unroll
          for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) xw[dy * 3 + dx] = X[(dy - 1) * PH_VW + dx - 1];
          float l1 = fabsf(xsub(yw[4], xw[4]));
          lsum[k] = (c == 0) ? l1 : xadd(lsum[k], l1);
          if (!a.no_ssim) {
            float mu_x = xdivc<9>(sum9(xw));
            float exx = xdivc<9>(sum9_prod(xw, xw));
            float exy = xdivc<9>(sum9_prod(xw, yw));
            SsimTerms t = ssim_terms(mu_x, mu_y, exx, eyy, exy);
            float sv = clamp01(t.v);
            ssum[k] = (c == 0) ? sv : xadd(ssum[k], sv);
            if (GRAD && k < 2) {
              float al, be, ga;
              ssim_coefs(t, al, be, ga);
              if (k == 0) { cf0[c * 3] = al; cf0[c * 3 + 1] = be; cf0[c * 3 + 2] = ga; }
              else        { cf1[c * 3] = al; cf1[c * 3 + 1] = be; cf1[c * 3 + 2] = ga; }
            }
          }
        }
      }
    }
    float rmin = 0.f, rmin_b = 0.f, lk01[2] = {0.f, 0.f};
    int idx = 0;
    const bool split = NC > 2 && !GRAD && a.min_reproj_b != nullptr;   // two independent 2-candidate mins in one pass
#pragma unroll
    for (int k = 0; k < NC; k++) {
      if (k < ncand) {
        float l1m = xdivc<3>(lsum[k]);
        float lk = a.no_ssim ? l1m : xadd(xmul(0.85f, xdivc<3>(ssum[k])), xmul(0.15f, l1m));
        if (DD && k < 2) lk01[k] = lk;
        if (AVG) rmin = (k == 0) ? lk : xmul(xadd(rmin, lk), 0.5f);   // .mean(1) of two: (l0 + l1) / 2
        else if (NC > 2 && (split || DD) && k >= 2) { if (k == 2 || lk < rmin_b) rmin_b = lk; }   // second min: candidates 2, 3
        else if (k == 0 || lk < rmin) { rmin = lk; idx = k; }
      }
    }
    bool both_dark = false;
    if (DD && a.selec_reproj) {   // :1058-1064, in the reference's order of assignments
      const bool d0 = (dmw[DD ? 4 : 0] >> 4) & 1u, d1 = (dmw[DD ? 4 : 0] >> 5) & 1u;
      if (d0) { rmin = lk01[1]; idx = 1; }
      if (d1) { rmin = lk01[0]; idx = 0; }
      if (d0 && d1) { rmin = 0.0f; both_dark = true; }
    }
    int mbit = 1;
    if (automask) {
      float ident = xadd(p_ident, xmul(p_noise, 0.00001f));
      mbit = (ident < rmin) ? 0 : 1;  // argmin([reproj, identity]) == 0, first index wins ties
    }
    if (DD && NC > 2 && ncand > 2) {   // the identity candidates were scored in this pass
      const float ident = xadd(rmin_b, xmul(p_noise, 0.00001f));
      mbit = (ident < rmin) ? 0 : 1;
      if (a.ignore_automask) mbit = 1;   // is_multi: the automask is replaced by the consistency mask (:1077-1085)
    }
    float w = (float)mbit;
    if (a.pixel_mask) w = xmul(w, p_mask);
    if (a.sample_mask) w = xmul(w, xsub(1.0f, __ldg(a.sample_mask + b)));
    const bool interior = ly >= tl.HL && ly < tl.HL + PH_TH && lx >= tl.HL && lx < tl.HL + PH_TW;
    if (interior) {
      if (a.min_reproj) a.min_reproj[po] = rmin;
      if (NC > 2 && split) a.min_reproj_b[po] = rmin_b;
      if (a.selection) a.selection[po] = (uint8_t)(idx | (mbit << 7));
      if (a.weight) a.weight[po] = w;
      if (DD && a.target_out) {   // the target as this scale's calls leave it (zero_img mutates it in place)
        const bool z = (dmw[DD ? 4 : 0] >> (ncand > 2 ? 3 : 1)) & 1u;
#pragma unroll
        for (int c = 0; c < 3; c++)
          a.target_out[((size_t)b * 3 + c) * HW + (size_t)gy * W + gx] = z ? 0.0f : sy[c * tl.VN + vc];
      }
      acc_loss += xmul(rmin, w);
      acc_w += w;
    }
    if (GRAD) {
      // warped candidates always carry a gradient; the temporal-hint candidates when the caller asked
      // for d/d syn (autograd then carries it into the warped images image_synthesis copied from)
      const bool syn_grad = (SYNG || !WARP) && a.grad_syn[0] != nullptr;
      const bool live = (idx < 2 || syn_grad) && w != 0.0f && !(DD && both_dark);
      lsel[i] = live ? idx : -1;
      lw[i] = AVG ? w * 0.5f : w;
      if (SYNG && live && idx > 1) *syn_flag = 1;   // (every writer stores 1: a benign race)
      if (AVG && live && !a.no_ssim) {
        const float sc = w * (0.5f * 0.85f / 27.0f);   // both candidates, half the weight each
#pragma unroll
        for (int j = 0; j < 9; j++) { coef[j * tl.LN + i] = cf0[j] * sc; coef[(9 + j) * tl.LN + i] = cf1[j] * sc; }
      } else if (live && !a.no_ssim) {
        const float sc = w * (0.85f / 27.0f);  // weight * 0.85 * (1/3 channels) * (1/9 window)
        if (NC <= 2 || idx < 2) {
#pragma unroll
          for (int j = 0; j < 9; j++) coef[j * tl.LN + i] = (idx == 0 ? cf0[j] : cf1[j]) * sc;
        } else {
          // rare path: re-derive the SSIM terms of the selected temporal-hint candidate
          for (int c = 0; c < 3; c++) {
            float yw[9], xw[9];
            const float* X = sx + (size_t)idx * tl.TS + c * tl.VN + vc;
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
              for (int dx = 0; dx < 3; dx++) {
                yw[dy * 3 + dx] = sy[c * tl.VN + vc + (dy - 1) * PH_VW + dx - 1];
                xw[dy * 3 + dx] = X[(dy - 1) * PH_VW + dx - 1];
              }
            SsimTerms t = ssim_terms(xdivc<9>(sum9(xw)), xdivc<9>(sum9(yw)), xdivc<9>(sum9_prod(xw, xw)),
                                     xdivc<9>(sum9_prod(yw, yw)), xdivc<9>(sum9_prod(xw, yw)));
            float al, be, ga;
            ssim_coefs(t, al, be, ga);
            coef[(c * 3) * tl.LN + i] = al * sc;
            coef[(c * 3 + 1) * tl.LN + i] = be * sc;
            coef[(c * 3 + 2) * tl.LN + i] = ga * sc;
          }
        }
      }
    }
  }

  // ---- phase C: gradients -------------------------------------------------------------------
  float gP[24];
#pragma unroll
  for (int j = 0; j < 24; j++) gP[j] = 0.0f;
  if (GRAD) {
    __syncthreads();
    for (int i = tid; i < PH_TW * PH_TH; i += PH_NT) {
      int qy = i / PH_TW, qx = i - qy * PH_TW;
      int gy = y0 + qy, gx = x0 + qx;
      if (gy >= H || gx >= W) continue;
      const int lc = (qy + tl.HL) * tl.LW + qx + tl.HL;
      const int vc = (qy + tl.HV) * PH_VW + qx + PH_OX;
      // d S / d (candidate image k) at this pixel, for a candidate that is given as an image: the
      // predictions of PRED mode, and the temporal-hint candidates (k >= 2) of either mode
      auto image_grad = [&](int k, float* out) {
        const size_t pq = (size_t)gy * W + gx;
        float g[3] = {0.f, 0.f, 0.f};
        float xq[3], yq[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { xq[c] = sx[(size_t)k * tl.TS + c * tl.VN + vc]; yq[c] = sy[c * tl.VN + vc]; }
        if (!a.no_ssim) {
          for (int dy = -1; dy <= 1; dy++) {
            int py = gy + dy;
            if (py < 0 || py >= H) continue;
            float my = ((py == 0 && dy == -1) || (py == H - 1 && dy == 1)) ? 2.0f : 1.0f;
            for (int dx = -1; dx <= 1; dx++) {
              int px = gx + dx;
              if (px < 0 || px >= W) continue;
              float m = my * (((px == 0 && dx == -1) || (px == W - 1 && dx == 1)) ? 2.0f : 1.0f);
              int li = lc + dy * tl.LW + dx;
              if (lsel[li] != k) continue;
#pragma unroll
              for (int c = 0; c < 3; c++)
                g[c] += m * (coef[(c * 3) * tl.LN + li] + 2.0f * xq[c] * coef[(c * 3 + 1) * tl.LN + li] +
                             yq[c] * coef[(c * 3 + 2) * tl.LN + li]);
            }
          }
        }
        if (lsel[lc] == k) {
          float wl = lw[lc] * (a.no_ssim ? (1.0f / 3.0f) : (0.15f / 3.0f));
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float d = yq[c] - xq[c];
            g[c] += d > 0.f ? -wl : (d < 0.f ? wl : 0.f);
          }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) out[((size_t)b * 3 + c) * HW + pq] = g[c];
      };
      if (!WARP) {
        // PRED mode: d S / d prediction for every candidate that can be selected, one at a time
        for (int k = 0; k < ncand; k++) {
          float* out = k < 2 ? a.grad_pred[k] : a.grad_syn[k - 2];
          if (out != nullptr) image_grad(k, out);
        }
        continue;
      }
      // (SYNG is a template parameter: the extra code costs the plain teacher pass 3% through the
      // instruction cache even when it never runs)
      // with SYNG the same neighbourhood walk also collects d S / d syn (the temporal-hint candidates 2, 3): one
      // selection load per neighbour decides which of the four accumulators it feeds
      const unsigned dmc = DD ? dm[vc] : 0u;
      float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f};
      float xq0[3], xq1[3], yq[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        xq0[c] = sx[c * tl.VN + vc];
        xq1[c] = ncand > 1 ? sx[(size_t)tl.TS + c * tl.VN + vc] : 0.0f;
        yq[c] = sy[c * tl.VN + vc];
      }
      if (!a.no_ssim) {
#pragma unroll
        for (int dy = -1; dy <= 1; dy++) {
          int py = gy + dy;
          if (py < 0 || py >= H) continue;
          // ReflectionPad2d(1): a border row sees its inner neighbour twice
          float my = ((py == 0 && dy == -1) || (py == H - 1 && dy == 1)) ? 2.0f : 1.0f;
#pragma unroll
          for (int dx = -1; dx <= 1; dx++) {
            int px = gx + dx;
            if (px < 0 || px >= W) continue;
            float m = my * (((px == 0 && dx == -1) || (px == W - 1 && dx == 1)) ? 2.0f : 1.0f);
            int li = lc + dy * tl.LW + dx;
            int s = lsel[li];
            // not live, or (SYNG) a temporal-hint candidate: its gradient goes to grad_syn in the loop after this
            // one, not into depth / pose
            if (SYNG ? (unsigned)s > 1u : s < 0) continue;
            if (DD) {   // the target value at this pixel as the selected call saw it
              const bool z = (dmc >> s) & 1u;
#pragma unroll
              for (int c = 0; c < 3; c++) {
                const float xq = s == 0 ? xq0[c] : xq1[c];
                const float v = m * (coef[(c * 3) * tl.LN + li] + 2.0f * xq * coef[(c * 3 + 1) * tl.LN + li] +
                                     (z ? 0.0f : yq[c]) * coef[(c * 3 + 2) * tl.LN + li]);
                if (s == 0) g0[c] += v; else g1[c] += v;
              }
              continue;
            }
            if (AVG) {
#pragma unroll
              for (int c = 0; c < 3; c++) {
                g0[c] += m * (coef[(c * 3) * tl.LN + li] + 2.0f * xq0[c] * coef[(c * 3 + 1) * tl.LN + li] +
                              yq[c] * coef[(c * 3 + 2) * tl.LN + li]);
                g1[c] += m * (coef[(9 + c * 3) * tl.LN + li] + 2.0f * xq1[c] * coef[(9 + c * 3 + 1) * tl.LN + li] +
                              yq[c] * coef[(9 + c * 3 + 2) * tl.LN + li]);
              }
              continue;
            }
#pragma unroll
            for (int c = 0; c < 3; c++) {
              float al = coef[(c * 3) * tl.LN + li], be = coef[(c * 3 + 1) * tl.LN + li],
                    ga = coef[(c * 3 + 2) * tl.LN + li];
              float xq = s == 0 ? xq0[c] : xq1[c];
              float v = m * (al + 2.0f * xq * be + yq[c] * ga);
              if (s == 0) g0[c] += v; else g1[c] += v;
            }
          }
        }
      }
      {
        int s = lsel[lc];
        if (s >= 0 && (!SYNG || s <= 1)) {
          float wl = lw[lc] * (a.no_ssim ? (1.0f / 3.0f) : (0.15f / 3.0f));
#pragma unroll
          for (int c = 0; c < 3; c++) {
            if (AVG) {   // lw already carries the 1/2
              const float d0 = yq[c] - xq0[c], d1 = yq[c] - xq1[c];
              g0[c] += d0 > 0.f ? -wl : (d0 < 0.f ? wl : 0.f);
              g1[c] += d1 > 0.f ? -wl : (d1 < 0.f ? wl : 0.f);
              continue;
            }
            float d = ((DD && ((dmc >> s) & 1u)) ? 0.0f : yq[c]) - (s == 0 ? xq0[c] : xq1[c]);      // target - pred
            float sg = d > 0.f ? -wl : (d < 0.f ? wl : 0.f);   // d|t-p|/dp = -sign(t-p)
            if (s == 0) g0[c] += sg; else g1[c] += sg;
          }
        }
      }
      if (DD && a.zero_img) {   // pred[mask] = 0 overwrites the warped pixel: no gradient reaches a dark one
        if ((dmc >> 4) & 1u) g0[0] = g0[1] = g0[2] = 0.0f;
        if ((dmc >> 5) & 1u) g1[0] = g1[1] = g1[2] = 0.0f;
      }
      const size_t po = (size_t)gy * W + gx;
      {
        float dv_in;
        if (lowres) {
          const UpAxis ay = up_axis(gy, a.depth_height, up_sy), ax = up_axis(gx, a.depth_width, up_sx);
          dv_in = upsample_at(a.depth + (size_t)b * dhw, a.depth_width, ay, ax);
        } else {
          dv_in = __ldg(a.depth + (size_t)b * HW + po);
        }
        float dv = a.depth_is_disp ? xdiv(1.0f, xadd(min_disp, xmul(disp_range, dv_in))) : dv_in;
        Ray ray = pixel_ray(geom->iK, (float)gx, (float)gy);
        float cam[3] = {dv * ray.x, dv * ray.y, dv * ray.z};
        float gdepth = 0.0f;
#pragma unroll
        for (int f = 0; f < 2; f++) {
          const float* g = f == 0 ? g0 : g1;
          if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f) continue;
          const float* P = geom->P[f];
          Sample s = project_pixel<CONV>(P, ray, dv, a.eps, H, W, &sdiv);
          Taps t = make_taps(s.ix, s.iy, H, W);
          const float* src = a.src[f] + (size_t)b * 3 * HW;
          float gix = 0.f, giy = 0.f;
#pragma unroll
          for (int c = 0; c < 3; c++) {
            float v00, v01, v10, v11;
            bilinear(src + c * HW, t, &v00, &v01, &v10, &v11);
            gix += g[c] * ((v01 - v00) * (1.0f - t.ty) + (v11 - v10) * t.ty);
            giy += g[c] * ((v10 - v00) * (1.0f - t.tx) + (v11 - v01) * t.tx);
          }
          gix *= s.gmx;  // d(ix)/d(px) == 1 for both conventions (normalise o unnormalise)
          giy *= s.gmy;
          float iz = 1.0f / s.Zp;
          float gX = gix * iz, gY = giy * iz;
          float gZ = -(gX * s.X + gY * s.Y) * iz;
          gdepth += gX * (P[0] * ray.x + P[1] * ray.y + P[2] * ray.z) +
                    gY * (P[4] * ray.x + P[5] * ray.y + P[6] * ray.z) +
                    gZ * (P[8] * ray.x + P[9] * ray.y + P[10] * ray.z);
          float* gp = gP + f * 12;
          gp[0] += gX * cam[0]; gp[1] += gX * cam[1]; gp[2] += gX * cam[2]; gp[3] += gX;
          gp[4] += gY * cam[0]; gp[5] += gY * cam[1]; gp[6] += gY * cam[2]; gp[7] += gY;
          gp[8] += gZ * cam[0]; gp[9] += gZ * cam[1]; gp[10] += gZ * cam[2]; gp[11] += gZ;
        }
        if (a.depth_is_disp) gdepth *= -disp_range * dv * dv;  // d depth / d disp
        a.grad_depth[(size_t)b * HW + po] = gdepth;
      }
    }
    // d S / d syn (the temporal-hint candidates 2, 3), in its own loop: a temporal-hint candidate wins only inside
    // and next to an instance, so most tiles have no such pixel (a CTA-uniform flag raised in phase B) and only
    // write their zeros; the others repeat the neighbourhood walk for the pixels that selected candidate 2 or 3.
    // Keeping this out of the loop above keeps that loop the plain teacher pass's.
    if (SYNG && WARP && a.grad_syn[0] != nullptr) {
      const bool has_syn = *syn_flag != 0;
      for (int i = tid; i < PH_TW * PH_TH; i += PH_NT) {
        const int qy = i / PH_TW, qx = i - qy * PH_TW;
        const int gy = y0 + qy, gx = x0 + qx;
        if (gy >= H || gx >= W) continue;
        float g2[3] = {0.f, 0.f, 0.f}, g3[3] = {0.f, 0.f, 0.f};
        if (has_syn) {
          const int lc = (qy + tl.HL) * tl.LW + qx + tl.HL;
          const int vc = (qy + tl.HV) * PH_VW + qx + PH_OX;
          float xq2[3], xq3[3], yq[3];
#pragma unroll
          for (int c = 0; c < 3; c++) {
            xq2[c] = sx[(size_t)2 * tl.TS + c * tl.VN + vc];
            xq3[c] = sx[(size_t)3 * tl.TS + c * tl.VN + vc];
            yq[c] = sy[c * tl.VN + vc];
          }
          if (!a.no_ssim) {
            for (int dy = -1; dy <= 1; dy++) {
              const int py = gy + dy;
              if (py < 0 || py >= H) continue;
              const float my = ((py == 0 && dy == -1) || (py == H - 1 && dy == 1)) ? 2.0f : 1.0f;
              for (int dx = -1; dx <= 1; dx++) {
                const int px = gx + dx;
                if (px < 0 || px >= W) continue;
                const float m = my * (((px == 0 && dx == -1) || (px == W - 1 && dx == 1)) ? 2.0f : 1.0f);
                const int li = lc + dy * tl.LW + dx;
                const int s = lsel[li];
                if (s < 2) continue;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                  const float xq = s == 2 ? xq2[c] : xq3[c];
                  const float v = m * (coef[(c * 3) * tl.LN + li] + 2.0f * xq * coef[(c * 3 + 1) * tl.LN + li] +
                                       yq[c] * coef[(c * 3 + 2) * tl.LN + li]);
                  if (s == 2) g2[c] += v; else g3[c] += v;
                }
              }
            }
          }
          const int s = lsel[lc];
          if (s > 1) {
            const float wl = lw[lc] * (a.no_ssim ? (1.0f / 3.0f) : (0.15f / 3.0f));
#pragma unroll
            for (int c = 0; c < 3; c++) {
              const float d = yq[c] - (s == 2 ? xq2[c] : xq3[c]);
              const float sg = d > 0.f ? -wl : (d < 0.f ? wl : 0.f);
              if (s == 2) g2[c] += sg; else g3[c] += sg;
            }
          }
        }
        const size_t pq = (size_t)gy * W + gx;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          a.grad_syn[0][((size_t)b * 3 + c) * HW + pq] = g2[c];
          a.grad_syn[1][((size_t)b * 3 + c) * HW + pq] = g3[c];
        }
      }
    }
  }

  // ---- per-CTA partials ---------------------------------------------------------------------
  const int warp = tid >> 5, lane = tid & 31;
  if (GRAD && WARP) {
    // the 24 d/d(K@T) sums and the two loss sums of the warp in one transposed reduction (31 shuffles instead of 130)
    float all[PH_NPART];
#pragma unroll
    for (int j = 0; j < 24; j++) all[j] = gP[j];
    all[24] = acc_loss;
    all[25] = acc_w;
    const float v = warp_sum_transposed<PH_NPART>(all, lane);
    if (lane < PH_NPART) red[warp * PH_NPART + lane] = v;
  } else {
    if (lane < 24) red[warp * PH_NPART + lane] = 0.0f;
    float v = warp_sum(acc_loss);
    if (lane == 0) red[warp * PH_NPART + 24] = v;
    v = warp_sum(acc_w);
    if (lane == 0) red[warp * PH_NPART + 25] = v;
  }
  __syncthreads();
  if (warp != 0) return;   // the hand-over below is warp 0's business: the other warps leave the SM to the next CTA
  const int tiles = gridDim.x * gridDim.y;
  if (lane < PH_NPART) {
    float s = 0.0f;
#pragma unroll
    for (int wv = 0; wv < PH_NT / 32; wv++) s += red[wv * PH_NPART + lane];
    size_t blk = (size_t)b * tiles + blockIdx.y * gridDim.x + blockIdx.x;
    a.partials[blk * PH_NPART + lane] = s;
  }
}

// Deterministic reduction of the per-CTA partials: grad_P (B,2,12) and the masked mean.  One CTA per sample,
// one warp per value, lanes stride the tiles with independent loads in flight; the cross-sample sum of the two
// loss scalars is done by whichever CTA finishes last (ticket, zeroed by photo_kernel's first CTA), in
// sample order, so the result does not depend on scheduling.  (Doing this inside photo_kernel's last CTAs was
// measured: the per-CTA fence + ticket round trip and the single-CTA tail cost 17 us per pass against 6 us for
// this kernel, which a scheduler can also move off the critical path - profiles/r2_notes.md.)
__global__ void __launch_bounds__(PH_NPART * 32) photo_finalize_kernel(float* __restrict__ partials, int batch,
                                                                      int tiles, float* __restrict__ sums,
                                                                      float* __restrict__ grad_P) {
  const int b = blockIdx.x, v = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tail = partials + (size_t)batch * tiles * PH_NPART;       // [batch][2] per-sample sums, then the tickets
  unsigned* ticket = reinterpret_cast<unsigned*>(tail + (size_t)batch * 2) + batch;
  __shared__ bool last;
  if (grad_P != nullptr || v >= 24) {   // without gradients only the two loss sums are live
    double s = 0.0;
    for (int t = lane; t < tiles; t += 32) s += (double)partials[((size_t)b * tiles + t) * PH_NPART + v];
    s = warp_sum(s);
    if (lane == 0) {
      if (v < 24) grad_P[b * 24 + v] = (float)s;
      else tail[b * 2 + (v - 24)] = (float)s;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == (unsigned)(batch - 1));
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double ls = 0.0, ws = 0.0;
    for (int i = 0; i < batch; i++) {
      ls += (double)reinterpret_cast<volatile float*>(tail)[i * 2];
      ws += (double)reinterpret_cast<volatile float*>(tail)[i * 2 + 1];
    }
    float lf = (float)ls, wf = (float)ws;
    sums[0] = lf;
    sums[1] = wf;
    sums[2] = lf / (wf + 1e-7f);   // loss_utils.py:113
    sums[3] = 0.0f;
  }
}

template <bool WARP, bool GRAD>
static void photo_dispatch(const mal_photo_args& a, const PhotoMaps& maps, int use_tma, int ncand, float min_disp,
                           float range, dim3 grid, size_t smem, cudaStream_t st) {
  const SizeDiv sdiv = size_div(a.height, a.width, a.convention);
  const bool lowres = WARP && a.depth_height > 0;
  const bool syng = WARP && GRAD && a.grad_syn[0] != nullptr;   // PRED mode handles grad_syn in its own branch
  const bool dd = a.zero_img || a.selec_reproj || a.target_out || a.identity_in_pass;   // DynamicDepth's compute_losses mode
#define MAL_PHOTO_GO(CONV_, LOW_, SYNG_, NC_, AVG_) \
  launch(photo_kernel<WARP, GRAD, CONV_, LOW_, SYNG_, NC_, AVG_>, grid, dim3(PH_NT), smem, st, a, maps, ncand, min_disp, range, use_tma, sdiv)
#define MAL_PHOTO_DD(LOW_, NC_) \
  launch(photo_kernel<WARP, GRAD, MAL_CONV_MANYDEPTH, LOW_, false, NC_, false, WARP>, grid, dim3(PH_NT), smem, st, a, maps, ncand, min_disp, range, use_tma, sdiv)
#define MAL_PHOTO_LAUNCH2(CONV_, NC_)                                       \
  do {                                                                      \
    if (lowres && syng) MAL_PHOTO_GO(CONV_, WARP, WARP && GRAD, NC_, false); \
    else if (lowres) MAL_PHOTO_GO(CONV_, WARP, false, NC_, false);           \
    else if (syng) MAL_PHOTO_GO(CONV_, false, WARP && GRAD, NC_, false);     \
    else MAL_PHOTO_GO(CONV_, false, false, NC_, false);                      \
  } while (0)
  // NC is the compiled-in candidate capacity: the 2-candidate passes never touch the temporal-hint
  // staging / rare-path code (instruction-cache pressure); avg_reprojection (2 candidates) is its own instance
#define MAL_PHOTO_LAUNCH(CONV_)                                   \
  do {                                                            \
    if (WARP && CONV_ == MAL_CONV_MANYDEPTH && dd) {              \
      if (lowres && ncand > 2) MAL_PHOTO_DD(WARP, 4);             \
      else if (lowres) MAL_PHOTO_DD(WARP, 2);                     \
      else if (ncand > 2) MAL_PHOTO_DD(false, 4);                 \
      else MAL_PHOTO_DD(false, 2);                                \
    } else if (a.avg_reprojection) {                              \
      if (lowres) MAL_PHOTO_GO(CONV_, WARP, false, 2, true);      \
      else MAL_PHOTO_GO(CONV_, false, false, 2, true);            \
    } else if (ncand > 2) MAL_PHOTO_LAUNCH2(CONV_, 4);            \
    else MAL_PHOTO_LAUNCH2(CONV_, 2);                             \
  } while (0)
  if (a.convention == MAL_CONV_MANYDEPTH) MAL_PHOTO_LAUNCH(MAL_CONV_MANYDEPTH);
  else MAL_PHOTO_LAUNCH(MAL_CONV_DUALREFINE);
#undef MAL_PHOTO_DD
#undef MAL_PHOTO_GO
#undef MAL_PHOTO_LAUNCH2
#undef MAL_PHOTO_LAUNCH
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_photo_partials_floats(int batch, int height, int width) {
  size_t tiles = (size_t)((width + PH_TW - 1) / PH_TW) * ((height + PH_TH - 1) / PH_TH);
  return (size_t)batch * tiles * PH_NPART + (size_t)batch * 2 + (size_t)batch + 1 + 3;   // + per-sample sums + tickets
}

extern "C" int mal_photo_forward(const mal_photo_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_photo_forward: args is NULL");
  const mal_photo_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height >= 3 && a.width >= 3, "mal_photo_forward: bad shape %dx%dx%d", a.batch,
              a.height, a.width);
  MAL_REQUIRE(a.batch <= 65535, "mal_photo_forward: batch %d exceeds gridDim.z", a.batch);
  MAL_REQUIRE(a.mode == MAL_PHOTO_WARP || a.mode == MAL_PHOTO_PRED, "mal_photo_forward: bad mode %d", a.mode);
  MAL_REQUIRE(a.convention == MAL_CONV_MANYDEPTH || a.convention == MAL_CONV_DUALREFINE,
              "mal_photo_forward: bad convention %d", a.convention);
  MAL_REQUIRE(a.target && a.src[0], "mal_photo_forward: target/src[0] pointers are required");
  const bool single = a.src[1] == nullptr;   // compute_reprojection_loss on one prediction
  if (single)
    MAL_REQUIRE(a.mode == MAL_PHOTO_PRED && !a.syn[0], "mal_photo_forward: a single candidate needs PRED mode, no syn");
  MAL_REQUIRE((a.syn[0] == nullptr) == (a.syn[1] == nullptr), "mal_photo_forward: give both syn candidates or none");
  MAL_REQUIRE(a.partials && a.sums, "mal_photo_forward: partials/sums workspaces are required");
  if (a.mode == MAL_PHOTO_WARP) {
    MAL_REQUIRE(a.depth && a.K && a.inv_K && a.T[0] && a.T[1], "mal_photo_forward: WARP mode needs depth,K,inv_K,T");
    if (a.with_grad) MAL_REQUIRE(a.grad_depth && a.grad_P, "mal_photo_forward: WARP+grad needs grad_depth, grad_P");
    if (a.with_grad) MAL_REQUIRE(!a.depth_b, "mal_photo_forward: the averaged (ensemble) disparity carries no gradient");
    if (a.depth_is_disp) MAL_REQUIRE(a.min_depth > 0 && a.max_depth > a.min_depth, "mal_photo_forward: bad depth range");
    MAL_REQUIRE((a.depth_height > 0) == (a.depth_width > 0) && a.depth_height >= 0,
                "mal_photo_forward: depth_height / depth_width must both be set (or both 0)");
    MAL_REQUIRE((a.warped[0] == nullptr) == (a.warped[1] == nullptr), "mal_photo_forward: give both warped sources or none");
    if (a.warped[0])
      MAL_REQUIRE(!a.depth_b && !a.zero_img && !a.selec_reproj && !a.target_out && !a.identity_in_pass,
                  "mal_photo_forward: `warped` goes with a plain WARP pass (no ensemble disparity, no DynamicDepth mode)");
  } else if (a.warped[0] || a.warped[1]) {
    MAL_REQUIRE(false, "mal_photo_forward: `warped` is a WARP-mode input");
  } else if (a.with_grad) {
    MAL_REQUIRE(a.grad_pred[0] && (single || a.grad_pred[1]), "mal_photo_forward: PRED+grad needs grad_pred");
  }
  const int ncand = single ? 1 : (a.syn[0] ? 4 : 2);
  const bool dd = a.zero_img || a.selec_reproj || a.target_out || a.identity_in_pass;
  if (dd) {
    MAL_REQUIRE(a.mode == MAL_PHOTO_WARP && a.convention == MAL_CONV_MANYDEPTH && !a.identity_min && !a.avg_reprojection &&
                    !a.min_reproj_b && !a.grad_syn[0],
                "mal_photo_forward: zero_img / selec_reproj (DynamicDepth) is a WARP-mode pass with the identity "
                "candidates given as `syn`; no identity_min, avg_reprojection, min_reproj_b or grad_syn");
    if (a.syn[0]) MAL_REQUIRE(a.noise || a.ignore_automask, "mal_photo_forward: the in-pass automask needs `noise`");
  } else {
    MAL_REQUIRE((a.identity_min == nullptr) == (a.noise == nullptr),
                "mal_photo_forward: automask needs identity_min and noise together");
  }
  if (a.min_reproj_b)
    MAL_REQUIRE(ncand == 4 && !a.with_grad && !a.identity_min && !a.avg_reprojection,
                "mal_photo_forward: min_reproj_b (a second min over candidates 2, 3) is a forward-only 4-candidate mode");
  if (a.avg_reprojection) {
    MAL_REQUIRE(ncand == 2, "mal_photo_forward: avg_reprojection averages exactly two candidates (no syn, no single)");
    MAL_REQUIRE(!(a.with_grad && a.mode == MAL_PHOTO_PRED),
                "mal_photo_forward: avg_reprojection with gradients is a WARP-mode feature (PRED mode: average the "
                "single-candidate maps of compute_reprojection_loss)");
  }
  const bool grad = a.with_grad != 0;
  const bool warp = a.mode == MAL_PHOTO_WARP;
  // disp_to_depth scalars exactly as python computes them (double), then rounded once to fp32
  const double lo = 1.0 / a.max_depth, hi = 1.0 / a.min_depth;
  const float min_disp = (float)lo, range = (float)(hi - lo);
  dim3 grid((a.width + PH_TW - 1) / PH_TW, (a.height + PH_TH - 1) / PH_TH, a.batch);
  MAL_REQUIRE(grid.y <= 65535, "mal_photo_forward: image too tall");
  size_t smem = photo_smem_bytes(grad, warp, ncand, a.avg_reprojection != 0, dd);
  cudaStream_t st = (cudaStream_t)stream;

  // TMA descriptors for every tensor the CTAs stage whole; any tensor TMA cannot address (rows that are not
  // 16-byte multiples, an unaligned base, an image smaller than the box) sends the call to the plain loader
  const PhotoTile tl = photo_tile(grad);
  PhotoMaps maps;
  memset(&maps, 0, sizeof(maps));
  bool tma = getenv("MAL_PHOTO_NO_TMA") == nullptr;
  const int W = a.width, H = a.height, B = a.batch;
  tma = tma && tile_map_encode(&maps.tgt, a.target, W, H, B * 3, PH_VW, tl.VH, 3);
  const bool pre = warp && a.warped[0] != nullptr;
  if (!warp || pre)
    for (int f = 0; f < (single ? 1 : 2); f++)
      tma = tma && tile_map_encode(&maps.src[f], pre ? a.warped[f] : a.src[f], W, H, B * 3, PH_VW, tl.VH, 3);
  if (ncand > 2)
    for (int k = 0; k < 2; k++) tma = tma && tile_map_encode(&maps.syn[k], a.syn[k], W, H, B * 3, PH_VW, tl.VH, 3);
  if (warp && a.depth_height == 0 && !pre) {
    tma = tma && tile_map_encode(&maps.depth, a.depth, W, H, B, PH_VW, tl.VH, 1);
    if (a.depth_b) tma = tma && tile_map_encode(&maps.depth_b, a.depth_b, W, H, B, PH_VW, tl.VH, 1);
  }
  const int use_tma = tma ? 1 : 0;

  if (warp) {
    if (grad) photo_dispatch<true, true>(a, maps, use_tma, ncand, min_disp, range, grid, smem, st);
    else photo_dispatch<true, false>(a, maps, use_tma, ncand, min_disp, range, grid, smem, st);
  } else {
    if (grad) photo_dispatch<false, true>(a, maps, use_tma, ncand, min_disp, range, grid, smem, st);
    else photo_dispatch<false, false>(a, maps, use_tma, ncand, min_disp, range, grid, smem, st);
  }
  int rc = check_launch("photo_kernel");
  if (rc || a.skip_finalize) return rc;
  return mal_photo_finalize(args, stream);
}

extern "C" int mal_photo_finalize(const mal_photo_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_photo_finalize: args is NULL");
  const mal_photo_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.height >= 3 && a.width >= 3 && a.partials && a.sums, "mal_photo_finalize: bad arguments");
  const bool grad = a.with_grad != 0 && a.mode == MAL_PHOTO_WARP;
  if (grad) MAL_REQUIRE(a.grad_P, "mal_photo_finalize: WARP+grad needs grad_P");
  const int tiles = ((a.width + PH_TW - 1) / PH_TW) * ((a.height + PH_TH - 1) / PH_TH);
  launch(photo_finalize_kernel, dim3(a.batch), dim3(PH_NPART * 32), 0, (cudaStream_t)stream, a.partials, a.batch,
         tiles, a.sums, grad ? a.grad_P : (float*)nullptr);
  return check_launch("photo_finalize_kernel");
}
