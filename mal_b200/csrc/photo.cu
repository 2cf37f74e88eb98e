// photo.cu - fused photometric loss (forward + un-normalised backward) for sm_100a.
//
// One CTA owns a TW x TH tile of one sample.  Phases (separated by __syncthreads):
//   A  stage the target tile (+halo) and every candidate tile in shared memory; in WARP mode
//      the two warped candidates are produced in place:  disp->depth->backproject->project->
//      bilinear gather (border)  (manydepth/layers.py:14-23,163-199, trainer.py:1122-1125)
//   B  per pixel of the loss region: 3x3 SSIM + L1 per candidate (layers.py:243-257,
//      loss_utils.py:46-55), min/argmin over candidates (:103), tie-break noise + automask
//      (:105-109, :27-44), multi-frame mask (:192-194); masked partial sums (:112-113).
//      With gradients, the SSIM derivative coefficients of the selected candidate are left in
//      shared memory for the ring of pixels around the tile.
//   C  (grad) per tile pixel: gather the 3x3 neighbourhood's coefficients (atomics-free SSIM
//      backward), add the L1 term, then chain through the bilinear sampler, the projection and
//      backprojection to d/d depth and d/d(K@T); per-CTA partials go to a workspace.
// A second tiny kernel reduces the per-CTA partials in a fixed order (deterministic).
//
// HBM traffic per pixel (WARP mode, 2+2 candidates, grad): reads 9 (target+2 src... gathers hit
// L2/L1) x 4 B x 3 ch + depth/noise/identity/mask 16 B, writes min_reproj 4 + sel 1 + grad 4 B.
#include "mal_math.cuh"

namespace mal {

constexpr int PH_TW = 32;
constexpr int PH_TH = 16;
constexpr int PH_NT = 256;
constexpr int PH_NPART = 26;  // 2 x 12 dL/dP + sum(w*reproj) + sum(w)

struct PhotoTile {
  int HV, HL;          // value halo, loss halo
  int VW, VH, VN;      // value tile dims
  int LW, LH, LN;      // loss tile dims
};
__host__ __device__ inline PhotoTile photo_tile(bool grad) {
  PhotoTile t;
  t.HV = grad ? 2 : 1;
  t.HL = grad ? 1 : 0;
  t.VW = PH_TW + 2 * t.HV; t.VH = PH_TH + 2 * t.HV; t.VN = t.VW * t.VH;
  t.LW = PH_TW + 2 * t.HL; t.LH = PH_TH + 2 * t.HL; t.LN = t.LW * t.LH;
  return t;
}
inline size_t photo_smem_bytes(bool grad, int ncand) {
  PhotoTile t = photo_tile(grad);
  size_t fl = sizeof(Geom) / 4 + 8 * PH_NPART + 8;
  fl += (size_t)3 * t.VN * (1 + ncand);
  if (grad) fl += (size_t)11 * t.LN;
  return fl * 4 + 16;
}

template <bool WARP, bool GRAD, int CONV, bool LOWRES, bool SYNG, int NC>
__global__ void __launch_bounds__(PH_NT, GRAD ? 3 : 4) photo_kernel(const mal_photo_args a, const int ncand,
                                                     const float min_disp, const float disp_range) {
  const PhotoTile tl = photo_tile(GRAD);
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * PH_TW, y0 = blockIdx.y * PH_TH;
  const int H = a.height, W = a.width;
  const size_t HW = (size_t)H * W;
  // low-resolution disparity, up-sampled on the fly (mal_photo_args.depth_height / depth_width)
  // (a template parameter: the extra addressing costs registers - 97 instead of 79 with gradients -
  // and with them the third resident CTA of the full-resolution case)
  constexpr bool lowres = WARP && LOWRES;
  const size_t dhw = (size_t)a.depth_height * a.depth_width;
  const float up_sy = lowres ? up_scale(a.depth_height, H) : 1.0f, up_sx = lowres ? up_scale(a.depth_width, W) : 1.0f;

  float* smem = reinterpret_cast<float*>(dyn_smem());
  Geom* geom = reinterpret_cast<Geom*>(smem);
  float* red = smem + sizeof(Geom) / 4;                 // [8][PH_NPART]
  float* sy = red + 8 * PH_NPART + 8;                   // [3][VN]
  float* sx = sy + 3 * tl.VN;                           // [ncand][3][VN]
  float* coef = sx + (size_t)ncand * 3 * tl.VN;         // [9][LN]      (GRAD)
  float* lw = coef + 9 * tl.LN;                         // [LN] weights (GRAD)
  int* lsel = reinterpret_cast<int*>(lw + tl.LN);       // [LN] selected warped candidate or -1

  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0) {
    // the finalize kernel's ticket (see photo_finalize_kernel); it launches after this kernel ends
    float* tail = a.partials + (size_t)gridDim.x * gridDim.y * gridDim.z * PH_NPART;
    *reinterpret_cast<unsigned*>(tail + (size_t)gridDim.z * 2) = 0u;
  }
  // ---- phase 0: camera constants --------------------------------------------------------
  if (WARP) {
    if (tid < 24) {
      int f = tid / 12, e = tid % 12;
      geom->P[f][e] = kt_entry(a.K + b * 16, a.T[f] + b * 16, e / 4, e % 4);
    } else if (tid < 33) {
      int e = tid - 24;
      geom->iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
    }
    __syncthreads();
  }

  // ---- phase A: stage value tiles ---------------------------------------------------------
  const float* tgt = a.target + (size_t)b * 3 * HW;
  for (int i = tid; i < tl.VN; i += PH_NT) {
    int ty = i / tl.VW, tx = i - ty * tl.VW;
    int ry = reflect_index(y0 - tl.HV + ty, H), rx = reflect_index(x0 - tl.HV + tx, W);
    ry = min(max(ry, 0), H - 1);
    rx = min(max(rx, 0), W - 1);
    size_t o = (size_t)ry * W + rx;
#pragma unroll
    for (int c = 0; c < 3; c++) sy[c * tl.VN + i] = __ldg(tgt + c * HW + o);
    if (NC > 2 && ncand > 2) {
#pragma unroll
      for (int k = 0; k < 2; k++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          sx[((2 + k) * 3 + c) * tl.VN + i] = __ldg(a.syn[k] + ((size_t)b * 3 + c) * HW + o);
    }
    if (WARP) {
      float dv;
      if (lowres) {   // F.interpolate(disp, [H, W], bilinear) read straight from the low-resolution plane
        const UpAxis ay = up_axis(ry, a.depth_height, up_sy), ax = up_axis(rx, a.depth_width, up_sx);
        dv = upsample_at(a.depth + (size_t)b * dhw, a.depth_width, ay, ax);
        if (a.depth_b) dv = xmul(xadd(dv, upsample_at(a.depth_b + (size_t)b * dhw, a.depth_width, ay, ax)), 0.5f);
      } else {
        dv = __ldg(a.depth + (size_t)b * HW + o);
        if (a.depth_b) dv = xmul(xadd(dv, __ldg(a.depth_b + (size_t)b * HW + o)), 0.5f);   // (a + b) / 2.0
      }
      if (a.depth_is_disp) dv = xdiv(1.0f, xadd(min_disp, xmul(disp_range, dv)));
      Ray ray = pixel_ray(geom->iK, (float)rx, (float)ry);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        Sample s = project_pixel<CONV>(geom->P[f], ray, dv, a.eps, H, W);
        Taps t = make_taps(s.ix, s.iy, H, W);
        const float* src = a.src[f] + (size_t)b * 3 * HW;
#pragma unroll
        for (int c = 0; c < 3; c++) sx[(f * 3 + c) * tl.VN + i] = bilinear(src + c * HW, t);
      }
    } else {
#pragma unroll
      for (int f = 0; f < 2; f++)
        if (f < ncand) {
#pragma unroll
          for (int c = 0; c < 3; c++)
            sx[(f * 3 + c) * tl.VN + i] = __ldg(a.src[f] + ((size_t)b * 3 + c) * HW + o);
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
    const int vc = (ly + tl.HV - tl.HL) * tl.VW + (lx + tl.HV - tl.HL);  // window centre in value tile
    // per-pixel planes are fetched now so their latency hides behind the SSIM arithmetic
    const size_t po = (size_t)b * HW + (size_t)gy * W + gx;
    float p_ident = 0.0f, p_noise = 0.0f, p_mask = 1.0f;
    if (automask) { p_ident = __ldg(a.identity_min + po); p_noise = __ldg(a.noise + po); }
    if (a.pixel_mask) p_mask = __ldg(a.pixel_mask + po);
    float ssum[NC], lsum[NC];   // NC: compiled-in candidate capacity (2 or 4); ncand <= NC are live
    float cf0[9], cf1[9];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float yw[9];
#pragma unroll
      for (int dy = 0; dy < 3; dy++)
#pragma unroll
        for (int dx = 0; dx < 3; dx++) yw[dy * 3 + dx] = sy[c * tl.VN + vc + (dy - 1) * tl.VW + dx - 1];
      float mu_y = 0.f, eyy = 0.f;
      if (!a.no_ssim) {
        mu_y = xdivc<9>(sum9(yw));
        eyy = xdivc<9>(sum9_prod(yw, yw));
      }
#pragma unroll
      for (int k = 0; k < NC; k++) {
        if (k < ncand) {
          const float* X = sx + (k * 3 + c) * tl.VN + vc;
          float xw[9];
#pragma unroll
          for (int dy = 0; dy < 3; dy++)
#pragma unroll
            for (int dx = 0; dx < 3; dx++) xw[dy * 3 + dx] = X[(dy - 1) * tl.VW + dx - 1];
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
    float rmin = 0.f;
    int idx = 0;
#pragma unroll
    for (int k = 0; k < NC; k++) {
      if (k < ncand) {
        float l1m = xdivc<3>(lsum[k]);
        float lk = a.no_ssim ? l1m : xadd(xmul(0.85f, xdivc<3>(ssum[k])), xmul(0.15f, l1m));
        if (k == 0 || lk < rmin) { rmin = lk; idx = k; }
      }
    }
    int mbit = 1;
    if (automask) {
      float ident = xadd(p_ident, xmul(p_noise, 0.00001f));
      mbit = (ident < rmin) ? 0 : 1;  // argmin([reproj, identity]) == 0, first index wins ties
    }
    float w = (float)mbit;
    if (a.pixel_mask) w = xmul(w, p_mask);
    if (a.sample_mask) w = xmul(w, xsub(1.0f, __ldg(a.sample_mask + b)));
    const bool interior = ly >= tl.HL && ly < tl.HL + PH_TH && lx >= tl.HL && lx < tl.HL + PH_TW;
    if (interior) {
      if (a.min_reproj) a.min_reproj[po] = rmin;
      if (a.selection) a.selection[po] = (uint8_t)(idx | (mbit << 7));
      if (a.weight) a.weight[po] = w;
      acc_loss += xmul(rmin, w);
      acc_w += w;
    }
    if (GRAD) {
      // warped candidates always carry a gradient; the temporal-hint candidates when the caller asked
      // for d/d syn (autograd then carries it into the warped images image_synthesis copied from)
      const bool syn_grad = (SYNG || !WARP) && a.grad_syn[0] != nullptr;
      const bool live = (idx < 2 || syn_grad) && w != 0.0f;
      lsel[i] = live ? idx : -1;
      lw[i] = w;
      if (live && !a.no_ssim) {
        const float sc = w * (0.85f / 27.0f);  // weight * 0.85 * (1/3 channels) * (1/9 window)
        if (NC <= 2 || idx < 2) {
#pragma unroll
          for (int j = 0; j < 9; j++) coef[j * tl.LN + i] = (idx == 0 ? cf0[j] : cf1[j]) * sc;
        } else {
          // rare path: re-derive the SSIM terms of the selected temporal-hint candidate
          for (int c = 0; c < 3; c++) {
            float yw[9], xw[9];
            const float* X = sx + (idx * 3 + c) * tl.VN + vc;
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
              for (int dx = 0; dx < 3; dx++) {
                yw[dy * 3 + dx] = sy[c * tl.VN + vc + (dy - 1) * tl.VW + dx - 1];
                xw[dy * 3 + dx] = X[(dy - 1) * tl.VW + dx - 1];
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
      const int vc = (qy + tl.HV) * tl.VW + qx + tl.HV;
      // d S / d (candidate image k) at this pixel, for a candidate that is given as an image: the
      // predictions of PRED mode, and the temporal-hint candidates (k >= 2) of either mode
      auto image_grad = [&](int k, float* out) {
        const size_t pq = (size_t)gy * W + gx;
        float g[3] = {0.f, 0.f, 0.f};
        float xq[3], yq[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { xq[c] = sx[(k * 3 + c) * tl.VN + vc]; yq[c] = sy[c * tl.VN + vc]; }
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
      if (SYNG && a.grad_syn[0] != nullptr)
        for (int k = 2; k < ncand; k++) image_grad(k, a.grad_syn[k - 2]);
      float g0[3] = {0.f, 0.f, 0.f}, g1[3] = {0.f, 0.f, 0.f};
      float xq0[3], xq1[3], yq[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        xq0[c] = sx[(0 * 3 + c) * tl.VN + vc];
        xq1[c] = ncand > 1 ? sx[(1 * 3 + c) * tl.VN + vc] : 0.0f;
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
            if (s < 0 || (SYNG && s > 1)) continue;   // only the warped candidates chain into depth / pose here
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
            float d = yq[c] - (s == 0 ? xq0[c] : xq1[c]);      // target - pred
            float sg = d > 0.f ? -wl : (d < 0.f ? wl : 0.f);   // d|t-p|/dp = -sign(t-p)
            if (s == 0) g0[c] += sg; else g1[c] += sg;
          }
        }
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
          Sample s = project_pixel<CONV>(P, ray, dv, a.eps, H, W);
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
  }

  // ---- per-CTA partials ---------------------------------------------------------------------
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int j = 0; j < 24; j++) {
    float v = (GRAD && WARP) ? warp_sum(gP[j]) : 0.0f;
    if (lane == 0) red[warp * PH_NPART + j] = v;
  }
  {
    float v = warp_sum(acc_loss);
    if (lane == 0) red[warp * PH_NPART + 24] = v;
    v = warp_sum(acc_w);
    if (lane == 0) red[warp * PH_NPART + 25] = v;
  }
  __syncthreads();
  if (tid < PH_NPART) {
    float s = 0.0f;
#pragma unroll
    for (int wv = 0; wv < PH_NT / 32; wv++) s += red[wv * PH_NPART + tid];
    size_t blk = ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    a.partials[blk * PH_NPART + tid] = s;
  }
}

// Deterministic reduction of the per-CTA partials: grad_P (B,2,12) and the masked mean.
// One CTA per sample, one warp per value, every lane keeps PF_TILES loads in flight, so the
// whole reduction is one or two memory round trips spread over B SMs.  The cross-sample sum of the
// two loss scalars is done by whichever CTA finishes last (ticket), in sample order, so the
// result does not depend on scheduling.  The ticket word lives behind the partials and is
// zeroed by photo_kernel's first CTA (stream order makes that visible here).
constexpr int PF_TILES = 8;
__global__ void __launch_bounds__(PH_NPART * 32) photo_finalize_kernel(float* __restrict__ partials, int batch,
                                                                      int tiles, float* __restrict__ sums,
                                                                      float* __restrict__ grad_P) {
  const int b = blockIdx.x, v = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tail = partials + (size_t)batch * tiles * PH_NPART;       // [batch][2] per-sample sums, then the ticket
  unsigned* ticket = reinterpret_cast<unsigned*>(tail + (size_t)batch * 2);
  __shared__ bool last;
  if (grad_P != nullptr || v >= 24) {   // without gradients only the two loss sums are live
    double s = 0.0;
    for (int t0 = 0; t0 < tiles; t0 += 32 * PF_TILES) {
      float x[PF_TILES];
#pragma unroll
      for (int k = 0; k < PF_TILES; k++) {
        const int t = t0 + k * 32 + lane;
        x[k] = t < tiles ? partials[((size_t)b * tiles + t) * PH_NPART + v] : 0.0f;
      }
#pragma unroll
      for (int k = 0; k < PF_TILES; k++) s += (double)x[k];
    }
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
static void photo_dispatch(const mal_photo_args& a, int ncand, float min_disp, float range, dim3 grid,
                           size_t smem, cudaStream_t st) {
  const bool lowres = WARP && a.depth_height > 0;
  const bool syng = WARP && GRAD && a.grad_syn[0] != nullptr;   // PRED mode handles grad_syn in its own branch
#define MAL_PHOTO_LAUNCH2(CONV_, NC_)                                                                        \
  do {                                                                                                      \
    if (lowres && syng) launch(photo_kernel<WARP, GRAD, CONV_, WARP, WARP && GRAD, NC_>, grid, dim3(PH_NT), smem, st, a, ncand, min_disp, range); \
    else if (lowres) launch(photo_kernel<WARP, GRAD, CONV_, WARP, false, NC_>, grid, dim3(PH_NT), smem, st, a, ncand, min_disp, range);          \
    else if (syng) launch(photo_kernel<WARP, GRAD, CONV_, false, WARP && GRAD, NC_>, grid, dim3(PH_NT), smem, st, a, ncand, min_disp, range);    \
    else launch(photo_kernel<WARP, GRAD, CONV_, false, false, NC_>, grid, dim3(PH_NT), smem, st, a, ncand, min_disp, range);                     \
  } while (0)
  // the candidate loops are compiled for 2 or 4 candidates: the 2-candidate passes (identity, ensemble,
  // student) run a kernel half the size of the 4-candidate teacher pass (instruction-cache pressure)
#define MAL_PHOTO_LAUNCH(CONV_)                  \
  do {                                           \
    if (ncand > 2) MAL_PHOTO_LAUNCH2(CONV_, 4);  \
    else MAL_PHOTO_LAUNCH2(CONV_, 2);            \
  } while (0)
  if (a.convention == MAL_CONV_MANYDEPTH) MAL_PHOTO_LAUNCH(MAL_CONV_MANYDEPTH);
  else MAL_PHOTO_LAUNCH(MAL_CONV_DUALREFINE);
#undef MAL_PHOTO_LAUNCH2
#undef MAL_PHOTO_LAUNCH
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_photo_partials_floats(int batch, int height, int width) {
  size_t tiles = (size_t)((width + PH_TW - 1) / PH_TW) * ((height + PH_TH - 1) / PH_TH);
  return (size_t)batch * tiles * PH_NPART + (size_t)batch * 2 + 4;   // + per-sample sums + ticket
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
  MAL_REQUIRE((a.identity_min == nullptr) == (a.noise == nullptr),
              "mal_photo_forward: automask needs identity_min and noise together");
  MAL_REQUIRE(a.partials && a.sums, "mal_photo_forward: partials/sums workspaces are required");
  if (a.mode == MAL_PHOTO_WARP) {
    MAL_REQUIRE(a.depth && a.K && a.inv_K && a.T[0] && a.T[1], "mal_photo_forward: WARP mode needs depth,K,inv_K,T");
    if (a.with_grad) MAL_REQUIRE(a.grad_depth && a.grad_P, "mal_photo_forward: WARP+grad needs grad_depth, grad_P");
    if (a.with_grad) MAL_REQUIRE(!a.depth_b, "mal_photo_forward: the averaged (ensemble) disparity carries no gradient");
    if (a.depth_is_disp) MAL_REQUIRE(a.min_depth > 0 && a.max_depth > a.min_depth, "mal_photo_forward: bad depth range");
    MAL_REQUIRE((a.depth_height > 0) == (a.depth_width > 0) && a.depth_height >= 0,
                "mal_photo_forward: depth_height / depth_width must both be set (or both 0)");
  } else if (a.with_grad) {
    MAL_REQUIRE(a.grad_pred[0] && (single || a.grad_pred[1]), "mal_photo_forward: PRED+grad needs grad_pred");
  }
  const int ncand = single ? 1 : (a.syn[0] ? 4 : 2);
  const bool grad = a.with_grad != 0;
  // disp_to_depth scalars exactly as python computes them (double), then rounded once to fp32
  const double lo = 1.0 / a.max_depth, hi = 1.0 / a.min_depth;
  const float min_disp = (float)lo, range = (float)(hi - lo);
  dim3 grid((a.width + PH_TW - 1) / PH_TW, (a.height + PH_TH - 1) / PH_TH, a.batch);
  MAL_REQUIRE(grid.y <= 65535, "mal_photo_forward: image too tall");
  size_t smem = photo_smem_bytes(grad, ncand);
  cudaStream_t st = (cudaStream_t)stream;
  if (a.mode == MAL_PHOTO_WARP) {
    if (grad) photo_dispatch<true, true>(a, ncand, min_disp, range, grid, smem, st);
    else photo_dispatch<true, false>(a, ncand, min_disp, range, grid, smem, st);
  } else {
    if (grad) photo_dispatch<false, true>(a, ncand, min_disp, range, grid, smem, st);
    else photo_dispatch<false, false>(a, ncand, min_disp, range, grid, smem, st);
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
