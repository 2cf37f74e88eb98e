// cost_volume.cu - plane-sweep matching cost volume (forward only) for sm_100a.
//
// Replaces ResnetEncoderMatching.match_features (manydepth/networks/resnet_encoder.py:151-233,
// dualrefine/networks/resnet_encoder.py:163-245) together with the head that consumes it in
// forward(): compute_confidence_mask (:255-262), the "viz" arg-min -> lowest_cost (:309-313,
// indices_to_disparity :247-253) and cost_volume *= confidence (:317).  The reference builds the
// volume under torch.no_grad(), so there is no backward.
//
// The reference materialises a (bins, C, h, w) warped tensor per sample (189 MB at 96x64x48x160,
// three times over); here nothing larger than the inputs and the (B, bins, h, w) result touches HBM.
//
// Work decomposition
//   kernel 1  cv_pack_kernel: lookup features NCHW -> channel-quad interleaved float4 planes
//             [C/4][h][w] (zero padded to a multiple of 16 channels) so that one 128-bit load fetches 4
//             channels of one bilinear tap.  (The quad sweep reads the current features in place; the
//             general kernel below packs them too.)
//   kernel 2  cv_sweep_kernel: one CTA = 32 consecutive pixels of one sample x all bins.
//             warp  <-> a 16-channel chunk (the cascade-sum granule of the reference, see below)
//             lane  <-> pixel
//             Bins are processed in groups of CV_BG.  Per group:
//               P  all threads: one projection per (pixel, bin) -> {tap offset | masked, the four bilinear
//                  weights} in smem (formed once here, not once per channel chunk: the blend below is
//                  fma-pipe-bound)
//               C  each warp sweeps the group's bins for its chunk.  The 2x2x16 texel block lives in
//                  registers and is only re-fetched when the integer tap origin moves.  Operand
//                  delivery is the real limit of this op: a sweep that loaded its 4 taps + 1 feature
//                  for each of the 566 M (pixel, bin, channel) evaluations would need 88 M L1
//                  wavefronts (~300 us at 128 B/clk/SM) - measured with a lane-per-bin variant that
//                  hit 82% L1 throughput at 449 us, profiles/r1_notes.md - so the taps have to be
//                  reused from registers across consecutive depth planes.
//                  Chunk sums go to smem.
//               F  all threads: combine the chunk sums in the reference's order, mean, edge mask,
//                  accumulate over lookup frames.
//             Epilogue: / (counts + 1e-7), per-pixel max over bins, missing fill, confidence,
//             first-index arg-min, coalesced plane-by-plane stores.
//   kernel 2' cv_sweep_quad_kernel (C <= 64, the default): four lanes per pixel, no block barrier in the sweep;
//             see its own comment further down.  Both sweeps have a DynamicDepth variant (cv_min, set_1 / pool
//             occlusion fill); the pool fill runs its own pre-passes (cv_project / cv_interior / cv_pack_cm /
//             cv_sample / cv_pool), described where they are defined.
//
// Arithmetic contract (bit-exact against torch CPU, pinned by tests/golden/cost_volume.npz):
//   projection / grid_sample arithmetic as in mal_math.cuh;  mean over channels = ATen cascade_sum
//   (SumKernel.cpp multi_row_sum): 16 channels summed sequentially from 0, chunk sums added
//   sequentially, then / C.
#include <cstdlib>

#include "mal_math.cuh"

namespace mal {

constexpr int CV_PX = 32;      // pixels per CTA (one per lane)
constexpr int CV_WARPS = 4;    // channel chunks in flight
constexpr int CV_NT = CV_PX * CV_WARPS;
constexpr int CV_BG = 16;      // bins per group
constexpr int CV_CHUNK = 16;   // channels per chunk (cascade_sum level step)
constexpr int CV_DEFAULT_MINB = 4;

struct CvGeom {
  float P[12];
  float iK[9];
  int live;   // lookup pose .sum() != 0
};

__host__ __device__ inline int cv_padded_channels(int C) { return (C + CV_CHUNK - 1) / CV_CHUNK * CV_CHUNK; }

inline size_t cv_smem_bytes(int nchunks, int num_bins) {
  size_t fl = sizeof(CvGeom) / 4 + 4;
  fl += (size_t)5 * CV_BG * CV_PX;                 // descriptors: off, the four blend weights
  fl += (size_t)nchunks * CV_BG * CV_PX;           // chunk sums
  fl += (size_t)2 * num_bins * CV_PX;              // cost, counts
  fl += (size_t)4 * CV_WARPS * CV_PX;              // epilogue scratch
  return fl * 4 + 16;
}

// NCHW -> [C/4][h][w] float4, zero padded to Cp channels.  One thread per (quad, pixel).
__global__ void __launch_bounds__(256) cv_pack_kernel(const float* __restrict__ src, float4* __restrict__ dst,
                                                      int C, int Cp, int hw, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int p = (int)(i % hw);
  long long r = i / hw;
  int q = (int)(r % (Cp / 4));
  long long img = r / (Cp / 4);
  const float* s = src + (img * C + (long long)q * 4) * hw + p;
  float4 v;
  v.x = (q * 4 + 0 < C) ? __ldg(s) : 0.0f;
  v.y = (q * 4 + 1 < C) ? __ldg(s + hw) : 0.0f;
  v.z = (q * 4 + 2 < C) ? __ldg(s + 2 * (size_t)hw) : 0.0f;
  v.w = (q * 4 + 3 < C) ? __ldg(s + 3 * (size_t)hw) : 0.0f;
  dst[i] = v;
}

// NCHW -> [C/16][h*w][4] float4 (chunk-major: the 16 channels of a texel's chunk are 64 contiguous bytes), zero
// padded.  Read by cv_pool_kernel, whose four quad lanes fetch one texel together.
__global__ void __launch_bounds__(256) cv_pack_cm_kernel(const float* __restrict__ src, float4* __restrict__ dst,
                                                         int C, int Cp, int hw, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 3);
  long long r = i >> 2;
  const int p = (int)(r % hw);
  r /= hw;
  const int chunk = (int)(r % (Cp / CV_CHUNK));
  const long long img = r / (Cp / CV_CHUNK);
  const int c0 = chunk * CV_CHUNK + q * 4;
  const float* sp = src + (img * C + c0) * hw + p;
  float4 v;
  v.x = (c0 + 0 < C) ? __ldg(sp) : 0.0f;
  v.y = (c0 + 1 < C) ? __ldg(sp + hw) : 0.0f;
  v.z = (c0 + 2 < C) ? __ldg(sp + 2 * (size_t)hw) : 0.0f;
  v.w = (c0 + 3 < C) ? __ldg(sp + 3 * (size_t)hw) : 0.0f;
  dst[i] = v;
}

__device__ __forceinline__ float4 ldg4(const float4* p) {
#ifdef MAL_EMU
  return *p;
#else
  return __ldg(p);
#endif
}

// sum_{c in quad} |bilinear(c) - cur(c)|, continuing the sequential chain in `acc`.  The blend and the
// subtraction run two channels per instruction (FFMA2 / FADD2, same IEEE results: mal_common.cuh).
__device__ __forceinline__ float quad_l1(float acc, const float4& a, const float4& b, const float4& c,
                                         const float4& d, const float4& cur, float nw_, float ne_, float sw_,
                                         float se_) {
  const pk2 nw = dup2(nw_), ne = dup2(ne_), sw = dup2(sw_), se = dup2(se_);
  const pk2 lo = x2fma(pack2(d.x, d.y), se, x2fma(pack2(c.x, c.y), sw, x2fma(pack2(b.x, b.y), ne, x2mul(pack2(a.x, a.y), nw))));
  const pk2 hi = x2fma(pack2(d.z, d.w), se, x2fma(pack2(c.z, c.w), sw, x2fma(pack2(b.z, b.w), ne, x2mul(pack2(a.z, a.w), nw))));
  const pk2 dlo = x2sub(lo, pack2(cur.x, cur.y)), dhi = x2sub(hi, pack2(cur.z, cur.w));
  acc = xadd(acc, fabsf(lo2(dlo)));
  acc = xadd(acc, fabsf(hi2(dlo)));
  acc = xadd(acc, fabsf(lo2(dhi)));
  acc = xadd(acc, fabsf(hi2(dhi)));
  return acc;
}


// ---- DynamicDepth extras (dynamicdepth/networks/resnet_encoder.py:191-202) ---------------------
constexpr int CV_OCC_BIT = 1 << 30;   // descriptor flag: the projected occlusion mask exceeds pool_th
constexpr int CV_ZERO_BIT = 1 << 29;  // ... and so does every neighbour of the pool window: the pooled value is 0

// F.grid_sample(occ_mask, pix_locs, zeros, bilinear, align_corners) > pool_th at one location
template <int CONV>
__device__ __forceinline__ bool occluded_at(const float* __restrict__ occ, const GridPoint& gp, int h, int w,
                                            float th) {
  const float ux = unnormalize<CONV>(gp.gx, w), uy = unnormalize<CONV>(gp.gy, h);
  if (!(ux > -2.0f && ux < (float)w + 1.0f && uy > -2.0f && uy < (float)h + 1.0f)) return 0.0f > th;
  Taps t = make_taps(ux, uy, h, w);
  return bilinear(occ, t) > th;
}

// warped value of 4 channels at an arbitrary location, zeros padding
__device__ __forceinline__ float4 bilinear4(const float4* __restrict__ plane, const Taps& t) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = t.v00 ? ldg4(plane + t.o00) : z, b = t.v01 ? ldg4(plane + t.o01) : z;
  const float4 c = t.v10 ? ldg4(plane + t.o10) : z, d = t.v11 ? ldg4(plane + t.o11) : z;
  float4 r;
  r.x = xfma(d.x, t.se, xfma(c.x, t.sw, xfma(b.x, t.ne, xmul(a.x, t.nw))));
  r.y = xfma(d.y, t.se, xfma(c.y, t.sw, xfma(b.y, t.ne, xmul(a.y, t.nw))));
  r.z = xfma(d.z, t.se, xfma(c.z, t.sw, xfma(b.z, t.ne, xmul(a.z, t.nw))));
  r.w = xfma(d.w, t.se, xfma(c.w, t.sw, xfma(b.w, t.ne, xmul(a.w, t.nw))));
  return r;
}

// the same from the chunk-major copy [chunk][h*w][4 quads]: `cell` points at quad q of texel 0
__device__ __forceinline__ float4 bilinear4_cm(const float4* __restrict__ cell, const Taps& t) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 a = t.v00 ? ldg4(cell + 4 * (size_t)t.o00) : z, b = t.v01 ? ldg4(cell + 4 * (size_t)t.o01) : z;
  const float4 c = t.v10 ? ldg4(cell + 4 * (size_t)t.o10) : z, d = t.v11 ? ldg4(cell + 4 * (size_t)t.o11) : z;
  float4 r;
  r.x = xfma(d.x, t.se, xfma(c.x, t.sw, xfma(b.x, t.ne, xmul(a.x, t.nw))));
  r.y = xfma(d.y, t.se, xfma(c.y, t.sw, xfma(b.y, t.ne, xmul(a.y, t.nw))));
  r.z = xfma(d.z, t.se, xfma(c.z, t.sw, xfma(b.z, t.ne, xmul(a.z, t.nw))));
  r.w = xfma(d.w, t.se, xfma(c.w, t.sw, xfma(b.w, t.ne, xmul(a.w, t.nw))));
  return r;
}

__device__ __forceinline__ float quad_l1_vals(float acc, const float4& v, const float4& cur) {
  acc = xadd(acc, fabsf(xsub(v.x, cur.x)));
  acc = xadd(acc, fabsf(xsub(v.y, cur.y)));
  acc = xadd(acc, fabsf(xsub(v.z, cur.z)));
  acc = xadd(acc, fabsf(xsub(v.w, cur.w)));
  return acc;
}

// "pool": an occluded sample takes, per channel, the max over its (2r+1)^3 neighbourhood in
// (bin, y, x) of the warped features with occluded entries zeroed (F.max_pool3d, implicit -inf
// padding).
// ---- "pool" with a descriptor volume -----------------------------------------------------------
// The pool window of an occluded sample reads its 26 neighbours in (bin, y, x).  Doing that inside the sweep (one
// lane walking its window while the other 31 wait, every neighbour re-projected, once per chunk) cost 4.0 of the
// 4.6 ms of the call at the Cityscapes bench shape.  Instead:
//   cv_project_kernel   projects every (lookup frame, bin, pixel) ONCE into a descriptor volume
//                       [B*F][bins][h*w] x {ux, uy, flags} (HBM is idle in this op);
//   cv_interior_kernel  marks the occluded samples whose whole window is occluded (the inside of a blob: the pooled
//                       value is 0) and appends the others - the rim, ~2 % of all samples - to a list;
//   cv_pool_kernel      one warp per listed sample: lanes = 8 neighbour slots x 4 channel quads, max over the slots
//                       by shuffles, then the |pooled - current| chunk sums in the reference's order -> `parts`;
//   the sweep           reads descriptors instead of projecting, and one float per (rim sample, chunk).
//   (the rim windows overlap: every un-occluded sample next to a blob is warped ONCE into a cache - cv_interior marks
//   them and hands out the cache slots, cv_sample fills them - and cv_pool takes its maxima over cached vectors)
// Workspace: CvDescLayout below.
constexpr int CV_DESC_EDGE = 1;   // pixel and sampling location pass the border masks (:203-212)
constexpr int CV_DESC_OCC = 2;    // projected occlusion mask > pool_th (:194-195)
constexpr int CV_DESC_ZERO = 4;   // ... and every sample of its pool window too: the pooled value is 0
constexpr int CV_DESC_NEED = 8;   // un-occluded, inside the pool window of a rim sample: its warped vector is cached
constexpr int CV_DESC_SLOT = 4;   // bits 4..31: cache slot + 1 (0: not cached)

__host__ __device__ inline size_t cv_desc_plane(int batch, int num_lookup, int num_bins, int hw) {
  return (size_t)batch * num_lookup * num_bins * hw;
}
// Workspace layout of the pool fill, in floats / ints (plane = B*F*bins*h*w samples).
struct CvDescLayout {
  size_t plane, ux, uy, flags, list, parts, counters, cm, list2, cache, total;
  int cap;   // cache capacity in samples
};
__host__ __device__ inline CvDescLayout cv_desc_layout(int batch, int num_lookup, int num_bins, int hw, int Cp) {
  CvDescLayout L;
  const int nchunks = Cp / CV_CHUNK;
  L.plane = cv_desc_plane(batch, num_lookup, num_bins, hw);
  L.ux = 0; L.uy = L.plane; L.flags = 2 * L.plane; L.list = 3 * L.plane; L.parts = 4 * L.plane;
  L.counters = (4 + (size_t)nchunks) * L.plane;                  // [0] rim samples, [1] cached samples
  L.cm = (L.counters + 4 + 3) / 4 * 4;                           // chunk-major lookup copy, 16-byte aligned
  L.list2 = L.cm + (size_t)batch * num_lookup * Cp * hw;
  const size_t cap = L.plane / 8 > 1024 ? L.plane / 8 : 1024;    // samples next to an occluded blob: ~2 % in practice
  L.cap = (int)(cap < 0x07ffffff ? cap : 0x07ffffff);            // (slot + 1) lives in bits 4..31 of the flag word
  L.cache = (L.list2 + L.cap + 3) / 4 * 4;
  L.total = L.cache + (size_t)L.cap * Cp;
  return L;
}

// Warp-aggregated list append: one atomicAdd per warp instead of one per element (a quarter of a million appends to
// one counter serialise otherwise).  Every lane of the warp must call it; returns the element's index or -1.
__device__ __forceinline__ int warp_append(int* counter, bool mine) {
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  if (m == 0u) return -1;
  const int lane = threadIdx.x & 31, leader = __ffs((int)m) - 1;
  int first = 0;
  if (lane == leader) first = atomicAdd(counter, __popc(m));
  first = __shfl_sync(0xffffffffu, first, leader);
  return mine ? first + __popc(m & ((1u << lane) - 1u)) : -1;
}

constexpr int CV_PJ = 8;   // planes per projection thread (the pixel's ray and occlusion row stay in registers)
template <int CONV>
__global__ void __launch_bounds__(256) cv_project_kernel(const mal_cost_volume_args a, const SizeDiv sdiv) {
  __shared__ CvGeom geom;
  const int h = a.height, w = a.width, hw = h * w, nb = a.num_bins;
  const int ngroups = (nb + CV_PJ - 1) / CV_PJ;
  const int grp = blockIdx.x % ngroups, bf = blockIdx.x / ngroups, b = bf / a.num_lookup;
  const int tid = threadIdx.x;
  if (tid < 12) {
    geom.P[tid] = kt_entry(a.K + b * 16, a.poses + (size_t)bf * 16, tid / 4, tid % 4);
  } else if (tid < 21) {
    const int e = tid - 12;
    geom.iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
  } else if (tid == 21) {
    const float* T = a.poses + (size_t)bf * 16;
    float s = 0.0f;
    for (int e = 0; e < 16; e++) s += T[e];
    geom.live = (s != 0.0f) ? 1 : 0;
  }
  __syncthreads();
  const int p = blockIdx.y * 256 + tid;
  if (p >= hw) return;
  const size_t plane = cv_desc_plane(a.batch, a.num_lookup, nb, hw);
  const int k0 = grp * CV_PJ, k1 = min(nb, k0 + CV_PJ);
  float* d_ux = a.desc + (size_t)bf * nb * hw + p;
  int* d_fl = reinterpret_cast<int*>(a.desc) + 2 * plane + (size_t)bf * nb * hw + p;
  if (!geom.live) {   // the sweep skips the frame; the passes over the flag plane must see "nothing here"
    for (int k = k0; k < k1; k++) d_fl[(size_t)k * hw] = 0;
    return;
  }
  const int py = p / w, px = p - py * w;
  const float* occ = nullptr;
  if (!(a.aug_mask && __ldg(a.aug_mask + b) != 0.0f)) occ = a.occ + (size_t)b * hw;
  const Ray ray = pixel_ray(geom.iK, (float)px, (float)py);
  const bool inner = py >= 2 && py < h - 2 && px >= 2 && px < w - 2;
#pragma unroll 2
  for (int k = k0; k < k1; k++) {
    const GridPoint gp = project_grid<CONV>(geom.P, ray, __ldg(a.bins + k), a.eps, h, w, &sdiv);
    const float xv = xmul(xadd(xmul(gp.gx, 0.5f), 0.5f), (float)(w - 1));
    const float yv = xmul(xadd(xmul(gp.gy, 0.5f), 0.5f), (float)(h - 1));
    int flags = 0;
    if (inner && xv >= 2.0f && xv <= (float)(w - 2) && yv >= 2.0f && yv <= (float)(h - 2)) flags |= CV_DESC_EDGE;
    if (occ && occluded_at<CONV>(occ, gp, h, w, a.pool_th)) flags |= CV_DESC_OCC;
    d_ux[(size_t)k * hw] = unnormalize<CONV>(gp.gx, w);
    d_ux[plane + (size_t)k * hw] = unnormalize<CONV>(gp.gy, h);
    d_fl[(size_t)k * hw] = flags;
  }
}

// Inside an occluded blob every sample of the pool window is occluded too: the pooled value is 0 and the sweep need
// not look.  One pass over the flag plane marks those samples (27 independent loads; out-of-range neighbours are
// clamped onto in-window ones, which leaves the AND unchanged), lists the others (the rim) and marks the
// un-occluded samples their windows reach.
__global__ void __launch_bounds__(256) cv_interior_kernel(const mal_cost_volume_args a, const int Cp) {
  const int h = a.height, w = a.width, hw = h * w, nb = a.num_bins, r = a.pool_radius;
  const int k = blockIdx.x % nb, bf = blockIdx.x / nb;
  const int p = blockIdx.y * 256 + threadIdx.x;
  const CvDescLayout L = cv_desc_layout(a.batch, a.num_lookup, nb, hw, Cp);
  int* base = reinterpret_cast<int*>(a.desc);
  int* fl = base + L.flags + (size_t)bf * nb * hw;
  const int mine = p < hw ? fl[(size_t)k * hw + p] : 0;
  const int py = p / w, px = p - py * w;
  bool rim = false;
  if (mine & CV_DESC_OCC) {
    int all = CV_DESC_OCC;
    for (int dk = -r; dk <= r; dk++) {
      const int kk = min(max(k + dk, 0), nb - 1);
      for (int dy = -r; dy <= r; dy++) {
        const int yy = min(max(py + dy, 0), h - 1);
#pragma unroll
        for (int dx = -3; dx <= 3; dx++) {
          if (dx < -r || dx > r) continue;
          const int xx = min(max(px + dx, 0), w - 1);
          all &= fl[(size_t)kk * hw + (size_t)yy * w + xx];   // (plain loads: the words change under this kernel)
        }
      }
    }
    // (other threads read and atomicOr these words meanwhile: ZERO goes onto occluded samples only, NEED onto
    // un-occluded ones only, and readers look at bit CV_DESC_OCC, which never changes)
    if (all & CV_DESC_OCC) fl[(size_t)k * hw + p] = mine | CV_DESC_ZERO;
    else rim = (mine & CV_DESC_EDGE) != 0;   // a rim sample the sweep will use
  }
  // cv_pool_kernel's work list (order is irrelevant: every entry owns its output slots)
  const int at = warp_append(base + L.counters, rim);
  if (!rim) return;
  base[L.list + at] = (int)(((size_t)bf * nb + k) * hw + p);
  for (int dk = -r; dk <= r; dk++) {
    const int kk = k + dk;
    if (kk < 0 || kk >= nb) continue;
    for (int dy = -r; dy <= r; dy++) {
      const int yy = py + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = -r; dx <= r; dx++) {
        const int xx = px + dx;
        if (xx < 0 || xx >= w) continue;
        int* nf = fl + (size_t)kk * hw + (size_t)yy * w + xx;
        if (*reinterpret_cast<volatile int*>(nf) & (CV_DESC_OCC | CV_DESC_NEED)) continue;
        // whoever sets NEED first hands the sample its cache slot (an overflowing one keeps slot 0: cv_pool_kernel
        // then warps it itself)
        if (atomicOr(nf, CV_DESC_NEED) & CV_DESC_NEED) continue;
        const int idx = atomicAdd(base + L.counters + 1, 1);
        if (idx >= L.cap) continue;
        atomicOr(nf, (idx + 1) << CV_DESC_SLOT);
        base[L.list2 + idx] = (int)((size_t)bf * nb * hw + (size_t)kk * hw + (size_t)yy * w + xx);
      }
    }
  }
}

// The warped feature vector (all channels) of every cached sample: half a warp per sample, lane = channel quad.
__global__ void __launch_bounds__(256) cv_sample_kernel(const mal_cost_volume_args a, const int Cp) {
  const int h = a.height, w = a.width, hw = h * w, nb = a.num_bins, nquads = Cp / 4;
  const CvDescLayout L = cv_desc_layout(a.batch, a.num_lookup, nb, hw, Cp);
  const int* base = reinterpret_cast<const int*>(a.desc);
  const int count = min(base[L.counters + 1], L.cap);
  const float4* lookcm = reinterpret_cast<const float4*>(a.desc + L.cm);
  float4* cache = reinterpret_cast<float4*>(a.desc + L.cache);
  const int half = threadIdx.x >> 4, qi0 = threadIdx.x & 15;
  const int nhalves = gridDim.x * (blockDim.x >> 4);
  for (int i = blockIdx.x * (blockDim.x >> 4) + half; i < count; i += nhalves) {
    const int o = __ldg(base + L.list2 + i);
    const int bf = o / (nb * hw);
    const float ux = __ldg(a.desc + L.ux + o), uy = __ldg(a.desc + L.uy + o);
    const bool any = ux > -2.0f && ux < (float)w + 1.0f && uy > -2.0f && uy < (float)h + 1.0f;
    // keep float -> int defined for wild coordinates: every tap is out of range anyway
    const Taps t = make_taps(any ? ux : -4.0f, any ? uy : -4.0f, h, w);
    for (int qi = qi0; qi < nquads; qi += 16) {
      const float4* cell = lookcm + (size_t)bf * nquads * hw + (size_t)(qi >> 2) * hw * 4 + (qi & 3);
      cache[(size_t)i * nquads + qi] = any ? bilinear4_cm(cell, t) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// One warp per rim sample: lane = neighbour slot (2) x channel quad (16), sixteen quads (four chunks) per pass.  The
// window's maxima come from the cached vectors; a sample without a slot (cache overflow) is warped here.
template <int R>   // pool radius as a constant (1 is the reference's default); 0: read it from the arguments
__global__ void __launch_bounds__(256, 4) cv_pool_kernel(const mal_cost_volume_args a, const int Cp) {
  const int h = a.height, w = a.width, hw = h * w, nb = a.num_bins, r = R ? R : a.pool_radius, side = 2 * r + 1;
  const int n = side * side * side, nquads = Cp / 4;
  const CvDescLayout L = cv_desc_layout(a.batch, a.num_lookup, nb, hw, Cp);
  const int* base = reinterpret_cast<const int*>(a.desc);
  const int count = base[L.counters];
  const float4* curq = reinterpret_cast<const float4*>(a.packed);
  const float4* lookcm = reinterpret_cast<const float4*>(a.desc + L.cm);
  const float4* cache = reinterpret_cast<const float4*>(a.desc + L.cache);
  float* parts = a.desc + L.parts;
  const int lane = threadIdx.x & 31, slot = lane >> 4, ql = lane & 15;
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += nwarps) {
    const int o = __ldg(base + L.list + i);
    const int bf = o / (nb * hw), rem = o - bf * nb * hw, k = rem / hw, p = rem - k * hw;
    const int py = p / w, px = p - py * w, b = bf / a.num_lookup;
    const size_t fbase = (size_t)bf * nb * hw;
    for (int q0 = 0; q0 < nquads; q0 += 16) {
      const int qi = q0 + ql;
      const bool qlive = qi < nquads;
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);   // the centre is occluded: contributes 0
      for (int j0 = 0; j0 < n; j0 += 32) {
        // the window's flag words, one per lane, in one round trip; -1: outside the volume
        int myword = -1;
        size_t myon = 0;
        {
          const int j = j0 + lane;
          const int kk = k + j / (side * side) - r, yy = py + (j / side) % side - r, xx = px + j % side - r;
          if (j < n && kk >= 0 && kk < nb && yy >= 0 && yy < h && xx >= 0 && xx < w) {
            myon = fbase + (size_t)kk * hw + (size_t)yy * w + xx;
            myword = __ldg(base + L.flags + myon);
          }
        }
#pragma unroll 4
        for (int jj = 0; jj < 32; jj += 2) {
          if (j0 + jj >= n) break;   // (warp-uniform)
          const int word = __shfl_sync(0xffffffffu, myword, jj + slot);
          const unsigned long long on = __shfl_sync(0xffffffffu, (unsigned long long)myon, jj + slot);
          if (word == -1 || (word & CV_DESC_OCC) || !qlive || j0 + jj + slot >= n) continue;   // x[mask] = 0
          float4 v;
          const int cs = (int)((unsigned)word >> CV_DESC_SLOT);
          if (cs) {
            v = ldg4(cache + (size_t)(cs - 1) * nquads + qi);
          } else {
            const float ux = __ldg(a.desc + L.ux + on), uy = __ldg(a.desc + L.uy + on);
            if (!(ux > -2.0f && ux < (float)w + 1.0f && uy > -2.0f && uy < (float)h + 1.0f)) continue;   // all taps are zero
            const Taps t = make_taps(ux, uy, h, w);
            v = bilinear4_cm(lookcm + (size_t)bf * nquads * hw + (size_t)(qi >> 2) * hw * 4 + (qi & 3), t);
          }
          m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
      }
      __syncwarp();
      m.x = fmaxf(m.x, __shfl_xor_sync(0xffffffffu, m.x, 16));
      m.y = fmaxf(m.y, __shfl_xor_sync(0xffffffffu, m.y, 16));
      m.z = fmaxf(m.z, __shfl_xor_sync(0xffffffffu, m.z, 16));
      m.w = fmaxf(m.w, __shfl_xor_sync(0xffffffffu, m.w, 16));
      // |pooled - current| summed over each chunk's 16 channels in order: the chain runs through the chunk's 4 lanes
      const float4 cur = (lane < 16 && qlive) ? ldg4(curq + ((size_t)b * nquads + qi) * hw + p) : make_float4(0.f, 0.f, 0.f, 0.f);
      float acc = 0.0f;
#pragma unroll
      for (int qq = 0; qq < 4; qq++) {
        if ((lane & 3) == qq) acc = quad_l1_vals(acc, m, cur);
        acc = __shfl_sync(0xffffffffu, acc, (lane & ~3) + qq);
      }
      if (lane < 16 && (lane & 3) == 0 && qlive) parts[(size_t)(qi >> 2) * L.plane + o] = acc;
    }
  }
}

template <int CONV, int MINB, bool DYN>
__global__ void __launch_bounds__(CV_NT, MINB) cv_sweep_kernel(const mal_cost_volume_args a, const int Cp, const SizeDiv sdiv) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = a.height, w = a.width, hw = h * w;
  const int nb = a.num_bins, nchunks = Cp / CV_CHUNK, nquads = Cp / 4;
  const int tiles = (hw + CV_PX - 1) / CV_PX;
  const int b = blockIdx.x / tiles;
  const int p = (blockIdx.x - b * tiles) * CV_PX + lane;   // this lane's pixel (flat index)
  const bool pix_ok = p < hw;
  const int py = pix_ok ? p / w : 0, px = pix_ok ? p - py * w : 0;

  float* smem = reinterpret_cast<float*>(dyn_smem());
  CvGeom* geom = reinterpret_cast<CvGeom*>(smem);
  int* d_off = reinterpret_cast<int*>(smem + sizeof(CvGeom) / 4 + 4);   // [BG][PX]
  float* d_w = reinterpret_cast<float*>(d_off + CV_BG * CV_PX);   // [4][BG][PX] blend weights nw, ne, sw, se: formed
                                                                  // once per (pixel, plane) here, not once per chunk
  float* part = d_w + 4 * CV_BG * CV_PX;                 // [nchunks][BG][PX]
  float* cost = part + (size_t)nchunks * CV_BG * CV_PX;  // [nb][PX]
  float* cnt = cost + (size_t)nb * CV_PX;                // [nb][PX]
  float* scr = cnt + (size_t)nb * CV_PX;                 // [4][WARPS][PX]

  const bool cv_min = DYN && a.cv_min;
  for (int i = tid; i < nb * CV_PX; i += CV_NT) { cost[i] = cv_min ? 1.0f : 0.0f; cnt[i] = 0.0f; }
  // a NaN / Inf among a pixel's current features also poisons its masked planes (NaN * 0 in the reference,
  // resnet_encoder.py:211-212); flagged per pixel by whichever warp holds the offending chunk
  __shared__ int s_poison[CV_PX];
  if (tid < CV_PX) s_poison[tid] = 0;
  // occlusion handling applies to samples whose matching augmentation is off (aug_mask == 0)
  const float* occ = nullptr;
  if (DYN && a.occ && a.occ_mode != 0 && !(a.aug_mask && __ldg(a.aug_mask + b) != 0.0f)) occ = a.occ + (size_t)b * hw;

  const float4* curq = reinterpret_cast<const float4*>(a.packed) + (size_t)b * nquads * hw;
  const float4* lookq = reinterpret_cast<const float4*>(a.packed) + (size_t)a.batch * nquads * hw;
  // the pixel's own border mask (current_mask[:, 2:-2, 2:-2], resnet_encoder.py:203-205)
  const bool inner = pix_ok && py >= 2 && py < h - 2 && px >= 2 && px < w - 2;

  for (int f = 0; f < a.num_lookup; f++) {
    __syncthreads();   // previous frame's geometry and descriptors are no longer read
    if (tid < 12) {
      geom->P[tid] = kt_entry(a.K + b * 16, a.poses + ((size_t)b * a.num_lookup + f) * 16, tid / 4, tid % 4);
    } else if (tid < 21) {
      int e = tid - 12;
      geom->iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
    } else if (tid == 21) {
      const float* T = a.poses + ((size_t)b * a.num_lookup + f) * 16;
      float s = 0.0f;
      for (int e = 0; e < 16; e++) s += T[e];
      geom->live = (s != 0.0f) ? 1 : 0;   // "ignore missing images", resnet_encoder.py:183-185
    }
    __syncthreads();
    if (!geom->live) continue;   // uniform across the CTA

    const Ray ray = pixel_ray(geom->iK, (float)px, (float)py);
    const float4* lq = lookq + ((size_t)b * a.num_lookup + f) * nquads * hw;
    // descriptor volume of this (sample, lookup frame): only with the pool fill and only where it applies
    const size_t dplane = DYN ? cv_desc_plane(a.batch, a.num_lookup, nb, hw) : 0;
    const float* desc = (DYN && a.desc && occ && a.occ_mode == MAL_CV_OCC_POOL)
                            ? a.desc + ((size_t)b * a.num_lookup + f) * nb * hw : nullptr;

    for (int g0 = 0; g0 < nb; g0 += CV_BG) {
      const int gn = min(CV_BG, nb - g0);
      // ---- P: projection descriptors ---------------------------------------------------------
      if (DYN && desc) {   // projected once by cv_project_kernel: all of the warp's loads go out together
        static_assert(CV_BG == 4 * CV_WARPS, "four planes of a group per warp");
        int fls[4];
        float uxs[4], uys[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = warp + CV_WARPS * i;
          const bool live = pix_ok && k < gn;
          const size_t o = live ? (size_t)(g0 + k) * hw + p : 0;
          fls[i] = live ? __ldg(reinterpret_cast<const int*>(desc) + 2 * dplane + o) : 0;
          uxs[i] = __ldg(desc + o);
          uys[i] = __ldg(desc + dplane + o);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = warp + CV_WARPS * i;
          if (k >= gn) continue;
          int off = -1;
          float tx = 0.0f, ty = 0.0f;
          if (fls[i] & CV_DESC_EDGE) {
            const float x0 = floorf(uxs[i]), y0 = floorf(uys[i]);
            off = min(max((int)y0, 0), h - 2) * w + min(max((int)x0, 0), w - 2);
            tx = xsub(uxs[i], x0);
            ty = xsub(uys[i], y0);
            if (fls[i] & CV_DESC_OCC) off |= CV_OCC_BIT | ((fls[i] & CV_DESC_ZERO) ? CV_ZERO_BIT : 0);
          }
          d_off[k * CV_PX + lane] = off;
          {
            const float e = xsub(1.0f, tx), sfr = xsub(1.0f, ty);
            d_w[(0 * CV_BG + k) * CV_PX + lane] = xmul(sfr, e);
            d_w[(1 * CV_BG + k) * CV_PX + lane] = xmul(sfr, tx);
            d_w[(2 * CV_BG + k) * CV_PX + lane] = xmul(ty, e);
            d_w[(3 * CV_BG + k) * CV_PX + lane] = xmul(ty, tx);
          }
        }
      } else
      for (int k = warp; k < gn; k += CV_WARPS) {
        int off = -1;
        float tx = 0.0f, ty = 0.0f;
        if (inner) {
          float depth = __ldg(a.bins + g0 + k);
          GridPoint gp = project_grid<CONV>(geom->P, ray, depth, a.eps, h, w, &sdiv);
          // edge mask on the sampling location (resnet_encoder.py:196-201)
          float xv = xmul(xadd(xmul(gp.gx, 0.5f), 0.5f), (float)(w - 1));
          float yv = xmul(xadd(xmul(gp.gy, 0.5f), 0.5f), (float)(h - 1));
          bool edge = xv >= 2.0f && xv <= (float)(w - 2) && yv >= 2.0f && yv <= (float)(h - 2);
          if (edge) {
            float ux = unnormalize<CONV>(gp.gx, w), uy = unnormalize<CONV>(gp.gy, h);
            float x0 = floorf(ux), y0 = floorf(uy);
            int xi = (int)x0, yi = (int)y0;
            // with the edge mask on, all four taps are inside the image for both conventions
            // (zeros padding never triggers); the clamp only guards pathological inputs
            xi = min(max(xi, 0), w - 2);
            yi = min(max(yi, 0), h - 2);
            off = yi * w + xi;
            tx = xsub(ux, x0);
            ty = xsub(uy, y0);
            if (DYN && occ && occluded_at<CONV>(occ, gp, h, w, a.pool_th)) off |= CV_OCC_BIT;
          }
        }
        d_off[k * CV_PX + lane] = off;
        {
          const float e = xsub(1.0f, tx), sfr = xsub(1.0f, ty);
          d_w[(0 * CV_BG + k) * CV_PX + lane] = xmul(sfr, e);
          d_w[(1 * CV_BG + k) * CV_PX + lane] = xmul(sfr, tx);
          d_w[(2 * CV_BG + k) * CV_PX + lane] = xmul(ty, e);
          d_w[(3 * CV_BG + k) * CV_PX + lane] = xmul(ty, tx);
        }
      }
      __syncthreads();

      // ---- C: channel sweep, one 16-channel chunk per warp pass ------------------------------
      for (int ch = warp; ch < nchunks; ch += CV_WARPS) {
        const float4* lqc = lq + (size_t)ch * 4 * hw;
        float4 cq[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
          cq[j] = pix_ok ? ldg4(curq + (size_t)(ch * 4 + j) * hw + p) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (!DYN) {
          bool bad = false;
#pragma unroll
          for (int j = 0; j < 4; j++)
            bad |= !(fabsf(cq[j].x) < INFINITY) || !(fabsf(cq[j].y) < INFINITY) || !(fabsf(cq[j].z) < INFINITY) ||
                   !(fabsf(cq[j].w) < INFINITY);
          if (bad) s_poison[lane] = 1;
        }
        float4 t00[4], t01[4], t10[4], t11[4];
        int coff = -1;
        const int* po = d_off + lane;
        const float* pw = d_w + lane;
        float* pp = part + (size_t)ch * CV_BG * CV_PX + lane;
        for (int k = 0; k < gn; k++) {
          int off = po[k * CV_PX];
          float acc = 0.0f;
          if (DYN && off >= 0 && (off & CV_OCC_BIT)) {
            if (a.occ_mode == MAL_CV_OCC_SET_1) {            // warped[mask] = 1.0
              const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
              for (int j = 0; j < 4; j++) acc = quad_l1_vals(acc, one, cq[j]);
            } else if (off & CV_ZERO_BIT) {                   // warped[mask] = max_pool3d(x)[mask], 0 inside a blob
              const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int j = 0; j < 4; j++) acc = quad_l1_vals(acc, zero, cq[j]);
            } else {                                          // ... computed by cv_pool_kernel on the rim
              acc = __ldg(a.desc + (4 + (size_t)ch) * dplane + ((size_t)b * a.num_lookup + f) * nb * hw +
                          (size_t)(g0 + k) * hw + p);
            }
            off = -1;
          }
          if (off >= 0) {
            if (off != coff) {
              coff = off;
              const float4* base = lqc + off;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                t00[j] = ldg4(base); t01[j] = ldg4(base + 1);
                const float4* row1 = base + w;
                t10[j] = ldg4(row1); t11[j] = ldg4(row1 + 1);
                base += hw;
              }
            }
            const float nw = pw[(0 * CV_BG + k) * CV_PX], ne = pw[(1 * CV_BG + k) * CV_PX];
            const float sw = pw[(2 * CV_BG + k) * CV_PX], se = pw[(3 * CV_BG + k) * CV_PX];
#pragma unroll
            for (int j = 0; j < 4; j++) acc = quad_l1(acc, t00[j], t01[j], t10[j], t11[j], cq[j], nw, ne, sw, se);
          }
          pp[k * CV_PX] = acc;
        }
      }
      __syncthreads();

      // ---- F: combine chunks, mean, accumulate over lookup frames ----------------------------
      for (int k = warp; k < gn; k += CV_WARPS) {
        if (d_off[k * CV_PX + lane] >= 0) {
          float s = part[(size_t)k * CV_PX + lane];                       // 0 + c0
          for (int ch = 1; ch < nchunks; ch++) s = xadd(s, part[((size_t)ch * CV_BG + k) * CV_PX + lane]);
          float diff = xdiv(s, (float)a.channels);                         // .mean(1), edge mask == 1
          int o = (g0 + k) * CV_PX + lane;
          if (cv_min) {   // diffs[diffs == 0] = 1.0; cost_volume = minimum(diffs, cost_volume)
            if (diff == 0.0f) diff = 1.0f;
            cost[o] = fminf(diff, cost[o]);
          } else {
            cost[o] = xadd(cost[o], diff);
            if (diff > 0.0f) cnt[o] = xadd(cnt[o], 1.0f);
          }
        } else if (!DYN && pix_ok && s_poison[lane]) {
          cost[(g0 + k) * CV_PX + lane] = NAN;   // masked plane of a poisoned pixel: NaN * 0
        }
      }
      // the next group's P phase rewrites the descriptors only after the barrier below
      __syncthreads();
    }
  }
  __syncthreads();

  // ---- epilogue -------------------------------------------------------------------------------
  // warp q owns a contiguous range of bins for every pixel of the tile
  const int per = (nb + CV_WARPS - 1) / CV_WARPS;
  const int k0 = min(nb, warp * per), k1 = min(nb, k0 + per);
  float vmax = -INFINITY;
  for (int k = k0; k < k1; k++) {
    int o = k * CV_PX + lane;
    float v;
    if (cv_min) v = (cost[o] == 1.0f) ? 0.0f : cost[o];   // cost_volume[cost_volume == 1] = 0
    else v = xdiv(cost[o], xadd(cnt[o], 1e-7f));         // cost_volume / (counts + 1e-7)
    cost[o] = v;
    vmax = fmaxf(vmax, v);
    if (v != v) s_poison[lane] = 2;   // torch.max propagates NaN, fmaxf drops it
  }
  scr[warp * CV_PX + lane] = vmax;
  __syncthreads();
  vmax = scr[lane];
#pragma unroll
  for (int q = 1; q < CV_WARPS; q++) vmax = fmaxf(vmax, scr[q * CV_PX + lane]);
  if (s_poison[lane] == 2) vmax = NAN;
  __syncthreads();

  float npos = 0.0f, best = INFINITY;
  int besti = 0x7fffffff;
  for (int k = k0; k < k1; k++) {
    int o = k * CV_PX + lane;
    float v = cost[o];
    float miss = (v == 0.0f) ? 1.0f : 0.0f;
    float out = v;
    if (a.set_missing_to_max) out = xadd(xmul(v, xsub(1.0f, miss)), xmul(vmax, miss));
    cost[o] = out;
    cnt[o] = miss;
    // compute_confidence_mask(cost_volume * (1 - missing_mask))
    if (xmul(out, xsub(1.0f, miss)) > 0.0f) npos += 1.0f;
    float viz = (out == 0.0f) ? 100.0f : out;       // viz_cost_vol[viz_cost_vol == 0] = 100
    // first min; the first NaN wins like torch.min; an all-+Inf column keeps its first plane
    if (besti == 0x7fffffff || viz < best || (viz != viz && best == best)) { best = viz; besti = k; }
  }
  scr[(0 * CV_WARPS + warp) * CV_PX + lane] = npos;
  scr[(1 * CV_WARPS + warp) * CV_PX + lane] = best;
  reinterpret_cast<int*>(scr)[(2 * CV_WARPS + warp) * CV_PX + lane] = besti;
  __syncthreads();
  float conf_n = 0.0f;
  best = INFINITY; besti = 0x7fffffff;
  bool have = false;
#pragma unroll
  for (int q = 0; q < CV_WARPS; q++) {
    conf_n += scr[(0 * CV_WARPS + q) * CV_PX + lane];
    float v = scr[(1 * CV_WARPS + q) * CV_PX + lane];
    int vi = reinterpret_cast<int*>(scr)[(2 * CV_WARPS + q) * CV_PX + lane];
    if (vi == 0x7fffffff) continue;
    if (!have || v < best || (v != v && best == best)) { best = v; besti = vi; have = true; }
  }
  const int thr = a.num_bins_threshold > 0 ? a.num_bins_threshold : nb;
  const float conf = (conf_n == (float)thr) ? 1.0f : 0.0f;
  besti = min(max(besti, 0), nb - 1);   // never index the bins out of range
  if (pix_ok) {
    const size_t vol = (size_t)b * nb * hw;
    for (int k = k0; k < k1; k++) {
      int o = k * CV_PX + lane;
      float out = cost[o];
      if (a.apply_confidence) out = xmul(out, conf);
      a.cost_volume[vol + (size_t)k * hw + p] = out;
      if (a.missing_mask) a.missing_mask[vol + (size_t)k * hw + p] = cnt[o];
    }
    if (warp == 0) {
      const size_t po = (size_t)b * hw + p;
      if (a.confidence) a.confidence[po] = conf;
      if (a.argmin) a.argmin[po] = besti;
      if (a.lowest_cost) a.lowest_cost[po] = xdiv(1.0f, __ldg(a.bins + besti));
    }
  }
}


// ---- sweep, variant Q: four lanes per pixel ------------------------------------------------------
// A warp owns 8 consecutive pixels; the 4 lanes of a pixel own the four 16-channel chunks (C <= 64).
// Lanes are chunk-major (lane = chunk * 8 + pixel): a quarter-warp - the unit a 128-bit load is processed
// in - then reads the taps of 8 consecutive pixels of ONE channel quad, 128 contiguous bytes.
// Per group of CQ_G = 8 depth planes every lane projects TWO planes (branch-free, so the two chains of
// IEEE divisions interleave); descriptors and chunk sums are exchanged through a per-warp shared-memory
// scratch (scalar stores + broadcast loads: about half the L1 wavefronts of the equivalent shuffles),
// and each lane then sweeps the 8 planes for its chunk with the 2x2x16 register cache (the projecting lane also
// forms the plane's four blend weights, so the sweep reads one float4 per plane).  No block barrier
// inside the sweep, and only 8 (not 32) pixels share a warp's re-fetch decision, so a texel block is
// re-fetched in ~20% instead of ~57% of the warp iterations (profiles/r1_notes.md).
// The bilinear blend and the subtraction run on packed fp32 pairs (FFMA2 / FADD2: two channels per
// instruction, same IEEE results, mal_common.cuh); only the sequential |.| accumulation is scalar:
// 3.5 instead of 6 issue slots per (pixel, plane, channel).  The current features are read in place
// (NCHW, once per lane); only the lookup features go through cv_pack_kernel.
constexpr int CQ_NT = 128;                 // 4 warps = 32 pixels per CTA
constexpr int CQ_G = 8;                    // depth planes per group: every lane projects two of them
// The count plane (how many lookup frames gave a non-zero difference) exists only when it is needed: with one lookup
// frame count == (cost > 0), with cv_min there is no count, and the missing flags of the epilogue fit a lane's
// register when it owns <= 32 planes.  Shared memory not taken is L1 for the texel re-fetches.
__host__ __device__ inline bool cq_needs_counts(int num_lookup, int cv_min, int num_bins) {
  return (num_lookup > 1 && !cv_min) || num_bins > 128;
}
inline size_t cq_smem_bytes(int num_bins, bool counts) {
  return ((size_t)(counts ? 2 : 1) * ((num_bins + 3) / 4 * 4) * CV_PX + 64) * 4;
}
// cost / count accumulators: [plane group][pixel][plane & 3] so that the 32 lanes of a warp
// (8 pixels x 4 planes of a group) hit 32 different banks
__device__ __forceinline__ int cq_idx(int k, int col) { return (k >> 2) * (4 * CV_PX) + col * 4 + (k & 3); }

struct CqTaps { pk2 v[4][2]; };   // one tap of a 16-channel chunk: 4 channel quads x {xy, zw}

__device__ __forceinline__ void cq_ld(pk2* dst, const char* p) {
  const float4 v = ldg4(reinterpret_cast<const float4*>(p));
  dst[0] = pack2(v.x, v.y);
  dst[1] = pack2(v.z, v.w);
}

// DYN: the DynamicDepth variant (dynamicdepth/networks/resnet_encoder.py:148-249): min over the lookup frames
// (cv_min), and an occluded sample's warped features replaced by 1.0 (set_1) or by the pooled maximum of its window
// (pool: 0 inside a blob, else the chunk sums cv_pool_kernel left in the workspace).  Either way the lane's 16-channel
// |.| sum of an occluded sample needs no texels: the two constant cases are formed once per lane.
template <int CONV, int MINB, bool DYN>
__global__ void __launch_bounds__(CQ_NT, MINB) cv_sweep_quad_kernel(const mal_cost_volume_args a, const int Cp, const SizeDiv sdiv) {
  __shared__ CvGeom geom;
  __shared__ __align__(16) float4 s_w[4][CQ_G][8];        // [warp][plane][pixel] bilinear weights {nw, ne, sw, se}
  __shared__ int s_o[4][CQ_G][8];                          // [warp][plane][pixel] tap origin or -1
  __shared__ __align__(16) float4 s_part[4][CQ_G][8];     // [warp][plane][pixel] chunk sums c0..c3
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // chunk of this lane, pixel column in the tile.  A quarter-warp (the unit a 128-bit load is
  // processed in) holds 8 consecutive pixels of ONE chunk: its 8 taps are 128 contiguous bytes.
  const int sub = lane >> 3, pl = lane & 7, col = warp * 8 + pl;
  const int h = a.height, w = a.width, hw = h * w;
  const int nb = a.num_bins, nchunks = Cp / CV_CHUNK, nquads = Cp / 4;
  const int tiles = (hw + CV_PX - 1) / CV_PX;
  const int b = blockIdx.x / tiles;
  const int p = (blockIdx.x - b * tiles) * CV_PX + col;
  const bool pix_ok = p < hw;
  const int py = pix_ok ? p / w : 0, px = pix_ok ? p - py * w : 0;
  const bool inner = pix_ok && py >= 2 && py < h - 2 && px >= 2 && px < w - 2;
  const bool active = sub < nchunks;       // lanes beyond the last chunk still project their plane

  float* cost = reinterpret_cast<float*>(dyn_smem());   // [nb][CV_PX]
  float* cnt = cost + (size_t)((nb + 3) / 4 * 4) * CV_PX;   // only with has_cnt
  const bool cv_min = DYN && a.cv_min;
  const bool has_cnt = cq_needs_counts(a.num_lookup, cv_min, nb);
  for (int k = sub; k < nb; k += 4) {
    cost[cq_idx(k, col)] = cv_min ? 1.0f : 0.0f;
    if (has_cnt) cnt[cq_idx(k, col)] = 0.0f;
  }
  // occlusion handling applies to samples whose matching augmentation is off (aug_mask == 0)
  const float* occ = nullptr;
  if (DYN && a.occ && a.occ_mode != 0 && !(a.aug_mask && __ldg(a.aug_mask + b) != 0.0f)) occ = a.occ + (size_t)b * hw;

  const float4* lookq = reinterpret_cast<const float4*>(a.packed) + (size_t)a.batch * nquads * hw;
  pk2 ncur[4][2];   // -current features of this lane's chunk (w - cur == w + (-cur))
  // A NaN / Inf among the pixel's current features poisons its masked planes too: the reference multiplies
  // |warped - cur|.mean(1) by the edge mask (NaN * 0 = NaN, resnet_encoder.py:211-212) instead of skipping them.
  int poisoned = 0;
  {
    // read straight from the NCHW input (once per lane: 16 scalar loads; channels beyond C are zero
    // padding): the current features need no packing pass
    const float* cur = a.current + (size_t)b * a.channels * hw + p;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int ch = sub * 16 + j * 4 + e;
        v[e] = (pix_ok && active && ch < a.channels) ? __ldg(cur + (size_t)ch * hw) : 0.0f;
        if (!DYN) poisoned |= !(fabsf(v[e]) < INFINITY);
      }
      ncur[j][0] = pack2(-v[0], -v[1]);
      ncur[j][1] = pack2(-v[2], -v[3]);
    }
  }
  poisoned |= __shfl_xor_sync(0xffffffffu, poisoned, 8);
  poisoned |= __shfl_xor_sync(0xffffffffu, poisoned, 16);
  const pk2 one2 = dup2(1.0f), mone2 = dup2(-1.0f);
  const float inv_channels = (a.channels & (a.channels - 1)) == 0 ? 1.0f / (float)a.channels : 0.0f;
  // (DYN) sum over the lane's 16 channels of |1 - cur| and |0 - cur|, in channel order from 0
  float l1_one = 0.0f, l1_zero = 0.0f;
  if (DYN && occ) {
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
      for (int hh = 0; hh < 2; hh++) {
        const pk2 d1 = x2add(one2, ncur[q][hh]);
        l1_one = xadd(xadd(l1_one, fabsf(lo2(d1))), fabsf(hi2(d1)));
        l1_zero = xadd(xadd(l1_zero, fabsf(lo2(ncur[q][hh]))), fabsf(hi2(ncur[q][hh])));
      }
  }
  const bool pooled = DYN && occ && a.occ_mode == MAL_CV_OCC_POOL;
  const size_t dplane = DYN ? cv_desc_plane(a.batch, a.num_lookup, nb, hw) : 0;

  for (int f = 0; f < a.num_lookup; f++) {
    __syncthreads();
    if (tid < 12) {
      geom.P[tid] = kt_entry(a.K + b * 16, a.poses + ((size_t)b * a.num_lookup + f) * 16, tid / 4, tid % 4);
    } else if (tid < 21) {
      int e = tid - 12;
      geom.iK[e] = a.inv_K[b * 16 + (e / 3) * 4 + e % 3];
    } else if (tid == 21) {
      const float* T = a.poses + ((size_t)b * a.num_lookup + f) * 16;
      float s = 0.0f;
      for (int e = 0; e < 16; e++) s += T[e];
      geom.live = (s != 0.0f) ? 1 : 0;
    }
    __syncthreads();
    if (!geom.live) continue;
    const Ray ray = pixel_ray(geom.iK, (float)px, (float)py);
    const float4* lqc = lookq + ((size_t)b * a.num_lookup + f) * nquads * hw + (size_t)sub * 4 * hw;
    const char* lbase = reinterpret_cast<const char*>(lqc);
    const size_t row_stride = (size_t)w * 16, quad_stride = (size_t)hw * 16;
    CqTaps t00, t01, t10, t11;
    int coff = -1;

    for (int k0 = 0; k0 < nb; k0 += CQ_G) {
      // ---- this lane's two planes of the group: projection descriptors --------------------------
      // branch-free, so that the two dependent chains (each ends in four IEEE divisions) interleave
      int off[2];
      float tx[2], ty[2];
      int pfl[2] = {0, 0};   // (DYN, pool) the pre-passes' verdicts, requested ahead of the projection arithmetic
      if (DYN && pooled) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int kk = k0 + sub + 4 * u;
          if (pix_ok && kk < nb)
            pfl[u] = __ldg(reinterpret_cast<const int*>(a.desc) + 2 * dplane + ((size_t)b * a.num_lookup + f) * nb * hw +
                           (size_t)kk * hw + p);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int kk = k0 + sub + 4 * u;
        const float depth = __ldg(a.bins + min(kk, nb - 1));
        const GridPoint gp = project_grid<CONV>(geom.P, ray, depth, a.eps, h, w, &sdiv);
        const float xv = xmul(xadd(xmul(gp.gx, 0.5f), 0.5f), (float)(w - 1));
        const float yv = xmul(xadd(xmul(gp.gy, 0.5f), 0.5f), (float)(h - 1));
        const bool ok = inner && kk < nb && xv >= 2.0f && xv <= (float)(w - 2) && yv >= 2.0f && yv <= (float)(h - 2);
        const float ux = unnormalize<CONV>(gp.gx, w), uy = unnormalize<CONV>(gp.gy, h);
        const float x0 = floorf(ux), y0 = floorf(uy);
        // with the edge mask on, all four taps are inside the image; the clamp guards pathological inputs
        const int xi = min(max((int)x0, 0), w - 2), yi = min(max((int)y0, 0), h - 2);
        off[u] = ok ? yi * w + xi : -1;
        tx[u] = xsub(ux, x0);
        ty[u] = xsub(uy, y0);
        if (DYN && occ && ok) {
          if (pooled) {   // the pre-passes' verdicts (cv_project_kernel / cv_interior_kernel)
            if (pfl[u] & CV_DESC_OCC) off[u] |= CV_OCC_BIT | ((pfl[u] & CV_DESC_ZERO) ? CV_ZERO_BIT : 0);
          } else if (occluded_at<CONV>(occ, gp, h, w, a.pool_th)) {
            off[u] |= CV_OCC_BIT;
          }
        }
      }
      if (!__any_sync(0xffffffffu, (off[0] & off[1]) >= 0 || poisoned)) continue;   // 8 pixels x 8 planes all masked
#pragma unroll
      for (int u = 0; u < 2; u++) {
        // the four blend weights, formed once per (pixel, plane) here instead of once per lane in the sweep
        // (the same operations: 1 - t as fma(t, -1, 1) = RN(1 - t))
        const float e = xsub(1.0f, tx[u]), sfr = xsub(1.0f, ty[u]);
        s_w[warp][sub + 4 * u][pl] = make_float4(xmul(sfr, e), xmul(sfr, tx[u]), xmul(ty[u], e), xmul(ty[u], tx[u]));
        s_o[warp][sub + 4 * u][pl] = off[u];
      }
      __syncwarp();
      // ---- the planes of the group, this lane's chunk ------------------------------------------
      float* part = reinterpret_cast<float*>(&s_part[warp][0][0]);
#pragma unroll 1
      for (int j0 = 0; j0 < CQ_G; j0 += 4) {
        float acc[4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int j = j0 + jj;
          const int o = s_o[warp][j][pl];
          acc[jj] = 0.0f;
          if (DYN && o >= 0 && (o & CV_OCC_BIT)) {   // an occluded sample: no texels needed
            if (active) {
              if (a.occ_mode == MAL_CV_OCC_SET_1) acc[jj] = l1_one;          // warped[mask] = 1.0
              else if (o & CV_ZERO_BIT) acc[jj] = l1_zero;                    // pooled value 0 inside a blob
              else acc[jj] = __ldg(a.desc + (4 + (size_t)sub) * dplane + ((size_t)b * a.num_lookup + f) * nb * hw +
                                   (size_t)(k0 + j) * hw + p);              // the rim: cv_pool_kernel's chunk sum
            }
          } else if (o >= 0 && active) {
            if (o != coff) {
              coff = o;
              // two 64-bit pointer increments per channel quad; the taps are fixed offsets from them
              const char* r0 = lbase + (size_t)(unsigned)o * 16;
#pragma unroll
              for (int q = 0; q < 4; q++) {
                const char* r1 = r0 + row_stride;
                cq_ld(t00.v[q], r0); cq_ld(t01.v[q], r0 + 16);
                cq_ld(t10.v[q], r1); cq_ld(t11.v[q], r1 + 16);
                r0 += quad_stride;
              }
            }
            const float4 wv = s_w[warp][j][pl];
            const pk2 nw = dup2(wv.x), ne = dup2(wv.y), sw = dup2(wv.z), se = dup2(wv.w);
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < 4; q++) {
#pragma unroll
              for (int hh = 0; hh < 2; hh++) {
                const pk2 wq = x2fma(t11.v[q][hh], se, x2fma(t10.v[q][hh], sw, x2fma(t01.v[q][hh], ne, x2mul(t00.v[q][hh], nw))));
                const pk2 d = x2add(wq, ncur[q][hh]);
                s = xadd(s, fabsf(lo2(d)));
                s = xadd(s, fabsf(hi2(d)));
              }
            }
            acc[jj] = s;
          }
        }
        // chunk sums of each plane; inactive / masked lanes contribute +0
#pragma unroll
        for (int jj = 0; jj < 4; jj++) part[((j0 + jj) * 8 + pl) * 4 + sub] = acc[jj];
      }
      __syncwarp();
      // ---- gather the chunk sums in the reference's order, mean over channels, accumulate -------
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (off[u] >= 0) {
          const float4 c = s_part[warp][sub + 4 * u][pl];
          const float s_mine = xadd(xadd(xadd(c.x, c.y), c.z), c.w);
          // .mean(1); edge mask == 1.  A power-of-two channel count divides exactly by multiplying.
          const float diff = inv_channels != 0.0f ? xmul(s_mine, inv_channels) : xdiv(s_mine, (float)a.channels);
          const int o = cq_idx(k0 + sub + 4 * u, col);
          if (cv_min) {   // diffs[diffs == 0] = 1.0; cost_volume = minimum(diffs, cost_volume)
            cost[o] = fminf(diff == 0.0f ? 1.0f : diff, cost[o]);
          } else {
            cost[o] = xadd(cost[o], diff);
            if (has_cnt && diff > 0.0f) cnt[o] = xadd(cnt[o], 1.0f);
          }
        } else if (poisoned && pix_ok && k0 + sub + 4 * u < nb) {
          cost[cq_idx(k0 + sub + 4 * u, col)] = NAN;   // masked plane of a poisoned pixel: NaN * 0
        }
      }
    }
  }
  __syncwarp();

  // ---- epilogue: lane `sub` of a pixel owns planes sub, sub+4, ... ---------------------------------
  float vmax = -INFINITY;
  int has_nan = 0;
  for (int k = sub; k < nb; k += 4) {
    const int o = cq_idx(k, col);
    // (one lookup frame: counts = (diff > 0) = (cost > 0); a NaN cost counts 0 either way)
    const float c0 = cost[o];
    const float v = cv_min ? (c0 == 1.0f ? 0.0f : c0)                                             // cost_volume[cost_volume == 1] = 0
                           : xdiv(c0, xadd(has_cnt ? cnt[o] : (c0 > 0.0f ? 1.0f : 0.0f), 1e-7f));   // cost_volume / (counts + 1e-7)
    cost[o] = v;
    vmax = fmaxf(vmax, v);
    has_nan |= (v != v);
  }
  vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 8));
  vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 16));
  has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, 8);
  has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, 16);
  if (has_nan) vmax = NAN;   // torch.max propagates NaN, fmaxf drops it
  float npos = 0.0f, best = INFINITY;
  int besti = 0x7fffffff;
  unsigned missbits = 0u;   // (no count plane) the missing flags of this lane's planes sub, sub + 4, ...
  for (int k = sub; k < nb; k += 4) {
    const int o = cq_idx(k, col);
    const float v = cost[o];
    const float miss = (v == 0.0f) ? 1.0f : 0.0f;
    float out = v;
    if (a.set_missing_to_max) out = xadd(xmul(v, xsub(1.0f, miss)), xmul(vmax, miss));
    cost[o] = out;
    if (has_cnt) cnt[o] = miss;
    else if (v == 0.0f) missbits |= 1u << (k >> 2);
    if (xmul(out, xsub(1.0f, miss)) > 0.0f) npos += 1.0f;
    const float viz = (out == 0.0f) ? 100.0f : out;
    // first min; the first NaN wins like torch.min (and like cv_sweep_kernel); an all-+Inf column keeps
    // its first plane
    if (besti == 0x7fffffff || viz < best || (viz != viz && best == best)) { best = viz; besti = k; }
  }
#pragma unroll
  for (int m = 8; m <= 16; m <<= 1) {
    npos += __shfl_xor_sync(0xffffffffu, npos, m);
    const float ov = __shfl_xor_sync(0xffffffffu, best, m);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, m);
    const bool on = ov != ov, mn = best != best;
    const bool take = on ? (!mn || oi < besti) : (!mn && (ov < best || (ov == best && oi < besti)));
    if (take) { best = ov; besti = oi; }
  }
  besti = min(max(besti, 0), nb - 1);   // never index the bins out of range
  const int thr = a.num_bins_threshold > 0 ? a.num_bins_threshold : nb;
  const float conf = (npos == (float)thr) ? 1.0f : 0.0f;
  if (pix_ok) {
    const size_t vol = (size_t)b * nb * hw;
    for (int k = sub; k < nb; k += 4) {
      const int o = cq_idx(k, col);
      float out = cost[o];
      if (a.apply_confidence) out = xmul(out, conf);
      a.cost_volume[vol + (size_t)k * hw + p] = out;
      if (a.missing_mask)
        a.missing_mask[vol + (size_t)k * hw + p] = has_cnt ? cnt[o] : (float)((missbits >> (k >> 2)) & 1u);
    }
    if (sub == 0) {
      const size_t po = (size_t)b * hw + p;
      if (a.confidence) a.confidence[po] = conf;
      if (a.argmin) a.argmin[po] = besti;
      if (a.lowest_cost) a.lowest_cost[po] = xdiv(1.0f, __ldg(a.bins + besti));
    }
  }
}

}  // namespace mal

using namespace mal;

extern "C" size_t mal_cost_volume_workspace_floats(int batch, int channels, int height, int width, int num_lookup) {
  return (size_t)batch * (1 + num_lookup) * cv_padded_channels(channels) * height * width;
}

extern "C" size_t mal_cost_volume_desc_floats(int batch, int channels, int num_lookup, int num_bins, int height,
                                              int width) {
  return cv_desc_layout(batch, num_lookup, num_bins, height * width, cv_padded_channels(channels)).total;
}

extern "C" int mal_cost_volume_forward(const mal_cost_volume_args* args, mal_stream_t stream) {
  MAL_REQUIRE(args != nullptr, "mal_cost_volume_forward: args is NULL");
  const mal_cost_volume_args& a = *args;
  MAL_REQUIRE(a.batch > 0 && a.channels > 0 && a.height >= 5 && a.width >= 5 && a.num_lookup > 0 && a.num_bins > 0,
              "mal_cost_volume_forward: bad shape B=%d C=%d h=%d w=%d F=%d bins=%d", a.batch, a.channels, a.height,
              a.width, a.num_lookup, a.num_bins);
  MAL_REQUIRE(a.channels <= 256, "mal_cost_volume_forward: %d channels (> 256) changes the reference's summation tree",
              a.channels);
  MAL_REQUIRE(a.convention == MAL_CONV_MANYDEPTH || a.convention == MAL_CONV_DUALREFINE,
              "mal_cost_volume_forward: bad convention %d", a.convention);
  MAL_REQUIRE(a.occ_mode == MAL_CV_OCC_NONE || a.occ_mode == MAL_CV_OCC_SET_1 || a.occ_mode == MAL_CV_OCC_POOL,
              "mal_cost_volume_forward: bad occ_mode %d", a.occ_mode);
  if (a.occ_mode == MAL_CV_OCC_POOL) MAL_REQUIRE(a.pool_radius >= 0 && a.pool_radius <= 3, "mal_cost_volume_forward: pool_radius %d", a.pool_radius);
  MAL_REQUIRE(a.current && a.lookup && a.poses && a.K && a.inv_K && a.bins && a.cost_volume && a.packed,
              "mal_cost_volume_forward: current/lookup/poses/K/inv_K/bins/cost_volume/packed are required");
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = cv_padded_channels(a.channels);
  const int hw = a.height * a.width;
  const bool dyn = a.cv_min || (a.occ && a.occ_mode != MAL_CV_OCC_NONE);
  // kernel choice: the four-lanes-per-pixel sweep handles up to 4 chunks (C <= 64), with or without the
  // DynamicDepth extras; MAL_CV_KERNEL=lane forces the one-pixel-per-lane kernel (tuning only)
  bool quad = Cp / CV_CHUNK <= 4;
  const bool pool = a.occ && a.occ_mode == MAL_CV_OCC_POOL;
  if (const char* e = getenv("MAL_CV_KERNEL")) quad = quad && e[0] != 'l';
  {
    long long total = (long long)a.batch * (Cp / 4) * hw;
    if (!quad || pool)   // the quad sweep reads the current features in place (NCHW); cv_pool_kernel the packed ones
      launch(cv_pack_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, a.current,
             reinterpret_cast<float4*>(a.packed), a.channels, Cp, hw, total);
    long long total_l = total * a.num_lookup;
    launch(cv_pack_kernel, dim3((unsigned)((total_l + 255) / 256)), dim3(256), 0, st, a.lookup,
           reinterpret_cast<float4*>(a.packed) + total, a.channels, Cp, hw, total_l);
    int rc = check_launch("cv_pack_kernel");
    if (rc) return rc;
  }
  const int tiles = (hw + CV_PX - 1) / CV_PX;
  const SizeDiv sdiv = size_div(a.height, a.width, a.convention);
  if (a.occ && a.occ_mode == MAL_CV_OCC_POOL) {
    // the pool fill's descriptor volume: every (lookup frame, bin, pixel) projected once, blob interiors marked
    MAL_REQUIRE(a.desc != nullptr, "mal_cost_volume_forward: MAL_CV_OCC_POOL needs the desc workspace "
                "(mal_cost_volume_desc_floats)");
    const size_t rows = (size_t)a.batch * a.num_lookup * a.num_bins;
    MAL_REQUIRE(rows < (1u << 31) && (hw + 255) / 256 <= 65535, "mal_cost_volume_forward: descriptor grid too large");
    dim3 pgrid((unsigned)rows, (unsigned)((hw + 255) / 256));
    dim3 jgrid((unsigned)((size_t)a.batch * a.num_lookup * ((a.num_bins + CV_PJ - 1) / CV_PJ)), (unsigned)((hw + 255) / 256));
    if (a.convention == MAL_CONV_MANYDEPTH) launch(cv_project_kernel<MAL_CONV_MANYDEPTH>, jgrid, dim3(256), 0, st, a, sdiv);
    else launch(cv_project_kernel<MAL_CONV_DUALREFINE>, jgrid, dim3(256), 0, st, a, sdiv);
    const CvDescLayout L = cv_desc_layout(a.batch, a.num_lookup, a.num_bins, hw, Cp);
    MAL_REQUIRE(L.plane < (1u << 31), "mal_cost_volume_forward: descriptor volume too large for 32-bit sample indices");
    cudaMemsetAsync(reinterpret_cast<int*>(a.desc) + L.counters, 0, 4 * sizeof(int), st);
    launch(cv_interior_kernel, pgrid, dim3(256), 0, st, a, Cp);
    {
      const long long total_l = (long long)a.batch * a.num_lookup * (Cp / 4) * hw;
      launch(cv_pack_cm_kernel, dim3((unsigned)((total_l + 255) / 256)), dim3(256), 0, st, a.lookup,
             reinterpret_cast<float4*>(a.desc + L.cm), a.channels, Cp, hw, total_l);
    }
    launch(cv_sample_kernel, dim3(148 * 8), dim3(256), 0, st, a, Cp);
    if (a.pool_radius == 1) launch(cv_pool_kernel<1>, dim3(148 * 8), dim3(256), 0, st, a, Cp);
    else launch(cv_pool_kernel<0>, dim3(148 * 8), dim3(256), 0, st, a, Cp);
    int rc = check_launch("cv_project_kernel / cv_interior_kernel / cv_pool_kernel");
    if (rc) return rc;
  }
  const size_t smem = cv_smem_bytes(Cp / CV_CHUNK, a.num_bins);
  MAL_REQUIRE(smem <= 227 * 1024, "mal_cost_volume_forward: %d bins x %d channels need %zu B of shared memory",
              a.num_bins, a.channels, smem);
  dim3 grid((unsigned)((size_t)a.batch * tiles));
  // resident CTAs per SM the register allocation is tuned for (profiles/r1_notes.md); the
  // environment override exists for tuning runs only
  int minb = CV_DEFAULT_MINB;
  if (const char* e = getenv("MAL_CV_MINB")) minb = atoi(e);
#define MAL_CV_LAUNCH(CONV_)                                                                         \
  do {                                                                                               \
    if (dyn && minb <= 3) launch(cv_sweep_kernel<CONV_, 3, true>, grid, dim3(CV_NT), smem, st, a, Cp, sdiv);     \
    else if (dyn) launch(cv_sweep_kernel<CONV_, 4, true>, grid, dim3(CV_NT), smem, st, a, Cp, sdiv);        \
    else if (minb <= 3) launch(cv_sweep_kernel<CONV_, 3, false>, grid, dim3(CV_NT), smem, st, a, Cp, sdiv); \
    else if (minb == 4) launch(cv_sweep_kernel<CONV_, 4, false>, grid, dim3(CV_NT), smem, st, a, Cp, sdiv); \
    else launch(cv_sweep_kernel<CONV_, 5, false>, grid, dim3(CV_NT), smem, st, a, Cp, sdiv);                \
  } while (0)
  if (quad) {
    const size_t qsmem = cq_smem_bytes(a.num_bins, cq_needs_counts(a.num_lookup, a.cv_min, a.num_bins));
#define MAL_CQ_GO(CONV_, DYN_)                                                                                  \
  do {                                                                                                          \
    if (minb <= 3) launch(cv_sweep_quad_kernel<CONV_, 3, DYN_>, grid, dim3(CQ_NT), qsmem, st, a, Cp, sdiv);         \
    else if (minb == 4) launch(cv_sweep_quad_kernel<CONV_, 4, DYN_>, grid, dim3(CQ_NT), qsmem, st, a, Cp, sdiv);    \
    else launch(cv_sweep_quad_kernel<CONV_, 5, DYN_>, grid, dim3(CQ_NT), qsmem, st, a, Cp, sdiv);                   \
  } while (0)
#define MAL_CQ_LAUNCH(CONV_)            \
  do {                                  \
    if (dyn) MAL_CQ_GO(CONV_, true);    \
    else MAL_CQ_GO(CONV_, false);       \
  } while (0)
    if (a.convention == MAL_CONV_MANYDEPTH) MAL_CQ_LAUNCH(MAL_CONV_MANYDEPTH);
    else MAL_CQ_LAUNCH(MAL_CONV_DUALREFINE);
#undef MAL_CQ_GO
#undef MAL_CQ_LAUNCH
    return check_launch("cv_sweep_quad_kernel");
  }
  if (a.convention == MAL_CONV_MANYDEPTH) MAL_CV_LAUNCH(MAL_CONV_MANYDEPTH);
  else MAL_CV_LAUNCH(MAL_CONV_DUALREFINE);
#undef MAL_CV_LAUNCH
  return check_launch("cv_sweep_kernel");
}
