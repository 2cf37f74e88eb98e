// mal_tma.cuh - TMA (cp.async.bulk.tensor) tile loads + mbarrier completion for sm_100a.
//
// A TileMap describes a contiguous fp32 tensor (W, H, N) - N image planes of H x W - and a fixed box
// (bw, bh, bn).  tma_load_3d() makes the TMA unit copy the box whose corner is (x, y, n) into shared memory
// as dense [bn][bh][bw] floats; elements outside the tensor arrive as 0 (the callers reflect the one row /
// column ReflectionPad2d needs afterwards).  One elected thread issues the copies; everybody waits on the
// mbarrier, so no thread spends registers or issue slots moving the tile (SASS: UTMALDG, SYNCS).
//
// Host side: the descriptor is a CUtensorMap built with cuTensorMapEncodeTiled (looked up through
// cudaGetDriverEntryPoint, so the library does not link libcuda) and passed to the kernel inside a
// __grid_constant__ parameter.  TMA needs a 16-byte aligned base, W * 4 a multiple of 16 and a box no larger
// than the tensor; tile_map_encode() returns false otherwise and the caller falls back to the plain loader.
// The box corner's x must be a multiple of 4 elements (16 bytes) as well - the kernel's tiling guarantees it.
//
// The CPU twin (MAL_EMU) keeps the geometry in a plain struct and copies synchronously.
#pragma once
#include "mal_common.cuh"

#ifndef MAL_EMU
#include <cuda.h>
#include <cudaTypedefs.h>
#endif

namespace mal {

#ifdef MAL_EMU
struct alignas(64) TileMap {
  const float* base;
  int W, H, N, bw, bh, bn;
};
inline bool tile_map_encode(TileMap* m, const float* base, int W, int H, int N, int bw, int bh, int bn) {
  if (base == nullptr || ((uintptr_t)base & 15) || (W & 3) || W < bw || H < bh || N < bn) return false;
  *m = TileMap{base, W, H, N, bw, bh, bn};
  return true;
}
__device__ __forceinline__ void mbar_init(unsigned long long*, int) {}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, unsigned) {}
__device__ __forceinline__ void mbar_wait(unsigned long long*, unsigned) {}
__device__ __forceinline__ void tma_load_3d(float* dst, const TileMap* m, int x, int y, int n, unsigned long long*) {
  // the hardware raises "illegal instruction" when the box does not start on a 16-byte boundary of the row
  // (measured on B200: tools/ubench/tma_check.cu) or the destination is not 128-byte aligned
  if ((x & 3) != 0 || ((uintptr_t)dst & 127) != 0) { fprintf(stderr, "cuda_emu: misaligned TMA box (x=%d, dst=%p)\n", x, (void*)dst); abort(); }
  for (int c = 0; c < m->bn; c++)
    for (int j = 0; j < m->bh; j++)
      for (int i = 0; i < m->bw; i++) {
        const int gx = x + i, gy = y + j, gn = n + c;
        const bool in = gx >= 0 && gx < m->W && gy >= 0 && gy < m->H && gn >= 0 && gn < m->N;
        dst[(c * m->bh + j) * m->bw + i] = in ? m->base[((size_t)gn * m->H + gy) * m->W + gx] : 0.0f;
      }
}
#else
typedef CUtensorMap TileMap;

inline PFN_cuTensorMapEncodeTiled_v12000 tile_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

inline bool tile_map_encode(TileMap* m, const float* base, int W, int H, int N, int bw, int bh, int bn) {
  if (base == nullptr || ((uintptr_t)base & 15) || (W & 3) || (bw & 3) || W < bw || H < bh || N < bn) return false;
  if (bw > 256 || bh > 256 || bn > 256) return false;
  PFN_cuTensorMapEncodeTiled_v12000 enc = tile_map_encoder();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
// box corner (x, y, n) -> dst (128-byte aligned shared memory), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_3d(float* dst, const TileMap* m, int x, int y, int n, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_addr(dst)),
      "l"(reinterpret_cast<unsigned long long>(m)), "r"(x), "r"(y), "r"(n), "r"(smem_addr(bar))
      : "memory");
}
#endif

}  // namespace mal
