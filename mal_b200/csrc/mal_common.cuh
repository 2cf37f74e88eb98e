// mal_common.cuh - shared plumbing for the mal_b200 CUDA sources (sm_100a).
//
// * error reporting for the C-ABI (no C++ exception crosses the boundary)
// * kernel launch helper (also the hook the host-side test emulator uses)
// * exact, never-contracted fp32 arithmetic (x*) for everything that feeds a
//   selection comparison, see DESIGN.md "Arithmetic contract"
// * block reductions
#pragma once

#ifdef MAL_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mal_b200.h"

namespace mal {

// ------------------------------------------------------------------ errors
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MAL_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return MAL_OK;
}
#define MAL_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::mal::fail(MAL_ERR_ARGUMENT, __VA_ARGS__); \
  } while (0)

// ------------------------------------------------------------------ launch
// One entry point for every kernel launch so the host emulator can intercept it.
template <class... KArgs, class... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                   Args... args) {
#ifdef MAL_EMU
  (void)stream;
  emu::launch(grid, block, smem, [=]() { kernel(args...); });
#else
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kernel<<<grid, block, smem, stream>>>(args...);
#endif
}

// dynamic shared memory base, 16-byte aligned
__device__ __forceinline__ unsigned char* dyn_smem() {
#ifdef MAL_EMU
  unsigned char* p = emu::current()->smem.data();
  return p + ((16 - ((uintptr_t)p & 15)) & 15);
#else
  extern __shared__ __align__(16) unsigned char mal_dyn_smem_[];
  return mal_dyn_smem_;
#endif
}

// ------------------------------------------------------------------ exact fp32
// Round-to-nearest single operations that the compiler may not fuse or reorder.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// Correctly rounded x / C for a compile-time constant C (Markstein: q = RN(x*r), rem = fma(-C,q,x),
// RN(q + rem*r)); identical to IEEE division for every finite x except the sign of -0 (checked
// exhaustively over all 2^32 inputs: tests/test_exact_arithmetic.py with MAL_EXHAUSTIVE=1).  3 instructions instead of ~10.
template <int C>
__device__ __forceinline__ float xdivc(float x) {
  const float r = 1.0f / (float)C;
  float q = __fmul_rn(x, r);
  float rem = __fmaf_rn(-(float)C, q, x);
  return __fmaf_rn(rem, r, q);
}

// ------------------------------------------------------------------ packed exact fp32 (f32x2)
// sm_100a issues two independent IEEE round-to-nearest fp32 operations per lane with one
// instruction (FFMA2 / FADD2 / FMUL2 on a 64-bit register pair).  Each half is the same
// correctly-rounded operation as its scalar x* counterpart, so the arithmetic contract is
// unchanged; what changes is the issue-slot cost (tools/ubench/f32x2.cu: an FFMA2 holds the fma
// pipe for two cycles but one issue slot, so loads, shuffles and integer work issue beside it).
//
// HAZARD (ptxas 12.9, measured: tools/ubench/x2check.cu): unlike the scalar forms, ptxas contracts an
// explicit mul.rn.f32x2 whose result feeds an add.rn/sub.rn.f32x2 into ONE FFMA2 (single rounding),
// with or without --fmad=false, and it sees through fma(p, 1, s) and fma(a, b, -0) rewrites.  A
// product that feeds a packed add or subtract must therefore use x2mul_nf (".ftz" on the multiply
// blocks the contraction; it differs from IEEE only when an operand or the product is subnormal).
// x2mul feeding x2fma's addend, or a store, is safe: there is nothing to contract.
typedef unsigned long long pk2;   // {lo, hi} fp32 pair in an aligned register pair
#ifdef MAL_EMU
__device__ __forceinline__ pk2 pack2(float lo, float hi) {
  pk2 r; float v[2] = {lo, hi}; memcpy(&r, v, 8); return r;
}
__device__ __forceinline__ void unpack2(pk2 p, float& lo, float& hi) {
  float v[2]; memcpy(v, &p, 8); lo = v[0]; hi = v[1];
}
__device__ __forceinline__ pk2 x2mul(pk2 a, pk2 b) {
  float a0, a1, b0, b1; unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(xmul(a0, b0), xmul(a1, b1));
}
__device__ __forceinline__ pk2 x2add(pk2 a, pk2 b) {
  float a0, a1, b0, b1; unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(xadd(a0, b0), xadd(a1, b1));
}
__device__ __forceinline__ pk2 x2fma(pk2 a, pk2 b, pk2 c) {
  float a0, a1, b0, b1, c0, c1; unpack2(a, a0, a1); unpack2(b, b0, b1); unpack2(c, c0, c1);
  return pack2(xfma(a0, b0, c0), xfma(a1, b1, c1));
}
__device__ __forceinline__ pk2 x2sub(pk2 a, pk2 b) {
  float a0, a1, b0, b1; unpack2(a, a0, a1); unpack2(b, b0, b1);
  return pack2(xsub(a0, b0), xsub(a1, b1));
}
__device__ __forceinline__ pk2 x2mul_nf(pk2 a, pk2 b) { return x2mul(a, b); }
#else
__device__ __forceinline__ pk2 pack2(float lo, float hi) {
  pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(pk2 p, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p));
}
__device__ __forceinline__ pk2 x2mul(pk2 a, pk2 b) {
  pk2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ pk2 x2add(pk2 a, pk2 b) {
  pk2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ pk2 x2fma(pk2 a, pk2 b, pk2 c) {
  pk2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ pk2 x2sub(pk2 a, pk2 b) {
  pk2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ pk2 x2mul_nf(pk2 a, pk2 b) {
  pk2 d; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
#endif
__device__ __forceinline__ pk2 dup2(float v) { return pack2(v, v); }
__device__ __forceinline__ float lo2(pk2 p) { float a, b; unpack2(p, a, b); return a; }
__device__ __forceinline__ float hi2(pk2 p) { float a, b; unpack2(p, a, b); return b; }
// |p| per half (folds into the abs modifier of the consuming FADD2)
__device__ __forceinline__ pk2 abs2(pk2 p) { float a, b; unpack2(p, a, b); return pack2(fabsf(a), fabsf(b)); }
// packed form of xdivc: correctly rounded p / C per half
template <int C>
__device__ __forceinline__ pk2 x2divc(pk2 x) {
  const pk2 r = dup2(1.0f / (float)C);
  const pk2 q = x2mul(x, r);
  const pk2 rem = x2fma(dup2(-(float)C), q, x);
  return x2fma(rem, r, q);
}

// ------------------------------------------------------------------ reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums of N <= 32 per-lane values over the warp, all at once: lane j < N returns sum over the lanes of v[j] (the other
// lanes return padding).  Each butterfly step halves the values a lane still carries (a lane keeps the half its
// partner sends the other half of), so the whole reduction is 31 shuffles instead of 5 N (N = 24: 31 against 120).
// Fixed order: deterministic.
template <int N>
__device__ __forceinline__ float warp_sum_transposed(const float (&v)[N], int lane) {
  static_assert(N <= 32, "one value per lane at most");
  float a[16], b[8], c[4], d[2];
  const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const float lo = i < N ? v[i < N ? i : 0] : 0.0f, hi = i + 16 < N ? v[i + 16 < N ? i + 16 : 0] : 0.0f;
    a[i] = (b16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b16 ? lo : hi, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) b[i] = (b8 ? a[i + 8] : a[i]) + __shfl_xor_sync(0xffffffffu, b8 ? a[i] : a[i + 8], 8);
#pragma unroll
  for (int i = 0; i < 4; i++) c[i] = (b4 ? b[i + 4] : b[i]) + __shfl_xor_sync(0xffffffffu, b4 ? b[i] : b[i + 4], 4);
#pragma unroll
  for (int i = 0; i < 2; i++) d[i] = (b2 ? c[i + 2] : c[i]) + __shfl_xor_sync(0xffffffffu, b2 ? c[i] : c[i + 2], 2);
  return (b1 ? d[1] : d[0]) + __shfl_xor_sync(0xffffffffu, b1 ? d[0] : d[1], 1);
}

__device__ __forceinline__ int reflect_index(int i, int n) {
  // ReflectionPad2d(1) index map, also safe for |overshoot| <= n-1
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

}  // namespace mal
