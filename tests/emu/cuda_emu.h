// cuda_emu.h - a tiny CPU emulator of the CUDA subset the mal_b200 kernels use.
//
// TEST INFRASTRUCTURE ONLY.  It lets the *same* .cu sources that nvcc compiles for
// sm_100a be compiled by g++ and executed on the host so that indexing, tiling,
// barrier placement and the bit-exact arithmetic contract can be checked in the
// CPU-only test suite (there is no GPU in the build container).  The product library
// libmal_b200.so never contains any of this and the mal_b200 package never loads the
// emulator build: see tests/emu/build_emu.py.
//
// Model: one OS thread; every CUDA thread of a block is a ucontext fiber; blocks run
// one after another.  __syncthreads(), __syncwarp() and the warp shuffles are yield
// points handled by a scheduler that enforces barrier semantics.  Global/shared
// atomics are plain read-modify-writes (execution is sequential).
#pragma once
#ifndef MAL_EMU
#error "cuda_emu.h is only for the MAL_EMU host build"
#endif

#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

typedef struct CUstream_st* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

namespace emu {

enum State { RUNNABLE = 0, AT_WARP = 1, AT_SYNC = 2, DONE = 3 };

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  int state = RUNNABLE;
  uint3 tid;
  unsigned linear;
};

struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  std::function<void()> body;
  std::vector<unsigned char> smem;
  std::vector<uint64_t> xchg;  // one 8-byte slot per thread for shuffles / votes
};

inline Block*& current() { static Block* b = nullptr; return b; }
inline uint3& tidx() { static uint3 v; return v; }
inline uint3& bidx() { static uint3 v; return v; }
inline dim3& bdim() { static dim3 v; return v; }
inline dim3& gdim() { static dim3 v; return v; }

static const size_t kStack = 256 * 1024;

inline void yield(int st) {
  Block* b = current();
  Fiber& f = b->fibers[b->cur];
  f.state = st;
  swapcontext(&f.ctx, &b->sched);
}

inline void trampoline() {
  Block* b = current();
  b->body();
  b->fibers[b->cur].state = DONE;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

inline void run_block(Block& blk, dim3 block) {
  unsigned n = block.x * block.y * block.z;
  if (blk.fibers.size() != n) {
    for (auto& f : blk.fibers) free(f.stack);
    blk.fibers.assign(n, Fiber());
    for (auto& f : blk.fibers) f.stack = (char*)malloc(kStack);
  }
  blk.xchg.assign(n, 0);
  for (unsigned i = 0; i < n; i++) {
    Fiber& f = blk.fibers[i];
    f.state = RUNNABLE;
    f.linear = i;
    f.tid = uint3{i % block.x, (i / block.x) % block.y, i / (block.x * block.y)};
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStack;
    f.ctx.uc_link = &blk.sched;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  current() = &blk;
  unsigned done = 0;
  while (done < n) {
    bool progressed = false;
    for (unsigned i = 0; i < n; i++) {
      Fiber& f = blk.fibers[i];
      if (f.state == RUNNABLE || f.state == AT_WARP) {
        blk.cur = (int)i;
        tidx() = f.tid;
        swapcontext(&blk.sched, &f.ctx);
        progressed = true;
      }
    }
    done = 0;
    unsigned at_sync = 0;
    for (auto& f : blk.fibers) { done += f.state == DONE; at_sync += f.state == AT_SYNC; }
    if (at_sync && at_sync + done == n) {
      // CUDA requires every non-exited thread to reach the barrier
      for (auto& f : blk.fibers) if (f.state == AT_SYNC) f.state = RUNNABLE;
    } else if (!progressed && done < n) {
      fprintf(stderr, "cuda_emu: deadlock (%u at barrier, %u done of %u)\n", at_sync, done, n);
      abort();
    }
  }
}

inline Block& the_block() { static Block b; return b; }

template <class Body>
inline void launch(dim3 grid, dim3 block, size_t smem, Body body) {
  Block& blk = the_block();
  blk.smem.assign(smem + 16, 0xCD);  // poison: kernels must not rely on zeroed smem
  blk.body = body;
  bdim() = block;
  gdim() = grid;
  for (unsigned z = 0; z < grid.z; z++)
    for (unsigned y = 0; y < grid.y; y++)
      for (unsigned x = 0; x < grid.x; x++) {
        bidx() = uint3{x, y, z};
        memset(blk.smem.data(), 0xCD, blk.smem.size());
        run_block(blk, block);
      }
}

inline unsigned lane() { return current()->fibers[current()->cur].linear & 31u; }
inline unsigned warp_base() { return current()->fibers[current()->cur].linear & ~31u; }

template <class T>
inline T exchange(T v, unsigned src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  Block* b = current();
  unsigned me = b->fibers[b->cur].linear;
  uint64_t bits = 0;
  memcpy(&bits, &v, sizeof(T));
  b->xchg[me] = bits;
  yield(AT_WARP);  // everyone in the warp publishes
  unsigned src = (me & ~31u) + (src_lane & 31u);
  T r = v;
  if (src < b->fibers.size() && b->fibers[src].state != DONE) memcpy(&r, &b->xchg[src], sizeof(T));
  yield(AT_WARP);  // everyone reads before the slot is reused
  return r;
}

inline unsigned ballot(int pred) {
  Block* b = current();
  unsigned me = b->fibers[b->cur].linear;
  b->xchg[me] = pred ? 1 : 0;
  yield(AT_WARP);
  unsigned base = me & ~31u, r = 0;
  for (unsigned l = 0; l < 32 && base + l < b->fibers.size(); l++)
    if (b->fibers[base + l].state != DONE && b->xchg[base + l]) r |= 1u << l;
  yield(AT_WARP);
  return r;
}

}  // namespace emu

#define threadIdx (emu::tidx())
#define blockIdx (emu::bidx())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

static inline void __syncthreads() { emu::yield(emu::AT_SYNC); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::yield(emu::AT_WARP); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::exchange(v, (unsigned)src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::exchange(v, emu::lane() ^ (unsigned)m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  unsigned l = emu::lane();
  return emu::exchange(v, l + d < 32 ? l + d : l);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  unsigned l = emu::lane();
  return emu::exchange(v, l >= d ? l - d : l);
}
static inline unsigned __ballot_sync(unsigned, int p) { return emu::ballot(p); }
static inline int __any_sync(unsigned, int p) { return emu::ballot(p) != 0; }
static inline int __all_sync(unsigned m, int p) { return emu::ballot(!p) == 0; }
static inline unsigned __activemask() { return 0xffffffffu; }

#define MAL_EMU_ATOMICS(T)                                                              \
  static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }              \
  static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }       \
  static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }       \
  static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
MAL_EMU_ATOMICS(float)
MAL_EMU_ATOMICS(double)
MAL_EMU_ATOMICS(int)
MAL_EMU_ATOMICS(unsigned)
MAL_EMU_ATOMICS(unsigned long long)
static inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p = o | v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }

template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline float __fmul_rn(float a, float b) { return a * b; }      // TU is built with -ffp-contract=off
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
#define __expf(a) expf(a)
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float2int_rd(float f) { return (int)floorf(f); }
static inline int __float2int_rz(float f) { return (int)f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
using std::max;
using std::min;
