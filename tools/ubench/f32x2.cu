// Micro-benchmark: issue rate of scalar vs packed (f32x2) FP32 instructions on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }

constexpr int CH = 8, IT = 64, REP = 8;
// mode 0: FFMA x CH chains; 1: FFMA2; 2: FADD; 3: FADD2; 4: FMUL2; 5: FFMA + IADD mix; 6: FFMA2 + 2 IADD mix; 7: FFMA2 + LDS
template <int MODE>
__global__ void bench(float* out, long long* cyc, float s, float t) {
  __shared__ float sm[1024];
  sm[threadIdx.x] = s * threadIdx.x;
  __syncthreads();
  float x[CH]; u64 y[CH]; int z[CH];
  for (int i = 0; i < CH; i++) { x[i] = s * (i + threadIdx.x); float2 v = make_float2(x[i], x[i] + 1.f); y[i] = *(u64*)&v; z[i] = i + threadIdx.x; }
  float2 tv = make_float2(t, t * 1.5f); u64 t2 = *(u64*)&tv;
  float2 sv = make_float2(s, s * 0.5f); u64 s2 = *(u64*)&sv;
  long long c0 = clock64();
#pragma unroll 1
  for (int it = 0; it < IT; it++) {
#pragma unroll
    for (int r = 0; r < REP; r++)
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (MODE == 0) x[i] = fma1(x[i], t, s);
      if (MODE == 1) y[i] = fma2(y[i], t2, s2);
      if (MODE == 2) x[i] = add1(x[i], t);
      if (MODE == 3) y[i] = add2(y[i], t2);
      if (MODE == 4) y[i] = mul2(y[i], t2);
      if (MODE == 5) { x[i] = fma1(x[i], t, s); asm volatile("add.s32 %0, %0, %1;" : "+r"(z[i]) : "r"(it)); }
      if (MODE == 6) { y[i] = fma2(y[i], t2, s2); asm volatile("add.s32 %0, %0, %1;" : "+r"(z[i]) : "r"(it)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(it)); }
      if (MODE == 7) { y[i] = fma2(y[i], t2, s2); float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((threadIdx.x + i * 32 + it) & 1023)))); x[i] = v; }
      if (MODE == 8) { y[i] = fma2(y[i], t2, s2); x[i] = fma1(x[i], t, s); }
    }
  }
  long long c1 = clock64();
  float acc = 0; for (int i = 0; i < CH; i++) { float2 v = *(float2*)&y[i]; acc += x[i] + v.x + v.y + z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
template <int MODE> void run(const char* name, int instr_per_iter, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  bench<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f); cudaDeviceSynchronize();
  bench<MODE><<<148, threads>>>(out, cyc, 1.0001f, 0.9999f); cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
  double warps_per_smsp = threads / 32 / 4.0;
  double winstr = (double)IT * REP * CH * instr_per_iter * warps_per_smsp;   // warp instructions per SMSP
  printf("%-28s threads=%4d  cycles=%9.0f  warp-instr/clk/SMSP=%.3f  clk per (listed) group per SMSP=%.3f\n", name, threads, c, winstr / c, c / ((double)IT * REP * CH * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {256, 1024}) {
    run<0>("FFMA", 1, th); run<1>("FFMA2", 1, th); run<2>("FADD", 1, th); run<3>("FADD2", 1, th); run<4>("FMUL2", 1, th);
    run<5>("FFMA+IADD", 2, th); run<6>("FFMA2+IADD+XOR", 3, th); run<7>("FFMA2+LDS", 2, th); run<8>("FFMA2+FFMA", 2, th);
  }
  return 0;
}
