// Checks on the GPU that the packed helpers of mal_common.cuh equal their scalar counterparts bit for bit.
#include <cstdio>
#include <cstdlib>
#include "../../mal_b200/csrc/mal_math.cuh"
using namespace mal;
__global__ void k(const float* a, const float* b, const float* c, int n, int* bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a0 = a[i], a1 = a[(i + 1) % n], b0 = b[i], b1 = b[(i + 7) % n], c0 = c[i], c1 = c[(i + 3) % n];
  pk2 A = pack2(a0, a1), B = pack2(b0, b1), C = pack2(c0, c1);
  float r0, r1;
  auto chk = [&](int id, pk2 r, float s0, float s1) {
    unpack2(r, r0, r1);
    if (__float_as_int(r0) != __float_as_int(s0) || __float_as_int(r1) != __float_as_int(s1)) atomicAdd(bad + id, 1);
  };
  chk(0, x2mul(A, B), xmul(a0, b0), xmul(a1, b1));
  chk(1, x2add(A, B), xadd(a0, b0), xadd(a1, b1));
  chk(2, x2sub(A, B), xsub(a0, b0), xsub(a1, b1));
  chk(3, x2fma(A, B, C), xfma(a0, b0, c0), xfma(a1, b1, c1));
  chk(4, x2divc<9>(A), xdivc<9>(a0), xdivc<9>(a1));
  chk(5, x2divc<3>(A), xdivc<3>(a0), xdivc<3>(a1));
  chk(6, x2add(C, abs2(x2sub(A, B))), xadd(c0, fabsf(xsub(a0, b0))), xadd(c1, fabsf(xsub(a1, b1))));
  chk(7, x2mul(A, dup2(b0)), xmul(a0, b0), xmul(a1, b0));
  chk(8, x2fma(A, dup2(b0), dup2(c0)), xfma(a0, b0, c0), xfma(a1, b0, c0));
  chk(9, x2add(x2mul_nf(dup2(0.85f), A), x2mul_nf(dup2(0.15f), B)), xadd(xmul(0.85f, a0), xmul(0.15f, b0)), xadd(xmul(0.85f, a1), xmul(0.15f, b1)));
  // window sums
  float w[9], v[9]; pk2 W[9], V[9];
  for (int j = 0; j < 9; j++) { w[j] = a[(i + j) % n]; v[j] = b[(i + 2 * j) % n]; W[j] = pack2(w[j], v[j]); V[j] = dup2(c[(i + j) % n]); }
  float cc[9]; for (int j = 0; j < 9; j++) cc[j] = c[(i + j) % n];
  chk(10, sum9(W), sum9(w), sum9(v));
  chk(11, sum9_prod(W, W), sum9_prod(w, w), sum9_prod(v, v));
  chk(12, sum9_prod(W, V), sum9_prod(w, cc), sum9_prod(v, cc));
  // ssim
  float mu_y = xdivc<9>(sum9(cc)), eyy = xdivc<9>(sum9_prod(cc, cc)), myy = xmul(mu_y, mu_y), sgy = xsub(eyy, myy);
  SsimTerms t0, t1;
  ssim_terms2(x2divc<9>(sum9(W)), mu_y, myy, sgy, x2divc<9>(sum9_prod(W, W)), x2divc<9>(sum9_prod(W, V)), t0, t1);
  SsimTerms s0 = ssim_terms(xdivc<9>(sum9(w)), mu_y, xdivc<9>(sum9_prod(w, w)), eyy, xdivc<9>(sum9_prod(w, cc)));
  SsimTerms s1 = ssim_terms(xdivc<9>(sum9(v)), mu_y, xdivc<9>(sum9_prod(v, v)), eyy, xdivc<9>(sum9_prod(v, cc)));
  chk(13, pack2(t0.v, t1.v), s0.v, s1.v);
  chk(14, pack2(t0.n, t1.n), s0.n, s1.n);
  chk(15, pack2(t0.d, t1.d), s0.d, s1.d);
}
int main() {
  const int n = 1 << 20;
  float *h = (float*)malloc(3 * n * 4), *d; int* bad;
  srand(1);
  for (int i = 0; i < 3 * n; i++) h[i] = (float)rand() / RAND_MAX;
  cudaMalloc(&d, 3 * n * 4); cudaMalloc(&bad, 64 * 4); cudaMemset(bad, 0, 64 * 4);
  cudaMemcpy(d, h, 3 * n * 4, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(d, d + n, d + 2 * n, n, bad);
  int hb[16]; cudaMemcpy(hb, bad, 64, cudaMemcpyDeviceToHost);
  int tot = 0;
  for (int i = 0; i < 16; i++) { printf("check %2d: %d mismatches\n", i, hb[i]); tot += hb[i]; }
  printf("%s\n", tot ? "FAIL" : "OK");
  return tot != 0;
}
