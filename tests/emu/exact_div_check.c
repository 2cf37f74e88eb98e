/* Exhaustive check of the 3-instruction division by a small constant used in mal_common.cuh
 * (xdivc<C>): q = RN(x * r), rem = fma(-C, q, x), result = fma(rem, r, q) with r = RN(1 / C),
 * against IEEE x / C for EVERY finite float x.  Test infrastructure (tests/test_exact_arithmetic.py). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float divc(float x, float C) {
  const float r = 1.0f / C;
  float q = x * r;
  float rem = fmaf(-C, q, x);
  return fmaf(rem, r, q);
}

int main(int argc, char** argv) {
  /* argv[1]: stride over the 2^32 bit patterns (1 = exhaustive; a prime stride samples every exponent) */
  const long long stride = argc > 1 ? atoll(argv[1]) : 1;
  /* 9, 3: avg_pool2d windows and channel means (xdivc); the rest: Project3D's image-size divisors
     (size_div_verified in mal_math.cuh - keep the two lists in step) */
  const float consts[] = {9.0f, 3.0f, 47.0f, 95.0f, 127.0f, 159.0f, 191.0f, 255.0f, 383.0f, 511.0f, 639.0f, 1023.0f};
  const int nconst = (int)(sizeof(consts) / sizeof(consts[0]));
  unsigned long long bad_total = 0;
  for (int ci = 0; ci < nconst; ci++) {
    const float C = consts[ci];
    unsigned long long bad = 0, bad_sign = 0;
#pragma omp parallel for reduction(+ : bad, bad_sign) schedule(static)
    for (long long bits = 0; bits < (1LL << 32); bits += stride) {
      uint32_t u = (uint32_t)bits;
      float x;
      memcpy(&x, &u, 4);
      if (!isfinite(x)) continue;
      float want = x / C, got = divc(x, C);
      uint32_t a, b;
      memcpy(&a, &want, 4);
      memcpy(&b, &got, 4);
      if (a != b) {
        if (want == 0.0f && got == 0.0f) bad_sign++;   /* -0 vs +0 only */
        else bad++;
      }
    }
    printf("C=%g mismatches=%llu signed_zero_only=%llu\n", C, bad, bad_sign);
    bad_total += bad;
  }
  return bad_total ? 1 : 0;
}
