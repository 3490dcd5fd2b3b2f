/*
 * letkf_b200_math.h -- bit-defined real32 front-end arithmetic shared by the CUDA
 * kernels (device) and the CPU oracle (host).
 *
 * Why this exists: the reference builds yo / Yb in real32
 * (module_letkf_core.f90:323-325,431-452) and only then promotes to real64
 * (module_letkf_core.f90:645-647).  A 1-ulp difference in one real32 operation is a
 * 6e-8 relative perturbation -- 600x the 1e-10 FP64 parity bar.  Fortran leaves the
 * evaluation order of sum()/dot_product() and the value of exp() to the compiler, so
 * this header DEFINES them once:
 *   - every real32 +,-,*,/ and sqrt is a single IEEE-754 round-to-nearest operation,
 *     never contracted into an FMA (device: __f*_rn intrinsics; host: build with
 *     -ffp-contract=off);
 *   - sums are sequential, left to right (member 0 .. k-1);
 *   - exp() of a real32 argument is computed by lk_expf() below: a fixed sequence of
 *     IEEE double FMAs, identical bit for bit on host and device, rounded once to real32.
 *
 * Reference constants restated here: gc1999 = 2*sqrt(10/3) (module_param.f90:116),
 * search radius r2 = gc1999*gc1999 (module_localization.f90:202), Gaspari-Cohn
 * coefficients (module_localization.f90:339-351).
 */
#ifndef LETKF_B200_MATH_H
#define LETKF_B200_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LK_HD __host__ __device__ __forceinline__
#else
#define LK_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define LK_MUL(a, b) __fmul_rn((a), (b))
#define LK_ADD(a, b) __fadd_rn((a), (b))
#define LK_SUB(a, b) __fsub_rn((a), (b))
#define LK_DIV(a, b) __fdiv_rn((a), (b))
#define LK_SQRT(a) __fsqrt_rn((a))
#define LK_FMA64(a, b, c) __fma_rn((a), (b), (c))
#else
/* host: translation units including this header are compiled with -ffp-contract=off */
#define LK_MUL(a, b) ((float)((float)(a) * (float)(b)))
#define LK_ADD(a, b) ((float)((float)(a) + (float)(b)))
#define LK_SUB(a, b) ((float)((float)(a) - (float)(b)))
#define LK_DIV(a, b) ((float)((float)(a) / (float)(b)))
#define LK_SQRT(a) (sqrtf((float)(a)))
#define LK_FMA64(a, b, c) (fma((double)(a), (double)(b), (double)(c)))
#endif

/* real32 parameter expressions, evaluated as the Fortran compiler must (IEEE real32). */
LK_HD float lk_sqrt_10_3(void) { return LK_SQRT(LK_DIV(10.0f, 3.0f)); }
LK_HD float lk_gc1999(void) { return LK_MUL(2.0f, lk_sqrt_10_3()); }
/* squared, normalised search radius (module_localization.f90:202) */
LK_HD float lk_search_r2(void) {
  float g = lk_gc1999();
  return LK_MUL(g, g);
}

/*
 * exp(x) for a real32 x, result real32.  Algorithm: n = rint(x*log2(e));
 * r = x - n*ln2 (two-term Cody-Waite in double, |r| <= 0.3466); e^r by the degree-13
 * Taylor polynomial in Horner form (truncation < 5e-18), every step one double FMA;
 * scale by 2^n built from the exponent bits; round once to real32.  Valid for
 * |x| < 700; the hot path only uses x = 0.25*r2 in [0, 3.34].
 */
LK_HD float lk_expf(float xf) {
  const double x = (double)xf;
  const double LOG2E = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01;
  const double LN2_LO = 1.90821492927058770002e-10;
  const double n = rint(x * LOG2E);
  double r = LK_FMA64(-n, LN2_HI, x);
  r = LK_FMA64(-n, LN2_LO, r);
  double p = 1.0 / 6227020800.0; /* 1/13! */
  p = LK_FMA64(p, r, 1.0 / 479001600.0);
  p = LK_FMA64(p, r, 1.0 / 39916800.0);
  p = LK_FMA64(p, r, 1.0 / 3628800.0);
  p = LK_FMA64(p, r, 1.0 / 362880.0);
  p = LK_FMA64(p, r, 1.0 / 40320.0);
  p = LK_FMA64(p, r, 1.0 / 5040.0);
  p = LK_FMA64(p, r, 1.0 / 720.0);
  p = LK_FMA64(p, r, 1.0 / 120.0);
  p = LK_FMA64(p, r, 1.0 / 24.0);
  p = LK_FMA64(p, r, 1.0 / 6.0);
  p = LK_FMA64(p, r, 0.5);
  p = LK_FMA64(p, r, 1.0);
  p = LK_FMA64(p, r, 1.0);
  const int64_t e = (int64_t)n + 1023;
  uint64_t bits = (uint64_t)e << 52;
  double scale;
#if defined(__CUDA_ARCH__)
  scale = __longlong_as_double((long long)bits);
#else
  memcpy(&scale, &bits, sizeof(double));
#endif
  return (float)(p * scale);
}

/*
 * Gaspari & Cohn (1999) 5th-order piecewise rational, real32 Horner forms exactly as
 * module_localization.f90:333-364 (x already normalised by the length scale).
 * NOTE: in real32 the second branch is slightly negative for
 * z >~ 1.958; the caller takes sqrt() of it and the reference propagates the NaN.
 */
LK_HD float lk_gaspari_cohn(float x) {
  const float a = lk_sqrt_10_3();
  const float a1 = -0.25f, a2 = 0.5f, a3 = 0.625f;
  const float a4 = -LK_DIV(5.0f, 3.0f);
  const float a5 = 1.0f;
  const float b1 = LK_DIV(1.0f, 12.0f);
  const float b2 = -0.5f, b3 = 0.625f;
  const float b4 = LK_DIV(5.0f, 3.0f);
  const float b5 = -5.0f, b6 = 4.0f;
  const float b7 = -LK_DIV(2.0f, 3.0f);
  const float z = LK_DIV(x, a);
  if (z <= 1.0f) {
    float t = LK_ADD(LK_MUL(a1, z), a2);
    t = LK_ADD(LK_MUL(z, t), a3);
    t = LK_ADD(LK_MUL(z, t), a4);
    return LK_ADD(LK_MUL(LK_MUL(z, z), t), a5);
  } else if (z <= 2.0f) {
    float t = LK_ADD(LK_MUL(b1, z), b2);
    t = LK_ADD(LK_MUL(z, t), b3);
    t = LK_ADD(LK_MUL(z, t), b4);
    t = LK_ADD(LK_MUL(z, t), b5);
    return LK_ADD(LK_ADD(LK_MUL(z, t), b6), LK_DIV(b7, z));
  }
  return 0.0f;
}

/*
 * Localisation applied to the observation ERROR (module_letkf_core.f90:439-450,516-523):
 *   weight_function /= 1 : error_inv = 1 / (err * exp(0.25*r2))
 *   weight_function == 1 : error_inv = sqrt(GC(sqrt(r2))) / err
 */
LK_HD float lk_error_inv(float err, float r2, int weight_function) {
  if (weight_function != 1) {
    return LK_DIV(1.0f, LK_MUL(err, lk_expf(LK_MUL(0.25f, r2))));
  }
  return LK_DIV(LK_SQRT(lk_gaspari_cohn(LK_SQRT(r2))), err);
}

/* 1/(hclr*1e3) and friends (module_localization.f90:76-82,234-240) */
LK_HD float lk_clr_inv(float clr_km) { return LK_DIV(1.0f, LK_MUL(clr_km, 1e3f)); }

/* distance of a coordinate from an interval (module_kdtree2.f90:1460-1477) */
LK_HD float lk_dis2_from_bnd(float x, float amin, float amax) {
  if (x > amax) {
    float d = LK_SUB(x, amax);
    return LK_MUL(d, d);
  }
  if (x < amin) {
    float d = LK_SUB(amin, x);
    return LK_MUL(d, d);
  }
  return 0.0f;
}

#endif /* LETKF_B200_MATH_H */
