// Scalar helpers shared by the eigensolver kernels: MUFU-seeded reciprocal / reciprocal square root
// with Newton refinement (the IEEE double sqrt and divide are ~30-instruction dependent chains, far
// too slow inside a Jacobi step), tolerances, and the rotation generator.
#pragma once

#include "letkf_internal.cuh"

namespace lk {

__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <typename T>
struct Fast;
template <>
struct Fast<double> {
  static __device__ __forceinline__ double rsqrt(double x) {  // x > 0 within real32 range; ~1 ulp
    double r = (double)rsqrt_approx((float)x);
    const double h = 0.5 * x;
    double e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-h * r, r, 0.5);
    r = fma(r, e, r);
    return r;
  }
  static __device__ __forceinline__ double rcp(double x) {
    double r = (double)rcp_approx((float)x);
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
  }
  // one Newton step (relative error ~1e-13): enough for a rotation ANGLE
  static __device__ __forceinline__ double rsqrt1(double x) {
    double r = (double)rsqrt_approx((float)x);
    const double e = fma(-0.5 * x * r, r, 0.5);
    return fma(r, e, r);
  }
  static __device__ __forceinline__ double rcp1(double x) {
    double r = (double)rcp_approx((float)x);
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
  }
  static __device__ __forceinline__ double tol2(int k) { return 4.930380657631324e-32 * k; }  // (eps sqrt k)^2
};
template <>
struct Fast<float> {
  static __device__ __forceinline__ float rsqrt(float x) {
    float r = rsqrt_approx(x);
    const float h = 0.5f * x;
    const float e = fmaf(-h * r, r, 0.5f);
    return fmaf(r, e, r);
  }
  static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
  static __device__ __forceinline__ float rsqrt1(float x) { return rsqrt_approx(x); }
  static __device__ __forceinline__ float rcp1(float x) { return __frcp_rn(x); }
  static __device__ __forceinline__ float tol2(int k) { return 1.4210855e-14f * k; }
};

// Jacobi rotation for a column pair with squared norms alpha, beta and inner product gamma:
// t = tan(theta) of the rotation that orthogonalises the pair, c = cos, s = sin.  With `rot` false the
// identity is returned.  FASTROT evaluates the ANGLE in real32 (MUFU rsqrt/rcp); c = rsqrt(1+t^2),
// s = c t stay in working precision so the rotation is orthogonal to working accuracy and only the
// annihilation of gamma is approximate (residual cosine ~1e-7 of the old one).  FASTROT needs squared
// norms inside the real32 range.
template <typename T, bool FASTROT>
__device__ __forceinline__ void jacobi_rotation(T alpha, T beta, T gamma, bool rot, T &c, T &s, T &t) {
  const T delta = beta - alpha;
  T tt;
  if (FASTROT) {
    const float df = (float)delta, gf = (float)gamma;
    const float x = fmaxf(fmaf(df, df, 4.f * gf * gf), 1e-30f);
    const float h = x * rsqrt_approx(x);
    tt = (T)((df >= 0.f ? 2.f : -2.f) * gf * rcp_approx(fabsf(df) + h));
  } else {
    T x = fma(delta, delta, T(4) * gamma * gamma);
    x = rot ? x : T(1);
    const T h = x * Fast<T>::rsqrt1(x);
    tt = (delta >= T(0) ? T(2) : T(-2)) * gamma * Fast<T>::rcp1(fabs(delta) + h);
  }
  const T cc = Fast<T>::rsqrt(fma(tt, tt, T(1)));
  c = T(1);
  s = T(0);
  t = T(0);
  if (rot) {
    t = tt;
    c = cc;
    s = cc * tt;
  }
}

}  // namespace lk
