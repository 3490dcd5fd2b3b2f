// "Matrix-function" form of letkf_solve (module_letkf_core.f90:649-679) shared by the k = 32 warp
// kernel (kernels_fcn32.cu) and the CTA-per-unit kernel (kernels_fcn.cu).
//
// The reference diagonalises C = (k-1)/rho I + Yb^T Yb with ?syevd and forms Pa~ = C^-1 and
// Wa = sqrt(k-1) C^(-1/2) with two k^3 gemms (module_eigen.f90:37-108).  Every grid point then only
// APPLIES them to a handful of k-vectors:  xa = xmean + xb'.(C^-1 b) + sqrt(k-1) C^(-1/2) xb'.
// Neither eigenvectors nor eigenvalues are needed for that, only the matrix function of C acting on
// vectors.  So here:
//   1. C = Q T Q^T, T tridiagonal, by Householder reflections (4/3 k^3 flops -- the one O(k^3) step);
//   2. T^(-1/2) z ~ sqrt(a) sum_j c_j (T + a beta_j I)^-1 z with the N = 32 poles of the
//      trapezoid rule in the conformal variable of Hale, Higham & Trefethen, "Computing A^alpha, log(A)
//      and related matrix functions by contour integrals", SIAM J. Numer. Anal. 46 (2008), method 3
//      (identical to Zolotarev's best rational approximation of x^(-1/2) on [a, a 2^q]); relative
//      error < 1e-15 for condition numbers up to 2^20, < 5e-13 up to 2^27 (checked on the host when the
//      table is built, fcn_poles.cu).  Each pole is a shifted SPD tridiagonal solve (LDL^T, O(k)),
//      all 32 poles run in the 32 lanes of a warp;
//   3. x' . C^-1 b = (T^(-1/2) Q^T x') . (T^(-1/2) Q^T b): the mean update needs no second function.
// The spectrum bound: every eigenvalue of C is >= mu = (k-1)/rho by construction, and <= the
// Gershgorin bound of T, so the interval [a, a 2^q] is known per unit without an eigensolve.
// The result is the same function of C the reference evaluates (basis invariant), to ~1e-14.
#pragma once

#include "letkf_internal.cuh"

namespace lk {

constexpr int FCN_NP = 32;    // poles = lanes
constexpr int FCN_QMAX = 40;  // intervals [1, 2^q], q = 1..QMAX
constexpr int FCN_QSAFE = 27; // up to here the 32-pole expansion is good to < 1e-12; beyond, the call fails loudly
// table layout: poles[(q * 2 + 0) * 32 + j] = c_j, poles[(q * 2 + 1) * 32 + j] = beta_j

// interval index for a spectrum inside [a, lmax]
__device__ __forceinline__ int fcn_interval(double lmax, double a) {
  const double kap = lmax / a;
  if (!(kap >= 1.0)) return 1;  // also NaN: the result is NaN anyway
  if (!(kap < 1.0995e12)) return FCN_QMAX;
  int ex;
  frexp(kap, &ex);  // kap = f 2^ex, 0.5 <= f < 1
  return ex < 1 ? 1 : (ex > FCN_QMAX ? FCN_QMAX : ex);
}
__device__ __forceinline__ double fcn_lower_edge(double mu) { return mu * (1.0 - 9.5367431640625e-7); }  // 1 - 2^-20

constexpr unsigned FULLF = 0xffffffffu;
constexpr int FCN_SEG = 16;  // check-point distance of the forward sweep

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLF, v, o);
  return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULLF, v, o));
  return v;
}

// sum v[0..16) over the warp; afterwards v[0] in lane l is the total of element l >> 1
__device__ __forceinline__ void treduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int n = 16, mask = 16; n > 1; n >>= 1, mask >>= 1) {
    const bool up = lane & mask;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const double send = up ? v[i] : v[i + n / 2];
      const double keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(FULLF, send, mask);
    }
  }
  v[0] += __shfl_xor_sync(FULLF, v[0], 1);
}

// 1/x for a pivot: MUFU seed + two Newton steps (~1 ulp; the IEEE divide is a ~25-instruction dependent
// chain, and there are k of them in sequence per pole).  Pivots are >= mu > 0 and far inside the real32
// range of the seed; NaN propagates.
__device__ __forceinline__ double fcn_rcp(double x) {
  float s;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"((float)x));
  double r = (double)s;
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// Reciprocal LDL^T pivots of T + beta I for this lane's pole: rp[i*32 + lane], i = 0..k-1 (one warp).
__device__ __forceinline__ void pole_pivots(int k, const double *d, const double *e, double beta, double *rp,
                                            int lane) {
  double rprev = fcn_rcp(d[0] + beta);
  rp[lane] = rprev;
  for (int i = 1; i < k; ++i) {
    const double l = e[i - 1] * rprev;
    rprev = fcn_rcp((d[i] + beta) - l * e[i - 1]);
    rp[i * 32 + lane] = rprev;
  }
}

// g = sum_j cw_j (T + beta_j)^-1 z for the vector in zb[0..k) (overwritten by g); one warp, lane = pole,
// cw = sqrt(a) c_lane.  rp: reciprocal pivots [i*32 + lane]; e: off-diagonal of T;
// ck: (k/FCN_SEG + 1) * 32 doubles of scratch.  The forward sweep L y = z keeps y only at the start of
// every 16-row segment; the backward sweep recomputes a segment into registers, so no k x 32 array is
// stored.  The 16 values of a segment are summed over the poles with one transposed butterfly.
__device__ __forceinline__ void pole_solve(double *zb, int k, const double *e, const double *rp, double cw, double *ck,
                                           int lane) {
  double yp = zb[0];
  ck[lane] = yp;
  for (int i = 1; i < k; ++i) {
    const double l = e[i - 1] * rp[(i - 1) * 32 + lane];
    yp = fma(-l, yp, zb[i]);
    if ((i & (FCN_SEG - 1)) == 0) ck[(i / FCN_SEG) * 32 + lane] = yp;
  }
  double unext = 0.0;
  for (int seg = (k - 1) / FCN_SEG; seg >= 0; --seg) {
    const int i0 = seg * FCN_SEG;
    const int len = k - i0 < FCN_SEG ? k - i0 : FCN_SEG;
    double ys[FCN_SEG];
    ys[0] = ck[seg * 32 + lane];
#pragma unroll
    for (int t = 1; t < FCN_SEG; ++t) {
      ys[t] = 0.0;
      if (t < len) {
        const int i = i0 + t;
        const double l = e[i - 1] * rp[(i - 1) * 32 + lane];
        ys[t] = fma(-l, ys[t - 1], zb[i]);
      }
    }
#pragma unroll
    for (int t = FCN_SEG - 1; t >= 0; --t) {
      if (t < len) {
        const int i = i0 + t;
        const double ee = i + 1 < k ? e[i] : 0.0;
        const double u = fma(-ee, unext, ys[t]) * rp[i * 32 + lane];
        unext = u;
        ys[t] = cw * u;
      }
    }
    treduce16(ys, lane);
    __syncwarp();  // every lane has read zb[i0 .. i0+len) of this segment
    if (!(lane & 1) && (lane >> 1) < len) zb[i0 + (lane >> 1)] = ys[0];
  }
  __syncwarp();
}

}  // namespace lk
