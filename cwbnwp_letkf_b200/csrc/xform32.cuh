// k = 32 transform of one analysis unit, shared by transform32_kernel and by the fused epilogue of the
// warm-started eigensolver: lane i holds row i of U in u[], `scale` = sqrt(k-1)/sqrt(lambda_lane),
// `wb` = wbar_lane.  Follows module_letkf_core.f90:671-698 (see kernels_xform.cu for the derivation).
#pragma once

#include "letkf_internal.cuh"

namespace lk {

struct Xform32Args {
  const int32_t *unit_pt;   // chunk-relative point of each unit
  const int32_t *nanflag;
  int64_t npts_total, pt_base;
  int nfields;
  float *var;
  int use_rtpp;
  float rtpp_alpha;
  int use_rtps;
  float rtps_alpha;
  double *xa_raw;
};

template <typename T, int N>
__device__ __forceinline__ void treduce32(T (&v)[N], int lane) {
#pragma unroll
  for (int n = N, mask = 16; n > 1; n >>= 1, mask >>= 1) {
    const bool up = lane & mask;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const T send = up ? v[i] : v[i + n / 2];
      const T keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
  }
}

template <typename T>
__device__ __forceinline__ T wsum32(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sequential (member 0..31) real32 sum of one value per lane, as the oracle defines sum()
__device__ __forceinline__ float seq_sum32(float v) {
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < 32; ++m) s = LK_ADD(s, __shfl_sync(0xffffffffu, v, m));
  return s;
}

// buf: 32 T of shared memory private to the warp
template <typename T>
__device__ __forceinline__ void transform32_unit(const T (&u)[32], T scale, T wb, bool isnan_unit, int64_t pt,
                                                 const Xform32Args &xa_, T *buf, int lane) {
  const float ninv = LK_DIV(1.0f, 32.0f);
  for (int f = 0; f < xa_.nfields; ++f) {
    float *v = xa_.var + (int64_t)f * xa_.npts_total * 32;
    const float xb = v[(int64_t)lane * xa_.npts_total + pt];             // core:228
    const T xmean = (T)LK_MUL(seq_sum32(xb), ninv);                       // core:671 (real32)
    const T xp = (T)xb - xmean;                                           // core:672
    const T sdot = wsum32(xp * wb);
    T pr[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) pr[j] = u[j] * xp;
    treduce32<T, 32>(pr, lane);                                           // lane j: (U^T xb')_j
    __syncwarp();
    buf[lane] = pr[0] * scale;
    __syncwarp();
    T y = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) y = fma(u[j], buf[j], y);
    T xa = xmean + (sdot + y);                                            // core:673-675
    if (isnan_unit) xa = xa * T(NAN);
    if (xa_.xa_raw) xa_.xa_raw[pt * 32 + lane] = (double)xa;
    float xa32 = (float)xa;                                               // core:679
    if (xa_.use_rtpp || xa_.use_rtps) {                                   // core:684-698
      const float xa_mean = LK_MUL(seq_sum32(xa32), ninv);
      float xap = LK_SUB(xa32, xa_mean);
      if (xa_.use_rtpp) {
        const float t1 = LK_MUL(LK_SUB(1.0f, xa_.rtpp_alpha), xap);
        xap = (float)((T)t1 + (T)xa_.rtpp_alpha * xp);
      }
      if (xa_.use_rtps) {
        // dot_product(xb',xb') in working precision, sequential like the oracle
        T d = 0;
#pragma unroll
        for (int m = 0; m < 32; ++m) {
          const T x = __shfl_sync(0xffffffffu, xp, m);
          d += x * x;
        }
        const float xb_std = (float)d;
        const float xa_std = seq_sum32(LK_MUL(xap, xap));
        const float fac =
            LK_ADD(LK_SUB(LK_MUL(xa_.rtps_alpha, LK_SQRT(LK_DIV(xb_std, xa_std))), xa_.rtps_alpha), 1.0f);
        xap = LK_MUL(xap, fac);
      }
      xa32 = LK_ADD(xa_mean, xap);                                        // core:697
    }
    v[(int64_t)lane * xa_.npts_total + pt] = xa32;                        // core:229
    __syncwarp();
  }
}

}  // namespace lk
