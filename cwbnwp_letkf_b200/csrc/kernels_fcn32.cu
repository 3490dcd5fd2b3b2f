// k = 32, FP64: letkf_solve without an eigendecomposition, one WARP per analysis unit (fcn_common.cuh).
//
// Lane i holds row i of the symmetric matrix C in 32 registers.  Householder step j eliminates column j
// below the sub-diagonal: the column is register a[j] of the lanes > j, so forming the reflector is one
// warp sum; p = tau A v reads v from a 256-byte shared-memory broadcast, and the rank-2 update
// a[l] -= v_i w_l + w_i v_l is lane-local with compile-time register indices (the 30 steps are fully
// unrolled: straight-line code, no register permutation).  The reflector overwrites the eliminated column
// (register a[j] of lanes > j), so Q is applied to vectors from registers with one warp sum per reflector.
// Then the 32 pole solves of T^(-1/2) z run one per lane (pole_solve), and the epilogue of letkf_solve
// (core:671-698) follows while everything is still on chip: per unit the kernel reads C (8 KB), b and the
// field column, and writes the field column.
#include <cstdlib>

#include "fcn_common.cuh"

namespace lk {

namespace {

constexpr int K32 = 32;

// z = Q^T y (FORWARD) or Q y for NV vectors at once (independent warp sums overlap)
template <int NV, bool FORWARD>
__device__ __forceinline__ void apply_q32(const double (&a)[K32], const double *tt, double (&y)[NV], int lane) {
#pragma unroll
  for (int s = 0; s < K32 - 2; ++s) {
    const int j = FORWARD ? s : K32 - 3 - s;
    const double vj = lane > j ? a[j] : 0.0;
    double dot[NV];
#pragma unroll
    for (int n = 0; n < NV; ++n) dot[n] = vj * y[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int n = 0; n < NV; ++n) dot[n] += __shfl_xor_sync(FULLF, dot[n], o);
    const double t = tt[j];
#pragma unroll
    for (int n = 0; n < NV; ++n) y[n] = fma(-(t * dot[n]), vj, y[n]);
  }
}

// Warp sums of three values at once, results in every lane: a transposed butterfly (the first two stages halve
// what travels) and three broadcasts -- 18 SHFL.32 and 6 DADD instead of 30 and 15, and ONE dependent chain.
__device__ __forceinline__ void wsum3(double &v0, double &v1, double &v2, int lane) {
  const bool u16 = lane & 16, u8 = lane & 8;
  double k0 = u16 ? v2 : v0, k1 = u16 ? 0.0 : v1;
  const double s0 = u16 ? v0 : v2, s1 = u16 ? v1 : 0.0;
  k0 += __shfl_xor_sync(FULLF, s0, 16);
  k1 += __shfl_xor_sync(FULLF, s1, 16);
  double k = u8 ? k1 : k0;
  k += __shfl_xor_sync(FULLF, u8 ? k0 : k1, 8);
  k += __shfl_xor_sync(FULLF, k, 4);
  k += __shfl_xor_sync(FULLF, k, 2);
  k += __shfl_xor_sync(FULLF, k, 1);
  v0 = __shfl_sync(FULLF, k, 0);
  v1 = __shfl_sync(FULLF, k, 8);
  v2 = __shfl_sync(FULLF, k, 16);
}

// sequential (member 0..31) real32 sum of one value per lane, as the oracle defines sum(): the values go
// through 128 bytes of shared memory (8 broadcast loads) instead of 32 dependent shuffles
__device__ __forceinline__ float seq_sum32f(float v, float *buf, int lane) {
  __syncwarp();
  buf[lane] = v;
  __syncwarp();
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < 32; m += 4) {
    const float4 t = *reinterpret_cast<const float4 *>(buf + m);
    s = LK_ADD(LK_ADD(LK_ADD(LK_ADD(s, t.x), t.y), t.z), t.w);
  }
  return s;
}

// per-warp shared memory (doubles)
constexpr int SM_VW = 0;                  // 64: (v_l, w_l) pairs
constexpr int SM_D = 64;                  // 32
constexpr int SM_E = 96;                  // 32
constexpr int SM_T = 128;                 // 32 tau
constexpr int SM_GB = 160;                // 32 g_b
constexpr int SM_Z = 192;                 // 32 work vector
constexpr int SM_CK = 224;                // 96 check points
constexpr int SM_F = 320;                 // 16 doubles = 32 floats: sequential-sum staging
constexpr int SM_X = 336;                 // 32 xb'
constexpr int SM_RP = 368;                // 1024 reciprocal pivots
constexpr int SM_WARP = 368 + 1024;       // 1392 doubles = 10.9 KB

template <int MINB>
__global__ void __launch_bounds__(128, MINB) fcn32_kernel(FcnArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t unit = (int64_t)blockIdx.x * 4 + w;
  if (unit >= A.nunits || (A.nunits_dev && unit >= *A.nunits_dev)) return;
  double *sm = reinterpret_cast<double *>(smem_raw) + (size_t)w * SM_WARP;
  double *vw = sm + SM_VW, *dd = sm + SM_D, *ee = sm + SM_E, *tt = sm + SM_T, *gb = sm + SM_GB, *zz = sm + SM_Z;
  double *ck = sm + SM_CK, *rp = sm + SM_RP, *xs = sm + SM_X;
  float *fb = reinterpret_cast<float *>(sm + SM_F);

  double a[K32];
  {
    const double *Cu = A.C + unit * (int64_t)(K32 * K32) + (int64_t)lane * K32;  // full symmetric (gram32)
#pragma unroll
    for (int j = 0; j < K32; j += 2) {
      const double2 t = *reinterpret_cast<const double2 *>(Cu + j);
      a[j] = t.x;
      a[j + 1] = t.y;
    }
  }

  const bool isnan_unit = A.nanflag[unit] != 0;
  const int64_t upt = A.unit_pt[unit];
  const double sk = sqrt(31.0);
  const float ninv = LK_DIV(1.0f, 32.0f);
  const bool have_field = A.var != nullptr && A.nz > 0 && A.nfields > 0;

  // Two vectors ride along with the tridiagonalisation and come out as Q^T y: b and the first field column.
  // Their reflector dot products share the step's own reduction (wsum3), so they cost no extra warp sums.
  double yb = A.bvec[unit * K32 + lane];
  float xb_first = 0.f;
  double xmean_first = 0.0, xp_first = 0.0;
  if (have_field) {
    xb_first = A.var[(int64_t)lane * A.npts_total + A.pt_base + upt];          // core:228
    xmean_first = (double)LK_MUL(seq_sum32f(xb_first, fb, lane), ninv);          // core:671 (real32)
    xp_first = (double)xb_first - xmean_first;                                   // core:672
  }
  double yx = xp_first;

  // ---- Householder tridiagonalisation, lane = row ----
#pragma unroll
  for (int j = 0; j < K32 - 2; ++j) {
    // (Lane j holds the same column in its own row by symmetry, but only to rounding: a reflector whose tau comes
    // from the row while v comes from the column is not orthogonal when the column is tiny -- measured 3e-7 in
    // Wa -- so the norm is reduced over the lanes that hold v.)
    const double xj = lane > j ? a[j] : 0.0;
    const double sigma = wsum(xj * xj);
    const double x1 = __shfl_sync(FULLF, a[j], j + 1);
    if (lane == j) dd[j] = a[j];
    const bool zero = !(sigma > 1e-290) && sigma == sigma;  // NaN falls through and propagates
    const double rn = rsqrt(zero ? 1.0 : sigma);
    const double nrm = zero ? 0.0 : sigma * rn;
    const double alpha = zero ? 0.0 : -copysign(nrm, x1);
    const double tau = zero ? 0.0 : rn * __drcp_rn(nrm + fabs(x1));
    const double v = lane == j + 1 ? x1 - alpha : (lane > j + 1 ? xj : 0.0);
    vw[2 * lane] = v;
    __syncwarp();
    double p0 = 0.0, p1 = 0.0;
#pragma unroll
    for (int l = j + 1; l < K32; ++l) {
      if ((l - j) & 1)
        p0 = fma(a[l], vw[2 * l], p0);
      else
        p1 = fma(a[l], vw[2 * l], p1);
    }
    double p = lane > j ? (p0 + p1) * tau : 0.0;
    double r0 = p * v, r1 = v * yb, r2 = v * yx;
    wsum3(r0, r1, r2, lane);
    const double K = 0.5 * tau * r0;
    yb = fma(-(tau * r1), v, yb);  // y <- H_j y
    yx = fma(-(tau * r2), v, yx);
    const double wv = fma(-K, v, p);
    vw[2 * lane + 1] = wv;
    __syncwarp();
#pragma unroll
    for (int l = j + 1; l < K32; ++l) {
      const double2 t = *reinterpret_cast<const double2 *>(vw + 2 * l);
      a[l] = fma(-v, t.y, fma(-wv, t.x, a[l]));
    }
    a[j] = v;  // reflector j lives in register a[j] of lanes > j
    if (lane == 0) {
      ee[j] = alpha;
      tt[j] = tau;
    }
    __syncwarp();
  }
  if (lane == 30) dd[30] = a[30];
  if (lane == 31) {
    dd[31] = a[31];
    ee[30] = a[30];
    ee[31] = 0.0;
  }
  __syncwarp();

  // ---- spectrum bound, poles, pivots ----
  double cw;
  {
    const double el = lane > 0 ? fabs(ee[lane - 1]) : 0.0, er = lane < 31 ? fabs(ee[lane]) : 0.0;
    const double g = wmax(dd[lane] + el + er);
    const double aedge = fcn_lower_edge(A.mu);
    const int q = fcn_interval(g, aedge);
    if (q > FCN_QSAFE && lane == 0 && A.qmax) atomicMax(A.qmax, q);
    cw = sqrt(aedge) * A.poles[(q * 2 + 0) * FCN_NP + lane];
    pole_pivots(K32, dd, ee, aedge * A.poles[(q * 2 + 1) * FCN_NP + lane], rp, lane);
  }

  // ---- g_b = T^(-1/2) Q^T b ----
  gb[lane] = yb;
  __syncwarp();
  pole_solve(gb, K32, ee, rp, cw, ck, lane);
  const double gbl = gb[lane];

  // ---- fields (every level and field that shares these weights) ----
  bool first = true;
  for (int lev = 0; lev < (have_field ? A.nz : 1); ++lev)
    for (int f = 0; f < (have_field ? A.nfields : 1); ++f) {
      const int64_t pt = A.pt_base + (int64_t)lev * A.level_stride + upt;
      float *v = have_field ? A.var + (int64_t)f * A.npts_total * K32 : nullptr;
      double xmean = xmean_first, xp = xp_first, y = yx;
      if (!first && have_field) {
        const float xb = v[(int64_t)lane * A.npts_total + pt];             // core:228
        xmean = (double)LK_MUL(seq_sum32f(xb, fb, lane), ninv);            // core:671 (real32)
        xp = (double)xb - xmean;                                           // core:672
        double yy[1] = {xp};
        apply_q32<1, true>(a, tt, yy, lane);
        y = yy[0];
      }
      first = false;
      if (!have_field) break;
      zz[lane] = y;
      __syncwarp();
      pole_solve(zz, K32, ee, rp, cw, ck, lane);
      double yy[1] = {zz[lane]};
      const double sdot = wsum(yy[0] * gbl);                               // xb' . wbar (core:673)
      apply_q32<1, false>(a, tt, yy, lane);
      double xa = xmean + (sdot + sk * yy[0]);                             // core:673-675
      if (isnan_unit) xa = xa * (double)NAN;
      if (A.xa_raw) A.xa_raw[pt * K32 + lane] = xa;
      float xa32 = (float)xa;                                              // core:679
      if (A.use_rtpp || A.use_rtps) {                                      // core:684-698
        const float xa_mean = LK_MUL(seq_sum32f(xa32, fb, lane), ninv);
        float xap = LK_SUB(xa32, xa_mean);
        if (A.use_rtpp) {
          const float t1 = LK_MUL(LK_SUB(1.0f, A.rtpp_alpha), xap);
          xap = (float)((double)t1 + (double)A.rtpp_alpha * xp);
        }
        if (A.use_rtps) {
          __syncwarp();
          xs[lane] = xp;
          __syncwarp();
          double dsum = 0;  // dot_product(xb', xb') in working precision, sequential like the oracle
#pragma unroll
          for (int m = 0; m < 32; m += 2) {
            const double2 t = *reinterpret_cast<const double2 *>(xs + m);
            dsum += t.x * t.x;
            dsum += t.y * t.y;
          }
          const float xb_std = (float)dsum;
          const float xa_std = seq_sum32f(LK_MUL(xap, xap), fb, lane);
          const float fac =
              LK_ADD(LK_SUB(LK_MUL(A.rtps_alpha, LK_SQRT(LK_DIV(xb_std, xa_std))), A.rtps_alpha), 1.0f);
          xap = LK_MUL(xap, fac);
        }
        xa32 = LK_ADD(xa_mean, xap);                                       // core:697
      }
      v[(int64_t)lane * A.npts_total + pt] = xa32;                         // core:229
      __syncwarp();
    }

  // ---- parity dump: wbar = Q T^(-1/2) g_b, Wa = sqrt(k-1) Q T^(-1/2) Q^T ----
  if (A.wbar_out) {
    zz[lane] = gbl;
    __syncwarp();
    pole_solve(zz, K32, ee, rp, cw, ck, lane);
    double y[1] = {zz[lane]};
    apply_q32<1, false>(a, tt, y, lane);
    A.wbar_out[upt * K32 + lane] = y[0];
    __syncwarp();
  }
  if (A.Wa_out)
    for (int m = 0; m < K32; ++m) {
      double y[1] = {lane == m ? 1.0 : 0.0};
      apply_q32<1, true>(a, tt, y, lane);
      zz[lane] = y[0];
      __syncwarp();
      pole_solve(zz, K32, ee, rp, cw, ck, lane);
      y[0] = zz[lane];
      apply_q32<1, false>(a, tt, y, lane);
      A.Wa_out[upt * (K32 * K32) + m * K32 + lane] = sk * y[0];
      __syncwarp();
    }
}

}  // namespace

void launch_fcn32_solve(cudaStream_t s, const FcnArgs &a) {
  if (a.nunits == 0) return;
  LK_REQUIRE(a.k == 32, "fcn32: k must be 32");
  const size_t smem = sizeof(double) * 4 * SM_WARP;
  static const int minb = [] {
    const char *e = getenv("LETKF_B200_FCN32_MINB");  // resident CTAs per SM the kernel is compiled for (tuning knob)
    return e ? atoi(e) : 3;
  }();
  const int64_t nblk = (a.nunits + 3) / 4;
  LK_REQUIRE(nblk < ((int64_t)1 << 31), "fcn32: too many units for one launch");
  auto launch = [&](auto kern) {
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)nblk, 128, smem, s>>>(a);
  };
  if (minb == 4)
    launch(fcn32_kernel<4>);
  else if (minb == 5)
    launch(fcn32_kernel<5>);
  else
    launch(fcn32_kernel<3>);
  LK_CUDA(cudaGetLastError());
}

}  // namespace lk
