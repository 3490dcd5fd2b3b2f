// k = 32 fast paths of the Gram and transform stages (FP64 build): one warp per analysis unit.
//
// gram32: C = Yb Yb^T + mu I and b = Yb yo on the FP64 tensor pipe.  C is symmetric, so only the
// 10 lower 8x8 tiles of the 4x4 tile grid are accumulated: per 4 observation rows one lane loads
// 4 perturbations (member 8I + lane/4 of row r + lane%4, I = 0..3), forms yb = pert*error_inv in
// real32 (single rounding, module_letkf_core.f90:452/525), promotes, and the SAME register serves
// as the A fragment of tile row I and the B fragment of tile column I of
// mma.sync.m8n8k4.f64 -- 10 DMMA per 4 rows instead of 32 DFMA per row and lane.
//
// transform32: xa = xb_mean + xb'.wbar + sqrt(k-1) U Lambda^(-1/2) U^T xb' with lane i holding row i
// of U (the layout eig32 writes), one transposed butterfly for U^T xb' and a shared-memory
// broadcast for the second product; RTPP/RTPS exactly as kernels_xform.cu.
#include <cstdlib>

#include "gram32.cuh"
#include "xform32.cuh"

namespace lk {

constexpr unsigned FULLM = 0xffffffffu;

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
    gram32_dmma_kernel(TreeViews tv, int64_t nunits, const int32_t *__restrict__ unit_pt, double mu,
                       double *__restrict__ C, double *__restrict__ bvec, int32_t *__restrict__ nanflag) {
  const int lane = threadIdx.x & 31;
  const int64_t unit = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (unit >= nunits || (tv.nunits_dev && unit >= *tv.nunits_dev)) return;
  const bool wn = gram32_unit(tv, unit_pt[unit], mu, C + unit * 1024, 32, bvec + unit * 32, lane);
  if (lane == 0) nanflag[unit] = wn ? 1 : 0;
}

void launch_gram32(cudaStream_t s, const TreeViews &tv, int64_t nunits, const int32_t *unit_pt, double mu,
                   double *C, double *b, int32_t *nanflag) {
  if (nunits == 0) return;
  // resident CTAs per SM the kernel is compiled for: 6 = 80 registers (no cap), 7 = 72, 8 = 64 (spills).
  // Measured on a 225 x 225 x 50 grid: 161.6 / 155.4 / 160.8 ms -- 28 warps per SM at 72 registers is the best.
  static const int minb = [] {
    const char *e = getenv("LETKF_B200_GRAM32_MINB");
    return e ? atoi(e) : 7;
  }();
  const unsigned grid = (unsigned)((nunits + 3) / 4);
  if (minb == 8)
    gram32_dmma_kernel<8><<<grid, 128, 0, s>>>(tv, nunits, unit_pt, mu, C, b, nanflag);
  else if (minb == 7)
    gram32_dmma_kernel<7><<<grid, 128, 0, s>>>(tv, nunits, unit_pt, mu, C, b, nanflag);
  else
    gram32_dmma_kernel<6><<<grid, 128, 0, s>>>(tv, nunits, unit_pt, mu, C, b, nanflag);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

// ---- transform, k = 32 ------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
    transform32_kernel(int64_t nunits, const T *__restrict__ U, const T *__restrict__ lam,
                       const T *__restrict__ wbar, Xform32Args xa) {
  __shared__ __align__(16) T sbuf[4][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t unit = (int64_t)blockIdx.x * 4 + w;
  if (unit >= nunits) return;
  const int64_t pt = xa.pt_base + xa.unit_pt[unit];
  T u[32];
  {
    const T *Uu = U + unit * 1024 + (int64_t)lane * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) u[j] = Uu[j];
  }
  const T scale = sqrt((T)31) / sqrt(lam[unit * 32 + lane]);  // lane j: sqrt(k-1)/sqrt(lambda_j)
  transform32_unit<T>(u, scale, wbar[unit * 32 + lane], xa.nanflag[unit] != 0, pt, xa, sbuf[w], lane);
}

template <typename T>
void launch_transform32(cudaStream_t s, int64_t nunits, const int32_t *unit_pt, int64_t npts_total, int64_t pt_base,
                        const T *U, const T *lam, const T *wbar, const int32_t *nanflag, int nfields, float *var,
                        int use_rtpp, float rtpp_alpha, int use_rtps, float rtps_alpha, double *xa_raw) {
  if (nunits == 0 || nfields == 0) return;
  Xform32Args xa{unit_pt, nanflag, npts_total, pt_base, nfields, var, use_rtpp, rtpp_alpha, use_rtps, rtps_alpha, xa_raw};
  transform32_kernel<T><<<(unsigned)((nunits + 3) / 4), 128, 0, s>>>(nunits, U, lam, wbar, xa);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_transform32<double>(cudaStream_t, int64_t, const int32_t *, int64_t, int64_t, const double *,
                                         const double *, const double *, const int32_t *, int, float *, int, float,
                                         int, float, double *);
template void launch_transform32<float>(cudaStream_t, int64_t, const int32_t *, int64_t, int64_t, const float *,
                                        const float *, const float *, const int32_t *, int, float *, int, float, int,
                                        float, double *);

// weights dump for the row-major U of the k = 32 path: Wa = sqrt(31) U Lambda^(-1/2) U^T
template <typename T>
__global__ void __launch_bounds__(256)
    weights_dump32_kernel(int64_t nunits, const int32_t *__restrict__ unit_pt, const T *__restrict__ U,
                          const T *__restrict__ lam, const T *__restrict__ wbar, double *__restrict__ wbar_out,
                          double *__restrict__ Wa_out) {
  __shared__ T sc[32];
  const int64_t unit = blockIdx.x;
  if (unit >= nunits) return;
  const int64_t pt = unit_pt[unit];
  const T *Uu = U + unit * 1024;  // U[i][l] at i*32 + l
  if (threadIdx.x < 32) sc[threadIdx.x] = sqrt((T)31) / sqrt(lam[unit * 32 + threadIdx.x]);
  __syncthreads();
  if (wbar_out && threadIdx.x < 32) wbar_out[pt * 32 + threadIdx.x] = (double)wbar[unit * 32 + threadIdx.x];
  if (Wa_out)
    for (int e = threadIdx.x; e < 1024; e += blockDim.x) {
      const int i = e & 31, j = e >> 5;
      T a = 0;
      for (int l = 0; l < 32; ++l) a += Uu[i * 32 + l] * sc[l] * Uu[j * 32 + l];
      Wa_out[pt * 1024 + e] = (double)a;
    }
}
template <typename T>
void launch_weights_dump32(cudaStream_t s, int64_t nunits, const int32_t *unit_pt, const T *U, const T *lam,
                           const T *wbar, double *wbar_out, double *Wa_out) {
  if (nunits == 0) return;
  weights_dump32_kernel<T><<<(unsigned)nunits, 256, 0, s>>>(nunits, unit_pt, U, lam, wbar, wbar_out, Wa_out);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_weights_dump32<double>(cudaStream_t, int64_t, const int32_t *, const double *, const double *,
                                            const double *, double *, double *);
template void launch_weights_dump32<float>(cudaStream_t, int64_t, const int32_t *, const float *, const float *,
                                           const float *, double *, double *);

}  // namespace lk
