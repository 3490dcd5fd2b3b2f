// Weight application: the tail of letkf_solve (module_letkf_core.f90:662-698).
//
// The reference materialises W = wbar 1^T + sqrt(k-1) Pa~^(1/2) with two k^3 gemms and applies
// it with ?gemv('t').  Each grid point transforms single k-vectors, so here the eigenpairs are
// applied directly:
//   xa(m) = xb_mean + xb' . wbar + sqrt(k-1) [ U Lambda^(-1/2) U^T xb' ](m)
// which is O(k^2) per field and equals the reference's result up to working-precision rounding.
// RTPP / RTPS then follow the reference's mixed real32/real64 expressions literally
// (module_letkf_core.f90:684-698) with sequential real32 sums.
//
// Generic path: one CTA per analysis unit, looping over the fields that share the weights.
#include "letkf_internal.cuh"

namespace lk {

template <typename T>
__device__ __forceinline__ T warp_sum_x(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
    transform_kernel(int k, int64_t nunits, const int32_t *__restrict__ unit_pt, int64_t npts_total,
                     int64_t pt_base, const T *__restrict__ U, const T *__restrict__ lam,
                     const T *__restrict__ wbar, const int32_t *__restrict__ nanflag, int nfields,
                     float *__restrict__ var, int use_rtpp, float rtpp_alpha, int use_rtps,
                     float rtps_alpha, double *__restrict__ xa_raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *xp = reinterpret_cast<T *>(smem_raw);  // [k] xb'
  T *zs = xp + k;                           // [k]
  T *red = zs + k;                          // [8]
  float *xb32 = reinterpret_cast<float *>(red + 8);  // [k]
  float *xa32 = xb32 + k;                            // [k]
  __shared__ float s_f[4];
  __shared__ T s_mean;

  const int64_t unit = blockIdx.x;
  if (unit >= nunits) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int64_t pt = pt_base + unit_pt[unit];
  const T *Uu = U + unit * (int64_t)k * k;
  const T *lu = lam + unit * (int64_t)k;
  const T *wb = wbar + unit * (int64_t)k;
  const float ninv = LK_DIV(1.0f, (float)k);
  const bool isnan_unit = nanflag[unit] != 0;
  const T sk = sqrt((T)(k - 1));  // sqrt(dble(nmember-1)) / sqrt(float(nmember-1)) (core:666/668)

  for (int f = 0; f < nfields; ++f) {
    float *v = var + (int64_t)f * npts_total * k;
    __syncthreads();
    for (int m = tid; m < k; m += nt) xb32[m] = v[(int64_t)m * npts_total + pt];  // core:228
    __syncthreads();
    if (tid == 0) {  // xb_mean = sum(xb) * nmember_inv in real32 (core:671)
      float s = 0.f;
      for (int m = 0; m < k; ++m) s = LK_ADD(s, xb32[m]);
      s_mean = (T)LK_MUL(s, ninv);
    }
    __syncthreads();
    const T xmean = s_mean;
    for (int m = tid; m < k; m += nt) xp[m] = (T)xb32[m] - xmean;  // core:672
    __syncthreads();
    // z = U^T xb' scaled by sqrt(k-1)/sqrt(lambda); warp per column, coalesced over rows
    for (int j = warp; j < k; j += nw) {
      const T *uj = Uu + (int64_t)j * k;
      T a = 0;
      for (int i = lane; i < k; i += 32) a += uj[i] * xp[i];
      a = warp_sum_x(a);
      if (lane == 0) zs[j] = a * sk / sqrt(lu[j]);
    }
    // s = xb' . wbar
    {
      T a = 0;
      for (int i = tid; i < k; i += nt) a += xp[i] * wb[i];
      a = warp_sum_x(a);
      if (lane == 0) red[warp] = a;
    }
    __syncthreads();
    T sdot = 0;
    for (int w = 0; w < nw; ++w) sdot += red[w];
    for (int m = tid; m < k; m += nt) {
      T a = 0;
      for (int j = 0; j < k; ++j) a += Uu[m + (int64_t)j * k] * zs[j];
      T xa = xmean + (sdot + a);  // core:673-675
      if (isnan_unit) xa = xa * T(NAN);
      if (xa_raw) xa_raw[pt * k + m] = (double)xa;
      xa32[m] = (float)xa;  // core:679
    }
    __syncthreads();
    if (use_rtpp || use_rtps) {  // core:684-698
      if (tid == 0) {
        float s = 0.f;
        for (int m = 0; m < k; ++m) s = LK_ADD(s, xa32[m]);
        s_f[0] = LK_MUL(s, ninv);  // xa_mean
      }
      __syncthreads();
      const float xa_mean = s_f[0];
      for (int m = tid; m < k; m += nt) {
        float xap = LK_SUB(xa32[m], xa_mean);
        if (use_rtpp) {  // real32*real32 + real32*T, assigned to real32 (core:689)
          const float t1 = LK_MUL(LK_SUB(1.0f, rtpp_alpha), xap);
          xap = (float)((T)t1 + (T)rtpp_alpha * xp[m]);
        }
        xa32[m] = xap;
      }
      __syncthreads();
      if (use_rtps) {
        if (tid == 0) {  // core:692-694
          T d = 0;
          for (int m = 0; m < k; ++m) d += xp[m] * xp[m];
          const float xb_std = (float)d;
          float xa_std = 0.f;
          for (int m = 0; m < k; ++m) xa_std = LK_ADD(xa_std, LK_MUL(xa32[m], xa32[m]));
          s_f[1] = LK_ADD(LK_SUB(LK_MUL(rtps_alpha, LK_SQRT(LK_DIV(xb_std, xa_std))), rtps_alpha), 1.0f);
        }
        __syncthreads();
        const float fac = s_f[1];
        for (int m = tid; m < k; m += nt) xa32[m] = LK_MUL(xa32[m], fac);
        __syncthreads();
      }
      for (int m = tid; m < k; m += nt) xa32[m] = LK_ADD(xa_mean, xa32[m]);  // core:697
      __syncthreads();
    }
    for (int m = tid; m < k; m += nt) v[(int64_t)m * npts_total + pt] = xa32[m];  // core:229
  }
}

template <typename T>
void launch_transform(cudaStream_t s, int k, int64_t nunits, const int32_t *unit_pt, int64_t npts_total,
                      int64_t pt_base, const T *U, const T *lam, const T *wbar, const int32_t *nanflag,
                      int nfields, float *var, int use_rtpp, float rtpp_alpha, int use_rtps,
                      float rtps_alpha, double *xa_raw) {
  if (nunits == 0 || nfields == 0) return;
  const size_t smem = sizeof(T) * (2 * (size_t)k + 8) + sizeof(float) * 2 * (size_t)k;
  transform_kernel<T><<<(unsigned)nunits, 256, smem, s>>>(k, nunits, unit_pt, npts_total, pt_base, U, lam,
                                                          wbar, nanflag, nfields, var, use_rtpp, rtpp_alpha,
                                                          use_rtps, rtps_alpha, xa_raw);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_transform<double>(cudaStream_t, int, int64_t, const int32_t *, int64_t, int64_t,
                                       const double *, const double *, const double *, const int32_t *, int,
                                       float *, int, float, int, float, double *);
template void launch_transform<float>(cudaStream_t, int, int64_t, const int32_t *, int64_t, int64_t,
                                      const float *, const float *, const float *, const int32_t *, int,
                                      float *, int, float, int, float, double *);

// ---- parity dump: wbar and Wa = sqrt(k-1) U Lambda^(-1/2) U^T (core:662-668) -------------------
template <typename T>
__global__ void __launch_bounds__(256)
    weights_dump_kernel(int k, int64_t nunits, const int32_t *__restrict__ unit_pt, const T *__restrict__ U,
                        const T *__restrict__ lam, const T *__restrict__ wbar, double *__restrict__ wbar_out,
                        double *__restrict__ Wa_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sc = reinterpret_cast<T *>(smem_raw);  // [k]
  const int64_t unit = blockIdx.x;
  if (unit >= nunits) return;
  const int64_t pt = unit_pt[unit];
  const T *Uu = U + unit * (int64_t)k * k;
  const T sk = sqrt((T)(k - 1));
  for (int j = threadIdx.x; j < k; j += blockDim.x) sc[j] = sk / sqrt(lam[unit * (int64_t)k + j]);
  __syncthreads();
  if (wbar_out)
    for (int i = threadIdx.x; i < k; i += blockDim.x) wbar_out[pt * k + i] = (double)wbar[unit * (int64_t)k + i];
  if (Wa_out)
    for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
      const int i = e % k, j = e / k;
      T a = 0;
      for (int l = 0; l < k; ++l) a += Uu[i + (int64_t)l * k] * sc[l] * Uu[j + (int64_t)l * k];
      Wa_out[pt * (int64_t)k * k + e] = (double)a;
    }
}

template <typename T>
void launch_weights_dump(cudaStream_t s, int k, int64_t nunits, const int32_t *unit_pt, const T *U,
                         const T *lam, const T *wbar, double *wbar_out, double *Wa_out) {
  if (nunits == 0) return;
  weights_dump_kernel<T><<<(unsigned)nunits, 256, sizeof(T) * k, s>>>(k, nunits, unit_pt, U, lam, wbar,
                                                                      wbar_out, Wa_out);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_weights_dump<double>(cudaStream_t, int, int64_t, const int32_t *, const double *,
                                          const double *, const double *, double *, double *);
template void launch_weights_dump<float>(cudaStream_t, int, int64_t, const int32_t *, const float *,
                                         const float *, const float *, double *, double *);

// ---- FMA-peak micro-benchmark (roofline denominator of the eigen / Gram stages) ----------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters) {
  T a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
  const T x = (T)1.0000001, y = (T)1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i] * x + y;
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == (T)123.456) out[0] = s;
}

// kind 2: FP64 tensor pipe alone (mma.sync.m8n8k4.f64, 8 independent accumulator tiles per warp);
// kind 3: the same DMMA stream interleaved with an equal number of independent DFMA chains in every warp --
// tells whether the FP64 tensor pipe and the FP64 FMA pipe overlap on this GPU (reported: DMMA flop + FMA flop).
template <bool WITH_FMA>
__global__ void __launch_bounds__(256) dmma_peak_kernel(double *out, int iters) {
  double c[8][2], a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i][0] = c[i][1] = 0.0;
    a[i] = (threadIdx.x + i) * 1e-3;
  }
  const double fa = 1.0 + 1e-9 * threadIdx.x, fb = 1e-9, x = 1.0000001, y = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1])
                     : "d"(fa), "d"(fb));
      if (WITH_FMA) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = a[i] * x + y;
      }
    }
  }
  double sacc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc += c[i][0] + c[i][1] + a[i];
  if (sacc == 123.456) out[0] = sacc;
}

double run_fma_peak(cudaStream_t s, int kind) {
  int dev = 0, sms = 0;
  LK_CUDA(cudaGetDevice(&dev));
  LK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void *out = nullptr;
  LK_CUDA(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  LK_CUDA(cudaEventCreate(&e0));
  LK_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, iters = kind == 0 ? 4000 : (kind == 1 ? 16000 : 2000);
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    LK_CUDA(cudaEventRecord(e0, s));
    if (kind == 0)
      fma_peak_kernel<double><<<blocks, 256, 0, s>>>((double *)out, iters);
    else if (kind == 1)
      fma_peak_kernel<float><<<blocks, 256, 0, s>>>((float *)out, iters);
    else if (kind == 2)
      dmma_peak_kernel<false><<<blocks, 256, 0, s>>>((double *)out, iters);
    else
      dmma_peak_kernel<true><<<blocks, 256, 0, s>>>((double *)out, iters);
    launch_counter()++;
    LK_CUDA(cudaEventRecord(e1, s));
    LK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    LK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // per thread and iteration: 64 FMA (kinds 0, 1); 32 DMMA of 2*8*8*4/32 flop per lane (+ 64 FMA, kind 3)
    const double per_thread = kind <= 1 ? 2.0 * 64.0 : (32.0 * 16.0 + (kind == 3 ? 2.0 * 64.0 : 0.0));
    const double flops = per_thread * iters * 256.0 * blocks;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // namespace lk
