// k = 32 eigensolver: one warp per matrix, the whole factor in registers.
//
// Same method as kernels_eig.cu (Cholesky C = L L^T, one-sided Jacobi on the columns of L), laid out
// for a warp: lane i owns ROW i of the factor in 32 registers, so a column rotation
// (g_p, g_q) <- (c g_p - s g_q, s g_p + c g_q) is purely lane-local with compile-time register
// indices.  Only the pair inner products cross lanes: the 16 products of a round-robin step are
// summed with one transposed butterfly (16 shuffled doubles, not 16 x 5), after which lanes 2n
// and 2n+1 both hold gamma_n and compute the rotation of pair n.  Column norms are carried along
// with the rotation update formulas (alpha - t gamma, beta + t gamma) and refreshed exactly once
// per sweep; (c, s) reach every lane through a 256-byte shared-memory broadcast.  Pairs are always
// the register pairs (2n, 2n+1): after every step the registers are permuted by the fixed
// round-robin rotation (31 steps return to the identity), so the step body is one small loop and
// stays in the instruction cache.
//
// Double-precision sqrt / divide in the rotation are replaced by MUFU seeds + Newton steps
// (2 iterations: full double accuracy); exact orthogonality only needs c^2 + s^2 = 1, which
// c = rsqrt(1 + t^2), s = c t delivers.
#include <cstdlib>

#include "eig_common.cuh"
#include "xform32.cuh"

namespace lk {

constexpr int K32 = 32;
constexpr int JACOBI32_CAP = 30;  // == LK_JACOBI_CAP32 (letkf_internal.cuh)
constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ constexpr int rr_pos(int m) { return (m & 1) ? 31 - (m >> 1) : (m >> 1); }
__host__ __device__ constexpr int rr_reg(int p) { return p <= 15 ? 2 * p : 2 * (31 - p) + 1; }
// register that feeds register m in the round-robin rotation
__host__ __device__ constexpr int rr_src(int m) {
  return rr_reg(rr_pos(m) == 0 ? 0 : (rr_pos(m) == 1 ? 31 : rr_pos(m) - 1));
}

// sum v[0..N) over the warp; afterwards v[0] in lane l is the total of element (l >> (5 - log2 N))
// -- for N = 16 element l>>1, for N = 32 element l.
template <typename T, int N>
__device__ __forceinline__ void transposed_reduce(T (&v)[N], int lane) {
#pragma unroll
  for (int n = N, mask = 16; n > 1; n >>= 1, mask >>= 1) {
    const bool up = lane & mask;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const T send = up ? v[i] : v[i + n / 2];
      const T keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, mask);
    }
  }
  if (N == 16) v[0] += __shfl_xor_sync(FULL, v[0], 1);
}

// Cholesky of the matrix whose row `lane` (entries j <= lane) is in g[]; leaves L there (zeros above
// the diagonal).  colbuf: 32 T of shared memory private to the warp.  Returns false if a pivot is
// not positive (warp-uniform).
template <typename T>
__device__ __forceinline__ bool warp_cholesky32(T (&g)[K32], int lane, T *colbuf) {
  bool ok = true;
#pragma unroll
  for (int j = 0; j < K32; ++j) {
    const T piv = __shfl_sync(FULL, g[j], j);
    ok = ok && (piv > T(0));
    const T rinv = Fast<T>::rsqrt(piv > T(0) ? piv : T(1));
    const T l = lane >= j ? g[j] * rinv : T(0);
    g[j] = l;
    if (j < K32 - 1) {
      __syncwarp();
      colbuf[lane] = l;
      __syncwarp();
#pragma unroll
      for (int m = j + 1; m < K32; ++m) g[m] = fma(-l, colbuf[m], g[m]);
    }
  }
#pragma unroll
  for (int m = 1; m < K32; ++m)
    if (m > lane) g[m] = T(0);
  return ok;
}

// one-sided Jacobi on the columns held as registers; returns the sweep count.
// csbuf: 32 T ((c,s) x 16 pairs); part: 16*32 T of shared memory private to the warp for the
// gamma reduction.  stop2: once a whole sweep has seen only cos^2 <= stop2, quadratic convergence
// puts every cosine after that sweep below ~stop2 (measured: 1e-6 -> 4e-13, 2.5e-9 -> 2e-16), so
// the sweep that would merely verify convergence is skipped.
// FASTROT: the rotation ANGLE t is evaluated in real32 (MUFU rsqrt/rcp, ~8 short-latency ops instead
// of ~15 dependent FP64 ops); c = rsqrt(1 + t^2), s = c t stay in working precision, so every
// rotation is orthogonal to working accuracy and only the annihilation of gamma is approximate
// (residual cosine ~1e-7 x the old one, far inside the quadratic-convergence budget).  Needs squared
// column norms within real32 range, which holds for LETKF matrices (eigenvalues >= (k-1)/rho).
template <typename T, bool FASTROT>
__device__ __forceinline__ int warp_jacobi32(T (&g)[K32], int lane, T *csbuf, T *part, T stop2) {
  const T tol2 = Fast<T>::tol2(K32);
  // lane that holds, before a rotation step, the norm this lane's register slot receives
  const int p = (lane & 1) ? 31 - (lane >> 1) : (lane >> 1);
  const int pp = p == 0 ? 0 : (p == 1 ? 31 : p - 1);
  const int src_lane = pp <= 15 ? 2 * pp : 2 * (31 - pp) + 1;
  const T *mypart = part + (lane >> 1) * 32 + (lane & 1) * 16;
  int sweeps = 0;
  for (; sweeps < JACOBI32_CAP; ++sweeps) {
    // exact squared column norms: lane l <- ||column in register slot l||^2
    T d;
    {
      T sq[K32];
#pragma unroll
      for (int j = 0; j < K32; ++j) sq[j] = g[j] * g[j];
      transposed_reduce<T, K32>(sq, lane);
      d = sq[0];
    }
    int rotated = 0, big = 0;
#pragma unroll 1
    for (int step = 0; step < K32 - 1; ++step) {
      // gamma_n = sum over lanes of g[2n] g[2n+1]: partial products through shared memory; lane
      // 2n+h adds the 16 partials of lanes 16h..16h+15, then one exchange with its partner.  The
      // read index is skewed by the lane so that every bank serves exactly two lanes.
#pragma unroll
      for (int n = 0; n < 16; ++n) part[n * 32 + lane] = g[2 * n] * g[2 * n + 1];
      __syncwarp();
      T gamma;
      {
        T acc[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r & 3] += mypart[(r + lane) & 15];
        gamma = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        gamma += __shfl_xor_sync(FULL, gamma, 1);
      }
      const T dpart = __shfl_xor_sync(FULL, d, 1);
      const T alpha = (lane & 1) ? dpart : d;
      const T beta = (lane & 1) ? d : dpart;
      const T g2 = gamma * gamma, ab = alpha * beta;
      const bool rot = g2 > tol2 * ab;
      big |= g2 > stop2 * ab;
      T c, s, t;
      jacobi_rotation<T, FASTROT>(alpha, beta, gamma, rot, c, s, t);
      rotated |= rot;
      d = (lane & 1) ? fma(t, gamma, d) : fma(-t, gamma, d);  // beta + t gamma | alpha - t gamma
      if (!(lane & 1)) {
        csbuf[lane] = c;      // pair n = lane/2 -> csbuf[2n], csbuf[2n+1]
        csbuf[lane + 1] = s;
      }
      __syncwarp();
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const T cn = csbuf[2 * n], sn = csbuf[2 * n + 1];
        const T a = g[2 * n], b = g[2 * n + 1];
        g[2 * n] = fma(cn, a, -(sn * b));
        g[2 * n + 1] = fma(sn, a, cn * b);
      }
      // round-robin rotation of the register slots and of the norms that travel with them
      {
        T ng[K32];
#pragma unroll
        for (int m = 0; m < K32; ++m) ng[m] = g[rr_src(m)];
#pragma unroll
        for (int m = 0; m < K32; ++m) g[m] = ng[m];
      }
      d = __shfl_sync(FULL, d, src_lane);
      __syncwarp();  // csbuf / part are rewritten by the next step
    }
    if (!__any_sync(FULL, rotated) || !__any_sync(FULL, big)) return sweeps + 1;
  }
  return JACOBI32_CAP + 1;  // not converged: reported as an error by the caller (api.cu)
}

// MODE 0: C (row-major lower / symmetric full, SPD), b -> U^T-by-rows (U[i][j] at i*32+j), lam, wbar
// MODE 1: A (column-major, lower referenced) -> W ascending, V column-major
template <typename T, int MODE, int MINB>
__global__ void __launch_bounds__(128, MINB)
    eig32_warp_kernel(int64_t n, T *__restrict__ Cio, const T *__restrict__ bvec, T *__restrict__ lam,
                      T *__restrict__ wbar, const T *__restrict__ Ain, T *__restrict__ Wout,
                      T *__restrict__ Vout, int32_t *__restrict__ sweeps_max) {
  __shared__ __align__(16) T sbuf[4][64 + 16 * 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t u = (int64_t)blockIdx.x * 4 + w;
  if (u >= n) return;
  T *buf = sbuf[w];
  T g[K32];
  T shift = T(0), scale = T(1);

  if (MODE == 0) {
    const T *Cu = Cio + u * (int64_t)(K32 * K32) + (int64_t)lane * K32;
#pragma unroll
    for (int j = 0; j < K32; j += 2) {
      const T x0 = Cu[j], x1 = Cu[j + 1];
      g[j] = j <= lane ? x0 : T(0);
      g[j + 1] = j + 1 <= lane ? x1 : T(0);
    }
    warp_cholesky32<T>(g, lane, buf);
  } else {
    const T *A = Ain + u * (int64_t)(K32 * K32);
    T full[K32];
#pragma unroll
    for (int j = 0; j < K32; ++j) full[j] = j <= lane ? A[lane + j * K32] : A[j + lane * K32];
    // scale by the largest |diagonal| so that the float seeds of rsqrt/rcp stay in range
    T dmax = fabs(full[0]);
#pragma unroll
    for (int j = 1; j < K32; ++j) dmax = j == lane ? fabs(full[j]) : dmax;
    dmax = lane == 0 ? fabs(full[0]) : dmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = max(dmax, __shfl_xor_sync(FULL, dmax, o));
    scale = dmax > T(0) ? dmax : T(1);
    const T iscale = T(1) / scale;
#pragma unroll
    for (int j = 0; j < K32; ++j) full[j] *= iscale;
#pragma unroll
    for (int j = 0; j < K32; ++j) g[j] = j <= lane ? full[j] : T(0);
    if (!warp_cholesky32<T>(g, lane, buf)) {
      // not positive definite: shift by a Gershgorin bound and factor again
      T off = T(0), dg = T(0);
#pragma unroll
      for (int j = 0; j < K32; ++j) {
        off += j == lane ? T(0) : fabs(full[j]);
        dg = j == lane ? full[j] : dg;
      }
      T lo = dg - off, sc = fabs(dg) + off;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(FULL, lo, o));
        sc = max(sc, __shfl_xor_sync(FULL, sc, o));
      }
      shift = sc * T(1e-3) - min(lo, T(0));
#pragma unroll
      for (int j = 0; j < K32; ++j) g[j] = j <= lane ? full[j] + (j == lane ? shift : T(0)) : T(0);
      warp_cholesky32<T>(g, lane, buf);
    }
  }

  // MODE 0 feeds the LETKF weights (1e-10 bar): stop once a sweep saw only |cos| <= 1e-7.
  // MODE 1 is the general eigensolver: |cos| <= 1e-9 before the last sweep.
  const T stop2 = MODE == 0 ? T(1e-14) : (sizeof(T) == 8 ? T(1e-18) : T(1e-9));
  const int sweeps = warp_jacobi32<T, MODE == 0>(g, lane, buf, buf + 64, stop2);
  if (lane == 0 && sweeps_max) atomicMax(sweeps_max, sweeps);

  // eigenvalues = squared column norms (lane j <- lambda_j), eigenvectors = normalised columns
  T lambda;
  {
    T sq[K32];
#pragma unroll
    for (int j = 0; j < K32; ++j) sq[j] = g[j] * g[j];
    transposed_reduce<T, K32>(sq, lane);
    lambda = sq[0];
  }
  __syncwarp();
  buf[lane] = Fast<T>::rsqrt(lambda);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < K32; ++j) g[j] *= buf[j];

  if (MODE == 0) {
    // wbar = U diag(1/lambda) U^T b   (eig:37-76 + core:651-652)
    const T bi = bvec[u * K32 + lane];
    T z;
    {
      T pr[K32];
#pragma unroll
      for (int j = 0; j < K32; ++j) pr[j] = g[j] * bi;
      transposed_reduce<T, K32>(pr, lane);
      z = pr[0] / lambda;
    }
    __syncwarp();
    buf[lane] = z;
    __syncwarp();
    T wb = T(0);
#pragma unroll
    for (int j = 0; j < K32; ++j) wb = fma(g[j], buf[j], wb);
    T *Uo = Cio + u * (int64_t)(K32 * K32) + (int64_t)lane * K32;
#pragma unroll
    for (int j = 0; j < K32; ++j) Uo[j] = g[j];
    lam[u * K32 + lane] = lambda;
    wbar[u * K32 + lane] = wb;
  } else {
    // ascending order like LAPACK
    int rank = 0;
#pragma unroll
    for (int l = 0; l < K32; ++l) {
      const T ll = __shfl_sync(FULL, lambda, l);
      rank += (ll < lambda) || (ll == lambda && l < lane);
    }
    Wout[u * K32 + rank] = (lambda - shift) * scale;
    T *V = Vout + u * (int64_t)(K32 * K32);
#pragma unroll
    for (int j = 0; j < K32; ++j) {
      const int rj = __shfl_sync(FULL, rank, j);
      V[rj * K32 + lane] = g[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Warm-started variant for the LETKF pipeline.  Consecutive analysis units are neighbouring grid
// points (x fastest) whose matrices differ little, so a warp walks a run of RUN consecutive units
// and starts each solve in the eigenbasis of the previous one:
//     C' = U_prev^T C U_prev  (nearly diagonal)  ->  Cholesky + Jacobi  ->  U',  U = U_prev U'.
// Any orthogonal U_prev gives the exact decomposition C = U Sigma^2 U^T, so this changes the number
// of sweeps (measured 7.2 -> ~4 on config M), not the result.  The three 32^3 products run in the
// row-per-lane layout with the second operand broadcast from shared memory.
// Shared memory per warp: A = U_prev (8 KB), B = scratch (T = C U_prev, U', and the Jacobi reduction
// buffer), 64 values for (c,s).
constexpr int LDA32 = 34;  // row stride of the shared-memory matrices: 16-byte aligned rows, <= 4-way conflicts

__device__ __forceinline__ void dmma884e(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// 32x32x32 product on the FP64 tensor pipe: D = op(A) B with fragments fetched by the functors
// fa(row, col) -> A[row][col], fb(row, col) -> B[row][col]; D is stored to `out` (row stride LDA32).
// The tensor pipe is otherwise idle in this kernel, so these products overlap with the FP64-pipe
// Jacobi work of the other resident warps.
template <typename FA, typename FB>
__device__ __forceinline__ void warp_gemm32_dmma(FA fa, FB fb, double *out, int lane) {
  const int lr = lane >> 2, lc = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int I = 0; I < 4; ++I)
#pragma unroll
    for (int J = 0; J < 4; ++J) acc[I][J][0] = acc[I][J][1] = 0.0;
#pragma unroll 2
  for (int kk = 0; kk < 8; ++kk) {
    double a[4], b[4];
#pragma unroll
    for (int I = 0; I < 4; ++I) a[I] = fa(8 * I + lr, 4 * kk + lc);
#pragma unroll
    for (int J = 0; J < 4; ++J) b[J] = fb(4 * kk + lc, 8 * J + lr);
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = 0; J < 4; ++J) dmma884e(acc[I][J][0], acc[I][J][1], a[I], b[J]);
  }
  __syncwarp();  // every lane has finished reading the operands (out may alias one of them)
#pragma unroll
  for (int I = 0; I < 4; ++I)
#pragma unroll
    for (int J = 0; J < 4; ++J) {
      double2 v;
      v.x = acc[I][J][0];
      v.y = acc[I][J][1];
      *reinterpret_cast<double2 *>(out + (8 * I + lr) * LDA32 + 8 * J + 2 * lc) = v;
    }
  __syncwarp();
}

__global__ void __launch_bounds__(128, 3)
    eig32_chain_kernel(int RUN, int64_t n, double *__restrict__ Cio, const double *__restrict__ bvec,
                       double *__restrict__ lam, double *__restrict__ wbar, int32_t *__restrict__ sweeps_max,
                       int32_t *__restrict__ sweeps_sum, bool fused, Xform32Args xa) {
  using T = double;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T *base = reinterpret_cast<T *>(smem_raw) + (size_t)w * (2 * 32 * LDA32 + 64);
  T *Abuf = base;                 // U_prev, [l][j] with row stride LDA32
  T *Bbuf = base + 32 * LDA32;    // scratch (T = C U_prev, C', U', Jacobi reduction buffer), stride LDA32
  T *cs = Bbuf + 32 * LDA32;      // 64
  const int64_t u0 = ((int64_t)blockIdx.x * 4 + w) * RUN;
  if (u0 >= n) return;
  const int64_t u1 = u0 + RUN < n ? u0 + RUN : n;
  int sw_max = 0, sw_sum = 0;
  bool prev_ok = false;
  for (int64_t u = u0; u < u1; ++u) {
    const bool warm = prev_ok;
    const T *Cu = Cio + u * (int64_t)(K32 * K32);
    const int ldc = K32;
    T g[K32];
    if (warm) {
      // T = C U_prev  (C from global memory, L1/L2 resident)
      warp_gemm32_dmma([&](int r, int c) { return Cu[r * ldc + c]; },
                       [&](int r, int c) { return Abuf[r * LDA32 + c]; }, Bbuf, lane);
      // C' = U_prev^T T   (written over T)
      warp_gemm32_dmma([&](int r, int c) { return Abuf[c * LDA32 + r]; },
                       [&](int r, int c) { return Bbuf[r * LDA32 + c]; }, Bbuf, lane);
#pragma unroll
      for (int j = 0; j < K32; ++j) g[j] = Bbuf[lane * LDA32 + j];
      __syncwarp();
    } else {
#pragma unroll
      for (int j = 0; j < K32; ++j) g[j] = Cu[(int64_t)lane * ldc + j];
      __syncwarp();
    }
#pragma unroll
    for (int j = 1; j < K32; ++j) g[j] = j <= lane ? g[j] : T(0);
    warp_cholesky32<T>(g, lane, cs);
    const int sweeps = warp_jacobi32<T, true>(g, lane, cs, Bbuf, T(1e-14));
    sw_max = sweeps > sw_max ? sweeps : sw_max;
    sw_sum += sweeps;
    T lambda;
    {
      T sq[K32];
#pragma unroll
      for (int j = 0; j < K32; ++j) sq[j] = g[j] * g[j];
      transposed_reduce<T, K32>(sq, lane);
      lambda = sq[0];
    }
    __syncwarp();
    cs[lane] = Fast<T>::rsqrt(lambda);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K32; ++j) g[j] *= cs[j];  // row `lane` of U'
    __syncwarp();
    if (warm) {
      // U = U_prev U'  (result lands in Abuf: it is the next U_prev)
#pragma unroll
      for (int j = 0; j < K32; ++j) Bbuf[lane * LDA32 + j] = g[j];
      __syncwarp();
      warp_gemm32_dmma([&](int r, int c) { return Abuf[r * LDA32 + c]; },
                       [&](int r, int c) { return Bbuf[r * LDA32 + c]; }, Abuf, lane);
#pragma unroll
      for (int j = 0; j < K32; ++j) g[j] = Abuf[lane * LDA32 + j];
    } else {
#pragma unroll
      for (int j = 0; j < K32; ++j) Abuf[lane * LDA32 + j] = g[j];
    }
    __syncwarp();
    // a unit whose matrix is not finite / not positive (real32 Gaspari-Cohn NaN rows) must
    // not seed its neighbour
    prev_ok = !__any_sync(FULL, !(lambda > T(0)) || !(lambda < T(1e300)));
    // wbar = U diag(1/lambda) U^T b
    const T bi = bvec[u * K32 + lane];
    T z;
    {
      T pr[K32];
#pragma unroll
      for (int j = 0; j < K32; ++j) pr[j] = g[j] * bi;
      transposed_reduce<T, K32>(pr, lane);
      z = pr[0] / lambda;
    }
    cs[lane] = z;
    __syncwarp();
    T wb = T(0);
#pragma unroll
    for (int j = 0; j < K32; ++j) wb = fma(g[j], cs[j], wb);
    __syncwarp();
    if (fused) {
      // transform the fields right here: U never goes to memory (saves 16 KB of traffic per unit)
      const T scale = sqrt((T)31) * Fast<T>::rsqrt(lambda);
      const bool isnan_unit = xa.nanflag[u] != 0;
      transform32_unit<T>(g, scale, wb, isnan_unit, xa.pt_base + xa.unit_pt[u], xa, cs, lane);
    } else {
      T *Uo = Cio + u * (int64_t)(K32 * K32) + (int64_t)lane * K32;
#pragma unroll
      for (int j = 0; j < K32; ++j) Uo[j] = g[j];
      lam[u * K32 + lane] = lambda;
      wbar[u * K32 + lane] = wb;
    }
    __syncwarp();
  }
  if (lane == 0) {
    if (sweeps_max) atomicMax(sweeps_max, sw_max);
    if (sweeps_sum) atomicAdd(sweeps_sum, sw_sum);
  }
}

template <typename T>
void launch_eig32_solve(cudaStream_t s, int64_t n, T *C_inout_U, const T *b, T *lam, T *wbar,
                        int32_t *sweeps_max, const Xform32Args *fuse) {
  if (n == 0) return;
  static const int occ = [] {
    const char *e = getenv("LETKF_B200_EIG_OCC");
    return e ? atoi(e) : 4;
  }();
  static const int chain = [] {
    const char *e = getenv("LETKF_B200_EIG_CHAIN");
    return e ? atoi(e) : 16;
  }();
  if (chain > 1 && sizeof(T) == 8) {
    const int RUN = chain;
    const size_t smem = sizeof(double) * 4 * (2 * 32 * LDA32 + 64);
    auto kern = eig32_chain_kernel;
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nwarp = (n + RUN - 1) / RUN;
    kern<<<(unsigned)((nwarp + 3) / 4), 128, smem, s>>>(RUN, n, reinterpret_cast<double *>(C_inout_U),
                                                         reinterpret_cast<const double *>(b),
                                                         reinterpret_cast<double *>(lam),
                                                         reinterpret_cast<double *>(wbar), sweeps_max,
                                                         sweeps_max ? sweeps_max + 1 : nullptr, fuse != nullptr,
                                                         fuse ? *fuse : Xform32Args{});
    launch_counter()++;
    LK_CUDA(cudaGetLastError());
    return;
  }
  LK_REQUIRE(fuse == nullptr, "fused transform needs the chained FP64 kernel");
  if (occ == 4)
    eig32_warp_kernel<T, 0, 4><<<(unsigned)((n + 3) / 4), 128, 0, s>>>(n, C_inout_U, b, lam, wbar, nullptr, nullptr,
                                                                        nullptr, sweeps_max);
  else
    eig32_warp_kernel<T, 0, 3><<<(unsigned)((n + 3) / 4), 128, 0, s>>>(n, C_inout_U, b, lam, wbar, nullptr, nullptr,
                                                                        nullptr, sweeps_max);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template <typename T>
void launch_syevd32(cudaStream_t s, int64_t n, const T *A, T *W, T *V, int32_t *sweeps_max) {
  if (n == 0) return;
  eig32_warp_kernel<T, 1, 2><<<(unsigned)((n + 3) / 4), 128, 0, s>>>(n, nullptr, nullptr, nullptr, nullptr, A, W, V,
                                                                      sweeps_max);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_eig32_solve<double>(cudaStream_t, int64_t, double *, const double *, double *, double *,
                                         int32_t *, const Xform32Args *);
template void launch_eig32_solve<float>(cudaStream_t, int64_t, float *, const float *, float *, float *, int32_t *,
                                        const Xform32Args *);
bool eig32_can_fuse() {
  const char *e = getenv("LETKF_B200_EIG_CHAIN");
  const char *f = getenv("LETKF_B200_FUSE");
  return (!e || atoi(e) > 1) && (!f || atoi(f) != 0);
}
template void launch_syevd32<double>(cudaStream_t, int64_t, const double *, double *, double *, int32_t *);
template void launch_syevd32<float>(cudaStream_t, int64_t, const float *, float *, float *, int32_t *);

}  // namespace lk
