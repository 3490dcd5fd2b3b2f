// Batched symmetric eigensolver -- the ?syevd('V','L') of module_eigen.f90:49/66, one matrix per
// analysis unit -- and the Pa~ b product that follows it (module_letkf_core.f90:651-652).
//
// Method: for a symmetric positive definite C (the LETKF matrix (k-1)/rho I + Yb Yb^T always is),
// factor C = L L^T (Cholesky) and orthogonalise the COLUMNS of L by one-sided (Hestenes) Jacobi
// rotations, L J1 J2 ... = U Sigma.  Then C = U Sigma^2 U^T: eigenvalues are the squared column
// norms, eigenvectors the normalised columns -- no separate eigenvector accumulation, and the
// Cholesky factor preconditions the iteration (Veselic-Hari), so 5-7 sweeps reach working
// precision.  General symmetric input (the stand-alone eigensolver entry point) is shifted by a
// Gershgorin bound to make it definite and the shift is removed from the eigenvalues.
//
// This file holds the generic block-per-matrix kernel (any k <= 256, either precision): the
// factor lives in shared memory when k*k*sizeof(T) fits, otherwise in place in global memory
// (L2-resident).  Column pairs of a round-robin ordering are processed one pair per warp with
// shuffle reductions for the three inner products.
#include "letkf_internal.cuh"

namespace lk {

template <typename T>
struct Eps;
template <>
struct Eps<double> {
  static __device__ __forceinline__ double tol(int k) { return 2.220446049250313e-16 * sqrt((double)k); }
  static __device__ __forceinline__ double tiny() { return 1e-290; }
};
template <>
struct Eps<float> {
  static __device__ __forceinline__ float tol(int k) { return 1.1920929e-7f * sqrtf((float)k); }
  static __device__ __forceinline__ float tiny() { return 1e-30f; }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Cholesky in place on the lower triangle of G (column-major, ld = k), then zero the strict
// upper triangle.  Whole CTA participates.
template <typename T>
__device__ bool block_cholesky(T *G, int k) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  bool ok = true;  // every thread reads the same pivots, so this is CTA-uniform
  for (int j = 0; j < k; ++j) {
    __syncthreads();
    const T piv = G[j + (size_t)j * k];
    if (!(piv > T(0))) ok = false;
    const T d = sqrt(piv);
    __syncthreads();
    if (tid == 0) G[j + (size_t)j * k] = d;
    const T dinv = T(1) / d;
    for (int i = j + 1 + tid; i < k; i += nt) G[i + (size_t)j * k] *= dinv;
    __syncthreads();
    for (int c = j + 1 + warp; c < k; c += nw) {
      const T f = G[c + (size_t)j * k];
      for (int i = c + lane; i < k; i += 32) G[i + (size_t)c * k] -= G[i + (size_t)j * k] * f;
    }
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += nt) {
    const int i = e % k, j = e / k;
    if (i < j) G[e] = T(0);
  }
  __syncthreads();
  return ok;
}

// One-sided Jacobi sweeps on the columns of G until every pair is orthogonal to tolerance.
// Returns the number of sweeps.  Round-robin (tournament) ordering on kk = k rounded up to even.
template <typename T>
__device__ int block_jacobi(T *G, int k) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int kk = (k + 1) & ~1;
  const int half = kk / 2, nm1 = kk - 1;
  const T tol = Eps<T>::tol(k);
  int sweeps = 0;
  for (; sweeps < 60; ++sweeps) {
    int rotated = 0;
    for (int step = 0; step < nm1; ++step) {
      for (int pi = warp; pi < half; pi += nw) {
        int p, q;
        if (pi == 0) {
          p = nm1;
          q = step;
        } else {
          p = (step + pi) % nm1;
          q = (step - pi + nm1) % nm1;
        }
        if (p > q) {
          const int t = p;
          p = q;
          q = t;
        }
        if (q >= k) continue;  // dummy column of the odd-k padding
        T *gp = G + (size_t)p * k, *gq = G + (size_t)q * k;
        T a = 0, b = 0, g = 0;
        for (int i = lane; i < k; i += 32) {
          const T x = gp[i], y = gq[i];
          a += x * x;
          b += y * y;
          g += x * y;
        }
        a = warp_sum(a);
        b = warp_sum(b);
        g = warp_sum(g);
        if (fabs(g) > tol * sqrt(a * b) && fabs(g) > Eps<T>::tiny()) {
          const T zeta = (b - a) / (T(2) * g);
          const T t = (zeta >= 0 ? T(1) : T(-1)) / (fabs(zeta) + sqrt(T(1) + zeta * zeta));
          const T c = T(1) / sqrt(T(1) + t * t);
          const T s = c * t;
          for (int i = lane; i < k; i += 32) {
            const T x = gp[i], y = gq[i];
            gp[i] = c * x - s * y;
            gq[i] = s * x + c * y;
          }
          rotated = 1;
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) {
      ++sweeps;
      break;
    }
  }
  return sweeps;
}

// mode 0: LETKF solve.  in: C (full symmetric, SPD), b.  out: U (in place of C), lam, wbar.
// mode 1: ?syevd.      in: A (lower).                  out: W ascending, V.
template <typename T, int MODE, bool SMEM>
__global__ void __launch_bounds__(256)
    eig_block_kernel(int k, int64_t n, T *__restrict__ Cio, const T *__restrict__ bvec, T *__restrict__ lam,
                     T *__restrict__ wbar, const T *__restrict__ Ain, T *__restrict__ Wout,
                     T *__restrict__ Vout, int32_t *__restrict__ sweeps_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *>(smem_raw);
  const int64_t u = blockIdx.x;
  if (u >= n) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  T *vec1 = sm;        // [k]
  T *vec2 = sm + k;    // [k]
  T *Gs = sm + 2 * k;  // [k*k] when SMEM
  T *Gg = MODE == 0 ? Cio + u * (int64_t)k * k : Vout + u * (int64_t)k * k;
  T *G = SMEM ? Gs : Gg;
  __shared__ T s_shift;

  if (MODE == 0) {
    if (SMEM)
      for (int e = tid; e < k * k; e += nt) G[e] = Gg[e];
    if (tid == 0) s_shift = T(0);
    __syncthreads();
    block_cholesky(G, k);  // C is SPD by construction; a NaN input propagates
  } else {
    const T *A = Ain + u * (int64_t)k * k;
    __shared__ T red_lo[256], red_sc[256];
    // First try the matrix as it is (an SPD input keeps its full relative accuracy); if a pivot
    // fails, shift by a Gershgorin bound so that it becomes definite and factor again.
    for (int attempt = 0; attempt < 2; ++attempt) {
      __syncthreads();
      for (int e = tid; e < k * k; e += nt) {  // symmetrise from the lower triangle
        const int i = e % k, j = e / k;
        G[e] = i >= j ? A[i + (size_t)j * k] : A[j + (size_t)i * k];
      }
      if (tid == 0 && attempt == 0) s_shift = T(0);
      __syncthreads();
      if (attempt == 1) {
        T lowest = sizeof(T) == 8 ? T(1e300) : T(3e38);
        T scale = T(0);
        for (int i = tid; i < k; i += nt) {
          T off = 0;
          for (int j = 0; j < k; ++j)
            if (j != i) off += fabs(G[i + (size_t)j * k]);
          lowest = min(lowest, G[i + (size_t)i * k] - off);
          scale = max(scale, fabs(G[i + (size_t)i * k]) + off);
        }
        red_lo[tid] = lowest;
        red_sc[tid] = scale;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
          if (tid < o) {
            red_lo[tid] = min(red_lo[tid], red_lo[tid + o]);
            red_sc[tid] = max(red_sc[tid], red_sc[tid + o]);
          }
          __syncthreads();
        }
        if (tid == 0) s_shift = red_sc[0] * T(1e-3) - min(red_lo[0], T(0));
        __syncthreads();
        const T sh = s_shift;
        for (int i = tid; i < k; i += nt) G[i + (size_t)i * k] += sh;
        __syncthreads();
      }
      if (block_cholesky(G, k)) break;
    }
  }
  const int sweeps = block_jacobi(G, k);
  if (tid == 0 && sweeps_max) atomicMax(sweeps_max, sweeps);

  // column norms -> eigenvalues; normalise columns -> eigenvectors
  for (int j = warp; j < k; j += nw) {
    T *gj = G + (size_t)j * k;
    T a = 0;
    for (int i = lane; i < k; i += 32) a += gj[i] * gj[i];
    a = warp_sum(a);
    const T inv = T(1) / sqrt(a);
    for (int i = lane; i < k; i += 32) gj[i] *= inv;
    if (lane == 0) vec1[j] = a;  // lambda (+ shift)
  }
  __syncthreads();

  if (MODE == 0) {
    // wbar = U diag(1/lam) U^T b   (inverse_matrix + ?gemv + ?symv, eig:37-76, core:651-652)
    const T *b = bvec + u * (int64_t)k;
    for (int j = warp; j < k; j += nw) {
      const T *gj = G + (size_t)j * k;
      T a = 0;
      for (int i = lane; i < k; i += 32) a += gj[i] * b[i];
      a = warp_sum(a);
      if (lane == 0) vec2[j] = a / vec1[j];
    }
    __syncthreads();
    for (int i = tid; i < k; i += nt) {
      T a = 0;
      for (int j = 0; j < k; ++j) a += G[i + (size_t)j * k] * vec2[j];
      wbar[u * (int64_t)k + i] = a;
      lam[u * (int64_t)k + i] = vec1[i];
    }
    if (SMEM)
      for (int e = tid; e < k * k; e += nt) Gg[e] = G[e];
  } else {
    // ascending order like LAPACK: rank by counting
    const T sh = s_shift;
    for (int j = tid; j < k; j += nt) {
      const T lj = vec1[j];
      int r = 0;
      for (int l = 0; l < k; ++l) {
        const T ll = vec1[l];
        r += (ll < lj) || (ll == lj && l < j);
      }
      ((int *)vec2)[j] = r;
      Wout[u * (int64_t)k + r] = lj - sh;
    }
    __syncthreads();
    if (SMEM) {
      for (int e = tid; e < k * k; e += nt) {
        const int i = e % k, j = e / k;
        Gg[i + (size_t)((int *)vec2)[j] * k] = G[e];
      }
    } else {
      // in-place column permutation in global memory: cycle-follow, one thread per row
      for (int i = tid; i < k; i += nt) {
        // rows are independent; permute row i across columns using a register-free cycle walk
        for (int start = 0; start < k; ++start) {
          // process each cycle once, from its smallest index
          int c = ((int *)vec2)[start];
          bool smallest = true;
          while (c != start) {
            if (c < start) { smallest = false; break; }
            c = ((int *)vec2)[c];
          }
          if (!smallest) continue;
          T carry = G[i + (size_t)start * k];
          int dst = ((int *)vec2)[start];
          while (dst != start) {
            const T tmp = G[i + (size_t)dst * k];
            G[i + (size_t)dst * k] = carry;
            carry = tmp;
            dst = ((int *)vec2)[dst];
          }
          G[i + (size_t)start * k] = carry;
        }
      }
    }
  }
}

template <typename T>
static size_t eig_smem_bytes(int k, bool smem_matrix) {
  return sizeof(T) * (2 * (size_t)k + (smem_matrix ? (size_t)k * k : 0));
}

template <typename T, int MODE>
static void launch_eig_generic(cudaStream_t s, int k, int64_t n, T *Cio, const T *b, T *lam, T *wbar,
                               const T *A, T *W, T *V, int32_t *sweeps_max) {
  if (n == 0) return;
  LK_REQUIRE(k >= 2 && k <= LETKF_B200_MAX_MEMBERS, "eigensolver: need 2 <= k <= 256");
  const bool smem_ok = eig_smem_bytes<T>(k, true) <= 200 * 1024;
  const size_t smem = eig_smem_bytes<T>(k, smem_ok);
  for (int64_t u0 = 0; u0 < n; u0 += 1 << 30) {
    const int64_t nu = std::min<int64_t>(n - u0, 1 << 30);
    const int64_t o2 = u0 * (int64_t)k * k, o1 = u0 * (int64_t)k;
    if (smem_ok) {
      auto kern = eig_block_kernel<T, MODE, true>;
      LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<(unsigned)nu, 256, smem, s>>>(k, nu, Cio ? Cio + o2 : nullptr, b ? b + o1 : nullptr,
                                           lam ? lam + o1 : nullptr, wbar ? wbar + o1 : nullptr,
                                           A ? A + o2 : nullptr, W ? W + o1 : nullptr, V ? V + o2 : nullptr,
                                           sweeps_max);
    } else {
      auto kern = eig_block_kernel<T, MODE, false>;
      kern<<<(unsigned)nu, 256, smem, s>>>(k, nu, Cio ? Cio + o2 : nullptr, b ? b + o1 : nullptr,
                                           lam ? lam + o1 : nullptr, wbar ? wbar + o1 : nullptr,
                                           A ? A + o2 : nullptr, W ? W + o1 : nullptr, V ? V + o2 : nullptr,
                                           sweeps_max);
    }
    launch_counter()++;
  }
  LK_CUDA(cudaGetLastError());
}

template <typename T>
void launch_eig_solve(cudaStream_t s, int k, int64_t nunits, T *C_inout_U, const T *b, T *lam, T *wbar,
                      int32_t *sweeps_max) {
  launch_eig_generic<T, 0>(s, k, nunits, C_inout_U, b, lam, wbar, nullptr, nullptr, nullptr, sweeps_max);
}
template <typename T>
void launch_syevd(cudaStream_t s, int k, int64_t batch, const T *A, T *W, T *V, int32_t *sweeps_max) {
  launch_eig_generic<T, 1>(s, k, batch, nullptr, nullptr, nullptr, nullptr, A, W, V, sweeps_max);
}

template void launch_eig_solve<double>(cudaStream_t, int, int64_t, double *, const double *, double *,
                                       double *, int32_t *);
template void launch_eig_solve<float>(cudaStream_t, int, int64_t, float *, const float *, float *, float *,
                                      int32_t *);
template void launch_syevd<double>(cudaStream_t, int, int64_t, const double *, double *, double *, int32_t *);
template void launch_syevd<float>(cudaStream_t, int, int64_t, const float *, float *, float *, int32_t *);

}  // namespace lk
