// Batched symmetric eigensolver, any k <= 256 -- the ?syevd('V','L') of module_eigen.f90:49/66, one
// matrix per analysis unit -- and the Pa~ b product that follows it (module_letkf_core.f90:651-652).
//
// Method (same as the k = 32 warp kernel): for SPD C factor C = L L^T and orthogonalise the COLUMNS of
// L by one-sided Jacobi rotations, L J1 J2 ... = U Sigma, so C = U Sigma^2 U^T: eigenvalues are squared
// column norms, eigenvectors the normalised columns; no eigenvector accumulation, and the Cholesky
// factor preconditions the iteration.  General symmetric input (stand-alone entry point) is first
// tried unshifted and otherwise shifted by a Gershgorin bound.
//
// One CTA per matrix; the factor lives in shared memory (k <= 160 FP64 / 224 FP32) or in place in
// global memory (L2) for larger k.  A plain shared-memory Jacobi is bandwidth bound: each rotation
// moves 32 B per row pair for 5 FMAs, 3x more than the 128 B/clk/SM the SM delivers per FP64 FMA.
// So the sweep is REGISTER BLOCKED: columns are grouped in blocks of 4; a warp loads two blocks
// (8 columns, rows split over the lanes) into registers, performs the 16 cross rotations as 4 rounds
// of 4 disjoint pairs, and stores them back -- 4x less shared-memory traffic per rotation.  Block
// pairs follow a round-robin schedule (nb-1 outer steps of nb/2 disjoint block pairs, one per warp);
// the 6 pairs inside each block are rotated once per sweep in a separate pass that also refreshes
// the exact column norms.  Within a round the 4 inner products are summed with one transposed
// butterfly so that every 8-lane group ends up with one gamma and computes one rotation; squared
// norms are carried in shared memory with the rotation update formulas.
#include <cstdlib>

#include "eig_common.cuh"

namespace lk {

constexpr unsigned FULLW = 0xffffffffu;

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLW, v, o);
  return v;
}

// 4 per-lane partial sums -> lane l holds the warp total of element l>>3
template <typename T>
__device__ __forceinline__ T reduce4(T v0, T v1, T v2, T v3, int lane) {
  const bool u16 = lane & 16, u8 = lane & 8;
  const T w0 = (u16 ? v2 : v0) + __shfl_xor_sync(FULLW, u16 ? v0 : v2, 16);
  const T w1 = (u16 ? v3 : v1) + __shfl_xor_sync(FULLW, u16 ? v1 : v3, 16);
  T z = (u8 ? w1 : w0) + __shfl_xor_sync(FULLW, u8 ? w0 : w1, 8);
  z += __shfl_xor_sync(FULLW, z, 4);
  z += __shfl_xor_sync(FULLW, z, 2);
  z += __shfl_xor_sync(FULLW, z, 1);
  return z;
}
// 2 per-lane partial sums -> lane l holds the warp total of element l>>4
template <typename T>
__device__ __forceinline__ T reduce2(T v0, T v1, int lane) {
  const bool u16 = lane & 16;
  T z = (u16 ? v1 : v0) + __shfl_xor_sync(FULLW, u16 ? v0 : v1, 16);
  z += __shfl_xor_sync(FULLW, z, 8);
  z += __shfl_xor_sync(FULLW, z, 4);
  z += __shfl_xor_sync(FULLW, z, 2);
  z += __shfl_xor_sync(FULLW, z, 1);
  return z;
}

// Cholesky in place on the lower triangle of G (column-major, leading dimension ld), then zero the
// strict upper triangle.  Whole CTA participates.  Returns false if a pivot is not positive.
template <typename T>
__device__ bool block_cholesky(T *G, int k, int ld) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  bool ok = true;  // every thread reads the same pivots, so this is CTA-uniform
  for (int j = 0; j < k; ++j) {
    __syncthreads();
    const T piv = G[j + (size_t)j * ld];
    if (!(piv > T(0))) ok = false;
    const T d = sqrt(piv);
    __syncthreads();
    if (tid == 0) G[j + (size_t)j * ld] = d;
    const T dinv = T(1) / d;
    for (int i = j + 1 + tid; i < k; i += nt) G[i + (size_t)j * ld] *= dinv;
    __syncthreads();
    for (int c = j + 1 + warp; c < k; c += nw) {
      const T f = G[c + (size_t)j * ld];
      for (int i = c + lane; i < k; i += 32) G[i + (size_t)c * ld] -= G[i + (size_t)j * ld] * f;
    }
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += nt) {
    const int i = e % k, j = e / k;
    if (i < j) G[i + (size_t)j * ld] = T(0);
  }
  __syncthreads();
  return ok;
}

// One rotation round inside a warp's register tile.  NP = pairs rotated simultaneously (4 for a block
// pair, 2 inside a block); pair i rotates register columns PA[i], PB[i] (compile time) which are the
// matrix columns colA, colB of THIS lane's pair (the pair with index lane / (32/NP)).
template <typename T, int RPL, int NC, int NP, bool FASTROT>
struct RoundOp {
  template <typename PA, typename PB>
  static __device__ __forceinline__ void run(T (&x)[NC][RPL], PA pa, PB pb, int colA, int colB, T *nrm, int lane,
                                             T tol2, T stop2, int &rotated, int &big) {
    constexpr int GRP = 32 / NP;
    T part[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      T a = T(0);
#pragma unroll
      for (int m = 0; m < RPL; ++m) a = fma(x[pa(i)][m], x[pb(i)][m], a);
      part[i] = a;
    }
    T gamma;
    if (NP == 4)
      gamma = reduce4(part[0], part[1], part[NP > 2 ? 2 : 0], part[NP > 2 ? 3 : 0], lane);
    else
      gamma = reduce2(part[0], part[1], lane);
    const T alpha = nrm[colA], beta = nrm[colB];
    const T g2 = gamma * gamma, ab = alpha * beta;
    const bool rot = g2 > tol2 * ab;
    big |= g2 > stop2 * ab;
    rotated |= rot;
    T c, s, t;
    jacobi_rotation<T, FASTROT>(alpha, beta, gamma, rot, c, s, t);
    __syncwarp();
    if ((lane & (GRP - 1)) == 0) {
      nrm[colA] = fma(-t, gamma, alpha);
      nrm[colB] = fma(t, gamma, beta);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const T ci = __shfl_sync(FULLW, c, i * GRP), si = __shfl_sync(FULLW, s, i * GRP);
#pragma unroll
      for (int m = 0; m < RPL; ++m) {
        const T a = x[pa(i)][m], b = x[pb(i)][m];
        x[pa(i)][m] = fma(ci, a, -(si * b));
        x[pb(i)][m] = fma(si, a, ci * b);
      }
    }
  }
};

// ---- warp-level building blocks of a sweep ---------------------------------------------------------
// P*: first column of a 4-column sub-block (column stride ld*); ncol*: index of that column in the norm
// array.  The two sub-blocks of a pair may live in different memories (see block_jacobi_rb).

// the 6 pairs inside one sub-block, plus its exact squared column norms
template <typename T, int RPL, bool FASTROT>
__device__ __forceinline__ void sub_self(T *P, int ld, int ncol, const bool (&rowok)[RPL], T *nrm, int lane, T tol2,
                                         T stop2, int &rotated, int &big) {
  T x[4][RPL];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int m = 0; m < RPL; ++m) x[c][m] = rowok[m] ? P[(size_t)c * ld + lane + 32 * m] : T(0);
  {
    T sq[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T a = T(0);
#pragma unroll
      for (int m = 0; m < RPL; ++m) a = fma(x[c][m], x[c][m], a);
      sq[c] = a;
    }
    const T nn = reduce4(sq[0], sq[1], sq[2], sq[3], lane);
    if ((lane & 7) == 0) nrm[ncol + (lane >> 3)] = nn;
    __syncwarp();
  }
  const int h = lane >> 4;  // which of the 2 simultaneous pairs this lane works for
  // round 0: (0,1) (2,3)   round 1: (0,2) (1,3)   round 2: (0,3) (1,2)
  RoundOp<T, RPL, 4, 2, FASTROT>::run(
      x, [](int i) { return 2 * i; }, [](int i) { return 2 * i + 1; }, ncol + 2 * h, ncol + 2 * h + 1, nrm, lane, tol2,
      stop2, rotated, big);
  RoundOp<T, RPL, 4, 2, FASTROT>::run(
      x, [](int i) { return i; }, [](int i) { return i + 2; }, ncol + h, ncol + h + 2, nrm, lane, tol2, stop2, rotated, big);
  RoundOp<T, RPL, 4, 2, FASTROT>::run(
      x, [](int i) { return i; }, [](int i) { return 3 - i; }, ncol + h, ncol + 3 - h, nrm, lane, tol2, stop2, rotated, big);
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int m = 0; m < RPL; ++m)
      if (rowok[m]) P[(size_t)c * ld + lane + 32 * m] = x[c][m];
}

// the 16 cross pairs of two sub-blocks: 4 rounds of 4 disjoint pairs, all from registers
template <typename T, int RPL, bool FASTROT>
__device__ __forceinline__ void sub_cross(T *PA, int ldA, T *PB, int ldB, int ncolA, int ncolB,
                                          const bool (&rowok)[RPL], T *nrm, int lane, T tol2, T stop2, int &rotated,
                                          int &big) {
  T x[8][RPL];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      x[c][m] = rowok[m] ? PA[(size_t)c * ldA + lane + 32 * m] : T(0);
      x[4 + c][m] = rowok[m] ? PB[(size_t)c * ldB + lane + 32 * m] : T(0);
    }
  const int i = lane >> 3;  // this lane's pair within a round
  RoundOp<T, RPL, 8, 4, FASTROT>::run(
      x, [](int i) { return i; }, [](int i) { return 4 + i; }, ncolA + i, ncolB + i, nrm, lane, tol2, stop2, rotated, big);
  RoundOp<T, RPL, 8, 4, FASTROT>::run(
      x, [](int i) { return i; }, [](int i) { return 4 + ((i + 1) & 3); }, ncolA + i, ncolB + ((i + 1) & 3), nrm, lane,
      tol2, stop2, rotated, big);
  RoundOp<T, RPL, 8, 4, FASTROT>::run(
      x, [](int i) { return i; }, [](int i) { return 4 + ((i + 2) & 3); }, ncolA + i, ncolB + ((i + 2) & 3), nrm, lane,
      tol2, stop2, rotated, big);
  RoundOp<T, RPL, 8, 4, FASTROT>::run(
      x, [](int i) { return i; }, [](int i) { return 4 + ((i + 3) & 3); }, ncolA + i, ncolB + ((i + 3) & 3), nrm, lane,
      tol2, stop2, rotated, big);
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int m = 0; m < RPL; ++m)
      if (rowok[m]) {
        PA[(size_t)c * ldA + lane + 32 * m] = x[c][m];
        PB[(size_t)c * ldB + lane + 32 * m] = x[4 + c][m];
      }
}

// round-robin partner schedule on nbk (even) items: pair j of step os
__device__ __forceinline__ void rr_pair(int j, int os, int nbk, int &I, int &J) {
  const int nm1 = nbk - 1;
  if (j == 0) {
    I = nm1;
    J = os;
  } else {
    I = (os + j) % nm1;
    J = (os - j + nm1) % nm1;
  }
}

// One-sided Jacobi sweeps, register blocked.  G: kp columns (kp multiple of 8, columns >= k are zero) of ld
// rows (lanes only touch rows < nrows).  Columns [0, nres) are read and written in R (column stride ldr)
// instead: the caller keeps as many columns of a matrix that lives in global memory resident in shared
// memory as fit, which bounds the L2 working set (k = 256 FP64: 148 CTAs x 512 KB thrash the L2 into
// DRAM otherwise).  nrm: kp values of shared memory.  Returns the number of sweeps.
template <typename T, int RPL, bool FASTROT, bool MIXED = false>
__device__ int block_jacobi_rb(T *G, int k, int kp, int ld, int nrows, T *nrm, T stop2, T *R = nullptr, int ldr = 0,
                               int nres = 0) {
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int nb = kp / 4, npair = nb / 2, nm1 = nb - 1;
  const T tol2 = Fast<T>::tol2(k);
  bool rowok[RPL];
#pragma unroll
  for (int m = 0; m < RPL; ++m) rowok[m] = lane + 32 * m < nrows;
  // (MIXED is a template flag so that the all-shared-memory instantiation keeps plain LDS / STS addressing)
  auto colp = [&](int c) { return (MIXED && c < nres) ? R + (size_t)c * ldr : G + (size_t)c * ld; };
  auto cold = [&](int c) { return (MIXED && c < nres) ? ldr : ld; };
  int sweeps = 0;
  for (; sweeps < LK_JACOBI_CAP; ++sweeps) {
    int rotated = 0, big = 0;
    // pass 1: pairs inside each block of 4 columns, plus exact column norms
    for (int I = warp; I < nb; I += nw)
      sub_self<T, RPL, FASTROT>(colp(4 * I), cold(4 * I), 4 * I, rowok, nrm, lane, tol2, stop2, rotated, big);
    __syncthreads();
    // pass 2: all cross pairs of every block pair, round-robin over the blocks
    for (int os = 0; os < nm1; ++os) {
      for (int j = warp; j < npair; j += nw) {
        int I, J;
        rr_pair(j, os, nb, I, J);
        sub_cross<T, RPL, FASTROT>(colp(4 * I), cold(4 * I), colp(4 * J), cold(4 * J), 4 * I, 4 * J, rowok, nrm, lane,
                                   tol2, stop2, rotated, big);
      }
      __syncthreads();
    }
    const int any_rot = __syncthreads_or(rotated);
    const int any_big = __syncthreads_or(big);
    if (!any_rot || !any_big) return sweeps + 1;
  }
  return LK_JACOBI_CAP + 1;  // not converged: reported as an error by the caller (api.cu)
}

__device__ __forceinline__ void dmma884b(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// kp x kp product D = A B on the FP64 tensor pipe by the whole CTA (kp % 16 == 0, at most 4 blocks of
// 16x16 outputs per warp).  fa(i,l) -> A[i][l], fb(l,j) -> B[l][j] as double (0 outside k).  The result
// is stored column-major to Gout (leading dimension ld), which may alias an operand: all operands are
// read before the first store.
template <typename T, typename FA, typename FB>
__device__ __forceinline__ void block_gemm_dmma(FA fa, FB fb, T *Gout, int ld, int kp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int nb16 = kp / 16, nblk = nb16 * nb16;
  double acc[4][4][2];
#pragma unroll
  for (int bi = 0; bi < 4; ++bi) {
#pragma unroll
    for (int t = 0; t < 4; ++t) acc[bi][t][0] = acc[bi][t][1] = 0.0;
    const int blk = warp + bi * nw;
    if (blk < nblk) {
      const int br = 16 * (blk / nb16), bc = 16 * (blk % nb16);
#pragma unroll 2
      for (int kk = 0; kk < kp; kk += 4) {
        const double a0 = fa(br + lr, kk + lc), a1 = fa(br + 8 + lr, kk + lc);
        const double b0 = fb(kk + lc, bc + lr), b1 = fb(kk + lc, bc + 8 + lr);
        dmma884b(acc[bi][0][0], acc[bi][0][1], a0, b0);
        dmma884b(acc[bi][1][0], acc[bi][1][1], a0, b1);
        dmma884b(acc[bi][2][0], acc[bi][2][1], a1, b0);
        dmma884b(acc[bi][3][0], acc[bi][3][1], a1, b1);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int bi = 0; bi < 4; ++bi) {
    const int blk = warp + bi * nw;
    if (blk < nblk) {
      const int br = 16 * (blk / nb16), bc = 16 * (blk % nb16);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int row = br + 8 * (t >> 1) + lr, col = bc + 8 * (t & 1) + 2 * lc;
        Gout[row + (size_t)col * ld] = (T)acc[bi][t][0];
        Gout[row + (size_t)(col + 1) * ld] = (T)acc[bi][t][1];
      }
    }
  }
  __syncthreads();
}

// Out-of-place variant for matrices with more than 4 output blocks per warp (k % 32 == 0): each warp walks
// 32 x 32 output blocks (16 DMMA per 8 fragment loads) and stores them straight to Out, which must not
// alias the operands.
template <typename T, typename FA, typename FB>
__device__ __forceinline__ void block_gemm_dmma_oop(FA fa, FB fb, T *Out, int ldo, int kp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int nb32 = kp / 32;
  __syncthreads();
  for (int blk = warp; blk < nb32 * nb32; blk += nw) {
    const int br = 32 * (blk % nb32), bc = 32 * (blk / nb32);
    double acc[4][4][2];
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) acc[ti][tj][0] = acc[ti][tj][1] = 0.0;
#pragma unroll 4
    for (int kk = 0; kk < kp; kk += 4) {
      double a[4], b[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        a[t] = fa(br + 8 * t + lr, kk + lc);
        b[t] = fb(kk + lc, bc + 8 * t + lr);
      }
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) dmma884b(acc[ti][tj][0], acc[ti][tj][1], a[ti], b[tj]);
    }
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        const int row = br + 8 * ti + lr, col = bc + 8 * tj + 2 * lc;
        Out[row + (size_t)col * ldo] = (T)acc[ti][tj][0];
        Out[row + (size_t)(col + 1) * ldo] = (T)acc[ti][tj][1];
      }
  }
  __syncthreads();
}

// Cholesky of a matrix that lives in global memory / L2 (k % 32 == 0, ld = k, lower triangle valid), blocked
// left-looking over 32-column panels staged in shared memory (panel: 32 x (k + 4)):
//   panel = C[c0:k, c0:c0+32] - L[c0:k, 0:c0] L[c0:c0+32, 0:c0]^T   (FP64 tensor pipe, one 32 x 32 block per warp)
//   factor the panel in shared memory (column by column, as block_cholesky), write its lower part back.
// The unblocked algorithm on global memory pays an L2 round trip per column update (k^3/3 of them).
template <typename T>
__device__ bool block_cholesky_panel(T *G, int k, T *panel) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int ldp = k + 4;
  bool ok = true;
  for (int c0 = 0; c0 < k; c0 += 32) {
    const int rows = k - c0;
    __syncthreads();  // earlier panels are in global memory
    for (int rb = warp; rb < rows / 32; rb += nw) {
      double acc[4][4][2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) acc[ti][tj][0] = acc[ti][tj][1] = 0.0;
      const T *Arow = G + c0 + 32 * rb + lr;  // + 8 ti + (kk + lc) k
      const T *Brow = G + c0 + lr;            // + 8 tj + (kk + lc) k
#pragma unroll 2
      for (int kk = 0; kk < c0; kk += 4) {
        const size_t off = (size_t)(kk + lc) * k;
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          a[t] = (double)Arow[8 * t + off];
          b[t] = (double)Brow[8 * t + off];
        }
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
          for (int tj = 0; tj < 4; ++tj) dmma884b(acc[ti][tj][0], acc[ti][tj][1], a[ti], b[tj]);
      }
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) {
          const int r = 32 * rb + 8 * ti + lr, c = 8 * tj + 2 * lc;
          panel[r + (size_t)c * ldp] = (T)((double)G[c0 + r + (size_t)(c0 + c) * k] - acc[ti][tj][0]);
          panel[r + (size_t)(c + 1) * ldp] = (T)((double)G[c0 + r + (size_t)(c0 + c + 1) * k] - acc[ti][tj][1]);
        }
    }
    for (int jj = 0; jj < 32; ++jj) {
      __syncthreads();
      const T piv = panel[jj + (size_t)jj * ldp];
      if (!(piv > T(0))) ok = false;
      const T d = sqrt(piv);
      __syncthreads();
      if (tid == 0) panel[jj + (size_t)jj * ldp] = d;
      const T dinv = T(1) / d;
      for (int i = jj + 1 + tid; i < rows; i += nt) panel[i + (size_t)jj * ldp] *= dinv;
      __syncthreads();
      for (int c = jj + 1 + warp; c < 32; c += nw) {
        const T f = panel[c + (size_t)jj * ldp];
        for (int i = c + lane; i < rows; i += 32) panel[i + (size_t)c * ldp] -= panel[i + (size_t)jj * ldp] * f;
      }
    }
    __syncthreads();
    for (int e = tid; e < 32 * rows; e += nt) {
      const int i = e % rows, c = e / rows;
      if (i >= c) G[c0 + i + (size_t)(c0 + c) * k] = panel[i + (size_t)c * ldp];
    }
  }
  __syncthreads();
  for (int e = tid; e < k * k; e += nt) {
    const int i = e % k, j = e / k;
    if (i < j) G[i + (size_t)j * k] = T(0);
  }
  __syncthreads();
  return ok;
}

// mode 0: LETKF solve.  in: C (column-major lower triangle, SPD), b.  out: U (column-major, in place of C), lam, wbar.
// mode 1: ?syevd.      in: A (lower).                  out: W ascending, V.
template <typename T, int MODE, int RPL, bool SMEM>
__global__ void __launch_bounds__(RPL >= 5 ? 256 : 512)
    eig_blk_kernel(int k, int64_t n, int run, T *__restrict__ Cio, const T *__restrict__ bvec, T *__restrict__ lam,
                   T *__restrict__ wbar, const T *__restrict__ Ain, T *__restrict__ Wout, T *__restrict__ Vout,
                   int32_t *__restrict__ sweeps_max, int32_t *__restrict__ sweeps_sum, T *__restrict__ scratch,
                   int nres) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sm = reinterpret_cast<T *>(smem_raw);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int kp = (k + 7) & ~7;
  // shared-memory columns are padded by 4 rows: fragment loads of the warm-start products then spread
  // over the banks; in-place global storage requires k % 32 == 0 (checked on the host)
  const int ld = SMEM ? 32 * RPL + 4 : k;
  T *vec1 = sm;             // [kp]
  T *vec2 = sm + kp;        // [kp]
  T *nrm = sm + 2 * kp;     // [kp]
  T *Gs = sm + 3 * kp;      // [kp * ld] when SMEM, else the 32 x (k + 4) Cholesky panel / nres resident columns
  __shared__ T s_shift;
  __shared__ T red_lo[512], red_sc[512];
  // Warm start (MODE 0, shared-memory path, kp % 16 == 0, <= 4 output blocks per warp): the CTA walks
  // `run` consecutive units and solves each in the eigenbasis of the previous one, exactly as the k = 32
  // chained kernel does; U_prev is read back from global memory (L2), the three products run on the
  // FP64 tensor pipe.
  // Larger matrices (k % 32 == 0, `scratch` given): the same, with out-of-place products through a
  // k x k scratch matrix per CTA in global memory (L2); the grid is then persistent and strides over runs.
  const bool chain_small = MODE == 0 && SMEM && run > 1 && kp % 16 == 0 && (kp / 16) * (kp / 16) <= 4 * nw;
  const bool chain_big = MODE == 0 && !chain_small && run > 1 && scratch != nullptr && k % 32 == 0;
  const bool can_chain = chain_small || chain_big;
  T *S = chain_big ? scratch + (size_t)blockIdx.x * k * k : nullptr;

  if (SMEM) {
    for (int e = tid; e < kp * ld; e += nt) Gs[e] = T(0);
    __syncthreads();
  }
  for (int64_t rb = blockIdx.x; rb * run < n; rb += gridDim.x) {
    const int64_t u0 = rb * run, u1 = u0 + run < n ? u0 + run : n;
    bool prev_ok = false;
    for (int64_t u = u0; u < u1; ++u) {
      T *Gg = MODE == 0 ? Cio + u * (int64_t)k * k : Vout + u * (int64_t)k * k;
      T *G = SMEM ? Gs : Gg;
      const bool warm = can_chain && prev_ok;
      const T *Up = MODE == 0 && u > u0 ? Cio + (u - 1) * (int64_t)k * k : nullptr;  // U of the previous unit
      if (MODE == 0) {
        __syncthreads();
        if (SMEM)  // Gram kernels deliver the column-major lower triangle: symmetrise while loading
          for (int e = tid; e < k * k; e += nt) {
            const int i = e % k, j = e / k;
            G[i + (size_t)j * ld] = i >= j ? Gg[i + (size_t)j * k] : Gg[j + (size_t)i * k];
          }
        if (tid == 0) s_shift = T(0);
        __syncthreads();
        if (warm && chain_big) {
          // S = C U_prev, then C' = U_prev^T S -> G (C: full in shared memory, lower triangle in global)
          block_gemm_dmma_oop<T>(
              [&](int i, int l) {
                return SMEM ? (double)G[i + (size_t)l * ld] : (double)(i >= l ? G[i + (size_t)l * ld] : G[l + (size_t)i * ld]);
              },
              [&](int l, int j) { return (double)Up[l + (size_t)j * k]; }, S, k, k);
          block_gemm_dmma_oop<T>([&](int i, int l) { return (double)Up[l + (size_t)i * k]; },
                                 [&](int l, int j) { return (double)S[l + (size_t)j * k]; }, G, ld, k);
        } else if (warm) {
          // T = C U_prev, then C' = U_prev^T T (both written over G)
          block_gemm_dmma<T>([&](int i, int l) { return (double)G[i + (size_t)l * ld]; },
                             [&](int l, int j) { return (l < k && j < k) ? (double)Up[l + (size_t)j * k] : 0.0; }, G, ld, kp);
          block_gemm_dmma<T>([&](int i, int l) { return (l < k && i < k) ? (double)Up[l + (size_t)i * k] : 0.0; },
                             [&](int l, int j) { return (double)G[l + (size_t)j * ld]; }, G, ld, kp);
        }
        // C is SPD by construction; a NaN input propagates
        if (SMEM)
          block_cholesky(G, k, ld);
        else
          block_cholesky_panel(G, k, Gs);
      } else {
        const T *A = Ain + u * (int64_t)k * k;
        // First try the matrix as it is (an SPD input keeps its full relative accuracy); if a pivot
        // fails, shift by a Gershgorin bound so that it becomes definite and factor again.
        for (int attempt = 0; attempt < 2; ++attempt) {
          __syncthreads();
          for (int e = tid; e < k * k; e += nt) {  // symmetrise from the lower triangle
            const int i = e % k, j = e / k;
            G[i + (size_t)j * ld] = i >= j ? A[i + (size_t)j * k] : A[j + (size_t)i * k];
          }
          if (tid == 0 && attempt == 0) s_shift = T(0);
          __syncthreads();
          if (attempt == 1) {
            T lowest = sizeof(T) == 8 ? T(1e300) : T(3e38);
            T scale = T(0);
            for (int i = tid; i < k; i += nt) {
              T off = 0;
              for (int j = 0; j < k; ++j)
                if (j != i) off += fabs(G[i + (size_t)j * ld]);
              lowest = min(lowest, G[i + (size_t)i * ld] - off);
              scale = max(scale, fabs(G[i + (size_t)i * ld]) + off);
            }
            red_lo[tid] = lowest;
            red_sc[tid] = scale;
            __syncthreads();
            if (tid == 0) {
              T lo = red_lo[0], sc = red_sc[0];
              for (int i = 1; i < nt; ++i) {
                lo = min(lo, red_lo[i]);
                sc = max(sc, red_sc[i]);
              }
              s_shift = sc * T(1e-3) - min(lo, T(0));
            }
            __syncthreads();
            const T sh = s_shift;
            for (int i = tid; i < k; i += nt) G[i + (size_t)i * ld] += sh;
            __syncthreads();
          }
          if (SMEM ? block_cholesky(G, k, ld) : block_cholesky_panel(G, k, Gs)) break;
        }
      }

      // MODE 0 feeds the LETKF weights (1e-10 bar): stop once a sweep saw only |cos| <= 1e-7 and take the
      // rotation angle in real32.  MODE 1 is the general eigensolver: |cos| <= 1e-9, angle in working precision.
      const T stop2 = MODE == 0 ? T(1e-14) : (sizeof(T) == 8 ? T(1e-18) : T(1e-9));
      int sweeps;
      if (SMEM) {
        sweeps = block_jacobi_rb<T, RPL, MODE == 0>(G, k, kp, ld, k, nrm, stop2);
      } else {
        // the first nres columns stay in shared memory for the sweeps (over the Cholesky panel, now free)
        for (int e = tid; e < nres * k; e += nt) Gs[e] = G[e];
        __syncthreads();
        sweeps = block_jacobi_rb<T, RPL, MODE == 0, true>(G, k, kp, ld, k, nrm, stop2, Gs, k, nres);
        for (int e = tid; e < nres * k; e += nt) G[e] = Gs[e];
        __syncthreads();
      }
      if (tid == 0 && sweeps_max) atomicMax(sweeps_max, sweeps);
      if (tid == 0 && sweeps_sum) atomicAdd(sweeps_sum, sweeps);

      // column norms -> eigenvalues; normalise columns -> eigenvectors
      for (int j = warp; j < k; j += nw) {
        T *gj = G + (size_t)j * ld;
        T a = 0;
        for (int i = lane; i < k; i += 32) a += gj[i] * gj[i];
        a = warp_sum(a);
        const T inv = T(1) / sqrt(a);
        for (int i = lane; i < k; i += 32) gj[i] *= inv;
        if (lane == 0) vec1[j] = a;  // lambda (+ shift)
      }
      __syncthreads();

      if (MODE == 0) {
        if (warm && chain_big) {
          // U = U_prev U'
          block_gemm_dmma_oop<T>([&](int i, int l) { return (double)Up[i + (size_t)l * k]; },
                                 [&](int l, int j) { return (double)G[l + (size_t)j * ld]; }, S, k, k);
          for (int e = tid; e < k * k; e += nt) G[(e % k) + (size_t)(e / k) * ld] = S[e];
          __syncthreads();
        } else if (warm) {
          // U = U_prev U'
          block_gemm_dmma<T>([&](int i, int l) { return (i < k && l < k) ? (double)Up[i + (size_t)l * k] : 0.0; },
                             [&](int l, int j) { return (double)G[l + (size_t)j * ld]; }, G, ld, kp);
        }
        {
          // a unit whose matrix is not finite / not positive must not seed its neighbour
          int bad = 0;
          for (int j = tid; j < k; j += nt) bad |= !(vec1[j] > T(0)) || !(vec1[j] < T(1e30));
          prev_ok = !__syncthreads_or(bad);
        }
        // wbar = U diag(1/lam) U^T b   (inverse_matrix + ?gemv + ?symv, eig:37-76, core:651-652)
        const T *b = bvec + u * (int64_t)k;
        for (int j = warp; j < k; j += nw) {
          const T *gj = G + (size_t)j * ld;
          T a = 0;
          for (int i = lane; i < k; i += 32) a += gj[i] * b[i];
          a = warp_sum(a);
          if (lane == 0) vec2[j] = a / vec1[j];
        }
        __syncthreads();
        for (int i = tid; i < k; i += nt) {
          T a = 0;
          for (int j = 0; j < k; ++j) a += G[i + (size_t)j * ld] * vec2[j];
          wbar[u * (int64_t)k + i] = a;
          lam[u * (int64_t)k + i] = vec1[i];
        }
        if (SMEM)
          for (int e = tid; e < k * k; e += nt) Gg[e] = G[(e % k) + (size_t)(e / k) * ld];
      } else {
        // ascending order like LAPACK: rank by counting
        const T sh = s_shift;
        int *rank = reinterpret_cast<int *>(vec2);
        for (int j = tid; j < k; j += nt) {
          const T lj = vec1[j];
          int r = 0;
          for (int l = 0; l < k; ++l) {
            const T ll = vec1[l];
            r += (ll < lj) || (ll == lj && l < j);
          }
          rank[j] = r;
          Wout[u * (int64_t)k + r] = lj - sh;
        }
        __syncthreads();
        if (SMEM) {
          for (int e = tid; e < k * k; e += nt) {
            const int i = e % k, j = e / k;
            Gg[i + (size_t)rank[j] * k] = G[i + (size_t)j * ld];
          }
        } else {
          // in-place column permutation in global memory: every thread owns rows and walks the cycles
          for (int i = tid; i < k; i += nt) {
            for (int start = 0; start < k; ++start) {
              int c = rank[start];
              bool smallest = true;
              while (c != start) {
                if (c < start) {
                  smallest = false;
                  break;
                }
                c = rank[c];
              }
              if (!smallest) continue;
              T carry = G[i + (size_t)start * k];
              int dst = rank[start];
              while (dst != start) {
                const T tmp = G[i + (size_t)dst * k];
                G[i + (size_t)dst * k] = carry;
                carry = tmp;
                dst = rank[dst];
              }
              G[i + (size_t)start * k] = carry;
            }
          }
        }
      }
      __syncthreads();
    }  // unit loop
  }  // run loop
}

// warm-start scratch of the persistent large-k kernel: grow-only, owned by the calling context
static unsigned char *eig_scratch(size_t bytes) {
  CtxShared *c = current_ctx_shared();
  LK_REQUIRE(c != nullptr, "eigensolver launched outside a library context");
  c->eig_scratch.ensure(bytes);
  return c->eig_scratch.p;
}

template <typename T, int MODE, int RPL>
static void launch_rpl(cudaStream_t s, int k, int64_t n, T *Cio, const T *b, T *lam, T *wbar, const T *A, T *W,
                       T *V, int32_t *sweeps_max) {
  const int kp = (k + 7) & ~7;
  const size_t smem_full = sizeof(T) * (3 * (size_t)kp + (size_t)kp * (32 * RPL + 4));
  const bool smem_ok = smem_full <= 216 * 1024;
  LK_REQUIRE(smem_ok || k % 32 == 0,
             "eigensolver: k above the shared-memory limit (160 FP64 / 224 FP32) must be a multiple of 32");
  // in-global variant: Cholesky panel, then (FP64) as many resident columns as fit beside ~10 KB of static
  // arrays.  FP32 matrices (256 KB at k = 256) stay in L2 and run two CTAs per SM instead.
  const int nres =
      (smem_ok || sizeof(T) == 4) ? 0 : std::min(k, (int)((216 * 1024 - sizeof(T) * 3 * kp) / (sizeof(T) * k)) & ~3);
  const size_t smem = smem_ok ? smem_full
                              : sizeof(T) * (3 * (size_t)kp + std::max(32 * ((size_t)k + 4), (size_t)nres * k));
  int nwarps = std::max(1, std::min(kp / 8, RPL >= 5 ? 8 : 16));
  const int threads = 32 * nwarps;
  static const int chain = [] {
    const char *e = getenv("LETKF_B200_EIG_CHAIN");
    return e ? atoi(e) : 8;
  }();
  // units per CTA: a run of neighbouring grid points solved with warm starts (MODE 0 only)
  const bool chain_small = MODE == 0 && smem_ok && chain > 1 && kp % 16 == 0 && (kp / 16) * (kp / 16) <= 4 * nwarps;
  const bool chain_big = MODE == 0 && !chain_small && chain > 1 && k % 32 == 0;
  const int run = chain_big ? 16 : chain_small ? 8 : 1;
  int64_t nblocks = (n + run - 1) / run;
  int32_t *ssum = sweeps_max ? sweeps_max + 1 : nullptr;
  auto launch = [&](auto kern) {
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    T *scratch = nullptr;
    if (chain_big) {  // persistent grid: one k x k scratch matrix per resident CTA
      int dev = 0, sms = 0, occ = 0;
      LK_CUDA(cudaGetDevice(&dev));
      LK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      LK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
      nblocks = std::min<int64_t>(nblocks, (int64_t)sms * std::max(occ, 1));
      scratch = reinterpret_cast<T *>(eig_scratch((size_t)nblocks * k * k * sizeof(T)));
    }
    LK_REQUIRE(nblocks < ((int64_t)1 << 31), "eigensolver: batch too large for one launch");
    kern<<<(unsigned)nblocks, threads, smem, s>>>(k, n, run, Cio, b, lam, wbar, A, W, V, sweeps_max, ssum, scratch,
                                                  nres);
  };
  if (smem_ok)
    launch(eig_blk_kernel<T, MODE, RPL, true>);
  else
    launch(eig_blk_kernel<T, MODE, RPL, false>);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

template <typename T, int MODE>
static void launch_eig_generic(cudaStream_t s, int k, int64_t n, T *Cio, const T *b, T *lam, T *wbar, const T *A,
                               T *W, T *V, int32_t *sweeps_max) {
  if (n == 0) return;
  LK_REQUIRE(k >= 2 && k <= LETKF_B200_MAX_MEMBERS, "eigensolver: need 2 <= k <= 256");
  const int rpl = (k + 31) / 32;
  switch (rpl) {
    case 1: launch_rpl<T, MODE, 1>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    case 2: launch_rpl<T, MODE, 2>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    case 3: launch_rpl<T, MODE, 3>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    case 4: launch_rpl<T, MODE, 4>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    case 5: launch_rpl<T, MODE, 5>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    case 6: launch_rpl<T, MODE, 6>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    case 7: launch_rpl<T, MODE, 7>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
    default: launch_rpl<T, MODE, 8>(s, k, n, Cio, b, lam, wbar, A, W, V, sweeps_max); break;
  }
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
