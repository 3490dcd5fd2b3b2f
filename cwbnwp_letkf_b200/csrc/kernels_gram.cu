// Fused R-localisation + local Gram matrix:
//   C = Yb Yb^T + (k-1)/rho I   (?syrk, module_letkf_core.f90:649/656)
//   b = Yb yo                    (?gemv, module_letkf_core.f90:651/658)
// with the rows of Yb / yo produced on the fly from the per-ob perturbations exactly as
// letkf_yoyb does per grid point (module_letkf_core.f90:443-452,516-525): error_inv from the
// localisation function in real32, then yb = bg*error_inv and yo = omm*error_inv as single
// real32 roundings, only then promoted to the working precision T.
//
// Generic path (any k <= 256): CTAs of 16x16 threads; CTA (unit, panel pair) computes one
// (16 KT)^2 block of C (one block for k <= 128, 2x2 blocks for k = 256).  Rows are staged 32
// at a time in shared memory (coalesced 4*k-byte reads of the ob-major perturbation table);
// thread (ty,tx) owns the interleaved KT x KT register tile C[pi+ty+16a][pj+tx+16b], so the
// shared reads are one broadcast and one conflict-free vector per step.  Candidates that fail QC are
// staged as zero rows: they add exact zeros, so C and b equal the reference's sums over the
// surviving rows.
#include "letkf_internal.cuh"

namespace lk {

constexpr int kRows = 32;

struct RowMeta {
  const float *pert;
  float ei;
  float yo;
  int pass;
};

template <typename T, int KT>
__global__ void __launch_bounds__(256)
    gram_kernel(TreeViews tv, int k, int64_t nunits, const int32_t *__restrict__ unit_pt, T mu,
                T *__restrict__ C, T *__restrict__ bvec, int32_t *__restrict__ nanflag) {
  constexpr int PW = 16 * KT;            // panel width
  const int KP = (k + 15) & ~15;         // staged row length
  const int nb = (k + PW - 1) / PW;
  const int pi = (blockIdx.y / nb) * PW, pj = (blockIdx.y % nb) * PW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *yb_s = reinterpret_cast<T *>(smem_raw);                 // [kRows][KP]
  T *yo_s = yb_s + kRows * KP;                               // [kRows]
  RowMeta *meta = reinterpret_cast<RowMeta *>(yo_s + kRows); // [kRows]
  __shared__ int s_nan;

  const int64_t unit = blockIdx.x;
  if (unit >= nunits) return;
  const int64_t q = unit_pt[unit];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_nan = 0;

  T acc[KT][KT];
#pragma unroll
  for (int a = 0; a < KT; ++a)
#pragma unroll
    for (int b = 0; b < KT; ++b) acc[a][b] = T(0);
  T bacc[KT];
#pragma unroll
  for (int b = 0; b < KT; ++b) bacc[b] = T(0);

  for (int t = 0; t < tv.ntrees; ++t) {
    const TreeView &TV = tv.t[t];
    const int ncand = TV.cnt[q] * TV.nact;
    for (int c0 = 0; c0 < ncand; c0 += kRows) {
      __syncthreads();  // previous batch fully consumed
      if (tid < kRows) {
        RowMeta m;
        m.pert = nullptr;
        m.ei = 0.f;
        m.yo = 0.f;
        m.pass = 0;
        const int c = c0 + tid;
        if (c < ncand) {
          const int j = c / TV.nact, a = c - j * TV.nact;
          const int64_t o = (int64_t)(TV.idx[q * TV.nalloc + j] - 1) * TV.nvar + TV.act[a];
          if (TV.pass[o]) {
            m.pass = 1;
            m.ei = lk_error_inv(TV.err[o], TV.r2[q * TV.nalloc + j], tv.weight_function);
            m.yo = LK_MUL(TV.omm[o], m.ei);
            m.pert = TV.pert + o * k;
            if (m.ei != m.ei) s_nan = 1;  // sqrt of a negative Gaspari-Cohn value
          }
        }
        meta[tid] = m;
        yo_s[tid] = (T)m.yo;
      }
      __syncthreads();
      for (int r = warp; r < kRows; r += 8) {
        const RowMeta m = meta[r];
        for (int i = lane; i < KP; i += 32) {
          float v = 0.f;
          if (m.pass && i < k) v = LK_MUL(__ldg(m.pert + i), m.ei);
          yb_s[r * KP + i] = (T)v;
        }
      }
      __syncthreads();
#pragma unroll 4
      for (int r = 0; r < kRows; ++r) {
        T ra[KT], rb[KT];
#pragma unroll
        for (int a = 0; a < KT; ++a) ra[a] = (pi + ty + 16 * a < KP) ? yb_s[r * KP + pi + ty + 16 * a] : T(0);
#pragma unroll
        for (int b = 0; b < KT; ++b) rb[b] = (pj + tx + 16 * b < KP) ? yb_s[r * KP + pj + tx + 16 * b] : T(0);
#pragma unroll
        for (int a = 0; a < KT; ++a)
#pragma unroll
          for (int b = 0; b < KT; ++b) acc[a][b] += ra[a] * rb[b];
        if (ty == 0 && pi == 0) {
          const T y = yo_s[r];
#pragma unroll
          for (int b = 0; b < KT; ++b) bacc[b] += rb[b] * y;
        }
      }
    }
  }
  __syncthreads();
  T *Cu = C + unit * (int64_t)k * k;
#pragma unroll
  for (int a = 0; a < KT; ++a)
#pragma unroll
    for (int b = 0; b < KT; ++b) {
      const int i = pi + ty + 16 * a, j = pj + tx + 16 * b;
      if (i < k && j < k) Cu[(int64_t)i * k + j] = acc[a][b] + (i == j ? mu : T(0));
    }
  if (ty == 0 && pi == 0) {
#pragma unroll
    for (int b = 0; b < KT; ++b) {
      const int j = pj + tx + 16 * b;
      if (j < k) bvec[unit * (int64_t)k + j] = bacc[b];
    }
  }
  if (tid == 0 && blockIdx.y == 0) nanflag[unit] = s_nan;
}

template <typename T>
void launch_gram(cudaStream_t s, const TreeViews &tv, int k, int64_t nunits, const int32_t *unit_pt, T mu,
                 T *C, T *b, int32_t *nanflag) {
  if (nunits == 0) return;
  const int kt = (k + 15) / 16;
  LK_REQUIRE(k <= LETKF_B200_MAX_MEMBERS, "launch_gram: nmember > 256");
  const int KP = (k + 15) & ~15;
  auto go = [&](auto ktc) {
    constexpr int KT = decltype(ktc)::value;
    const int nb = (k + 16 * KT - 1) / (16 * KT);
    const size_t smem = (size_t)kRows * KP * sizeof(T) + kRows * sizeof(T) + kRows * sizeof(RowMeta);
    auto kern = gram_kernel<T, KT>;
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t u0 = 0; u0 < nunits; u0 += 1 << 30) {  // grid.x limit is 2^31-1; chunks are far smaller
      const int64_t nu = std::min<int64_t>(nunits - u0, 1 << 30);
      kern<<<dim3((unsigned)nu, (unsigned)(nb * nb)), 256, smem, s>>>(tv, k, nu, unit_pt + u0, mu,
                                                                    C + u0 * (int64_t)k * k, b + u0 * k,
                                                                    nanflag + u0);
    }
  };
  if (kt <= 1) go(std::integral_constant<int, 1>{});
  else if (kt <= 2) go(std::integral_constant<int, 2>{});
  else if (kt <= 4) go(std::integral_constant<int, 4>{});
  else if (kt <= 6) go(std::integral_constant<int, 6>{});
  else go(std::integral_constant<int, 8>{});
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

template void launch_gram<double>(cudaStream_t, const TreeViews &, int, int64_t, const int32_t *, double,
                                  double *, double *, int32_t *);
template void launch_gram<float>(cudaStream_t, const TreeViews &, int, int64_t, const int32_t *, float,
                                 float *, float *, int32_t *);

}  // namespace lk
