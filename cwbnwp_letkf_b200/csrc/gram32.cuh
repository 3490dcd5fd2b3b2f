// k = 32 localise + Gram for one analysis unit on the FP64 tensor pipe (see kernels_k32.cu for the
// description).  Shared by gram32_dmma_kernel (result to global memory) and by the fully fused
// per-unit kernel (result to shared memory, kernels_fused32.cu).
#pragma once

#include "letkf_internal.cuh"

namespace lk {

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Accumulates C = Yb Yb^T + mu I (full symmetric, element (r,c) stored at Cout[r*ldc + c]) and
// b = Yb yo (bout[0..31]) for grid point q of the current chunk.  Returns true if a NaN row was seen.
// One warp; Cout / bout may be shared or global memory.
__device__ __forceinline__ bool gram32_unit(const TreeViews &tv, int64_t q, double mu, double *Cout, int ldc,
                                            double *bout, int lane) {
  constexpr unsigned FULLG = 0xffffffffu;
  const int lr = lane >> 2, lc = lane & 3;
  double acc[10][2];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t][0] = acc[t][1] = 0.0;
  double bacc[4] = {0.0, 0.0, 0.0, 0.0};
  bool anynan = false;

  for (int t = 0; t < tv.ntrees; ++t) {
    const TreeView &TV = tv.t[t];
    const int ncand = TV.cnt[q] * TV.nact;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
      // candidate (tree entry, slot) c0 + lane -> row metadata
      bool pass = false;
      float ei = 0.f, yo = 0.f;
      const float *pr = TV.pert;
      const int c = c0 + lane;
      if (c < ncand) {
        const int j = c / TV.nact, a = c - j * TV.nact;
        const int64_t o = (int64_t)(TV.idx[q * TV.nalloc + j] - 1) * TV.nvar + TV.act[a];
        if (TV.pass[o]) {
          pass = true;
          ei = lk_error_inv(TV.err[o], TV.r2[q * TV.nalloc + j], tv.weight_function);
          yo = LK_MUL(TV.omm[o], ei);
          pr = TV.pert + o * 32;
          anynan = anynan || (ei != ei);
        }
      }
      const unsigned pmask = __ballot_sync(FULLG, pass);
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) {
        if (((pmask >> (4 * g8)) & 0xFu) == 0u) continue;  // warp-uniform
        const int src = 4 * g8 + lc;
        const float e = __shfl_sync(FULLG, ei, src);
        const float y = __shfl_sync(FULLG, yo, src);
        const float *p = (const float *)__shfl_sync(FULLG, (unsigned long long)pr, src);
        const bool ok = (pmask >> src) & 1u;
        double a[4];
#pragma unroll
        for (int I = 0; I < 4; ++I) {
          const float v = ok ? LK_MUL(__ldg(p + 8 * I + lr), e) : 0.f;
          a[I] = (double)v;
        }
        const double yd = (double)y;
#pragma unroll
        for (int I = 0; I < 4; ++I) bacc[I] = fma(a[I], yd, bacc[I]);
        int tix = 0;
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
          for (int J = 0; J <= I; ++J) {
            dmma884(acc[tix][0], acc[tix][1], a[I], a[J]);
            ++tix;
          }
      }
    }
  }
  {
    int tix = 0;
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = 0; J <= I; ++J) {
        const int row = 8 * I + lr, col = 8 * J + 2 * lc;
        double2 v;
        v.x = acc[tix][0] + (row == col ? mu : 0.0);
        v.y = acc[tix][1] + (row == col + 1 ? mu : 0.0);
        *reinterpret_cast<double2 *>(Cout + row * ldc + col) = v;
        if (I != J) {  // mirror: the warm-started eigensolver multiplies with full rows of C
          Cout[col * ldc + row] = v.x;
          Cout[(col + 1) * ldc + row] = v.y;
        }
        ++tix;
      }
  }
#pragma unroll
  for (int I = 0; I < 4; ++I) {
    double s = bacc[I];
    s += __shfl_xor_sync(FULLG, s, 1);
    s += __shfl_xor_sync(FULLG, s, 2);
    if (lc == 0) bout[8 * I + lr] = s;
  }
  return __any_sync(FULLG, anynan);
}

}  // namespace lk
