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
  if (unit >= nunits || (tv.nunits_dev && unit >= *tv.nunits_dev)) return;
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

// =====================================================================================================
// Tensor-core Gram for any k (FP64 accumulation on mma.sync.m8n8k4.f64), used for every k != 32.
//
// C is tiled in 8x8 MMA tiles grouped into 32x32 super-blocks; only the S(S+1)/2 lower super-blocks
// (S = ceil(k/32)) are computed.  A warp owns up to two super-blocks (16 tiles = 32 accumulator doubles
// each); a CTA of up to 8 warps owns up to 16 of them and a unit is spread over as many CTAs as
// needed (k = 256: 36 super-blocks -> 3 CTAs).  Rows are staged 32 at a time as doubles in shared
// memory with a row stride = 8 (mod 16) doubles, so that the fragment load of a 4-row group
// (lane -> row lane%4, member 8I + lane/4) costs the minimal two wavefronts.  The same fragment is the
// A operand of tile row I and the B operand of tile column I.
// =====================================================================================================
namespace lk {

__device__ __forceinline__ void dmma884g(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <typename T>
__global__ void __launch_bounds__(256)
    gram_dmma_kernel(TreeViews tv, int k, int S, int sb_per_cta, int64_t nunits, const int32_t *__restrict__ unit_pt,
                     double mu, T *__restrict__ C, T *__restrict__ bvec, int32_t *__restrict__ nanflag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int KP = 32 * S;     // padded member count
  const int LDS_ = KP + 8;   // row stride in doubles: 8 (mod 16)
  double *yb_s = reinterpret_cast<double *>(smem_raw);  // [kRows][LDS_]
  double *yo_s = yb_s + kRows * LDS_;                   // [kRows]
  RowMeta *meta = reinterpret_cast<RowMeta *>(yo_s + kRows);
  __shared__ int s_nan;
  __shared__ unsigned s_pmask;

  const int64_t unit = blockIdx.x;
  if (unit >= nunits || (tv.nunits_dev && unit >= *tv.nunits_dev)) return;
  const int64_t q = unit_pt[unit];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int NSB = S * (S + 1) / 2;
  if (tid == 0) s_nan = 0;

  // this warp's super-blocks (si >= sj), linear index over the lower triangle
  int sbi[2], sbj[2];
  bool has[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const int lin = blockIdx.y * sb_per_cta + warp * 2 + s;
    has[s] = (warp * 2 + s) < sb_per_cta && lin < NSB;
    int i = 0, rem = has[s] ? lin : 0;
    while (rem > i) {
      rem -= i + 1;
      ++i;
    }
    sbi[s] = i;
    sbj[s] = rem;
  }
  double acc[2][16][2];
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[s][t][0] = acc[s][t][1] = 0.0;
  double bacc = 0.0;  // b[tid] for the CTA with blockIdx.y == 0

  for (int t = 0; t < tv.ntrees; ++t) {
    const TreeView &TV = tv.t[t];
    const int ncand = TV.cnt[q] * TV.nact;
    for (int c0 = 0; c0 < ncand; c0 += kRows) {
      __syncthreads();
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
            if (m.ei != m.ei) s_nan = 1;
          }
        }
        meta[tid] = m;
        yo_s[tid] = (double)m.yo;
        const unsigned pm = __ballot_sync(0xffffffffu, m.pass != 0);
        if (tid == 0) s_pmask = pm;
      }
      __syncthreads();
      const unsigned pmask = s_pmask;
      for (int r = warp; r < kRows; r += nw) {
        const RowMeta m = meta[r];
        for (int i = lane; i < KP; i += 32) {
          float v = 0.f;
          if (m.pass && i < k) v = LK_MUL(__ldg(m.pert + i), m.ei);
          yb_s[r * LDS_ + i] = (double)v;
        }
      }
      __syncthreads();
      if (blockIdx.y == 0 && tid < k) {
#pragma unroll 8
        for (int r = 0; r < kRows; ++r) bacc = fma(yb_s[r * LDS_ + tid], yo_s[r], bacc);
      }
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!has[s]) continue;
        const bool diag = sbi[s] == sbj[s];
#pragma unroll 2
        for (int g8 = 0; g8 < 8; ++g8) {
          if (((pmask >> (4 * g8)) & 0xFu) == 0u) continue;
          const double *row = yb_s + (4 * g8 + lc) * LDS_ + lr;
          double fa[4], fb[4];
#pragma unroll
          for (int I = 0; I < 4; ++I) fa[I] = row[32 * sbi[s] + 8 * I];
          if (diag) {
#pragma unroll
            for (int I = 0; I < 4; ++I) fb[I] = fa[I];
          } else {
#pragma unroll
            for (int I = 0; I < 4; ++I) fb[I] = row[32 * sbj[s] + 8 * I];
          }
#pragma unroll
          for (int I = 0; I < 4; ++I)
#pragma unroll
            for (int J = 0; J < 4; ++J) {
              if (diag && J > I) continue;  // upper tiles of a diagonal super-block are not needed
              dmma884g(acc[s][I * 4 + J][0], acc[s][I * 4 + J][1], fa[I], fb[J]);
            }
        }
      }
    }
  }
  __syncthreads();
  T *Cu = C + unit * (int64_t)k * k;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    if (!has[s]) continue;
    const bool diag = sbi[s] == sbj[s];
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = 0; J < 4; ++J) {
        if (diag && J > I) continue;
        const int row = 32 * sbi[s] + 8 * I + lr, col = 32 * sbj[s] + 8 * J + 2 * lc;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = col + e;
          if (row < k && cc < k) {
            const double v = acc[s][I * 4 + J][e] + (row == cc ? mu : 0.0);
            // element (row, cc), row >= cc except inside diagonal tiles: store where the column-major
            // lower triangle lives (and its mirror, so diagonal tiles are complete)
            Cu[(int64_t)cc * k + row] = (T)v;
            if (diag && I == J) Cu[(int64_t)row * k + cc] = (T)v;
          }
        }
      }
  }
  if (blockIdx.y == 0) {
    if (tid < k) bvec[unit * (int64_t)k + tid] = (T)bacc;
    if (tid == 0) nanflag[unit] = s_nan;
  }
}

template <typename T>
void launch_gram_dmma(cudaStream_t s, const TreeViews &tv, int k, int64_t nunits, const int32_t *unit_pt, T mu,
                      T *C, T *b, int32_t *nanflag) {
  if (nunits == 0) return;
  LK_REQUIRE(k <= LETKF_B200_MAX_MEMBERS, "launch_gram_dmma: nmember > 256");
  const int S = (k + 31) / 32;
  const int NSB = S * (S + 1) / 2;
  const int ncta = (NSB + 15) / 16;
  const int sb_per_cta = (NSB + ncta - 1) / ncta;
  int nwarps = (sb_per_cta + 1) / 2;
  nwarps = std::max(nwarps, (k + 31) / 32);  // the b-vector needs k threads in CTA 0
  const size_t smem = sizeof(double) * ((size_t)kRows * (32 * S + 8) + kRows) + kRows * sizeof(RowMeta);
  auto kern = gram_dmma_kernel<T>;
  LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<dim3((unsigned)nunits, (unsigned)ncta), 32 * nwarps, smem, s>>>(tv, k, S, sb_per_cta, nunits, unit_pt,
                                                                          (double)mu, C, b, nanflag);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_gram_dmma<double>(cudaStream_t, const TreeViews &, int, int64_t, const int32_t *, double,
                                       double *, double *, int32_t *);
template void launch_gram_dmma<float>(cudaStream_t, const TreeViews &, int, int64_t, const int32_t *, float, float *,
                                      float *, int32_t *);

}  // namespace lk

// =====================================================================================================
// TMA-staged variant of the tensor-core Gram (k % 4 == 0).  The rows of a batch are 4k-byte contiguous
// lines of the ob-major perturbation table; a producer warp gathers them with 1-D bulk asynchronous
// copies (cp.async.bulk.shared.global, completion on an mbarrier) into a ring of shared-memory stages,
// so the gathers of the next batches are in flight while the current one feeds the tensor pipe.  Rows are kept as
// the raw real32 perturbations (row stride 4k+32 bytes: the 4 rows of an MMA fragment fall in distinct
// bank groups); yb = pert*error_inv (single real32 rounding) and the promotion to double happen in
// registers when a fragment is loaded, so there is no separate conversion pass.
// =====================================================================================================
namespace lk {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct RowMeta2 {
  float ei, yo;
  int pass;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kStages = 3;

// Warp-specialised: the LAST warp of the CTA is the producer.  Per batch of kRows candidate rows it
// evaluates the row metadata (QC verdict, localised 1/error, scaled departure), publishes it in shared
// memory, and gathers the passing rows with bulk asynchronous copies that complete on the stage's `full`
// mbarrier.  The other warps are consumers: wait for `full`, feed the tensor pipe from the stage, arrive on
// the stage's `empty` mbarrier.  No CTA-wide barrier inside the loop, kStages batches in flight.
// NS = 32 x 32 super-blocks per consumer warp.  One per warp keeps every SM sub-partition busy (k = 256: 12
// warps, 3 per sub-partition; two left 6 of 8 warps active and was 1.7x slower); two per warp (168
// registers, hence the smaller block) where the blocks split evenly (k = 96, 128).
template <typename T, int NS>
__global__ void __launch_bounds__(NS == 2 ? 288 : 512)
    gram_tma_kernel(TreeViews tv, int k, int S, int sb_per_cta, int64_t nunits, const int32_t *__restrict__ unit_pt,
                    double mu, T *__restrict__ C, T *__restrict__ bvec, int32_t *__restrict__ nanflag) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int rowb = 4 * k + 32;  // bytes per staged row
  RowMeta2 *meta = reinterpret_cast<RowMeta2 *>(smem_raw + (size_t)kStages * kRows * rowb);  // [kStages][kRows]
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
  __shared__ int s_nan;
  __shared__ unsigned s_pmask[kStages];
  __shared__ int s_last[kStages];

  const int64_t unit = blockIdx.x;
  if (unit >= nunits || (tv.nunits_dev && unit >= *tv.nunits_dev)) return;
  const int64_t q = unit_pt[unit];
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nc = (blockDim.x >> 5) - 1;  // consumer warps
  const int lr = lane >> 2, lc = lane & 3;
  const int NSB = S * (S + 1) / 2;
  if (tid == 0) {
    s_nan = 0;
    for (int st = 0; st < kStages; ++st) {
      mbar_init(&full[st], 1);
      mbar_init(&empty[st], nc);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == nc) {
    // ---------------------------------------- producer ---------------------------------------------
    int it_t = 0, it_c0 = 0;  // tree index and candidate offset of the current batch
    auto advance_to_valid = [&]() {  // skip exhausted / empty trees
      while (it_t < tv.ntrees && it_c0 >= tv.t[it_t].cnt[q] * tv.t[it_t].nact) {
        ++it_t;
        it_c0 = 0;
      }
    };
    advance_to_valid();
    for (int bt = 0;; ++bt) {
      const int slot = bt % kStages;
      if (bt >= kStages) mbar_wait(&empty[slot], (uint32_t)((bt / kStages - 1) & 1));
      RowMeta2 m{0.f, 0.f, 0};
      const float *src = nullptr;
      bool last = true;
      if (it_t < tv.ntrees) {
        const TreeView &TV = tv.t[it_t];
        const int ncand = TV.cnt[q] * TV.nact;
        const int c = it_c0 + lane;
        if (c < ncand) {
          const int j = c / TV.nact, a = c - j * TV.nact;
          const int64_t o = (int64_t)(TV.idx[q * TV.nalloc + j] - 1) * TV.nvar + TV.act[a];
          if (TV.pass[o]) {
            m.pass = 1;
            m.ei = lk_error_inv(TV.err[o], TV.r2[q * TV.nalloc + j], tv.weight_function);
            m.yo = LK_MUL(TV.omm[o], m.ei);
            src = TV.pert + o * k;
            if (m.ei != m.ei) s_nan = 1;
          }
        }
        it_c0 += kRows;
        advance_to_valid();
        last = it_t >= tv.ntrees;
      }
      meta[slot * kRows + lane] = m;
      const unsigned pm = __ballot_sync(0xffffffffu, m.pass != 0);
      if (lane == 0) {
        s_pmask[slot] = pm;
        s_last[slot] = last ? 1 : 0;
      }
      __syncwarp();  // metadata of all lanes ordered before the (releasing) arrive of lane 0
      if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(__popc(pm) * 4 * k));
      __syncwarp();
      if (m.pass) tma_bulk_g2s(smem_raw + ((size_t)slot * kRows + lane) * rowb, src, (uint32_t)(4 * k), &full[slot]);
      if (last) break;
    }
    return;
  }

  // ------------------------------------------ consumers ----------------------------------------------
  int sbi[NS], sbj[NS];
  bool has[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int lin = blockIdx.y * sb_per_cta + warp * NS + s;
    has[s] = (warp * NS + s) < sb_per_cta && lin < NSB;
    int i = 0, rem = has[s] ? lin : 0;
    while (rem > i) {
      rem -= i + 1;
      ++i;
    }
    sbi[s] = i;
    sbj[s] = rem;
  }
  double acc[NS][16][2];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int t = 0; t < 16; ++t) acc[s][t][0] = acc[s][t][1] = 0.0;
  double bacc = 0.0;

  for (int bt = 0;; ++bt) {
    const int slot = bt % kStages;
    mbar_wait(&full[slot], (uint32_t)((bt / kStages) & 1));
    const unsigned pmask = s_pmask[slot];
    const bool last = s_last[slot] != 0;
    const unsigned char *st = smem_raw + (size_t)slot * kRows * rowb;
    const RowMeta2 *mt = meta + slot * kRows;
    if (blockIdx.y == 0 && tid < k) {
#pragma unroll 4
      for (int r = 0; r < kRows; ++r) {
        if (!((pmask >> r) & 1u)) continue;
        const float v = LK_MUL(*reinterpret_cast<const float *>(st + (size_t)r * rowb + 4 * tid), mt[r].ei);
        bacc = fma((double)v, (double)mt[r].yo, bacc);
      }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (!has[s]) continue;
      const bool diag = sbi[s] == sbj[s];
#pragma unroll 2
      for (int g8 = 0; g8 < 8; ++g8) {
        if (((pmask >> (4 * g8)) & 0xFu) == 0u) continue;
        const int r = 4 * g8 + lc;
        const bool ok = (pmask >> r) & 1u;
        const float e = mt[r].ei;
        const float *row = reinterpret_cast<const float *>(st + (size_t)r * rowb) + lr;
        double fa[4], fb[4];
#pragma unroll
        for (int I = 0; I < 4; ++I) {
          const int m = 32 * sbi[s] + 8 * I;
          fa[I] = (ok && m + lr < k) ? (double)LK_MUL(row[m], e) : 0.0;
        }
        if (diag) {
#pragma unroll
          for (int I = 0; I < 4; ++I) fb[I] = fa[I];
        } else {
#pragma unroll
          for (int I = 0; I < 4; ++I) {
            const int m = 32 * sbj[s] + 8 * I;
            fb[I] = (ok && m + lr < k) ? (double)LK_MUL(row[m], e) : 0.0;
          }
        }
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
          for (int J = 0; J < 4; ++J) {
            if (diag && J > I) continue;
            dmma884g(acc[s][I * 4 + J][0], acc[s][I * 4 + J][1], fa[I], fb[J]);
          }
      }
    }
    __syncwarp();  // every lane has read the stage before lane 0 hands it back
    if (lane == 0) mbar_arrive(&empty[slot]);
    if (last) break;
  }
  T *Cu = C + unit * (int64_t)k * k;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    if (!has[s]) continue;
    const bool diag = sbi[s] == sbj[s];
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = 0; J < 4; ++J) {
        if (diag && J > I) continue;
        const int row = 32 * sbi[s] + 8 * I + lr, col = 32 * sbj[s] + 8 * J + 2 * lc;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cc = col + e;
          if (row < k && cc < k) {
            const double v = acc[s][I * 4 + J][e] + (row == cc ? mu : 0.0);
            Cu[(int64_t)cc * k + row] = (T)v;
            if (diag && I == J) Cu[(int64_t)row * k + cc] = (T)v;
          }
        }
      }
  }
  if (blockIdx.y == 0) {
    if (tid < k) bvec[unit * (int64_t)k + tid] = (T)bacc;
    if (tid == 0) nanflag[unit] = s_nan;
  }
}

template <typename T>
void launch_gram_tma(cudaStream_t s, const TreeViews &tv, int k, int64_t nunits, const int32_t *unit_pt, T mu, T *C,
                     T *b, int32_t *nanflag) {
  if (nunits == 0) return;
  LK_REQUIRE(k <= LETKF_B200_MAX_MEMBERS && k % 4 == 0, "launch_gram_tma: k must be a multiple of 4, <= 256");
  const int S = (k + 31) / 32;
  const int NSB = S * (S + 1) / 2;
  const int ncta = (NSB + 15) / 16;
  const int sb_per_cta = (NSB + ncta - 1) / ncta;
  // two super-blocks per warp when that splits evenly (k = 96, 128: fewer barriers per DMMA), else one
  const bool two = sb_per_cta % 2 == 0 && sb_per_cta <= 10;
  const int ns = two ? 2 : 1;
  int nwarps = (sb_per_cta + ns - 1) / ns;
  nwarps = std::max(nwarps, (k + 31) / 32);
  const size_t smem = (size_t)kStages * kRows * (4 * k + 32) + kStages * kRows * sizeof(RowMeta2);
  auto launch = [&](auto kern) {
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // + 1: the producer warp (k <= 256: at most 15 consumer warps)
    LK_REQUIRE(nwarps + 1 <= (two ? 9 : 16), "launch_gram_tma: too many warps");
    kern<<<dim3((unsigned)nunits, (unsigned)ncta), 32 * (nwarps + 1), smem, s>>>(tv, k, S, sb_per_cta, nunits, unit_pt,
                                                                            (double)mu, C, b, nanflag);
  };
  if (two)
    launch(gram_tma_kernel<T, 2>);
  else
    launch(gram_tma_kernel<T, 1>);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}
template void launch_gram_tma<double>(cudaStream_t, const TreeViews &, int, int64_t, const int32_t *, double, double *,
                                      double *, int32_t *);
template void launch_gram_tma<float>(cudaStream_t, const TreeViews &, int, int64_t, const int32_t *, float, float *,
                                     float *, int32_t *);

}  // namespace lk
