// Internal declarations shared by the translation units of libletkf_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/letkf_b200.h"
#include "../../include/letkf_b200_math.h"

namespace lk {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define LK_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw lk::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                      ":" + std::to_string(__LINE__) + ")");                               \
  } while (0)

#define LK_REQUIRE(cond, msg)                \
  do {                                       \
    if (!(cond)) throw lk::Error(msg);       \
  } while (0)

// ---- device buffer -------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void ensure(size_t count) {  // grow-only
    if (count <= n) return;
    release();
    LK_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
};

// ---- flattened kdtree2 (module_kdtree2.f90), device layout ----------------------------
// 64-byte node: four float4 loads.  Positions l,u index the rearranged point array
// (0-based, inclusive).  cut_dim is 0-based; left < 0 marks a terminal node.
struct alignas(16) KdNodeDev {
  float cut_val, cut_l, cut_r;
  int32_t cut_dim;
  int32_t l, u, left, right;
  float lo[3];
  float hi0;
  float hi1, hi2;
  int32_t pad0, pad1;
};
static_assert(sizeof(KdNodeDev) == 64, "node must be 64 bytes");

struct HostTree {
  int dim = 0, n = 0;
  std::vector<KdNodeDev> nodes;  // preorder; root = 0
  std::vector<float4> pts;       // rearranged_data(:,i) + original 1-based index in .w (as int bits)
  std::vector<int32_t> ind;      // permutation (1-based values)
};
// host build, identical topology / permutation to kdtree2_create (kd2:598-834,897-979)
void build_kdtree_host(const float *xyz /*[n][3] already normalised*/, int n, int dim, HostTree &out);
void selftest_host_search(const HostTree &ht, int64_t nq, const float *xyz, float hinv, float vinv, int nalloc,
                          int32_t *cnt, int32_t *idx, float *r2out);

struct DevTree {
  int dim = 0, n = 0, nnodes = 0;
  float hclr = 0, vclr = 0;  // cache key
  DevBuf<KdNodeDev> nodes;
  DevBuf<float4> pts;
};

// ---- observations ----------------------------------------------------------------------
struct ObsDev {
  int family = 0, type = 0, n = 0, nvar = 0;
  std::vector<float> h_xyz;     // host copy of xyz[3,n] for the tree build
  DevBuf<float> xyz;            // [n][3]
  DevBuf<float> obs;            // [n][nvar]
  DevBuf<float> error;          // [n][nvar] (gts)
  // grid-point independent half of letkf_yoyb, computed once in set_obs:
  DevBuf<float> pert;           // [n][nvar][k]  bg - mean (ob-major, member contiguous)
  DevBuf<float> mean;           // [n][nvar]
  DevBuf<float> stdv;           // [n][nvar]     sqrt(dot(bg,bg)*nmember_1_inv)
  DevBuf<uint8_t> anyqc;        // [n][nvar]     any(qc >= 0)   (radar: 1)
  // per-variable half (depends on err_muti / err_rej / norain): filled per analyze
  DevBuf<float> err;            // [n][nvar]     error*err_muti or namelist error
  DevBuf<float> omm;            // [n][nvar]     obs - mean
  DevBuf<uint8_t> pass;         // [n][nvar]     row survives is_assim + qc + gross-error checks
  std::map<std::pair<float, float>, std::unique_ptr<DevTree>> trees;  // keyed by (hclr, vclr)
};

// one active tree of the current variable, as the kernels see it
struct TreeView {
  const KdNodeDev *nodes;
  const float4 *pts;
  int dim;
  float hinv, vinv;     // 1/(hclr*1e3), 1/(vclr*1e3) (vinv unused for dim 2)
  int nalloc;           // max_lz_pts
  int nvar;             // slots of this type
  int nact;             // assimilated slots
  int act[LETKF_B200_MAX_SLOTS];
  const float *pert, *omm, *err;
  const uint8_t *pass;
  // per-chunk search output
  int32_t *cnt;         // [chunk]
  int32_t *idx;         // [chunk][nalloc]
  float *r2;            // [chunk][nalloc]
  int family, type;
};
struct TreeViews {
  int ntrees;
  int weight_function;
  // Number of analysis units of the current chunk as the compaction left it ON THE DEVICE (null: the launch
  // argument is exact).  Lets the host queue a chunk's Gram / solve kernels with the chunk size as the grid,
  // without waiting for the count: units beyond *nunits_dev exit at once.
  const int32_t *nunits_dev;
  TreeView t[LETKF_B200_MAX_TYPES];
};

// ---- kernels (host launchers) -----------------------------------------------------------
void launch_obs_static(cudaStream_t s, int k, int n, int nvar, bool gts, const float *hdxb,
                       const int32_t *qc, float *pert, float *mean, float *stdv, uint8_t *anyqc);
void launch_obs_config(cudaStream_t s, int n, int nvar, bool gts, bool is_dbz, const float *obs,
                       const float *error, const float *mean, const float *stdv,
                       const uint8_t *anyqc, const letkf_b200_type_config &tc, float norain,
                       float *err, float *omm, uint8_t *pass);
void launch_search(cudaStream_t s, const TreeView &tv, int64_t nq, const float *xyz /*[nq][3]*/);
// rows (passing candidates) per point and "has any list entry" flags
void launch_count_rows(cudaStream_t s, const TreeViews &tv, int64_t nq, int32_t *p);

template <typename T>
void launch_gram(cudaStream_t s, const TreeViews &tv, int k, int64_t nunits, const int32_t *unit_pt,
                 T mu, T *C /*[nunits][k][k]*/, T *b /*[nunits][k]*/, int32_t *nanflag);
// tensor-core (FP64 mma) variant for any k; writes the column-major lower triangle of C
template <typename T>
void launch_gram_dmma(cudaStream_t s, const TreeViews &tv, int k, int64_t nunits, const int32_t *unit_pt, T mu,
                      T *C, T *b, int32_t *nanflag);
// same with the rows gathered by TMA bulk copies into a double-buffered stage (k % 4 == 0)
template <typename T>
void launch_gram_tma(cudaStream_t s, const TreeViews &tv, int k, int64_t nunits, const int32_t *unit_pt, T mu, T *C,
                     T *b, int32_t *nanflag);
// C,b -> U (orthonormal eigenvectors, column-major), lam (unsorted), wbar = U diag(1/lam) U^T b
template <typename T>
void launch_eig_solve(cudaStream_t s, int k, int64_t nunits, T *C_inout_U, const T *b, T *lam,
                      T *wbar, int32_t *sweeps_max);
// plain eigensolver: A (lower) -> W ascending, V
template <typename T>
void launch_syevd(cudaStream_t s, int k, int64_t batch, const T *A, T *W, T *V, int32_t *sweeps_max);
template <typename T>
void launch_transform(cudaStream_t s, int k, int64_t nunits, const int32_t *unit_pt, int64_t npts_total,
                      int64_t pt_base, const T *U, const T *lam, const T *wbar, const int32_t *nanflag,
                      int nfields, float *var, int use_rtpp, float rtpp_alpha, int use_rtps,
                      float rtps_alpha, double *xa_raw /*optional [npts][k]*/);
template <typename T>
void launch_weights_dump(cudaStream_t s, int k, int64_t nunits, const int32_t *unit_pt, const T *U,
                         const T *lam, const T *wbar, double *wbar_out, double *Wa_out);
void launch_tune_q(cudaStream_t s, int k, int64_t npts, float *var, int64_t p0 = 0, int64_t n = -1);
void launch_yoyb_rows(cudaStream_t s, const TreeViews &tv, int k, int64_t nq, const int64_t *row_offset,
                      float *yo, float *yb);
double run_fma_peak(cudaStream_t s, int kind);

// ---- k = 32 fast paths (warp per analysis unit; U stored by rows: U[i][j] at i*32+j) ----------
void launch_gram32(cudaStream_t s, const TreeViews &tv, int64_t nunits, const int32_t *unit_pt, double mu,
                   double *C, double *b, int32_t *nanflag);
struct Xform32Args;
// fuse != nullptr: apply the transform in the solver's epilogue (FP64 chained kernel only); U/lam/wbar are
// then not written
template <typename T>
void launch_eig32_solve(cudaStream_t s, int64_t n, T *C_inout_U, const T *b, T *lam, T *wbar, int32_t *sweeps_max,
                        const Xform32Args *fuse);
bool eig32_can_fuse();
template <typename T>
void launch_syevd32(cudaStream_t s, int64_t n, const T *A, T *W, T *V, int32_t *sweeps_max);
template <typename T>
void launch_transform32(cudaStream_t s, int64_t nunits, const int32_t *unit_pt, int64_t npts_total, int64_t pt_base,
                        const T *U, const T *lam, const T *wbar, const int32_t *nanflag, int nfields, float *var,
                        int use_rtpp, float rtpp_alpha, int use_rtps, float rtps_alpha, double *xa_raw);
template <typename T>
void launch_weights_dump32(cudaStream_t s, int64_t nunits, const int32_t *unit_pt, const T *U, const T *lam,
                           const T *wbar, double *wbar_out, double *Wa_out);

// ---- matrix-function solver (fcn_common.cuh): FP64 replacement of eigen + transform (+ weights dump) ----
struct FcnArgs {
  int k;
  int64_t nunits;
  double *C;            // [nunits][k][k]: lower triangle (column-major) or full symmetric; overwritten
  const double *bvec;   // [nunits][k]
  const int32_t *unit_pt, *nanflag;
  double mu;            // (k-1)/rho: lower bound of the spectrum
  const double *poles;  // device copy of the pole table
  const int32_t *nunits_dev;  // device-side unit count (null: nunits is exact), see TreeViews
  int32_t *qmax;        // atomicMax of the interval index q over the units (spectrum inside [a, a 2^q]); may be null
  // transform: var == nullptr skips it.  Point of (unit, level) = pt_base + level*level_stride + unit_pt[unit]
  int64_t npts_total, pt_base, level_stride;
  int nz, nfields;
  float *var;
  int use_rtpp;
  float rtpp_alpha;
  int use_rtps;
  float rtps_alpha;
  double *xa_raw;
  // parity dump (either may be null), indexed by unit_pt
  double *wbar_out, *Wa_out;
};
void launch_fcn_solve(cudaStream_t s, const FcnArgs &a);     // CTA per unit, any k
void launch_fcn32_solve(cudaStream_t s, const FcnArgs &a);   // warp per unit, k = 32
const std::vector<double> &fcn_pole_table_host();

// Per-context state the kernel launchers need (launch count, the warm-start scratch of the large-k
// Jacobi kernel).  Every C-ABI entry point binds its context to the calling thread for the duration of
// the call (CtxBind in api.cu); nothing is process-global, so contexts on different devices / threads do
// not share or race on it.
struct CtxShared {
  int64_t launches = 0;
  DevBuf<unsigned char> eig_scratch;
};
CtxShared *&current_ctx_shared();
int64_t &launch_counter();

// sweep caps of the Jacobi eigensolvers; a kernel reports cap + 1 when the stop criterion was not met
constexpr int LK_JACOBI_CAP = 40, LK_JACOBI_CAP32 = 30;

}  // namespace lk
