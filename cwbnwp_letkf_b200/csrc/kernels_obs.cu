// Observation-space kernels: the grid-point independent half of letkf_yoyb
// (module_letkf_core.f90:430-437,497-510), row counting, the debug yo/yb dump, letkf_tune_q.
//
// All real32 arithmetic here goes through the LK_* single-rounding operations of
// letkf_b200_math.h in the order the oracle defines (sums over members run 0..k-1).
#include "letkf_internal.cuh"

namespace lk {

CtxShared *&current_ctx_shared() {
  static thread_local CtxShared *cur = nullptr;
  return cur;
}
int64_t &launch_counter() {
  static thread_local int64_t unbound = 0;  // launches outside any context (none in practice)
  CtxShared *c = current_ctx_shared();
  return c ? c->launches : unbound;
}

// ---- once per set_obs: mean, perturbations, spread, any(qc>=0) -------------------------------
// hdxb is the reference array: gts hdxb(nvar,n,0:k-1), radar hdxb(n,0:k-1) (nvar = 1).
// One thread per (ob, slot); reads are coalesced across threads for each member.
__global__ void obs_static_kernel(int k, int n, int nvar, bool gts, const float *__restrict__ hdxb,
                                  const int32_t *__restrict__ qc, float *__restrict__ pert,
                                  float *__restrict__ mean_o, float *__restrict__ std_o,
                                  uint8_t *__restrict__ anyqc) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // o = i*nvar + s
  const int64_t total = (int64_t)n * nvar;
  if (o >= total) return;
  const float ninv = LK_DIV(1.0f, (float)k);          // module_param.f90:129
  const float n1inv = LK_DIV(1.0f, (float)(k - 1));   // module_param.f90:130
  const int64_t mstride = total;                      // member stride of hdxb / qc
  float sum = 0.0f;
  bool any = !gts;
  for (int m = 0; m < k; ++m) {
    sum = LK_ADD(sum, hdxb[o + mstride * m]);
    if (gts) any = any || (qc[o + mstride * m] >= 0);
  }
  const float mean = LK_MUL(sum, ninv);               // core:431/498
  float dot = 0.0f;
  float *pr = pert + o * k;
  for (int m = 0; m < k; ++m) {
    const float d = LK_SUB(hdxb[o + mstride * m], mean);  // core:432/499
    pr[m] = d;
    dot = LK_ADD(dot, LK_MUL(d, d));
  }
  mean_o[o] = mean;
  std_o[o] = LK_SQRT(LK_MUL(dot, n1inv));             // core:434/501
  anyqc[o] = any ? 1 : 0;
}

void launch_obs_static(cudaStream_t s, int k, int n, int nvar, bool gts, const float *hdxb,
                       const int32_t *qc, float *pert, float *mean, float *stdv, uint8_t *anyqc) {
  const int64_t total = (int64_t)n * nvar;
  if (total == 0) return;
  const int bs = 128;
  obs_static_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, s>>>(k, n, nvar, gts, hdxb, qc, pert,
                                                                     mean, stdv, anyqc);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

// ---- once per variable: error, innovation, QC verdict ----------------------------------------
struct SlotCfg {
  int is_assim[LETKF_B200_MAX_SLOTS];
  float err_muti[LETKF_B200_MAX_SLOTS];
  float err_rej[LETKF_B200_MAX_SLOTS];
};

__global__ void obs_config_kernel(int n, int nvar, bool gts, bool is_dbz, const float *__restrict__ obs,
                                  const float *__restrict__ error, const float *__restrict__ mean,
                                  const float *__restrict__ stdv, const uint8_t *__restrict__ anyqc,
                                  SlotCfg sc, float norain, float *__restrict__ err_o,
                                  float *__restrict__ omm_o, uint8_t *__restrict__ pass_o) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (int64_t)n * nvar) return;
  const int s = (int)(o % nvar);
  const float ob = obs[o];
  const float mu = mean[o];
  const float sd = stdv[o];
  const float err = gts ? LK_MUL(error[o], sc.err_muti[s]) : sc.err_muti[0];  // core:435 / 488,502
  const float omm = LK_SUB(ob, mu);                                          // core:433/500
  const float lim = LK_MUL(LK_SQRT(LK_ADD(LK_MUL(sd, sd), LK_MUL(err, err))), sc.err_rej[gts ? s : 0]);
  const bool gross = fabsf(omm) > lim;                                       // core:437/505/509
  bool pass;
  if (gts) {
    pass = sc.is_assim[s] && anyqc[o] && !gross;                             // core:429,437
  } else if (is_dbz) {
    pass = !(gross && ob != norain) && !(ob == norain && mu == norain);      // core:504-507
  } else {
    pass = !gross;                                                           // core:509
  }
  err_o[o] = err;
  omm_o[o] = omm;
  pass_o[o] = pass ? 1 : 0;
}

void launch_obs_config(cudaStream_t s, int n, int nvar, bool gts, bool is_dbz, const float *obs,
                       const float *error, const float *mean, const float *stdv, const uint8_t *anyqc,
                       const letkf_b200_type_config &tc, float norain, float *err, float *omm,
                       uint8_t *pass) {
  const int64_t total = (int64_t)n * nvar;
  if (total == 0) return;
  SlotCfg sc;
  for (int i = 0; i < LETKF_B200_MAX_SLOTS; ++i) {
    sc.is_assim[i] = tc.is_assim[i];
    sc.err_muti[i] = tc.err_muti[i];
    sc.err_rej[i] = tc.err_rej[i];
  }
  const int bs = 256;
  obs_config_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, s>>>(n, nvar, gts, is_dbz, obs, error, mean,
                                                                     stdv, anyqc, sc, norain, err, omm, pass);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

// ---- rows per point: p = #candidates whose pass flag is set ----------------------------------
__global__ void count_rows_kernel(TreeViews tv, int64_t nq, int32_t *__restrict__ p_out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  int p = 0;
  for (int t = 0; t < tv.ntrees; ++t) {
    const TreeView &T = tv.t[t];
    const int c = T.cnt[q];
    const int32_t *idx = T.idx + q * T.nalloc;
    for (int j = 0; j < c; ++j) {
      const int64_t o = (int64_t)(idx[j] - 1) * T.nvar;
      for (int a = 0; a < T.nact; ++a) p += T.pass[o + T.act[a]];
    }
  }
  p_out[q] = p;
}

void launch_count_rows(cudaStream_t s, const TreeViews &tv, int64_t nq, int32_t *p) {
  if (nq == 0) return;
  const int bs = 128;
  count_rows_kernel<<<(unsigned)((nq + bs - 1) / bs), bs, 0, s>>>(tv, nq, p);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

// ---- debug dump of yo / yb in the reference row order (letkf_yoyb, core:300-595) -------------
// One warp per point; candidates are walked in order (tree, list entry, slot) 32 at a time and
// the passing ones are compacted with a ballot, so the row order is the reference's.
__global__ void yoyb_rows_kernel(TreeViews tv, int k, int64_t nq, const int64_t *__restrict__ row_offset,
                                 float *__restrict__ yo, float *__restrict__ yb) {
  const int lane = threadIdx.x & 31;
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  int64_t row = row_offset[q];
  for (int t = 0; t < tv.ntrees; ++t) {
    const TreeView &T = tv.t[t];
    const int ncand = T.cnt[q] * T.nact;
    for (int c0 = 0; c0 < ncand; c0 += 32) {
      const int c = c0 + lane;
      bool pass = false;
      float ei = 0.f, yov = 0.f;
      const float *pr = nullptr;
      if (c < ncand) {
        const int j = c / T.nact, a = c - j * T.nact;
        const int64_t o = (int64_t)(T.idx[q * T.nalloc + j] - 1) * T.nvar + T.act[a];
        pass = T.pass[o] != 0;
        if (pass) {
          ei = lk_error_inv(T.err[o], T.r2[q * T.nalloc + j], tv.weight_function);
          yov = LK_MUL(T.omm[o], ei);  // core:451/524
          pr = T.pert + o * k;
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, pass);
      if (pass) yo[row + __popc(mask & ((1u << lane) - 1))] = yov;
      for (unsigned mm = mask; mm; mm &= mm - 1) {
        const int src = __ffs(mm) - 1;
        const float e = __shfl_sync(0xffffffffu, ei, src);
        const float *p = (const float *)__shfl_sync(0xffffffffu, (unsigned long long)pr, src);
        const int64_t r = row + __popc(mask & ((1u << src) - 1));
        for (int m = lane; m < k; m += 32) yb[r * k + m] = LK_MUL(p[m], e);  // core:452/525
      }
      row += __popc(mask);
    }
  }
}

void launch_yoyb_rows(cudaStream_t s, const TreeViews &tv, int k, int64_t nq, const int64_t *row_offset,
                      float *yo, float *yb) {
  if (nq == 0) return;
  const int bs = 128;
  const int64_t threads = nq * 32;
  yoyb_rows_kernel<<<(unsigned)((threads + bs - 1) / bs), bs, 0, s>>>(tv, k, nq, row_offset, yo, yb);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

// ---- letkf_tune_q (module_letkf_core.f90:702-733) ---------------------------------------------
// One thread per grid point: coalesced across points for every member.  ratio = sum(var) /
// sum(var, var>0) in real32, sequential over members; 0/0 -> NaN like the reference.
// Points [p0, p0 + n) of a field whose member stride is npts.
__global__ void tune_q_kernel(int k, int64_t npts, int64_t p0, int64_t n, float *__restrict__ var) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t pt = p0 + i;
  float s_all = 0.f, s_pos = 0.f;
  for (int m = 0; m < k; ++m) {
    const float v = var[(int64_t)m * npts + pt];
    s_all = LK_ADD(s_all, v);
    if (v > 0.f) s_pos = LK_ADD(s_pos, v);
  }
  const float ratio = LK_DIV(s_all, s_pos);
  for (int m = 0; m < k; ++m) {
    const float v = var[(int64_t)m * npts + pt];
    var[(int64_t)m * npts + pt] = (v < 0.f) ? 0.f : LK_MUL(ratio, v);
  }
}

void launch_tune_q(cudaStream_t s, int k, int64_t npts, float *var, int64_t p0, int64_t n) {
  if (n < 0) n = npts - p0;
  if (n <= 0) return;
  const int bs = 256;
  tune_q_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, s>>>(k, npts, p0, n, var);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

}  // namespace lk
