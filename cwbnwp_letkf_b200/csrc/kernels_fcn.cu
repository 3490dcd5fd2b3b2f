// letkf_solve without an eigendecomposition, any k <= 256, FP64: one CTA per analysis unit.
// See fcn_common.cuh for the method.  Stages per unit:
//   1. blocked Householder tridiagonalisation C = Q T Q^T (LAPACK dsytrd / dlatrd organisation: inside
//      a panel of NB columns the trailing matrix is only READ -- one symmetric matrix-vector product per
//      column, corrected with the panel's V, W -- and updated once per panel by the rank-2NB product
//      A -= V W^T + W V^T).  The matrix lives in shared memory for k <= 128, otherwise in place in global
//      memory (L2): its bytes are then streamed once per column instead of three times, and never more
//      than 512 KB per CTA are live (the Jacobi path kept three such matrices per CTA).
//      The reflectors overwrite the eliminated columns (as LAPACK stores them).
//   2. spectrum bound (Gershgorin of T), pole table row q, LDL^T pivots of T + a beta_j I for the 32 poles.
//   3. for every vector (b, the field perturbations of every level that shares the weights, unit
//      vectors for the parity dump): z = Q^T y, g = T^(-1/2) z by the pole solves (lane = pole, forward
//      sweep check-pointed every 16 rows so that no k x 32 array is stored), then u = Q g and the
//      epilogue of letkf_solve (core:671-698).  One warp per vector.
#include "fcn_common.cuh"

namespace lk {

namespace {

constexpr int FCN_NB = 16;    // panel width
constexpr int FCN_NVW = 4;    // vectors in flight (one warp each)

// trailing update A[r][l] -= sum_t V[t][r] W[t][l] + W[t][r] V[t][l] for r, l in [s, k): one warp per
// (group of 32*RT rows, 4 columns); a thread owns rows R0 + lane + 32 a so that every load is coalesced /
// conflict free.  V, W: [t][kv] in shared memory (zero beyond k).
template <int RT>
__device__ __forceinline__ void trailing_update(double *A, int ld, int k, int s, int pb, const double *V,
                                                const double *W, int kv, int warp, int lane, int nw) {
  const int nt = k - s;
  const int nrg = (nt + 32 * RT - 1) / (32 * RT), ncq = (nt + 3) / 4;
  for (int tile = warp; tile < nrg * ncq; tile += nw) {
    const int rg = tile % nrg, cq = tile / nrg;
    const int R0 = s + 32 * RT * rg + lane, l0 = s + 4 * cq;
    double acc[RT][4], old[RT][4];
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        acc[a][b] = 0.0;
        const int r = R0 + 32 * a, l = l0 + b;
        old[a][b] = (r < k && l < k) ? A[r + (size_t)l * ld] : 0.0;
      }
    for (int t = 0; t < pb; ++t) {
      const double *Vt = V + (size_t)t * kv, *Wt = W + (size_t)t * kv;
      double vr[RT], wr[RT], vl[4], wl[4];
#pragma unroll
      for (int a = 0; a < RT; ++a) {
        const int r = R0 + 32 * a;
        vr[a] = r < k ? Vt[r] : 0.0;
        wr[a] = r < k ? Wt[r] : 0.0;
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        vl[b] = l0 + b < k ? Vt[l0 + b] : 0.0;
        wl[b] = l0 + b < k ? Wt[l0 + b] : 0.0;
      }
#pragma unroll
      for (int a = 0; a < RT; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(vr[a], wl[b], fma(wr[a], vl[b], acc[a][b]));
    }
#pragma unroll
    for (int a = 0; a < RT; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int r = R0 + 32 * a, l = l0 + b;
        if (r < k && l < k) A[r + (size_t)l * ld] = old[a][b] - acc[a][b];
      }
  }
}

// y <- H_c y for the stored reflectors, c ascending (Q^T y) or descending (Q y).  y[m] is element
// lane + 32 m.  Reflector c occupies A[c+1 .. k-1][c].
template <int RPL, bool FORWARD>
__device__ __forceinline__ void apply_reflectors(double (&y)[RPL], const double *A, int ld, int k, const double *tau,
                                                 int lane) {
  const int nref = k - 2;
  if (nref <= 0) return;
  auto loadv = [&](int c, double (&v)[RPL]) {
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      const int i = lane + 32 * m;
      v[m] = (i > c && i < k) ? A[i + (size_t)c * ld] : 0.0;
    }
  };
  double v[RPL], vn[RPL];
  loadv(FORWARD ? 0 : nref - 1, v);
  for (int s = 0; s < nref; ++s) {
    const int c = FORWARD ? s : nref - 1 - s;
    if (s + 1 < nref) loadv(FORWARD ? c + 1 : c - 1, vn);  // prefetch the next reflector
    double dot = 0.0;
#pragma unroll
    for (int m = 0; m < RPL; ++m) dot = fma(v[m], y[m], dot);
    dot = wsum(dot) * tau[c];
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      y[m] = fma(-dot, v[m], y[m]);
      v[m] = vn[m];
    }
  }
}

enum { VK_NONE = 0, VK_B = 1, VK_FIELD = 2, VK_WBAR = 3, VK_UNIT = 4 };

template <int RPL>
__global__ void __launch_bounds__(512)
    fcn_blk_kernel(FcnArgs a, int a_smem) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = a.k;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const int kp = (k + 3) & ~3;  // padded vector length (row stride of V, W)
  const int nb = FCN_NB;
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *d = sm;                 // [kp]
  double *e = d + kp;             // [kp]
  double *tau = e + kp;           // [kp]
  double *gb = tau + kp;          // [kp]  T^(-1/2) Q^T b
  double *xcol = gb + kp;         // [kp]
  double *red = xcol + kp;        // [32]
  double *red2 = red + 32;        // [32]
  double *s1 = red2 + 32;         // [32]
  double *s2 = s1 + 32;           // [32]
  double *part = s2 + 32;         // [nw * 32]
  double *VW = part + nw * 32;    // [2 * nb * kp], later rp[k * 32]
  const int vw_len = (2 * nb * kp > 32 * k) ? 2 * nb * kp : 32 * k;
  double *wscr = VW + vw_len;     // FCN_NVW x per-warp scratch
  const int ckn = (k / FCN_SEG + 1) * 32;
  const int wscr_len = 3 * kp + ckn;  // zb[kp], xp[kp], f32 pair [kp], ck[ckn]
  double *As = wscr + FCN_NVW * wscr_len;  // [k * k] when a_smem
  __shared__ int s_q;
  __shared__ double s_aedge;

  const int64_t unit = blockIdx.x;
  if (unit >= a.nunits) return;
  double *Cg = a.C + unit * (int64_t)k * k;
  double *A = a_smem ? As : Cg;
  const int ld = k;
  double *V = VW, *W = VW + (size_t)nb * kp;

  // ---- load / symmetrise (the Gram kernels deliver the column-major lower triangle) ----
  for (int x = tid; x < 2 * nb * kp; x += nt) VW[x] = 0.0;
  if (a_smem) {
    for (int x = tid; x < k * k; x += nt) {
      const int i = x % k, j = x / k;
      As[x] = i >= j ? Cg[i + (size_t)j * k] : Cg[j + (size_t)i * k];
    }
  } else {
    for (int x = tid; x < k * k; x += nt) {
      const int i = x % k, j = x / k;
      if (i < j) Cg[x] = Cg[j + (size_t)i * k];
    }
  }
  __syncthreads();

  // ---- 1. tridiagonalisation ----
  const int r = tid;  // thread per row
  for (int c0 = 0; c0 < k - 2; c0 += nb) {
    const int pb = (k - 2 - c0) < nb ? (k - 2 - c0) : nb;
    for (int i = 0; i < pb; ++i) {
      const int c = c0 + i;
      // phase 1: column c with the pending updates of this panel
      double colv = 0.0;
      if (r >= c && r < k) {
        colv = A[r + (size_t)c * ld];
        for (int t = 0; t < i; ++t)
          colv -= V[t * kp + r] * W[t * kp + c] + W[t * kp + r] * V[t * kp + c];
        xcol[r] = colv;
      }
      {
        double sq = (r > c && r < k) ? colv * colv : 0.0;
        sq = wsum(sq);
        if (lane == 0) red[warp] = sq;
      }
      __syncthreads();
      // phase 2: reflector
      double sigma = 0.0;
      for (int w = 0; w < nw; ++w) sigma += red[w];
      const double x1 = xcol[c + 1];
      const double nrm = sqrt(sigma);
      const bool zero = nrm == 0.0;
      const double alpha = zero ? 0.0 : -copysign(nrm, x1);
      const double tauc = zero ? 0.0 : 1.0 / (nrm * (nrm + fabs(x1)));
      const double v_r = (r == c + 1) ? x1 - alpha : ((r > c + 1 && r < k) ? colv : 0.0);
      if (r < k) V[i * kp + r] = v_r;
      if (tid == 0) {
        d[c] = xcol[c];
        e[c] = alpha;
        tau[c] = tauc;
      }
      __syncthreads();
      // phase 3: p = A v over rows c+1.., split over the warps; dots of v with the panel's V, W
      {
        const int m = k - c - 1;
        const int nrb = (m + 31) >> 5;
        const int nsplit = nw / nrb > 1 ? nw / nrb : 1;
        const double *vi = V + i * kp;
        for (int item = warp; item < nrb * nsplit; item += nw) {
          const int rb = item / nsplit, sp = item - rb * nsplit;
          int row = c + 1 + 32 * rb + lane;
          const bool rowok = row < k;
          row = rowok ? row : k - 1;
          const int lb = c + 1 + (int)(((long long)m * sp) / nsplit);
          const int le = c + 1 + (int)(((long long)m * (sp + 1)) / nsplit);
          const double *Ar = A + row;
          double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
          int l = lb;
          for (; l + 8 <= le; l += 8) {
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x[u] = Ar[(size_t)(l + u) * ld];
            acc0 = fma(x[0], vi[l + 0], acc0);
            acc1 = fma(x[1], vi[l + 1], acc1);
            acc2 = fma(x[2], vi[l + 2], acc2);
            acc3 = fma(x[3], vi[l + 3], acc3);
            acc0 = fma(x[4], vi[l + 4], acc0);
            acc1 = fma(x[5], vi[l + 5], acc1);
            acc2 = fma(x[6], vi[l + 6], acc2);
            acc3 = fma(x[7], vi[l + 7], acc3);
          }
          for (; l < le; ++l) acc0 = fma(Ar[(size_t)l * ld], vi[l], acc0);
          part[item * 32 + lane] = rowok ? (acc0 + acc1) + (acc2 + acc3) : 0.0;
        }
        for (int dt = warp; dt < 2 * i; dt += nw) {
          const int t = dt >> 1;
          const double *src = (dt & 1) ? V + t * kp : W + t * kp;
          double acc = 0.0;
          for (int rr = c + 1 + lane; rr < k; rr += 32) acc = fma(src[rr], vi[rr], acc);
          acc = wsum(acc);
          if (lane == 0) ((dt & 1) ? s2 : s1)[t] = acc;
        }
      }
      __syncthreads();
      // phase 4: p = tau (A_eff v), K = tau/2 p.v
      double pr = 0.0;
      if (r > c && r < k) {
        const int m = k - c - 1;
        const int nrb = (m + 31) >> 5;
        const int nsplit = nw / nrb > 1 ? nw / nrb : 1;
        const int rr = r - (c + 1), rb = rr >> 5, ln = rr & 31;
        for (int sp = 0; sp < nsplit; ++sp) pr += part[(rb * nsplit + sp) * 32 + ln];
        for (int t = 0; t < i; ++t) pr -= V[t * kp + r] * s1[t] + W[t * kp + r] * s2[t];
        pr *= tauc;
      }
      {
        double pv = wsum(pr * v_r);
        if (lane == 0) red2[warp] = pv;
      }
      __syncthreads();
      // phase 5: w = p - K v; store the reflector over the eliminated column
      double K = 0.0;
      for (int w = 0; w < nw; ++w) K += red2[w];
      K *= 0.5 * tauc;
      if (r < k) {
        W[i * kp + r] = fma(-K, v_r, pr);
        if (r > c) A[r + (size_t)c * ld] = v_r;
      }
      __syncthreads();
    }
    // trailing update with this panel
    {
      const int s = c0 + pb, ntr = k - s;
      if (ntr > 64)
        trailing_update<4>(A, ld, k, s, pb, V, W, kp, warp, lane, nw);
      else if (ntr > 32)
        trailing_update<2>(A, ld, k, s, pb, V, W, kp, warp, lane, nw);
      else
        trailing_update<1>(A, ld, k, s, pb, V, W, kp, warp, lane, nw);
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (k >= 2) {
      d[k - 2] = A[(k - 2) + (size_t)(k - 2) * ld];
      e[k - 2] = A[(k - 1) + (size_t)(k - 2) * ld];
      tau[k - 2] = 0.0;
    }
    d[k - 1] = A[(k - 1) + (size_t)(k - 1) * ld];
    e[k - 1] = 0.0;
    tau[k - 1] = 0.0;
  }
  __syncthreads();

  // ---- 2. spectrum bound, pole row, pivots (warp 0) ----
  double *rp = VW;
  if (warp == 0) {
    double g = 0.0;
    for (int i = lane; i < k; i += 32) {
      const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i + 1 < k ? fabs(e[i]) : 0.0;
      g = fmax(g, d[i] + el + er);
    }
    g = wmax(g);
    const double aedge = fcn_lower_edge(a.mu);
    const int q = fcn_interval(g, aedge);
    if (lane == 0) {
      s_q = q;
      s_aedge = aedge;
    }
    pole_pivots(k, d, e, aedge * a.poles[(q * 2 + 1) * FCN_NP + lane], rp, lane);
  }
  __syncthreads();
  const double cw = sqrt(s_aedge) * a.poles[(s_q * 2 + 0) * FCN_NP + lane];

  // ---- 3. vectors ----
  const bool dump = a.wbar_out != nullptr || a.Wa_out != nullptr;
  const int nfv = a.var ? a.nz * a.nfields : 0;
  // list: [b][fields ...][pad so that wbar is not in the first batch][wbar][unit vectors]
  const int nvb = nw < FCN_NVW ? nw : FCN_NVW;  // vectors per batch: one warp each
  const int i_wbar = dump ? ((1 + nfv) > nvb ? (1 + nfv) : nvb) : -1;
  const int nvec = dump ? i_wbar + 1 + (a.Wa_out ? k : 0) : 1 + nfv;
  const bool isnan_unit = a.nanflag[unit] != 0;
  const int64_t upt = a.unit_pt[unit];
  const float ninv = LK_DIV(1.0f, (float)k);
  const double sk = sqrt((double)(k - 1));  // core:666/668
  const int slot = nw - 1 - warp;           // vector slot of this warp (warp 0 last)
  double *zb = wscr + (size_t)(slot < nvb ? slot : 0) * wscr_len;
  double *xp = zb + kp;
  float *xb32 = reinterpret_cast<float *>(xp + kp);
  float *xa32 = xb32 + kp;
  double *ck = xp + 2 * kp;

  for (int v0 = 0; v0 < nvec; v0 += nvb) {
    const int vi = v0 + slot;
    int kind = VK_NONE, sub = 0;
    if (slot < nvb && vi < nvec) {
      if (vi == 0) kind = VK_B;
      else if (vi <= nfv) { kind = VK_FIELD; sub = vi - 1; }
      else if (vi == i_wbar) kind = VK_WBAR;
      else if (dump && vi > i_wbar) { kind = VK_UNIT; sub = vi - i_wbar - 1; }
    }
    double y[RPL];
    double xmean = 0.0;
    int64_t pt = 0;
    float *vfield = nullptr;
    if (kind != VK_NONE) {
      if (kind == VK_B) {
        const double *bv = a.bvec + unit * (int64_t)k;
#pragma unroll
        for (int m = 0; m < RPL; ++m) { const int i = lane + 32 * m; y[m] = i < k ? bv[i] : 0.0; }
      } else if (kind == VK_FIELD) {
        const int lev = sub / a.nfields, f = sub - lev * a.nfields;
        pt = a.pt_base + (int64_t)lev * a.level_stride + upt;
        vfield = a.var + (int64_t)f * a.npts_total * k;
        for (int i = lane; i < k; i += 32) xb32[i] = vfield[(int64_t)i * a.npts_total + pt];  // core:228
        __syncwarp();
        float s = 0.f;  // xb_mean = sum(xb) * nmember_inv in real32, sequential (core:671)
        for (int i = 0; i < k; ++i) s = LK_ADD(s, xb32[i]);
        xmean = (double)LK_MUL(s, ninv);
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
          const int i = lane + 32 * m;
          y[m] = i < k ? (double)xb32[i] - xmean : 0.0;  // core:672
          if (i < k) xp[i] = y[m];
        }
      } else if (kind == VK_WBAR) {
#pragma unroll
        for (int m = 0; m < RPL; ++m) { const int i = lane + 32 * m; y[m] = i < k ? gb[i] : 0.0; }
      } else {
#pragma unroll
        for (int m = 0; m < RPL; ++m) y[m] = (lane + 32 * m == sub) ? 1.0 : 0.0;
      }
      if (kind != VK_WBAR) apply_reflectors<RPL, true>(y, A, ld, k, tau, lane);  // z = Q^T y
#pragma unroll
      for (int m = 0; m < RPL; ++m) { const int i = lane + 32 * m; if (i < k) zb[i] = y[m]; }
      __syncwarp();
      pole_solve(zb, k, e, rp, cw, ck, lane);  // zb <- T^(-1/2) z
    }
    if (v0 == 0) {  // publish g_b
      __syncthreads();
      if (slot == 0) for (int i = lane; i < k; i += 32) gb[i] = zb[i];
      __syncthreads();
    }
    if (kind == VK_FIELD || kind == VK_WBAR || kind == VK_UNIT) {
      double sdot = 0.0;
#pragma unroll
      for (int m = 0; m < RPL; ++m) {
        const int i = lane + 32 * m;
        y[m] = i < k ? zb[i] : 0.0;
        if (kind == VK_FIELD && i < k) sdot = fma(y[m], gb[i], sdot);
      }
      sdot = wsum(sdot);  // xb' . wbar (core:673)
      apply_reflectors<RPL, false>(y, A, ld, k, tau, lane);  // u = Q g
      if (kind == VK_WBAR) {
        if (a.wbar_out)
#pragma unroll
          for (int m = 0; m < RPL; ++m) { const int i = lane + 32 * m; if (i < k) a.wbar_out[upt * k + i] = y[m]; }
      } else if (kind == VK_UNIT) {
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
          const int i = lane + 32 * m;
          if (i < k) a.Wa_out[upt * (int64_t)k * k + (int64_t)sub * k + i] = sk * y[m];
        }
      } else {
        // epilogue of letkf_solve (core:673-698), as kernels_xform.cu
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
          const int i = lane + 32 * m;
          if (i < k) {
            double xa = xmean + (sdot + sk * y[m]);
            if (isnan_unit) xa = xa * (double)NAN;
            if (a.xa_raw) a.xa_raw[pt * k + i] = xa;
            xa32[i] = (float)xa;  // core:679
          }
        }
        __syncwarp();
        if (a.use_rtpp || a.use_rtps) {
          float s = 0.f;
          for (int i = 0; i < k; ++i) s = LK_ADD(s, xa32[i]);
          const float xa_mean = LK_MUL(s, ninv);
          __syncwarp();
          for (int i = lane; i < k; i += 32) {
            float xap = LK_SUB(xa32[i], xa_mean);
            if (a.use_rtpp) {  // core:689
              const float t1 = LK_MUL(LK_SUB(1.0f, a.rtpp_alpha), xap);
              xap = (float)((double)t1 + (double)a.rtpp_alpha * xp[i]);
            }
            xa32[i] = xap;
          }
          __syncwarp();
          if (a.use_rtps) {  // core:692-694
            double dd = 0.0;
            float xa_std = 0.f;
            for (int i = 0; i < k; ++i) {
              dd += xp[i] * xp[i];
              xa_std = LK_ADD(xa_std, LK_MUL(xa32[i], xa32[i]));
            }
            const float xb_std = (float)dd;
            const float fac =
                LK_ADD(LK_SUB(LK_MUL(a.rtps_alpha, LK_SQRT(LK_DIV(xb_std, xa_std))), a.rtps_alpha), 1.0f);
            __syncwarp();
            for (int i = lane; i < k; i += 32) xa32[i] = LK_MUL(xa32[i], fac);
            __syncwarp();
          }
          for (int i = lane; i < k; i += 32) xa32[i] = LK_ADD(xa_mean, xa32[i]);  // core:697
          __syncwarp();
        }
        for (int i = lane; i < k; i += 32) vfield[(int64_t)i * a.npts_total + pt] = xa32[i];  // core:229
      }
    }
    __syncwarp();
  }
}

}  // namespace

size_t fcn_blk_smem(int k, int threads, bool a_smem) {
  const int kp = (k + 3) & ~3, nw = threads / 32;
  const size_t vw = std::max<size_t>((size_t)2 * FCN_NB * kp, (size_t)32 * k);
  const size_t wscr = (size_t)3 * kp + (size_t)(k / FCN_SEG + 1) * 32;
  size_t n = (size_t)5 * kp + 4 * 32 + (size_t)nw * 32 + vw + FCN_NVW * wscr;
  if (a_smem) n += (size_t)k * k;
  return n * sizeof(double);
}

void launch_fcn_solve(cudaStream_t s, const FcnArgs &a) {
  if (a.nunits == 0) return;
  const int k = a.k;
  LK_REQUIRE(k >= 2 && k <= LETKF_B200_MAX_MEMBERS, "fcn solver: need 2 <= k <= 256");
  int threads = ((2 * k + 31) / 32) * 32;
  threads = std::max(64, std::min(512, threads));
  const bool a_smem = fcn_blk_smem(k, threads, true) <= 220 * 1024;
  const size_t smem = fcn_blk_smem(k, threads, a_smem);
  LK_REQUIRE(a.nunits < ((int64_t)1 << 31), "fcn solver: too many units for one launch");
  auto launch = [&](auto kern) {
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)a.nunits, threads, smem, s>>>(a, a_smem ? 1 : 0);
  };
  switch ((k + 31) / 32) {
    case 1: launch(fcn_blk_kernel<1>); break;
    case 2: launch(fcn_blk_kernel<2>); break;
    case 3: launch(fcn_blk_kernel<3>); break;
    case 4: launch(fcn_blk_kernel<4>); break;
    case 5: launch(fcn_blk_kernel<5>); break;
    case 6: launch(fcn_blk_kernel<6>); break;
    case 7: launch(fcn_blk_kernel<7>); break;
    default: launch(fcn_blk_kernel<8>); break;
  }
  LK_CUDA(cudaGetLastError());
}

}  // namespace lk
