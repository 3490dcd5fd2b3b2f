// letkf_solve without an eigendecomposition, any k <= 256, FP64: one CTA per analysis unit.
// See fcn_common.cuh for the method.  Stages per unit:
//   1. blocked Householder tridiagonalisation C = Q T Q^T (LAPACK dsytrd / dlatrd organisation: inside
//      a panel of NB columns the trailing matrix is only READ -- one symmetric matrix-vector product per
//      column, corrected with the panel's V, W -- and updated once per panel by the rank-2NB product
//      A -= V W^T + W V^T on the FP64 tensor pipe).  The matrix is processed in place in global memory and
//      stays L2 resident: its bytes are streamed once per column instead of three times, never more than the
//      k x k doubles of C are live per CTA (the Jacobi path kept three such matrices per CTA), and the small
//      shared-memory footprint (panel + vectors) lets many CTAs share an SM.
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
constexpr int FCN_NVW = 2;    // vectors in flight (one warp each): b and one field column is the common case

__device__ __forceinline__ void dmma884f(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Trailing update on the FP64 tensor pipe: A[r][l] -= sum_t P[t][r] Q[t][l] with P = [V; W], Q = [W; V], one
// warp per 32 x 32 output tile (16 mma.m8n8k4 per 8 fragment loads).  VW: rows [0, nb) = V, [nb, 2nb) = W,
// row stride kv = 4 (mod 16) doubles so that the fragment loads are bank-conflict free.  Rows [pb, pb4) of
// V and W must be zero.  One instruction per 256 multiply-adds instead of one per 32.
template <typename AP>
__device__ __forceinline__ void trailing_update_dmma(AP A, int ld, int k, int s, int pb, const double *VW, int kv,
                                                     int nb, int warp, int lane, int nw) {
  const int nt = k - s, nt32 = (nt + 31) >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int pb4 = (pb + 3) & ~3;
  for (int tile = warp; tile < nt32 * nt32; tile += nw) {
    const int tj = tile / nt32, ti = tile - tj * nt32;
    const int r0 = s + 32 * ti, l0 = s + 32 * tj;
    const bool interior = r0 + 32 <= k && l0 + 32 <= k;  // warp-uniform
    double acc[4][4][2];
#pragma unroll
    for (int I = 0; I < 4; ++I)
#pragma unroll
      for (int J = 0; J < 4; ++J) acc[I][J][0] = acc[I][J][1] = 0.0;
    int ro[4], lo[4];
#pragma unroll
    for (int I = 0; I < 4; ++I) {
      const int r = r0 + 8 * I + lr, l = l0 + 8 * I + lr;
      ro[I] = interior ? r : (r < k ? r : k - 1);
      lo[I] = interior ? l : (l < k ? l : k - 1);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const double *P = VW + (half ? nb * kv : 0) + lc * kv, *Q = VW + (half ? 0 : nb * kv) + lc * kv;
      for (int t0 = 0; t0 < pb4; t0 += 4, P += 4 * kv, Q += 4 * kv) {
        double a[4], b[4];
#pragma unroll
        for (int I = 0; I < 4; ++I) {
          a[I] = P[ro[I]];
          b[I] = Q[lo[I]];
        }
#pragma unroll
        for (int I = 0; I < 4; ++I)
#pragma unroll
          for (int J = 0; J < 4; ++J) dmma884f(acc[I][J][0], acc[I][J][1], a[I], b[J]);
      }
    }
#pragma unroll
    for (int J = 0; J < 4; ++J) {
      const int col = l0 + 8 * J + 2 * lc;
#pragma unroll
      for (int I = 0; I < 4; ++I) {
        const int row = r0 + 8 * I + lr;
        if (interior || (row < k && col < k)) A[row + (size_t)col * ld] -= acc[I][J][0];
        if (interior || (row < k && col + 1 < k)) A[row + (size_t)(col + 1) * ld] -= acc[I][J][1];
      }
    }
  }
}

// One share of p = A v (rows c+1 .. k-1): a warp takes R row blocks (rows row0 + 32 a) and a contiguous
// range of columns; the column pointer steps by ld, the row offsets are immediates.  Rows >= k read
// whatever follows (the caller pads the allocation) and are masked when the partial sums are stored.
// part[((item * R) + a) * 32 + lane].
template <int R, typename AP>
__device__ __forceinline__ void symv_part(AP A, int ld, int k, int c, const double *vi, double *part, int warp,
                                          int lane, int nw) {
  const int m = k - c - 1;
  const int ngrp = (m + 32 * R - 1) / (32 * R);
  int lg = 0;  // nsplit = largest power of two <= nw / ngrp
  while ((ngrp << (lg + 1)) <= nw) ++lg;
  const int nsplit = 1 << lg;
  for (int item = warp; item < (ngrp << lg); item += nw) {
    const int g = item >> lg, sp = item & (nsplit - 1);
    const int row0 = c + 1 + 32 * R * g + lane;
    const int lb = c + 1 + ((m * sp) >> lg), le = c + 1 + ((m * (sp + 1)) >> lg);
    double acc[R];
#pragma unroll
    for (int a = 0; a < R; ++a) acc[a] = 0.0;
    AP pa = A + (size_t)lb * ld + row0;
    int l = lb;
    for (; l + 4 <= le; l += 4) {
      double x[4][R];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int a = 0; a < R; ++a) x[u][a] = pa[32 * a];
        pa += ld;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double vl = vi[l + u];
#pragma unroll
        for (int a = 0; a < R; ++a) acc[a] = fma(x[u][a], vl, acc[a]);
      }
    }
    for (; l < le; ++l, pa += ld) {
      const double vl = vi[l];
#pragma unroll
      for (int a = 0; a < R; ++a) acc[a] = fma(pa[32 * a], vl, acc[a]);
    }
#pragma unroll
    for (int a = 0; a < R; ++a) part[(item * R + a) * 32 + lane] = (row0 + 32 * a < k) ? acc[a] : 0.0;
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

// shared-memory layout (doubles), shared by the kernel and the host-side size computation
struct FcnSmem {
  int kp, nw, part_len, vw_len, wscr_len, total;
  __host__ __device__ FcnSmem(int k, int nw_) {
    kp = ((k + 15) & ~15) + 4;  // row stride of V, W: 4 (mod 16) doubles (conflict-free mma fragment loads)
    nw = nw_;
    part_len = nw * 4 * 32;
    vw_len = (2 * FCN_NB * kp > 32 * k) ? 2 * FCN_NB * kp : 32 * k;
    wscr_len = 3 * kp + (k / FCN_SEG + 1) * 32;  // zb[kp], xp[kp], f32 pair [kp], ck
    total = 5 * kp + 4 * 32 + part_len + vw_len + FCN_NVW * wscr_len;
  }
};

template <int RPL>
__global__ void __launch_bounds__(256, 2)
    fcn_blk_kernel(FcnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = a.k;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  const FcnSmem L(k, nw);
  const int kp = L.kp;  // padded vector length (row stride of V, W)
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
  double *part = s2 + 32;         // [nw * 4 * 32]
  double *VW = part + L.part_len; // [2 * nb * kp], later rp[k * 32]
  double *wscr = VW + L.vw_len;   // FCN_NVW x per-warp scratch
  const int wscr_len = L.wscr_len;
  __shared__ int s_q;
  __shared__ double s_aedge;

  const int64_t unit = blockIdx.x;
  if (unit >= a.nunits || (a.nunits_dev && unit >= *a.nunits_dev)) return;
  double *Cg = a.C + unit * (int64_t)k * k;
  double *A = Cg;  // processed in place in global memory: L2 resident (rows >= k of a column read the next column)
  const int ld = k;
  double *V = VW, *W = VW + (size_t)nb * kp;

  // ---- load / symmetrise (the Gram kernels deliver the column-major lower triangle) ----
  for (int x = tid; x < 2 * nb * kp; x += nt) VW[x] = 0.0;
  for (int j = warp; j < k; j += nw)
    for (int i = lane; i < j; i += 32) Cg[i + (size_t)j * k] = Cg[j + (size_t)i * k];
  __syncthreads();

  // ---- 1. tridiagonalisation ----
  const int r = tid;  // thread per row
  for (int c0 = 0; c0 < k - 2; c0 += nb) {
    const int pb = (k - 2 - c0) < nb ? (k - 2 - c0) : nb;
    for (int i = 0; i < pb; ++i) {
      const int c = c0 + i;
      const int m = k - c - 1;          // rows / columns of the trailing block
      const int nrb = (m + 31) >> 5;
      const int RR = nrb >= 4 ? 4 : (nrb >= 2 ? 2 : 1);  // row blocks per warp in the product (CTA-uniform)
      // phase 1: column c with the pending updates of this panel
      double colv = 0.0;
      if (r >= c && r < k) {
        colv = A[r + (size_t)c * ld];
        const double *pv = V + r, *pw = W + r, *qv = V + c, *qw = W + c;
        for (int t = 0; t < i; ++t, pv += kp, pw += kp, qv += kp, qw += kp)
          colv -= pv[0] * qw[0] + pw[0] * qv[0];
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
      const bool zero = !(sigma > 1e-290) && sigma == sigma;  // NaN falls through and propagates
      const double rn = rsqrt(zero ? 1.0 : sigma);
      const double nrm = zero ? 0.0 : sigma * rn;
      const double alpha = zero ? 0.0 : -copysign(nrm, x1);
      const double tauc = zero ? 0.0 : rn * __drcp_rn(nrm + fabs(x1));
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
        const double *vi = V + i * kp;
        if (RR == 4)
          symv_part<4>(A, ld, k, c, vi, part, warp, lane, nw);
        else if (RR == 2)
          symv_part<2>(A, ld, k, c, vi, part, warp, lane, nw);
        else
          symv_part<1>(A, ld, k, c, vi, part, warp, lane, nw);
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
        const int lgR = RR == 4 ? 2 : (RR == 2 ? 1 : 0);
        const int ngrp = (nrb + RR - 1) >> lgR;
        int lg = 0;
        while ((ngrp << (lg + 1)) <= nw) ++lg;
        const int rr = r - (c + 1), rb = rr >> 5, ln = rr & 31;
        const int g = rb >> lgR, ai = rb & (RR - 1);
        const double *pp = part + (((g << lg) * RR + ai) << 5) + ln;
        for (int sp = 0; sp < (1 << lg); ++sp, pp += RR * 32) pr += pp[0];
        const double *pv = V + r, *pw = W + r;
        for (int t = 0; t < i; ++t, pv += kp, pw += kp) pr -= pv[0] * s1[t] + pw[0] * s2[t];
        pr *= tauc;
      }
      {
        double pvs = wsum(pr * v_r);
        if (lane == 0) red2[warp] = pvs;
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
    // trailing update with this panel (FP64 tensor pipe)
    {
      const int s = c0 + pb;
      const int pb4 = (pb + 3) & ~3;
      if (pb4 != pb) {  // last panel: rows [pb, pb4) of V and W still hold the previous panel
        for (int x = tid; x < (pb4 - pb) * kp; x += nt) {
          V[pb * kp + x] = 0.0;
          W[pb * kp + x] = 0.0;
        }
        __syncthreads();
      }
      trailing_update_dmma(A, ld, k, s, pb, VW, kp, nb, warp, lane, nw);
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
      if (q > FCN_QSAFE && a.qmax) atomicMax(a.qmax, q);
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

size_t fcn_blk_smem(int k, int threads) { return sizeof(double) * (size_t)FcnSmem(k, threads / 32).total; }

void launch_fcn_solve(cudaStream_t s, const FcnArgs &a) {
  if (a.nunits == 0) return;
  const int k = a.k;
  LK_REQUIRE(k >= 2 && k <= LETKF_B200_MAX_MEMBERS, "fcn solver: need 2 <= k <= 256");
  // One thread per matrix row in the panel phases and no more: small CTAs, many per SM (k = 64: 28 KB of shared
  // memory and 2 warps per CTA, 8 CTAs per SM).  Measured on B200 (profiles/README.md): keeping the matrix in
  // shared memory (one or two CTAs per SM) is 1.4x SLOWER than leaving it in L2 with more CTAs resident, and
  // CTAs larger than k threads lose to their own barriers.
  const int threads = std::max(64, ((k + 31) / 32) * 32);
  const size_t smem = fcn_blk_smem(k, threads);
  LK_REQUIRE(a.nunits < ((int64_t)1 << 31), "fcn solver: too many units for one launch");
  auto launch = [&](auto kern) {
    LK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)a.nunits, threads, smem, s>>>(a);
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
