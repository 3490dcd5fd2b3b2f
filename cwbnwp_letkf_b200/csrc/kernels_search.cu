// Fixed-radius local observation search: get_lz -> kdtree2_r_nearest
// (module_localization.f90:188-331, module_kdtree2.f90:1118-1179,1381-1477,1619-1712).
//
// One thread per query walks the flattened tree with an explicit stack.  The recursion of the
// reference (closer child first, then the farther child iff the cut-plane distance and the
// node-box distance stay within r2) is reproduced exactly: the pruning test of the farther
// child depends only on the node and the query, so it is evaluated on the way down and the
// child is pushed; popping restores the reference's visiting order.  Hits are stored in
// visiting order until max_lz_pts is reached; the reference keeps walking after that without
// storing anything (module_kdtree2.f90:1696-1706), so stopping there returns the same list.
//
// Distances accumulate dimension by dimension in real32 with single roundings (kd2:1677-1681).
// Neighbouring threads are neighbouring grid points (x fastest), so their walks mostly coincide
// and node / bucket loads coalesce into a few L2 sectors.
#include <cstring>

#include "letkf_internal.cuh"

namespace lk {

template <typename V>
__host__ __device__ __forceinline__ V ld_ro(const V *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
__host__ __device__ __forceinline__ int f2i(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int i;
  memcpy(&i, &f, sizeof(int));
  return i;
#endif
}

// The walk for one (already normalised) query.  __host__ __device__ so that the host-only
// self-test entry point (letkf_b200_selftest_host_search) can check tree + walk against the
// oracle on a machine without a GPU; the pipeline only ever runs it inside search_kernel.
template <int DIM>
__host__ __device__ __forceinline__ int search_one(const KdNodeDev *__restrict__ nodes,
                                                   const float4 *__restrict__ pts, float q0, float q1,
                                                   float q2, float r2, int nalloc, int32_t *__restrict__ oi,
                                                   float *__restrict__ od) {
  int nfound = 0;
  int stack[48];
  int sp = 0;
  int cur = 0;
  const float4 *n4 = reinterpret_cast<const float4 *>(nodes);
  while (true) {
    int l, u;
    while (true) {  // descend to a terminal node (kd2:1402-1456)
      const float4 a = ld_ro(n4 + (size_t)cur * 4 + 0);  // cut_val, cut_l, cut_r, cut_dim
      const int4 b = ld_ro(reinterpret_cast<const int4 *>(n4 + (size_t)cur * 4 + 1));  // l,u,left,right
      if (b.z < 0) {
        l = b.x;
        u = b.y;
        break;
      }
      const int cd = f2i(a.w);
      const float qval = cd == 0 ? q0 : (cd == 1 ? q1 : q2);
      int closer, farther;
      float dis;
      if (qval < a.x) {  // kd2:1415-1418
        closer = b.z;
        farther = b.w;
        const float d = LK_SUB(a.z, qval);
        dis = LK_MUL(d, d);
      } else {  // kd2:1420-1423
        closer = b.w;
        farther = b.z;
        const float d = LK_SUB(a.y, qval);
        dis = LK_MUL(d, d);
      }
      if (dis <= r2) {  // kd2:1433-1448
        const float4 lo = ld_ro(n4 + (size_t)cur * 4 + 2);  // lo0 lo1 lo2 hi0
        const float4 hi = ld_ro(n4 + (size_t)cur * 4 + 3);  // hi1 hi2
        bool visit = true;
        if (cd != 0) {
          dis = LK_ADD(dis, lk_dis2_from_bnd(q0, lo.x, lo.w));
          visit = !(dis > r2);
        }
        if (visit && cd != 1) {
          dis = LK_ADD(dis, lk_dis2_from_bnd(q1, lo.y, hi.x));
          visit = !(dis > r2);
        }
        if (DIM == 3 && visit && cd != 2) {
          dis = LK_ADD(dis, lk_dis2_from_bnd(q2, lo.z, hi.y));
          visit = !(dis > r2);
        }
        if (visit) stack[sp++] = farther;
      }
      cur = closer;
    }
    // terminal node: scan the bucket in storage order (kd2:1654-1707).  A bucket holds at most 13
    // points (kd2:505,737); all of them are fetched first so that the loads are in flight together
    // (the scan itself has data-dependent exits, which would otherwise serialise 13 L2 round trips).
    float4 pv[13];
#pragma unroll
    for (int j = 0; j < 13; ++j)
      if (l + j <= u) pv[j] = ld_ro(pts + l + j);
#pragma unroll
    for (int j = 0; j < 13; ++j) {
      if (l + j > u) break;
      const float4 p = pv[j];
      float d = LK_SUB(p.x, q0);
      float sd = LK_MUL(d, d);
      if (sd > r2) continue;
      d = LK_SUB(p.y, q1);
      sd = LK_ADD(sd, LK_MUL(d, d));
      if (sd > r2) continue;
      if (DIM == 3) {
        d = LK_SUB(p.z, q2);
        sd = LK_ADD(sd, LK_MUL(d, d));
        if (sd > r2) continue;
      }
      oi[nfound] = f2i(p.w);
      od[nfound] = sd;
      ++nfound;
      if (nfound >= nalloc) return nfound;
    }
    if (sp == 0) break;
    cur = stack[--sp];
  }
  return nfound;
}

template <int DIM>
__global__ void __launch_bounds__(128)
    search_kernel(const KdNodeDev *__restrict__ nodes, const float4 *__restrict__ pts, int64_t nq,
                  const float *__restrict__ xyz, float hinv, float vinv, float r2, int nalloc,
                  int32_t *__restrict__ out_cnt, int32_t *__restrict__ out_idx, float *__restrict__ out_r2) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  // normalise by the localisation length scales (loc:243-246,299-302)
  const float q0 = LK_MUL(xyz[q * 3 + 0], hinv);
  const float q1 = LK_MUL(xyz[q * 3 + 1], hinv);
  const float q2 = DIM == 3 ? LK_MUL(xyz[q * 3 + 2], vinv) : 0.f;
  out_cnt[q] = search_one<DIM>(nodes, pts, q0, q1, q2, r2, nalloc, out_idx + q * nalloc, out_r2 + q * nalloc);
}

// Host-only self-test of tree build + walk (no GPU needed).  NOT part of any product path.
void selftest_host_search(const HostTree &ht, int64_t nq, const float *xyz, float hinv, float vinv, int nalloc,
                          int32_t *cnt, int32_t *idx, float *r2out) {
  const float r2 = lk_search_r2();
  for (int64_t q = 0; q < nq; ++q) {
    const float q0 = LK_MUL(xyz[q * 3 + 0], hinv), q1 = LK_MUL(xyz[q * 3 + 1], hinv);
    const float q2 = ht.dim == 3 ? LK_MUL(xyz[q * 3 + 2], vinv) : 0.f;
    cnt[q] = ht.dim == 3 ? search_one<3>(ht.nodes.data(), ht.pts.data(), q0, q1, q2, r2, nalloc, idx + q * nalloc,
                                         r2out + q * nalloc)
                         : search_one<2>(ht.nodes.data(), ht.pts.data(), q0, q1, q2, r2, nalloc, idx + q * nalloc,
                                         r2out + q * nalloc);
  }
}

void launch_search(cudaStream_t s, const TreeView &tv, int64_t nq, const float *xyz) {
  if (nq == 0) return;
  const int bs = 128;
  const unsigned grid = (unsigned)((nq + bs - 1) / bs);
  const float r2 = lk_search_r2();
  if (tv.dim == 3)
    search_kernel<3><<<grid, bs, 0, s>>>(tv.nodes, tv.pts, nq, xyz, tv.hinv, tv.vinv, r2, tv.nalloc, tv.cnt,
                                         tv.idx, tv.r2);
  else
    search_kernel<2><<<grid, bs, 0, s>>>(tv.nodes, tv.pts, nq, xyz, tv.hinv, tv.vinv, r2, tv.nalloc, tv.cnt,
                                         tv.idx, tv.r2);
  launch_counter()++;
  LK_CUDA(cudaGetLastError());
}

}  // namespace lk
