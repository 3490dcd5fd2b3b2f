// Host-side construction of the flattened k-d tree the GPU search walks.
//
// Bit-exact local observation lists require the tree of kdtree2_create
// (module_kdtree2.f90:598-680): the same node ranges, the same cut dimensions, and -- because
// the ball query keeps the FIRST max_lz_pts hits of its depth-first walk
// (module_kdtree2.f90:1696-1706) -- the same point permutation inside every bucket.  That
// permutation is defined by a sequential quickselect (select_on_coordinate,
// module_kdtree2.f90:897-929), so the build runs on the host; nodes of disjoint index ranges are
// independent and are built by parallel tasks.  The result is uploaded once per
// (type, hclr, vclr) and cached across variables.
//
// Written iteratively over an explicit work list (the reference recurses): a top-down pass
// does the splits and keeps the APPROXIMATE boxes the split-dimension choice uses
// (module_kdtree2.f90:761-776), a bottom-up pass forms the true boxes and cut values
// (module_kdtree2.f90:821-831).
#include <omp.h>

#include <algorithm>
#include <cstring>

#include "letkf_internal.cuh"

namespace lk {

namespace {

constexpr int kBucket = 12;  // module_kdtree2.f90:505

struct Builder {
  const float *v;  // [n][3]
  int dim, n;
  std::vector<int32_t> ind;  // 0-based values
  // preorder node numbering is known in closed form because every split is at the index
  // midpoint: subtree over a range of `len` points has nodes(len) nodes.
  std::vector<KdNodeDev> nodes;
  std::vector<float> alo, ahi;  // approximate boxes [node][3]

  static int64_t nodes_of(int64_t len) {  // number of nodes of a subtree holding len points
    if (len - 1 <= kBucket) return 1;
    const int64_t l = 1, u = len;
    const int64_t m = (l + u) / 2;
    return 1 + nodes_of(m - l + 1) + nodes_of(u - m);
  }

  inline float coord(int c, int pos) const { return v[(size_t)ind[pos] * 3 + c]; }

  void spread(int c, int l, int u, float &lo, float &hi) const {
    float smin = coord(c, l), smax = smin;
    for (int i = l + 1; i <= u; ++i) {
      const float x = coord(c, i);
      if (x < smin) smin = x;
      if (x > smax) smax = x;
    }
    lo = smin;
    hi = smax;
  }

  // Lomuto quickselect with first-element pivot, exactly the reference's move sequence
  void select(int c, int k, int l, int u) {
    while (l < u) {
      const int t = ind[l];
      const float pv = v[(size_t)t * 3 + c];
      int m = l;
      for (int i = l + 1; i <= u; ++i) {
        if (coord(c, i) < pv) {
          ++m;
          std::swap(ind[m], ind[i]);
        }
      }
      std::swap(ind[l], ind[m]);
      if (m <= k) l = m + 1;
      if (m >= k) u = m - 1;
    }
  }

  // top-down: node `me` over positions [l,u] (0-based inclusive), parent index or -1
  void split(int me, int l, int u, int parent) {
    KdNodeDev &nd = nodes[me];
    std::memset(&nd, 0, sizeof(nd));
    nd.l = l;
    nd.u = u;
    nd.left = nd.right = -1;
    nd.cut_dim = -1;
    float *lo = &alo[(size_t)me * 3], *hi = &ahi[(size_t)me * 3];
    if ((u - l) <= kBucket) {
      for (int c = 0; c < dim; ++c) spread(c, l, u, lo[c], hi[c]);
      return;
    }
    for (int c = 0; c < dim; ++c) {
      if (parent < 0 || c == nodes[parent].cut_dim)
        spread(c, l, u, lo[c], hi[c]);
      else {
        lo[c] = alo[(size_t)parent * 3 + c];
        hi[c] = ahi[(size_t)parent * 3 + c];
      }
    }
    int c = 0;
    float best = hi[0] - lo[0];
    for (int i = 1; i < dim; ++i) {
      const float s = hi[i] - lo[i];
      if (s > best) {
        best = s;
        c = i;
      }
    }
    // Fortran m = (l+u)/2 on 1-based bounds
    const int m = ((l + 1) + (u + 1)) / 2 - 1;
    select(c, m, l, u);
    nd.cut_dim = c;
    const int left = me + 1;
    const int right = left + (int)nodes_of(m - l + 1);
    nd.left = left;
    nd.right = right;
    const bool big = (u - l) > 4096;
    if (big) {
#pragma omp task default(shared) firstprivate(left, l, m, me)
      split(left, l, m, me);
#pragma omp task default(shared) firstprivate(right, m, u, me)
      split(right, m + 1, u, me);
#pragma omp taskwait
    } else {
      split(left, l, m, me);
      split(right, m + 1, u, me);
    }
  }
};

}  // namespace

void build_kdtree_host(const float *xyz, int n, int dim, HostTree &out) {
  LK_REQUIRE(n > 0 && (dim == 2 || dim == 3), "build_kdtree_host: bad arguments");
  Builder b;
  b.v = xyz;
  b.dim = dim;
  b.n = n;
  b.ind.resize(n);
  for (int i = 0; i < n; ++i) b.ind[i] = i;
  const int64_t nn = Builder::nodes_of(n);
  b.nodes.resize(nn);
  b.alo.assign((size_t)nn * 3, 0.f);
  b.ahi.assign((size_t)nn * 3, 0.f);
#pragma omp parallel
  {
#pragma omp single
    b.split(0, 0, n - 1, -1);
  }
  // bottom-up: children follow their parent in preorder, so a reverse scan sees them first
  for (int64_t i = nn - 1; i >= 0; --i) {
    KdNodeDev &nd = b.nodes[i];
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    if (nd.left < 0) {
      for (int c = 0; c < dim; ++c) {
        lo[c] = b.alo[(size_t)i * 3 + c];
        hi[c] = b.ahi[(size_t)i * 3 + c];
      }
    } else {
      const KdNodeDev &L = b.nodes[nd.left], &R = b.nodes[nd.right];
      const float Llo[3] = {L.lo[0], L.lo[1], L.lo[2]}, Lhi[3] = {L.hi0, L.hi1, L.hi2};
      const float Rlo[3] = {R.lo[0], R.lo[1], R.lo[2]}, Rhi[3] = {R.hi0, R.hi1, R.hi2};
      const int c = nd.cut_dim;
      nd.cut_r = Rlo[c];
      nd.cut_l = Lhi[c];
      nd.cut_val = (nd.cut_l + nd.cut_r) / 2;
      for (int d = 0; d < dim; ++d) {
        hi[d] = std::max(Lhi[d], Rhi[d]);
        lo[d] = std::min(Llo[d], Rlo[d]);
      }
    }
    nd.lo[0] = lo[0];
    nd.lo[1] = lo[1];
    nd.lo[2] = lo[2];
    nd.hi0 = hi[0];
    nd.hi1 = hi[1];
    nd.hi2 = hi[2];
  }
  out.dim = dim;
  out.n = n;
  out.nodes = std::move(b.nodes);
  out.pts.resize(n);
  out.ind.resize(n);
  for (int i = 0; i < n; ++i) {
    const int o = b.ind[i];
    float4 p;
    p.x = xyz[(size_t)o * 3 + 0];
    p.y = xyz[(size_t)o * 3 + 1];
    p.z = dim == 3 ? xyz[(size_t)o * 3 + 2] : 0.f;
    const int32_t o1 = o + 1;
    std::memcpy(&p.w, &o1, sizeof(float));
    out.pts[i] = p;
    out.ind[i] = o + 1;
  }
}

}  // namespace lk
