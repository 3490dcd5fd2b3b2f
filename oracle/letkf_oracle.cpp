/*
 * letkf_oracle.cpp -- CPU restatement of the CWBNWP-LETKF per-grid-point local analysis.
 *
 * TEST INFRASTRUCTURE ONLY (see letkf_oracle.h).  PARITY UNPINNED by the reference's own
 * tests (it has none and cannot be compiled here); pinned by tests/test_oracle_*.py.
 *
 * Each function cites the reference file:line it follows (paths under /root/reference).
 * Conventions the Fortran leaves to the compiler are fixed as in
 * include/letkf_b200_math.h: real32 ops are single IEEE roundings without FMA
 * contraction (this file is compiled with -ffp-contract=off), sum()/dot_product() run
 * left to right, exp() is lk_expf().  BLAS/LAPACK calls go to the same routines the
 * reference calls (?syrk ?syevd ?gemm ?gemv ?symv), here from scipy's bundled
 * OpenBLAS 0.3.31.dev / LAPACK 3.12.0 (the reference links Fujitsu SSL2, not vendored).
 */
#include "letkf_oracle.h"

#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../include/letkf_b200_math.h"

/* ---- BLAS / LAPACK (Fortran ABI, LP64, scipy_ prefix) -------------------------------- */
extern "C" {
void scipy_dsyrk_(const char *, const char *, const int *, const int *, const double *,
                  const double *, const int *, const double *, double *, const int *, size_t,
                  size_t);
void scipy_ssyrk_(const char *, const char *, const int *, const int *, const float *,
                  const float *, const int *, const float *, float *, const int *, size_t, size_t);
void scipy_dsyevd_(const char *, const char *, const int *, double *, const int *, double *,
                   double *, const int *, int *, const int *, int *, size_t, size_t);
void scipy_ssyevd_(const char *, const char *, const int *, float *, const int *, float *, float *,
                   const int *, int *, const int *, int *, size_t, size_t);
void scipy_dgemm_(const char *, const char *, const int *, const int *, const int *, const double *,
                  const double *, const int *, const double *, const int *, const double *,
                  double *, const int *, size_t, size_t);
void scipy_sgemm_(const char *, const char *, const int *, const int *, const int *, const float *,
                  const float *, const int *, const float *, const int *, const float *, float *,
                  const int *, size_t, size_t);
void scipy_dgemv_(const char *, const int *, const int *, const double *, const double *,
                  const int *, const double *, const int *, const double *, double *, const int *,
                  size_t);
void scipy_sgemv_(const char *, const int *, const int *, const float *, const float *,
                  const int *, const float *, const int *, const float *, float *, const int *,
                  size_t);
void scipy_dsymv_(const char *, const int *, const double *, const double *, const int *,
                  const double *, const int *, const double *, double *, const int *, size_t);
void scipy_ssymv_(const char *, const int *, const float *, const float *, const int *,
                  const float *, const int *, const float *, float *, const int *, size_t);
void scipy_openblas_set_num_threads(int);
}

static thread_local std::string g_err;
static int fail(const std::string &m) {
  g_err = m;
  return -1;
}
extern "C" const char *or_last_error(void) { return g_err.c_str(); }

/* reference enums (module_param.f90:28-57, 93-97) */
enum { GTS_SOUND = 1, GTS_SYNOP = 2, GTS_GPSPW = 8, GTS_METAR = 10, GTS_SHIPS = 11, NUM_GTS = 29 };
enum { RAD_DBZ = 1, RAD_VR = 2, RAD_ZDR = 3, RAD_KDP = 4, NUM_RADAR = 4 };

/* ======================================================================================
 * kdtree2 (module_kdtree2.f90).  Indices l,u,ind are 1-based as in the reference.
 * ====================================================================================== */
struct KdNode {
  int cut_dim = 0;          /* 1-based; 0 for terminal nodes (kd2:744) */
  float cut_val = 0, cut_val_left = 0, cut_val_right = 0;
  int l = 0, u = 0;
  int left = -1, right = -1; /* node indices, -1 = null */
  float lower[3] = {0, 0, 0}, upper[3] = {0, 0, 0};
};

struct or_kdtree {
  int dimen = 0, n = 0;
  std::vector<float> the_data;   /* [3,n] column-major copy (kd2:634) */
  std::vector<int> ind;          /* 1-based values, stored at [0..n) for positions 1..n */
  std::vector<float> rearranged; /* [dimen,n] (kd2:670-675) */
  std::vector<KdNode> nodes;
  int root = -1;
  float v(int c, int idx1) const { return the_data[(size_t)(idx1 - 1) * 3 + (c - 1)]; }
  int &I(int pos1) { return ind[pos1 - 1]; }
  int I(int pos1) const { return ind[pos1 - 1]; }
};

static const int bucket_size = 12; /* kd2:505 */

/* spread_in_coordinate (kd2:931-979) */
static void spread_in_coordinate(const or_kdtree &t, int c, int l, int u, float &lo, float &hi) {
  float smin = t.v(c, t.I(l));
  float smax = smin;
  int i;
  for (i = l + 2; i <= u; i += 2) {
    float lmin = t.v(c, t.I(i - 1));
    float lmax = t.v(c, t.I(i));
    if (lmin > lmax) std::swap(lmin, lmax);
    if (smin > lmin) smin = lmin;
    if (smax < lmax) smax = lmax;
  }
  if (i == u + 1) {
    float last = t.v(c, t.I(u));
    if (smin > last) smin = last;
    if (smax < last) smax = last;
  }
  lo = smin;
  hi = smax;
}

/* select_on_coordinate (kd2:897-929) */
static void select_on_coordinate(or_kdtree &t, int c, int k, int li, int ui) {
  int l = li, u = ui;
  while (l < u) {
    int tt = t.I(l);
    int m = l;
    for (int i = l + 1; i <= u; ++i) {
      if (t.v(c, t.I(i)) < t.v(c, tt)) {
        m = m + 1;
        std::swap(t.I(m), t.I(i));
      }
    }
    std::swap(t.I(l), t.I(m));
    if (m <= k) l = m + 1;
    if (m >= k) u = m - 1;
  }
}

/* build_tree_for_range (kd2:696-834); parent = -1 for the root */
static int build_tree_for_range(or_kdtree &t, int l, int u, int parent) {
  if (u < l) return -1; /* kd2:731-735 */
  const int dimen = t.dimen;
  const int me = (int)t.nodes.size();
  t.nodes.emplace_back();
  if ((u - l) <= bucket_size) { /* kd2:737-749 terminal node: true bounding box */
    KdNode nd;
    for (int i = 1; i <= dimen; ++i) spread_in_coordinate(t, i, l, u, nd.lower[i - 1], nd.upper[i - 1]);
    nd.cut_dim = 0;
    nd.cut_val = 0.0f;
    nd.l = l;
    nd.u = u;
    t.nodes[me] = nd;
    return me;
  }
  KdNode nd;
  for (int i = 1; i <= dimen; ++i) { /* kd2:761-773 approximate box */
    bool recompute = true;
    if (parent >= 0 && i != t.nodes[parent].cut_dim) recompute = false;
    if (recompute) {
      spread_in_coordinate(t, i, l, u, nd.lower[i - 1], nd.upper[i - 1]);
    } else {
      nd.lower[i - 1] = t.nodes[parent].lower[i - 1];
      nd.upper[i - 1] = t.nodes[parent].upper[i - 1];
    }
  }
  int c = 1; /* maxloc: first maximum (kd2:776) */
  {
    float best = nd.upper[0] - nd.lower[0];
    for (int i = 2; i <= dimen; ++i) {
      float s = nd.upper[i - 1] - nd.lower[i - 1];
      if (s > best) {
        best = s;
        c = i;
      }
    }
  }
  const int m = (l + u) / 2; /* kd2:784 */
  select_on_coordinate(t, c, m, l, u);
  nd.cut_dim = c;
  nd.l = l;
  nd.u = u;
  t.nodes[me] = nd; /* children read parent's cut_dim and approximate box */
  const int left = build_tree_for_range(t, l, m, me);
  const int right = build_tree_for_range(t, m + 1, u, me);
  KdNode &r = t.nodes[me];
  r.left = left;
  r.right = right;
  if (right < 0) { /* kd2:812-815 (unreachable for u-l > bucket_size; kept for fidelity) */
    for (int i = 0; i < 3; ++i) {
      r.lower[i] = t.nodes[left].lower[i];
      r.upper[i] = t.nodes[left].upper[i];
    }
    r.cut_val_left = t.nodes[left].upper[c - 1];
    r.cut_val = r.cut_val_left;
  } else if (left < 0) { /* kd2:816-819 */
    for (int i = 0; i < 3; ++i) {
      r.lower[i] = t.nodes[right].lower[i];
      r.upper[i] = t.nodes[right].upper[i];
    }
    r.cut_val_right = t.nodes[right].lower[c - 1];
    r.cut_val = r.cut_val_right;
  } else { /* kd2:820-832 */
    r.cut_val_right = t.nodes[right].lower[c - 1];
    r.cut_val_left = t.nodes[left].upper[c - 1];
    r.cut_val = (r.cut_val_left + r.cut_val_right) / 2;
    for (int i = 0; i < dimen; ++i) {
      r.upper[i] = std::max(t.nodes[left].upper[i], t.nodes[right].upper[i]);
      r.lower[i] = std::min(t.nodes[left].lower[i], t.nodes[right].lower[i]);
    }
  }
  return me;
}

/* kdtree2_create (kd2:598-680): sort=.false., rearrange=.true. defaults */
extern "C" or_kdtree *or_kd_create(const float *data, int n, int dim) {
  or_kdtree *t = new or_kdtree();
  t->dimen = dim;
  t->n = n;
  t->the_data.assign(data, data + (size_t)3 * n);
  t->ind.resize(n);
  for (int j = 1; j <= n; ++j) t->I(j) = j; /* kd2:689-692 */
  t->nodes.reserve(n / 6 + 16);
  t->root = build_tree_for_range(*t, 1, n, -1);
  t->rearranged.resize((size_t)dim * n);
  for (int i = 1; i <= n; ++i)
    for (int c = 1; c <= dim; ++c) t->rearranged[(size_t)(i - 1) * dim + (c - 1)] = t->v(c, t->I(i));
  return t;
}
extern "C" void or_kd_destroy(or_kdtree *t) { delete t; }

struct SearchRec { /* tree_search_record (kd2:565-588), fixed-ball subset */
  int dimen, nfound, nalloc;
  float ballsize;
  bool overflow;
  const float *qv; /* 0-based */
  int32_t *res_idx;
  float *res_dis;
  long found_total; /* extra: in-ball count including overflowed hits (not in reference) */
};

/* process_terminal_node_fixedball (kd2:1619-1712), rearrange=.true. branch */
static void process_terminal_node_fixedball(const or_kdtree &t, SearchRec &sr, const KdNode &node) {
  const int dimen = sr.dimen;
  const float ballsize = sr.ballsize;
  int nfound = sr.nfound;
  for (int i = node.l; i <= node.u; ++i) {
    float sd = 0.0f;
    bool out = false;
    for (int k = 1; k <= dimen; ++k) {
      const float d = t.rearranged[(size_t)(i - 1) * dimen + (k - 1)] - sr.qv[k - 1];
      sd = sd + d * d;
      if (sd > ballsize) {
        out = true;
        break;
      }
    }
    if (out) continue;
    const int indexofi = t.I(i);
    sr.found_total++;
    nfound = nfound + 1;
    if (nfound > sr.nalloc) { /* kd2:1697-1702: overflow, nothing more is stored */
      sr.overflow = true;
      nfound = sr.nalloc;
      break;
    } else {
      sr.res_dis[nfound - 1] = sd;
      sr.res_idx[nfound - 1] = indexofi;
    }
  }
  sr.nfound = nfound;
}

/* search (kd2:1381-1457) */
static void search(const or_kdtree &t, SearchRec &sr, int node_i) {
  const KdNode &node = t.nodes[node_i];
  if (!(node.left >= 0 && node.right >= 0)) {
    process_terminal_node_fixedball(t, sr, node);
    return;
  }
  const float *qv = sr.qv;
  const int cut_dim = node.cut_dim;
  const float qval = qv[cut_dim - 1];
  int ncloser, nfarther;
  float dis;
  if (qval < node.cut_val) {
    ncloser = node.left;
    nfarther = node.right;
    const float d = node.cut_val_right - qval;
    dis = d * d;
  } else {
    ncloser = node.right;
    nfarther = node.left;
    const float d = node.cut_val_left - qval;
    dis = d * d;
  }
  if (ncloser >= 0) search(t, sr, ncloser);
  if (nfarther >= 0) {
    const float ballsize = sr.ballsize;
    if (dis <= ballsize) {
      for (int i = 1; i <= sr.dimen; ++i) {
        if (i != cut_dim) {
          dis = dis + lk_dis2_from_bnd(qv[i - 1], node.lower[i - 1], node.upper[i - 1]);
          if (dis > ballsize) return;
        }
      }
      search(t, sr, nfarther);
    }
  }
}

/* kdtree2_r_nearest (kd2:1118-1179).  Returns nfound (<= nalloc); idx 1-based. */
extern "C" int or_kd_r_nearest(const or_kdtree *t, const float *qv, float r2, int nalloc,
                               int32_t *idx, float *dis, int *nfound_total) {
  SearchRec sr;
  sr.qv = qv;
  sr.ballsize = r2;
  sr.nfound = 0;
  sr.res_idx = idx;
  sr.res_dis = dis;
  sr.nalloc = nalloc;
  sr.overflow = false;
  sr.dimen = t->dimen;
  sr.found_total = 0;
  if (t->root >= 0) search(*t, sr, t->root);
  if (nfound_total) *nfound_total = (int)sr.found_total;
  return sr.nfound;
}

/* kdtree2_r_nearest_brute_force idea (kd2:1755-1793): O(n) scan in ORIGINAL index order
 * with the same real32 distance arithmetic; returns every in-ball index (up to nalloc). */
extern "C" int or_kd_brute(const or_kdtree *t, const float *qv, float r2, int nalloc, int32_t *idx,
                           float *dis) {
  int nf = 0;
  for (int i = 1; i <= t->n; ++i) {
    float sd = 0.0f;
    for (int k = 1; k <= t->dimen; ++k) {
      const float d = t->v(k, i) - qv[k - 1];
      sd = sd + d * d;
    }
    if (sd <= r2) {
      if (nf < nalloc) {
        idx[nf] = i;
        dis[nf] = sd;
      }
      nf++;
    }
  }
  return nf;
}

extern "C" int or_kd_num_nodes(const or_kdtree *t) { return (int)t->nodes.size(); }
extern "C" void or_kd_dump(const or_kdtree *t, int32_t *cut_dim, float *cut_val, float *cut_l,
                           float *cut_r, int32_t *l, int32_t *u, int32_t *left, int32_t *right,
                           float *box, int32_t *ind) {
  for (size_t i = 0; i < t->nodes.size(); ++i) {
    const KdNode &n = t->nodes[i];
    cut_dim[i] = n.cut_dim;
    cut_val[i] = n.cut_val;
    cut_l[i] = n.cut_val_left;
    cut_r[i] = n.cut_val_right;
    l[i] = n.l;
    u[i] = n.u;
    left[i] = n.left;
    right[i] = n.right;
    for (int c = 0; c < 3; ++c) {
      box[i * 6 + c * 2 + 0] = n.lower[c];
      box[i * 6 + c * 2 + 1] = n.upper[c];
    }
  }
  for (int i = 0; i < t->n; ++i) ind[i] = t->ind[i];
}

/* ======================================================================================
 * context: obs containers + trees + per-thread eigen workspace
 * ====================================================================================== */
struct ObsType {
  int n = 0, nvar = 0;
  const float *xyz = nullptr, *obs = nullptr, *error = nullptr, *hdxb = nullptr;
  const int32_t *qc = nullptr;
};
struct TreeEntry { /* kdtree_type (loc:12-15) + what get_lz re-derives from the namelist */
  int family, mytype, cfg_index, dim;
  std::unique_ptr<or_kdtree> tree;
};
struct EigenWs { /* module eigen state (eig:4-12), one per thread */
  int n = 0, lwork = 0, liwork = 0;
  int prec = -1;
  std::vector<double> work_d, eval_d, evect_d;
  std::vector<float> work_s, eval_s, evect_s;
  std::vector<int> iwork;
};
struct or_ctx {
  int k = 0;
  bool real64 = true;
  float nmember_inv = 0, nmember_1_inv = 0; /* par:121-131 */
  ObsType gts[NUM_GTS + 1], rad[NUM_RADAR + 1];
  std::vector<TreeEntry> trees; /* gts trees first, then radar (core:217-218) */
  int n_gts_trees = 0;
};

extern "C" or_ctx *or_create(int nmember, int real64) {
  scipy_openblas_set_num_threads(1);
  or_ctx *c = new or_ctx();
  c->k = nmember;
  c->real64 = real64 != 0;
  c->nmember_inv = 1.0f / nmember;         /* par:129 */
  c->nmember_1_inv = 1.0f / (nmember - 1); /* par:130 */
  return c;
}
extern "C" void or_destroy(or_ctx *c) { delete c; }

extern "C" int or_set_obs(or_ctx *c, int family, int type, int n, int nvar, const float *xyz,
                          const float *obs, const float *error, const float *hdxb,
                          const int32_t *qc) {
  ObsType *o;
  if (family == 0) {
    if (type < 1 || type > NUM_GTS) return fail("or_set_obs: bad gts type");
    o = &c->gts[type];
  } else if (family == 1) {
    if (type < 1 || type > NUM_RADAR) return fail("or_set_obs: bad radar type");
    o = &c->rad[type];
  } else
    return fail("or_set_obs: bad family");
  o->n = n;
  o->nvar = nvar;
  o->xyz = xyz;
  o->obs = obs;
  o->error = error;
  o->hdxb = hdxb;
  o->qc = qc;
  return 0;
}

static const or_type_config *find_cfg(const or_var_config *cfg, int family, int type, int *index) {
  for (int i = 0; i < cfg->ntypes; ++i)
    if (cfg->types[i].family == family && cfg->types[i].type == type) {
      if (index) *index = i;
      return &cfg->types[i];
    }
  return nullptr;
}

/* build_tree for one family (loc:35-167) */
static int build_family(or_ctx *c, const or_var_config *cfg, int family) {
  struct L { int obs_type, cfg_index; float hclr_inv, vclr_inv; };
  std::vector<L> list;
  float vclr_inv = 0.0f; /* the function-local that loc:151 tests (value of the LAST type) */
  const int ntypes_family = family == 0 ? (int)NUM_GTS : (int)NUM_RADAR;
  for (int obs_type = 1; obs_type <= ntypes_family; ++obs_type) {
    const ObsType &o = family == 0 ? c->gts[obs_type] : c->rad[obs_type];
    if (o.n <= 0) continue; /* loc:58,99 */
    if (family == 0) { /* loc:59-72 */
      if (!(obs_type == GTS_SYNOP || obs_type == GTS_METAR || obs_type == GTS_SHIPS ||
            obs_type == GTS_SOUND || obs_type == GTS_GPSPW))
        continue;
    }
    int ci = -1;
    const or_type_config *tc = find_cfg(cfg, family, obs_type, &ci);
    if (!tc) continue; /* namelist default use_it = .false. (cfg:8,29) */
    if (tc->use_it && tc->hclr > 0.0f) { /* loc:74,113 */
      const float hclr_inv = 1.0f / (tc->hclr * 1e3f);
      if (tc->vclr > 0.0f)
        vclr_inv = 1.0f / (tc->vclr * 1e3f);
      else
        vclr_inv = -1.0f;
      list.push_back({obs_type, ci, hclr_inv, vclr_inv});
    }
  }
  if (list.empty()) return 0; /* loc:91,130 */
  for (const L &cur : list) {
    const ObsType &o = family == 0 ? c->gts[cur.obs_type] : c->rad[cur.obs_type];
    std::vector<float> xyz(o.xyz, o.xyz + (size_t)3 * o.n); /* loc:148 */
    for (int i = 0; i < o.n; ++i) {                          /* loc:149 */
      xyz[(size_t)3 * i + 0] = xyz[(size_t)3 * i + 0] * cur.hclr_inv;
      xyz[(size_t)3 * i + 1] = xyz[(size_t)3 * i + 1] * cur.hclr_inv;
    }
    int dim;
    if (vclr_inv > 0.0f) { /* loc:151: NOT current%vclr_inv */
      dim = 3;
      for (int i = 0; i < o.n; ++i) xyz[(size_t)3 * i + 2] = xyz[(size_t)3 * i + 2] * cur.vclr_inv;
    } else {
      dim = 2;
      for (int i = 0; i < o.n; ++i) xyz[(size_t)3 * i + 2] = -1.0f;
    }
    /* get_lz picks the QUERY dimension per type from the namelist (loc:245-253,301-309).
     * A mismatch is an out-of-bounds read of qv in the reference: undefined, refused here. */
    const int qdim = cfg->types[cur.cfg_index].vclr > 0.0f ? 3 : 2;
    if (qdim != dim)
      return fail("build_tree: family mixes 2-D and 3-D localisation (reference behaviour undefined, SURVEY Q3)");
    TreeEntry te;
    te.family = family;
    te.mytype = cur.obs_type;
    te.cfg_index = cur.cfg_index;
    te.dim = dim;
    te.tree.reset(or_kd_create(xyz.data(), o.n, dim)); /* loc:160 */
    c->trees.push_back(std::move(te));
  }
  return (int)list.size();
}

extern "C" void or_destroy_trees(or_ctx *c) { /* loc:169-186 */
  c->trees.clear();
  c->n_gts_trees = 0;
}

extern "C" int or_build_trees(or_ctx *c, const or_var_config *cfg) { /* core:63-64 */
  or_destroy_trees(c);
  int a = build_family(c, cfg, 0);
  if (a < 0) return a;
  c->n_gts_trees = a;
  int b = build_family(c, cfg, 1);
  if (b < 0) return b;
  return a + b;
}

/* get_lz (loc:188-331) for both families; results per tree */
extern "C" int or_get_lz(or_ctx *c, const or_var_config *cfg, const float xyz[3], int stride,
                         int32_t *out_family, int32_t *out_type, int32_t *out_n, int32_t *out_idx,
                         float *out_r2) {
  const float r2 = lk_search_r2(); /* loc:202 */
  int any = 0;
  for (size_t t = 0; t < c->trees.size(); ++t) {
    const TreeEntry &te = c->trees[t];
    const or_type_config &tc = cfg->types[te.cfg_index];
    if (tc.max_lz_pts > stride) return fail("or_get_lz: stride < max_lz_pts");
    const float hclr_inv = 1.0f / (tc.hclr * 1e3f); /* loc:234,290 */
    float vclr_inv;
    if (tc.vclr > 0.0f)
      vclr_inv = 1.0f / (tc.vclr * 1e3f);
    else
      vclr_inv = -1.0f;
    float tmp[3];
    tmp[0] = xyz[0] * hclr_inv; /* loc:243,299 */
    tmp[1] = xyz[1] * hclr_inv;
    tmp[2] = 0.0f;
    if (vclr_inv > 0.0f) tmp[2] = xyz[2] * vclr_inv; /* loc:246,302 */
    int nlz = or_kd_r_nearest(te.tree.get(), tmp, r2, tc.max_lz_pts, out_idx + (size_t)t * stride,
                              out_r2 + (size_t)t * stride, nullptr);
    out_family[t] = te.family;
    out_type[t] = te.mytype;
    out_n[t] = nlz;
    if (nlz > 0) any = 1;
  }
  return any;
}

/* letkf_yoyb (core:300-595).  The linked list (core:558-578) is replaced by direct
 * appends to yo / yb: same order, same values. */
extern "C" int or_yoyb(or_ctx *c, const or_var_config *cfg, int ntrees, int stride,
                       const int32_t *fam, const int32_t *type, const int32_t *nn,
                       const int32_t *idxs, const float *r2s, int pmax, float *yo, float *yb) {
  const int k = c->k;
  std::vector<float> bg(k);
  int total = 0;
  for (int t = 0; t < ntrees; ++t) {
    if (nn[t] <= 0) continue; /* core:337,475: idx not allocated */
    const or_type_config *tc = find_cfg(cfg, fam[t], type[t], nullptr);
    if (!tc) return fail("or_yoyb: no config for type");
    if (fam[t] == 0) {
      const ObsType &o = c->gts[type[t]];
      const int nvar = o.nvar;
      bool is_assim[OR_MAX_SLOTS];
      for (int s = 0; s < nvar; ++s) is_assim[s] = tc->hclr > 0.0f ? (tc->is_assim[s] != 0) : false; /* core:355-363 */
      for (int j = 0; j < nn[t]; ++j) { /* core:421 */
        const int idx = idxs[(size_t)t * stride + j];
        const float r2 = r2s[(size_t)t * stride + j];
        for (int s = 0; s < nvar; ++s) { /* core:428 */
          if (!is_assim[s]) continue;
          bool anyqc = false; /* any(qc(k,idx,:) >= 0) core:429 */
          for (int m = 0; m < k; ++m)
            if (o.qc[(size_t)s + (size_t)nvar * ((idx - 1) + (size_t)o.n * m)] >= 0) {
              anyqc = true;
              break;
            }
          if (!anyqc) continue;
          float sum = 0.0f;
          for (int m = 0; m < k; ++m) {
            bg[m] = o.hdxb[(size_t)s + (size_t)nvar * ((idx - 1) + (size_t)o.n * m)];
            sum = sum + bg[m];
          }
          const float mean = sum * c->nmember_inv; /* core:431 */
          float dot = 0.0f;
          for (int m = 0; m < k; ++m) {
            bg[m] = bg[m] - mean; /* core:432 */
            dot = dot + bg[m] * bg[m];
          }
          float omm = o.obs[(size_t)s + (size_t)nvar * (idx - 1)] - mean; /* core:433 */
          const float std_ = sqrtf(dot * c->nmember_1_inv);               /* core:434 */
          const float err = o.error[(size_t)s + (size_t)nvar * (idx - 1)] * tc->err_muti[s]; /* core:435 */
          if (fabsf(omm) > sqrtf(std_ * std_ + err * err) * tc->err_rej[s]) continue; /* core:437 */
          const float error_inv = lk_error_inv(err, r2, cfg->weight_function); /* core:443-450 */
          omm = omm * error_inv;
          if (total >= pmax) return fail("or_yoyb: pmax too small");
          yo[total] = omm;
          for (int m = 0; m < k; ++m) yb[(size_t)total * k + m] = bg[m] * error_inv; /* core:452 */
          total++;
        }
      }
    } else {
      const ObsType &o = c->rad[type[t]];
      const bool is_assim = tc->hclr > 0.0f; /* core:487 */
      const float err = tc->err_muti[0];    /* core:488,502 */
      const float err_rej = tc->err_rej[0];
      if (!is_assim) continue;
      for (int j = 0; j < nn[t]; ++j) { /* core:492 */
        const int idx = idxs[(size_t)t * stride + j];
        const float r2 = r2s[(size_t)t * stride + j];
        float sum = 0.0f;
        for (int m = 0; m < k; ++m) {
          bg[m] = o.hdxb[(size_t)(idx - 1) + (size_t)o.n * m]; /* core:497 */
          sum = sum + bg[m];
        }
        const float mean = sum * c->nmember_inv;
        float dot = 0.0f;
        for (int m = 0; m < k; ++m) {
          bg[m] = bg[m] - mean;
          dot = dot + bg[m] * bg[m];
        }
        const float ob = o.obs[idx - 1];
        float omm = ob - mean;
        const float std_ = sqrtf(dot * c->nmember_1_inv);
        const bool gross = fabsf(omm) > sqrtf(std_ * std_ + err * err) * err_rej;
        if (type[t] == RAD_DBZ) { /* core:504-507 */
          if (gross && ob != cfg->norain_value) continue;
          if (ob == cfg->norain_value && mean == cfg->norain_value) continue;
        } else {
          if (gross) continue; /* core:509 */
        }
        const float error_inv = lk_error_inv(err, r2, cfg->weight_function); /* core:516-523 */
        omm = omm * error_inv;
        if (total >= pmax) return fail("or_yoyb: pmax too small");
        yo[total] = omm;
        for (int m = 0; m < k; ++m) yb[(size_t)total * k + m] = bg[m] * error_inv;
        total++;
      }
    }
  }
  return total;
}

/* set_optimal_workspace_for_eigen (eig:16-35) */
static void eigen_ws_init(EigenWs &w, int n, bool real64) {
  if (w.n == n && w.prec == (int)real64) return;
  w.n = n;
  w.prec = (int)real64;
  int info = 0, m1 = -1;
  int iq = 0;
  if (real64) {
    w.eval_d.assign(n, 0.0);
    w.evect_d.assign((size_t)n * n, 0.0);
    double q = 0;
    scipy_dsyevd_("V", "L", &n, w.evect_d.data(), &n, w.eval_d.data(), &q, &m1, &iq, &m1, &info, 1, 1);
    w.lwork = (int)q;
    w.liwork = iq;
    w.work_d.assign(w.lwork, 0.0);
  } else {
    w.eval_s.assign(n, 0.0f);
    w.evect_s.assign((size_t)n * n, 0.0f);
    float q = 0;
    scipy_ssyevd_("V", "L", &n, w.evect_s.data(), &n, w.eval_s.data(), &q, &m1, &iq, &m1, &info, 1, 1);
    w.lwork = (int)q;
    w.liwork = iq;
    w.work_s.assign(w.lwork, 0.0f);
  }
  w.iwork.assign(w.liwork, 0);
}
static thread_local EigenWs tl_ws;

template <typename T> struct Blas;
template <> struct Blas<double> {
  static void syrk(int n, int kk, double alpha, const double *a, double beta, double *c) {
    scipy_dsyrk_("L", "N", &n, &kk, &alpha, a, &n, &beta, c, &n, 1, 1);
  }
  static void syevd(EigenWs &w, int n, int *info) {
    scipy_dsyevd_("V", "L", &n, w.evect_d.data(), &n, w.eval_d.data(), w.work_d.data(), &w.lwork,
                  w.iwork.data(), &w.liwork, info, 1, 1);
  }
  static void gemm_nt(int n, const double *a, const double *b, double *c) {
    const double one = 1, zero = 0;
    scipy_dgemm_("N", "T", &n, &n, &n, &one, a, &n, b, &n, &zero, c, &n, 1, 1);
  }
  static void gemv(const char *tr, int m, int n, double alpha, const double *a, const double *x,
                   double beta, double *y) {
    const int one = 1;
    scipy_dgemv_(tr, &m, &n, &alpha, a, &m, x, &one, &beta, y, &one, 1);
  }
  static void symv(int n, const double *a, const double *x, double *y) {
    const double one = 1, zero = 0;
    const int i1 = 1;
    scipy_dsymv_("L", &n, &one, a, &n, x, &i1, &zero, y, &i1, 1);
  }
  static double *eval(EigenWs &w) { return w.eval_d.data(); }
  static double *evect(EigenWs &w) { return w.evect_d.data(); }
};
template <> struct Blas<float> {
  static void syrk(int n, int kk, float alpha, const float *a, float beta, float *c) {
    scipy_ssyrk_("L", "N", &n, &kk, &alpha, a, &n, &beta, c, &n, 1, 1);
  }
  static void syevd(EigenWs &w, int n, int *info) {
    scipy_ssyevd_("V", "L", &n, w.evect_s.data(), &n, w.eval_s.data(), w.work_s.data(), &w.lwork,
                  w.iwork.data(), &w.liwork, info, 1, 1);
  }
  static void gemm_nt(int n, const float *a, const float *b, float *c) {
    const float one = 1, zero = 0;
    scipy_sgemm_("N", "T", &n, &n, &n, &one, a, &n, b, &n, &zero, c, &n, 1, 1);
  }
  static void gemv(const char *tr, int m, int n, float alpha, const float *a, const float *x,
                   float beta, float *y) {
    const int one = 1;
    scipy_sgemv_(tr, &m, &n, &alpha, a, &m, x, &one, &beta, y, &one, 1);
  }
  static void symv(int n, const float *a, const float *x, float *y) {
    const float one = 1, zero = 0;
    const int i1 = 1;
    scipy_ssymv_("L", &n, &one, a, &n, x, &i1, &zero, y, &i1, 1);
  }
  static float *eval(EigenWs &w) { return w.eval_s.data(); }
  static float *evect(EigenWs &w) { return w.evect_s.data(); }
};

/* letkf_solve (core:598-700) with inverse_matrix (eig:37-76) and sqrt_matrix (eig:78-108).
 * T = double is the REAL64 branch, T = float the default-real branch. */
template <typename T>
static void solve_impl(or_ctx *c, const float *xb, int nobs, const float *yo, const float *yb,
                       float inflat, bool use_rtpp, float rtpp_alpha, bool use_rtps,
                       float rtps_alpha, float *xa, double *wbar_out, double *Wa_out,
                       double *xa_raw) {
  const int n = c->k;
  EigenWs &ws = tl_ws;
  eigen_ws_init(ws, n, sizeof(T) == 8);
  std::vector<T> identity((size_t)n * n, T(0)), tmp((size_t)n * n), w((size_t)n * n),
      wbar2d((size_t)n * n), wbar(n), xb_prime(n);
  for (int i = 0; i < n; ++i) identity[(size_t)i * n + i] = T(1); /* core:628-637 */
  std::vector<T> yb_r((size_t)n * nobs), yo_r(nobs);             /* core:642-647 */
  for (size_t i = 0; i < (size_t)n * nobs; ++i) yb_r[i] = (T)yb[i];
  for (int i = 0; i < nobs; ++i) yo_r[i] = (T)yo[i];
  const T inflat_r = (T)inflat;
  Blas<T>::syrk(n, nobs, T(1), yb_r.data(), inflat_r, identity.data()); /* core:649/656 */
  /* inverse_matrix (eig:37-76) */
  T *evect = Blas<T>::evect(ws), *eval = Blas<T>::eval(ws);
  std::memcpy(evect, identity.data(), sizeof(T) * n * n); /* ?copy eig:48/65 */
  int info = 0;
  Blas<T>::syevd(ws, n, &info); /* eig:49/66; info ignored by the reference */
  for (int i = 0; i < n; ++i) { /* eig:51-54 */
    eval[i] = T(1) / eval[i];
    for (int r = 0; r < n; ++r) tmp[(size_t)i * n + r] = evect[(size_t)i * n + r] * eval[i];
  }
  Blas<T>::gemm_nt(n, tmp.data(), evect, identity.data());                      /* eig:56/73 */
  Blas<T>::gemv("N", n, nobs, T(1), yb_r.data(), yo_r.data(), T(0), wbar.data()); /* core:651 */
  Blas<T>::symv(n, identity.data(), wbar.data(), xb_prime.data());              /* core:652 */
  for (int j = 0; j < n; ++j) /* wbar2d = spread(xb_prime, 2, nmember) core:662 */
    for (int i = 0; i < n; ++i) wbar2d[(size_t)j * n + i] = xb_prime[i];
  if (wbar_out)
    for (int i = 0; i < n; ++i) wbar_out[i] = (double)xb_prime[i];
  /* sqrt_matrix (eig:78-108): eval already holds 1/lambda */
  for (int i = 0; i < n; ++i) {
    const T s = std::sqrt(eval[i]);
    for (int r = 0; r < n; ++r) tmp[(size_t)i * n + r] = evect[(size_t)i * n + r] * s;
  }
  Blas<T>::gemm_nt(n, tmp.data(), evect, w.data());
  const T sk = sizeof(T) == 8 ? (T)std::sqrt((double)(n - 1)) : (T)sqrtf((float)(n - 1));
  for (size_t i = 0; i < (size_t)n * n; ++i) wbar2d[i] = wbar2d[i] + sk * w[i]; /* ?axpy core:666/668 */
  if (Wa_out)
    for (size_t i = 0; i < (size_t)n * n; ++i) Wa_out[i] = (double)(sk * w[i]);
  float xsum = 0.0f; /* xb_mean = sum(xb) * nmember_inv : real32 expression (core:671) */
  for (int i = 0; i < n; ++i) xsum = xsum + xb[i];
  const T xb_mean = (T)(xsum * c->nmember_inv);
  for (int i = 0; i < n; ++i) xb_prime[i] = (T)xb[i] - xb_mean; /* core:672 */
  for (int i = 0; i < n; ++i) wbar[i] = xb_mean;                /* core:673 */
  Blas<T>::gemv("T", n, n, T(1), wbar2d.data(), xb_prime.data(), T(1), wbar.data()); /* core:675/677 */
  for (int i = 0; i < n; ++i) xa[i] = (float)wbar[i];                                /* core:679 */
  if (xa_raw)
    for (int i = 0; i < n; ++i) xa_raw[i] = (double)wbar[i];
  if (use_rtpp || use_rtps) { /* core:684-698 */
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s = s + xa[i];
    const float xa_mean = s * c->nmember_inv;
    std::vector<float> xa_prime(n);
    for (int i = 0; i < n; ++i) xa_prime[i] = xa[i] - xa_mean;
    if (use_rtpp) { /* mixed real32 * T expression, assigned to real32 (core:689) */
      const float oma = 1.0f - rtpp_alpha;
      for (int i = 0; i < n; ++i)
        xa_prime[i] = (float)((T)(oma * xa_prime[i]) + (T)rtpp_alpha * xb_prime[i]);
    }
    if (use_rtps) { /* core:692-694 */
      T d = T(0);
      for (int i = 0; i < n; ++i) d = d + xb_prime[i] * xb_prime[i];
      const float xb_std = (float)d;
      float xa_std = 0.0f;
      for (int i = 0; i < n; ++i) xa_std = xa_std + xa_prime[i] * xa_prime[i];
      const float f = rtps_alpha * sqrtf(xb_std / xa_std) - rtps_alpha + 1.0f;
      for (int i = 0; i < n; ++i) xa_prime[i] = xa_prime[i] * f;
    }
    for (int i = 0; i < n; ++i) xa[i] = xa_mean + xa_prime[i]; /* core:697 */
  }
}

extern "C" int or_solve(or_ctx *c, const float *xb, int p, const float *yo, const float *yb,
                        float inflat, int use_rtpp, float rtpp_alpha, int use_rtps,
                        float rtps_alpha, float *xa, double *wbar, double *Wa, double *xa_raw) {
  if (p <= 0) return fail("or_solve: p <= 0");
  if (c->real64)
    solve_impl<double>(c, xb, p, yo, yb, inflat, use_rtpp != 0, rtpp_alpha, use_rtps != 0,
                       rtps_alpha, xa, wbar, Wa, xa_raw);
  else
    solve_impl<float>(c, xb, p, yo, yb, inflat, use_rtpp != 0, rtpp_alpha, use_rtps != 0,
                      rtps_alpha, xa, wbar, Wa, xa_raw);
  return 0;
}

/* the grid-point loop (core:209-240) */
extern "C" int or_analyze(or_ctx *c, const or_var_config *cfg, int64_t npts, const float *xyz_grid,
                          int nfields, float *var, int nthreads, int64_t *np_out,
                          int64_t *rows_out) {
  const int k = c->k;
  const int ntrees = (int)c->trees.size();
  int stride = 1, pmax = 0;
  for (const TreeEntry &te : c->trees) {
    const or_type_config &tc = cfg->types[te.cfg_index];
    stride = std::max(stride, tc.max_lz_pts);
    pmax += tc.max_lz_pts * (te.family == 0 ? tc.nvar : 1);
  }
  const float inflat = (float)(k - 1) / cfg->multi_infl; /* core:68 */
  int64_t np = 0, rows = 0;
  int err = 0;
  std::string errmsg;
  if (ntrees == 0) { /* core:66 */
    if (np_out) *np_out = 0;
    if (rows_out) *rows_out = 0;
    return 0;
  }
#pragma omp parallel num_threads(nthreads) reduction(+ : np, rows)
  {
    std::vector<int32_t> fam(ntrees), typ(ntrees), nn(ntrees), idx((size_t)ntrees * stride);
    std::vector<float> r2((size_t)ntrees * stride), yo(std::max(pmax, 1)),
        yb((size_t)std::max(pmax, 1) * k), xb(k), xa(k);
#pragma omp for schedule(dynamic, 64)
    for (int64_t pt = 0; pt < npts; ++pt) {
      if (err) continue;
      const float xyz[3] = {xyz_grid[3 * pt], xyz_grid[3 * pt + 1], xyz_grid[3 * pt + 2]};
      int any = or_get_lz(c, cfg, xyz, stride, fam.data(), typ.data(), nn.data(), idx.data(), r2.data());
      if (any < 0) {
#pragma omp critical
        { err = 1; errmsg = g_err; }
        continue;
      }
      if (!any) continue; /* core:220 */
      int p = or_yoyb(c, cfg, ntrees, stride, fam.data(), typ.data(), nn.data(), idx.data(),
                      r2.data(), pmax, yo.data(), yb.data());
      if (p < 0) {
#pragma omp critical
        { err = 1; errmsg = g_err; }
        continue;
      }
      if (p == 0) continue; /* core:226 */
      np += 1;
      rows += p;
      for (int f = 0; f < nfields; ++f) {
        float *v = var + (size_t)f * npts * k;
        for (int m = 0; m < k; ++m) xb[m] = v[(size_t)m * npts + pt]; /* core:228 */
        or_solve(c, xb.data(), p, yo.data(), yb.data(), inflat, cfg->use_rtpp, cfg->rtpp_alpha,
                 cfg->use_rtps, cfg->rtps_alpha, xa.data(), nullptr, nullptr, nullptr);
        for (int m = 0; m < k; ++m) v[(size_t)m * npts + pt] = xa[m]; /* core:229 */
      }
    }
  }
  if (err) return fail(errmsg);
  if (np_out) *np_out = np;
  if (rows_out) *rows_out = rows;
  return 0;
}

/* letkf_tune_q (core:702-733) */
extern "C" void or_tune_q(int k, int64_t npts, float *q) {
  std::vector<float> var(k);
  for (int64_t pt = 0; pt < npts; ++pt) {
    float s_all = 0.0f, s_pos = 0.0f;
    for (int m = 0; m < k; ++m) {
      var[m] = q[(size_t)m * npts + pt];
      s_all = s_all + var[m];
      if (var[m] > 0.0f) s_pos = s_pos + var[m];
    }
    const float ratio = s_all / s_pos; /* core:719; 0/0 = NaN when all members are 0 */
    for (int m = 0; m < k; ++m) {
      if (var[m] < 0.0f)
        var[m] = 0.0f;
      else
        var[m] = ratio * var[m];
      q[(size_t)m * npts + pt] = var[m];
    }
  }
}

/* BASELINE config E: LAPACK ?syevd('V','L') per matrix (eig:49/66) */
extern "C" int or_syevd_batch(int k, int64_t batch, int real64, const void *A, void *W, void *V,
                              int nthreads) {
  int bad = 0;
#pragma omp parallel num_threads(nthreads)
  {
    EigenWs ws;
    eigen_ws_init(ws, k, real64 != 0);
#pragma omp for schedule(dynamic, 16)
    for (int64_t b = 0; b < batch; ++b) {
      int info = 0;
      if (real64) {
        std::memcpy(ws.evect_d.data(), (const double *)A + (size_t)b * k * k, sizeof(double) * k * k);
        Blas<double>::syevd(ws, k, &info);
        std::memcpy((double *)W + (size_t)b * k, ws.eval_d.data(), sizeof(double) * k);
        std::memcpy((double *)V + (size_t)b * k * k, ws.evect_d.data(), sizeof(double) * k * k);
      } else {
        std::memcpy(ws.evect_s.data(), (const float *)A + (size_t)b * k * k, sizeof(float) * k * k);
        Blas<float>::syevd(ws, k, &info);
        std::memcpy((float *)W + (size_t)b * k, ws.eval_s.data(), sizeof(float) * k);
        std::memcpy((float *)V + (size_t)b * k * k, ws.evect_s.data(), sizeof(float) * k * k);
      }
      if (info != 0) {
#pragma omp atomic write
        bad = 1;
      }
    }
  }
  return bad ? fail("syevd: info != 0") : 0;
}

extern "C" float or_gc1999(void) { return lk_gc1999(); }
extern "C" float or_search_r2(void) { return lk_search_r2(); }
extern "C" float or_gaspari_cohn(float x) { return lk_gaspari_cohn(x); }
extern "C" float or_expf(float x) { return lk_expf(x); }
extern "C" float or_error_inv(float err, float r2, int wf) { return lk_error_inv(err, r2, wf); }
