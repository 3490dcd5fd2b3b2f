// C ABI of libletkf_b200.so (include/letkf_b200.h): context, observation upload, tree cache and
// the per-variable pipeline  search -> count/compact -> localise+Gram -> eigen -> transform.
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "letkf_internal.cuh"
#include "xform32.cuh"

using namespace lk;

static thread_local std::string g_last_error;

// reference enums (module_param.f90:28-57,93-97)
enum { GTS_SOUND = 1, GTS_SYNOP = 2, GTS_GPSPW = 8, GTS_METAR = 10, GTS_SHIPS = 11, NUM_GTS = 29 };
enum { RAD_DBZ = 1, NUM_RADAR = 4 };

struct ActiveTree {
  ObsDev *obs;
  const letkf_b200_type_config *tc;
  DevTree *tree;
  int dim;
};

struct letkf_b200_ctx {
  int k = 0;
  bool real64 = true;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  std::map<std::pair<int, int>, std::unique_ptr<ObsDev>> obs;  // (family, type), iteration = reference order
  int64_t chunk_override = 0;
  int nz_hint = 1;             // letkf_b200_set_levels: points p and p + l*(npts/nz) share (x, y)
  bool force_generic = false;  // LETKF_B200_GENERIC=1: use the generic block-per-unit kernels for every k
  // per-call device staging for the host-pointer entry points
  DevBuf<float> d_xyz, d_var;
  // double-buffered slabs + copy streams of the pipelined host-pointer analyze
  DevBuf<float> slab_xyz[2], slab_var[2];
  cudaStream_t cs_in = nullptr, cs_out = nullptr;
  cudaEvent_t ev_in[2] = {}, ev_done[2] = {}, ev_out[2] = {};
  // per-chunk scratch
  DevBuf<int32_t> cnt[LETKF_B200_MAX_TYPES], idx[LETKF_B200_MAX_TYPES];
  DevBuf<float> r2[LETKF_B200_MAX_TYPES];
  DevBuf<int32_t> p, unit_pt, nanflag, counters;  // counters: [0]=nunits [1]=sweeps max
  DevBuf<int64_t> rows_sum;
  DevBuf<unsigned char> cub_tmp;
  DevBuf<unsigned char> C, b, lam, wbar;  // sized in bytes for the working precision
  std::vector<cudaEvent_t> io_ev;         // chunk-granular host IO: [2*i] upload done, [2*i+1] chunk analysed
  CtxShared shared;                       // launch counter, eigensolver scratch (bound per call, CtxBind)
  // host-sync-free chunk loop of the FP64 path: per-chunk unit / row counts stay on the device, stage events per chunk
  DevBuf<int32_t> chunk_cnt;
  DevBuf<int64_t> chunk_rows;
  std::vector<cudaEvent_t> st_ev;         // 4 per chunk: start, after search + compaction, after Gram, after solve
  // FP64 solve = Householder tridiagonalisation + pole expansion of C^(-1/2) (fcn_common.cuh); the
  // Jacobi eigensolver path stays selectable (LETKF_B200_SOLVER=jacobi) and serves the FP32 build
  bool use_fcn = true;
  DevBuf<double> poles;
};

// binds the context's shared state to the calling thread for one C-ABI call
struct CtxBind {
  CtxShared *prev;
  explicit CtxBind(letkf_b200_ctx *c) : prev(current_ctx_shared()) { current_ctx_shared() = c ? &c->shared : nullptr; }
  ~CtxBind() { current_ctx_shared() = prev; }
};

template <typename F>
static int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_last_error = e.what();
    cudaGetLastError();  // clear a sticky-less error state
    return 1;
  }
}

template <typename F>
static int guarded(letkf_b200_ctx *c, F &&f) {
  CtxBind bind(c);
  const int rc = guarded(std::forward<F>(f));
  if (rc != 0 && c) {
    // a failed call must not leave copies in flight on the caller's arrays (the Fortran shim stops the
    // program and may deallocate them): drain the compute and the two copy streams before returning
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->cs_in) cudaStreamSynchronize(c->cs_in);
    if (c->cs_out) cudaStreamSynchronize(c->cs_out);
    cudaGetLastError();
  }
  return rc;
}

extern "C" const char *letkf_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int letkf_b200_version(void) { return 100; }

extern "C" int letkf_b200_init(letkf_b200_ctx **out, int nmember, int real64, int device) {
  return guarded([&] {
    LK_REQUIRE(out != nullptr, "letkf_b200_init: null ctx pointer");
    LK_REQUIRE(nmember >= 2 && nmember <= LETKF_B200_MAX_MEMBERS, "letkf_b200_init: need 2 <= nmember <= 256");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(std::string("letkf_b200_init: no CUDA device (") + cudaGetErrorString(e) +
                  "); this library has no CPU path");
    LK_REQUIRE(device >= 0 && device < ndev, "letkf_b200_init: bad device ordinal");
    LK_CUDA(cudaSetDevice(device));
    auto c = std::make_unique<letkf_b200_ctx>();
    c->k = nmember;
    c->real64 = real64 != 0;
    c->device = device;
    const char *fg = getenv("LETKF_B200_GENERIC");
    c->force_generic = fg && fg[0] == '1';
    const char *sv = getenv("LETKF_B200_SOLVER");
    c->use_fcn = c->real64 && !(sv && std::string(sv) == "jacobi");
    LK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &ev : c->ev) LK_CUDA(cudaEventCreate(&ev));
    if (c->use_fcn) {
      const std::vector<double> &tab = fcn_pole_table_host();
      c->poles.ensure(tab.size());
      LK_CUDA(cudaMemcpy(c->poles.p, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    }
    *out = c.release();
  });
}

extern "C" int letkf_b200_finalize(letkf_b200_ctx *c) {
  return guarded(c, [&] {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &ev : c->ev)
      if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    for (auto &ev : c->io_ev) cudaEventDestroy(ev);
    for (auto &ev : c->st_ev) cudaEventDestroy(ev);
    if (c->cs_in) {
      cudaStreamDestroy(c->cs_in);
      cudaStreamDestroy(c->cs_out);
      for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c->ev_in[i]);
        cudaEventDestroy(c->ev_done[i]);
        cudaEventDestroy(c->ev_out[i]);
      }
    }
    delete c;
  });
}

extern "C" int64_t letkf_b200_launch_count(letkf_b200_ctx *c) { return c ? c->shared.launches : 0; }
extern "C" void *letkf_b200_stream(letkf_b200_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int letkf_b200_set_levels(letkf_b200_ctx *c, int nz) {
  if (!c || nz < 1) return 1;
  c->nz_hint = nz;
  return 0;
}
extern "C" int letkf_b200_set_chunk(letkf_b200_ctx *c, int64_t n) {
  if (!c || n < 0) return 1;
  c->chunk_override = n;
  return 0;
}

// ---- observations ------------------------------------------------------------------------------
static void set_obs_impl(letkf_b200_ctx *c, int family, int type, int n, int nvar, const float *xyz,
                         const float *obs, const float *error, const float *hdxb, const int32_t *qc,
                         bool on_device) {
  LK_REQUIRE(c, "null context");
  LK_CUDA(cudaSetDevice(c->device));
  LK_REQUIRE(family == LETKF_B200_GTS || family == LETKF_B200_RADAR, "set_obs: bad family");
  const bool gts = family == LETKF_B200_GTS;
  LK_REQUIRE(type >= 1 && type <= (gts ? NUM_GTS : NUM_RADAR), "set_obs: bad type");
  LK_REQUIRE(n >= 0 && nvar >= 1 && nvar <= LETKF_B200_MAX_SLOTS, "set_obs: bad n / nvar");
  LK_REQUIRE(gts || nvar == 1, "set_obs: radar types have one slot");
  const auto key = std::make_pair(family, type);
  if (n == 0) {
    c->obs.erase(key);
    return;
  }
  LK_REQUIRE(xyz && obs && hdxb, "set_obs: null array");
  LK_REQUIRE(!gts || (error && qc), "set_obs: gts needs error and qc");
  auto o = std::make_unique<ObsDev>();
  o->family = family;
  o->type = type;
  o->n = n;
  o->nvar = nvar;
  const size_t nn = (size_t)n * nvar, k = c->k;
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  o->h_xyz.resize((size_t)3 * n);
  LK_CUDA(cudaMemcpyAsync(o->h_xyz.data(), xyz, sizeof(float) * 3 * n,
                          on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost, c->stream));
  o->obs.ensure(nn);
  LK_CUDA(cudaMemcpyAsync(o->obs.p, obs, sizeof(float) * nn, kind, c->stream));
  if (gts) {
    o->error.ensure(nn);
    LK_CUDA(cudaMemcpyAsync(o->error.p, error, sizeof(float) * nn, kind, c->stream));
  }
  o->pert.ensure(nn * k);
  o->mean.ensure(nn);
  o->stdv.ensure(nn);
  o->anyqc.ensure(nn);
  o->err.ensure(nn);
  o->omm.ensure(nn);
  o->pass.ensure(nn);
  // hdxb / qc are only needed for the one-off statistics: stage, reduce, free
  DevBuf<float> d_h;
  DevBuf<int32_t> d_q;
  const float *hp = hdxb;
  const int32_t *qp = qc;
  if (!on_device) {
    d_h.ensure(nn * k);
    LK_CUDA(cudaMemcpyAsync(d_h.p, hdxb, sizeof(float) * nn * k, kind, c->stream));
    hp = d_h.p;
    if (gts) {
      d_q.ensure(nn * k);
      LK_CUDA(cudaMemcpyAsync(d_q.p, qc, sizeof(int32_t) * nn * k, kind, c->stream));
      qp = d_q.p;
    }
  }
  launch_obs_static(c->stream, c->k, n, nvar, gts, hp, qp, o->pert.p, o->mean.p, o->stdv.p, o->anyqc.p);
  LK_CUDA(cudaStreamSynchronize(c->stream));
  c->obs[key] = std::move(o);
}

extern "C" int letkf_b200_set_obs(letkf_b200_ctx *c, int family, int type, int n, int nvar, const float *xyz,
                                  const float *obs, const float *error, const float *hdxb,
                                  const int32_t *qc) {
  return guarded(c, [&] { set_obs_impl(c, family, type, n, nvar, xyz, obs, error, hdxb, qc, false); });
}
extern "C" int letkf_b200_set_obs_dev(letkf_b200_ctx *c, int family, int type, int n, int nvar,
                                      const float *xyz, const float *obs, const float *error,
                                      const float *hdxb, const int32_t *qc) {
  return guarded(c, [&] { set_obs_impl(c, family, type, n, nvar, xyz, obs, error, hdxb, qc, true); });
}
extern "C" int letkf_b200_clear_obs(letkf_b200_ctx *c) {
  return guarded(c, [&] {
    LK_REQUIRE(c, "null context");
    LK_CUDA(cudaSetDevice(c->device));
    LK_CUDA(cudaStreamSynchronize(c->stream));
    c->obs.clear();
  });
}

// ---- build_tree (module_localization.f90:35-167) -------------------------------------------------
static DevTree *get_tree(letkf_b200_ctx *c, ObsDev &o, float hclr, float vclr, int dim) {
  const auto key = std::make_pair(hclr, dim == 3 ? vclr : -1.0f);
  auto it = o.trees.find(key);
  if (it != o.trees.end()) return it->second.get();
  const float hinv = lk_clr_inv(hclr);                      // loc:76,115
  const float vinv = dim == 3 ? lk_clr_inv(vclr) : -1.0f;   // loc:78-82,117-121
  std::vector<float> xyz(o.h_xyz);
  for (int i = 0; i < o.n; ++i) {  // loc:149-157
    xyz[(size_t)3 * i + 0] = LK_MUL(xyz[(size_t)3 * i + 0], hinv);
    xyz[(size_t)3 * i + 1] = LK_MUL(xyz[(size_t)3 * i + 1], hinv);
    xyz[(size_t)3 * i + 2] = dim == 3 ? LK_MUL(xyz[(size_t)3 * i + 2], vinv) : -1.0f;
  }
  HostTree ht;
  build_kdtree_host(xyz.data(), o.n, dim, ht);
  auto dt = std::make_unique<DevTree>();
  dt->dim = dim;
  dt->n = o.n;
  dt->nnodes = (int)ht.nodes.size();
  dt->hclr = hclr;
  dt->vclr = vclr;
  dt->nodes.ensure(ht.nodes.size());
  dt->pts.ensure(ht.pts.size());
  LK_CUDA(cudaMemcpyAsync(dt->nodes.p, ht.nodes.data(), sizeof(KdNodeDev) * ht.nodes.size(),
                          cudaMemcpyHostToDevice, c->stream));
  LK_CUDA(cudaMemcpyAsync(dt->pts.p, ht.pts.data(), sizeof(float4) * ht.pts.size(), cudaMemcpyHostToDevice,
                          c->stream));
  LK_CUDA(cudaStreamSynchronize(c->stream));
  DevTree *raw = dt.get();
  o.trees[key] = std::move(dt);
  return raw;
}

static const letkf_b200_type_config *find_cfg(const letkf_b200_var_config *cfg, int family, int type) {
  for (int i = 0; i < cfg->ntypes; ++i)
    if (cfg->types[i].family == family && cfg->types[i].type == type) return &cfg->types[i];
  return nullptr;
}

// Active observation types of this variable in the reference's visiting order (gts then radar,
// ascending enum), with their trees built / fetched and the per-variable QC arrays refreshed.
static std::vector<ActiveTree> prepare_trees(letkf_b200_ctx *c, const letkf_b200_var_config *cfg) {
  LK_REQUIRE(cfg && cfg->ntypes >= 0 && cfg->ntypes <= LETKF_B200_MAX_TYPES, "bad var_config");
  std::vector<ActiveTree> act;
  for (int family = 0; family < 2; ++family) {
    std::vector<ActiveTree> fam;
    bool last_3d = false;
    for (auto &kv : c->obs) {
      if (kv.first.first != family) continue;
      ObsDev &o = *kv.second;
      if (o.n <= 0) continue;  // loc:58,99
      if (family == LETKF_B200_GTS &&
          !(o.type == GTS_SYNOP || o.type == GTS_METAR || o.type == GTS_SHIPS || o.type == GTS_SOUND ||
            o.type == GTS_GPSPW))
        continue;  // loc:59-72
      const letkf_b200_type_config *tc = find_cfg(cfg, family, o.type);
      if (!tc) continue;
      if (!(tc->use_it && tc->hclr > 0.0f)) continue;  // loc:74,113
      LK_REQUIRE(tc->max_lz_pts >= 1, "max_lz_pts must be >= 1");
      LK_REQUIRE(tc->nvar == o.nvar, "var_config nvar does not match the observation set");
      last_3d = tc->vclr > 0.0f;
      fam.push_back({&o, tc, nullptr, 0});
    }
    for (auto &a : fam) {
      // The reference sizes every tree of a family with the LAST type's vclr (loc:151) but
      // queries each with its own (loc:245,301); a mismatch reads past the query vector.
      LK_REQUIRE((a.tc->vclr > 0.0f) == last_3d,
                 "family mixes 2-D and 3-D localisation (reference behaviour undefined, SURVEY Q3)");
      a.dim = last_3d ? 3 : 2;
      a.tree = get_tree(c, *a.obs, a.tc->hclr, a.tc->vclr, a.dim);
      const bool gts = family == LETKF_B200_GTS;
      launch_obs_config(c->stream, a.obs->n, a.obs->nvar, gts, !gts && a.obs->type == RAD_DBZ, a.obs->obs.p,
                        a.obs->error.p, a.obs->mean.p, a.obs->stdv.p, a.obs->anyqc.p, *a.tc,
                        cfg->norain_value, a.obs->err.p, a.obs->omm.p, a.obs->pass.p);
      act.push_back(a);
    }
  }
  LK_REQUIRE(act.size() <= LETKF_B200_MAX_TYPES, "too many active trees");
  return act;
}

static TreeViews make_views(letkf_b200_ctx *c, const letkf_b200_var_config *cfg,
                            const std::vector<ActiveTree> &act, int64_t chunk) {
  TreeViews tv;
  std::memset(&tv, 0, sizeof(tv));
  tv.ntrees = (int)act.size();
  tv.weight_function = cfg->weight_function;
  for (int t = 0; t < tv.ntrees; ++t) {
    const ActiveTree &a = act[t];
    TreeView &v = tv.t[t];
    v.nodes = a.tree->nodes.p;
    v.pts = a.tree->pts.p;
    v.dim = a.dim;
    v.hinv = lk_clr_inv(a.tc->hclr);                        // loc:234,290
    v.vinv = a.dim == 3 ? lk_clr_inv(a.tc->vclr) : -1.0f;   // loc:236-240,292-296
    v.nalloc = a.tc->max_lz_pts;
    v.nvar = a.obs->nvar;
    v.nact = 0;
    for (int s = 0; s < a.obs->nvar; ++s) {
      const bool on = a.obs->family == LETKF_B200_GTS ? a.tc->is_assim[s] != 0 : true;  // core:355-363,487
      if (on) v.act[v.nact++] = s;
    }
    v.pert = a.obs->pert.p;
    v.omm = a.obs->omm.p;
    v.err = a.obs->err.p;
    v.pass = a.obs->pass.p;
    c->cnt[t].ensure(chunk);
    c->idx[t].ensure((size_t)chunk * v.nalloc);
    c->r2[t].ensure((size_t)chunk * v.nalloc);
    v.cnt = c->cnt[t].p;
    v.idx = c->idx[t].p;
    v.r2 = c->r2[t].p;
    v.family = a.obs->family;
    v.type = a.obs->type;
  }
  return tv;
}

static int64_t pick_chunk(letkf_b200_ctx *c, const std::vector<ActiveTree> &act, int64_t npts, size_t tsize) {
  if (c->chunk_override > 0) return std::min<int64_t>(c->chunk_override, std::max<int64_t>(npts, 1));
  size_t per_pt = 16 + ((size_t)c->k * c->k + 3 * (size_t)c->k) * tsize;
  for (auto &a : act) per_pt += 4 + 8 * (size_t)a.tc->max_lz_pts;
  size_t free_b = 0, total_b = 0;
  LK_CUDA(cudaMemGetInfo(&free_b, &total_b));
  // large chunks matter for large k: the persistent eigensolver grid quantises a chunk into rounds of
  // 148 runs (k = 256: 530 KB per unit, 40 GB = 79k units = 33 rounds instead of 10)
  const size_t budget = std::min<size_t>((size_t)40 << 30, free_b / 3);
  int64_t chunk = (int64_t)(budget / per_pt);
  chunk = std::max<int64_t>(1024, std::min<int64_t>(chunk, 1 << 18));
  return std::min<int64_t>(chunk, std::max<int64_t>(npts, 1));
}

struct ChunkOut {  // optional parity outputs of run_pipeline
  int32_t *p = nullptr;       // device [npts]
  double *wbar = nullptr;     // device [npts][k]
  double *Wa = nullptr;       // device [npts][k][k]
  double *xa_raw = nullptr;   // device [npts][k]
  bool transform = true;
  // Host-pointer call: the ensemble fields stay on the host and are moved chunk by chunk -- chunk i+1..
  // uploads and chunk i-1 downloads run on two copy streams while chunk i is analysed (rows of the
  // member-slowest [nfields*k][npts] matrix: one 2-D copy per chunk and direction).
  const float *h_in = nullptr;  // host var, read
  float *h_out = nullptr;       // host var, written (chunks that were analysed)
};

template <typename T>
static void run_pipeline(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts, const float *d_xyz,
                         int nfields, float *d_var, letkf_b200_stats *st, const ChunkOut &co) {
  const int k = c->k;
  cudaStream_t s = c->stream;
  letkf_b200_stats stats;
  std::memset(&stats, 0, sizeof(stats));
  stats.npts = npts;
  LK_CUDA(cudaEventRecord(c->ev[0], s));
  std::vector<ActiveTree> act = prepare_trees(c, cfg);
  stats.ntrees = (int)act.size();
  LK_CUDA(cudaEventRecord(c->ev[1], s));
  float ms_search = 0, ms_gram = 0, ms_eig = 0, ms_xf = 0;
  if (!act.empty() && npts > 0) {  // core:66: nothing to do without trees
    // When every active type is localised in 2-D only, the local observation lists -- hence Yb, C and
    // the weights -- depend on (x, y) alone: all levels of a column share them.  With the level count
    // declared (letkf_b200_set_levels) the search / Gram / eigen stages run once per column and only the
    // transform runs per point.  Same arithmetic on the same inputs as solving every level: exact.
    bool all2d = true;
    for (auto &a : act) all2d = all2d && a.dim == 2;
    const int nz = (all2d && c->nz_hint > 1 && npts % c->nz_hint == 0 && !co.p && !co.wbar && !co.Wa) ? c->nz_hint : 1;
    const int64_t nsearch = npts / nz;  // columns (or all points)
    const int64_t chunk = pick_chunk(c, act, nsearch, sizeof(T));
    TreeViews tv = make_views(c, cfg, act, chunk);
    c->p.ensure(chunk);
    c->unit_pt.ensure(chunk);
    c->nanflag.ensure(chunk);
    c->counters.ensure(4);
    c->rows_sum.ensure(1);
    size_t tmp1 = 0, tmp2 = 0;
    cub::CountingInputIterator<int32_t> cit(0);
    LK_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp1, cit, c->p.p, c->unit_pt.p, c->counters.p, (int)chunk, s));
    LK_CUDA(cub::DeviceReduce::Sum(nullptr, tmp2, c->p.p, c->rows_sum.p, (int)chunk, s));
    c->cub_tmp.ensure(std::max(tmp1, tmp2) + 256);
    LK_CUDA(cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(int32_t), s));
    const float inflat = LK_DIV((float)(k - 1), cfg->multi_infl);  // core:68 (real32)
    const T mu = (T)inflat;                                       // core:645
    const bool host_io = co.h_in != nullptr && nz == 1 && nfields > 0;
    const int io_rows = nfields * k;
    if (host_io) {
      const int64_t nchunks = (nsearch + chunk - 1) / chunk;
      while ((int64_t)c->io_ev.size() < 2 * nchunks) {
        cudaEvent_t e;
        LK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->io_ev.push_back(e);
      }
      for (int64_t ci = 0; ci < nchunks; ++ci) {  // all uploads are queued now; they stream in chunk order
        const int64_t c0 = ci * chunk, nq = std::min(chunk, nsearch - c0);
        LK_CUDA(cudaMemcpy2DAsync(d_var + c0, sizeof(float) * npts, co.h_in + c0, sizeof(float) * npts,
                                  sizeof(float) * nq, io_rows, cudaMemcpyHostToDevice, c->cs_in));
        LK_CUDA(cudaEventRecord(c->io_ev[2 * ci], c->cs_in));
      }
    }
    // letkf_tune_q is pointwise, so with chunked host IO it runs per chunk, before the chunk is downloaded
    const bool tq_chunk = host_io && cfg->tune_q && co.transform;
    // hand a finished chunk to the download stream (chunks that were not analysed are unchanged on the
    // host, unless tune_q touches every point)
    auto chunk_out = [&](int64_t ci, int64_t c0, int64_t nq, bool need_upload_wait) {
      if (need_upload_wait) LK_CUDA(cudaStreamWaitEvent(s, c->io_ev[2 * ci], 0));
      if (tq_chunk)
        for (int f = 0; f < nfields; ++f) launch_tune_q(s, k, npts, d_var + (int64_t)f * npts * k, c0, nq);
      LK_CUDA(cudaEventRecord(c->io_ev[2 * ci + 1], s));
      LK_CUDA(cudaStreamWaitEvent(c->cs_out, c->io_ev[2 * ci + 1], 0));
      LK_CUDA(cudaMemcpy2DAsync(co.h_out + c0, sizeof(float) * npts, d_var + c0, sizeof(float) * npts,
                                sizeof(float) * nq, io_rows, cudaMemcpyDeviceToHost, c->cs_out));
    };
    // FP64 default path: the whole variable is queued without a host synchronisation.  The compaction leaves
    // the unit count of a chunk on the device; Gram and solve are launched with the chunk size as the grid and
    // drop the surplus units themselves (TreeViews::nunits_dev).  A stalled host thread (another process polling
    // the driver, the scheduler) then no longer leaves the GPU idle between chunks -- measured as an occasional
    // +9 % step whose stage times summed to the normal total.
    const bool async_chunks = c->use_fcn && sizeof(T) == 8;
    const int64_t nchunks_all = (nsearch + chunk - 1) / chunk;
    if (async_chunks) {
      c->chunk_cnt.ensure(nchunks_all);
      c->chunk_rows.ensure(nchunks_all);
      LK_CUDA(cudaMemsetAsync(c->chunk_cnt.p, 0, sizeof(int32_t) * nchunks_all, s));
      LK_CUDA(cudaMemsetAsync(c->chunk_rows.p, 0, sizeof(int64_t) * nchunks_all, s));
      while ((int64_t)c->st_ev.size() < 4 * nchunks_all) {
        cudaEvent_t e;
        LK_CUDA(cudaEventCreate(&e));
        c->st_ev.push_back(e);
      }
      const bool fast32 = k == 32 && !c->force_generic;
      c->C.ensure((size_t)chunk * k * k * sizeof(T) + 4096);  // the solve may read (and discard) up to 127 rows past a column
      c->b.ensure((size_t)chunk * k * sizeof(T));
      T *C = reinterpret_cast<T *>(c->C.p), *b = reinterpret_cast<T *>(c->b.p);
      for (int64_t c0 = 0; c0 < nsearch; c0 += chunk) {
        const int64_t nq = std::min(chunk, nsearch - c0);
        const int64_t ci = c0 / chunk;
        cudaEvent_t *ev4 = &c->st_ev[4 * ci];
        LK_CUDA(cudaEventRecord(ev4[0], s));
        for (int t = 0; t < tv.ntrees; ++t) launch_search(s, tv.t[t], nq, d_xyz + c0 * 3);
        launch_count_rows(s, tv, nq, c->p.p);
        size_t tb = c->cub_tmp.n;
        LK_CUDA(cub::DeviceSelect::Flagged(c->cub_tmp.p, tb, cit, c->p.p, c->unit_pt.p, c->chunk_cnt.p + ci, (int)nq, s));
        tb = c->cub_tmp.n;
        LK_CUDA(cub::DeviceReduce::Sum(c->cub_tmp.p, tb, c->p.p, c->chunk_rows.p + ci, (int)nq, s));
        launch_counter() += 2;
        LK_CUDA(cudaEventRecord(ev4[1], s));
        if (co.p) LK_CUDA(cudaMemcpyAsync(co.p + c0, c->p.p, sizeof(int32_t) * nq, cudaMemcpyDeviceToDevice, s));
        if (host_io) LK_CUDA(cudaStreamWaitEvent(s, c->io_ev[2 * ci], 0));  // this chunk's fields are in HBM
        tv.nunits_dev = c->chunk_cnt.p + ci;
        if (fast32)
          launch_gram32(s, tv, nq, c->unit_pt.p, (double)mu, reinterpret_cast<double *>(C), reinterpret_cast<double *>(b),
                        c->nanflag.p);
        else if (c->force_generic)
          launch_gram<T>(s, tv, k, nq, c->unit_pt.p, mu, C, b, c->nanflag.p);
        else if (k % 4 == 0)
          launch_gram_tma<T>(s, tv, k, nq, c->unit_pt.p, mu, C, b, c->nanflag.p);
        else
          launch_gram_dmma<T>(s, tv, k, nq, c->unit_pt.p, mu, C, b, c->nanflag.p);
        LK_CUDA(cudaEventRecord(ev4[2], s));
        FcnArgs fa;
        fa.k = k;
        fa.nunits = nq;
        fa.nunits_dev = tv.nunits_dev;
        fa.C = reinterpret_cast<double *>(C);
        fa.bvec = reinterpret_cast<const double *>(b);
        fa.unit_pt = c->unit_pt.p;
        fa.nanflag = c->nanflag.p;
        fa.mu = (double)mu;
        fa.poles = c->poles.p;
        fa.qmax = c->counters.p + 3;
        fa.npts_total = npts;
        fa.pt_base = c0;
        fa.level_stride = nsearch;
        fa.nz = nz;
        fa.nfields = nfields;
        fa.var = (co.transform && nfields > 0) ? d_var : nullptr;
        fa.use_rtpp = cfg->use_rtpp;
        fa.rtpp_alpha = cfg->rtpp_alpha;
        fa.use_rtps = cfg->use_rtps;
        fa.rtps_alpha = cfg->rtps_alpha;
        fa.xa_raw = co.xa_raw;
        fa.wbar_out = co.wbar ? co.wbar + c0 * k : nullptr;
        fa.Wa_out = co.Wa ? co.Wa + c0 * (int64_t)k * k : nullptr;
        if (fast32)
          launch_fcn32_solve(s, fa);
        else
          launch_fcn_solve(s, fa);
        launch_counter()++;
        LK_CUDA(cudaEventRecord(ev4[3], s));
        if (host_io) chunk_out(ci, c0, nq, false);
      }
      tv.nunits_dev = nullptr;
      std::vector<int32_t> h_cnt(nchunks_all);
      std::vector<int64_t> h_rows(nchunks_all);
      LK_CUDA(cudaMemcpyAsync(h_cnt.data(), c->chunk_cnt.p, sizeof(int32_t) * nchunks_all, cudaMemcpyDeviceToHost, s));
      LK_CUDA(cudaMemcpyAsync(h_rows.data(), c->chunk_rows.p, sizeof(int64_t) * nchunks_all, cudaMemcpyDeviceToHost, s));
      LK_CUDA(cudaStreamSynchronize(s));
      for (int64_t ci = 0; ci < nchunks_all; ++ci) {
        stats.npts_analysed += (int64_t)h_cnt[ci] * nz;
        stats.units += h_cnt[ci];
        stats.rows += h_rows[ci] * nz;
        float ms = 0;
        LK_CUDA(cudaEventElapsedTime(&ms, c->st_ev[4 * ci], c->st_ev[4 * ci + 1]));
        ms_search += ms;
        LK_CUDA(cudaEventElapsedTime(&ms, c->st_ev[4 * ci + 1], c->st_ev[4 * ci + 2]));
        ms_gram += ms;
        LK_CUDA(cudaEventElapsedTime(&ms, c->st_ev[4 * ci + 2], c->st_ev[4 * ci + 3]));
        ms_eig += ms;
      }
    }
    for (int64_t c0 = 0; !async_chunks && c0 < nsearch; c0 += chunk) {
      const int64_t nq = std::min(chunk, nsearch - c0);
      const int64_t ci = c0 / chunk;
      LK_CUDA(cudaEventRecord(c->ev[2], s));
      for (int t = 0; t < tv.ntrees; ++t) launch_search(s, tv.t[t], nq, d_xyz + c0 * 3);
      launch_count_rows(s, tv, nq, c->p.p);
      size_t tb = c->cub_tmp.n;
      LK_CUDA(cub::DeviceSelect::Flagged(c->cub_tmp.p, tb, cit, c->p.p, c->unit_pt.p, c->counters.p, (int)nq, s));
      tb = c->cub_tmp.n;
      LK_CUDA(cub::DeviceReduce::Sum(c->cub_tmp.p, tb, c->p.p, c->rows_sum.p, (int)nq, s));
      launch_counter() += 2;
      LK_CUDA(cudaEventRecord(c->ev[3], s));
      int32_t h_cnt[2] = {0, 0};
      int64_t h_rows = 0;
      LK_CUDA(cudaMemcpyAsync(h_cnt, c->counters.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      LK_CUDA(cudaMemcpyAsync(&h_rows, c->rows_sum.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
      LK_CUDA(cudaStreamSynchronize(s));
      const int64_t nunits = h_cnt[0];
      stats.npts_analysed += nunits * nz;
      stats.units += nunits;
      stats.rows += h_rows * nz;
      if (co.p) LK_CUDA(cudaMemcpyAsync(co.p + c0, c->p.p, sizeof(int32_t) * nq, cudaMemcpyDeviceToDevice, s));
      if (nunits > 0) {
        if (host_io) LK_CUDA(cudaStreamWaitEvent(s, c->io_ev[2 * ci], 0));  // this chunk's fields are in HBM
        c->C.ensure((size_t)nunits * k * k * sizeof(T) + 4096);  // the solve may read (and discard) up to 127 rows past a column
        c->b.ensure((size_t)nunits * k * sizeof(T));
        c->lam.ensure((size_t)nunits * k * sizeof(T));
        c->wbar.ensure((size_t)nunits * k * sizeof(T));
        T *C = reinterpret_cast<T *>(c->C.p), *b = reinterpret_cast<T *>(c->b.p);
        T *lam = reinterpret_cast<T *>(c->lam.p), *wbar = reinterpret_cast<T *>(c->wbar.p);
        const bool fast32 = k == 32 && !c->force_generic;  // warp-per-unit register kernels
        // k = 32 FP64 with one transform per unit and no parity dump: the transform runs in the
        // eigensolver's epilogue and U never leaves the registers
        // (the FP64 default path never gets here: async_chunks above)
        const bool fuse = fast32 && sizeof(T) == 8 && nz == 1 && co.transform && nfields > 0 && !co.wbar &&
                          !co.Wa && eig32_can_fuse();
        Xform32Args xargs{c->unit_pt.p, c->nanflag.p, npts, c0, nfields, d_var, cfg->use_rtpp, cfg->rtpp_alpha,
                          cfg->use_rtps, cfg->rtps_alpha, co.xa_raw};
        if (fast32 && sizeof(T) == 8)
          launch_gram32(s, tv, nunits, c->unit_pt.p, (double)mu, reinterpret_cast<double *>(C),
                        reinterpret_cast<double *>(b), c->nanflag.p);
        else if (!c->force_generic && !(fast32 && sizeof(T) == 4)) {
          static const bool use_tma = [] {
            const char *e = getenv("LETKF_B200_GRAM_TMA");
            return !e || atoi(e) != 0;
          }();
          if (use_tma && k % 4 == 0)
            launch_gram_tma<T>(s, tv, k, nunits, c->unit_pt.p, mu, C, b, c->nanflag.p);
          else
            launch_gram_dmma<T>(s, tv, k, nunits, c->unit_pt.p, mu, C, b, c->nanflag.p);
        }
        else
          launch_gram<T>(s, tv, k, nunits, c->unit_pt.p, mu, C, b, c->nanflag.p);
        LK_CUDA(cudaEventRecord(c->ev[4], s));
        {
        if (fast32)
          launch_eig32_solve<T>(s, nunits, C, b, lam, wbar, c->counters.p + 1, fuse ? &xargs : nullptr);
        else
          launch_eig_solve<T>(s, k, nunits, C, b, lam, wbar, c->counters.p + 1);
        LK_CUDA(cudaEventRecord(c->ev[5], s));
        if (co.wbar || co.Wa) {
          double *wo = co.wbar ? co.wbar + c0 * k : nullptr;
          double *Wo = co.Wa ? co.Wa + c0 * (int64_t)k * k : nullptr;
          if (fast32)
            launch_weights_dump32<T>(s, nunits, c->unit_pt.p, C, lam, wbar, wo, Wo);
          else
            launch_weights_dump<T>(s, k, nunits, c->unit_pt.p, C, lam, wbar, wo, Wo);
        }
        if (co.transform && nfields > 0 && !fuse) {
          for (int l = 0; l < nz; ++l) {  // nz > 1: the same weights serve every level of the column
            const int64_t base = c0 + (int64_t)l * nsearch;
            if (fast32)
              launch_transform32<T>(s, nunits, c->unit_pt.p, npts, base, C, lam, wbar, c->nanflag.p, nfields, d_var,
                                    cfg->use_rtpp, cfg->rtpp_alpha, cfg->use_rtps, cfg->rtps_alpha, co.xa_raw);
            else
              launch_transform<T>(s, k, nunits, c->unit_pt.p, npts, base, C, lam, wbar, c->nanflag.p, nfields, d_var,
                                  cfg->use_rtpp, cfg->rtpp_alpha, cfg->use_rtps, cfg->rtps_alpha, co.xa_raw);
          }
        }
        }
        LK_CUDA(cudaEventRecord(c->ev[6], s));
        if (host_io) chunk_out(ci, c0, nq, false);
        LK_CUDA(cudaEventSynchronize(c->ev[6]));
        float ms = 0;
        LK_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]));
        ms_gram += ms;
        LK_CUDA(cudaEventElapsedTime(&ms, c->ev[4], c->ev[5]));
        ms_eig += ms;
        LK_CUDA(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]));
        ms_xf += ms;
      }
      else if (host_io && tq_chunk) {
        chunk_out(ci, c0, nq, true);
      }
      float ms = 0;
      LK_CUDA(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
      ms_search += ms;
    }
    if (host_io) {
      LK_CUDA(cudaStreamSynchronize(c->cs_in));
      LK_CUDA(cudaStreamSynchronize(c->cs_out));
    }
    if (cfg->tune_q && co.transform && !tq_chunk)
      for (int f = 0; f < nfields; ++f) launch_tune_q(s, k, npts, d_var + (int64_t)f * npts * k);  // core:252-278
    int32_t h_sw[3] = {0, 0, 0};
    LK_CUDA(cudaMemcpyAsync(h_sw, c->counters.p + 1, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LK_CUDA(cudaStreamSynchronize(s));
    // the 32-pole expansion of C^(-1/2) is accurate to < 1e-12 for condition numbers up to 2^27; a unit beyond
    // that (observation errors ~4 orders of magnitude below the spread) is refused rather than answered badly
    LK_REQUIRE(h_sw[2] <= 27, "local analysis matrix with condition number above 2^27 (interval index " +
                                  std::to_string(h_sw[2]) + "): outside the range of the FP64 solve; "
                                  "LETKF_B200_SOLVER=jacobi handles it");
    stats.max_sweeps = h_sw[0];
    stats.sweeps_sum = h_sw[1];
    LK_REQUIRE(h_sw[0] <= (k == 32 && !c->force_generic ? LK_JACOBI_CAP32 : LK_JACOBI_CAP),
               "Jacobi eigensolver did not converge within its sweep cap: the analysis of this variable is invalid");
  }
  LK_CUDA(cudaEventRecord(c->ev[7], s));
  LK_CUDA(cudaEventSynchronize(c->ev[7]));
  LK_CUDA(cudaEventElapsedTime(&stats.ms_tree, c->ev[0], c->ev[1]));
  LK_CUDA(cudaEventElapsedTime(&stats.ms_total, c->ev[0], c->ev[7]));
  stats.ms_search = ms_search;
  stats.ms_gram = ms_gram;
  stats.ms_eigen = ms_eig;
  stats.ms_transform = ms_xf;
  if (st) *st = stats;
}

static void analyze_dev_impl(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts,
                             const float *d_xyz, int nfields, float *d_var, letkf_b200_stats *st) {
  LK_REQUIRE(c && cfg, "null context / config");
  LK_REQUIRE(npts >= 0 && nfields >= 0, "negative size");
  LK_REQUIRE(npts == 0 || (d_xyz && (nfields == 0 || d_var)), "null array");
  LK_CUDA(cudaSetDevice(c->device));
  ChunkOut co;
  if (c->real64)
    run_pipeline<double>(c, cfg, npts, d_xyz, nfields, d_var, st, co);
  else
    run_pipeline<float>(c, cfg, npts, d_xyz, nfields, d_var, st, co);
}

extern "C" int letkf_b200_analyze_dev(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts,
                                      const float *xyz, int nfields, float *var, letkf_b200_stats *st) {
  return guarded(c, [&] { analyze_dev_impl(c, cfg, npts, xyz, nfields, var, st); });
}

static void ensure_copy_streams(letkf_b200_ctx *c) {
  if (c->cs_in) return;
  LK_CUDA(cudaStreamCreateWithFlags(&c->cs_in, cudaStreamNonBlocking));
  LK_CUDA(cudaStreamCreateWithFlags(&c->cs_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    LK_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    LK_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    LK_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
  }
}

// Host-pointer entry point.  By default the coordinates are copied in and the fields stream through the
// device chunk by chunk on two copy streams while other chunks are analysed (on a B200 host the two bulk
// transfers of config M cost ~50 ms of a 1.3 s step when they are not overlapped).  With LETKF_B200_SLAB=<points>
// the grid is instead streamed through double-buffered slabs (slab i+1 copied in and slab i-1 copied
// out on two copy streams while slab i is analysed) -- for grids that do not fit in HBM; measured
// slower than the single pass when everything fits (per-slab overheads > hidden copy time).  Points
// are independent, so slabbing does not change results.
extern "C" int letkf_b200_analyze(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts,
                                  const float *xyz, int nfields, float *var, letkf_b200_stats *st) {
  return guarded(c, [&] {
    LK_REQUIRE(c && cfg, "null context / config");
    LK_REQUIRE(npts >= 0 && nfields >= 0, "negative size");
    LK_REQUIRE(npts == 0 || (xyz && (nfields == 0 || var)), "null array");
    LK_CUDA(cudaSetDevice(c->device));
    const int k = c->k;
    int64_t slab = (int64_t)1 << 40;
    if (const char *e = getenv("LETKF_B200_SLAB"))  // test knob; values <= 0 keep the default
      if (atoll(e) > 0) slab = std::max<int64_t>(32, atoll(e));
    if (npts <= slab + slab / 2 || c->nz_hint > 1) {  // small grid, or whole columns needed: one piece
      const size_t nv = (size_t)npts * k * nfields;
      c->d_xyz.ensure((size_t)npts * 3 + 1);
      c->d_var.ensure(nv + 1);
      LK_CUDA(cudaMemcpyAsync(c->d_xyz.p, xyz, sizeof(float) * 3 * npts, cudaMemcpyHostToDevice, c->stream));
      // The fields move chunk by chunk on two copy streams, overlapped with the analysis of the other chunks
      // (run_pipeline).  Not when all levels of a column share weights: then one copy in, one copy out.
      const bool chunk_io = c->nz_hint == 1 && nfields > 0 && npts > 0 && !getenv("LETKF_B200_BULK_IO");
      if (chunk_io) {
        ensure_copy_streams(c);
        ChunkOut co;
        co.h_in = var;
        co.h_out = var;
        if (c->real64)
          run_pipeline<double>(c, cfg, npts, c->d_xyz.p, nfields, c->d_var.p, st, co);
        else
          run_pipeline<float>(c, cfg, npts, c->d_xyz.p, nfields, c->d_var.p, st, co);
        LK_CUDA(cudaStreamSynchronize(c->stream));
        return;
      }
      LK_CUDA(cudaMemcpyAsync(c->d_var.p, var, sizeof(float) * nv, cudaMemcpyHostToDevice, c->stream));
      analyze_dev_impl(c, cfg, npts, c->d_xyz.p, nfields, c->d_var.p, st);
      LK_CUDA(cudaMemcpyAsync(var, c->d_var.p, sizeof(float) * nv, cudaMemcpyDeviceToHost, c->stream));
      LK_CUDA(cudaStreamSynchronize(c->stream));
      return;
    }
    const int rows = nfields * k;  // field f, member m is row f*k+m of a [rows][npts] host matrix
    ensure_copy_streams(c);
    for (int i = 0; i < 2; ++i) {
      c->slab_xyz[i].ensure((size_t)slab * 3);
      c->slab_var[i].ensure((size_t)slab * rows + 1);
    }
    const int64_t nslab = (npts + slab - 1) / slab;
    auto copy_in = [&](int64_t i) {
      const int b = (int)(i & 1);
      const int64_t s0 = i * slab, ns = std::min(slab, npts - s0);
      if (i >= 2) LK_CUDA(cudaStreamWaitEvent(c->cs_in, c->ev_out[b], 0));  // buffer b drained by slab i-2
      LK_CUDA(cudaMemcpyAsync(c->slab_xyz[b].p, xyz + s0 * 3, sizeof(float) * 3 * ns, cudaMemcpyHostToDevice, c->cs_in));
      if (rows > 0)
        LK_CUDA(cudaMemcpy2DAsync(c->slab_var[b].p, sizeof(float) * ns, var + s0, sizeof(float) * npts,
                                  sizeof(float) * ns, rows, cudaMemcpyHostToDevice, c->cs_in));
      LK_CUDA(cudaEventRecord(c->ev_in[b], c->cs_in));
    };
    letkf_b200_stats tot;
    std::memset(&tot, 0, sizeof(tot));
    copy_in(0);
    for (int64_t i = 0; i < nslab; ++i) {
      const int b = (int)(i & 1);
      const int64_t s0 = i * slab, ns = std::min(slab, npts - s0);
      if (i + 1 < nslab) copy_in(i + 1);
      LK_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[b], 0));
      letkf_b200_stats one;
      analyze_dev_impl(c, cfg, ns, c->slab_xyz[b].p, nfields, c->slab_var[b].p, &one);
      LK_CUDA(cudaEventRecord(c->ev_done[b], c->stream));
      LK_CUDA(cudaStreamWaitEvent(c->cs_out, c->ev_done[b], 0));
      if (rows > 0)
        LK_CUDA(cudaMemcpy2DAsync(var + s0, sizeof(float) * npts, c->slab_var[b].p, sizeof(float) * ns,
                                  sizeof(float) * ns, rows, cudaMemcpyDeviceToHost, c->cs_out));
      LK_CUDA(cudaEventRecord(c->ev_out[b], c->cs_out));
      tot.npts += one.npts;
      tot.npts_analysed += one.npts_analysed;
      tot.rows += one.rows;
      tot.units += one.units;
      tot.ntrees = one.ntrees;
      tot.max_sweeps = std::max(tot.max_sweeps, one.max_sweeps);
      tot.sweeps_sum += one.sweeps_sum;
      tot.ms_tree += one.ms_tree;
      tot.ms_search += one.ms_search;
      tot.ms_gram += one.ms_gram;
      tot.ms_eigen += one.ms_eigen;
      tot.ms_transform += one.ms_transform;
      tot.ms_total += one.ms_total;
    }
    LK_CUDA(cudaStreamSynchronize(c->cs_out));
    LK_CUDA(cudaStreamSynchronize(c->stream));
    if (st) *st = tot;
  });
}

extern "C" int letkf_b200_tune_q_dev(letkf_b200_ctx *c, int64_t npts, float *var) {
  return guarded(c, [&] {
    LK_REQUIRE(c && (npts == 0 || var), "null argument");
    LK_CUDA(cudaSetDevice(c->device));
    launch_tune_q(c->stream, c->k, npts, var);
    LK_CUDA(cudaStreamSynchronize(c->stream));
  });
}
extern "C" int letkf_b200_tune_q(letkf_b200_ctx *c, int64_t npts, float *var) {
  return guarded(c, [&] {
    LK_REQUIRE(c && (npts == 0 || var), "null argument");
    LK_CUDA(cudaSetDevice(c->device));
    const size_t nv = (size_t)npts * c->k;
    c->d_var.ensure(nv + 1);
    LK_CUDA(cudaMemcpyAsync(c->d_var.p, var, sizeof(float) * nv, cudaMemcpyHostToDevice, c->stream));
    launch_tune_q(c->stream, c->k, npts, c->d_var.p);
    LK_CUDA(cudaMemcpyAsync(var, c->d_var.p, sizeof(float) * nv, cudaMemcpyDeviceToHost, c->stream));
    LK_CUDA(cudaStreamSynchronize(c->stream));
  });
}

// ---- stage-level entry points ----------------------------------------------------------------------
extern "C" int letkf_b200_search(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts,
                                 const float *xyz, int32_t *ntrees, int32_t *family, int32_t *type,
                                 int32_t *stride, int32_t *count, int32_t *idx, float *r2) {
  return guarded(c, [&] {
    LK_REQUIRE(c && cfg && ntrees, "null argument");
    LK_CUDA(cudaSetDevice(c->device));
    std::vector<ActiveTree> act = prepare_trees(c, cfg);
    *ntrees = (int32_t)act.size();
    for (size_t t = 0; t < act.size(); ++t) {
      if (family) family[t] = act[t].obs->family;
      if (type) type[t] = act[t].obs->type;
      if (stride) stride[t] = act[t].tc->max_lz_pts;
    }
    if (!idx || act.empty() || npts == 0) {
      LK_CUDA(cudaStreamSynchronize(c->stream));
      return;
    }
    LK_REQUIRE(xyz && count && r2, "null array");
    const int64_t chunk = std::min<int64_t>(npts, 1 << 16);
    TreeViews tv = make_views(c, cfg, act, chunk);
    c->d_xyz.ensure((size_t)npts * 3);
    LK_CUDA(cudaMemcpyAsync(c->d_xyz.p, xyz, sizeof(float) * 3 * npts, cudaMemcpyHostToDevice, c->stream));
    for (int64_t c0 = 0; c0 < npts; c0 += chunk) {
      const int64_t nq = std::min(chunk, npts - c0);
      int64_t off = 0;
      for (int t = 0; t < tv.ntrees; ++t) {
        launch_search(c->stream, tv.t[t], nq, c->d_xyz.p + c0 * 3);
        const int na = tv.t[t].nalloc;
        LK_CUDA(cudaMemcpyAsync(count + (int64_t)t * npts + c0, tv.t[t].cnt, sizeof(int32_t) * nq,
                                cudaMemcpyDeviceToHost, c->stream));
        LK_CUDA(cudaMemcpyAsync(idx + off + c0 * na, tv.t[t].idx, sizeof(int32_t) * nq * na,
                                cudaMemcpyDeviceToHost, c->stream));
        LK_CUDA(cudaMemcpyAsync(r2 + off + c0 * na, tv.t[t].r2, sizeof(float) * nq * na, cudaMemcpyDeviceToHost,
                                c->stream));
        off += npts * na;
      }
      LK_CUDA(cudaStreamSynchronize(c->stream));
    }
  });
}

extern "C" int letkf_b200_yoyb(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts,
                               const float *xyz, int64_t *row_offset, float *yo, float *yb) {
  return guarded(c, [&] {
    LK_REQUIRE(c && cfg && row_offset && (npts == 0 || xyz), "null argument");
    LK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    std::vector<ActiveTree> act = prepare_trees(c, cfg);
    row_offset[0] = 0;
    if (act.empty() || npts == 0) {
      for (int64_t i = 0; i < npts; ++i) row_offset[i + 1] = 0;
      LK_CUDA(cudaStreamSynchronize(s));
      return;
    }
    LK_REQUIRE(npts <= (1 << 20), "letkf_b200_yoyb is a parity entry point: npts <= 2^20");
    TreeViews tv = make_views(c, cfg, act, npts);
    c->d_xyz.ensure((size_t)npts * 3);
    c->p.ensure(npts);
    LK_CUDA(cudaMemcpyAsync(c->d_xyz.p, xyz, sizeof(float) * 3 * npts, cudaMemcpyHostToDevice, s));
    for (int t = 0; t < tv.ntrees; ++t) launch_search(s, tv.t[t], npts, c->d_xyz.p);
    launch_count_rows(s, tv, npts, c->p.p);
    std::vector<int32_t> hp(npts);
    LK_CUDA(cudaMemcpyAsync(hp.data(), c->p.p, sizeof(int32_t) * npts, cudaMemcpyDeviceToHost, s));
    LK_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < npts; ++i) row_offset[i + 1] = row_offset[i] + hp[i];
    if (!yo) return;
    LK_REQUIRE(yb, "yb is null");
    const int64_t rows = row_offset[npts];
    if (rows == 0) return;
    DevBuf<int64_t> d_off;
    DevBuf<float> d_yo, d_yb;
    d_off.ensure(npts + 1);
    d_yo.ensure(rows);
    d_yb.ensure((size_t)rows * c->k);
    LK_CUDA(cudaMemcpyAsync(d_off.p, row_offset, sizeof(int64_t) * (npts + 1), cudaMemcpyHostToDevice, s));
    launch_yoyb_rows(s, tv, c->k, npts, d_off.p, d_yo.p, d_yb.p);
    LK_CUDA(cudaMemcpyAsync(yo, d_yo.p, sizeof(float) * rows, cudaMemcpyDeviceToHost, s));
    LK_CUDA(cudaMemcpyAsync(yb, d_yb.p, sizeof(float) * rows * c->k, cudaMemcpyDeviceToHost, s));
    LK_CUDA(cudaStreamSynchronize(s));
  });
}

extern "C" int letkf_b200_weights(letkf_b200_ctx *c, const letkf_b200_var_config *cfg, int64_t npts,
                                  const float *xyz, const float *xb, int32_t *p, double *wbar, double *Wa,
                                  double *xa_raw) {
  return guarded(c, [&] {
    LK_REQUIRE(c && cfg && (npts == 0 || xyz), "null argument");
    LK_REQUIRE(!xa_raw || xb, "xa_raw needs xb");
    LK_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int k = c->k;
    DevBuf<int32_t> d_p;
    DevBuf<double> d_wbar, d_Wa, d_raw;
    c->d_xyz.ensure((size_t)npts * 3 + 1);
    LK_CUDA(cudaMemcpyAsync(c->d_xyz.p, xyz, sizeof(float) * 3 * npts, cudaMemcpyHostToDevice, s));
    ChunkOut co;
    d_p.ensure(npts + 1);
    LK_CUDA(cudaMemsetAsync(d_p.p, 0, sizeof(int32_t) * npts, s));
    co.p = d_p.p;
    if (wbar) {
      d_wbar.ensure((size_t)npts * k);
      LK_CUDA(cudaMemsetAsync(d_wbar.p, 0, sizeof(double) * npts * k, s));
      co.wbar = d_wbar.p;
    }
    if (Wa) {
      d_Wa.ensure((size_t)npts * k * k);
      LK_CUDA(cudaMemsetAsync(d_Wa.p, 0, sizeof(double) * npts * k * k, s));
      co.Wa = d_Wa.p;
    }
    float *d_var = nullptr;
    int nfields = 0;
    letkf_b200_var_config cfg2 = *cfg;
    cfg2.use_rtpp = cfg2.use_rtps = 0;
    cfg2.tune_q = 0;
    if (xa_raw) {
      d_raw.ensure((size_t)npts * k);
      LK_CUDA(cudaMemsetAsync(d_raw.p, 0, sizeof(double) * npts * k, s));
      co.xa_raw = d_raw.p;
      c->d_var.ensure((size_t)npts * k + 1);
      LK_CUDA(cudaMemcpyAsync(c->d_var.p, xb, sizeof(float) * npts * k, cudaMemcpyHostToDevice, s));
      d_var = c->d_var.p;
      nfields = 1;
    }
    co.transform = xa_raw != nullptr;
    if (c->real64)
      run_pipeline<double>(c, &cfg2, npts, c->d_xyz.p, nfields, d_var, nullptr, co);
    else
      run_pipeline<float>(c, &cfg2, npts, c->d_xyz.p, nfields, d_var, nullptr, co);
    if (p) LK_CUDA(cudaMemcpyAsync(p, d_p.p, sizeof(int32_t) * npts, cudaMemcpyDeviceToHost, s));
    if (wbar) LK_CUDA(cudaMemcpyAsync(wbar, d_wbar.p, sizeof(double) * npts * k, cudaMemcpyDeviceToHost, s));
    if (Wa) LK_CUDA(cudaMemcpyAsync(Wa, d_Wa.p, sizeof(double) * npts * k * k, cudaMemcpyDeviceToHost, s));
    if (xa_raw) LK_CUDA(cudaMemcpyAsync(xa_raw, d_raw.p, sizeof(double) * npts * k, cudaMemcpyDeviceToHost, s));
    LK_CUDA(cudaStreamSynchronize(s));
  });
}

template <typename T>
static void syevd_dev(letkf_b200_ctx *c, int k, int64_t batch, const void *A, void *W, void *V, int32_t *sweeps) {
  c->counters.ensure(4);
  LK_CUDA(cudaMemsetAsync(c->counters.p, 0, 4 * sizeof(int32_t), c->stream));
  if (k == 32 && !c->force_generic)
    launch_syevd32<T>(c->stream, batch, (const T *)A, (T *)W, (T *)V, c->counters.p + 1);
  else
    launch_syevd<T>(c->stream, k, batch, (const T *)A, (T *)W, (T *)V, c->counters.p + 1);
  int32_t h_sw = 0;
  LK_CUDA(cudaMemcpyAsync(&h_sw, c->counters.p + 1, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  LK_CUDA(cudaStreamSynchronize(c->stream));
  if (sweeps) *sweeps = h_sw;
  LK_REQUIRE(h_sw <= (k == 32 && !c->force_generic ? LK_JACOBI_CAP32 : LK_JACOBI_CAP),
             "syevd_batched: the Jacobi iteration did not converge within its sweep cap for at least one matrix");
}

extern "C" int letkf_b200_syevd_batched_dev(letkf_b200_ctx *c, int k, int64_t batch, int real64, const void *A,
                                            void *W, void *V, int32_t *sweeps) {
  return guarded(c, [&] {
    LK_REQUIRE(c && (batch == 0 || (A && W && V)), "null argument");
    LK_REQUIRE(A != V, "A and V must not alias");
    LK_CUDA(cudaSetDevice(c->device));
    if (real64)
      syevd_dev<double>(c, k, batch, A, W, V, sweeps);
    else
      syevd_dev<float>(c, k, batch, A, W, V, sweeps);
  });
}

extern "C" int letkf_b200_syevd_batched(letkf_b200_ctx *c, int k, int64_t batch, int real64, const void *A,
                                        void *W, void *V, int32_t *sweeps) {
  return guarded(c, [&] {
    LK_REQUIRE(c && (batch == 0 || (A && W && V)), "null argument");
    LK_CUDA(cudaSetDevice(c->device));
    const size_t ts = real64 ? 8 : 4;
    const size_t nA = (size_t)batch * k * k * ts, nW = (size_t)batch * k * ts;
    DevBuf<unsigned char> dA, dW, dV;
    dA.ensure(nA + 16);
    dW.ensure(nW + 16);
    dV.ensure(nA + 16);
    LK_CUDA(cudaMemcpyAsync(dA.p, A, nA, cudaMemcpyHostToDevice, c->stream));
    if (real64)
      syevd_dev<double>(c, k, batch, dA.p, dW.p, dV.p, sweeps);
    else
      syevd_dev<float>(c, k, batch, dA.p, dW.p, dV.p, sweeps);
    LK_CUDA(cudaMemcpyAsync(W, dW.p, nW, cudaMemcpyDeviceToHost, c->stream));
    LK_CUDA(cudaMemcpyAsync(V, dV.p, nA, cudaMemcpyDeviceToHost, c->stream));
    LK_CUDA(cudaStreamSynchronize(c->stream));
  });
}

extern "C" int letkf_b200_fma_peak(letkf_b200_ctx *c, int kind, double *tflops) {
  return guarded(c, [&] {
    LK_REQUIRE(c && tflops, "null argument");
    LK_CUDA(cudaSetDevice(c->device));
    *tflops = run_fma_peak(c->stream, kind);
  });
}

// ---- host-only self-test (no GPU): kdtree2_create + the search walk on the CPU -----------------
// Lets the CPU test tier compare the product's tree builder and walk with the oracle before any
// GPU time is spent.  Never called by the pipeline; the pipeline has no CPU path.
extern "C" int letkf_b200_selftest_host_search(int n, const float *obs_xyz, float hclr, float vclr, int64_t nq,
                                               const float *xyz_grid, int max_lz_pts, int32_t *ind_out,
                                               int32_t *nnodes_out, int32_t *count, int32_t *idx, float *r2) {
  return guarded([&] {
    LK_REQUIRE(n > 0 && obs_xyz && hclr > 0 && max_lz_pts >= 1, "selftest: bad arguments");
    const int dim = vclr > 0.0f ? 3 : 2;
    const float hinv = lk_clr_inv(hclr), vinv = dim == 3 ? lk_clr_inv(vclr) : -1.0f;
    std::vector<float> xyz(obs_xyz, obs_xyz + (size_t)3 * n);
    for (int i = 0; i < n; ++i) {
      xyz[(size_t)3 * i + 0] = LK_MUL(xyz[(size_t)3 * i + 0], hinv);
      xyz[(size_t)3 * i + 1] = LK_MUL(xyz[(size_t)3 * i + 1], hinv);
      xyz[(size_t)3 * i + 2] = dim == 3 ? LK_MUL(xyz[(size_t)3 * i + 2], vinv) : -1.0f;
    }
    HostTree ht;
    build_kdtree_host(xyz.data(), n, dim, ht);
    if (ind_out) std::memcpy(ind_out, ht.ind.data(), sizeof(int32_t) * n);
    if (nnodes_out) *nnodes_out = (int32_t)ht.nodes.size();
    if (nq > 0) selftest_host_search(ht, nq, xyz_grid, hinv, vinv, max_lz_pts, count, idx, r2);
  });
}
