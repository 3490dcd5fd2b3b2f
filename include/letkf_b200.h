/*
 * letkf_b200.h -- C ABI of the B200 (sm_100a) local-analysis library.
 *
 * This is the drop-in boundary for the hot path of lopunch/CWBNWP-LETKF: the body of the
 * grid-point loop of letkf_driver (module_letkf_core.f90:209-240) together with the
 * build_tree / destroy_tree calls that bracket it (module_letkf_core.f90:63-64,295) and
 * everything they call (module_localization.f90, module_kdtree2.f90, module_eigen.f90,
 * letkf_yoyb / letkf_solve in module_letkf_core.f90:300-700).  The reference has no FFI;
 * its seam is Fortran module procedures, so every entry point below names the reference
 * procedure(s) it replaces.  The Fortran side binds these with ISO_C_BINDING
 * (cwbnwp_letkf_b200/fortran/letkf_b200_mod.f90).
 *
 * Conventions
 *   - plain C: pointers + sizes, int status return (0 = ok); on failure
 *     letkf_b200_last_error() describes it (the reference convention is `stop "msg"`,
 *     e.g. module_letkf_core.f90:161 -- the Fortran shim prints the message and stops).
 *   - arrays are Fortran column-major exactly as the reference declares them; indices
 *     returned to the caller are 1-based like kdtree2's.
 *   - functions ending in _dev take DEVICE pointers (data already resident in HBM);
 *     the others take HOST pointers and do their own transfers.  STREAM ORDER: the library
 *     launches on its own non-blocking stream (letkf_b200_stream); it does not wait for the
 *     stream that produced a caller's device buffer.  The caller must synchronise its producing
 *     stream (or make letkf_b200_stream wait on an event of it) before a _dev call.  Every call
 *     returns only after its own device work has completed, so results may be consumed on any
 *     stream afterwards.  Host arrays passed to letkf_b200_analyze are streamed with asynchronous
 *     copies: page-locked (pinned) memory lets those copies overlap the analysis; pageable
 *     arrays work but serialise them.  On failure all copies are drained before the call returns.  The caller owns every
 *     array it passes; the library keeps no host pointer after a call returns
 *     (set_obs copies to the device).
 *   - one call at a time per context (the reference hot path is non-reentrant too:
 *     module-level trees module_localization.f90:30-31, eigen workspace module_eigen.f90:4-12).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef LETKF_B200_H
#define LETKF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LETKF_B200_MAX_SLOTS 5
#define LETKF_B200_MAX_TYPES 16
#define LETKF_B200_MAX_MEMBERS 256

/* observation families and types: the reference enums */
#define LETKF_B200_GTS 0   /* gts_structure   module_gts_omboma.f90:13-22 */
#define LETKF_B200_RADAR 1 /* radar_structure module_radar.f90:13-16      */

/* One observation type's namelist slice for ONE updated variable: gts_config /
 * radar_variable_config / gts_variable_config of module_config.f90:7-34 evaluated at ivar. */
typedef struct {
  int32_t family;     /* LETKF_B200_GTS | LETKF_B200_RADAR */
  int32_t type;       /* module_param.f90:28-57 (gts) / :93-97 (radar) */
  int32_t use_it;     /* %use_it */
  int32_t max_lz_pts; /* %max_lz_pts */
  float hclr;         /* %hclr(ivar) [km]; <= 0: type unused for this variable */
  float vclr;         /* %vclr(ivar) [km]; <= 0: 2-D localisation */
  int32_t nvar;       /* observation slots: 5 synop/ships/metar (u,v,t,p,q), 4 sound (u,v,t,q),
                         1 gpspw, 1 radar (module_letkf_core.f90:339-418,470) */
  int32_t is_assim[LETKF_B200_MAX_SLOTS]; /* %<slot>%is_assim(ivar) */
  float err_muti[LETKF_B200_MAX_SLOTS];   /* %<slot>%err_muti ; radar: [0] = %error */
  float err_rej[LETKF_B200_MAX_SLOTS];    /* %<slot>%err_rej */
} letkf_b200_type_config;

/* Everything letkf_driver reads from the namelists for one var_update entry. */
typedef struct {
  int32_t ntypes;
  int32_t weight_function; /* 0 Gaussian, 1 Gaspari-Cohn (module_config.f90:58) */
  float norain_value;      /* module_config.f90:46 */
  float multi_infl;        /* multi_infl(ivar): inflat = (nmember-1)/multi_infl (core:68) */
  int32_t use_rtpp;
  float rtpp_alpha;
  int32_t use_rtps;
  float rtps_alpha;
  int32_t tune_q; /* run letkf_tune_q on the updated field(s) (core:252-278) */
  letkf_b200_type_config types[LETKF_B200_MAX_TYPES];
} letkf_b200_var_config;

typedef struct {
  int64_t npts;          /* grid points swept (core:209-213) */
  int64_t npts_analysed; /* points that reached letkf_solve, p > 0 (core:226) */
  int64_t rows;          /* sum of p over analysed points */
  int64_t units;         /* distinct weight sets solved (= analysed points, or columns when
                            every active type is 2-D and all levels share one obs list) */
  int32_t ntrees;
  int32_t max_sweeps;    /* largest Jacobi sweep count in the eigen stage */
  float ms_tree, ms_search, ms_gram, ms_eigen, ms_transform, ms_total; /* device time per stage */
  int64_t sweeps_sum;    /* total Jacobi sweeps over all units (k = 32 path; 0 if not counted) */
} letkf_b200_stats;

typedef struct letkf_b200_ctx letkf_b200_ctx;

/* ---- life cycle ----------------------------------------------------------------------
 * init  replaces set_optimal_workspace_for_eigen(nmember) (module_eigen.f90:16, called at
 *       cwb_letkf.f90:35) and set_ensemble_constants (module_param.f90:126).
 *       real64 != 0 selects the -DREAL64 arithmetic (Makefile:9, the shipped default).
 * finalize replaces destroy_eigen_array (module_eigen.f90:110, cwb_letkf.f90:63). */
int letkf_b200_init(letkf_b200_ctx **ctx, int nmember, int real64, int device);
int letkf_b200_finalize(letkf_b200_ctx *ctx);
const char *letkf_b200_last_error(void);
int letkf_b200_version(void);

/* ---- observations --------------------------------------------------------------------
 * Hands one platform / radar type to the library once the reference's distribute step has
 * completed (wait_jobs, module_letkf_core.f90:50).  Mirrors gts_structure
 * (module_gts_omboma.f90:13-22: xyz(3,n) obs(nvar,n) error(nvar,n) hdxb(nvar,n,0:k-1)
 * qc(nvar,n,0:k-1)) and radar_structure (module_radar.f90:13-16: obs(n) hdxb(n,0:k-1);
 * nvar = 1, error = qc = NULL).  Also performs the grid-point independent half of
 * letkf_yoyb once: ensemble mean, perturbations, spread, innovation (core:430-434,497-501). */
int letkf_b200_set_obs(letkf_b200_ctx *ctx, int family, int type, int n, int nvar,
                       const float *xyz, const float *obs, const float *error,
                       const float *hdxb, const int32_t *qc);
/* same, arrays already in device memory (multi-GPU: hdxb after the NCCL all-gather that
 * mirrors mpi_iallgatherv at module_gts_omboma.f90:601-605 / module_radar.f90:179) */
int letkf_b200_set_obs_dev(letkf_b200_ctx *ctx, int family, int type, int n, int nvar,
                           const float *xyz, const float *obs, const float *error,
                           const float *hdxb, const int32_t *qc);
int letkf_b200_clear_obs(letkf_b200_ctx *ctx);

/* ---- the hot path --------------------------------------------------------------------
 * Replaces, for one updated variable: build_tree x2 (core:63-64), the loop core:209-240
 * (get_lz, letkf_yoyb, letkf_solve incl. RTPP/RTPS), optionally letkf_tune_q (core:252-278),
 * and destroy_tree (core:295).
 *   xyz_grid[3,npts] : x,y from proj%lonlat_to_xy and alt(i,j,k), metres (core:211-214)
 *   var[npts,k]      : the memory of var(loc_nx,loc_ny,nz,0:nmember-1) (core:85), member
 *                      slowest, updated in place; nfields fields that share this
 *                      configuration are stacked (field f at var + f*npts*k), which lets
 *                      variables with identical localisation/inflation share one set of
 *                      weights (e.g. the eight hydrometeor variables of input.nml:7,37-38).
 *   stats may be NULL. */
int letkf_b200_analyze(letkf_b200_ctx *ctx, const letkf_b200_var_config *cfg, int64_t npts,
                       const float *xyz_grid, int nfields, float *var, letkf_b200_stats *stats);
int letkf_b200_analyze_dev(letkf_b200_ctx *ctx, const letkf_b200_var_config *cfg, int64_t npts,
                           const float *xyz_grid, int nfields, float *var,
                           letkf_b200_stats *stats);

/* letkf_tune_q alone (module_letkf_core.f90:702-733) on var[npts,k] */
int letkf_b200_tune_q(letkf_b200_ctx *ctx, int64_t npts, float *var);
int letkf_b200_tune_q_dev(letkf_b200_ctx *ctx, int64_t npts, float *var);

/* ---- stage-level entry points (parity tests / benchmarks) ----------------------------
 * search: build_tree + get_lz for npts points (module_localization.f90:35-331).  Trees are
 * reported in the reference's visiting order (gts then radar, ascending type).  For tree t,
 * point i: count[t*npts+i] entries at idx/r2[offset_t + i*stride_t ...], offset_t =
 * sum_{u<t} npts*stride_u, stride_t = max_lz_pts.  idx is 1-based, r2 the normalised
 * squared distance, both in kdtree2's visiting order.  Call with idx == NULL to query
 * ntrees/family/type/stride only. */
int letkf_b200_search(letkf_b200_ctx *ctx, const letkf_b200_var_config *cfg, int64_t npts,
                      const float *xyz_grid, int32_t *ntrees, int32_t *family, int32_t *type,
                      int32_t *stride, int32_t *count, int32_t *idx, float *r2);

/* letkf_yoyb for npts points (module_letkf_core.f90:300-595): rows in the reference order,
 * row_offset[npts+1] (prefix sums of p), yo[rows], yb[k,rows].  With yo == NULL only
 * row_offset is filled (sizing pass). */
int letkf_b200_yoyb(letkf_b200_ctx *ctx, const letkf_b200_var_config *cfg, int64_t npts,
                    const float *xyz_grid, int64_t *row_offset, float *yo, float *yb);

/* letkf_solve internals for npts points (module_letkf_core.f90:649-679): p[npts],
 * wbar[k,npts] = Pa~ Yb yo, Wa[k,k,npts] = sqrt(k-1) Pa~^(1/2) (both as double; zero where
 * p == 0), and, if xb != NULL (xb[npts,k] like var), xa_raw[k,npts] = the analysis in
 * working precision before the real32 cast / RTPP / RTPS (core:675).  Any output may be NULL. */
int letkf_b200_weights(letkf_b200_ctx *ctx, const letkf_b200_var_config *cfg, int64_t npts,
                       const float *xyz_grid, const float *xb, int32_t *p, double *wbar,
                       double *Wa, double *xa_raw);

/* Batched symmetric eigensolver: the ?syevd('V','L') of module_eigen.f90:49/66 for `batch`
 * k x k matrices.  A[b] column-major, lower triangle referenced; W[b] ascending; V[b]
 * column-major orthonormal eigenvectors.  real64: double, else float.  sweeps (may be NULL)
 * receives the largest Jacobi sweep count. */
int letkf_b200_syevd_batched(letkf_b200_ctx *ctx, int k, int64_t batch, int real64, const void *A,
                             void *W, void *V, int32_t *sweeps);
int letkf_b200_syevd_batched_dev(letkf_b200_ctx *ctx, int k, int64_t batch, int real64,
                                 const void *A, void *W, void *V, int32_t *sweeps);

/* device FMA-peak micro-benchmark (roofline denominator of the eigen stage): returns the
 * sustained TFLOP/s of dependent-chain-free FMA loops.  kind: 0 = FP64 FMA, 1 = FP32 FMA,
 * 2 = FP64 tensor pipe (mma.sync.m8n8k4.f64), 3 = 2 and 0 interleaved in every warp (sum of both) */
int letkf_b200_fma_peak(letkf_b200_ctx *ctx, int kind, double *tflops);

/* Host-only self-test, no GPU needed and NOT part of any product path: builds the k-d tree of one
 * observation type exactly as the pipeline does on the host (kdtree2_create,
 * module_kdtree2.f90:598-834) and runs the device search routine compiled for the host on nq
 * queries.  obs_xyz[3,n] metres; ind_out[n] = kdtree2's permutation `ind` (1-based);
 * count[nq], idx/r2[nq*max_lz_pts]. */
int letkf_b200_selftest_host_search(int n, const float *obs_xyz, float hclr, float vclr, int64_t nq,
                                    const float *xyz_grid, int max_lz_pts, int32_t *ind_out,
                                    int32_t *nnodes_out, int32_t *count, int32_t *idx, float *r2);

/* Host-only self-test, no GPU needed and NOT part of any product path: copies the pole table of the
 * FP64 solve (C^(-1/2) ~ sqrt(a) sum_j c_j (C + a beta_j)^-1 on a spectrum inside [a, a 2^q]; layout
 * [q][0][j] = c_j, [q][1][j] = beta_j, q = 0..40, j = 0..31) into out[0..cap) and returns its length in
 * doubles (-1 on failure).  It replaces nothing in the reference: module_eigen.f90:37-108 obtains the same
 * functions of C from ?syevd.  The CPU test tier checks the table against x^(-1/2). */
int letkf_b200_selftest_pole_table(double *out, int cap);

/* number of kernels this library launched since init (bench.py's gpu_launches) */
int64_t letkf_b200_launch_count(letkf_b200_ctx *ctx);
/* stream the library launches on (cudaStream_t), for CUDA-event timing by the caller */
void *letkf_b200_stream(letkf_b200_ctx *ctx);
/* Declares the vertical structure of the grid handed to analyze: npts = ncol*nz with point
 * p + l*ncol directly above point p (the memory order of var(loc_nx,loc_ny,nz,:), core:85).  When
 * every active observation type of a variable is localised in 2-D only (vclr <= 0 -- MU, P, PH in
 * input.nml:45-46,52,...), all levels of a column have the same local observations and weights; the
 * library then searches / solves once per column and only transforms per point.  nz = 1 (default)
 * disables the sharing. */
int letkf_b200_set_levels(letkf_b200_ctx *ctx, int nz);
/* tuning knob: points per pipeline chunk (0 = automatic) */
int letkf_b200_set_chunk(letkf_b200_ctx *ctx, int64_t chunk_points);

#ifdef __cplusplus
}
#endif
#endif /* LETKF_B200_H */
