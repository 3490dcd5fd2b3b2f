/*
 * letkf_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cwbnwp_letkf_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / reported CPU baseline.
 *
 * PARITY UNPINNED: the reference (lopunch/CWBNWP-LETKF) ships no tests, fixtures or
 * golden vectors, and cannot be compiled in this environment (no Fortran compiler, MPI,
 * NetCDF or Fujitsu SSL2).  This oracle is a line-by-line restatement of the reference
 * algorithm; it is pinned by its own known-answer tests (tests/test_oracle_*.py):
 * brute-force radius search, hand-derived DFS truncation order, closed-form scalar
 * Kalman update, LETKF identities, LAPACK residuals.
 */
#ifndef LETKF_ORACLE_H
#define LETKF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_MAX_SLOTS 5

/* one observation type's namelist slice for ONE variable (module_config.f90:7-34) */
typedef struct {
  int32_t family;     /* 0 = gts (module_gts_omboma.f90:13-22), 1 = radar (module_radar.f90:13-16) */
  int32_t type;       /* reference enum (module_param.f90:28-57 gts, :93-97 radar) */
  int32_t use_it;
  int32_t max_lz_pts;
  float hclr;         /* hclr(ivar) [km]; <=0 : type not used for this variable */
  float vclr;         /* vclr(ivar) [km]; <=0 : 2-D localisation */
  int32_t nvar;       /* slots: 5 synop/ships/metar, 4 sound, 1 gpspw, 1 radar */
  int32_t is_assim[OR_MAX_SLOTS];
  float err_muti[OR_MAX_SLOTS]; /* radar: [0] = namelist error */
  float err_rej[OR_MAX_SLOTS];
} or_type_config;

typedef struct {
  int32_t ntypes;
  int32_t weight_function; /* 0 Gaussian, 1 Gaspari-Cohn (module_config.f90:58) */
  float norain_value;      /* module_config.f90:46 */
  float multi_infl;        /* multi_infl(ivar) */
  int32_t use_rtpp;
  float rtpp_alpha;
  int32_t use_rtps;
  float rtps_alpha;
  or_type_config types[16];
} or_var_config;

typedef struct or_ctx or_ctx;

/* nmember, real64!=0 -> the -DREAL64 build (Makefile:9); blas: path of the OpenBLAS .so */
or_ctx *or_create(int nmember, int real64);
void or_destroy(or_ctx *);
const char *or_last_error(void);

/* obs arrays are Fortran column-major; pointers are borrowed (caller keeps them alive) */
int or_set_obs(or_ctx *, int family, int type, int n, int nvar, const float *xyz /*[3,n]*/,
               const float *obs /*[nvar,n]*/, const float *error /*[nvar,n] gts only*/,
               const float *hdxb /*[nvar,n,k]*/, const int32_t *qc /*[nvar,n,k] gts only*/);

/* build_tree for both families (module_localization.f90:35-167); returns #trees, <0 error */
int or_build_trees(or_ctx *, const or_var_config *);
void or_destroy_trees(or_ctx *);

/* get_lz for one grid point over every tree (module_localization.f90:188-331).
 * Trees are visited gts first then radar, ascending type.  For tree t:
 * out_type[t], out_n[t], and idx/r2 written at out_idx + t*stride (1-based idx). */
int or_get_lz(or_ctx *, const or_var_config *, const float xyz[3], int stride, int32_t *out_family,
              int32_t *out_type, int32_t *out_n, int32_t *out_idx, float *out_r2);

/* letkf_yoyb (module_letkf_core.f90:300-595) from lists produced by or_get_lz.
 * yo[p], yb[k,p] column-major; returns p (<= pmax) */
int or_yoyb(or_ctx *, const or_var_config *, int ntrees, int stride, const int32_t *fam,
            const int32_t *type, const int32_t *n, const int32_t *idx, const float *r2, int pmax,
            float *yo, float *yb);

/* letkf_solve (module_letkf_core.f90:598-700).  Optional debug outputs (may be NULL):
 * wbar[k] = Pa~ Yb yo, Wa[k,k] = sqrt(k-1) Pa~^1/2 (as double even in the real32 build),
 * xa_raw[k] = analysis before the real32 cast / RTPP / RTPS. */
int or_solve(or_ctx *, const float *xb, int p, const float *yo, const float *yb, float inflat,
             int use_rtpp, float rtpp_alpha, int use_rtps, float rtps_alpha, float *xa,
             double *wbar, double *Wa, double *xa_raw);

/* whole hot loop (module_letkf_core.f90:209-240) over npts points.
 * xyz_grid[3,npts]; var[npts, k] member-slowest, updated in place; nfields fields share
 * the configuration (field f at var + f*npts*k).  np_out: #points with p>0.
 * nthreads OpenMP threads (1 = the reference's serial order). */
int or_analyze(or_ctx *, const or_var_config *, int64_t npts, const float *xyz_grid, int nfields,
               float *var, int nthreads, int64_t *np_out, int64_t *rows_out);

/* letkf_tune_q (module_letkf_core.f90:702-733) on var[npts,k] */
void or_tune_q(int nmember, int64_t npts, float *var);

/* ---- low-level kdtree2 access for the pins (module_kdtree2.f90) ---- */
typedef struct or_kdtree or_kdtree;
or_kdtree *or_kd_create(const float *data /*[3,n]*/, int n, int dim);
void or_kd_destroy(or_kdtree *);
int or_kd_r_nearest(const or_kdtree *, const float *qv, float r2, int nalloc, int32_t *idx,
                    float *dis, int *nfound_total);
int or_kd_brute(const or_kdtree *, const float *qv, float r2, int nalloc, int32_t *idx,
                float *dis);
/* flattened tree dump for hand-derived pins: nodes in preorder */
int or_kd_num_nodes(const or_kdtree *);
void or_kd_dump(const or_kdtree *, int32_t *cut_dim, float *cut_val, float *cut_l, float *cut_r,
                int32_t *l, int32_t *u, int32_t *left, int32_t *right, float *box /*[nn,3,2]*/,
                int32_t *ind /*[n]*/);

/* LAPACK ?syevd('V','L') on a batch, for BASELINE config E; A[b] column-major k*k */
int or_syevd_batch(int k, int64_t batch, int real64, const void *A, void *W, void *V, int nthreads);

/* reference constants / scalar functions for pins */
float or_gc1999(void);
float or_search_r2(void);
float or_gaspari_cohn(float x);
float or_expf(float x);
float or_error_inv(float err, float r2, int weight_function);

#ifdef __cplusplus
}
#endif
#endif
