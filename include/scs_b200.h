/*
 * scs_b200.h -- C ABI of libscsb200.so, the B200-native (sm_100a) backend for the SCS
 * ADMM iteration.  Plain pointers and sizes only: no torch / CUDA types cross this line.
 *
 * Three groups of entry points, each citing the reference interface it replaces
 * ("S/" = /root/reference/scs_source/):
 *
 *  (1) the public SCS C API   (S/include/scs.h:271-338)  -- device-resident ADMM loop;
 *      this is what scs/scspy.c + scs/scsobject.h bind (scsobject.h:520,903,986,1217,1240).
 *  (2) the ScsLinSysWork plugin ABI (S/include/linsys.h:25-71) with the reference's
 *      host-pointer semantics, so the reference core (scs.c compiled with -DINDIRECT=1)
 *      can link this library as its linear-system backend unchanged.
 *  (3) scs_b200_* : device-resident replacements of the reference's internal host helpers
 *      that its C tests link against (SCS(proj_dual_cone) cones.c:1544, SCS(accum_by_*)
 *      scs_matrix.c:135-199, aa_apply/aa_safeguard aa.c:822-901), exposed with host
 *      buffers for parity tests, plus batch / multi-GPU and measurement hooks that have
 *      no reference counterpart.
 *
 * Struct layouts below are bit-identical to the reference's non-DLONG, non-SFLOAT,
 * non-USE_SPECTRAL_CONES build (scs_types.h:14-33; GPU builds force 32-bit ints,
 * meson.build:169-174).
 */
#ifndef SCS_B200_H_GUARD
#define SCS_B200_H_GUARD

#ifdef __cplusplus
extern "C" {
#endif

typedef int scs_int;      /* S/include/scs_types.h:14-25 (DLONG off) */
typedef double scs_float; /* S/include/scs_types.h:27-33 (SFLOAT off) */

/* exit flags, S/include/scs.h:33-42 */
#define SCS_INFEASIBLE_INACCURATE (-7)
#define SCS_UNBOUNDED_INACCURATE (-6)
#define SCS_SIGINT (-5)
#define SCS_FAILED (-4)
#define SCS_INDETERMINATE (-3)
#define SCS_INFEASIBLE (-2)
#define SCS_UNBOUNDED (-1)
#define SCS_UNFINISHED (0)
#define SCS_SOLVED (1)
#define SCS_SOLVED_INACCURATE (2)

/* S/include/aa_stats.h:21-42 */
typedef struct {
  scs_int iter;
  scs_int n_accept;
  scs_int n_reject_lapack;
  scs_int n_reject_rank0;
  scs_int n_reject_nonfinite;
  scs_int n_reject_weight_cap;
  scs_int n_safeguard_reject;
  scs_int last_rank;
  scs_float last_aa_norm;
  scs_float last_regularization;
} AaStats;

/* CSC matrix, S/include/scs.h:47-58 */
typedef struct {
  scs_float *x;
  scs_int *i;
  scs_int *p;
  scs_int m;
  scs_int n;
} ScsMatrix;

/* S/include/scs.h:61-101 */
typedef struct {
  scs_int normalize;
  scs_float scale;
  scs_int adaptive_scale;
  scs_float rho_x;
  scs_int max_iters;
  scs_float eps_abs;
  scs_float eps_rel;
  scs_float eps_infeas;
  scs_float alpha;
  scs_float time_limit_secs;
  scs_int verbose;
  scs_int warm_start;
  scs_int acceleration_lookback;
  scs_int acceleration_interval;
  scs_int acceleration_type_1;
  scs_float acceleration_regularization;
  scs_float acceleration_relaxation;
  const char *write_data_filename; /* scs_init dumps the problem there (S/src/rw.c:240-260 format) */
  const char *log_csv_filename;    /* scs_solve appends one row per iteration (S/src/rw.c:317-476) */
} ScsSettings;

/* S/include/scs.h:104-119 */
typedef struct {
  scs_int m;
  scs_int n;
  ScsMatrix *A; /* m x n CSC, row indices sorted within a column */
  ScsMatrix *P; /* n x n upper-triangular CSC, or NULL */
  scs_float *b;
  scs_float *c;
} ScsData;

/* S/include/scs.h:122-173 (spectral cones compiled out) */
typedef struct {
  scs_int z;
  scs_int l;
  scs_float *bu;
  scs_float *bl;
  scs_int bsize;
  scs_int *q;
  scs_int qsize;
  scs_int *s;
  scs_int ssize;
  scs_int *cs;
  scs_int cssize;
  scs_int ep;
  scs_int ed;
  scs_float *p;
  scs_int psize;
} ScsCone;

/* S/include/scs.h:181-188 */
typedef struct {
  scs_float *x;
  scs_float *y;
  scs_float *s;
} ScsSolution;

/* S/include/scs.h:191-244 */
typedef struct {
  scs_int iter;
  char status[128];
  char lin_sys_solver[128];
  scs_int status_val;
  scs_int scale_updates;
  scs_float pobj;
  scs_float dobj;
  scs_float res_pri;
  scs_float res_dual;
  scs_float gap;
  scs_float res_infeas;
  scs_float res_unbdd_a;
  scs_float res_unbdd_p;
  scs_float setup_time; /* milliseconds */
  scs_float solve_time; /* milliseconds */
  scs_float scale;
  scs_float comp_slack;
  scs_int rejected_accel_steps;
  scs_int accepted_accel_steps;
  AaStats aa_stats;
  scs_float lin_sys_time; /* milliseconds, CUDA events */
  scs_float cone_time;    /* milliseconds, CUDA events */
  scs_float accel_time;   /* milliseconds, CUDA events */
} ScsInfo;

typedef struct SCS_WORK ScsWork;                /* opaque, S/include/scs.h:30 */
typedef struct SCS_LIN_SYS_WORK ScsLinSysWork;  /* opaque, S/include/scs.h:28 */

/* ---------------------------------------------------------------------------------- */
/* (1) public SCS API -- S/include/scs.h:271-338.  Same names, arguments, ownership   */
/*     (inputs deep-copied, scs.h:253-255) and error behaviour (NULL / SCS_FAILED with */
/*     NaN-filled outputs, scs.c:316-359) as the reference.                            */
/* ---------------------------------------------------------------------------------- */
ScsWork *scs_init(const ScsData *d, const ScsCone *k, const ScsSettings *stgs); /* scs.h:271 */
scs_int scs_update(ScsWork *w, scs_float *b, scs_float *c);                     /* scs.h:285 */
scs_int scs_solve(ScsWork *w, ScsSolution *sol, ScsInfo *info,
                  scs_int warm_start);                                          /* scs.h:300 */
void scs_finish(ScsWork *w);                                                    /* scs.h:308 */
scs_int scs(const ScsData *d, const ScsCone *k, const ScsSettings *stgs,
            ScsSolution *sol, ScsInfo *info);                                   /* scs.h:323 */
void scs_set_default_settings(ScsSettings *stgs);                               /* scs.h:331 */
const char *scs_version(void);                                                  /* scs.h:338 */

/* Problem data files, S/include/rw.h:15-21 (SCS(write_data) / SCS(read_data), S/src/rw.c:240-315): the
 * reference's native binary layout, readable and writable by either library (integer width of the
 * file is converted on read).  write: target is stgs->write_data_filename; read: allocates *d, *k,
 * *stgs, released with scs_b200_free_data (SCS(free_data), S/src/util.c).  Host only, no device. */
scs_int scs_b200_write_data(const ScsData *d, const ScsCone *k, const ScsSettings *stgs);
scs_int scs_b200_read_data(const char *filename, ScsData **d, ScsCone **k, ScsSettings **stgs);
void scs_b200_free_data(ScsData *d, ScsCone *k, ScsSettings *stgs);
/* header line of the per-iteration CSV trace (S/src/rw.c:333-402, without the spectral-cone columns) */
const char *scs_b200_csv_header(void);

/* ---------------------------------------------------------------------------------- */
/* (2) linear-system plugin ABI -- S/include/linsys.h:25-71.  Host pointers in and     */
/*     out (H2D/D2H inside the call); A, P, diag_r are copied, never retained.         */
/* ---------------------------------------------------------------------------------- */
ScsLinSysWork *scs_init_lin_sys_work(const ScsMatrix *A, const ScsMatrix *P,
                                     const scs_float *diag_r);                  /* linsys.h:25 */
void scs_free_lin_sys_work(ScsLinSysWork *w);                                   /* linsys.h:33 */
scs_int scs_solve_lin_sys(ScsLinSysWork *w, scs_float *b, const scs_float *s,
                          scs_float tol);                                       /* linsys.h:53 */
scs_int scs_update_lin_sys_diag_r(ScsLinSysWork *w,
                                  const scs_float *new_diag_r);                 /* linsys.h:64 */
const char *scs_get_lin_sys_method(void);                                       /* linsys.h:71 */
/* total CG iterations so far (reference: ScsLinSysWork.tot_cg_its, cpu/indirect/private.h:28) */
scs_int scs_b200_lin_sys_cg_its(const ScsLinSysWork *w);

/* ---------------------------------------------------------------------------------- */
/* (3) device-resident helpers with host buffers (parity-test surface)                 */
/* ---------------------------------------------------------------------------------- */
typedef struct SCS_B200_CONE_WORK ScsB200ConeWork;
/* SCS(init_cone), cones.c:1490-1530 */
ScsB200ConeWork *scs_b200_init_cone(const ScsCone *k, scs_int m);
/* SCS(proj_dual_cone), cones.c:1544-1588: x (len m, host) <- Pi_{K*}^{R}(x).  D = scal->D
 * (len m) or NULL, r_y (len m) or NULL.  Returns 0, <0 on failure. */
scs_int scs_b200_proj_dual_cone(scs_float *x, ScsB200ConeWork *c, const scs_float *D,
                                const scs_float *r_y);
/* SCS(finish_cone), cones.c:284-338 */
void scs_b200_finish_cone(ScsB200ConeWork *c);
/* Measurement hook (no reference counterpart): `reps` device-resident projections of the
 * slowly drifting input x_r = x0 + r*step*x1 (r_y = 1); returns mean ms per projection (CUDA
 * events on the workspace stream); *sweeps_out = mean Jacobi sweeps per PSD cone. */
double scs_b200_bench_proj_cone(ScsB200ConeWork *c, const scs_float *x0, const scs_float *x1,
                                scs_float step, scs_int reps, scs_int warmup, double *sweeps_out);

/* root_plus, S/src/scs.c:667-688 (static there; exercised by S/test/problems/test_root_plus.h): the device
 * kernel of the ADMM loop on host buffers.  g, p, mu, r: nm entries; tau_scale = diag_r[n+m]; eta = v[n+m].
 * Returns tau (NaN on failure). */
scs_float scs_b200_root_plus(const scs_float *g, const scs_float *p, const scs_float *mu, const scs_float *r,
                             scs_int nm, scs_float tau_scale, scs_float eta);

/* SCS(accum_by_a / accum_by_atrans / accum_by_p), scs_matrix.c:135-199: y += A x etc.
 * A is CSC; P upper-triangular CSC.  Host buffers. Returns 0 on success. */
scs_int scs_b200_accum_by_a(const ScsMatrix *A, const scs_float *x, scs_float *y);
scs_int scs_b200_accum_by_atrans(const ScsMatrix *A, const scs_float *x, scs_float *y);
scs_int scs_b200_accum_by_p(const ScsMatrix *P, const scs_float *x, scs_float *y);

/* Anderson acceleration, S/include/aa.h: aa_init / aa_apply / aa_safeguard / aa_reset /
 * aa_finish / aa_get_stats with host vectors (state S,Y,D stays in HBM). */
typedef struct SCS_B200_AA_WORK ScsB200AaWork;
ScsB200AaWork *scs_b200_aa_init(scs_int dim, scs_int mem, scs_int min_len, scs_int type1,
                                scs_float regularization, scs_float relaxation,
                                scs_float safeguard_factor, scs_float max_weight_norm,
                                scs_int ir_max_steps);
scs_float scs_b200_aa_apply(scs_float *f, const scs_float *x, ScsB200AaWork *a);
scs_int scs_b200_aa_safeguard(scs_float *f_new, scs_float *x_new, ScsB200AaWork *a);
void scs_b200_aa_reset(ScsB200AaWork *a);
AaStats scs_b200_aa_get_stats(ScsB200AaWork *a);
void scs_b200_aa_finish(ScsB200AaWork *a);

/* ---------------------------------------------------------------------------------- */
/* device selection, batch sharding, measurement (no reference counterpart)            */
/* ---------------------------------------------------------------------------------- */
/* CUDA device used by workspaces created afterwards on the calling thread (default 0). */
scs_int scs_b200_set_device(scs_int device);
scs_int scs_b200_device_count(void);

/* Counters of one workspace since scs_init (all device work of this backend). */
typedef struct {
  long long kernel_launches;  /* kernels of this library launched for this workspace */
  long long cg_iters;         /* total CG iterations */
  long long admm_iters;       /* total ADMM iterations */
  long long spmv_calls;       /* SpMV-type kernel launches */
  double spmv_ms;             /* device time in SpMV kernels when timing is enabled */
  double algorithmic_bytes;   /* SURVEY.md 8(d) byte model, accumulated per iteration */
  long long h2d_bytes;        /* bytes copied host->device by this workspace */
  long long d2h_bytes;        /* bytes copied device->host by this workspace */
  long long collectives;      /* NCCL all-reduces issued (row-partitioned mode) */
  long long collective_bytes; /* bytes all-reduced */
  long long tiled_a;          /* 1 when A x runs on the tiled shared-memory SpMV engine (csrc/tiled.cuh) */
  long long tiled_g;          /* 1 when A'z + P p runs on it */
  long long tiled_slots;      /* stored slots of both tiled operators, padding included */
  long long tiled_nnz;        /* non-zeros they hold */
} ScsB200Stats;
scs_int scs_b200_get_stats(const ScsWork *w, ScsB200Stats *out);

/* SpMV micro-benchmark used by bench.py for the roofline of the dominant kernel:
 * which = 0: z = R_y^{-1} A p (CSR of A), 1: Gp = A' z + P p + R_x p (CSR of A', fused).
 * Runs `reps` launches on the workspace's stream, returns average ms per launch measured
 * with CUDA events on that stream; *alg_bytes gets the algorithmic bytes of one launch. */
double scs_b200_bench_spmv(ScsWork *w, scs_int which, scs_int reps, double *alg_bytes);

/* Per-CTA profile of the last launch of a tiled operator (which = 0: A, 1: [A' | P]): 4 doubles per CTA
 * = duration in us, us spent streaming non-zeros, modelled cost of its item list, items.  Returns the CTA
 * count (0 when the operator runs on the row engine). */
scs_int scs_b200_tiled_profile(ScsWork *w, scs_int which, double *out, scs_int cap);

/* Tile geometry of the tiled SpMV engine in this build: {rows per row bin, columns per column bin, warps per
 * CTA, x-slice stages} (csrc/tiled.cuh). */
void scs_b200_tiled_geometry(scs_int out[4]);

/* Host-only test hook: the work plan of the tiled SpMV engine for a synthetic cell map (no device needed).
 * hg[rb * ncb + cb] = groups of the (row bin, column bin) cell, p1 / p2 = row pointers (p2 may be NULL).
 * items_out: 5 ints per item {cta, row bin, slot, first index into seq_out, one past the last};
 * binfo_out: 2 ints per row bin {first slot, pieces}; cost_out: modelled cost per CTA.  Returns the number
 * of items, -1 when a buffer is too small. */
scs_int scs_b200_tiled_plan(scs_int nrows, scs_int ncb, const scs_int *hg, const scs_int *p1, const scs_int *p2,
                            scs_int sms, scs_int *items_out, scs_int items_cap, scs_int *seq_out, scs_int seq_cap,
                            scs_int *binfo_out, double *cost_out, scs_int *ncta_out, scs_int *nseq_out);

/* Iteration marks: the next scs_solve records a CUDA event on the workspace stream at the
 * top of ADMM iteration `begin_iter` and of `end_iter` (or at loop exit if earlier), and
 * switches per-launch event timing of the SpMV kernels on between them. */
scs_int scs_b200_set_marks(ScsWork *w, scs_int begin_iter, scs_int end_iter);
typedef struct {
  double ms;                 /* device time between the two marks (CUDA events) */
  long long iters;           /* ADMM iterations between the marks */
  long long cg_iters;        /* CG iterations between the marks */
  long long kernel_launches; /* kernels of this library launched between the marks */
  double algorithmic_bytes;  /* SURVEY.md 8(d) byte model between the marks */
  double spmv_a_ms, spmv_g_ms;           /* summed device time of the two SpMV kernels */
  long long spmv_a_launches, spmv_g_launches; /* real (not early-exit) launches timed */
  double bytes_a, bytes_g;   /* algorithmic bytes of ONE launch of each kernel */
} ScsB200Marks;
scs_int scs_b200_get_marks(const ScsWork *w, ScsB200Marks *out);

/* Row-partitioned single problem over the GPUs of one box (no reference counterpart).
 * One process per GPU.  Rank 0 obtains an id with scs_b200_dist_unique_id (128 bytes), every
 * rank calls scs_b200_dist_init(rank, world, id) after scs_b200_set_device and before
 * scs_init.  Workspaces created afterwards take the FULL problem on every rank, keep a
 * cone-aligned block of rows of A and the columns those rows touch (columns touched by one rank only
 * are private to it, the others are replicated), and return the full (x, y, s).  Collectives per CG
 * iteration: one sum all-reduce of the shared block of A_g' z_g with the scalar p'Gp riding along, and
 * one gather of two reduction scalars -- both kernels over CUDA-IPC-mapped peer memory when the ranks can map each
 * other (SCS_B200_DIST_P2P=0: ncclAllReduce); a few scalar gathers per ADMM iteration; Anderson acceleration
 * gathers one (mem x (2 mem + 1)) trapezoid per rank (NCCL).  world == 1 with SCS_B200_DIST_SELFTEST=1 in the
 * environment runs the same code path on one GPU (collectives degenerate to copies). */
scs_int scs_b200_dist_unique_id(void *out128);
scs_int scs_b200_dist_init(scs_int rank, scs_int world, const void *id128);
void scs_b200_dist_finalize(void);
/* host-only: the block rank `rank` of `world` would own: out = {row0, m_local, nnz_local, z, l,
 * bsize, qsize, ssize, cssize, ep, ed, psize} */
scs_int scs_b200_dist_partition(const ScsData *d, const ScsCone *k, scs_int rank, scs_int world, scs_int out[12]);
/* host-only: the local problem rank `rank` of `world` would build.  sizes = {row0, m_loc, n_shared, n_loc,
 * nnz(A_loc), nnz(P_loc)}; the local column order is [shared columns (replicated) | columns private to the
 * rank], loc2glob maps it back.  Array pointers may be NULL (a first call then only returns the sizes). */
scs_int scs_b200_dist_local(const ScsData *d, const ScsCone *k, scs_int rank, scs_int world, scs_int sizes[6],
                            scs_int *loc2glob, scs_int *Ap, scs_int *Ai, scs_float *Ax, scs_int *Pp, scs_int *Pi,
                            scs_float *Px, scs_float *c_loc);
scs_int scs_b200_dist_rank(void);
scs_int scs_b200_dist_world(void);

/* Solve `count` independent problems on the current device (no reference counterpart: the
 * reference API is single-problem; BASELINE.json configs[4]).  Arrays of pointers,
 * one per problem; `sol[i]` vectors are allocated when NULL, exactly as scs_solve does.  Problems
 * that fit (every cone of the default build -- zero, nonneg, box, second-order, PSD up to order 32, complex PSD up
 * to order 16, exponential, power; shared-memory footprint <= 200 KB, cold start,
 * lookback <= 10) are solved by the batch engine: ONE kernel launch, one CTA per problem with all
 * state in shared memory (csrc/batch.cu); the others go through the streaming engine one after
 * another.  Batch sharding across GPUs is done by the caller: one process per GPU, problem
 * i -> rank i % world, no collective.  `streams` is reserved.  Returns 0, or the worst failure code. */
scs_int scs_b200_solve_batch(scs_int count, const ScsData *const *d, const ScsCone *const *k,
                             const ScsSettings *stgs, ScsSolution *const *sol, ScsInfo *info,
                             scs_int streams);
/* Host-only: the plan scs_b200_solve_batch would make (no device is touched).  fused_out[count]: 1 = batch kernel,
 * 0 = streaming engine, -1 = fails validation; dims_out[8] = {shared-memory bytes per CTA, resident dense inverse,
 * kernel instantiation with exp / power / PSD cones, largest PSD order (complex: 2 cs), PSD workspaces, cones with a
 * boundary, power cones, box bounds}.  Returns the number of members the batch kernel takes, or -1. */
scs_int scs_b200_batch_plan(scs_int count, const ScsData *const *d, const ScsCone *const *k, const ScsSettings *stgs,
                            scs_int *fused_out, scs_int *dims_out);
/* how the last scs_b200_solve_batch of this process ran: out = {problems in the batch engine,
 * problems streamed, batch-kernel ms (CUDA events), host packing ms, H2D bytes, D2H bytes,
 * CTAs launched, shared memory per CTA, 1 if the batch kernel ran in direct mode (resident dense
 * inverse) / 0 for PCG, total CG iterations, total ADMM iterations, 0, then SM cycles summed over the
 * problems: equilibration, factorisation, linear solves, Anderson acceleration, residual checks, total} */
scs_int scs_b200_batch_stats(double out[18]);

#ifdef __cplusplus
}
#endif
#endif
