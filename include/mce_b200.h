/* mce_b200.h -- C ABI of libmce_b200.so, the B200-native replacement for the per-step term propagation of the
 * Multivariate Cauchy Estimator in natsnyder1/CauchyFriendly.
 *
 * Every entry point replaces one member of the reference's `struct CauchyEstimator`
 * (/root/reference/include/cauchy_estimator.hpp, cited per function below); include/cauchy_estimator.hpp in
 * this repository wraps them back into that struct so the reference's callers (cauchy_windows.hpp, pycauchy.hpp,
 * src/ *.cpp, the MATLAB mex shims) compile unchanged.  Conventions are the reference's: caller-owned row-major
 * `double*` inputs that are consumed during the call (Gamma is d x pncc row-major; H is one 1 x d row per call;
 * B and u may be NULL), complex outputs as interleaved (re, im) doubles, errors as the reference's
 * numeric_moment_errors bit field (cauchy_constants.hpp:104-116).  No torch / CUDA types cross this boundary.
 *
 * All functions return 0 (or a non-negative value) on success and a negative code on failure;
 * mce_last_error() returns a static message for the calling thread's last failure.  There is no CPU fallback:
 * mce_create() fails with MCE_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef MCE_B200_H_
#define MCE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mce_handle mce_handle;

enum {
  MCE_OK = 0,
  MCE_ERR_NO_DEVICE = -1,
  MCE_ERR_BAD_ARG = -2,
  MCE_ERR_CUDA = -3,
  MCE_ERR_STATE = -4,     /* e.g. stepping past num_estimation_steps (the reference exit(1)s, est:1220-1225) */
  MCE_ERR_CAPACITY = -5,
  MCE_ERR_FTR = -6,       /* the term-reduction rounds did not reach a fixed point (cannot happen for finite inputs) */
  MCE_ERR_SCAN = -7       /* the exact scan of Re fz (early G_SCALE_FACTOR of large steps) disagreed with the serial chain: internal error */
};

typedef struct mce_options {
  int device;                 /* CUDA device ordinal; -1 = current device                                    */
  int tr_search_order[12];    /* TR_SEARCH_IDXS_ORDERING (cauchy_constants.hpp:66); {0,1,2,...} by default   */
  int print_basic_info;       /* reference quirk A.9(iii): when set, moments are re-evaluated after FTR      */
  int fast_moments;           /* 0 (default): every moment sum in the reference's serial order (bit-identical to NUM_CPUS=1: 114
                                 dependent chains at d = 7, 5 ns per slot); 1: two-level tree reductions (deterministic; Im fz, mean
                                 and covariance differ from the serial order by ~1e-9 / ~1e-6 of their largest entry), except Re fz,
                                 which an exact parallel scan keeps bit-identical -- and with it G_SCALE_FACTOR and every count, key
                                 and G value of the following steps                                                  */
  int group_split_threshold;  /* reduction groups with more members are split over several CTAs; 0 = default (192),
                                 -1 = never split.  The results do not depend on it.                          */
  int phase_timing;           /* 1: fill the per-phase ms_* fields of mce_step_stats (adds a stream synchronisation
                                 after every phase of a step); 0 (default): only the CUDA-event totals are measured */
  int lean_group_kernel;      /* 1: every group runs the lean variant of the G-table kernel (no second value table in shared
                                 memory; the engine otherwise uses it only for tables the standard variant cannot hold).
                                 The results do not depend on it.                                             */
  int early_scale_min_slots;  /* steps with at least this many (parent + child) slots take G_SCALE_FACTOR from an exact parallel scan of
                                 Re fz, so that the G-table kernels start while the serial moment chains still run; 0 = default
                                 (400000), -1 = never.  The results do not depend on it.                      */
  int fast_moments_min_slots; /* fast_moments applies to steps with at least this many slots (below, the dependent chains are cheap and
                                 stay, bit-exact); 0 = default (300000)                                        */
  int reserved[2];
} mce_options;

/* Fills `o` with the defaults (device -1, identity search order). */
void mce_default_options(mce_options* o);

/* CauchyEstimator::CauchyEstimator(A0,p0,b0,steps,d,cmcc,pncc,p,print) -- est:87-185.
 * root_point[d] and b_pert[max_shape] are the two vectors the reference draws with libc rand() in its constructor
 * (est:125-128, cell_enumeration.hpp:467-470); the host shim keeps drawing them so both implementations see the
 * same values.  max_shape = (steps-1)*pncc + d  (est:97). */
mce_handle* mce_create(int d, int cmcc, int pncc, int p, int steps, const double* A0, const double* p0, const double* b0,
                       const double* root_point, const double* b_pert, const mce_options* opts);
void mce_destroy(mce_handle* h);                                                   /* ~CauchyEstimator, est:1396 */

/* int CauchyEstimator::step(msmt,Phi,Gamma,beta,H,gamma,B,u) -- est:1211-1245. Returns numeric_moment_errors (>= 0). */
int mce_step(mce_handle* h, double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H,
             double gamma, const double* B, const double* u);

/* Public fields read by the callers after step(): fz, conditional_mean, conditional_variance (est:67-69),
 * G_SCALE_FACTOR, Nt, master_step, numeric_moment_errors, terms_per_shape, shape_range. */
typedef struct mce_moments {
  double fz[2];               /* the public `fz` field: 1+0i after a non-final step unless print_basic_info (est:1172-1176) */
  double fz_after_mu[2];      /* normalisation factor right after the measurement update (what est:795 prints)              */
  double mean[2 * 16];
  double cov[2 * 16 * 16];
  double mean_after_mu[2 * 16];       /* moments as checked right after the measurement update; they differ from mean/cov  */
  double cov_after_mu[2 * 16 * 16];   /* only with print_basic_info (quirk A.9 iii: recomputed from the tables after FTR)   */
  double g_scale_factor;
  int numeric_moment_errors;
  int Nt;                     /* terms after the step (after FTR, or after MU on the window's last step)      */
  int Nt_after_muc;           /* "Total Terms after MUC" (est:791)                                            */
  int master_step;
  int skip_post_mu;
} mce_moments;
int mce_get_moments(mce_handle* h, mce_moments* out);
int mce_shape_range(mce_handle* h);
int mce_get_terms_per_shape(mce_handle* h, int* counts /*[shape_range]*/, int after_muc);

void mce_set_master_step(mce_handle* h, int master_step);   /* callers write this field: cauchy_windows.hpp:538,659 */
int mce_reset(mce_handle* h);                                                       /* reset(), est:1247-1300 */
int mce_reinitialize_start_statistics(mce_handle* h, const double* A0, const double* p0, const double* b0); /* est:1302 */
/* The initial term as the next first step reads it, WITHOUT changing the statistics reset() re-seeds from.  The reference
 * keeps that term in childterms_workspace (setup_first_term, cauchy_term.hpp:772-785); callers write it directly
 * (pycauchy_single_step_reset, pycauchy.hpp:818).  Only valid while master_step == 0. */
int mce_set_first_term(mce_handle* h, const double* A, const double* p, const double* b);
/* b <- b + sign*delta on every term: finalize_extended_moments (sign=-1, delta=Re mean; est:1358-1394) and
 * shift_cf_by_bias (sign=+1; est:1312-1328). A no-op after the window's last step, like the reference. */
int mce_shift_b(mce_handle* h, const double* delta, double sign);
int mce_deterministic_time_prop(mce_handle* h, const double* Phi, const double* B, const double* u); /* est:1331-1355 */

/* Host mirror of the term list for the reference's side consumers (cpdf_ndim.hpp:680-692 reads A, p, b, m, the
 * parent table and enc_B of every term) and for the parity tests.  Call with NULL arrays to get the sizes. */
int mce_export_shape(mce_handle* h, int m, int* n_terms, long long* n_cells_total, double* A /*[n][m*d]*/, double* p /*[n][m]*/,
                     double* b /*[n][d]*/, int* cells /*[n]*/, uint32_t* keys /*[sum cells]*/, double* G /*[sum cells][2]*/);

/* Point-wise 1-D marginal conditional pdf, evaluated on the device from the resident term list (no term export).
 * Replaces PointWiseNDimCauchyCPDF::evaluate_1D_marginal_cpdf (cpdf_ndim.hpp:1233-1354) as it is driven by
 * CauchyCPDFGridDispatcher1D (cpdf_ndim.hpp:2017-2139; pycauchy_get_marginal_1D_pointwise_cpdf, pycauchy.hpp:890-931):
 * the first point is evaluated term by term, the others from the per-term cache, every sum in the reference's term order.
 * `bar_nu` [d] is the direction the reference draws with rand() for hyperplanes orthogonal to the marginal axis
 * (cpdf_ndim.hpp:385-389).
 *   mce_cpdf_grid_count     number of grid points of reset_grid(low, high, res)            (cpdf_ndim.hpp:2055-2072)
 *   mce_marginal_1d_points  ys[k] = f(xs[k]); returns n, 0 when no tables exist (last step of the window), < 0 on error
 *   mce_marginal_1d_grid    fills xy[n][2] = (x, f(x)) like CauchyPoint2D points[]; returns n (<= n_cap) or as above
 *   mce_cpdf_last_ms        device time (CUDA events) of the last evaluation */
int mce_cpdf_grid_count(double grid_low, double grid_high, double grid_res);
int mce_marginal_1d_points(mce_handle* h, int marg_idx, const double* bar_nu, int n, const double* xs, double* ys);
int mce_marginal_1d_grid(mce_handle* h, int marg_idx, const double* bar_nu, double grid_low, double grid_high, double grid_res,
                         double* xy, int n_cap);
/* Point-wise 2-D marginal cpdf of the state pair (marg_idx1 < marg_idx2): PointWiseNDimCauchyCPDF::evaluate_2D_marginal_cpdf
 * (cpdf_ndim.hpp:1356-1455) as driven by CauchyCPDFGridDispatcher2D (cpdf_ndim.hpp:1774-1919; pycauchy.hpp:822-872).  The
 * grid is y-major like the reference's CauchyPoint3D points[]: xyz[(i*nx + j)] = (x_j, y_i, f).  Sums run in the reference's
 * term order; atan2 / sin / cos are the device's (values agree with the reference to rounding noise, not bit for bit). */
int mce_marginal_2d_points(mce_handle* h, int marg_idx1, int marg_idx2, const double* bar_nu, int n, const double* xs, const double* ys, double* zs);
int mce_marginal_2d_grid(mce_handle* h, int marg_idx1, int marg_idx2, const double* bar_nu, double xlo, double xhi, double xres,
                         double ylo, double yhi, double yres, double* xyz /*[n_cap][3]*/, int n_cap, int* nx_out, int* ny_out);
double mce_cpdf_last_ms(mce_handle* h);

/* Statistics of the last step: device milliseconds per phase and algorithmic byte counts (bench.py roofline). */
typedef struct mce_step_stats {
  double ms_total, ms_tp, ms_mu, ms_moments, ms_regroup, ms_ftr, ms_gtable, ms_compact;
  long long parents, slots, terms_after_muc, groups, survivors;
  long long bytes_gtable_algorithmic;   /* SURVEY.md 8(d) formula with the actual per-term cell counts */
  long long bytes_step_algorithmic;
  long long kernel_launches;
  int ftr_rounds_max;
  int diag_unmodelled_alias, diag_hash_overflow;
  double ev_step_ms;                    /* CUDA-event time of the whole step on the engine's stream            */
  double ev_gtable_ms;                  /* CUDA-event time of the group (B-table + G-table) kernel launches     */
  long long gtable_launches;
  long long cells_parents, cells_survivors;   /* actual table cells read / written by the group kernel          */
  long long split_groups;               /* reduction groups large enough to be split over several CTAs          */
  double ev_moments_ms;                 /* CUDA-event time of the moment sums on the side stream                  */
  double ev_ftr_ms;                     /* CUDA-event time of the term reduction (sorts, rounds, group lists)      */
  double ev_mu_ms;                      /* CUDA-event time from the start of the step to the end of the measurement update */
  long long gtable_lean_launches;       /* group-kernel launches that ran the lean variant                          */
} mce_step_stats;
int mce_get_step_stats(mce_handle* h, mce_step_stats* out);

/* ---- ONE estimator partitioned over several ranks (one process per GPU; SURVEY.md 8e, csrc/mce_kern_part.h) ----
 * Replaces the pthread split of the term list (cauchy_estimator.hpp:886-895, 1604).  Every rank creates the estimator with
 * identical arguments and makes identical calls; every term lives on exactly ONE rank.  Per step a rank propagates its own
 * parents, the new terms are routed to the rank that owns their reduction key (all-to-all; range splitters snapped to gaps wider
 * than the reduction window, so term reduction stays global), the owner fetches the tables of the parents its terms descend
 * from, builds the G-tables and keeps the survivors.  Moments: mode 0 (default) adds ALL ranks' slots in the reference's
 * order on every rank (bit-identical to one GPU); mode 1 adds per-rank serial sums in rank order (scales; results depend
 * on the world size in the last bits).  mce_get_moments / mce_get_terms_per_shape return GLOBAL values on every rank;
 * mce_export_shape returns the rank's own terms, mce_shard_export_gpos their positions in the canonical (one-GPU) order.
 * Native transport: NCCL (libnccl.so.2 opened at run time: grouped ncclSend/ncclRecv, ncclAllGather, ncclAllReduce on the
 * estimator's stream).  Rank 0 calls mce_shard_unique_id, ships the 128 bytes to the other ranks by any means (e.g.
 * torch.distributed.broadcast), then every rank calls mce_shard_init.
 * Callback transport: the library calls `fn` for every exchange (op 0: in-place all-gather of world chunks of n bytes at
 * base, chunk `rank` valid on entry; op 1: in-place sum of n uint32 over the ranks; op 2: base points to a
 * mce_alltoallv_args, n = world); the pointers are device pointers on the CUDA build.  Requires at most 16 hyperplanes per term. */
typedef struct { const void* send; void* recv; const long long* soff; const long long* scnt; const long long* roff; const long long* rcnt; } mce_alltoallv_args;
typedef struct {
  int rank, world, owned_terms, imported_parents, local_parents, pad_;
  long long bytes_terms, bytes_parents, bytes_moments, bytes_keys;   /* received from other ranks during the last step */
  double ms_stage[8];   /* CUDA-event times of the exchange stages of the last step: keys + splitters, destinations, term pack + exchange,
                           unpack, import list, requests, parent-record pack + exchange, import store */
} mce_shard_stats;
typedef int (*mce_exchange_fn)(void* ctx, int op, void* base, long long n);
int mce_shard_unique_id(int device, void* id128);
int mce_shard_init(mce_handle* h, int rank, int world, const void* id128);
int mce_shard_init_callback(mce_handle* h, int rank, int world, mce_exchange_fn fn, void* ctx);
int mce_shard_set_moments_mode(mce_handle* h, int mode);
int mce_shard_export_gpos(mce_handle* h, int* out /*[cap]*/, int cap);     /* returns the number of local terms */
int mce_shard_get_stats(mce_handle* h, mce_shard_stats* out);

/* Test hook: keep a host copy of the post-MUC term list and FTR flag arrays of the last step. */
/* Device self-test of the branch-free IEEE division used by the cpdf grid kernel (csrc/mce_math.h: div_nobranch) on n
 * pseudo-random operand pairs: out[0] = pairs flagged valid whose value differs from a / b (must be 0), out[1] = valid pairs. */
int mce_debug_div_selftest(mce_handle* h, long long n, unsigned long long seed, unsigned long long* out /*[2]*/);
/* Test hook for the moment sums (csrc/mce_kern_prop.h: KMomentsSerial): fz = sum g, mean_j = sum g y_j, cov_jk = -sum (g y_j) y_k over
 * n caller-supplied slots, every real accumulator in slot order exactly like the dependent chain of cauchy_estimator.hpp:307-338.
 * g: n complex values, y: n x d complex values (d may be 0: only fz), out: 2 * (1 + d + d*d) doubles. */
int mce_debug_moment_sums(mce_handle* h, long long n, int d, const double* g, const double* y, double* out);
/* Test hook for the exact scan of one serial-order sum (csrc/mce_kern_prop.h: KSumScan, used for Re fz of a partitioned estimator):
 * out[0] = g[0].re + g[1].re + ... added in order, bit for bit; out[1] = how often the scan fell back to the literal loop; out[2] = tiles (of 8192
 * addends) applied in O(1) from the summaries computed in parallel (KSumTileSums / KSumTileMaps). */
int mce_debug_sum_scan(mce_handle* h, long long n, const double* g /* n complex */, double* out /*[3]*/);
/* Test hook: the per-slot moment inputs of the last step (g: n complex, y: n x d complex; either may be NULL), at most `cap` slots are copied;
 * returns the slot count of that step.  The moment kernels add exactly these values (cauchy_estimator.hpp:307-338). */
long long mce_debug_export_slots(mce_handle* h, long long cap, double* g, double* y);
int mce_debug_capture(mce_handle* h, int enable);
int mce_debug_muc_shape(mce_handle* h, int m, int* n_terms, double* A, double* p, double* q, double* b, double* cd /*[n][2]*/,
                        int* meta /*[n][8]*/, uint8_t* cmap /*[n][32]*/, int8_t* csmap /*[n][32]*/, int* F /*[n]*/);

const char* mce_last_error(void);
const char* mce_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCE_B200_H_ */
