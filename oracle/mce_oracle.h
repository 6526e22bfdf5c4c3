/* mce_oracle.h -- TEST INFRASTRUCTURE (checker). Plain-C restatement of the reference's CPU algorithm
 * for the MCE per-step term propagation, serial order (the reference built with NUM_CPUS = 1).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this;
 * the product (cauchyfriendly_b200/csrc, libmce_b200.so) never does.
 *
 * Parity pin: tests/test_oracle_vs_reference.py compares this restatement with the real reference
 * (oracle/_ref/ref_run_cpu1, compiled from /root/reference by oracle/Makefile) bit for bit on every
 * scenario under tests/golden/, and with the committed golden dumps when the reference binary is absent.
 */
#ifndef MCE_ORACLE_H_
#define MCE_ORACLE_H_

#include <complex.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t key; double complex value; } mceo_kcv; /* gtable.hpp:90-94 KeyCValue */

typedef struct mceo_term {      /* cauchy_term.hpp:46-68 CauchyTerm */
  int m, d;
  double *A, *p, *q, *b;
  double c_val, d_val;
  int cells_gtable_p; mceo_kcv* gtable_p;
  int cells_gtable; int* enc_B; mceo_kcv* gtable;
  int enc_lhp; unsigned Horthog_flag;
  uint8_t* c_map; int8_t* cs_map;
  int phc, pbc, z, is_new_child;
} mceo_term;

typedef struct mceo_arena { char** pages; size_t* used; size_t* cap; int n_pages; } mceo_arena;

typedef struct mceo {
  int d, cmcc, pncc, p, Nt, num_estimation_steps, master_step, shape_range;
  mceo_term** terms_dp; int* terms_per_shape;
  double *A0, *p0, *b0, *root_point, *b_pert;
  int tr_order[12];
  double G_SCALE_FACTOR;
  double complex fz, last_fz, *mean, *var, *last_mean, *last_var;
  int numeric_moment_errors, skip_post_mu;
  int print_basic_info;          /* selects quirk A.9(iii): recompute moments after FTR */
  /* statistics of the last step, for dumps */
  int* muc_counts; int Nt_muc; double complex fz_mu;
  int Nt_removed_last;
  mceo_arena gen[2]; int cur_gen; mceo_arena step_arena;
  void (*after_muc)(struct mceo*, void*); void* after_muc_arg;
  int* last_F; int last_F_shape;   /* filled per shape when after_ftr_shape is set */
  void (*after_ftr_shape)(struct mceo*, int m, const int* F, int n, void*); void* after_ftr_arg;
} mceo;

mceo* mceo_create(int d, int cmcc, int pncc, int p, int steps, const double* A0, const double* p0, const double* b0,
                  const double* root_point, const double* b_pert, const int* tr_order);
int  mceo_step(mceo* e, double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H,
               double gamma, const double* B, const double* u);
void mceo_shift_b(mceo* e, const double* delta);   /* b <- b - delta on every term (est:1365-1383) */
/* point-wise 1-D marginal cpdf on a grid (cpdf_ndim.hpp:1233-1354, 2055-2139); returns the number of grid points */
int  mceo_marginal_1d_grid(const mceo* e, int marg_idx, const double* bar_nu, double grid_low, double grid_high, double grid_res,
                           double* xs, double* ys);
/* point-wise 2-D marginal cpdf on a grid (cpdf_ndim.hpp:1356-1455, 1474-1747, 1816-1919); out[ny*nx][3] = (x, y, z) */
int  mceo_marginal_2d_grid(const mceo* e, int idx1, int idx2, const double* bar_nu, const double* gx, const double* gy, double* out);
void mceo_reset(mceo* e);
void mceo_destroy(mceo* e);

/* small pieces exposed so that tests can pin the device restatements of libgcc / glibc arithmetic */
void mceo_cdiv(double a, double b, double c, double d, double* re, double* im);   /* (a+ib)/(c+id) via __divdc3 */
void mceo_cmul(double a, double b, double c, double d, double* re, double* im);   /* via __muldc3 */
double mceo_cabs(double a, double b);                                             /* glibc cabs */

#ifdef __cplusplus
}
#endif
#endif
