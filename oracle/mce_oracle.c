/* mce_oracle.c -- TEST INFRASTRUCTURE (checker), see mce_oracle.h.
 *
 * Plain-C restatement of the reference's serial CPU path (NUM_CPUS = 1) for
 * CauchyEstimator::step() (/root/reference/include/cauchy_estimator.hpp:1211).  Every function cites the
 * reference file:line it follows.  Operation order is kept (dot products left to right, row-major loops,
 * no FMA: build with -ffp-contract=off) because the reference's epsilon decisions depend on it
 * (BASELINE.md section 3).  Complex arithmetic uses C99 `double complex`, i.e. the same libgcc
 * __divdc3 / __muldc3 and glibc cabs the reference binary calls.
 *
 * File abbreviations: est = cauchy_estimator.hpp, term = cauchy_term.hpp, ce = cell_enumeration.hpp,
 * flat = flattening.hpp, gs = eval_gs.hpp, gt = gtable.hpp, tr = term_reduction.hpp, util = cauchy_util.hpp,
 * la = cauchy_linalg.hpp, cst = cauchy_constants.hpp.
 */
#include "mce_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants, cst:29-73, 88-116 ---- */
static const double RECIPRICAL_TWO_PI = 1.0 / (2.0 * M_PI);
#define COALIGN_TP_EPS 1e-8
#define MU_EPS 1e-10
#define COALIGN_MU_EPS COALIGN_TP_EPS
#define TERM_APPROXIMATION_EPS 1e-15
#define DCE_STORAGE_MULT 4
#define PLU_EPS 1e-15
#define COND_EPS 1e12
#define REDUCTION_EPS 1e-8
#define THRESHOLD_FZ_IMAG_TO_REAL 1e-3
#define HARD_LIMIT_IMAGINARY_MEAN 0.001
#define THRESHOLD_MEAN_IMAG_TO_REAL 1e-1
#define HARD_LIMIT_IMAGINARY_COVARIANCE 2000
#define THRESHOLD_COVARIANCE_IMAG_TO_REAL 10
#define COV_EIGENVALUE_TOLERANCE -1e-5
enum { ERROR_COVARIANCE_UNSTABLE_ANY_STEP = 0, ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT = 1,
       ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT = 2, ERROR_COVARIANCE_AT_CURRENT_STEP_DNE = 3,
       ERROR_MEAN_UNSTABLE_ANY_STEP = 4, ERROR_MEAN_UNSTABLE_CURRENT_STEP_FINAL_MSMT = 5,
       ERROR_MEAN_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT = 6, ERROR_MEAN_AT_CURRENT_STEP_DNE = 7,
       ERROR_FZ_UNSTABLE = 8, ERROR_FZ_NEGATIVE = 9 };
#define kEmpty 0xffffffffu

/* ---- arenas (stand-in for the paged ChunkedPacked* stores, util:578, 1253, 1377) ---- */
static void* arena_alloc(mceo_arena* a, size_t bytes) {
  bytes = (bytes + 15) & ~(size_t)15;
  if (a->n_pages == 0 || a->used[a->n_pages - 1] + bytes > a->cap[a->n_pages - 1]) {
    size_t cap = bytes > ((size_t)8 << 20) ? bytes : ((size_t)8 << 20);
    a->pages = (char**)realloc(a->pages, sizeof(char*) * (a->n_pages + 1));
    a->used = (size_t*)realloc(a->used, sizeof(size_t) * (a->n_pages + 1));
    a->cap = (size_t*)realloc(a->cap, sizeof(size_t) * (a->n_pages + 1));
    a->pages[a->n_pages] = (char*)malloc(cap);
    a->used[a->n_pages] = 0; a->cap[a->n_pages] = cap; a->n_pages++;
  }
  void* p = a->pages[a->n_pages - 1] + a->used[a->n_pages - 1];
  a->used[a->n_pages - 1] += bytes;
  return p;
}
static void arena_reset(mceo_arena* a) {
  for (int i = 0; i < a->n_pages; i++) free(a->pages[i]);
  free(a->pages); free(a->used); free(a->cap);
  memset(a, 0, sizeof(*a));
}

/* ---- tiny linalg, la:331-366, 483-520, 559-641 ---- */
static double dot_prod(const double* x, const double* y, int n) { double z = 0.0; for (int i = 0; i < n; i++) z += x[i] * y[i]; return z; }
static double sum_vec(const double* x, int n) { double s = 0; for (int i = 0; i < n; i++) s += x[i]; return s; }

/* ---- cell counts, ce:34-56 ---- */
static unsigned long long binomialCoeff(int n, int k) {
  unsigned long long res = 1;
  if (k > n - k) k = n - k;
  for (int i = 0; i < k; ++i) { res *= (n - i); res /= (i + 1); }
  return res;
}
static int cell_count_central(int hyp, int dim) {
  if (hyp < dim) return 1 << hyp;
  unsigned long long fc = 0;
  for (int i = 0; i < dim; i++) fc += binomialCoeff(hyp - 1, i);
  return (int)(2 * fc);
}
static int cell_count_general(int hyp, int dim) {
  if (hyp < dim) return 1 << hyp;
  unsigned long long fc = 0;
  for (int i = 0; i < dim + 1; i++) fc += binomialCoeff(hyp, i);
  return (int)fc;
}

/* ---- u32 -> u32 open-addressing set, gt:97-129, 207-262 (only membership/value semantics matter) ---- */
typedef struct { uint32_t key, value; } kv32;
static uint32_t hash32(uint32_t k, uint32_t cap) { k ^= k >> 16; k *= 0x85ebca6b; k ^= k >> 13; k *= 0xc2b2ae35; k ^= k >> 16; return k % cap; }
static void hs_insert(kv32* t, uint32_t cap, uint32_t key, uint32_t value) {
  uint32_t s = hash32(key, cap);
  for (;;) { if (t[s].key == kEmpty || t[s].key == key) { t[s].key = key; t[s].value = value; return; } s = (s + 1) % cap; }
}
static kv32* hs_find(kv32* t, uint32_t cap, uint32_t key) {
  uint32_t s = hash32(key, cap);
  for (uint32_t it = 0; it < cap; it++) { if (t[s].key == key) return t + s; if (t[s].key == kEmpty) return NULL; s = (s + 1) % cap; }
  return NULL;
}

/* ---- G-table lookups (BINSEARCH_STORAGE + HALF_STORAGE), gt:283-301, gs:94-153 ---- */
static int binsearch(const mceo_kcv* g, uint32_t target, int n) {
  int low = 0, high = n - 1;
  while (low <= high) {
    int mid = (low + high) / 2;
    uint32_t mk = g[mid].key;
    if (mk == target) return mid;
    else if (mk > target) high = mid - 1;
    else low = mid + 1;
  }
  return -1;
}
static double complex g_num_binsearch(int enc_l, int two_to_phc_minus1, int rev_phc_mask, const mceo_kcv* gp, int n) {
  if (enc_l & two_to_phc_minus1) {
    int idx = binsearch(gp, (uint32_t)(rev_phc_mask ^ enc_l), n);
    if (idx == -1) return 0;
    return conj(gp[idx].value);
  }
  int idx = binsearch(gp, (uint32_t)enc_l, n);
  if (idx == -1) return 0;
  return gp[idx].value;
}
static int cmp_kcv(const void* a, const void* b) { return (int)(((const mceo_kcv*)a)->key - ((const mceo_kcv*)b)->key); } /* gt:303 */

void mceo_cdiv(double a, double b, double c, double d, double* re, double* im) {
  double complex r = CMPLX(a, b) / CMPLX(c, d); *re = creal(r); *im = cimag(r);
}
void mceo_cmul(double a, double b, double c, double d, double* re, double* im) {
  double complex r = CMPLX(a, b) * CMPLX(c, d); *re = creal(r); *im = cimag(r);
}
double mceo_cabs(double a, double b) { return cabs(CMPLX(a, b)); }

/* ---- util:128-210 precoalign_Gamma_beta ---- */
static int precoalign_Gamma_beta(const double* Gamma, const double* beta, int cmcc, int d, double* tG, double* tb) {
  for (int i = 0; i < cmcc; i++) { for (int j = 0; j < d; j++) tG[i * d + j] = Gamma[j * cmcc + i]; tb[i] = beta[i]; }
  for (int i = 0; i < cmcc; i++) {
    double nf = 0; double* x = tG + i * d;           /* normalize_l1, util:113-125 */
    for (int j = 0; j < d; j++) nf += fabs(x[j]);
    for (int j = 0; j < d; j++) x[j] /= nf;
    tb[i] *= nf;
  }
  char F[64]; memset(F, 1, sizeof(F));
  for (int i = 0; i < cmcc - 1; i++) if (F[i]) {
    const double* gr = tG + i * d;
    for (int j = i + 1; j < cmcc; j++) if (F[j]) {
      const double* gc = tG + j * d; int pos = 1, neg = 1;
      for (int l = 0; l < d; l++) {
        if (pos) pos &= fabs(gr[l] - gc[l]) < COALIGN_MU_EPS;
        if (neg) neg &= fabs(gr[l] + gc[l]) < COALIGN_MU_EPS;
        if (!(pos || neg)) break;
      }
      if (pos || neg) { F[j] = 0; tb[i] += tb[j]; }
    }
  }
  int tc = 0; for (int i = 0; i < cmcc; i++) tc += F[i];
  if (tc != cmcc) {
    int uc = 1;
    for (int j = 1; j < cmcc; j++) if (F[j]) {
      if (uc < j) { memcpy(tG + uc * d, tG + j * d, d * sizeof(double)); tb[uc] = tb[j]; }
      uc++;
    }
    return tc;
  }
  return cmcc;
}

/* ---- term:435-456 time_prop ---- */
static void time_prop(mceo_term* t, const double* Phi, const double* B, const double* u, int cmcc) {
  const int m = t->m, d = t->d;
  double work[m * d > d ? m * d : d];
  memcpy(work, t->A, m * d * sizeof(double));
  for (int i = 0; i < m; i++) for (int j = 0; j < d; j++) {        /* A @ Phi.T, la:352-366 */
    double sum = 0.0;
    for (int k = 0; k < d; k++) sum += work[i * d + k] * Phi[k + j * d];
    t->A[i * d + j] = sum;
  }
  memcpy(work, t->b, d * sizeof(double));
  for (int i = 0; i < d; i++) { double sum = 0.0; for (int j = 0; j < d; j++) sum += Phi[i * d + j] * work[j]; t->b[i] = sum; }
  if (cmcc > 0) {
    for (int i = 0; i < d; i++) { double sum = 0.0; for (int j = 0; j < cmcc; j++) sum += B[i * cmcc + j] * u[j]; work[i] = sum; }
    for (int i = 0; i < d; i++) t->b[i] += 1.0 * work[i];
  }
}
/* ---- term:458-472 normalize_hps ---- */
static void normalize_hps(mceo_term* t, int set_q) {
  const int m = t->m, d = t->d;
  if (set_q) memcpy(t->q, t->p, m * sizeof(double));
  for (int i = 0; i < m; i++) {
    double norm1 = 0;
    for (int j = 0; j < d; j++) norm1 += fabs(t->A[i * d + j]);
    t->p[i] *= norm1;
    for (int j = 0; j < d; j++) t->A[i * d + j] /= norm1;
  }
}
/* ---- term:475-531 tp_coalign ---- */
static int tp_coalign(mceo_term* t, const double* Gamma_T, const double* beta, int cmcc) {
  normalize_hps(t, 0);
  const int d = t->d; int m = t->m;
  char F[64]; memset(F, 1, sizeof(F));
  for (int j = 0; j < cmcc; j++) {
    const double* gr = Gamma_T + j * d;
    for (int k = 0; k < m; k++) if (F[j]) {
      const double* ac = t->A + k * d; int pos = 1, neg = 1;
      for (int l = 0; l < d; l++) {
        if (pos) pos &= fabs(gr[l] - ac[l]) < COALIGN_MU_EPS;
        if (neg) neg &= fabs(gr[l] + ac[l]) < COALIGN_MU_EPS;
        if (!(pos || neg)) break;
      }
      if (pos || neg) { F[j] = 0; t->p[k] += beta[j]; }
    }
  }
  for (int i = 0; i < cmcc; i++) if (F[i]) { memcpy(t->A + m * d, Gamma_T + i * d, d * sizeof(double)); t->p[m++] = beta[i]; }
  t->m = m;
  return m;
}

/* ---- la:1180-1390 PLU / solve_trf / cond('1') ---- */
static int PLU(double* A, int* P, int n, double tol) {
  double temp_row[n];
  for (int j = 0; j < n; ++j) {
    double pivot = tol; int pivot_ind = -1;
    for (int i = j; i < n; ++i) if (fabs(A[i * n + j]) > fabs(pivot)) { pivot = A[i * n + j]; pivot_ind = i; }
    if (pivot_ind == -1) return 1;
    if (pivot_ind != j) {
      memcpy(temp_row, A + j * n, n * sizeof(double));
      memcpy(A + j * n, A + pivot_ind * n, n * sizeof(double));
      memcpy(A + pivot_ind * n, temp_row, n * sizeof(double));
    }
    P[j] = pivot_ind;
    for (int k = j + 1; k < n; ++k) {
      A[k * n + j] /= A[j * n + j];
      double temp = A[k * n + j];
      for (int q = j + 1; q < n; q++) A[k * n + q] -= temp * A[j * n + q];
    }
  }
  return 0;
}
static void back_solve(const double* LU, double* b, int n) {
  for (int i = n - 1; i >= 0; i--) { double sol = b[i]; for (int j = n - 1; j > i; j--) sol -= LU[i * n + j] * b[j]; b[i] = sol / LU[i * n + i]; }
}
static void forward_solve(const double* LU, double* b, int n) {
  for (int i = 0; i < n; i++) { double sol = b[i]; for (int j = 0; j < i; j++) sol -= LU[i * n + j] * b[j]; b[i] = sol; }
}
static void perm_T(const int* P, int* P_T, int n) {        /* la:1261-1275 */
  int Preg[n];
  for (int i = 0; i < n; i++) Preg[i] = i;
  for (int i = 0; i < n; i++) { int t = Preg[i]; Preg[i] = Preg[P[i]]; Preg[P[i]] = t; }
  for (int i = 0; i < n; i++) P_T[Preg[i]] = i;
}
static void solve_trf(const double* LU, const int* P, const double* b, double* x, int n) {
  int P_T[n]; perm_T(P, P_T, n);
  for (int i = 0; i < n; i++) x[P_T[i]] = b[i];
  forward_solve(LU, x, n); back_solve(LU, x, n);
}
static double matrix_one_norm(const double* A, int m, int n) {
  double mx = -1;
  for (int i = 0; i < n; i++) { double v = 0; for (int j = 0; j < m; j++) v += fabs(A[j * n + i]); if (v > mx) mx = v; }
  return mx;
}
static double cond1(double* A, double* work, int* P, int n, double tol) {
  int P_T[n];
  double norm_val = matrix_one_norm(A, n, n);
  if (PLU(A, P, n, tol) == 1) return DBL_MAX;
  memset(work, 0, n * n * sizeof(double));
  perm_T(P, P_T, n);
  for (int i = 0; i < n; i++) { work[i * n + P_T[i]] = 1; forward_solve(A, work + i * n, n); back_solve(A, work + i * n, n); }
  for (int i = 0; i < n; i++) for (int j = i; j < n; j++) { double t = work[i * n + j]; work[i * n + j] = work[j * n + i]; work[j * n + i] = t; }
  return norm_val * matrix_one_norm(work, n, n);
}

/* ---- ce:676-855 make_time_prop_btable (DCE-TP). Writes term->enc_B / cells_gtable. ---- */
static int next_combo(int* c, int k, int n) {   /* lexicographic successor == std::prev_permutation order of ce:71-92 */
  int i = k - 1;
  while (i >= 0 && c[i] == n - k + i) i--;
  if (i < 0) return 0;
  c[i]++;
  for (int j = i + 1; j < k; j++) c[j] = c[j - 1] + 1;
  return 1;
}
static void make_time_prop_btable(mceo* e, mceo_term* term) {
  const int m = term->m, d = term->d;
  if (m < d) { int n = 1 << (m - 1); for (int i = 0; i < n; i++) term->enc_B[i] = i; return; }   /* cells_gtable keeps its stale value (ce:681-687) */
  if (m == d) { term->cells_gtable = 0; return; }   /* combo_counts[m] = 0 for m <= d (ce:454-459): the serial path yields an EMPTY table */
  const int phc = term->phc;
  double Ac[d * d], work[d * d], bc[d], vertex[d]; int P[d], combo[d];
  const double* b_pert = e->b_pert; const double* A = term->A;
  const int two_to_d = 1 << d, cells_parent = term->cells_gtable_p, cells_gen = cell_count_general(m, d);
  const uint32_t cap = (uint32_t)cells_gen * DCE_STORAGE_MULT;
  kv32* hsh = (kv32*)malloc(sizeof(kv32) * cap); memset(hsh, 0xff, sizeof(kv32) * cap);
  int* inter = (int*)malloc(sizeof(int) * (cells_gen + 1));
  /* visited flags F[2^m] (ce:735): a second hash set keeps this restatement usable for large m */
  const uint32_t vcap = (uint32_t)two_to_d * (uint32_t)binomialCoeff(m, d) * 2u + 16u;
  kv32* vis = (kv32*)malloc(sizeof(kv32) * vcap); memset(vis, 0xff, sizeof(kv32) * vcap);
  const int two_to_phc_minus1 = 1 << (phc - 1), two_to_m_minus1 = 1 << (m - 1);
  const int rev_phc_mask = (1 << phc) - 1, rev_m_mask = (1 << m) - 1;
  int count_set = 0;
  for (int j = 0; j < d; j++) combo[j] = j;
  do {
    for (int j = 0; j < d; j++) { memcpy(Ac + j * d, A + combo[j] * d, d * sizeof(double)); bc[j] = b_pert[combo[j]]; }
    double cn = cond1(Ac, work, P, d, PLU_EPS);
    if (cn > COND_EPS) continue;
    solve_trf(Ac, P, bc, vertex, d);
    int enc_sv_niv = 0;
    for (int ac = 0, ci = 0; ac < m; ac++) {
      if (ci < d && combo[ci] == ac) { ci++; continue; }
      if ((dot_prod(A + ac * d, vertex, d) - b_pert[ac]) < 0) enc_sv_niv |= (1 << ac);
    }
    for (int j = 0; j < two_to_d; j++) {
      int enc_sv = enc_sv_niv;
      for (int k = 0; k < d; k++) if ((j >> k) & 1) enc_sv |= (1 << combo[k]);     /* SSav, ce:60-69 */
      if (hs_find(vis, vcap, (uint32_t)enc_sv) == NULL) {
        hs_insert(vis, vcap, (uint32_t)enc_sv, 0);
        int enc_psv = enc_sv & rev_phc_mask;
        if (enc_psv & two_to_phc_minus1) enc_psv ^= rev_phc_mask;
        if (binsearch(term->gtable_p, (uint32_t)enc_psv, cells_parent) != -1) {
          hs_insert(hsh, cap, (uint32_t)enc_sv, (uint32_t)count_set);
          inter[count_set++] = enc_sv;
        }
      }
    }
    if (count_set == cells_gen) break;
  } while (next_combo(combo, d, m));
  char* F = (char*)malloc(count_set + 1); memset(F, 1, count_set + 1);
  int cnt = 0;
  for (int i = 0; i < count_set; i++) if (F[i]) {
    int b = inter[i], b_rev = b ^ rev_m_mask;
    kv32* q = hs_find(hsh, cap, (uint32_t)b_rev);
    if (q) { F[i] = 0; F[q->value] = 0; term->enc_B[cnt++] = (b & two_to_m_minus1) ? b_rev : b; }
  }
  term->cells_gtable = cnt;
  free(F); free(vis); free(inter); free(hsh);
}

/* ---- term:84-310 msmt_update. Children payloads are carved from the step arena. ---- */
static int msmt_update(mceo* e, mceo_term* par, mceo_term* child_terms, double msmt, const double* H, double gamma,
                       int first_update, int last_update) {
  const int m = par->m, d = par->d;
  double mu[(m + 1) * d], rho[m + 1]; int sign_AH[m + 1], F_int[m + 1];
  memcpy(mu, par->A, m * d * sizeof(double)); memcpy(rho, par->p, m * sizeof(double));
  rho[m] = gamma; memset(mu + m * d, 0, d * sizeof(double)); F_int[m] = 1;
  for (int l = 0; l < m; l++) {
    double* mu_il = mu + l * d;
    double H_mu_il = dot_prod(H, mu_il, d), a = fabs(H_mu_il);
    if (a < MU_EPS) { rho[l] = par->p[l]; sign_AH[l] = 1; F_int[l] = 0; }
    else {
      double s = 1.0 / H_mu_il;
      for (int i = 0; i < d; i++) mu_il[i] *= s;
      rho[l] = par->p[l] * a; sign_AH[l] = (H_mu_il > 0) ? 1 : -1; F_int[l] = 1;
    }
  }
  double zeta = msmt - dot_prod(H, par->b, d);
  int nchild = 0;
  for (int t = 0; t < m + 1; t++) if (F_int[t]) {
    mceo_term* child;
    if (t == m) child = par;
    else {
      child = child_terms + nchild++;
      memset(child, 0, sizeof(*child));
      child->m = m; child->d = d;
      child->A = (double*)arena_alloc(&e->step_arena, sizeof(double) * m * d);
      child->p = (double*)arena_alloc(&e->step_arena, sizeof(double) * m);
      child->q = (double*)arena_alloc(&e->step_arena, sizeof(double) * m);
      child->b = (double*)arena_alloc(&e->step_arena, sizeof(double) * d);
      child->c_map = (uint8_t*)arena_alloc(&e->step_arena, m);
      child->cs_map = (int8_t*)arena_alloc(&e->step_arena, m);
    }
    if (!first_update) { child->gtable_p = par->gtable_p; child->cells_gtable_p = par->cells_gtable_p; child->phc = par->phc; child->pbc = m; child->z = t; }
    child->c_val = zeta; child->d_val = rho[t];
    const double* mu_it = mu + t * d;
    for (int i = 0; i < d; i++) child->b[i] = par->b[i] + zeta * mu_it[i];
    int l = 0; unsigned hofs = 0;
    for (int _l = 0; _l < m + 1; _l++) if (_l != t) {
      double* A_tl = child->A + l * d; const double* mu_il = mu + _l * d;
      child->p[l] = rho[_l];
      if (F_int[_l]) for (int i = 0; i < d; i++) A_tl[i] = mu_il[i] - mu_it[i];
      else { memcpy(A_tl, mu_il, d * sizeof(double)); hofs |= (1u << l); }
      l++;
    }
    child->Horthog_flag = hofs;
  }
  par->is_new_child = 0;
  if (!first_update) {
    int enc_sgn_AH = 0, mask_last_bit = 1 << (m - 1);
    for (int l = 0; l < m; l++) if (sign_AH[l] == -1) enc_sgn_AH |= (1 << l);
    par->enc_lhp = enc_sgn_AH;
    if (par->phc < m) par->enc_lhp &= (1 << par->phc) - 1;
    if (enc_sgn_AH & mask_last_bit) enc_sgn_AH ^= (1 << m) - 1;
    for (int i = 0; i < nchild; i++) {
      child_terms[i].enc_lhp = par->enc_lhp; child_terms[i].is_new_child = 1;
      child_terms[i].enc_B = par->enc_B; child_terms[i].cells_gtable = par->cells_gtable;
    }
    if (!last_update) for (int i = 0; i < par->cells_gtable; i++) par->enc_B[i] ^= enc_sgn_AH;
  }
  return nchild;
}

/* ---- term:312-401 eval_g_yei ---- */
static double complex eval_g_yei(const mceo_term* t, const double* root_point, double complex* yei, int first_update) {
  const int m = t->m, d = t->d;
  double sign_A[m], tmp_yei[d], ygi;
  memset(tmp_yei, 0, d * sizeof(double));
  if (t->Horthog_flag) {
    ygi = 0;
    for (int l = 0; l < m; l++) {
      sign_A[l] = dot_prod(t->A + l * d, root_point, d) > 0 ? 1 : -1;
      double sc = t->p[l] * sign_A[l];
      for (int i = 0; i < d; i++) tmp_yei[i] += sc * t->A[l * d + i];
      if (!(t->Horthog_flag & (1u << l))) ygi += t->p[l] * sign_A[l];
    }
  } else {
    for (int l = 0; l < m; l++) {
      sign_A[l] = dot_prod(t->A + l * d, root_point, d) > 0 ? 1 : -1;
      double sc = t->p[l] * sign_A[l];
      for (int i = 0; i < d; i++) tmp_yei[i] += sc * t->A[l * d + i];
    }
    ygi = dot_prod(t->p, sign_A, m);
  }
  double complex g_num_p, g_num_m;
  if (first_update) { g_num_p = 1; g_num_m = 1; }
  else {
    const int phc = t->phc, two_to_phc_minus1 = 1 << (phc - 1), rev_phc_mask = (1 << phc) - 1;
    int enc_lp = 0, enc_lm = 0, k = 0;
    for (int l = 0; l < m; l++) {
      if (k < phc) {
        if (k == t->z) { enc_lm |= (1 << k); k++; if (k == phc) continue; }
        if (sign_A[l] < 0) { enc_lp |= (1 << k); enc_lm |= (1 << k); }
        k++;
      }
    }
    g_num_p = g_num_binsearch(enc_lp ^ t->enc_lhp, two_to_phc_minus1, rev_phc_mask, t->gtable_p, t->cells_gtable_p);
    g_num_m = g_num_binsearch(enc_lm ^ t->enc_lhp, two_to_phc_minus1, rev_phc_mask, t->gtable_p, t->cells_gtable_p);
  }
  double complex g_val = g_num_p / CMPLX(ygi + t->d_val, t->c_val) - g_num_m / CMPLX(ygi - t->d_val, t->c_val);
  g_val *= RECIPRICAL_TWO_PI;
  for (int j = 0; j < d; j++) yei[j] = CMPLX(-tmp_yei[j], t->b[j]);
  return g_val;
}
/* ---- term:403-433 eval_g_yei_after_ftr ---- */
static double complex eval_g_yei_after_ftr(const mceo_term* t, const double* root_point, double complex* yei) {
  const int m = t->m, d = t->d; int enc_sv = 0; double tmp_yei[d];
  memset(tmp_yei, 0, d * sizeof(double));
  for (int l = 0; l < m; l++) {
    double s = dot_prod(t->A + l * d, root_point, d) > 0 ? 1 : -1;
    double sc = t->p[l] * s;
    for (int i = 0; i < d; i++) tmp_yei[i] += sc * t->A[l * d + i];
    if (s < 0) enc_sv |= 1 << l;
  }
  double complex g = g_num_binsearch(enc_sv, 1 << (m - 1), (1 << m) - 1, t->gtable_p, t->cells_gtable_p);
  for (int j = 0; j < d; j++) yei[j] = CMPLX(-tmp_yei[j], t->b[j]);
  return g;
}
static void accumulate_moment(mceo* e, double complex g_val, const double complex* yei) {   /* est:318-325 */
  const int d = e->d;
  e->fz += g_val;
  for (int j = 0; j < d; j++) {
    double complex y = yei[j];
    e->mean[j] += g_val * y;
    for (int k = 0; k < d; k++) e->var[j * d + k] -= g_val * y * yei[k];
  }
}
static void cache_moments(mceo* e, mceo_term* parent, mceo_term* children, int nchild) {       /* est:307-338 */
  double complex yei[e->d];
  double complex g = eval_g_yei(parent, e->root_point, yei, 0);
  accumulate_moment(e, g, yei);
  for (int i = 0; i < nchild; i++) { g = eval_g_yei(children + i, e->root_point, yei, 0); accumulate_moment(e, g, yei); }
}

/* ---- covariance_checker, util:1882-2004. Eigenvalues by cyclic Jacobi on the lower triangle
 *      (the reference's NR tred2/tqli, eig_solve.hpp:263-404, also reads the lower triangle only). ---- */
static void sym_eigvals(const double* A, double* ev, int n) {
  double a[n * n];
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) a[i * n + j] = a[j * n + i] = A[i * n + j];
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) off += a[i * n + j] * a[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
      if (fabs(a[p * n + q]) < 1e-300) continue;
      double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * a[p * n + q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < n; k++) { double akp = a[k * n + p], akq = a[k * n + q]; a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq; }
      for (int k = 0; k < n; k++) { double apk = a[p * n + k], aqk = a[q * n + k]; a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk; }
    }
  }
  for (int i = 0; i < n; i++) ev[i] = a[i * n + i];
}
static int covariance_checker(const double complex* cov, int d) {
  int flags = 0; double cr[d * d], ci[d * d], eig[d];
  for (int i = 0; i < d * d; i++) { cr[i] = creal(cov[i]); ci[i] = cimag(cov[i]); }
  sym_eigvals(cr, eig, d);
  for (int i = 0; i < d; i++) if (eig[i] < COV_EIGENVALUE_TOLERANCE) flags |= 1 << 0;
  for (int i = 0; i < d - 1; i++) {
    double sig_ii = sqrt(cr[i * d + i]);
    for (int j = i + 1; j < d; j++) { double corr = cr[i * d + j] / (sig_ii * sqrt(cr[j * d + j])); if (fabs(corr) > 1) flags |= 1 << 1; }
  }
  for (int i = 0; i < d; i++) for (int j = i; j < d; j++) {
    double ratio = fabs(ci[i * d + j]) / (fabs(cr[i * d + j]) + 1e-15);
    if (ratio > THRESHOLD_COVARIANCE_IMAG_TO_REAL) flags |= 1 << 2;
  }
  for (int i = 0; i < d * d; i++) if (fabs(ci[i]) > HARD_LIMIT_IMAGINARY_COVARIANCE) flags |= 1 << 3;
  return flags;
}
/* ---- est:359-461 moments_numerical_check ---- */
static void moments_numerical_check(mceo* e) {
  const int d = e->d, p = e->p;
  int first_msmt = (e->master_step % p) == 0, not_last = (e->master_step % p) != (p - 1), last = (e->master_step % p) == (p - 1);
  if (first_msmt) {
    e->numeric_moment_errors |= (1 << ERROR_MEAN_AT_CURRENT_STEP_DNE) | (1 << ERROR_COVARIANCE_AT_CURRENT_STEP_DNE);
    e->numeric_moment_errors &= ~((1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_FINAL_MSMT) | (1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT) |
                                  (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT) | (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT));
  }
  int nw = 0;
  if (creal(e->fz) <= 0) nw |= (1 << ERROR_FZ_NEGATIVE);
  if (fabs(cimag(e->fz) / (1e-15 + creal(e->fz))) > THRESHOLD_FZ_IMAG_TO_REAL) nw |= (1 << ERROR_FZ_UNSTABLE);
  int mean_okay = 1;
  for (int i = 0; i < d; i++) {
    double mr = fabs(creal(e->mean[i])), mi = fabs(cimag(e->mean[i])), ratio = mi / (1e-15 + mr);
    if ((ratio > THRESHOLD_MEAN_IMAG_TO_REAL) || (mi > HARD_LIMIT_IMAGINARY_MEAN)) {
      nw |= (1 << ERROR_MEAN_UNSTABLE_ANY_STEP);
      if (not_last) nw |= (1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT);
      if (last) nw |= (1 << ERROR_MEAN_UNSTABLE_CURRENT_STEP_FINAL_MSMT);
      mean_okay = 0;
    }
  }
  int cov_okay = 1;
  if (covariance_checker(e->var, d)) {
    nw |= (1 << ERROR_COVARIANCE_UNSTABLE_ANY_STEP);
    if (not_last) nw |= (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_NOT_FINAL_MSMT);
    if (last) nw |= (1 << ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT);
    cov_okay = 0;
  }
  e->numeric_moment_errors |= nw;
  if (mean_okay) e->numeric_moment_errors &= ~(1 << ERROR_MEAN_AT_CURRENT_STEP_DNE);
  if (cov_okay) e->numeric_moment_errors &= ~(1 << ERROR_COVARIANCE_AT_CURRENT_STEP_DNE);
  if (mean_okay) memcpy(e->last_mean, e->mean, d * sizeof(double complex)); else memcpy(e->mean, e->last_mean, d * sizeof(double complex));
  if (cov_okay) memcpy(e->last_var, e->var, d * d * sizeof(double complex)); else memcpy(e->var, e->last_var, d * d * sizeof(double complex));
  if (!mean_okay && !cov_okay) e->fz = e->last_fz; else e->last_fz = e->fz;
}
/* ---- est:340-357 finalize_cached_moments ---- */
static void finalize_moments_core(mceo* e) {
  const int d = e->d;
  e->G_SCALE_FACTOR = RECIPRICAL_TWO_PI / creal(e->fz);
  double complex Ifz = CMPLX(0, creal(e->fz));
  for (int i = 0; i < d; i++) e->mean[i] /= Ifz;
  for (int i = 0; i < d; i++) for (int j = 0; j < d; j++)
    e->var[i * d + j] = (e->var[i * d + j] / e->fz) - e->mean[i] * e->mean[j];
}
/* ---- est:524-602 compute_moments ---- */
static void compute_moments(mceo* e, int before_ftr) {
  const int d = e->d; int first_step = (e->master_step == 0); double complex yei[d];
  e->fz = 0; memset(e->mean, 0, d * sizeof(double complex)); memset(e->var, 0, d * d * sizeof(double complex));
  for (int m = 1; m < e->shape_range; m++)
    for (int i = 0; i < e->terms_per_shape[m]; i++) {
      double complex g = before_ftr ? eval_g_yei(e->terms_dp[m] + i, e->root_point, yei, first_step)
                                    : eval_g_yei_after_ftr(e->terms_dp[m] + i, e->root_point, yei);
      accumulate_moment(e, g, yei);
    }
  finalize_moments_core(e);
}

/* ---- term:533-745 mu_coalign ---- */
static int mu_coalign(mceo_term* t) {
  normalize_hps(t, 1);
  const int m = t->m, d = t->d;
  char F[m], Hu[m];
  memset(F, 1, m); memset(t->c_map, 255, m); memset(t->cs_map, 1, m);
  if (t->Horthog_flag) for (int j = 0; j < m; j++) Hu[j] = (t->Horthog_flag & (1u << j)) != 0;
  int unique_count = 0;
  for (int j = 0; j < m - 1; j++) if (F[j]) {
    const double* Ar = t->A + j * d;
    t->c_map[j] = (uint8_t)unique_count;
    for (int k = j + 1; k < m; k++) if (F[k]) {
      const double* Ac = t->A + k * d; int pos = 1, neg = 1;
      for (int l = 0; l < d; l++) {
        if (pos) pos &= fabs(Ar[l] - Ac[l]) < COALIGN_MU_EPS;
        if (neg) neg &= fabs(Ar[l] + Ac[l]) < COALIGN_MU_EPS;
        if (!(pos || neg)) break;
      }
      if (pos) {
        if (t->Horthog_flag) {
          int hj = Hu[j], hk = Hu[k];
          if (!hj && !hk) t->q[j] += t->q[k];
          else if (hj && !hk) { t->q[j] = t->q[k]; Hu[j] = 0; }
          else if (!hj && hk) Hu[k] = 0;
          else { Hu[k] = 0; t->q[j] += t->q[k]; }     /* term:611-628 (warning branch) */
        } else t->q[j] += t->q[k];
        F[k] = 0; t->p[j] += t->p[k]; t->c_map[k] = (uint8_t)unique_count; t->cs_map[k] = 1;
      }
      if (neg) {
        if (t->Horthog_flag) {
          int hj = Hu[j], hk = Hu[k];
          if (!hj && !hk) t->q[j] -= t->q[k];
          else if (hj && !hk) { t->q[j] = -t->q[k]; Hu[j] = 0; }
          else if (!hj && hk) Hu[k] = 0;
          else { fprintf(stderr, "[mce_oracle] MUC #2: two H-orthogonal hyperplanes coalign (term:674-691)\n"); exit(1); }
        } else t->q[j] -= t->q[k];
        F[k] = 0; t->p[j] += t->p[k]; t->c_map[k] = (uint8_t)unique_count; t->cs_map[k] = -1;
      }
    }
    unique_count += 1;
  }
  if (F[m - 1]) t->c_map[m - 1] = (uint8_t)unique_count;
  int new_shape = 0;
  for (int i = 0; i < m; i++) new_shape += F[i];
  if (new_shape != m) {
    unique_count = 1;
    for (int j = 1; j < m; j++) if (F[j]) {
      if (unique_count < j) { memcpy(t->A + unique_count * d, t->A + j * d, d * sizeof(double)); t->p[unique_count] = t->p[j]; t->q[unique_count] = t->q[j]; }
      unique_count++;
    }
    if (t->Horthog_flag) {
      unsigned nf = 0; unique_count = 0;
      for (int j = 0; j < m; j++) if (F[j]) nf |= ((unsigned)Hu[j] << unique_count++);
      t->Horthog_flag = nf;
    }
    t->m = new_shape;
  }
  return t->m;
}

/* ---- tr:30-80 build_ordered_point_maps + tr:89-276 fast_term_reduction ---- */
typedef struct { double p; int pi; } PointMap;
static int compare_pointmap(const void* a, const void* b) {
  double p1 = ((const PointMap*)a)->p, p2 = ((const PointMap*)b)->p;
  return p1 > p2 ? 1 : (p1 < p2 ? -1 : 0);
}
static int ftr_binary_search(double qp, const double* arr, int n, int less_than) {
  int high = n - 1, low = 0;
  while (low <= high) { int mid = (high + low) / 2; if (qp < arr[mid]) high = mid - 1; else low = mid + 1; }
  return less_than ? low - 1 : high + 1;
}
static void fast_term_reduction(mceo* e, mceo_term* terms, int* F, int n, int m, int d) {
  const double ep = REDUCTION_EPS;
  double** op = (double**)malloc(d * sizeof(double*)); int** fm = (int**)malloc(d * sizeof(int*)); int** bm = (int**)malloc(d * sizeof(int*));
  PointMap* pm = (PointMap*)malloc((size_t)n * sizeof(PointMap));
  for (int i = 0; i < d; i++) {
    op[i] = (double*)malloc((size_t)n * sizeof(double)); fm[i] = (int*)malloc((size_t)n * sizeof(int)); bm[i] = (int*)malloc((size_t)n * sizeof(int));
    for (int j = 0; j < n; j++) { pm[j].p = terms[j].b[i]; pm[j].pi = j; }
    qsort(pm, n, sizeof(PointMap), compare_pointmap);
    for (int j = 0; j < n; j++) { op[i][j] = pm[j].p; bm[i][j] = pm[j].pi; fm[i][pm[j].pi] = j; }
  }
  free(pm);
  int* cand = (int*)malloc((size_t)n * sizeof(int));
  const int* sidx = e->tr_order;
  for (int i = 0; i < n; i++) if (F[i] == i) {
    const double* point = terms[i].b;
    int cc = 0;
    {                                                     /* construct_candidate_list, tr:108-137 */
      const double* o = op[sidx[0]]; const int* b = bm[sidx[0]]; double qp = point[sidx[0]];
      int lti = ftr_binary_search(qp - ep, o, n, 1), gti = ftr_binary_search(qp + ep, o, n, 0);
      if ((gti - lti) > 2)
        for (int k = lti + 1; k < gti; k++) { int pi = b[k]; if (pi > i && F[pi] == pi) cand[cc++] = pi; }
    }
    if (!cc) continue;
    for (int j = 1; j < d; j++) {                          /* prune_candidate_list, tr:139-157 */
      const double* o = op[sidx[j]]; const int* f = fm[sidx[j]]; double qp = point[sidx[j]];
      int lti = ftr_binary_search(qp - ep, o, n, 1), gti = ftr_binary_search(qp + ep, o, n, 0);
      ++lti; --gti;
      int ncc = cc;
      for (int k = cc - 1; k > -1; --k) { int opi = f[cand[k]]; if (opi < lti || opi > gti) { --ncc; int t = cand[k]; cand[k] = cand[ncc]; cand[ncc] = t; } }
      cc = ncc;
      if (cc == 0) break;
    }
    if (!cc) continue;
    const double* pi_ = terms[i].p; const double* Ai = terms[i].A;
    for (int j = 0; j < cc; j++) {
      int cl = cand[j], ok = 1; const double* pj = terms[cl].p;
      for (int k = 0; k < m; k++) if (fabs(pj[k] - pi_[k]) > ep) { ok = 0; break; }
      if (ok) {
        const double* Aj = terms[cl].A;
        for (int k = 0; k < m; k++) {                       /* quirk A.9(i): indexes Ai[k], Aj[k] (tr:245-249) */
          int pos = 1, neg = 1;
          for (int l = 0; l < d; l++) {
            double ar = Ai[k], ac = Aj[k];
            if (pos) pos &= fabs(ar - ac) < ep;
            if (neg) neg &= fabs(ar + ac) < ep;
            if (!(pos || neg)) break;
          }
          if (!(pos || neg)) { ok = 0; break; }
        }
      }
      if (ok) F[cl] = i;
    }
  }
  free(cand);
  for (int i = 0; i < d; i++) { free(op[i]); free(fm[i]); free(bm[i]); }
  free(op); free(fm); free(bm);
}

/* ---- ce:206-371 make_new_child_btable. Output into `out` (capacity >= 2*cells_parent); returns cell count. ---- */
static int make_new_child_btable(const mceo_term* term, const int* B_mu, int cells_parent, int* out) {
  const int m = term->m, d = term->d;
  if (m <= d) { int n = 1 << (m - 1); for (int i = 0; i < n; i++) out[i] = i; return n; }
  const int pbc = term->pbc, z = term->z;
  const uint32_t cap = (uint32_t)cells_parent * DCE_STORAGE_MULT;
  kv32* hsh = (kv32*)malloc(sizeof(kv32) * cap); memset(hsh, 0xff, sizeof(kv32) * cap);
  char* F = (char*)malloc(cells_parent > pbc ? cells_parent : pbc); memset(F, 1, cells_parent);
  const int coal = m < pbc;
  int* Buc = coal ? (int*)malloc(sizeof(int) * 2 * (cells_parent + 1)) : out;
  for (int j = 0; j < cells_parent; j++) hs_insert(hsh, cap, (uint32_t)B_mu[j], (uint32_t)j);
  const int shift_high = z + 1, shift_z = pbc - 1, mask_low = (1 << z) - 1, mask_z = 1 << z, mask_hbit = 1 << shift_z, rev = (1 << pbc) - 1;
  int count_B = 0;
  for (int j = 0; j < cells_parent; j++) if (F[j]) {
    int b = B_mu[j], bq = b ^ mask_z;
    if (bq & mask_hbit) bq ^= rev;
    kv32* q = hs_find(hsh, cap, (uint32_t)bq);
    if (q) {
      F[j] = 0; F[q->value] = 0;
      int z_bit = (b & mask_z) >> z;
      int csv1 = ((b >> shift_high) << z) | (b & mask_low) | (z_bit << shift_z), csv2 = csv1 ^ mask_hbit;
      Buc[count_B++] = (csv1 & mask_hbit) ? csv1 ^ rev : csv1;
      Buc[count_B++] = (csv2 & mask_hbit) ? csv2 ^ rev : csv2;
    }
  }
  int cells = count_B;
  if (coal) {
    int bit_mask[32], count_coal = 0;
    memset(F, 1, pbc);
    for (int j = 0; j < pbc; j++) { int c = term->c_map[j]; if (F[c]) { F[c] = 0; bit_mask[count_coal++] = 1 << j; } }
    const uint32_t cap2 = (uint32_t)(count_B + 1) * DCE_STORAGE_MULT;
    kv32* h2 = (kv32*)malloc(sizeof(kv32) * cap2); memset(h2, 0xff, sizeof(kv32) * cap2);
    int count_Bc = 0;
    for (int j = 0; j < count_B; j++) {
      int bc = 0, b = Buc[j];
      for (int l = 0; l < count_coal; l++) if (b & bit_mask[l]) bc |= (1 << l);
      if (!hs_find(h2, cap2, (uint32_t)bc)) { hs_insert(h2, cap2, (uint32_t)bc, 0); out[count_Bc++] = bc; }
    }
    cells = count_Bc;
    free(h2); free(Buc);
  }
  free(F); free(hsh);
  return cells;
}
/* ---- ce:584-625 update_btable (sigma from the first component with |.| >= eps in both rows) ---- */
static int orientation_sigma(const double* A_i, const double* A_j, int m, int d) {
  int sigma = 0;
  for (int k = 0, kd = 0; k < m; k++, kd += d) {
    int l = 0;
    while ((fabs(A_i[kd + l]) < REDUCTION_EPS) || (fabs(A_j[kd + l]) < REDUCTION_EPS)) l++;
    if ((A_i[kd + l] * A_j[kd + l]) < 0) sigma |= (1 << k);
  }
  return sigma;
}
static void update_btable(const double* A_i, int* bt_i, const double* A_j, int* bt_j, int cells, int m, int d) {
  int sigma = orientation_sigma(A_i, A_j, m, d);
  if (sigma & (1 << (m - 1))) sigma ^= (1 << m) - 1;
  if (bt_j == NULL) { if (sigma) for (int k = 0; k < cells; k++) bt_i[k] ^= sigma; }
  else { if (sigma) for (int k = 0; k < cells; k++) bt_j[k] = bt_i[k] ^ sigma; else memcpy(bt_j, bt_i, cells * sizeof(int)); }
}

/* ---- flat:77-255 make_gtable. Fills term->gtable (sorted by key); returns 1 when the term is negligible. ---- */
static int make_gtable(mceo_term* term, double G_SCALE_FACTOR) {
  const int m = term->m, phc = term->phc;
  const int two_to_phc_minus1 = 1 << (phc - 1), rev_phc_mask = (1 << phc) - 1;
  int sign_b[m], negligible = 1;
  const double c_val = term->c_val, d_val = term->d_val;
  double p_sum_squared = sum_vec(term->p, m); p_sum_squared *= p_sum_squared;
  const int num_cells = term->cells_gtable, z_idx = term->z, enc_lhp = term->enc_lhp;
  const int Horthog_flag = (int)term->Horthog_flag; const double* q = term->q;
  for (int j = 0; j < num_cells; j++) {
    int b_enc = term->enc_B[j], enc_lp, enc_lm;
    double ygi = 0;
    if (Horthog_flag) {
      for (int k = 0; k < m; k++) { sign_b[k] = ((b_enc >> k) & 1) ? -1 : 1; if (!(Horthog_flag & (1 << k))) ygi += q[k] * sign_b[k]; }
    } else for (int k = 0; k < m; k++) { sign_b[k] = ((b_enc >> k) & 1) ? -1 : 1; ygi += q[k] * sign_b[k]; }
    if (!term->is_new_child) { enc_lp = b_enc & rev_phc_mask; enc_lm = enc_lp; }
    else {
      int k = 0, l = 0; enc_lp = 0; enc_lm = 0;
      while (k < phc) {
        if (k == z_idx) { enc_lm |= (1 << k); k++; if (k == phc) break; }
        int b_val = term->c_map == NULL ? sign_b[l] : sign_b[term->c_map[l]] * term->cs_map[l];
        if (b_val < 0) { enc_lp |= (1 << k); enc_lm |= (1 << k); }
        k++; l++;
      }
    }
    double complex g_num_p = g_num_binsearch(enc_lp ^ enc_lhp, two_to_phc_minus1, rev_phc_mask, term->gtable_p, term->cells_gtable_p);
    double complex g_num_m = g_num_binsearch(enc_lm ^ enc_lhp, two_to_phc_minus1, rev_phc_mask, term->gtable_p, term->cells_gtable_p);
    mceo_kcv kv; kv.key = (uint32_t)b_enc;
    kv.value = g_num_p / CMPLX(ygi + d_val, c_val) - g_num_m / CMPLX(ygi - d_val, c_val);
    kv.value *= G_SCALE_FACTOR;
    term->gtable[j] = kv;
    if (negligible && (p_sum_squared * cabs(kv.value)) > TERM_APPROXIMATION_EPS) negligible = 0;
  }
  qsort(term->gtable, num_cells, sizeof(mceo_kcv), cmp_kcv);
  return negligible;
}
/* ---- flat:259-315 add_gtables + gs:264-292 gs_add_binsearch ---- */
static void add_gtables(mceo_term* ti, mceo_term* tj) {
  const int m = ti->m, d = ti->d, n = ti->cells_gtable;
  const int two_to_m_minus1 = 1 << (m - 1), rev_b = (1 << m) - 1;
  int sigma = orientation_sigma(ti->A, tj->A, m, d);
  for (int k = 0; k < n; k++) {
    int enc_bi = ti->enc_B[k], enc_bj = enc_bi ^ sigma, use_conj = 0;
    if (enc_bj & two_to_m_minus1) { use_conj = 1; enc_bj ^= rev_b; }
    int ii = binsearch(ti->gtable, (uint32_t)enc_bi, n), jj = binsearch(tj->gtable, (uint32_t)enc_bj, n);
    if (ii > -1 && jj > -1) { if (use_conj) ti->gtable[ii].value += conj(tj->gtable[jj].value); else ti->gtable[ii].value += tj->gtable[jj].value; }
  }
}
/* ---- flat:14-67 make_gtable_first ---- */
static void make_gtable_first(mceo_term* term, double G_SCALE_FACTOR) {
  const int m = term->d;
  for (int j = 0; j < term->cells_gtable; j++) term->enc_B[j] = j;
  for (int j = 0; j < term->cells_gtable; j++) {
    double ygi = 0;
    for (int k = 0; k < m; k++) if (!(term->Horthog_flag & (1u << k))) { double s = ((j >> k) & 1) == 0 ? 1 : -1; ygi += term->p[k] * s; }
    mceo_kcv kv; kv.key = (uint32_t)j;
    kv.value = 1.0 / CMPLX(ygi + term->d_val, term->c_val) - 1.0 / CMPLX(ygi - term->d_val, term->c_val);
    kv.value *= G_SCALE_FACTOR;
    term->gtable[j] = kv;
  }
  qsort(term->gtable, term->cells_gtable, sizeof(mceo_kcv), cmp_kcv);
}

/* Deep copy of an accepted root into the next generation (become_parent term:748-755 + reduce_store util:1399-1406). */
static void keep_term(mceo* e, mceo_arena* ng, mceo_term* dst, const mceo_term* src) {
  const int m = src->m, d = src->d, n = src->cells_gtable;
  *dst = *src;
  dst->A = (double*)arena_alloc(ng, sizeof(double) * m * d); memcpy(dst->A, src->A, sizeof(double) * m * d);
  dst->p = (double*)arena_alloc(ng, sizeof(double) * m); memcpy(dst->p, src->p, sizeof(double) * m);
  dst->b = (double*)arena_alloc(ng, sizeof(double) * d); memcpy(dst->b, src->b, sizeof(double) * d);
  dst->q = NULL; dst->c_map = NULL; dst->cs_map = NULL;
  dst->enc_B = (int*)arena_alloc(ng, sizeof(int) * (n ? n : 1)); memcpy(dst->enc_B, src->enc_B, sizeof(int) * n);
  dst->gtable_p = (mceo_kcv*)arena_alloc(ng, sizeof(mceo_kcv) * (n ? n : 1)); memcpy(dst->gtable_p, src->gtable, sizeof(mceo_kcv) * n);
  dst->phc = m; dst->cells_gtable_p = n; dst->gtable = NULL; dst->is_new_child = 0;
  (void)e;
}

/* ---- flat:318-568 make_gtables for one shape. B-table pointer sharing is kept exactly as in the
 *      reference (children alias their parent's enc_B; in-place re-orientation, flat:433-441). ---- */
static void make_gtables(mceo* e, mceo_arena* ng, int* Nt_reduced, int* Nt_removed, mceo_term* terms, mceo_term* ftr_terms,
                         const int* F, int** fwd, const int* fwd_counts, int Nt_shape, int m, int d) {
  int nred = 0, nrem = 0;
  /* scratch: fresh B/G memory per candidate (the reference reuses chunk memory; contents are what matter).
   * No table of any shape exceeds cell_count_central(max_shape, d) / 2 cells. */
  const int cap_cells = cell_count_central(e->shape_range - 1, d) + 16;
  int* Sroot = (int*)malloc(sizeof(int) * cap_cells); int* Smem = (int*)malloc(sizeof(int) * cap_cells);
  mceo_kcv* Groot = (mceo_kcv*)malloc(sizeof(mceo_kcv) * cap_cells); mceo_kcv* Gmem = (mceo_kcv*)malloc(sizeof(mceo_kcv) * cap_cells);
  for (int j = 0; j < Nt_shape; j++) if (F[j] == j) {
    int rt_idx = j; mceo_term* child_j = terms + rt_idx;
    if (child_j->is_new_child) {
      const int* parent_B = child_j->enc_B; int ncp = child_j->cells_gtable;
      child_j->cells_gtable = make_new_child_btable(child_j, parent_B, ncp, Sroot);
      child_j->enc_B = Sroot;
    }
    child_j->gtable = Groot;
    if (make_gtable(child_j, e->G_SCALE_FACTOR)) rt_idx = -1;
    const int ncomb = fwd_counts[j]; int k = 0;
    if (rt_idx == -1) {                                      /* flat:412-489 root re-election */
      int cells_grp = child_j->cells_gtable; int* bt_grp = child_j->enc_B; const double* A_lfr = child_j->A;
      while (k < ncomb) {
        int cp = fwd[j][k++]; mceo_term* ck = terms + cp;
        if (ck->is_new_child) {
          ck->enc_B = bt_grp; ck->cells_gtable = cells_grp;
          update_btable(A_lfr, ck->enc_B, ck->A, NULL, ck->cells_gtable, m, d);
          ck->gtable = Groot;
        } else {
          bt_grp = ck->enc_B;
          cells_grp = ck->cells_gtable;      /* equal / fewer / more: in all three cases the group count becomes child_k's */
          ck->gtable = Groot;
        }
        if (!make_gtable(ck, e->G_SCALE_FACTOR)) { rt_idx = cp; child_j = ck; break; }
        else A_lfr = ck->A;
      }
    }
    while (k < ncomb) {                                      /* flat:491-550 members */
      int cp = fwd[j][k++]; mceo_term* ck = terms + cp;
      if (ck->is_new_child) {
        ck->cells_gtable = child_j->cells_gtable; ck->enc_B = Smem;
        update_btable(child_j->A, child_j->enc_B, ck->A, ck->enc_B, ck->cells_gtable, m, d);
      } else if (ck->cells_gtable != child_j->cells_gtable) {
        if (ck->cells_gtable > child_j->cells_gtable) {
          ck->cells_gtable = child_j->cells_gtable;
          update_btable(child_j->A, child_j->enc_B, ck->A, ck->enc_B, ck->cells_gtable, m, d);
        } else {
          ck->cells_gtable = child_j->cells_gtable; ck->enc_B = Smem;
          update_btable(child_j->A, child_j->enc_B, ck->A, ck->enc_B, ck->cells_gtable, m, d);
        }
      }
      ck->gtable = Gmem;
      if (!make_gtable(ck, e->G_SCALE_FACTOR)) add_gtables(child_j, ck);
    }
    if (rt_idx != -1) keep_term(e, ng, ftr_terms + nred++, child_j);
    else nrem++;
  }
  free(Sroot); free(Smem); free(Groot); free(Gmem);
  *Nt_reduced = nred; *Nt_removed = nrem;
}

/* ---- est:981-1177 fast_term_reduction_and_create_gtables ---- */
static void ftr_and_gtables(mceo* e) {
  if (e->skip_post_mu) return;
  mceo_arena* ng = &e->gen[1 - e->cur_gen];
  mceo_term** ftr_dp = (mceo_term**)calloc(e->shape_range, sizeof(mceo_term*));
  int Nt_reduced = 0, Nt_removed = 0;
  for (int m = 0; m < e->shape_range; m++) {
    const int n = e->terms_per_shape[m];
    if (n > 0) {
      mceo_term* terms = e->terms_dp[m];
      int* F = (int*)malloc(sizeof(int) * n);
      for (int i = 0; i < n; i++) F[i] = i;
      fast_term_reduction(e, terms, F, n, m, e->d);
      if (e->after_ftr_shape) e->after_ftr_shape(e, m, F, n, e->after_ftr_arg);
      /* ForwardFlagArray, util:360-408 */
      int* cnt = (int*)calloc(n, sizeof(int)); int** fwd = (int**)calloc(n, sizeof(int*)); int nroots = 0;
      for (int j = 0; j < n; j++) { if (F[j] != j) cnt[F[j]]++; else nroots++; }
      for (int j = 0; j < n; j++) if (cnt[j]) { fwd[j] = (int*)malloc(sizeof(int) * cnt[j]); cnt[j] = 0; }
      for (int j = 0; j < n; j++) if (F[j] != j) fwd[F[j]][cnt[F[j]]++] = j;
      ftr_dp[m] = (mceo_term*)malloc(sizeof(mceo_term) * (nroots ? nroots : 1));
      int nred = 0, nrem = 0;
      make_gtables(e, ng, &nred, &nrem, terms, ftr_dp[m], F, fwd, cnt, n, m, e->d);
      Nt_reduced += nred; Nt_removed += nrem; e->terms_per_shape[m] = nred;
      for (int j = 0; j < n; j++) free(fwd[j]);
      free(fwd); free(cnt); free(F);
    } else ftr_dp[m] = (mceo_term*)malloc(1);
    free(e->terms_dp[m]);
  }
  e->Nt = Nt_reduced; e->Nt_removed_last = Nt_removed;
  free(e->terms_dp); e->terms_dp = ftr_dp;
  /* swap generations (swap_gtables util:895): the old parents' tables and this step's scratch die here */
  arena_reset(&e->gen[e->cur_gen]); arena_reset(&e->step_arena);
  e->cur_gen = 1 - e->cur_gen;
  if (e->print_basic_info) compute_moments(e, 0); else e->fz = 1;
}

/* ---- est:604-832 step_tp_to_muc (serial) ---- */
static void step_tp_to_muc(mceo* e, double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H,
                           double gamma, const double* B, const double* u) {
  const int d = e->d;
  double tmp_Gamma[d * (e->pncc > 0 ? e->pncc : 1)], tmp_beta[e->pncc > 0 ? e->pncc : 1];
  int tmp_pncc = 0;
  const int with_tp = (e->master_step % e->p) == 0;
  e->fz = 0; memset(e->mean, 0, d * sizeof(double complex)); memset(e->var, 0, d * d * sizeof(double complex));
  if (with_tp) tmp_pncc = precoalign_Gamma_beta(Gamma, beta, e->pncc, d, tmp_Gamma, tmp_beta);
  int* new_tps = (int*)calloc(e->shape_range, sizeof(int));
  size_t Nt_alloc = 0;
  for (int m = 1; m < e->shape_range; m++) Nt_alloc += (size_t)e->terms_per_shape[m] * (m + tmp_pncc);
  mceo_term* new_children = (mceo_term*)malloc(sizeof(mceo_term) * (e->skip_post_mu ? e->shape_range + 1 : Nt_alloc + 1));
  size_t Nt_new = 0;
  const int MS = e->shape_range - 1;
  for (int m = 1; m < e->shape_range; m++) {
    const int n = e->terms_per_shape[m];
    mceo_term* terms = e->terms_dp[m];
    for (int i = 0; i < n; i++) {
      mceo_term* parent = terms + i;
      /* transfer_term_to_workspace, term:787-798 */
      double* wA = (double*)arena_alloc(&e->step_arena, sizeof(double) * (MS + 1) * d);
      double* wp = (double*)arena_alloc(&e->step_arena, sizeof(double) * (MS + 1));
      double* wq = (double*)arena_alloc(&e->step_arena, sizeof(double) * (MS + 1));
      double* wb = (double*)arena_alloc(&e->step_arena, sizeof(double) * d);
      memcpy(wA, parent->A, sizeof(double) * parent->m * d); memcpy(wp, parent->p, sizeof(double) * parent->m); memcpy(wb, parent->b, sizeof(double) * d);
      parent->A = wA; parent->p = wp; parent->q = wq; parent->b = wb; parent->c_map = NULL; parent->cs_map = NULL;
      if (with_tp) {
        time_prop(parent, Phi, B, u, e->cmcc);
        int m_tp = tp_coalign(parent, tmp_Gamma, tmp_beta, tmp_pncc);
        if (!e->skip_post_mu) {
          if (parent->m == parent->phc) {
            int* Bp = parent->enc_B;
            parent->enc_B = (int*)arena_alloc(&e->step_arena, sizeof(int) * (parent->cells_gtable_p + 1));
            memcpy(parent->enc_B, Bp, parent->cells_gtable_p * sizeof(int));
          } else {
            parent->enc_B = (int*)arena_alloc(&e->step_arena, sizeof(int) * (cell_count_central(m_tp, d) / 2 + 1));
            make_time_prop_btable(e, parent);
          }
        }
      }
      const int m_pre = parent->m;
      mceo_term* children = e->skip_post_mu ? new_children : new_children + Nt_new;
      int nchild = msmt_update(e, parent, children, msmt, H, gamma, 0, e->skip_post_mu);
      Nt_new += nchild;
      cache_moments(e, parent, children, nchild);
      if (!e->skip_post_mu) {
        normalize_hps(parent, 1);
        new_tps[parent->m]++;
        for (int j = 0; j < nchild; j++) {
          new_tps[mu_coalign(children + j)]++;
          if (!(children[j].m < m_pre)) { children[j].c_map = NULL; children[j].cs_map = NULL; }   /* util:1290-1299 */
        }
      } else new_tps[parent->m] += nchild + 1;
    }
  }
  e->Nt += (int)Nt_new;
  finalize_moments_core(e);
  moments_numerical_check(e);
  if (!e->skip_post_mu) {
    mceo_term** ndp = (mceo_term**)calloc(e->shape_range, sizeof(mceo_term*));
    for (int m = 0; m < e->shape_range; m++) ndp[m] = (mceo_term*)malloc(sizeof(mceo_term) * (new_tps[m] ? new_tps[m] : 1));
    memset(new_tps, 0, e->shape_range * sizeof(int));
    for (int s = 0; s < e->shape_range; s++) {
      for (int i = 0; i < e->terms_per_shape[s]; i++) { int m = e->terms_dp[s][i].m; ndp[m][new_tps[m]++] = e->terms_dp[s][i]; }
      free(e->terms_dp[s]);
    }
    for (size_t i = 0; i < Nt_new; i++) { int m = new_children[i].m; ndp[m][new_tps[m]++] = new_children[i]; }
    free(e->terms_dp); e->terms_dp = ndp;
  }
  memcpy(e->terms_per_shape, new_tps, e->shape_range * sizeof(int));
  free(new_tps); free(new_children);
}

/* ---- est:1179-1208 step_first ---- */
static void step_first(mceo* e, double msmt, const double* H, double gamma) {
  const int d = e->d;
  mceo_term* terms = e->terms_dp[d];
  e->Nt = msmt_update(e, terms, terms + 1, msmt, H, gamma, 1, 0) + 1;
  e->terms_per_shape[d] = e->Nt;
  compute_moments(e, 1);
  mceo_arena* ng = &e->gen[1 - e->cur_gen];
  mceo_term kept[d + 1];
  for (int i = 0; i < e->Nt; i++) {
    terms[i].m = d; terms[i].d = d;
    terms[i].cells_gtable = cell_count_central(d, d) / 2;
    terms[i].enc_B = (int*)arena_alloc(&e->step_arena, sizeof(int) * terms[i].cells_gtable);
    terms[i].gtable = (mceo_kcv*)arena_alloc(&e->step_arena, sizeof(mceo_kcv) * terms[i].cells_gtable);
    make_gtable_first(terms + i, e->G_SCALE_FACTOR);
    keep_term(e, ng, kept + i, terms + i);
  }
  memcpy(terms, kept, sizeof(mceo_term) * e->Nt);
  arena_reset(&e->gen[e->cur_gen]); arena_reset(&e->step_arena);
  e->cur_gen = 1 - e->cur_gen;
  if (e->print_basic_info) compute_moments(e, 0); else e->fz = 1;
  memcpy(e->last_mean, e->mean, d * sizeof(double complex)); memcpy(e->last_var, e->var, d * d * sizeof(double complex));
  e->last_fz = e->fz;
}

static void setup_first_term(mceo* e) {           /* term:772-785 */
  const int d = e->d;
  mceo_term* t = e->terms_dp[d];
  memset(t, 0, sizeof(mceo_term) * (d + 1));
  t->m = d; t->d = d;
  t->A = (double*)arena_alloc(&e->step_arena, sizeof(double) * d * d); memcpy(t->A, e->A0, sizeof(double) * d * d);
  t->p = (double*)arena_alloc(&e->step_arena, sizeof(double) * d); memcpy(t->p, e->p0, sizeof(double) * d);
  t->q = (double*)arena_alloc(&e->step_arena, sizeof(double) * d);
  t->b = (double*)arena_alloc(&e->step_arena, sizeof(double) * d); memcpy(t->b, e->b0, sizeof(double) * d);
}

mceo* mceo_create(int d, int cmcc, int pncc, int p, int steps, const double* A0, const double* p0, const double* b0,
                  const double* root_point, const double* b_pert, const int* tr_order) {
  mceo* e = (mceo*)calloc(1, sizeof(mceo));
  e->d = d; e->cmcc = cmcc; e->pncc = pncc; e->p = p; e->Nt = 1; e->master_step = 0;
  e->num_estimation_steps = p * steps;
  int max_hp_shape = d > 1 ? (steps - 1) * pncc + d : d + pncc;          /* est:97 */
  e->shape_range = max_hp_shape + 1;
  e->terms_per_shape = (int*)calloc(e->shape_range, sizeof(int)); e->terms_per_shape[d] = 1;
  e->muc_counts = (int*)calloc(e->shape_range, sizeof(int));
  e->terms_dp = (mceo_term**)calloc(e->shape_range, sizeof(mceo_term*));
  for (int i = 0; i < e->shape_range; i++) e->terms_dp[i] = (mceo_term*)malloc(sizeof(mceo_term) * (i == d ? d + 1 : 1));
  e->mean = (double complex*)calloc(d, sizeof(double complex)); e->var = (double complex*)calloc(d * d, sizeof(double complex));
  e->last_mean = (double complex*)calloc(d, sizeof(double complex)); e->last_var = (double complex*)calloc(d * d, sizeof(double complex));
  e->root_point = (double*)malloc(sizeof(double) * d); memcpy(e->root_point, root_point, sizeof(double) * d);
  e->b_pert = (double*)calloc(max_hp_shape + 1, sizeof(double)); memcpy(e->b_pert, b_pert, sizeof(double) * max_hp_shape);
  for (int i = 0; i < 12; i++) e->tr_order[i] = tr_order ? tr_order[i] : i;
  e->A0 = (double*)malloc(sizeof(double) * d * d); memcpy(e->A0, A0, sizeof(double) * d * d);
  e->p0 = (double*)malloc(sizeof(double) * d); memcpy(e->p0, p0, sizeof(double) * d);
  e->b0 = (double*)malloc(sizeof(double) * d); memcpy(e->b0, b0, sizeof(double) * d);
  setup_first_term(e);
  return e;
}

int mceo_step(mceo* e, double msmt, const double* Phi, const double* Gamma, const double* beta, const double* H,
              double gamma, const double* B, const double* u) {
  if (e->numeric_moment_errors & (1 << ERROR_FZ_NEGATIVE)) return e->numeric_moment_errors;      /* est:1214-1219 */
  if (e->master_step == e->num_estimation_steps) { fprintf(stderr, "[mce_oracle] master_step == num_estimation_steps\n"); exit(1); }
  e->skip_post_mu = (e->master_step == (e->num_estimation_steps - 1));                              /* SKIP_LAST_STEP */
  if (e->master_step == 0) {
    step_first(e, msmt, H, gamma);
    e->Nt_muc = e->Nt; memcpy(e->muc_counts, e->terms_per_shape, e->shape_range * sizeof(int)); e->fz_mu = e->last_fz;
  } else {
    step_tp_to_muc(e, msmt, Phi, Gamma, beta, H, gamma, B, u);
    e->Nt_muc = e->Nt; memcpy(e->muc_counts, e->terms_per_shape, e->shape_range * sizeof(int)); e->fz_mu = e->fz;
    if (e->after_muc) e->after_muc(e, e->after_muc_arg);
    ftr_and_gtables(e);
  }
  e->master_step++;
  return e->numeric_moment_errors;
}

void mceo_shift_b(mceo* e, const double* delta) {         /* est:1365-1383 */
  if (e->skip_post_mu) return;
  for (int m = 1; m < e->shape_range; m++)
    for (int i = 0; i < e->terms_per_shape[m]; i++)
      for (int j = 0; j < e->d; j++) e->terms_dp[m][i].b[j] -= delta[j];
}


/* ---- point-wise 1-D marginal cpdf on a grid -------------------------------------------
 * Restates PointWiseNDimCauchyCPDF::evaluate_1D_marginal_cpdf (cpdf_ndim.hpp:1233-1354) as it is driven by
 * CauchyCPDFGridDispatcher1D::evaluate_point_grid (cpdf_ndim.hpp:2074-2139): the first grid point is evaluated term by
 * term with two G-table lookups and two complex divisions while the per-term cache {a, b, w, s} (cpdf_ndim.hpp:216-222)
 * is filled; every further point is the cached sum  sum_t (a x + b) / (w^2 + (x - s)^2)  in term order.  The grid itself
 * is reset_grid (cpdf_ndim.hpp:2055-2072).  Returns the number of grid points, 0 when the estimator holds no tables
 * (window's last step, SKIP_LAST_STEP) -- xs / ys may be NULL to query the count. */
int mceo_marginal_1d_grid(const mceo* e, int marg_idx, const double* bar_nu, double grid_low, double grid_high, double grid_res,
                          double* xs, double* ys) {
  if (e->master_step < 1 || marg_idx < 0 || marg_idx >= e->d || !(grid_high > grid_low) || !(grid_res > 0)) return -1;
  if (e->master_step == e->num_estimation_steps) return 0;                      /* cpdf_ndim.hpp:2079-2083 */
  const int n_pts = (int)((grid_high - grid_low + grid_res - 1e-15) / grid_res) + 1;
  if (!xs || !ys) return n_pts;
  for (int i = 0; i < n_pts; i++) { double g = grid_low + i * grid_res; if (g > grid_high) g = grid_high; xs[i] = g; }
  const int d = e->d;
  const double norm_factor = creal(e->fz);
  double* cache = (double*)malloc(sizeof(double) * 4 * (size_t)(e->Nt > 0 ? e->Nt : 1));
  int count = 0;
  { /* first point: uncached evaluation + cache set-up (cpdf_ndim.hpp:1283-1340) */
    const double x1 = xs[0];
    double fx = 0;
    for (int m = 1; m < e->shape_range; m++) {
      const int two_to_m_minus1 = 1 << (m - 1), rev_m_mask = (1 << m) - 1;
      for (int i = 0; i < e->terms_per_shape[m]; i++) {
        const mceo_term* t = e->terms_dp[m] + i;
        const double b_c = t->b[marg_idx] - x1;
        double p_cc = 0; int lhs = 0, rhs = 0;
        for (int j = 0; j < m; j++) {
          const double A_cj = t->A[j * d + marg_idx], f = fabs(A_cj);
          p_cc += t->p[j] * f;
          if (f > 1e-15) { if (A_cj > 0) lhs |= 1 << j; else rhs |= 1 << j; }
          else { if (dot_prod(t->A + j * d, bar_nu, d) > 0) lhs |= 1 << j; else rhs |= 1 << j; }
        }
        const double complex gl = g_num_binsearch(lhs, two_to_m_minus1, rev_m_mask, t->gtable_p, t->cells_gtable_p);
        const double complex gr = g_num_binsearch(rhs, two_to_m_minus1, rev_m_mask, t->gtable_p, t->cells_gtable_p);
        const double complex gv = gl / CMPLX(p_cc, b_c) - gr / CMPLX(-p_cc, b_c);
        fx += creal(gv);
        double* c = cache + 4 * (size_t)count++;
        const double s = t->b[marg_idx], w = p_cc, c_kk = creal(gr), d_kk = cimag(gr);
        c[0] = d_kk / M_PI; c[1] = (c_kk * w - d_kk * s) / M_PI; c[2] = w; c[3] = s;
      }
    }
    ys[0] = creal((double complex)(fx * (1.0 / (2.0 * M_PI)) / norm_factor));
  }
  for (int k = 1; k < n_pts; k++) {   /* cached points (cpdf_ndim.hpp:1266-1281) */
    const double x1 = xs[k];
    double fx = 0;
    for (int i = 0; i < count; i++) {
      const double* c = cache + 4 * (size_t)i;
      const double w2 = c[2] * c[2];
      double x1ms = x1 - c[3];
      x1ms *= x1ms;
      fx += (c[0] * x1 + c[1]) / (w2 + x1ms);
    }
    ys[k] = fx / norm_factor;
  }
  free(cache);
  return n_pts;
}

/* ---- point-wise 2-D marginal cpdf on a grid -------------------------------------------
 * Restates PointWiseNDimCauchyCPDF::evaluate_2D_marginal_cpdf (cpdf_ndim.hpp:1356-1455) with its helpers
 * marg2d_extract_2D_HPA (:31-40), marg2d_remove_zeros_and_coalign (:42-139), marg2d_get_cell_wall_angles (:142-170),
 * marg2d_get_SVs (:176-201), marg2d_eval_term_for_cpdf (:1474-1658) and marg2d_cached_eval_term_for_cpdf (:1660-1747),
 * as driven by CauchyCPDFGridDispatcher2D (reset_grid :1816-1848, evaluate_point_grid :1850-1919).  The per-term cache
 * (angles' sines / cosines, gamma reals, G values) is built once; every grid point is the cached sum in term order (the
 * uncached evaluation of the first point performs the same operations on the same values).  Uses libm atan2/sin/cos like
 * the reference.  out = [ny*nx][3] (x, y, z), y-major like the dispatcher.  Returns the number of points. */
#define MCEO_ZERO_HP 32
#define MCEO_COALIGN_MU_EPS 1e-8
#define MCEO_MU_EPS 1e-10
#define MCEO_INTEGRAL_GAMMA_EPS 1e-8
static int cmp_dless(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return x > y ? 1 : (x < y ? -1 : 0); }
typedef struct { int m; double b[2]; double *sin_t, *cos_t, *g1, *g2; double complex* gv; } mceo_c2d;

static double c2d_eval(const mceo_c2d* c, double x1, double x2) {       /* cpdf_ndim.hpp:1660-1747 */
  const double gam1_imag = c->b[0] - x1, gam2_imag = c->b[1] - x2;
  const int check_gamma1 = fabs(gam1_imag) < MCEO_INTEGRAL_GAMMA_EPS;
  double term_integral = 0;
  for (int i = 0; i < c->m; i++) {
    const double sin_t1 = c->sin_t[i], cos_t1 = c->cos_t[i], sin_t2 = c->sin_t[i + 1], cos_t2 = c->cos_t[i + 1];
    double complex gamma1 = CMPLX(c->g1[i], gam1_imag), gamma2 = CMPLX(c->g2[i], gam2_imag), lo, hi, cell;
    int fast = 1;
    if (check_gamma1 && fabs(c->g1[i]) < MCEO_INTEGRAL_GAMMA_EPS) {
      fast = 0;
      if (fabs(c->g2[i]) < MCEO_INTEGRAL_GAMMA_EPS && fabs(gam2_imag) < MCEO_INTEGRAL_GAMMA_EPS) { fprintf(stderr, "marg2d: singular gamma\n"); exit(1); }
    }
    if (fast) {
      gamma2 *= gamma1; gamma1 *= gamma1;
      lo = sin_t1 / (gamma1 * cos_t1 + gamma2 * sin_t1);
      hi = sin_t2 / (gamma1 * cos_t2 + gamma2 * sin_t2);
      cell = hi - lo;
      cell *= c->gv[i];
      term_integral += creal(cell);
    } else {
      lo = (gamma1 * sin_t1 - gamma2 * cos_t1) / (gamma1 * cos_t1 + gamma2 * sin_t1);
      hi = (gamma1 * sin_t2 - gamma2 * cos_t2) / (gamma1 * cos_t2 + gamma2 * sin_t2);
      cell = hi - lo;
      gamma1 *= gamma1; gamma2 *= gamma2;
      cell *= c->gv[i] / (gamma1 + gamma2);
      term_integral += creal(cell);
    }
  }
  return term_integral;
}

int mceo_marginal_2d_grid(const mceo* e, int idx1, int idx2, const double* bar_nu, const double* gx /*lo,hi,res*/, const double* gy,
                          double* out) {
  if (e->master_step < 1 || idx1 < 0 || idx1 >= idx2 || idx2 >= e->d || !(gx[1] > gx[0]) || !(gy[1] > gy[0]) || !(gx[2] > 0) || !(gy[2] > 0)) return -1;
  if (e->master_step == e->num_estimation_steps) return 0;
  const int nx = (int)((gx[1] - gx[0] + gx[2] - 1e-15) / gx[2]) + 1, ny = (int)((gy[1] - gy[0] + gy[2] - 1e-15) / gy[2]) + 1;
  if (!out) return nx * ny;
  const int d = e->d, MS = e->shape_range;
  mceo_c2d* cache = (mceo_c2d*)malloc(sizeof(mceo_c2d) * (size_t)(e->Nt > 0 ? e->Nt : 1));
  int count = 0;
  double* wA = malloc(sizeof(double) * 2 * MS); double* wA2 = malloc(sizeof(double) * 2 * MS); double* wp = malloc(sizeof(double) * MS);
  double* wp2 = malloc(sizeof(double) * MS); int* c_map = malloc(sizeof(int) * MS); int* cs_map = malloc(sizeof(int) * MS);
  int* F = malloc(sizeof(int) * MS); int* F_idxs = malloc(sizeof(int) * MS);
  double* thetas = malloc(sizeof(double) * 2 * MS); double* SVs = malloc(sizeof(double) * MS * MS); double* As = malloc(sizeof(double) * 2 * MS);
  for (int m = 1; m < e->shape_range; m++) {
    for (int ti = 0; ti < e->terms_per_shape[m]; ti++) {
      const mceo_term* t = e->terms_dp[m] + ti;
      double b2[2];
      memcpy(wp, t->p, sizeof(double) * m);
      for (int i = 0; i < m; i++) { wA[2 * i] = t->A[i * d + idx1]; wA[2 * i + 1] = t->A[i * d + idx2]; }     /* :31-40 */
      b2[0] = t->b[idx1]; b2[1] = t->b[idx2];
      /* marg2d_remove_zeros_and_coalign, :42-139 */
      for (int i = 0; i < m; i++) F[i] = 1;
      for (int i = 0; i < m; i++) {
        double* ai = wA + 2 * i; const double f0 = fabs(ai[0]), f1 = fabs(ai[1]);
        if (f0 < MCEO_MU_EPS && f1 < MCEO_MU_EPS) { c_map[i] = MCEO_ZERO_HP; cs_map[i] = MCEO_ZERO_HP; F_idxs[i] = MCEO_ZERO_HP; F[i] = 0; continue; }
        const double sum_a = f0 + f1; ai[0] /= sum_a; ai[1] /= sum_a; wp[i] *= sum_a;
      }
      int mn = 0;
      for (int i = 0; i < m; i++) {
        if (!F[i]) continue;
        c_map[i] = mn; cs_map[i] = 1; F_idxs[i] = i; wp2[mn] = wp[i];
        const double* ai = wA + 2 * i;
        for (int j = i + 1; j < m; j++) {
          if (!F[j]) continue;
          const double* aj = wA + 2 * j;
          if (fabs(ai[0] - aj[0]) < MCEO_COALIGN_MU_EPS && fabs(ai[1] - aj[1]) < MCEO_COALIGN_MU_EPS) { c_map[j] = mn; cs_map[j] = 1; F[j] = 0; F_idxs[j] = i; wp2[mn] += wp[j]; continue; }
          if (fabs(ai[0] + aj[0]) < MCEO_COALIGN_MU_EPS && fabs(ai[1] + aj[1]) < MCEO_COALIGN_MU_EPS) { c_map[j] = mn; cs_map[j] = -1; F[j] = 0; F_idxs[j] = i; wp2[mn] += wp[j]; continue; }
        }
        mn++;
      }
      if (mn < m) { mn = 0; for (int i = 0; i < m; i++) if (F_idxs[i] == i) { wA2[2 * mn] = wA[2 * i]; wA2[2 * mn + 1] = wA[2 * i + 1]; mn++; } }
      else memcpy(wA2, wA, sizeof(double) * 2 * m);
      const int use_maps = mn != m;
      /* marg2d_get_cell_wall_angles, :142-170 */
      for (int i = 0, k = 0; i < mn; i++) {
        const double* a = wA2 + 2 * i; double pt[2];
        if (fabs(a[0]) < fabs(a[1])) { pt[0] = 1; pt[1] = -a[0] / a[1]; } else { pt[0] = -a[1] / a[0]; pt[1] = 1; }
        double t1 = atan2(pt[1], pt[0]);
        if (t1 < 0) t1 += M_PI;
        thetas[k++] = t1; thetas[k++] = t1 + M_PI;
      }
      qsort(thetas, 2 * mn, sizeof(double), cmp_dless);
      /* marg2d_get_SVs (no flip), :176-201 */
      for (int i = 0; i < mn; i++) {
        const double tt = (thetas[i + 1] + thetas[i]) / 2.0, p0 = cos(tt), p1 = sin(tt);
        for (int j = 0; j < mn; j++) { const double sum = wA2[2 * j] * p0 + wA2[2 * j + 1] * p1; SVs[i * mn + j] = sum > 0 ? 1 : -1; }
      }
      for (int i = 0; i < mn; i++) { As[2 * i] = wA2[2 * i] * wp2[i]; As[2 * i + 1] = wA2[2 * i + 1] * wp2[i]; }
      mceo_c2d* c = cache + count++;
      c->m = mn; c->b[0] = b2[0]; c->b[1] = b2[1];
      c->sin_t = malloc(sizeof(double) * (mn + 1)); c->cos_t = malloc(sizeof(double) * (mn + 1));
      c->g1 = malloc(sizeof(double) * (mn + 1)); c->g2 = malloc(sizeof(double) * (mn + 1)); c->gv = malloc(sizeof(double complex) * (mn + 1));
      const int two_to_mp_minus1 = 1 << (m - 1), rev_mp_mask = (1 << m) - 1;
      for (int i = 0; i < mn; i++) {                                                              /* :1548-1600 */
        const double* SV = SVs + i * mn; int enc = 0;
        if (!use_maps) { for (int j = 0; j < mn; j++) if (SV[j] < 0) enc |= 1 << j; }
        else for (int j = 0; j < m; j++) {
          if (c_map[j] == MCEO_ZERO_HP) { if (dot_prod(t->A + j * d, bar_nu, d) < 0) enc |= 1 << j; }
          else { const int sgn = (int)(SV[c_map[j]] * cs_map[j]); if (sgn < 0) enc |= 1 << j; }
        }
        c->gv[i] = g_num_binsearch(enc, two_to_mp_minus1, rev_mp_mask, t->gtable_p, t->cells_gtable_p);
        double g1 = 0, g2 = 0;
        for (int j = 0; j < mn; j++) { g1 -= As[2 * j] * SV[j]; g2 -= As[2 * j + 1] * SV[j]; }
        c->g1[i] = g1; c->g2[i] = g2;
        c->sin_t[i] = sin(thetas[i]); c->cos_t[i] = cos(thetas[i]);
      }
      c->sin_t[mn] = sin(thetas[mn]); c->cos_t[mn] = cos(thetas[mn]);
    }
  }
  const double norm_factor = creal(e->fz), R2PI = 1.0 / (2.0 * M_PI);
  for (int i = 0; i < ny; i++) {
    double y = gy[0] + i * gy[2]; if (y > gy[1]) y = gy[1];
    for (int j = 0; j < nx; j++) {
      double x = gx[0] + j * gx[2]; if (x > gx[1]) x = gx[1];
      double fx = 0;
      for (int k = 0; k < count; k++) fx += c2d_eval(cache + k, x, y);
      double* o = out + 3 * ((size_t)i * nx + j);
      o[0] = x; o[1] = y; o[2] = 2 * fx * R2PI * R2PI / norm_factor;                         /* :1446 */
    }
  }
  for (int k = 0; k < count; k++) { free(cache[k].sin_t); free(cache[k].cos_t); free(cache[k].g1); free(cache[k].g2); free(cache[k].gv); }
  free(cache); free(wA); free(wA2); free(wp); free(wp2); free(c_map); free(cs_map); free(F); free(F_idxs); free(thetas); free(SVs); free(As);
  return nx * ny;
}

void mceo_reset(mceo* e) {                                 /* est:1247-1300 */
  arena_reset(&e->gen[0]); arena_reset(&e->gen[1]); arena_reset(&e->step_arena); e->cur_gen = 0;
  for (int i = 0; i < e->shape_range; i++) { free(e->terms_dp[i]); e->terms_dp[i] = (mceo_term*)malloc(sizeof(mceo_term) * (i == e->d ? e->d + 1 : 1)); }
  memset(e->terms_per_shape, 0, e->shape_range * sizeof(int)); e->terms_per_shape[e->d] = 1;
  e->Nt = 1; e->master_step = 0; e->numeric_moment_errors = 0;
  setup_first_term(e);
}

void mceo_destroy(mceo* e) {
  arena_reset(&e->gen[0]); arena_reset(&e->gen[1]); arena_reset(&e->step_arena);
  for (int i = 0; i < e->shape_range; i++) free(e->terms_dp[i]);
  free(e->terms_dp); free(e->terms_per_shape); free(e->muc_counts); free(e->mean); free(e->var); free(e->last_mean); free(e->last_var);
  free(e->root_point); free(e->b_pert); free(e->A0); free(e->p0); free(e->b0); free(e);
}
