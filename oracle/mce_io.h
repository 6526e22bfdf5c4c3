/* mce_io.h -- scenario / dump file formats shared by the TEST INFRASTRUCTURE:
 *   oracle/mce_oracle.c   (plain-C restatement of the reference path)
 *   oracle/ref_run.cpp    (runner that #includes the real reference headers)
 *   tests/                (python reader: tests/mceio.py)
 * This header is checker-side code; nothing in the product path includes it.
 *
 * Scenario file ("MCES"): an open-loop recording of every argument the caller
 * passes to CauchyEstimator::step() (/root/reference/include/cauchy_estimator.hpp:1211)
 * plus the two random vectors the reference draws with libc rand():
 *   root_point (cauchy_estimator.hpp:125-128) and b_pert (cell_enumeration.hpp:467-470).
 *
 * Dump file ("MCED"): a flat list of named little-endian arrays.
 */
#ifndef MCE_IO_H_
#define MCE_IO_H_

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCES_MAGIC 0x5345434dU /* "MCES" */
#define MCED_MAGIC 0x4445434dU /* "MCED" */
#define MCE_MAX_D 16
#define MCE_MAX_SHAPE 32

enum { MCE_SHIFT_NONE = 0, MCE_SHIFT_OWN_MEAN = 1, MCE_SHIFT_EXPLICIT = 2 };

typedef struct {
  double msmt, gamma;
  double Phi[MCE_MAX_D * MCE_MAX_D];
  double Gamma[MCE_MAX_D * MCE_MAX_D];
  double beta[MCE_MAX_D];
  double H[MCE_MAX_D];
  int has_Bu;
  double B[MCE_MAX_D * MCE_MAX_D];
  double u[MCE_MAX_D];
  int shift_kind;             /* finalize_extended_moments after the step? */
  double delta[MCE_MAX_D];    /* explicit shift b <- b - delta (MCE_SHIFT_EXPLICIT) */
} mces_step;

typedef struct {
  int d, cmcc, pncc, p, steps; /* steps = window steps passed to the constructor */
  int n_records;               /* number of step() calls recorded (<= p*steps)  */
  int tr_order[12];            /* TR_SEARCH_IDXS_ORDERING, cauchy_constants.hpp:66 */
  double root_point[MCE_MAX_D];
  double b_pert[MCE_MAX_SHAPE];
  double A0[MCE_MAX_D * MCE_MAX_D], p0[MCE_MAX_D], b0[MCE_MAX_D];
  mces_step* rec;
} mces_scenario;

static inline void mces__wd(FILE* f, const double* x, int n) { fwrite(x, sizeof(double), (size_t)n, f); }
static inline void mces__rd(FILE* f, double* x, int n) {
  if (fread(x, sizeof(double), (size_t)n, f) != (size_t)n) { fprintf(stderr, "mces: short read\n"); exit(2); }
}
static inline void mces__wi(FILE* f, int v) { int32_t t = v; fwrite(&t, 4, 1, f); }
static inline int mces__ri(FILE* f) {
  int32_t t; if (fread(&t, 4, 1, f) != 1) { fprintf(stderr, "mces: short read\n"); exit(2); } return t;
}

static inline int mces_max_shape(const mces_scenario* s) {
  /* cauchy_estimator.hpp:97 */
  return s->d > 1 ? (s->steps - 1) * s->pncc + s->d : s->d + s->pncc;
}

static inline void mces_write(const char* path, const mces_scenario* s) {
  FILE* f = fopen(path, "wb");
  if (!f) { perror(path); exit(2); }
  uint32_t magic = MCES_MAGIC; fwrite(&magic, 4, 1, f);
  mces__wi(f, 1);
  mces__wi(f, s->d); mces__wi(f, s->cmcc); mces__wi(f, s->pncc); mces__wi(f, s->p);
  mces__wi(f, s->steps); mces__wi(f, s->n_records);
  for (int i = 0; i < 12; i++) mces__wi(f, s->tr_order[i]);
  const int d = s->d, ms = mces_max_shape(s);
  mces__wd(f, s->root_point, d); mces__wi(f, ms); mces__wd(f, s->b_pert, ms);
  mces__wd(f, s->A0, d * d); mces__wd(f, s->p0, d); mces__wd(f, s->b0, d);
  for (int k = 0; k < s->n_records; k++) {
    const mces_step* r = s->rec + k;
    mces__wd(f, &r->msmt, 1); mces__wd(f, &r->gamma, 1);
    mces__wd(f, r->Phi, d * d); mces__wd(f, r->Gamma, d * s->pncc); mces__wd(f, r->beta, s->pncc);
    mces__wd(f, r->H, d);
    mces__wi(f, r->has_Bu);
    if (r->has_Bu) { mces__wd(f, r->B, d * s->cmcc); mces__wd(f, r->u, s->cmcc); }
    mces__wi(f, r->shift_kind);
    mces__wd(f, r->delta, d);
  }
  fclose(f);
}

static inline void mces_read(const char* path, mces_scenario* s) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  memset(s, 0, sizeof(*s));
  uint32_t magic; if (fread(&magic, 4, 1, f) != 1 || magic != MCES_MAGIC) { fprintf(stderr, "%s: not a MCES file\n", path); exit(2); }
  (void)mces__ri(f);
  s->d = mces__ri(f); s->cmcc = mces__ri(f); s->pncc = mces__ri(f); s->p = mces__ri(f);
  s->steps = mces__ri(f); s->n_records = mces__ri(f);
  for (int i = 0; i < 12; i++) s->tr_order[i] = mces__ri(f);
  const int d = s->d;
  mces__rd(f, s->root_point, d); int ms = mces__ri(f); mces__rd(f, s->b_pert, ms);
  mces__rd(f, s->A0, d * d); mces__rd(f, s->p0, d); mces__rd(f, s->b0, d);
  s->rec = (mces_step*)calloc((size_t)s->n_records, sizeof(mces_step));
  for (int k = 0; k < s->n_records; k++) {
    mces_step* r = s->rec + k;
    mces__rd(f, &r->msmt, 1); mces__rd(f, &r->gamma, 1);
    mces__rd(f, r->Phi, d * d); mces__rd(f, r->Gamma, d * s->pncc); mces__rd(f, r->beta, s->pncc);
    mces__rd(f, r->H, d);
    r->has_Bu = mces__ri(f);
    if (r->has_Bu) { mces__rd(f, r->B, d * s->cmcc); mces__rd(f, r->u, s->cmcc); }
    r->shift_kind = mces__ri(f);
    mces__rd(f, r->delta, d);
  }
  fclose(f);
}

/* ---- dump container ---------------------------------------------------- */
enum { MCED_F64 = 0, MCED_I32 = 1, MCED_U32 = 2, MCED_U8 = 3, MCED_I8 = 4, MCED_C128 = 5 };
static inline size_t mced_esize(int dt) {
  switch (dt) { case MCED_F64: return 8; case MCED_I32: case MCED_U32: return 4;
                case MCED_U8: case MCED_I8: return 1; default: return 16; }
}
static inline FILE* mced_open(const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) { perror(path); exit(2); }
  uint32_t magic = MCED_MAGIC; fwrite(&magic, 4, 1, f);
  return f;
}
static inline void mced_put(FILE* f, const char* name, int dtype, int ndim, const uint64_t* dims, const void* data) {
  uint32_t nl = (uint32_t)strlen(name);
  fwrite(&nl, 4, 1, f); fwrite(name, 1, nl, f);
  uint32_t dt = (uint32_t)dtype, nd = (uint32_t)ndim;
  fwrite(&dt, 4, 1, f); fwrite(&nd, 4, 1, f);
  size_t n = 1;
  for (int i = 0; i < ndim; i++) { fwrite(&dims[i], 8, 1, f); n *= (size_t)dims[i]; }
  if (n) fwrite(data, mced_esize(dtype), n, f);
}
static inline void mced_put1(FILE* f, const char* name, int dtype, uint64_t n, const void* data) {
  mced_put(f, name, dtype, 1, &n, data);
}
static inline void mced_put2(FILE* f, const char* name, int dtype, uint64_t n0, uint64_t n1, const void* data) {
  uint64_t dims[2] = {n0, n1};
  mced_put(f, name, dtype, 2, dims, data);
}

#ifdef __cplusplus
}
#endif
#endif /* MCE_IO_H_ */
