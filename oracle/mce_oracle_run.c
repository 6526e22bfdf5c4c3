/* mce_oracle_run.c -- TEST INFRASTRUCTURE. Replays a scenario through the plain-C oracle (mce_oracle.c)
 * and writes the same dump layout as oracle/ref_run.cpp, so the two can be diffed array by array. */
#include "mce_io.h"
#include "mce_oracle.h"
#include <math.h>
#include <time.h>

static uint64_t mix64(uint64_t k) {
  uint64_t x = (k + 1) * 0x9E3779B97F4A7C15ULL;
  x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
  return x;
}
static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

typedef struct { FILE* f; int step; int full; int MS; } dump_ctx;

static void name(char* buf, const dump_ctx* c, const char* phase, int m, const char* leaf) {
  sprintf(buf, "s%d/%s/m%d/%s", c->step, phase, m, leaf);
}

static void dump_muc(mceo* e, void* arg) {
  dump_ctx* c = (dump_ctx*)arg;
  if (!c->f || !c->full || e->skip_post_mu) return;
  const int d = e->d, MS = c->MS; char nm[128];
  for (int m = 1; m < e->shape_range; m++) {
    const int n = e->terms_per_shape[m];
    if (n <= 0) continue;
    double* A = malloc(sizeof(double) * n * m * d); double* p = malloc(sizeof(double) * n * m); double* q = malloc(sizeof(double) * n * m);
    double* b = malloc(sizeof(double) * n * d); double* cd = malloc(sizeof(double) * n * 2); int* meta = malloc(sizeof(int) * n * 8);
    uint8_t* cmap = malloc((size_t)n * MS); int8_t* csmap = malloc((size_t)n * MS);
    memset(cmap, 255, (size_t)n * MS); memset(csmap, 0, (size_t)n * MS);
    for (int i = 0; i < n; i++) {
      const mceo_term* t = e->terms_dp[m] + i;
      memcpy(A + (size_t)i * m * d, t->A, sizeof(double) * m * d); memcpy(p + (size_t)i * m, t->p, sizeof(double) * m);
      memcpy(q + (size_t)i * m, t->q, sizeof(double) * m); memcpy(b + (size_t)i * d, t->b, sizeof(double) * d);
      cd[2 * i] = t->c_val; cd[2 * i + 1] = t->d_val;
      int* me = meta + (size_t)i * 8;
      me[0] = t->phc; me[1] = t->pbc; me[2] = t->z; me[3] = t->enc_lhp; me[4] = (int)t->Horthog_flag; me[5] = t->is_new_child;
      me[6] = t->cells_gtable; me[7] = t->c_map != NULL;
      if (t->c_map) for (int l = 0; l < t->pbc; l++) { cmap[(size_t)i * MS + l] = t->c_map[l]; csmap[(size_t)i * MS + l] = t->cs_map[l]; }
    }
    name(nm, c, "muc", m, "A"); mced_put2(c->f, nm, MCED_F64, n, m * d, A);
    name(nm, c, "muc", m, "p"); mced_put2(c->f, nm, MCED_F64, n, m, p);
    name(nm, c, "muc", m, "q"); mced_put2(c->f, nm, MCED_F64, n, m, q);
    name(nm, c, "muc", m, "b"); mced_put2(c->f, nm, MCED_F64, n, d, b);
    name(nm, c, "muc", m, "cd"); mced_put2(c->f, nm, MCED_F64, n, 2, cd);
    name(nm, c, "muc", m, "meta"); mced_put2(c->f, nm, MCED_I32, n, 8, meta);
    name(nm, c, "muc", m, "cmap"); mced_put2(c->f, nm, MCED_U8, n, MS, cmap);
    name(nm, c, "muc", m, "csmap"); mced_put2(c->f, nm, MCED_I8, n, MS, csmap);
    free(A); free(p); free(q); free(b); free(cd); free(meta); free(cmap); free(csmap);
  }
}
static void dump_F(mceo* e, int m, const int* F, int n, void* arg) {
  dump_ctx* c = (dump_ctx*)arg; (void)e;
  if (!c->f || !c->full) return;
  char nm[128]; name(nm, c, "muc", m, "F");
  mced_put1(c->f, nm, MCED_I32, n, F);
}
static void dump_ftr(mceo* e, dump_ctx* c) {
  const int d = e->d; char nm[128];
  for (int m = 1; m < e->shape_range; m++) {
    const int n = e->terms_per_shape[m];
    if (n <= 0) continue;
    uint64_t sum_cells = 0, hx = 0, hs = 0; double sumG = 0, sump = 0, sumb = 0;
    for (int i = 0; i < n; i++) {
      const mceo_term* t = e->terms_dp[m] + i; uint64_t th = 0;
      for (int k = 0; k < t->cells_gtable_p; k++) { th += mix64(t->gtable_p[k].key); sumG += cabs(t->gtable_p[k].value); }
      sum_cells += t->cells_gtable_p; hx ^= th; hs += th * (uint64_t)(i + 1);
      for (int l = 0; l < m; l++) sump += t->p[l];
      for (int l = 0; l < d; l++) sumb += fabs(t->b[l]);
    }
    uint32_t dig[8] = {(uint32_t)n, 0, (uint32_t)sum_cells, (uint32_t)(sum_cells >> 32), (uint32_t)hx, (uint32_t)(hx >> 32), (uint32_t)hs, (uint32_t)(hs >> 32)};
    name(nm, c, "ftr", m, "digest"); mced_put1(c->f, nm, MCED_U32, 8, dig);
    double fd[3] = {sumG, sump, sumb};
    name(nm, c, "ftr", m, "fdigest"); mced_put1(c->f, nm, MCED_F64, 3, fd);
    if (!c->full) continue;
    double* A = malloc(sizeof(double) * n * m * d); double* p = malloc(sizeof(double) * n * m); double* b = malloc(sizeof(double) * n * d);
    int* cells = malloc(sizeof(int) * n);
    uint32_t* keys = malloc(sizeof(uint32_t) * (sum_cells + 1)); double* G = malloc(sizeof(double) * 2 * (sum_cells + 1)); int* encB = malloc(sizeof(int) * (sum_cells + 1));
    size_t o = 0;
    for (int i = 0; i < n; i++) {
      const mceo_term* t = e->terms_dp[m] + i;
      memcpy(A + (size_t)i * m * d, t->A, sizeof(double) * m * d); memcpy(p + (size_t)i * m, t->p, sizeof(double) * m); memcpy(b + (size_t)i * d, t->b, sizeof(double) * d);
      cells[i] = t->cells_gtable_p;
      for (int k = 0; k < t->cells_gtable_p; k++, o++) { keys[o] = t->gtable_p[k].key; G[2 * o] = creal(t->gtable_p[k].value); G[2 * o + 1] = cimag(t->gtable_p[k].value); encB[o] = t->enc_B[k]; }
    }
    name(nm, c, "ftr", m, "A"); mced_put2(c->f, nm, MCED_F64, n, m * d, A);
    name(nm, c, "ftr", m, "p"); mced_put2(c->f, nm, MCED_F64, n, m, p);
    name(nm, c, "ftr", m, "b"); mced_put2(c->f, nm, MCED_F64, n, d, b);
    name(nm, c, "ftr", m, "cells"); mced_put1(c->f, nm, MCED_I32, n, cells);
    name(nm, c, "ftr", m, "keys"); mced_put1(c->f, nm, MCED_U32, sum_cells, keys);
    name(nm, c, "ftr", m, "G"); mced_put1(c->f, nm, MCED_C128, sum_cells, G);
    name(nm, c, "ftr", m, "encB"); mced_put1(c->f, nm, MCED_I32, sum_cells, encB);
    free(A); free(p); free(b); free(cells); free(keys); free(G); free(encB);
  }
}

int main(int argc, char** argv) {
  const char* scen = NULL; const char* out = NULL; int full_upto = 0, max_steps = 1 << 30, verbose = 0, time_only = 0, print_info = 0;
  double cg[3] = {0, 0, 0}; const char* csteps = NULL; double cg2[6] = {0, 0, 0, 0, 0, 0}; int with_2d = 0;   /* --cpdf2d xlo xhi xres ylo yhi yres */   /* --cpdf1d lo hi res step[,step..]: same arrays as oracle/ref_cpdf.cpp */
  for (int i = 1; i < argc; i++) {
    if (!strcmp(argv[i], "--full-upto")) full_upto = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--max-steps")) max_steps = atoi(argv[++i]);
    else if (!strcmp(argv[i], "--verbose")) verbose = 1;
    else if (!strcmp(argv[i], "--time-only")) time_only = 1;
    else if (!strcmp(argv[i], "--print-basic-info")) print_info = 1;
    else if (!strcmp(argv[i], "--no-F")) {}
    else if (!strcmp(argv[i], "--cpdf2d")) { for (int k = 0; k < 6; k++) cg2[k] = atof(argv[i + 1 + k]); with_2d = 1; i += 6; }
    else if (!strcmp(argv[i], "--cpdf1d")) { cg[0] = atof(argv[i + 1]); cg[1] = atof(argv[i + 2]); cg[2] = atof(argv[i + 3]); csteps = argv[i + 4]; i += 4; }
    else if (!scen) scen = argv[i];
    else out = argv[i];
  }
  if (!scen || (!out && !time_only)) { fprintf(stderr, "usage: mce_oracle_run scenario.mces out.mced [--full-upto K] [--max-steps N] [--time-only]\n"); return 2; }
  mces_scenario sc; mces_read(scen, &sc);
  const int d = sc.d;
  mceo* e = mceo_create(d, sc.cmcc, sc.pncc, sc.p, sc.steps, sc.A0, sc.p0, sc.b0, sc.root_point, sc.b_pert, sc.tr_order);
  e->print_basic_info = print_info;
  dump_ctx ctx; ctx.f = time_only ? NULL : mced_open(out); ctx.MS = e->shape_range - 1;
  if (ctx.f) { int hdr[6] = {d, sc.cmcc, sc.pncc, sc.p, sc.steps, 1}; mced_put1(ctx.f, "header", MCED_I32, 6, hdr); }
  e->after_muc = dump_muc; e->after_muc_arg = &ctx; e->after_ftr_shape = dump_F; e->after_ftr_arg = &ctx;
  const int nrec = sc.n_records < max_steps ? sc.n_records : max_steps;
  for (int k = 0; k < nrec; k++) {
    mces_step* r = sc.rec + k; char nm[128];
    ctx.step = k + 1; ctx.full = (k + 1) <= full_upto;
    const int first = e->master_step == 0, with_tp = (e->master_step % e->p) == 0;
    double t0 = now_ms();
    int err = mceo_step(e, r->msmt, r->Phi, r->Gamma, r->beta, r->H, r->gamma, r->has_Bu ? r->B : NULL, r->has_Bu ? r->u : NULL);
    double ms = now_ms() - t0;
    if (ctx.f) {
      int info[6] = {with_tp && !first, e->skip_post_mu, e->Nt_muc, e->Nt, err, first};
      sprintf(nm, "s%d/info", k + 1); mced_put1(ctx.f, nm, MCED_I32, 6, info);
      sprintf(nm, "s%d/muc/counts", k + 1); mced_put1(ctx.f, nm, MCED_I32, e->shape_range, e->muc_counts);
      double mom[2 * (1 + MCE_MAX_D + MCE_MAX_D * MCE_MAX_D)]; int o = 0;
      mom[o++] = creal(e->fz_mu); mom[o++] = cimag(e->fz_mu);
      for (int i = 0; i < d; i++) { mom[o++] = creal(e->mean[i]); mom[o++] = cimag(e->mean[i]); }
      for (int i = 0; i < d * d; i++) { mom[o++] = creal(e->var[i]); mom[o++] = cimag(e->var[i]); }
      sprintf(nm, "s%d/moments", k + 1); mced_put1(ctx.f, nm, MCED_C128, 1 + d + d * d, mom);
      sprintf(nm, "s%d/gscale", k + 1); mced_put1(ctx.f, nm, MCED_F64, 1, &e->G_SCALE_FACTOR);
      double tms[2] = {ms, 0}; sprintf(nm, "s%d/ms", k + 1); mced_put1(ctx.f, nm, MCED_F64, 2, tms);
      if (!e->skip_post_mu) {
        sprintf(nm, "s%d/ftr/counts", k + 1); mced_put1(ctx.f, nm, MCED_I32, e->shape_range, e->terms_per_shape);
        dump_ftr(e, &ctx);
      }
    }
    if (verbose || time_only) printf("step %d: after MUC %d, after FTR %d, %.3f ms err=%d\n", k + 1, e->Nt_muc, e->Nt, ms, err);
    if (r->shift_kind == MCE_SHIFT_OWN_MEAN) { double dl[MCE_MAX_D]; for (int i = 0; i < d; i++) dl[i] = creal(e->mean[i]); mceo_shift_b(e, dl); }
    else if (r->shift_kind == MCE_SHIFT_EXPLICIT) mceo_shift_b(e, r->delta);
    if (csteps && ctx.f) {
      int want = 0; { char buf[256]; strncpy(buf, csteps, 255); buf[255] = 0; for (char* t = strtok(buf, ","); t; t = strtok(NULL, ",")) want |= atoi(t) == k + 1; }
      if (want) {
        double bar_nu[MCE_MAX_D];
        for (int i = 0; i < d; i++) bar_nu[i] = 0.25 + 1.5 * sc.root_point[i] - (int)(1.5 * sc.root_point[i]);
        const int np = mceo_marginal_1d_grid(e, 0, bar_nu, cg[0], cg[1], cg[2], NULL, NULL);
        if (np > 0) {
          double* xs = malloc(sizeof(double) * np); double* ys = malloc(sizeof(double) * np); double* xy = malloc(sizeof(double) * 2 * np);
          for (int idx = 0; idx < d; idx++) {
            mceo_marginal_1d_grid(e, idx, bar_nu, cg[0], cg[1], cg[2], xs, ys);
            for (int i = 0; i < np; i++) { xy[2 * i] = xs[i]; xy[2 * i + 1] = ys[i]; }
            sprintf(nm, "s%d/cpdf1d/i%d", k + 1, idx); mced_put2(ctx.f, nm, MCED_F64, np, 2, xy);
          }
          free(xs); free(ys); free(xy);
        }
        if (with_2d) {
          const int n2 = mceo_marginal_2d_grid(e, 0, 1, bar_nu, cg2, cg2 + 3, NULL);
          if (n2 > 0) {
            double* o3 = malloc(sizeof(double) * 3 * n2);
            for (int a = 0; a < d - 1; a++) {
              const int i1 = (a < d - 2 || d == 2) ? a : 0, i2 = (a < d - 2 || d == 2) ? a + 1 : d - 1;
              mceo_marginal_2d_grid(e, i1, i2, bar_nu, cg2, cg2 + 3, o3);
              sprintf(nm, "s%d/cpdf2d/i%d_%d", k + 1, i1, i2); mced_put2(ctx.f, nm, MCED_F64, n2, 3, o3);
            }
            free(o3);
          }
        }
      }
    }
  }
  if (ctx.f) fclose(ctx.f);
  return 0;
}
