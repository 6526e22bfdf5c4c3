// ref_run.cpp -- TEST INFRASTRUCTURE. Replays a recorded scenario (oracle/mce_io.h) through the
// UNMODIFIED reference estimator and dumps its state after every phase of every step.
//
// It is compiled by oracle/Makefile against the reference headers where they lie
// (/root/reference/include, reached through the symlink overlay oracle/_ref/overlay_cpuN so that
// NUM_CPUS -- a `const int` in cauchy_constants.hpp:69 -- can be generated per variant).  No
// reference source is copied into this repository; the binary lands in oracle/_ref/ (git-ignored).
//
// The only reference logic restated here is the 25-line body of CauchyEstimator::step()
// (cauchy_estimator.hpp:1211-1245), split so that the post-MUC state can be dumped before
// fast_term_reduction_and_create_gtables() consumes it.  Everything it calls is the reference.
#include "cauchy_estimator.hpp"  // resolved through -I oracle/_ref/overlay_cpuN/include
#include "mce_io.h"
#include <chrono>
#include <string>
#include <vector>
#include <algorithm>

static uint64_t mix64(uint64_t k) {
  uint64_t x = (k + 1) * 0x9E3779B97F4A7C15ULL;
  x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 32;
  return x;
}

struct Opts {
  const char* scenario = nullptr;
  const char* out = nullptr;
  int full_upto = 0;      // full term/table dumps for steps <= full_upto
  int dump_F = 1;         // dump serial FTR flag arrays (full steps only)
  int max_steps = 1 << 30;
  int quiet = 1;
  int time_only = 0;      // no dumps at all, just per-step wall time
  int print_info = 0;     // construct the estimator with print_basic_info = true (quirk A.9 iii: moments recomputed after FTR)
};

static void put_i32(FILE* f, const std::string& name, const std::vector<int>& v) {
  mced_put1(f, name.c_str(), MCED_I32, v.size(), v.data());
}

// Dumps every term of shape m after MUC (what FTR consumes).
static void dump_muc_shape(FILE* f, const std::string& pre, CauchyEstimator& est, int m, int MS) {
  const int d = est.d;
  const int n = est.terms_per_shape[m];
  CauchyTerm* terms = est.terms_dp[m];
  std::vector<double> A((size_t)n * m * d), p((size_t)n * m), q((size_t)n * m), b((size_t)n * d), cd((size_t)n * 2);
  std::vector<int> meta((size_t)n * 8);
  std::vector<uint8_t> cmap((size_t)n * MS, 255);
  std::vector<int8_t> csmap((size_t)n * MS, 0);
  for (int i = 0; i < n; i++) {
    CauchyTerm* t = terms + i;
    memcpy(&A[(size_t)i * m * d], t->A, sizeof(double) * m * d);
    memcpy(&p[(size_t)i * m], t->p, sizeof(double) * m);
    memcpy(&q[(size_t)i * m], t->q, sizeof(double) * m);
    memcpy(&b[(size_t)i * d], t->b, sizeof(double) * d);
    cd[2 * i] = t->c_val; cd[2 * i + 1] = t->d_val;
    int* me = &meta[(size_t)i * 8];
    me[0] = t->phc; me[1] = t->pbc; me[2] = t->z; me[3] = t->enc_lhp; me[4] = (int)t->Horthog_flag;
    me[5] = t->is_new_child; me[6] = t->cells_gtable; me[7] = (t->c_map != NULL);
    if (t->c_map != NULL)
      for (int l = 0; l < t->pbc; l++) { cmap[(size_t)i * MS + l] = t->c_map[l]; csmap[(size_t)i * MS + l] = t->cs_map[l]; }
  }
  mced_put2(f, (pre + "/A").c_str(), MCED_F64, n, m * d, A.data());
  mced_put2(f, (pre + "/p").c_str(), MCED_F64, n, m, p.data());
  mced_put2(f, (pre + "/q").c_str(), MCED_F64, n, m, q.data());
  mced_put2(f, (pre + "/b").c_str(), MCED_F64, n, d, b.data());
  mced_put2(f, (pre + "/cd").c_str(), MCED_F64, n, 2, cd.data());
  mced_put2(f, (pre + "/meta").c_str(), MCED_I32, n, 8, meta.data());
  mced_put2(f, (pre + "/cmap").c_str(), MCED_U8, n, MS, cmap.data());
  mced_put2(f, (pre + "/csmap").c_str(), MCED_I8, n, MS, csmap.data());
}

// Serial FTR flag array for shape m, produced by the reference's own free functions
// (term_reduction.hpp:30,159) on private helper buffers; the estimator state is not touched.
static void dump_F_shape(FILE* f, const std::string& pre, CauchyEstimator& est, int m) {
  const int d = est.d;
  const int n = est.terms_per_shape[m];
  FastTermRedHelper h;
  h.init(d, n > 0 ? n : 1);
  memcpy(h.F_TR, h.F, n * sizeof(int));
  build_ordered_point_maps(est.terms_dp[m], h.ordered_points, h.forward_map, h.backward_map, n, d, false);
  fast_term_reduction(est.terms_dp[m], h.F_TR, h.ordered_points, h.forward_map, h.backward_map, REDUCTION_EPS, n, m, d);
  mced_put1(f, (pre + "/F").c_str(), MCED_I32, n, h.F_TR);
  h.deinit();
}

// Terms + parent tables after FTR / G-table construction (the next step's parents).
static void dump_ftr_shape(FILE* f, const std::string& pre, CauchyEstimator& est, int m, bool full) {
  const int d = est.d;
  const int n = est.terms_per_shape[m];
  CauchyTerm* terms = est.terms_dp[m];
  uint64_t sum_cells = 0, hx = 0, hs = 0;
  double sumG = 0, sump = 0, sumb = 0;
  for (int i = 0; i < n; i++) {
    CauchyTerm* t = terms + i;
    uint64_t th = 0;
    for (int c = 0; c < t->cells_gtable_p; c++) {
      th += mix64(t->gtable_p[c].key);
      sumG += cabs(t->gtable_p[c].value);
    }
    sum_cells += t->cells_gtable_p;
    hx ^= th; hs += th * (uint64_t)(i + 1);
    for (int l = 0; l < m; l++) sump += t->p[l];
    for (int l = 0; l < d; l++) sumb += fabs(t->b[l]);
  }
  uint32_t dig[8] = {(uint32_t)n, 0, (uint32_t)sum_cells, (uint32_t)(sum_cells >> 32),
                     (uint32_t)hx, (uint32_t)(hx >> 32), (uint32_t)hs, (uint32_t)(hs >> 32)};
  mced_put1(f, (pre + "/digest").c_str(), MCED_U32, 8, dig);
  double fd[3] = {sumG, sump, sumb};
  mced_put1(f, (pre + "/fdigest").c_str(), MCED_F64, 3, fd);
  if (!full) return;
  std::vector<double> A((size_t)n * m * d), p((size_t)n * m), b((size_t)n * d);
  std::vector<int> cells(n), encB;
  std::vector<uint32_t> keys;
  std::vector<double> G;
  for (int i = 0; i < n; i++) {
    CauchyTerm* t = terms + i;
    memcpy(&A[(size_t)i * m * d], t->A, sizeof(double) * m * d);
    memcpy(&p[(size_t)i * m], t->p, sizeof(double) * m);
    memcpy(&b[(size_t)i * d], t->b, sizeof(double) * d);
    cells[i] = t->cells_gtable_p;
    for (int c = 0; c < t->cells_gtable_p; c++) {
      keys.push_back(t->gtable_p[c].key);
      G.push_back(creal(t->gtable_p[c].value)); G.push_back(cimag(t->gtable_p[c].value));
      encB.push_back(t->enc_B[c]);
    }
  }
  mced_put2(f, (pre + "/A").c_str(), MCED_F64, n, m * d, A.data());
  mced_put2(f, (pre + "/p").c_str(), MCED_F64, n, m, p.data());
  mced_put2(f, (pre + "/b").c_str(), MCED_F64, n, d, b.data());
  mced_put1(f, (pre + "/cells").c_str(), MCED_I32, n, cells.data());
  mced_put1(f, (pre + "/keys").c_str(), MCED_U32, keys.size(), keys.data());
  mced_put1(f, (pre + "/G").c_str(), MCED_C128, keys.size(), G.data());
  mced_put1(f, (pre + "/encB").c_str(), MCED_I32, encB.size(), encB.data());
}

int main(int argc, char** argv) {
  Opts o;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "--full-upto") o.full_upto = atoi(argv[++i]);
    else if (a == "--no-F") o.dump_F = 0;
    else if (a == "--max-steps") o.max_steps = atoi(argv[++i]);
    else if (a == "--verbose") o.quiet = 0;
    else if (a == "--time-only") o.time_only = 1;
    else if (a == "--print-basic-info") o.print_info = 1;
    else if (!o.scenario) o.scenario = argv[i];
    else o.out = argv[i];
  }
  if (!o.scenario || (!o.out && !o.time_only)) {
    fprintf(stderr, "usage: ref_run scenario.mces out.mced [--full-upto K] [--no-F] [--max-steps N] [--time-only]\n");
    return 2;
  }
  mces_scenario sc;
  mces_read(o.scenario, &sc);
  const int d = sc.d;
  set_tr_search_idxs_ordering(sc.tr_order, d < 12 ? d : 12);

  CauchyEstimator est(sc.A0, sc.p0, sc.b0, sc.steps, d, sc.cmcc, sc.pncc, sc.p, o.print_info != 0);
  // Replace the rand()-drawn vectors by the recorded ones so every implementation sees the same values.
  for (int i = 0; i < d; i++) est.root_point[i] = sc.root_point[i];
  const int MS = est.shape_range - 1;
  for (int t = 0; t < NUM_CPUS; t++)
    for (int i = 0; i < MS; i++) est.dce_helper[t].b_pert[i] = sc.b_pert[i];

  FILE* f = o.time_only ? NULL : mced_open(o.out);
  if (f) {
    int hdr[6] = {d, sc.cmcc, sc.pncc, sc.p, sc.steps, NUM_CPUS};
    mced_put1(f, "header", MCED_I32, 6, hdr);
  }
  const int nrec = std::min(sc.n_records, o.max_steps);
  std::vector<double> step_ms;
  for (int k = 0; k < nrec; k++) {
    mces_step* r = sc.rec + k;
    const std::string sp = "s" + std::to_string(k + 1);
    const bool full = (k + 1) <= o.full_upto;
    auto t0 = std::chrono::steady_clock::now();
    // ---- restated body of CauchyEstimator::step(), cauchy_estimator.hpp:1211-1245 ----
    est.set_function_pointers();
    if (est.numeric_moment_errors & (1 << ERROR_FZ_NEGATIVE)) { fprintf(stderr, "ERROR_FZ_NEGATIVE at step %d\n", k + 1); break; }
    if (est.master_step == est.num_estimation_steps) { fprintf(stderr, "master_step == num_estimation_steps\n"); break; }
    est.skip_post_mu = SKIP_LAST_STEP && (est.master_step == (est.num_estimation_steps - 1));
    const bool first = est.master_step == 0;
    const bool with_tp = (est.master_step % est.p) == 0;
    double* Bp = r->has_Bu ? r->B : NULL;
    double* up = r->has_Bu ? r->u : NULL;
    double ms_muc = 0;
    C_COMPLEX_TYPE fz_mu = 0;
    std::vector<int> muc_counts(est.shape_range, 0);
    int Nt_muc = 0;
    if (first) {
      est.step_first(r->msmt, r->H, r->gamma);
      fz_mu = est.last_fz;
      Nt_muc = est.Nt;
      for (int m = 0; m < est.shape_range; m++) muc_counts[m] = est.terms_per_shape[m];
    } else {
      if ((NUM_CPUS == 1) || (est.Nt < MIN_TERMS_PER_THREAD_TP_TO_MUC))
        est.step_tp_to_muc(r->msmt, r->Phi, r->Gamma, r->beta, r->H, r->gamma, Bp, up);
      else
        est.threaded_step_tp_to_muc(r->msmt, r->Phi, r->Gamma, r->beta, r->H, r->gamma, Bp, up);
      ms_muc = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      fz_mu = est.fz;
      Nt_muc = est.Nt;
      for (int m = 0; m < est.shape_range; m++) muc_counts[m] = est.terms_per_shape[m];
      if (f && full && !est.skip_post_mu) {
        for (int m = 1; m < est.shape_range; m++)
          if (est.terms_per_shape[m] > 0) {
            dump_muc_shape(f, sp + "/muc/m" + std::to_string(m), est, m, MS);
            if (o.dump_F) dump_F_shape(f, sp + "/muc/m" + std::to_string(m), est, m);
          }
      }
      t0 = std::chrono::steady_clock::now();
      est.fast_term_reduction_and_create_gtables();
    }
    est.master_step++;
    // ---- end of restated step() body ----
    double ms = ms_muc + std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    step_ms.push_back(ms);
    if (f) {
      std::vector<int> info = {with_tp && !first, est.skip_post_mu, Nt_muc, est.Nt, est.numeric_moment_errors, first};
      put_i32(f, sp + "/info", info);
      put_i32(f, sp + "/muc/counts", muc_counts);
      std::vector<double> mom;
      mom.push_back(creal(fz_mu)); mom.push_back(cimag(fz_mu));
      for (int i = 0; i < d; i++) { mom.push_back(creal(est.conditional_mean[i])); mom.push_back(cimag(est.conditional_mean[i])); }
      for (int i = 0; i < d * d; i++) { mom.push_back(creal(est.conditional_variance[i])); mom.push_back(cimag(est.conditional_variance[i])); }
      mced_put1(f, (sp + "/moments").c_str(), MCED_C128, 1 + d + d * d, mom.data());
      mced_put1(f, (sp + "/gscale").c_str(), MCED_F64, 1, &est.G_SCALE_FACTOR);
      double tms[2] = {ms, ms_muc};
      mced_put1(f, (sp + "/ms").c_str(), MCED_F64, 2, tms);
      if (!est.skip_post_mu) {
        std::vector<int> ftr_counts(est.shape_range, 0);
        for (int m = 0; m < est.shape_range; m++) ftr_counts[m] = est.terms_per_shape[m];
        put_i32(f, sp + "/ftr/counts", ftr_counts);
        for (int m = 1; m < est.shape_range; m++)
          if (est.terms_per_shape[m] > 0) dump_ftr_shape(f, sp + "/ftr/m" + std::to_string(m), est, m, full);
      }
    }
    if (!o.quiet || o.time_only)
      printf("step %d: after MUC %d, after FTR %d, %.3f ms (muc %.3f ms) err=%d\n", k + 1, Nt_muc, est.Nt, ms, ms_muc, est.numeric_moment_errors);
    // finalize_extended_moments (cauchy_estimator.hpp:1358): the recorded delta replays the closed loop open-loop.
    if (r->shift_kind == MCE_SHIFT_OWN_MEAN) {
      double xb[MCE_MAX_D] = {0};
      est.finalize_extended_moments(xb);
    } else if (r->shift_kind == MCE_SHIFT_EXPLICIT) {
      if (!est.skip_post_mu)
        for (int m = 1; m < est.shape_range; m++)
          for (int i = 0; i < est.terms_per_shape[m]; i++) sub_vecs(est.terms_dp[m][i].b, r->delta, d);
    }
  }
  if (f) fclose(f);
  fflush(stdout);
  _exit(0);  // skip the estimator destructor (it is slow and irrelevant for a dump tool)
}
