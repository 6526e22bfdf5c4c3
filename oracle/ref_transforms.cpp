// ref_transforms.cpp -- TEST INFRASTRUCTURE. Golden vectors for the small transforms of the estimator:
// CauchyEstimator::deterministic_time_prop (cauchy_estimator.hpp:1331-1355) and shift_cf_by_bias (:1312-1328).
// Replays `k` steps of a scenario through the UNMODIFIED reference, then applies
//     deterministic_time_prop(T, NULL, NULL)   [T = I + 0.1 Phi of record k]      -> dump "t1/..."
//     shift_cf_by_bias(bias)                   [bias_j = 0.01 (j + 1)]            -> dump "t2/..."
//     deterministic_time_prop(T, B, u)         [only when the scenario has B, u]  -> dump "t3/..."
// and replays the remaining records, dumping the moments of every later step ("s<i>/moments").
// Compiled by oracle/Makefile against the reference headers where they lie; nothing of the reference is copied.
#include "cauchy_estimator.hpp"
#include "mce_io.h"
#include <string>
#include <vector>

static void dump_terms(FILE* f, const std::string& pre, CauchyEstimator& est) {
  const int d = est.d;
  for (int m = 1; m < est.shape_range; m++) {
    const int n = est.terms_per_shape[m];
    if (n <= 0) continue;
    std::vector<double> A((size_t)n * m * d), p((size_t)n * m), b((size_t)n * d);
    for (int i = 0; i < n; i++) {
      CauchyTerm* t = est.terms_dp[m] + i;
      memcpy(&A[(size_t)i * m * d], t->A, sizeof(double) * m * d);
      memcpy(&p[(size_t)i * m], t->p, sizeof(double) * m);
      memcpy(&b[(size_t)i * d], t->b, sizeof(double) * d);
    }
    const std::string q = pre + "/m" + std::to_string(m);
    mced_put2(f, (q + "/A").c_str(), MCED_F64, n, m * d, A.data());
    mced_put2(f, (q + "/p").c_str(), MCED_F64, n, m, p.data());
    mced_put2(f, (q + "/b").c_str(), MCED_F64, n, d, b.data());
  }
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: ref_transforms scenario.mces out.mced k\n"); return 2; }
  mces_scenario sc;
  mces_read(argv[1], &sc);
  const int d = sc.d, k0 = atoi(argv[3]);
  set_tr_search_idxs_ordering(sc.tr_order, d < 12 ? d : 12);
  CauchyEstimator est(sc.A0, sc.p0, sc.b0, sc.steps, d, sc.cmcc, sc.pncc, sc.p, false);
  for (int i = 0; i < d; i++) est.root_point[i] = sc.root_point[i];
  const int MS = est.shape_range - 1;
  for (int t = 0; t < NUM_CPUS; t++)
    for (int i = 0; i < MS; i++) est.dce_helper[t].b_pert[i] = sc.b_pert[i];
  FILE* f = mced_open(argv[2]);
  int hdr[6] = {d, sc.cmcc, sc.pncc, sc.p, sc.steps, NUM_CPUS};
  mced_put1(f, "header", MCED_I32, 6, hdr);
  auto step = [&](int k) {
    mces_step* r = sc.rec + k;
    est.step(r->msmt, r->Phi, r->Gamma, r->beta, r->H, r->gamma, r->has_Bu ? r->B : NULL, r->has_Bu ? r->u : NULL);
    if (r->shift_kind == MCE_SHIFT_OWN_MEAN) { double xb[MCE_MAX_D] = {0}; est.finalize_extended_moments(xb); }
    else if (r->shift_kind == MCE_SHIFT_EXPLICIT && !est.skip_post_mu)
      for (int m = 1; m < est.shape_range; m++)
        for (int i = 0; i < est.terms_per_shape[m]; i++) sub_vecs(est.terms_dp[m][i].b, r->delta, d);
  };
  for (int k = 0; k < k0; k++) step(k);
  double T[MCE_MAX_D * MCE_MAX_D], bias[MCE_MAX_D];
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++) T[i * d + j] = (i == j ? 1.0 : 0.0) + 0.1 * sc.rec[k0].Phi[i * d + j];
  for (int j = 0; j < d; j++) bias[j] = 0.01 * (j + 1);
  mced_put1(f, "T", MCED_F64, d * d, T);
  mced_put1(f, "bias", MCED_F64, d, bias);
  est.deterministic_time_prop(T, NULL, NULL);
  dump_terms(f, "t1", est);
  est.shift_cf_by_bias(bias);
  dump_terms(f, "t2", est);
  if (sc.rec[k0].has_Bu) {
    est.deterministic_time_prop(T, sc.rec[k0].B, sc.rec[k0].u);
    dump_terms(f, "t3", est);
  }
  for (int k = k0; k < sc.n_records; k++) {
    step(k);
    std::vector<double> mom;
    mom.push_back(creal(est.fz)); mom.push_back(cimag(est.fz));
    for (int i = 0; i < d; i++) { mom.push_back(creal(est.conditional_mean[i])); mom.push_back(cimag(est.conditional_mean[i])); }
    for (int i = 0; i < d * d; i++) { mom.push_back(creal(est.conditional_variance[i])); mom.push_back(cimag(est.conditional_variance[i])); }
    mced_put1(f, ("s" + std::to_string(k + 1) + "/moments").c_str(), MCED_C128, 1 + d + d * d, mom.data());
    int info[2] = {est.Nt, est.numeric_moment_errors};
    mced_put1(f, ("s" + std::to_string(k + 1) + "/info").c_str(), MCED_I32, 2, info);
  }
  fclose(f);
  fflush(stdout);
  _exit(0);
}
