// ref_cpdf.cpp -- TEST INFRASTRUCTURE. Replays a recorded scenario (oracle/mce_io.h) through the UNMODIFIED reference
// estimator and, after each requested step, evaluates the reference's own point-wise 1-D marginal cpdf on a grid
// (CauchyCPDFGridDispatcher1D::evaluate_point_grid, cpdf_ndim.hpp:2074-2139, one thread) for every state index.
// Output: an MCED dump with   s<k>/cpdf1d/i<idx>  = [n_points][2] (x, y)   and   cpdf1d/bar_nu, cpdf1d/grid.
//
// Compiled by oracle/Makefile against the reference headers where they lie (symlink overlay, NUM_CPUS = 1); nothing
// of the reference is copied into this repository.  The only thing set from outside is `bar_nu`, which the reference
// draws with libc rand() (cpdf_ndim.hpp:385-389): the recorded vector replaces the draw so that every implementation
// sees the same values.
#include "cpdf_ndim.hpp"  // resolved through -I oracle/_ref/overlay_cpu1/include
#include "mce_io.h"
#include <string>
#include <vector>

int main(int argc, char** argv) {
  if (argc < 7) {
    fprintf(stderr, "usage: ref_cpdf scenario.mces out.mced grid_low grid_high grid_res step[,step...] [--time]\n");
    return 2;
  }
  const char* scen = argv[1];
  const char* out = argv[2];
  const double glo = atof(argv[3]), ghi = atof(argv[4]), gres = atof(argv[5]);
  std::vector<int> steps;
  { char* s = strdup(argv[6]); for (char* t = strtok(s, ","); t; t = strtok(NULL, ",")) steps.push_back(atoi(t)); }
  bool timing = false;
  // --2d xlo xhi xres ylo yhi yres : additionally evaluate the 2-D marginal of every state pair (i, i+1) and (0, d-1)
  double g2[6] = {0, 0, 0, 0, 0, 0}; bool with_2d = false;
  for (int i = 7; i < argc; i++) {
    if (std::string(argv[i]) == "--time") timing = true;
    else if (std::string(argv[i]) == "--2d" && i + 6 < argc) { for (int k = 0; k < 6; k++) g2[k] = atof(argv[i + 1 + k]); with_2d = true; i += 6; }
  }

  mces_scenario sc;
  mces_read(scen, &sc);
  const int d = sc.d;
  set_tr_search_idxs_ordering(sc.tr_order, d < 12 ? d : 12);
  CauchyEstimator est(sc.A0, sc.p0, sc.b0, sc.steps, d, sc.cmcc, sc.pncc, sc.p, false);
  for (int i = 0; i < d; i++) est.root_point[i] = sc.root_point[i];
  const int MS = est.shape_range - 1;
  for (int t = 0; t < NUM_CPUS; t++)
    for (int i = 0; i < MS; i++) est.dce_helper[t].b_pert[i] = sc.b_pert[i];

  PointWiseNDimCauchyCPDF cpdf(&est);
  double bar_nu[MCE_MAX_D];
  for (int i = 0; i < d; i++) { bar_nu[i] = 0.25 + 1.5 * sc.root_point[i] - (int)(1.5 * sc.root_point[i]); cpdf.bar_nu[i] = bar_nu[i]; }
  CauchyCPDFGridDispatcher1D grid(&cpdf, glo, ghi, gres, NULL);
  CauchyCPDFGridDispatcher2D* grid2 = with_2d ? new CauchyCPDFGridDispatcher2D(&cpdf, g2[0], g2[1], g2[2], g2[3], g2[4], g2[5], NULL) : NULL;

  FILE* f = mced_open(out);
  int hdr[6] = {d, sc.cmcc, sc.pncc, sc.p, sc.steps, NUM_CPUS};
  mced_put1(f, "header", MCED_I32, 6, hdr);
  mced_put1(f, "cpdf1d/bar_nu", MCED_F64, d, bar_nu);
  double g3[3] = {glo, ghi, gres};
  mced_put1(f, "cpdf1d/grid", MCED_F64, 3, g3);
  if (with_2d) mced_put1(f, "cpdf2d/grid", MCED_F64, 6, g2);

  int last = 0; for (int s : steps) last = s > last ? s : last;
  for (int k = 0; k < sc.n_records && k < last; k++) {
    mces_step* r = sc.rec + k;
    double* Bp = r->has_Bu ? r->B : NULL;
    double* up = r->has_Bu ? r->u : NULL;
    est.step(r->msmt, r->Phi, r->Gamma, r->beta, r->H, r->gamma, Bp, up);
    if (r->shift_kind == MCE_SHIFT_OWN_MEAN) {
      double xb[MCE_MAX_D] = {0};
      est.finalize_extended_moments(xb);
    } else if (r->shift_kind == MCE_SHIFT_EXPLICIT) {
      if (!est.skip_post_mu)
        for (int m = 1; m < est.shape_range; m++)
          for (int i = 0; i < est.terms_per_shape[m]; i++) sub_vecs(est.terms_dp[m][i].b, r->delta, d);
    }
    bool want = false;
    for (int s : steps) want |= (s == k + 1);
    if (!want) continue;
    for (int idx = 0; idx < d; idx++) {
      CPUTimer tmr; tmr.tic();
      if (grid.evaluate_point_grid(idx, 1, false)) continue;   // the window's last step has no tables (SKIP_LAST_STEP)
      tmr.toc(false);
      if (timing) printf("step %d idx %d: %d terms x %d points in %d ms\n", k + 1, idx, est.Nt, grid.num_grid_points, tmr.cpu_time_used);
      const std::string name = "s" + std::to_string(k + 1) + "/cpdf1d/i" + std::to_string(idx);
      mced_put2(f, name.c_str(), MCED_F64, grid.num_grid_points, 2, (double*)grid.points);
    }
    if (with_2d) {
      for (int a = 0; a < d - 1; a++) {
        const int i1 = a < d - 2 || d == 2 ? a : 0, i2 = a < d - 2 || d == 2 ? a + 1 : d - 1;     // (0,1), (1,2), ..., and (0, d-1) last
        CPUTimer tmr; tmr.tic();
        if (grid2->evaluate_point_grid(i1, i2, 1, false)) continue;
        tmr.toc(false);
        if (timing) printf("step %d pair %d,%d: %d terms x %d points in %d ms\n", k + 1, i1, i2, est.Nt, grid2->num_grid_points, tmr.cpu_time_used);
        const std::string name = "s" + std::to_string(k + 1) + "/cpdf2d/i" + std::to_string(i1) + "_" + std::to_string(i2);
        mced_put2(f, name.c_str(), MCED_F64, grid2->num_grid_points, 3, (double*)grid2->points);
      }
    }
  }
  fclose(f);
  fflush(stdout);
  _exit(0);
}
