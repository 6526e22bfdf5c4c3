// ref_gen_homing.cpp -- TEST INFRASTRUCTURE. Records the reference's homing-missile EMCE example
// (/root/reference/src/homing_missile.cpp, BASELINE.json configs[1]) as an open-loop scenario file: the example is CLOSED
// loop (the guidance command is computed from the estimate and changes the next measurement), so the reference estimator
// itself runs in the loop here and every argument it is given is written down.  The model, simulator, guidance law and
// estimator are the reference's own code (the example's translation unit is #included with main() renamed); only the
// driver loop of its single-window test (homing_missile.cpp:357-455) is restated, with the author's recommended seed
// (:416, :586-588) and the window depth of the sliding-window run (8).  During the first `depth` steps of the windowed
// example (test_homing_missile, :457-716) the fullest window IS this estimator, so the recording equals its inputs.
#define main ref_example_main
#include "src/homing_missile.cpp"  // resolved through -I oracle/_ref/overlay_cpu1
#undef main
#include "mce_io.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_gen_homing out.mces [depth=8]\n"); return 2; }
  const int foo_steps = argc > 2 ? atoi(argv[2]) : 8;
  const double SCALE_BETA = 5.0, RADAR_SAS_PARAM = 1.3;          // defaults of test_homing_missile (:463-464)
  const int n = 3, p = 1, cmcc = 1, pncc = 1, sim_num_steps = 99, total_steps = sim_num_steps + 1;
  BallisticMissileConstants bsc(RADAR_SAS_PARAM);
  double Phi[n * n], Gamma[n * pncc], B[n * cmcc], H[p * n], u_feedback[cmcc];
  const double sigma_w0 = sqrt((2.0 / bsc.tau * bsc.E_at2) / bsc.dt);
  const double sigma_v0 = sqrt((bsc.R1 + bsc.R2 / pow(bsc.t_f, 2)) / bsc.dt);
  int ftr_ordering[3] = {1, 2, 0};
  set_tr_search_idxs_ordering(ftr_ordering, n);
  double xhat_ce[n], x_ce[n];
  double beta[pncc] = {sigma_w0 * GAUSS_TO_CAUCHY_NOISE / SCALE_BETA};
  double gamma[p], A0[n * n];
  double p0[n] = {sqrt(bsc.E_yt2) * GAUSS_TO_CAUCHY_NOISE, sqrt(bsc.E_vt2) * GAUSS_TO_CAUCHY_NOISE, sqrt(bsc.E_at2) * GAUSS_TO_CAUCHY_NOISE};
  double b0[n] = {0.0, 0.0, 0.0};
  CauchyDynamicsUpdateContainer duc;
  duc.n = n; duc.cmcc = cmcc; duc.p = p; duc.pncc = pncc; duc.dt = bsc.dt; duc.step = 1;
  duc.Phi = Phi; duc.B = B; duc.Gamma = Gamma; duc.H = H; duc.beta = beta; duc.gamma = gamma; duc.x = x_ce; duc.u = u_feedback;
  duc.other_stuff = &bsc;
  HomingSimulation hs(total_steps, &bsc);
  double z[p];
  srand(1658778374u);
  hs.reset_counters();
  hs.simulate_telegraph_process_noise_and_measurement_noise();
  memset(u_feedback, 0, cmcc * sizeof(double));
  memset(x_ce, 0, n * sizeof(double)); memset(xhat_ce, 0, n * sizeof(double));
  init_nonlinear_missile_dynamics(x_ce, Phi, B, Gamma, H, bsc.tau, bsc.Vc, bsc.t_f, bsc.dt, n);
  memcpy(A0, Phi, n * n * sizeof(double));
  reflect_array(A0, n, n);
  gamma[0] = sigma_v0 * _RADAR_TO_CAUCHY_NOISE;
  duc.step = 1;
  CauchyEstimator cauchyEst(A0, p0, b0, foo_steps, n, 0, pncc, p, false);   // extended: controls enter through the deterministic part (cauchy_windows.hpp:429)
  mces_scenario sc;
  memset(&sc, 0, sizeof(sc));
  sc.d = n; sc.cmcc = 0; sc.pncc = pncc; sc.p = p; sc.steps = foo_steps; sc.n_records = foo_steps * p;
  for (int i = 0; i < 12; i++) sc.tr_order[i] = i < 3 ? ftr_ordering[i] : i;
  memcpy(sc.root_point, cauchyEst.root_point, n * sizeof(double));
  memcpy(sc.b_pert, cauchyEst.dce_helper[0].b_pert, (cauchyEst.shape_range - 1) * sizeof(double));
  memcpy(sc.A0, A0, sizeof(A0)); memcpy(sc.p0, p0, sizeof(p0)); memcpy(sc.b0, b0, sizeof(b0));
  sc.rec = (mces_step*)calloc(sc.n_records, sizeof(mces_step));
  for (int t = 1; t < NUM_CPUS; t++) memcpy(cauchyEst.dce_helper[t].b_pert, sc.b_pert, (cauchyEst.shape_range - 1) * sizeof(double));
  int k = 0;
  for (int i = 0; i < foo_steps; i++) {
    hs.step_simulation(xhat_ce, u_feedback, &duc, NULL, NULL, z, NULL);
    if (i > 0) ballistic_missile_nonlinear_radar_full_update_callback(&duc);
    for (int j = 0; j < p; j++) {
      double zbar[p];
      ballistic_missile_ece_msmt_model(&duc, zbar);
      ballistic_missile_nonlinear_radar_msmt_update_callback(&duc);
      const double msmt = z[j] - zbar[j];
      mces_step* r = sc.rec + k++;
      r->msmt = msmt; r->gamma = gamma[j];
      memcpy(r->Phi, Phi, n * n * sizeof(double)); memcpy(r->Gamma, Gamma, n * pncc * sizeof(double));
      memcpy(r->beta, beta, pncc * sizeof(double)); memcpy(r->H, H + j * n, n * sizeof(double));
      cauchyEst.step(msmt, Phi, Gamma, beta, H + j * n, gamma[j], NULL, NULL);
      r->shift_kind = MCE_SHIFT_EXPLICIT;
      for (int l = 0; l < n; l++) r->delta[l] = creal(cauchyEst.conditional_mean[l]);
      cauchyEst.finalize_extended_moments(duc.x);
      // the windowed example feeds the bank's estimate back into the guidance law (:690-697); for these steps that is this estimator's
      memcpy(xhat_ce, duc.x, n * sizeof(double));
      printf("MU %d: Nt=%d err=%d xhat = %.16f %.16f %.16f\n", k, cauchyEst.Nt, cauchyEst.numeric_moment_errors, duc.x[0], duc.x[1], duc.x[2]);
    }
  }
  mces_write(argv[1], &sc);
  fflush(stdout);
  _exit(0);
}
