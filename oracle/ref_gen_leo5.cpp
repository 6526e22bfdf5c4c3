// ref_gen_leo5.cpp -- TEST INFRASTRUCTURE. Records the 5-state LEO EMCE example
// (/root/reference/src/leo_satellite_5state.cpp:465-606, BASELINE.json configs[2]) as an open-loop scenario file, exactly
// like ref_gen_leo7.cpp does for the 7-state example: the model, simulator and estimator are the reference's own code
// (the example's translation unit is #included with main() renamed); only the driver loop is restated, without the EKF
// baseline (which draws no random numbers and does not feed the Cauchy estimator).
#define main ref_example_main
#include "src/leo_satellite_5state.cpp"  // resolved through -I oracle/_ref/overlay_cpu1
#undef main
#include "mce_io.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: ref_gen_leo5 out.mces [foo_steps=7]\n"); return 2; }
  const int foo_steps = argc > 2 ? atoi(argv[2]) : 7;
  srand(0);  // leo_satellite_5state.cpp:467-469
  count = 0;
  leo_satellite_5state leo;
  const int n = leo.n, cmcc = 0, p = leo.p;
  int pncc = n;                 // the simulation samples from the 5x5 W first (:476)
  const int sim_num_steps = 300;
  double Phi[n * n], Gamma[n * leo.pncc], H[p * n];
  double x0[n];
  memcpy(x0, leo.x0, n * sizeof(double));
  for (int i = 0; i < 4; i++) x0[i] = random_normal(leo.x0[i], leo.alpha_pv_gauss);
  x0[4] = random_normal(leo.x0[4], leo.alpha_density_gauss);
  double x0_kf[n];
  memcpy(x0_kf, leo.x0, n * sizeof(double));
  KalmanDynamicsUpdateContainer kduc;
  kduc.n = n; kduc.pncc = pncc; kduc.cmcc = cmcc; kduc.p = p; kduc.dt = leo.dt; kduc.step = 0;
  kduc.Phi = Phi; kduc.Gamma = Gamma; kduc.H = H; kduc.B = NULL; kduc.u = NULL;
  kduc.W = leo.Wd; kduc.V = leo.V; kduc.x = x0_kf; kduc.other_stuff = &leo;
  SimulationLogger sim_log(NULL, sim_num_steps, x0, &kduc, &leo_5state_simulation_transition_model, &leo_5state_simulation_measurement_model);
  sim_log.run_simulation_and_log();
  pncc = leo.pncc;

  double x0_ce[n];
  memcpy(x0_ce, sim_log.true_state_history, n * sizeof(double));
  double beta[pncc], gamma[p];
  beta[0] = leo.beta_cauchy;
  for (int i = 0; i < p; i++) gamma[i] = leo.std_dev_gps * leo.GAUSS_TO_CAUCHY;
  leo_5state_transition_model_jacobians(Phi, Gamma, x0_ce, &leo);
  double A0[n * n];
  memcpy(A0, Phi, n * n * sizeof(double));
  reflect_array(A0, n, n);
  double p0[n], b0[n];
  for (int i = 0; i < 4; i++) p0[i] = leo.alpha_pv_cauchy;
  p0[4] = leo.alpha_density_cauchy;
  memset(b0, 0, n * sizeof(double));
  CauchyDynamicsUpdateContainer duc;
  duc.n = n; duc.cmcc = cmcc; duc.p = p; duc.pncc = pncc; duc.x = x0_ce; duc.dt = leo.dt; duc.step = 0;
  duc.Phi = Phi; duc.Gamma = Gamma; duc.u = NULL; duc.B = NULL; duc.H = H; duc.beta = beta; duc.gamma = gamma;
  duc.other_stuff = &leo;
  int ftr_idx_ordering[5] = {3, 2, 4, 1, 0};
  set_tr_search_idxs_ordering(ftr_idx_ordering, 5);

  CauchyEstimator cauchyEst(A0, p0, b0, foo_steps, n, cmcc, pncc, p, false);
  mces_scenario sc;
  memset(&sc, 0, sizeof(sc));
  sc.d = n; sc.cmcc = cmcc; sc.pncc = pncc; sc.p = p; sc.steps = foo_steps; sc.n_records = foo_steps * p;
  for (int i = 0; i < 12; i++) sc.tr_order[i] = i < 5 ? ftr_idx_ordering[i] : i;
  memcpy(sc.root_point, cauchyEst.root_point, n * sizeof(double));
  memcpy(sc.b_pert, cauchyEst.dce_helper[0].b_pert, (cauchyEst.shape_range - 1) * sizeof(double));
  memcpy(sc.A0, A0, sizeof(A0)); memcpy(sc.p0, p0, sizeof(p0)); memcpy(sc.b0, b0, sizeof(b0));
  sc.rec = (mces_step*)calloc(sc.n_records, sizeof(mces_step));
  for (int t = 1; t < NUM_CPUS; t++) memcpy(cauchyEst.dce_helper[t].b_pert, sc.b_pert, (cauchyEst.shape_range - 1) * sizeof(double));
  int k = 0;
  for (int i = 0; i < foo_steps; i++) {
    double* msmts = sim_log.msmt_history + (i + 1) * p;
    ece_leo_5state_transition_model_and_jacobians(&duc);
    for (int j = 0; j < p; j++) {
      double zbar[p];
      ece_leo_5state_measurement_jacobian(&duc);
      ece_leo_5state_measurement_model(&duc, zbar);
      double msmt = msmts[j] - zbar[j];
      mces_step* r = sc.rec + k++;
      r->msmt = msmt; r->gamma = gamma[j];
      memcpy(r->Phi, Phi, n * n * sizeof(double)); memcpy(r->Gamma, Gamma, n * pncc * sizeof(double));
      memcpy(r->beta, beta, pncc * sizeof(double)); memcpy(r->H, H + j * n, n * sizeof(double));
      cauchyEst.step(msmt, Phi, Gamma, beta, H + j * n, gamma[j], NULL, NULL);
      r->shift_kind = MCE_SHIFT_EXPLICIT;
      for (int l = 0; l < n; l++) r->delta[l] = creal(cauchyEst.conditional_mean[l]);
      cauchyEst.finalize_extended_moments(duc.x);
      printf("MU %d: Nt=%d err=%d\n", k, cauchyEst.Nt, cauchyEst.numeric_moment_errors);
    }
  }
  mces_write(argv[1], &sc);
  fflush(stdout);
  _exit(0);
}
