// winbank_dropin.cpp -- the reference's sliding-window bank (SlidingWindowManager, cauchy_windows.hpp:729-1165) on the
// inputs of its own example src/window_manager.cpp:5-103 (3-state system, srand(11), 200 simulated steps, 8 windows),
// with a log directory so that the bank's outputs can be compared: the example itself passes log_dir = NULL and writes
// nothing.  Compiled twice like pycauchy_dropin.cpp (reference NUM_CPUS = 1 build -> golden; drop-in header + libmce_b200.so).
// usage: winbank_dropin <log_dir> [num_steps]
#include "../include/cauchy_windows.hpp"

int main(int argc, char** argv)
{
    if(argc < 2) { printf("usage: %s <log_dir> [num_steps]\n", argv[0]); return 2; }
    const int n = 3, pncc = 1, cmcc = 0, p = 1;
    double Phi[n*n] = {1.4, -0.6, -1.0,  -0.2,  1.0,  0.5,  0.6, -0.6, -0.2};
    double Gamma[n*pncc] = {.1, 0.3, -0.2};
    double H[n] = {1.0, 0.5, 0.2};
    double beta[pncc] = {0.1};
    double gamma[p] = {0.2};
    double A0[n*n] = {1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0};
    double p0[n] = {0.10, 0.08, 0.05};
    double b0[n] = {0, 0, 0};
    srand(11);
    CauchyDynamicsUpdateContainer duc;
    duc.n = n; duc.pncc = pncc; duc.p = p; duc.cmcc = cmcc;
    duc.Phi = Phi; duc.Gamma = Gamma; duc.H = H; duc.B = NULL; duc.u = NULL;
    duc.beta = beta; duc.gamma = gamma;
    duc.step = 0; duc.dt = 0; duc.other_stuff = NULL; duc.x = NULL;
    const int num_steps = argc > 2 ? atoi(argv[2]) : 200;
    SimulationLogger sim_log(NULL, num_steps, b0, &duc, cauchy_lti_transition_model, cauchy_lti_measurement_model);
    sim_log.run_simulation_and_log();
    const int total_steps = num_steps + 1, num_windows = 8;
    SlidingWindowManager swm(num_windows, total_steps, A0, p0, b0, &duc, false, false, false, false, NULL, NULL, NULL, NULL, argv[1]);
    // Every window process inherited this RNG state at fork() and drew its root_point (est:125-128) and b_pert
    // (cell_enumeration.hpp:467-470) from it; the same draws here give the values all windows use (for replays elsewhere).
    double rp_bp[n + (n + num_windows - 1)];
    for(int i = 0; i < n; i++) rp_bp[i] = 1.0 + random_uniform();
    for(int i = 0; i < n + num_windows - 1; i++) rp_bp[n + i] = 2*random_uniform() - 1;
    for(int i = 0; i < total_steps; i++)
        swm.step(sim_log.msmt_history + i*p, NULL);
    swm.shutdown();
    char path[4096];
    sprintf(path, "%s/msmts.txt", argv[1]);          // the simulated measurement sequence, for replays through other front ends
    FILE* f = fopen(path, "w");                        // %.17g round-trips a double; the reference's loggers print %.16lf, which does not
    for(int i = 0; i < total_steps*p; i++) fprintf(f, "%.17g\n", sim_log.msmt_history[i]);
    fclose(f);
    sprintf(path, "%s/root_point_b_pert.txt", argv[1]);
    f = fopen(path, "w");
    for(int i = 0; i < n + (n + num_windows - 1); i++) fprintf(f, "%.17g\n", rp_bp[i]);
    fclose(f);
    printf("winbank drop-in done: %d measurements\n", total_steps);
    return 0;
}
