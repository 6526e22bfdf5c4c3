// pycauchy_dropin.cpp -- drives the reference's Swig shim (scripts/swig/cauchy/pycauchy.hpp, plain C++: the same functions
// pycauchy.i wraps for Python and the MATLAB mex files call) from a C++ main, so that the shim can be checked without swig.
// Compiled twice from an overlay of the reference tree: once with the reference's own cauchy_estimator.hpp (NUM_CPUS = 1;
// oracle/Makefile -> oracle/_ref/ex_pycauchy_cpu1, whose output is the golden text tests/golden/ex_pycauchy_cpu1.txt) and
// once with this repository's drop-in header + libmce_b200.so (tools/build_dropin.sh -> build/dropin/pycauchy_dropin).
// Both must print the same text.  Exercises the paths a PySlidingWindowManager uses to re-initialise a window
// (pycauchy.hpp:798-819): reset() re-seeding from A0_init/p0_init/b0_init written in place, the master_step == 0 branch
// that writes the first term through setup_first_term, and deterministic transforms before and after the first step.
#include "../scripts/swig/cauchy/pycauchy.hpp"

static void one_step(void* h, double z, int n)
{
    double *oPhi, *oGam, *oB, *oH, *obeta, *ogamma, *fz, *xhat, *Phat, *cfz, *cx, *cP; int* err;
    int sPhi, sGam, sB, sH, sbeta, sgamma, sfz, sx, sP, scfz, scx, scP, serr;
    pycauchy_single_step_ltiv(h, &z, 1, NULL, 0, &oPhi, &sPhi, &oGam, &sGam, &oB, &sB, &oH, &sH, &obeta, &sbeta, &ogamma, &sgamma,
                              &fz, &sfz, &xhat, &sx, &Phat, &sP, &cfz, &scfz, &cx, &scx, &cP, &scP, &err, &serr);
    printf("z %.17g fz %.17g cerr_fz %.17g err %d terms %d\n  xhat", z, fz[0], cfz[0], err[0], pycauchy_single_step_get_number_of_terms(h));
    for(int i = 0; i < n; i++) printf(" %.17g", xhat[i]);
    printf("\n  Phat");
    for(int i = 0; i < n*n; i++) printf(" %.17g", Phat[i]);
    printf("\n  cerr %.17g %.17g\n", cx[0], cP[0]);
    free(oPhi); free(oGam); free(oB); free(oH); free(obeta); free(ogamma); free(fz); free(xhat); free(Phat); free(cfz); free(cx); free(cP); free(err);
}

int main()
{
    const int n = 3;
    // the reference's 3-state example system, src/cauchy_estimator.cpp:97-110
    double Phi[9] = {1.4, -0.6, -1.0, -0.2, 1.0, 0.5, 0.6, -0.6, -0.2};
    double Gamma[3] = {.1, 0.3, -0.2}, H[3] = {1.0, 0.5, 0.2}, beta[1] = {0.1}, gamma[1] = {0.2};
    double A0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p0[3] = {0.10, 0.08, 0.05}, b0[3] = {0, 0, 0};
    double zs[10] = {-1.2172011200334241, -0.35943271347277583, -0.52353301003957098, 0.5855389648301792, -0.8048243525901404,
                     0.34053610027255954, 1.0580483915838776, -0.55152999529515989, -0.72879029737003309, -0.82415138330170357};
    // re-initialisation statistics of the kind speyers_window_init produces (a rotated basis, new weights, a shifted centre)
    double A0b[9] = {0.8, -0.6, 0.0, 0.6, 0.8, 0.0, 0.0, 0.0, 1.0}, p0b[3] = {0.21, 0.07, 0.11}, b0b[3] = {0.05, -0.12, 0.3};
    double A0c[9] = {1.0, 0.2, -0.1, 0.0, 1.0, 0.3, 0.0, 0.0, 1.0}, p0c[3] = {0.3, 0.15, 0.09}, b0c[3] = {-0.2, 0.1, 0.05};
    double Trans[9] = {1.0, 0.1, 0.005, 0.0, 1.0, 0.1, 0.0, 0.0, 0.9}, bias[3] = {0.01, -0.02, 0.03};
    srand(7);
    void* h = pycauchy_initialize_lti(6, A0, 9, p0, 3, b0, 3, Phi, 9, Gamma, 3, NULL, 0, beta, 1, H, 3, gamma, 1, 0.0, 0, false);
    printf("# a: 4 steps from the constructor's statistics\n");
    for(int k = 0; k < 4; k++) one_step(h, zs[k], n);
    printf("# b: reset mid-window with new A0, p0, b0 (master_step != 0 -> reset())\n");
    pycauchy_single_step_reset(h, A0b, 9, p0b, 3, b0b, 3, NULL, 0);
    for(int k = 0; k < 6; k++) one_step(h, zs[k + 2], n);          // runs to the end of the window
    printf("# c: reset after the last step keeping the statistics, then the master_step == 0 branch with new ones\n");
    pycauchy_single_step_reset(h, NULL, 0, NULL, 0, NULL, 0, NULL, 0);
    pycauchy_single_step_reset(h, A0c, 9, p0c, 3, b0c, 3, NULL, 0);
    for(int k = 0; k < 3; k++) one_step(h, zs[k + 1], n);
    printf("# d: reset() must come back to A0b (the last statistics written into A0_init)... which the step-0 branch did not touch\n");
    pycauchy_single_step_reset(h, NULL, 0, NULL, 0, NULL, 0, NULL, 0);
    for(int k = 0; k < 2; k++) one_step(h, zs[k + 4], n);
    printf("# e: deterministic transform of the initial term, two steps, a transform mid-window, one more step\n");
    pycauchy_single_step_reset(h, A0, 9, p0, 3, b0, 3, NULL, 0);
    pycauchy_single_step_deterministic_transform(h, Trans, 9, bias, 3);
    for(int k = 0; k < 2; k++) one_step(h, zs[k], n);
    pycauchy_single_step_deterministic_transform(h, Trans, 9, bias, 3);
    one_step(h, zs[2], n);
    pycauchy_single_step_shutdown(h);
    printf("pycauchy drop-in done\n");
    return 0;
}
