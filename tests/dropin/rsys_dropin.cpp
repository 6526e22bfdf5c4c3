// rsys_dropin.cpp -- the relative-system readers of the reference (include/cauchy_prediction.hpp: get_marg2d_relative_and_transformed_cpdf :653,
// eval_rel_sys_moments_for_term :207, grid_eval_marg2d_relative_and_transformed_cpdf :1091), reached the way the Swig module and the mex files
// reach them (pycauchy_single_step_eval_2d_rsys_cpdf, scripts/swig/cauchy/pycauchy.hpp:677): two estimators -- a primary and a secondary system --
// are stepped, then the 2-D cpdf of the transformed relative state and its moments are formed from BOTH term lists (O(Nt_p x Nt_s) term pairs).
// These readers live outside step(); behind this repository's drop-in header they walk the host mirror of the GPU-resident term lists
// (include/cauchy_estimator.hpp: sync_host_mirror).  Compiled twice like pycauchy_dropin.cpp -- unmodified reference (oracle/Makefile ->
// oracle/_ref/ex_rsys_cpu1 -> tests/golden/ex_rsys_cpu1.txt) and drop-in header + libmce_b200.so (tools/build_dropin.sh) -- and both must print
// the same text.
#include "../scripts/swig/cauchy/pycauchy.hpp"

static void step(void* h, double z)
{
    double *oPhi, *oGam, *oB, *oH, *obeta, *ogamma, *fz, *xhat, *Phat, *cfz, *cx, *cP; int* err;
    int sPhi, sGam, sB, sH, sbeta, sgamma, sfz, sx, sP, scfz, scx, scP, serr;
    pycauchy_single_step_ltiv(h, &z, 1, NULL, 0, &oPhi, &sPhi, &oGam, &sGam, &oB, &sB, &oH, &sH, &obeta, &sbeta, &ogamma, &sgamma,
                              &fz, &sfz, &xhat, &sx, &Phat, &sP, &cfz, &scfz, &cx, &scx, &cP, &scP, &err, &serr);
    printf("  z %.17g fz %.17g err %d terms %d xhat %.17g %.17g %.17g\n", z, fz[0], err[0], pycauchy_single_step_get_number_of_terms(h), xhat[0], xhat[1], xhat[2]);
    free(oPhi); free(oGam); free(oB); free(oH); free(obeta); free(ogamma); free(fz); free(xhat); free(Phat); free(cfz); free(cx); free(cP); free(err);
}

static void rsys(void* s, void* p, double* Trel, double eps)
{
    double *fz, *xh, *Ph, *cfz, *cxh, *cPh, *grid; int sfz, sxh, sPh, scfz, scxh, scPh, sgrid, nx, ny;
    pycauchy_single_step_eval_2d_rsys_cpdf(Trel, 6, s, p, eps, -0.6, 0.6, 0.2, -0.45, 0.45, 0.15,
                                           &fz, &sfz, &xh, &sxh, &Ph, &sPh, &cfz, &scfz, &cxh, &scxh, &cPh, &scPh, &grid, &sgrid, &nx, &ny);
    printf("  rsys fz %.17g (%.3e) xhat %.17g %.17g Phat %.17g %.17g %.17g %.17g grid %d x %d\n", fz[0], cfz[0], xh[0], xh[1], Ph[0], Ph[1], Ph[2], Ph[3], nx, ny);
    for(int i = 0; i < nx * ny; i++) printf("  pt %.17g %.17g %.17g\n", grid[3*i], grid[3*i+1], grid[3*i+2]);
    free(fz); free(xh); free(Ph); free(cfz); free(cxh); free(cPh); free(grid);
}

int main()
{
    // the reference's 3-state example system (src/cauchy_estimator.cpp:97-110) seen by two estimators with different start statistics and measurements
    double Phi[9] = {1.4, -0.6, -1.0, -0.2, 1.0, 0.5, 0.6, -0.6, -0.2};
    double Gamma[3] = {.1, 0.3, -0.2}, H[3] = {1.0, 0.5, 0.2}, beta[1] = {0.1}, gamma[1] = {0.2};
    double A0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p0[3] = {0.10, 0.08, 0.05}, b0[3] = {0, 0, 0};
    double A0s[9] = {0.8, -0.6, 0.0, 0.6, 0.8, 0.0, 0.0, 0.0, 1.0}, p0s[3] = {0.21, 0.07, 0.11}, b0s[3] = {0.05, -0.12, 0.3};
    double zp[5] = {-1.2172011200334241, -0.35943271347277583, -0.52353301003957098, 0.5855389648301792, -0.8048243525901404};
    double zs[5] = {0.34053610027255954, 1.0580483915838776, -0.55152999529515989, -0.72879029737003309, -0.82415138330170357};
    double Trel[6] = {1, 0, 0, 0, 1, 0};             // relative state projected onto its first two components
    double Trot[6] = {0.6, 0.8, 0, -0.8, 0.6, 0.5};  // ... and a rotated / sheared projection
    srand(11);
    void* p = pycauchy_initialize_lti(6, A0, 9, p0, 3, b0, 3, Phi, 9, Gamma, 3, NULL, 0, beta, 1, H, 3, gamma, 1, 0.0, 0, false);
    void* s = pycauchy_initialize_lti(6, A0s, 9, p0s, 3, b0s, 3, Phi, 9, Gamma, 3, NULL, 0, beta, 1, H, 3, gamma, 1, 0.0, 0, false);
    for(int k = 0; k < 4; k++)
    {
        printf("# step %d\n", k + 1);
        step(p, zp[k]); step(s, zs[k]);
        if(k >= 1) rsys(s, p, Trel, 1e-12);
    }
    printf("# rotated projection, coarser term approximation\n");
    rsys(s, p, Trot, 1e-8);
    pycauchy_single_step_shutdown(p);
    pycauchy_single_step_shutdown(s);
    printf("rsys drop-in done\n");
    return 0;
}
