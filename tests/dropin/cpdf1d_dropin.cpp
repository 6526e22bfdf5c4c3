// cpdf1d_dropin.cpp -- TEST PROGRAM. Runs the 3-state example system of the reference's src/cauchy_estimator.cpp
// (:97-110) on the GPU estimator (this repository's drop-in cauchy_estimator.hpp) and, after every step, evaluates the
// 1-D marginal cpdf of every state twice on the same grid:
//   * with the reference's OWN CPU code (CauchyCPDFGridDispatcher1D over the host mirror of the device term list,
//     cpdf_ndim.hpp compiled unchanged from the overlay), and
//   * with the device dispatcher (include/cpdf_b200.hpp -> mce_marginal_1d_grid),
// and compares the two grids bit for bit.  Built by tools/build_dropin.sh; prints "cpdf1d drop-in OK".
#include "cauchy_estimator.hpp"   // the overlay resolves this to the B200 drop-in
#include "cpdf_ndim.hpp"
#include "cpdf_b200.hpp"

int main()
{
    const int n = 3, cmcc = 0, pncc = 1, p = 1, steps = 8;
    double Phi[n*n] = {1.4, -0.6, -1.0,  -0.2, 1.0, 0.5,  0.6, -0.6, -0.2};
    double Gamma[n*pncc] = {.1, 0.3, -0.2};
    double H[n] = {1.0, 0.5, 0.2};
    double beta[pncc] = {0.1};
    double gamma[p] = {0.2};
    double A0[n*n] = {1,0,0, 0,1,0, 0,0,1};
    double p0[n] = {0.10, 0.08, 0.05};
    double b0[n] = {0, 0, 0};
    double zs[steps] = {0.022172011200334241, -0.11943271347277583, -1.22353301003957098, -1.4055389648301792,
                        -1.34053610027255954, 0.4580483915838776, 0.65152999529515989, 0.52378648722334};
    CauchyEstimator est(A0, p0, b0, steps, n, cmcc, pncc, p, false);
    PointWiseNDimCauchyCPDF cpdf(&est);
    CauchyCPDFGridDispatcher1D ref_grid(&cpdf, -2.0, 2.0, 0.01, NULL);
    CauchyCPDFGridDispatcher1D_B200 dev_grid(&cpdf, -2.0, 2.0, 0.01, NULL);
    CauchyCPDFGridDispatcher2D ref_grid2(&cpdf, -1.0, 1.0, 0.1, -1.5, 1.5, 0.15, NULL);
    CauchyCPDFGridDispatcher2D_B200 dev_grid2(&cpdf, -1.0, 1.0, 0.1, -1.5, 1.5, 0.15, NULL);
    long long compared = 0, differing = 0, compared2 = 0; double worst2 = 0;
    for(int k = 0; k < steps - 1; k++)      // the window's last step has no tables (SKIP_LAST_STEP)
    {
        est.step(zs[k], Phi, Gamma, beta, H, gamma[0], NULL, NULL);
        est.sync_host_mirror();             // host copy of the term list for the reference's CPU reader
        for(int idx = 0; idx < n; idx++)
        {
            cpdf.master_step_of_cached_1d_terms = -1;        // force the reference to rebuild its cache
            if(ref_grid.evaluate_point_grid(idx, 1, false)) { printf("reference grid refused at step %d\n", k+1); return 1; }
            if(dev_grid.evaluate_point_grid(idx, 1, false)) { printf("device grid refused at step %d\n", k+1); return 1; }
            if(ref_grid.num_grid_points != dev_grid.num_grid_points) { printf("grid size mismatch\n"); return 1; }
            for(int i = 0; i < ref_grid.num_grid_points; i++)
            {
                compared++;
                if(memcmp(&ref_grid.points[i], &dev_grid.points[i], sizeof(CauchyPoint2D)) != 0)
                {
                    if(differing++ < 5)
                        printf("step %d idx %d point %d: reference (%.17g, %.17g) device (%.17g, %.17g)\n", k+1, idx, i,
                               ref_grid.points[i].x, ref_grid.points[i].y, dev_grid.points[i].x, dev_grid.points[i].y);
                }
            }
        }
        // 2-D marginals: same term order, but the device's atan2 / sin / cos differ from glibc's in the last bits
        for(int pair = 0; pair < 2 && k < 6; pair++)
        {
            int i1 = 0, i2 = 1 + pair;
            cpdf.reset_2D_marginal_cpdf();
            if(ref_grid2.evaluate_point_grid(i1, i2, 1, false) || dev_grid2.evaluate_point_grid(i1, i2, 1, false)) { printf("2-D grid refused at step %d\n", k+1); return 1; }
            double zmax = 0;
            for(int i = 0; i < ref_grid2.num_grid_points; i++) zmax = fmax(zmax, fabs(ref_grid2.points[i].z));
            for(int i = 0; i < ref_grid2.num_grid_points; i++)
            {
                compared2++;
                if(ref_grid2.points[i].x != dev_grid2.points[i].x || ref_grid2.points[i].y != dev_grid2.points[i].y) { printf("2-D grid coordinates differ\n"); return 1; }
                worst2 = fmax(worst2, fabs(ref_grid2.points[i].z - dev_grid2.points[i].z) / zmax);
            }
        }
        printf("step %d: %d terms, f(0) = %.12e %.12e %.12e\n", k+1, est.Nt, dev_grid.points[200].y, ref_grid.points[200].y, ref_grid.points[0].y);
    }
    printf("compared %lld grid values, %lld differ\n", compared, differing);
    printf("compared %lld 2-D grid values, worst |dz| / max|z| = %.3e\n", compared2, worst2);
    if(worst2 > 1e-9) { printf("2-D marginal outside tolerance\n"); return 1; }
    if(differing == 0 && compared > 0)
        printf("cpdf1d drop-in OK\n");
    return differing == 0 ? 0 : 1;
}
