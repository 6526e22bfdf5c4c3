// cpdf_b200.hpp -- device-side replacements for the reference's 1-D and 2-D marginal cpdf grid dispatchers.
//
// `CauchyCPDFGridDispatcher1D_B200` has the members callers of `CauchyCPDFGridDispatcher1D` use
// (/root/reference/include/cpdf_ndim.hpp:2017-2230; pycauchy.hpp:890-931): the constructor arguments, `points`,
// `num_grid_points`, `reset_grid`, `evaluate_point_grid(marg_idx, num_threads, with_timing)` and `log_point_grid()`.
// The grid is evaluated on the GPU from the resident term list through the C ABI (mce_marginal_1d_grid); no term leaves
// the device and `num_threads` is ignored.  Values, grid and log files are identical to the reference's.
// Include it after cauchy_estimator.hpp (this repository's drop-in) and cpdf_ndim.hpp; it needs the reference's
// `PointWiseNDimCauchyCPDF` only for `bar_nu` (the random direction drawn at cpdf_ndim.hpp:385-389) and `cauchyEst`.
#ifndef _CPDF_B200_HPP_
#define _CPDF_B200_HPP_

#include "cauchy_estimator.hpp"
#include "cpdf_ndim.hpp"
#include "mce_b200.h"

struct CauchyCPDFGridDispatcher1D_B200
{
    PointWiseNDimCauchyCPDF* cpdf;
    CauchyPoint2D* points;
    int num_grid_points;
    char* log_dir;
    int tags[10];
    int tag_counts[10];
    int num_tags;
    int last_marg_idx;
    double grid_low, grid_high, grid_res;

    CauchyCPDFGridDispatcher1D_B200(PointWiseNDimCauchyCPDF* _cpdf, double _grid_low, double _grid_high, double _grid_res, char* _log_dir = NULL)
    {
        points = NULL; cpdf = _cpdf; num_tags = 0; last_marg_idx = -1;
        reset_grid(_grid_low, _grid_high, _grid_res);
        if(_log_dir != NULL)
        {
            int len = strlen(_log_dir);
            log_dir = (char*) malloc((len + 1) * sizeof(char));
            strcpy(log_dir, _log_dir);
            if(log_dir[len-1] == '/')
                log_dir[len-1] = '\0';
            check_dir_and_create(log_dir);
        }
        else
            log_dir = NULL;
    }

    void reset_grid(double _grid_low, double _grid_high, double _grid_res)      // cpdf_ndim.hpp:2055-2072
    {
        assert(_grid_high > _grid_low);
        assert(_grid_res > 0);
        grid_low = _grid_low; grid_high = _grid_high; grid_res = _grid_res;
        num_grid_points = mce_cpdf_grid_count(grid_low, grid_high, grid_res);
        points = (CauchyPoint2D*) realloc(points, num_grid_points * sizeof(CauchyPoint2D));
        null_ptr_check(points);
        for(int i = 0; i < num_grid_points; i++)
        {
            double grid_point = grid_low + i * grid_res;
            if(grid_point > grid_high)
                grid_point = grid_high;
            points[i].x = grid_point;
            points[i].y = -1;
        }
    }

    int evaluate_point_grid(int marg_idx, int num_threads, bool with_timing = false)   // cpdf_ndim.hpp:2074-2139
    {
        CauchyEstimator* est = cpdf->cauchyEst;
        assert(marg_idx < est->d);
        assert(marg_idx > -1);
        (void) num_threads;
        if( (est->master_step == est->num_estimation_steps) && (SKIP_LAST_STEP == true) )
        {
            printf(YEL "[WARN CauchyCPDFGridDispatcher1D:] Cannot evaluate cauchy estimator cpdf for the last step since SKIP_LAST_STEP == true! (The G Tables were not created, as they were skipped!)" NC "\n");
            return 1;
        }
        int rc = mce_marginal_1d_grid(est->handle, marg_idx, cpdf->bar_nu, grid_low, grid_high, grid_res, (double*) points, num_grid_points);
        if(rc < 0)
        {
            printf(RED "[CauchyCPDFGridDispatcher1D/B200] %s" NC "\n", mce_last_error());
            exit(1);
        }
        if(rc == 0)
            return 1;
        last_marg_idx = marg_idx;
        if(with_timing)
            printf("1D Grid Eval Step %d:\n  Computing %d gridpoints of %d CF terms took: %.3lf ms on the device\n", est->master_step, num_grid_points, est->Nt, mce_cpdf_last_ms(est->handle));
        return 0;
    }

    // {log_dir}/cpdf_{marg_idx}_{count}.bin, {log_dir}/grid_elems_{marg_idx}.txt -- cpdf_ndim.hpp:2141-2202
    int log_point_grid()
    {
        if(log_dir == NULL)
        {
            printf(YEL "[WARN CauchyCPDFGridDispatcher1D:]\n  Cannot Log! The log directory was not set!" NC "\n");
            return 1;
        }
        CauchyEstimator* est = cpdf->cauchyEst;
        if( ((est->master_step == est->num_estimation_steps) && (SKIP_LAST_STEP == true)) || last_marg_idx < 0 )
            return 1;
        int tag = last_marg_idx, tag_idx = -1;
        for(int i = 0; i < num_tags; i++)
            if(tags[i] == tag)
                tag_idx = i;
        if(tag_idx == -1)
        {
            tag_idx = num_tags; tags[tag_idx] = tag; tag_counts[tag_idx] = 0; num_tags++;
        }
        int tag_count = ++tag_counts[tag_idx];
        char* path = (char*) malloc((strlen(log_dir) + 64) * sizeof(char));
        sprintf(path, "%s/grid_elems_%d.txt", log_dir, tag);
        FILE* dims_file = fopen(path, tag_count == 1 ? "w" : "a");
        if(dims_file == NULL) { printf(RED "[ERROR CauchyCPDFGridDispatcher1D:] Could not open %s" NC "\n", path); exit(1); }
        fprintf(dims_file, "%d\n", num_grid_points);
        sprintf(path, "%s/cpdf_%d_%d.bin", log_dir, tag, tag_count);
        FILE* data_file = fopen(path, "wb");
        if(data_file == NULL) { printf(RED "[ERROR CauchyCPDFGridDispatcher1D:] Could not open %s" NC "\n", path); exit(1); }
        fwrite(points, sizeof(CauchyPoint2D), num_grid_points, data_file);
        fclose(data_file); fclose(dims_file); free(path);
        return 0;
    }

    ~CauchyCPDFGridDispatcher1D_B200()
    {
        free(points);
        if(log_dir != NULL) free(log_dir);
    }
};

// Device-side replacement for CauchyCPDFGridDispatcher2D (cpdf_ndim.hpp:1774-2007; pycauchy.hpp:822-872): same members
// (`points` of CauchyPoint3D, y-major; `num_points_x`, `num_points_y`, `num_grid_points`; reset_grid; evaluate_point_grid;
// log_point_grid).  Sums run in the reference's term order; atan2 / sin / cos are the device's.
struct CauchyCPDFGridDispatcher2D_B200
{
    PointWiseNDimCauchyCPDF* cpdf;
    CauchyPoint3D* points;
    int num_grid_points;
    int num_points_x;
    int num_points_y;
    char* log_dir;
    int tags[25*2];
    int tag_counts[25];
    int num_tags;
    int last_idxs[2];
    double glx, ghx, grx, gly, ghy, gry;

    CauchyCPDFGridDispatcher2D_B200(PointWiseNDimCauchyCPDF* _cpdf, double grid_low_x, double grid_high_x, double grid_res_x,
        double grid_low_y, double grid_high_y, double grid_res_y, char* _log_dir = NULL)
    {
        points = NULL; cpdf = _cpdf; num_tags = 0; last_idxs[0] = last_idxs[1] = -1;
        reset_grid(grid_low_x, grid_high_x, grid_res_x, grid_low_y, grid_high_y, grid_res_y);
        if(_log_dir != NULL)
        {
            int len = strlen(_log_dir);
            log_dir = (char*) malloc((len + 1) * sizeof(char));
            strcpy(log_dir, _log_dir);
            if(log_dir[len-1] == '/')
                log_dir[len-1] = '\0';
            check_dir_and_create(log_dir);
        }
        else
            log_dir = NULL;
    }

    void reset_grid(double grid_low_x, double grid_high_x, double grid_res_x, double grid_low_y, double grid_high_y, double grid_res_y)   // cpdf_ndim.hpp:1816-1848
    {
        assert(grid_high_x > grid_low_x);
        assert(grid_high_y > grid_low_y);
        assert(grid_res_x > 0);
        assert(grid_res_y > 0);
        glx = grid_low_x; ghx = grid_high_x; grx = grid_res_x; gly = grid_low_y; ghy = grid_high_y; gry = grid_res_y;
        num_points_x = mce_cpdf_grid_count(glx, ghx, grx);
        num_points_y = mce_cpdf_grid_count(gly, ghy, gry);
        num_grid_points = num_points_x * num_points_y;
        points = (CauchyPoint3D*) realloc(points, num_grid_points * sizeof(CauchyPoint3D));
        null_ptr_check(points);
        for(int i = 0; i < num_points_y; i++)
        {
            double gy = gly + i * gry; if(gy > ghy) gy = ghy;
            for(int j = 0; j < num_points_x; j++)
            {
                double gx = glx + j * grx; if(gx > ghx) gx = ghx;
                points[i*num_points_x + j].x = gx; points[i*num_points_x + j].y = gy; points[i*num_points_x + j].z = -1;
            }
        }
    }

    int evaluate_point_grid(int marg_idx1, int marg_idx2, int num_threads, bool with_timing = false)   // cpdf_ndim.hpp:1850-1919
    {
        CauchyEstimator* est = cpdf->cauchyEst;
        assert(marg_idx1 < marg_idx2);
        assert(marg_idx2 < est->d);
        assert(marg_idx1 > -1);
        (void) num_threads;
        if( (est->master_step == est->num_estimation_steps) && (SKIP_LAST_STEP == true) )
        {
            printf(YEL "[WARN CauchyCPDFGridDispatcher2D:] Cannot evaluate cauchy estimator cpdf for the last step since SKIP_LAST_STEP == true! (The G Tables were not created, as they were skipped!)" NC "\n");
            return 1;
        }
        int rc = mce_marginal_2d_grid(est->handle, marg_idx1, marg_idx2, cpdf->bar_nu, glx, ghx, grx, gly, ghy, gry, (double*) points, num_grid_points, NULL, NULL);
        if(rc < 0)
        {
            printf(RED "[CauchyCPDFGridDispatcher2D/B200] %s" NC "\n", mce_last_error());
            exit(1);
        }
        if(rc == 0)
            return 1;
        last_idxs[0] = marg_idx1; last_idxs[1] = marg_idx2;
        if(with_timing)
            printf("2D Grid Eval Step %d:\n  Computing %d gridpoints of %d CF terms took: %.3lf ms on the device\n", est->master_step, num_grid_points, est->Nt, mce_cpdf_last_ms(est->handle));
        return 0;
    }

    // {log_dir}/cpdf_{idx1}{idx2}_{count}.bin, {log_dir}/grid_elems_{idx1}{idx2}.txt -- cpdf_ndim.hpp:1921-1983
    int log_point_grid()
    {
        if(log_dir == NULL)
        {
            printf(YEL "[WARN CauchyCPDFGridDispatcher2D:]\n  Cannot Log! The log directory was not set!" NC "\n");
            return 1;
        }
        CauchyEstimator* est = cpdf->cauchyEst;
        if( ((est->master_step == est->num_estimation_steps) && (SKIP_LAST_STEP == true)) || last_idxs[0] < 0 )
            return 1;
        int tag_idx = -1;
        for(int i = 0; i < num_tags; i++)
            if(tags[2*i] == last_idxs[0] && tags[2*i+1] == last_idxs[1])
                tag_idx = i;
        if(tag_idx == -1)
        {
            tag_idx = num_tags; tags[2*tag_idx] = last_idxs[0]; tags[2*tag_idx+1] = last_idxs[1]; tag_counts[tag_idx] = 0; num_tags++;
        }
        int tag_count = ++tag_counts[tag_idx];
        char* path = (char*) malloc((strlen(log_dir) + 64) * sizeof(char));
        sprintf(path, "%s/grid_elems_%d%d.txt", log_dir, last_idxs[0], last_idxs[1]);
        FILE* dims_file = fopen(path, tag_count == 1 ? "w" : "a");
        if(dims_file == NULL) { printf(RED "[ERROR CauchyCPDFGridDispatcher2D:] Could not open %s" NC "\n", path); exit(1); }
        fprintf(dims_file, "%d,%d\n", num_points_x, num_points_y);
        sprintf(path, "%s/cpdf_%d%d_%d.bin", log_dir, last_idxs[0], last_idxs[1], tag_count);
        FILE* data_file = fopen(path, "wb");
        if(data_file == NULL) { printf(RED "[ERROR CauchyCPDFGridDispatcher2D:] Could not open %s" NC "\n", path); exit(1); }
        fwrite(points, sizeof(CauchyPoint3D), num_grid_points, data_file);
        fclose(data_file); fclose(dims_file); free(path);
        return 0;
    }

    ~CauchyCPDFGridDispatcher2D_B200()
    {
        free(points);
        if(log_dir != NULL) free(log_dir);
    }
};

#endif // _CPDF_B200_HPP_
