"""Host-side mirror of the reference's `struct CauchyEstimator` (include/cauchy_estimator.hpp:47-1432) over the
C ABI of libmce_b200.so.  Same member names, argument meaning and error behaviour as the reference, so code written
against the reference (src/*.cpp, scripts/swig/cauchy/cauchy_estimator.py:515 PyCauchyEstimator) ports line by line.
All term-list work happens on the GPU; this file only marshals arguments."""
import ctypes as ct
import os

import numpy as np

from . import _capi

# numeric_moment_errors bits, cauchy_constants.hpp:104-116
ERROR_FZ_NEGATIVE = 9
ERROR_FZ_UNSTABLE = 8
ERROR_MEAN_AT_CURRENT_STEP_DNE = 7
ERROR_COVARIANCE_AT_CURRENT_STEP_DNE = 3


def _dp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_double)) if a is not None else None


def _f64(x, n=None):
    if x is None:
        return None
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError("expected %d values, got %d" % (n, a.size))
    return a


class CauchyEstimator:
    """CauchyEstimator(A0, p0, b0, steps, d, cmcc, pncc, p, print_basic_info) -- est:87.

    root_point / b_pert are the vectors the reference draws with libc rand() in its constructor (est:125-128,
    cell_enumeration.hpp:467-470); pass recorded values to reproduce a reference run, or leave None to draw them
    from `seed` with the same distributions (1 + U(0,1], 2U - 1)."""

    def __init__(self, A0, p0, b0, steps, d, cmcc, pncc, p, print_basic_info=False, root_point=None, b_pert=None,
                 tr_search_idxs_ordering=None, device=-1, seed=0, fast_moments=False, _lib=None):
        self._lib = _lib if _lib is not None else _capi.load()   # _lib: test hook (tests/emu), never set by product code
        self.d, self.cmcc, self.pncc, self.p = int(d), int(cmcc), int(pncc), int(p)
        self.num_estimation_steps = self.p * int(steps)
        max_shape = (int(steps) - 1) * self.pncc + self.d if self.d > 1 else self.d + self.pncc
        rng = np.random.RandomState(seed)
        if root_point is None:
            root_point = 1.0 + (rng.randint(0, 2**31 - 1, self.d) + 1.0) / 2.0**31
        if b_pert is None:
            b_pert = 2.0 * (rng.randint(0, 2**31 - 1, max_shape) + 1.0) / 2.0**31 - 1.0
        self.root_point = _f64(root_point, self.d)
        bp = np.zeros(max(max_shape, 1) + _capi.MAXM)
        bp[: len(b_pert)] = np.asarray(b_pert, np.float64)[: len(bp)]
        self.b_pert = bp
        o = _capi.MceOptions()
        self._lib.mce_default_options(ct.byref(o))
        o.device = int(device)
        o.print_basic_info = int(bool(print_basic_info))
        o.fast_moments = int(bool(fast_moments))
        if tr_search_idxs_ordering is not None:
            for i, v in enumerate(list(tr_search_idxs_ordering)[:12]):
                o.tr_search_order[i] = int(v)
        self._A0, self._p0, self._b0 = _f64(A0, self.d * self.d), _f64(p0, self.d), _f64(b0, self.d)
        self._h = self._lib.mce_create(self.d, self.cmcc, self.pncc, self.p, int(steps), _dp(self._A0), _dp(self._p0), _dp(self._b0),
                                       _dp(self.root_point), _dp(self.b_pert), ct.byref(o))
        if not self._h:
            raise RuntimeError(self._lib.mce_last_error().decode())
        self.shape_range = self._lib.mce_shape_range(self._h)
        self.print_basic_info = bool(print_basic_info)
        self.win_num = 0
        self._refresh()

    # ---- public fields of the reference struct ----
    def _refresh(self):
        m = _capi.MceMoments()
        self._lib.mce_get_moments(self._h, ct.byref(m))
        d = self.d
        self.fz = complex(m.fz[0], m.fz[1])
        self.fz_after_mu = complex(m.fz_after_mu[0], m.fz_after_mu[1])
        self.conditional_mean = np.array(m.mean[: 2 * d]).view(np.complex128).copy()
        self.conditional_variance = np.array(m.cov[: 2 * d * d]).view(np.complex128).reshape(d, d).copy()
        self.G_SCALE_FACTOR = m.g_scale_factor
        self.numeric_moment_errors = m.numeric_moment_errors
        self.Nt = m.Nt
        self.Nt_after_muc = m.Nt_after_muc
        self._master_step = m.master_step
        self.skip_post_mu = bool(m.skip_post_mu)

    @property
    def master_step(self):
        return self._master_step

    @master_step.setter
    def master_step(self, v):          # callers write this field (cauchy_windows.hpp:538,659; pycauchy.hpp:776)
        self._lib.mce_set_master_step(self._h, int(v))
        self._master_step = int(v)

    @property
    def terms_per_shape(self):
        c = (ct.c_int * self.shape_range)()
        self._lib.mce_get_terms_per_shape(self._h, c, 0)
        return np.array(c[:], np.int32)

    def set_win_num(self, win_num):
        self.win_num = int(win_num)

    def set_function_pointers(self):   # est:192: storage mode is fixed (BINSEARCH + HALF storage) in this build
        pass

    # ---- the hot path ----
    def step(self, msmt, Phi, Gamma, beta, H, gamma, B=None, u=None):
        """int step(msmt, Phi, Gamma, beta, H, gamma, B, u) -- est:1211. Returns numeric_moment_errors."""
        d = self.d
        Phi_, Gam_, beta_, H_ = _f64(Phi), _f64(Gamma), _f64(beta), _f64(H, d)
        B_, u_ = _f64(B), _f64(u)
        rc = self._lib.mce_step(self._h, float(msmt), _dp(Phi_), _dp(Gam_), _dp(beta_), _dp(H_), float(gamma), _dp(B_), _dp(u_))
        if rc < 0:
            raise RuntimeError(self._lib.mce_last_error().decode())
        self._refresh()
        return rc

    def finalize_extended_moments(self, x_bar):
        """est:1358-1394: b <- b - Re(mean) on every term; mean += x_bar; x_bar <- Re(mean) (in place)."""
        delta = np.ascontiguousarray(self.conditional_mean.real)
        self._lib.mce_shift_b(self._h, _dp(delta), -1.0)
        self.conditional_mean = self.conditional_mean + np.asarray(x_bar, np.float64)
        x_bar[:] = self.conditional_mean.real
        return x_bar

    def shift_cf_by_bias(self, bias):
        self._lib.mce_shift_b(self._h, _dp(_f64(bias, self.d)), 1.0)

    def deterministic_time_prop(self, Phi, B=None, u=None):
        rc = self._lib.mce_deterministic_time_prop(self._h, _dp(_f64(Phi, self.d * self.d)), _dp(_f64(B)), _dp(_f64(u)))
        if rc < 0:
            raise RuntimeError(self._lib.mce_last_error().decode())

    def reset(self):
        self._lib.mce_reset(self._h)
        self._refresh()

    def reinitialize_start_statistics(self, A0, p0, b0):
        self._A0, self._p0, self._b0 = _f64(A0, self.d * self.d), _f64(p0, self.d), _f64(b0, self.d)
        self._lib.mce_reinitialize_start_statistics(self._h, _dp(self._A0), _dp(self._p0), _dp(self._b0))

    def step_stats(self):
        s = _capi.MceStepStats()
        self._lib.mce_get_step_stats(self._h, ct.byref(s))
        return {n: getattr(s, n) for n, _ in s._fields_}

    def export_shape(self, m):
        """Host copy of the terms with m hyperplanes: A [n, m*d], p [n, m], b [n, d], cells [n], keys, G (concatenated)."""
        d = self.d
        n, tot = ct.c_int(0), ct.c_longlong(0)
        self._lib.mce_export_shape(self._h, m, ct.byref(n), ct.byref(tot), None, None, None, None, None, None)
        n, tot = n.value, tot.value
        A, p, b = np.zeros((n, m * d)), np.zeros((n, m)), np.zeros((n, d))
        cells, keys, G = np.zeros(n, np.int32), np.zeros(tot, np.uint32), np.zeros(tot, np.complex128)
        if n:
            n2, t2 = ct.c_int(0), ct.c_longlong(0)
            self._lib.mce_export_shape(self._h, m, ct.byref(n2), ct.byref(t2), _dp(A), _dp(p), _dp(b),
                                       cells.ctypes.data_as(ct.POINTER(ct.c_int)), keys.ctypes.data_as(ct.POINTER(ct.c_uint32)),
                                       G.view(np.float64).ctypes.data_as(ct.POINTER(ct.c_double)))
        return dict(A=A, p=p, b=b, cells=cells, keys=keys, G=G)

    # ---- point-wise marginal cpdf (cauchy_estimator.py:1003-1031 / pycauchy.hpp:890-931), evaluated on the device ----
    def _bar_nu(self):
        # the reference draws this direction with rand() when the cpdf object is created (cpdf_ndim.hpp:385-389: 2 U(0,1])
        if getattr(self, "bar_nu", None) is None:
            self.bar_nu = 2.0 * (np.random.RandomState(12345).randint(0, 2**31 - 1, self.d) + 1.0) / 2.0**31
        return _f64(self.bar_nu, self.d)

    def get_marginal_1D_pointwise_cpdf(self, marg_idx, gridx_low, gridx_high, gridx_resolution, log_dir=None):
        """X, Y of the marginal cpdf of state `marg_idx` on the grid [low, high] with step `resolution` (same grid, same
        values and -- when `log_dir` is given -- the same files as the reference: cpdf_ndim.hpp:2055-2072, 2141-2202)."""
        if self.master_step < 1:
            print("Cannot evaluate Cauchy Estimator 1D Marginal CPDF before it has been stepped!")
            return None, None
        lo, hi, res, idx = float(gridx_low), float(gridx_high), float(gridx_resolution), int(marg_idx)
        assert hi > lo and res > 0 and -1 < idx < self.d
        n = self._lib.mce_cpdf_grid_count(lo, hi, res)
        xy = np.zeros((n, 2))
        bar_nu = self._bar_nu()
        rc = self._lib.mce_marginal_1d_grid(self._h, idx, _dp(bar_nu), lo, hi, res, _dp(xy), n)
        if rc < 0:
            raise RuntimeError(self._lib.mce_last_error().decode())
        if rc == 0:
            print("[WARN CauchyCPDFGridDispatcher1D:] Cannot evaluate cauchy estimator cpdf for the last step since SKIP_LAST_STEP == true!")
            return np.zeros(0), np.zeros(0)
        if log_dir:
            self._log_1d_grid(str(log_dir).rstrip("/"), idx, xy)
        return xy[:, 0].copy(), xy[:, 1].copy()

    def get_1D_pointwise_cpdf(self, gridx_low, gridx_high, gridx_resolution, log_dir=None):   # cauchy_estimator.py:1033
        if self.d != 1:
            print("Cannot evaluate Cauchy Estimator 1D CPDF for a {}-state system!".format(self.d))
            return None, None
        return self.get_marginal_1D_pointwise_cpdf(0, gridx_low, gridx_high, gridx_resolution, log_dir)

    def get_marginal_2D_pointwise_cpdf(self, marg_idx1, marg_idx2, gridx_low, gridx_high, gridx_resolution, gridy_low, gridy_high,
                                       gridy_resolution, log_dir=None, reset_cache=True):
        """X, Y, Z (each [num_gridy, num_gridx]) of the marginal cpdf of the state pair, like cauchy_estimator.py:954-989;
        with `log_dir` the reference's files are written (cpdf_ndim.hpp:1921-1983).  `reset_cache` is accepted and ignored:
        the device evaluation keeps no cache between calls."""
        if self.master_step < 1:
            print("Cannot evaluate Cauchy Estimator Marginal 2D CPDF before it has been stepped!")
            return None, None, None
        i1, i2 = int(marg_idx1), int(marg_idx2)
        xl, xh, xr, yl, yh, yr = (float(v) for v in (gridx_low, gridx_high, gridx_resolution, gridy_low, gridy_high, gridy_resolution))
        assert xh > xl and yh > yl and xr > 0 and yr > 0
        assert -1 < i1 < i2 < self.d
        nx, ny = self._lib.mce_cpdf_grid_count(xl, xh, xr), self._lib.mce_cpdf_grid_count(yl, yh, yr)
        xyz = np.zeros((nx * ny, 3))
        bar_nu = self._bar_nu()
        rc = self._lib.mce_marginal_2d_grid(self._h, i1, i2, _dp(bar_nu), xl, xh, xr, yl, yh, yr, _dp(xyz), nx * ny, None, None)
        if rc < 0:
            raise RuntimeError(self._lib.mce_last_error().decode())
        if rc == 0:
            print("[WARN CauchyCPDFGridDispatcher2D:] Cannot evaluate cauchy estimator cpdf for the last step since SKIP_LAST_STEP == true!")
            return np.zeros((0, 0)), np.zeros((0, 0)), np.zeros((0, 0))
        if log_dir:
            ld = str(log_dir).rstrip("/")
            os.makedirs(ld, exist_ok=True)
            counts = self.__dict__.setdefault("_cpdf_log_counts", {})
            k = counts.get((ld, i1, i2), 0)
            with open(os.path.join(ld, "grid_elems_%d%d.txt" % (i1, i2)), "w" if k == 0 else "a") as f:
                f.write("%d,%d\n" % (nx, ny))
            xyz.tofile(os.path.join(ld, "cpdf_%d%d_%d.bin" % (i1, i2, k + 1)))
            counts[(ld, i1, i2)] = k + 1
        g = xyz.reshape(ny, nx, 3)
        return g[:, :, 0].copy(), g[:, :, 1].copy(), g[:, :, 2].copy()

    def get_2D_pointwise_cpdf(self, gridx_low, gridx_high, gridx_resolution, gridy_low, gridy_high, gridy_resolution, log_dir=None):
        if self.d != 2:                                            # cauchy_estimator.py:991-1001
            print("Cannot evaluate Cauchy Estimator 2D CPDF for a {}-state system!".format(self.d))
            return None, None, None
        return self.get_marginal_2D_pointwise_cpdf(0, 1, gridx_low, gridx_high, gridx_resolution, gridy_low, gridy_high, gridy_resolution, log_dir)

    def marginal_1d_points(self, marg_idx, xs):
        """f(xs[k]) for arbitrary points; the first point is evaluated uncached like the grid dispatcher's first point."""
        xs = np.ascontiguousarray(xs, np.float64)
        ys = np.zeros_like(xs)
        bar_nu = self._bar_nu()
        rc = self._lib.mce_marginal_1d_points(self._h, int(marg_idx), _dp(bar_nu), len(xs), _dp(xs), _dp(ys))
        if rc < 0:
            raise RuntimeError(self._lib.mce_last_error().decode())
        return ys if rc > 0 else np.zeros(0)

    def cpdf_last_ms(self):
        return self._lib.mce_cpdf_last_ms(self._h)

    def _log_1d_grid(self, log_dir, marg_idx, xy):
        # CauchyCPDFGridDispatcher1D::log_point_grid (cpdf_ndim.hpp:2141-2202): binary (x, y) pairs appended per call to
        # cpdf_<idx>_<count>.bin and one "<num_points>" row per call in grid_elems_<idx>.txt
        os.makedirs(log_dir, exist_ok=True)
        counts = self.__dict__.setdefault("_cpdf_log_counts", {})
        k = counts.get((log_dir, marg_idx), 0)
        with open(os.path.join(log_dir, "grid_elems_%d.txt" % marg_idx), "w" if k == 0 else "a") as f:
            f.write("%d\n" % len(xy))
        xy.astype(np.float64).tofile(os.path.join(log_dir, "cpdf_%d_%d.bin" % (marg_idx, k + 1)))   # tag_count starts at 1
        counts[(log_dir, marg_idx)] = k + 1

    def print_conditional_mean_variance(self):   # est:513-522
        print("fz: %.16f + %.16fj" % (self.fz.real, self.fz.imag))
        print("Conditional Mean:\n", self.conditional_mean)
        print("Conditional Variance:\n", self.conditional_variance)

    def shutdown(self):
        if getattr(self, "_h", None):
            self._lib.mce_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass
