"""Sliding-window bank over B200 estimators -- host-side mirror of the reference's PySlidingWindowManager
(scripts/swig/cauchy/cauchy_estimator.py:1092-1333) and of SlidingWindowManager::step (include/cauchy_windows.hpp:1025-1165)
for LTI / LTV systems.

W estimators ("windows") run staggered: at steady state window w has folded in 1..W time steps; the fullest window without
covariance error flags provides the estimate, and the window that just emptied is re-seeded from that estimate with
Speyer's initialisation (cauchy_util.hpp:23-112) and the last measurement.

Multi-GPU: windows are independent estimators, so they are sharded over ranks (window w lives on rank w % world_size, one
process per GPU, torch.distributed).  The only exchange per step is an all-gather of each window's (count, error code, fz,
mean, covariance) -- 3 + n + n*n doubles per window -- after which every rank runs the same selection logic; there is no
data-path collective.  This is the reference's own unit of distribution (one forked process per window,
cauchy_windows.hpp:353-376)."""
import math
import os

import numpy as np

from .estimator import CauchyEstimator

COV_UNSTABLE_FINAL = 1 << 1   # ERROR_COVARIANCE_UNSTABLE_CURRENT_STEP_FINAL_MSMT (cauchy_constants.hpp:114)
COV_DNE = 1 << 3              # ERROR_COVARIANCE_AT_CURRENT_STEP_DNE
MEAN_DNE = 1 << 7             # ERROR_MEAN_AT_CURRENT_STEP_DNE
FZ_NEGATIVE = 1 << 9          # ERROR_FZ_NEGATIVE


def _pythag(a, b):      # evs_pythag, eig_solve.hpp:254-261
    absa, absb = abs(a), abs(b)
    if absa > absb:
        t = absb / absa
        return absa * math.sqrt(1.0 + t * t)
    if absb == 0.0:
        return 0.0
    t = absa / absb
    return absb * math.sqrt(1.0 + t * t)


def _esign(a, b):       # EVS_SIGN, eig_solve.hpp:11
    return abs(a) if b >= 0.0 else -abs(a)


def sym_eig(A, n):
    """Eigenvalues / eigenvectors of a symmetric matrix exactly as the reference computes them (sym_eig, eig_solve.hpp:404-427:
    Householder tridiagonalisation evs_tred2 :263-331 followed by QL with implicit shifts evs_tqli :334-389), operation for
    operation in IEEE double -- LAPACK's eigenvectors differ from these in sign and last bits, and the windows re-seeded from
    them would drift from the reference's by ~1e-9.  Returns (evals[n], evecs[n][n]) with eigenvector k in COLUMN k."""
    a = [[float(A[i][j]) for j in range(n)] for i in range(n)]
    d = [0.0] * n
    e = [0.0] * n
    for i in range(n - 1, 0, -1):           # evs_tred2
        l = i - 1
        h = scale = 0.0
        if l > 0:
            for k in range(l + 1):
                scale += abs(a[i][k])
            if scale == 0.0:
                e[i] = a[i][l]
            else:
                for k in range(l + 1):
                    a[i][k] /= scale
                    h += a[i][k] * a[i][k]
                f = a[i][l]
                g = -math.sqrt(h) if f >= 0.0 else math.sqrt(h)
                e[i] = scale * g
                h -= f * g
                a[i][l] = f - g
                f = 0.0
                for j in range(l + 1):
                    a[j][i] = a[i][j] / h
                    g = 0.0
                    for k in range(j + 1):
                        g += a[j][k] * a[i][k]
                    for k in range(j + 1, l + 1):
                        g += a[k][j] * a[i][k]
                    e[j] = g / h
                    f += e[j] * a[i][j]
                hh = f / (h + h)
                for j in range(l + 1):
                    f = a[i][j]
                    e[j] = g = e[j] - hh * f
                    for k in range(j + 1):
                        a[j][k] -= (f * e[k] + g * a[i][k])
        else:
            e[i] = a[i][l]
        d[i] = h
    d[0] = 0.0
    e[0] = 0.0
    for i in range(n):
        l = i
        if d[i] != 0.0:
            for j in range(l):
                g = 0.0
                for k in range(l):
                    g += a[i][k] * a[k][j]
                for k in range(l):
                    a[k][j] -= g * a[k][i]
        d[i] = a[i][i]
        a[i][i] = 1.0
        for j in range(l):
            a[j][i] = a[i][j] = 0.0
    z = a                                   # evs_tqli
    for i in range(1, n):
        e[i - 1] = e[i]
    e[n - 1] = 0.0
    for l in range(n):
        it = 0
        while True:
            m = l
            while m < n - 1:
                dd = abs(d[m]) + abs(d[m + 1])
                if abs(e[m]) + dd == dd:
                    break
                m += 1
            if m != l:
                if it == 30:
                    it += 1
                    break
                it += 1
                g = (d[l + 1] - d[l]) / (2.0 * e[l])
                r = _pythag(g, 1.0)
                g = d[m] - d[l] + e[l] / (g + _esign(r, g))
                s = c = 1.0
                p = 0.0
                i = m - 1
                broke = False
                while i >= l:
                    f = s * e[i]
                    b = c * e[i]
                    r = _pythag(f, g)
                    e[i + 1] = r
                    if r == 0.0:
                        d[i + 1] -= p
                        e[m] = 0.0
                        broke = True
                        break
                    s = f / r
                    c = g / r
                    g = d[i + 1] - p
                    r = (d[i] - g) * s + 2.0 * c * b
                    p = s * r
                    d[i + 1] = g + p
                    g = c * r - b
                    for k in range(n):
                        f = z[k][i + 1]
                        z[k][i + 1] = s * z[k][i] + c * f
                        z[k][i] = c * z[k][i] - s * f
                    i -= 1
                if broke and r == 0.0 and i >= l:
                    continue
                d[l] -= p
                e[l] = g
                e[m] = 0.0
            if m == l:
                break
    return d, z


def speyers_window_init(x1_hat, Var, H, gamma, z1):
    """Speyer's window initialisation restated operation for operation from cauchy_util.hpp:23-112 (window_var_boost = NULL):
    returns A0 (n x n), p0, b0 of a one-term characteristic function whose first measurement update reproduces (x1_hat, Var).
    Scalar IEEE-double loops in the reference's order (n <= 8, once per bank step), so the re-seeded window starts from the same
    bits as the reference's."""
    n = len(x1_hat)
    x1 = [float(v) for v in x1_hat]
    V = [[float(np.asarray(Var)[i][j]) for j in range(n)] for i in range(n)]
    Hh = [float(v) for v in np.asarray(H, np.float64).reshape(-1)]
    ev, _ = sym_eig(V, n)
    if any(w < -1e-5 for w in ev):          # COV_EIGENVALUE_TOLERANCE (:46-68): make the covariance more positive definite
        boost = -1e-5 - min(ev)
        for i in range(n):
            V[i][i] += boost
    HH = [[Hh[i] * Hh[j] for j in range(n)] for i in range(n)]                 # inner_mat_prod(H, work, 1, N): sum of one product
    HH = [[0.0 + HH[i][j] for j in range(n)] for i in range(n)]
    W2 = [[0.0] * n for _ in range(n)]
    for i in range(n):                      # matmatmul(Var, work, work2)
        for j in range(n):
            sm = 0.0
            for k in range(n):
                sm += V[i][k] * HH[k][j]
            W2[i][j] = sm
    W = [[0.0] * n for _ in range(n)]
    for i in range(n):                      # matmatmul(work2, Var, work)
        for j in range(n):
            sm = 0.0
            for k in range(n):
                sm += W2[i][k] * V[k][j]
            W[i][j] = sm
    Hx = 0.0
    for i in range(n):
        Hx += Hh[i] * x1[i]
    scale = gamma * gamma + (z1 - Hx) * (z1 - Hx)
    inv = 1.0 / scale
    M = [[V[i][j] + (W[i][j] * inv) * 1.0 for j in range(n)] for i in range(n)]     # scale_mat(work, 1/scale); add_mat(A_0, Var, work, 1.0)
    eigs, vec = sym_eig(M, n)               # vec: eigenvector k in column k (A_0 before reflect_array)
    scale2 = (z1 - Hx) / scale
    VH = [0.0] * n
    for i in range(n):                      # matvecmul(Var, H, work)
        sm = 0.0
        for j in range(n):
            sm += V[i][j] * Hh[j]
        VH[i] = sm
    b0 = [x1[i] - VH[i] * scale2 for i in range(n)]
    HA = [0.0] * n
    for i in range(n):                      # matvecmul(A_0, H, work2, N, N, true): sum_j A_0[i + j*N] * H[j]
        sm = 0.0
        for j in range(n):
            sm += vec[j][i] * Hh[j]
        HA[i] = sm
    HVH = 0.0
    for i in range(n):
        HVH += Hh[i] * VH[i]
    scale3 = (scale + HVH) / gamma
    p0 = []
    for i in range(n):
        v = eigs[i] / scale3
        v *= HA[i]
        v /= float((HA[i] > 0) - (HA[i] < 0))
        p0.append(v)
    A0 = np.array([[vec[j][i] for j in range(n)] for i in range(n)], np.float64)     # reflect_array: rows are the eigenvectors
    return A0, np.array(p0, np.float64), np.array(b0, np.float64)


LOG_NAMES = ("cond_means.txt", "cond_covars.txt", "norm_factors.txt", "cerr_cond_means.txt", "cerr_cond_covars.txt",
             "cerr_norm_factors.txt", "numeric_error_codes.txt")


def _fmt(values):
    # log_double_array_to_file, array_logging.hpp:60-70: "%.16lf" separated by single blanks
    return " ".join("%.16f" % float(v) for v in np.atleast_1d(values))


class WindowLogFiles:
    """The reference's window-manager log layout: `<log_dir>/<name>` hold the best window per step as "<win_idx>:<values>"
    lines, `<log_dir>/windows/win<i>/<name>` every window's own history without the prefix (cauchy_windows.hpp:1237-1373;
    read back by scripts/cauchy_plotter.py and cauchy_estimator.py:1337-1373 load_cauchy_log_folder)."""

    def __init__(self, log_dir, num_windows, log_windows=True):
        self.log_dir = log_dir.rstrip("/")
        os.makedirs(self.log_dir, exist_ok=True)
        self.best = [open(os.path.join(self.log_dir, nme), "w") for nme in LOG_NAMES]
        self.wins = None
        if log_windows:
            self.wins = []
            for w in range(num_windows):
                d = os.path.join(self.log_dir, "windows", "win%d" % w)
                os.makedirs(d, exist_ok=True)
                self.wins.append([open(os.path.join(d, nme), "w") for nme in LOG_NAMES])

    @staticmethod
    def _rows(row, n):
        nn = n * n
        return (_fmt(row[3:3 + n]), _fmt(row[3 + n:3 + n + nn]), _fmt(row[2]), _fmt(row[4 + n + nn]), _fmt(row[5 + n + nn]),
                _fmt(row[3 + n + nn]), "%d" % int(row[1]))

    def write_best(self, win_idx, row, n):
        for f, txt in zip(self.best, self._rows(row, n)):
            f.write("%d:%s\n" % (win_idx, txt))
            f.flush()

    def write_window(self, w, row, n):
        if self.wins is None:
            return
        for f, txt in zip(self.wins[w], self._rows(row, n)):
            f.write(txt + "\n")
            f.flush()

    def close(self):
        for f in self.best + ([g for fs in self.wins for g in fs] if self.wins else []):
            f.close()


def load_window_data(path):          # cauchy_estimator.py:1337-1340
    return np.array([[float(v) for v in line.split(":")[1].split(" ")] for line in open(path)])


def load_data(path):                 # cauchy_estimator.py:1343-1346
    return np.array([[float(v) for v in line.split(" ")] for line in open(path)])


def load_cauchy_log_folder(log_dir, with_win_logging=True):      # cauchy_estimator.py:1348-1373
    log_dir = log_dir if log_dir.endswith("/") else log_dir + "/"
    rd = load_window_data if with_win_logging else load_data
    covars = rd(log_dir + "cond_covars.txt")
    n = int(np.sqrt(covars.shape[1]))
    return {"x": rd(log_dir + "cond_means.txt"), "P": covars.reshape(covars.shape[0], n, n), "cerr_x": rd(log_dir + "cerr_cond_means.txt"),
            "cerr_P": rd(log_dir + "cerr_cond_covars.txt"), "cerr_fz": rd(log_dir + "cerr_norm_factors.txt")}


class SlidingWindowBank:
    """LTI/LTV sliding-window Cauchy estimator (W windows of depth W).

    step(msmts, controls) mirrors PySlidingWindowManager.step and returns (xhat, Phat, wavg_xhat, wavg_Phat)."""

    def __init__(self, num_windows, A0, p0, b0, Phi, B, Gamma, beta, H, gamma, *, estimator_cls=CauchyEstimator, est_kwargs=None,
                 dist=None, seed=0, debug_print=False, concurrent=False, log_dir=None, log_windows=True, selection="python"):
        # The reference has two window managers with different usable-window rules: PySlidingWindowManager skips windows whose
        # covariance is unstable on the final measurement or does not exist (cauchy_estimator.py:1181-1204, selection="python");
        # the C++ SlidingWindowManager skips fz < 0, covariance-DNE and mean-DNE (stratgey_choose_fullest_window_first,
        # cauchy_windows.hpp:1578-1610, selection="cpp").
        assert selection in ("python", "cpp")
        self.selection = selection
        self.W = int(num_windows)
        self.n = int(np.asarray(p0).size)
        self.Phi = np.asarray(Phi, np.float64).reshape(self.n, self.n)
        self.Gamma = np.asarray(Gamma, np.float64).reshape(self.n, -1)
        self.pncc = self.Gamma.shape[1]
        self.beta = np.asarray(beta, np.float64).reshape(self.pncc)
        self.H = np.asarray(H, np.float64).reshape(-1, self.n)
        self.p = self.H.shape[0]
        self.gamma = np.asarray(gamma, np.float64).reshape(self.p)
        self.B = None if B is None else np.asarray(B, np.float64).reshape(self.n, -1)
        self.cmcc = 0 if self.B is None else self.B.shape[1]
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.debug_print = debug_print
        # concurrent=True steps this rank's windows from a thread pool: every estimator owns its CUDA streams and the C ABI
        # releases the GIL, so the launch-latency-bound young windows overlap with the full one on the same GPU
        self._pool = None
        if concurrent:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=max(1, (int(num_windows) + self.world - 1) // self.world))
        kw = dict(est_kwargs or {})
        # every window is created with the same (seeded) root point / perturbation, on the rank that owns it
        self.ests = {}
        for w in range(self.W):
            if w % self.world == self.rank:
                self.ests[w] = estimator_cls(A0, p0, b0, self.W, self.n, self.cmcc, self.pncc, self.p, seed=seed, **kw)
                self.ests[w].set_win_num(w + 1)
        self.win_counts = np.zeros(self.W, np.int64)
        self.step_idx = 0
        self.moment_info = {"x": [], "P": [], "fz": [], "win_idx": [], "err_code": []}
        self.avg_moment_info = {"x": [], "P": [], "win_idx": [], "err_code": []}
        n = self.n
        # per window: count, err, Re fz, mean[n], cov[n*n], then Im fz, max |Im mean|, max |Im cov| (last measurement)
        self._stats = np.zeros((self.W, 3 + n + n * n + 3))
        self._last_msmts = None
        # the reference's log files (cauchy_windows.hpp:1237-1476, array_logging.hpp:60-98), written by rank 0
        self._log = None
        if log_dir is not None and self.rank == 0:
            self._log = WindowLogFiles(str(log_dir), self.W, log_windows)

    # ---- one estimator, one time step: p measurement updates (PyCauchyEstimator._call_step, cauchy_estimator.py:612-656) ----
    def _step_window(self, w, msmts, controls, first_msmt=0):
        est = self.ests[w]
        u = None if controls is None or self.cmcc == 0 else np.asarray(controls, np.float64)
        for i in range(first_msmt, self.p):
            est.step(msmts[i], self.Phi, self.Gamma, self.beta, self.H[i], self.gamma[i], self.B, u)
        n = self.n
        row = self._stats[w]
        row[1] = est.numeric_moment_errors
        row[2] = est.fz.real            # the public field both reference managers report: 1 after a non-final step (est:1172-1176)
        row[3:3 + n] = est.conditional_mean.real
        row[3 + n:3 + n + n * n] = est.conditional_variance.real.ravel()
        row[3 + n + n * n] = est.fz.imag                                             # save_window_data, cauchy_windows.hpp:1478-1503
        row[4 + n + n * n] = np.abs(est.conditional_mean.imag).max()
        row[5 + n + n * n] = np.abs(est.conditional_variance.imag).max()

    def _exchange(self):
        """All ranks end up with every window's statistics (the only communication of a step)."""
        if self.dist is None or self.world == 1:
            return
        import torch
        mine = torch.from_numpy(self._stats.copy())
        owner = torch.tensor([w % self.world for w in range(self.W)])
        mine[owner != self.rank] = 0
        backend = self.dist.get_backend()
        if backend == "nccl":
            mine = mine.cuda()
        self.dist.all_reduce(mine)              # rows are disjoint across ranks: a sum is a gather
        self._stats[:] = mine.cpu().numpy()

    def _best_window(self):
        okays = np.zeros(self.W, bool)
        idxs = []
        for i in range(self.W):
            if self.win_counts[i] > 0:
                err = int(self._stats[i, 1])
                bad = (err & (FZ_NEGATIVE | COV_DNE | MEAN_DNE)) if self.selection == "cpp" else (err & (COV_UNSTABLE_FINAL | COV_DNE))
                if not bad:
                    idxs.append((i, self.win_counts[i]))
                    okays[i] = True
        if self.step_idx == 0:
            best, okays[0] = 0, True
        else:
            if not idxs:
                raise RuntimeError("No window is available without an error code!")
            best = sorted(idxs, key=lambda x: x[1], reverse=True)[0][0]
        return best, okays

    def _mean_cov(self, w):
        n = self.n
        return self._stats[w, 3:3 + n].copy(), self._stats[w, 3 + n:3 + n + n * n].reshape(n, n).copy()

    def step(self, msmts, controls=None):
        msmts = np.asarray(msmts, np.float64).reshape(self.p)
        min_idx = max_idx = None
        if self.step_idx == 0:
            if 0 in self.ests:
                self._step_window(0, msmts, controls)
            self.win_counts[0] += 1
        else:
            max_idx = int(np.argmax(self.win_counts))
            min_idx = int(np.argmin(self.win_counts))
            mine = [w for w in range(self.W) if self.win_counts[w] > 0 and w in self.ests]
            if self._pool is not None and len(mine) > 1:
                list(self._pool.map(lambda w: self._step_window(w, msmts, controls), mine))
            else:
                for w in mine:
                    self._step_window(w, msmts, controls)
            for w in range(self.W):
                if self.win_counts[w] > 0:
                    self.win_counts[w] += 1
        self._stats[:, 0] = self.win_counts
        self._exchange()
        best, okays = self._best_window()
        xhat, Phat = self._mean_cov(best)
        self.moment_info["x"].append(xhat); self.moment_info["P"].append(Phat)
        self.moment_info["fz"].append(self._stats[best, 2]); self.moment_info["win_idx"].append(best)
        self.moment_info["err_code"].append(int(self._stats[best, 1]))
        # weighted average over the usable windows (cauchy_estimator.py:1218-1259)
        wsum, xavg, Pavg, err_or = 0.0, np.zeros(self.n), np.zeros((self.n, self.n)), 0
        for w in range(self.W):
            if self.win_counts[w] > 0 and okays[w]:
                f = self.win_counts[w] / self.W
                x, P = self._mean_cov(w)
                wsum += f; xavg += x * f; Pavg += P * f; err_or |= int(self._stats[w, 1])
        xavg /= wsum; Pavg /= wsum
        self.avg_moment_info["x"].append(xavg); self.avg_moment_info["P"].append(Pavg)
        self.avg_moment_info["win_idx"].append(-1); self.avg_moment_info["err_code"].append(err_or)
        if self._log is not None:           # sequential_logger_best_window / _all_windows, cauchy_windows.hpp:1378-1421
            self._log.write_best(best, self._stats[best], self.n)
            for w in range(self.W):
                if self.win_counts[w] > 0:
                    self._log.write_window(w, self._stats[w], self.n)
        # re-seed the empty window about the best window's estimate and the last measurement (reset_about_estimator, :888-923)
        if self.step_idx > 0:
            if min_idx in self.ests:
                A0, p0, b0 = speyers_window_init(xhat, Phat, self.H[self.p - 1], self.gamma[self.p - 1], msmts[self.p - 1])
                est = self.ests[min_idx]
                est.reset()
                est.reinitialize_start_statistics(A0, p0, b0)
                self._step_window(min_idx, msmts, None, first_msmt=self.p - 1)
                est.master_step = self.p
            self.win_counts[min_idx] += 1
            if self.win_counts[max_idx] == self.W:
                if max_idx in self.ests:
                    self.ests[max_idx].reset()
                self.win_counts[max_idx] = 0
        self.step_idx += 1
        return xhat, Phat, xavg, Pavg

    def shutdown(self):
        if self._log is not None:
            self._log.close()
            self._log = None
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None
        for e in self.ests.values():
            e.shutdown()
        self.ests = {}
