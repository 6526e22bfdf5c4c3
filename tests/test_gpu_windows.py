"""The sliding-window bank (cauchyfriendly_b200/windows.py, mirror of PySlidingWindowManager) on the GPU path: the same
host logic must give the same estimates over libmce_b200.so as over the emulated kernels (both bit-exact restatements)."""
import functools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(est_cls, concurrent=False):
    from cauchyfriendly_b200.windows import SlidingWindowBank
    Phi = np.array([[1.4, -0.6, -1.0], [-0.2, 1.0, 0.5], [0.6, -0.6, -0.2]])
    Gamma = np.array([.1, .3, -.2]); H = np.array([[1.0, .5, .2]])
    rng = np.random.RandomState(3)
    x = np.zeros(3); zs = []
    for _ in range(14):
        x = Phi @ x + Gamma * 0.1 * rng.standard_cauchy(); zs.append(H[0] @ x + 0.2 * rng.standard_cauchy())
    bank = SlidingWindowBank(5, np.eye(3), [.1, .08, .05], np.zeros(3), Phi, None, Gamma, [.1], H, [.2], estimator_cls=est_cls, seed=5, concurrent=concurrent)
    out = []
    for z in zs:
        xh, Ph, xa, Pa = bank.step([z])
        out.append(np.concatenate([xh, Ph.ravel(), xa, Pa.ravel(), [bank.moment_info["win_idx"][-1]]]))
    bank.shutdown()
    return np.array(out)


def test_window_bank_gpu_matches_emulated_kernels():
    from cauchyfriendly_b200.estimator import CauchyEstimator
    from harness import load_emu
    gpu = _run(CauchyEstimator)
    emu = _run(functools.partial(CauchyEstimator, _lib=load_emu()))
    assert np.array_equal(gpu, emu)
    assert np.isfinite(gpu).all() and gpu.shape[0] == 14


def test_window_bank_concurrent_windows_same_results():
    """Windows stepped from a thread pool (one estimator per thread, own CUDA streams) give the same estimates."""
    from cauchyfriendly_b200.estimator import CauchyEstimator
    assert np.array_equal(_run(CauchyEstimator), _run(CauchyEstimator, concurrent=True))


def _load_best(path):
    """`<win_idx>:<values>` lines of the reference's bank logs (cauchy_windows.hpp:1378-1400)."""
    idx, rows = [], []
    for line in open(path):
        a, b = line.split(":")
        idx.append(int(a)); rows.append([float(v) for v in b.split()])
    return np.array(idx), np.array(rows)


def test_window_bank_matches_the_reference_window_manager(tmp_path):
    """The Python bank against the REFERENCE's SlidingWindowManager: tests/golden/winbank_cpu1/ holds the log files the
    unmodified NUM_CPUS=1 reference wrote for the inputs of src/window_manager.cpp (srand(11), 201 measurements, 8 windows;
    tests/dropin/winbank_dropin.cpp) together with the measurement sequence.  Replaying the measurements through
    SlidingWindowBank must select the same window at every step and write the same means / covariances / normalisation
    factors into the same log layout -- digit for digit (16 decimals): the estimators are bit-exact, the bank's Speyer
    initialisation restates the reference's own symmetric eigen-solver operation for operation (windows.py::sym_eig), and
    the windows get the root_point / b_pert the reference's window processes drew."""
    import os
    from harness import ROOT
    from cauchyfriendly_b200.estimator import CauchyEstimator
    from cauchyfriendly_b200.windows import SlidingWindowBank
    gold = os.path.join(ROOT, "tests", "golden", "winbank_cpu1")
    zs = np.loadtxt(os.path.join(gold, "msmts.txt"))
    Phi = np.array([[1.4, -0.6, -1.0], [-0.2, 1.0, 0.5], [0.6, -0.6, -0.2]])
    Gamma = np.array([.1, .3, -.2]); H = np.array([[1.0, .5, .2]])
    rb = np.loadtxt(os.path.join(gold, "root_point_b_pert.txt"))       # the vectors every window of the reference run drew with rand()
    bank = SlidingWindowBank(8, np.eye(3), [.10, .08, .05], np.zeros(3), Phi, None, Gamma, [.1], H, [.2], estimator_cls=CauchyEstimator,
                             est_kwargs=dict(root_point=rb[:3], b_pert=rb[3:]), log_dir=str(tmp_path / "logs"), log_windows=False, selection="cpp")
    for z in zs:
        bank.step([z])
    bank.shutdown()
    worst = {}
    for name in ("cond_means.txt", "cond_covars.txt", "norm_factors.txt", "cerr_cond_means.txt", "cerr_cond_covars.txt", "cerr_norm_factors.txt"):
        gi, gv = _load_best(os.path.join(gold, name))
        oi, ov = _load_best(str(tmp_path / "logs" / name))
        assert gv.shape == ov.shape and gi.shape == oi.shape, name
        assert np.array_equal(gi, oi), "%s: a different window was selected at steps %s" % (name, np.nonzero(gi != oi)[0][:10])
        rel = np.max(np.abs(gv - ov), axis=1) / np.max(np.abs(gv), axis=1)
        worst[name] = float(rel.max())
        assert np.array_equal(gv, ov), "%s: worst row deviates by %.3e (step %d)" % (name, rel.max(), int(rel.argmax()))
    gi, gv = _load_best(os.path.join(gold, "numeric_error_codes.txt"))
    oi, ov = _load_best(str(tmp_path / "logs" / "numeric_error_codes.txt"))
    assert np.array_equal(gv, ov)
    print("window bank vs reference manager, worst relative row deviation:", worst)
