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
