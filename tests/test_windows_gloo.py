"""N > 1 path on CPU: the sliding-window bank shards its windows over torch.distributed ranks (gloo, world_size 2) and must
produce exactly the estimates of the unsharded bank.  The estimators run on the sequential emulation backend here (test
infrastructure, tests/emu); on a GPU box the same host logic drives libmce_b200.so (tests/test_gpu_windows.py)."""
import functools
import os
import subprocess
import sys

import numpy as np

from harness import ROOT

WORKER = r"""
import os, sys, functools
import numpy as np
sys.path.insert(0, os.environ["MCE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MCE_ROOT"], "tests"))
import torch.distributed as dist
from harness import load_emu
from cauchyfriendly_b200.estimator import CauchyEstimator
from cauchyfriendly_b200.windows import SlidingWindowBank
world = int(os.environ.get("WORLD_SIZE", "1"))
d = None
if world > 1:
    dist.init_process_group("gloo")
    d = dist
lib = load_emu()
Phi = np.array([[1.4, -0.6, -1.0], [-0.2, 1.0, 0.5], [0.6, -0.6, -0.2]])
Gamma = np.array([.1, .3, -.2]); H = np.array([[1.0, .5, .2]]); beta = [.1]; gamma = [.2]
rng = np.random.RandomState(3)
x = np.zeros(3); zs = []
for _ in range(10):
    x = Phi @ x + Gamma * 0.1 * rng.standard_cauchy(); zs.append(H[0] @ x + 0.2 * rng.standard_cauchy())
bank = SlidingWindowBank(4, np.eye(3), [.1, .08, .05], np.zeros(3), Phi, None, Gamma, beta, H, gamma,
                         estimator_cls=functools.partial(CauchyEstimator, _lib=lib), dist=d, seed=5, log_dir=os.environ.get("MCE_LOG"))
out = []
for z in zs:
    xh, Ph, xa, Pa = bank.step([z])
    out.append(np.concatenate([xh, Ph.ravel(), xa, Pa.ravel(), [bank.moment_info["win_idx"][-1]]]))
bank.shutdown()
if d is None or d.get_rank() == 0:
    np.save(os.environ["MCE_OUT"], np.array(out))
if d is not None:
    d.barrier(); d.destroy_process_group()
"""


def _run(world, out, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MCE_ROOT=ROOT, MCE_OUT=str(out), MCE_LOG=str(out) + ".log")
    if world == 1:
        subprocess.check_call([sys.executable, str(script)], env=env, timeout=600)
    else:
        subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                               "--master-port", "29517", str(script)], env=env, timeout=900)
    return np.load(out)


def test_sharded_window_bank_matches_single_process(tmp_path):
    from harness import load_emu
    load_emu(rebuild=False)
    a = _run(1, tmp_path / "w1.npy", tmp_path)
    b = _run(2, tmp_path / "w2.npy", tmp_path)
    assert a.shape == b.shape and a.shape[0] == 10
    assert np.array_equal(a, b)
    assert np.isfinite(a).all()
    assert set(a[:, -1].astype(int)) - {0} != set()      # windows other than the first did become "best" after the warm-up
    # the reference's log layout (cauchy_windows.hpp:1237-1421), written by rank 0 and read back with the reference's parsers
    from cauchyfriendly_b200.windows import LOG_NAMES, load_cauchy_log_folder, load_data
    for world_dir in (str(tmp_path / "w1.npy") + ".log", str(tmp_path / "w2.npy") + ".log"):
        logs = load_cauchy_log_folder(world_dir)
        assert logs["x"].shape == (10, 3) and logs["P"].shape == (10, 3, 3)
        assert np.allclose(logs["x"], a[:, :3], rtol=0, atol=1e-15) and np.allclose(logs["P"].reshape(10, 9), a[:, 3:12], rtol=0, atol=1e-15)
        best = [int(line.split(":")[0]) for line in open(os.path.join(world_dir, "cond_means.txt"))]
        assert best == list(a[:, -1].astype(int))
        assert sorted(os.listdir(world_dir)) == sorted(list(LOG_NAMES) + ["windows"])
        w0 = load_data(os.path.join(world_dir, "windows", "win0", "cond_means.txt"))
        assert w0.shape[1] == 3 and w0.shape[0] >= 4
        codes = [int(line.split(":")[1]) for line in open(os.path.join(world_dir, "numeric_error_codes.txt"))]
        assert len(codes) == 10


def test_speyer_init_reproduces_mean_and_covariance():
    """A one-term CF built by Speyer's initialisation must return (x1_hat, Var) after its first measurement update:
    checked in closed form through the oracle-validated first step of the emulated kernels."""
    from cauchyfriendly_b200.estimator import CauchyEstimator
    from cauchyfriendly_b200.windows import speyers_window_init
    from harness import load_emu
    lib = load_emu()
    rng = np.random.RandomState(0)
    n = 3
    Q = rng.randn(n, n); Var = Q @ Q.T + 0.5 * np.eye(n)
    x1 = rng.randn(n); H = np.array([1.0, 0.5, 0.2]); gamma = 0.2; z = 0.37
    A0, p0, b0 = speyers_window_init(x1, Var, H, gamma, z)
    est = CauchyEstimator(A0, p0, b0, 4, n, 0, 1, 1, _lib=lib)
    est.step(z, np.eye(n), np.ones(n), [0.1], H, gamma)
    assert np.allclose(est.conditional_mean.real, x1, rtol=1e-9, atol=1e-12)
    assert np.allclose(est.conditional_variance.real, Var, rtol=1e-8, atol=1e-11)
    est.shutdown()
