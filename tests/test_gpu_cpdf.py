"""GPU parity of the point-wise 1-D marginal cpdf (mce_kern_cpdf.h) through the C ABI: bit-identical to the grids of the
UNMODIFIED reference (tests/golden/*.cpdf.mced, produced by oracle/ref_cpdf.cpp) after small and large steps of five
scenarios, including 102 897 terms x 7 states of the 7-state LEO GPS window, and to the plain-C oracle run live on a
different grid."""
import os
import subprocess

import numpy as np
import pytest

from harness import ROOT, cpdf_steps, load_product, oracle_cpdf1d, run_cpdf1d
from mceio import read_dump, read_scenario

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _same(a, b, n):
    assert a.shape == b.shape, n
    assert np.array_equal(a.view(np.uint64), np.ascontiguousarray(b).view(np.uint64)), "%s: max abs diff %.3e" % (n, np.abs(a - b).max())


@pytest.mark.parametrize("name", ["lti3", "lti4_2pnoise", "syn5", "leo5", "leo7", "lti2", "lti3_3msmts", "lti4_2msmts", "syn8", "homing3"])
def test_cpdf1d_matches_reference_golden(name):
    gold = read_dump(os.path.join(GOLD, name + ".cpdf.mced"))
    got = run_cpdf1d(load_product(), read_scenario(os.path.join(GOLD, name + ".mces")), gold)
    names = [n for n in gold if "/cpdf1d/i" in n]
    assert len(names) == len(cpdf_steps(gold)) * int(gold["header"][0])
    for n in names:
        _same(gold[n], got[n], n)
        y = gold[n][:, 1]
        assert np.all(np.isfinite(y))


def test_cpdf1d_matches_oracle_on_another_grid(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    scen = os.path.join(GOLD, "lti3.mces")
    lo, hi, res, steps = -1.3, 0.9, 0.003, [4, 9]
    ref = oracle_cpdf1d(scen, str(tmp_path / "o.mced"), lo, hi, res, steps)
    gold = dict(ref)
    sc = read_scenario(scen)
    gold["cpdf1d/grid"] = np.array([lo, hi, res])
    gold["cpdf1d/bar_nu"] = np.array([0.25 + 1.5 * v - int(1.5 * v) for v in sc.root_point[: sc.d]])
    got = run_cpdf1d(load_product(), sc, gold)
    for n in [k for k in ref if "/cpdf1d/i" in k]:
        _same(ref[n], got[n], n)
    # the marginal integrates to ~1 over a window that holds the mass (sanity of the normalisation)
    y = got["s9/cpdf1d/i0"]
    assert 0.5 < y[:, 1].sum() * res <= 1.0 + 1e-9


def test_python_mirror_matches_capi(tmp_path):
    """cauchyfriendly_b200.CauchyEstimator.get_marginal_1D_pointwise_cpdf (cauchy_estimator.py:1003) returns the same grid
    and writes the reference's log files (cpdf_ndim.hpp:2141-2202)."""
    from cauchyfriendly_b200 import CauchyEstimator
    sc = read_scenario(os.path.join(GOLD, "lti3.mces"))
    gold = read_dump(os.path.join(GOLD, "lti3.cpdf.mced"))
    lo, hi, res = [float(v) for v in gold["cpdf1d/grid"]]
    est = CauchyEstimator(sc.A0, sc.p0, sc.b0, sc.steps, sc.d, sc.cmcc, sc.pncc, sc.p, root_point=sc.root_point, b_pert=sc.b_pert,
                          tr_search_idxs_ordering=sc.tr_order)
    est.bar_nu = gold["cpdf1d/bar_nu"]
    assert est.get_marginal_1D_pointwise_cpdf(0, lo, hi, res) == (None, None)
    for k in range(5):
        r = sc.rec[k]
        est.step(r.msmt, r.Phi, r.Gamma, r.beta, r.H, r.gamma)
    X, Y = est.get_marginal_1D_pointwise_cpdf(1, lo, hi, res, log_dir=str(tmp_path / "log"))
    _same(gold["s5/cpdf1d/i1"], np.stack([X, Y], 1), "python mirror")
    raw = np.fromfile(str(tmp_path / "log" / "cpdf_1_1.bin")).reshape(-1, 2)
    _same(gold["s5/cpdf1d/i1"], raw, "log file")
    assert open(str(tmp_path / "log" / "grid_elems_1.txt")).read().split() == [str(len(X))]
    assert est.cpdf_last_ms() > 0
    est.shutdown()


def test_branch_free_division_equals_ieee_division():
    """div_nobranch (mce_math.h) replays nvcc's division sequence without its branch; whenever it flags its result valid the
    value must be the IEEE quotient.  3 x 10^9 operand pairs: all exponents, exponents near 1, mantissa edge patterns."""
    import ctypes as ct
    from harness import Session
    lib = load_product()
    s = Session(lib, read_scenario(os.path.join(GOLD, "lti3.mces")))
    try:
        tot_ok = 0
        for seed in (1, 2, 3):
            out = (ct.c_ulonglong * 2)()
            assert lib.mce_debug_div_selftest(s.h, 1000 * 1000 * 1000, seed * 7919, out) == 0
            assert out[0] == 0, "%d of %d flagged-valid quotients differ from a / b" % (out[0], out[1])
            tot_ok += out[1]
        assert tot_ok > 2 * 10**9          # the flag is set for the bulk of the pairs (the test is not vacuous)
    finally:
        s.close()


# ---- 2-D marginal -------------------------------------------------------------------------------------------------
# The sums run in the reference's term order and every arithmetic step restates the reference's, but the cache holds
# atan2 / sin / cos values: CUDA's functions (<= 2 ulp) are not glibc's, so parity is a tolerance, written here.
CPDF2D_RTOL = 1e-9          # |z - z_ref| <= CPDF2D_RTOL * max|z_ref| over the grid


@pytest.mark.parametrize("name", ["lti3", "lti4_2pnoise", "syn5", "leo5", "leo7", "lti2", "lti3_3msmts", "lti4_2msmts", "syn8", "homing3"])
def test_cpdf2d_matches_reference_golden(name):
    from harness import run_cpdf2d
    gold = read_dump(os.path.join(GOLD, name + ".cpdf.mced"))
    names = [n for n in gold if "/cpdf2d/i" in n]
    assert names
    got = run_cpdf2d(load_product(), read_scenario(os.path.join(GOLD, name + ".mces")), gold)
    for n in names:
        a, b = gold[n], got[n]
        assert a.shape == b.shape, n
        assert np.array_equal(a[:, :2], b[:, :2]), n + ": grid coordinates"
        scale = np.abs(a[:, 2]).max()
        err = np.abs(a[:, 2] - b[:, 2]).max()
        assert err <= CPDF2D_RTOL * scale, "%s: max |dz| %.3e vs scale %.3e" % (n, err, scale)


def test_python_mirror_2d(tmp_path):
    from cauchyfriendly_b200 import CauchyEstimator
    sc = read_scenario(os.path.join(GOLD, "lti3.mces"))
    gold = read_dump(os.path.join(GOLD, "lti3.cpdf.mced"))
    g = [float(v) for v in gold["cpdf2d/grid"]]
    est = CauchyEstimator(sc.A0, sc.p0, sc.b0, sc.steps, sc.d, sc.cmcc, sc.pncc, sc.p, root_point=sc.root_point, b_pert=sc.b_pert,
                          tr_search_idxs_ordering=sc.tr_order)
    est.bar_nu = gold["cpdf1d/bar_nu"]
    for k in range(5):
        r = sc.rec[k]
        est.step(r.msmt, r.Phi, r.Gamma, r.beta, r.H, r.gamma)
    X, Y, Z = est.get_marginal_2D_pointwise_cpdf(0, 1, *g, log_dir=str(tmp_path / "log2"))
    ref = gold["s5/cpdf2d/i0_1"]
    ny, nx = Z.shape
    assert np.array_equal(np.stack([X.ravel(), Y.ravel()], 1), ref[:, :2])
    assert np.abs(Z.ravel() - ref[:, 2]).max() <= CPDF2D_RTOL * np.abs(ref[:, 2]).max()
    raw = np.fromfile(str(tmp_path / "log2" / "cpdf_01_1.bin")).reshape(-1, 3)
    assert np.array_equal(raw[:, 2], Z.ravel())
    assert open(str(tmp_path / "log2" / "grid_elems_01.txt")).read().strip() == "%d,%d" % (nx, ny)
    est.shutdown()
