"""Point-wise 1-D marginal cpdf, CPU side: the plain-C oracle restatement
(oracle/mce_oracle.c: mceo_marginal_1d_grid) and the emulated device kernels (mce_kern_cpdf.h over tests/emu) against the
golden grids produced by the UNMODIFIED reference (oracle/ref_cpdf.cpp -> tests/golden/*.cpdf.mced, tools/make_golden.sh).
Every value must be bit-identical: the sums run in the reference's term order."""
import os
import subprocess

import numpy as np
import pytest

from harness import ROOT, cpdf_steps, load_emu, oracle_cpdf1d, run_cpdf1d
from mceio import read_dump, read_scenario

GOLD = os.path.join(ROOT, "tests", "golden")
# scenario -> last step replayed on the CPU
ORACLE_CASES = {"lti3": 10, "lti4_2pnoise": 6, "syn5": 6, "leo5": 5, "lti2": 9, "lti3_3msmts": 12, "lti4_2msmts": 9, "syn8": 4, "homing3": 6}
EMU_CASES = {"lti3": 8, "lti4_2pnoise": 3, "syn5": 4, "leo5": 5, "leo7": 6, "lti2": 9, "lti3_3msmts": 12, "lti4_2msmts": 7, "syn8": 4, "homing3": 6}


@pytest.fixture(scope="module", autouse=True)
def _build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def _check(gold, got, steps):
    names = [n for n in gold if "/cpdf1d/i" in n and int(n.split("/")[0][1:]) in steps]
    assert names
    for n in names:
        assert n in got, n
        assert gold[n].shape == got[n].shape, n
        assert np.array_equal(gold[n].view(np.uint64), np.ascontiguousarray(got[n]).view(np.uint64)), \
            "%s: max abs diff %.3e" % (n, np.abs(gold[n] - got[n]).max())


@pytest.mark.parametrize("name", sorted(ORACLE_CASES))
def test_oracle_cpdf_matches_reference_golden(name, tmp_path):
    gold = read_dump(os.path.join(GOLD, name + ".cpdf.mced"))
    steps = [k for k in cpdf_steps(gold) if k <= ORACLE_CASES[name]]
    lo, hi, res = [float(v) for v in gold["cpdf1d/grid"]]
    got = oracle_cpdf1d(os.path.join(GOLD, name + ".mces"), str(tmp_path / "o.mced"), lo, hi, res, steps)
    _check(gold, got, steps)


@pytest.fixture(scope="module")
def emu():
    return load_emu(rebuild=True)


@pytest.mark.parametrize("name", sorted(EMU_CASES))
def test_emulated_cpdf_kernels_match_golden(emu, name):
    gold = read_dump(os.path.join(GOLD, name + ".cpdf.mced"))
    steps = [k for k in cpdf_steps(gold) if k <= EMU_CASES[name]]
    got = run_cpdf1d(emu, read_scenario(os.path.join(GOLD, name + ".mces")), gold, max_step=EMU_CASES[name])
    _check(gold, got, steps)


def test_cpdf_misuse_is_reported(emu):
    import ctypes as ct
    from harness import Session, _dp
    sc = read_scenario(os.path.join(GOLD, "lti3.mces"))
    s = Session(emu, sc)
    try:
        xy = np.zeros((4, 2)); nu = np.ones(sc.d)
        assert emu.mce_cpdf_grid_count(0.0, -1.0, 0.1) < 0            # grid_high <= grid_low
        assert emu.mce_marginal_1d_grid(s.h, 0, _dp(nu), 0.0, 0.3, 0.1, _dp(xy), 4) < 0   # not stepped yet (cpdf_ndim.hpp:1239)
        for r in sc.rec[:2]:
            s.step(r)
        assert emu.mce_marginal_1d_grid(s.h, 0, _dp(nu), 0.0, 0.3, 0.1, _dp(xy), 4) == 4
        assert emu.mce_marginal_1d_grid(s.h, sc.d, _dp(nu), 0.0, 0.3, 0.1, _dp(xy), 4) < 0   # state index out of range
        assert emu.mce_marginal_1d_grid(s.h, 0, _dp(nu), 0.0, 0.3, 0.1, _dp(xy), 3) < 0      # capacity too small
        assert abs(xy[:, 1].min()) >= 0
    finally:
        s.close()


# ---- 2-D marginal: on the CPU both the oracle and the emulated kernels call the host libm, like the reference ----
def _check2d(gold, got, steps):
    names = [n for n in gold if "/cpdf2d/i" in n and int(n.split("/")[0][1:]) in steps]
    assert names
    for n in names:
        assert n in got, n
        assert np.array_equal(gold[n].view(np.uint64), np.ascontiguousarray(got[n]).view(np.uint64)), \
            "%s: max abs diff %.3e" % (n, np.abs(gold[n] - got[n]).max())


@pytest.mark.parametrize("name", ["lti3", "lti4_2pnoise", "syn5", "leo5", "lti2", "lti3_3msmts", "lti4_2msmts", "syn8", "homing3"])
def test_oracle_cpdf2d_matches_reference_golden(name, tmp_path):
    gold = read_dump(os.path.join(GOLD, name + ".cpdf.mced"))
    steps = [k for k in cpdf_steps(gold) if k <= {"lti3": 8, "lti4_2pnoise": 6, "syn5": 6, "leo5": 5, "lti2": 9, "lti3_3msmts": 12, "lti4_2msmts": 9, "syn8": 4, "homing3": 6}[name]]
    lo, hi, res = [float(v) for v in gold["cpdf1d/grid"]]
    out = str(tmp_path / "o.mced")
    cmd = [os.path.join(ROOT, "oracle", "_build", "mce_oracle_run"), os.path.join(GOLD, name + ".mces"), out, "--max-steps", str(max(steps)),
           "--cpdf1d", repr(lo), repr(hi), repr(res), ",".join(str(k) for k in steps), "--cpdf2d"] + [repr(float(v)) for v in gold["cpdf2d/grid"]]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    _check2d(gold, read_dump(out), steps)


@pytest.mark.parametrize("name,last", [("lti3", 8), ("syn5", 4), ("leo5", 5), ("leo7", 6), ("lti2", 9), ("lti3_3msmts", 12), ("lti4_2msmts", 7), ("syn8", 4), ("homing3", 6)])
def test_emulated_cpdf2d_kernels_match_golden(emu, name, last):
    from harness import run_cpdf2d
    gold = read_dump(os.path.join(GOLD, name + ".cpdf.mced"))
    steps = [k for k in cpdf_steps(gold) if k <= last]
    got = run_cpdf2d(emu, read_scenario(os.path.join(GOLD, name + ".mces")), gold, max_step=last)
    _check2d(gold, got, steps)
