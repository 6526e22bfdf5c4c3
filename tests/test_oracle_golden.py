"""Pins the plain-C oracle (oracle/mce_oracle.c) against the golden dumps produced by the UNMODIFIED reference
(oracle/_ref/ref_run_cpu1, see tools/make_golden.sh): every array bit for bit -- hyperplanes, weights, coalignment
maps, FTR flag arrays, sorted sign-vector keys, G values, term counts and moments."""
import os
import subprocess

import pytest

from compare import compare_dumps
from harness import ROOT, oracle_dump
from mceio import read_dump

GOLD = os.path.join(ROOT, "tests", "golden")
# scenario -> number of steps replayed on the CPU (kept small enough for a few-minute CPU suite)
CASES = {"lti3": 9, "lti2": 10, "lti4": 7, "lti3_3msmts": 15, "lti4_2pnoise": 6, "lti4_2msmts": 11,
         "syn2": 12, "syn3": 9, "syn4": 7, "syn5": 6, "syn6": 5, "syn7": 5, "syn8": 4, "leo7": 8, "leo5": 9, "homing3": 8,
         "homing_real": 8, "leo7_deep16": 8, "lti3_deep": 9, "lti4_2pnoise_deep": 6, "lti3_3msmts_deep": 12}


@pytest.fixture(scope="module", autouse=True)
def _build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])


def _upto(d, k):
    return {n: v for n, v in d.items() if n == "header" or int(n.split("/")[0][1:]) <= k}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name, tmp_path):
    steps = CASES[name]
    gold = _upto(read_dump(os.path.join(GOLD, name + ".ref.mced")), steps)
    got = oracle_dump(os.path.join(GOLD, name + ".mces"), str(tmp_path / "o.mced"), full_upto=100, max_steps=steps)
    # the oracle dumped every step in full; compare exactly the arrays the golden file holds
    got = {n: v for n, v in got.items() if n in gold}
    probs = compare_dumps(gold, got, float_rtol=0.0)
    assert not probs, "\n".join(probs[:20])


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_run_cpu1")), reason="reference runner not built")
def test_oracle_matches_live_reference(tmp_path):
    """Same check against the reference binary itself (when it travelled with the repo), on a case outside the golden set."""
    scen = os.path.join(GOLD, "syn4.mces")
    ref = oracle_dump(scen, str(tmp_path / "r.mced"), full_upto=100, max_steps=7, use_ref=True)
    got = oracle_dump(scen, str(tmp_path / "o.mced"), full_upto=100, max_steps=7)
    probs = compare_dumps(ref, got, float_rtol=0.0)
    assert not probs, "\n".join(probs[:20])


def test_reference_8_thread_build_delta_report(capsys):
    """Informational: the reference ships with NUM_CPUS = 8 (cauchy_constants.hpp:69) and its results depend
    on that constant -- threads walk the terms in another order, other reduction-group roots are elected, partial moment sums
    are added in another order.  Parity in this repository is against the NUM_CPUS = 1 build.  This test reads
    the 8-thread golden (tests/golden/leo7.ref8.mced, made by oracle/_ref/ref_run_cpu8) and reports the per-MU deltas; it
    asserts equality only where the two builds MUST agree: the reference stays single-threaded below its threading threshold
    (MUs 1-4 of the LEO7 window), so counts and moments there are bit-identical."""
    import numpy as np
    g1 = read_dump(os.path.join(GOLD, "leo7.ref.mced"))
    g8 = read_dump(os.path.join(GOLD, "leo7.ref8.mced"))
    rows = []
    for k in range(1, 13):
        c1, c8 = int(g1["s%d/muc/counts" % k].sum()), int(g8["s%d/muc/counts" % k].sum())
        f1 = int(g1["s%d/ftr/counts" % k].sum()) if "s%d/ftr/counts" % k in g1 else -1
        f8 = int(g8["s%d/ftr/counts" % k].sum()) if "s%d/ftr/counts" % k in g8 else -1
        m1, m8 = g1["s%d/moments" % k], g8["s%d/moments" % k]
        d = 7
        fz_rel = abs(m1[0] - m8[0]) / abs(m1[0])
        mean_rel = float(np.max(np.abs(m1[1:1 + d] - m8[1:1 + d])) / np.max(np.abs(m1[1:1 + d])))
        cov_rel = float(np.max(np.abs(m1[1 + d:] - m8[1 + d:])) / np.max(np.abs(m1[1 + d:])))
        rows.append((k, c1, c8, f1, f8, fz_rel, mean_rel, cov_rel))
        if k <= 4:
            assert (c1, f1) == (c8, f8) and fz_rel == 0 and mean_rel == 0 and cov_rel == 0
    with capsys.disabled():
        print("\nreference NUM_CPUS=1 vs NUM_CPUS=8, LEO7 window: MU | after MUC (1 / 8 threads) | after FTR (1 / 8) | fz, mean, cov deltas relative to max|.|")
        for r in rows:
            print("  %2d | %8d %8d | %7d %7d | %.2e %.2e %.2e" % r)
    assert rows[-1][1] != rows[-1][2], "the 8-thread golden no longer differs from the serial one: regenerate the report in DESIGN.md"
