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
         "lti3_deep": 9, "lti4_2pnoise_deep": 6, "leo5_deep": 8, "lti3_3msmts_deep": 12}


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
