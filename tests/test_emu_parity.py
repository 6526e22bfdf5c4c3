"""CPU-only logic test of the kernel bodies: the sequential emulation backend (tests/emu, test infrastructure) runs
the same mce_kern_*.h code the GPU runs and must reproduce the reference's golden dumps bit for bit.  This does not
replace the `-m gpu` parity tests; it catches logic regressions on machines without a GPU."""
import os

import pytest

from compare import compare_dumps
from harness import ROOT, load_emu, run_scenario
from mceio import read_dump, read_scenario

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {"lti3": (8, 5), "lti2": (10, 6), "lti4": (6, 4), "lti3_3msmts": (12, 7), "lti4_2pnoise": (5, 3), "lti4_2msmts": (9, 5),
         "syn2": (12, 3), "syn3": (8, 3), "syn5": (5, 3), "syn7": (4, 2), "syn8": (4, 2), "leo7": (6, 3), "leo5": (7, 4), "homing3": (8, 5),
         # declared deeper than replayed: max_shape > 16 routes through KTpDce / KGTable (sort + hash variants)
         "homing_real": (8, 5), "leo7_deep16": (6, 3), "lti3_deep": (8, 4), "lti4_2pnoise_deep": (5, 3), "lti3_3msmts_deep": (12, 5)}


def _upto(d, k):
    return {n: v for n, v in d.items() if n == "header" or int(n.split("/")[0][1:]) <= k}


@pytest.fixture(scope="module")
def emu():
    return load_emu(rebuild=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_kernels_match_golden(emu, name):
    steps, full = CASES[name]
    sc = read_scenario(os.path.join(GOLD, name + ".mces"))
    gold = _upto(read_dump(os.path.join(GOLD, name + ".ref.mced")), steps)
    got = run_scenario(emu, sc, full_upto=full, max_steps=steps, capture=True)
    got = {n: v for n, v in got.items() if n in gold}
    probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12})
    assert not probs, "\n".join(probs[:20])


@pytest.mark.parametrize("name,steps,full,split", [("lti3", 8, 5, 3), ("lti2", 10, 6, 2), ("lti4_2msmts", 8, 5, 5), ("syn3", 8, 3, 4)])
def test_split_reduction_groups_match_golden(emu, name, steps, full, split):
    """Reduction groups above the split threshold take the multi-CTA path (root election / member parts / ordered sum);
    with a tiny threshold most groups do, and every array must still equal the reference's."""
    sc = read_scenario(os.path.join(GOLD, name + ".mces"))
    gold = _upto(read_dump(os.path.join(GOLD, name + ".ref.mced")), steps)
    got = run_scenario(emu, sc, full_upto=full, max_steps=steps, capture=True, split=split)
    assert max(got["s%d/stats" % k][12] for k in range(2, steps + 1)) > 0, "no group was split"
    got = {n: v for n, v in got.items() if n in gold}
    probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12})
    assert not probs, "\n".join(probs[:20])


@pytest.mark.parametrize("name,steps,full,split", [("lti3", 8, 5, 0), ("lti3", 8, 5, 3), ("leo7", 6, 3, 0), ("lti4_2msmts", 8, 5, 0), ("homing3", 8, 5, 0), ("syn2", 10, 3, 2)])
def test_lean_group_kernel_matches_golden(emu, name, steps, full, split):
    """The lean variant of the G-table kernel (one value table in shared memory instead of two; taken by the engine for tables the standard
    variant cannot hold) forced onto every group, with and without split groups."""
    sc = read_scenario(os.path.join(GOLD, name + ".mces"))
    gold = _upto(read_dump(os.path.join(GOLD, name + ".ref.mced")), steps)
    got = run_scenario(emu, sc, full_upto=full, max_steps=steps, capture=True, split=split, lean=True)
    assert max(got["s%d/lean/stats" % k][0] for k in range(2, steps + 1)) > 0, "the lean variant did not run"
    got = {n: v for n, v in got.items() if n in gold}
    probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12})
    assert not probs, "\n".join(probs[:20])


@pytest.mark.parametrize("name,steps,full", [("lti3", 8, 5), ("leo7", 6, 3), ("homing3", 8, 5), ("lti4_2pnoise", 5, 3), ("lti3_deep", 8, 4)])
def test_early_scale_from_the_exact_scan_matches_golden(emu, name, steps, full):
    """Large steps take G_SCALE_FACTOR from the exact scan of Re fz (KSumScan) and finish the serial moment chains beside the G-table kernels; with the
    threshold at one slot every step does.  The engine itself checks the scan against the chain bit for bit (MCE_ERR_SCAN)."""
    sc = read_scenario(os.path.join(GOLD, name + ".mces"))
    gold = _upto(read_dump(os.path.join(GOLD, name + ".ref.mced")), steps)
    got = run_scenario(emu, sc, full_upto=full, max_steps=steps, capture=True, early_scale=1)
    got = {n: v for n, v in got.items() if n in gold}
    probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12})
    assert not probs, "\n".join(probs[:20])


@pytest.mark.parametrize("name,steps,full", [("leo7", 6, 3), ("lti3", 8, 5)])
def test_fast_moments_keep_every_discrete_result(emu, name, steps, full):
    """mce_options.fast_moments: tree sums for every moment but Re fz, which the exact scan keeps bit-identical: counts, keys, hyperplanes and G values unchanged."""
    import numpy as np
    sc = read_scenario(os.path.join(GOLD, name + ".mces"))
    gold = _upto(read_dump(os.path.join(GOLD, name + ".ref.mced")), steps)
    got = run_scenario(emu, sc, full_upto=full, max_steps=steps, capture=True, fast_moments=2)       # threshold 2 slots: every step takes the tree + scan path
    got = {n: v for n, v in got.items() if n in gold}
    probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12}, skip=lambda n: n.endswith("/moments"))
    for n in gold:
        if n.endswith("/moments") and not n.startswith("s1/"):
            if gold[n][0].real.tobytes() != got[n][0].real.tobytes():
                probs.append("%s: Re fz is not bit-identical" % n)
            a, b, d = gold[n], got[n], sc.d
            if np.max(np.abs(a[1:1 + d] - b[1:1 + d])) > 1e-8 * np.max(np.abs(a[1:1 + d])) or np.max(np.abs(a[1 + d:] - b[1 + d:])) > 1e-4 * np.max(np.abs(a[1 + d:])):
                probs.append("%s: mean / covariance beyond the reordering noise" % n)
    assert not probs, "\n".join(probs[:20])
