"""deterministic_time_prop and shift_cf_by_bias (cauchy_estimator.hpp:1331-1355, 1312-1328) against golden
vectors of the UNMODIFIED reference (oracle/ref_transforms.cpp -> tests/golden/*.transforms.mced): the term lists right after each
transform and the moments of every later step, bit for bit.  CPU: the emulated kernels; GPU: libmce_b200.so."""
import os

import numpy as np
import pytest

from compare import compare_dumps
from harness import ROOT, load_emu, load_product, run_transforms
from mceio import read_dump, read_scenario

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {"lti3": 4, "homing3": 4}        # scenario -> steps replayed before the transforms (homing3 adds the (T, B, u) form)


def _check(lib, name):
    gold = read_dump(os.path.join(GOLD, name + ".transforms.mced"))
    got = run_transforms(lib, read_scenario(os.path.join(GOLD, name + ".mces")), gold, CASES[name])
    want = {n: v for n, v in gold.items() if n not in ("header", "T", "bias")}
    assert any(n.startswith("t2/") for n in want) and any(n.endswith("/moments") for n in want)
    if name == "homing3":
        assert any(n.startswith("t3/") for n in want)
    probs = compare_dumps(want, {n: v for n, v in got.items() if n in want}, float_rtol=0.0)
    assert not probs, "\n".join(probs[:20])


@pytest.mark.parametrize("name", sorted(CASES))
def test_transforms_emulated(name):
    _check(load_emu(rebuild=False), name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_transforms_gpu(name):
    _check(load_product(), name)
