"""Pins the fp64 complex helpers of csrc/mce_math.h (restatements of libgcc __divdc3 / __muldc3 and glibc hypot) against
the host libraries the reference binary calls: 4 M random and edge-case operands, bit for bit."""
import os
import subprocess

from harness import ROOT


def test_complex_div_mul_abs_match_libgcc_glibc(tmp_path):
    exe = str(tmp_path / "tmath")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "cauchyfriendly_b200", "csrc"),
                           os.path.join(ROOT, "tests", "emu", "test_math.cpp"), "-o", exe, "-lm"])
    out = subprocess.check_output([exe]).decode()
    assert "bad_div=0 bad_mul=0 bad_abs=0" in out, out
