"""SURVEY section 8f, relative-system / prediction readers (cauchy_prediction.hpp:207-1190), reached through the Swig shim exactly as the reference's Python
and MATLAB bindings reach them (pycauchy_single_step_eval_2d_rsys_cpdf, pycauchy.hpp:677).  They are O(Nt_primary x Nt_secondary) readers of two term
lists OUTSIDE step(): behind the drop-in header they run unchanged over the host mirror of the device-resident term lists (MCE_AUTO_MIRROR=1).
tests/dropin/rsys_dropin.cpp is compiled twice -- unmodified reference (golden tests/golden/ex_rsys_cpu1.txt, tools/make_golden_dropin.py) and drop-in
header + library -- and every result line (estimator moments, relative-system normaliser / mean / covariance, all grid values) must agree digit for digit."""
import os
import subprocess

import pytest

from harness import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden", "ex_rsys_cpu1.txt")


def result_lines(text):
    """Everything the program prints itself; the reference's own progress / timing chatter is dropped."""
    return [l for l in text.splitlines() if l.startswith(("#", "  z ", "  rsys", "  pt", "rsys drop-in"))]


def _run(exe, cwd):
    env = dict(os.environ)
    env["MCE_AUTO_MIRROR"] = "1"
    return subprocess.run([exe], cwd=cwd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600).stdout.decode(errors="replace")


def _check(text):
    gold = open(GOLDEN).read().splitlines()
    ours = result_lines(text)
    assert len(gold) == 221 and gold[-1] == "rsys drop-in done"
    assert len(ours) == len(gold), text[-1500:]
    bad = [(g, o) for g, o in zip(gold, ours) if g != o]
    assert not bad, "first difference:\n  reference %s\n  drop-in   %s" % bad[0]


def test_relative_system_readers_over_the_emulated_kernels(tmp_path):
    """CPU twin: the same drop-in program linked with tests/emu (the kernel bodies run sequentially).  Needs the reference headers at build time."""
    ov = os.path.join(ROOT, "build", "overlay_cpu1")
    if not os.path.isdir("/root/reference/include"):
        pytest.skip("reference headers absent (the program is the reference's pycauchy.hpp compiled against the drop-in header)")
    from harness import load_emu
    load_emu()
    if not os.path.isdir(ov):
        subprocess.check_call(["bash", os.path.join(ROOT, "tools", "build_dropin.sh")], stdout=subprocess.DEVNULL)
    emu_dir = os.path.join(ROOT, "tests", "emu", "_build")
    exe = str(tmp_path / "rsys_emu")
    subprocess.check_call(["g++", "-O2", "-w", "-ffp-contract=off", "-I" + os.path.join(ov, "include"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ov, "tests", "rsys_dropin.cpp"), "-o", exe, "-L" + emu_dir, "-lmce_emu", "-Wl,-rpath," + emu_dir, "-lm", "-lpthread"])
    _check(_run(exe, str(tmp_path)))


@pytest.mark.gpu
def test_relative_system_readers_on_the_gpu_path(tmp_path):
    exe = os.path.join(ROOT, "build", "dropin", "rsys_dropin")
    if not os.path.exists(exe):
        pytest.skip("drop-in binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    _check(_run(exe, str(tmp_path)))
