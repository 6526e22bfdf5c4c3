"""Drop-in check: the reference's OWN example programs (src/cauchy_estimator.cpp, src/leo_satellite_7state_gps.cpp),
compiled unchanged against include/cauchy_estimator.hpp + libmce_b200.so (tools/build_dropin.sh), print the same term
counts and moments as the same programs built from the unmodified reference (oracle/_ref/ex_*_cpu1)."""
import os
import re
import subprocess

import pytest

from harness import ROOT

pytestmark = pytest.mark.gpu


def _parse(text):
    """[(terms after MUC/MU, fz string, mean row string)] for every 'after MU' moment block."""
    out = []
    blocks = re.split(r"Moment Information \(after MU\)", text)
    counts = re.findall(r"Total Terms after MUC?: (\d+)", text)
    for b in blocks[1:]:
        fz = re.search(r"fz: (\S+) \+ (\S+)j", b)
        mean = re.search(r"Conditional Mean:\s*\n([^\n]*)\n", b)
        out.append((fz.group(1), fz.group(2), mean.group(1).strip() if mean else ""))
    return counts, out


def _run(exe, cwd):
    return subprocess.run([exe], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500).stdout.decode(errors="replace")


def _golden(name):
    counts, moments = [], []
    for line in open(os.path.join(ROOT, "tests", "golden", name + ".txt")):
        if line.startswith("counts "):
            counts = line.split()[1:]
        elif line.startswith("moment "):
            a, b, c = line[len("moment "):].rstrip("\n").split(" | ")
            moments.append((a, b, c))
    return counts, moments


@pytest.mark.parametrize("ref_name,our_name", [("ex_cauchy_estimator_cpu1", "cauchy_estimator"), ("ex_leo7_cpu1", "leo_satellite_7state_gps")])
def test_reference_example_runs_on_the_gpu_path(ref_name, our_name, tmp_path):
    """The golden text files hold the parsed output of the unmodified reference example (NUM_CPUS=1 build), generated in
    the build container (the 7-state example needs ~4 CPU-minutes there); our binary is the same example source compiled
    against the drop-in header.  Both examples print with print_basic_info = true, i.e. they also exercise quirk A.9(iii)."""
    our_exe = os.path.join(ROOT, "build", "dropin", our_name)
    if not os.path.exists(our_exe):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    work = tmp_path / "bin"
    work.mkdir()
    (tmp_path / "log" / "leo7" / "dense" / "w5").mkdir(parents=True)
    ours = _run(our_exe, str(work))
    c_ref, m_ref = _golden(ref_name)
    c_our, m_our = _parse(ours)
    assert len(m_ref) >= 10
    assert c_our == c_ref, "term counts differ:\n%s\n%s\n%s" % (c_ref, c_our, ours[-1500:])
    assert m_our == m_ref


def test_reference_window_manager_runs_on_the_gpu_path(tmp_path):
    """src/window_manager.cpp unchanged: SlidingWindowManager forks 8 window processes, each creating its own estimator
    (and CUDA context) through the drop-in header; 201 measurements of the 3-state system must go through."""
    exe = os.path.join(ROOT, "build", "dropin", "window_manager")
    if not os.path.exists(exe):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    out = _run(exe, str(tmp_path))
    m = re.search(r"The Simulation of (\d+) measurements took (\S+) seconds; rate = (\S+) hz", out)
    assert m, out[-2000:]
    assert int(m.group(1)) == 201 and float(m.group(3)) > 0
    assert "All children have exited" in out


def test_device_cpdf_dispatcher_equals_the_reference_cpu_cpdf(tmp_path):
    """tests/dropin/cpdf1d_dropin.cpp: the reference's own CauchyCPDFGridDispatcher1D (cpdf_ndim.hpp, compiled unchanged) reads
    the host mirror of the device term list and must produce the same 401-point grids, bit for bit, as the device
    dispatcher of include/cpdf_b200.hpp, for every state after each of 7 steps of the 3-state example."""
    exe = os.path.join(ROOT, "build", "dropin", "cpdf1d_dropin")
    if not os.path.exists(exe):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    out = _run(exe, str(tmp_path))
    assert "cpdf1d drop-in OK" in out, out[-2000:]
    m = re.search(r"compared (\d+) grid values, (\d+) differ", out)
    assert m and int(m.group(1)) == 7 * 3 * 401 and int(m.group(2)) == 0


def test_leo5_example_runs_on_the_gpu_path(tmp_path):
    """src/leo_satellite_5state.cpp unchanged (BASELINE.json configs[2]: 5-state LEO EMCE, 14 measurement updates, closed loop
    through finalize_extended_moments): same counts and moments as the NUM_CPUS=1 reference (tests/golden/ex_leo5_cpu1.txt)."""
    exe = os.path.join(ROOT, "build", "dropin", "leo_satellite_5state")
    if not os.path.exists(exe):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    work = tmp_path / "bin"
    work.mkdir()
    (tmp_path / "log" / "leo5" / "dense" / "w8").mkdir(parents=True)
    c_our, m_our = _parse(_run(exe, str(work)))
    c_ref, m_ref = _golden("ex_leo5_cpu1")
    assert len(m_ref) == 14
    assert c_our == c_ref
    assert m_our == m_ref


def test_swig_shim_reset_paths_match_the_reference(tmp_path):
    """tests/dropin/pycauchy_dropin.cpp: the reference's pycauchy.hpp (what the Swig module and the mex files call) driven from a
    C++ main -- steps, pycauchy_single_step_reset in both of its branches (reset() re-seeding from A0_init written in place;
    setup_first_term at master_step == 0), deterministic transforms before and after the first step.  The text must equal the
    output of the same program built from the unmodified reference (tests/golden/ex_pycauchy_cpu1.txt), digit for digit."""
    exe = os.path.join(ROOT, "build", "dropin", "pycauchy_dropin")
    if not os.path.exists(exe):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    ours = [l for l in _run(exe, str(tmp_path)).splitlines() if l.startswith(("#", "z ", "  xhat", "  Phat", "  cerr", "pycauchy"))]
    gold = open(os.path.join(ROOT, "tests", "golden", "ex_pycauchy_cpu1.txt")).read().splitlines()
    assert len(gold) == 78 and gold[-1] == "pycauchy drop-in done"
    diff = [(i, a, b) for i, (a, b) in enumerate(zip(gold, ours)) if a != b]
    assert len(ours) == len(gold) and not diff, "first differing lines:\n" + "\n".join("%d\n  ref %s\n  got %s" % t for t in diff[:5])


def _bank_logs_equal(gold_dir, got_dir, files):
    bad = []
    for f in files:
        a = open(os.path.join(gold_dir, f)).read().split()
        b = open(os.path.join(got_dir, f)).read().split()
        if a != b:
            n = sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
            first = next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
            bad.append("%s: %d of %d tokens differ (%d vs %d tokens); first at token %d: ref %s, got %s" % (
                f, n, len(a), len(a), len(b), first, a[first:first + 3], b[first:first + 3]))
    return bad


def test_reference_window_bank_logs_match_the_reference(tmp_path):
    """The reference's SlidingWindowManager (8 forked windows, each with its own estimator and CUDA context) on the inputs of
    src/window_manager.cpp (srand(11), 201 measurements), logging enabled (tests/dropin/winbank_dropin.cpp): the bank's
    conditional means / covariances / normalisation factors / error codes -- printed with 16 decimals -- must equal the logs
    of the same program built from the unmodified NUM_CPUS=1 reference (tests/golden/winbank_cpu1/)."""
    exe = os.path.join(ROOT, "build", "dropin", "winbank_dropin")
    if not os.path.exists(exe):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    logs = tmp_path / "logs"
    logs.mkdir()
    out = subprocess.run([exe, str(logs)], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500).stdout.decode(errors="replace")
    assert "winbank drop-in done: 201 measurements" in out, out[-2000:]
    files = ["cond_means.txt", "cond_covars.txt", "norm_factors.txt", "numeric_error_codes.txt", "cerr_cond_means.txt", "cerr_cond_covars.txt", "cerr_norm_factors.txt"]
    bad = _bank_logs_equal(os.path.join(ROOT, "tests", "golden", "winbank_cpu1"), str(logs), files)
    assert not bad, "\n".join(bad)


def test_homing_missile_example_matches_the_reference(tmp_path):
    """src/homing_missile.cpp unchanged (BASELINE.json configs[1]: 3-state nonlinear EMCE with a control input, 8 windows, 99
    steps, CLOSED LOOP -- the guidance command is computed from the bank's estimate, so any deviation feeds back).  time() is
    pinned to the author's seed by an LD_PRELOAD shim (tests/dropin/fixed_time.c).  Controls, measurements and every bank log
    must equal those of the unmodified NUM_CPUS=1 reference (tests/golden/homing_cpu1/)."""
    exe = os.path.join(ROOT, "build", "dropin", "homing_missile")
    shim = os.path.join(ROOT, "build", "dropin", "fixed_time.so")
    if not os.path.exists(exe) or not os.path.exists(shim):
        pytest.skip("drop-in example binaries not built (needs /root/reference at build time: tools/build_dropin.sh)")
    logs = tmp_path / "hm"
    logs.mkdir()
    env = dict(os.environ, LD_PRELOAD=shim)
    out = subprocess.run([exe, "8", str(logs), "5.0", "1.3", "1"], cwd=str(tmp_path), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=1500).stdout.decode(errors="replace")
    assert "Seeding with 1658778374" in out and "All children have exited" in out, out[-2000:]
    files = ["cond_means.txt", "cond_covars.txt", "norm_factors.txt", "numeric_error_codes.txt", "cerr_cond_means.txt", "cerr_cond_covars.txt", "cerr_norm_factors.txt",
             "cauchy_controls.txt", "cauchy_with_controller_msmts.txt", "cauchy_with_controller_true_states.txt"]
    bad = _bank_logs_equal(os.path.join(ROOT, "tests", "golden", "homing_cpu1"), str(logs / "w8_bs5_sas13" / "mct1"), files)
    if bad:
        print("\n".join(bad))
    assert not bad, bad[0][:300]
