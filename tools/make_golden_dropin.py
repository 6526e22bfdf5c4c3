#!/usr/bin/env python
"""Golden outputs of the reference's own host-side callers of the estimator, produced by the UNMODIFIED reference built with
NUM_CPUS = 1 (oracle/Makefile -> oracle/_ref/ex_*_cpu1; needs /root/reference at build time).  tests/test_gpu_dropin.py runs the
same sources compiled against include/cauchy_estimator.hpp + libmce_b200.so (tools/build_dropin.sh) and compares.

  ex_pycauchy_cpu1.txt   stdout of tests/dropin/pycauchy_dropin.cpp (the Swig shim pycauchy.hpp driven from C++)
  ex_leo5_cpu1.txt       counts + moments printed by src/leo_satellite_5state.cpp (BASELINE.json configs[2])
  ex_rsys_cpu1.txt       the result lines of tests/dropin/rsys_dropin.cpp: the relative-system readers of cauchy_prediction.hpp over two estimators
  winbank_cpu1/          log files of the 8-window bank on the inputs of src/window_manager.cpp (tests/dropin/winbank_dropin.cpp)
  homing_cpu1/           log files of src/homing_missile.cpp (BASELINE.json configs[1]; closed loop, 8 windows, 99 steps) with
                         time() pinned to the author's seed 1658778374 by tests/dropin/fixed_time.c (LD_PRELOAD)
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD = os.path.join(ROOT, "tests", "golden")
BANK_FILES = ["cond_means.txt", "cond_covars.txt", "norm_factors.txt", "numeric_error_codes.txt", "cerr_cond_means.txt", "cerr_cond_covars.txt", "cerr_norm_factors.txt"]
HOMING_FILES = BANK_FILES + ["cauchy_controls.txt", "cauchy_with_controller_msmts.txt", "cauchy_with_controller_true_states.txt"]


def run(cmd, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, check=True).stdout.decode(errors="replace")


def parsed_example(text):
    from test_gpu_dropin import _parse
    counts, moments = _parse(text)
    return "counts " + " ".join(counts) + "\n" + "".join("moment %s | %s | %s\n" % m for m in moments)


def main():
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(GOLD, "ex_pycauchy_cpu1.txt"), "w").write(run([os.path.join(REF, "ex_pycauchy_cpu1")], td))
        from test_rsys_dropin import result_lines
        open(os.path.join(GOLD, "ex_rsys_cpu1.txt"), "w").write("\n".join(result_lines(run([os.path.join(REF, "ex_rsys_cpu1")], td))) + "\n")
        wb = os.path.join(td, "wb")
        os.makedirs(wb)
        run([os.path.join(REF, "ex_winbank_cpu1"), wb], td)
        os.makedirs(os.path.join(GOLD, "winbank_cpu1"), exist_ok=True)
        for f in BANK_FILES + ["msmts.txt", "root_point_b_pert.txt"]:
            shutil.copy(os.path.join(wb, f), os.path.join(GOLD, "winbank_cpu1", f))
        hm = os.path.join(td, "hm")
        os.makedirs(hm)
        run([os.path.join(REF, "ex_homing_cpu1"), "8", hm, "5.0", "1.3", "1"], td, {"LD_PRELOAD": os.path.join(REF, "fixed_time.so")})
        os.makedirs(os.path.join(GOLD, "homing_cpu1"), exist_ok=True)
        for f in HOMING_FILES:
            shutil.copy(os.path.join(hm, "w8_bs5_sas13", "mct1", f), os.path.join(GOLD, "homing_cpu1", f))
        if "--leo5" in sys.argv:        # 2.5 CPU-minutes
            os.makedirs(os.path.join(td, "l5", "bin"))
            os.makedirs(os.path.join(td, "l5", "log", "leo5", "dense", "w8"))
            open(os.path.join(GOLD, "ex_leo5_cpu1.txt"), "w").write(parsed_example(run([os.path.join(REF, "ex_leo5_cpu1")], os.path.join(td, "l5", "bin"))))


if __name__ == "__main__":
    main()
