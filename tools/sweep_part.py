#!/usr/bin/env python
"""Synthetic LTI term-count sweep at scale (BASELINE.json configs[4]), one GPU or ONE estimator
partitioned over N GPUs (run under torchrun for N > 1).  For every state dimension n:
  1. parity: the first K steps (what the CPU reference finishes in about a minute) are replayed with full term-list merging
     and compared bit for bit -- counts, key digests, moments -- with oracle/_ref/ref_run_cpu1 run on the same scenario file;
  2. scale: the window is stepped until the next step would exceed the term cap; per-step terms and times are printed, and
     the counts / moments of the N-GPU run are compared with the one-GPU run of rank 0 on the steps both reach.
Usage: [torchrun --nproc-per-node N] tools/sweep_part.py [--cap 5e7] [--cap1 1.5e7] [--dims 3 4] [--moments ordered|allreduce] [--cpu-steps K] [--emu]"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from compare import compare_dumps  # noqa: E402
from gen_scenarios import lti  # noqa: E402
from harness import Session, load_emu, load_product, run_scenario, run_scenario_partitioned  # noqa: E402
from mceio import read_dump, write_scenario  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cap", type=float, default=5e7, help="term cap per step of the (partitioned) run")
ap.add_argument("--cap1", type=float, default=1.5e7, help="term cap per step of the one-GPU comparison run")
ap.add_argument("--dims", dest="n", type=int, nargs="*", default=[3])
ap.add_argument("--moments", default="ordered")
ap.add_argument("--cpu-steps", type=int, default=9)
ap.add_argument("--emu", action="store_true", help="CPU smoke test: emulated kernels + gloo callback transport")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = int(os.environ.get("LOCAL_RANK", "0"))
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")          # ships the NCCL id, the verdicts and (parity phase) the merged term lists
lib = load_emu() if args.emu else load_product()
transport = "callback" if args.emu else "nccl"
os.makedirs(args.out, exist_ok=True)


def system(n):
    steps = 17 - n                                   # max_shape = steps - 1 + n <= 16
    rng = np.random.RandomState(1000 + n)
    Phi = rng.uniform(-1, 1, (n, n)); Gam = rng.uniform(-1, 1, n); H = rng.uniform(-1, 1, n)
    Phi *= 0.95 / np.max(np.abs(np.linalg.eigvals(Phi)))
    x = np.zeros(n); zs = []
    for _ in range(steps):
        x = Phi @ x + Gam * 0.1 * rng.standard_cauchy(); zs.append(H @ x + 0.2 * rng.standard_cauchy())
    return lti("syn%d_deep" % n, Phi, Gam, H, [0.1], [0.2], np.eye(n), np.full(n, .1), np.zeros(n), zs, steps, seed=100 + n)[1]


def timed_run(sc, cap, partitioned):
    """Steps the window until the next step would exceed `cap` terms; returns per-step rows (step, global terms after MUC, survivors,
    ms (max over ranks when partitioned), fz / mean bits)."""
    s = Session(lib, sc, device=dev)
    if partitioned:
        from cauchyfriendly_b200.shard import init_term_sharding
        init_term_sharding(s.h, dist, lib=lib, transport=transport, device=dev, moments=args.moments)
    rows = []
    try:
        for rep in range(2):                         # pass 0 sizes the buffers
            lib.mce_reset(s.h)
            rows = []
            if partitioned:
                dist.barrier()
            for k, r in enumerate(sc.rec):
                t0 = time.perf_counter()
                s.step(r)
                dt = (time.perf_counter() - t0) * 1e3
                mo = s.moments()
                st = s.stats()
                rows.append((k + 1, int(st.terms_after_muc), int(mo.Nt), dt, (mo.fz_after_mu[0], tuple(mo.mean[:2 * sc.d])), tuple(s.counts(False).tolist())))
                if st.terms_after_muc * 3.6 > cap or k + 2 >= len(sc.rec):      # the next step would exceed the cap (the last step only sums moments)
                    break
    finally:
        s.close()
    if partitioned:
        ms = [None] * world
        dist.all_gather_object(ms, [r[3] for r in rows])
        rows = [r[:3] + (max(m[i] for m in ms),) + r[4:] for i, r in enumerate(rows)]
    return rows


summary = []
for n in args.n:
    sc = system(n)
    scen_path = os.path.join(args.out, "sweep_syn%d_deep.mces" % n)
    if rank == 0:
        write_scenario(scen_path, sc)
    # ---- 1. parity against the unmodified reference on the steps the CPU reaches ----
    K = min(args.cpu_steps, len(sc.rec) - 1)
    verdict = "skipped"
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "ref_run_cpu1")
    if K > 0 and os.path.exists(ref_bin):
        if world > 1:
            got = run_scenario_partitioned(lib, sc, dist, full_upto=0, max_steps=K, transport=transport, moments="ordered", device=dev)
        else:
            got = run_scenario(lib, sc, full_upto=0, max_steps=K, device=dev)
        if rank == 0:
            dump = os.path.join(args.out, "sweep_syn%d_deep.ref.mced" % n)
            t0 = time.time()
            subprocess.check_call([ref_bin, scen_path, dump, "--full-upto", "0", "--max-steps", str(K)], stdout=subprocess.DEVNULL)
            gold = read_dump(dump)
            got = {k: v for k, v in got.items() if k in gold}
            probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12}, skip=lambda nm: "/muc/m" in nm or nm.endswith("/stats"))
            verdict = "bit-exact (counts, key digests, moments; %d arrays)" % len(got) if not probs else "DIFFERS: " + "; ".join(probs[:3])
            print("n=%d parity vs ref_run_cpu1 on steps 1..%d (%d rank%s, reference took %.1f s): %s" % (n, K, world, "s" if world > 1 else "", time.time() - t0, verdict), flush=True)
    # ---- 2. scale ----
    rows = timed_run(sc, args.cap, world > 1)
    rows1 = timed_run(sc, args.cap1, False) if (world > 1 and rank == 0) else None
    if rank == 0:
        for r in rows:
            one = next((q for q in rows1 if q[0] == r[0]), None) if rows1 else None
            agree = "" if one is None else ("  one GPU: %8.2f ms, counts %s, fz/mean %s" % (one[3], "equal" if one[5] == r[5] and one[1] == r[1] else "DIFFER", "bit-equal" if one[4] == r[4] else "differ"))
            print("n=%d step %2d: terms after MUC %10d  survivors %9d  %9.2f ms%s" % (n, r[0], r[1], r[2], r[3], agree), flush=True)
        child = sum(r[1] for r in rows) - sum(r[2] for r in rows[:-1]) - 1
        ms = sum(r[3] for r in rows)
        entry = {"n": n, "gpus": world, "moments": args.moments if world > 1 else "ordered", "steps_run": rows[-1][0], "child_terms": int(child), "ms": ms,
                 "child_terms_per_s": child / (ms * 1e-3), "largest_step_terms": rows[-1][1], "largest_step_ms": rows[-1][3], "parity_vs_reference": verdict}
        if rows1:
            common = [r for r in rows if any(q[0] == r[0] for q in rows1)]
            ms_n = sum(r[3] for r in common); ms_1 = sum(q[3] for q in rows1 if any(r[0] == q[0] for r in common))
            entry["one_gpu_ms_same_steps"] = ms_1; entry["ms_same_steps"] = ms_n; entry["speedup_same_steps"] = ms_1 / ms_n
            entry["counts_equal_one_gpu"] = all(q[5] == r[5] and q[1] == r[1] for r in common for q in rows1 if q[0] == r[0])
        summary.append(entry)
        print("n=%d: %d steps, %d child terms in %.1f ms -> %.2f M child terms/s on %d GPU(s)" % (n, rows[-1][0], child, ms, child / ms / 1e3, world), flush=True)
if rank == 0:
    print(json.dumps({"sweep": summary, "term_cap": args.cap, "gpus": world}))
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
