#!/usr/bin/env python
"""ONE estimator partitioned over real GPUs (run under torchrun, one rank per GPU): every rank replays a scenario holding only
its own terms (NCCL send/recv, all-gather and all-reduce issued by the library, csrc/mce_kern_part.h); the merged results are
checked against the golden dump on every rank, and the window time is compared with the one-GPU run on rank 0.
Usage: torchrun --nproc-per-node N tools/shard_check.py [scenario] [steps] [full_upto] [ordered|allreduce]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist  # noqa: E402

from cauchyfriendly_b200.shard import init_term_sharding  # noqa: E402
from compare import compare_dumps  # noqa: E402
from harness import Session, load_product, run_scenario, run_scenario_partitioned  # noqa: E402
from cauchyfriendly_b200._capi import MceShardStats  # noqa: E402
import ctypes as ct  # noqa: E402
from mceio import SHIFT_EXPLICIT, read_dump, read_scenario  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "leo7"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
full = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = sys.argv[4] if len(sys.argv) > 4 else "ordered"
dist.init_process_group("gloo")                 # only ships the NCCL id and the final verdicts
rank, world, dev = dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
lib = load_product()
gold_dir = os.path.join(ROOT, "tests", "golden")
sc = read_scenario(os.path.join(gold_dir, name + ".mces"))
gold_path = os.path.join(gold_dir, name + ".ref.mced")
QUICK = bool(os.environ.get("SHARD_QUICK"))      # only the phase breakdown
if QUICK:
    gold, gold_kind = {}, "nothing"
elif os.path.exists(gold_path):
    gold = {n: v for n, v in read_dump(gold_path).items() if n == "header" or int(n.split("/")[0][1:]) <= steps}
    gold_kind = "the reference's golden dump"
else:             # no reference dump for this scenario (too deep for the CPU): the one-GPU run of this library, itself pinned on the shallower windows
    box = [None]
    if rank == 0:
        one = run_scenario(lib, sc, full_upto=0, max_steps=steps, device=dev)
        box[0] = {n: v for n, v in one.items() if not n.endswith("/stats")}
    dist.broadcast_object_list(box, src=0)
    gold = box[0]
    gold_kind = "the one-GPU run"

# 1. parity of the sharded run, on every rank
if not QUICK:
    xs = []
    def on_step(k, s, out):
        st = MceShardStats(); lib.mce_shard_get_stats(s.h, ct.byref(st))
        xs.append((k, st.local_parents, st.owned_terms, st.imported_parents, st.bytes_terms, st.bytes_parents, st.bytes_moments, st.bytes_keys))
    got = run_scenario_partitioned(lib, sc, dist, full_upto=full, max_steps=steps, transport="nccl", moments=mode, device=dev, on_step=on_step)
    got = {n: v for n, v in got.items() if n in gold}
    def skip(n):
        return "/muc/m" in n or ("/ftr/m" in n and n.split("/")[-1] in ("A", "p", "b", "cells", "keys", "G", "encB") and int(n.split("/")[0][1:]) > full)
    if mode == "ordered":
        probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12}, skip=skip)
    elif mode == "hybrid":      # everything bit-exact except Im fz / mean / covariance (rank-ordered partial sums); Re fz must be bit-exact
        import numpy as np
        probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12}, skip=lambda n: skip(n) or n.endswith("/moments"))
        worst = [0.0, 0.0]
        for n in gold:
            if n.endswith("/moments") and n in got:
                a, b = gold[n], got[n]
                if a[0].real.tobytes() != b[0].real.tobytes():
                    probs.append("%s: Re fz is not bit-identical" % n)
                worst[0] = max(worst[0], float(np.max(np.abs(a[1:1 + sc.d] - b[1:1 + sc.d])) / np.max(np.abs(a[1:1 + sc.d]))))
                worst[1] = max(worst[1], float(np.max(np.abs(a[1 + sc.d:] - b[1 + sc.d:])) / np.max(np.abs(a[1 + sc.d:]))))
        if rank == 0:
            print("hybrid moments: Re fz bit-identical on every step; largest deviation of the mean %.2e, of the covariance %.2e (relative to the largest entry)" % tuple(worst), flush=True)
    else:
        probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-6, r"/gscale$": 1e-12}, skip=lambda n: skip(n) or n.endswith("/G") or n.endswith("/moments"))
    allx = [None] * world
    dist.all_gather_object(allx, xs)
    if rank == 0:
        print("step: per rank (local parents, owned terms, imported parents, MB received: terms, parent tables, moment slots, keys)")
        for i in range(len(xs)):
            print("  %2d: " % xs[i][0] + " | ".join("%d %d %d %.1f %.1f %.1f %.1f" % (a[i][1], a[i][2], a[i][3], a[i][4] / 1e6, a[i][5] / 1e6, a[i][6] / 1e6, a[i][7] / 1e6) for a in allx), flush=True)
    # the one known gap of the one-GPU path itself (flattening.hpp:516-530 at MU 12-13 of the 15-MU LEO7 window)
    known = [q for q in probs if name == "leo7_w5" and (q.startswith("s12/ftr/m10/digest") or q.startswith("s13/ftr/m11/digest"))]
    probs = [q for q in probs if q not in known]
    verdicts = [None] * world
    dist.all_gather_object(verdicts, ("OK" + (" (apart from the %d key digests of the known aliasing gap, identical to the one-GPU result)" % len(known) if known else "")) if not probs else "; ".join(probs[:5]))
    if rank == 0:
        print("parity vs %s (%s, %d steps, %d ranks):" % (gold_kind, name, steps, world), verdicts, flush=True)


    # 2. window time: sharded (all ranks) vs unsharded (rank 0 alone)
    def window_ms(s, reps=3, sync=True):
        best, per = 1e30, None
        for _ in range(reps):
            lib.mce_reset(s.h)
            if sync:
                dist.barrier()
            ts = []
            t0 = time.perf_counter()
            for r in sc.rec[:steps]:
                t1 = time.perf_counter()
                s.step(r)
                if r.shift_kind == SHIFT_EXPLICIT:
                    s.shift_b(r.delta, -1.0)
                ts.append((time.perf_counter() - t1) * 1e3)
            tot = (time.perf_counter() - t0) * 1e3
            if tot < best:
                best, per = tot, ts
        return best, per


    s = Session(lib, sc, device=dev)
    init_term_sharding(s.h, dist, lib=lib, transport="nccl", device=dev, moments=mode)
    t_sh, per_sh = window_ms(s)
    s.close()
    ts = [None] * world
    dist.all_gather_object(ts, t_sh)
    if rank == 0:
        s1 = Session(lib, sc, device=dev)
        best, per1 = window_ms(s1, sync=False)
        s1.close()
        print("ms per MU, one GPU:      " + " ".join("%.2f" % v for v in per1))
        print("ms per MU, %d ranks (r0): " % world + " ".join("%.2f" % v for v in per_sh), flush=True)
        print("window time (%s moments): %d-rank partitioned estimator %.2f ms (max over ranks), single GPU %.2f ms, speed-up %.2fx" % (mode, world, max(ts), best, best / max(ts)), flush=True)
# 3. where the time goes: one more pass with a stream synchronisation after every phase (mce_options.phase_timing)
s = Session(lib, sc, device=dev, phase_timing=True)
init_term_sharding(s.h, dist, lib=lib, transport="nccl", device=dev, moments=mode)
rows = []
for rep in range(2):
    lib.mce_reset(s.h)
    dist.barrier()
    rows = []
    for r in sc.rec[:steps]:
        s.step(r)
        st = s.stats()
        ss = MceShardStats(); lib.mce_shard_get_stats(s.h, ct.byref(ss))
        rows.append((st.ms_total, st.ms_tp, st.ms_mu, st.ms_moments, st.ms_regroup, st.ms_ftr, st.ms_gtable, st.ms_compact, st.ev_moments_ms) + tuple(ss.ms_stage))
        if r.shift_kind == SHIFT_EXPLICIT:
            s.shift_b(r.delta, -1.0)
s.close()
allr = [None] * world
dist.all_gather_object(allr, rows)
if rank == 0:
    print("phase times (ms, max over ranks; phase timing on): MU total | tp mu moments(+gather) regroup+exchange ftr gtable(+mask sync) compact(+rank assign) | moment chain")
    for k in range(len(rows)):
        mx = [max(a[k][i] for a in allr) for i in range(17)]
        print("  %2d %7.2f | %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f | %6.2f | exchange stages: keys %.2f dest %.2f terms %.2f unpack %.2f implist %.2f req %.2f parents %.2f store %.2f" % tuple([k + 1] + mx), flush=True)
dist.barrier()
dist.destroy_process_group()
