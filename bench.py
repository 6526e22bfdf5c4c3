#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 MCE term-propagation path.

Metric (BASELINE.json): child terms/s (and ms per MCE step) on the 7-state LEO GPS sliding-window estimator
(configs[3]: d=7, 3 GPS measurements per time step, 4 time steps = 12 measurement updates, replayed open-loop
from tests/golden/leo7.mces, which records the reference example src/leo_satellite_7state_gps.cpp).

One bench *step* = one full pass of that 12-MU window through CauchyEstimator::step() on a fresh estimator
(mce_reset between passes): 2.0 M child terms are generated and globally de-duplicated per pass.
  value  : child terms / s, device time (CUDA events on the engine's stream, summed over the 12 step() calls)
  e2e    : same metric, wall clock around the reference-facing C-ABI calls with HOST buffers (every step() call
           uploads Phi/Gamma/H/... and downloads the moment sums; nothing is cached between passes)
  N > 1  : ONE window partitioned over the GPUs (every term lives on one rank: terms are routed to the owners of their
           reduction keys over NCCL send/recv, parent tables are fetched from their home ranks, the moment sums run in the
           reference's order on every rank -- csrc/mce_kern_part.h), scaling = "strong".  The window is the example's
           sliding-window depth (num_windows = 5, leo_satellite_7state_gps.cpp:585: 15 MUs, 17 M child terms) so that there is
           something to partition; the line also carries the same window on ONE GPU measured in the same run
           (`one_gpu_same_window`) and the round-1 figure of N independent 12-MU windows (`replicas`).
           --shard windows restores the replica run as the main line (scaling = "weak").
--impl reference times the reference's own CPU implementation (oracle/_ref/ref_run_cpu8, the unmodified
reference compiled with its default NUM_CPUS = 8) on the SAME full window (one pass; `config` is identical in both
arms) and, for repeatability, on the prefix MUs 1..9, which both arms report as `matched`.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SCEN = os.path.join(ROOT, "tests", "golden", "leo7.mces")
WORKLOAD = "leo7_gps_d7_p3_window4 (12 MUs, tests/golden/leo7.mces)"
SCEN_DEEP = os.path.join(ROOT, "tests", "golden", "leo7_w5.mces")
WORKLOAD_DEEP = "leo7_gps_d7_p3_window5 (15 MUs, tests/golden/leo7_w5.mces)"
METRIC = "child_terms_per_sec"


def _workload(args):
    """N = 1: the 12-MU window of the reference's own example run; N > 1 (one window partitioned): its 15-MU sliding-window depth."""
    deep = args.window == 5 or (args.window == 0 and args.gpus > 1 and args.shard == "terms")
    return (SCEN_DEEP, WORKLOAD_DEEP) if deep else (SCEN, WORKLOAD)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                              timeout=5).decode().strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def _parse_ref_lines(text):
    rows = []
    for m in re.finditer(r"step (\d+): after MUC (\d+), after FTR (\d+), ([0-9.]+) ms", text):
        rows.append((int(m.group(1)), int(m.group(2)), int(m.group(3)), float(m.group(4))))
    return rows


def _child_terms(rows):
    """child terms of MU k = terms after MUC - terms that entered the step (previous after-FTR count)."""
    tot, prev = 0, 1
    for _, muc, ftr, _ in rows:
        tot += muc - prev
        prev = ftr
    return tot


def run_reference_sample(n_mu, scen=None):
    scen = scen or SCEN
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_run_cpu8")
    kind = "reference"
    if not os.path.exists(exe):
        exe = os.path.join(ROOT, "oracle", "_build", "mce_oracle_run")
        kind = "port"
        if not os.path.exists(exe):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    t0 = time.time()
    out = subprocess.check_output([exe, scen, "--time-only", "--max-steps", str(n_mu)], stderr=subprocess.DEVNULL).decode()
    wall = time.time() - t0
    rows = _parse_ref_lines(out)
    step_ms = sum(r[3] for r in rows)
    return kind, rows, step_ms, wall


MATCHED_MUS = 9          # prefix of the window both arms also report on its own (cheap enough to repeat on the CPU)


def _config(n_mus, workload=WORKLOAD):
    """Identical in both arms: the workload (inputs) both time.  Everything arm-specific lives in other keys of the line --
    including the child-term count: the reference's 8-thread build walks the terms in another order than its 1-thread build
    (the canonical order this repository reproduces bit for bit), elects other reduction-group roots and from MU 10 on carries
    ~13 % more terms through the same window (2 402 755 child terms against 2 115 431)."""
    return {"workload": workload, "mus_per_step": n_mus}


def reference_arm(args):
    """The reference's own CPU implementation (unmodified, NUM_CPUS = 8 pthreads) on this box's host cores.
    `value` is measured on the SAME work the GPU arm times: the full 12-MU window, run --ref-full-passes times (default 1,
    about half a minute).  The remaining steps / warm-ups are the bounded sample MUs 1..9 of the same window, reported as
    `matched` (the GPU arm reports the same prefix from its per-MU CUDA events), so that the K + W passes end within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scen, workload = _workload(args)
    if scen == SCEN_DEEP:
        return reference_arm_deep(args, scen, workload)
    for _ in range(min(1, args.warmup)):
        run_reference_sample(MATCHED_MUS)
    full = []
    for _ in range(max(1, args.ref_full_passes)):
        kind, rows, ms, _ = run_reference_sample(12)
        full.append((rows, ms))
    pre_ms, pre_child, n_pre = 0.0, 0, 0
    budget_s = 120.0
    t0 = time.time()
    for _ in range(max(0, args.steps - len(full))):
        if time.time() - t0 > budget_s:
            break
        kind, rows, ms, _ = run_reference_sample(MATCHED_MUS)
        pre_ms += ms
        pre_child += _child_terms(rows)
        n_pre += 1
    if n_pre == 0:          # the prefix of a full pass is the same computation
        rows, _ = full[0]
        pre_ms = sum(r[3] for r in rows[:MATCHED_MUS]); pre_child = _child_terms(rows[:MATCHED_MUS]); n_pre = 1
    tot_ms = sum(ms for _, ms in full)
    tot_child = sum(_child_terms(rows) for rows, _ in full)
    value = tot_child / (tot_ms / 1e3)
    cores = 8 if kind == "reference" else 1
    who = "unmodified reference, NUM_CPUS=8 pthreads" if kind == "reference" else "plain-C oracle port, 1 thread"
    sample = "the full 12-MU window, %d pass(es) (%d child terms, %.1f s of step() time each); %d further passes of MUs 1..%d as `matched`; %s" % (
        len(full), tot_child // len(full), tot_ms / 1e3 / len(full), n_pre, MATCHED_MUS, who)
    heaviest = max(full[0][0], key=lambda r: r[3])
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "child terms/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": tot_ms / len(full), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": _config(12), "child_terms_per_step": tot_child // len(full),
            "matched": {"mus": MATCHED_MUS, "child_terms": pre_child // n_pre, "value": pre_child / (pre_ms / 1e3), "ms": pre_ms / n_pre, "passes": n_pre},
            "detail": {"full_window_passes": len(full), "heaviest_mu": {"mu": heaviest[0], "ms": heaviest[3], "terms_after_muc": heaviest[1], "survivors": heaviest[2]},
                       "ms_per_mu": [r[3] for r in full[0][0]]},
            "cpu_baseline": {"value": value, "unit": "child terms/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "child terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def reference_arm_deep(args, scen, workload):
    """N > 1: the GPU arm partitions the 15-MU window.  One pass of that window takes the reference ~7 minutes on 8 threads (its
    throughput falls with depth), so each step is the bounded sample MUs 1..12 of the same window (about 20 s); `value` is the
    reference's child terms/s on that sample -- an UPPER bound of its throughput on the whole window."""
    n_mu = 12
    passes = []
    t0 = time.time()
    for _ in range(max(1, args.steps)):
        kind, rows, ms, _ = run_reference_sample(n_mu, scen)
        passes.append((rows, ms))
        if time.time() - t0 > 150.0:
            break
    tot_ms = sum(ms for _, ms in passes); tot_child = sum(_child_terms(rows) for rows, _ in passes)
    value = tot_child / (tot_ms / 1e3)
    cores = 8 if kind == "reference" else 1
    sample = "MUs 1..%d of the 15-MU window, %d pass(es) (%d child terms, %.1f s of step() time each); %s" % (
        n_mu, len(passes), tot_child // len(passes), tot_ms / 1e3 / len(passes), "unmodified reference, NUM_CPUS=8 pthreads" if kind == "reference" else "plain-C oracle port, 1 thread")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "child terms/s", "n_gpus": args.gpus, "steps": len(passes),
            "warmup": 0, "ms_per_step": tot_ms / len(passes), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": _config(15, workload), "child_terms_per_step": tot_child // len(passes),
            "detail": {"ms_per_mu": [r[3] for r in passes[0][0]], "sampled_mus": n_mu},
            "cpu_baseline": {"value": value, "unit": "child terms/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "child terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gpu_arm(args):
    import numpy as np
    import torch
    from harness import Session, load_product
    from mceio import SHIFT_EXPLICIT, read_scenario
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries ONE JSON line: libraries that chat on it (NCCL prints its version on communicator creation) are sent to stderr, the line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load_product()
    scen, workload = _workload(args)
    sc = read_scenario(scen)
    s = Session(lib, sc, device=local_rank)
    term_sharded = dist is not None and args.shard == "terms"
    if term_sharded:
        from cauchyfriendly_b200.shard import init_term_sharding
        init_term_sharding(s.h, dist, lib=lib, transport="nccl", device=local_rank, moments=args.moments)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def one_pass(collect):
        ev_ms, wall0, rows, gt_ms, gt_bytes, gt_launch, h2d, d2h, launches, phases = 0.0, time.perf_counter(), [], 0.0, 0, 0, 0, 0, 0, None
        prev = 1
        child = 0
        heaviest = (0.0, 0)
        pre_ev, pre_child, pre_wall = 0.0, 0, 0.0
        for k, r in enumerate(sc.rec):
            s.step(r)
            st = s.stats()
            ev_ms += st.ev_step_ms
            child += st.terms_after_muc - prev
            if k < MATCHED_MUS:
                pre_ev += st.ev_step_ms; pre_child += st.terms_after_muc - prev; pre_wall = time.perf_counter() - wall0
            prev = st.survivors if k + 1 < len(sc.rec) else st.terms_after_muc
            gt_ms += st.ev_gtable_ms
            gt_bytes += st.bytes_gtable_algorithmic
            gt_launch += st.gtable_launches
            launches += st.kernel_launches
            d = sc.d
            h2d += 8 * (2 + d * d + d * sc.pncc + sc.pncc + d)      # msmt, gamma, Phi, Gamma, beta, H
            d2h += 16 * (1 + d + d * d) + 8 * 66                   # moment sums + per-shape counters
            if st.ev_step_ms > heaviest[0]:
                heaviest = (st.ev_step_ms, k + 1, st.terms_after_muc, st.survivors)
            if r.shift_kind == SHIFT_EXPLICIT:
                s.shift_b(r.delta, -1.0)
        mo = s.moments()
        wall = time.perf_counter() - wall0
        if term_sharded and collect:
            from cauchyfriendly_b200._capi import MceShardStats
            import ctypes as ct
            ss = MceShardStats(); lib.mce_shard_get_stats(s.h, ct.byref(ss))
            phases = dict(owned_terms=ss.owned_terms, imported_parents=ss.imported_parents, local_parents=ss.local_parents,
                          mb_terms=ss.bytes_terms / 1e6, mb_parent_tables=ss.bytes_parents / 1e6, mb_keys=ss.bytes_keys / 1e6)
        lib.mce_reset(s.h)
        return dict(ev_ms=ev_ms, wall_s=wall, child=child, gt_ms=gt_ms, gt_bytes=gt_bytes, gt_launch=gt_launch, h2d=h2d, d2h=d2h,
                    launches=launches, heaviest=heaviest, Nt=mo.Nt, pre_ev=pre_ev, pre_child=pre_child, pre_wall=pre_wall, shard=phases)

    for _ in range(max(3, args.warmup)):
        one_pass(False)
        l2_flush.fill_(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t_region0 = time.perf_counter()
    acc = []
    for _ in range(args.steps):
        l2_flush.fill_(1)                      # L2 flush between timed iterations
        torch.cuda.synchronize()
        acc.append(one_pass(True))
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    region_s = time.perf_counter() - t_region0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ev_s = sum(a["ev_ms"] for a in acc) / 1e3
    wall_s = sum(a["wall_s"] for a in acc)
    child = sum(a["child"] for a in acc)
    if dist:
        t = torch.tensor([ev_s, wall_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ev_s, wall_s = float(t[0]), float(t[1])
        if not term_sharded:                  # term sharding: every rank works on the same window
            c = torch.tensor([child], dtype=torch.int64, device="cuda")
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            child = int(c[0])
    # N > 1, one window partitioned: the same window on ONE GPU (rank 0 alone) and the replica figure (one 12-MU window per GPU)
    secondary = {}
    if term_sharded:
        def solo(scn, passes):
            sc1 = read_scenario(scn)
            s1 = Session(lib, sc1, device=local_rank)
            ev, ch = 0.0, 0
            for it in range(passes + 1):
                prev = 1
                for k, r in enumerate(sc1.rec):
                    s1.step(r)
                    st = s1.stats()
                    if it > 0:
                        ev += st.ev_step_ms; ch += st.terms_after_muc - prev
                    prev = st.survivors if k + 1 < len(sc1.rec) else st.terms_after_muc
                    if r.shift_kind == SHIFT_EXPLICIT:
                        s1.shift_b(r.delta, -1.0)
                lib.mce_reset(s1.h)
            s1.close()
            return ev / 1e3, ch
        if rank == 0:
            ev1, ch1 = solo(scen, 2)
            secondary["one_gpu_same_window"] = {"value": ch1 / ev1, "ms_per_step": 1e3 * ev1 / 2, "passes": 2}
        dist.barrier()
        evr, chr_ = solo(SCEN, 3)
        t = torch.tensor([evr], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([chr_], dtype=torch.int64, device="cuda"); dist.all_reduce(c, op=dist.ReduceOp.SUM)
        secondary["replicas"] = {"value": int(c[0]) / float(t[0]), "ms_per_step": 1e3 * float(t[0]) / 3, "workload": WORKLOAD, "scaling": "weak", "passes": 3}
        gathered = [None] * world
        dist.all_gather_object(gathered, acc[-1]["shard"])
        secondary["partition_last_mu"] = gathered
    if world == 1:
        # mce_options.fast_moments (off by default): no dependent moment chain -- tree sums, Re fz from the exact scan -- so every count, key and G value stays
        # bit-identical and Im fz / mean / covariance move by reordering noise (~1e-9 / ~1e-6 of their largest entry).  Reported beside the headline, never as it.
        s2 = Session(lib, sc, device=local_rank, fast_moments=True)
        ev2, ch2 = 0.0, 0
        for it in range(4):
            l2_flush.fill_(1); torch.cuda.synchronize()
            prev = 1
            for k, r in enumerate(sc.rec):
                s2.step(r)
                st = s2.stats()
                if it > 0:
                    ev2 += st.ev_step_ms; ch2 += st.terms_after_muc - prev
                prev = st.survivors if k + 1 < len(sc.rec) else st.terms_after_muc
                if r.shift_kind == SHIFT_EXPLICIT:
                    s2.shift_b(r.delta, -1.0)
            lib.mce_reset(s2.h)
        s2.close()
        secondary["fast_moments_option"] = {"value": ch2 / (ev2 / 1e3), "ms_per_step": ev2 / 3, "passes": 3,
                                            "note": "mce_options.fast_moments = 1 (not the default): counts, keys and G bit-identical, Im fz / mean / covariance within reordering noise"}
    if rank == 0:
        peaks, which = _peaks()
        a0 = acc[-1]
        gt_ms_per_launch = sum(a["gt_ms"] for a in acc) / max(1, sum(a["gt_launch"] for a in acc))
        gt_bytes_per_launch = sum(a["gt_bytes"] for a in acc) / max(1, sum(a["gt_launch"] for a in acc))
        achieved = gt_bytes_per_launch / (gt_ms_per_launch * 1e-3) / 1e9 if gt_ms_per_launch > 0 else 0.0
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic_r02.json")      # ncu pass of tools/gpu_profile.sh on this build (not measured in this run)
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("gtable_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": child / ev_s, "unit": "child terms/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": 1e3 * ev_s / args.steps, "higher_is_better": True, "scaling": "strong" if term_sharded else "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": _config(len(sc.rec), workload), "child_terms_per_step": child // (args.steps * (1 if term_sharded else world)),
                "matched": {"mus": MATCHED_MUS, "child_terms": a0["pre_child"], "value": sum(a["pre_child"] for a in acc) / (sum(a["pre_ev"] for a in acc) / 1e3),
                            "e2e_value": sum(a["pre_child"] for a in acc) / sum(a["pre_wall"] for a in acc), "ms": sum(a["pre_ev"] for a in acc) / len(acc), "passes": len(acc)},
                "detail": {"ms_per_mu_mean": 1e3 * ev_s / (args.steps * len(sc.rec)),
                           "heaviest_mu": {"mu": a0["heaviest"][1], "ms": a0["heaviest"][0], "terms_after_muc": a0["heaviest"][2], "survivors": a0["heaviest"][3]},
                           "parallelism": ("one window partitioned over %d gpus: terms routed to the owners of their reduction keys (NCCL send/recv), parent tables fetched from their home ranks" if term_sharded else "window-per-gpu x%d") % world, "l2": "flushed between timed iterations (256 MiB fill)",
                           "moments": "reference serial order (bit-exact)" if not term_sharded or args.moments == "ordered" else (
                               "Re fz bit-exact (exact scan of the serial chain over all ranks' slots => counts, keys and G bit-exact); Im fz, mean, covariance: per-rank two-level sums added in rank order" if args.moments == "hybrid"
                               else "per-rank serial sums added in rank order (last bits depend on N)"), "gtable_share_of_step": sum(a["gt_ms"] for a in acc) / (1e3 * ev_s)},
                "e2e": {"value": child / wall_s, "unit": "child terms/s", "h2d_bytes_per_step": a0["h2d"], "d2h_bytes_per_step": a0["d2h"]},
                "gpu_launches": int(sum(a["launches"] for a in acc)),
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "kernel": "KGTable (child B-table + G-table build per reduction group)", "achieved": achieved,
                             "peak": peaks["hbm_gbs"], "peak_source": which, "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                             "traffic": traffic, "traffic_source": "profiles/traffic_r02.json (ncu dram__bytes_read + dram__bytes_write over every KGTable launch of one cold pass, tools/gpu_profile.sh)" if traffic else None, "bytes_per_launch_algorithmic": gt_bytes_per_launch, "ms_per_launch": gt_ms_per_launch},
                "timed_region_wall_s": region_s}
        line.update(secondary)
        if "one_gpu_same_window" in secondary:
            line["strong_scaling_speedup"] = secondary["one_gpu_same_window"]["ms_per_step"] / line["ms_per_step"]
        # CPU baseline: the reference itself on this box's host cores, bounded sample (MUs 1..9 of the same window)
        if world == 1 and not args.no_cpu_baseline:
            kind, rows, ms, _ = run_reference_sample(12 if not args.short_cpu_baseline else MATCHED_MUS)
            cb = _child_terms(rows) / (ms / 1e3)
            line["cpu_baseline"] = {"value": cb, "unit": "child terms/s", "cores": 8 if kind == "reference" else 1, "kind": kind,
                                    "sample": "%s of the 12-MU window, one pass (%d child terms, %.1f s of step() time), %s" % (
                                        "all 12 MUs" if not args.short_cpu_baseline else "MUs 1..%d" % MATCHED_MUS,
                                        _child_terms(rows), ms / 1e3, "unmodified reference NUM_CPUS=8" if kind == "reference" else "plain-C oracle, 1 thread")}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    s.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--short-cpu-baseline", action="store_true", help="cpu_baseline leg on MUs 1..9 only (2 s instead of ~30 s)")
    ap.add_argument("--ref-full-passes", type=int, default=1, help="--impl reference: passes of the FULL window `value` is measured on")
    ap.add_argument("--shard", default="terms", choices=["windows", "terms"],
                    help="N > 1: ONE window partitioned over the GPUs (default, strong scaling) or one independent window per GPU (weak scaling)")
    ap.add_argument("--window", type=int, default=0, choices=[0, 4, 5], help="time steps of the window: 4 (12 MUs) or 5 (15 MUs); 0 = 4 at N = 1, 5 for a partitioned window")
    ap.add_argument("--moments", default="hybrid", choices=["ordered", "hybrid", "allreduce"],
                    help="partitioned window: every moment sum in the reference's order on every rank (ordered: all bit-exact, the chain does not scale); Re fz by an exact scan over all ranks' slots "
                         "and the other sums per rank (hybrid, default: every count, key and G bit-exact); or all sums per rank (allreduce)")
    a = ap.parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        gpu_arm(a)
