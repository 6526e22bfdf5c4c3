#!/usr/bin/env python
"""Term-level sharding on real GPUs (run under torchrun, one rank per GPU): every rank replays a scenario with the DCE-TP and
G-table kernels split over the ranks (NCCL all-gathers issued by the library), checks its own results against the golden
dump, and the window time is compared with the unsharded run on rank 0.
Usage: torchrun --nproc-per-node N tools/shard_check.py [scenario] [steps] [full_upto]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch.distributed as dist  # noqa: E402

from cauchyfriendly_b200.shard import init_term_sharding  # noqa: E402
from compare import compare_dumps  # noqa: E402
from harness import Session, load_product, run_scenario  # noqa: E402
from mceio import SHIFT_EXPLICIT, read_dump, read_scenario  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "leo7"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
full = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dist.init_process_group("gloo")                 # only ships the NCCL id and the final verdicts
rank, world, dev = dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
lib = load_product()
gold_dir = os.path.join(ROOT, "tests", "golden")
sc = read_scenario(os.path.join(gold_dir, name + ".mces"))
gold = {n: v for n, v in read_dump(os.path.join(gold_dir, name + ".ref.mced")).items() if n == "header" or int(n.split("/")[0][1:]) <= steps}

# 1. parity of the sharded run, on every rank
got = run_scenario(lib, sc, full_upto=full, max_steps=steps, capture=True, device=dev,
                   on_create=lambda s: init_term_sharding(s.h, dist, lib=lib, transport="nccl", device=dev))
got = {n: v for n, v in got.items() if n in gold}
probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12})
verdicts = [None] * world
dist.all_gather_object(verdicts, "OK" if not probs else "; ".join(probs[:5]))
if rank == 0:
    print("parity vs golden (%s, %d steps, %d ranks):" % (name, steps, world), verdicts, flush=True)


# 2. window time: sharded (all ranks) vs unsharded (rank 0 alone)
def window_ms(s, reps=4):
    best = 1e30
    for _ in range(reps):
        lib.mce_reset(s.h)
        dist.barrier()
        t0 = time.perf_counter()
        for r in sc.rec[:steps]:
            s.step(r)
            if r.shift_kind == SHIFT_EXPLICIT:
                s.shift_b(r.delta, -1.0)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


s = Session(lib, sc, device=dev)
init_term_sharding(s.h, dist, lib=lib, transport="nccl", device=dev)
t_sh = window_ms(s)
s.close()
ts = [None] * world
dist.all_gather_object(ts, t_sh)
if rank == 0:
    s1 = Session(lib, sc, device=dev)
    best = 1e30
    for _ in range(4):
        lib.mce_reset(s1.h)
        t0 = time.perf_counter()
        for r in sc.rec[:steps]:
            s1.step(r)
            if r.shift_kind == SHIFT_EXPLICIT:
                s1.shift_b(r.delta, -1.0)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    s1.close()
    print("window time: %d-rank term sharding %.2f ms (max over ranks), single GPU %.2f ms, speed-up %.2fx" % (world, max(ts), best, best / max(ts)), flush=True)
dist.barrier()
dist.destroy_process_group()
