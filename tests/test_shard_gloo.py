"""N > 1 path on CPU: term-level sharding of ONE estimator over two torch.distributed ranks (gloo).  Each rank runs the
sequential emulation of the kernels (test infrastructure) on its share of the DCE-TP parents and of the reduction groups
and exchanges the results through the library's callback transport; both ranks must reproduce the reference's golden
dumps bit for bit, exactly like the unsharded run."""
import os
import subprocess
import sys

import pytest

from harness import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["MCE_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MCE_ROOT"], "tests"))
import torch.distributed as dist
from compare import compare_dumps
from harness import load_emu, run_scenario
from mceio import read_dump, read_scenario
from cauchyfriendly_b200.shard import init_term_sharding
dist.init_process_group("gloo")
lib = load_emu()
name, steps, full = os.environ["MCE_CASE"].split(":")
steps, full = int(steps), int(full)
gold_dir = os.path.join(os.environ["MCE_ROOT"], "tests", "golden")
sc = read_scenario(os.path.join(gold_dir, name + ".mces"))
gold = {n: v for n, v in read_dump(os.path.join(gold_dir, name + ".ref.mced")).items() if n == "header" or int(n.split("/")[0][1:]) <= steps}
got = run_scenario(lib, sc, full_upto=full, max_steps=steps, capture=True, split=int(os.environ.get("MCE_SPLIT", "0")),
                   on_create=lambda s: init_term_sharding(s.h, dist, lib=lib, transport="callback"))
got = {n: v for n, v in got.items() if n in gold}
probs = compare_dumps(gold, got, float_rtol=0.0, float_names_rtol={r"fdigest$": 1e-12})
from cauchyfriendly_b200 import shard
if shard.EXCHANGES[0] < 10:
    probs.append("the exchange layer was not used (%d calls)" % shard.EXCHANGES[0])
open(os.environ["MCE_OUT"] + ".%d" % dist.get_rank(), "w").write("\n".join(probs[:20]) if probs else "OK")
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("case,split", [("lti3:8:5", 0), ("lti4_2msmts:8:5", 0), ("leo7:6:3", 0), ("lti3:8:5", 3)])
def test_term_sharded_estimator_matches_golden_on_every_rank(case, split, tmp_path):
    from harness import load_emu
    load_emu(rebuild=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = str(tmp_path / "res")
    env = dict(os.environ, MCE_ROOT=ROOT, MCE_OUT=out, MCE_CASE=case, MCE_SPLIT=str(split))
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29519", str(script)], env=env, timeout=900)
    for r in range(2):
        assert open(out + ".%d" % r).read() == "OK", "rank %d: %s" % (r, open(out + ".%d" % r).read())
