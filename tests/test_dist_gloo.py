"""World-size-2 test of the multi-GPU plumbing on CPU (gloo): read sharding + the reductions bench.py
uses.  There is no data-path collective to test: ranks never exchange read data."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["C3_ROOT"])
import numpy as np
from c3poa_b200.dist import Group, shard_range
from c3poa_b200 import synth
from oracle import pyoracle as O            # CPU stand-in for the per-rank work (tests only)
g = Group("gloo")
d = synth.make_reads(10, insert_len=300, repeats=3, seed=77)
lo, hi = shard_range(10, g.world, g.rank)
sp = [synth.SPLINT1, synth.revcomp(synth.SPLINT1)]
idx = np.array([1 if s == "-" else 0 for s in d["strand"]], dtype=np.int32)
r = O.consensus_batch(d["seqs"][lo:hi], sp, idx[lo:hi])
n_ok = g.allsum(float((r["results"]["status"] == 0).sum()))
t = g.allmax(float(g.rank + 1))
g.barrier()
total = g.allsum(float(hi - lo))
if g.rank == 0:
    print(json.dumps(dict(world=g.world, n_ok=n_ok, tmax=t, total=total)))
g.close()
'''


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, C3_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == 2 and r["tmax"] == 2.0 and r["total"] == 10.0
    # same answer as a single process over all reads
    import numpy as np
    sys.path.insert(0, ROOT)
    from c3poa_b200 import synth
    from oracle import pyoracle as O
    d = synth.make_reads(10, insert_len=300, repeats=3, seed=77)
    sp = [synth.SPLINT1, synth.revcomp(synth.SPLINT1)]
    idx = np.array([1 if s == "-" else 0 for s in d["strand"]], dtype=np.int32)
    full = O.consensus_batch(d["seqs"], sp, idx)
    assert r["n_ok"] == float((full["results"]["status"] == 0).sum())


def test_strong_scaling_shares_cover_the_batch():
    """bench.py --scaling strong: the per-rank shares add up to the batch and differ by at most one read."""
    import bench
    for total in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            shares = [bench.rank_share(total, world, r) for r in range(world)]
            assert sum(shares) == total and max(shares) - min(shares) <= 1
