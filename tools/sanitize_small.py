"""Small end-to-end batch for compute-sanitizer runs (memcheck / racecheck / synccheck / initcheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from c3poa_b200 import synth
from c3poa_b200.api import GpuConsensus, ReadBatch

d = synth.make_reads(24, insert_len=300, repeat_range=(1, 6), seed=3)
rng = np.random.default_rng(1)
sp2 = synth.random_seq(rng, 600).tobytes().decode()          # multi-pass conk
d2 = synth.make_reads(4, insert_len=250, repeats=3, seed=4, splints={"S2": sp2})
seqs = d["seqs"] + d2["seqs"]
splints = [synth.SPLINT1, synth.revcomp(synth.SPLINT1), sp2, synth.revcomp(sp2)]
idx = [1 if s == "-" else 0 for s in d["strand"]] + [3 if s == "-" else 2 for s in d2["strand"]]
g = GpuConsensus(0)
out = g.consensus_batch(ReadBatch.from_strings(seqs, splints, np.array(idx, dtype=np.int32)), max_peaks=32, cons_cap=4096)
print("status", np.unique(out["results"]["status"], return_counts=True))
r = g.poa_batch([[seqs[0][:400], seqs[0][5:390]], [seqs[1][:300]] * 3], want_msa=True)
print("poa", r["status"])
g.close()
