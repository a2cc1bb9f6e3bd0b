import sys, numpy as np
sys.path.insert(0, '.')
from c3poa_b200 import synth
from c3poa_b200.api import GpuConsensus, ReadBatch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
blob, off, st = synth.make_batch(n, seed=3)
sp = synth.SPLINT1 + synth.revcomp(synth.SPLINT1)
b = ReadBatch(blob, off, np.frombuffer(sp.encode(), dtype=np.uint8).copy(), np.array([0, 284, 568], dtype=np.int32), st.astype(np.int32))
g = GpuConsensus(0)
g.stage(b)
for i in range(5):
    g.run(max_peaks=16, cons_cap=2048)
    t = g.timings()
print({k: round(v, 2) for k, v in t.items()})
