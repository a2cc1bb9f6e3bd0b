"""One POA-only launch of the group kernel on synthetic cfg2 groups, for ncu captures (small, so that the replay
passes stay cheap).  usage: python tools/grp_ncu_run.py [reads] [mode]"""
import sys

import numpy as np

sys.path.insert(0, '.')
from c3poa_b200 import synth  # noqa: E402
from c3poa_b200.api import GpuConsensus, ReadBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
mode = sys.argv[2] if len(sys.argv) > 2 else "grp"
blob, off, st = synth.make_batch(n, seed=3)
sp = synth.SPLINT1 + synth.revcomp(synth.SPLINT1)
b = ReadBatch(blob, off, np.frombuffer(sp.encode(), dtype=np.uint8).copy(), np.array([0, 284, 568], dtype=np.int32), st.astype(np.int32))
g = GpuConsensus(0, poa_mode=mode)
out = g.consensus_batch(b, max_peaks=16, cons_cap=2048)
print('timings', g.timings(), 'fast kernel', g.lane_counts(), 'ok', int((out['results']['status'] == 0).sum()))
