"""Per-phase cycle shares of c3_poa_graph_kernel (warp-cycles per phase, lane 0 of every warp) (library built with -DC3L_PROF into build/variants/lib_gprof.so).
usage: python tools/grp_prof_run.py [reads]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, '.')
os.environ.setdefault('C3POA_GPU_LIB', 'build/variants/lib_gprof.so')
from c3poa_b200 import synth, _lib  # noqa: E402
from c3poa_b200.api import GpuConsensus, ReadBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
L = _lib.load()
blob, off, st = synth.make_batch(n, seed=3)
sp = synth.SPLINT1 + synth.revcomp(synth.SPLINT1)
b = ReadBatch(blob, off, np.frombuffer(sp.encode(), dtype=np.uint8).copy(), np.array([0, 284, 568], dtype=np.int32), st.astype(np.int32))
g = GpuConsensus(0, poa_mode="grp")
z = (C.c_ulonglong * 24)()
g.consensus_batch(b, max_peaks=16, cons_cap=2048)
L.c3_debug_lane_prof(z, 1)
out = g.consensus_batch(b, max_peaks=16, cons_cap=2048)
L.c3_debug_lane_prof(z, 0)
v = list(z)
names = ['load state', 'backtrack', 'merge', 'reorder', 'prepare', 'consensus+end']
tot = max(sum(v[6:12]), 1)
for i, nm in enumerate(names):
    print(f'{nm:18s} {v[6+i]/1e9:10.3f} Gcycles {100*v[6+i]/tot:6.2f} %')
print('timings', g.timings(), 'grp', g.lane_counts(), 'ok', int((out['results']['status'] == 0).sum()))
