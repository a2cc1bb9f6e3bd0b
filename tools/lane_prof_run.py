"""Per-phase cycle shares of c3_poa_lane_kernel (library built with -DC3L_PROF into build/variants/lib_lprof.so).
usage: python tools/lane_prof_run.py [reads]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, '.')
os.environ.setdefault('C3POA_GPU_LIB', 'build/variants/lib_lprof.so')
from c3poa_b200 import synth, _lib  # noqa: E402
from c3poa_b200.api import GpuConsensus, ReadBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
L = _lib.load()
blob, off, st = synth.make_batch(n, seed=3)
sp = synth.SPLINT1 + synth.revcomp(synth.SPLINT1)
b = ReadBatch(blob, off, np.frombuffer(sp.encode(), dtype=np.uint8).copy(), np.array([0, 284, 568], dtype=np.int32), st.astype(np.int32))
g = GpuConsensus(0, poa_mode="lane")
z = (C.c_ulonglong * 24)()
g.consensus_batch(b, max_peaks=16, cons_cap=2048)
if hasattr(L, 'c3_debug_lane_prof'):
    L.c3_debug_lane_prof(z, 1)
out = g.consensus_batch(b, max_peaks=16, cons_cap=2048)
if hasattr(L, 'c3_debug_lane_prof'):
    L.c3_debug_lane_prof(z, 0)
v = list(z)
names = ['item_begin', 'align_begin', 'source_row', 'row_setup', 'row_compute', 'align_end(rest)', 'backtrack', 'merge', 'item_end']
tot = max(sum(v[:9]), 1)
for i, nm in enumerate(names):
    print(f'{nm:18s} {v[i]/1e9:10.3f} Gcycles {100*v[i]/tot:6.2f} %')
print('warp row steps', v[9], 'mean mv', v[10] / max(v[9], 1), 'cycles/row-step compute', v[4] / max(v[9], 1), 'setup', v[3] / max(v[9], 1))
print('backtrack trips', v[11], 'cycles/trip', v[6] / max(v[11], 1), 'F trips', v[12])
print('row loop per row-step: load-wait', v[13] / max(v[9], 1), 'arith', v[14] / max(v[9], 1), 'store+argmax', v[15] / max(v[9], 1))
print('row pre/post per row-step: setup-records', v[16] / max(v[9], 1), 'fold', v[17] / max(v[9], 1), 'codes+first loads', v[18] / max(v[9], 1), 'epilogue (records of next rows arrive)', v[19] / max(v[9], 1))
print('timings', g.timings(), 'lane', g.lane_counts(), 'ok', int((out['results']['status'] == 0).sum()))
