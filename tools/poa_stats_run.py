import os, sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
os.environ['C3POA_GPU_LIB'] = 'build/variants/lib_stats.so'
from c3poa_b200 import synth, _lib
from c3poa_b200.api import GpuConsensus, ReadBatch
L = _lib.load()
blob, off, st = synth.make_batch(8000, seed=3)
sp = synth.SPLINT1 + synth.revcomp(synth.SPLINT1)
b = ReadBatch(blob, off, np.frombuffer(sp.encode(), dtype=np.uint8).copy(), np.array([0,284,568],dtype=np.int32), st.astype(np.int32))
g = GpuConsensus(0)
z = (C.c_ulonglong*16)()
L.c3_debug_stats(z, 1)
out = g.consensus_batch(b, max_peaks=16, cons_cap=2048)
L.c3_debug_stats(z, 0)
v = list(z)
names = ['rows','pred0_not_in_ring','multi_pred_rows','sum_npre','sum_ng','rows_ng>32','sum_width','passes','passes_mono','spec_rounds','spec_steps','generic_steps','merge_nondel_ops','merge_complex_ops']
for n,x in zip(names, v): print(f'{n:20s} {x:12d}  per_row {x/max(v[0],1):.3f}')
print('reads ok', int((out['results']['status']==0).sum()))
