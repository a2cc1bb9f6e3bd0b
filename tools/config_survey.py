"""Run the BASELINE.json config shapes at moderate scale: status histogram, throughput, GCUPS per stage."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from c3poa_b200 import synth
from c3poa_b200.api import GpuConsensus, ReadBatch

def resolve(d):
    names = sorted(d["splints"]); sp = []
    for n in names: sp += [d["splints"][n], synth.revcomp(d["splints"][n])]
    idx = [2 * names.index(s) + (1 if st == "-" else 0) for s, st in zip(d["splint_name"], d["strand"])]
    return sp, np.array(idx, dtype=np.int32)

mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = GpuConsensus(0, poa_mode=mode)
rng = np.random.default_rng(4)
sp4 = {"Splint1": synth.SPLINT1, **{f"Splint{k}": synth.random_seq(rng, 284).tobytes().decode() for k in range(2, 5)}}
cfgs = {
    "cfg1 1kb x3-5": dict(n_reads=4000, insert_len=1000, repeat_range=(3, 5), seed=1),
    "cfg3 500bp x15-30": dict(n_reads=3000, insert_len=500, repeat_range=(15, 30), seed=3),
    "cfg4 3-5kb x2-4": dict(n_reads=1500, insert_len=(3000, 5000), repeat_range=(2, 4), seed=4, flank=(300, 2500)),
    "cfg5 mixed 4 splints": dict(n_reads=4000, insert_choices=[500, 1000, 2000, 4000], repeat_range=(2, 10), seed=5, splints=sp4),
}
for name, kw in cfgs.items():
    kw = dict(kw); kw["n_reads"] *= scale
    d = synth.make_reads(**kw)
    sp, idx = resolve(d)
    b = ReadBatch.from_strings(d["seqs"], sp, idx)
    cap = 2 * int(np.diff(b.off).max())
    for _ in range(2):
        out = g.consensus_batch(b, max_peaks=64, cons_cap=min(cap, 20000))
    t = g.timings(); R = out["results"]
    u, c = np.unique(R["status"], return_counts=True)
    cells = int(R["poa_cells"].sum())
    print(f"{name:22s} reads {b.n:5d} bases {int(b.off[-1])/1e6:7.1f}M status {dict(zip(map(int,u),map(int,c)))} "
          f"total {t['total_ms']:8.1f} ms ({b.n/t['total_ms']*1e3:9.0f} reads/s) conk {t['conk_ms']:7.1f} peaks {t['peaks_ms']:6.1f} "
          f"poa {t['poa_ms']:8.1f} ms  POA {cells/t['poa_ms']/1e6:6.1f} GCUPS  max nodes {int(R['poa_nodes'].max())}  lane given/done {g.lane_counts()}")
