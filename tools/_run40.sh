cd $GRAFT_REPO_ROOT
python - <<'PY'
import cProfile, pstats, sys, os, time, tempfile, io
sys.path.insert(0, ".")
import numpy as np
from c3poa_b200 import synth, driver
n = 200000
tmp = tempfile.mkdtemp(prefix="c3drv_")
blob, off, st = synth.make_batch(n, seed=11)
rng = np.random.default_rng(1)
with open(f"{tmp}/reads.fastq", "wb") as f:
    q = (rng.integers(7, 21, size=int(np.diff(off).max())).astype(np.uint8) + 33).tobytes()
    for i in range(n):
        s = blob[off[i]:off[i + 1]].tobytes()
        f.write(b"@r%07d\n" % i + s + b"\n+\n" + q[:len(s)] + b"\n")
names = [f"r{i:07d}" for i in range(n)]
os.makedirs(f"{tmp}/out/tmp")
synth.write_psl(f"{tmp}/out/tmp/splint_to_read_alignments.psl", names, ["Splint1"] * n, ["-" if x else "+" for x in st])
open(f"{tmp}/splint.fasta", "w").write(f">Splint1\n{synth.SPLINT1}\n")
args = driver.parse_args(["-r", f"{tmp}/reads.fastq", "-s", f"{tmp}/splint.fasta", "-o", f"{tmp}/out", "--batch", "100000", "--inflight", "1"])
pr = cProfile.Profile()
t0 = time.time()
pr.enable(); tot = driver.main(args); pr.disable()
dt = time.time() - t0
print(f"driver {n} reads in {dt:.1f} s -> {n/dt:.0f} reads/s  {tot}")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
PY
