"""End-to-end wall clock of the driver (FASTQ in, FASTA/FASTQ out) on a synthetic cfg2 file."""
import os, sys, time, tempfile
sys.path.insert(0, ".")
import numpy as np
from c3poa_b200 import synth, driver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
tmp = tempfile.mkdtemp(prefix="c3drv_")
blob, off, st = synth.make_batch(n, seed=11)
rng = np.random.default_rng(1)
t0 = time.time()
with open(f"{tmp}/reads.fastq", "wb") as f:
    q = (rng.integers(7, 21, size=int(np.diff(off).max())).astype(np.uint8) + 33).tobytes()
    for i in range(n):
        s = blob[off[i]:off[i + 1]].tobytes()
        f.write(b"@r%07d\n" % i + s + b"\n+\n" + q[:len(s)] + b"\n")
names = [f"r{i:07d}" for i in range(n)]
os.makedirs(f"{tmp}/out/tmp")
synth.write_psl(f"{tmp}/out/tmp/splint_to_read_alignments.psl", names, ["Splint1"] * n, ["-" if x else "+" for x in st])
open(f"{tmp}/splint.fasta", "w").write(f">Splint1\n{synth.SPLINT1}\n")
print("wrote fastq in", round(time.time() - t0, 1), "s", os.path.getsize(f"{tmp}/reads.fastq") / 1e6, "MB")
batch = sys.argv[2] if len(sys.argv) > 2 else "10000"
for infl in (1, 2):
    t0 = time.time()
    tot = driver.main(driver.parse_args(["-r", f"{tmp}/reads.fastq", "-s", f"{tmp}/splint.fasta", "-o", f"{tmp}/out", "--batch", batch, "--inflight", str(infl)]))
    dt = time.time() - t0
    print(f"inflight={infl}: driver {n} reads in {dt:.1f} s -> {n/dt:.0f} reads/s  {tot}")
