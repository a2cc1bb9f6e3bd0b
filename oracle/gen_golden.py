#!/usr/bin/env python3
"""Generate tests/golden/* by running the REFERENCE's own Python for the parts of
the hot path that exist in-tree:

  * bin/savitzky_golay.py + bin/call_peaks.py   (stage 2)       -> stage2.npz
  * C3POa.py analyze_reads (peak shift, split, header naming)   -> split.json
  * bin/consensus.py pairwise_consensus (2-repeat path)         -> pairwise.json

Run in the build container only (needs /root/reference); the fixtures are
committed because the reference does not travel to the GPU box.  conk and
pyabpoa are NOT available anywhere (parity unpinned), so the profiles fed to
stage 2 come from the oracle's conk restatement and from adversarial shapes.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

# shims the reference needs on numpy >= 1.24 (np.int / np.mat were removed)
np.int = int            # noqa
np.mat = np.asmatrix    # noqa
for name in ("mappy", "conk", "pyabpoa"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["conk"].conk = types.ModuleType("conk.conk")
sys.modules["conk.conk"] = sys.modules["conk"].conk
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "bin"))

from call_peaks import call_peaks            # noqa: E402  (reference)
from savitzky_golay import savitzky_golay    # noqa: E402  (reference)
from consensus import pairwise_consensus     # noqa: E402  (reference)
import C3POa as REFMOD                       # noqa: E402  (reference)

from c3poa_b200 import synth                 # noqa: E402
from oracle import pyoracle as O             # noqa: E402


def stage2_cases():
    rng = np.random.default_rng(7)
    cases = []
    # (a) conk-restatement profiles of synthetic reads, several shapes
    for seed, (il, k) in enumerate([(1000, 5), (1000, 3), (500, 8), (600, 2), (2000, 2), (300, 1)]):
        d = synth.make_reads(2, insert_len=il, repeats=k, seed=100 + seed)
        for i in range(2):
            sp = synth.SPLINT1 if d["strand"][i] == "+" else synth.revcomp(synth.SPLINT1)
            cases.append(O.conk(sp, d["seqs"][i], 20))
    # (b) no-splint read (random sequence): gate `max < 6*median` must reject
    cases.append(O.conk(synth.SPLINT1, synth.random_seq(rng, 3000).tobytes().decode(), 20))
    # (c) adversarial integer shapes: plateaus, equal-height peaks, peaks near the ends,
    #     peaks closer than min_dist, constant profile
    n = 2500
    base = rng.integers(5, 15, size=n).astype(np.int32)
    p = base.copy(); p[400:410] += 4000; p[1100:1110] += 4000; p[1800:1810] += 4000      # equal heights
    cases.append(p)
    p = base.copy(); p[300:302] += 9000; p[650:652] += 8000; p[1300:1302] += 9000; p[1500:1502] += 9500   # < min_dist apart
    cases.append(p)
    p = base.copy(); p[0:30] += 5000; p[n - 30:] += 5000; p[1200:1230] += 5000           # edge peaks
    cases.append(p)
    cases.append(np.full(n, 7, dtype=np.int32))                                          # constant
    p = np.zeros(n, dtype=np.int32); p[1000:1100] = 600; p[1700:1800] = 600              # flat tops, zero median
    cases.append(p)
    p = base.copy(); p[::2] += 1; p[700:760] += 300                                      # weak peak under the gate
    cases.append(p)
    return cases


def gen_stage2(out_dir):
    cases = stage2_cases()
    arrs = {}
    for i, prof in enumerate(cases):
        sm = np.asarray(prof)
        for _ in range(3):
            sm = savitzky_golay(sm, 41, 2, deriv=0, rate=1)
        peaks = np.asarray(call_peaks(prof, 500, 3, 41, 2), dtype=np.int64)
        arrs[f"profile_{i}"] = np.asarray(prof, dtype=np.int32)
        arrs[f"smoothed_{i}"] = np.asarray(sm, dtype=np.float64)
        arrs[f"peaks_{i}"] = peaks
        # one more min_dist to pin the distance rule
        arrs[f"peaks_d120_{i}"] = np.asarray(call_peaks(prof, 120, 3, 41, 2), dtype=np.int64)
    arrs["n_cases"] = np.array(len(cases))
    arrs["coef"] = O.sg_coeffs(41, 2)
    np.savez_compressed(os.path.join(out_dir, "stage2.npz"), **arrs)
    print("stage2 cases:", len(cases))


def gen_split(out_dir):
    """Drive the reference's analyze_reads with a stubbed conk (returns a designed profile)
    and a capturing determine_consensus; record bounds and the FASTA header."""
    rng = np.random.default_rng(11)
    cases = []
    tmp = tempfile.mkdtemp(prefix="c3poa_golden_")
    os.mkdir(os.path.join(tmp, "Splint1"))
    args = types.SimpleNamespace(out_path=tmp + "/", mdistcutoff=500, zero=True)
    captured = {}

    def fake_dc(args_, read, subreads, sub_qual, dangling, qual_dangling, racon, tmp_dir, subread_file):
        captured["subs"] = [len(s) for s in subreads]
        captured["sub_seqs"] = list(subreads)
        captured["dang"] = list(dangling)
        return ("ACGT" * 25, len(subreads))

    REFMOD.determine_consensus = fake_dc
    ls = 284
    designs = [
        [300, 1584, 2868, 4152, 5436],                 # regular
        [300, 1584, 2868, 4152, 5436, 6000],           # short last unit -> filtered
        [300, 1575, 2875, 4125, 5475],                 # rounding to 50 incl. .5 cases (1275, 1300, 1250, 1350)
        [50, 1325, 2600],                              # first peak <= 100 after shift? (50+142=192 > 100)
        [10, 1290],                                    # two peaks
        [1500],                                        # single peak -> two dangling halves
        [300, 1584, 2868, 6990],                       # last peak shifted beyond read end -> dropped
        [300, 900, 2184, 3468, 4752, 5300],            # outliers at both ends
        [],                                            # no peaks
        [6950],                                        # only peak dropped by the end filter
        [120, 1145, 2170, 3195, 4220, 5245, 6270],     # 1025 -> rounds to 1000 (half-even on 20.5)
        [120, 1195, 2270, 3345],                       # 1075 -> 21.5 -> 22 -> 1100
    ]
    lr = 7000
    for idx, pk in enumerate(designs):
        seq = synth.random_seq(rng, lr).tobytes().decode()
        qual = (rng.integers(7, 21, size=lr).astype(np.uint8) + 33).tobytes().decode()
        name = f"g{idx:03d}"
        REFMOD.conk.conk = lambda s, q, p, _pk=pk: np.zeros(len(q))   # profile content unused: call_peaks stubbed
        REFMOD.call_peaks = lambda scores, md, it, w, o, _pk=pk: np.asarray(_pk, dtype=np.int64)
        captured.clear()
        fa = os.path.join(tmp, "Splint1", f"tmp{idx}", "R2C2_Consensus.fasta")
        REFMOD.analyze_reads(args, [(name, seq, qual)], {"Splint1": ["A" * ls, "T" * ls]},
                             {name: ["Splint1", "+"]}, {"Splint1"}, idx, "racon")
        header = open(fa).read().split("\n")[0] if os.path.exists(fa) else ""
        # recover bounds from the captured substrings (unique random sequence)
        subs = [(seq.index(s), seq.index(s) + len(s)) for s in captured.get("sub_seqs", [])]
        dang = [(seq.index(s) if s else 0, (seq.index(s) if s else 0) + len(s)) for s in captured.get("dang", [])]
        cases.append(dict(peaks=pk, ls=ls, lr=lr, called=bool(captured), subs=subs, dang=dang,
                          header=header, avg_qual=round(sum(ord(x) - 33 for x in qual) / lr, 2), name=name))
    with open(os.path.join(out_dir, "split.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("split cases:", len(cases))


def gen_pairwise(out_dir):
    """pairwise_consensus over MSA rows of two sequences (oracle POA msa supplies the rows)."""
    rng = np.random.default_rng(13)
    cases = []
    for t in range(8):
        a = synth.random_seq(rng, 300 + 40 * t)
        s1 = synth.mutate(rng, a).tobytes().decode()
        s2 = synth.mutate(rng, a).tobytes().decode()
        q1 = (rng.integers(3, 40, size=len(s1)).astype(np.uint8) + 33).tobytes().decode()
        q2 = (rng.integers(3, 40, size=len(s2)).astype(np.uint8) + 33).tobytes().decode()
        msa = O.poa_msa([s1, s2], out_cons=False, out_msa=True)["msa"]
        cons = pairwise_consensus(msa, [s1, s2], [q1, q2])
        cases.append(dict(s1=s1, s2=s2, q1=q1, q2=q2, msa=msa, cons=cons))
    with open(os.path.join(out_dir, "pairwise.json"), "w") as f:
        json.dump(cases, f)
    print("pairwise cases:", len(cases))


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    gen_stage2(out)
    gen_split(out)
    gen_pairwise(out)
