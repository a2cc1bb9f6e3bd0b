"""CPU tests: the oracle against the golden vectors produced by the REFERENCE's own Python
(oracle/gen_golden.py; stage 2, split, header naming, pairwise consensus), plus self-consistency
properties of the two parity-unpinned restatements (conk, abPOA)."""
import json
import os

import numpy as np
import pytest

from c3poa_b200 import synth
from c3poa_b200.pairwise import pairwise_consensus

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_stage2_matches_reference_golden(oracle):
    z = np.load(os.path.join(GOLD, "stage2.npz"))
    coef = z["coef"]
    assert np.array_equal(coef, oracle.sg_coeffs(41, 2))
    near_ties = 0
    for i in range(int(z["n_cases"])):
        prof = z[f"profile_{i}"]
        for md, key in ((500, "peaks_"), (120, "peaks_d120_")):
            pk, sm, med = oracle.call_peaks(prof, md, coef=coef)
            assert np.array_equal(pk, z[f"{key}{i}"]), (i, md)
        ref = z[f"smoothed_{i}"]
        # north_star tolerance for smoothed profiles: 1e-5 relative (we are ~1e-13)
        assert np.all(np.abs(sm - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0))
        assert np.max(np.abs(sm - ref)) < 1e-9
        d = np.abs(np.diff(ref))
        near_ties += int(((d > 0) & (d < 1e-9)).sum())
    assert near_ties < 50          # reported, not assumed zero


def test_savgol_constant_and_symmetry(oracle):
    coef = oracle.sg_coeffs(41, 2)
    assert abs(coef.sum() - 1.0) < 1e-12 and np.allclose(coef, coef[::-1], atol=1e-15)
    h = 20
    k = np.arange(-h, h + 1)
    closed = (3 * (3 * h * h + 3 * h - 1) - 15 * k * k) / ((2 * h + 3) * (2 * h + 1) * (2 * h - 1))
    assert np.max(np.abs(coef - closed)) < 1e-15
    out = oracle.savgol(np.full(500, 7.0), coef)
    assert np.max(np.abs(out - 7.0)) < 1e-12


def test_split_matches_reference_golden(oracle):
    for c in json.load(open(os.path.join(GOLD, "split.json"))):
        skip, pks, sb, db = oracle.split(c["peaks"], c["ls"], c["lr"])
        assert skip == (not c["called"]), c["peaks"]
        if c["called"]:
            assert sb.tolist() == c["subs"] and db.tolist() == c["dang"], c["peaks"]


def test_header_naming_matches_reference_golden():
    from c3poa_b200.driver import header
    rng = np.random.default_rng(11)      # same stream as oracle/gen_golden.py:gen_split
    for c in json.load(open(os.path.join(GOLD, "split.json"))):
        rng.integers(0, 4, size=c["lr"])                      # the sequence draw
        qual = (rng.integers(7, 21, size=c["lr"]).astype(np.uint8) + 33).tobytes().decode()
        if c["called"]:
            assert header(c["name"], qual, c["lr"], len(c["subs"]), 100) == c["header"]


def test_pairwise_matches_reference_golden(oracle):
    for c in json.load(open(os.path.join(GOLD, "pairwise.json"))):
        assert pairwise_consensus(c["msa"], [c["s1"], c["s2"]], [c["q1"], c["q2"]]) == c["cons"]
        # the oracle's MSA rows are deterministic and spell the inputs
        msa = oracle.poa_msa([c["s1"], c["s2"]], out_cons=False, out_msa=True)["msa"]
        assert msa == c["msa"]
        assert msa[0].replace("-", "") == c["s1"] and msa[1].replace("-", "") == c["s2"]


def test_conk_profile_properties(oracle):
    """Indirect pins for the parity-unpinned conk restatement: peaks land on splint centres."""
    d = synth.make_reads(6, insert_len=700, repeats=4, seed=3, both_strands=False, err=(0.02, 0.01, 0.01))
    ls = len(synth.SPLINT1)
    for seq in d["seqs"]:
        prof = oracle.conk(synth.SPLINT1, seq, 20)
        assert prof.shape == (len(seq),) and prof.min() >= 0
        pk, _, _ = oracle.call_peaks(prof)
        assert len(pk) == 5
        assert np.all(np.abs(np.diff(pk) - (700 + ls)) < 40)
    # an exact copy of the splint at offset p gives the maximum at d = p
    rng = np.random.default_rng(0)
    seq = synth.random_seq(rng, 500).tobytes().decode() + synth.SPLINT1 + synth.random_seq(rng, 500).tobytes().decode()
    prof = oracle.conk(synth.SPLINT1, seq, 20)
    assert int(np.argmax(prof)) == 500
    assert prof[500] >= 5 * ls * (ls + 1) // 2      # at least the sum 5 + 10 + ... along the diagonal


def test_poa_properties(oracle):
    """Indirect pins for the parity-unpinned abPOA restatement: error-free copies, identity gain."""
    rng = np.random.default_rng(5)
    a = synth.random_seq(rng, 600).tobytes().decode()
    assert oracle.poa_msa([a, a, a, a])["cons"] == a
    truth = synth.random_seq(rng, 800)
    subs = [synth.mutate(rng, truth).tobytes().decode() for _ in range(9)]
    t = truth.tobytes().decode()

    def ident(x):
        import difflib
        return difflib.SequenceMatcher(None, x, t, autojunk=False).ratio()
    c3, c9 = oracle.poa_msa(subs[:3])["cons"], oracle.poa_msa(subs)["cons"]
    assert ident(c9) > ident(c3) > max(ident(s) for s in subs[:3]) - 0.01
    assert ident(c9) > 0.985
    r = oracle.poa_msa(subs[:4], out_msa=True)
    assert all(m.replace("-", "") == s for m, s in zip(r["msa"], subs[:4]))
    assert len({len(m) for m in r["msa"]}) == 1


def test_oracle_batch_threads_agree(oracle):
    d = synth.make_reads(24, insert_len=500, repeat_range=(1, 5), seed=8)
    sp = [synth.SPLINT1, synth.revcomp(synth.SPLINT1)]
    idx = np.array([1 if s == "-" else 0 for s in d["strand"]], dtype=np.int32)
    a = oracle.consensus_batch(d["seqs"], sp, idx, n_threads=1)
    b = oracle.consensus_batch(d["seqs"], sp, idx, n_threads=4)
    assert np.array_equal(a["results"], b["results"]) and np.array_equal(a["cons"], b["cons"])
    assert a["rc"] == 0 and (a["results"]["status"] == 0).sum() > 10


def test_mixed_batch_generator_and_bench_compare(oracle):
    """synth.make_mixed_batch (bench cfg3-cfg5 workloads): classes drawn per read, splint/strand indices consistent with
    the bases; the oracle finds the repeats; bench.compare_with_oracle counts exactly the reads that differ."""
    import bench
    from c3poa_b200 import synth
    blob, off, sp_idx, splints = bench.make_workload("cfg5", 48, 11)
    assert off.size == 49 and off[-1] == blob.size and len(splints) == 8 and sp_idx.min() >= 0 and sp_idx.max() < 8
    b2, o2, i2, _ = bench.make_workload("cfg5", 48, 11)
    assert np.array_equal(blob, b2) and np.array_equal(off, o2) and np.array_equal(sp_idx, i2)      # deterministic
    seqs = [blob[off[i]:off[i + 1]].tobytes().decode() for i in range(48)]
    r = oracle.consensus_batch(seqs, splints, sp_idx, n_threads=4, max_peaks=16, cons_cap=10240)
    st = r["results"]["status"]
    assert (st >= 0).all() and ((st == 0) | (st == 2)).sum() >= 40         # the right splint in the right orientation
    assert len(set(np.diff(off) // 1500)) > 4                              # mixed lengths
    assert bench.compare_with_oracle(r, r, 48) == 0
    r2 = {k: np.array(v, copy=True) for k, v in r.items() if k != "rc"}
    i = int(np.flatnonzero(st == 0)[0])
    r2["cons"][i, 3] ^= 1
    j = int(np.flatnonzero(st == 0)[1])
    r2["results"]["poa_cells"][j] += 1
    assert bench.compare_with_oracle(r2, r, 48) == 2
    blob3, off3, idx3, sp3 = bench.make_workload("cfg3", 6, 5)
    r3 = oracle.consensus_batch([blob3[off3[i]:off3[i + 1]].tobytes().decode() for i in range(6)], sp3, idx3, n_threads=4,
                                max_peaks=64, cons_cap=1536)
    assert (r3["results"]["n_sub"] >= 12).all() and (r3["results"]["status"] == 0).all()


def test_abpoa_named_switches_in_oracle(oracle):
    """the two recalled-but-unverified upstream branches are switches (default off).  int8 lanes only
    matter below ~25 nt; both leave the consensus of clean repeats alone and keep the run valid."""
    from c3poa_b200 import synth
    rng = np.random.default_rng(3)
    a = synth.random_seq(rng, 600)
    g = [synth.mutate(rng, a).tobytes().decode() for _ in range(5)]
    base = oracle.poa_msa(g)
    i8 = oracle.poa_msa(g, para=oracle.default_para(int8_lanes=1))
    assert i8["cons"] == base["cons"] and i8["cells"] == base["cells"] and i8["node_n"] == base["node_n"]
    ec = oracle.poa_msa(g, para=oracle.default_para(end_clamp=1))
    assert ec["cells"] <= base["cells"] and len(ec["cons"]) > 500
    tiny = ["ACGTACGTACGTAC", "ACGTACGAACGTAC", "ACGTACGTACGTAC"]
    t0, t1 = oracle.poa_msa(tiny), oracle.poa_msa(tiny, para=oracle.default_para(int8_lanes=1))
    assert t0["cons"] == t1["cons"] == tiny[0] and t1["cells"] >= t0["cells"]          # granule 32 instead of 16


def test_oracle_against_real_pyabpoa_and_conk_when_importable(oracle):
    """The oracle pinned to the real natives wherever they are importable (skipped offline: parity unpinned)."""
    pa = pytest.importorskip("pyabpoa")
    from c3poa_b200 import synth
    rng = np.random.default_rng(22)
    aligner = pa.msa_aligner(match=5)
    for L in (200, 700, 1284, 2600):
        a = synth.random_seq(rng, L)
        g = [synth.mutate(rng, a).tobytes().decode() for _ in range(5)]
        assert oracle.poa_msa(g)["cons"] == aligner.msa(g, True, False).cons_seq[0], L
    conk = pytest.importorskip("conk")
    seq = synth.make_reads(1, insert_len=600, repeats=4, seed=3)["seqs"][0]
    assert np.array_equal(np.asarray(oracle.conk(synth.SPLINT1, seq, 20)), np.asarray(conk.conk(synth.SPLINT1, seq, 20)))
