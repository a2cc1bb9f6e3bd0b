"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors produced by the reference's own Python.  Bit-exact for integer / index / byte outputs;
smoothed profiles: bit-exact against the oracle, <= 1e-5 relative against the reference (north_star)."""
import json
import os

import numpy as np
import pytest

from c3poa_b200 import synth
from c3poa_b200.api import ReadBatch, default_poa_params, sg_coeffs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _resolve_splints(d):
    names = sorted(d["splints"])
    sp = []
    for nme in names:
        sp += [d["splints"][nme], synth.revcomp(d["splints"][nme])]
    idx = [2 * names.index(s) + (1 if st == "-" else 0) for s, st in zip(d["splint_name"], d["strand"])]
    return sp, np.array(idx, dtype=np.int32)


def _mixed_reads(seed=5):
    rng = np.random.default_rng(seed)
    sp2 = synth.random_seq(rng, 150).tobytes().decode()
    sp3 = synth.random_seq(rng, 700).tobytes().decode()       # > 512 rows: multi-pass conk
    d1 = synth.make_reads(12, insert_len=1000, repeats=5, seed=seed)
    d2 = synth.make_reads(8, insert_len=400, repeat_range=(1, 6), seed=seed + 1, splints={"S2": sp2})
    d3 = synth.make_reads(6, insert_len=800, repeats=3, seed=seed + 2, splints={"S3": sp3})
    seqs, sps, idx = [], [], []
    for d in (d1, d2, d3):
        sp, ix = _resolve_splints(d)
        idx += list(ix + len(sps))
        sps += sp
        seqs += d["seqs"]
    # Ns and lower case
    s = list(seqs[0]); s[100] = "N"; s[2000] = "n"; s[2001] = "a"; seqs[0] = "".join(s)
    seqs.append(synth.random_seq(rng, 1000).tobytes().decode()); idx.append(0)      # no splint
    seqs.append(synth.random_seq(rng, 33).tobytes().decode()); idx.append(1)        # tiny read
    return seqs, sps, np.array(idx, dtype=np.int32)


def test_conk_parity(gpu, oracle):
    seqs, sps, idx = _mixed_reads()
    b = ReadBatch.from_strings(seqs, sps, idx)
    prof = gpu.conk_batch(b, penalty=20)
    bad = []
    for i, s in enumerate(seqs):
        ref = oracle.conk(sps[idx[i]], s, 20)
        got = prof[b.off[i]:b.off[i + 1]]
        if not np.array_equal(ref, got):
            w = np.flatnonzero(ref != got)
            bad.append((i, len(s), len(sps[idx[i]]), int(w[0]), int(w.size), int(ref[w[0]]), int(got[w[0]])))
    assert not bad, f"conk mismatches (read, Lr, Ls, first d, count, ref, got): {bad[:8]}"


def test_conk_penalties(gpu, oracle):
    d = synth.make_reads(4, insert_len=300, repeats=2, seed=9)
    sp, idx = _resolve_splints(d)
    b = ReadBatch.from_strings(d["seqs"], sp, idx)
    for pen in (1, 7, 20, 50):
        prof = gpu.conk_batch(b, penalty=pen)
        for i, s in enumerate(d["seqs"]):
            assert np.array_equal(oracle.conk(sp[idx[i]], s, pen), prof[b.off[i]:b.off[i + 1]]), (pen, i)


def test_peaks_golden(gpu, oracle):
    z = np.load(os.path.join(GOLD, "stage2.npz"))
    n = int(z["n_cases"])
    coef = z["coef"]
    assert np.array_equal(coef, sg_coeffs(41, 2))
    profs = [z[f"profile_{i}"] for i in range(n)]
    off = np.zeros(n + 1, dtype=np.int64); off[1:] = np.cumsum([p.size for p in profs])
    blob = np.concatenate(profs).astype(np.int32)
    for md, key in ((500, "peaks_"), (120, "peaks_d120_")):
        r = gpu.peaks_batch(blob, off, min_dist=md, coef=coef, want_smoothed=True)
        for i in range(n):
            got = r["peaks"][i, :r["n_peaks"][i]]
            assert r["n_peaks"][i] >= 0, (i, r["n_peaks"][i])
            assert np.array_equal(got, z[f"{key}{i}"]), (md, i, got.tolist(), z[f"{key}{i}"].tolist())
            sm = r["smoothed"][off[i]:off[i + 1]]
            ref = z[f"smoothed_{i}"]
            # reference numpy (BLAS summation order): 1e-5 relative (north_star tolerance)
            assert np.all(np.abs(sm - ref) <= 1e-5 * np.maximum(np.abs(ref), 1.0)), i
            # oracle (same fixed summation order): bit-exact
            _, osm, omed = oracle.call_peaks(profs[i], md, coef=coef)
            assert np.array_equal(sm.view(np.int64), osm.view(np.int64)), f"smoothed not bit-exact, case {i}"
            assert r["median"][i] == omed, (i, r["median"][i], omed)


def _poa_groups(seed=3):
    rng = np.random.default_rng(seed)
    groups = []
    for L, k in ((300, 3), (500, 5), (1284, 5), (784, 12), (200, 8), (1500, 4), (64, 3), (40, 20)):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(k)])
    a = synth.random_seq(rng, 400).tobytes().decode()
    groups.append([a, a, a])                                   # identical copies
    groups.append([a])                                         # single sequence
    groups.append([a, a[:350], a[50:], a[:200] + a[230:]])     # length outliers, big deletion
    s = list(a); s[10] = "N"; s[200] = "N"
    groups.append(["".join(s), a, a[:100] + "N" + a[100:]])    # N bases
    b = synth.random_seq(rng, 400).tobytes().decode()
    groups.append([a, b, a, b, a])                             # unrelated sequences mixed
    return groups


def test_poa_parity(gpu, oracle):
    groups = _poa_groups()
    r = gpu.poa_batch(groups)
    bad = []
    for i, g in enumerate(groups):
        o = oracle.poa_msa(g)
        if r["status"][i] != 0 or r["cons"][i] != o["cons"] or r["cells"][i] != o["cells"] or r["nodes"][i] != o["node_n"]:
            bad.append((i, int(r["status"][i]), len(r["cons"][i]), len(o["cons"]), int(r["cells"][i]), int(o["cells"]),
                        int(r["nodes"][i]), int(o["node_n"])))
    assert not bad, f"poa mismatches (group, status, |cons| gpu/oracle, cells gpu/oracle, nodes gpu/oracle): {bad}"


def test_lane_kernel_serves_what_it_covers(gpu, oracle):
    """The fast kernel of the mode (grp: group kernel, lane: thread-per-read kernel) finishes every plain consensus
    group itself (no silent fallback to the warp kernel)."""
    rng = np.random.default_rng(12)
    groups = []
    for L in rng.integers(40, 1500, size=100):
        a = synth.random_seq(rng, int(L))
        groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(int(rng.integers(3, 8)))])
    r = gpu.poa_batch(groups)
    given, done = gpu.lane_counts()
    if gpu.poa_mode == "auto":
        assert (given, done) == (0, 0)                            # a batch this small goes to the warp kernel
    else:
        assert (given, done) == (len(groups), len(groups))        # grp: the group kernel; lane: the lane kernel
    for i in range(0, len(groups), 7):
        o = oracle.poa_msa(groups[i])
        assert r["status"][i] == 0 and r["cons"][i] == o["cons"] and r["cells"][i] == o["cells"] and r["nodes"][i] == o["node_n"], i


def test_poa_pairwise_msa(gpu, oracle):
    """2-repeat path: the two MSA rows (abpoa_generate_rc_msa order) and the reference's pairwise consensus."""
    import json
    from c3poa_b200.pairwise import pairwise_consensus
    cases = json.load(open(os.path.join(GOLD, "pairwise.json")))
    groups = [[c["s1"], c["s2"]] for c in cases]
    rng = np.random.default_rng(17)
    for L in (50, 700, 1300, 2500):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a).tobytes().decode(), synth.mutate(rng, a).tobytes().decode()])
    groups.append([groups[0][0], groups[0][0]])
    r = gpu.poa_batch(groups, want_msa=True)
    for i, g in enumerate(groups):
        o = oracle.poa_msa(g, out_cons=False, out_msa=True)
        assert r["status"][i] == 0 and r["msa"][i] == o["msa"], i
        assert r["msa"][i][0].replace("-", "") == g[0] and r["msa"][i][1].replace("-", "") == g[1]
    for i, c in enumerate(cases):      # golden: reference bin/consensus.py on these rows
        assert r["msa"][i] == c["msa"]
        assert pairwise_consensus(r["msa"][i], [c["s1"], c["s2"]], [c["q1"], c["q2"]]) == c["cons"]


def test_poa_error_free_copies(gpu):
    rng = np.random.default_rng(1)
    seqs = [synth.random_seq(rng, L).tobytes().decode() for L in (100, 333, 1000, 2048)]
    r = gpu.poa_batch([[s] * k for s, k in zip(seqs, (3, 4, 5, 6))])
    assert list(r["status"]) == [0, 0, 0, 0]
    assert r["cons"] == seqs


def _check_fused(gpu, oracle, d, cons_cap=4096, **kw):
    sp, idx = _resolve_splints(d)
    b = ReadBatch.from_strings(d["seqs"], sp, idx)
    out = gpu.consensus_batch(b, max_peaks=64, cons_cap=cons_cap, **kw)
    ref = oracle.consensus_batch(d["seqs"], sp, idx, max_peaks=64, cons_cap=cons_cap, n_threads=8)
    R, G = ref["results"], out["results"]
    bad = []
    for i in range(b.n):
        ok = (R["status"][i] == G["status"][i] and R["n_peaks"][i] == G["n_peaks"][i] and R["n_sub"][i] == G["n_sub"][i]
              and R["n_dang"][i] == G["n_dang"][i])
        if ok and R["status"][i] in (0, 2):
            npk, ns, nd = R["n_peaks"][i], R["n_sub"][i], R["n_dang"][i]
            ok = (np.array_equal(ref["peaks"][i, :npk], out["peaks"][i, :npk])
                  and np.array_equal(ref["sub_bounds"][i, :ns], out["sub_bounds"][i, :ns])
                  and np.array_equal(ref["dang_bounds"][i, :nd], out["dang_bounds"][i, :nd]))
        if ok and R["status"][i] == 2 and R["n_sub"][i] == 2:      # pairwise path: [row0 | row1] of the MSA
            L = R["cons_len"][i]
            ok = (L == G["cons_len"][i] and L > 0 and np.array_equal(ref["cons"][i, :2 * L], out["cons"][i, :2 * L])
                  and R["poa_cells"][i] == G["poa_cells"][i])
        if ok and R["status"][i] == 0:
            ok = (R["cons_len"][i] == G["cons_len"][i]
                  and np.array_equal(ref["cons"][i, :R["cons_len"][i]], out["cons"][i, :G["cons_len"][i]])
                  and R["poa_cells"][i] == G["poa_cells"][i])
        if not ok:
            bad.append((i, [int(x) for x in (R["status"][i], G["status"][i], R["n_peaks"][i], G["n_peaks"][i],
                                             R["n_sub"][i], G["n_sub"][i], R["cons_len"][i], G["cons_len"][i],
                                             R["poa_cells"][i], G["poa_cells"][i])]))
    assert not bad, f"{len(bad)} of {b.n} reads differ; first: {bad[:5]}"
    return out


def test_fused_cfg2_like(gpu, oracle):
    d = synth.make_reads(192, insert_len=1000, repeats=5, seed=21)
    out = _check_fused(gpu, oracle, d)
    assert (out["results"]["status"] == 0).mean() > 0.9


def test_fused_cfg1_mixed_repeats(gpu, oracle):
    d = synth.make_reads(128, insert_len=1000, repeat_range=(1, 5), seed=22)
    _check_fused(gpu, oracle, d)


def test_fused_short_insert_deep(gpu, oracle):
    d = synth.make_reads(24, insert_len=500, repeat_range=(15, 30), seed=23)
    _check_fused(gpu, oracle, d)


def test_fused_long_insert(gpu, oracle):
    d = synth.make_reads(12, insert_len=(3000, 5000), repeat_range=(2, 4), seed=24, flank=(300, 2500))
    _check_fused(gpu, oracle, d, cons_cap=8192)


def test_fused_multi_splint(gpu, oracle):
    rng = np.random.default_rng(4)
    splints = {"Splint1": synth.SPLINT1}
    for k in range(2, 5):
        splints[f"Splint{k}"] = synth.random_seq(rng, 284).tobytes().decode()
    d = synth.make_reads(96, insert_choices=[500, 1000, 2000], repeat_range=(2, 8), seed=25, splints=splints)
    _check_fused(gpu, oracle, d)


def test_properties_at_scale(gpu):
    """Size-independent properties on a batch too large for the oracle to be the checker."""
    d = synth.make_reads(4000, insert_len=1000, repeats=5, seed=26, err=(0.0, 0.0, 0.0))
    sp, idx = _resolve_splints(d)
    b = ReadBatch.from_strings(d["seqs"], sp, idx)
    out = gpu.consensus_batch(b, max_peaks=64, cons_cap=2048)
    R = out["results"]
    assert np.all(R["status"] == 0), np.unique(R["status"], return_counts=True)
    assert np.all(R["n_sub"] == 5) and np.all(R["n_peaks"] == 6)
    unit = len(synth.SPLINT1) + 1000
    assert np.all(R["cons_len"] == unit)
    # error-free copies: the consensus is the repeat unit itself, wherever the peak offsets landed
    for i in range(0, b.n, 97):
        a, e = out["sub_bounds"][i, 0]
        assert out["cons"][i, :unit].tobytes().decode() == d["seqs"][i][a:e]
    # idempotence: running the same batch again gives identical bytes
    out2 = gpu.consensus_batch(b, max_peaks=64, cons_cap=2048)
    assert np.array_equal(out["cons"], out2["cons"]) and np.array_equal(out["results"], out2["results"])


def test_driver_end_to_end(gpu, oracle, tmp_path):
    """The C3POa-compatible driver: CLI flags, PSL hook (BLAT skipped), c3poa.log, Splint_N/ layout,
    header naming and the pre-polish consensus of every repeat count (>=3 POA, 2 pairwise, 1 copy)."""
    from c3poa_b200 import driver
    from c3poa_b200.fastx import fastx_read
    from c3poa_b200.pairwise import pairwise_consensus
    d = synth.make_reads(60, insert_len=700, repeat_range=(1, 5), seed=31)
    short = synth.make_reads(3, insert_len=100, repeats=1, seed=32)          # below --lencutoff
    out = tmp_path / "out"
    (out / "tmp").mkdir(parents=True)
    names = d["names"] + [f"s{i}" for i in range(3)]
    synth.write_fastq(tmp_path / "reads.fastq", names, d["seqs"] + short["seqs"], d["quals"] + short["quals"])
    (tmp_path / "splint.fasta").write_text(f">Splint1\n{synth.SPLINT1}\n")
    synth.write_psl(out / "tmp" / "splint_to_read_alignments.psl", d["names"], d["splint_name"], d["strand"])
    args = driver.parse_args(["-r", str(tmp_path / "reads.fastq"), "-s", str(tmp_path / "splint.fasta"),
                              "-o", str(out), "-l", "1000", "-d", "500"])
    totals = driver.main(args)
    log = (out / "c3poa.log").read_text()
    assert "Total reads: 63" in log and "Under len cutoff: 3" in log and "No splint reads: 0" in log
    cons = {n: s for n, s, _ in fastx_read(str(out / "Splint1" / "R2C2_Consensus.fasta"))}
    subs = list(fastx_read(str(out / "Splint1" / "R2C2_Subreads.fastq")))
    assert totals["errors"] == 0 and totals["consensus"] == len(cons) > 40 and 0 < totals["pairwise"] < len(cons)
    sp, idx = _resolve_splints(d)
    ref = oracle.consensus_batch(d["seqs"], sp, idx, max_peaks=128, cons_cap=8192)
    R = ref["results"]
    seen = 0
    for i, name in enumerate(d["names"]):
        ns = int(R["n_sub"][i])
        sb = ref["sub_bounds"][i, :ns]
        if R["status"][i] == 0:
            exp = ref["cons"][i, :R["cons_len"][i]].tobytes().decode()
        elif R["status"][i] == 2 and ns == 2:
            L = R["cons_len"][i]
            rows = [ref["cons"][i, :L].tobytes().decode(), ref["cons"][i, L:2 * L].tobytes().decode()]
            exp = pairwise_consensus(rows, [d["seqs"][i][a:b] for a, b in sb], [d["quals"][i][a:b] for a, b in sb])
        else:
            continue
        hdr = driver.header(name, d["quals"][i], len(d["seqs"][i]), ns, len(exp))[1:]
        assert cons.get(hdr) == exp, (i, hdr)
        seen += 1
    assert seen == len(cons)
    assert {s[0].rsplit("_", 1)[0] for s in subs} <= set(d["names"])
    have_cons = {h.rsplit("_", 4)[0] for h in cons}
    assert sum(1 for s in subs if s[0].endswith("_1") and s[0][:-2] in have_cons) == len(cons)


def test_driver_polish_flag_runs_one_process_per_batch(gpu, tmp_path):
    """--polish: the driver hands each batch to ONE polishing process and writes what comes back.  racon is
    external; a stand-in appends a marker to every target, so every plain consensus in the output must carry it
    (the 2-repeat pairwise consensi are not polished by this path) and the call count must equal the batch count."""
    import stat
    from c3poa_b200 import driver
    from c3poa_b200.fastx import fastx_read
    if gpu.poa_mode != "auto":
        pytest.skip("one mode is enough for the plumbing")
    d = synth.make_reads(50, insert_len=600, repeats=4, seed=41)
    out = tmp_path / "out"
    (out / "tmp").mkdir(parents=True)
    synth.write_fastq(tmp_path / "reads.fastq", d["names"], d["seqs"], d["quals"])
    (tmp_path / "splint.fasta").write_text(f">Splint1\n{synth.SPLINT1}\n")
    synth.write_psl(out / "tmp" / "splint_to_read_alignments.psl", d["names"], d["splint_name"], d["strand"])
    fake = tmp_path / "fake_racon"
    fake.write_text("#!/usr/bin/env python3\nimport sys\nopen(sys.argv[3] + '.calls', 'a').write('x')\n"
                    "open(%r, 'a').write('x')\n"
                    "for r in open(sys.argv[3]).read().split('>')[1:]:\n"
                    "    name, seq = r.split('\\n')[:2]\n    print('>' + name); print(seq + 'GATTACA')\n" % str(tmp_path / "calls"))
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    (tmp_path / "config").write_text(f"racon\t{fake}\nblat\tblat\n")
    args = driver.parse_args(["-r", str(tmp_path / "reads.fastq"), "-s", str(tmp_path / "splint.fasta"), "-o", str(out),
                              "-c", str(tmp_path / "config"), "--polish", "--batch", "20"])
    totals = driver.main(args)
    cons = list(fastx_read(str(out / "Splint1" / "R2C2_Consensus.fasta")))
    assert totals["errors"] == 0 and totals["polished"] == len(cons) >= 40
    assert all(s.endswith("GATTACA") and int(n.rsplit("_", 1)[1]) == len(s) for n, s, _ in cons)
    assert len((tmp_path / "calls").read_text()) == 3                     # 50 reads in batches of 20


def test_poa_parameter_sweep(gpu, oracle):
    """abPOA keyword arguments other than the reference's (match=5): scoring, band and SIMD-granule variants."""
    rng = np.random.default_rng(41)
    groups = []
    for L, k in ((250, 4), (900, 3), (1284, 5), (600, 7)):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a, 0.05, 0.04, 0.04).tobytes().decode() for _ in range(k)])
    variants = [dict(match=2), dict(match=5, mismatch=2), dict(gap_open1=6, gap_ext1=1, gap_open2=30, gap_ext2=1),
                dict(wb=4, wf=0.0), dict(wb=40, wf=0.05), dict(wb=-1), dict(simd_bits=128), dict(simd_bits=512),
                dict(match=1, mismatch=1, gap_open1=1, gap_ext1=1, gap_open2=2, gap_ext2=1)]
    for kw in variants:
        r = gpu.poa_batch(groups, params=default_poa_params(**kw))
        for i, g in enumerate(groups):
            o = oracle.poa_msa(g, para=oracle.default_para(**kw))
            assert r["status"][i] == 0 and r["cons"][i] == o["cons"] and r["cells"][i] == o["cells"] \
                and r["nodes"][i] == o["node_n"], (kw, i, int(r["status"][i]), int(r["cells"][i]), int(o["cells"]))


def test_poa_band_too_narrow_fails_loudly(gpu, oracle):
    """extra_b = extra_f = 0: the alignment leaves the band; oracle and GPU both report a backtrack failure."""
    rng = np.random.default_rng(41)
    a = synth.random_seq(rng, 900)
    g = [synth.mutate(rng, a, 0.05, 0.04, 0.04).tobytes().decode() for _ in range(3)]
    with pytest.raises(RuntimeError):
        oracle.poa_msa(g, para=oracle.default_para(wb=0, wf=0.0))
    r = gpu.poa_batch([g], params=default_poa_params(wb=0, wf=0.0))
    assert r["status"][0] < 0 and r["cons"][0] == "", (int(r["status"][0]), len(r["cons"][0]))   # -206 backtrack / -203 cell pool


def test_wide_int32_mode_and_long_sequences(gpu, oracle):
    """qlen*5 > 32767-10 switches the reference build to int32 lanes (band granule 8): sequences > 6.5 kb."""
    rng = np.random.default_rng(43)
    a = synth.random_seq(rng, 7000)
    g = [synth.mutate(rng, a, 0.03, 0.02, 0.02).tobytes().decode() for _ in range(3)]
    r = gpu.poa_batch([g])
    o = oracle.poa_msa(g)
    assert r["status"][0] == 0 and r["cons"][0] == o["cons"] and r["cells"][0] == o["cells"]


def test_peaks_edge_cases(gpu, oracle):
    """Ragged batch: short, flat, monotone and spiky profiles in one call; every count must match the oracle."""
    rng = np.random.default_rng(47)
    profs = [np.zeros(30, np.int32), np.full(1000, 3, np.int32), np.arange(2000, dtype=np.int32),
             np.arange(2000, dtype=np.int32)[::-1].copy(), rng.integers(0, 50, 5000).astype(np.int32),
             (rng.integers(0, 20, 9000) + 5000 * (np.arange(9000) % 1500 == 700)).astype(np.int32),
             np.zeros(20, np.int32)]
    off = np.zeros(len(profs) + 1, dtype=np.int64); off[1:] = np.cumsum([p.size for p in profs])
    r = gpu.peaks_batch(np.concatenate(profs), off, min_dist=500, want_smoothed=True)
    for i, p in enumerate(profs):
        if p.size < 22:
            assert r["n_peaks"][i] == -1          # the reference's padding needs 21 samples; flagged, never silent
            continue
        pk, sm, med = oracle.call_peaks(p, 500)
        assert r["n_peaks"][i] == len(pk) and np.array_equal(r["peaks"][i, :len(pk)], pk), i
        assert np.array_equal(r["smoothed"][off[i]:off[i + 1]].view(np.int64), sm.view(np.int64)), i


def test_capacity_errors_are_reported(gpu):
    """Overflows surface as negative per-read status codes (never silently wrong)."""
    d = synth.make_reads(6, insert_len=400, repeats=4, seed=51)
    sp, idx = _resolve_splints(d)
    b = ReadBatch.from_strings(d["seqs"], sp, idx)
    out = gpu.consensus_batch(b, max_peaks=16, cons_cap=64)         # consensus does not fit 64 bytes
    assert np.all(out["results"]["status"] == -209)
    out = gpu.consensus_batch(b, max_peaks=2, cons_cap=2048)        # more peaks than max_peaks
    assert np.all(out["results"]["status"] < 0)
    out = gpu.consensus_batch(b, max_peaks=16, cons_cap=2048)
    assert np.all(out["results"]["status"] == 0)


def test_cfg5_like_mixed_batch_properties(gpu, oracle):
    """cfg 5 shape (mixed inserts, 4 splints, 2-10 repeats): oracle parity on a sample + batch invariants."""
    rng = np.random.default_rng(4)
    splints = {"Splint1": synth.SPLINT1}
    for k in range(2, 5):
        splints[f"Splint{k}"] = synth.random_seq(rng, 284).tobytes().decode()
    d = synth.make_reads(1500, insert_choices=[500, 1000, 2000, 4000], repeat_range=(2, 10), seed=55, splints=splints)
    sp, idx = _resolve_splints(d)
    b = ReadBatch.from_strings(d["seqs"], sp, idx)
    out = gpu.consensus_batch(b, max_peaks=32, cons_cap=16384)
    R = out["results"]
    assert np.all(R["status"] >= 0), np.unique(R["status"], return_counts=True)
    ok = R["status"] == 0
    assert ok.mean() > 0.8
    # consensus length tracks the repeat unit (insert + splint) within 3 %
    unit = np.array([len(t) + 284 for t in d["truth"]])
    assert np.all(np.abs(R["cons_len"][ok] - unit[ok]) <= 0.06 * unit[ok] + 10)
    # subread bounds are ordered, inside the read, and contiguous peak to peak
    for i in range(0, b.n, 37):
        ns = R["n_sub"][i]
        sb = out["sub_bounds"][i, :ns]
        assert np.all(sb[:, 0] < sb[:, 1]) and sb.min(initial=0) >= 0 and sb.max(initial=0) <= len(d["seqs"][i])
    sample = list(range(0, b.n, 25))
    ref = oracle.consensus_batch([d["seqs"][i] for i in sample], sp, idx[sample], max_peaks=32, cons_cap=16384, n_threads=8)
    for k, i in enumerate(sample):
        assert ref["results"]["status"][k] == R["status"][i] and ref["results"]["cons_len"][k] == R["cons_len"][i], i
        L = R["cons_len"][i] * (2 if (R["status"][i] == 2 and R["n_sub"][i] == 2) else 1)
        assert np.array_equal(ref["cons"][k, :L], out["cons"][i, :L]), i


def test_lane_and_warp_kernels_agree_at_scale(gpu):
    """45 000 cfg2 reads: `auto` hands the batch to the group kernel; the warp-per-read kernel must return the same
    bytes for every read (status, peaks, bounds, consensus, DP cell counts, graph sizes)."""
    if gpu.poa_mode != "auto":
        pytest.skip("runs once, switching modes itself")
    blob, off, strand = synth.make_batch(45000, insert_len=1000, repeats=5, seed=77)
    sp = synth.SPLINT1 + synth.revcomp(synth.SPLINT1)
    b = ReadBatch(blob, off, np.frombuffer(sp.encode(), dtype=np.uint8).copy(), np.array([0, 284, 568], dtype=np.int32),
                  strand.astype(np.int32))
    try:
        a = gpu.consensus_batch(b, max_peaks=16, cons_cap=2048)
        given, done = gpu.lane_counts()
        assert given >= 44000 and done >= given - 20, (given, done)     # the group kernel ran, and finished (nearly) all it took:
        # a read with a row wider than the arena's 8 vectors goes to the warp kernel, compared below like the rest
        a = {k: np.array(v, copy=True) for k, v in a.items()}
        gpu.set_poa_mode("warp")
        w = gpu.consensus_batch(b, max_peaks=16, cons_cap=2048)
        assert gpu.lane_counts()[0] == 0
    finally:
        gpu.set_poa_mode("auto")
    for f in ("status", "n_peaks", "n_sub", "cons_len", "poa_cells", "poa_nodes"):
        assert np.array_equal(a["results"][f], w["results"][f]), f
    assert (a["results"]["status"] == 0).mean() > 0.95
    assert np.array_equal(a["peaks"], w["peaks"]) and np.array_equal(a["sub_bounds"], w["sub_bounds"])
    L = a["results"]["cons_len"]
    mask = np.arange(a["cons"].shape[1])[None, :] < L[:, None]
    assert np.array_equal(a["cons"][mask], w["cons"][mask])


def test_driver_multi_process_sharding(gpu, tmp_path):
    """--gpus N: one process per GPU, reads sharded by index, per-rank tmp dirs concatenated (here both
    ranks are mapped onto the box's GPUs modulo the device count, so the path runs on a 1-GPU box too)."""
    from c3poa_b200 import _lib, driver
    from c3poa_b200.fastx import fastx_read
    d = synth.make_reads(40, insert_len=500, repeat_range=(3, 5), seed=61)
    (tmp_path / "a" / "tmp").mkdir(parents=True); (tmp_path / "b" / "tmp").mkdir(parents=True)
    synth.write_fastq(tmp_path / "reads.fastq", d["names"], d["seqs"], d["quals"])
    (tmp_path / "splint.fasta").write_text(f">Splint1\n{synth.SPLINT1}\n")
    for o in ("a", "b"):
        synth.write_psl(tmp_path / o / "tmp" / "splint_to_read_alignments.psl", d["names"], d["splint_name"], d["strand"])
    base = ["-r", str(tmp_path / "reads.fastq"), "-s", str(tmp_path / "splint.fasta")]
    driver.main(driver.parse_args(base + ["-o", str(tmp_path / "a")]))
    ngpu = _lib.load().c3_device_count()
    os.environ["C3POA_DEVICE_MODULO"] = str(ngpu)
    try:
        driver.main(driver.parse_args(base + ["-o", str(tmp_path / "b"), "--gpus", "2", "--batch", "7"]))
    finally:
        del os.environ["C3POA_DEVICE_MODULO"]
    # many small batches, three in flight on one GPU
    (tmp_path / "c" / "tmp").mkdir(parents=True)
    synth.write_psl(tmp_path / "c" / "tmp" / "splint_to_read_alignments.psl", d["names"], d["splint_name"], d["strand"])
    driver.main(driver.parse_args(base + ["-o", str(tmp_path / "c"), "--batch", "9", "--inflight", "3"]))
    c = {n: s for n, s, _ in fastx_read(str(tmp_path / "c" / "Splint1" / "R2C2_Consensus.fasta"))}
    a = {n: s for n, s, _ in fastx_read(str(tmp_path / "a" / "Splint1" / "R2C2_Consensus.fasta"))}
    b = {n: s for n, s, _ in fastx_read(str(tmp_path / "b" / "Splint1" / "R2C2_Consensus.fasta"))}
    assert a == b == c and len(a) == 40              # output order is unspecified in the reference: compare as sets
    assert not any(p.name.startswith("tmp") for p in (tmp_path / "b" / "Splint1").iterdir())


def test_splint_assignment_on_gpu(gpu, tmp_path):
    """f-3: the conk kernel over every splint x strand finds the known splint and strand of synthetic reads;
    reads without a splint stay below the acceptance fraction; the driver gives the same consensi with
    --assign gpu as with the PSL."""
    from c3poa_b200 import driver
    from c3poa_b200.fastx import fastx_read
    rng = np.random.default_rng(71)
    splints = {"Splint1": synth.SPLINT1}
    for k in range(2, 5):
        splints[f"Splint{k}"] = synth.random_seq(rng, 284).tobytes().decode()
    d = synth.make_reads(160, insert_choices=[500, 1000, 2000], repeat_range=(1, 6), seed=72, splints=splints)
    sp_names = sorted(splints)
    cands = [s for n in sp_names for s in (splints[n], synth.revcomp(splints[n]))]
    truth = np.array([2 * sp_names.index(s) + (1 if st == "-" else 0) for s, st in zip(d["splint_name"], d["strand"])])
    junk = [synth.random_seq(rng, 3000).tobytes().decode() for _ in range(10)]
    b = ReadBatch.from_strings(d["seqs"] + junk, cands, np.zeros(170, dtype=np.int32))
    best, scores = gpu.assign_splints(b.blob, b.off, cands)
    assert np.array_equal(best[:160], truth)
    perfect = 5 * 284 * 285 // 2
    top = scores[best, np.arange(170)]
    assert top[:160].min() > 0.05 * perfect > top[160:].max(), (top[:160].min(), top[160:].max())
    second = np.sort(scores[:, :160], axis=0)[-2]
    assert np.all(top[:160] > 3 * second)                       # clear margin over every other candidate
    # driver: --assign gpu reproduces the PSL-driven run
    for o in ("psl", "gpu"):
        (tmp_path / o / "tmp").mkdir(parents=True)
    synth.write_fastq(tmp_path / "reads.fastq", d["names"], d["seqs"], d["quals"])
    (tmp_path / "splint.fasta").write_text("".join(f">{n}\n{s}\n" for n, s in splints.items()))
    synth.write_psl(tmp_path / "psl" / "tmp" / "splint_to_read_alignments.psl", d["names"], d["splint_name"], d["strand"])
    base = ["-r", str(tmp_path / "reads.fastq"), "-s", str(tmp_path / "splint.fasta")]
    driver.main(driver.parse_args(base + ["-o", str(tmp_path / "psl")]))
    driver.main(driver.parse_args(base + ["-o", str(tmp_path / "gpu"), "--assign", "gpu"]))
    for n in sp_names:
        a = {x: s for x, s, _ in fastx_read(str(tmp_path / "psl" / n / "R2C2_Consensus.fasta"))}
        g = {x: s for x, s, _ in fastx_read(str(tmp_path / "gpu" / n / "R2C2_Consensus.fasta"))}
        assert a == g and len(a) > 10, n
    assert "No splint reads: 0" in (tmp_path / "gpu" / "c3poa.log").read_text()


def test_fused_very_long_reads(gpu, oracle):
    """cfg 4 upper end: 30-55 kb concatemers (5 kb inserts x 5-9), plus a read whose single subread is too long for
    the graph (> 65 000 columns is rejected loudly, not mis-computed)."""
    d = synth.make_reads(4, insert_len=5000, repeat_range=(5, 9), seed=81, flank=(500, 3000))
    out = _check_fused(gpu, oracle, d, cons_cap=16384)
    assert np.all(out["results"]["status"] == 0) and out["results"]["cons_len"].min() > 5000
    rng = np.random.default_rng(82)
    a = synth.random_seq(rng, 70000).tobytes().decode()
    from c3poa_b200.api import GpuError
    with pytest.raises(GpuError, match="too long"):
        gpu.poa_batch([[a, a[:69000], a]], cons_cap=80000)


def test_drop_in_shims(gpu, oracle):
    """The per-call drop-ins keep the reference's signatures and return types:
    conk.conk(splint, seq, penalty), call_peaks(scores, min_dist, iters, window, order),
    pyabpoa.msa_aligner(match=5).msa(seqs, out_cons, out_msa)."""
    from c3poa_b200.shims.conk import conk
    from c3poa_b200.shims.call_peaks import call_peaks
    import c3poa_b200.shims.pyabpoa as poa
    d = synth.make_reads(3, insert_len=400, repeats=4, seed=91, both_strands=False)
    seq = d["seqs"][0]
    scores = conk.conk(synth.SPLINT1, seq, 20)
    assert np.array_equal(np.asarray(scores), oracle.conk(synth.SPLINT1, seq, 20))
    peaks = call_peaks(scores, 500, 3, 41, 2)
    ref_peaks, _, _ = oracle.call_peaks(np.asarray(scores), 500)
    assert isinstance(peaks, np.ndarray) and peaks.dtype == np.int64 and np.array_equal(peaks, ref_peaks)
    assert list(peaks + len(synth.SPLINT1) // 2)                       # usable as C3POa.py:125-127 uses it
    assert call_peaks(np.full(800, 3), 500, 3, 41, 2) == []            # gate not passed -> [] like the reference
    _, pk, sb, _ = oracle.split(ref_peaks, len(synth.SPLINT1), len(seq))
    subs = [seq[a:b] for a, b in sb]
    res = poa.msa_aligner(match=5).msa(subs, out_cons=True, out_msa=True)
    assert res.cons_seq[0] == oracle.poa_msa(subs)["cons"] and res.n_seq == len(subs)
    res2 = poa.msa_aligner(match=5).msa(subs[:2], out_cons=False, out_msa=True)
    assert res2.msa_seq == oracle.poa_msa(subs[:2], out_cons=False, out_msa=True)["msa"] and not res2.cons_seq


@pytest.mark.gpu
def test_auto_group_kernel_vs_oracle_12k(gpu, oracle):
    """12 000 reads of the bench workload's shape: the smallest batch `auto` hands to the group kernel.  Every output of the
    fused C-ABI call (c3_consensus_batch) against the oracle on the same reads."""
    if gpu.poa_mode != "auto":
        pytest.skip("runs once")
    import bench
    n = 12000
    blob, off, sp_idx, splints = bench.make_workload("cfg2_1kb_x5", n, 4242)
    b = ReadBatch(blob, off, np.frombuffer("".join(splints).encode(), dtype=np.uint8).copy(),
                  np.array([0, 284, 568], dtype=np.int32), np.ascontiguousarray(sp_idx, dtype=np.int32))
    out = gpu.consensus_batch(b, max_peaks=16, cons_cap=2048)
    given, done = gpu.lane_counts()
    assert given >= 11500 and done >= given - 20, (given, done)
    t = gpu.timings()
    assert t["poa_dp_launches"] >= 4 and t["poa_graph_launches"] >= 5 and t["poa_dp_ms"] > 0
    seqs = [blob[off[i]:off[i + 1]].tobytes().decode() for i in range(n)]
    r = oracle.consensus_batch(seqs, splints, sp_idx, n_threads=os.cpu_count() or 1, max_peaks=16, cons_cap=2048)
    assert bench.compare_with_oracle(out, r, n) == 0
    assert np.array_equal(out["dang_bounds"], r["dang_bounds"])
    assert (out["results"]["status"] == 0).mean() > 0.95


@pytest.mark.gpu
def test_mixed_batch_bulk_to_group_kernel_tail_to_warp(gpu, oracle):
    """cfg5-like mix (inserts 500-4000, 2-10 repeats, 4 splints) padded with short reads so that `auto` sends the bulk of
    short subreads to the group kernel and the long / 2-repeat ones to the warp kernel in the same call; a sample of both
    kinds against the oracle."""
    if gpu.poa_mode != "auto":
        pytest.skip("runs once")
    import bench
    n = 44000
    blob, off, sp_idx, splints = bench.make_workload("cfg5", n, 99)
    sp_off = np.zeros(len(splints) + 1, dtype=np.int32)
    sp_off[1:] = np.cumsum([len(x) for x in splints])
    b = ReadBatch(blob, off, np.frombuffer("".join(splints).encode(), dtype=np.uint8).copy(), sp_off,
                  np.ascontiguousarray(sp_idx, dtype=np.int32))
    out = gpu.consensus_batch(b, max_peaks=16, cons_cap=10240)
    given, done = gpu.lane_counts()
    res = out["results"]
    in_poa = int((res["n_sub"] >= 2).sum())
    assert given >= 24000 and done >= given - 50 and in_poa - done > 1000, (given, done, in_poa)
    assert (res["status"] < 0).sum() == 0
    m = 600
    seqs = [blob[off[i]:off[i + 1]].tobytes().decode() for i in range(m)]
    r = oracle.consensus_batch(seqs, splints, sp_idx[:m], n_threads=os.cpu_count() or 1, max_peaks=16, cons_cap=10240)
    assert bench.compare_with_oracle(out, r, m) == 0


@pytest.mark.gpu
def test_abpoa_named_switches_on_gpu(gpu, oracle):
    """c3_set_abpoa_switches: with either switch set every read runs through the warp kernel, which implements both; same
    bytes, cell counts and graph sizes as the oracle with the same switches."""
    rng = np.random.default_rng(8)
    groups = [["ACGTACGTACGTAC", "ACGTACGAACGTAC", "ACGTACGTACGTAC"], ["ACGTTGCAAC"] * 3]
    for L in (18, 60, 300, 900, 1500):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a, 0.06, 0.05, 0.05).tobytes().decode() for _ in range(4)])
    try:
        for i8, ec in ((1, 0), (0, 1), (1, 1)):
            gpu.set_abpoa_switches(bool(i8), bool(ec))
            r = gpu.poa_batch(groups)
            assert gpu.lane_counts() == (0, 0)
            for k, g in enumerate(groups):
                o = oracle.poa_msa(g, para=oracle.default_para(int8_lanes=i8, end_clamp=ec))
                assert r["status"][k] == 0 and r["cons"][k] == o["cons"] and r["cells"][k] == o["cells"] \
                    and r["nodes"][k] == o["node_n"], (i8, ec, k)
    finally:
        gpu.set_abpoa_switches(False, False)


@pytest.mark.gpu
def test_against_real_pyabpoa_and_conk_when_importable(gpu, oracle):
    """wherever the real natives are importable (pyabpoa 1.0.5, conk) the GPU path and the oracle are
    compared with them and the suite reports `oracle = real`; offline (this image) the test is skipped and parity with
    upstream stays unpinned."""
    pa = pytest.importorskip("pyabpoa")
    rng = np.random.default_rng(21)
    groups = []
    for L in (300, 800, 1284, 1284, 2200):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(5)])
    r = gpu.poa_batch(groups)
    aligner = pa.msa_aligner(match=5)
    for k, g in enumerate(groups):
        real = aligner.msa(g, True, False).cons_seq[0]
        assert oracle.poa_msa(g)["cons"] == real, ("oracle vs pyabpoa", k)
        assert r["cons"][k] == real, ("GPU vs pyabpoa", k)
    conk = pytest.importorskip("conk")
    seq = synth.make_reads(1, insert_len=600, repeats=4, seed=3)["seqs"][0]
    real_prof = np.asarray(conk.conk(synth.SPLINT1, seq, 20), dtype=np.int64)
    b = ReadBatch.from_strings([seq], [synth.SPLINT1], np.zeros(1, dtype=np.int32))
    assert np.array_equal(gpu.conk_batch(b, 20).astype(np.int64), real_prof)
    print("oracle = real")


@pytest.mark.gpu
def test_driver_zero_repeat_records_and_long_read_batches(gpu, oracle, tmp_path):
    """Two things the advisor found in round 1.  (1) Zero-repeat reads (one splint): the reference's zero_repeats writes the
    two dangling halves as @name_0 / @name_1 before it overlaps them with mappy; the driver writes those records and
    counts the reads in c3poa.log instead of dropping them silently (-z switches the records off).  (2) A read far longer
    than the rest of its batch is held back and run in a batch of its own (the consensus buffers are dense), with the
    same result."""
    if gpu.poa_mode != "auto":
        pytest.skip("driver creates its own handles")
    from c3poa_b200 import driver
    from c3poa_b200.fastx import fastx_read
    d = synth.make_reads(30, insert_len=700, repeat_range=(3, 5), seed=71)
    z = synth.make_reads(4, insert_len=1500, repeats=0, seed=72, flank=(600, 1200))
    lg = synth.make_reads(1, insert_len=1000, repeats=40, seed=73)
    names = d["names"] + [f"z{i}" for i in range(4)] + ["long0"]
    seqs, quals = d["seqs"] + z["seqs"] + lg["seqs"], d["quals"] + z["quals"] + lg["quals"]
    assert len(lg["seqs"][0]) > 40000
    for sub, extra in (("a", []), ("b", ["-z"])):
        out = tmp_path / sub
        (out / "tmp").mkdir(parents=True)
        synth.write_fastq(tmp_path / "reads.fastq", names, seqs, quals)
        (tmp_path / "splint.fasta").write_text(f">Splint1\n{synth.SPLINT1}\n")
        synth.write_psl(out / "tmp" / "splint_to_read_alignments.psl", names, ["Splint1"] * len(names),
                        d["strand"] + z["strand"] + lg["strand"])
        totals = driver.main(driver.parse_args(["-r", str(tmp_path / "reads.fastq"), "-s", str(tmp_path / "splint.fasta"),
                                                "-o", str(out), "-l", "1000", "-d", "500"] + extra))
        cons = {n.rsplit("_", 4)[0]: s for n, s, _ in fastx_read(str(out / "Splint1" / "R2C2_Consensus.fasta"))}
        subs = [s[0] for s in fastx_read(str(out / "Splint1" / "R2C2_Subreads.fastq"))]
        log = (out / "c3poa.log").read_text()
        assert totals["errors"] == 0 and totals.get("held_back_long_reads") == 1 and totals.get("zero", 0) >= 3
        assert f"Zero-repeat reads without consensus" in log and log.rstrip().endswith(str(totals["zero"]))
        zrec = [n for n in subs if n.startswith("z")]
        if extra:
            assert not zrec and totals.get("zero_records", 0) == 0
        else:
            assert totals["zero_records"] == totals["zero"] and len(zrec) == 2 * totals["zero"]
            assert {n[-2:] for n in zrec} == {"_0", "_1"}
        # the long read: same consensus as the oracle's, produced by the batch of its own
        sp = [synth.SPLINT1, synth.revcomp(synth.SPLINT1)]
        ref = oracle.consensus_batch(lg["seqs"], sp, np.array([1 if lg["strand"][0] == "-" else 0], dtype=np.int32),
                                     max_peaks=128, cons_cap=4096)
        assert ref["results"]["status"][0] == 0 and ref["results"]["n_sub"][0] >= 30
        assert cons["long0"] == ref["cons"][0, :ref["results"]["cons_len"][0]].tobytes().decode()


@pytest.mark.gpu
def test_conk_packed_kernel_small_cases(gpu, oracle):
    """c3_conk2_kernel (two reads per warp in 16-bit halves; what batches of >= 4096 reads get): forced onto small and
    awkward inputs -- reads without partner, pairs of very different length, N bases, two splints of different length,
    splint lengths around the lane multiples, penalties 1-50 -- and compared with the oracle read by read."""
    if gpu.poa_mode != "auto":
        pytest.skip("runs once")
    rng = np.random.default_rng(5)
    os.environ["C3POA_CONK_PACKED_MIN"] = "1"
    try:
        for ls in (31, 32, 33, 200, 284, 288, 289, 479, 480):
            sps = [synth.random_seq(rng, ls).tobytes().decode(), synth.random_seq(rng, max(20, ls - 37)).tobytes().decode()]
            seqs, idx = [], []
            for k, L in enumerate((900, 4100, 60, 2049, 2047, 5000, 1500, 33, 7000)):
                sp = k % 2
                core = synth.random_seq(rng, max(10, L // 3)).tobytes().decode()
                s = (core + sps[sp] + core + sps[sp] + core)[:L]
                if k == 1:
                    s = s[:50] + "N" + s[51:200] + "n" + s[201:]
                seqs.append(s); idx.append(sp)
            b = ReadBatch.from_strings(seqs, sps, np.array(idx, dtype=np.int32))
            for pen in ((20,) if ls not in (284, 33) else (1, 7, 20, 50)):
                prof = gpu.conk_batch(b, penalty=pen)
                for i, s in enumerate(seqs):
                    assert np.array_equal(oracle.conk(sps[idx[i]], s, pen), prof[b.off[i]:b.off[i + 1]]), (ls, pen, i)
        # the same inputs through the one-read kernel give the same bytes
        os.environ["C3POA_CONK_INT32"] = "1"
        assert np.array_equal(gpu.conk_batch(b, penalty=20), prof)
    finally:
        os.environ.pop("C3POA_CONK_PACKED_MIN", None); os.environ.pop("C3POA_CONK_INT32", None)
