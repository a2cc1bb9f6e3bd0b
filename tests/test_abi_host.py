"""CPU tests: the C-ABI library loads and exports every symbol include/c3poa_gpu.h declares (no
compute calls without a GPU), the product fails loudly without a device, and the host-side logic."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "c3poa_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(c3_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from c3poa_b200 import _lib, build
    build.build()
    L = _lib.load()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/c3poa_gpu.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert b"sm_100a" in L.c3_version()


def test_struct_layouts_match_header():
    import ctypes as C
    from c3poa_b200 import _lib
    assert C.sizeof(_lib.PoaParams) == 40 and _lib.RESULT_DTYPE.itemsize == 32 and C.sizeof(_lib.Timings) == 32


def test_no_cpu_fallback_without_device():
    from c3poa_b200 import _lib
    if _lib.load().c3_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from c3poa_b200.api import GpuConsensus, GpuError
    with pytest.raises(GpuError):
        GpuConsensus(0)


def test_product_never_imports_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "c3poa_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", ""), os.path.join(dp, f)
    out = subprocess.run([sys.executable, "-c", "import sys; import c3poa_b200, c3poa_b200.api, c3poa_b200.driver; "
                          "print(any(m.startswith('oracle') for m in sys.modules))"], cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stderr


def test_readbatch_packing_and_synth_determinism():
    from c3poa_b200 import synth
    from c3poa_b200.api import ReadBatch
    d1 = synth.make_reads(5, insert_len=300, repeats=2, seed=42)
    d2 = synth.make_reads(5, insert_len=300, repeats=2, seed=42)
    assert d1["seqs"] == d2["seqs"] and d1["strand"] == d2["strand"]
    b = ReadBatch.from_strings(d1["seqs"], [synth.SPLINT1, synth.revcomp(synth.SPLINT1)], [0, 1, 0, 1, 0])
    assert b.n == 5 and b.off[-1] == sum(map(len, d1["seqs"])) and b.seq(3) == d1["seqs"][3]
    assert b.sp_off.tolist() == [0, 284, 568]
    blob, off, st = synth.make_batch(300, insert_len=200, repeats=3, seed=9)
    blob2, off2, st2 = synth.make_batch(300, insert_len=200, repeats=3, seed=9, workers=1)
    assert np.array_equal(blob, blob2) and np.array_equal(off, off2) and np.array_equal(st, st2)
    assert set(np.unique(blob)) <= set(b"ACGT")
    assert synth.revcomp(synth.revcomp(synth.SPLINT1)) == synth.SPLINT1


def test_driver_psl_and_cli(tmp_path):
    from c3poa_b200 import driver, synth
    from c3poa_b200.fastx import fastx_read
    d = synth.make_reads(6, insert_len=300, repeats=2, seed=1)
    fq, psl = tmp_path / "r.fastq", tmp_path / "a.psl"
    synth.write_fastq(fq, d["names"], d["seqs"], d["quals"])
    synth.write_psl(psl, d["names"], d["splint_name"], d["strand"])
    got = list(fastx_read(str(fq)))
    assert [g[0] for g in got] == d["names"] and [g[1] for g in got] == d["seqs"] and got[0][2] == d["quals"][0]
    ad, aset, no = driver.read_psl(str(psl), d["names"] + ["ghost"])
    assert no == 1 and aset == {"Splint1"} and ad[d["names"][2]] == ("Splint1", d["strand"][2])
    a = driver.parse_args(["-r", str(fq), "-s", "s.fa", "-o", str(tmp_path), "-l", "500", "-d", "300", "-z", "-co"])
    assert (a.lencutoff, a.mdistcutoff, a.zero, a.compress_output, a.groupSize, a.numThreads) == (500, 300, False, True, 1000, 1)
    assert driver.header("r1", chr(33 + 20) * 5200, 5200, 3, 1250) == ">r1_20.0_5200_3_1250"


def test_shard_range_covers_everything():
    from c3poa_b200.dist import shard_range
    for n in (0, 1, 7, 100000, 100003):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_cxx_ingest_matches_python_reader(tmp_path):
    """f-2: the library's FASTQ/FASTA reader (plain + gzip) against the pure-Python reader."""
    import gzip
    import shutil
    from c3poa_b200 import synth
    from c3poa_b200.fastx import fastx_read
    from c3poa_b200.ingest import FastqBatches
    d = synth.make_reads(37, insert_len=250, repeat_range=(1, 4), seed=3)
    short = synth.make_reads(5, insert_len=100, repeats=1, seed=4, flank=(10, 20))
    names = d["names"] + [f"short{i} extra comment" for i in range(5)]
    fq = tmp_path / "r.fastq"
    synth.write_fastq(fq, names, d["seqs"] + short["seqs"], d["quals"] + short["quals"])
    with open(fq, "rb") as f, gzip.open(str(fq) + ".gz", "wb") as g:
        shutil.copyfileobj(f, g)
    ref = [r for r in fastx_read(str(fq))]
    for path in (str(fq), str(fq) + ".gz"):
        for min_len in (0, 700):
            fb = FastqBatches(path, min_len=min_len, max_reads=8, max_bases=40000, pinned=False)
            got = []
            for b in fb:
                assert b["n"] <= 8 and b["off"][-1] <= 40000
                for i in range(b["n"]):
                    a, e = b["off"][i], b["off"][i + 1]
                    got.append((b["names"][i], b["blob"][a:e].tobytes().decode(), b["qual"][a:e].tobytes().decode()))
                    assert b["qual_sum"][i] == sum(ord(c) - 33 for c in got[-1][2])
            exp = [r for r in ref if len(r[1]) >= min_len]
            assert got == exp and fb.n_short.value == len(ref) - len(exp)
    fa = tmp_path / "s.fasta"
    fa.write_text(">A desc\nACGT\nAC\n>B\nTTTT\n")
    b = next(FastqBatches(str(fa), pinned=False))
    assert b["names"] == ["A", "B"] and b["blob"].tobytes() == b"ACGTACTTTT" and b["off"].tolist() == [0, 6, 10]
