"""The group kernel's warp body (c3poa_b200/csrc/poa_grp.cuh) run on the CPU by a fiber warp emulator
(tests/emul/warp_emu.cpp + tests/emul/grp_emul.cu -- the same code the GPU runs, 32 lanes with real shuffles /
ballots / syncwarps) against the oracle: consensus bytes, DP cell counts and graph sizes bit-exact.
Needs nvcc (host code only is executed); no GPU."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c3poa_b200 import synth  # noqa: E402

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
_ENC = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _ENC[_c] = _i
    _ENC[_c | 0x20] = _i


def build_emul():
    out = os.path.join(ROOT, "build", "grp_emul.so")
    src = os.path.join(ROOT, "tests", "emul", "grp_emul.cu")
    rt = os.path.join(ROOT, "tests", "emul", "warp_emu.cpp")
    deps = [src, rt] + [os.path.join(ROOT, "c3poa_b200", "csrc", f) for f in ("poa_grp.cuh", "poa_graph.cuh", "poa_lane.cuh", "poa.cuh", "common.cuh")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        obj = os.path.join(ROOT, "build", "warp_emu.o")
        subprocess.run(["g++", "-O2", "-fPIC", "-c", rt, "-o", obj], check=True)
        subprocess.run([NVCC, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-shared", "-ccbin", "/usr/bin/g++", "-o", out, src, obj], check=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    return build_emul()


def run_emul(lib, groups, para=None, msa2=0, min_seqs=1, node_cap=None, vs_shift=None, rv_shift=3, n_warps=1, graph_gl=32, want_rows=False):
    n = len(groups)
    max_seqs = max(len(g) for g in groups)
    blob = []
    item_base = np.zeros(n, dtype=np.int64)
    bounds = np.zeros((n, max_seqs, 2), dtype=np.int32)
    nseq = np.zeros(n, dtype=np.int32)
    pos = 0
    max_q = 1
    for i, g in enumerate(groups):
        item_base[i] = pos
        o = 0
        nseq[i] = len(g)
        for k, s in enumerate(g):
            bounds[i, k] = (o, o + len(s))
            o += len(s)
            max_q = max(max_q, len(s))
            blob.append(s)
        pos += o
    codes = _ENC[np.frombuffer("".join(blob).encode(), dtype=np.uint8)] if pos else np.zeros(1, dtype=np.uint8)
    codes = np.ascontiguousarray(codes)
    p = dict(match=5, mismatch=4, o1=4, e1=2, o2=24, e2=1, wb=10, wf=0.01, simd_bits=256)
    p.update(para or {})
    if node_cap is None:
        node_cap = (2 + max_q + (max_seqs - 1) * (max_q * 35 // 100 + 16) + 31) & ~31
    node_cap = min(node_cap, 65504)
    cigar_cap = (max_q + node_cap + 64 + 1) & ~1
    qp_stride = (max_q + 48) & ~15
    if vs_shift is None:
        w = p["wb"] + int(p["wf"] * max_q)
        need = (2 * w + 1 + 64) // 16 + 2
        vs_shift = 3
        while (1 << vs_shift) < need:
            vs_shift += 1
    rv_shift = min(rv_shift, vs_shift)         # rv_shift 2 (with vs_shift 3) selects the 4-lanes-per-read instantiation
    cons_cap = max_q * (4 if want_rows else 2) + 64
    cons = np.zeros((n, cons_cap), dtype=np.uint8)
    status = np.full(n, -1, dtype=np.int32)
    clen = np.zeros(n, dtype=np.int32)
    nodes = np.zeros(n, dtype=np.int32)
    cells = np.zeros(n, dtype=np.int64)
    done = np.zeros(n, dtype=np.int32)
    lib.c3g_emul_batch.restype = C.c_int
    lib.c3g_emul_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.c3g_emul_batch(n, codes.ctypes.data, item_base.ctypes.data, bounds.ctypes.data, nseq.ctypes.data, max_seqs,
                            min_seqs, msa2, p["match"], p["mismatch"], p["o1"], p["e1"], p["o2"], p["e2"], p["wb"],
                            p["wf"], p["simd_bits"], node_cap, cigar_cap, qp_stride, vs_shift, rv_shift, n_warps, graph_gl,
                            cons.ctypes.data, cons_cap, status.ctypes.data, clen.ctypes.data, nodes.ctypes.data,
                            cells.ctypes.data, done.ctypes.data)
    assert rc == 0, "warp emulator reported a deadlock"
    out = dict(done=done, status=status, cells=cells, nodes=nodes,
               cons=[cons[i, :clen[i]].tobytes().decode() for i in range(n)])
    if want_rows:      # 2-sequence groups with msa2: [row0 | row1], clen columns each
        out["rows"] = [[cons[i, :clen[i]].tobytes().decode(), cons[i, clen[i]:2 * clen[i]].tobytes().decode()] for i in range(n)]
    return out


def _groups(seed=3):
    rng = np.random.default_rng(seed)
    groups = []
    for L, k in ((300, 3), (500, 5), (1284, 5), (784, 12), (200, 8), (1500, 4), (64, 3), (40, 20), (17, 4), (900, 7)):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(k)])
    a = synth.random_seq(rng, 400).tobytes().decode()
    groups.append([a, a, a])
    groups.append([a])
    groups.append([a, a[:350], a[50:], a[:200] + a[230:]])
    s = list(a); s[10] = "N"; s[200] = "N"
    groups.append(["".join(s), a, a[:100] + "N" + a[100:]])
    b = synth.random_seq(rng, 400).tobytes().decode()
    groups.append([a, b, a, b, a])
    groups.append([a, b])
    for L in rng.integers(30, 700, size=40):
        c = synth.random_seq(rng, int(L))
        groups.append([synth.mutate(rng, c, 0.06, 0.05, 0.05).tobytes().decode() for _ in range(int(rng.integers(2, 7)))])
    return groups


def _check(r, groups, oracle, para=None, must_finish=True):
    bad = []
    for i, g in enumerate(groups):
        if not r["done"][i] and not must_finish:
            continue
        o = oracle.poa_msa(g, para=para)
        if not r["done"][i] or r["status"][i] != 0 or r["cons"][i] != o["cons"] or r["cells"][i] != o["cells"] \
                or r["nodes"][i] != o["node_n"]:
            bad.append((i, int(r["done"][i]), int(r["status"][i]), len(r["cons"][i]), len(o["cons"]),
                        int(r["cells"][i]), int(o["cells"]), int(r["nodes"][i]), int(o["node_n"])))
    assert not bad, f"(group, done, status, |cons| grp/oracle, cells grp/oracle, nodes grp/oracle): {bad}"


def test_grp_body_matches_oracle(emul, oracle):
    groups = _groups()
    _check(run_emul(emul, groups), groups, oracle)
    _check(run_emul(emul, groups, graph_gl=1), groups, oracle)        # small-wave variant of the backtrack (eager loads)


def test_grp_bench_shape(emul, oracle):
    """The bench workload's shape (1 kb insert + 284 nt splint, 5 subreads, 4/3/3 % errors): every read is finished
    by the group kernel (no exactness-guard or capacity fallback) and equals the oracle."""
    rng = np.random.default_rng(2025)
    groups = []
    for _ in range(12):
        a = synth.random_seq(rng, 1284)
        groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(5)])
    r = run_emul(emul, groups)
    assert all(r["done"]) and not any(r["status"])
    _check(r, groups, oracle)


def test_grp_wide_and_deep(emul, oracle):
    """Bands wider than one pass of 8 vectors (long sequences), rings of 8 and 16 vectors, deep graphs."""
    rng = np.random.default_rng(77)
    groups = []
    a = synth.random_seq(rng, 4200)
    groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(3)])
    a = synth.random_seq(rng, 2600)
    groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(4)])
    a = synth.random_seq(rng, 500)
    groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(24)])
    for rv in (3, 4):
        _check(run_emul(emul, groups, rv_shift=rv), groups, oracle)


def test_grp_pairwise_msa_rows(emul, oracle):
    """2-sequence groups with msa2: the group path returns the two MSA rows [row0 | row1] (abPOA's rc_msa order), equal to
    the oracle's; groups of more sequences in the same batch still return their consensus."""
    rng = np.random.default_rng(17)
    groups = []
    for L in (50, 300, 700, 1300, 2500):
        a = synth.random_seq(rng, L)
        groups.append([synth.mutate(rng, a).tobytes().decode(), synth.mutate(rng, a).tobytes().decode()])
    groups.append([groups[0][0], groups[0][0]])
    a = synth.random_seq(rng, 400)
    groups.append([synth.mutate(rng, a, 0.08, 0.06, 0.06).tobytes().decode(), synth.mutate(rng, a, 0.08, 0.06, 0.06).tobytes().decode()])
    groups.append([synth.mutate(rng, a).tobytes().decode() for _ in range(4)])
    n = len(groups)
    max_q = max(len(s) for g in groups for s in g)
    lib = emul
    # run_emul returns cons[:cons_len]; the rows need 2 * cons_len bytes: call the harness directly through a wider slice
    r = run_emul(lib, groups, msa2=1, want_rows=True)
    assert all(r["done"]) and not any(r["status"])
    for i, g in enumerate(groups[:-1]):
        o = oracle.poa_msa(g, out_cons=False, out_msa=True)
        assert r["rows"][i] == o["msa"], i
        assert r["rows"][i][0].replace("-", "") == g[0] and r["rows"][i][1].replace("-", "") == g[1]
    assert r["cons"][-1] == oracle.poa_msa(groups[-1])["cons"]


def test_grp_four_lane_variant(emul, oracle):
    """GL = 4 (8 reads per warp, 4-vector ring; what short sequences get): the standard groups -- including bands wider than
    one pass of 4 lanes -- and 16 short-insert groups that fill two warps, bit-exact vs the oracle."""
    groups = _groups()
    _check(run_emul(emul, groups, vs_shift=3, rv_shift=2), groups, oracle, must_finish=False)
    rng = np.random.default_rng(91)
    short = []
    for _ in range(16):
        a = synth.random_seq(rng, int(rng.integers(500, 900)))
        short.append([synth.mutate(rng, a).tobytes().decode() for _ in range(int(rng.integers(3, 9)))])
    r = run_emul(emul, short, vs_shift=3, rv_shift=2, n_warps=2)
    assert all(r["done"])
    _check(r, short, oracle)


def test_grp_declines_what_it_does_not_cover(emul):
    rng = np.random.default_rng(5)
    a = synth.random_seq(rng, 300)
    g = [synth.mutate(rng, a).tobytes().decode() for _ in range(4)]
    pair = g[:2]
    r = run_emul(emul, [g, pair], msa2=1)
    assert list(r["done"]) == [1, 1]                       # (2-sequence groups: see test_grp_pairwise_msa_rows)
    assert not run_emul(emul, [g], para=dict(simd_bits=128))["done"][0]
    assert not run_emul(emul, [g], para=dict(wb=-1))["done"][0]
    assert not run_emul(emul, [g], node_cap=320)["done"][0]
    assert not run_emul(emul, [g], para=dict(wb=200), vs_shift=3)["done"][0]      # band wider than the arena rows
