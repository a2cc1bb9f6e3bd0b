"""CPU fuzz of the lane kernel's per-thread phases (tests/emul/lane_emul.cu, the same __host__ __device__ code the GPU
runs, 32 states in lockstep) against the oracle: consensus bytes, DP cell counts, graph sizes.

    python tools/fuzz_lane_emul.py shapes [seed]     deep / long / divergent / unrelated-mix groups (768 groups)
    python tools/fuzz_lane_emul.py ties              low-complexity, homopolymer, identical, rotated, extreme indels (640)

Needs build/lane_emul.so (built by `pytest tests/test_lane_emul.py`).  Round-1 result: 0 mismatches, 0 declined in both.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from c3poa_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
import test_lane_emul as T  # noqa: E402

lib = C.CDLL(os.path.join(ROOT, "build", "lane_emul.so"))


def check(groups, sm_vec, tally):
    r = T.run_emul(lib, groups, sm_vec=sm_vec)
    for i, g in enumerate(groups):
        tally["total"] += 1
        if not r["done"][i]:
            tally["declined"] += 1
            continue
        o = pyoracle.poa_msa(g)
        if r["cons"][i] != o["cons"] or r["cells"][i] != o["cells"] or r["nodes"][i] != o["node_n"]:
            tally["bad"] += 1
            print("MISMATCH group", i, len(g), len(g[0]))


def shapes(seed):
    rng = np.random.default_rng(seed)
    tally = dict(total=0, declined=0, bad=0)
    for rnd in range(12):
        groups = []
        for _ in range(64):
            kind = rng.integers(0, 6)
            if kind == 0: L = int(rng.integers(20, 120)); k = int(rng.integers(3, 31))
            elif kind == 1: L = int(rng.integers(400, 1600)); k = int(rng.integers(3, 7))
            elif kind == 2: L = int(rng.integers(100, 600)); k = int(rng.integers(10, 31))
            elif kind == 3: L = int(rng.integers(1500, 3000)); k = int(rng.integers(3, 5))
            else: L = int(rng.integers(200, 1300)); k = int(rng.integers(3, 8))
            a = synth.random_seq(rng, L)
            sub, ins, dele = [(0.04, 0.03, 0.03), (0.10, 0.08, 0.08), (0.01, 0.01, 0.01), (0.2, 0.1, 0.1)][int(rng.integers(0, 4))]
            g = [synth.mutate(rng, a, sub, ins, dele).tobytes().decode() for _ in range(k)]
            if kind == 5:   # an unrelated sequence mixed in, and a truncated one
                g[1] = synth.random_seq(rng, L).tobytes().decode(); g[2] = g[2][:max(10, L // 3)]
            groups.append(g)
        check(groups, int(rng.integers(3, 10)), tally)
        print("round", rnd, tally, flush=True)
    return tally


def ties():
    rng = np.random.default_rng(7)
    tally = dict(total=0, declined=0, bad=0)

    def lowc(L, alpha):   # low complexity: few bases, short repeats -> many score ties
        motif = "".join(rng.choice(list(alpha), size=int(rng.integers(1, 5))))
        return np.frombuffer((motif * (L // len(motif) + 1))[:L].encode(), dtype=np.uint8).copy()
    for rnd in range(10):
        groups = []
        for _ in range(64):
            kind = int(rng.integers(0, 5)); L = int(rng.integers(16, 700)); k = int(rng.integers(3, 12))
            if kind == 0: a = lowc(L, "AC")
            elif kind == 1: a = lowc(L, "ACGT")
            elif kind == 3: a = np.frombuffer(("A" * L).encode(), dtype=np.uint8).copy()
            else: a = synth.random_seq(rng, L)
            rates = [(0.3, 0.15, 0.15), (0.0, 0.0, 0.0), (0.05, 0.2, 0.0), (0.05, 0.0, 0.2), (0.04, 0.03, 0.03)][int(rng.integers(0, 5))]
            g = [synth.mutate(rng, a, *rates).tobytes().decode() or "A" for _ in range(k)]
            if kind == 4:
                g = g[::-1] + [g[0][len(g[0]) // 2:] + g[0][:len(g[0]) // 2]]     # rotated copy
            groups.append(g)
        check(groups, int(rng.integers(2, 9)), tally)
        print("round", rnd, tally, flush=True)
    return tally


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "shapes"
    t = shapes(int(sys.argv[2]) if len(sys.argv) > 2 else 1) if mode == "shapes" else ties()
    sys.exit(1 if t["bad"] else 0)
