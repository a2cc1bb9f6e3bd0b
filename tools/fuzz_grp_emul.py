"""CPU fuzz of the group kernel's warp body (tests/emul/grp_emul.cu on the fiber warp emulator: the same code the GPU
runs) against the oracle: consensus bytes, DP cell counts, graph sizes.

    python tools/fuzz_grp_emul.py shapes [seed] [rounds]   deep / long / divergent / unrelated-mix groups (32 per round)
    python tools/fuzz_grp_emul.py ties [seed] [rounds]     low-complexity, homopolymer, identical, rotated, extreme indels
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from c3poa_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
import test_grp_emul as T  # noqa: E402

lib = T.build_emul()


def check(groups, rv_shift, tally):
    r = T.run_emul(lib, groups, rv_shift=rv_shift)
    for i, g in enumerate(groups):
        tally["total"] += 1
        if not r["done"][i]:
            tally["declined"] += 1
            continue
        o = pyoracle.poa_msa(g)
        if r["cons"][i] != o["cons"] or r["cells"][i] != o["cells"] or r["nodes"][i] != o["node_n"]:
            tally["bad"] += 1
            print("MISMATCH group", i, len(g), len(g[0]), flush=True)
            np.save("/tmp/fuzz_bad_group.npy", np.array(g, dtype=object), allow_pickle=True)


def shapes(seed, rounds):
    rng = np.random.default_rng(seed)
    tally = dict(total=0, declined=0, bad=0)
    for rnd in range(rounds):
        groups = []
        for _ in range(32):
            kind = rng.integers(0, 7)
            if kind == 0: L = int(rng.integers(20, 120)); k = int(rng.integers(3, 31))
            elif kind == 1: L = int(rng.integers(400, 1600)); k = int(rng.integers(3, 7))
            elif kind == 2: L = int(rng.integers(100, 600)); k = int(rng.integers(10, 31))
            elif kind == 3: L = int(rng.integers(1500, 3000)); k = int(rng.integers(3, 5))
            elif kind == 6: L = int(rng.integers(3000, 5200)); k = 3
            else: L = int(rng.integers(200, 1300)); k = int(rng.integers(3, 8))
            a = synth.random_seq(rng, L)
            sub, ins, dele = [(0.04, 0.03, 0.03), (0.10, 0.08, 0.08), (0.01, 0.01, 0.01), (0.2, 0.1, 0.1)][int(rng.integers(0, 4))]
            g = [synth.mutate(rng, a, sub, ins, dele).tobytes().decode() for _ in range(k)]
            if kind == 5:   # an unrelated sequence mixed in, and a truncated one
                g[1] = synth.random_seq(rng, L).tobytes().decode(); g[2] = g[2][:max(10, L // 3)]
            if rng.integers(0, 8) == 0:   # a few N bases
                s = list(g[0]); s[len(s) // 2] = "N"; g[0] = "".join(s)
            groups.append(g)
        check(groups, int(rng.integers(3, 5)), tally)
        print("round", rnd, tally, flush=True)
    return tally


def ties(seed, rounds):
    rng = np.random.default_rng(seed)
    tally = dict(total=0, declined=0, bad=0)

    def lowc(L, alpha):   # low complexity: few bases, short repeats -> many score ties
        motif = "".join(rng.choice(list(alpha), size=int(rng.integers(1, 5))))
        return np.frombuffer((motif * (L // len(motif) + 1))[:L].encode(), dtype=np.uint8).copy()
    for rnd in range(rounds):
        groups = []
        for _ in range(32):
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
        check(groups, int(rng.integers(3, 5)), tally)
        print("round", rnd, tally, flush=True)
    return tally


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "shapes"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    t = shapes(seed, rounds) if mode == "shapes" else ties(seed, rounds)
    sys.exit(1 if t["bad"] else 0)
