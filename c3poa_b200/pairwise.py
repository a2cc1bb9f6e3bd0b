"""Quality-aware consensus of two aligned sequences: the 2-repeat path of the reference
(/root/reference/bin/consensus.py:4-81, called from bin/determine_consensus.py:33-41 with the two
MSA rows abPOA returns).  Re-stated here column-run-wise; results are checked against the
reference's own function through tests/golden/pairwise.json."""
from __future__ import annotations


def _expand_quality(row: str, qual: str) -> str:
    """Quality string stretched to the row's length (normalizeLen, consensus.py:50-74): a gap column
    takes the first quality when nothing precedes it, else the integer mean of its two neighbours;
    trailing gaps repeat the last produced character."""
    out = []
    qi = 0
    nq = len(qual)
    for ch in row:
        if qi >= nq:
            break
        if ch != "-":
            out.append(qual[qi]); qi += 1
        elif qi == 0:
            out.append(qual[0])
        else:
            out.append(chr(int((ord(qual[qi - 1]) + ord(qual[qi])) / 2)))
    missing = len(row) - len(out)
    if missing > 0:
        # the reference appends newQuality[-1] once per trailing '-' it finds
        k = 0
        while k < len(row) and row[-1 - k] == "-":
            out.append(out[-1]); k += 1
    return "".join(out)


def pairwise_consensus(rows, subreads, quals) -> str:
    """rows: the two MSA rows; subreads/quals: the two ungapped sequences and their quality strings."""
    qual_of = {s: q for s, q in zip(subreads, quals)}      # later duplicate wins, as in the reference dict
    a, b = rows[0], rows[1]
    qa = _expand_quality(a, qual_of[a.replace("-", "")])
    qb = _expand_quality(b, qual_of[b.replace("-", "")])
    out = []
    i, n = 0, len(a)
    while i != n:
        ca, cb = a[i], b[i]
        if ca == cb:
            out.append(ca)
        if ca != cb and ca != "-" and cb != "-":
            out.append(ca if ord(qa[i]) > ord(qb[i]) else cb)
        if ca == "-" or cb == "-":
            gapped = a if ca == "-" else b
            run = 1
            try:
                while gapped[i + run] == "-":
                    run += 1
            except IndexError:                              # a gap that reaches the end counts as length 1
                run = 1
            ma = sum(map(ord, qa[i:i + run])) / run
            mb = sum(map(ord, qb[i:i + run])) / run
            out.append(a[i:i + run] if ma > mb else b[i:i + run])
            i += run
            continue
        i += 1
    return "".join(out).replace("-", "")
