"""Minimal FASTA/FASTQ reader and revcomp so that the driver does not hard-require mappy for I/O
(the reference uses mm.fastx_read / mm.revcomp: /root/reference/C3POa.py:201,232-234,239)."""
from __future__ import annotations

import gzip

_COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")


def revcomp(seq: str) -> str:
    return seq.encode().translate(_COMP)[::-1].decode()


def fastx_read(path: str):
    """Yields (name, seq, qual) like mappy.fastx_read(path, read_comment=False); qual is None for FASTA."""
    with open(path, "rb") as probe:                       # gzip by magic bytes, like the C++ ingest (not by suffix)
        is_gz = probe.read(2) == b"\x1f\x8b"
    op = gzip.open if is_gz else open
    with op(path, "rt") as f:
        line = f.readline()
        while line:
            if line.startswith("@"):
                name = line[1:].split()[0] if line[1:].strip() else ""
                seq = f.readline().rstrip("\n")
                f.readline()
                qual = f.readline().rstrip("\n")
                yield name, seq, qual
                line = f.readline()
            elif line.startswith(">"):
                name = line[1:].split()[0] if line[1:].strip() else ""
                parts = []
                line = f.readline()
                while line and not line.startswith(">"):
                    parts.append(line.strip())
                    line = f.readline()
                yield name, "".join(parts), None
            else:
                line = f.readline()
