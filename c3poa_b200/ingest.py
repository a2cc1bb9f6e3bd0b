"""FASTQ/FASTA ingest through the library's C++ reader (c3_fastq_*): length filter + packed batches in the
layout GpuConsensus.stage() takes.  Replaces the reference's two mappy.fastx_read passes
(/root/reference/C3POa.py:201-206,239-254) without Python work per base."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class FastqBatches:
    """Iterates over batches: dict(n, blob uint8, off int64[n+1], qual uint8, names list[str], qual_sum int64[n])."""

    def __init__(self, path: str, min_len: int = 0, max_reads: int = 50000, max_bases: int = 1 << 29, pinned=True,
                 nbuf: int = 1):
        """nbuf > 1 rotates that many buffer sets, so a batch stays valid while the next nbuf-1 are read
        (for callers that keep batches in flight on the GPU)."""
        self._L = _lib.load()
        h = C.c_void_p()
        if self._L.c3_fastq_open(path.encode(), C.byref(h)) != 0:
            raise OSError(f"cannot open {path}")
        self._h = h
        self.min_len, self.max_reads, self.max_bases = int(min_len), int(max_reads), int(max_bases)
        self.n_short = C.c_int64(0)
        alloc = self._alloc_pinned if pinned else (lambda n, dt: np.empty(n, dtype=dt))
        self._sets = []
        for _ in range(max(1, nbuf)):
            self._sets.append(dict(seq=alloc(self.max_bases, np.uint8), qual=alloc(self.max_bases, np.uint8),
                                   off=np.zeros(self.max_reads + 1, dtype=np.int64),
                                   qsum=np.zeros(self.max_reads, dtype=np.int64)))
        self._cur = 0
        self._names = np.zeros(self.max_reads * 64 + 4096, dtype=np.uint8)
        self._name_off = np.zeros(self.max_reads + 1, dtype=np.int64)

    def _alloc_pinned(self, n, dt):
        from .api import PinnedArray
        if not hasattr(self, "_pins"):
            self._pins = []
        p = PinnedArray((n,), dt)
        self._pins.append(p)
        return p.array

    def __iter__(self):
        return self

    def __next__(self):
        b = self._sets[self._cur]
        self._cur = (self._cur + 1) % len(self._sets)
        n = self._L.c3_fastq_next(self._h, self.max_reads, self.max_bases, self.min_len, b["seq"].ctypes.data,
                                  b["qual"].ctypes.data, b["off"].ctypes.data, self._names.ctypes.data,
                                  self._names.size, self._name_off.ctypes.data, b["qsum"].ctypes.data,
                                  C.byref(self.n_short))
        if n < 0:
            raise RuntimeError(f"c3_fastq_next failed ({n}): a record larger than the batch buffers?")
        if n == 0:
            raise StopIteration
        tot = int(b["off"][n])
        nb = self._names[:int(self._name_off[n])].tobytes().split(b"\x00")[:n]
        return dict(n=n, blob=b["seq"][:tot], off=b["off"][:n + 1], qual=b["qual"][:tot],
                    names=[x.decode() for x in nb], qual_sum=b["qsum"][:n],
                    names_raw=self._names[:int(self._name_off[n])].copy(), name_off=self._name_off[:n + 1].copy())

    def close(self):
        if getattr(self, "_h", None):
            self._L.c3_fastq_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_names(names):
    """NUL-terminated names back to back + start offsets (the layout c3_fastq_next returns and c3_format_batch takes)."""
    enc = [x.encode() for x in names]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) + 1 for x in enc])
    raw = np.frombuffer(b"".join(x + b"\x00" for x in enc), dtype=np.uint8).copy() if enc else np.zeros(1, dtype=np.uint8)
    return raw, off


def format_batch(out, names_raw, name_off, blob, qual, off, qual_sum, group=None, which_group=0):
    """Consensus FASTA + subread FASTQ records of the status-0 reads of one output group, formatted by the library
    (c3_format_batch; /root/reference/C3POa.py:167-173, bin/determine_consensus.py:57-77).
    Returns (fasta, fastq, stats dict); the two texts are uint8 arrays (write them as they are: no copy)."""
    L = _lib.load()
    R = np.ascontiguousarray(out["results"])
    n = R.shape[0]
    sb = np.ascontiguousarray(out["sub_bounds"], dtype=np.int32)
    db = np.ascontiguousarray(out["dang_bounds"], dtype=np.int32)
    cons = np.ascontiguousarray(out["cons"])
    max_peaks, cons_cap = sb.shape[1], cons.shape[1]
    names_raw = np.ascontiguousarray(names_raw, dtype=np.uint8)
    name_off = np.ascontiguousarray(name_off, dtype=np.int64)
    blob = np.ascontiguousarray(blob); off = np.ascontiguousarray(off, dtype=np.int64)
    qual = None if qual is None else np.ascontiguousarray(qual)
    qual_sum = np.ascontiguousarray(qual_sum, dtype=np.int64)
    sel = np.ones(n, dtype=bool) if group is None else (np.asarray(group) == which_group)
    ok = sel & (R["status"] == 0)
    name_len = np.diff(name_off)[:n]
    cons_cap_b = int((name_len[ok] + 96).sum() + R["cons_len"][ok].sum()) + 64
    sub_cap_b = int(2 * np.diff(off)[ok].sum() + ((R["n_sub"][ok] + 2) * (name_len[ok] + 32)).sum()) + 64
    oc = np.empty(cons_cap_b, dtype=np.uint8); osb = np.empty(sub_cap_b, dtype=np.uint8)
    lc, ls = C.c_int64(0), C.c_int64(0)
    st = np.zeros(4, dtype=np.int64)
    g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    rc = L.c3_format_batch(n, names_raw.ctypes.data, name_off.ctypes.data, blob.ctypes.data,
                           None if qual is None else qual.ctypes.data, off.ctypes.data, qual_sum.ctypes.data,
                           R.ctypes.data, sb.ctypes.data, db.ctypes.data, max_peaks, cons.ctypes.data, cons_cap,
                           None if g is None else g.ctypes.data, int(which_group), oc.ctypes.data, cons_cap_b,
                           C.byref(lc), osb.ctypes.data, sub_cap_b, C.byref(ls), st.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"c3_format_batch failed ({rc})")
    return oc[:lc.value], osb[:ls.value], dict(consensus=int(st[0]), no_peaks=int(st[1]),
                                                                   left=int(st[2]), errors=int(st[3]))
