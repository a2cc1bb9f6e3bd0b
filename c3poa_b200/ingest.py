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
                    names=[x.decode() for x in nb], qual_sum=b["qsum"][:n])

    def close(self):
        if getattr(self, "_h", None):
            self._L.c3_fastq_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
