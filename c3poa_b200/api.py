"""Host-side API over the C ABI: batch containers and the GpuConsensus handle.

Mirrors the reference's three call sites batch-wise (see include/c3poa_gpu.h):
  conk.conk            -> GpuConsensus.conk_batch        (C3POa.py:123)
  call_peaks           -> GpuConsensus.peaks_batch       (C3POa.py:124)
  msa_aligner().msa    -> GpuConsensus.poa_batch         (bin/determine_consensus.py:30-47)
  analyze_reads body   -> GpuConsensus.consensus_batch   (C3POa.py:112-165, pre-racon)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import PoaParams, RESULT_DTYPE, Timings


def sg_coeffs(window: int = 41, order: int = 2) -> np.ndarray:
    """Savitzky-Golay coefficients exactly as the reference computes them
    (/root/reference/bin/savitzky_golay.py:27-31, deriv=0, rate=1)."""
    half = (window - 1) // 2
    b = np.asmatrix([[k ** i for i in range(order + 1)] for k in range(-half, half + 1)])
    return np.ascontiguousarray(np.linalg.pinv(b).A[0], dtype=np.float64)


def default_poa_params(**kw) -> PoaParams:
    p = PoaParams()
    _lib.load().c3_default_poa_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _pack(strs):
    """list[str|bytes] -> (uint8 blob, int64 offsets)."""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in strs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    blob = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return blob, off


@dataclass
class ReadBatch:
    """Packed reads + strand-resolved splints, ready for the device."""
    blob: np.ndarray          # uint8 ASCII
    off: np.ndarray           # int64 [n+1]
    sp_blob: np.ndarray       # uint8 ASCII
    sp_off: np.ndarray        # int32 [n_splints+1]
    sp_idx: np.ndarray        # int32 [n]

    @property
    def n(self) -> int:
        return self.off.size - 1

    @classmethod
    def from_strings(cls, seqs, splints, splint_idx):
        blob, off = _pack(seqs)
        sp_blob, sp_off = _pack(splints)
        return cls(blob, off, sp_blob, sp_off.astype(np.int32), np.ascontiguousarray(splint_idx, dtype=np.int32))

    def seq(self, i: int) -> str:
        return self.blob[self.off[i]:self.off[i + 1]].tobytes().decode()


class PinnedArray:
    """numpy view over page-locked host memory owned by the library."""

    def __init__(self, shape, dtype):
        self._L = _lib.load()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._p = self._L.c3_host_alloc(max(n, 1))
        if not self._p:
            raise MemoryError("c3_host_alloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._p:
                self.array = None
                self._L.c3_host_free(self._p)
                self._p = None
        except Exception:
            pass


class GpuError(RuntimeError):
    pass


class GpuConsensus:
    """One handle per device (not thread-safe; use one per host thread/GPU)."""

    POA_MODES = {"auto": 0, "warp": 1, "lane": 2, "grp": 3}

    def __init__(self, device: int = 0, poa_mode: str = "auto"):
        self._L = _lib.load()
        h = C.c_void_p()
        rc = self._L.c3_init(device, C.byref(h))
        if rc != 0:
            raise GpuError(f"c3_init(device={device}) failed with {rc}: no usable CUDA device "
                           "(the GPU stages have no CPU fallback)")
        self._h = h
        self.device = device
        if poa_mode != "auto":
            self.set_poa_mode(poa_mode)

    def set_poa_mode(self, mode: str):
        """POA kernel choice: 'auto' (the group path -- 8 lanes per read for the DP, one thread per read for the graph phases
        -- for the reads with a mean subread length <= 2 600 and <= 32 subreads when a batch holds >= 12 000 of them, the
        warp-per-read kernel for the rest), 'warp' (warp-per-read kernel only), 'lane' (round 1's thread-per-read kernel
        whenever eligible), 'grp' (the group path whenever eligible).  Same results."""
        self._ck(self._L.c3_set_poa_mode(self._h, self.POA_MODES[mode]), "c3_set_poa_mode")

    def set_abpoa_switches(self, int8_lanes: bool = False, end_clamp: bool = False):
        """Named switches of the abPOA restatement; both off by default."""
        self._ck(self._L.c3_set_abpoa_switches(self._h, int(int8_lanes), int(end_clamp)), "c3_set_abpoa_switches")

    def lane_counts(self):
        """(reads given to the lane kernel, reads it finished) of the last poa_batch / run."""
        a, b = C.c_int32(), C.c_int32()
        self._ck(self._L.c3_lane_counts(self._h, C.byref(a), C.byref(b)), "c3_lane_counts")
        return a.value, b.value

    def close(self):
        if getattr(self, "_h", None):
            self._L.c3_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise GpuError(f"{what} failed ({rc}): {self._L.c3_last_error(self._h).decode()}")

    def timings(self) -> dict:
        t = Timings()
        self._L.c3_get_timings(self._h, C.byref(t))
        return {k: getattr(t, k) for k, _ in Timings._fields_}

    def int_peak_ops(self) -> float:
        v = C.c_double()
        self._ck(self._L.c3_measure_int_peak(self._h, C.byref(v)), "c3_measure_int_peak")
        return v.value

    # ---- B1 ----
    def conk_batch(self, batch: ReadBatch, penalty: int = 20):
        """Returns the int32 profile blob (CSR with batch.off)."""
        prof = np.empty(int(batch.off[-1]), dtype=np.int32)
        self._ck(self._L.c3_conk_batch(self._h, batch.n, batch.blob.ctypes.data, batch.off.ctypes.data,
                                       batch.sp_off.size - 1, batch.sp_blob.ctypes.data, batch.sp_off.ctypes.data,
                                       batch.sp_idx.ctypes.data, penalty, prof.ctypes.data), "c3_conk_batch")
        return prof

    # ---- f-3: splint assignment ----
    def assign_splints(self, blob: np.ndarray, off: np.ndarray, candidates, penalty: int = 20):
        """candidates: list[str] (every splint in both orientations).  Returns (best int32[n], scores int32[c, n])."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = off.size - 1
        cb, coff = _pack(candidates)
        coff = coff.astype(np.int32)
        best = np.zeros(n, dtype=np.int32)
        scores = np.zeros((len(candidates), n), dtype=np.int32)
        self._ck(self._L.c3_assign_splints(self._h, n, blob.ctypes.data, off.ctypes.data, len(candidates),
                                           cb.ctypes.data, coff.ctypes.data, penalty, best.ctypes.data,
                                           scores.ctypes.data), "c3_assign_splints")
        return best, scores

    # ---- B2 ----
    def peaks_batch(self, prof: np.ndarray, off: np.ndarray, min_dist=500, iters=3, window=41, order=2,
                    coef=None, want_smoothed=False, max_peaks=64, height_mult=3.0, gate_mult=6.0):
        prof = np.ascontiguousarray(prof, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = off.size - 1
        coef = sg_coeffs(window, order) if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        peaks = np.zeros((n, max_peaks), dtype=np.int32)
        npk = np.zeros(n, dtype=np.int32)
        med = np.zeros(n, dtype=np.float64)
        sm = np.empty(prof.size, dtype=np.float64) if want_smoothed else None
        self._ck(self._L.c3_peaks_batch(self._h, n, prof.ctypes.data, off.ctypes.data, coef.ctypes.data, coef.size,
                                        iters, min_dist, height_mult, gate_mult,
                                        sm.ctypes.data if want_smoothed else None, med.ctypes.data,
                                        peaks.ctypes.data, max_peaks, npk.ctypes.data), "c3_peaks_batch")
        return dict(peaks=peaks, n_peaks=npk, median=med, smoothed=sm)

    # ---- B3 ----
    def poa_batch(self, groups, params: PoaParams | None = None, cons_cap: int | None = None, want_msa: bool = False):
        """groups: list of list[str].  Returns dict(cons=list[str], status, cells, nodes[, msa]).
        want_msa: groups of exactly two sequences return their two MSA rows (msa[g] = [row0, row1])
        instead of a consensus, as the reference's 2-repeat path asks abPOA for."""
        params = params or default_poa_params()
        flat = [s for g in groups for s in g]
        blob, off = _pack(flat)
        goff = np.zeros(len(groups) + 1, dtype=np.int32)
        goff[1:] = np.cumsum([len(g) for g in groups])
        n = len(groups)
        if cons_cap is None:
            cons_cap = int(max((len(s) for s in flat), default=1)) * (4 if want_msa else 2) + 64
        msa_cap = cons_cap // 2
        msa = np.zeros((len(flat), msa_cap), dtype=np.uint8) if want_msa else None
        msa_len = np.zeros(n, dtype=np.int32)
        cons = np.zeros((n, cons_cap), dtype=np.uint8)
        clen = np.zeros(n, dtype=np.int32)
        cells = np.zeros(n, dtype=np.int64)
        nodes = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        self._ck(self._L.c3_poa_batch(self._h, n, blob.ctypes.data, off.ctypes.data, goff.ctypes.data,
                                      C.byref(params), cons.ctypes.data, cons_cap, clen.ctypes.data,
                                      cells.ctypes.data, nodes.ctypes.data, status.ctypes.data,
                                      msa.ctypes.data if want_msa else None, msa_cap if want_msa else 0,
                                      msa_len.ctypes.data if want_msa else None), "c3_poa_batch")
        out = dict(cons=[cons[i, :clen[i]].tobytes().decode() for i in range(n)], status=status, cells=cells,
                   nodes=nodes)
        if want_msa:
            out["msa"] = [[msa[goff[g] + k, :msa_len[g]].tobytes().decode() for k in range(2)] if msa_len[g] > 0 else []
                          for g in range(n)]
        return out

    # ---- B4 ----
    def stage(self, batch: ReadBatch):
        self._ck(self._L.c3_stage(self._h, batch.n, batch.blob.ctypes.data, batch.off.ctypes.data,
                                  batch.sp_off.size - 1, batch.sp_blob.ctypes.data, batch.sp_off.ctypes.data,
                                  batch.sp_idx.ctypes.data), "c3_stage")
        self._n = batch.n

    def run(self, penalty=20, min_dist=500, iters=3, window=41, order=2, coef=None, params=None,
            max_peaks=64, cons_cap=4096):
        coef = sg_coeffs(window, order) if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        params = params or default_poa_params()
        self._ck(self._L.c3_run(self._h, penalty, coef.ctypes.data, coef.size, iters, min_dist, C.byref(params),
                                max_peaks, cons_cap), "c3_run")
        self._max_peaks, self._cons_cap = max_peaks, cons_cap

    def fetch(self, out=None):
        n, mp, cc = self._n, self._max_peaks, self._cons_cap
        if out is None:
            out = dict(peaks=np.zeros((n, mp), dtype=np.int32), sub_bounds=np.zeros((n, mp, 2), dtype=np.int32),
                       dang_bounds=np.zeros((n, 2, 2), dtype=np.int32), cons=np.zeros((n, cc), dtype=np.uint8),
                       results=np.zeros(n, dtype=RESULT_DTYPE))
        self._ck(self._L.c3_fetch(self._h, out["peaks"].ctypes.data, out["sub_bounds"].ctypes.data,
                                  out["dang_bounds"].ctypes.data, out["cons"].ctypes.data,
                                  out["results"].ctypes.data), "c3_fetch")
        return out

    def consensus_batch(self, batch: ReadBatch, out=None, penalty=20, min_dist=500, iters=3, window=41, order=2,
                        coef=None, params=None, max_peaks=64, cons_cap=4096):
        """B4 as one call of the fused C-ABI entry point `c3_consensus_batch` (host buffers in, host buffers out): what
        the reference's `for read in reads:` loop (C3POa.py:112-165) is replaced by.  stage / run / fetch are the same
        three steps for callers that keep a batch resident."""
        coef = sg_coeffs(window, order) if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        params = params or default_poa_params()
        n = batch.n
        if out is None:
            out = dict(peaks=np.zeros((n, max_peaks), dtype=np.int32), sub_bounds=np.zeros((n, max_peaks, 2), dtype=np.int32),
                       dang_bounds=np.zeros((n, 2, 2), dtype=np.int32), cons=np.zeros((n, cons_cap), dtype=np.uint8),
                       results=np.zeros(n, dtype=RESULT_DTYPE))
        self._ck(self._L.c3_consensus_batch(self._h, n, batch.blob.ctypes.data, batch.off.ctypes.data,
                                            batch.sp_off.size - 1, batch.sp_blob.ctypes.data, batch.sp_off.ctypes.data,
                                            batch.sp_idx.ctypes.data, penalty, coef.ctypes.data, coef.size, iters, min_dist,
                                            C.byref(params), max_peaks, cons_cap, out["peaks"].ctypes.data,
                                            out["sub_bounds"].ctypes.data, out["dang_bounds"].ctypes.data,
                                            out["cons"].ctypes.data, out["results"].ctypes.data), "c3_consensus_batch")
        self._n, self._max_peaks, self._cons_cap = n, max_peaks, cons_cap
        return out


def pairwise_rows(out, i):
    """The two MSA rows of a 2-repeat read (status 2, n_sub 2) from a fused-batch result."""
    L = int(out["results"]["cons_len"][i])
    row = out["cons"][i]
    return [row[:L].tobytes().decode(), row[L:2 * L].tobytes().decode()]
