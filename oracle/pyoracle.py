"""ctypes front-end of the CPU oracle (oracle/c3poa_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(c3poa_b200/) never imports this module.

Parity status: stage 2 / 3a are pinned by the reference's own Python (see
oracle/gen_golden.py and tests/golden/); conk and abPOA 1.0.5 are PARITY
UNPINNED (neither is vendored in /root/reference nor installable offline).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libc3poa_oracle.so")
    src = os.path.join(_HERE, "c3poa_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libc3poa_oracle.so"])
    return so


class PoaPara(C.Structure):
    _fields_ = [("match", C.c_int), ("mismatch", C.c_int), ("gap_open1", C.c_int), ("gap_ext1", C.c_int),
                ("gap_open2", C.c_int), ("gap_ext2", C.c_int), ("wb", C.c_int), ("wf", C.c_double),
                ("simd_bits", C.c_int), ("int8_lanes", C.c_int), ("end_clamp", C.c_int)]


class PoaStats(C.Structure):
    _fields_ = [("cells", C.c_int64), ("node_n", C.c_int32), ("n_aln", C.c_int32),
                ("last_score", C.c_int32), ("status", C.c_int32)]


class ReadResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_peaks", C.c_int32), ("n_sub", C.c_int32), ("n_dang", C.c_int32),
                ("cons_len", C.c_int32), ("pad", C.c_int32), ("poa_cells", C.c_int64), ("conk_cells", C.c_int64)]


RESULT_DTYPE = np.dtype([("status", "<i4"), ("n_peaks", "<i4"), ("n_sub", "<i4"), ("n_dang", "<i4"),
                         ("cons_len", "<i4"), ("poa_nodes", "<i4"), ("poa_cells", "<i8"), ("conk_cells", "<i8")])


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.c3o_conk.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
        L.c3o_savgol.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.c3o_call_peaks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.c3o_split.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p,
                                C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_int)]
        L.c3o_poa_default_para.argtypes = [C.POINTER(PoaPara)]
        L.c3o_poa_msa.argtypes = [C.POINTER(PoaPara), C.c_int, C.POINTER(C.c_char_p), C.c_void_p,
                                  C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_int,
                                  C.POINTER(C.c_int), C.POINTER(PoaStats), C.c_void_p]
        L.c3o_consensus_batch.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_char_p),
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_int, C.POINTER(PoaPara), C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


def sg_coeffs(window: int = 41, order: int = 2) -> np.ndarray:
    """The reference's coefficient line (bin/savitzky_golay.py:27-31), deriv=0, rate=1."""
    half = (window - 1) // 2
    b = np.asmatrix([[k ** i for i in range(order + 1)] for k in range(-half, half + 1)])
    return np.ascontiguousarray(np.linalg.pinv(b).A[0], dtype=np.float64)


def default_para(**kw) -> PoaPara:
    p = PoaPara()
    lib().c3o_poa_default_para(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def conk(splint: str, seq: str, penalty: int = 20) -> np.ndarray:
    out = np.empty(len(seq), dtype=np.int32)
    rc = lib().c3o_conk(splint.encode(), len(splint), seq.encode(), len(seq), penalty, out.ctypes.data)
    if rc:
        raise RuntimeError(f"c3o_conk rc={rc}")
    return out


def savgol(y: np.ndarray, coef: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(y)
    rc = lib().c3o_savgol(y.ctypes.data, y.size, coef.ctypes.data, coef.size, out.ctypes.data)
    if rc:
        raise RuntimeError(f"c3o_savgol rc={rc}")
    return out


def call_peaks(scores, min_dist=500, iters=3, window=41, order=2, coef=None, max_peaks=4096):
    """Returns (peaks int32[], smoothed float64[], median)."""
    scores = np.ascontiguousarray(scores, dtype=np.int32)
    coef = sg_coeffs(window, order) if coef is None else coef
    sm = np.empty(scores.size, dtype=np.float64)
    med = C.c_double()
    pk = np.empty(max_peaks, dtype=np.int32)
    n = lib().c3o_call_peaks(scores.ctypes.data, scores.size, min_dist, iters, coef.ctypes.data, coef.size,
                             sm.ctypes.data, C.byref(med), pk.ctypes.data, max_peaks)
    if n < 0:
        raise RuntimeError(f"c3o_call_peaks rc={n}")
    return pk[:n].copy(), sm, med.value


def split(peaks, ls: int, lr: int):
    """Returns (skip, shifted_peaks, sub_bounds[n,2], dang_bounds[m,2])."""
    pk = np.ascontiguousarray(peaks, dtype=np.int32).copy()
    n = C.c_int(pk.size)
    sb = np.zeros((max(pk.size, 1), 2), dtype=np.int32)
    db = np.zeros((2, 2), dtype=np.int32)
    ns, nd = C.c_int(), C.c_int()
    skip = lib().c3o_split(pk.ctypes.data, C.byref(n), ls, lr, sb.ctypes.data, C.byref(ns), db.ctypes.data, C.byref(nd))
    return bool(skip), pk[:n.value].copy(), sb[:ns.value].copy(), db[:nd.value].copy()


def poa_msa(seqs, out_cons=True, out_msa=False, para=None, debug=False):
    """Returns dict(cons=str, msa=list[str], stats=..., dbg=ndarray)."""
    para = para or default_para()
    n = len(seqs)
    arr = (C.c_char_p * n)(*[s.encode() for s in seqs])
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    tot = int(lens.sum()) + 16
    cons = np.zeros(tot, dtype=np.uint8)
    cl, ml = C.c_int(0), C.c_int(0)
    msa = np.zeros((n, tot), dtype=np.uint8) if out_msa else None
    st = PoaStats()
    dbg = np.zeros((max(n, 1), 4), dtype=np.int32)
    rc = lib().c3o_poa_msa(C.byref(para), n, arr, lens.ctypes.data,
                           cons.ctypes.data if out_cons else None, tot, C.byref(cl),
                           msa.ctypes.data if out_msa else None, tot, C.byref(ml), C.byref(st),
                           dbg.ctypes.data if debug else None)
    if rc:
        raise RuntimeError(f"c3o_poa_msa rc={rc}")
    return dict(cons=cons[:cl.value].tobytes().decode() if out_cons else "",
                msa=[msa[i, :ml.value].tobytes().decode() for i in range(n)] if out_msa else [],
                cells=st.cells, node_n=st.node_n, n_aln=st.n_aln, last_score=st.last_score, dbg=dbg[:st.n_aln])


def consensus_batch(seqs, splints, splint_idx, penalty=20, min_dist=500, iters=3, window=41, order=2,
                    para=None, max_peaks=64, cons_cap=None, n_threads=1):
    """Whole per-read path on the CPU (baseline leg).  seqs: list[str]; splints: list[str] already
    strand-resolved; splint_idx[i] indexes splints."""
    para = para or default_para()
    n = len(seqs)
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    blob = "".join(seqs).encode()
    sp = (C.c_char_p * len(splints))(*[s.encode() for s in splints])
    sl = np.array([len(s) for s in splints], dtype=np.int32)
    si = np.ascontiguousarray(splint_idx, dtype=np.int32)
    coef = sg_coeffs(window, order)
    if cons_cap is None:
        cons_cap = int(max(len(s) for s in seqs)) if n else 1
    peaks = np.zeros((n, max_peaks), dtype=np.int32)
    sb = np.zeros((n, max_peaks, 2), dtype=np.int32)
    db = np.zeros((n, 2, 2), dtype=np.int32)
    cons = np.zeros((n, cons_cap), dtype=np.uint8)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    rc = lib().c3o_consensus_batch(n, blob, off.ctypes.data, len(splints), sp, sl.ctypes.data, si.ctypes.data,
                                   penalty, min_dist, iters, coef.ctypes.data, coef.size, C.byref(para),
                                   max_peaks, peaks.ctypes.data, sb.ctypes.data, db.ctypes.data, cons_cap,
                                   cons.ctypes.data, res.ctypes.data, n_threads)
    return dict(rc=rc, results=res, peaks=peaks, sub_bounds=sb, dang_bounds=db, cons=cons)
