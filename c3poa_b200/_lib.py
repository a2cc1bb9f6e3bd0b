"""ctypes binding of libc3poa_gpu.so (the C ABI declared in include/c3poa_gpu.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device can
be initialised, importing/creating a handle raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("C3POA_GPU_LIB") or os.path.join(_HERE, "libc3poa_gpu.so")   # env: tuning variants only

EXPORTS = [
    "c3_version", "c3_device_count", "c3_init", "c3_destroy", "c3_last_error", "c3_default_poa_params",
    "c3_get_timings", "c3_conk_batch", "c3_peaks_batch", "c3_poa_batch", "c3_stage", "c3_run", "c3_fetch",
    "c3_consensus_batch", "c3_measure_int_peak", "c3_host_alloc", "c3_host_free",
    "c3_fastq_open", "c3_fastq_next", "c3_fastq_close", "c3_assign_splints", "c3_set_poa_mode", "c3_lane_counts",
    "c3_format_batch", "c3_set_abpoa_switches",
]


class PoaParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32), ("gap_ext1", C.c_int32),
                ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32), ("wb", C.c_int32), ("simd_bits", C.c_int32),
                ("wf", C.c_double)]


class Timings(C.Structure):
    _fields_ = [("encode_ms", C.c_float), ("conk_ms", C.c_float), ("peaks_ms", C.c_float), ("split_ms", C.c_float),
                ("poa_ms", C.c_float), ("total_ms", C.c_float), ("kernel_launches", C.c_int32),
                ("poa_items", C.c_int32), ("poa_dp_ms", C.c_float), ("poa_graph_ms", C.c_float),
                ("poa_warp_ms", C.c_float), ("poa_lane_ms", C.c_float), ("poa_dp_launches", C.c_int32),
                ("poa_graph_launches", C.c_int32)]


RESULT_DTYPE = np.dtype([("status", "<i4"), ("n_peaks", "<i4"), ("n_sub", "<i4"), ("n_dang", "<i4"),
                         ("cons_len", "<i4"), ("poa_nodes", "<i4"), ("poa_cells", "<i8")])
assert RESULT_DTYPE.itemsize == 32

_LIB = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises LibraryMissing (never falls back to a CPU path)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise LibraryMissing(f"{SO_PATH} not found: build it with `python -m c3poa_b200.build` "
                             "(there is no CPU fallback for the GPU stages)")
    L = C.CDLL(SO_PATH)
    vp, i32, i64p = C.c_void_p, C.c_int32, C.c_void_p
    L.c3_version.restype = C.c_char_p
    L.c3_device_count.restype = C.c_int
    L.c3_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.c3_destroy.argtypes = [vp]
    L.c3_destroy.restype = None
    L.c3_last_error.argtypes = [vp]
    L.c3_last_error.restype = C.c_char_p
    L.c3_default_poa_params.argtypes = [C.POINTER(PoaParams)]
    L.c3_default_poa_params.restype = None
    L.c3_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.c3_set_poa_mode.argtypes = [vp, i32]
    L.c3_set_abpoa_switches.argtypes = [vp, i32, i32]
    L.c3_lane_counts.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.c3_conk_batch.argtypes = [vp, i32, vp, i64p, i32, vp, vp, vp, i32, vp]
    L.c3_peaks_batch.argtypes = [vp, i32, vp, i64p, vp, i32, i32, i32, C.c_double, C.c_double, vp, vp, vp, i32, vp]
    L.c3_poa_batch.argtypes = [vp, i32, vp, i64p, vp, C.POINTER(PoaParams), vp, i32, vp, vp, vp, vp, vp, i32, vp]
    L.c3_stage.argtypes = [vp, i32, vp, i64p, i32, vp, vp, vp]
    L.c3_run.argtypes = [vp, i32, vp, i32, i32, i32, C.POINTER(PoaParams), i32, i32]
    L.c3_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
    L.c3_consensus_batch.argtypes = [vp, i32, vp, i64p, i32, vp, vp, vp, i32, vp, i32, i32, i32,
                                     C.POINTER(PoaParams), i32, i32, vp, vp, vp, vp, vp]
    L.c3_measure_int_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.c3_host_alloc.argtypes = [C.c_size_t]
    L.c3_host_alloc.restype = vp
    L.c3_host_free.argtypes = [vp]
    L.c3_host_free.restype = None
    L.c3_assign_splints.argtypes = [vp, i32, vp, i64p, i32, vp, vp, i32, vp, vp]
    L.c3_fastq_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.c3_fastq_next.argtypes = [vp, i32, C.c_int64, i32, vp, vp, vp, vp, C.c_int64, vp, vp, vp]
    L.c3_fastq_close.argtypes = [vp]
    L.c3_fastq_close.restype = None
    L.c3_format_batch.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32, vp, C.c_int64, vp, vp,
                                  C.c_int64, vp, vp]
    _LIB = L
    return L
