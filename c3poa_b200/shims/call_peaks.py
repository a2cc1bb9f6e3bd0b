"""Drop-in for bin/call_peaks.py: call_peaks(scores, min_dist, iters, window, order) on the GPU."""
import numpy as np

from ..api import GpuConsensus

_GPU = None


def call_peaks(scores, min_dist, iters, window, order):
    global _GPU
    if _GPU is None:
        _GPU = GpuConsensus(0)
    s = np.ascontiguousarray(scores, dtype=np.int32)
    r = _GPU.peaks_batch(s, np.array([0, s.size], dtype=np.int64), min_dist=int(min_dist), iters=int(iters),
                         window=int(window), order=int(order))
    n = int(r["n_peaks"][0])
    if n < 0:
        raise RuntimeError(f"GPU call_peaks failed with status {n}")
    return [] if n == 0 else r["peaks"][0, :n].astype(np.int64)
