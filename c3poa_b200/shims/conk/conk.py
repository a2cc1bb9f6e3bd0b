"""conk.conk(splint, seq, penalty) -> 1-D int32 array, one score per read offset (GPU, no CPU fallback)."""
import numpy as np

from ...api import GpuConsensus, ReadBatch

_GPU = None


def _gpu():
    global _GPU
    if _GPU is None:
        _GPU = GpuConsensus(0)
    return _GPU


def conk(splint: str, seq: str, penalty: int):
    b = ReadBatch.from_strings([seq], [splint], np.zeros(1, dtype=np.int32))
    return _gpu().conk_batch(b, penalty=int(penalty))
