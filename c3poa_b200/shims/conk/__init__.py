"""Drop-in for the `conk` package used by the reference (`from conk import conk`; call site
/root/reference/C3POa.py:10,123): conk.conk(splint, seq, penalty) -> scores, on the GPU."""
from . import conk  # noqa: F401
