"""snapatac2_b200 -- B200-native matrix-free spectral embedding.

Drop-in for one hot path of SnapATAC2: ``snap.tl.spectral`` with
``distance_metric='cosine'`` (``snapatac2.tools._embedding.spectral`` ->
``internal.spectral_embedding``).  Hand-written sm_100a CUDA kernels behind a
C ABI (``include/snapb200.h``), driven through ctypes.  No CPU fallback.
``pp.knn`` is the consumer next to the path: the exact neighbour graph of ``X_spectral``.
"""

from . import tl, pp, synth, dist      # noqa: F401
from ._adata import MiniAnnData       # noqa: F401
from .engine import Engine            # noqa: F401

__all__ = ["tl", "pp", "synth", "dist", "MiniAnnData", "Engine"]
__version__ = "0.1.0"
