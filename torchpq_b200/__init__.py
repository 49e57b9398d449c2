"""torchpq_b200 -- B200-native (sm_100a) IVFPQ search path behind torchpq.index.IVFPQIndex.search().

Importing this package loads ``libtpq_b200.so``; there is no CPU / PyTorch fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from .index import IVFPQIndex, ScanLayout
from . import fn
from .patch import patch, unpatch

__all__ = ["IVFPQIndex", "ScanLayout", "fn", "patch", "unpatch"]
