"""bmnas -- B200-native engine behind the BM-NAS ``models.search.darts`` drop-in surface.

Host side: launch plans (program.py), autograd glue (runtime.py), fused optimiser
(optim.py), search-step driver with CUDA graphs and NCCL data parallelism (search.py).
Device side: libbmnas_b200.so (csrc/*.cu) reached through the C ABI in
include/bmnas_b200.h.  There is no CPU or library fallback.
"""
from . import native, rng  # noqa: F401
from .rng import manual_seed  # noqa: F401

__all__ = ['native', 'rng', 'manual_seed']
