"""concurrentproject_b200 -- B200-native score-only Smith-Waterman/Gotoh behind the reference's
score-function entry points (algoGPU.h).  The product is concurrentproject_b200/lib/libswb200.so
(hand-written sm_100a CUDA, C ABI in include/); this package is its thin host-side mirror."""
from . import rng  # noqa: F401  (pure numpy, safe without the CUDA library)

__all__ = ["rng", "api", "fasta", "ring"]


def __getattr__(name):
    if name in ("api", "fasta", "ring"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
