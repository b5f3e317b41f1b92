"""magi-1_b200: B200-native (sm_100a) implementation of MAGI-1's chunk-causal DiT attention hot path.

The directory name follows the repository; import it as `magi_b200` (see magi_b200.py at the repo root).
Everything here is the host-side mirror of the reference's operator interface for this path
(inference/model/dit/dit_module.py:776-1246, inference/infra/parallelism/context_parallel.py); the compute lives in
csrc/ (hand-written CUDA behind the C-ABI of include/magi_attn.h).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
