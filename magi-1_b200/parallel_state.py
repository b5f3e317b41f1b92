"""Minimal context-parallel group holder: the three accessors the hot path uses from the reference's Megatron-style
inference/infra/distributed/parallel_state.py (get_cp_group :499, get_cp_world_size :621, get_cp_rank :629).
The rest of that module (tp/pp/dp groups) is outside the hot path."""
from typing import Optional

import torch.distributed as dist

_CP_GROUP = None
_CP_RANKS = None


def initialize_model_parallel(cp_size: Optional[int] = None, backend: Optional[str] = None) -> None:
    """All ranks of the job form one context-parallel group (the reference asserts cp * pp == world, pp = 1 on this
    path: dist_utils.py:48)."""
    global _CP_GROUP, _CP_RANKS
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    world = dist.get_world_size()
    if cp_size is not None and cp_size != world:
        raise ValueError(f"cp_size ({cp_size}) must equal the world size ({world}) on this path")
    _CP_RANKS = list(range(world))
    _CP_GROUP = dist.new_group(_CP_RANKS, backend=backend) if backend else dist.group.WORLD


def set_cp_group(group, ranks=None) -> None:
    global _CP_GROUP, _CP_RANKS
    _CP_GROUP, _CP_RANKS = group, ranks


def destroy_model_parallel() -> None:
    global _CP_GROUP, _CP_RANKS
    _CP_GROUP, _CP_RANKS = None, None


def get_cp_group():
    return _CP_GROUP


def get_cp_world_size() -> int:
    if _CP_GROUP is None or not dist.is_initialized():
        return 1
    return dist.get_world_size(group=_CP_GROUP)


def get_cp_rank() -> int:
    if _CP_GROUP is None or not dist.is_initialized():
        return 0
    return dist.get_rank(group=_CP_GROUP)


def get_tp_world_size(with_context_parallel: bool = False) -> int:
    """tp_size is always 1 in the reference (dist_utils.py:53-59), so the "tp-cp" size is the cp size
    (dit_module.py:794)."""
    return get_cp_world_size() if with_context_parallel else 1


def cp_global_rank(cp_rank: int) -> int:
    """Global rank of a member of the cp group (needed by point-to-point ops of the ring-KV exchange)."""
    if _CP_RANKS is not None:
        return _CP_RANKS[cp_rank]
    return cp_rank
