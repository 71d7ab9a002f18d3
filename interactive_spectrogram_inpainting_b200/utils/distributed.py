"""Rank helpers with the reference's semantics
(interactive_spectrogram_inpainting/utils/distributed.py:5-22) plus the note sharding used
by code extraction: every note goes to exactly one rank, nothing is padded or dropped."""
from typing import Tuple

import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized()


def is_master_process() -> bool:
    return not is_distributed() or dist.get_rank() == 0


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if is_distributed():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_size(total: int, rank: int, world_size: int) -> int:
    """Notes owned by ``rank``: total // world (+1 for the first total % world ranks) -- the
    count ``DistributedEvalSampler`` computes (utils/distributed.py:17-22)."""
    return total // world_size + int(rank < total % world_size)


def shard_indices(total: int, rank: int, world_size: int) -> range:
    """The strided shard ``rank, rank + world, ...`` of ``DistributedSampler(shuffle=False)``
    (extract_code.py:196-198) without its padding: no note is duplicated."""
    return range(rank, total, world_size)


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous alternative ``[start, stop)`` with the same per-rank counts; keeps a
    rank's notes adjacent in memory, which is what the synthetic benchmark uses."""
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + int(rank < extra)
