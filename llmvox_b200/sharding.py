"""Session-parallel sharding across GPUs: sessions never interact, so session i lives on rank
i mod world and no collective ever touches the data path.  Pure host logic."""
from __future__ import annotations

from typing import List, Sequence, TypeVar

T = TypeVar("T")


def shard_indices(n_items: int, world_size: int, rank: int) -> List[int]:
    """Indices of the items rank `rank` owns: round-robin, so that loads differ by at most one item."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_items, world_size))


def shard(items: Sequence[T], world_size: int, rank: int) -> List[T]:
    return [items[i] for i in shard_indices(len(items), world_size, rank)]


def owner(item_index: int, world_size: int) -> int:
    return item_index % world_size


def unshard(per_rank: Sequence[Sequence[T]]) -> List[T]:
    """Inverse of `shard` given every rank's list (e.g. after a host-side gather of results)."""
    world = len(per_rank)
    n = sum(len(p) for p in per_rank)
    out: List[T] = [None] * n  # type: ignore[list-item]
    for r, part in enumerate(per_rank):
        for k, v in enumerate(part):
            out[r + k * world] = v
    return out
