"""Per-rank work partitioning of the reference's batch-test drivers.

`split_list_into_chunks` mirrors the helper every driver defines and applies to its list of test pairs before starting
one process per GPU (/root/reference/stage2_batchtest_inpaint_model.py:25-31,266-285;
stage1_batchtest_prior_model.py:30-36,170-183; stage3_batchtest_refined_model.py): contiguous chunks of
`len(lst) // n` items, the remainder appended to the last chunk, order preserved; fewer items than ranks is an error
there (`range()` with step 0) and here.  Quirk kept: only ONE trailing chunk is merged, so for short lists with
`len % n >= len // n` the function returns more than `n` chunks and the drivers (which index `data_list[rank]`) skip the
extra items — `rank_shard` reproduces that.  Ranks never exchange data afterwards (no data-path collective)."""
from __future__ import annotations

from typing import List, Sequence, TypeVar

T = TypeVar("T")


def split_list_into_chunks(lst: Sequence[T], n: int) -> List[List[T]]:
    if n <= 0:
        raise ValueError("n must be positive")
    size = len(lst) // n
    if size == 0:
        raise ValueError("range() arg 3 must not be zero")      # the reference's failure for len(lst) < n
    chunks = [list(lst[i:i + size]) for i in range(0, len(lst), size)]
    if len(chunks) > n:                                         # ONE trailing chunk is merged, exactly as the reference does:
        tail = chunks.pop()                                     # when len % n >= len // n (short lists only) more than n
        chunks[-1].extend(tail)                                 # chunks remain and the drivers never run the extra ones
    return chunks


def rank_shard(lst: Sequence[T], rank: int, world_size: int) -> List[T]:
    """The items rank `rank` of `world_size` processes works on (what `data_list[rank]` is in the drivers)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    return split_list_into_chunks(lst, world_size)[rank]
