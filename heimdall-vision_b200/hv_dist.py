"""Multi-GPU host logic: frames are independent (detection.rs has no temporal state), so batches and camera streams
shard across ranks with NO collective on the data path.  The only exchange is the all-reduce of the 32 x u64 line
statistics vector (hv_line_stats: frames inspected / rejected, defects, area histogram; dashboard.py:38-46,
heimdall/core/system.py:168-175).  One process per GPU; torch.distributed is the plumbing (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

STATS_WORDS = 32
STATS_FIELDS = ["frames_inspected", "frames_rejected", "total_defects", "total_components", "total_defect_area",
                "total_fg_pixels"] + [f"area_hist_{i}" for i in range(16)] + ["capacity_errors"]


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n_items for `rank` (contiguous slices of ceil(N/G) frames)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    per = -(-n_items // world)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def stream_owner(stream_id: int, world: int) -> int:
    """Camera-stream affinity: stream s is always handled by GPU s mod G (no frame is ever split across GPUs)."""
    return stream_id % world


def streams_of(rank: int, world: int, n_streams: int) -> List[int]:
    return [s for s in range(n_streams) if stream_owner(s, world) == rank]


class CudaArrayView:
    """Wraps a raw device pointer (hv_stats_device_ptr) so torch can alias it: torch.as_tensor(view, device=...)."""

    def __init__(self, ptr: int, n_words: int = STATS_WORDS):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2}


def allreduce_line_stats(local_stats, group=None, async_op: bool = False):
    """Sum the per-rank statistics vectors in place (int64 tensor of STATS_WORDS, any device).  u64 counters are
    reduced as int64: identical bit patterns as long as the totals stay below 2^63."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(local_stats, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def stats_dict(vec) -> dict:
    v = [int(x) for x in vec[:len(STATS_FIELDS)]]
    d = dict(zip(STATS_FIELDS, v))
    d["defect_rate"] = d["frames_rejected"] / d["frames_inspected"] if d["frames_inspected"] else 0.0
    return d
