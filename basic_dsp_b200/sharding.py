"""Row sharding of batched work across the GPUs of one box (one process per GPU).

The hot path has no cross-row data flow (the reference's matrix type processes rows one after the
other, matrix/src/time_freq.rs:52-74), so a batch is split into contiguous row blocks, every rank
works on its own block with no collective on the data path, and only the *timing* is reduced
(max over ranks) through torch.distributed."""
from __future__ import annotations


def row_shard(total_rows: int, world: int, rank: int):
    """Contiguous block of rows owned by `rank`: the first (total_rows % world) ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(total_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """max-reduction of a per-rank scalar (elapsed time); identity when not running distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def weak_scaling_throughput(units_per_rank: int, steps: int, elapsed_s_max: float, world: int) -> float:
    """Whole-job units/s when every rank processed units_per_rank*steps units in elapsed_s_max."""
    return world * units_per_rank * steps / elapsed_s_max
