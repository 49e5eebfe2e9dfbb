"""Host-side logic of the multi-GPU runs: how the stream of read pairs is dealt to ranks / devices, and how per-rank device
times are combined.  Read pairs are independent, so there is no collective on the data path: every rank
(one process per GPU) trims its own contiguous shard of batches; the only exchange is the max-reduction of the timings.
Kept free of CUDA so that it can be tested with the gloo backend on CPU."""
from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    batches: tuple  # global batch indices owned by this rank, in processing order
    first_pair: tuple  # first global pair index of each owned batch


def shard_batches(n_batches_per_rank: int, pairs_per_batch: int, rank: int, world: int) -> Shard:
    """Weak scaling: every rank owns `n_batches_per_rank` batches; batch b of rank r is global batch r*n+b, i.e. ranks own
    disjoint contiguous slices of the stream and together cover [0, world*n*pairs_per_batch)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    batches = tuple(rank * n_batches_per_rank + b for b in range(n_batches_per_rank))
    return Shard(rank, world, batches, tuple(g * pairs_per_batch for g in batches))


def round_robin_device(slot: int, n_devices: int) -> int:
    """Device of a slot inside one process (the C ABI's rule: slot s runs on device_ids[s % n_devices])."""
    return slot % n_devices


def reduce_max(value: float, dist=None, device=None) -> float:
    """Max over ranks of a per-rank measurement (device-side elapsed time); identity without a process group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(pairs_per_rank: int, world: int, max_seconds: float) -> float:
    """Whole-job pairs/s: all ranks' pairs over the slowest rank's time."""
    return world * pairs_per_rank / max_seconds
