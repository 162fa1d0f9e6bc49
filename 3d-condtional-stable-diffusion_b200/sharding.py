"""Multi-GPU sharding of the sampling path: independent sample batches, one process per GPU, NO collective on the
data path.  Samples are numbered globally; the Philox counter carries the global sample index, so the
latents of sample i do not depend on how many GPUs the job uses."""
from __future__ import annotations

import os


def partition(total: int, world: int, rank: int):
    """Contiguous slice of ``total`` samples owned by ``rank``: (start, count); the first total % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def dist_env():
    """(world, rank, local_rank) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_context(context_ids, world: int, rank: int):
    """Slice a per-sample context-id sequence for this rank."""
    start, count = partition(len(context_ids), world, rank)
    return list(context_ids[start:start + count]), start


def generate_sharded(model, total_samples: int, latent_shape, context_ids=None, seed: int = 1234, world=None, rank=None, **kw):
    """Run model.generate on this rank's slice of ``total_samples``; returns (latents, first_global_sample_index).
    No communication: callers gather on the host if they need all volumes in one place."""
    w, r, _ = dist_env()
    world = w if world is None else world
    rank = r if rank is None else rank
    start, count = partition(total_samples, world, rank)
    if count == 0:
        return None, start
    ctx = None
    if context_ids is not None:
        ctx = list(context_ids[start:start + count])
    shape = (count,) + tuple(latent_shape)
    lat = model.generate(shape, seed=seed, sample_id0=start, **({"context": ctx} if ctx is not None else {}), **kw)
    return lat, start
