"""Multi-GPU MSM: the one place the path shards.

The terms of a large MSM are partitioned by rank (one process per GPU); every rank runs the full local pipeline on
its slice; the only exchange is an all-gather of the 144-byte Jacobian partial results -- NCCL has no reduction over
curve points, so the "all-reduce of partial sums" is gather + a local add kernel.  Blob batches and NTTs do not shard
(independent units: replicas only)."""
import numpy as np

__all__ = ["shard_bounds", "ShardedMsm"]


def shard_bounds(n: int, rank: int, world: int):
    """contiguous, balanced slice [lo, hi) of n terms for `rank` of `world` (first n % world ranks get one more)"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class ShardedMsm:
    """Fixed-base MSM over `world` GPUs.  Each rank passes ITS slice of the bases at construction and ITS slice of
    the scalars per call; every rank gets the full result.

    local_msm(points_slice) -> object with .mult_device(out_ptr, n, scalars_ptr, batch, stream)
    all_gather(dst, src): torch.distributed.all_gather_into_tensor (or a stand-in)
    g1_sum(out_ptr, points_ptr, n, stream): device add of the gathered partials
    The three callables are injected so the host logic can be exercised on CPU (gloo) with stand-ins."""

    def __init__(self, local_handle, rank, world, all_gather, g1_sum, alloc):
        self.h, self.rank, self.world = local_handle, rank, world
        self.all_gather, self.g1_sum = all_gather, g1_sum
        self.partial = alloc(18)
        self.gathered = alloc(18 * world)
        self.total = alloc(18)

    def mult(self, scalars_ptr, n_local, stream=0):
        self.h.mult_device(self.partial.data_ptr(), n_local, scalars_ptr, 1, stream)
        if self.world == 1:
            return self.partial
        self.all_gather(self.gathered, self.partial)
        self.g1_sum(self.total.data_ptr(), self.gathered.data_ptr(), self.world, stream)
        return self.total
