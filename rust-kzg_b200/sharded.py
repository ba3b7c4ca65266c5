"""Multi-GPU MSM: the one place the path shards (BASELINE configs[4]).

The terms of a large MSM are partitioned by rank (one process per GPU); every rank runs the full local pipeline on
its slice; the only exchange is an all-gather of the 144-byte Jacobian partial results -- NCCL has no reduction over
curve points, so the "all-reduce of partial sums" is gather + a local add kernel.  The exchange lives in the library
(b200_msm_sharded_*, include/b200_kzg.h): local MSM, ncclAllGather and the add are enqueued on one stream.  This class is
the Python-side stub: it ships rank 0's NCCL unique id to the other ranks through a caller-supplied broadcast (any
transport: torch.distributed, MPI, a file) and forwards the calls.  Blob batches and NTTs do not shard (independent
units: replicas only)."""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ["shard_bounds", "ShardedMsm"]


def shard_bounds(n: int, rank: int, world: int):
    """contiguous, balanced slice [lo, hi) of n terms for `rank` of `world` (first n % world ranks get one more)"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


class _LibBackend:
    """the product path: libb200kzg.so"""

    @staticmethod
    def _L():
        from . import lib
        return lib()

    def unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        if self._L().b200_msm_sharded_unique_id(buf) != 0:
            raise _lib.B200Error("b200_msm_sharded_unique_id failed (NCCL not loadable in this process?)")
        return bytes(buf)

    def prepare(self, pts, rank, world, uid):
        idb = (C.c_uint8 * 128).from_buffer_copy(uid)
        h = self._L().b200_msm_sharded_prepare(pts.ctypes.data_as(C.c_void_p), pts.shape[0], rank, world, idb)
        if not h:
            raise _lib.B200Error("b200_msm_sharded_prepare failed (no CUDA device, NCCL or memory); no CPU fallback")
        return h

    def mult(self, h, sc):
        out = np.zeros(18, np.uint64)
        _lib.check(self._L().b200_msm_sharded_mult(h, out.ctypes.data_as(C.c_void_p), sc.shape[0], sc.ctypes.data_as(C.c_void_p)))
        return out

    def mult_device(self, h, out_ptr, n_local, scalars_ptr, stream):
        _lib.check(self._L().b200_msm_sharded_mult_device(h, C.c_void_p(out_ptr), n_local, C.c_void_p(scalars_ptr), C.c_void_p(stream)))

    def local(self, h):
        return self._L().b200_msm_sharded_local(h)

    def free(self, h):
        self._L().b200_msm_sharded_free(h)


class ShardedMsm:
    """Fixed-base MSM over `world` GPUs, one process per GPU.  Each rank passes ITS slice of the bases at construction and
    ITS slice of the scalars per call; every rank gets the full result.

    broadcast(b: bytes) -> bytes: returns rank 0's argument on every rank (only called when world > 1).
    backend: the library by default; the CPU tests inject a stand-in with the same five methods."""

    def __init__(self, local_affine_points, rank, world, broadcast=None, backend=None):
        self.be = backend or _LibBackend()
        self.rank, self.world = rank, world
        pts = np.ascontiguousarray(np.asarray(local_affine_points, dtype=np.uint64).reshape(-1, 12))
        self.n_local = pts.shape[0]
        uid = bytes(128)
        if world > 1:
            if broadcast is None:
                raise ValueError("world > 1 needs a broadcast callable for the NCCL unique id")
            uid = broadcast(self.be.unique_id() if rank == 0 else bytes(128))
            if len(uid) != 128:
                raise ValueError("broadcast must return the 128-byte id of rank 0")
        self.uid = uid
        self.h = self.be.prepare(pts, rank, world, uid)

    def mult(self, local_scalars):
        """host scalars of this rank's slice -> the full sum (Jacobian, 18 u64) on every rank"""
        sc = np.ascontiguousarray(np.asarray(local_scalars, dtype=np.uint64).reshape(-1, 4))
        return self.be.mult(self.h, sc)

    def mult_device(self, out_ptr, n_local, scalars_ptr, stream=0):
        """device pointers, asynchronous on `stream`"""
        self.be.mult_device(self.h, out_ptr, n_local, scalars_ptr, stream)

    def local_handle(self):
        """the rank-local prepared MSM handle (for b200_msm_info / profiling)"""
        return self.be.local(self.h)

    def close(self):
        if self.h:
            self.be.free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
