"""world_size-2 gloo test of the multi-GPU MSM host logic: sharding, the unique-id bootstrap (rank 0 creates, everyone
receives), gather of the 144-byte partials + local add.  The library backend is replaced by an oracle stand-in so it
runs on CPU; bench.py --gpus N and tests/test_gpu_sharded.py drive the same ShardedMsm class over libb200kzg.so."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rust_kzg_b200 as B
    from oracle import c_oracle as K
    text = open(os.path.join(ROOT, "rust-kzg_b200", "data", "trusted_setup.txt")).read()
    L = K.p1s_to_affine(K.KZGSettings(text).g1_lagrange_brp)
    rng = np.random.default_rng(7)                      # same stream on every rank: the global problem
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    pts = np.tile(L, (n // 4096 + 1, 1))[:n]
    lo, hi = B.shard_bounds(n, rank, world)

    class Backend:                                      # stand-in for libb200kzg.so: the oracle computes, gloo gathers
        def unique_id(self):
            assert rank == 0                            # only rank 0 may create the id
            return bytes(range(128))

        def prepare(self, p, r, w, uid):
            assert (r, w) == (rank, world) and np.array_equal(p, pts[lo:hi])
            return {"uid": uid}

        def mult(self, h, s):
            part = K.msm_affine(pts[lo:hi], s) if hi > lo else np.zeros(18, np.uint64)
            gathered = torch.zeros(18 * world, dtype=torch.int64)
            dist.all_gather_into_tensor(gathered, torch.from_numpy(part.view(np.int64).copy()))
            g = gathered.numpy().view(np.uint64).reshape(world, 18)
            acc = np.zeros(18, np.uint64)
            for i in range(world):
                acc = K.p1_add(acc, g[i])
            return acc

        def free(self, h):
            pass

    def bcast(b):
        box = [b]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    sm = B.ShardedMsm(pts[lo:hi], rank, world, broadcast=bcast, backend=Backend())
    total = sm.mult(sc[lo:hi])
    full = K.msm_affine(pts, sc, nthreads=2)
    q.put((rank, K.p1_compress(total) == K.p1_compress(full) and sm.uid == bytes(range(128)), (lo, hi)))
    sm.close()
    dist.destroy_process_group()


def test_sharded_msm_fails_loudly_without_a_device():
    """the product backend has no CPU path: preparing a shard without a GPU raises"""
    import rust_kzg_b200 as B
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.B200Error):
        B.ShardedMsm(np.zeros((8, 12), np.uint64), 0, 1)


@pytest.mark.parametrize("n", [5000, 4097])
def test_sharded_msm_world2(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    bounds = sorted(b for _, _, b in res)
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == n
