"""world_size-2 gloo test of the multi-GPU MSM host logic (sharding, all-gather of 144-byte partials, local add).
The device pieces are replaced by oracle stand-ins so it runs on CPU; the GPU path uses the same ShardedMsm class."""
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

    class Handle:                                       # stand-in for PreparedMsm over this rank's slice
        def mult_device(self, out_ptr, n_local, scalars_ptr, batch, stream):
            res = K.msm_affine(pts[lo:hi], sc[lo:hi]) if hi > lo else np.zeros(18, np.uint64)
            holder["partial"].copy_(torch.from_numpy(res.view(np.int64)))

    holder = {}

    def alloc(k):
        t = torch.zeros(k, dtype=torch.int64)
        if "partial" not in holder:
            holder["partial"] = t
        return t

    def g1_sum(out_ptr, pts_ptr, k, stream):
        g = sm.gathered.numpy().view(np.uint64).reshape(k, 18)
        acc = np.zeros(18, np.uint64)
        for i in range(k):
            acc = K.p1_add(acc, g[i])
        sm.total.copy_(torch.from_numpy(acc.view(np.int64)))

    sm = B.ShardedMsm(Handle(), rank, world, dist.all_gather_into_tensor, g1_sum, alloc)
    total = sm.mult(0, hi - lo)
    full = K.msm_affine(pts, sc, nthreads=2)
    q.put((rank, K.p1_compress(total.numpy().view(np.uint64)) == K.p1_compress(full), (lo, hi)))
    dist.destroy_process_group()


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
