"""The library's multi-GPU MSM entry points (b200_msm_sharded_*, include/b200_kzg.h) on the GPU.  world = 1 exercises the
whole call path except the collective on any box; world = 2 (two processes, NCCL) runs when two devices are visible."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_world1_matches_oracle(B, K, lagrange_affine):
    rng = np.random.default_rng(41)
    n = 3 * 4096
    pts = np.tile(lagrange_affine, (3, 1))
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    sm = B.ShardedMsm(pts, 0, 1)
    want = K.p1_compress(K.msm_affine(pts, sc, nthreads=os.cpu_count() or 1))
    assert K.p1_compress(sm.mult(sc)) == want
    assert K.p1_compress(sm.mult(sc[:5000])) == K.p1_compress(K.msm_affine(pts[:5000], sc[:5000], nthreads=4))
    sm.close()


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import rust_kzg_b200 as B
    from oracle import c_oracle as K
    text = open(os.path.join(ROOT, "rust-kzg_b200", "data", "trusted_setup.txt")).read()
    L = K.p1s_to_affine(K.KZGSettings(text).g1_lagrange_brp)
    rng = np.random.default_rng(7)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    pts = np.tile(L, (n // 4096 + 1, 1))[:n]
    lo, hi = B.shard_bounds(n, rank, world)

    def bcast(b):
        t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    sm = B.ShardedMsm(pts[lo:hi], rank, world, broadcast=bcast)
    got = sm.mult(sc[lo:hi])
    d_sc = torch.from_numpy(sc[lo:hi].view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    sm.mult_device(d_out.data_ptr(), hi - lo, d_sc.data_ptr(), 0)
    torch.cuda.synchronize()
    full = K.p1_compress(K.msm_affine(pts, sc, nthreads=4))
    q.put((rank, K.p1_compress(got) == full and K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == full))
    sm.close()
    dist.destroy_process_group()


def test_sharded_world2_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 20000, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
