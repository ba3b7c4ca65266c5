"""Window width / segment-fold sweep of the 64-blob commitment and proof MSMs (B200_BLOB_C, B200_BLOB_FOLD); every
configuration is checked byte-for-byte against the first one and against the oracle on three blobs.
Arguments: commit_c:commit_fold[:proof_c:proof_fold] (B200_PROOF_C, B200_PROOF_FOLD; default 13:2).
Run under gpurun:  python scripts/blob_window_sweep.py 12:-1 12:2 13:2 13:3 14:3"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rust_kzg_b200 as B  # noqa: E402


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    cfgs = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(12, -1), (12, 2), (13, 2), (13, 3), (14, 3)]
    K, osettings, _ = bench.load_bases()
    osettings.set_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(bench.SEED)
    nb = 64
    blobs = bench.rand_blobs(rng, nb)
    zs = bench.rand_blobs(rng, 1)[0, :32 * nb].reshape(nb, 32).copy()
    d_blobs = torch.from_numpy(blobs).cuda()
    d_z = torch.from_numpy(zs).cuda()
    d_out = torch.zeros((nb, 48), dtype=torch.uint8, device="cuda")
    d_y = torch.zeros((nb, 32), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(nb, dtype=torch.int32, device="cuda")
    exp_c = {i: K.blob_to_kzg_commitment(blobs[i].tobytes(), osettings) for i in (0, 17, 63)}
    ref = None
    res = {}
    for cfg in cfgs:
        c, fold = cfg[0], cfg[1]
        pc, pfold = (cfg[2], cfg[3]) if len(cfg) >= 4 else (13, 2)     # c:fold[:proof_c:proof_fold]
        os.environ["B200_BLOB_C"] = str(c)
        os.environ["B200_BLOB_FOLD"] = str(fold)
        os.environ["B200_PROOF_C"] = str(pc)
        os.environ["B200_PROOF_FOLD"] = str(pfold)
        os.environ["B200_BLOB_C0"] = str(cfg[4] if len(cfg) >= 5 else 0)                # ...[:commit_c0]
        ts = B.KZGSettings.load_trusted_setup_file()
        ms_c = timed(lambda: ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), nb, d_st.data_ptr(), 0))
        comm = d_out.cpu().numpy().copy()
        ms_p = timed(lambda: ts.compute_kzg_proof_device(d_out.data_ptr(), d_y.data_ptr(), d_blobs.data_ptr(), d_z.data_ptr(), nb,
                                                         d_st.data_ptr(), 0, 0))
        proofs = d_out.cpu().numpy().copy()
        ok = all(comm[i].tobytes() == exp_c[i] for i in exp_c)
        if ref is None:
            ref = (comm, proofs)
        ok = ok and np.array_equal(comm, ref[0]) and np.array_equal(proofs, ref[1])
        key = ":".join(str(x) for x in cfg)
        res[key] = {"commit_ms": ms_c, "proof_ms": ms_p, "parity_ok": bool(ok)}
        print(key, res[key], flush=True)
        ts.free()
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "blob_window_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
