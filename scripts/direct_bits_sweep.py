"""Window-width sweep of the direct-lookup tables (B200_BLOB_DIRECT_BITS, B200_FK20_DIRECT_BITS): load time, commitment and
proof batches of 1..64 blobs, cell proofs of 64 blobs; every configuration is compared byte-for-byte with the first one and
three commitments with the oracle.  Run under gpurun:  python scripts/direct_bits_sweep.py 8 11 13"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rust_kzg_b200 as B  # noqa: E402
from blob_window_sweep import timed  # noqa: E402


def main():
    widths = [int(a) for a in sys.argv[1:]] or [8, 11, 13]
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
    out = {}
    for c in widths:
        os.environ["B200_BLOB_DIRECT_BITS"] = str(c)
        os.environ["B200_FK20_DIRECT_BITS"] = str(c)
        t0 = time.perf_counter()
        ts = B.KZGSettings.load_trusted_setup_file()
        load_s = time.perf_counter() - t0
        row = {"load_s": round(load_s, 3), "free_gb_after_load": round(torch.cuda.mem_get_info()[0] / 2**30, 1)}
        for n in (1, 2, 4, 8, 16, 32, 64):
            row[f"commit_{n}_ms"] = round(timed(lambda: ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), n,
                                                                                         d_st.data_ptr(), 0)), 4)
        comm = d_out.cpu().numpy().copy()
        for n in (1, 16, 64):
            row[f"proof_{n}_ms"] = round(timed(lambda: ts.compute_kzg_proof_device(d_out.data_ptr(), d_y.data_ptr(), d_blobs.data_ptr(),
                                                                                   d_z.data_ptr(), n, d_st.data_ptr(), 0, 0)), 4)
        proofs = d_out.cpu().numpy().copy()
        t0 = time.perf_counter()
        cp1 = ts.compute_cell_proofs_batch(blobs[:1])
        row["fk_first_call_s"] = round(time.perf_counter() - t0, 3)
        cp = ts.compute_cell_proofs_batch(blobs)
        t0 = time.perf_counter()
        for _ in range(3):
            ts.compute_cell_proofs_batch(blobs)
        row["cell_proofs_64_ms"] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
        row["free_gb_after_fk"] = round(torch.cuda.mem_get_info()[0] / 2**30, 1)
        row["oracle_ok"] = all(comm[i].tobytes() == exp_c[i] for i in exp_c)
        cur = (comm.tobytes(), proofs.tobytes(), np.asarray(cp).tobytes(), np.asarray(cp1).tobytes())
        if ref is None:
            ref = cur
        row["same_as_first"] = cur == ref
        out[str(c)] = row
        print(c, json.dumps(row), flush=True)
        ts.free()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_direct_bits_sweep.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
