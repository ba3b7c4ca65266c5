"""Window-width sweep of the prepared fixed-base MSM at 2^LOGN terms (B200_MSM_C), parity-checked against the
folded-scalar oracle before timing.  Run under gpurun:  python scripts/msm_window_sweep.py 20 16 18 19 20 21"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rust_kzg_b200 as B  # noqa: E402


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    cs = [int(x) for x in sys.argv[2:]] or [16, 18, 19, 20]
    n = 1 << logn
    K, _, L = bench.load_bases()
    rng = np.random.default_rng(bench.SEED)
    sc = bench.rand_fr(rng, n)
    exp = K.p1_compress(bench.folded_expectation(K, L, sc, os.cpu_count() or 1))
    pts = np.tile(L, (n // 4096, 1))
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    res = {}
    for c in cs:
        os.environ["B200_MSM_C"] = str(c)
        msm = B.PreparedMsm(pts)
        for _ in range(3):
            msm.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
        torch.cuda.synchronize()
        ok = K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == exp
        msm.set_profiling(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            msm.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
        e1.record()
        torch.cuda.synchronize()
        acc_ms, runs = msm.profile_read()
        info = msm.info()
        res[c] = {"ms": e0.elapsed_time(e1) / 10, "accumulate_ms": acc_ms / max(runs, 1), "parity_ok": bool(ok), **info}
        print(c, res[c], flush=True)
        msm.close()
        del msm
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"logn": logn, "results": res}, open(os.path.join(ROOT, "gpurun_out", "msm_window_sweep_%d.json" % logn), "w"), indent=1)


if __name__ == "__main__":
    main()
