#!/usr/bin/env python3
"""MSM 2^logn on the current environment's engine settings (B200_ACC_KARA, B200_MSM_C, ...): whole-MSM and accumulate-kernel
time (CUDA events) + parity against the folded-scalar oracle.  One JSON line.   python scripts/acc_time.py [logn] [tag]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import rust_kzg_b200 as B  # noqa: E402
import bench  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
tag = sys.argv[2] if len(sys.argv) > 2 else ""
K, s, L = bench.load_bases()
n = 1 << logn
rng = np.random.default_rng(bench.SEED)
sc = bench.rand_fr(rng, n)
pts = np.tile(L, (n // 4096, 1))
exp = K.p1_compress(bench.folded_expectation(K, L, sc, os.cpu_count() or 1))
d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
h = B.PreparedMsm(pts)
ms = bench.timed_events(torch, lambda: h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0), reps=20, warm=5)
ok = K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == exp
h.set_profiling(True)
for _ in range(10):
    h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
torch.cuda.synchronize()
acc_ms, runs = h.profile_read()
print(json.dumps({"tag": tag, "logn": logn, "ms": ms, "accumulate_ms": acc_ms / max(runs, 1), "parity_ok": bool(ok), "info": h.info(),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("B200_")}}))
