"""Device-resident fft_fr timings over the size sweep (CUDA events, 20 reps after 5 warm-ups); prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B

rng = np.random.default_rng(3)
fs = B.FFTSettings(20)
out = {}
for logn in (9, 10, 11, 12, 13, 14, 16, 18, 20):
    m = 1 << logn
    a = rng.integers(0, 1 << 62, size=(m, 4), dtype=np.uint64)
    d_in = torch.from_numpy(a.view(np.int64)).cuda()
    d_o = torch.zeros_like(d_in)
    for _ in range(5):
        fs.fft_fr_device(d_o.data_ptr(), d_in.data_ptr(), m, False, 1, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fs.fft_fr_device(d_o.data_ptr(), d_in.data_ptr(), m, False, 1, 0)
    e1.record()
    torch.cuda.synchronize()
    out["2^%d" % logn] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
print(json.dumps({"variant": os.environ.get("B200_NTT_VARIANT", "0"), "us": out}))
