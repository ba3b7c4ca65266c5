"""fft_g1 of size 128 (the FK20 transform) over batches of 8..256 transforms, plain and fused double stages: is a stage
latency-bound (flat in the batch) or throughput-bound (linear)?  Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_kzg_b200 as B  # noqa: E402
from oracle import c_oracle as K  # noqa: E402
from blob_window_sweep import timed  # noqa: E402

text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
pts = np.ascontiguousarray(K.KZGSettings(text).g1_lagrange_brp, dtype=np.uint64).reshape(-1, 18)
fs = B.FFTSettings(7)
n = 128
res = {}
for batch in (1, 2, 8, 16, 32, 64, 128, 256):
    src = np.tile(pts, (max(1, batch * n // len(pts)), 1))[:batch * n]
    d_in = torch.from_numpy(src.view(np.int64)).cuda()
    d_out = torch.zeros_like(d_in)
    row = {}
    for fuse in (("0", "2", "3", "6") if batch <= 8 else ("0", "2", "3")):
        for split in ("0", "1"):
            os.environ["B200_FFT_G1_FUSE"] = fuse
            os.environ["B200_FFT_G1_SPLIT"] = split
            row["fuse" + fuse + ("_split" if split == "1" else "") + "_ms"] = round(
                timed(lambda: fs.fft_g1_device(d_out.data_ptr(), d_in.data_ptr(), n, False, batch, 0), reps=3, warm=1), 3)
    del os.environ["B200_FFT_G1_FUSE"], os.environ["B200_FFT_G1_SPLIT"]
    row["auto_ms"] = round(timed(lambda: fs.fft_g1_device(d_out.data_ptr(), d_in.data_ptr(), n, False, batch, 0), reps=3, warm=1), 3)
    res[batch] = row
    print(batch, json.dumps(row), flush=True)
