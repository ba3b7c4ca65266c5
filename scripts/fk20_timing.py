"""FK20 cell-proof batch timing (64 blobs, host numpy in / out); knobs come from the environment (B200_FK20_C ...)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B

rng = np.random.default_rng(1)
ts = B.KZGSettings.load_trusted_setup_file()
blobs = rng.integers(0, 256, size=(64, 4096, 32), dtype=np.uint8)
blobs[:, :, 0] = 0
blobs = blobs.reshape(64, -1)
ts.compute_cell_proofs_batch(blobs[:1])
ts.compute_cell_proofs_batch(blobs)
t = time.perf_counter()
for _ in range(3):
    ts.compute_cell_proofs_batch(blobs)
print(json.dumps({"fk20_c": os.environ.get("B200_FK20_C", "default"), "proofs_ms_per_64_blobs": (time.perf_counter() - t) / 3 * 1e3}))
