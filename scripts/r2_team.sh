cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_eip4844.py tests/test_gpu_das7594.py -m gpu -x -q 2>&1 | tail -2
for team in 0 16; do B200_FK20_TEAM=$team python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0, '.')
import rust_kzg_b200 as B
rng = np.random.default_rng(1)
ts = B.KZGSettings.load_trusted_setup_file()
blobs = rng.integers(0, 256, size=(16, 4096, 32), dtype=np.uint8); blobs[:, :, 0] = 0
blobs = blobs.reshape(16, -1)
ref = ts.compute_cell_proofs_batch(blobs)
out = []
for n in (1, 2, 4, 8, 16):
    ts.compute_cell_proofs_batch(blobs[:n])
    t = time.perf_counter()
    for _ in range(5):
        p = ts.compute_cell_proofs_batch(blobs[:n])
    out.append("%d: %.2f ms" % (n, (time.perf_counter() - t) / 5 * 1e3))
    assert np.array_equal(p, ref[:n])
print("team", os.environ["B200_FK20_TEAM"], " ".join(out), flush=True)
PY
done
