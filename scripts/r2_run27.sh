cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_ntt.py -m gpu -x -q -k "g1" 2>&1 | tail -3
python scripts/fft_g1_batch_timing.py 2>&1 | tail -7
timeout 900 python -m pytest tests/test_gpu_eip4844.py tests/test_gpu_das7594.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import rust_kzg_b200 as B
rng = np.random.default_rng(1)
ts = B.KZGSettings.load_trusted_setup_file()
blobs = rng.integers(0, 256, size=(64, 4096, 32), dtype=np.uint8); blobs[:, :, 0] = 0
blobs = blobs.reshape(64, -1)
cb = np.zeros((64, 128, 2048), np.uint8); pb = np.zeros((64, 128, 48), np.uint8)
for n in (1, 4, 8, 16, 32, 64):
    ts.compute_cells_and_kzg_proofs_batch(blobs[:n], cb[:n], pb[:n])
    t = time.perf_counter()
    for _ in range(3):
        ts.compute_cells_and_kzg_proofs_batch(blobs[:n], cb[:n], pb[:n])
    both = (time.perf_counter() - t) / 3 * 1e3
    print(n, 'cells+proofs %.2f ms -> %.0f blobs/s' % (both, n / both * 1e3))
PY
