cd $GRAFT_REPO_ROOT
(time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) 2>&1 | tail -7
python scripts/fk20_timing.py 2>&1 | tail -2
B200_FFT_G1_FUSE=1 python scripts/fk20_timing.py 2>&1 | tail -2
python scripts/verify_timing.py 2>&1 | tail -12
python - <<'PY'
import os, sys, numpy as np, torch, json
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
import rust_kzg_b200 as B, bench
K, osettings, L = bench.load_bases()
fs = B.FFTSettings(15)
pts = np.ascontiguousarray(np.tile(osettings.g1_monomial, (8, 1)))
d_pts = torch.from_numpy(pts.view(np.int64)).cuda(); d_res = torch.zeros_like(d_pts)
for fuse in ("0", "1"):
    os.environ["B200_FFT_G1_FUSE"] = fuse
    ms = bench.timed_events(torch, lambda: fs.fft_g1_device(d_res.data_ptr(), d_pts.data_ptr(), 1 << 15, False, 1, 0), reps=3, warm=1)
    print("fft_g1 2^15 fuse", fuse, ms)
PY
