cd $GRAFT_REPO_ROOT
timeout 1400 python -m pytest tests/test_gpu_msm.py tests/test_gpu_threads.py tests/test_c_consumer.py -m gpu -x -q 2>&1 | tail -4
python - <<'PY'
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import rust_kzg_b200 as B
from oracle import c_oracle as K
text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
L = K.p1s_to_affine(K.KZGSettings(text).g1_lagrange_brp)
rng = np.random.default_rng(1)
sc = rng.integers(0, 1 << 62, size=(64 * 4096, 4), dtype=np.uint64)
for direct, bits in (("0", "11"), ("1", "8"), ("1", "11"), ("1", "13")):
    os.environ["B200_MSM_DIRECT"] = direct; os.environ["B200_MSM_DIRECT_BITS"] = bits
    t0 = time.perf_counter(); h = B.PreparedMsm(L); prep = time.perf_counter() - t0
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda(); d_out = torch.zeros((64, 18), dtype=torch.int64, device="cuda")
    row = {}
    for batch in (1, 8, 64):
        for _ in range(3): h.mult_device(d_out.data_ptr(), 4096, d_sc.data_ptr(), batch, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): h.mult_device(d_out.data_ptr(), 4096, d_sc.data_ptr(), batch, 0)
        e1.record(); torch.cuda.synchronize()
        row["dev_batch%d_ms" % batch] = round(e0.elapsed_time(e1) / 10, 4)
    one = sc[:4096].copy()
    for _ in range(3): h.mult(one)
    t0 = time.perf_counter()
    for _ in range(20): h.mult(one)
    row["host_call_ms"] = round((time.perf_counter() - t0) / 20 * 1e3, 4)
    print("direct", direct, "bits", h.info()["direct_bits"], "prepare_s %.2f" % prep, row, flush=True)
    h.close()
PY
