"""Quick GPU probe: integer-pipe microbenchmark + MSM timings through the device-pointer ABI (not a bench line)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B
from oracle import c_oracle as K

out = {}
out["microbench"] = B.microbench_int()
print(out["microbench"], flush=True)

text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
s = K.KZGSettings(text, nthreads=os.cpu_count())
L = K.p1s_to_affine(s.g1_lagrange_brp)
rng = np.random.default_rng(1)


def rand_fr(n):
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    return a


for logn in [int(x) for x in (sys.argv[1:] or ["12", "16", "20"])]:
    n = 1 << logn
    pts = np.tile(L, (max(1, n // 4096), 1))[:n]
    t0 = time.time()
    h = B.PreparedMsm(pts)
    torch.cuda.synchronize()
    t_prep = time.time() - t0
    sc = rand_fr(n)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    got = d_out.cpu().numpy().view(np.uint64)
    # folded-scalar oracle
    folded = sc[:4096].copy() if n >= 4096 else sc
    for k in range(1, n // 4096):
        folded = K.fr_add(folded, sc[k * 4096:(k + 1) * 4096])
    exp = K.msm_affine(L[:min(n, 4096)], folded, nthreads=os.cpu_count())
    ok = K.p1_compress(got) == K.p1_compress(exp)
    t0 = time.time()
    host = h.mult(sc)
    t_host = time.time() - t0
    info = h.info()
    rec = dict(logn=logn, ms=ms, mpts_per_s=n / ms / 1e3, ok=bool(ok), prep_s=t_prep, host_call_ms=t_host * 1e3, **info)
    print(rec, flush=True)
    out["msm_%d" % logn] = rec
    h.close()

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
