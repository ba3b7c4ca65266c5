"""A/B timing of MSM 2^20 accumulate variants selected by environment variables (not a bench line)."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rust_kzg_b200 as B
    from oracle import c_oracle as K
    logn = int(os.environ.get("LOGN", "20"))
    n = 1 << logn
    text = open(os.path.join(os.path.dirname(B.LIB_PATH), "data", "trusted_setup.txt")).read()
    L = K.p1s_to_affine(K.KZGSettings(text).g1_lagrange_brp)
    pts = np.tile(L, (max(1, n // 4096), 1))[:n]
    rng = np.random.default_rng(1)
    sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    h = B.PreparedMsm(pts)
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    for _ in range(3):
        h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
    torch.cuda.synchronize()
    h.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
    e1.record()
    torch.cuda.synchronize()
    acc, runs = h.profile_read()
    folded = sc[:4096].copy()
    for k in range(1, n // 4096):
        folded = K.fr_add(folded, sc[k * 4096:(k + 1) * 4096])
    ok = K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == K.p1_compress(K.msm_affine(L, folded, nthreads=8))
    print(json.dumps({"ms": e0.elapsed_time(e1) / 10, "acc_ms": acc / runs, "ok": bool(ok), **h.info()}))
else:
    for env in sys.argv[1:]:
        e = dict(os.environ)
        for kv in env.split(","):
            if "=" in kv:
                k, v = kv.split("=")
                e[k] = v
        out = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(env, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:], flush=True)
