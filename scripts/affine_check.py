#!/usr/bin/env python3
"""Batch-affine bucket accumulation (k_accumulate_affine) vs the XYZZ task kernel and the oracle: parity on uniform and
adversarial inputs that exercise every exceptional case (P + P, P + (-P), infinity operands), then device-side timing of
the 2^20-term MSM for both paths and the slot counts K.  Writes one JSON document to stdout.
    python scripts/affine_check.py [logn=20]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import rust_kzg_b200 as B  # noqa: E402
from oracle import c_oracle as K  # noqa: E402
import bench  # noqa: E402

R = bench.R_MOD


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    Kx, s, L = bench.load_bases()
    cores = os.cpu_count() or 1
    out = {"parity": {}, "timing": {}}
    # ---- parity at 2^17 (smallest size that takes the affine path), several window widths ----------------------------
    n = 1 << 17
    reps = n // 4096
    rng = np.random.default_rng(11)
    pts = np.tile(L, (reps, 1))
    pts_inf = pts.copy()
    pts_inf[5::97] = 0                                   # points at infinity in the table
    base4096 = bench.rand_fr(rng, 4096)
    neg4096 = K.fr_sub(np.zeros((4096, 4), np.uint64), base4096)
    cases = {
        "uniform": bench.rand_fr(rng, n),
        "tiled_same_scalar": np.tile(base4096, (reps, 1)),                       # 32 copies of every (point, digit): P + P everywhere
        "plus_minus": np.concatenate([np.tile(base4096, (reps // 2, 1)), np.tile(neg4096, (reps // 2, 1))]),   # P + (-P): sum is infinity
        "all_equal": bench.adversarial_scalars(K, rng, n, "all_equal"),
        "r_minus_1": bench.adversarial_scalars(K, rng, n, "r_minus_1"),
        "below_2^64": bench.adversarial_scalars(K, rng, n, "below_2^64"),
        "ten_percent_zero": bench.adversarial_scalars(K, rng, n, "ten_percent_zero"),
        "zeros": np.zeros((n, 4), np.uint64),
    }
    for cwin in (0, 13, 20):
        env = {"B200_MSM_AFFINE": 1}
        if cwin:
            env["B200_MSM_C"] = cwin
        for tag, p in (("", pts), ("+inf_points", pts_inf)):
            h = with_env(env, lambda: B.PreparedMsm(p))
            for name, sc in cases.items():
                if tag and name not in ("uniform", "tiled_same_scalar", "plus_minus"):
                    continue
                sc = np.ascontiguousarray(sc)
                got = h.mult(sc)
                aff = h.info()["accumulate"]
                if tag:
                    keep = np.ones(n, bool)
                    keep[5::97] = False
                    sc_eff = sc.copy()
                    sc_eff[~keep] = 0
                else:
                    sc_eff = sc
                want = bench.folded_expectation(K, L, sc_eff, cores)
                out["parity"]["c=%s %s%s" % (cwin or "auto", name, tag)] = {"ok": bool(K.p1_compress(got) == K.p1_compress(want)), "path": aff}
            h.close()
    # ---- timing at 2^logn: XYZZ vs affine with K = 8, 10, 12, 16 -------------------------------------------------------
    n = 1 << logn
    rng = np.random.default_rng(bench.SEED)
    sc = bench.rand_fr(rng, n)
    pts = np.tile(L, (n // 4096, 1))
    exp = K.p1_compress(bench.folded_expectation(K, L, sc, cores))
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    confs = [("xyzz", {"B200_MSM_AFFINE": 0})] + [("affine_k%d" % k, {"B200_MSM_AFFINE": 1, "B200_MSM_AFFINE_K": k}) for k in (8, 10, 12, 16)]
    for name, env in confs:
        try:
            h = with_env(env, lambda: B.PreparedMsm(pts))
            ms = bench.timed_events(torch, lambda: h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0), reps=10, warm=3)
            ok = K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == exp
            h.set_profiling(True)
            for _ in range(5):
                h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
            torch.cuda.synchronize()
            acc_ms, runs = h.profile_read()
            st = h.last_stats()
            out["timing"][name] = {"ms": ms, "accumulate_ms": acc_ms / max(runs, 1), "parity_ok": bool(ok), "path": h.info()["accumulate"],
                                   "entries": st["entries"], "partials": st["tasks"], "adds": st["adds"]}
            h.close()
        except Exception as e:
            out["timing"][name] = {"error": repr(e)[:300]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
