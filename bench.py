#!/usr/bin/env python3
"""bench.py -- headline benchmark of the rust-kzg hot path on B200 (contract in the task statement).

A "step" is one pass of the hot path over one batch of synthetic input.  Workload at every N (weak scaling):
BASELINE.json configs[1], an MSM of 2^LOG_N (default 2^20) uniformly random Fr scalars x trusted-setup G1 points per
GPU, bit-exact vs the CPU oracle.  For N > 1 the 2^LOG_N * N terms are sharded by rank through the
library's own multi-GPU entry point (b200_msm_sharded_mult_device): local MSM, NCCL all-gather of the 144-byte partial
results, local add -- NCCL cannot add curve points.

Units.  `value` / `e2e.value` are BASELINE's "G1-adds/s" = 16 canonical bucket additions per term (canonical c = 16,
BASELINE.md section 2) for BOTH arms, i.e. a points/s ratio; `points_per_s` is the plain term rate and
`adds_per_s_actual` counts the additions this engine really performs (N*W/t + reduce adds with the
implementation's own W), read back from the device.  `value` has scalars resident in HBM; `e2e` goes through the
reference-facing C ABI call mult_pippenger_prepared with pinned HOST scalars (H2D + D2H inside the timed region);
`e2e_pageable` is the same call on pageable memory (what a Rust Vec<Fr> is).

`roofline` is the BINDING roofline of the dominant kernel -- the 32-bit integer multiply pipe --,
`hbm_roofline` the HBM one the north star also asks for.  Extra keys report the other BASELINE metrics: blobs/s
(batch 64, and through the unmodified single-blob c-kzg symbols from 1 / 4 / 16 host threads), the Fr NTT / DAS sweep
2^12..2^20, fft_g1, timed adversarial inputs, and BASELINE configs[4] (2^24 terms sharded over the N GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = int(os.environ.get("B200_BENCH_LOGN", "20"))
LOG_BIG = int(os.environ.get("B200_BENCH_LOGBIG", "24"))     # BASELINE configs[4]: total terms sharded over the N GPUs
ADDS_PER_TERM = 16          # canonical c = 16
BYTES_PER_TERM = 128        # 32 B scalar + 96 B affine point
SEED = 0x4B5A47
METRIC = "MSM G1-adds/sec at 2^%d (per GPU, weak scaling); blobs/sec blob_to_kzg_commitment in extra" % LOG_N
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def workload_config():
    """the part of `config` both arms share word for word (the driver compares them)"""
    return {"workload": "MSM 2^%d random Fr scalars x EIP-4844 trusted-setup G1 points (4096 Lagrange points tiled), per GPU" % LOG_N,
            "terms_per_gpu": 1 << LOG_N, "seed": SEED, "prng": "numpy PCG64",
            "adds_unit": "16 canonical bucket additions per term (BASELINE.md section 2) for every arm, whatever window "
                         "width an engine uses; points_per_s and adds_per_s_actual are reported beside it"}


def rand_fr(rng, n):
    """uniform-looking Montgomery Fr limbs below 2^254 < r (numpy PCG64, seed SEED)"""
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    return a


def rand_blobs(rng, n):
    b = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    b[:, :, 0] = 0   # canonical field elements (kzg-bench/src/tests/eip_4844.rs:28-37)
    return b.reshape(n, 131072)


def int_to_limbs(v):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class ClockSampler:
    """SM clock + throttle reasons sampled every 20 ms during the timed region (B200_PROFILING.md recipe; NVML is what
    nvidia-smi reads -- a piped `nvidia-smi -lms` block-buffers its output, so the library is polled directly)"""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.sm, self.bits, self.mx, self.index = [], 0, None, index
        self.stop_flag, self.thread = threading.Event(), None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return

        def loop():
            while not self.stop_flag.is_set():
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    pass
                time.sleep(0.02)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def mark(self):
        """forget samples taken so far (called right before the timed region)"""
        self.sm, self.bits = [], 0

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=1.0)
        reasons = sorted(n for b, n in self.REASONS.items() if self.bits & b)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": reasons, "samples": len(self.sm)}


def load_bases():
    """affine Lagrange bases of the EIP-4844 trusted setup, bit-reversed, via the oracle's parser (host, one-time)"""
    from oracle import c_oracle as K
    text = open(os.path.join(ROOT, "rust-kzg_b200", "data", "trusted_setup.txt")).read()
    s = K.KZGSettings(text, nthreads=os.cpu_count() or 1)
    return K, s, K.p1s_to_affine(s.g1_lagrange_brp)


def fold_scalars(K, sc):
    """sum_k sc[j + 4096 k] mod r for j < 4096, by halving (log2 passes of one vectorised fr_add each)"""
    assert sc.shape[0] % 4096 == 0
    cur = sc
    while cur.shape[0] > 4096:
        blocks = cur.shape[0] // 4096
        if blocks & 1:
            head, cur = cur[:4096], cur[4096:]
            cur = np.concatenate([K.fr_add(cur[:4096], head), cur[4096:]])
            blocks -= 1
        half = (blocks // 2) * 4096
        cur = K.fr_add(cur[:half], cur[half:])
    return np.ascontiguousarray(cur)


def folded_expectation(K, L, sc, nthreads):
    """exact oracle for tiled bases P_i = L[i mod 4096]: fold the scalars, then a 4096-term CPU MSM"""
    return K.msm_affine(L, fold_scalars(K, sc), nthreads=nthreads)


def cpu_msm_pass(K, L, n_terms, nthreads, steps=1, warmup=0):
    """the oracle's restatement of tiling_parallel_pippenger on the host cores over n_terms terms of the workload"""
    rng = np.random.default_rng(SEED)
    sc = rand_fr(rng, n_terms)
    pts = np.tile(L, (max(1, n_terms // 4096), 1))[:n_terms]
    for _ in range(warmup):
        K.msm_affine(pts, sc, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(steps):
        K.msm_affine(pts, sc, nthreads=nthreads)
    dt = (time.perf_counter() - t0) / steps
    return n_terms * ADDS_PER_TERM / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust reference cannot be built here) on all
    host threads over the SAME workload as the b200 arm -- the whole 2^LOG_N-term MSM per step.  Only when
    (steps + warmup) full passes would exceed four minutes is the step cut to a prefix, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, s, L = load_bases()
    cores = os.cpu_count() or 1
    n = 1 << LOG_N
    _, probe = cpu_msm_pass(K, L, 1 << 16, cores)               # ~60 ms on 16 threads: sizes the run
    est = probe * (n / (1 << 16)) * (args.steps + args.warmup)
    full = True
    while n > (1 << 14) and est > 240.0:
        n >>= 1
        est /= 2
        full = False
    value, dt = cpu_msm_pass(K, L, n, cores, steps=args.steps, warmup=args.warmup)
    cfg = workload_config()
    cfg["sample"] = "whole workload per step" if full else "first 2^%d terms per step (time bound)" % (n.bit_length() - 1)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "G1-adds/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": cfg, "points_per_s": value / ADDS_PER_TERM,
        "cpu_baseline": {"value": value, "unit": "G1-adds/s", "cores": cores, "kind": "port",
                         "sample": "%s of the 2^%d-term MSM, %.2f s per pass on %d threads; tiling_parallel_pippenger restatement "
                                   "(oracle/kzg_oracle.c), not blst assembly" % ("all" if full else "first 2^%d terms" % (n.bit_length() - 1),
                                                                                 LOG_N, dt, cores)},
        "e2e": {"value": value, "unit": "G1-adds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def dist_max(torch, dist, world, x):
    if world == 1:
        return x
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extra", action="store_true", help="skip the blob / NTT / adversarial / 2^24 extra metrics")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import rust_kzg_b200 as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or B.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << LOG_N
    K, osettings, L = load_bases()
    cores = os.cpu_count() or 1

    def bcast(b):
        t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    # ---- inputs: rank r owns terms [r*n, (r+1)*n) of the N*n-term MSM; bases tile the 4096 setup points ----------
    rng = np.random.default_rng(SEED + rank)
    sc = rand_fr(rng, n)
    pts = np.tile(L, (n // 4096, 1)) if n >= 4096 else L[:n]
    smsm = B.ShardedMsm(pts, rank, world, broadcast=bcast)           # the library's multi-GPU entry point (world 1: no NCCL)
    msm = B.PreparedMsm(None, _borrowed_handle=smsm.local_handle())  # view of the rank-local handle: info / profiling
    h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory()
    d_sc = h_sc.cuda(non_blocking=True)
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    stream = 0  # the library launches on the legacy default stream = torch's current stream, so torch events see it

    def step():
        smsm.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples from here are dropped at mark(); kept ones span the timed region + e2e leg
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    info = msm.info()
    # parity of the thing being timed: the sum over ALL ranks' terms vs the folded-scalar oracle
    if world > 1:
        folded = torch.from_numpy(fold_scalars(K, sc).view(np.int64)).cuda()
        allf = [torch.zeros_like(folded) for _ in range(world)]
        dist.all_gather(allf, folded)
        tot = allf[0].cpu().numpy().view(np.uint64).reshape(4096, 4)
        for t in allf[1:]:
            tot = K.fr_add(tot, t.cpu().numpy().view(np.uint64).reshape(4096, 4))
        exp = K.msm_affine(L, np.ascontiguousarray(tot), nthreads=cores)
    else:
        exp = folded_expectation(K, L, sc, cores)
    if K.p1_compress(d_out.cpu().numpy().view(np.uint64)) != K.p1_compress(exp):
        raise SystemExit("bench.py: MSM result differs from the oracle -- refusing to report a number")

    msm.set_profiling(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    acc_ms_sum, acc_runs = msm.profile_read()
    msm.set_profiling(False)
    stats = msm.last_stats()   # work counters of the last step, read back from the device
    ms_step = dist_max(torch, dist, world, ms_total) / args.steps
    value = world * n * ADDS_PER_TERM / (ms_step * 1e-3)

    # ---- e2e: the reference-facing call with host scalars, H2D + D2H inside the timed region ----------------------
    h_np = h_sc.numpy().view(np.uint64).reshape(n, 4)

    def e2e_leg(arr, reps):
        for _ in range(2):
            smsm.mult(arr)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = smsm.mult(arr)
        dt = dist_max(torch, dist, world, (time.perf_counter() - t0) / reps)
        if K.p1_compress(r) != K.p1_compress(exp):
            raise SystemExit("bench.py: e2e MSM result differs from the oracle")
        return dt

    e2e_steps = max(3, min(args.steps, 10))
    t_e2e = e2e_leg(h_np, e2e_steps)                 # pinned host scalars
    t_e2e_pg = e2e_leg(sc, e2e_steps)                # pageable host scalars (plain numpy memory)
    e2e_value = world * n * ADDS_PER_TERM / t_e2e
    clocks = sampler.stop() if rank == 0 else None

    # ---- extras every rank takes part in: BASELINE configs[4] and blob-batch replicas -----------------------------
    shared_extra = {}
    if not args.no_extra:
        del d_sc, h_sc
        smsm.close()
        msm.close()
        torch.cuda.empty_cache()
        try:
            shared_extra["msm_2p%d" % LOG_BIG] = big_msm(B, K, L, torch, dist, rank, world, cores, bcast)
        except Exception as e:
            shared_extra["msm_2p%d" % LOG_BIG] = {"error": repr(e)[:300]}
        if world > 1:
            try:
                shared_extra["blob_replicas"] = blob_replicas(B, K, osettings, torch, dist, rank, world)
            except Exception as e:
                shared_extra["blob_replicas"] = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines of the dominant kernel (the bucket accumulation), timed live with CUDA events on its stream ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    acc_ms = acc_ms_sum / max(acc_runs, 1)
    traffic = None
    try:   # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by configuration; null when not captured
        tr = json.load(open(os.path.join(ROOT, "profiles", "accumulate_traffic.json")))
        traffic = tr.get("by_config", {}).get("2^%d,c=%d,%s" % (LOG_N, info["c"], info.get("accumulate", "xyzz")))
    except Exception:
        pass
    mb = B.microbench_int()
    mul_per_add = info.get("fp_mul_per_add", 10.0)   # 8M + 2S (XYZZ mixed addition) or the batch-affine count
    imad_per_add = mul_per_add * 300                  # Fp mul = 144 + 144 + 12 32x32 multiply-adds
    acc_adds = stats["adds"]["accumulate"]
    achieved_int = acc_adds * imad_per_add / (acc_ms * 1e-3) / 1e12 if acc_ms else None
    roofline = {"bound": "int32-mul pipe (IMAD.WIDE.U32, 32x32+64 multiply-add; FMA-heavy pipe)", "unit": "T multiply-adds/s",
                "achieved": achieved_int, "peak": mb["imad_per_s"] / 1e12,
                "frac": achieved_int / (mb["imad_per_s"] / 1e12) if achieved_int else None,
                "traffic": traffic, "kernel": info.get("accumulate_kernel", "k_accumulate"), "kernel_ms": acc_ms,
                "kernel_share_of_step": acc_ms / ms_step if acc_ms else None,
                "peak_source": "b200_microbench_int (dependent-operand IMAD.WIDE loop) measured in this run; theoretical "
                               "32/clk/SM x 148 SMs x SM clock",
                "fp_mul_per_s_measured": mb["fpmul_per_s"], "adds_per_launch": acc_adds, "fp_mul_per_add": mul_per_add,
                "note": "achieved = additions the kernel performs (read back from the device) x %.1f Fp multiplications x 300 "
                        "32-bit multiply-adds / kernel time; this path is bound by the integer multiply pipe, not HBM "
                        "(SURVEY.md 8d)" % mul_per_add}
    achieved_gbs = BYTES_PER_TERM * n / (acc_ms * 1e-3) / 1e9 if acc_ms else None
    hbm_roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved_gbs / hbm_peak if achieved_gbs else None, "traffic": traffic,
                    "traffic_gbs": traffic / (acc_ms * 1e-3) / 1e9 if traffic and acc_ms else None,
                    "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback",
                    "note": "algorithmic bytes = 128 B per term (32 B scalar + 96 B point); the fraction is small by the nature "
                            "of the problem (94k multiply-adds per 128 bytes)"}

    # ---- CPU baseline beside it: oracle port on the host cores over the whole workload ------------------------------
    cpu_value, cpu_dt = cpu_msm_pass(K, L, n, cores)
    cpu = {"value": cpu_value, "unit": "G1-adds/s", "cores": cores, "kind": "port",
           "sample": "the whole 2^%d-term MSM once, %.2f s on %d threads; C restatement of tiling_parallel_pippenger "
                     "(oracle/kzg_oracle.c), not blst assembly" % (LOG_N, cpu_dt, cores)}

    cfg = workload_config()
    cfg.update({"window_bits": info["c"], "windows": info["W"], "table_bytes": info["table_bytes"],
                "l2": "inputs larger than L2: %d MiB scalars + %.1f GiB table per step" % (n * 32 >> 20, info["table_bytes"] / 2 ** 30),
                "parity": "compressed result == folded-scalar oracle (checked before timing, and on both e2e legs)",
                "multi_gpu": "terms sharded by rank inside the library (b200_msm_sharded_mult_device): NCCL all-gather of "
                             "144 B partial results + local add" if world > 1 else "single GPU"})
    out = {
        "metric": METRIC, "value": value, "unit": "G1-adds/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": cfg,
        "points_per_s": world * n / (ms_step * 1e-3),
        "adds_per_s_actual": world * stats["adds_total"] / (ms_step * 1e-3), "adds_actual_per_step": stats,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "G1-adds/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 144,
                "ms_per_step": t_e2e * 1e3, "points_per_s": world * n / t_e2e,
                "api": "mult_pippenger_prepared / b200_msm_sharded_mult (C ABI), pinned host scalars"},
        "e2e_pageable": {"value": world * n * ADDS_PER_TERM / t_e2e_pg, "unit": "G1-adds/s", "ms_per_step": t_e2e_pg * 1e3,
                         "api": "same call, pageable host scalars (plain heap memory)"},
        "gpu_launches": (info["launches"] + (1 if world > 1 else 0)) * args.steps,
        "roofline": roofline, "hbm_roofline": hbm_roofline, "cpu_baseline": cpu,
    }
    extra = dict(shared_extra)
    if not args.no_extra and world == 1:
        for name, fn in (("g1_lincomb_4096", lambda: lincomb_4096(B, K, L, torch)),
                         ("variable_base_e2e", lambda: variable_base(B, K, torch, pts, sc, exp)),
                         ("adversarial_msm", lambda: adversarial_msm(B, K, L, torch, pts, ms_step)),
                         ("blobs", lambda: blob_metrics(B, K, osettings, torch)),
                         ("threads", lambda: thread_metrics()),
                         ("ntt", lambda: ntt_metrics(B, K, osettings, torch))):
            try:
                res = fn()
                if name in ("blobs", "ntt"):
                    extra.update(res)
                else:
                    extra[name] = res
            except Exception as e:  # extras must not take the headline down
                extra[name] = {"error": repr(e)[:300]}
    if extra:
        out["extra"] = extra
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(out))


def lincomb_4096(B, K, L, torch):
    """mult_pippenger_prepared at the size the reference calls g1_lincomb with (4096 Lagrange points, kzg/src/eip_4844.rs:463-476):
    one call host to host and device-resident, on the handle's direct-lookup table and on the bucket pipeline beside it"""
    rng = np.random.default_rng(SEED + 9)
    n = L.shape[0]
    sc = np.ascontiguousarray(rand_fr(rng, n))
    want = K.p1_compress(K.msm_affine(L, sc, nthreads=os.cpu_count() or 1))
    out = {"npoints": int(n), "api": "prepare_msm / mult_pippenger_prepared (C ABI), pageable host scalars"}
    for label, direct in (("direct_table", "1"), ("bucket_pipeline", "0")):
        os.environ["B200_MSM_DIRECT"] = direct
        try:
            h = B.PreparedMsm(L)
        finally:
            del os.environ["B200_MSM_DIRECT"]
        got = h.mult(sc)
        for _ in range(5):
            h.mult(sc)
        t0 = time.perf_counter()
        for _ in range(50):
            h.mult(sc)
        host_ms = (time.perf_counter() - t0) / 50 * 1e3
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
        dev_ms = timed_events(torch, lambda: h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0), reps=20, warm=5)
        out[label] = {"host_call_ms": host_ms, "device_ms": dev_ms, "direct_bits": h.info()["direct_bits"],
                      "parity_ok": bool(K.p1_compress(got) == want and K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == want)}
        h.close()
    return out


def timed_events(torch, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def big_msm(B, K, L, torch, dist, rank, world, cores, bcast):
    """BASELINE configs[4]: ONE MSM of 2^LOG_BIG terms sharded over the N GPUs (strong scaling: the total is fixed, each
    rank holds 2^LOG_BIG / N terms and their table rows).  Every rank gets the full result; checked against the
    folded-scalar oracle of the whole problem."""
    total = 1 << LOG_BIG
    lo, hi = B.shard_bounds(total, rank, world)
    nl = hi - lo
    assert lo % 4096 == 0 and nl % 4096 == 0
    rng = np.random.default_rng(SEED + 1000 + rank)
    sc = rand_fr(rng, nl)
    t0 = time.perf_counter()
    sm = B.ShardedMsm(np.tile(L, (nl // 4096, 1)), rank, world, broadcast=bcast)
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    for _ in range(2):
        sm.mult_device(d_out.data_ptr(), nl, d_sc.data_ptr(), 0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sm.mult_device(d_out.data_ptr(), nl, d_sc.data_ptr(), 0)
    e1.record()
    torch.cuda.synchronize()
    ms = dist_max(torch, dist, world, e0.elapsed_time(e1) / reps)
    folded = fold_scalars(K, sc)
    if world > 1:
        f = torch.from_numpy(folded.view(np.int64)).cuda()
        allf = [torch.zeros_like(f) for _ in range(world)]
        dist.all_gather(allf, f)
        folded = allf[0].cpu().numpy().view(np.uint64).reshape(4096, 4)
        for t in allf[1:]:
            folded = K.fr_add(folded, t.cpu().numpy().view(np.uint64).reshape(4096, 4))
    ok = K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == K.p1_compress(K.msm_affine(L, np.ascontiguousarray(folded), nthreads=cores))
    info = B.PreparedMsm(None, _borrowed_handle=sm.local_handle()).info()
    sm.close()
    del d_sc
    torch.cuda.empty_cache()
    return {"total_terms": total, "terms_per_gpu": nl, "n_gpus": world, "ms": ms, "points_per_s": total / (ms * 1e-3),
            "g1_adds_per_s": total * ADDS_PER_TERM / (ms * 1e-3), "scaling": "strong", "parity_ok": bool(ok),
            "window_bits": info["c"], "table_bytes_per_gpu": info["table_bytes"], "prepare_s": t_prep,
            "timing": "CUDA events, device-resident scalars, max over ranks"}


def blob_replicas(B, K, osettings, torch, dist, rank, world):
    """blob batches do not shard: N independent replicas, one 64-blob commitment batch per GPU per step"""
    rng = np.random.default_rng(SEED + 77 + rank)
    nb = 64
    blobs = rand_blobs(rng, nb)
    ts = B.KZGSettings.load_trusted_setup_file()
    d_blobs = torch.from_numpy(blobs).cuda()
    d_out = torch.zeros((nb, 48), dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(nb, dtype=torch.int32, device="cuda")
    fn = lambda: ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), nb, d_status.data_ptr(), 0)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    ms = dist_max(torch, dist, world, timed_events(torch, fn, reps=10, warm=0))
    ok = d_out[5].cpu().numpy().tobytes() == K.blob_to_kzg_commitment(blobs[5].tobytes(), osettings)
    okt = torch.tensor([1.0 if ok else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    ts.free()
    return {"blob_to_kzg_commitment_blobs_per_s": world * nb / (ms * 1e-3), "ms_per_batch": ms, "batch_per_gpu": nb,
            "n_gpus": world, "parity_ok": bool(okt.item() == 1.0), "mode": "replicas only, no collective"}


def variable_base(B, K, torch, pts, sc, exp):
    """variable-base call of the reference's L4 benchmark shape (bench_g1_lincomb: points + scalars travel with every
    call, kzg-bench/src/benches/lincomb.rs:35-46; 116.75 ms on L4 + sppark)"""
    n = sc.shape[0]
    h_pts = torch.from_numpy(pts.view(np.int64)).pin_memory()
    h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory()
    hp, hs = h_pts.numpy().view(np.uint64).reshape(n, 12), h_sc.numpy().view(np.uint64).reshape(n, 4)
    B.mult_pippenger(hp, hs)
    t0 = time.perf_counter()
    for _ in range(3):
        rv = B.mult_pippenger(hp, hs)
    t_var = (time.perf_counter() - t0) / 3
    return {"api": "mult_pippenger (C ABI), pinned host points + scalars", "ms_per_call": t_var * 1e3, "points_per_s": n / t_var,
            "h2d_bytes_per_call": 128 * n, "parity_ok": bool(K.p1_compress(rv) == K.p1_compress(exp))}


def adversarial_scalars(K, rng, n, kind):
    """Montgomery-form scalar vectors of the reference's edge suites (kzg-bench/src/tests/bls12_381.rs:282-296) and of the
    consensus vectors (all-equal elements).  The library receives blst_fr = value * 2^256 mod r; the VALUE is what counts
    for the digit distribution, so it is chosen first and converted (Montgomery product with R^2)."""
    def to_mont(canon):   # (n, 4) canonical limbs -> Montgomery limbs, through the oracle's batched Montgomery product
        rr = np.tile(int_to_limbs((1 << 512) % R_MOD), (canon.shape[0], 1))
        return K.fr_mul(np.ascontiguousarray(canon), rr)
    if kind == "all_equal":                       # every scalar the same value (the all-0x02 blob of the golden vectors)
        return np.tile(int_to_limbs(2 * (1 << 256) % R_MOD), (n, 1))
    if kind == "r_minus_1":
        return np.tile(int_to_limbs((R_MOD - 1) * (1 << 256) % R_MOD), (n, 1))
    if kind == "below_2^64":                      # small scalars: 64 random bits each, all distinct
        canon = np.zeros((n, 4), np.uint64)
        canon[:, 0] = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
        return to_mont(canon)
    if kind == "ten_percent_zero":
        a = rand_fr(rng, n)
        a[rng.random(n) < 0.1] = 0
        return a
    raise ValueError(kind)


def adversarial_msm(B, K, L, torch, pts, uniform_ms):
    """the headline MSM on adversarial scalar distributions, timed (device-resident) and checked against the oracle"""
    n = pts.shape[0]
    msm = B.PreparedMsm(pts)
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    out = {"uniform_ms": uniform_ms}
    rng = np.random.default_rng(SEED + 5)
    worst = 0.0
    for kind in ("all_equal", "r_minus_1", "below_2^64", "ten_percent_zero"):
        sc = np.ascontiguousarray(adversarial_scalars(K, rng, n, kind))
        d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
        ms = timed_events(torch, lambda: msm.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0), reps=5, warm=3)
        ok = K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == K.p1_compress(folded_expectation(K, L, sc, os.cpu_count() or 1))
        out[kind] = {"ms": ms, "vs_uniform": ms / uniform_ms, "parity_ok": bool(ok)}
        worst = max(worst, ms / uniform_ms)
        del d_sc
    out["worst_vs_uniform"] = worst
    msm.close()
    return out


def thread_metrics():
    """blobs/s through the UNMODIFIED single-blob c-kzg symbols from a C pthread driver (examples/ckzg_threads.c): T host
    threads, one pageable blob per call -- the way the reference's consumers call them (kzg/src/eip_4844.rs:770-816)"""
    exe = "/tmp/b200_ckzg_threads_bench"
    lib_dir = os.path.join(ROOT, "rust-kzg_b200")
    subprocess.check_call(["gcc", "-O2", "-pthread", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "ckzg_threads.c"),
                           "-L" + lib_dir, "-lb200kzg", "-Wl,-rpath," + lib_dir, "-o", exe])
    setup = os.path.join(lib_dir, "data", "trusted_setup.txt")
    out = {"driver": "examples/ckzg_threads.c (pthreads, pageable malloc'd blobs, every result compared with the single-threaded one)"}
    for op, key in (("commit", "blob_to_kzg_commitment"), ("blob_proof", "compute_blob_kzg_proof")):
        rec = {}
        for t in (1, 4, 16):
            calls = 200 if t == 1 else 400 if t == 4 else 300
            r = subprocess.run([exe, setup, op, str(t), str(calls), "4"], capture_output=True, text=True, timeout=300)
            line = json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else {"error": r.stderr[-200:]}
            rec["e2e_threads%d_blobs_per_s" % t] = line.get("per_s")
            rec["threads%d_bit_exact" % t] = bool(r.returncode == 0 and line.get("mismatches") == 0 and line.get("errors") == 0)
        out[key] = rec
    # compute_cells_and_kzg_proofs per blob from 1 / 9 threads (a block's blobs under a parallel iterator): shared FK20 passes
    rec = {}
    for t in (1, 9):
        r = subprocess.run([exe, setup, "cells", str(t), "12", "2"], capture_output=True, text=True, timeout=300)
        line = json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else {"error": r.stderr[-200:]}
        rec["e2e_threads%d_blobs_per_s" % t] = line.get("per_s")
        rec["threads%d_mean_batch" % t] = line.get("mean_batch")
        rec["threads%d_bit_exact" % t] = bool(r.returncode == 0 and line.get("mismatches") == 0 and line.get("errors") == 0)
    out["compute_cells_and_kzg_proofs"] = rec
    # verify_blob_kzg_proof one at a time from 1 / 16 threads: concurrent requests are checked as one batch
    rec = {}
    for t in (1, 16):
        r = subprocess.run([exe, setup, "verify", str(t), "30", "4"], capture_output=True, text=True, timeout=300)
        line = json.loads(r.stdout.strip().splitlines()[-1]) if r.stdout.strip() else {"error": r.stderr[-200:]}
        rec["e2e_threads%d_per_s" % t] = line.get("per_s")
        rec["threads%d_mean_batch" % t] = line.get("mean_batch")
        rec["threads%d_all_true" % t] = bool(r.returncode == 0 and line.get("mismatches") == 0 and line.get("errors") == 0)
    out["verify_blob_kzg_proof"] = rec
    return out


def blob_metrics(B, K, osettings, torch):
    """the other half of BASELINE's metric: blobs/s (batch 64, BASELINE config 3), adversarial blobs, EIP-7594, verification"""
    ex = {}
    rng = np.random.default_rng(SEED)
    nb = 64
    blobs = rand_blobs(rng, nb)
    ts = B.KZGSettings.load_trusted_setup_file()
    h_blobs = torch.from_numpy(blobs).pin_memory()
    d_blobs = h_blobs.cuda()
    d_out = torch.zeros((nb, 48), dtype=torch.uint8, device="cuda")
    d_y = torch.zeros((nb, 32), dtype=torch.uint8, device="cuda")
    d_status = torch.zeros(nb, dtype=torch.int32, device="cuda")
    timed = lambda fn, reps=10, warm=3: timed_events(torch, fn, reps, warm)

    ms = timed(lambda: ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), nb, d_status.data_ptr(), 0))
    comm = d_out.cpu().numpy()
    osettings.set_threads(os.cpu_count() or 1)
    ok = all(comm[i].tobytes() == K.blob_to_kzg_commitment(blobs[i].tobytes(), osettings) for i in (0, 17, 63))
    ex["blob_to_kzg_commitment"] = {"blobs_per_s": nb / (ms * 1e-3), "ms_per_batch": ms, "batch": nb, "parity_ok": bool(ok),
                                    "launches_per_batch": ts.launches()}
    uniform_ms = ms
    # adversarial blob contents on the same 64-blob batch (timed, parity-checked): all elements equal (the all-0x02
    # golden blob), elements < 2^64, 10 % zero elements, all elements r - 1
    adv = {"uniform_ms": uniform_ms}
    worst = 0.0
    for kind in ("all_equal", "below_2^64", "ten_percent_zero", "r_minus_1"):
        if kind == "all_equal":
            one = np.zeros((4096, 32), np.uint8)
            one[:, 31] = 2
        elif kind == "r_minus_1":
            one = np.tile(np.frombuffer((R_MOD - 1).to_bytes(32, "big"), np.uint8), (4096, 1))
        elif kind == "below_2^64":
            one = np.zeros((4096, 32), np.uint8)
            one[:, 24:] = rng.integers(0, 256, size=(4096, 8), dtype=np.uint8)
        else:
            one = rand_blobs(rng, 1).reshape(4096, 32).copy()
            one[rng.random(4096) < 0.1] = 0
        ab = np.ascontiguousarray(np.tile(one.reshape(1, -1), (nb, 1)))
        d_ab = torch.from_numpy(ab).cuda()
        ms_a = timed(lambda: ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_ab.data_ptr(), nb, d_status.data_ptr(), 0), reps=5, warm=2)
        ok_a = d_out[3].cpu().numpy().tobytes() == K.blob_to_kzg_commitment(ab[3].tobytes(), osettings) and int(d_status.sum().item()) == 0
        adv[kind] = {"ms_per_batch": ms_a, "vs_uniform": ms_a / uniform_ms, "parity_ok": bool(ok_a)}
        worst = max(worst, ms_a / uniform_ms)
    adv["worst_vs_uniform"] = worst
    ex["adversarial_blobs"] = adv
    # e2e through the C ABI with pinned / pageable host blobs
    h_out = torch.zeros((nb, 48), dtype=torch.uint8).pin_memory()
    reps = 5
    ts.blob_to_kzg_commitment_batch_ptr(h_out.data_ptr(), h_blobs.data_ptr(), nb)
    t0 = time.perf_counter()
    for _ in range(reps):
        ts.blob_to_kzg_commitment_batch_ptr(h_out.data_ptr(), h_blobs.data_ptr(), nb)
    dt = (time.perf_counter() - t0) / reps
    ex["blob_to_kzg_commitment"]["e2e_blobs_per_s"] = nb / dt
    ex["blob_to_kzg_commitment"]["e2e_parity_ok"] = bool(np.array_equal(h_out.numpy(), comm))
    pg_out = np.zeros((nb, 48), np.uint8)
    ts.blob_to_kzg_commitment_batch(blobs, pg_out)
    t0 = time.perf_counter()
    for _ in range(reps):
        ts.blob_to_kzg_commitment_batch(blobs, pg_out)
    ex["blob_to_kzg_commitment"]["e2e_pageable_blobs_per_s"] = nb / ((time.perf_counter() - t0) / reps)
    ex["blob_to_kzg_commitment"]["e2e_pageable_parity_ok"] = bool(np.array_equal(pg_out, comm))
    # compute_kzg_proof on device-resident blobs (z given), compute_blob_kzg_proof end to end (incl. host SHA-256)
    zs = rand_blobs(rng, 1)[0, :32 * nb].reshape(nb, 32).copy()
    d_z = torch.from_numpy(zs).cuda()
    ms = timed(lambda: ts.compute_kzg_proof_device(d_out.data_ptr(), d_y.data_ptr(), d_blobs.data_ptr(), d_z.data_ptr(), nb,
                                                   d_status.data_ptr(), 0, 0))
    p0, y0 = K.compute_kzg_proof(blobs[5].tobytes(), zs[5].tobytes(), osettings)
    ex["compute_kzg_proof"] = {"blobs_per_s": nb / (ms * 1e-3), "ms_per_batch": ms, "batch": nb,
                               "parity_ok": bool(d_out[5].cpu().numpy().tobytes() == p0 and d_y[5].cpu().numpy().tobytes() == y0)}
    h_comm = torch.from_numpy(comm.copy()).pin_memory()
    ts.compute_blob_kzg_proof_batch_ptr(h_out.data_ptr(), h_blobs.data_ptr(), h_comm.data_ptr(), nb)
    t0 = time.perf_counter()
    for _ in range(reps):
        ts.compute_blob_kzg_proof_batch_ptr(h_out.data_ptr(), h_blobs.data_ptr(), h_comm.data_ptr(), nb)
    dt = (time.perf_counter() - t0) / reps
    pb = K.compute_blob_kzg_proof(blobs[9].tobytes(), comm[9].tobytes(), osettings)
    ex["compute_blob_kzg_proof"] = {"e2e_blobs_per_s": nb / dt, "ms_per_batch": dt * 1e3, "batch": nb,
                                    "parity_ok": bool(h_out[9].numpy().tobytes() == pb)}
    # single-blob latency through the c-kzg entry point (52.4 ms blst 1 core, 6.35 ms 16 cores, 5.21 ms L4+sppark)
    one = h_blobs[0].numpy().tobytes()
    for _ in range(3):
        ts.blob_to_kzg_commitment(one)
    t0 = time.perf_counter()
    for _ in range(10):
        c1 = ts.blob_to_kzg_commitment(one)
    ex["blob_to_kzg_commitment"]["single_blob_ms"] = (time.perf_counter() - t0) / 10 * 1e3
    ex["blob_to_kzg_commitment"]["single_blob_parity_ok"] = bool(c1 == comm[0].tobytes())
    # sustained rate on a stream of 512 blobs through the same host-pointer calls: chunks of 64 rotate over the lanes, so
    # the latency-bound tail of one chunk overlaps the accumulation of the next
    big = 512
    h_big = torch.from_numpy(np.tile(blobs, (big // nb, 1))).pin_memory()
    h_big_out = torch.zeros((big, 48), dtype=torch.uint8).pin_memory()
    ts.blob_to_kzg_commitment_batch_ptr(h_big_out.data_ptr(), h_big.data_ptr(), big)
    t0 = time.perf_counter()
    ts.blob_to_kzg_commitment_batch_ptr(h_big_out.data_ptr(), h_big.data_ptr(), big)
    dt = time.perf_counter() - t0
    ex["blob_to_kzg_commitment"]["e2e_stream512_blobs_per_s"] = big / dt
    ex["blob_to_kzg_commitment"]["e2e_stream512_parity_ok"] = bool(np.array_equal(h_big_out.numpy()[:nb], comm) and
                                                                   np.array_equal(h_big_out.numpy()[-nb:], comm))
    h_big_comm = torch.from_numpy(np.tile(comm, (big // nb, 1))).pin_memory()
    h_big_proof = torch.zeros((big, 48), dtype=torch.uint8).pin_memory()
    ts.compute_blob_kzg_proof_batch_ptr(h_big_proof.data_ptr(), h_big.data_ptr(), h_big_comm.data_ptr(), big)
    t0 = time.perf_counter()
    ts.compute_blob_kzg_proof_batch_ptr(h_big_proof.data_ptr(), h_big.data_ptr(), h_big_comm.data_ptr(), big)
    dt = time.perf_counter() - t0
    ex["compute_blob_kzg_proof"]["e2e_stream512_blobs_per_s"] = big / dt
    ex["compute_blob_kzg_proof"]["e2e_stream512_parity_ok"] = bool(np.array_equal(h_big_proof.numpy()[:nb], h_out.numpy()))
    # CPU beside it: the oracle port, one blob per thread-parallel MSM
    t0 = time.perf_counter()
    for i in range(4):
        K.blob_to_kzg_commitment(blobs[i].tobytes(), osettings)
    ex["blob_to_kzg_commitment"]["cpu_port_blobs_per_s"] = 4 / (time.perf_counter() - t0)
    ex["blob_to_kzg_commitment"]["cpu_cores"] = os.cpu_count()
    # EIP-7594 producer: cells (NTT-8192) and FK20 cell proofs, 64 blobs, host numpy in and out
    try:
        ts.compute_cell_proofs_batch(blobs[:1])           # builds the 128 x 64 FK20 table on first use
        ts.compute_cells_batch(blobs)                     # first call allocates the device / NTT workspaces
        t0 = time.perf_counter()
        for _ in range(3):
            cells = ts.compute_cells_batch(blobs)
        dt_cells = (time.perf_counter() - t0) / 3
        t0 = time.perf_counter()
        for _ in range(3):
            proofs = ts.compute_cell_proofs_batch(blobs)
        dt_proofs = (time.perf_counter() - t0) / 3
        # both outputs in one pass (b200_compute_cells_and_kzg_proofs_batch) into caller-owned arrays
        cells_buf = np.zeros((nb, 128, 2048), np.uint8)
        proofs_buf = np.zeros((nb, 128, 48), np.uint8)
        ts.compute_cells_and_kzg_proofs_batch(blobs, cells_buf, proofs_buf)
        t0 = time.perf_counter()
        for _ in range(3):
            ts.compute_cells_and_kzg_proofs_batch(blobs, cells_buf, proofs_buf)
        dt_both = (time.perf_counter() - t0) / 3
        both_ok = bool(np.array_equal(cells_buf, cells) and np.array_equal(proofs_buf, proofs))
        t0 = time.perf_counter()
        oc, op = K.compute_cells_and_kzg_proofs(blobs[3].tobytes(), osettings)
        dt_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        K.compute_cells_and_kzg_proofs(blobs[4].tobytes(), osettings)
        dt_cpu = min(dt_cpu, time.perf_counter() - t0)   # the first call also builds the oracle's FK20 columns
        ex["compute_cells_and_kzg_proofs"] = {
            "blobs_per_s": nb / dt_both, "ms_per_batch": dt_both * 1e3, "cells_only_ms_per_batch": dt_cells * 1e3,
            "proofs_only_ms_per_batch": dt_proofs * 1e3, "batch": nb,
            "parity_ok": bool(both_ok and np.asarray(cells[3]).tobytes() == b"".join(oc) and np.asarray(proofs[3]).tobytes() == b"".join(op)),
            "cpu_port_blobs_per_s": 1.0 / dt_cpu, "cpu_cores": os.cpu_count()}
    except Exception as e:  # an extra: never lose the headline line over it
        ex["compute_cells_and_kzg_proofs"] = {"error": repr(e)[:200]}
    # verification side: pairing on the device, host byte arrays in, bool out; CPU port beside it
    try:
        proofs64 = h_out.numpy().copy()                      # compute_blob_kzg_proof results of the 64 blobs above
        ok_all = ts.verify_blob_kzg_proof_batch(blobs, comm, proofs64)
        t0 = time.perf_counter()
        for _ in range(reps):
            ts.verify_blob_kzg_proof_batch(blobs, comm, proofs64)
        dt = (time.perf_counter() - t0) / reps
        badp = proofs64.copy()
        badp[7] = proofs64[8]
        rej = ts.verify_blob_kzg_proof_batch(blobs, comm, badp)
        t0 = time.perf_counter()
        ts.verify_blob_kzg_proof_batch(h_big.numpy(), h_big_comm.numpy(), h_big_proof.numpy())
        dt_big = time.perf_counter() - t0
        ts.verify_blob_kzg_proof(blobs[3], comm[3], proofs64[3])          # first call allocates the coalescer's pinned staging
        t0 = time.perf_counter()
        for _ in range(reps):
            one_ok = ts.verify_blob_kzg_proof(blobs[3], comm[3], proofs64[3])
        dt_one = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        cpu_ok = K.verify_blob_kzg_proof_batch([blobs[i].tobytes() for i in range(8)], [comm[i].tobytes() for i in range(8)],
                                               [proofs64[i].tobytes() for i in range(8)], osettings)
        dt_cpu = time.perf_counter() - t0
        ex["verify_blob_kzg_proof_batch"] = {
            "e2e_blobs_per_s": nb / dt, "ms_per_batch": dt * 1e3, "batch": nb, "accepts_valid": bool(ok_all and one_ok and cpu_ok),
            "rejects_swapped_proof": bool(rej is False), "e2e_stream512_blobs_per_s": big / dt_big,
            "single_verify_blob_kzg_proof_ms": dt_one * 1e3, "cpu_port_blobs_per_s": 8 / dt_cpu, "cpu_cores": 1}
    except Exception as e:
        ex["verify_blob_kzg_proof_batch"] = {"error": repr(e)[:200]}
    ts.free()
    return ex


def ntt_metrics(B, K, osettings, torch):
    """Fr NTT and DAS extension at EVERY size 2^12..2^20 (BASELINE config 4), each checked against the oracle; fft_g1 2^15"""
    ex = {}
    rng = np.random.default_rng(SEED + 9)
    cores = os.cpu_count() or 1
    fs = B.FFTSettings(20)
    ofs = K.FFTSettings(20)
    imad = None
    try:
        imad = B.microbench_int()["imad_per_s"]
    except Exception:
        pass
    ntt = {}
    for logn in range(12, 21):
        m = 1 << logn
        data = rand_fr(rng, m)
        d_in = torch.from_numpy(data.view(np.int64)).cuda()
        d_o = torch.zeros_like(d_in)
        ms = timed_events(torch, lambda: fs.fft_fr_device(d_o.data_ptr(), d_in.data_ptr(), m, False, 1, 0))
        rec = {"ms": ms, "elements_per_s": m / (ms * 1e-3), "butterflies_per_s": (m // 2) * logn / (ms * 1e-3),
               "hbm_gbs_algorithmic": 64 * m * (2 if logn > 11 else 1) / (ms * 1e-3) / 1e9,
               "parity_ok": bool(np.array_equal(d_o.cpu().numpy().view(np.uint64).reshape(m, 4), ofs.fft_fr(data, False, nthreads=cores)))}
        if imad:
            # integer roofline of the transform: (n/2 log n butterflies + n inter-pass twiddles) x 136 IMAD.WIDE per Fr
            # multiplication against the measured IMAD.WIDE peak
            rec["int_roofline_frac"] = ((m // 2) * logn + m) * 136 / (ms * 1e-3) / imad
        ms_i = timed_events(torch, lambda: fs.fft_fr_device(d_in.data_ptr(), d_o.data_ptr(), m, True, 1, 0), reps=3, warm=1)
        rec["inverse_ms"] = ms_i
        rec["roundtrip_ok"] = bool(np.array_equal(d_in.cpu().numpy().view(np.uint64).reshape(m, 4), data))
        ntt["2^%d" % logn] = rec
    ex["fft_fr"] = ntt
    das = {}
    for logn in range(12, 21):   # n/2 even-index evaluations -> n/2 odd-index ones (data_availability_sampling.rs:78-100)
        h = 1 << (logn - 1)
        evens = rand_fr(rng, h)
        d_in = torch.from_numpy(evens.view(np.int64)).cuda()
        d_o = torch.zeros_like(d_in)
        ms = timed_events(torch, lambda: fs.das_fft_extension_device(d_o.data_ptr(), d_in.data_ptr(), h, 1, 0))
        das["2^%d" % logn] = {"ms": ms, "evens": h, "elements_per_s": h / (ms * 1e-3),
                              "parity_ok": bool(np.array_equal(d_o.cpu().numpy().view(np.uint64).reshape(h, 4), ofs.das_fft_extension(evens)))}
    ex["das_fft_extension"] = das
    # fft_g1 at the reference's own bench size, scale 15 (18.84 s on one core, 4.96 s on 16); input = the 4096
    # monomial setup points tiled.  Checked by the slow-DFT identity out[k] = sum_j w^(jk) P_j on sampled outputs (each a
    # 2^15-term CPU MSM of the oracle) and against the oracle's own fft_g1 at 2^7.
    try:
        g1m = osettings.g1_monomial
        nn = 1 << 15
        pts = np.ascontiguousarray(np.tile(g1m, (nn // 4096, 1)))
        d_pts = torch.from_numpy(pts.view(np.int64)).cuda()
        d_res = torch.zeros_like(d_pts)
        ms = timed_events(torch, lambda: fs.fft_g1_device(d_res.data_ptr(), d_pts.data_ptr(), nn, False, 1, 0), reps=3, warm=1)
        res = d_res.cpu().numpy().view(np.uint64).reshape(nn, 18)
        roots = ofs.roots_of_unity                      # (2^20 + 1) x 4, Montgomery
        stride = (1 << 20) // nn
        aff = K.p1s_to_affine(pts)
        dft_ok = True
        for k in (1, 12345, nn - 1):
            idx = (np.arange(nn, dtype=np.int64) * k % nn) * stride
            want = K.msm_affine(aff, np.ascontiguousarray(roots[idx]), nthreads=cores)
            dft_ok = dft_ok and K.p1_compress(res[k]) == K.p1_compress(want)
        small = fs.fft_g1(g1m[:128], False)
        want = ofs.fft_g1(g1m[:128], False)
        ok = all(K.p1_compress(small[i]) == K.p1_compress(want[i]) for i in range(128))
        ex["fft_g1"] = {"2^15": {"ms": ms, "points_per_s": nn / (ms * 1e-3), "slow_dft_identity_ok": bool(dft_ok),
                                 "checked_outputs": [1, 12345, nn - 1]}, "parity_ok_2^7": bool(ok)}
    except Exception as e:
        ex["fft_g1"] = {"error": repr(e)[:200]}
    fs.close()
    return ex


if __name__ == "__main__":
    main()
