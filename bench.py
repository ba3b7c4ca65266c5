#!/usr/bin/env python3
"""bench.py -- headline benchmark of the rust-kzg hot path on B200 (contract in the task statement).

A "step" is one pass of the hot path over one batch of synthetic input.  Workload at every N (weak scaling):
BASELINE.json configs[1], an MSM of 2^20 uniformly random Fr scalars x trusted-setup G1 points per GPU, bit-exact
vs the CPU oracle.  For N > 1 the 2^20*N terms are sharded by rank: every rank runs its local MSM,
the 144-byte partial results are all-gathered over NCCL and summed locally (NCCL cannot add curve points).

Unit: "G1-adds/s" = canonical bucket additions per second = 16 per term (canonical c = 16 ->
16 mixed additions per term), so both arms (this one and --impl reference) are counted identically.  points/s =
value / 16.  `value` has scalars resident in HBM; `e2e` goes through the reference-facing C ABI call
mult_pippenger_prepared with pinned HOST scalars (H2D + D2H inside the timed region).

Extra keys report the other BASELINE metric (blobs/s for blob_to_kzg_commitment / compute_blob_kzg_proof, batch 64)
and the Fr NTT at 2^20.

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
ADDS_PER_TERM = 16          # canonical c = 16
BYTES_PER_TERM = 128        # 32 B scalar + 96 B affine point
IMAD_PER_ADD = 10 * 300     # mixed add = 8M + 2S ~ 10 Fp mul; Fp mul = 144 + 144 + 12 32x32 multiply-adds
SEED = 0x4B5A47
METRIC = "MSM G1-adds/sec at 2^20 (per GPU, weak scaling); blobs/sec blob_to_kzg_commitment in extra"


def rand_fr(rng, n):
    """uniform-looking Montgomery Fr limbs below 2^254 < r (numpy PCG64, seed SEED)"""
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    return a


def rand_blobs(rng, n):
    b = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    b[:, :, 0] = 0   # canonical field elements (kzg-bench/src/tests/eip_4844.rs:28-37)
    return b.reshape(n, 131072)


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


def folded_expectation(K, L, sc, nthreads):
    """exact oracle for tiled bases P_i = L[i mod 4096]: fold the scalars, then a 4096-term CPU MSM"""
    folded = sc[:4096].copy()
    for k in range(1, sc.shape[0] // 4096):
        folded = K.fr_add(folded, sc[k * 4096:(k + 1) * 4096])
    return K.msm_affine(L, folded, nthreads=nthreads)


def cpu_baseline(K, L, n_terms, nthreads, steps=1, warmup=0):
    """the oracle's restatement of tiling_parallel_pippenger on the host cores over the first n_terms terms"""
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
    host threads, a bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, s, L = load_bases()
    cores = os.cpu_count() or 1
    # size the sample so that (steps + warmup) passes end within a few minutes: ~24 us of CPU per term per core
    budget_s = 90.0
    per_term = 30e-6 / cores
    n = 1 << 18
    while n > (1 << 12) and n * per_term * (args.steps + args.warmup) > budget_s:
        n >>= 1
    value, dt = cpu_baseline(K, L, n, cores, steps=args.steps, warmup=args.warmup)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "G1-adds/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "MSM over the first 2^%d of 2^%d random Fr scalars x tiled EIP-4844 trusted-setup G1 points, "
                               "CPU, all host threads" % (n.bit_length() - 1, LOG_N), "seed": SEED},
        "cpu_baseline": {"value": value, "unit": "G1-adds/s", "cores": cores, "kind": "port",
                         "sample": "first 2^%d terms of the 2^%d-term MSM, tiling_parallel_pippenger restatement "
                                   "(oracle/kzg_oracle.c), not blst assembly" % (n.bit_length() - 1, LOG_N)},
        "e2e": {"value": value, "unit": "G1-adds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extra", action="store_true", help="skip the blob / NTT extra metrics")
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

    # ---- inputs: rank r owns terms [r*n, (r+1)*n) of the N*n-term MSM; bases tile the 4096 setup points ----------
    rng = np.random.default_rng(SEED + rank)
    sc = rand_fr(rng, n)
    pts = np.tile(L, (n // 4096, 1)) if n >= 4096 else L[:n]
    msm = B.PreparedMsm(pts)
    h_sc = torch.from_numpy(sc.view(np.int64)).pin_memory()
    d_sc = h_sc.cuda(non_blocking=True)
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    d_all = torch.zeros((world, 18), dtype=torch.int64, device="cuda")
    d_total = torch.zeros(18, dtype=torch.int64, device="cuda")
    stream = 0  # the library launches on the legacy default stream = torch's current stream, so torch events see it

    def step():
        msm.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_out)
            B.g1_sum_device(d_total.data_ptr(), d_all.data_ptr(), world, stream)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples from here are dropped at mark(); kept ones span the timed region + e2e leg
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    info = msm.info()
    # parity of the thing being timed: local result vs the folded-scalar oracle
    exp = folded_expectation(K, L, sc, os.cpu_count() or 1)
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
    n_entries, n_tasks = msm.last_counts()   # non-zero digits sorted into buckets / accumulate tasks of the last step
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n * ADDS_PER_TERM / (ms_step * 1e-3)

    # ---- e2e: the reference-facing call with host scalars (pinned), H2D + D2H inside the timed region -----------
    h_np = h_sc.numpy().view(np.uint64).reshape(n, 4)
    for _ in range(2):
        msm.mult(h_np)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        res = msm.mult(h_np)
    t_e2e = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([t_e2e], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    if K.p1_compress(res) != K.p1_compress(exp):
        raise SystemExit("bench.py: e2e MSM result differs from the oracle")
    e2e_value = world * n * ADDS_PER_TERM / t_e2e
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_accumulate), measured live with CUDA events on its stream -----------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    acc_ms = acc_ms_sum / max(acc_runs, 1)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "accumulate_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    achieved_gbs = BYTES_PER_TERM * n / (acc_ms * 1e-3) / 1e9 if acc_ms else None
    mb = B.microbench_int()
    adds_per_launch = n_entries - n_tasks   # a task's first point is a load, every other entry one mixed addition
    roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak if achieved_gbs else None, "traffic": traffic,
                "kernel": "k_accumulate", "kernel_ms": acc_ms, "kernel_share_of_step": acc_ms / ms_step if acc_ms else None,
                "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback",
                "note": "the path is bound by the integer multiply pipe, not HBM (SURVEY.md 8d); see int_roofline"}
    int_roofline = {"bound": "FMA-heavy pipe (IMAD.WIDE.U32, 32x32+64 multiply-add)", "unit": "T multiply-adds/s",
                    "achieved": adds_per_launch * IMAD_PER_ADD / (acc_ms * 1e-3) / 1e12 if acc_ms else None,
                    "peak": mb["imad_per_s"] / 1e12, "peak_source": "b200_microbench_int (dependent-operand IMAD.WIDE loop), measured in this run",
                    "fp_mul_per_s_measured": mb["fpmul_per_s"],
                    "adds_per_launch": adds_per_launch, "entries": n_entries, "tasks": n_tasks,
                    "note": "achieved counts 3000 algorithmic multiply-adds per bucket addition (10 Fp mul x 300) over the "
                            "additions k_accumulate actually performs (entries - tasks, read back from the device); every "
                            "IMAD.WIDE form issues at 32/clk/SM on sm_100a (measured, with or without carry), ncu shows "
                            "the FMA-heavy pipe 86% busy in k_accumulate (profiles/r01_accumulate_full.md)"}
    if int_roofline["achieved"]:
        int_roofline["frac"] = int_roofline["achieved"] / int_roofline["peak"]

    # ---- CPU baseline beside it: oracle port on the host cores, bounded sample ------------------------------------
    cores = os.cpu_count() or 1
    n_cpu = 1 << min(LOG_N, 20 if cores >= 8 else 18)   # ~10-30 core-seconds of CPU work
    cpu_value, cpu_dt = cpu_baseline(K, L, n_cpu, cores)
    cpu = {"value": cpu_value, "unit": "G1-adds/s", "cores": cores, "kind": "port",
           "sample": "first 2^%d terms of the 2^%d-term MSM, %.2f s on %d threads; C restatement of "
                     "tiling_parallel_pippenger (oracle/kzg_oracle.c), not blst assembly" % (n_cpu.bit_length() - 1, LOG_N, cpu_dt, cores)}

    out = {
        "metric": METRIC, "value": value, "unit": "G1-adds/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": "MSM 2^%d random Fr scalars x EIP-4844 trusted-setup G1 points (4096 Lagrange points tiled), "
                               "per GPU; prepared fixed-base table resident in HBM" % LOG_N,
                   "points_per_s": value / ADDS_PER_TERM,
                   "adds_unit": "16 canonical bucket additions per term (BASELINE.md section 2) for every arm, whatever "
                                "window width the engine uses; int_roofline counts the additions actually performed",
                   "window_bits": info["c"], "windows": info["W"],
                   "table_bytes": info["table_bytes"], "seed": SEED, "prng": "numpy PCG64",
                   "l2": "inputs larger than L2: %d MiB scalars + %.1f GiB table per step" % (n * 32 >> 20, info["table_bytes"] / 2 ** 30),
                   "parity": "compressed result == folded-scalar oracle (checked before timing)",
                   "multi_gpu": "terms sharded by rank, NCCL all-gather of 144 B partial results + local add" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "G1-adds/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 144,
                "ms_per_step": t_e2e * 1e3, "api": "mult_pippenger_prepared (C ABI), pinned host scalars"},
        "gpu_launches": (info["launches"] + (1 if world > 1 else 0)) * args.steps,
        "roofline": roofline, "int_roofline": int_roofline, "cpu_baseline": cpu,
    }
    if not args.no_extra and world == 1:
        try:
            # variable-base call of the reference's L4 benchmark shape (bench_g1_lincomb, points + scalars travel with
            # every call: kzg-bench/src/benches/lincomb.rs:35-46; 116.75 ms on L4 + sppark)
            h_pts = torch.from_numpy(pts.view(np.int64)).pin_memory()
            hp = h_pts.numpy().view(np.uint64).reshape(n, 12)
            B.mult_pippenger(hp, h_np)
            t0 = time.perf_counter()
            for _ in range(3):
                rv = B.mult_pippenger(hp, h_np)
            t_var = (time.perf_counter() - t0) / 3
            out["variable_base_e2e"] = {"api": "mult_pippenger (C ABI), pinned host points + scalars",
                                        "ms_per_call": t_var * 1e3, "points_per_s": n / t_var,
                                        "h2d_bytes_per_call": 128 * n, "parity_ok": bool(K.p1_compress(rv) == K.p1_compress(exp))}
            del h_pts, hp
        except Exception as e:
            out["variable_base_e2e"] = {"error": repr(e)}
        msm.close()
        try:
            out["extra"] = extra_metrics(B, K, osettings, torch)
        except Exception as e:  # extras must not take the headline down
            out["extra"] = {"error": repr(e)}
    if world > 1:
        dist.destroy_process_group()
    print(json.dumps(out))


def extra_metrics(B, K, osettings, torch):
    """the other half of BASELINE's metric: blobs/s (batch 64, BASELINE config 3) and the Fr NTT (config 4)"""
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

    def timed(fn, reps=10, warm=3):
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

    ms = timed(lambda: ts.blob_to_kzg_commitment_device(d_out.data_ptr(), d_blobs.data_ptr(), nb, d_status.data_ptr(), 0))
    comm = d_out.cpu().numpy()
    osettings.set_threads(os.cpu_count() or 1)
    ok = all(comm[i].tobytes() == K.blob_to_kzg_commitment(blobs[i].tobytes(), osettings) for i in (0, 17, 63))
    ex["blob_to_kzg_commitment"] = {"blobs_per_s": nb / (ms * 1e-3), "ms_per_batch": ms, "batch": nb, "parity_ok": bool(ok),
                                    "launches_per_batch": ts.launches()}
    # e2e through the C ABI with pinned host blobs
    h_out = torch.zeros((nb, 48), dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    reps = 5
    ts.blob_to_kzg_commitment_batch_ptr(h_out.data_ptr(), h_blobs.data_ptr(), nb)
    t0 = time.perf_counter()
    for _ in range(reps):
        ts.blob_to_kzg_commitment_batch_ptr(h_out.data_ptr(), h_blobs.data_ptr(), nb)
    dt = (time.perf_counter() - t0) / reps
    ex["blob_to_kzg_commitment"]["e2e_blobs_per_s"] = nb / dt
    ex["blob_to_kzg_commitment"]["e2e_parity_ok"] = bool(np.array_equal(h_out.numpy(), comm))
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
    # sustained rate on a stream of 512 blobs through the same host-pointer calls: chunks of 64 alternate between two
    # lanes, so the latency-bound tail of one chunk overlaps the accumulation of the next
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
        t0 = time.perf_counter()
        oc, op = K.compute_cells_and_kzg_proofs(blobs[3].tobytes(), osettings)
        dt_cpu = time.perf_counter() - t0
        t0 = time.perf_counter()
        K.compute_cells_and_kzg_proofs(blobs[4].tobytes(), osettings)
        dt_cpu = min(dt_cpu, time.perf_counter() - t0)   # the first call also builds the oracle's FK20 columns
        ex["compute_cells_and_kzg_proofs"] = {
            "blobs_per_s": nb / (dt_cells + dt_proofs), "cells_ms_per_batch": dt_cells * 1e3, "proofs_ms_per_batch": dt_proofs * 1e3,
            "batch": nb, "parity_ok": bool(np.asarray(cells[3]).tobytes() == b"".join(oc) and np.asarray(proofs[3]).tobytes() == b"".join(op)),
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
    # Fr NTT sweep (BASELINE config 4)
    fs = B.FFTSettings(20)
    ofs = K.FFTSettings(20)
    ntt = {}
    for logn in (12, 16, 20):
        m = 1 << logn
        data = rand_fr(rng, m)
        d_in = torch.from_numpy(data.view(np.int64)).cuda()
        d_o = torch.zeros_like(d_in)
        ms = timed(lambda: fs.fft_fr_device(d_o.data_ptr(), d_in.data_ptr(), m, False, 1, 0))
        rec = {"ms": ms, "elements_per_s": m / (ms * 1e-3), "butterflies_per_s": (m // 2) * logn / (ms * 1e-3),
               "hbm_gbs_algorithmic": 64 * m * (2 if logn > 11 else 1) / (ms * 1e-3) / 1e9}
        if logn == 20:
            # integer roofline of the transform: (n/2 log n butterflies + n inter-pass twiddles) x 136 IMAD.WIDE per Fr
            # multiplication against the measured IMAD.WIDE peak
            try:
                imad = B.microbench_int()["imad_per_s"]
                rec["int_roofline_frac"] = ((m // 2) * logn + m) * 136 / (ms * 1e-3) / imad
            except Exception:
                pass
        if logn <= 16:
            rec["parity_ok"] = bool(np.array_equal(d_o.cpu().numpy().view(np.uint64).reshape(m, 4), ofs.fft_fr(data, False, nthreads=os.cpu_count() or 1)))
        ntt["2^%d" % logn] = rec
    ex["fft_fr"] = ntt
    # DAS extension sweep (BASELINE config 4): n/2 even-index evaluations -> n/2 odd-index ones (data_availability_sampling.rs:78-100)
    das = {}
    for logn in (12, 16, 20):
        h = 1 << (logn - 1)
        evens = rand_fr(rng, h)
        d_in = torch.from_numpy(evens.view(np.int64)).cuda()
        d_o = torch.zeros_like(d_in)
        ms = timed(lambda: fs.das_fft_extension_device(d_o.data_ptr(), d_in.data_ptr(), h, 1, 0))
        rec = {"ms": ms, "evens": h, "elements_per_s": h / (ms * 1e-3)}
        if logn <= 16:
            rec["parity_ok"] = bool(np.array_equal(d_o.cpu().numpy().view(np.uint64).reshape(h, 4), ofs.das_fft_extension(evens)))
        das["2^%d" % logn] = rec
    ex["das_fft_extension"] = das
    # fft_g1 at the reference's own bench size, scale 15 (18.84 s on one core, 4.96 s on 16); input = the 4096
    # monomial setup points tiled; parity against the oracle at 2^7
    try:
        g1m = osettings.g1_monomial
        pts = np.ascontiguousarray(np.tile(g1m, (8, 1)))
        d_pts = torch.from_numpy(pts.view(np.int64)).cuda()
        d_res = torch.zeros_like(d_pts)
        ms = timed(lambda: fs.fft_g1_device(d_res.data_ptr(), d_pts.data_ptr(), 1 << 15, False, 1, 0), reps=3, warm=1)
        small = fs.fft_g1(g1m[:128], False)
        want = ofs.fft_g1(g1m[:128], False)
        ok = all(K.p1_compress(small[i]) == K.p1_compress(want[i]) for i in range(128))
        ex["fft_g1"] = {"2^15": {"ms": ms, "points_per_s": (1 << 15) / (ms * 1e-3)}, "parity_ok_2^7": bool(ok)}
    except Exception as e:
        ex["fft_g1"] = {"error": repr(e)[:200]}
    fs.close()
    return ex


if __name__ == "__main__":
    main()
