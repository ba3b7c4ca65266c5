"""A plain-C program (examples/ckzg_roundtrip.c) links libb200kzg.so through include/b200_kzg.h and drives the c-kzg-4844
entry points the way a language binding would.  CPU: it compiles, links and fails loudly without a device (no fallback).
GPU: commit -> prove -> verify and cells -> recover -> verify_cells round trips succeed."""
import os
import subprocess

import pytest

from conftest import ROOT, SETUP_PATH

EXE = "/tmp/b200_ckzg_roundtrip"
EXE_T = "/tmp/b200_ckzg_threads"


def _build():
    lib_dir = os.path.join(ROOT, "rust-kzg_b200")
    if not os.path.exists(os.path.join(lib_dir, "libb200kzg.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ckzg_roundtrip.c"), "-L" + lib_dir, "-lb200kzg",
                           "-Wl,-rpath," + lib_dir, "-o", EXE])
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-pthread", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ckzg_threads.c"), "-L" + lib_dir, "-lb200kzg",
                           "-Wl,-rpath," + lib_dir, "-o", EXE_T])


def test_c_consumer_builds_and_refuses_without_gpu():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE, SETUP_PATH], capture_output=True, text=True)
    assert r.returncode == 3 and "load_trusted_setup_file" in r.stderr     # C_KZG_ERROR: no device, no CPU path
    r = subprocess.run([EXE_T, SETUP_PATH, "commit", "4", "2"], capture_output=True, text=True)
    assert r.returncode == 3 and "load_trusted_setup_file" in r.stderr


@pytest.mark.gpu
def test_c_consumer_round_trips():
    _build()
    r = subprocess.run([EXE, SETUP_PATH], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("op,threads", [("commit", 16), ("blob_proof", 8), ("proof", 8), ("mixed", 12), ("cells", 6), ("verify", 8)])
def test_c_pthread_consumer_is_bit_exact_under_coalescing(op, threads):
    """examples/ckzg_threads.c: N pthreads call the unmodified single-blob symbols on pageable blobs; concurrent requests
    share launch sequences (csrc/coalesce.cuh) and every result must equal the one computed with a single thread active
    (cells: compute_cells_and_kzg_proofs, 128 cells + 128 proofs per call; verify: verify_blob_kzg_proof, all true)"""
    import json
    _build()
    r = subprocess.run([EXE_T, SETUP_PATH, op, str(threads), "6", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["mismatches"] == 0 and line["errors"] == 0 and line["isolation_failures"] == 0
    assert line["calls"] == threads * 6
