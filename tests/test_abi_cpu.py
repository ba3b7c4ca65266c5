"""The C-ABI library loads and exports every symbol include/*.h declares; compute entry points fail LOUDLY without a
GPU (no CPU fallback); host-side logic that needs no device (SHA-256, setup parsing, sharding) works.  No GPU needed."""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "b200_kzg.h")


@pytest.fixture(scope="module")
def B():
    import rust_kzg_b200
    if not os.path.exists(rust_kzg_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return rust_kzg_b200


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(\w+)\s*\(", src)))


def test_header_symbols_are_exported(B):
    names = _declared_symbols()
    assert len(names) >= 35
    # the reference-facing names must be exactly the ones the reference's FFI binds
    for must in ("prepare_msm", "mult_pippenger_prepared", "mult_pippenger", "load_trusted_setup", "load_trusted_setup_file",
                 "free_trusted_setup", "blob_to_kzg_commitment", "compute_kzg_proof", "compute_blob_kzg_proof",
                 "verify_kzg_proof", "verify_blob_kzg_proof", "verify_blob_kzg_proof_batch", "compute_challenge",
                 "bytes_to_kzg_commitment", "bytes_from_bls_field", "compute_cells_and_kzg_proofs",
                 "recover_cells_and_kzg_proofs", "verify_cell_kzg_proof_batch",
                 "compute_verify_cell_kzg_proof_batch_challenge"):
        assert must in names
    out = subprocess.check_output(["nm", "-D", "--defined-only", B.LIB_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    lib = C.CDLL(B.LIB_PATH)
    for n in names:
        getattr(lib, n)


def test_library_is_sm100a_native(B):
    """the shipped kernels are sm_100a SASS (no PTX-JIT fallback for other architectures)"""
    out = subprocess.run(["cuobjdump", "-lelf", B.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:200]
    assert not re.search(r"sm_(7|8|9)\d", out)


def test_no_cpu_fallback(B):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert B.device_count() == 0
    pts = np.zeros((8, 12), np.uint64)
    with pytest.raises(B.B200Error):
        B.PreparedMsm(pts)
    with pytest.raises(B.B200Error, match="no CUDA device"):
        B.mult_pippenger(pts, np.zeros((8, 4), np.uint64))
    with pytest.raises(B.B200Error):
        B.FFTSettings(4)
    with pytest.raises(B.KzgError) as e:
        B.KZGSettings.load_trusted_setup_file()
    assert e.value.code == 2     # C_KZG_ERROR, not a silent CPU path
    with pytest.raises(B.B200Error):
        B.microbench_int()


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under rust-kzg_b200/ may import, link or execute it"""
    pkg = os.path.join(ROOT, "rust-kzg_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|kzg_oracle|libkzg_oracle|c_oracle|\bko_\w+\s*\()")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(txt), os.path.join(dirpath, f)
    # and the shared library does not link it
    import rust_kzg_b200
    out = subprocess.run(["ldd", rust_kzg_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_sha256(B):
    from rust_kzg_b200 import eip4844
    rng = np.random.default_rng(2)
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 1000, 131152):
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert eip4844.sha256(msg) == hashlib.sha256(msg).digest()
        assert eip4844.sha256(msg, portable=True) == hashlib.sha256(msg).digest()


def test_setup_text_rejections_need_no_device(B, tmp_path):
    """wrong counts / truncated / bad hex are BADARGS from the parser (kzg/src/eip_4844.rs:151-228), before any GPU use"""
    toks = open(B.default_trusted_setup_path()).read().split()
    for name, tk in {"g1": ["4095"] + toks[1:], "g2": [toks[0], "64"] + toks[2:], "trunc": toks[:700],
                     "hex": toks[:2] + ["zz" + toks[2][2:]] + toks[3:], "empty": []}.items():
        p = tmp_path / (name + ".txt")
        p.write_text("\n".join(tk) + "\n")
        with pytest.raises(B.KzgError) as e:
            B.KZGSettings.load_trusted_setup_file(str(p))
        assert e.value.code == 1, name
    with pytest.raises(B.KzgError) as e:
        B.KZGSettings.load_trusted_setup(bytes(48 * 4095), bytes(48 * 4096), bytes(96 * 65))
    assert e.value.code == 1
    B.lib().free_trusted_setup(None)


def test_shard_bounds(B):
    for n in (0, 1, 7, 1 << 20, (1 << 24) + 5):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = B.shard_bounds(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
            sizes = [B.shard_bounds(n, r, world)[1] - B.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        B.shard_bounds(10, 2, 2)
