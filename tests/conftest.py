import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SETUP_PATH = os.path.join(ROOT, "rust-kzg_b200", "data", "trusted_setup.txt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def K():
    """the C oracle binding (test infrastructure)"""
    from oracle import c_oracle
    return c_oracle


@pytest.fixture(scope="session")
def setup_text():
    with open(SETUP_PATH) as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle_settings(K, setup_text):
    return K.KZGSettings(setup_text)


@pytest.fixture(scope="session")
def lagrange_affine(K, oracle_settings):
    """the 4096 bit-reversed Lagrange bases as blst_p1_affine rows (n,12) u64"""
    return K.p1s_to_affine(oracle_settings.g1_lagrange_brp)


@pytest.fixture(scope="session")
def vectors():
    with open(os.path.join(GOLDEN, "vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_blobs():
    with open(os.path.join(GOLDEN, "blobs.bin"), "rb") as f:
        data = f.read()
    return [data[i:i + 131072] for i in range(0, len(data), 131072)]


@pytest.fixture(scope="session")
def golden_cells():
    """pool of the well-formed 2048-byte cells the EIP-7594 vectors refer to by index"""
    with open(os.path.join(GOLDEN, "cells.bin"), "rb") as f:
        data = f.read()
    return [data[i:i + 2048] for i in range(0, len(data), 2048)]


def cell_of(ref, pool):
    """a cell reference of vectors.json: pool index, or {"hex": ...} for malformed ones (may raise ValueError)"""
    return pool[ref] if isinstance(ref, int) else bytes.fromhex(ref["hex"][2:])


@pytest.fixture(scope="session")
def B():
    """the product package; GPU tests only"""
    import rust_kzg_b200
    return rust_kzg_b200


R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
P_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB


def rand_ints(rng, n, mod):
    """n uniform integers in [0, mod) from a numpy Generator (wide draw, then reduce)"""
    nbytes = (mod.bit_length() + 7) // 8 + 8
    raw = rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)
    return [int.from_bytes(raw[i].tobytes(), "little") % mod for i in range(n)]


def rand_fr_mont(rng, n):
    """n uniform Fr in Montgomery limbs, (n,4) u64 -- any 4x64-bit pattern below r is a valid Montgomery value of a
    uniform element, so draw canonical-looking limbs directly (fast path for large n)"""
    out = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)   # < 2^254 < r
    return out
