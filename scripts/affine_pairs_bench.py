#!/usr/bin/env python3
"""b200_bench_affine_pairs: rate of independent batch-affine pair additions (one tree round without bucket logic), against
the XYZZ task kernel's 2.65 G additions/s.  python scripts/affine_pairs_bench.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_kzg_b200 as B
from rust_kzg_b200 import _lib

L = B.lib()
out = {}
for npairs in (6_800_000, 1_700_000, 400_000):
    for K in (8, 16, 32, 64):
        ms = C.c_double()
        _lib.check(L.b200_bench_affine_pairs(13 << 20, npairs, K, C.byref(ms)))
        out["pairs=%d,K=%d" % (npairs, K)] = {"ms": ms.value, "G_adds_per_s": npairs / ms.value / 1e6}
print(json.dumps(out, indent=1))
