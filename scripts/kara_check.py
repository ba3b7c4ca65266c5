import sys, os, json, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import rust_kzg_b200 as B
from oracle import c_oracle as K
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
from conftest import rand_ints, R_MOD, P_MOD
import ctypes as C
L = B.lib()
def field(which, op, a, b=None):
    out = np.zeros_like(a)
    fn = L.b200_selftest_fp if which == "fp" else L.b200_selftest_fr
    from rust_kzg_b200 import _lib
    _lib.check(fn(op, out.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p) if b is not None else None, a.shape[0]))
    return out
rng = np.random.default_rng(3)
res = {}
for which, mod, conv, omul in (("fp", P_MOD, K.fp_from_ints, K.fp_mul), ("fr", R_MOD, K.fr_from_ints, K.fr_mul)):
    xs = rand_ints(rng, 4000, mod) + [0, 1, mod - 1, mod - 2, 2, (mod - 1) // 2] * 4
    ys = rand_ints(rng, 4000, mod) + [0, mod - 1, mod - 1, 1, mod - 2, (mod + 1) // 2] * 4
    a, b = conv(xs), conv(ys)
    res[which + "_mul_ok"] = bool(np.array_equal(field(which, 128 + 0, a, b), omul(a, b)))
    res[which + "_sqr_ok"] = bool(np.array_equal(field(which, 128 + 8, a), omul(a, a)))
    # raw limb patterns (Montgomery residues are just numbers < mod): all-ones halves exercise the carry bits of a0 + a1
    w = a.shape[1]
    edge = np.zeros((6, w), np.uint64)
    edge[0] = conv([mod - 1])[0]; edge[1, : w // 2] = np.uint64(0xFFFFFFFFFFFFFFFF); edge[2, w // 2 - 1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    edge[3, : w // 2] = np.uint64(0xFFFFFFFFFFFFFFFF); edge[3, w // 2:] = np.uint64(0x0FFFFFFFFFFFFFFF) if which == "fr" else np.uint64(0x00FFFFFFFFFFFFFF)
    edge[3, w - 1] = np.uint64(0x1000000000000000) if which == "fr" else np.uint64(0x0100000000000000)
    edge[4, 0] = 1; edge[5, w // 2] = 1
    ee = np.ascontiguousarray(np.repeat(edge, 6, axis=0)); ff = np.ascontiguousarray(np.tile(edge, (6, 1)))
    res[which + "_edge_ok"] = bool(np.array_equal(field(which, 128 + 0, ee, ff), omul(ee, ff)))
for f, name in ((0, "fp"), (1, "fr")):
    for m, mn in ((0, "cios"), (5, "kara")):
        res["%s_%s_Gmul_s" % (name, mn)] = B.microbench_mul(f, m) / 1e9
print(json.dumps(res, indent=1))
