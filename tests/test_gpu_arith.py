"""Device field / point arithmetic vs the CPU oracle, one operation at a time (bit-exact, Montgomery limbs)."""
import ctypes as C

import numpy as np
import pytest

from conftest import P_MOD, R_MOD, rand_ints

pytestmark = pytest.mark.gpu


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _field(B, which, op, a, b=None):
    from rust_kzg_b200 import _lib
    out = np.zeros_like(a)
    fn = B.lib().b200_selftest_fp if which == "fp" else B.lib().b200_selftest_fr
    _lib.check(fn(op, _p(out), _p(a), _p(b) if b is not None else None, a.shape[0]))
    return out


EDGE_FP = [0, 1, 2, P_MOD - 1, P_MOD - 2, (P_MOD - 1) // 2, (1 << 380), (1 << 381) - 1 - ((1 << 381) - 1 >= P_MOD) * 0]
EDGE_FR = [0, 1, 2, R_MOD - 1, R_MOD - 2, (R_MOD - 1) // 2, 1 << 254, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]


@pytest.mark.parametrize("which", ["fp", "fr"])
def test_field_ops_match_oracle(B, K, which):
    rng = np.random.default_rng(0x4B5A47)
    mod = P_MOD if which == "fp" else R_MOD
    edges = [v % mod for v in (EDGE_FP if which == "fp" else EDGE_FR)]
    n = 4096
    xs = edges + rand_ints(rng, n - len(edges), mod)
    ys = rand_ints(rng, n - len(edges), mod) + edges
    conv = K.fp_from_ints if which == "fp" else K.fr_from_ints
    a, b = conv(xs), conv(ys)
    omul, oadd, osub = (K.fp_mul, K.fp_add, K.fp_sub) if which == "fp" else (K.fr_mul, K.fr_add, K.fr_sub)
    assert np.array_equal(_field(B, which, 0, a, b), omul(a, b))
    assert np.array_equal(_field(B, which, 1, a, b), oadd(a, b))
    assert np.array_equal(_field(B, which, 2, a, b), osub(a, b))
    zero = np.zeros_like(a)
    assert np.array_equal(_field(B, which, 3, a), osub(zero, a))
    # squares and products against every edge value
    assert np.array_equal(_field(B, which, 0, a, a), omul(a, a))
    # Montgomery conversions: to_mont(canonical limbs) == oracle's Montgomery form, and back
    w = 6 if which == "fp" else 4
    canon = np.array([[(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(w)] for v in xs], dtype=np.uint64)
    assert np.array_equal(_field(B, which, 5, canon), a)
    assert np.array_equal(_field(B, which, 6, a), canon)
    # the three multiplier variants (default unrolled carry chain, +16 compact carry chain, +32 radix-2^28) agree
    for var in (16, 32, 64):   # 64: a*b on the FP64 pipe (DFMA), reduction on IMAD.WIDE
        assert np.array_equal(_field(B, which, var, a, b), omul(a, b))
        assert np.array_equal(_field(B, which, var, a, a), omul(a, a))
        assert np.array_equal(_field(B, which, var + 5, canon), a)
        assert np.array_equal(_field(B, which, var + 6, a), canon)
    # squaring has its own code path in the radix-2^28 multiplier
    assert np.array_equal(_field(B, which, 32 + 8, a), omul(a, a))
    assert np.array_equal(_field(B, which, 8, a), omul(a, a))


@pytest.mark.parametrize("which", ["fp", "fr"])
def test_field_inverse(B, K, which):
    rng = np.random.default_rng(7)
    mod = P_MOD if which == "fp" else R_MOD
    xs = [1, 2, mod - 1] + rand_ints(rng, 253, mod)
    conv = K.fp_from_ints if which == "fp" else K.fr_from_ints
    a = conv(xs)
    oinv = K.fp_inv if which == "fp" else K.fr_inv
    assert np.array_equal(_field(B, which, 4, a), oinv(a))
    assert np.array_equal(_field(B, which, 16 + 4, a), oinv(a))
    assert np.array_equal(_field(B, which, 7, a), oinv(a))          # Fermat cross-check of the bingcd inverse (op 4)
    assert np.array_equal(_field(B, which, 9, a), oinv(a))          # ... and the bit-serial binary Euclid
    z = np.zeros((4, a.shape[1]), np.uint64)
    assert not _field(B, which, 4, z).any()                         # inverse(0) == 0


def _rand_points(K, rng, n):
    """n random multiples of the generator (Jacobian, non-trivial Z after adds)"""
    dat = open(__import__("os").path.join(__import__("conftest").GOLDEN, "g1_compressed_valid_test_vectors.dat"), "rb").read()
    G = K.p1_uncompress(dat[48:96])
    pts = []
    for s in K.fr_from_ints(rand_ints(rng, n, R_MOD)):
        pts.append(K.p1_mult(G, s))
    return np.array(pts)


@pytest.mark.parametrize("mixed", [0, 1])
def test_point_add_all_cases(B, K, mixed):
    from rust_kzg_b200 import _lib
    rng = np.random.default_rng(11)
    n = 64
    a = _rand_points(K, rng, n)
    b = _rand_points(K, rng, n)
    inf = np.zeros(18, np.uint64)
    # exceptional cases of the addition law: P+P, P+(-P), inf+P, P+inf, inf+inf
    b[0] = a[0]                                   # doubling
    neg = a[1].copy()
    negy = K.fp_sub(np.zeros((1, 6), np.uint64), neg[6:12].reshape(1, 6))[0]
    neg[6:12] = negy
    b[1] = neg                                    # cancellation
    a[2] = inf
    b[3] = inf
    a[4] = inf
    b[4] = inf
    # same point, different Jacobian representative: b[5] = 2*a[5] - a[5]
    b[5] = K.p1_add(K.p1_double(a[5]), np.concatenate([a[5][:6], K.fp_sub(np.zeros((1, 6), np.uint64), a[5][6:12].reshape(1, 6))[0], a[5][12:]]))
    out = np.zeros_like(a)
    _lib.check(B.lib().b200_selftest_p1_add(_p(out), _p(a), _p(b), n, mixed))
    for i in range(n):
        exp = K.p1_add(a[i], b[i])
        assert K.p1_compress(out[i]) == K.p1_compress(exp), i
    comp = np.zeros((n, 48), np.uint8)
    _lib.check(B.lib().b200_selftest_p1_compress(_p(comp), _p(out), n))
    for i in range(n):
        assert comp[i].tobytes() == K.p1_compress(out[i]), i


def test_compress_kat(B, K):
    """compress(i*G), i < 1000 (zkcrypto/bls12_381/src/tests/g1_compressed_valid_test_vectors.dat) on the device"""
    import os
    from conftest import GOLDEN
    from rust_kzg_b200 import _lib
    dat = open(os.path.join(GOLDEN, "g1_compressed_valid_test_vectors.dat"), "rb").read()
    G = K.p1_uncompress(dat[48:96])
    acc = K.p1_uncompress(dat[:48])
    pts = []
    for _ in range(1000):
        pts.append(acc)
        acc = K.p1_add(acc, G)
    pts = np.array(pts)
    comp = np.zeros((1000, 48), np.uint8)
    _lib.check(B.lib().b200_selftest_p1_compress(_p(comp), _p(pts), 1000))
    assert comp.tobytes() == dat


def test_quad_glv_scalar_multiplication_edges(B, K, lagrange_affine):
    """the lane-quad scalar multiplication splits k = k1 + k2 z^2 (GLV): scalars at and around the split boundaries, the ends
    of the range, points at infinity and repeated points, against the oracle's naive MSM"""
    from conftest import R_MOD, rand_ints
    z2 = 0xd201000000010000 ** 2
    edge = [0, 1, 2, z2 - 1, z2, z2 + 1, 2 * z2, 7 * z2 - 3, (1 << 128) - 1, 1 << 128, (1 << 128) + 1, (z2 - 1) * z2, z2 * z2 - z2,
            R_MOD - 1, R_MOD - 2, R_MOD - z2, R_MOD - z2 - 1, R_MOD - z2 + 1, (R_MOD - 1) // 2, 0xF, 0xF0, 0xFFFFFFFF << 124]
    rng = np.random.default_rng(41)
    for n, ints in ((len(edge), edge), (1, [R_MOD - 1]), (9, rand_ints(rng, 9, R_MOD)), (64, rand_ints(rng, 64, R_MOD)),
                    (203, rand_ints(rng, 203, R_MOD))):
        pts = lagrange_affine[:n].copy()
        if n > 8:
            pts[3] = 0                      # infinity
            pts[5] = pts[4]                 # P twice: the tree sum meets P + P
            ints = list(ints)
            ints[5] = ints[4]
        sc = K.fr_from_ints([v % R_MOD for v in ints])
        got = B.selftest_lincomb_quads(pts, sc)
        want = K.msm_affine(pts, sc, nthreads=2)
        assert K.p1_compress(got) == K.p1_compress(want), n
