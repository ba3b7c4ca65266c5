"""MSM through the C ABI (prepare_msm / mult_pippenger_prepared / mult_pippenger) vs the CPU oracle.
Parity is on the 48-byte compressed encoding, the reference's own definition (fuzz/src/lib.rs:81-95)."""
import numpy as np
import pytest

from conftest import R_MOD, rand_fr_mont, rand_ints

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _bucket_pipeline_only():
    """4096-point handles would take the direct-lookup table for full-length calls (csrc/capi_msm.cu); the tests of this module
    pin the bucket pipeline and its window plans, so they switch it off -- test_direct_lookup_handles covers the table path"""
    import os
    os.environ["B200_MSM_DIRECT"] = "0"
    yield
    del os.environ["B200_MSM_DIRECT"]


@pytest.fixture(scope="module")
def prepared(B, lagrange_affine):
    h = B.PreparedMsm(lagrange_affine)
    yield h
    h.close()


def _same(K, got, exp):
    assert K.p1_compress(got) == K.p1_compress(exp)


@pytest.mark.parametrize("n", [4096, 4095, 1000, 129, 64, 33, 8, 7, 2, 1])
def test_prepared_matches_oracle(B, K, prepared, lagrange_affine, n):
    rng = np.random.default_rng(n)
    sc = K.fr_from_ints(rand_ints(rng, n, R_MOD))
    got = prepared.mult(sc)
    exp = K.msm_affine(lagrange_affine[:n], sc)
    _same(K, got, exp)


def test_variable_base_matches_oracle(B, K, lagrange_affine):
    rng = np.random.default_rng(5)
    for n in (4096, 300, 31, 8, 3, 1):
        sc = K.fr_from_ints(rand_ints(rng, n, R_MOD))
        got = B.mult_pippenger(lagrange_affine[:n], sc)
        _same(K, got, K.msm_affine(lagrange_affine[:n], sc))


def test_edge_scalars(B, K, prepared, lagrange_affine):
    """all-zero, all-one, r-1, small scalars, 10% zeros (kzg-bench/src/tests/bls12_381.rs:282-296)"""
    n = 4096
    rng = np.random.default_rng(9)
    cases = {
        "zeros": [0] * n,
        "ones": [1] * n,
        "r-1": [R_MOD - 1] * n,
        "small": [int(x) for x in rng.integers(0, 1 << 16, n)],
        "same": [0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % R_MOD] * n,
        "window-edges": [(1 << (16 * (i % 16))) * 0x8000 % R_MOD for i in range(n)],
    }
    sparse = rand_ints(rng, n, R_MOD)
    for i in range(n):
        if rng.random() < 0.1:
            sparse[i] = 0
    cases["10pct-zero"] = sparse
    for name, ints in cases.items():
        sc = K.fr_from_ints(ints)
        exp = K.msm_affine(lagrange_affine, sc)
        assert K.p1_compress(prepared.mult(sc)) == K.p1_compress(exp), name
        assert K.p1_compress(B.mult_pippenger(lagrange_affine, sc)) == K.p1_compress(exp), name + "/variable"


def test_infinity_and_repeated_points(B, K, lagrange_affine):
    """10% points at infinity; repeated points; P and -P in one call (SURVEY.md section 7 hard parts)"""
    n = 2048
    rng = np.random.default_rng(21)
    pts = lagrange_affine[:n].copy()
    for i in range(n):
        if rng.random() < 0.1:
            pts[i] = 0                      # infinity = all-zero affine
    pts[100:200] = pts[300]                 # repeats
    negp = pts[301].copy()
    negp[6:12] = K.fp_sub(np.zeros((1, 6), np.uint64), negp[6:12].reshape(1, 6))[0]
    pts[400:450] = negp
    pts[450:500] = pts[301]
    sc_i = rand_ints(rng, n, R_MOD)
    for i in range(400, 500):
        sc_i[i] = sc_i[400]                 # P and -P with equal scalars land in the same bucket and cancel
    sc = K.fr_from_ints(sc_i)
    exp = K.msm_affine(pts, sc)
    _same(K, B.mult_pippenger(pts, sc), exp)
    h = B.PreparedMsm(pts)
    _same(K, h.mult(sc), exp)
    h.close()


def test_generator_sum_closed_form(B, K):
    """sum (i+1)*G, i < 255 == (255*256/2)*G (kzg-bench/src/tests/bls12_381.rs:184-221)"""
    import os
    from conftest import GOLDEN
    dat = open(os.path.join(GOLDEN, "g1_compressed_valid_test_vectors.dat"), "rb").read()
    G = K.p1_uncompress(dat[48:96])
    n = 255
    pts = K.p1s_to_affine(np.tile(G, (n, 1)))
    sc = K.fr_from_ints(range(1, n + 1))
    exp = K.p1_mult(G, K.fr_from_ints([n * (n + 1) // 2])[0])
    _same(K, B.mult_pippenger(pts, sc), exp)
    h = B.PreparedMsm(pts)
    _same(K, h.mult(sc), exp)
    h.close()


def test_prefix_lengths_reuse_table(B, K, prepared, lagrange_affine):
    """every prefix length 0..128 with one prepared table (kzg-bench/src/tests/bls12_381.rs:298-387)"""
    rng = np.random.default_rng(3)
    sc = K.fr_from_ints(rand_ints(rng, 128, R_MOD))
    for n in list(range(0, 20)) + [31, 32, 33, 63, 64, 65, 127, 128]:
        if n == 0:
            got = np.zeros(18, np.uint64)
            from rust_kzg_b200 import _lib
            import ctypes as C
            _lib.check(B.lib().mult_pippenger_prepared(prepared.h, got.ctypes.data_as(C.c_void_p), 0, None))
            assert K.p1_is_inf(got)
            continue
        _same(K, prepared.mult(sc[:n]), K.msm_affine(lagrange_affine[:n], sc[:n]))


def test_batch_matches_single(B, K, lagrange_affine):
    import os
    os.environ["B200_MSM_MAX_BATCH"] = "8"
    try:
        h = B.PreparedMsm(lagrange_affine)
    finally:
        del os.environ["B200_MSM_MAX_BATCH"]
    rng = np.random.default_rng(17)
    sc = rand_fr_mont(rng, 8 * 4096)
    outs = h.mult_batch(sc, 8)
    for b in range(8):
        _same(K, outs[b], K.msm_affine(lagrange_affine, sc[b * 4096:(b + 1) * 4096], nthreads=8))
    h.close()


def test_large_tiled_bases_folded_oracle(B, K, lagrange_affine):
    """2^16 terms over tiled bases P_i = L[i mod 4096]: exact cheap oracle by folding scalars."""
    n = 1 << 16
    rng = np.random.default_rng(31)
    sc = rand_fr_mont(rng, n)
    pts = np.tile(lagrange_affine, (n // 4096, 1))
    folded = sc[:4096].copy()
    for k in range(1, n // 4096):
        folded = K.fr_add(folded, sc[k * 4096:(k + 1) * 4096])
    exp = K.msm_affine(lagrange_affine, folded, nthreads=8)
    h = B.PreparedMsm(pts)
    _same(K, h.mult(sc), exp)
    h.close()
    _same(K, B.mult_pippenger(pts, sc), exp)


def test_adversarial_distributions_at_scale(B, K, lagrange_affine):
    """2^16 terms whose digits all collide: every term lands in the same few buckets, so a bucket holds up to 2^16
    entries and is cut into ~1000 tasks + a warp-parallel combine (bucket load imbalance)."""
    n = 1 << 16
    reps = n // 4096
    pts = np.tile(lagrange_affine, (reps, 1))
    h = B.PreparedMsm(pts)
    rng = np.random.default_rng(41)
    s0, s1 = rand_ints(rng, 2, R_MOD)
    cases = {
        "all-equal": [s0] * n,
        "two-values": [s0 if (i * 7) % 3 else s1 for i in range(n)],
        "all-r-minus-1": [R_MOD - 1] * n,
        "one-hot-window": [1 << 200] * n,
    }
    for name, ints in cases.items():
        sc = K.fr_from_ints(ints[:4096 * 2])          # conversion is slow in Python: build two tiles, then repeat
        sc = np.tile(sc, (reps // 2, 1))
        ints_eff = (ints[:4096 * 2]) * (reps // 2)
        folded = [sum(ints_eff[j + 4096 * k] for k in range(reps)) % R_MOD for j in range(4096)]
        exp = K.msm_affine(lagrange_affine, K.fr_from_ints(folded), nthreads=8)
        assert K.p1_compress(h.mult(sc)) == K.p1_compress(exp), name
        assert K.p1_compress(B.mult_pippenger(pts, sc)) == K.p1_compress(exp), name + "/variable"
    h.close()


def _with_env(name, value, fn):
    import os
    old = os.environ.get(name)
    os.environ[name] = str(value)
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[name]
        else:
            os.environ[name] = old


@pytest.mark.parametrize("c", [17, 18, 20, 22])
def test_wide_windows_segment_fold(B, K, lagrange_affine, c):
    """windows wider than 16 bits: the bucket set is folded by segments of 2^(c-16) buckets before the 15-bit reduce
    (k_segment_fold); same group element as the reference's p1_integrate_buckets (tiling_pippenger_ops.rs:21-45)."""
    n = 1 << 14
    rng = np.random.default_rng(100 + c)
    sc = rand_fr_mont(rng, n)
    pts = np.tile(lagrange_affine, (n // 4096, 1))
    folded = sc[:4096].copy()
    for k in range(1, n // 4096):
        folded = K.fr_add(folded, sc[k * 4096:(k + 1) * 4096])
    exp = K.msm_affine(lagrange_affine, folded, nthreads=8)
    h = _with_env("B200_MSM_C", c, lambda: B.PreparedMsm(pts))
    assert h.info()["c"] == c
    _same(K, h.mult(sc), exp)
    # edge scalars: every term in the top bucket of a segment / bucket 0 / cancelling signs
    for name, ints in {"r-1": [R_MOD - 1] * 4096, "ones": [1] * 4096, "zeros": [0] * 4096,
                       "half-window": [(1 << (c - 1)) % R_MOD] * 4096,
                       "small": [int(x) for x in rng.integers(0, 1 << 20, 4096)]}.items():
        s4 = K.fr_from_ints(ints)
        assert K.p1_compress(h.mult(s4)) == K.p1_compress(K.msm_affine(lagrange_affine, s4, nthreads=8)), name
    h.close()
    # variable base through the same reduce (one bucket set per window + Horner)
    if c == 20:
        # engines of the variable-base call are cached per capacity: 2^15 terms is a capacity no other test uses, so the
        # override is what this engine is built with
        pts2, sc2 = np.concatenate([pts, pts]), np.concatenate([sc, sc])
        got = _with_env("B200_MSM_VC", c, lambda: B.mult_pippenger(pts2, sc2))
        _same(K, got, K.p1_add(exp, exp))


@pytest.mark.parametrize("c,c0", [(13, 9), (15, 1), (20, 16), (12, 4)])
def test_narrow_first_window(B, K, lagrange_affine, c, c0):
    """window 0 of c0 bits, the others of c bits, so that the top window ends exactly at bit 256 (MsmConfig::c0): same
    digits-times-table-rows identity as uniform windows, checked on random, blob-like (< 2^248) and edge scalars."""
    n = 4096
    rng = np.random.default_rng(200 + c)
    h = _with_env("B200_MSM_C", c, lambda: _with_env("B200_MSM_C0", c0, lambda: B.PreparedMsm(lagrange_affine)))
    assert h.info()["c"] == c
    cases = {"random": rand_ints(rng, n, R_MOD), "blob-like": rand_ints(rng, n, 1 << 248), "r-1": [R_MOD - 1] * n,
             "ones": [1] * n, "low-window-edge": [((1 << (c0 - 1)) + (1 << 255) % R_MOD) % R_MOD] * n if c0 > 1 else [3] * n,
             "zeros": [0] * n}
    for name, ints in cases.items():
        sc = K.fr_from_ints(ints)
        assert K.p1_compress(h.mult(sc)) == K.p1_compress(K.msm_affine(lagrange_affine, sc, nthreads=8)), name
    h.close()


@pytest.mark.parametrize("logn", [20, 21])
def test_baseline_config_default_plan_folded_oracle(B, K, lagrange_affine, logn):
    """BASELINE configs[1] itself (2^20 terms, the auto-selected window plan with the segment fold) and the per-GPU share
    of configs[4] (2^21): host-pointer call, device-pointer call and an adversarial vector, against the folded-scalar
    oracle (tiled bases P_i = L[i mod 4096])."""
    import torch
    n = 1 << logn
    rng = np.random.default_rng(300 + logn)
    sc = rand_fr_mont(rng, n)
    pts = np.tile(lagrange_affine, (n // 4096, 1))
    cur = sc
    while cur.shape[0] > 4096:
        half = cur.shape[0] // 2
        cur = K.fr_add(cur[:half], cur[half:])
    exp = K.p1_compress(K.msm_affine(lagrange_affine, np.ascontiguousarray(cur), nthreads=8))
    h = B.PreparedMsm(pts)
    c = B.msm_plan(n, True)["c"]
    assert h.info()["c"] == c and c >= 17          # the default plan at this size is a wide window: the fold is on the path
    assert K.p1_compress(h.mult(sc)) == exp
    d_sc = torch.from_numpy(sc.view(np.int64)).cuda()
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    h.mult_device(d_out.data_ptr(), n, d_sc.data_ptr(), 1, 0)
    torch.cuda.synchronize()
    assert K.p1_compress(d_out.cpu().numpy().view(np.uint64)) == exp
    st = h.last_stats()
    assert st["entries"] > 0 and st["adds"]["accumulate"] == st["entries"] - st["tasks"] and st["fold_bits"] == c - 16
    if logn == 20:
        # every scalar equal: all 2^20 entries of a window in ONE bucket (the all-0x02 consensus blob at MSM scale)
        same = np.tile(K.fr_from_ints([0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % R_MOD]), (n, 1))
        tot = K.fr_from_ints([(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % R_MOD) * (n // 4096) % R_MOD] * 4096)
        assert K.p1_compress(h.mult(same)) == K.p1_compress(K.msm_affine(lagrange_affine, tot, nthreads=8))
    h.close()


def test_batch_affine_accumulation_option(B, K, lagrange_affine):
    """k_accumulate_affine (B200_MSM_AFFINE=1: chunked lock-step affine additions with a shared inversion; off by default,
    see profiles/r02_affine.md) against the folded-scalar oracle, on inputs that hit every exceptional case of the affine
    addition law: P + P (the same scalar on every copy of a point), P + (-P) (s and r - s), points at infinity in the
    table, all-equal scalars (one bucket per window), zeros."""
    n = 1 << 17                                   # smallest table whose calls take the affine path
    reps = n // 4096
    rng = np.random.default_rng(77)
    pts = np.tile(lagrange_affine, (reps, 1))
    pts_inf = pts.copy()
    pts_inf[5::97] = 0
    base = rand_fr_mont(rng, 4096)
    neg = K.fr_sub(np.zeros((4096, 4), np.uint64), base)
    cases = {
        "uniform": rand_fr_mont(rng, n),
        "same-scalar-per-point": np.tile(base, (reps, 1)),
        "plus-minus": np.concatenate([np.tile(base, (reps // 2, 1)), np.tile(neg, (reps // 2, 1))]),
        "all-equal": np.tile(base[:1], (n, 1)),
        "zeros": np.zeros((n, 4), np.uint64),
    }

    def fold(sc):
        cur = sc
        while cur.shape[0] > 4096:
            half = cur.shape[0] // 2
            cur = K.fr_add(cur[:half], cur[half:])
        return np.ascontiguousarray(cur)

    for tree in (0, 1):
        for p, tag in ((pts, ""), (pts_inf, "+inf")):
            h = _with_env("B200_MSM_AFFINE", 1, lambda: _with_env("B200_MSM_AFFINE_TREE", tree, lambda: B.PreparedMsm(p)))
            for name, sc in cases.items():
                sc = np.ascontiguousarray(sc)
                eff = sc.copy()
                if tag:
                    eff[5::97] = 0
                got = h.mult(sc)
                assert h.info()["accumulate"] == "affine"
                assert K.p1_compress(got) == K.p1_compress(K.msm_affine(lagrange_affine, fold(eff), nthreads=8)), (tree, tag, name)
            h.close()


def test_scalar_randomisation_and_its_subgroup_guard(B, K, lagrange_affine):
    """Prepared tables multiply every scalar by a per-base rho_i over a table of rho_i^-1 * P_i (csrc/msm.cu): the same
    group element for bases of the prime-order subgroup -- checked against the oracle with randomisation on and off -- and
    NOT applied when a base lies outside the subgroup (the identity rho * rho^-1 = 1 holds mod r only), where the plain
    table must still give the exact integer combination the oracle computes."""
    rng = np.random.default_rng(91)
    n = 4096
    sc = K.fr_from_ints(rand_ints(rng, n, R_MOD))
    want = K.p1_compress(K.msm_affine(lagrange_affine, sc, nthreads=8))
    on = B.PreparedMsm(lagrange_affine)
    assert on.info()["randomized"] is True
    off = _with_env("B200_MSM_RANDOMIZE", 0, lambda: B.PreparedMsm(lagrange_affine))
    assert off.info()["randomized"] is False
    assert K.p1_compress(on.mult(sc)) == want and K.p1_compress(off.mult(sc)) == want
    on.close()
    off.close()
    # a curve point outside G1: decode random abscissas until one is on the curve (blst_p1_uncompress does not test the
    # subgroup); almost every curve point has a cofactor component
    outside = None
    for t in range(200):
        x = int(rng.integers(1, 1 << 62)) * 0x10001 + t
        b = bytearray(x.to_bytes(48, "big"))
        b[0] |= 0x80
        try:
            cand = K.p1_uncompress_affine(bytes(b))
        except Exception:
            continue
        if not K.p1_in_g1(K.p1_from_affine(cand)):
            outside = cand
            break
    assert outside is not None
    pts = lagrange_affine[:256].copy()
    pts[17] = outside
    h = B.PreparedMsm(pts)
    assert h.info()["randomized"] is False                      # the guard: one base outside the subgroup turns it off
    s256 = K.fr_from_ints(rand_ints(rng, 256, R_MOD))
    assert K.p1_compress(h.mult(s256)) == K.p1_compress(K.msm_affine(pts, s256, nthreads=4))
    h.close()


@pytest.mark.parametrize("bits", [11, 8, 13])
def test_direct_lookup_handles(B, K, lagrange_affine, bits):
    """fixed-base handles of 4096 points sum full-length calls by direct lookups in a table of every signed-digit multiple
    (csrc/fk20_direct.cu, one launch, Jacobian out): random, blob-like and edge scalars, bases at infinity and repeated, the
    batch entry, coalesced callers, and prefixes (which stay on the bucket pipeline) against the oracle"""
    import os
    import threading
    env = {"B200_MSM_DIRECT": "1", "B200_MSM_DIRECT_BITS": str(bits), "B200_DIRECT_RESERVE_GB": "2"}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        pts = lagrange_affine.copy()
        pts[5] = 0                                   # the point at infinity (all-zero affine)
        pts[9] = pts[8]                              # a repeated base
        h = B.PreparedMsm(pts)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    assert h.info()["direct_bits"] == bits
    rng = np.random.default_rng(300 + bits)
    n = 4096
    half = 1 << (bits - 1)
    cases = {"random": rand_ints(rng, n, R_MOD), "blob-like": rand_ints(rng, n, 1 << 248), "r-1": [R_MOD - 1] * n, "ones": [1] * n,
             "zeros": [0] * n, "half-window": [half] * n, "half-window+1": [half + 1] * n,
             "window-borrow": [((half << bits) | half) % R_MOD] * n, "all-ones-bits": [(1 << 254) - 1] * n}
    for name, ints in cases.items():
        sc = K.fr_from_ints(ints)
        assert K.p1_compress(h.mult(sc)) == K.p1_compress(K.msm_affine(pts, sc, nthreads=8)), name
    # prefix: the bucket pipeline on the first 1000 bases
    sc = K.fr_from_ints(rand_ints(rng, 1000, R_MOD))
    assert K.p1_compress(h.mult(sc)) == K.p1_compress(K.msm_affine(pts[:1000], sc, nthreads=8))
    # batch of 5 vectors in one launch; CTA ranges cross the vector boundaries
    sc5 = K.fr_from_ints(rand_ints(rng, 5 * n, R_MOD))
    got = h.mult_batch(sc5, 5)
    for v in range(5):
        assert K.p1_compress(got[v]) == K.p1_compress(K.msm_affine(pts, sc5[v * n:(v + 1) * n], nthreads=8)), v
    # coalesced callers
    want = [K.p1_compress(got[v]) for v in range(5)]
    errs = []

    def worker(k):
        try:
            for rep in range(3):
                v = (k + rep) % 5
                assert K.p1_compress(h.mult(sc5[v * n:(v + 1) * n])) == want[v]
        except Exception as e:                      # noqa: BLE001
            errs.append(repr(e))
    th = [threading.Thread(target=worker, args=(k,)) for k in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    h.close()
