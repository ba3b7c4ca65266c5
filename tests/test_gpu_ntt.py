"""Fr NTT / DAS extension through the C ABI vs the oracle and the reference's KATs (bit-exact Montgomery limbs)."""
import numpy as np
import pytest

from conftest import R_MOD, rand_fr_mont, rand_ints

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fs20(B):
    s = B.FFTSettings(20)
    yield s
    s.close()


def test_roots_tables_match_oracle(B, K):
    for scale in (0, 1, 4, 13):
        d, o = B.FFTSettings(scale), K.FFTSettings(scale)
        assert np.array_equal(d.get_roots_of_unity(), o.roots_of_unity)
        assert np.array_equal(d.get_brp_roots_of_unity(), o.brp_roots_of_unity)
        assert np.array_equal(d.get_reversed_roots_of_unity(), o.reverse_roots_of_unity)
        d.close()


def test_inverse_fft_kat(B, K, kats):
    """kzg-bench/src/tests/fft_fr.rs:49-84"""
    fs = B.FFTSettings(4)
    out = fs.fft_fr(K.fr_from_ints(range(16)), True)
    got = [[int(x) for x in K.fr_to_ints([row])[0:1]] for row in out]
    exp = [sum(v << (64 * i) for i, v in enumerate(row)) for row in kats["inv_fft_expected"]]
    assert [g[0] for g in got] == exp
    fs.close()


def test_das_extension_kat(B, K, kats):
    """kzg-bench/src/tests/das.rs:4-31"""
    fs = B.FFTSettings(4)
    odds = fs.das_fft_extension(K.fr_from_ints(range(8)))
    exp = [sum(v << (64 * i) for i, v in enumerate(row)) for row in kats["das_expected_u"]]
    assert K.fr_to_ints(odds) == exp
    fs.close()


@pytest.mark.parametrize("logn", [0, 1, 2, 3, 5, 8, 10, 11, 12, 13, 15, 16, 17, 20])
@pytest.mark.parametrize("inverse", [False, True])
def test_fft_matches_oracle(B, K, fs20, logn, inverse):
    n = 1 << logn
    rng = np.random.default_rng(100 + logn)
    data = rand_fr_mont(rng, n)
    ofs = K.FFTSettings(20)
    exp = ofs.fft_fr(data, inverse, nthreads=8)
    got = fs20.fft_fr(data, inverse)
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("logn", [9, 10, 11, 12, 13, 14])
def test_cluster_kernel_matches_two_passes(B, K, fs20, logn):
    """2^9 .. 2^14 points run both passes in one launch on a thread-block cluster of eight CTAs with the inter-pass transpose
    through distributed shared memory (csrc/ntt.cu, k_ntt_cluster); B200_NTT_CLUSTER=0 keeps the two-launch form.  Same limbs from
    both and from the oracle: forward, inverse, the DAS extension (twisted input) and a batch of five transforms on the device"""
    import os
    import torch
    n = 1 << logn
    rng = np.random.default_rng(400 + logn)
    data = rand_fr_mont(rng, 5 * n)
    ofs = K.FFTSettings(20)
    res = {}
    for mode in ("16", "8", "0"):
        os.environ["B200_NTT_CLUSTER"] = mode
        try:
            fwd = fs20.fft_fr(data[:n], False)
            inv = fs20.fft_fr(data[:n], True)
            das = fs20.das_fft_extension(data[:n])
            d_in = torch.from_numpy(data.view(np.int64)).cuda()
            d_out = torch.zeros_like(d_in)
            fs20.fft_fr_device(d_out.data_ptr(), d_in.data_ptr(), n, True, 5, 0)
            torch.cuda.synchronize()
            res[mode] = (fwd, inv, das, d_out.cpu().numpy().view(np.uint64))
            assert fs20.launches() == (2 if mode == "0" else 1), (mode, fs20.launches())
        finally:
            del os.environ["B200_NTT_CLUSTER"]
    for mode in ("16", "8"):
        for a, b in zip(res[mode], res["0"]):
            assert np.array_equal(a, b), mode
    assert np.array_equal(res["16"][0], ofs.fft_fr(data[:n], False, nthreads=8))
    assert np.array_equal(res["16"][1], ofs.fft_fr(data[:n], True, nthreads=8))
    assert np.array_equal(res["16"][2], ofs.das_fft_extension(data[:n]))
    assert np.array_equal(res["16"][3][4 * n:], ofs.fft_fr(data[4 * n:], True, nthreads=8))


def test_fft_slow_dft_and_roundtrip(B, K):
    """fft_fr vs the O(n^2) DFT at 2^12 and forward/inverse roundtrip (kzg-bench/src/tests/fft_fr.rs:5-46)"""
    fs = B.FFTSettings(12)
    ofs = K.FFTSettings(12)
    data = K.fr_from_ints(range(4096))
    fwd = fs.fft_fr(data, False)
    assert np.array_equal(fwd, ofs.fft_fr_slow(data, False))
    assert np.array_equal(fs.fft_fr(fwd, True), data)
    fs.close()


def test_stride_invariance(B, K):
    """kzg-bench/src/tests/fft_fr.rs:87-106"""
    a, b = B.FFTSettings(9), B.FFTSettings(12)
    data = K.fr_from_ints(range(512))
    assert np.array_equal(a.fft_fr(data), b.fft_fr(data))
    a.close()
    b.close()


def test_bad_lengths_error(B, K, fs20):
    from rust_kzg_b200 import B200Error
    fs = B.FFTSettings(4)
    with pytest.raises(B200Error, match="longer than the available max width"):
        fs.fft_fr(K.fr_from_ints(range(32)))
    with pytest.raises(B200Error, match="power-of-two"):
        fs.fft_fr(K.fr_from_ints(range(12)))
    with pytest.raises(B200Error, match="longer than the available max width"):
        fs.das_fft_extension(K.fr_from_ints(range(16)))
    with pytest.raises(B200Error, match="power-of-two"):
        fs.das_fft_extension(K.fr_from_ints(range(3)))
    fs.close()
    assert B.lib().b200_fft_settings_new(32) is None


@pytest.mark.parametrize("scale", list(range(1, 16)) + [20])
def test_das_extension_random(B, K, fs20, scale):
    """odds from evens: matches the oracle's recursion, and the upper half of the IFFT of the interleaved vector is
    zero (kzg-bench/src/tests/das.rs:35-68)"""
    width = 1 << scale
    rng = np.random.default_rng(scale)
    evens = rand_fr_mont(rng, width // 2)
    odds = fs20.das_fft_extension(evens)
    if scale <= 16:
        assert np.array_equal(odds, K.FFTSettings(20).das_fft_extension(evens))
    data = np.empty((width, 4), np.uint64)
    data[0::2] = evens
    data[1::2] = odds
    coeffs = fs20.fft_fr(data, True)
    assert not coeffs[width // 2:].any()


@pytest.mark.parametrize("logn", [0, 1, 3, 6, 8])
@pytest.mark.parametrize("inverse", [False, True])
def test_fft_g1_matches_oracle(B, K, oracle_settings, logn, inverse):
    """FFTG1::fft_g1 vs the oracle's restatement of fft_g1_fast (kzg-bench/src/tests/fft_g1.rs); compressed bytes"""
    n = 1 << logn
    fs, ofs = B.FFTSettings(10), K.FFTSettings(10)
    pts = oracle_settings.g1_lagrange_brp[100:100 + n].copy()
    if n >= 8:
        pts[3] = 0                          # a point at infinity
        pts[5] = pts[4]                     # a repeated point
    got = fs.fft_g1(pts, inverse)
    exp = ofs.fft_g1(pts, inverse)
    for i in range(n):
        assert K.p1_compress(got[i]) == K.p1_compress(exp[i]), i
    fs.close()


def test_fft_g1_roundtrip_and_slow(B, K, oracle_settings):
    fs, ofs = B.FFTSettings(6), K.FFTSettings(6)
    pts = oracle_settings.g1_monomial[:64].copy()
    fwd = fs.fft_g1(pts, False)
    slow = ofs.fft_g1_slow(pts, False)
    back = fs.fft_g1(fwd, True)
    for i in range(64):
        assert K.p1_compress(fwd[i]) == K.p1_compress(slow[i]), i
        assert K.p1_compress(back[i]) == K.p1_compress(pts[i]), i
    from rust_kzg_b200 import B200Error
    with pytest.raises(B200Error, match="power-of-two"):
        fs.fft_g1(pts[:12])
    with pytest.raises(B200Error, match="longer than the available max width"):
        fs.fft_g1(np.tile(pts, (2, 1)))
    fs.close()


@pytest.mark.parametrize("logn,inverse", [(10, False), (10, True), (12, False)])
def test_fft_g1_matches_oracle_mid_sizes(B, K, oracle_settings, logn, inverse):
    """fft_g1 at 2^10 / 2^12 (several butterfly stages per lane quad, every twiddle class) vs the oracle's fft_g1_fast"""
    n = 1 << logn
    fs, ofs = B.FFTSettings(12), K.FFTSettings(12)
    pts = np.ascontiguousarray(np.tile(oracle_settings.g1_monomial, (max(1, n // 4096), 1))[:n]).copy()
    pts[7] = 0                              # a point at infinity
    pts[9] = pts[8]                         # a repeated point
    got = fs.fft_g1(pts, inverse)
    exp = ofs.fft_g1(pts, inverse)
    assert np.array_equal(K.p1s_to_affine(got), K.p1s_to_affine(exp))
    fs.close()


def test_fft_g1_roundtrip_2p15(B, K, oracle_settings):
    """the reference's bench size (kzg-bench/src/benches/fft.rs: scale 15): inverse(forward(x)) == x on all 2^15 points,
    and three outputs against the slow-DFT identity out[k] = sum_j w^(jk) P_j (a 2^15-term oracle MSM each)"""
    n = 1 << 15
    fs, ofs = B.FFTSettings(15), K.FFTSettings(15)
    pts = np.ascontiguousarray(np.tile(oracle_settings.g1_monomial, (n // 4096, 1)))
    fwd = fs.fft_g1(pts, False)
    back = fs.fft_g1(fwd, True)
    aff = K.p1s_to_affine(pts)
    assert np.array_equal(K.p1s_to_affine(back), aff)
    roots = ofs.roots_of_unity
    for k in (1, 4097, n - 1):
        idx = (np.arange(n, dtype=np.int64) * k) % n
        want = K.msm_affine(aff, np.ascontiguousarray(roots[idx]), nthreads=8)
        assert K.p1_compress(fwd[k]) == K.p1_compress(want), k
    fs.close()


def test_fft_fr_above_2p22_three_passes(B, K):
    """the reference accepts scales up to 31 (blst/src/types/fft_settings.rs:28-58); above 2^22 points the device adds a third
    pass (csrc/ntt.cu).  2^23: forward vs the oracle, inverse(forward(x)) == x, and the DAS extension's zero upper half."""
    n = 1 << 23
    fs, ofs = B.FFTSettings(23), K.FFTSettings(23)
    rng = np.random.default_rng(23)
    data = rand_fr_mont(rng, n)
    fwd = fs.fft_fr(data, False)
    assert np.array_equal(fwd, ofs.fft_fr(data, False, nthreads=8))
    assert np.array_equal(fs.fft_fr(fwd, True), data)
    del ofs
    fs.close()


@pytest.mark.parametrize("logn", [2, 3, 5, 6, 7, 8])
def test_fft_g1_fused_and_plain_stages_agree(B, K, oracle_settings, logn):
    """fft_g1 runs its stages fused -- up to six at once for a lone small transform ((4^R - 1) / 3 independent scalar
    multiplications per 2^R points), in triples (21 per eight points) or pairs (five per four points, csrc/fft_g1.cu) -- while
    the launch is small and stage by stage otherwise: every form against the oracle, forward and inverse, for every residue
    of log n"""
    import os
    n = 1 << logn
    ofs = K.FFTSettings(10)
    pts = oracle_settings.g1_monomial[7:7 + n].copy()
    if n >= 8:
        pts[2] = 0
    want = {inv: K.p1s_to_affine(ofs.fft_g1(pts, inv)) for inv in (False, True)}
    # ... each with one quad per scalar multiplication and with two (one GLV half each, B200_FFT_G1_SPLIT)
    for fuse in ("6", "4", "3", "2", "0"):
        for split in ("1", "0"):
            os.environ["B200_FFT_G1_FUSE"] = fuse
            os.environ["B200_FFT_G1_SPLIT"] = split
            try:
                fs = B.FFTSettings(10)
                for inv in (False, True):
                    assert np.array_equal(K.p1s_to_affine(fs.fft_g1(pts, inv)), want[inv]), (fuse, split, inv)
                fs.close()
            finally:
                del os.environ["B200_FFT_G1_FUSE"], os.environ["B200_FFT_G1_SPLIT"]
