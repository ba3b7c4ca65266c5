"""Pin the CPU oracle (test infrastructure) against every golden vector / KAT the reference holds for the path
(SURVEY.md section 8c).  Runs without a GPU."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, R_MOD


def H(x):
    return bytes.fromhex(x[2:])


def _blob(case, golden_blobs):
    return golden_blobs[case["blob"]] if "blob" in case else bytes(case["blob_len"])


def test_scale2_roots_of_unity(K, kats):
    """blst/src/consts.rs:17-50 == 7^((r-1)/2^i)"""
    from oracle import kzg_oracle as O
    for i, row in enumerate(kats["scale2_root_of_unity"]):
        want = sum(v << (64 * k) for k, v in enumerate(row))
        assert K.fr_to_ints([K.scale2_root_of_unity(i)])[0] == want
        assert O.scale2_root_of_unity(i) == want


def test_fft_and_das_kats(K, kats):
    from oracle import kzg_oracle as O
    exp_fft = [sum(v << (64 * k) for k, v in enumerate(row)) for row in kats["inv_fft_expected"]]
    exp_das = [sum(v << (64 * k) for k, v in enumerate(row)) for row in kats["das_expected_u"]]
    fs = K.FFTSettings(4)
    assert K.fr_to_ints(fs.fft_fr(K.fr_from_ints(range(16)), True)) == exp_fft
    assert K.fr_to_ints(fs.das_fft_extension(K.fr_from_ints(range(8)))) == exp_das
    pfs = O.FFTSettings(4)
    assert pfs.fft_fr(list(range(16)), True) == exp_fft
    assert pfs.das_fft_extension(list(range(8))) == exp_das


def test_expected_powers_and_bytes(K, kats):
    """kzg-bench/src/tests/eip_4844.rs:48-69"""
    x = K.fr_from_ints([32930439])
    acc = K.fr_from_ints([1])
    for row in kats["expected_powers"]:
        assert K.fr_to_ints(acc)[0] == sum(v << (64 * k) for k, v in enumerate(row))
        acc = K.fr_mul(acc, x)
    b = (329).to_bytes(32, "big")
    assert K.fr_to_bytes(K.fr_from_bytes(b)) == b
    with pytest.raises(K.OracleError):
        K.fr_from_bytes(R_MOD.to_bytes(32, "big"))


def test_g1_compressed_kat(K):
    """zkcrypto/bls12_381/src/tests/g1_compressed_valid_test_vectors.dat: compress(i*G), i < 1000"""
    from oracle import kzg_oracle as O
    dat = open(os.path.join(GOLDEN, "g1_compressed_valid_test_vectors.dat"), "rb").read()
    G, acc = K.p1_uncompress(dat[48:96]), K.p1_uncompress(dat[:48])
    pacc, pG = O.INF, O.g1_from_affine(O.G1_GEN)
    for i in range(1000):
        enc = dat[48 * i:48 * i + 48]
        assert K.p1_compress(acc) == enc
        assert K.p1_compress(K.p1_uncompress(enc)) == enc
        if i % 50 == 0:
            assert O.g1_compress(pacc) == enc and K.p1_in_g1(acc)
        acc = K.p1_add(acc, G)
        if i % 50 == 49:
            pacc = O.g1_mul(pG, i + 1)
        elif i % 50 == 0 and i:
            pass


def test_commitment_vectors(K, oracle_settings, vectors, golden_blobs):
    for c in vectors["blob_to_kzg_commitment"]:
        try:
            out = "0x" + K.blob_to_kzg_commitment(_blob(c, golden_blobs), oracle_settings).hex()
        except K.OracleError:
            out = None
        assert out == c["output"], c["name"]


def test_compute_challenge_vectors(K, vectors, golden_blobs):
    for c in vectors["compute_challenge"]:
        assert "0x" + K.compute_challenge(_blob(c, golden_blobs), H(c["commitment"])).hex() == c["output"], c["name"]


def test_compute_kzg_proof_vectors(K, oracle_settings, vectors, golden_blobs):
    for c in vectors["compute_kzg_proof"]:
        try:
            p, y = K.compute_kzg_proof(_blob(c, golden_blobs), H(c["z"]), oracle_settings)
            out = ["0x" + p.hex(), "0x" + y.hex()]
        except (K.OracleError, ValueError):
            out = None
        assert out == c["output"], c["name"]


def test_compute_blob_kzg_proof_vectors(K, oracle_settings, vectors, golden_blobs):
    for c in vectors["compute_blob_kzg_proof"]:
        try:
            out = "0x" + K.compute_blob_kzg_proof(_blob(c, golden_blobs), H(c["commitment"]), oracle_settings).hex()
        except (K.OracleError, ValueError):
            out = None
        assert out == c["output"], c["name"]


def test_compute_cells_vectors(K, oracle_settings, vectors, golden_blobs):
    """the NTT golden vectors: cells = BRP(NTT_8192(INTT_4096(BRP(blob)))) (kzg/src/das.rs:244-275)"""
    for c in vectors["compute_cells"]:
        try:
            cells = K.compute_cells(_blob(c, golden_blobs), oracle_settings)
        except K.OracleError:
            cells = None
        if c["output"] is None:
            assert cells is None, c["name"]
            continue
        assert hashlib.sha256(b"".join(cells)).hexdigest() == c["output"]["all_sha256"], c["name"]
        assert "0x" + cells[0].hex() == c["output"]["cell0"] and "0x" + cells[127].hex() == c["output"]["cell127"]
        assert [hashlib.sha256(x).hexdigest() for x in cells] == c["output"]["cell_sha256"]


def test_compute_cells_and_kzg_proofs_vectors(K, setup_text, vectors, golden_blobs):
    """FK20 (kzg/src/das.rs:660-696): cells and all 128 proofs of every golden case"""
    s = K.KZGSettings(setup_text, nthreads=8)
    for c in vectors["compute_cells_and_kzg_proofs"]:
        try:
            cells, proofs = K.compute_cells_and_kzg_proofs(_blob(c, golden_blobs), s)
        except K.OracleError:
            cells = proofs = None
        if c["output"] is None:
            assert cells is None, c["name"]
            continue
        assert hashlib.sha256(b"".join(cells)).hexdigest() == c["output"]["cells_sha256"], c["name"]
        assert ["0x" + p.hex() for p in proofs] == c["output"]["proofs"], c["name"]


def test_commitment_and_proof_kats(K, oracle_settings, kats):
    k = kats["commitment_kat"]
    assert "0x" + K.blob_to_kzg_commitment(H(k["blob0"]) + bytes(131072 - 32), oracle_settings).hex() == k["commitment"]
    k = kats["proof_kat"]
    p, _ = K.compute_kzg_proof(H(k["blob0"]) + bytes(131072 - 32), H(k["z"]), oracle_settings)
    assert "0x" + p.hex() == k["proof"]


def test_python_oracle_agrees_on_a_vector(K, setup_text, vectors, golden_blobs):
    """the independent big-int restatement reproduces one full commitment + proof vector"""
    from oracle import kzg_oracle as O
    s = O.load_trusted_setup(setup_text)
    c = [x for x in vectors["blob_to_kzg_commitment"] if x["name"].endswith("case_valid_blob_5")][0]
    assert "0x" + O.blob_to_kzg_commitment(golden_blobs[c["blob"]], s).hex() == c["output"]
    c = [x for x in vectors["compute_blob_kzg_proof"] if x["name"].endswith("case_valid_blob_0")][0]
    assert "0x" + O.compute_blob_kzg_proof(golden_blobs[c["blob"]], H(c["commitment"]), s).hex() == c["output"]


def test_msm_variants_agree(K, lagrange_affine, oracle_settings):
    """sequential / parallel Pippenger restatements and the naive sum give one group element; prefix lengths,
    zero scalars and points at infinity (kzg-bench/src/tests/bls12_381.rs:184-387)"""
    rng = np.random.default_rng(1)
    from conftest import rand_ints
    for n in (0, 1, 7, 8, 33, 300):
        sc = K.fr_from_ints(rand_ints(rng, n, R_MOD)) if n else np.zeros((0, 4), np.uint64)
        pts = oracle_settings.g1_lagrange_brp[:n]
        a = K.g1_lincomb(pts, sc, n, nthreads=1)
        b = K.g1_lincomb(pts, sc, n, nthreads=4)
        c = K.msm_naive(pts, sc)
        d = K.msm_affine(lagrange_affine[:n], sc, nthreads=2)
        assert K.p1_compress(a) == K.p1_compress(b) == K.p1_compress(c) == K.p1_compress(d), n
    n = 512
    ints = rand_ints(rng, n, R_MOD)
    pts = oracle_settings.g1_lagrange_brp[:n].copy()
    for i in range(0, n, 10):
        ints[i] = 0
        pts[i + 1] = 0
    sc = K.fr_from_ints(ints)
    assert K.p1_compress(K.g1_lincomb(pts, sc, n, nthreads=3)) == K.p1_compress(K.msm_naive(pts, sc))


def test_fft_properties(K):
    """slow DFT vs fast, roundtrip, stride invariance, DAS zero upper half (kzg-bench/src/tests/fft_fr.rs, das.rs)"""
    from conftest import rand_fr_mont
    rng = np.random.default_rng(3)
    fs = K.FFTSettings(10)
    data = rand_fr_mont(rng, 256)
    fwd = fs.fft_fr(data)
    assert np.array_equal(fwd, fs.fft_fr_slow(data))
    assert np.array_equal(fs.fft_fr(fwd, True), data)
    assert np.array_equal(fwd, K.FFTSettings(8).fft_fr(data))
    assert np.array_equal(fs.fft_fr(data, nthreads=4), fwd)
    for scale in range(1, 10):
        w = 1 << scale
        evens = rand_fr_mont(rng, w // 2)
        odds = fs.das_fft_extension(evens)
        inter = np.empty((w, 4), np.uint64)
        inter[0::2], inter[1::2] = evens, odds
        assert not fs.fft_fr(inter, True)[w // 2:].any()
    with pytest.raises(K.OracleError):
        fs.fft_fr(data[:12])
    with pytest.raises(K.OracleError):
        fs.das_fft_extension(rand_fr_mont(rng, 1024))


# ---- verification vectors: pin the oracle's pairing (kzg-bench/src/tests/eip_4844.rs:676-1010) ------------------
def _blob_any(ref, golden_blobs):
    if "blob" in ref:
        return golden_blobs[ref["blob"]]
    return bytes(ref.get("blob_len", 0))


def _run(K, fn):
    try:
        return fn()
    except (K.OracleError, ValueError):
        return None


def test_g2_generator_is_setup_point_zero(K, oracle_settings):
    g2 = K.g2_monomial(oracle_settings)
    assert np.array_equal(g2[0], K.p2_generator())
    assert all(K.lib.ko_p2_affine_on_curve(g2[i].ctypes.data) for i in range(65))


def test_verify_kzg_proof_vectors(K, oracle_settings, vectors):
    cases = vectors["verify_kzg_proof"]
    assert len(cases) == 122
    seen = {True: 0, False: 0, None: 0}
    for c in cases:
        got = _run(K, lambda: K.verify_kzg_proof(H(c["commitment"]), H(c["z"]), H(c["y"]), H(c["proof"]), oracle_settings))
        assert got == c["output"], c["name"]
        seen[got] += 1
    assert all(seen.values())


def test_verify_blob_kzg_proof_vectors(K, oracle_settings, vectors, golden_blobs):
    cases = vectors["verify_blob_kzg_proof"]
    assert len(cases) == 29
    for c in cases:
        got = _run(K, lambda: K.verify_blob_kzg_proof(_blob_any(c, golden_blobs), H(c["commitment"]), H(c["proof"]),
                                                      oracle_settings))
        assert got == c["output"], c["name"]


def test_verify_blob_kzg_proof_batch_vectors(K, oracle_settings, vectors, golden_blobs):
    cases = vectors["verify_blob_kzg_proof_batch"]
    assert len(cases) == 24
    for c in cases:
        got = _run(K, lambda: K.verify_blob_kzg_proof_batch([_blob_any(b, golden_blobs) for b in c["blobs"]],
                                                            [H(x) for x in c["commitments"]],
                                                            [H(x) for x in c["proofs"]], oracle_settings))
        assert got == c["output"], c["name"]


def test_pairing_bilinearity(K, oracle_settings):
    """e([a]G1, [s]G2) == e([a s^1]... ) is not checkable without s; use e([a]P, Q) == e(P, Q)^a via e([a]P, Q) e(-[a]P, Q) and
    the KZG identity on the setup itself: e(g1_monomial[1], g2[0]) == e(g1_monomial[0], g2[1])  (eip_4844.rs:1005-1020)"""
    g1m, g2 = oracle_settings.g1_monomial, K.g2_monomial(oracle_settings)
    assert K.pairings_verify(g1m[1], g2[0], g1m[0], g2[1])
    assert K.pairings_verify(g1m[5], g2[3], g1m[2], g2[6])
    assert not K.pairings_verify(g1m[5], g2[3], g1m[2], g2[5])
    a = K.fr_from_ints([0x1234567890ABCDEF1234])[0]
    assert K.pairings_verify(K.p1_mult(g1m[1], a), g2[0], K.p1_mult(g1m[0], a), g2[1])


# ---- EIP-7594 recovery / cell verification vectors (kzg-bench/src/tests/eip_7594.rs:190-468) ----------------------
def test_cell_batch_challenge_vectors(K, vectors, golden_cells):
    from conftest import cell_of
    cases = vectors["compute_verify_cell_kzg_proof_batch_challenge"]
    assert len(cases) == 10
    for c in cases:
        got = _run(K, lambda: "0x" + K.compute_verify_cell_kzg_proof_batch_challenge(
            [H(x) for x in c["commitments"]], c["commitment_indices"], c["cell_indices"],
            [cell_of(x, golden_cells) for x in c["cells"]], [H(x) for x in c["proofs"]]).hex())
        assert got == c["output"], c["name"]


def test_verify_cell_kzg_proof_batch_vectors(K, oracle_settings, vectors, golden_cells):
    from conftest import cell_of
    cases = vectors["verify_cell_kzg_proof_batch"]
    assert len(cases) == 32
    seen = {True: 0, False: 0, None: 0}
    for c in cases:
        got = _run(K, lambda: K.verify_cell_kzg_proof_batch([H(x) for x in c["commitments"]], c["cell_indices"],
                                                            [cell_of(x, golden_cells) for x in c["cells"]],
                                                            [H(x) for x in c["proofs"]], oracle_settings))
        assert got == c["output"], c["name"]
        seen[got] += 1
    assert all(seen.values())


def test_recover_cells_and_kzg_proofs_vectors(K, setup_text, vectors, golden_cells):
    from conftest import cell_of
    s = K.KZGSettings(setup_text, nthreads=os.cpu_count() or 1)
    cases = vectors["recover_cells_and_kzg_proofs"]
    assert len(cases) == 18
    for c in cases:
        got = _run(K, lambda: K.recover_cells_and_kzg_proofs(c["cell_indices"], [cell_of(x, golden_cells) for x in c["cells"]], s))
        want = c["output"]
        if want is None:
            assert got is None, c["name"]
        else:
            assert got is not None, c["name"]
            assert got[0] == [cell_of(x, golden_cells) for x in want["cells"]], c["name"]
            assert got[1] == [H(x) for x in want["proofs"]], c["name"]


def test_setup_in_monomial_form_is_rejected(K, setup_text):
    """is_trusted_setup_in_lagrange_form (kzg/src/eip_4844.rs:1005-1020): monomial points in the Lagrange slot -> Err"""
    toks = setup_text.split()
    n1, n2 = int(toks[0]), int(toks[1])
    g2, mono = toks[2 + n1:2 + n1 + n2], toks[2 + n1 + n2:]
    with pytest.raises(K.OracleError):
        K.KZGSettings(" ".join(toks[:2] + mono + g2 + mono))
