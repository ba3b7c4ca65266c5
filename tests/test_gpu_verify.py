"""Verification side of the c-kzg-4844 ABI on the device (pairing included) vs the reference's consensus-spec
vectors (kzg-bench/src/test_vectors/verify_*, runners kzg-bench/src/tests/eip_4844.rs:676-1010) and vs the oracle on
seeded random inputs."""
import numpy as np
import pytest

from conftest import R_MOD, rand_ints

pytestmark = pytest.mark.gpu


def H(x):
    return bytes.fromhex(x[2:])


@pytest.fixture(scope="module")
def ts(B):
    s = B.KZGSettings.load_trusted_setup_file()
    yield s
    s.free()


def _blob_any(ref, golden_blobs):
    if "blob" in ref:
        return golden_blobs[ref["blob"]]
    return bytes(ref.get("blob_len", 0))


def _run(B, fn):
    try:
        return fn()
    except (B.KzgError, ValueError):
        return None


def test_g2_points_match_oracle(K, ts, oracle_settings):
    """g2_values_monomial of CKZGSettings = blst_p2_from_affine(uncompress(...)) of all 65 points"""
    got = ts.array("g2_values_monomial", 65, 36)
    want = K.g2_monomial(oracle_settings)
    assert np.array_equal(got[:, :24], want)
    one = K.fp_from_ints([1])[0]
    assert np.array_equal(got[:, 24:30], np.tile(one, (65, 1))) and not got[:, 30:].any()


def test_pairing_kernel_vs_oracle(K, ts, oracle_settings):
    """the device Miller loop + final exponentiation on setup identities and their perturbations"""
    g1m = oracle_settings.g1_monomial
    g2 = K.g2_monomial(oracle_settings)
    q = {0: g2[0], 1: g2[1], 2: g2[64]}
    # e([s^k]G1, G2) == e([s^(k-1)]G1, [s]G2), e([s^64+k]G1, G2) == e([s^k] G1, [s^64]G2)
    cases = [(g1m[1], 0, g1m[0], 1), (g1m[7], 0, g1m[6], 1), (g1m[64], 0, g1m[0], 2), (g1m[70], 0, g1m[6], 2),
             (g1m[65], 1, g1m[2], 2), (g1m[3], 0, g1m[3], 0), (g1m[2], 0, g1m[0], 1), (g1m[5], 1, g1m[5], 2),
             (g1m[64], 0, g1m[1], 2)]
    a = K.fr_from_ints([0xDEADBEEF12345678901234567890])[0]
    cases.append((K.p1_mult(g1m[1], a), 0, K.p1_mult(g1m[0], a), 1))
    inf = np.zeros(18, np.uint64)
    cases += [(inf, 0, inf, 1), (inf, 0, g1m[0], 1), (g1m[1], 0, inf, 1)]
    seen = set()
    for a1, qa, b1, qb in cases:
        want = K.pairings_verify(a1, q[qa], b1, q[qb])
        assert ts.pairings_verify(a1, qa, b1, qb) == want
        seen.add(want)
    assert seen == {True, False}


def test_verify_kzg_proof_vectors(B, ts, vectors):
    cases = vectors["verify_kzg_proof"]
    assert len(cases) == 122
    for c in cases:
        got = _run(B, lambda: ts.verify_kzg_proof(H(c["commitment"]), H(c["z"]), H(c["y"]), H(c["proof"])))
        assert got == c["output"], c["name"]


def test_verify_blob_kzg_proof_vectors(B, ts, vectors, golden_blobs):
    cases = vectors["verify_blob_kzg_proof"]
    assert len(cases) == 29
    for c in cases:
        got = _run(B, lambda: ts.verify_blob_kzg_proof(_blob_any(c, golden_blobs), H(c["commitment"]), H(c["proof"])))
        assert got == c["output"], c["name"]


def test_verify_blob_kzg_proof_batch_vectors(B, ts, vectors, golden_blobs):
    cases = vectors["verify_blob_kzg_proof_batch"]
    assert len(cases) == 24
    for c in cases:
        got = _run(B, lambda: ts.verify_blob_kzg_proof_batch([_blob_any(b, golden_blobs) for b in c["blobs"]],
                                                             [H(x) for x in c["commitments"]], [H(x) for x in c["proofs"]]))
        assert got == c["output"], c["name"]


def _random_blobs(rng, n):
    blobs = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    return blobs.reshape(n, -1)


def test_round_trip_batch_70(B, K, ts, oracle_settings):
    """produce -> verify round trip at more than one chunk (70 > max_batch 64); flipping one proof or one blob byte
    must flip the verdict, and the oracle agrees on a sample"""
    rng = np.random.default_rng(11)
    n = 70
    blobs = _random_blobs(rng, n)
    comm = ts.blob_to_kzg_commitment_batch(blobs)
    proofs = ts.compute_blob_kzg_proof_batch(blobs, comm)
    assert ts.verify_blob_kzg_proof_batch(blobs, comm, proofs) is True
    for i in (0, 33, 69):
        assert ts.verify_blob_kzg_proof(blobs[i], comm[i], proofs[i]) is True
    bad = proofs.copy()
    bad[41] = proofs[40]
    assert ts.verify_blob_kzg_proof_batch(blobs, comm, bad) is False
    assert ts.verify_blob_kzg_proof(blobs[41], comm[41], bad[41]) is False
    blobs2 = blobs.copy()
    blobs2[66, 31] ^= 1
    assert ts.verify_blob_kzg_proof_batch(blobs2, comm, proofs) is False
    sub = slice(38, 43)
    want = K.verify_blob_kzg_proof_batch([blobs[i].tobytes() for i in range(38, 43)], [comm[i].tobytes() for i in range(38, 43)],
                                         [bad[i].tobytes() for i in range(38, 43)], oracle_settings)
    assert ts.verify_blob_kzg_proof_batch(blobs[sub], comm[sub], bad[sub]) == want is False
    assert ts.verify_blob_kzg_proof_batch(blobs[:0], comm[:0], proofs[:0]) is True


def test_verify_kzg_proof_batch_random_points(B, K, ts, oracle_settings):
    """compute_kzg_proof at random z, then the (C, z, y, proof) batch verifier; wrong y is rejected; non-canonical y errors"""
    rng = np.random.default_rng(12)
    n = 9
    blobs = _random_blobs(rng, n)
    zs = np.array([list(v.to_bytes(32, "big")) for v in rand_ints(rng, n, R_MOD)], dtype=np.uint8)
    comm = ts.blob_to_kzg_commitment_batch(blobs)
    proofs, ys = ts.compute_kzg_proof_batch(blobs, zs)
    assert ts.verify_kzg_proof_batch(comm, zs, ys, proofs) is True
    for i in range(n):
        assert ts.verify_kzg_proof(comm[i], zs[i], ys[i], proofs[i]) is True
        assert K.verify_kzg_proof(comm[i].tobytes(), zs[i].tobytes(), ys[i].tobytes(), proofs[i].tobytes(), oracle_settings)
    ys2 = ys.copy()
    ys2[4, 31] ^= 1
    assert ts.verify_kzg_proof_batch(comm, zs, ys2, proofs) is False
    assert ts.verify_kzg_proof(comm[4], zs[4], ys2[4], proofs[4]) is False
    ys3 = ys.copy()
    ys3[2] = np.frombuffer(R_MOD.to_bytes(32, "big"), np.uint8)
    with pytest.raises(B.KzgError):
        ts.verify_kzg_proof_batch(comm, zs, ys3, proofs)


def test_setup_in_monomial_form_is_rejected(B, setup_text):
    """load_trusted_setup runs the reference's pairing check (kzg/src/eip_4844.rs:1005-1020, 1064-1068) on the device"""
    toks = setup_text.split()
    n1, n2 = int(toks[0]), int(toks[1])
    unhex = lambda ts_: bytes.fromhex("".join(ts_))
    lag, g2, mono = unhex(toks[2:2 + n1]), unhex(toks[2 + n1:2 + n1 + n2]), unhex(toks[2 + n1 + n2:])
    s = B.KZGSettings.load_trusted_setup(mono, lag, g2)
    s.free()
    with pytest.raises(B.KzgError) as e:
        B.KZGSettings.load_trusted_setup(mono, mono, g2)
    assert e.value.code == 1
    bad_g2 = bytearray(g2)
    bad_g2[96 + 95] ^= 1                      # [s]G2 with a flipped x bit: off the curve (or a different point)
    try:
        s2 = B.KZGSettings.load_trusted_setup(mono, lag, bytes(bad_g2))
        s2.free()
        loaded = True
    except B.KzgError as e2:
        loaded = False
        assert e2.code == 1
    assert loaded in (True, False)
