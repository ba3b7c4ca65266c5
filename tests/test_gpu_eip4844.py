"""c-kzg-4844 commitment / proof functions through the C ABI vs the reference's consensus-spec vectors
(kzg-bench/src/test_vectors/*, extracted to tests/golden/) and vs the oracle on random blobs."""
import numpy as np
import pytest

from conftest import R_MOD

pytestmark = pytest.mark.gpu


def H(x):
    return bytes.fromhex(x[2:])


@pytest.fixture(scope="module")
def ts(B):
    s = B.KZGSettings.load_trusted_setup_file()
    yield s
    s.free()


def _blob_of(case, golden_blobs):
    if "blob" in case:
        return golden_blobs[case["blob"]]
    return bytes([0]) * case["blob_len"]      # malformed length: content irrelevant


def test_settings_arrays_match_oracle(B, K, ts, oracle_settings):
    """CKZGSettings host arrays (kzg_settings_to_c, blst/src/eip_4844.rs:40-144) are the reference's values"""
    assert np.array_equal(ts.array("roots_of_unity", 8193, 4), oracle_settings.fs.roots_of_unity)
    assert np.array_equal(ts.array("brp_roots_of_unity", 8192, 4), oracle_settings.fs.brp_roots_of_unity)
    assert np.array_equal(ts.array("reverse_roots_of_unity", 8193, 4), oracle_settings.fs.reverse_roots_of_unity)
    assert np.array_equal(ts.array("g1_values_lagrange_brp", 4096, 18), oracle_settings.g1_lagrange_brp)
    assert np.array_equal(ts.array("g1_values_monomial", 4096, 18), oracle_settings.g1_monomial)


def test_x_ext_fft_columns_match_oracle(B, K, ts, oracle_settings):
    """the host x_ext_fft_columns a TryFrom<&CKZGSettings>-style reader walks (blst/src/types/kzg_settings.rs:398-417):
    128 non-NULL rows of 64 points, the same group elements as FsKZGSettings::new builds (:84-101)"""
    got = ts.x_ext_fft_columns()
    want = K.x_ext_fft_columns(oracle_settings)
    assert got.shape == want.shape == (128, 64, 18)
    for r in (0, 1, 63, 64, 127):
        for o in (0, 17, 63):
            assert K.p1_compress(got[r, o]) == K.p1_compress(want[r, o]), (r, o)
    # all of them, through one batched comparison of affine forms
    assert np.array_equal(K.p1s_to_affine(got.reshape(-1, 18)), K.p1s_to_affine(want.reshape(-1, 18)))


def test_blob_to_kzg_commitment_vectors(B, ts, vectors, golden_blobs):
    for c in vectors["blob_to_kzg_commitment"]:
        try:
            out = "0x" + ts.blob_to_kzg_commitment(_blob_of(c, golden_blobs)).hex()
        except B.KzgError:
            out = None
        assert out == c["output"], c["name"]


def test_compute_kzg_proof_vectors(B, ts, vectors, golden_blobs):
    for c in vectors["compute_kzg_proof"]:
        try:
            p, y = ts.compute_kzg_proof(_blob_of(c, golden_blobs), H(c["z"]))
            out = ["0x" + p.hex(), "0x" + y.hex()]
        except (B.KzgError, ValueError):
            out = None
        assert out == c["output"], c["name"]


def test_compute_blob_kzg_proof_vectors(B, ts, vectors, golden_blobs):
    for c in vectors["compute_blob_kzg_proof"]:
        try:
            out = "0x" + ts.compute_blob_kzg_proof(_blob_of(c, golden_blobs), H(c["commitment"])).hex()
        except (B.KzgError, ValueError):
            out = None
        assert out == c["output"], c["name"]


def test_commitment_and_proof_kats(B, ts, kats):
    """kzg-bench/src/tests/eip_4844.rs:85-175"""
    k = kats["commitment_kat"]
    blob = H(k["blob0"]) + bytes(131072 - 32)
    assert "0x" + ts.blob_to_kzg_commitment(blob).hex() == k["commitment"]
    k = kats["proof_kat"]
    blob = H(k["blob0"]) + bytes(131072 - 32)
    proof, _ = ts.compute_kzg_proof(blob, H(k["z"]))
    assert "0x" + proof.hex() == k["proof"]


def _rand_blobs(rng, n):
    """kzg-bench/src/tests/eip_4844.rs:28-37: random bytes with the top byte of every element zeroed"""
    b = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    b[:, :, 0] = 0
    return b.reshape(n, 131072)


def test_batches_match_oracle(B, K, ts, oracle_settings):
    rng = np.random.default_rng(0x4B5A47)
    n = 5
    blobs = _rand_blobs(rng, n)
    blobs[1, 32:] = 0                                      # sparse blob
    blobs[2] = np.tile(blobs[2, :32], 4096)                # all elements equal (one hot bucket per window)
    oracle_settings.set_threads(8)
    comm = ts.blob_to_kzg_commitment_batch(blobs)
    for i in range(n):
        assert comm[i].tobytes() == K.blob_to_kzg_commitment(blobs[i].tobytes(), oracle_settings), i
    zs = _rand_blobs(rng, 1)[0, : 32 * n].reshape(n, 32).copy()
    # z inside the evaluation domain (kzg/src/eip_4844.rs:458-462, 484-510): a bit-reversed root of unity
    zs[3] = np.frombuffer(K.fr_to_bytes(oracle_settings.fs.brp_roots_of_unity[5]), dtype=np.uint8)
    zs[4] = np.frombuffer(K.fr_to_bytes(oracle_settings.fs.brp_roots_of_unity[0]), dtype=np.uint8)   # z = 1
    proofs, ys = ts.compute_kzg_proof_batch(blobs, zs)
    for i in range(n):
        ep, ey = K.compute_kzg_proof(blobs[i].tobytes(), zs[i].tobytes(), oracle_settings)
        assert (proofs[i].tobytes(), ys[i].tobytes()) == (ep, ey), i
    bp = ts.compute_blob_kzg_proof_batch(blobs, comm)
    for i in range(n):
        assert bp[i].tobytes() == K.compute_blob_kzg_proof(blobs[i].tobytes(), comm[i].tobytes(), oracle_settings), i
    oracle_settings.set_threads(1)


def test_full_batch_of_64_and_chunking(B, K, ts, oracle_settings):
    """BASELINE config 3: 64 blobs per call; 70 exercises the chunking past the context capacity"""
    rng = np.random.default_rng(64)
    blobs = _rand_blobs(rng, 70)
    comm = ts.blob_to_kzg_commitment_batch(blobs)
    oracle_settings.set_threads(8)
    for i in (0, 31, 63, 64, 69):
        assert comm[i].tobytes() == K.blob_to_kzg_commitment(blobs[i].tobytes(), oracle_settings), i
    proofs = ts.compute_blob_kzg_proof_batch(blobs, comm)
    for i in (0, 63, 69):
        assert proofs[i].tobytes() == K.compute_blob_kzg_proof(blobs[i].tobytes(), comm[i].tobytes(), oracle_settings), i
    oracle_settings.set_threads(1)
    # size-independent property: the commitment is linear in the blob -> commit(a) + commit(b) == commit(a + b)
    a = K.fr_from_ints([int.from_bytes(blobs[0, 32 * i:32 * i + 32].tobytes(), "big") for i in range(4096)])
    b = K.fr_from_ints([int.from_bytes(blobs[1, 32 * i:32 * i + 32].tobytes(), "big") for i in range(4096)])
    s = K.fr_add(a, b)
    sum_blob = b"".join(K.fr_to_bytes(x) for x in s)
    lhs = K.p1_add(K.p1_uncompress(comm[0].tobytes()), K.p1_uncompress(comm[1].tobytes()))
    assert K.p1_compress(lhs) == ts.blob_to_kzg_commitment(sum_blob)


def test_error_cases(B, K, ts, golden_blobs):
    """kzg-bench/src/tests/c_bindings.rs:65-97, 584-616 and SURVEY.md appendix B.8"""
    good = golden_blobs[3]
    bad = bytearray(good)
    bad[64:96] = R_MOD.to_bytes(32, "big")                 # element == r: non-canonical
    with pytest.raises(B.KzgError):
        ts.blob_to_kzg_commitment(bytes(bad))
    with pytest.raises(B.KzgError):
        ts.compute_kzg_proof(bytes(bad), bytes(32))
    with pytest.raises(B.KzgError):
        ts.compute_kzg_proof(good, R_MOD.to_bytes(32, "big"))          # z == r
    comm = ts.blob_to_kzg_commitment(good)
    with pytest.raises(B.KzgError):
        ts.compute_blob_kzg_proof(bytes(bad), comm)
    # commitment at infinity is accepted
    inf = bytes([0xC0]) + bytes(47)
    assert ts.compute_blob_kzg_proof(good, inf) == K.compute_blob_kzg_proof(good, inf, K.KZGSettings(open(B.default_trusted_setup_path()).read()))
    # not compressed / not on curve / wrong subgroup
    with pytest.raises(B.KzgError):
        ts.compute_blob_kzg_proof(good, bytes(48))
    x = 5
    P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
    found = None
    while found is None:                                   # a curve point outside G1: x such that x^3+4 is a square
        y2 = (x ** 3 + 4) % P
        y = pow(y2, (P + 1) // 4, P)
        if y * y % P == y2:
            enc = bytearray(x.to_bytes(48, "big"))
            enc[0] |= 0x80 | (0x20 if y > (P - 1) // 2 else 0)
            pt = K.p1_uncompress(bytes(enc))
            if not K.p1_in_g1(pt):
                found = bytes(enc)
        x += 1
    with pytest.raises(B.KzgError):
        ts.compute_blob_kzg_proof(good, found)
    # a batch with one bad element fails as a whole
    blobs = np.frombuffer(good + bytes(bad), dtype=np.uint8).reshape(2, 131072)
    with pytest.raises(B.KzgError):
        ts.blob_to_kzg_commitment_batch(blobs)


def test_malformed_trusted_setups(B, tmp_path):
    """load_trusted_setup error paths (kzg-bench/src/tests/fixtures/*): wrong counts, bad hex, bad points"""
    text = open(B.default_trusted_setup_path()).read()
    toks = text.split()
    cases = {
        "wrong_g1_count": ["4095"] + toks[1:],
        "wrong_g2_count": [toks[0], "64"] + toks[2:],
        "truncated": toks[:500],
        "bad_hex": toks[:2] + ["zz" + toks[2][2:]] + toks[3:],
        "point_not_on_curve": toks[:2] + [toks[2][:-2] + ("00" if toks[2][-2:] != "00" else "01")] + toks[3:],
    }
    for name, tk in cases.items():
        p = tmp_path / (name + ".txt")
        p.write_text("\n".join(tk) + "\n")
        with pytest.raises(B.KzgError):
            B.KZGSettings.load_trusted_setup_file(str(p))
    # free is idempotent and NULL-safe; a freed handle is rejected (kzg-bench/src/tests/c_bindings.rs free tests)
    s = B.KZGSettings.load_trusted_setup_file()
    s.free()
    s.free()
    assert s.c.g1_values_lagrange_brp is None and s.c.roots_of_unity is None
    B.lib().free_trusted_setup(None)
    s.loaded = False
    with pytest.raises(B.KzgError):
        s.blob_to_kzg_commitment(bytes(131072))


def test_load_from_bytes(B, K, setup_text):
    """load_trusted_setup (bytes form, blst/src/eip_4844.rs:180-222)"""
    from oracle import kzg_oracle as O
    mono, lag, g2 = O.load_trusted_setup_string(setup_text)
    s = B.KZGSettings.load_trusted_setup(b"".join(mono), b"".join(lag), b"".join(g2))
    assert s.blob_to_kzg_commitment(bytes(131072)) == bytes([0xC0]) + bytes(47)
    s.free()
    with pytest.raises(B.KzgError):
        B.KZGSettings.load_trusted_setup(b"".join(mono[:-1]), b"".join(lag), b"".join(g2))


def test_sha256_paths(B):
    import hashlib
    from rust_kzg_b200 import eip4844
    rng = np.random.default_rng(2)
    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 1000, 131152):
        msg = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert eip4844.sha256(msg) == hashlib.sha256(msg).digest()
        assert eip4844.sha256(msg, portable=True) == hashlib.sha256(msg).digest()


def test_compute_cells_vectors(B, ts, vectors, golden_blobs):
    """the reference's NTT golden vectors (kzg-bench/src/test_vectors/compute_cells): INTT_4096 + NTT_8192 on the device"""
    import hashlib
    for c in vectors["compute_cells"]:
        try:
            cells = ts.compute_cells(_blob_of(c, golden_blobs))
        except B.KzgError:
            cells = None
        if c["output"] is None:
            assert cells is None, c["name"]
            continue
        assert [hashlib.sha256(x).hexdigest() for x in cells] == c["output"]["cell_sha256"], c["name"]
        assert "0x" + cells[0].hex() == c["output"]["cell0"] and "0x" + cells[127].hex() == c["output"]["cell127"]


def test_compute_cells_batch(B, K, ts, oracle_settings):
    rng = np.random.default_rng(77)
    blobs = _rand_blobs(rng, 3)
    out = ts.compute_cells_batch(blobs)
    for i in range(3):
        exp = K.compute_cells(blobs[i].tobytes(), oracle_settings)
        assert out[i].tobytes() == b"".join(exp), i


def test_compute_cells_and_kzg_proofs_vectors(B, ts, vectors, golden_blobs):
    """kzg-bench/src/test_vectors/compute_cells_and_kzg_proofs: cells + FK20 proofs on the device"""
    import hashlib
    for c in vectors["compute_cells_and_kzg_proofs"]:
        try:
            cells, proofs = ts.compute_cells_and_kzg_proofs(_blob_of(c, golden_blobs))
        except B.KzgError:
            cells = proofs = None
        if c["output"] is None:
            assert cells is None, c["name"]
            continue
        assert hashlib.sha256(b"".join(cells)).hexdigest() == c["output"]["cells_sha256"], c["name"]
        assert ["0x" + p.hex() for p in proofs] == c["output"]["proofs"], c["name"]


def test_cell_proofs_batch_matches_oracle(B, K, ts, oracle_settings):
    rng = np.random.default_rng(78)
    blobs = _rand_blobs(rng, 18)             # more than one FK20 chunk (capacity 16)
    oracle_settings.set_threads(8)
    out = ts.compute_cell_proofs_batch(blobs)
    for i in (0, 15, 17):
        _, exp = K.compute_cells_and_kzg_proofs(blobs[i].tobytes(), oracle_settings, want_cells=False)
        assert out[i].tobytes() == b"".join(exp), i
    oracle_settings.set_threads(1)
    # both outputs NULL is rejected (kzg/src/das.rs:250-252)
    import ctypes as C
    assert B.lib().compute_cells_and_kzg_proofs(None, None, blobs[0].ctypes.data_as(C.c_void_p), C.byref(ts.c)) == 1


def test_cells_and_proofs_batch_in_one_pass(B, K, ts, oracle_settings):
    """b200_compute_cells_and_kzg_proofs_batch: the shared blob -> monomial pass with the cells copied out under the FK20
    kernels gives the bytes of the two separate calls and of the oracle; an invalid blob fails the call before any output"""
    rng = np.random.default_rng(79)
    blobs = _rand_blobs(rng, 5)
    cells, proofs = ts.compute_cells_and_kzg_proofs_batch(blobs)
    assert np.array_equal(cells, ts.compute_cells_batch(blobs))
    assert np.array_equal(proofs, ts.compute_cell_proofs_batch(blobs))
    oracle_settings.set_threads(8)
    oc, op = K.compute_cells_and_kzg_proofs(blobs[4].tobytes(), oracle_settings)
    oracle_settings.set_threads(1)
    assert cells[4].tobytes() == b"".join(oc) and proofs[4].tobytes() == b"".join(op)
    c1, p1 = ts.compute_cells_and_kzg_proofs(blobs[4])           # the c-kzg symbol is the batch of one
    assert b"".join(c1) == cells[4].tobytes() and b"".join(p1) == proofs[4].tobytes()
    bad = blobs.copy()
    bad[2, :32] = 0xff                                            # element 0 of blob 2 >= r
    cells_out = np.full((5, 128, 2048), 7, np.uint8)
    proofs_out = np.full((5, 128, 48), 7, np.uint8)
    with pytest.raises(B.KzgError):
        ts.compute_cells_and_kzg_proofs_batch(bad, cells_out, proofs_out)
    assert (cells_out == 7).all() and (proofs_out == 7).all()


def test_helper_exports_compute_challenge_vectors(B, K, vectors, golden_blobs):
    """compute_challenge / bytes_to_kzg_commitment / bytes_from_bls_field (blst/src/eip_4844.rs:498-530) on the reference's
    compute_challenge vectors"""
    for c in vectors["compute_challenge"]:
        p1 = B.bytes_to_kzg_commitment(H(c["commitment"]))
        assert K.p1_compress(p1) == H(c["commitment"])
        fr = B.compute_challenge(golden_blobs[c["blob"]], p1)
        assert "0x" + B.bytes_from_bls_field(fr).hex() == c["output"], c["name"]
    with pytest.raises(B.KzgError):
        B.bytes_to_kzg_commitment(bytes(48))           # compression flag missing
    x = K.fr_from_ints([R_MOD - 5])[0]
    assert B.bytes_from_bls_field(x) == (R_MOD - 5).to_bytes(32, "big")


def test_direct_small_batch_msm_matches_bucket_engine(B, K, oracle_settings):
    """batches of up to B200_BLOB_DIRECT blobs take the direct-lookup MSM (csrc/fk20_direct.cu: no buckets, one launch), larger
    ones the bucket engine; every table width (13-bit default, 11 and 8 when HBM is short) and the bucket engine must give the
    oracle's bytes, also for adversarial blob contents, and agree at the switch-over"""
    import os
    rng = np.random.default_rng(55)
    blobs = rng.integers(0, 256, size=(26, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    blobs[1] = 0                                            # zero polynomial: commitment = infinity
    blobs[2, :, :31] = 0
    blobs[2, :, 31] = 2                                     # the all-0x02 consensus blob
    blobs[3] = np.frombuffer((R_MOD - 1).to_bytes(32, "big"), np.uint8)   # every element r - 1
    blobs = blobs.reshape(26, -1)
    zs = blobs[0].reshape(4096, 32)[:26].copy()
    want_c = [K.blob_to_kzg_commitment(blobs[i].tobytes(), oracle_settings) for i in range(26)]
    want_p = [K.compute_kzg_proof(blobs[i].tobytes(), zs[i].tobytes(), oracle_settings) for i in (0, 2, 3, 25)]
    for direct, bits in (("24", "13"), ("24", "11"), ("64", "8"), ("0", "13")):
        os.environ["B200_BLOB_DIRECT"] = direct
        os.environ["B200_BLOB_DIRECT_BITS"] = bits
        os.environ["B200_DIRECT_RESERVE_GB"] = "2"         # other settings objects of this session hold tables too
        try:
            ts = B.KZGSettings.load_trusted_setup_file()
        finally:
            del os.environ["B200_BLOB_DIRECT"], os.environ["B200_BLOB_DIRECT_BITS"], os.environ["B200_DIRECT_RESERVE_GB"]
        info = ts.direct_tables()
        assert info["blob_bits"] == (0 if direct == "0" else int(bits)) and info["blob_max_batch"] == int(direct), info
        for n in (1, 3, 24, 25, 26):
            got = ts.blob_to_kzg_commitment_batch(blobs[:n])
            assert [bytes(g) for g in got] == want_c[:n], (direct, bits, n)
        proofs, ys = ts.compute_kzg_proof_batch(blobs[:26], zs)
        for k, i in enumerate((0, 2, 3, 25)):
            assert (bytes(proofs[i]), bytes(ys[i])) == tuple(want_p[k]), (direct, bits, i)
        p1, y1 = ts.compute_kzg_proof(blobs[2], zs[2])      # single call: a batch of one
        assert (p1, y1) == tuple(want_p[1])
        ts.free()


def test_direct_tables_step_down_when_hbm_is_short(B, K, oracle_settings):
    """the tables must leave B200_DIRECT_RESERVE_GB free: with a reserve nothing can satisfy the settings object loads without
    them (bucket engines), with one only a narrow table fits the width steps down -- same bytes either way"""
    import os
    import torch
    rng = np.random.default_rng(57)
    blobs = _rand_blobs(rng, 2)
    want = [K.blob_to_kzg_commitment(blobs[i].tobytes(), oracle_settings) for i in range(2)]
    free_gb = torch.cuda.mem_get_info()[0] / 2**30
    # (reserve, widths the Lagrange table may end up with)
    cases = [(100000, (0,))]
    if free_gb > 12:
        cases.append((int(free_gb) - 10, (8, 11)))       # 10 GiB of headroom: the 7.5 GiB table fits, the 30 GiB one does not
    for reserve, allowed in cases:
        os.environ["B200_DIRECT_RESERVE_GB"] = str(reserve)
        try:
            ts = B.KZGSettings.load_trusted_setup_file()
        finally:
            del os.environ["B200_DIRECT_RESERVE_GB"]
        assert ts.direct_tables()["blob_bits"] in allowed, (reserve, ts.direct_tables(), free_gb)
        got = ts.blob_to_kzg_commitment_batch(blobs)
        assert [bytes(g) for g in got] == want, reserve
        ts.free()


def test_fk20_lincomb_table_widths_agree(B, K, oracle_settings):
    """the FK20 lincomb stage by direct lookups on 13- / 11- / 8-bit tables and on the bucket engine (B200_FK20_DIRECT=0):
    identical cell proofs, and blob 0's equal the oracle's (kzg/src/das.rs:660-696)"""
    import os
    rng = np.random.default_rng(56)
    blobs = rng.integers(0, 256, size=(3, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    blobs[1, :, :] = 0
    blobs[1, :, 31] = 1
    blobs = blobs.reshape(3, -1)
    _, want = K.compute_cells_and_kzg_proofs(blobs[0].tobytes(), oracle_settings)
    ref = None
    for direct, bits in (("1", "13"), ("1", "11"), ("1", "8"), ("0", "8")):
        os.environ["B200_FK20_DIRECT"] = direct
        os.environ["B200_FK20_DIRECT_BITS"] = bits
        os.environ["B200_BLOB_DIRECT"] = "0"               # leave the HBM to the table under test (60 GiB at 13 bits)
        os.environ["B200_DIRECT_RESERVE_GB"] = "2"
        try:
            ts = B.KZGSettings.load_trusted_setup_file()
            got = ts.compute_cell_proofs_batch(blobs)      # the tables are built on first use
        finally:
            del os.environ["B200_FK20_DIRECT"], os.environ["B200_FK20_DIRECT_BITS"], os.environ["B200_BLOB_DIRECT"]
            del os.environ["B200_DIRECT_RESERVE_GB"]
        assert ts.direct_tables()["fk20_bits"] == (int(bits) if direct == "1" else 0), ts.direct_tables()
        assert [bytes(g) for g in got[0]] == [bytes(w) for w in want], (direct, bits)
        if ref is None:
            ref = got.copy()
        assert np.array_equal(got, ref), (direct, bits)
        ts.free()
