"""EIP-7594 recovery and cell-proof batch verification on the device vs the reference's consensus-spec vectors
(kzg-bench/src/test_vectors/{recover_cells_and_kzg_proofs, verify_cell_kzg_proof_batch,
compute_verify_cell_kzg_proof_batch_challenge}, runners kzg-bench/src/tests/eip_7594.rs:190-468) and round trips."""
import numpy as np
import pytest

from conftest import cell_of

pytestmark = pytest.mark.gpu


def H(x):
    return bytes.fromhex(x[2:])


@pytest.fixture(scope="module")
def ts(B):
    s = B.KZGSettings.load_trusted_setup_file()
    yield s
    s.free()


def _run(B, fn):
    try:
        return fn()
    except (B.KzgError, ValueError, OverflowError):
        return None


def test_cell_batch_challenge_vectors(B, K, ts, vectors, golden_cells):
    cases = vectors["compute_verify_cell_kzg_proof_batch_challenge"]
    assert len(cases) == 10
    for c in cases:
        got = _run(B, lambda: ts.compute_verify_cell_kzg_proof_batch_challenge(
            [H(x) for x in c["commitments"]], c["commitment_indices"], c["cell_indices"],
            [cell_of(x, golden_cells) for x in c["cells"]], [H(x) for x in c["proofs"]]))
        out = None if got is None else "0x" + K.fr_to_bytes(got).hex()
        assert out == c["output"], c["name"]


def test_verify_cell_kzg_proof_batch_vectors(B, ts, vectors, golden_cells):
    cases = vectors["verify_cell_kzg_proof_batch"]
    assert len(cases) == 32
    for c in cases:
        got = _run(B, lambda: ts.verify_cell_kzg_proof_batch([H(x) for x in c["commitments"]], c["cell_indices"],
                                                             [cell_of(x, golden_cells) for x in c["cells"]],
                                                             [H(x) for x in c["proofs"]]))
        assert got == c["output"], c["name"]


def test_recover_cells_and_kzg_proofs_vectors(B, ts, vectors, golden_cells):
    cases = vectors["recover_cells_and_kzg_proofs"]
    assert len(cases) == 18
    for c in cases:
        got = _run(B, lambda: ts.recover_cells_and_kzg_proofs(c["cell_indices"], [cell_of(x, golden_cells) for x in c["cells"]]))
        want = c["output"]
        if want is None:
            assert got is None, c["name"]
        else:
            assert got is not None, c["name"]
            assert got[0] == [cell_of(x, golden_cells) for x in want["cells"]], c["name"]
            assert got[1] == [H(x) for x in want["proofs"]], c["name"]


def test_cells_round_trip_random_blob(B, K, ts, oracle_settings):
    """compute -> drop cells -> recover -> verify, on a seeded random blob; recovery of inconsistent cells agrees with the
    oracle (same sequence of exact field operations)"""
    rng = np.random.default_rng(21)
    blob = rng.integers(0, 256, size=(4096, 32), dtype=np.uint8)
    blob[:, 0] = 0
    blob = blob.tobytes()
    cells, proofs = ts.compute_cells_and_kzg_proofs(blob)
    commitment = ts.blob_to_kzg_commitment(blob)
    keep = sorted(rng.choice(128, size=77, replace=False).tolist())
    rc, rp = ts.recover_cells_and_kzg_proofs(keep, [cells[i] for i in keep])
    assert rc == cells and rp == proofs
    rc2, none = ts.recover_cells_and_kzg_proofs(keep[:64], [cells[i] for i in keep[:64]], want_proofs=False)
    assert rc2 == cells and none is None
    idx = rng.choice(128, size=40, replace=True).tolist()
    assert ts.verify_cell_kzg_proof_batch([commitment] * 40, idx, [cells[i] for i in idx], [proofs[i] for i in idx]) is True
    bad = [proofs[i] for i in idx]
    bad[7] = proofs[(idx[7] + 1) % 128]
    assert ts.verify_cell_kzg_proof_batch([commitment] * 40, idx, [cells[i] for i in idx], bad) is False
    # inconsistent input: one provided cell replaced by another; still a well-defined computation
    sub = keep[:70]
    cs = [cells[i] for i in sub]
    cs[5] = cells[(sub[5] + 1) % 128]
    got = ts.recover_cells_and_kzg_proofs(sub, cs, want_proofs=False)[0]
    want = K.recover_cells_and_kzg_proofs(sub, cs, oracle_settings, want_proofs=False)[0]
    assert got == want


def test_verify_cells_many_blobs_shuffled(B, K, ts, oracle_settings):
    """384 cells of three blobs in random order (three unique commitments, repeated cells): accepted; a cell moved to the
    wrong commitment is rejected; the oracle agrees on a 24-cell sample"""
    rng = np.random.default_rng(22)
    blobs = rng.integers(0, 256, size=(3, 4096, 32), dtype=np.uint8)
    blobs[:, :, 0] = 0
    blobs = [b.tobytes() for b in blobs.reshape(3, -1)]
    comms = [ts.blob_to_kzg_commitment(b) for b in blobs]
    per_blob = [ts.compute_cells_and_kzg_proofs(b) for b in blobs]
    items = [(k, i) for k in range(3) for i in range(128)] + [(1, 5), (1, 5), (2, 127)]
    order = rng.permutation(len(items))
    items = [items[o] for o in order]
    c = [comms[k] for k, i in items]
    idx = [i for k, i in items]
    cells = [per_blob[k][0][i] for k, i in items]
    proofs = [per_blob[k][1][i] for k, i in items]
    assert ts.verify_cell_kzg_proof_batch(c, idx, cells, proofs) is True
    wrong = list(c)
    wrong[17] = comms[(items[17][0] + 1) % 3]
    assert ts.verify_cell_kzg_proof_batch(wrong, idx, cells, proofs) is False
    sub = slice(10, 34)
    assert K.verify_cell_kzg_proof_batch(c[sub], idx[sub], cells[sub], proofs[sub], oracle_settings) is True
    assert K.verify_cell_kzg_proof_batch(wrong[sub], idx[sub], cells[sub], proofs[sub], oracle_settings) is False
    assert ts.verify_cell_kzg_proof_batch(wrong[sub], idx[sub], cells[sub], proofs[sub]) is False
