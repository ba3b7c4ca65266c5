#!/usr/bin/env python3
"""Extract the reference's consensus-spec golden vectors for the hot path into compact fixtures.

Run HERE (the build container), where /root/reference exists; the GPU box only sees the output.
Source: /root/reference/kzg-bench/src/test_vectors/<fn>/kzg-mainnet/<case>/data.yaml
        (runners: kzg-bench/src/tests/eip_4844.rs:538-1010, kzg-bench/src/tests/eip_7594.rs:23-468)
Output: tests/golden/blobs.bin       unique well-formed (131072 B) blobs, concatenated
        tests/golden/vectors.json    per-function case lists; blobs referenced by index ("blob": N) or,
                                     for malformed lengths, just the length ("blob_len")
        tests/golden/cells.bin       unique well-formed (2048 B) cells of the EIP-7594 vectors, concatenated;
                                     vectors.json refers to them by index
        tests/golden/cells_sha256.json  sha256 of each of the 128 cells for compute_cells (full cells would be
                                     256 KiB per case; cell 0 and cell 127 are kept verbatim)
Only data (test vectors) is extracted, no reference source code.
"""
import hashlib, json, os, sys, yaml

REF = "/root/reference/kzg-bench/src/test_vectors"
OUT = os.path.dirname(os.path.abspath(__file__))

blobs, blob_index = [], {}

def unhex(s):
    assert s.startswith("0x"), s[:10]
    return bytes.fromhex(s[2:])

def blob_ref(hexstr):
    try:
        b = unhex(hexstr)
    except ValueError:
        return {"blob_bad_hex_len": len(hexstr) - 2}
    if len(b) != 131072:
        # malformed length: only the length matters (bytes_to_blob rejects it, kzg/src/eip_4844.rs:867-874)
        return {"blob_len": len(b), "blob_fill": sorted(set(b))[:4]}
    h = hashlib.sha256(b).digest()
    if h not in blob_index:
        blob_index[h] = len(blobs)
        blobs.append(b)
    return {"blob": blob_index[h]}

def cases(fn):
    d = os.path.join(REF, fn, "kzg-mainnet")
    for name in sorted(os.listdir(d)):
        with open(os.path.join(d, name, "data.yaml")) as f:
            yield name, yaml.safe_load(f)

vec = {}
vec["blob_to_kzg_commitment"] = [
    dict(name=n, output=y["output"], **blob_ref(y["input"]["blob"])) for n, y in cases("blob_to_kzg_commitment")]
vec["compute_kzg_proof"] = [
    dict(name=n, z=y["input"]["z"], output=y["output"], **blob_ref(y["input"]["blob"]))
    for n, y in cases("compute_kzg_proof")]
vec["compute_blob_kzg_proof"] = [
    dict(name=n, commitment=y["input"]["commitment"], output=y["output"], **blob_ref(y["input"]["blob"]))
    for n, y in cases("compute_blob_kzg_proof")]
vec["compute_challenge"] = [
    dict(name=n, commitment=y["input"]["commitment"], output=y["output"], **blob_ref(y["input"]["blob"]))
    for n, y in cases("compute_challenge")]
cc = []
for n, y in cases("compute_cells"):
    out = y["output"]
    e = dict(name=n, **blob_ref(y["input"]["blob"]))
    if out is None:
        e["output"] = None
    else:
        assert len(out) == 128
        e["output"] = dict(cell_sha256=[hashlib.sha256(unhex(c)).hexdigest() for c in out],
                           cell0=out[0], cell127=out[127],
                           all_sha256=hashlib.sha256(b"".join(unhex(c) for c in out)).hexdigest())
    cc.append(e)
vec["compute_cells"] = cc
fk = []
for n, y in cases("compute_cells_and_kzg_proofs"):
    out = y["output"]
    e = dict(name=n, **blob_ref(y["input"]["blob"]))
    if out is None:
        e["output"] = None
    else:
        cells, proofs = out
        assert len(cells) == 128 and len(proofs) == 128
        e["output"] = dict(cells_sha256=hashlib.sha256(b"".join(unhex(c) for c in cells)).hexdigest(), proofs=proofs)
    fk.append(e)
vec["compute_cells_and_kzg_proofs"] = fk

# verification vectors (kzg-bench/src/tests/eip_4844.rs:676-1010): outputs are booleans or null (error)
vec["verify_kzg_proof"] = [
    dict(name=n, commitment=y["input"]["commitment"], z=y["input"]["z"], y=y["input"]["y"], proof=y["input"]["proof"],
         output=y["output"]) for n, y in cases("verify_kzg_proof")]
vec["verify_blob_kzg_proof"] = [
    dict(name=n, commitment=y["input"]["commitment"], proof=y["input"]["proof"], output=y["output"],
         **blob_ref(y["input"]["blob"])) for n, y in cases("verify_blob_kzg_proof")]
vec["verify_blob_kzg_proof_batch"] = [
    dict(name=n, blobs=[blob_ref(b) for b in y["input"]["blobs"]], commitments=y["input"]["commitments"],
         proofs=y["input"]["proofs"], output=y["output"]) for n, y in cases("verify_blob_kzg_proof_batch")]

# EIP-7594 recovery / cell verification (kzg-bench/src/tests/eip_7594.rs:190-468): cells are 2048-byte strings that repeat
# across cases, so they go into a pool (cells.bin) and are referenced by index; malformed ones stay inline.
cells_pool, cell_index = [], {}

def cell_ref(x):
    if isinstance(x, list):   # cosets_evals: 64 field elements
        x = "0x" + "".join(e[2:] for e in x)
    try:
        b = unhex(x)
    except ValueError:
        return {"hex": x}
    if len(b) != 2048:
        return {"hex": x}
    h = hashlib.sha256(b).digest()
    if h not in cell_index:
        cell_index[h] = len(cells_pool)
        cells_pool.append(b)
    return cell_index[h]

vc = []
for n, y in cases("verify_cell_kzg_proof_batch"):
    i = y["input"]
    vc.append(dict(name=n, commitments=i["commitments"], cell_indices=i["cell_indices"],
                   cells=[cell_ref(c) for c in (i["cells"] or [])], proofs=i["proofs"], output=y["output"]))
vec["verify_cell_kzg_proof_batch"] = vc
rc = []
for n, y in cases("recover_cells_and_kzg_proofs"):
    i, out = y["input"], y["output"]
    e = dict(name=n, cell_indices=i["cell_indices"], cells=[cell_ref(c) for c in (i["cells"] or [])])
    if out is None:
        e["output"] = None
    else:
        cells, proofs = out
        assert len(cells) == 128 and len(proofs) == 128
        e["output"] = dict(cells=[cell_ref(c) for c in cells], proofs=proofs)
    rc.append(e)
vec["recover_cells_and_kzg_proofs"] = rc
ch = []
for n, y in cases("compute_verify_cell_kzg_proof_batch_challenge"):
    i = y["input"]
    ch.append(dict(name=n, commitments=i["commitments"], commitment_indices=i["commitment_indices"],
                   cell_indices=i["cell_indices"], cells=[cell_ref(c) for c in (i["cosets_evals"] or [])],
                   proofs=i["proofs"], output=y["output"]))
vec["compute_verify_cell_kzg_proof_batch_challenge"] = ch
with open(os.path.join(OUT, "cells.bin"), "wb") as f:
    for b in cells_pool:
        f.write(b)
print("unique cells:", len(cells_pool))

with open(os.path.join(OUT, "blobs.bin"), "wb") as f:
    for b in blobs:
        f.write(b)
with open(os.path.join(OUT, "vectors.json"), "w") as f:
    json.dump(vec, f, indent=0)
print("unique blobs:", len(blobs), {k: len(v) for k, v in vec.items()})

# --- other fixtures the reference's tests pin for this path (data, not source) -------------------------------
import shutil
# mainnet trusted setup (text format parsed by kzg/src/eip_4844.rs:151-228) -- also the bench's fixed bases
shutil.copyfile("/root/reference/kzg-bench/src/trusted_setup.txt",
                os.path.join(OUT, "..", "..", "rust-kzg_b200", "data", "trusted_setup.txt"))
# 1000 x 48 B = compress(i*G), i = 0..999 (zkcrypto/bls12_381/src/tests/mod.rs:3-52)
shutil.copyfile("/root/reference/zkcrypto/bls12_381/src/tests/g1_compressed_valid_test_vectors.dat",
                os.path.join(OUT, "g1_compressed_valid_test_vectors.dat"))

# --- hard-coded known-answer tables in the reference's Rust tests (numbers only) ------------------------------
import re
def u64_tables(path, name):
    src = open(path).read()
    i = src.index("=", src.index(name))
    j = src.index("];\n", i)
    rows = re.findall(r"\[([^\[\]]+)\]", src[i:j])
    out = []
    for r in rows:
        vals = [int(t.strip().replace("_", ""), 0) for t in r.split(",") if t.strip()]
        if len(vals) == 4:
            out.append(vals)
    return out
T = "/root/reference/kzg-bench/src/tests/"
kats = {
  # kzg-bench/src/tests/fft_fr.rs:49-84 : inverse fft_fr of data[i]=i, scale 4
  "inv_fft_expected": u64_tables(T + "fft_fr.rs", "inv_fft_expected"),
  # kzg-bench/src/tests/das.rs:4-31 : das_fft_extension of evens[i]=i, scale 4
  "das_expected_u": u64_tables(T + "das.rs", "expected_u"),
  # kzg-bench/src/tests/eip_4844.rs:48-60 : powers of 32930439
  "expected_powers": u64_tables(T + "eip_4844.rs", "EXPECTED_POWERS"),
  # blst/src/consts.rs:17-50
  "scale2_root_of_unity": u64_tables("/root/reference/blst/src/consts.rs", "SCALE2_ROOT_OF_UNITY"),
  # kzg-bench/src/tests/eip_4844.rs:85-121
  "commitment_kat": {"blob0": "0x14629a3a39f7b854e6aa49aa2edb450267eac2c14bb2d4f97a0b81a3f57055ad",
     "commitment": "0x91a5e1c143820d2e7bec38a5404c5145807cb88c0abbbecbcb4bccc83a4b417326e337574cff43303f8a6648ecbee7ac"},
  # kzg-bench/src/tests/eip_4844.rs:124-175
  "proof_kat": {"blob0": "0x69386e69dbae0357b399b8d645a57a3062dfbe00bd8e97170b9bdd6bc6168a13",
     "z": "0x03ea4fb841b4f9e01aa917c5e40dbd67efb4b8d4d9052069595f0647feba320d",
     "proof": "0xb21f8f9b85e52fd9c4a6d4fb4e9a27ebdc5a09c3f5ca17f6bcd85c26f04953b0e6925607aaebed1087e5cc2fe4b2b356"},
}
assert len(kats["inv_fft_expected"]) == 16 and len(kats["das_expected_u"]) == 8
assert len(kats["expected_powers"]) == 11 and len(kats["scale2_root_of_unity"]) == 32
with open(os.path.join(OUT, "kats.json"), "w") as f:
    json.dump(kats, f, indent=0)
print("kats ok")
