/* ckzg_roundtrip.c -- a plain-C consumer of libb200kzg.so through the c-kzg-4844 ABI (include/b200_kzg.h, section B2):
 * exactly the calls a language binding of c-kzg-4844 makes (compare kzg-bench/src/tests/c_bindings.rs in the reference).
 *   gcc -O2 -Iinclude examples/ckzg_roundtrip.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckzg_roundtrip
 *   /tmp/ckzg_roundtrip rust-kzg_b200/data/trusted_setup.txt
 * Prints "ok" and exits 0 when commit -> prove -> verify, cells -> recover -> verify_cells all agree. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200_kzg.h"

#define CHECK(x) do { if (!(x)) { fprintf(stderr, "FAILED: %s (line %d)\n", #x, __LINE__); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s trusted_setup.txt\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "r");
    CHECK(f != NULL);
    KZGSettings s;
    C_KZG_RET rc = load_trusted_setup_file(&s, f);
    fclose(f);
    if (rc != C_KZG_OK) { fprintf(stderr, "load_trusted_setup_file: %d (no CUDA device?)\n", rc); return 3; }

    Blob *blob = (Blob *)calloc(1, sizeof(Blob));
    unsigned x = 12345;
    for (size_t i = 0; i < sizeof(blob->bytes); i++) {      /* field elements with a zero top byte are canonical */
        x = x * 1103515245u + 12345u;
        blob->bytes[i] = (i % 32 == 0) ? 0 : (uint8_t)(x >> 16);
    }
    KZGCommitment c;
    KZGProof p, pz;
    Bytes32 z, y;
    bool ok = false;
    CHECK(blob_to_kzg_commitment(&c, blob, &s) == C_KZG_OK);
    CHECK(compute_blob_kzg_proof(&p, blob, &c, &s) == C_KZG_OK);
    CHECK(verify_blob_kzg_proof(&ok, blob, &c, &p, &s) == C_KZG_OK && ok);
    memcpy(z.bytes, blob->bytes + 64, 32);
    CHECK(compute_kzg_proof(&pz, &y, blob, &z, &s) == C_KZG_OK);
    CHECK(verify_kzg_proof(&ok, &c, &z, &y, &pz, &s) == C_KZG_OK && ok);
    y.bytes[31] ^= 1;
    CHECK(verify_kzg_proof(&ok, &c, &z, &y, &pz, &s) == C_KZG_OK && !ok);
    blob->bytes[0] = 0xff;                                   /* >= r: every entry point must refuse the blob */
    CHECK(blob_to_kzg_commitment(&c, blob, &s) == C_KZG_BADARGS);
    blob->bytes[0] = 0;

    Cell *cells = (Cell *)malloc(128 * sizeof(Cell)), *half = (Cell *)malloc(64 * sizeof(Cell)), *rec = (Cell *)malloc(128 * sizeof(Cell));
    KZGProof *proofs = (KZGProof *)malloc(128 * sizeof(KZGProof)), *rproofs = (KZGProof *)malloc(128 * sizeof(KZGProof));
    uint64_t idx[128];
    Bytes48 comms[128];
    CHECK(compute_cells_and_kzg_proofs(cells, proofs, blob, &s) == C_KZG_OK);
    for (int i = 0; i < 64; i++) { idx[i] = 2 * i + 1; half[i] = cells[2 * i + 1]; }
    CHECK(recover_cells_and_kzg_proofs(rec, rproofs, idx, half, 64, &s) == C_KZG_OK);
    CHECK(memcmp(rec, cells, 128 * sizeof(Cell)) == 0 && memcmp(rproofs, proofs, 128 * sizeof(KZGProof)) == 0);
    for (int i = 0; i < 128; i++) { idx[i] = i; comms[i] = c; }
    CHECK(verify_cell_kzg_proof_batch(&ok, comms, idx, cells, proofs, 128, &s) == C_KZG_OK && ok);
    proofs[5] = proofs[6];
    CHECK(verify_cell_kzg_proof_batch(&ok, comms, idx, cells, proofs, 128, &s) == C_KZG_OK && !ok);

    free_trusted_setup(&s);
    CHECK(s.g1_values_lagrange_brp == NULL);
    free(blob); free(cells); free(half); free(rec); free(proofs); free(rproofs);
    puts("ok");
    return 0;
}
