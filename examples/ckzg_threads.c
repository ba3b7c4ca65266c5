/* ckzg_threads.c -- a multithreaded plain-C consumer of the UNMODIFIED c-kzg-4844 single-blob entry points, the way the
 * reference's own consumers call them: many host threads, one blob per call (rayon par_chunks over blobs,
 * kzg/src/eip_4844.rs:770-816, into blst/src/eip_4844.rs:163-175, 274-291, 476-496).  Input blobs live in pageable
 * heap memory (malloc), like a Rust Vec<Blob>.
 *
 *   gcc -O2 -pthread -Iinclude examples/ckzg_threads.c -Lrust-kzg_b200 -lb200kzg -Wl,-rpath,$PWD/rust-kzg_b200 -o /tmp/ckzg_threads
 *   /tmp/ckzg_threads rust-kzg_b200/data/trusted_setup.txt <op> <threads> <calls_per_thread> [distinct_blobs]
 *     op: commit | proof | blob_proof | mixed | cells | verify   (mixed: thread t runs op t % 3; cells: compute_cells_and_kzg_proofs;
 *         verify: verify_blob_kzg_proof on valid triples, every answer must be true)
 *
 * Every thread first computes its reference outputs with ONE thread active (nothing to coalesce with), then all threads
 * make one untimed concurrent call each (first multi-request batches allocate staging and load kernels), then they run
 * concurrently under the clock and every result is compared byte for byte with the single-threaded one.  Prints one JSON line:
 *   {"op": ..., "threads": T, "calls": N, "seconds": s, "per_s": N / s, "mismatches": 0, "errors": 0}
 * and exits 0 only if there were no mismatches and no errors. */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "b200_kzg.h"

enum { OP_COMMIT = 0, OP_PROOF = 1, OP_BLOB_PROOF = 2, OP_MIXED = 3, OP_CELLS = 4, OP_VERIFY = 5 };

typedef struct {
    int id, op, calls, nblobs;
    const KZGSettings *s;
    Blob *blobs;          /* nblobs pageable blobs of this thread */
    Bytes48 *commit;      /* per blob: reference commitment */
    Bytes48 *proof;       /* per blob: reference proof of this thread's op */
    Bytes32 *y;           /* per blob: reference y (OP_PROOF) */
    Cell *ref_cells;      /* per blob: 128 reference cells (OP_CELLS) */
    KZGProof *ref_cproofs; /* per blob: 128 reference cell proofs (OP_CELLS) */
    Bytes32 z;
    long mismatches, errors;
    pthread_barrier_t *start, *warm;
} Worker;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int run_one(const Worker *w, int b, Bytes48 *out, Bytes32 *y) {
    switch (w->op) {
        case OP_COMMIT: return blob_to_kzg_commitment(out, &w->blobs[b], w->s);
        case OP_PROOF: return compute_kzg_proof(out, y, &w->blobs[b], &w->z, w->s);
        default: return compute_blob_kzg_proof(out, &w->blobs[b], &w->commit[b], w->s);
    }
}

static void *worker_main(void *arg) {
    Worker *w = (Worker *)arg;
    Cell *cells = w->op == OP_CELLS ? (Cell *)malloc(128 * sizeof(Cell)) : NULL;
    KZGProof *cproofs = w->op == OP_CELLS ? (KZGProof *)malloc(128 * sizeof(KZGProof)) : NULL;
    /* one untimed concurrent round first: the first multi-request batches allocate staging and load kernels */
    pthread_barrier_wait(w->warm);
    {
        Bytes48 out;
        Bytes32 y;
        bool ok;
        if (w->op == OP_CELLS) compute_cells_and_kzg_proofs(cells, cproofs, &w->blobs[0], w->s);
        else if (w->op == OP_VERIFY) verify_blob_kzg_proof(&ok, &w->blobs[0], &w->commit[0], &w->proof[0], w->s);
        else run_one(w, 0, &out, &y);
    }
    pthread_barrier_wait(w->warm);             /* every warm call has returned: the main thread snapshots the counters */
    pthread_barrier_wait(w->start);
    for (int i = 0; i < w->calls; i++) {
        int b = i % w->nblobs;
        if (w->op == OP_VERIFY) {
            bool ok = false;
            if (verify_blob_kzg_proof(&ok, &w->blobs[b], &w->commit[b], &w->proof[b], w->s) != C_KZG_OK) { w->errors++; continue; }
            if (!ok) w->mismatches++;
            continue;
        }
        if (w->op == OP_CELLS) {
            memset(cproofs, 0, 128 * sizeof(KZGProof));
            if (compute_cells_and_kzg_proofs(cells, cproofs, &w->blobs[b], w->s) != C_KZG_OK) { w->errors++; continue; }
            if (memcmp(cells, w->ref_cells + (size_t)b * 128, 128 * sizeof(Cell)) != 0) w->mismatches++;
            if (memcmp(cproofs, w->ref_cproofs + (size_t)b * 128, 128 * sizeof(KZGProof)) != 0) w->mismatches++;
            continue;
        }
        Bytes48 out;
        Bytes32 y;
        memset(&out, 0, sizeof out);
        memset(&y, 0, sizeof y);
        if (run_one(w, b, &out, &y) != C_KZG_OK) { w->errors++; continue; }
        const Bytes48 *want = w->op == OP_COMMIT ? &w->commit[b] : &w->proof[b];
        if (memcmp(&out, want, 48) != 0) w->mismatches++;
        if (w->op == OP_PROOF && memcmp(&y, &w->y[b], 32) != 0) w->mismatches++;
    }
    free(cells);
    free(cproofs);
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s trusted_setup.txt commit|proof|blob_proof|mixed|cells|verify threads calls_per_thread [distinct_blobs]\n", argv[0]); return 2; }
    const char *opname = argv[2];
    int op = !strcmp(opname, "commit") ? OP_COMMIT : !strcmp(opname, "proof") ? OP_PROOF : !strcmp(opname, "blob_proof") ? OP_BLOB_PROOF
             : !strcmp(opname, "mixed") ? OP_MIXED : !strcmp(opname, "cells") ? OP_CELLS : !strcmp(opname, "verify") ? OP_VERIFY : -1;
    int T = atoi(argv[3]), calls = atoi(argv[4]), nblobs = argc > 5 ? atoi(argv[5]) : 4;
    if (op < 0 || T < 1 || T > 256 || calls < 1 || nblobs < 1) { fprintf(stderr, "bad arguments\n"); return 2; }
    FILE *f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 2; }
    KZGSettings s;
    C_KZG_RET rc = load_trusted_setup_file(&s, f);
    fclose(f);
    if (rc != C_KZG_OK) { fprintf(stderr, "load_trusted_setup_file: %d (no CUDA device?)\n", rc); return 3; }

    Worker *w = (Worker *)calloc((size_t)T, sizeof(Worker));
    pthread_barrier_t start, warm;
    pthread_barrier_init(&start, NULL, (unsigned)T + 1);
    pthread_barrier_init(&warm, NULL, (unsigned)T + 1);
    unsigned x = 2463534242u;
    for (int t = 0; t < T; t++) {
        w[t].id = t; w[t].op = op == OP_MIXED ? t % 3 : op; w[t].calls = calls; w[t].nblobs = nblobs; w[t].s = &s; w[t].start = &start; w[t].warm = &warm;
        w[t].blobs = (Blob *)malloc((size_t)nblobs * sizeof(Blob));
        w[t].commit = (Bytes48 *)calloc((size_t)nblobs, sizeof(Bytes48));
        w[t].proof = (Bytes48 *)calloc((size_t)nblobs, sizeof(Bytes48));
        w[t].y = (Bytes32 *)calloc((size_t)nblobs, sizeof(Bytes32));
        for (int b = 0; b < nblobs; b++)
            for (size_t i = 0; i < sizeof(Blob); i++) {      /* field elements with a zero top byte are canonical */
                x ^= x << 13; x ^= x >> 17; x ^= x << 5;
                w[t].blobs[b].bytes[i] = (i % 32 == 0) ? 0 : (uint8_t)(x >> 11);
            }
        memcpy(w[t].z.bytes, w[t].blobs[0].bytes + 96, 32);
        /* single-threaded reference outputs: nothing else is in flight, so each call is a batch of one */
        for (int b = 0; b < nblobs; b++) {
            if (blob_to_kzg_commitment(&w[t].commit[b], &w[t].blobs[b], &s) != C_KZG_OK) { fprintf(stderr, "reference commitment failed\n"); return 4; }
            if (w[t].op == OP_CELLS) continue;
            if (w[t].op == OP_VERIFY) {         /* the proof to verify: compute_blob_kzg_proof of the same blob */
                if (compute_blob_kzg_proof(&w[t].proof[b], &w[t].blobs[b], &w[t].commit[b], &s) != C_KZG_OK) { fprintf(stderr, "reference proof failed\n"); return 4; }
                continue;
            }
            if (w[t].op != OP_COMMIT && run_one(&w[t], b, &w[t].proof[b], &w[t].y[b]) != C_KZG_OK) { fprintf(stderr, "reference proof failed\n"); return 4; }
        }
        if (w[t].op == OP_CELLS) {
            w[t].ref_cells = (Cell *)malloc((size_t)nblobs * 128 * sizeof(Cell));
            w[t].ref_cproofs = (KZGProof *)malloc((size_t)nblobs * 128 * sizeof(KZGProof));
            for (int b = 0; b < nblobs; b++)
                if (compute_cells_and_kzg_proofs(w[t].ref_cells + (size_t)b * 128, w[t].ref_cproofs + (size_t)b * 128, &w[t].blobs[b], &s) != C_KZG_OK) {
                    fprintf(stderr, "reference cells failed\n"); return 4;
                }
        }
    }
    uint64_t st0[5], st[5];
    b200_kzg_coalesce_stats(&s, st0);          /* counters up to here belong to the single-threaded reference pass */
    uint64_t cst0[2];
    b200_kzg_cells_coalesce_stats(&s, cst0);
    uint64_t vst0[3];
    b200_kzg_verify_coalesce_stats(&s, vst0);
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    for (int t = 0; t < T; t++) pthread_create(&th[t], NULL, worker_main, &w[t]);
    pthread_barrier_wait(&warm);               /* the untimed round runs between these two waits */
    pthread_barrier_wait(&warm);
    b200_kzg_coalesce_stats(&s, st0);          /* ... and its batches do not count */
    b200_kzg_cells_coalesce_stats(&s, cst0);
    b200_kzg_verify_coalesce_stats(&s, vst0);
    pthread_barrier_wait(&start);
    double t0 = now_s();
    long mism = 0, errs = 0;
    for (int t = 0; t < T; t++) { pthread_join(th[t], NULL); mism += w[t].mismatches; errs += w[t].errors; }
    double dt = now_s() - t0;
    b200_kzg_coalesce_stats(&s, st);
    for (int i = 0; i < 4; i++) st[i] -= st0[i];
    if (op == OP_VERIFY) {
        uint64_t vs[3];
        b200_kzg_verify_coalesce_stats(&s, vs);
        st[0] = vs[0] - vst0[0]; st[1] = vs[1] - vst0[1]; st[2] = st[3] = st[4] = 0;
    }
    if (op == OP_CELLS) {                       /* its own queue: batches and requests only */
        uint64_t cs[2];
        b200_kzg_cells_coalesce_stats(&s, cs);
        st[0] = cs[0] - cst0[0]; st[1] = cs[1] - cst0[1]; st[2] = st[3] = st[4] = 0;
    }
    /* an invalid blob among valid concurrent callers must fail alone (per-request status, not per-batch) */
    long isolation_failures = 0;
    {
        Blob *bad = (Blob *)malloc(sizeof(Blob));
        memcpy(bad, &w[0].blobs[0], sizeof(Blob));
        bad->bytes[0] = 0xff;
        Bytes48 out;
        if (blob_to_kzg_commitment(&out, bad, &s) != C_KZG_BADARGS) isolation_failures++;
        free(bad);
    }
    printf("{\"op\": \"%s\", \"threads\": %d, \"calls\": %ld, \"seconds\": %.6f, \"per_s\": %.1f, \"mismatches\": %ld, \"errors\": %ld, "
           "\"isolation_failures\": %ld, \"input_memory\": \"pageable (malloc)\", \"batches\": %llu, \"mean_batch\": %.2f, "
           "\"max_batch\": %llu, \"mean_lane_wait_us\": %.1f, \"mean_lane_exec_us\": %.1f}\n",
           opname, T, (long)T * calls, dt, (double)T * calls / dt, mism, errs, isolation_failures, (unsigned long long)st[0],
           st[0] ? (double)st[1] / (double)st[0] : 0.0, (unsigned long long)st[4], st[0] ? 1e-3 * (double)st[2] / (double)st[0] : 0.0,
           st[0] ? 1e-3 * (double)st[3] / (double)st[0] : 0.0);
    free_trusted_setup(&s);
    return (mism || errs || isolation_failures) ? 1 : 0;
}
