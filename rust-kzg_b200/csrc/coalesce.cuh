// coalesce.cuh -- packing concurrent single-item calls into one launch sequence (host side, no device code).
//
// The reference's consumers call the SINGLE-item functions from many threads: rayon par_chunks over blobs
// (kzg/src/eip_4844.rs:770-816) into blob_to_kzg_commitment / compute_blob_kzg_proof (blst/src/eip_4844.rs:163-175,
// 274-291, 476-496), and g1_lincomb on a Send + Sync precomputation handle (kzg/src/msm/sppark.rs:24-44).  One 4096-term
// item cannot fill this GPU (one blob: 0.56 ms, latency-bound; 64 blobs: 2.9 ms), so concurrent calls are combined:
//
//   claim    a caller takes a slot of the OPEN batch of its kind (or opens one and becomes its leader),
//   stage    copies its own inputs into the batch's pinned staging -- every caller its own, in parallel --,
//   run      the leader waits for the device resource (a lane / the handle), closes the batch, runs it, publishes,
//   consume  every caller picks up its own result and status; the last one returns the batch to the pool.
//
// While the device is busy the next batch fills; no timer is involved, and an uncontended call runs immediately as a
// batch of one.  Batches are pooled (pinned staging is allocated once per batch object).
#pragma once
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b200 {

struct CoBatchBase {
    int kind = 0, consumed = 0;
    std::atomic<int> claimed{0};   // written under the queue's mutex; read without it by a lingering leader
    std::atomic<int> ready{0};
    bool done = false;
    int rc = 0;   // failure of the whole batch (device error); per-item validity travels in the payload
    std::condition_variable cv;
};

// Batch: derives from CoBatchBase and owns the staging.  KINDS: independent open batches (one per entry point).
template <class Batch, int KINDS>
struct CoQueue {
    std::mutex mu;
    std::condition_variable cv_free;
    std::vector<std::unique_ptr<Batch>> all;
    std::vector<Batch*> free_list;
    Batch* open[KINDS] = {};
    size_t max_batches = 4;

    struct Claim {
        Batch* b;
        int idx;
        bool leader;
    };
    // make(): allocate a new Batch (may throw; nothing is left half-done)
    template <class Make>
    Claim claim(int kind, int cap, Make&& make) {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            Batch* b = open[kind];
            if (b && b->claimed.load() < cap) return Claim{b, b->claimed.fetch_add(1), false};
            if (b) open[kind] = nullptr;   // full: its leader will run it; start the next one
            if (!free_list.empty()) {
                b = free_list.back();
                free_list.pop_back();
            } else if (all.size() < max_batches) {
                std::unique_ptr<Batch> nb = make();
                b = nb.get();
                all.push_back(std::move(nb));
            } else {
                cv_free.wait(lk);
                continue;
            }
            b->kind = kind; b->claimed.store(1); b->consumed = 0; b->ready.store(0); b->done = false; b->rc = 0;
            open[kind] = b;
            return Claim{b, 0, true};
        }
    }
    static void staged(Batch* b) { b->ready.fetch_add(1, std::memory_order_release); }
    // leader, before close(): how many slots have been claimed so far
    static int claimed_so_far(Batch* b) { return b->claimed.load(std::memory_order_relaxed); }   // no lock: the poll must not starve claimers
    // leader: no more claims; returns the item count once every claimed slot has been staged
    int close(Batch* b) {
        int n;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (open[b->kind] == b) open[b->kind] = nullptr;
            n = b->claimed.load();
        }
        while (b->ready.load(std::memory_order_acquire) < n) std::this_thread::yield();
        return n;
    }
    // leader: publish the outcome (also the only exit when the device resource could not be taken)
    void publish(Batch* b, int rc) {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (open[b->kind] == b) open[b->kind] = nullptr;
            b->rc = rc;
            b->done = true;
        }
        b->cv.notify_all();
    }
    void wait(Batch* b) {
        std::unique_lock<std::mutex> lk(mu);
        b->cv.wait(lk, [&] { return b->done; });
    }
    // every caller, after reading its slot (claimed is final once done is set)
    void consume(Batch* b) {
        std::lock_guard<std::mutex> lk(mu);
        if (++b->consumed == b->claimed.load()) {
            free_list.push_back(b);
            cv_free.notify_one();
        }
    }
};

}  // namespace b200
