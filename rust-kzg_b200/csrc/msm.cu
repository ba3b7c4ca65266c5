// msm.cu -- bucket-method MSM over BLS12-381 G1, hand-written for sm_100a.  See msm.cuh for the pipeline.
#include "msm.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "g1.cuh"
#include "g1_quad.cuh"
#include "warp_inverse.cuh"
#include "util.cuh"

namespace b200 {

static constexpr int kMinTaskLen = 4;     // shortest accumulate task (entries) for latency-bound calls
static constexpr int kAccThreads = 128;  // accumulate CTA: 4 warps, one per SM sub-partition
static constexpr int kReduceBits = 15;   // bucket-index bits the marginal-sum reduce handles (three 5-bit digits)
static constexpr int kMaxWindow = 22;    // wider windows than kReduceBits + 1 go through k_segment_fold first

// ---------------------------------------------------------------------------------------------------------------
// 1 / 3: signed digits of every scalar; histogram (SCATTER = false) or counting-sort scatter (SCATTER = true).
// Scalars are read with two 128-bit loads per thread, consecutive threads read consecutive scalars (coalesced).
// The digit of window j is  raw = bits[o_j, o_j+cw) + carry  (o_j = c0 + c*(j-1), cw = c; window 0: o = 0, cw = c0 <= c);
// raw > 2^(cw-1) becomes raw - 2^cw with carry 1, so |digit| <= 2^(c-1): bucket |digit| - 1 of 2^(c-1) buckets; windows
// reaching bit 256 (c0 + c*(W-1) >= 256, s < r < 0.91 * 2^255) leave no final carry.  Same digit set as the reference's
// Booth recoding (kzg/src/msm/pippenger_utils.rs:251-281): sum_j digit_j 2^(o_j) = s (tests/test_msm_plan_cpu.py).
// ---------------------------------------------------------------------------------------------------------------
// Scalar randomisation for fixed-base tables (MsmConfig::randomize).  The table is built over Q_i = rho_i^-1 * P_i and every
// scalar is multiplied by rho_i before it is cut into digits:  sum (s_i rho_i) Q_i = sum s_i P_i  for points of the
// prime-order subgroup (checked at prepare: k_check_subgroup).  rho_i is pseudo-random and different for every base, so
// whatever the caller's scalars look like -- all equal (the all-0x02 consensus blob), small, r - 1, zero top bytes -- the
// DIGITS are uniform: no bucket ever collects a whole window, any window width is safe, and the timing no longer depends
// on the input (c = 18 / 19 / 21 cost 2.3x because of top-window collisions; all-equal scalars 1.6 - 2.3x).
// The multiplication replaces the Montgomery conversion the kernel does anyway, so it is free.  Zero scalars stay zero.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// canonical rho_i, 0 < rho_i < 2^254 < r
__device__ __forceinline__ fr_t rho_of(uint64_t seed, uint64_t index) {
    fr_t r;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint64_t w = splitmix64(seed ^ (index * 4 + j) * 0xd1342543de82ef95ull);
        if (j == 3) w &= 0x3fffffffffffffffull;
        r.v[2 * j] = (uint32_t)w;
        r.v[2 * j + 1] = (uint32_t)(w >> 32);
    }
    if (r.is_zero()) r.v[0] = 1;
    return r;
}

template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const uint4* __restrict__ scalars, size_t n, size_t row_stride, size_t total,
                                                int c, int c0, int W, int nb, int fixed, int mont, uint32_t* __restrict__ ctr,
                                                uint32_t* __restrict__ entries, size_t period, size_t period_n, size_t gid0,
                                                uint64_t rho_seed) {
    size_t gid = gid0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t vec = gid / n, i = gid - vec * n;
    fr_t s = load_field_ro<fr_t>(scalars + 2 * gid);
    if (rho_seed) {
        // canonical (s rho_i): Montgomery product of (s R) and rho, or of s and (rho R)
        fr_t rho = rho_of(rho_seed, (uint64_t)((vec % period) * period_n + i));
        s = mont ? s * rho : s * rho.to_mont();
    } else if (mont) {
        s = s.from_mont();
    }
    uint32_t w[9];
#pragma unroll
    for (int k = 0; k < 8; k++) w[k] = s.v[k];
    w[8] = 0;
    uint32_t carry = 0;
    for (int j = 0; j < W; j++) {
        // window j: cw bits from bit o (window 0 may be narrower than the others, see MsmConfig::c0)
        const int o = j ? c0 + c * (j - 1) : 0, cw = j ? c : c0;
        const uint32_t mask = (1u << cw) - 1;
        uint32_t raw = 0;
        if (o < 256) {
            int word = o >> 5, sh = o & 31;
            uint64_t two = ((uint64_t)w[word + 1] << 32) | w[word];
            raw = (uint32_t)(two >> sh) & mask;
        }
        raw += carry;
        uint32_t neg = raw > (1u << (cw - 1));
        uint32_t mag = neg ? (1u << cw) - raw : raw;
        carry = neg;
        if (mag != 0) {
            size_t group = fixed ? vec : vec * W + j;
            size_t key = group * nb + (mag - 1);
            // top window: it may hold only a few scalar bits (c = 13: bit 247 and up, and blob elements are < 2^248),
            // so thousands of neighbouring scalars hit the same one or two keys -- lanes of a warp with the same key
            // share ONE atomic there (same-address atomics serialise in L2)
            uint32_t cnt = 1, rank = 0;
            bool leader = true;
            unsigned peers = 0;
            const bool agg = j == W - 1;
            if (agg) {
                peers = __match_any_sync(__activemask(), key);
                cnt = __popc(peers);
                rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
                leader = rank == 0;
            }
            if (SCATTER) {
                uint32_t pos = 0;
                if (leader) pos = atomicAdd(&ctr[key], cnt);
                if (agg) pos = __shfl_sync(peers, pos, __ffs(peers) - 1) + rank;
                uint32_t idx = (uint32_t)(fixed ? (size_t)j * row_stride + (vec % period) * period_n + i : i);
                entries[pos] = idx | (neg << 31);
            } else {
                if (leader) atomicAdd(&ctr[key], cnt);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2: exclusive scan of u32 (three small kernels; n up to ~2^21 keys).  out[n] = total.
static constexpr int kScanBlock = 1024;
static constexpr int kScanItems = 4;  // per thread -> 4096 per CTA

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    uint32_t base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[31];
    __syncthreads();
    return base + x - v;
}
// f: 0 = identity; 0 < f < 2^30: ceil(x / f) (task counts of a bucket with x entries);
// f = kScanSegments | S: `in` holds bucket OFFSETS and the value is the number of chunks of the global S-entry grid the
// bucket [in[i], in[i+1]) touches -- its segment count in the batch-affine accumulation (k_accumulate_affine)
static constexpr uint32_t kScanSegments = 0x40000000u;
__device__ __forceinline__ uint32_t scan_map(const uint32_t* __restrict__ in, size_t i, uint32_t f) {
    const uint32_t x = in[i];
    if (f == 0) return x;
    if (f & kScanSegments) {
        const uint32_t S = f & (kScanSegments - 1), x1 = in[i + 1];
        return x1 > x ? (x1 - 1) / S - x / S + 1 : 0;
    }
    return (x + f - 1) / f;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_partial(const uint32_t* __restrict__ in, size_t n, uint32_t f,
                                                             uint32_t* __restrict__ block_sums) {
    size_t base = ((size_t)blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) s += scan_map(in, base + k, f);
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_blocks(uint32_t* block_sums, int nblocks) {
    // single CTA, nblocks <= kScanBlock * kScanItems
    uint32_t v[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        int idx = threadIdx.x * kScanItems + k;
        v[k] = idx < nblocks ? block_sums[idx] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        int idx = threadIdx.x * kScanItems + k;
        if (idx < nblocks) block_sums[idx] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) block_sums[nblocks] = total;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_final(const uint32_t* __restrict__ in, size_t n, uint32_t f,
                                                           const uint32_t* __restrict__ block_sums,
                                                           uint32_t* __restrict__ out, uint32_t* __restrict__ out2) {
    size_t base = ((size_t)blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
    uint32_t v[kScanItems], s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = base + k < n ? scan_map(in, base + k, f) : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) {
            out[base + k] = ex;
            if (out2) out2[base + k] = ex;
        }
        ex += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanBlock - 1) {
        out[n] = block_sums[gridDim.x];
        if (out2) out2[n] = block_sums[gridDim.x];
    }
}
// out[i] = exclusive prefix of f(in[i]) for i < n; out[n] = total; out2 (optional) gets a copy.
static int scan_exclusive(const uint32_t* in, size_t n, uint32_t f, uint32_t* out, uint32_t* out2, uint32_t* tmp,
                          cudaStream_t st) {
    int nblocks = (int)div_up(n, (size_t)kScanBlock * kScanItems);
    if (nblocks > kScanBlock * kScanItems) throw CudaError(-1, "scan_exclusive: too many keys");
    k_scan_partial<<<nblocks, kScanBlock, 0, st>>>(in, n, f, tmp);
    k_scan_blocks<<<1, kScanBlock, 0, st>>>(tmp, nblocks);
    k_scan_final<<<nblocks, kScanBlock, 0, st>>>(in, n, f, tmp, out, out2);
    B200_LAUNCH_CHECK();
    return 3;
}

// ---------------------------------------------------------------------------------------------------------------
// 4: tasks.  Bucket `key` with cnt entries becomes ceil(cnt/L) tasks; tasks are counting-sorted by length
// (descending) so the 32 lanes of a warp run the same trip count and long tasks start first.
// size_hist layout: [0..L] histogram, [L+1..2L+1] base, [2L+2..3L+2] cursor.
__global__ void __launch_bounds__(256) k_task_hist(const uint32_t* __restrict__ counts, size_t nkeys, int L,
                                                   uint32_t* __restrict__ size_hist) {
    extern __shared__ uint32_t sh[];
    for (int k = threadIdx.x; k <= L; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key < nkeys) {
        uint32_t cnt = counts[key];
        if (cnt) {
            uint32_t tc = (cnt + L - 1) / L, rem = cnt - (tc - 1) * L;
            atomicAdd(&sh[0], 1u);   // slot 0 (no task has length 0) counts the non-empty buckets: MsmEngine::last_stats
            atomicAdd(&sh[rem], 1u);
            if (tc > 1) atomicAdd(&sh[L], tc - 1);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k <= L; k += blockDim.x)
        if (sh[k]) atomicAdd(&size_hist[k], sh[k]);
}
__global__ void k_task_bases(uint32_t* size_hist, int L) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t run = 0;
        for (int s = L; s >= 1; s--) {
            size_hist[L + 1 + s] = run;
            size_hist[2 * L + 2 + s] = 0;
            run += size_hist[s];
        }
    }
}
__global__ void __launch_bounds__(256) k_task_emit(const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                   const uint32_t* __restrict__ task_base, size_t nkeys, int L,
                                                   uint32_t* __restrict__ size_hist, uint32_t* __restrict__ sorted_tasks) {
    size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cnt = key < nkeys ? counts[key] : 0;
    const bool live = cnt != 0;
    uint32_t tc = 0, rem = 0xffffffffu, start = 0, slot = 0;
    if (live) {
        tc = (cnt + L - 1) / L;
        rem = cnt - (tc - 1) * L;
        start = offsets[key];
        slot = task_base[key];
    }
    const uint32_t* base = size_hist + L + 1;
    uint32_t* cur = size_hist + 2 * L + 2;
    if (tc > 1) {
        uint32_t p = base[L] + atomicAdd(&cur[L], tc - 1);
        for (uint32_t q = 0; q + 1 < tc; q++) {
            sorted_tasks[3 * (size_t)(p + q)] = start + q * L;
            sorted_tasks[3 * (size_t)(p + q) + 1] = L;
            sorted_tasks[3 * (size_t)(p + q) + 2] = slot + q;
        }
    }
    // the cursor of a length class is a hot address (2^19 buckets share ~40 classes at c = 20): lanes of a warp with the
    // same class take their slots with ONE atomic (whole warps reach this point: the grid is a multiple of 32 threads)
    const unsigned peers = __match_any_sync(0xffffffffu, rem);
    if (live) {
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        uint32_t first = 0;
        if (lane == leader) first = atomicAdd(&cur[rem], (uint32_t)__popc(peers));
        first = __shfl_sync(peers, first, leader);
        uint32_t p = base[rem] + first + __popc(peers & ((1u << lane) - 1));
        sorted_tasks[3 * (size_t)p] = start + (tc - 1) * L;
        sorted_tasks[3 * (size_t)p + 1] = rem;
        sorted_tasks[3 * (size_t)p + 2] = slot + tc - 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 5: bucket accumulation -- the dominant kernel.  One thread per task; the running XYZZ sum lives in registers, each
// step gathers one 96-byte affine point (six 128-bit read-only loads) and does a mixed addition (8M + 2S).
// MINB = CTAs per SM the register allocation is held to: 3 (<= 168 registers) or 4 (<= 128)
template <bool PREFETCH, int MINB>
__global__ void __launch_bounds__(kAccThreads, MINB) k_accumulate(const uint8_t* __restrict__ table, const uint32_t stride,
                                                            const uint32_t* __restrict__ entries,
                                                            const uint32_t* __restrict__ sorted_tasks,
                                                            const uint32_t* __restrict__ n_tasks_ptr,
                                                            uint8_t* __restrict__ partials) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_tasks_ptr) return;
    uint32_t start = sorted_tasks[3 * t], len = sorted_tasks[3 * t + 1], slot = sorted_tasks[3 * t + 2];
    xyzz_t acc = xyzz_t::inf();
    const uint32_t* e = entries + start;
    if (PREFETCH) {
        // software pipeline: the gather of point k+1 (index load, then six 128-bit loads) is in flight during the
        // ~4000-instruction addition of point k
        uint32_t v = e[0];
        affine_t p = load_affine(table + (size_t)(v & 0x7fffffffu) * stride);
        for (uint32_t k = 0; k < len; k++) {
            affine_t cur = p;
            uint32_t cv = v;
            if (k + 1 < len) {
                v = e[k + 1];
                p = load_affine(table + (size_t)(v & 0x7fffffffu) * stride);
            }
            cur.y = cur.y.cneg(cv >> 31);
            xyzz_add_affine(acc, cur);
        }
    } else {
        for (uint32_t k = 0; k < len; k++) {
            uint32_t v = e[k];
            affine_t p = load_affine(table + (size_t)(v & 0x7fffffffu) * stride);
            p.y = p.y.cneg(v >> 31);
            xyzz_add_affine(acc, p);
        }
    }
    store_xyzz(partials + (size_t)slot * 192, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// 5b: batch-affine bucket accumulation (the reference's own fastest CPU tables work the same way: WbitsTable's
// batch-affine tree, kzg/src/msm/wbits.rs:442-488, and arkmsm's BatchAdder, kzg/src/msm/arkmsm/batch_adder.rs).
//
// A mixed XYZZ addition costs 8M + 2S = 10 field multiplications.  Adding two AFFINE points costs 3 (lambda = dy/dx,
// x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1) once 1/dx is known, and Montgomery's trick gives 1/dx for a
// whole batch at 3 multiplications each plus ONE inversion: 6 per addition.  What makes it pay on this machine is the
// batch: one CTA inverts 128 lanes x K slots = 1280 denominators at a time, so the binary-Euclid inversion (ALU pipe)
// and the product tree that shares it (~5 multiplications per lane and step) are a few per cent of the work.
//
// Work decomposition: the sorted entry list is cut into CHUNKS of S consecutive entries -- a global grid, independent
// of the bucket boundaries -- and every thread owns K chunks ("slots"; chunk c = k * G + g for thread g of G).  In step
// t every slot consumes entry t of its chunk: the first entry of a bucket (or of the chunk) starts a new affine
// accumulator, every other entry is added to it.  All slots of all threads advance in lock step, so the whole machine
// does the same amount of work (no tail, whatever the bucket sizes are) and every step is one batch inversion per CTA.
// A bucket that straddles chunk boundaries leaves one partial sum per chunk it touches ("segment"); the partial slots
// are numbered bucket-major (task_base, from the segment-count scan), so the existing k_bucket_combine folds them and
// the reduce stages are unchanged.  The running sums live in an L2-resident scratch array (96 B per slot), the prefix
// products of Montgomery's trick in shared memory.
//
// Exceptional cases, all explicit: infinity in the table or as the running sum (after P + (-P)), P + P (tangent: the
// denominator becomes 2y and the numerator 3x^2), P + (-P) (infinity; no denominator).
static constexpr int kBaThreads = 128;
static constexpr uint32_t kEntryStart = 1u << 30;           // entry flag: first entry of its bucket
static constexpr uint32_t kEntryIndex = kEntryStart - 1;    // table index bits (sign stays in bit 31)
enum { BA_NOP = 0, BA_LOAD = 1, BA_ADD = 2, BA_SKIP = 3, BA_COPY = 4, BA_DBL = 5, BA_INF = 6 };

// after the scatter: flag the first entry of every non-empty bucket; count the non-empty buckets (size_hist[0])
__global__ void __launch_bounds__(256) k_mark_starts(const uint32_t* __restrict__ offsets, size_t nkeys, uint32_t* __restrict__ entries,
                                                     uint32_t* __restrict__ nonempty) {
    size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    if (key < nkeys) {
        uint32_t o = offsets[key];
        live = offsets[key + 1] > o;
        if (live) entries[o] |= kEntryStart;
    }
    unsigned m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(nonempty, (uint32_t)__popc(m));
}
// partial slot of the first segment of every chunk: chunk c starts at entry c * S inside bucket `key` (the last key
// whose offset is <= c * S), which began floor(offsets[key] / S) chunks earlier
__global__ void __launch_bounds__(256) k_chunk_first(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ task_base,
                                                     size_t nkeys, uint32_t S, size_t nchunks, uint32_t* __restrict__ first_slot) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const size_t p = c * S;
    if (p >= offsets[nkeys]) return;
    size_t lo = 0, hi = nkeys;                      // offsets[lo] <= p < offsets[hi]
    while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) lo = mid; else hi = mid;
    }
    first_slot[c] = task_base[lo] + (uint32_t)(c - offsets[lo] / S);
}

static __device__ __noinline__ fp_t ba_mul(fp_t a, fp_t b) { return a * b; }

// 1 / v for the 128 values of a CTA with ONE field inversion: product tree over the lanes (pairs by shuffle, the 127
// internal nodes in shared memory, each level computed by as few warps as it has nodes), inversion of the root by warp
// 0, then the inverses walk down the tree: inv(child) = inv(parent) * sibling.  ~5 multiplications per lane instead of
// the 12 of two warp-wide scans.  v != 0 on every lane; all 128 threads must call.
// sm_t / sm_i: 128 field elements each (index 1..127 used, heap order: children of i are 2i and 2i+1).
__device__ __forceinline__ fp_t cta_batch_inverse(const fp_t& v, uint8_t* sm_t, uint8_t* sm_i) {
    const int tid = threadIdx.x;
    const fp_t sib = shfl_xor_fp(v, 1);
    {
        fp_t pr = ba_mul(v, sib);                                  // node 64 + tid / 2 (both lanes of a pair compute it)
        if (!(tid & 1)) store_field(sm_t + (size_t)(64 + (tid >> 1)) * 48, pr);
    }
    __syncthreads();
#pragma unroll 1
    for (int size = 32; size >= 1; size >>= 1) {
        if (tid < size) {
            const int i = size + tid;
            fp_t pr = ba_mul(load_field<fp_t>(sm_t + (size_t)(2 * i) * 48), load_field<fp_t>(sm_t + (size_t)(2 * i + 1) * 48));
            store_field(sm_t + (size_t)i * 48, pr);
        }
        __syncthreads();
    }
    if (tid < 32) {
        fp_t r = load_field<fp_t>(sm_t + 48).inverse();            // identical on the 32 lanes: no divergence
        if (tid == 0) store_field(sm_i + 48, r);
    }
    __syncthreads();
#pragma unroll 1
    for (int size = 1; size <= 32; size <<= 1) {                   // parents [size, 2 size) -> children [2 size, 4 size)
        if (tid < 2 * size) {
            const int j = 2 * size + tid;
            fp_t r = ba_mul(load_field<fp_t>(sm_i + (size_t)(j >> 1) * 48), load_field<fp_t>(sm_t + (size_t)(j ^ 1) * 48));
            store_field(sm_i + (size_t)j * 48, r);
        }
        __syncthreads();
    }
    return ba_mul(load_field<fp_t>(sm_i + (size_t)(64 + (tid >> 1)) * 48), sib);
}

// TREE = false (default): every WARP inverts its own 32 x K denominators (two shuffle scans + one inversion, identical on
// all lanes: warp_inverse.cuh) -- no barrier anywhere, so the warps of an SM drift out of phase and one warp's inversion
// (ALU pipe) runs under the other warps' multiplications (FMA-heavy pipe).  TREE = true: one inversion per CTA through the
// shared-memory product tree above (fewer multiplications, but every step ends in a serial, barrier-separated phase).
template <int K, bool TREE>
__global__ void __launch_bounds__(kBaThreads, 3) k_accumulate_affine(const uint8_t* __restrict__ table, const uint32_t stride,
                                                                   const uint32_t* __restrict__ entries,
                                                                   const uint32_t* __restrict__ n_entries_ptr,
                                                                   const uint32_t* __restrict__ first_slot, const uint32_t S,
                                                                   uint8_t* __restrict__ acc_buf, uint8_t* __restrict__ partials) {
    extern __shared__ __align__(16) uint8_t ba_smem[];
    uint4* sm_prefix = reinterpret_cast<uint4*>(ba_smem);                  // [K][3][128] : conflict-free 16-byte columns
    uint8_t* sm_t = ba_smem + (size_t)K * 3 * kBaThreads * 16;             // product tree
    uint8_t* sm_i = sm_t + 128 * 48;                                       // inverses of the tree nodes
    uint32_t* sm_next = reinterpret_cast<uint32_t*>(sm_i + 128 * 48);      // [K][128] next partial slot of every chunk
    const uint32_t tid = threadIdx.x;
    const size_t G = (size_t)gridDim.x * kBaThreads, g = (size_t)blockIdx.x * kBaThreads + tid;
    const size_t E = *n_entries_ptr;
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const size_t c = (size_t)k * G + g;
        sm_next[k * kBaThreads + tid] = c * S < E ? first_slot[c] : 0u;
    }
#pragma unroll 1
    for (uint32_t step = 0; step < S; step++) {
        uint64_t modes = 0;
        fp_t running = fp_t::one();
        // ---- forward: denominators and their prefix products ------------------------------------------------------
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const size_t c = (size_t)k * G + g, pos = c * S + step;
            if (pos >= E) continue;
            const uint32_t e = entries[pos];
            uint32_t mode = BA_LOAD;
            if (step != 0 && !(e & kEntryStart)) {
                const uint8_t* pp = table + (size_t)(e & kEntryIndex) * stride;
                const uint8_t* ap = acc_buf + c * 96;
                const fp_t x2 = load_field_ro<fp_t>(pp), x1 = load_field<fp_t>(ap);
                mode = BA_ADD;
                const bool z1 = x1.is_zero(), z2 = x2.is_zero();
                if (z1 || z2) {                       // possibly an infinity operand (x = 0 is also a valid abscissa)
                    if (z2 && load_field_ro<fp_t>(pp + 48).is_zero()) mode = BA_SKIP;
                    else if (z1 && load_field<fp_t>(ap + 48).is_zero()) mode = BA_COPY;
                }
                if (mode == BA_ADD) {
                    fp_t d = x2 - x1;
                    if (d.is_zero()) {                // same abscissa: P + P or P + (-P)
                        const fp_t y1 = load_field<fp_t>(ap + 48), y2 = load_field_ro<fp_t>(pp + 48).cneg(e >> 31);
                        if (y1 == y2 && !y1.is_zero()) { mode = BA_DBL; d = y1.dbl(); } else mode = BA_INF;
                    }
                    if (mode != BA_INF) {
                        uint4* pf = sm_prefix + (size_t)k * 3 * kBaThreads + tid;
                        pf[0] = make_uint4(running.v[0], running.v[1], running.v[2], running.v[3]);
                        pf[kBaThreads] = make_uint4(running.v[4], running.v[5], running.v[6], running.v[7]);
                        pf[2 * kBaThreads] = make_uint4(running.v[8], running.v[9], running.v[10], running.v[11]);
                        running = running * d;
                    }
                }
            }
            modes |= (uint64_t)mode << (3 * k);
        }
        // ---- one inversion for the 128 x K denominators of the CTA (step 0 only loads) ------------------------------
        fp_t inv = fp_t::one();
        if (step != 0) inv = TREE ? cta_batch_inverse(running, sm_t, sm_i) : warp_inverse(running);
        // ---- backward: peel the inverses off and finish the additions ---------------------------------------------
#pragma unroll 1
        for (int k = K - 1; k >= 0; k--) {
            const uint32_t mode = (uint32_t)(modes >> (3 * k)) & 7u;
            if (mode == BA_NOP) continue;
            const size_t c = (size_t)k * G + g, pos = c * S + step;
            const uint32_t e = entries[pos];
            const uint8_t* pp = table + (size_t)(e & kEntryIndex) * stride;
            uint8_t* ap = acc_buf + c * 96;
            affine_t res;
            if (mode == BA_LOAD || mode == BA_COPY) {
                res = load_affine(pp);
                res.y = res.y.cneg(e >> 31);
            } else if (mode == BA_SKIP) {
                res = load_affine(ap);
            } else if (mode == BA_INF) {
                res.x = fp_t::zero();
                res.y = fp_t::zero();
            } else {
                const affine_t a = load_affine(ap);
                affine_t b = load_affine(pp);
                b.y = b.y.cneg(e >> 31);
                const uint4* pf = sm_prefix + (size_t)k * 3 * kBaThreads + tid;
                fp_t pre;
                {
                    uint4 t0 = pf[0], t1 = pf[kBaThreads], t2 = pf[2 * kBaThreads];
                    pre.v[0] = t0.x; pre.v[1] = t0.y; pre.v[2] = t0.z; pre.v[3] = t0.w;
                    pre.v[4] = t1.x; pre.v[5] = t1.y; pre.v[6] = t1.z; pre.v[7] = t1.w;
                    pre.v[8] = t2.x; pre.v[9] = t2.y; pre.v[10] = t2.z; pre.v[11] = t2.w;
                }
                fp_t d, num;
                if (mode == BA_DBL) {
                    d = a.y.dbl();
                    fp_t xx = ba_mul(a.x, a.x);
                    num = xx.dbl() + xx;              // 3 x^2
                } else {
                    d = b.x - a.x;
                    num = b.y - a.y;
                }
                const fp_t inv_d = pre * inv;         // 1 / d
                inv = inv * d;
                const fp_t lam = num * inv_d;
                res.x = lam.sqr() - a.x - b.x;
                res.y = lam * (a.x - res.x) - a.y;
            }
            // the segment ends with this entry when the chunk does or the next entry starts a bucket
            const bool last = step + 1 == S || pos + 1 >= E || (entries[pos + 1] & kEntryStart);
            if (last) {
                const uint32_t slot = sm_next[k * kBaThreads + tid]++;
                xyzz_t out;
                const bool inf = res.is_inf();
                out.x = res.x; out.y = res.y;
                out.zz = inf ? fp_t::zero() : fp_t::one();
                out.zzz = out.zz;
                store_xyzz(partials + (size_t)slot * 192, out);
            } else {
                store_affine(ap, res);
            }
        }
    }
}
// variant on the Karatsuba multiplier (B200_ACC_KARA=1; mont.cuh MODE 5): 252 instead of 288 wide multiplies per field
// multiplication, paid for with ~120 more additions on the otherwise idle ALU pipe
template <int MINB>
__global__ void __launch_bounds__(kAccThreads, MINB) k_accumulate_kara(const uint8_t* __restrict__ table, const uint32_t stride,
                                                                     const uint32_t* __restrict__ entries,
                                                                     const uint32_t* __restrict__ sorted_tasks,
                                                                     const uint32_t* __restrict__ n_tasks_ptr,
                                                                     uint8_t* __restrict__ partials) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_tasks_ptr) return;
    uint32_t start = sorted_tasks[3 * t], len = sorted_tasks[3 * t + 1], slot = sorted_tasks[3 * t + 2];
    ck::xyzz_t acc = ck::xyzz_t::inf();
    const uint32_t* e = entries + start;
    uint32_t v = e[0];
    ck::affine_t p = ck::load_affine(table + (size_t)(v & 0x7fffffffu) * stride);
    for (uint32_t k = 0; k < len; k++) {
        ck::affine_t cur = p;
        uint32_t cv = v;
        if (k + 1 < len) {
            v = e[k + 1];
            p = ck::load_affine(table + (size_t)(v & 0x7fffffffu) * stride);
        }
        cur.y = cur.y.cneg(cv >> 31);
        ck::xyzz_add_affine(acc, cur);
    }
    ck::store_xyzz(partials + (size_t)slot * 192, acc);
}

// variant with the field multiplication out of line (B200_ACC_CALL=1): same work, ~20x smaller loop body
__global__ void __launch_bounds__(kAccThreads) k_accumulate_call(const uint8_t* __restrict__ table, const uint32_t stride,
                                                                 const uint32_t* __restrict__ entries,
                                                                 const uint32_t* __restrict__ sorted_tasks,
                                                                 const uint32_t* __restrict__ n_tasks_ptr,
                                                                 uint8_t* __restrict__ partials) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_tasks_ptr) return;
    uint32_t start = sorted_tasks[3 * t], len = sorted_tasks[3 * t + 1], slot = sorted_tasks[3 * t + 2];
    cl::xyzz_t acc = cl::xyzz_t::inf();
    const uint32_t* e = entries + start;
    for (uint32_t k = 0; k < len; k++) {
        uint32_t v = e[k];
        cl::affine_t p = cl::load_affine(table + (size_t)(v & 0x7fffffffu) * stride);
        p.y = p.y.cneg(v >> 31);
        cl::xyzz_add_affine(acc, p);
    }
    cl::store_xyzz(partials + (size_t)slot * 192, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// 6: bucket reduction  S_g = sum_b (b+1) * B_b  per group.  A running sum over 2^(c-1) buckets is a chain of
// dependent point additions (the reference's p1_integrate_buckets, kzg/src/msm/tiling_pippenger_ops.rs:21-45), and a
// single thread needs several microseconds per addition, so the chain is replaced by a shallow, wide form.
// Write the bucket index in D <= 3 "digits" b = sum_a v_a << sh_a (each digit <= 5 bits).  Then
//        sum_b b * B_b = sum_a 2^(sh_a) * sum_v v * M[a][v],     M[a][v] = sum of the buckets whose digit a equals v,
// i.e. D*32 independent plain sums (tree reductions) followed by D weighted sums over <= 32 items, each done by one
// warp with a suffix scan in registers (sum_v v*M_v = sum_{k>=1} Suf_k).  Depth ~ 30 additions instead of 2^c.

// inclusive suffix scan: lane l gets sum_{m >= l} v_m
__device__ __forceinline__ xyzz_t warp_suffix_scan_xyzz(xyzz_t v, int width) {
    int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int d = 1; d < width; d <<= 1) {   // lanes >= width hold infinity
        xyzz_t o = shfl_down_xyzz(v, d);
        if (lane + d >= 32) o = xyzz_t::inf();
        xyzz_add(v, o);
    }
    return v;
}

// 6a: buckets that were cut into several tasks: fold the task partials into the first slot.
// Buckets with 2..32 partials: one thread per bucket when there are many buckets (throughput-bound), or, SUB = true,
// kCombLanes lanes per bucket when there are few (latency-bound: strided partial sums, then a quad tree -- the chain is
// ceil(tc/kCombLanes) + log2(kCombLanes) additions deep instead of tc - 1).
// WARP = true: a small grid of warps strides over the keys and folds buckets with more than 32 partials.
static constexpr int kCombLanes = 8;
// warp_min: buckets with more partials than this go to the WARP form.  32 when most buckets are cut (the per-thread
// chains are what the machine is filled with); 4 when cut buckets are the exception (skewed digits, e.g. a top window
// with one or two scalar bits: a 32-partial chain on one thread would be 0.4 ms of latency).
template <bool WARP, bool SUB>
__global__ void __launch_bounds__(128) k_bucket_combine(uint8_t* __restrict__ partials, const uint32_t* __restrict__ task_base,
                                                        size_t nkeys, uint32_t warp_min) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (WARP) {
        // rare case (adversarial scalar distributions): each warp scans 32 keys per step with coalesced loads and
        // folds the few buckets that have more than 32 partials
        const int lane = threadIdx.x & 31;
        const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
        for (size_t key0 = (gid >> 5) * 32; key0 < nkeys; key0 += nwarps * 32) {
            size_t key = key0 + lane;
            uint32_t my0 = key < nkeys ? task_base[key] : 0, my1 = key < nkeys ? task_base[key + 1] : 0;
            unsigned big = __ballot_sync(0xffffffffu, my1 - my0 > warp_min);
            while (big) {
                int src = __ffs(big) - 1;
                big &= big - 1;
                uint32_t s0 = __shfl_sync(0xffffffffu, my0, src), s1 = __shfl_sync(0xffffffffu, my1, src);
                xyzz_t acc = xyzz_t::inf();
                for (uint32_t s = s0 + lane; s < s1; s += 32) {
                    xyzz_t part = load_xyzz(partials + (size_t)s * 192);
                    xyzz_add(acc, part);
                }
                fp_t q = seg_sum_quad(acc, 32);
                if (lane < 4) store_field(partials + (size_t)s0 * 192 + quad_store_offset(), q);
            }
        }
        return;
    }
    if (SUB) {
        const size_t key = gid / kCombLanes;
        const int sub = (int)(gid % kCombLanes);
        uint32_t s0 = 0, s1 = 0;
        if (key < nkeys) { s0 = task_base[key]; s1 = task_base[key + 1]; }
        const uint32_t tc = s1 - s0;
        const bool live = tc >= 2 && tc <= warp_min;
        // a warp none of whose buckets was cut has nothing to do
        if (!__any_sync(0xffffffffu, live)) return;
        xyzz_t acc = xyzz_t::inf();
        if (live) {
            for (uint32_t s = s0 + sub; s < s1; s += kCombLanes) {
                xyzz_t part = load_xyzz(partials + (size_t)s * 192);
                xyzz_add(acc, part);
            }
        }
        fp_t q = seg_sum_quad(acc, kCombLanes);
        if (live && sub < 4) store_field(partials + (size_t)s0 * 192 + quad_store_offset(), q);
        return;
    }
    const size_t key = gid;
    if (key >= nkeys) return;
    uint32_t s0 = task_base[key], s1 = task_base[key + 1];
    uint32_t tc = s1 - s0;
    if (tc < 2 || tc > warp_min) return;
    xyzz_t acc = load_xyzz(partials + (size_t)s0 * 192);
    for (uint32_t s = s0 + 1; s < s1; s++) {
        xyzz_t part = load_xyzz(partials + (size_t)s * 192);
        xyzz_add(acc, part);
    }
    store_xyzz(partials + (size_t)s0 * 192, acc);
}

struct AxisPlan {
    int D;
    int w[3];   // digit widths (bits), sum = c - 1
    int sh[3];  // digit bit offsets
};

// 6a': wide windows (c > 16).  The marginal form below handles 15 bucket-index bits; a wider bucket set is first folded
// by SEGMENTS of 2^kf consecutive buckets, one thread per segment (2^15 segments per group: a throughput kernel, the
// chains are 2 * 2^kf additions).  With b = hi * 2^kf + lo and weights b + 1:
//      sum_b (b+1) B_b = 2^kf * sum_hi (hi+1) T_hi  -  sum_hi R_hi,
//      T_hi = sum_lo B_(hi,lo),   R_hi = sum_lo (2^kf - 1 - lo) B_(hi,lo)   (an ascending running sum: acc += run; run += B)
// so the T_hi go through the 15-bit reduce unchanged and the R_hi only need a plain sum.  This is the reference's
// p1_integrate_buckets (kzg/src/msm/tiling_pippenger_ops.rs:21-45) applied per segment instead of over the whole set.
// The additions go through the out-of-line multiplier (namespace cl): the kernel is two point additions in a short loop,
// ~4 KiB of code instead of ~90 KiB -- with one or two warps per scheduler instruction fetch is exposed latency.
__global__ void __launch_bounds__(128) k_segment_fold(const uint8_t* __restrict__ partials, const uint32_t* __restrict__ task_base,
                                                      size_t nseg, int kf, uint8_t* __restrict__ seg_t, uint8_t* __restrict__ seg_r) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const size_t key0 = s << kf;
    cl::xyzz_t run = cl::xyzz_t::inf(), acc = cl::xyzz_t::inf();
#pragma unroll 1
    for (int lo = 0; lo < (1 << kf); lo++) {
        cl::xyzz_add(acc, run);
        uint32_t s0 = task_base[key0 + lo];
        if (task_base[key0 + lo + 1] > s0) {
            cl::xyzz_t part = cl::load_xyzz(partials + (size_t)s0 * 192);
            cl::xyzz_add(run, part);
        }
    }
    cl::store_xyzz(seg_t + s * 192, run);
    cl::store_xyzz(seg_r + s * 192, acc);
}
// identity slot map for point arrays that have exactly one slot per key (the folded segments)
__global__ void k_iota(uint32_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

// 6b: marginal sums.  CTA (v, a, g) adds up the nb >> w[a] buckets of group g whose digit a equals v.
template <int kMargThreads>
__global__ void __launch_bounds__(kMargThreads) k_marginals(const uint8_t* __restrict__ partials,
                                                            const uint32_t* __restrict__ task_base, int nb, AxisPlan ap,
                                                            uint8_t* __restrict__ marg,
                                                            const uint8_t* __restrict__ partials_r, uint8_t* __restrict__ marg_r) {
    __shared__ __align__(16) uint8_t sh[(kMargThreads / 32) * 192];
    const int v = blockIdx.x;
    int a = blockIdx.y;
    const size_t g = blockIdx.z;
    int wa, sa;
    if (a == ap.D) {
        // extra grid row (wide windows only): plain sum of the segment running sums R, as the 32 marginals of one
        // 5-bit digit over the second point array (k_group_finish adds them up)
        partials = partials_r;
        marg = marg_r;
        a = 0; wa = nb >= 32 ? 5 : 0; sa = 0;
    } else {
        wa = ap.w[a]; sa = ap.sh[a];
    }
    uint8_t* dst = marg + ((g * 3 + a) * 32 + v) * 192;
    if (v >= (1 << wa)) {
        if (threadIdx.x == 0) store_xyzz(dst, xyzz_t::inf());
        return;
    }
    const int count = nb >> wa;
    xyzz_t acc = xyzz_t::inf();
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        int b = ((i >> sa) << (sa + wa)) | (v << sa) | (i & ((1 << sa) - 1));
        size_t key = g * nb + b;
        uint32_t s0 = task_base[key];
        if (task_base[key + 1] > s0) {
            xyzz_t part = load_xyzz(partials + (size_t)s0 * 192);
            xyzz_add(acc, part);
        }
    }
    // warp sums (quad-distributed), then the kMargThreads/32 warp results are folded by warp 0, one quad each
    fp_t q = seg_sum_quad(acc, 32);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane < 4) store_field(sh + wid * 192 + quad_store_offset(), q);
    __syncthreads();
    if (wid == 0) {
        constexpr int nw = kMargThreads / 32;
        fp_t a = (lane >> 2) < nw ? load_field<fp_t>(sh + (lane >> 2) * 192 + quad_store_offset()) : fp_t::zero();
        a = quad_tree(a, 4 * nw);
        if (lane < 4) store_field(dst + quad_store_offset(), a);
    }
}

static constexpr int kMargSerial = 8;  // buckets summed serially per lane before the shuffle tree
// 6b': the same marginal sums for many groups (blob batches): S lanes per marginal instead of a CTA, so that the
// tree steps waste few lanes -- with 64 groups x 40 marginals the CTA form spends most of its FMA-pipe time on
// additions with the point at infinity.  S = 2^log_s lanes sum count/S buckets each, then a
// log_s-step shuffle tree inside the S-lane group.
__global__ void __launch_bounds__(128, 3) k_marginals_sub(const uint8_t* __restrict__ partials, const uint32_t* __restrict__ task_base,
                                                       int nb, AxisPlan ap, size_t groups, uint8_t* __restrict__ marg,
                                                       const uint8_t* __restrict__ partials_r, uint8_t* __restrict__ marg_r) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int a = blockIdx.y;  // one grid row per digit axis: the axes are independent and run concurrently
    int wa, sa;
    if (a == ap.D) {
        // extra row after a segment fold: plain sum of the R points as the 32 marginals of one 5-bit digit (see k_marginals)
        partials = partials_r;
        marg = marg_r;
        a = 0; wa = nb >= 32 ? 5 : 0; sa = 0;
    } else {
        wa = ap.w[a]; sa = ap.sh[a];
    }
    int log_s = 0;
    while (log_s < 5 && ((nb >> wa) >> log_s) > kMargSerial) log_s++;
    const int S = 1 << log_s;
    const size_t total_threads = (groups << wa) << log_s;
    const bool live = gid < total_threads;
    size_t mi = (live ? gid : total_threads - 1) >> log_s;  // marginal index = g * 2^wa + v
    int sub = (int)(gid & (S - 1));
    size_t g = mi >> wa;
    int v = (int)(mi & ((1 << wa) - 1));
    const int count = nb >> wa;
    xyzz_t acc = xyzz_t::inf();
    if (live) {
        for (int i = sub; i < count; i += S) {
            int b = ((i >> sa) << (sa + wa)) | (v << sa) | (i & ((1 << sa) - 1));
            size_t key = g * nb + b;
            uint32_t s0 = task_base[key];
            if (task_base[key + 1] > s0) {
                xyzz_t part = load_xyzz(partials + (size_t)s0 * 192);
                xyzz_add(acc, part);
            }
        }
    }
    uint8_t* dst = marg + ((g * 3 + a) * 32 + v) * 192;
    if (S >= 4) {
        fp_t q = seg_sum_quad(acc, S);
        if (live && sub < 4) store_field(dst + quad_store_offset(), q);
    } else {
        if (S == 2) {
            xyzz_t o = shfl_down_xyzz(acc, 1);
            if (sub) o = xyzz_t::inf();
            xyzz_add(acc, o);
        }
        if (live && sub == 0) store_xyzz(dst, acc);
    }
}

// 6c: one CTA per group: weighted sums sum_v v * M_(a,v) over the <= 32 marginals of every digit axis a, scaled by 2^sh[a],
// plus the plain sum (the "+1" of the weights b + 1).  The weights are decomposed into BITS: warp (a, k) sums the marginals
// whose index has bit k set (one masked quad tree, all 15 of them side by side), then one warp per axis runs the five-bit
// Horner sum_k 2^k S_(a,k) and its 2^sh[a] -- 38 + 37 us of dependent additions instead of the 65 + 38 us of a five-level
// suffix scan followed by a tree.  Everything runs on quad-distributed points (g1_quad.cuh).
// marg_r != nullptr (wide windows, see k_segment_fold): warp 16 sums the 32 plain marginals of the R_hi and the result is
// 2^kf * (15-bit reduce of the T_hi) - sum R.
// Launch shape: the 17 sums (15 bit sums, the plain sum, the R sum) are spread over six CTAs of three warps per group, so
// that every warp keeps its full register budget; the CTA that finishes last (a counter per group) runs the Horner stage.
static constexpr int kFinSlots = 17, kFinCtas = 6, kFinThreads = 96;
__global__ void __launch_bounds__(kFinThreads) k_group_finish(const uint8_t* __restrict__ marg, AxisPlan ap,
                                                              uint8_t* __restrict__ group_sums, uint8_t* __restrict__ out_jac,
                                                              const uint8_t* __restrict__ marg_r, int kf, uint8_t* __restrict__ scratch,
                                                              unsigned* __restrict__ counters) {
    __shared__ __align__(16) uint8_t sh[3 * 192];             // the axis results
    __shared__ int sh_last;
    const size_t g = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int w = blockIdx.x * 3 + wid;                       // slot 0 .. 17 (17: idle)
    const int off = quad_store_offset();
    uint8_t* slots = scratch + g * kFinSlots * 192;
    {
        const int a = w < 15 ? w / 5 : 0, k = w % 5;
        const bool bit_warp = w < 15 && a < ap.D && k < ap.w[a];
        const bool plain_warp = w == 15, r_warp = w == 16 && marg_r != nullptr;
        if (bit_warp || plain_warp || r_warp) {
            const uint8_t* src = r_warp ? marg_r + ((g * 3) * 32 + lane) * 192 : marg + ((g * 3 + a) * 32 + lane) * 192;
            xyzz_t m = load_xyzz(src);                         // lanes beyond the axis width read infinity
            if (bit_warp && !((lane >> k) & 1)) m = xyzz_t::inf();
            fp_t q = seg_sum_quad(m, 32);
            if (lane < 4) store_field(slots + w * 192 + off, q);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&counters[g], 1u);
        sh_last = done == gridDim.x - 1;
        if (sh_last) counters[g] = 0;                         // ready for the next run
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    auto slot = [&](int idx) {                                // written by other CTAs: read through L2
        fp_t v;
        const uint4* src = reinterpret_cast<const uint4*>(slots + idx * 192 + off);
#pragma unroll
        for (int q = 0; q < 3; q++) {
            uint4 t = __ldcg(src + q);
            v.v[4 * q] = t.x; v.v[4 * q + 1] = t.y; v.v[4 * q + 2] = t.z; v.v[4 * q + 3] = t.w;
        }
        return v;
    };
    if (wid < ap.D) {
        const int a = wid, top = ap.w[a] - 1;                  // Horner over the bits of the digit, then the digit's position
        fp_t acc = top >= 0 ? slot(a * 5 + top) : fp_t::zero();
        for (int k = top - 1; k >= 0; k--) {
            acc = quad_dbl(acc);
            acc = quad_add(acc, slot(a * 5 + k));
        }
        for (int k = 0; k < ap.sh[a]; k++) acc = quad_dbl(acc);
        if (lane < 4) store_field(sh + a * 192 + off, acc);
    }
    __syncthreads();
    if (wid == 0) {
        fp_t acc = slot(15);                                  // the "+1" of the weights b+1: plain sum of all buckets (axis 0 marginals)
        for (int a = 0; a < ap.D; a++) acc = quad_add(acc, load_field<fp_t>(sh + a * 192 + off));
        if (marg_r) {
            for (int k = 0; k < kf; k++) acc = quad_dbl(acc);
            fp_t r = slot(16);
            if ((lane & 3) == 1) r = r.neg();                 // -(X, Y, ZZ, ZZZ) = (X, -Y, ZZ, ZZZ); infinity stays all-zero
            acc = quad_add(acc, r);
        }
        if (group_sums && lane < 4) store_field(group_sums + g * 192 + off, acc);
        if (out_jac) {
            // Jacobian (X*ZZ, Y*ZZZ, ZZ), see xyzz_to_jac; infinity is all-zero in both forms
            fp_t t = acc * shfl_xor_fp(acc, 2);
            if (lane < 2) store_field(out_jac + g * 144 + lane * 48, t);
            if (lane == 2) store_field(out_jac + g * 144 + 96, acc);
        }
    }
}
// 7 (VARIABLE): result = sum_j 2^(c*j) S_j, Horner from the top window (as tiling_pippenger does,
// kzg/src/msm/tiling_pippenger_ops.rs:106-138).  One warp, quad-distributed: W*c doublings of three levels each.
__global__ void __launch_bounds__(32) k_horner(const uint8_t* __restrict__ group_sums, int W, int c, uint8_t* __restrict__ out_jac) {
    const int lane = threadIdx.x & 31, off = quad_store_offset();
    fp_t acc = load_field<fp_t>(group_sums + (size_t)(W - 1) * 192 + off);
    for (int j = W - 2; j >= 0; j--) {
        for (int k = 0; k < c; k++) acc = quad_dbl(acc);
        acc = quad_add(acc, load_field<fp_t>(group_sums + (size_t)j * 192 + off));
    }
    fp_t t = acc * shfl_xor_fp(acc, 2);
    if (lane < 2) store_field(out_jac + lane * 48, t);
    if (lane == 2) store_field(out_jac + 96, acc);
}

// randomisation, step 1: are all bases in the prime-order subgroup (or infinity)?  One lane quad per point.
__global__ void __launch_bounds__(32) k_check_subgroup(const uint8_t* __restrict__ table, size_t stride, size_t n, int* __restrict__ bad) {
    const size_t q = (size_t)blockIdx.x * 8 + (threadIdx.x >> 2);
    const int role = threadIdx.x & 3;
    const bool live = q < n;
    affine_t a = load_affine(table + (live ? q : 0) * stride);
    const bool inf = a.is_inf();
    fp_t comp = role == 0 ? a.x : role == 1 ? a.y : fp_t::one();
    if (inf) comp = fp_t::zero();
    const bool ok = quad_in_subgroup(a, comp);
    if (role == 0 && live && !inf && !ok) *bad = 1;
}
// randomisation, step 2: P_i <- rho_i^-1 * P_i in place (row 0 of the table), one lane quad per point: GLV scalar
// multiplication on the quad arithmetic, back to affine with one warp-shared field inversion
__global__ void __launch_bounds__(32) k_scale_points(uint8_t* __restrict__ table, size_t stride, size_t n, uint64_t seed) {
    __shared__ __align__(16) uint8_t qtable[kQuadTableBytes];
    const size_t q = (size_t)blockIdx.x * 8 + (threadIdx.x >> 2);
    const int role = threadIdx.x & 3, base = threadIdx.x & 28;
    const bool live = q < n;
    const size_t i = live ? q : 0;
    affine_t a = load_affine(table + i * stride);
    const bool inf = a.is_inf();
    fp_t comp = role == 0 ? a.x : role == 1 ? a.y : fp_t::one();
    if (inf) comp = fp_t::zero();
    // k = rho_i^-1 mod r, canonical (the four lanes of a quad compute the same value)
    fr_t k = warp_inverse(rho_of(seed, (uint64_t)i).to_mont()).from_mont();
    comp = quad_mul_scalar(comp, k.v, qtable);
    // affine: x = X / ZZ, y = Y / ZZZ with 1/ZZ = ZZZ^-2 ZZ^2 (ZZ^3 = ZZZ^2)
    const fp_t zz = shfl_idx_fp(comp, base | 2), zzz = shfl_idx_fp(comp, base | 3);
    const bool rinf = zz.is_zero();
    const fp_t izzz = warp_inverse(rinf ? fp_t::one() : zzz);
    const fp_t izz = izzz.sqr() * zz.sqr();
    fp_t out = role == 0 ? comp * izz : comp * izzz;
    if (rinf) out = fp_t::zero();
    if (live && role < 2) store_field(table + i * stride + role * 48, out);
}

// table rows for FIXED engines: row j = 2^(c*j) * P_i, affine.  One thread per point walks all rows
// (c doublings in XYZZ, then back to affine with one warp-shared field inversion).  One-time cost at prepare.
__global__ void __launch_bounds__(128) k_build_rows(uint8_t* __restrict__ table, size_t n, int W, int c, int c0, size_t stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    if (!live) i = n - 1;
    cc::affine_t p = cc::load_affine(table + i * stride);
    for (int j = 1; j < W; j++) {
        cc::xyzz_t q = cc::affine_to_xyzz(p);
        for (int k = 0; k < (j == 1 ? c0 : c); k++) cc::xyzz_dbl(q);   // row j = 2^(c0 + (j-1) c) * P
        const bool inf = q.is_inf();
        // 1/ZZZ; then 1/ZZ = ZZZ^-2 * ZZ^2  (ZZ^3 = ZZZ^2), as xyzz_to_affine
        cc::fp_t izzz = warp_inverse(inf ? cc::fp_t::one() : q.zzz);
        cc::fp_t izz = izzz.sqr() * q.zz.sqr();
        p = inf ? cc::affine_t{cc::fp_t::zero(), cc::fp_t::zero()} : cc::affine_t{q.x * izz, q.y * izzz};
        if (live) cc::store_affine(table + ((size_t)j * n + i) * stride, p);
    }
}

// Jacobian -> 48-byte compressed (blst_p1_compress): one thread per point, one warp-shared field inversion.
// brp_bits > 0: output index = bit-reversal of the low brp_bits bits of i (reverse_bit_order per group of 2^brp_bits)
__global__ void __launch_bounds__(32) k_compress(const uint8_t* __restrict__ jac, uint8_t* __restrict__ out, int count, int brp_bits) {
    const int lane = threadIdx.x;
    const int i = blockIdx.x * blockDim.x + lane;
    const bool live = i < count;
    int o = i;
    if (brp_bits) {
        int low = i & ((1 << brp_bits) - 1);
        o = (i - low) | (int)(__brev((unsigned)low) >> (32 - brp_bits));
    }
    cc::jac_t p = live ? cc::load_jac(jac + (size_t)i * 144) : cc::jac_t::inf();
    const bool inf = p.is_inf();
    cc::fp_t inv = warp_inverse(inf ? cc::fp_t::one() : p.z);
    if (!live) return;
    cc::affine_t a{cc::fp_t::zero(), cc::fp_t::zero()};
    if (!inf) {
        cc::fp_t zi2 = inv.sqr();
        a = cc::affine_t{p.x * zi2, p.y * zi2 * inv};
    }
    cc::affine_compress(out + (size_t)o * 48, a);
}
// sum of `count` Jacobian points by one warp (multi-GPU combine: count = number of ranks).  Lane i takes points i,
// i + 32, ...; the 32 lane sums are folded by the quad tree of g1_quad.cuh (one plain addition, then every addition on a
// lane quad: ~4 us each instead of ~13 us), over as few levels as `count` needs.  The result stays quad-distributed
// and is written as Jacobian (X*ZZ, Y*ZZZ, ZZ), infinity all-zero, like k_group_finish.
__global__ void __launch_bounds__(32) k_g1_sum(const uint8_t* __restrict__ jac, uint8_t* __restrict__ out, int count) {
    const int lane = threadIdx.x;
    xyzz_t acc = xyzz_t::inf();
    for (int i = lane; i < count; i += 32) {
        xyzz_t p = jac_to_xyzz(load_jac(jac + (size_t)i * 144));
        xyzz_add(acc, p);
    }
    int S = 4;
    while (S < 32 && S < count) S <<= 1;
    fp_t q = seg_sum_quad(acc, S);
    fp_t t = q * shfl_xor_fp(q, 2);
    if (lane < 2) store_field(out + lane * 48, t);
    if (lane == 2) store_field(out + 96, q);
}
void launch_g1_sum(const void* jac_dev, void* out_jac_dev, int count, cudaStream_t stream) {
    k_g1_sum<<<1, 32, 0, stream>>>((const uint8_t*)jac_dev, (uint8_t*)out_jac_dev, count);
    B200_LAUNCH_CHECK();
}
__global__ void k_jac_to_affine(const uint8_t* __restrict__ jac, uint8_t* __restrict__ aff, int count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    cc::store_affine(aff + (size_t)i * 96, cc::jac_to_affine(cc::load_jac(jac + (size_t)i * 144)));
}
void launch_jac_to_affine(const void* jac_dev, void* affine_dev, int count, cudaStream_t stream) {
    if (count <= 0) return;
    k_jac_to_affine<<<div_up(count, 64), 64, 0, stream>>>((const uint8_t*)jac_dev, (uint8_t*)affine_dev, count);
    B200_LAUNCH_CHECK();
}
void launch_points_to_compressed(const void* jac_dev, uint8_t* out48_dev, int count, cudaStream_t stream, int brp_bits) {
    if (count <= 0) return;
    k_compress<<<div_up(count, 32), 32, 0, stream>>>((const uint8_t*)jac_dev, out48_dev, count, brp_bits);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// Microbenchmark (b200_bench_affine_pairs): the rate of INDEPENDENT batch-affine pair additions -- K pairs per thread,
// two gathered 96-byte points per pair from a large table, prefix products in global memory, one inversion per CTA through
// the shared-memory product tree -- i.e. one round of a tree-shaped batch-affine accumulation without any bucket logic.
// It bounds from above what such a design could reach against the XYZZ task kernel's 2.65 G additions/s.
template <int K>
__global__ void __launch_bounds__(kBaThreads, 3) k_affine_pairs_bench(const uint8_t* __restrict__ table, uint32_t npoints, size_t npairs,
                                                                    uint8_t* __restrict__ prefix, uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t sm_t[128 * 48];
    __shared__ __align__(16) uint8_t sm_i[128 * 48];
    const uint32_t tid = threadIdx.x;
    const size_t p0 = ((size_t)blockIdx.x * kBaThreads + tid) * K;
    uint4* pf = reinterpret_cast<uint4*>(prefix) + ((size_t)blockIdx.x * K * 3) * kBaThreads + tid;
    fp_t running = fp_t::one();
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const size_t p = p0 + k;
        if (p >= npairs) break;
        const uint32_t i1 = (uint32_t)(splitmix64(p) % npoints), i2 = (uint32_t)(splitmix64(~p) % npoints);
        const fp_t x1 = load_field_ro<fp_t>(table + (size_t)i1 * 96), x2 = load_field_ro<fp_t>(table + (size_t)i2 * 96);
        fp_t d = x2 - x1;
        if (d.is_zero()) d = fp_t::one();
        uint4* q = pf + (size_t)k * 3 * kBaThreads;
        q[0] = make_uint4(running.v[0], running.v[1], running.v[2], running.v[3]);
        q[kBaThreads] = make_uint4(running.v[4], running.v[5], running.v[6], running.v[7]);
        q[2 * kBaThreads] = make_uint4(running.v[8], running.v[9], running.v[10], running.v[11]);
        running = running * d;
    }
    fp_t inv = cta_batch_inverse(running, sm_t, sm_i);
#pragma unroll 1
    for (int k = K - 1; k >= 0; k--) {
        const size_t p = p0 + k;
        if (p >= npairs) continue;
        const uint32_t i1 = (uint32_t)(splitmix64(p) % npoints), i2 = (uint32_t)(splitmix64(~p) % npoints);
        const affine_t a = load_affine(table + (size_t)i1 * 96), b = load_affine(table + (size_t)i2 * 96);
        const uint4* q = pf + (size_t)k * 3 * kBaThreads;
        fp_t pre;
        {
            uint4 t0 = q[0], t1 = q[kBaThreads], t2 = q[2 * kBaThreads];
            pre.v[0] = t0.x; pre.v[1] = t0.y; pre.v[2] = t0.z; pre.v[3] = t0.w;
            pre.v[4] = t1.x; pre.v[5] = t1.y; pre.v[6] = t1.z; pre.v[7] = t1.w;
            pre.v[8] = t2.x; pre.v[9] = t2.y; pre.v[10] = t2.z; pre.v[11] = t2.w;
        }
        fp_t d = b.x - a.x;
        if (d.is_zero()) d = fp_t::one();
        const fp_t inv_d = pre * inv;
        inv = inv * d;
        const fp_t lam = (b.y - a.y) * inv_d;
        affine_t r;
        r.x = lam.sqr() - a.x - b.x;
        r.y = lam * (a.x - r.x) - a.y;
        store_affine(out + p * 96, r);
    }
}
// returns the kernel time in ms (CUDA events) for npairs additions over a table of npoints pseudo-random "points"
float bench_affine_pairs(uint32_t npoints, size_t npairs, int K, cudaStream_t st) {
    uint8_t* table = dev_alloc<uint8_t>((size_t)npoints * 96);
    uint8_t* out = dev_alloc<uint8_t>(npairs * 96);
    const size_t threads = (npairs + K - 1) / K, blocks = (threads + kBaThreads - 1) / kBaThreads;
    uint8_t* prefix = dev_alloc<uint8_t>(blocks * kBaThreads * (size_t)K * 48);
    // any field elements do: the cost of the arithmetic does not depend on the values (top limb cleared: < p)
    B200_CUDA_CHECK(cudaMemsetAsync(table, 0x5a, (size_t)npoints * 96, st));
    k_iota<<<div_up((size_t)npoints * 24, 256), 256, 0, st>>>((uint32_t*)table, (size_t)npoints * 24);   // distinct low limbs
    cudaEvent_t e0, e1;
    B200_CUDA_CHECK(cudaEventCreate(&e0));
    B200_CUDA_CHECK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        B200_CUDA_CHECK(cudaEventRecord(e0, st));
        if (K == 8) k_affine_pairs_bench<8><<<(unsigned)blocks, kBaThreads, 0, st>>>(table, npoints, npairs, prefix, out);
        else if (K == 16) k_affine_pairs_bench<16><<<(unsigned)blocks, kBaThreads, 0, st>>>(table, npoints, npairs, prefix, out);
        else if (K == 64) k_affine_pairs_bench<64><<<(unsigned)blocks, kBaThreads, 0, st>>>(table, npoints, npairs, prefix, out);
        else k_affine_pairs_bench<32><<<(unsigned)blocks, kBaThreads, 0, st>>>(table, npoints, npairs, prefix, out);
        B200_CUDA_CHECK(cudaEventRecord(e1, st));
        B200_CUDA_CHECK(cudaEventSynchronize(e1));
        B200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(table); cudaFree(out); cudaFree(prefix);
    B200_LAUNCH_CHECK();
    return ms;
}

typedef void (*ba_kernel_t)(const uint8_t*, uint32_t, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t, uint8_t*, uint8_t*);
static ba_kernel_t ba_kernel(int k, bool tree) {
    switch (k) {
        case 8: return tree ? k_accumulate_affine<8, true> : k_accumulate_affine<8, false>;
        case 12: return tree ? k_accumulate_affine<12, true> : k_accumulate_affine<12, false>;
        case 16: return tree ? k_accumulate_affine<16, true> : k_accumulate_affine<16, false>;
        default: return tree ? k_accumulate_affine<10, true> : k_accumulate_affine<10, false>;
    }
}

MsmEngine::MsmEngine(const MsmConfig& cfg, const void* points, bool host_points, cudaStream_t stream, const MsmEngine* share_table)
    : cfg_(cfg) {
    if (!cfg_.fixed || cfg_.c0 <= 0 || cfg_.c0 > cfg_.c) cfg_.c0 = cfg_.c;
    if (cfg_.c < 2 || cfg_.c > kMaxWindow || cfg_.c0 + cfg_.c * (cfg_.W - 1) < 256)
        throw CudaError(-1, "MsmEngine: bad window configuration");
    if (cfg_.L < 1 || cfg_.L > 1024) throw CudaError(-1, "MsmEngine: bad task length");
    if (!cfg_.fixed) cfg_.max_batch = 1;
    if (!cfg_.fixed || cfg_.bases_period < 1) cfg_.bases_period = 1;
    const size_t table_points = (size_t)cfg_.bases_period * cfg_.n;
    nb_ = 1 << (cfg_.c - 1);
    groups_max_ = cfg_.fixed ? (size_t)cfg_.max_batch : (size_t)cfg_.W;
    keys_max_ = groups_max_ * nb_;
    entries_max_ = (size_t)cfg_.max_batch * cfg_.n * cfg_.W;
    if (entries_max_ >= (1ull << 32) || (cfg_.fixed ? table_points * cfg_.W : cfg_.n) >= (1ull << 31))
        throw CudaError(-1, "MsmEngine: problem too large for 32-bit entry indices");
    tasks_max_ = entries_max_ / cfg_.L + keys_max_ + 1;
    if (!cfg_.fixed) cfg_.affine = false;
    if (getenv("B200_MSM_AFFINE_K")) cfg_.affine_k = atoi(getenv("B200_MSM_AFFINE_K"));
    if (cfg_.affine_k != 8 && cfg_.affine_k != 10 && cfg_.affine_k != 12 && cfg_.affine_k != 16) cfg_.affine_k = 10;
    ba_tree_ = getenv("B200_MSM_AFFINE_TREE") && atoi(getenv("B200_MSM_AFFINE_TREE")) != 0;
    if (cfg_.affine) {
        // one CTA per (SM, resident slot); the kernel is persistent, so the grid is exactly what fits
        int dev = 0, sms = 148, occ = 0;
        B200_CUDA_CHECK(cudaGetDevice(&dev));
        B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        ba_smem_ = (size_t)cfg_.affine_k * 3 * kBaThreads * 16 + 2 * 128 * 48 + (size_t)cfg_.affine_k * kBaThreads * 4;
        auto kern = ba_kernel(cfg_.affine_k, ba_tree_);
        B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba_smem_));
        B200_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBaThreads, ba_smem_));
        if (occ < 1) throw CudaError(-1, "MsmEngine: k_accumulate_affine does not fit an SM");
        ba_blocks_ = sms * occ;
        ba_chunks_max_ = (size_t)ba_blocks_ * kBaThreads * cfg_.affine_k;
        // every non-empty bucket and every chunk contributes at most one partial sum
        tasks_max_ = std::max(tasks_max_, keys_max_ + ba_chunks_max_ + 1);
        first_slot_ = dev_alloc<uint32_t>(ba_chunks_max_);
        acc_buf_ = dev_alloc<uint8_t>(ba_chunks_max_ * 96);
    }
    size_t rows = cfg_.fixed ? cfg_.W : 1;
    stride_ = cfg_.affine ? 128 : 96;
    table_bytes_ = rows * table_points * stride_;
    if (share_table) {
        const MsmConfig& o = share_table->cfg_;
        if (!cfg_.fixed || !o.fixed || o.c != cfg_.c || o.c0 != cfg_.c0 || o.W != cfg_.W || o.n != cfg_.n || o.bases_period != cfg_.bases_period ||
            share_table->stride_ != stride_)
            throw CudaError(-1, "MsmEngine: shared table has a different layout");
        table_ = share_table->table_;
        rho_seed_ = share_table->rho_seed_;
        owns_table_ = false;
        points = nullptr;
    } else {
        table_ = dev_alloc<uint8_t>(table_bytes_);
    }
    counts_ = dev_alloc<uint32_t>(keys_max_ + 1);
    offsets_ = dev_alloc<uint32_t>(keys_max_ + 1);
    cursor_ = dev_alloc<uint32_t>(keys_max_ + 1);
    task_base_ = dev_alloc<uint32_t>(keys_max_ + 1);
    entries_ = dev_alloc<uint32_t>(entries_max_);
    sorted_tasks_ = dev_alloc<uint32_t>(3 * tasks_max_);
    size_hist_ = dev_alloc<uint32_t>(3 * (cfg_.L + 1));
    scan_tmp_ = dev_alloc<uint32_t>(kScanBlock * kScanItems + 1);
    partials_ = dev_alloc<uint8_t>(tasks_max_ * 192);
    chunk_sums_ = dev_alloc<uint8_t>(groups_max_ * 3 * 32 * 192);  // marginal sums [group][axis][32]
    fin_scratch_ = dev_alloc<uint8_t>(groups_max_ * kFinSlots * 192);   // k_group_finish: bit sums of every group
    fin_cnt_ = dev_alloc<unsigned>(groups_max_);
    B200_CUDA_CHECK(cudaMemsetAsync(fin_cnt_, 0, groups_max_ * sizeof(unsigned), stream));
    group_sums_ = dev_alloc<uint8_t>(groups_max_ * 192);
    kf_ = std::max(cfg_.c - 1 > kReduceBits ? cfg_.c - 1 - kReduceBits : 0, std::min(cfg_.fold, cfg_.c - 2));
    if (kf_ > 0) {
        // folded segments (T and R point per segment), their identity slot map, marginals of the R sums
        const size_t nseg = keys_max_ >> kf_;
        seg_t_ = dev_alloc<uint8_t>(nseg * 192);
        seg_r_ = dev_alloc<uint8_t>(nseg * 192);
        seg_ident_ = dev_alloc<uint32_t>(nseg + 1);
        chunk_sums_r_ = dev_alloc<uint8_t>(groups_max_ * 3 * 32 * 192);
        k_iota<<<div_up(nseg + 1, 256), 256, 0, stream>>>(seg_ident_, nseg + 1);
        B200_LAUNCH_CHECK();
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    if (points) {
        const cudaMemcpyKind kind = host_points ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
        if (stride_ == 96) B200_CUDA_CHECK(cudaMemcpyAsync(table_, points, table_points * 96, kind, stream));
        else B200_CUDA_CHECK(cudaMemcpy2DAsync(table_, stride_, points, 96, 96, table_points, kind, stream));
        if (cfg_.fixed && cfg_.randomize) {
            // scalar randomisation needs bases of the prime-order subgroup: check, then scale row 0 by rho_i^-1
            int* d_bad = dev_alloc<int>(1);
            int bad = 0;
            B200_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), stream));
            k_check_subgroup<<<div_up(table_points, 8), 32, 0, stream>>>((const uint8_t*)table_, stride_, table_points, d_bad);
            cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
            cudaFree(d_bad);
            B200_CUDA_CHECK(e);
            if (!bad) {
                rho_seed_ = cfg_.rho_seed ? cfg_.rho_seed : 0x4b5a47b200ull;
                k_scale_points<<<div_up(table_points, 8), 32, 0, stream>>>((uint8_t*)table_, stride_, table_points, rho_seed_);
                B200_LAUNCH_CHECK();
            }
        }
        if (cfg_.fixed && cfg_.W > 1) {
            k_build_rows<<<div_up(table_points, 128), 128, 0, stream>>>((uint8_t*)table_, table_points, cfg_.W, cfg_.c, cfg_.c0, stride_);
            B200_LAUNCH_CHECK();
        }
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
}

void MsmEngine::profile_read(double* accumulate_ms_sum, int* runs) {
    double sum = 0;
    for (int i = 0; i < prof_count_; i++) {
        float ms = 0;
        B200_CUDA_CHECK(cudaEventSynchronize(prof_ev_[2 * i + 1]));
        B200_CUDA_CHECK(cudaEventElapsedTime(&ms, prof_ev_[2 * i], prof_ev_[2 * i + 1]));
        sum += ms;
    }
    if (accumulate_ms_sum) *accumulate_ms_sum = sum;
    if (runs) *runs = prof_count_;
    prof_count_ = 0;
}

void MsmEngine::last_counts(size_t* entries, size_t* tasks, cudaStream_t stream) {
    uint32_t e = 0, t = 0;
    if (last_nkeys_) {
        B200_CUDA_CHECK(cudaMemcpyAsync(&e, offsets_ + last_nkeys_, 4, cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaMemcpyAsync(&t, task_base_ + last_nkeys_, 4, cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    if (entries) *entries = e;
    if (tasks) *tasks = t;
}

// [entries, tasks, non-empty buckets, keys, fold bits, buckets per group seen by the marginal reduce, digit axes, groups]
void MsmEngine::last_stats(uint64_t out[8], cudaStream_t stream) {
    uint32_t e = 0, t = 0, ne = 0;
    if (last_nkeys_) {
        B200_CUDA_CHECK(cudaMemcpyAsync(&e, offsets_ + last_nkeys_, 4, cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaMemcpyAsync(&t, task_base_ + last_nkeys_, 4, cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaMemcpyAsync(&ne, size_hist_, 4, cudaMemcpyDeviceToHost, stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    const int bits = cfg_.c - 1 - kf_;
    out[0] = e; out[1] = t; out[2] = ne; out[3] = last_nkeys_; out[4] = (uint64_t)kf_; out[5] = (uint64_t)(nb_ >> kf_);
    out[6] = (uint64_t)std::max(1, (bits + 4) / 5); out[7] = last_nkeys_ / (size_t)nb_;
}

MsmEngine::~MsmEngine() {
    for (auto& e : prof_ev_)
        if (e) cudaEventDestroy(e);
    for (auto& e : copy_ev_)
        if (e) cudaEventDestroy(e);
    if (copy_start_) cudaEventDestroy(copy_start_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    cudaFree(first_slot_); cudaFree(acc_buf_);
    if (owns_table_) cudaFree(table_);
    cudaFree(counts_); cudaFree(offsets_); cudaFree(cursor_); cudaFree(task_base_); cudaFree(entries_);
    cudaFree(sorted_tasks_); cudaFree(size_hist_); cudaFree(scan_tmp_); cudaFree(partials_); cudaFree(chunk_sums_); cudaFree(fin_scratch_); cudaFree(fin_cnt_);
    cudaFree(group_sums_);
    cudaFree(seg_t_); cudaFree(seg_r_); cudaFree(seg_ident_); cudaFree(chunk_sums_r_);
}

void MsmEngine::set_points(const void* points_dev, size_t npoints, cudaStream_t stream) {
    if (cfg_.fixed) throw CudaError(-1, "set_points on a FIXED engine");
    if (npoints > cfg_.n) throw CudaError(-1, "set_points: too many points");
    B200_CUDA_CHECK(cudaMemcpyAsync(table_, points_dev, npoints * 96, cudaMemcpyDeviceToDevice, stream));
}

void MsmEngine::run(const void* scalars_dev, size_t npoints, int batch, bool mont, void* out_dev, cudaStream_t st,
                    const void* scalars_host) {
    if (npoints > cfg_.n || batch < 1 || batch > cfg_.max_batch) throw CudaError(-1, "MsmEngine::run: bad sizes");
    const int c = cfg_.c, W = cfg_.W, c0 = cfg_.c0;
    const size_t groups = cfg_.fixed ? (size_t)batch : (size_t)W;
    const size_t nkeys = groups * nb_;
    const size_t total = (size_t)batch * npoints;
    // task length: cfg_.L when the call fills the machine; short calls (one blob, 2^12-point MSMs) are latency-bound
    // -- a thread's serial chain is L additions of ~10 us each with one warp per scheduler -- so they are cut into
    // more, shorter tasks (down to kMinTaskLen) and the partial sums folded by the sub-warp trees of k_bucket_combine
    int L = cfg_.L;
    {
        // two warps on every SM sub-partition; B200_MSM_FILL_DIV > 1 aims at a fraction of the machine (several lanes of
        // a settings object running side by side each get their share: fewer, longer tasks, less combine work)
        static const int fill_div = std::max(1, getenv("B200_MSM_FILL_DIV") ? atoi(getenv("B200_MSM_FILL_DIV")) : 1);
        const size_t fill = (size_t)148 * 4 * 32 * 2 / fill_div;
        size_t want = std::max<size_t>(kMinTaskLen, total * W / fill);
        if ((size_t)L > want) L = (int)want;
        while (L < cfg_.L && total * W / L + nkeys + 1 > tasks_max_) L++;
    }
    int launches = 0;
    if (total == 0) {
        B200_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, (size_t)batch * 144, st));
        launches_ = 0;
        return;
    }
    B200_CUDA_CHECK(cudaMemsetAsync(counts_, 0, (nkeys + 1) * sizeof(uint32_t), st));
    B200_CUDA_CHECK(cudaMemsetAsync(size_hist_, 0, 3 * (L + 1) * sizeof(uint32_t), st));
    // 1 digits + histogram
    const size_t row_stride = (size_t)cfg_.bases_period * cfg_.n;
    if (scalars_host) {
        // host scalars: copy in chunks on a second stream and histogram each chunk as it lands, so the PCIe transfer
        // overlaps the first kernel instead of preceding it
        if (!copy_stream_) {
            B200_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
            for (auto& e : copy_ev_) B200_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            B200_CUDA_CHECK(cudaEventCreateWithFlags(&copy_start_, cudaEventDisableTiming));
        }
        const int nchunks = total >= (1u << 16) ? kCopyChunks : 1;
        const size_t per = (total + nchunks - 1) / nchunks;
        B200_CUDA_CHECK(cudaEventRecord(copy_start_, st));           // the staging buffer is free once st reaches here
        B200_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, copy_start_, 0));
        for (int k = 0; k < nchunks; k++) {
            size_t lo = (size_t)k * per, hi = std::min(total, lo + per);
            if (lo >= hi) break;
            B200_CUDA_CHECK(cudaMemcpyAsync((uint8_t*)scalars_dev + lo * 32, (const uint8_t*)scalars_host + lo * 32, (hi - lo) * 32,
                                            cudaMemcpyHostToDevice, copy_stream_));
            B200_CUDA_CHECK(cudaEventRecord(copy_ev_[k], copy_stream_));
            B200_CUDA_CHECK(cudaStreamWaitEvent(st, copy_ev_[k], 0));
            k_digits<false><<<div_up(hi - lo, 256), 256, 0, st>>>((const uint4*)scalars_dev, npoints, row_stride, hi, c, c0, W, nb_, cfg_.fixed,
                                                                 mont, counts_, nullptr, (size_t)cfg_.bases_period, cfg_.n, lo, rho_seed_);
            launches++;
        }
    } else {
        k_digits<false><<<div_up(total, 256), 256, 0, st>>>((const uint4*)scalars_dev, npoints, row_stride, total, c, c0, W, nb_,
                                                            cfg_.fixed, mont, counts_, nullptr, (size_t)cfg_.bases_period, cfg_.n, 0, rho_seed_);
        launches++;
    }
    // batch-affine accumulation when the call fills the machine: S entries per chunk, chunk grid = every slot of every thread
    const size_t entries_bound = total * W;
    uint32_t S = 0;
    if (cfg_.affine && entries_bound >= ((size_t)1 << 21)) {
        S = (uint32_t)((entries_bound + ba_chunks_max_ - 1) / ba_chunks_max_);
        if (S < 4) S = 4;
    }
    last_affine_ = S != 0;
    // 2 offsets (and a working copy for the scatter cursors), task bases
    launches += scan_exclusive(counts_, nkeys, 0, offsets_, cursor_, scan_tmp_, st);
    if (S) launches += scan_exclusive(offsets_, nkeys, kScanSegments | S, task_base_, nullptr, scan_tmp_, st);   // segments per bucket
    else launches += scan_exclusive(counts_, nkeys, (uint32_t)L, task_base_, nullptr, scan_tmp_, st);
    // 3 scatter
    k_digits<true><<<div_up(total, 256), 256, 0, st>>>((const uint4*)scalars_dev, npoints, row_stride, total, c, c0, W, nb_,
                                                       cfg_.fixed, mont, cursor_, entries_, (size_t)cfg_.bases_period, cfg_.n, 0, rho_seed_);
    launches++;
    const bool prof = profiling_ && prof_count_ < kProfSlots;
    auto prof_begin = [&] {
        if (!prof) return;
        for (int k = 0; k < 2; k++)
            if (!prof_ev_[2 * prof_count_ + k]) B200_CUDA_CHECK(cudaEventCreate(&prof_ev_[2 * prof_count_ + k]));
        B200_CUDA_CHECK(cudaEventRecord(prof_ev_[2 * prof_count_], st));
    };
    if (S) {
        // 4' bucket-start flags, first partial slot of every chunk
        const size_t nchunks = (entries_bound + S - 1) / S;
        k_mark_starts<<<div_up(nkeys, 256), 256, 0, st>>>(offsets_, nkeys, entries_, size_hist_);
        k_chunk_first<<<div_up(nchunks, 256), 256, 0, st>>>(offsets_, task_base_, nkeys, S, nchunks, first_slot_);
        launches += 2;
        // 5' accumulate: persistent grid, every slot of every thread advances one entry per step
        prof_begin();
        ba_kernel(cfg_.affine_k, ba_tree_)<<<ba_blocks_, kBaThreads, ba_smem_, st>>>((const uint8_t*)table_, (uint32_t)stride_, entries_, offsets_ + nkeys,
                                                                         first_slot_, S, acc_buf_, (uint8_t*)partials_);
    } else {
        // 4 tasks sorted by length
        k_task_hist<<<div_up(nkeys, 256), 256, (L + 1) * sizeof(uint32_t), st>>>(counts_, nkeys, L, size_hist_);
        k_task_bases<<<1, 32, 0, st>>>(size_hist_, L);
        k_task_emit<<<div_up(nkeys, 256), 256, 0, st>>>(counts_, offsets_, task_base_, nkeys, L, size_hist_, sorted_tasks_);
        launches += 3;
        // 5 accumulate: grid sized for the worst case, surplus threads exit on the device-side task count
        prof_begin();
        size_t tasks_bound = std::min(tasks_max_, total * W / L + nkeys + 1);
        static const bool acc_call = getenv("B200_ACC_CALL") && atoi(getenv("B200_ACC_CALL"));
        static const bool acc_prefetch = !getenv("B200_ACC_PREFETCH") || atoi(getenv("B200_ACC_PREFETCH"));  // default on: -2.2% at 2^20
        static const int acc_kara = getenv("B200_ACC_KARA") ? atoi(getenv("B200_ACC_KARA")) : 0;
        if (acc_kara)
            (acc_kara == 2 ? k_accumulate_kara<2> : k_accumulate_kara<3>)<<<div_up(tasks_bound, kAccThreads), kAccThreads, 0, st>>>(
                (const uint8_t*)table_, (uint32_t)stride_, entries_, sorted_tasks_, task_base_ + nkeys, (uint8_t*)partials_);
        else if (acc_call)
            k_accumulate_call<<<div_up(tasks_bound, kAccThreads), kAccThreads, 0, st>>>((const uint8_t*)table_, (uint32_t)stride_, entries_,
                                                                                       sorted_tasks_, task_base_ + nkeys, (uint8_t*)partials_);
        else {
            static const int acc_occ = getenv("B200_ACC_OCC") ? atoi(getenv("B200_ACC_OCC")) : 3;
            auto kern = acc_prefetch ? (acc_occ == 4 ? k_accumulate<true, 4> : k_accumulate<true, 3>)
                                     : (acc_occ == 4 ? k_accumulate<false, 4> : k_accumulate<false, 3>);
            kern<<<div_up(tasks_bound, kAccThreads), kAccThreads, 0, st>>>((const uint8_t*)table_, (uint32_t)stride_, entries_, sorted_tasks_,
                                                                           task_base_ + nkeys, (uint8_t*)partials_);
        }
    }
    if (prof) {
        B200_CUDA_CHECK(cudaEventRecord(prof_ev_[2 * prof_count_ + 1], st));
        prof_count_++;
    }
    launches++;
    // 6 reduce
    // (batch-affine path: almost every bucket has 2-4 segment sums -- they are what the machine is filled with)
    const uint32_t warp_min = !S && (double)total * W / (double)nkeys <= 0.5 * L ? 4 : 32;
    if (nkeys <= 8192)
        k_bucket_combine<false, true><<<div_up(nkeys * kCombLanes, 128), 128, 0, st>>>((uint8_t*)partials_, task_base_, nkeys, warp_min);
    else
        k_bucket_combine<false, false><<<div_up(nkeys, 128), 128, 0, st>>>((uint8_t*)partials_, task_base_, nkeys, warp_min);
    k_bucket_combine<true, false><<<148 * 4, 128, 0, st>>>((uint8_t*)partials_, task_base_, nkeys, warp_min);
    // wide windows: fold segments of 2^kf buckets first; the marginal reduce then sees nbr = 2^15 "buckets" T_hi per group
    const int kf = kf_;
    const int nbr = nb_ >> kf;
    const uint8_t* rpart = (const uint8_t*)partials_;
    const uint32_t* rbase = task_base_;
    if (kf) {
        const size_t nseg = groups * nbr;
        k_segment_fold<<<div_up(nseg, 128), 128, 0, st>>>((const uint8_t*)partials_, task_base_, nseg, kf, (uint8_t*)seg_t_, (uint8_t*)seg_r_);
        launches++;
        rpart = (const uint8_t*)seg_t_;
        rbase = seg_ident_;
    }
    AxisPlan ap{};
    {
        int bits = c - 1 - kf;
        ap.D = (bits + 4) / 5;
        if (ap.D < 1) ap.D = 1;
        int off = 0;
        for (int a = 0; a < 3; a++) {
            int wa = a < ap.D ? (bits - off + (ap.D - a) - 1) / (ap.D - a) : 0;
            ap.w[a] = wa;
            ap.sh[a] = off;
            off += wa;
        }
    }
    if (groups * 32 * ap.D >= 2048) {
        // many groups: lane-efficient sub-warp marginals, one grid row per digit axis (plus one for the R points after
        // a segment fold).  Slots of digit values that do not exist (v >= 2^w) must read as infinity (all-zero XYZZ).
        B200_CUDA_CHECK(cudaMemsetAsync(chunk_sums_, 0, groups * 3 * 32 * 192, st));
        if (kf) B200_CUDA_CHECK(cudaMemsetAsync(chunk_sums_r_, 0, groups * 3 * 32 * 192, st));
        size_t max_threads = 0;
        for (int a = 0; a < ap.D + (kf ? 1 : 0); a++) {
            const int wa = a < ap.D ? ap.w[a] : (nbr >= 32 ? 5 : 0);
            int log_s = 0;
            while (log_s < 5 && ((nbr >> wa) >> log_s) > kMargSerial) log_s++;
            max_threads = std::max(max_threads, (groups << wa) << log_s);
        }
        k_marginals_sub<<<dim3(div_up(max_threads, 128), ap.D + (kf ? 1 : 0)), 128, 0, st>>>(
            rpart, rbase, nbr, ap, groups, (uint8_t*)chunk_sums_, (const uint8_t*)seg_r_, (uint8_t*)chunk_sums_r_);
        launches += 3;
    } else {
        // one CTA per marginal; 256 threads once a marginal covers >= 1024 buckets (shorter serial chains).
        // Wide windows: one more grid row sums the R_hi (32 plain marginals, added up by k_group_finish's fourth warp).
        const dim3 mgrid(32, ap.D + (kf ? 1 : 0), (unsigned)groups);
        if ((nbr >> ap.w[0]) >= 1024)
            k_marginals<256><<<mgrid, 256, 0, st>>>(rpart, rbase, nbr, ap, (uint8_t*)chunk_sums_, (const uint8_t*)seg_r_,
                                                    (uint8_t*)chunk_sums_r_);
        else
            k_marginals<128><<<mgrid, 128, 0, st>>>(rpart, rbase, nbr, ap, (uint8_t*)chunk_sums_, (const uint8_t*)seg_r_,
                                                    (uint8_t*)chunk_sums_r_);
        launches += 3;
    }
    const uint8_t* marg_r = kf ? (const uint8_t*)chunk_sums_r_ : nullptr;
    if (cfg_.fixed) {
        k_group_finish<<<dim3(kFinCtas, (unsigned)groups), kFinThreads, 0, st>>>((const uint8_t*)chunk_sums_, ap, nullptr, (uint8_t*)out_dev, marg_r, kf,
                                                                              (uint8_t*)fin_scratch_, fin_cnt_);
        launches += 1;
    } else {
        k_group_finish<<<dim3(kFinCtas, (unsigned)groups), kFinThreads, 0, st>>>((const uint8_t*)chunk_sums_, ap, (uint8_t*)group_sums_, nullptr, marg_r, kf,
                                                                              (uint8_t*)fin_scratch_, fin_cnt_);
        k_horner<<<1, 32, 0, st>>>((const uint8_t*)group_sums_, W, c, (uint8_t*)out_dev);
        launches += 2;
    }
    B200_LAUNCH_CHECK();
    launches_ = launches;
    last_nkeys_ = nkeys;
}

}  // namespace b200
