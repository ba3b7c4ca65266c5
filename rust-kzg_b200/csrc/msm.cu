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
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_digits(const uint4* __restrict__ scalars, size_t n, size_t row_stride, size_t total,
                                                int c, int c0, int W, int nb, int fixed, int mont, uint32_t* __restrict__ ctr,
                                                uint32_t* __restrict__ entries, size_t period, size_t period_n, size_t gid0) {
    size_t gid = gid0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t vec = gid / n, i = gid - vec * n;
    fr_t s = load_field_ro<fr_t>(scalars + 2 * gid);
    if (mont) s = s.from_mont();
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
// f: 0 = identity; f > 0: ceil(x / f) (task counts); f | 0x80000000: floor(x / f) (pair counts)
__device__ __forceinline__ uint32_t scan_map(uint32_t x, uint32_t f) {
    if (f == 0) return x;
    if (f & 0x80000000u) return x / (f & 0x7fffffffu);
    return (x + f - 1) / f;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_partial(const uint32_t* __restrict__ in, size_t n, uint32_t f,
                                                             uint32_t* __restrict__ block_sums) {
    size_t base = ((size_t)blockIdx.x * kScanBlock + threadIdx.x) * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) { uint32_t x = in[base + k]; s += scan_map(x, f); }
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
        uint32_t x = base + k < n ? in[base + k] : 0;
        v[k] = scan_map(x, f);
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
__global__ void __launch_bounds__(kAccThreads, MINB) k_accumulate(const uint8_t* __restrict__ table,
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
        affine_t p = load_affine(table + (size_t)(v & 0x7fffffffu) * 96);
        for (uint32_t k = 0; k < len; k++) {
            affine_t cur = p;
            uint32_t cv = v;
            if (k + 1 < len) {
                v = e[k + 1];
                p = load_affine(table + (size_t)(v & 0x7fffffffu) * 96);
            }
            cur.y = cur.y.cneg(cv >> 31);
            xyzz_add_affine(acc, cur);
        }
    } else {
        for (uint32_t k = 0; k < len; k++) {
            uint32_t v = e[k];
            affine_t p = load_affine(table + (size_t)(v & 0x7fffffffu) * 96);
            p.y = p.y.cneg(v >> 31);
            xyzz_add_affine(acc, p);
        }
    }
    store_xyzz(partials + (size_t)slot * 192, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// 5a: batch-affine pre-reduction.  A mixed XYZZ addition costs 8M + 2S; adding two AFFINE points costs 3M once
// 1/(x2 - x1) is known, and inverting K denominators together (Montgomery's trick: 3M each + one inversion) makes
// that 6M per addition.  The inversion itself is the binary-Euclid routine, which runs on the ALU pipe while the
// multiplications keep the FMA-heavy pipe busy.  Buckets are summed as trees: every round adds adjacent pairs of each
// bucket's point list (all pairs of all buckets are independent, so a thread takes K consecutive pairs of the flat
// pair list, across bucket boundaries), halving the lists; after a few rounds the XYZZ task kernel finishes the rest.
// All exceptional cases are explicit: infinity operands, P + P (tangent slope) and P + (-P) (result infinity).
static constexpr int kAffK = 16;  // pairs per thread = denominators per inversion

struct AffinePair {
    affine_t p1, p2;
};
// gather mode (round 0): points come from the table through the sorted entries (index | sign << 31)
__device__ __forceinline__ affine_t load_round_point(const uint8_t* __restrict__ table, const uint32_t* __restrict__ entries,
                                                     const uint8_t* __restrict__ buf, size_t pos, bool gather) {
    if (gather) {
        uint32_t v = entries[pos];
        affine_t p = load_affine(table + (size_t)(v & 0x7fffffffu) * 96);
        if (!p.is_inf()) p.y = p.y.cneg(v >> 31);
        return p;
    }
    return load_affine(buf + pos * 96);
}
// denominator of the slope: x2 - x1, or 2*y1 when the points coincide; zero means "no inversion needed"
__device__ __forceinline__ fp_t pair_denominator(const affine_t& p1, const affine_t& p2) {
    if (p1.is_inf() || p2.is_inf()) return fp_t::zero();
    fp_t dx = p2.x - p1.x;
    if (!dx.is_zero()) return dx;
    if (p1.y == p2.y) return p1.y.dbl();
    return fp_t::zero();  // P + (-P)
}
__global__ void __launch_bounds__(128) k_affine_round(const uint8_t* __restrict__ table, const uint32_t* __restrict__ entries,
                                                      const uint8_t* __restrict__ in_buf, uint8_t* __restrict__ out_buf,
                                                      const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ counts,
                                                      const uint32_t* __restrict__ pair_base, size_t nkeys, int gather) {
    const size_t total_pairs = pair_base[nkeys];
    size_t p0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kAffK;
    if (p0 >= total_pairs) return;
    const int npairs = (int)min((size_t)kAffK, total_pairs - p0);
    // locate the bucket of the first pair: last key with pair_base[key] <= p0
    size_t lo = 0, hi = nkeys;
    while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if (pair_base[mid] <= p0) lo = mid; else hi = mid;
    }
    size_t key0 = lo;
    uint32_t k0 = (uint32_t)(p0 - pair_base[key0]);

    fp_t prefix[kAffK];
    fp_t acc = fp_t::one();
    {
        size_t key = key0;
        uint32_t k = k0, npk = counts[key] >> 1, off = offsets[key];
#pragma unroll 1
        for (int i = 0; i < npairs; i++) {
            while (k >= npk) { key++; k = 0; npk = counts[key] >> 1; off = offsets[key]; }
            affine_t a = load_round_point(table, entries, in_buf, (size_t)off + 2 * k, gather);
            affine_t b = load_round_point(table, entries, in_buf, (size_t)off + 2 * k + 1, gather);
            fp_t d = pair_denominator(a, b);
            prefix[i] = acc;
            if (!d.is_zero()) acc = acc * d;
            k++;
        }
    }
    fp_t inv = acc.inverse();
    // second sweep, backwards, recomputing the (cheap) denominators: inv_i = prefix_i * inv; inv *= d_i
    {
        // re-walk forward to find the position of the last pair, then go back; simpler: recompute positions per pair
        // by walking forward again and storing (key, k) compactly
        uint32_t pk_key[kAffK], pk_k[kAffK];
        size_t key = key0;
        uint32_t k = k0, npk = counts[key] >> 1;
#pragma unroll 1
        for (int i = 0; i < npairs; i++) {
            while (k >= npk) { key++; k = 0; npk = counts[key] >> 1; }
            pk_key[i] = (uint32_t)key;
            pk_k[i] = k;
            k++;
        }
#pragma unroll 1
        for (int i = npairs - 1; i >= 0; i--) {
            uint32_t off = offsets[pk_key[i]];
            uint32_t kk = pk_k[i];
            affine_t a = load_round_point(table, entries, in_buf, (size_t)off + 2 * kk, gather);
            affine_t b = load_round_point(table, entries, in_buf, (size_t)off + 2 * kk + 1, gather);
            affine_t r;
            if (a.is_inf()) {
                r = b;
            } else if (b.is_inf()) {
                r = a;
            } else {
                fp_t dx = b.x - a.x;
                bool dbl = dx.is_zero();
                if (dbl && !(a.y == b.y)) {
                    r.x = fp_t::zero();
                    r.y = fp_t::zero();  // P + (-P)
                } else {
                    fp_t d = dbl ? a.y.dbl() : dx;
                    fp_t inv_i = prefix[i] * inv;
                    inv = inv * d;
                    fp_t num;
                    if (dbl) {
                        fp_t xx = a.x.sqr();
                        num = xx.dbl() + xx;  // 3 x^2
                    } else {
                        num = b.y - a.y;
                    }
                    fp_t lam = num * inv_i;
                    fp_t x3 = lam.sqr() - a.x - b.x;
                    r.x = x3;
                    r.y = lam * (a.x - x3) - a.y;
                }
            }
            store_affine(out_buf + ((size_t)off + kk) * 96, r);
        }
    }
}
// odd bucket sizes: the unpaired last point moves to the end of the halved list; then counts := ceil(counts / 2)
__global__ void __launch_bounds__(256) k_affine_leftover(const uint8_t* __restrict__ table, const uint32_t* __restrict__ entries,
                                                         const uint8_t* __restrict__ in_buf, uint8_t* __restrict__ out_buf,
                                                         const uint32_t* __restrict__ offsets, uint32_t* __restrict__ counts,
                                                         size_t nkeys, int gather) {
    size_t key = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nkeys) return;
    uint32_t c = counts[key];
    if (c & 1) {
        uint32_t off = offsets[key];
        affine_t p = load_round_point(table, entries, in_buf, (size_t)off + c - 1, gather);
        store_affine(out_buf + ((size_t)off + (c >> 1)) * 96, p);
    }
    counts[key] = (c + 1) >> 1;
}
// the XYZZ task kernel on an already materialised (and partly reduced) point list: contiguous reads, no gather
__global__ void __launch_bounds__(kAccThreads) k_accumulate_direct(const uint8_t* __restrict__ buf, const uint32_t* __restrict__ sorted_tasks,
                                                                   const uint32_t* __restrict__ n_tasks_ptr, uint8_t* __restrict__ partials) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *n_tasks_ptr) return;
    uint32_t start = sorted_tasks[3 * t], len = sorted_tasks[3 * t + 1], slot = sorted_tasks[3 * t + 2];
    xyzz_t acc = xyzz_t::inf();
    for (uint32_t k = 0; k < len; k++) {
        affine_t p = load_affine(buf + ((size_t)start + k) * 96);
        xyzz_add_affine(acc, p);
    }
    store_xyzz(partials + (size_t)slot * 192, acc);
}

// variant with the field multiplication out of line (B200_ACC_CALL=1): same work, ~20x smaller loop body
__global__ void __launch_bounds__(kAccThreads) k_accumulate_call(const uint8_t* __restrict__ table,
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
        cl::affine_t p = cl::load_affine(table + (size_t)(v & 0x7fffffffu) * 96);
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

// 6c: one CTA per group, one warp per digit axis: weighted sum over the <= 32 marginals, scale by 2^sh, combine.
// Everything after the suffix scan runs on quad-distributed points (g1_quad.cuh).
// marg_r != nullptr (wide windows, see k_segment_fold): the fourth warp sums the 32 plain marginals of the R_hi and the
// result is 2^kf * (15-bit reduce of the T_hi) - sum R.
__global__ void __launch_bounds__(128) k_group_finish(const uint8_t* __restrict__ marg, AxisPlan ap,
                                                      uint8_t* __restrict__ group_sums, uint8_t* __restrict__ out_jac,
                                                      const uint8_t* __restrict__ marg_r, int kf) {
    __shared__ __align__(16) uint8_t sh[5 * 192];
    const size_t g = blockIdx.x;
    const int lane = threadIdx.x & 31, a = threadIdx.x >> 5;
    const int off = quad_store_offset();
    if (a < ap.D) {
        xyzz_t m = load_xyzz(marg + ((g * 3 + a) * 32 + lane) * 192);
        const int width = 1 << ap.w[a];                  // marginals of this axis; the other lanes read infinity
        xyzz_t suf = warp_suffix_scan_xyzz(m, width);    // Suf_l = sum_{v >= l} M_v ; Suf_0 = sum of all buckets
        if (a == 0 && lane == 0) store_xyzz(sh + 3 * 192, suf);
        if (lane == 0) suf = xyzz_t::inf();              // sum_v v*M_v = sum_{k >= 1} Suf_k
        fp_t w = seg_sum_quad(suf, width < 4 ? 4 : width);
        for (int k = 0; k < ap.sh[a]; k++) w = quad_dbl(w);
        if (lane < 4) store_field(sh + a * 192 + off, w);
    } else if (a == 3 && marg_r) {
        xyzz_t m = load_xyzz(marg_r + ((g * 3) * 32 + lane) * 192);
        fp_t w = seg_sum_quad(m, 32);
        if (lane < 4) store_field(sh + 4 * 192 + off, w);
    }
    __syncthreads();
    if (a == 0) {
        fp_t acc = load_field<fp_t>(sh + 3 * 192 + off);  // the "+1" of the weights b+1
        for (int k = 0; k < ap.D; k++) acc = quad_add(acc, load_field<fp_t>(sh + k * 192 + off));
        if (marg_r) {
            for (int k = 0; k < kf; k++) acc = quad_dbl(acc);
            fp_t r = load_field<fp_t>(sh + 4 * 192 + off);
            if ((lane & 3) == 1) r = r.neg();             // -(X, Y, ZZ, ZZZ) = (X, -Y, ZZ, ZZZ); infinity stays all-zero
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

// table rows for FIXED engines: row j = 2^(c*j) * P_i, affine.  One thread per point walks all rows
// (c doublings in XYZZ, then back to affine with one warp-shared field inversion).  One-time cost at prepare.
__global__ void __launch_bounds__(128) k_build_rows(uint8_t* __restrict__ table, size_t n, int W, int c, int c0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    if (!live) i = n - 1;
    cc::affine_t p = cc::load_affine(table + i * 96);
    for (int j = 1; j < W; j++) {
        cc::xyzz_t q = cc::affine_to_xyzz(p);
        for (int k = 0; k < (j == 1 ? c0 : c); k++) cc::xyzz_dbl(q);   // row j = 2^(c0 + (j-1) c) * P
        const bool inf = q.is_inf();
        // 1/ZZZ; then 1/ZZ = ZZZ^-2 * ZZ^2  (ZZ^3 = ZZZ^2), as xyzz_to_affine
        cc::fp_t izzz = warp_inverse(inf ? cc::fp_t::one() : q.zzz);
        cc::fp_t izz = izzz.sqr() * q.zz.sqr();
        p = inf ? cc::affine_t{cc::fp_t::zero(), cc::fp_t::zero()} : cc::affine_t{q.x * izz, q.y * izzz};
        if (live) cc::store_affine(table + ((size_t)j * n + i) * 96, p);
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
    size_t rows = cfg_.fixed ? cfg_.W : 1;
    table_bytes_ = rows * table_points * 96;
    if (share_table) {
        const MsmConfig& o = share_table->cfg_;
        if (!cfg_.fixed || !o.fixed || o.c != cfg_.c || o.c0 != cfg_.c0 || o.W != cfg_.W || o.n != cfg_.n || o.bases_period != cfg_.bases_period)
            throw CudaError(-1, "MsmEngine: shared table has a different layout");
        table_ = share_table->table_;
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
    group_sums_ = dev_alloc<uint8_t>(groups_max_ * 192);
    pair_base_ = dev_alloc<uint32_t>(keys_max_ + 1);
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
    {
        const char* e = getenv("B200_AFFINE_ROUNDS");
        // OFF by default: measured on B200 (MSM 2^20) the rounds make the accumulation phase 16.1 ms instead of 6.4 ms --
        // one binary-Euclid inversion per 16 additions costs more ALU-pipe time than the 4 multiplications it saves
        // per addition (profiles/r01_multiplier_variants.md).  Kept as a tested option for larger inversion batches.
        max_rounds_ = e && *e ? atoi(e) : 0;
        // two point buffers of one point per entry; only for fixed-base engines and while they stay within 16 GiB
        if (cfg_.fixed && max_rounds_ > 0 && entries_max_ * 96 * 2 <= (16ull << 30)) {
            aff_buf_[0] = dev_alloc<uint8_t>(entries_max_ * 96);
            aff_buf_[1] = dev_alloc<uint8_t>(entries_max_ * 96);
        }
    }
    if (points) {
        B200_CUDA_CHECK(cudaMemcpyAsync(table_, points, table_points * 96, host_points ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, stream));
        if (cfg_.fixed && cfg_.W > 1) {
            k_build_rows<<<div_up(table_points, 128), 128, 0, stream>>>((uint8_t*)table_, table_points, cfg_.W, cfg_.c, cfg_.c0);
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
    cudaFree(pair_base_); cudaFree(aff_buf_[0]); cudaFree(aff_buf_[1]);
    if (owns_table_) cudaFree(table_);
    cudaFree(counts_); cudaFree(offsets_); cudaFree(cursor_); cudaFree(task_base_); cudaFree(entries_);
    cudaFree(sorted_tasks_); cudaFree(size_hist_); cudaFree(scan_tmp_); cudaFree(partials_); cudaFree(chunk_sums_);
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
        const size_t fill = (size_t)148 * 4 * 32 * 2;   // two warps on every SM sub-partition
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
                                                                 mont, counts_, nullptr, (size_t)cfg_.bases_period, cfg_.n, lo);
            launches++;
        }
    } else {
        k_digits<false><<<div_up(total, 256), 256, 0, st>>>((const uint4*)scalars_dev, npoints, row_stride, total, c, c0, W, nb_,
                                                            cfg_.fixed, mont, counts_, nullptr, (size_t)cfg_.bases_period, cfg_.n, 0);
        launches++;
    }
    // 2 offsets (and a working copy for the scatter cursors), task bases
    launches += scan_exclusive(counts_, nkeys, 0, offsets_, cursor_, scan_tmp_, st);
    launches += scan_exclusive(counts_, nkeys, (uint32_t)L, task_base_, nullptr, scan_tmp_, st);
    // 3 scatter
    k_digits<true><<<div_up(total, 256), 256, 0, st>>>((const uint4*)scalars_dev, npoints, row_stride, total, c, c0, W, nb_,
                                                       cfg_.fixed, mont, cursor_, entries_, (size_t)cfg_.bases_period, cfg_.n, 0);
    launches++;
    // 5a batch-affine rounds (FIXED engines with room for the two point buffers): halve every bucket's list R times
    const bool prof = profiling_ && prof_count_ < kProfSlots;
    if (prof) {
        for (int k = 0; k < 2; k++)
            if (!prof_ev_[2 * prof_count_ + k]) B200_CUDA_CHECK(cudaEventCreate(&prof_ev_[2 * prof_count_ + k]));
        B200_CUDA_CHECK(cudaEventRecord(prof_ev_[2 * prof_count_], st));
    }
    int rounds = 0;
    const uint8_t* reduced = nullptr;  // buffer holding the current point lists once rounds > 0
    if (aff_buf_[0]) {
        // expected entries per bucket decides how many halvings pay: stop while a round still fills the machine
        double per_bucket = (double)total * W / (double)nkeys;
        size_t pairs = (size_t)((double)total * W / 2);
        while (per_bucket >= 8.0 && pairs >= (size_t)kAffK * 32768 && rounds < max_rounds_) {
            rounds++;
            per_bucket /= 2;
            pairs /= 2;
        }
        size_t pair_bound = total * W / 2;
        for (int r = 0; r < rounds; r++) {
            launches += scan_exclusive(counts_, nkeys, 0x80000002u, pair_base_, nullptr, scan_tmp_, st);
            const uint8_t* in = r == 0 ? nullptr : aff_buf_[(r - 1) & 1];
            uint8_t* out = aff_buf_[r & 1];
            size_t threads = pair_bound / kAffK + 1;
            k_affine_round<<<div_up(threads, 128), 128, 0, st>>>((const uint8_t*)table_, entries_, in, out, offsets_, counts_, pair_base_,
                                                                nkeys, r == 0);
            k_affine_leftover<<<div_up(nkeys, 256), 256, 0, st>>>((const uint8_t*)table_, entries_, in, out, offsets_, counts_, nkeys,
                                                                 r == 0);
            launches += 2;
            pair_bound = pair_bound / 2 + nkeys;
            reduced = out;
        }
        if (rounds) launches += scan_exclusive(counts_, nkeys, (uint32_t)L, task_base_, nullptr, scan_tmp_, st);
    }
    // 4 tasks sorted by length
    k_task_hist<<<div_up(nkeys, 256), 256, (L + 1) * sizeof(uint32_t), st>>>(counts_, nkeys, L, size_hist_);
    k_task_bases<<<1, 32, 0, st>>>(size_hist_, L);
    k_task_emit<<<div_up(nkeys, 256), 256, 0, st>>>(counts_, offsets_, task_base_, nkeys, L, size_hist_, sorted_tasks_);
    launches += 3;
    // 5 accumulate: grid sized for the worst case, surplus threads exit on the device-side task count
    size_t tasks_bound = std::min(tasks_max_, total * W / L + nkeys + 1);
    static const bool acc_call = getenv("B200_ACC_CALL") && atoi(getenv("B200_ACC_CALL"));
    static const bool acc_prefetch = !getenv("B200_ACC_PREFETCH") || atoi(getenv("B200_ACC_PREFETCH"));  // default on: -2.2% at 2^20
    if (reduced)
        k_accumulate_direct<<<div_up(tasks_bound, kAccThreads), kAccThreads, 0, st>>>(reduced, sorted_tasks_, task_base_ + nkeys,
                                                                                     (uint8_t*)partials_);
    else if (acc_call)
        k_accumulate_call<<<div_up(tasks_bound, kAccThreads), kAccThreads, 0, st>>>((const uint8_t*)table_, entries_, sorted_tasks_,
                                                                                   task_base_ + nkeys, (uint8_t*)partials_);
    else {
        static const int acc_occ = getenv("B200_ACC_OCC") ? atoi(getenv("B200_ACC_OCC")) : 3;
        auto kern = acc_prefetch ? (acc_occ == 4 ? k_accumulate<true, 4> : k_accumulate<true, 3>)
                                 : (acc_occ == 4 ? k_accumulate<false, 4> : k_accumulate<false, 3>);
        kern<<<div_up(tasks_bound, kAccThreads), kAccThreads, 0, st>>>((const uint8_t*)table_, entries_, sorted_tasks_,
                                                                       task_base_ + nkeys, (uint8_t*)partials_);
    }
    if (prof) {
        B200_CUDA_CHECK(cudaEventRecord(prof_ev_[2 * prof_count_ + 1], st));
        prof_count_++;
    }
    launches++;
    // 6 reduce
    const uint32_t warp_min = (double)total * W / (double)nkeys <= 0.5 * L ? 4 : 32;
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
        k_group_finish<<<(unsigned)groups, 128, 0, st>>>((const uint8_t*)chunk_sums_, ap, nullptr, (uint8_t*)out_dev, marg_r, kf);
        launches += 1;
    } else {
        k_group_finish<<<(unsigned)groups, 128, 0, st>>>((const uint8_t*)chunk_sums_, ap, (uint8_t*)group_sums_, nullptr, marg_r, kf);
        k_horner<<<1, 32, 0, st>>>((const uint8_t*)group_sums_, W, c, (uint8_t*)out_dev);
        launches += 2;
    }
    B200_LAUNCH_CHECK();
    launches_ = launches;
    last_nkeys_ = nkeys;
}

}  // namespace b200
