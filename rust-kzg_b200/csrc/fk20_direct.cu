// fk20_direct.cu -- fixed-base lincombs as DIRECT table lookups instead of a bucket pass:
//   * the 128 lincombs of 64 fixed points per blob in compute_fk20_proofs (g1_lincomb_batch over x_ext_fft_columns,
//     kzg/src/das.rs:676-680; the reference precomputes BGMW tables for them, kzg/src/msm/bgmw.rs:306-380),
//   * the 4096-term MSMs over the Lagrange points of blob_to_kzg_commitment / compute_kzg_proof
//     (g1_lincomb_fast over g1_values_lagrange_brp, kzg/src/eip_4844.rs:463-476, :520-524).
//
// HBM is large, so EVERY signed-digit multiple of every window of every fixed point is tabulated: for point P, window
// j < W = ceil(256 / c) and digit d = 1 .. 2^(c-1) the table holds d * 2^(cj) * P in affine form, item-major:
//     entry(P, j, d) = table[((P * W + j) << (c - 1)) + d - 1]            (96 bytes each)
// A lincomb is then the plain sum of (points x W) looked-up points.  Booth digits d_j = raw_j + bit_{cj-1} - 2^c top_j
// need no carry chain, so every (point, window) ITEM is independent: a team of threads strides over the items of a vector,
// each thread sums its share by mixed additions with the next entry's gather in flight, and trees fold the team.
// No sort, no buckets, no per-lincomb reduction kernels; the work per vector is points x W additions, which is what makes
// a WIDE window worth its table: c = 13 costs 20 additions per point instead of the 32 of c = 8 for
// 4096 x 20 x 4096 x 96 B = 30 GiB (Lagrange points) and 60 GiB (FK20 columns) of the 180 GB.
#include "eip4844.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "g1.cuh"
#include "g1_quad.cuh"
#include "util.cuh"
#include "warp_inverse.cuh"

namespace b200 {

static constexpr int kDChunk = 16;

int direct_windows(int c) { return (256 + c - 1) / c; }   // scalars are < 2^255: the top window never carries out
size_t direct_table_bytes(size_t npts, int c) { return (npts * direct_windows(c) << (c - 1)) * 96; }

// rows: the fixed-base rows of an MSM engine with the same window width, row j = 2^(cj) * P_pt (affine), row-major
// [W][npts] with row_stride bytes per point.  One thread per item (pt, j): the 2^(c-1) multiples by repeated mixed
// addition, converted to affine sixteen at a time with one field inversion per warp (Montgomery's trick inside the
// thread, warp_inverse across the lanes).
__global__ void __launch_bounds__(128) k_direct_build(const uint8_t* __restrict__ rows, size_t row_stride, uint8_t* __restrict__ table,
                                                      size_t npts, int W, int c) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // grid covers exactly npts * W threads
    const size_t pt = t / W, j = t % W;
    const int D = 1 << (c - 1);
    cc::affine_t base = cc::load_affine(rows + (j * npts + pt) * row_stride);
    uint8_t* out = table + (t << (c - 1)) * 96;
    cc::xyzz_t acc = cc::affine_to_xyzz(base);
#pragma unroll 1
    for (int c0 = 0; c0 < D; c0 += kDChunk) {
        cc::xyzz_t pts[kDChunk];
        cc::fp_t pre[kDChunk];
        cc::fp_t run = cc::fp_t::one();
        for (int e = 0; e < kDChunk; e++) {
            pts[e] = acc;
            pre[e] = run;
            run = run * (acc.is_inf() ? cc::fp_t::one() : acc.zzz);
            cc::xyzz_add_affine(acc, base);
        }
        cc::fp_t inv = warp_inverse(run);
        for (int e = kDChunk - 1; e >= 0; e--) {
            cc::affine_t a{cc::fp_t::zero(), cc::fp_t::zero()};
            if (!pts[e].is_inf()) {
                cc::fp_t izzz = inv * pre[e];
                inv = inv * pts[e].zzz;
                cc::fp_t s = pts[e].zz * izzz;           // 1/ZZ = (ZZ / ZZZ)^2 because ZZ^3 = ZZZ^2
                a.x = pts[e].x * s.sqr();
                a.y = pts[e].y * izzz;
            }
            cc::store_affine(out + (size_t)(c0 + e) * 96, a);
        }
    }
}
void launch_direct_build(const void* rows, size_t row_stride, void* table, size_t npts, int c, cudaStream_t st) {
    const int W = direct_windows(c);
    if (c < 5 || c > 16 || (npts * W) % 128) throw CudaError(-1, "direct table: unsupported shape");
    k_direct_build<<<(unsigned)(npts * W / 128), 128, 0, st>>>((const uint8_t*)rows, row_stride, (uint8_t*)table, npts, W, c);
    B200_LAUNCH_CHECK();
}

// AR selects the arithmetic instantiation: b200:: (multiplier inlined) or b200::cl:: (multiplier behind a call, a loop body
// that fits the instruction cache -- ncu showed "no instruction" as the second largest stall of the inlined form here).
struct ArInline {
    typedef b200::fp_t fp;
    typedef b200::affine_t affine;
    typedef b200::xyzz_t xyzz;
    static __device__ __forceinline__ affine load(const void* p) { return b200::load_affine(p); }
    static __device__ __forceinline__ void add(xyzz& acc, const affine& p) { b200::xyzz_add_affine(acc, p); }
};
struct ArCall {
    typedef b200::cl::fp_t fp;
    typedef b200::cl::affine_t affine;
    typedef b200::cl::xyzz_t xyzz;
    static __device__ __forceinline__ affine load(const void* p) { return b200::cl::load_affine(p); }
    static __device__ __forceinline__ void add(xyzz& acc, const affine& p) { b200::cl::xyzz_add_affine(acc, p); }
};

// Booth digit of window j (width c) of a canonical little-endian scalar: raw + (bit below the window) - 2^c * (top bit)
__device__ __forceinline__ int booth_digit(const uint32_t* __restrict__ sc, int j, int c) {
    const int o = c * j - 1;
    uint64_t v;
    if (o < 0) {
        v = (uint64_t)sc[0] << 1;
    } else {
        const int w = o >> 5;
        const uint32_t lo = sc[w], hi = w + 1 < 8 ? sc[w + 1] : 0u;
        v = (((uint64_t)hi << 32) | lo) >> (o & 31);
    }
    const uint32_t u = (uint32_t)v & ((2u << c) - 1u);
    return (int)(u >> 1) + (int)(u & 1u) - (int)((u >> c) << c);
}
// the sum of P items first, first + step, ... of one vector: sc = the vector's scalars, tab = entry 0 of its item 0
template <class AR>
__device__ __forceinline__ xyzz_t direct_chain(const uint32_t* __restrict__ sc, const uint8_t* __restrict__ tab, uint32_t first,
                                               uint32_t step, int P, int W, int c) {
    auto fetch = [&](uint32_t it, int& d) -> const uint8_t* {
        const uint32_t pt = it / (uint32_t)W, j = it - pt * (uint32_t)W;
        d = booth_digit(sc + pt * 8, (int)j, c);
        const int mag = d < 0 ? -d : d;
        return tab + (((size_t)it << (c - 1)) + (size_t)(mag ? mag - 1 : 0)) * 96;
    };
    typename AR::xyzz acc = AR::xyzz::inf();
    int d;
    uint32_t it = first;
    typename AR::affine p = AR::load(fetch(it, d));
#pragma unroll 1
    for (int i = 0; i < P; i++) {
        typename AR::affine cur = p;
        const int cd = d;
        if (i + 1 < P) {
            it += step;
            p = AR::load(fetch(it, d));
        }
        if (cd == 0) cur = typename AR::affine{AR::fp::zero(), AR::fp::zero()};   // adding infinity: same instruction stream
        cur.y = cur.y.cneg(cd < 0);
        AR::add(acc, cur);
    }
    xyzz_t a2;                                                // the trees run on the inlined-multiplier types (same layout)
#pragma unroll
    for (int k = 0; k < 12; k++) { a2.x.v[k] = acc.x.v[k]; a2.y.v[k] = acc.y.v[k]; a2.zzz.v[k] = acc.zzz.v[k]; a2.zz.v[k] = acc.zz.v[k]; }
    return a2;
}

// FK20: one warp per lincomb v (= blob * period + row): out[v] = sum_i scalars[v][i] * column[row][i], npv points each
template <class AR, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 384 / (32 * WARPS)) k_direct_lincomb(const uint8_t* __restrict__ scalars,
                                                                                const uint8_t* __restrict__ table,
                                                                                uint8_t* __restrict__ out_jac, int nvec, int period,
                                                                                int npv, int W, int c) {
    const int lane = threadIdx.x & 31;
    const size_t v = (size_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (v >= (size_t)nvec) return;                       // whole warps leave together
    const size_t row = v % period;
    const uint32_t* sc = reinterpret_cast<const uint32_t*>(scalars + v * npv * 32);
    const uint8_t* tab = table + ((row * npv * W) << (c - 1)) * 96;
    xyzz_t a2 = direct_chain<AR>(sc, tab, lane, 32, npv * W / 32, W, c);
    xyzz_t total = warp_sum_xyzz(a2);
    if (lane == 0) store_jac(out_jac + v * 144, xyzz_to_jac(total));
}
void launch_direct_lincomb(const void* scalars, const void* table, void* out_jac, int nvec, int period, int npv, int c, cudaStream_t st) {
    // measured (scripts/fk20_timing.py, proofs of 64 blobs, c = 8): inlined multiplier, 4 warps per CTA 19.5 ms; behind a call
    // 18.9; inlined, 2 warps 19.3; behind a call, 2 warps per CTA 18.8 (default: smaller code, finer-grained last wave)
    static const int variant = getenv("B200_FK20_LINCOMB") ? atoi(getenv("B200_FK20_LINCOMB")) : 3;
    const int W = direct_windows(c);
    if ((npv * W) % 32) throw CudaError(-1, "direct lincomb: items per vector must fill whole warps");
    const uint8_t *s8 = (const uint8_t*)scalars, *t8 = (const uint8_t*)table;
    uint8_t* o8 = (uint8_t*)out_jac;
    if (variant == 1) k_direct_lincomb<ArCall, 4><<<div_up(nvec, 4), 128, 0, st>>>(s8, t8, o8, nvec, period, npv, W, c);
    else if (variant == 2) k_direct_lincomb<ArInline, 2><<<div_up(nvec, 2), 64, 0, st>>>(s8, t8, o8, nvec, period, npv, W, c);
    else if (variant == 3) k_direct_lincomb<ArCall, 2><<<div_up(nvec, 2), 64, 0, st>>>(s8, t8, o8, nvec, period, npv, W, c);
    else k_direct_lincomb<ArInline, 4><<<div_up(nvec, 4), 128, 0, st>>>(s8, t8, o8, nvec, period, npv, W, c);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// Batches of npts-term MSMs over the Lagrange points (blob_to_kzg_commitment / compute_kzg_proof) by the same direct
// lookups.  The bucket pipeline is a chain of ~20 dependent, mostly tiny kernels whose reduction tail (bucket combine,
// marginal sums, weighted sums) costs ~0.45 ms however little work there is, and that latency is what concurrent
// single-blob callers queue behind (coalesce.cuh).  Direct form: the items of a vector are cut into SLICES of 128 (thread t
// of a CTA takes item 128 s + t of slice s), and the nvec * spv slices of the whole batch are dealt out evenly and
// contiguously to the G CTAs of ONE wave (two CTAs per SM): every SM gets the same number of additions whatever the batch
// size; a CTA whose range crosses a vector boundary works on both vectors in turn.  Per (CTA, vector) piece: the chain of
// additions, a warp tree and a tree over the CTA's four warps leave one partial sum; the CTA that finishes a vector's last
// partial (a counter per vector) folds its partials, inverts ZZZ and writes the 48-byte compressed result.  One launch per
// batch; chain length + three trees + one inversion.
__device__ __forceinline__ uint32_t direct_cta_of(uint32_t slice, uint32_t S, uint32_t G) {   // the CTA whose range holds `slice`
    return (uint32_t)((((uint64_t)slice + 1) * G + S - 1) / S) - 1;
}
#ifndef B200_DIRECT_MINB
#define B200_DIRECT_MINB 2   // measured: 3 CTAs per SM (168 registers, spills in the chain) 2.39 ms per 64 blobs against 2.31
#endif
template <class AR>
__global__ void __launch_bounds__(128, B200_DIRECT_MINB) k_direct_msm(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ table,
                                                    uint8_t* __restrict__ partials, int npts, int W, int c, uint32_t spv, uint32_t S,
                                                    unsigned* __restrict__ counters, uint8_t* __restrict__ out48,
                                                    uint8_t* __restrict__ out_jac, unsigned long long* trace, int period) {
    unsigned long long ts[8];
    auto stamp = [&](int k) { if (trace) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); ts[k] = t; } };
    stamp(0);
    __shared__ __align__(16) uint8_t sh[4 * 192];
    __shared__ int sh_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t G = gridDim.x;
    const uint32_t b0 = (uint32_t)((uint64_t)blockIdx.x * S / G), b1 = (uint32_t)(((uint64_t)blockIdx.x + 1) * S / G);
#pragma unroll 1
    for (uint32_t vec = b0 / spv; vec * spv < b1; vec++) {
        const uint32_t v0 = vec * spv, s0 = (b0 > v0 ? b0 : v0) - v0, s1 = (b1 < v0 + spv ? b1 : v0 + spv) - v0;
        const uint32_t* sc = reinterpret_cast<const uint32_t*>(scalars + (size_t)vec * npts * 32);
        // period > 1 (FK20): vector v sums over the point block v mod period, whose entries start (v mod period) * npts * W items in
        const uint8_t* tab = table + ((((size_t)(vec % (uint32_t)period) * npts * W) << (c - 1)) * 96);
        xyzz_t a2 = direct_chain<AR>(sc, tab, s0 * 128u + threadIdx.x, 128u, (int)(s1 - s0), W, c);
        stamp(1);
        // warp tree, then the CTA's four warp sums on four quads of warp 0: one partial per (CTA, vector)
        fp_t q = seg_sum_quad(a2, 32);
        stamp(2);
        if (lane < 4) store_field(sh + wid * 192 + quad_store_offset(), q);
        __syncthreads();
        const uint32_t first_cta = direct_cta_of(v0, S, G), m = direct_cta_of(v0 + spv - 1, S, G) - first_cta + 1;
        uint8_t* base = partials + (size_t)vec * 128 * 192;   // m <= 128 slots per vector
        if (wid == 0) {
            fp_t cq = lane < 16 ? load_field<fp_t>(sh + (lane >> 2) * 192 + quad_store_offset()) : fp_t::zero();
            fp_t t = quad_tree(cq, 16);
            if (lane < 4) store_field(base + (size_t)(blockIdx.x - first_cta) * 192 + quad_store_offset(), t);
        }
        stamp(3);
        // last CTA of this vector?
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(&counters[vec], 1u);
            sh_last = done == m - 1;
            if (sh_last) counters[vec] = 0;                   // ready for the next launch on this lane
        }
        __syncthreads();
        if (!sh_last) continue;                               // CTA-uniform
        __threadfence();
        // fold the m <= 128 partials with the CTA's 128 threads
        xyzz_t a;
        if (threadIdx.x < m) {
            const uint4* src = reinterpret_cast<const uint4*>(base + (size_t)threadIdx.x * 192);
            uint4 tmp[12];
#pragma unroll
            for (int k = 0; k < 12; k++) tmp[k] = __ldcg(src + k);   // written by other CTAs: read through L2
            const uint32_t* wv = reinterpret_cast<const uint32_t*>(tmp);
#pragma unroll
            for (int k = 0; k < 12; k++) { a.x.v[k] = wv[k]; a.y.v[k] = wv[12 + k]; a.zzz.v[k] = wv[24 + k]; a.zz.v[k] = wv[36 + k]; }
        } else {
            a = xyzz_t::inf();
        }
        stamp(4);
        fp_t tot = fp_t::zero();
        if (m > 32) {
            fp_t q2 = seg_sum_quad(a, 32);
            if (lane < 4) store_field(sh + wid * 192 + quad_store_offset(), q2);
            __syncthreads();
            if (wid == 0) {
                fp_t c2 = lane < 16 ? load_field<fp_t>(sh + (lane >> 2) * 192 + quad_store_offset()) : fp_t::zero();
                tot = quad_tree(c2, 16);                      // quad 0: (X, Y, ZZ, ZZZ) of the vector's sum
            }
        } else if (wid == 0) {
            tot = seg_sum_quad(a, 32);                        // one warp holds every partial
        }
        stamp(5);
        if (wid == 0 && out_jac) {
            // Jacobian (X ZZ, Y ZZZ, ZZ) as k_group_finish writes it (blst_p1); infinity stays all-zero.  No inversion.
            const fp_t jc = tot * shfl_xor_fp(tot, 2);
            if (lane < 2) store_field(out_jac + (size_t)vec * 144 + lane * 48, jc);
            if (lane == 2) store_field(out_jac + (size_t)vec * 144 + 96, tot);
        } else if (wid == 0) {
            // compressed form (blst_p1_compress): x = X / ZZ, y = Y / ZZZ with 1/ZZ = ZZZ^-2 ZZ^2; one inversion, uniform over the warp
            const fp_t zz = shfl_idx_fp(tot, 2), zzz = shfl_idx_fp(tot, 3);
            const bool inf = zz.is_zero();
            const fp_t izzz = (inf ? fp_t::one() : zzz).inverse();
            stamp(6);
            const fp_t izz = izzz.sqr() * zz.sqr();
            const fp_t coord = lane == 0 ? tot * izz : tot * izzz;    // lane 0: x, lane 1: y
            const fp_t yv = shfl_idx_fp(coord, 1);
            if (lane == 0) {
                cc::affine_t r;
#pragma unroll
                for (int k = 0; k < 12; k++) { r.x.v[k] = inf ? 0u : coord.v[k]; r.y.v[k] = inf ? 0u : yv.v[k]; }
                cc::affine_compress(out48 + (size_t)vec * 48, r);
                stamp(7);
                if (trace && vec == 0)
                    for (int k = 0; k < 8; k++) trace[k] = ts[k];
            }
        }
        __syncthreads();                                      // sh is reused by the next piece
    }
}
// blst_fr (Montgomery) -> canonical little-endian scalars, the form the Booth digits are read from
__global__ void k_fr_from_mont(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_field(out + i * 32, load_field_ro<fr_t>(in + i * 32).from_mont());
}
void launch_fr_from_mont(const void* in, void* out, size_t n, cudaStream_t st) {
    if (!n) return;
    k_fr_from_mont<<<(unsigned)div_up(n, (size_t)256), 256, 0, st>>>((const uint8_t*)in, (uint8_t*)out, n);
    B200_LAUNCH_CHECK();
}
void launch_direct_msm_compressed(const void* scalars, const void* table, void* partials, unsigned* counters, uint8_t* out48, int nvec,
                                  int npts, int c, cudaStream_t st) {
    launch_direct_msm(scalars, table, partials, counters, out48, nullptr, nvec, npts, c, st);
}
// scalars: nvec x npts canonical little-endian 32-byte scalars; partials: workspace of nvec * 128 XYZZ points (192 bytes);
// counters: nvec zero-initialised words (left zero); results: out48 = nvec compressed points, or (out_jac != nullptr) nvec blst_p1
void launch_direct_msm(const void* scalars, const void* table, void* partials, unsigned* counters, uint8_t* out48, uint8_t* out_jac,
                       int nvec, int npts, int c, cudaStream_t st, int period) {
    const int W = direct_windows(c);
    if (((size_t)npts * W) % 128) throw CudaError(-1, "direct MSM: items per vector must fill whole CTAs");
    static const int wave = [] {                              // resident CTAs of this kernel on the whole device
        int dev = 0, sms = 148, per_sm = 2;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_direct_msm<ArCall>, 128, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 2;
        }
        return per_sm * sms;
    }();
    const uint32_t spv = (uint32_t)((size_t)npts * W / 128), S = spv * (uint32_t)nvec;
    // at least five slices per CTA, and no vector cut into more than the 128 pieces its fold handles
    uint32_t G = std::min<uint32_t>(std::min<uint32_t>((uint32_t)wave, 128u * (uint32_t)nvec), std::max<uint32_t>(1u, S / 5));
    auto cta_of = [&](uint32_t slice) { return (uint32_t)((((uint64_t)slice + 1) * G + S - 1) / S) - 1; };
    for (;; G--) {
        uint32_t worst = 0;
        for (uint32_t v = 0; v < (uint32_t)nvec; v++) worst = std::max(worst, cta_of(v * spv + spv - 1) - cta_of(v * spv) + 1);
        if (worst <= 128 || G == 1) break;
    }
    // B200_DIRECT_TRACE=1 (debugging aid): %globaltimer stamps of the CTA that finishes vector 0, printed after a sync
    static unsigned long long* trace = [] {
        unsigned long long* t = nullptr;
        if (getenv("B200_DIRECT_TRACE") && atoi(getenv("B200_DIRECT_TRACE"))) cudaMallocManaged(&t, 8 * sizeof(unsigned long long));
        return t;
    }();
    k_direct_msm<ArCall><<<G, 128, 0, st>>>((const uint8_t*)scalars, (const uint8_t*)table, (uint8_t*)partials, npts, W, c, spv, S, counters,
                                           out48, out_jac, trace, period < 1 ? 1 : period);
    B200_LAUNCH_CHECK();
    if (trace) {
        cudaStreamSynchronize(st);
        fprintf(stderr, "direct trace (ns) nvec=%d c=%d G=%u slices/CTA=%.1f: loop %llu warp-tree %llu cta-tree %llu ticket+load %llu fold %llu inverse %llu compress %llu\n",
                nvec, c, G, (double)S / G, trace[1] - trace[0], trace[2] - trace[1], trace[3] - trace[2], trace[4] - trace[3], trace[5] - trace[4],
                trace[6] - trace[5], trace[7] - trace[6]);
    }
}

}  // namespace b200
