// fk20_direct.cu -- the 128 lincombs of 64 fixed points per blob in compute_fk20_proofs (g1_lincomb_batch over
// x_ext_fft_columns, kzg/src/das.rs:676-680; the reference precomputes BGMW tables for them, kzg/src/msm/bgmw.rs:306-380)
// as DIRECT table lookups instead of a bucket pass.
//
// With only 64 points per lincomb and 8192 lincombs per 64-blob batch the bucket method spends as much time sorting and
// reducing 8192 tiny bucket sets (one CTA each, ~45 % of the MSM stage, profiles/r01_verify.md) as adding points.  HBM is
// large, so every multiple is tabulated: for column point P (8192 of them), window j < 32 and digit d = 1..128 the table
// holds d * 2^(8j) * P in affine form -- 8192 * 32 * 128 * 96 B = 3 GiB.  A lincomb is then the plain sum of 64 * 32
// looked-up points: one warp per lincomb, lane j owns window j (Booth digits d_j = byte_j + bit_{8j-1} - 256 bit_{8j+7}
// need no carry chain), 64 mixed additions per lane with the next entry's gather in flight, one warp tree at the end.
// No sort, no buckets, no per-lincomb reduction kernels.
#include "eip4844.cuh"

#include <algorithm>
#include <cstdlib>
#include "g1.cuh"
#include "g1_quad.cuh"
#include "util.cuh"
#include "warp_inverse.cuh"

namespace b200 {

static constexpr int kDPts = 8192, kDW = 32, kDDigits = 128, kDCell = 64, kDChunk = 16;

size_t fk_direct_table_bytes() { return (size_t)kDPts * kDW * kDDigits * 96; }
size_t direct_table_bytes(size_t npts) { return npts * kDW * kDDigits * 96; }

// rows: the fixed-base rows of the MSM engine, row j = 2^(8j) * P_pt (affine), row-major [32][8192].
// One thread per (pt, j): the 128 multiples by repeated mixed addition, converted to affine sixteen at a time with one
// field inversion per warp (Montgomery's trick inside the thread, warp_inverse across the lanes).
__global__ void __launch_bounds__(128) k_fk_direct_build(const uint8_t* __restrict__ rows, uint8_t* __restrict__ table, size_t npts) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // grid covers exactly npts * 32 threads (npts % 4 == 0)
    const size_t pt = t / kDW, j = t % kDW;
    cc::affine_t base = cc::load_affine(rows + (j * npts + pt) * 96);
    uint8_t* out = table + ((pt * kDW + j) * kDDigits) * 96;
    cc::xyzz_t acc = cc::affine_to_xyzz(base);
    for (int c0 = 0; c0 < kDDigits; c0 += kDChunk) {
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
void launch_fk_direct_build(const void* rows, void* table, cudaStream_t st) { launch_direct_build(rows, table, kDPts, st); }
// rows: [32][npts] fixed-base rows for c = 8 (96-byte stride); table: npts x 32 x 128 affine points
void launch_direct_build(const void* rows, void* table, size_t npts, cudaStream_t st) {
    if (npts % 4) throw CudaError(-1, "direct table: point count must be a multiple of 4");
    k_fk_direct_build<<<(unsigned)(npts * kDW / 128), 128, 0, st>>>((const uint8_t*)rows, (uint8_t*)table, npts);
    B200_LAUNCH_CHECK();
}

__device__ __forceinline__ const uint8_t* fk_entry(const uint8_t* table, size_t row, int i, int lane, int d) {
    int mag = d < 0 ? -d : d;
    return table + ((((row * kDCell + i) * kDW + lane) * kDDigits) + (mag ? mag - 1 : 0)) * 96;
}
// one warp per lincomb v (= blob * 128 + row): out[v] = sum_i scalars[v][i] * column[row][i]
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
template <class AR, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 384 / (32 * WARPS)) k_fk_direct_lincomb(const uint8_t* __restrict__ scalars,
                                                                                   const uint8_t* __restrict__ table,
                                                                                   uint8_t* __restrict__ out_jac, int nvec) {
    const int lane = threadIdx.x & 31;
    const size_t v = (size_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (v >= (size_t)nvec) return;                       // whole warps leave together
    const size_t row = v % 128;
    const uint32_t* sc = reinterpret_cast<const uint32_t*>(scalars + v * kDCell * 32);
    // Booth digit of window `lane` of scalar i: byte + (bit below) - 256 * (top bit of the byte)
    auto digit = [&](int i) -> int {
        uint32_t w = sc[i * 8 + (lane >> 2)];
        uint32_t byte = (w >> ((lane & 3) * 8)) & 0xffu;
        uint32_t below = lane == 0 ? 0u : ((lane & 3) ? (w >> ((lane & 3) * 8 - 1)) & 1u : sc[i * 8 + (lane >> 2) - 1] >> 31);
        return (int)byte + (int)below - (int)((byte >> 7) << 8);
    };
    typename AR::xyzz acc = AR::xyzz::inf();
    int d = digit(0);
    typename AR::affine p = AR::load(fk_entry(table, row, 0, lane, d));
#pragma unroll 1
    for (int i = 0; i < kDCell; i++) {
        typename AR::affine cur = p;
        const int cd = d;
        if (i + 1 < kDCell) {
            d = digit(i + 1);
            p = AR::load(fk_entry(table, row, i + 1, lane, d));
        }
        if (cd == 0) cur = typename AR::affine{AR::fp::zero(), AR::fp::zero()};   // adding infinity: same instruction stream
        cur.y = cur.y.cneg(cd < 0);
        AR::add(acc, cur);
    }
    // the tree sum runs on the inlined-multiplier types (same memory layout)
    xyzz_t a2;
#pragma unroll
    for (int k = 0; k < 12; k++) { a2.x.v[k] = acc.x.v[k]; a2.y.v[k] = acc.y.v[k]; a2.zzz.v[k] = acc.zzz.v[k]; a2.zz.v[k] = acc.zz.v[k]; }
    xyzz_t total = warp_sum_xyzz(a2);
    if (lane == 0) store_jac(out_jac + v * 144, xyzz_to_jac(total));
}
void launch_fk_direct_lincomb(const void* scalars, const void* table, void* out_jac, int nvec, cudaStream_t st) {
    // measured (scripts/fk20_timing.py, proofs of 64 blobs): inlined multiplier, 4 warps per CTA 19.5 ms; behind a call 18.9;
    // inlined, 2 warps 19.3; behind a call, 2 warps per CTA 18.8 (default: smaller code, finer-grained last wave)
    static const int variant = getenv("B200_FK20_LINCOMB") ? atoi(getenv("B200_FK20_LINCOMB")) : 3;
    const uint8_t *s8 = (const uint8_t*)scalars, *t8 = (const uint8_t*)table;
    uint8_t* o8 = (uint8_t*)out_jac;
    if (variant == 1) k_fk_direct_lincomb<ArCall, 4><<<div_up(nvec, 4), 128, 0, st>>>(s8, t8, o8, nvec);
    else if (variant == 2) k_fk_direct_lincomb<ArInline, 2><<<div_up(nvec, 2), 64, 0, st>>>(s8, t8, o8, nvec);
    else if (variant == 3) k_fk_direct_lincomb<ArCall, 2><<<div_up(nvec, 2), 64, 0, st>>>(s8, t8, o8, nvec);
    else k_fk_direct_lincomb<ArInline, 4><<<div_up(nvec, 4), 128, 0, st>>>(s8, t8, o8, nvec);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// Small batches of 4096-term MSMs over the Lagrange points (blob_to_kzg_commitment / compute_kzg_proof of 1 .. 32 blobs) by
// the same direct lookups.  The bucket pipeline is a chain of ~20 dependent, mostly tiny kernels whose reduction tail
// (bucket combine, marginal sums, weighted sums) costs ~0.45 ms however little work there is: one blob takes 0.55 ms, and
// that latency is what concurrent single-blob callers queue behind (coalesce.cuh).  Direct form: a warp takes P consecutive
// points, lane j owns window j (8-bit Booth digits need no carry chain), P mixed additions per lane with the next gather in
// flight, one warp tree; a second kernel folds the 4096 / P partial sums of every vector.  No sort, no buckets: two
// launches, chain length P + 2 trees.  table: 4096 x 32 x 128 affine points (1.5 GiB, built once per settings object).
// FUSED: the CTA that finishes a vector's last partial sum (a counter per vector) also folds the vector's m partials and
// writes the 48-byte compressed result -- one launch per batch instead of three (direct sums, fold, compression).
template <class AR, bool FUSED>
__global__ void __launch_bounds__(128) k_direct_msm_partial(const uint8_t* __restrict__ scalars, const uint8_t* __restrict__ table,
                                                            uint8_t* __restrict__ partials, int npts, int P, size_t nwarps,
                                                            unsigned* __restrict__ counters, uint8_t* __restrict__ out48) {
    __shared__ __align__(16) uint8_t sh[4 * 192];
    __shared__ int sh_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t w = (size_t)blockIdx.x * 4 + wid;             // nwarps is a multiple of 4: whole CTAs are live
    const size_t per_vec = (size_t)npts / P, v = w / per_vec, p0 = (w % per_vec) * P;
    const uint32_t* sc = reinterpret_cast<const uint32_t*>(scalars + (v * npts + p0) * 32);
    auto digit = [&](int i) -> int {                          // Booth digit of window `lane` of scalar i (see k_fk_direct_lincomb)
        uint32_t wd = sc[i * 8 + (lane >> 2)];
        uint32_t byte = (wd >> ((lane & 3) * 8)) & 0xffu;
        uint32_t below = lane == 0 ? 0u : ((lane & 3) ? (wd >> ((lane & 3) * 8 - 1)) & 1u : sc[i * 8 + (lane >> 2) - 1] >> 31);
        return (int)byte + (int)below - (int)((byte >> 7) << 8);
    };
    auto entry = [&](int i, int d) -> const uint8_t* {
        const int mag = d < 0 ? -d : d;
        return table + ((((p0 + i) * kDW + lane) * kDDigits) + (mag ? mag - 1 : 0)) * 96;
    };
    typename AR::xyzz acc = AR::xyzz::inf();
    int d = digit(0);
    typename AR::affine p = AR::load(entry(0, d));
#pragma unroll 1
    for (int i = 0; i < P; i++) {
        typename AR::affine cur = p;
        const int cd = d;
        if (i + 1 < P) {
            d = digit(i + 1);
            p = AR::load(entry(i + 1, d));
        }
        if (cd == 0) cur = typename AR::affine{AR::fp::zero(), AR::fp::zero()};
        cur.y = cur.y.cneg(cd < 0);
        AR::add(acc, cur);
    }
    xyzz_t a2;
#pragma unroll
    for (int k = 0; k < 12; k++) { a2.x.v[k] = acc.x.v[k]; a2.y.v[k] = acc.y.v[k]; a2.zzz.v[k] = acc.zzz.v[k]; a2.zz.v[k] = acc.zz.v[k]; }
    // warp tree, then the CTA's four warp sums on four quads of warp 0: one partial per CTA (4 P points)
    fp_t q = seg_sum_quad(a2, 32);
    if (lane < 4) store_field(sh + wid * 192 + quad_store_offset(), q);
    __syncthreads();
    if (wid == 0) {
        fp_t c = lane < 16 ? load_field<fp_t>(sh + (lane >> 2) * 192 + quad_store_offset()) : fp_t::zero();
        fp_t t = quad_tree(c, 16);
        if (lane < 4) store_field(partials + (size_t)blockIdx.x * 192 + quad_store_offset(), t);
    }
    if (!FUSED) return;
    // last CTA of this vector?  (partials of a vector are contiguous: m = npts / (4 P) CTAs per vector)
    const int m = npts / (4 * P);
    const size_t vec = (size_t)blockIdx.x / m;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(&counters[vec], 1u);
        sh_last = done == (unsigned)m - 1;
        if (sh_last) counters[vec] = 0;                       // ready for the next launch on this lane
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    // fold the m partials (16 <= m <= 128) with the CTA's 128 threads, exactly as k_direct_msm_reduce does
    const uint8_t* base = partials + vec * m * 192;
    xyzz_t a;
    if ((int)threadIdx.x < m) {
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
    __syncthreads();                                          // sh is reused below
    fp_t q2 = seg_sum_quad(a, 32);
    if (lane < 4) store_field(sh + wid * 192 + quad_store_offset(), q2);
    __syncthreads();
    if (wid != 0) return;
    fp_t c2 = lane < 16 ? load_field<fp_t>(sh + (lane >> 2) * 192 + quad_store_offset()) : fp_t::zero();
    fp_t tot = quad_tree(c2, 16);                             // quad 0: (X, Y, ZZ, ZZZ) of the vector's sum
    // compressed form (blst_p1_compress): x = X / ZZ, y = Y / ZZZ with 1/ZZ = ZZZ^-2 ZZ^2; one inversion, uniform over the warp
    const fp_t zz = shfl_idx_fp(tot, 2), zzz = shfl_idx_fp(tot, 3);
    const bool inf = zz.is_zero();
    const fp_t izzz = (inf ? fp_t::one() : zzz).inverse();
    const fp_t izz = izzz.sqr() * zz.sqr();
    const fp_t coord = lane == 0 ? tot * izz : tot * izzz;    // lane 0: x, lane 1: y
    const fp_t yv = shfl_idx_fp(coord, 1);
    if (lane == 0) {
        cc::affine_t r;
#pragma unroll
        for (int k = 0; k < 12; k++) { r.x.v[k] = inf ? 0u : coord.v[k]; r.y.v[k] = inf ? 0u : yv.v[k]; }
        cc::affine_compress(out48 + vec * 48, r);
    }
}
// one CTA of 32 .. 256 threads per vector: m = 4096 / (4 P) partial sums (16 <= m <= 128, a power of two) -> Jacobian result
__global__ void __launch_bounds__(256) k_direct_msm_reduce(const uint8_t* __restrict__ partials, int m, uint8_t* __restrict__ out_jac) {
    __shared__ __align__(16) uint8_t sh[8 * 192];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint8_t* base = partials + (size_t)blockIdx.x * m * 192;
    xyzz_t a = (int)threadIdx.x < m ? load_xyzz(base + (size_t)threadIdx.x * 192) : xyzz_t::inf();
    for (int i = threadIdx.x + blockDim.x; i < m; i += blockDim.x) {
        xyzz_t b = load_xyzz(base + (size_t)i * 192);
        xyzz_add(a, b);
    }
    fp_t q = seg_sum_quad(a, 32);
    if (lane < 4) store_field(sh + wid * 192 + quad_store_offset(), q);
    __syncthreads();
    if (wid == 0) {
        fp_t c = (lane >> 2) < nw ? load_field<fp_t>(sh + (lane >> 2) * 192 + quad_store_offset()) : fp_t::zero();
        fp_t t = quad_tree(c, 32);                           // the (at most 8) warp sums, one quad each
        fp_t j = t * shfl_xor_fp(t, 2);                      // Jacobian (X ZZ, Y ZZZ, ZZ); infinity stays all-zero
        if (lane < 2) store_field(out_jac + (size_t)blockIdx.x * 144 + lane * 48, j);
        if (lane == 2) store_field(out_jac + (size_t)blockIdx.x * 144 + 96, t);
    }
}
static int direct_p(int nvec, int npts) {
    // P points per warp: as few as keep one wave of warps on the machine (148 SMs x 12 warps), between 8 and 64;
    // four warps per CTA leave one partial sum per 4 P points
    int P = 8;
    while (P < 64 && (size_t)nvec * npts / P > 148 * 12) P <<= 1;
    return P;
}
// scalars: nvec x npts canonical little-endian 32-byte scalars; partials: workspace of nvec * npts / 32 XYZZ points
void launch_direct_msm(const void* scalars, const void* table, void* partials, void* out_jac, int nvec, int npts, cudaStream_t st) {
    const int P = direct_p(nvec, npts);
    const size_t nwarps = (size_t)nvec * npts / P;
    const int m = npts / (4 * P);                              // partial sums per vector: 128 .. 16
    k_direct_msm_partial<ArCall, false><<<(unsigned)(nwarps / 4), 128, 0, st>>>((const uint8_t*)scalars, (const uint8_t*)table,
                                                                             (uint8_t*)partials, npts, P, nwarps, nullptr, nullptr);
    k_direct_msm_reduce<<<nvec, std::max(32, std::min(m, 256)), 0, st>>>((const uint8_t*)partials, m, (uint8_t*)out_jac);
    B200_LAUNCH_CHECK();
}
// the same sums, written as 48-byte compressed points by ONE launch; counters: nvec zero-initialised words (left zero)
void launch_direct_msm_compressed(const void* scalars, const void* table, void* partials, unsigned* counters, uint8_t* out48, int nvec,
                                  int npts, cudaStream_t st) {
    const int P = direct_p(nvec, npts);
    const size_t nwarps = (size_t)nvec * npts / P;
    k_direct_msm_partial<ArCall, true><<<(unsigned)(nwarps / 4), 128, 0, st>>>((const uint8_t*)scalars, (const uint8_t*)table,
                                                                            (uint8_t*)partials, npts, P, nwarps, counters, out48);
    B200_LAUNCH_CHECK();
}

}  // namespace b200
