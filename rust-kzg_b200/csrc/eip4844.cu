// eip4844.cu -- device kernels and host driver of the batched EIP-4844 commitment / proof path.  See eip4844.cuh.
#include "eip4844.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "g1.cuh"
#include "g1_quad.cuh"
#include "util.cuh"
#include "warp_inverse.cuh"
#include "wire.cuh"

namespace b200 {

// bytes_to_blob (kzg/src/eip_4844.rs:867-880) = 4096 x Fr::from_bytes (blst/src/types/fr.rs:64-86): big-endian,
// canonical (< r) or the whole blob is rejected.  Writes the canonical little-endian scalar (MSM input) and,
// if poly != nullptr, the Montgomery form (polynomial arithmetic).  One thread per field element, 32 B in, coalesced.
__global__ void __launch_bounds__(256) k_blob_to_fr(const uint8_t* __restrict__ blobs, size_t total, uint8_t* __restrict__ scalars,
                                                    uint8_t* __restrict__ poly, int* __restrict__ status) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    fr_t v;
    load_be32(blobs + gid * 32, v.v);
    if (!lt_r(v.v)) status[gid / kFieldElementsPerBlob] = 1;
    if (scalars) store_field(scalars + gid * 32, v);
    if (poly) store_field(poly + gid * 32, v.to_mont());
}

// z: Fr::from_bytes (reject >= r) or hash_to_bls_field (reduce mod r) -> Montgomery
__global__ void k_z_to_fr(const uint8_t* __restrict__ z_bytes, int n, int reduce, uint8_t* __restrict__ z_out, int* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t v;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint8_t* p = z_bytes + (size_t)i * 32 + 4 * (7 - k);
        v.v[k] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
    }
    if (!reduce && !lt_r(v.v)) status[i] = 1;
    // hash_to_bls_field reduces mod r (blst_fr_from_scalar, blst/src/types/fr.rs:88-107): 2^256 < 3r, so at most two
    // subtractions bring the value into [0, r) before the Montgomery conversion (which needs a reduced operand)
    for (int it = 0; it < 2; it++) {
        if (lt_r(v.v)) break;
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint64_t d = (uint64_t)v.v[k] - FrParams::mod(k) - borrow;
            v.v[k] = (uint32_t)d;
            borrow = (uint32_t)(d >> 63);
        }
    }
    store_field(z_out + (size_t)i * 32, v.to_mont());
}
// Fr::to_bytes (blst/src/types/fr.rs:127-136)
__global__ void k_fr_to_bytes(const uint8_t* __restrict__ fr_mont, int n, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t v = load_field<fr_t>(fr_mont + (size_t)i * 32).from_mont();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t w = v.v[7 - k];
        uint8_t* p = out + (size_t)i * 32 + 4 * k;
        p[0] = w >> 24; p[1] = w >> 16; p[2] = w >> 8; p[3] = w;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// G1 decompression (blst_p1_uncompress via FsG1::from_bytes, blst/src/types/g1.rs:65-87; encoding:
// zkcrypto/bls12_381/src/notes/serialization.rs:1-29).  One thread per point; sqrt = y2^((p+1)/4).
__device__ __forceinline__ bool uncompress_point(const uint8_t* in, cc::affine_t& out) {
    using cc::fp_t;
    uint32_t b0 = in[0];
    uint32_t cflag = b0 >> 7, iflag = (b0 >> 6) & 1, sflag = (b0 >> 5) & 1;
    fp_t x;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const uint8_t* p = in + 4 * (11 - k);
        x.v[k] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
    }
    x.v[11] &= 0x1fffffffu;
    out.x = fp_t::zero();
    out.y = fp_t::zero();
    if (!cflag) return false;
    if (iflag) return !sflag && x.is_zero();
    // x < p ?
    bool lt = false;
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        uint32_t m = FpParams::mod(i);
        if (x.v[i] != m) { lt = x.v[i] < m; break; }
    }
    if (!lt) return false;
    fp_t xm = x.to_mont();
    fp_t four = fp_t::one().dbl().dbl();
    fp_t y2 = xm.sqr() * xm + four;
    const uint32_t E[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                            0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};  // (p+1)/4
    fp_t y = y2.pow_words(E);
    if (y.sqr() != y2) return false;
    if (cc::fp_is_lex_largest(y) != (bool)sflag) y = y.neg();
    out.x = xm;
    out.y = y;
    return true;
}
__global__ void __launch_bounds__(64) k_uncompress_g1(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int* __restrict__ flags, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cc::affine_t a;
    bool ok = uncompress_point(in + (size_t)i * 48, a);
    if (!ok && flags) flags[i] = 1;
    cc::store_affine(out + (size_t)i * 96, a);
}
void launch_uncompress_g1(const uint8_t* in48_dev, void* affine_out_dev, int* flags_dev, int n, cudaStream_t st) {
    if (n <= 0) return;
    k_uncompress_g1<<<div_up(n, 64), 64, 0, st>>>(in48_dev, (uint8_t*)affine_out_dev, flags_dev, n);
    B200_LAUNCH_CHECK();
}

// affine -> Jacobian (blst_p1_from_affine), optionally writing to the bit-reversed position
// (reverse_bit_order of the Lagrange points, kzg/src/eip_4844.rs:1070)
__global__ void k_affine_to_jac(const uint8_t* __restrict__ aff, uint8_t* __restrict__ jac, uint8_t* __restrict__ aff_out, int n, int log_n, int brp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    cc::affine_t a = cc::load_affine(aff + (size_t)i * 96);
    int j = brp ? (int)(__brev((unsigned)i) >> (32 - log_n)) : i;
    cc::jac_t p{a.x, a.y, a.is_inf() ? cc::fp_t::zero() : cc::fp_t::one()};
    cc::store_jac(jac + (size_t)j * 144, p);
    if (aff_out) cc::store_affine(aff_out + (size_t)j * 96, a);
}

// ---------------------------------------------------------------------------------------------------------------
// compute_kzg_proof_rust up to the MSM (kzg/src/eip_4844.rs:437-519) with
// evaluate_polynomial_in_evaluation_form (:954-1003) fused in.  One CTA per blob:
//   d_i = z - w_i;  inv_i = 1/d_i by the batch-inversion product trick of fr_batch_inv (:882-914), per thread over
//   its 16 elements with one inversion per WARP (warp_inverse.cuh: prefix/suffix products by shuffle scans, every lane
//   inverts the same total -- 32 different binary-Euclid inversions in a warp would run the union of their branches);
//   y = (z^4096 - 1)/4096 * sum p_i w_i inv_i;   q_i = (p_i - y)/(w_i - z) = (y - p_i) inv_i.
// The reference's second batch inversion is unnecessary: 1/(w_i - z) = -inv_i.
// z equal to a domain point w_m (:458-462, 484-510): y = p_m, q_m = (1/z) sum_{i != m} (p_i - y) w_i inv_i.
static constexpr int kQThreads = 256;
static constexpr int kQPer = (int)(kFieldElementsPerBlob / kQThreads);  // 16

__device__ __forceinline__ fr_t block_sum_fr(fr_t v, uint8_t* sh /* 8 Fr */) {
    // warp tree through shuffles, then the 8 warp leaders through shared memory; result valid in every thread
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        fr_t o;
#pragma unroll
        for (int k = 0; k < 8; k++) o.v[k] = __shfl_down_sync(0xffffffffu, v.v[k], d);
        v = v + o;
    }
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) store_field(sh + wid * 32, v);
    __syncthreads();
    fr_t total = load_field<fr_t>(sh);
    for (int w = 1; w < kQThreads / 32; w++) total = total + load_field<fr_t>(sh + w * 32);
    return total;
}

__global__ void __launch_bounds__(kQThreads) k_quotient(const uint8_t* __restrict__ poly, const uint8_t* __restrict__ z_all,
                                                        const uint8_t* __restrict__ domain, uint8_t* __restrict__ q_out,
                                                        uint8_t* __restrict__ y_out) {
    extern __shared__ __align__(16) uint8_t sm[];  // 4096 Fr scratch (prefix products, then inverses) + 8 Fr + flags
    uint8_t* pref = sm;
    uint8_t* red = sm + kFieldElementsPerBlob * 32;
    int* found = reinterpret_cast<int*>(red + 8 * 32);
    const size_t blob = blockIdx.x;
    const uint8_t* p = poly + blob * kFieldElementsPerBlob * 32;
    const fr_t z = load_field<fr_t>(z_all + blob * 32);
    if (threadIdx.x == 0) *found = -1;
    __syncthreads();

    // element k of this thread is i = k * 256 + tid (coalesced across the CTA)
    fr_t acc = fr_t::one();
#pragma unroll 1
    for (int k = 0; k < kQPer; k++) {
        int i = k * kQThreads + threadIdx.x;
        fr_t d = z - load_field_ro<fr_t>(domain + (size_t)i * 32);
        if (d.is_zero()) { *found = i; d = fr_t::one(); }
        store_field(pref + (size_t)i * 32, acc);
        acc = acc * d;
    }
    fr_t inv = warp_inverse(acc);   // acc != 0: zero differences were replaced by one above
    fr_t ysum = fr_t::zero();
#pragma unroll 1
    for (int k = kQPer - 1; k >= 0; k--) {
        int i = k * kQThreads + threadIdx.x;
        fr_t w = load_field_ro<fr_t>(domain + (size_t)i * 32);
        fr_t d = z - w;
        bool at_root = d.is_zero();
        if (at_root) d = fr_t::one();
        fr_t inv_i = load_field<fr_t>(pref + (size_t)i * 32) * inv;
        inv = inv * d;
        store_field(pref + (size_t)i * 32, inv_i);
        if (!at_root) ysum = ysum + inv_i * w * load_field_ro<fr_t>(p + (size_t)i * 32);
    }
    fr_t total = block_sum_fr(ysum, red);
    const int m = *found;
    fr_t y;
    if (m >= 0) {
        y = load_field_ro<fr_t>(p + (size_t)m * 32);
    } else {
        // y = total / 4096 * (z^4096 - 1)
        fr_t zn = z;
#pragma unroll 1
        for (int s = 0; s < 12; s++) zn = zn.sqr();
        fr_t n_inv = fr_t::zero();
        n_inv.v[7] = 0x00100000u;  // 4096^-1 in Montgomery form: 2^-12 * 2^256 = 2^244
        y = total * n_inv * (zn - fr_t::one());
    }
    // quotient
    fr_t qm_sum = fr_t::zero();
#pragma unroll 1
    for (int k = 0; k < kQPer; k++) {
        int i = k * kQThreads + threadIdx.x;
        fr_t pi = load_field_ro<fr_t>(p + (size_t)i * 32);
        fr_t inv_i = load_field<fr_t>(pref + (size_t)i * 32);
        fr_t q;
        if (i == m) {
            q = fr_t::zero();
        } else {
            q = (y - pi) * inv_i;
            if (m >= 0) qm_sum = qm_sum + (pi - y) * load_field_ro<fr_t>(domain + (size_t)i * 32) * inv_i;
        }
        store_field(q_out + (blob * kFieldElementsPerBlob + i) * 32, q.from_mont());
    }
    if (m >= 0) {
        fr_t t = block_sum_fr(qm_sum, red);
        if (threadIdx.x == 0) {
            fr_t qm = t * z.inverse();
            store_field(q_out + (blob * kFieldElementsPerBlob + m) * 32, qm.from_mont());
        }
    }
    if (threadIdx.x == 0) store_field(y_out + blob * 32, y);
}

// ---------------------------------------------------------------------------------------------------------------
// compute_cells (kzg/src/das.rs:244-275, 618-629): cells = BRP( NTT_8192( INTT_4096( BRP(blob) ) || 0 ) ).
// in: n x 4096 Montgomery; out: n x 8192, element brev12(i) of the low half, upper half zero
__global__ void __launch_bounds__(256) k_cells_brp_in(const uint8_t* __restrict__ poly, uint8_t* __restrict__ out, size_t total) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t blob = gid / kFieldElementsPerBlob, i = gid % kFieldElementsPerBlob;
    size_t r = __brev((unsigned)i) >> 20;
    fr_t v = load_field_ro<fr_t>(poly + gid * 32);
    store_field(out + (blob * kFieldElementsPerBlob + r) * 32, v);
}
// zero-pad the 4096 monomial coefficients to 8192
__global__ void __launch_bounds__(256) k_cells_pad(const uint8_t* __restrict__ mono, uint8_t* __restrict__ out, size_t total) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t blob = gid / (2 * kFieldElementsPerBlob), i = gid % (2 * kFieldElementsPerBlob);
    fr_t v = i < kFieldElementsPerBlob ? load_field_ro<fr_t>(mono + (blob * kFieldElementsPerBlob + i) * 32) : fr_t::zero();
    store_field(out + gid * 32, v);
}
// bit-reverse (13 bits) and serialise: cell c, element e = position 64 c + e of the bit-reversed extended evaluation
__global__ void __launch_bounds__(256) k_cells_out(const uint8_t* __restrict__ ext, uint8_t* __restrict__ cells, size_t total) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t blob = gid / (2 * kFieldElementsPerBlob), i = gid % (2 * kFieldElementsPerBlob);
    size_t r = __brev((unsigned)i) >> 19;
    fr_t v = load_field_ro<fr_t>(ext + (blob * 2 * kFieldElementsPerBlob + r) * 32).from_mont();
    uint32_t* o = reinterpret_cast<uint32_t*>(cells + gid * 32);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = __byte_perm(v.v[7 - k], 0, 0x0123);
}

void launch_cells_out(const void* ext_dev, uint8_t* cells_dev, size_t total, cudaStream_t st) {
    k_cells_out<<<div_up(total, 256), 256, 0, st>>>((const uint8_t*)ext_dev, cells_dev, total);
    B200_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------------
// FK20 cell proofs (compute_fk20_proofs, kzg/src/das.rs:660-696; setup blst/src/types/kzg_settings.rs:38-101)
constexpr int kCellSize = 64, kFkK = 64, kFkK2 = 128;

// setup: x_ext[offset][t] = g1_monomial[4096 - 64 - 1 - offset - 64 t] for t < 63, infinity otherwise (Jacobian)
__global__ void k_fk_gather_x(const uint8_t* __restrict__ monomial_jac, uint8_t* __restrict__ x_ext) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= kCellSize * kFkK2) return;
    int offset = gid / kFkK2, t = gid % kFkK2;
    uint4* dst = reinterpret_cast<uint4*>(x_ext + (size_t)gid * 144);
    if (t < kFkK - 1) {
        int src = (int)kFieldElementsPerBlob - kCellSize - 1 - offset - t * kCellSize;
        const uint4* q = reinterpret_cast<const uint4*>(monomial_jac + (size_t)src * 144);
#pragma unroll
        for (int k = 0; k < 9; k++) dst[k] = q[k];
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) dst[k] = make_uint4(0, 0, 0, 0);
    }
}
// setup: table[row * 64 + offset] = affine(points[offset][row])
__global__ void __launch_bounds__(64) k_fk_table(const uint8_t* __restrict__ points_jac, uint8_t* __restrict__ table_aff) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= kCellSize * kFkK2) return;
    int row = gid / kCellSize, offset = gid % kCellSize;
    cc::jac_t p = cc::load_jac(points_jac + ((size_t)offset * kFkK2 + row) * 144);
    cc::store_affine(table_aff + (size_t)gid * 96, cc::jac_to_affine(p));
}
// toeplitz_coeffs_stride (kzg/src/das.rs:626-658) for every (blob, offset): 128 Fr each
__global__ void __launch_bounds__(256) k_fk_toeplitz(const uint8_t* __restrict__ mono, uint8_t* __restrict__ out, size_t total,
                                                     size_t mono_stride) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t t = gid % kFkK2, bi = gid / kFkK2, i = bi % kCellSize, b = bi / kCellSize;
    const size_t d = kFieldElementsPerBlob - 1;
    fr_t v = fr_t::zero();
    if (t == 0) {
        v = load_field_ro<fr_t>(mono + (b * mono_stride + d - i) * 32);
    } else if (t > (size_t)kFkK2 - (kFkK - 1)) {  // t = 2r - j, j = 1 .. r-2
        size_t j = kFkK2 - t;
        v = load_field_ro<fr_t>(mono + (b * mono_stride + d - i - j * kCellSize) * 32);
    }
    store_field(out + gid * 32, v);
}
// coeffs[j][i] = fft[(b, i)][j]: transpose into the MSM's scalar layout [vector = b*128 + j][i], canonical form.
// The scalars are multiplied by 1/128 here: the inverse fft_g1 that follows the lincombs is linear, so scaling its
// inputs replaces the [n^-1] scalar multiplication of every output point (blst/src/fft_g1.rs:74-79) by one Fr product.
__global__ void __launch_bounds__(256) k_fk_transpose(const uint8_t* __restrict__ fft, uint8_t* __restrict__ scalars, size_t total,
                                                      const uint8_t* __restrict__ inv_n) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    size_t i = gid % kCellSize, vj = gid / kCellSize, j = vj % kFkK2, b = vj / kFkK2;
    fr_t v = (load_field_ro<fr_t>(fft + ((b * kCellSize + i) * kFkK2 + j) * 32) * load_field_ro<fr_t>(inv_n)).from_mont();
    store_field(scalars + gid * 32, v);
}

static int env_int_local(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}
// The direct-lookup tables (fk20_direct.cu) trade HBM for additions: the widest window <= want whose table still leaves
// B200_DIRECT_RESERVE_GB (default 40) of the device free for everything else in the process; 0 when not even the 8-bit one fits.
int pick_direct_bits(size_t npts, int want) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
    const size_t reserve = (size_t)env_int_local("B200_DIRECT_RESERVE_GB", 40) << 30;
    const int cand[3] = {want, 11, 8};
    for (int c : cand)
        if (c >= 5 && c <= 16 && c <= want && direct_table_bytes(npts, c) + reserve <= free_b) return c;
    return 0;
}
// table of every digit multiple of `points` (n affine points; period > 1: `points` holds period blocks of n): built from the
// fixed-base rows of a throw-away engine with the same window width; nullptr when the allocation fails
void* build_direct_table(const void* points, size_t n, int period, int c, cudaStream_t st) {
    void* p = nullptr;
    if (cudaMalloc(&p, direct_table_bytes(n * period, c)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    MsmConfig rc;
    rc.c = c; rc.W = direct_windows(c); rc.fixed = true; rc.n = n; rc.max_batch = 1; rc.L = 64; rc.randomize = false;
    rc.bases_period = period;
    MsmEngine rows(rc, points, false, st);                    // rows 2^(cj) * P_i; dropped after the build
    launch_direct_build(rows.table(), rows.table_stride(), p, n * period, c, st);
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    return p;
}

KzgSettingsDev::KzgSettingsDev(const uint8_t* g1_monomial, const uint8_t* g1_lagrange, int max_batch, cudaStream_t st)
    : max_batch_(max_batch) {
    const int n = (int)kFieldElementsPerBlob;
    uint8_t* comp = dev_alloc<uint8_t>((size_t)2 * n * 48);
    uint8_t* aff = dev_alloc<uint8_t>((size_t)2 * n * 96);
    uint8_t* aff_brp = dev_alloc<uint8_t>((size_t)n * 96);
    int* flags = dev_alloc<int>(2 * n);
    lagrange_jac_ = dev_alloc<uint8_t>((size_t)n * 144);
    monomial_jac_ = dev_alloc<uint8_t>((size_t)n * 144);
    B200_CUDA_CHECK(cudaMemcpyAsync(comp, g1_lagrange, (size_t)n * 48, cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(comp + (size_t)n * 48, g1_monomial, (size_t)n * 48, cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemsetAsync(flags, 0, 2 * n * sizeof(int), st));
    launch_uncompress_g1(comp, aff, flags, 2 * n, st);
    k_affine_to_jac<<<div_up(n, 128), 128, 0, st>>>(aff, (uint8_t*)lagrange_jac_, aff_brp, n, 12, 1);
    k_affine_to_jac<<<div_up(n, 128), 128, 0, st>>>(aff + (size_t)n * 96, (uint8_t*)monomial_jac_, nullptr, n, 12, 0);
    B200_LAUNCH_CHECK();
    std::vector<int> hflags(2 * n);
    B200_CUDA_CHECK(cudaMemcpyAsync(hflags.data(), flags, 2 * n * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    bool bad = false;
    for (int f : hflags) bad |= f != 0;
    if (!bad) {
        fs_.reset(new FFTSettingsDev(13, st));  // FIELD_ELEMENTS_PER_EXT_BLOB = 8192 (kzg/src/eip_4844.rs:1072-1077)
        MsmConfig cfg;
        // scalar randomisation (msm.cu): the Lagrange points of a real setup are in the prime-order subgroup (the engine checks)
        cfg.randomize = env_int_local("B200_BLOB_RANDOMIZE", 1) != 0;
        // randomised scalars are uniform in Fr whatever the blob bytes are: 13-bit windows (W = 20; the top window holds bits
        // 247..254) with a 2-bit segment fold, as for the proofs -- 64 commitments 2.73 ms against 3.10 ms with c = 12, whose
        // top window would hold 3 bits and put 7/8 of a blob's elements into four buckets (scripts/blob_window_sweep.py).
        // Without randomisation blob elements usually have a zero top byte and c = 12 (bits 240..247 fill window 20) is the
        // better split: 2.85 ms.
        cfg.c = env_int_local("B200_BLOB_C", cfg.randomize ? 13 : 12);
        cfg.fold = env_int_local("B200_BLOB_FOLD", cfg.randomize ? 2 : -1);
        cfg.c0 = env_int_local("B200_BLOB_C0", 0);
        cfg.W = (256 + cfg.c - 1) / cfg.c;
        cfg.fixed = true;
        cfg.n = n;
        cfg.max_batch = max_batch;
        cfg.L = env_int_local("B200_BLOB_L", 64);
        // The proof MSM runs over quotient values, which are uniform in Fr: 13-bit windows (W = 20, the top window still
        // holds 8 random bits) with a 2-bit segment fold in front of the marginal sums cost 9 % less than c = 12
        // (3.03 vs 3.33 ms per 64 proofs).  Blob elements usually have a zero top byte, which leaves window 19 of a
        // 13-bit split with the single bit 247: half of a blob's elements in ONE bucket (64 blobs: 2.98 vs 2.86 ms), so
        // the commitment MSM keeps c = 12, where bits 240..247 fill window 20 (scripts/blob_window_sweep.py).
        MsmConfig cfg_q = cfg;
        cfg_q.c = env_int_local("B200_PROOF_C", 13);
        cfg_q.fold = env_int_local("B200_PROOF_FOLD", 2);
        cfg_q.c0 = 0;
        cfg_q.W = (256 + cfg_q.c - 1) / cfg_q.c;
        for (Lane& ln : lanes_) {
            // one table per window layout, shared by the lanes: what the L2 has to hold does not grow with the lane count
            ln.msm.reset(new MsmEngine(cfg, aff_brp, false, st, &ln == lanes_ ? nullptr : lanes_[0].msm.get()));
            const bool same = cfg_q.c == cfg.c && cfg_q.c0 == cfg.c0 && cfg_q.W == cfg.W && cfg_q.randomize == cfg.randomize;
            ln.msm_q.reset(new MsmEngine(cfg_q, aff_brp, false, st, same ? lanes_[0].msm.get() : &ln == lanes_ ? nullptr : lanes_[0].msm_q.get()));
            ln.scalars = dev_alloc<uint8_t>((size_t)max_batch * n * 32);
            ln.poly = dev_alloc<uint8_t>((size_t)max_batch * n * 32);
            ln.z = dev_alloc<uint8_t>((size_t)max_batch * 32);
            ln.y = dev_alloc<uint8_t>((size_t)max_batch * 32);
            ln.out_jac = dev_alloc<uint8_t>((size_t)max_batch * 144);
        }
        // direct lookup table of the Lagrange points (fk20_direct.cu) for batches of up to direct_max_ blobs: 13-bit windows,
        // 30 GiB, 20 additions per element instead of 32 (B200_BLOB_DIRECT_BITS; narrower when HBM is short).
        // B200_BLOB_DIRECT=0 or a failed allocation leaves every batch to the bucket engine.
        direct_max_ = std::min(max_batch, env_int_local("B200_BLOB_DIRECT", 64));
        if (direct_max_ > 0) {
            direct_c_ = pick_direct_bits(n, env_int_local("B200_BLOB_DIRECT_BITS", 13));
            if (direct_c_) lag_direct_ = build_direct_table(aff_brp, n, 1, direct_c_, st);
            if (lag_direct_) {
                for (Lane& ln : lanes_) {
                    ln.direct_part = dev_alloc<uint8_t>((size_t)direct_max_ * 128 * 192);
                    ln.direct_cnt = dev_alloc<unsigned>(direct_max_);
                    B200_CUDA_CHECK(cudaMemsetAsync(ln.direct_cnt, 0, direct_max_ * sizeof(unsigned), st));
                }
                B200_CUDA_CHECK(cudaStreamSynchronize(st));
            } else {
                direct_max_ = direct_c_ = 0;
            }
        }
        // the 4096 domain = first half of the bit-reversed 8192 roots (kzg/src/eip_4844.rs:463, 976)
        domain_ = dev_alloc<uint8_t>((size_t)n * 32);
        B200_CUDA_CHECK(cudaMemcpyAsync(domain_, fs_->brp_roots_dev(), (size_t)n * 32, cudaMemcpyDeviceToDevice, st));
        B200_CUDA_CHECK(cudaFuncSetAttribute(k_quotient, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(kFieldElementsPerBlob * 32 + 8 * 32 + 64)));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    cudaFree(comp); cudaFree(aff); cudaFree(aff_brp); cudaFree(flags);
    if (bad) {
        cudaFree(lagrange_jac_); cudaFree(monomial_jac_);
        lagrange_jac_ = monomial_jac_ = nullptr;
        throw CudaError(1, "Failed to uncompress");  // FsG1::from_bytes error text (blst/src/types/g1.rs:82)
    }
}

KzgSettingsDev::~KzgSettingsDev() {
    cudaFree(lagrange_jac_); cudaFree(monomial_jac_); cudaFree(domain_);
    for (Lane& ln : lanes_) { cudaFree(ln.scalars); cudaFree(ln.poly); cudaFree(ln.z); cudaFree(ln.y); cudaFree(ln.out_jac); cudaFree(ln.direct_part); cudaFree(ln.direct_cnt); }
    cudaFree(lag_direct_);
    cudaFree(cells_a_); cudaFree(cells_b_);
    cudaFree(fk_a_); cudaFree(fk_b_); cudaFree(fk_pts_); cudaFree(fk_direct_); cudaFree(fk_team_part_); cudaFree(fk_team_cnt_);
    cudaFree(g2_affine_); cudaFree(g2_jac_); cudaFree(g2_lines_); cudaFree(vf_buf_); cudaFree(das_buf_);
}

// Small batches: direct table lookups, ONE launch (fk20_direct.cu: sums, fold and compression fused through a last-CTA
// counter); batches that fill the machine: the bucket engine, then the compression kernel.
int KzgSettingsDev::lagrange_msm(int lane, MsmEngine& eng, int n, uint8_t* out48, cudaStream_t st) {
    Lane& ln = lanes_[lane % kLanes];
    if (lag_direct_ && n <= direct_max_) {
        launch_direct_msm_compressed(ln.scalars, lag_direct_, ln.direct_part, ln.direct_cnt, out48, n, (int)kFieldElementsPerBlob, direct_c_, st);
        return 1;
    }
    eng.run(ln.scalars, kFieldElementsPerBlob, n, false, ln.out_jac, st);
    launch_points_to_compressed(ln.out_jac, out48, n, st);
    return eng.launches_per_run() + 1;
}

void KzgSettingsDev::blob_to_commitments(const uint8_t* blobs, int n, uint8_t* out48, int* status, cudaStream_t st, int lane) {
    if (n < 1 || n > max_batch_) throw CudaError(-1, "blob batch exceeds the settings' capacity");
    Lane& ln = lanes_[lane % kLanes];
    size_t total = (size_t)n * kFieldElementsPerBlob;
    k_blob_to_fr<<<div_up(total, 256), 256, 0, st>>>(blobs, total, (uint8_t*)ln.scalars, nullptr, status);
    B200_LAUNCH_CHECK();
    launches_ = 1 + lagrange_msm(lane, *ln.msm, n, out48, st);
}

void KzgSettingsDev::compute_proofs(const uint8_t* blobs, const uint8_t* z_bytes, int z_reduce, int n, uint8_t* proofs48,
                                    uint8_t* y32, int* status, cudaStream_t st, int lane) {
    if (n < 1 || n > max_batch_) throw CudaError(-1, "blob batch exceeds the settings' capacity");
    Lane& ln = lanes_[lane % kLanes];
    size_t total = (size_t)n * kFieldElementsPerBlob;
    k_blob_to_fr<<<div_up(total, 256), 256, 0, st>>>(blobs, total, nullptr, (uint8_t*)ln.poly, status);
    k_z_to_fr<<<div_up(n, 64), 64, 0, st>>>(z_bytes, n, z_reduce, (uint8_t*)ln.z, status);
    k_quotient<<<n, kQThreads, kFieldElementsPerBlob * 32 + 8 * 32 + 64, st>>>((const uint8_t*)ln.poly, (const uint8_t*)ln.z,
                                                                              (const uint8_t*)domain_, (uint8_t*)ln.scalars,
                                                                              (uint8_t*)ln.y);
    B200_LAUNCH_CHECK();
    const int msm_launches = lagrange_msm(lane, *ln.msm_q, n, proofs48, st);
    int extra = 0;
    if (y32) {
        k_fr_to_bytes<<<div_up(n, 64), 64, 0, st>>>((const uint8_t*)ln.y, n, y32);
        B200_LAUNCH_CHECK();
        extra = 1;
    }
    launches_ = 3 + extra + msm_launches;
}

void KzgSettingsDev::compute_cells(const uint8_t* blobs, int n, uint8_t* cells_out, int* status, cudaStream_t st) {
    if (n < 1 || n > max_batch_) throw CudaError(-1, "blob batch exceeds the settings' capacity");
    size_t total = (size_t)n * kFieldElementsPerBlob;
    if (!cells_a_) {
        cells_a_ = dev_alloc<uint8_t>((size_t)max_batch_ * 2 * kFieldElementsPerBlob * 32);
        cells_b_ = dev_alloc<uint8_t>((size_t)max_batch_ * 2 * kFieldElementsPerBlob * 32);
    }
    k_blob_to_fr<<<div_up(total, 256), 256, 0, st>>>(blobs, total, nullptr, (uint8_t*)lanes_[0].poly, status);
    k_cells_brp_in<<<div_up(total, 256), 256, 0, st>>>((const uint8_t*)lanes_[0].poly, (uint8_t*)cells_a_, total);
    B200_LAUNCH_CHECK();
    fs_->fft_fr(cells_a_, cells_b_, kFieldElementsPerBlob, true, n, st);             // poly_lagrange_to_monomial
    k_cells_pad<<<div_up(2 * total, 256), 256, 0, st>>>((const uint8_t*)cells_b_, (uint8_t*)cells_a_, 2 * total);
    B200_LAUNCH_CHECK();
    fs_->fft_fr(cells_a_, cells_b_, 2 * kFieldElementsPerBlob, false, n, st);
    k_cells_out<<<div_up(2 * total, 256), 256, 0, st>>>((const uint8_t*)cells_b_, cells_out, 2 * total);
    B200_LAUNCH_CHECK();
    launches_ = 7;
}

// x_ext_fft_columns of FsKZGSettings::new (blst/src/types/kzg_settings.rs:84-101) for the host KZGSettings struct:
// out[row][offset] (128 x 64 blst_p1, row-major) = toeplitz_part_1 of offset `offset`, entry `row`.  Device buffer out.
__global__ void __launch_bounds__(128) k_fk_columns_out(const uint8_t* __restrict__ points_jac, uint8_t* __restrict__ out) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= kCellSize * kFkK2) return;
    int row = gid / kCellSize, offset = gid % kCellSize;
    const uint4* src = reinterpret_cast<const uint4*>(points_jac + ((size_t)offset * kFkK2 + row) * 144);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)gid * 144);
#pragma unroll
    for (int k = 0; k < 9; k++) dst[k] = src[k];
}
void KzgSettingsDev::x_ext_fft_columns(void* out_dev, cudaStream_t st) {
    const int npts = kCellSize * kFkK2;  // 8192
    uint8_t* x_ext = dev_alloc<uint8_t>((size_t)npts * 144);
    uint8_t* points = dev_alloc<uint8_t>((size_t)npts * 144);
    try {
        k_fk_gather_x<<<div_up(npts, 128), 128, 0, st>>>((const uint8_t*)monomial_jac_, x_ext);
        B200_LAUNCH_CHECK();
        fs_->fft_g1(x_ext, points, kFkK2, false, kCellSize, st);
        k_fk_columns_out<<<div_up(npts, 128), 128, 0, st>>>(points, (uint8_t*)out_dev);
        B200_LAUNCH_CHECK();
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (...) {
        cudaFree(x_ext); cudaFree(points);
        throw;
    }
    cudaFree(x_ext); cudaFree(points);
}

void KzgSettingsDev::ensure_fk20(cudaStream_t st) {
    if (fk_ready_) return;
    fk_batch_ = std::max(1, std::min(max_batch_, env_int_local("B200_FK20_BATCH", 64)));
    const int npts = kCellSize * kFkK2;  // 8192
    uint8_t* x_ext = dev_alloc<uint8_t>((size_t)npts * 144);
    uint8_t* points = dev_alloc<uint8_t>((size_t)npts * 144);
    uint8_t* table = dev_alloc<uint8_t>((size_t)npts * 96);
    k_fk_gather_x<<<div_up(npts, 128), 128, 0, st>>>((const uint8_t*)monomial_jac_, x_ext);
    B200_LAUNCH_CHECK();
    // toeplitz_part_1: 64 forward fft_g1 of size 128 over the zero-extended vectors (kzg_settings.rs:38-61)
    fs_->fft_g1(x_ext, points, kFkK2, false, kCellSize, st);
    k_fk_table<<<div_up(npts, 64), 64, 0, st>>>(points, table);
    B200_LAUNCH_CHECK();
    // direct lookup table of every digit multiple (fk20_direct.cu): 13-bit windows, 60 GiB of HBM for a lincomb stage of
    // 64 x 20 additions without buckets (B200_FK20_DIRECT_BITS; narrower when HBM is short).  B200_FK20_DIRECT=0 or a failed
    // allocation: the bucket engine (8-bit windows, one bucket set per lincomb) serves the lincombs.
    if (env_int_local("B200_FK20_DIRECT", 1)) {
        fk_direct_c_ = pick_direct_bits(npts, env_int_local("B200_FK20_DIRECT_BITS", 13));
        if (fk_direct_c_) fk_direct_ = build_direct_table(table, kCellSize, kFkK2, fk_direct_c_, st);
    }
    if (!fk_direct_) {
        MsmConfig cfg;
        cfg.c = env_int_local("B200_FK20_C", 8);
        cfg.W = (256 + cfg.c - 1) / cfg.c;
        cfg.fixed = true;
        cfg.n = kCellSize;
        cfg.max_batch = fk_batch_ * kFkK2;
        cfg.L = 64;
        cfg.bases_period = kFkK2;
        fk_msm_.reset(new MsmEngine(cfg, table, false, st));
    }
    // small batches: the lincombs by teams of CTAs per vector (k_direct_msm with per-vector point blocks) instead of one warp
    // per lincomb -- one blob's 128 lincombs are 128 warps of 40 additions each, a 0.5 ms chain on an empty machine
    fk_team_max_ = fk_direct_ && (kCellSize * direct_windows(fk_direct_c_)) % 128 == 0 ? std::min(fk_batch_, env_int_local("B200_FK20_TEAM", 2)) : 0;   // measured: 1 blob 3.91 -> 3.59 ms, 2: 3.93 -> 3.72, 4: equal, 8 and up the warp form wins
    if (fk_team_max_ > 0) {
        fk_team_part_ = dev_alloc<uint8_t>((size_t)fk_team_max_ * kFkK2 * 128 * 192);
        fk_team_cnt_ = dev_alloc<unsigned>((size_t)fk_team_max_ * kFkK2);
        B200_CUDA_CHECK(cudaMemsetAsync(fk_team_cnt_, 0, (size_t)fk_team_max_ * kFkK2 * sizeof(unsigned), st));
    }
    fs_->prepare_g1((size_t)fk_batch_ * kFkK2);   // no buffer growth or lazy kernel load inside a later multi-blob pass
    fk_ready_ = true;
    fk_a_ = dev_alloc<uint8_t>((size_t)fk_batch_ * kCellSize * kFkK2 * 32);
    fk_b_ = dev_alloc<uint8_t>((size_t)fk_batch_ * kCellSize * kFkK2 * 32);
    fk_pts_ = dev_alloc<uint8_t>((size_t)fk_batch_ * kFkK2 * 144);
    if (!cells_a_) {
        cells_a_ = dev_alloc<uint8_t>((size_t)max_batch_ * 2 * kFieldElementsPerBlob * 32);
        cells_b_ = dev_alloc<uint8_t>((size_t)max_batch_ * 2 * kFieldElementsPerBlob * 32);
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(x_ext); cudaFree(points); cudaFree(table);
}

void KzgSettingsDev::compute_cell_proofs(const uint8_t* blobs, int n, uint8_t* proofs48, int* status, cudaStream_t st) {
    ensure_fk20(st);
    if (n < 1 || n > fk_batch_) throw CudaError(-1, "blob batch exceeds the FK20 capacity");
    size_t total = (size_t)n * kFieldElementsPerBlob;
    // polynomial in monomial form (poly_lagrange_to_monomial, kzg/src/das.rs:618-629)
    k_blob_to_fr<<<div_up(total, 256), 256, 0, st>>>(blobs, total, nullptr, (uint8_t*)lanes_[0].poly, status);
    k_cells_brp_in<<<div_up(total, 256), 256, 0, st>>>((const uint8_t*)lanes_[0].poly, (uint8_t*)cells_a_, total);
    B200_LAUNCH_CHECK();
    fs_->fft_fr(cells_a_, cells_b_, kFieldElementsPerBlob, true, n, st);
    fk20_from_mono(cells_b_, kFieldElementsPerBlob, n, proofs48, st);
    launches_ = 8 + (fk_direct_ ? 1 : fk_msm_->launches_per_run()) + 2 * 9;
}
// compute_cells_and_kzg_proofs with both outputs (kzg/src/das.rs:244-292): one pass from blob bytes to the monomial form, shared
// by the cells (NTT-8192) and the FK20 proofs.  cells_done is recorded once cells_out is complete, so the caller can copy the
// cells out on another stream while the proofs -- 80 % of the call -- are still running.
void KzgSettingsDev::compute_cells_and_proofs(const uint8_t* blobs, int n, uint8_t* cells_out, uint8_t* proofs48, int* status,
                                              cudaStream_t st, cudaEvent_t cells_done) {
    ensure_fk20(st);
    if (n < 1 || n > fk_batch_ || n > max_batch_) throw CudaError(-1, "blob batch exceeds the FK20 capacity");
    compute_cells(blobs, n, cells_out, status, st);
    if (cells_done) B200_CUDA_CHECK(cudaEventRecord(cells_done, st));
    // cells_a_ still holds the zero-padded monomial form (8192 Fr per blob): fft_fr is out of place
    fk20_from_mono(cells_a_, 2 * kFieldElementsPerBlob, n, proofs48, st);
    launches_ = 7 + 5 + (fk_direct_ ? 1 : fk_msm_->launches_per_run()) + 2 * 9;
}
// compute_fk20_proofs (kzg/src/das.rs:660-696) from polynomials in monomial form: blob b's coefficients 0..4095 start at
// mono + b * stride Fr (Montgomery)
void KzgSettingsDev::fk20_from_mono(const void* mono, size_t stride, int n, uint8_t* proofs48, cudaStream_t st) {
    ensure_fk20(st);
    if (n < 1 || n > fk_batch_) throw CudaError(-1, "blob batch exceeds the FK20 capacity");
    // Toeplitz coefficient vectors and their 128-point transforms
    size_t tt = (size_t)n * kCellSize * kFkK2;
    k_fk_toeplitz<<<div_up(tt, 256), 256, 0, st>>>((const uint8_t*)mono, (uint8_t*)fk_a_, tt, stride);
    B200_LAUNCH_CHECK();
    fs_->fft_fr(fk_a_, fk_b_, kFkK2, false, n * kCellSize, st);
    k_fk_transpose<<<div_up(tt, 256), 256, 0, st>>>((const uint8_t*)fk_b_, (uint8_t*)fk_a_, tt, (const uint8_t*)fs_->inv_pow2_dev(7));
    B200_LAUNCH_CHECK();
    // g1_lincomb_batch: 128 lincombs of 64 fixed points per blob (kzg/src/das.rs:676-680)
    if (fk_direct_ && n <= fk_team_max_)
        launch_direct_msm(fk_a_, fk_direct_, fk_team_part_, fk_team_cnt_, nullptr, (uint8_t*)fk_pts_, n * kFkK2, kCellSize, fk_direct_c_, st, kFkK2);
    else if (fk_direct_) launch_direct_lincomb(fk_a_, fk_direct_, fk_pts_, n * kFkK2, kFkK2, kCellSize, fk_direct_c_, st);
    else fk_msm_->run(fk_a_, kCellSize, n * kFkK2, false, fk_pts_, st);
    // h = inverse fft_g1, upper half := identity, forward fft_g1 (:682-695)
    fs_->fft_g1(fk_pts_, fk_pts_, kFkK2, true, n, st, /*apply_scale=*/false);
    B200_CUDA_CHECK(cudaMemset2DAsync((uint8_t*)fk_pts_ + (size_t)kFkK * 144, (size_t)kFkK2 * 144, 0, (size_t)kFkK * 144, n, st));
    fs_->fft_g1(fk_pts_, fk_pts_, kFkK2, false, n, st);
    launch_points_to_compressed(fk_pts_, proofs48, n * kFkK2, st, 7);  // reverse_bit_order(proofs) (:287)
}

// G1::from_bytes + "!is_inf && !is_valid -> Err" of the verify_* functions (kzg/src/eip_4844.rs:601-606, 655-660,
// 720-736) and of compute_blob_kzg_proof (:556-558): affine out (may be nullptr), status[i % status_mod] = 1 when point i
// is malformed, off the curve or outside the subgroup.
// One lane QUAD per point: decoding (a 381-bit square-root chain) runs redundantly on the four lanes, then the subgroup
// test -- (beta x, y) == -[z^2] P, two 64-bit ladders (eprint 2021/1130 sec. 6, zkcrypto/bls12_381/src/g1.rs:401-410) --
// runs on the quad arithmetic of g1_quad.cuh: 3 / 4 multiplication levels per doubling / addition instead of 9 / 14
// multiplications, and every quad of the warp follows the same instruction stream (the scalar is a constant).
__device__ __forceinline__ bool uncompress_point_u(const uint8_t* in, affine_t& out) {
    uint32_t b0 = in[0];
    uint32_t cflag = b0 >> 7, iflag = (b0 >> 6) & 1, sflag = (b0 >> 5) & 1;
    fp_t x;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const uint8_t* p = in + 4 * (11 - k);
        x.v[k] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
    }
    x.v[11] &= 0x1fffffffu;
    out.x = fp_t::zero();
    out.y = fp_t::zero();
    if (!cflag) return false;
    if (iflag) return !sflag && x.is_zero();
    bool lt = false;
#pragma unroll
    for (int i = 11; i >= 0; i--) {
        uint32_t m = FpParams::mod(i);
        if (x.v[i] != m) { lt = x.v[i] < m; break; }
    }
    if (!lt) return false;
    fp_t xm = x.to_mont();
    fp_t four = fp_t::one().dbl().dbl();
    fp_t y2 = xm.sqr() * xm + four;
    const uint32_t E[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                            0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};  // (p+1)/4
    fp_t y = y2.pow_words(E);
    if (y.sqr() != y2) return false;
    if (fp_is_lex_largest(y) != (bool)sflag) y = y.neg();
    out.x = xm;
    out.y = y;
    return true;
}
// mode 0: decode + subgroup test; 1: decode only (status for malformed / off-curve input); 2: subgroup test only, on the
// affine points a mode-1 launch left in `out` (so that the test can run beside the caller's critical path)
__global__ void __launch_bounds__(32) k_decode_g1_checked(const uint8_t* __restrict__ in, int n, uint8_t* __restrict__ out,
                                                          int* __restrict__ status, int status_mod, int mode) {
    const int q = blockIdx.x * 8 + (threadIdx.x >> 2), role = threadIdx.x & 3;
    const bool live = q < n;
    const int i = live ? q : 0;
    affine_t a;
    bool ok = true;
    if (mode == 2) a = load_affine(out + (size_t)i * 96);
    else ok = uncompress_point_u(in + (size_t)i * 48, a);
    const bool inf = a.is_inf();
    bool in_g1 = true;
    if (mode != 1) {
        fp_t comp = role == 0 ? a.x : role == 1 ? a.y : fp_t::one();
        if (inf) comp = fp_t::zero();
        in_g1 = quad_in_subgroup(a, comp);                  // every lane of the warp takes part
    }
    if (role == 0 && live) {
        if (ok && !inf) ok = in_g1;
        if (!ok) status[i % status_mod] = 1;
        if (out && mode != 2) store_affine(out + (size_t)i * 96, a);
    }
}
void launch_decode_g1_checked(const uint8_t* in48_dev, void* affine_out_dev, int* status_dev, int n, cudaStream_t st, int status_mod) {
    if (n < 1) return;
    k_decode_g1_checked<<<div_up(n, 8), 32, 0, st>>>(in48_dev, n, (uint8_t*)affine_out_dev, status_dev, status_mod > 0 ? status_mod : n, 0);
    B200_LAUNCH_CHECK();
}
// the two halves of launch_decode_g1_checked as separate launches: decode (affine_out_dev required), then the subgroup test
void launch_decode_g1_unchecked(const uint8_t* in48_dev, void* affine_out_dev, int* status_dev, int n, cudaStream_t st, int status_mod) {
    if (n < 1) return;
    k_decode_g1_checked<<<div_up(n, 8), 32, 0, st>>>(in48_dev, n, (uint8_t*)affine_out_dev, status_dev, status_mod > 0 ? status_mod : n, 1);
    B200_LAUNCH_CHECK();
}
void launch_subgroup_g1(const void* affine_dev, int* status_dev, int n, cudaStream_t st, int status_mod) {
    if (n < 1) return;
    k_decode_g1_checked<<<div_up(n, 8), 32, 0, st>>>(nullptr, n, (uint8_t*)const_cast<void*>(affine_dev), status_dev, status_mod > 0 ? status_mod : n, 2);
    B200_LAUNCH_CHECK();
}
void launch_affine_to_jac(const void* affine_dev, void* jac_dev, int n, cudaStream_t st) {
    if (n < 1) return;
    k_affine_to_jac<<<div_up(n, 128), 128, 0, st>>>((const uint8_t*)affine_dev, (uint8_t*)jac_dev, nullptr, n, 0, 0);
    B200_LAUNCH_CHECK();
}
void launch_fr_from_bytes(const uint8_t* bytes32_dev, int n, int reduce, void* fr_mont_dev, int* status_dev, cudaStream_t st) {
    if (n < 1) return;
    k_z_to_fr<<<div_up(n, 64), 64, 0, st>>>(bytes32_dev, n, reduce, (uint8_t*)fr_mont_dev, status_dev);
    B200_LAUNCH_CHECK();
}
void launch_fr_to_bytes(const void* fr_mont_dev, int n, uint8_t* bytes32_dev, cudaStream_t st) {
    if (n < 1) return;
    k_fr_to_bytes<<<div_up(n, 64), 64, 0, st>>>((const uint8_t*)fr_mont_dev, n, bytes32_dev);
    B200_LAUNCH_CHECK();
}

// y_i = p_i(z_i) (evaluate_polynomial_in_evaluation_form, kzg/src/eip_4844.rs:954-1003) for the blob verifiers
// (:662-666, 700-718): the same fused kernel as the proof path; the quotient it also produces is discarded.
void KzgSettingsDev::evaluate_blobs(const uint8_t* blobs, const uint8_t* z_bytes, int z_reduce, int n, uint8_t* z32, uint8_t* y32,
                                    int* status, cudaStream_t st, int lane) {
    if (n < 1 || n > max_batch_) throw CudaError(-1, "blob batch exceeds the settings' capacity");
    Lane& ln = lanes_[lane % kLanes];
    size_t total = (size_t)n * kFieldElementsPerBlob;
    k_blob_to_fr<<<div_up(total, 256), 256, 0, st>>>(blobs, total, nullptr, (uint8_t*)ln.poly, status);
    k_z_to_fr<<<div_up(n, 64), 64, 0, st>>>(z_bytes, n, z_reduce, (uint8_t*)ln.z, status);
    k_quotient<<<n, kQThreads, kFieldElementsPerBlob * 32 + 8 * 32 + 64, st>>>((const uint8_t*)ln.poly, (const uint8_t*)ln.z,
                                                                              (const uint8_t*)domain_, (uint8_t*)ln.scalars,
                                                                              (uint8_t*)ln.y);
    k_fr_to_bytes<<<div_up(n, 64), 64, 0, st>>>((const uint8_t*)ln.z, n, z32);
    k_fr_to_bytes<<<div_up(n, 64), 64, 0, st>>>((const uint8_t*)ln.y, n, y32);
    B200_LAUNCH_CHECK();
    launches_ = 5;
}

void KzgSettingsDev::validate_commitments(const uint8_t* commitments48, int n, int* status, cudaStream_t st) {
    if (n < 1) return;
    launch_decode_g1_checked(commitments48, nullptr, status, n, st);
}

}  // namespace b200
