// das7594.cu -- EIP-7594 recovery and cell-proof batch verification on the device
// (DAS::recover_cells_and_kzg_proofs, kzg/src/das.rs:101-207 with recover_cells :549-616;
//  DAS::verify_cell_kzg_proof_batch, kzg/src/das.rs:294-388 with its helpers :697-900).
//
// Recovery is the reference's sequence of exact field operations (so it agrees even on inconsistent inputs):
//   E0 = evaluations in natural order with 0 at the missing cells, Z(x) = prod (x^64 - w^(64 m)) over missing cells m,
//   (E0 . Z) -> coefficients -> coset 7 -> divide by Z on the coset -> back: five 8192-point NTTs + one for Z on the
//   coset, with the pointwise work fused into small kernels around them.  The reconstructed coefficient vector IS the
//   monomial form FK20 needs, so the reference's extra inverse transform (poly_lagrange_to_monomial, :186-188) is free.
// Cell verification reduces to two short linear combinations and one pairing against the tabulated [s^64]G2 lines
// (verify.cu): n proofs, m unique commitments and the 64 monomial points of the aggregated interpolation polynomial.
#include "eip4844.cuh"
#include "g1.cuh"
#include "pairing.cuh"
#include "util.cuh"
#include "warp_inverse.cuh"
#include "wire.cuh"

#include <vector>

namespace b200 {

static constexpr int kExt = 8192, kCells = 128, kCellFr = 64;

__device__ __forceinline__ fr_t fr_pow_small(fr_t base, uint32_t e) {
    fr_t acc = fr_t::one();
    while (e) {
        if (e & 1) acc = acc * base;
        base = base.sqr();
        e >>= 1;
    }
    return acc;
}
__device__ __forceinline__ fr_t fr_from_u32(uint32_t v) {
    fr_t t = fr_t::zero();
    t.v[0] = v;
    return t.to_mont();
}

// ---- recovery ---------------------------------------------------------------------------------------------------------
// provided cell i, element e -> natural-order slot brev13(64 * idx_i + e); out must be zeroed beforehand
__global__ void __launch_bounds__(256) k_rc_scatter(const uint8_t* __restrict__ cells, const uint32_t* __restrict__ idx, int n,
                                                    uint8_t* __restrict__ out, int* __restrict__ status) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * kCellFr) return;
    int i = gid / kCellFr, e = gid % kCellFr;
    fr_t v;
    load_be32(cells + (size_t)gid * 32, v.v);
    if (!lt_r(v.v)) status[0] = 1;
    uint32_t pos = idx[i] * kCellFr + e;
    uint32_t r = __brev(pos) >> 19;
    store_field(out + (size_t)r * 32, v.to_mont());
}
// vanishing polynomial of the missing cells (compute_vanishing_polynomial_from_roots, das.rs:488-514, spread with stride
// 64, :516-547): thread j owns coefficient j of prod (x - root_i); out (8192 Fr) must be zeroed beforehand.
// missing[i] = brev7(cell index); root_i = roots_of_unity[missing[i] * 64]
__global__ void __launch_bounds__(128) k_rc_vanishing(const uint32_t* __restrict__ missing, int nm, const uint8_t* __restrict__ roots,
                                                      uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t sh[2][(kCells + 1) * 32];
    const int j = threadIdx.x;
    fr_t cur = j == 0 ? fr_t::one() : fr_t::zero();   // polynomial "1"
    int buf = 0;
    for (int i = 0; i < nm; i++) {
        store_field(sh[buf] + j * 32, cur);
        __syncthreads();
        fr_t neg = load_field_ro<fr_t>(roots + (size_t)missing[i] * 64 * 32).neg();
        fr_t lower = j > 0 ? load_field<fr_t>(sh[buf] + (j - 1) * 32) : fr_t::zero();
        cur = cur * neg + lower;                       // (x - root) * poly
        buf ^= 1;
    }
    if (j <= nm) store_field(out + (size_t)j * kCellFr * 32, cur);
}
// out[i] = a[i] * b[i]
__global__ void __launch_bounds__(256) k_fr_mul(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_field(out + (size_t)i * 32, load_field<fr_t>(a + (size_t)i * 32) * load_field<fr_t>(b + (size_t)i * 32));
}
// shift_poly (das.rs:454-460): out[i] = a[i] * f^i with f = 7 or 1/7
__global__ void __launch_bounds__(256) k_fr_shift(const uint8_t* __restrict__ a, uint8_t* __restrict__ out, int n, int inverse) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t f = fr_from_u32(7);
    if (inverse) f = f.inverse();
    store_field(out + (size_t)i * 32, load_field<fr_t>(a + (size_t)i * 32) * fr_pow_small(f, (uint32_t)i));
}
// out[i] = a[i] / b[i]   (batch_inverse + product, das.rs:596-602)
__global__ void __launch_bounds__(256) k_fr_div(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;   // n is a multiple of the warp size: whole warps are live
    if (i >= n) return;
    fr_t d = load_field<fr_t>(b + (size_t)i * 32);
    const bool z = d.is_zero();                      // cannot happen for Z on the coset; 0 -> 0 as the scalar inverse does
    fr_t inv = warp_inverse(z ? fr_t::one() : d);
    if (z) inv = fr_t::zero();
    store_field(out + (size_t)i * 32, load_field<fr_t>(a + (size_t)i * 32) * inv);
}
// natural-order evaluations -> bit-reversed order (reverse_bit_order of 8192 elements)
__global__ void __launch_bounds__(256) k_fr_brp13(const uint8_t* __restrict__ a, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kExt) return;
    store_field(out + (size_t)(__brev((unsigned)i) >> 19) * 32, load_field<fr_t>(a + (size_t)i * 32));
}

uint8_t* KzgSettingsDev::ensure_das_ws(size_t bytes) {
    if (bytes > das_bytes_) {
        cudaFree(das_buf_);
        das_buf_ = nullptr;
        das_bytes_ = 0;
        das_buf_ = dev_alloc<uint8_t>(bytes);
        das_bytes_ = bytes;
    }
    return (uint8_t*)das_buf_;
}

void KzgSettingsDev::recover_cells(const uint8_t* cells, const uint64_t* cell_idx, int n, uint8_t* cells_out, uint8_t* proofs48,
                                   int* status, cudaStream_t st) {
    if (n < kCells / 2 || n > kCells) throw CudaError(1, "cell count out of range");
    const size_t V = (size_t)kExt * 32;
    uint8_t* w = ensure_das_ws(6 * V + 4096);
    uint8_t *e0 = w, *van = w + V, *t1 = w + 2 * V, *t2 = w + 3 * V, *t3 = w + 4 * V, *t4 = w + 5 * V;
    uint32_t* d_idx = reinterpret_cast<uint32_t*>(w + 6 * V);        // n cell indices, then the missing list
    std::vector<uint32_t> h_idx(2 * kCells, 0);
    bool present[kCells] = {false};
    for (int i = 0; i < n; i++) { h_idx[i] = (uint32_t)cell_idx[i]; present[cell_idx[i]] = true; }
    int nm = 0;
    for (int c = 0; c < kCells; c++)
        if (!present[c]) {   // reverse_bits_limited(128, c)
            uint32_t rb = 0;
            for (int b = 0; b < 7; b++) rb |= ((uint32_t)(c >> b) & 1u) << (6 - b);
            h_idx[kCells + nm++] = rb;
        }
    B200_CUDA_CHECK(cudaMemcpyAsync(d_idx, h_idx.data(), 2 * kCells * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemsetAsync(e0, 0, 2 * V, st));               // e0 and van
    k_rc_scatter<<<div_up((size_t)n * kCellFr, 256), 256, 0, st>>>(cells, d_idx, n, e0, status);
    B200_LAUNCH_CHECK();
    const uint8_t* mono = nullptr;                                      // 8192 monomial coefficients (Montgomery)
    const uint8_t* evals = nullptr;                                     // 8192 evaluations, natural order
    if (nm == 0) {
        evals = e0;
        if (proofs48) {
            fs_->fft_fr(e0, t1, kExt, true, 1, st);                     // poly_lagrange_to_monomial (das.rs:186-188)
            mono = t1;
        }
    } else {
        k_rc_vanishing<<<1, 128, 0, st>>>(d_idx + kCells, nm, (const uint8_t*)fs_->roots_dev(), van);
        B200_LAUNCH_CHECK();
        fs_->fft_fr(van, t1, kExt, false, 1, st);                       // Z on the domain
        k_fr_mul<<<kExt / 256, 256, 0, st>>>(e0, t1, t2, kExt);         // (E . Z); E is 0 where cells are missing
        fs_->fft_fr(t2, t1, kExt, true, 1, st);                         // coefficients of E Z
        k_fr_shift<<<kExt / 256, 256, 0, st>>>(t1, t2, kExt, 0);
        fs_->fft_fr(t2, t1, kExt, false, 1, st);                        // E Z over the coset (coset_fft)
        k_fr_shift<<<kExt / 256, 256, 0, st>>>(van, t2, kExt, 0);
        fs_->fft_fr(t2, t3, kExt, false, 1, st);                        // Z over the coset
        k_fr_div<<<kExt / 256, 256, 0, st>>>(t1, t3, t2, kExt);
        fs_->fft_fr(t2, t1, kExt, true, 1, st);                         // coset_ifft ...
        k_fr_shift<<<kExt / 256, 256, 0, st>>>(t1, t4, kExt, 1);        // ... reconstructed coefficients
        fs_->fft_fr(t4, t3, kExt, false, 1, st);                        // all 8192 evaluations
        B200_LAUNCH_CHECK();
        mono = t4;
        evals = t3;
    }
    launch_cells_out(evals, cells_out, kExt, st);
    if (proofs48) fk20_from_mono(mono, kExt, 1, proofs48, st);
}

// ---- cell verification -----------------------------------------------------------------------------------------------
// r^i for i < n (compute_powers, kzg/src/eip_4844.rs:316-326), one thread per power; cells -> Montgomery with the canonical
// check of Fr::from_bytes
__global__ void __launch_bounds__(256) k_vc_powers(const uint8_t* __restrict__ r_mont, int n, uint8_t* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_field(out + (size_t)i * 32, fr_pow_small(load_field<fr_t>(r_mont), (uint32_t)i));
}
__global__ void __launch_bounds__(256) k_vc_cells(const uint8_t* __restrict__ cells, int n, uint8_t* __restrict__ out, int* __restrict__ status) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * kCellFr) return;
    fr_t v;
    load_be32(cells + (size_t)gid * 32, v.v);
    if (!lt_r(v.v)) status[gid / kCellFr] = 1;
    store_field(out + (size_t)gid * 32, v.to_mont());
}
// aggregated_column_cells (das.rs:791-803) in bit-reversed order inside each column: one CTA per column, thread k
__global__ void __launch_bounds__(64) k_vc_columns(const uint8_t* __restrict__ cells_mont, const uint32_t* __restrict__ cell_idx,
                                                   const uint8_t* __restrict__ rp, int n, uint8_t* __restrict__ agg) {
    const int col = blockIdx.x, k = threadIdx.x;
    fr_t acc = fr_t::zero();
    for (int i = 0; i < n; i++)
        if (cell_idx[i] == (uint32_t)col)
            acc = acc + load_field<fr_t>(cells_mont + ((size_t)i * kCellFr + k) * 32) * load_field<fr_t>(rp + (size_t)i * 32);
    store_field(agg + ((size_t)col * kCellFr + (__brev((unsigned)k) >> 26)) * 32, acc);
}
// terms of the two lincombs.  Segment 0: (proof_i, r^i).  Segment 1: (C_j, w_j), (g1_monomial[k], -poly_k),
// (proof_i, r^i h_i^64); L = n + m + 64, both segments padded to L.
//   w_j = sum of r^i over cells of commitment j                                   (das.rs:697-741)
//   poly_k = sum over columns c of INTT64(column c)[k] * roots[8192 - brev7(c)]^k   (das.rs:805-833, 743-778)
//   h_i^64 = roots[brev7(cell_idx_i) * 64]                                        (das.rs:844-880)
__global__ void __launch_bounds__(256) k_vc_terms(const uint8_t* __restrict__ comm_aff, int m, const uint32_t* __restrict__ comm_idx,
                                                  const uint32_t* __restrict__ cell_idx, const uint8_t* __restrict__ proof_aff,
                                                  const uint8_t* __restrict__ rp, const uint8_t* __restrict__ col_poly,
                                                  const uint8_t* __restrict__ roots, const uint8_t* __restrict__ monomial_jac, int n,
                                                  uint8_t* __restrict__ pts, uint8_t* __restrict__ scalars) {
    const size_t L = (size_t)n + m + kCellFr;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L) return;
    affine_t inf{fp_t::zero(), fp_t::zero()};
    affine_t p0 = inf, p1;
    fr_t s0 = fr_t::zero(), s1;
    if (t < (size_t)n) {
        fr_t r = load_field<fr_t>(rp + t * 32);
        p0 = p1 = load_affine(proof_aff + t * 96);
        s0 = r;
        uint32_t hb = __brev(cell_idx[t]) >> 25;
        s1 = r * load_field_ro<fr_t>(roots + (size_t)hb * kCellFr * 32);
    } else if (t < (size_t)n + m) {
        const uint32_t j = (uint32_t)(t - n);
        fr_t w = fr_t::zero();
        for (int i = 0; i < n; i++)
            if (comm_idx[i] == j) w = w + load_field<fr_t>(rp + (size_t)i * 32);
        p1 = load_affine(comm_aff + (size_t)j * 96);
        s1 = w;
    } else {
        const uint32_t k = (uint32_t)(t - n - m);
        fr_t acc = fr_t::zero();
        for (int c = 0; c < kCells; c++) {
            uint32_t cb = __brev((unsigned)c) >> 25;
            uint32_t e = ((uint32_t)(kExt - cb) * k) & (kExt - 1);   // roots[8192 - cb]^k = roots[(8192 - cb) k mod 8192]
            acc = acc + load_field<fr_t>(col_poly + ((size_t)c * kCellFr + k) * 32) * load_field_ro<fr_t>(roots + (size_t)e * 32);
        }
        p1 = load_affine(monomial_jac + (size_t)k * 144);            // blst_p1 with z = 1: the first 96 bytes are the affine point
        s1 = acc.neg();
    }
    store_affine(pts + t * 96, p0);
    store_field(scalars + t * 32, s0.from_mont());
    store_affine(pts + (L + t) * 96, p1);
    store_field(scalars + (L + t) * 32, s1.from_mont());
}

// argument checks of compute_verify_cell_kzg_proof_batch_challenge (blst/src/eip_7594.rs:50-80): every commitment and
// proof must decode (no subgroup check there), every cell element must be canonical; status[0] = 1 otherwise
__global__ void k_any_flag(const int* __restrict__ flags, int n, int* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) status[0] = 1;
}
size_t check_challenge_ws_bytes(int m, int n) {
    return (size_t)(m + n) * 96 + (size_t)n * kCellFr * 32 + (size_t)(m + 2 * n + 1) * sizeof(int);
}
// context-free (the reference's compute_verify_cell_kzg_proof_batch_challenge takes no settings, blst/src/eip_7594.rs:35-97):
// w = device workspace of check_challenge_ws_bytes(m, n) bytes
void launch_check_challenge_inputs(uint8_t* w, const uint8_t* commitments48, int m, const uint8_t* cells, const uint8_t* proofs48, int n,
                                   int* status, cudaStream_t st) {
    uint8_t* aff = w;
    uint8_t* cells_m = w + (size_t)(m + n) * 96;
    int* flags = reinterpret_cast<int*>(cells_m + (size_t)n * kCellFr * 32);
    const int nf = m + 2 * n;
    B200_CUDA_CHECK(cudaMemsetAsync(flags, 0, (size_t)(nf + 1) * sizeof(int), st));
    if (m > 0) launch_uncompress_g1(commitments48, aff, flags, m, st);
    if (n > 0) {
        launch_uncompress_g1(proofs48, aff + (size_t)m * 96, flags + m, n, st);
        k_vc_cells<<<div_up((size_t)n * kCellFr, 256), 256, 0, st>>>(cells, n, cells_m, flags + m + n);
        B200_LAUNCH_CHECK();
    }
    if (nf > 0) {
        k_any_flag<<<div_up(nf, 256), 256, 0, st>>>(flags, nf, status);
        B200_LAUNCH_CHECK();
    }
}
void KzgSettingsDev::check_challenge_inputs(const uint8_t* commitments48, int m, const uint8_t* cells, const uint8_t* proofs48, int n,
                                            int* status, cudaStream_t st) {
    launch_check_challenge_inputs(ensure_das_ws(check_challenge_ws_bytes(m, n)), commitments48, m, cells, proofs48, n, status, st);
}

void KzgSettingsDev::verify_cells(const uint8_t* commitments48, int m, const uint32_t* comm_idx, const uint32_t* cell_idx,
                                  const uint8_t* cells, const uint8_t* proofs48, const uint8_t* r32, int n, int* status, int* result,
                                  cudaStream_t st) {
    if (!g2_lines_) throw CudaError(-1, "trusted setup was loaded without G2 points");
    if (n < 1 || m < 1 || m > n) throw CudaError(-1, "verify_cells: bad counts");
    const size_t L = (size_t)n + m + kCellFr, blocks = (2 * L + 7) / 8;   // lincomb2_and_pair may use two quads per term
    const size_t colb = (size_t)kCells * kCellFr * 32;
    size_t bytes = (size_t)m * 96 + (size_t)n * 96 + (size_t)n * kCellFr * 32 + (size_t)n * 32 + 64 + 2 * colb + 2 * L * (96 + 32) +
                   2 * blocks * 192 + 2 * 192 + 4 * kMillerLines * kLineBytes + 1024;
    uint8_t* w = ensure_das_ws(bytes);
    uint8_t* proof_aff = w;   w += (size_t)n * 96;
    uint8_t* comm_aff = w;    w += (size_t)m * 96;
    uint8_t* cells_m = w;     w += (size_t)n * kCellFr * 32;
    uint8_t* rp = w;          w += (size_t)n * 32;
    uint8_t* r = w;           w += 64;
    uint8_t* agg = w;         w += colb;
    uint8_t* colp = w;        w += colb;
    uint8_t* pts = w;         w += 2 * L * 96;
    uint8_t* scalars = w;     w += 2 * L * 32;
    uint8_t* partials = w;    w += 2 * blocks * 192;
    uint8_t* sums = w;        w += 2 * 192;
    uint8_t* scratch = (uint8_t*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    // commitments report into status[0..m) (m <= n), proofs and cells into status[i]
    if (commitments48 == proofs48 + (size_t)n * 48) {
        launch_decode_g1_checked(proofs48, proof_aff, status, n + m, st, n);   // one launch; commitment j -> status[j]
    } else {
        launch_decode_g1_checked(commitments48, comm_aff, status, m, st);
        launch_decode_g1_checked(proofs48, proof_aff, status, n, st);
    }
    k_vc_cells<<<div_up((size_t)n * kCellFr, 256), 256, 0, st>>>(cells, n, cells_m, status);
    launch_fr_from_bytes(r32, 1, 1, r, status, st);
    k_vc_powers<<<div_up(n, 256), 256, 0, st>>>(r, n, rp);
    k_vc_columns<<<kCells, kCellFr, 0, st>>>(cells_m, cell_idx, rp, n, agg);
    B200_LAUNCH_CHECK();
    fs_->fft_fr(agg, colp, kCellFr, true, kCells, st);   // 128 inverse transforms of 64 points; unused columns are zero
    k_vc_terms<<<div_up(L, 256), 256, 0, st>>>(comm_aff, m, comm_idx, cell_idx, proof_aff, rp, colp, (const uint8_t*)fs_->roots_dev(),
                                             (const uint8_t*)monomial_jac_, n, pts, scalars);
    B200_LAUNCH_CHECK();
    // e(final, G2) == e(proof_lincomb, [s^64]G2)   (das.rs:381-387)
    lincomb2_and_pair(pts, scalars, L, partials, sums, scratch, 2, 0, result, st);
}

}  // namespace b200
