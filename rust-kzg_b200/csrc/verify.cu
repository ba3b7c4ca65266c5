// verify.cu -- the verification side of EIP-4844 on the device: verify_kzg_proof, verify_blob_kzg_proof and
// verify_blob_kzg_proof_batch (kzg/src/eip_4844.rs:328-435, 586-866; pairing: blst/src/kzg_proofs.rs:74-100).
//
// One code path serves all three.  With r_i the Fiat-Shamir powers (r_0 = 1, so n = 1 is the single check of
// blst/src/types/kzg_settings.rs:178-196):
//      A = sum r_i proof_i
//      B = sum r_i C_i  +  sum (r_i z_i) proof_i  -  (sum r_i y_i) G1           (the reference forms C_i - [y_i]G1 with n
//                                                                                scalar multiplications; same point)
//      accept  <=>  e(A, [s]G2) == e(B, G2)   <=>   e(-A, [s]G2) e(B, G2) == 1
// Both G2 arguments are fixed points of the setup: their Miller-loop lines are tabulated at load time (pairing.cuh).
// The two linear combinations are short (n and 2n + 1 terms) and latency-bound, so they run as one scalar
// multiplication per lane quad (g1_quad.cuh) followed by quad trees, not through the bucket engine.
#include "eip4844.cuh"
#include "g1.cuh"
#include "g1_quad.cuh"
#include "pairing.cuh"
#include "util.cuh"

#include <vector>

namespace b200 {

// ---- G2 setup points --------------------------------------------------------------------------------------------
// affine out: 4 Fp per point (x.re, x.im, y.re, y.im); jac out: blst_p2 {x, y, z} with z = 1 (0 for infinity)
__global__ void __launch_bounds__(32) k_g2_decode(const uint8_t* __restrict__ in96, int n, uint8_t* __restrict__ affine,
                                                  uint8_t* __restrict__ jac, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp2_t x, y;
    if (!g2_uncompress(in96 + (size_t)i * 96, x, y)) flags[i] = 1;
    store_fp2(affine + (size_t)i * 192, x);
    store_fp2(affine + (size_t)i * 192 + 96, y);
    bool inf = x.is_zero() && y.is_zero();
    store_fp2(jac + (size_t)i * 288, x);
    store_fp2(jac + (size_t)i * 288 + 96, y);
    store_fp2(jac + (size_t)i * 288 + 192, inf ? fp2_t::zero() : fp2_t::one());
}
// line tables for the points idx[0..m) of the affine array; table t at lines + t * 68 * 288
struct G2PrepArgs { int idx[4]; int m; };
__global__ void __launch_bounds__(32) k_g2_prepare(const uint8_t* __restrict__ affine, G2PrepArgs args, uint8_t* __restrict__ lines) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= args.m) return;
    const uint8_t* q = affine + (size_t)args.idx[t] * 192;
    g2_prepare_lines(load_fp2(q), load_fp2(q + 96), lines + (size_t)t * kMillerLines * kLineBytes);
}

// ---- terms of the two linear combinations -------------------------------------------------------------------------
// Segment 0 (A) and segment 1 (B), each padded to L = 2n + 1 terms: points affine (96 B), scalars canonical (32 B).
// One thread per item: r^i by square-and-multiply on i (compute_powers, kzg/src/eip_4844.rs:316-326, without the serial
// chain); the G1 term's scalar -sum r^i y_i is reduced by one CTA afterwards.
__device__ __forceinline__ fr_t fr_pow_u32(fr_t base, uint32_t e) {
    fr_t acc = fr_t::one();
    while (e) {
        if (e & 1) acc = acc * base;
        base = base.sqr();
        e >>= 1;
    }
    return acc;
}
__global__ void __launch_bounds__(128) k_verify_terms(const uint8_t* __restrict__ comm_aff, const uint8_t* __restrict__ proof_aff,
                                                      const uint8_t* __restrict__ z_mont, const uint8_t* __restrict__ y_mont,
                                                      const uint8_t* __restrict__ r_mont, int n, uint8_t* __restrict__ pts,
                                                      uint8_t* __restrict__ scalars, uint8_t* __restrict__ ry) {
    const size_t L = 2 * (size_t)n + 1;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    affine_t inf{fp_t::zero(), fp_t::zero()};
    if (i < (size_t)n) {
        fr_t r = n > 1 ? load_field<fr_t>(r_mont) : fr_t::one();
        fr_t rp = fr_pow_u32(r, (uint32_t)i);
        affine_t c = load_affine(comm_aff + i * 96), p = load_affine(proof_aff + i * 96);
        fr_t z = load_field<fr_t>(z_mont + i * 32), y = load_field<fr_t>(y_mont + i * 32);
        fr_t rpc = rp.from_mont();
        store_affine(pts + i * 96, p);
        store_field(scalars + i * 32, rpc);
        store_affine(pts + (L + i) * 96, c);
        store_field(scalars + (L + i) * 32, rpc);
        store_affine(pts + (L + n + i) * 96, p);
        store_field(scalars + (L + n + i) * 32, (rp * z).from_mont());
        store_field(ry + i * 32, rp * y);
    } else {
        store_affine(pts + i * 96, inf);                      // padding of segment 0
        store_field(scalars + i * 32, fr_t::zero());
    }
}
// scalar of the generator term: -(sum of ry[0..n)), one CTA
__global__ void __launch_bounds__(256) k_verify_gen_term(const uint8_t* __restrict__ ry, int n, uint8_t* __restrict__ pts,
                                                         uint8_t* __restrict__ scalars) {
    __shared__ __align__(16) uint8_t sh[8 * 32];
    fr_t acc = fr_t::zero();
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc = acc + load_field<fr_t>(ry + (size_t)i * 32);
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        fr_t o;
#pragma unroll
        for (int k = 0; k < 8; k++) o.v[k] = __shfl_down_sync(0xffffffffu, acc.v[k], d);
        acc = acc + o;
    }
    if ((threadIdx.x & 31) == 0) store_field(sh + (threadIdx.x >> 5) * 32, acc);
    __syncthreads();
    if (threadIdx.x == 0) {
        fr_t total = load_field<fr_t>(sh);
        for (int w = 1; w < 8; w++) total = total + load_field<fr_t>(sh + w * 32);
        const size_t L = 2 * (size_t)n + 1;
        affine_t g{fp_cast<fp_t>(pf_const(G1_GEN_AFFINE[0])), fp_cast<fp_t>(pf_const(G1_GEN_AFFINE[1]))};
        store_affine(pts + (L + 2 * (size_t)n) * 96, g);
        store_field(scalars + (L + 2 * (size_t)n) * 32, total.neg().from_mont());
    }
}
// One lane quad per term: partial[seg][block] = sum of the block's eight [k_i] P_i, in XYZZ (192 B).
// SPLIT (few terms: a single verification): two quads per term, one GLV half of the scalar multiplication each -- a 20 %
// shorter chain; the block's tree sums the halves like any other summands.
template <bool SPLIT>
__global__ void __launch_bounds__(32) k_lincomb_quads(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ scalars, int L,
                                                      uint8_t* __restrict__ partial) {
    __shared__ __align__(16) uint8_t table[kQuadTableBytes];
    const int q2 = blockIdx.x * 8 + (threadIdx.x >> 2);
    const int q = SPLIT ? q2 >> 1 : q2;
    const int role = threadIdx.x & 3;
    const bool live = q < L;
    const size_t t = (size_t)blockIdx.y * L + (live ? q : 0);
    fp_t x = load_field<fp_t>(pts + t * 96), y = load_field<fp_t>(pts + t * 96 + 48);
    const bool inf = !live || (x.is_zero() && y.is_zero());
    fp_t comp = role == 0 ? x : role == 1 ? y : fp_t::one();
    if (inf) comp = fp_t::zero();
    fr_t k = load_field<fr_t>(scalars + t * 32);
    comp = SPLIT ? quad_mul_scalar_half(comp, k.v, table, q2 & 1) : quad_mul_scalar(comp, k.v, table);
    comp = quad_tree(comp, 32);
    if (threadIdx.x < 4) store_field(partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 192 + quad_store_offset(), comp);
}
// out[seg] = sum of m partials (XYZZ), one warp per segment
__global__ void __launch_bounds__(32) k_quad_sum(const uint8_t* __restrict__ partial, int m, uint8_t* __restrict__ out) {
    const uint8_t* p = partial + (size_t)blockIdx.x * m * 192;
    fp_t acc = fp_t::zero();
    for (int base = 0; base < m; base += 8) {
        int q = base + (threadIdx.x >> 2);
        fp_t c = q < m ? load_field<fp_t>(p + (size_t)q * 192 + quad_store_offset()) : fp_t::zero();
        acc = quad_add(acc, c);
    }
    acc = quad_tree(acc, 32);
    if (threadIdx.x < 4) store_field(out + (size_t)blockIdx.x * 192 + quad_store_offset(), acc);
}
// Jacobian -> XYZZ for pairings_verify's operands
__global__ void k_jac_to_xyzz2(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, uint8_t* __restrict__ out) {
    if (threadIdx.x >= 2) return;
    cc::xyzz_t p = cc::jac_to_xyzz(cc::load_jac(threadIdx.x ? b : a));
    cc::store_xyzz(out + threadIdx.x * 192, p);
}

// ---- the pairing check: one warp --------------------------------------------------------------------------------
struct PairingArgs {
    const uint8_t* g1_xyzz[4];   // P_i, XYZZ
    const uint8_t* lines[4];     // line table of Q_i
    int neg[4];                  // use -P_i
    int n;
};
static constexpr int kPairScratchBytes = 4 * kMillerLines * kLineBytes;  // scaled lines of up to four pairs
__global__ void __launch_bounds__(kPairThreads) k_pairing_check(PairingArgs args, uint8_t* __restrict__ scratch, int* __restrict__ result) {
    __shared__ __align__(16) pf_t f[12];
    __shared__ __align__(16) pf_t ws[kW12FinalExpScratch];
    const int tid = threadIdx.x;
    bool skip[4];
    // evaluate every line at P_i, one (pair, line) per thread.  P_i = (X/ZZ, Y/ZZZ) is NOT normalised: the
    // line is scaled by ZZ*ZZZ in Fp instead -- (l2 ZZ ZZZ, l1 X ZZZ, l0 Y ZZ) -- and factors from Fp* vanish in the final
    // exponentiation ((p^12 - 1)/r is a multiple of p - 1), which saves a field inversion per pair.
    for (int p = 0; p < args.n; p++) skip[p] = load_field<pf_t>(args.g1_xyzz[p] + 144).is_zero();
    for (int w = tid; w < args.n * kMillerLines; w += kPairThreads) {
        const int p = w / kMillerLines, idx = w - p * kMillerLines;
        const uint8_t* g = args.g1_xyzz[p];
        pf_t X = load_field<pf_t>(g), Y = load_field<pf_t>(g + 48), ZZZ = load_field<pf_t>(g + 96), ZZ = load_field<pf_t>(g + 144);
        pf_t px = X * ZZZ, py = Y * ZZ, sc = ZZ * ZZZ;
        if (args.neg[p]) py = py.neg();
        const uint8_t* l = args.lines[p] + (size_t)idx * kLineBytes;
        uint8_t* o = scratch + ((size_t)p * kMillerLines + idx) * kLineBytes;
        store_fp2(o, load_fp2(l + 192).scale(sc));
        store_fp2(o + 96, load_fp2(l + 96).scale(px));
        store_fp2(o + 192, load_fp2(l).scale(py));
    }
    __syncthreads();
    pf_t* prod = ws + 132;
    w12_set_one(f);
    int idx = 0;
#pragma unroll 1
    for (int b = 61; b >= -1; b--) {
        const int steps = (b >= 0 && ((kBlsXHalf >> b) & 1)) ? 2 : 1;
#pragma unroll 1
        for (int s = 0; s < steps; s++, idx++)
            for (int p = 0; p < args.n; p++)
                if (!skip[p]) w12_mul_sparse(f, f, reinterpret_cast<const pf_t*>(scratch + ((size_t)p * kMillerLines + idx) * kLineBytes), prod);
        if (b >= 0) w12_mul(f, f, f, prod);
    }
    w12_conj(f, f);
    w12_final_exp(f, ws);
    bool one = w12_is_one(f);
    if (tid == 0) *result = one ? 1 : 0;
}

// test hook: out (Jacobian, device) = sum k_i P_i through the quad scalar multiplication path (affine points, Montgomery scalars)
__global__ void k_mont_to_canon(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) store_field(out + (size_t)i * 32, load_field<fr_t>(in + (size_t)i * 32).from_mont());
}
__global__ void __launch_bounds__(32) k_xyzz_to_jac1(const uint8_t* __restrict__ in, uint8_t* __restrict__ out) {
    if (threadIdx.x == 0) cc::store_jac(out, cc::xyzz_to_jac(cc::load_xyzz(in)));
}
void selftest_lincomb_quads(const void* points_affine_dev, const void* scalars_mont_dev, int n, void* out_jac_dev, cudaStream_t st) {
    const int blocks = (n + 7) / 8;
    uint8_t* canon = dev_alloc<uint8_t>((size_t)n * 32);
    uint8_t* partials = dev_alloc<uint8_t>((size_t)blocks * 192 + 192);
    k_mont_to_canon<<<div_up(n, 128), 128, 0, st>>>((const uint8_t*)scalars_mont_dev, canon, n);
    k_lincomb_quads<false><<<dim3((unsigned)blocks, 1), 32, 0, st>>>((const uint8_t*)points_affine_dev, canon, n, partials);
    k_quad_sum<<<1, 32, 0, st>>>(partials, blocks, partials + (size_t)blocks * 192);
    k_xyzz_to_jac1<<<1, 32, 0, st>>>(partials + (size_t)blocks * 192, (uint8_t*)out_jac_dev);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(canon);
    cudaFree(partials);
    B200_CUDA_CHECK(e);
    B200_LAUNCH_CHECK();
}

// ---- host side ----------------------------------------------------------------------------------------------------
void KzgSettingsDev::load_g2(const uint8_t* g2_monomial, int count, cudaStream_t st) {
    if (count < 65) throw CudaError(1, "Invalid number of g2 points in trusted setup");
    uint8_t* comp = dev_alloc<uint8_t>((size_t)count * 96);
    int* flags = dev_alloc<int>(count);
    g2_affine_ = dev_alloc<uint8_t>((size_t)count * 192);
    g2_jac_ = dev_alloc<uint8_t>((size_t)count * 288);
    g2_lines_ = dev_alloc<uint8_t>((size_t)3 * kMillerLines * kLineBytes);
    B200_CUDA_CHECK(cudaMemcpyAsync(comp, g2_monomial, (size_t)count * 96, cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemsetAsync(flags, 0, count * sizeof(int), st));
    k_g2_decode<<<div_up(count, 32), 32, 0, st>>>(comp, count, (uint8_t*)g2_affine_, (uint8_t*)g2_jac_, flags);
    G2PrepArgs pa{{0, 1, 64, 0}, 3};
    k_g2_prepare<<<1, 32, 0, st>>>((const uint8_t*)g2_affine_, pa, (uint8_t*)g2_lines_);
    B200_LAUNCH_CHECK();
    std::vector<int> h(count);
    B200_CUDA_CHECK(cudaMemcpyAsync(h.data(), flags, count * sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(comp);
    cudaFree(flags);
    for (int f : h)
        if (f) {
            cudaFree(g2_lines_);
            g2_lines_ = nullptr;
            throw CudaError(1, "Failed to uncompress");  // FsG2::from_bytes (blst/src/types/g2.rs:66)
        }
}

// workspace layout for n items (L = 2n + 1):
//   [comm_aff n*96][proof_aff n*96][z n*32][y n*32][r 32][ry n*32][pts 2L*96][scalars 2L*32][partials 2*ceil(2L/8)*192][sums 2*192]
//   [pair scratch]
void KzgSettingsDev::ensure_verify_ws(size_t n) {
    if (n <= vf_cap_ && vf_buf_) return;
    cudaFree(vf_buf_);
    size_t L = 2 * n + 1, blocks = (2 * L + 7) / 8;   // lincomb2_and_pair may use two quads per term
    size_t bytes = n * (96 + 96 + 32 + 32 + 32) + 64 + 2 * L * (96 + 32) + 2 * blocks * 192 + 2 * 192 + kPairScratchBytes + 256;
    vf_buf_ = dev_alloc<uint8_t>(bytes);
    vf_cap_ = n;
}

static void run_pairing(const uint8_t* p0, const uint8_t* l0, int neg0, const uint8_t* p1, const uint8_t* l1, int neg1, uint8_t* scratch,
                        int* result, cudaStream_t st) {
    PairingArgs a{};
    a.n = 2;
    a.g1_xyzz[0] = p0; a.lines[0] = l0; a.neg[0] = neg0;
    a.g1_xyzz[1] = p1; a.lines[1] = l1; a.neg[1] = neg1;
    k_pairing_check<<<1, kPairThreads, 0, st>>>(a, scratch, result);
    B200_LAUNCH_CHECK();
}

// segment 0 = pts[0..L), segment 1 = pts[L..2L) (affine, canonical scalars): result = e(-S0, Q[qa]) e(S1, Q[qb]) == 1
void KzgSettingsDev::lincomb2_and_pair(const uint8_t* pts, const uint8_t* scalars, size_t L, uint8_t* partials, uint8_t* sums,
                                       uint8_t* scratch, int qa, int qb, int* result, cudaStream_t st) {
    // few terms (single verifications): two quads per term, one GLV half each (the launch is a handful of warps either way)
    const bool split = L <= 256;
    const size_t blocks = split ? (2 * L + 7) / 8 : (L + 7) / 8;
    if (split) k_lincomb_quads<true><<<dim3((unsigned)blocks, 2), 32, 0, st>>>(pts, scalars, (int)L, partials);
    else k_lincomb_quads<false><<<dim3((unsigned)blocks, 2), 32, 0, st>>>(pts, scalars, (int)L, partials);
    k_quad_sum<<<2, 32, 0, st>>>(partials, (int)blocks, sums);
    B200_LAUNCH_CHECK();
    const uint8_t* lines = (const uint8_t*)g2_lines_;
    const size_t tb = (size_t)kMillerLines * kLineBytes;
    run_pairing(sums, lines + qa * tb, 1, sums + 192, lines + qb * tb, 0, scratch, result, st);
}

// decode + subgroup-check the 2n points into the workspace (may run early, on another stream, while z / y are still
// being produced); one launch when the two arrays are contiguous
void KzgSettingsDev::verify_decode(const uint8_t* commitments48, const uint8_t* proofs48, int n, int* status, cudaStream_t st,
                                   cudaEvent_t decoded) {
    ensure_verify_ws(n);
    uint8_t* comm_aff = (uint8_t*)vf_buf_;
    uint8_t* proof_aff = comm_aff + (size_t)n * 96;
    const bool contiguous = proofs48 == commitments48 + (size_t)n * 48;
    if (!decoded) {
        if (contiguous) {
            launch_decode_g1_checked(commitments48, comm_aff, status, 2 * n, st, n);
        } else {
            launch_decode_g1_checked(commitments48, comm_aff, status, n, st);
            launch_decode_g1_checked(proofs48, proof_aff, status, n, st);
        }
        return;
    }
    // coordinates first (a square-root chain per point), the subgroup ladders after the event
    if (contiguous) {
        launch_decode_g1_unchecked(commitments48, comm_aff, status, 2 * n, st, n);
    } else {
        launch_decode_g1_unchecked(commitments48, comm_aff, status, n, st);
        launch_decode_g1_unchecked(proofs48, proof_aff, status, n, st);
    }
    B200_CUDA_CHECK(cudaEventRecord(decoded, st));
    launch_subgroup_g1(comm_aff, status, 2 * n, st, n);      // comm_aff and proof_aff are adjacent in the workspace
}

void KzgSettingsDev::verify_batch(const uint8_t* commitments48, const uint8_t* proofs48, const uint8_t* z32, const uint8_t* y32,
                                  int z_reduce, const uint8_t* r32, int n, int* status, int* result, cudaStream_t st, bool skip_decode) {
    if (!g2_lines_) throw CudaError(-1, "trusted setup was loaded without G2 points");
    if (n < 1) throw CudaError(-1, "verify_batch needs at least one item");
    ensure_verify_ws(n);
    const size_t L = 2 * (size_t)n + 1, blocks = (2 * L + 7) / 8;
    uint8_t* w = (uint8_t*)vf_buf_;
    uint8_t* comm_aff = w;                 w += (size_t)n * 96;
    uint8_t* proof_aff = w;                w += (size_t)n * 96;
    uint8_t* z = w;                        w += (size_t)n * 32;
    uint8_t* y = w;                        w += (size_t)n * 32;
    uint8_t* r = w;                        w += 64;
    uint8_t* ry = w;                       w += (size_t)n * 32;
    uint8_t* pts = w;                      w += 2 * L * 96;
    uint8_t* scalars = w;                  w += 2 * L * 32;
    uint8_t* partials = w;                 w += 2 * blocks * 192;
    uint8_t* sums = w;                     w += 2 * 192;
    uint8_t* scratch = (uint8_t*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    if (!skip_decode) verify_decode(commitments48, proofs48, n, status, st);
    launch_fr_from_bytes(z32, n, z_reduce, z, status, st);
    launch_fr_from_bytes(y32, n, 0, y, status, st);
    if (n > 1) launch_fr_from_bytes(r32, 1, 1, r, status, st);  // hash_to_bls_field never fails: status untouched
    k_verify_terms<<<div_up(L, 128), 128, 0, st>>>(comm_aff, proof_aff, z, y, r, n, pts, scalars, ry);
    k_verify_gen_term<<<1, 256, 0, st>>>(ry, n, pts, scalars);
    B200_LAUNCH_CHECK();
    lincomb2_and_pair(pts, scalars, L, partials, sums, scratch, 1, 0, result, st);
    launches_ = 9 + (n > 1);
}

void KzgSettingsDev::pairings_verify(const void* a1_jac, int qa, const void* b1_jac, int qb, int* result, cudaStream_t st) {
    if (!g2_lines_) throw CudaError(-1, "trusted setup was loaded without G2 points");
    if (qa < 0 || qa > 2 || qb < 0 || qb > 2) throw CudaError(-1, "bad G2 table index");
    ensure_verify_ws(1);
    uint8_t* w = (uint8_t*)vf_buf_;
    uint8_t* sums = w;
    uint8_t* scratch = (uint8_t*)(((uintptr_t)(w + 2 * 192) + 255) & ~(uintptr_t)255);
    k_jac_to_xyzz2<<<1, 32, 0, st>>>((const uint8_t*)a1_jac, (const uint8_t*)b1_jac, sums);
    const uint8_t* lines = (const uint8_t*)g2_lines_;
    const size_t tb = (size_t)kMillerLines * kLineBytes;
    run_pairing(sums, lines + qa * tb, 1, sums + 192, lines + qb * tb, 0, scratch, result, st);
}

}  // namespace b200
