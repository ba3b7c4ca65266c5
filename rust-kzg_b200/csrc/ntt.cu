// ntt.cu -- Fr NTT / inverse NTT / DAS extension kernels for sm_100a.  See ntt.cuh for the decomposition.
#include "ntt.cuh"

#include <algorithm>
#include <cstdlib>

#include <cooperative_groups.h>

#include "mont.cuh"
#include "util.cuh"

namespace b200 {

namespace cg = cooperative_groups;

static constexpr int kNttThreads = 256;
static constexpr int kMaxLogM = 11;          // largest in-CTA sub-transform: 2^11 points = 64 KiB of shared memory
static constexpr int kTileElems = 1 << 11;   // G * m points per CTA

struct NttPass {
    const uint8_t* in;
    uint8_t* out;
    int log_m;              // sub-transform size m = 2^log_m
    int G;                  // columns handled by one CTA
    size_t ncols;           // columns per transform
    size_t in_cstride, in_rstride, out_cstride, out_rstride;  // element strides
    int in_col_fast, out_col_fast;  // which index runs fastest across consecutive threads (coalescing)
    const uint8_t* roots;   // w^0 .. w^nmax
    size_t nmax;
    size_t unit_m;          // nmax / m
    int inverse;
    size_t tw_unit;         // 0: none; else outputs (q, col) are multiplied by w_N^(q*col), w_N = roots[tw_unit]
    size_t twist_unit;      // 0: none; else input element with natural index j is multiplied by roots[j*twist_unit]
    const uint8_t* scale;   // nullptr or one Fr multiplied into every output (n^-1)
    size_t batch_stride;      // elements between consecutive transforms of the batch (input)
    size_t out_batch_stride;  // ... and in the output (differs when the transforms are the rows of a larger one)
};

__device__ __forceinline__ fr_t root_at(const NttPass& p, size_t e_units) {
    // e_units = exponent * unit, 0 <= e_units <= nmax
    size_t idx = p.inverse ? p.nmax - e_units : e_units;
    return load_field_ro<fr_t>(p.roots + idx * 32);
}

// Shared-memory slot of tile element i.  The XOR swizzle spreads the 8-element groups a thread owns in the first
// butterfly group (stride-256-byte accesses, otherwise an 8-way bank conflict) over the banks.
__device__ __forceinline__ int slot(int i) { return i ^ ((i >> 3) & 7); }

// R consecutive radix-2 DIT stages (t .. t+R-1) on the 2^R elements whose in-column indices differ only in bits
// t .. t+R-1, entirely in registers: one shared-memory round trip and one barrier per R stages instead of per stage.
// The butterfly is the reference's (blst/src/fft_fr.rs:98-103): lo' = lo + w*hi, hi' = lo - w*hi.
// Stage twiddles w_m^k, k < m/2, come from a per-CTA copy in shared memory (filled once per CTA, shared by its G
// columns): seven of them are consumed per 3-stage group, and fetching them through L1/L2 showed up as the largest
// stall (long scoreboard, ncu r01) of a kernel that is otherwise bound by the multiply pipe.
template <int R>
__device__ __forceinline__ void butterfly_group(uint8_t* sm, const uint8_t* tw, const NttPass& p, int m, int t, int unit) {
    constexpr int E = 1 << R;
    const int units_per_col = m >> R;
    const int g = unit / units_per_col, rem = unit - g * units_per_col;
    const int lo = rem & ((1 << t) - 1), hi = rem >> t;
    const int base = (hi << (t + R)) | lo;
    fr_t x[E];
#pragma unroll
    for (int j = 0; j < E; j++) x[j] = load_field<fr_t>(sm + (size_t)slot(g * m + base + (j << t)) * 32);
#pragma unroll
    for (int s = 0; s < R; s++) {
        // stage t+s pairs j and j + 2^s; the twiddle exponent is (index mod 2^(t+s)) * m / 2^(t+s+1)
#pragma unroll
        for (int b = 0; b < (1 << s); b++) {
            const int low = lo + (b << t);  // index mod 2^(t+s) of the pairs whose low s group-bits equal b
            fr_t w;
            const bool has_w = low != 0;
            if (has_w) w = load_field<fr_t>(tw + ((size_t)low << (p.log_m - 1 - (t + s))) * 32);
#pragma unroll
            for (int j = 0; j < E; j++) {
                if ((j & ((1 << (s + 1)) - 1)) != b) continue;  // j has bit s clear and low s bits == b
                fr_t u = x[j], v = x[j + (1 << s)];
                if (has_w) v = v * w;
                x[j] = u + v;
                x[j + (1 << s)] = u - v;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < E; j++) store_field(sm + (size_t)slot(g * m + base + (j << t)) * 32, x[j]);
}

// all log_m stages of the G sub-transforms of a tile in shared memory: groups of three (RMAX = 3) or two stages, then
// whatever is left (two or one); ends with a barrier
template <int RMAX>
__device__ __forceinline__ void tile_butterflies(uint8_t* sm, const uint8_t* tw, const NttPass& p, int m, int tile) {
    int t = 0;
    if (RMAX >= 3) {
        for (; t + 3 <= p.log_m; t += 3) {
            for (int u = threadIdx.x; u < (tile >> 3); u += blockDim.x) butterfly_group<3>(sm, tw, p, m, t, u);
            __syncthreads();
        }
    } else {
        for (; t + 2 <= p.log_m - 1 || t + 2 == p.log_m; t += 2) {
            for (int u = threadIdx.x; u < (tile >> 2); u += blockDim.x) butterfly_group<2>(sm, tw, p, m, t, u);
            __syncthreads();
        }
    }
    if (p.log_m - t == 2) {
        for (int u = threadIdx.x; u < (tile >> 2); u += blockDim.x) butterfly_group<2>(sm, tw, p, m, t, u);
        __syncthreads();
    } else if (p.log_m - t == 1) {
        for (int u = threadIdx.x; u < (tile >> 1); u += blockDim.x) butterfly_group<1>(sm, tw, p, m, t, u);
        __syncthreads();
    }
}

template <int RMAX, int MINB>
__global__ void __launch_bounds__(kNttThreads, MINB) k_ntt_pass(NttPass p) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int m = 1 << p.log_m;
    const int tile = p.G * m;
    uint8_t* tw = sm + (size_t)tile * 32;
    for (int k = threadIdx.x; k < (m >> 1); k += blockDim.x) store_field(tw + (size_t)k * 32, root_at(p, (size_t)k * p.unit_m));
    const size_t col0 = (size_t)blockIdx.x * p.G;
    const uint8_t* in = p.in + (size_t)blockIdx.y * p.batch_stride * 32;
    uint8_t* out = p.out + (size_t)blockIdx.y * p.out_batch_stride * 32;

    // load (optionally twisted), bit-reversed row index into shared memory
    for (int e = threadIdx.x; e < tile; e += blockDim.x) {
        int g, r;
        if (p.in_col_fast) { g = e % p.G; r = e / p.G; } else { g = e >> p.log_m; r = e & (m - 1); }
        size_t col = col0 + g;
        fr_t x = fr_t::zero();
        if (col < p.ncols) {
            size_t j = col * p.in_cstride + (size_t)r * p.in_rstride;
            x = load_field_ro<fr_t>(in + j * 32);
            if (p.twist_unit) x = x * load_field_ro<fr_t>(p.roots + (j * p.twist_unit) * 32);
        }
        int rr = p.log_m ? (int)(__brev((unsigned)r) >> (32 - p.log_m)) : 0;
        store_field(sm + (size_t)slot(g * m + rr) * 32, x);
    }
    __syncthreads();

    tile_butterflies<RMAX>(sm, tw, p, m, tile);

    // store (optionally with the inter-pass twiddle and the 1/n scale)
    fr_t sc;
    if (p.scale) sc = load_field_ro<fr_t>(p.scale);
    for (int e = threadIdx.x; e < tile; e += blockDim.x) {
        int g, q;
        if (p.out_col_fast) { g = e % p.G; q = e / p.G; } else { g = e >> p.log_m; q = e & (m - 1); }
        size_t col = col0 + g;
        if (col >= p.ncols) continue;
        fr_t x = load_field<fr_t>(sm + (size_t)slot(g * m + q) * 32);
        if (p.tw_unit && q && col) x = x * root_at(p, (size_t)q * col * p.tw_unit);
        if (p.scale) x = x * sc;
        store_field(out + (col * p.out_cstride + (size_t)q * p.out_rstride) * 32, x);
    }
}

// Both passes of a small transform in ONE launch by a thread-block cluster of C = 8 or 16 CTAs: CTA r transforms the
// columns [r n1/C, (r+1) n1/C) (pass 1, p1), applies the twiddle w_n^(i2 j1) and writes every value straight into the shared
// memory of the CTA that owns row i2 in pass 2 (distributed shared memory: the inter-pass transpose never touches HBM),
// one cluster barrier, then CTA r transforms the rows [r n2/C, (r+1) n2/C) it now holds (p2) and writes natural order.
// One launch instead of two and no scratch round trip: the small-size latency is launches and dependent phases, not work.
static constexpr int kClusterMinLog = 9, kClusterMaxLog = 14;    // 2^14 on 8 CTAs: two tiles of 2048 points + twiddles = 132 KiB per CTA
static constexpr int kClusterMaxShm = 2 * (1 << (kClusterMaxLog - 3)) * 32 + 2 * (1 << 6) * 32;
__global__ void __launch_bounds__(kNttThreads) k_ntt_cluster(NttPass p1, NttPass p2) {
    extern __shared__ __align__(16) uint8_t sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank(), C = (int)cluster.num_blocks();   // 8 or 16 CTAs (launch attribute)
    const int m1 = 1 << p1.log_m, m2 = 1 << p2.log_m;           // n2 (column length), n1 (row length)
    const int G1 = m2 / C, G2 = m1 / C;                           // columns / rows per CTA
    const int tile = G1 * m1;                                     // = G2 * m2 = n / C elements
    uint8_t* tile_a = sm;
    uint8_t* tile_b = sm + (size_t)tile * 32;
    uint8_t* tw1 = tile_b + (size_t)tile * 32;
    uint8_t* tw2 = tw1 + (size_t)(m1 >> 1) * 32;
    for (int k = threadIdx.x; k < (m1 >> 1); k += blockDim.x) store_field(tw1 + (size_t)k * 32, root_at(p1, (size_t)k * p1.unit_m));
    for (int k = threadIdx.x; k < (m2 >> 1); k += blockDim.x) store_field(tw2 + (size_t)k * 32, root_at(p2, (size_t)k * p2.unit_m));
    const uint8_t* in = p1.in + (size_t)blockIdx.y * p1.batch_stride * 32;
    uint8_t* out = p2.out + (size_t)blockIdx.y * p2.out_batch_stride * 32;
    const size_t col0 = (size_t)r * G1;
    // pass 1 load: column fastest across threads (consecutive columns are consecutive in memory), bit-reversed row index
    for (int e = threadIdx.x; e < tile; e += blockDim.x) {
        const int g = e % G1, row = e / G1;
        const size_t j = (col0 + g) * p1.in_cstride + (size_t)row * p1.in_rstride;
        fr_t x = load_field_ro<fr_t>(in + j * 32);
        if (p1.twist_unit) x = x * load_field_ro<fr_t>(p1.roots + (j * p1.twist_unit) * 32);
        const int rr = (int)(__brev((unsigned)row) >> (32 - p1.log_m));
        store_field(tile_a + (size_t)slot(g * m1 + rr) * 32, x);
    }
    __syncthreads();
    tile_butterflies<2>(tile_a, tw1, p1, m1, tile);
    cluster.sync();                                               // every CTA of the cluster is running: its shared memory can be written
    // exchange: value (i2 = q, j1 = col) times w_n^(q col) goes to row q's owner, at the bit-reversed position of col in the row
    for (int e = threadIdx.x; e < tile; e += blockDim.x) {
        const int g = e % G1, q = e / G1;
        const size_t col = col0 + g;
        fr_t x = load_field<fr_t>(tile_a + (size_t)slot(g * m1 + q) * 32);
        if (q && col) x = x * root_at(p1, (size_t)q * col * p1.tw_unit);
        const int dst = q / G2, gl = q - dst * G2;
        const int cc = (int)(__brev((unsigned)col) >> (32 - p2.log_m));
        uint8_t* remote = cluster.map_shared_rank(tile_b, dst);
        store_field(remote + (size_t)slot(gl * m2 + cc) * 32, x);
    }
    cluster.sync();
    tile_butterflies<2>(tile_b, tw2, p2, m2, tile);
    // pass 2 store: row i2 = r G2 + g, output element i1 = q at X[i2 + n2 i1]
    fr_t sc;
    if (p2.scale) sc = load_field_ro<fr_t>(p2.scale);
    for (int e = threadIdx.x; e < tile; e += blockDim.x) {
        const int g = e % G2, q = e / G2;
        const size_t i2 = (size_t)r * G2 + g;
        fr_t x = load_field<fr_t>(tile_b + (size_t)slot(g * m2 + q) * 32);
        if (p2.scale) x = x * sc;
        store_field(out + (i2 * p2.out_cstride + (size_t)q * p2.out_rstride) * 32, x);
    }
}

// ---- settings: roots of unity on the device -------------------------------------------------------------------
// pw[b] = root^(2^b), b <= scale  (single thread)
__global__ void k_root_powers(const uint32_t* exp_words /*8*/, uint8_t* pw, int scale, uint8_t* inv_n /*32 Fr*/) {
    if (threadIdx.x || blockIdx.x) return;
    // SCALE2_ROOT_OF_UNITY[scale] = 7^((r-1)/2^scale) (blst/src/consts.rs:14-50); the exponent arrives precomputed
    fr_t seven = fr_t::zero();
    seven.v[0] = 7;
    seven = seven.to_mont();
    fr_t acc = fr_t::one();
    for (int i = 255; i >= 0; i--) {
        acc = acc.sqr();
        if ((exp_words[i >> 5] >> (i & 31)) & 1) acc = acc * seven;
    }
    for (int b = 0; b <= scale; b++) {
        store_field(pw + b * 32, acc);
        acc = acc.sqr();
    }
    // inv_n[k] = (2^k)^-1: halve repeatedly from 1 (multiply by 2^-1 = (r+1)/2)
    fr_t two = fr_t::one() + fr_t::one();
    fr_t half = two.inverse();
    fr_t cur = fr_t::one();
    for (int k = 0; k < 32; k++) {
        store_field(inv_n + k * 32, cur);
        cur = cur * half;
    }
}
// roots[i] = root^i for i in [0, n], by binary expansion over pw[]
__global__ void k_roots_table(const uint8_t* pw, uint8_t* roots, uint8_t* brp, size_t n, int scale) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    fr_t acc = fr_t::one();
    for (int b = 0; b <= scale; b++)
        if ((i >> b) & 1) acc = acc * load_field_ro<fr_t>(pw + b * 32);
    store_field(roots + i * 32, acc);
    if (i < n) {
        size_t r = scale ? (size_t)(__brevll((unsigned long long)i) >> (64 - scale)) : 0;
        store_field(brp + r * 32, acc);
    }
}

FFTSettingsDev::FFTSettingsDev(int scale, cudaStream_t st) : scale_(scale) {
    if (scale < 0 || scale >= 32) throw CudaError(1, "Scale is expected to be within root of unity matrix row size");
    max_width_ = (size_t)1 << scale;
    roots_ = dev_alloc<uint8_t>((max_width_ + 1) * 32 + 33 * 32 + 32 * 32 + 32);
    brp_roots_ = dev_alloc<uint8_t>(max_width_ * 32);
    // (r - 1) >> scale as 8 little-endian words (plain integer shifting, no field arithmetic on the host)
    uint32_t e[8] = {0x00000000u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    for (int s = 0; s < scale; s++)
        for (int i = 0; i < 8; i++) e[i] = (e[i] >> 1) | (i < 7 ? e[i + 1] << 31 : 0);
    uint8_t* tail = (uint8_t*)roots_ + (max_width_ + 1) * 32;  // [pw 33 Fr][inv_n 32 Fr][exp 32 B]
    uint8_t* pw = tail;
    uint8_t* inv_n = tail + 33 * 32;
    uint32_t* exp_dev = (uint32_t*)(inv_n + 32 * 32);
    B200_CUDA_CHECK(cudaMemcpyAsync(exp_dev, e, 32, cudaMemcpyHostToDevice, st));
    k_root_powers<<<1, 32, 0, st>>>(exp_dev, pw, scale, inv_n);
    k_roots_table<<<div_up(max_width_ + 1, 128), 128, 0, st>>>(pw, (uint8_t*)roots_, (uint8_t*)brp_roots_, max_width_, scale);
    B200_LAUNCH_CHECK();
    const int max_shm = (kTileElems + (1 << (kMaxLogM - 1))) * 32;
    B200_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_shm));
    B200_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_pass<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_shm));
    B200_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterMaxShm));
    B200_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));   // clusters of 16
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
}

FFTSettingsDev::~FFTSettingsDev() {
    cudaFree(roots_);
    cudaFree(brp_roots_);
    cudaFree(scratch_);
    cudaFree(scratch2_);
    cudaFree(scratch3_);
    cudaFree(g1_work_);
    cudaFree(g1_tmp_);
}

void FFTSettingsDev::ensure_scratch(size_t elems) {
    if (elems <= scratch_elems_) return;
    cudaFree(scratch_);
    cudaFree(scratch2_);
    scratch_ = dev_alloc<uint8_t>(elems * 32);
    scratch2_ = dev_alloc<uint8_t>(elems * 32);
    scratch_elems_ = elems;
}

// Kernel shape.  Default: radix-4 register groups (two stages per shared-memory round trip), 80 registers, three CTAs per
// SM -- measured on B200 (scripts/ntt_timing.py): 2^12 57 -> 24 us, 2^16 74 -> 31 us, 2^20 282 -> 254 us against the
// radix-8 groups at 128 registers / two CTAs per SM, because the transform is bound by the multiply pipe and needs the
// extra warps to keep it fed (ncu: FMA-heavy pipe 68 % busy, top stall math-pipe throttle).  B200_NTT_VARIANT=3 selects
// the radix-8 form.  Tile = points per CTA: small transforms use small tiles so that the pass fills more SMs.
static int ntt_variant() {
    static const int v = getenv("B200_NTT_VARIANT") ? atoi(getenv("B200_NTT_VARIANT")) : 2;
    return v;
}
static int tile_elems(size_t n) {
    static const int t = getenv("B200_NTT_TILE") ? atoi(getenv("B200_NTT_TILE")) : 0;
    if (t) return t;
    if (ntt_variant() == 3) return kTileElems;
    // measured (scripts/ntt_timing.py, B200_NTT_TILE sweep): up to 2^14 points 128-point tiles spread a pass over more SMs
    // (2^10 21.0 -> 16.9 us, 2^12 23.0 -> 20.9, 2^14 27.1 -> 23.0); 2^16 is best at 512 (31.2 vs 35.4), 2^18 and up at 1024
    if (n <= ((size_t)1 << 14)) return kTileElems / 16;
    return n <= ((size_t)1 << 16) ? kTileElems / 4 : kTileElems / 2;
}
static void launch_pass(const NttPass& p, int batch, cudaStream_t st) {
    int m = 1 << p.log_m;
    unsigned blocks = div_up(p.ncols, (size_t)p.G);
    const int variant = ntt_variant();
    const size_t shm = ((size_t)p.G * m + (m >> 1)) * 32;
    if (variant == 3) k_ntt_pass<3, 2><<<dim3(blocks, (unsigned)batch), kNttThreads, shm, st>>>(p);
    else k_ntt_pass<2, 3><<<dim3(blocks, (unsigned)batch), kNttThreads, shm, st>>>(p);
    B200_LAUNCH_CHECK();
}

// batch transforms of 2^k <= 2^22 points: one CTA per transform, or two passes (see ntt.cuh).  in_bstride / out_bstride:
// elements between consecutive transforms; output element i of a transform goes to out[i * out_mul] (out_mul > 1 when the
// transforms are the rows of a three-pass transform and their outputs interleave); tmp: batch * 2^k elements of scratch.
void FFTSettingsDev::transform(const void* in, void* out, int k, bool inverse, int batch, const uint8_t* scale_ptr, size_t twist_unit,
                               size_t in_bstride, size_t out_bstride, size_t out_mul, void* tmp, cudaStream_t st) {
    const size_t n = (size_t)1 << k;
    NttPass p{};
    p.roots = (const uint8_t*)roots_;
    p.nmax = max_width_;
    p.inverse = inverse;
    p.batch_stride = in_bstride;
    p.out_batch_stride = out_bstride;
    // one CTA per transform when there are many of them (or they are tiny); a lone mid-size transform is split into two
    // passes so that it spreads over several SMs (2^11 points: 50 -> ~25 us)
    if (k <= kMaxLogM && (k <= 8 || batch >= 16)) {
        p.in = (const uint8_t*)in; p.out = (uint8_t*)out;
        p.log_m = k; p.G = 1; p.ncols = 1;
        p.in_cstride = 0; p.in_rstride = 1; p.out_cstride = 0; p.out_rstride = out_mul;
        p.in_col_fast = 0; p.out_col_fast = 0;
        p.unit_m = max_width_ >> k;
        p.tw_unit = 0; p.twist_unit = twist_unit;
        p.scale = scale_ptr;
        launch_pass(p, batch, st);
        launches_ += 1;
        return;
    }
    int k2 = (k + 1) / 2, k1 = k - k2;  // n2 = 2^k2 (pass 1, strided columns), n1 = 2^k1 (pass 2, contiguous rows)
    size_t n1 = (size_t)1 << k1, n2 = (size_t)1 << k2;
    // B200_NTT_CLUSTER (per call: tests toggle it): 0 never, 8 / 16 that cluster size for every 2^9 .. 2^14, unset: by size
    const int cl_env = getenv("B200_NTT_CLUSTER") ? atoi(getenv("B200_NTT_CLUSTER")) : -1;
    const int max_auto = getenv("B200_NTT_CLUSTER_MAXLOG") ? atoi(getenv("B200_NTT_CLUSTER_MAXLOG")) : 12;
    int ctas = cl_env == 8 || cl_env == 16 ? cl_env : cl_env == 0 ? 0 : (k <= max_auto ? (k >= 11 ? 16 : 8) : 0);
    if (ctas == 16 && k1 < 4) ctas = 8;                        // at least one column and one row per CTA
    if (ctas && k >= kClusterMinLog && k <= kClusterMaxLog && ((size_t)2 * (n / ctas) + (n2 >> 1) + (n1 >> 1)) * 32 <= (size_t)kClusterMaxShm) {
        // one launch: a cluster of eight CTAs per transform, the inter-pass transpose through distributed shared memory
        NttPass p2 = p;
        p.in = (const uint8_t*)in;
        p.log_m = k2; p.ncols = n1;
        p.in_cstride = 1; p.in_rstride = n1;
        p.unit_m = max_width_ >> k2;
        p.tw_unit = max_width_ >> k; p.twist_unit = twist_unit;
        p2.out = (uint8_t*)out;
        p2.log_m = k1; p2.ncols = n2;
        p2.out_cstride = out_mul; p2.out_rstride = n2 * out_mul;
        p2.unit_m = max_width_ >> k1;
        p2.scale = scale_ptr;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)ctas, (unsigned)batch);
        lc.blockDim = dim3(kNttThreads);
        lc.dynamicSmemBytes = (2 * (n / ctas) + (n2 >> 1) + (n1 >> 1)) * 32;
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        B200_CUDA_CHECK(cudaLaunchKernelEx(&lc, k_ntt_cluster, p, p2));
        launches_ += 1;
        return;
    }
    // pass 1: for every column j1 < n1, transform the n2 points x[j1 + n1*j2]; times w_n^(i2*j1); in-place layout
    p.in = (const uint8_t*)in; p.out = (uint8_t*)tmp;
    p.out_batch_stride = n;
    p.log_m = k2; p.G = std::max(1, tile_elems(n) >> k2); p.ncols = n1;
    p.in_cstride = 1; p.in_rstride = n1; p.out_cstride = 1; p.out_rstride = n1;
    p.in_col_fast = 1; p.out_col_fast = 1;
    p.unit_m = max_width_ >> k2;
    p.tw_unit = max_width_ >> k; p.twist_unit = twist_unit;
    p.scale = nullptr;
    launch_pass(p, batch, st);
    // pass 2: for every i2 < n2, transform the n1 contiguous points y[n1*i2 + j1]; X[i2 + n2*i1]
    p.in = (const uint8_t*)tmp; p.out = (uint8_t*)out;
    p.batch_stride = n;
    p.out_batch_stride = out_bstride;
    p.log_m = k1; p.G = std::max(1, tile_elems(n) >> k1); p.ncols = n2;
    p.in_cstride = n1; p.in_rstride = 1; p.out_cstride = out_mul; p.out_rstride = n2 * out_mul;
    p.in_col_fast = 0; p.out_col_fast = 1;
    p.unit_m = max_width_ >> k1;
    p.tw_unit = 0; p.twist_unit = 0;
    p.scale = scale_ptr;
    launch_pass(p, batch, st);
    launches_ += 2;
}

void FFTSettingsDev::run_passes(const void* in, void* out, size_t n, bool inverse, int batch, bool scale,
                                size_t twist_unit, cudaStream_t st) {
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    const uint8_t* inv_n = (const uint8_t*)roots_ + (max_width_ + 1) * 32 + 33 * 32;
    const uint8_t* scale_ptr = scale ? inv_n + k * 32 : nullptr;
    if (k <= 2 * kMaxLogM) {
        if (k > kMaxLogM || !(k <= 8 || batch >= 16)) ensure_scratch((size_t)batch * n);
        transform(in, out, k, inverse, batch, scale_ptr, twist_unit, n, n, 1, scratch_, st);
        return;
    }
    // Above 2^22 points (the reference accepts scales up to 31, blst/src/types/fft_settings.rs:28-58): one more level of
    // the same decomposition, n = n1 * 2^11.  Pass A transforms the n1 strided columns of 2^11 points and applies the
    // twiddle w_n^(i2 j1); the 2^11 rows of n1 <= 2^20 contiguous points are then ordinary (two-pass) transforms whose
    // outputs interleave: X[i2 + 2^11 i1].  Three full sweeps over the data instead of two.
    const int kc = kMaxLogM, kr = k - kc;
    const size_t nc = (size_t)1 << kc, nr = (size_t)1 << kr;
    ensure_scratch((size_t)batch * n);
    if (scratch3_elems_ < n) {
        cudaFree(scratch3_);
        scratch3_ = nullptr; scratch3_elems_ = 0;
        scratch3_ = dev_alloc<uint8_t>(n * 32);
        scratch3_elems_ = n;
    }
    for (int b = 0; b < batch; b++) {
        const uint8_t* src = (const uint8_t*)in + (size_t)b * n * 32;
        uint8_t* dst = (uint8_t*)out + (size_t)b * n * 32;
        NttPass p{};
        p.roots = (const uint8_t*)roots_;
        p.nmax = max_width_;
        p.inverse = inverse;
        p.batch_stride = n; p.out_batch_stride = n;
        p.in = src; p.out = (uint8_t*)scratch3_;
        p.log_m = kc; p.G = std::max(1, tile_elems(n) >> kc); p.ncols = nr;
        p.in_cstride = 1; p.in_rstride = nr; p.out_cstride = 1; p.out_rstride = nr;
        p.in_col_fast = 1; p.out_col_fast = 1;
        p.unit_m = max_width_ >> kc;
        p.tw_unit = max_width_ >> k; p.twist_unit = twist_unit;
        p.scale = nullptr;
        launch_pass(p, 1, st);
        launches_ += 1;
        // rows: 2^11 transforms of nr points, input row i2 at scratch3 + i2 * nr, output element i1 at dst[i2 + 2^11 * i1]
        transform(scratch3_, dst, kr, inverse, (int)nc, scale_ptr, 0, nr, 1, nc, scratch_, st);
    }
}

void FFTSettingsDev::fft_fr(const void* in_dev, void* out_dev, size_t n, bool inverse, int batch, cudaStream_t st) {
    // argument checks of fft_fr_output (blst/src/fft_fr.rs:118-133)
    if (n > max_width_) throw CudaError(1, "Supplied list is longer than the available max width");
    if (n == 0 || (n & (n - 1))) throw CudaError(1, "A list with power-of-two length expected");
    launches_ = 0;
    run_passes(in_dev, out_dev, n, inverse, batch, inverse, 0, st);
}

void FFTSettingsDev::das_fft_extension(const void* evens_dev, void* odds_dev, size_t n, int batch, cudaStream_t st) {
    // argument checks of das_fft_extension (blst/src/data_availability_sampling.rs:79-87)
    if (n == 0) throw CudaError(1, "A non-zero list ab expected");
    if (n & (n - 1)) throw CudaError(1, "A list with power-of-two length expected");
    if (n * 2 > max_width_) throw CudaError(1, "Supplied list is longer than the available max width");
    launches_ = 0;
    // odds = NTT_n( INTT_n(evens) .* w_2n^j ): the unique degree < n extension evaluated on the odd coset
    ensure_scratch((size_t)batch * n);
    void* coeffs = scratch2_;
    run_passes(evens_dev, coeffs, n, true, batch, true, 0, st);
    run_passes(coeffs, odds_dev, n, false, batch, false, max_width_ / (2 * n), st);
}

}  // namespace b200
