"""CPU oracle (pure-Python big integers) for the rust-kzg hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* of the reference's algorithm for the MSM / NTT / EIP-4844 commitment+proof path,
used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker.  The product
(rust-kzg_b200/) never imports it.  The arithmetic of the parity target lives in the un-vendored crate
blst 0.3.16 (Cargo.lock:530-532); what is restated here is the published BLS12-381 arithmetic plus the reference's
own call sites.  Parity is PINNED: tests/test_oracle_golden.py checks this file against the reference's
consensus-spec vectors and KATs (tests/golden/, extracted by tests/golden/make_golden.py).

Each function cites the reference file:line (relative to /root/reference) it follows.
Big-int arithmetic is slow: use for small cases; oracle/kzg_oracle.c is the fast restatement.
"""
import hashlib

# ---------------------------------------------------------------------------------------------------------------
# constants (zkcrypto/bls12_381/src/fp.rs:71, scalar.rs:77, g1.rs:176; blst/src/consts.rs:52-116)
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
B = 4
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
FIELD_ELEMENTS_PER_BLOB = 4096          # kzg/src/eip_4844.rs:32
FIELD_ELEMENTS_PER_EXT_BLOB = 8192      # kzg/src/eth/mod.rs
FIELD_ELEMENTS_PER_CELL = 64
CELLS_PER_EXT_BLOB = 128
BYTES_PER_BLOB = 131072
FIAT_SHAMIR_PROTOCOL_DOMAIN = b"FSBLOBVERIFY_V1_"   # kzg/src/eip_4844.rs:52-54
PRIMITIVE_ROOT = 7                      # SCALE2_ROOT_OF_UNITY[i] = 7^((r-1)/2^i)  (blst/src/consts.rs:14-50, checked in tests)


class OracleError(ValueError):
    """Mirrors the reference's Err(String) results."""


# ---------------------------------------------------------------------------------------------------------------
# Fr: wire format (blst/src/types/fr.rs:64-107, 127-136)
def fr_from_bytes(b: bytes) -> int:
    if len(b) != 32:
        raise OracleError("Invalid byte length")
    v = int.from_bytes(b, "big")
    if v >= R:                           # blst_scalar_fr_check
        raise OracleError("Invalid scalar")
    return v


def fr_from_bytes_unchecked(b: bytes) -> int:
    return int.from_bytes(b, "big") % R  # blst_fr_from_scalar reduces (fr.rs:88-107)


def fr_to_bytes(v: int) -> bytes:
    return int(v % R).to_bytes(32, "big")


def fr_inv(a: int) -> int:
    return pow(a, R - 2, R)


def fr_to_u64_arr(v: int):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def fr_from_u64_arr(a) -> int:
    return sum(int(x) << (64 * i) for i, x in enumerate(a)) % R


# ---------------------------------------------------------------------------------------------------------------
# G1, Jacobian coordinates over Python ints.  INF is Z == 0 (blst/src/types/g1.rs:151-177).
INF = (0, 0, 0)


def g1_is_inf(p):
    return p[2] == 0


def g1_from_affine(a):
    return INF if a is None else (a[0], a[1], 1)


def g1_to_affine(p):
    if p[2] == 0:
        return None
    zi = pow(p[2], P - 2, P)
    zi2 = zi * zi % P
    return (p[0] * zi2 % P, p[1] * zi2 * zi % P)


def g1_dbl(p):
    X, Y, Z = p
    if Z == 0 or Y == 0:
        return INF
    A = X * X % P
    Bq = Y * Y % P
    C = Bq * Bq % P
    D = 2 * ((X + Bq) * (X + Bq) - A - C) % P
    E = 3 * A % P
    F = E * E % P
    X3 = (F - 2 * D) % P
    Y3 = (E * (D - X3) - 8 * C) % P
    Z3 = 2 * Y * Z % P
    return (X3, Y3, Z3)


def g1_add(p, q):
    """add_or_double semantics (blst_p1_add_or_double, blst/src/types/g1.rs:102-108)."""
    if p[2] == 0:
        return q
    if q[2] == 0:
        return p
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    Z1Z1 = Z1 * Z1 % P
    Z2Z2 = Z2 * Z2 % P
    U1 = X1 * Z2Z2 % P
    U2 = X2 * Z1Z1 % P
    S1 = Y1 * Z2 * Z2Z2 % P
    S2 = Y2 * Z1 * Z1Z1 % P
    if U1 == U2:
        return g1_dbl(p) if S1 == S2 else INF
    H = (U2 - U1) % P
    I = 4 * H * H % P
    J = H * I % P
    r = 2 * (S2 - S1) % P
    V = U1 * I % P
    X3 = (r * r - J - 2 * V) % P
    Y3 = (r * (V - X3) - 2 * S1 * J) % P
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % P
    return (X3, Y3, Z3)


def g1_neg(p):
    return (p[0], (-p[1]) % P, p[2])


def g1_mul(p, k: int):
    """blst_p1_mult (blst/src/types/g1.rs:242-273): plain double-and-add is the same group element."""
    k %= R
    acc = INF
    for bit in bin(k)[2:] if k else "":
        acc = g1_dbl(acc)
        if bit == "1":
            acc = g1_add(acc, p)
    return acc


def g1_eq(p, q):
    return g1_to_affine(p) == g1_to_affine(q)


def g1_on_curve_affine(a):
    x, y = a
    return (y * y - x * x * x - B) % P == 0


def g1_in_subgroup(p):
    """blst_p1_in_g1 (g1.rs:110-119): membership in the order-r subgroup."""
    if p[2] == 0:
        return True
    acc = INF                     # [r]P with r not reduced
    for bit in bin(R)[2:]:
        acc = g1_dbl(acc)
        if bit == "1":
            acc = g1_add(acc, p)
    return acc[2] == 0


# compressed encoding (zkcrypto/bls12_381/src/notes/serialization.rs:1-29, g1.rs:221, 337-390;
# blst_p1_compress / blst_p1_uncompress via blst/src/types/g1.rs:65-100)
def g1_compress(p) -> bytes:
    a = g1_to_affine(p)
    if a is None:
        return bytes([0xC0]) + bytes(47)
    x, y = a
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > (P - 1) // 2:          # lexicographically largest
        out[0] |= 0x20
    return bytes(out)


def g1_uncompress(b: bytes):
    """-> Jacobian point; raises on malformed input.  No subgroup check (from_bytes does not do one)."""
    if len(b) != 48:
        raise OracleError("Invalid byte length")
    c_flag, i_flag, s_flag = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c_flag:
        raise OracleError("Failed to uncompress")
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if i_flag:
        if s_flag or x != 0:
            raise OracleError("Failed to uncompress")
        return INF
    if x >= P:
        raise OracleError("Failed to uncompress")
    y2 = (x * x * x + B) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        raise OracleError("Failed to uncompress")
    if (y > (P - 1) // 2) != bool(s_flag):
        y = P - y
    return (x, y, 1)


# ---------------------------------------------------------------------------------------------------------------
# bit-reversal (kzg/src/common_utils.rs:6-34)
def reverse_bit_order(vals):
    n = len(vals)
    if n == 0:
        raise OracleError("Values can not be empty")
    if n == 1:
        return list(vals)
    if n & (n - 1):
        raise OracleError("Values length has to be a power of 2")
    bits = n.bit_length() - 1
    out = list(vals)
    for i in range(n):
        r = int(format(i, "0%db" % bits)[::-1], 2)
        if r > i:
            out[i], out[r] = out[r], out[i]
    return out


# ---------------------------------------------------------------------------------------------------------------
# FFTSettings (blst/src/types/fft_settings.rs:28-58, 90-106)
class FFTSettings:
    def __init__(self, scale: int):
        if scale >= 32:
            raise OracleError("Scale is expected to be within root of unity matrix row size")
        self.max_width = 1 << scale
        self.root_of_unity = scale2_root_of_unity(scale)
        roots = [1]
        for _ in range(self.max_width):
            roots.append(roots[-1] * self.root_of_unity % R)
        assert roots[-1] == 1 and (self.max_width == 1 or roots[self.max_width // 2] == R - 1)
        self.roots_of_unity = roots                                   # max_width + 1 entries
        self.brp_roots_of_unity = reverse_bit_order(roots[:-1])
        self.reverse_roots_of_unity = roots[::-1]

    # blst/src/fft_fr.rs:112-165
    def fft_fr(self, data, inverse=False):
        n = len(data)
        if n > self.max_width:
            raise OracleError("Supplied list is longer than the available max width")
        if n == 0 or n & (n - 1):
            raise OracleError("A list with power-of-two length expected")
        stride = self.max_width // n
        roots = self.reverse_roots_of_unity if inverse else self.roots_of_unity
        out = _fft_fr_fast(list(data), roots, stride)
        if inverse:
            inv = fr_inv(n % R)
            out = [v * inv % R for v in out]
        return out

    # blst/src/data_availability_sampling.rs:14-100
    def das_fft_extension(self, evens):
        n = len(evens)
        if n == 0:
            raise OracleError("A non-zero list ab expected")
        if n & (n - 1):
            raise OracleError("A list with power-of-two length expected")
        if n * 2 > self.max_width:
            raise OracleError("Supplied list is longer than the available max width")
        stride = self.max_width // (n * 2)
        odds = list(evens)
        self._das_stride(odds, 0, n, stride)
        inv = fr_inv(n % R)
        return [v * inv % R for v in odds]

    def _das_stride(self, a, lo, n, stride):
        ru, rr = self.roots_of_unity, self.reverse_roots_of_unity
        if n < 2:
            return
        if n == 2:
            x = (a[lo] + a[lo + 1]) % R
            y = (a[lo] - a[lo + 1]) % R
            yr = y * ru[stride] % R
            a[lo], a[lo + 1] = (x + yr) % R, (x - yr) % R
            return
        half = n // 2
        for i in range(half):
            t1 = (a[lo + i] + a[lo + half + i]) % R
            t2 = (a[lo + i] - a[lo + half + i]) % R
            a[lo + half + i] = t2 * rr[i * 2 * stride] % R
            a[lo + i] = t1
        self._das_stride(a, lo, half, stride * 2)
        self._das_stride(a, lo + half, half, stride * 2)
        for i in range(half):
            x, y = a[lo + i], a[lo + half + i]
            yr = y * ru[(1 + 2 * i) * stride] % R
            a[lo + i], a[lo + half + i] = (x + yr) % R, (x - yr) % R


def scale2_root_of_unity(scale: int) -> int:
    return pow(PRIMITIVE_ROOT, (R - 1) >> scale, R)


def _fft_fr_fast(data, roots, roots_stride):
    """Radix-2 DIT, natural in / natural out (blst/src/fft_fr.rs:49-108), iterative form of the same recursion."""
    n = len(data)
    if n == 1:
        return data
    even = _fft_fr_fast(data[0::2], roots, roots_stride * 2)
    odd = _fft_fr_fast(data[1::2], roots, roots_stride * 2)
    half = n // 2
    out = [0] * n
    for i in range(half):
        yr = odd[i] * roots[i * roots_stride] % R
        out[i] = (even[i] + yr) % R
        out[i + half] = (even[i] - yr) % R
    return out


def fft_fr_slow(data, roots, roots_stride):
    """O(n^2) DFT (blst/src/fft_fr.rs:168-186)."""
    n = len(data)
    return [sum(data[j] * roots[((i * j) % n) * roots_stride] for j in range(n)) % R for i in range(n)]


# ---------------------------------------------------------------------------------------------------------------
# MSM (kzg/src/msm/msm_impls.rs:114-148; tiling_pippenger_ops.rs; pippenger_utils.rs)
def pippenger_window_size(npoints: int) -> int:      # pippenger_utils.rs:300-317
    wbits = npoints.bit_length()
    if wbits > 13:
        return wbits - 4
    if wbits > 5:
        return wbits - 3
    return 2


def booth_encode(wval: int, sz: int) -> int:         # pippenger_utils.rs:251-256 (u64 arithmetic)
    M = (1 << 64) - 1
    mask = (0 - (wval >> sz)) & M
    wval = (wval + 1) >> 1
    return ((wval ^ mask) - mask) & M


def get_wval(scalar: int, off: int, bits: int) -> int:
    """get_wval_limb (pippenger_utils.rs:231-244) returns the bits with trash above; callers mask."""
    return scalar >> off


def msm_naive(points, scalars):
    """len < 8 path of msm() (msm_impls.rs:126-133) and the tests' expected value (bls12_381.rs:238-243)."""
    acc = INF
    for p, s in zip(points, scalars):
        acc = g1_add(acc, g1_mul(p, s))
    return acc


def msm_pippenger(points, scalars):
    """pippenger() (msm_impls.rs:40-61) -> tiling_pippenger (tiling_pippenger_ops.rs:106-138): signed-digit
    (Booth) windows processed MSB first; buckets integrated by the running-sum of p1_integrate_buckets (:21-45)."""
    pts, scs = [], []
    for p, s in zip(points, scalars):
        if not g1_is_inf(p):                         # filter infinity (msm_impls.rs:51-56)
            a = g1_to_affine(p)
            pts.append((a[0], a[1], 1))
            scs.append(s % R)
    n = len(pts)
    if n == 0:
        return INF
    window = pippenger_window_size(n)
    nbuckets = 1 << (window - 1)

    def tile(bit0, wbits, cbits):
        buckets = [INF] * nbuckets
        wmask = (1 << (wbits + 1)) - 1
        z = 1 if bit0 == 0 else 0
        b0 = bit0 - (z ^ 1)
        wb = wbits + (z ^ 1)
        for p, s in zip(pts, scs):
            wval = ((get_wval(s, b0, wb) << z) & wmask)
            wval = booth_encode(wval, cbits)
            sign = (wval >> cbits) & 1
            idx = wval & ((1 << cbits) - 1)
            if idx:
                buckets[idx - 1] = g1_add(buckets[idx - 1], g1_neg(p) if sign else p)
        # p1_integrate_buckets with wbits = cbits-1
        m = (1 << (cbits - 1)) - 1
        ret = buckets[m]
        acc = buckets[m]
        while m:
            m -= 1
            acc = g1_add(acc, buckets[m])
            ret = g1_add(ret, acc)
        return ret

    wbits = 255 % window
    cbits = wbits + 1
    bit0 = 255
    ret = INF
    while True:
        bit0 -= wbits
        if bit0 == 0:
            break
        ret = g1_add(ret, tile(bit0, wbits, cbits))
        for _ in range(window):
            ret = g1_dbl(ret)
        cbits = window
        wbits = window
    ret = g1_add(ret, tile(0, wbits, cbits))
    return ret


def g1_lincomb(points, scalars, length=None):
    """G1LinComb::g1_lincomb (kzg/src/lib.rs:142-181) -> msm() (msm_impls.rs:114-148)."""
    if length is None:
        length = min(len(points), len(scalars))
    points, scalars = points[:length], scalars[:length]
    if length < 8:
        return msm_naive(points, scalars)
    return msm_pippenger(points, scalars)


# ---------------------------------------------------------------------------------------------------------------
# trusted setup (kzg/src/eip_4844.rs:151-228, 1022-1086)
class KZGSettings:
    pass


def load_trusted_setup_string(text: str):
    tok = text.split()
    n1, n2 = int(tok[0]), int(tok[1])
    if n1 != FIELD_ELEMENTS_PER_BLOB or n2 != 65:
        raise OracleError("Incorrect trusted setup format")
    body = tok[2:]
    if len(body) != 2 * n1 + n2:
        raise OracleError("Incorrect trusted setup format")
    g1_lagrange = [bytes.fromhex(t) for t in body[:n1]]
    g2_monomial = [bytes.fromhex(t) for t in body[n1:n1 + n2]]
    g1_monomial = [bytes.fromhex(t) for t in body[n1 + n2:]]
    return g1_monomial, g1_lagrange, g2_monomial


def load_trusted_setup(text: str, need_monomial=False) -> KZGSettings:
    g1_monomial_b, g1_lagrange_b, g2_b = load_trusted_setup_string(text)
    s = KZGSettings()
    s.g1_lagrange_brp = reverse_bit_order([g1_uncompress(b) for b in g1_lagrange_b])   # eip_4844.rs:1055-1070
    s.g1_monomial = [g1_uncompress(b) for b in g1_monomial_b] if need_monomial else None
    s.g2_monomial_bytes = g2_b
    s.fs = FFTSettings(13)                                                              # eip_4844.rs:1072-1077
    return s


# ---------------------------------------------------------------------------------------------------------------
# EIP-4844 (kzg/src/eip_4844.rs)
def bytes_to_blob(blob: bytes):                       # :867-880
    if len(blob) != BYTES_PER_BLOB:
        raise OracleError("Invalid blob: Invalid byte length")
    return [fr_from_bytes(blob[i:i + 32]) for i in range(0, BYTES_PER_BLOB, 32)]


def blob_to_kzg_commitment(blob: bytes, s: KZGSettings) -> bytes:     # :278-314
    poly = bytes_to_blob(blob)
    return g1_compress(g1_lincomb(s.g1_lagrange_brp, poly, FIELD_ELEMENTS_PER_BLOB))


def compute_challenge(poly, commitment_bytes: bytes) -> int:           # :920-945
    h = hashlib.sha256()
    h.update(FIAT_SHAMIR_PROTOCOL_DOMAIN)
    h.update((0).to_bytes(8, "big"))
    h.update(FIELD_ELEMENTS_PER_BLOB.to_bytes(8, "big"))
    for v in poly:
        h.update(fr_to_bytes(v))
    h.update(commitment_bytes)
    return fr_from_bytes_unchecked(h.digest())


def fr_batch_inv(a):                                  # :882-914
    if not a:
        raise OracleError("Length is less than 0.")
    acc = 1
    out = []
    for v in a:
        out.append(acc)
        acc = acc * v % R
    if acc == 0:
        raise OracleError("Zero input")
    acc = fr_inv(acc)
    for i in range(len(a) - 1, -1, -1):
        out[i] = out[i] * acc % R
        acc = acc * a[i] % R
    return out


def evaluate_polynomial_in_evaluation_form(poly, x: int, s: KZGSettings) -> int:   # :954-1003
    roots = s.fs.brp_roots_of_unity
    n = FIELD_ELEMENTS_PER_BLOB
    inv_in = []
    for i in range(n):
        if x == roots[i]:
            return poly[i]
        inv_in.append((x - roots[i]) % R)
    inv = fr_batch_inv(inv_in)
    out = 0
    for i in range(n):
        out = (out + inv[i] * roots[i] % R * poly[i]) % R
    out = out * fr_inv(n) % R
    out = out * ((pow(x, n, R) - 1) % R) % R
    return out


def compute_kzg_proof_fr(poly, z: int, s: KZGSettings):                 # :437-519 -> (proof point, y)
    n = FIELD_ELEMENTS_PER_BLOB
    roots = s.fs.brp_roots_of_unity
    y = evaluate_polynomial_in_evaluation_form(poly, z, s)
    q = [0] * n
    inv_in = [0] * n
    m = 0
    for i in range(n):
        if z == roots[i]:
            m = i + 1
            inv_in[i] = 1
            continue
        q[i] = (poly[i] - y) % R
        inv_in[i] = (roots[i] - z) % R
    inv = fr_batch_inv(inv_in)
    q = [q[i] * inv[i] % R for i in range(n)]
    if m:
        m -= 1
        q[m] = 0
        for i in range(n):
            if i != m:
                inv_in[i] = (z - roots[i]) % R * z % R
        inv = fr_batch_inv(inv_in)
        for i in range(n):
            if i != m:
                t = (poly[i] - y) % R * roots[i] % R * inv[i] % R
                q[m] = (q[m] + t) % R
    return g1_lincomb(s.g1_lagrange_brp, q, n), y, q


def compute_kzg_proof(blob: bytes, z_bytes: bytes, s: KZGSettings):    # compute_kzg_proof_raw :521-539
    poly = bytes_to_blob(blob)
    z = fr_from_bytes(z_bytes)
    proof, y, _ = compute_kzg_proof_fr(poly, z, s)
    return g1_compress(proof), fr_to_bytes(y)


def compute_blob_kzg_proof(blob: bytes, commitment_bytes: bytes, s: KZGSettings) -> bytes:   # :541-584
    poly = bytes_to_blob(blob)
    c = g1_uncompress(commitment_bytes)
    if not g1_is_inf(c) and not g1_in_subgroup(c):
        raise OracleError("Invalid commitment")
    z = compute_challenge(poly, g1_compress(c))
    proof, _, _ = compute_kzg_proof_fr(poly, z, s)
    return g1_compress(proof)


# EIP-7594 compute_cells (kzg/src/das.rs:244-275, 618-629): the NTT golden vectors
def compute_cells(blob: bytes, s: KZGSettings):
    poly = bytes_to_blob(blob)
    fs = s.fs
    monomial = fs.fft_fr(reverse_bit_order(poly), True)                                 # poly_lagrange_to_monomial
    ext = fs.fft_fr(monomial + [0] * 4096, False)
    ext = reverse_bit_order(ext)
    return [b"".join(fr_to_bytes(v) for v in ext[c * 64:(c + 1) * 64]) for c in range(CELLS_PER_EXT_BLOB)]

