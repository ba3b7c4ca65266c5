"""Host-side logic of the MSM engine that needs no GPU: the window plan (b200_msm_plan), and integer models of the three
identities the kernels rely on -- the signed-digit recoding with a narrower window 0 (k_digits), the segment fold
(k_segment_fold) and the digit-marginal reduce (k_marginals / k_group_finish).  Integers stand in for curve points: every
identity is linear in the bucket sums, so it holds in G1 iff it holds in Z."""
import os
import random

import pytest

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


@pytest.fixture(scope="module")
def B():
    import rust_kzg_b200 as B
    return B


def _env(**kw):
    class _E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return _E()


def test_plan_invariants(B):
    for lg in range(0, 27):
        for fixed in (True, False):
            p = B.msm_plan(1 << lg, fixed)
            assert 4 <= p["c"] <= (22 if fixed else 20)
            assert p["c0"] + (p["W"] - 1) * p["c"] >= 256, p       # the top window absorbs the last carry
            assert p["W"] == -(-256 // p["c"])
            # wide windows are folded down to the 15-bit reduce; 13-bit fixed-base windows take the measured 2-bit fold
            assert p["fold_bits"] == max(p["c"] - 16, 2 if (fixed and p["c"] == 13) else 0, 0)
            # scalars are < r < 2^255 and uniform on a randomised table: the top window must keep enough of its bits that its
            # buckets are not overloaded -- 2^(c - t) / W, its load relative to the average, stays below 4
            t = 255 - (p["W"] - 1) * p["c"]
            if t <= 0:                       # c divides 255 (15, 17): the last window only ever sees a carry that cannot occur,
                t += p["c"]                  # the window below it is the top one and is full
            assert 2.0 ** (p["c"] - t) / p["W"] < 4, p
    assert B.msm_plan(1 << 20)["c"] == 20 and B.msm_plan(1 << 20)["W"] == 13     # measured optimum (profiles/r01_window_sweeps.md)
    assert B.msm_plan(1 << 21)["c"] == 20
    assert B.msm_plan(1 << 19)["c"] == 16 and B.msm_plan(4096)["c"] == 13
    assert B.msm_plan(1 << 20, False)["c"] == 16                                  # one bucket set per window: no wide windows


def test_plan_env_overrides_are_clamped(B):
    with _env(B200_MSM_C=40, B200_MSM_VC=40):
        assert B.msm_plan(4096)["c"] == 22 and B.msm_plan(4096, False)["c"] == 20
    with _env(B200_MSM_C=1):
        assert B.msm_plan(4096)["c"] == 4
    with _env(B200_MSM_C=13, B200_MSM_C0=9):
        p = B.msm_plan(4096)
        assert (p["c"], p["c0"], p["W"]) == (13, 9, 20) and p["c0"] + 19 * 13 == 256
        assert B.msm_plan(4096, False)["c0"] == B.msm_plan(4096, False)["c"]      # variable base: uniform windows only


def recode(s, c, c0, W):
    """k_digits (csrc/msm.cu): window 0 has c0 bits, the others c bits; raw > half becomes raw - 2^cw with a carry"""
    digits, carry = [], 0
    for j in range(W):
        o, cw = (0, c0) if j == 0 else (c0 + c * (j - 1), c)
        raw = ((s >> o) & ((1 << cw) - 1) if o < 256 else 0) + carry
        neg = raw > (1 << (cw - 1))
        digits.append(raw - (1 << cw) if neg else raw)
        carry = int(neg)
    return digits, carry


@pytest.mark.parametrize("c,c0", [(16, 16), (20, 20), (13, 13), (13, 9), (15, 1), (12, 4), (20, 16), (22, 22), (8, 8)])
def test_signed_digits_reconstruct_the_scalar(c, c0):
    W = -(-256 // c)
    assert c0 + (W - 1) * c >= 256
    rnd = random.Random(c * 100 + c0)
    scalars = [0, 1, R_MOD - 1, R_MOD >> 1, (1 << 248) - 1, 1 << 247, (1 << (c0 - 1)) if c0 > 1 else 1, (1 << 255) % R_MOD]
    scalars += [rnd.randrange(R_MOD) for _ in range(300)] + [rnd.randrange(1 << 248) for _ in range(100)]
    nb = 1 << (c - 1)
    for s in scalars:
        d, carry = recode(s, c, c0, W)
        assert carry == 0, "a final carry would be lost"                      # needs s < r < 0.91 * 2^255
        assert all(abs(x) <= nb for x in d)                                      # bucket |d| - 1 < nb
        assert abs(d[0]) <= 1 << max(c0 - 1, 0)
        offs = [0] + [c0 + c * (j - 1) for j in range(1, W)]
        assert sum(x << o for x, o in zip(d, offs)) == s                         # table row j holds 2^offs[j] * P


@pytest.mark.parametrize("bits,kf", [(19, 4), (16, 1), (12, 2), (21, 6), (11, 3)])
def test_segment_fold_identity(bits, kf):
    """sum_b (b+1) B_b = 2^kf * sum_hi (hi+1) T_hi - sum_hi R_hi with the kernel's recurrence acc += run; run += B"""
    rnd = random.Random(bits * 10 + kf)
    nb = 1 << min(bits, 14)            # the identity does not depend on the size; keep the model fast
    Bk = [rnd.randrange(-10 ** 6, 10 ** 6) if rnd.random() < 0.7 else 0 for _ in range(nb)]   # 0 = empty bucket
    want = sum((b + 1) * Bk[b] for b in range(nb))
    T, Rs = [], []
    for hi in range(nb >> kf):
        run = acc = 0
        for lo in range(1 << kf):
            acc += run
            run += Bk[(hi << kf) + lo]
        T.append(run)
        Rs.append(acc)
    assert (sum((hi + 1) * t for hi, t in enumerate(T)) << kf) - sum(Rs) == want


def axis_plan(bits):
    """MsmEngine::run: D = ceil(bits / 5) digits of near-equal width, least significant first"""
    D = max(1, (bits + 4) // 5)
    w, sh, off = [], [], 0
    for a in range(D):
        wa = (bits - off + (D - a) - 1) // (D - a)
        w.append(wa)
        sh.append(off)
        off += wa
    return D, w, sh


@pytest.mark.parametrize("bits", [1, 4, 5, 7, 9, 11, 15])
def test_digit_marginal_reduce_identity(bits):
    """sum_b (b+1) B_b = sum_b B_b + sum_a 2^sh_a sum_v v M[a][v],  M[a][v] = sum of the buckets whose digit a is v"""
    D, w, sh = axis_plan(bits)
    assert sum(w) == bits and D <= 3 and max(w) <= 5
    rnd = random.Random(bits)
    nb = 1 << bits
    Bk = [rnd.randrange(-1000, 1000) for _ in range(nb)]
    total = sum(Bk)
    acc = total
    for a in range(D):
        M = [0] * (1 << w[a])
        for b in range(nb):
            M[(b >> sh[a]) & ((1 << w[a]) - 1)] += Bk[b]
        # warp suffix scan form: sum_v v M_v = sum_{k>=1} Suf_k
        suf = [sum(M[k:]) for k in range(len(M))]
        assert sum(suf[1:]) == sum(v * m for v, m in enumerate(M))
        acc += sum(suf[1:]) << sh[a]
    assert acc == sum((b + 1) * Bk[b] for b in range(nb))
