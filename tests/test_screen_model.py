"""Host model of the tcgen05 epilogue's single-precision screen (ldx_triangle_mma.cu: fast_pair2, true_n11), run
against the pinned oracle.  The CUDA kernel itself is checked bit for bit on the GPU (tests/test_parity_gpu.py); this
file checks the MATHEMATICS the kernel relies on, on the CPU, so that a wrong bound or a wrong allele-flip identity is
caught without a GPU:

  * minor-allele encoding: r2 and |D'| from (n11', a', c') with a' = min(n1a, N - n1a) equal those of the true counts,
    with  m = Dn > 0 ? N * min(a', c') - a'c' : a'c'  as the D' bound (calc_ld.py:63-76) and true_n11() as the inverse;
  * every pair the screen ACCEPTS carries the reference's rounded values (the oracle's packed word), for reciprocals
    perturbed by the full 2^-23 that rcp.approx.ftz.f32 is allowed;
  * the share of deferred pairs stays small.

The float32 FMA is modelled as round32(a * b + c) in float64 (the products here are exact in float64).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ld_oracle  # noqa: E402

F = np.float32
U = 2.0 ** -24
C_DP, C_R2 = F(8 * U), F(12 * U)                 # SCREEN_C_DP / SCREEN_C_R2
MAGIC = F(12582912.0)
R2_INT0, DP_INT0 = 0x8000, 0x80000000


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def rcp32(x, skew):
    """rcp.approx.ftz.f32 within its documented 2^-23 relative error; `skew` in [-1, 1] picks where in that band."""
    with np.errstate(divide="ignore"):
        r = (1.0 / x.astype(np.float64)) * (1.0 + skew * 2.0 ** -23)
    return r.astype(F)


def true_n11(n11p, a_t, b_t, n):
    fa, fb = 2 * a_t > n, 2 * b_t > n
    return np.where(fa, np.where(fb, n11p - n + a_t + b_t, b_t - n11p), np.where(fb, a_t - n11p, n11p))


def minor_n11(n11, a_t, b_t, n):
    """What the tensor cores count when rows with 2 * n1 > N enter complemented."""
    fa, fb = 2 * a_t > n, 2 * b_t > n
    return np.where(fa, np.where(fb, n - a_t - b_t + n11, b_t - n11), np.where(fb, a_t - n11, n11))


def screen(n, n11p, a_t, b_t, skew):
    """-> (packed word, deferred flag) as fast_pair2 computes them."""
    nf = F(n)
    a = np.minimum(a_t, n - a_t)
    c = np.minimum(b_t, n - b_t)
    af, cf = a.astype(F), c.astype(F)
    lim_g = 1.0e4 * 16.0 * 1.1102230246251565e-16 * n * n
    lim_r2, lim_dp = F(0.5 - (lim_g + 1.0e-5)), F(0.5 - (0.5 * lim_g + 1.0e-5))
    ra = rcp32((a * (n - a)).astype(F), skew)
    rc = rcp32((c * (n - c)).astype(F), -skew)
    acc = (n11p * 128).astype(F)
    nP = (-af) * cf
    fD = fma32(acc, np.full_like(acc, nf * F(0.0078125)), nP)
    mp = fma32(np.full_like(acc, nf), np.minimum(af, cf), nP)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        r1 = rcp32(np.where(fD > 0, mp, nP), skew)
        rinv = ra * rc
        d4 = fD * F(1.0e4)
        x_dp = d4 * r1
        x_r2 = (d4 * fD) * rinv
        t_dp, t_r2 = x_dp + MAGIC, x_r2 + MAGIC
        f_dp = fma32(t_dp + (-MAGIC), np.full_like(acc, F(-1)), x_dp)
        f_r2 = fma32(t_r2 + (-MAGIC), np.full_like(acc, F(-1)), x_r2)
        l_dp = fma32(x_dp, np.full_like(acc, -C_DP), np.full_like(acc, lim_dp))
        l_r2 = fma32(x_r2, np.full_like(acc, -C_R2), np.full_like(acc, lim_r2))
        mono = np.isinf(rinv)
        fine = (np.abs(f_dp) <= l_dp) & (np.abs(f_r2) <= l_r2) & (fD != 0)
    slow = ~(fine | mono)
    k_dp = t_dp.view(np.uint32) & 0xFFFF
    k_r2 = t_r2.view(np.uint32) & 0xFFFF
    word = np.where(mono, np.uint32(DP_INT0 | R2_INT0), (k_dp << 16) | k_r2).astype(np.uint32)
    return word, slow


def random_counts(rng, n, size):
    """Count triples with every regime represented: common x common, rare x rare, near-perfect LD, monomorphic."""
    a_t = np.where(rng.random(size) < 0.5, rng.integers(0, n + 1, size), np.minimum(rng.geometric(0.01, size), n))
    b_t = np.where(rng.random(size) < 0.5, rng.integers(0, n + 1, size), np.minimum(rng.geometric(0.01, size), n))
    flip = rng.random(size) < 0.3
    a_t = np.where(flip, n - a_t, a_t)
    lo, hi = np.maximum(0, a_t + b_t - n), np.minimum(a_t, b_t)
    u = rng.random(size)
    shape = rng.integers(0, 4, size)
    n11 = np.where(shape == 0, lo, np.where(shape == 1, hi, lo + np.floor(u * (hi - lo + 1)).astype(np.int64)))
    near = shape == 3                                         # around independence: Dn near 0
    n11 = np.where(near, np.clip(np.round(a_t * b_t / n).astype(np.int64) + rng.integers(-1, 2, size), lo, hi), n11)
    return np.clip(n11, lo, hi).astype(np.int64), a_t.astype(np.int64), b_t.astype(np.int64)


@pytest.mark.parametrize("n", [2, 7, 198, 1006, 5008, 5792, 8192])
def test_minor_allele_identities(n):
    rng = np.random.default_rng(n)
    n11, a_t, b_t = random_counts(rng, n, 20000)
    n11p = minor_n11(n11, a_t, b_t, n)
    a, c = np.minimum(a_t, n - a_t), np.minimum(b_t, n - b_t)
    assert (n11p >= 0).all() and (n11p <= np.minimum(a, c)).all()
    assert (true_n11(n11p, a_t, b_t, n) == n11).all()
    # |Dn| is invariant, and the D' bound of calc_ld.py:63-76 on the true counts equals the minor-allele form
    dn_t = n11 * n - a_t * b_t
    dn_m = n11p * n - a * c
    assert (np.abs(dn_t) == np.abs(dn_m)).all()
    m_t = np.where(dn_t >= 0, np.minimum(a_t * (n - b_t), (n - a_t) * b_t), np.minimum(a_t * b_t, (n - a_t) * (n - b_t)))
    m_m = np.where(dn_m > 0, n * np.minimum(a, c) - a * c, a * c)
    nz = dn_t != 0
    assert (m_t[nz] == m_m[nz]).all()
    assert ((a_t * (n - a_t)) * (b_t * (n - b_t)) == (a * (n - a)) * (c * (n - c))).all()


@pytest.mark.parametrize("n", [2, 7, 198, 1006, 5008, 5792, 8192])
@pytest.mark.parametrize("skew", [-1.0, 0.0, 1.0])
def test_accepted_pairs_carry_the_reference_rounding(n, skew):
    rng = np.random.default_rng(1000 + n)
    n11, a_t, b_t = random_counts(rng, n, 120000)
    n11p = minor_n11(n11, a_t, b_t, n)
    word, slow = screen(n, n11p, a_t, b_t, skew)
    want = ld_oracle.packed_words(n, n11.astype(np.int32), a_t.astype(np.int32), b_t.astype(np.int32))
    ok = ~slow
    bad = np.flatnonzero(ok & (word != want))
    assert bad.size == 0, [(int(n11[i]), int(a_t[i]), int(b_t[i]), hex(word[i]), hex(want[i])) for i in bad[:5]]
    # nothing but rounding boundaries, exact zeros of D between polymorphic variants and N beyond the guard band's reach is deferred
    if n <= 5792:
        poly = (a_t > 0) & (a_t < n) & (b_t > 0) & (b_t < n) & (n11 * n != a_t * b_t)
        assert slow[poly].mean() < 0.03, slow[poly].mean()
