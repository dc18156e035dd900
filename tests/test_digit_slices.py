"""The arithmetic of the INT8 acquisition product (csrc/acquire_i8.cu), restated with numpy integers on the host:
digit extraction, int32 level sums, Horner combination.  No GPU involved -- this pins the *scheme* (the claims of
DESIGN.md 3a: level sums fit int32, levels 0..7 reproduce the fixed-point product to < 2^-49 of the row scale, the
result is closer to the exact product than a float64 dot product of the same data)."""
from fractions import Fraction

import numpy as np
import pytest

SLICES = 7


def w_digits(w_row):
    """row_exp_kernel + slice_w_kernel: per-row exponent, 55-bit fixed point, balanced base-256 digits; a[p], p = 0 most significant."""
    m = np.abs(w_row).max()
    e = (int(np.floor(np.log2(m))) if m > 0 else 0) + 2
    fx = np.rint(np.ldexp(w_row, 55 - e)).astype(np.int64)
    exact_fx = fx.copy()
    a = np.zeros((SLICES, len(w_row)), dtype=np.int64)
    for mth in range(SLICES):
        d = ((fx & 0xFF) ^ 0x80) - 0x80                     # (signed char)(fx & 0xFF)
        a[SLICES - 1 - mth] = d
        fx = (fx - d) >> 8
    assert not fx.any()
    return e, a, exact_fx


def k_digits(k):
    """panel_i8_kernel, unsigned variant: fx = round(k * 2^54); its 7 low bytes are the digits; b[q], q = 0 most significant."""
    fx = np.rint(np.ldexp(k, 54)).astype(np.int64)
    b = np.stack([(fx >> (8 * (SLICES - 1 - q))) & 0xFF for q in range(SLICES)])
    assert ((fx >> 56) == 0).all()
    return b, fx


def product_like_the_kernel(e, a, b):
    """trigemm_i8_kernel: S_t = sum_{p+q=t} a_p . b_q for t = 0..7 (int32 on the GPU), Horner in 2^-8, scale 2^(e+1-14)."""
    S = []
    for t in range(8):
        s = 0
        for p in range(SLICES):
            q = t - p
            if 0 <= q < SLICES:
                s += int(a[p] @ b[q])
        assert abs(s) < 2 ** 31, (t, s)                      # the TMEM accumulators are int32
        S.append(s)
    acc = np.float64(S[7])
    for t in range(6, -1, -1):
        acc = acc * np.float64(2.0 ** -8) + np.float64(S[t])    # a * 2^-8 is exact, so this equals the kernel's fma
    return float(acc) * 2.0 ** (e + 1 - 14), S


@pytest.mark.parametrize("n,seed", [(512, 0), (4096, 1), (8192, 2)])
def test_digit_slice_product_matches_the_exact_fixed_point_product(n, seed):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(n) * np.exp(-rng.random(n) * 8.0)           # entries over several orders of magnitude
    w[rng.integers(n)] = 37.5                                           # a dominant "diagonal" entry sets the row scale
    k = np.exp(-0.5 * rng.random(n) * 40.0)                             # kernel values in (0, 1], mostly small
    k[rng.integers(n)] = 1.0 + 1e-4                                     # the jitter quirk: slightly above one
    e, a, fxw = w_digits(w)
    b, fxk = k_digits(k)
    v, S = product_like_the_kernel(e, a, b)
    exact_fixed = Fraction(sum(int(x) * int(y) for x, y in zip(fxw, fxk)), 1) * Fraction(2) ** (e - 55 - 54)
    row_scale = 2.0 ** e
    assert abs(Fraction(v) - exact_fixed) <= Fraction(row_scale) * Fraction(2) ** -49
    # against the exact product of the ORIGINAL doubles: no worse than a float64 dot product's error bound
    exact = sum(Fraction(float(x)) * Fraction(float(y)) for x, y in zip(w, k))
    bound = n * 2.0 ** -53 * float(np.abs(w) @ k)                       # classic bound of a length-n fp64 dot product
    assert abs(Fraction(v) - exact) <= Fraction(bound) + Fraction(row_scale) * Fraction(2) ** -49
    assert abs(v - float(w @ k)) <= 2.0 * bound


def test_level_sums_cannot_overflow_int32_at_the_supported_sizes():
    # unsigned panel digits (n_pad <= 8192): 7 pairs * K * 128 * 255; signed digits (n_pad <= 16384): 7 * K * 2^14
    assert 7 * 8192 * 128 * 255 < 2 ** 31
    assert 7 * 16384 * 2 ** 14 < 2 ** 31
    assert 7 * 16384 * 128 * 255 >= 2 ** 31          # why the unsigned variant stops at 8192 rows


@pytest.mark.parametrize("K", [512, 4096, 8192])
def test_worst_case_of_the_dropped_levels_is_5K_2pow_minus_62_of_the_row_scale(K):
    """Adversarial digits (every W digit +127, every panel byte 255): the dropped levels t = 8..12 then reach their
    maximum, which the documentation states as 5 K 2^-62 of the row scale 2^e."""
    a = np.full((SLICES, K), 127, dtype=np.int64)
    b = np.full((SLICES, K), 255, dtype=np.int64)
    dropped = Fraction(0)
    for t in range(8, 2 * SLICES - 1):
        s = sum(int(a[p] @ b[t - p]) for p in range(SLICES) if 0 <= t - p < SLICES)
        dropped += Fraction(s) * Fraction(2) ** (-14 - 8 * t) * 2          # relative to the row scale 2^e: 2^{e+1-14-8t} / 2^e
    assert dropped <= Fraction(5 * K) * Fraction(2) ** -62 * Fraction(101, 100)
    assert dropped >= Fraction(5 * K) * Fraction(2) ** -62 * Fraction(95, 100)
