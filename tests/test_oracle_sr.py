"""CPU tests of the oracle's per-element stochastic rounding (extension, piquant_oracle.h: orc_quantize_sr).
Parity unpinned (the reference has no such mode); pinned here: the Philox4x32-10 core against the Random123
known-answer vectors, the rounding rule against a direct numpy restatement, and the statistical property the
mode exists for -- unbiasedness per element, which the reference's one-threshold-per-call mode does not have."""
from __future__ import annotations

import numpy as np
import pytest

from helpers import unpack
from oracle import port
from oracle.port import BF16, BITS, F32, INT8, STOCHASTIC, UINT2, UINT4, UINT8, bf16_bits_to_f32, f32_to_bf16_bits


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, "philox4x32 10" lines
    assert port.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert port.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert port.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def numpy_sr(x32: np.ndarray, bits: int, scale: float, zp: int, key: int, base: int = 0) -> np.ndarray:
    n = x32.size
    j = base + np.arange(n, dtype=np.int64)
    words = {}
    k = np.zeros(n, dtype=np.int64)
    for i in range(n):
        g = int(j[i]) >> 3
        if g not in words:
            words[g] = port.philox4x32_10([g & 0xFFFFFFFF, g >> 32, 0, 0], [key & 0xFFFFFFFF, key >> 32])
        k[i] = (words[g][(int(j[i]) & 7) >> 1] >> (16 * (int(j[i]) & 1))) & 0xFFFF
    p = (x32 * (np.float32(1.0) / np.float32(scale))).astype(np.float32)
    t = np.floor(p.astype(np.float64) + (k + 0.5) / 65536.0).astype(np.int64)
    return np.clip(t + zp, 0, (1 << bits) - 1).astype(np.uint8)


@pytest.mark.parametrize("dt_in", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("dt_out", (UINT2, UINT4, UINT8), ids=("u2", "u4", "u8"))
def test_sr_rule_matches_numpy_restatement(dt_in, dt_out):
    rng = np.random.default_rng(41)
    bits = BITS[dt_out]
    for n, base in ((1, 0), (7, 0), (1000, 0), (1000, 4096), (333, 2**36)):
        x = rng.uniform(-2, 2, n).astype(np.float32)
        x[: min(n, 6)] = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5], np.float32)[: min(n, 6)]
        xin = x if dt_in == F32 else f32_to_bf16_bits(x)
        x32 = x if dt_in == F32 else bf16_bits_to_f32(xin)
        for scale, zp, key in ((0.1, 3, 1), (2.0 / ((1 << bits) - 1), (1 << bits) // 2, 0xDEADBEEFCAFEF00D)):
            q = port.quantize_sr(xin, dt_out, scale, zp, key, base)
            assert np.array_equal(unpack(q, dt_out, n), numpy_sr(x32, bits, scale, zp, key, base)), (n, base, scale)


def test_sr_stream_is_a_function_of_the_element_index():
    """Quantizing a suffix with base = its offset reproduces the suffix of the whole tensor (host-pointer chunking relies on it)."""
    rng = np.random.default_rng(42)
    x = rng.uniform(-1, 1, 5000).astype(np.float32)
    whole = port.quantize_sr(x, UINT8, 0.01, 128, 77)
    for off in (8, 64, 4096):
        assert np.array_equal(port.quantize_sr(x[off:], UINT8, 0.01, 128, 77, base=off), whole[off:])
    assert not np.array_equal(port.quantize_sr(x, UINT8, 0.01, 128, 78), whole)      # another key, other bits


def test_sr_is_unbiased_per_element_and_per_call_threshold_is_not():
    n = 200_000
    x = np.full(n, 0.3 * 0.01, np.float32)              # every element sits 0.3 of a step above 128
    q = port.quantize_sr(x, UINT8, 0.01, 128, 2024).astype(np.float64)
    p = float(np.float32(0.3 * 0.01) * (np.float32(1) / np.float32(0.01)))
    assert set(np.unique(q)) == {128.0, 129.0}
    assert abs(q.mean() - (128 + p)) < 4 * np.sqrt(0.21 / n)          # 4 sigma of a Bernoulli(0.3) mean
    # the reference's mode: one threshold for the whole call -> all elements round the same way, mean off by 0.3 or 0.7
    q1 = port.quantize(x, UINT8, 0.01, 128, STOCHASTIC, xi=0.6).astype(np.float64)
    assert q1.std() == 0.0 and abs(q1.mean() - (128 + p)) > 0.29
    # integers never move, whatever the random bits
    xi = (np.arange(-100, 100, dtype=np.float32) * np.float32(0.5))
    qi = port.quantize_sr(xi, UINT8, 0.5, 128, 5)
    assert np.array_equal(qi, np.clip(np.arange(-100, 100) + 128, 0, 255).astype(np.uint8))


def test_sr_signed_and_special_values():
    x = np.array([np.nan, np.inf, -np.inf, 3e9, -3e9, 1e20, 8388609.0, -8388609.0, 16777216.0, 1e-40], np.float32)
    with np.errstate(all="ignore"):
        q = port.quantize_sr(x, UINT8, 1.0, 0, 9)
    # NaN / +-inf / huge: (int64)p by the x86 rule (INT64_MIN for NaN, inf and |p| >= 2^63), then + zp and clamp
    assert list(q) == [0, 0, 0, 255, 0, 0, 255, 0, 255, 0]
    rng = np.random.default_rng(43)
    y = rng.uniform(-1, 1, 1001).astype(np.float32)
    a = port.quantize_sr(y, INT8, 2 / 255, -1, 11)
    b = port.quantize_sr(y, UINT8, 2 / 255, 127, 11)
    assert np.array_equal(a, b ^ 0x80)


def test_sr_requantize_is_quantize_then_generic_dequantize():
    """orc_requantize_sr = per-element quantize followed by the generic dequant_step of orc_requantize (f32: one multiply)."""
    rng = np.random.default_rng(44)
    x = rng.uniform(-1.5, 1.5, 3001).astype(np.float32)
    for dt, scale, zp in ((UINT8, 2 / 255, 128), (UINT4, 0.2, 7), (INT8, 0.01, -3)):
        q = port.quantize_sr(x, dt, scale, zp, 31337, base=64)
        y = port.requantize_sr(x, dt, scale, zp, 31337, base=64)
        assert np.array_equal(y.view(np.uint32), port.dequantize(q, dt, x.size, F32, scale, zp).view(np.uint32))
    acc = rng.uniform(-1, 1, x.size).astype(np.float32)
    y = port.requantize_sr(x, UINT8, 2 / 255, 128, 7, op=1, out=acc.copy())
    q = port.quantize_sr(x, UINT8, 2 / 255, 128, 7)
    assert np.array_equal(y.view(np.uint32), port.dequantize(q, UINT8, x.size, F32, 2 / 255, 128, 1, out=acc.copy()).view(np.uint32))
