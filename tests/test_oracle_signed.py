"""CPU tests of the oracle's signed extension (ORC_INT2/4/8, piquant_oracle.h).

The reference has no signed dtypes at this commit, so there is nothing to pin these to ("parity unpinned" in the
oracle header).  What can be checked is that the definition -- intN is the offset-binary view of the pinned uintN
functions -- agrees with the textbook formula  q = clamp(round_half_away(x / scale) + zp, -2^(N-1), 2^(N-1) - 1),
two's complement fields, on ordinary inputs, and that the parameter formula is the reference's with q_min = -2^(N-1)
(reference src/piquant.cpp:245-258).
"""
from __future__ import annotations

import numpy as np
import pytest

from oracle import port
from oracle.port import (ADD, BF16, BITS, F32, INT2, INT4, INT8, NEAREST, SET, SIGNED, STOCHASTIC, UINT2, UINT4, UINT8,
                         bf16_bits_to_f32, f32_to_bf16_bits)

SIGNED_TYPES = (INT2, INT4, INT8)


def unpack_signed(q: np.ndarray, dt: int, numel: int) -> np.ndarray:
    bits = BITS[dt]
    per = 8 // bits
    shifts = (np.arange(per, dtype=np.uint8) * bits)[None, :]
    f = ((q[:, None] >> shifts) & ((1 << bits) - 1)).reshape(-1)[:numel].astype(np.int64)
    return np.where(f >= (1 << (bits - 1)), f - (1 << bits), f)


def textbook(x32: np.ndarray, dt_in: int, dt: int, scale: float, zp: int, mode: int, xi: float) -> np.ndarray:
    bits = BITS[dt]
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    inv = np.float32(1.0) / np.float32(scale)
    p = (x32 * inv).astype(np.float32)
    if mode == STOCHASTIC:
        tr = np.trunc(p)
        dec = np.abs(p - tr)
        t = tr + np.where(np.float32(xi) < dec, np.where(p < 0, -1.0, 1.0), 0.0)
    elif dt_in == F32 and dt == INT2:                       # no SIMD body in the reference for f32 -> 2 bit: std::round
        t = np.sign(p) * np.floor(np.abs(p.astype(np.float64)) + 0.5)
    else:                                                   # SIMD body: trunc(p +- 0.5) in float32
        t = np.trunc((p + np.copysign(np.float32(0.5), p)).astype(np.float32))
    return np.clip(t.astype(np.int64) + zp, lo, hi)


@pytest.mark.parametrize("dt", SIGNED_TYPES)
@pytest.mark.parametrize("dt_in", (F32, BF16))
@pytest.mark.parametrize("mode", (NEAREST, STOCHASTIC))
def test_signed_quantize_is_the_textbook_formula(dt, dt_in, mode):
    rng = np.random.default_rng(21)
    bits = BITS[dt]
    for n in (1, 2, 3, 5, 17, 1000, 4099):
        for scale, zp in ((2.0 / ((1 << bits) - 1), 0), (0.037, -1), (0.5, (1 << (bits - 1)) - 1), (1.0, -(1 << (bits - 1))), (0.01, 3)):
            x = rng.uniform(-2.0, 2.0, n).astype(np.float32)
            x[: min(n, 8)] = (np.array([0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5], np.float32) * np.float32(scale))[: min(n, 8)]
            xin = x if dt_in == F32 else f32_to_bf16_bits(x)
            x32 = x if dt_in == F32 else bf16_bits_to_f32(xin)
            q = port.quantize(xin, dt, scale, zp, mode, xi=0.3)
            assert q.size == (n * bits + 7) // 8
            assert np.array_equal(unpack_signed(q, dt, n), textbook(x32, dt_in, dt, scale, zp, mode, 0.3)), (n, scale, zp)
            # fields of elements that do not exist stay zero
            if (n * bits) % 8:
                assert q[-1] >> ((n * bits) % 8) == 0


@pytest.mark.parametrize("dt", SIGNED_TYPES)
def test_signed_is_offset_binary_of_unsigned(dt):
    rng = np.random.default_rng(22)
    bits = BITS[dt]
    udt = SIGNED[dt]
    off = 1 << (bits - 1)
    sign = {2: 0xAA, 4: 0x88, 8: 0x80}[bits]
    n = 4096
    x = rng.uniform(-3, 3, n).astype(np.float32)
    x[:6] = [np.nan, np.inf, -np.inf, 3e9, -3e9, 1e20]
    with np.errstate(all="ignore"):
        qs = port.quantize(x, dt, 0.05, -2, NEAREST)
        qu = port.quantize(x, udt, 0.05, -2 + off, NEAREST)
    assert np.array_equal(qs, qu ^ sign)
    for dt_out in (F32, BF16):
        for op in (SET, ADD):
            acc = rng.uniform(-1, 1, n).astype(np.float32)
            acc = acc if dt_out == F32 else f32_to_bf16_bits(acc)
            a = port.dequantize(qs, dt, n, dt_out, 0.05, -2, op, out=acc.copy())
            b = port.dequantize(qu, udt, n, dt_out, 0.05, -2 + off, op, out=acc.copy())
            assert np.array_equal(a, b)
    a = port.requantize(x[6:], dt, 0.05, -2, NEAREST, 0.0, SET)
    b = port.requantize(x[6:], udt, 0.05, -2 + off, NEAREST, 0.0, SET)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("dt", SIGNED_TYPES)
def test_signed_dequantize_f32_is_q_minus_zp_times_scale(dt):
    """u8/u4 -> f32 (and every SET cell up to rounding): float(q - zp) * scale with the SIGNED q and zp."""
    rng = np.random.default_rng(23)
    bits = BITS[dt]
    n = 1001
    fields = rng.integers(0, 1 << bits, n).astype(np.uint8)
    per = 8 // bits
    pad = (-n) % per
    f = np.concatenate([fields, np.zeros(pad, np.uint8)]).reshape(-1, per)
    q = np.zeros(f.shape[0], np.uint8)
    for k in range(per):
        q |= (f[:, k] << (k * bits)).astype(np.uint8)
    qs = unpack_signed(q, dt, n)
    for scale, zp in ((0.1, 0), (0.037, -3), (2.0, 5)):
        got = port.dequantize(q, dt, n, F32, scale, zp, SET)
        want = ((qs - zp).astype(np.float32) * np.float32(scale)).astype(np.float32)
        if dt == INT2:      # generic kernel: same formula through int64 -> float
            assert np.array_equal(got, want)
        else:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        got_b = bf16_bits_to_f32(port.dequantize(q, dt, n, BF16, scale, zp, SET))
        assert np.all(np.abs(got_b - want) <= np.abs(want) * 2.0**-7 + scale * 2.0**-6)


def test_signed_params_known_answers():
    # reference formula (src/piquant.cpp:245-258) with q_min = -2^(N-1):  zp = clamp(round(q_min - min/scale))
    s, z = port.params_from_minmax(-1.0, 1.0, INT8)
    assert s == np.float32(2.0 / 255.0) and z == -1            # round(-128 + 127.5) = round(-0.5) = -1 (half away)
    s, z = port.params_from_minmax(-3.0, 5.0, INT8)
    assert s == np.float32(8.0 / 255.0) and z == -32           # round(-128 + 95.625)
    s, z = port.params_from_minmax(-1.0, 1.0, INT4)
    assert s == np.float32(2.0 / 15.0) and z == -1             # round(-8 + 7.5) = round(-0.5) = -1 (half away)
    s, z = port.params_from_minmax(-1.0, 1.0, INT2)
    assert s == np.float32(2.0 / 3.0) and z == -1              # round(-2 + 1.5) = round(-0.5) = -1
    s, z = port.params_from_minmax(0.0, 1.0, INT8)
    assert z == -128
    s, z = port.params_from_minmax(-1.0, 0.0, INT8)
    assert z == 127
    for dt in SIGNED_TYPES:                                     # constant input: scale 1, signed midpoint
        assert port.params_from_minmax(0.25, 0.25, dt) == (1.0, -1)
    # same scale as the unsigned type of the same width
    for dt in SIGNED_TYPES:
        assert port.params_from_minmax(-0.7, 1.9, dt)[0] == port.params_from_minmax(-0.7, 1.9, SIGNED[dt])[0]


@pytest.mark.parametrize("dt", SIGNED_TYPES)
def test_signed_round_trip_within_half_a_step(dt):
    rng = np.random.default_rng(24)
    x = rng.uniform(-1, 1, 20000).astype(np.float32)
    s, z = port.compute_quant_params(x, dt)
    q = port.quantize(x, dt, s, z, NEAREST)
    y = port.dequantize(q, dt, x.size, F32, s, z, SET)
    assert np.abs(y - x).max() <= 0.5 * s * (1 + 1e-5) + 1e-7
