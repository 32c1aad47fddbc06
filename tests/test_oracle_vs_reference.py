"""Pins the CPU oracle (oracle/piquant_oracle.c) against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref/libpiquant_ref.so, see oracle/Makefile).

* ORC_SEM_REF emulation must equal the reference byte for byte, on random AND adversarial inputs,
  for every cell of the reference's dispatch tables (kernels.inl:108-149).
* ORC_SEM_BODY (what the CUDA library implements) may differ from the reference only at the inputs
  where the reference disagrees with itself (SIMD body vs scalar tail).

CPU-only; skipped when the reference library could not be built (no /root/reference).
"""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from helpers import (DEQUANT_CELLS, QUANT_CELLS, aligned, as_f32, cell_id, infer_xi, make_input,
                     special_values, unpack)
from oracle import port, ref
from oracle.port import (ADD, BF16, BITS, F32, NEAREST, SEM_BODY, SEM_REF, SET, STOCHASTIC, UINT2, UINT4,
                         UINT8, f32_to_bf16_bits, packed_bytes)

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libpiquant_ref.so not built")

NT = 4  # thread count of the reference context AND of the oracle's partition emulation
needs_avx512 = pytest.mark.skipif(not ref.cpu_isa().startswith("avx512"),
                                  reason="ORC_SEM_REF emulates the AVX-512 build's head/body/tail split")


@pytest.fixture(scope="module")
def ctx():
    c = ref.Context(NT)
    yield c
    c.close()


# ------------------------------------------------------------------------------------------------
# quantize, nearest
# ------------------------------------------------------------------------------------------------

@needs_avx512
@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_nearest_random_bit_exact(ctx, cell):
    dt_in, dt_out = cell
    rng = np.random.default_rng(0x9032002)            # seed of the reference's own tests (test/quant.cpp:31)
    for it in range(25):
        n = int(rng.integers(1, 20000)) if it else 1
        scale = float(np.float32(rng.uniform(0.1, 1.0)))
        zp = int(rng.integers(-128, 256))
        x = make_input(rng, n, dt_in, -1.0 * (1 + it % 3 * 20), 1.0 * (1 + it % 3 * 20))
        off = int(rng.integers(0, 16))
        o_ref = aligned(packed_bytes(dt_out, n), off=off)
        o_emu = aligned(packed_bytes(dt_out, n), off=off)
        ctx.quantize(x, dt_out, scale, zp, NEAREST, out=o_ref)
        port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_REF, nthreads=NT, out=o_emu)
        o_body = port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_BODY)
        assert np.array_equal(o_ref, o_emu), f"SEM_REF != reference (n={n}, scale={scale}, zp={zp})"
        # on U(-a,a) data |x/scale| == pred(0.5) has probability ~1e-10 per element: BODY == reference too
        assert np.array_equal(o_ref, o_body), f"SEM_BODY != reference (n={n}, scale={scale}, zp={zp})"


@needs_avx512
@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_nearest_adversarial_bit_exact(ctx, cell):
    """NaN, +-inf, |p| >= 2^31, ties, pred(0.5), zero points outside int32: the emulation of the
    reference (x86 cvtt 'integer indefinite', wrapping add, head/body/tail split) matches exactly."""
    dt_in, dt_out = cell
    rng = np.random.default_rng(1)
    qmax = (1 << BITS[dt_out]) - 1
    body_differs = 0
    for scale in (1.0, 0.25, 0.1, 0.0078431):
        sp = special_values(scale)
        for zp in (0, 1, 7, 128, 255, -1, -128, 2**31 - 1, -2**31, 2**31, 2**40 + 3, -2**40 - 5):
            for n_extra in (0, 1, 3, 64, 130, 1000):
                pad = rng.uniform(-2, 2, n_extra).astype(np.float32)
                xf = np.concatenate([pad, sp, pad, sp])
                x = xf if dt_in == F32 else f32_to_bf16_bits(xf)
                n = x.size
                off = int(rng.integers(0, 16))
                o_ref = aligned(packed_bytes(dt_out, n), off=off)
                o_emu = aligned(packed_bytes(dt_out, n), off=off)
                ctx.quantize(x, dt_out, scale, zp, NEAREST, out=o_ref)
                port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_REF, nthreads=NT, out=o_emu)
                assert np.array_equal(o_ref, o_emu), f"scale={scale} zp={zp} n={n} off={off}"
                # BODY semantics: may differ from the reference ONLY where trunc(p+-.5) != round(p)
                o_body = port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_BODY)
                if not np.array_equal(o_body, o_ref):
                    body_differs += 1
                    a, b = unpack(o_body, dt_out, n), unpack(o_ref, dt_out, n)
                    inv = np.float32(1.0) / np.float32(scale)
                    with np.errstate(all="ignore"):
                        p = np.abs(as_f32(x) * inv)
                    idx = np.flatnonzero(a != b)
                    ok = (p[idx] == np.float32(0.49999997)) | ((p[idx] >= 2.0**23) & (p[idx] < 2.0**24))
                    assert ok.all(), f"unexpected BODY/reference difference at p={p[idx][~ok][:4]}"
    if (dt_in, dt_out) == (F32, UINT8):       # 64-wide bodies: some pred(0.5) always lands in a scalar tail
                                              # (bf16 cannot represent pred(0.5)*scale at all)
        assert body_differs > 0               # the self-inconsistency is real and we exercised it
    assert qmax in (3, 15, 255)


# ------------------------------------------------------------------------------------------------
# quantize, stochastic (one xi per call, drawn inside the reference from random_device)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_stochastic_consistent_with_one_threshold(ctx, cell):
    dt_in, dt_out = cell
    rng = np.random.default_rng(7)
    qmax = (1 << BITS[dt_out]) - 1
    pinned = 0
    for it in range(30):
        n = int(rng.integers(1000, 12000))
        scale = float(np.float32(rng.uniform(0.1, 1.0)))
        zp = int(rng.integers(0, qmax + 1))
        x = make_input(rng, n, dt_in)
        o_ref = ctx.quantize(x, dt_out, scale, zp, STOCHASTIC)
        xi = infer_xi(as_f32(x), scale, zp, qmax, unpack(o_ref, dt_out, n))
        assert xi is not None, "reference output is not explainable by a single per-call threshold"
        o_emu = port.quantize(x, dt_out, scale, zp, STOCHASTIC, xi=xi, semantics=SEM_REF, nthreads=NT)
        o_body = port.quantize(x, dt_out, scale, zp, STOCHASTIC, xi=xi, semantics=SEM_BODY)
        assert np.array_equal(o_ref, o_emu)
        assert np.array_equal(o_ref, o_body)           # stochastic has no SIMD body: both semantics coincide
        pinned += 1
    assert pinned == 30


# ------------------------------------------------------------------------------------------------
# dequantize
# ------------------------------------------------------------------------------------------------

@needs_avx512
@pytest.mark.parametrize("cell", DEQUANT_CELLS, ids=cell_id)
def test_dequantize_bit_exact(ctx, cell):
    """All 12 dequantize cells: emulation == reference bit for bit (including the FMA contraction GCC
    applies to the f32 ADD accumulate and the u2->f32 tail that ignores ADD, dequantize.inl:72-86)."""
    dt_in, dt_out, op = cell
    rng = np.random.default_rng(2)
    for it in range(30):
        n = int(rng.integers(1, 3000)) if it % 3 else int(rng.integers(3000, 40000))
        q = rng.integers(0, 256, packed_bytes(dt_in, n)).astype(np.uint8)
        scale = float(np.float32(rng.uniform(0.001, 1.0)))
        zp = int(rng.integers(0, 1 << BITS[dt_in]))
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
        o_ref, o_emu, o_body = prev.copy(), prev.copy(), prev.copy()
        ctx.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=o_ref)
        port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=o_emu, semantics=SEM_REF, nthreads=NT)
        port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=o_body, semantics=SEM_BODY)
        assert np.array_equal(o_ref, o_emu), f"n={n} scale={scale} zp={zp}"
        if dt_out == F32:
            assert np.array_equal(o_ref, o_body)       # f32 outputs: body and tail formulas coincide
        else:
            # bf16 outputs: scalar tails round twice / use (q-zp)*s instead of fma(q,s,-zp*s) (so q == zp
            # gives exactly 0 in a tail and the rounding residue of zp*s in the body); at most one bf16
            # ulp of the value apart and far inside the 0.5*scale tolerance of the north star
            a, b = as_f32(o_ref).astype(np.float64), as_f32(o_body).astype(np.float64)
            d = np.abs(a - b)
            if op == SET:
                assert d.max() <= 0.5 * scale
            qmax = (1 << BITS[dt_in]) - 1      # |dequantized term| <= qmax*scale is itself rounded to bf16 in ADD tails
            assert (d <= (np.maximum(np.abs(a), np.abs(b)) + qmax * scale) * 2.0**-7 + scale * 1e-6).all()


def test_dequantize_uint2_f32_add_tail_bug_is_reproduced(ctx):
    """Reference quirk: dequant_uint2 SETs its 1-3 element tail even for ADD."""
    n = 7
    q = np.array([0b11100100, 0b00011011], dtype=np.uint8)
    prev = np.full(n, 100.0, dtype=np.float32)
    o_ref = ctx.dequantize(q, UINT2, n, F32, 1.0, 0, ADD, out=prev.copy())
    o_emu = port.dequantize(q, UINT2, n, F32, 1.0, 0, ADD, out=prev.copy())
    assert np.array_equal(o_ref, o_emu)
    assert o_ref[:4].tolist() == [100.0, 101.0, 102.0, 103.0]
    assert o_ref[4:].tolist() == [3.0, 2.0, 1.0]          # tail overwritten, not accumulated


# ------------------------------------------------------------------------------------------------
# fused quantize->dequantize ("requantize", C++ API only: piquant.hpp:276-285)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_io,dt_q,op", list(itertools.product((F32, BF16), (UINT2, UINT4, UINT8), (SET, ADD))),
                         ids=lambda v: str(v))
def test_requantize_nearest_bit_exact(ctx, dt_io, dt_q, op):
    if not hasattr(ref.lib(), "piquant_ref_shim_requantize"):
        pytest.skip("reference library built without oracle/ref_shim.cpp")
    rng = np.random.default_rng(3)
    fma = ref.cpu_isa() not in ("sse42", "generic")     # scalar loop is contracted only in FMA-enabled TUs
    for it in range(10):
        n = int(rng.integers(1, 20000))
        x = make_input(rng, n, dt_io)
        scale, zp = ctx.compute_quant_params(x, dt_q)
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_io == F32 else f32_to_bf16_bits(prev)
        o_ref = ctx.requantize(x, dt_q, scale, zp, NEAREST, op, out=prev.copy())
        o_emu = port.requantize(x, dt_q, scale, zp, NEAREST, 0.0, op, out=prev.copy(), fma_add=fma)
        assert np.array_equal(o_ref, o_emu), f"n={n}"


# ------------------------------------------------------------------------------------------------
# min/max -> (scale, zero_point)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_in", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("dt_q", (UINT2, UINT4, UINT8), ids=("u2", "u4", "u8"))
def test_compute_quant_params_bit_equal(ctx, dt_in, dt_q):
    rng = np.random.default_rng(4)
    cases = [make_input(rng, int(rng.integers(1, 50000)), dt_in, lo, hi)
             for lo, hi in ((-1, 1), (0, 1), (1, 2), (-5, -1), (-1e-3, 1e3), (-3e38, 3e38))]
    consts = [np.full(100, v, np.float32) for v in (42.0, 0.0, -7.5)]
    cases += [c if dt_in == F32 else f32_to_bf16_bits(c) for c in consts]
    cases += [np.array([-1, 1], np.float32) if dt_in == F32 else f32_to_bf16_bits(np.array([-1, 1], np.float32))]
    for x in cases:
        s_ref, z_ref = ctx.compute_quant_params(x, dt_q)
        s_emu, z_emu = port.compute_quant_params(x, dt_q)
        assert np.float32(s_ref).tobytes() == np.float32(s_emu).tobytes() and z_ref == z_emu


# ------------------------------------------------------------------------------------------------
# degenerate scales: 0, +-inf, NaN, negative, denormal, huge -- the oracle still equals the reference
# ------------------------------------------------------------------------------------------------

DEGENERATE_SCALES = (0.0, -0.0, float("inf"), float("nan"), -1.0, -0.037, 1e-45, 1e-30, 1e30, 3.4e38)


@needs_avx512
@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_degenerate_scales_bit_exact(ctx, cell):
    dt_in, dt_out = cell
    rng = np.random.default_rng(31)
    for scale in DEGENERATE_SCALES:
        for zp in (0, 3, 128, -7, 2**31 - 1):
            n = int(rng.integers(200, 3000))
            x = make_input(rng, n, dt_in, -300.0, 300.0)
            sp = special_values(1.0)
            x[5:5 + sp.size] = sp if dt_in == F32 else f32_to_bf16_bits(sp)
            for mode, xi in ((NEAREST, 0.0),):
                o_ref = aligned(packed_bytes(dt_out, n))
                o_emu = aligned(packed_bytes(dt_out, n))
                ctx.quantize(x, dt_out, scale, zp, mode, out=o_ref)
                with np.errstate(all="ignore"):
                    port.quantize(x, dt_out, scale, zp, mode, xi=xi, semantics=SEM_REF, nthreads=NT, out=o_emu)
                assert np.array_equal(o_ref, o_emu), f"scale={scale} zp={zp} n={n}"


@needs_avx512
@pytest.mark.parametrize("cell", DEQUANT_CELLS, ids=cell_id)
def test_dequantize_degenerate_scales_bit_exact(ctx, cell):
    dt_in, dt_out, op = cell
    rng = np.random.default_rng(32)
    for scale in DEGENERATE_SCALES:
        for zp in (0, 5, 255, 2**22 + 1, -2**31):
            n = int(rng.integers(100, 3000))
            q = rng.integers(0, 256, packed_bytes(dt_in, n)).astype(np.uint8)
            prev = rng.uniform(-100, 100, n).astype(np.float32)
            prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
            o_ref, o_emu = prev.copy(), prev.copy()
            ctx.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=o_ref)
            with np.errstate(all="ignore"):
                port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=o_emu, semantics=SEM_REF, nthreads=NT)
            if dt_out == F32:
                a, b = o_ref.view(np.uint32), o_emu.view(np.uint32)
                nan_a, nan_b = np.isnan(o_ref), np.isnan(o_emu)
                assert np.array_equal(nan_a, nan_b) and np.array_equal(a[~nan_a], b[~nan_b]), f"scale={scale} zp={zp} n={n}"
            else:
                # bf16 outputs below 2^-126: the AVX-512-BF16 instruction (vcvtneps2bf16) treats denormal inputs as zero,
                # the reference's software conversion and scalar tails keep them -- ISA-dependent, like NaN payloads; the
                # oracle (and the CUDA library) keep them.  Compare with denormals flushed to signed zero on both sides.
                def ftz(v):
                    return np.where((v & 0x7F80) == 0, v & 0x8000, v)
                nan_a, nan_b = (o_ref & 0x7FFF) > 0x7F80, (o_emu & 0x7FFF) > 0x7F80
                assert np.array_equal(nan_a, nan_b) and np.array_equal(ftz(o_ref[~nan_a]), ftz(o_emu[~nan_b])), f"scale={scale} zp={zp} n={n}"
