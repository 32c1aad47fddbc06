"""GPU parity: the sm_100a kernels, called through the C ABI of libpiquant.so, against the CPU oracle
(oracle/piquant_oracle.c, SEM_BODY semantics = what the CUDA library implements) and against the
committed golden vectors produced by the unmodified reference.

Bar: bit-exact for every integer / packed output, for (scale, zero_point) and for every f32 output;
bf16 outputs bit-exact except NaN payloads (hardware cvt gives the canonical NaN).
"""
from __future__ import annotations

import itertools
from pathlib import Path

import numpy as np
import pytest

from helpers import DEQUANT_CELLS, QUANT_CELLS, as_f32, cell_id, make_input, special_values, unpack
from oracle import port
from oracle.port import (ADD, BF16, BITS, F32, NEAREST, SEM_BODY, SET, STOCHASTIC, UINT2, UINT4, UINT8, f32_to_bf16_bits,
                         packed_bytes)

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).parent / "golden" / "piquant_golden.npz"
DTN = {"f32": F32, "bf16": BF16, "u2": UINT2, "u4": UINT4, "u8": UINT8}
EDGE_SIZES = (1, 2, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 1023, 1025, 4095, 4097)
VARIANTS = (1, 2)     # 1 = direct LDG/STG kernels, 2 = TMA ring kernels


@pytest.fixture(scope="module", params=VARIANTS, ids=("direct", "tma"))
def gpu(request):
    from gpu_util import Gpu
    return Gpu(variant=request.param)


@pytest.fixture(scope="module")
def gpu0():
    from gpu_util import Gpu
    return Gpu(variant=0)


def bf16_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """bit-equal, except that any NaN matches any NaN"""
    an = (a & 0x7FFF) > 0x7F80
    bn = (b & 0x7FFF) > 0x7F80
    return bool(np.array_equal(an, bn) and np.array_equal(a[~an], b[~bn]))


# ------------------------------------------------------------------------------------------------
# quantize
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_nearest_sizes_bit_exact(gpu, cell):
    dt_in, dt_out = cell
    rng = np.random.default_rng(0x9032002)      # seed of the reference's own tests (test/quant.cpp:31)
    sizes = list(EDGE_SIZES) + [int(rng.integers(5000, 15000)) for _ in range(6)] + [100003, 1 << 20, (1 << 20) + 77]
    for n in sizes:
        scale = float(np.float32(rng.uniform(0.1, 1.0)))
        zp = int(rng.integers(-128, 256))
        x = make_input(rng, n, dt_in, -3.0, 3.0)
        want = port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_BODY)
        got = gpu.quantize(x, dt_out, scale, zp, NEAREST)
        assert np.array_equal(got, want), f"n={n} scale={scale} zp={zp}: {np.flatnonzero(got != want)[:8]}"


@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_nearest_any_alignment(gpu, cell):
    """void* ABI: no alignment promise.  Every (input, output) byte phase takes the head / ragged / byte path."""
    dt_in, dt_out = cell
    rng = np.random.default_rng(5)
    isz = 4 if dt_in == F32 else 2
    for in_off, out_off in itertools.product((0, isz, 3 * isz, 16, 16 + isz), (0, 1, 3, 7, 13, 16)):
        n = int(rng.integers(900, 9000))
        x = make_input(rng, n, dt_in)
        scale, zp = port.compute_quant_params(x, dt_out)
        want = port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_BODY)
        got = gpu.quantize(x, dt_out, scale, zp, NEAREST, in_off=in_off, out_off=out_off)
        assert np.array_equal(got, want), f"in_off={in_off} out_off={out_off} n={n}"


@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_nearest_adversarial(gpu, cell):
    """NaN, +-inf, |x/scale| >= 2^31, ties, pred(0.5), zero points outside int32: x86 'integer indefinite'
    conversion and wrapping adds are reproduced exactly."""
    dt_in, dt_out = cell
    rng = np.random.default_rng(1)
    for scale in (1.0, 0.25, 0.1, 0.0078431):
        sp = special_values(scale)
        for zp in (0, 1, 7, 128, 255, -1, -128, 2**31 - 1, -2**31, 2**31, 2**40 + 3, -2**40 - 5, 2**29, 2**29 + 1, -2**29 - 1):
            pad = rng.uniform(-2, 2, 130).astype(np.float32)
            xf = np.concatenate([pad, sp, pad, sp, pad[:7]])
            x = xf if dt_in == F32 else f32_to_bf16_bits(xf)
            want = port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_BODY)
            got = gpu.quantize(x, dt_out, scale, zp, NEAREST)
            assert np.array_equal(got, want), f"scale={scale} zp={zp}: {np.flatnonzero(got != want)[:8]}"


@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_stochastic_bit_exact_for_a_given_threshold(gpu, cell):
    dt_in, dt_out = cell
    rng = np.random.default_rng(7)
    qmax = (1 << BITS[dt_out]) - 1
    for n in (1, 17, 4097, 12345, 1 << 18):
        for xi in (0.0, 0.25, 0.5, 0.999):
            scale = float(np.float32(rng.uniform(0.1, 1.0)))
            zp = int(rng.integers(0, qmax + 1))
            x = make_input(rng, n, dt_in, -4.0, 4.0)
            if n > 100:
                sp = special_values(scale)
                x[10:10 + sp.size] = sp if dt_in == F32 else f32_to_bf16_bits(sp)
            want = port.quantize(x, dt_out, scale, zp, STOCHASTIC, xi=xi, semantics=SEM_BODY)
            got = gpu.quantize(x, dt_out, scale, zp, STOCHASTIC, xi=xi)
            assert np.array_equal(got, want), f"n={n} xi={xi}"
    # big zero points take the 64-bit path
    x = make_input(rng, 5000, dt_in, -4.0, 4.0)
    for zp in (2**40 + 3, -2**40 - 5, 2**63 - 1, -2**63):
        want = port.quantize(x, dt_out, 0.5, zp, STOCHASTIC, xi=0.3, semantics=SEM_BODY)
        got = gpu.quantize(x, dt_out, 0.5, zp, STOCHASTIC, xi=0.3)
        assert np.array_equal(got, want), f"zp={zp}"


def test_quantize_stochastic_default_draws_one_threshold_per_call(gpu0):
    """Without an explicit threshold every call draws one xi in [0,1) shared by all elements
    (reference src/piquant.cpp:199-201): constant input quantizes to a constant."""
    x = np.full(10000, 0.3, np.float32)
    seen = set()
    for _ in range(40):
        q = gpu0.quantize(x, UINT8, 1.0, 0, STOCHASTIC, xi=None)
        assert (q == q[0]).all() and q[0] in (0, 1)
        xi = gpu0.ctx.last_stochastic_threshold
        assert 0.0 <= xi < 1.0 and q[0] == (1 if xi < np.float32(0.3) else 0)
        seen.add(int(q[0]))
    assert seen == {0, 1}


def test_quantize_matches_reference_golden(gpu):
    g = np.load(GOLDEN)
    for key in [str(k) for k in g["__keys__"] if str(k).startswith("quant/")]:
        _, dti, dto, mode, n = key.split("/")
        x, out = g[key + "/x"], g[key + "/out"]
        scale, zp, xi = g[key + "/p"]
        got = gpu.quantize(x, DTN[dto], float(scale), int(zp), STOCHASTIC if mode == "st" else NEAREST, xi=float(xi) if mode == "st" else None)
        assert np.array_equal(got, out), key


# ------------------------------------------------------------------------------------------------
# dequantize
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cell", DEQUANT_CELLS, ids=cell_id)
def test_dequantize_bit_exact(gpu, cell):
    dt_in, dt_out, op = cell
    rng = np.random.default_rng(2)
    sizes = list(EDGE_SIZES) + [int(rng.integers(3000, 40000)) for _ in range(5)] + [1 << 20, (1 << 20) + 5]
    for n in sizes:
        q = rng.integers(0, 256, packed_bytes(dt_in, n)).astype(np.uint8)
        scale = float(np.float32(rng.uniform(0.001, 1.0)))
        zp = int(rng.integers(0, 1 << BITS[dt_in]))
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
        want = port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=prev.copy(), semantics=SEM_BODY)
        got = gpu.dequantize(q, dt_in, n, dt_out, scale, zp, op, prev=prev)
        if dt_out == F32:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"n={n} scale={scale} zp={zp}"
        else:
            assert bf16_equal(got, want), f"n={n} scale={scale} zp={zp}"


@pytest.mark.parametrize("cell", DEQUANT_CELLS, ids=cell_id)
def test_dequantize_any_alignment_and_odd_zero_points(gpu, cell):
    dt_in, dt_out, op = cell
    rng = np.random.default_rng(3)
    osz = 4 if dt_out == F32 else 2
    for in_off, out_off in itertools.product((0, 1, 3, 4, 8, 17), (0, osz, 3 * osz, 16, 32 + osz)):
        n = int(rng.integers(700, 6000))
        q = rng.integers(0, 256, packed_bytes(dt_in, n)).astype(np.uint8)
        zp = int(rng.choice([0, 3, -7, 300, -2**31, 2**31 - 1, 2**40 + 1, -2**35]))
        scale = float(np.float32(rng.uniform(0.01, 2.0)))
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
        want = port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=prev.copy(), semantics=SEM_BODY)
        got = gpu.dequantize(q, dt_in, n, dt_out, scale, zp, op, prev=prev, in_off=in_off, out_off=out_off)
        if dt_out == F32:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"in_off={in_off} out_off={out_off} n={n} zp={zp}"
        else:
            assert bf16_equal(got, want), f"in_off={in_off} out_off={out_off} n={n} zp={zp}"


def test_dequantize_uint2_f32_add_tail_quirk_is_kept(gpu):
    """The reference's generic u2->f32 kernel SETs its 1-3 element tail even for ADD (dequantize.inl:72-86)."""
    q = np.array([0b11100100, 0b00011011], dtype=np.uint8)
    prev = np.full(7, 100.0, np.float32)
    got = gpu.dequantize(q, UINT2, 7, F32, 1.0, 0, ADD, prev=prev)
    assert got.tolist() == [100.0, 101.0, 102.0, 103.0, 3.0, 2.0, 1.0]


def test_dequantize_matches_reference_golden(gpu):
    g = np.load(GOLDEN)
    for key in [str(k) for k in g["__keys__"] if str(k).startswith("dequant/")]:
        _, dti, dto, op, n = key.split("/")
        q, prev, out = g[key + "/q"], g[key + "/prev"], g[key + "/out"]
        scale, zp = g[key + "/p"]
        got = gpu.dequantize(q, DTN[dti], int(n), DTN[dto], float(scale), int(zp), ADD if op == "add" else SET, prev=prev)
        if dto == "f32":
            assert np.array_equal(got, out), key
        else:
            # the golden's last n % body-width elements come from the reference's scalar tails, which round
            # twice (see test_oracle_golden.py); the SIMD-body part must be bit-equal
            body = {"u8": 64, "u4": 128, "u2": 256}[dti]
            nb = (int(n) // 4 // body) * body
            assert np.array_equal(got[:nb], out[:nb]), key
            a, b = port.bf16_bits_to_f32(got).astype(np.float64), port.bf16_bits_to_f32(out).astype(np.float64)
            qmax = (1 << BITS[DTN[dti]]) - 1
            assert (np.abs(a - b) <= (np.maximum(np.abs(a), np.abs(b)) + qmax * scale) * 2.0**-7 + scale * 1e-6).all(), key


# ------------------------------------------------------------------------------------------------
# round trips and fused requantize
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_round_trip_within_half_a_step(gpu0, cell):
    """north star: |dequantize(quantize(x)) - x| <= 0.5 * scale (+ bf16 rounding of the output)."""
    dt_in, dt_q = cell
    rng = np.random.default_rng(11)
    n = 200003
    x = make_input(rng, n, dt_in)
    scale, zp = gpu0.compute_quant_params(x, dt_q)
    q = gpu0.quantize(x, dt_q, scale, zp, NEAREST)
    y = gpu0.dequantize(q, dt_q, n, dt_in, scale, zp, SET)
    err = np.abs(as_f32(y).astype(np.float64) - as_f32(x).astype(np.float64))
    tol = 0.5 * scale * (1 + 1e-6) + (2.0**-8 * np.abs(as_f32(x)) + 2.0**-8 * scale if dt_in == BF16 else 1e-7)
    assert (err <= tol).all(), f"max err {err.max()} vs {scale}"


@pytest.mark.parametrize("dt_io,dt_q,op,mode", list(itertools.product((F32, BF16), (UINT2, UINT4, UINT8), (SET, ADD), (NEAREST, STOCHASTIC))),
                         ids=lambda v: str(v))
def test_requantize_bit_exact(gpu0, dt_io, dt_q, op, mode):
    rng = np.random.default_rng(3)
    for n, off in ((1, 0), (7, 0), (4099, 0), (20000, 4), (1 << 18, 0)):
        x = make_input(rng, n, dt_io, -2.0, 2.0)
        scale, zp = port.compute_quant_params(x, dt_q)
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_io == F32 else f32_to_bf16_bits(prev)
        want = port.requantize(x, dt_q, scale, zp, mode, 0.4, op, out=prev.copy(), fma_add=True)
        got = gpu0.requantize(x, dt_q, scale, zp, mode, 0.4, op, prev=prev, in_off=off, out_off=off)
        if dt_io == F32:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"n={n}"
        else:
            assert bf16_equal(got, want), f"n={n}"


@pytest.mark.parametrize("dt_io", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("mode", (NEAREST, STOCHASTIC), ids=("nearest", "stochastic"))
def test_requantize_adversarial(gpu0, dt_io, mode):
    """The float-domain fast path of requantize.cu (magic-number rounding, clamp as two float min/max) against the
    oracle where it could go wrong: ties and their neighbours, results that round to zero from below (+0.0, never
    -0.0), NaN / inf / huge values inside an otherwise ordinary vector, zero points on both sides of the fast-path
    limits, and thresholds at the ends of [0, 1)."""
    rng = np.random.default_rng(11)
    one_below = float(np.nextafter(np.float32(1.0), np.float32(0.0)))
    for dt_q in (UINT2, UINT4, UINT8):
        qmax = (1 << BITS[dt_q]) - 1
        for scale in (1.0, 0.5, 2.0 / 255, 0.037):
            for zp in (0, 1, qmax // 2, qmax, 255, 256, -1, -7, 1000, 2**22, 2**22 + 1, -2**22 - 1, 2**31 - 1, 2**40 + 3):
                n = 4096 + 37
                x = make_input(rng, n, dt_io, -3.0 * scale * (qmax + 2), 3.0 * scale * (qmax + 2))
                sp = special_values(scale)
                grid = (np.arange(-40, 41, dtype=np.float32) * np.float32(0.25) * np.float32(scale)).astype(np.float32)
                near = np.concatenate([np.nextafter(grid, np.float32(np.inf)), np.nextafter(grid, np.float32(-np.inf))])
                extra = np.concatenate([sp, grid, near]).astype(np.float32)
                x[5:5 + extra.size] = extra if dt_io == F32 else f32_to_bf16_bits(extra)
                prev = rng.uniform(-1, 1, n).astype(np.float32)
                prev[::7] = -0.0
                prev = prev if dt_io == F32 else f32_to_bf16_bits(prev)
                for xi in ((0.4,) if mode == NEAREST else (0.0, 1e-30, 0.5, one_below)):
                    for op in (SET, ADD):
                        with np.errstate(all="ignore"):
                            want = port.requantize(x, dt_q, scale, zp, mode, xi, op, out=prev.copy(), fma_add=True)
                        got = gpu0.requantize(x, dt_q, scale, zp, mode, xi, op, prev=prev)
                        what = f"dt_q={dt_q} scale={scale} zp={zp} xi={xi} op={op}"
                        if dt_io == F32:
                            bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
                            bad = bad[~(np.isnan(got[bad]) & np.isnan(want[bad]))]
                            assert bad.size == 0, f"{what}: x={x[bad[:4]]} got={got[bad[:4]]} want={want[bad[:4]]}"
                        else:
                            assert bf16_equal(got, want), what


@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_stochastic_extreme_thresholds(gpu, cell):
    """sign(r) * ceil(|r| - xi), the form the kernels use, against the literal trunc / compare / add of
    quantize.inl:8-19 at the thresholds where a rounding slip would show: 0, denormal, tiny, just below 1."""
    dt_in, dt_out = cell
    rng = np.random.default_rng(12)
    qmax = (1 << BITS[dt_out]) - 1
    one_below = float(np.nextafter(np.float32(1.0), np.float32(0.0)))
    for xi in (0.0, 1e-45, 1e-30, 5.9604645e-08, 0.49999997, 0.5, 0.50000006, one_below):
        for scale, zp in ((1.0, 0), (0.25, qmax // 2), (2.0 / 255, qmax), (0.037, -3)):
            n = 3000
            x = make_input(rng, n, dt_in, -2.0 * scale * (qmax + 2), 2.0 * scale * (qmax + 2))
            sp = special_values(scale)
            grid = (np.arange(-40, 41, dtype=np.float32) * np.float32(0.25) * np.float32(scale)).astype(np.float32)
            near = np.concatenate([np.nextafter(grid, np.float32(np.inf)), np.nextafter(grid, np.float32(-np.inf))])
            extra = np.concatenate([sp, grid, near]).astype(np.float32)
            x[5:5 + extra.size] = extra if dt_in == F32 else f32_to_bf16_bits(extra)
            with np.errstate(all="ignore"):
                want = port.quantize(x, dt_out, scale, zp, STOCHASTIC, xi=xi, semantics=SEM_BODY)
            got = gpu.quantize(x, dt_out, scale, zp, STOCHASTIC, xi=xi)
            assert np.array_equal(got, want), f"xi={xi} scale={scale} zp={zp}: {np.flatnonzero(got != want)[:8]}"


def test_bf16_to_2bit_exhaustive_over_all_65536_inputs(gpu0):
    """bf16 -> uint2 / int2 runs on three threshold compares placed by the host (quantize.cu: quant_thresholds); the input
    domain is small enough to try EVERY bf16 bit pattern: once grouped so that whole vectors stay inside the fast path's
    domain (sorted by magnitude, then shuffled in blocks), once fully shuffled (NaN / inf / huge values in most vectors ->
    exact fallback), for scales from 1e-3 to 123, zero points inside and outside the range, both rounding modes."""
    from oracle.port import INT2
    rng = np.random.default_rng(61)
    allbits = np.arange(65536, dtype=np.uint16)
    by_mag = allbits[np.argsort((allbits & 0x7FFF).astype(np.int64), kind="stable")]
    blocks = by_mag.reshape(-1, 4096).copy()
    for b in blocks:
        rng.shuffle(b)
    x = np.concatenate([blocks.reshape(-1), rng.permutation(allbits), by_mag[:40000], allbits[:12345]])
    for dt_out in (UINT2, INT2):
        for scale in (2.0 / 3.0, 0.5, 0.037, 1e-3, 123.456, 1.0, 3e-5):
            for zp in (0, 1, 2, 3, -1, 5, -2):
                for mode, xi in ((NEAREST, 0.0), (STOCHASTIC, 0.0), (STOCHASTIC, 0.3), (STOCHASTIC, 0.99999994)):
                    with np.errstate(all="ignore"):
                        want = port.quantize(x, dt_out, scale, zp, mode, xi=xi, semantics=SEM_BODY)
                    got = gpu0.quantize(x, dt_out, scale, zp, mode, xi=xi)
                    bad = np.flatnonzero(got != want)
                    assert bad.size == 0, (f"dt={dt_out} scale={scale} zp={zp} mode={mode} xi={xi}: first bad byte {bad[:4]}, "
                                           f"inputs {[hex(v) for v in x[bad[0] * 4: bad[0] * 4 + 4]]} got {got[bad[0]]:#x} want {want[bad[0]]:#x}")


# ------------------------------------------------------------------------------------------------
# min/max -> (scale, zero_point)
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_in", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("dt_q", (UINT2, UINT4, UINT8), ids=("u2", "u4", "u8"))
def test_compute_quant_params_bit_equal(gpu0, dt_in, dt_q):
    rng = np.random.default_rng(4)
    isz = 4 if dt_in == F32 else 2
    cases = [(make_input(rng, n, dt_in, lo, hi), off)
             for (lo, hi), n, off in zip(((-1, 1), (0, 1), (1, 2), (-5, -1), (-1e-3, 1e3), (-3e38, 3e38), (-1, 1), (-2, 7)),
                                         (1, 7, 50000, 4097, 1 << 20, 33333, (1 << 22) + 3, 12345),
                                         (0, 0, 0, isz, 0, 3 * isz, 16, 32 - isz))]
    consts = [np.full(100, v, np.float32) for v in (42.0, 0.0, -7.5)]
    cases += [(c if dt_in == F32 else f32_to_bf16_bits(c), 0) for c in consts]
    nan_mix = rng.uniform(-1, 1, 10000).astype(np.float32)
    nan_mix[::7] = np.nan
    cases.append((nan_mix if dt_in == F32 else f32_to_bf16_bits(nan_mix), 0))
    for x, off in cases:
        want = port.compute_quant_params(x, dt_q)
        got = gpu0.compute_quant_params(x, dt_q, in_off=off)
        assert np.float32(got[0]).tobytes() == np.float32(want[0]).tobytes() and got[1] == want[1], f"n={x.size} off={off}: {got} vs {want}"


def test_quant_params_match_reference_golden(gpu0):
    g = np.load(GOLDEN)
    for key in [str(k) for k in g["__keys__"] if str(k).startswith("params/")]:
        dtq = key.split("/")[-1]
        x = g[key + "/x"]
        bits, zp = g[key + "/p"]
        s, z = gpu0.compute_quant_params(x, DTN[dtq])
        assert int(np.float32(s).view(np.uint32)) == int(bits) and z == int(zp), key


def test_known_answers(gpu0):
    """SURVEY section 8c values probed from the reference build; identity KAT of test/quant.cpp:198-217."""
    pm1 = np.array([-1, 1], np.float32)
    assert gpu0.compute_quant_params(pm1, UINT4) == (pytest.approx(0.13333334028720856, abs=0), 8)
    assert gpu0.compute_quant_params(pm1, UINT2) == (pytest.approx(0.6666666865348816, abs=0), 2)
    assert gpu0.compute_quant_params(np.array([-3, 5, 1], np.float32), UINT8) == (pytest.approx(0.0313725508749485, abs=0), 96)
    assert gpu0.compute_quant_params(np.array([1, 2], np.float32), UINT8) == (pytest.approx(0.003921568859368563, abs=0), 0)
    c = np.full(8191, 42.0, np.float32)
    s, z = gpu0.compute_quant_params(c, UINT8)
    assert (s, z) == (1.0, 127)
    q = gpu0.quantize(c, UINT8, s, z)
    y = gpu0.dequantize(q, UINT8, c.size, F32, s, z, ADD, prev=np.zeros_like(c))
    assert np.abs(y - 42.0).max() <= 1e-6


# ------------------------------------------------------------------------------------------------
# device-resident parameters (one-shot quantize): bit-identical to the two-step host-parameter path
# ------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_quantize_auto_equals_params_then_quantize(gpu, cell):
    import torch
    from gpu_util import DT, to_dev, to_host
    from piquant import RoundMode

    dt_in, dt_out = cell
    rng = np.random.default_rng(21)
    for n, lo, hi in ((1, -1, 1), (7, 0, 1), (4099, -3, 5), (250_001, -1e-3, 1e3), (3_000_000, -1, 1), (30_000_000, -2, 2)):
        x = make_input(rng, n, dt_in, lo, hi)
        want_params = port.compute_quant_params(x, dt_out)
        want_q = port.quantize(x, dt_out, *want_params, NEAREST, semantics=SEM_BODY)
        d_in = to_dev(x)
        d_out = to_dev(np.zeros(packed_bytes(dt_out, n), np.uint8))
        got_params = gpu.ctx.quantize_auto_ptr(d_in.data_ptr(), DT[dt_in], d_out.data_ptr(), DT[dt_out], n, RoundMode.NEAREST)
        assert np.float32(got_params[0]).tobytes() == np.float32(want_params[0]).tobytes() and got_params[1] == want_params[1], f"n={n}"
        assert np.array_equal(to_host(d_out, np.uint8), want_q), f"n={n}"
        # the same through the asynchronous pieces + dequantize from the device-resident parameters
        meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
        d_out2 = to_dev(np.zeros(packed_bytes(dt_out, n), np.uint8))
        gpu.ctx.compute_meta_async_ptr(d_in.data_ptr(), DT[dt_in], n, DT[dt_out], meta.data_ptr())
        gpu.ctx.quantize_meta_async_ptr(d_in.data_ptr(), DT[dt_in], d_out2.data_ptr(), DT[dt_out], n, RoundMode.NEAREST, meta.data_ptr())
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_in == F32 else f32_to_bf16_bits(prev)
        d_acc = to_dev(prev)
        from piquant import ReduceOp
        gpu.ctx.dequantize_meta_async_ptr(d_out2.data_ptr(), DT[dt_out], d_acc.data_ptr(), DT[dt_in], n, ReduceOp.ADD, meta.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(to_host(d_out2, np.uint8), want_q), f"n={n}"
        want_acc = port.dequantize(want_q, dt_out, n, dt_in, *want_params, ADD, out=prev.copy(), semantics=SEM_BODY)
        got_acc = to_host(d_acc, prev.dtype)
        assert (np.array_equal(got_acc.view(np.uint32), want_acc.view(np.uint32)) if dt_in == F32 else bf16_equal(got_acc, want_acc)), f"n={n}"
        import piquant.torch as pt
        assert pt.meta_to_host(meta) == (pytest.approx(want_params[0], abs=0), want_params[1])


def test_device_params_kernel_matches_host_arithmetic(gpu0):
    """The one-thread parameter kernel evaluates the reference's double-precision formula exactly like the host."""
    import torch
    import piquant.torch as pt
    from gpu_util import DT, to_dev

    rng = np.random.default_rng(22)
    cases = [np.array(v, np.float32) for v in ([-1, 1], [0, 1], [1, 2], [-5, -1], [42, 42], [-3, 5, 1], [-3e38, 3e38], [1e-30, 2e-30],
                                               [0, 0], [-0.0, 0.0], [1e-45, 2e-45], [-7.5, -7.5])]
    cases += [rng.uniform(-10 ** rng.uniform(-6, 6), 10 ** rng.uniform(-6, 6), int(rng.integers(2, 2000))).astype(np.float32) for _ in range(60)]
    for x in cases:
        for dq in (UINT2, UINT4, UINT8):
            meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
            d = to_dev(x)
            gpu0.ctx.compute_meta_async_ptr(d.data_ptr(), DT[F32], x.size, DT[dq], meta.data_ptr())
            s, z = pt.meta_to_host(meta)
            ws, wz = port.compute_quant_params(x, dq)
            assert np.float32(s).tobytes() == np.float32(ws).tobytes() and z == wz, (x[:4], dq, (s, z), (ws, wz))


# ------------------------------------------------------------------------------------------------
# fuzz: random cell / size / alignment / zero point and DEGENERATE scales (0, inf, NaN, negative, denormal)
# ------------------------------------------------------------------------------------------------

_SCALES = (1.0, 0.5, 2.0 / 255, 1e-3, 123.456, 1e-30, 1e30, 1e-45, 3.4e38, 0.0, -0.0, float("inf"), -1.0, -0.037, float("nan"))


def test_fuzz_quantize_against_oracle(gpu):
    rng = np.random.default_rng(0xF022)
    for it in range(150):
        dt_in, dt_out = QUANT_CELLS[int(rng.integers(len(QUANT_CELLS)))]
        n = int(rng.choice([1, 2, 3, 5, 31, 257, 1000, 4099, 65_537, 300_001]))
        scale = float(np.float32(_SCALES[int(rng.integers(len(_SCALES)))]))
        zp = int(rng.choice([0, 1, 3, 8, 128, 255, -5, 1000, -1000, 2**31 - 1, -2**31, 2**33 + 7]))
        mode = int(rng.integers(2))
        xi = float(np.float32(rng.uniform(0, 0.999)))
        x = make_input(rng, n, dt_in, -float(rng.choice([1, 3, 300, 1e6])), float(rng.choice([1, 3, 300, 1e6])))
        if n > 40 and rng.random() < 0.5:
            sp = special_values(1.0 if not np.isfinite(scale) or scale == 0 else abs(scale))
            sp = sp if dt_in == F32 else f32_to_bf16_bits(sp)
            x[3:3 + sp.size] = sp[: max(0, min(sp.size, n - 3))]
        isz = 4 if dt_in == F32 else 2
        in_off, out_off = int(rng.choice([0, isz, 16, 32 + isz])), int(rng.choice([0, 1, 5, 16]))
        with np.errstate(all="ignore"):
            want = port.quantize(x, dt_out, scale, zp, mode, xi=xi, semantics=SEM_BODY)
        got = gpu.quantize(x, dt_out, scale, zp, mode, xi=xi, in_off=in_off, out_off=out_off)
        assert np.array_equal(got, want), f"it={it} cell=({dt_in},{dt_out}) n={n} scale={scale} zp={zp} mode={mode} xi={xi} offs=({in_off},{out_off})"


def test_fuzz_dequantize_against_oracle(gpu):
    rng = np.random.default_rng(0xF023)
    for it in range(150):
        dt_in, dt_out, op = DEQUANT_CELLS[int(rng.integers(len(DEQUANT_CELLS)))]
        n = int(rng.choice([1, 2, 3, 5, 31, 257, 1000, 4099, 65_537, 300_001]))
        scale = float(np.float32(_SCALES[int(rng.integers(len(_SCALES)))]))
        zp = int(rng.choice([0, 1, 3, 8, 128, 255, -5, 1000, 2**22, 2**22 + 1, -2**22 - 1, 2**31 - 1, -2**31, 2**33 + 7]))
        q = rng.integers(0, 256, packed_bytes(dt_in, n)).astype(np.uint8)
        prev = rng.uniform(-100, 100, n).astype(np.float32)
        prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
        osz = 4 if dt_out == F32 else 2
        in_off, out_off = int(rng.choice([0, 1, 4, 16])), int(rng.choice([0, osz, 16, 32 + osz]))
        with np.errstate(all="ignore"):
            want = port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=prev.copy(), semantics=SEM_BODY)
        got = gpu.dequantize(q, dt_in, n, dt_out, scale, zp, op, prev=prev, in_off=in_off, out_off=out_off)
        ok = np.array_equal(got.view(np.uint32), want.view(np.uint32)) if dt_out == F32 else bf16_equal(got, want)
        if dt_out == F32 and not ok:            # f32 NaN payloads: any NaN matches any NaN
            gn, wn = np.isnan(got), np.isnan(want)
            ok = np.array_equal(gn, wn) and np.array_equal(got[~gn].view(np.uint32), want[~wn].view(np.uint32))
        assert ok, f"it={it} cell=({dt_in},{dt_out},{op}) n={n} scale={scale} zp={zp} offs=({in_off},{out_off})"
