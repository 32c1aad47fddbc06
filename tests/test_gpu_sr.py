"""GPU parity for per-element stochastic rounding (extension: PIQUANT_CUDA_STOCHASTIC_PER_ELEMENT, piquant_cuda.h):
the Philox4x32-10 + floor(p + u) kernels through the C ABI against the oracle (orc_quantize_sr), bit for bit -- every
cell, ragged sizes, misaligned buffers (byte-granular kernel), special values, big zero points, signed dtypes, the chunked
host-pointer path (the random stream is a function of the element index), device-resident parameters, the key API."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from helpers import QUANT_CELLS, cell_id, make_input, special_values
from oracle import port
from oracle.port import BF16, BITS, F32, INT4, INT8, UINT4, UINT8, f32_to_bf16_bits, packed_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu0():
    from gpu_util import Gpu
    return Gpu(variant=0)


@pytest.mark.parametrize("cell", QUANT_CELLS, ids=cell_id)
def test_sr_quantize_bit_exact(gpu0, cell):
    dt_in, dt_out = cell
    bits = BITS[dt_out]
    isz = 4 if dt_in == F32 else 2
    rng = np.random.default_rng(51)
    for n in (1, 2, 3, 7, 8, 9, 63, 64, 65, 1000, 4097, 12345, 1 << 18, (1 << 20) + 3):
        for scale, zp in ((2.0 / ((1 << bits) - 1), (1 << bits) // 2), (0.037, 0), (0.5, -3), (1.0, 2**31 - 1), (0.25, 2**40 + 7)):
            x = make_input(rng, n, dt_in, -4.0, 4.0)
            if n > 100:
                sp = special_values(scale)
                x[10:10 + sp.size] = sp if dt_in == F32 else f32_to_bf16_bits(sp)
            key = int(rng.integers(0, 2**63))
            # (isz, 3): output not 16-byte aligned by 3 -> ragged head not a multiple of 8 elements -> byte-granular kernel
            for in_off, out_off in ((0, 0),) if n not in (12345, 1 << 18) else ((0, 0), (isz, 3), (0, 8), (32, 16)):
                with np.errstate(all="ignore"):
                    want = port.quantize_sr(x, dt_out, scale, zp, key)
                got = gpu0.quantize_sr(x, dt_out, scale, zp, key, in_off=in_off, out_off=out_off)
                assert np.array_equal(got, want), f"n={n} scale={scale} zp={zp} offs=({in_off},{out_off}): {np.flatnonzero(got != want)[:8]}"


@pytest.mark.parametrize("dt_out", (INT8, INT4), ids=("i8", "i4"))
def test_sr_signed_bit_exact(gpu0, dt_out):
    rng = np.random.default_rng(52)
    for dt_in in (F32, BF16):
        x = make_input(rng, 100_001, dt_in, -1.0, 1.0)
        s, z = port.compute_quant_params(x, dt_out)
        assert np.array_equal(gpu0.quantize_sr(x, dt_out, s, z, 1234), port.quantize_sr(x, dt_out, s, z, 1234))


def test_sr_host_pointer_path_continues_the_random_stream(gpu0):
    """A host tensor larger than one pipeline chunk (8 Mi elements) is quantized chunk by chunk; element i must still use
    the random bits of index i of the whole tensor."""
    from gpu_util import DT
    from piquant import RoundMode
    rng = np.random.default_rng(53)
    n = (8 << 20) * 2 + 12345
    x = rng.uniform(-1, 1, n).astype(np.float32)
    out = np.zeros(n, np.uint8)
    gpu0.ctx.set_sr_key(99)
    try:
        gpu0.ctx.quantize_ptr(x.ctypes.data, DT[F32], out.ctypes.data, DT[UINT8], n, 2 / 255, 128, RoundMode.STOCHASTIC_PER_ELEMENT)
    finally:
        gpu0.ctx.set_sr_key(None)
    assert np.array_equal(out, port.quantize_sr(x, UINT8, 2 / 255, 128, 99))


def test_sr_key_api_and_statistics(gpu0):
    import torch
    from gpu_util import DT
    from piquant import RoundMode
    ctx = gpu0.ctx
    n = 4_000_000
    x = torch.full((n,), 0.3 * 0.01, device="cuda")
    q = torch.empty(n, dtype=torch.uint8, device="cuda")

    def run():
        ctx.quantize_ptr(x.data_ptr(), DT[F32], q.data_ptr(), DT[UINT8], n, 0.01, 128, RoundMode.STOCHASTIC_PER_ELEMENT)
        torch.cuda.synchronize()
        return q.clone()

    ctx.seed(7)
    a = run(); ka = ctx.last_sr_key
    b = run(); kb = ctx.last_sr_key
    assert ka != kb and not torch.equal(a, b)                 # a fresh key per call
    ctx.seed(7)
    assert torch.equal(run(), a) and ctx.last_sr_key == ka    # replayable after seeding
    ctx.set_sr_key(ka)
    assert torch.equal(run(), a) and torch.equal(run(), a)    # fixed key
    ctx.set_sr_key(None)
    p = float(np.float32(0.3 * 0.01) * (np.float32(1) / np.float32(0.01)))
    m = a.double().mean().item()
    assert abs(m - (128 + p)) < 4 * np.sqrt(0.21 / n), m      # unbiased: 4 sigma of a Bernoulli mean
    # the reference's per-call mode rounds all elements the same way
    ctx.set_stochastic_threshold(0.6)
    ctx.quantize_ptr(x.data_ptr(), DT[F32], q.data_ptr(), DT[UINT8], n, 0.01, 128, RoundMode.STOCHASTIC)
    torch.cuda.synchronize()
    ctx.set_stochastic_threshold(None)
    assert q.min().item() == q.max().item()


@pytest.mark.parametrize("dt_in,dt_out", list(itertools.product((F32, BF16), (UINT8, UINT4))), ids=lambda v: str(v))
def test_sr_with_device_resident_parameters(gpu0, dt_in, dt_out):
    import torch
    from gpu_util import DT
    from piquant import RoundMode
    ctx = gpu0.ctx
    rng = np.random.default_rng(54)
    n = 300_001
    x = make_input(rng, n, dt_in, -1.5, 0.75)
    scale, zp = port.compute_quant_params(x, dt_out)
    d_in = torch.from_numpy(x.view(np.uint8)).cuda()
    d_out = torch.zeros(packed_bytes(dt_out, n), dtype=torch.uint8, device="cuda")
    ctx.set_sr_key(4242)
    try:
        s2, z2 = ctx.quantize_auto_ptr(d_in.data_ptr(), DT[dt_in], d_out.data_ptr(), DT[dt_out], n, RoundMode.STOCHASTIC_PER_ELEMENT)
    finally:
        ctx.set_sr_key(None)
    assert (s2, z2) == (scale, zp)
    assert np.array_equal(d_out.cpu().numpy(), port.quantize_sr(x, dt_out, scale, zp, 4242))


def test_sr_torch_surface(gpu0):
    import torch
    import piquant.torch as pt
    x = torch.empty(1_000_000, device="cuda").uniform_(-1, 1)
    scale, zp = pt.compute_quant_params(x, dtype=torch.quint8)
    q = pt.quantize(x, scale=scale, zero_point=zp, dtype=torch.quint8, round_mode="stochastic_per_element")
    y = pt.dequantize(q, scale=scale, zero_point=zp, dtype=torch.float32)
    err = (y - x)
    assert err.abs().max().item() <= scale * (1 + 1e-5) + 1e-7          # never more than one step away
    assert abs(err.mean().item()) < 5 * scale * 0.41 / 1000              # and unbiased (sigma of U-shaped error <= 0.41 step, n = 1e6)
    r = pt.requantize(x, scale=scale, zero_point=zp, dtype=torch.quint8, round_mode="stochastic_per_element")
    assert (r - x).abs().max().item() <= scale * (1 + 1e-5) + 1e-7 and abs((r - x).mean().item()) < 5 * scale * 0.41 / 1000


@pytest.mark.parametrize("dt_io", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("op", (0, 1), ids=("set", "add"))
def test_sr_requantize_bit_exact(gpu0, dt_io, op):
    """Fused quantize -> dequantize with per-element stochastic rounding: the float-domain fast path (floor by two round-down
    adds, no conversion), the exact fallback (special values), ragged sizes and the one-element-per-thread kernel (misaligned
    buffers) against orc_requantize_sr, bit for bit."""
    from oracle.port import INT4, UINT2
    rng = np.random.default_rng(55)
    isz = 4 if dt_io == F32 else 2
    for dt_q in (UINT8, UINT4, UINT2, INT8, INT4):
        bits = BITS[dt_q]
        for n in (1, 7, 8, 9, 4099, 100_003, 1 << 18):
            for scale, zp in ((2.0 / ((1 << bits) - 1), 1), (0.037, 0), (0.5, 300), (1.0, 2**40 + 7)):
                x = make_input(rng, n, dt_io, -3.0, 3.0)
                if n > 100:
                    sp = special_values(scale)
                    x[10:10 + sp.size] = sp if dt_io == F32 else f32_to_bf16_bits(sp)
                prev = rng.uniform(-1, 1, n).astype(np.float32)
                prev = prev if dt_io == F32 else f32_to_bf16_bits(prev)
                key = int(rng.integers(0, 2**63))
                for in_off, out_off in ((0, 0),) if n != 100_003 else ((0, 0), (isz, isz), (isz, 2 * isz), (0, 16)):
                    with np.errstate(all="ignore"):
                        want = port.requantize_sr(x, dt_q, scale, zp, key, op=op, out=prev.copy())
                    got = gpu0.requantize_sr(x, dt_q, scale, zp, key, op=op, prev=prev, in_off=in_off, out_off=out_off)
                    what = f"dt_q={dt_q} n={n} scale={scale} zp={zp} offs=({in_off},{out_off})"
                    if dt_io == F32:
                        bad = np.flatnonzero(got.view(np.uint32) != want.view(np.uint32))
                        assert bad.size == 0, f"{what}: x={x[bad[:4]]} got={got[bad[:4]]} want={want[bad[:4]]}"
                    else:
                        an, bn = (got & 0x7FFF) > 0x7F80, (want & 0x7FFF) > 0x7F80
                        assert np.array_equal(an, bn) and np.array_equal(got[~an], want[~bn]), what
