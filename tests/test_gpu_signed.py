"""GPU parity for the signed extension dtypes (INT2 / INT4 / INT8, include/piquant_cuda.h): the sm_100a kernels through
the C ABI against the oracle's definition (offset-binary view of the pinned unsigned functions, piquant_oracle.h), both
kernel families (direct and TMA ring), any alignment, special values, device-resident parameters and the torch surface.
Bit-exact bar as for the unsigned types."""
from __future__ import annotations

import itertools

import numpy as np
import pytest

from helpers import make_input, special_values
from oracle import port
from oracle.port import (ADD, BF16, BITS, F32, INT2, INT4, INT8, NEAREST, SEM_BODY, SET, SIGNED, STOCHASTIC, f32_to_bf16_bits,
                         packed_bytes)

pytestmark = pytest.mark.gpu

SIGNED_TYPES = (INT2, INT4, INT8)
NAMES = {F32: "f32", BF16: "bf16", INT2: "i2", INT4: "i4", INT8: "i8"}
CELLS = list(itertools.product((F32, BF16), SIGNED_TYPES))


def cid(c):
    return "-".join(NAMES[v] for v in c)


@pytest.fixture(scope="module", params=(1, 2), ids=("direct", "tma"))
def gpu(request):
    from gpu_util import Gpu
    return Gpu(variant=request.param)


@pytest.fixture(scope="module")
def gpu0():
    from gpu_util import Gpu
    return Gpu(variant=0)


def bf16_equal(a, b):
    an = (a & 0x7FFF) > 0x7F80
    bn = (b & 0x7FFF) > 0x7F80
    return bool(np.array_equal(an, bn) and np.array_equal(a[~an], b[~bn]))


@pytest.mark.parametrize("cell", CELLS, ids=cid)
@pytest.mark.parametrize("mode", (NEAREST, STOCHASTIC), ids=("nearest", "stochastic"))
def test_signed_quantize_bit_exact(gpu, cell, mode):
    dt_in, dt_out = cell
    bits = BITS[dt_out]
    rng = np.random.default_rng(31)
    isz = 4 if dt_in == F32 else 2
    for n in (1, 2, 3, 5, 17, 63, 64, 65, 4097, 12345, 1 << 18):
        for scale, zp in ((2.0 / ((1 << bits) - 1), -1), (0.037, 0), (0.5, (1 << (bits - 1)) - 1), (1.0, -(1 << (bits - 1))),
                          (0.25, 2**31 - 1), (0.25, -2**40 - 5)):
            x = make_input(rng, n, dt_in, -4.0, 4.0)
            if n > 100:
                sp = special_values(scale)
                x[10:10 + sp.size] = sp if dt_in == F32 else f32_to_bf16_bits(sp)
            in_off, out_off = (0, 0) if n != 12345 else (isz, 3)
            with np.errstate(all="ignore"):
                want = port.quantize(x, dt_out, scale, zp, mode, xi=0.3, semantics=SEM_BODY)
            got = gpu.quantize(x, dt_out, scale, zp, mode, xi=0.3, in_off=in_off, out_off=out_off)
            assert np.array_equal(got, want), f"n={n} scale={scale} zp={zp}: {np.flatnonzero(got != want)[:8]}"


@pytest.mark.parametrize("dt_in", SIGNED_TYPES, ids=lambda d: NAMES[d])
@pytest.mark.parametrize("dt_out", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("op", (SET, ADD), ids=("set", "add"))
def test_signed_dequantize_bit_exact(gpu, dt_in, dt_out, op):
    rng = np.random.default_rng(32)
    osz = 4 if dt_out == F32 else 2
    for n in (1, 3, 7, 64, 65, 1001, 4099, 100_003, 1 << 18):
        for scale, zp in ((0.01, -1), (0.5, 3), (2.0 / 255, -128), (1.0, 2**22 + 9), (1.0, -2**35)):
            q = rng.integers(0, 256, packed_bytes(dt_in, n), dtype=np.uint8)
            prev = rng.uniform(-1, 1, n).astype(np.float32)
            prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
            in_off, out_off = (0, 0) if n != 100_003 else (1, osz)
            want = port.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=prev.copy())
            got = gpu.dequantize(q, dt_in, n, dt_out, scale, zp, op, prev=prev, in_off=in_off, out_off=out_off)
            if dt_out == F32:
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"n={n} scale={scale} zp={zp}"
            else:
                assert bf16_equal(got, want), f"n={n} scale={scale} zp={zp}"


@pytest.mark.parametrize("dt_io,dt_q,op,mode", list(itertools.product((F32, BF16), SIGNED_TYPES, (SET, ADD), (NEAREST, STOCHASTIC))),
                         ids=lambda v: str(v))
def test_signed_requantize_bit_exact(gpu0, dt_io, dt_q, op, mode):
    rng = np.random.default_rng(33)
    for n in (1, 7, 4099, 1 << 16):
        x = make_input(rng, n, dt_io, -2.0, 2.0)
        scale, zp = port.compute_quant_params(x, dt_q)
        prev = rng.uniform(-1, 1, n).astype(np.float32)
        prev = prev if dt_io == F32 else f32_to_bf16_bits(prev)
        want = port.requantize(x, dt_q, scale, zp, mode, 0.4, op, out=prev.copy(), fma_add=True)
        got = gpu0.requantize(x, dt_q, scale, zp, mode, 0.4, op, prev=prev)
        if dt_io == F32:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"n={n}"
        else:
            assert bf16_equal(got, want), f"n={n}"


@pytest.mark.parametrize("dt_in", (F32, BF16), ids=("f32", "bf16"))
@pytest.mark.parametrize("dt_q", SIGNED_TYPES, ids=lambda d: NAMES[d])
def test_signed_compute_quant_params_bit_equal(gpu0, dt_in, dt_q):
    rng = np.random.default_rng(34)
    for n in (1, 2, 1000, 100_001):
        for lo, hi in ((-1.0, 1.0), (0.0, 1.0), (-5.0, -1.0), (-1e-3, 2e-3), (-300.0, 700.0)):
            x = make_input(rng, n, dt_in, lo, hi)
            assert gpu0.compute_quant_params(x, dt_q) == port.compute_quant_params(x, dt_q), (n, lo, hi)
    x = np.full(100, 0.25, np.float32)
    x = x if dt_in == F32 else f32_to_bf16_bits(x)
    assert gpu0.compute_quant_params(x, dt_q) == (1.0, -1)


@pytest.mark.parametrize("cell", CELLS, ids=cid)
def test_signed_quantize_auto_and_device_params(gpu0, cell):
    """min/max -> device parameter block -> quantize / dequantize without a host round trip, signed target."""
    import torch
    from gpu_util import DT
    from piquant import ReduceOp, RoundMode
    dt_in, dt_out = cell
    rng = np.random.default_rng(35)
    ctx = gpu0.ctx
    for n in (1000, 300_001):
        x = make_input(rng, n, dt_in, -1.5, 0.75)
        scale, zp = port.compute_quant_params(x, dt_out)
        want = port.quantize(x, dt_out, scale, zp, NEAREST, semantics=SEM_BODY)
        d_in = torch.from_numpy(x.view(np.uint8)).cuda()
        d_out = torch.zeros(packed_bytes(dt_out, n), dtype=torch.uint8, device="cuda")
        s2, z2 = ctx.quantize_auto_ptr(d_in.data_ptr(), DT[dt_in], d_out.data_ptr(), DT[dt_out], n, RoundMode.NEAREST)
        assert (s2, z2) == (scale, zp)
        assert np.array_equal(d_out.cpu().numpy(), want)
        # explicit meta path: compute_meta -> quantize_meta -> dequantize_meta
        meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
        d_out.zero_()
        ctx.compute_meta_async_ptr(d_in.data_ptr(), DT[dt_in], n, DT[dt_out], meta.data_ptr())
        ctx.quantize_meta_async_ptr(d_in.data_ptr(), DT[dt_in], d_out.data_ptr(), DT[dt_out], n, RoundMode.NEAREST, meta.data_ptr())
        back = torch.zeros(n, dtype=torch.float32, device="cuda")
        ctx.dequantize_meta_async_ptr(d_out.data_ptr(), DT[dt_out], back.data_ptr(), DT[F32], n, ReduceOp.SET, meta.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), want)
        want_back = port.dequantize(want, dt_out, n, F32, scale, zp, SET)
        assert np.array_equal(back.cpu().numpy().view(np.uint32), want_back.view(np.uint32))
        m = meta.cpu().numpy()
        assert m[:4].view(np.float32)[0] == np.float32(scale) and m[8:16].view(np.int64)[0] == zp


def test_signed_is_offset_binary_of_unsigned_on_gpu(gpu0):
    """int8 bytes == uint8 bytes of zero point + 128 with the top bit flipped, on 2^24 elements (all vector paths)."""
    rng = np.random.default_rng(36)
    n = (1 << 24) + 5
    x = rng.uniform(-1, 1, n).astype(np.float32)
    for sdt, sign in ((INT8, 0x80), (INT4, 0x88), (INT2, 0xAA)):
        bits = BITS[sdt]
        s, z = port.compute_quant_params(x, sdt)
        a = gpu0.quantize(x, sdt, s, z)
        b = gpu0.quantize(x, SIGNED[sdt], s, z + (1 << (bits - 1)))
        mask = np.full(a.size, sign, np.uint8)
        if (n * bits) % 8:
            mask[-1] &= (1 << ((n * bits) % 8)) - 1          # fields of elements that do not exist stay zero
        assert np.array_equal(a, b ^ mask)


def test_torch_surface_int8(gpu0):
    import torch
    import piquant.torch as pt
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.empty(100_000, device="cuda").uniform_(-1, 1, generator=g)
    scale, zp = pt.compute_quant_params(x, dtype=torch.int8)
    q = pt.quantize(x, scale=scale, zero_point=zp, dtype=torch.int8)
    assert q.dtype == torch.int8 and q.shape == x.shape and q.device == x.device
    want = torch.clamp(torch.round(x / scale) + zp, -128, 127).to(torch.int8)
    assert (q.to(torch.int32) - want.to(torch.int32)).abs().max().item() <= 1       # torch.round is half-to-even
    assert (q != want).float().mean().item() < 1e-3
    y = pt.dequantize(q, scale=scale, zero_point=zp, dtype=torch.float32)
    assert (y - x).abs().max().item() <= 0.5 * scale * (1 + 1e-5) + 1e-7
    # against torch's own per-tensor affine qint8
    tq = torch.quantize_per_tensor(x.cpu(), scale, zp, torch.qint8)
    assert (tq.int_repr().to(torch.int32) - q.cpu().to(torch.int32)).abs().max().item() <= 1
    assert torch.allclose(tq.dequantize(), y.cpu(), atol=scale + 1e-6)
