"""Maximum sizes: tensors with MORE than 2^32 elements (a B200 holds them: 4.3 G f32 = 17 GB).

The reference indexes with size_t / int64 throughout (kernels.inl:108-196, piquant.cpp:277-381); so must the
kernels.  An element-wise oracle run at this size is out of reach, so each test checks
  (i)  whole-tensor call == the same call on 2^30-element slices (each far below any 32-bit limit), bit for bit;
  (ii) windows straddling element 2^31, element 2^32 and the ragged end against the oracle, bit for bit;
  (iii) min/max finds extremes planted beyond element 2^32.
"""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

N = (1 << 32) + (1 << 20) + 37          # ragged: not a multiple of any vector or pack width
SLICE = 1 << 30
WIN = 1 << 16


@pytest.fixture(scope="module")
def pt():
    import piquant.torch as pt
    return pt


@pytest.fixture(scope="module")
def x_huge():
    if torch.cuda.mem_get_info()[0] < 80 * 2**30:
        pytest.skip("needs 80 GiB of free device memory")
    x = torch.empty(N, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(32)
    for lo in range(0, N, SLICE):
        x[lo:lo + SLICE].uniform_(-1, 1, generator=g)
    yield x
    del x
    torch.cuda.empty_cache()


def windows():
    return (0, (1 << 31) - WIN // 2, (1 << 32) - WIN // 2, N - WIN - 37)


def raw_bytes(q: "torch.Tensor", nbytes: int) -> "torch.Tensor":
    return torch.empty(0, dtype=torch.uint8, device=q.device).set_(q.untyped_storage())[:nbytes]


def test_quantize_f32_u8_beyond_2_32(pt, x_huge):
    from oracle import port
    x = x_huge
    s, z = 2.0 / 255.0, 128
    q = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8)
    assert q.numel() == N
    for lo in range(0, N, SLICE):
        part = pt.quantize(x[lo:lo + SLICE], scale=s, zero_point=z, dtype=torch.uint8)
        assert torch.equal(q[lo:lo + SLICE], part), f"slice at {lo}"
    for lo in windows():
        hi = min(N, lo + WIN + 37)
        assert np.array_equal(q[lo:hi].cpu().numpy(), port.quantize(x[lo:hi].cpu().numpy(), port.UINT8, s, z)), lo
    # the TMA-ring kernels (variant 2) index the same way
    import piquant
    ctx_tma = piquant.Context()
    ctx_tma.set_kernel_variant(2)
    assert torch.equal(pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8, ctx=ctx_tma), q)
    y_direct = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32)
    y_tma = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32, ctx=ctx_tma)
    assert torch.equal(y_direct, y_tma)
    del y_direct, y_tma
    # stochastic with a fixed threshold against the oracle windows
    ctx = piquant.Context()
    ctx.set_stochastic_threshold(0.3125)
    qs = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8, round_mode="stochastic", ctx=ctx)
    for lo in windows():
        hi = min(N, lo + WIN + 37)
        want = port.quantize(x[lo:hi].cpu().numpy(), port.UINT8, s, z, mode=port.STOCHASTIC, xi=0.3125)
        assert np.array_equal(qs[lo:hi].cpu().numpy(), want), lo
    del qs

    # dequantize ADD back into a 17 GB accumulator, then SET; both against slices and the oracle windows
    acc = torch.full((N,), 0.25, device="cuda")
    pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32, reduce_op="add", out=acc)
    for lo in range(0, N, SLICE):
        part = torch.full((min(SLICE, N - lo),), 0.25, device="cuda")
        pt.dequantize(q[lo:lo + SLICE], scale=s, zero_point=z, dtype=torch.float32, reduce_op="add", out=part)
        assert torch.equal(acc[lo:lo + SLICE], part), f"slice at {lo}"
        del part
    for lo in windows():
        hi = min(N, lo + WIN + 37)
        prev = np.full(hi - lo, 0.25, dtype=np.float32)
        want = port.dequantize(q[lo:hi].cpu().numpy(), port.UINT8, hi - lo, port.F32, s, z, op=port.ADD, out=prev)
        assert np.array_equal(acc[lo:hi].cpu().numpy().view(np.uint32), want.view(np.uint32)), lo
    del acc

    # fused requantize of the whole tensor == dequantize(quantize(x)) on the windows
    y = pt.requantize(x, scale=s, zero_point=z, dtype=torch.uint8)
    for lo in windows():
        hi = min(N, lo + WIN + 37)
        want = port.dequantize(q[lo:hi].cpu().numpy(), port.UINT8, hi - lo, port.F32, s, z)
        assert np.array_equal(y[lo:hi].cpu().numpy().view(np.uint32), want.view(np.uint32)), lo


def test_minmax_beyond_2_32(pt, x_huge):
    import piquant
    x = x_huge
    old = (x[(1 << 32) + 5].item(), x[N - 1].item())
    x[(1 << 32) + 5] = 7.0
    x[N - 1] = -9.0
    try:
        ctx = piquant.Context()
        for dt, cdt in ((torch.quint8, piquant.DataType.UINT8), (torch.quint2x4, piquant.DataType.UINT2)):
            s, z = pt.compute_quant_params(x, dtype=dt, ctx=ctx)
            want = ctx.params_from_minmax(-9.0, 7.0, cdt)
            assert (np.float32(s), z) == (np.float32(want[0]), want[1])
    finally:
        x[(1 << 32) + 5], x[N - 1] = old


def test_bf16_u2_beyond_2_32(pt, x_huge):
    """Packed sub-byte cells: 4 elements per byte, so element 2^32 sits at byte 2^30; the last byte is partial."""
    from oracle import port
    xb = x_huge.bfloat16()
    s, z = 2.0 / 3.0, 2
    nbytes = (N + 3) // 4
    q = pt.quantize(xb, scale=s, zero_point=z, dtype=torch.quint2x4)
    raw = raw_bytes(q, nbytes)
    for lo in range(0, N, SLICE):
        part = pt.quantize(xb[lo:lo + SLICE], scale=s, zero_point=z, dtype=torch.quint2x4)
        nb = (min(SLICE, N - lo) + 3) // 4
        assert torch.equal(raw[lo // 4:lo // 4 + nb], raw_bytes(part, nb)), f"slice at {lo}"
    y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.bfloat16)
    assert y.numel() == N
    for lo in windows():
        lo -= lo % 4
        hi = N if lo + WIN + 64 >= N else lo + WIN
        xs = xb[lo:hi].view(torch.int16).cpu().numpy().view(np.uint16)
        want_q = port.quantize(xs, port.UINT2, s, z)
        got_q = raw[lo // 4:lo // 4 + want_q.size].cpu().numpy()
        assert np.array_equal(got_q, want_q), lo
        want_y = port.dequantize(want_q, port.UINT2, hi - lo, port.BF16, s, z)
        assert np.array_equal(y[lo:hi].view(torch.int16).cpu().numpy().view(np.uint16), want_y), lo
