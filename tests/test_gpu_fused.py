"""Round-2 additions on the GPU, each against the oracle AND against the unfused path it replaces:

* the reduction tail that also evaluates the parameter arithmetic (one launch instead of min/max + params kernel);
* dequantize-ADD fused with min/max + parameters of the sums (the receiver side of a ring reduce-scatter hop);
* dequantize-SET that also forwards the packed bytes (the all-gather hop);
* many tensors in one launch (``piquant_cuda_quantize_batch``);
* ``*_on_stream`` entry points: explicit device + stream per call, threads sharing one context without a lock,
  per-stream reduction scratch (the round-1 advisor finding);
* pageable host tensors through the bounce-buffer pipeline;
* parameter blocks flagged as failed stop the kernels that consume them.
"""
from __future__ import annotations

import struct
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from helpers import QUANT_DTYPES, FLOAT_DTYPES, DT_NAME, as_f32, make_input  # noqa: E402
from oracle import port  # noqa: E402
from oracle.port import ADD, BF16, F32, NEAREST, SEM_BODY, SET, STOCHASTIC, UINT2, UINT4, UINT8, packed_bytes  # noqa: E402

SIZES = (1, 63, 64, 4097, 65_536, 1_000_003)


def _ctx():
    import piquant
    return piquant.Context()


def _site():
    return torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream


def _dev_bytes(a: np.ndarray, off: int = 0) -> "torch.Tensor":
    """raw bytes of `a` in CUDA memory, `off` bytes past a 256-byte aligned address"""
    raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    buf = torch.zeros(raw.size + off + 256, dtype=torch.uint8, device="cuda")
    t = buf[off:off + raw.size]
    if raw.size:
        t.copy_(torch.from_numpy(raw))
    return t


def _meta_tuple(meta: "torch.Tensor"):
    raw = meta.cpu().numpy().tobytes()
    scale_bits, error, zp = struct.unpack_from("<Iiq", raw, 0)
    return scale_bits, error, zp


def _f32_bits(v: float) -> int:
    return struct.unpack("<I", struct.pack("<f", v))[0]


# -------------------------------------------------------------------------------------------------------------
# reduction tail with folded parameter arithmetic
# -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_in", FLOAT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("dt_q", QUANT_DTYPES, ids=lambda d: DT_NAME[d])
def test_compute_meta_one_launch_matches_oracle_params(dt_in, dt_q):
    from gpu_util import DT
    ctx = _ctx()
    rng = np.random.default_rng(5)
    dev, st = _site()
    for n in SIZES:
        for lo, hi in ((-1.0, 1.0), (0.25, 7.0), (-300.0, -2.0), (3.0, 3.0)):
            x = make_input(rng, n, dt_in, lo, hi)
            d_x = _dev_bytes(x)
            meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
            before = ctx.kernel_launches
            ctx.compute_meta_on_stream(d_x.data_ptr(), DT[dt_in], n, DT[dt_q], meta.data_ptr(), 0, dev, st)
            assert ctx.kernel_launches - before == 1, "min/max and the parameter arithmetic must be ONE launch"
            want = port.compute_quant_params(x, dt_q)
            assert _meta_tuple(meta) == (_f32_bits(want[0]), 0, want[1]), (n, lo, hi)


# -------------------------------------------------------------------------------------------------------------
# dequantize-ADD + min/max + parameters (ring reduce-scatter hop)
# -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_out", FLOAT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("dt_q", QUANT_DTYPES, ids=lambda d: DT_NAME[d])
def test_dequantize_add_minmax_equals_the_two_passes(dt_q, dt_out):
    from gpu_util import DT
    ctx = _ctx()
    rng = np.random.default_rng(11)
    dev, st = _site()
    odt = np.float32 if dt_out == F32 else np.uint16
    for n in SIZES + (262_144 * 3 + 5,):
        for in_off, out_off in ((0, 0), (0, 16), (4, 0), (1, 2 if dt_out == BF16 else 4)):
            src = make_input(rng, n, F32, -2.0, 3.0)
            s_in, z_in = port.compute_quant_params(src, dt_q)
            q = port.quantize(src, dt_q, s_in, z_in, NEAREST, semantics=SEM_BODY)
            acc = make_input(rng, n, dt_out, -1.0, 1.0)
            want_out = port.dequantize(q, dt_q, n, dt_out, s_in, z_in, ADD, out=acc.copy(), semantics=SEM_BODY)
            for dt_next in (UINT8, UINT4):
                want_params = port.compute_quant_params(want_out, dt_next)
                d_q = _dev_bytes(q, in_off)
                d_acc = _dev_bytes(acc, out_off)
                meta_in = torch.zeros(64, dtype=torch.uint8, device="cuda")
                # parameters of the incoming chunk as a device block: produce it from a 2-element tensor with the same range
                d_src = _dev_bytes(src)
                ctx.compute_meta_on_stream(d_src.data_ptr(), DT[F32], n, DT[dt_q], meta_in.data_ptr(), 0, dev, st)
                nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
                nxt2 = torch.zeros(64, dtype=torch.uint8, device="cuda")
                before = ctx.kernel_launches
                ctx.dequantize_add_minmax_on_stream(d_q.data_ptr(), DT[dt_q], d_acc.data_ptr(), DT[dt_out], n, meta_in.data_ptr(), DT[dt_next],
                                                    nxt.data_ptr(), nxt2.data_ptr(), dev, st)
                launched = ctx.kernel_launches - before
                got_out = d_acc.cpu().numpy().view(odt)
                assert np.array_equal(got_out.view(np.uint8), want_out.view(np.uint8)), (n, in_off, out_off)
                assert _meta_tuple(nxt) == (_f32_bits(want_params[0]), 0, want_params[1]), (n, in_off, out_off, dt_next)
                assert torch.equal(nxt, nxt2)
                vector_aligned = n >= 64 and (in_off % 4 == 0) and (out_off % 16 == 0)
                if vector_aligned and in_off == 0 and out_off == 0:
                    assert launched == 1, "aligned buffers must take the ONE fused kernel"


@pytest.mark.parametrize("dt_out", FLOAT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("dt_q", QUANT_DTYPES, ids=lambda d: DT_NAME[d])
def test_dequantize_sum_minmax_equals_successive_add_calls(dt_q, dt_out):
    """``piquant_cuda_dequantize_sum_minmax_on_stream``: up to 8 packed sources folded into the accumulator in ONE pass ==
    that many piquant_dequantize(ADD) calls of the reference, in order, and the parameters of the sums."""
    from gpu_util import DT
    ctx = _ctx()
    rng = np.random.default_rng(17)
    dev, st = _site()
    odt = np.float32 if dt_out == F32 else np.uint16
    for n, n_src in ((1, 1), (63, 2), (64, 3), (4097, 7), (65_536, 8), (1_000_003, 7), (262_144 * 3 + 5, 5)):
        for in_off, out_off in ((0, 0), (0, 16), (4, 0), (1, 2 if dt_out == BF16 else 4)):
            acc = make_input(rng, n, dt_out, -1.0, 1.0)
            want_out = acc.copy()
            d_qs, metas = [], []
            for k in range(n_src):
                src = make_input(rng, n, F32, -2.0 - k, 3.0 + 0.5 * k)      # every source has its own range -> its own parameters
                s_in, z_in = port.compute_quant_params(src, dt_q)
                q = port.quantize(src, dt_q, s_in, z_in, NEAREST, semantics=SEM_BODY)
                want_out = port.dequantize(q, dt_q, n, dt_out, s_in, z_in, ADD, out=want_out, semantics=SEM_BODY)
                d_qs.append(_dev_bytes(q, in_off))
                meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
                ctx.compute_meta_on_stream(_dev_bytes(src).data_ptr(), DT[F32], n, DT[dt_q], meta.data_ptr(), 0, dev, st)
                metas.append(meta)
            want_params = port.compute_quant_params(want_out, UINT8)
            d_acc = _dev_bytes(acc, out_off)
            nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
            nxt2 = torch.zeros(64, dtype=torch.uint8, device="cuda")
            before = ctx.kernel_launches
            ctx.dequantize_sum_minmax_on_stream([t.data_ptr() for t in d_qs], DT[dt_q], d_acc.data_ptr(), DT[dt_out], n,
                                                [m.data_ptr() for m in metas], DT[UINT8], nxt.data_ptr(), nxt2.data_ptr(), dev, st)
            launched = ctx.kernel_launches - before
            got_out = d_acc.cpu().numpy().view(odt)
            assert np.array_equal(got_out.view(np.uint8), want_out.view(np.uint8)), (n, n_src, in_off, out_off)
            assert _meta_tuple(nxt) == (_f32_bits(want_params[0]), 0, want_params[1]), (n, n_src, in_off, out_off)
            assert torch.equal(nxt, nxt2)
            if n >= 64 and in_off == 0 and out_off == 0:
                assert launched == 1, "aligned buffers must take the ONE multi-source kernel"


def test_dequantize_sum_minmax_mixed_alignment_and_flagged_source():
    """Sources that do not share a byte phase take the separate launches (same results); a flagged parameter block stops
    the whole launch and flags the block it was to produce."""
    from gpu_util import DT
    ctx = _ctx()
    rng = np.random.default_rng(19)
    dev, st = _site()
    n = 300_001
    acc = make_input(rng, n, F32, -1.0, 1.0)
    want_out = acc.copy()
    d_qs, metas = [], []
    for k, off in enumerate((0, 4, 16)):
        src = make_input(rng, n, F32, -1.0, 2.0 + k)
        s_in, z_in = port.compute_quant_params(src, UINT4)
        q = port.quantize(src, UINT4, s_in, z_in, NEAREST, semantics=SEM_BODY)
        want_out = port.dequantize(q, UINT4, n, F32, s_in, z_in, ADD, out=want_out, semantics=SEM_BODY)
        d_qs.append(_dev_bytes(q, off))
        meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
        ctx.compute_meta_on_stream(_dev_bytes(src).data_ptr(), DT[F32], n, DT[UINT4], meta.data_ptr(), 0, dev, st)
        metas.append(meta)
    d_acc = _dev_bytes(acc)
    nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
    before = ctx.kernel_launches
    ctx.dequantize_sum_minmax_on_stream([t.data_ptr() for t in d_qs], DT[UINT4], d_acc.data_ptr(), DT[F32], n, [m.data_ptr() for m in metas],
                                        DT[UINT8], nxt.data_ptr(), 0, dev, st)
    assert ctx.kernel_launches - before == 3
    assert np.array_equal(d_acc.cpu().numpy().view(np.uint8), want_out.view(np.uint8))
    want = port.compute_quant_params(want_out, UINT8)
    assert _meta_tuple(nxt) == (_f32_bits(want[0]), 0, want[1])
    # flagged source (error word = 1): accumulator untouched, produced block flagged
    bad = metas[1].clone()
    bad[4:8] = torch.tensor([1, 0, 0, 0], dtype=torch.uint8, device="cuda")
    d_qs = [_dev_bytes(np.zeros(packed_bytes(UINT4, n), dtype=np.uint8)) for _ in range(3)]
    d_acc = _dev_bytes(acc)
    nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.dequantize_sum_minmax_on_stream([t.data_ptr() for t in d_qs], DT[UINT4], d_acc.data_ptr(), DT[F32], n,
                                        [metas[0].data_ptr(), bad.data_ptr(), metas[2].data_ptr()], DT[UINT8], nxt.data_ptr(), 0, dev, st)
    assert np.array_equal(d_acc.cpu().numpy().view(np.uint8), acc.view(np.uint8))
    assert _meta_tuple(nxt)[1] == 1


def test_dequantize_add_minmax_ignores_nan_like_the_minmax_kernel():
    from gpu_util import DT
    ctx = _ctx()
    dev, st = _site()
    n = 70_000
    rng = np.random.default_rng(3)
    acc = rng.uniform(-1, 1, n).astype(np.float32)
    acc[[5, 4099, n - 1]] = np.nan
    q = rng.integers(0, 256, n, dtype=np.uint8)
    want_out = port.dequantize(q, UINT8, n, F32, 0.01, 128, ADD, out=acc.copy(), semantics=SEM_BODY)
    want = port.compute_quant_params(want_out, UINT8)
    d_q, d_acc = _dev_bytes(q), _dev_bytes(acc)
    meta_in = torch.zeros(64, dtype=torch.uint8, device="cuda")
    two = _dev_bytes(np.array([-128 * 0.01, 127 * 0.01], dtype=np.float32))
    ctx.compute_meta_on_stream(two.data_ptr(), DT[F32], 2, DT[UINT8], meta_in.data_ptr(), 0, dev, st)
    assert _meta_tuple(meta_in)[2] == 128
    nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.dequantize_add_minmax_on_stream(d_q.data_ptr(), DT[UINT8], d_acc.data_ptr(), DT[F32], n, meta_in.data_ptr(), DT[UINT8], nxt.data_ptr(), 0, dev, st)
    got = d_acc.cpu().numpy().view(np.float32)
    s_in = struct.unpack("<f", meta_in.cpu().numpy().tobytes()[:4])[0]
    want_out = port.dequantize(q, UINT8, n, F32, s_in, 128, ADD, out=acc.copy(), semantics=SEM_BODY)
    want = port.compute_quant_params(want_out, UINT8)
    nan = np.isnan(want_out)
    assert nan.sum() == 3 and np.array_equal(np.isnan(got), nan)          # (NaN payloads differ between x86 and the GPU's fma)
    assert np.array_equal(got[~nan].view(np.uint32), want_out[~nan].view(np.uint32))
    assert _meta_tuple(nxt) == (_f32_bits(want[0]), 0, want[1])


# -------------------------------------------------------------------------------------------------------------
# dequantize-SET + forward (ring all-gather hop)
# -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_out", FLOAT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("dt_q", QUANT_DTYPES, ids=lambda d: DT_NAME[d])
def test_dequantize_forward_equals_dequantize_plus_copy(dt_q, dt_out):
    from gpu_util import DT
    ctx = _ctx()
    rng = np.random.default_rng(17)
    dev, st = _site()
    odt = np.float32 if dt_out == F32 else np.uint16
    for n in SIZES:
        for in_off, fwd_off in ((0, 0), (64, 32), (0, 8), (3, 3)):
            src = make_input(rng, n, F32, -1.0, 2.0)
            s, z = port.compute_quant_params(src, dt_q)
            q = port.quantize(src, dt_q, s, z, NEAREST, semantics=SEM_BODY)
            want = port.dequantize(q, dt_q, n, dt_out, s, z, SET, semantics=SEM_BODY)
            d_q = _dev_bytes(q, in_off)
            d_out = _dev_bytes(np.zeros(n, dtype=odt))
            d_fwd = _dev_bytes(np.full(q.size + 64, 0xEE, np.uint8), fwd_off)
            meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
            fmeta = torch.zeros(64, dtype=torch.uint8, device="cuda")
            d_src = _dev_bytes(src)
            ctx.compute_meta_on_stream(d_src.data_ptr(), DT[F32], n, DT[dt_q], meta.data_ptr(), 0, dev, st)
            ctx.dequantize_forward_on_stream(d_q.data_ptr(), DT[dt_q], d_out.data_ptr(), DT[dt_out], n, meta.data_ptr(), d_fwd.data_ptr(), fmeta.data_ptr(), dev, st)
            assert np.array_equal(d_out.cpu().numpy().view(odt).view(np.uint8), want.view(np.uint8)), (n, in_off, fwd_off)
            f = d_fwd.cpu().numpy()
            assert np.array_equal(f[:q.size], q) and (f[q.size:] == 0xEE).all(), (n, in_off, fwd_off)
            assert torch.equal(meta, fmeta)


# -------------------------------------------------------------------------------------------------------------
# many tensors, one launch
# -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_in", FLOAT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("dt_q", QUANT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("mode", (NEAREST, STOCHASTIC), ids=("nearest", "stochastic"))
def test_quantize_batch_equals_one_call_per_tensor(dt_in, dt_q, mode):
    from gpu_util import DT, MODE
    ctx = _ctx()
    rng = np.random.default_rng(23)
    dev, st = _site()
    sizes = [1, 3, 17, 64, 1000, 4096, 4097, 16384 * 3 + 7, 100_003, 250_000] + [int(v) for v in rng.integers(1, 60_000, 300)]
    xs, d_in, d_out, items, params = [], [], [], [], []
    for i, n in enumerate(sizes):
        x = make_input(rng, n, dt_in, -1.0 - i % 3, 1.0 + i % 5)
        if n > 40:
            x_f = as_f32(x).copy()
            x_f[7] = np.nan
            x_f[11] = np.inf
            x = x_f if dt_in == F32 else port.f32_to_bf16_bits(x_f)
        s, z = port.compute_quant_params(make_input(rng, 64, F32, -1.0 - i % 3, 1.0 + i % 5), dt_q)
        esz = 4 if dt_in == F32 else 2
        a = _dev_bytes(x, (i % 4) * esz * (1 if i % 7 else 0))               # some inputs off the 32-byte grid
        o = _dev_bytes(np.full(packed_bytes(dt_q, n) + 32, 0xAA, np.uint8), i % 5)
        xs.append(x); d_in.append(a); d_out.append(o); params.append((s, z))
        items.append((a.data_ptr(), o.data_ptr(), n, s, z))
    ctx.set_stochastic_threshold(0.37 if mode == STOCHASTIC else None)
    before = ctx.kernel_launches
    ctx.quantize_batch(items, DT[dt_in], DT[dt_q], MODE[mode], dev, st)
    assert ctx.kernel_launches - before == (len(sizes) + 255) // 256
    torch.cuda.synchronize()
    for i, n in enumerate(sizes):
        s, z = params[i]
        want = port.quantize(xs[i], dt_q, s, z, mode, xi=0.37, semantics=SEM_BODY)
        got = d_out[i].cpu().numpy()
        assert np.array_equal(got[:want.size], want), (i, n)
        assert (got[want.size:] == 0xAA).all(), (i, n, "wrote past the end")


def test_quantize_batch_torch_surface_and_empty_items():
    import piquant.torch as pt
    ts = [torch.rand(n, device="cuda") * 2 - 1 for n in (1000, 0, 77, 1_000_000)]
    scales, zps = [2 / 255] * 4, [128] * 4
    outs = pt.quantize_batch(ts, scales=scales, zero_points=zps, dtype=torch.uint8)
    for t, o in zip(ts, outs):
        assert torch.equal(o, pt.quantize(t, scale=2 / 255, zero_point=128, dtype=torch.uint8))


# -------------------------------------------------------------------------------------------------------------
# REVERSE tile order, one-shot quantize
# -------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("dt_in", FLOAT_DTYPES, ids=lambda d: DT_NAME[d])
@pytest.mark.parametrize("dt_q", QUANT_DTYPES, ids=lambda d: DT_NAME[d])
def test_reverse_tile_order_writes_the_same_bytes(dt_in, dt_q):
    from gpu_util import DT
    import piquant
    ctx = _ctx()
    rng = np.random.default_rng(29)
    dev, st = _site()
    for n in (100, 16384 * 2, 16384 * 5 + 77, 1_000_003):
        x = make_input(rng, n, dt_in)
        s, z = port.compute_quant_params(x, dt_q)
        want = port.quantize(x, dt_q, s, z, NEAREST, semantics=SEM_BODY)
        d_x = _dev_bytes(x)
        meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
        ctx.compute_meta_on_stream(d_x.data_ptr(), DT[dt_in], n, DT[dt_q], meta.data_ptr(), 0, dev, st)
        out = torch.zeros(want.size, dtype=torch.uint8, device="cuda")
        ctx.quantize_meta_on_stream(d_x.data_ptr(), DT[dt_in], out.data_ptr(), DT[dt_q], n, piquant.RoundMode.NEAREST, meta.data_ptr(),
                                    piquant.Context.FLAG_REVERSE, dev, st)
        assert np.array_equal(out.cpu().numpy(), want), n


def test_quantize_auto_is_two_launches_and_matches_two_calls():
    import piquant.torch as pt
    import piquant
    ctx = piquant.Context()
    for n in (1000, 27_264_000):
        x = torch.rand(n, device="cuda") * 3 - 1
        before = ctx.kernel_launches
        q, s, z = pt.quantize_auto(x, dtype=torch.uint8, ctx=ctx)
        assert ctx.kernel_launches - before == 2, "min/max(+params) and quantize"
        s2, z2 = pt.compute_quant_params(x, dtype=torch.uint8, ctx=ctx)
        assert (s, z) == (s2, z2)
        assert torch.equal(q, pt.quantize(x, scale=s2, zero_point=z2, dtype=torch.uint8, ctx=ctx))


# -------------------------------------------------------------------------------------------------------------
# flagged parameter blocks
# -------------------------------------------------------------------------------------------------------------

def test_kernels_skip_work_on_a_flagged_parameter_block():
    """An empty tensor yields a negative scale: the reference aborts in compute_quant_params (src/piquant.cpp:373).  On the device
    the block is flagged instead, every kernel that is handed it leaves its output alone, and the flag travels on."""
    from gpu_util import DT
    import piquant
    ctx = _ctx()
    dev, st = _site()
    meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.compute_meta_on_stream(0, DT[F32], 0, DT[UINT8], meta.data_ptr(), 0, dev, st)
    assert _meta_tuple(meta)[1] == 1
    x = torch.rand(100_000, device="cuda")
    out = torch.full((100_000,), 0x5A, dtype=torch.uint8, device="cuda")
    ctx.quantize_meta_on_stream(x.data_ptr(), DT[F32], out.data_ptr(), DT[UINT8], x.numel(), piquant.RoundMode.NEAREST, meta.data_ptr(), 0, dev, st)
    assert bool((out == 0x5A).all())
    acc = torch.full((100_000,), 7.0, device="cuda")
    ctx.dequantize_meta_on_stream(out.data_ptr(), DT[UINT8], acc.data_ptr(), DT[F32], x.numel(), piquant.ReduceOp.ADD, meta.data_ptr(), dev, st)
    assert bool((acc == 7.0).all())
    nxt = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.dequantize_add_minmax_on_stream(out.data_ptr(), DT[UINT8], acc.data_ptr(), DT[F32], x.numel(), meta.data_ptr(), DT[UINT8], nxt.data_ptr(), 0, dev, st)
    assert bool((acc == 7.0).all()) and _meta_tuple(nxt)[1] == 1
    with pytest.raises(ValueError):
        import piquant.torch as pt
        pt.meta_to_host(nxt)


# -------------------------------------------------------------------------------------------------------------
# streams and threads on ONE context
# -------------------------------------------------------------------------------------------------------------

def test_two_streams_one_context_device_resident_parameters():
    """compute_meta + quantize_meta sequences on two streams of one context, interleaved call by call: each stream has its own
    reduction scratch / ticket (the round-1 finding: one shared ticket could be incremented by two grids at once)."""
    from gpu_util import DT
    import piquant
    ctx = _ctx()
    dev = torch.cuda.current_device()
    streams = [torch.cuda.Stream() for _ in range(2)]
    rng = np.random.default_rng(31)
    n = 3_000_000
    xs = [rng.uniform(-1 - k, 2 + k, n).astype(np.float32) for k in range(2)]
    d_x = [_dev_bytes(x) for x in xs]
    torch.cuda.synchronize()
    metas = [[torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(8)] for _ in range(2)]
    outs = [[torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(8)] for _ in range(2)]
    for it in range(8):
        for k in range(2):
            st = streams[k].cuda_stream
            ctx.compute_meta_on_stream(d_x[k].data_ptr(), DT[F32], n, DT[UINT8], metas[k][it].data_ptr(), 0, dev, st)
            ctx.quantize_meta_on_stream(d_x[k].data_ptr(), DT[F32], outs[k][it].data_ptr(), DT[UINT8], n, piquant.RoundMode.NEAREST,
                                        metas[k][it].data_ptr(), 0, dev, st)
    torch.cuda.synchronize()
    for k in range(2):
        s, z = port.compute_quant_params(xs[k], UINT8)
        want = port.quantize(xs[k], UINT8, s, z, NEAREST, semantics=SEM_BODY)
        for it in range(8):
            assert _meta_tuple(metas[k][it]) == (_f32_bits(s), 0, z), (k, it)
            assert np.array_equal(outs[k][it].cpu().numpy(), want), (k, it)


def test_async_minmax_on_one_stream_then_blocking_params_on_another():
    """the advisor's scenario: compute_meta (asynchronous) on stream A, compute_quant_params (synchronous) on stream B right after"""
    from gpu_util import DT
    ctx = _ctx()
    dev = torch.cuda.current_device()
    a, b = torch.cuda.Stream(), torch.cuda.Stream()
    rng = np.random.default_rng(37)
    x_big = rng.uniform(-5, 9, 40_000_000).astype(np.float32)
    x_small = rng.uniform(-1, 1, 100_000).astype(np.float32)
    d_big, d_small = _dev_bytes(x_big), _dev_bytes(x_small)
    torch.cuda.synchronize()
    for _ in range(10):
        meta = torch.zeros(64, dtype=torch.uint8, device="cuda")
        ctx.compute_meta_on_stream(d_big.data_ptr(), DT[F32], x_big.size, DT[UINT8], meta.data_ptr(), 0, dev, a.cuda_stream)
        got = ctx.compute_quant_params_on_stream(d_small.data_ptr(), DT[F32], x_small.size, DT[UINT4], dev, b.cuda_stream)
        assert got == port.compute_quant_params(x_small, UINT4)
        a.synchronize()
        want = port.compute_quant_params(x_big, UINT8)
        assert _meta_tuple(meta) == (_f32_bits(want[0]), 0, want[1])


def test_threads_share_one_context_without_a_lock():
    """piquant.torch passes device and stream with every call: nothing in the context is mutated, so no caller-side lock."""
    import piquant
    import piquant.torch as pt

    ctx = piquant.Context()
    errors: list = []

    def work(seed: int) -> None:
        stream = torch.cuda.Stream()
        try:
            with torch.cuda.stream(stream):
                for it in range(12):
                    n = 700_000 + 64 * seed + it
                    x = torch.full((n,), float(seed + 1), device="cuda")
                    x[seed] = -1.0
                    s, z = pt.compute_quant_params(x, dtype=torch.uint8, ctx=ctx)
                    q = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8, ctx=ctx)
                    y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32, ctx=ctx)
                    q2, s2, z2 = pt.quantize_auto(x, dtype=torch.uint8, ctx=ctx)
                    if (s2, z2) != (s, z) or not torch.equal(q, q2) or (y - x).abs().max().item() > 0.5 * s * 1.001:
                        errors.append((seed, it))
            stream.synchronize()
        except Exception as e:      # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
    assert ctx.get_stream() == 0, "the torch surface must not touch the context's stream"


def test_cpu_tensor_after_a_call_on_another_stream_does_not_use_a_stale_stream():
    import piquant.torch as pt
    side = torch.cuda.Stream()
    x = torch.rand(50_000, device="cuda")
    with torch.cuda.stream(side):
        pt.quantize(x, scale=1 / 255, zero_point=0, dtype=torch.uint8)
    xc = torch.rand(50_000)
    q = pt.quantize(xc, scale=1 / 255, zero_point=0, dtype=torch.uint8)
    assert not q.is_cuda
    assert torch.equal(q, pt.quantize(xc.cuda(), scale=1 / 255, zero_point=0, dtype=torch.uint8).cpu())


# -------------------------------------------------------------------------------------------------------------
# pageable host tensors (what the reference's own callers pass)
# -------------------------------------------------------------------------------------------------------------

def test_pageable_host_tensors_through_the_bounce_pipeline():
    """> 8 MiB pageable tensors are moved by the library's copy workers through pinned bounce buffers; several chunks, ragged end"""
    from gpu_util import DT, MODE, OP
    ctx = _ctx()
    rng = np.random.default_rng(41)
    n = 3 * (8 << 20) + 12_345                  # three full pipeline chunks and a ragged fourth
    x = rng.uniform(-1, 1, n).astype(np.float32)
    s, z = ctx.compute_quant_params_ptr_float32(x.ctypes.data, DT[UINT8], n)
    assert (s, z) == port.compute_quant_params(x, UINT8)
    q = np.empty(n, dtype=np.uint8)
    ctx.quantize_ptr(x.ctypes.data, DT[F32], q.ctypes.data, DT[UINT8], n, s, z, MODE[0])
    want_q = port.quantize(x, UINT8, s, z, NEAREST, semantics=SEM_BODY)
    assert np.array_equal(q, want_q)
    acc = rng.uniform(-1, 1, n).astype(np.float32)
    want_acc = port.dequantize(want_q, UINT8, n, F32, s, z, ADD, out=acc.copy(), semantics=SEM_BODY)
    ctx.dequantize_ptr(q.ctypes.data, DT[UINT8], acc.ctypes.data, DT[F32], n, s, z, OP[1])
    assert np.array_equal(acc.view(np.uint32), want_acc.view(np.uint32))
    # bf16 -> u4, pageable in, device out
    xb = port.f32_to_bf16_bits(x[: (8 << 20) + 999])
    s4, z4 = port.compute_quant_params(xb, UINT4)
    out = torch.zeros(packed_bytes(UINT4, xb.size), dtype=torch.uint8, device="cuda")
    ctx.quantize_ptr(xb.ctypes.data, DT[BF16], out.data_ptr(), DT[UINT4], xb.size, s4, z4, MODE[0])
    assert np.array_equal(out.cpu().numpy(), port.quantize(xb, UINT4, s4, z4, NEAREST, semantics=SEM_BODY))


def test_prepared_quantize_batch_equals_per_call_and_takes_new_parameters():
    """piquant.torch.QuantizeBatch: descriptors built once, run() == quantize() per tensor; set_params is honoured."""
    import piquant.torch as pt
    ctx = _ctx()
    g = torch.Generator(device="cuda").manual_seed(5)
    tensors = [torch.rand(n, device="cuda", generator=g) * 4 - 2 for n in (1, 63, 4097, 100_003, 1_000_000)]
    params = [pt.compute_quant_params(t, dtype=torch.quint4x2, ctx=ctx) for t in tensors]
    batch = pt.QuantizeBatch(tensors, scales=[p[0] for p in params], zero_points=[p[1] for p in params], dtype=torch.quint4x2, ctx=ctx)
    before = ctx.kernel_launches
    outs = batch.run()
    assert ctx.kernel_launches - before == 1
    raw = lambda t: torch.empty(0, dtype=torch.uint8, device="cuda").set_(t.untyped_storage())      # noqa: E731
    for t, o, (s, z) in zip(tensors, outs, params):
        assert torch.equal(raw(o), raw(pt.quantize(t, scale=s, zero_point=z, dtype=torch.quint4x2, ctx=ctx)))
    batch.set_params(2, 0.5, 3)
    outs = batch.run()
    assert torch.equal(raw(outs[2]), raw(pt.quantize(tensors[2], scale=0.5, zero_point=3, dtype=torch.quint4x2, ctx=ctx)))
    assert torch.equal(raw(outs[4]), raw(pt.quantize(tensors[4], scale=params[4][0], zero_point=params[4][1], dtype=torch.quint4x2, ctx=ctx)))
