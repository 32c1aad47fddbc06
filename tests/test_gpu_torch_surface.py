"""piquant.torch on the GPU, written to read like the reference's own pytest suite
(reference python/tests/test_torch.py:23-53) with the tensors on `cuda`, plus what the B200 build adds:
CPU tensors through the host-pointer pipeline, `out=` accumulators, the current-stream contract."""
from __future__ import annotations

import math
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
TORCH_FLOAT_TYPES = (torch.bfloat16, torch.float32)
TORCH_QUANT_TYPES = (torch.quint8, torch.quint4x2, torch.quint2x4)
random.seed(128)


def numel() -> int:
    return random.randint(1, 128)


@pytest.fixture(scope="module")
def pt():
    import piquant.torch as pt
    return pt


@pytest.mark.parametrize("dtype_in", TORCH_FLOAT_TYPES)
@pytest.mark.parametrize("dtype_quantized", TORCH_QUANT_TYPES)
def test_compute_quant_config(pt, dtype_in, dtype_quantized):
    gen = torch.Generator(device="cuda").manual_seed(128)
    tensor = torch.empty(numel(), numel(), numel(), numel() % 16 + 1, dtype=dtype_in, device="cuda")
    tensor.uniform_(-1.0, 1.0, generator=gen)
    scale, zero_point = pt.compute_quant_params(tensor, dtype=dtype_quantized)
    assert scale > 0 and not math.isnan(scale) and not math.isinf(scale)
    # same answer as the reference's formula evaluated by torch
    mn, mx = tensor.float().min().item(), tensor.float().max().item()
    qmax = {torch.quint8: 255, torch.quint4x2: 15, torch.quint2x4: 3}[dtype_quantized]
    assert scale == pytest.approx((mx - mn) / qmax, rel=1e-6)
    assert zero_point == max(0, min(qmax, round(-mn / ((mx - mn) / qmax))))


@pytest.mark.parametrize("device", ("cuda", "cpu"))
@pytest.mark.parametrize("dtype_in", TORCH_FLOAT_TYPES)
@pytest.mark.parametrize("dtype_quantized", TORCH_QUANT_TYPES)
def test_quantize_roundtrip(pt, dtype_in, dtype_quantized, device):
    """reference test_quantize_roundtrip: dequantized output vs torch.quantize_per_tensor and vs the input."""
    gen = torch.Generator().manual_seed(128)
    inp = torch.empty(numel(), numel(), numel(), numel() % 16 + 1, dtype=dtype_in).uniform_(-1.0, 1.0, generator=gen).to(device)
    scale, zero_point = pt.compute_quant_params(inp, dtype=dtype_quantized)
    quantized_pi = pt.quantize(inp, zero_point=zero_point, scale=scale, dtype=dtype_quantized)
    assert quantized_pi.shape == inp.shape and quantized_pi.dtype == dtype_quantized and quantized_pi.device == inp.device
    dequantized_pi = pt.dequantize(quantized_pi, scale=scale, zero_point=zero_point, dtype=dtype_in)
    assert dequantized_pi.dtype == inp.dtype and dequantized_pi.device == inp.device and dequantized_pi.shape == inp.shape
    quantized_torch = torch.quantize_per_tensor(inp.float().cpu(), scale=scale, zero_point=zero_point, dtype=dtype_quantized)
    dequantized_torch = quantized_torch.dequantize().to(dtype_in)
    # torch rounds half-to-even, pi-quant half-away-from-zero: at exact ties of x/scale (a ~1e-5 fraction of
    # random f32 data) the two differ by one step; everywhere else they agree to 1e-3 like in the reference's test
    diff = (dequantized_torch.float() - dequantized_pi.cpu().float()).abs()
    assert diff.max().item() <= scale + 2.0**-7
    assert (diff > 1e-3).float().mean().item() < 1e-3
    assert torch.allclose(dequantized_torch, inp.cpu(), atol=scale * 0.5 + 1e-3)
    assert torch.allclose(dequantized_pi.cpu(), inp.cpu(), atol=scale * 0.5 + 1e-3)


def test_non_contiguous_inputs_and_uint8_alias(pt):
    x = torch.rand(64, 96, device="cuda") * 2 - 1
    xt = x.t()                                         # not contiguous: the surface makes it so
    s, z = pt.compute_quant_params(xt, dtype=torch.uint8)
    assert (s, z) == pt.compute_quant_params(x, dtype=torch.quint8)
    q = pt.quantize(xt, scale=s, zero_point=z, dtype=torch.uint8)
    assert q.dtype == torch.uint8 and q.shape == xt.shape
    want = torch.clamp(torch.round(xt.contiguous() / s) + z, 0, 255).to(torch.uint8)
    assert (q.int() - want.int()).abs().max().item() <= 1           # torch.round is half-to-even: ties may differ by one
    y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32)
    assert torch.allclose(y, xt.contiguous(), atol=0.5 * s + 1e-6)


def test_dequantize_add_accumulates_into_out(pt):
    """ring-reduce step: dequantize with reduce_op='add' into an accumulator (reference README.md:29)."""
    n = 100_003
    acc = torch.zeros(n, device="cuda")
    total = torch.zeros(n, device="cuda", dtype=torch.float64)
    for i in range(4):
        x = torch.rand(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(i)) * 2 - 1
        s, z = pt.compute_quant_params(x, dtype=torch.quint8)
        q = pt.quantize(x, scale=s, zero_point=z, dtype=torch.quint8)
        out = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32, reduce_op="add", out=acc)
        assert out is acc
        total += pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32).double()
    assert torch.allclose(acc.double(), total, atol=1e-5)
    fresh = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32, reduce_op="add")     # no out=: starts from zeros
    assert torch.equal(fresh, pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32))


def test_requantize_equals_quantize_then_dequantize(pt):
    x = torch.rand(250_001, device="cuda") * 6 - 3
    for dt in TORCH_QUANT_TYPES:
        s, z = pt.compute_quant_params(x, dtype=dt)
        fused = pt.requantize(x, scale=s, zero_point=z, dtype=dt)
        two = pt.dequantize(pt.quantize(x, scale=s, zero_point=z, dtype=dt), scale=s, zero_point=z, dtype=torch.float32)
        # the fused pass uses the scalar rounding step everywhere (std::round); it differs from the SIMD-body
        # formula only at |x/scale| = pred(0.5), absent from this data
        assert torch.equal(fused, two)


def test_calls_are_ordered_on_the_current_stream(pt):
    """Work is enqueued on torch's current stream: no explicit synchronisation between producer ops,
    piquant calls and consumer ops, also on a side stream."""
    side = torch.cuda.Stream()
    n = 4_000_000
    with torch.cuda.stream(side):
        x = torch.full((n,), 0.25, device="cuda")
        x.mul_(2.0)                                                   # producer on the side stream
        q = pt.quantize(x, scale=0.5, zero_point=3, dtype=torch.quint8)
        y = pt.dequantize(q, scale=0.5, zero_point=3, dtype=torch.float32)
        ok = (y == 0.5).all()                                         # consumer on the side stream
    side.synchronize()
    assert bool(ok)


def test_host_pointer_pipeline_matches_device_path(pt):
    """CPU tensors (pageable and pinned) stream through the GPU in chunks; several chunks and a ragged tail."""
    import piquant
    from piquant import DataType as D, ReduceOp, RoundMode

    n = (8 << 20) * 2 + 12_345                                          # 3 pipeline chunks
    ctx = piquant.Context()
    x = torch.empty(n).uniform_(-1, 1, generator=torch.Generator().manual_seed(3))
    s, z = pt.compute_quant_params(x, dtype=torch.quint4x2, ctx=ctx)           # host min/max path
    xd = x.cuda()
    assert (s, z) == pt.compute_quant_params(xd, dtype=torch.quint4x2, ctx=ctx)
    qd = pt.quantize(xd, scale=s, zero_point=z, dtype=torch.quint4x2, ctx=ctx)
    nbytes = (n + 1) // 2
    raw_d = torch.empty(0, dtype=torch.uint8, device="cuda").set_(qd.untyped_storage())[:nbytes].cpu()
    for pinned in (False, True):
        xi = x.pin_memory() if pinned else x
        qh = torch.empty(nbytes, dtype=torch.uint8)
        qh = qh.pin_memory() if pinned else qh
        ctx.quantize_ptr(xi.data_ptr(), D.F32, qh.data_ptr(), D.UINT4, n, s, z, RoundMode.NEAREST)
        assert torch.equal(qh, raw_d), f"pinned={pinned}"
        acc = torch.full((n,), 1.5)
        acc = acc.pin_memory() if pinned else acc
        ctx.dequantize_ptr(qh.data_ptr(), D.UINT4, acc.data_ptr(), D.F32, n, s, z, ReduceOp.ADD)
        want = pt.dequantize(qd, scale=s, zero_point=z, dtype=torch.float32, reduce_op="add", out=torch.full((n,), 1.5, device="cuda"), ctx=ctx)
        assert torch.equal(acc, want.cpu()), f"pinned={pinned}"
    # mixed: host input, device output
    qm = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    ctx.quantize_ptr(x.data_ptr(), D.F32, qm.data_ptr(), D.UINT4, n, s, z, RoundMode.NEAREST)
    torch.cuda.synchronize()
    assert torch.equal(qm.cpu(), raw_d)


def test_full_size_properties_1e9(pt):
    """BASELINE's full size (numel = 1e9): size-independent properties instead of an element-wise oracle run.
    (i) the 256-bin histogram of the quantized bytes equals the histogram computed by torch from the same
    formula; (ii) dequantize(quantize(x)) is within 0.5*scale of x everywhere; (iii) quantize is idempotent
    on its own dequantized output."""
    n = 1_000_000_000
    if torch.cuda.mem_get_info()[0] < 16 * 2**30:
        pytest.skip("needs 16 GiB of free device memory")
    x = torch.empty(n, device="cuda").uniform_(-1, 1, generator=torch.Generator(device="cuda").manual_seed(0))
    s, z = pt.compute_quant_params(x, dtype=torch.quint8)
    assert z == 128 and abs(s - 2 / 255) < 1e-6
    q = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8)
    hist = torch.bincount(q.view(-1).int(), minlength=256)
    assert int(hist.sum()) == n and int(hist[1:255].min()) > 0
    y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32)
    blk = 1 << 27
    worst = 0.0
    for i in range(0, n, blk):
        worst = max(worst, (y[i:i + blk] - x[i:i + blk]).abs().max().item())
    assert worst <= 0.5 * s * (1 + 1e-6) + 1e-7
    q2 = pt.quantize(y, scale=s, zero_point=z, dtype=torch.uint8)
    assert torch.equal(q, q2)
    # slices against the oracle, bit for bit (first, middle, last 4 Mi elements)
    from oracle import port
    for lo in (0, n // 2 - 77, n - (1 << 22)):
        xs = x[lo:lo + (1 << 22)].cpu().numpy()
        assert np.array_equal(q[lo:lo + (1 << 22)].cpu().numpy(), port.quantize(xs, port.UINT8, s, z))


def test_full_size_c2_bf16_quint4x2_round_trip_1e8(pt):
    """BASELINE config 2 at its full size: bf16 -> quint4x2 -> bf16, numel = 1e8.  Properties: error <= 0.5*scale
    (+ the bf16 rounding of the output), idempotence of quantize on dequantized data, 16 used levels, and the first /
    last 4 Mi elements against the oracle bit for bit."""
    n = 100_000_000
    x = torch.empty(n, device="cuda").uniform_(-1, 1, generator=torch.Generator(device="cuda").manual_seed(2)).bfloat16()
    s, z = pt.compute_quant_params(x, dtype=torch.quint4x2)
    q = pt.quantize(x, scale=s, zero_point=z, dtype=torch.quint4x2)
    y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.bfloat16)
    err = (y.float() - x.float()).abs().max().item()
    assert err <= 0.5 * s + 2.0**-8 * (1.0 + s)
    q2 = pt.quantize(y, scale=s, zero_point=z, dtype=torch.quint4x2)
    raw = torch.empty(0, dtype=torch.uint8, device="cuda").set_(q.untyped_storage())[: n // 2]
    raw2 = torch.empty(0, dtype=torch.uint8, device="cuda").set_(q2.untyped_storage())[: n // 2]
    assert torch.equal(raw, raw2)
    levels = torch.bincount((raw & 15).int(), minlength=16) + torch.bincount((raw >> 4).int(), minlength=16)
    # level 0 needs x/scale <= -7.5, i.e. x < -1 with inv_scale = fl(1/scale) = 7.4999995: unreachable on [-1, 1]
    assert int(levels.sum()) == n and int(levels[1:].min()) > 0
    from oracle import port
    m = 1 << 22
    for lo in (0, n - m):
        xs = x[lo:lo + m].view(torch.int16).cpu().numpy().view(np.uint16)
        want_q = port.quantize(xs, port.UINT4, s, z)
        assert np.array_equal(raw[lo // 2:(lo + m) // 2].cpu().numpy(), want_q)
        want_y = port.dequantize(want_q, port.UINT4, m, port.BF16, s, z)
        assert np.array_equal(y[lo:lo + m].view(torch.int16).cpu().numpy().view(np.uint16), want_y)


def test_full_size_c4_stochastic_and_c5_add_1e9(pt):
    """BASELINE configs 4 and 5 at numel = 1e9.  Stochastic: every element is either the truncated or the
    away-from-zero neighbour, consistent with ONE threshold (the one the context reports), and differs from
    nearest by at most one step.  ADD: linear -- accumulating the same quantized tensor k times gives k times
    the SET result, bit for bit where the sums are exact."""
    import piquant

    n = 1_000_000_000
    if torch.cuda.mem_get_info()[0] < 24 * 2**30:
        pytest.skip("needs 24 GiB of free device memory")
    ctx = piquant.Context()
    x = torch.empty(n, device="cuda").uniform_(-1, 1, generator=torch.Generator(device="cuda").manual_seed(4))
    s, z = pt.compute_quant_params(x, dtype=torch.uint8, ctx=ctx)
    qn = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8, ctx=ctx)
    qs = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8, round_mode="stochastic", ctx=ctx)
    xi = ctx.last_stochastic_threshold
    assert 0.0 <= xi < 1.0
    blk = 1 << 27
    inv = torch.tensor(1.0, dtype=torch.float32) / torch.tensor(s, dtype=torch.float32)
    for i in range(0, n, blk):
        r = x[i:i + blk] * inv.item()
        tr = torch.trunc(r)
        away = (r - tr).abs() > xi
        want = torch.clamp(tr + torch.where(away, torch.sign(r), torch.zeros_like(r)) + z, 0, 255).to(torch.uint8)
        assert torch.equal(qs[i:i + blk], want)
        assert (qs[i:i + blk].int() - qn[i:i + blk].int()).abs().max().item() <= 1
    del qs, x
    # C5: u8 -> f32 with the ADD store op is linear in the number of accumulations
    set_once = pt.dequantize(qn, scale=s, zero_point=z, dtype=torch.float32, ctx=ctx)
    acc = torch.zeros(n, device="cuda")
    for _ in range(4):
        pt.dequantize(qn, scale=s, zero_point=z, dtype=torch.float32, reduce_op="add", out=acc, ctx=ctx)
    # every accumulation is ONE fma(d, s, acc) in f32; d*s + acc is exact in f64 here (33 significant bits), so
    # rounding the f64 sum to f32 reproduces the fma bit for bit
    for i in range(0, n, blk):
        d = (qn[i:i + blk].int() - z).double()
        ref = torch.zeros(d.numel(), device="cuda", dtype=torch.float32)
        for _ in range(4):
            ref = (ref.double() + d * s).float()
        assert torch.equal(acc[i:i + blk], ref)
        assert torch.allclose(acc[i:i + blk], 4 * set_once[i:i + blk], rtol=0, atol=4e-7)


def test_full_size_c3_minmax_1e9_both_dtypes(pt):
    n = 1_000_000_000
    x = torch.empty(n, device="cuda").uniform_(-3, 5, generator=torch.Generator(device="cuda").manual_seed(3))
    x[123_456_789] = -7.25
    x[n - 1] = 9.5
    for t in (x, x.bfloat16()):
        mn, mx = t.float().min().item(), t.float().max().item()
        for dt, qmax in ((torch.quint8, 255), (torch.quint4x2, 15), (torch.quint2x4, 3)):
            s, z = pt.compute_quant_params(t, dtype=dt)
            want_s = np.float32((np.float64(mx) - np.float64(mn)) / qmax)
            want_z = int(max(0.0, min(float(qmax), float(np.round(0.0 - np.float64(mn) / ((np.float64(mx) - np.float64(mn)) / qmax))))))
            assert np.float32(s) == want_s and z == want_z


def test_device_parameter_path_is_cuda_graph_capturable(pt):
    """min/max -> parameters -> quantize -> dequantize-ADD with parameters that never leave the GPU has no host
    synchronisation, so the whole sequence can be captured once into a CUDA graph and replayed on new data."""
    import piquant
    from piquant import DataType as D, ReduceOp, RoundMode

    n = 3_000_007
    ctx = piquant.Context()
    x = torch.empty(n, device="cuda")
    q = torch.empty(n, dtype=torch.uint8, device="cuda")
    acc = torch.zeros(n, device="cuda")
    meta = pt.new_meta(x.device)

    def sequence():
        ctx.compute_meta_async_ptr(x.data_ptr(), D.F32, n, D.UINT8, meta.data_ptr())
        ctx.quantize_meta_async_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, RoundMode.NEAREST, meta.data_ptr())
        ctx.dequantize_meta_async_ptr(q.data_ptr(), D.UINT8, acc.data_ptr(), D.F32, n, ReduceOp.ADD, meta.data_ptr())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ctx.set_stream(side.cuda_stream)
        x.uniform_(-1, 1)
        sequence()                                   # warm-up outside the capture: per-device scratch gets allocated here
    side.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        sequence()
    launches = ctx.kernel_launches
    total = torch.zeros(n, device="cuda", dtype=torch.float64)
    acc.zero_()
    for i in range(5):
        x.uniform_(-(i + 1), i + 1, generator=torch.Generator(device="cuda").manual_seed(i))
        graph.replay()
        torch.cuda.synchronize()
        s, z = pt.meta_to_host(meta)
        assert (s, z) == pt.compute_quant_params(x, dtype=torch.quint8)       # same parameters as the host-synchronous path
        assert torch.equal(q, pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8))
        total += pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32).double()
    assert ctx.kernel_launches == launches           # replays issue no new launches through the library
    assert torch.allclose(acc.double(), total, atol=1e-4)


def test_quantize_auto_with_host_tensors_and_zero_copy_mode(pt, monkeypatch):
    """quantize_auto on CPU tensors falls back to the two synchronous calls; PIQUANT_CUDA_HOST_MODE=zerocopy makes
    kernels read pinned host memory in place over PCIe (measured slower than the copy-engine pipeline, kept as an option)."""
    import piquant

    n = 1_234_567
    x = torch.empty(n).uniform_(-2, 3, generator=torch.Generator().manual_seed(9))
    want_s, want_z = pt.compute_quant_params(x.cuda(), dtype=torch.quint8)
    want_q = pt.quantize(x.cuda(), scale=want_s, zero_point=want_z, dtype=torch.uint8).cpu()
    q, s, z = pt.quantize_auto(x, dtype=torch.uint8)
    assert (s, z) == (want_s, want_z) and torch.equal(q, want_q) and q.device.type == "cpu"
    monkeypatch.setenv("PIQUANT_CUDA_HOST_MODE", "zerocopy")
    ctx = piquant.Context()
    xp = x.pin_memory()
    before = ctx.kernel_launches
    assert pt.compute_quant_params(xp, dtype=torch.quint8, ctx=ctx) == (want_s, want_z)
    qp = torch.empty(n, dtype=torch.uint8).pin_memory()
    pt.quantize(xp, scale=want_s, zero_point=want_z, dtype=torch.uint8, ctx=ctx, out=qp)
    assert torch.equal(qp, want_q)
    assert ctx.kernel_launches - before == 2          # one kernel each, no staging chunks


def test_managed_memory_pointers(pt):
    """cudaMallocManaged buffers are device-accessible: the kernels run on them in place."""
    try:
        from cuda.bindings import runtime as rt
    except Exception:
        try:
            from cuda import cudart as rt
        except Exception:
            pytest.skip("cuda-python not importable")
    import ctypes

    import piquant
    from piquant import DataType as D, RoundMode

    n = 1_000_003
    err, px = rt.cudaMallocManaged(4 * n, rt.cudaMemAttachGlobal)
    assert err == rt.cudaError_t.cudaSuccess
    err, pq_ = rt.cudaMallocManaged(n, rt.cudaMemAttachGlobal)
    assert err == rt.cudaError_t.cudaSuccess
    try:
        x = np.ctypeslib.as_array((ctypes.c_float * n).from_address(int(px)))
        x[:] = np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)
        ctx = piquant.Context()
        s, z = ctx.compute_quant_params_ptr_float32(int(px), D.UINT8, n)
        ctx.quantize_ptr(int(px), D.F32, int(pq_), D.UINT8, n, s, z, RoundMode.NEAREST)
        ctx.synchronize()
        torch.cuda.synchronize()
        q = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(int(pq_))).copy()
        from oracle import port
        assert (s, z) == port.compute_quant_params(x, port.UINT8)
        assert np.array_equal(q, port.quantize(np.ascontiguousarray(x), port.UINT8, s, z))
    finally:
        rt.cudaFree(px)
        rt.cudaFree(pq_)
