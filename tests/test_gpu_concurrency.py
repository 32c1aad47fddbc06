"""Threads, streams and contexts: the reference's context methods are const and callable from several threads
(reference src/piquant.cpp:194-211); here that means a mutex around the dispatcher, stream-ordered launches and
one self-resetting work counter per stream for the persistent TMA kernels."""
from __future__ import annotations

import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _work(ctx, stream, seed: int, variant: int, iters: int, errors: list) -> None:
    import piquant.torch as pt

    try:
        with torch.cuda.stream(stream):
            g = torch.Generator(device="cuda").manual_seed(seed)
            for it in range(iters):
                n = 300_000 + 4099 * ((seed + it) % 7)
                x = torch.empty(n, device="cuda").uniform_(-1 - seed, 1 + seed, generator=g)
                s, z = pt.compute_quant_params(x, dtype=torch.quint8, ctx=ctx)
                q = pt.quantize(x, scale=s, zero_point=z, dtype=torch.uint8, ctx=ctx)
                y = pt.dequantize(q, scale=s, zero_point=z, dtype=torch.float32, ctx=ctx)
                want = torch.clamp(torch.trunc(x * (torch.tensor(1.0) / torch.tensor(s)).item() + torch.where(x >= 0, 0.5, -0.5)) + z, 0, 255)
                if not torch.equal(q.float(), want) or (y - x).abs().max().item() > 0.5 * s * 1.001:
                    errors.append((seed, it, n))
        stream.synchronize()
    except Exception as e:      # noqa: BLE001
        errors.append((seed, repr(e)))


@pytest.mark.parametrize("variant", (1, 2), ids=("direct", "tma"))
def test_threads_with_own_context_and_stream(variant):
    import piquant

    errors: list = []
    threads = []
    for i in range(6):
        ctx = piquant.Context()
        ctx.set_kernel_variant(variant)
        threads.append(threading.Thread(target=_work, args=(ctx, torch.cuda.Stream(), i, variant, 12, errors)))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


@pytest.mark.parametrize("variant", (1, 2), ids=("direct", "tma"))
def test_threads_sharing_one_context_on_different_streams(variant):
    """piquant.torch binds the context to the caller's current stream on every call; with a shared context the
    bind + launch of two threads may interleave, so callers sharing a context must share the stream or serialise --
    here each thread serialises its bind+call pairs with a lock, the launches themselves still overlap on the GPU."""
    import piquant
    import piquant.torch as pt

    ctx = piquant.Context()
    ctx.set_kernel_variant(variant)
    lock = threading.Lock()
    errors: list = []

    def work(seed: int) -> None:
        stream = torch.cuda.Stream()
        try:
            with torch.cuda.stream(stream):
                for it in range(10):
                    n = 1_000_000 + 64 * seed + it
                    x = torch.full((n,), float(seed + 1), device="cuda")
                    with lock:
                        q = pt.quantize(x, scale=0.5, zero_point=1, dtype=torch.uint8, ctx=ctx)
                    with lock:
                        y = pt.dequantize(q, scale=0.5, zero_point=1, dtype=torch.float32, ctx=ctx)
                    if not bool((y == float(seed + 1)).all()):
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
    assert ctx.kernel_launches == 8 * 10 * 2


def test_many_streams_recycle_the_work_counter_table():
    """More distinct streams than counter slots (16): the table is recycled after a device sync."""
    import piquant
    import piquant.torch as pt

    ctx = piquant.Context()
    ctx.set_kernel_variant(2)
    x = torch.rand(2_000_000, device="cuda") * 2 - 1
    ref = pt.quantize(x, scale=2 / 255, zero_point=128, dtype=torch.uint8, ctx=ctx)
    torch.cuda.synchronize()
    for _ in range(40):
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            q = pt.quantize(x, scale=2 / 255, zero_point=128, dtype=torch.uint8, ctx=ctx)
        st.synchronize()
        assert torch.equal(q, ref)


def test_threads_sharing_one_context_draw_their_stochastic_numbers_safely():
    """Stochastic calls from several threads on ONE context and one stream: the per-call threshold / Philox key comes from a
    generator inside the context (the reference's is thread_local, src/piquant.cpp:194-195); concurrent draws must neither crash
    nor hand out a torn value -- every output is consistent with SOME single threshold in [0, 1) resp. some key."""
    import piquant
    from piquant import DataType as D, RoundMode

    ctx = piquant.Context()
    errors: list = []
    n = 200_000

    def work(seed: int) -> None:
        try:
            x = torch.full((n,), 0.3, device="cuda")
            q = torch.empty(n, dtype=torch.uint8, device="cuda")
            for it in range(40):
                mode = RoundMode.STOCHASTIC if it % 2 == 0 else RoundMode.STOCHASTIC_PER_ELEMENT
                ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, 1.0, 0, mode)
                torch.cuda.synchronize()
                lo, hi = int(q.min().item()), int(q.max().item())
                if mode == RoundMode.STOCHASTIC and not (lo == hi and lo in (0, 1)):
                    errors.append((seed, it, lo, hi))                  # one threshold per call: constant input -> constant output
                if mode == RoundMode.STOCHASTIC_PER_ELEMENT and not (lo == 0 and hi == 1 and abs(q.float().mean().item() - 0.3) < 0.01):
                    errors.append((seed, it, lo, hi, q.float().mean().item()))
        except Exception as e:      # noqa: BLE001
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
    assert ctx.kernel_launches == 6 * 40
