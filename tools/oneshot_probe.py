#!/usr/bin/env python
"""Where does the time of the one-shot quantize (min/max -> parameters -> quantize) go, and does L2 help the second pass?

    python tools/oneshot_probe.py            (one B200)

For f32 -> uint8 at several sizes, CUDA-event times (rotating over distinct buffers so that nothing is warm by accident):
  minmax            the reduction alone (evict_first loads)
  minmax_keep       the reduction alone with L2::evict_last loads
  quant_cold        quantize alone
  pair_fwd          min/max (keep) + quantize in tile order        (device-resident parameters, one stream, no sync)
  pair_rev          min/max (keep) + quantize from the END of the tensor (LaunchCfg::reverse)
  pair_nokeep_fwd   min/max (evict_first) + quantize in tile order
  auto_wall         piquant_cuda_quantize_auto, wall clock incl. its one synchronisation
  two_calls_wall    compute_quant_params + quantize, wall clock (two synchronisations)
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import piquant  # noqa: E402
from piquant import DataType as D, RoundMode  # noqa: E402


def main() -> None:
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = piquant.Context()
    st = torch.cuda.current_stream().cuda_stream
    LOCAL, KEEP, REV = piquant.Context.FLAG_LOCAL, piquant.Context.FLAG_KEEP_IN_L2, piquant.Context.FLAG_REVERSE
    pool = torch.empty(1_000_000_000, dtype=torch.float32, device=dev).uniform_(-1, 1)
    qpool = torch.empty(1_000_000_000, dtype=torch.uint8, device=dev)
    meta = torch.zeros(64, dtype=torch.uint8, device=dev)
    out4 = torch.zeros(4, dtype=torch.float32, device=dev)
    print(f"{'numel':>12} {'MB':>7} | " + " ".join(f"{k:>16}" for k in ("minmax", "minmax_keep", "quant_cold", "pair_fwd", "pair_rev", "pair_nokeep_fwd", "auto_wall", "two_calls_wall")))
    for n in (2_000_000, 4_000_000, 8_000_000, 12_000_000, 16_000_000, 20_000_000, 27_264_000, 40_000_000, 64_000_000):
        wins = max(3, min(32, pool.numel() // n))
        xs = [pool[i * n:(i + 1) * n] for i in range(wins)]
        qs = [qpool[i * n:(i + 1) * n] for i in range(wins)]
        s, z = ctx.compute_quant_params_ptr_float32(xs[0].data_ptr(), D.UINT8, n)

        def ev(fn, reps=40):
            for k in range(4):
                fn(k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps):
                fn(4 + k)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps * 1e3

        def wall(fn, reps=40):
            for k in range(4):
                fn(k)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(reps):
                fn(4 + k)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / reps * 1e6

        def mm(flags):
            return lambda k: ctx.minmax_on_stream(xs[k % wins].data_ptr(), D.F32, n, out4.data_ptr(), LOCAL | flags, 0, st)

        def quant(k):
            ctx.quantize_on_stream(xs[k % wins].data_ptr(), D.F32, qs[k % wins].data_ptr(), D.UINT8, n, s, z, RoundMode.NEAREST, 0, st)

        def pair(mm_flags, q_flags):
            def fn(k):
                x, q = xs[k % wins], qs[k % wins]
                ctx.compute_meta_on_stream(x.data_ptr(), D.F32, n, D.UINT8, meta.data_ptr(), LOCAL | mm_flags, 0, st)
                ctx.quantize_meta_on_stream(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, RoundMode.NEAREST, meta.data_ptr(), q_flags, 0, st)
            return fn

        def auto(k):
            ctx.quantize_auto_on_stream(xs[k % wins].data_ptr(), D.F32, qs[k % wins].data_ptr(), D.UINT8, n, RoundMode.NEAREST, 0, st)

        def two(k):
            s_, z_ = ctx.compute_quant_params_on_stream(xs[k % wins].data_ptr(), D.F32, n, D.UINT8, 0, st)
            ctx.quantize_on_stream(xs[k % wins].data_ptr(), D.F32, qs[k % wins].data_ptr(), D.UINT8, n, s_, z_, RoundMode.NEAREST, 0, st)

        # NOTE compute_meta_on_stream turns KEEP on by itself for tensors <= 96 MB; the no-keep pair therefore uses min/max + quantize by value
        def pair_nokeep(k):
            x, q = xs[k % wins], qs[k % wins]
            ctx.minmax_on_stream(x.data_ptr(), D.F32, n, out4.data_ptr(), LOCAL, 0, st)
            ctx.quantize_on_stream(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, s, z, RoundMode.NEAREST, 0, st)

        row = [ev(mm(0)), ev(mm(KEEP)), ev(quant), ev(pair(KEEP, 0)), ev(pair(KEEP, REV)), ev(pair_nokeep), wall(auto), wall(two)]
        print(f"{n:>12} {4 * n / 1e6:>7.1f} | " + " ".join(f"{v:>16.2f}" for v in row), flush=True)


if __name__ == "__main__":
    main()
