#!/usr/bin/env python
"""Generates tests/golden/piquant_golden.npz from the UNMODIFIED reference (oracle/_ref/libpiquant_ref.so,
built from /root/reference by oracle/Makefile).  Run in the build container, where /root/reference
exists; the fixture is committed so that the GPU box (no /root/reference) can check against it.

    python tests/golden/make_golden.py

Every record is produced by calling the reference's own C ABI (include/piquant.h:42-85) with a
4-thread context; stochastic records also store the per-call threshold xi inferred from the output
(see tests/helpers.py:infer_xi) so that the call can be replayed deterministically.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

from helpers import DEQUANT_CELLS, QUANT_CELLS, DT_NAME, OP_NAME, as_f32, infer_xi, make_input, special_values, unpack  # noqa: E402
from oracle import ref  # noqa: E402
from oracle.port import (ADD, BF16, BITS, F32, NEAREST, SET, STOCHASTIC, UINT2, UINT4, UINT8,  # noqa: E402
                         f32_to_bf16_bits, packed_bytes)

NT = 4
SIZES = (1, 7, 1037, 2051)


def main() -> None:
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    ctx = ref.Context(NT)
    rng = np.random.default_rng(20260101)
    rec: dict[str, np.ndarray] = {}
    meta = []

    # --- quantize, nearest + stochastic
    for dt_in, dt_out in QUANT_CELLS:
        qmax = (1 << BITS[dt_out]) - 1
        for n in SIZES:
            for mode in (NEAREST, STOCHASTIC):
                x = make_input(rng, n, dt_in)
                if n == 1037 and mode == NEAREST:      # sprinkle the reference's corner cases in the middle
                    sp = special_values(0.25)
                    # keep |x/scale| == pred(0.5) out: there the reference's answer depends on whether the
                    # element falls in a SIMD body or a scalar tail (thread count / alignment), see DESIGN.md
                    sp = sp[np.abs(sp / np.float32(0.25)) != np.float32(0.49999997)]
                    sp = sp if dt_in == F32 else f32_to_bf16_bits(sp)
                    x[100:100 + sp.size] = sp
                scale, zp = (0.25, 3) if n == 1037 else ctx.compute_quant_params(x, dt_out)
                out = ctx.quantize(x, dt_out, scale, zp, mode)
                xi = -1.0
                if mode == STOCHASTIC:
                    got = infer_xi(as_f32(x), scale, zp, qmax, unpack(out, dt_out, n))
                    assert got is not None
                    xi = got
                key = f"quant/{DT_NAME[dt_in]}/{DT_NAME[dt_out]}/{'st' if mode else 'nr'}/{n}"
                rec[key + "/x"] = x
                rec[key + "/out"] = out
                rec[key + "/p"] = np.array([scale, zp, xi], dtype=np.float64)
                meta.append(key)

    # --- dequantize SET / ADD
    for dt_in, dt_out, op in DEQUANT_CELLS:
        for n in SIZES:
            q = rng.integers(0, 256, packed_bytes(dt_in, n)).astype(np.uint8)
            scale = float(np.float32(rng.uniform(0.01, 1.0)))
            zp = int(rng.integers(0, 1 << BITS[dt_in]))
            prev = rng.uniform(-1, 1, n).astype(np.float32)
            prev = prev if dt_out == F32 else f32_to_bf16_bits(prev)
            out = ctx.dequantize(q, dt_in, n, dt_out, scale, zp, op, out=prev.copy())
            key = f"dequant/{DT_NAME[dt_in]}/{DT_NAME[dt_out]}/{OP_NAME[op]}/{n}"
            rec[key + "/q"] = q
            rec[key + "/prev"] = prev
            rec[key + "/out"] = out
            rec[key + "/p"] = np.array([scale, zp], dtype=np.float64)
            meta.append(key)

    # --- compute_quant_params known answers + random
    kats = {
        "pm1": np.array([-1, 1], np.float32),
        "m3_5_1": np.array([-3, 5, 1], np.float32),
        "const42": np.full(100, 42.0, np.float32),
        "one_two": np.array([1, 2], np.float32),
        "u11": rng.uniform(-1, 1, 4099).astype(np.float32),
        "wide": rng.uniform(-300, 7, 2051).astype(np.float32),
    }
    for name, xf in kats.items():
        for dt_in in (F32, BF16):
            x = xf if dt_in == F32 else f32_to_bf16_bits(xf)
            for dt_q in (UINT2, UINT4, UINT8):
                s, z = ctx.compute_quant_params(x, dt_q)
                key = f"params/{name}/{DT_NAME[dt_in]}/{DT_NAME[dt_q]}"
                rec[key + "/x"] = x
                rec[key + "/p"] = np.array([np.float32(s).view(np.uint32), z], dtype=np.int64)
                meta.append(key)

    rec["__keys__"] = np.array(meta)
    rec["__info__"] = np.array([f"reference threads={NT} isa={ref.cpu_isa()}"])
    out_path = Path(__file__).with_name("piquant_golden.npz")
    np.savez_compressed(out_path, **rec)
    print(f"wrote {out_path} ({out_path.stat().st_size / 1024:.0f} KiB, {len(meta)} records)")


if __name__ == "__main__":
    main()
