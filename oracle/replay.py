"""CPU replays of the quantized all-reduce algorithms of ``piquant.distributed`` -- TEST INFRASTRUCTURE ONLY.

Both functions restate a collective with nothing but the oracle's ``compute_quant_params`` / ``quantize`` /
``dequantize`` (``oracle.port``: the restatement of reference src/piquant.cpp:222-259 and src/kernels/*.inl), chunk
by chunk and in the order the GPU implementation applies them, so that a GPU result can be compared BIT FOR BIT.
Inputs are the per-rank host arrays (float32, or uint16 holding bf16 bits); ``shard_bounds`` is
``piquant.distributed.shard_bounds`` (a pure function), ``align`` its ``SHARD_ALIGN``.  Used by tests/test_gpu_multi.py
and by the parity field of bench.py's all-reduce leg.
"""
from __future__ import annotations

import numpy as np

from . import port as orc


def _lanes(n: int, world: int, lanes: int, align: int):
    if n < lanes * world * align:
        lanes = 1
    per = n // lanes // align * align
    return [(lane * per, (lane + 1) * per if lane < lanes - 1 else n) for lane in range(lanes)]


def direct_all_reduce(inputs, qdt: int, fdt: int, shard_bounds, align: int, lanes: int = 1) -> np.ndarray:
    """``quantized_all_reduce_(algorithm="direct")``: chunk c is owned by rank c; every other rank quantizes ITS chunk c
    with that chunk's own parameters, the owner adds the dequantized chunks to its float chunk in rank order
    (dequantize with the ADD store op), quantizes the sums once and EVERY rank takes the dequantized values of those
    packed bytes.  Every lane is an independent all-reduce of its contiguous part.  Returns the result's bytes."""
    world, n = len(inputs), inputs[0].size
    out = np.empty_like(inputs[0])
    for p0, p1 in _lanes(n, world, lanes, align):
        for c in range(world):
            b, e = shard_bounds(p1 - p0, world, c)
            b, e = b + p0, e + p0
            if e == b:
                continue
            acc = inputs[c][b:e].copy()
            for r in range(world):
                if r == c:
                    continue
                s, z = orc.compute_quant_params(inputs[r][b:e], qdt)
                acc = orc.dequantize(orc.quantize(inputs[r][b:e], qdt, s, z), qdt, e - b, fdt, s, z, orc.ADD, out=acc)
            s, z = orc.compute_quant_params(acc, qdt)
            out[b:e] = orc.dequantize(orc.quantize(acc, qdt, s, z), qdt, e - b, fdt, s, z, orc.SET)
    return out.view(np.uint8)


def ring_all_reduce(inputs, qdt: int, fdt: int, shard_bounds, align: int, lanes: int = 1) -> np.ndarray:
    """``quantized_all_reduce_(algorithm="ring")``: chunk c starts on rank c, is quantized with the parameters of the
    running sum at every hop and accumulated with dequantize-ADD on the next rank; the rank that holds the complete sum
    quantizes it once more and EVERY rank takes the dequantized values of those packed bytes."""
    world, n = len(inputs), inputs[0].size
    out = np.empty_like(inputs[0])
    per = n // lanes // align * align           # (the ring form keeps the requested lane count whatever the size)
    for lane in range(lanes):
        p0, p1 = lane * per, ((lane + 1) * per if lane < lanes - 1 else n)
        for c in range(world):
            b, e = shard_bounds(p1 - p0, world, c)
            b, e = b + p0, e + p0
            if e == b:
                continue
            acc = inputs[c][b:e].copy()
            for k in range(1, world):
                s, z = orc.compute_quant_params(acc, qdt)
                q = orc.quantize(acc, qdt, s, z)
                acc = orc.dequantize(q, qdt, e - b, fdt, s, z, orc.ADD, out=inputs[(c + k) % world][b:e].copy())
            s, z = orc.compute_quant_params(acc, qdt)
            out[b:e] = orc.dequantize(orc.quantize(acc, qdt, s, z), qdt, e - b, fdt, s, z, orc.SET)
    return out.view(np.uint8)
