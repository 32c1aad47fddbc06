"""N > 1 host logic on CPUs: world_size-2 `gloo` process group.

What runs here is everything of the sharded path EXCEPT the CUDA kernel: contiguous shard bounds,
the one {-min, max} MAX all-reduce, and the native library's host-side parameter arithmetic.  The
local min/max of a shard -- a CUDA kernel in production -- is computed with torch on the CPU inside
this TEST only; results must equal the oracle's whole-tensor answer bit for bit on every rank."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, numel: int, q) -> None:
    for p in (str(ROOT), str(ROOT / "pi-quant_b200"), str(ROOT / "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import port as orc
        from piquant import distributed as pd

        rng = np.random.default_rng(123)
        x = rng.uniform(-3, 7, numel).astype(np.float32)
        b, e = pd.shard_bounds(numel, world, rank)
        shard = torch.from_numpy(x[b:e])
        fmax = torch.finfo(torch.float32).max
        if shard.numel():
            local = torch.stack([-shard.min(), shard.max()]).float()     # stand-in for the min/max kernel (test only)
        else:
            local = torch.tensor([-fmax, -fmax])
        mn, mx = pd.combine_minmax(local)
        out = {}
        for name, tdt, odt in (("u8", torch.quint8, orc.UINT8), ("u4", torch.quint4x2, orc.UINT4), ("u2", torch.quint2x4, orc.UINT2)):
            got = pd.params_from_minmax(mn, mx, tdt)
            want = orc.compute_quant_params(x, odt)
            out[name] = (got, want)
            # shards are independent: quantizing a shard with the shared parameters gives exactly the
            # bytes of the whole-tensor result (boundaries never split a packed byte)
            scale, zp = want
            whole = orc.quantize(x, odt, scale, zp)
            part = orc.quantize(x[b:e], odt, scale, zp) if e > b else np.zeros(0, np.uint8)
            per = 8 // orc.BITS[odt]
            assert b % per == 0
            assert np.array_equal(part, whole[b // per: b // per + part.size])
        q.put((rank, (b, e), out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("numel", (1_000_003, 300, 5))
def test_sharded_quant_params_world_size_2(numel):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, numel, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bounds = sorted(r[1] for r in results)
    assert bounds[0][0] == 0 and bounds[-1][1] == numel and bounds[0][1] == bounds[1][0]
    for _, _, out in results:
        for name, (got, want) in out.items():
            assert np.float32(got[0]).tobytes() == np.float32(want[0]).tobytes() and got[1] == want[1], (name, got, want)


def test_shard_bounds_properties():
    for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from piquant.distributed import SHARD_ALIGN, shard_bounds

    for numel in (0, 1, 127, 128, 129, 1000, 27_264_000, 1_000_000_000, 10**9 + 7):
        for world in (1, 2, 4, 8):
            prev_end = 0
            for rank in range(world):
                b, e = shard_bounds(numel, world, rank)
                assert b == prev_end and e >= b
                if rank < world - 1:
                    assert e % SHARD_ALIGN == 0
                prev_end = e
            assert prev_end == numel
    # BASELINE config 5: 1e9 over 8 GPUs = 8 shards of 125 M elements
    assert [shard_bounds(10**9, 8, r) for r in (0, 7)] == [(0, 125_000_000), (875_000_000, 10**9)]


def test_ring_schedule_is_an_all_reduce():
    """quantized_all_reduce_'s chunk schedule, simulated for W ranks with exact arithmetic: what a rank sends
    at step s is what its successor expects, reduce-scatter leaves chunk (r+1)%W complete on rank r, and the
    all-gather spreads every reduced chunk to every rank."""
    for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from piquant.distributed import ring_schedule

    for world in (2, 3, 4, 8):
        data = [np.arange(world * 3, dtype=np.float64).reshape(world, 3) * (r + 1) for r in range(world)]
        want = sum(data)
        sched = [ring_schedule(world, r) for r in range(world)]
        for phase, combine in ((0, np.add), (1, lambda old, new: new)):
            for s in range(world - 1):
                msgs = [data[r][sched[r][phase][s][0]].copy() for r in range(world)]
                for r in range(world):
                    src = (r - 1) % world
                    assert sched[src][phase][s][0] == sched[r][phase][s][1]
                    data[r][sched[r][phase][s][1]] = combine(data[r][sched[r][phase][s][1]], msgs[src])
            if phase == 0:
                for r in range(world):
                    assert np.array_equal(data[r][(r + 1) % world], want[(r + 1) % world])
        for r in range(world):
            assert np.array_equal(data[r], want)


def test_gather_pieces_cover_the_chunk_at_vector_aligned_cuts():
    """The piece table of the multicast gather: a partition of [0, numel) into <= pieces non-empty ranges whose interior cuts are
    multiples of 256 elements; the same on every rank by construction (pure function)."""
    from piquant.distributed import gather_pieces

    for numel in (0, 1, 255, 256, 257, 1000, 4096, 33_554_432, 33_554_432 + 77, 1_000_003):
        for pieces in (1, 2, 4, 7):
            got = gather_pieces(numel, pieces)
            assert len(got) <= pieces
            if numel == 0:
                assert got == []
                continue
            assert got[0][0] == 0 and got[-1][1] == numel
            for (lo, hi), nxt in zip(got, got[1:] + [None]):
                assert hi > lo and lo % 256 == 0
                if nxt is not None:
                    assert nxt[0] == hi
            if numel >= pieces * 512:                      # large chunks are cut evenly (within one alignment unit)
                sizes = [hi - lo for lo, hi in got]
                assert len(got) == pieces and max(sizes) - min(sizes) <= 512
