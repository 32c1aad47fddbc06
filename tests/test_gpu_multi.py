"""Sharded path on real GPUs (needs >= 2 devices; `gpurun --gpus 2 -- python -m pytest tests -m gpu`):
both routes to whole-tensor parameters -- torch.distributed all-reduce and the native library's own
NCCL communicator -- must return, on every rank, exactly the oracle's single-tensor answer; shards
quantized independently must concatenate to the whole-tensor result."""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, numel: int, q) -> None:
    for p in (str(ROOT), str(ROOT / "pi-quant_b200"), str(ROOT / "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import piquant
        import piquant.torch as pt
        from oracle import port as orc
        from piquant import distributed as pd

        rng = np.random.default_rng(321)
        x = rng.uniform(-2, 5, numel).astype(np.float32)
        b, e = pd.shard_bounds(numel, world, rank)
        shard = torch.from_numpy(x[b:e]).cuda()
        ctx = piquant.Context()
        res = {}
        for name, tdt, odt in (("u8", torch.quint8, orc.UINT8), ("u4", torch.quint4x2, orc.UINT4)):
            want = orc.compute_quant_params(x, odt)
            got_torch = pd.compute_quant_params_sharded(shard, dtype=tdt, ctx=ctx)
            pd.init_native_comm(ctx)
            got_native = pt.compute_quant_params(shard, dtype=tdt, ctx=ctx)
            pd.destroy_native_comm(ctx)
            got_local = pt.compute_quant_params(shard, dtype=tdt, ctx=ctx)      # local again after destroy
            want_local = orc.compute_quant_params(x[b:e], odt)
            qs = pt.quantize(shard, scale=want[0], zero_point=want[1], dtype=tdt, ctx=ctx)
            nbytes = orc.packed_bytes(odt, e - b)
            raw = torch.empty(0, dtype=torch.uint8, device=qs.device).set_(qs.untyped_storage())[:nbytes].cpu().numpy()
            whole = orc.quantize(x, odt, want[0], want[1])
            per = 8 // orc.BITS[odt]
            res[name] = (got_torch == want, got_native == want, got_local == want_local,
                         bool(np.array_equal(raw, whole[b // per: b // per + nbytes])))
        # device-resident parameters of a SHARDED tensor: with a communicator the parameter block holds whole-tensor values
        pd.init_native_comm(ctx)
        meta = pt.new_meta(shard.device)
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.compute_meta_async_ptr(shard.data_ptr(), piquant.DataType.F32, shard.numel(), piquant.DataType.UINT8, meta.data_ptr())
        res["sharded_meta"] = (pt.meta_to_host(meta) == orc.compute_quant_params(x, orc.UINT8),)
        pd.destroy_native_comm(ctx)
        # quantized ring all-reduce: every rank ends with bit-identical values, close to the exact sum
        for tdt, qdt, transport, rmode in ((torch.float32, torch.quint8, "nccl", "nearest"), (torch.bfloat16, torch.quint8, "nccl", "nearest"),
                                           (torch.float32, torch.quint4x2, "nccl", "nearest"), (torch.float32, torch.quint8, "p2p", "nearest"),
                                           (torch.bfloat16, torch.quint4x2, "p2p", "nearest"), (torch.float32, torch.quint8, "p2p", "nearest"),
                                           (torch.float32, torch.quint8, "p2p", "stochastic_per_element"),
                                           (torch.float32, torch.quint8, "nccl", "stochastic_per_element")):
            tol_steps = 1.0 if rmode == "nearest" else 2.0          # per-element SR: up to one step per hop instead of half a step
            g = torch.Generator(device="cuda").manual_seed(100 + rank)
            t = (torch.rand(1_000_003, device="cuda", generator=g) * 2 - 1).to(tdt)
            exact = t.double().clone()
            dist.all_reduce(exact)
            pd.quantized_all_reduce_(t, dtype=qdt, ctx=ctx, transport=transport, round_mode=rmode)
            gathered = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            identical = all(torch.equal(gathered[0].view(torch.uint8), gi.view(torch.uint8)) for gi in gathered)
            qmax = {torch.quint8: 255, torch.quint4x2: 15}[qdt]
            step = 2.0 * world / qmax                     # |sum| <= world, so every hop's scale is <= 2*world/qmax
            err = (t.double() - exact).abs().max().item()
            bound = step * (0.5 * world + 0.5) * tol_steps + (0.02 * world if tdt == torch.bfloat16 else 1e-5)
            ok = (identical, err <= bound)
            if rmode != "nearest":      # unbiased: the mean error over 1e6 elements is far below one step
                ok += (abs((t.double() - exact).mean().item()) < 0.01 * step,)
            res[f"ring_{transport}_{tdt}_{qdt}_{rmode}_{len(res)}"] = ok
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_sharded_params_and_quantize_two_gpus():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 3_000_001, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        for name, flags in res.items():
            assert all(flags), (rank, name, flags)
