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
            # both transports of the native exchange: {-min, max} swapped inside the min/max kernel through peer-mapped
            # mailboxes (default where the GPUs can map each other's memory), and ncclAllReduce
            transports = []
            got_native = None
            for tr in ((2, 1) if ctx.comm_transport == 2 else (1,)):
                ctx.comm_set_transport(tr)
                assert ctx.comm_transport == tr
                for _ in range(3):          # several exchanges in a row: the mailbox parity / sequence logic
                    got = pt.compute_quant_params(shard, dtype=tdt, ctx=ctx)
                    got_native = got if got_native is None or got_native == got else ("mismatch", got_native, got)
                transports.append(tr)
            res[name + "_transports"] = (len(transports) >= 1,)
            # a host-resident shard goes through the same exchange
            got_host = pt.compute_quant_params(shard.cpu(), dtype=tdt, ctx=ctx)
            res[name + "_host_shard"] = (got_host == want,)
            ctx.comm_set_transport(0)
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
        # ... unless the caller asks for this shard only (what the ring reduction does per chunk)
        ctx.compute_meta_on_stream(shard.data_ptr(), piquant.DataType.F32, shard.numel(), piquant.DataType.UINT8, meta.data_ptr(),
                                   piquant.Context.FLAG_LOCAL, rank, torch.cuda.current_stream().cuda_stream)
        res["local_meta_with_comm"] = (pt.meta_to_host(meta) == orc.compute_quant_params(x[b:e], orc.UINT8),)
        # one-shot quantize of a shard with whole-tensor parameters: one reduction launch (exchange inside) + one quantize launch
        q_auto, s_auto, z_auto = pt.quantize_auto(shard, dtype=torch.uint8, ctx=ctx)
        whole8 = orc.quantize(x, orc.UINT8, *orc.compute_quant_params(x, orc.UINT8))
        res["sharded_quantize_auto"] = ((s_auto, z_auto) == orc.compute_quant_params(x, orc.UINT8), bool(np.array_equal(q_auto.cpu().numpy(), whole8[b:e])))
        pd.destroy_native_comm(ctx)
        # quantized ring all-reduce: every rank ends with bit-identical values, close to the exact sum
        # the NVSwitch form: two all-to-all exchanges, ONE multi-source reduce kernel; replayed on the CPU with the oracle
        # the gather exchange through the NVSwitch multicast address (forced on: 2 ranks would not pick it), pieces with their own flags
        for tdt, qdt, numel, lanes, graph in ((torch.float32, torch.quint8, 1_000_003, 1, False), (torch.bfloat16, torch.quint4x2, 300_007, 2, False),
                                              (torch.float32, torch.quint2x4, 70_001, 1, False), (torch.float32, torch.quint8, 100, 1, False),
                                              (torch.float32, torch.quint8, 4_194_304, 1, True)):
            t = torch.zeros(numel, device="cuda", dtype=tdt)
            plan = pd.QuantizedAllReduce(t, dtype=qdt, ctx=ctx, lanes=lanes, multicast=True) if graph else None
            if (plan.plan.multicast if graph else True):
                for rep in range(3):
                    g = torch.Generator(device="cuda").manual_seed(700 + 10 * rep + rank)
                    t.copy_((torch.rand(numel, device="cuda", generator=g) * 2 - 1).to(tdt))
                    inputs = [torch.empty_like(t) for _ in range(world)]
                    dist.all_gather(inputs, t)
                    if graph:
                        plan()
                    else:
                        pd.quantized_all_reduce_(t, dtype=qdt, ctx=ctx, algorithm="direct", lanes=lanes, multicast=True)
                    want = _direct_on_the_oracle(orc, pd, [i.cpu() for i in inputs], qdt, lanes)
                    res[f"multicast_{tdt}_{qdt}_{numel}_graph{graph}_{rep}"] = (bool(np.array_equal(_bits(t.cpu()), want)),)
        for tdt, qdt, rmode, numel, lanes in ((torch.float32, torch.quint8, "nearest", 1_000_003, 1), (torch.bfloat16, torch.quint8, "nearest", 1_000_003, 1),
                                              (torch.float32, torch.quint4x2, "nearest", 300_007, 1), (torch.bfloat16, torch.quint2x4, "nearest", 70_001, 1),
                                              (torch.float32, torch.quint8, "nearest", 100, 1), (torch.float32, torch.quint8, "nearest", 4_194_304, 1),
                                              (torch.float32, torch.quint8, "nearest", 1_000_003, 2), (torch.bfloat16, torch.quint4x2, "nearest", 1_000_003, 3),
                                              (torch.float32, torch.quint8, "nearest", 200, 2),
                                              (torch.float32, torch.quint8, "stochastic_per_element", 1_000_003, 2)):
            g = torch.Generator(device="cuda").manual_seed(500 + rank)
            t = (torch.rand(numel, device="cuda", generator=g) * 2 - 1).to(tdt)
            exact = t.double().clone()
            dist.all_reduce(exact)
            inputs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(inputs, t)
            pd.quantized_all_reduce_(t, dtype=qdt, ctx=ctx, transport="p2p", round_mode=rmode, algorithm="direct", lanes=lanes)
            key = f"direct_{tdt}_{qdt}_{rmode}_{numel}_lanes{lanes}"
            if rmode == "nearest":
                want = _direct_on_the_oracle(orc, pd, [i.cpu() for i in inputs], qdt, lanes)
                res[key + "_bit_exact_vs_oracle"] = (bool(np.array_equal(_bits(t.cpu()), want)),)
            gathered = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(gathered, t)
            identical = all(torch.equal(gathered[0], gi) for gi in gathered)
            levels = {torch.quint8: 255, torch.quint4x2: 15, torch.quint2x4: 3}[qdt]
            step = 2.0 * world / levels
            err = (t.double() - exact).abs().max().item()
            # every input is rounded once (<= step/2 each, at the input's own scale 2/levels), the sum once more
            bound = (0.5 if rmode == "nearest" else 1.0) * ((world - 1) * 2.0 / levels + step) + (0.05 * world if tdt == torch.bfloat16 else 1e-5)
            res[key] = (identical, err <= bound)
        # the direct algorithm with its exchanges handed to NCCL (all_to_all_single + all_gather_into_tensor): same bits
        for tdt, qdt, numel in ((torch.float32, torch.quint8, 1_000_003), (torch.bfloat16, torch.quint4x2, 300_007),
                                (torch.float32, torch.quint2x4, 70_001), (torch.float32, torch.quint8, 100)):
            g = torch.Generator(device="cuda").manual_seed(300 + rank)
            t = (torch.rand(numel, device="cuda", generator=g) * 2 - 1).to(tdt)
            inputs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(inputs, t)
            pd.quantized_all_reduce_(t, dtype=qdt, ctx=ctx, transport="nccl", algorithm="direct")
            want = _direct_on_the_oracle(orc, pd, [i.cpu() for i in inputs], qdt, 1)
            res[f"direct_over_nccl_{tdt}_{qdt}_{numel}"] = (bool(np.array_equal(_bits(t.cpu()), want)),)
        # the same collective captured into a CUDA graph for a persistent tensor, replayed with new contents
        for tdt, qdt, numel, lanes in ((torch.float32, torch.quint8, 1_000_003, 2), (torch.bfloat16, torch.quint4x2, 300_007, 1)):
            t = torch.zeros(numel, device="cuda", dtype=tdt)
            plan = pd.QuantizedAllReduce(t, dtype=qdt, ctx=ctx, lanes=lanes)
            for rep in range(3):
                g = torch.Generator(device="cuda").manual_seed(900 + 10 * rep + rank)
                t.copy_((torch.rand(numel, device="cuda", generator=g) * 2 - 1).to(tdt))
                inputs = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(inputs, t)
                plan()
                want = _direct_on_the_oracle(orc, pd, [i.cpu() for i in inputs], qdt, lanes)
                res[f"graph_replay_{tdt}_{qdt}_{rep}"] = (bool(np.array_equal(_bits(t.cpu()), want)),)
        for tdt, qdt, transport, rmode, lanes in ((torch.float32, torch.quint8, "nccl", "nearest", 1), (torch.bfloat16, torch.quint8, "nccl", "nearest", 1),
                                                  (torch.float32, torch.quint4x2, "nccl", "nearest", 1), (torch.float32, torch.quint8, "p2p", "nearest", 1),
                                                  (torch.bfloat16, torch.quint4x2, "p2p", "nearest", 1), (torch.float32, torch.quint8, "p2p", "nearest", 1),
                                                  (torch.float32, torch.quint8, "p2p", "stochastic_per_element", 1),
                                                  (torch.float32, torch.quint4x2, "auto", "nearest", 1), (torch.bfloat16, torch.quint8, "auto", "nearest", 1),
                                                  (torch.float32, torch.quint8, "p2p", "nearest", 2), (torch.bfloat16, torch.quint8, "p2p", "nearest", 3),
                                                  (torch.float32, torch.quint8, "nccl", "stochastic_per_element", 1)):
            tol_steps = 1.0 if rmode == "nearest" else 2.0          # per-element SR: up to one step per hop instead of half a step
            g = torch.Generator(device="cuda").manual_seed(100 + rank)
            t = (torch.rand(1_000_003, device="cuda", generator=g) * 2 - 1).to(tdt)
            exact = t.double().clone()
            dist.all_reduce(exact)
            inputs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(inputs, t)
            pd.quantized_all_reduce_(t, dtype=qdt, ctx=ctx, transport=transport, round_mode=rmode, lanes=lanes, algorithm="ring")
            if rmode == "nearest":
                # the whole collective replayed on the CPU with the oracle, hop by hop: the GPU result must be bit-identical
                want = _ring_on_the_oracle(orc, pd, [i.cpu() for i in inputs], qdt, lanes)
                res[f"ring_bit_exact_{transport}_{tdt}_{qdt}_{len(res)}"] = (bool(np.array_equal(_bits(t.cpu()), want)),)
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
            res[f"ring_{transport}_{tdt}_{qdt}_{rmode}_lanes{lanes}_{len(res)}"] = ok
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def _bits(t):
    import torch
    return t.contiguous().view(torch.uint8).numpy().copy()


def _host_views(orc, inputs, qdt):
    import torch
    is_bf16 = inputs[0].dtype == torch.bfloat16
    odt = {torch.quint8: orc.UINT8, torch.quint4x2: orc.UINT4, torch.quint2x4: orc.UINT2}[qdt]
    host = [(i.view(torch.int16).numpy().view(np.uint16).copy() if is_bf16 else i.numpy().copy()) for i in inputs]
    return host, odt, (orc.BF16 if is_bf16 else orc.F32)


def _direct_on_the_oracle(orc, pd, inputs, qdt, lanes=1):
    """the collective replayed on the CPU with nothing but the oracle's three functions (oracle/replay.py)"""
    from oracle import replay
    host, odt, fdt = _host_views(orc, inputs, qdt)
    return replay.direct_all_reduce(host, odt, fdt, pd.shard_bounds, pd.SHARD_ALIGN, lanes)


def _ring_on_the_oracle(orc, pd, inputs, qdt, lanes=1):
    from oracle import replay
    host, odt, fdt = _host_views(orc, inputs, qdt)
    return replay.ring_all_reduce(host, odt, fdt, pd.shard_bounds, pd.SHARD_ALIGN, lanes)


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
