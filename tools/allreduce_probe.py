#!/usr/bin/env python
"""Quantized all-reduce variants next to NCCL's f32 all-reduce (torchrun --nproc-per-node N tools/allreduce_probe.py [log2 numel]).
CUDA events, max over ranks; development tool -- the judged numbers come from bench.py's quantized_ring_all_reduce leg."""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import piquant  # noqa: E402
from piquant import distributed as pd  # noqa: E402


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = piquant.Context()
    sizes = [int(a) for a in sys.argv[1:]] or [28]
    for lg in sizes:
        n = 1 << lg
        base = torch.empty(n, dtype=torch.float32, device=dev).uniform_(-1, 1)
        work = torch.empty_like(base)

        def timed(fn, reps=8):
            for _ in range(3):
                work.copy_(base)
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            tot = 0.0
            for _ in range(reps):
                work.copy_(base)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            t = torch.tensor([tot / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        rows = [("nccl f32 all_reduce", timed(lambda: dist.all_reduce(work)))]
        exact = work.clone()
        for name, kw in (("ring p2p, 2 lanes", dict(transport="p2p", algorithm="ring", lanes=2)),
                         ("direct, 1 lane", dict(transport="p2p", algorithm="direct", lanes=1)),
                         ("direct, 2 staggered lanes", dict(transport="p2p", algorithm="direct", lanes=2)),
                         ("direct u4, 1 lane", dict(transport="p2p", algorithm="direct", lanes=1, dtype=torch.quint4x2)),
                         ):
            kw = dict(kw)
            qd = kw.pop("dtype", torch.quint8)
            ms = timed(lambda: pd.quantized_all_reduce_(work, dtype=qd, ctx=ctx, **kw))
            err = (work - exact).abs().max()
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            rows.append((name, ms, float(err.item())))
        # where the time goes: CUDA events at the phase boundaries of one eager call (rank 0's view)
        plan = pd._DirectPlan(n, torch.float32, torch.quint8, dev, None, ctx, piquant.RoundMode.NEAREST, 1)
        for rep in range(3):
            work.copy_(base)
            torch.cuda.synchronize()
            dist.barrier()
            plan.trace = [] if rep == 2 else None
            plan.enqueue(work)
            torch.cuda.synchronize()
        if rank == 0:
            t0 = plan.trace[0][2]
            print("  timeline of one eager direct all-reduce (1 lane), us since start, rank 0:")
            for lane, label, ev in plan.trace:
                print(f"    {t0.elapsed_time(ev) * 1e3:9.1f}  {label}")
        # what the links give: every rank copies a packed chunk to `fan` different peers at once, one stream per peer
        import torch.distributed._symmetric_memory as symm_mem
        nb = n // world
        buf = symm_mem.empty(world * nb, dtype=torch.uint8, device=dev)
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        src = torch.empty(nb, dtype=torch.uint8, device=dev)
        streams = [torch.cuda.Stream() for _ in range(world - 1)]
        for fan in (1, 2, 3, 4, 7):
            if fan > world - 1:
                continue
            def burst():
                cur = torch.cuda.current_stream()
                for k in range(world - 1):
                    st = streams[k % fan]
                    if k < fan:
                        st.wait_stream(cur)
                    peer = (rank + 1 + k) % world
                    ctx.copy_on_stream(int(hdl.buffer_ptrs[peer]) + rank * nb, src.data_ptr(), nb, local, st.cuda_stream)
                for st in streams[:fan]:
                    cur.wait_stream(st)
            ms = timed(burst, reps=5)
            if rank == 0:
                print(f"  link probe: {world - 1} copies of {nb / 1e6:.1f} MB to {world - 1} peers, {fan} in flight: {ms * 1e3:8.1f} us  "
                      f"= {(world - 1) * nb / ms / 1e6:7.1f} GB/s out per GPU (every GPU sending and receiving)")
        # NVSwitch multicast: ONE copy-engine transfer to the multicast address lands in every rank's buffer at the same offset
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if rank == 0:
            print(f"  multicast_ptr = {mc:#x}  (has_multicast_support: {getattr(type(hdl), 'has_multicast_support', None)})")
        if mc:
            try:
                buf.zero_()
                src.fill_(rank + 1)
                torch.cuda.synchronize()
                dist.barrier()
                ctx.copy_on_stream(mc + rank * nb, src.data_ptr(), nb, local, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                dist.barrier()
                ok = all(int(buf[k * nb].item()) == k + 1 and int(buf[(k + 1) * nb - 1].item()) == k + 1 for k in range(world))
                okt = torch.tensor([int(ok)], device=dev)
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)

                def mc_burst():
                    ctx.copy_on_stream(mc + rank * nb, src.data_ptr(), nb, local, torch.cuda.current_stream().cuda_stream)
                ms = timed(mc_burst, reps=5)
                if rank == 0:
                    print(f"  multicast probe: every rank broadcasts {nb / 1e6:.1f} MB through the multicast address with ONE copy-engine transfer: "
                          f"{ms * 1e3:8.1f} us = {(world - 1) * nb / ms / 1e6:7.1f} GB/s INTO every GPU; all replicas correct: {bool(okt.item())}")
            except Exception as e:      # noqa: BLE001
                if rank == 0:
                    print(f"  multicast probe failed: {type(e).__name__}: {str(e)[:200]}")
        del buf, hdl
        # A/B in one run: gather exchange by seven unicast copies or through the NVSwitch multicast address (pieces + flags)
        for rep in range(2):
            for mcast in (False, True):
                plan = pd._DirectPlan(n, torch.float32, torch.quint8, dev, None, ctx, piquant.RoundMode.NEAREST, 1, mcast)
                name = "multicast" if plan.multicast else "unicast"
                ms = timed(lambda: plan.enqueue(work))
                err = (work - exact).abs().max()
                dist.all_reduce(err, op=dist.ReduceOp.MAX)
                rows.append((f"direct eager, {name} gather (rep {rep})", ms, float(err.item())))
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    plan.enqueue(work)
                ms = timed(g.replay)
                err = (work - exact).abs().max()
                dist.all_reduce(err, op=dist.ReduceOp.MAX)
                rows.append((f"direct graph, {name} gather (rep {rep})", ms, float(err.item())))
                if rep == 1 and plan.multicast:
                    work.copy_(base)
                    torch.cuda.synchronize()
                    dist.barrier()
                    plan.trace = []
                    plan.enqueue(work)
                    torch.cuda.synchronize()
                    if rank == 0:
                        t0 = plan.trace[0][2]
                        print("  timeline of one eager direct all-reduce with the multicast gather, us since start, rank 0:")
                        for lane, label, ev in plan.trace:
                            print(f"    {t0.elapsed_time(ev) * 1e3:9.1f}  {label}")
                    plan.trace = None
                del g, plan
        for lanes in (1,):
            for qd, qn in ((torch.quint8, "u8"), (torch.quint4x2, "u4")):
                plan = pd.QuantizedAllReduce(work, dtype=qd, ctx=ctx, lanes=lanes)
                ms = timed(plan)
                err = (work - exact).abs().max()
                dist.all_reduce(err, op=dist.ReduceOp.MAX)
                rows.append((f"direct {qn} CUDA GRAPH, {lanes} staggered lanes", ms, float(err.item())))
                del plan
        if rank == 0:
            print(f"numel = 2^{lg} f32 ({4 * n / 1e6:.0f} MB), {world} GPUs")
            for r in rows:
                extra = f"   {rows[0][1] / r[1]:5.2f}x nccl   max_abs_err {r[2]:.4f}" if len(r) > 2 else ""
                print(f"  {r[0]:48s} {r[1] * 1e3:9.1f} us{extra}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
