#!/usr/bin/env python
"""What does one hop of the quantized ring cost, kernel by kernel?  (2 GPUs: torchrun --nproc-per-node 2 tools/ring_probe.py)

Times, with CUDA events on rank 0 while rank 1 does the same, for the chunk of an 8-GPU ring over 2^28 f32 (33.5 M elements):
  quantize f32->u8 into LOCAL memory / into the neighbour's slot over NVLink, direct kernels and TMA (bulk-store) kernels
  dequantize-ADD + min/max (the fused accumulate)
  dequantize-SET + forward to LOCAL / PEER memory
  a plain copy of the packed payload to the peer (copy engine / copy kernel), the symmetric-memory barrier
"""
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
from piquant import DataType as D, ReduceOp, RoundMode  # noqa: E402


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem

    n = (1 << 28) // 8
    meta = 64
    slot = (meta + n + 255) // 256 * 256
    buf = symm_mem.empty(2 * slot, dtype=torch.uint8, device=dev)
    hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
    peer = int(hdl.buffer_ptrs[(rank + 1) % world])
    mine = buf.data_ptr()
    ctx = piquant.Context()
    st = torch.cuda.current_stream().cuda_stream
    LOCAL, REV = piquant.Context.FLAG_LOCAL, piquant.Context.FLAG_REVERSE
    xs = [torch.empty(n, dtype=torch.float32, device=dev).uniform_(-1, 1) for _ in range(6)]
    keep = torch.empty(slot, dtype=torch.uint8, device=dev)
    m = torch.zeros(64, dtype=torch.uint8, device=dev)
    m2 = torch.zeros(64, dtype=torch.uint8, device=dev)
    ctx.compute_meta_on_stream(xs[0].data_ptr(), D.F32, n, D.UINT8, m.data_ptr(), LOCAL, local, st)
    buf[:meta].copy_(m)
    keep[:meta].copy_(m)
    ctx.quantize_meta_on_stream(xs[0].data_ptr(), D.F32, mine + meta, D.UINT8, n, RoundMode.NEAREST, m.data_ptr(), 0, local, st)
    torch.cuda.synchronize()
    dist.barrier()

    def ev(fn, reps=20):
        for k in range(3):
            fn(k)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(reps):
            fn(3 + k)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rows = []
    for variant, vname in ((1, "direct"), (2, "tma")):
        ctx.set_kernel_variant(variant)
        rows.append((f"quantize -> local   [{vname}]", ev(lambda k: ctx.quantize_meta_on_stream(xs[k % 6].data_ptr(), D.F32, keep.data_ptr() + meta, D.UINT8, n, RoundMode.NEAREST, m.data_ptr(), 0, local, st)), 5 * n))
        rows.append((f"quantize -> PEER    [{vname}]", ev(lambda k: ctx.quantize_meta_on_stream(xs[k % 6].data_ptr(), D.F32, peer + meta, D.UINT8, n, RoundMode.NEAREST, m.data_ptr(), 0, local, st)), 5 * n))
    ctx.set_kernel_variant(0)
    rows.append(("min/max + params (first hop)", ev(lambda k: ctx.compute_meta_on_stream(xs[k % 6].data_ptr(), D.F32, n, D.UINT8, m2.data_ptr(), LOCAL, local, st)), 4 * n))
    rows.append(("dequantize-ADD + min/max + params (fused)", ev(lambda k: ctx.dequantize_add_minmax_on_stream(mine + meta, D.UINT8, xs[k % 6].data_ptr(), D.F32, n, mine, D.UINT8, m2.data_ptr(), 0, local, st)), 9 * n))
    rows.append(("dequantize-ADD (plain)", ev(lambda k: ctx.dequantize_meta_on_stream(mine + meta, D.UINT8, xs[k % 6].data_ptr(), D.F32, n, ReduceOp.ADD, mine, local, st)), 9 * n))
    rows.append(("dequantize-SET (plain)", ev(lambda k: ctx.dequantize_meta_on_stream(mine + meta, D.UINT8, xs[k % 6].data_ptr(), D.F32, n, ReduceOp.SET, mine, local, st)), 5 * n))
    rows.append(("dequantize-SET + forward -> local", ev(lambda k: ctx.dequantize_forward_on_stream(mine + meta, D.UINT8, xs[k % 6].data_ptr(), D.F32, n, mine, keep.data_ptr() + meta, keep.data_ptr(), local, st)), 6 * n))
    rows.append(("dequantize-SET + forward -> PEER", ev(lambda k: ctx.dequantize_forward_on_stream(mine + meta, D.UINT8, xs[k % 6].data_ptr(), D.F32, n, mine, peer + slot + meta, peer + slot, local, st)), 6 * n))
    peer_t = hdl.get_buffer((rank + 1) % world, (2 * slot,), torch.uint8)
    rows.append(("copy packed payload -> PEER (torch copy_)", ev(lambda k: peer_t[slot:slot + n].copy_(keep[:n], non_blocking=True)), n))
    rows.append(("symmetric-memory barrier", ev(lambda k: hdl.barrier(channel=0)), 0))
    # quantize -> PEER overlapped with the fused accumulate of another chunk on a second stream (what two lanes do)
    side = torch.cuda.Stream()

    def overlapped(k):
        side.wait_stream(torch.cuda.current_stream())
        ctx.quantize_meta_on_stream(xs[k % 3].data_ptr(), D.F32, peer + meta, D.UINT8, n, RoundMode.NEAREST, m.data_ptr(), 0, local, st)
        ctx.dequantize_add_minmax_on_stream(mine + slot + meta, D.UINT8, xs[3 + k % 3].data_ptr(), D.F32, n, mine, D.UINT8, m2.data_ptr(), 0, local, side.cuda_stream)
        torch.cuda.current_stream().wait_stream(side)
    buf[slot:slot + meta].copy_(m)
    rows.append(("quantize -> PEER  ||  fused accumulate (two streams)", ev(overlapped), 14 * n))
    if rank == 0:
        print(f"chunk = {n} f32 elements ({4 * n / 1e6:.1f} MB), packed {n / 1e6:.1f} MB; times are the max over the {world} ranks")
        for name, us, nbytes in rows:
            print(f"{name:58s} {us:9.2f} us   {nbytes / us / 1e3 if us else 0:8.1f} GB/s (algorithmic)   link {n / us / 1e3:7.1f} GB/s")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
