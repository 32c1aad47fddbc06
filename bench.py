#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200 pi-quant library.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): Gelem/s of f32 -> uint8 nearest-rounding quantize with a per-tensor scale.
Workload: numel = 1e9 elements PER GPU (the size the metric's roofline target is quoted on), x ~ U(-1,1),
(scale, zero_point) from compute_quant_params.  One step = one quantize pass over the rank's shard
= ONE kernel launch through the C ABI (piquant_quantize on device pointers).  Shards are independent
(no collective on the data path), so N GPUs is weak scaling: value = N * 1e9 / max-over-ranks step time.

Printed JSON keys beyond the base contract:
  roofline      dominant kernel vs the measured HBM copy peak (MEASURED_PEAKS.json); algorithmic 5 B/elem
  e2e           same metric through piquant_quantize with PINNED HOST buffers: H2D + kernel + D2H per step
  cpu_baseline  the unmodified reference (oracle/_ref/libpiquant_ref.so) on this box's host cores, N=1 only
  extra         the other BASELINE configs, timed the same way (kernel-only, per GPU); the reference's own small-tensor
                benchmark recipe; pageable-host throughput; per-config byte comparison with the CPU reference
  strong        run on ALL ranks at every N: the BASELINE configs as the contract spells them -- ONE tensor sharded over the
                N GPUs (C3 f32[1e9] compute_quant_params incl. the cross-rank exchange, C5 u8[1e9]->f32 ADD, f32->u8 and
                bf16->u4 at 27.264 M and 1e9 total) with per-rank parity against the CPU oracle in the same run
--impl reference times the reference's own CPU implementation (all host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "pi-quant_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "f32->uint8 nearest quantize throughput"
UNIT = "Gelem/s"
NUMEL = int(os.environ.get("PIQUANT_BENCH_NUMEL", 1_000_000_000))
BYTES_PER_ELEM = 5.0          # 4 B read + 1 B written


def measured_peak() -> tuple[float, str]:
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"f32->uint8 nearest-round quantize, per-tensor scale, numel={NUMEL} per GPU (BASELINE metric at its 1e9 size)",
        "numel_per_gpu": NUMEL,
        "shards": n_gpus,
        "input": "x ~ U(-1,1) f32, scale/zero_point from compute_quant_params",
        "l2": "inputs larger than L2: 5 GB of traffic per step vs 126 MB L2, no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 8:
                self.rows.append(parts)

    def mark(self) -> int:
        return len(self.rows)

    def stop(self) -> None:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, lo: int, hi: int, window: str) -> dict:
        rows = self.rows[lo:hi] or self.rows
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "window": "nvidia-smi unavailable"}
        busy = [r for r in rows if r[3].isdigit() and int(r[3]) >= 50] or rows
        sm = [float(r[0]) for r in busy if r[0].replace(".", "").isdigit()]
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[4 + i].lower().startswith("active") for r in busy):
                reasons.append(name)
        power = [float(r[2]) for r in busy if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(rows[0][1]) if rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(busy), "power_w_max": max(power) if power else None, "window": window}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def bind_host_to_gpu_numa_node(torch, index: int) -> dict:
    """Pin this process (and so the pages of the pinned host buffers it allocates next) to the NUMA node the GPU hangs off.
    On a multi-socket 8-GPU box eight ranks copying from one node's memory share that node's DRAM and the inter-socket link
    (measured in round 1: 21 GB/s per GPU instead of 55).  Best effort: any failure leaves the affinity untouched."""
    info: dict = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text().strip())
        info.update({"pci": bdf, "node": node})
        nodes = [d for d in Path("/sys/devices/system/node").glob("node[0-9]*")]
        info["nodes"] = len(nodes)
        if node < 0 or len(nodes) < 2:
            return info
        cpus: set = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update({"bound": True, "cpus": len(cpus)})
    except Exception as e:      # noqa: BLE001 -- never let topology probing break the benchmark
        info["error"] = repr(e)[:120]
    return info


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import piquant
    from piquant import DataType as D, ReduceOp, RoundMode

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path behind libpiquant.so)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = piquant.Context()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    n = NUMEL
    gen = torch.Generator(device=dev).manual_seed(rank)
    x = torch.empty(n, dtype=torch.float32, device=dev).uniform_(-1, 1, generator=gen)
    q = torch.empty(n, dtype=torch.uint8, device=dev)
    scale, zp = ctx.compute_quant_params_ptr_float32(x.data_ptr(), D.UINT8, n)

    def step() -> None:
        ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, scale, zp, RoundMode.NEAREST)

    sampler = ClockSampler(local) if rank == 0 else None
    # W untimed warm-up steps, then keep the same loop running until 1 s has passed: a power-capped B200 needs
    # several hundred ms under load before clocks settle (throughput drifts by +-3 % until then, profiles/r1_placement_probe.txt)
    warm_steps, t_w = 0, time.perf_counter()
    while warm_steps < max(args.warmup, 3) or time.perf_counter() - t_w < 1.0:
        step()
        warm_steps += 1
        if warm_steps % 64 == 0:
            torch.cuda.synchronize()
    barrier()
    launches0 = ctx.kernel_launches
    mark0 = sampler.mark() if sampler else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.kernel_launches - launches0
    window = "timed region"
    if sampler and total_ms < 600.0:
        # the timed region is shorter than a few nvidia-smi periods: keep the identical loop running (untimed)
        # so that the clock samples are taken under the same load
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        window = "timed region + 1.0 s untimed continuation of the same step loop"
    mark1 = sampler.mark() if sampler else 0
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with pinned host buffers ---------------------------------
    affinity0 = os.sched_getaffinity(0)
    numa = bind_host_to_gpu_numa_node(torch, local)
    xh = torch.empty(n, dtype=torch.float32).pin_memory()
    qh = torch.empty(n, dtype=torch.uint8).pin_memory()
    xh.copy_(x)
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step() -> None:       # synchronous: H2D chunks | kernel | D2H chunks, overlapped inside the library
        ctx.quantize_ptr(xh.data_ptr(), D.F32, qh.data_ptr(), D.UINT8, n, scale, zp, RoundMode.NEAREST)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    barrier()
    e2e_value = world * n / e2e_s / 1e9
    # what the link itself can do: a plain pinned -> device copy of the same 4 GB (and device -> pinned of the 1 GB)
    t0 = time.perf_counter()
    for _ in range(2):
        x.copy_(xh, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbps = 2 * 4 * n / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter()
    for _ in range(2):
        qh.copy_(q, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbps = 2 * n / (time.perf_counter() - t0) / 1e9
    e2e_ok = bool(torch.equal(qh[: 1 << 24], q[: 1 << 24].cpu()))
    # ... and what the link does when both directions run at once on every rank (the pipeline's actual traffic pattern:
    # 4 B/element up and 1 B/element down share the host's memory system with the other ranks' copies)
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        with torch.cuda.stream(up):
            x.copy_(xh, non_blocking=True)
        with torch.cuda.stream(down):
            qh.copy_(q, non_blocking=True)          # 1 byte down per 4 bytes up, the pipeline's own ratio
    torch.cuda.synchronize()
    bidir_s = max_over_ranks(time.perf_counter() - t0)
    bidir_h2d_gbps = 2 * 4 * n / bidir_s / 1e9
    barrier()
    # the literal drop-in case: PAGEABLE host tensors (what piquant.torch callers of the reference pass, reference
    # python/src/piquant/torch.py:87,117) -- copy workers <-> pinned bounce buffers <-> copy engines
    n_pg = min(n, 1 << 28)
    xp = torch.empty(n_pg, dtype=torch.float32)
    xp.copy_(xh[:n_pg])
    qp = torch.empty(n_pg, dtype=torch.uint8)
    ctx.quantize_ptr(xp.data_ptr(), D.F32, qp.data_ptr(), D.UINT8, n_pg, scale, zp, RoundMode.NEAREST)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.quantize_ptr(xp.data_ptr(), D.F32, qp.data_ptr(), D.UINT8, n_pg, scale, zp, RoundMode.NEAREST)
    pageable_s = max_over_ranks((time.perf_counter() - t0) / 3)
    pageable_ok = bool(torch.equal(qp[: 1 << 24], qh[: 1 << 24]))
    barrier()
    del xp, qp
    os.sched_setaffinity(0, affinity0)      # the CPU baseline below must see every host core again

    # ---- the other BASELINE configs, kernel-only, this rank's GPU -------------------------------
    extra = {}
    if rank == 0:
        extra = run_extra(torch, ctx, D, RoundMode, ReduceOp, x, q, scale, zp, dev)
    barrier()
    strong = strong_scaling_legs(torch, dist, piquant, D, RoundMode, ReduceOp, x, q, world, rank, dev, barrier, max_over_ranks)
    if world > 1:
        ring = ring_allreduce_leg(torch, dist, ctx, world, rank, dev)
        if rank == 0:
            extra["quantized_ring_all_reduce"] = ring

    # ---- CPU baseline beside it (rank 0, N=1) ----------------------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1:
        cpu_baseline, parity = cpu_reference_leg(xh.numpy(), qh.numpy(), scale, zp, ctx, D, RoundMode, ReduceOp)

    if sampler:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_ELEM * n / (ms_per_step * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get("quantize_f32_u8_1e9_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "warmup_steps_run": warm_steps,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32->u8 (f32 multiply/add, int32 clamp)", "data": "synthetic U(-1,1), seeded per rank",
        "config": workload_config(world),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": BYTES_PER_ELEM * n,
                     "kernel": "quantize f32->u8 (one launch per step)", "of_nominal_8000": round(achieved / 8000.0, 4)},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": 4 * n * world, "d2h_bytes_per_step": n * world,
                "steps": e2e_steps, "path": "piquant_quantize(host pinned in, host pinned out): chunked H2D | kernel | D2H pipeline",
                "bound": "pcie", "h2d_GBps_achieved_per_gpu": round(4 * n / e2e_s / 1e9, 1), "h2d_GBps_plain_memcpy": round(h2d_gbps, 1),
                "d2h_GBps_plain_memcpy": round(d2h_gbps, 1), "frac_of_link": round((4 * n / e2e_s / 1e9) / h2d_gbps, 4),
                "h2d_GBps_plain_memcpy_both_directions_busy": round(bidir_h2d_gbps, 1),
                "frac_of_link_both_directions_busy": round((4 * n / e2e_s / 1e9) / bidir_h2d_gbps, 4),
                "output_matches_device_path": e2e_ok, "host_numa": numa,
                "pageable_host": {"value": round(world * n_pg / pageable_s / 1e9, 3), "unit": UNIT, "numel_per_gpu": n_pg,
                                  "h2d_GBps_per_gpu": round(4 * n_pg / pageable_s / 1e9, 1), "output_matches_pinned_path": pageable_ok,
                                  "path": "piquant_quantize(PAGEABLE host in, PAGEABLE host out): host copy workers <-> pinned bounce ring <-> H2D | kernel | D2H",
                                  "note": "what a caller of the reference's piquant.torch passes (CPU torch tensors)"}},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(mark0, mark1, window),
        "cpu_baseline": cpu_baseline,
        "parity_vs_cpu_reference": parity,
        "strong": strong,
        "extra": extra,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def time_launches(torch, fn, reps: int, warm: int = 3) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def run_extra(torch, ctx, D, RoundMode, ReduceOp, x, q, scale, zp, dev) -> dict:
    peak, _ = measured_peak()
    out = {}

    def rec(name, numel, bytes_per_elem, t, note=""):
        gbs = bytes_per_elem * numel / t / 1e9
        out[name] = {"numel": numel, "ms": round(t * 1e3, 4), "Gelem/s": round(numel / t / 1e9, 1), "GB/s": round(gbs, 1),
                     "frac_of_measured_peak": round(gbs / peak, 4), "note": note}

    n = x.numel()
    # C1: the reference's own benchmark size; 3 rotating buffer pairs (409 MB) so that L2 (126 MB) cannot serve it
    n1 = 27_264_000
    if n >= 3 * n1:
        xs = [x[i * n1:(i + 1) * n1] for i in range(3)]
        qs = [q[i * n1:(i + 1) * n1] for i in range(3)]
        state = {"i": 0}

        def c1():
            i = state["i"] = (state["i"] + 1) % 3
            ctx.quantize_ptr(xs[i].data_ptr(), D.F32, qs[i].data_ptr(), D.UINT8, n1, scale, zp, RoundMode.NEAREST)
        rec("C1_f32_u8_nearest_27.264M", n1, 5, time_launches(torch, c1, 60, 6), "3 rotating buffer pairs, back-to-back launches")
        rec("C1_f32_u8_nearest_27.264M_hot_L2", n1, 5, time_launches(torch, lambda: ctx.quantize_ptr(xs[0].data_ptr(), D.F32, qs[0].data_ptr(), D.UINT8, n1, scale, zp, RoundMode.NEAREST), 60, 6),
            "the SAME buffer pair every launch: 136 MB of traffic against a 126 MB L2, partly L2-resident -- not an HBM number")
        # the same 60 launches captured once into a CUDA graph and replayed: no host launch cost at all (report both)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                for _ in range(60):
                    c1()
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            rec("C1_f32_u8_nearest_27.264M_cuda_graph_replay", n1, 5, time_launches(torch, graph.replay, 5, 2) / 60,
                "60 launches (3 rotating buffer pairs) captured into one CUDA graph, replayed")
            del graph
        except Exception as e:      # noqa: BLE001
            ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            out["C1_f32_u8_nearest_27.264M_cuda_graph_replay"] = {"error": repr(e)[:200]}
        # the only figure the reference publishes for this path: a README bar chart, ~1.7 s per 1000 runs at this size on an
        # EPYC 9654 (~16 Gelem/s, read off the chart, +-5 %).  Other hardware, so context, not vs_baseline.
        out["C1_f32_u8_nearest_27.264M"]["reference_readme_chart_Gelem/s"] = 16.0
        out["C1_f32_u8_nearest_27.264M"]["ratio_to_readme_chart"] = round(out["C1_f32_u8_nearest_27.264M"]["Gelem/s"] / 16.0, 1)
    # C2: bf16 -> quint4x2 quantize and dequantize round trip, numel = 1e8
    n2 = min(100_000_000, n)
    xb = x[:n2].to(torch.bfloat16)
    q4 = torch.empty((n2 + 1) // 2, dtype=torch.uint8, device=dev)
    yb = torch.empty(n2, dtype=torch.bfloat16, device=dev)
    s4, z4 = ctx.compute_quant_params_ptr_bfloat16(xb.data_ptr(), D.UINT4, n2)
    rec("C2_bf16_u4_quantize_1e8", n2, 2.5, time_launches(torch, lambda: ctx.quantize_ptr(xb.data_ptr(), D.BF16, q4.data_ptr(), D.UINT4, n2, s4, z4, RoundMode.NEAREST), 20),
        "250 MB per launch: partly L2-resident on a 126 MB L2")
    rec("C2_u4_bf16_dequantize_1e8", n2, 2.5, time_launches(torch, lambda: ctx.dequantize_ptr(q4.data_ptr(), D.UINT4, yb.data_ptr(), D.BF16, n2, s4, z4, ReduceOp.SET), 20))
    err = (yb.float() - xb.float()).abs().max().item()
    out["C2_round_trip_max_abs_err_over_scale"] = round(err / s4, 4)
    del xb, q4, yb
    # C3: compute_quant_params (min/max reduce, synchronous: returns host scalars)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        ctx.compute_quant_params_ptr_float32(x.data_ptr(), D.UINT8, n)
    rec("C3_compute_quant_params_f32", n, 4, (time.perf_counter() - t0) / reps, "wall clock, includes the stream sync and 16 B D2H")
    # one-shot: compute_quant_params + quantize with device-resident parameters (three kernels, ONE sync) against the
    # two separate calls (two syncs); at the reference's benchmark size the tensor (109 MB) fits the 126 MB L2, so the
    # quantize pass re-reads it from L2, not HBM
    n1 = min(27_264_000, n)
    x1, q1 = x[:n1], q[:n1]

    def two_step():
        s_, z_ = ctx.compute_quant_params_ptr_float32(x1.data_ptr(), D.UINT8, n1)
        ctx.quantize_ptr(x1.data_ptr(), D.F32, q1.data_ptr(), D.UINT8, n1, s_, z_, RoundMode.NEAREST)
        torch.cuda.synchronize()

    def one_shot():
        ctx.quantize_auto_ptr(x1.data_ptr(), D.F32, q1.data_ptr(), D.UINT8, n1, RoundMode.NEAREST)

    for name, fn in (("params_then_quantize_two_calls_27.264M", two_step), ("quantize_auto_one_shot_27.264M", one_shot)):
        for _ in range(5):
            fn()
        t0 = time.perf_counter()
        for _ in range(50):
            fn()
        t = (time.perf_counter() - t0) / 50
        out[name] = {"numel": n1, "us": round(t * 1e6, 2), "Gelem/s": round(n1 / t / 1e9, 1), "note": "wall clock incl. synchronisation; min/max + params + quantize"}
    # per-launch distribution of the headline kernel (one event pair per launch) and the cost of a call on a tiny tensor
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    for e0, e1 in evs:
        e0.record()
        ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, scale, zp, RoundMode.NEAREST)
        e1.record()
    torch.cuda.synchronize()
    per = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    out["headline_per_launch_ms"] = {"best": round(per[0], 4), "median": round(per[len(per) // 2], 4), "worst": round(per[-1], 4),
                                     "best_GBps": round(5 * n / (per[0] * 1e-3) / 1e9, 1), "median_GBps": round(5 * n / (per[len(per) // 2] * 1e-3) / 1e9, 1)}
    tiny = 4096
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, tiny, scale, zp, RoundMode.NEAREST)
    t_issue = (time.perf_counter() - t0) / 2000
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / 2000
    out["call_overhead_us_numel_4096"] = {"host_issue": round(t_issue * 1e6, 2), "throughput_back_to_back": round(t_total * 1e6, 2),
                                         "note": "reference ABI (piquant_quantize) FROM PYTHON: cffi call + pointer classification (2 driver queries) + cudaLaunchKernelEx (PDL); the same call from C costs 2.2 us, an empty <<<>>> launch 2.3 us (tools/call_overhead.cu, profiles/r2_call_overhead_c_abi.txt)"}
    # the same call with device and stream passed along (piquant_cuda_quantize_on_stream: what piquant.torch uses): no classification
    dev_i, st_i = dev.index, torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        ctx.quantize_on_stream(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, tiny, scale, zp, RoundMode.NEAREST, dev_i, st_i)
    t_issue = (time.perf_counter() - t0) / 2000
    torch.cuda.synchronize()
    t_total = (time.perf_counter() - t0) / 2000
    out["call_overhead_us_numel_4096_on_stream"] = {"host_issue": round(t_issue * 1e6, 2), "throughput_back_to_back": round(t_total * 1e6, 2),
                                                   "note": "piquant_cuda_quantize_on_stream from Python: cffi call + slot lookup + cudaLaunchKernelEx (PDL)"}
    out["small_tensor_regime"] = small_tensor_leg(torch, ctx, dev)
    # C4: stochastic rounding
    rec("C4_f32_u8_stochastic", n, 5, time_launches(torch, lambda: ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.UINT8, n, scale, zp, RoundMode.STOCHASTIC), 10))
    # C4 as BASELINE.json spells it ("f32->int8"): the reference has no signed dtype at this commit (mapped to
    # uint8 above); INT8 is this library's extension (piquant_cuda.h), the same kernel on the offset-binary view
    zp_i8 = zp - 128
    rec("C4_f32_int8_stochastic_signed_extension", n, 5,
        time_launches(torch, lambda: ctx.quantize_ptr(x.data_ptr(), D.F32, q.data_ptr(), D.INT8, n, scale, zp_i8, RoundMode.STOCHASTIC), 10),
        "INT8 = extension dtype; zero point = uint8 zero point - 128")
    # C5: dequantize with ADD store op into an f32 accumulator (one 1/8 shard of 1e9 and the full size)
    n5 = max(n // 8, 1)
    acc = torch.zeros(n5, dtype=torch.float32, device=dev)
    rec("C5_u8_f32_dequantize_add_shard", n5, 9, time_launches(torch, lambda: ctx.dequantize_ptr(q.data_ptr(), D.UINT8, acc.data_ptr(), D.F32, n5, scale, zp, ReduceOp.ADD), 20),
        "one 1/8 shard of numel")
    del acc
    return out


def strong_scaling_legs(torch, dist, piquant, D, RoundMode, ReduceOp, x, q, world, rank, dev, barrier, max_over_ranks) -> dict:
    """The BASELINE configs as the contract spells them -- ONE tensor of `total` elements sharded contiguously over the N GPUs
    (piquant.distributed.shard_bounds; reference src/piquant.cpp:132-176 splits a tensor over threads the same way) -- on ALL
    ranks, with parity against the CPU oracle in the same run.  The global tensor is the concatenation of every rank's
    x[b:e] (rank-seeded U(-1,1)); shards are independent, the only exchange is {-min, max} inside compute_quant_params.
    Per config: T_N = max over ranks of the per-rank time, T_1 = rank 0 alone on the whole tensor, efficiency = T_1 / (N T_N)."""
    from piquant import distributed as pd

    peak, _ = measured_peak()
    out: dict = {"n_gpus": world, "sharding": "contiguous, boundaries at multiples of 64 elements (piquant.distributed.shard_bounds)"}
    ctx = piquant.Context()
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = [piquant.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init_rank(uid[0], world, rank)
    solo = piquant.Context()            # no communicator: rank 0 alone on the whole tensor (T_1)
    solo.set_stream(stream.cuda_stream)
    n_all = x.numel()

    def flags_all(*flags) -> bool:
        t = torch.tensor([1 if all(flags) else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def device_time(fn, reps, warm=3):
        """CUDA-event time per call of fn(k) (k = launch index, for rotating windows), max over ranks"""
        for k in range(warm):
            fn(k)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(reps):
            fn(warm + k)
        e1.record()
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1) / reps * 1e-3)

    def wall_time(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return max_over_ranks((time.perf_counter() - t0) / reps)

    def only_rank0(measure):
        """run `measure` on rank 0 while the other ranks wait; every rank gets the value"""
        v = measure() if rank == 0 else 0.0
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def oracle_windows(n_r):
        """4 Mi-element windows at both ends of this rank's shard (the shard boundaries of the global tensor)"""
        w = min(n_r, 1 << 22)
        return [(0, w)] if n_r <= (1 << 23) else [(0, w), (n_r - w, n_r)]

    from oracle import port as orc      # the CPU oracle, here ONLY as the checker of the GPU results (never timed, never the product)

    # ---- C3: compute_quant_params of f32[total] sharded, incl. the cross-rank exchange ----------------------------
    total = n_all
    b, e = pd.shard_bounds(total, world, rank)
    shard = x[b:e]
    n_r = e - b
    res = ctx.compute_quant_params_ptr_float32(shard.data_ptr(), D.UINT8, n_r)
    # parity: every rank must hold the parameters of the WHOLE tensor, bit for bit.  Independent route: torch reductions +
    # torch.distributed + the oracle's restatement of the reference's double arithmetic.
    mm = torch.stack([-shard.min(), shard.max()]) if n_r else torch.full((2,), -3.4028234663852886e38, device=dev)
    if world > 1:
        dist.all_reduce(mm, op=dist.ReduceOp.MAX)
    want = orc.params_from_minmax(float(-mm[0].item()), float(mm[1].item()), orc.UINT8)
    c3 = {"numel_total": total, "numel_per_gpu": n_r, "scale": res[0], "zero_point": res[1],
          "parity_params_bit_equal_on_every_rank": flags_all(res == want)}
    transports = {}
    if world > 1:
        for name, tr in (("peer_memory_in_kernel", 2), ("nccl_allreduce", 1)):
            if tr == 2 and ctx.comm_transport != 2:
                transports[name] = "unavailable (ranks cannot map each other's memory)"
                continue
            ctx.comm_set_transport(tr)
            t = wall_time(lambda: ctx.compute_quant_params_ptr_float32(shard.data_ptr(), D.UINT8, n_r), 30)
            ok = flags_all(ctx.compute_quant_params_ptr_float32(shard.data_ptr(), D.UINT8, n_r) == want)
            transports[name] = {"ms": round(t * 1e3, 4), "parity": ok}
        ctx.comm_set_transport(0)
        c3["exchange"] = transports
    t_n = wall_time(lambda: ctx.compute_quant_params_ptr_float32(shard.data_ptr(), D.UINT8, n_r), 30)
    t_1 = only_rank0(lambda: wall_time_local(torch, lambda: solo.compute_quant_params_ptr_float32(x.data_ptr(), D.UINT8, total), 20))
    c3.update({"ms": round(t_n * 1e3, 4), "ms_one_gpu_whole_tensor": round(t_1 * 1e3, 4), "Gelem/s": round(total / t_n / 1e9, 1),
               "strong_scaling_efficiency": round(t_1 / (world * t_n), 4), "transport": {0: "none", 1: "nccl_allreduce", 2: "peer_memory_in_kernel"}[ctx.comm_transport],
               "roofline_ms_per_gpu_at_measured_peak": round(4 * n_r / (peak * 1e9) * 1e3, 4),
               "note": "wall clock of the synchronous call (returns host scalars): min/max kernel with the exchange inside + stream sync; max over ranks"})
    out["C3_compute_quant_params_f32_1e9_sharded"] = c3
    scale_g, zp_g = res

    # ---- quantize: f32->u8 and bf16->u4, 27.264 M and 1e9 total, sharded ----------------------------------------
    xb_all = x.to(torch.bfloat16)
    for dt_name, src, DIN, DQ, odq, bpe, per in (("f32_u8", x, D.F32, D.UINT8, orc.UINT8, 5.0, 1), ("bf16_u4", xb_all, D.BF16, D.UINT4, orc.UINT4, 2.5, 2)):
        for total in (27_264_000, n_all):
            b, e = pd.shard_bounds(total, world, rank)
            n_r = e - b
            # whole-tensor parameters through the communicator (bf16: its own min/max)
            if DIN == D.F32:
                s_g, z_g = ctx.compute_quant_params_ptr_float32(src[b:e].data_ptr(), DQ, n_r)
            else:
                s_g, z_g = ctx.compute_quant_params_ptr_bfloat16(src[b:e].data_ptr(), DQ, n_r)
            # rotating windows over the rank's 1e9-element buffer, so that L2 (126 MB) never serves a small shard twice
            windows = max(1, min(64, n_all // max(n_r, 1)))
            offs = [(b + j * n_r) if (b + (j + 1) * n_r) <= n_all else b for j in range(windows)] if total < n_all else [b]
            offs = [o - o % 64 for o in offs]

            ptrs = [(src[o:o + n_r].data_ptr(), q[o // per:].data_ptr()) for o in offs]      # (slicing a tensor costs more host time than the call)

            def launch(k, n_r=n_r, ptrs=ptrs, DIN=DIN, DQ=DQ, s_g=s_g, z_g=z_g):
                pi, po = ptrs[k % len(ptrs)]
                ctx.quantize_ptr(pi, DIN, po, DQ, n_r, s_g, z_g, RoundMode.NEAREST)
            reps = 200 if total < n_all else 20
            t_n = device_time(launch, reps)
            t_graph = None
            if total < n_all:
                # a shard this small is issue-bound from Python (a call costs 4-5 us of host time, the kernel less): the same launches
                # captured once into a CUDA graph and replayed show what the GPU side takes
                try:
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
                        for k in range(64):
                            launch(k)
                    ctx.set_stream(stream.cuda_stream)
                    t_graph = device_time(lambda k: graph.replay(), 5, 2) / 64
                    del graph
                except Exception:      # noqa: BLE001
                    ctx.set_stream(stream.cuda_stream)
            o1 = [j * total for j in range(max(1, min(32, n_all // total)))]
            t_1 = only_rank0(lambda: device_time_local(torch, lambda k: solo.quantize_ptr(src[o1[k % len(o1)]:].data_ptr(), DIN, q[o1[k % len(o1)] // per:].data_ptr(), DQ, total, s_g, z_g, RoundMode.NEAREST), reps))
            # parity: the shard's packed bytes at both shard boundaries == the oracle on the same elements with the global parameters
            launch(0)
            torch.cuda.synchronize()
            o = offs[0]
            ok = True
            for w0, w1 in oracle_windows(n_r):
                w0 -= w0 % 64
                xin = src[o + w0:o + w1]
                host = xin.cpu().numpy() if DIN == D.F32 else xin.view(torch.int16).cpu().numpy().view("uint16")
                got = q[(o + w0) // per:(o + w0) // per + orc.packed_bytes(odq, w1 - w0)].cpu().numpy()
                ok = ok and bool((orc.quantize(host, odq, s_g, z_g) == got).all())
            gbs = bpe * total / t_n / 1e9
            out[f"quantize_{dt_name}_{'27.264M' if total < n_all else '1e9'}_sharded"] = {
                "numel_total": total, "numel_per_gpu": n_r, "us": round(t_n * 1e6, 2), "us_one_gpu_whole_tensor": round(t_1 * 1e6, 2),
                "Gelem/s": round(total / t_n / 1e9, 1), "GB/s_aggregate": round(gbs, 1), "frac_of_measured_peak_x_N": round(gbs / (peak * world), 4),
                "strong_scaling_efficiency": round(t_1 / (world * t_n), 4), "parity_sharded": flags_all(ok),
                "note": "CUDA events over back-to-back launches on rotating windows, max over ranks; parity = packed bytes of both shard-boundary windows vs the CPU oracle on every rank"}
            if t_graph is not None:
                out[f"quantize_{dt_name}_27.264M_sharded"].update({
                    "us_cuda_graph_replay": round(t_graph * 1e6, 2), "GB/s_aggregate_cuda_graph_replay": round(bpe * total / t_graph / 1e9, 1),
                    "strong_scaling_efficiency_cuda_graph_replay": round(t_1 / (world * t_graph), 4),
                    "note_cuda_graph": "64 of the same launches captured once and replayed: no host issue cost; T_1 is the eager one-GPU time"})
    del xb_all

    # ---- C5: u8[total] -> f32 dequantize with the ADD store op, sharded ------------------------------------------
    total = n_all
    b, e = pd.shard_bounds(total, world, rank)
    n_r = e - b
    acc = torch.zeros(n_r, dtype=torch.float32, device=dev)
    ctx.quantize_ptr(x[b:e].data_ptr(), D.F32, q[b:e].data_ptr(), D.UINT8, n_r, scale_g, zp_g, RoundMode.NEAREST)

    def add(_k=0):
        ctx.dequantize_ptr(q[b:e].data_ptr(), D.UINT8, acc.data_ptr(), D.F32, n_r, scale_g, zp_g, ReduceOp.ADD)
    t_n = device_time(add, 20)
    acc.zero_()
    add()
    add()
    torch.cuda.synchronize()
    ok = True
    for w0, w1 in oracle_windows(n_r):
        qw = q[b + w0:b + w1].cpu().numpy()
        wantw = orc.dequantize(qw, orc.UINT8, w1 - w0, orc.F32, scale_g, zp_g, orc.ADD,
                               out=orc.dequantize(qw, orc.UINT8, w1 - w0, orc.F32, scale_g, zp_g, orc.ADD))
        ok = ok and bool((acc[w0:w1].cpu().numpy().view("uint32") == wantw.view("uint32")).all())
    del acc

    def solo_c5():
        acc1 = torch.zeros(total, dtype=torch.float32, device=dev)
        t = device_time_local(torch, lambda k: solo.dequantize_ptr(q.data_ptr(), D.UINT8, acc1.data_ptr(), D.F32, total, scale_g, zp_g, ReduceOp.ADD), 10)
        del acc1
        return t
    t_1 = only_rank0(solo_c5)
    gbs = 9.0 * total / t_n / 1e9
    out["C5_u8_f32_dequantize_add_1e9_sharded"] = {
        "numel_total": total, "numel_per_gpu": n_r, "ms": round(t_n * 1e3, 4), "ms_one_gpu_whole_tensor": round(t_1 * 1e3, 4),
        "Gelem/s": round(total / t_n / 1e9, 1), "GB/s_aggregate": round(gbs, 1), "frac_of_measured_peak_x_N": round(gbs / (peak * world), 4),
        "strong_scaling_efficiency": round(t_1 / (world * t_n), 4), "parity_sharded": flags_all(ok),
        "note": "all ranks concurrently, 9 B/element; parity = two ADD passes into zeros vs the oracle, both shard-boundary windows, every rank"}
    if world > 1:
        ctx.comm_destroy()
    return out


def small_tensor_leg(torch, ctx, dev) -> dict:
    """The reference's own Python benchmark (reference python/benchmark/benchmark.py:16-23): NUMEL = 1e6 f32, 1000 runs of
    piquant.torch.quantize per quantized dtype (each run allocates its output, like the reference's), total seconds -- here on CUDA
    tensors, next to the SAME thousand tensors handed over as one batch (piquant.torch.quantize_batch: one launch per 256)."""
    import piquant.torch as pt

    runs, numel = 1000, 1_000_000
    res = {"numel": numel, "runs": runs, "recipe": "reference python/benchmark/benchmark.py: torch.rand(NUMEL), compute_quant_params once, quantize x 1000"}
    for tdt, name, bpe in ((torch.quint8, "quint8", 5.0), (torch.quint4x2, "quint4x2", 4.5), (torch.quint2x4, "quint2x4", 4.25)):
        try:
            t1 = torch.rand(numel, dtype=torch.float32, device=dev)
            s_, z_ = pt.compute_quant_params(t1, dtype=tdt, ctx=ctx)
            keep: list = []

            def per_call():
                keep.clear()
                for _ in range(runs):
                    keep.append(pt.quantize(t1, scale=s_, zero_point=z_, dtype=tdt, ctx=ctx))

            def batched():
                keep.clear()
                keep.extend(pt.quantize_batch([t1] * runs, scales=[s_] * runs, zero_points=[z_] * runs, dtype=tdt, ctx=ctx))

            row = {}
            for label, fn in (("per_call", per_call), ("one_batch", batched)):
                fn()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                t = time.perf_counter() - t0
                row[label] = {"total_s": round(t, 5), "us_per_tensor": round(t / runs * 1e6, 2), "Gelem/s": round(runs * numel / t / 1e9, 1),
                              "GB/s": round(bpe * runs * numel / t / 1e9, 1)}
            # the same batch with outputs allocated beforehand (the allocation of 1000 quantized torch tensors is most of the
            # time above; what remains is mostly Python turning 1000 tensors into descriptors): wall clock of the call and the
            # time between two CUDA events around it
            outs = [torch.empty(t1.shape, dtype=tdt, device=dev) for _ in range(runs)]
            args = dict(scales=[s_] * runs, zero_points=[z_] * runs, dtype=tdt, ctx=ctx, outs=outs)
            pt.quantize_batch([t1] * runs, **args)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            pt.quantize_batch([t1] * runs, **args)
            e1.record()
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            dev_t = e0.elapsed_time(e1) * 1e-3
            row["one_batch_preallocated_outputs"] = {"total_s": round(t, 5), "us_per_tensor": round(t / runs * 1e6, 2),
                                                     "device_us_per_tensor": round(dev_t / runs * 1e6, 3),
                                                     "device_GB/s": round(bpe * runs * numel / dev_t / 1e9, 1)}
            # a PREPARED batch (piquant.torch.QuantizeBatch: descriptor array built once): what is left is one native call
            prepared = pt.QuantizeBatch([t1] * runs, **args)
            prepared.run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            prepared.run()
            e1.record()
            torch.cuda.synchronize()
            t = time.perf_counter() - t0
            dev_t = e0.elapsed_time(e1) * 1e-3
            row["prepared_batch"] = {"total_s": round(t, 5), "us_per_tensor": round(t / runs * 1e6, 3), "device_us_per_tensor": round(dev_t / runs * 1e6, 3),
                                     "device_GB/s": round(bpe * runs * numel / dev_t / 1e9, 1)}
            del prepared
            if name == "quint8":
                # ... and with 1000 DISTINCT input tensors (the recipe above reads one 4 MB tensor a thousand times: every CTA column
                # of the batch hits the same L2 lines at once)
                many = torch.rand(runs, numel, dtype=torch.float32, device=dev)
                ins = [many[i] for i in range(runs)]
                prepared = pt.QuantizeBatch(ins, **args)
                prepared.run()
                torch.cuda.synchronize()
                e0.record()
                prepared.run()
                e1.record()
                torch.cuda.synchronize()
                dev_t = e0.elapsed_time(e1) * 1e-3
                row["prepared_batch_distinct_inputs"] = {"device_us_per_tensor": round(dev_t / runs * 1e6, 3),
                                                         "device_GB/s": round(bpe * runs * numel / dev_t / 1e9, 1),
                                                         "note": "5 GB of HBM traffic in one batch: the kernel's own rate"}
                del many, ins, prepared
            del outs
            a, b = keep[0], pt.quantize(t1, scale=s_, zero_point=z_, dtype=tdt, ctx=ctx)
            row["batch_equals_per_call"] = bool(torch.equal(torch.empty(0, dtype=torch.uint8, device=dev).set_(a.untyped_storage()),
                                                            torch.empty(0, dtype=torch.uint8, device=dev).set_(b.untyped_storage())))
            res[name] = row
            keep.clear()
        except Exception as e:      # noqa: BLE001
            res[name] = {"error": repr(e)[:200]}
    res["note"] = "the same 4 MB tensor every run, as in the reference's script: L2-resident on the GPU (and cache-resident on the CPU)"
    return res


def wall_time_local(torch, fn, reps, warm=3) -> float:
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def device_time_local(torch, fn, reps, warm=3) -> float:
    for k in range(warm):
        fn(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        fn(warm + k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def ring_allreduce_leg(torch, dist, ctx, world, rank, dev) -> dict:
    """The caller pattern the ADD store op exists for (reference README.md:29): SUM all-reduce of an f32 tensor with
    uint8 transport (piquant.distributed.quantized_all_reduce_) next to NCCL's f32 all-reduce of the same tensor."""
    from piquant import distributed as pd

    n = 1 << 28
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    base = torch.empty(n, dtype=torch.float32, device=dev).uniform_(-1, 1, generator=gen)
    work = torch.empty_like(base)

    def timed(fn, reps=5):
        for _ in range(2):
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

    def identical_on_every_rank() -> bool:
        """two wrapping checksums of the result's bit pattern, compared across ranks"""
        v = work.view(torch.int32).to(torch.int64)
        c = torch.stack([v.sum(), (v * (torch.arange(n, device=dev, dtype=torch.int64) % 8191 + 1)).sum()])
        lo, hi = c.clone(), c.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return bool(torch.equal(lo, hi))

    def error_stats(exact):
        d = (work - exact).double()
        st = torch.stack([d.abs().max(), d.mean().abs()])
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
        return round(float(st[0].item()), 5), float(f"{st[1].item():.3e}")

    ms_nccl = timed(lambda: dist.all_reduce(work))
    exact = work.clone()
    res = {"numel": n, "ms_nccl_f32": round(ms_nccl, 3)}
    bus = 2 * (world - 1) / world * n * 4 / 1e9
    res["nccl_busbw_GBps"] = round(bus / (ms_nccl * 1e-3), 1)
    for key, kw in (("quantized_u8_ring_nccl_sendrecv", dict(transport="nccl", algorithm="ring")),
                    ("quantized_u8_ring_p2p_fused_2_lanes", dict(transport="p2p", algorithm="ring", lanes=2)),
                    ("quantized_u8_ring_p2p_fused_2_lanes_stochastic_per_element",
                     dict(transport="p2p", algorithm="ring", lanes=2, round_mode="stochastic_per_element")),
                    ("quantized_u8_direct_over_nccl_collectives", dict(transport="nccl", algorithm="direct")),
                    ("quantized_u8_direct_all_to_all", dict(transport="p2p", algorithm="direct")),
                    ("quantized_u8_direct_all_to_all_stochastic_per_element", dict(transport="p2p", algorithm="direct", round_mode="stochastic_per_element")),
                    ("quantized_u8_direct_all_to_all_cuda_graph", dict(graph=True)),
                    ("quantized_u4_direct_all_to_all_cuda_graph", dict(graph=True, qdtype="quint4x2"))):
        try:
            kw = dict(kw)
            qdtype = getattr(torch, kw.pop("qdtype", "quint8"))
            if kw.pop("graph", False):
                plan = pd.QuantizedAllReduce(work, dtype=qdtype, ctx=ctx)
                ms = timed(plan)
                del plan
            else:
                ms = timed(lambda: pd.quantized_all_reduce_(work, dtype=qdtype, ctx=ctx, **kw))
            mx, mean = error_stats(exact)
            res[key] = {"ms": round(ms, 3), "speedup_vs_nccl_f32": round(ms_nccl / ms, 3), "effective_busbw_GBps": round(bus / (ms * 1e-3), 1),
                        "max_abs_err": mx, "abs_mean_err": mean, "bit_identical_on_every_rank": identical_on_every_rank()}
        except Exception as e:      # noqa: BLE001  (symmetric memory not available on this box)
            res[key] = f"unavailable: {type(e).__name__}: {str(e)[:120]}"
    # parity in the same run: a 1 M-element all-reduce of every form replayed on the CPU with the oracle (oracle/replay.py), bit for bit
    try:
        import numpy as np
        from oracle import port as orc, replay
        m = 1_000_003
        small = torch.empty(m, dtype=torch.float32, device=dev).uniform_(-1, 1, generator=gen)
        gathered = [torch.empty_like(small) for _ in range(world)]
        dist.all_gather(gathered, small)
        host = [g.cpu().numpy() for g in gathered]
        want = {"direct": replay.direct_all_reduce(host, orc.UINT8, orc.F32, pd.shard_bounds, pd.SHARD_ALIGN, 1),
                "ring": replay.ring_all_reduce(host, orc.UINT8, orc.F32, pd.shard_bounds, pd.SHARD_ALIGN, 1)}
        checks = {}
        for name, run in (("direct", lambda t: pd.quantized_all_reduce_(t, dtype=torch.quint8, ctx=ctx, transport="p2p", algorithm="direct", lanes=1)),
                          ("direct_over_nccl", lambda t: pd.quantized_all_reduce_(t, dtype=torch.quint8, ctx=ctx, transport="nccl", algorithm="direct")),
                          ("direct_cuda_graph", lambda t: pd.QuantizedAllReduce(t, dtype=torch.quint8, ctx=ctx, lanes=1)()),
                          ("ring", lambda t: pd.quantized_all_reduce_(t, dtype=torch.quint8, ctx=ctx, transport="p2p", algorithm="ring", lanes=1))):
            t = small.clone()
            run(t)
            ok = torch.tensor([int(np.array_equal(t.cpu().numpy().view(np.uint8), want[name.split("_")[0]]))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            checks[name] = bool(ok.item())
        res["parity_vs_oracle_replay"] = {"numel": m, "bit_exact_on_every_rank": checks,
                                          "what": "GPU result vs the collective replayed with the CPU oracle's compute_quant_params/quantize/dequantize"}
    except Exception as e:      # noqa: BLE001
        res["parity_vs_oracle_replay"] = f"unavailable: {type(e).__name__}: {str(e)[:160]}"
    res["note"] = ("ring reduce-scatter + all-gather, [64 B params | u8 payload] per hop, no host sync; a reduce-scatter hop = quantize + ONE fused "
                   "dequantize-ADD/min-max/params kernel; p2p_fused = kernels store into / forward to the neighbour's slot over NVLink peer memory, "
                   "nothing is sent; abs_mean_err shows the bias nearest rounding accumulates per hop and per-element stochastic rounding does not; "
                   "direct_over_nccl_collectives = the direct algorithm with one all_to_all_single + one all_gather_into_tensor (no peer memory needed); "
                   "direct_all_to_all = the NVSwitch form: quantize each chunk once, copy engines move it to its owner, ONE multi-source "
                   "dequantize-sum kernel reduces, the packed sums are broadcast by copy engines and dequantized: 2 quantizations per value "
                   "whatever the world size, 1 barrier + per-slot arrival flags written by the copy engine; cuda_graph = the same collective of a "
                   "persistent tensor captured once (piquant.distributed.QuantizedAllReduce) and replayed")
    return res


def cpu_reference_leg(x_host, q_gpu_host, scale, zp, ctx, D, RoundMode, ReduceOp):
    """The unmodified reference on this box's host cores (oracle/_ref), bounded to ~10 s, and a full-size
    bit comparison of its output with the GPU library's -- for the headline and for every other BASELINE config
    (the GPU side of those comparisons runs through the same C ABI on the same host arrays)."""
    import numpy as np

    try:
        from oracle import ref
        if not ref.available():
            raise RuntimeError("oracle/_ref not built")
    except Exception as e:   # fall back to the scalar C port on a small sample
        from oracle import port
        m = min(x_host.size, 20_000_000)
        t0 = time.perf_counter()
        out = port.quantize(x_host[:m], port.UINT8, scale, zp)
        t = time.perf_counter() - t0
        mism = int((out != q_gpu_host[:m]).sum())
        return ({"value": round(m / t / 1e9, 4), "unit": UNIT, "cores": 1, "kind": "port", "sample": f"first {m} elements, 1 pass ({e})"},
                {"numel": m, "mismatches": mism})
    cores = os.cpu_count() or 1
    c = ref.Context(cores)
    n = x_host.size
    out = np.empty(n, dtype=np.uint8)
    c.quantize(x_host, ref_dt("UINT8"), scale, zp, 0, out=out)       # warm-up pass (also the parity pass)
    passes, t0 = 0, time.perf_counter()
    while True:
        c.quantize(x_host, ref_dt("UINT8"), scale, zp, 0, out=out)
        passes += 1
        if time.perf_counter() - t0 > 10.0 or passes >= 400:        # ~10 s of CPU work
            break
    t = (time.perf_counter() - t0) / passes
    mism = int(np.count_nonzero(out != q_gpu_host))

    # the other BASELINE configs on the same cores, a few passes each (a reported baseline beside the GPU lines in `extra`)
    def timed(fn, max_s=2.0, max_passes=20):
        fn()
        k, t1 = 0, time.perf_counter()
        while True:
            fn()
            k += 1
            if time.perf_counter() - t1 > max_s or k >= max_passes:
                break
        return (time.perf_counter() - t1) / k, k

    others = {}
    try:
        from oracle.port import f32_to_bf16_bits
        n2 = min(100_000_000, n)
        xb = f32_to_bf16_bits(x_host[:n2])
        s4, z4 = c.compute_quant_params(xb, ref_dt("UINT4"))
        q4 = np.empty((n2 + 1) // 2, dtype=np.uint8)
        yb = np.empty(n2, dtype=np.uint16)
        acc = np.zeros(n // 8, dtype=np.float32)
        legs = {
            "C2_bf16_u4_quantize_1e8": (n2, lambda: c.quantize(xb, ref_dt("UINT4"), s4, z4, 0, out=q4)),
            "C2_u4_bf16_dequantize_1e8": (n2, lambda: c.dequantize(q4, ref_dt("UINT4"), n2, ref_dt("BF16"), s4, z4, 0, out=yb)),
            "C3_compute_quant_params_f32": (n, lambda: c.compute_quant_params(x_host, ref_dt("UINT8"))),
            "C4_f32_u8_stochastic": (n, lambda: c.quantize(x_host, ref_dt("UINT8"), scale, zp, 1, out=out)),
            "C5_u8_f32_dequantize_add_shard": (n // 8, lambda: c.dequantize(out[: n // 8], ref_dt("UINT8"), n // 8, ref_dt("F32"), scale, zp, 1, out=acc)),
        }
        for name, (m, fn) in legs.items():
            tt, k = timed(fn)
            others[name] = {"numel": m, "ms": round(tt * 1e3, 3), "Gelem/s": round(m / tt / 1e9, 3), "passes": k}
    except Exception as e:      # noqa: BLE001
        others["error"] = repr(e)[:200]
    try:        # the reference's own small-tensor recipe on these cores: 1e6 elements x 1000 runs per dtype (benchmark.py:16-23)
        small = {}
        x1 = np.ascontiguousarray(np.abs(x_host[:1_000_000]))
        for name, dtn in (("quint8", "UINT8"), ("quint4x2", "UINT4"), ("quint2x4", "UINT2")):
            s1, z1 = c.compute_quant_params(x1, ref_dt(dtn))
            o1 = np.empty(orc_packed(ref_dt(dtn), x1.size), dtype=np.uint8)
            c.quantize(x1, ref_dt(dtn), s1, z1, 0, out=o1)
            t1 = time.perf_counter()
            for _ in range(1000):
                c.quantize(x1, ref_dt(dtn), s1, z1, 0, out=o1)
            tt = time.perf_counter() - t1
            small[name] = {"total_s": round(tt, 5), "us_per_tensor": round(tt * 1e3, 2), "Gelem/s": round(1e9 / tt / 1e9, 2)}
        others["small_tensor_regime_1e6_x_1000"] = small
    except Exception as e:      # noqa: BLE001
        others["small_tensor_regime_1e6_x_1000"] = {"error": repr(e)[:200]}
    per_config = {"headline_f32_u8_nearest_1e9": {"numel": n, "mismatches": mism}}
    try:
        from oracle import port as orc
        bf = orc.bf16_bits_to_f32
        # C2: bf16 -> u4 bytes, u4 -> bf16 bits, and the round-trip error of BOTH implementations on the same data
        q4g = np.empty_like(q4)
        ctx.quantize_ptr(xb.ctypes.data, D.BF16, q4g.ctypes.data, D.UINT4, n2, s4, z4, RoundMode.NEAREST)
        ybg = np.empty_like(yb)
        ctx.dequantize_ptr(q4.ctypes.data, D.UINT4, ybg.ctypes.data, D.BF16, n2, s4, z4, ReduceOp.SET)
        per_config["C2_bf16_u4_quantize_1e8"] = {"numel": n2, "mismatches": int(np.count_nonzero(q4g != q4))}
        per_config["C2_u4_bf16_dequantize_1e8"] = {"numel": n2, "mismatches": int(np.count_nonzero(ybg != yb))}
        xf = bf(xb)
        per_config["C2_round_trip_max_abs_err_over_scale"] = {
            "gpu": round(float(np.abs(bf(ybg) - xf).max()) / s4, 4), "cpu_reference": round(float(np.abs(bf(yb) - xf).max()) / s4, 4),
            "note": "above 0.5 on both: the excess is the bf16 rounding of the dequantized output, not the quantizer"}
        del xf, q4g, ybg
        # C3: parameters of the whole tensor
        per_config["C3_compute_quant_params_f32"] = {
            "gpu": list(ctx.compute_quant_params_ptr_float32(x_host.ctypes.data, D.UINT8, n)), "cpu_reference": list(c.compute_quant_params(x_host, ref_dt("UINT8")))}
        per_config["C3_compute_quant_params_f32"]["equal"] = per_config["C3_compute_quant_params_f32"]["gpu"] == per_config["C3_compute_quant_params_f32"]["cpu_reference"]
        # C4: the reference's stochastic output (its threshold cannot be set: one value per call from an unreachable RNG,
        # src/piquant.cpp:194-201) brackets that threshold; the GPU is run with a threshold from the bracket and must match bytewise
        m4 = min(n, 200_000_000)
        ref4 = c.quantize(x_host[:m4], ref_dt("UINT8"), scale, zp, 1)
        xi = orc.infer_stochastic_threshold(x_host[:m4], scale, zp, 255, ref4)
        if xi is None:
            per_config["C4_f32_u8_stochastic"] = {"numel": m4, "error": "no single threshold explains the reference's output"}
        else:
            g4 = np.empty_like(ref4)
            ctx.set_stochastic_threshold(xi)
            ctx.quantize_ptr(x_host[:m4].ctypes.data, D.F32, g4.ctypes.data, D.UINT8, m4, scale, zp, RoundMode.STOCHASTIC)
            ctx.set_stochastic_threshold(None)
            per_config["C4_f32_u8_stochastic"] = {"numel": m4, "threshold_inferred_from_reference_output": xi, "mismatches": int(np.count_nonzero(g4 != ref4))}
            del g4
        del ref4
        # C5: one ADD pass into a non-zero accumulator
        n5 = n // 8
        prev = x_host[n5:2 * n5].copy()
        acc_r, acc_g = prev.copy(), prev.copy()
        c.dequantize(out[:n5], ref_dt("UINT8"), n5, ref_dt("F32"), scale, zp, 1, out=acc_r)
        ctx.dequantize_ptr(out[:n5].ctypes.data, D.UINT8, acc_g.ctypes.data, D.F32, n5, scale, zp, ReduceOp.ADD)
        per_config["C5_u8_f32_dequantize_add_shard"] = {"numel": n5, "mismatches": int(np.count_nonzero(acc_g.view(np.uint32) != acc_r.view(np.uint32)))}
    except Exception as e:      # noqa: BLE001
        per_config["error"] = repr(e)[:300]
    c.close()
    return ({"value": round(n / t / 1e9, 3), "unit": UNIT, "cores": cores, "kind": "reference", "isa": ref.cpu_isa(),
             "sample": f"numel={n} (the full workload), {passes} passes, {cores} threads, mean",
             "other_configs_same_cores": others},
            {"numel": n, "mismatches": mism, "what": "GPU output vs reference CPU output on the same host arrays, byte for byte (bf16 / f32 results: bit patterns)",
             "per_config": per_config})


def ref_dt(name: str) -> int:
    from oracle import port
    return getattr(port, name)


def orc_packed(dt: int, numel: int) -> int:
    from oracle import port
    return port.packed_bytes(dt, numel)


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------

def run_reference(args) -> None:
    if int(os.environ.get("RANK", 0)) != 0:
        return
    import numpy as np

    from oracle import port
    kind, cores = "reference", os.cpu_count() or 1
    try:
        from oracle import ref
        if not ref.available():
            raise RuntimeError
        c = ref.Context(cores)
        isa = ref.cpu_isa()

        def quant(xs, out):
            c.quantize(xs, port.UINT8, scale, zp, 0, out=out)
    except Exception:
        kind, cores, isa = "port", 1, "scalar C"

        def quant(xs, out):
            port.quantize(xs, port.UINT8, scale, zp, out=out)
    n = NUMEL if kind == "reference" else min(NUMEL, 50_000_000)
    rng = np.random.default_rng(0)
    x = np.empty(n, dtype=np.float32)
    blk = 1 << 24
    for i in range(0, n, blk):
        x[i:i + blk] = rng.random(min(blk, n - i), dtype=np.float32) * 2 - 1
    scale, zp = port.compute_quant_params(x[: 1 << 24], port.UINT8) if kind == "port" else c.compute_quant_params(x, port.UINT8)
    out = np.empty(n, dtype=np.uint8)
    t0 = time.perf_counter()
    quant(x, out)
    t1 = time.perf_counter() - t0
    steps, warm = args.steps, max(args.warmup, 1)
    m = n
    if t1 * (steps + warm) > 240.0:          # bound the whole run to a few minutes
        m = max(1 << 20, int(n * 240.0 / (t1 * (steps + warm))))
    xs, outs = x[:m], out[:m]
    for _ in range(warm):
        quant(xs, outs)
    t0 = time.perf_counter()
    for _ in range(steps):
        quant(xs, outs)
    t = (time.perf_counter() - t0) / steps
    value = m / t / 1e9
    sample = f"numel={m} per step ({'the full workload' if m == NUMEL else 'bounded sample of numel=' + str(NUMEL)}), {cores} threads, isa={isa}"
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": round(t * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32->u8 (f32 multiply/add, int32 clamp)", "data": "synthetic U(-1,1)", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def claim_stdout() -> None:
    """Only the JSON line may reach stdout: libraries (NCCL prints its version banner there) are sent to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
