"""piquant.distributed -- a tensor sharded contiguously over the GPUs of one box.

The reference has no distributed code at all; this module is the B200 side of
SURVEY.md section 8(e).  Shards are independent for quantize / dequantize / requantize -- every rank
simply calls ``piquant.torch`` on its shard with the shared ``(scale, zero_point)``, no collective.
The only exchange step is the whole-tensor min/max behind ``compute_quant_params``: each rank
reduces its shard on its GPU and the ranks combine ``{-min, max}`` with ONE 2-float MAX all-reduce
(NCCL over NVLink on GPUs; any ``torch.distributed`` backend works, which is how the host logic is
tested with ``gloo`` on CPUs).  min/max of non-NaN floats is exact and associative, so every rank
gets bit-identical parameters, equal to the single-GPU / CPU-reference result.

Two equivalent routes are provided:
* ``compute_quant_params_sharded``   -- min/max kernel -> ``torch.distributed.all_reduce(MAX)``;
* ``init_native_comm``               -- hands the native library its own NCCL communicator, after which
  plain ``piquant.torch.compute_quant_params`` / ``piquant_compute_quant_params_*`` do the same
  all-reduce inside the C library (kernel and ncclAllReduce on one stream, one host sync).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import Context, DataType, ReduceOp, RoundMode
from .torch import _QUANT_TYPES, _site, torch_to_piquant_dtype

SHARD_ALIGN = 64    # elements: a multiple of every pack width (4 for uint2); 128 B of bf16, 16 B of packed uint2


def shard_bounds(numel: int, world_size: int, rank: int, align: int = SHARD_ALIGN) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous shard.  Boundaries are multiples of ``align`` elements so that
    packed bytes never straddle two shards and both streams stay vector-aligned; the last rank takes the
    remainder (the reference splits ranges over threads the same way, reference src/piquant.cpp:145-157)."""
    assert 0 <= rank < world_size and numel >= 0
    per = (numel // world_size) // align * align
    begin = per * rank
    end = numel if rank == world_size - 1 else per * (rank + 1)
    return begin, end


def combine_minmax(neg_min_max: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> Tuple[float, float]:
    """All-reduce(MAX) a 2-float tensor ``{-min, max}`` over ``group``; returns the global (min, max)."""
    assert neg_min_max.numel() == 2 and neg_min_max.dtype == torch.float32
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(neg_min_max, op=dist.ReduceOp.MAX, group=group)
    mn, mx = (-neg_min_max[0]).item(), neg_min_max[1].item()
    return mn, mx


def local_neg_min_max(shard: torch.Tensor, ctx: Context = Context.get()) -> torch.Tensor:
    """``{-min, max}`` of a CUDA shard as a 2-float CUDA tensor (asynchronous: one kernel on the current
    stream).  An empty shard contributes the identity ``{-FLT_MAX, -FLT_MAX}``."""
    assert shard.is_cuda, "the min/max kernel runs on the GPU; there is no CPU path"
    out4 = torch.empty(4, dtype=torch.float32, device=shard.device)
    if shard.numel() == 0:
        fmax = torch.finfo(torch.float32).max
        return torch.full((2,), -fmax, dtype=torch.float32, device=shard.device)
    shard = shard if shard.is_contiguous() else shard.contiguous()
    device, stream = _site(shard)
    ctx.minmax_on_stream(shard.data_ptr(), torch_to_piquant_dtype(shard.dtype), shard.numel(), out4.data_ptr(), Context.FLAG_LOCAL,
                         device, stream)
    return out4[2:4]


def compute_quant_params_sharded(shard: torch.Tensor, *, dtype: torch.dtype, group: Optional[dist.ProcessGroup] = None,
                                 ctx: Context = Context.get()) -> Tuple[float, int]:
    """Whole-tensor ``(scale, zero_point)`` from this rank's shard; identical on every rank."""
    assert dtype in _QUANT_TYPES, f"Unsupported quantized dtype: {dtype}. Must be one of {list(_QUANT_TYPES)}"
    mn, mx = combine_minmax(local_neg_min_max(shard, ctx), group)
    return params_from_minmax(mn, mx, dtype)


def params_from_minmax(mn: float, mx: float, dtype: torch.dtype) -> Tuple[float, int]:
    """The reference's double-precision scale / zero-point arithmetic (reference src/piquant.cpp:245-258),
    executed by the native library on the host."""
    return Context.params_from_minmax(mn, mx, torch_to_piquant_dtype(dtype))


def init_native_comm(ctx: Context = Context.get(), group: Optional[dist.ProcessGroup] = None) -> None:
    """Give ``ctx`` its own NCCL communicator spanning ``group`` (bootstrapped through torch.distributed).
    Afterwards ``compute_quant_params`` on this context returns whole-tensor parameters."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = [Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init_rank(uid[0], world, rank)


def destroy_native_comm(ctx: Context = Context.get()) -> None:
    ctx.comm_destroy()


__all__ = ["SHARD_ALIGN", "shard_bounds", "combine_minmax", "local_neg_min_max", "compute_quant_params_sharded",
           "params_from_minmax", "init_native_comm", "destroy_native_comm", "ring_schedule", "gather_pieces", "quantized_all_reduce_", "QuantizedAllReduce",
           "DataType"]


# ----------------------------------------------------------------------------------------------------
# quantized ring all-reduce: the caller pattern the reference's ADD store op exists for
# (reference README.md:29 "useful for ring-reduction operations")
# ----------------------------------------------------------------------------------------------------

def ring_schedule(world_size: int, rank: int):
    """Chunk indices of a ring all-reduce.  Returns (reduce_scatter, all_gather): two lists of
    ``(send_chunk, recv_chunk)`` per step; data always flows rank -> rank+1.  After reduce-scatter rank r
    holds the complete sum of chunk (r + 1) % world_size."""
    w, r = world_size, rank
    reduce_scatter = [((r - s) % w, (r - s - 1) % w) for s in range(w - 1)]
    all_gather = [((r + 1 - s) % w, (r - s) % w) for s in range(w - 1)]
    return reduce_scatter, all_gather


_P2P_SLOTS: dict = {}
_SIDE_STREAMS: dict = {}


def _p2p_slots(nbytes: int, device: torch.device, group, lane: int = 0):
    """Two receive slots per rank in symmetric memory (every rank can address every other rank's slots over
    NVLink).  Cached per (group, device, size, lane): the rendezvous is a collective and costs milliseconds."""
    import torch.distributed._symmetric_memory as symm_mem

    grp = group if group is not None else dist.group.WORLD
    key = (grp.group_name, device.index, nbytes, lane)
    if key not in _P2P_SLOTS:
        buf = symm_mem.empty(2 * nbytes, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, grp)
        _P2P_SLOTS[key] = (buf, hdl)
    return _P2P_SLOTS[key]


def _peer_memory_available(device: torch.device, group) -> bool:
    """is symmetric memory available on this box / build?  (a collective the first time: every rank must ask)"""
    try:
        _p2p_slots(256, device, group, -1)
        return True
    except Exception:      # noqa: BLE001
        return False


def _side_stream(device: torch.device, lane: int) -> "torch.cuda.Stream":
    key = (device.index, lane)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def quantized_all_reduce_(tensor: torch.Tensor, *, dtype: torch.dtype = torch.quint8, group: Optional[dist.ProcessGroup] = None,
                          ctx: Context = Context.get(), transport: str = "auto", round_mode: str = "nearest", lanes: Optional[int] = None,
                          algorithm: str = "auto", multicast: Optional[bool] = None) -> torch.Tensor:
    """In-place SUM all-reduce of a contiguous CUDA float32 / bfloat16 tensor with quantized transport.

    ``algorithm="direct"`` (needs peer memory; see ``_DirectPlan``) is the NVSwitch-native form: every chunk crosses the
    links once per phase like in a ring, but in ONE all-to-all exchange per phase instead of world-1 dependent hops,
    every value is quantized twice in total instead of ``world`` times, and the reduce step is one pass.  Measured on
    8 B200s, 2^28 f32, uint8 transport: 1.09 ms against 1.40 ms for the ring below and 2.6 ms for NCCL's f32 all-reduce
    (profiles/r2_allreduce_probe_n8.txt); captured into a CUDA graph (``QuantizedAllReduce``) 1.06 ms.
    ``algorithm="auto"`` (default) = direct.  Its ``transport``: ``"p2p"`` (copy engines into peer memory; ``"auto"`` picks it
    when the GPUs can map each other's memory) or ``"nccl"`` -- the same algorithm, same results bit for bit, with the two
    exchanges done by ONE ``all_to_all_single`` and ONE ``all_gather_into_tensor`` of the packed slots: the form for boxes
    without peer access (the 2 * (world - 1) send/recv hops of the ring over NCCL are slower than NCCL's own f32 all-reduce
    beyond 2 GPUs: 0.39x at 8).
    ``multicast=True`` (direct form, opt-in): broadcast the reduced chunks through the NVSwitch multicast address of the symmetric
    buffer; bit-identical results, measured slower than the default at 8 GPUs (see ``_DirectPlan``).
    ``lanes=None``: 2 for the direct form on 2 GPUs and tensors of 256 MB or more, else 1.  The rest of this text describes
    ``algorithm="ring"``.

    Ring reduce-scatter + ring all-gather over NVLink; every hop carries ``[64-byte parameter block | packed
    payload]`` -- 1, 1/2 or 1/4 byte per element instead of 4 (or 2).  A reduce-scatter hop is TWO passes over the
    chunk: the sender quantizes it with parameters that are already on the device, the receiver runs ONE fused kernel --
    dequantize with the ADD store op into its accumulator chunk AND min/max of the sums AND the reference's
    scale / zero-point arithmetic for the next hop (``piquant_cuda_dequantize_add_minmax_on_stream``): the chunk a rank
    accumulates at hop s is exactly the chunk it sends at hop s + 1, so no separate min/max pass ever reads it.
    Parameters never visit the host; the whole collective is enqueued without a single synchronisation.  In the
    all-gather phase the owner of a reduced chunk dequantizes its own packed bytes too, so every rank ends with
    bit-identical values.  The result is the sum up to quantization error (<= 0.5 * scale per hop and element).
    ``round_mode="stochastic_per_element"`` rounds every element with its own Philox random number instead (extension,
    ``piquant_cuda.h``): the error per hop is then <= 1 * scale but has zero mean, so it averages out over steps and ranks
    instead of accumulating as a bias -- what gradient compression wants; each rank draws its own key per hop.

    ``transport="nccl"``: each hop is an NCCL send/recv of the packed buffer.  ``transport="p2p"``: nothing is ever
    sent -- the sender's quantize kernel stores its packed output DIRECTLY into the receiver's slot through NVLink peer
    memory (torch symmetric memory), the fused receiver kernel drops the next hop's parameter block into the next
    receiver's slot, and in the all-gather phase the kernel that dequantizes a received chunk also stores its packed
    bytes on to the next rank (``piquant_cuda_dequantize_forward_on_stream``): compute and transfer are one kernel per
    direction, with one stream-ordered barrier per hop publishing the slot.  ``transport="auto"``: p2p when symmetric
    memory is available, else nccl.

    ``lanes`` (p2p only): the tensor is cut into that many contiguous parts, each reduced by its own ring on its own
    stream, the hops enqueued alternately.  A hop alternates between an NVLink-bound kernel (the quantize that stores into
    peer memory) and an HBM-bound one (the fused accumulate); with two lanes one lane's link phase overlaps the other's
    memory phase.  Every lane is an independent all-reduce of its part: the result is a sum with per-part, per-chunk scales."""
    assert tensor.is_cuda and tensor.is_contiguous() and tensor.dtype in (torch.float32, torch.bfloat16)
    assert dtype in _QUANT_TYPES
    rmode = {"nearest": RoundMode.NEAREST, "stochastic_per_element": RoundMode.STOCHASTIC_PER_ELEMENT}[round_mode]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return tensor
    if transport not in ("nccl", "p2p", "auto"):
        raise ValueError(f"unknown transport {transport!r}")
    if algorithm not in ("ring", "direct", "auto"):
        raise ValueError(f"unknown algorithm {algorithm!r}")
    if algorithm == "auto":
        algorithm = "direct"
    if algorithm == "direct":
        if transport == "auto":
            transport = "p2p" if _peer_memory_available(tensor.device, group) else "nccl"
        if transport == "nccl":
            return _direct_all_reduce_nccl(tensor, dtype, group, ctx, rmode)
        return _direct_all_reduce(tensor, dtype, group, ctx, rmode, lanes if lanes is not None else _auto_lanes(tensor, world), multicast)
    lanes = 1 if lanes is None else lanes
    fdt, qdt = torch_to_piquant_dtype(tensor.dtype), torch_to_piquant_dtype(dtype)
    meta = Context.META_BYTES
    device = tensor.device.index
    LOCAL, REVERSE = Context.FLAG_LOCAL, Context.FLAG_REVERSE
    reduce_scatter, all_gather = ring_schedule(world, rank)
    own = (rank + 1) % world

    def ring(flat, stream, lane, use_p2p):
        """generator: enqueues one ring all-reduce of `flat` on `stream`, yielding after every hop"""
        bounds = [shard_bounds(flat.numel(), world, i) for i in range(world)]
        qbytes = [qdt.storage_bytes(e - b) for b, e in bounds]
        slot_bytes = (meta + max(qbytes) + 255) // 256 * 256

        def chunk(i):
            b, e = bounds[i]
            return flat[b:e]

        def first_params(i, meta_ptr):      # min/max + parameter arithmetic of chunk i in ONE launch; never the communicator's whole-tensor path
            c = chunk(i)
            if c.numel():
                ctx.compute_meta_on_stream(c.data_ptr(), fdt, c.numel(), qdt, meta_ptr, LOCAL, device, stream)

        def quantize_to(i, payload_ptr, meta_ptr, flags=0):
            c = chunk(i)
            if c.numel():
                ctx.quantize_meta_on_stream(c.data_ptr(), fdt, payload_ptr, qdt, c.numel(), rmode, meta_ptr, flags, device, stream)

        def accumulate(i, base_ptr, next_meta_ptr, next_meta_copy_ptr):      # chunk i += [meta | packed] at base_ptr; parameters of the sums out
            c = chunk(i)
            if c.numel():
                ctx.dequantize_add_minmax_on_stream(base_ptr + meta, qdt, c.data_ptr(), fdt, c.numel(), base_ptr, qdt, next_meta_ptr,
                                                    next_meta_copy_ptr, device, stream)

        def scatter(i, base_ptr):           # chunk i = dequantize([meta | packed] at base_ptr)
            c = chunk(i)
            if c.numel():
                ctx.dequantize_meta_on_stream(base_ptr + meta, qdt, c.data_ptr(), fdt, c.numel(), ReduceOp.SET, base_ptr, device, stream)

        def scatter_forward(i, base_ptr, fwd_base_ptr):      # ... and store the same [meta | packed] on to fwd_base_ptr (peer memory)
            c = chunk(i)
            if c.numel():
                ctx.dequantize_forward_on_stream(base_ptr + meta, qdt, c.data_ptr(), fdt, c.numel(), base_ptr, fwd_base_ptr + meta, fwd_base_ptr,
                                                 device, stream)

        keep = torch.empty(slot_bytes, dtype=torch.uint8, device=flat.device)      # [meta | packed] of the chunk this rank owns
        if use_p2p:
            local_slots, hdl = _p2p_slots(slot_bytes, flat.device, group, lane)
            my_base, nxt_base = local_slots.data_ptr(), int(hdl.buffer_ptrs[(rank + 1) % world])
            peer_hdr0 = hdl.get_buffer((rank + 1) % world, (meta,), torch.uint8)       # header of the neighbour's slot 0
            cur = torch.empty(meta, dtype=torch.uint8, device=flat.device)             # parameters of the chunk to send next
            hdl.barrier(channel=0)               # nobody is still reading the slots of a previous call
            first_params(reduce_scatter[0][0], cur.data_ptr())
            peer_hdr0.copy_(cur, non_blocking=True)
            step = 0
            for send_i, recv_i in reduce_scatter:
                off, nxt_off = (step % 2) * slot_bytes, ((step + 1) % 2) * slot_bytes
                # quantize straight into the neighbour's slot over NVLink (its header is there already); from hop 1 on the chunk
                # was just written by `accumulate`, so it is read from its end -- the part L2 still holds
                quantize_to(send_i, nxt_base + off + meta, cur.data_ptr(), REVERSE if step else 0)
                hdl.barrier(channel=0)                           # my slot `off` now holds chunk recv_i from my predecessor
                yield
                # ONE kernel: chunk += dequantize(slot); parameters of the sums -> `cur` and -> the header of the neighbour's NEXT slot
                accumulate(recv_i, my_base + off, cur.data_ptr(), nxt_base + nxt_off)
                step += 1
                yield
            # the owner keeps exactly what everybody else will receive: quantize the reduced chunk once, dequantize that
            keep[:meta].copy_(cur, non_blocking=True)
            quantize_to(own, keep.data_ptr() + meta, cur.data_ptr(), REVERSE)
            src = keep.data_ptr()
            for send_i, recv_i in all_gather:
                off = (step % 2) * slot_bytes
                scatter_forward(send_i, src, nxt_base + off)     # dequantize what I hold of chunk send_i AND store its packed bytes into the neighbour's slot
                hdl.barrier(channel=0)
                src = my_base + off                              # chunk recv_i has arrived; it is dequantized (and forwarded) by the next iteration
                step += 1
                yield
            scatter(all_gather[-1][1], src)
            return

        nxt = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
        prv = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
        send_buf, recv_buf, spare = keep, torch.empty_like(keep), torch.empty_like(keep)

        def exchange(send_buf, send_i, recv_buf, recv_i):
            ops = [dist.P2POp(dist.isend, send_buf[: meta + qbytes[send_i]], nxt, group=group),
                   dist.P2POp(dist.irecv, recv_buf[: meta + qbytes[recv_i]], prv, group=group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()                                       # stream-ordered: the current stream waits, the host does not

        first_params(reduce_scatter[0][0], send_buf.data_ptr())
        for hop, (send_i, recv_i) in enumerate(reduce_scatter):
            quantize_to(send_i, send_buf.data_ptr() + meta, send_buf.data_ptr(), REVERSE if hop else 0)
            exchange(send_buf, send_i, recv_buf, recv_i)
            accumulate(recv_i, recv_buf.data_ptr(), send_buf.data_ptr(), 0)     # next hop's parameters land in the send buffer's header
            yield
        quantize_to(own, send_buf.data_ptr() + meta, send_buf.data_ptr(), REVERSE)
        scatter(own, send_buf.data_ptr())            # the owner keeps exactly what everybody else will receive
        for send_i, recv_i in all_gather:
            exchange(send_buf, send_i, recv_buf, recv_i)
            scatter(recv_i, recv_buf.data_ptr())
            send_buf, recv_buf, spare = recv_buf, spare, send_buf     # forward what was just received
            yield

    flat = tensor.view(-1)
    use_p2p = transport == "p2p"
    if transport == "auto":
        use_p2p = _peer_memory_available(tensor.device, group)
    lanes = max(1, int(lanes)) if use_p2p else 1
    main = torch.cuda.current_stream(tensor.device)
    if lanes == 1 or flat.numel() < lanes * world * SHARD_ALIGN:
        for _ in ring(flat, main.cuda_stream, 0, use_p2p):
            pass
        return tensor
    # several lanes: contiguous parts, one ring and one stream each, hops enqueued alternately
    per = flat.numel() // lanes // SHARD_ALIGN * SHARD_ALIGN
    parts = [flat[i * per: (i + 1) * per if i < lanes - 1 else flat.numel()] for i in range(lanes)]
    streams = [main] + [_side_stream(tensor.device, i) for i in range(1, lanes)]
    for st in streams[1:]:
        st.wait_stream(main)
    gens = []
    for i, (part, st) in enumerate(zip(parts, streams)):
        gens.append((st, ring(part, st.cuda_stream, i, True)))
    while gens:
        for st, g in list(gens):
            with torch.cuda.stream(st):          # torch-side work of the lane (barriers, header copies) goes to the lane's stream too
                try:
                    next(g)
                except StopIteration:
                    gens.remove((st, g))
    for st in streams[1:]:
        main.wait_stream(st)
    return tensor


def gather_pieces(numel: int, pieces: int):
    """[lo, hi) element ranges that cut a chunk of ``numel`` elements into at most ``pieces`` non-empty parts at multiples of 256
    elements (64 packed bytes at 2 bits: every part starts vector-aligned on both sides).  A pure function of its arguments, so
    every rank derives the same table for every chunk."""
    cuts = sorted({min(numel, (numel * k // pieces) // 256 * 256) for k in range(pieces)} | {numel})
    return [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


_COPY_STREAMS: dict = {}


def _copy_streams(device: torch.device, lane: int, n: int):
    """streams that carry nothing but copy-engine transfers"""
    have = _COPY_STREAMS.setdefault((device.index, lane), [])
    while len(have) < n:
        have.append(torch.cuda.Stream(device=device))
    return have[:n]


class _DirectPlan:
    """Quantized all-reduce as two all-to-all exchanges over NVSwitch (every GPU reaches every peer at full link rate).

    Chunk c of the tensor (``shard_bounds``) is owned by rank c.

    1. scatter: for every other rank j, ONE launch computes min/max + parameters of my chunk j, one launch quantizes it
       into a local staging slot ``[64-byte parameter block | packed payload]``, and a COPY ENGINE moves the slot into rank
       j's receive slot number ``rank`` and then writes rank j's arrival flag for it -- two ``cudaMemcpyAsync`` on the
       plan's copy stream (``piquant_cuda_copy_on_stream``), no kernel: the SMs go on with chunk j + 1 while chunk j is on
       the wire, and the copies follow each other without a gap (measured on 8 B200s: one copy in flight per GPU moves
       568 GB/s out of every GPU at once, two or more in flight 450-510 GB/s -- profiles/r2_allreduce_probe_n8.txt).
    2. reduce: once the world-1 flags are up, ONE kernel folds the received slots into my own float chunk in rank order
       -- exactly world-1 successive dequantize(ADD) calls, in one pass -- and produces the parameters of the sums
       (``piquant_cuda_dequantize_sum_minmax_on_stream``); the sums are quantized once and the copy engine broadcasts
       ``[parameters | packed sums]`` to every peer's gather slot number ``rank`` (+ flag) while this rank dequantizes its own.
    3. gather: every rank dequantizes (SET) each gathered slot as soon as its flag is up, in the order the slots arrive.
       Owners dequantize the same bytes they sent, so all ranks end with bit-identical values.

    Every element is quantized twice whatever the world size (the ring quantizes the running sum at every hop), one
    barrier per call (nobody may still be reading the slots of the previous call; it also starts the ranks in lockstep)
    instead of 2 * (world - 1), and nothing synchronises with the host.

    ``lanes``: the tensor is cut into that many contiguous parts, each an independent all-reduce with its own slots and
    flags, run STAGGERED on the same two streams: scatter A, scatter B, reduce A, reduce B, gather A, gather B.  The
    copy stream is a FIFO, so the link carries B's scatter while the SMs reduce A, and A's broadcast while they reduce
    B: the link -- the bound of this collective -- never waits for a kernel except at the very start."""

    CH_BARRIER, CH_SCATTER, CH_GATHER = 0, 1, 2             # signal-pad channels (torch symmetric memory: one u32 per channel and rank)

    GATHER_PIECES = 4                                       # multicast gather: pieces per reduced chunk, each with its own arrival flag

    def __init__(self, numel: int, float_dtype: torch.dtype, dtype: torch.dtype, device: torch.device, group, ctx: Context, rmode: RoundMode,
                 lanes: int = 1, multicast: Optional[bool] = None):
        """Everything that allocates or rendezvouses happens here, so that ``enqueue`` only launches (it may run inside a
        CUDA graph capture)."""
        self.group, self.ctx, self.rmode, self.dev, self.numel, self.float_dtype = group, ctx, rmode, device, numel, float_dtype
        self.barrier_on_main = True
        self.trace = None                                   # a list: CUDA events at the phase boundaries of the next enqueue (tools/allreduce_probe.py)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.fdt, self.qdt = torch_to_piquant_dtype(float_dtype), torch_to_piquant_dtype(dtype)
        world, rank, meta = self.world, self.rank, Context.META_BYTES
        self.others = [(rank + d) % world for d in range(1, world)]     # staggered: at any moment every rank receives from one sender
        self.arrivals = [(rank - d) % world for d in range(1, world)]   # ... so this is the order in which slots arrive here
        lanes = max(1, int(lanes))
        if numel < lanes * world * SHARD_ALIGN:
            lanes = 1
        per = numel // lanes // SHARD_ALIGN * SHARD_ALIGN
        self.parts = [(i * per, (i + 1) * per if i < lanes - 1 else numel) for i in range(lanes)]
        self.copy_stream = _copy_streams(device, 0, 1)[0]
        self.one = torch.ones(1, dtype=torch.int32, device=device)      # what a raised flag holds
        self.lanes = []
        for lane, (p0, p1) in enumerate(self.parts):
            bounds = [shard_bounds(p1 - p0, world, i) for i in range(world)]
            qbytes = [self.qdt.storage_bytes(e - b) for b, e in bounds]
            slot_bytes = (meta + max(qbytes) + 255) // 256 * 256
            # symmetric memory: [world scatter slots | world gather slots]; slot k of the first half receives from
            # rank k, slot k of the second half holds the reduced chunk k
            # ... followed by 4 KB of arrival flags for the multicast gather: flag [owner][piece], one u32 each
            local, hdl = _p2p_slots(world * slot_bytes + 2048, device, group, lane=-2 - lane)       # (allocates 2 * nbytes: the two halves)
            flags_off = 2 * world * slot_bytes
            local[flags_off:].zero_()
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            pieces = [gather_pieces(e - b, self.GATHER_PIECES) for b, e in bounds]
            self.lanes.append(dict(bounds=bounds, qbytes=qbytes, slot_bytes=slot_bytes, local=local, hdl=hdl,
                                   peer_base=[int(hdl.buffer_ptrs[i]) for i in range(world)],
                                   peer_pad=[int(hdl.signal_pad_ptrs[i]) for i in range(world)],
                                   mc=mc, flags_off=flags_off, pieces=pieces,
                                   stage=torch.empty(world * slot_bytes, dtype=torch.uint8, device=device)))
        # NVSwitch multicast for the gather exchange (opt-in): one copy-engine transfer per piece reaches every rank.  A single
        # 33.5 MB multicast per rank delivers 0.75-0.85 TB/s INTO every GPU against 0.48-0.56 TB/s for seven unicast copies, but
        # all seven slots then arrive together and the dequantizes queue up behind them; cut into 4 pieces with their own flags
        # (so that dequantizing overlaps the arrivals) the exchange measured SLOWER than the unicast form at 8 GPUs -- 1.12 ms
        # against 1.08 ms for the whole collective, the first piece landing 400 us after the reduce
        # (profiles/r2_allreduce_probe_n8_multicast_ab.txt) -- so the default stays unicast.
        have_mc = all(L["mc"] for L in self.lanes) and world * self.GATHER_PIECES * 4 <= 2048
        self.multicast = bool(multicast) and have_mc
        torch.cuda.synchronize(device)

    def _lane_phases(self, flat: torch.Tensor, lane: int, main: "torch.cuda.Stream"):
        """generator: enqueues one all-reduce of `flat` on `main` and the copy stream, yielding between its three phases"""
        L = self.lanes[lane]
        ctx, world, rank, fdt, qdt, rmode = self.ctx, self.world, self.rank, self.fdt, self.qdt, self.rmode
        meta, device, st, side = Context.META_BYTES, self.dev.index, main.cuda_stream, self.copy_stream
        LOCAL, REVERSE = Context.FLAG_LOCAL, Context.FLAG_REVERSE
        bounds, qbytes, slot_bytes, hdl, peer_base, peer_pad = L["bounds"], L["qbytes"], L["slot_bytes"], L["hdl"], L["peer_base"], L["peer_pad"]
        my_base, stage = L["local"].data_ptr(), L["stage"].data_ptr()
        rs_off = lambda k: k * slot_bytes                                      # noqa: E731
        ag_off = lambda k: (world + k) * slot_bytes                            # noqa: E731

        def chunk(i):
            b, e = bounds[i]
            return flat[b:e]

        def send(src_ptr, dst_rank, dst_off, nbytes, channel):
            """copy engine: payload, then the arrival flag (stream order = arrival order; flag word [channel][rank] of dst's pad)"""
            ctx.copy_on_stream(peer_base[dst_rank] + dst_off, src_ptr, nbytes, device, side.cuda_stream)
            ctx.copy_on_stream(peer_pad[dst_rank] + 4 * (world * channel + rank), self.one.data_ptr(), 4, device, side.cuda_stream)

        def copies_follow_main():
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)

        def mark(label):
            if self.trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(main)
                self.trace.append((lane, label, ev))

        mark("start")
        # Nobody may still be reading the slots of a previous call when the first copy of this one lands: one barrier.  It sits on
        # the MAIN stream although only the copies need it: it also starts the ranks in lockstep, which is what keeps the staggered
        # schedule staggered (every receiver fed by one sender at a time).  Measured on 8 GPUs in one run: barrier on the main
        # stream 1.06-1.08 ms, on the copy stream (first chunk quantized while the ranks meet) 1.18-1.23 ms -- the skewed ranks'
        # copies collide at the receivers (profiles/r2_allreduce_probe_n8_barrier_ab.txt).
        if self.barrier_on_main:
            hdl.barrier(channel=self.CH_BARRIER)
        else:                                                # (A/B switch of tools/allreduce_probe.py)
            with torch.cuda.stream(side):
                hdl.barrier(channel=self.CH_BARRIER)
        mark("barrier")
        for k, j in enumerate(self.others):
            c = chunk(j)
            if c.numel():
                base = stage + k * slot_bytes
                ctx.compute_meta_on_stream(c.data_ptr(), fdt, c.numel(), qdt, base, LOCAL, device, st)
                ctx.quantize_meta_on_stream(c.data_ptr(), fdt, base + meta, qdt, c.numel(), rmode, base, REVERSE, device, st)
                copies_follow_main()
                send(base, j, rs_off(rank), meta + qbytes[j], self.CH_SCATTER)
                mark(f"quantized chunk {j}")
        yield
        mine = chunk(rank)
        # where [parameters | packed sums] of my chunk are produced: my own gather slot, or -- the multicast transfer writes EVERY
        # replica of that slot, mine included, and must not copy onto itself -- a local staging slot
        own_slot = stage + (world - 1) * slot_bytes if self.multicast else my_base + ag_off(rank)
        if mine.numel():
            for k in self.arrivals:
                hdl.wait_signal(k, self.CH_SCATTER)        # my scatter slots are complete (the kernel lowers the flag again)
            mark("all scatter slots arrived")
            srcs = [my_base + rs_off(k) for k in range(world) if k != rank]
            for g in range(0, len(srcs), Context.MAX_SUM_SOURCES):       # one launch up to 9 ranks; the last launch's parameters are the sums'
                part = srcs[g:g + Context.MAX_SUM_SOURCES]
                ctx.dequantize_sum_minmax_on_stream([p + meta for p in part], qdt, mine.data_ptr(), fdt, mine.numel(), part, qdt, own_slot, 0,
                                                    device, st)
            ctx.quantize_meta_on_stream(mine.data_ptr(), fdt, own_slot + meta, qdt, mine.numel(), rmode, own_slot, REVERSE, device, st)
            mark("reduced + quantized own chunk")
            copies_follow_main()
            if self.multicast:
                mc_slot, mc_flags = L["mc"] + ag_off(rank), L["mc"] + L["flags_off"] + 4 * rank * self.GATHER_PIECES
                for p, (lo, hi) in enumerate(L["pieces"][rank]):
                    b0 = 0 if p == 0 else meta + qdt.storage_bytes(lo)
                    b1 = meta + qdt.storage_bytes(hi)
                    ctx.copy_on_stream(mc_slot + b0, own_slot + b0, b1 - b0, device, side.cuda_stream)     # ONE transfer, every rank's slot
                    ctx.copy_on_stream(mc_flags + 4 * p, self.one.data_ptr(), 4, device, side.cuda_stream)  # ... then every rank's flag
            else:
                for j in self.others:
                    send(own_slot, j, ag_off(rank), meta + qbytes[rank], self.CH_GATHER)
            # the owner takes the dequantized values of exactly the bytes everybody else receives
            ctx.dequantize_meta_on_stream(own_slot + meta, qdt, mine.data_ptr(), fdt, mine.numel(), ReduceOp.SET, own_slot, device, st)
        yield
        if self.multicast:
            # all owners broadcast at once, so piece p of every chunk arrives at about the same time: dequantize round by round
            my_flags = my_base + L["flags_off"]
            for p in range(self.GATHER_PIECES):
                for j in self.arrivals:
                    if p < len(L["pieces"][j]):
                        lo, hi = L["pieces"][j][p]
                        src = my_base + ag_off(j)
                        ctx.wait_flag_on_stream(my_flags + 4 * (j * self.GATHER_PIECES + p), device, st)
                        ctx.dequantize_meta_on_stream(src + meta + qdt.storage_bytes(lo), qdt, chunk(j)[lo:hi].data_ptr(), fdt, hi - lo,
                                                      ReduceOp.SET, src, device, st)
                mark(f"dequantized piece {p} of every chunk")
            return
        for j in self.arrivals:
            c = chunk(j)
            if c.numel():
                src = my_base + ag_off(j)
                hdl.wait_signal(j, self.CH_GATHER)         # gather slot j is complete
                ctx.dequantize_meta_on_stream(src + meta, qdt, c.data_ptr(), fdt, c.numel(), ReduceOp.SET, src, device, st)
                mark(f"dequantized chunk {j}")

    def enqueue(self, tensor: torch.Tensor) -> torch.Tensor:
        """Launch the collective on the current stream (and the plan's copy stream); no allocation, no synchronisation."""
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.numel() == self.numel and tensor.dtype == self.float_dtype
        assert tensor.device == self.dev
        flat = tensor.view(-1)
        main = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(main)                 # (a capture needs the fork; eagerly: staging slots of the previous call are free)
        gens = [self._lane_phases(flat[p0:p1], i, main) for i, (p0, p1) in enumerate(self.parts)]
        while gens:                                        # phase by phase, lane after lane: scatter A, scatter B, reduce A, ...
            for g in list(gens):
                try:
                    next(g)
                except StopIteration:
                    gens.remove(g)
        main.wait_stream(self.copy_stream)                 # the staging slots and my gather slots are free again
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(main)
            self.trace.append((0, "copy stream joined", ev))
        return tensor


_DIRECT_PLANS: dict = {}


def _direct_all_reduce_nccl(tensor: torch.Tensor, dtype: torch.dtype, group, ctx: Context, rmode: RoundMode) -> torch.Tensor:
    """The direct algorithm (``_DirectPlan``: chunk c owned by rank c, every value quantized twice, one-pass multi-source
    reduce) with its two exchanges handed to NCCL: ONE ``all_to_all_single`` of ``[parameters | packed chunk]`` slots and ONE
    ``all_gather_into_tensor`` of the reduced slots.  Same arithmetic in the same order as the peer-memory form -- results are
    bit-identical to it and to the CPU replay -- for boxes where the GPUs cannot map each other's memory; nothing overlaps
    (NCCL needs all of a rank's slots before it starts), so the peer-memory form is the faster one where it exists."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    fdt, qdt = torch_to_piquant_dtype(tensor.dtype), torch_to_piquant_dtype(dtype)
    meta, dev = Context.META_BYTES, tensor.device
    device, st = dev.index, torch.cuda.current_stream(dev).cuda_stream
    LOCAL, REVERSE = Context.FLAG_LOCAL, Context.FLAG_REVERSE
    flat = tensor.view(-1)
    bounds = [shard_bounds(flat.numel(), world, i) for i in range(world)]
    qbytes = [qdt.storage_bytes(e - b) for b, e in bounds]
    slot_bytes = (meta + max(qbytes) + 255) // 256 * 256
    send = torch.empty(world * slot_bytes, dtype=torch.uint8, device=dev)      # slot j: my chunk j for its owner, rank j
    recv = torch.empty_like(send)                                              # slot k: rank k's version of MY chunk
    for j in range(world):
        b, e = bounds[j]
        if j != rank and e > b:
            base = send.data_ptr() + j * slot_bytes
            ctx.compute_meta_on_stream(flat[b:e].data_ptr(), fdt, e - b, qdt, base, LOCAL, device, st)
            ctx.quantize_meta_on_stream(flat[b:e].data_ptr(), fdt, base + meta, qdt, e - b, rmode, base, REVERSE, device, st)
    dist.all_to_all_single(recv, send, group=group)
    b, e = bounds[rank]
    mine = torch.empty(slot_bytes, dtype=torch.uint8, device=dev)              # [parameters | packed sums] of my chunk
    gathered = send                                                            # (its slots have been delivered: reuse the buffer)
    if e > b:
        srcs = [recv.data_ptr() + k * slot_bytes for k in range(world) if k != rank]
        for g in range(0, len(srcs), Context.MAX_SUM_SOURCES):
            part = srcs[g:g + Context.MAX_SUM_SOURCES]
            ctx.dequantize_sum_minmax_on_stream([p + meta for p in part], qdt, flat[b:e].data_ptr(), fdt, e - b, part, qdt, mine.data_ptr(), 0,
                                                device, st)
        ctx.quantize_meta_on_stream(flat[b:e].data_ptr(), fdt, mine.data_ptr() + meta, qdt, e - b, rmode, mine.data_ptr(), REVERSE, device, st)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    for j in range(world):                                                     # owners too: every rank dequantizes the same bytes
        bj, ej = bounds[j]
        if ej > bj:
            src = gathered.data_ptr() + j * slot_bytes
            ctx.dequantize_meta_on_stream(src + meta, qdt, flat[bj:ej].data_ptr(), fdt, ej - bj, ReduceOp.SET, src, device, st)
    return tensor


def _auto_lanes(tensor: torch.Tensor, world: int) -> int:
    """Two staggered lanes pay on 2 GPUs for large tensors (2^28 f32: 1.01 -> 0.76-0.82 ms); on an NVSwitch box with more ranks the
    halved copies lose more link efficiency than the overlap wins (8 GPUs: 1.06 -> 1.18 ms), and below ~256 MB the second lane only
    doubles the host work (profiles/r2_allreduce_probe_n2.txt, _n8.txt)."""
    return 2 if world == 2 and tensor.numel() * tensor.element_size() >= (256 << 20) else 1


def _direct_all_reduce(tensor: torch.Tensor, dtype: torch.dtype, group, ctx: Context, rmode: RoundMode, lanes: int = 1,
                       multicast: Optional[bool] = None) -> torch.Tensor:
    grp = group if group is not None else dist.group.WORLD
    key = (grp.group_name, tensor.device.index, tensor.numel(), tensor.dtype, dtype, id(ctx), rmode, lanes, multicast)
    plan = _DIRECT_PLANS.get(key)
    if plan is None:
        plan = _DIRECT_PLANS[key] = _DirectPlan(tensor.numel(), tensor.dtype, dtype, tensor.device, group, ctx, rmode, lanes, multicast)
    return plan.enqueue(tensor)


class QuantizedAllReduce:
    """The direct quantized all-reduce of ONE persistent tensor (a gradient bucket), captured into a CUDA graph.

    The collective is ~75 launches, copies and flag operations per lane; enqueued from Python they cost more host time than
    the GPU needs for tensors below ~1 GB.  Captured once, ``plan()`` replays the whole exchange -- kernels, copy-engine
    transfers into peer memory, flags -- with one graph launch on the current stream.  Every rank must construct and call
    the object in the same order (construction is a collective: symmetric-memory rendezvous).  Nearest rounding only: a
    captured launch would replay the same random stream.  Plans of equal size share their receive slots (the barrier at
    the start of every call keeps successive calls apart): replay them on ONE stream, like any collective of a group."""

    def __init__(self, tensor: torch.Tensor, *, dtype: torch.dtype = torch.quint8, group: Optional[dist.ProcessGroup] = None,
                 ctx: Context = Context.get(), lanes: Optional[int] = None, multicast: Optional[bool] = None):
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.dtype in (torch.float32, torch.bfloat16)
        assert dtype in _QUANT_TYPES
        self.tensor = tensor
        self.world = dist.get_world_size(group)
        if self.world == 1:
            self.graph = None
            return
        lanes = lanes if lanes is not None else _auto_lanes(tensor, self.world)
        self.plan = _DirectPlan(tensor.numel(), tensor.dtype, dtype, tensor.device, group, ctx, RoundMode.NEAREST, lanes, multicast)
        # one tiny eager call: the library's per-device state (slot table, two allocations) exists before the capture starts
        warm = torch.zeros(SHARD_ALIGN * self.world, dtype=tensor.dtype, device=tensor.device)
        meta = torch.zeros(Context.META_BYTES, dtype=torch.uint8, device=tensor.device)
        device, stream = _site(warm)
        ctx.compute_meta_on_stream(warm.data_ptr(), torch_to_piquant_dtype(tensor.dtype), warm.numel(), torch_to_piquant_dtype(dtype),
                                   meta.data_ptr(), Context.FLAG_LOCAL, device, stream)
        torch.cuda.synchronize(tensor.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.plan.enqueue(tensor)

    def __call__(self) -> torch.Tensor:
        if self.graph is not None:
            self.graph.replay()
        return self.tensor
