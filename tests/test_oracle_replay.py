"""The CPU replays of the quantized all-reduce algorithms (oracle/replay.py) are what the 2-GPU tests and the bench compare
the GPU collectives with bit for bit; here they are checked against the exact sum (CPU only)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "pi-quant_b200"))

from oracle import port as orc, replay  # noqa: E402


def _shard_bounds(numel, world, rank, align=64):
    # piquant.distributed.shard_bounds restated (importing the package needs the CUDA library; tests/test_sharding_gloo.py covers the original)
    per = (numel // world) // align * align
    return per * rank, (numel if rank == world - 1 else per * (rank + 1))


@pytest.mark.parametrize("world", (2, 3, 8))
@pytest.mark.parametrize("lanes", (1, 2))
@pytest.mark.parametrize("fn", (replay.direct_all_reduce, replay.ring_all_reduce), ids=("direct", "ring"))
def test_replay_is_the_sum_up_to_quantization_error(world, lanes, fn):
    rng = np.random.default_rng(world * 10 + lanes)
    n = 50_003
    inputs = [rng.uniform(-1, 1, n).astype(np.float32) for _ in range(world)]
    exact = np.sum(np.stack(inputs).astype(np.float64), axis=0)
    got = fn([i.copy() for i in inputs], orc.UINT8, orc.F32, _shard_bounds, 64, lanes).view(np.float32)
    assert got.shape == (n,)
    step_in, step_sum = 2.0 / 255, 2.0 * world / 255
    if fn is replay.direct_all_reduce:          # every input rounded once at its own scale, the sum once more
        bound = 0.5 * ((world - 1) * step_in + step_sum) + 1e-5
    else:                                       # the running sum is rounded at every hop, at most at the final scale
        bound = 0.5 * world * step_sum + 1e-5
    assert np.abs(got - exact).max() <= bound


def test_direct_replay_small_tensor_falls_back_to_one_lane_and_handles_empty_chunks():
    rng = np.random.default_rng(1)
    inputs = [rng.uniform(-1, 1, 100).astype(np.float32) for _ in range(2)]
    a = replay.direct_all_reduce([i.copy() for i in inputs], orc.UINT8, orc.F32, _shard_bounds, 64, 1)
    b = replay.direct_all_reduce([i.copy() for i in inputs], orc.UINT8, orc.F32, _shard_bounds, 64, 2)
    assert np.array_equal(a, b)                 # 100 < 2 lanes * 2 ranks * 64: one lane; chunk 0 is empty, rank 1 owns everything
    assert np.abs(a.view(np.float32) - (inputs[0] + inputs[1])).max() <= 0.5 * (2 / 255 + 4 / 255) + 1e-5
