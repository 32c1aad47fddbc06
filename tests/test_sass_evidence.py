"""What the shipped cubin contains (runs on the CPU: cuobjdump disassembles libpiquant.so without a GPU).

B200_PROFILING.md, "What proves a Blackwell-native kernel": the PTX names never appear in SASS, these do --
UBLKCP.S.G / UBLKCP.G.S  cp.async.bulk (TMA, 1-D bulk copies) of the ring kernels, SYNCS.* their mbarriers,
LDG.E...256 / STG.E...256  32-byte global accesses (one full DRAM sector per thread), sm_100 only,
I2IP.U{8,4,2}.S32.SAT    clamp + pack of two quantized elements in one instruction,
FMNMX3.NAN               3-input NaN-propagating max: the range witness of the speculative group quantize,
ACQBULK / PREEXIT        griddepcontrol.wait / launch_dependents (programmatic dependent launch)."""
from __future__ import annotations

import re
import shutil
import subprocess
import sys
from collections import Counter
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "pi-quant_b200" / "piquant" / "libpiquant.so"


@pytest.fixture(scope="module")
def sass() -> str:
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    sys.path.insert(0, str(ROOT / "pi-quant_b200"))
    import build as pq_build

    pq_build.build()
    return subprocess.run([exe, "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout


def test_only_sm_100a_code_is_shipped(sass):
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs


def test_blackwell_mnemonics_present(sass):
    ops = Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", sass, flags=re.M))
    def count(prefix):
        return sum(v for k, v in ops.items() if k.startswith(prefix))
    assert count("UBLKCP.S.G") >= 18 and count("UBLKCP.G.S") >= 18          # 18 quantize + 12 dequantize TMA kernels
    assert count("SYNCS.ARRIVE.TRANS64") > 0 and count("SYNCS.PHASECHK.TRANS64.TRYWAIT") > 0
    assert any(re.match(r"LDG\.E\..*256", k) for k in ops), "no 256-bit global loads"
    assert any(re.match(r"STG\.E\..*256", k) for k in ops), "no 256-bit global stores"
    for bits in (8, 4, 2):
        assert count(f"I2IP.U{bits}.S32.SAT") > 0
    assert count("FMNMX3.NAN") > 0
    assert count("ACQBULK") > 0 and count("PREEXIT") > 0
    assert count("HMMA") == 0 and count("UTC") == 0                          # elementwise path: no tensor-core instructions


def test_no_kernel_uses_local_memory_outside_the_cold_fallbacks():
    """Streaming kernels must not spill: local-memory traffic (STL / LDL) would show up as extra DRAM bytes.
    The only stack frames allowed are the few bytes the EXACT FALLBACK of the bf16 stochastic kernels spills since those
    kernels were capped at 32 registers (8 CTAs per SM): the fallback only runs for vectors holding
    NaN / inf / huge values, the speculative path of the same kernels has no local-memory instruction."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(exe).exists():
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout
    frames = dict(re.findall(r"Function (\S+):\s*\n\s*REG:\d+ STACK:(\d+)", out))
    assert len(frames) > 100, "cuobjdump -res-usage output not understood"
    allowed = re.compile(r"quant_(stream|batch)_kernelILi1ELi[248]ELi2E|quant_bf16_u2_threshold_kernelILi2E")   # <bf16, *, stochastic> (the batch kernel runs the same tile code)
    for fn, stack in frames.items():
        if allowed.search(fn):
            assert int(stack) <= 64, (fn, stack)
        else:
            assert int(stack) == 0, (fn, stack)


def test_speculative_paths_of_the_capped_kernels_do_not_touch_local_memory(sass):
    """In the kernels that may spill, every STL / LDL sits in the exact fallback: none between the vector loads of a tile and
    the first packed store that follows them."""
    for fn in ("_ZN2pq19quant_stream_kernelILi1ELi4ELi2EEEvNS_9QuantArgsE", "_ZN2pq30quant_bf16_u2_threshold_kernelILi2EEEvNS_9QuantArgsE"):
        body = sass.split(f"Function : {fn}")[1].split("Function :")[0]
        ops = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", body, flags=re.M)
        first_load = next(i for i, o in enumerate(ops) if o.startswith("LDG.E") and "256" in o)
        bras = [i for i, o in enumerate(ops) if i > first_load and o == "BRA"]
        fast_end = bras[1]      # loads, speculative steps, witness branch to the fallback, packing, jump over the fallback
        assert not [o for o in ops[first_load:fast_end] if o.startswith(("STL", "LDL"))]
        assert any(o.startswith(("I2IP", "HSET2")) for o in ops[first_load:fast_end])
