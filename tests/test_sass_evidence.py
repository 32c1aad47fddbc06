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


def test_no_kernel_uses_local_memory(sass):
    """Streaming kernels must not spill: local-memory traffic (STL / LDL) would show up as extra DRAM bytes."""
    local = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?((?:STL|LDL)[A-Z0-9_.]*)", sass, flags=re.M)
    assert not local, Counter(local)
