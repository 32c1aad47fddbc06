"""CPU oracle for the pi-quant hot path -- TEST INFRASTRUCTURE ONLY.

Two checkers live here, neither is ever imported by the product (``pi-quant_b200/``):

* ``oracle.port``  -- ctypes binding of ``liboracle.so`` (``piquant_oracle.c``), our plain-C
  restatement of the reference's arithmetic.  Travels everywhere, needs only gcc.
* ``oracle.ref``   -- ctypes binding of ``_ref/libpiquant_ref.so``, the UNMODIFIED reference
  compiled from ``/root/reference`` by ``oracle/Makefile``.  Used to pin the port and as the CPU
  baseline of ``bench.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
PORT_LIB = HERE / "liboracle.so"
REF_LIB = HERE / "_ref" / "libpiquant_ref.so"
REFERENCE_SRC = Path(os.environ.get("PIQUANT_REFERENCE_SRC", "/root/reference"))


def build(ref: bool | None = None, quiet: bool = True) -> None:
    """Compile the C restatement, and the real reference when its sources are present.

    ``ref=None`` builds the reference only if ``/root/reference`` exists and the library is
    missing (the GPU box has no ``/root/reference``; it uses the prebuilt file that travelled
    with the snapshot).
    """
    out = subprocess.DEVNULL if quiet else None
    src = HERE / "piquant_oracle.c"
    if not PORT_LIB.exists() or PORT_LIB.stat().st_mtime < max(src.stat().st_mtime, (HERE / "piquant_oracle.h").stat().st_mtime):
        subprocess.check_call(["make", "-C", str(HERE), "oracle"], stdout=out)
    if ref is None:
        ref = (REFERENCE_SRC / "src" / "piquant.cpp").exists() and not REF_LIB.exists()
    if ref:
        subprocess.check_call(["make", "-C", str(HERE), "ref", f"REF={REFERENCE_SRC}"], stdout=out, stderr=out)
    # the reference's own gtest suites, compiled unmodified against this repo's piquant.hpp + libpiquant.so
    # (acceptance test of the drop-in boundary; runs on the GPU box, see tests/test_gpu_reference_gtests.py)
    lib = HERE.parent / "pi-quant_b200" / "piquant" / "libpiquant.so"
    if (REFERENCE_SRC / "test" / "quant.cpp").exists() and lib.exists():
        subprocess.check_call(["make", "-C", str(HERE), "reftests", f"REF={REFERENCE_SRC}"], stdout=out, stderr=out)


REF_TESTS = HERE / "_ref" / "piquant_ref_tests_b200"


def have_ref() -> bool:
    return REF_LIB.exists()
