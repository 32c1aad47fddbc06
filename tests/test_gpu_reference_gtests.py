"""The reference's OWN gtest suites (reference test/quant.cpp, dequant.cpp, requant.cpp, quant_config.cpp with
test/naive.hpp as their scalar oracle), compiled unmodified against this repo's include/piquant.hpp and linked
against libpiquant.so (recipe: `make -C oracle reftests`, output oracle/_ref/piquant_ref_tests_b200 -- built in
the dev container where /root/reference exists, shipped to the GPU box as a binary, never committed).

All 65 cases of the reference's suite must pass on the B200 library: quantize uint4 exact / uint8 +-1 against
quantize_naive, 24 dequantize round trips, 20 fused requantize round trips, 12 x 100 random-range runs, the
identity KAT.  Their buffers are std::vectors, so this also exercises the host-pointer pipeline end to end."""
from __future__ import annotations

import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu


def test_reference_gtest_suite_passes_on_the_cuda_library():
    from oracle import REF_TESTS

    if not REF_TESTS.exists():
        pytest.skip("oracle/_ref/piquant_ref_tests_b200 not built (needs /root/reference at build time)")
    r = subprocess.run([str(REF_TESTS), "--gtest_color=no"], capture_output=True, text=True, timeout=1200)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    m = re.search(r"\[  PASSED  \] (\d+) tests", r.stdout)
    assert m and int(m.group(1)) == 65, tail
    assert "FAILED" not in r.stdout
