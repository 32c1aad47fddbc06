"""The CPU oracle (SEM_BODY, the semantics the CUDA library implements) against the committed golden
vectors produced by the unmodified reference (tests/golden/make_golden.py).  Runs anywhere: it needs
neither /root/reference nor oracle/_ref."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest

from oracle import port
from oracle.port import ADD, BF16, F32, NEAREST, SEM_BODY, SET, STOCHASTIC, UINT2, UINT4, UINT8

GOLDEN = Path(__file__).parent / "golden" / "piquant_golden.npz"
DT = {"f32": F32, "bf16": BF16, "u2": UINT2, "u4": UINT4, "u8": UINT8}


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def keys(prefix: str) -> list[str]:
    return [str(k) for k in np.load(GOLDEN)["__keys__"] if str(k).startswith(prefix)]


@pytest.mark.parametrize("key", keys("quant/"))
def test_quantize_matches_golden(golden, key):
    _, dti, dto, mode, n = key.split("/")
    x, out = golden[key + "/x"], golden[key + "/out"]
    scale, zp, xi = golden[key + "/p"]
    got = port.quantize(x, DT[dto], float(scale), int(zp), STOCHASTIC if mode == "st" else NEAREST,
                        xi=float(xi), semantics=SEM_BODY)
    assert np.array_equal(got, out)


@pytest.mark.parametrize("key", keys("dequant/"))
def test_dequantize_matches_golden(golden, key):
    _, dti, dto, op, n = key.split("/")
    q, prev, out = golden[key + "/q"], golden[key + "/prev"], golden[key + "/out"]
    scale, zp = golden[key + "/p"]
    got = port.dequantize(q, DT[dti], int(n), DT[dto], float(scale), int(zp), ADD if op == "add" else SET,
                          out=prev.copy(), semantics=SEM_BODY)
    if dto == "f32":
        assert np.array_equal(got, out)
    else:
        # bf16 outputs: the golden holds the reference's scalar-tail results for the last n % body-width
        # elements (different rounding order, see test_oracle_vs_reference.py); one bf16 ulp at most
        a = port.bf16_bits_to_f32(got).astype(np.float64)
        b = port.bf16_bits_to_f32(out).astype(np.float64)
        qmax = (1 << port.BITS[DT[dti]]) - 1
        assert (np.abs(a - b) <= (np.maximum(np.abs(a), np.abs(b)) + qmax * scale) * 2.0**-7 + scale * 1e-6).all()
        body = {"u8": 64, "u4": 128, "u2": 256}[dti]
        nb = (int(n) // 4 // body) * body     # elements surely inside a SIMD body of the 4-thread reference run
        assert np.array_equal(got[:nb], out[:nb])


@pytest.mark.parametrize("key", keys("params/"))
def test_quant_params_match_golden(golden, key):
    dtq = key.split("/")[-1]
    x = golden[key + "/x"]
    bits, zp = golden[key + "/p"]
    s, z = port.compute_quant_params(x, DT[dtq])
    assert int(np.float32(s).view(np.uint32)) == int(bits) and z == int(zp)


def test_known_answers_from_survey():
    """values probed from the reference build."""
    pm1 = np.array([-1, 1], np.float32)
    assert port.compute_quant_params(pm1, UINT4) == (pytest.approx(0.13333334028720856, abs=0), 8)
    assert port.compute_quant_params(pm1, UINT2) == (pytest.approx(0.6666666865348816, abs=0), 2)
    assert port.compute_quant_params(np.array([-3, 5, 1], np.float32), UINT8) == (pytest.approx(0.0313725508749485, abs=0), 96)
    c = np.full(10, 42.0, np.float32)
    assert [port.compute_quant_params(c, d) for d in (UINT8, UINT4, UINT2)] == [(1.0, 127), (1.0, 7), (1.0, 1)]
    assert port.compute_quant_params(np.array([1, 2], np.float32), UINT8) == (pytest.approx(0.003921568859368563, abs=0), 0)
    with pytest.raises(ValueError):
        port.compute_quant_params(np.zeros(0, np.float32), UINT8)      # reference aborts on empty input
