"""CPU-side checks of the drop-in boundary: libpiquant.so builds, loads without a GPU, exports every
symbol that include/*.h declares with the reference's enum values, never initialises CUDA at context
creation, and fails LOUDLY (abort, like the reference's panic()) when asked to compute without a device."""
from __future__ import annotations

import ctypes
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "pi-quant_b200"
LIB = PKG / "piquant" / "libpiquant.so"


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, str(PKG))
    import build as pq_build      # pi-quant_b200/build.py

    pq_build.build()
    assert LIB.exists()
    return ctypes.CDLL(str(LIB))


def declared_functions(header: Path) -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return re.findall(r"\b(piquant_\w+)\s*\(", text)


def test_exports_every_declared_symbol(lib):
    names = declared_functions(ROOT / "include" / "piquant.h") + declared_functions(ROOT / "include" / "piquant_cuda.h")
    assert len(names) >= 21
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/ but not exported by libpiquant.so"
    # the reference ABI, name by name (reference include/piquant.h:42-85)
    for name in ("piquant_context_create", "piquant_context_destroy", "piquant_quantize", "piquant_dequantize",
                 "piquant_compute_quant_params_float32", "piquant_compute_quant_params_bfloat16"):
        assert name in names


def test_exports_nothing_else(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert exported and all(s.startswith("piquant_") for s in exported), exported


def test_enum_values_match_reference_abi():
    """reference include/piquant.h:23-40, locked by static_asserts in reference src/capi.cpp:9-13"""
    text = (ROOT / "include" / "piquant.h").read_text()
    for name, value in (("PIQUANT_NEAREST", 0), ("PIQUANT_STOCHASTIC", 1), ("PIQUANT_REDUCE_OP_SET", 0),
                        ("PIQUANT_REDUCE_OP_ADD", 1), ("PIQUANT_DTYPE_F32", 0), ("PIQUANT_DTYPE_BF16", 1),
                        ("PIQUANT_DTYPE_UINT2", 2), ("PIQUANT_DTYPE_UINT4", 3), ("PIQUANT_DTYPE_UINT8", 4)):
        assert re.search(rf"\b{name}\s*=\s*{value}\b", text), name


def test_python_package_mirrors_reference_surface(lib):
    import piquant
    from piquant import Context, DataType, ReduceOp, RoundMode

    # the reference's five values keep their ABI numbers (reference include/piquant.h:33-40); 5..7 are the signed extension
    assert [(m.name, m.value) for m in DataType][:5] == [("F32", 0), ("BF16", 1), ("UINT2", 2), ("UINT4", 3), ("UINT8", 4)]
    assert [(m.name, m.value) for m in DataType][5:] == [("INT2", 5), ("INT4", 6), ("INT8", 7)]
    assert DataType.INT4.is_signed and DataType.INT4.is_quantized and not DataType.UINT4.is_signed and DataType.INT4.storage_bytes(7) == 4
    assert RoundMode.NEAREST.value == 0 and RoundMode.STOCHASTIC.value == 1
    assert ReduceOp.SET.value == 0 and ReduceOp.ADD.value == 1
    assert DataType.UINT4.bit_size == 4 and DataType.UINT4.is_quantized and DataType.BF16.is_dequantized
    assert DataType.UINT2.storage_bytes(7) == 2 and DataType.UINT4.storage_bytes(7) == 4 and DataType.F32.storage_bytes(7) == 28
    for name in ("quantize_ptr", "dequantize_ptr", "compute_quant_params_ptr_float32", "compute_quant_params_ptr_bfloat16"):
        assert callable(getattr(Context, name))
    import piquant.torch as pt
    for name in ("compute_quant_params", "quantize", "dequantize", "torch_to_piquant_dtype", "piquant_to_torch_dtype"):
        assert callable(getattr(pt, name))
    ctx = Context(4)                      # no CUDA call: works on a machine without a GPU
    ctx.set_stochastic_threshold(0.25)
    ctx.seed(123)
    assert ctx.kernel_launches == 0
    assert Context.get() is Context.get()
    assert piquant.cuda_device_count() >= 0


def test_params_from_minmax_is_host_arithmetic(lib):
    """The double-precision scale / zero-point math runs on the host; check it against the oracle
    without any GPU (reference src/piquant.cpp:245-258)."""
    from oracle import port
    from piquant import Context, DataType

    dt = {port.UINT2: DataType.UINT2, port.UINT4: DataType.UINT4, port.UINT8: DataType.UINT8,
          port.INT2: DataType.INT2, port.INT4: DataType.INT4, port.INT8: DataType.INT8}
    cases = [(-1.0, 1.0), (0.0, 1.0), (1.0, 2.0), (-5.0, -1.0), (42.0, 42.0), (-3.0, 5.0), (-3e38, 3e38), (1e-30, 2e-30)]
    for mn, mx in cases:
        for d in dt:
            assert Context.params_from_minmax(mn, mx, dt[d]) == port.params_from_minmax(mn, mx, d)


_ABORT_SNIPPET = r"""
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
lib.piquant_context_create.restype = ctypes.c_void_p
ctx = lib.piquant_context_create(ctypes.c_size_t(1))
{call}
print("survived")
"""


def run_snippet(call: str) -> subprocess.CompletedProcess:
    return subprocess.run([sys.executable, "-c", _ABORT_SNIPPET.format(call=call), str(LIB)], capture_output=True, text=True)


def test_invalid_dtype_combination_aborts_like_the_reference(lib):
    """reference src/piquant.cpp:288-289: wrong dtype class -> panic() -> abort()."""
    buf = "b = ctypes.create_string_buffer(64)\n"
    r = run_snippet(buf + "lib.piquant_quantize(ctypes.c_void_p(ctx), b, 4, b, 0, ctypes.c_size_t(8), ctypes.c_float(1.0), ctypes.c_int64(0), 0)")
    assert r.returncode == -6 and "must be a dequantized type" in r.stderr and "survived" not in r.stdout


def test_extension_values_are_validated_before_any_cuda_call(lib):
    """Round mode 2 (per-element SR) and dtypes 5..7 (signed) are accepted, anything beyond aborts -- checked with numel == 0
    calls, which return before touching CUDA, so this runs without a GPU."""
    buf = "b = ctypes.create_string_buffer(64)\n"
    ok = run_snippet(buf + "\n".join([
        "lib.piquant_quantize(ctypes.c_void_p(ctx), b, 0, b, 7, ctypes.c_size_t(0), ctypes.c_float(1.0), ctypes.c_int64(0), 2)",
        "lib.piquant_dequantize(ctypes.c_void_p(ctx), b, 6, b, 1, ctypes.c_size_t(0), ctypes.c_float(1.0), ctypes.c_int64(0), 1)",
        "lib.piquant_cuda_requantize(ctypes.c_void_p(ctx), b, 0, b, 5, ctypes.c_size_t(0), ctypes.c_float(1.0), ctypes.c_int64(0), 2, 0)",
        "lib.piquant_cuda_set_sr_key(ctypes.c_void_p(ctx), ctypes.c_uint64(5))",
        "lib.piquant_cuda_last_sr_key.restype = ctypes.c_uint64",
        "assert lib.piquant_cuda_last_sr_key(ctypes.c_void_p(ctx)) == 0",
    ]))
    assert ok.returncode == 0 and "survived" in ok.stdout, ok.stderr
    bad_mode = run_snippet(buf + "lib.piquant_quantize(ctypes.c_void_p(ctx), b, 0, b, 4, ctypes.c_size_t(8), ctypes.c_float(1.0), ctypes.c_int64(0), 3)")
    assert bad_mode.returncode == -6 and "invalid round mode" in bad_mode.stderr
    bad_dtype = run_snippet(buf + "lib.piquant_quantize(ctypes.c_void_p(ctx), b, 0, b, 8, ctypes.c_size_t(8), ctypes.c_float(1.0), ctypes.c_int64(0), 0)")
    assert bad_dtype.returncode == -6 and "must be a quantized type" in bad_dtype.stderr
    # signed parameter arithmetic on the host: the reference's formula with q_min = -2^(N-1)
    from oracle import port
    from piquant import Context, DataType
    assert Context.params_from_minmax(-1.0, 1.0, DataType.INT8) == port.params_from_minmax(-1.0, 1.0, port.INT8) == (np.float32(2 / 255), -1)


def test_round2_entry_points_validate_their_arguments_before_any_cuda_call(lib):
    """The fused / multi-source / flag entry points check dtype classes, source counts and pointers first (reference convention:
    message + abort, src/piquant.cpp:88-98), so a bad call dies with its own message even on a box without a GPU."""
    pre = ("b = ctypes.create_string_buffer(256)\n"
           "arr = (ctypes.c_void_p * 9)(*[ctypes.addressof(b)] * 9)\n"
           "V = ctypes.c_void_p\n")
    cases = [
        # 9 sources: one more than PIQUANT_CUDA_MAX_SUM_SOURCES
        ("lib.piquant_cuda_dequantize_sum_minmax_on_stream(V(ctx), arr, arr, ctypes.c_size_t(9), 4, b, 0, ctypes.c_size_t(64), 4, b, None, 0, None)",
         "between 1 and 8 sources"),
        ("lib.piquant_cuda_dequantize_sum_minmax_on_stream(V(ctx), arr, arr, ctypes.c_size_t(0), 4, b, 0, ctypes.c_size_t(64), 4, b, None, 0, None)",
         "between 1 and 8 sources"),
        # a float dtype where a quantized one belongs
        ("lib.piquant_cuda_dequantize_sum_minmax_on_stream(V(ctx), arr, arr, ctypes.c_size_t(2), 0, b, 0, ctypes.c_size_t(64), 4, b, None, 0, None)",
         "must be a quantized type"),
        ("lib.piquant_cuda_dequantize_add_minmax_on_stream(V(ctx), b, 4, b, 0, ctypes.c_size_t(64), b, 1, b, None, 0, None)",
         "is not a quantization type"),
        # the flag of piquant_cuda_wait_flag_on_stream is a 4-byte word
        ("lib.piquant_cuda_wait_flag_on_stream(V(ctx), V(ctypes.addressof(b) + 2), 0, None)", "4-byte aligned"),
        ("lib.piquant_cuda_wait_flag_on_stream(V(ctx), None, 0, None)", "4-byte aligned"),
        ("lib.piquant_cuda_copy_on_stream(V(ctx), None, b, ctypes.c_size_t(8), 0, None)", "must not be NULL"),
    ]
    for call, message in cases:
        r = run_snippet(pre + call)
        assert r.returncode == -6 and message in r.stderr, (call, r.returncode, r.stderr[-300:])
    # empty work returns before CUDA is touched
    ok = run_snippet(pre + "lib.piquant_cuda_copy_on_stream(V(ctx), b, b, ctypes.c_size_t(0), 0, None)")
    assert ok.returncode == 0 and "survived" in ok.stdout, ok.stderr


def test_compute_without_a_gpu_aborts_loudly(lib):
    import piquant

    if piquant.cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    buf = "b = ctypes.create_string_buffer(64)\n"
    r = run_snippet(buf + "lib.piquant_quantize(ctypes.c_void_p(ctx), b, 0, b, 4, ctypes.c_size_t(8), ctypes.c_float(1.0), ctypes.c_int64(0), 0)")
    assert r.returncode == -6 and "no usable CUDA device" in r.stderr and "survived" not in r.stdout


_FACADE_PROGRAM = r"""
#include <piquant.hpp>
#include <vector>
int main() {
    using namespace piquant;
    static_assert(dtype_traits<uint4_t>::type_code == dtype::uint4 && dtype_limits<uint4_t>::max == 15);
    static_assert(dtype_info_of(dtype::uint2).bit_size == 2 && sizeof(bfp16_t) == 2);
    static_assert(static_cast<float>(bfp16_t{1.0f} + bfp16_t{0.5f}) == 1.5f);
    static_assert(bfp16_t{3.14159f}.bits == 0x4049);
    context ctx{4};                               // no CUDA call
    ctx.set_stochastic_threshold(0.5f);
    std::vector<float> x; std::vector<uint4_t> q;
    ctx.quantize_generic<float, uint4_t>(x, q, 1.0f, 0, round_mode::nearest);      // numel == 0: returns before touching CUDA
    return ctx.native() != nullptr ? 0 : 1;
}
"""


def test_cxx_facade_compiles_and_links_against_the_library(lib, tmp_path):
    """include/piquant.hpp mirrors the reference's C++ API (reference include/piquant.hpp:20-339) header-only on the C ABI."""
    src = tmp_path / "facade.cpp"
    src.write_text(_FACADE_PROGRAM)
    exe = tmp_path / "facade"
    subprocess.run(["g++", "-std=c++20", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), f"-L{LIB.parent}", "-lpiquant",
                    f"-Wl,-rpath,{LIB.parent}", "-o", str(exe)], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_headers_are_valid_c99(tmp_path):
    """piquant.h is the reference's C99 API (reference include/piquant.h:1-3): both public headers must compile as C."""
    src = tmp_path / "c99.c"
    src.write_text('#include <piquant.h>\n#include <piquant_cuda.h>\n'
                   'int main(void) { piquant_cuda_meta_t m; m.error = 0; return (int)sizeof(m) == 64 ? m.error : 1; }\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", f"-I{ROOT / 'include'}", "-c", str(src), "-o", str(tmp_path / "c99.o")],
                   check=True, capture_output=True, text=True)
