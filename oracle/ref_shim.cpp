// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's fused quantize->dequantize entry point exists only in its C++ API
// (include/piquant.hpp:276-285, not in piquant.h).  This shim is compiled INTO
// oracle/_ref/libpiquant_ref.so against the reference's own header (found under $(REF) at build
// time, never copied) and exposes that one method with C linkage so ctypes can reach it.
#include <piquant.hpp>

#include <bit>
#include <cstddef>
#include <cstdint>
#include <span>

extern "C" __attribute__((visibility("default")))
void piquant_ref_shim_requantize(void* ctx, const void* in, int dtype_in_out, void* out, int quant_type,
                                 std::size_t numel, float scale, std::int64_t zero_point, int mode, int op) {
    using namespace piquant;
    const std::size_t bytes = numel * dtype_info_of(static_cast<dtype>(dtype_in_out)).stride;
    std::bit_cast<context*>(ctx)->quantize_dequantize_fused(
        std::span<const std::byte>{static_cast<const std::byte*>(in), bytes},
        static_cast<dtype>(dtype_in_out),
        std::span<std::byte>{static_cast<std::byte*>(out), bytes},
        static_cast<dtype>(quant_type),
        scale, zero_point,
        static_cast<round_mode>(mode),
        static_cast<reduce_op>(op));
}
