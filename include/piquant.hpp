// piquant.hpp -- C++20 facade of the B200 pi-quant library, header-only on top of the C ABI.
//
// Source-compatible with the reference's C++ API (reference include/piquant.hpp:20-339): the same
// namespace, enums, element types (uint2_t, uint4_t, bfp16_t), dtype tables / traits / limits and the
// same `piquant::context` members, so that code written against the reference -- its gtest suites under
// reference test/*.cpp included -- compiles unchanged and links against libpiquant.so alone.  Where the
// reference's context owns a thread pool (reference src/piquant.cpp:113-211), this one owns a
// piquant_context_t* and every member is a thin inline forward to piquant.h / piquant_cuda.h:
// spans may point to host memory (synchronous, streamed through the GPU) or to CUDA device memory
// (asynchronous on the context's stream).
#pragma once

#include <array>
#include <bit>
#include <concepts>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <memory>
#include <span>
#include <string_view>
#include <type_traits>
#include <utility>

#include "piquant.h"
#include "piquant_cuda.h"

#define QUANT_EXPORT

namespace piquant {

// ---- enums: numeric values are the C ABI's (piquant.h), checked below ---------------------------------
enum class round_mode { nearest, stochastic, count_ };
enum class reduce_op { set, add, count_ };
// int2 / int4 / int8: signed extension of the B200 library (piquant_cuda.h), not in the reference at this commit
enum class dtype { f32 = 0, bf16, uint2, uint4, uint8, int2, int4, int8, count_ };

static_assert(static_cast<int>(round_mode::nearest) == PIQUANT_NEAREST && static_cast<int>(round_mode::stochastic) == PIQUANT_STOCHASTIC);
static_assert(static_cast<int>(reduce_op::set) == PIQUANT_REDUCE_OP_SET && static_cast<int>(reduce_op::add) == PIQUANT_REDUCE_OP_ADD);
static_assert(static_cast<int>(dtype::f32) == PIQUANT_DTYPE_F32 && static_cast<int>(dtype::bf16) == PIQUANT_DTYPE_BF16 &&
              static_cast<int>(dtype::uint2) == PIQUANT_DTYPE_UINT2 && static_cast<int>(dtype::uint4) == PIQUANT_DTYPE_UINT4 &&
              static_cast<int>(dtype::uint8) == PIQUANT_DTYPE_UINT8 && static_cast<int>(dtype::int2) == PIQUANT_CUDA_DTYPE_INT2 &&
              static_cast<int>(dtype::int4) == PIQUANT_CUDA_DTYPE_INT4 && static_cast<int>(dtype::int8) == PIQUANT_CUDA_DTYPE_INT8);

// ---- element types ------------------------------------------------------------------------------------

// One storage byte of a bit-packed tensor: holds 8/Bits elements, element k in bits [k*Bits, k*Bits+Bits).
template <unsigned Bits>
struct packed_uint final {
    using packed_storage = std::uint8_t;
    packed_storage bits {};

    constexpr packed_uint() noexcept = default;
    constexpr packed_uint(int v) noexcept : bits {static_cast<packed_storage>(v)} {}
    constexpr auto operator==(packed_uint rhs) const noexcept -> bool { return bits == rhs.bits; }
    constexpr auto operator==(packed_storage rhs) const noexcept -> bool { return bits == rhs; }
    constexpr explicit operator std::uint8_t() const noexcept { return bits; }
    constexpr explicit operator std::int64_t() const noexcept { return bits; }
};
using uint2_t = packed_uint<2>;
using uint4_t = packed_uint<4>;

using fp32_t = float;

// bfloat16 as raw bits.  f32 -> bf16 rounds to nearest even and keeps NaNs quiet; bf16 -> f32 is a shift.
// Arithmetic is done in f32 and rounded back (this is what the fused requantize kernel reproduces).
struct bfp16_t final {
    using packed_storage = std::uint16_t;
    packed_storage bits {};

    constexpr bfp16_t() noexcept = default;
    constexpr bfp16_t(fp32_t value) noexcept {
        const auto u {std::bit_cast<std::uint32_t>(value)};
        const bool is_nan {(u & 0x7fffffffu) > 0x7f800000u};
        bits = static_cast<packed_storage>(is_nan ? (u >> 16) | 64u : (u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    }
    constexpr explicit operator fp32_t() const noexcept { return std::bit_cast<fp32_t>(static_cast<std::uint32_t>(bits) << 16); }
    constexpr auto operator==(bfp16_t rhs) const noexcept -> bool { return bits == rhs.bits; }
    constexpr auto operator==(packed_storage rhs) const noexcept -> bool { return bits == rhs; }

#define PIQUANT_BF16_OP(op)                                                                                      \
    constexpr auto operator op(bfp16_t rhs) const noexcept -> bfp16_t { return {static_cast<fp32_t>(*this) op static_cast<fp32_t>(rhs)}; } \
    constexpr auto operator op##=(bfp16_t rhs) noexcept -> bfp16_t& { return *this = *this op rhs; }
    PIQUANT_BF16_OP(+)
    PIQUANT_BF16_OP(-)
    PIQUANT_BF16_OP(*)
    PIQUANT_BF16_OP(/)
#undef PIQUANT_BF16_OP
};

static_assert(sizeof(uint2_t) == 1 && sizeof(uint4_t) == 1 && sizeof(bfp16_t) == 2);

// ---- dtype table ----------------------------------------------------------------------------------------
struct dtype_flags final {
    enum $ { none = 0, is_quant = 1 << 0, is_float = 1 << 1, is_int = 1 << 2, is_signed = 1 << 3, is_packed = 1 << 4 };
};

struct dtype_info final {
    std::string_view name;
    std::size_t stride;       // bytes of one storage unit
    std::size_t bit_size;     // bits of one element
    std::underlying_type_t<dtype_flags::$> flags;
};

inline constexpr std::array<dtype_info, static_cast<std::size_t>(dtype::count_)> dtype_infos {{
    {"f32", 4, 32, dtype_flags::is_float | dtype_flags::is_signed},
    {"bf16", 2, 16, dtype_flags::is_float | dtype_flags::is_signed},
    {"uint2", 1, 2, dtype_flags::is_quant | dtype_flags::is_int | dtype_flags::is_packed},
    {"uint4", 1, 4, dtype_flags::is_quant | dtype_flags::is_int | dtype_flags::is_packed},
    {"uint8", 1, 8, dtype_flags::is_quant | dtype_flags::is_int},
    {"int2", 1, 2, dtype_flags::is_quant | dtype_flags::is_int | dtype_flags::is_packed | dtype_flags::is_signed},
    {"int4", 1, 4, dtype_flags::is_quant | dtype_flags::is_int | dtype_flags::is_packed | dtype_flags::is_signed},
    {"int8", 1, 8, dtype_flags::is_quant | dtype_flags::is_int | dtype_flags::is_signed},
}};
[[nodiscard]] constexpr auto dtype_info_of(dtype dt) noexcept -> const dtype_info& { return dtype_infos[static_cast<std::size_t>(dt)]; }

template <typename T> concept is_float_type = std::is_floating_point_v<T> || std::is_same_v<T, bfp16_t>;
template <typename T> concept is_quant_type = std::is_integral_v<T> || std::is_same_v<T, uint2_t> || std::is_same_v<T, uint4_t>;
template <typename T> concept is_dtype = is_float_type<T> || is_quant_type<T>;

template <typename T> requires is_dtype<T> struct dtype_traits final {};
template <> struct dtype_traits<fp32_t> { static constexpr dtype type_code {dtype::f32}; };
template <> struct dtype_traits<bfp16_t> { static constexpr dtype type_code {dtype::bf16}; };
template <> struct dtype_traits<uint2_t> { static constexpr dtype type_code {dtype::uint2}; };
template <> struct dtype_traits<uint4_t> { static constexpr dtype type_code {dtype::uint4}; };
template <> struct dtype_traits<std::uint8_t> { static constexpr dtype type_code {dtype::uint8}; };
template <> struct dtype_traits<std::int8_t> { static constexpr dtype type_code {dtype::int8}; };

template <typename> struct dtype_limits final {};
template <> struct dtype_limits<fp32_t> final {
    static constexpr fp32_t min {std::numeric_limits<fp32_t>::lowest()};
    static constexpr fp32_t max {std::numeric_limits<fp32_t>::max()};
};
template <> struct dtype_limits<bfp16_t> final {
    static constexpr bfp16_t min {std::bit_cast<fp32_t>(0xff7f0000u)};      // bits 0xFF7F: most negative finite bf16
    static constexpr bfp16_t max {std::bit_cast<fp32_t>(0x7f7f0000u)};      // bits 0x7F7F
};
template <> struct dtype_limits<uint2_t> final { static constexpr std::uint8_t min {0}, max {3}; };
template <> struct dtype_limits<uint4_t> final { static constexpr std::uint8_t min {0}, max {15}; };
template <> struct dtype_limits<std::uint8_t> final { static constexpr std::uint8_t min {0}, max {255}; };

// ---- context ------------------------------------------------------------------------------------------
class context final {
public:
    // num_threads is accepted for source compatibility; the GPU grid replaces the thread pool.
    explicit context(std::size_t num_threads) : m_ctx {piquant_context_create(num_threads), &piquant_context_destroy} {}
    context(const context&) = delete;
    context(context&&) = delete;
    auto operator=(const context&) -> context& = delete;
    auto operator=(context&&) -> context& = delete;
    ~context() = default;

    // out = quantize(in); out must hold exactly ceil(numel * bits / 8) bytes (reference src/piquant.cpp:277-308)
    auto quantize(std::span<const std::byte> in, dtype dtype_in, std::span<std::byte> out, dtype dtype_out, fp32_t scale,
                  std::int64_t zero_point, round_mode mode) const -> void {
        const std::size_t numel {in.size() / dtype_info_of(dtype_in).stride};
        expect(out.size() == storage_bytes(dtype_out, numel), "quantize: output span has the wrong size");
        piquant_quantize(m_ctx.get(), in.data(), static_cast<piquant_dtype_t>(dtype_in), out.data(), static_cast<piquant_dtype_t>(dtype_out),
                         numel, scale, zero_point, static_cast<piquant_round_mode_t>(mode));
    }

    // out (op)= dequantize(in); in must hold exactly ceil(numel * bits / 8) bytes (reference src/piquant.cpp:310-340)
    auto dequantize(std::span<const std::byte> in, dtype dtype_in, std::span<std::byte> out, dtype dtype_out, fp32_t scale,
                    std::int64_t zero_point, reduce_op op) const -> void {
        const std::size_t numel {out.size() / dtype_info_of(dtype_out).stride};
        expect(in.size() == storage_bytes(dtype_in, numel), "dequantize: input span has the wrong size");
        piquant_dequantize(m_ctx.get(), in.data(), static_cast<piquant_dtype_t>(dtype_in), out.data(), static_cast<piquant_dtype_t>(dtype_out),
                           numel, scale, zero_point, static_cast<piquant_reduce_op_t>(op));
    }

    // out (op)= dequantize(quantize(in)), unpacked, same float type in and out (reference src/piquant.cpp:342-369)
    auto quantize_dequantize_fused(std::span<const std::byte> in, dtype dtype_in_out, std::span<std::byte> out, dtype quant_type,
                                   fp32_t scale, std::int64_t zero_point, round_mode mode, reduce_op op) const -> void {
        expect(in.size() == out.size(), "quantize_dequantize_fused: input and output spans must have the same length");
        piquant_cuda_requantize(m_ctx.get(), in.data(), static_cast<piquant_dtype_t>(dtype_in_out), out.data(),
                                static_cast<piquant_dtype_t>(quant_type), in.size() / dtype_info_of(dtype_in_out).stride, scale, zero_point,
                                static_cast<piquant_round_mode_t>(mode), static_cast<piquant_reduce_op_t>(op));
    }

    [[nodiscard]] auto compute_quant_config_from_data(std::span<const fp32_t> x, dtype quant_dst_dtype) const -> std::pair<fp32_t, std::int64_t> {
        std::pair<fp32_t, std::int64_t> r {};
        piquant_compute_quant_params_float32(m_ctx.get(), x.data(), x.size(), static_cast<piquant_dtype_t>(quant_dst_dtype), &r.first, &r.second);
        return r;
    }
    [[nodiscard]] auto compute_quant_config_from_data(std::span<const bfp16_t> x, dtype quant_dst_dtype) const -> std::pair<fp32_t, std::int64_t> {
        std::pair<fp32_t, std::int64_t> r {};
        piquant_compute_quant_params_bfloat16(m_ctx.get(), reinterpret_cast<const std::uint16_t*>(x.data()), x.size(),
                                              static_cast<piquant_dtype_t>(quant_dst_dtype), &r.first, &r.second);
        return r;
    }

    // typed conveniences
    template <typename IN, typename OUT> requires is_float_type<IN> && is_quant_type<OUT>
    auto quantize_generic(std::span<const IN> in, std::span<OUT> out, fp32_t scale, std::int64_t zero_point, round_mode mode) -> void {
        quantize(std::as_bytes(in), dtype_traits<IN>::type_code, std::as_writable_bytes(out), dtype_traits<OUT>::type_code, scale, zero_point, mode);
    }
    template <typename IN, typename OUT> requires is_quant_type<IN> && is_float_type<OUT>
    auto dequantize_generic(std::span<const IN> in, std::span<OUT> out, fp32_t scale, std::int64_t zero_point, reduce_op op) -> void {
        dequantize(std::as_bytes(in), dtype_traits<IN>::type_code, std::as_writable_bytes(out), dtype_traits<OUT>::type_code, scale, zero_point, op);
    }
    template <typename INOUT, typename QUANT> requires is_float_type<INOUT> && is_quant_type<QUANT>
    auto quantize_dequantize_fused_generic(std::span<const INOUT> in, std::span<INOUT> out, fp32_t scale, std::int64_t zero_point,
                                           round_mode mode, reduce_op op) -> void {
        quantize_dequantize_fused(std::as_bytes(in), dtype_traits<INOUT>::type_code, std::as_writable_bytes(out), dtype_traits<QUANT>::type_code,
                                  scale, zero_point, mode, op);
    }

    // ---- CUDA side (no counterpart in the reference) ----
    [[nodiscard]] auto native() const noexcept -> piquant_context_t* { return m_ctx.get(); }
    auto set_stream(void* cuda_stream) const -> void { piquant_cuda_set_stream(m_ctx.get(), cuda_stream); }
    auto synchronize() const -> void { piquant_cuda_synchronize(m_ctx.get()); }
    auto set_stochastic_threshold(fp32_t xi) const -> void { piquant_cuda_set_stochastic_threshold(m_ctx.get(), xi); }

    // Kept for source compatibility with code that names these types (reference include/piquant.hpp:313-335);
    // the CUDA dispatcher does not use them.
    class pimpl;
    enum class command_type { quant, dequant, quant_dequant };
    struct quant_descriptor final {
        command_type type {command_type::quant};
        const std::byte* in {};
        std::byte* out {};
        std::int64_t numel {};
        fp32_t scale {};
        std::int64_t zero_point {};
        dtype dt_in {};
        dtype dt_out {};
        round_mode rounding {};
        reduce_op reducing {};
        fp32_t rnd_threshold {};
    };

private:
    [[nodiscard]] static constexpr auto storage_bytes(dtype dt, std::size_t numel) noexcept -> std::size_t {
        const auto& info {dtype_info_of(dt)};
        if (info.bit_size >= 8) return numel * info.stride;
        const std::size_t per_byte {8u / info.bit_size};
        return (numel + per_byte - 1) / per_byte * info.stride;
    }
    static auto expect(bool ok, const char* what) -> void {
        if (!ok) [[unlikely]] {     // the library's error convention: message + abort
            std::fprintf(stderr, "\x1b[31mpiquant: %s\x1b[0m\n", what);
            std::abort();
        }
    }

    std::unique_ptr<piquant_context_t, void (*)(piquant_context_t*)> m_ctx;
};

}  // namespace piquant
