// Syntax-check stand-in for XLA's FFI header (xla/ffi/api/ffi.h, shipped by jaxlib under jax.ffi.include_dir()).
// TEST INFRASTRUCTURE: just enough declarations for `g++ -fsyntax-only` of marl_sat_b200/csrc/xla_ffi_shim.cc in
// an image without JAX (tests/test_host_abi.py).  It checks that the shim parses and that every handler's
// argument list type-checks against the msat_* prototypes, not XLA's binding machinery.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace xla::ffi {
enum DataType { U8, S8, S32, U32, F32, F64, S64, PRED };
template <DataType> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S8> { using type = int8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<PRED> { using type = bool; };
struct Dimensions {
    const int64_t* ptr; size_t n;
    size_t size() const { return n; }
    int64_t operator[](size_t i) const { return ptr[i]; }
    int64_t back() const { return ptr[n - 1]; }
};
template <DataType T> struct Buffer {
    typename NativeOf<T>::type* typed_data() const { return nullptr; }
    Dimensions dimensions() const { return {nullptr, 0}; }
    size_t element_count() const { return 0; }
};
template <DataType T> struct ResultHolder {
    Buffer<T> b;
    Buffer<T>* operator->() { return &b; }
};
template <DataType T> using ResultBuffer = ResultHolder<T>;
enum class ErrorCode { kInvalidArgument, kInternal };
struct Error {
    static Error Success() { return {}; }
    Error() = default;
    Error(ErrorCode, std::string) {}
};
template <typename S> struct PlatformStream {};
struct Binding {
    template <typename T> Binding& Ctx() { return *this; }
    template <typename T> Binding& Arg() { return *this; }
    template <typename T> Binding& Ret() { return *this; }
    template <typename T> Binding& Attr(const char*) { return *this; }
};
struct Ffi { static Binding Bind() { return {}; } };
}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
    extern "C" void* name() { (void)(binding); return reinterpret_cast<void*>(&impl); }
