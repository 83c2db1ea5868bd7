// Minimal DECLARATION STUB of the xla::ffi surface that bijax_b200/csrc/xla_ffi_shim.cc uses -- test infrastructure only.
//
// jaxlib's real header (jax.ffi.include_dir()/xla/ffi/api/ffi.h) is not available in this image.  This file is NOT a copy of
// it: it declares, from the documented public API, just the names the shim touches (Buffer<dtype>, ResultBuffer, Error,
// ErrorCode, PlatformStream, Ffi::Bind().Ctx/Arg/Ret/Attr, XLA_FFI_DEFINE_HANDLER_SYMBOL) so that
// tests/test_ffi_shim_cpu.py can type-check the shim with `g++ -fsyntax-only`.  Its binder records the C++ type every
// Ctx / Arg / Ret / Attr contributes and the DEFINE macro static_asserts that the handler is callable with exactly that
// list and returns Error -- the same contract the real binder enforces.  Nothing here can execute a custom call.
#ifndef BJX_TEST_XLA_FFI_STUB_H_
#define BJX_TEST_XLA_FFI_STUB_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

enum class DataType { U8, S32, U32, S64, F32 };
inline constexpr DataType U8 = DataType::U8;
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType U32 = DataType::U32;
inline constexpr DataType S64 = DataType::S64;
inline constexpr DataType F32 = DataType::F32;

namespace internal {
template <DataType>
struct Native;
template <>
struct Native<DataType::U8> { using type = uint8_t; };
template <>
struct Native<DataType::S32> { using type = int32_t; };
template <>
struct Native<DataType::U32> { using type = uint32_t; };
template <>
struct Native<DataType::S64> { using type = int64_t; };
template <>
struct Native<DataType::F32> { using type = float; };
}  // namespace internal

template <typename T>
class Span {
 public:
  size_t size() const;
  const T& operator[](size_t i) const;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename internal::Native<dtype>::type;
  T* typed_data() const;
  size_t element_count() const;
  Span<const int64_t> dimensions() const;
};

template <typename T>
class Result {
 public:
  T* operator->() const;
  T& operator*() const;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented, kResourceExhausted };

class Error {
 public:
  Error(ErrorCode code, std::string message);
  static Error Success();
};

template <typename T>
struct PlatformStream {};

namespace internal {
template <typename T>
struct CtxType;
template <typename T>
struct CtxType<PlatformStream<T>> { using type = T; };
}  // namespace internal

template <typename... Ts>
struct Binding {
  template <typename T>
  Binding<Ts..., typename internal::CtxType<T>::type> Ctx() const;
  template <typename T>
  Binding<Ts..., T> Arg() const;
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const;
  template <typename T>
  Binding<Ts..., T> Attr(const char* name) const;
  template <typename F>
  static constexpr bool Accepts = std::is_invocable_r_v<Error, F, Ts...>;
  static constexpr size_t kArity = sizeof...(Ts);
};

struct Ffi {
  static Binding<> Bind();
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                        \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>,                                              \
                #name ": the handler's parameter list does not match its Ffi::Bind() chain (or it does not return Error)"); \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame* call_frame)

#endif  // BJX_TEST_XLA_FFI_STUB_H_
