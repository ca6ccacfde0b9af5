// MOCK of the slice of jaxlib's xla/ffi/api/ffi.h that csrc/xla/xla_ffi.cc uses.  TEST INFRASTRUCTURE: it lets the
// build image (no JAX) check that the adapter translation unit parses and type-checks; it is NOT the real API and is
// never shipped.  The real header comes from jax.ffi.include_dir().
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

namespace xla { namespace ffi {

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
enum class DataType { F32, F64, S32, U8 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T> class ErrorOr {
 public:
  ErrorOr() = default;
  ErrorOr(T v) : v_(std::move(v)) {}
  bool has_value() const { return v_.has_value(); }
  T& value() { return *v_; }
 private:
  std::optional<T> v_;
};

template <typename T> struct Span {
  const T* p = nullptr; size_t n = 0;
  const T* begin() const { return p; }
  const T* end() const { return p + n; }
};

class AnyBuffer {
 public:
  DataType element_type() const { return dt_; }
  Span<int64_t> dimensions() const { return Span<int64_t>{dims_.data(), dims_.size()}; }
  void* untyped_data() const { return data_; }
  size_t element_count() const { size_t n = 1; for (auto d : dims_) n *= (size_t)d; return n; }
 private:
  DataType dt_ = DataType::F32;
  std::vector<int64_t> dims_;
  void* data_ = nullptr;
};

template <typename T> class Result {
 public:
  T* operator->() { return &v_; }
 private:
  T v_;
};

class RemainingArgs {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<T> get(size_t) const { return ErrorOr<T>(); }
};
class RemainingRets {
 public:
  size_t size() const { return 0; }
  template <typename T> ErrorOr<Result<T>> get(size_t) const { return ErrorOr<Result<T>>(); }
};
class Dictionary {
 public:
  template <typename T> ErrorOr<T> get(const char*) const { return ErrorOr<T>(); }
};
class ScratchAllocator {
 public:
  std::optional<void*> Allocate(size_t) { return std::nullopt; }
};
template <typename S> struct PlatformStream {};

class Ffi {
 public:
  struct Binding {
    template <typename T> Binding& Ctx() { return *this; }
    Binding& RemainingArgs() { return *this; }
    Binding& RemainingRets() { return *this; }
    Binding& Attrs() { return *this; }
  };
  static Binding Bind() { return Binding(); }
};

}}  // namespace xla::ffi

// the real macro defines an extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*); the mock only instantiates the impl
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                 \
  extern "C" void* symbol(void*) {                                           \
    auto b = (binding); (void)b;                                             \
    auto fn = &impl; (void)fn;                                               \
    return nullptr;                                                          \
  }
