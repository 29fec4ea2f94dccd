// TEST STAND-IN for xla/ffi/api/ffi.h (jaxlib is not installable in this image).
//
// Provides just enough of the XLA FFI C++ API for multitreegp_b200/csrc/mtgp_xla_ffi.cc to COMPILE and for
// its handler bodies (RolloutImpl, PointwiseImpl) to be EXECUTED by tests/test_ffi_adapter.py with device
// buffers shaped as `jax.ffi.ffi_call(..., vmap_method="expand_dims")` delivers them: Buffer<dtype> (dimensions,
// typed_data, element_count), Span, Error, ScratchAllocator, ResultBuffer, and a binding DSL whose calls are
// no-ops.  The call-frame decoding of the real header is NOT reproduced: XLA_FFI_DEFINE_HANDLER_SYMBOL defines a
// symbol that reports "stand-in".  With the real header (jax.ffi.include_dir()) the same source builds the real
// handlers.  Not part of the product.
#pragma once
#define MTGP_FAKE_XLA_FFI 1

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

namespace xla::ffi {

enum DataType { F32 = 11, U32 = 8, S32 = 4 };
template <DataType dt> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<U32> { using type = uint32_t; };
template <> struct NativeType<S32> { using type = int32_t; };

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  const T* data() const { return d_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T& back() const { return d_[n_ - 1]; }
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }

 private:
  const T* d_ = nullptr;
  size_t n_ = 0;
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  Buffer() = default;
  Buffer(T* data, std::vector<int64_t> dims) : data_(data), dims_(std::move(dims)) {}
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
  T* typed_data() const { return data_; }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims_) n *= (size_t)d;
    return n;
  }

 private:
  T* data_ = nullptr;
  std::vector<int64_t> dims_;
};

template <DataType dt>
class ResultBuffer {
 public:
  explicit ResultBuffer(Buffer<dt>* b) : b_(b) {}
  Buffer<dt>* operator->() const { return b_; }

 private:
  Buffer<dt>* b_;
};

class Error {
 public:
  static Error Success() { return Error(0, ""); }
  static Error InvalidArgument(std::string m) { return Error(3, std::move(m)); }
  static Error Internal(std::string m) { return Error(13, std::move(m)); }
  bool failure() const { return code_ != 0; }
  int code() const { return code_; }
  const std::string& message() const { return msg_; }

 private:
  Error(int c, std::string m) : code_(c), msg_(std::move(m)) {}
  int code_;
  std::string msg_;
};

// device scratch that lives as long as the call (XLA frees it after the handler returns; so does this one)
class ScratchAllocator {
 public:
  ScratchAllocator() = default;
  ScratchAllocator(const ScratchAllocator&) = delete;
  ScratchAllocator(ScratchAllocator&& o) noexcept : ptrs_(std::move(o.ptrs_)) { o.ptrs_.clear(); }
  ~ScratchAllocator() {
    for (void* p : ptrs_) cudaFree(p);
  }
  std::optional<void*> Allocate(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) return std::nullopt;
    ptrs_.push_back(p);
    return p;
  }

 private:
  std::vector<void*> ptrs_;
};

template <typename T> struct PlatformStream {};

// binding DSL: every call is a no-op that returns the binder
struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace xla::ffi

// the real macro defines  extern "C" XLA_FFI_Error* sym(XLA_FFI_CallFrame*)  that decodes the call frame
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                      \
  extern "C" __attribute__((visibility("default"))) void* sym(void* /*frame*/) { \
    (void)(binding);                                                           \
    (void)&impl;                                                               \
    return (void*)"stand-in for xla/ffi/api/ffi.h: handler not callable through a call frame"; \
  }
