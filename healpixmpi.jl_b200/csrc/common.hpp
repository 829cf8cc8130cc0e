// common.hpp — error plumbing shared by all translation units of libhpsht_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/hpsht_b200.h"

namespace hpsht {

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);

#define HP_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      throw ::hpsht::Error(HPSHT_ERR_CUDA, std::string(#call) + ": " +                     \
                                               cudaGetErrorString(e__) + " (" + __FILE__ + \
                                               ":" + std::to_string(__LINE__) + ")");      \
    }                                                                                      \
  } while (0)

#define HP_REQUIRE(cond, code, msg)                       \
  do {                                                    \
    if (!(cond)) throw ::hpsht::Error((code), (msg));     \
  } while (0)

// Translate C++ exceptions into status codes at the C boundary.
template <class F>
static inline int guarded(F&& f) {
  try {
    f();
    return HPSHT_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return HPSHT_ERR_INTERNAL;
  } catch (...) {
    set_last_error("unknown exception");
    return HPSHT_ERR_INTERNAL;
  }
}

// Simple RAII device buffer.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    release();
    if (count == 0) return;
    HP_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
  }
  void ensure(size_t count) {
    if (count > n) alloc(count);
  }
  void upload(const T* h, size_t count, cudaStream_t s = 0) {
    ensure(count);
    if (count) HP_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s = 0) {
    if (n) HP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

static inline bool is_device_pointer(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace hpsht
