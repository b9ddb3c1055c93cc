// dev_util.cuh — small RAII helpers shared by the C ABI translation unit and the device paths.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

namespace sdg {

#define CUDA_OK(x)                                                                                              \
  do {                                                                                                          \
    cudaError_t e_ = (x);                                                                                       \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_));           \
  } while (0)

// cudaFuncSetAttribute is per DEVICE: remember, per kernel instantiation, on which devices the attribute has been set (one process
// may hold contexts on several devices, sdg_config.device).  Setting it twice from two threads is harmless.
inline bool firstUseOnThisDevice(std::atomic<unsigned long long>& mask) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (mask.load(std::memory_order_acquire) & bit) return false;
  mask.fetch_or(bit, std::memory_order_acq_rel);
  return true;
}

void setLastError(const char* msg);   // sdg_api.cu: message returned by sdg_last_error()

template <typename T>
struct DevBuf {
  T* p = nullptr; size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void alloc(size_t count) { release(); n = count; if (count) CUDA_OK(cudaMalloc(&p, count * sizeof(T))); }
  void upload(const std::vector<T>& v, cudaStream_t s = 0) { alloc(v.size()); if (!v.empty()) { CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s)); CUDA_OK(cudaStreamSynchronize(s)); } }
  void zero(cudaStream_t s = 0) { if (n) CUDA_OK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { release(); }
};

}  // namespace sdg
