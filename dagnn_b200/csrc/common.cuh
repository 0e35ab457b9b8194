// Shared helpers for libdagnn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "dagnn_b200.h"

namespace dagnn {

// thread-local error string behind dagnn_last_error()
char* err_buf();
int set_err(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(DAGNN_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return DAGNN_OK;
}

#define DAGNN_CUDA_OK(expr)                                                                   \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) return dagnn::set_err(DAGNN_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define DAGNN_REQUIRE(cond, msg)                                              \
  do {                                                                        \
    if (!(cond)) return dagnn::set_err(DAGNN_E_INVALID, "%s (%s)", msg, #cond); \
  } while (0)

// Function attributes (dynamic shared memory limit, cluster opt-in) are per device: run `f` once per device ordinal,
// thread-safe (the reference's DataParallel calls forward from one host thread per GPU, tg/data_parallel.py:60-61).
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  std::mutex mu;
  bool done[kMaxDevices] = {};
};
template <typename F>
inline int per_device_once(PerDeviceOnce& o, int* dev_out, F&& f) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_err(DAGNN_E_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= kMaxDevices) return set_err(DAGNN_E_UNSUPPORTED, "device ordinal %d", dev);
  if (dev_out) *dev_out = dev;
  std::lock_guard<std::mutex> lk(o.mu);
  if (o.done[dev]) return DAGNN_OK;
  const int rc = f(dev);
  if (rc == DAGNN_OK) o.done[dev] = true;
  return rc;
}

__host__ __device__ inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

}  // namespace dagnn
