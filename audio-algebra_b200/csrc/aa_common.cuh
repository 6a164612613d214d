// Shared host/device helpers for libaa_b200 (error reporting, launch accounting, small device utils).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include "../../include/aa_b200.h"

namespace aa {

void set_error(const char* fmt, ...);          // thread-local message, defined in api.cu
extern std::atomic<int64_t> g_launch_count;    // defined in api.cu

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define AA_REQUIRE(cond, ...)                         \
  do {                                                \
    if (!(cond)) {                                    \
      ::aa::set_error(__VA_ARGS__);                   \
      return AA_ERR_INVALID_ARG;                      \
    }                                                 \
  } while (0)

#define AA_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::aa::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return AA_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define AA_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::aa::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return AA_ERR_CUDA;                                                                  \
    }                                                                                      \
    ::aa::count_launch();                                                                  \
  } while (0)

inline int num_sms() {   // SM count of the CURRENT device (cached per device: one process may drive several GPUs)
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: opt in once per (kernel, device).
template <typename K>
inline cudaError_t ensure_dyn_smem(K kernel, int bytes) {
  struct Slot { const void* k; int bytes[64]; };
  static Slot slots[96] = {};
  static int n_slots = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  const void* key = reinterpret_cast<const void*>(kernel);
  std::lock_guard<std::mutex> g(mu);
  Slot* sl = nullptr;
  for (int i = 0; i < n_slots; ++i)
    if (slots[i].k == key) { sl = &slots[i]; break; }
  if (sl && sl->bytes[dev] >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (!sl && n_slots < 96) { sl = &slots[n_slots++]; sl->k = key; }
  if (sl) sl->bytes[dev] = bytes;
  return e;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// streaming (read-once) 128-bit global load that does not allocate in L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

}  // namespace aa
