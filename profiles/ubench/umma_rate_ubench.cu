// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M = 128) as a function of N and kind, operands resident in shared
// memory (SWIZZLE_128B K-major tiles, contents irrelevant), 512 back-to-back MMAs into one TMEM accumulator, one CTA per SM
// on all SMs (so that clocks are at load).  Prints clocks per MMA and the fraction of the kind's peak MAC rate that this is.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate_ubench.bin umma_rate_ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int KIND>   // 0: tf32, 1: f16 (bf16 operands)
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128) rate(int n, int same_a, long long* out, int row_shift = 0) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 * 4 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t b0 = base + 4 * 16384;
    const long long t0 = clock64();
    for (int i = 0; i < 512; ++i) {
      const uint32_t a0 = base + (same_a ? 0u : (uint32_t)(i & 3) * 16384u) + (uint32_t)row_shift * 128u;   // row-shifted view (conv taps)   // 4 different A tiles in rotation, or always the same one
      umma<KIND>(tmem, make_desc_sw128(a0) + (uint64_t)(2 * (i & 3)), make_desc_sw128(b0) + (uint64_t)(2 * (i & 3)), idesc, i != 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 4 * 16384 + 32768 + 1024;
  cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int kind = 0; kind < 2; ++kind)
    for (int same_a = 0; same_a < 2; ++same_a)
      for (int n : {32, 64, 128, 256}) {
        long long best = 1LL << 60;
        for (int rep = 0; rep < 5; ++rep) {
          if (kind == 0) rate<0><<<sms, 128, smem>>>(n, same_a, d); else rate<1><<<sms, 128, smem>>>(n, same_a, d);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
          long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          if (h < best) best = h;
        }
        const double clk = best / 512.0;
        const double macs = 128.0 * n * (kind == 0 ? 8 : 16), peak = kind == 0 ? 2048.0 : 4096.0;
        printf("kind::%s  M=128 N=%3d K=%2d  A tiles %s : %6.1f clk per MMA  -> %5.1f %% of %4.0f MAC/clk/SM\n", kind == 0 ? "tf32" : "f16 ", n,
               kind == 0 ? 8 : 16, same_a ? "same   " : "rotated", clk, 100.0 * macs / clk / peak, peak);
      }
  // row-shifted A descriptors (start address not a multiple of the 1024-byte swizzle atom): same rate?
  for (int shift : {0, 1, 3, 8, 9}) {
    long long best = 1LL << 60;
    for (int rep = 0; rep < 5; ++rep) {
      rate<0><<<sms, 128, smem>>>(128, 1, d, shift);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (h < best) best = h;
    }
    printf("kind::tf32  M=128 N=128 K= 8  A start shifted by %d rows : %6.1f clk per MMA\n", shift, best / 512.0);
  }
  return 0;
}
