// Micro-benchmark / probe: can a SWIZZLE_128B K-major tcgen05 operand start at an arbitrary ROW of a tile that TMA (or, here, a
// hand-written copy of TMA's layout) placed at a 1024-byte aligned address?  If yes, the 7 taps of a dilated k7 convolution can
// read one activation tile (+ halo) through 7 row-shifted descriptors instead of 7 separately loaded tiles.
//   layout: element (row r, k) of a [rows x 32 fp32] tile at  r*128 + (((k/4) ^ (r & 7)) * 16) + (k%4)*4   (r & 7 = address bits 7..9)
//   probe:  D[128 x 64] = A[r0 .. r0+127][0..31] * B[64 x 32]^T  for r0 in {0,1,3,8,9,27}, descriptor base_offset = 0 or (r0 & 7)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sw128_rowshift_ubench.bin sw128_rowshift_ubench.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int RA = 160, N = 64, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int r0, int use_bo) {
  extern __shared__ unsigned char raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  unsigned char* p = raw + (base - smem_u32(raw));
  float* sA = reinterpret_cast<float*>(p);                  // RA rows x 128 B
  float* sB = reinterpret_cast<float*>(p + RA * 128);       // N rows x 128 B (RA * 128 is a multiple of 1024)
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < RA * K; e += 128) { const int r = e / K, k = e % K; sA[r * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3)] = A[e]; }
  for (int e = tid; e < N * K; e += 128) { const int r = e / K, k = e % K; sB[r * 32 + (((k >> 2) ^ (r & 7)) << 2) + (k & 3)] = B[e]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = base + (uint32_t)r0 * 128u, b0 = base + RA * 128;
    for (int k = 0; k < 4; ++k)
      umma_tf32(tmem, make_desc_sw128(a0, use_bo ? (uint32_t)r0 : 0u) + (uint64_t)(2 * k), make_desc_sw128(b0, 0) + (uint64_t)(2 * k), idesc, k != 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    for (int j = 0; j < 32; ++j) D[tid * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> A(RA * K), B(N * K), D(128 * N);
  for (int r = 0; r < RA; ++r) for (int k = 0; k < K; ++k) A[r * K + k] = (float)((r * 7 + k * 3) % 11 - 5);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = (float)((n * 5 + k) % 7 - 3);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = RA * 128 + N * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[6] = {0, 1, 3, 8, 9, 27};
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int r0 : shifts) {
      probe<<<1, 128, smem>>>(dA, dB, dD, r0, use_bo);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          float ref = 0.f;
          for (int k = 0; k < K; ++k) ref += A[(m + r0) * K + k] * B[n * K + k];
          if (ref != D[m * N + n]) ++bad;
        }
      printf("row shift %2d  base_offset field %s : %s (%d of %d wrong)\n", r0, use_bo ? "(r0 & 7)" : "0       ", bad ? "WRONG" : "exact", bad, 128 * N);
    }
  return 0;
}
