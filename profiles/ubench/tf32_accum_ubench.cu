// Micro-benchmark: how accurate is the fp32 accumulation of tcgen05.mma kind::tf32 on sm_100a?
// D[128 x 64] = A[128 x K] * B[64 x K]^T from one CTA, operands in the no-swizzle K-major panel layout.
//   mode 0: operands pre-rounded to TF32 (every product exact in fp32), one TMEM accumulator chain over all of K
//           -> the error against the float64 sum is the accumulator's alone
//   mode 1: full fp32 operands, 3-term hi/lo split (a_lo w_hi + a_hi w_lo + a_hi w_hi), one chain
//   mode 2: as mode 1, but the chain is cut every `flush` K-steps: TMEM -> registers, summed there in fp32 (round to nearest)
// Inputs uniform in [0,1) (sign-coherent sums expose a truncation bias as a negative mean error) or in [-0.5,0.5).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_accum_ubench tf32_accum_ubench.cu
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int M = 128, N = 64, KC = 16;   // K batch staged in shared memory per round

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(128 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da),
               "l"(db), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ float rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
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

__global__ void __launch_bounds__(128) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K,
                                                   int mode, int flush) {
  __shared__ __align__(128) float sA[2][KC / 4][M][4];   // [hi|lo][panel][row][4 k]
  __shared__ __align__(128) float sB[2][KC / 4][N][4];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  float sum[N];
#pragma unroll
  for (int j = 0; j < N; ++j) sum[j] = 0.f;
  uint32_t phase = 0;
  int steps_in_chain = 0;
  auto drain = [&]() {   // accumulator -> registers (fp32 round-to-nearest adds)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) sum[c + j] += __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };
  for (int k0 = 0; k0 < K; k0 += KC) {
    for (int e = tid; e < M * KC; e += 128) {
      const int r = e / KC, k = e % KC;
      const float v = A[(size_t)r * K + k0 + k];
      const float hi = (mode == 0) ? v : rna(v);
      sA[0][k / 4][r][k % 4] = hi;
      sA[1][k / 4][r][k % 4] = rna(v - hi);
    }
    for (int e = tid; e < N * KC; e += 128) {
      const int r = e / KC, k = e % KC;
      const float v = B[(size_t)r * K + k0 + k];
      const float hi = (mode == 0) ? v : rna(v);
      sB[0][k / 4][r][k % 4] = hi;
      sB[1][k / 4][r][k % 4] = rna(v - hi);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kk = 0; kk < KC / 8; ++kk) {
        const uint32_t acc0 = (steps_in_chain + kk) != 0 ? 1u : 0u;
        const uint64_t ahi = make_desc_ns(smem_u32(&sA[0][2 * kk][0][0]), M * 16), alo = make_desc_ns(smem_u32(&sA[1][2 * kk][0][0]), M * 16);
        const uint64_t bhi = make_desc_ns(smem_u32(&sB[0][2 * kk][0][0]), N * 16), blo = make_desc_ns(smem_u32(&sB[1][2 * kk][0][0]), N * 16);
        if (mode == 0) {
          umma_tf32(tmem, ahi, bhi, idesc, acc0);
        } else {
          umma_tf32(tmem, alo, bhi, idesc, acc0);
          umma_tf32(tmem, ahi, blo, idesc, 1u);
          umma_tf32(tmem, ahi, bhi, idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    steps_in_chain += KC / 8;
    // everyone waits for the MMAs of this round (shared memory is reused next round)
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra Dn;\nbra W;\nDn:\n}\n" ::"r"(smem_u32(&bar)), "r"(phase)
                 : "memory");
    phase ^= 1u;
    if (mode == 2 && steps_in_chain >= flush) { drain(); steps_in_chain = 0; }
    __syncthreads();
  }
  if (steps_in_chain > 0) drain();
  for (int j = 0; j < N; ++j) D[(size_t)tid * N + j] = sum[j];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static float host_rna_tf32(float v) {   // round to nearest (ties away) at 10 mantissa bits
  uint32_t u;
  memcpy(&u, &v, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

int main() {
  const int Ks[3] = {256, 1024, 4096};
  for (int sym = 0; sym < 2; ++sym)
    for (int K : Ks) {
      std::vector<float> A((size_t)M * K), B((size_t)N * K);
      srand(1234 + K);
      for (auto& v : A) v = (float)rand() / RAND_MAX - (sym ? 0.5f : 0.f);
      for (auto& v : B) v = (float)rand() / RAND_MAX - (sym ? 0.5f : 0.f);
      float *dA, *dB, *dD;
      cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, (size_t)M * N * 4);
      struct Cfg { int mode, flush; const char* name; } cfgs[] = {{0, 0, "tf32 exact products, one chain"}, {1, 0, "3xTF32, one chain"},
                                                                 {2, 16, "3xTF32, chain cut every 16 K-steps"}, {2, 4, "3xTF32, chain cut every 4 K-steps"}};
      for (const Cfg& c : cfgs) {
        std::vector<float> Ah = A, Bh = B;
        if (c.mode == 0) { for (auto& v : Ah) v = host_rna_tf32(v); for (auto& v : Bh) v = host_rna_tf32(v); }
        cudaMemcpy(dA, Ah.data(), Ah.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bh.data(), Bh.size() * 4, cudaMemcpyHostToDevice);
        gemm_kernel<<<1, 128>>>(dA, dB, dD, K, c.mode, c.flush);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<float> D((size_t)M * N);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double num = 0, den = 0, bias = 0, num32 = 0, bias32 = 0;
        for (int i = 0; i < M; ++i)
          for (int j = 0; j < N; ++j) {
            double ref = 0;
            float s32 = 0.f;
            for (int k = 0; k < K; ++k) { ref += (double)Ah[(size_t)i * K + k] * Bh[(size_t)j * K + k]; s32 = fmaf(Ah[(size_t)i * K + k], Bh[(size_t)j * K + k], s32); }
            const double e = D[(size_t)i * N + j] - ref, e32 = s32 - ref;
            num += e * e; den += ref * ref; bias += e; num32 += e32 * e32; bias32 += e32;
          }
        const double rms_ref = sqrt(den / (M * N));
        printf("inputs %-12s K=%4d  %-38s rel_l2 %.3e  mean_err/rms %+.3e   | sequential fp32 FMA: rel_l2 %.3e  mean_err/rms %+.3e\n",
               sym ? "[-0.5,0.5)" : "[0,1)", K, c.name, sqrt(num / den), bias / (M * N) / rms_ref, sqrt(num32 / den), bias32 / (M * N) / rms_ref);
      }
      cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
  return 0;
}
