// Microbenchmark: scalar FFMA vs packed FFMA2 issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int ILP>
__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b, int iters) {
  float v[ILP];
  #pragma unroll
  for (int i=0;i<ILP;i++) v[i] = threadIdx.x*0.001f + i;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) v[i] = fmaf(v[i], a, b);
  }
  float s=0; 
  #pragma unroll
  for (int i=0;i<ILP;i++) s+=v[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP>
__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b, int iters) {
  float2 v[ILP];
  float2 aa = make_float2(a,a), bb = make_float2(b,b);
  #pragma unroll
  for (int i=0;i<ILP;i++) v[i] = make_float2(threadIdx.x*0.001f + i, i*0.5f);
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) v[i] = __ffma2_rn(v[i], aa, bb);
  }
  float s=0; 
  #pragma unroll
  for (int i=0;i<ILP;i++) s+=v[i].x+v[i].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// mixed: fadd2 + ffma2 + LDS.128 like an FFT
template<int ILP>
__global__ void __launch_bounds__(256) k_fadd2(float* out, float a, float b, int iters) {
  float2 v[ILP];
  float2 bb = make_float2(b,a);
  #pragma unroll
  for (int i=0;i<ILP;i++) v[i] = make_float2(threadIdx.x*0.001f + i, i*0.5f);
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int i=0;i<ILP;i++) v[i] = __fadd2_rn(v[i], bb);
  }
  float s=0; 
  #pragma unroll
  for (int i=0;i<ILP;i++) s+=v[i].x+v[i].y;
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k_lds128(float* out, int iters) {
  extern __shared__ float4 sm[];
  for (int i=threadIdx.x;i<2048;i+=blockDim.x) sm[i]=make_float4(i,i,i,i);
  __syncthreads();
  float4 acc=make_float4(0,0,0,0);
  int idx=threadIdx.x;
  for (int it=0; it<iters; ++it) {
    #pragma unroll
    for (int j=0;j<8;j++){ float4 t=sm[(idx+j*256)&2047]; acc.x+=t.x; acc.y+=t.y; acc.z+=t.z; acc.w+=t.w; }
    idx=(idx+32)&2047;
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc.x+acc.y+acc.z+acc.w;
}
int main(){
  float* d; CK(cudaMalloc(&d, 148*8*256*4*4));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=20000; float ms;
  for (int rep=0; rep<2; ++rep) {
  for (int blocksPerSM : {1,2,4}) {
    int grid=148*blocksPerSM;
    k_ffma<16><<<grid,256>>>(d,1.0001f,0.5f,iters); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_ffma<16><<<grid,256>>>(d,1.0001f,0.5f,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    double fma = (double)grid*256*16*iters; 
    printf("FFMA   blocks/SM=%d  %.3f ms  %.2f TFMA/s  (%.1f lanes/SM/clk @1.965GHz)\n", blocksPerSM, ms, fma/ms/1e9, fma/ms/1e-3/148/1.965e9);
    cudaEventRecord(e0); k_ffma2<16><<<grid,256>>>(d,1.0001f,0.5f,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    printf("FFMA2  blocks/SM=%d  %.3f ms  %.2f TFMA/s  (%.1f lanes/SM/clk)\n", blocksPerSM, ms, 2*fma/ms/1e9, 2*fma/ms/1e-3/148/1.965e9);
    cudaEventRecord(e0); k_fadd2<16><<<grid,256>>>(d,1.0001f,0.5f,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    printf("FADD2  blocks/SM=%d  %.3f ms  %.2f Tadd/s  (%.1f lanes/SM/clk)\n", blocksPerSM, ms, 2*fma/ms/1e9, 2*fma/ms/1e-3/148/1.965e9);
  }}
  {
    int grid=148*4; int it2=4000;
    k_lds128<<<grid,256,32768>>>(d,it2); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_lds128<<<grid,256,32768>>>(d,it2); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms,e0,e1);
    double bytes=(double)grid*256*8*16*it2;
    printf("LDS.128 %.3f ms %.1f TB/s (%.1f B/SM/clk)\n", ms, bytes/ms/1e9, bytes/ms/1e-3/148/1.965e9);
  }
  return 0;
}
