// Internal interface between encoder.cu (layer table, fp32 path, sequencing) and conv_tc.cu (bf16 tcgen05 path).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "../../include/aa_b200.h"

namespace aa {

enum { ROLE_PLAIN = 0, ROLE_RES_FIRST = 1, ROLE_RES_SECOND = 2 };

struct ConvLayer {
  int cin, cout, k, stride, dil, pad;
  int elu;    // ELU after bias (+ residual)
  int role;   // ROLE_RES_FIRST: input is also the residual of the next layer; ROLE_RES_SECOND: adds it
};

int build_layer_table(const AaEncoderCfg& cfg, std::vector<ConvLayer>& layers);

struct TcState;
int tc_create(TcState** st, const std::vector<ConvLayer>& layers);
void tc_destroy(TcState* st);
void tc_invalidate_weights(TcState* st);
int64_t tc_workspace_bytes(const std::vector<ConvLayer>& layers, int64_t batch, int64_t n);
int tc_forward(TcState* st, const std::vector<ConvLayer>& layers, const std::vector<float*>& w, const std::vector<float*>& b,
               const float* const* stems_host, const float* faders_host, int n_stems, int64_t batch, int64_t n, int apply_tanh,
               float* y, void* workspace, cudaStream_t stream);

// 3xTF32 tensor-core path (conv_tf32.cuh): fp32-grade results, same call shape as the bf16 path
struct TfState;
bool tf_eligible(const std::vector<ConvLayer>& layers);
int tf_create(TfState** st, const std::vector<ConvLayer>& layers);
void tf_destroy(TfState* st);
void tf_invalidate_weights(TfState* st);
int64_t tf_workspace_bytes(const std::vector<ConvLayer>& layers, int64_t batch, int64_t n);
int tf_forward(TfState* st, const std::vector<ConvLayer>& layers, const std::vector<float*>& w, const std::vector<float*>& b,
               const float* const* stems_host, const float* faders_host, int n_stems, int64_t batch, int64_t n, int apply_tanh,
               float* y, void* workspace, cudaStream_t stream);

}  // namespace aa
