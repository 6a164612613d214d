// bf16 tensor-core path of the encoder (tcgen05 implicit-GEMM Conv1d).  Placeholder until the kernel lands.
#include "aa_common.cuh"
#include "encoder.cuh"

namespace aa {
struct TcState { int dummy; };
int tc_create(TcState** st, const std::vector<ConvLayer>&) { *st = nullptr; set_error("bf16 encoder path not built"); return AA_ERR_UNSUPPORTED; }
void tc_destroy(TcState*) {}
void tc_invalidate_weights(TcState*) {}
int64_t tc_workspace_bytes(const std::vector<ConvLayer>&, int64_t, int64_t) { return 256; }
int tc_forward(TcState*, const std::vector<ConvLayer>&, const std::vector<float*>&, const std::vector<float*>&, const float* const*,
               const float*, int, int64_t, int64_t, int, float*, void*, cudaStream_t) {
  set_error("bf16 encoder path not built");
  return AA_ERR_UNSUPPORTED;
}
}  // namespace aa
