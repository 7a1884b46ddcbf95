// placeholder until the tcgen05 kernel lands (next commit): reports "unsupported" so bf16 mode uses the SIMT kernel
#include "common.cuh"
namespace regat {
bool gemm_tc_supported(int, int, int, int, int, const void*, int, const void*, int) { return false; }
int gemm_tc(int, int, int, int, int, const void*, int, const void*, int, void*, int, int, const EpiArgs&, int, cudaStream_t) {
  set_error("gemm_tc: not built");
  return REGAT_ERR_UNSUPPORTED;
}
}  // namespace regat
