// tcgen05 implicit-GEMM convolution back end -- placeholder until the kernels land.
#include "common.cuh"

int tc_prepare(lc_codec* h) {
  if (h->precision != LC_PREC_FP32) {
    lc_set_error("tensor-core precision modes are not built yet");
    return LC_EUNSUPPORTED;
  }
  return LC_OK;
}
size_t tc_workspace_bytes(const lc_codec*, int) { return 0; }
int tc_encode(lc_codec*, const float*, const ChunkGeom&, int, float, float, int, int, float*, void*, cudaStream_t) {
  lc_set_error("tensor-core path not built");
  return LC_EUNSUPPORTED;
}
int tc_decode(lc_codec*, const float*, const ChunkGeom&, int, int, float*, void*, cudaStream_t) {
  lc_set_error("tensor-core path not built");
  return LC_EUNSUPPORTED;
}
