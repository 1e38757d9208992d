// placeholder until the tcgen05 path lands
#include "model.cuh"
namespace sag {
int launch_gather_gemm_umma(int precision, const float*, const float*, float*, const GatherGeom&, const Epilogue&, cudaStream_t) {
  set_error("precision %d (tcgen05 path) is not built", precision);
  return SAG_EUNSUPPORTED;
}
}
