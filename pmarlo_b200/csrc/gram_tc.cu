// gram_tc.cu -- K3 (tensor-core path): placeholder until the tcgen05 kernel lands.
#include "common.cuh"
namespace pmb {
bool gram_tcgen05_supported(int, int64_t, const float*) { return false; }
int gram_tcgen05(const float*, int64_t, int, int64_t, const uint8_t*, int, int, const float*,
                 const float*, double*, void*, size_t, cudaStream_t) {
  set_error("tcgen05 Gram path not built");
  return PMB_EUNSUPPORTED;
}
}  // namespace pmb
