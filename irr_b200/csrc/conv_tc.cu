// tcgen05 implicit-GEMM conv path — placeholder until the kernel lands.
#include "common.cuh"
namespace irr {
bool tc_supported(int, int, int, int, int) { return false; }
size_t tc_packed_bytes(int, int, int, int) { return 0; }
int tc_pack(const float*, void*, int, int, int, int, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return IRR_E_UNSUPPORTED;
}
int tc_conv(const float*, long long, const void*, const float*, const float*, long long, float*, long long, int, int,
            int, int, int, int, int, int, float, float, int, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return IRR_E_UNSUPPORTED;
}
}  // namespace irr
