// placeholder until the tcgen05 path lands: the BF16 numeric mode reports UDH_ENOSUP.
#include "conv_tc.cuh"

namespace udh {

size_t tc_workspace_bytes(int, int, int) { return 0; }

int tc_cnn_fwd_convs(const float*, const size_t*, const float*, const float*, void*, const size_t*, size_t, int, int,
                     cudaStream_t) {
  set_error("UDH_NUMERIC_BF16 is not available in this build");
  return UDH_ENOSUP;
}

int tc_cnn_bwd_convs(const float*, const size_t*, const float*, const float*, float*, float*, float*, void*, const size_t*,
                     size_t, int, int, cudaStream_t) {
  set_error("UDH_NUMERIC_BF16 is not available in this build");
  return UDH_ENOSUP;
}

}  // namespace udh
