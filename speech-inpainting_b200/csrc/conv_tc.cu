// bf16 tcgen05 implicit-GEMM convolution - placeholder until the TMA/tcgen05 kernel lands (next commit).
#include "common.cuh"
extern "C" int sib_conv1d_bf16(const sib_conv_desc*, const void*, const void*, const float*, const void*, void*, void*,
                               sib_stream_t) {
  sib::set_error("sib_conv1d_bf16: not built in this revision");
  return SIB_ERR_UNSUPPORTED;
}
