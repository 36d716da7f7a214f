// K2 / K2b, SNB_PREC_BF16 back end (tcgen05 / TMEM / bulk-copy).  Placeholder until the kernel lands.
#include "common.cuh"
#include "handle.h"
namespace snb {
size_t tc_packed_bytes(const snb_handle_s*) { return 16; }
int tc_pack_weights(snb_handle_s*, void*, cudaStream_t) { set_error("bf16 back end not built yet"); return 3; }
size_t tc_workspace_bytes(const snb_handle_s*, int64_t, int64_t) { return 0; }
size_t tc_bwd_scratch_bytes(const snb_handle_s*, int64_t, int64_t) { return 0; }
int tc_forward(const snb_handle_s*, const float*, const float*, int64_t, int64_t, const float*, const float*, float*, float*,
               void*, cudaStream_t) { set_error("bf16 back end not built yet"); return 3; }
int tc_backward(const snb_handle_s*, const float*, const float*, int64_t, int64_t, const float*, const float*, const float*,
                const float*, const float*, const void*, void*, float*, float*, float*, float*, float* const*, cudaStream_t) {
  set_error("bf16 back end not built yet"); return 3;
}
}  // namespace snb
