// (iv) impl 2: tcgen05 3xTF32 candidate generation + exact fp32 rescoring.  (placeholder
// until the tensor-core kernel lands; fails loudly rather than falling back.)
#include "common.cuh"

namespace srb {
int score_topk_tc(const srb_topk_desc* d, cudaStream_t st) {
  (void)d;
  (void)st;
  set_error("topk: impl 2 (tcgen05) is not built in this version");
  return SRB_ERR_ARG;
}
}  // namespace srb

extern "C" int64_t srb_topk_workspace_bytes(int32_t n_q, int32_t n_items, int32_t d, int32_t k) {
  (void)n_q;
  (void)n_items;
  (void)d;
  (void)k;
  return 0;
}
