// gram_tc.cu — tcgen05 tensor-core kernels (Gram, scores). Placeholder until the tcgen05 path lands.
#include "common.cuh"
namespace srb {
void gram_tcgen05(srb_ctx *, const __half *, const __half *, uint64_t, uint32_t, double *) {
    throw Error(SRB_ERR_UNSUPPORTED, "tcgen05 Gram kernel not built yet: use gram_mode=1");
}
void scores_tcgen05(srb_ctx *, const __half *, const __half *, uint64_t, uint32_t, const double *, const double *, uint32_t, double *) {
    throw Error(SRB_ERR_UNSUPPORTED, "tcgen05 scores kernel not built yet: use gram_mode=1");
}
}  // namespace srb
