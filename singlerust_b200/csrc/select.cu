// select.cu — K5: highly-variable-gene selection. Replaces select_features(HighlyVariable(n)),
// src/memory/processing/dim_red/mod.rs:135-140: per-gene variance (nonzero-only, one-pass form) ->
// stable sort by DESCENDING variance -> first n indices in that order (ties keep ascending index).
// 30 k keys: a stable LSD radix sort (cub::DeviceRadixSort, stable by construction) on the device keeps the
// pipeline free of a host round-trip; the cost is negligible next to the nnz passes.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace srb {

__global__ void iota_kernel(uint32_t *a, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}
// partial_cmp treats -0.0 == +0.0; the radix order does not. Canonicalise so ties stay index-ordered.
__global__ void canon_zero_nan_kernel(double *v, uint64_t n, uint32_t *nan_flag) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = v[i];
    if (x != x) atomicOr(nan_flag, 1u);
    if (x == 0.0) v[i] = 0.0;
}

void select_hvg_device(srb_mat *m, uint64_t n_top, Buf &d_idx_u32, uint64_t *n_out, bool check_nan) {
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    // per-gene variance = Column direction (dim_red/mod.rs:136)
    const uint64_t M = m->ncols;
    SRB_REQUIRE(M > 0, SRB_ERR_INVALID_ARG, "matrix has no columns");
    Buf var = dev_alloc(s, 8 * M), var_sorted = dev_alloc(s, 8 * M);
    Buf idx = dev_alloc(s, 4 * M), idx_sorted = dev_alloc(s, 4 * M);
    Buf flag = dev_zeros(s, 4);
    if (m->format == SRB_CSR) minor_variance_from_moments(m, var->as<double>(), false);
    else major_variance(m, var->as<double>());
    StageTimer t(c, ST_HVG);
    const unsigned g = (unsigned)((M + 255) / 256);
    SRB_LAUNCH(canon_zero_nan_kernel, g, 256, 0, s, var->as<double>(), M, flag->as<uint32_t>());
    SRB_LAUNCH(iota_kernel, g, 256, 0, s, idx->as<uint32_t>(), M);
    size_t tmp_bytes = 0;
    SRB_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, var->as<double>(), var_sorted->as<double>(),
                                                       idx->as<uint32_t>(), idx_sorted->as<uint32_t>(), (int)M, 0, 64, s));
    Buf tmp = dev_alloc(s, tmp_bytes);
    SRB_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp->p, tmp_bytes, var->as<double>(), var_sorted->as<double>(),
                                                       idx->as<uint32_t>(), idx_sorted->as<uint32_t>(), (int)M, 0, 64, s));
    g_launches.fetch_add(4, std::memory_order_relaxed);  // onesweep: histogram + 3 passes (approx.)
    if (check_nan) {
        uint32_t h = 0;
        SRB_CUDA(cudaMemcpyAsync(&h, flag->p, 4, cudaMemcpyDeviceToHost, s));
        SRB_CUDA(cudaStreamSynchronize(s));
        SRB_REQUIRE(!h, SRB_ERR_NAN, "NaN variance in HighlyVariable selection (the reference panics in sort_by(partial_cmp().unwrap()), dim_red/mod.rs:138)");
    }
    d_idx_u32 = idx_sorted;
    *n_out = n_top < M ? n_top : M;
}

}  // namespace srb
