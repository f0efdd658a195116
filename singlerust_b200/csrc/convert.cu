// convert.cu — device CSC -> CSR conversion, so the selected densify (convert_to_array_f64_csc_selected,
// src/shared/mod.rs:261-290) and pca_inplace work on CSC-stored X exactly like on CSR. Stable LSD radix sort of the
// entry positions by row index: entries arrive in column-major order, so inside every row the columns stay ascending
// (canonical CSR). Rare path (the headline pipeline is CSR); built from cub primitives.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace srb {

__global__ void row_hist_kernel(const uint32_t *__restrict__ idx, uint64_t nnz, unsigned long long *__restrict__ cnt) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&cnt[idx[k]], 1ULL);
}
// colid[k] = column of entry k (expands the CSC offsets); pos[k] = k
__global__ void expand_major_kernel(const int64_t *__restrict__ off, uint64_t nmajor, uint32_t *__restrict__ colid,
                                    uint32_t *__restrict__ pos) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t c = warp; c < nmajor; c += nwarps)
        for (int64_t k = off[c] + lane; k < off[c + 1]; k += 32) colid[k] = (uint32_t)c, pos[k] = (uint32_t)k;
}
template <typename VT>
__global__ void gather_kernel(const uint32_t *__restrict__ pos_sorted, const uint32_t *__restrict__ colid,
                              const VT *__restrict__ val, uint64_t nnz, uint32_t *__restrict__ out_idx, VT *__restrict__ out_val) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t p = pos_sorted[k];
        out_idx[k] = colid[p];
        out_val[k] = val[p];
    }
}

// Returns a CSR matrix with the same logical content as the CSC matrix m (pending transforms applied first).
srb_mat *csc_to_csr(srb_mat *m) {
    SRB_REQUIRE(m->format == SRB_CSC, SRB_ERR_INVALID_ARG, "csc_to_csr needs a CSC matrix");
    if (m->has_pending()) materialize(m, false);
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    const uint64_t nnz = m->st->nnz, ncols = m->ncols, nrows = m->nrows;
    SRB_REQUIRE(nnz < (1ull << 32), SRB_ERR_UNSUPPORTED, "CSC -> CSR conversion supports fewer than 2^32 stored entries");
    auto st = std::make_shared<Structure>();
    st->nmajor = nrows, st->nminor = ncols, st->nnz = nnz;
    st->offsets = dev_zeros(s, 8 * (nrows + 1));
    st->indices = dev_alloc(s, 4 * std::max<uint64_t>(nnz, 1));
    std::unique_ptr<srb_mat> out(new srb_mat());
    out->ctx = c, out->format = SRB_CSR, out->nrows = nrows, out->ncols = ncols, out->st = st;
    out->vdtype = m->vdtype, out->src_dtype = m->src_dtype;
    out->global_row0 = 0, out->global_nrows = nrows;
    out->values = dev_alloc(s, (m->vdtype == SRB_F32 ? 4 : 8) * std::max<uint64_t>(nnz, 1));
    if (nnz == 0) return out.release();
    const unsigned g = (unsigned)std::min<uint64_t>((nnz + 255) / 256, (uint64_t)c->sm_count * 16);
    Buf cnt = dev_zeros(s, 8 * (nrows + 1));
    SRB_LAUNCH(row_hist_kernel, g, 256, 0, s, m->st->indices->as<uint32_t>(), nnz, cnt->as<unsigned long long>());
    size_t tb = 0;
    SRB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt->as<int64_t>(), st->offsets->as<int64_t>(), (int64_t)(nrows + 1), s));
    Buf tmp = dev_alloc(s, tb);
    SRB_CUDA(cub::DeviceScan::ExclusiveSum(tmp->p, tb, cnt->as<int64_t>(), st->offsets->as<int64_t>(), (int64_t)(nrows + 1), s));
    Buf colid = dev_alloc(s, 4 * nnz), pos = dev_alloc(s, 4 * nnz), pos_sorted = dev_alloc(s, 4 * nnz), key_sorted = dev_alloc(s, 4 * nnz);
    SRB_LAUNCH(expand_major_kernel, (unsigned)std::min<uint64_t>((ncols + 7) / 8, (uint64_t)c->sm_count * 32), 256, 0, s,
               m->st->offsets->as<int64_t>(), ncols, colid->as<uint32_t>(), pos->as<uint32_t>());
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < nrows) ++end_bit;
    tb = 0;
    SRB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, m->st->indices->as<uint32_t>(), key_sorted->as<uint32_t>(), pos->as<uint32_t>(),
                                             pos_sorted->as<uint32_t>(), (int64_t)nnz, 0, end_bit, s));
    Buf tmp2 = dev_alloc(s, tb);
    SRB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2->p, tb, m->st->indices->as<uint32_t>(), key_sorted->as<uint32_t>(), pos->as<uint32_t>(),
                                             pos_sorted->as<uint32_t>(), (int64_t)nnz, 0, end_bit, s));
    g_launches.fetch_add(6, std::memory_order_relaxed);
    if (m->vdtype == SRB_F32)
        SRB_LAUNCH((gather_kernel<float>), g, 256, 0, s, pos_sorted->as<uint32_t>(), colid->as<uint32_t>(), m->values->as<float>(), nnz, st->indices->as<uint32_t>(), out->values->as<float>());
    else
        SRB_LAUNCH((gather_kernel<double>), g, 256, 0, s, pos_sorted->as<uint32_t>(), colid->as<uint32_t>(), m->values->as<double>(), nnz, st->indices->as<uint32_t>(), out->values->as<double>());
    SRB_CUDA(cudaStreamSynchronize(s));
    return out.release();
}

}  // namespace srb
