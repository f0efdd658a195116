// subset.cu — N2: row / column compaction of a device-resident compressed matrix, the device half of
// filter_cells / filter_genes (src/memory/processing/mod.rs:86-146, 245-299), which the reference hands to
// IMAnnData::subset{,_inplace} (anndata-memory) after building a boolean mask from per-cell / per-gene counts and
// sums. HBM-bound stream kernels: one warp per kept major line, ballot compaction of its kept entries.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace srb {

// kept entries per OLD major line (0 for dropped lines)
__global__ void __launch_bounds__(256) subset_count_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                           const uint8_t *__restrict__ keep_major, const uint8_t *__restrict__ keep_minor,
                                                           uint64_t nmajor, int64_t *__restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        int64_t c = 0;
        if (!keep_major || keep_major[r]) {
            const int64_t a = off[r], b = off[r + 1];
            if (!keep_minor) {
                c = b - a;
            } else {
                for (int64_t k = a + lane; k < b; k += 32) c += keep_minor[idx[k]] ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            }
        }
        if (lane == 0) cnt[r] = c;
    }
}
__global__ void flags_to_i64_kernel(const uint8_t *__restrict__ keep, uint64_t n, int64_t *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = keep ? (keep[i] ? 1 : 0) : 1;
    if (i == n) out[i] = 0;
}
// new offsets: new_off[new_line] = prefix of kept counts; old line r maps to new line line_map[r]
__global__ void subset_offsets_kernel(const int64_t *__restrict__ cnt_prefix, const int64_t *__restrict__ line_map,
                                      const uint8_t *__restrict__ keep_major, uint64_t nmajor, uint64_t new_nmajor,
                                      int64_t *__restrict__ new_off) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nmajor && (!keep_major || keep_major[r])) new_off[line_map[r]] = cnt_prefix[r];
    if (r == nmajor) new_off[new_nmajor] = cnt_prefix[nmajor];
}
template <typename VT>
__global__ void __launch_bounds__(256) subset_fill_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                          const VT *__restrict__ val, const uint8_t *__restrict__ keep_major,
                                                          const uint8_t *__restrict__ keep_minor, const int64_t *__restrict__ minor_map,
                                                          const int64_t *__restrict__ cnt_prefix, uint64_t nmajor,
                                                          uint32_t *__restrict__ out_idx, VT *__restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        if (keep_major && !keep_major[r]) continue;
        const int64_t a = off[r], b = off[r + 1];
        int64_t o = cnt_prefix[r];
        for (int64_t k0 = a; k0 < b; k0 += 32) {
            const int64_t k = k0 + lane;
            uint32_t c = 0;
            bool keep = false;
            if (k < b) {
                c = idx[k];
                keep = !keep_minor || keep_minor[c];
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int64_t pos = o + __popc(mask & ((1u << lane) - 1u));
                out_idx[pos] = keep_minor ? (uint32_t)minor_map[c] : c;
                out_val[pos] = val[k];
            }
            o += __popc(mask);
        }
    }
}

static void exclusive_scan_i64(cudaStream_t s, const int64_t *in, int64_t *out, uint64_t n) {
    size_t tb = 0;
    SRB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int64_t)n, s));
    Buf tmp = dev_alloc(s, tb);
    SRB_CUDA(cub::DeviceScan::ExclusiveSum(tmp->p, tb, in, out, (int64_t)n, s));
    g_launches.fetch_add(2, std::memory_order_relaxed);
}

// keep_rows / keep_cols: host byte masks (nullable = keep all)
srb_mat *subset_matrix(srb_mat *m, const uint8_t *keep_rows, const uint8_t *keep_cols) {
    if (m->has_pending()) materialize(m, false);
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    const uint8_t *h_major = m->format == SRB_CSR ? keep_rows : keep_cols;
    const uint8_t *h_minor = m->format == SRB_CSR ? keep_cols : keep_rows;
    const uint64_t nmajor = m->nmajor(), nminor = m->nminor(), nnz = m->st->nnz;
    Buf d_major, d_minor;
    if (h_major) {
        d_major = dev_alloc(s, nmajor + 1);
        SRB_CUDA(cudaMemcpyAsync(d_major->p, h_major, nmajor, cudaMemcpyHostToDevice, s));
    }
    if (h_minor) {
        d_minor = dev_alloc(s, nminor + 1);
        SRB_CUDA(cudaMemcpyAsync(d_minor->p, h_minor, nminor, cudaMemcpyHostToDevice, s));
    }
    const uint8_t *km = d_major ? d_major->as<uint8_t>() : nullptr, *kn = d_minor ? d_minor->as<uint8_t>() : nullptr;
    auto nb = [](uint64_t n) { return (unsigned)std::max<uint64_t>(1, (n + 255) / 256); };
    // line and minor index maps (exclusive scans of the keep flags)
    Buf fl = dev_alloc(s, 8 * (nmajor + 1)), line_map = dev_alloc(s, 8 * (nmajor + 1));
    SRB_LAUNCH(flags_to_i64_kernel, nb(nmajor + 1), 256, 0, s, km, nmajor, fl->as<int64_t>());
    exclusive_scan_i64(s, fl->as<int64_t>(), line_map->as<int64_t>(), nmajor + 1);
    Buf fm = dev_alloc(s, 8 * (nminor + 1)), minor_map = dev_alloc(s, 8 * (nminor + 1));
    SRB_LAUNCH(flags_to_i64_kernel, nb(nminor + 1), 256, 0, s, kn, nminor, fm->as<int64_t>());
    exclusive_scan_i64(s, fm->as<int64_t>(), minor_map->as<int64_t>(), nminor + 1);
    // kept entries per line and their prefix
    Buf cnt = dev_zeros(s, 8 * (nmajor + 1)), cnt_prefix = dev_alloc(s, 8 * (nmajor + 1));
    const unsigned wgrid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nmajor + 7) / 8, (uint64_t)c->sm_count * 32));
    if (nmajor) SRB_LAUNCH(subset_count_kernel, wgrid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), km, kn, nmajor, cnt->as<int64_t>());
    exclusive_scan_i64(s, cnt->as<int64_t>(), cnt_prefix->as<int64_t>(), nmajor + 1);
    int64_t h[3];
    SRB_CUDA(cudaMemcpyAsync(&h[0], line_map->as<int64_t>() + nmajor, 8, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaMemcpyAsync(&h[1], minor_map->as<int64_t>() + nminor, 8, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaMemcpyAsync(&h[2], cnt_prefix->as<int64_t>() + nmajor, 8, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    const uint64_t new_nmajor = (uint64_t)h[0], new_nminor = (uint64_t)h[1], new_nnz = (uint64_t)h[2];
    auto st = std::make_shared<Structure>();
    st->nmajor = new_nmajor, st->nminor = new_nminor, st->nnz = new_nnz;
    st->offsets = dev_zeros(s, 8 * (new_nmajor + 1));
    st->indices = dev_alloc(s, 4 * std::max<uint64_t>(new_nnz, 1));
    std::unique_ptr<srb_mat> out(new srb_mat());
    out->ctx = c, out->format = m->format, out->st = st;
    out->nrows = m->format == SRB_CSR ? new_nmajor : new_nminor;
    out->ncols = m->format == SRB_CSR ? new_nminor : new_nmajor;
    out->vdtype = m->vdtype, out->src_dtype = m->src_dtype;
    out->global_row0 = 0, out->global_nrows = out->nrows;
    out->values = dev_alloc(s, (m->vdtype == SRB_F32 ? 4 : 8) * std::max<uint64_t>(new_nnz, 1));
    SRB_LAUNCH(subset_offsets_kernel, nb(nmajor + 1), 256, 0, s, cnt_prefix->as<int64_t>(), line_map->as<int64_t>(), km, nmajor, new_nmajor, st->offsets->as<int64_t>());
    if (nmajor && nnz) {
        if (m->vdtype == SRB_F32)
            SRB_LAUNCH((subset_fill_kernel<float>), wgrid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), km, kn, minor_map->as<int64_t>(), cnt_prefix->as<int64_t>(), nmajor, st->indices->as<uint32_t>(), out->values->as<float>());
        else
            SRB_LAUNCH((subset_fill_kernel<double>), wgrid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<double>(), km, kn, minor_map->as<int64_t>(), cnt_prefix->as<int64_t>(), nmajor, st->indices->as<uint32_t>(), out->values->as<double>());
    }
    SRB_CUDA(cudaStreamSynchronize(s));
    return out.release();
}

}  // namespace srb

using namespace srb;

extern "C" int32_t srb_mat_subset(srb_mat *m, const uint8_t *keep_rows, const uint8_t *keep_cols, srb_mat **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(m && m->ctx && m->st && out, SRB_ERR_INVALID_ARG, "null argument");
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    srb_mat *r = subset_matrix(m, keep_rows, keep_cols);
    if (m->ctx->nranks > 1) r->global_row0 = 0, r->global_nrows = r->nrows;  // caller re-declares the shard extent
    *out = r;
    SRB_API_END
}
