// synth.cu — device generator of the synthetic count matrix: entry (i, j) exists iff
// mix32(rowkey(i) + j * 0x9E3779B1) < thr[j] * depth(i); integer-only, so the CPU twin (oracle/srb_oracle.c,
// orc_synth_*) produces the identical CSR bit for bit. Benchmark / test input only — not part of the reference.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace srb {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t row_key(uint32_t seed, uint64_t row) { return mix32(seed ^ mix32((uint32_t)row + 0x9E3779B9U)); }
__device__ __forceinline__ uint32_t row_depth(uint32_t r, int skew) { return skew ? 32768U + (mix32(r ^ 0xA511E9B3U) % 98304U) : 65536U; }
__device__ __forceinline__ bool entry(uint32_t r, uint32_t depth, uint32_t col, const uint32_t *__restrict__ thr,
                                      const uint32_t *__restrict__ amp, float *v) {
    const uint32_t h = mix32(r + col * 0x9E3779B1U);
    unsigned long long t = ((unsigned long long)thr[col] * depth) >> 16;
    if (t > 0xFFFFFFFFULL) t = 0xFFFFFFFFULL;
    if ((unsigned long long)h >= t) return false;
    const uint32_t g = mix32(h ^ 0x68E31DA4U);
    uint32_t k = g ? (uint32_t)(__ffs((int)g) - 1) : 32u;
    if (k > 15) k = 15;
    *v = (float)(1U + ((k * (16U + amp[col])) >> 4));
    return true;
}

__global__ void __launch_bounds__(256) synth_count_kernel(uint32_t seed, int skew, uint64_t row0, uint64_t nrows, uint32_t ncols,
                                                          const uint32_t *__restrict__ thr, const uint32_t *__restrict__ amp,
                                                          int64_t *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t i = warp; i < nrows; i += nwarps) {
        const uint32_t r = row_key(seed, row0 + i), d = row_depth(r, skew);
        uint32_t c = 0;
        float v;
        for (uint32_t j0 = 0; j0 < ncols; j0 += 32) {
            const uint32_t j = j0 + lane;
            const bool p = j < ncols && entry(r, d, j, thr, amp, &v);
            c += __popc(__ballot_sync(0xffffffffu, p));
        }
        if (lane == 0) counts[i] = c;
    }
}
__global__ void __launch_bounds__(256) synth_fill_kernel(uint32_t seed, int skew, uint64_t row0, uint64_t nrows, uint32_t ncols,
                                                         const uint32_t *__restrict__ thr, const uint32_t *__restrict__ amp,
                                                         const int64_t *__restrict__ off, uint32_t *__restrict__ idx,
                                                         float *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t i = warp; i < nrows; i += nwarps) {
        const uint32_t r = row_key(seed, row0 + i), d = row_depth(r, skew);
        int64_t o = off[i];
        for (uint32_t j0 = 0; j0 < ncols; j0 += 32) {
            const uint32_t j = j0 + lane;
            float v = 0.f;
            const bool p = j < ncols && entry(r, d, j, thr, amp, &v);
            const unsigned mask = __ballot_sync(0xffffffffu, p);
            if (p) {
                const int64_t pos = o + __popc(mask & ((1u << lane) - 1u));
                idx[pos] = j;
                val[pos] = v;
            }
            o += __popc(mask);
        }
    }
}

}  // namespace srb

using namespace srb;

extern "C" int32_t srb_synth_csr(srb_ctx *ctx, uint32_t seed, int32_t skew, uint64_t row0, uint64_t nrows, uint32_t ncols,
                                 const uint32_t *thr, const uint32_t *amp, srb_mat **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx && thr && amp && out, SRB_ERR_INVALID_ARG, "null argument");
    SRB_REQUIRE(ncols > 0, SRB_ERR_INVALID_ARG, "ncols must be positive");
    SRB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    Buf d_thr = dev_alloc(s, 4 * (size_t)ncols), d_amp = dev_alloc(s, 4 * (size_t)ncols);
    SRB_CUDA(cudaMemcpyAsync(d_thr->p, thr, 4 * (size_t)ncols, cudaMemcpyHostToDevice, s));
    SRB_CUDA(cudaMemcpyAsync(d_amp->p, amp, 4 * (size_t)ncols, cudaMemcpyHostToDevice, s));
    auto st = std::make_shared<Structure>();
    st->nmajor = nrows, st->nminor = ncols;
    st->offsets = dev_zeros(s, 8 * (nrows + 1));
    Buf counts = dev_zeros(s, 8 * (nrows + 1));
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nrows + 7) / 8, (uint64_t)ctx->sm_count * 32));
    if (nrows) SRB_LAUNCH(synth_count_kernel, grid, 256, 0, s, seed, skew, row0, nrows, ncols, d_thr->as<uint32_t>(), d_amp->as<uint32_t>(), counts->as<int64_t>());
    // exclusive scan over nrows+1 entries (the last input is 0) -> offsets[nrows] = nnz
    size_t tmp_bytes = 0;
    SRB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts->as<int64_t>(), st->offsets->as<int64_t>(), (int64_t)(nrows + 1), s));
    Buf tmp = dev_alloc(s, tmp_bytes);
    SRB_CUDA(cub::DeviceScan::ExclusiveSum(tmp->p, tmp_bytes, counts->as<int64_t>(), st->offsets->as<int64_t>(), (int64_t)(nrows + 1), s));
    int64_t nnz = 0;
    SRB_CUDA(cudaMemcpyAsync(&nnz, st->offsets->as<int64_t>() + nrows, 8, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    st->nnz = (uint64_t)nnz;
    st->indices = dev_alloc(s, 4 * (size_t)std::max<int64_t>(nnz, 1));
    std::unique_ptr<srb_mat> m(new srb_mat());
    m->ctx = ctx, m->format = SRB_CSR, m->nrows = nrows, m->ncols = ncols, m->st = st;
    m->vdtype = SRB_F32, m->src_dtype = SRB_F32;
    m->values = dev_alloc(s, 4 * (size_t)std::max<int64_t>(nnz, 1));
    m->global_row0 = 0, m->global_nrows = nrows;
    if (nrows) SRB_LAUNCH(synth_fill_kernel, grid, 256, 0, s, seed, skew, row0, nrows, ncols, d_thr->as<uint32_t>(), d_amp->as<uint32_t>(), st->offsets->as<int64_t>(), st->indices->as<uint32_t>(), m->values->as<float>());
    SRB_CUDA(cudaStreamSynchronize(s));
    *out = m.release();
    SRB_API_END
}
