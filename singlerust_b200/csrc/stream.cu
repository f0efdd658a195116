// stream.cu — K10 (first form): chunked accumulation for backed (out-of-core) data. Replaces the chunk drivers
// shared::statistics::{number,sum}::chunked (src/shared/statistics/mod.rs:17-41, 59-83) and the chunk helpers
// csr.rs:48-74,112-143 / csc.rs:45-68,105-131. Each pushed chunk is uploaded, reduced on the device with the same
// kernels as the whole-matrix path, and folded into running device accumulators:
//   major direction: results are written at the chunk's running major offset (the reference drops that offset and
//                    mis-places them — SURVEY §10; deviation documented in DESIGN.md)
//   minor direction: count/sum/sumsq totals += chunk moments (chunk order => deterministic)
#include <algorithm>
#include <memory>

#include "common.cuh"

struct srb_stream {
    srb_ctx *ctx = nullptr;
    int format = SRB_CSR;
    uint64_t nrows_total = 0, ncols_total = 0;
    uint64_t nmajor_total = 0, nminor = 0, major_pos = 0;
    srb::Buf major_cnt;  // u32[nmajor_total]
    srb::Buf major_sum;  // f64[nmajor_total]
    srb::Buf major_var;  // f64[nmajor_total]
    srb::Buf minor_cnt, minor_sum, minor_sq;  // f64[nminor]
    // optional residency: pushed chunks are appended to one device-resident matrix (srb_stream_finish_matrix)
    bool retain = false, stats = true;
    uint64_t nnz_pos = 0, nnz_cap = 0, nnz_hint = 0;
    int vdtype = -1, src_dtype = SRB_F32;
    srb::Buf r_off, r_idx, r_val;
};

namespace srb {
__global__ void axpy1_kernel(double *__restrict__ acc, const double *__restrict__ x, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += x[i];
}
__global__ void counts_at_kernel(const int64_t *__restrict__ off, uint32_t *__restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(off[i + 1] - off[i]);
}
__global__ void stream_var_kernel(const double *cnt, const double *sum, const double *sq, uint64_t n, double *out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double r = 0.0;
    if (cnt[j] > 0.0) {
        const double mean = sum[j] / cnt[j];
        r = __dsub_rn(__ddiv_rn(sq[j], cnt[j]), __dmul_rn(mean, mean));
    }
    out[j] = r;
}
__global__ void rebase_offsets_kernel(const int64_t *__restrict__ src, int64_t *__restrict__ dst, uint64_t n, int64_t base) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + base;  // dst already points at the chunk's first line; entry n is the running total
}
__global__ void f64_to_u32_kernel2(const double *in, uint32_t *out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}
static unsigned nblk(uint64_t n) { return (unsigned)std::max<uint64_t>(1, (n + 255) / 256); }
}  // namespace srb

using namespace srb;

extern "C" {

int32_t srb_stream_begin(srb_ctx *ctx, int32_t format, uint64_t nrows_total, uint64_t ncols_total, srb_stream **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx && out, SRB_ERR_INVALID_ARG, "null argument");
    SRB_REQUIRE(format == SRB_CSR || format == SRB_CSC, SRB_ERR_INVALID_ARG, "format must be CSR or CSC");
    SRB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::unique_ptr<srb_stream> st(new srb_stream());
    st->ctx = ctx, st->format = format, st->nrows_total = nrows_total, st->ncols_total = ncols_total;
    st->nmajor_total = format == SRB_CSR ? nrows_total : ncols_total;
    st->nminor = format == SRB_CSR ? ncols_total : nrows_total;
    st->major_cnt = dev_zeros(s, 4 * (st->nmajor_total + 1));
    st->major_sum = dev_zeros(s, 8 * (st->nmajor_total + 1));
    st->major_var = dev_zeros(s, 8 * (st->nmajor_total + 1));
    st->minor_cnt = dev_zeros(s, 8 * (st->nminor + 1));
    st->minor_sum = dev_zeros(s, 8 * (st->nminor + 1));
    st->minor_sq = dev_zeros(s, 8 * (st->nminor + 1));
    *out = st.release();
    SRB_API_END
}

int32_t srb_stream_push(srb_stream *st, uint64_t nmajor_chunk, uint64_t nnz, const void *offsets, const void *indices,
                        int32_t idx_width, const void *values, int32_t dtype) {
    SRB_API_BEGIN
    SRB_REQUIRE(st, SRB_ERR_INVALID_ARG, "null stream");
    SRB_REQUIRE(st->major_pos + nmajor_chunk <= st->nmajor_total, SRB_ERR_INVALID_ARG, "more lines pushed than announced");
    srb_ctx *c = st->ctx;
    cudaStream_t s = c->stream;
    srb_mat *chunk = nullptr;
    const uint64_t nr = st->format == SRB_CSR ? nmajor_chunk : st->nrows_total;
    const uint64_t nc = st->format == SRB_CSR ? st->ncols_total : nmajor_chunk;
    int32_t rc = srb_mat_upload(c, st->format, nr, nc, nnz, offsets, indices, idx_width, values, dtype, &chunk);
    if (rc != SRB_OK) return rc;
    std::unique_ptr<srb_mat> guard(chunk);
    if (st->retain) {
        if (st->vdtype < 0) st->vdtype = chunk->vdtype, st->src_dtype = chunk->src_dtype;
        SRB_REQUIRE(st->vdtype == chunk->vdtype, SRB_ERR_INVALID_ARG, "all chunks of a resident stream must share one value dtype class");
        const size_t vsz = st->vdtype == SRB_F32 ? 4 : 8;
        if (st->nnz_pos + nnz > st->nnz_cap) {  // grow geometrically (block cache: old blocks are recycled)
            const uint64_t ncap = std::max<uint64_t>(std::max<uint64_t>(st->nnz_pos + nnz, st->nnz_hint), st->nnz_cap + st->nnz_cap / 2 + 1024);
            Buf ni = dev_alloc(s, 4 * ncap), nv = dev_alloc(s, vsz * ncap);
            if (st->nnz_pos) {
                SRB_CUDA(cudaMemcpyAsync(ni->p, st->r_idx->p, 4 * st->nnz_pos, cudaMemcpyDeviceToDevice, s));
                SRB_CUDA(cudaMemcpyAsync(nv->p, st->r_val->p, vsz * st->nnz_pos, cudaMemcpyDeviceToDevice, s));
            }
            st->r_idx = ni, st->r_val = nv, st->nnz_cap = ncap;
        }
        if (nnz) {
            SRB_CUDA(cudaMemcpyAsync(st->r_idx->as<uint32_t>() + st->nnz_pos, chunk->st->indices->p, 4 * nnz, cudaMemcpyDeviceToDevice, s));
            SRB_CUDA(cudaMemcpyAsync((char *)st->r_val->p + vsz * st->nnz_pos, chunk->values->p, vsz * nnz, cudaMemcpyDeviceToDevice, s));
        }
        SRB_LAUNCH(rebase_offsets_kernel, nblk(nmajor_chunk + 1), 256, 0, s, chunk->st->offsets->as<int64_t>(),
                   st->r_off->as<int64_t>() + st->major_pos, nmajor_chunk + 1, (int64_t)st->nnz_pos);
        st->nnz_pos += nnz;
    }
    {
        // chunk moments are local (never reduced over ranks); the caller reduces at the end if it shards chunks
        if (!st->stats) {
            // residency only
        } else if (nmajor_chunk) {
            major_sum_absmax(chunk);
            SRB_LAUNCH(counts_at_kernel, nblk(nmajor_chunk), 256, 0, s, chunk->st->offsets->as<int64_t>(), st->major_cnt->as<uint32_t>() + st->major_pos, nmajor_chunk);
            SRB_CUDA(cudaMemcpyAsync(st->major_sum->as<double>() + st->major_pos, chunk->major.sum->p, 8 * nmajor_chunk, cudaMemcpyDeviceToDevice, s));
            major_variance(chunk, st->major_var->as<double>() + st->major_pos);
        }
        if (st->stats) ensure_minor_moments(chunk, /*local_only=*/true);
        const uint64_t M = st->nminor;
        if (st->stats && M) {
            SRB_LAUNCH(axpy1_kernel, nblk(M), 256, 0, s, st->minor_cnt->as<double>(), chunk->minor.cnt->as<double>(), M);
            SRB_LAUNCH(axpy1_kernel, nblk(M), 256, 0, s, st->minor_sum->as<double>(), chunk->minor.sum->as<double>(), M);
            SRB_LAUNCH(axpy1_kernel, nblk(M), 256, 0, s, st->minor_sq->as<double>(), chunk->minor.sq->as<double>(), M);
        }
    }
    st->major_pos += nmajor_chunk;
    SRB_CUDA(cudaStreamSynchronize(s));  // host arrays are only borrowed for the duration of the call
    SRB_API_END
}

static bool stream_dir_is_major(const srb_stream *st, int direction) { return (direction == SRB_ROW) == (st->format == SRB_CSR); }

int32_t srb_stream_number(srb_stream *st, int32_t direction, uint32_t *out) {
    SRB_API_BEGIN
    SRB_REQUIRE(st && out, SRB_ERR_INVALID_ARG, "null argument");
    cudaStream_t s = st->ctx->stream;
    if (stream_dir_is_major(st, direction)) {
        SRB_CUDA(cudaMemcpyAsync(out, st->major_cnt->p, 4 * st->nmajor_total, cudaMemcpyDeviceToHost, s));
    } else {
        Buf t = dev_alloc(s, 4 * (st->nminor + 1));
        if (st->nminor) SRB_LAUNCH(f64_to_u32_kernel2, nblk(st->nminor), 256, 0, s, st->minor_cnt->as<double>(), t->as<uint32_t>(), st->nminor);
        SRB_CUDA(cudaMemcpyAsync(out, t->p, 4 * st->nminor, cudaMemcpyDeviceToHost, s));
    }
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_API_END
}

int32_t srb_stream_sum(srb_stream *st, int32_t direction, double *out) {
    SRB_API_BEGIN
    SRB_REQUIRE(st && out, SRB_ERR_INVALID_ARG, "null argument");
    cudaStream_t s = st->ctx->stream;
    if (stream_dir_is_major(st, direction)) SRB_CUDA(cudaMemcpyAsync(out, st->major_sum->p, 8 * st->nmajor_total, cudaMemcpyDeviceToHost, s));
    else SRB_CUDA(cudaMemcpyAsync(out, st->minor_sum->p, 8 * st->nminor, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_API_END
}

int32_t srb_stream_variance(srb_stream *st, int32_t direction, double *out) {
    SRB_API_BEGIN
    SRB_REQUIRE(st && out, SRB_ERR_INVALID_ARG, "null argument");
    cudaStream_t s = st->ctx->stream;
    if (stream_dir_is_major(st, direction)) {
        SRB_CUDA(cudaMemcpyAsync(out, st->major_var->p, 8 * st->nmajor_total, cudaMemcpyDeviceToHost, s));
    } else {
        Buf t = dev_alloc(s, 8 * (st->nminor + 1));
        if (st->nminor) SRB_LAUNCH(stream_var_kernel, nblk(st->nminor), 256, 0, s, st->minor_cnt->as<double>(), st->minor_sum->as<double>(), st->minor_sq->as<double>(), st->nminor, t->as<double>());
        SRB_CUDA(cudaMemcpyAsync(out, t->p, 8 * st->nminor, cudaMemcpyDeviceToHost, s));
    }
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_API_END
}

int32_t srb_stream_set_retain(srb_stream *st, uint64_t nnz_hint, int32_t keep_statistics) {
    SRB_API_BEGIN
    SRB_REQUIRE(st, SRB_ERR_INVALID_ARG, "null stream");
    SRB_REQUIRE(st->major_pos == 0, SRB_ERR_INVALID_ARG, "set_retain must precede the first push");
    cudaStream_t s = st->ctx->stream;
    st->retain = true;
    st->stats = keep_statistics != 0;
    st->r_off = dev_zeros(s, 8 * (st->nmajor_total + 1));
    st->nnz_hint = nnz_hint;
    SRB_API_END
}

int32_t srb_stream_finish_matrix(srb_stream *st, srb_mat **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(st && out, SRB_ERR_INVALID_ARG, "null argument");
    SRB_REQUIRE(st->retain, SRB_ERR_INVALID_ARG, "stream was not opened with srb_stream_set_retain");
    SRB_REQUIRE(st->major_pos == st->nmajor_total, SRB_ERR_INVALID_ARG, "fewer lines pushed than announced");
    auto sx = std::make_shared<Structure>();
    sx->nmajor = st->nmajor_total, sx->nminor = st->nminor, sx->nnz = st->nnz_pos;
    sx->offsets = st->r_off;
    cudaStream_t s = st->ctx->stream;
    sx->indices = st->r_idx ? st->r_idx : dev_alloc(s, 4);
    std::unique_ptr<srb_mat> m(new srb_mat());
    m->ctx = st->ctx, m->format = st->format, m->nrows = st->nrows_total, m->ncols = st->ncols_total, m->st = sx;
    m->vdtype = st->vdtype < 0 ? SRB_F32 : st->vdtype, m->src_dtype = st->src_dtype;
    m->values = st->r_val ? st->r_val : dev_alloc(s, 8);
    m->global_row0 = 0, m->global_nrows = st->nrows_total;
    st->r_off.reset(), st->r_idx.reset(), st->r_val.reset();
    st->retain = false;
    *out = m.release();
    SRB_API_END
}

int32_t srb_stream_free(srb_stream *st) {
    SRB_API_BEGIN
    delete st;
    SRB_API_END
}

}  // extern "C"
