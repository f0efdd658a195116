// api.cu — C ABI entry points of libsrb200 (include/srb200.h), context / upload / download plumbing and the
// direction dispatch that mirrors src/shared/statistics/mod.rs (ArrayData::{CsrMatrix,CscMatrix} -> helper).
#include <algorithm>
#include <cmath>
#include <chrono>
#include <map>
#include <set>
#include <mutex>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace srb {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};
void set_last_error(const std::string &msg) { t_last_error = msg; }

// ---- device memory: per-stream block cache ----------------------------------------------------------------------
// Every buffer of a context lives on that context's single stream, so a block released by ~DevBuf can be handed to
// the next request in program order without any event bookkeeping. Steady-state steps therefore never reach the
// driver allocator (measured: cudaMallocAsync pool growth stalls of 40-700 ms per step otherwise). Best fit within
// 25 %; blocks beyond the cache budget (3/4 of the device) are returned to the driver.
struct BlockCache {
    std::multimap<size_t, void *> free_blocks;
    size_t cached_bytes = 0;
};
static std::mutex g_cache_mu;
static std::map<cudaStream_t, BlockCache> g_caches;
static std::set<cudaStream_t> g_live_streams;
static std::set<const srb_ctx *> g_live_ctx;

void register_stream(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_live_streams.insert(s);
}
void unregister_stream(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_live_streams.erase(s);
    g_caches.erase(s);
}
bool ctx_alive(const srb_ctx *c) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    return g_live_ctx.count(c) != 0;
}
// free blocks kept per stream: three quarters of the device (135 GB on a B200). A request the driver cannot satisfy releases
// the cache and retries (DevBuf::DevBuf), so a generous budget cannot cause an out-of-memory error; a tight one made the
// 4M-cell step (57 GB of per-step buffers on top of a 48 GB matrix) go back to cudaMalloc / cudaFree every step (353 vs 83 ms)
static size_t cache_budget() {
    static const size_t v = [] {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || total_b == 0) {
            (void)cudaGetLastError();
            return (size_t)64 << 30;
        }
        return total_b / 4 * 3;
    }();
    return v;
}

static size_t round_block(size_t n) {
    if (n < 512) return 512;
    if (n < (1u << 20)) return (n + 511) & ~size_t(511);
    return (n + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
}

DevBuf::DevBuf(size_t n, cudaStream_t s) : bytes(n), st(s) {
    cap = round_block(n);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        BlockCache &c = g_caches[s];
        auto it = c.free_blocks.lower_bound(cap);
        if (it != c.free_blocks.end() && it->first <= cap + cap / 4 + (size_t(1) << 20)) {
            p = it->second;
            cap = it->first;
            c.cached_bytes -= cap;
            c.free_blocks.erase(it);
            return;
        }
    }
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        release_all_cached_blocks();  // give every stream's free blocks back to the driver and retry once
        e = cudaMalloc(&p, cap);
    }
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        throw Error(SRB_ERR_OOM, std::string("cudaMalloc(") + std::to_string(cap) + "): " + cudaGetErrorString(e));
    }
}
DevBuf::~DevBuf() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (!g_live_streams.count(st)) {  // the owning context is gone: nothing may be queued on its stream any more
        cudaFree(p);
        return;
    }
    BlockCache &c = g_caches[st];
    c.free_blocks.emplace(cap, p);
    c.cached_bytes += cap;
    while (c.cached_bytes > cache_budget() && !c.free_blocks.empty()) {
        auto it = std::prev(c.free_blocks.end());
        cudaFree(it->second);
        c.cached_bytes -= it->first;
        c.free_blocks.erase(it);
    }
}
void release_cached_blocks(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto f = g_caches.find(s);
    if (f == g_caches.end()) return;
    cudaStreamSynchronize(s);
    for (auto &kv : f->second.free_blocks) cudaFree(kv.second);
    f->second.free_blocks.clear();
    f->second.cached_bytes = 0;
}
void release_all_cached_blocks() {
    std::vector<cudaStream_t> streams;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (auto &kv : g_caches)
            if (g_live_streams.count(kv.first)) streams.push_back(kv.first);
    }
    for (cudaStream_t st : streams) release_cached_blocks(st);
}
Buf dev_alloc(cudaStream_t st, size_t bytes) { return std::make_shared<DevBuf>(bytes, st); }
Buf dev_zeros(cudaStream_t st, size_t bytes) {
    Buf b = dev_alloc(st, bytes);
    SRB_CUDA(cudaMemsetAsync(b->p, 0, bytes ? bytes : 16, st));
    return b;
}

static bool debug_timing() {
    static int v = -1;
    if (v < 0) v = getenv("SRB_DEBUG_TIMING") ? 1 : 0;
    return v == 1;
}
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void trace_point(const char *label) {
    if (!debug_timing()) return;
    static double last = 0.0;
    const double t = now_ms();
    fprintf(stderr, "[srb]   %-28s +%.2f ms\n", label, last == 0.0 ? 0.0 : t - last);
    last = t;
}
StageTimer::StageTimer(srb_ctx *ctx, int stage) : c(ctx), s(stage) {
    if (!c->ev_used[s]) {
        cudaEventRecord(c->ev0[s], c->stream);
        c->ev_used[s] = true;
    }
    if (debug_timing()) {
        cudaStreamSynchronize(c->stream);
        t0 = now_ms();
    }
}
StageTimer::~StageTimer() {
    cudaEventRecord(c->ev1[s], c->stream);
    if (debug_timing()) {
        const double t1 = now_ms();
        cudaStreamSynchronize(c->stream);
        fprintf(stderr, "[srb] stage %d: host enqueue %.2f ms, until done %.2f ms\n", s, t1 - t0, now_ms() - t0);
    }
}

static void check_mat(const srb_mat *m) {
    SRB_REQUIRE(m && m->ctx && m->st, SRB_ERR_INVALID_ARG, "null matrix handle");
    SRB_REQUIRE(ctx_alive(m->ctx), SRB_ERR_INVALID_ARG, "the matrix outlived its context (srb_ctx_destroy was called first)");
}
static void check_dir(int d) { SRB_REQUIRE(d == SRB_ROW || d == SRB_COLUMN, SRB_ERR_INVALID_ARG, "direction must be 0 (Row) or 1 (Column)"); }

static void d2h(srb_ctx *c, void *host, const void *dev, size_t bytes) {
    if (bytes) SRB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
}

__global__ void f64_to_u32_kernel(const double *__restrict__ in, uint32_t *__restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}
__global__ void offsets_to_counts_kernel(const int64_t *__restrict__ off, uint32_t *__restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(off[i + 1] - off[i]);
}
__global__ void sqrt_kernel(double *a, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = sqrt(a[i]);
}

static unsigned nb(uint64_t n) { return (unsigned)std::max<uint64_t>(1, (n + 255) / 256); }

// number: csr.rs:16-38 / csc.rs:15-35
static void number_device(srb_mat *m, int direction, uint32_t *d_out) {
    cudaStream_t s = m->ctx->stream;
    if (m->dir_is_major(direction)) {
        const uint64_t n = m->nmajor();
        if (n) SRB_LAUNCH(offsets_to_counts_kernel, nb(n), 256, 0, s, m->st->offsets->as<int64_t>(), d_out, n);
    } else {
        ensure_minor_moments(m);
        const uint64_t n = m->nminor();
        if (n) SRB_LAUNCH(f64_to_u32_kernel, nb(n), 256, 0, s, m->minor.cnt->as<double>(), d_out, n);
    }
}
static void sum_device(srb_mat *m, int direction, double *d_out) {
    cudaStream_t s = m->ctx->stream;
    if (m->dir_is_major(direction)) {
        major_sum_absmax(m);
        SRB_CUDA(cudaMemcpyAsync(d_out, m->major.sum->p, sizeof(double) * m->nmajor(), cudaMemcpyDeviceToDevice, s));
    } else {
        ensure_minor_moments(m);
        SRB_CUDA(cudaMemcpyAsync(d_out, m->minor.sum->p, sizeof(double) * m->nminor(), cudaMemcpyDeviceToDevice, s));
    }
}
static void variance_device(srb_mat *m, int direction, double *d_out, bool sqrt_it) {
    if (m->dir_is_major(direction)) {
        major_variance(m, d_out);
        const uint64_t n = m->nmajor();
        if (sqrt_it && n) SRB_LAUNCH(sqrt_kernel, nb(n), 256, 0, m->ctx->stream, d_out, n);
    } else {
        minor_variance_from_moments(m, d_out, sqrt_it);
    }
}
static uint64_t out_len(const srb_mat *m, int direction) { return direction == SRB_ROW ? m->nrows : m->ncols; }

}  // namespace srb

using namespace srb;

extern "C" {

const char *srb_version(void) { return "srb200 0.1 (sm_100a)"; }
const char *srb_last_error_message(void) { return t_last_error.c_str(); }
uint64_t srb_kernel_launch_count(void) { return g_launches.load(); }

int32_t srb_ctx_create(int32_t device, srb_ctx **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(out, SRB_ERR_INVALID_ARG, "out is null");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw Error(SRB_ERR_CUDA, "no CUDA device available: libsrb200 has no CPU fallback");
    }
    SRB_REQUIRE(device >= 0 && device < ndev, SRB_ERR_INVALID_ARG, "device index out of range");
    SRB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SRB_CUDA(cudaGetDeviceProperties(&prop, device));
    SRB_REQUIRE(prop.major == 10, SRB_ERR_CUDA, std::string("libsrb200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
    std::unique_ptr<srb_ctx> c(new srb_ctx());
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    SRB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < ST_COUNT; ++i) {
        SRB_CUDA(cudaEventCreate(&c->ev0[i]));
        SRB_CUDA(cudaEventCreate(&c->ev1[i]));
    }
    register_stream(c->stream);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_live_ctx.insert(c.get());
    }
    *out = c.release();
    SRB_API_END
}

int32_t srb_ctx_destroy(srb_ctx *ctx) {
    SRB_API_BEGIN
    if (!ctx) return SRB_OK;
    SRB_REQUIRE(ctx_alive(ctx), SRB_ERR_INVALID_ARG, "context already destroyed");
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_live_ctx.erase(ctx);
    }
    release_cached_blocks(ctx->stream);
    unregister_stream(ctx->stream);
    comm_destroy(ctx);
    eig_destroy(ctx);
    if (ctx->up_ring) cudaFreeHost(ctx->up_ring);
    for (int i = 0; i < srb_ctx::kUpSlots; ++i)
        if (ctx->up_ev[i]) cudaEventDestroy(ctx->up_ev[i]);
    for (int i = 0; i < srb_ctx::kUpChunkEvents; ++i) {
        if (ctx->up_cev[i]) cudaEventDestroy(ctx->up_cev[i]);
        if (ctx->up_sev[i]) cudaEventDestroy(ctx->up_sev[i]);
    }
    for (int i = 0; i < ST_COUNT; ++i) {
        cudaEventDestroy(ctx->ev0[i]);
        cudaEventDestroy(ctx->ev1[i]);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    SRB_API_END
}

int32_t srb_ctx_set_value_mode(srb_ctx *ctx, int32_t mode) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    SRB_REQUIRE(mode == SRB_VALUES_COMPACT || mode == SRB_VALUES_FAITHFUL, SRB_ERR_INVALID_ARG, "bad value mode");
    ctx->value_mode = mode;
    SRB_API_END
}

int32_t srb_ctx_set_eig_mode(srb_ctx *ctx, int32_t mode) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    SRB_REQUIRE(mode == SRB_EIG_SYEVD || mode == SRB_EIG_CHFSI, SRB_ERR_INVALID_ARG, "bad eig mode");
    ctx->eig_mode = mode;
    SRB_API_END
}

int32_t srb_ctx_last_eig(srb_ctx *ctx, int32_t *solver, int32_t *block_products, int32_t *outer_iterations, double *max_residual) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    if (solver) *solver = ctx->last_eig_mode;
    if (block_products) *block_products = ctx->last_eig_products;
    if (outer_iterations) *outer_iterations = ctx->last_eig_outer;
    if (max_residual) *max_residual = ctx->last_eig_residual;
    SRB_API_END
}

int32_t srb_ctx_synchronize(srb_ctx *ctx) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    SRB_CUDA(cudaStreamSynchronize(ctx->stream));
    SRB_API_END
}

void *srb_ctx_stream(srb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int32_t srb_mat_set_shard(srb_mat *m, uint64_t global_row0, uint64_t global_nrows) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(m->format == SRB_CSR, SRB_ERR_UNSUPPORTED, "row sharding needs CSR");
    SRB_REQUIRE(global_row0 + m->nrows <= global_nrows, SRB_ERR_INVALID_ARG, "shard exceeds the global row count");
    m->global_row0 = global_row0, m->global_nrows = global_nrows;
    SRB_API_END
}

int32_t srb_mat_free(srb_mat *m) {
    SRB_API_BEGIN
    if (m) {
        if (ctx_alive(m->ctx)) cudaSetDevice(m->ctx->device);
        delete m;  // buffers of a dead context go straight back to the driver (~DevBuf)
    }
    SRB_API_END
}

int32_t srb_mat_clone(srb_mat *m, srb_mat **out) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(out, SRB_ERR_INVALID_ARG, "out is null");
    // copy-on-write: the clone shares structure and value buffer; the first transform of either side
    // writes into a fresh buffer (materialize() sees use_count > 1)
    *out = new srb_mat(*m);
    SRB_API_END
}

int32_t srb_mat_info(srb_mat *m, uint64_t *nrows, uint64_t *ncols, uint64_t *nnz, int32_t *format, int32_t *value_dtype) {
    SRB_API_BEGIN
    check_mat(m);
    if (nrows) *nrows = m->nrows;
    if (ncols) *ncols = m->ncols;
    if (nnz) *nnz = m->st->nnz;
    if (format) *format = m->format;
    if (value_dtype) {
        int dt = m->vdtype;
        if (m->has_pending() && m->ctx->value_mode == SRB_VALUES_FAITHFUL && (m->pend_scale || m->src_dtype != SRB_F32)) dt = SRB_F64;
        *value_dtype = dt;
    }
    SRB_API_END
}

int32_t srb_mat_download(srb_mat *m, uint64_t *offsets, uint64_t *indices, double *values_f64, float *values_f32) {
    SRB_API_BEGIN
    check_mat(m);
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint64_t nnz = m->st->nnz;
    if (offsets) d2h(c, offsets, m->st->offsets->p, sizeof(int64_t) * (m->nmajor() + 1));
    const uint64_t chunk = std::min<uint64_t>(std::max<uint64_t>(nnz, 1), 1ull << 26);
    if (indices && nnz) {
        Buf stage = dev_alloc(s, chunk * 8);
        for (uint64_t o = 0; o < nnz; o += chunk) {
            const uint64_t len = std::min(chunk, nnz - o);
            SRB_LAUNCH((convert_kernel<uint32_t, uint64_t>), grid_for(c, len), 256, 0, s, m->st->indices->as<uint32_t>() + o, stage->as<uint64_t>(), len);
            d2h(c, indices + o, stage->p, len * 8);
            SRB_CUDA(cudaStreamSynchronize(s));
        }
    }
    if ((values_f64 || values_f32) && nnz) {
        if (m->has_pending()) materialize(m, false);
        if (values_f64) {
            if (m->vdtype == SRB_F64) d2h(c, values_f64, m->values->p, nnz * 8);
            else {
                Buf stage = dev_alloc(s, chunk * 8);
                for (uint64_t o = 0; o < nnz; o += chunk) {
                    const uint64_t len = std::min(chunk, nnz - o);
                    SRB_LAUNCH((convert_kernel<float, double>), grid_for(c, len), 256, 0, s, m->values->as<float>() + o, stage->as<double>(), len);
                    d2h(c, values_f64 + o, stage->p, len * 8);
                    SRB_CUDA(cudaStreamSynchronize(s));
                }
            }
        }
        if (values_f32) {
            if (m->vdtype == SRB_F32) d2h(c, values_f32, m->values->p, nnz * 4);
            else {
                Buf stage = dev_alloc(s, chunk * 4);
                for (uint64_t o = 0; o < nnz; o += chunk) {
                    const uint64_t len = std::min(chunk, nnz - o);
                    SRB_LAUNCH((convert_kernel<double, float>), grid_for(c, len), 256, 0, s, m->values->as<double>() + o, stage->as<float>(), len);
                    d2h(c, values_f32 + o, stage->p, len * 4);
                    SRB_CUDA(cudaStreamSynchronize(s));
                }
            }
        }
    }
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_API_END
}

// ---- statistics --------------------------------------------------------------------------------------
int32_t srb_number(srb_mat *m, int32_t direction, uint32_t *out) {
    SRB_API_BEGIN
    check_mat(m), check_dir(direction);
    SRB_REQUIRE(out, SRB_ERR_INVALID_ARG, "out is null");
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    const uint64_t n = out_len(m, direction);
    Buf d = dev_alloc(m->ctx->stream, sizeof(uint32_t) * (n ? n : 1));
    number_device(m, direction, d->as<uint32_t>());
    d2h(m->ctx, out, d->p, sizeof(uint32_t) * n);
    SRB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    SRB_API_END
}

int32_t srb_sum(srb_mat *m, int32_t direction, double *out) {
    SRB_API_BEGIN
    check_mat(m), check_dir(direction);
    SRB_REQUIRE(out, SRB_ERR_INVALID_ARG, "out is null");
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    const uint64_t n = out_len(m, direction);
    Buf d = dev_alloc(m->ctx->stream, sizeof(double) * (n ? n : 1));
    sum_device(m, direction, d->as<double>());
    d2h(m->ctx, out, d->p, sizeof(double) * n);
    SRB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    SRB_API_END
}

static int32_t variance_like(srb_mat *m, int32_t direction, double *out, bool sq) {
    SRB_API_BEGIN
    check_mat(m), check_dir(direction);
    SRB_REQUIRE(out, SRB_ERR_INVALID_ARG, "out is null");
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    const uint64_t n = out_len(m, direction);
    Buf d = dev_alloc(m->ctx->stream, sizeof(double) * (n ? n : 1));
    variance_device(m, direction, d->as<double>(), sq);
    d2h(m->ctx, out, d->p, sizeof(double) * n);
    SRB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    SRB_API_END
}
int32_t srb_variance(srb_mat *m, int32_t direction, double *out) { return variance_like(m, direction, out, false); }
int32_t srb_std_dev(srb_mat *m, int32_t direction, double *out) { return variance_like(m, direction, out, true); }

int32_t srb_min_max(srb_mat *m, int32_t direction, double *out_min, double *out_max) {
    SRB_API_BEGIN
    check_mat(m), check_dir(direction);
    SRB_REQUIRE(out_min && out_max, SRB_ERR_INVALID_ARG, "out is null");
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    const uint64_t n = out_len(m, direction);
    Buf d = dev_alloc(m->ctx->stream, sizeof(double) * 2 * (n ? n : 1));
    if (m->dir_is_major(direction)) major_min_max(m, d->as<double>(), d->as<double>() + n);
    else minor_min_max(m, d->as<double>(), d->as<double>() + n);
    d2h(m->ctx, out_min, d->p, sizeof(double) * n);
    d2h(m->ctx, out_max, d->as<double>() + n, sizeof(double) * n);
    SRB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    SRB_API_END
}

int32_t srb_qc_all(srb_mat *m, uint32_t *num_per_cell, uint32_t *num_per_gene, double *expr_per_cell,
                   double *expr_per_gene, double *variance_per_cell, double *variance_per_gene,
                   double *std_dev_per_cell, double *std_dev_per_gene) {
    SRB_API_BEGIN
    check_mat(m);
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint64_t nr = m->nrows, ncol = m->ncols;
    Buf cell_u = dev_alloc(s, 4 * (nr + 1)), gene_u = dev_alloc(s, 4 * (ncol + 1));
    Buf cell_d = dev_alloc(s, 8 * (nr + 1)), gene_d = dev_alloc(s, 8 * (ncol + 1));
    if (num_per_cell) { number_device(m, SRB_ROW, cell_u->as<uint32_t>()); d2h(c, num_per_cell, cell_u->p, 4 * nr); }
    if (num_per_gene) { number_device(m, SRB_COLUMN, gene_u->as<uint32_t>()); d2h(c, num_per_gene, gene_u->p, 4 * ncol); }
    if (expr_per_cell) { sum_device(m, SRB_ROW, cell_d->as<double>()); d2h(c, expr_per_cell, cell_d->p, 8 * nr); SRB_CUDA(cudaStreamSynchronize(s)); }
    if (expr_per_gene) { sum_device(m, SRB_COLUMN, gene_d->as<double>()); d2h(c, expr_per_gene, gene_d->p, 8 * ncol); SRB_CUDA(cudaStreamSynchronize(s)); }
    if (variance_per_cell || std_dev_per_cell) {
        variance_device(m, SRB_ROW, cell_d->as<double>(), false);
        if (variance_per_cell) d2h(c, variance_per_cell, cell_d->p, 8 * nr);
        if (std_dev_per_cell) {
            if (nr) SRB_LAUNCH(sqrt_kernel, nb(nr), 256, 0, s, cell_d->as<double>(), nr);
            d2h(c, std_dev_per_cell, cell_d->p, 8 * nr);
        }
        SRB_CUDA(cudaStreamSynchronize(s));
    }
    if (variance_per_gene || std_dev_per_gene) {
        variance_device(m, SRB_COLUMN, gene_d->as<double>(), false);
        if (variance_per_gene) d2h(c, variance_per_gene, gene_d->p, 8 * ncol);
        if (std_dev_per_gene) {
            if (ncol) SRB_LAUNCH(sqrt_kernel, nb(ncol), 256, 0, s, gene_d->as<double>(), ncol);
            d2h(c, std_dev_per_gene, gene_d->p, 8 * ncol);
        }
    }
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_API_END
}

// per-gene moments of the CURRENT values of a CSR row chunk (pending transforms applied): count, sum, sum of squares.
// Chunk-local: never reduced over ranks (pass 1 of the out-of-core pipeline adds them up itself).
int32_t srb_gene_moments(srb_mat *m, double *count, double *sum, double *sumsq) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(m->format == SRB_CSR, SRB_ERR_UNSUPPORTED, "gene moments of a chunk need CSR (genes = minor axis)");
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    ensure_minor_moments(m, /*local_only=*/true);
    const size_t bytes = sizeof(double) * m->ncols;
    if (count) d2h(c, count, m->minor.cnt->p, bytes);
    if (sum) d2h(c, sum, m->minor.sum->p, bytes);
    if (sumsq) d2h(c, sumsq, m->minor.sq->p, bytes);
    SRB_CUDA(cudaStreamSynchronize(c->stream));
    SRB_API_END
}

// ---- normalisation / transform -------------------------------------------------------------------------
int32_t srb_normalize_total_inplace(srb_mat *m, double target_sum, int32_t direction) {
    SRB_API_BEGIN
    check_mat(m), check_dir(direction);
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    set_pending_normalize(m, target_sum, direction);
    SRB_API_END
}

int32_t srb_log1p_inplace(srb_mat *m) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_CUDA(cudaSetDevice(m->ctx->device));
    set_pending_log1p(m);
    SRB_API_END
}

// ---- feature selection ------------------------------------------------------------------------------
int32_t srb_select_hvg(srb_mat *m, uint64_t n_top, uint64_t *out_idx, uint64_t *out_n) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(out_idx && out_n, SRB_ERR_INVALID_ARG, "out is null");
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    Buf d_idx;
    uint64_t n = 0;
    select_hvg_device(m, n_top, d_idx, &n, true);
    std::vector<uint32_t> h(n);
    d2h(c, h.data(), d_idx->p, 4 * n);
    SRB_CUDA(cudaStreamSynchronize(c->stream));
    for (uint64_t i = 0; i < n; ++i) out_idx[i] = h[i];
    *out_n = n;
    SRB_API_END
}

int32_t srb_select_var_threshold(srb_mat *m, double threshold, uint64_t *out_idx, uint64_t *out_n) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(out_idx && out_n, SRB_ERR_INVALID_ARG, "out is null");
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    const uint64_t ncol = m->ncols;
    Buf d = dev_alloc(c->stream, 8 * (ncol + 1));
    variance_device(m, SRB_COLUMN, d->as<double>(), false);
    std::vector<double> v(ncol);
    d2h(c, v.data(), d->p, 8 * ncol);
    SRB_CUDA(cudaStreamSynchronize(c->stream));
    uint64_t n = 0;
    for (uint64_t j = 0; j < ncol; ++j)
        if (v[j] > threshold) out_idx[n++] = j;  // dim_red/mod.rs:148-153
    *out_n = n;
    SRB_API_END
}

// ---- densify / PCA --------------------------------------------------------------------------------------
static Buf upload_selection(srb_mat *m, const uint64_t *col_sel, uint64_t n_sel) {
    SRB_REQUIRE(col_sel || n_sel == 0, SRB_ERR_INVALID_ARG, "col_sel is null");
    std::vector<uint32_t> h(n_sel);
    for (uint64_t j = 0; j < n_sel; ++j) {
        // select_info_elem_to_indices, shared/utils/mod.rs:8-13: "Index out of bounds"
        SRB_REQUIRE(col_sel[j] < m->ncols, SRB_ERR_INDEX_OOB, "selected column index out of bounds");
        h[j] = (uint32_t)col_sel[j];
    }
    Buf d = dev_alloc(m->ctx->stream, 4 * (n_sel ? n_sel : 1));
    if (n_sel) SRB_CUDA(cudaMemcpyAsync(d->p, h.data(), 4 * n_sel, cudaMemcpyHostToDevice, m->ctx->stream));
    SRB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return d;
}

int32_t srb_densify_selected(srb_mat *m, const uint64_t *col_sel, uint64_t n_sel, double *out) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(out || n_sel == 0 || m->nrows == 0, SRB_ERR_INVALID_ARG, "out is null");
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    std::unique_ptr<srb_mat> twin;  // convert_to_array_f64_csc_selected (shared/mod.rs:261-290): via a CSR twin
    if (m->format == SRB_CSC) {
        twin.reset(csc_to_csr(m));
        m = twin.get();
    }
    Buf d_sel = upload_selection(m, col_sel, n_sel);
    if (n_sel == 0 || m->nrows == 0) return SRB_OK;
    const uint64_t rows_per = std::max<uint64_t>(1, std::min<uint64_t>(m->nrows, (1ull << 27) / n_sel));
    Buf stage = dev_alloc(c->stream, 8 * rows_per * n_sel);
    for (uint64_t r0 = 0; r0 < m->nrows; r0 += rows_per) {
        const uint64_t nr = std::min(rows_per, m->nrows - r0);
        densify_selected_f64(m, d_sel->as<uint32_t>(), n_sel, stage->as<double>(), r0, nr);
        d2h(c, out + r0 * n_sel, stage->p, 8 * nr * n_sel);
        SRB_CUDA(cudaStreamSynchronize(c->stream));
    }
    SRB_API_END
}

static void reset_stage_timers(srb_ctx *c) {
    for (int i = 0; i < ST_COUNT; ++i) c->ev_used[i] = false;
}

int32_t srb_pca(srb_mat *m, const uint64_t *col_sel, uint64_t n_sel, uint64_t k, int32_t center, int32_t scale,
                int32_t gram_mode, double *scores, double *components, double *explained_variance_ratio) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(n_sel >= 1 && k >= 1, SRB_ERR_INVALID_ARG, "need at least one feature and one component");
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    reset_stage_timers(c);
    std::unique_ptr<srb_mat> twin;  // CSC-stored X: PCA runs on a CSR twin (cells = rows)
    if (m->format == SRB_CSC) {
        twin.reset(csc_to_csr(m));
        m = twin.get();
    }
    Buf d_sel = upload_selection(m, col_sel, n_sel);
    PcaOut o{scores, components, explained_variance_ratio};
    pca_run(m, d_sel->as<uint32_t>(), n_sel, std::min<uint64_t>(k, n_sel), center != 0, scale != 0, gram_mode, o);
    SRB_API_END
}

int32_t srb_pipeline_normalize_hvg_pca(srb_mat *m, double target_sum, uint64_t n_top, uint64_t k, int32_t center,
                                       int32_t scale, int32_t gram_mode, uint64_t *hvg_out, double *scores,
                                       double *components, double *explained_variance_ratio) {
    SRB_API_BEGIN
    check_mat(m);
    SRB_REQUIRE(n_top >= 1 && k >= 1, SRB_ERR_INVALID_ARG, "need at least one feature and one component");
    srb_ctx *c = m->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    reset_stage_timers(c);
    SRB_TRACE("pipeline begin");
    set_pending_normalize(m, target_sum, SRB_ROW);
    SRB_TRACE("set_pending_normalize");
    set_pending_log1p(m);
    SRB_TRACE("set_pending_log1p");
    Buf d_idx;
    uint64_t n_sel = 0;
    {
        select_hvg_device(m, n_top, d_idx, &n_sel, true);
    }
    SRB_TRACE("select_hvg_device");
    if (hvg_out) {
        std::vector<uint32_t> h(n_sel);
        d2h(c, h.data(), d_idx->p, 4 * n_sel);
        SRB_CUDA(cudaStreamSynchronize(c->stream));
        for (uint64_t i = 0; i < n_sel; ++i) hvg_out[i] = h[i];
    }
    PcaOut o{scores, components, explained_variance_ratio};
    std::unique_ptr<srb_mat> twin;
    if (m->format == SRB_CSC) {
        twin.reset(csc_to_csr(m));
        m = twin.get();
    }
    pca_run(m, d_idx->as<uint32_t>(), n_sel, std::min<uint64_t>(k, n_sel), center != 0, scale != 0, gram_mode, o);
    SRB_API_END
}

int32_t srb_last_stage_ms(srb_ctx *ctx, float *out_ms, int32_t n) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx && out_ms, SRB_ERR_INVALID_ARG, "null argument");
    SRB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i) {
        out_ms[i] = 0.f;
        if (i < ST_COUNT && ctx->ev_used[i]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev0[i], ctx->ev1[i]) == cudaSuccess) out_ms[i] = ms;
            else cudaGetLastError();
        }
    }
    SRB_API_END
}

}  // extern "C"
