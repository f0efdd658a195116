// common.cuh — internal types of libsrb200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/srb200.h"

namespace srb {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

void set_last_error(const std::string &msg);
extern std::atomic<uint64_t> g_launches;

#define SRB_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) {                                                                             \
            throw srb::Error(_e == cudaErrorMemoryAllocation ? SRB_ERR_OOM : SRB_ERR_CUDA,                   \
                             std::string(#expr) + ": " + cudaGetErrorString(_e));                            \
        }                                                                                                    \
    } while (0)

#define SRB_REQUIRE(cond, code, msg)                                                                         \
    do {                                                                                                     \
        if (!(cond)) throw srb::Error((code), (msg));                                                        \
    } while (0)

// every kernel launch goes through this so gpu_launches is an honest count
#define SRB_LAUNCH(kernel, grid, block, smem, stream, ...)                                                   \
    do {                                                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                          \
        srb::g_launches.fetch_add(1, std::memory_order_relaxed);                                             \
        SRB_CUDA(cudaGetLastError());                                                                        \
    } while (0)

// ABI boundary: translate exceptions into status codes + thread-local message. A non-sticky error left behind by an
// earlier call of this thread (ours or another library's) is dropped first, so SRB_LAUNCH never reports a stale one.
#define SRB_API_BEGIN                                                                                        \
    try {                                                                                                    \
        (void)cudaGetLastError();
#define SRB_API_END                                                                                          \
    return SRB_OK;                                                                                           \
    }                                                                                                        \
    catch (const srb::Error &e) {                                                                            \
        srb::set_last_error(e.what());                                                                       \
        return e.code;                                                                                       \
    }                                                                                                        \
    catch (const std::bad_alloc &) {                                                                         \
        srb::set_last_error("host allocation failed");                                                       \
        return SRB_ERR_OOM;                                                                                  \
    }                                                                                                        \
    catch (const std::exception &e) {                                                                        \
        srb::set_last_error(e.what());                                                                       \
        return SRB_ERR_INVALID_ARG;                                                                          \
    }

// Handles (matrices, streams) may outlive their context in a garbage-collected host language: contexts and their
// streams are registered while alive, and a handle freed after its context only gives its memory back to the driver.
void register_stream(cudaStream_t s);
void unregister_stream(cudaStream_t s);
bool ctx_alive(const srb_ctx *c);
// device buffer owned by one context stream; returned to that stream's block cache when the last reference drops
void release_cached_blocks(cudaStream_t s);
void release_all_cached_blocks();
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0, cap = 0;
    cudaStream_t st = nullptr;
    DevBuf(size_t n, cudaStream_t s);
    ~DevBuf();
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};
using Buf = std::shared_ptr<DevBuf>;
Buf dev_alloc(cudaStream_t st, size_t bytes);
Buf dev_zeros(cudaStream_t st, size_t bytes);

enum Stage { ST_ROWSUM = 0, ST_FUSED = 1, ST_HVG = 2, ST_DENSIFY = 3, ST_GRAM = 4, ST_EIG = 5, ST_SCORES = 6, ST_ALLREDUCE = 7, ST_ALLREDUCE_GRAM = 8, ST_COUNT = 9 };

struct NcclApi;  // dlopen'ed table (comm.cu)

}  // namespace srb

struct srb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int value_mode = SRB_VALUES_COMPACT;
    int sm_count = 148;
    size_t smem_optin = 0;
    // NCCL (optional)
    void *comm = nullptr;
    int rank = 0, nranks = 1;
    // stage timing
    cudaEvent_t ev0[srb::ST_COUNT] = {}, ev1[srb::ST_COUNT] = {};
    bool ev_used[srb::ST_COUNT] = {};
    // cuSOLVER handle (lazy, eig.cu) and the high-priority side stream its small latency-bound kernels run on, so
    // that they are scheduled ahead of another context's bandwidth-bound kernels when batches are pipelined
    void *solver = nullptr;
    void *blas = nullptr;  // cublasHandle_t of the ChFSI eigensolver (lazy, eig.cu)
    int eig_mode = -1;     // srb_eig_mode, -1 = the process default (SRB_EIG_MODE)
    int last_eig_mode = 0, last_eig_products = 0, last_eig_outer = 0;  // what the last K8 did: 0 syevd, 1 chfsi, 2 chfsi fell back
    double last_eig_residual = 0.0;
    void *solver_params = nullptr;
    cudaStream_t eig_stream = nullptr;
    cudaEvent_t eig_in = nullptr, eig_out = nullptr;
    // pinned staging ring of the packed upload path (upload.cu: upload_packed), lazily allocated
    static constexpr int kUpSlots = 4;
    int upload_mode = -1;  // srb_upload_mode, -1 = the process default (SRB_UPLOAD_PACK)
    uint64_t last_upload_h2d = 0;  // bytes the last srb_mat_upload moved over the link
    int last_upload_packed = 0;
    void *up_ring = nullptr;
    size_t up_ring_bytes = 0;
    cudaEvent_t up_ev[kUpSlots] = {};
    bool up_ev_used[kUpSlots] = {};
    // per-chunk completion events of the balanced upload (upload.cu: upload_balanced): how far the link is behind the host
    static constexpr int kUpChunkEvents = 16;
    cudaEvent_t up_cev[kUpChunkEvents] = {}, up_sev[kUpChunkEvents] = {};  // end / start of a chunk's copies
    int last_upload_chunks = 0, last_upload_idx_packed = 0, last_upload_val_packed = 0;  // what the balanced upload chose
};

namespace srb {

// immutable index structure, shared between clones of a matrix
struct Structure {
    uint64_t nmajor = 0, nminor = 0, nnz = 0;
    Buf offsets;  // int64[nmajor+1]
    Buf indices;  // uint32[nnz]
    // stripe split cache for the minor-moments kernel: splits[(S-1) * nmajor] (int64), for S = nstripes
    int nstripes = 0;
    Buf splits;
};

// exact integer per-minor-line moments (see minor_moments.cu): count, sum of q, sum of q^2 as 3x32-bit limbs
struct MinorMoments {
    bool valid = false;
    bool exact_path = false;  // fixed-point SMEM path (true) or fp64 global-atomic path (false)
    Buf acc;                  // u64[6 * nminor]: cnt, sumA, sumB, sqA, sqB, sqC (32-bit limbs; after allreduce = global)
    Buf fexp;                 // int32[1]: F (power-of-two scale), device-resident
    Buf cnt, sum, sq;         // finalised f64[nminor] each (counts are exact integers in f64)
    bool reduced = false;     // allreduced over ranks
};

struct MajorStats {
    bool valid = false;
    Buf sum;     // f64[nmajor]
    Buf absmax;  // f64[nmajor]
    Buf absmin;  // f64[nmajor]: min |v| over the non-zero stored values of the line (+inf if none)
    Buf flags;   // u32[3]: [0] any negative, [1] any non-finite, [2] any non-integer
};

}  // namespace srb

struct srb_mat {
    srb_ctx *ctx = nullptr;
    int format = SRB_CSR;
    uint64_t nrows = 0, ncols = 0;
    std::shared_ptr<srb::Structure> st;
    srb::Buf values;
    int vdtype = SRB_F32;  // device storage of `values`
    int src_dtype = SRB_F32;  // dtype the host handed over (decides FAITHFUL promotion)
    // deferred transforms (applied in this order): v * scale[line] then log1p
    srb::Buf pend_scale;  // f64[nmajor] or f64[nminor]
    bool pend_scale_major = true;
    bool pend_log1p = false;
    srb::Buf pend_bound;  // f64[2] device: {upper bound of |value|, lower bound of the non-zero |value|} after scaling (before log1p)
    // caches of the CURRENT (post-pending) logical values; dropped on mutation
    srb::MajorStats major;
    srb::MinorMoments minor;
    srb::Buf absmax_all;  // f64[2] device: {max |v|, min non-zero |v|} of the materialised values (valid iff non-null)
    // sharding
    uint64_t global_row0 = 0, global_nrows = 0;

    bool has_pending() const { return (bool)pend_scale || pend_log1p; }
    uint64_t nmajor() const { return st->nmajor; }
    uint64_t nminor() const { return st->nminor; }
    bool dir_is_major(int direction) const { return (direction == SRB_ROW) == (format == SRB_CSR); }
};

namespace srb {

// debug tracing of host-side time (SRB_DEBUG_TIMING=1)
void trace_point(const char *label);
#define SRB_TRACE(label) srb::trace_point(label)

struct StageTimer {
    srb_ctx *c;
    int s;
    double t0 = 0.0;
    StageTimer(srb_ctx *ctx, int stage);
    ~StageTimer();
};

// ---- kernels / host launchers implemented across the .cu files -------------------------------------
// major_stats.cu
void major_sum_absmax(srb_mat *m);                               // fills m->major (on materialised values)
void major_variance(srb_mat *m, double *d_out);                  // two-pass per line, NaN for empty
void major_min_max(srb_mat *m, double *d_min, double *d_max);
// minor_moments.cu
void materialize(srb_mat *m, bool want_moments, bool local_only = false);  // apply pending transforms (fused with moments)
void ensure_minor_moments(srb_mat *m, bool local_only = false);  // moments of current logical values (local_only: never reduced over ranks)
void minor_min_max(srb_mat *m, double *d_min, double *d_max);
void minor_variance_from_moments(srb_mat *m, double *d_out, bool sqrt_it);
void set_pending_normalize(srb_mat *m, double target, int direction);
void set_pending_log1p(srb_mat *m);
void reduce_minor_moments_over_ranks(srb_mat *m);
// select.cu
void select_hvg_device(srb_mat *m, uint64_t n_top, Buf &d_idx_u32, uint64_t *n_out, bool check_nan);
// pca.cu
struct PcaOut {
    double *scores, *components, *evr;
};
void pca_run(srb_mat *m, const uint32_t *d_sel, uint64_t n_sel, uint64_t k, bool center, bool scale, int gram_mode,
             PcaOut out);
void densify_selected_f64(srb_mat *m, const uint32_t *d_sel, uint64_t n_sel, double *d_out, uint64_t row0, uint64_t nrows);
// gram_tc.cu (tcgen05)
// d_used: the leading columns of the panels that hold selected genes (0 = all dpad); the rest is zero padding
void gram_tcgen05(srb_ctx *ctx, const __half *Xh, const __half *Xl, uint64_t n, uint32_t dpad, double *G, uint32_t d_used = 0);
// scores = Z' W - 1 bias^T (bias[kpad]: the rank-one correction of the sparse panel shift, zeros when every column is centred)
void scores_tcgen05(srb_ctx *ctx, const __half *Xh, const __half *Xl, uint64_t n, uint32_t dpad, const double *W,
                    uint32_t kpad, uint32_t k, const double *bias, double *scores);
// convert.cu
srb_mat *csc_to_csr(srb_mat *m);
// eig.cu
// C overwritten by eigenvectors (column-major, ascending); returns the number of eigenpairs computed (d, or topk when the
// range solver is selected): pair j (ascending) is column j / d_evals[j], so component c is column (count-1-c)
uint32_t sym_eig_desc(srb_ctx *ctx, double *d_C, uint32_t d, uint32_t topk, double *d_evals);
// comm.cu
void allreduce_u64_sum(srb_ctx *ctx, uint64_t *d_buf, size_t n);
void allreduce_f64_sum(srb_ctx *ctx, double *d_buf, size_t n);
void allreduce_f64_min(srb_ctx *ctx, double *d_buf, size_t n);
void allreduce_f64_max(srb_ctx *ctx, double *d_buf, size_t n);
void allgather_f64(srb_ctx *ctx, cudaStream_t stream, double *d_buf, size_t count_per_rank);
void comm_destroy(srb_ctx *ctx);
void eig_destroy(srb_ctx *ctx);

// element-wise dtype conversion (upload.cu: host dtype -> device storage; api.cu: device storage -> download dtype)
template <typename SRC, typename DST>
__global__ void convert_kernel(const SRC *__restrict__ src, DST *__restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = (DST)src[i];
}
// grid of 256-thread blocks for a grid-stride loop over n elements, capped at 16 blocks per SM
static inline unsigned grid_for(const srb_ctx *c, uint64_t n) {
    const uint64_t blocks = (n + 255) / 256, cap = (uint64_t)c->sm_count * 16;
    return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// misc device helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// order-preserving map double <-> u64 (for atomicMin/Max on doubles)
__device__ __forceinline__ unsigned long long f64_to_ordered(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ULL) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ordered_to_f64(unsigned long long u) {
    u = (u & 0x8000000000000000ULL) ? (u & 0x7FFFFFFFFFFFFFFFULL) : ~u;
    return __longlong_as_double((long long)u);
}

}  // namespace srb
