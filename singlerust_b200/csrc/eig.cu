// eig.cu — K8: symmetric eigendecomposition of the d x d correlation matrix (d = #selected genes, 2000 in the
// headline config). This is the small dense step the reference hands to single_algebra's SVD backends
// (LAPACK/faer, Cargo.toml:13-15,42); here cuSOLVER's fp64 syevd (a library call for a latency-bound O(d^3)
// step that is not on the HBM/tensor hot path).
#include <cusolverDn.h>

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace srb {

#define SRB_CUSOLVER(expr)                                                                                   \
    do {                                                                                                     \
        cusolverStatus_t _e = (expr);                                                                        \
        if (_e != CUSOLVER_STATUS_SUCCESS)                                                                   \
            throw srb::Error(SRB_ERR_CUDA, std::string(#expr) + ": cusolver status " + std::to_string((int)_e)); \
    } while (0)

// C (row- or column-major: symmetric) is overwritten by the eigenvectors (column-major, ascending eigenvalues);
// evals receives the ascending eigenvalues.
uint32_t sym_eig_desc(srb_ctx *ctx, double *d_C, uint32_t d, uint32_t topk, double *d_evals) {
    cudaStream_t s = ctx->stream;
    if (!ctx->solver) {
        cusolverDnHandle_t h;
        SRB_CUSOLVER(cusolverDnCreate(&h));
        ctx->solver = h;
        int lo = 0, hi = 0;
        SRB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically lowest = highest priority
        SRB_CUDA(cudaStreamCreateWithPriority(&ctx->eig_stream, cudaStreamNonBlocking, hi));
        SRB_CUDA(cudaEventCreateWithFlags(&ctx->eig_in, cudaEventDisableTiming));
        SRB_CUDA(cudaEventCreateWithFlags(&ctx->eig_out, cudaEventDisableTiming));
    }
    cusolverDnHandle_t h = (cusolverDnHandle_t)ctx->solver;
    cudaStream_t es = ctx->eig_stream;
    SRB_CUSOLVER(cusolverDnSetStream(h, es));
    static int use_x = -1;  // SRB_EIG_X=1: the 64-bit generic API (cusolverDnXsyevd) instead of the legacy Dsyevd
    static int use_range = -1;  // SRB_EIG_RANGE=1: cusolverDnDsyevdx on the index range of the top-k eigenvalues only
    if (use_x < 0) {
        const char *e = getenv("SRB_EIG_X");
        use_x = (e && e[0] == '1') ? 1 : 0;
        const char *r = getenv("SRB_EIG_RANGE");
        use_range = (r && r[0] == '1') ? 1 : 0;
    }
    if (use_range && topk < d) {
        int lw = 0, meig = 0;
        const int il = (int)(d - topk + 1), iu = (int)d;
        SRB_CUSOLVER(cusolverDnDsyevdx_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, (int)d, d_C, (int)d,
                                                  0.0, 0.0, il, iu, &meig, d_evals, &lw));
        Buf wk = dev_alloc(s, sizeof(double) * (size_t)std::max(lw, 1));
        Buf inf = dev_zeros(s, sizeof(int));
        SRB_CUDA(cudaEventRecord(ctx->eig_in, s));
        SRB_CUDA(cudaStreamWaitEvent(es, ctx->eig_in, 0));
        SRB_CUSOLVER(cusolverDnDsyevdx(h, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, (int)d, d_C, (int)d, 0.0, 0.0, il,
                                       iu, &meig, d_evals, wk->as<double>(), lw, inf->as<int>()));
        int hi = 0;
        SRB_CUDA(cudaMemcpyAsync(&hi, inf->p, sizeof(int), cudaMemcpyDeviceToHost, es));
        SRB_CUDA(cudaEventRecord(ctx->eig_out, es));
        SRB_CUDA(cudaStreamWaitEvent(s, ctx->eig_out, 0));
        SRB_CUDA(cudaStreamSynchronize(es));
        SRB_REQUIRE(hi == 0 && meig == (int)topk, SRB_ERR_NAN, "syevdx failed (info=" + std::to_string(hi) + ", meig=" + std::to_string(meig) + ")");
        return topk;
    }
    Buf info = dev_zeros(s, sizeof(int));
    Buf work;
    int lwork = 0;
    size_t wdev = 0, whost = 0;
    std::vector<char> hwork;
    if (!ctx->solver_params) {
        cusolverDnParams_t prm;
        SRB_CUSOLVER(cusolverDnCreateParams(&prm));
        ctx->solver_params = prm;
    }
    if (use_x) {
        SRB_CUSOLVER(cusolverDnXsyevd_bufferSize(h, (cusolverDnParams_t)ctx->solver_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                                                 (int64_t)d, CUDA_R_64F, d_C, (int64_t)d, CUDA_R_64F, d_evals, CUDA_R_64F, &wdev, &whost));
        work = dev_alloc(s, std::max<size_t>(wdev, 16));
        hwork.resize(std::max<size_t>(whost, 16));
    } else {
        SRB_CUSOLVER(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)d, d_C, (int)d, d_evals, &lwork));
        work = dev_alloc(s, sizeof(double) * (size_t)std::max(lwork, 1));
    }
    // main stream -> eig stream (inputs ready, scratch buffers owned) ...
    SRB_CUDA(cudaEventRecord(ctx->eig_in, s));
    SRB_CUDA(cudaStreamWaitEvent(es, ctx->eig_in, 0));
    if (use_x)
        SRB_CUSOLVER(cusolverDnXsyevd(h, (cusolverDnParams_t)ctx->solver_params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int64_t)d,
                                      CUDA_R_64F, d_C, (int64_t)d, CUDA_R_64F, d_evals, CUDA_R_64F, work->p, wdev, hwork.data(), whost, info->as<int>()));
    else
        SRB_CUSOLVER(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)d, d_C, (int)d, d_evals, work->as<double>(), lwork, info->as<int>()));
    int hinfo = 0;
    SRB_CUDA(cudaMemcpyAsync(&hinfo, info->p, sizeof(int), cudaMemcpyDeviceToHost, es));
    // ... and back: everything enqueued on the main stream afterwards (including reuse of the scratch blocks through
    // the block cache) is ordered after the eigensolver
    SRB_CUDA(cudaEventRecord(ctx->eig_out, es));
    SRB_CUDA(cudaStreamWaitEvent(s, ctx->eig_out, 0));
    SRB_CUDA(cudaStreamSynchronize(es));
    SRB_REQUIRE(hinfo == 0, SRB_ERR_NAN, "syevd did not converge / illegal value (info=" + std::to_string(hinfo) + "): NaN in the correlation matrix?");
    return d;
}

void eig_destroy(srb_ctx *ctx) {
    if (ctx->solver_params) cusolverDnDestroyParams((cusolverDnParams_t)ctx->solver_params);
    ctx->solver_params = nullptr;
    if (ctx->solver) cusolverDnDestroy((cusolverDnHandle_t)ctx->solver);
    ctx->solver = nullptr;
    if (ctx->eig_stream) cudaStreamDestroy(ctx->eig_stream);
    if (ctx->eig_in) cudaEventDestroy(ctx->eig_in);
    if (ctx->eig_out) cudaEventDestroy(ctx->eig_out);
    ctx->eig_stream = nullptr, ctx->eig_in = nullptr, ctx->eig_out = nullptr;
}

}  // namespace srb
