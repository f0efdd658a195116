// eig.cu — K8: symmetric eigendecomposition of the d x d correlation matrix (d = #selected genes, 2000 in the
// headline config). This is the small dense step the reference hands to single_algebra's SVD backends
// (LAPACK/faer, Cargo.toml:13-15,42); here cuSOLVER's fp64 syevd (a library call for a latency-bound O(d^3)
// step that is not on the HBM/tensor hot path).
#include <cusolverDn.h>

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
void sym_eig_desc(srb_ctx *ctx, double *d_C, uint32_t d, double *d_evals) {
    cudaStream_t s = ctx->stream;
    if (!ctx->solver) {
        cusolverDnHandle_t h;
        SRB_CUSOLVER(cusolverDnCreate(&h));
        ctx->solver = h;
    }
    cusolverDnHandle_t h = (cusolverDnHandle_t)ctx->solver;
    SRB_CUSOLVER(cusolverDnSetStream(h, s));
    int lwork = 0;
    SRB_CUSOLVER(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)d, d_C, (int)d, d_evals, &lwork));
    Buf work = dev_alloc(s, sizeof(double) * (size_t)std::max(lwork, 1));
    Buf info = dev_zeros(s, sizeof(int));
    SRB_CUSOLVER(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)d, d_C, (int)d, d_evals, work->as<double>(), lwork, info->as<int>()));
    int hinfo = 0;
    SRB_CUDA(cudaMemcpyAsync(&hinfo, info->p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_REQUIRE(hinfo == 0, SRB_ERR_NAN, "syevd did not converge / illegal value (info=" + std::to_string(hinfo) + "): NaN in the correlation matrix?");
}

void eig_destroy(srb_ctx *ctx) {
    if (ctx->solver) cusolverDnDestroy((cusolverDnHandle_t)ctx->solver);
    ctx->solver = nullptr;
}

}  // namespace srb
