// eig.cu — K8: symmetric eigendecomposition of the d x d correlation matrix (d = #selected genes, 2000 in the
// headline config). This is the small dense step the reference hands to single_algebra's SVD backends
// (LAPACK/faer, Cargo.toml:13-15,42); here cuSOLVER's fp64 syevd (a library call for a latency-bound O(d^3)
// step that is not on the HBM/tensor hot path).
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

// built-in default of SRB_EIG_MODE: 0 = syevd, 1 = chfsi (measured at the bench size: 27.4 -> 11.0 ms, loadings equal to 7e-13)
#define SRB_EIG_DEFAULT_MODE 1

namespace srb {

#define SRB_CUSOLVER(expr)                                                                                   \
    do {                                                                                                     \
        cusolverStatus_t _e = (expr);                                                                        \
        if (_e != CUSOLVER_STATUS_SUCCESS)                                                                   \
            throw srb::Error(SRB_ERR_CUDA, std::string(#expr) + ": cusolver status " + std::to_string((int)_e)); \
    } while (0)

#define SRB_CUBLAS(expr)                                                                                     \
    do {                                                                                                     \
        cublasStatus_t _e = (expr);                                                                          \
        if (_e != CUBLAS_STATUS_SUCCESS)                                                                     \
            throw srb::Error(SRB_ERR_CUDA, std::string(#expr) + ": cublas status " + std::to_string((int)_e)); \
    } while (0)

// =====================================================================================================================
// Top-k eigenpairs by Chebyshev-filtered subspace iteration (ChFSI) — the GEMM-shaped alternative to syevd.
//
// PCA needs the k (= 50) leading eigenpairs of the d x d (d = 2000) correlation matrix; the explained-variance ratio
// needs only trace(C) beyond that. cuSOLVER's syevd spends 20 of its 28 ms in the Householder tridiagonalisation, one
// latency-bound column at a time (~10 us per column, fp32 no faster: measured 24.5 vs 27.7 ms), while an fp64 GEMM of
// the same size takes 0.49 ms on this part (33 TFLOP/s). So:
//   1. bounds: L = 40 Krylov steps (CGS2-orthogonalised) -> Ritz values of the L x L projection give a safe lower bound
//      of the spectrum, an upper bound, and a coarse density of states from which the first cut is read;
//   2. filter: a block Y (d x b, b ~ 4k) is multiplied by a degree-m Chebyshev polynomial of C that is bounded by 1 on
//      [lo, cut] and grows like cosh(m acosh x) above it — one DGEMM per degree (three-term recurrence on C - cI);
//      m is capped so the top of the spectrum is amplified by <= 1e8 and the block stays numerically full rank, and the
//      filter is repeated R times with a CholeskyQR2 in between until the k-th eigenvalue has gained ~1e11 on the cut;
//   3. Rayleigh-Ritz on the block (b x b syevd), residuals of the top k; repeat from 2 with the Ritz values as the new
//      cut / bounds until max ||C v - theta v|| <= 1e-11 |theta_1|.
// Any doubt (Cholesky breakdown, non-finite numbers, no convergence in 6 rounds, Krylov breakdown) returns false and the
// caller runs syevd on the untouched matrix. NumPy prototype of the same flow: 49 block products for the bench's flat
// Marchenko-Pastur spectrum, 20-110 for spiked / power-law / clustered spectra, eigenvectors within 1e-9.
// =====================================================================================================================
__global__ void fill_random_kernel(double *__restrict__ a, uint64_t n, uint64_t seed) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;  // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        a[i] = (double)(int64_t)z * (1.0 / 9223372036854775808.0);  // uniform in (-1, 1)
    }
}
// dst = src - c I  (both d x d, column-major)
__global__ void shift_copy_kernel(const double *__restrict__ src, double *__restrict__ dst, uint32_t d, double c) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)d * d) return;
    const uint32_t r = (uint32_t)(i % d), col = (uint32_t)(i / d);
    dst[i] = src[i] - (r == col ? c : 0.0);
}
__global__ void symmetrize_kernel(double *__restrict__ g, uint32_t b) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b * b) return;
    const uint32_t r = i % b, c = i / b;
    if (r < c) {
        const double m = 0.5 * (g[(size_t)c * b + r] + g[(size_t)r * b + c]);
        g[(size_t)c * b + r] = m;
        g[(size_t)r * b + c] = m;
    }
}
// one block per wanted pair j: out[j] = || cw[:, j] - theta[j0 + j] * y[:, j0 + j] ||_2
__global__ void residual_norms_kernel(const double *__restrict__ cw, const double *__restrict__ y, const double *__restrict__ theta,
                                      uint32_t d, uint32_t j0, double *__restrict__ out) {
    const uint32_t j = blockIdx.x;
    const double th = theta[j0 + j];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {
        const double r = cw[(size_t)j * d + i] - th * y[(size_t)(j0 + j) * d + i];
        acc += r * r;
    }
    __shared__ double sh[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[j] = sqrt(v);
    }
}

// ---- Krylov bounds: hand-written Lanczos step (replaces ~7 cuBLAS launches per step) -------------------------------
// w = C v for a SYMMETRIC column-major C: w[i] = <C[:, i], v>, so every warp streams one contiguous column (coalesced
// 256-byte warp loads, C stays in L2 between steps: 32 MB at d = 2000). The product is written twice: into the next basis
// slot (to be orthogonalised) and into CV[:, j] (kept for the projected matrix H = V^T C V).
__global__ void __launch_bounds__(256) symv_cols_kernel(const double *__restrict__ C, const double *__restrict__ v, uint32_t d,
                                                        double *__restrict__ w, double *__restrict__ cv) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = gridDim.x * 8;
    for (uint32_t col = warp; col < d; col += nwarps) {
        const double *c = C + (size_t)col * d;
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t i = lane;
        for (; i + 224 < d; i += 256) {  // 8 independent 256-byte warp loads in flight
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] += c[i + 32 * u] * v[i + 32 * u];
        }
        for (; i < d; i += 32) a[0] += c[i] * v[i];
        const double acc = warp_sum(((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])));
        if (lane == 0) w[col] = acc, cv[col] = acc;
    }
}
__device__ __forceinline__ double block_sum_1024(double v, double *sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x & 31) < (blockDim.x >> 5) ? sh[threadIdx.x & 31] : 0.0;
    return warp_sum(t);
}
// v0 = normalised splitmix64 noise (one CTA)
__global__ void __launch_bounds__(1024) lanczos_init_kernel(double *__restrict__ v, uint32_t d, uint64_t seed) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {
        uint64_t z = ((uint64_t)i + 1) * 0x9E3779B97F4A7C15ull + seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const double x = (double)(int64_t)z * (1.0 / 9223372036854775808.0);
        v[i] = x;
        acc += x * x;
    }
    const double inv = rsqrt(block_sum_1024(acc, sh));
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) v[i] *= inv;
}
// One CTA: w = V[:, j+1] (holding C v_j) is orthogonalised twice against V[:, 0..j] (CGS2), beta_j = ||w|| goes to
// sc[2 + j], v_{j+1} = w / beta_j. sc[1] = ||C v_0|| is the reference magnitude of the breakdown test (flag[0]).
// The coefficients h = V^T w are formed 8 basis columns at a time with every thread summing over its own elements
// (8 x elements-per-thread independent loads in flight; a first version with one warp per column was latency-bound:
// 45 us per step). Dynamic shared memory: w (d doubles) + h (j + 1) + the per-warp partial sums (32 x 8).
__global__ void __launch_bounds__(1024) lanczos_orth_kernel(double *__restrict__ V, uint32_t d, int j, double *__restrict__ sc,
                                                            uint32_t *__restrict__ flag) {
    extern __shared__ double lz[];
    __shared__ double sh[32];
    double *w = lz, *h = lz + d, *part = h + (j + 1);  // part[8][32]
    double *wg = V + (size_t)(j + 1) * d;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double nrm0 = 0.0;
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {
        const double x = wg[i];
        w[i] = x;
        nrm0 += x * x;
    }
    nrm0 = block_sum_1024(nrm0, sh);
    if (j == 0 && threadIdx.x == 0) sc[1] = sqrt(nrm0);
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
        for (int c0 = 0; c0 <= j; c0 += 8) {
            double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {
                const double wi = w[i];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c0 + u <= j) a[u] += V[(size_t)(c0 + u) * d + i] * wi;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double r = warp_sum(a[u]);
                if (lane == 0) part[u * 32 + warp] = r;
            }
            __syncthreads();
            if (warp < 8 && c0 + (int)warp <= j) {
                const double r = warp_sum(part[warp * 32 + lane]);
                if (lane == 0) h[c0 + warp] = r;
            }
            __syncthreads();
        }
        for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {  // w -= V h
            double x = w[i];
            for (int c = 0; c <= j; ++c) x -= V[(size_t)c * d + i] * h[c];
            w[i] = x;
        }
        __syncthreads();
    }
    double nrm = 0.0;
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) nrm += w[i] * w[i];
    nrm = sqrt(block_sum_1024(nrm, sh));
    const double ref = j == 0 ? sqrt(nrm0) : sc[1];
    const bool bad = !(nrm > 1e-13 * fabs(ref)) || !isfinite(nrm);
    const double inv = bad ? 0.0 : 1.0 / nrm;
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) wg[i] = w[i] * inv;
    if (threadIdx.x == 0) {
        sc[2 + j] = nrm;
        if (bad) atomicOr(flag, 1u);
    }
}
// H[a][b] = <V[:, a], CV[:, b]>  (L x L, one CTA per entry)
__global__ void __launch_bounds__(128) krylov_project_kernel(const double *__restrict__ V, const double *__restrict__ CV, uint32_t d, int L,
                                                             double *__restrict__ H) {
    const int a = blockIdx.x % L, b = blockIdx.x / L;
    const double *va = V + (size_t)a * d, *cb = CV + (size_t)b * d;
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < d; i += 128) acc += va[i] * cb[i];
    __shared__ double sh[4];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) H[(size_t)a * L + b] = (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

// Cholesky factor of the b x b Gram matrix of the block (CholeskyQR), one CTA, the triangle held in shared memory
// (b (b + 1) / 2 doubles: 148 KB at b = 192). G is column-major with the UPPER triangle valid; on exit the upper triangle
// holds R with G = R^T R (what cusolverDnDpotrf(UPPER) leaves, at a quarter of its ~0.2 ms). info = the 1-based index of a
// non-positive / non-finite pivot, else 0. Right-looking: column k is scaled, then the trailing triangle is updated.
__global__ void __launch_bounds__(1024) chol_upper_kernel(double *__restrict__ G, uint32_t b, int *__restrict__ info) {
    extern __shared__ double Lt[];  // L[i][j], j <= i, at i (i + 1) / 2 + j   (L = R^T)
    __shared__ double s_piv;
    __shared__ int s_bad;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    for (uint32_t e = tid; e < b * b; e += nt) {
        const uint32_t i = e % b, j = e / b;  // G(i, j), column-major
        if (i <= j) Lt[(size_t)j * (j + 1) / 2 + i] = G[e];
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (uint32_t k = 0; k < b; ++k) {
        if (tid == 0) {
            const double p = Lt[(size_t)k * (k + 1) / 2 + k];
            if (!(p > 0.0) || !isfinite(p)) {
                if (!s_bad) s_bad = (int)k + 1;
                s_piv = 1.0;
            } else {
                s_piv = sqrt(p);
            }
        }
        __syncthreads();
        const double piv = s_piv;
        for (uint32_t i = k + tid; i < b; i += nt) {
            double &x = Lt[(size_t)i * (i + 1) / 2 + k];
            x = i == k ? piv : x / piv;
        }
        __syncthreads();
        // trailing update: L[i][j] -= L[i][k] L[j][k] for k < j <= i < b; one warp per row, lanes along the row
        for (uint32_t i = k + 1 + (tid >> 5); i < b; i += nt >> 5) {
            double *row = Lt + (size_t)i * (i + 1) / 2;
            const double lik = row[k];
            for (uint32_t j = k + 1 + (tid & 31); j <= i; j += 32) row[j] -= lik * Lt[(size_t)j * (j + 1) / 2 + k];
        }
        __syncthreads();
    }
    for (uint32_t e = tid; e < b * b; e += nt) {
        const uint32_t i = e % b, j = e / b;
        if (i <= j) G[e] = Lt[(size_t)j * (j + 1) / 2 + i];
    }
    if (tid == 0) *info = s_bad;
}

// cyclic Jacobi on a small symmetric matrix (row-major n x n, destroyed); eigenvalues ascending, vecs[i * n + j] =
// component i of eigenvector j. Host side: the L x L Krylov projection only.
static void jacobi_eigh(std::vector<double> &a, int n, std::vector<double> &evals, std::vector<double> &vecs) {
    vecs.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) vecs[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; ++i) {
            diag += a[(size_t)i * n + i] * a[(size_t)i * n + i];
            for (int j = i + 1; j < n; ++j) off += a[(size_t)i * n + j] * a[(size_t)i * n + j];
        }
        if (off <= 1e-30 * (diag + off)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s2 = t * c;
                for (int r = 0; r < n; ++r) {  // columns p, q
                    const double arp = a[(size_t)r * n + p], arq = a[(size_t)r * n + q];
                    a[(size_t)r * n + p] = c * arp - s2 * arq;
                    a[(size_t)r * n + q] = s2 * arp + c * arq;
                }
                for (int r = 0; r < n; ++r) {  // rows p, q
                    const double apr = a[(size_t)p * n + r], aqr = a[(size_t)q * n + r];
                    a[(size_t)p * n + r] = c * apr - s2 * aqr;
                    a[(size_t)q * n + r] = s2 * apr + c * aqr;
                }
                for (int r = 0; r < n; ++r) {
                    const double vrp = vecs[(size_t)r * n + p], vrq = vecs[(size_t)r * n + q];
                    vecs[(size_t)r * n + p] = c * vrp - s2 * vrq;
                    vecs[(size_t)r * n + q] = s2 * vrp + c * vrq;
                }
            }
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return a[(size_t)x * n + x] < a[(size_t)y * n + y]; });
    evals.resize(n);
    std::vector<double> sorted((size_t)n * n);
    for (int j = 0; j < n; ++j) {
        evals[j] = a[(size_t)order[j] * n + order[j]];
        for (int i = 0; i < n; ++i) sorted[(size_t)i * n + j] = vecs[(size_t)i * n + order[j]];
    }
    vecs.swap(sorted);
}

struct ChfsiStats {
    int block_products = 0, cholqr = 0, rayleigh_ritz = 0, outer = 0;
    double max_residual = 0.0;
};

// SRB_EIG_TRACE=1: wall-clock per phase (synchronising — for probes only, never in a timed run)
struct PhaseTrace {
    bool on;
    cudaStream_t s;
    double t0, acc[5] = {0, 0, 0, 0, 0};  // krylov, filter, cholqr, rayleigh-ritz, other
    static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
    explicit PhaseTrace(cudaStream_t st) : s(st) {
        const char *e = getenv("SRB_EIG_TRACE");
        on = e && e[0] == '1';
        if (on) cudaStreamSynchronize(s), t0 = now();
    }
    void mark(int phase) {
        if (!on) return;
        cudaStreamSynchronize(s);
        const double t = now();
        acc[phase] += t - t0, t0 = t;
    }
    void report(const ChfsiStats &st) const {
        if (on)
            fprintf(stderr, "[srb] chfsi phases (ms): krylov %.3f filter %.3f cholqr %.3f rayleigh-ritz %.3f other %.3f | products %d cholqr %d outer %d\n",
                    acc[0], acc[1], acc[2], acc[3], acc[4], st.block_products, st.cholqr, st.outer);
    }
};

// Leading k eigenpairs of the symmetric d x d matrix d_C (column-major, untouched unless true is returned): on success
// columns 0..k-1 of d_C hold the eigenvectors of the k largest eigenvalues in ASCENDING order and d_evals[0..k-1] the values.
static bool chfsi_topk(srb_ctx *ctx, cublasHandle_t bl, cusolverDnHandle_t so, cudaStream_t es, double *d_C, uint32_t d, uint32_t k,
                       double *d_evals) {
    // tunables (round-2 sweeps without a rebuild; the defaults are the measured configuration):
    //   SRB_CHFSI_KRYLOV  Krylov steps for the bounds (8..40, default 16: on the spectra of the 1M / 2M / 4M-cell bench
    //                     matrices the resulting estimates give 58-70 block products every time, where 24 steps give 51
    //                     on a lucky draw and 85-130, or a second outer round, on others)   SRB_CHFSI_BLOCK  block width b
    //   SRB_CHFSI_TARGET  log10 of the gain of the k-th eigenvalue over the cut per outer round (default 11)
    static const int L = [] {
        const char *e = getenv("SRB_CHFSI_KRYLOV");
        return e ? std::max(8, std::min(40, atoi(e))) : 16;
    }();
    static const uint32_t b_env = [] {
        const char *e = getenv("SRB_CHFSI_BLOCK");
        return e ? (uint32_t)std::max(0, atoi(e)) / 32 * 32 : 0u;
    }();
    static const double kTarget = [] {
        const char *e = getenv("SRB_CHFSI_TARGET");
        return std::pow(10.0, e ? std::max(6.0, std::min(14.0, atof(e))) : 11.0);
    }();
    constexpr int kMaxOuter = 6, kMaxRounds = 24, kMaxDegree = 32, kMaxProducts = 400;
    constexpr double kAmpCap = 1e8, kTol = 1e-11;
    // block width: 3 k (or k + 96) rounded up to 32 — 160 at k = 50 (measured at the bench size: 6.7 ms against 7.5 ms for 192:
    // the same 51 block products, each narrower, and a smaller Rayleigh-Ritz problem)
    uint32_t b = std::min<uint32_t>(d / 4, ((std::max<uint32_t>(3 * k, k + 96) + 31) / 32) * 32);
    if (b_env >= k + 16 && b_env <= d / 2) b = b_env;
    if (b < k + 16 || d < 512) return false;
    const double one = 1.0, zero = 0.0, minus1 = -1.0;
    const size_t dd = (size_t)d * d, db = (size_t)d * b;
    ChfsiStats st;
    // ---- workspace ----
    Buf bV = dev_alloc(es, 8 * (size_t)d * (L + 1)), bCV = dev_alloc(es, 8 * (size_t)d * L), bH = dev_alloc(es, 8 * (size_t)L * L);
    Buf bsc = dev_zeros(es, 8 * (L + 4)), bflag = dev_zeros(es, 4 * 64);
    Buf bCs = dev_alloc(es, 8 * dd), bY0 = dev_alloc(es, 8 * db), bY1 = dev_alloc(es, 8 * db), bW = dev_alloc(es, 8 * db);
    Buf bG = dev_alloc(es, 8 * (size_t)b * b), bth = dev_alloc(es, 8 * b), bres = dev_alloc(es, 8 * k), bCW = dev_alloc(es, 8 * (size_t)d * k);
    double *V = bV->as<double>(), *CV = bCV->as<double>(), *H = bH->as<double>(), *sc = bsc->as<double>();
    uint32_t *flag = bflag->as<uint32_t>();  // [0] Krylov breakdown, [1 + i] LAPACK infos
    int *infos = reinterpret_cast<int *>(flag + 1);
    int n_info = 0;
    double *Cs = bCs->as<double>(), *Y = bY0->as<double>(), *Yb = bY1->as<double>(), *W = bW->as<double>();
    double *G = bG->as<double>(), *theta = bth->as<double>(), *res = bres->as<double>(), *CW = bCW->as<double>();
    int lw_potrf = 0, lw_syevd = 0;
    SRB_CUSOLVER(cusolverDnDpotrf_bufferSize(so, CUBLAS_FILL_MODE_UPPER, (int)b, G, (int)b, &lw_potrf));
    SRB_CUSOLVER(cusolverDnDsyevd_bufferSize(so, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)b, G, (int)b, theta, &lw_syevd));
    // SRB_CHFSI_RR=syevj: Jacobi instead of divide-and-conquer for the b x b Rayleigh-Ritz problem (probe)
    static const bool rr_jacobi = [] {
        const char *e = getenv("SRB_CHFSI_RR");
        return e && !strcmp(e, "syevj");
    }();
    syevjInfo_t jinfo = nullptr;
    int lw_syevj = 0;
    if (rr_jacobi) {
        SRB_CUSOLVER(cusolverDnCreateSyevjInfo(&jinfo));
        cusolverDnXsyevjSetTolerance(jinfo, 1e-15);
        cusolverDnXsyevjSetMaxSweeps(jinfo, 30);
        cusolverDnXsyevjSetSortEig(jinfo, 1);
        SRB_CUSOLVER(cusolverDnDsyevj_bufferSize(so, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)b, G, (int)b, theta, &lw_syevj, jinfo));
    }
    struct JGuard {
        syevjInfo_t p;
        ~JGuard() { if (p) cusolverDnDestroySyevjInfo(p); }
    } jguard{jinfo};
    Buf bwork = dev_alloc(es, 8 * (size_t)std::max(std::max(std::max(lw_potrf, lw_syevd), lw_syevj), 1));
    double *work = bwork->as<double>();
    SRB_CUBLAS(cublasSetStream(bl, es));
    PhaseTrace trace(es);

    // ---- 1. Krylov bounds: v_{j+1} = normalise((I - V V^T)^2 C v_j), two own launches per step ----
    if ((size_t)8 * (d + L + 1 + 256) > ctx->smem_optin) return false;  // w must fit one CTA's shared memory: syevd instead
    const size_t lz_smem = 8 * ((size_t)d + L + 1 + 256);
    SRB_CUDA(cudaFuncSetAttribute(lanczos_orth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz_smem));
    SRB_LAUNCH(lanczos_init_kernel, 1, 1024, 0, es, V, d, 0x5EEDC0DEull);
    for (int j = 0; j < L; ++j) {
        SRB_LAUNCH(symv_cols_kernel, (d + 7) / 8, 256, 0, es, d_C, V + (size_t)j * d, d, V + (size_t)(j + 1) * d, CV + (size_t)j * d);
        SRB_LAUNCH(lanczos_orth_kernel, 1, 1024, lz_smem, es, V, d, j, sc, flag);
    }
    SRB_LAUNCH(krylov_project_kernel, (unsigned)(L * L), 128, 0, es, V, CV, d, L, H);
    std::vector<double> hH((size_t)L * L), hsc(L + 4);
    uint32_t hflag0 = 0;
    SRB_CUDA(cudaMemcpyAsync(hH.data(), H, 8 * hH.size(), cudaMemcpyDeviceToHost, es));
    SRB_CUDA(cudaMemcpyAsync(hsc.data(), sc, 8 * hsc.size(), cudaMemcpyDeviceToHost, es));
    SRB_CUDA(cudaMemcpyAsync(&hflag0, flag, 4, cudaMemcpyDeviceToHost, es));
    SRB_CUDA(cudaStreamSynchronize(es));
    trace.mark(0);
    if (hflag0) return false;
    for (double x : hH)
        if (!std::isfinite(x)) return false;
    for (int i = 0; i < L; ++i)
        for (int j = i + 1; j < L; ++j) hH[(size_t)i * L + j] = hH[(size_t)j * L + i] = 0.5 * (hH[(size_t)i * L + j] + hH[(size_t)j * L + i]);
    std::vector<double> ritz, S;
    jacobi_eigh(hH, L, ritz, S);
    const double beta_last = hsc[2 + L - 1];  // norm of the next Krylov direction: residual of Ritz pair i = beta |S[L-1][i]|
    const double span = ritz[L - 1] - ritz[0];
    if (!(span > 0.0)) return false;
    double lo = ritz[0] - beta_last * fabs(S[(size_t)(L - 1) * L + 0]) - 0.01 * span;
    double up = ritz[L - 1] + beta_last * fabs(S[(size_t)(L - 1) * L + (L - 1)]);
    // density of states from the Ritz weights (first components squared): walk down from the top
    double cut = ritz[0], lamk = ritz[L - 1], cw = 0.0;
    bool have_k = false, have_cut = false;
    for (int i = L - 1; i >= 0; --i) {
        cw += S[(size_t)0 * L + i] * S[(size_t)0 * L + i];
        if (!have_k && cw >= (double)k / d) lamk = ritz[i], have_k = true;
        if (!have_cut && cw >= 0.8 * b / d) cut = ritz[i], have_cut = true;
    }
    {
        // The step function can cross k / d one Ritz value too early; an optimistic lamk makes the first outer round
        // under-filter (residual a few 1e-11, a second Rayleigh-Ritz step). The midpoint-rule CDF from the top, linearly
        // interpolated between the Ritz values, is smoother: take the lower of the two estimates.
        const double p = (double)k / d;
        double acc = 0.0, c_prev = 0.0, t_prev = up, q = ritz[0];
        bool found = false;
        for (int i = L - 1; i >= 0; --i) {  // nodes from the top; (c, t) = (CDF from the top at the node, Ritz value)
            const double w = S[(size_t)0 * L + i] * S[(size_t)0 * L + i], t = ritz[i], c = acc + 0.5 * w;
            if (!found && p <= c) {
                q = c > c_prev ? t_prev + (p - c_prev) / (c - c_prev) * (t - t_prev) : t;
                found = true;
            }
            c_prev = c, t_prev = t, acc += w;
        }
        lamk = std::min(lamk, q);
    }
    cut = std::min(cut, ritz[L - 1] - 0.02 * span);
    cut = std::max(cut, lo + 0.05 * span);
    lamk = std::max(lamk, cut + 0.01 * span);

    // ---- 2./3. filter + Rayleigh-Ritz ----
    SRB_LAUNCH(fill_random_kernel, (unsigned)std::min<size_t>((db + 255) / 256, 4096), 256, 0, es, Y, (uint64_t)db, 0xC4EB5EEDull);
    // column shards of the filter (row-sharded jobs only; every rank reaches this point with the same matrix and the same
    // decisions: the Lanczos bounds, CholeskyQR and Rayleigh-Ritz steps are replicated and deterministic)
    static const bool shard_off = [] {
        const char *e = getenv("SRB_CHFSI_SHARD");
        return e && e[0] == '0';
    }();
    const bool shard = !shard_off && ctx->nranks > 1 && ctx->comm && b % (uint32_t)ctx->nranks == 0 && b / (uint32_t)ctx->nranks >= 8;
    const uint32_t shard_cols = shard ? b / (uint32_t)ctx->nranks : b;
    const uint32_t shard_col0 = shard ? (uint32_t)ctx->rank * shard_cols : 0;
    const size_t chol_smem = 8 * (size_t)b * (b + 1) / 2;
    // SRB_CHFSI_CHOL=own selects the single-CTA kernel (measured SLOWER than cusolverDnDpotrf in its unblocked form: 0.38 vs
    // 0.2 ms at b = 192 — one latency-bound column at a time; kept for a blocked rewrite)
    static const bool want_own_chol = [] {
        const char *e = getenv("SRB_CHFSI_CHOL");
        return e && !strcmp(e, "own");
    }();
    const bool own_chol = want_own_chol && chol_smem + 1024 <= ctx->smem_optin;
    if (own_chol) SRB_CUDA(cudaFuncSetAttribute(chol_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem));
    auto cholqr = [&](double *Ycur) {
        SRB_CUBLAS(cublasDsyrk(bl, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, (int)b, (int)d, &one, Ycur, (int)d, &zero, G, (int)b));
        if (own_chol) SRB_LAUNCH(chol_upper_kernel, 1, 1024, chol_smem, es, G, b, infos + n_info);
        else SRB_CUSOLVER(cusolverDnDpotrf(so, CUBLAS_FILL_MODE_UPPER, (int)b, G, (int)b, work, lw_potrf, infos + n_info));
        ++n_info;
        SRB_CUBLAS(cublasDtrsm(bl, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)d, (int)b, &one, G, (int)b, Ycur, (int)d));
    };
    std::vector<double> hth(b), hres(k);
    std::vector<int> hinfo(63);
    bool converged = false;
    for (int outer = 0; outer < kMaxOuter && !converged; ++outer) {
        ++st.outer;
        const double e = 0.5 * (cut - lo), c = 0.5 * (cut + lo);
        if (!(e > 0.0) || !std::isfinite(e)) return false;
        const double xtop = std::max((up - c) / e, 1.0 + 1e-12), xk = std::max((lamk - c) / e, 1.0 + 1e-9);
        const int m = (int)std::max(2.0, std::min((double)kMaxDegree, std::floor(std::acosh(kAmpCap) / std::acosh(xtop))));
        const double amp = std::cosh(m * std::acosh(xk));
        // a later outer round only tops up what the last Rayleigh-Ritz step showed missing (with a margin of 1e3): a residual
        // of 3e-11 against the 1e-11 bound costs ~18 more block products, not a second full filter
        const double target_now = outer == 0 ? kTarget : std::min(kTarget, std::max(1e3, 1e3 * st.max_residual / kTol));
        int rounds = (int)std::ceil(std::log(target_now) / std::log(std::max(amp, 1.0001)));
        rounds = std::max(1, std::min(rounds, outer == 0 ? 3 : kMaxRounds));
        // the smallest degree that reaches the target in exactly `rounds` rounds (a ceil() on the round count would
        // otherwise overshoot by up to a whole round: 76 instead of 51 block products at the bench size)
        const int m_full = m;
        const int m_trim = (int)std::ceil(std::acosh(std::pow(target_now, 1.0 / rounds)) / std::acosh(xk));
        const int m_use = std::max(2, std::min(m_full, m_trim));
        const double amp_top_use = std::cosh(m_use * std::acosh(xtop));
        if (n_info + 2 * rounds + 1 > 60) return false;
        // work budget: beyond ~400 block products the iteration would cost more than the syevd it replaces
        if (st.block_products + rounds * m_use > kMaxProducts) return false;
        trace.mark(4);
        SRB_LAUNCH(shift_copy_kernel, (unsigned)((dd + 255) / 256), 256, 0, es, d_C, Cs, d, c);
        for (int r = 0; r < rounds; ++r) {
            // scaled Chebyshev recurrence (Zhou & Saad): Y_1 = (s1/e) Cs Y_0 ; Y_{i+1} = (2 s_{i+1}/e) Cs Y_i - s_i s_{i+1} Y_{i-1}
            const double sigma1 = e / (up - c);
            double sigma = sigma1, a1 = sigma1 / e;
            // The filter acts on every column independently. In a row-sharded job the correlation matrix is replicated
            // bit for bit (allreduced), so each rank filters its own b / nranks columns and the block is allgathered once
            // per round (2.5 MB over NVLink) instead of every rank filtering all b columns.
            const size_t co = (size_t)shard_col0 * d;
            SRB_CUBLAS(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, (int)d, (int)shard_cols, (int)d, &a1, Cs, (int)d, Y + co, (int)d, &zero, Yb + co, (int)d));
            double *Yp = Y, *Yc = Yb;
            for (int i = 2; i <= m_use; ++i) {
                const double sn = 1.0 / (2.0 / sigma1 - sigma), al = 2.0 * sn / e, be = -sigma * sn;
                SRB_CUBLAS(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, (int)d, (int)shard_cols, (int)d, &al, Cs, (int)d, Yc + co, (int)d, &be, Yp + co, (int)d));
                std::swap(Yp, Yc);
                sigma = sn;
            }
            st.block_products += m_use;
            if (shard_cols != b) allgather_f64(ctx, es, Yc, (size_t)shard_cols * d);
            if (Yc != Y) std::swap(Y, Yb);  // Y = filtered block, Yb = scratch
            trace.mark(1);
            // CholeskyQR: one pass leaves an orthogonality error of ~eps * cond(Y)^2. Before Rayleigh-Ritz the block must be
            // orthonormal to working precision (two passes); between filter rounds it only has to stay well conditioned,
            // which one pass guarantees while the round amplified the top by <= 1e7 (error <= 1e-2)
            cholqr(Y);
            if (r + 1 == rounds || amp_top_use > 1e7) cholqr(Y);
            ++st.cholqr;
            trace.mark(2);
        }
        // Rayleigh-Ritz
        SRB_CUBLAS(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, (int)d, (int)b, (int)d, &one, d_C, (int)d, Y, (int)d, &zero, W, (int)d));
        SRB_CUBLAS(cublasDgemm(bl, CUBLAS_OP_T, CUBLAS_OP_N, (int)b, (int)b, (int)d, &one, Y, (int)d, W, (int)d, &zero, G, (int)b));
        SRB_LAUNCH(symmetrize_kernel, (b * b + 255) / 256, 256, 0, es, G, b);
        if (rr_jacobi)
            SRB_CUSOLVER(cusolverDnDsyevj(so, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)b, G, (int)b, theta, work, lw_syevj, infos + n_info, jinfo));
        else
            SRB_CUSOLVER(cusolverDnDsyevd(so, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)b, G, (int)b, theta, work, lw_syevd, infos + n_info));
        ++n_info;
        ++st.block_products, ++st.rayleigh_ritz;
        SRB_CUBLAS(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, (int)d, (int)b, (int)b, &one, Y, (int)d, G, (int)b, &zero, Yb, (int)d));
        SRB_CUBLAS(cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, (int)d, (int)k, (int)b, &one, W, (int)d, G + (size_t)(b - k) * b, (int)b, &zero, CW, (int)d));
        std::swap(Y, Yb);  // Y = Ritz vectors (ascending)
        SRB_LAUNCH(residual_norms_kernel, k, 256, 0, es, CW, Y, theta, d, b - k, res);
        SRB_CUDA(cudaMemcpyAsync(hth.data(), theta, 8 * b, cudaMemcpyDeviceToHost, es));
        SRB_CUDA(cudaMemcpyAsync(hres.data(), res, 8 * k, cudaMemcpyDeviceToHost, es));
        SRB_CUDA(cudaMemcpyAsync(hinfo.data(), infos, 4 * n_info, cudaMemcpyDeviceToHost, es));
        SRB_CUDA(cudaStreamSynchronize(es));
        trace.mark(3);
        for (int i = 0; i < n_info; ++i)
            if (hinfo[i] != 0) return false;
        n_info = 0;
        double rmax = 0.0;
        for (uint32_t j = 0; j < k; ++j) {
            if (!std::isfinite(hres[j])) return false;
            rmax = std::max(rmax, hres[j]);
        }
        for (uint32_t j = 0; j < b; ++j)
            if (!std::isfinite(hth[j])) return false;
        st.max_residual = rmax / std::max(fabs(hth[b - 1]), 1e-300);
        converged = st.max_residual <= kTol;
        // the block's smallest Ritz value raises the cut when the first one left more than b eigenvalues above it; it only
        // lowers it when the k-th Ritz value shows the cut sat above wanted eigenvalues (a badly converged last vector has a
        // Ritz value deep inside the bulk, which would blunt the next filter)
        lamk = hth[b - k], up = std::max(up, hth[b - 1]);
        cut = lamk <= cut ? hth[0] : std::max(cut, hth[0]);
        if (!(cut > lo)) lo = cut - 0.05 * (up - cut);
    }
    ctx->last_eig_products = st.block_products, ctx->last_eig_outer = st.outer, ctx->last_eig_residual = st.max_residual;
    trace.report(st);
    if (!converged) return false;
    SRB_CUDA(cudaMemcpyAsync(d_C, Y + (size_t)(b - k) * d, 8 * (size_t)d * k, cudaMemcpyDeviceToDevice, es));
    SRB_CUDA(cudaMemcpyAsync(d_evals, theta + (b - k), 8 * k, cudaMemcpyDeviceToDevice, es));
    return true;
}

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
        register_stream(ctx->eig_stream);
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
    static int eig_mode = -1;  // SRB_EIG_MODE: "chfsi" (default: top-k by Chebyshev-filtered subspace iteration) | "syevd"
    if (eig_mode < 0) {
        const char *e = getenv("SRB_EIG_MODE");
        eig_mode = (e && !strcmp(e, "chfsi")) ? 1 : (e && !strcmp(e, "syevd")) ? 0 : SRB_EIG_DEFAULT_MODE;
    }
    ctx->last_eig_mode = 0;
    const int mode = ctx->eig_mode >= 0 ? ctx->eig_mode : eig_mode;
    if (mode == 1 && topk < d && d >= 1024 && topk * 8 <= d) {
        if (!ctx->blas) {
            cublasHandle_t bh;
            SRB_CUBLAS(cublasCreate(&bh));
            ctx->blas = bh;
        }
        SRB_CUDA(cudaEventRecord(ctx->eig_in, s));
        SRB_CUDA(cudaStreamWaitEvent(es, ctx->eig_in, 0));
        // the work buffers live in the eig stream's own block cache (allocated, used and released on that stream only)
        bool ok = false;
        try {
            ok = chfsi_topk(ctx, (cublasHandle_t)ctx->blas, h, es, d_C, d, topk, d_evals);
        } catch (...) {
            // leave the cached handle and the two streams in a defined state before the error travels on: host pointer
            // mode again, and nothing of the eig stream still in flight when its buffers return to the block caches
            cublasSetPointerMode((cublasHandle_t)ctx->blas, CUBLAS_POINTER_MODE_HOST);
            cudaEventRecord(ctx->eig_out, es);
            cudaStreamWaitEvent(s, ctx->eig_out, 0);
            cudaStreamSynchronize(es);
            throw;
        }
        ctx->last_eig_mode = ok ? 1 : 2;
        SRB_CUDA(cudaEventRecord(ctx->eig_out, es));
        SRB_CUDA(cudaStreamWaitEvent(s, ctx->eig_out, 0));
        SRB_CUDA(cudaStreamSynchronize(es));
        if (ok) return topk;
        // fall through: syevd on the untouched matrix
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
    if (const char *dump = getenv("SRB_EIG_DUMP")) {  // probe: the full spectrum (ascending f64) for offline solver tuning
        std::vector<double> ev(d);
        SRB_CUDA(cudaMemcpy(ev.data(), d_evals, 8 * (size_t)d, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(dump, "wb")) {
            fwrite(ev.data(), 8, d, f);
            fclose(f);
        }
    }
    return d;
}

void eig_destroy(srb_ctx *ctx) {
    if (ctx->blas) cublasDestroy((cublasHandle_t)ctx->blas);
    ctx->blas = nullptr;
    if (ctx->eig_stream) release_cached_blocks(ctx->eig_stream), unregister_stream(ctx->eig_stream);
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
