// pca.cu — K6..K9: PCA on the selected (HVG) columns. Replaces pca_inplace, src/memory/processing/dim_red/mod.rs:24-94
// (select -> convert_to_array_f64_selected (shared/mod.rs:230-259) -> PCABuilder.fit/transform) with a Gram
// formulation that never builds the n x d f64 block the reference allocates (16 GB at 1M x 2000):
//
//   mu = sum_j / n, sigma^2 = sumsq_j / n - mu^2   from the per-gene moments already computed (ddof = 0, ALL cells)
//   Z  = (X - mu)/sigma : n x d block of the standardised selected columns (pca/mod.rs:87-111 semantics), held as
//        split-fp16 panels Zh + Zl (22-bit mantissa, |z - (zh+zl)| <= 2^-22 |z|), row-major [n][dpad]. The block is
//        centred BEFORE the Gram product: subtracting n mu mu^T afterwards would turn any accumulation bias of the
//        tensor-core sums into a rank-one perturbation along mu/sigma ("why centre first").
//   C  = Z^T Z                        (d x d, tensor cores or fp64 CUDA cores; allreduced over row shards)
//   C  = V L V^T (symmetric eig);  components = V[:, :k];  ratio = L[:k] / trace(C)   (pca/mod.rs:131-151)
//   scores = Z V_k                                                          (pca/mod.rs:156-185)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace srb {

static unsigned nb(uint64_t n, unsigned t = 256) { return (unsigned)std::max<uint64_t>(1, (n + t - 1) / t); }

// SRB_PCA_SHIFT=center: every panel column is centred before the Gram product (round-1 behaviour); default "sparse": sparse
// genes keep their zeros and are corrected by the rank-one term afterwards (see sel_stats_kernel)
static int sparse_shift_mode() {
    static const int v = [] {
        const char *e = getenv("SRB_PCA_SHIFT");
        return (e && !strcmp(e, "center")) ? 0 : 1;
    }();
    return v;
}

// lut[col] = position in the selection, -1 if not selected; duplicates: the LAST position wins (HashMap insert
// order, shared/mod.rs:241)
__global__ void lut_fill_kernel(int *lut, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lut[i] = -1;
}
__global__ void lut_set_kernel(int *lut, const uint32_t *sel, uint64_t n_sel) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_sel) atomicMax(&lut[sel[j]], (int)j);
}

// 16-bit LUT for the panel builder (60 KB at 30 k genes: stays in L1); 0xFFFF = not selected
__global__ void lut16_kernel(const int *__restrict__ lut, uint16_t *__restrict__ lut16, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lut16[i] = lut[i] < 0 ? (uint16_t)0xFFFF : (uint16_t)lut[i];
}
__global__ void shis_kernel(const float *__restrict__ shf, const float *__restrict__ isf, float2 *__restrict__ shis, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) shis[i] = make_float2(shf[i], isf[i]);
}

// reference-shaped dense block (parity/debug): all rows [row0,row0+nrows) x selection, f64 row-major
template <typename VT>
__global__ void __launch_bounds__(256) densify_f64_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                          const VT *__restrict__ val, const int *__restrict__ lut,
                                                          uint64_t row0, uint64_t nrows, uint64_t n_sel, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t r = warp; r < nrows; r += nwarps) {
        const int64_t a = off[row0 + r], b = off[row0 + r + 1];
        for (int64_t k = a + lane; k < b; k += 32) {
            const int p = lut[idx[k]];
            if (p >= 0) out[r * n_sel + (uint64_t)p] = (double)val[k];
        }
    }
}

void densify_selected_f64(srb_mat *m, const uint32_t *d_sel, uint64_t n_sel, double *d_out, uint64_t row0, uint64_t nrows) {
    if (m->has_pending()) materialize(m, false);
    cudaStream_t s = m->ctx->stream;
    const uint64_t M = m->nminor();
    Buf lut = dev_alloc(s, 4 * (M + 1));
    SRB_LAUNCH(lut_fill_kernel, nb(M), 256, 0, s, lut->as<int>(), M);
    SRB_LAUNCH(lut_set_kernel, nb(n_sel), 256, 0, s, lut->as<int>(), d_sel, n_sel);
    SRB_CUDA(cudaMemsetAsync(d_out, 0, 8 * nrows * n_sel, s));
    const unsigned grid = (unsigned)std::min<uint64_t>((nrows + 7) / 8, (uint64_t)m->ctx->sm_count * 32);
    if (m->vdtype == SRB_F32)
        SRB_LAUNCH((densify_f64_kernel<float>), grid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), lut->as<int>(), row0, nrows, n_sel, d_out);
    else
        SRB_LAUNCH((densify_f64_kernel<double>), grid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<double>(), lut->as<int>(), row0, nrows, n_sel, d_out);
}

// K6: CSR rows -> standardised split-fp16 panels, one warp per row, no shared memory (64 resident warps per SM hide the
// HBM latency). Step 1 streams the row of implicit-zero constants (0 - shift_j) * inv_sd_j to global memory with
// 16-byte stores; step 2 overwrites the stored entries with 2-byte scattered stores — they hit the lines step 1 just
// put into L2, so DRAM sees each line once. __syncwarp() orders the two steps inside the warp.
// HBM: 8 B/nnz in, 4*dpad B/row out.
// (Two cp.async.bulk-staged variants were measured in round 2 and removed: 8 warps sharing one tile, the row of an entry found
// by binary search, a barrier per tile: 12.5 ms; row-owning consumer warps reading a shared-memory ring like K1, selection
// bitmap in shared memory: 4.95 ms. This kernel with 8 loads per array in flight per lane: 4.3-4.4 ms. Staging the loads does
// not help here because the stall is the dependent chain LUT -> (shift, 1/sd) -> store of the selected 7 %, not the stream.
// Loading a value only after the LUT said its entry is selected (45 % of the value sectors instead of all): 4.06 vs 4.12 ms,
// not kept.)
template <typename VT>
__global__ void __launch_bounds__(256) densify_panels_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                             const VT *__restrict__ val, const uint16_t *__restrict__ lut,
                                                             const float2 *__restrict__ shis,
                                                             const __half *__restrict__ zc_h, const __half *__restrict__ zc_l,
                                                             uint64_t nrows, uint32_t dpad, __half *__restrict__ Xh,
                                                             __half *__restrict__ Xl) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    const uint32_t nvec = dpad / 8;  // uint4 = 8 halves
    const uint4 *ch = reinterpret_cast<const uint4 *>(zc_h), *cl = reinterpret_cast<const uint4 *>(zc_l);
    constexpr int kBatch = 8;
    for (uint64_t r = warp; r < nrows; r += nwarps) {
        const int64_t a = off[r], b = off[r + 1];
        uint4 *oh = reinterpret_cast<uint4 *>(Xh + r * dpad), *ol = reinterpret_cast<uint4 *>(Xl + r * dpad);
        for (uint32_t i = lane; i < nvec; i += 32) oh[i] = ch[i], ol[i] = cl[i];
        __syncwarp();
        __half *rh = Xh + r * dpad, *rl = Xl + r * dpad;
        for (int64_t k0 = a + lane; k0 < b; k0 += 32 * kBatch) {
            uint32_t cc[kBatch];
            float vv[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int64_t k = k0 + 32 * u;
                const bool in = k < b;
                cc[u] = in ? idx[k] : 0xFFFFFFFFu;
                vv[u] = in ? (float)val[k] : 0.f;
            }
            uint32_t pp[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) pp[u] = cc[u] != 0xFFFFFFFFu ? (uint32_t)lut[cc[u]] : 0xFFFFu;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const uint32_t p = pp[u];
                if (p != 0xFFFFu) {
                    // fp32 is enough here: the split keeps 22 bits of z, fp32 carries 24
                    const float2 si = shis[p];
                    const float z = (vv[u] - si.x) * si.y;
                    const __half h = __float2half_rn(z);
                    rh[p] = h;
                    rl[p] = __float2half_rn(z - __half2float(h));
                }
            }
        }
        __syncwarp();
    }
}

// default variant: register double buffer (SRB_DENSIFY_PIPE=0 selects the plain batched kernel above)
template <typename VT, int kBatch>
__global__ void __launch_bounds__(256, kBatch == 4 ? 4 : 3) densify_panels_pipe_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                             const VT *__restrict__ val, const uint16_t *__restrict__ lut,
                                                             const float2 *__restrict__ shis,
                                                             const __half *__restrict__ zc_h, const __half *__restrict__ zc_l,
                                                             uint64_t nrows, uint32_t dpad, __half *__restrict__ Xh,
                                                             __half *__restrict__ Xl) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    const uint32_t nvec = dpad / 8;  // uint4 = 8 halves
    const uint4 *ch = reinterpret_cast<const uint4 *>(zc_h), *cl = reinterpret_cast<const uint4 *>(zc_l);
    // the row bounds are fetched one row ahead and the first batch of (index, value) loads is issued BEFORE the 8 KB of
    // implicit-zero constants are written, so neither the offset load nor the fill leaves the warp without loads in flight
    uint64_t r = warp;
    int64_t a_next = 0, b_next = 0;
    if (r < nrows) a_next = off[r], b_next = off[r + 1];
    for (; r < nrows; r += nwarps) {
        const int64_t a = a_next, b = b_next;
        const uint64_t rn = r + nwarps;
        if (rn < nrows) a_next = off[rn], b_next = off[rn + 1];
        // software pipeline: the next batch of (index, value) loads is in flight while the current batch walks the
        // dependent chain index -> LUT -> (shift, 1/sd) -> store
        uint32_t cc[kBatch], cn[kBatch];
        float vv[kBatch], vn[kBatch];
        int64_t k0 = a + lane;
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int64_t k = k0 + 32 * u;
            cc[u] = k < b ? idx[k] : 0xFFFFFFFFu;
            vv[u] = k < b ? (float)val[k] : 0.f;
        }
        uint4 *oh = reinterpret_cast<uint4 *>(Xh + r * dpad), *ol = reinterpret_cast<uint4 *>(Xl + r * dpad);
        for (uint32_t i = lane; i < nvec; i += 32) oh[i] = ch[i], ol[i] = cl[i];
        __syncwarp();
        __half *rh = Xh + r * dpad, *rl = Xl + r * dpad;
        for (; k0 < b; k0 += 32 * kBatch) {
            const int64_t kn = k0 + 32 * kBatch;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int64_t k = kn + 32 * u;
                cn[u] = k < b ? idx[k] : 0xFFFFFFFFu;
                vn[u] = k < b ? (float)val[k] : 0.f;
            }
            uint32_t pp[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) pp[u] = cc[u] != 0xFFFFFFFFu ? (uint32_t)lut[cc[u]] : 0xFFFFu;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const uint32_t p = pp[u];
                if (p != 0xFFFFu) {
                    const float2 si = shis[p];
                    const float z = (vv[u] - si.x) * si.y;
                    const __half h = __float2half_rn(z);
                    rh[p] = h;
                    rl[p] = __float2half_rn(z - __half2float(h));
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) cc[u] = cn[u], vv[u] = vn[u];
        }
        __syncwarp();
    }
}

// K7 (validation path): G += X^T X in fp64 on CUDA cores. 64x64 output tile per CTA, upper-triangular tiles only,
// split over row ranges (fp64 REDs into G). x = (double)xh + (double)xl.
static constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256) gram_f64_kernel(const __half *__restrict__ Xh, const __half *__restrict__ Xl,
                                                       uint64_t nrows, uint32_t dpad, uint64_t rows_per_split,
                                                       double *__restrict__ G) {
    // decode upper-triangular tile (ti <= tj)
    const uint32_t nt = dpad / GT;
    uint32_t t = blockIdx.x, ti = 0;
    while (t >= nt - ti) t -= nt - ti, ++ti;
    const uint32_t tj = ti + t;
    const uint64_t r0 = (uint64_t)blockIdx.y * rows_per_split;
    const uint64_t r1 = min(r0 + rows_per_split, nrows);
    __shared__ double As[GK][GT + 1], Bs[GK][GT + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    double acc[4][4] = {};
    for (uint64_t rk = r0; rk < r1; rk += GK) {
        for (int e = threadIdx.x; e < GK * GT; e += 256) {
            const int kk = e / GT, c = e % GT;
            const uint64_t r = rk + kk;
            double a = 0.0, b = 0.0;
            if (r < r1) {
                const uint64_t ia = r * dpad + (uint64_t)ti * GT + c, ib = r * dpad + (uint64_t)tj * GT + c;
                a = (double)__half2float(Xh[ia]) + (double)__half2float(Xl[ia]);
                b = (double)__half2float(Xh[ib]) + (double)__half2float(Xl[ib]);
            }
            As[kk][c] = a, Bs[kk][c] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i], bv[i] = Bs[kk][tx * 4 + i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t gi = ti * GT + ty * 4 + i, gj = tj * GT + tx * 4 + j;
            if (acc[i][j] != 0.0) atomicAdd(&G[(uint64_t)gi * dpad + gj], acc[i][j]);
        }
}
// copy the upper triangle (tile-wise computed: entries with tile(i) <= tile(j)) onto the lower one
__global__ void gram_mirror_kernel(double *G, uint32_t dpad, uint32_t tile) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (uint64_t)dpad * dpad) return;
    const uint32_t i = (uint32_t)(e / dpad), j = (uint32_t)(e % dpad);
    if (i / tile > j / tile) G[e] = G[(uint64_t)j * dpad + i];
}

// Row-sharded jobs exchange only what is needed of the symmetric Gram matrix: the upper triangle of its leading d x d block,
// packed row by row (16 MB instead of 33.5 MB at d = 2000, dpad = 2048). T[i d - i (i - 1) / 2 + (j - i)] = G[i][j], i <= j.
__global__ void tri_pack_kernel(const double *__restrict__ G, uint32_t dpad, uint32_t d, double *__restrict__ T) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (uint64_t)d * d) return;
    const uint64_t i = e / d, j = e % d;
    if (i <= j) T[i * d - i * (i - 1) / 2 + (j - i)] = G[i * dpad + j];
}
__global__ void tri_unpack_kernel(const double *__restrict__ T, uint32_t dpad, uint32_t d, double *__restrict__ G) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (uint64_t)d * d) return;
    const uint64_t i = e / d, j = e % d;
    if (i <= j) {
        const double v = T[i * d - i * (i - 1) / 2 + (j - i)];
        G[i * dpad + j] = v;
        G[j * dpad + i] = v;
    }
}
// sum of the local Gram matrices over the ranks (the one exchange of the PCA stage)
static void allreduce_gram(srb_ctx *c, double *G, uint32_t dpad, uint64_t n_sel) {
    cudaStream_t s = c->stream;
    const uint32_t d = (uint32_t)n_sel;
    const uint64_t tri = (uint64_t)d * (d + 1) / 2;
    Buf T = dev_alloc(s, 8 * tri);
    SRB_LAUNCH(tri_pack_kernel, nb((uint64_t)d * d), 256, 0, s, G, dpad, d, T->as<double>());
    allreduce_f64_sum(c, T->as<double>(), (size_t)tri);
    SRB_LAUNCH(tri_unpack_kernel, nb((uint64_t)d * d), 256, 0, s, T->as<double>(), dpad, d, G);
}

// per selected gene: mean over ALL cells and 1/std (ddof 0) from the global per-gene moments.
// Panel column j holds z' = (x - s_j) / sd_j. For a SPARSE gene (mean^2 <= 0.25 var: most cells do not express it) the panel
// shift s_j is 0 instead of the mean, so the implicit zeros of the count matrix stay exact zeros in the fp16 panels (95 % of
// the entries at the bench density) and the products that reach the tensor cores are mostly zero: fewer terms per fp32
// accumulation chunk (less rounding) and far less switching in the MMA datapath, which runs at the 1 kW power cap. The
// column is then off by the constant w_j = (mean_j - s_j) / sd_j:  Z = Z' - 1 w^T,  Z^T Z = Z'^T Z' - n w w^T  (Z'^T 1 = n w),
// scores = Z' V - 1 (w^T V): both corrections are exact fp64 rank-one terms (corr_kernel, components_kernel -> bias). The
// cancellation is benign exactly when w is small, which is the criterion; dense genes keep the centre-first form (w_j = 0),
// whose reason is given at the top of the file.
__global__ void sel_stats_kernel(const uint32_t *__restrict__ sel, uint64_t n_sel, uint32_t dpad, const double *__restrict__ sum,
                                 const double *__restrict__ sq, double n_cells, int center, int scale, int sparse_shift,
                                 double *__restrict__ shift, double *__restrict__ inv_sd, double *__restrict__ wres, float *__restrict__ shf,
                                 float *__restrict__ isf, __half *__restrict__ zc_h, __half *__restrict__ zc_l,
                                 uint32_t *__restrict__ flag) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dpad) return;
    if (j >= n_sel) {
        shift[j] = 0.0, inv_sd[j] = 0.0, wres[j] = 0.0, shf[j] = 0.f, isf[j] = 0.f;
        zc_h[j] = __float2half(0.f), zc_l[j] = __float2half(0.f);
        return;
    }
    const uint32_t g = sel[j];
    const double mean = sum[g] / n_cells;
    double var = __dsub_rn(__ddiv_rn(sq[g], n_cells), __dmul_rn(mean, mean));
    if (var < 0.0) var = 0.0;
    const double sd = sqrt(var);
    double is = 1.0;
    if (scale) {
        if (sd > 0.0) is = 1.0 / sd;
        else { is = 0.0; atomicOr(flag, 1u); }  // the reference divides by zero here (NaN columns)
    }
    const double sh_full = center ? mean : 0.0;  // pca/mod.rs:98-104: subtract only when centring
    const bool sparse = sparse_shift && center && mean * mean <= 0.25 * var;
    const double sh = sparse ? 0.0 : sh_full;    // what the panels subtract
    shift[j] = sh_full, inv_sd[j] = is, wres[j] = (sh_full - sh) * is, shf[j] = (float)sh, isf[j] = (float)is;
    const double z0 = (0.0 - sh) * is;
    const __half h = __double2half(z0);
    zc_h[j] = h;
    zc_l[j] = __double2half(z0 - (double)__half2float(h));
}
// compact copy C[i][j] = G[i][j] (n_sel x n_sel out of dpad x dpad). The diagonal is replaced by its exact value
// sum_cells z^2 = inv_sd^2 (sumsq - 2 shift sum + n shift^2) from the fp64 gene moments (= n when centring and
// scaling): the diagonal sums only positive terms, which is where fp32 chunk accumulation has a systematic bias.
__global__ void corr_kernel(const double *__restrict__ G, uint32_t dpad, uint64_t n_sel, const uint32_t *__restrict__ sel,
                            const double *__restrict__ sum, const double *__restrict__ sq, const double *__restrict__ shift,
                            const double *__restrict__ inv_sd, const double *__restrict__ wres, double n_cells, double *__restrict__ C) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_sel * n_sel) return;
    const uint64_t i = e / n_sel, j = e % n_sel;
    double v = G[i * dpad + j] - n_cells * wres[i] * wres[j];  // Z^T Z = Z'^T Z' - n w w^T (sparse panel shift)
    if (i == j) {
        const uint32_t g = sel[i];
        const double sh = shift[i], is = inv_sd[i];
        v = is * is * (sq[g] - 2.0 * sh * sum[g] + n_cells * sh * sh);
    }
    C[e] = v;
}
__global__ void trace_kernel(const double *__restrict__ C, uint64_t n_sel, double *__restrict__ out) {
    double s = 0.0;
    for (uint64_t i = threadIdx.x; i < n_sel; i += blockDim.x) s += C[i * n_sel + i];
    s = warp_sum(s);
    __shared__ double part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (unsigned w = 0; w < blockDim.x / 32; ++w) t += part[w];
        out[0] = t;
    }
}
// eigenvectors come back column-major with ascending eigenvalues: column (n_sel-1-c) is component c.
// Sign convention: the entry of largest magnitude of every component is positive (deterministic across runs).
// comps[j][c] (row-major n_sel x k), W[p][c] = comps[p][c] (row-major dpad x kpad, zero padded), evr[c] = lambda_c / trace
__global__ void components_kernel(const double *__restrict__ evec, const double *__restrict__ evals_asc, uint32_t npairs, uint64_t n_sel,
                                  uint32_t k, uint32_t kpad, const double *__restrict__ trace, const double *__restrict__ wres,
                                  double *__restrict__ comps, double *__restrict__ W, double *__restrict__ evr, double *__restrict__ bias) {
    const uint32_t c = blockIdx.x;
    if (c >= k) return;
    const double *v = evec + (uint64_t)(npairs - 1 - c) * n_sel;
    __shared__ double s_val[256];
    __shared__ int s_idx[256];
    double best = -1.0;
    int bi = 0;
    for (uint64_t j = threadIdx.x; j < n_sel; j += blockDim.x) {
        const double a = fabs(v[j]);
        if (a > best) best = a, bi = (int)j;
    }
    s_val[threadIdx.x] = best, s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double ob = s_val[threadIdx.x + o];
            const int oi = s_idx[threadIdx.x + o];
            if (ob > s_val[threadIdx.x] || (ob == s_val[threadIdx.x] && oi < s_idx[threadIdx.x])) s_val[threadIdx.x] = ob, s_idx[threadIdx.x] = oi;
        }
        __syncthreads();
    }
    const double sgn = v[s_idx[0]] < 0.0 ? -1.0 : 1.0;
    __syncthreads();
    double b = 0.0;  // bias[c] = sum_j w_j V[j][c]: the constant the scores of the shifted panels are off by
    for (uint64_t j = threadIdx.x; j < n_sel; j += blockDim.x) {
        const double x = sgn * v[j];
        comps[j * k + c] = x;
        W[j * kpad + c] = x;
        b += wres[j] * x;
    }
    s_val[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_val[threadIdx.x] += s_val[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) evr[c] = evals_asc[npairs - 1 - c] / trace[0], bias[c] = s_val[0];
}

// K9 (validation path): scores[r][c] = sum_p z[r][p] W[p][c] in fp64 on CUDA cores. One warp per 4 rows; lane l owns
// components c0+l and c0+l+32; W rows are read once per 4 rows (L1), z values are warp-broadcast.
__global__ void __launch_bounds__(256) scores_simt_kernel(const __half *__restrict__ Xh, const __half *__restrict__ Xl,
                                                          uint64_t nrows, uint32_t dpad, const double *__restrict__ W,
                                                          uint32_t kpad, uint32_t k, uint32_t c0, const double *__restrict__ bias,
                                                          double *__restrict__ scores) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t r0 = warp * 4; r0 < nrows; r0 += nwarps * 4) {
        double acc[4][2] = {};
        for (uint32_t p0 = 0; p0 < dpad; p0 += 32) {
            double x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t r = min(r0 + i, nrows - 1);
                x[i] = (double)__half2float(Xh[r * dpad + p0 + lane]) + (double)__half2float(Xl[r * dpad + p0 + lane]);
            }
#pragma unroll 8
            for (int src = 0; src < 32; ++src) {
                const double *wr = W + (uint64_t)(p0 + src) * kpad + c0;
                const double w0 = wr[lane], w1 = wr[lane + 32];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double xv = __shfl_sync(0xffffffffu, x[i], src);
                    acc[i][0] += xv * w0;
                    acc[i][1] += xv * w1;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint64_t r = r0 + i;
            if (r < nrows) {
                if (c0 + lane < k) scores[r * k + c0 + lane] = acc[i][0] - bias[c0 + lane];
                if (c0 + lane + 32 < k) scores[r * k + c0 + lane + 32] = acc[i][1] - bias[c0 + lane + 32];
            }
        }
    }
}

void pca_run(srb_mat *m, const uint32_t *d_sel, uint64_t n_sel, uint64_t k, bool center, bool scale, int gram_mode,
             PcaOut out) {
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    SRB_REQUIRE(gram_mode == 0 || gram_mode == 1, SRB_ERR_INVALID_ARG, "gram_mode must be 0 (tcgen05) or 1 (fp64)");
    const uint64_t n = m->nrows, M = m->nminor();
    const double n_cells = (double)(c->nranks > 1 ? m->global_nrows : m->nrows);
    SRB_REQUIRE(n_cells >= 2, SRB_ERR_INVALID_ARG, "PCA needs at least two cells");
    SRB_TRACE("pca_run begin");
    ensure_minor_moments(m);  // also applies pending transforms (fused)
    SRB_TRACE("ensure_minor_moments");

    const uint32_t dpad = (uint32_t)((n_sel + 255) / 256 * 256);
    const uint32_t kpad = (uint32_t)((k + 63) / 64 * 64);
    Buf lut = dev_alloc(s, 4 * (M + 1));
    SRB_LAUNCH(lut_fill_kernel, nb(M), 256, 0, s, lut->as<int>(), M);
    SRB_LAUNCH(lut_set_kernel, nb(n_sel), 256, 0, s, lut->as<int>(), d_sel, n_sel);

    Buf flag = dev_zeros(s, 4);
    Buf shift = dev_alloc(s, 8 * dpad), inv_sd = dev_alloc(s, 8 * dpad), zc_h = dev_alloc(s, 2 * dpad), zc_l = dev_alloc(s, 2 * dpad);
    Buf shf = dev_alloc(s, 4 * dpad), isf = dev_alloc(s, 4 * dpad), wres = dev_alloc(s, 8 * dpad);
    SRB_LAUNCH(sel_stats_kernel, nb(dpad), 256, 0, s, d_sel, n_sel, dpad, m->minor.sum->as<double>(), m->minor.sq->as<double>(),
               n_cells, center ? 1 : 0, scale ? 1 : 0, sparse_shift_mode(), shift->as<double>(), inv_sd->as<double>(), wres->as<double>(),
               shf->as<float>(), isf->as<float>(), zc_h->as<__half>(), zc_l->as<__half>(), flag->as<uint32_t>());

    SRB_TRACE("lut + sel_stats");
    SRB_REQUIRE(n_sel < 65535, SRB_ERR_UNSUPPORTED, "at most 65534 selected features");
    Buf lut16 = dev_alloc(s, 2 * (M + 1)), shis = dev_alloc(s, 8 * dpad);
    SRB_LAUNCH(lut16_kernel, nb(M), 256, 0, s, lut->as<int>(), lut16->as<uint16_t>(), M);
    SRB_LAUNCH(shis_kernel, nb(dpad), 256, 0, s, shf->as<float>(), isf->as<float>(), shis->as<float2>(), dpad);
    // K6 panels
    Buf Xh = dev_alloc(s, 2 * (size_t)std::max<uint64_t>(n, 1) * dpad), Xl = dev_alloc(s, 2 * (size_t)std::max<uint64_t>(n, 1) * dpad);
    if (n) {
        StageTimer t(c, ST_DENSIFY);
        const size_t smem = 0;
        const unsigned grid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)c->sm_count * 8 * 4);
        static int pipe = -1;
        if (pipe < 0) {
            const char *e = getenv("SRB_DENSIFY_PIPE");
            pipe = (e && e[0] == '0') ? 0 : 1;  // default: register double-buffered variant (4.40 vs 4.68 ms at L)
        }
        static const int batch8 = [] {
            const char *e = getenv("SRB_DENSIFY_BATCH");
            return (e && atoi(e) == 4) ? 0 : 1;  // 8 loads per array in flight per lane: 4.30-4.38 ms against 4.64 for 4 (measured)
        }();
        if (m->vdtype == SRB_F32 && pipe && batch8) {
            SRB_LAUNCH((densify_panels_pipe_kernel<float, 8>), grid, 256, smem, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), lut16->as<uint16_t>(), shis->as<float2>(), zc_h->as<__half>(), zc_l->as<__half>(), n, dpad, Xh->as<__half>(), Xl->as<__half>());
        } else if (m->vdtype == SRB_F32 && pipe) {
            SRB_LAUNCH((densify_panels_pipe_kernel<float, 4>), grid, 256, smem, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), lut16->as<uint16_t>(), shis->as<float2>(), zc_h->as<__half>(), zc_l->as<__half>(), n, dpad, Xh->as<__half>(), Xl->as<__half>());
        } else if (m->vdtype == SRB_F32) {
            SRB_LAUNCH((densify_panels_kernel<float>), grid, 256, smem, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), lut16->as<uint16_t>(), shis->as<float2>(), zc_h->as<__half>(), zc_l->as<__half>(), n, dpad, Xh->as<__half>(), Xl->as<__half>());
        } else {
            SRB_LAUNCH((densify_panels_kernel<double>), grid, 256, smem, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<double>(), lut16->as<uint16_t>(), shis->as<float2>(), zc_h->as<__half>(), zc_l->as<__half>(), n, dpad, Xh->as<__half>(), Xl->as<__half>());
        }
    }
    SRB_TRACE("panels alloc + densify enqueue");
    // K7 Gram
    Buf G = dev_zeros(s, 8 * (size_t)dpad * dpad);
    {
        StageTimer t(c, ST_GRAM);
        if (gram_mode == 0) {
            gram_tcgen05(c, Xh->as<__half>(), Xl->as<__half>(), n, dpad, G->as<double>(), (uint32_t)n_sel);
        } else if (n) {
            const uint32_t nt = dpad / GT;
            const uint32_t ntiles = nt * (nt + 1) / 2;
            uint64_t splits = std::max<uint64_t>(1, std::min<uint64_t>((n + 511) / 512, (uint64_t)c->sm_count * 8 / ntiles + 1));
            const uint64_t rps = ((n + splits - 1) / splits + GK - 1) / GK * GK;
            splits = (n + rps - 1) / rps;
            SRB_LAUNCH(gram_f64_kernel, dim3(ntiles, (unsigned)splits), 256, 0, s, Xh->as<__half>(), Xl->as<__half>(), n, dpad, rps, G->as<double>());
            SRB_LAUNCH(gram_mirror_kernel, nb((uint64_t)dpad * dpad), 256, 0, s, G->as<double>(), dpad, (uint32_t)GT);
        }
    }
    SRB_TRACE("gram");
    if (c->nranks > 1) {
        StageTimer t(c, ST_ALLREDUCE_GRAM);
        allreduce_gram(c, G->as<double>(), dpad, n_sel);
    }
    // K8: correlation matrix + symmetric eigendecomposition
    Buf C = dev_alloc(s, 8 * n_sel * n_sel), evals = dev_alloc(s, 8 * n_sel), tr = dev_alloc(s, 8);
    Buf comps = dev_alloc(s, 8 * n_sel * k), W = dev_zeros(s, 8 * (size_t)dpad * kpad), evr = dev_alloc(s, 8 * k), bias = dev_zeros(s, 8 * kpad);
    {
        StageTimer t(c, ST_EIG);
        SRB_LAUNCH(corr_kernel, nb(n_sel * n_sel), 256, 0, s, G->as<double>(), dpad, n_sel, d_sel, m->minor.sum->as<double>(), m->minor.sq->as<double>(), shift->as<double>(), inv_sd->as<double>(), wres->as<double>(), n_cells, C->as<double>());
        SRB_LAUNCH(trace_kernel, 1, 256, 0, s, C->as<double>(), n_sel, tr->as<double>());
        const uint32_t npairs = sym_eig_desc(c, C->as<double>(), (uint32_t)n_sel, (uint32_t)k, evals->as<double>());
        SRB_LAUNCH(components_kernel, (unsigned)k, 256, 0, s, C->as<double>(), evals->as<double>(), npairs, n_sel, (uint32_t)k, kpad, tr->as<double>(), wres->as<double>(), comps->as<double>(), W->as<double>(), evr->as<double>(), bias->as<double>());
    }
    SRB_TRACE("eig");
    // K9 scores
    Buf scores = dev_alloc(s, 8 * std::max<uint64_t>(n, 1) * k);
    if (n) {
        StageTimer t(c, ST_SCORES);
        if (gram_mode == 0 && k <= 64) {
            scores_tcgen05(c, Xh->as<__half>(), Xl->as<__half>(), n, dpad, W->as<double>(), kpad, (uint32_t)k, bias->as<double>(), scores->as<double>());
        } else {
            const unsigned grid = (unsigned)std::min<uint64_t>((n + 31) / 32, (uint64_t)c->sm_count * 16);
            for (uint32_t c0 = 0; c0 < k; c0 += 64)
                SRB_LAUNCH(scores_simt_kernel, grid, 256, 0, s, Xh->as<__half>(), Xl->as<__half>(), n, dpad, W->as<double>(), kpad, (uint32_t)k, c0, bias->as<double>(), scores->as<double>());
        }
    }
    if (out.scores && n) SRB_CUDA(cudaMemcpyAsync(out.scores, scores->p, 8 * n * k, cudaMemcpyDeviceToHost, s));
    if (out.components) SRB_CUDA(cudaMemcpyAsync(out.components, comps->p, 8 * n_sel * k, cudaMemcpyDeviceToHost, s));
    if (out.evr) SRB_CUDA(cudaMemcpyAsync(out.evr, evr->p, 8 * k, cudaMemcpyDeviceToHost, s));
    SRB_TRACE("scores enqueue");
    uint32_t hflag = 0;
    SRB_CUDA(cudaMemcpyAsync(&hflag, flag->p, 4, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_TRACE("final sync");
    SRB_REQUIRE(!hflag, SRB_ERR_NAN, "a selected feature has zero variance over all cells while scale=true (the reference divides by zero: NaN scores)");
}

}  // namespace srb

// =====================================================================================================================
// Out-of-core PCA (BASELINE.json config 5): the same kernels, driven chunk by chunk for data that does not
// fit the GPUs' HBM. The reference has no chunked PCA (src/backed/processing/mod.rs is empty; its chunk drivers stop at
// number / sum, src/shared/statistics/mod.rs:17-83). Three passes over the row chunks, all arithmetic on the device:
//   1. per-gene moments of the transformed chunk (srb_gene_moments) — the caller adds the O(genes) vectors and picks the
//      features exactly like select_features does (dim_red/mod.rs:123-156);
//   2. srb_pca_stream_push_gram: standardise the chunk with the GLOBAL mean / std (K6) and add its Gram matrix (K7);
//      srb_pca_stream_fit: correlation matrix, eigenpairs (K8), loadings and explained-variance ratio;
//   3. srb_pca_stream_transform: scores of a chunk (K6 + K9).
// =====================================================================================================================
struct srb_pca_stream {
    srb_ctx *ctx = nullptr;
    uint64_t M = 0, n_sel = 0, k = 0, rows_seen = 0, ncells_total = 0;
    double n_cells = 0.0;
    bool center = true, scale = true, fitted = false;
    int gram_mode = 0;
    uint32_t dpad = 0, kpad = 0;
    srb::Buf sel, sum, sq, lut16, shis, zc_h, zc_l, shift, inv_sd, wres, bias, flag, G, W, comps, evr;
};

namespace srb {

__global__ void add_inplace_kernel(double *__restrict__ acc, const double *__restrict__ x, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc[i] += x[i];
}

// K6 for one chunk with the stream's global statistics: panels Xh / Xl (chunk rows x dpad)
static void stream_densify(srb_pca_stream *ps, srb_mat *m, Buf &Xh, Buf &Xl) {
    srb_ctx *c = ps->ctx;
    cudaStream_t s = c->stream;
    const uint64_t n = m->nrows;
    const uint32_t dpad = ps->dpad;
    Xh = dev_alloc(s, 2 * (size_t)std::max<uint64_t>(n, 1) * dpad), Xl = dev_alloc(s, 2 * (size_t)std::max<uint64_t>(n, 1) * dpad);
    if (!n) return;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 7) / 8, (uint64_t)c->sm_count * 8 * 4);
    if (m->vdtype == SRB_F32)
        SRB_LAUNCH((densify_panels_pipe_kernel<float, 4>), grid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), ps->lut16->as<uint16_t>(), ps->shis->as<float2>(), ps->zc_h->as<__half>(), ps->zc_l->as<__half>(), n, dpad, Xh->as<__half>(), Xl->as<__half>());
    else
        SRB_LAUNCH((densify_panels_kernel<double>), grid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<double>(), ps->lut16->as<uint16_t>(), ps->shis->as<float2>(), ps->zc_h->as<__half>(), ps->zc_l->as<__half>(), n, dpad, Xh->as<__half>(), Xl->as<__half>());
}

static void stream_check_chunk(const srb_pca_stream *ps, const srb_mat *m) {
    SRB_REQUIRE(ps && m && m->ctx && m->st, SRB_ERR_INVALID_ARG, "null handle");
    SRB_REQUIRE(ctx_alive(ps->ctx), SRB_ERR_INVALID_ARG, "the PCA stream outlived its context");
    SRB_REQUIRE(m->ctx == ps->ctx, SRB_ERR_INVALID_ARG, "chunk and PCA stream belong to different contexts");
    SRB_REQUIRE(m->format == SRB_CSR, SRB_ERR_UNSUPPORTED, "out-of-core PCA takes CSR row chunks");
    SRB_REQUIRE(m->ncols == ps->M, SRB_ERR_INVALID_ARG, "chunk has a different number of genes");
}

}  // namespace srb

using namespace srb;

extern "C" {

int32_t srb_pca_stream_begin(srb_ctx *ctx, uint64_t ncols, uint64_t ncells_total, const double *gene_sum, const double *gene_sumsq,
                             const uint64_t *col_sel, uint64_t n_sel, uint64_t k, int32_t center, int32_t scale, int32_t gram_mode,
                             srb_pca_stream **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx && out && gene_sum && gene_sumsq && col_sel, SRB_ERR_INVALID_ARG, "null argument");
    SRB_REQUIRE(gram_mode == 0 || gram_mode == 1, SRB_ERR_INVALID_ARG, "gram_mode must be 0 (tcgen05) or 1 (fp64)");
    SRB_REQUIRE(ncells_total >= 2, SRB_ERR_INVALID_ARG, "PCA needs at least two cells");
    SRB_REQUIRE(n_sel >= 1 && n_sel < 65535 && k >= 1 && k <= n_sel, SRB_ERR_INVALID_ARG, "need 1 <= k <= n_sel < 65535");
    SRB_REQUIRE(ncols < (1ull << 32), SRB_ERR_INVALID_ARG, "too many genes");
    SRB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::vector<uint32_t> h(n_sel);
    for (uint64_t i = 0; i < n_sel; ++i) {
        SRB_REQUIRE(col_sel[i] < ncols, SRB_ERR_INDEX_OOB, "Index out of bounds");
        h[i] = (uint32_t)col_sel[i];
    }
    std::unique_ptr<srb_pca_stream> ps(new srb_pca_stream());
    ps->ctx = ctx, ps->M = ncols, ps->n_sel = n_sel, ps->k = k, ps->n_cells = (double)ncells_total, ps->ncells_total = ncells_total;
    ps->center = center != 0, ps->scale = scale != 0, ps->gram_mode = gram_mode;
    const uint32_t dpad = ps->dpad = (uint32_t)((n_sel + 255) / 256 * 256);
    ps->kpad = (uint32_t)((k + 63) / 64 * 64);
    const uint64_t M = ncols;
    ps->sel = dev_alloc(s, 4 * n_sel), ps->sum = dev_alloc(s, 8 * (M + 1)), ps->sq = dev_alloc(s, 8 * (M + 1));
    SRB_CUDA(cudaMemcpyAsync(ps->sel->p, h.data(), 4 * n_sel, cudaMemcpyHostToDevice, s));
    if (M) {
        SRB_CUDA(cudaMemcpyAsync(ps->sum->p, gene_sum, 8 * M, cudaMemcpyHostToDevice, s));
        SRB_CUDA(cudaMemcpyAsync(ps->sq->p, gene_sumsq, 8 * M, cudaMemcpyHostToDevice, s));
    }
    Buf lut = dev_alloc(s, 4 * (M + 1));
    SRB_LAUNCH(lut_fill_kernel, nb(M), 256, 0, s, lut->as<int>(), M);
    SRB_LAUNCH(lut_set_kernel, nb(n_sel), 256, 0, s, lut->as<int>(), ps->sel->as<uint32_t>(), n_sel);
    ps->flag = dev_zeros(s, 4);
    ps->shift = dev_alloc(s, 8 * dpad), ps->inv_sd = dev_alloc(s, 8 * dpad), ps->zc_h = dev_alloc(s, 2 * dpad), ps->zc_l = dev_alloc(s, 2 * dpad);
    Buf shf = dev_alloc(s, 4 * dpad), isf = dev_alloc(s, 4 * dpad);
    ps->wres = dev_alloc(s, 8 * dpad);
    SRB_LAUNCH(sel_stats_kernel, nb(dpad), 256, 0, s, ps->sel->as<uint32_t>(), n_sel, dpad, ps->sum->as<double>(), ps->sq->as<double>(),
               ps->n_cells, ps->center ? 1 : 0, ps->scale ? 1 : 0, sparse_shift_mode(), ps->shift->as<double>(), ps->inv_sd->as<double>(),
               ps->wres->as<double>(), shf->as<float>(), isf->as<float>(), ps->zc_h->as<__half>(), ps->zc_l->as<__half>(), ps->flag->as<uint32_t>());
    ps->lut16 = dev_alloc(s, 2 * (M + 1)), ps->shis = dev_alloc(s, 8 * dpad);
    SRB_LAUNCH(lut16_kernel, nb(M), 256, 0, s, lut->as<int>(), ps->lut16->as<uint16_t>(), M);
    SRB_LAUNCH(shis_kernel, nb(dpad), 256, 0, s, shf->as<float>(), isf->as<float>(), ps->shis->as<float2>(), dpad);
    ps->G = dev_zeros(s, 8 * (size_t)dpad * dpad);
    SRB_CUDA(cudaStreamSynchronize(s));  // the host staging vectors are done with
    *out = ps.release();
    SRB_API_END
}

int32_t srb_pca_stream_push_gram(srb_pca_stream *ps, srb_mat *chunk) {
    SRB_API_BEGIN
    stream_check_chunk(ps, chunk);
    SRB_REQUIRE(!ps->fitted, SRB_ERR_INVALID_ARG, "push_gram after fit");
    srb_ctx *c = ps->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    materialize(chunk, false);  // pending normalise / log1p are applied to the chunk's values
    const uint64_t n = chunk->nrows;
    if (n) {
        const uint32_t dpad = ps->dpad;
        Buf Xh, Xl;
        stream_densify(ps, chunk, Xh, Xl);
        Buf Gc = dev_zeros(s, 8 * (size_t)dpad * dpad);
        // a chunk smaller than one tensor-core tile of cells goes through the fp64 CUDA-core Gram (exact, and tiny)
        if (ps->gram_mode == 0 && n >= 256) {
            gram_tcgen05(c, Xh->as<__half>(), Xl->as<__half>(), n, dpad, Gc->as<double>(), (uint32_t)ps->n_sel);
        } else {
            const uint32_t nt = dpad / GT;
            const uint32_t ntiles = nt * (nt + 1) / 2;
            uint64_t splits = std::max<uint64_t>(1, std::min<uint64_t>((n + 511) / 512, (uint64_t)c->sm_count * 8 / ntiles + 1));
            const uint64_t rps = ((n + splits - 1) / splits + GK - 1) / GK * GK;
            splits = (n + rps - 1) / rps;
            SRB_LAUNCH(gram_f64_kernel, dim3(ntiles, (unsigned)splits), 256, 0, s, Xh->as<__half>(), Xl->as<__half>(), n, dpad, rps, Gc->as<double>());
            SRB_LAUNCH(gram_mirror_kernel, nb((uint64_t)dpad * dpad), 256, 0, s, Gc->as<double>(), dpad, (uint32_t)GT);
        }
        SRB_LAUNCH(add_inplace_kernel, (unsigned)std::min<uint64_t>(((uint64_t)dpad * dpad + 255) / 256, (uint64_t)c->sm_count * 16), 256, 0, s,
                   ps->G->as<double>(), Gc->as<double>(), (uint64_t)dpad * dpad);
        ps->rows_seen += n;
    }
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_API_END
}

int32_t srb_pca_stream_fit(srb_pca_stream *ps, double *components, double *explained_variance_ratio) {
    SRB_API_BEGIN
    SRB_REQUIRE(ps, SRB_ERR_INVALID_ARG, "null handle");
    SRB_REQUIRE(!ps->fitted, SRB_ERR_INVALID_ARG, "already fitted");
    srb_ctx *c = ps->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint64_t n_sel = ps->n_sel, k = ps->k;
    const uint32_t dpad = ps->dpad, kpad = ps->kpad;
    {
        // every cell must have been pushed exactly once (over all ranks): the diagonal of C comes from the caller's global
        // moments, the off-diagonals from the pushed chunks, and the two must describe the same cells
        uint64_t seen = ps->rows_seen;
        if (c->nranks > 1) {
            Buf d_seen = dev_alloc(s, 8);
            SRB_CUDA(cudaMemcpyAsync(d_seen->p, &seen, 8, cudaMemcpyHostToDevice, s));
            allreduce_u64_sum(c, d_seen->as<uint64_t>(), 1);
            SRB_CUDA(cudaMemcpyAsync(&seen, d_seen->p, 8, cudaMemcpyDeviceToHost, s));
            SRB_CUDA(cudaStreamSynchronize(s));
        }
        SRB_REQUIRE(seen == ps->ncells_total, SRB_ERR_INVALID_ARG,
                    "srb_pca_stream_fit: " + std::to_string(seen) + " cells were pushed but the stream was opened for " + std::to_string(ps->ncells_total));
    }
    if (c->nranks > 1) allreduce_gram(c, ps->G->as<double>(), dpad, n_sel);  // ranks hold disjoint chunks
    Buf C = dev_alloc(s, 8 * n_sel * n_sel), evals = dev_alloc(s, 8 * n_sel), tr = dev_alloc(s, 8);
    ps->comps = dev_alloc(s, 8 * n_sel * k), ps->W = dev_zeros(s, 8 * (size_t)dpad * kpad), ps->evr = dev_alloc(s, 8 * k);
    ps->bias = dev_zeros(s, 8 * kpad);
    SRB_LAUNCH(corr_kernel, nb(n_sel * n_sel), 256, 0, s, ps->G->as<double>(), dpad, n_sel, ps->sel->as<uint32_t>(), ps->sum->as<double>(),
               ps->sq->as<double>(), ps->shift->as<double>(), ps->inv_sd->as<double>(), ps->wres->as<double>(), ps->n_cells, C->as<double>());
    SRB_LAUNCH(trace_kernel, 1, 256, 0, s, C->as<double>(), n_sel, tr->as<double>());
    const uint32_t npairs = sym_eig_desc(c, C->as<double>(), (uint32_t)n_sel, (uint32_t)k, evals->as<double>());
    SRB_LAUNCH(components_kernel, (unsigned)k, 256, 0, s, C->as<double>(), evals->as<double>(), npairs, n_sel, (uint32_t)k, kpad, tr->as<double>(),
               ps->wres->as<double>(), ps->comps->as<double>(), ps->W->as<double>(), ps->evr->as<double>(), ps->bias->as<double>());
    if (components) SRB_CUDA(cudaMemcpyAsync(components, ps->comps->p, 8 * n_sel * k, cudaMemcpyDeviceToHost, s));
    if (explained_variance_ratio) SRB_CUDA(cudaMemcpyAsync(explained_variance_ratio, ps->evr->p, 8 * k, cudaMemcpyDeviceToHost, s));
    uint32_t hflag = 0;
    SRB_CUDA(cudaMemcpyAsync(&hflag, ps->flag->p, 4, cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_REQUIRE(!hflag, SRB_ERR_NAN, "a selected feature has zero variance over all cells while scale=true (the reference divides by zero: NaN scores)");
    ps->G.reset();
    ps->fitted = true;
    SRB_API_END
}

int32_t srb_pca_stream_transform(srb_pca_stream *ps, srb_mat *chunk, double *scores) {
    SRB_API_BEGIN
    stream_check_chunk(ps, chunk);
    SRB_REQUIRE(ps->fitted, SRB_ERR_INVALID_ARG, "transform before fit");
    SRB_REQUIRE(scores || chunk->nrows == 0, SRB_ERR_INVALID_ARG, "scores is null");
    srb_ctx *c = ps->ctx;
    SRB_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    materialize(chunk, false);
    const uint64_t n = chunk->nrows, k = ps->k;
    if (n) {
        Buf Xh, Xl;
        stream_densify(ps, chunk, Xh, Xl);
        Buf out = dev_alloc(s, 8 * n * k);
        if (ps->gram_mode == 0 && k <= 64 && n >= 256) {
            scores_tcgen05(c, Xh->as<__half>(), Xl->as<__half>(), n, ps->dpad, ps->W->as<double>(), ps->kpad, (uint32_t)k, ps->bias->as<double>(), out->as<double>());
        } else {
            const unsigned grid = (unsigned)std::min<uint64_t>((n + 31) / 32, (uint64_t)c->sm_count * 16);
            for (uint32_t c0 = 0; c0 < k; c0 += 64)
                SRB_LAUNCH(scores_simt_kernel, grid, 256, 0, s, Xh->as<__half>(), Xl->as<__half>(), n, ps->dpad, ps->W->as<double>(), ps->kpad, (uint32_t)k, c0, ps->bias->as<double>(), out->as<double>());
        }
        SRB_CUDA(cudaMemcpyAsync(scores, out->p, 8 * n * k, cudaMemcpyDeviceToHost, s));
        SRB_CUDA(cudaStreamSynchronize(s));
    }
    SRB_API_END
}

int32_t srb_pca_stream_free(srb_pca_stream *ps) {
    SRB_API_BEGIN
    if (ps) {
        if (ctx_alive(ps->ctx)) {
            cudaSetDevice(ps->ctx->device);
            cudaStreamSynchronize(ps->ctx->stream);
        }
        delete ps;
    }
    SRB_API_END
}

}  // extern "C"
