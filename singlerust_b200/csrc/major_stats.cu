// major_stats.cu — K1: segmented reductions along the MAJOR axis of a compressed matrix (per cell on CSR,
// per gene on CSC). Replaces the Row branches of src/shared/statistics/helper/csr.rs (sum 87-93, variance
// 158-170, min/max 200-210) and the Column branches of helper/csc.rs.
//
// HBM-bound: 4 B/nnz (f32 values) + 8 B/line of offsets; column indices are never read.
// Mapping: LPR lanes cooperate on one line (LPR = 32 for long lines, 8 for short ones), coalesced strided
// loads with 4 independent loads in flight per lane, f64 accumulation, shuffle tree at the end.
#include <algorithm>
#include <cstdlib>

#include "bulk.cuh"
#include "common.cuh"

namespace srb {

enum { MODE_SUM_ABSMAX = 0, MODE_VARIANCE = 1, MODE_MINMAX = 2 };

template <int LPR>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int LPR>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int LPR>
__device__ __forceinline__ double group_min(double v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename VT, int LPR, int MODE>
__global__ void __launch_bounds__(256) major_reduce_kernel(const int64_t *__restrict__ off, const VT *__restrict__ val,
                                                           uint64_t nmajor, double *__restrict__ o0,
                                                           double *__restrict__ o1, double *__restrict__ o2,
                                                           uint32_t *__restrict__ flags) {
    const int lane = threadIdx.x % LPR;
    const uint64_t group = ((uint64_t)blockIdx.x * 256 + threadIdx.x) / LPR;
    const uint64_t ngroups = (uint64_t)gridDim.x * 256 / LPR;
    const uint64_t iters = (nmajor + ngroups - 1) / ngroups;  // uniform trip count: shuffles stay convergent
    uint32_t bad = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        const uint64_t i = group + it * ngroups;
        const bool act = i < nmajor;
        const int64_t a = act ? off[i] : 0;
        const int64_t b = act ? off[i + 1] : 0;
        if (MODE == MODE_SUM_ABSMAX || MODE == MODE_VARIANCE) {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0, mx = 0, mnz = INFINITY;
            int64_t k = a + lane;
            if (MODE == MODE_SUM_ABSMAX && sizeof(VT) == 4) {
                // f32 fast path: range / sign / finiteness / integrality tracked on the raw bit patterns (integer
                // min/max and one OR per element) so that the loop stays bandwidth- rather than issue-bound;
                // 16 independent loads in flight per lane.
                uint32_t orbits = 0, maxab = 0, minab1 = 0xFFFFFFFFu, nonint = 0;
                const float *fv = reinterpret_cast<const float *>(val);
                constexpr int U = 16;
                for (; k < b; k += U * LPR) {
                    float v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) v[u] = (k + u * LPR < b) ? fv[k + u * LPR] : 0.f;
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t bits = __float_as_uint(v[u]), ab = bits & 0x7FFFFFFFu;
                        orbits |= bits;
                        maxab = max(maxab, ab);
                        minab1 = min(minab1, ab - 1u);  // zeros wrap to 0xFFFFFFFF and drop out of the minimum
                        nonint |= (uint32_t)(v[u] != truncf(v[u]));
                    }
#pragma unroll
                    for (int u = 0; u < U; u += 4) {
                        s0 += (double)v[u], s1 += (double)v[u + 1], s2 += (double)v[u + 2], s3 += (double)v[u + 3];
                    }
                }
                mx = (double)__uint_as_float(maxab);
                mnz = (minab1 == 0xFFFFFFFFu) ? INFINITY : (double)__uint_as_float(minab1 + 1u);
                bad |= (orbits >> 31) | (2u * (uint32_t)(maxab >= 0x7F800000u)) | (4u * nonint);
            } else {
            for (; k + 3 * LPR < b; k += 4 * LPR) {
                const double v0 = (double)val[k], v1 = (double)val[k + LPR], v2 = (double)val[k + 2 * LPR],
                             v3 = (double)val[k + 3 * LPR];
                s0 += v0, s1 += v1, s2 += v2, s3 += v3;
                if (MODE == MODE_SUM_ABSMAX) {
                    mx = fmax(fmax(mx, fabs(v0)), fmax(fabs(v1), fmax(fabs(v2), fabs(v3))));
                    mnz = fmin(fmin(mnz, v0 != 0 ? fabs(v0) : INFINITY), fmin(v1 != 0 ? fabs(v1) : INFINITY, fmin(v2 != 0 ? fabs(v2) : INFINITY, v3 != 0 ? fabs(v3) : INFINITY)));
                    bad |= (v0 < 0) | (v1 < 0) | (v2 < 0) | (v3 < 0);
                    bad |= 2u * (!isfinite(v0) | !isfinite(v1) | !isfinite(v2) | !isfinite(v3));
                    bad |= 4u * ((v0 != rint(v0)) | (v1 != rint(v1)) | (v2 != rint(v2)) | (v3 != rint(v3)));
                }
            }
            for (; k < b; k += LPR) {
                const double v0 = (double)val[k];
                s0 += v0;
                if (MODE == MODE_SUM_ABSMAX) {
                    mx = fmax(mx, fabs(v0));
                    mnz = fmin(mnz, v0 != 0 ? fabs(v0) : INFINITY);
                    bad |= (v0 < 0) | (2u * !isfinite(v0)) | (4u * (v0 != rint(v0)));
                }
            }
            }
            const double sum = group_sum<LPR>((s0 + s1) + (s2 + s3));
            if (MODE == MODE_SUM_ABSMAX) {
                mx = group_max<LPR>(mx);
                mnz = group_min<LPR>(mnz);
                if (act && lane == 0) {
                    o0[i] = sum;
                    o1[i] = mx;
                    o2[i] = mnz;
                }
            } else {
                // variance_whole_helper major branch: mean = sum/count; sum((v-mean)^2)/count; 0/0 -> NaN
                const double cnt = (double)(uint32_t)(b - a);
                const double mean = sum / cnt;
                double q0 = 0, q1 = 0;
                int64_t k2 = a + lane;
                for (; k2 + LPR < b; k2 += 2 * LPR) {
                    const double d0 = (double)val[k2] - mean, d1 = (double)val[k2 + LPR] - mean;
                    q0 += d0 * d0, q1 += d1 * d1;
                }
                for (; k2 < b; k2 += LPR) {
                    const double d0 = (double)val[k2] - mean;
                    q0 += d0 * d0;
                }
                const double ss = group_sum<LPR>(q0 + q1);
                if (act && lane == 0) o0[i] = ss / cnt;
            }
        } else {
            double mn = INFINITY, mx = -INFINITY;
            for (int64_t k = a + lane; k < b; k += LPR) {
                const double v = (double)val[k];
                mn = fmin(mn, v), mx = fmax(mx, v);
            }
            mn = group_min<LPR>(mn), mx = group_max<LPR>(mx);
            if (act && lane == 0) {
                o0[i] = mn;
                o1[i] = mx;
            }
        }
    }
    if (MODE == MODE_SUM_ABSMAX && bad) {
        if (bad & 1u) atomicOr(flags, 1u);
        if (bad & 2u) atomicOr(flags + 1, 1u);
        if (bad & 4u) atomicOr(flags + 2, 1u);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K1, bulk-staged form (f32 values, sum / range / flags): the contiguous run of values that belongs to a CTA's rows is
// streamed tile by tile (2048 values = 8 KB) into a 4-stage shared-memory ring by cp.async.bulk (one producer thread; the
// copy engine keeps up to 32 KB per CTA, ~220 KB per SM, in flight without a register or an issue slot of the reducing
// warps), and two consumer warps reduce whole rows out of shared memory: lane-strided conflict-free reads, f64 partial
// sums per lane, warp-shuffle / redux at the end of the row. A tile holds ~1.4 rows of the bench matrix, so the two
// consumer warps of a CTA work on neighbouring rows of the same tile; parallelism comes from 7 resident CTAs per SM.
// Every consumer warp waits for and releases every tile exactly once, in order (rows may span tiles).
// ---------------------------------------------------------------------------------------------------------------------
namespace k1b {
constexpr int TILE = 2048, STAGES = 4;
}

template <int CONSUMERS>
__global__ void __launch_bounds__(32 * (1 + CONSUMERS)) major_sum_bulk_kernel(const int64_t *__restrict__ off, const float *__restrict__ val,
                                                                      uint64_t nmajor, uint32_t rows_per_cta, double *__restrict__ o_sum,
                                                                      double *__restrict__ o_max, double *__restrict__ o_min,
                                                                      uint32_t *__restrict__ flags) {
    using namespace k1b;
    constexpr int THREADS = 32 * (1 + CONSUMERS);
    (void)THREADS;
    __shared__ __align__(128) float tile[STAGES][TILE];
    __shared__ __align__(8) uint64_t bars[2 * STAGES];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t r0 = (uint64_t)blockIdx.x * rows_per_cta;
    if (r0 >= nmajor) return;
    const uint64_t r1 = min(r0 + (uint64_t)rows_per_cta, nmajor);
    const int64_t a0 = off[r0], b1 = off[r1];
    const int64_t base = a0 & ~(int64_t)3;  // 16-byte aligned start of the CTA's run
    const uint32_t ntiles = (uint32_t)((b1 - base + TILE - 1) / TILE);
    const uint32_t full0 = bulk::smem_u32(&bars[0]), empty0 = bulk::smem_u32(&bars[STAGES]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) bulk::mbar_init(full0 + 8 * s, 1), bulk::mbar_init(empty0 + 8 * s, CONSUMERS);
        bulk::mbar_init_fence();
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            for (uint32_t t = 0; t < ntiles; ++t) {
                const uint32_t s = t % STAGES;
                bulk::mbar_wait(empty0 + 8 * s, ((t / STAGES) & 1) ^ 1);
                const int64_t lo = base + (int64_t)t * TILE;
                const uint32_t n = (uint32_t)min((int64_t)TILE, (b1 - lo + 3) & ~(int64_t)3);  // whole 16-byte units
                bulk::mbar_arrive_expect_tx(full0 + 8 * s, 4 * n);
                bulk::copy_g2s(bulk::smem_u32(&tile[s][0]), val + lo, 4 * n, full0 + 8 * s);
            }
        }
        return;
    }
    const uint32_t cw = warp - 1;
    uint32_t t_cur = 0, bad = 0;
    bool have = false;  // tile t_cur has been waited for
    auto release_until = [&](uint32_t t) {  // wait for and release the tiles before t
        while (t_cur < t) {
            if (!have) bulk::mbar_wait(full0 + 8 * (t_cur % STAGES), (t_cur / STAGES) & 1);
            __syncwarp();
            if (lane == 0) bulk::mbar_arrive(empty0 + 8 * (t_cur % STAGES));
            ++t_cur, have = false;
        }
    };
    for (uint64_t r = r0 + cw; r < r1; r += CONSUMERS) {
        const int64_t a = off[r], b = off[r + 1];
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        uint32_t orbits = 0, maxab = 0, minab1 = 0xFFFFFFFFu, fracbits = 0;
        int64_t pos = a;
        while (pos < b) {
            const uint32_t t = (uint32_t)((pos - base) / TILE);
            release_until(t);
            if (!have) bulk::mbar_wait(full0 + 8 * (t % STAGES), (t / STAGES) & 1), have = true;
            const int64_t tile_lo = base + (int64_t)t * TILE;
            const int e = (int)(min(b, tile_lo + TILE) - tile_lo);
            const float *tp = tile[t % STAGES];
            int i = (int)(pos - tile_lo) + (int)lane;
            // range / sign / finiteness / integrality on the raw bit patterns, two elements per 3-input instruction
            // (LOP3, VIMNMX3): |v| = bits & 0x7FFFFFFF; zeros wrap to 0xFFFFFFFF in (|v| - 1) and drop out of the minimum;
            // v - trunc(v) has a non-zero pattern exactly for fractions, NaN and inf
#define SRB_K1_PAIR(VA, VB)                                                                              \
    do {                                                                                                 \
        const uint32_t ba = __float_as_uint(VA), bb = __float_as_uint(VB);                               \
        const uint32_t aa = ba & 0x7FFFFFFFu, ab = bb & 0x7FFFFFFFu;                                     \
        orbits |= ba | bb;                                                                               \
        maxab = __vimax3_u32(maxab, aa, ab);                                                             \
        minab1 = __vimin3_u32(minab1, aa - 1u, ab - 1u);                                                 \
        fracbits |= __float_as_uint((VA) - truncf(VA)) | __float_as_uint((VB) - truncf(VB));             \
    } while (0)
            for (; i + 96 < e; i += 128) {
                const float v0 = tp[i], v1 = tp[i + 32], v2 = tp[i + 64], v3 = tp[i + 96];
                SRB_K1_PAIR(v0, v1);
                SRB_K1_PAIR(v2, v3);
                s0 += (double)v0, s1 += (double)v1, s2 += (double)v2, s3 += (double)v3;
            }
            for (; i < e; i += 32) {
                const float v0 = tp[i];
                SRB_K1_PAIR(v0, v0);
                s0 += (double)v0;
            }
#undef SRB_K1_PAIR
            pos = tile_lo + e;
        }
        const double sum = warp_sum((s0 + s1) + (s2 + s3));
        maxab = __reduce_max_sync(0xffffffffu, maxab);
        minab1 = __reduce_min_sync(0xffffffffu, minab1);
        orbits = __reduce_or_sync(0xffffffffu, orbits);
        const uint32_t nonint = (__reduce_or_sync(0xffffffffu, fracbits) & 0x7FFFFFFFu) ? 1u : 0u;
        if (lane == 0) {
            o_sum[r] = sum;
            o_max[r] = (double)__uint_as_float(maxab);
            o_min[r] = (minab1 == 0xFFFFFFFFu) ? INFINITY : (double)__uint_as_float(minab1 + 1u);
        }
        bad |= (orbits >> 31) | (2u * (uint32_t)(maxab >= 0x7F800000u)) | (4u * nonint);
    }
    release_until(ntiles);
    if (lane == 0 && bad) {
        if (bad & 1u) atomicOr(flags, 1u);
        if (bad & 2u) atomicOr(flags + 1, 1u);
        if (bad & 4u) atomicOr(flags + 2, 1u);
    }
}

template <int MODE>
static void launch_major(srb_mat *m, double *o0, double *o1, double *o2, uint32_t *flags) {
    srb_ctx *c = m->ctx;
    const uint64_t n = m->nmajor();
    if (n == 0) return;
    const double avg = (double)m->st->nnz / (double)n;
    const bool wide = avg >= 96.0;
    const int lpr = wide ? 32 : 8;
    const uint64_t groups_per_cta = 256 / lpr;
    uint64_t grid = (n + groups_per_cta - 1) / groups_per_cta;
    const uint64_t cap = (uint64_t)c->sm_count * 8 * 4;  // 8 resident CTAs/SM, 4 waves
    if (grid > cap) grid = cap;
    const int64_t *off = m->st->offsets->as<int64_t>();
#define GO(VT, L) SRB_LAUNCH((major_reduce_kernel<VT, L, MODE>), (unsigned)grid, 256, 0, c->stream, off, m->values->as<VT>(), n, o0, o1, o2, flags)
    if (m->vdtype == SRB_F32) {
        if (wide) GO(float, 32); else GO(float, 8);
    } else {
        if (wide) GO(double, 32); else GO(double, 8);
    }
#undef GO
}

void major_sum_absmax(srb_mat *m) {
    if (m->has_pending()) materialize(m, false);
    if (m->major.valid) return;
    srb_ctx *c = m->ctx;
    cudaStream_t st = c->stream;
    const uint64_t n = m->nmajor();
    m->major.sum = dev_zeros(st, sizeof(double) * (n + 1));
    m->major.absmax = dev_zeros(st, sizeof(double) * (n + 1));
    m->major.absmin = dev_zeros(st, sizeof(double) * (n + 1));
    m->major.flags = dev_zeros(st, sizeof(uint32_t) * 4);
    static const int bulk_on = [] {
        const char *e = getenv("SRB_K1_BULK");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    // bulk-staged kernel: f32 storage and lines long enough that a tile holds only a few of them
    if (bulk_on && n && m->vdtype == SRB_F32 && (double)m->st->nnz / (double)n >= 128.0) {
        uint64_t rpc = (n + (uint64_t)c->sm_count * 28 - 1) / ((uint64_t)c->sm_count * 28);
        rpc = std::max<uint64_t>(16, std::min<uint64_t>(rpc, 4096));
        const unsigned grid = (unsigned)((n + rpc - 1) / rpc);
        static const int consumers = [] {
            const char *e = getenv("SRB_K1_CONSUMERS");
            return e ? atoi(e) : 4;  // measured at the bench size: 2 -> 1.36 ms, 3 -> 1.15, 4 -> 1.05 (0.88 of the HBM peak), 6 -> 1.07, 8 -> 1.10
        }();
        if (consumers == 4)
            SRB_LAUNCH(major_sum_bulk_kernel<4>, grid, 160, 0, st, m->st->offsets->as<int64_t>(), m->values->as<float>(), n, (uint32_t)rpc,
                       m->major.sum->as<double>(), m->major.absmax->as<double>(), m->major.absmin->as<double>(), m->major.flags->as<uint32_t>());
        else
            SRB_LAUNCH(major_sum_bulk_kernel<2>, grid, 96, 0, st, m->st->offsets->as<int64_t>(), m->values->as<float>(), n, (uint32_t)rpc,
                       m->major.sum->as<double>(), m->major.absmax->as<double>(), m->major.absmin->as<double>(), m->major.flags->as<uint32_t>());
    } else {
        launch_major<MODE_SUM_ABSMAX>(m, m->major.sum->as<double>(), m->major.absmax->as<double>(), m->major.absmin->as<double>(),
                                      m->major.flags->as<uint32_t>());
    }
    m->major.valid = true;
}

void major_variance(srb_mat *m, double *d_out) {
    if (m->has_pending()) materialize(m, false);
    launch_major<MODE_VARIANCE>(m, d_out, nullptr, nullptr, nullptr);
}

void major_min_max(srb_mat *m, double *d_min, double *d_max) {
    if (m->has_pending()) materialize(m, false);
    launch_major<MODE_MINMAX>(m, d_min, d_max, nullptr, nullptr);
}

}  // namespace srb
