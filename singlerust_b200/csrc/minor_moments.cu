// minor_moments.cu — K2/K3/K4: per-MINOR-line moments (per gene on CSR) and the deferred value transforms
// (total-count scale, log1p), fused into one pass over the nnz.
//
// Replaces: helper/csr.rs Column branches (number 29-36, sum 94-100, variance 172-186) and helper/csc.rs Row
// branches; scale/mod.rs:59-89,141-173 (value scaling) and transform/mod.rs:36-56 (log1p).
//
// Why fixed point: per-gene sums are a scatter-reduce into ~30 k bins. On sm_100a only 32-bit INTEGER
// shared-memory atomics are native (ATOMS.ADD); f32/f64/u64 shared atomics compile to CAS spin loops and global
// fp64 REDs run at <1 lane/clk/SM — both far from the HBM roofline. So every value is quantised once to
// q = rint(v * 2^F) < 2^28 (F chosen on the device from the exact max |v|), and count / sum(q) / sum(q^2) are
// accumulated EXACTLY in shared-memory integer limbs with carry propagation, flushed as 32-bit limbs into
// 64-bit global accumulators (no carries needed there), and combined at the end. Integer addition is
// associative, so the moments are bit-identical for any grid shape, any shard count and any GPU count.
// Quantisation error: |v - q 2^-F| <= 2^-(F+1) <= max|v| * 2^-28 per value (for f32 data of similar magnitude
// this is below the f32 ulp), unbiased; see DESIGN.md for the bound on the variance.
//
// Shared-memory bins: 4 words / gene  [sum_lo | sq_lo | sq_mid | packed(count:14, sum_hi:12, sq_hi:6)]
// => 16 B/gene; genes are split into S column stripes so a stripe fits in <= ~160 KB; a CTA owns one
// (row block, stripe) and reads only the contiguous part of each row that falls into its stripe (column
// indices are sorted within a row; split points are found once per structure by binary search).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace srb {

static constexpr int kFusedThreads = 1024;
static constexpr uint32_t kMaxRowsPerBlock = 8192;   // packed count field (14 bits) and carry fields
static constexpr size_t kBinSmemBudget = 160 * 1024;  // per CTA
static constexpr int kMaxStripes = 8;

// ---------------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------------
__global__ void splits_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx, uint64_t nmajor,
                              int S, uint32_t W, int64_t *__restrict__ splits) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t total = nmajor * (uint64_t)(S - 1);
    if (t >= total) return;
    const uint64_t r = t % nmajor;
    const int s = (int)(t / nmajor);  // boundary between stripe s and s+1
    const uint32_t bound = (uint32_t)(s + 1) * W;
    int64_t lo = off[r], hi = off[r + 1];
    while (lo < hi) {  // first k with idx[k] >= bound
        const int64_t mid = (lo + hi) >> 1;
        if (idx[mid] < bound) lo = mid + 1; else hi = mid;
    }
    splits[(uint64_t)s * nmajor + r] = lo;
}

// scale[i] = 0 if sum == 0 else target / sum   (scale/mod.rs:9-15, 93-99)
__global__ void line_scale_kernel(const double *__restrict__ sum, uint64_t n, double target, double *__restrict__ scale) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scale[i] = (sum[i] == 0.0) ? 0.0 : target / sum[i];
}

// range[0] = max_i amax[i]*|b[i]| ; range[1] = min_i amin[i]*|b[i]| over the strictly positive products (b null => 1).
// Non-negative doubles order like their bit patterns, so integer atomics do the reduction.
__global__ void range_init_kernel(unsigned long long *range) {
    range[0] = 0ULL;
    range[1] = 0x7FF0000000000000ULL;  // +inf
}
__global__ void range_prod_kernel(const double *__restrict__ amax, const double *__restrict__ amin,
                                  const double *__restrict__ b, uint64_t n, unsigned long long *__restrict__ range) {
    double mx = 0.0, mn = INFINITY;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const double sc = b ? fabs(b[i]) : 1.0;
        const double hi = fabs(amax[i]) * sc, lo = fabs(amin[i]) * sc;
        if (hi == hi) mx = fmax(mx, hi);
        else mx = INFINITY;  // NaN => unbounded
        if (lo > 0.0) mn = fmin(mn, lo);
    }
    mx = warp_max(mx), mn = warp_min(mn);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(range, (unsigned long long)__double_as_longlong(mx));
        atomicMin(range + 1, (unsigned long long)__double_as_longlong(mn));
    }
}
__global__ void range_mul_kernel(double *out, const double *a, const double *b) {
    out[0] = a[0] * b[0];
    out[1] = a[1] * b[1];
}

// F such that rint(bound' * 2^F) < 2^28, bound' = bound (or log1p(bound)) with a safety margin
__global__ void fexp_kernel(const double *__restrict__ bound, int do_log1p, int *__restrict__ fexp) {
    double b = bound[0];
    if (do_log1p) b = log1p(b);
    b *= 1.0 + 1e-6;
    int e = 0;
    if (b > 0.0 && isfinite(b)) e = ilogb(b) + 1;  // b < 2^e
    int F = 28 - e;
    if (F > 100) F = 100;
    if (F < -900) F = -900;
    fexp[0] = F;
}

// ---------------------------------------------------------------------------------------------------
// value transform shared by all kernels: x = v * scale ; optional log1p. COMPACT (float) and FAITHFUL (double)
// ---------------------------------------------------------------------------------------------------
template <typename VTO>
struct Xform;
template <>
struct Xform<float> {
    // f32 pipeline: the scale is rounded once to f32 (rel 6e-8), log1pf is <= 1 ulp
    template <typename VTI>
    static __device__ __forceinline__ float apply(VTI v, double sc, bool has_scale, bool lg) {
        float x = (float)v;
        if (has_scale) x *= (float)sc;
        if (lg) x = log1pf(x);
        return x;
    }
};
template <>
struct Xform<double> {
    // reference arithmetic: (v as f64) * scale, f64::ln_1p  (scale/mod.rs:66-72, transform/mod.rs:38-41)
    template <typename VTI>
    static __device__ __forceinline__ double apply(VTI v, double sc, bool has_scale, bool lg) {
        double x = (double)v;
        if (has_scale) x *= sc;
        if (lg) x = log1p(x);
        return x;
    }
};

struct FusedParams {
    const int64_t *off;
    const uint32_t *idx;
    const void *vin;
    void *vout;
    const double *scale;  // null => none
    int scale_major;
    int do_log1p;
    const int64_t *splits;
    int S;
    uint32_t W;
    uint64_t nmajor, nminor;
    uint32_t rows_per_block;
    const int *fexp;
    unsigned long long *acc;  // 6 * nminor: cnt, sumA, sumB, sqA, sqB, sqC
};

// ---------------------------------------------------------------------------------------------------
// (A) fused exact kernel
// ---------------------------------------------------------------------------------------------------
template <typename VTI, typename VTO, bool WRITE>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_exact_kernel(const FusedParams p) {
    // bins: 4 words per gene [sum_lo | sq_lo | sq_mid | packed(count:14, sum_hi:12, sq_hi:6)], stored in blocks of 32 genes
    // x 4 word-rows of 128 bytes: word w of gene g sits at byte (g >> 5) * 512 + w * 128 + (g & 31) * 4. The four atomics of
    // an entry share one base register (immediate offsets 0 / 128 / 256 / 384) and gene g still maps to bank g mod 32 (a
    // gene-interleaved layout, 16 bytes per gene, put every atomic on 8 of the 32 banks: measured 5.6 instead of 4.3 ms).
    // Addressed in the shared window directly (32-bit addresses, atom.shared): the generic-pointer form recomputed the window
    // base per entry (SASS: S2UR / UMOV / ULEA + 2 IMAD).
    extern __shared__ uint32_t bins[];
    const uint32_t W = p.W;
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(bins);
    asm volatile("" : "+r"(sbase));  // opaque: keeps the compiler from re-deriving the window base (3 uniform ops) per entry
    for (uint32_t i = threadIdx.x; i < 4 * W; i += kFusedThreads) bins[i] = 0;
    __syncthreads();

    const int s = (int)(blockIdx.x % (unsigned)p.S);
    const uint64_t rb = blockIdx.x / (unsigned)p.S;
    const uint32_t col_lo = (uint32_t)s * W;
    const uint64_t r0 = rb * p.rows_per_block;
    const uint64_t r1 = min(r0 + (uint64_t)p.rows_per_block, p.nmajor);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int F = p.fexp[0];
    const double qs = ldexp(1.0, F);
    const float qsf = (float)qs;
    const VTI *vin = reinterpret_cast<const VTI *>(p.vin);
    VTO *vout = reinterpret_cast<VTO *>(p.vout);  // may alias vin (in-place): same thread, same element

    // Memory-level parallelism: the row's segment bounds are fetched one row ahead and the nnz are read in batches of
    // kBatch independent 128-byte warp loads per array before any of them is consumed (the ncu source view of the first
    // version showed ~75 % of the stall samples on the first use of the loaded value / index).
    // The transform flags are uniform for the launch; the row loop is instantiated once per combination (SCALE: 0 none,
    // 1 per row, 2 per column; LG) so that no flag is tested and no scale converted per entry (with the inline-asm atomics
    // the compiler no longer unswitches the loop by itself).
    constexpr int kBatch = 8;
    auto seg_lo = [&](uint64_t r) { return (s == 0) ? p.off[r] : p.splits[(uint64_t)(s - 1) * p.nmajor + r]; };
    auto seg_hi = [&](uint64_t r) { return (s == p.S - 1) ? p.off[r + 1] : p.splits[(uint64_t)s * p.nmajor + r]; };
    auto rows = [&](auto scale_tag, auto lg_tag) {
        constexpr int SCALE = decltype(scale_tag)::value;
        constexpr bool LG = decltype(lg_tag)::value;
        uint64_t r = r0 + warp;
        int64_t a_next = 0, b_next = 0;
        if (r < r1) a_next = seg_lo(r), b_next = seg_hi(r);
        for (; r < r1; r += kFusedThreads / 32) {
            const int64_t a = a_next, b = b_next;
            const uint64_t rn = r + kFusedThreads / 32;
            if (rn < r1) a_next = seg_lo(rn), b_next = seg_hi(rn);
            const double sc_row = SCALE == 1 ? p.scale[r] : 1.0;
            const VTO sc_row_t = (VTO)sc_row;  // f32 pipeline: the scale is rounded once to f32 (Xform<float>)
            // 32-bit offsets relative to the segment start keep the address arithmetic to one IMAD.WIDE per access
            const uint32_t *ip = p.idx + a;
            const VTI *vp = vin + a;
            VTO *op = vout + a;
            const int len = (int)(b - a);
            auto consume = [&](uint32_t c, VTI v, int k) {
                VTO x = (VTO)v;
                if (SCALE == 1) x *= sc_row_t;
                if (SCALE == 2) x *= (VTO)p.scale[c];
                if (LG) x = sizeof(VTO) == 4 ? (VTO)log1pf((float)x) : (VTO)log1p((double)x);
                if (WRITE) op[k] = x;
                uint32_t q;
                if (sizeof(VTO) == 4) q = __float2uint_rn((float)x * qsf);
                else q = (uint32_t)min(__double2ull_rn((double)x * qs), 0xFFFFFFFFULL);
                const uint32_t g = c - col_lo;
                const uint32_t ga = sbase + (g << 2) + (g >> 5) * 384u;
                uint32_t o1, o2, o3;
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o1) : "r"(ga), "r"(q) : "memory");
                const unsigned long long q2 = (unsigned long long)q * q;
                const uint32_t l = (uint32_t)q2, h = (uint32_t)(q2 >> 32);
                asm volatile("atom.shared.add.u32 %0, [%1+128], %2;" : "=r"(o2) : "r"(ga), "r"(l) : "memory");
                uint32_t c1, add3, c3, t0;
                // carries through add.cc / addc instead of compare + select
                asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(t0), "=r"(c1) : "r"(o1), "r"(q));
                asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(t0), "=r"(add3) : "r"(o2), "r"(l), "r"(h));
                asm volatile("atom.shared.add.u32 %0, [%1+256], %2;" : "=r"(o3) : "r"(ga), "r"(add3) : "memory");
                asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(t0), "=r"(c3) : "r"(o3), "r"(add3));
                asm volatile("red.shared.add.u32 [%0+384], %1;" ::"r"(ga), "r"((1u << 18) + c1 * 64u + c3) : "memory");
            };
            int k0 = lane;
            // whole batches: every lane of the warp is in range, no predicate per entry
            for (; k0 - lane + 32 * kBatch <= len; k0 += 32 * kBatch) {
                uint32_t cc[kBatch];
                VTI vv[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) cc[u] = ip[k0 + 32 * u], vv[u] = vp[k0 + 32 * u];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) consume(cc[u], vv[u], k0 + 32 * u);
            }
            if (k0 < len) {  // the ragged tail of the segment
                uint32_t cc[kBatch];
                VTI vv[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int k = k0 + 32 * u;
                    const bool in = k < len;
                    cc[u] = in ? ip[k] : 0u;
                    vv[u] = in ? vp[k] : (VTI)0;
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int k = k0 + 32 * u;
                    if (k < len) consume(cc[u], vv[u], k);
                }
            }
        }
    };
    {
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        const bool lg = p.do_log1p != 0;
        if (p.scale == nullptr) {
            if (lg) rows(I0{}, std::true_type{}); else rows(I0{}, std::false_type{});
        } else if (p.scale_major) {
            if (lg) rows(I1{}, std::true_type{}); else rows(I1{}, std::false_type{});
        } else {
            if (lg) rows(I2{}, std::true_type{}); else rows(I2{}, std::false_type{});
        }
    }
    __syncthreads();
    unsigned long long *acc = p.acc;
    const uint64_t M = p.nminor;
    for (uint32_t g = threadIdx.x; g < W; g += kFusedThreads) {
        const uint32_t *bw = bins + (g >> 5) * 128u + (g & 31u);
        const uint4 bin = make_uint4(bw[0], bw[32], bw[64], bw[96]);  // sum_lo, sq_lo, sq_mid, packed
        const uint32_t pk = bin.w;
        const uint32_t cnt = pk >> 18;
        if (cnt == 0) continue;
        const uint64_t col = (uint64_t)col_lo + g;
        atomicAdd(&acc[col], (unsigned long long)cnt);
        atomicAdd(&acc[M + col], (unsigned long long)bin.x);
        const uint32_t sh = (pk >> 6) & 0xFFFu;
        if (sh) atomicAdd(&acc[2 * M + col], (unsigned long long)sh);
        atomicAdd(&acc[3 * M + col], (unsigned long long)bin.y);
        atomicAdd(&acc[4 * M + col], (unsigned long long)bin.z);
        const uint32_t qh = pk & 63u;
        if (qh) atomicAdd(&acc[5 * M + col], (unsigned long long)qh);
    }

}

// ---------------------------------------------------------------------------------------------------
// (B) transform only (no moments): one warp per line, pure streaming
// ---------------------------------------------------------------------------------------------------
template <typename VTI, typename VTO>
__global__ void __launch_bounds__(256) transform_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                        const VTI *__restrict__ vin, VTO *__restrict__ vout,
                                                        const double *__restrict__ scale, int scale_major,
                                                        int do_log1p, uint64_t nmajor) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    const bool has_scale = scale != nullptr;
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        const int64_t a = off[r], b = off[r + 1];
        const double sc_row = (has_scale && scale_major) ? scale[r] : 1.0;
#pragma unroll 4
        for (int64_t k = a + lane; k < b; k += 32) {
            const double sc = (has_scale && !scale_major) ? scale[idx[k]] : sc_row;
            vout[k] = Xform<VTO>::apply(vin[k], sc, has_scale, do_log1p != 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// (C) general moments (negative / non-finite data, or too many minor lines for stripes): fp64 global REDs
// ---------------------------------------------------------------------------------------------------
template <typename VT>
__global__ void __launch_bounds__(256) moments_general_kernel(const int64_t *__restrict__ off,
                                                              const uint32_t *__restrict__ idx,
                                                              const VT *__restrict__ val, uint64_t nmajor,
                                                              double *__restrict__ cnt, double *__restrict__ sum,
                                                              double *__restrict__ sq) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const uint64_t nwarps = (uint64_t)gridDim.x * 8;
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        const int64_t a = off[r], b = off[r + 1];
        for (int64_t k = a + lane; k < b; k += 32) {
            const uint32_t c = idx[k];
            const double v = (double)val[k];
            atomicAdd(&cnt[c], 1.0);
            atomicAdd(&sum[c], v);
            atomicAdd(&sq[c], v * v);
        }
    }
}

// exact accumulators -> doubles. sum = (sumA + sumB 2^32) 2^-F ; sq = (sqA + sqB 2^32 + sqC 2^64) 2^-2F
__global__ void finalize_exact_kernel(const unsigned long long *__restrict__ acc, uint64_t M, const int *__restrict__ fexp,
                                      double *__restrict__ cnt, double *__restrict__ sum, double *__restrict__ sq) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int F = fexp[0];
    cnt[j] = (double)acc[j];
    const double s = (double)acc[2 * M + j] * 4294967296.0 + (double)acc[M + j];
    // exact 128-bit recombination of the three limbs, then one rounding to double
    unsigned __int128 t = (unsigned __int128)acc[3 * M + j] + ((unsigned __int128)acc[4 * M + j] << 32) +
                          ((unsigned __int128)acc[5 * M + j] << 64);
    const unsigned long long thi = (unsigned long long)(t >> 64), tlo = (unsigned long long)t;
    const double q = (double)thi * 18446744073709551616.0 + (double)tlo;
    sum[j] = ldexp(s, -F);
    sq[j] = ldexp(q, -2 * F);
}

// variance_whole_helper minor branch (csr.rs:179-184): count>0 ? sq/cnt - mean^2 : 0.0
__global__ void minor_variance_kernel(const double *__restrict__ cnt, const double *__restrict__ sum,
                                      const double *__restrict__ sq, uint64_t M, int sqrt_it, double *__restrict__ out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double r = 0.0;
    if (cnt[j] > 0.0) {
        // no FMA contraction: the reference (and the oracle) round mean*mean before the subtraction, which makes
        // the variance of a constant line exactly 0.0 (HVG tie order depends on it)
        const double mean = sum[j] / cnt[j];
        r = __dsub_rn(__ddiv_rn(sq[j], cnt[j]), __dmul_rn(mean, mean));
    }
    out[j] = sqrt_it ? sqrt(r) : r;
}

// Exact-path variance: the integer identity  cnt*sum(q^2) - (sum q)^2 >= 0  is evaluated in 128-bit integers, so
// there is NO cancellation error at all (the reference's f64 one-pass form sq/cnt - mean^2, csr.rs:179-184, loses
// ~eps*E[x^2]/var); a gene whose stored values are all equal gets exactly 0.0 like the reference.
__global__ void minor_variance_exact_kernel(const unsigned long long *__restrict__ acc, uint64_t M, const int *__restrict__ fexp,
                                            int sqrt_it, double *__restrict__ out) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const unsigned long long cnt = acc[j];
    double r = 0.0;
    if (cnt > 0) {
        const unsigned __int128 S = (unsigned __int128)acc[M + j] + ((unsigned __int128)acc[2 * M + j] << 32);
        const unsigned __int128 Q = (unsigned __int128)acc[3 * M + j] + ((unsigned __int128)acc[4 * M + j] << 32) +
                                    ((unsigned __int128)acc[5 * M + j] << 64);
        const unsigned __int128 a = (unsigned __int128)cnt * Q, b = S * S;
        const unsigned __int128 N = a >= b ? a - b : 0;
        const double nd = (double)(unsigned long long)(N >> 64) * 18446744073709551616.0 + (double)(unsigned long long)N;
        r = ldexp(nd / ((double)cnt * (double)cnt), -2 * fexp[0]);
    }
    out[j] = sqrt_it ? sqrt(r) : r;
}

template <typename VT>
__global__ void __launch_bounds__(256) minor_minmax_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx,
                                                           const VT *__restrict__ val, uint64_t nnz,
                                                           unsigned long long *__restrict__ kmin,
                                                           unsigned long long *__restrict__ kmax) {
    for (uint64_t k = (uint64_t)blockIdx.x * 256 + threadIdx.x; k < nnz; k += (uint64_t)gridDim.x * 256) {
        const double v = (double)val[k];
        if (v != v) continue;  // f64::min/max ignore NaN operands
        const unsigned long long key = f64_to_ordered(v);
        const uint32_t c = idx[k];
        if (key < kmin[c]) atomicMin(&kmin[c], key);
        if (key > kmax[c]) atomicMax(&kmax[c], key);
    }
}
__global__ void minmax_init_kernel(unsigned long long *kmin, unsigned long long *kmax, uint64_t M) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    kmin[j] = f64_to_ordered(INFINITY);
    kmax[j] = f64_to_ordered(-INFINITY);
}
__global__ void minmax_decode_kernel(const unsigned long long *kmin, const unsigned long long *kmax, uint64_t M,
                                     double *mn, double *mx) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    mn[j] = ordered_to_f64(kmin[j]);
    mx[j] = ordered_to_f64(kmax[j]);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

static void ensure_splits(srb_mat *m, int S, uint32_t W) {
    Structure &st = *m->st;
    if (S <= 1) return;
    if (st.nstripes == S && st.splits) return;
    cudaStream_t s = m->ctx->stream;
    st.splits = dev_alloc(s, sizeof(int64_t) * st.nmajor * (size_t)(S - 1));
    const uint64_t total = st.nmajor * (uint64_t)(S - 1);
    SRB_LAUNCH(splits_kernel, blocks_for(total, 256), 256, 0, s, st.offsets->as<int64_t>(), st.indices->as<uint32_t>(),
               st.nmajor, S, W, st.splits->as<int64_t>());
    st.nstripes = S;
}

static void stripe_plan(const srb_mat *m, int *S, uint32_t *W) {
    const uint64_t M = m->nminor();
    const uint64_t per = kBinSmemBudget / 16;  // genes per stripe at most
    int s = (int)((M + per - 1) / per);
    if (s < 1) s = 1;
    uint32_t w = (uint32_t)((M + s - 1) / s);
    w = (w + 31u) & ~31u;
    *S = s;
    *W = w;
}

static Buf new_range(cudaStream_t s) {
    Buf r = dev_alloc(s, 2 * sizeof(double));
    SRB_LAUNCH(range_init_kernel, 1, 1, 0, s, r->as<unsigned long long>());
    return r;
}
// {max |v|, min non-zero |v|} of the stored values
static void ensure_range_all(srb_mat *m) {
    if (m->absmax_all) return;
    major_sum_absmax(m);
    cudaStream_t s = m->ctx->stream;
    m->absmax_all = new_range(s);
    if (m->nmajor())
        SRB_LAUNCH(range_prod_kernel, 64, 256, 0, s, m->major.absmax->as<double>(), m->major.absmin->as<double>(),
                   (const double *)nullptr, m->nmajor(), m->absmax_all->as<unsigned long long>());
}

// Decide whether the exact fixed-point path applies (one small D2H + sync):
//  * stored values finite and non-negative (count matrices and their normalised / log1p'd forms always are)
//  * moderate stripe count
//  * quantisation keeps every value to <= 2^-17 relative: either integer-valued data below 2^28 (exact), or a
//    dynamic range max/min <= 4096 of the values whose moments are taken
//  * FAITHFUL f64 storage asks for the reference's f64 accumulation instead
// dec[0] = hi, dec[1] = -lo, dec[2..4] = flags (as doubles) so that one MAX-allreduce makes the decision (and the
// fixed-point scale F derived from dec[0]) identical on every rank of a row-sharded job
__global__ void decision_pack_kernel(const double *__restrict__ range, const uint32_t *__restrict__ flags, double *__restrict__ dec) {
    dec[0] = range[0];
    dec[1] = -range[1];
    dec[2] = (double)flags[0], dec[3] = (double)flags[1], dec[4] = (double)flags[2];
}
static bool exact_path_ok(srb_mat *m, const Buf &range, bool pending, bool lg, int out_dtype, Buf &dec_out, bool reduce) {
    int S;
    uint32_t W;
    stripe_plan(m, &S, &W);
    if (S > kMaxStripes) return false;
    if (m->ctx->value_mode == SRB_VALUES_FAITHFUL && out_dtype == SRB_F64) return false;
    cudaStream_t s = m->ctx->stream;
    Buf dec = dev_alloc(s, 5 * sizeof(double));
    SRB_LAUNCH(decision_pack_kernel, 1, 1, 0, s, range->as<double>(), m->major.flags->as<uint32_t>(), dec->as<double>());
    if (reduce) allreduce_f64_max(m->ctx, dec->as<double>(), 5);
    double h[5];
    SRB_CUDA(cudaMemcpyAsync(h, dec->p, sizeof(h), cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    dec_out = dec;  // dec[0] is the (global) upper bound the fixed-point scale is derived from
    if (h[2] != 0.0 || h[3] != 0.0) return false;
    double hi = h[0], lo = -h[1];
    if (!(hi >= 0.0) || !std::isfinite(hi)) return false;
    if (lg) hi = std::log1p(hi), lo = std::log1p(lo);
    if (!pending && h[4] == 0.0 && hi < 268435456.0) return true;  // integers: exactly representable
    if (hi == 0.0 || std::isinf(lo)) return true;                  // all zeros
    return lo > 0.0 && hi / lo <= 4096.0;
}

template <typename VTI, typename VTO>
static void launch_fused(srb_mat *m, const FusedParams &p, bool write, unsigned grid, size_t smem) {
    cudaStream_t s = m->ctx->stream;
#define SRB_FUSED_GO(W)                                                                                                          \
    do {                                                                                                                         \
        SRB_CUDA(cudaFuncSetAttribute(fused_exact_kernel<VTI, VTO, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        SRB_LAUNCH((fused_exact_kernel<VTI, VTO, W>), grid, kFusedThreads, smem, s, p);                                          \
    } while (0)
    if (write) SRB_FUSED_GO(true);
    else SRB_FUSED_GO(false);
#undef SRB_FUSED_GO
}

// Apply the pending transforms; when want_moments, also produce the per-minor-line moments of the result — summed over
// the ranks of a row-sharded CSR unless local_only (chunk moments of the backed / out-of-core drivers).
void materialize(srb_mat *m, bool want_moments, bool local_only) {
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    const bool pending = m->has_pending();
    const bool reduce = !local_only && c->nranks > 1 && m->format == SRB_CSR;
    if (!pending && (!want_moments || (m->minor.valid && m->minor.reduced == reduce))) return;
    Structure &st = *m->st;
    const uint64_t M = st.nminor, N = st.nmajor, nnz = st.nnz;

    // output storage type: FAITHFUL promotes to f64 whenever a transform is applied (reference behaviour)
    int out_dtype = m->vdtype;
    if (pending && c->value_mode == SRB_VALUES_FAITHFUL && (m->pend_scale || m->src_dtype != SRB_F32)) out_dtype = SRB_F64;

    // the range of the values whose moments are taken
    Buf bound;
    bool exact = false;
    if (want_moments) {
        if (pending) {
            bound = m->pend_bound;  // set_pending_* computed it together with major stats
        } else {
            ensure_range_all(m);
            bound = m->absmax_all;
        }
        Buf dec;
        exact = m->major.valid && exact_path_ok(m, bound, pending, pending && m->pend_log1p, out_dtype, dec, reduce);
        if (exact) bound = dec;
    }

    SRB_TRACE("materialize: decision");
    Buf new_values = m->values;
    const bool need_new_buffer = pending && (out_dtype != m->vdtype || m->values.use_count() > 1);
    if (need_new_buffer) new_values = dev_alloc(s, (out_dtype == SRB_F32 ? 4 : 8) * (nnz ? nnz : 1));

    SRB_TRACE("materialize: value buffer");
    MinorMoments mm;
    Buf new_absmax;
    if (exact) {
        int S;
        uint32_t W;
        stripe_plan(m, &S, &W);
        ensure_splits(m, S, W);
        mm.exact_path = true;
        mm.acc = dev_zeros(s, sizeof(unsigned long long) * 6 * (M ? M : 1));
        mm.fexp = dev_alloc(s, sizeof(int));
        SRB_LAUNCH(fexp_kernel, 1, 1, 0, s, bound->as<double>(), (int)(pending && m->pend_log1p), mm.fexp->as<int>());
        FusedParams p;
        p.off = st.offsets->as<int64_t>();
        p.idx = st.indices->as<uint32_t>();
        p.vin = m->values->p;
        p.vout = new_values->p;
        p.scale = m->pend_scale ? m->pend_scale->as<double>() : nullptr;
        p.scale_major = m->pend_scale_major ? 1 : 0;
        p.do_log1p = m->pend_log1p ? 1 : 0;
        p.splits = S > 1 ? st.splits->as<int64_t>() : nullptr;
        p.S = S;
        p.W = W;
        p.nmajor = N;
        p.nminor = M;
        uint64_t rpb = (N + (uint64_t)c->sm_count * 2 - 1) / ((uint64_t)c->sm_count * 2);
        if (rpb < 32) rpb = 32;
        if (rpb > kMaxRowsPerBlock) rpb = kMaxRowsPerBlock;
        p.rows_per_block = (uint32_t)rpb;
        p.fexp = mm.fexp->as<int>();
        p.acc = mm.acc->as<unsigned long long>();
        const uint64_t nb = (N + rpb - 1) / rpb;
        const unsigned grid = (unsigned)(nb * (uint64_t)S);
        const size_t smem = (size_t)W * 16;
        if (grid > 0) {
            StageTimer t(c, ST_FUSED);
            const bool in32 = m->vdtype == SRB_F32, out32 = out_dtype == SRB_F32;
            if (in32 && out32) launch_fused<float, float>(m, p, pending, grid, smem);
            else if (in32 && !out32) launch_fused<float, double>(m, p, pending, grid, smem);
            else if (!in32 && !out32) launch_fused<double, double>(m, p, pending, grid, smem);
            else throw Error(SRB_ERR_INVALID_ARG, "f64 -> f32 demotion is never requested");
        }
        mm.cnt = dev_alloc(s, sizeof(double) * (M ? M : 1));
        mm.sum = dev_alloc(s, sizeof(double) * (M ? M : 1));
        mm.sq = dev_alloc(s, sizeof(double) * (M ? M : 1));
        if (reduce) {
            StageTimer t(c, ST_ALLREDUCE);
            allreduce_u64_sum(c, mm.acc->as<uint64_t>(), 6 * M);
            mm.reduced = true;
        }
        if (M) SRB_LAUNCH(finalize_exact_kernel, blocks_for(M, 256), 256, 0, s, mm.acc->as<unsigned long long>(), M,
                          mm.fexp->as<int>(), mm.cnt->as<double>(), mm.sum->as<double>(), mm.sq->as<double>());
        mm.valid = true;
    } else {
        if (pending && nnz) {
            StageTimer t(c, ST_FUSED);
            const unsigned grid = (unsigned)std::min<uint64_t>((N + 7) / 8, (uint64_t)c->sm_count * 32);
            const double *sc = m->pend_scale ? m->pend_scale->as<double>() : nullptr;
            const int sm = m->pend_scale_major ? 1 : 0, lg = m->pend_log1p ? 1 : 0;
            const int64_t *off = st.offsets->as<int64_t>();
            const uint32_t *idx = st.indices->as<uint32_t>();
            if (m->vdtype == SRB_F32 && out_dtype == SRB_F32)
                SRB_LAUNCH((transform_kernel<float, float>), grid, 256, 0, s, off, idx, m->values->as<float>(), new_values->as<float>(), sc, sm, lg, N);
            else if (m->vdtype == SRB_F32)
                SRB_LAUNCH((transform_kernel<float, double>), grid, 256, 0, s, off, idx, m->values->as<float>(), new_values->as<double>(), sc, sm, lg, N);
            else
                SRB_LAUNCH((transform_kernel<double, double>), grid, 256, 0, s, off, idx, m->values->as<double>(), new_values->as<double>(), sc, sm, lg, N);
        }
        if (want_moments) {
            mm.exact_path = false;
            mm.cnt = dev_zeros(s, sizeof(double) * (M ? M : 1));
            mm.sum = dev_zeros(s, sizeof(double) * (M ? M : 1));
            mm.sq = dev_zeros(s, sizeof(double) * (M ? M : 1));
            if (nnz) {
                const unsigned grid = (unsigned)std::min<uint64_t>((N + 7) / 8, (uint64_t)c->sm_count * 32);
                const int64_t *off = st.offsets->as<int64_t>();
                const uint32_t *idx = st.indices->as<uint32_t>();
                if (out_dtype == SRB_F32)
                    SRB_LAUNCH((moments_general_kernel<float>), grid, 256, 0, s, off, idx, new_values->as<float>(), N, mm.cnt->as<double>(), mm.sum->as<double>(), mm.sq->as<double>());
                else
                    SRB_LAUNCH((moments_general_kernel<double>), grid, 256, 0, s, off, idx, new_values->as<double>(), N, mm.cnt->as<double>(), mm.sum->as<double>(), mm.sq->as<double>());
            }
            if (reduce) {
                StageTimer t(c, ST_ALLREDUCE);
                allreduce_f64_sum(c, mm.cnt->as<double>(), M);
                allreduce_f64_sum(c, mm.sum->as<double>(), M);
                allreduce_f64_sum(c, mm.sq->as<double>(), M);
                mm.reduced = true;
            }
            mm.valid = true;
        }
    }

    if (pending) {
        m->values = new_values;
        m->vdtype = out_dtype;
        m->pend_scale.reset();
        m->pend_log1p = false;
        m->pend_bound.reset();
        m->major = MajorStats();
        m->absmax_all = new_absmax;  // unknown after a transform (null): recomputed by a K1 pass if ever needed
        m->minor = MinorMoments();
    }
    if (want_moments) m->minor = mm;
}

void ensure_minor_moments(srb_mat *m, bool local_only) {
    const bool reduce = !local_only && m->ctx->nranks > 1 && m->format == SRB_CSR;
    if (m->minor.valid && !m->has_pending() && m->minor.reduced == reduce) return;
    materialize(m, true, local_only);  // a cache of the other flavour (local vs reduced) is recomputed, never reused
}

void minor_variance_from_moments(srb_mat *m, double *d_out, bool sqrt_it) {
    ensure_minor_moments(m);
    const uint64_t M = m->nminor();
    if (!M) return;
    if (m->minor.exact_path)
        SRB_LAUNCH(minor_variance_exact_kernel, blocks_for(M, 256), 256, 0, m->ctx->stream, m->minor.acc->as<unsigned long long>(), M,
                   m->minor.fexp->as<int>(), sqrt_it ? 1 : 0, d_out);
    else
        SRB_LAUNCH(minor_variance_kernel, blocks_for(M, 256), 256, 0, m->ctx->stream, m->minor.cnt->as<double>(),
                   m->minor.sum->as<double>(), m->minor.sq->as<double>(), M, sqrt_it ? 1 : 0, d_out);
}

void minor_min_max(srb_mat *m, double *d_min, double *d_max) {
    if (m->has_pending()) materialize(m, false);
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    const uint64_t M = m->nminor(), nnz = m->st->nnz;
    if (!M) return;
    Buf keys = dev_alloc(s, sizeof(unsigned long long) * 2 * M);
    unsigned long long *kmin = keys->as<unsigned long long>(), *kmax = kmin + M;
    SRB_LAUNCH(minmax_init_kernel, blocks_for(M, 256), 256, 0, s, kmin, kmax, M);
    if (nnz) {
        const unsigned grid = (unsigned)std::min<uint64_t>((nnz + 255) / 256, (uint64_t)c->sm_count * 16);
        if (m->vdtype == SRB_F32)
            SRB_LAUNCH((minor_minmax_kernel<float>), grid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<float>(), nnz, kmin, kmax);
        else
            SRB_LAUNCH((minor_minmax_kernel<double>), grid, 256, 0, s, m->st->offsets->as<int64_t>(), m->st->indices->as<uint32_t>(), m->values->as<double>(), nnz, kmin, kmax);
    }
    SRB_LAUNCH(minmax_decode_kernel, blocks_for(M, 256), 256, 0, s, kmin, kmax, M, d_min, d_max);
    if (c->nranks > 1 && m->format == SRB_CSR) {
        allreduce_f64_min(c, d_min, M);
        allreduce_f64_max(c, d_max, M);
    }
}

// ---- normalisation bookkeeping (called from api.cu) --------------------------------------------------
void set_pending_normalize(srb_mat *m, double target, int direction) {
    srb_ctx *c = m->ctx;
    cudaStream_t s = c->stream;
    if (m->has_pending()) materialize(m, false);  // sums must be those of the current logical values
    const bool major = m->dir_is_major(direction);
    const uint64_t n = major ? m->nmajor() : m->nminor();
    Buf scale = dev_alloc(s, sizeof(double) * (n ? n : 1));
    Buf bound = new_range(s);
    {
        StageTimer t(c, ST_ROWSUM);
        major_sum_absmax(m);
    }
    if (major) {
        if (n) SRB_LAUNCH(line_scale_kernel, blocks_for(n, 256), 256, 0, s, m->major.sum->as<double>(), n, target, scale->as<double>());
        if (n) SRB_LAUNCH(range_prod_kernel, 64, 256, 0, s, m->major.absmax->as<double>(), m->major.absmin->as<double>(), scale->as<double>(), n, bound->as<unsigned long long>());
    } else {
        ensure_minor_moments(m);
        // on a sharded CSR the minor sums are already global (allreduced); scale is replicated
        if (n) SRB_LAUNCH(line_scale_kernel, blocks_for(n, 256), 256, 0, s, m->minor.sum->as<double>(), n, target, scale->as<double>());
        Buf srange = new_range(s);
        if (n) SRB_LAUNCH(range_prod_kernel, 64, 256, 0, s, scale->as<double>(), scale->as<double>(), (const double *)nullptr, n, srange->as<unsigned long long>());
        ensure_range_all(m);
        SRB_LAUNCH(range_mul_kernel, 1, 1, 0, s, bound->as<double>(), m->absmax_all->as<double>(), srange->as<double>());
    }
    if (!(target >= 0.0)) {
        // a negative target flips signs: the exact (non-negative) path is not applicable
        const double inf[2] = {INFINITY, 0.0};
        SRB_CUDA(cudaMemcpyAsync(bound->p, inf, sizeof(inf), cudaMemcpyHostToDevice, s));
        SRB_CUDA(cudaStreamSynchronize(s));
    }
    m->pend_scale = scale;
    m->pend_scale_major = major;
    m->pend_bound = bound;
    // caches describe the pre-transform values from now on: keep major (needed for flags) but drop minor
    m->minor = MinorMoments();
}

void set_pending_log1p(srb_mat *m) {
    if (m->pend_log1p) materialize(m, false);  // log1p(log1p(x)): apply the first one now
    if (!m->pend_scale) {
        ensure_range_all(m);  // range of the stored values (also computes the major flags)
        m->pend_bound = m->absmax_all;
    }
    m->pend_log1p = true;
    m->minor = MinorMoments();
}

}  // namespace srb
