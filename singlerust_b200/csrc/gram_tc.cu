// gram_tc.cu — K7 on the 5th-generation tensor cores: C = Z^T Z for the standardised HVG block, hand-written
// tcgen05 / TMEM / TMA (sm_100a only). Replaces the SVD of the n x d block that the reference delegates to
// single_algebra (dim_red/mod.rs:66) by the d x d Gram product it is equivalent to.
//
// Operands: Z is held as two fp16 panels Zh, Zl (z = zh + zl to 22 bits), row-major [n cells][dpad genes]. For
// C[i][j] = sum_cells Z[c][i] Z[c][j] both MMA operands are "MN-major" (the gene index is contiguous in memory, the
// contraction index = cell is strided): TMA boxes of 64 genes (128 B) x 64 cells land in shared memory in exactly the
// canonical MN-major SWIZZLE_128B layout ((64 mn) x (8 k) atoms of 1024 B, SBO = 1024 B between k-groups, LBO = one
// box = 8192 B between 64-gene groups).
// Precision: 3 MMAs per k-step (hi*hi + hi*lo + lo*hi; the dropped lo*lo term is 2^-22 relative). Every tcgen05.mma
// TRUNCATES the fp32 TMEM accumulator once, which shrinks same-sign sums by ~0.3 ulp per instruction (measured:
// 1e-5 relative after ~300 MMAs). So accumulation is three-level:
//   TMEM   fp32, 128 cells (24 MMAs) per chunk, two accumulator stages of 256 columns (ping-pong with the epilogue)
//   regs   fp32 round-to-nearest sums of 64 chunks, held by the 8 epilogue warps (one TMEM lane x 128 columns each)
//   global fp64 partial tile per work item, updated every 8192 cells
// Work decomposition: upper-triangular 128 x 256 tiles x a split of the cell range, one work item per CTA
// (72 tiles x 2 = 144 CTAs on 148 SMs at d = 2048); a small kernel sums the splits and mirrors the triangle.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2-9 = epilogue (TMEM lane quarter = warp_idx % 4, column half = (warp_idx - 2) / 4).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace srb {
namespace tc {

constexpr uint32_t BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr uint32_t BOX_BYTES = 64 * 2 * BK;            // 64 genes x 2 B x 64 cells = 8192
constexpr uint32_t A_BOXES = BM / 64, B_BOXES = BN / 64;
constexpr uint32_t STAGE_BYTES = (A_BOXES + B_BOXES) * 2 * BOX_BYTES;  // hi + lo = 98304
constexpr uint32_t STAGES = 2;
constexpr uint32_t CHUNK_KBLOCKS = 2;                  // 128 cells (24 MMAs) per TMEM accumulation chunk
constexpr uint32_t FLUSH_CHUNKS = 64;                  // register-level fp32 sums are folded into fp64 every 64 chunks
constexpr uint32_t THREADS = 192;                      // scores kernel
constexpr uint32_t GRAM_THREADS = 320;                 // 2 control warps + 8 epilogue warps
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, MN-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);              // start address       [0,14)
    d |= (uint64_t)((BOX_BYTES >> 4) & 0x3FFF) << 16;    // leading byte offset [16,30): next 64-wide MN group
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;        // stride byte offset  [32,46): next group of 8 k-rows
    d |= 1ull << 46;                                     // descriptor version = 1 (Blackwell)
    d |= 2ull << 61;                                     // layout type = SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = f16, both MN-major, M = 128, N = 256
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((BN >> 3) << 17) | ((BM >> 4) << 24);

struct GramParams {
    const uint2 *tiles;   // (ti, tj) per tile
    uint32_t ksplit;
    uint32_t kblocks_total;
    double *partial;      // [items][BM][BN]
};

__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_tc_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                             const __grid_constant__ CUtensorMap map_lo, const GramParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * STAGES;
    const uint32_t tfull_bar = bar_base + 16 * STAGES, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t item = blockIdx.x;
    const uint32_t tile = item / p.ksplit, ks = item % p.ksplit;
    const uint2 tij = p.tiles[tile];
    const uint32_t per = (p.kblocks_total + p.ksplit - 1) / p.ksplit;
    const uint32_t kb0 = min(ks * per, p.kblocks_total), kb1 = min(kb0 + per, p.kblocks_total);
    const uint32_t nkb = kb1 - kb0;
    const uint32_t nchunks = (nkb + CHUNK_KBLOCKS - 1) / CHUNK_KBLOCKS;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < STAGES; ++s) mbar_init(full_bar + 8 * s, 1), mbar_init(empty_bar + 8 * s, 1);
        for (uint32_t a = 0; a < 2; ++a) mbar_init(tfull_bar + 8 * a, 1), mbar_init(tempty_bar + 8 * a, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (uint32_t it = 0; it < nkb; ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty_bar + 8 * s, ph ^ 1);
                mbar_arrive_expect_tx(full_bar + 8 * s, STAGE_BYTES);
                const int32_t row = (int32_t)((kb0 + it) * BK);
                const uint32_t sa = base + s * STAGE_BYTES;
                const uint32_t a_hi = sa, a_lo = sa + A_BOXES * BOX_BYTES;
                const uint32_t b_hi = sa + 2 * A_BOXES * BOX_BYTES, b_lo = b_hi + B_BOXES * BOX_BYTES;
#pragma unroll
                for (uint32_t b = 0; b < A_BOXES; ++b) {
                    const int32_t g = (int32_t)(tij.x * BM + b * 64);
                    tma_load_2d(a_hi + b * BOX_BYTES, &map_hi, full_bar + 8 * s, g, row);
                    tma_load_2d(a_lo + b * BOX_BYTES, &map_lo, full_bar + 8 * s, g, row);
                }
#pragma unroll
                for (uint32_t b = 0; b < B_BOXES; ++b) {
                    const int32_t g = (int32_t)(tij.y * BN + b * 64);
                    tma_load_2d(b_hi + b * BOX_BYTES, &map_hi, full_bar + 8 * s, g, row);
                    tma_load_2d(b_lo + b * BOX_BYTES, &map_lo, full_bar + 8 * s, g, row);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        uint32_t it = 0;
        for (uint32_t c = 0; c < nchunks; ++c) {
            const uint32_t as = c & 1;
            mbar_wait(tempty_bar + 8 * as, ((c >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            const uint32_t kend = min((c + 1) * CHUNK_KBLOCKS, nkb);
            for (uint32_t kb = c * CHUNK_KBLOCKS; kb < kend; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full_bar + 8 * s, ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = base + s * STAGE_BYTES;
                    const uint32_t a_hi = sa, a_lo = sa + A_BOXES * BOX_BYTES;
                    const uint32_t b_hi = sa + 2 * A_BOXES * BOX_BYTES, b_lo = b_hi + B_BOXES * BOX_BYTES;
#pragma unroll
                    for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
                        const uint32_t ko = kk * UMMA_K * 128;  // 16 k-rows x 128 B
                        const uint64_t dah = make_desc_mn_sw128(a_hi + ko), dal = make_desc_mn_sw128(a_lo + ko);
                        const uint64_t dbh = make_desc_mn_sw128(b_hi + ko), dbl = make_desc_mn_sw128(b_lo + ko);
                        const uint32_t first = (kb == c * CHUNK_KBLOCKS && kk == 0) ? 0u : 1u;
                        tc_mma_f16(d_tmem, dal, dbh, IDESC, first);  // small cross terms first
                        tc_mma_f16(d_tmem, dah, dbl, IDESC, 1u);
                        tc_mma_f16(d_tmem, dah, dbh, IDESC, 1u);
                    }
                    tc_commit(empty_bar + 8 * s);                       // smem stage free once these MMAs retire
                    if (kb + 1 == kend) tc_commit(tfull_bar + 8 * as);  // accumulator chunk complete
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> fp32 registers (RN) -> fp64 partial tile ----------------
        const uint32_t q = warp & 3;             // TMEM lane quarter this warp may access
        const uint32_t half = (warp - 2) >> 2;   // which 128 of the 256 accumulator columns
        const uint32_t row = q * 32 + lane;
        double *prow = p.partial + (size_t)item * (BM * BN) + (size_t)row * BN + half * 128;
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.f;
        for (uint32_t c = 0; c < nchunks; ++c) {
            const uint32_t as = c & 1;
            mbar_wait(tfull_bar + 8 * as, (c >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (uint32_t g = 0; g < 4; ++g) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * BN + half * 128 + g * 32;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[g * 32 + i] += __uint_as_float(r[i]);
            }
            tc_fence_before();
            mbar_arrive(tempty_bar + 8 * as);
            if ((c + 1) % FLUSH_CHUNKS == 0 || c + 1 == nchunks) {
                double2 *pp = reinterpret_cast<double2 *>(prow);
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    double2 v = pp[i];
                    v.x += (double)acc[2 * i];
                    v.y += (double)acc[2 * i + 1];
                    pp[i] = v;
                    acc[2 * i] = 0.f, acc[2 * i + 1] = 0.f;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// G[i][j] = G[j][i] = sum over k-splits of the partial tile entry, for the computed entries with i <= j
__global__ void gram_reduce_kernel(const double *__restrict__ partial, const uint2 *__restrict__ tiles, uint32_t ntiles,
                                   uint32_t ksplit, uint32_t dpad, double *__restrict__ G) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_tile = (uint64_t)BM * BN;
    if (e >= per_tile * ntiles) return;
    const uint32_t t = (uint32_t)(e / per_tile);
    const uint32_t r = (uint32_t)((e % per_tile) / BN), c = (uint32_t)(e % BN);
    const uint32_t i = tiles[t].x * BM + r, j = tiles[t].y * BN + c;
    if (i > j) return;
    double s = 0.0;
    for (uint32_t ks = 0; ks < ksplit; ++ks) s += partial[((size_t)t * ksplit + ks) * per_tile + (size_t)r * BN + c];
    G[(uint64_t)i * dpad + j] = s;
    G[(uint64_t)j * dpad + i] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SRB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
        SRB_REQUIRE(f && qres == cudaDriverEntryPointSuccess, SRB_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
        fn = (EncodeTiledFn)f;
    }
    return fn;
}

static CUtensorMap panel_map(const __half *X, uint64_t n, uint32_t dpad) {
    CUtensorMap m;
    cuuint64_t dims[2] = {dpad, n};
    cuuint64_t strides[1] = {(cuuint64_t)dpad * 2};
    cuuint32_t box[2] = {64, BK};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, SRB_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return m;
}

}  // namespace tc

CUtensorMap gram_panel_map(const __half *X, uint64_t n, uint32_t dpad) { return tc::panel_map(X, n, dpad); }
void gram_tcgen05_pair(srb_ctx *ctx, const __half *Xh, const __half *Xl, uint64_t n, uint32_t dpad, double *G, uint32_t d_used);  // gram_tc2.cu

void gram_tcgen05(srb_ctx *ctx, const __half *Xh, const __half *Xl, uint64_t n, uint32_t dpad, double *G, uint32_t d_used) {
    using namespace tc;
    if (n == 0) return;
    {
        // CTA-pair kernel (cta_group::2) unless SRB_GRAM_PAIR=0
        static int use_pair = -1;
        if (use_pair < 0) {
            const char *e = getenv("SRB_GRAM_PAIR");
            use_pair = (e && e[0] == '0') ? 0 : 1;
        }
        if (use_pair) {
            gram_tcgen05_pair(ctx, Xh, Xl, n, dpad, G, d_used);
            return;
        }
    }
    cudaStream_t s = ctx->stream;
    SRB_REQUIRE(dpad % BN == 0, SRB_ERR_INVALID_ARG, "dpad must be a multiple of 256");
    const uint32_t NI = dpad / BM, NJ = dpad / BN;
    std::vector<uint2> tiles;
    for (uint32_t ti = 0; ti < NI; ++ti)
        for (uint32_t tj = ti / 2; tj < NJ; ++tj) tiles.push_back(make_uint2(ti, tj));
    const uint32_t ntiles = (uint32_t)tiles.size();
    const uint32_t kblocks = (uint32_t)((n + BK - 1) / BK);
    uint32_t ksplit = std::max<uint32_t>(1, (uint32_t)ctx->sm_count / ntiles);
    ksplit = std::min(ksplit, kblocks);
    const uint32_t items = ntiles * ksplit;
    Buf d_tiles = dev_alloc(s, sizeof(uint2) * ntiles);
    SRB_CUDA(cudaMemcpyAsync(d_tiles->p, tiles.data(), sizeof(uint2) * ntiles, cudaMemcpyHostToDevice, s));
    Buf partial = dev_zeros(s, sizeof(double) * (size_t)items * BM * BN);
    CUtensorMap mh = panel_map(Xh, n, dpad), ml = panel_map(Xl, n, dpad);
    GramParams p;
    p.tiles = d_tiles->as<uint2>();
    p.ksplit = ksplit;
    p.kblocks_total = kblocks;
    p.partial = partial->as<double>();
    SRB_CUDA(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SRB_LAUNCH(gram_tc_kernel, items, GRAM_THREADS, SMEM_BYTES, s, mh, ml, p);
    const uint64_t total = (uint64_t)ntiles * BM * BN;
    SRB_LAUNCH(gram_reduce_kernel, (unsigned)((total + 255) / 256), 256, 0, s, partial->as<double>(), d_tiles->as<uint2>(), ntiles, ksplit, dpad, G);
    SRB_CUDA(cudaStreamSynchronize(s));  // `tiles` (host vector) must outlive the async copy
}

// =====================================================================================================================
// K9 on tensor cores: scores = Z V_k  (pca/mod.rs:156-185 `transform`). M = 128 cells per tile, N = kpad = 64 components,
// K = dpad genes. Both operands K-major (genes contiguous): A = Z panel rows, B = Wt[comp][gene] (split-fp16 of the
// fp64 eigenvectors). 3 MMAs per k-step (zl*wh + zh*wl + zh*wh). Every tcgen05.mma truncates the fp32 TMEM accumulator
// (gram kernel header), so the 2048-gene contraction (384 MMAs) is split into 4 chunks with their own TMEM columns
// (96 MMAs each) that the epilogue adds in fp64: round 2 measured max score errors of 5e-5 rms with a single chain. The kernel is HBM-bound (reads the 8 GB of panels once); persistent CTAs, 4-stage TMA
// ring, two TMEM accumulator stages so the epilogue of tile t overlaps the loads/MMAs of tile t+1.
// =====================================================================================================================
namespace sc {
using namespace tc;
constexpr uint32_t SM = 128, SN = 64, SBK = 64;
constexpr uint32_t A_BYTES = SM * SBK * 2;          // 16384 per panel half
constexpr uint32_t B_BYTES = SN * SBK * 2;          // 8192
constexpr uint32_t S_STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 49152
constexpr uint32_t S_STAGES = 4;
constexpr uint32_t S_SMEM_BYTES = S_STAGES * S_STAGE_BYTES + 1024 + 256;
constexpr uint32_t S_KCHUNKS = 4;      // the gene contraction is accumulated in 4 separate TMEM chunks, summed in fp64
constexpr uint32_t S_TMEM_COLS = 512;  // 2 accumulator stages x 4 chunks x 64 columns
constexpr uint32_t S_IDESC = (1u << 4) | ((SN >> 3) << 17) | ((SM >> 4) << 24);  // f32 accum, f16 x f16, K-major both

// K-major SWIZZLE_128B descriptor: rows of 128 B (64 halves), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;  // SBO; LBO unused for a single swizzle atom along K
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

__global__ void __launch_bounds__(THREADS, 1) scores_tc_kernel(const __grid_constant__ CUtensorMap map_zh,
                                                               const __grid_constant__ CUtensorMap map_zl,
                                                               const __grid_constant__ CUtensorMap map_wh,
                                                               const __grid_constant__ CUtensorMap map_wl, uint64_t nrows,
                                                               uint32_t kblocks, uint32_t k, const double *__restrict__ bias,
                                                               double *__restrict__ scores) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + S_STAGES * S_STAGE_BYTES;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * S_STAGES;
    const uint32_t tfull_bar = bar_base + 16 * S_STAGES, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ntiles = (uint32_t)((nrows + SM - 1) / SM);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < S_STAGES; ++s) mbar_init(full_bar + 8 * s, 1), mbar_init(empty_bar + 8 * s, 1);
        for (uint32_t a = 0; a < 2; ++a) mbar_init(tfull_bar + 8 * a, 1), mbar_init(tempty_bar + 8 * a, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(S_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int32_t row = (int32_t)(t * SM);
                for (uint32_t kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % S_STAGES, ph = (it / S_STAGES) & 1;
                    mbar_wait(empty_bar + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(full_bar + 8 * s, S_STAGE_BYTES);
                    const uint32_t sa = base + s * S_STAGE_BYTES;
                    const int32_t g = (int32_t)(kb * SBK);
                    tma_load_2d(sa, &map_zh, full_bar + 8 * s, g, row);
                    tma_load_2d(sa + A_BYTES, &map_zl, full_bar + 8 * s, g, row);
                    tma_load_2d(sa + 2 * A_BYTES, &map_wh, full_bar + 8 * s, g, 0);
                    tma_load_2d(sa + 2 * A_BYTES + B_BYTES, &map_wl, full_bar + 8 * s, g, 0);
                }
            }
        }
    } else if (warp == 1) {
        uint32_t it = 0, lt = 0;
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
            const uint32_t as = lt & 1;
            mbar_wait(tempty_bar + 8 * as, ((lt >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t per = (kblocks + S_KCHUNKS - 1) / S_KCHUNKS;  // k-blocks per accumulation chunk
            for (uint32_t kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t s = it % S_STAGES, ph = (it / S_STAGES) & 1;
                const uint32_t d_tmem = tmem_base + as * (S_KCHUNKS * SN) + (kb / per) * SN;
                mbar_wait(full_bar + 8 * s, ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = base + s * S_STAGE_BYTES;
#pragma unroll
                    for (uint32_t kk = 0; kk < SBK / UMMA_K; ++kk) {
                        const uint32_t ko = kk * UMMA_K * 2;  // 16 halves = 32 B along the swizzled row
                        const uint64_t dzh = make_desc_k_sw128(sa + ko), dzl = make_desc_k_sw128(sa + A_BYTES + ko);
                        const uint64_t dwh = make_desc_k_sw128(sa + 2 * A_BYTES + ko), dwl = make_desc_k_sw128(sa + 2 * A_BYTES + B_BYTES + ko);
                        const uint32_t first = (kb % per == 0 && kk == 0) ? 0u : 1u;
                        tc_mma_f16(d_tmem, dzl, dwh, S_IDESC, first);
                        tc_mma_f16(d_tmem, dzh, dwl, S_IDESC, 1u);
                        tc_mma_f16(d_tmem, dzh, dwh, S_IDESC, 1u);
                    }
                    tc_commit(empty_bar + 8 * s);
                    if (kb + 1 == kblocks) tc_commit(tfull_bar + 8 * as);
                }
                __syncwarp();
            }
        }
    } else {
        const uint32_t q = warp & 3;
        uint32_t lt = 0;
        for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
            const uint32_t as = lt & 1;
            mbar_wait(tfull_bar + 8 * as, (lt >> 1) & 1);
            tc_fence_after();
            const uint64_t row = (uint64_t)t * SM + q * 32 + lane;
            const uint32_t nchunk = (kblocks + ((kblocks + S_KCHUNKS - 1) / S_KCHUNKS) - 1) / ((kblocks + S_KCHUNKS - 1) / S_KCHUNKS);
#pragma unroll
            for (uint32_t col0 = 0; col0 < SN; col0 += 32) {
                double acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = 0.0;
                for (uint32_t ch = 0; ch < nchunk; ++ch) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * (S_KCHUNKS * SN) + ch * SN + col0;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr)
                        : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] += (double)__uint_as_float(r[i]);
                }
                if (row < nrows) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < k) scores[row * k + col0 + i] = acc[i] - bias[col0 + i];
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar + 8 * as);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(S_TMEM_COLS) : "memory");
    }
}

// Wt_{h,l}[c][p] = split-fp16 of W[p][c]   (W: dpad x kpad fp64 row-major, zero padded)
__global__ void w_split_kernel(const double *__restrict__ W, uint32_t dpad, uint32_t kpad, __half *__restrict__ wh, __half *__restrict__ wl) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= dpad * kpad) return;
    const uint32_t c = e / dpad, pidx = e % dpad;
    const double w = W[(size_t)pidx * kpad + c];
    const __half h = __double2half(w);
    wh[e] = h;
    wl[e] = __double2half(w - (double)__half2float(h));
}

static CUtensorMap map2d(const __half *X, uint64_t rows, uint32_t cols, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)X, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SRB_REQUIRE(r == CUDA_SUCCESS, SRB_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return m;
}
}  // namespace sc

void scores_tcgen05(srb_ctx *ctx, const __half *Xh, const __half *Xl, uint64_t n, uint32_t dpad, const double *W, uint32_t kpad,
                    uint32_t k, const double *bias, double *scores) {
    using namespace sc;
    if (n == 0) return;
    SRB_REQUIRE(kpad == SN && k <= SN, SRB_ERR_UNSUPPORTED, "tensor-core scores support up to 64 components");
    SRB_REQUIRE(dpad % SBK == 0, SRB_ERR_INVALID_ARG, "dpad must be a multiple of 64");
    cudaStream_t s = ctx->stream;
    Buf wh = dev_alloc(s, 2 * (size_t)dpad * kpad), wl = dev_alloc(s, 2 * (size_t)dpad * kpad);
    SRB_LAUNCH(w_split_kernel, (dpad * kpad + 255) / 256, 256, 0, s, W, dpad, kpad, wh->as<__half>(), wl->as<__half>());
    CUtensorMap mzh = map2d(Xh, n, dpad, SM), mzl = map2d(Xl, n, dpad, SM);
    CUtensorMap mwh = map2d(wh->as<__half>(), kpad, dpad, SN), mwl = map2d(wl->as<__half>(), kpad, dpad, SN);
    const uint32_t ntiles = (uint32_t)((n + SM - 1) / SM);
    const unsigned grid = std::min<uint32_t>(ntiles, (uint32_t)ctx->sm_count);
    SRB_CUDA(cudaFuncSetAttribute(scores_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S_SMEM_BYTES));
    SRB_LAUNCH(scores_tc_kernel, grid, THREADS, S_SMEM_BYTES, s, mzh, mzl, mwh, mwl, n, dpad / SBK, k, bias, scores);
}

}  // namespace srb
