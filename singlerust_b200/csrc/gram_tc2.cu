// gram_tc2.cu — K7 with CTA pairs (tcgen05 cta_group::2): the same Gram product as gram_tc.cu on 256 x 256 pair tiles.
// Each CTA of a 2-CTA cluster loads its own 128 rows of the A operand and its own 128-column half of the B operand
// (64 KB per 64-cell k-block instead of 96 KB for a 128 x 256 single-CTA tile), the leader CTA issues one
// tcgen05.mma.cta_group::2 (M = 256, N = 256) per split term, and each CTA's epilogue warps fold their own 128 TMEM
// lanes into the fp64 partial tile. The single-CTA kernel was L2-fill bound (~40 B/clk/SM); this halves the operand
// bytes per MMA flop (3 stages x 64 KB).
// Schedule: 36 pair tiles of the upper triangle (d = 2000 -> 8 x 8 tile grid) x 2 halves of the cells = 72 work items on the
// 74 CTA pairs of the chip, one item each, all sweeping the cells in lock step: the 64-cell slab of the panels that one item
// pulls into L2 is what the other 35 tiles of the same half need at that moment, which is why DRAM sees 12 GB for 8.4 GB of
// panels and not 36 x that. ncu: tensor pipe 81 % active; the kernel lasts as long as ONE full item.
// SRB_GRAM_TRIM=1 runs the tiles of the last tile column with N trimmed to the selected genes (208 of 256 at d = 2000).
// Measured in round 2 and left OFF: 3 % fewer executed flops, same duration (8.12 vs 8.10 ms — the other 64 items set the
// makespan), and the trimmed items run ahead of the sweep, which breaks the lock step: DRAM reads 12.0 -> 21.4 GB, L2 hit
// rate 76 -> 65 % (profiles/r02_summary.md). The same effect rules out a stream-K style balanced split of (tile, k-block)
// work over all 74 pairs, which would otherwise be worth up to 7 %.
//
// Barrier protocol (all barriers exist in both CTAs at the same shared-memory offsets):
//   full[s]   lives in the leader; count 2 (one arrive per producer) + 2 x 64 KB of TMA transaction bytes
//             (cp.async.bulk.tensor...cta_group::2 signals the leader's barrier from both CTAs)
//   empty[s]  per CTA; released by tcgen05.commit.cta_group::2 multicast to both CTAs
//   tfull[a]  per CTA; same multicast commit at the end of an accumulation chunk
//   tempty[a] lives in the leader; count 512 = the 2 x 256 epilogue threads of both CTAs (remote arrive from the peer)
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace srb {
namespace tc2 {

constexpr uint32_t TM = 256, TN = 256, BK = 64, UMMA_K = 16;
constexpr uint32_t BOX_BYTES = 64 * 2 * BK;                // 8192
constexpr uint32_t HALF_BOXES = 2;                         // 128 genes per CTA per operand
constexpr uint32_t STAGE_BYTES = 2 * 2 * HALF_BOXES * BOX_BYTES;  // (A + B) x (hi + lo) x 2 boxes = 65536
constexpr uint32_t STAGES = 3;
constexpr uint32_t FLUSH_CELLS = 8192;  // fp32 register sums are folded into fp64 every 8192 cells
constexpr uint32_t THREADS = 320;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
// instruction descriptor: D = f32, A = B = f16, both MN-major, M = 256 (the pair), N = n (256, or less in the last tile column)
__host__ __device__ constexpr uint32_t idesc_n(uint32_t n) { return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((TM >> 4) << 24); }
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // cute::Sm100MmaPeerBitMask: address of the even (leader) CTA's copy

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
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
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 2-CTA TMA load: data lands in this CTA's shared memory, the transaction bytes are signalled on the leader's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & PEER_MASK)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((BOX_BYTES >> 4) & 0x3FFF) << 16;  // LBO: next 64-gene group
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;      // SBO: next 8 k-rows
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

struct Params {
    const uint2 *tiles;  // (ti, tj) in units of 256
    uint32_t ksplit;
    uint32_t kblocks_total;
    double *partial;     // [items][256][256]
    uint32_t chunk_kblocks;  // k-blocks (of 64 cells) per TMEM accumulation chunk
    uint32_t last_tj, last_n;  // tiles of the last tile column only need N = last_n (< 256) accumulator columns
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
    gram_tc2_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = base + STAGES * STAGE_BYTES;
    const uint32_t full_bar = bar_base, empty_bar = bar_base + 8 * STAGES;
    const uint32_t tfull_bar = bar_base + 16 * STAGES, tempty_bar = tfull_bar + 16;
    const uint32_t tmem_slot = tempty_bar + 16;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;

    const uint32_t item = blockIdx.x >> 1;
    const uint32_t tile = item / p.ksplit, ks = item % p.ksplit;
    const uint2 tij = p.tiles[tile];
    const uint32_t per = (p.kblocks_total + p.ksplit - 1) / p.ksplit;
    const uint32_t kb0 = min(ks * per, p.kblocks_total), kb1 = min(kb0 + per, p.kblocks_total);
    const uint32_t nkb = kb1 - kb0;
    const uint32_t n_cols = tij.y == p.last_tj ? p.last_n : TN;  // accumulator columns this tile really needs
    const uint32_t idesc = idesc_n(n_cols);
    const uint32_t CHUNK_KBLOCKS = p.chunk_kblocks;
    const uint32_t FLUSH_CHUNKS = max(1u, FLUSH_CELLS / (CHUNK_KBLOCKS * BK));
    const uint32_t nchunks = (nkb + CHUNK_KBLOCKS - 1) / CHUNK_KBLOCKS;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < STAGES; ++s) mbar_init(full_bar + 8 * s, 2), mbar_init(empty_bar + 8 * s, 1);
        for (uint32_t a = 0; a < 2; ++a) mbar_init(tfull_bar + 8 * a, 1), mbar_init(tempty_bar + 8 * a, 512);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs) ----------------
        if (lane == 0) {
            for (uint32_t it = 0; it < nkb; ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty_bar + 8 * s, ph ^ 1);
                if (leader) mbar_arrive_expect_tx(full_bar + 8 * s, 2 * STAGE_BYTES);
                else mbar_arrive_cluster(full_bar + 8 * s, 0);
                const int32_t row = (int32_t)((kb0 + it) * BK);
                const uint32_t sa = base + s * STAGE_BYTES;
                const uint32_t a_hi = sa, a_lo = sa + HALF_BOXES * BOX_BYTES;
                const uint32_t b_hi = sa + 2 * HALF_BOXES * BOX_BYTES, b_lo = b_hi + HALF_BOXES * BOX_BYTES;
#pragma unroll
                for (uint32_t b = 0; b < HALF_BOXES; ++b) {
                    const int32_t ga = (int32_t)(tij.x * TM + rank * 128 + b * 64);
                    const int32_t gb = (int32_t)(tij.y * TN + rank * (n_cols / 2) + b * 64);  // this CTA's N / 2 columns of B
                    tma_load_2d_pair(a_hi + b * BOX_BYTES, &map_hi, full_bar + 8 * s, ga, row);
                    tma_load_2d_pair(a_lo + b * BOX_BYTES, &map_lo, full_bar + 8 * s, ga, row);
                    tma_load_2d_pair(b_hi + b * BOX_BYTES, &map_hi, full_bar + 8 * s, gb, row);
                    tma_load_2d_pair(b_lo + b * BOX_BYTES, &map_lo, full_bar + 8 * s, gb, row);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (leader) {
            uint32_t it = 0;
            for (uint32_t c = 0; c < nchunks; ++c) {
                const uint32_t as = c & 1;
                mbar_wait(tempty_bar + 8 * as, ((c >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * TN;
                const uint32_t kend = min((c + 1) * CHUNK_KBLOCKS, nkb);
                for (uint32_t kb = c * CHUNK_KBLOCKS; kb < kend; ++kb, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(full_bar + 8 * s, ph);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = base + s * STAGE_BYTES;
                        const uint32_t a_hi = sa, a_lo = sa + HALF_BOXES * BOX_BYTES;
                        const uint32_t b_hi = sa + 2 * HALF_BOXES * BOX_BYTES, b_lo = b_hi + HALF_BOXES * BOX_BYTES;
#pragma unroll
                        for (uint32_t kk = 0; kk < BK / UMMA_K; ++kk) {
                            const uint32_t ko = kk * UMMA_K * 128;
                            const uint64_t dah = make_desc_mn_sw128(a_hi + ko), dal = make_desc_mn_sw128(a_lo + ko);
                            const uint64_t dbh = make_desc_mn_sw128(b_hi + ko), dbl = make_desc_mn_sw128(b_lo + ko);
                            const uint32_t first = (kb == c * CHUNK_KBLOCKS && kk == 0) ? 0u : 1u;
                            tc_mma_pair(d_tmem, dal, dbh, idesc, first);
                            tc_mma_pair(d_tmem, dah, dbl, idesc, 1u);
                            tc_mma_pair(d_tmem, dah, dbh, idesc, 1u);
                        }
                        tc_commit_pair(empty_bar + 8 * s);
                        if (kb + 1 == kend) tc_commit_pair(tfull_bar + 8 * as);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---------------- epilogue (both CTAs): own 128 TMEM lanes -> fp32 registers (RN) -> fp64 partial ----------------
        const uint32_t q = warp & 3;
        const uint32_t half = (warp - 2) >> 2;
        const uint32_t row = rank * 128 + q * 32 + lane;
        double *prow = p.partial + (size_t)item * (TM * TN) + (size_t)row * TN + half * 128;
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.f;
        for (uint32_t c = 0; c < nchunks; ++c) {
            const uint32_t as = c & 1;
            mbar_wait(tfull_bar + 8 * as, (c >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (uint32_t g = 0; g < 4; g += 2) {
                uint32_t r[64];
                const uint32_t taddr = tmem_base + ((q * 32u) << 16) + as * TN + half * 128 + g * 32;
#define SRB_LDTM32(R, ADDR)                                                                                                 \
                asm volatile(                                                                                               \
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                               \
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                              \
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
                    : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]),      \
                      "=r"(R[8]), "=r"(R[9]), "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), \
                      "=r"(R[16]), "=r"(R[17]), "=r"(R[18]), "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), \
                      "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]), "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31]) \
                    : "r"(ADDR)                                                                                             \
                    : "memory")
                uint32_t *r0 = r, *r1 = r + 32;
                SRB_LDTM32(r0, taddr);
                SRB_LDTM32(r1, taddr + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 64; ++i) acc[g * 32 + i] += __uint_as_float(r[i]);
            }
            tc_fence_before();
            if (leader) mbar_arrive_local(tempty_bar + 8 * as);
            else mbar_arrive_cluster(tempty_bar + 8 * as, 0);
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
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

__global__ void gram_reduce2_kernel(const double *__restrict__ partial, const uint2 *__restrict__ tiles, uint32_t ntiles,
                                    uint32_t ksplit, uint32_t dpad, double *__restrict__ G) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_tile = (uint64_t)TM * TN;
    if (e >= per_tile * ntiles) return;
    const uint32_t t = (uint32_t)(e / per_tile);
    const uint32_t r = (uint32_t)((e % per_tile) / TN), c = (uint32_t)(e % TN);
    const uint32_t i = tiles[t].x * TM + r, j = tiles[t].y * TN + c;
    if (i > j) return;
    double s = 0.0;
    for (uint32_t ks = 0; ks < ksplit; ++ks) s += partial[((size_t)t * ksplit + ks) * per_tile + (size_t)r * TN + c];
    G[(uint64_t)i * dpad + j] = s;
    G[(uint64_t)j * dpad + i] = s;
}

}  // namespace tc2

CUtensorMap gram_panel_map(const __half *X, uint64_t n, uint32_t dpad);  // gram_tc.cu

void gram_tcgen05_pair(srb_ctx *ctx, const __half *Xh, const __half *Xl, uint64_t n, uint32_t dpad, double *G, uint32_t d_used) {
    using namespace tc2;
    if (n == 0) return;
    cudaStream_t s = ctx->stream;
    SRB_REQUIRE(dpad % TN == 0, SRB_ERR_INVALID_ARG, "dpad must be a multiple of 256");
    const uint32_t NT = dpad / TN;
    std::vector<uint2> tiles;
    for (uint32_t ti = 0; ti < NT; ++ti)
        for (uint32_t tj = ti; tj < NT; ++tj) tiles.push_back(make_uint2(ti, tj));
    const uint32_t ntiles = (uint32_t)tiles.size();
    const uint32_t kblocks = (uint32_t)((n + BK - 1) / BK);
    uint32_t ksplit = std::max<uint32_t>(1, (uint32_t)(ctx->sm_count / 2) / ntiles);
    ksplit = std::min(ksplit, kblocks);
    const uint32_t items = ntiles * ksplit;
    Buf d_tiles = dev_alloc(s, sizeof(uint2) * ntiles);
    SRB_CUDA(cudaMemcpyAsync(d_tiles->p, tiles.data(), sizeof(uint2) * ntiles, cudaMemcpyHostToDevice, s));
    Buf partial = dev_zeros(s, sizeof(double) * (size_t)items * TM * TN);
    CUtensorMap mh = gram_panel_map(Xh, n, dpad), ml = gram_panel_map(Xl, n, dpad);
    Params p;
    p.tiles = d_tiles->as<uint2>();
    p.ksplit = ksplit;
    p.kblocks_total = kblocks;
    p.partial = partial->as<double>();
    {
        // the columns beyond d_used are padding: with SRB_GRAM_TRIM=1 the tiles of the last tile column run the MMA with
        // N = what is needed, rounded to 32 (each CTA of the pair supplies N / 2 columns, a multiple of 16)
        static const int trim = [] {
            const char *e = getenv("SRB_GRAM_TRIM");
            return (e && e[0] == '1') ? 1 : 0;  // off by default: see the header
        }();
        const uint32_t used = d_used ? std::min(d_used, dpad) : dpad;
        const uint32_t rem = used - (NT - 1) * TN;  // columns of the last tile column that hold selected genes
        p.last_tj = NT - 1;
        p.last_n = trim ? std::max<uint32_t>(32, std::min<uint32_t>(TN, (rem + 31) / 32 * 32)) : TN;
    }
    {
        static int chunk = -1;
        if (chunk < 0) {
            const char *e = getenv("SRB_GRAM_CHUNK");
            chunk = e ? std::max(1, atoi(e)) : 2;
        }
        p.chunk_kblocks = (uint32_t)chunk;
    }
    SRB_CUDA(cudaFuncSetAttribute(gram_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    SRB_LAUNCH(gram_tc2_kernel, items * 2, THREADS, SMEM_BYTES, s, mh, ml, p);
    const uint64_t total = (uint64_t)ntiles * TM * TN;
    SRB_LAUNCH(gram_reduce2_kernel, (unsigned)((total + 255) / 256), 256, 0, s, partial->as<double>(), d_tiles->as<uint2>(), ntiles, ksplit, dpad, G);
    SRB_CUDA(cudaStreamSynchronize(s));
}

}  // namespace srb
