// api.cu — C ABI entry points of libsrb200 (include/srb200.h), context / upload / download plumbing and the
// direction dispatch that mirrors src/shared/statistics/mod.rs (ArrayData::{CsrMatrix,CscMatrix} -> helper).
#include <algorithm>
#include <cmath>
#include <chrono>
#include <map>
#include <mutex>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "host_pack.h"

// built-in default of SRB_UPLOAD_PACK (kept in step with _ffi.UPLOAD_DEFAULT)
#define SRB_UPLOAD_DEFAULT_MODE SRB_UPLOAD_AUTO

namespace srb {

static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};
void set_last_error(const std::string &msg) { t_last_error = msg; }

// ---- device memory: per-stream block cache ----------------------------------------------------------------------
// Every buffer of a context lives on that context's single stream, so a block released by ~DevBuf can be handed to
// the next request in program order without any event bookkeeping. Steady-state steps therefore never reach the
// driver allocator (measured: cudaMallocAsync pool growth stalls of 40-700 ms per step otherwise). Best fit within
// 25 %; blocks beyond a 64 GB cache budget are returned to the driver.
struct BlockCache {
    std::multimap<size_t, void *> free_blocks;
    size_t cached_bytes = 0;
};
static std::mutex g_cache_mu;
static std::map<cudaStream_t, BlockCache> g_caches;
static constexpr size_t kCacheBudget = 64ull << 30;

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
        release_cached_blocks(s);  // give everything back and retry once
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
    BlockCache &c = g_caches[st];
    c.free_blocks.emplace(cap, p);
    c.cached_bytes += cap;
    while (c.cached_bytes > kCacheBudget && !c.free_blocks.empty()) {
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

// ---- conversion kernels ------------------------------------------------------------------------------
template <typename SRC, typename DST>
__global__ void convert_kernel(const SRC *__restrict__ src, DST *__restrict__ dst, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = (DST)src[i];
}
// narrow + bounds check; flags[0] |= 1 on out-of-range
template <typename SRC>
__global__ void narrow_index_kernel(const SRC *__restrict__ src, uint32_t *__restrict__ dst, uint64_t n, uint64_t bound,
                                    uint32_t *__restrict__ flags) {
    uint32_t bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = (uint64_t)src[i];
        bad |= (uint32_t)(v >= bound);
        dst[i] = (uint32_t)v;
    }
    if (bad) atomicOr(flags, 1u);
}
// canonical form check: offsets monotone, indices strictly increasing within a line. flags[1] |= 1 otherwise
__global__ void canonical_check_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx, uint64_t nmajor,
                                       uint64_t nnz, uint32_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t bad = 0;
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        const int64_t a = off[r], b = off[r + 1];
        if (a > b || a < 0 || (uint64_t)b > nnz) { bad = 1; continue; }
        for (int64_t k = a + 1 + lane; k < b; k += 32) bad |= (uint32_t)(idx[k] <= idx[k - 1]);
    }
    if (bad) atomicOr(flags + 1, 1u);
}

static unsigned grid_for(const srb_ctx *c, uint64_t n) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 16));
}

// host array (any supported dtype) -> device array of DST, through a bounded staging buffer
template <typename DST>
static void upload_convert(srb_ctx *c, const void *host, int dtype, uint64_t n, DST *d_dst) {
    if (n == 0) return;
    cudaStream_t s = c->stream;
    size_t esz;
    switch (dtype) {
        case SRB_I8: case SRB_U8: esz = 1; break;
        case SRB_I16: case SRB_U16: esz = 2; break;
        case SRB_I32: case SRB_U32: case SRB_F32: esz = 4; break;
        case SRB_F64: esz = 8; break;
        default: throw Error(SRB_ERR_UNSUPPORTED_DTYPE, "dtype not supported (the reference panics for I64/U64/Usize/Bool/String)");
    }
    if ((dtype == SRB_F32 && sizeof(DST) == 4) || (dtype == SRB_F64 && sizeof(DST) == 8)) {
        SRB_CUDA(cudaMemcpyAsync(d_dst, host, n * esz, cudaMemcpyHostToDevice, s));
        return;
    }
    const uint64_t chunk = std::min<uint64_t>(n, 1ull << 26);
    Buf stage = dev_alloc(s, chunk * esz);
    for (uint64_t o = 0; o < n; o += chunk) {
        const uint64_t len = std::min<uint64_t>(chunk, n - o);
        SRB_CUDA(cudaMemcpyAsync(stage->p, (const char *)host + o * esz, len * esz, cudaMemcpyHostToDevice, s));
        const unsigned g = grid_for(c, len);
        switch (dtype) {
            case SRB_I8: SRB_LAUNCH((convert_kernel<int8_t, DST>), g, 256, 0, s, stage->as<int8_t>(), d_dst + o, len); break;
            case SRB_U8: SRB_LAUNCH((convert_kernel<uint8_t, DST>), g, 256, 0, s, stage->as<uint8_t>(), d_dst + o, len); break;
            case SRB_I16: SRB_LAUNCH((convert_kernel<int16_t, DST>), g, 256, 0, s, stage->as<int16_t>(), d_dst + o, len); break;
            case SRB_U16: SRB_LAUNCH((convert_kernel<uint16_t, DST>), g, 256, 0, s, stage->as<uint16_t>(), d_dst + o, len); break;
            case SRB_I32: SRB_LAUNCH((convert_kernel<int32_t, DST>), g, 256, 0, s, stage->as<int32_t>(), d_dst + o, len); break;
            case SRB_U32: SRB_LAUNCH((convert_kernel<uint32_t, DST>), g, 256, 0, s, stage->as<uint32_t>(), d_dst + o, len); break;
            case SRB_F32: SRB_LAUNCH((convert_kernel<float, DST>), g, 256, 0, s, stage->as<float>(), d_dst + o, len); break;
            case SRB_F64: SRB_LAUNCH((convert_kernel<double, DST>), g, 256, 0, s, stage->as<double>(), d_dst + o, len); break;
        }
    }
}

static void upload_indices(srb_ctx *c, const void *host, int width, uint64_t n, uint64_t bound, uint32_t *d_dst,
                           uint32_t *d_flags) {
    if (n == 0) return;
    cudaStream_t s = c->stream;
    const uint64_t chunk = std::min<uint64_t>(n, 1ull << 26);
    Buf stage = dev_alloc(s, chunk * (size_t)width);
    for (uint64_t o = 0; o < n; o += chunk) {
        const uint64_t len = std::min<uint64_t>(chunk, n - o);
        SRB_CUDA(cudaMemcpyAsync(stage->p, (const char *)host + o * width, len * width, cudaMemcpyHostToDevice, s));
        if (width == 8)
            SRB_LAUNCH((narrow_index_kernel<uint64_t>), grid_for(c, len), 256, 0, s, stage->as<uint64_t>(), d_dst + o, len, bound, d_flags);
        else
            SRB_LAUNCH((narrow_index_kernel<uint32_t>), grid_for(c, len), 256, 0, s, stage->as<uint32_t>(), d_dst + o, len, bound, d_flags);
    }
}

// ---- packed upload ------------------------------------------------------------------------------------
// The reference's col_indices are `usize` (8 bytes); at the bench size they are 12 of the 18 GB one step moves over
// PCIe. SRB_UPLOAD_PACK=1 narrows them on the HOST (host_pack.cpp, a small thread pool) into a pinned staging ring —
// 2 bytes per entry when nminor <= 65 536, else 4 — so only 2-4 bytes per entry cross the link; packing chunk c+1
// overlaps the DMA of chunk c, and the value chunks are enqueued in between so the link never waits for the host.
// Pageable caller memory (a Rust Vec) is staged through the same ring with a threaded memcpy instead of the
// driver's single-threaded bounce buffer. The device widens to u32 and repeats the bounds check.
static int upload_pack_mode() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SRB_UPLOAD_PACK");
        if (!e) v = SRB_UPLOAD_DEFAULT_MODE;
        else if (!strcmp(e, "auto")) v = SRB_UPLOAD_AUTO;
        else if (!strcmp(e, "values")) v = SRB_UPLOAD_HOST_PACK_VALUES;
        else if (!strcmp(e, "adaptive")) v = SRB_UPLOAD_HOST_PACK_ADAPTIVE;
        else if (!strcmp(e, "delta")) v = SRB_UPLOAD_HOST_PACK_DELTA;
        else v = atoi(e) != 0 ? SRB_UPLOAD_HOST_PACK : SRB_UPLOAD_DEVICE_NARROW;
    }
    return v;
}
// host threads one context may use for packing: the ranks of a node share its cores
static int upload_threads(const srb_ctx *c) { return std::max(1, host_pack_threads() / std::max(1, c->nranks)); }
// AUTO: packing pays when the host narrows faster than the link moves the unpacked array (12 B per entry at ~55 GB/s =
// 4.6 G entries/s; one host thread packs ~0.9 G entries/s), i.e. with >= 6 threads, and only for arrays worth a ring
static int effective_upload_mode(const srb_ctx *c, uint64_t nnz) {
    const int mode = c->upload_mode >= 0 ? c->upload_mode : upload_pack_mode();
    if (mode == SRB_UPLOAD_AUTO) return (nnz >= (1ull << 20) && upload_threads(c) >= 6) ? SRB_UPLOAD_HOST_PACK : SRB_UPLOAD_DEVICE_NARROW;
    return mode;
}
static bool host_is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
static void ensure_upload_ring(srb_ctx *c, size_t bytes) {
    if (c->up_ring_bytes >= bytes) return;
    if (c->up_ring) {
        SRB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFreeHost(c->up_ring);
        c->up_ring = nullptr, c->up_ring_bytes = 0;
    }
    SRB_CUDA(cudaHostAlloc(&c->up_ring, bytes, cudaHostAllocDefault));
    c->up_ring_bytes = bytes;
    for (int i = 0; i < srb_ctx::kUpSlots; ++i)
        if (!c->up_ev[i]) SRB_CUDA(cudaEventCreateWithFlags(&c->up_ev[i], cudaEventDisableTiming));
}
// f32 chunk values that travelled as u8 / u16 (widths[chunk] = 1 | 2; 0 = the chunk was copied raw): rebuild the f32 array.
// Packed chunk c sits at byte offset 2 * c * chunk of `pk`.
__global__ void unpack_values_kernel(const uint8_t *__restrict__ pk, float *__restrict__ out, uint64_t n, int chunk_shift,
                                     const uint8_t *__restrict__ widths) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = i >> chunk_shift, base = c << chunk_shift;  // a single chunk (n < 2^shift) has c = 0
        const uint8_t w = widths[c];
        if (w == 1) out[i] = (float)pk[2 * base + (i - base)];
        else if (w == 2) out[i] = (float)reinterpret_cast<const uint16_t *>(pk)[i];
    }
}
// HOST_PACK_DELTA: rebuild the u32 indices from the one-byte gap codes (host_pack.cpp). One warp per line: a segmented
// inclusive scan in which an escape (code 255: the full index sits in the sorted side list) restarts the running sum.
__global__ void delta_decode_kernel(const uint8_t *__restrict__ code, const int64_t *__restrict__ off, uint64_t nmajor,
                                    const uint64_t *__restrict__ esc_pos, const uint32_t *__restrict__ esc_val, uint64_t n_esc,
                                    uint64_t bound, uint32_t *__restrict__ out, uint32_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t bad = 0;
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        const int64_t a = off[r], b = off[r + 1];
        uint32_t carry = 0;
        for (int64_t base = a; base < b; base += 32) {
            const int64_t i = base + lane;
            const bool valid = i < b;
            uint32_t v = valid ? code[i] : 0u;
            int reset = 0;
            if (valid && v == 255u) {  // binary search of the escape list for position i
                uint64_t lo = 0, hi = n_esc;
                while (lo < hi) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if (esc_pos[mid] < (uint64_t)i) lo = mid + 1;
                    else hi = mid;
                }
                if (lo < n_esc && esc_pos[lo] == (uint64_t)i) v = esc_val[lo];
                else bad = 1;  // cannot happen for codes produced by host_delta_encode
                reset = 1;
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t pv = __shfl_up_sync(0xffffffffu, v, o);
                const int pr = __shfl_up_sync(0xffffffffu, reset, o);
                if (lane >= o && !reset) v += pv, reset = pr;
            }
            const uint32_t col = reset ? v : v + carry;
            if (valid) {
                out[i] = col;
                bad |= (uint32_t)((uint64_t)col >= bound);
            }
            carry = __shfl_sync(0xffffffffu, col, 31);
        }
    }
    if (bad) atomicOr(flags, 1u);
}

// indices (always) and, when `values` is a bit-copy of the device storage (vsz bytes per entry), the values too.
// Returns the bytes that crossed the link.
static uint64_t upload_packed(srb_ctx *c, const void *indices, int width, uint64_t n, uint64_t bound, uint32_t *d_idx,
                              uint32_t *d_flags, const void *values, size_t vsz, void *d_val, int value_packing /* 0 never, 1 always, 2 when the host is ahead of the link */,
                              const void *offsets = nullptr, uint64_t nmajor = 0, const int64_t *d_offsets = nullptr) {
    if (n == 0) return 0;
    cudaStream_t s = c->stream;
    // delta coding (one byte per entry) needs trustworthy offsets to walk the lines; otherwise plain narrowing
    const bool delta = offsets && d_offsets && host_offsets_valid(offsets, width, nmajor, n);
    const int pw = delta ? 1 : (bound <= 65536 ? 2 : 4);
    DeltaEscapes esc;
    const int nthreads = upload_threads(c);
    const bool stage_vals = values && host_is_pageable(values);
    const bool pack_vals = values && vsz == 4 && value_packing != 0;  // f32 counts -> u8 / u16 where lossless
    constexpr int kChunkShift = 22;
    const uint64_t chunk = std::min<uint64_t>(n, 1ull << kChunkShift);
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    const size_t idx_bytes = (chunk * pw + 255) & ~size_t(255);
    const size_t val_bytes = stage_vals ? chunk * vsz : (pack_vals ? chunk * 2 : 0);
    const size_t slot_bytes = idx_bytes + ((val_bytes + 255) & ~size_t(255));
    ensure_upload_ring(c, slot_bytes * srb_ctx::kUpSlots);
    Buf dpk, dvpk;
    if (pw < 4) dpk = dev_alloc(s, n * pw);
    if (pack_vals) dvpk = dev_alloc(s, nchunks * chunk * 2);
    char *d_pk = pw < 4 ? dpk->as<char>() : (char *)d_idx;
    std::vector<uint8_t> widths(nchunks, 0);
    int vstate = pack_vals ? 1 : 0;  // 1: try u8, 2: try u16, 0: raw (sticky: a chunk that refuses widens all later ones)
    bool oob = false, any_packed = false;
    uint64_t ci = 0, link = 0;
    for (uint64_t o = 0; o < n; o += chunk, ++ci) {
        const uint64_t len = std::min<uint64_t>(chunk, n - o);
        const int slot = (int)(ci % srb_ctx::kUpSlots);
        // ADAPTIVE: the slot's previous DMA still running means the host is ahead of the link, so this chunk can afford the
        // extra host pass that halves its link bytes; when the link is the one waiting, the values go raw
        bool pack_this = value_packing == 1;
        if (value_packing == 2 && c->up_ev_used[slot]) pack_this = cudaEventQuery(c->up_ev[slot]) == cudaErrorNotReady;
        if (c->up_ev_used[slot]) SRB_CUDA(cudaEventSynchronize(c->up_ev[slot]));  // the slot's previous DMA is done
        char *h_idx = (char *)c->up_ring + slot_bytes * slot, *h_val = h_idx + idx_bytes;
        if (delta) oob |= host_delta_encode(indices, offsets, width, nmajor, o, len, (uint8_t *)h_idx, bound, nthreads, esc);
        else oob |= host_pack_indices((const char *)indices + o * width, width, len, h_idx, pw, bound, nthreads);
        SRB_CUDA(cudaMemcpyAsync(d_pk + o * pw, h_idx, len * pw, cudaMemcpyHostToDevice, s));
        link += len * pw;
        if (values) {
            const char *src = (const char *)values + o * vsz;
            int w = 0;
            while (vstate && pack_this) {
                if (host_pack_values_f32((const float *)src, len, h_val, vstate, nthreads)) {
                    w = vstate;
                    break;
                }
                vstate = vstate == 1 ? 2 : 0;
            }
            widths[ci] = (uint8_t)w;
            if (w) {
                SRB_CUDA(cudaMemcpyAsync(dvpk->as<char>() + 2 * o, h_val, len * w, cudaMemcpyHostToDevice, s));
                link += len * w;
                any_packed = true;
            } else {
                if (stage_vals) {
                    host_copy_parallel(src, h_val, len * vsz, nthreads);
                    src = h_val;
                }
                SRB_CUDA(cudaMemcpyAsync((char *)d_val + o * vsz, src, len * vsz, cudaMemcpyHostToDevice, s));
                link += len * vsz;
            }
        }
        SRB_CUDA(cudaEventRecord(c->up_ev[slot], s));
        c->up_ev_used[slot] = true;
    }
    bool sync_needed = false;
    if (delta) {
        const uint64_t ne = esc.pos.size();
        Buf dpos = dev_alloc(s, 8 * std::max<uint64_t>(ne, 1)), dval = dev_alloc(s, 4 * std::max<uint64_t>(ne, 1));
        if (ne) {
            SRB_CUDA(cudaMemcpyAsync(dpos->p, esc.pos.data(), 8 * ne, cudaMemcpyHostToDevice, s));
            SRB_CUDA(cudaMemcpyAsync(dval->p, esc.val.data(), 4 * ne, cudaMemcpyHostToDevice, s));
            link += 12 * ne;
        }
        const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nmajor + 7) / 8, (uint64_t)c->sm_count * 16));
        SRB_LAUNCH(delta_decode_kernel, g, 256, 0, s, dpk->as<uint8_t>(), d_offsets, nmajor, dpos->as<uint64_t>(), dval->as<uint32_t>(), ne, bound, d_idx, d_flags);
        sync_needed = true;  // `esc` (pageable) and the escape buffers are done with
    } else if (pw == 2) {
        SRB_LAUNCH((narrow_index_kernel<uint16_t>), grid_for(c, n), 256, 0, s, dpk->as<uint16_t>(), d_idx, n, bound, d_flags);
    }
    if (sync_needed) SRB_CUDA(cudaStreamSynchronize(s));
    if (any_packed) {
        Buf dw = dev_alloc(s, nchunks);
        SRB_CUDA(cudaMemcpyAsync(dw->p, widths.data(), nchunks, cudaMemcpyHostToDevice, s));
        SRB_LAUNCH(unpack_values_kernel, grid_for(c, n), 256, 0, s, dvpk->as<uint8_t>(), (float *)d_val, n, kChunkShift, dw->as<uint8_t>());
        SRB_CUDA(cudaStreamSynchronize(s));  // `widths` (pageable) and the staging ring are done with
    }
    if (oob) {
        SRB_CUDA(cudaStreamSynchronize(s));
        throw Error(SRB_ERR_INDEX_OOB, "minor index out of bounds");
    }
    return link;
}

static void check_mat(const srb_mat *m) { SRB_REQUIRE(m && m->ctx && m->st, SRB_ERR_INVALID_ARG, "null matrix handle"); }
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
    *out = c.release();
    SRB_API_END
}

int32_t srb_ctx_destroy(srb_ctx *ctx) {
    SRB_API_BEGIN
    if (!ctx) return SRB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    release_cached_blocks(ctx->stream);
    comm_destroy(ctx);
    eig_destroy(ctx);
    if (ctx->up_ring) cudaFreeHost(ctx->up_ring);
    for (int i = 0; i < srb_ctx::kUpSlots; ++i)
        if (ctx->up_ev[i]) cudaEventDestroy(ctx->up_ev[i]);
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

int32_t srb_ctx_set_upload_mode(srb_ctx *ctx, int32_t mode) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    SRB_REQUIRE(mode == SRB_UPLOAD_DEVICE_NARROW || mode == SRB_UPLOAD_HOST_PACK || mode == SRB_UPLOAD_AUTO ||
                    mode == SRB_UPLOAD_HOST_PACK_VALUES || mode == SRB_UPLOAD_HOST_PACK_ADAPTIVE || mode == SRB_UPLOAD_HOST_PACK_DELTA,
                SRB_ERR_INVALID_ARG, "bad upload mode");
    ctx->upload_mode = mode;
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

int32_t srb_ctx_last_upload(srb_ctx *ctx, uint64_t *h2d_bytes, int32_t *host_packed) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    if (h2d_bytes) *h2d_bytes = ctx->last_upload_h2d;
    if (host_packed) *host_packed = ctx->last_upload_packed;
    SRB_API_END
}

int32_t srb_ctx_synchronize(srb_ctx *ctx) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    SRB_CUDA(cudaStreamSynchronize(ctx->stream));
    SRB_API_END
}

void *srb_ctx_stream(srb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int32_t srb_mat_upload(srb_ctx *ctx, int32_t format, uint64_t nrows, uint64_t ncols, uint64_t nnz, const void *offsets,
                       const void *indices, int32_t idx_width, const void *values, int32_t dtype, srb_mat **out) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx && out, SRB_ERR_INVALID_ARG, "null ctx/out");
    SRB_REQUIRE(format == SRB_CSR || format == SRB_CSC, SRB_ERR_INVALID_ARG, "format must be CSR or CSC");
    SRB_REQUIRE(idx_width == 4 || idx_width == 8, SRB_ERR_INVALID_ARG, "idx_width must be 4 or 8");
    SRB_REQUIRE(offsets && (nnz == 0 || (indices && values)), SRB_ERR_INVALID_ARG, "null array");
    SRB_REQUIRE(dtype != SRB_I64 && dtype != SRB_U64 && dtype >= 0 && dtype <= SRB_F64, SRB_ERR_UNSUPPORTED_DTYPE,
                "dtype not supported (the reference panics for I64/U64/Usize/Bool/String, shared/mod.rs:117-126)");
    SRB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint64_t nmajor = format == SRB_CSR ? nrows : ncols;
    const uint64_t nminor = format == SRB_CSR ? ncols : nrows;
    SRB_REQUIRE(nminor < (1ull << 32) && nmajor < (1ull << 40), SRB_ERR_INVALID_ARG, "matrix too large");
    auto st = std::make_shared<Structure>();
    st->nmajor = nmajor, st->nminor = nminor, st->nnz = nnz;
    st->offsets = dev_alloc(s, sizeof(int64_t) * (nmajor + 1));
    st->indices = dev_alloc(s, sizeof(uint32_t) * (nnz ? nnz : 1));
    Buf flags = dev_zeros(s, sizeof(uint32_t) * 2);
    if (idx_width == 8) {
        // u64 -> i64 is a bit copy
        SRB_CUDA(cudaMemcpyAsync(st->offsets->p, offsets, sizeof(int64_t) * (nmajor + 1), cudaMemcpyHostToDevice, s));
    } else {
        upload_convert<int64_t>(ctx, offsets, SRB_U32, nmajor + 1, st->offsets->as<int64_t>());
    }
    const int up_mode = effective_upload_mode(ctx, nnz);
    const bool packed = up_mode == SRB_UPLOAD_HOST_PACK || up_mode == SRB_UPLOAD_HOST_PACK_VALUES || up_mode == SRB_UPLOAD_HOST_PACK_ADAPTIVE ||
                        up_mode == SRB_UPLOAD_HOST_PACK_DELTA;
    if (!packed) upload_indices(ctx, indices, idx_width, nnz, nminor, st->indices->as<uint32_t>(), flags->as<uint32_t>());
    std::unique_ptr<srb_mat> m(new srb_mat());
    m->ctx = ctx, m->format = format, m->nrows = nrows, m->ncols = ncols, m->st = st;
    m->src_dtype = dtype;
    m->global_row0 = 0, m->global_nrows = nrows;
    const bool f32_exact = dtype == SRB_I8 || dtype == SRB_U8 || dtype == SRB_I16 || dtype == SRB_U16 || dtype == SRB_F32;
    m->vdtype = f32_exact ? SRB_F32 : SRB_F64;
    m->values = dev_alloc(s, (f32_exact ? 4 : 8) * (nnz ? nnz : 1));
    // values whose host dtype is the device storage dtype travel as they are, interleaved with the index chunks
    const bool direct = (dtype == SRB_F32 && f32_exact) || (dtype == SRB_F64 && !f32_exact);
    uint64_t link_bytes = 0;
    if (packed)
        link_bytes = upload_packed(ctx, indices, idx_width, nnz, nminor, st->indices->as<uint32_t>(), flags->as<uint32_t>(),
                                   direct ? values : nullptr, f32_exact ? 4 : 8, m->values->p,
                                   up_mode == SRB_UPLOAD_HOST_PACK_VALUES ? 1 : up_mode == SRB_UPLOAD_HOST_PACK_ADAPTIVE ? 2 : 0,
                                   up_mode == SRB_UPLOAD_HOST_PACK_DELTA ? offsets : nullptr, nmajor, st->offsets->as<int64_t>());
    if (!(packed && direct)) {
        if (f32_exact) upload_convert<float>(ctx, values, dtype, nnz, m->values->as<float>());
        else upload_convert<double>(ctx, values, dtype, nnz, m->values->as<double>());
    }
    {
        static const size_t esz[10] = {1, 2, 4, 8, 1, 2, 4, 8, 4, 8};
        ctx->last_upload_h2d = (uint64_t)idx_width * (nmajor + 1) +
                               (packed ? link_bytes + (direct ? 0 : esz[dtype] * nnz) : ((uint64_t)idx_width + esz[dtype]) * nnz);
        ctx->last_upload_packed = packed ? 1 : 0;
    }
    if (nmajor) SRB_LAUNCH(canonical_check_kernel, grid_for(ctx, nmajor * 32), 256, 0, s, st->offsets->as<int64_t>(), st->indices->as<uint32_t>(), nmajor, nnz, flags->as<uint32_t>());
    uint32_t hflags[2];
    int64_t last = 0;
    SRB_CUDA(cudaMemcpyAsync(hflags, flags->p, sizeof(hflags), cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaMemcpyAsync(&last, st->offsets->as<int64_t>() + nmajor, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    SRB_CUDA(cudaStreamSynchronize(s));
    SRB_REQUIRE(!hflags[0], SRB_ERR_INDEX_OOB, "minor index out of bounds");
    SRB_REQUIRE((uint64_t)last == nnz, SRB_ERR_INVALID_ARG, "offsets[nmajor] != nnz");
    SRB_REQUIRE(!hflags[1], SRB_ERR_UNSUPPORTED, "non-canonical matrix (unsorted/duplicate indices or non-monotone offsets): the reference answers CsrNonCanonical with todo!()");
    *out = m.release();
    SRB_API_END
}

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
        cudaSetDevice(m->ctx->device);
        delete m;
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
