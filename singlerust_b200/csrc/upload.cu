// upload.cu — the host -> HBM path of libsrb200: srb_mat_upload (include/srb200.h) and everything behind it. The
// reference hands over nalgebra-sparse slices (`usize` offsets / indices, values of any DynCsrMatrix dtype;
// src/shared/statistics/helper/csr.rs:24,32,96); the device copy is int64 offsets, uint32 indices and f32 / f64 values,
// bounds- and canonical-form-checked. Because PCIe, not the GPU, bounds an end-to-end step, the index array can be packed
// on the host before it crosses the link (host_pack.cpp) — see the srb_upload_mode comment in the header.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <vector>

#include "common.cuh"
#include "host_pack.h"

// built-in default of SRB_UPLOAD_PACK
#define SRB_UPLOAD_DEFAULT_MODE SRB_UPLOAD_AUTO

namespace srb {

// ---- conversion kernels ------------------------------------------------------------------------------
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
// canonical form check: offsets start at 0 and are monotone, indices strictly increasing within a line. flags[1] |= 1 otherwise
__global__ void canonical_check_kernel(const int64_t *__restrict__ off, const uint32_t *__restrict__ idx, uint64_t nmajor,
                                       uint64_t nnz, uint32_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t bad = 0;
    if (warp == 0 && off[0] != 0) bad = 1;  // entries before offsets[0] would belong to no line (nalgebra-sparse rejects it)
    for (uint64_t r = warp; r < nmajor; r += nwarps) {
        const int64_t a = off[r], b = off[r + 1];
        if (a > b || a < 0 || (uint64_t)b > nnz) { bad = 1; continue; }
        for (int64_t k = a + 1 + lane; k < b; k += 32) bad |= (uint32_t)(idx[k] <= idx[k - 1]);
    }
    if (bad) atomicOr(flags + 1, 1u);
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
        else if (!strcmp(e, "balanced")) v = SRB_UPLOAD_BALANCED;
        else v = atoi(e) != 0 ? SRB_UPLOAD_HOST_PACK : SRB_UPLOAD_DEVICE_NARROW;
    }
    return v;
}
// host threads one context may use for packing: the ranks of a node share its cores
static int upload_threads(const srb_ctx *c) { return std::max(1, host_pack_threads() / std::max(1, c->nranks)); }
// AUTO: the balanced upload decides chunk by chunk from the measured packing time against the queued link work, whatever
// the number of host threads per rank; arrays too small to be worth a staging ring are narrowed on the device
static int effective_upload_mode(const srb_ctx *c, uint64_t nnz) {
    const int mode = c->upload_mode >= 0 ? c->upload_mode : upload_pack_mode();
    if (mode == SRB_UPLOAD_AUTO) return nnz >= (1ull << 20) ? SRB_UPLOAD_BALANCED : SRB_UPLOAD_DEVICE_NARROW;
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
// raw_chunk (may be null): raw_chunk[i >> chunk_shift] != 0 marks a chunk of entries that travelled unpacked — out[i]
// already holds their index (narrowed by its own kernel), which takes part in the scan as a restart value.
__global__ void delta_decode_kernel(const uint8_t *__restrict__ code, const int64_t *__restrict__ off, uint64_t nmajor,
                                    const uint64_t *__restrict__ esc_pos, const uint32_t *__restrict__ esc_val, uint64_t n_esc,
                                    uint64_t bound, uint32_t *__restrict__ out, uint32_t *__restrict__ flags,
                                    const uint8_t *__restrict__ raw_chunk, int chunk_shift) {
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
            const bool raw = valid && raw_chunk && raw_chunk[i >> chunk_shift];
            uint32_t v = valid ? (raw ? out[i] : (uint32_t)code[i]) : 0u;
            int reset = raw ? 1 : 0;
            if (valid && !raw && v == 255u) {  // binary search of the escape list for position i
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
            if (valid && !raw) {
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
        SRB_LAUNCH(delta_decode_kernel, g, 256, 0, s, dpk->as<uint8_t>(), d_offsets, nmajor, dpos->as<uint64_t>(), dval->as<uint32_t>(), ne, bound, d_idx, d_flags,
                   (const uint8_t *)nullptr, 0);
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


// ---- balanced upload (SRB_UPLOAD_BALANCED, what AUTO resolves to) ----------------------------------------------------
// Every chunk of 4 M entries decides for itself how it crosses PCIe: packed on the host (indices: one-byte gap codes, or
// 2 / 4-byte narrowing when the offsets cannot be trusted; f32 count values: u8 / u16 where lossless) or raw (the caller's
// bytes as they are, narrowed on the device). Packing trades host time for link bytes, so a chunk is packed exactly when
// the link still has at least that much work queued (bytes enqueued but not yet copied / the link rate >= the running
// estimate of the packing time): with many cores per GPU everything is packed and the link carries 5 B per entry instead of
// 12; with 2 threads per rank (8 ranks on a 16-core host) most chunks go raw and neither side waits for the other.
// Index packing is decided first (it saves 7 link bytes per 8 host bytes read; value packing 3 per 4). Pageable caller
// memory is always packed (staging it raw would cost the host more than coding it).
static double link_rate_bytes_per_ms() {  // read per upload (not cached): tests steer the raw / packed mix with it
    const char *e = getenv("SRB_LINK_GBS");
    const double g = e ? atof(e) : 50.0;
    return (g > 0.0 ? g : 50.0) * 1e6;
}
static double host_now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static uint64_t upload_balanced(srb_ctx *c, const void *indices, int width, uint64_t n, uint64_t bound, uint32_t *d_idx, uint32_t *d_flags,
                                const void *values, size_t vsz, void *d_val, const void *offsets, uint64_t nmajor, const int64_t *d_offsets) {
    c->last_upload_chunks = c->last_upload_idx_packed = c->last_upload_val_packed = 0;
    if (n == 0) return 0;
    cudaStream_t s = c->stream;
    const int nthreads = upload_threads(c);
    const bool delta = offsets && d_offsets && host_offsets_valid(offsets, width, nmajor, n);
    const int pw = delta ? 1 : (bound <= 65536 ? 2 : 4);
    const bool idx_pageable = host_is_pageable(indices), val_pageable = values && host_is_pageable(values);
    const bool idx_pack_pointless = !delta && pw == width;  // u32 source, wide minor dimension: packed == raw
    const bool vals_packable = values && vsz == 4;
    constexpr int kChunkShift = 22;
    const uint64_t chunk = std::min<uint64_t>(n, 1ull << kChunkShift);
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    // ring slot: [index part: packed codes, or the raw integers of a pageable source][value part: packed, or staged raw]
    const size_t idx_bytes = (chunk * (size_t)(idx_pageable ? width : pw) + 255) & ~size_t(255);
    const size_t val_bytes = values ? (chunk * (val_pageable ? vsz : 2) + 255) & ~size_t(255) : 0;
    const size_t slot_bytes = idx_bytes + val_bytes;
    ensure_upload_ring(c, slot_bytes * srb_ctx::kUpSlots);
    for (int i = 0; i < srb_ctx::kUpChunkEvents; ++i) {
        if (!c->up_cev[i]) SRB_CUDA(cudaEventCreate(&c->up_cev[i]));  // timing enabled: the copies' own duration gives the link rate
        if (!c->up_sev[i]) SRB_CUDA(cudaEventCreate(&c->up_sev[i]));
    }
    Buf dpk = pw < 4 ? dev_alloc(s, n * pw) : Buf();            // packed index codes (pw = 4 lands in d_idx directly)
    Buf draw = dev_alloc(s, chunk * (size_t)width);              // raw index chunk, narrowed right after its copy
    Buf dvpk = vals_packable ? dev_alloc(s, nchunks * chunk * 2) : Buf();
    std::vector<uint8_t> widths(nchunks, 0), raw_idx(nchunks, 0);
    DeltaEscapes esc;
    int vstate = vals_packable ? 1 : 0;  // 1: try u8, 2: try u16, 0: raw (sticky: a chunk that refuses widens all later ones)
    bool oob = false, any_val_packed = false, any_idx_raw = false;
    uint64_t link = 0, done = 0, enq_bytes = 0, done_bytes = 0;
    std::vector<uint64_t> cum_bytes(nchunks + 1, 0);  // bytes enqueued up to and including chunk i - 1
    double t_idx = 0.0, t_val = 0.0;                  // running estimates of the packing time of one chunk (ms)
    double acc_raw = 0.0, acc_val = 0.0;              // dithering accumulators of the raw-index / packed-value fractions
    uint64_t n_rate = 0;                              // link-rate samples taken
    int slot_use = 0;
    // link rate: SRB_LINK_GBS (default 50) until the first chunks have completed, then the running mean of what the copies
    // of this upload really achieved (bytes / duration between the chunk's start and end events). With 8 GPUs uploading at
    // once each gets ~18 GB/s of the host's memory and PCIe fabric, not 50, and a model that assumes 50 sends too much raw.
    const bool rate_fixed = getenv("SRB_LINK_GBS") != nullptr;
    double rate = link_rate_bytes_per_ms();
    uint64_t ci = 0;
    for (uint64_t o = 0; o < n; o += chunk, ++ci) {
        const uint64_t len = std::min<uint64_t>(chunk, n - o);
        // retire completed chunks (keeps the per-chunk events reusable; the decision itself is model-based, see below)
        while (done < ci && cudaEventQuery(c->up_cev[done % srb_ctx::kUpChunkEvents]) == cudaSuccess) {
            float ms = 0.f;
            const uint64_t cb = cum_bytes[done + 1] - cum_bytes[done];
            if (!rate_fixed && cb && cudaEventElapsedTime(&ms, c->up_sev[done % srb_ctx::kUpChunkEvents], c->up_cev[done % srb_ctx::kUpChunkEvents]) == cudaSuccess && ms > 0.f) {
                const double r = (double)cb / (double)ms;
                rate = n_rate++ == 0 ? r : 0.75 * rate + 0.25 * r;
            }
            ++done;
        }
        (void)cudaGetLastError();  // cudaErrorNotReady is not an error here
        done_bytes = cum_bytes[done];
        // Rate model. Per chunk of `len` entries: packing the indices costs the host t_idx (measured, running mean, with
        // the DMA engine competing for the same memory) and the link len*pw bytes; sending them raw costs the host nothing
        // and the link len*width bytes. With the values raw, host and link finish together when a fraction
        //     g = (t_idx - t_k) / (t_idx - t_k + t_r)        t_k, t_r = link time of a packed / raw chunk (indices + values)
        // of the chunks goes raw (g = 0 when the host packs faster than the link drains). Only when the host is FASTER than
        // the link is the spare host time spent on packing values, for the fraction f of chunks that equalises the two.
        double g = 0.0, f = 0.0;
        upload_mix(t_idx, t_val, len, pw, width, values ? vsz : 0, vstate, rate, &g, &f);
        // every 32nd chunk is packed regardless, so a pessimistic first timing (cold pages, pool start-up) cannot lock
        // the upload into the raw mode
        const bool probe = ci % 32 == 0;
        acc_raw += g, acc_val += f;
        bool pack_idx = !idx_pack_pointless;
        if (pack_idx && !idx_pageable && !probe && acc_raw >= 1.0) pack_idx = false, acc_raw -= 1.0;
        bool try_val = false;
        if (vstate != 0 && (probe || acc_val >= 1.0)) {
            try_val = true;
            if (!probe) acc_val -= 1.0;
        }
        // pageable values would have to be staged through the ring anyway (read 4 B + write 4 B per entry on the host):
        // packing them (read 4 B + write 1 B) is the cheaper way to get them there
        if (vstate != 0 && val_pageable) try_val = true;
        const bool need_slot = pack_idx || try_val || (values && val_pageable) || idx_pageable;
        char *h_idx = nullptr, *h_val = nullptr;
        if (need_slot) {
            const int slot = slot_use++ % srb_ctx::kUpSlots;
            if (c->up_ev_used[slot]) SRB_CUDA(cudaEventSynchronize(c->up_ev[slot]));  // the slot's previous copies are done
            h_idx = (char *)c->up_ring + slot_bytes * slot, h_val = h_idx + idx_bytes;
        }
        uint64_t bytes = 0;
        if (ci >= (uint64_t)srb_ctx::kUpChunkEvents && done + srb_ctx::kUpChunkEvents <= ci) {
            // the events about to be reused belong to a chunk that has not been seen complete yet
            SRB_CUDA(cudaEventSynchronize(c->up_cev[ci % srb_ctx::kUpChunkEvents]));
            done = ci - srb_ctx::kUpChunkEvents + 1;
        }
        // ---- indices ----
        if (pack_idx) {
            const double t0 = host_now_ms();
            if (delta) oob |= host_delta_encode(indices, offsets, width, nmajor, o, len, (uint8_t *)h_idx, bound, nthreads, esc);
            else oob |= host_pack_indices((const char *)indices + o * width, width, len, h_idx, pw, bound, nthreads);
            const double dt = (host_now_ms() - t0) * (double)chunk / (double)len;
            t_idx = t_idx == 0.0 ? dt : 0.5 * t_idx + 0.5 * dt;
            char *dst = pw < 4 ? dpk->as<char>() + o * pw : (char *)(d_idx + o);
            SRB_CUDA(cudaEventRecord(c->up_sev[ci % srb_ctx::kUpChunkEvents], s));  // after packing: the copies' own duration
            SRB_CUDA(cudaMemcpyAsync(dst, h_idx, len * pw, cudaMemcpyHostToDevice, s));
            if (pw == 2) SRB_LAUNCH((narrow_index_kernel<uint16_t>), grid_for(c, len), 256, 0, s, dpk->as<uint16_t>() + o, d_idx + o, len, bound, d_flags);
            bytes += len * pw;
            ++c->last_upload_idx_packed;
        } else {
            // pinned (or pack-pointless) source: the DMA engine reads the caller's array directly
            const char *src = (const char *)indices + o * width;
            if (idx_pageable && h_idx) {
                host_copy_parallel(src, h_idx, len * width, nthreads);
                src = h_idx;
            }
            SRB_CUDA(cudaEventRecord(c->up_sev[ci % srb_ctx::kUpChunkEvents], s));
            SRB_CUDA(cudaMemcpyAsync(draw->p, src, len * width, cudaMemcpyHostToDevice, s));
            if (width == 8) SRB_LAUNCH((narrow_index_kernel<uint64_t>), grid_for(c, len), 256, 0, s, draw->as<uint64_t>(), d_idx + o, len, bound, d_flags);
            else SRB_LAUNCH((narrow_index_kernel<uint32_t>), grid_for(c, len), 256, 0, s, draw->as<uint32_t>(), d_idx + o, len, bound, d_flags);
            bytes += len * width;
            raw_idx[ci] = 1, any_idx_raw = true;
        }
        // ---- values ----
        if (values) {
            const char *src = (const char *)values + o * vsz;
            int w = 0;
            if (try_val) {
                const double t0 = host_now_ms();
                while (vstate) {
                    if (host_pack_values_f32((const float *)src, len, h_val, vstate, nthreads)) {
                        w = vstate;
                        break;
                    }
                    vstate = vstate == 1 ? 2 : 0;
                }
                const double dt = (host_now_ms() - t0) * (double)chunk / (double)len;
                t_val = t_val == 0.0 ? dt : 0.5 * t_val + 0.5 * dt;
            }
            widths[ci] = (uint8_t)w;
            if (w) {
                SRB_CUDA(cudaMemcpyAsync(dvpk->as<char>() + 2 * o, h_val, len * w, cudaMemcpyHostToDevice, s));
                bytes += len * w;
                any_val_packed = true;
                ++c->last_upload_val_packed;
            } else {
                if (val_pageable) {
                    host_copy_parallel(src, h_val, len * vsz, nthreads);
                    src = h_val;
                }
                SRB_CUDA(cudaMemcpyAsync((char *)d_val + o * vsz, src, len * vsz, cudaMemcpyHostToDevice, s));
                bytes += len * vsz;
            }
        }
        if (need_slot) {
            const int slot = (slot_use - 1) % srb_ctx::kUpSlots;
            SRB_CUDA(cudaEventRecord(c->up_ev[slot], s));
            c->up_ev_used[slot] = true;
        }
        SRB_CUDA(cudaEventRecord(c->up_cev[ci % srb_ctx::kUpChunkEvents], s));
        enq_bytes += bytes, link += bytes;
        cum_bytes[ci + 1] = enq_bytes;
    }
    c->last_upload_chunks = (int)nchunks;
    bool sync_needed = false;
    if (delta && c->last_upload_idx_packed) {
        const uint64_t ne = esc.pos.size();
        Buf dpos = dev_alloc(s, 8 * std::max<uint64_t>(ne, 1)), dval = dev_alloc(s, 4 * std::max<uint64_t>(ne, 1));
        if (ne) {
            SRB_CUDA(cudaMemcpyAsync(dpos->p, esc.pos.data(), 8 * ne, cudaMemcpyHostToDevice, s));
            SRB_CUDA(cudaMemcpyAsync(dval->p, esc.val.data(), 4 * ne, cudaMemcpyHostToDevice, s));
            link += 12 * ne;
        }
        Buf draw_flags;
        if (any_idx_raw) {
            draw_flags = dev_alloc(s, nchunks);
            SRB_CUDA(cudaMemcpyAsync(draw_flags->p, raw_idx.data(), nchunks, cudaMemcpyHostToDevice, s));
        }
        const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nmajor + 7) / 8, (uint64_t)c->sm_count * 16));
        SRB_LAUNCH(delta_decode_kernel, g, 256, 0, s, dpk->as<uint8_t>(), d_offsets, nmajor, dpos->as<uint64_t>(), dval->as<uint32_t>(), ne, bound, d_idx,
                   d_flags, any_idx_raw ? draw_flags->as<uint8_t>() : (const uint8_t *)nullptr, kChunkShift);
        sync_needed = true;  // `esc`, `raw_idx` (pageable) and the escape buffers are done with
    }
    if (any_val_packed) {
        Buf dw = dev_alloc(s, nchunks);
        SRB_CUDA(cudaMemcpyAsync(dw->p, widths.data(), nchunks, cudaMemcpyHostToDevice, s));
        SRB_LAUNCH(unpack_values_kernel, grid_for(c, n), 256, 0, s, dvpk->as<uint8_t>(), (float *)d_val, n, kChunkShift, dw->as<uint8_t>());
        sync_needed = true;  // `widths` (pageable) and the staging ring are done with
    }
    if (sync_needed || oob) SRB_CUDA(cudaStreamSynchronize(s));
    if (oob) throw Error(SRB_ERR_INDEX_OOB, "minor index out of bounds");
    return link;
}
}  // namespace srb

using namespace srb;

extern "C" {

int32_t srb_ctx_set_upload_mode(srb_ctx *ctx, int32_t mode) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    SRB_REQUIRE(mode == SRB_UPLOAD_DEVICE_NARROW || mode == SRB_UPLOAD_HOST_PACK || mode == SRB_UPLOAD_AUTO ||
                    mode == SRB_UPLOAD_HOST_PACK_VALUES || mode == SRB_UPLOAD_HOST_PACK_ADAPTIVE || mode == SRB_UPLOAD_HOST_PACK_DELTA ||
                    mode == SRB_UPLOAD_BALANCED,
                SRB_ERR_INVALID_ARG, "bad upload mode");
    ctx->upload_mode = mode;
    SRB_API_END
}

int32_t srb_ctx_last_upload(srb_ctx *ctx, uint64_t *h2d_bytes, int32_t *host_packed) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    if (h2d_bytes) *h2d_bytes = ctx->last_upload_h2d;
    if (host_packed) *host_packed = ctx->last_upload_packed;
    SRB_API_END
}

int32_t srb_ctx_last_upload_chunks(srb_ctx *ctx, int32_t *chunks, int32_t *index_chunks_packed, int32_t *value_chunks_packed) {
    SRB_API_BEGIN
    SRB_REQUIRE(ctx, SRB_ERR_INVALID_ARG, "null ctx");
    if (chunks) *chunks = ctx->last_upload_chunks;
    if (index_chunks_packed) *index_chunks_packed = ctx->last_upload_idx_packed;
    if (value_chunks_packed) *value_chunks_packed = ctx->last_upload_val_packed;
    SRB_API_END
}

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
    const bool balanced = up_mode == SRB_UPLOAD_BALANCED;
    const bool packed = up_mode == SRB_UPLOAD_HOST_PACK || up_mode == SRB_UPLOAD_HOST_PACK_VALUES || up_mode == SRB_UPLOAD_HOST_PACK_ADAPTIVE ||
                        up_mode == SRB_UPLOAD_HOST_PACK_DELTA || balanced;
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
    if (balanced)
        link_bytes = upload_balanced(ctx, indices, idx_width, nnz, nminor, st->indices->as<uint32_t>(), flags->as<uint32_t>(),
                                     direct ? values : nullptr, f32_exact ? 4 : 8, m->values->p, offsets, nmajor, st->offsets->as<int64_t>());
    else if (packed)
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
        ctx->last_upload_packed = balanced ? (ctx->last_upload_idx_packed > 0 ? 1 : 0) : (packed ? 1 : 0);
        if (!balanced) ctx->last_upload_chunks = ctx->last_upload_idx_packed = ctx->last_upload_val_packed = 0;
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

}  // extern "C"
