// host_pack.h — host-side packing helpers of the upload path (host_pack.cpp; plain C++, no CUDA).
#pragma once
#include <cstdint>
#include <vector>

namespace srb {
// threads the upload path uses (SRB_UPLOAD_THREADS, default min(hardware threads, 16))
int host_pack_threads();
// dst[i] = (narrow) src[i] for i < n, src entries `src_width` (4|8) bytes, dst entries `dst_width` (2|4) bytes;
// returns true when any source value is >= bound. nthreads <= 0: the default.
bool host_pack_indices(const void *src, int src_width, uint64_t n, void *dst, int dst_width, uint64_t bound, int nthreads);
// dst[i] = (u8 | u16) src[i] when EVERY src[i] is an integer in [0, 2^(8 dst_width)) with the exact f32 bit pattern of
// that integer (so the device can rebuild the f32 array bit for bit); returns false otherwise (dst is then garbage)
bool host_pack_values_f32(const float *src, uint64_t n, void *dst, int dst_width, int nthreads);
// delta coding of the sorted minor indices (one byte per entry + an escape list); see host_pack.cpp
struct DeltaEscapes {
    std::vector<uint64_t> pos;  // global entry positions, ascending
    std::vector<uint32_t> val;  // the full index at that position
};
// offsets[0] == 0, monotone, offsets[nmajor] == nnz (the delta coder walks the lines and must not be led astray)
bool host_offsets_valid(const void *offs, int width, uint64_t nmajor, uint64_t nnz);
// codes for the entries [o, o + len) into dst[0 .. len), escapes appended to esc; returns true when any index >= bound
bool host_delta_encode(const void *cols, const void *offs, int width, uint64_t nmajor, uint64_t o, uint64_t len, uint8_t *dst,
                       uint64_t bound, int nthreads, DeltaEscapes &esc);
// The rate model of the BALANCED upload (upload.cu). Per chunk of `len` entries: packing the indices costs the host t_idx_ms
// (measured) and the link len * pw bytes; raw costs the host nothing and the link len * width bytes; values travel raw
// (vsz bytes) or packed (1 byte, 2 when vstate == 2) at a host cost of t_val_ms. *raw_index_fraction = the share of chunks
// whose indices should go raw so that host and link finish together (0 when the host packs faster than the link drains);
// *packed_value_fraction = the share whose values should be packed with the host time that is left over (0 otherwise).
void upload_mix(double t_idx_ms, double t_val_ms, uint64_t len, int pw, int width, size_t vsz, int vstate, double link_bytes_per_ms,
                double *raw_index_fraction, double *packed_value_fraction);
// threaded memcpy (pageable host memory -> pinned staging ring)
void host_copy_parallel(const void *src, void *dst, uint64_t bytes, int nthreads);
}  // namespace srb
