// host_pack.cpp — host side of the PCIe upload: pack the reference's `usize` index arrays (nalgebra-sparse
// col_indices(), 8 bytes per stored entry; used at src/shared/statistics/helper/csr.rs:32,96) into the narrowest
// integer that holds `nminor` (2 bytes up to 65 536 genes, else 4) BEFORE they cross PCIe, on a small thread pool.
// At the bench size (1.5 G entries) the u64 indices are 12 of the 18 GB a step uploads; packed they are 3 GB.
// Pure host code (compiled by g++, no CUDA): marshalling only — no statistic is computed here.
#include <algorithm>
#include <atomic>
#include <sched.h>

#include <condition_variable>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <pthread.h>
#include <thread>
#include <vector>

#include "host_pack.h"

namespace srb {

static std::atomic<bool> g_forked_child{false};  // set in a forked child: the pool's worker threads do not exist there

// ---- a persistent pool: run(n, fn) calls fn(part) for part in [0, n), part 0 on the calling thread ------------
class HostPool {
    std::vector<std::thread> workers_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int)> *job_ = nullptr;
    uint64_t generation_ = 0;
    int parts_ = 0, next_ = 0, pending_ = 0;
    bool stop_ = false;

    void loop() {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [&] { return stop_ || (generation_ != seen && next_ < parts_); });
            if (stop_) return;
            seen = generation_;
            while (next_ < parts_) {
                const int p = next_++;
                const std::function<void(int)> *j = job_;
                lk.unlock();
                (*j)(p);
                lk.lock();
                if (--pending_ == 0) cv_done_.notify_all();
            }
        }
    }

public:
    explicit HostPool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size() + 1; }

    void run(int parts, const std::function<void(int)> &fn) {
        if (parts <= 0) return;
        if (parts == 1 || workers_.empty() || g_forked_child.load()) {
            for (int p = 0; p < parts; ++p) fn(p);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mu_);  // one parallel region at a time
        std::unique_lock<std::mutex> lk(mu_);
        job_ = &fn, parts_ = parts, next_ = 0, pending_ = parts;
        ++generation_;
        cv_work_.notify_all();
        while (next_ < parts_) {  // the caller works too
            const int p = next_++;
            lk.unlock();
            fn(p);
            lk.lock();
            --pending_;
        }
        cv_done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr, parts_ = 0;
    }
};

// CPUs this process may really use: the affinity mask and the cgroup-v2 CPU quota (hardware_concurrency sees neither;
// a pool larger than the quota only adds context switches).
static unsigned usable_cpus() {
    unsigned n = std::max(1u, std::thread::hardware_concurrency());
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min<unsigned>(n, (unsigned)std::max(1, CPU_COUNT(&set)));
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char quota[32] = {0};
        long period = 0;
        if (fscanf(f, "%31s %ld", quota, &period) == 2 && strcmp(quota, "max") != 0 && period > 0) {
            const long q = atol(quota);
            if (q > 0) n = std::min<unsigned>(n, (unsigned)std::max(1L, q / period));
        }
        fclose(f);
    }
    return n;
}

int host_pack_threads() {
    static int n = [] {
        const char *e = getenv("SRB_UPLOAD_THREADS");
        int v = e ? atoi(e) : 0;
        if (v <= 0) v = (int)std::min<unsigned>(usable_cpus(), 16u);
        return std::min(v, 64);
    }();
    return n;
}

// A forked child inherits the pool object but not its worker threads: it must not wait for them.
static void mark_forked_child() { g_forked_child.store(true); }

static HostPool &pool() {
    // leaked on purpose: no static destructor has to join worker threads at process exit (or in a forked child)
    static HostPool *p = [] {
        pthread_atfork(nullptr, nullptr, mark_forked_child);
        return new HostPool(host_pack_threads() - 1);
    }();
    return *p;
}

// Parts per parallel region: 4 per thread (handed out dynamically by the pool), never smaller than `min_units`. With one
// part per thread a region lasts as long as its slowest thread, and in the pipelined end-to-end step one of the cores is
// shared with the host thread that drives the other batch's kernels: that thread's part took twice as long.
static int region_parts(int nthreads, uint64_t n, int min_shift) {
    static const int per_thread = [] {
        const char *e = getenv("SRB_PACK_PARTS_PER_THREAD");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 16 ? v : 4;
    }();
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)nthreads * (nthreads > 1 ? per_thread : 1), n >> min_shift));
}

// ---- the packing loops; cloned per ISA so the .so stays loadable on any x86-64 host -----------------------------
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define SRB_ISA_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define SRB_ISA_CLONES
#endif

// Bounds test without a compare (so the loop vectorises on AVX2, which has no unsigned 64-bit max): for bound <= 2^63,
// the top bit of  v | (bound - 1 - v)  is set iff v >= bound. The OR over the block is returned.
#define SRB_PACK_LOOP(NAME, SRC, DST)                                                                              \
    SRB_ISA_CLONES static uint64_t NAME(const SRC *__restrict__ s, DST *__restrict__ d, uint64_t n, uint64_t bm1) { \
        uint64_t acc = 0;                                                                                          \
        for (uint64_t i = 0; i < n; ++i) {                                                                         \
            const uint64_t v = (uint64_t)s[i];                                                                     \
            acc |= v | (bm1 - v);                                                                                  \
            d[i] = (DST)v;                                                                                         \
        }                                                                                                          \
        return acc;                                                                                                \
    }
SRB_PACK_LOOP(pack_u64_u16, uint64_t, uint16_t)
SRB_PACK_LOOP(pack_u64_u32, uint64_t, uint32_t)
SRB_PACK_LOOP(pack_u32_u16, uint32_t, uint16_t)
SRB_PACK_LOOP(pack_u32_u32, uint32_t, uint32_t)

static uint64_t pack_range(const void *src, int sw, void *dst, int dw, uint64_t a, uint64_t b, uint64_t bm1) {
    const uint64_t n = b - a;
    if (sw == 8 && dw == 2) return pack_u64_u16((const uint64_t *)src + a, (uint16_t *)dst + a, n, bm1);
    if (sw == 8 && dw == 4) return pack_u64_u32((const uint64_t *)src + a, (uint32_t *)dst + a, n, bm1);
    if (sw == 4 && dw == 2) return pack_u32_u16((const uint32_t *)src + a, (uint16_t *)dst + a, n, bm1);
    return pack_u32_u32((const uint32_t *)src + a, (uint32_t *)dst + a, n, bm1);
}

bool host_pack_indices(const void *src, int src_width, uint64_t n, void *dst, int dst_width, uint64_t bound, int nthreads) {
    if (n == 0) return false;
    if (bound == 0) return true;
    if (bound > (1ull << 63)) bound = 1ull << 63;
    const uint64_t bm1 = bound - 1;
    if (nthreads <= 0) nthreads = host_pack_threads();
    // >= 64 K entries per part: below that the fork/join costs more than the copy
    const int parts = region_parts(nthreads, n, 16);
    if (parts == 1) return (pack_range(src, src_width, dst, dst_width, 0, n, bm1) >> 63) != 0;
    std::vector<uint64_t> acc((size_t)parts, 0);
    const uint64_t per = ((n + parts - 1) / parts + 63) & ~uint64_t(63);  // parts start on 64-entry boundaries
    pool().run(parts, [&](int p) {
        const uint64_t a = std::min<uint64_t>(n, per * (uint64_t)p), b = std::min<uint64_t>(n, a + per);
        if (b > a) acc[(size_t)p] = pack_range(src, src_width, dst, dst_width, a, b, bm1);
    });
    uint64_t all = 0;
    for (uint64_t v : acc) all |= v;
    return (all >> 63) != 0;
}

void host_copy_parallel(const void *src, void *dst, uint64_t bytes, int nthreads) {
    if (bytes == 0) return;
    if (nthreads <= 0) nthreads = host_pack_threads();
    const int parts = region_parts(nthreads, bytes, 19);
    if (parts == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    const uint64_t per = ((bytes + parts - 1) / parts + 4095) & ~uint64_t(4095);
    pool().run(parts, [&](int p) {
        const uint64_t a = std::min<uint64_t>(bytes, per * (uint64_t)p), b = std::min<uint64_t>(bytes, a + per);
        if (b > a) memcpy((char *)dst + a, (const char *)src + a, b - a);
    });
}

void upload_mix(double t_idx, double t_val, uint64_t len, int pw, int width, size_t vsz, int vstate, double rate, double *g_out, double *f_out) {
    double g = 0.0, f = 0.0;
    const double t_ki = (double)len * pw / rate, t_ri = (double)len * width / rate;
    const double t_vr = vsz ? (double)len * (double)vsz / rate : 0.0, t_vp = vsz ? (double)len * (vstate == 2 ? 2 : 1) / rate : 0.0;
    if (t_idx > 0.0) {
        // host and link finish together when (1 - g) t_idx = (1 - g) t_k + g t_r
        const double t_k = t_ki + t_vr, t_r = t_ri + t_vr;
        if (t_idx > t_k) g = (t_idx - t_k) / (t_idx - t_k + t_r);
        // spare host time: t_idx + f t_val = t_k - f (t_vr - t_vp)
        else if (vstate && vsz == 4 && t_val > 0.0) f = std::min(1.0, (t_k - t_idx) / (t_val + t_vr - t_vp));
    }
    *g_out = g, *f_out = f;
}

// ---- count values: f32 -> u8 / u16 when that is lossless ------------------------------------------------------------
// Raw counts are small non-negative integers stored as f32 (the common h5ad case). A chunk whose every value is an
// integer in [0, 2^(8 dw)) travels as dw-byte integers. The test is bit-exact — the round trip through int32 must
// reproduce the f32 bit pattern, so -0.0, NaN, fractions and negatives all refuse — and written on the bit patterns
// without float compares so it vectorises: for a non-negative finite f32 the pattern is monotone in the value, hence
// "negative, NaN, inf or >= 65536" is one unsigned comparison against 0x477FFFFF (65535.996); such lanes are zeroed
// before the (then always defined) float -> int conversion and flagged.
#define SRB_PACK_VALUES_LOOP(NAME, DST)                                                                             \
    SRB_ISA_CLONES static uint32_t NAME(const float *__restrict__ s, DST *__restrict__ d, uint64_t n) {             \
        uint32_t bad = 0, range = 0;                                                                                \
        for (uint64_t i = 0; i < n; ++i) {                                                                          \
            uint32_t bits, bbits;                                                                                   \
            memcpy(&bits, s + i, 4);                                                                                \
            const uint32_t m = (uint32_t)((int32_t)(0x477FFFFFu - bits) >> 31);                                     \
            const uint32_t sb = bits & ~m;                                                                          \
            float c;                                                                                                \
            memcpy(&c, &sb, 4);                                                                                     \
            const int32_t q = (int32_t)c;                                                                           \
            const float back = (float)q;                                                                            \
            memcpy(&bbits, &back, 4);                                                                               \
            bad |= (bits ^ bbits) | m;                                                                              \
            range |= (uint32_t)q;                                                                                   \
            d[i] = (DST)q;                                                                                          \
        }                                                                                                           \
        return bad | (range >> (8 * sizeof(DST)));                                                                  \
    }
SRB_PACK_VALUES_LOOP(pack_f32_u8, uint8_t)
SRB_PACK_VALUES_LOOP(pack_f32_u16, uint16_t)

bool host_pack_values_f32(const float *src, uint64_t n, void *dst, int dst_width, int nthreads) {
    if (n == 0) return true;
    if (nthreads <= 0) nthreads = host_pack_threads();
    const int parts = region_parts(nthreads, n, 16);
    auto one = [&](uint64_t a, uint64_t b) -> uint32_t {
        return dst_width == 1 ? pack_f32_u8(src + a, (uint8_t *)dst + a, b - a) : pack_f32_u16(src + a, (uint16_t *)dst + a, b - a);
    };
    if (parts == 1) return one(0, n) == 0;
    std::vector<uint32_t> bad((size_t)parts, 0);
    const uint64_t per = ((n + parts - 1) / parts + 63) & ~uint64_t(63);
    pool().run(parts, [&](int p) {
        const uint64_t a = std::min<uint64_t>(n, per * (uint64_t)p), b = std::min<uint64_t>(n, a + per);
        if (b > a) bad[(size_t)p] = one(a, b);
    });
    uint32_t all = 0;
    for (uint32_t v : bad) all |= v;
    return all == 0;
}

// ---- delta coding of the sorted minor indices ------------------------------------------------------------------------
// Within a line the indices are strictly increasing (nalgebra-sparse invariant), so at 5 % density the gap to the
// previous index averages 20: one byte per entry instead of two. code = index - previous (previous = 0 at a line start)
// when that is < 255, else the escape 255 with the full index in a side list (position, value), in entry order. A
// duplicate (gap 0) is representable; an unsorted pair wraps to a huge gap and becomes an escape, so the device rebuilds
// EXACTLY the caller's array and its canonical-form check still sees what the caller passed.
#define SRB_DELTA_LOOP(NAME, SRC)                                                                                        \
    SRB_ISA_CLONES static uint64_t NAME(const SRC *__restrict__ c, uint8_t *__restrict__ d, uint64_t n, uint64_t bm1,    \
                                        uint64_t *n_escapes) {                                                          \
        uint64_t acc = 0, esc = 0;                                                                                      \
        for (uint64_t j = 0; j < n; ++j) { /* c[j - 1] exists: the run starts at the second entry of its line part */    \
            const uint64_t cur = (uint64_t)c[j], delta = cur - (uint64_t)c[(int64_t)j - 1];                              \
            const uint64_t hi = delta >> 8, big = 0 - ((hi | (0 - hi)) >> 63); /* all ones iff delta >= 256 (or wrapped) */ \
            const uint32_t code = (uint32_t)((delta | big) & 255);                                                      \
            acc |= cur | (bm1 - cur);                                                                                   \
            esc += (code + 1) >> 8;                                                                                     \
            d[j] = (uint8_t)code;                                                                                       \
        }                                                                                                               \
        *n_escapes = esc;                                                                                               \
        return acc;                                                                                                     \
    }
SRB_DELTA_LOOP(delta_run_u64, uint64_t)
SRB_DELTA_LOOP(delta_run_u32, uint32_t)

template <class IDX, class OFF>
static uint64_t delta_part(const IDX *cols, const OFF *offs, uint64_t nmajor, uint64_t a, uint64_t b, uint8_t *dst, uint64_t o,
                           uint64_t bm1, DeltaEscapes &esc) {
    if (a >= b) return 0;
    // line containing entry a: the last r with offs[r] <= a (empty lines before it are skipped by the search)
    uint64_t r = (uint64_t)(std::upper_bound(offs, offs + nmajor + 1, (OFF)a) - offs) - 1;
    uint64_t acc = 0, i = a;
    while (i < b) {
        while (r + 1 <= nmajor && (uint64_t)offs[r + 1] <= i) ++r;  // empty lines
        const uint64_t line_end = std::min<uint64_t>((uint64_t)offs[r + 1], b);
        // first entry of this line part
        const uint64_t cur = (uint64_t)cols[i], prev = (i == (uint64_t)offs[r]) ? 0 : (uint64_t)cols[i - 1];
        const uint64_t delta = cur - prev;
        acc |= cur | (bm1 - cur);
        if (delta < 255) {
            dst[i - o] = (uint8_t)delta;
        } else {
            dst[i - o] = 255;
            esc.pos.push_back(i), esc.val.push_back((uint32_t)cur);
        }
        // the rest of the line part, vectorised
        const uint64_t n = line_end - (i + 1);
        if (n) {
            uint64_t ne = 0;
            if (sizeof(IDX) == 8) acc |= delta_run_u64((const uint64_t *)(const void *)(cols + i + 1), dst + (i + 1 - o), n, bm1, &ne);
            else acc |= delta_run_u32((const uint32_t *)(const void *)(cols + i + 1), dst + (i + 1 - o), n, bm1, &ne);
            if (ne)
                for (uint64_t j = i + 1; j < line_end; ++j)
                    if (dst[j - o] == 255) esc.pos.push_back(j), esc.val.push_back((uint32_t)cols[j]);
        }
        i = line_end;
    }
    return acc;
}

bool host_offsets_valid(const void *offs, int width, uint64_t nmajor, uint64_t nnz) {
    uint64_t prev = 0;
    for (uint64_t r = 0; r <= nmajor; ++r) {
        const uint64_t v = width == 8 ? ((const uint64_t *)offs)[r] : (uint64_t)((const uint32_t *)offs)[r];
        if (v < prev || v > nnz || (r == 0 && v != 0)) return false;
        prev = v;
    }
    return prev == nnz;
}

bool host_delta_encode(const void *cols, const void *offs, int width, uint64_t nmajor, uint64_t o, uint64_t len, uint8_t *dst,
                       uint64_t bound, int nthreads, DeltaEscapes &esc) {
    if (len == 0) return false;
    if (bound == 0) return true;
    if (bound > (1ull << 63)) bound = 1ull << 63;
    const uint64_t bm1 = bound - 1;
    if (nthreads <= 0) nthreads = host_pack_threads();
    const int parts = region_parts(nthreads, len, 16);
    std::vector<DeltaEscapes> pe((size_t)parts);
    std::vector<uint64_t> acc((size_t)parts, 0);
    const uint64_t per = ((len + parts - 1) / parts + 63) & ~uint64_t(63);
    auto one = [&](int p) {
        const uint64_t a = o + std::min<uint64_t>(len, per * (uint64_t)p), b = std::min<uint64_t>(o + len, a + per);
        if (width == 8)
            acc[(size_t)p] = delta_part((const uint64_t *)cols, (const uint64_t *)offs, nmajor, a, b, dst, o, bm1, pe[(size_t)p]);
        else
            acc[(size_t)p] = delta_part((const uint32_t *)cols, (const uint32_t *)offs, nmajor, a, b, dst, o, bm1, pe[(size_t)p]);
    };
    if (parts == 1) one(0);
    else pool().run(parts, one);
    uint64_t all = 0;
    for (int p = 0; p < parts; ++p) {  // parts are in entry order, so the merged list stays sorted by position
        all |= acc[(size_t)p];
        esc.pos.insert(esc.pos.end(), pe[(size_t)p].pos.begin(), pe[(size_t)p].pos.end());
        esc.val.insert(esc.val.end(), pe[(size_t)p].val.begin(), pe[(size_t)p].val.end());
    }
    return (all >> 63) != 0;
}

}  // namespace srb

// test hook: delta-encode a whole index array in chunks of `chunk` entries (the upload's chunking); the escape arrays
// hold up to esc_cap entries (more escapes than that: returns -2 with *n_esc = the number needed)
extern "C" int32_t srb_host_delta_encode(const void *cols, const void *offs, int32_t width, uint64_t nmajor, uint64_t nnz, uint64_t bound,
                                         uint64_t chunk, int32_t nthreads, uint8_t *codes, uint64_t *esc_pos, uint32_t *esc_val,
                                         uint64_t esc_cap, uint64_t *n_esc, int32_t *out_of_bounds) {
    if ((width != 4 && width != 8) || !offs || (nnz && (!cols || !codes)) || chunk == 0) return -1;
    if (!srb::host_offsets_valid(offs, width, nmajor, nnz)) return -3;
    srb::DeltaEscapes esc;
    bool oob = false;
    for (uint64_t o = 0; o < nnz; o += chunk)
        oob |= srb::host_delta_encode(cols, offs, width, nmajor, o, std::min<uint64_t>(chunk, nnz - o), codes + o, bound, nthreads, esc);
    if (n_esc) *n_esc = esc.pos.size();
    if (out_of_bounds) *out_of_bounds = oob ? 1 : 0;
    if (esc.pos.size() > esc_cap) return -2;
    if (!esc.pos.empty()) {
        memcpy(esc_pos, esc.pos.data(), 8 * esc.pos.size());
        memcpy(esc_val, esc.val.data(), 4 * esc.val.size());
    }
    return 0;
}

extern "C" int32_t srb_upload_mix(double t_idx_ms, double t_val_ms, uint64_t len, int32_t packed_index_bytes, int32_t idx_width,
                                  int32_t value_bytes, double link_gbs, double *raw_index_fraction, double *packed_value_fraction) {
    if (!raw_index_fraction || !packed_value_fraction || !(link_gbs > 0.0) || len == 0) return -1;
    srb::upload_mix(t_idx_ms, t_val_ms, len, packed_index_bytes, idx_width, (size_t)value_bytes, value_bytes == 4 ? 1 : 0, link_gbs * 1e6,
                    raw_index_fraction, packed_value_fraction);
    return 0;
}

extern "C" int32_t srb_host_pack_values_f32(const float *src, uint64_t n, void *dst, int32_t dst_width, int32_t nthreads,
                                            int32_t *lossless) {
    if ((n && (!src || !dst)) || (dst_width != 1 && dst_width != 2)) return -1;
    const bool ok = srb::host_pack_values_f32(src, n, dst, dst_width, nthreads);
    if (lossless) *lossless = ok ? 1 : 0;
    return 0;
}

extern "C" int32_t srb_host_pack_indices(const void *src, int32_t src_width, uint64_t n, void *dst, int32_t dst_width,
                                         uint64_t bound, int32_t nthreads, int32_t *out_of_bounds) {
    if ((n && (!src || !dst)) || (src_width != 4 && src_width != 8) || (dst_width != 2 && dst_width != 4)) return -1;
    const bool oob = srb::host_pack_indices(src, src_width, n, dst, dst_width, bound, nthreads);
    if (out_of_bounds) *out_of_bounds = oob ? 1 : 0;
    return 0;
}
