/*
 * srb_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic of the SingleRust hot path, written from the
 * behaviour of the reference sources (paths relative to /root/reference):
 *   src/shared/statistics/helper/csr.rs   (number/sum/variance/min_max/std_dev, whole + chunk)
 *   src/shared/statistics/helper/csc.rs   (mirror for CSC)
 *   src/shared/statistics/mod.rs:17-41,59-83 (chunk drivers)
 *   src/memory/processing/scale/mod.rs    (total-count normalisation)
 *   src/memory/processing/transform/mod.rs (log1p)
 *   src/memory/processing/dim_red/mod.rs:123-156 (feature selection)
 *   src/shared/mod.rs:230-290             (selected densify)
 * The PCA arithmetic (external crate single_algebra 0.1.0-alpha.3, Cargo.toml:42) is restated in
 * oracle/pca_oracle.py (NumPy) from the in-tree dead-code spec src/shared/processing/pca/mod.rs:74-185.
 *
 * PARITY UNPINNED: the reference holds no golden vector / known-answer test for this path
 *  and cannot be compiled here (no Rust toolchain). The oracle is pinned instead
 * against hand-evaluated known answers (tests/golden/kat_4x5.json) and an
 * independent NumPy/SciPy statement (tests/test_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. Nothing under singlerust_b200/ may.
 *
 * Storage: a "compressed" matrix is (offsets u64[nmajor+1], indices u64[nnz], values T[nnz]);
 * CSR has major = row, CSC has major = column. `usize` of the Rust host is u64 here.
 * Direction: 0 = Row, 1 = Column (src/shared/mod.rs:39-42).
 *
 * Loop order and accumulation type (f64, serial, storage order) follow the reference so the
 * results are the bit-level stand-in for the uncompilable Rust; the *_omp variants are the
 * "best-effort parallel" CPU baseline (same arithmetic per line, thread-partitioned) and are
 * only used for timing.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * number (stored-entry counts). csr.rs:16-38 / csc.rs:15-35.
 * along_major != 0: offsets.windows(2) differences (CSR Row, CSC Column);
 * else histogram of the minor indices (CSR Column, CSC Row).
 * ---------------------------------------------------------------------------------------- */
void orc_number(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices, int along_major, u32 *out) {
    if (along_major) {
        for (u64 i = 0; i < nmajor; ++i) out[i] = (u32)(offsets[i + 1] - offsets[i]);
    } else {
        memset(out, 0, sizeof(u32) * nminor);
        u64 nnz = offsets[nmajor];
        for (u64 k = 0; k < nnz; ++k) out[indices[k]] += 1;
    }
}

/* number_chunk_helper, csr.rs:48-74 / csc.rs:45-68: accumulate into `reference` (length len).
 * faithful: the major direction adds the chunk's count at the CHUNK-LOCAL index i (the reference
 * discards the chunk offset, shared/statistics/mod.rs:24 — SURVEY §3.4/§10 defect).
 * major_offset is the corrected placement; pass 0 for the faithful behaviour. */
void orc_number_chunk(u64 nmajor, const u64 *offsets, const u64 *indices, int along_major, u64 major_offset,
                      u32 *reference, u64 len) {
    if (along_major) {
        for (u64 i = 0; i < nmajor; ++i) {
            u32 c = (u32)(offsets[i + 1] - offsets[i]);
            if (i + major_offset < len) reference[i + major_offset] += c;
        }
    } else {
        u64 nnz = offsets[nmajor] - offsets[0];
        const u64 *idx = indices;
        for (u64 k = 0; k < nnz; ++k)
            if (idx[k] < len) reference[idx[k]] += 1;
    }
}

#define DEFINE_FOR_TYPE(T, SUF)                                                                                    \
    /* sum_whole_helper, csr.rs:81-102 / csc.rs:74-95 */                                                          \
    void orc_sum_##SUF(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices, const T *values,          \
                       int along_major, double *out) {                                                            \
        if (along_major) {                                                                                        \
            for (u64 i = 0; i < nmajor; ++i) {                                                                    \
                double s = 0.0;                                                                                   \
                for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) s += (double)values[k];                         \
                out[i] = s;                                                                                       \
            }                                                                                                     \
        } else {                                                                                                  \
            for (u64 j = 0; j < nminor; ++j) out[j] = 0.0;                                                        \
            u64 nnz = offsets[nmajor];                                                                            \
            for (u64 k = 0; k < nnz; ++k) out[indices[k]] += (double)values[k];                                   \
        }                                                                                                         \
    }                                                                                                             \
    /* sum_chunk_helper, csr.rs:112-143 / csc.rs:105-131. Major direction OVERWRITES reference[i]               \
     * at the chunk-local index in the reference (defect); major_offset gives the corrected placement. */        \
    void orc_sum_chunk_##SUF(u64 nmajor, const u64 *offsets, const u64 *indices, const T *values,                \
                             int along_major, u64 major_offset, double *reference, u64 len) {                     \
        if (along_major) {                                                                                        \
            for (u64 i = 0; i < nmajor; ++i) {                                                                    \
                double s = 0.0;                                                                                   \
                for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) s += (double)values[k];                         \
                if (i + major_offset < len) reference[i + major_offset] = s;                                      \
            }                                                                                                     \
        } else {                                                                                                  \
            u64 nnz = offsets[nmajor];                                                                            \
            for (u64 k = 0; k < nnz; ++k)                                                                         \
                if (indices[k] < len) reference[indices[k]] += (double)values[k];                                 \
        }                                                                                                         \
    }                                                                                                             \
    /* variance_whole_helper, csr.rs:149-188 / csc.rs:137-176.                                                    \
     * major: nonzero-only two-pass population variance, empty line -> 0/0 = NaN.                                \
     * minor: one-pass  sq/cnt - mean^2 guarded by count>0 (empty -> 0.0). */                                     \
    void orc_variance_##SUF(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices, const T *values,     \
                            int along_major, double *out) {                                                       \
        if (along_major) {                                                                                        \
            for (u64 i = 0; i < nmajor; ++i) {                                                                    \
                double s = 0.0;                                                                                   \
                for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) s += (double)values[k];                         \
                double cnt = (double)(u32)(offsets[i + 1] - offsets[i]);                                          \
                double mean = s / cnt;                                                                            \
                double acc = 0.0;                                                                                 \
                for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) {                                               \
                    double d = (double)values[k] - mean;                                                          \
                    acc += d * d;                                                                                 \
                }                                                                                                 \
                out[i] = acc / cnt;                                                                               \
            }                                                                                                     \
        } else {                                                                                                  \
            double *sum = (double *)calloc(nminor ? nminor : 1, sizeof(double));                                  \
            double *sq = (double *)calloc(nminor ? nminor : 1, sizeof(double));                                   \
            u32 *cnt = (u32 *)calloc(nminor ? nminor : 1, sizeof(u32));                                           \
            u64 nnz = offsets[nmajor];                                                                            \
            for (u64 k = 0; k < nnz; ++k) sum[indices[k]] += (double)values[k];                                   \
            for (u64 k = 0; k < nnz; ++k) cnt[indices[k]] += 1;                                                   \
            for (u64 k = 0; k < nnz; ++k) {                                                                       \
                double v = (double)values[k];                                                                     \
                sq[indices[k]] += v * v;                                                                          \
            }                                                                                                     \
            for (u64 j = 0; j < nminor; ++j) {                                                                    \
                out[j] = 0.0;                                                                                     \
                if (cnt[j] > 0) {                                                                                 \
                    double mean = sum[j] / (double)cnt[j];                                                        \
                    out[j] = sq[j] / (double)cnt[j] - mean * mean;                                                \
                }                                                                                                 \
            }                                                                                                     \
            free(sum);                                                                                            \
            free(sq);                                                                                             \
            free(cnt);                                                                                            \
        }                                                                                                         \
    }                                                                                                             \
    /* std_dev_whole, csr.rs:225-228: sqrt of the variance, NaN propagates */                                     \
    void orc_std_dev_##SUF(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices, const T *values,      \
                           int along_major, double *out) {                                                        \
        orc_variance_##SUF(nmajor, nminor, offsets, indices, values, along_major, out);                           \
        u64 len = along_major ? nmajor : nminor;                                                                  \
        for (u64 i = 0; i < len; ++i) out[i] = sqrt(out[i]);                                                      \
    }                                                                                                             \
    /* min_max_whole_helper, csr.rs:194-223 / csc.rs:182-211; f64::min/max ignore a NaN operand (fmin/fmax) */    \
    void orc_min_max_##SUF(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices, const T *values,      \
                           int along_major, double *mn, double *mx) {                                             \
        u64 len = along_major ? nmajor : nminor;                                                                  \
        for (u64 i = 0; i < len; ++i) {                                                                           \
            mn[i] = INFINITY;                                                                                     \
            mx[i] = -INFINITY;                                                                                    \
        }                                                                                                         \
        for (u64 i = 0; i < nmajor; ++i)                                                                          \
            for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) {                                                   \
                u64 t = along_major ? i : indices[k];                                                             \
                double v = (double)values[k];                                                                     \
                mn[t] = fmin(mn[t], v);                                                                           \
                mx[t] = fmax(mx[t], v);                                                                           \
            }                                                                                                     \
    }                                                                                                             \
    /* scale_row / scale_col, scale/mod.rs:7-173: scale = 0 if sum==0 else target/sum; v *= scale[line];          \
     * result is always f64 (non-f64 input is converted first, scale/mod.rs:74-83). */                            \
    void orc_normalize_total_##SUF(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices,               \
                                   const T *values, int along_major, double target, double *out_values) {         \
        u64 len = along_major ? nmajor : nminor;                                                                  \
        double *scale = (double *)malloc(sizeof(double) * (len ? len : 1));                                       \
        orc_sum_##SUF(nmajor, nminor, offsets, indices, values, along_major, scale);                              \
        for (u64 i = 0; i < len; ++i) scale[i] = (scale[i] == 0.0) ? 0.0 : target / scale[i];                     \
        for (u64 i = 0; i < nmajor; ++i)                                                                          \
            for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) {                                                   \
                double v = (double)values[k];                                                                     \
                v *= scale[along_major ? i : indices[k]];                                                         \
                out_values[k] = v;                                                                                \
            }                                                                                                     \
        free(scale);                                                                                              \
    }                                                                                                             \
    /* convert_to_array_f64_{csr,csc}_selected, shared/mod.rs:230-290: dense[out_major? ...]. For CSR:            \
     * rows = major selection, cols = minor selection through a map (last duplicate wins, HashMap insert          \
     * order). Output is row-major (n_row_sel x n_col_sel) regardless of storage. */                              \
    void orc_densify_selected_##SUF(u64 nmajor, u64 nminor, const u64 *offsets, const u64 *indices,              \
                                    const T *values, int is_csc, const u64 *row_sel, u64 n_row_sel,               \
                                    const u64 *col_sel, u64 n_col_sel, double *dense) {                           \
        const u64 *major_sel = is_csc ? col_sel : row_sel;                                                        \
        u64 n_major_sel = is_csc ? n_col_sel : n_row_sel;                                                         \
        const u64 *minor_sel = is_csc ? row_sel : col_sel;                                                        \
        u64 n_minor_sel = is_csc ? n_row_sel : n_col_sel;                                                         \
        int64_t *map = (int64_t *)malloc(sizeof(int64_t) * (nminor ? nminor : 1));                                \
        for (u64 j = 0; j < nminor; ++j) map[j] = -1;                                                             \
        for (u64 j = 0; j < n_minor_sel; ++j) map[minor_sel[j]] = (int64_t)j;                                     \
        memset(dense, 0, sizeof(double) * n_row_sel * n_col_sel);                                                 \
        for (u64 o = 0; o < n_major_sel; ++o) {                                                                   \
            u64 mj = major_sel[o];                                                                                \
            if (mj >= nmajor) continue;                                                                           \
            for (u64 k = offsets[mj]; k < offsets[mj + 1]; ++k) {                                                 \
                int64_t p = map[indices[k]];                                                                      \
                if (p < 0) continue;                                                                              \
                if (is_csc)                                                                                       \
                    dense[(u64)p * n_col_sel + o] = (double)values[k];                                            \
                else                                                                                              \
                    dense[o * n_col_sel + (u64)p] = (double)values[k];                                            \
            }                                                                                                     \
        }                                                                                                         \
        free(map);                                                                                                \
    }

DEFINE_FOR_TYPE(float, f32)
DEFINE_FOR_TYPE(double, f64)

/* log1p_data, transform/mod.rs:8-62: f64 -> f64::ln_1p, f32 -> f32::ln_1p (in place, dtype kept). */
void orc_log1p_f64(u64 nnz, double *values) {
    for (u64 k = 0; k < nnz; ++k) values[k] = log1p(values[k]);
}
void orc_log1p_f32(u64 nnz, float *values) {
    for (u64 k = 0; k < nnz; ++k) values[k] = log1pf(values[k]);
}

/* select_features(HighlyVariable(n)), dim_red/mod.rs:135-140: stable sort of (index, variance) by
 * descending variance (partial_cmp; the reference panics on NaN — here: return -1), first n indices
 * in THAT order. Stable merge sort so that ties keep ascending index like Rust's sort_by. */
static void merge_sort_desc(u64 *idx, u64 *tmp, const double *v, u64 lo, u64 hi) {
    if (hi - lo < 2) return;
    u64 mid = lo + (hi - lo) / 2;
    merge_sort_desc(idx, tmp, v, lo, mid);
    merge_sort_desc(idx, tmp, v, mid, hi);
    u64 a = lo, b = mid, o = lo;
    while (a < mid && b < hi) {
        /* take from the right run only if strictly greater: keeps stability */
        if (v[idx[b]] > v[idx[a]])
            tmp[o++] = idx[b++];
        else
            tmp[o++] = idx[a++];
    }
    while (a < mid) tmp[o++] = idx[a++];
    while (b < hi) tmp[o++] = idx[b++];
    memcpy(idx + lo, tmp + lo, sizeof(u64) * (hi - lo));
}
int orc_select_hvg(const double *variances, u64 m, u64 n_top, u64 *out_idx) {
    for (u64 j = 0; j < m; ++j)
        if (isnan(variances[j])) return -1;
    u64 *idx = (u64 *)malloc(sizeof(u64) * (m ? m : 1));
    u64 *tmp = (u64 *)malloc(sizeof(u64) * (m ? m : 1));
    for (u64 j = 0; j < m; ++j) idx[j] = j;
    merge_sort_desc(idx, tmp, variances, 0, m);
    u64 take = n_top < m ? n_top : m;
    memcpy(out_idx, idx, sizeof(u64) * take);
    free(idx);
    free(tmp);
    return (int)0;
}
/* select_features(VarianceThreshold(t)), dim_red/mod.rs:148-153: v > t, ascending index order. */
u64 orc_select_var_threshold(const double *variances, u64 m, double t, u64 *out_idx) {
    u64 n = 0;
    for (u64 j = 0; j < m; ++j)
        if (variances[j] > t) out_idx[n++] = j;
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Best-effort parallel CPU baseline of the normalise -> log1p -> per-gene moments stage on a CSR
 * with the reference's host layout (u64 indices). Same per-line arithmetic; rows partitioned over
 * threads; per-thread gene partials merged in thread order. Timing use only.
 * Returns per-gene nonzero-only variance of the transformed values (csr.rs:172-186 semantics) and
 * writes the transformed f64 values.
 * ---------------------------------------------------------------------------------------- */
void orc_norm_log1p_genevar_omp_f32(u64 nrows, u64 ncols, const u64 *offsets, const u64 *indices,
                                    const float *values, double target, double *out_values, double *gene_sum,
                                    double *gene_sq, u32 *gene_cnt, double *gene_var) {
    int nt = orc_num_threads();
    double *psum = (double *)calloc((size_t)nt * ncols, sizeof(double));
    double *psq = (double *)calloc((size_t)nt * ncols, sizeof(double));
    u32 *pcnt = (u32 *)calloc((size_t)nt * ncols, sizeof(u32));
#pragma omp parallel
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        double *ls = psum + (size_t)t * ncols, *lq = psq + (size_t)t * ncols;
        u32 *lc = pcnt + (size_t)t * ncols;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < (int64_t)nrows; ++i) {
            double s = 0.0;
            for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) s += (double)values[k];
            double sc = (s == 0.0) ? 0.0 : target / s;
            for (u64 k = offsets[i]; k < offsets[i + 1]; ++k) {
                double v = log1p((double)values[k] * sc);
                out_values[k] = v;
                u64 c = indices[k];
                ls[c] += v;
                lq[c] += v * v;
                lc[c] += 1;
            }
        }
    }
    for (u64 j = 0; j < ncols; ++j) {
        double s = 0, q = 0;
        u32 c = 0;
        for (int t = 0; t < nt; ++t) {
            s += psum[(size_t)t * ncols + j];
            q += psq[(size_t)t * ncols + j];
            c += pcnt[(size_t)t * ncols + j];
        }
        gene_sum[j] = s;
        gene_sq[j] = q;
        gene_cnt[j] = c;
        gene_var[j] = 0.0;
        if (c > 0) {
            double mean = s / (double)c;
            gene_var[j] = q / (double)c - mean * mean;
        }
    }
    free(psum);
    free(psq);
    free(pcnt);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic count-matrix generator (CPU twin of the device generator in
 * singlerust_b200/csrc/synth.cu; integer-only so both sides are bit-identical). SURVEY §8(d).
 *   row key   r  = mix32(seed ^ mix32(row + 0x9E3779B9))
 *   presence  h  = mix32(r + col * 0x9E3779B1);   present iff h < ((thr[col] * depth(row)) >> 16)
 *   value     k  = min(ctz(mix32(h ^ 0x68E31DA4)), 15);  v = 1 + ((k * (16 + amp[col])) >> 4)
 *   depth(row)   = 65536 (no skew) or 32768 + (mix32(r ^ 0xA511E9B3) % 98304)  (0.5x .. 2x)
 * ---------------------------------------------------------------------------------------- */
static inline u32 mix32(u32 x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}
static inline u32 synth_row_key(u32 seed, u64 row) { return mix32(seed ^ mix32((u32)row + 0x9E3779B9U)); }
static inline u32 synth_depth(u32 r, int skew) { return skew ? 32768U + (mix32(r ^ 0xA511E9B3U) % 98304U) : 65536U; }
static inline int synth_entry(u32 r, u32 depth, u32 col, const u32 *thr, const u32 *amp, float *v) {
    u32 h = mix32(r + col * 0x9E3779B1U);
    u64 t = ((u64)thr[col] * depth) >> 16;
    if (t > 0xFFFFFFFFULL) t = 0xFFFFFFFFULL;
    if ((u64)h >= t) return 0;
    u32 g = mix32(h ^ 0x68E31DA4U);
    u32 k = g ? (u32)__builtin_ctz(g) : 32;
    if (k > 15) k = 15;
    *v = (float)(1U + ((k * (16U + amp[col])) >> 4));
    return 1;
}
/* pass 1: per-row counts for rows [row0, row0+nrows) of the global matrix */
void orc_synth_count(u32 seed, int skew, u64 row0, u64 nrows, u32 ncols, const u32 *thr, const u32 *amp,
                     u64 *offsets /* nrows+1 */) {
    offsets[0] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)nrows; ++i) {
        u32 r = synth_row_key(seed, row0 + (u64)i);
        u32 d = synth_depth(r, skew);
        u64 c = 0;
        float v;
        for (u32 j = 0; j < ncols; ++j) c += (u64)synth_entry(r, d, j, thr, amp, &v);
        offsets[i + 1] = c;
    }
    for (u64 i = 0; i < nrows; ++i) offsets[i + 1] += offsets[i];
}
/* pass 2: fill indices (u64) and values (f32) */
void orc_synth_fill(u32 seed, int skew, u64 row0, u64 nrows, u32 ncols, const u32 *thr, const u32 *amp,
                    const u64 *offsets, u64 *indices, float *values) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)nrows; ++i) {
        u32 r = synth_row_key(seed, row0 + (u64)i);
        u32 d = synth_depth(r, skew);
        u64 o = offsets[i];
        float v;
        for (u32 j = 0; j < ncols; ++j)
            if (synth_entry(r, d, j, thr, amp, &v)) {
                indices[o] = j;
                values[o] = v;
                ++o;
            }
    }
}
