/*
 * srb200.h — C ABI of libsrb200.so: the B200 (sm_100a) drop-in for SingleRust's sparse count-matrix
 * analytics path. Plain pointers and sizes only; every function returns an srb_status (0 = OK) and never
 * aborts or throws across the boundary; srb_last_error_message() gives the thread-local reason.
 *
 * The reference (pure Rust, /root/reference) has no FFI boundary; the seam sits where it hands raw
 * CSR/CSC slices (nalgebra-sparse row_offsets()/col_indices()/values(), `usize` = u64) to arithmetic.
 * Each entry point cites the reference function whose body it replaces. INTEGRATION.md shows the
 * Rust `extern "C"` block that binds these symbols 1:1.
 *
 * Model: upload once -> many ops on the device-resident matrix -> download. One srb_ctx = one GPU + one
 * CUDA stream (+ optionally one NCCL rank of a cell-row-sharded job). Calls on one ctx are serialised
 * by the caller (mirrors the RwLock in anndata-memory's IMArrayElement, src/memory/statistics/mod.rs:12).
 */
#ifndef SRB200_H
#define SRB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRB_VERSION_MAJOR 0
#define SRB_VERSION_MINOR 1

typedef enum srb_status {
    SRB_OK = 0,
    SRB_ERR_INVALID_ARG = -1,
    SRB_ERR_UNSUPPORTED_DTYPE = -2, /* I64/U64/Usize/Bool/String: the reference panics, shared/mod.rs:117-126 */
    SRB_ERR_INDEX_OOB = -3,
    SRB_ERR_CUDA = -4,
    SRB_ERR_NCCL = -5,
    SRB_ERR_OOM = -6,
    SRB_ERR_NAN = -7,        /* NaN variance in HVG sort: the reference panics, dim_red/mod.rs:138 */
    SRB_ERR_UNSUPPORTED = -8 /* container/op combination the reference answers with todo!() */
} srb_status;

/* src/shared/mod.rs:39-42 — identical discriminants */
typedef enum srb_direction { SRB_ROW = 0, SRB_COLUMN = 1 } srb_direction;

/* anndata DynCsrMatrix / DynCscMatrix variants accepted by match_dyn_cs{r,c}_matrix!, shared/mod.rs:110-150 */
typedef enum srb_dtype {
    SRB_I8 = 0, SRB_I16 = 1, SRB_I32 = 2, SRB_I64 = 3, /* I64: unsupported */
    SRB_U8 = 4, SRB_U16 = 5, SRB_U32 = 6, SRB_U64 = 7, /* U64: unsupported */
    SRB_F32 = 8, SRB_F64 = 9
} srb_dtype;

typedef enum srb_format { SRB_CSR = 0, SRB_CSC = 1 } srb_format;

/* width of the host offset/index integers: 8 = Rust usize (nalgebra-sparse), 4 = h5ad on-disk int32 */
typedef enum srb_index_width { SRB_IDX32 = 4, SRB_IDX64 = 8 } srb_index_width;

/* value storage policy after normalize_total / log1p:
 *   COMPACT  keep f32 storage when the input was f32-representable (halves HBM traffic; |rel err| <= 2e-7)
 *   FAITHFUL promote to f64 exactly like the reference does (scale/mod.rs:74-83, transform/mod.rs:48-55) */
typedef enum srb_value_mode { SRB_VALUES_COMPACT = 0, SRB_VALUES_FAITHFUL = 1 } srb_value_mode;

/* how srb_mat_upload / srb_stream_push move the index array over PCIe (results are identical):
 *   DEVICE_NARROW  copy the host integers as they are (8 bytes per entry for Rust usize) and narrow on the device
 *   HOST_PACK      narrow on the host (thread pool, pinned staging ring) to 2 bytes per entry when the minor dimension
 *                  is <= 65 536, else 4, overlap packing with the DMA, and stage pageable caller memory (a Rust Vec)
 *                  through the same ring
 *   HOST_PACK_VALUES  HOST_PACK, and f32 values travel as u8 / u16 for every chunk in which that is lossless (raw
 *                  counts; bit-identical f32 on the device). Pays only when the link, not the host, is the bottleneck:
 *                  on the 16-core bench host it is slower than HOST_PACK (measured, profiles/), so AUTO never picks it
 *   HOST_PACK_ADAPTIVE  HOST_PACK, and a chunk's values are packed only while the host is ahead of the link (the staging
 *                  slot's previous DMA is still in flight), so the two stay balanced (upload at the bench size: 147-168 ms).
 *   HOST_PACK_DELTA  HOST_PACK with the sorted indices of each line delta-coded to ONE byte per entry (gap to the previous
 *                  index; gaps >= 255, line starts beyond 254 and non-canonical pairs escape to a side list), rebuilt
 *                  exactly on the device (upload at the bench size: 147-155 ms against 171-180 for HOST_PACK, 330 raw).
 *   BALANCED       every 4 M-entry chunk is either packed on the host (indices: DELTA codes, or HOST_PACK narrowing when the
 *                  offsets cannot be trusted; f32 counts: u8 / u16 where lossless) or sent raw and narrowed on the device —
 *                  packed exactly when the copies already queued on the link take at least as long as packing the chunk
 *                  (measured, running estimate), so host and link stay busy for any number of host threads per rank
 *   AUTO           BALANCED for arrays of >= 2^20 entries, else DEVICE_NARROW
 * Process default: AUTO; environment SRB_UPLOAD_PACK (0 | 1 | auto | values | adaptive | delta | balanced) overrides it;
 * SRB_LINK_GBS (default 50) is the link rate BALANCED assumes when it converts queued bytes into time. */
typedef enum srb_upload_mode {
    SRB_UPLOAD_DEVICE_NARROW = 0, SRB_UPLOAD_HOST_PACK = 1, SRB_UPLOAD_AUTO = 2, SRB_UPLOAD_HOST_PACK_VALUES = 3,
    SRB_UPLOAD_HOST_PACK_ADAPTIVE = 4, SRB_UPLOAD_HOST_PACK_DELTA = 5, SRB_UPLOAD_BALANCED = 6
} srb_upload_mode;

/* K8, the eigensolver behind srb_pca (top-k eigenpairs of the n_sel x n_sel correlation matrix):
 *   SYEVD  cuSOLVER's full fp64 divide-and-conquer solver (28 ms at n_sel = 2000: latency-bound tridiagonalisation)
 *   CHFSI  Chebyshev-filtered subspace iteration for the k leading pairs, built from fp64 GEMMs (csrc/eig.cu); used when
 *          n_sel >= 1024 and 8 k <= n_sel, converged to ||C v - theta v|| <= 1e-11 |theta_1|, and falls back to SYEVD on
 *          any doubt (breakdown, non-finite values, no convergence)
 * Process default: CHFSI; environment SRB_EIG_MODE (syevd | chfsi) overrides it. */
typedef enum srb_eig_mode { SRB_EIG_SYEVD = 0, SRB_EIG_CHFSI = 1 } srb_eig_mode;

typedef struct srb_ctx srb_ctx;
typedef struct srb_mat srb_mat;

/* ---- library ------------------------------------------------------------------------------------ */
const char *srb_version(void);
const char *srb_last_error_message(void);
/* number of CUDA kernels launched by this library on the calling process since load (bench "gpu_launches") */
uint64_t srb_kernel_launch_count(void);

/* host helper of the HOST_PACK upload (no GPU involved; exposed so the packing can be tested and timed alone):
 * dst[i] = src[i] narrowed from src_width (4 | 8) to dst_width (2 | 4) bytes on nthreads (<= 0: default) host threads;
 * *out_of_bounds = 1 when any source value is >= bound. Returns 0, or -1 on a bad argument. */
int32_t srb_host_pack_indices(const void *src, int32_t src_width, uint64_t n, void *dst, int32_t dst_width,
                              uint64_t bound, int32_t nthreads, int32_t *out_of_bounds);

/* the value twin: f32 counts -> dst_width (1 | 2) byte integers; *lossless = 1 iff every value is an integer in
 * [0, 2^(8 dst_width)) whose f32 bit pattern the device can rebuild exactly (else dst is unspecified) */
int32_t srb_host_pack_values_f32(const float *src, uint64_t n, void *dst, int32_t dst_width, int32_t nthreads,
                                 int32_t *lossless);

/* the delta coder of the HOST_PACK_DELTA upload, for testing (no GPU): codes[i] = indices[i] - previous index of the line
 * (0 at a line start) when < 255, else 255 with (position, full index) appended to esc_pos / esc_val (capacity esc_cap;
 * *n_esc = number produced). `chunk` = entries per encoding call, as the upload chunks them. Returns 0, -1 bad argument,
 * -2 escape capacity too small, -3 offsets not monotone / not ending at nnz. */
int32_t srb_host_delta_encode(const void *indices, const void *offsets, int32_t idx_width, uint64_t nmajor, uint64_t nnz,
                              uint64_t bound, uint64_t chunk, int32_t nthreads, uint8_t *codes, uint64_t *esc_pos,
                              uint32_t *esc_val, uint64_t esc_cap, uint64_t *n_esc, int32_t *out_of_bounds);

/* the rate model of the BALANCED upload, for testing (no GPU): with a measured packing time of t_idx_ms / t_val_ms per chunk of
 * `len` entries and a link of link_gbs GB/s, the share of chunks whose indices should travel raw (idx_width bytes per entry
 * instead of packed_index_bytes) and the share whose f32 values should be packed to one byte, so that host and link finish
 * together. Returns 0, or -1 on a bad argument. */
int32_t srb_upload_mix(double t_idx_ms, double t_val_ms, uint64_t len, int32_t packed_index_bytes, int32_t idx_width,
                       int32_t value_bytes, double link_gbs, double *raw_index_fraction, double *packed_value_fraction);

/* ---- context ------------------------------------------------------------------------------------ */
int32_t srb_ctx_create(int32_t device, srb_ctx **out);
int32_t srb_ctx_destroy(srb_ctx *ctx);
int32_t srb_ctx_set_value_mode(srb_ctx *ctx, int32_t mode /* srb_value_mode */);
int32_t srb_ctx_set_upload_mode(srb_ctx *ctx, int32_t mode /* srb_upload_mode */);
int32_t srb_ctx_set_eig_mode(srb_ctx *ctx, int32_t mode /* srb_eig_mode */);
/* what K8 did in the last srb_pca / srb_pipeline_* call on this ctx: solver = 0 SYEVD, 1 CHFSI, 2 CHFSI fell back to SYEVD;
 * for CHFSI the number of d x b block products (DGEMMs), outer (Rayleigh-Ritz) iterations and the final
 * max ||C v - theta v|| / |theta_1| over the k pairs. Any pointer may be NULL. */
int32_t srb_ctx_last_eig(srb_ctx *ctx, int32_t *solver, int32_t *block_products, int32_t *outer_iterations, double *max_residual);
/* what the last srb_mat_upload on this ctx moved over the link: bytes, and whether the index array was host-packed */
int32_t srb_ctx_last_upload(srb_ctx *ctx, uint64_t *h2d_bytes, int32_t *host_packed);
/* BALANCED upload only: chunks of the last upload, and how many of them had their indices / values packed on the host */
int32_t srb_ctx_last_upload_chunks(srb_ctx *ctx, int32_t *chunks, int32_t *index_chunks_packed, int32_t *value_chunks_packed);
int32_t srb_ctx_synchronize(srb_ctx *ctx);
/* the cudaStream_t all work of this ctx is enqueued on (for CUDA-event timing by the caller) */
void *srb_ctx_stream(srb_ctx *ctx);

/* multi-GPU (new; the reference is single-process): one rank per GPU, matrix sharded by cell-row.
 * id128 is an ncclUniqueId (128 bytes) produced by srb_comm_unique_id on rank 0 and broadcast by the
 * launcher (bench.py uses torch.distributed for that rendezvous only). */
int32_t srb_comm_unique_id(void *id128);
int32_t srb_ctx_comm_init(srb_ctx *ctx, const void *id128, int32_t rank, int32_t nranks);

/* ---- matrix upload / download ------------------------------------------------------------------- */
/* Borrow host arrays for the duration of the call (Rust holds the read guard) and build the device copy.
 * offsets[nmajor+1], indices[nnz] of width idx_width; values[nnz] of `dtype`.
 * Indices must be < nminor (else SRB_ERR_INDEX_OOB, checked on device). Replaces nothing in the reference;
 * it is the residency step in front of every function below.
 * global_row0 / global_nrows describe this rank's shard of a row-sharded CSR (pass 0 / nrows when unsharded). */
int32_t srb_mat_upload(srb_ctx *ctx, int32_t format, uint64_t nrows, uint64_t ncols, uint64_t nnz,
                       const void *offsets, const void *indices, int32_t idx_width, const void *values,
                       int32_t dtype, srb_mat **out);
int32_t srb_mat_set_shard(srb_mat *m, uint64_t global_row0, uint64_t global_nrows);
int32_t srb_mat_free(srb_mat *m);
/* deep copy semantics (IMAnnData::deep_clone used by normalize_total / log1p_transform, processing/mod.rs:314-332),
 * implemented copy-on-write: the index structure (immutable) and the value buffer are shared until one side is
 * transformed, which then writes into a fresh buffer */
int32_t srb_mat_clone(srb_mat *m, srb_mat **out);
int32_t srb_mat_info(srb_mat *m, uint64_t *nrows, uint64_t *ncols, uint64_t *nnz, int32_t *format,
                     int32_t *value_dtype /* SRB_F32 or SRB_F64: current device storage */);
/* current values as f64 (what the reference holds after normalise) or f32; any of the pointers may be NULL */
int32_t srb_mat_download(srb_mat *m, uint64_t *offsets, uint64_t *indices, double *values_f64, float *values_f32);

/* Row / column subset (new device op for SURVEY §8f N2): the IMAnnData::subset step of filter_cells / filter_genes
 * (src/memory/processing/mod.rs:86-146, 245-299). keep_rows[nrows] / keep_cols[ncols] are host byte masks (non-zero =
 * keep; NULL = keep all). Produces a new compacted matrix (indices renumbered); the input is left untouched. */
int32_t srb_mat_subset(srb_mat *m, const uint8_t *keep_rows, const uint8_t *keep_cols, srb_mat **out);

/* synthetic count matrix generated on the device (bit-identical to oracle/srb_oracle.c
 * orc_synth_*): rows [row0, row0+nrows) of the global matrix for `seed`; thr/amp are host tables of ncols. */
int32_t srb_synth_csr(srb_ctx *ctx, uint32_t seed, int32_t skew, uint64_t row0, uint64_t nrows, uint32_t ncols,
                      const uint32_t *thr, const uint32_t *amp, srb_mat **out);

/* ---- statistics: shared::statistics::{number,sum,variance,minmax,stddev}::whole ------------------- */
/* All outputs are caller-allocated host buffers of length nrows (SRB_ROW) or ncols (SRB_COLUMN).
 * On a sharded matrix, SRB_ROW outputs cover the local rows; SRB_COLUMN outputs are allreduced. */
/* number_whole_helper: csr.rs:16-38, csc.rs:15-35 (counts stored entries incl. explicit zeros) */
int32_t srb_number(srb_mat *m, int32_t direction, uint32_t *out);
/* sum_whole_helper: csr.rs:81-102, csc.rs:74-95 */
int32_t srb_sum(srb_mat *m, int32_t direction, double *out);
/* variance_whole_helper: csr.rs:149-188, csc.rs:137-176 (nonzero-only; major two-pass/NaN, minor one-pass/0) */
int32_t srb_variance(srb_mat *m, int32_t direction, double *out);
/* std_dev_whole: csr.rs:225-228, csc.rs:213-216 */
int32_t srb_std_dev(srb_mat *m, int32_t direction, double *out);
/* min_max_whole_helper: csr.rs:194-223, csc.rs:182-211 (empty line -> +inf / -inf) */
int32_t srb_min_max(srb_mat *m, int32_t direction, double *out_min, double *out_max);
/* compute_qc_variables, memory/statistics/mod.rs:48-72: the 8 vectors of StatisticsContainer in one call
 * (two passes over the matrix instead of the reference's sixteen). Any pointer may be NULL. */
int32_t srb_qc_all(srb_mat *m, uint32_t *num_per_cell, uint32_t *num_per_gene, double *expr_per_cell,
                   double *expr_per_gene, double *variance_per_cell, double *variance_per_gene,
                   double *std_dev_per_cell, double *std_dev_per_gene);

/* ---- chunked accumulation: shared::statistics::{number,sum}::chunked (mod.rs:17-41, 59-83) --------- */
/* Streaming accumulator for backed (out-of-core) data. Push row-chunks (CSR) or column-chunks (CSC) in
 * order; the major offset of each chunk is tracked, so SRB_ROW results land at the correct global row
 * (the reference drops the offset: csr.rs:56-61,125-127 — documented deviation). */
typedef struct srb_stream srb_stream;
int32_t srb_stream_begin(srb_ctx *ctx, int32_t format, uint64_t nrows_total, uint64_t ncols_total, srb_stream **out);
int32_t srb_stream_push(srb_stream *s, uint64_t nmajor_chunk, uint64_t nnz, const void *offsets, const void *indices,
                        int32_t idx_width, const void *values, int32_t dtype);
int32_t srb_stream_number(srb_stream *s, int32_t direction, uint32_t *out);
int32_t srb_stream_sum(srb_stream *s, int32_t direction, double *out);
int32_t srb_stream_variance(srb_stream *s, int32_t direction, double *out); /* new: not in the reference */
/* Residency (new): keep the pushed chunks on the device as one matrix, so a backed data set that fits
 * the GPUs' HBM (180 GB each; ~6 M cells of 1500 nnz per GPU) is uploaded once, chunk by chunk, and then runs the same
 * normalise / HVG / PCA kernels as an in-memory matrix. Call before the first push; nnz_hint = 0 grows geometrically.
 * keep_statistics = 0 skips the per-chunk accumulators above. srb_stream_finish_matrix hands the matrix over. */
int32_t srb_stream_set_retain(srb_stream *s, uint64_t nnz_hint, int32_t keep_statistics);
int32_t srb_stream_finish_matrix(srb_stream *s, srb_mat **out);
int32_t srb_stream_free(srb_stream *s);

/* ---- normalisation / transform -------------------------------------------------------------------- */
/* normalize_total_inplace -> scale_row / scale_col: processing/mod.rs:303-312, scale/mod.rs:7-173.
 * scale = 0 if line sum == 0 else target/sum. Deferred: the multiply is fused into the next pass that
 * touches the values (DESIGN.md "deferred transforms"); results are identical to eager execution. */
int32_t srb_normalize_total_inplace(srb_mat *m, double target_sum, int32_t direction);
/* log1p_transform_inplace -> log1p_data: processing/mod.rs:324-326, transform/mod.rs:8-62 */
int32_t srb_log1p_inplace(srb_mat *m);

/* ---- feature selection: select_features, dim_red/mod.rs:123-156 ------------------------------------ */
/* HighlyVariable(n_top): indices in DESCENDING-variance order, ties by ascending index (stable sort).
 * *out_n = min(n_top, ncols). SRB_ERR_NAN if a variance is NaN (reference panics). */
int32_t srb_select_hvg(srb_mat *m, uint64_t n_top, uint64_t *out_idx, uint64_t *out_n);
/* VarianceThreshold(t): v > t, ascending index order */
int32_t srb_select_var_threshold(srb_mat *m, double threshold, uint64_t *out_idx, uint64_t *out_n);

/* ---- selected densify: convert_to_array_f64_selected, shared/mod.rs:230-315 ------------------------ */
/* all rows x selected columns, row-major f64, column j = col_sel[j]; parity/debug use (16 GB at 1M x 2000) */
int32_t srb_densify_selected(srb_mat *m, const uint64_t *col_sel, uint64_t n_sel, double *out);

/* ---- PCA: pca_inplace, dim_red/mod.rs:24-94 (+ single_algebra PCABuilder fit/transform) ------------- */
/* col_sel/n_sel: the selected features in selection order (from srb_select_hvg or the caller);
 * k = n_components (capped at n_sel like dim_red/mod.rs:52). center/scale as PCABuilder.
 * Outputs (host, any may be NULL): scores local_rows x k row-major (obsm["X_pca"]); components n_sel x k
 * row-major (V[:, :k], rows in selection order); explained_variance_ratio k.
 * gram_mode: 0 = tensor cores (tcgen05, split-fp16 operands, fp32 TMEM chunks, fp64 across chunks),
 *            1 = fp64 CUDA-core reference path (validation). */
int32_t srb_pca(srb_mat *m, const uint64_t *col_sel, uint64_t n_sel, uint64_t k, int32_t center, int32_t scale,
                int32_t gram_mode, double *scores, double *components, double *explained_variance_ratio);

/* One call for the headline pipeline on a raw-count matrix (BASELINE.json config 3):
 * normalize_total_inplace(target, Row) -> log1p_transform_inplace -> pca_inplace(HighlyVariable(n_top), k).
 * Equivalent to the three calls above; exists so a host can enqueue the whole step without intermediate
 * host round-trips. hvg_out receives the selection (n_top entries). */
int32_t srb_pipeline_normalize_hvg_pca(srb_mat *m, double target_sum, uint64_t n_top, uint64_t k, int32_t center,
                                       int32_t scale, int32_t gram_mode, uint64_t *hvg_out, double *scores,
                                       double *components, double *explained_variance_ratio);

/* ---- out-of-core PCA over CSR row chunks (new: SURVEY §8f N1 / BASELINE.json config 5; the reference's chunk drivers
 * stop at number / sum, src/shared/statistics/mod.rs:17-83, and src/backed/processing/mod.rs is empty) -------------------
 * Three passes over the chunks of a data set that does not fit HBM; a chunk is an srb_mat uploaded by the caller with the
 * transforms already requested (srb_normalize_total_inplace(Row) / srb_log1p_inplace are row-local, so a row chunk
 * transforms exactly like the whole matrix):
 *   pass 1  srb_gene_moments(chunk): count, sum, sum of squares per gene of the chunk's current values; the caller adds the
 *           O(genes) vectors over the chunks and selects the features as select_features does (dim_red/mod.rs:123-156)
 *   pass 2  srb_pca_stream_begin(global moments, selection) ... srb_pca_stream_push_gram(chunk) ... srb_pca_stream_fit
 *           (on a multi-rank ctx each rank pushes its own chunks and fit allreduces the Gram matrix)
 *   pass 3  srb_pca_stream_transform(chunk) -> scores of the chunk's rows (chunk rows x k, row-major)
 * Results equal srb_pca on the whole matrix up to the summation order of the Gram matrix. */
typedef struct srb_pca_stream srb_pca_stream;
int32_t srb_gene_moments(srb_mat *chunk, double *count, double *sum, double *sumsq); /* each ncols long; may be NULL */
int32_t srb_pca_stream_begin(srb_ctx *ctx, uint64_t ncols, uint64_t ncells_total, const double *gene_sum,
                             const double *gene_sumsq, const uint64_t *col_sel, uint64_t n_sel, uint64_t k, int32_t center,
                             int32_t scale, int32_t gram_mode, srb_pca_stream **out);
int32_t srb_pca_stream_push_gram(srb_pca_stream *ps, srb_mat *chunk);
int32_t srb_pca_stream_fit(srb_pca_stream *ps, double *components /* n_sel x k */, double *explained_variance_ratio /* k */);
int32_t srb_pca_stream_transform(srb_pca_stream *ps, srb_mat *chunk, double *scores /* chunk rows x k */);
int32_t srb_pca_stream_free(srb_pca_stream *ps);

/* per-stage device times (ms) of the last srb_pca / srb_pipeline_* call on this matrix's ctx, measured with
 * CUDA events on the ctx stream: [0] row sums, [1] fused normalise+log1p+gene moments, [2] hvg select,
 * [3] densify, [4] gram, [5] eig, [6] scores, [7] moments allreduce, [8] gram allreduce. n = capacity of out_ms. */
int32_t srb_last_stage_ms(srb_ctx *ctx, float *out_ms, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H */
