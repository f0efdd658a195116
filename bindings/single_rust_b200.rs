// single_rust_b200.rs — UNCOMPILED mirror of include/srb200.h for the reference crate (no Rust toolchain exists in the
// build image). Drop into src/b200/sys.rs and link with `cargo:rustc-link-lib=dylib=srb200`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct srb_ctx { _p: [u8; 0] }
#[repr(C)] pub struct srb_mat { _p: [u8; 0] }
#[repr(C)] pub struct srb_stream { _p: [u8; 0] }
#[repr(C)] pub struct srb_pca_stream { _p: [u8; 0] }

pub const SRB_ROW: i32 = 0;      // Direction::Row    (src/shared/mod.rs:39-42)
pub const SRB_COLUMN: i32 = 1;   // Direction::Column
pub const SRB_CSR: i32 = 0;
pub const SRB_CSC: i32 = 1;
pub const SRB_IDX64: i32 = 8;    // usize offsets / indices of nalgebra-sparse
// srb_dtype: I8=0 I16=1 I32=2 (I64=3) U8=4 U16=5 U32=6 (U64=7) F32=8 F64=9

#[link(name = "srb200")]
extern "C" {
    pub fn srb_version() -> *const c_char;
    pub fn srb_last_error_message() -> *const c_char;
    pub fn srb_kernel_launch_count() -> u64;
    pub fn srb_host_delta_encode(indices: *const c_void, offsets: *const c_void, idx_width: i32, nmajor: u64, nnz: u64,
                                 bound: u64, chunk: u64, nthreads: i32, codes: *mut u8, esc_pos: *mut u64,
                                 esc_val: *mut u32, esc_cap: u64, n_esc: *mut u64, out_of_bounds: *mut i32) -> i32;
    pub fn srb_host_pack_values_f32(src: *const f32, n: u64, dst: *mut c_void, dst_width: i32, nthreads: i32,
                                    lossless: *mut i32) -> i32;
    pub fn srb_host_pack_indices(src: *const c_void, src_width: i32, n: u64, dst: *mut c_void, dst_width: i32,
                                 bound: u64, nthreads: i32, out_of_bounds: *mut i32) -> i32;
    pub fn srb_ctx_create(device: i32, out: *mut *mut srb_ctx) -> i32;
    pub fn srb_ctx_destroy(ctx: *mut srb_ctx) -> i32;
    pub fn srb_ctx_set_value_mode(ctx: *mut srb_ctx, mode: i32) -> i32;
    pub fn srb_ctx_set_upload_mode(ctx: *mut srb_ctx, mode: i32) -> i32;   // 0 device-narrow, 1 host-pack, 2 auto, 3 host-pack + values, 4 host-pack + values while the host is ahead, 5 host-pack with delta-coded indices
    pub fn srb_ctx_set_eig_mode(ctx: *mut srb_ctx, mode: i32) -> i32;      // 0 syevd, 1 chfsi
    pub fn srb_ctx_last_eig(ctx: *mut srb_ctx, solver: *mut i32, block_products: *mut i32, outer_iterations: *mut i32,
                            max_residual: *mut f64) -> i32;
    pub fn srb_upload_mix(t_idx_ms: f64, t_val_ms: f64, len: u64, packed_index_bytes: i32, idx_width: i32, value_bytes: i32, link_gbs: f64, raw_index_fraction: *mut f64, packed_value_fraction: *mut f64) -> i32;
    pub fn srb_ctx_last_upload(ctx: *mut srb_ctx, h2d_bytes: *mut u64, host_packed: *mut i32) -> i32;
    pub fn srb_ctx_last_upload_chunks(ctx: *mut srb_ctx, chunks: *mut i32, index_chunks_packed: *mut i32, value_chunks_packed: *mut i32) -> i32;
    pub fn srb_ctx_synchronize(ctx: *mut srb_ctx) -> i32;
    pub fn srb_ctx_stream(ctx: *mut srb_ctx) -> *mut c_void;
    pub fn srb_last_stage_ms(ctx: *mut srb_ctx, out_ms: *mut f32, n: i32) -> i32;
    pub fn srb_comm_unique_id(id128: *mut c_void) -> i32;
    pub fn srb_ctx_comm_init(ctx: *mut srb_ctx, id128: *const c_void, rank: i32, nranks: i32) -> i32;

    pub fn srb_mat_upload(ctx: *mut srb_ctx, format: i32, nrows: u64, ncols: u64, nnz: u64,
                          offsets: *const c_void, indices: *const c_void, idx_width: i32,
                          values: *const c_void, dtype: i32, out: *mut *mut srb_mat) -> i32;
    pub fn srb_mat_set_shard(m: *mut srb_mat, global_row0: u64, global_nrows: u64) -> i32;
    pub fn srb_mat_clone(m: *mut srb_mat, out: *mut *mut srb_mat) -> i32;
    pub fn srb_mat_subset(m: *mut srb_mat, keep_rows: *const u8, keep_cols: *const u8, out: *mut *mut srb_mat) -> i32;
    pub fn srb_mat_free(m: *mut srb_mat) -> i32;
    pub fn srb_mat_info(m: *mut srb_mat, nrows: *mut u64, ncols: *mut u64, nnz: *mut u64, format: *mut i32,
                        value_dtype: *mut i32) -> i32;
    pub fn srb_synth_csr(ctx: *mut srb_ctx, seed: u32, skew: i32, row0: u64, nrows: u64, ncols: u32,
                         thr: *const u32, amp: *const u32, out: *mut *mut srb_mat) -> i32;
    pub fn srb_mat_download(m: *mut srb_mat, offsets: *mut u64, indices: *mut u64,
                            values_f64: *mut f64, values_f32: *mut f32) -> i32;

    pub fn srb_number(m: *mut srb_mat, direction: i32, out: *mut u32) -> i32;
    pub fn srb_sum(m: *mut srb_mat, direction: i32, out: *mut f64) -> i32;
    pub fn srb_variance(m: *mut srb_mat, direction: i32, out: *mut f64) -> i32;
    pub fn srb_std_dev(m: *mut srb_mat, direction: i32, out: *mut f64) -> i32;
    pub fn srb_min_max(m: *mut srb_mat, direction: i32, mn: *mut f64, mx: *mut f64) -> i32;
    pub fn srb_qc_all(m: *mut srb_mat, num_per_cell: *mut u32, num_per_gene: *mut u32,
                      expr_per_cell: *mut f64, expr_per_gene: *mut f64,
                      variance_per_cell: *mut f64, variance_per_gene: *mut f64,
                      std_dev_per_cell: *mut f64, std_dev_per_gene: *mut f64) -> i32;

    pub fn srb_normalize_total_inplace(m: *mut srb_mat, target_sum: f64, direction: i32) -> i32;
    pub fn srb_log1p_inplace(m: *mut srb_mat) -> i32;
    pub fn srb_select_hvg(m: *mut srb_mat, n_top: u64, out_idx: *mut u64, out_n: *mut u64) -> i32;
    pub fn srb_select_var_threshold(m: *mut srb_mat, t: f64, out_idx: *mut u64, out_n: *mut u64) -> i32;
    pub fn srb_densify_selected(m: *mut srb_mat, col_sel: *const u64, n_sel: u64, out: *mut f64) -> i32;
    pub fn srb_pca(m: *mut srb_mat, col_sel: *const u64, n_sel: u64, k: u64, center: i32, scale: i32,
                   gram_mode: i32, scores: *mut f64, components: *mut f64, evr: *mut f64) -> i32;

    pub fn srb_pipeline_normalize_hvg_pca(m: *mut srb_mat, target_sum: f64, n_top: u64, k: u64, center: i32, scale: i32,
                                          gram_mode: i32, hvg_out: *mut u64, scores: *mut f64, components: *mut f64,
                                          evr: *mut f64) -> i32;

    pub fn srb_gene_moments(chunk: *mut srb_mat, count: *mut f64, sum: *mut f64, sumsq: *mut f64) -> i32;
    pub fn srb_pca_stream_begin(ctx: *mut srb_ctx, ncols: u64, ncells_total: u64, gene_sum: *const f64, gene_sumsq: *const f64,
                                col_sel: *const u64, n_sel: u64, k: u64, center: i32, scale: i32, gram_mode: i32,
                                out: *mut *mut srb_pca_stream) -> i32;
    pub fn srb_pca_stream_push_gram(ps: *mut srb_pca_stream, chunk: *mut srb_mat) -> i32;
    pub fn srb_pca_stream_fit(ps: *mut srb_pca_stream, components: *mut f64, evr: *mut f64) -> i32;
    pub fn srb_pca_stream_transform(ps: *mut srb_pca_stream, chunk: *mut srb_mat, scores: *mut f64) -> i32;
    pub fn srb_pca_stream_free(ps: *mut srb_pca_stream) -> i32;

    pub fn srb_stream_begin(ctx: *mut srb_ctx, format: i32, nrows_total: u64, ncols_total: u64,
                            out: *mut *mut srb_stream) -> i32;
    pub fn srb_stream_push(s: *mut srb_stream, nmajor_chunk: u64, nnz: u64, offsets: *const c_void,
                           indices: *const c_void, idx_width: i32, values: *const c_void, dtype: i32) -> i32;
    pub fn srb_stream_number(s: *mut srb_stream, direction: i32, out: *mut u32) -> i32;
    pub fn srb_stream_sum(s: *mut srb_stream, direction: i32, out: *mut f64) -> i32;
    pub fn srb_stream_variance(s: *mut srb_stream, direction: i32, out: *mut f64) -> i32;
    pub fn srb_stream_set_retain(s: *mut srb_stream, nnz_hint: u64, keep_statistics: i32) -> i32;
    pub fn srb_stream_finish_matrix(s: *mut srb_stream, out: *mut *mut srb_mat) -> i32;
    pub fn srb_stream_free(s: *mut srb_stream) -> i32;
}

fn check(rc: i32) -> anyhow::Result<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(srb_last_error_message()) }.to_string_lossy().into_owned();
    anyhow::bail!("srb200 error {rc}: {msg}")      // the crate's convention is anyhow::Result everywhere
}
