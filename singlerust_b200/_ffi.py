"""ctypes binding of libsrb200.so (include/srb200.h). The same symbols a Rust `extern "C"` block binds
(INTEGRATION.md). There is NO CPU fallback: if the library is missing or no sm_100 GPU is present, calls raise."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ROW, COLUMN = 0, 1
CSR, CSC = 0, 1
VALUES_COMPACT, VALUES_FAITHFUL = 0, 1
GRAM_TENSOR, GRAM_FP64 = 0, 1
EIG_SYEVD, EIG_CHFSI = 0, 1
UPLOAD_DEVICE_NARROW, UPLOAD_HOST_PACK, UPLOAD_AUTO, UPLOAD_HOST_PACK_VALUES, UPLOAD_HOST_PACK_ADAPTIVE = 0, 1, 2, 3, 4
UPLOAD_HOST_PACK_DELTA = 5
UPLOAD_BALANCED = 6

DTYPES = {np.dtype(np.int8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
          np.dtype(np.uint8): 4, np.dtype(np.uint16): 5, np.dtype(np.uint32): 6, np.dtype(np.uint64): 7,
          np.dtype(np.float32): 8, np.dtype(np.float64): 9}
F32, F64 = 8, 9

STAGES = ["row_sums", "fused_norm_log1p_moments", "hvg_select", "densify", "gram", "eig", "scores", "allreduce_moments",
          "allreduce_gram"]


class SrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"srb200 error {code}: {msg}")
        self.code = code


def lib_path() -> str:
    return os.path.join(_HERE, "libsrb200.so")


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise RuntimeError(f"{p} is missing: run `python -m singlerust_b200.build` (no CPU fallback exists)")
        L = C.CDLL(p)
        L.srb_version.restype = C.c_char_p
        L.srb_last_error_message.restype = C.c_char_p
        L.srb_kernel_launch_count.restype = C.c_uint64
        L.srb_ctx_stream.restype = C.c_void_p
        L.srb_ctx_stream.argtypes = [C.c_void_p]
        vp, u64, i32, f64 = C.c_void_p, C.c_uint64, C.c_int32, C.c_double
        sig = {
            "srb_ctx_create": [i32, C.POINTER(vp)],
            "srb_ctx_destroy": [vp],
            "srb_ctx_set_value_mode": [vp, i32],
            "srb_ctx_set_upload_mode": [vp, i32],
            "srb_host_pack_indices": [vp, i32, u64, vp, i32, u64, i32, C.POINTER(i32)],
            "srb_upload_mix": [f64, f64, u64, i32, i32, i32, f64, C.POINTER(f64), C.POINTER(f64)],
            "srb_ctx_last_upload": [vp, C.POINTER(u64), C.POINTER(i32)],
            "srb_ctx_last_upload_chunks": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
            "srb_ctx_set_eig_mode": [vp, i32],
            "srb_ctx_last_eig": [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(f64)],
            "srb_host_pack_values_f32": [vp, u64, vp, i32, i32, C.POINTER(i32)],
            "srb_host_delta_encode": [vp, vp, i32, u64, u64, u64, u64, i32, vp, vp, vp, u64, C.POINTER(u64), C.POINTER(i32)],
            "srb_ctx_synchronize": [vp],
            "srb_comm_unique_id": [vp],
            "srb_ctx_comm_init": [vp, vp, i32, i32],
            "srb_mat_upload": [vp, i32, u64, u64, u64, vp, vp, i32, vp, i32, C.POINTER(vp)],
            "srb_mat_set_shard": [vp, u64, u64],
            "srb_mat_free": [vp],
            "srb_mat_clone": [vp, C.POINTER(vp)],
            "srb_mat_subset": [vp, vp, vp, C.POINTER(vp)],
            "srb_mat_info": [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(i32), C.POINTER(i32)],
            "srb_mat_download": [vp, vp, vp, vp, vp],
            "srb_synth_csr": [vp, C.c_uint32, i32, u64, u64, C.c_uint32, vp, vp, C.POINTER(vp)],
            "srb_number": [vp, i32, vp],
            "srb_sum": [vp, i32, vp],
            "srb_variance": [vp, i32, vp],
            "srb_std_dev": [vp, i32, vp],
            "srb_min_max": [vp, i32, vp, vp],
            "srb_qc_all": [vp] + [vp] * 8,
            "srb_stream_begin": [vp, i32, u64, u64, C.POINTER(vp)],
            "srb_stream_push": [vp, u64, u64, vp, vp, i32, vp, i32],
            "srb_stream_number": [vp, i32, vp],
            "srb_stream_sum": [vp, i32, vp],
            "srb_stream_variance": [vp, i32, vp],
            "srb_stream_set_retain": [vp, u64, i32],
            "srb_stream_finish_matrix": [vp, C.POINTER(vp)],
            "srb_stream_free": [vp],
            "srb_normalize_total_inplace": [vp, f64, i32],
            "srb_log1p_inplace": [vp],
            "srb_select_hvg": [vp, u64, vp, C.POINTER(u64)],
            "srb_select_var_threshold": [vp, f64, vp, C.POINTER(u64)],
            "srb_densify_selected": [vp, vp, u64, vp],
            "srb_pca": [vp, vp, u64, u64, i32, i32, i32, vp, vp, vp],
            "srb_pipeline_normalize_hvg_pca": [vp, f64, u64, u64, i32, i32, i32, vp, vp, vp, vp],
            "srb_last_stage_ms": [vp, vp, i32],
            "srb_gene_moments": [vp, vp, vp, vp],
            "srb_pca_stream_begin": [vp, u64, u64, vp, vp, vp, u64, u64, i32, i32, i32, C.POINTER(vp)],
            "srb_pca_stream_push_gram": [vp, vp],
            "srb_pca_stream_fit": [vp, vp, vp],
            "srb_pca_stream_transform": [vp, vp, vp],
            "srb_pca_stream_free": [vp],
        }
        for name, args in sig.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int32
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        raise SrbError(rc, lib().srb_last_error_message().decode(errors="replace"))


def _ptr(a):
    """numpy array | torch tensor | int address | None -> void*"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"], "host arrays must be contiguous"
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def version() -> str:
    return lib().srb_version().decode()


def host_pack_indices(src: np.ndarray, dst_width: int, bound: int, nthreads: int = 0):
    """Host-side narrowing used by the HOST_PACK upload (no GPU). Returns (packed array, any value >= bound)."""
    assert src.dtype in (np.uint32, np.uint64, np.int32, np.int64) and src.flags["C_CONTIGUOUS"]
    dst = np.empty(src.shape[0], dtype=np.uint16 if dst_width == 2 else np.uint32)
    oob = C.c_int32(0)
    rc = lib().srb_host_pack_indices(_ptr(src), src.dtype.itemsize, src.shape[0], _ptr(dst), dst_width, bound, nthreads, C.byref(oob))
    if rc != 0:
        raise ValueError("srb_host_pack_indices: bad argument")
    return dst, bool(oob.value)


def host_delta_encode(offsets: np.ndarray, indices: np.ndarray, bound: int, chunk: int = 1 << 22, nthreads: int = 0):
    """Delta coder of the HOST_PACK_DELTA upload (no GPU). Returns (codes u8, escape positions, escape values, oob)."""
    assert offsets.dtype == indices.dtype and offsets.dtype in (np.uint64, np.uint32)
    offsets, indices = np.ascontiguousarray(offsets), np.ascontiguousarray(indices)
    nnz, cap = indices.shape[0], max(1024, indices.shape[0])
    codes, pos, val = np.empty(nnz, np.uint8), np.empty(cap, np.uint64), np.empty(cap, np.uint32)
    n, oob = C.c_uint64(0), C.c_int32(0)
    rc = lib().srb_host_delta_encode(_ptr(indices), _ptr(offsets), indices.dtype.itemsize, offsets.shape[0] - 1, nnz, bound, chunk,
                                     nthreads, _ptr(codes), _ptr(pos), _ptr(val), cap, C.byref(n), C.byref(oob))
    if rc != 0:
        raise ValueError(f"srb_host_delta_encode: {rc}")
    return codes, pos[:n.value].copy(), val[:n.value].copy(), bool(oob.value)


def host_pack_values_f32(src: np.ndarray, dst_width: int, nthreads: int = 0):
    """Host-side lossless narrowing of f32 counts (no GPU). Returns (packed array, lossless)."""
    assert src.dtype == np.float32 and src.flags["C_CONTIGUOUS"]
    dst = np.empty(src.shape[0], dtype=np.uint8 if dst_width == 1 else np.uint16)
    ok = C.c_int32(0)
    if lib().srb_host_pack_values_f32(_ptr(src), src.shape[0], _ptr(dst), dst_width, nthreads, C.byref(ok)) != 0:
        raise ValueError("srb_host_pack_values_f32: bad argument")
    return dst, bool(ok.value)


def upload_mix(t_idx_ms, t_val_ms, n_entries, packed_index_bytes=1, idx_width=8, value_bytes=4, link_gbs=50.0):
    """Rate model of the BALANCED upload (no GPU): (share of chunks with raw indices, share of chunks with packed values)."""
    g, f = C.c_double(0.0), C.c_double(0.0)
    if lib().srb_upload_mix(t_idx_ms, t_val_ms, n_entries, packed_index_bytes, idx_width, value_bytes, link_gbs, C.byref(g), C.byref(f)) != 0:
        raise ValueError("srb_upload_mix: bad argument")
    return g.value, f.value


def kernel_launch_count() -> int:
    return int(lib().srb_kernel_launch_count())


class Context:
    """One GPU + one CUDA stream (+ one NCCL rank)."""

    def __init__(self, device: int = 0, value_mode: int = VALUES_COMPACT):
        self._h = C.c_void_p()
        check(lib().srb_ctx_create(device, C.byref(self._h)))
        self.device = device
        self.rank, self.nranks = 0, 1
        if value_mode != VALUES_COMPACT:
            self.set_value_mode(value_mode)

    def set_value_mode(self, mode):
        check(lib().srb_ctx_set_value_mode(self._h, mode))

    def set_upload_mode(self, mode):
        check(lib().srb_ctx_set_upload_mode(self._h, mode))

    def set_eig_mode(self, mode):
        check(lib().srb_ctx_set_eig_mode(self._h, mode))

    def last_eig(self):
        """dict(solver 'syevd' | 'chfsi' | 'chfsi->syevd', block_products, outer_iterations, max_residual) of the last K8."""
        s, p, o, r = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        check(lib().srb_ctx_last_eig(self._h, C.byref(s), C.byref(p), C.byref(o), C.byref(r)))
        return dict(solver=("syevd", "chfsi", "chfsi->syevd")[s.value], block_products=p.value, outer_iterations=o.value,
                    max_residual=r.value)

    def last_upload(self):
        """(bytes the last upload moved over the link, index array was host-packed)"""
        b, p = C.c_uint64(0), C.c_int32(0)
        check(lib().srb_ctx_last_upload(self._h, C.byref(b), C.byref(p)))
        return int(b.value), bool(p.value)

    def last_upload_chunks(self):
        """BALANCED upload: (chunks, chunks whose indices were host-packed, chunks whose values were host-packed)"""
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(lib().srb_ctx_last_upload_chunks(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def synchronize(self):
        check(lib().srb_ctx_synchronize(self._h))

    @property
    def stream(self) -> int:
        return int(lib().srb_ctx_stream(self._h) or 0)

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().srb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        check(lib().srb_ctx_comm_init(self._h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def last_stage_ms(self) -> dict:
        out = np.zeros(len(STAGES), dtype=np.float32)
        check(lib().srb_last_stage_ms(self._h, _ptr(out), len(STAGES)))
        return dict(zip(STAGES, out.tolist()))

    def close(self):
        if self._h:
            lib().srb_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMatrix:
    """Device-resident compressed matrix (one row shard)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self._h = ctx, handle

    # ---- construction ----
    @classmethod
    def upload(cls, ctx, fmt, nrows, ncols, offsets, indices, values, nnz=None, idx_width=None, dtype=None):
        """offsets/indices: uint64 (Rust usize) or uint32/int32 arrays (or raw addresses with idx_width given)."""
        if isinstance(offsets, np.ndarray):
            if offsets.dtype.itemsize == 8:
                offsets, indices = np.ascontiguousarray(offsets, np.uint64), np.ascontiguousarray(indices, np.uint64)
                idx_width = 8
            else:
                offsets, indices = np.ascontiguousarray(offsets, np.uint32), np.ascontiguousarray(indices, np.uint32)
                idx_width = 4
            nnz = int(offsets[-1]) if nnz is None else nnz
        if isinstance(values, np.ndarray):
            values = np.ascontiguousarray(values)
            dtype = DTYPES[values.dtype]
        h = C.c_void_p()
        check(lib().srb_mat_upload(ctx._h, fmt, nrows, ncols, nnz, _ptr(offsets), _ptr(indices), idx_width, _ptr(values),
                                   dtype, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_scipy(cls, ctx, m, index_dtype=np.uint64):
        import scipy.sparse as sp
        fmt = CSR if (sp.isspmatrix_csr(m) or isinstance(m, sp.csr_array)) else CSC
        return cls.upload(ctx, fmt, m.shape[0], m.shape[1], m.indptr.astype(index_dtype), m.indices.astype(index_dtype),
                          m.data)

    @classmethod
    def synth(cls, ctx, seed, nrows, ncols, thr, amp, row0=0, skew=False):
        thr = np.ascontiguousarray(thr, np.uint32)
        amp = np.ascontiguousarray(amp, np.uint32)
        h = C.c_void_p()
        check(lib().srb_synth_csr(ctx._h, seed, int(skew), row0, nrows, ncols, _ptr(thr), _ptr(amp), C.byref(h)))
        return cls(ctx, h)

    def set_shard(self, global_row0, global_nrows):
        check(lib().srb_mat_set_shard(self._h, global_row0, global_nrows))

    def clone(self):
        h = C.c_void_p()
        check(lib().srb_mat_clone(self._h, C.byref(h)))
        return DeviceMatrix(self.ctx, h)

    def subset(self, keep_rows=None, keep_cols=None):
        """Row / column compaction (IMAnnData::subset): boolean masks, None = keep all."""
        kr = None if keep_rows is None else np.ascontiguousarray(keep_rows, np.uint8)
        kc = None if keep_cols is None else np.ascontiguousarray(keep_cols, np.uint8)
        nr, nc = self.shape
        assert kr is None or kr.size == nr
        assert kc is None or kc.size == nc
        h = C.c_void_p()
        check(lib().srb_mat_subset(self._h, _ptr(kr), _ptr(kc), C.byref(h)))
        return DeviceMatrix(self.ctx, h)

    def free(self):
        if self._h:
            lib().srb_mat_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- info / download ----
    def info(self):
        nr, nc, nnz = C.c_uint64(), C.c_uint64(), C.c_uint64()
        fmt, vd = C.c_int32(), C.c_int32()
        check(lib().srb_mat_info(self._h, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(fmt), C.byref(vd)))
        return dict(nrows=nr.value, ncols=nc.value, nnz=nnz.value, format=fmt.value, value_dtype=vd.value)

    @property
    def shape(self):
        i = self.info()
        return (i["nrows"], i["ncols"])

    def _len(self, direction):
        i = self.info()
        return i["nrows"] if direction == ROW else i["ncols"]

    def download(self, structure=True, values="f64"):
        i = self.info()
        nmajor = i["nrows"] if i["format"] == CSR else i["ncols"]
        off = np.zeros(nmajor + 1, np.uint64) if structure else None
        idx = np.zeros(i["nnz"], np.uint64) if structure else None
        v64 = np.zeros(i["nnz"], np.float64) if values == "f64" else None
        v32 = np.zeros(i["nnz"], np.float32) if values == "f32" else None
        check(lib().srb_mat_download(self._h, _ptr(off), _ptr(idx), _ptr(v64), _ptr(v32)))
        return off, idx, (v64 if v64 is not None else v32)

    # ---- statistics ----
    def number(self, direction):
        out = np.zeros(self._len(direction), np.uint32)
        check(lib().srb_number(self._h, direction, _ptr(out)))
        return out

    def _f64(self, fn, direction):
        out = np.zeros(self._len(direction), np.float64)
        check(fn(self._h, direction, _ptr(out)))
        return out

    def sum(self, direction):
        return self._f64(lib().srb_sum, direction)

    def variance(self, direction):
        return self._f64(lib().srb_variance, direction)

    def std_dev(self, direction):
        return self._f64(lib().srb_std_dev, direction)

    def min_max(self, direction):
        n = self._len(direction)
        mn, mx = np.zeros(n), np.zeros(n)
        check(lib().srb_min_max(self._h, direction, _ptr(mn), _ptr(mx)))
        return mn, mx

    def qc_all(self):
        nr, nc = self.shape
        o = dict(num_per_cell=np.zeros(nr, np.uint32), num_per_gene=np.zeros(nc, np.uint32),
                 expr_per_cell=np.zeros(nr), expr_per_gene=np.zeros(nc), variance_per_cell=np.zeros(nr),
                 variance_per_gene=np.zeros(nc), std_dev_per_cell=np.zeros(nr), std_dev_per_gene=np.zeros(nc))
        check(lib().srb_qc_all(self._h, *[_ptr(o[k]) for k in
                                          ("num_per_cell", "num_per_gene", "expr_per_cell", "expr_per_gene",
                                           "variance_per_cell", "variance_per_gene", "std_dev_per_cell",
                                           "std_dev_per_gene")]))
        return o

    def gene_moments(self):
        """(count, sum, sum of squares) per gene of the current values of this CSR (chunk); never reduced over ranks."""
        nc = self.shape[1]
        cnt, s, q = np.zeros(nc), np.zeros(nc), np.zeros(nc)
        check(lib().srb_gene_moments(self._h, _ptr(cnt), _ptr(s), _ptr(q)))
        return cnt, s, q

    # ---- transforms ----
    def normalize_total_inplace(self, target_sum, direction=ROW):
        check(lib().srb_normalize_total_inplace(self._h, float(target_sum), direction))

    def log1p_inplace(self):
        check(lib().srb_log1p_inplace(self._h))

    # ---- selection / PCA ----
    def select_hvg(self, n_top):
        out = np.zeros(min(n_top, self.shape[1]), np.uint64)
        n = C.c_uint64()
        check(lib().srb_select_hvg(self._h, n_top, _ptr(out), C.byref(n)))
        return out[:n.value]

    def select_var_threshold(self, t):
        out = np.zeros(self.shape[1], np.uint64)
        n = C.c_uint64()
        check(lib().srb_select_var_threshold(self._h, float(t), _ptr(out), C.byref(n)))
        return out[:n.value].copy()

    def densify_selected(self, col_sel):
        sel = np.ascontiguousarray(col_sel, np.uint64)
        out = np.zeros((self.shape[0], sel.size), np.float64)
        check(lib().srb_densify_selected(self._h, _ptr(sel), sel.size, _ptr(out)))
        return out

    def pca(self, col_sel, k, center=True, scale=True, gram_mode=GRAM_TENSOR, want_scores=True, scores_out=None):
        sel = np.ascontiguousarray(col_sel, np.uint64)
        k = min(int(k), sel.size)
        nr = self.shape[0]
        scores = scores_out if scores_out is not None else (np.zeros((nr, k)) if want_scores else None)
        comps, evr = np.zeros((sel.size, k)), np.zeros(k)
        check(lib().srb_pca(self._h, _ptr(sel), sel.size, k, int(center), int(scale), gram_mode, _ptr(scores), _ptr(comps),
                            _ptr(evr)))
        return dict(scores=scores, components=comps, explained_variance_ratio=evr, selection=sel)

    def pipeline_normalize_hvg_pca(self, target_sum, n_top, k, center=True, scale=True, gram_mode=GRAM_TENSOR,
                                   want_scores=True, scores_out=None, want_outputs=True):
        nr, nc = self.shape
        n_sel = min(n_top, nc)
        k = min(int(k), n_sel)
        hvg = np.zeros(n_sel, np.uint64) if want_outputs else None
        scores = scores_out if scores_out is not None else (np.zeros((nr, k)) if (want_scores and want_outputs) else None)
        comps = np.zeros((n_sel, k)) if want_outputs else None
        evr = np.zeros(k) if want_outputs else None
        check(lib().srb_pipeline_normalize_hvg_pca(self._h, float(target_sum), n_top, k, int(center), int(scale), gram_mode,
                                                   _ptr(hvg), _ptr(scores), _ptr(comps), _ptr(evr)))
        return dict(scores=scores, components=comps, explained_variance_ratio=evr, selection=hvg)


class PcaStream:
    """Out-of-core PCA over CSR row chunks (srb_pca_stream_*): push_gram every chunk, fit, then transform every chunk."""

    def __init__(self, ctx, ncols, ncells_total, gene_sum, gene_sumsq, col_sel, k, center=True, scale=True, gram_mode=GRAM_TENSOR):
        self.ctx = ctx
        self.sel = np.ascontiguousarray(col_sel, np.uint64)
        self.k = min(int(k), self.sel.size)
        gs, gq = np.ascontiguousarray(gene_sum, np.float64), np.ascontiguousarray(gene_sumsq, np.float64)
        assert gs.shape == (ncols,) and gq.shape == (ncols,)
        self._h = C.c_void_p()
        check(lib().srb_pca_stream_begin(ctx._h, ncols, ncells_total, _ptr(gs), _ptr(gq), _ptr(self.sel), self.sel.size, self.k,
                                         int(center), int(scale), gram_mode, C.byref(self._h)))

    def push_gram(self, chunk: "DeviceMatrix"):
        check(lib().srb_pca_stream_push_gram(self._h, chunk._h))

    def fit(self):
        comps, evr = np.zeros((self.sel.size, self.k)), np.zeros(self.k)
        check(lib().srb_pca_stream_fit(self._h, _ptr(comps), _ptr(evr)))
        return comps, evr

    def transform(self, chunk: "DeviceMatrix", out=None):
        n = chunk.shape[0]
        out = np.zeros((n, self.k)) if out is None else out
        assert out.shape == (n, self.k) and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        check(lib().srb_pca_stream_transform(self._h, chunk._h, _ptr(out)))
        return out

    def free(self):
        if self._h:
            check(lib().srb_pca_stream_free(self._h))
            self._h = C.c_void_p()


class ChunkStream:
    """shared::statistics::{number,sum}::chunked accumulator (srb_stream_*)."""

    def __init__(self, ctx, fmt, nrows_total, ncols_total):
        self.ctx, self.fmt, self.nrows, self.ncols = ctx, fmt, nrows_total, ncols_total
        self._h = C.c_void_p()
        check(lib().srb_stream_begin(ctx._h, fmt, nrows_total, ncols_total, C.byref(self._h)))

    def push(self, offsets, indices, values):
        offsets, indices = np.ascontiguousarray(offsets, np.uint64), np.ascontiguousarray(indices, np.uint64)
        values = np.ascontiguousarray(values)
        check(lib().srb_stream_push(self._h, offsets.size - 1, int(offsets[-1]), _ptr(offsets), _ptr(indices), 8,
                                    _ptr(values), DTYPES[values.dtype]))

    def set_retain(self, nnz_hint=0, keep_statistics=True):
        check(lib().srb_stream_set_retain(self._h, nnz_hint, int(keep_statistics)))

    def finish_matrix(self) -> "DeviceMatrix":
        h = C.c_void_p()
        check(lib().srb_stream_finish_matrix(self._h, C.byref(h)))
        return DeviceMatrix(self.ctx, h)

    def _len(self, direction):
        return self.nrows if direction == ROW else self.ncols

    def number(self, direction):
        out = np.zeros(self._len(direction), np.uint32)
        check(lib().srb_stream_number(self._h, direction, _ptr(out)))
        return out

    def sum(self, direction):
        out = np.zeros(self._len(direction))
        check(lib().srb_stream_sum(self._h, direction, _ptr(out)))
        return out

    def variance(self, direction):
        out = np.zeros(self._len(direction))
        check(lib().srb_stream_variance(self._h, direction, _ptr(out)))
        return out

    def free(self):
        if self._h:
            lib().srb_stream_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
