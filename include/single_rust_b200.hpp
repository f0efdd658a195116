// single_rust_b200.hpp — C++17 host side above the C ABI (include/srb200.h), header-only.
//
// The reference (SingleRust, /root/reference) is Rust; this image has no Rust toolchain, so the host side that a Rust
// maintainer would write over `extern "C"` (bindings/single_rust_b200.rs) is written here in C++ with the
// reference's own module / function names, argument order and error behaviour:
//
//   single_rust::shared::{Direction, ComputationMode, FeatureSelection, FlexValue}     src/shared/mod.rs:17-102
//   single_rust::memory::statistics::{compute_number, compute_sum, compute_variance,
//                compute_min_max, compute_std_dev, compute_qc_variables, qc_vars_inplace} src/memory/statistics/mod.rs:10-103
//   single_rust::memory::processing::{filter_cells{,_inplace}, filter_genes{,_inplace},
//                normalize_total{,_inplace}, log1p_transform{,_inplace},
//                select_features, pca_inplace}                                          src/memory/processing/**
//   single_rust::backed::statistics::{compute_number, compute_sum}                      src/backed/statistics/mod.rs:5-45
//
// `anyhow::Result<T>` becomes "returns T or throws single_rust::Error" (code = srb_status, what() = the library's
// message); the reference's panics / todo!()s surface as the same exception instead of aborting. Every function body is
// one or a few C-ABI calls: no arithmetic on matrix data happens on the host, and there is no CPU fallback.
#ifndef SINGLE_RUST_B200_HPP
#define SINGLE_RUST_B200_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <map>
#include <numeric>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "srb200.h"

namespace single_rust {

/// anyhow::Error stand-in: `code` is the srb_status of the failing ABI call (or SRB_ERR_INVALID_ARG for host checks).
class Error : public std::runtime_error {
public:
    int32_t code;
    Error(int32_t c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

namespace detail {
inline void check(int32_t rc) {
    if (rc != SRB_OK) throw Error(rc, srb_last_error_message());
}
}  // namespace detail

// =====================================================================================================================
// shared vocabulary — src/shared/mod.rs
// =====================================================================================================================
namespace shared {

/// src/shared/mod.rs:39-60 (same discriminants; they cross the ABI as int32)
enum class Direction : int32_t { Row = SRB_ROW, Column = SRB_COLUMN };
inline bool is_row(Direction d) { return d == Direction::Row; }

/// src/shared/mod.rs:25-37
struct ComputationMode {
    std::optional<size_t> chunk;  // nullopt = Whole
    static ComputationMode Whole() { return {}; }
    static ComputationMode Chunked(size_t n) { return ComputationMode{n}; }
    bool is_whole() const { return !chunk.has_value(); }
};

/// src/shared/mod.rs:17-23
struct FeatureSelection {
    enum class Kind { HighlyVariableCol, HighlyVariable, Randomized, VarianceThreshold, None } kind = Kind::None;
    std::string column;  // HighlyVariableCol
    size_t n = 0;        // HighlyVariable / Randomized
    double threshold = 0.0;
    static FeatureSelection HighlyVariableCol(std::string name) { return {Kind::HighlyVariableCol, std::move(name), 0, 0.0}; }
    static FeatureSelection HighlyVariable(size_t n) { return {Kind::HighlyVariable, {}, n, 0.0}; }
    static FeatureSelection Randomized(size_t n) { return {Kind::Randomized, {}, n, 0.0}; }
    static FeatureSelection VarianceThreshold(double t) { return {Kind::VarianceThreshold, {}, 0, t}; }
    static FeatureSelection None() { return {}; }
};

/// src/shared/mod.rs:62-102
struct FlexValue {
    enum class Kind { Absolute, Relative, None } kind = Kind::None;
    uint32_t absolute = 0;
    double relative = 0.0;
    static FlexValue Absolute(uint32_t v) { return {Kind::Absolute, v, 0.0}; }
    static FlexValue Relative(double v) { return {Kind::Relative, 0, v}; }
    static FlexValue None() { return {}; }
    bool is_absolute() const { return kind == Kind::Absolute; }
    bool is_relative() const { return kind == Kind::Relative; }
    bool is_none() const { return kind == Kind::None; }
};

}  // namespace shared

// =====================================================================================================================
// host matrices in the reference's layout (nalgebra-sparse CsrMatrix / CscMatrix: usize offsets + indices)
// =====================================================================================================================
template <class T> struct dtype_of;
#define SRB_HPP_DTYPE(T, D) \
    template <> struct dtype_of<T> { static constexpr int32_t value = D; }
SRB_HPP_DTYPE(int8_t, SRB_I8);
SRB_HPP_DTYPE(int16_t, SRB_I16);
SRB_HPP_DTYPE(int32_t, SRB_I32);
SRB_HPP_DTYPE(int64_t, SRB_I64);  // accepted by the type system like DynCsrMatrix::I64, refused at run time (shared/mod.rs:117)
SRB_HPP_DTYPE(uint8_t, SRB_U8);
SRB_HPP_DTYPE(uint16_t, SRB_U16);
SRB_HPP_DTYPE(uint32_t, SRB_U32);
SRB_HPP_DTYPE(uint64_t, SRB_U64);
SRB_HPP_DTYPE(float, SRB_F32);
SRB_HPP_DTYPE(double, SRB_F64);
#undef SRB_HPP_DTYPE

enum class Format : int32_t { Csr = SRB_CSR, Csc = SRB_CSC };

/// Borrowed view of a compressed matrix: what `row_offsets()/col_indices()/values()` hand out (csr.rs:24,32,96).
template <class T>
struct CsView {
    Format format = Format::Csr;
    uint64_t nrows = 0, ncols = 0;
    const uint64_t *offsets = nullptr;  // nmajor + 1
    const uint64_t *indices = nullptr;  // nnz
    const T *values = nullptr;          // nnz
    uint64_t nmajor() const { return format == Format::Csr ? nrows : ncols; }
    uint64_t nnz() const { return offsets ? offsets[nmajor()] : 0; }
};

/// Owning twin (what download() returns).
template <class T>
struct CsMatrix {
    Format format = Format::Csr;
    uint64_t nrows = 0, ncols = 0;
    std::vector<uint64_t> offsets, indices;
    std::vector<T> values;
    CsView<T> view() const { return {format, nrows, ncols, offsets.data(), indices.data(), values.data()}; }
};

/// Row-major dense block (ndarray Array2<f64>): obsm["X_pca"], varm["PCA_loadings"].
struct Array2 {
    size_t rows = 0, cols = 0;
    std::vector<double> data;
    double &operator()(size_t r, size_t c) { return data[r * cols + c]; }
    double operator()(size_t r, size_t c) const { return data[r * cols + c]; }
};

// =====================================================================================================================
// device handles (RAII over srb_ctx / srb_mat)
// =====================================================================================================================
class Device {
    srb_ctx *h_ = nullptr;

public:
    explicit Device(int32_t index = 0, srb_value_mode mode = SRB_VALUES_COMPACT) {
        detail::check(srb_ctx_create(index, &h_));
        if (mode != SRB_VALUES_COMPACT) detail::check(srb_ctx_set_value_mode(h_, mode));
    }
    ~Device() { srb_ctx_destroy(h_); }
    Device(const Device &) = delete;
    Device &operator=(const Device &) = delete;
    srb_ctx *handle() const { return h_; }
    void set_upload_mode(srb_upload_mode m) { detail::check(srb_ctx_set_upload_mode(h_, m)); }
    void set_eig_mode(srb_eig_mode m) { detail::check(srb_ctx_set_eig_mode(h_, m)); }
    /// Multi-GPU (new; the reference is single-process): one Device per rank, X sharded by cell-row
    /// (DeviceMatrix::set_shard). `id` comes from unique_id() on rank 0 and reaches the other ranks through the launcher.
    struct CommId {
        unsigned char bytes[128];
    };
    static CommId unique_id() {
        CommId id{};
        detail::check(srb_comm_unique_id(id.bytes));
        return id;
    }
    void comm_init(const CommId &id, int32_t rank, int32_t nranks) { detail::check(srb_ctx_comm_init(h_, id.bytes, rank, nranks)); }
    void synchronize() { detail::check(srb_ctx_synchronize(h_)); }
};

struct PcaResult {
    Array2 scores;      // local rows x k           (transform; obsm["X_pca"])
    Array2 components;  // n_selected x k           (V[:, :k], rows in selection order)
    std::vector<double> explained_variance_ratio;  // k
};

class DeviceMatrix {
    srb_mat *h_ = nullptr;

public:
    DeviceMatrix() = default;
    explicit DeviceMatrix(srb_mat *h) : h_(h) {}
    ~DeviceMatrix() { srb_mat_free(h_); }
    DeviceMatrix(DeviceMatrix &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    DeviceMatrix &operator=(DeviceMatrix &&o) noexcept {
        if (this != &o) {
            srb_mat_free(h_);
            h_ = o.h_, o.h_ = nullptr;
        }
        return *this;
    }
    DeviceMatrix(const DeviceMatrix &) = delete;
    DeviceMatrix &operator=(const DeviceMatrix &) = delete;
    srb_mat *handle() const { return h_; }

    template <class T>
    static DeviceMatrix upload(Device &dev, const CsView<T> &m) {
        srb_mat *h = nullptr;
        detail::check(srb_mat_upload(dev.handle(), (int32_t)m.format, m.nrows, m.ncols, m.nnz(), m.offsets, m.indices,
                                     SRB_IDX64, m.values, dtype_of<T>::value, &h));
        return DeviceMatrix(h);
    }

    struct Info {
        uint64_t nrows, ncols, nnz;
        Format format;
        int32_t value_dtype;
    };
    Info info() const {
        Info i{};
        int32_t f = 0;
        detail::check(srb_mat_info(h_, &i.nrows, &i.ncols, &i.nnz, &f, &i.value_dtype));
        i.format = (Format)f;
        return i;
    }
    uint64_t len(shared::Direction d) const {
        const Info i = info();
        return shared::is_row(d) ? i.nrows : i.ncols;
    }

    /// this rank's rows are [global_row0, global_row0 + nrows) of a matrix of global_nrows cells: per-gene results are
    /// then allreduced over the ranks of the Device's communicator, per-cell results stay local
    void set_shard(uint64_t global_row0, uint64_t global_nrows) { detail::check(srb_mat_set_shard(h_, global_row0, global_nrows)); }
    DeviceMatrix clone() const {  // IMAnnData::deep_clone of X: copy-on-write on the device
        srb_mat *h = nullptr;
        detail::check(srb_mat_clone(h_, &h));
        return DeviceMatrix(h);
    }
    DeviceMatrix subset(const std::vector<uint8_t> *keep_rows, const std::vector<uint8_t> *keep_cols) const {
        srb_mat *h = nullptr;
        detail::check(srb_mat_subset(h_, keep_rows ? keep_rows->data() : nullptr, keep_cols ? keep_cols->data() : nullptr, &h));
        return DeviceMatrix(h);
    }
    /// current values as f64 — what the reference holds after normalise (scale/mod.rs:82)
    CsMatrix<double> download() const {
        const Info i = info();
        CsMatrix<double> m;
        m.format = i.format, m.nrows = i.nrows, m.ncols = i.ncols;
        m.offsets.resize((i.format == Format::Csr ? i.nrows : i.ncols) + 1);
        m.indices.resize(i.nnz), m.values.resize(i.nnz);
        detail::check(srb_mat_download(h_, m.offsets.data(), m.indices.data(), m.values.data(), nullptr));
        return m;
    }

    // shared::statistics::{number,sum,variance,stddev,minmax}::whole — src/shared/statistics/mod.rs:7,49,91,104,117
    std::vector<uint32_t> number(shared::Direction d) const {
        std::vector<uint32_t> out(len(d));
        detail::check(srb_number(h_, (int32_t)d, out.data()));
        return out;
    }
    std::vector<double> sum(shared::Direction d) const {
        std::vector<double> out(len(d));
        detail::check(srb_sum(h_, (int32_t)d, out.data()));
        return out;
    }
    std::vector<double> variance(shared::Direction d) const {
        std::vector<double> out(len(d));
        detail::check(srb_variance(h_, (int32_t)d, out.data()));
        return out;
    }
    std::vector<double> std_dev(shared::Direction d) const {
        std::vector<double> out(len(d));
        detail::check(srb_std_dev(h_, (int32_t)d, out.data()));
        return out;
    }
    std::pair<std::vector<double>, std::vector<double>> min_max(shared::Direction d) const {
        std::vector<double> mn(len(d)), mx(len(d));
        detail::check(srb_min_max(h_, (int32_t)d, mn.data(), mx.data()));
        return {std::move(mn), std::move(mx)};
    }

    /// count, sum and sum of squares per gene of the current values of this CSR (chunk): pass 1 of the out-of-core pipeline
    struct GeneMoments {
        std::vector<double> count, sum, sumsq;
    };
    GeneMoments gene_moments() const {
        const size_t m = (size_t)info().ncols;
        GeneMoments g{std::vector<double>(m), std::vector<double>(m), std::vector<double>(m)};
        detail::check(srb_gene_moments(h_, g.count.data(), g.sum.data(), g.sumsq.data()));
        return g;
    }

    void normalize_total_inplace(double target_sum, shared::Direction d) {
        detail::check(srb_normalize_total_inplace(h_, target_sum, (int32_t)d));
    }
    void log1p_inplace() { detail::check(srb_log1p_inplace(h_)); }

    std::vector<uint64_t> select_hvg(uint64_t n_top) const {
        std::vector<uint64_t> out(std::min<uint64_t>(n_top, info().ncols));
        uint64_t n = 0;
        detail::check(srb_select_hvg(h_, n_top, out.data(), &n));
        out.resize(n);
        return out;
    }
    std::vector<uint64_t> select_var_threshold(double t) const {
        std::vector<uint64_t> out(info().ncols);
        uint64_t n = 0;
        detail::check(srb_select_var_threshold(h_, t, out.data(), &n));
        out.resize(n);
        return out;
    }
    /// convert_to_array_f64_selected (shared/mod.rs:292-315): all rows x selected columns
    Array2 densify_selected(const std::vector<uint64_t> &cols) const {
        Array2 a{(size_t)info().nrows, cols.size(), {}};
        a.data.resize(a.rows * a.cols);
        detail::check(srb_densify_selected(h_, cols.data(), cols.size(), a.data.data()));
        return a;
    }
    PcaResult pca(const std::vector<uint64_t> &sel, uint64_t k, bool center, bool scale, int32_t gram_mode = 0) const {
        PcaResult r;
        const size_t n = (size_t)info().nrows;
        r.scores = {n, (size_t)k, std::vector<double>(n * k)};
        r.components = {sel.size(), (size_t)k, std::vector<double>(sel.size() * k)};
        r.explained_variance_ratio.resize(k);
        detail::check(srb_pca(h_, sel.data(), sel.size(), k, center, scale, gram_mode, r.scores.data.data(),
                              r.components.data.data(), r.explained_variance_ratio.data()));
        return r;
    }
};

// =====================================================================================================================
// IMAnnData — the slice of anndata_memory::IMAnnData the path touches: X on the device, obs / var columns, obsm / varm
// =====================================================================================================================
using Column = std::variant<std::vector<uint32_t>, std::vector<double>, std::vector<uint8_t> /* bool */>;

namespace detail {
template <class V>
std::vector<V> take(const std::vector<V> &v, const std::vector<uint8_t> &mask) {
    std::vector<V> out;
    for (size_t i = 0; i < v.size() && i < mask.size(); ++i)
        if (mask[i]) out.push_back(v[i]);
    return out;
}
inline Column take(const Column &c, const std::vector<uint8_t> &mask) {
    return std::visit([&](const auto &v) -> Column { return take(v, mask); }, c);
}
inline Array2 take_rows(const Array2 &a, const std::vector<uint8_t> &mask) {
    Array2 out{0, a.cols, {}};
    for (size_t r = 0; r < a.rows; ++r)
        if (mask[r]) {
            out.data.insert(out.data.end(), a.data.begin() + r * a.cols, a.data.begin() + (r + 1) * a.cols);
            ++out.rows;
        }
    return out;
}
}  // namespace detail

class IMAnnData {
    DeviceMatrix x_;

public:
    std::map<std::string, Column> obs, var;
    std::map<std::string, Array2> obsm, varm;
    std::vector<double> explained_variance_ratio;  // computed and dropped by the reference (dim_red/mod.rs:77-88); kept

    explicit IMAnnData(DeviceMatrix x) : x_(std::move(x)) {}
    template <class T>
    IMAnnData(Device &dev, const CsView<T> &m) : x_(DeviceMatrix::upload(dev, m)) {}

    DeviceMatrix &x() { return x_; }
    const DeviceMatrix &x() const { return x_; }
    uint64_t n_obs() const { return x_.info().nrows; }
    uint64_t n_vars() const { return x_.info().ncols; }

    IMAnnData deep_clone() const {
        IMAnnData c(x_.clone());
        c.obs = obs, c.var = var, c.obsm = obsm, c.varm = varm, c.explained_variance_ratio = explained_variance_ratio;
        return c;
    }
    /// IMAnnData::subset / subset_inplace by boolean masks (processing/mod.rs:113-118, 140-145)
    IMAnnData subset(const std::vector<uint8_t> *keep_obs, const std::vector<uint8_t> *keep_var) const {
        IMAnnData out(x_.subset(keep_obs, keep_var));
        for (const auto &kv : obs) out.obs[kv.first] = keep_obs ? detail::take(kv.second, *keep_obs) : kv.second;
        for (const auto &kv : var) out.var[kv.first] = keep_var ? detail::take(kv.second, *keep_var) : kv.second;
        for (const auto &kv : obsm) out.obsm[kv.first] = keep_obs ? detail::take_rows(kv.second, *keep_obs) : kv.second;
        for (const auto &kv : varm) out.varm[kv.first] = keep_var ? detail::take_rows(kv.second, *keep_var) : kv.second;
        return out;
    }
    void subset_inplace(const std::vector<uint8_t> *keep_obs, const std::vector<uint8_t> *keep_var) {
        *this = subset(keep_obs, keep_var);
    }
};

// =====================================================================================================================
// memory::statistics — src/memory/statistics/mod.rs
// =====================================================================================================================
namespace memory {
namespace statistics {
using shared::Direction;

inline std::vector<uint32_t> compute_number(const IMAnnData &adata, Direction direction) { return adata.x().number(direction); }   // :10-15
inline std::vector<double> compute_sum(const IMAnnData &adata, Direction direction) { return adata.x().sum(direction); }           // :17-22
inline std::vector<double> compute_variance(const IMAnnData &adata, Direction direction) { return adata.x().variance(direction); } // :24-29
inline std::pair<std::vector<double>, std::vector<double>> compute_min_max(const IMAnnData &adata, Direction direction) {            // :31-39
    return adata.x().min_max(direction);
}
inline std::vector<double> compute_std_dev(const IMAnnData &adata, Direction direction) { return adata.x().std_dev(direction); }   // :41-46

/// src/memory/statistics/structs/mod.rs:1-10 (field names kept)
struct StatisticsContainer {
    std::vector<uint32_t> num_per_cell, num_per_gene;
    std::vector<double> expr_per_gene, expr_per_cell, variance_per_gene, variance_per_cell, std_dev_per_cell, std_dev_per_gene;
};

/// :48-72 — one ABI call, two passes over the matrix instead of the reference's sixteen
inline StatisticsContainer compute_qc_variables(const IMAnnData &adata) {
    const auto i = adata.x().info();
    StatisticsContainer s;
    s.num_per_cell.resize(i.nrows), s.expr_per_cell.resize(i.nrows), s.variance_per_cell.resize(i.nrows), s.std_dev_per_cell.resize(i.nrows);
    s.num_per_gene.resize(i.ncols), s.expr_per_gene.resize(i.ncols), s.variance_per_gene.resize(i.ncols), s.std_dev_per_gene.resize(i.ncols);
    detail::check(srb_qc_all(adata.x().handle(), s.num_per_cell.data(), s.num_per_gene.data(), s.expr_per_cell.data(),
                             s.expr_per_gene.data(), s.variance_per_cell.data(), s.variance_per_gene.data(),
                             s.std_dev_per_cell.data(), s.std_dev_per_gene.data()));
    return s;
}

/// :74-103 — column names exactly as the reference writes them
inline void qc_vars_inplace(IMAnnData &adata) {
    StatisticsContainer d = compute_qc_variables(adata);
    adata.obs["num_genes_per_cell"] = std::move(d.num_per_cell);
    adata.obs["sum_expr_per_cell"] = std::move(d.expr_per_cell);
    adata.obs["var_expr_per_cell"] = std::move(d.variance_per_cell);
    adata.obs["std_dev_per_cell"] = std::move(d.std_dev_per_cell);
    adata.var["num_cells_per_gene"] = std::move(d.num_per_gene);
    adata.var["sum_expr_per_gene"] = std::move(d.expr_per_gene);
    adata.var["var_expr_per_gene"] = std::move(d.variance_per_gene);
    adata.var["std_dev_per_gene"] = std::move(d.std_dev_per_gene);
}
}  // namespace statistics

// =====================================================================================================================
// memory::processing — src/memory/processing/{mod.rs, scale, transform, dim_red}
// =====================================================================================================================
namespace processing {
using shared::Direction;
using shared::FeatureSelection;
using shared::FlexValue;

/// ndarray_stats::interpolate::Linear on the sorted values: index q (n - 1), lower + (higher - lower) * frac
inline double linear_quantile(std::vector<double> v, double q) {
    if (v.empty()) throw Error(SRB_ERR_INVALID_ARG, "Error calculating percentile: empty input");
    if (!(q >= 0.0 && q <= 1.0)) throw Error(SRB_ERR_INVALID_ARG, "Error calculating percentile: q outside [0, 1]");
    for (double x : v)
        if (std::isnan(x)) throw Error(SRB_ERR_NAN, "NaN in the per-line sums (noisy_float n64 panics in the reference)");
    std::sort(v.begin(), v.end());
    const double pos = q * (double)(v.size() - 1);
    const size_t lo = (size_t)std::floor(pos), hi = (size_t)std::ceil(pos);
    return v[lo] + (v[hi] - v[lo]) * (pos - (double)lo);
}

/// processing/mod.rs:148-174: f64::MIN / f64::MAX when the limit is not Relative
inline std::pair<double, double> calculate_percentiles(const std::vector<double> &values, const FlexValue &lower_lim, const FlexValue &upper_lim) {
    const double lo = lower_lim.is_relative() ? linear_quantile(values, lower_lim.relative) : std::numeric_limits<double>::lowest();
    const double hi = upper_lim.is_relative() ? linear_quantile(values, upper_lim.relative) : std::numeric_limits<double>::max();
    return {lo, hi};
}

/// processing/mod.rs:32-83 (cells) / :193-243 (genes): Absolute limits compare the stored-entry COUNT, Relative limits
/// compare the SUM against its percentile; the nine (lower, upper) arms reduce to two independent conjuncts.
inline std::vector<uint8_t> create_filter_mask(size_t n, const std::vector<uint32_t> &counts, const std::vector<double> &sums,
                                               const FlexValue &lower_lim, const FlexValue &upper_lim, double lower_percentile,
                                               double upper_percentile) {
    std::vector<uint8_t> keep(n, 1);
    for (size_t i = 0; i < n; ++i) {
        bool k = true;
        if (lower_lim.is_absolute()) k = k && counts[i] >= lower_lim.absolute;
        else if (lower_lim.is_relative()) k = k && sums[i] >= lower_percentile;
        if (upper_lim.is_absolute()) k = k && counts[i] <= upper_lim.absolute;
        else if (upper_lim.is_relative()) k = k && sums[i] <= upper_percentile;
        keep[i] = k;
    }
    return keep;
}

namespace detail_filter {
inline std::vector<uint8_t> mask_for(const IMAnnData &adata, const FlexValue &lower_lim, const FlexValue &upper_lim, Direction d) {
    const bool need_count = lower_lim.is_absolute() || upper_lim.is_absolute();
    std::vector<uint32_t> counts;  // calculate_{cell,gene}_stats, processing/mod.rs:16-30, 176-191
    if (need_count) counts = adata.x().number(d);
    const std::vector<double> sums = adata.x().sum(d);
    const auto [lo, hi] = calculate_percentiles(sums, lower_lim, upper_lim);
    return create_filter_mask(sums.size(), counts, sums, lower_lim, upper_lim, lo, hi);
}
}  // namespace detail_filter

inline void filter_cells_inplace(IMAnnData &adata, FlexValue lower_lim, FlexValue upper_lim) {  // :86-121
    const auto mask = detail_filter::mask_for(adata, lower_lim, upper_lim, Direction::Row);
    adata.subset_inplace(&mask, nullptr);
}
inline IMAnnData filter_cells(const IMAnnData &adata, FlexValue lower_lim, FlexValue upper_lim) {  // :123-146
    const auto mask = detail_filter::mask_for(adata, lower_lim, upper_lim, Direction::Row);
    return adata.subset(&mask, nullptr);
}
inline void filter_genes_inplace(IMAnnData &adata, FlexValue lower_lim, FlexValue upper_lim) {  // :245-271
    const auto mask = detail_filter::mask_for(adata, lower_lim, upper_lim, Direction::Column);
    adata.subset_inplace(nullptr, &mask);
}
inline IMAnnData filter_genes(const IMAnnData &adata, FlexValue lower_lim, FlexValue upper_lim) {  // :273-299
    const auto mask = detail_filter::mask_for(adata, lower_lim, upper_lim, Direction::Column);
    return adata.subset(nullptr, &mask);
}

/// processing/mod.rs:303-312 -> scale::scale_row / scale_col (scale/mod.rs:7-173)
inline void normalize_total_inplace(IMAnnData &adata, double target_sum, Direction direction) {
    adata.x().normalize_total_inplace(target_sum, direction);
}
/// :314-322 — deep_clone + in-place
inline IMAnnData normalize_total(const IMAnnData &adata, double target_sum, Direction direction) {
    IMAnnData n = adata.deep_clone();
    normalize_total_inplace(n, target_sum, direction);
    return n;
}
/// :324-326 -> transform::log1p_data (transform/mod.rs:8-62)
inline void log1p_transform_inplace(IMAnnData &adata) { adata.x().log1p_inplace(); }
/// :328-332
inline IMAnnData log1p_transform(const IMAnnData &adata) {
    IMAnnData n = adata.deep_clone();
    log1p_transform_inplace(n);
    return n;
}

/// dim_red/mod.rs:123-156
inline std::vector<uint64_t> select_features(const IMAnnData &adata, const FeatureSelection &fs) {
    using K = FeatureSelection::Kind;
    switch (fs.kind) {
        case K::HighlyVariableCol: {
            auto it = adata.var.find(fs.column);
            if (it == adata.var.end()) throw Error(SRB_ERR_INVALID_ARG, "Error accessing column '" + fs.column + "'");
            const auto *b = std::get_if<std::vector<uint8_t>>(&it->second);
            if (!b) throw Error(SRB_ERR_INVALID_ARG, "Column '" + fs.column + "' is not boolean");
            std::vector<uint64_t> out;
            for (size_t i = 0; i < b->size(); ++i)
                if ((*b)[i]) out.push_back(i);
            return out;
        }
        case K::HighlyVariable: return adata.x().select_hvg(fs.n);
        case K::Randomized: {  // thread_rng in the reference: not reproducible there either
            std::vector<uint64_t> idx(adata.n_vars());
            std::iota(idx.begin(), idx.end(), 0);
            std::mt19937_64 rng{std::random_device{}()};
            std::shuffle(idx.begin(), idx.end(), rng);
            idx.resize(std::min<size_t>(fs.n, idx.size()));
            return idx;
        }
        case K::VarianceThreshold: return adata.x().select_var_threshold(fs.threshold);
        case K::None: {
            std::vector<uint64_t> idx(adata.n_vars());
            std::iota(idx.begin(), idx.end(), 0);
            return idx;
        }
    }
    throw Error(SRB_ERR_INVALID_ARG, "unknown FeatureSelection");
}

/// The SVD back-end argument of the reference (`svd_mode: S`, FaerSVD / LapackSVD from single_algebra): accepted and
/// ignored — the SVD is replaced by the equivalent Gram + symmetric eigendecomposition on the device.
enum class SVDMode { Lapack, Faer };

/// dim_red/mod.rs:24-94. Defaults as the reference: n_components 2 (capped at #features), center / scale true.
/// `n_threads` (rayon pool) is accepted and ignored. Stores obsm["X_pca"]; the loadings and the explained-variance
/// ratio, which the reference computes and drops (:77-88), are kept in varm["PCA_loadings"] (zero-filled to all genes
/// like attach_pca_results :108-118) and adata.explained_variance_ratio.
inline void pca_inplace(IMAnnData &adata, std::optional<size_t> n_components, std::optional<bool> center, std::optional<bool> scale,
                        std::optional<size_t> n_threads, const FeatureSelection &feature_selection, SVDMode svd_mode = SVDMode::Lapack,
                        int32_t gram_mode = 0) {
    (void)n_threads, (void)svd_mode;
    const std::vector<uint64_t> sel = select_features(adata, feature_selection);
    // dense.column(1) panics in the reference when fewer than two features are selected (:38-39)
    if (sel.size() < 2) throw Error(SRB_ERR_INVALID_ARG, "pca_inplace needs at least two selected features (the reference panics here)");
    const size_t k = std::min<size_t>(n_components.value_or(2), sel.size());
    PcaResult r = adata.x().pca(sel, k, center.value_or(true), scale.value_or(true), gram_mode);
    Array2 full{(size_t)adata.n_vars(), k, {}};
    full.data.assign(full.rows * k, 0.0);
    for (size_t j = 0; j < sel.size(); ++j)
        std::copy_n(r.components.data.begin() + j * k, k, full.data.begin() + (size_t)sel[j] * k);
    adata.obsm["X_pca"] = std::move(r.scores);
    adata.varm["PCA_loadings"] = std::move(full);
    adata.explained_variance_ratio = std::move(r.explained_variance_ratio);
}

}  // namespace processing
}  // namespace memory

// =====================================================================================================================
// backed — src/backed/statistics/mod.rs (+ the pipeline the reference's empty src/backed/processing lacks)
// =====================================================================================================================
namespace backed {
using shared::ComputationMode;
using shared::Direction;

/// Stand-in for anndata::AnnData<B: Backend>: X is reachable as a whole or through the chunk iterator
/// (ArrayElemOp::iter, src/shared/statistics/mod.rs:24,66): row chunks for CSR, column chunks for CSC, in order.
template <class T>
class ChunkSource {
public:
    virtual ~ChunkSource() = default;
    virtual uint64_t n_obs() const = 0;
    virtual uint64_t n_vars() const = 0;
    virtual Format format() const = 0;
    /// calls f(chunk) for consecutive chunks of at most chunk_size major lines; offsets of a chunk start at 0
    virtual void for_each_chunk(size_t chunk_size, const std::function<void(const CsView<T> &)> &f) const = 0;
    virtual CsView<T> whole() const = 0;
};

/// A ChunkSource over a host matrix (what an HDF5-backed store would produce chunk by chunk).
template <class T>
class HostChunkSource : public ChunkSource<T> {
    CsView<T> m_;

public:
    explicit HostChunkSource(CsView<T> m) : m_(m) {}
    uint64_t n_obs() const override { return m_.nrows; }
    uint64_t n_vars() const override { return m_.ncols; }
    Format format() const override { return m_.format; }
    CsView<T> whole() const override { return m_; }
    void for_each_chunk(size_t chunk_size, const std::function<void(const CsView<T> &)> &f) const override {
        if (chunk_size == 0) throw Error(SRB_ERR_INVALID_ARG, "chunk size must be positive");
        const uint64_t nmajor = m_.nmajor();
        std::vector<uint64_t> off;
        for (uint64_t s = 0; s < nmajor; s += chunk_size) {
            const uint64_t e = std::min<uint64_t>(nmajor, s + chunk_size), base = m_.offsets[s];
            off.resize(e - s + 1);
            for (uint64_t i = s; i <= e; ++i) off[i - s] = m_.offsets[i] - base;
            CsView<T> c{m_.format, m_.format == Format::Csr ? e - s : m_.nrows, m_.format == Format::Csr ? m_.ncols : e - s,
                        off.data(), m_.indices + base, m_.values + base};
            f(c);
        }
    }
};

// ---- on-disk chunk store -----------------------------------------------------------------------------------------------
// A directory with meta.txt ("csr|csc nrows ncols"), indptr.npy, indices.npy, data.npy — the three arrays an .h5ad X group
// holds (there is no HDF5 library in this image). Written by singlerust_b200.anndata.BackedAnnData.write_store or
// write_store() below; chunks are read with fseek/fread only when the iterator reaches them.
namespace store_detail {
struct NpyInfo {
    std::string descr;     // e.g. "<f4", "<u8"
    uint64_t count = 0;    // elements (1-D arrays only)
    long long offset = 0;  // byte offset of the data
};
inline NpyInfo npy_header(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error(SRB_ERR_INVALID_ARG, "cannot open " + path);
    unsigned char pre[12];
    NpyInfo info;
    bool ok = std::fread(pre, 1, 10, f) == 10 && std::memcmp(pre, "\x93NUMPY", 6) == 0;
    size_t hlen = 0;
    if (ok && pre[6] == 1) {
        hlen = pre[8] | (size_t)pre[9] << 8;
        info.offset = 10;
    } else if (ok && (pre[6] == 2 || pre[6] == 3)) {
        ok = std::fread(pre + 10, 1, 2, f) == 2;
        hlen = pre[8] | (size_t)pre[9] << 8 | (size_t)pre[10] << 16 | (size_t)pre[11] << 24;
        info.offset = 12;
    } else {
        ok = false;
    }
    std::string h(hlen, ' ');
    ok = ok && std::fread(&h[0], 1, hlen, f) == hlen;
    std::fclose(f);
    if (!ok) throw Error(SRB_ERR_INVALID_ARG, path + ": not a .npy file");
    info.offset += (long long)hlen;
    const size_t d = h.find("'descr'"), fo = h.find("'fortran_order'"), sh = h.find("'shape'");
    if (d == std::string::npos || sh == std::string::npos) throw Error(SRB_ERR_INVALID_ARG, path + ": unreadable .npy header");
    const size_t q0 = h.find('\'', h.find(':', d)), q1 = h.find('\'', q0 + 1);
    info.descr = h.substr(q0 + 1, q1 - q0 - 1);
    if (fo != std::string::npos && h.compare(h.find(':', fo) + 1, 5, " True") == 0) throw Error(SRB_ERR_UNSUPPORTED, path + ": fortran order");
    const size_t p0 = h.find('(', sh), p1 = h.find(')', p0);
    const std::string shape = h.substr(p0 + 1, p1 - p0 - 1);
    if (std::count(shape.begin(), shape.end(), ',') > 1) throw Error(SRB_ERR_UNSUPPORTED, path + ": not one-dimensional");
    info.count = shape.find_first_of("0123456789") == std::string::npos ? 0 : std::stoull(shape);
    return info;
}
inline void read_at(const std::string &path, long long offset, void *dst, size_t bytes) {
    if (!bytes) return;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error(SRB_ERR_INVALID_ARG, "cannot open " + path);
    const bool ok = fseeko(f, (off_t)offset, SEEK_SET) == 0 && std::fread(dst, 1, bytes, f) == bytes;
    std::fclose(f);
    if (!ok) throw Error(SRB_ERR_INVALID_ARG, "short read from " + path);
}
template <class T> struct npy_descr;
#define SRB_HPP_DESCR(T, S) \
    template <> struct npy_descr<T> { static const char *value() { return S; } }
SRB_HPP_DESCR(int8_t, "|i1");
SRB_HPP_DESCR(uint8_t, "|u1");
SRB_HPP_DESCR(int16_t, "<i2");
SRB_HPP_DESCR(uint16_t, "<u2");
SRB_HPP_DESCR(int32_t, "<i4");
SRB_HPP_DESCR(uint32_t, "<u4");
SRB_HPP_DESCR(int64_t, "<i8");
SRB_HPP_DESCR(uint64_t, "<u8");
SRB_HPP_DESCR(float, "<f4");
SRB_HPP_DESCR(double, "<f8");
#undef SRB_HPP_DESCR
template <class T>
void write_npy(const std::string &path, const T *data, uint64_t n) {
    std::string h = std::string("{'descr': '") + npy_descr<T>::value() + "', 'fortran_order': False, 'shape': (" + std::to_string(n) + ",), }";
    while ((10 + h.size() + 1) % 64) h.push_back(' ');
    h.push_back('\n');
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw Error(SRB_ERR_INVALID_ARG, "cannot create " + path);
    const unsigned char pre[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(h.size() & 255), (unsigned char)(h.size() >> 8)};
    const bool ok = std::fwrite(pre, 1, 10, f) == 10 && std::fwrite(h.data(), 1, h.size(), f) == h.size() &&
                    (n == 0 || std::fwrite(data, sizeof(T), n, f) == n);
    std::fclose(f);
    if (!ok) throw Error(SRB_ERR_INVALID_ARG, "short write to " + path);
}
}  // namespace store_detail

template <class T>
void write_store(const std::string &dir, const CsView<T> &m) {
    FILE *f = std::fopen((dir + "/meta.txt").c_str(), "w");
    if (!f) throw Error(SRB_ERR_INVALID_ARG, "cannot create " + dir + "/meta.txt (the directory must exist)");
    std::fprintf(f, "%s %llu %llu\n", m.format == Format::Csr ? "csr" : "csc", (unsigned long long)m.nrows, (unsigned long long)m.ncols);
    std::fclose(f);
    store_detail::write_npy(dir + "/indptr.npy", m.offsets, m.nmajor() + 1);
    store_detail::write_npy(dir + "/indices.npy", m.indices, m.nnz());
    store_detail::write_npy(dir + "/data.npy", m.values, m.nnz());
}

/// ChunkSource over an on-disk chunk store: only indptr is held in memory; index / value chunks are read on demand.
template <class T>
class StoreChunkSource : public ChunkSource<T> {
    std::string dir_;
    Format format_ = Format::Csr;
    uint64_t nrows_ = 0, ncols_ = 0;
    std::vector<uint64_t> indptr_;
    store_detail::NpyInfo idx_, val_;
    mutable CsMatrix<T> whole_;  // filled by whole()

    void read_lines(uint64_t s, uint64_t e, std::vector<uint64_t> &off, std::vector<uint64_t> &idx, std::vector<T> &val) const {
        const uint64_t a = indptr_[s], b = indptr_[e], n = b - a;
        off.resize(e - s + 1);
        for (uint64_t i = s; i <= e; ++i) off[i - s] = indptr_[i] - a;
        idx.resize(n), val.resize(n);
        if (idx_.descr == "<u8" || idx_.descr == "<i8") {
            store_detail::read_at(dir_ + "/indices.npy", idx_.offset + (long long)(8 * a), idx.data(), 8 * n);
        } else {  // 4-byte on-disk indices (the h5ad convention) are widened to the usize layout
            std::vector<uint32_t> tmp(n);
            store_detail::read_at(dir_ + "/indices.npy", idx_.offset + (long long)(4 * a), tmp.data(), 4 * n);
            std::copy(tmp.begin(), tmp.end(), idx.begin());
        }
        store_detail::read_at(dir_ + "/data.npy", val_.offset + (long long)(sizeof(T) * a), val.data(), sizeof(T) * n);
    }

public:
    explicit StoreChunkSource(const std::string &dir) : dir_(dir) {
        FILE *f = std::fopen((dir + "/meta.txt").c_str(), "r");
        if (!f) throw Error(SRB_ERR_INVALID_ARG, "cannot open " + dir + "/meta.txt");
        char fmt[8] = {0};
        unsigned long long r = 0, c = 0;
        const int got = std::fscanf(f, "%7s %llu %llu", fmt, &r, &c);
        std::fclose(f);
        if (got != 3 || (std::strcmp(fmt, "csr") && std::strcmp(fmt, "csc"))) throw Error(SRB_ERR_INVALID_ARG, dir + "/meta.txt: expected 'csr|csc nrows ncols'");
        format_ = std::strcmp(fmt, "csr") ? Format::Csc : Format::Csr, nrows_ = r, ncols_ = c;
        const store_detail::NpyInfo ip = store_detail::npy_header(dir + "/indptr.npy");
        idx_ = store_detail::npy_header(dir + "/indices.npy"), val_ = store_detail::npy_header(dir + "/data.npy");
        const uint64_t nmajor = format_ == Format::Csr ? nrows_ : ncols_;
        if (ip.count != nmajor + 1 || idx_.count != val_.count) throw Error(SRB_ERR_INVALID_ARG, dir + ": arrays do not match meta.txt");
        if (val_.descr != store_detail::npy_descr<T>::value()) throw Error(SRB_ERR_UNSUPPORTED_DTYPE, dir + "/data.npy holds " + val_.descr + ", not " + store_detail::npy_descr<T>::value());
        if (idx_.descr != "<u8" && idx_.descr != "<i8" && idx_.descr != "<u4" && idx_.descr != "<i4") throw Error(SRB_ERR_UNSUPPORTED_DTYPE, dir + "/indices.npy: " + idx_.descr);
        indptr_.resize(nmajor + 1);
        if (ip.descr == "<u8" || ip.descr == "<i8") {
            store_detail::read_at(dir + "/indptr.npy", ip.offset, indptr_.data(), 8 * (nmajor + 1));
        } else if (ip.descr == "<u4" || ip.descr == "<i4") {
            std::vector<uint32_t> tmp(nmajor + 1);
            store_detail::read_at(dir + "/indptr.npy", ip.offset, tmp.data(), 4 * (nmajor + 1));
            std::copy(tmp.begin(), tmp.end(), indptr_.begin());
        } else {
            throw Error(SRB_ERR_UNSUPPORTED_DTYPE, dir + "/indptr.npy: " + ip.descr);
        }
        if (indptr_.front() != 0 || indptr_.back() != idx_.count) throw Error(SRB_ERR_INVALID_ARG, dir + ": indptr does not span the index array");
    }
    uint64_t n_obs() const override { return nrows_; }
    uint64_t n_vars() const override { return ncols_; }
    Format format() const override { return format_; }
    void for_each_chunk(size_t chunk_size, const std::function<void(const CsView<T> &)> &f) const override {
        if (chunk_size == 0) throw Error(SRB_ERR_INVALID_ARG, "chunk size must be positive");
        const uint64_t nmajor = indptr_.size() - 1;
        std::vector<uint64_t> off, idx;
        std::vector<T> val;
        for (uint64_t s = 0; s < nmajor; s += chunk_size) {
            const uint64_t e = std::min<uint64_t>(nmajor, s + chunk_size);
            read_lines(s, e, off, idx, val);
            f(CsView<T>{format_, format_ == Format::Csr ? e - s : nrows_, format_ == Format::Csr ? ncols_ : e - s, off.data(), idx.data(), val.data()});
        }
    }
    /// reads the whole matrix into memory (ComputationMode::Whole)
    CsView<T> whole() const override {
        if (whole_.offsets.empty()) {
            whole_.format = format_, whole_.nrows = nrows_, whole_.ncols = ncols_;
            read_lines(0, indptr_.size() - 1, whole_.offsets, whole_.indices, whole_.values);
        }
        return whole_.view();
    }
};

/// RAII over srb_stream: shared::statistics::{number,sum}::chunked (src/shared/statistics/mod.rs:17-41, 59-83)
class ChunkStream {
    srb_stream *h_ = nullptr;
    uint64_t nrows_, ncols_;

public:
    ChunkStream(Device &dev, Format fmt, uint64_t nrows_total, uint64_t ncols_total) : nrows_(nrows_total), ncols_(ncols_total) {
        detail::check(srb_stream_begin(dev.handle(), (int32_t)fmt, nrows_total, ncols_total, &h_));
    }
    ~ChunkStream() { srb_stream_free(h_); }
    ChunkStream(const ChunkStream &) = delete;
    ChunkStream &operator=(const ChunkStream &) = delete;
    void set_retain(uint64_t nnz_hint, bool keep_statistics) { detail::check(srb_stream_set_retain(h_, nnz_hint, keep_statistics)); }
    template <class T>
    void push(const CsView<T> &c) {
        detail::check(srb_stream_push(h_, c.nmajor(), c.nnz(), c.offsets, c.indices, SRB_IDX64, c.values, dtype_of<T>::value));
    }
    uint64_t len(Direction d) const { return shared::is_row(d) ? nrows_ : ncols_; }
    std::vector<uint32_t> number(Direction d) {
        std::vector<uint32_t> out(len(d));
        detail::check(srb_stream_number(h_, (int32_t)d, out.data()));
        return out;
    }
    std::vector<double> sum(Direction d) {
        std::vector<double> out(len(d));
        detail::check(srb_stream_sum(h_, (int32_t)d, out.data()));
        return out;
    }
    std::vector<double> variance(Direction d) {
        std::vector<double> out(len(d));
        detail::check(srb_stream_variance(h_, (int32_t)d, out.data()));
        return out;
    }
    DeviceMatrix finish_matrix() {
        srb_mat *m = nullptr;
        detail::check(srb_stream_finish_matrix(h_, &m));
        return DeviceMatrix(m);
    }
};

namespace statistics {
namespace detail_stream {
template <class T>
void feed(ChunkStream &st, const ChunkSource<T> &adata, size_t chunk) {
    adata.for_each_chunk(chunk, [&](const CsView<T> &c) { st.push(c); });
}
}  // namespace detail_stream

/// src/backed/statistics/mod.rs:5-24. Chunked Row-direction results land at the chunk's global offset (the reference
/// drops the offset, csr.rs:56-61 — documented deviation).
template <class T>
std::vector<uint32_t> compute_number(Device &dev, const ChunkSource<T> &adata, Direction direction, ComputationMode mode) {
    if (mode.is_whole()) return DeviceMatrix::upload(dev, adata.whole()).number(direction);
    ChunkStream st(dev, adata.format(), adata.n_obs(), adata.n_vars());
    detail_stream::feed(st, adata, *mode.chunk);
    return st.number(direction);
}
/// src/backed/statistics/mod.rs:26-45
template <class T>
std::vector<double> compute_sum(Device &dev, const ChunkSource<T> &adata, Direction direction, ComputationMode mode) {
    if (mode.is_whole()) return DeviceMatrix::upload(dev, adata.whole()).sum(direction);
    ChunkStream st(dev, adata.format(), adata.n_obs(), adata.n_vars());
    detail_stream::feed(st, adata, *mode.chunk);
    return st.sum(direction);
}
/// not in the reference (only number and sum exist for backed data): same accumulator, one more output
template <class T>
std::vector<double> compute_variance(Device &dev, const ChunkSource<T> &adata, Direction direction, ComputationMode mode) {
    if (mode.is_whole()) return DeviceMatrix::upload(dev, adata.whole()).variance(direction);
    ChunkStream st(dev, adata.format(), adata.n_obs(), adata.n_vars());
    detail_stream::feed(st, adata, *mode.chunk);
    return st.variance(direction);
}
}  // namespace statistics

/// RAII over srb_pca_stream: out-of-core PCA over CSR row chunks (push_gram every chunk, fit, transform every chunk)
class PcaStream {
    srb_pca_stream *h_ = nullptr;
    size_t n_sel_, k_;

public:
    PcaStream(Device &dev, uint64_t ncols, uint64_t ncells_total, const std::vector<double> &gene_sum, const std::vector<double> &gene_sumsq,
              const std::vector<uint64_t> &sel, size_t k, bool center, bool scale, int32_t gram_mode = 0)
        : n_sel_(sel.size()), k_(std::min(k, sel.size())) {
        if (gene_sum.size() != ncols || gene_sumsq.size() != ncols) throw Error(SRB_ERR_INVALID_ARG, "gene moments must have one entry per gene");
        detail::check(srb_pca_stream_begin(dev.handle(), ncols, ncells_total, gene_sum.data(), gene_sumsq.data(), sel.data(), sel.size(), k_,
                                           center, scale, gram_mode, &h_));
    }
    ~PcaStream() { srb_pca_stream_free(h_); }
    PcaStream(const PcaStream &) = delete;
    PcaStream &operator=(const PcaStream &) = delete;
    size_t k() const { return k_; }
    void push_gram(const DeviceMatrix &chunk) { detail::check(srb_pca_stream_push_gram(h_, chunk.handle())); }
    /// returns (components n_sel x k, explained-variance ratio k)
    std::pair<Array2, std::vector<double>> fit() {
        Array2 comps{n_sel_, k_, std::vector<double>(n_sel_ * k_)};
        std::vector<double> evr(k_);
        detail::check(srb_pca_stream_fit(h_, comps.data.data(), evr.data()));
        return {std::move(comps), std::move(evr)};
    }
    /// scores of the chunk's rows into out[0 .. rows * k)
    void transform(const DeviceMatrix &chunk, double *out) { detail::check(srb_pca_stream_transform(h_, chunk.handle(), out)); }
};

namespace processing {
/// select_features (dim_red/mod.rs:123-156, HighlyVariable arm) on per-gene moments accumulated over chunks: nonzero-only
/// one-pass variance (helper/csr.rs:172-186), stable descending sort, ties keep ascending index. O(genes) on the host,
/// like the reference's own sort.
inline std::vector<uint64_t> select_hvg_from_moments(const std::vector<double> &count, const std::vector<double> &sum,
                                                     const std::vector<double> &sumsq, size_t n_top) {
    const size_t m = count.size();
    std::vector<double> var(m, 0.0);
    for (size_t j = 0; j < m; ++j)
        if (count[j] > 0) {
            const double mean = sum[j] / count[j];
            var[j] = sumsq[j] / count[j] - mean * mean;
            if (std::isnan(var[j])) throw Error(SRB_ERR_NAN, "NaN variance in the HVG sort (the reference panics here)");
        }
    std::vector<uint64_t> idx(m);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return var[a] > var[b]; });
    idx.resize(std::min(n_top, m));
    return idx;
}

/// The headline pipeline for data that does NOT fit the GPU: three passes over the CSR row chunks, one chunk resident at
/// a time (srb_gene_moments, srb_pca_stream_*). New functionality: the reference's src/backed/processing/mod.rs is empty.
struct OutOfCoreResult {
    Array2 scores, components;
    std::vector<double> explained_variance_ratio;
    std::vector<uint64_t> selection;
};
template <class T>
OutOfCoreResult normalize_hvg_pca_out_of_core(Device &dev, const ChunkSource<T> &adata, ComputationMode mode, double target_sum = 1e4,
                                              size_t n_top_genes = 2000, size_t n_components = 50, bool center = true, bool scale = true,
                                              int32_t gram_mode = 0) {
    if (mode.is_whole() || adata.format() != Format::Csr) throw Error(SRB_ERR_INVALID_ARG, "the out-of-core pipeline streams CSR row chunks");
    const uint64_t n = adata.n_obs(), m = adata.n_vars();
    auto transformed = [&](const CsView<T> &c) {
        DeviceMatrix d = DeviceMatrix::upload(dev, c);
        d.normalize_total_inplace(target_sum, Direction::Row);  // row-local: a row chunk normalises like the whole matrix
        d.log1p_inplace();
        return d;
    };
    std::vector<double> cnt(m, 0.0), sum(m, 0.0), sq(m, 0.0);
    adata.for_each_chunk(*mode.chunk, [&](const CsView<T> &c) {  // pass 1
        const auto g = transformed(c).gene_moments();
        for (uint64_t j = 0; j < m; ++j) cnt[j] += g.count[j], sum[j] += g.sum[j], sq[j] += g.sumsq[j];
    });
    OutOfCoreResult r;
    r.selection = select_hvg_from_moments(cnt, sum, sq, n_top_genes);
    if (r.selection.size() < 2) throw Error(SRB_ERR_INVALID_ARG, "pca needs at least two selected features (the reference panics here)");
    PcaStream ps(dev, m, n, sum, sq, r.selection, n_components, center, scale, gram_mode);
    adata.for_each_chunk(*mode.chunk, [&](const CsView<T> &c) { ps.push_gram(transformed(c)); });  // pass 2
    auto fit = ps.fit();
    r.components = std::move(fit.first), r.explained_variance_ratio = std::move(fit.second);
    r.scores = Array2{(size_t)n, ps.k(), std::vector<double>((size_t)n * ps.k())};
    uint64_t row = 0;
    adata.for_each_chunk(*mode.chunk, [&](const CsView<T> &c) {  // pass 3
        ps.transform(transformed(c), r.scores.data.data() + row * ps.k());
        row += c.nrows;
    });
    return r;
}

/// Stream the backed X to the device chunk by chunk and keep it resident (srb_stream_set_retain): new functionality,
/// the reference's src/backed/processing/mod.rs is empty.
template <class T>
IMAnnData load_resident(Device &dev, const ChunkSource<T> &adata, ComputationMode mode, uint64_t nnz_hint = 0) {
    if (mode.is_whole()) return IMAnnData(DeviceMatrix::upload(dev, adata.whole()));
    ChunkStream st(dev, adata.format(), adata.n_obs(), adata.n_vars());
    st.set_retain(nnz_hint, false);
    statistics::detail_stream::feed(st, adata, *mode.chunk);
    return IMAnnData(st.finish_matrix());
}
/// normalize_total(Row) -> log1p -> pca_inplace(HighlyVariable(n)) over backed data (BASELINE.json config 5)
template <class T>
IMAnnData normalize_hvg_pca(Device &dev, const ChunkSource<T> &adata, ComputationMode mode, double target_sum = 1e4,
                            size_t n_top_genes = 2000, size_t n_components = 50, bool center = true, bool scale = true) {
    IMAnnData d = load_resident(dev, adata, mode);
    memory::processing::normalize_total_inplace(d, target_sum, Direction::Row);
    memory::processing::log1p_transform_inplace(d);
    memory::processing::pca_inplace(d, n_components, center, scale, std::nullopt, shared::FeatureSelection::HighlyVariable(n_top_genes));
    return d;
}
}  // namespace processing
}  // namespace backed

}  // namespace single_rust
#endif  // SINGLE_RUST_B200_HPP
