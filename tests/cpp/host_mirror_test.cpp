// host_mirror_test.cpp — parity test of the C++ host side (include/single_rust_b200.hpp) written the way the
// reference's own tests read (tests/test_basic_stats.rs, tests/test_basic_load.rs, processing/mod.rs:335-481):
// load a matrix, call memory::statistics / memory::processing / backed::statistics, compare.
//   host_mirror_test --cpu-check   no GPU needed: host logic + "fails loudly without a device"
//   host_mirror_test               needs the B200: known answers of SURVEY.md §9 and differential checks against the
//                                  CPU oracle (oracle/srb_oracle.c — test infrastructure, linked only here)
#include <cinttypes>
#include <cstdio>
#include <cstdlib>

#include "single_rust_b200.hpp"

using namespace single_rust;
using shared::ComputationMode;
using shared::Direction;
using shared::FeatureSelection;
using shared::FlexValue;

extern "C" {  // oracle/srb_oracle.c
void orc_number(uint64_t nmajor, uint64_t nminor, const uint64_t *offsets, const uint64_t *indices, int along_major, uint32_t *out);
void orc_sum_f32(uint64_t, uint64_t, const uint64_t *, const uint64_t *, const float *, int, double *);
void orc_variance_f32(uint64_t, uint64_t, const uint64_t *, const uint64_t *, const float *, int, double *);
void orc_std_dev_f32(uint64_t, uint64_t, const uint64_t *, const uint64_t *, const float *, int, double *);
void orc_min_max_f32(uint64_t, uint64_t, const uint64_t *, const uint64_t *, const float *, int, double *, double *);
void orc_normalize_total_f32(uint64_t, uint64_t, const uint64_t *, const uint64_t *, const float *, int, double, double *);
void orc_log1p_f64(uint64_t nnz, double *values);
void orc_variance_f64(uint64_t, uint64_t, const uint64_t *, const uint64_t *, const double *, int, double *);
int orc_select_hvg(const double *variances, uint64_t m, uint64_t n_top, uint64_t *out_idx);
}

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        ++g_checks;                                                          \
        if (!(cond)) {                                                       \
            ++g_fail;                                                        \
            fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
        }                                                                    \
    } while (0)

static bool same_or_both_nan(double a, double b) { return (std::isnan(a) && std::isnan(b)) || a == b; }
static bool close(double a, double b, double rtol, double atol = 0.0) {
    if (std::isnan(a) || std::isnan(b)) return std::isnan(a) && std::isnan(b);
    if (std::isinf(a) || std::isinf(b)) return a == b;
    return std::fabs(a - b) <= atol + rtol * std::fabs(b);
}
template <class A, class B>
static bool all_equal(const A &a, const B &b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (!same_or_both_nan((double)a[i], (double)b[i])) return false;
    return true;
}
static bool all_close(const std::vector<double> &a, const std::vector<double> &b, double rtol, double atol = 0.0) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i)
        if (!close(a[i], b[i], rtol, atol)) {
            fprintf(stderr, "  [%zu] got %.17g want %.17g\n", i, a[i], b[i]);
            return false;
        }
    return true;
}

// ---- the 4 x 5 known-answer matrix of SURVEY.md §9 --------------------------------------------------------------------
static const std::vector<uint64_t> kIndptr = {0, 3, 5, 5, 8}, kIndices = {0, 2, 3, 0, 1, 0, 2, 3};
static const std::vector<double> kData = {1, 2, 3, 4, 5, 2, 2, 6};
static const double NaN = std::numeric_limits<double>::quiet_NaN(), Inf = std::numeric_limits<double>::infinity();

// ---- a seeded random CSR in the reference's layout --------------------------------------------------------------------
static CsMatrix<float> random_csr(uint64_t n, uint64_t m, double density, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> u(0.0, 1.0);
    std::uniform_int_distribution<int> cnt(1, 49);
    CsMatrix<float> a;
    a.format = Format::Csr, a.nrows = n, a.ncols = m;
    a.offsets.push_back(0);
    for (uint64_t i = 0; i < n; ++i) {
        for (uint64_t j = 0; j < m; ++j)
            if (i != 1 && j != 0 && u(rng) < density) {  // row 1 and gene 0 stay empty
                a.indices.push_back(j);
                a.values.push_back((float)cnt(rng));
            }
        a.offsets.push_back(a.indices.size());
    }
    return a;
}

static int cpu_check(const char *store_dir) {
    // (1) no device: construction fails loudly, nothing falls back to the CPU
    bool threw = false;
    try {
        Device dev(0);
    } catch (const Error &e) {
        threw = true;
        CHECK(e.code == SRB_ERR_CUDA);
        CHECK(std::string(e.what()).find("no CPU fallback") != std::string::npos);
    }
    CHECK(threw);
    // (2) host logic of the filters: ndarray-stats Linear quantile and the nine FlexValue arms
    using namespace memory::processing;
    CHECK(linear_quantile({4, 1, 3, 2}, 0.0) == 1.0);
    CHECK(linear_quantile({4, 1, 3, 2}, 1.0) == 4.0);
    CHECK(linear_quantile({4, 1, 3, 2}, 0.5) == 2.5);
    CHECK(close(linear_quantile({4, 1, 3, 2}, 0.1), 1.3, 1e-15));
    auto pr = calculate_percentiles({1, 2, 3, 4, 5}, FlexValue::Relative(0.25), FlexValue::None());
    CHECK(pr.first == 2.0 && pr.second == std::numeric_limits<double>::max());
    pr = calculate_percentiles({1, 2, 3, 4, 5}, FlexValue::Absolute(3), FlexValue::Relative(0.75));
    CHECK(pr.first == std::numeric_limits<double>::lowest() && pr.second == 4.0);
    const std::vector<uint32_t> counts = {1, 5, 3, 9};
    const std::vector<double> sums = {10, 50, 30, 90};
    auto mk = [&](FlexValue lo, FlexValue hi, double lp, double up) { return create_filter_mask(4, counts, sums, lo, hi, lp, up); };
    CHECK((mk(FlexValue::Absolute(3), FlexValue::Absolute(5), 0, 0) == std::vector<uint8_t>{0, 1, 1, 0}));
    CHECK((mk(FlexValue::Relative(0.1), FlexValue::Relative(0.9), 30, 50) == std::vector<uint8_t>{0, 1, 1, 0}));
    CHECK((mk(FlexValue::Absolute(3), FlexValue::Relative(0.9), 0, 50) == std::vector<uint8_t>{0, 1, 1, 0}));
    CHECK((mk(FlexValue::Relative(0.1), FlexValue::Absolute(5), 30, 0) == std::vector<uint8_t>{0, 1, 1, 0}));
    CHECK((mk(FlexValue::Absolute(3), FlexValue::None(), 0, 0) == std::vector<uint8_t>{0, 1, 1, 1}));
    CHECK((mk(FlexValue::None(), FlexValue::Absolute(5), 0, 0) == std::vector<uint8_t>{1, 1, 1, 0}));
    CHECK((mk(FlexValue::Relative(0.1), FlexValue::None(), 30, 0) == std::vector<uint8_t>{0, 1, 1, 1}));
    CHECK((mk(FlexValue::None(), FlexValue::Relative(0.9), 0, 50) == std::vector<uint8_t>{1, 1, 1, 0}));
    CHECK((mk(FlexValue::None(), FlexValue::None(), 0, 0) == std::vector<uint8_t>{1, 1, 1, 1}));
    bool nan_threw = false;
    try {
        linear_quantile({1.0, NaN}, 0.5);
    } catch (const Error &e) {
        nan_threw = e.code == SRB_ERR_NAN;
    }
    CHECK(nan_threw);
    // (3) the chunk iterator stand-in: row chunks with offsets rebased to 0, in order, covering every entry once
    CsView<double> kat{Format::Csr, 4, 5, kIndptr.data(), kIndices.data(), kData.data()};
    backed::HostChunkSource<double> src(kat);
    uint64_t rows = 0, nnz = 0, calls = 0;
    src.for_each_chunk(3, [&](const CsView<double> &c) {
        CHECK(c.offsets[0] == 0 && c.ncols == 5);
        CHECK(c.indices == kIndices.data() + nnz && c.values == kData.data() + nnz);
        rows += c.nrows, nnz += c.nnz(), ++calls;
    });
    CHECK(rows == 4 && nnz == 8 && calls == 2);
    // (3b) the on-disk chunk store: write, reopen, iterate — every chunk equals the in-memory iterator's
    if (store_dir) {
        CsMatrix<float> a = random_csr(257, 61, 0.2, 7);
        backed::write_store(store_dir, a.view());
        backed::StoreChunkSource<float> disk(store_dir);
        backed::HostChunkSource<float> mem(a.view());
        CHECK(disk.n_obs() == 257 && disk.n_vars() == 61 && disk.format() == Format::Csr);
        for (size_t chunk : {1u, 50u, 256u, 257u, 1000u}) {
            std::vector<std::vector<uint64_t>> offs, idxs;
            std::vector<std::vector<float>> vals;
            mem.for_each_chunk(chunk, [&](const CsView<float> &c) {
                offs.emplace_back(c.offsets, c.offsets + c.nmajor() + 1);
                idxs.emplace_back(c.indices, c.indices + c.nnz());
                vals.emplace_back(c.values, c.values + c.nnz());
            });
            size_t i = 0;
            disk.for_each_chunk(chunk, [&](const CsView<float> &c) {
                CHECK(i < offs.size() && c.ncols == 61 && c.nmajor() + 1 == offs[i].size());
                CHECK(std::equal(offs[i].begin(), offs[i].end(), c.offsets) && std::equal(idxs[i].begin(), idxs[i].end(), c.indices) &&
                      std::equal(vals[i].begin(), vals[i].end(), c.values));
                ++i;
            });
            CHECK(i == offs.size());
        }
        CsView<float> w = disk.whole();
        CHECK(w.nnz() == a.indices.size() && std::equal(a.values.begin(), a.values.end(), w.values));
        bool wrong_dtype = false;
        try {
            backed::StoreChunkSource<double> bad(store_dir);
        } catch (const Error &e) {
            wrong_dtype = e.code == SRB_ERR_UNSUPPORTED_DTYPE;
        }
        CHECK(wrong_dtype);
    }
    // (4) vocabulary
    CHECK((int32_t)Direction::Row == 0 && (int32_t)Direction::Column == 1 && shared::is_row(Direction::Row));
    CHECK(ComputationMode::Whole().is_whole() && !ComputationMode::Chunked(7).is_whole() && *ComputationMode::Chunked(7).chunk == 7);
    CHECK(FlexValue::Absolute(3).is_absolute() && FlexValue::Relative(0.5).is_relative() && FlexValue::None().is_none());
    CHECK(dtype_of<float>::value == SRB_F32 && dtype_of<uint16_t>::value == SRB_U16 && dtype_of<int64_t>::value == SRB_I64);
    printf("cpu-check: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}

static void kat_checks(Device &dev, Device &faithful) {
    CsView<double> kat{Format::Csr, 4, 5, kIndptr.data(), kIndices.data(), kData.data()};
    IMAnnData adata(dev, kat);
    namespace st = memory::statistics;
    CHECK(adata.n_obs() == 4 && adata.n_vars() == 5);
    CHECK(all_equal(st::compute_number(adata, Direction::Row), std::vector<uint32_t>{3, 2, 0, 3}));
    CHECK(all_equal(st::compute_number(adata, Direction::Column), std::vector<uint32_t>{3, 1, 2, 2, 0}));
    CHECK(all_equal(st::compute_sum(adata, Direction::Row), std::vector<double>{6, 9, 0, 10}));
    CHECK(all_equal(st::compute_sum(adata, Direction::Column), std::vector<double>{7, 5, 4, 9, 0}));
    CHECK(all_close(st::compute_variance(adata, Direction::Row), {0.6666666666666666, 0.25, NaN, 3.555555555555556}, 1e-12));
    CHECK(all_close(st::compute_variance(adata, Direction::Column), {1.5555555555555545, 0, 0, 2.25, 0}, 1e-12, 1e-15));
    CHECK(all_close(st::compute_std_dev(adata, Direction::Row), {0.816496580927726, 0.5, NaN, 1.8856180831641267}, 1e-12));
    CHECK(all_close(st::compute_std_dev(adata, Direction::Column), {1.2472191289246466, 0, 0, 1.5, 0}, 1e-12, 1e-15));
    auto mmr = st::compute_min_max(adata, Direction::Row);
    CHECK(all_equal(mmr.first, std::vector<double>{1, 4, Inf, 2}) && all_equal(mmr.second, std::vector<double>{3, 5, -Inf, 6}));
    auto mmc = st::compute_min_max(adata, Direction::Column);
    CHECK(all_equal(mmc.first, std::vector<double>{1, 5, 2, 3, Inf}) && all_equal(mmc.second, std::vector<double>{4, 5, 2, 6, -Inf}));
    // qc_vars_inplace writes the reference's column names (memory/statistics/mod.rs:80-97)
    st::qc_vars_inplace(adata);
    for (const char *c : {"num_genes_per_cell", "sum_expr_per_cell", "var_expr_per_cell", "std_dev_per_cell"}) CHECK(adata.obs.count(c) == 1);
    for (const char *c : {"num_cells_per_gene", "sum_expr_per_gene", "var_expr_per_gene", "std_dev_per_gene"}) CHECK(adata.var.count(c) == 1);
    CHECK(all_equal(std::get<std::vector<uint32_t>>(adata.obs["num_genes_per_cell"]), std::vector<uint32_t>{3, 2, 0, 3}));
    CHECK(all_equal(std::get<std::vector<double>>(adata.var["sum_expr_per_gene"]), std::vector<double>{7, 5, 4, 9, 0}));

    // normalize_total (non-inplace = deep_clone + inplace, processing/mod.rs:314-322) leaves the input untouched
    namespace pr = memory::processing;
    IMAnnData f(faithful, kat);
    IMAnnData norm = pr::normalize_total(f, 10.0, Direction::Row);
    CHECK(all_equal(f.x().download().values, kData));
    CHECK(all_close(norm.x().download().values,
                    {1.6666666666666667, 3.3333333333333335, 5, 4.444444444444445, 5.555555555555555, 2, 2, 6}, 1e-15));
    IMAnnData lg = pr::log1p_transform(norm);
    CHECK(all_close(lg.x().download().values,
                    {0.9808292530117262, 1.4663370687934272, 1.791759469228055, 1.6945957207744073, 1.8803128665695001,
                     1.0986122886681096, 1.0986122886681096, 1.9459101490553132}, 1e-14));
    CHECK(all_close(st::compute_variance(lg, Direction::Column),
                    {0.09761462948178345, 0, 0.033805378479553116, 0.005940608022801719, 0}, 1e-6, 1e-9));
    CHECK(all_equal(pr::select_features(lg, FeatureSelection::HighlyVariable(3)), std::vector<uint64_t>{0, 2, 3}));
    CHECK(all_equal(pr::select_features(lg, FeatureSelection::HighlyVariable(5)), std::vector<uint64_t>{0, 2, 3, 1, 4}));
    CHECK(all_equal(pr::select_features(lg, FeatureSelection::None()), std::vector<uint64_t>{0, 1, 2, 3, 4}));
    CHECK(all_equal(pr::select_features(lg, FeatureSelection::VarianceThreshold(0.01)), std::vector<uint64_t>{0, 2}));
    CHECK(pr::select_features(lg, FeatureSelection::Randomized(3)).size() == 3);
    lg.var["highly_variable"] = std::vector<uint8_t>{1, 0, 0, 1, 0};
    CHECK(all_equal(pr::select_features(lg, FeatureSelection::HighlyVariableCol("highly_variable")), std::vector<uint64_t>{0, 3}));
    bool threw = false;
    try {
        pr::select_features(lg, FeatureSelection::HighlyVariableCol("missing"));
    } catch (const Error &) {
        threw = true;
    }
    CHECK(threw);
    Array2 d = lg.x().densify_selected({0, 2, 3});
    CHECK(d.rows == 4 && d.cols == 3 && close(d(0, 1), 1.4663370687934272, 1e-14) && d(2, 0) == 0 && d(2, 1) == 0 && d(2, 2) == 0);
    // unsupported dtype: the reference panics for I64 (shared/mod.rs:117); here an Error
    std::vector<int64_t> i64(kData.begin(), kData.end());
    CsView<int64_t> bad{Format::Csr, 4, 5, kIndptr.data(), kIndices.data(), i64.data()};
    threw = false;
    try {
        IMAnnData x(dev, bad);
    } catch (const Error &e) {
        threw = e.code == SRB_ERR_UNSUPPORTED_DTYPE;
    }
    CHECK(threw);
}

static void random_checks(Device &dev, Device &faithful) {
    const uint64_t n = 3000, m = 500;
    CsMatrix<float> a = random_csr(n, m, 0.06, 42);
    const uint64_t *off = a.offsets.data(), *idx = a.indices.data();
    const float *val = a.values.data();
    IMAnnData adata(dev, a.view());
    namespace st = memory::statistics;
    for (Direction d : {Direction::Row, Direction::Column}) {
        const int major = d == Direction::Row;
        const uint64_t len = major ? n : m;
        std::vector<uint32_t> num(len);
        std::vector<double> sum(len), var(len), sd(len), mn(len), mx(len);
        orc_number(n, m, off, idx, major, num.data());
        orc_sum_f32(n, m, off, idx, val, major, sum.data());
        orc_variance_f32(n, m, off, idx, val, major, var.data());
        orc_std_dev_f32(n, m, off, idx, val, major, sd.data());
        orc_min_max_f32(n, m, off, idx, val, major, mn.data(), mx.data());
        CHECK(all_equal(st::compute_number(adata, d), num));  // bit-exact
        CHECK(all_equal(st::compute_sum(adata, d), sum));     // integer counts: exact in f64
        CHECK(all_close(st::compute_variance(adata, d), var, 1e-5, 1e-9));
        CHECK(all_close(st::compute_std_dev(adata, d), sd, 1e-5, 1e-6));
        auto mm = st::compute_min_max(adata, d);
        CHECK(all_equal(mm.first, mn) && all_equal(mm.second, mx));
    }
    // backed::statistics, Whole and Chunked, against the same oracle numbers
    backed::HostChunkSource<float> src(a.view());
    std::vector<uint32_t> num_c(m), num_r(n);
    std::vector<double> sum_c(m), sum_r(n);
    orc_number(n, m, off, idx, 0, num_c.data());
    orc_number(n, m, off, idx, 1, num_r.data());
    orc_sum_f32(n, m, off, idx, val, 0, sum_c.data());
    orc_sum_f32(n, m, off, idx, val, 1, sum_r.data());
    for (ComputationMode mode : {ComputationMode::Whole(), ComputationMode::Chunked(1), ComputationMode::Chunked(777), ComputationMode::Chunked(5000)}) {
        if (!mode.is_whole() && *mode.chunk == 1 && n > 500) continue;  // one row per chunk is exercised on the small matrix below
        CHECK(all_equal(backed::statistics::compute_number(dev, src, Direction::Column, mode), num_c));
        CHECK(all_equal(backed::statistics::compute_number(dev, src, Direction::Row, mode), num_r));
        CHECK(all_equal(backed::statistics::compute_sum(dev, src, Direction::Column, mode), sum_c));
        CHECK(all_equal(backed::statistics::compute_sum(dev, src, Direction::Row, mode), sum_r));
    }
    {
        CsView<double> kat{Format::Csr, 4, 5, kIndptr.data(), kIndices.data(), kData.data()};
        backed::HostChunkSource<double> ks(kat);
        CHECK(all_equal(backed::statistics::compute_sum(dev, ks, Direction::Row, ComputationMode::Chunked(1)), std::vector<double>{6, 9, 0, 10}));
        CHECK(all_equal(backed::statistics::compute_number(dev, ks, Direction::Column, ComputationMode::Chunked(1)), std::vector<uint32_t>{3, 1, 2, 2, 0}));
    }

    // normalise + log1p in FAITHFUL (f64) mode against the oracle, then HVG order bit-exact
    namespace pr = memory::processing;
    IMAnnData f(faithful, a.view());
    pr::normalize_total_inplace(f, 1e4, Direction::Row);
    pr::log1p_transform_inplace(f);
    std::vector<double> want(a.values.size());
    orc_normalize_total_f32(n, m, off, idx, val, 1, 1e4, want.data());
    orc_log1p_f64(want.size(), want.data());
    CHECK(all_close(f.x().download().values, want, 1e-13));
    std::vector<double> gv(m);
    orc_variance_f64(n, m, off, idx, want.data(), 0, gv.data());
    CHECK(all_close(memory::statistics::compute_variance(f, Direction::Column), gv, 1e-9, 1e-12));
    std::vector<uint64_t> hv(50);
    CHECK(orc_select_hvg(gv.data(), m, 50, hv.data()) == 0);
    CHECK(all_equal(pr::select_features(f, FeatureSelection::HighlyVariable(50)), hv));

    // filters: masks from the oracle's counts / sums through the same FlexValue arms, then subset
    std::vector<uint32_t> cnt_r(n);
    std::vector<double> s_r(n);
    orc_number(n, m, off, idx, 1, cnt_r.data());
    orc_sum_f32(n, m, off, idx, val, 1, s_r.data());
    const FlexValue lo = FlexValue::Absolute(25), hi = FlexValue::Relative(0.9);
    auto pc = pr::calculate_percentiles(s_r, lo, hi);
    auto mask = pr::create_filter_mask(n, cnt_r, s_r, lo, hi, pc.first, pc.second);
    const uint64_t kept = std::accumulate(mask.begin(), mask.end(), uint64_t(0));
    IMAnnData adata2(dev, a.view());
    memory::statistics::qc_vars_inplace(adata2);
    IMAnnData fc = pr::filter_cells(adata2, lo, hi);
    CHECK(fc.n_obs() == kept && kept > 0 && kept < n && fc.n_vars() == m);
    CHECK(std::get<std::vector<uint32_t>>(fc.obs["num_genes_per_cell"]).size() == kept);
    std::vector<double> want_sums;
    for (uint64_t i = 0; i < n; ++i)
        if (mask[i]) want_sums.push_back(s_r[i]);
    CHECK(all_equal(memory::statistics::compute_sum(fc, Direction::Row), want_sums));
    pr::filter_genes_inplace(adata2, FlexValue::Absolute(1), FlexValue::None());  // drops the empty gene 0
    CHECK(adata2.n_vars() < m && adata2.n_obs() == n);
    auto ng = memory::statistics::compute_number(adata2, Direction::Column);
    CHECK(std::all_of(ng.begin(), ng.end(), [](uint32_t c) { return c >= 1; }));

    // pca_inplace: shapes, ordering of the explained-variance ratio, orthonormal components, tensor-core Gram == fp64 Gram
    IMAnnData p(dev, a.view());
    pr::normalize_total_inplace(p, 1e4, Direction::Row);
    pr::log1p_transform_inplace(p);
    IMAnnData p64 = p.deep_clone();
    pr::pca_inplace(p, 10, std::nullopt, std::nullopt, 8, FeatureSelection::HighlyVariable(128), pr::SVDMode::Faer);
    pr::pca_inplace(p64, 10, true, true, std::nullopt, FeatureSelection::HighlyVariable(128), pr::SVDMode::Lapack, /*gram_mode=*/1);
    const Array2 &sc = p.obsm["X_pca"], &ld = p.varm["PCA_loadings"];
    CHECK(sc.rows == n && sc.cols == 10 && ld.rows == m && ld.cols == 10 && p.explained_variance_ratio.size() == 10);
    for (size_t c = 1; c < 10; ++c) CHECK(p.explained_variance_ratio[c] <= p.explained_variance_ratio[c - 1]);
    CHECK(all_close(p.explained_variance_ratio, p64.explained_variance_ratio, 1e-5));
    for (size_t c1 = 0; c1 < 10; ++c1)
        for (size_t c2 = c1; c2 < 10; ++c2) {
            double dot = 0;
            for (size_t j = 0; j < m; ++j) dot += ld(j, c1) * ld(j, c2);
            CHECK(std::fabs(dot - (c1 == c2 ? 1.0 : 0.0)) < 1e-8);
        }
    // default n_components = 2; fewer than two features is an error, not a panic
    IMAnnData q(dev, a.view());
    pr::pca_inplace(q, std::nullopt, std::nullopt, std::nullopt, std::nullopt, FeatureSelection::HighlyVariable(16));
    CHECK(q.obsm["X_pca"].cols == 2);
    bool threw = false;
    try {
        pr::pca_inplace(q, 2, true, true, std::nullopt, FeatureSelection::HighlyVariable(1));
    } catch (const Error &) {
        threw = true;
    }
    CHECK(threw);
    // backed pipeline (config 5 in miniature): chunked upload, resident, same scores as the in-memory path
    IMAnnData bk = backed::processing::normalize_hvg_pca(dev, src, ComputationMode::Chunked(700), 1e4, 128, 10);
    CHECK(all_close(bk.obsm["X_pca"].data, p.obsm["X_pca"].data, 1e-9, 1e-9));
    CHECK(all_close(bk.explained_variance_ratio, p.explained_variance_ratio, 1e-12));
}

// out-of-core pipeline (three passes over row chunks) against the resident one on the same matrix
static int out_of_core_check() {
    Device dev(0);
    CsMatrix<float> a = random_csr(5000, 1200, 0.05, 11);
    // plant three cell programmes so that the two leading components are well separated
    for (uint64_t i = 0; i < a.nrows; ++i)
        for (uint64_t p = a.offsets[i]; p < a.offsets[i + 1]; ++p)
            if (a.indices[p] % 3 == i % 3) a.values[p] *= 4.0f;
    backed::HostChunkSource<float> src(a.view());
    namespace pr = memory::processing;
    for (int32_t gram_mode : {1, 0}) {  // 1: fp64 Gram (only the summation order differs), 0: tensor cores (1e-6-level Gram)
        const double tol = gram_mode == 1 ? 1e-7 : 1e-4;
        IMAnnData ref(dev, a.view());
        pr::normalize_total_inplace(ref, 1e4, Direction::Row);
        pr::log1p_transform_inplace(ref);
        pr::pca_inplace(ref, 2, true, true, std::nullopt, FeatureSelection::HighlyVariable(200), pr::SVDMode::Lapack, gram_mode);
        auto sel_ref = pr::select_features(ref, FeatureSelection::HighlyVariable(200));
        for (size_t chunk : {700u, 5000u, 64u}) {
            auto r = backed::processing::normalize_hvg_pca_out_of_core(dev, src, ComputationMode::Chunked(chunk), 1e4, 200, 2, true, true, gram_mode);
            CHECK(all_equal(r.selection, sel_ref));
            CHECK(all_close(r.explained_variance_ratio, ref.explained_variance_ratio, tol));
            const Array2 &want = ref.obsm["X_pca"];
            CHECK(r.scores.rows == want.rows && r.scores.cols == 2 && r.components.rows == 200 && r.components.cols == 2);
            double scale = 0, err = 0;
            for (size_t c = 0; c < 2; ++c) {
                double dot = 0;
                for (size_t i = 0; i < want.rows; ++i) dot += r.scores(i, c) * want(i, c);
                const double sgn = dot < 0 ? -1.0 : 1.0;
                for (size_t i = 0; i < want.rows; ++i) {
                    err = std::max(err, std::fabs(sgn * r.scores(i, c) - want(i, c)));
                    scale = std::max(scale, std::fabs(want(i, c)));
                }
            }
            CHECK(err <= tol * scale);
        }
    }
    printf("out-of-core: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--cpu-check") return cpu_check(argc > 2 ? argv[2] : nullptr);
    if (argc > 1 && std::string(argv[1]) == "--out-of-core") {
        try {
            return out_of_core_check();
        } catch (const Error &e) {
            fprintf(stderr, "single_rust::Error %d: %s\n", e.code, e.what());
            return 2;
        }
    }
    if (argc > 3 && std::string(argv[1]) == "--store-sums") {  // per-chunk value sums of a chunk store written by someone else
        backed::StoreChunkSource<float> disk(argv[2]);
        disk.for_each_chunk((size_t)std::atoll(argv[3]), [&](const CsView<float> &c) {
            double s = 0;
            uint64_t is = 0;
            for (uint64_t i = 0; i < c.nnz(); ++i) s += c.values[i], is += c.indices[i];
            printf("%llu %llu %.17g %llu\n", (unsigned long long)c.nmajor(), (unsigned long long)c.nnz(), s, (unsigned long long)is);
        });
        return 0;
    }
    try {
        Device dev(0), faithful(0, SRB_VALUES_FAITHFUL);
        kat_checks(dev, faithful);
        random_checks(dev, faithful);
        dev.set_upload_mode(SRB_UPLOAD_HOST_PACK);  // the same checks through the packed upload
        kat_checks(dev, faithful);
    } catch (const Error &e) {
        fprintf(stderr, "single_rust::Error %d: %s\n", e.code, e.what());
        return 2;
    }
    printf("host_mirror_test: %d checks, %d failed, kernel launches %" PRIu64 "\n", g_checks, g_fail, srb_kernel_launch_count());
    return g_fail ? 1 : 0;
}
