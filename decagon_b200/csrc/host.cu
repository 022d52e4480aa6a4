// Host-only pieces: error slot, COO -> canonical CSR, transpose, sampler thresholds.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include <algorithm>
#include <numeric>

#include "dgn_internal.cuh"

namespace dgn {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

// Stable counting sort by row, then a stable sort by column inside each row: entries are
// ordered by (row, col) and duplicates keep their input order (scipy's sum_duplicates would
// merge them; preprocess_graph never emits duplicates, minibatch.py:80-93).
void csr_from_coo(int n_rows, int n_cols, int64_t nnz, const int32_t *rows, const int32_t *cols, const float *vals,
                  HostCsr &out) {
    DGN_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "csr_from_coo: negative size");
    DGN_REQUIRE(nnz < (int64_t)INT32_MAX, "csr_from_coo: nnz %lld does not fit int32 offsets", (long long)nnz);
    out.n_rows = n_rows;
    out.n_cols = n_cols;
    out.rowptr.assign((size_t)n_rows + 1, 0);
    out.col.resize((size_t)nnz);
    out.val.resize((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) {
        DGN_REQUIRE(rows[e] >= 0 && rows[e] < n_rows && cols[e] >= 0 && cols[e] < n_cols,
                    "csr_from_coo: entry %lld = (%d, %d) outside %d x %d", (long long)e, rows[e], cols[e], n_rows,
                    n_cols);
        out.rowptr[(size_t)rows[e] + 1]++;
    }
    for (int r = 0; r < n_rows; ++r) out.rowptr[r + 1] += out.rowptr[r];
    std::vector<int> cursor(out.rowptr.begin(), out.rowptr.end() - 1);
    std::vector<int64_t> src((size_t)nnz);
    for (int64_t e = 0; e < nnz; ++e) src[(size_t)cursor[rows[e]]++] = e;
    for (int r = 0; r < n_rows; ++r) {
        auto b = src.begin() + out.rowptr[r], e = src.begin() + out.rowptr[r + 1];
        bool sorted = true;
        for (auto it = b; sorted && it + 1 < e; ++it) sorted = cols[*it] <= cols[*(it + 1)];
        if (!sorted) std::stable_sort(b, e, [&](int64_t x, int64_t y) { return cols[x] < cols[y]; });
    }
    for (int64_t e = 0; e < nnz; ++e) {
        out.col[(size_t)e] = cols[src[(size_t)e]];
        out.val[(size_t)e] = vals[src[(size_t)e]];
    }
}

void csr_transpose(const HostCsr &a, HostCsr &out) {
    out.n_rows = a.n_cols;
    out.n_cols = a.n_rows;
    out.rowptr.assign((size_t)a.n_cols + 1, 0);
    out.col.resize(a.col.size());
    out.val.resize(a.val.size());
    for (int c : a.col) out.rowptr[(size_t)c + 1]++;
    for (int r = 0; r < a.n_cols; ++r) out.rowptr[r + 1] += out.rowptr[r];
    std::vector<int> cursor(out.rowptr.begin(), out.rowptr.end() - 1);
    for (int r = 0; r < a.n_rows; ++r)
        for (int e = a.rowptr[r]; e < a.rowptr[r + 1]; ++e) {
            int dst = cursor[a.col[e]]++;
            out.col[dst] = r;
            out.val[dst] = a.val[e];
        }
}

}  // namespace dgn

using namespace dgn;

extern "C" const char *dgn_last_error(void) { return g_error; }
extern "C" int dgn_version(void) { return 100; }

extern "C" int dgn_csr_from_coo(int32_t n_rows, int32_t n_cols, int64_t nnz, const int32_t *coo_rows,
                                const int32_t *coo_cols, const float *vals, int32_t *rowptr_out, int32_t *col_out,
                                float *val_out) {
    try {
        HostCsr c;
        csr_from_coo(n_rows, n_cols, nnz, coo_rows, coo_cols, vals, c);
        std::copy(c.rowptr.begin(), c.rowptr.end(), rowptr_out);
        std::copy(c.col.begin(), c.col.end(), col_out);
        std::copy(c.val.begin(), c.val.end(), val_out);
        return DGN_OK;
    } catch (const Failure &f) {
        return f.code;
    } catch (const std::exception &e) {  // std::bad_alloc and friends must not cross the C ABI
        set_error("host error: %s", e.what());
        return DGN_ERR_INVALID;
    }
}

// d^0.75 = sqrt(d) * sqrt(sqrt(d)): sqrt is correctly rounded, so the table is identical on
// every host (libm pow is not).  thr[v] = min(floor(cum[v] / total * 2^32), 2^32 - 1).
extern "C" int dgn_sampler_thresholds(const double *degrees, int32_t n, uint32_t *thresholds_out) {
    try {
        DGN_REQUIRE(n > 0, "sampler: empty degree table");
        std::vector<double> cum((size_t)n);
        double run = 0.0;
        for (int v = 0; v < n; ++v) {
            DGN_REQUIRE(degrees[v] >= 0.0, "sampler: negative degree at %d", v);
            double s = sqrt(degrees[v]);
            run += s * sqrt(s);
            cum[(size_t)v] = run;
        }
        DGN_REQUIRE(run > 0.0, "sampler: all degrees are zero");
        for (int v = 0; v < n; ++v) {
            double t = floor(cum[(size_t)v] / run * 4294967296.0);
            thresholds_out[v] = t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
        }
        return DGN_OK;
    } catch (const Failure &f) {
        return f.code;
    } catch (const std::exception &e) {  // std::bad_alloc and friends must not cross the C ABI
        set_error("host error: %s", e.what());
        return DGN_ERR_INVALID;
    }
}
