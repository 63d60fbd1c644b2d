// quant_dev_hostexec.cpp -- TEST BUILD ONLY.  The passes of the device packer (quant_dev.inl) executed
// serially on the host, so the CPU test-suite can pin the pipeline logic (sort keys, run ids, order-free
// reductions, the integer x87 addition) against the host packer where no GPU exists.  Not linked into
// libbambu_em.so and never used by the product; the CUDA executor of quant_dev.cu is checked against the
// same host packer by the -m gpu tests.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/bambu_em.h"
#include "quant_pack.h"

#define QD_FN inline
#define QD_LAMBDA [=]

inline void qd_or(uint32_t *p, uint32_t v) { *p |= v; }
inline void qd_inc(uint32_t *p) { ++*p; }
inline void qd_add(double *p, double v) { *p += v; }
inline void qd_min(uint64_t *p, uint64_t v) { if (v < *p) *p = v; }
inline void qd_max(uint64_t *p, uint64_t v) { if (v > *p) *p = v; }

#include "quant_dev.inl"
#include "r_round.h"

namespace {

struct HostExec {
    std::vector<void *> blocks;
    ~HostExec() { for (void *p : blocks) free(p); }
    template <class T>
    T *alloc(size_t n)
    {
        void *p = malloc(std::max<size_t>(n * sizeof(T), 1));
        if (!p) throw std::bad_alloc();
        memset(p, 0xA5, std::max<size_t>(n * sizeof(T), 1));      // uninitialised on the device too
        blocks.push_back(p);
        return static_cast<T *>(p);
    }
    void need(int) {}
    void lap(const char *) {}
    size_t mark() const { return 0; }
    void release(size_t) {}
    void fill(void *p, int byte, size_t bytes) { memset(p, byte, bytes); }
    template <class F>
    void pfor(size_t n, F f)
    {
        // reversed on purpose: nothing in the pipeline may depend on the order of a parallel loop
        for (size_t i = n; i-- > 0;) f(i);
    }
    void minmax(const int64_t *x, size_t n, int64_t *out)
    {
        out[0] = *std::min_element(x, x + n);
        out[1] = *std::max_element(x, x + n);
    }
    uint32_t scan(const uint32_t *in, uint32_t *out, size_t n)
    {
        uint32_t run = 0;
        for (size_t i = 0; i < n; ++i) { const uint32_t t = in[i]; out[i] = run; run += t; }
        return run;
    }
    void sort_pairs(uint64_t *&k, uint32_t *&v, uint64_t *&k2, uint32_t *&v2, size_t n, int bits)
    {
        if (n < 2) return;
        const uint64_t mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
        std::vector<uint32_t> idx(n);
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (k[a] & mask) < (k[b] & mask); });
        for (size_t i = 0; i < n; ++i) { k2[i] = k[idx[i]]; v2[i] = v[idx[i]]; }
        std::swap(k, k2);
        std::swap(v, v2);
    }
    template <class T> void d2h(T *dst, const T *src, size_t n) { if (n) memcpy(dst, src, n * sizeof(T)); }
    template <class T> void h2d(T *dst, const T *src, size_t n) { if (n) memcpy(dst, src, n * sizeof(T)); }
};

thread_local std::string g_err;

template <class T>
void take(std::vector<T> &dst, const T *src, size_t n) { dst.assign(src, src + n); }

}  // namespace

extern "C" {

const char *qd_hostexec_last_error(void) { return g_err.c_str(); }

// returns 0 ok, 1 "unsupported by the device formulation", BAMBU_EM_EINVAL for the reference's input errors.
// Arrays come back through caller-provided buffers sized for n rows (sizes[] = n_groups, n_rows, n_cols, n_entries,
// n_out, n_unobserved).
int qd_hostexec_pack(int64_t n, const int64_t *gene, const int64_t *tx, const int64_t *eq, const double *nobs,
                     const uint8_t *equal, const double *rcw, const double *txl, int32_t d_mode, int64_t *sizes,
                     int64_t *grp_row_ptr, int64_t *grp_col_ptr, int64_t *row_nnz_ptr, int32_t *col_idx, double *aval,
                     uint8_t *nz_flags, double *Y, double *K, uint8_t *col_flags, int32_t *txid, int64_t *out_txid,
                     int64_t *unobserved_txid, int64_t *gene_grp_id, double *d_rate)
{
    try {
        HostExec x;
        qd::Input in;
        in.n = (size_t)n; in.gene = gene; in.tx = tx; in.eq = eq; in.nobs = nobs; in.equal = equal;
        in.rcw = d_mode ? rcw : nullptr; in.txl = d_mode ? txl : nullptr; in.d_mode = d_mode != 0;
        qd::Result R;
        const int rc = qd::pack(x, in, R, g_err);
        if (rc == qd::kUnsupported) return 1;
        if (rc == qd::kInvalid) return BAMBU_EM_EINVAL;
        sizes[0] = R.n_groups; sizes[1] = R.n_rows; sizes[2] = R.n_cols; sizes[3] = (int64_t)R.n_entries;
        sizes[4] = (int64_t)R.out_txid.size(); sizes[5] = (int64_t)R.unobserved_txid.size();
        memcpy(grp_row_ptr, R.grp_row_ptr, 8 * ((size_t)R.n_groups + 1));
        memcpy(grp_col_ptr, R.grp_col_ptr, 8 * ((size_t)R.n_groups + 1));
        memcpy(row_nnz_ptr, R.row_nnz_ptr, 8 * ((size_t)R.n_rows + 1));
        if (R.n_entries) {
            memcpy(col_idx, R.col_idx, 4 * R.n_entries); memcpy(aval, R.aval, 8 * R.n_entries);
            memcpy(nz_flags, R.nz_flags, R.n_entries);
            memcpy(Y, R.Y, 8 * (size_t)R.n_cols); memcpy(K, R.K, 8 * (size_t)R.n_cols); memcpy(col_flags, R.col_flags, R.n_cols);
            memcpy(txid, R.txid, 4 * (size_t)R.n_rows);
        }
        std::copy(R.out_txid.begin(), R.out_txid.end(), out_txid);
        std::copy(R.unobserved_txid.begin(), R.unobserved_txid.end(), unobserved_txid);
        std::copy(R.gene_grp_id.begin(), R.gene_grp_id.end(), gene_grp_id);
        *d_rate = R.d_rate;
        return 0;
    } catch (const std::bad_alloc &) {
        g_err = "out of memory";
        return BAMBU_EM_ENOMEM;
    }
}

// the integer x87 addition against the real thing
double qd_hostexec_ext_sum(const double *x, int64_t n)
{
    qd::Ext acc;
    acc.m = 0; acc.e = 0;
    for (int64_t i = 0; i < n; ++i) acc = qd::ext_add(acc, qd::ext_from_double(x[i]));
    bool oor = false;
    const double d = qd::ext_to_double(acc, &oor);
    return oor ? NAN : d;
}
// R's round(x, digits) as the device epilogue of the cohort stage does it (r_round.h)
void qd_hostexec_r_round(const double *x, int64_t n, int digits, double *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = qd::r_round(x[i], digits);
}
double qd_hostexec_ld_sum(const double *x, int64_t n)
{
    long double s = 0;
    for (int64_t i = 0; i < n; ++i) s += (long double)x[i];
    return (double)s;
}

}  // extern "C"
