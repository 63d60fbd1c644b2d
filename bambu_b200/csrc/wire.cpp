// wire.cpp -- arena (v1, include/bambu_em.h) -> wire form v2 ("live arena"), native host code.
//
// The host->device copy of the arena is what bounds a cohort end to end (8 ranks share one host's PCIe), so the
// packer sends only what the EM reads.  Exact eliminations, the same two the device pack applies to the v1 arena:
//   * a column with Y_j == 0 contributes +-0 / NaN->0 to every sum it takes part in (src/em.cpp:48-49, 98-99);
//   * an entry with aval == 0 contributes +-0.
// What a dropped entry still influences is rs_i = sum_j X_ij over ALL entries of the row (src/em.cpp:33): it is
// evaluated here -- Armadillo's two accumulators by column parity, the order em_sell.cuh / the oracle use -- and
// travels as one double per row.  Entries keep the parity of their ORIGINAL column index (bit 0 of the index),
// which is all the summation order needs.
//
// Per column the packer sends the integer count and the gene's K where both are integers below 2^31 and
// Y == nobs / K holds bit for bit (R/bambu-quantify_utilityFunctions.R:221-228 builds Y exactly so): the device
// then forms Y with the same IEEE division.  Otherwise Y and K travel as doubles (col_mode 0).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/bambu_em.h"

struct bambu_wire {
    std::vector<int64_t> grp_row_ptr, grp_col_ptr, grp_ent_ptr, nnz_nonzero;
    std::vector<uint32_t> row_end, eq_bits, multi_bits;
    std::vector<double> rs, aval;
    std::vector<uint8_t> cidx, colY, colK;
    int32_t col_mode = 1;
};

static thread_local std::string g_wire_error;

static inline int idx_width(int64_t NL) { return NL <= 32767 ? 2 : 4; }

extern "C" {

const char *bambu_wire_last_error(void) { return g_wire_error.c_str(); }

int bambu_wire_free(bambu_wire *w)
{
    delete w;
    return BAMBU_EM_OK;
}

int bambu_wire_create(bambu_wire **out, int64_t n_groups, const int64_t *grp_row_ptr, const int64_t *grp_col_ptr,
                      const int64_t *row_nnz_ptr, const int32_t *col_idx, const double *aval, const uint8_t *nz_flags,
                      const double *Y, const double *K, const uint8_t *col_flags)
{
    if (!out) { g_wire_error = "NULL output"; return BAMBU_EM_EINVAL; }
    *out = nullptr;
    if (n_groups < 0 || !grp_row_ptr || !grp_col_ptr || !row_nnz_ptr) { g_wire_error = "NULL pointer array"; return BAMBU_EM_EINVAL; }
    for (int64_t g = 0; g < n_groups; ++g)
        if (grp_row_ptr[g + 1] < grp_row_ptr[g] || grp_col_ptr[g + 1] < grp_col_ptr[g]) {
            g_wire_error = "group pointer arrays must be non-decreasing";
            return BAMBU_EM_EINVAL;
        }
    const int64_t rows = grp_row_ptr[n_groups], cols = grp_col_ptr[n_groups];
    for (int64_t r = 0; r < rows; ++r)
        if (row_nnz_ptr[r + 1] < row_nnz_ptr[r]) { g_wire_error = "row_nnz_ptr must be non-decreasing"; return BAMBU_EM_EINVAL; }
    const int64_t nnz = row_nnz_ptr[rows];
    if ((nnz && (!col_idx || !aval || !nz_flags)) || (cols && (!Y || !K || !col_flags))) { g_wire_error = "NULL payload pointer"; return BAMBU_EM_EINVAL; }
    bambu_wire *W = new (std::nothrow) bambu_wire;
    if (!W) { g_wire_error = "out of memory"; return BAMBU_EM_ENOMEM; }
    try {
        // live columns and their compact index
        std::vector<int64_t> live_of((size_t)cols, -1);
        W->grp_row_ptr.assign(grp_row_ptr, grp_row_ptr + n_groups + 1);
        W->grp_col_ptr.assign((size_t)n_groups + 1, 0);
        W->grp_ent_ptr.assign((size_t)n_groups + 1, 0);
        W->nnz_nonzero.assign((size_t)n_groups, 0);
        bool ints = true;
        int64_t nl = 0;
        for (int64_t g = 0; g < n_groups; ++g) {
            for (int64_t j = grp_col_ptr[g]; j < grp_col_ptr[g + 1]; ++j) {
                if (!std::isfinite(K[j])) { throw std::string("K must be finite (it is a sum of read counts)"); }
                if (Y[j] != 0.0) {       // NaN counts as live, like on the device
                    live_of[(size_t)j] = nl++ - W->grp_col_ptr[(size_t)g];
                    if (ints) {
                        const double k = K[j], n = std::nearbyint(Y[j] * k);
                        if (!(k >= 1.0 && k < 2147483648.0 && k == std::floor(k) && n >= 0.0 && n < 2147483648.0 && n / k == Y[j]))
                            ints = false;
                    }
                }
            }
            W->grp_col_ptr[(size_t)g + 1] = nl;
        }
        W->col_mode = ints ? 1 : 0;
        const size_t cw = ints ? 4 : 8;
        W->colY.resize(cw * (size_t)nl);
        W->colK.resize(cw * (size_t)nl);
        W->multi_bits.assign(((size_t)nl + 31) / 32 + 1, 0u);
        {
            int64_t q = 0;
            for (int64_t j = 0; j < cols; ++j) {
                if (!(Y[j] != 0.0)) continue;
                if (ints) {
                    const int32_t k = (int32_t)K[j], n = (int32_t)std::nearbyint(Y[j] * K[j]);
                    memcpy(&W->colY[4 * (size_t)q], &n, 4);
                    memcpy(&W->colK[4 * (size_t)q], &k, 4);
                } else {
                    memcpy(&W->colY[8 * (size_t)q], &Y[j], 8);
                    memcpy(&W->colK[8 * (size_t)q], &K[j], 8);
                }
                if (col_flags[j] & 1) W->multi_bits[(size_t)q >> 5] |= 1u << (q & 31);
                ++q;
            }
        }
        // rows: rs over all entries, live entries compacted
        W->rs.resize((size_t)rows);
        W->row_end.resize((size_t)rows);
        W->aval.reserve((size_t)nnz / 2 + 16);
        std::vector<uint32_t> idx;          // the group's indices, narrowed when the group is written out
        int64_t ne = 0;
        for (int64_t g = 0; g < n_groups; ++g) {
            const int64_t c0 = grp_col_ptr[g], N = grp_col_ptr[g + 1] - c0;
            const int64_t NL = W->grp_col_ptr[(size_t)g + 1] - W->grp_col_ptr[(size_t)g];
            idx.clear();
            int64_t nz = 0;
            for (int64_t r = grp_row_ptr[g]; r < grp_row_ptr[g + 1]; ++r) {
                double a1 = 0.0, a2 = 0.0;
                int64_t prev = -1;
                for (int64_t k = row_nnz_ptr[r]; k < row_nnz_ptr[r + 1]; ++k) {
                    const int64_t j = col_idx[k];
                    if (j < 0 || j >= N) throw std::string("col_idx outside its group's column range");
                    if (j <= prev) throw std::string("col_idx not strictly increasing inside a row");
                    prev = j;
                    const double a = aval[k];
                    if (j & 1) a2 += a; else a1 += a;
                    if (a != 0.0) {
                        ++nz;
                        const int64_t lj = live_of[(size_t)(c0 + j)];
                        if (lj >= 0) {
                            idx.push_back((uint32_t)((lj << 1) | (j & 1)));
                            W->aval.push_back(a);
                            if (nz_flags[k] & 1) {
                                const size_t e = (size_t)ne + idx.size() - 1;
                                if ((e >> 5) >= W->eq_bits.size()) W->eq_bits.resize((e >> 5) + 1024, 0u);
                                W->eq_bits[e >> 5] |= 1u << (e & 31);
                            }
                        }
                    }
                }
                W->rs[(size_t)r] = a1 + a2;
                if (idx.size() > 0xffffffffull) throw std::string("a group has more than 2^32 live entries");
                W->row_end[(size_t)r] = (uint32_t)idx.size();
            }
            W->nnz_nonzero[(size_t)g] = nz;
            // index block: 4-byte aligned, u16 while the live column index fits 15 bits
            W->cidx.resize((W->cidx.size() + 3) / 4 * 4, 0);
            const size_t at = W->cidx.size();
            if (idx_width(NL) == 2) {
                W->cidx.resize(at + 2 * idx.size());
                uint16_t *d = reinterpret_cast<uint16_t *>(W->cidx.data() + at);
                for (size_t k = 0; k < idx.size(); ++k) d[k] = (uint16_t)idx[k];
            } else {
                W->cidx.resize(at + 4 * idx.size());
                if (!idx.empty()) memcpy(W->cidx.data() + at, idx.data(), 4 * idx.size());
            }
            ne += (int64_t)idx.size();
            W->grp_ent_ptr[(size_t)g + 1] = ne;
        }
        W->eq_bits.resize(((size_t)ne + 31) / 32 + 1, 0u);
        W->cidx.resize((W->cidx.size() + 3) / 4 * 4 + 4, 0);
    } catch (const std::string &e) {
        g_wire_error = e;
        delete W;
        return BAMBU_EM_EINVAL;
    } catch (const std::bad_alloc &) {
        g_wire_error = "out of memory";
        delete W;
        return BAMBU_EM_ENOMEM;
    } catch (...) {
        g_wire_error = "unknown exception";
        delete W;
        return BAMBU_EM_EINVAL;
    }
    *out = W;
    return BAMBU_EM_OK;
}

int bambu_wire_sizes(const bambu_wire *w, int64_t *n_groups, int64_t *n_rows, int64_t *n_cols, int64_t *n_entries,
                     int64_t *cidx_bytes, int32_t *col_mode, int64_t *total_bytes)
{
    if (!w) return BAMBU_EM_EINVAL;
    const int64_t G = (int64_t)w->grp_row_ptr.size() - 1;
    if (n_groups) *n_groups = G;
    if (n_rows) *n_rows = w->grp_row_ptr.back();
    if (n_cols) *n_cols = w->grp_col_ptr.back();
    if (n_entries) *n_entries = w->grp_ent_ptr.back();
    if (cidx_bytes) *cidx_bytes = (int64_t)w->cidx.size();
    if (col_mode) *col_mode = w->col_mode;
    if (total_bytes)
        *total_bytes = (int64_t)(8 * 3 * (w->grp_row_ptr.size()) + 4 * w->row_end.size() + 8 * w->rs.size() + 8 * w->aval.size() +
                                 w->cidx.size() + 4 * w->eq_bits.size() + 4 * w->multi_bits.size() + w->colY.size() + w->colK.size());
    return BAMBU_EM_OK;
}

int bambu_wire_arrays(const bambu_wire *w, const int64_t **grp_row_ptr, const int64_t **grp_col_ptr,
                      const int64_t **grp_ent_ptr, const uint32_t **row_end, const double **rs, const double **aval,
                      const void **cidx, const uint32_t **eq_bits, const void **colY, const void **colK,
                      const uint32_t **multi_bits, const int64_t **nnz_nonzero)
{
    if (!w) return BAMBU_EM_EINVAL;
    if (grp_row_ptr) *grp_row_ptr = w->grp_row_ptr.data();
    if (grp_col_ptr) *grp_col_ptr = w->grp_col_ptr.data();
    if (grp_ent_ptr) *grp_ent_ptr = w->grp_ent_ptr.data();
    if (row_end) *row_end = w->row_end.data();
    if (rs) *rs = w->rs.data();
    if (aval) *aval = w->aval.data();
    if (cidx) *cidx = w->cidx.data();
    if (eq_bits) *eq_bits = w->eq_bits.data();
    if (colY) *colY = w->colY.data();
    if (colK) *colK = w->colK.data();
    if (multi_bits) *multi_bits = w->multi_bits.data();
    if (nnz_nonzero) *nnz_nonzero = w->nnz_nonzero.data();
    return BAMBU_EM_OK;
}

}  // extern "C"
