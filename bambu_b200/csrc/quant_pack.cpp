// quant_pack.cpp -- native host side of bambu.quantDT around the EM: the long table -> the arena
// (addAval ... assignGroups ... getInputList) and the merge of the estimates back by txid.
//
// Same semantics, step by step, as the reference's R code (bambu 3.3.5):
//   bambu.quantDT                     R/bambu-quantify.R:53-74
//   addAval / simplifyNames / calculateDegradationRate / modifyAvaluewithDegradation_rate /
//   removeUnObservedGenes             R/bambu-quantify_utilityFunctions.R:196-315
//   initialiseOutput / filterTxRc / assignGroups / getInputList     :322-371
//   modifyQuantOut / removeDuplicates :438-455
// and as bambu_b200/quantify.py (the vectorised numpy mirror), against which the tests compare it
// bit for bit.  R's sum()/mean() accumulate in long double; x86-64 long double is that type.
// Sorting is LSD radix on dense integer keys, everything else is linear passes.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/bambu_em.h"
#include "quant_pack.h"

namespace {

using qp::ld;
using qp::bits_for;
using qp::r_median;

thread_local std::string g_pack_error;

// ---- sorting: LSD radix on (64-bit key, 32-bit payload) pairs, 11-bit digits, streaming passes ----
struct KV {
    uint64_t k;
    uint32_t v;
};

void radix_kv(std::vector<KV> &a, int key_bits)
{
    const size_t n = a.size();
    if (n < 2) return;
    std::vector<KV> tmp(n);
    for (int sh = 0; sh < key_bits; sh += 11) {
        uint32_t cnt[2049] = {0};
        for (size_t i = 0; i < n; ++i) ++cnt[((a[i].k >> sh) & 2047u) + 1];
        if (cnt[((a[0].k >> sh) & 2047u) + 1] == n) continue;      // constant digit
        for (int b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
        for (size_t i = 0; i < n; ++i) tmp[cnt[(a[i].k >> sh) & 2047u]++] = a[i];
        a.swap(tmp);
    }
}

typedef std::vector<uint32_t> Perm;

// dense ids of a 64-bit code column.  by_first_appearance: numbered like R's match(x, unique(x));
// otherwise by sorted value (order preserving).  Small value ranges use a direct table.
uint32_t dense_ids(const int64_t *x, size_t n, bool by_first_appearance, std::vector<uint32_t> &id,
                   std::vector<int64_t> *values = nullptr)
{
    id.resize(n);
    if (!n) { if (values) values->clear(); return 0; }
    int64_t lo = x[0], hi = x[0];
    for (size_t i = 1; i < n; ++i) { lo = std::min(lo, x[i]); hi = std::max(hi, x[i]); }
    const uint64_t range = (uint64_t)hi - (uint64_t)lo;
    if (range < (uint64_t)(8 * n + (1u << 20))) {
        std::vector<uint32_t> tab((size_t)range + 1, UINT32_MAX);
        uint32_t k = 0;
        if (by_first_appearance) {
            for (size_t i = 0; i < n; ++i) {
                uint32_t &t = tab[(size_t)((uint64_t)x[i] - (uint64_t)lo)];
                if (t == UINT32_MAX) t = k++;
                id[i] = t;
            }
        } else {
            for (size_t i = 0; i < n; ++i) tab[(size_t)((uint64_t)x[i] - (uint64_t)lo)] = 0;
            if (values) values->clear();
            for (size_t v = 0; v <= range; ++v)
                if (tab[v] != UINT32_MAX) { tab[v] = k++; if (values) values->push_back((int64_t)((uint64_t)lo + v)); }
            for (size_t i = 0; i < n; ++i) id[i] = tab[(size_t)((uint64_t)x[i] - (uint64_t)lo)];
        }
        return k;
    }
    // general case: sort (value, row)
    std::vector<KV> a(n);
    for (size_t i = 0; i < n; ++i) a[i] = KV{(uint64_t)x[i] - (uint64_t)lo, (uint32_t)i};
    radix_kv(a, bits_for(range));
    std::vector<uint32_t> run_of(n);
    std::vector<uint32_t> first_row;      // smallest row of each run (stable sort: the first one met)
    uint32_t k = 0;
    if (values) values->clear();
    for (size_t i = 0; i < n; ++i) {
        if (i && a[i].k != a[i - 1].k) ++k;
        if (!i || a[i].k != a[i - 1].k) { first_row.push_back(a[i].v); if (values) values->push_back((int64_t)(a[i].k + (uint64_t)lo)); }
        run_of[a[i].v] = k;
    }
    ++k;
    if (!by_first_appearance) { id.swap(run_of); return k; }
    std::vector<KV> r(k);
    for (uint32_t q = 0; q < k; ++q) r[q] = KV{first_row[q], q};
    radix_kv(r, bits_for(n));
    std::vector<uint32_t> rank(k);
    for (uint32_t q = 0; q < k; ++q) rank[r[q].v] = q;
    for (size_t i = 0; i < n; ++i) id[i] = rank[run_of[i]];
    return k;
}

// ids of the distinct (a, b) pairs (a < na, b < nb), in deterministic (sorted) order
uint32_t pair_ids(const std::vector<uint32_t> &a, const std::vector<uint32_t> &b, uint32_t na, uint32_t nb,
                  std::vector<uint32_t> &id)
{
    const size_t n = a.size();
    const int bb = bits_for(nb ? nb - 1 : 0), ba = bits_for(na ? na - 1 : 0);
    std::vector<KV> kv(n);
    for (size_t i = 0; i < n; ++i) kv[i] = KV{((uint64_t)a[i] << bb) | b[i], (uint32_t)i};
    radix_kv(kv, ba + bb);
    id.resize(n);
    uint32_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        if (i && kv[i].k != kv[i - 1].k) ++k;
        id[kv[i].v] = k;
    }
    return n ? k + 1 : 0;
}

// R's pmax / pmin without na.rm: NA if either argument is NA
inline double r_pmax(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : (a > b ? a : b); }
inline double r_pmin(double a, double b) { return (std::isnan(a) || std::isnan(b)) ? NAN : (a < b ? a : b); }

}  // namespace

void bambu_quant_set_error(const std::string &s) { g_pack_error = s; }      // for quant_dev.cu

extern "C" {

const char *bambu_quant_last_error(void) { return g_pack_error.c_str(); }

int bambu_quant_pack_free(bambu_quant_pack *p)
{
    delete p;
    return BAMBU_EM_OK;
}

// Everything of bambu.quantDT up to (not including) the EM (R/bambu-quantify.R:53-62).
int bambu_quant_pack_create(bambu_quant_pack **out, int64_t n, const int64_t *gene_code, const int64_t *txid_in,
                            const int64_t *eq_in, const double *nobs, const uint8_t *equal, const double *rcWidth,
                            const double *txlen, int32_t degradation_bias)
{
    if (!out) { g_pack_error = "NULL output"; return BAMBU_EM_EINVAL; }
    *out = nullptr;
    if (n < 0 || n > (int64_t)UINT32_MAX - 8) { g_pack_error = "row count out of range"; return BAMBU_EM_EINVAL; }
    if (n && (!gene_code || !txid_in || !eq_in || !nobs)) {
        g_pack_error = "Columns GENEID, txid, eqClassId, nobs, are missing from object.";   // R/...utilityFunctions.R:199-203
        return BAMBU_EM_EINVAL;
    }
    // the arena carries txid as int32 (R's integer): anything wider would be truncated and merged under a wrong id
    for (int64_t i = 0; i < n; ++i)
        if (txid_in[i] < INT32_MIN || txid_in[i] > INT32_MAX) { g_pack_error = "txid outside the 32-bit integer range"; return BAMBU_EM_EINVAL; }
    bambu_quant_pack *P = new (std::nothrow) bambu_quant_pack;
    if (!P) { g_pack_error = "out of memory"; return BAMBU_EM_ENOMEM; }
    const size_t N = (size_t)n;
    const bool timing = getenv("BAMBU_PACK_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[quant_pack] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };
    try {
        // ---- simplifyNames (:236-241) + dense order-preserving codes for txid / eqClassId ----------
        std::vector<uint32_t> gene, tx, eq;
        std::vector<int64_t> tx_values;
        const uint32_t G = dense_ids(gene_code, N, true, gene);
        const uint32_t NT = dense_ids(txid_in, N, false, tx, &tx_values);
        const uint32_t NE = dense_ids(eq_in, N, false, eq);
        std::vector<uint8_t> eqf(N, 0);
        if (equal) for (size_t i = 0; i < N; ++i) eqf[i] = equal[i] ? 1 : 0;

        lap("dense ids");
        // (gene, eqClass) classes and (gene, tx) pairs
        std::vector<uint32_t> cls, gt;
        const uint32_t NC = pair_ids(gene, eq, G, NE, cls);
        const uint32_t NGT = pair_ids(gene, tx, G, NT, gt);
        // a class carries one nobs value (the reference takes unique(gene_sid, eqClassId, nobs) and would
        // otherwise count the class twice; getInputList then breaks) -- bambu_b200/quantify.py raises too
        {
            std::vector<double> v(NC);
            std::vector<uint8_t> seen(NC, 0);
            for (size_t i = 0; i < N; ++i) {
                if (!seen[cls[i]]) { seen[cls[i]] = 1; v[cls[i]] = nobs[i]; }
                else if (!(v[cls[i]] == nobs[i]) && !(std::isnan(v[cls[i]]) && std::isnan(nobs[i])))
                    throw std::string("an equivalence class carries more than one nobs value");
            }
        }

        lap("pair ids + nobs check");
        // ---- calculateDegradationRate (:247-269) ---------------------------------------------------
        const bool d_mode = degradation_bias != 0;
        double d_rate = NAN;
        if (d_mode) {
            std::vector<ld> gnobs(G, 0), gdobs(G, 0);
            std::vector<uint8_t> has(G, 0), has_par(G, 0), c_seen(NC, 0), p_seen(NC, 0);
            std::vector<double> glen(G, -INFINITY);
            for (size_t i = 0; i < N; ++i) {      // unique() keeps first occurrences in table order
                const uint32_t c = cls[i], g = gene[i];
                if (!c_seen[c]) { c_seen[c] = 1; gnobs[g] += (ld)nobs[i]; has[g] = 1; }
                if (!eqf[i] && !p_seen[c]) { p_seen[c] = 1; gdobs[g] += (ld)nobs[i]; has_par[g] = 1; }
                const double L = txlen ? txlen[i] : 1.0;
                if (L > glen[g]) glen[g] = L;
            }
            std::vector<double> nb(G), db(G);      // every gene code owns at least one row, so has[] is all ones
            for (uint32_t g = 0; g < G; ++g) { nb[g] = (double)gnobs[g]; db[g] = (double)gdobs[g]; }
            d_rate = qp::degradation_rate(G, nb.data(), db.data(), has_par.data(), glen.data());
        }
        P->d_rate = d_rate;

        lap("degradation rate");
        // ---- modifyAvaluewithDegradation_rate (:276-296) ---------------------------------------------
        std::vector<uint8_t> multi(N, 0);
        {
            // distinct transcripts per class: rows sorted by (class, tx)
            const int bt = bits_for(NT ? NT - 1 : 0);
            std::vector<KV> kv(N);
            for (size_t i = 0; i < N; ++i) kv[i] = KV{((uint64_t)cls[i] << bt) | tx[i], (uint32_t)i};
            radix_kv(kv, bits_for(NC ? NC - 1 : 0) + bt);
            std::vector<uint32_t> ntx(NC, 0);
            for (size_t i = 0; i < N; ++i)
                if (!i || kv[i].k != kv[i - 1].k) ++ntx[cls[kv[i].v]];
            for (size_t i = 0; i < N; ++i) multi[i] = ntx[cls[i]] > 1;
        }
        std::vector<double> aval(N, 1.0);
        if (d_mode) {
            std::vector<double> part(N);
            for (size_t i = 0; i < N; ++i) part[i] = (rcWidth ? rcWidth[i] : 0.0) * d_rate / 1000;
            std::vector<ld> psum(NGT, 0);
            for (size_t i = 0; i < N; ++i)
                if (multi[i] && !eqf[i]) psum[gt[i]] += (ld)part[i];            // R sum(): table order
            std::vector<uint8_t> isna(N, 1);
            for (size_t i = 0; i < N; ++i)
                if (multi[i]) { aval[i] = eqf[i] ? 1 - (double)psum[gt[i]] : part[i]; isna[i] = 0; }
                else aval[i] = NAN;
            if (d_rate == 0) {
                std::vector<uint8_t> allpar(NC, 1);
                for (size_t i = 0; i < N; ++i) if (!((!eqf[i]) && multi[i])) allpar[cls[i]] = 0;
                for (size_t i = 0; i < N; ++i) if (allpar[cls[i]]) aval[i] = 0.01;
            }
            for (size_t i = 0; i < N; ++i) {
                aval[i] = r_pmax(r_pmin(aval[i], 1.0), 0.0);                                   // pmax(pmin(aval,1),0)
                if (multi[i] && eqf[i]) aval[i] = r_pmin(1.0, r_pmax(aval[i], part[i]));
                if (!multi[i]) aval[i] = 1.0;
            }
        }

        lap("A-values");
        // ---- removeUnObservedGenes (:301-315) ----------------------------------------------------------
        std::vector<ld> gsum(G, 0);
        for (size_t i = 0; i < N; ++i) gsum[gene[i]] += (ld)nobs[i];
        std::vector<uint8_t> keep(N, 1);
        {
            std::vector<uint8_t> seen(NGT, 0);
            for (size_t i = 0; i < N; ++i)
                if ((double)gsum[gene[i]] == 0) {
                    keep[i] = 0;
                    if (!seen[gt[i]]) { seen[gt[i]] = 1; P->unobserved_txid.push_back(txid_in[i]); }
                }
        }
        // ---- K, n.obs (:221-228) -----------------------------------------------------------------------
        std::vector<ld> Kg(G, 0);
        {
            std::vector<uint8_t> seen(NC, 0);
            for (size_t i = 0; i < N; ++i)
                if (keep[i] && !seen[cls[i]]) { seen[cls[i]] = 1; Kg[gene[i]] += (ld)nobs[i]; }
        }
        // initialiseOutput (:322-327): the right_join orders rows by first appearance of (gene_sid, eqClassId);
        // unique txid in that order
        {
            std::vector<uint32_t> rank(NC, UINT32_MAX);
            uint32_t r = 0;
            for (size_t i = 0; i < N; ++i) if (keep[i] && rank[cls[i]] == UINT32_MAX) rank[cls[i]] = r++;
            std::vector<KV> kv;
            kv.reserve(N);
            for (size_t i = 0; i < N; ++i) if (keep[i]) kv.push_back(KV{rank[cls[i]], (uint32_t)i});
            radix_kv(kv, bits_for(r));       // stable: table order inside a class
            Perm p(kv.size());
            for (size_t q = 0; q < kv.size(); ++q) p[q] = kv[q].v;
            std::vector<uint8_t> seen(NT, 0);
            for (uint32_t i : p) if (!seen[tx[i]]) { seen[tx[i]] = 1; P->out_txid.push_back(txid_in[i]); }
        }
        lap("K, n.obs, initialiseOutput");
        // ---- filterTxRc (:332-342) ---------------------------------------------------------------------
        {
            std::vector<ld> tsum(NGT, 0);
            for (size_t i = 0; i < N; ++i) if (keep[i]) tsum[gt[i]] += (ld)nobs[i];
            for (size_t i = 0; i < N; ++i) if (keep[i] && !((double)tsum[gt[i]] > 0)) keep[i] = 0;
        }
        Perm rows;
        rows.reserve(N);
        for (size_t i = 0; i < N; ++i) if (keep[i]) rows.push_back((uint32_t)i);

        lap("filterTxRc");
        // ---- assignGroups (:347-356) -------------------------------------------------------------------
        // distinct txid and distinct eqClassId per gene among the kept rows
        std::vector<uint32_t> n_tx(G, 0), n_eq(G, 0);
        {
            std::vector<uint8_t> seen(NGT, 0), seen_c(NC, 0);
            for (uint32_t i : rows) {
                if (!seen[gt[i]]) { seen[gt[i]] = 1; ++n_tx[gene[i]]; }
                if (!seen_c[cls[i]]) { seen_c[cls[i]] = 1; ++n_eq[gene[i]]; }
            }
        }
        std::vector<uint32_t> gi_of_gene;
        const std::vector<int64_t> gids = qp::assign_groups(n_tx, n_eq, gi_of_gene);
        lap("assignGroups");
        // ---- getInputList + getAMat (:359-371, 427-431) as one arena --------------------------------
        const uint32_t NG = (uint32_t)gids.size();
        std::vector<uint32_t> gi(N, 0);
        for (uint32_t i : rows) gi[i] = gi_of_gene[gene[i]];
        // arena rows: unique (group, txid) ascending; arena columns: unique (group, gene_sid, eqClassId) ascending
        std::vector<uint32_t> arow(N, 0), acol(N, 0);
        uint32_t n_arow = 0, n_acol = 0;
        const int b_tx = bits_for(NT ? NT - 1 : 0), b_eq = bits_for(NE ? NE - 1 : 0), b_g = bits_for(G ? G - 1 : 0),
                  b_gi = bits_for(NG ? NG - 1 : 0);
        if (b_gi + b_g + b_eq > 64) throw std::string("too many groups x genes x classes for the 64-bit column key");
        {
            std::vector<KV> kv(rows.size());
            for (size_t q = 0; q < rows.size(); ++q) kv[q] = KV{((uint64_t)gi[rows[q]] << b_tx) | tx[rows[q]], rows[q]};
            radix_kv(kv, b_gi + b_tx);
            for (size_t k = 0; k < kv.size(); ++k) {
                if (k && kv[k].k != kv[k - 1].k) ++n_arow;
                arow[kv[k].v] = n_arow;
            }
            if (!kv.empty()) ++n_arow;
            P->txid.resize(n_arow);
            P->grp_row_ptr.assign(NG + 1, 0);
            std::vector<uint8_t> seen(n_arow, 0);
            for (const KV &e : kv) {
                const uint32_t i = e.v;
                if (!seen[arow[i]]) { seen[arow[i]] = 1; P->txid[arow[i]] = (int32_t)tx_values[tx[i]]; ++P->grp_row_ptr[gi[i] + 1]; }
            }
            for (uint32_t g = 0; g < NG; ++g) P->grp_row_ptr[g + 1] += P->grp_row_ptr[g];
        }
        {
            std::vector<KV> kv(rows.size());
            for (size_t q = 0; q < rows.size(); ++q) {
                const uint32_t i = rows[q];
                kv[q] = KV{((((uint64_t)gi[i] << b_g) | gene[i]) << b_eq) | eq[i], i};
            }
            radix_kv(kv, b_gi + b_g + b_eq);
            for (size_t k = 0; k < kv.size(); ++k) {
                if (k && kv[k].k != kv[k - 1].k) ++n_acol;
                acol[kv[k].v] = n_acol;
            }
            if (!kv.empty()) ++n_acol;
            P->Y.assign(n_acol, NAN);
            P->K.assign(n_acol, NAN);
            P->col_flags.assign(n_acol, 0);
            P->grp_col_ptr.assign(NG + 1, 0);
            std::vector<uint8_t> seen(n_acol, 0);
            for (const KV &e : kv) {
                const uint32_t i = e.v, c = acol[i];
                if (!seen[c]) {
                    seen[c] = 1;
                    ++P->grp_col_ptr[gi[i] + 1];
                    const double Kv = (double)Kg[gene[i]];
                    P->K[c] = Kv;
                    P->Y[c] = nobs[i] / Kv;
                    P->col_flags[c] = multi[i] ? BAMBU_COL_MULTI : 0;
                }
            }
            for (uint32_t g = 0; g < NG; ++g) P->grp_col_ptr[g + 1] += P->grp_col_ptr[g];
        }
        {
            const int b_c = bits_for(n_acol ? n_acol - 1 : 0);
            std::vector<KV> kv(rows.size());
            for (size_t q = 0; q < rows.size(); ++q) kv[q] = KV{((uint64_t)arow[rows[q]] << b_c) | acol[rows[q]], rows[q]};
            radix_kv(kv, bits_for(n_arow ? n_arow - 1 : 0) + b_c);
            const size_t E = kv.size();
            P->col_idx.resize(E);
            P->aval.resize(E);
            P->nz_flags.resize(E);
            P->row_nnz_ptr.assign((size_t)n_arow + 1, 0);
            for (size_t k = 0; k < E; ++k) {
                const uint32_t i = kv[k].v;
                if (k && kv[k].k == kv[k - 1].k)
                    throw std::string("duplicated (txid, gene, eqClassId) rows in readClassDt");
                P->col_idx[k] = (int32_t)(acol[i] - P->grp_col_ptr[gi[i]]);
                P->aval[k] = aval[i];
                P->nz_flags[k] = eqf[i] ? BAMBU_NZ_EQUAL : 0;
                ++P->row_nnz_ptr[arow[i] + 1];
            }
            for (uint32_t r = 0; r < n_arow; ++r) P->row_nnz_ptr[r + 1] += P->row_nnz_ptr[r];
        }
        P->gene_grp_id = gids;
        if (P->grp_row_ptr.empty()) { P->grp_row_ptr.assign(1, 0); P->grp_col_ptr.assign(1, 0); }
        if (P->row_nnz_ptr.empty()) P->row_nnz_ptr.assign(1, 0);
    } catch (const std::string &e) {
        g_pack_error = e;
        delete P;
        return BAMBU_EM_EINVAL;
    } catch (const std::bad_alloc &) {
        g_pack_error = "out of memory";
        delete P;
        return BAMBU_EM_ENOMEM;
    } catch (const std::exception &e) {
        g_pack_error = e.what();
        delete P;
        return BAMBU_EM_EINVAL;
    } catch (...) {
        g_pack_error = "unknown exception in the host packer";
        delete P;
        return BAMBU_EM_EINVAL;
    }
    *out = P;
    return BAMBU_EM_OK;
}

int bambu_quant_pack_sizes(const bambu_quant_pack *p, int64_t *n_groups, int64_t *n_rows, int64_t *n_cols, int64_t *nnz,
                           int64_t *n_out, int64_t *n_unobserved, double *d_rate)
{
    if (!p) return BAMBU_EM_EINVAL;
    if (n_groups) *n_groups = (int64_t)p->grp_row_ptr.size() - 1;
    if (n_rows) *n_rows = p->grp_row_ptr.back();
    if (n_cols) *n_cols = p->grp_col_ptr.back();
    if (nnz) *nnz = p->row_nnz_ptr.back();
    if (n_out) *n_out = (int64_t)p->out_txid.size();
    if (n_unobserved) *n_unobserved = (int64_t)p->unobserved_txid.size();
    if (d_rate) *d_rate = p->d_rate;
    return BAMBU_EM_OK;
}

// borrowed pointers into the pack (valid until bambu_quant_pack_free)
int bambu_quant_pack_arena(const bambu_quant_pack *p, const int64_t **grp_row_ptr, const int64_t **grp_col_ptr,
                           const int64_t **row_nnz_ptr, const int32_t **col_idx, const double **aval,
                           const uint8_t **nz_flags, const double **Y, const double **K, const uint8_t **col_flags,
                           const int32_t **txid)
{
    if (!p) return BAMBU_EM_EINVAL;
    if (grp_row_ptr) *grp_row_ptr = p->grp_row_ptr.data();
    if (grp_col_ptr) *grp_col_ptr = p->grp_col_ptr.data();
    if (row_nnz_ptr) *row_nnz_ptr = p->row_nnz_ptr.data();
    if (col_idx) *col_idx = p->col_idx.data();
    if (aval) *aval = p->aval.data();
    if (nz_flags) *nz_flags = p->nz_flags.data();
    if (Y) *Y = p->Y.data();
    if (K) *K = p->K.data();
    if (col_flags) *col_flags = p->col_flags.data();
    if (txid) *txid = p->txid.data();
    return BAMBU_EM_OK;
}

int bambu_quant_pack_bookkeeping(const bambu_quant_pack *p, const int64_t **out_txid, const int64_t **unobserved_txid,
                                 const int64_t **gene_grp_id)
{
    if (!p) return BAMBU_EM_EINVAL;
    if (out_txid) *out_txid = p->out_txid.data();
    if (unobserved_txid) *unobserved_txid = p->unobserved_txid.data();
    if (gene_grp_id) *gene_grp_id = p->gene_grp_id.data();
    return BAMBU_EM_OK;
}

// modifyQuantOut (positional; a txid estimated in several groups keeps the LAST group's value), rbind of
// the unobserved genes' zero rows, removeDuplicates (sum by txid, first-appearance order)
// (R/bambu-quantify_utilityFunctions.R:438-455, R/bambu-quantify.R:70-72).
// est = 3 x n_rows as bambu_em_batched returns it.  Outputs: n_result rows (<= n_unobserved + n_out).
int bambu_quant_merge(const bambu_quant_pack *p, const double *est, int64_t *txid_out, double *counts,
                      double *full_length, double *unique_counts, int64_t *n_result)
{
    if (!p || !txid_out || !counts || !full_length || !unique_counts || !n_result) return BAMBU_EM_EINVAL;
    const size_t n_out = p->out_txid.size(), n_rows = (size_t)p->grp_row_ptr.back();
    if (n_rows && !est) return BAMBU_EM_EINVAL;
    // txid -> position lookups: a direct table when the txid range is small (the usual case: txids are row
    // numbers of the annotation), a hash map otherwise
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t t : p->out_txid) { lo = std::min(lo, t); hi = std::max(hi, t); }
    for (int64_t t : p->unobserved_txid) { lo = std::min(lo, t); hi = std::max(hi, t); }
    const size_t n_all = n_out + p->unobserved_txid.size();
    const bool direct = n_all && (uint64_t)hi - (uint64_t)lo < (uint64_t)(8 * n_all + (1u << 20));
    std::vector<uint32_t> tab;
    std::unordered_map<int64_t, uint32_t> map;
    auto reset_index = [&]() {
        if (direct) tab.assign((size_t)((uint64_t)hi - (uint64_t)lo) + 1, UINT32_MAX);
        else { map.clear(); map.reserve(n_all * 2 + 16); }
    };
    auto find = [&](int64_t t) -> uint32_t {          // UINT32_MAX = absent
        if (direct) return (t < lo || t > hi) ? UINT32_MAX : tab[(size_t)((uint64_t)t - (uint64_t)lo)];
        auto it = map.find(t);
        return it == map.end() ? UINT32_MAX : it->second;
    };
    auto put = [&](int64_t t, uint32_t v) {
        if (direct) tab[(size_t)((uint64_t)t - (uint64_t)lo)] = v;
        else map.emplace(t, v);
    };
    reset_index();
    for (size_t q = 0; q < n_out; ++q) put(p->out_txid[q], (uint32_t)q);
    std::vector<double> v0(n_out, 0.0), v1(n_out, 0.0), v2(n_out, 0.0);
    for (size_t r = 0; r < n_rows; ++r) {          // arena row order = rbind order of the groups: later rows win
        const uint32_t q = find((int64_t)p->txid[r]);
        if (q == UINT32_MAX) continue;
        v0[q] = est[r]; v1[q] = est[n_rows + r]; v2[q] = est[2 * n_rows + r];
    }
    reset_index();
    // removeDuplicates' sum(): out_txid holds every txid once and the unobserved rows are zeros, so a slot receives
    // zeros and at most one value -- 0 + v is exact in any precision, the double accumulators below ARE R's
    // long double sums rounded back.
    int64_t k = 0;
    auto add = [&](int64_t t, double a, double b, double c) {
        uint32_t q = find(t);
        if (q == UINT32_MAX) {
            q = (uint32_t)k;
            put(t, q);
            txid_out[k++] = t;
            counts[q] = 0.0; full_length[q] = 0.0; unique_counts[q] = 0.0;
        }
        counts[q] += a; full_length[q] += b; unique_counts[q] += c;
    };
    for (int64_t t : p->unobserved_txid) add(t, 0.0, 0.0, 0.0);
    for (size_t q = 0; q < n_out; ++q) add(p->out_txid[q], v0[q], v1[q], v2[q]);
    *n_result = k;
    return BAMBU_EM_OK;
}

// bambu.quantDT in one call with the HOST packer: long table in, estimates by txid out
// (pack -> bambu_em_batched -> merge).  bambu_quantDT (quant_dev.cu) packs on the GPU and comes here for the
// tables its device formulation does not cover.
// Output arrays must hold at least as many rows as there are distinct txids in the input (<= n).
int bambu_quantDT_hostpack(int64_t n, const int64_t *gene_code, const int64_t *txid, const int64_t *eqClassId, const double *nobs,
                  const uint8_t *equal, const double *rcWidth, const double *txlen, int32_t degradation_bias,
                  int32_t maxiter, double minvalue, double conv, int64_t *txid_out, double *counts, double *full_length,
                  double *unique_counts, int64_t *n_result, double *d_rate_out)
{
    bambu_quant_pack *p = nullptr;
    int rc = bambu_quant_pack_create(&p, n, gene_code, txid, eqClassId, nobs, equal, rcWidth, txlen, degradation_bias);
    if (rc) return rc;
    const int64_t ng = (int64_t)p->grp_row_ptr.size() - 1, rows = p->grp_row_ptr.back();
    std::vector<double> est((size_t)std::max<int64_t>(3 * rows, 1));
    std::vector<int32_t> iters((size_t)std::max<int64_t>(ng, 1)), status((size_t)std::max<int64_t>(ng, 1));
    // keep payload pointers non-NULL for degenerate tables
    static const double zd = 0; static const int32_t zi = 0; static const uint8_t zb = 0;
    if (ng)
        rc = bambu_em_batched(ng, p->grp_row_ptr.data(), p->grp_col_ptr.data(), p->row_nnz_ptr.data(),
                              p->col_idx.empty() ? &zi : p->col_idx.data(), p->aval.empty() ? &zd : p->aval.data(),
                              p->nz_flags.empty() ? &zb : p->nz_flags.data(), p->Y.empty() ? &zd : p->Y.data(),
                              p->K.empty() ? &zd : p->K.data(), p->col_flags.empty() ? &zb : p->col_flags.data(), maxiter,
                              minvalue, conv, est.data(), iters.data(), status.data());
    if (rc) g_pack_error = bambu_em_last_error();
    if (!rc) rc = bambu_quant_merge(p, est.data(), txid_out, counts, full_length, unique_counts, n_result);
    if (!rc && d_rate_out) *d_rate_out = p->d_rate;
    bambu_quant_pack_free(p);
    return rc;
}

}  // extern "C"
