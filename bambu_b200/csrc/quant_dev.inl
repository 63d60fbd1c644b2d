// quant_dev.inl -- the long table -> arena pipeline of bambu.quantDT (R/bambu-quantify.R:53-62) written
// as data-parallel passes: radix sorts on dense integer keys, run-length ids, scans, and flat loops
// whose bodies only use order-free atomics (integer add/min/max, adds of integer-valued doubles).
//
// The same step-by-step semantics as csrc/quant_pack.cpp (which follows the R code line by line and is
// what the tests compare against, byte for byte); here every "first occurrence in table order" is the
// minimum row index of a run of a stable sort, and the one order-sensitive reduction of the reference
// -- R's long double sum() of the partial a-values of a transcript (:280) -- is a sequential loop over
// the transcript's rows in table order with the x87 extended addition done in integer arithmetic.
//
// The pipeline is a template over an executor X that provides memory, parallel loops, a stable LSD
// radix sort of (u64 key, u32 payload) pairs, an exclusive scan and host<->device copies:
//   quant_dev.cu            CudaExec  -- the product: kernels for sm_100a, arena left in HBM
//   quant_dev_hostexec.cpp  HostExec  -- the same passes run serially on the host; built only by the
//                                        CPU tests to pin the pipeline logic where no GPU exists
// Inputs the parallel formulation does not reproduce exactly (non-integer or negative nobs, negative
// or non-finite widths / lengths / degradation rate) are reported as kUnsupported and the caller uses
// the host packer for them.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "quant_pack.h"
#include "x87_sum.h"

namespace qd {

enum { kOk = 0, kUnsupported = 1, kInvalid = 2 };
enum { kStUnsupported = 1u, kStTwoNobs = 2u, kStDupRows = 4u };
enum { kColGene = 0, kColTx, kColEq, kColNobs, kColEqual, kColTxlen, kColRcWidth, kNumCols };

struct Input {           // device pointers (equal / rcw / txl may be null)
    size_t n;
    const int64_t *gene, *tx, *eq;
    const double *nobs;
    const uint8_t *equal;
    const double *rcw, *txl;
    bool d_mode;
};

struct Result {          // arena in executor memory + the small bookkeeping lists on the host
    uint32_t n_groups = 0, n_rows = 0, n_cols = 0;
    size_t n_entries = 0;
    int64_t *grp_row_ptr = nullptr, *grp_col_ptr = nullptr, *row_nnz_ptr = nullptr;
    int32_t *col_idx = nullptr, *txid = nullptr;
    double *aval = nullptr, *Y = nullptr, *K = nullptr;
    uint8_t *nz_flags = nullptr, *col_flags = nullptr;
    std::vector<int64_t> out_txid, unobserved_txid, gene_grp_id;
    double d_rate = NAN;
};

// R's pmax / pmin without na.rm: NA if either argument is NA
QD_FN double r_pmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }
QD_FN double r_pmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }

// order-preserving map of non-NaN doubles onto u64 (for an atomic max)
QD_FN uint64_t ordered_bits(double d)
{
    uint64_t b;
    memcpy(&b, &d, 8);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
QD_FN double from_ordered_bits(uint64_t u)
{
    const uint64_t b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// Scratch every sort shares.
template <class X>
struct Scratch {
    uint64_t *k, *k2;
    uint32_t *v, *v2, *flag, *run;
};

// after a sort of n keys: run[j] = id of the run of equal keys position j belongs to; returns the number of runs
template <class X>
uint32_t run_ids(X &x, Scratch<X> &s, size_t n)
{
    if (!n) return 0;
    const uint64_t *k = s.k;
    uint32_t *flag = s.flag;
    x.pfor(n, QD_LAMBDA(size_t j) { flag[j] = (j + 1 < n && k[j + 1] != k[j]) ? 1u : 0u; });
    return x.scan(s.flag, s.run, n) + 1;
}

// dense ids of a 64-bit column in sorted order of the values (order preserving).  With first_appearance
// the ids are renumbered like R's match(x, unique(x)).  values (optional): the distinct values, by id.
template <class X>
uint32_t dense_ids(X &x, Scratch<X> &s, const int64_t *col, size_t n, int64_t lo, int64_t hi, bool first_appearance,
                   uint32_t *id, int64_t **values)
{
    const int bits = qp::bits_for((uint64_t)hi - (uint64_t)lo);
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(n, QD_LAMBDA(size_t i) { k[i] = (uint64_t)col[i] - (uint64_t)lo; v[i] = (uint32_t)i; });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, n, bits);
    const uint32_t count = run_ids(x, s, n);
    const uint64_t *k = s.k;
    const uint32_t *v = s.v, *run = s.run;
    if (values) {
        int64_t *val = x.template alloc<int64_t>(count);
        *values = val;
        x.pfor(n, QD_LAMBDA(size_t j) { if (j == 0 || k[j] != k[j - 1]) val[run[j]] = (int64_t)(k[j] + (uint64_t)lo); });
    }
    if (!first_appearance) {
        x.pfor(n, QD_LAMBDA(size_t j) { id[v[j]] = run[j]; });
        return count;
    }
    // rank the runs by their first row (the sort is stable, so the head of a run is its smallest row)
    const size_t mk = x.mark();
    uint32_t *run_of = x.template alloc<uint32_t>(n);
    uint32_t *rank = x.template alloc<uint32_t>(count);
    uint64_t *fk = x.template alloc<uint64_t>(count), *fk2 = x.template alloc<uint64_t>(count);
    uint32_t *fv = x.template alloc<uint32_t>(count), *fv2 = x.template alloc<uint32_t>(count);
    x.pfor(n, QD_LAMBDA(size_t j) {
        run_of[v[j]] = run[j];
        if (j == 0 || k[j] != k[j - 1]) { fk[run[j]] = v[j]; fv[run[j]] = run[j]; }
    });
    x.sort_pairs(fk, fv, fk2, fv2, count, qp::bits_for(n));
    {
        const uint32_t *sv = fv;
        x.pfor(count, QD_LAMBDA(size_t q) { rank[sv[q]] = (uint32_t)q; });
    }
    x.pfor(n, QD_LAMBDA(size_t i) { id[i] = rank[run_of[i]]; });
    x.release(mk);
    return count;
}

template <class X>
int pack(X &x, const Input &in, Result &R, std::string &err)
{
    const size_t N = in.n;
    R = Result();
    if (N == 0) {
        R.grp_row_ptr = x.template alloc<int64_t>(1);
        R.grp_col_ptr = x.template alloc<int64_t>(1);
        R.row_nnz_ptr = x.template alloc<int64_t>(1);
        x.fill(R.grp_row_ptr, 0, 8); x.fill(R.grp_col_ptr, 0, 8); x.fill(R.row_nnz_ptr, 0, 8);
        return kOk;
    }
    const int64_t *gene_in = in.gene, *tx_in = in.tx, *eq_in = in.eq;
    const double *nobs = in.nobs, *rcw = in.rcw, *txl = in.txl;
    const uint8_t *equal = in.equal;
    const bool d_mode = in.d_mode;
    const int bN = qp::bits_for(N);

    Scratch<X> s;
    s.k = x.template alloc<uint64_t>(N); s.k2 = x.template alloc<uint64_t>(N);
    s.v = x.template alloc<uint32_t>(N); s.v2 = x.template alloc<uint32_t>(N);
    s.flag = x.template alloc<uint32_t>(N); s.run = x.template alloc<uint32_t>(N);
    uint32_t *status = x.template alloc<uint32_t>(1);
    x.fill(status, 0, 4);

    // The executor may still be copying the table in (column by column, in the order gene, txid, eqClassId,
    // nobs, equal, txlen, rcWidth): x.need(c) orders the work that follows behind column c's arrival.
    // ---- value ranges of the id columns ------------------------------------------------------------------
    int64_t mm[6];
    {
        int64_t *d_mm = x.template alloc<int64_t>(6);
        x.need(kColGene); x.minmax(gene_in, N, d_mm);
        x.need(kColTx); x.minmax(tx_in, N, d_mm + 2);
        x.need(kColEq); x.minmax(eq_in, N, d_mm + 4);
        x.d2h(mm, d_mm, 6);
    }
    x.lap("minmax");
    // the arena carries txid as int32 (R's integer): anything wider would be truncated and merged under a wrong id
    if (N && (mm[2] < INT32_MIN || mm[3] > INT32_MAX)) { err = "txid outside the 32-bit integer range"; return kInvalid; }
    // ---- simplifyNames (:236-241) + dense order-preserving codes for txid / eqClassId -----------
    uint32_t *gene = x.template alloc<uint32_t>(N), *tx = x.template alloc<uint32_t>(N), *eq = x.template alloc<uint32_t>(N);
    int64_t *tx_values = nullptr;
    const uint32_t G = dense_ids(x, s, gene_in, N, mm[0], mm[1], true, gene, nullptr);
    const uint32_t NT = dense_ids(x, s, tx_in, N, mm[2], mm[3], false, tx, &tx_values);
    const uint32_t NE = dense_ids(x, s, eq_in, N, mm[4], mm[5], false, eq, nullptr);
    const int b_g = qp::bits_for(G - 1), b_tx = qp::bits_for(NT - 1), b_eq = qp::bits_for(NE - 1);

    x.lap("dense ids");
    // ---- what the parallel formulation relies on: counts are non-negative integers whose sums stay exact ----
    uint8_t *eqf = x.template alloc<uint8_t>(N);
    {
        const double nobs_limit = 9007199254740992.0 / (double)N;
        x.need(kColNobs); x.need(kColEqual);
        x.pfor(N, QD_LAMBDA(size_t i) {
            const double y = nobs[i];
            if (!(y >= 0.0) || !(y <= nobs_limit) || y != (double)(int64_t)y) qd_or(status, kStUnsupported);
            eqf[i] = (equal && equal[i]) ? 1 : 0;
        });
    }
    // ---- (gene, eqClass) classes: id, first row; a class carries one nobs value -----------------
    uint32_t *cls = x.template alloc<uint32_t>(N);
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(N, QD_LAMBDA(size_t i) { k[i] = ((uint64_t)gene[i] << b_eq) | eq[i]; v[i] = (uint32_t)i; });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, N, b_g + b_eq);
    const uint32_t NC = run_ids(x, s, N);
    uint32_t *cls_first = x.template alloc<uint32_t>(NC);
    {
        const uint64_t *k = s.k;
        const uint32_t *v = s.v, *run = s.run;
        x.pfor(N, QD_LAMBDA(size_t j) {
            cls[v[j]] = run[j];
            if (j == 0 || k[j] != k[j - 1]) cls_first[run[j]] = v[j];
        });
        x.pfor(N, QD_LAMBDA(size_t j) { if (nobs[v[j]] != nobs[cls_first[run[j]]]) qd_or(status, kStTwoNobs); });
    }
    x.lap("classes");
    // ---- (gene, tx) pairs: id, rows of each pair in table order ---------------------------------------
    uint32_t *gt = x.template alloc<uint32_t>(N), *gt_perm = x.template alloc<uint32_t>(N);
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(N, QD_LAMBDA(size_t i) { k[i] = ((uint64_t)gene[i] << b_tx) | tx[i]; v[i] = (uint32_t)i; });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, N, b_g + b_tx);
    const uint32_t NGT = run_ids(x, s, N);
    uint32_t *gt_start = x.template alloc<uint32_t>((size_t)NGT + 1);
    {
        const uint64_t *k = s.k;
        const uint32_t *v = s.v, *run = s.run;
        x.pfor(N, QD_LAMBDA(size_t j) {
            gt[v[j]] = run[j];
            gt_perm[j] = v[j];
            if (j == 0 || k[j] != k[j - 1]) gt_start[run[j]] = (uint32_t)j;
            if (j == N - 1) gt_start[run[j] + 1] = (uint32_t)N;
        });
    }

    x.lap("gene-tx pairs");
    // ---- per-gene sums over the distinct classes (calculateDegradationRate :247-257; K :221-228) ----
    uint8_t *cls_par = x.template alloc<uint8_t>(NC);
    double *gnobs = x.template alloc<double>(G), *gdobs = x.template alloc<double>(G);
    uint8_t *has_par = x.template alloc<uint8_t>(G);
    uint64_t *glen = x.template alloc<uint64_t>(G);
    x.fill(cls_par, 0, NC); x.fill(gnobs, 0, 8 * (size_t)G); x.fill(gdobs, 0, 8 * (size_t)G); x.fill(has_par, 0, G);
    x.pfor(N, QD_LAMBDA(size_t i) { if (!eqf[i]) cls_par[cls[i]] = 1; });
    x.pfor(NC, QD_LAMBDA(size_t c) {
        const uint32_t r = cls_first[c], g = gene[r];
        qd_add(gnobs + g, nobs[r]);
        if (cls_par[c]) { qd_add(gdobs + g, nobs[r]); has_par[g] = 1; }
    });
    double d_rate = NAN;
    if (d_mode) {
        x.need(kColTxlen); x.need(kColRcWidth);
        x.pfor(N, QD_LAMBDA(size_t i) {      // lengths > 0 and widths >= 0, finite
            bool bad = false;
            if (rcw) { const double w = rcw[i]; bad |= !(w >= 0.0) || !(w <= 1.7e308); }
            if (txl) { const double l = txl[i]; bad |= !(l > 0.0) || !(l <= 1.7e308); }
            if (bad) qd_or(status, kStUnsupported);
        });
        const uint64_t neg_inf = ordered_bits(-INFINITY);
        x.pfor(G, QD_LAMBDA(size_t g) { glen[g] = neg_inf; });
        x.pfor(N, QD_LAMBDA(size_t i) { qd_max(glen + gene[i], ordered_bits(txl ? txl[i] : 1.0)); });
        std::vector<double> h_nb(G), h_db(G), h_len(G);
        std::vector<uint8_t> h_par(G);
        std::vector<uint64_t> h_lenb(G);
        x.d2h(h_nb.data(), gnobs, G); x.d2h(h_db.data(), gdobs, G); x.d2h(h_par.data(), has_par, G); x.d2h(h_lenb.data(), glen, G);
        for (uint32_t g = 0; g < G; ++g) h_len[g] = from_ordered_bits(h_lenb[g]);
        uint32_t st;
        x.d2h(&st, status, 1);
        if (st & kStUnsupported) return kUnsupported;
        d_rate = qp::degradation_rate(G, h_nb.data(), h_db.data(), h_par.data(), h_len.data());
        if (!(d_rate >= 0.0) || !(d_rate <= 1.7e308)) return kUnsupported;
    }
    R.d_rate = d_rate;

    x.lap("gene sums + degradation rate");
    // ---- multi_align: classes with more than one distinct transcript (:276-279) -----------------------
    uint32_t *ntx = x.template alloc<uint32_t>(NC);
    x.fill(ntx, 0, 4 * (size_t)NC);
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(N, QD_LAMBDA(size_t i) { k[i] = ((uint64_t)cls[i] << b_tx) | tx[i]; v[i] = (uint32_t)i; });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, N, qp::bits_for(NC - 1) + b_tx);
    {
        const uint64_t *k = s.k;
        x.pfor(N, QD_LAMBDA(size_t j) { if (j == 0 || k[j] != k[j - 1]) qd_inc(ntx + (uint32_t)(k[j] >> b_tx)); });
    }

    x.lap("multi_align");
    // ---- per (gene, tx): nobs total (filterTxRc :333-334) and the long double sum of the partial a-values
    double *psum = x.template alloc<double>(NGT), *tsum = x.template alloc<double>(NGT);
    x.pfor(NGT, QD_LAMBDA(size_t p) {
        Ext acc;
        acc.m = 0; acc.e = 0;
        double ts = 0.0;
        for (uint32_t j = gt_start[p]; j < gt_start[p + 1]; ++j) {
            const uint32_t i = gt_perm[j];
            ts += nobs[i];
            if (d_mode && ntx[cls[i]] > 1 && !eqf[i]) acc = ext_add(acc, ext_from_double((rcw ? rcw[i] : 0.0) * d_rate / 1000));
        }
        bool oor = false;
        psum[p] = ext_to_double(acc, &oor);
        tsum[p] = ts;
        if (oor) qd_or(status, kStUnsupported);
    });
    // ---- modifyAvaluewithDegradation_rate (:280-296) -----------------------------------------------------
    double *aval_row = x.template alloc<double>(N);
    if (!d_mode) {
        x.pfor(N, QD_LAMBDA(size_t i) { aval_row[i] = 1.0; });
    } else {
        uint8_t *allpar = nullptr;
        if (d_rate == 0) {
            allpar = x.template alloc<uint8_t>(NC);
            x.fill(allpar, 1, NC);
            x.pfor(N, QD_LAMBDA(size_t i) { if (!(!eqf[i] && ntx[cls[i]] > 1)) allpar[cls[i]] = 0; });
        }
        x.pfor(N, QD_LAMBDA(size_t i) {
            const bool multi = ntx[cls[i]] > 1;
            const double part = (rcw ? rcw[i] : 0.0) * d_rate / 1000;
            double a = multi ? (eqf[i] ? 1 - psum[gt[i]] : part) : NAN;
            if (allpar && allpar[cls[i]]) a = 0.01;
            a = r_pmax(r_pmin(a, 1.0), 0.0);
            if (multi && eqf[i]) a = r_pmin(1.0, r_pmax(a, part));
            if (!multi) a = 1.0;
            aval_row[i] = a;
        });
    }

    x.lap("a-values");
    // ---- removeUnObservedGenes (:301-315): genes whose nobs are all zero; their txids in first-appearance order
    {
        const size_t mk = x.mark();
        uint32_t *pos = x.template alloc<uint32_t>(NGT), *fl = x.template alloc<uint32_t>(NGT);
        x.pfor(NGT, QD_LAMBDA(size_t p) { fl[p] = gnobs[gene[gt_perm[gt_start[p]]]] == 0 ? 1u : 0u; });
        const uint32_t nu = x.scan(fl, pos, NGT);
        if (nu) {
            uint64_t *k = s.k;
            uint32_t *v = s.v;
            x.pfor(NGT, QD_LAMBDA(size_t p) { if (fl[p]) { k[pos[p]] = gt_perm[gt_start[p]]; v[pos[p]] = 0; } });
            x.sort_pairs(s.k, s.v, s.k2, s.v2, nu, bN);
            int64_t *o = x.template alloc<int64_t>(nu);
            const uint64_t *ks = s.k;
            x.pfor(nu, QD_LAMBDA(size_t q) { o[q] = tx_in[ks[q]]; });
            R.unobserved_txid.resize(nu);
            x.d2h(R.unobserved_txid.data(), o, nu);
        }
        x.release(mk);
    }
    x.lap("unobserved");
    // ---- initialiseOutput (:322-327): txids of the observed genes, ordered by the first appearance of their
    //      class and by table order inside it ------------------------------------------------------------
    {
        const size_t mk = x.mark();
        uint64_t *txkey = x.template alloc<uint64_t>(NT);
        uint32_t *fl = x.template alloc<uint32_t>(NT), *pos = x.template alloc<uint32_t>(NT);
        x.fill(txkey, 0xff, 8 * (size_t)NT);
        x.pfor(N, QD_LAMBDA(size_t i) {
            if (gnobs[gene[i]] != 0) qd_min(txkey + tx[i], ((uint64_t)cls_first[cls[i]] << bN) | (uint64_t)i);
        });
        x.pfor(NT, QD_LAMBDA(size_t t) { fl[t] = txkey[t] != ~0ull ? 1u : 0u; });
        const uint32_t no = x.scan(fl, pos, NT);
        if (no) {
            uint64_t *k = s.k;
            uint32_t *v = s.v;
            x.pfor(NT, QD_LAMBDA(size_t t) { if (fl[t]) { k[pos[t]] = txkey[t]; v[pos[t]] = 0; } });
            x.sort_pairs(s.k, s.v, s.k2, s.v2, no, 2 * bN);
            int64_t *o = x.template alloc<int64_t>(no);
            const uint64_t *ks = s.k;
            const uint64_t row_mask = (bN >= 64) ? ~0ull : ((1ull << bN) - 1);
            x.pfor(no, QD_LAMBDA(size_t q) { o[q] = tx_in[ks[q] & row_mask]; });
            R.out_txid.resize(no);
            x.d2h(R.out_txid.data(), o, no);
        }
        x.release(mk);
    }

    x.lap("initialiseOutput");
    // ---- filterTxRc (:332-342): rows of observed genes whose (gene, tx) has nobs > 0 -------------------
    uint8_t *keep_gt = x.template alloc<uint8_t>(NGT);
    x.pfor(NGT, QD_LAMBDA(size_t p) { keep_gt[p] = (gnobs[gene[gt_perm[gt_start[p]]]] != 0 && tsum[p] > 0) ? 1 : 0; });
    uint32_t *rows = nullptr;
    size_t E = 0;
    {
        uint32_t *fl = s.flag, *pos = s.run;
        x.pfor(N, QD_LAMBDA(size_t i) { fl[i] = keep_gt[gt[i]]; });
        E = x.scan(fl, pos, N);
        rows = x.template alloc<uint32_t>(E ? E : 1);
        uint32_t *rw = rows;
        x.pfor(N, QD_LAMBDA(size_t i) { if (fl[i]) rw[pos[i]] = (uint32_t)i; });
    }
    x.lap("filterTxRc");
    // ---- assignGroups (:347-356): distinct txids x distinct classes per gene among the kept rows -------
    std::vector<uint32_t> h_ntx(G), h_neq(G), h_gi;
    {
        const size_t mk = x.mark();
        uint32_t *n_tx = x.template alloc<uint32_t>(G), *n_eq = x.template alloc<uint32_t>(G);
        uint8_t *cls_keep = x.template alloc<uint8_t>(NC);
        x.fill(n_tx, 0, 4 * (size_t)G); x.fill(n_eq, 0, 4 * (size_t)G); x.fill(cls_keep, 0, NC);
        x.pfor(NGT, QD_LAMBDA(size_t p) { if (keep_gt[p]) qd_inc(n_tx + gene[gt_perm[gt_start[p]]]); });
        {
            const uint32_t *rw = rows;
            x.pfor(E, QD_LAMBDA(size_t q) { cls_keep[cls[rw[q]]] = 1; });
        }
        x.pfor(NC, QD_LAMBDA(size_t c) { if (cls_keep[c]) qd_inc(n_eq + gene[cls_first[c]]); });
        x.d2h(h_ntx.data(), n_tx, G); x.d2h(h_neq.data(), n_eq, G);
        x.release(mk);
    }
    R.gene_grp_id = qp::assign_groups(h_ntx, h_neq, h_gi);
    const uint32_t NG = (uint32_t)R.gene_grp_id.size();
    const int b_gi = qp::bits_for(NG ? NG - 1 : 0);
    if (b_gi + b_g + b_eq > 64) { err = "too many groups x genes x classes for the 64-bit column key"; return kInvalid; }
    uint32_t *gi_of_gene = x.template alloc<uint32_t>(G);
    x.h2d(gi_of_gene, h_gi.data(), G);

    x.lap("assignGroups");
    // ---- getInputList + getAMat (:359-371, 427-431) as one arena -----------------------------------------
    R.n_groups = NG;
    R.grp_row_ptr = x.template alloc<int64_t>((size_t)NG + 1);
    R.grp_col_ptr = x.template alloc<int64_t>((size_t)NG + 1);
    x.fill(R.grp_row_ptr, 0, 8 * ((size_t)NG + 1)); x.fill(R.grp_col_ptr, 0, 8 * ((size_t)NG + 1));
    if (E == 0) {
        R.row_nnz_ptr = x.template alloc<int64_t>(1);
        x.fill(R.row_nnz_ptr, 0, 8);
        uint32_t st;
        x.d2h(&st, status, 1);
        if (st & kStUnsupported) return kUnsupported;
        if (st & kStTwoNobs) { err = "an equivalence class carries more than one nobs value"; return kInvalid; }
        return kOk;
    }
    uint32_t *arow = x.template alloc<uint32_t>(N), *acol = x.template alloc<uint32_t>(N);
    const uint32_t *rw = rows;
    // rows: unique (group, txid) ascending
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(E, QD_LAMBDA(size_t q) { const uint32_t i = rw[q]; k[q] = ((uint64_t)gi_of_gene[gene[i]] << b_tx) | tx[i]; v[q] = i; });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, E, b_gi + b_tx);
    const uint32_t n_arow = run_ids(x, s, E);
    R.n_rows = n_arow;
    R.txid = x.template alloc<int32_t>(n_arow);
    {
        const uint64_t *k = s.k;
        const uint32_t *v = s.v, *run = s.run;
        int32_t *txid = R.txid;
        int64_t *grp = R.grp_row_ptr;
        x.pfor(E, QD_LAMBDA(size_t j) {
            arow[v[j]] = run[j];
            if (j == 0 || k[j] != k[j - 1]) txid[run[j]] = (int32_t)tx_values[tx[v[j]]];
            if (j == 0 || (k[j] >> b_tx) != (k[j - 1] >> b_tx)) grp[k[j] >> b_tx] = run[j];
            if (j == E - 1) grp[NG] = n_arow;
        });
    }
    x.lap("arena rows");
    // columns: unique (group, gene_sid, eqClassId) ascending
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(E, QD_LAMBDA(size_t q) {
            const uint32_t i = rw[q];
            k[q] = ((((uint64_t)gi_of_gene[gene[i]] << b_g) | gene[i]) << b_eq) | eq[i];
            v[q] = i;
        });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, E, b_gi + b_g + b_eq);
    const uint32_t n_acol = run_ids(x, s, E);
    R.n_cols = n_acol;
    R.Y = x.template alloc<double>(n_acol);
    R.K = x.template alloc<double>(n_acol);
    R.col_flags = x.template alloc<uint8_t>(n_acol);
    {
        const uint64_t *k = s.k;
        const uint32_t *v = s.v, *run = s.run;
        double *Y = R.Y, *K = R.K;
        uint8_t *cf = R.col_flags;
        int64_t *grp = R.grp_col_ptr;
        const int sh = b_g + b_eq;
        x.pfor(E, QD_LAMBDA(size_t j) {
            const uint32_t i = v[j], c = run[j];
            acol[i] = c;
            if (j == 0 || k[j] != k[j - 1]) {
                const double Kv = gnobs[gene[i]];
                K[c] = Kv;
                Y[c] = nobs[i] / Kv;
                cf[c] = ntx[cls[i]] > 1 ? 1 : 0;
            }
            if (j == 0 || (k[j] >> sh) != (k[j - 1] >> sh)) grp[k[j] >> sh] = c;
            if (j == E - 1) grp[NG] = n_acol;
        });
    }
    x.lap("arena columns");
    // entries: by (row, column)
    const int b_c = qp::bits_for(n_acol - 1);
    {
        uint64_t *k = s.k;
        uint32_t *v = s.v;
        x.pfor(E, QD_LAMBDA(size_t q) { const uint32_t i = rw[q]; k[q] = ((uint64_t)arow[i] << b_c) | acol[i]; v[q] = i; });
    }
    x.sort_pairs(s.k, s.v, s.k2, s.v2, E, qp::bits_for(n_arow - 1) + b_c);
    R.n_entries = E;
    R.col_idx = x.template alloc<int32_t>(E);
    R.aval = x.template alloc<double>(E);
    R.nz_flags = x.template alloc<uint8_t>(E);
    R.row_nnz_ptr = x.template alloc<int64_t>((size_t)n_arow + 1);
    {
        const uint64_t *k = s.k;
        const uint32_t *v = s.v;
        int32_t *col_idx = R.col_idx;
        double *aval = R.aval;
        uint8_t *nzf = R.nz_flags;
        int64_t *rnp = R.row_nnz_ptr;
        const int64_t *gcp = R.grp_col_ptr;
        x.pfor(E, QD_LAMBDA(size_t j) {
            const uint32_t i = v[j];
            if (j && k[j] == k[j - 1]) qd_or(status, kStDupRows);
            col_idx[j] = (int32_t)((int64_t)acol[i] - gcp[gi_of_gene[gene[i]]]);
            aval[j] = aval_row[i];
            nzf[j] = eqf[i] ? 1 : 0;
            if (j == 0 || (k[j] >> b_c) != (k[j - 1] >> b_c)) rnp[k[j] >> b_c] = (int64_t)j;
            if (j == E - 1) rnp[n_arow] = (int64_t)E;
        });
    }
    x.lap("arena entries");
    uint32_t st;
    x.d2h(&st, status, 1);
    if (st & kStUnsupported) return kUnsupported;
    if (st & kStTwoNobs) { err = "an equivalence class carries more than one nobs value"; return kInvalid; }
    if (st & kStDupRows) { err = "duplicated (txid, gene, eqClassId) rows in readClassDt"; return kInvalid; }
    return kOk;
}

}  // namespace qd
