// equirc.cpp -- equivalence read classes from the distance table: genEquiRCs and helpers
// (R/bambu-quantify_utilityFunctions.R:46-192), native host code.  The producer of the long table
// bambu.quantDT takes.  Same steps, same orders as the Python mirror bambu_b200/equirc.py (which is pinned on the
// reference's bundled chr9 run and on an independent restatement, tests/test_equirc_random.py); everything here works
// on integer columns -- the caller codes GENEID and read-class names as integers.
//
//   processMinEquiRC (:159-181)                two minimal classes per annotated transcript (all members partial,
//                                              minRC = 1; the owner full length, the others partial)
//   genEquiRCsBasedOnObservedReads (:74-101)   compatible rows of the distance table inside the read class's own gene,
//                                              class key = sorted signed txids of the read class (negative = equal)
//   getUniCountPerEquiRC (:118-132)            nobs / rcWidth per class over the DISTINCT (class, widths, count, gene)
//                                              rows (distinct() before sum(): the reference's quirk)
//   createEqClassToTxMapping + addEmptyRC      full_join with the minimal classes, NA -> 0, max per class, distinct(),
//   (:140-152, 185-192)                        eqClassId = cur_group_id() by first appearance
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/bambu_em.h"

namespace {

struct VecHash {
    size_t operator()(const std::vector<int64_t> &v) const
    {
        uint64_t h = 0x9e3779b97f4a7c15ull ^ v.size();
        for (int64_t x : v) { h ^= (uint64_t)x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};

struct Interner {          // key (sorted signed txids) -> dense id, first-appearance order
    std::unordered_map<std::vector<int64_t>, int64_t, VecHash> ids;
    std::vector<std::vector<int64_t>> keys;
    int64_t operator()(const std::vector<int64_t> &k)
    {
        auto it = ids.find(k);
        if (it != ids.end()) return it->second;
        const int64_t id = (int64_t)keys.size();
        ids.emplace(k, id);
        keys.push_back(k);
        return id;
    }
};

struct Key4 {              // (class, gene, txid, equal)
    int64_t c, g, t;
    bool e;
    bool operator==(const Key4 &o) const { return c == o.c && g == o.g && t == o.t && e == o.e; }
};
struct Key4Hash {
    size_t operator()(const Key4 &k) const
    {
        uint64_t h = (uint64_t)k.c * 0x9e3779b97f4a7c15ull;
        h ^= (uint64_t)k.g + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.t + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.e + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

struct Key5 {              // (class, first exon width, total width, read count, gene): getUniCountPerEquiRC's distinct()
    int64_t c, g;
    double few, tw, cnt;
    bool operator==(const Key5 &o) const { return c == o.c && g == o.g && few == o.few && tw == o.tw && cnt == o.cnt; }
};
struct Key5Hash {
    size_t operator()(const Key5 &k) const
    {
        auto bits = [](double d) { uint64_t b; memcpy(&b, &d, 8); return b; };
        uint64_t h = (uint64_t)k.c * 0x9e3779b97f4a7c15ull;
        for (uint64_t x : {(uint64_t)k.g, bits(k.few), bits(k.tw), bits(k.cnt)}) h ^= x + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

struct MinRow { Key4 k; int64_t minRC; };      // minRC: 1, or 0 for "absent" (NA -> 0 after the join)

}  // namespace

struct bambu_equirc {
    std::vector<int64_t> gene, txid, eqClassId, minRC;
    std::vector<double> nobs, rcWidth, txlen;
    std::vector<uint8_t> equal;
};

static thread_local std::string g_equirc_error;

extern "C" {

const char *bambu_equirc_last_error(void) { return g_equirc_error.c_str(); }

int bambu_equirc_free(bambu_equirc *p)
{
    delete p;
    return BAMBU_EM_OK;
}

int bambu_equirc_create(bambu_equirc **out, int64_t n_tx, const int64_t *ann_txid, const int64_t *ann_gene,
                        const int64_t *eqcls_ptr, const int64_t *eqcls_val, const double *ann_txlen, int64_t n_rc,
                        const int64_t *rc_gene, const double *rc_first_exon_width, const double *rc_total_width,
                        int64_t n_dt, const int64_t *dt_rc, const int64_t *dt_txid, const double *dt_read_count,
                        const uint8_t *dt_equal, const int64_t *dt_gene, const uint8_t *dt_unidentified)
{
    if (!out) { g_equirc_error = "NULL output"; return BAMBU_EM_EINVAL; }
    *out = nullptr;
    if (n_tx < 0 || n_rc < 0 || n_dt < 0 || (n_tx && (!ann_txid || !ann_gene || !eqcls_ptr || !ann_txlen)) ||
        (n_rc && (!rc_gene || !rc_first_exon_width || !rc_total_width)) ||
        (n_dt && (!dt_rc || !dt_txid || !dt_read_count || !dt_equal || !dt_gene))) {
        g_equirc_error = "NULL or negative argument";
        return BAMBU_EM_EINVAL;
    }
    bambu_equirc *P = new (std::nothrow) bambu_equirc;
    if (!P) { g_equirc_error = "out of memory"; return BAMBU_EM_ENOMEM; }
    try {
        // ---- processMinEquiRC + transcript lengths ----------------------------------------------------------------
        Interner intern;
        std::unordered_map<int64_t, double> txlen;
        txlen.reserve((size_t)n_tx * 2 + 16);
        std::vector<MinRow> ann_rows;                               // bind_rows order
        for (int64_t i = 0; i < n_tx; ++i) {
            if (eqcls_ptr[i + 1] < eqcls_ptr[i]) throw std::string("eqClassById pointers decrease");
            txlen[ann_txid[i]] = ann_txlen[i];
            std::vector<int64_t> cls(eqcls_val + eqcls_ptr[i], eqcls_val + eqcls_ptr[i + 1]);
            const int64_t cid = intern(cls);                        // annotation keys are stored sorted (unsigned = all partial)
            for (int64_t m : cls) ann_rows.push_back(MinRow{Key4{cid, ann_gene[i], m, false}, 1});
        }
        {
            std::unordered_set<Key4, Key4Hash> seen;
            for (int64_t i = 0; i < n_tx; ++i) {
                std::vector<int64_t> cls(eqcls_val + eqcls_ptr[i], eqcls_val + eqcls_ptr[i + 1]);
                std::vector<int64_t> key(cls.size());
                for (size_t q = 0; q < cls.size(); ++q) key[q] = cls[q] == ann_txid[i] ? -cls[q] : cls[q];
                std::sort(key.begin(), key.end());
                const int64_t cid = intern(key);
                for (int64_t m : cls) {
                    const Key4 k{cid, ann_gene[i], m, m == ann_txid[i]};
                    if (seen.insert(k).second) ann_rows.push_back(MinRow{k, 0});      // distinct()
                }
            }
        }
        // ---- genEquiRCsBasedOnObservedReads -------------------------------------------------------------------------
        std::vector<std::vector<int64_t>> by_rc((size_t)n_rc);
        for (int64_t k = 0; k < n_dt; ++k) {
            if (dt_unidentified && dt_unidentified[k]) continue;
            if (dt_rc[k] < 0 || dt_rc[k] >= n_rc) throw std::string("readClassId outside the read-class table");
            by_rc[(size_t)dt_rc[k]].push_back(k);
        }
        struct Obs { int64_t cid, gene; double few, tw, cnt; bool equal; };
        std::vector<Obs> obs;
        for (int64_t rc = 0; rc < n_rc; ++rc) {                     // X[Y] join: order of rowData(readClass)
            std::vector<int64_t> ks;
            for (int64_t k : by_rc[(size_t)rc]) if (dt_gene[k] == rc_gene[rc]) ks.push_back(k);
            if (ks.empty()) continue;
            std::vector<int64_t> key;
            for (int64_t k : ks) key.push_back(dt_equal[k] ? -dt_txid[k] : dt_txid[k]);
            std::sort(key.begin(), key.end());
            const int64_t cid = intern(key);
            for (int64_t k : ks)
                obs.push_back(Obs{cid, dt_gene[k], rc_first_exon_width[rc], rc_total_width[rc], dt_read_count[k], dt_equal[k] != 0});
        }
        // ---- getUniCountPerEquiRC -----------------------------------------------------------------------------------------
        const size_t NK = intern.keys.size();
        std::vector<uint8_t> any_equal(NK, 0), has_width(NK, 0);
        std::vector<double> nobs(NK, 0.0), width(NK, 0.0);
        for (const Obs &o : obs) any_equal[(size_t)o.cid] |= o.equal ? 1 : 0;
        std::vector<Key5> distinct;
        {
            std::unordered_set<Key5, Key5Hash> seen;
            for (const Obs &o : obs) {
                const Key5 k{o.cid, o.gene, o.few, o.tw, o.cnt};
                if (seen.insert(k).second) distinct.push_back(k);
            }
        }
        for (const Key5 &d : distinct) {
            nobs[(size_t)d.c] += d.cnt;
            const double wv = any_equal[(size_t)d.c] ? d.tw : d.few;
            if (!has_width[(size_t)d.c] || wv > width[(size_t)d.c]) { width[(size_t)d.c] = wv; has_width[(size_t)d.c] = 1; }
        }
        struct Count { int64_t c, g; };
        std::vector<Count> counts;
        {
            std::unordered_set<Key4, Key4Hash> seen;
            for (const Key5 &d : distinct)
                if (seen.insert(Key4{d.c, d.g, 0, false}).second) counts.push_back(Count{d.c, d.g});
        }
        // ---- createEqClassToTxMapping, full_join with the minimal classes (addEmptyRC), NA -> 0 ----------------------------
        std::unordered_map<Key4, std::vector<int64_t>, Key4Hash> min_rows;
        min_rows.reserve(ann_rows.size() * 2 + 16);
        for (const MinRow &r : ann_rows) min_rows[r.k].push_back(r.minRC);
        struct Joined { Key4 k; double n, w; int64_t m; };
        std::vector<Joined> joined;
        std::unordered_set<Key4, Key4Hash> used;
        for (const Count &cn : counts) {
            for (int64_t s : intern.keys[(size_t)cn.c]) {
                const Key4 k{cn.c, cn.g, s < 0 ? -s : s, s < 0};
                auto it = min_rows.find(k);
                if (it == min_rows.end()) joined.push_back(Joined{k, nobs[(size_t)cn.c], width[(size_t)cn.c], 0});
                else {
                    used.insert(k);
                    for (int64_t m : it->second) joined.push_back(Joined{k, nobs[(size_t)cn.c], width[(size_t)cn.c], m});
                }
            }
        }
        for (const MinRow &r : ann_rows)                            // unmatched rows of the right table follow
            if (!used.count(r.k)) joined.push_back(Joined{r.k, 0.0, 0.0, r.minRC});
        // group_by(eqClassById): max of nobs, rcWidth, minRC
        const size_t NK2 = intern.keys.size();
        std::vector<double> top_n(NK2, 0.0), top_w(NK2, 0.0);
        std::vector<int64_t> top_m(NK2, 0);
        for (const Joined &j : joined) {
            top_n[(size_t)j.k.c] = std::max(top_n[(size_t)j.k.c], j.n);
            top_w[(size_t)j.k.c] = std::max(top_w[(size_t)j.k.c], j.w);
            top_m[(size_t)j.k.c] = std::max(top_m[(size_t)j.k.c], j.m);
        }
        // distinct(), cur_group_id() by first appearance
        std::vector<int64_t> eq_id(NK2, 0);
        int64_t next_id = 0;
        std::unordered_set<Key4, Key4Hash> seen;
        seen.reserve(joined.size() * 2 + 16);
        for (const Joined &j : joined) {
            if (!seen.insert(j.k).second) continue;
            if (!eq_id[(size_t)j.k.c]) eq_id[(size_t)j.k.c] = ++next_id;
            auto tl = txlen.find(j.k.t);
            if (tl == txlen.end()) throw std::string("a txid of the distance table is not in the annotation");
            P->gene.push_back(j.k.g);
            P->txid.push_back(j.k.t);
            P->eqClassId.push_back(eq_id[(size_t)j.k.c]);
            P->nobs.push_back(top_n[(size_t)j.k.c]);
            P->equal.push_back(j.k.e ? 1 : 0);
            P->rcWidth.push_back(top_w[(size_t)j.k.c]);
            P->txlen.push_back(tl->second);
            P->minRC.push_back(top_m[(size_t)j.k.c]);
        }
    } catch (const std::string &e) {
        g_equirc_error = e;
        delete P;
        return BAMBU_EM_EINVAL;
    } catch (const std::bad_alloc &) {
        g_equirc_error = "out of memory";
        delete P;
        return BAMBU_EM_ENOMEM;
    } catch (...) {
        g_equirc_error = "unknown exception";
        delete P;
        return BAMBU_EM_EINVAL;
    }
    *out = P;
    return BAMBU_EM_OK;
}

int bambu_equirc_size(const bambu_equirc *p, int64_t *n_rows)
{
    if (!p || !n_rows) return BAMBU_EM_EINVAL;
    *n_rows = (int64_t)p->txid.size();
    return BAMBU_EM_OK;
}

int bambu_equirc_table(const bambu_equirc *p, const int64_t **gene, const int64_t **txid, const int64_t **eqClassId,
                       const double **nobs, const uint8_t **equal, const double **rcWidth, const double **txlen,
                       const int64_t **minRC)
{
    if (!p) return BAMBU_EM_EINVAL;
    if (gene) *gene = p->gene.data();
    if (txid) *txid = p->txid.data();
    if (eqClassId) *eqClassId = p->eqClassId.data();
    if (nobs) *nobs = p->nobs.data();
    if (equal) *equal = p->equal.data();
    if (rcWidth) *rcWidth = p->rcWidth.data();
    if (txlen) *txlen = p->txlen.data();
    if (minRC) *minRC = p->minRC.data();
    return BAMBU_EM_OK;
}

}  // extern "C"
