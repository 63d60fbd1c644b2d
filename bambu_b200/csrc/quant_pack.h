// quant_pack.h -- what the host packer (quant_pack.cpp) and the device packer (quant_dev.cu) share:
// the pack object behind the bambu_quant_* accessors and the two small per-gene steps that stay on
// the host in both (the degradation-rate median and the group assignment; G values each).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

struct bambu_quant_pack {
    // the arena
    std::vector<int64_t> grp_row_ptr, grp_col_ptr, row_nnz_ptr;
    std::vector<int32_t> col_idx, txid;
    std::vector<double> aval, Y, K;
    std::vector<uint8_t> nz_flags, col_flags;
    // bookkeeping of initialiseOutput / removeUnObservedGenes / assignGroups
    std::vector<int64_t> out_txid, unobserved_txid, gene_grp_id;
    double d_rate = NAN;
};

namespace qp {

typedef long double ld;

inline int bits_for(uint64_t max_value)      // bits needed to hold values 0..max_value
{
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

inline double r_median(std::vector<double> x)   // median(x, na.rm = TRUE) with mean()'s long double refinement
{
    x.erase(std::remove_if(x.begin(), x.end(), [](double v) { return std::isnan(v); }), x.end());
    const size_t n = x.size();
    if (!n) return NAN;
    std::nth_element(x.begin(), x.begin() + n / 2, x.end());          // selection instead of a full sort
    if (n & 1) return x[n / 2];
    const ld b = x[n / 2], a = *std::max_element(x.begin(), x.begin() + n / 2);
    const ld s = (a + b) / (ld)2;
    const ld t = ((a - s) + (b - s)) / (ld)2;
    return (double)(s + t);
}

// calculateDegradationRate's last step (R/bambu-quantify_utilityFunctions.R:258-269) from the per-gene
// sums: nobs of the gene's classes, nobs of its classes with a partial (non-equal) row, longest transcript.
inline double degradation_rate(size_t G, const double *gene_nobs, const double *gene_partial_nobs,
                               const uint8_t *has_partial, const double *gene_len)
{
    std::vector<double> vals_keep, vals_all;
    for (size_t g = 0; g < G; ++g) {
        const double nb = gene_nobs[g];
        const double db = has_partial[g] ? gene_partial_nobs[g] : NAN;
        const double v = (db / nb) * 1000 / gene_len[g];
        vals_all.push_back(v);
        if (nb >= 30 && (nb - db) >= 5) vals_keep.push_back(v);     // NA comparisons are not TRUE
    }
    return r_median(vals_keep.empty() ? vals_all : vals_keep);
}

// assignGroups (R/bambu-quantify_utilityFunctions.R:347-356) for the genes that kept rows (n_tx > 0):
// tr_dimension = n_tx * n_eq, ascending (stable in gene_sid), gene_grp_id = cumsum %/% 1000 + 1.
// Returns the sorted distinct group ids; group_index[g] = position of gene g's id in that list.
inline std::vector<int64_t> assign_groups(const std::vector<uint32_t> &n_tx, const std::vector<uint32_t> &n_eq,
                                          std::vector<uint32_t> &group_index)
{
    const size_t G = n_tx.size();
    struct DG { uint64_t dim; uint32_t gene; };
    std::vector<DG> a, b;
    a.reserve(G);
    uint64_t dmax = 0;
    for (uint32_t g = 0; g < G; ++g)
        if (n_tx[g]) { a.push_back(DG{(uint64_t)n_tx[g] * n_eq[g], g}); dmax = std::max(dmax, a.back().dim); }
    // order(tr_dimension): stable LSD radix on the dimension (genes enter in ascending gene_sid)
    b.resize(a.size());
    for (int sh = 0; sh < bits_for(dmax); sh += 16) {
        std::vector<uint32_t> cnt(65537, 0);
        for (const DG &e : a) ++cnt[((e.dim >> sh) & 0xffffu) + 1];
        for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
        for (const DG &e : a) b[cnt[(e.dim >> sh) & 0xffffu]++] = e;
        a.swap(b);
    }
    // the cumulative sum only grows, so the group ids come out sorted: number them as they appear
    std::vector<int64_t> gids;
    group_index.assign(G, 0);
    uint64_t cum = 0;
    for (const DG &e : a) {
        cum += e.dim;
        const int64_t id = (int64_t)(cum / 1000 + 1);
        if (gids.empty() || gids.back() != id) gids.push_back(id);
        group_index[e.gene] = (uint32_t)(gids.size() - 1);
    }
    return gids;
}

}  // namespace qp
