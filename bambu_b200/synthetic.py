"""Synthetic read-class sets of the shapes ``BASELINE.json:configs`` names.

The generators follow SURVEY.md 8(d).  They produce what the reference's
R-side packing (``R/bambu-quantify.R:56-61``) would hand to
``abundance_quantification``: per gene a transcript x read-class pattern with
A-values built by the degradation-bias rule of
``modifyAvaluewithDegradation_rate``
(``R/bambu-quantify_utilityFunctions.R:276-296``), ``K``/``n.obs`` per column
(``:221-228``) and gene groups assigned by the literal ``assignGroups`` rule
(``:347-356``).  Everything is seeded ``numpy.random.default_rng``; there is no
real data here.

    config 2  single sample, 60k genes / 250k tx / 500k read classes / ~2M nnz
    config 3  5k genes incl. 50 giant loci (2-4k tx, 100-200k read classes each)
    config 4  cohort: config-2 annotation shared, per-sample counts / d_rate / 20 % of
              the multi-transcript classes re-drawn
    config 5  like 4 with 75k genes / 400k tx / 800k read classes / ~3.2M nnz
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .arena import Arena

GROUP_CELLS = 1000  # R/bambu-quantify_utilityFunctions.R:352  cumN %/% 1000 + 1


# --------------------------------------------------------------------------
@dataclass
class Annotation:
    """Sample-invariant part: gene sizes, grouping, and the base pattern of the
    multi-transcript read classes."""
    m: np.ndarray            # transcripts per gene
    n: np.ndarray            # read classes per gene (m unique + (n-m) free)
    lam: float               # expected extra transcripts per free class
    grp: np.ndarray          # gene_grp_id per gene (1-based, as in R)
    seed: int
    giant: np.ndarray        # bool per gene: sampled with the giant-locus sampler
    giant_k: float = 19.0

    @property
    def n_genes(self):
        return len(self.m)


def _gene_sizes(G, sum_m, cap, rng):
    mean_x = sum_m / G - 1.0
    p = 1.0 / (1.0 + mean_x)
    m = 1 + np.minimum(rng.geometric(p, G) - 1, cap - 1)
    # rescale to hit sum_m exactly: +-1 on random genes
    diff = int(sum_m - m.sum())
    while diff != 0:
        step = 1 if diff > 0 else -1
        ok = np.nonzero((m + step >= 1) & (m + step <= cap))[0]
        pick = rng.choice(ok, size=min(abs(diff), len(ok)), replace=False)
        m[pick] += step
        diff = int(sum_m - m.sum())
    return m.astype(np.int64)


def assign_groups(tr_dimension: np.ndarray) -> np.ndarray:
    """``assignGroups``: stable ascending order of ``tr_dimension``, running sum,
    ``%/% 1000 + 1`` (R/bambu-quantify_utilityFunctions.R:347-356)."""
    order = np.argsort(tr_dimension, kind="stable")
    cum = np.cumsum(tr_dimension[order])
    grp = np.empty(len(tr_dimension), dtype=np.int64)
    grp[order] = cum // GROUP_CELLS + 1
    return grp


def _solve_lambda(m, n_free, target_extra):
    """lambda with sum_g n_free_g * min(m_g - 1, lambda) = target_extra."""
    lo, hi = 0.0, float(m.max())
    if (n_free * (m - 1)).sum() <= target_extra:
        return hi
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if (n_free * np.minimum(m - 1, mid)).sum() < target_extra:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def make_annotation(n_genes=60_000, sum_m=250_000, sum_n=500_000, nnz=2_000_000, seed=20260002,
                    n_giant=0, cap=200) -> Annotation:
    rng = np.random.default_rng(seed)
    G_small = n_genes - n_giant
    m = _gene_sizes(G_small, sum_m, cap, rng)
    n = 2 * m
    diff = int(sum_n - n.sum())
    if diff != 0:  # spread the remainder over random genes (keeps n >= m)
        pick = rng.choice(G_small, size=abs(diff), replace=True)
        np.add.at(n, pick, 1 if diff > 0 else 0)
    lam = _solve_lambda(m, n - m, nnz - n.sum())
    giant = np.zeros(n_genes, dtype=bool)
    if n_giant:
        gm = rng.integers(2000, 4001, n_giant)
        gn = rng.integers(100_000, 200_001, n_giant)
        pos = np.sort(rng.choice(n_genes, n_giant, replace=False))
        mm = np.empty(n_genes, np.int64)
        nn = np.empty(n_genes, np.int64)
        giant[pos] = True
        mm[giant], nn[giant] = gm, gn
        mm[~giant], nn[~giant] = m, n
        m, n = mm, nn
    grp = assign_groups(m * n)
    return Annotation(m=m, n=n, lam=lam, grp=grp, seed=seed, giant=giant)


# --------------------------------------------------------------------------
def _free_class_entries(m, n_free, lam, rng):
    """Pattern of the free (non-unique) classes of ordinary genes: class picks
    one anchor transcript plus every other transcript of its gene with
    probability min(1, lam/(m-1))  =>  k ~ 1 + Binomial(m-1, p_g).
    Returns (gene, class_local(0..n_free-1), tx_local) of every entry."""
    G = len(m)
    ncls = int(n_free.sum())
    cls_gene = np.repeat(np.arange(G), n_free)
    cls_local = np.arange(ncls) - np.repeat(np.cumsum(n_free) - n_free, n_free)
    cm = m[cls_gene]
    # all (class, tx) pairs of ordinary genes
    pair_cls = np.repeat(np.arange(ncls), cm)
    pair_tx = np.arange(int(cm.sum())) - np.repeat(np.cumsum(cm) - cm, cm)
    anchor = (rng.random(ncls) * cm).astype(np.int64)
    p = np.minimum(1.0, lam / np.maximum(cm - 1, 1))
    keep = (rng.random(len(pair_cls)) < p[pair_cls]) | (pair_tx == anchor[pair_cls])
    pair_cls, pair_tx = pair_cls[keep], pair_tx[keep]
    return cls_gene[pair_cls], cls_local[pair_cls], pair_tx


def _giant_class_entries(m_g, n_free, k_mean, rng):
    """Giant locus: k ~ 1 + Binomial(m-1, k_mean/m) transcripts per class,
    drawn without enumerating all m x n pairs."""
    k = 1 + rng.binomial(m_g - 1, k_mean / m_g, n_free)
    cls = np.repeat(np.arange(n_free), k)
    tx = rng.integers(0, m_g, len(cls))
    key = np.unique(cls.astype(np.int64) * m_g + tx)   # duplicates collapse (k shrinks slightly)
    return key // m_g, key % m_g


def sample_arena(ann: Annotation, sample: int = 0, d_rate: float | None = 0.1, resample_frac: float = 0.0,
                 observed_frac: float = 0.4, seed: int | None = None, long_table: bool = False):
    """One sample's arena.  ``sample`` seeds the counts; the pattern is the
    annotation's base pattern (seed ``ann.seed``) with ``resample_frac`` of the
    free classes re-drawn for this sample; ``d_rate=None`` draws U(0.02, 0.3)."""
    m, n, G = ann.m, ann.n, ann.n_genes
    base = np.random.default_rng(ann.seed + 7919)
    rng = np.random.default_rng((ann.seed if seed is None else seed) + 1_000_003 * (sample + 1))
    if d_rate is None:
        d_rate = float(rng.uniform(0.02, 0.3))
    small = ~ann.giant
    n_free = n - m

    # ---- pattern: unique class per tx + free classes -----------------------
    gs = np.nonzero(small)[0]
    eg, ec, et = _free_class_entries(m[gs], n_free[gs], ann.lam, base)
    if resample_frac > 0:
        ncls = int(n_free[gs].sum())
        redo = rng.random(ncls) < resample_frac
        cls_off = np.cumsum(n_free[gs]) - n_free[gs]
        cid = cls_off[eg] + ec
        keepm = ~redo[cid]
        eg2, ec2, et2 = _free_class_entries(m[gs], n_free[gs], ann.lam, rng)
        cid2 = cls_off[eg2] + ec2
        takem = redo[cid2]
        eg = np.concatenate((eg[keepm], eg2[takem]))
        ec = np.concatenate((ec[keepm], ec2[takem]))
        et = np.concatenate((et[keepm], et2[takem]))
    eg = gs[eg]
    parts_g, parts_c, parts_t = [eg], [ec], [et]
    for g in np.nonzero(ann.giant)[0]:
        c, t = _giant_class_entries(int(m[g]), int(n_free[g]), ann.giant_k, base)
        parts_g.append(np.full(len(c), g, np.int64))
        parts_c.append(c)
        parts_t.append(t)
    # unique classes: tx t of gene g <-> free-class index n_free + t
    ug = np.repeat(np.arange(G), m)
    ut = np.arange(int(m.sum())) - np.repeat(np.cumsum(m) - m, m)
    parts_g.append(ug)
    parts_c.append(n_free[ug] + ut)
    parts_t.append(ut)
    eg = np.concatenate(parts_g)
    ec = np.concatenate(parts_c)
    et = np.concatenate(parts_t)

    # classes get a random position inside their gene (eqClassId order is arbitrary)
    col_off = np.cumsum(n) - n                     # first class of gene (gene order)
    ncols = int(n.sum())
    pos_key = base.random(ncols)
    gene_of_col = np.repeat(np.arange(G), n)
    order = np.lexsort((pos_key, gene_of_col))     # permutation inside each gene
    newpos = np.empty(ncols, np.int64)
    newpos[order] = np.arange(ncols)
    ecol = newpos[col_off[eg] + ec]                # class id, global in gene order
    row_off = np.cumsum(m) - m
    erow = row_off[eg] + et                        # tx id, global in gene order

    # ---- per class: multiplicity, width; per entry: equal flag -------------
    k_col = np.bincount(ecol, minlength=ncols)
    multi = k_col > 1
    w_col = base.uniform(200.0, 3000.0, ncols)
    has_eq = base.random(ncols) < 0.5
    # the entry flagged equal: the one with the smallest random key in its class
    ekey = base.random(len(ecol))
    o = np.lexsort((ekey, ecol))
    first = np.ones(len(o), bool)
    first[1:] = ecol[o][1:] != ecol[o][:-1]
    equal = np.zeros(len(ecol), bool)
    equal[o[first]] = True
    equal &= has_eq[ecol]

    # ---- A-values (R/bambu-quantify_utilityFunctions.R:276-296) ------------
    part = w_col[ecol] * d_rate / 1000.0
    is_partial_multi = multi[ecol] & ~equal
    psum = np.bincount(erow, weights=np.where(is_partial_multi, part, 0.0), minlength=int(m.sum()))
    aval = np.where(equal, 1.0 - psum[erow], part)
    aval = np.clip(aval, 0.0, 1.0)
    aval = np.where(multi[ecol] & equal, np.minimum(1.0, np.maximum(aval, part)), aval)
    aval = np.where(multi[ecol], aval, 1.0)

    # ---- counts -------------------------------------------------------------
    r, mean = 0.5, 40.0
    nobs = np.where(rng.random(ncols) < observed_frac,
                    1.0 + rng.negative_binomial(r, r / (r + mean), ncols), 0.0)
    tx_obs = np.bincount(erow, weights=nobs[ecol], minlength=int(m.sum()))
    dead_tx = np.nonzero(tx_obs == 0)[0]
    if len(dead_tx):   # give the transcript's unique class one observation draw
        g_of = np.repeat(np.arange(G), m)[dead_tx]
        ucol = newpos[col_off[g_of] + n_free[g_of] + (dead_tx - row_off[g_of])]
        nobs[ucol] = 1.0 + rng.negative_binomial(r, r / (r + mean), len(ucol))
    # nobs is indexed by final class position; gene of final position p is gene_of_col[p]
    Kg = np.bincount(gene_of_col, weights=nobs, minlength=G)
    K = Kg[gene_of_col]
    Y = nobs / K

    if long_table:
        # the same sample as the long table bambu.quantDT starts from (R/bambu-quantify.R:53-56):
        # one row per (class, transcript), rows in random order.  txlen is drawn so that the
        # a-values differ from the arena's (the degradation rate is estimated from the table).
        trng = np.random.default_rng(ann.seed + 104_729)
        txlen = trng.integers(400, 6000, int(m.sum())).astype(np.float64)
        perm = rng.permutation(len(ecol))
        return {"GENEID": eg[perm].astype(np.int64), "txid": (erow[perm] + 1).astype(np.int64),
                "eqClassId": (ecol[perm] + 1).astype(np.int64), "nobs": nobs[ecol[perm]],
                "equal": equal[perm], "rcWidth": np.minimum(w_col[ecol[perm]], txlen[erow[perm]]),
                "txlen": txlen[erow[perm]]}

    # ---- arrange genes group by group ---------------------------------------
    gorder = np.lexsort((np.arange(G), ann.grp))   # by group id, then gene_sid
    grank = np.empty(G, np.int64)
    grank[gorder] = np.arange(G)
    m_o, n_o = m[gorder], n[gorder]
    row_new_off = np.cumsum(m_o) - m_o
    col_new_off = np.cumsum(n_o) - n_o
    row_map = np.repeat(row_new_off[grank] - row_off, m) + np.arange(int(m.sum()))
    col_map = np.repeat(col_new_off[grank] - col_off, n) + np.arange(ncols)
    grp_sorted = ann.grp[gorder]
    gb = np.nonzero(np.concatenate(([True], grp_sorted[1:] != grp_sorted[:-1], [True])))[0]
    grp_row_ptr = np.concatenate((row_new_off, [m.sum()]))[gb].astype(np.int64)
    grp_col_ptr = np.concatenate((col_new_off, [ncols]))[gb].astype(np.int64)

    r_new = row_map[erow]
    c_new = col_map[ecol]
    o = np.lexsort((c_new, r_new))
    r_new, c_new = r_new[o], c_new[o]
    grp_of_row = np.repeat(np.arange(len(gb) - 1), np.diff(grp_row_ptr))
    col_idx = (c_new - grp_col_ptr[grp_of_row[r_new]]).astype(np.int32)
    row_nnz_ptr = np.concatenate(([0], np.cumsum(np.bincount(r_new, minlength=int(m.sum()))))).astype(np.int64)
    inv_col = np.empty(ncols, np.int64)
    inv_col[col_map] = np.arange(ncols)
    inv_row = np.empty(int(m.sum()), np.int64)
    inv_row[row_map] = np.arange(int(m.sum()))
    a = Arena(grp_row_ptr, grp_col_ptr, row_nnz_ptr, col_idx, aval[o].astype(np.float64),
              equal[o].astype(np.uint8), Y[inv_col], K[inv_col], multi[inv_col].astype(np.uint8),
              (inv_row + 1).astype(np.int32))
    a.meta = {"d_rate": d_rate, "sample": sample}
    return a


# --------------------------------------------------------------------------
def config2(seed=20260002, scale=1.0) -> Arena:
    """Synthetic single sample (BASELINE.json configs[1]).  ``scale`` shrinks
    every size proportionally (tests)."""
    ann = annotation_for("config2", seed, scale)
    return sample_arena(ann, 0, d_rate=0.1)


def annotation_for(name: str, seed: int | None = None, scale: float = 1.0) -> Annotation:
    s = lambda x: max(1, int(round(x * scale)))  # noqa: E731
    if name in ("config2", "config4"):
        return make_annotation(s(60_000), s(250_000), s(500_000), s(2_000_000),
                               seed if seed is not None else (20260002 if name == "config2" else 20260004))
    if name == "config3":
        ng = max(1, int(round(50 * scale)))
        G = s(5_000)
        Gs = G - ng
        return make_annotation(G, int(round(Gs * 250 / 60)), int(round(Gs * 500 / 60)), int(round(Gs * 2000 / 60)),
                               seed if seed is not None else 20260003, n_giant=ng)
    if name == "config5":
        return make_annotation(s(75_000), s(400_000), s(800_000), s(3_200_000),
                               seed if seed is not None else 20260005)
    raise ValueError(name)


def cohort_sample(ann: Annotation, s: int) -> Arena:
    """Sample ``s`` of a cohort (configs 4/5): fresh counts, 20 % of the free
    classes re-drawn, per-sample d_rate ~ U(0.02, 0.3)."""
    return sample_arena(ann, s, d_rate=None, resample_frac=0.2)


# -------------------------------------------------------------------------- config 3, locus by locus
# ``sample_arena`` on the whole config-3 annotation sorts 1.5e8 entries at once (minutes, 20 GB).  A giant locus
# is its own gene group (tr_dimension >= 2e8 >> 1000), so the same workload can be drawn one locus at a time --
# every locus from its own seed -- which is also what lets a rank generate ONLY the loci of its shard.
@dataclass
class Config3Plan:
    small: Annotation          # the ordinary genes (their groups come first, as assignGroups orders them)
    giant_m: np.ndarray        # transcripts / read classes of every giant locus, in group order
    giant_n: np.ndarray
    seed: int

    @property
    def n_giant(self):
        return len(self.giant_m)

    def giant_expected_entries(self):
        """a-priori entry count of a locus: one unique class per transcript + k ~ 1 + Binomial(m-1, 19/m) per free class"""
        return self.giant_m + (self.giant_n - self.giant_m) * (1.0 + 19.0 * (self.giant_m - 1) / self.giant_m)


def config3_plan(scale: float = 1.0, seed: int = 20260003) -> Config3Plan:
    ng = max(1, int(round(50 * scale)))
    Gs = max(1, int(round(5_000 * scale))) - ng
    small = make_annotation(Gs, int(round(Gs * 250 / 60)), int(round(Gs * 500 / 60)), int(round(Gs * 2000 / 60)), seed)
    rng = np.random.default_rng(seed + 31)
    gm = rng.integers(2000, 4001, ng)
    gn = rng.integers(100_000, 200_001, ng)
    o = np.argsort(gm * gn, kind="stable")          # assignGroups: ascending tr_dimension
    return Config3Plan(small, gm[o].astype(np.int64), gn[o].astype(np.int64), seed)


def config3_locus(plan: Config3Plan, k: int, sample: int = 0) -> Arena:
    """Giant locus ``k`` of the plan as a one-group arena (drawn by ``sample_arena`` from the locus' own seed)."""
    ann = Annotation(m=plan.giant_m[k:k + 1].copy(), n=plan.giant_n[k:k + 1].copy(), lam=0.0, grp=np.ones(1, np.int64),
                     seed=plan.seed + 1009 * (k + 1), giant=np.ones(1, bool))
    return sample_arena(ann, sample)


def _config3_locus_job(args):
    scale, seed, k, sample = args
    a = config3_locus(config3_plan(scale, seed), k, sample)
    return (a.grp_row_ptr, a.grp_col_ptr, a.row_nnz_ptr, a.col_idx, a.aval, a.nz_flags, a.Y, a.K, a.col_flags, a.txid)


def config3_arena(scale: float = 1.0, seed: int = 20260003, loci=None, small: bool = True, workers: int = 1,
                  sample: int = 0, mp_context: str = "fork") -> Arena:
    """Config 3 (BASELINE.json configs[2]): the ordinary genes' groups followed by the giant loci ``loci`` (default:
    all), each its own group.  ``workers`` > 1 draws the loci in a process pool (``mp_context``: use "forkserver"
    from a process that already holds a CUDA context)."""
    plan = config3_plan(scale, seed)
    loci = list(range(plan.n_giant)) if loci is None else [int(k) for k in loci]
    jobs = [(scale, seed, k, sample) for k in loci]
    if workers > 1 and len(jobs) > 1:
        import multiprocessing as mp
        with mp.get_context(mp_context).Pool(min(workers, len(jobs))) as pool:
            parts = [Arena(*p) for p in pool.map(_config3_locus_job, jobs, chunksize=1)]
    else:
        parts = [Arena(*_config3_locus_job(j)) for j in jobs]
    head = [sample_arena(plan.small, sample)] if small else []
    return Arena.concatenate(head + parts)
