"""genEquiRCs (bambu_b200/equirc.py) against the independent loop-level restatement
(oracle/equirc_ref.py) on RANDOM annotations / read classes / distance tables -- beyond the bundled chr9 run that
pins both on the reference's stored outputs (tests/test_chr9_e2e.py).  The generator produces the situations the
R code treats specially: read classes matching several transcripts, full-length (`equal`) and partial matches mixed,
read classes whose widths and counts coincide (distinct()-before-sum, R/bambu-quantify_utilityFunctions.R:118-132),
incompatible rows (`_unidentified`), rows in another gene than the read class's own, transcripts without reads."""
import numpy as np
import pytest


def random_fixture(rng, n_genes=12):
    ann = {"TXNAME": [], "GENEID": [], "txid": [], "eqClassById_values": [], "eqClassById_end": [], "exon_width": [],
           "exon_end": []}
    gene_tx = {}
    txid = 0
    for g in range(n_genes):
        gene = f"G{g:03d}"
        m = int(rng.integers(1, 6))
        ids = list(range(txid + 1, txid + m + 1))
        txid += m
        gene_tx[gene] = ids
        for t in ids:
            ann["TXNAME"].append(f"T{t}")
            ann["GENEID"].append(gene)
            ann["txid"].append(t)
            # minimal equivalent class of the transcript: itself plus some transcripts of the same gene, sorted
            others = [u for u in ids if u != t and rng.random() < 0.4]
            cls = sorted([t] + others)
            ann["eqClassById_values"].extend(cls)
            ann["eqClassById_end"].append(len(ann["eqClassById_values"]))
            ne = int(rng.integers(1, 5))
            ann["exon_width"].extend(int(x) for x in rng.integers(50, 900, ne))
            ann["exon_end"].append(len(ann["exon_width"]))
    ann["names"] = list(ann["TXNAME"])
    rc = {"names": [], "exon_width": [], "exon_end": [], "exon_rank": [], "GENEID": []}
    dt = {"annotationTxId": [], "txid": [], "readClassId": [], "readCount": [], "compatible": [], "equal": [], "GENEID": []}
    genes = list(gene_tx)
    n_rc = int(rng.integers(n_genes, 4 * n_genes))
    width_pool = [(int(a), int(b)) for a, b in zip(rng.integers(50, 400, 6), rng.integers(400, 2000, 6))]
    for k in range(n_rc):
        name = f"rc.{k + 1}"
        gene = genes[int(rng.integers(0, len(genes)))]
        rc["names"].append(name)
        rc["GENEID"].append(gene)
        # widths from a small pool so that different read classes coincide in (first exon, total, count)
        first, total = width_pool[int(rng.integers(0, len(width_pool)))]
        ne = int(rng.integers(1, 4))
        ws = [first] + [int((total - first) // max(ne - 1, 1))] * (ne - 1)
        ranks = list(range(1, ne + 1))
        perm = rng.permutation(ne)
        rc["exon_width"].extend(int(ws[p]) for p in perm)
        rc["exon_rank"].extend(float(ranks[p]) for p in perm)
        rc["exon_end"].append(len(rc["exon_width"]))
        if rng.random() < 0.1:
            continue                                  # a read class without any row in the distance table
        count = int(rng.integers(1, 4))
        txs = gene_tx[gene]
        hit = [t for t in txs if rng.random() < 0.6] or [txs[0]]
        eq_done = False
        for t in hit:
            equal = (not eq_done) and rng.random() < 0.3
            eq_done |= equal
            dt["annotationTxId"].append(f"T{t}"); dt["txid"].append(t); dt["readClassId"].append(name)
            dt["readCount"].append(count); dt["compatible"].append(True); dt["equal"].append(bool(equal)); dt["GENEID"].append(gene)
        if rng.random() < 0.2:                        # an incompatible row: dropped
            dt["annotationTxId"].append(f"{gene}_unidentified"); dt["txid"].append(hit[0]); dt["readClassId"].append(name)
            dt["readCount"].append(count); dt["compatible"].append(False); dt["equal"].append(False); dt["GENEID"].append(gene)
        if rng.random() < 0.15 and len(genes) > 1:    # a row in another gene than the read class's own: dropped by the join
            og = genes[(genes.index(gene) + 1) % len(genes)]
            t = gene_tx[og][0]
            dt["annotationTxId"].append(f"T{t}"); dt["txid"].append(t); dt["readClassId"].append(name)
            dt["readCount"].append(count); dt["compatible"].append(True); dt["equal"].append(False); dt["GENEID"].append(og)
    order = rng.permutation(len(dt["txid"]))          # the distance table comes in no particular order
    dt = {k: [v[i] for i in order] for k, v in dt.items()}
    return {"annotation": ann, "readClasses": rc, "distTable": dt}


def canon(tab):
    """rows as sorted tuples, class ids replaced by their row sets (the numbering is checked separately)"""
    n = len(tab["txid"])
    return sorted((str(tab["GENEID"][k]), int(tab["txid"][k]), int(tab["eqClassId"][k]), float(tab["nobs"][k]), bool(tab["equal"][k]),
                   float(tab["rcWidth"][k]), float(tab["txlen"][k]), int(tab["minRC"][k])) for k in range(n))


@pytest.mark.parametrize("seed", range(12))
def test_product_genEquiRCs_equals_the_restatement(seed):
    from bambu_b200.equirc import AnnotationClasses, genEquiRCs
    from oracle.equirc_ref import gen_equi_rcs
    rng = np.random.default_rng(7000 + seed)
    fx = random_fixture(rng, n_genes=int(rng.integers(3, 40)))
    want = gen_equi_rcs(fx)
    got = genEquiRCs(fx["distTable"], fx["readClasses"], fx["annotation"])
    assert len(got["txid"]) == len(want["txid"])
    for col in ("GENEID", "txid", "eqClassId", "nobs", "equal", "rcWidth", "txlen"):
        assert list(got[col]) == list(want[col]), col          # same rows in the same order, same class numbering
    # the native host code (csrc/equirc.cpp, bambu_equirc_create): same rows, same order, same numbering
    from bambu_b200.equirc import genEquiRCs_native
    nat = genEquiRCs_native(fx["distTable"], fx["readClasses"], fx["annotation"])
    for col in ("GENEID", "txid", "eqClassId", "nobs", "equal", "rcWidth", "txlen", "minRC"):
        assert list(nat[col]) == list(got[col]), col
    # the annotation side cached for a cohort gives the same table
    cached = AnnotationClasses(fx["annotation"])
    again = genEquiRCs(fx["distTable"], fx["readClasses"], cached)
    assert canon(again) == canon(got)


def test_random_fixture_through_the_whole_host_chain(oracle_mod):
    """distance table -> genEquiRCs -> bambu.quantDT (oracle EM) -> assays: mass is conserved (every compatible read
    of an observed class is explained by some transcript)."""
    from bambu_b200.equirc import genEquiRCs
    from bambu_b200.quantify import bambu_quantDT
    rng = np.random.default_rng(99)
    fx = random_fixture(rng, n_genes=25)
    tab = genEquiRCs(fx["distTable"], fx["readClasses"], fx["annotation"])
    q = bambu_quantDT(tab, {"degradationBias": True},
                      em=lambda a, maxiter, conv, minvalue: oracle_mod.em_batched(a, maxiter=maxiter, conv=conv, minvalue=minvalue))
    cls_nobs = {}
    for k in range(len(tab["txid"])):
        cls_nobs[(tab["GENEID"][k], int(tab["eqClassId"][k]))] = float(tab["nobs"][k])
    assert abs(q["counts"].sum() - sum(cls_nobs.values())) < 1e-6 * max(1.0, sum(cls_nobs.values()))
