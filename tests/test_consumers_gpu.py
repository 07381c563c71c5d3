"""Parity of the consumers of the pairwise matrices against the oracle's
literal restatement of the reference: filter_high_relatedness (index sets must be identical),
the Weir & Goudet statistics (FP64; the reference averages in long double, the GPU in
double-double: tolerance 1e-12, in practice the same double), and the O(n m) scans
(integer counts, bit-exact)."""
import numpy as np
import pytest

import tidypopgen_b200 as tpg
from conftest import EXAMPLE_7x6, same
from oracle import oracle as O
from tidypopgen_b200 import _lib as L

pytestmark = pytest.mark.gpu

WG_ATOL = 1e-12


def fbm_of(raw, n, m):
    return tpg.FBMCode256.from_bytes(raw, n, m)


def related_matrix(rng, n, n_fam, noise=0.02):
    """A KING-like symmetric matrix: families of 2-5 members with kinship 0.1-0.3, background
    noise around 0 (negative values included: the filter takes |.|)."""
    m = rng.normal(0, noise, size=(n, n))
    m = (m + m.T) / 2
    left = list(rng.permutation(n))
    for _ in range(n_fam):
        size = int(rng.integers(2, 6))
        if len(left) < size:
            break
        fam = [left.pop() for _ in range(size)]
        for a in fam:
            for b in fam:
                if a < b:
                    m[a, b] = m[b, a] = rng.uniform(0.1, 0.3)
    np.fill_diagonal(m, 0.5)
    return m


def check_filter(mat, thr):
    ref = O.filter_high_relatedness(mat, thr)
    if ref["error"]:
        with pytest.raises(L.TpgError, match="missing value where TRUE/FALSE needed"):
            tpg.filter_high_relatedness(mat, kings_threshold=thr)
        return None
    passed, removed, keep = tpg.filter_high_relatedness(mat, kings_threshold=thr)
    n = mat.shape[0]
    assert np.array_equal(keep, ref["keep"])
    assert passed == [str(k + 1) for k in ref["order"] if ref["keep"][k]]     # var_names[new_order][!col_to_delete]
    assert removed == [str(k + 1) for k in range(n) if not ref["keep"][k]]
    return keep


def test_filter_families_reference_test(families):
    """test_filter_high_relatedness.R: families.bed at 0.2 — pairs (9,10) and (11,12) are above;
    after filtering, a second pass keeps everybody."""
    X = tpg.FBMCode256(families["path"], 12, 961)
    king = tpg.snp_king(X)
    keep = check_filter(king, 0.2)
    assert keep.sum() == 10 and (keep[8] ^ keep[9]) and (keep[10] ^ keep[11])
    kept = np.flatnonzero(keep)
    assert check_filter(king[np.ix_(kept, kept)], 0.2).all()
    # with ids from a gen_tibble (R/filter_high_relatedness.R:37-40)
    gt = tpg.gen_tibble(X, indiv_id=[f"id{k}" for k in range(12)])
    p, r, k = tpg.filter_high_relatedness(tpg.pairwise_king(gt, as_matrix=True), gt, kings_threshold=0.2)
    assert sorted(r) == sorted(f"id{i}" for i in np.flatnonzero(~keep)) and len(p) == 10 and k.sum() == 10


@pytest.mark.parametrize("n,fam,thr", [(2, 1, 0.05), (3, 1, 0.05), (40, 6, 0.088), (150, 25, 0.088), (300, 60, 0.177),
                                       (257, 40, 0.01), (64, 0, 0.088), (500, 120, 0.088)])
def test_filter_random_parity(n, fam, thr):
    rng = np.random.default_rng(n * 7 + fam)
    check_filter(related_matrix(rng, n, fam), thr)


def test_filter_edge_cases():
    rng = np.random.default_rng(3)
    # one individual (:46-53)
    p, r, k = tpg.filter_high_relatedness(np.array([[0.5]]), kings_threshold=0.1)
    assert p == ["1"] and r == [] and k.tolist() == [True]
    # exact ties (dyadic values): both sides must break them the same way
    t = np.full((6, 6), 0.25)
    np.fill_diagonal(t, 0.5)
    check_filter(t, 0.125)
    # nothing above the threshold: NaNs are never compared, nobody is removed
    m = related_matrix(rng, 30, 0)
    m[3, 7] = m[7, 3] = np.nan
    assert check_filter(m, 0.2).all()
    # a NaN met while pairs above the threshold remain: R stops, so do we
    errors = 0
    for seed in range(12):
        r2 = np.random.default_rng(100 + seed)
        m = related_matrix(r2, 30, 5)
        a, b = r2.choice(30, 2, replace=False)
        m[a, b] = m[b, a] = np.nan
        errors += O.filter_high_relatedness(m, 0.088)["error"]
        check_filter(m, 0.088)               # follows the oracle: same error, or same result
    assert errors >= 1
    # infinities (KING's x/0) behave like very related pairs
    m = related_matrix(rng, 50, 8)
    m[4, 9] = m[9, 4] = np.inf
    m[20, 30] = m[30, 20] = -np.inf
    check_filter(m, 0.088)
    # an all-NaN individual sorts last and is left alone when nothing else is above the threshold
    m = related_matrix(rng, 20, 0)
    m[5, :] = np.nan
    m[:, 5] = np.nan
    assert check_filter(m, 0.4).all()


def test_filter_large_runs_and_is_idempotent():
    """n = 6000 (out of the oracle's reach: O(n^2) per decision): idempotence — the kept set has
    no pair above the threshold, and filtering it again keeps everybody."""
    rng = np.random.default_rng(9)
    m = related_matrix(rng, 6000, 900)
    p, r, keep = tpg.filter_high_relatedness(m, kings_threshold=0.088)
    kept = np.flatnonzero(keep)
    sub = np.abs(m[np.ix_(kept, kept)])
    np.fill_diagonal(sub, 0)
    assert sub.max() <= 0.088 and 0 < len(r) < 6000
    assert tpg.filter_high_relatedness(m[np.ix_(kept, kept)], kings_threshold=0.088)[2].all()


def test_wg17_example(example_gt):
    """7 x 6 example of test_indiv_inbreeding.R / test_pop_fis_fst.R (pop1 = a,b,e; pop2 = c,d;
    pop3 = f,g): inbreeding = diag(beta.dosage) = diag(GRM) - 1."""
    gt = tpg.gen_tibble(EXAMPLE_7x6, indiv_id=list("abcdefg"))
    inb = tpg.indiv_inbreeding(gt)
    a = O.snp_allele_sharing(example_gt["raw"], 7, 6)["as"]
    ref, mb = O.indiv_inbreeding(a)
    assert np.allclose(inb, ref, rtol=0, atol=WG_ATOL) and abs(mb - 0.584126984126984) < 1e-14
    assert abs(inb[0] - (0.396947 - 1)) < 1e-6 and abs(inb[1] - (1.519084 - 1)) < 1e-6
    grm = tpg.pairwise_grm(gt)
    assert np.allclose(inb, np.diag(grm) - 1, rtol=0, atol=1e-12)
    gt.population = ["pop1", "pop1", "pop2", "pop2", "pop1", "pop3", "pop3"]
    grp = np.array([0, 0, 1, 1, 0, 2, 2], np.int32)
    fst = tpg.pop_fst(gt, include_global=True)
    fis = tpg.pop_fis(gt, include_global=True)
    rf, ri = O.pop_fst(a, grp, 3), O.pop_fis_wg17(a, grp, 3)
    assert list(fst) == ["pop1", "pop2", "pop3", "global"] and list(fis) == list(fst)
    assert np.allclose(list(fst.values()), rf, rtol=0, atol=WG_ATOL, equal_nan=True)
    assert np.allclose(list(fis.values()), ri, rtol=0, atol=WG_ATOL, equal_nan=True)
    # grouped indiv_inbreeding: one vector per population from that population's own matrix
    per_pop = tpg.indiv_inbreeding(gt)
    sub = tpg.gen_tibble(EXAMPLE_7x6[[0, 1, 4]], indiv_id=list("abe"))
    assert len(per_pop) == 3 and np.allclose(per_pop[0], tpg.indiv_inbreeding(sub), rtol=0, atol=WG_ATOL)
    with pytest.raises(ValueError, match="grouped"):
        tpg.pop_fst(tpg.gen_tibble(EXAMPLE_7x6))


@pytest.mark.parametrize("n,m,npop", [(60, 400, 3), (257, 900, 7), (500, 300, 40)])
def test_wg17_random_parity(n, m, npop):
    rng = np.random.default_rng(n + npop)
    raw = rng.choice(np.array([0, 1, 2, 3], np.uint8), size=m * n, p=[.4, .3, .2, .1])
    X = fbm_of(raw, n, m)
    a = tpg.snp_allele_sharing(X)
    ref_a = O.snp_allele_sharing(raw, n, m)["as"]
    assert same(a, ref_a)
    grp = rng.integers(0, npop, n).astype(np.int32)
    grp[:npop] = np.arange(npop)                 # every population is present
    gt = tpg.gen_tibble(X, population=[f"p{g:03d}" for g in grp])
    fst, fis = tpg.pop_fst(gt, include_global=True, allele_sharing_mat=a), tpg.pop_fis(gt, include_global=True, allele_sharing_mat=a)
    assert np.allclose(list(fst.values()), O.pop_fst(ref_a, grp, npop), rtol=0, atol=WG_ATOL, equal_nan=True)
    assert np.allclose(list(fis.values()), O.pop_fis_wg17(ref_a, grp, npop), rtol=0, atol=WG_ATOL, equal_nan=True)
    gt.population = None
    assert np.allclose(tpg.indiv_inbreeding(gt, allele_sharing_mat=a), O.indiv_inbreeding(ref_a)[0], rtol=0, atol=WG_ATOL)


# ------------------------------------------------------------------ scans through code256

def test_scans_example(example_gt):
    gt = tpg.gen_tibble(EXAMPLE_7x6)
    g = EXAMPLE_7x6
    counts = tpg.indiv_het_obs(gt, as_counts=True)
    assert counts[:, 0].tolist() == np.nansum(g == 1, axis=1).tolist()
    assert counts[:, 1].tolist() == np.isnan(g).sum(axis=1).tolist()
    # test_indiv_het_obs.R:35-44: het_obs = rowSums(g == 1) / (ncol - n_na)
    assert same(tpg.indiv_het_obs(gt), np.nansum(g == 1, axis=1) / (6 - np.isnan(g).sum(axis=1)))
    # test_indiv_missingness.R: n_na / n_loci
    assert same(tpg.indiv_missingness(gt), np.isnan(g).sum(axis=1) / 6)
    assert same(tpg.loci_missingness(gt), np.isnan(g).sum(axis=0) / 7)
    f, v = O.alt_freq(example_gt["raw"], 7, 6)
    assert same(tpg.loci_alt_freq(gt), f)
    assert np.array_equal(tpg.loci_alt_freq(gt, as_counts=True)[:, 1], v)


@pytest.mark.parametrize("case", [(929, 3000, "all", "all"), (300, 2000, "shuffle", "shuffle"), (1500, 700, "window", "every3"),
                                  (64, 5000, "dup", "all"), (2500, 513, "subset", "subset"),
                                  (1027, 801, "window", "all"), (4102, 333, "all", "all")])
@pytest.mark.parametrize("src_kind", ["file", "device", "pinned"])
@pytest.mark.parametrize("table_kernel", [0, 1, 2, 3], ids=["auto", "lut", "rounds", "warp"])
def test_scans_random_parity(case, src_kind, table_kernel, tpg_cfg):
    """Every case on every scan kernel (tpg_config.scan_table): the two table-free builds (standard code tables,
    contiguous rows and loci — warp per locus from 512 individuals on, 16-locus rounds below; the other cases
    fall back by themselves) and the lookup-table one."""
    import torch
    from test_parity_gpu import pick

    tpg_cfg(scan_table=table_kernel)
    n_tot, m_tot, rk, ck = case
    rng = np.random.default_rng(n_tot + m_tot)
    raw = rng.choice(np.array([0, 1, 2, 3, 4, 5, 6, 200], np.uint8), size=m_tot * n_tot, p=[.35, .25, .2, .08, .04, .03, .03, .02])
    X = fbm_of(raw, n_tot, m_tot)
    rows, cols = pick(rk, n_tot, rng), pick(ck, m_tot, rng)
    n = n_tot if rows is None else len(rows)
    ploidy = rng.choice([1.0, 2.0], size=n, p=[.3, .7])
    kw = {}
    if src_kind == "device":
        dev = torch.from_numpy(raw).cuda()
        kw["src"] = (L.SRC_DEVICE, dev.data_ptr(), dev)
    if src_kind == "pinned":                      # DMA'd raw from the caller's pinned memory
        pin = torch.from_numpy(raw).pin_memory()
        kw["src"] = (L.SRC_HOST, pin.data_ptr(), pin)
    kw["gpu_chunk_loci"] = 512
    for code in (O.CODE_012, O.CODE_IMPUTE_PRED):
        r = tpg.scan_counts(X, rows, cols, code256=code, ploidy=ploidy, **kw)
        het, na = O.gt_ind_hetero(raw, n_tot, m_tot, rows, cols, code256=code)
        assert np.array_equal(r["het_n"], het) and np.array_equal(r["na_n"], na)
        f, v = O.alt_freq(raw, n_tot, m_tot, rows, cols, code256=code, ploidy=ploidy)
        assert same(r["loc0"], f) and np.array_equal(O.is_na(r["loc0"]), O.is_na(f)) and np.array_equal(r["n_valid"], v)
        c, _ = O.alt_freq(raw, n_tot, m_tot, rows, cols, code256=code, ploidy=ploidy, as_counts=True)
        assert np.array_equal(tpg.scan_counts(X, rows, cols, code256=code, ploidy=ploidy, as_counts=True, **kw)["loc0"], c)


def test_scans_reject_fractional_codes_and_polyploids():
    X = fbm_of(np.zeros(12, np.uint8), 3, 4)
    code = O.CODE_012.copy()
    code[7] = 0.5
    with pytest.raises(L.TpgError, match="integer dosages"):
        tpg.scan_counts(X, code256=code)
    with pytest.raises(L.TpgError, match="diploid"):
        tpg.scan_counts(X, ploidy=[2, 4, 2])


def test_qc_report_indiv_families(families):
    """test_qc_report_indiv.R / test_filter_high_relatedness.R: the report's to_keep is the filter's."""
    X = tpg.FBMCode256(families["path"], 12, 961)
    gt = tpg.gen_tibble(X)
    rep = tpg.qc_report_indiv(gt, kings_threshold=0.2)
    het, na = O.gt_ind_hetero(families["raw"], 12, 961)
    assert same(rep["het_obs"], het / (961 - na)) and same(rep["missingness"], na / 961)
    king = O.snp_king(families["raw"], 12, 961, block_size_=O.block_size(12))["king"]
    assert np.array_equal(rep["to_keep"], O.filter_high_relatedness(king, 0.2)["keep"]) and rep["to_keep"].sum() == 10
    assert same(rep["king"], king)
    rel, k2 = tpg.king_filter(gt, 0.2)            # matrix never leaves HBM
    assert k2 is None and np.array_equal(rel[2], rep["to_keep"])
    assert rel[0] == tpg.filter_high_relatedness(king, gt, kings_threshold=0.2)[0]
    assert "to_keep" not in tpg.qc_report_indiv(gt)
    assert tpg.qc_report_indiv(gt, kings_threshold="first")["to_keep"].sum() == 10


@pytest.mark.parametrize("n,npop,contiguous", [(12, 3, True), (700, 9, False), (1500, 40, False), (900, 1, True)])
def test_grouped_king_filter_equals_one_job_per_group(n, npop, contiguous):
    """qc_report_indiv on a grouped gen_tibble (R/qc_report_indiv.R:110-181): one KING + one filter per group.
    tpg_king_filter_groups does it in ONE block-diagonal job; every group's matrix and flags equal the
    oracle's per-group computation, and to_keep is assembled in the reference's order."""
    rng = np.random.default_rng(n + npop)
    m = 1500
    raw = rng.choice(np.array([0, 1, 2, 3], np.uint8), size=m * n, p=[.45, .3, .2, .05])
    g = raw.reshape(m, n)
    for a in range(0, n - 1, 7):                      # plant related pairs (near-duplicates) inside the data
        g[:, a + 1] = np.where(rng.random(m) < 0.9, g[:, a], g[:, a + 1])
    raw = g.reshape(-1)
    pops = np.sort(rng.integers(0, npop, n)) if contiguous else rng.integers(0, npop, n)
    labels = [f"pop{p:03d}" for p in pops]
    gt = tpg.gen_tibble(fbm_of(raw, n, m), population=labels)
    keep, mats, keys = tpg.king_filter_groups(gt, labels, 0.177, return_king=True)
    assert keys == sorted(set(labels))
    bs = O.block_size(n)
    for lab in keys:
        idx = np.flatnonzero(np.array(labels) == lab)
        ref_king = O.snp_king(raw, n, m, (idx + 1).astype(np.int32), None, block_size_=bs)["king"]
        assert same(mats[lab], ref_king), lab
        ref_keep = O.filter_high_relatedness(ref_king, 0.177)["keep"] if len(idx) > 1 else np.array([True])
        assert np.array_equal(keep[idx], ref_keep), lab
    assert (~keep).sum() > 0 or npop >= 40
    rep = tpg.qc_report_indiv(gt, kings_threshold="first", grouped=True)
    first_seen = list(dict.fromkeys(labels))
    assert np.array_equal(rep["to_keep"], np.concatenate([keep[np.array(labels) == lab] for lab in first_seen]))
    if contiguous:
        assert np.array_equal(rep["to_keep"], keep)
