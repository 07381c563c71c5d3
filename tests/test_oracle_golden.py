"""Pins the CPU oracle against every golden vector the reference holds for this path
(SURVEY.md §8c).  No GPU."""
import numpy as np

from conftest import EXAMPLE_7x6, TOY_3x6, TOY_3x6_NA, same, to_bk
from oracle import oracle as O


def king_r(X):
    """The reference's own R restatement, tests/testthat/test_snp_king.R:29-46."""
    X0 = np.nan_to_num((X == 0).astype(float))
    X1 = np.nan_to_num((X == 1).astype(float))
    X2 = np.nan_to_num((X == 2).astype(float))
    num = X1 @ X1.T - 2 * (X0 @ X2.T + X2 @ X0.T)
    valid = (~np.isnan(X)).astype(float)
    Ni = X1 @ valid.T
    Nj = Ni.T
    with np.errstate(all="ignore"):
        return num / (2 * np.minimum(Ni, Nj)) + 0.5 - 0.25 * (Ni + Nj) / np.minimum(Ni, Nj)


def test_split_len_matches_reference_block_pattern():
    # test_snp_king.R:80: block.size = 300 on 961 loci -> 240/240/241/240 (exercises the pad column)
    assert O.split_len(961, 300)[:, 2].tolist() == [240, 240, 241, 240]
    assert O.split_len(6, 3).tolist() == [[1, 3, 3], [4, 6, 3]]
    assert O.split_len(6, 2)[:, 2].tolist() == [2, 2, 2]
    assert O.split_len(5, 100).tolist() == [[1, 5, 5]]
    # block_size(n) = floor(2^30 / (8 n)); SURVEY §8 table
    assert [O.block_size(n) for n in (929, 2504, 50000, 10000)] == [144475, 53601, 2684, 13421]
    # padding-quirk constants of the benchmark shapes
    assert O.n_short(650000, O.block_size(929)) == 0
    assert O.n_short(1000000, O.block_size(2504)) == 8
    assert O.n_short(500000, O.block_size(50000)) == 38
    assert O.n_short(1000000, O.block_size(10000)) == 50


def test_king_toy_identical_to_king_r():
    # test_snp_king.R:50-58
    bk = to_bk(TOY_3x6)
    r = O.snp_king(bk, 3, 6)
    assert same(r["king"], king_r(TOY_3x6))
    assert same(O.snp_king(bk, 3, 6, block_size_=3)["king"], r["king"])


def test_king_families_identical_to_king_r_and_block_invariant(families):
    # test_snp_king.R:66-81
    f = families
    X = f["raw"].reshape(f["m"], f["n"]).T.astype(float)
    X[X == 3] = np.nan
    r = O.snp_king(f["raw"], f["n"], f["m"])
    assert same(r["king"], king_r(X))
    assert same(O.snp_king(f["raw"], f["n"], f["m"], block_size_=300)["king"], r["king"])
    # known entries
    assert r["K"][0, 1] == -32 and r["N_Aa_i"][0, 1] == 326 and r["N_Aa_i"][1, 0] == 309
    assert abs(r["king"][0, 1] - (-0.0655)) < 5e-5


def test_king_families_vs_king_binary(families):
    # test_snp_king.R:84-131: KING's own output, tolerance 0.06 there; we hold its print precision
    f = families
    king = O.snp_king(f["raw"], f["n"], f["m"])["king"]
    for i, j, v in f["kin0"]:
        assert abs(king[i - 1, j - 1] - v) < 6e-5
        assert abs(king[j - 1, i - 1] - v) < 6e-5
    assert np.allclose(np.diag(king), 0.5)


def test_ibs_toy():
    # test_snp_ibs.R:28-36
    bk = to_bk(TOY_3x6)
    r = O.snp_ibs(bk, 3, 6, type="raw_counts")
    assert r["ibs"][0, 1] == sum([1, 2, 2, 1, 1, 2]) == 9
    assert same(O.snp_ibs(bk, 3, 6, type="raw_counts", block_size_=3)["ibs"], r["ibs"])
    # test_snp_ibs.R:58-65: rows {1,3}, loci {2,3,5,6}
    sub = O.snp_ibs(bk, 3, 6, ind_row=[1, 3], ind_col=[2, 3, 5, 6], type="raw_counts")
    assert sub["ibs"][0, 1] == sum(np.array([1, 1, 2, 1, 2, 1])[[1, 2, 4, 5]]) == 6


def test_ibs_families_vs_plink(families):
    # test_snp_ibs.R:69-104: round(, 6) then equal
    f = families
    r = O.snp_ibs(f["raw"], f["n"], f["m"], type="proportion")
    assert np.array_equal(np.round(r["out"], 6), f["mibs"])
    assert r["ibs"][0, 1] == 1253 and r["valid_n"][0, 1] == 1790
    adj = O.snp_ibs(f["raw"], f["n"], f["m"], type="adjusted_counts")["out"]
    assert same(adj, r["out"] * f["m"])


def matching(X):
    """hierfstat::matching by its definition: mean over jointly typed loci of 1 (same
    homozygote) / 0 (opposite homozygotes) / 0.5 (a heterozygote involved)."""
    n = X.shape[0]
    out = np.full((n, n), np.nan)
    for i in range(n):
        for j in range(n):
            ok = ~np.isnan(X[i]) & ~np.isnan(X[j])
            if ok.any():
                out[i, j] = np.mean(0.5 * (1 + (X[i, ok] - 1) * (X[j, ok] - 1)))
    return out


def test_allele_sharing_toy_vs_matching():
    # test_pairwise_allele_sharing.R:29-46
    bk = to_bk(TOY_3x6_NA)
    r = O.snp_allele_sharing(bk, 3, 6)
    assert np.allclose(r["as"], matching(TOY_3x6_NA), rtol=0, atol=1e-15)
    assert same(O.snp_allele_sharing(bk, 3, 6, block_size_=3)["as"], r["as"])  # two equal blocks


def test_allele_sharing_pad_quirk():
    # src/snp_as.cpp:56-64: every block shorter than the longest adds +1 to every numerator
    bk = to_bk(TOY_3x6_NA)
    base = O.snp_allele_sharing(bk, 3, 6)
    q = O.snp_allele_sharing(bk, 3, 6, block_size_=4)  # blocks 3 + 3 -> equal: no quirk
    assert same(q["num"], base["num"])
    bk7 = to_bk(EXAMPLE_7x6[:, :5])
    b1 = O.snp_allele_sharing(bk7, 7, 5)
    b2 = O.snp_allele_sharing(bk7, 7, 5, block_size_=3)  # blocks 2 + 3 -> one short block
    assert O.split_len(5, 3)[:, 2].tolist() == [2, 3]
    assert same(b2["num"], b1["num"] + 1) and same(b2["denom"], b1["denom"])


def test_example_gt_known_values(example_gt):
    # values computed from the formulas for load_example_gt()
    e = example_gt
    assert np.array_equal(e["raw"], to_bk(EXAMPLE_7x6))
    k = O.snp_king(e["raw"], 7, 6)
    assert k["K"][0].tolist() == [4, 1, 1, 2, 2, 0, 2]
    assert k["N_Aa_i"][0].tolist() == [4, 3, 3, 4, 4, 3, 4]
    assert k["king"][0, 1] == 0 and k["king"][0, 2] == 0.125 and k["king"][1, 4] == -1.25
    assert np.isnan(k["king"][1, 2]) and np.isnan(k["king"][5, 6])  # test_autoplot_pairwise.R:17-24
    i = O.snp_ibs(e["raw"], 7, 6, type="raw_counts")
    assert i["ibs"][0].tolist() == [12, 8, 7, 10, 9, 6, 7]
    assert i["valid_n"][0].tolist() == [12, 10, 10, 12, 12, 10, 10]
    a = O.snp_allele_sharing(e["raw"], 7, 6)["as"]
    assert np.allclose(a[0], [2 / 3, 0.7, 0.6, 2 / 3, 7 / 12, 0.6, 0.5], rtol=0, atol=1e-15)
    assert np.allclose(a, matching(EXAMPLE_7x6), rtol=0, atol=1e-15)
    grm, mb = O.grm_from_as(a)
    assert abs(mb - 0.584126984126984) < 1e-15
    assert abs(grm[0, 0] - 0.396947) < 1e-6 and abs(grm[1, 1] - 1.519084) < 1e-6


def test_grm_is_twice_beta_dosage(example_gt):
    # test_pairwise_grm.R:33-40: 2 * hierfstat::beta.dosage(inb = FALSE) == (M - mb)/(1 - mb) * 2
    a = matching(EXAMPLE_7x6)
    off = a[~np.eye(7, dtype=bool)]
    mb = off.mean()
    grm, mb2 = O.grm_from_as(O.snp_allele_sharing(example_gt["raw"], 7, 6)["as"])
    assert abs(mb - mb2) < 1e-15
    assert np.allclose(grm, 2 * (a - mb) / (1 - mb), rtol=0, atol=1e-14)


def test_raw_byte_rule_imputed_and_ploidy():
    # bytes > 2 are missing for KING/IBS, bytes >= 3 for AS — imputed 4/5/6 included
    rng = np.random.default_rng(0)
    g = rng.integers(0, 3, size=(40, 9)).astype(np.uint8)  # [m, n]
    miss = rng.random(g.shape) < 0.2
    a = np.where(miss, 3, g).astype(np.uint8)
    b = np.where(miss, 4 + g, g).astype(np.uint8)  # "imputed" bytes
    for fn, key in ((O.snp_king, "king"), (O.snp_allele_sharing, "as")):
        assert same(fn(a, 9, 40)[key], fn(b, 9, 40)[key])
    assert same(O.snp_ibs(a, 9, 40)["out"], O.snp_ibs(b, 9, 40)["out"])


def test_na_payload():
    bk = to_bk(np.array([[0.0, np.nan], [np.nan, 1.0]]))
    a = O.snp_allele_sharing(bk, 2, 2)["as"]
    assert O.is_na(a)[0, 1] and O.is_na(a)[1, 0] and not O.is_na(a)[0, 0]


# ------------------------------------------------------------------ consumers

def test_filter_high_relatedness_reference_property(families):
    """test_filter_high_relatedness.R: on families.bed at 0.2 only (9,10) and (11,12) are above the
    threshold (KING binary: 0.2741 and 0.2248); one of each pair goes, and a second pass over the
    kept individuals keeps everybody."""
    king = O.snp_king(families["raw"], 12, 961)["king"]
    r = O.filter_high_relatedness(king, 0.2)
    keep = r["keep"]
    assert not r["error"] and keep.sum() == 10 and (keep[8] ^ keep[9]) and (keep[10] ^ keep[11])
    kept = np.flatnonzero(keep)
    assert O.filter_high_relatedness(king[np.ix_(kept, kept)], 0.2)["keep"].all()
    # the visiting order is by decreasing mean |kinship|, ties stable
    a = np.abs(king)
    np.fill_diagonal(a, np.nan)
    cm = np.nanmean(a, axis=0)
    assert np.all(np.diff(cm[r["order"]]) <= 1e-15)


def test_wg17_example_values(example_gt):
    """indiv_inbreeding = diag(hierfstat::beta.dosage) = diag(GRM) - 1 (test_indiv_inbreeding.R:42-44,
    test_pairwise_grm.R); SURVEY G2: GRM[1,1] = 0.396947, GRM[2,2] = 1.519084, mb = 0.584126984126984."""
    a = O.snp_allele_sharing(example_gt["raw"], 7, 6)["as"]
    inb, mb = O.indiv_inbreeding(a)
    g, _ = O.grm_from_as(a)
    assert abs(mb - 0.584126984126984) < 1e-14
    assert np.allclose(inb, np.diag(g) - 1, rtol=0, atol=1e-15)
    assert abs(inb[0] + 0.603053) < 1e-6 and abs(inb[1] - 0.519084) < 1e-6
    grp = np.array([0, 0, 1, 1, 0, 2, 2], np.int32)
    fst, fis = O.pop_fst(a, grp, 3), O.pop_fis_wg17(a, grp, 3)
    # definitional cross-check with plain numpy means
    M = a.copy()
    np.fill_diagonal(M, np.nan)
    Fs = np.array([np.nanmean(M[np.ix_(grp == p, grp == p)]) for p in range(3)])
    Mb = np.mean([np.nanmean(M[np.ix_(grp == i, grp == j)]) for i in range(3) for j in range(i)])
    assert np.allclose(fst[:3], (Fs - Mb) / (1 - Mb), rtol=0, atol=1e-14) and abs(fst[3] - np.mean(fst[:3])) < 1e-14
    Mii = np.diag(a) * 2 - 1
    assert np.allclose(fis[:3], [np.mean((Mii[grp == p] - Fs[p]) / (1 - Fs[p])) for p in range(3)], rtol=0, atol=1e-14)


def test_scans_example(example_gt):
    """indiv_het_obs / indiv_missingness / loci_alt_freq on the 7 x 6 example (test_indiv_het_obs.R,
    test_indiv_missingness.R: counts by rowSums on the genotype matrix)."""
    g = EXAMPLE_7x6
    het, na = O.gt_ind_hetero(example_gt["raw"], 7, 6)
    assert het.tolist() == np.nansum(g == 1, axis=1).tolist() and na.tolist() == np.isnan(g).sum(axis=1).tolist()
    f, v = O.alt_freq(example_gt["raw"], 7, 6)
    assert np.allclose(f, np.nansum(g, axis=0) / (2 * (~np.isnan(g)).sum(axis=0))) and v.tolist() == (2 * (~np.isnan(g)).sum(axis=0)).tolist()
