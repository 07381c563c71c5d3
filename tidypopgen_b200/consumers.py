"""Consumers of the pairwise matrices, run where the matrix is:
filter_high_relatedness (the KING consumer behind qc_report_indiv) and the Weir & Goudet
(2017) statistics built on allele sharing (indiv_inbreeding, pop_fst, pop_fis).

Same names, arguments, return values and error text as the reference
(R/filter_high_relatedness.R:26-145, R/indiv_inbreeding.R:23-95, R/pop_fst.R:31-76,
R/pop_fis.R:52-191); the arithmetic is in libtpg_b200.so (csrc/relate.cu, csrc/wg17.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .pairwise import GenTibble


def _square(matrix):
    a = np.asarray(matrix)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("matrix should be a square matrix")
    return np.asfortranarray(a, dtype=np.float64)


def filter_high_relatedness(matrix, x: GenTibble | None = None, kings_threshold=None, verbose: bool = False,
                            device: int = 0):
    """R/filter_high_relatedness.R:26-145.  Returns [passed_filter, to_remove, keep]: the ids
    to retain (in the order the filter visited them), the ids to remove, and a boolean per
    individual in the original order."""
    if kings_threshold is None:
        raise TypeError("argument \"kings_threshold\" is missing, with no default")
    a = _square(matrix)
    n = a.shape[0]
    names = getattr(matrix, "names", None)
    if names is None:  # :33-41
        names = [str(k + 1) for k in range(n)] if x is None else list(x.id)
    names = list(names)
    keep = np.zeros(n, np.int32)
    order = np.zeros(n, np.int32)
    removed = C.c_int64(0)
    eb = L.errbuf()
    rc = L.lib().tpg_filter_high_relatedness(device, a.ctypes.data, 0, n, float(kings_threshold), keep.ctypes.data,
                                             order.ctypes.data, C.byref(removed), eb, len(eb))
    L.check(rc, eb)
    if verbose and removed.value == 0:
        print(f"All correlations <= {kings_threshold}")
    kept = keep.astype(bool)
    passed = [names[k] for k in order if kept[k]]
    passed_set = set(passed)
    to_remove = [nm for nm in names if nm not in passed_set]
    return [passed, to_remove, np.array([nm in passed_set for nm in names])]


def king_filter(x: GenTibble, kings_threshold, return_king: bool = False, block_size: int | None = None, **kw):
    """pairwise_king + filter_high_relatedness in one native call; the n x n matrix stays in HBM
    (returned only when return_king=True).  Returns ([passed, to_remove, keep], king or None)."""
    from . import snp

    if block_size is None:
        block_size = snp.block_size(len(x))
    job, keepalive = snp.make_job(x.fbm, x.row_index, x.big_index, L.STAT_KING, block_size, **kw)
    n = len(x)
    keep, order = np.zeros(n, np.int32), np.zeros(n, np.int32)
    removed = C.c_int64(0)
    king = np.zeros((n, n), np.float64, order="F") if return_king else None
    st = L.Stats()
    eb = L.errbuf()
    rc = L.lib().tpg_king_filter(C.byref(job), float(kings_threshold), keep.ctypes.data, order.ctypes.data,
                                 C.byref(removed), None if king is None else king.ctypes.data, C.byref(st), eb, len(eb))
    del keepalive
    L.check(rc, eb)
    snp.last_stats = st.as_dict()
    kept = keep.astype(bool)
    names = list(x.id)
    passed = [names[k] for k in order if kept[k]]
    return [passed, [nm for k, nm in enumerate(names) if not kept[k]], kept], king


def king_filter_groups(x: GenTibble, group, kings_threshold, return_king: bool = False, block_size: int | None = None, **kw):
    """One pairwise_king + filter_high_relatedness PER GROUP (R/qc_report_indiv.R:161,169-180) in one native
    call (tpg_king_filter_groups): the library sorts the individuals by group, one KING job computes only the
    tiles that hold a same-group pair, and every diagonal block is filtered in HBM.
    group: one label per individual.  Returns (keep per individual [job order], {label: KING matrix} or None,
    labels in dplyr's group_keys order)."""
    from . import snp

    labels = sorted(set(group))                       # dplyr::group_keys sorts the keys
    code = {lab: k for k, lab in enumerate(labels)}
    g = np.ascontiguousarray([code[v] for v in group], dtype=np.int32)
    n = len(x)
    if len(g) != n:
        raise ValueError("group must have one entry per individual")
    if block_size is None:
        block_size = snp.block_size(len(x))
    job, keepalive = snp.make_job(x.fbm, x.row_index, x.big_index, L.STAT_KING, block_size, **kw)
    keep = np.zeros(n, np.int32)
    removed = np.zeros(len(labels), np.int64)
    sizes = np.bincount(g, minlength=len(labels))
    king = np.zeros(int((sizes.astype(np.int64) ** 2).sum()), np.float64) if return_king else None
    st = L.Stats()
    eb = L.errbuf()
    rc = L.lib().tpg_king_filter_groups(C.byref(job), g.ctypes.data, len(labels), float(kings_threshold), keep.ctypes.data,
                                        removed.ctypes.data, None if king is None else king.ctypes.data, C.byref(st),
                                        eb, len(eb))
    del keepalive
    L.check(rc, eb)
    snp.last_stats = st.as_dict()
    mats = None
    if king is not None:
        mats, off = {}, 0
        for lab, ng in zip(labels, sizes):
            mats[lab] = king[off:off + ng * ng].reshape(ng, ng).T.copy()   # column-major n_g x n_g
            off += ng * ng
    return keep.astype(bool), mats, labels


def _wg17(as_mat, group, npop, want_inb=False, want_fst=False, want_fis=False, device: int = 0, x: GenTibble | None = None,
          block_size: int | None = None, **kw):
    """tpg_wg17 on a given allele-sharing matrix, or — as_mat None — tpg_as_wg17: the matrix is
    computed from x's genotypes and reduced without leaving HBM."""
    n = len(x) if as_mat is None else np.asarray(as_mat).shape[0]
    inb = np.zeros(n) if want_inb else None
    fst = np.zeros(npop + 1) if want_fst else None
    fis = np.zeros(npop + 1) if want_fis else None
    mb = np.zeros(2)
    g = None if group is None else np.ascontiguousarray(group, dtype=np.int32)
    ptr = lambda a: None if a is None else a.ctypes.data  # noqa: E731
    eb = L.errbuf()
    if as_mat is None:
        from . import snp

        if block_size is None:
            block_size = snp.block_size(len(x))
        job, keepalive = snp.make_job(x.fbm, x.row_index, x.big_index, L.STAT_AS, block_size, device=device, **kw)
        st = L.Stats()
        rc = L.lib().tpg_as_wg17(C.byref(job), ptr(g), npop, ptr(inb), ptr(fst), ptr(fis), mb.ctypes.data, None,
                                 C.byref(st), eb, len(eb))
        del keepalive
        snp.last_stats = st.as_dict()
    else:
        a = _square(as_mat)
        rc = L.lib().tpg_wg17(device, a.ctypes.data, 0, n, ptr(g), npop, ptr(inb), ptr(fst), ptr(fis), mb.ctypes.data,
                              eb, len(eb))
    L.check(rc, eb)
    return inb, fst, fis, mb


def _groups(x: GenTibble):
    """factor(group_indices(.x)): populations in sorted order of their keys, 0-based codes."""
    if x.population is None:
        raise ValueError(".x should be a grouped gen_tibble")
    keys = sorted(set(x.population))
    code = {k: i for i, k in enumerate(keys)}
    return keys, np.array([code[p] for p in x.population], dtype=np.int32)


def indiv_inbreeding(x: GenTibble, method: str = "WG17", allele_sharing_mat=None, **kw):
    """R/indiv_inbreeding.R:23-95.  Ungrouped: one vector; grouped: a list with one vector per
    population, each computed from that population's own allele-sharing matrix."""
    if method != "WG17":
        raise ValueError("'arg' should be 'WG17'")
    if x.population is not None:
        if allele_sharing_mat is not None:
            raise ValueError("allele_sharing_mat can not be provided for grouped_df objects")
        keys, code = _groups(x)
        out = []
        for k in range(len(keys)):
            sub = x[np.flatnonzero(code == k)]
            sub.population = None
            out.append(indiv_inbreeding(sub, method=method, **kw))
        return out
    inb, _, _, _ = _wg17(allele_sharing_mat, None, 1, want_inb=True, x=x, **kw)
    return inb


def pop_fst(x: GenTibble, include_global: bool = False, allele_sharing_mat=None, **kw):
    """R/pop_fst.R:31-76: population-specific Fst of Weir & Goudet (2017); a dict keyed by
    population (plus "global" when asked)."""
    keys, code = _groups(x)
    _, fst, _, _ = _wg17(allele_sharing_mat, code, len(keys), want_fst=True, x=x, **kw)
    out = {k: fst[i] for i, k in enumerate(keys)}
    if include_global:
        out["global"] = fst[len(keys)]
    return out


def pop_fis(x: GenTibble, method: str = "WG17", include_global: bool = False, allele_sharing_mat=None, **kw):
    """pop_fis(method = "WG17") — R/pop_fis.R:137-191 (the Nei87 method is a per-locus
    statistic outside this path)."""
    if method != "WG17":
        raise NotImplementedError("only method = 'WG17' is built on the pairwise path")
    keys, code = _groups(x)
    _, _, fis, _ = _wg17(allele_sharing_mat, code, len(keys), want_fis=True, x=x, **kw)
    out = {k: fis[i] for i, k in enumerate(keys)}
    if include_global:
        out["global"] = fis[len(keys)]
    return out


# ------------------------------------------------------------------ O(n m) scans through code256

CODE_012 = np.array([0.0, 1.0, 2.0] + [np.nan] * 253)  # bigsnpr::CODE_012, the gen_tibble default


def scan_counts(X, ind_row=None, ind_col=None, code256=None, ploidy=None, as_counts=False, **kw):
    """One tpg_scan call: dict(het_n, na_n, loc0, n_valid, loc_na, stats)."""
    from . import snp

    job, keep = snp.make_job(X, ind_row, ind_col, L.STAT_AS, 1, **kw)
    n = job.n if job.ind_row else X.nrow
    m = job.m if job.ind_col else X.ncol
    if n <= 0 or m <= 0:
        raise ValueError("ind.row and ind.col must not be empty")
    code = np.ascontiguousarray(CODE_012 if code256 is None else code256, dtype=np.float64)
    if code.shape != (256,):
        raise ValueError("code256 must hold 256 values")
    pl = None if ploidy is None else np.ascontiguousarray(ploidy, dtype=np.float64)
    if pl is not None and pl.shape != (n,):
        raise ValueError("ploidy must have one value per individual")
    het, na = np.zeros(n, np.int32), np.zeros(n, np.int32)
    loc0, loc1, loc_na = np.zeros(m), np.zeros(m), np.zeros(m, np.int32)
    st = L.Stats()
    eb = L.errbuf()
    rc = L.lib().tpg_scan(C.byref(job), code.ctypes.data, None if pl is None else pl.ctypes.data, int(as_counts),
                          het.ctypes.data, na.ctypes.data, loc0.ctypes.data, loc1.ctypes.data, loc_na.ctypes.data,
                          C.byref(st), eb, len(eb))
    del keep
    L.check(rc, eb)
    return {"het_n": het, "na_n": na, "loc0": loc0, "n_valid": loc1, "loc_na": loc_na, "stats": st.as_dict()}


def _code(x: GenTibble):
    return x.fbm.code256 if getattr(x.fbm, "code256", None) is not None else CODE_012


def indiv_het_obs(x: GenTibble, as_counts: bool = False, **kw):
    """R/indiv_het_obs.R:33-65: het_n / (ncol(X) - na_n) — ncol of the FBM, as the reference
    writes it — or the (het_n, na_n) count matrix."""
    r = scan_counts(x.fbm, x.row_index, x.big_index, code256=_code(x), **kw)
    if as_counts:
        return np.stack([r["het_n"], r["na_n"]], axis=1)  # columns het_n, na_n
    return r["het_n"] / (x.fbm.ncol - r["na_n"])


def indiv_missingness(x: GenTibble, as_counts: bool = False, **kw):
    """R/indiv_missingness.R:35-77: na_n (/ count_loci(x))."""
    r = scan_counts(x.fbm, x.row_index, x.big_index, code256=_code(x), **kw)
    return r["na_n"] if as_counts else r["na_n"] / len(x.big_index)


def loci_alt_freq(x: GenTibble, as_counts: bool = False, ploidy=None, **kw):
    """loci_alt_freq for diploid / pseudohaploid data (R/loci_alt_freq.R:328-379 over
    src/alt_freq_dip_pseudo_cpp.cpp): the frequency vector, or the (n_alt, n_valid) matrix."""
    r = scan_counts(x.fbm, x.row_index, x.big_index, code256=_code(x), ploidy=ploidy, as_counts=as_counts, **kw)
    if as_counts:
        return np.stack([r["loc0"], r["n_valid"]], axis=1)
    return r["loc0"]


def loci_missingness(x: GenTibble, as_counts: bool = False, **kw):
    """R/loci_missingness.R:38-74: missing individuals per locus (/ number of individuals)."""
    r = scan_counts(x.fbm, x.row_index, x.big_index, code256=_code(x), **kw)
    return r["loc_na"] if as_counts else r["loc_na"] / len(x)


def qc_report_indiv(x: GenTibble, kings_threshold=None, grouped: bool = False, **kw):
    """qc_report_indiv for a diploid gen_tibble (R/qc_report_indiv.R:45-101; grouped_df method :110-181):
    het_obs, missingness and — with kings_threshold (a number, "first" = 0.177 or "second" =
    0.088) — to_keep from filter_high_relatedness on the KING matrix, which is attached as
    "king".  grouped=True is the method for a gen_tibble grouped by `population`: one KING and one
    filter per group, and to_keep assembled exactly as the reference does — the groups' flags
    concatenated in order of first appearance (:174-177), which lines up with the rows when the
    individuals of a group are contiguous."""
    if kings_threshold is not None and not isinstance(kings_threshold, (int, float)):
        if kings_threshold not in ("first", "second"):
            raise ValueError("kings_threshold must be a numeric or one of 'first' or 'second'")
        kings_threshold = 0.177 if kings_threshold == "first" else 0.088
    n_loci = len(x.big_index)
    r = scan_counts(x.fbm, x.row_index, x.big_index, code256=_code(x), **kw)
    rep = {"het_obs": r["het_n"] / (n_loci - r["na_n"]), "missingness": r["na_n"] / n_loci}
    if kings_threshold is not None and grouped:
        if x.population is None:
            raise ValueError(".x should be a grouped gen_tibble")
        keep, mats, _ = king_filter_groups(x, list(x.population), kings_threshold, return_king=True, **kw)
        pop = np.asarray(x.population, dtype=object)
        first_seen = list(dict.fromkeys(x.population))                   # unique(qc_report$group)
        rep["group"] = list(x.population)
        rep["to_keep"] = np.concatenate([keep[pop == lab] for lab in first_seen])
        rep["to_keep_by_individual"] = keep
        rep["id"] = list(x.id)
        rep["king"] = mats
    elif kings_threshold is not None:
        # one native call: KING is computed, filtered in HBM and shipped once for attr(to_keep, "king")
        rel, king = king_filter(x, kings_threshold, return_king=True, **kw)
        rep["to_keep"] = rel[2]
        rep["id"] = list(x.id)
        rep["king"] = king
    return rep
