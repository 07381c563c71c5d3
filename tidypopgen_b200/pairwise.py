"""gen_tibble-level API: pairwise_king, pairwise_ibs, pairwise_allele_sharing, pairwise_grm.

Mirrors the reference's L5 wrappers (R/pairwise_king.R:33-56, R/pairwise_ibs.R:34-74,
R/pairwise_allele_sharing.R:26-49, R/pairwise_grm.R:30-51): extract (FBM, ind.row,
ind.col) from the gen_tibble, call the snp_* driver, attach dimnames / class, or tidy the
matrix into (item1, item2, value) rows in `combn` order (R/gt_helper_functions.R:26-38).

`GenTibble` is the minimum of tidypopgen's gen_tibble this path reads: `id`, the FBM behind
`genotypes`, the row indices `vctrs::vec_data(x$genotypes)` and the loci table's
`big_index` (R/pairwise_king.R:39-41) — both 1-based, arbitrary order / subset.  Building,
merging, saving and the rest of gen_tibble stay the reference's R code.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from . import snp
from .fbm import FBMCode256


class PairwiseMatrix(np.ndarray):
    """An n x n matrix with dimnames and R class c("pairwise_matrix", "matrix", "array")."""

    r_class = ("pairwise_matrix", "matrix", "array")

    def __new__(cls, a, names):
        obj = np.asarray(a).view(cls)
        obj.names = list(names)
        return obj

    def __array_finalize__(self, obj):
        self.names = getattr(obj, "names", None)


@dataclass
class PairwiseTbl:
    """tidy_dist_matrix(): one row per unordered pair, order (1,2),(1,3),...,(2,3),..."""

    item1: list
    item2: list
    value: np.ndarray
    r_class: tuple = ("pairwise_tbl", "tbl_df", "tbl", "data.frame")

    def __len__(self):
        return len(self.value)


def tidy_dist_matrix(mat: np.ndarray, names) -> PairwiseTbl:
    """R/gt_helper_functions.R:26-38: xy <- t(combn(colnames(mat), 2)); value = mat[xy]."""
    a = np.asarray(mat)
    if a.ndim != 2:
        raise TypeError("mat should be a matrix")
    i, j = np.triu_indices(a.shape[0], k=1)  # row-major upper triangle == combn order
    names = list(names)
    return PairwiseTbl([names[k] for k in i], [names[k] for k in j], np.array(a[i, j]))


@dataclass
class GenTibble:
    id: list
    fbm: FBMCode256
    row_index: np.ndarray  # vctrs::vec_data(x$genotypes): 1-based rows of the FBM
    big_index: np.ndarray  # show_loci(x)$big_index: 1-based columns of the FBM
    loci_names: list = field(default_factory=list)
    population: list | None = None

    def __len__(self):
        return len(self.id)

    def __getitem__(self, rows):
        """x[rows, ] — 0-based positions (numpy style) or a boolean mask."""
        rows = np.asarray(rows)
        if rows.dtype == bool:
            rows = np.flatnonzero(rows)
        return replace(self, id=[self.id[k] for k in rows], row_index=self.row_index[rows],
                       population=None if self.population is None else [self.population[k] for k in rows])

    def select_loci(self, which):
        """select_loci(x, which) — 0-based positions in the current loci table, or names."""
        which = list(which)
        if which and isinstance(which[0], str):
            pos = {nm: k for k, nm in enumerate(self.loci_names)}
            which = [pos[w] for w in which]
        which = np.asarray(which, dtype=np.int64)
        return replace(self, big_index=self.big_index[which],
                       loci_names=[self.loci_names[k] for k in which] if self.loci_names else [])


def gen_tibble(x, indiv_id=None, loci_names=None, population=None, backingfile=None) -> GenTibble:
    """gen_tibble(x = <matrix>, ...) for an in-memory dosage matrix (NaN = missing), or
    wrap an existing FBMCode256."""
    fbm = x if isinstance(x, FBMCode256) else FBMCode256.from_matrix(x, backingfile)
    n, m = fbm.nrow, fbm.ncol
    ids = list(indiv_id) if indiv_id is not None else [f"ind{k + 1}" for k in range(n)]
    lnames = list(loci_names) if loci_names is not None else [f"rs{k + 1}" for k in range(m)]
    return GenTibble(ids, fbm, np.arange(1, n + 1, dtype=np.int32), np.arange(1, m + 1, dtype=np.int32),
                     lnames, None if population is None else list(population))


def _default_bs(x: GenTibble) -> int:
    return snp.block_size(len(x))  # bigstatsr::block_size(nrow(x)): rows of the gen_tibble


def pairwise_king(x: GenTibble, as_matrix: bool = False, block_size: int | None = None, **kw):
    """R/pairwise_king.R:33-56 (note: passes block_size(nrow(x)) un-multiplied, unlike
    snp_king's own default — kept as in the reference)."""
    if block_size is None:
        block_size = _default_bs(x)
    king = snp.snp_king(x.fbm, ind_row=x.row_index, ind_col=x.big_index, block_size=block_size, **kw)
    if as_matrix:
        return PairwiseMatrix(king, x.id)
    return tidy_dist_matrix(king, x.id)


def pairwise_ibs(x: GenTibble, as_matrix: bool = False, type: str = "proportion",
                 block_size: int | None = None, **kw):
    """R/pairwise_ibs.R:34-74.  type = "raw_counts" returns the two FBMs tagged with
    `indiv_names` and ignores as_matrix, as the reference does."""
    if block_size is None:
        block_size = _default_bs(x)
    res = snp.snp_ibs(x.fbm, ind_row=x.row_index, ind_col=x.big_index, type=type, block_size=block_size, **kw)
    if isinstance(res, dict):
        res["ibs"].attrs["indiv_names"] = list(x.id)
        res["valid_n"].attrs["indiv_names"] = list(x.id)
        return res
    if as_matrix:
        return PairwiseMatrix(res, x.id)
    return tidy_dist_matrix(res, x.id)


def pairwise_allele_sharing(x: GenTibble, as_matrix: bool = False, block_size: int | None = None, **kw):
    """R/pairwise_allele_sharing.R:26-49."""
    if block_size is None:
        block_size = _default_bs(x)
    a = snp.snp_allele_sharing(x.fbm, ind_row=x.row_index, ind_col=x.big_index, block_size=block_size, **kw)
    if as_matrix:
        return PairwiseMatrix(a, x.id)
    return tidy_dist_matrix(a, x.id)


def pairwise_grm(x: GenTibble, allele_sharing_mat=None, block_size: int | None = None, **kw):
    """R/pairwise_grm.R:30-51: GRM = 2 (AS - mb) / (1 - mb), mb = mean off-diagonal AS."""
    if block_size is None:
        block_size = _default_bs(x)
    if allele_sharing_mat is None:
        g = snp.snp_grm(x.fbm, ind_row=x.row_index, ind_col=x.big_index, block_size=block_size, **kw)
    else:
        g, _ = snp.grm_from_allele_sharing(np.asarray(allele_sharing_mat), device=kw.get("device", 0))
    return PairwiseMatrix(g, x.id)
