"""B200-native pairwise-individual similarity path of tidypopgen.

KING-robust kinship, identity by state, allele sharing and the Weir & Goudet GRM, computed
by libtpg_b200.so (hand-written sm_100a kernels behind the C ABI of include/tpg_b200.h).
This package is the thin host-side mirror of the reference's R interface for that path;
everything else of tidypopgen stays the reference's R code.
"""
from .fbm import FBMCode256, FBMDouble  # noqa: F401
from .pairwise import (  # noqa: F401
    GenTibble,
    PairwiseMatrix,
    PairwiseTbl,
    gen_tibble,
    pairwise_allele_sharing,
    pairwise_grm,
    pairwise_ibs,
    pairwise_king,
    tidy_dist_matrix,
)
from .consumers import (  # noqa: F401
    filter_high_relatedness,
    indiv_het_obs,
    indiv_inbreeding,
    indiv_missingness,
    king_filter,
    king_filter_groups,
    loci_alt_freq,
    loci_missingness,
    pop_fis,
    pop_fst,
    qc_report_indiv,
    scan_counts,
)
from .snp import (  # noqa: F401
    block_size,
    cut_by_size,
    grm_from_allele_sharing,
    snp_allele_sharing,
    snp_allele_sharing_counts,
    snp_grm,
    snp_ibs,
    snp_king,
    snp_king_counts,
    snp_pairwise_many,
)
from ._lib import config  # noqa: F401

__version__ = "0.2.0"
