# Replacement bodies for the three L4 drivers.  Signatures, defaults and the class check
# are the reference's, verbatim (R/snp_king.R:32-41, R/snp_ibs.R:42-50,
# R/snp_allele_sharing.R:33-38); only the block loop + .Call per block + big_apply finish
# are replaced by one .Call.  `block.size` stays in the signature: it no longer sizes any
# buffer, but allele sharing still depends on it through the reference's padding quirk.

snp_king <- function(
  X, # nolint start
  ind.row = bigstatsr::rows_along(X),
  ind.col = bigstatsr::cols_along(X),
  block.size = bigstatsr::block_size(nrow(X)) * 4
) {
  # nolint end
  if (!inherits(X, "FBM.code256")) {
    stop("X should be a FBM.code256 matrix")
  }
  gpu_snp_king(X, ind.row, ind.col, block.size,
               device = getOption("tidypopgen.gpu.device", 0L))
}

snp_ibs <- function(
  X, # nolint start
  ind.row = bigstatsr::rows_along(X),
  ind.col = bigstatsr::cols_along(X),
  type = c("proportion", "adjusted_counts", "raw_counts"),
  block.size = bigstatsr::block_size(nrow(X))
) {
  # nolint end
  type <- match.arg(type)
  device <- getOption("tidypopgen.gpu.device", 0L)
  if (type == "raw_counts") {
    n <- length(ind.row)
    IBS <- bigstatsr::FBM(n, n, init = 0) # nolint
    IBS_valid_loci <- bigstatsr::FBM(n, n, init = 0) # nolint
    gpu_snp_ibs_counts(IBS, IBS_valid_loci, X, ind.row, ind.col, block.size, device)
    return(list(ibs = IBS, valid_n = IBS_valid_loci))
  }
  gpu_snp_ibs(X, ind.row, ind.col, type == "adjusted_counts", block.size, device)
}

snp_allele_sharing <- function(
  X, # nolint start
  ind.row = bigstatsr::rows_along(X),
  ind.col = bigstatsr::cols_along(X),
  block.size = bigstatsr::block_size(nrow(X))
) {
  # nolint end
  gpu_snp_allele_sharing(X, ind.row, ind.col, block.size,
                         pad_quirk = getOption("tidypopgen.gpu.reference_pad_quirk", TRUE),
                         device = getOption("tidypopgen.gpu.device", 0L))
}
# pairwise_king / pairwise_ibs / pairwise_allele_sharing / pairwise_grm
# (R/pairwise_*.R) are unchanged: they only call the three drivers above.
