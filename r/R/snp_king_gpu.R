# Replacement bodies for the three L4 drivers.  Signatures, defaults and the class check
# are the reference's, verbatim (R/snp_king.R:32-41, R/snp_ibs.R:42-50,
# R/snp_allele_sharing.R:33-38); only the block loop + .Call per block + big_apply finish
# are replaced by one .Call.  `block.size` stays in the signature: it no longer sizes any
# buffer, but allele sharing still depends on it through the reference's padding quirk.

# options(tidypopgen.gpu.*) -> the library's tpg_config (include/tpg_b200.h), applied before every job.
#   tidypopgen.gpu.devices   0-based CUDA ordinals; several = one job spread over them (tpg_pairwise_multi),
#                            e.g. options(tidypopgen.gpu.devices = 0:7) for pairwise_king on 8 x B200
#   tidypopgen.gpu.<field>   any field of tpg_config: gram_route, pack_mode, io_mode, host_threads,
#                            cache_counts, host_chunk_mb, ...
#   tidypopgen.gpu.reference_pad_quirk  FALSE = block-size independent allele sharing
tpg_gpu_devices <- function() {
  as.integer(getOption("tidypopgen.gpu.devices", getOption("tidypopgen.gpu.device", 0L)))
}
tpg_gpu_config <- function() {
  fields <- c("gram_route", "gram_ctas", "gram_ksplit", "gram_max_kblocks", "gram_clusters",
              "gram_no_trim", "gram_wide", "gram_l2_rows", "gram_l2_cols", "gram_order", "pack_mode", "pk2_i8", "no_compact", "gather_smem_kb", "gather_layout", "io_mode", "no_thp", "decode_k4_v1",
              "host_threads", "scan_loci_per_cta", "scan_no_indiv", "scan_table", "cache_counts", "cache_genotypes", "result_shells",
              "device_chunk_loci", "host_chunk_mb", "link_gbps", "hybrid_factor")
  opts <- lapply(stats::setNames(fields, fields), function(f) getOption(paste0("tidypopgen.gpu.", f)))
  gpu_set_config(opts[!vapply(opts, is.null, logical(1))])
  invisible(NULL)
}

# Drops the buffers and caches the backend keeps on the GPUs (after a FBM was rewritten in place, or to give the
# memory back): tpg_gpu_release()
tpg_gpu_release <- function(devices = tpg_gpu_devices()) {
  gpu_release(as.integer(devices))
  invisible(NULL)
}

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
  tpg_gpu_config()
  gpu_snp_king(X, ind.row, ind.col, block.size, tpg_gpu_devices())
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
  tpg_gpu_config()
  devices <- tpg_gpu_devices()
  if (type == "raw_counts") {
    n <- length(ind.row)
    IBS <- bigstatsr::FBM(n, n, init = 0) # nolint
    IBS_valid_loci <- bigstatsr::FBM(n, n, init = 0) # nolint
    # the library maps IBS$backingfile / IBS_valid_loci$backingfile and writes the counts there
    gpu_snp_ibs_counts(IBS, IBS_valid_loci, X, ind.row, ind.col, block.size, devices)
    return(list(ibs = IBS, valid_n = IBS_valid_loci))
  }
  gpu_snp_ibs(X, ind.row, ind.col, type == "adjusted_counts", block.size, devices)
}

snp_allele_sharing <- function(
  X, # nolint start
  ind.row = bigstatsr::rows_along(X),
  ind.col = bigstatsr::cols_along(X),
  block.size = bigstatsr::block_size(nrow(X))
) {
  # nolint end
  tpg_gpu_config()
  gpu_snp_allele_sharing(X, ind.row, ind.col, block.size,
                         pad_quirk = getOption("tidypopgen.gpu.reference_pad_quirk", TRUE),
                         devices = tpg_gpu_devices())
}

# Not in the reference: KING, IBS proportion and allele sharing of one (X, ind.row, ind.col) from a single
# pass over the genotypes.  options(tidypopgen.gpu.cache_counts = 1L) gives the same saving without a new
# function: the second and third of pairwise_king / pairwise_ibs / pairwise_allele_sharing on the same
# gen_tibble finish from the count matrices the first one left on the device.
snp_pairwise_many <- function(X, ind.row = bigstatsr::rows_along(X), ind.col = bigstatsr::cols_along(X), # nolint
                              block.size = bigstatsr::block_size(nrow(X))) {
  if (!inherits(X, "FBM.code256")) stop("X should be a FBM.code256 matrix")
  tpg_gpu_config()
  gpu_snp_pairwise_many(X, ind.row, ind.col, block.size,
                        getOption("tidypopgen.gpu.reference_pad_quirk", TRUE), tpg_gpu_devices())
}
# pairwise_king / pairwise_ibs / pairwise_allele_sharing / pairwise_grm
# (R/pairwise_*.R) are unchanged: they only call the three drivers above.

# ---- consumers of the matrices (SURVEY.md section 8f) --------------------------------------------
# filter_high_relatedness(): the names handling (:33-44), the single-individual branch (:46-53) and
# the return value (:138-144) stay as in R/filter_high_relatedness.R; the O(n^2)-per-pair loop
# (:56-135) becomes one call.
filter_high_relatedness_core <- function(matrix, var_names, kings_threshold,
                                         device = getOption("tidypopgen.gpu.device", 0L)) {
  res <- gpu_filter_high_relatedness(matrix, kings_threshold, device)
  passed_filter <- var_names[res$order][as.logical(res$keep[res$order])]
  to_remove <- var_names[!var_names %in% passed_filter]
  list(passed_filter, to_remove, var_names %in% passed_filter)
}

# indiv_het_obs.vctrs_bigSNP (R/indiv_het_obs.R:33-65): one scan instead of big_apply(gt_ind_hetero)
indiv_het_obs_gpu <- function(.x, as_counts = FALSE) {
  X <- attr(.x, "fbm") # nolint
  res <- gpu_scan(X, vctrs::vec_data(.x), attr(.x, "loci")$big_index, indiv_ploidy(.x),
                  device = getOption("tidypopgen.gpu.device", 0L))
  if (!as_counts) {
    return(res$het_n / (ncol(X) - res$na_n))
  }
  cbind(het_n = res$het_n, na_n = res$na_n)
}

# pop_fst (R/pop_fst.R:31-76) from a precomputed or freshly computed allele-sharing matrix
pop_fst_gpu <- function(.x, include_global = FALSE, allele_sharing_mat = NULL) {
  if (!inherits(.x, "grouped_df")) stop(".x should be a grouped gen_tibble")
  if (is.null(allele_sharing_mat)) allele_sharing_mat <- pairwise_allele_sharing(.x, as_matrix = TRUE)
  n_pop <- max(dplyr::group_indices(.x))
  res <- gpu_wg17(allele_sharing_mat, dplyr::group_indices(.x) - 1L, n_pop,
                  device = getOption("tidypopgen.gpu.device", 0L))
  fst <- if (include_global) res$fst else res$fst[seq_len(n_pop)]
  names(fst) <- c(dplyr::group_keys(.x) %>% dplyr::pull(1), if (include_global) "global")
  fst
}
