// gpu_glue.cpp — Rcpp glue between tidypopgen's R drivers and libtpg_b200 (include/tpg_b200.h).
//
// Drop this file into tidypopgen's src/ and run Rcpp::compileAttributes(): the
// `// [[Rcpp::export]]` tags regenerate R/RcppExports.R and the CallEntries table of
// src/RcppExports.cpp (the package registers routines with R_useDynamicSymbols(dll, FALSE),
// src/RcppExports.cpp:374-377, so unregistered symbols are invisible to .Call).
//
// It replaces the three per-block entry points
//   increment_king_numerator  src/snp_king.cpp:21-74   (R/RcppExports.R:76-78)
//   increment_ibs_counts      src/snp_ibs.cpp:22-74    (R/RcppExports.R:72-74)
//   increment_as_counts       src/snp_as.cpp:22-67     (R/RcppExports.R:68-70)
// by ONE call per job.  Nothing here touches the R API off the main thread, and no C++
// exception crosses a CUDA frame: the C ABI returns a code + message, turned into
// Rcpp::stop() exactly where myassert() used to throw (src/snp_king.cpp:11).
//
// Every entry takes `devices` = options(tidypopgen.gpu.devices) (0-based CUDA ordinals): one device runs
// tpg_pairwise, several run tpg_pairwise_multi — the north-star target, pairwise_king on 8 x B200, is
// options(tidypopgen.gpu.devices = 0:7); pairwise_king(x).
//
// Not BUILT in this repository (no R toolchain in the image), but type-checked: tests/test_abi.py compiles
// this file with -fsyntax-only against include/tpg_b200.h and a stand-in for <Rcpp.h> (tests/stubs), so
// every tpg_* call below agrees with the ABI, and the raw_counts path (file-backed outputs) is exercised
// through ctypes with exactly the fields this file sets (tests/test_round2_gpu.py).  See INTEGRATION.md.
#include <Rcpp.h>
#include "tpg_b200.h"

using namespace Rcpp;

namespace {

void fill_job(tpg_job& job, Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
              int stat, double block_size, bool pad_quirk, const IntegerVector& devices, std::string& path_keep) {
  std::memset(&job, 0, sizeof job);
  job.abi_version = TPG_ABI_VERSION;
  job.stat = stat;
  job.src_kind = TPG_SRC_FILE;              // stream the .bk; bigstatsr's mmap is not needed
  job.device = devices.size() > 0 ? devices.begin()[0] : 0;
  path_keep = as<std::string>(BM["backingfile"]);   // FBM RefClass field (bk path)
  job.bk_path = path_keep.c_str();
  job.nrow_total = as<double>(BM["nrow"]);
  job.ncol_total = as<double>(BM["ncol"]);
  job.ind_row = rowInd.begin();             // 1-based, exactly as R passes them
  job.n = rowInd.size();
  job.ind_col = colInd.begin();
  job.m = colInd.size();
  job.block_size = static_cast<int64_t>(block_size);
  job.pad_quirk = pad_quirk ? 1 : 0;
  job.shard_mode = TPG_SHARD_NONE;
  job.shard_world = 1;
}

// One job on options(tidypopgen.gpu.devices): a single device -> tpg_pairwise; several ->
// tpg_pairwise_multi (one host thread per GPU inside the library, counts summed over NVLink,
// finished once; shard_mode TPG_SHARD_NONE lets the library pick loci or tile sharding).
void run(const tpg_job& job, tpg_result& res, const IntegerVector& devices) {
  char err[1024] = {0};
  const int rc = devices.size() > 1
                     ? tpg_pairwise_multi(&job, &res, devices.begin(), static_cast<int32_t>(devices.size()),
                                          TPG_SHARD_NONE, err, sizeof err)
                     : tpg_pairwise(&job, &res, err, sizeof err);
  if (rc != TPG_OK) stop("tidypopgen GPU backend: %s", err);
}

}  // namespace

// options(tidypopgen.gpu.<field>) -> tpg_config (include/tpg_b200.h).  `opts` is a named list holding any
// subset of the struct's fields (r/R/snp_king_gpu.R: tpg_gpu_config() collects them); missing fields keep
// their defaults.  Called by the R drivers before every job: a few integer stores.
// [[Rcpp::export]]
void gpu_set_config(List opts) {
  tpg_config c;
  tpg_config_init(&c);
#define TPG_OPT_I(f) if (opts.containsElementNamed(#f)) c.f = as<int>(opts[#f]);
#define TPG_OPT_D(f) if (opts.containsElementNamed(#f)) c.f = as<double>(opts[#f]);
  TPG_OPT_I(gram_route) TPG_OPT_I(gram_ctas) TPG_OPT_I(gram_ksplit) TPG_OPT_I(gram_max_kblocks)
  TPG_OPT_I(gram_clusters) TPG_OPT_I(gram_no_trim) TPG_OPT_I(gram_wide) TPG_OPT_I(gram_l2_rows) TPG_OPT_I(gram_l2_cols) TPG_OPT_I(gram_order) TPG_OPT_I(pack_mode) TPG_OPT_I(pk2_i8) TPG_OPT_I(no_compact) TPG_OPT_I(gather_smem_kb) TPG_OPT_I(gather_layout)
  TPG_OPT_I(io_mode) TPG_OPT_I(no_thp) TPG_OPT_I(decode_k4_v1) TPG_OPT_I(host_threads)
  TPG_OPT_I(scan_loci_per_cta) TPG_OPT_I(scan_no_indiv) TPG_OPT_I(scan_table) TPG_OPT_I(cache_counts) TPG_OPT_I(cache_genotypes) TPG_OPT_I(result_shells)
  TPG_OPT_D(link_gbps) TPG_OPT_D(hybrid_factor)
  if (opts.containsElementNamed("device_chunk_loci")) c.device_chunk_loci = static_cast<int64_t>(as<double>(opts["device_chunk_loci"]));
  if (opts.containsElementNamed("host_chunk_mb")) c.host_chunk_mb = static_cast<int64_t>(as<double>(opts["host_chunk_mb"]));
#undef TPG_OPT_I
#undef TPG_OPT_D
  char err[512] = {0};
  if (tpg_config_set(&c, err, sizeof err) != TPG_OK) stop("tidypopgen GPU backend: %s", err);
}

// Frees what the library keeps on the given devices between jobs — staging and plane buffers, the count-matrix
// and genotype caches (tpg_config.cache_counts / cache_genotypes).  Call it after rewriting a FBM in place.
// [[Rcpp::export]]
void gpu_release(const IntegerVector& devices) {
  for (long k = 0; k < devices.size(); ++k) tpg_release(devices.begin()[k]);
}

// snp_king(): returns the n x n KING-robust matrix (R/snp_king.R:32-103).
// [[Rcpp::export]]
NumericMatrix gpu_snp_king(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
                           double block_size, const IntegerVector& devices) {
  const int n = rowInd.size();
  NumericMatrix out(n, n);                  // R owns the result buffer (column-major doubles)
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_KING, block_size, true, devices, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = out.begin();
  run(job, res, devices);
  return out;
}

// snp_ibs(type = "proportion" | "adjusted_counts"): n x n matrix (R/snp_ibs.R:89-101).
// [[Rcpp::export]]
NumericMatrix gpu_snp_ibs(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
                          bool adjusted, double block_size, const IntegerVector& devices) {
  const int n = rowInd.size();
  NumericMatrix out(n, n);
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, adjusted ? TPG_STAT_IBS_ADJUSTED : TPG_STAT_IBS_PROPORTION,
           block_size, true, devices, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = out.begin();
  run(job, res, devices);
  return out;
}

// snp_ibs(type = "raw_counts"): fills the two caller-owned double FBMs `ibs` and `valid_n`
// (bigstatsr::FBM(n, n), R/snp_ibs.R:54-56,84-85).  The reference aliases their mappings through
// FBM_RW2arma (src/snp_ibs.cpp:8-15); here the library maps the two backing files itself
// (tpg_result.out0_path / out1_path: a double FBM's .bk is header-less column-major doubles), so no
// bigstatsr header is needed.  bigstatsr's own mapping of the same files is MAP_SHARED: it sees the
// counts as soon as this call returns.
// [[Rcpp::export]]
void gpu_snp_ibs_counts(Environment ibs, Environment valid_n, Environment BM,
                        const IntegerVector& rowInd, const IntegerVector& colInd,
                        double block_size, const IntegerVector& devices) {
  const std::string p_ibs = as<std::string>(ibs["backingfile"]);
  const std::string p_val = as<std::string>(valid_n["backingfile"]);
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_IBS_RAW, block_size, true, devices, keep);
  std::memset(&res, 0, sizeof res);
  res.out0_path = p_ibs.c_str();
  res.out1_path = p_val.c_str();
  run(job, res, devices);
}

// KING + IBS proportion + allele sharing of the same (X, ind.row, ind.col) from ONE pass over the
// genotypes (tpg_pairwise_many): what pairwise_king, pairwise_ibs and pairwise_allele_sharing called one
// after the other on a gen_tibble would each decode and multiply for themselves
// (R/snp_king.R:32, R/snp_ibs.R:42, R/snp_allele_sharing.R:33).
// [[Rcpp::export]]
List gpu_snp_pairwise_many(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
                           double block_size, bool pad_quirk, const IntegerVector& devices) {
  const int n = rowInd.size();
  NumericMatrix king(n, n), ibs(n, n), as_(n, n);
  tpg_job job; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_KING, block_size, pad_quirk, devices, keep);
  const int32_t stats[3] = {TPG_STAT_KING, TPG_STAT_IBS_PROPORTION, TPG_STAT_AS};
  tpg_result res[3];
  std::memset(res, 0, sizeof res);
  res[0].out0 = king.begin();
  res[1].out0 = ibs.begin();
  res[2].out0 = as_.begin();
  char err[1024] = {0};
  if (tpg_pairwise_many(&job, stats, 3, res, devices.begin(), static_cast<int32_t>(devices.size()), TPG_SHARD_NONE,
                        err, sizeof err) != TPG_OK)
    stop("tidypopgen GPU backend: %s", err);
  return List::create(_["king"] = king, _["ibs"] = ibs, _["allele_sharing"] = as_);
}

// snp_allele_sharing(): n x n matrix with NA_real_ where no locus is jointly typed
// (R/snp_allele_sharing.R:33-82).  pad_quirk = TRUE reproduces src/snp_as.cpp:56-64.
// [[Rcpp::export]]
NumericMatrix gpu_snp_allele_sharing(Environment BM, const IntegerVector& rowInd,
                                     const IntegerVector& colInd, double block_size,
                                     bool pad_quirk, const IntegerVector& devices) {
  const int n = rowInd.size();
  NumericMatrix out(n, n);
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_AS, block_size, pad_quirk, devices, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = out.begin();
  run(job, res, devices);
  return out;
}

// ------------------------------------------------------------------ consumers of the matrices
// (include/tpg_b200.h "consumers of the matrices")

// filter_high_relatedness(): keep / order of R/filter_high_relatedness.R:26-145 for a matrix in R memory.
// [[Rcpp::export]]
List gpu_filter_high_relatedness(NumericMatrix matrix, double kings_threshold, int device = 0) {
  const int n = matrix.nrow();
  if (matrix.ncol() != n) stop("matrix should be square");
  IntegerVector keep(n), order(n);
  int64_t removed = 0;
  char err[1024] = {0};
  if (tpg_filter_high_relatedness(device, matrix.begin(), /*on_device=*/0, n, kings_threshold, keep.begin(),
                                  order.begin(), &removed, err, sizeof err) != TPG_OK)
    stop("%s", err);  // R's own message where the reference stops on a NaN comparison
  return List::create(_["keep"] = keep, _["order"] = order + 1, _["n_removed"] = (double) removed);
}

// qc_report_indiv(kings_threshold = ...): KING + filter in one call; the matrix stays in HBM unless
// want_king (R/qc_report_indiv.R:85-99 attaches it as attr(to_keep, "king")).
// [[Rcpp::export]]
List gpu_king_filter(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd, double block_size,
                     double kings_threshold, bool want_king = true, int device = 0) {
  const int n = rowInd.size();
  tpg_job job; std::string path;
  IntegerVector devs(1);
  devs.begin()[0] = device;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_KING, block_size, true, devs, path);
  IntegerVector keep(n), order(n);
  NumericMatrix king(want_king ? n : 0, want_king ? n : 0);
  int64_t removed = 0;
  char err[1024] = {0};
  if (tpg_king_filter(&job, kings_threshold, keep.begin(), order.begin(), &removed, want_king ? king.begin() : nullptr,
                      nullptr, err, sizeof err) != TPG_OK)
    stop("%s", err);
  return List::create(_["keep"] = keep, _["order"] = order + 1, _["king"] = king);
}

// indiv_inbreeding / pop_fst / pop_fis(method = "WG17") from an allele-sharing matrix in R memory
// (R/indiv_inbreeding.R:56-81, R/pop_fst.R:31-76, R/pop_fis.R:137-191).  group = group_indices(.x) - 1L.
// [[Rcpp::export]]
List gpu_wg17(NumericMatrix allele_sharing, const IntegerVector& group, int npop, int device = 0) {
  const int n = allele_sharing.nrow();
  NumericVector inb(n), fst(npop + 1), fis(npop + 1), mb(2);
  char err[1024] = {0};
  const bool grouped = npop >= 2;
  if (tpg_wg17(device, allele_sharing.begin(), 0, n, grouped ? group.begin() : nullptr, grouped ? npop : 1, inb.begin(),
               grouped ? fst.begin() : nullptr, grouped ? fis.begin() : nullptr, mb.begin(), err, sizeof err) != TPG_OK)
    stop("%s", err);
  return List::create(_["inbreeding"] = inb, _["fst"] = fst, _["fis"] = fis, _["mb"] = mb);
}

// gt_ind_hetero + alt_freq_dip_pseudo_cpp + the NA row of big_counts in one pass
// (src/gt_ind_hetero.cpp:12-42, src/alt_freq_dip_pseudo_cpp.cpp:8-58, R/loci_missingness.R:49-60).
// [[Rcpp::export]]
List gpu_scan(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd, const NumericVector& ploidy,
              bool as_counts = false, int device = 0) {
  const int n = rowInd.size(), m = colInd.size();
  tpg_job job; std::string path;
  IntegerVector devs(1);
  devs.begin()[0] = device;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_AS, 1, true, devs, path);
  NumericVector code = BM["code256"];          // these callers DO decode through code256
  IntegerVector het(n), na(n), loc_na(m);
  NumericVector loc0(m), loc1(m);
  char err[1024] = {0};
  if (tpg_scan(&job, code.begin(), ploidy.begin(), as_counts ? 1 : 0, het.begin(), na.begin(), loc0.begin(), loc1.begin(),
               loc_na.begin(), nullptr, err, sizeof err) != TPG_OK)
    stop("tidypopgen GPU backend: %s", err);
  return List::create(_["het_n"] = het, _["na_n"] = na, _[as_counts ? "n_alt" : "freq"] = loc0, _["n_valid"] = loc1,
                      _["n_na"] = loc_na);
}
