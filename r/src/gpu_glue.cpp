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
// Not BUILT in this repository (no R toolchain in the image), but type-checked: tests/test_abi.py compiles
// this file with -fsyntax-only against include/tpg_b200.h and a stand-in for <Rcpp.h> (tests/stubs), so
// every tpg_* call below agrees with the ABI.  See INTEGRATION.md.
#include <Rcpp.h>
#include "tpg_b200.h"

using namespace Rcpp;

namespace {

void fill_job(tpg_job& job, Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
              int stat, double block_size, bool pad_quirk, int device, std::string& path_keep) {
  std::memset(&job, 0, sizeof job);
  job.abi_version = TPG_ABI_VERSION;
  job.stat = stat;
  job.src_kind = TPG_SRC_FILE;              // stream the .bk; bigstatsr's mmap is not needed
  job.device = device;
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

void run(const tpg_job& job, tpg_result& res) {
  char err[1024] = {0};
  const int rc = tpg_pairwise(&job, &res, err, sizeof err);
  if (rc != TPG_OK) stop("tidypopgen GPU backend: %s", err);
}

}  // namespace

// snp_king(): returns the n x n KING-robust matrix (R/snp_king.R:32-103).
// [[Rcpp::export]]
NumericMatrix gpu_snp_king(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
                           double block_size, int device = 0) {
  const int n = rowInd.size();
  NumericMatrix out(n, n);                  // R owns the result buffer (column-major doubles)
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_KING, block_size, true, device, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = out.begin();
  run(job, res);
  return out;
}

// snp_ibs(type = "proportion" | "adjusted_counts"): n x n matrix (R/snp_ibs.R:89-101).
// [[Rcpp::export]]
NumericMatrix gpu_snp_ibs(Environment BM, const IntegerVector& rowInd, const IntegerVector& colInd,
                          bool adjusted, double block_size, int device = 0) {
  const int n = rowInd.size();
  NumericMatrix out(n, n);
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, adjusted ? TPG_STAT_IBS_ADJUSTED : TPG_STAT_IBS_PROPORTION,
           block_size, true, device, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = out.begin();
  run(job, res);
  return out;
}

// snp_ibs(type = "raw_counts"): fills the two caller-owned double FBMs `ibs` and `valid_n`
// (bigstatsr::FBM(n, n), R/snp_ibs.R:54-56,84-85) through their writable mappings, the same
// FBM_RW2arma aliasing the reference uses (src/snp_ibs.cpp:8-15).
// [[Rcpp::export]]
void gpu_snp_ibs_counts(Environment ibs, Environment valid_n, Environment BM,
                        const IntegerVector& rowInd, const IntegerVector& colInd,
                        double block_size, int device = 0) {
  // XPtr<FBM_RW> xp = ibs["address_rw"]; double* p = (double*) xp->matrix();   (bigstatsr)
  // kept as comments here: needs <bigstatsr/BMAcc.h>, which this repository does not vendor.
  double* p_ibs = nullptr;   // = static_cast<double*>(XPtr<FBM_RW>(ibs["address_rw"])->matrix());
  double* p_val = nullptr;   // = static_cast<double*>(XPtr<FBM_RW>(valid_n["address_rw"])->matrix());
  if (!p_ibs || !p_val) stop("gpu_snp_ibs_counts: build with bigstatsr headers to map the FBMs");
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_IBS_RAW, block_size, true, device, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = p_ibs;
  res.out1 = p_val;
  run(job, res);
}

// snp_allele_sharing(): n x n matrix with NA_real_ where no locus is jointly typed
// (R/snp_allele_sharing.R:33-82).  pad_quirk = TRUE reproduces src/snp_as.cpp:56-64.
// [[Rcpp::export]]
NumericMatrix gpu_snp_allele_sharing(Environment BM, const IntegerVector& rowInd,
                                     const IntegerVector& colInd, double block_size,
                                     bool pad_quirk = true, int device = 0) {
  const int n = rowInd.size();
  NumericMatrix out(n, n);
  tpg_job job; tpg_result res; std::string keep;
  fill_job(job, BM, rowInd, colInd, TPG_STAT_AS, block_size, pad_quirk, device, keep);
  std::memset(&res, 0, sizeof res);
  res.out0 = out.begin();
  run(job, res);
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
  fill_job(job, BM, rowInd, colInd, TPG_STAT_KING, block_size, true, device, path);
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
  fill_job(job, BM, rowInd, colInd, TPG_STAT_AS, 1, true, device, path);
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
