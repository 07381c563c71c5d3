/*
 * oracle.c — CPU restatement of tidypopgen's pairwise-individual similarity path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tidypopgen_b200/ may import, link or call this
 * file; it is the checker for tests/, __graft_entry__.smoke() and bench.py's CPU baseline.
 *
 * The reference cannot be executed here (no R; its C++ needs the un-vendored bigstatsr /
 * RcppArmadillo headers), so every function below restates the reference line by line and
 * cites it (paths relative to the tidypopgen source tree).  The restatement is pinned
 * against the reference's own golden vectors in tests/test_oracle_golden.py (KING binary
 * output test_king.kin0, PLINK output test_plinkIBS.mibs, the bundled 7x6 example, the 3x6
 * toys of the reference's testthat files).  Parity NOT pinned by any reference test:
 * split_len rounding / block_size() (third-party bigparallelr / bigstatsr, restated from
 * their published source) and imputed bytes 4/5/6 — see DESIGN.md.
 *
 * Deliberately literal: FP64 indicator scratch matrices of n x max(block size) columns, one
 * pass per loci block, FP64 products accumulated into n x n FP64 matrices, then the R
 * finishing formulas in R's operator order.  Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ block partition */

/* bigstatsr::block_size(n): max(1, floor(getOption("bigstatsr.block.sizeGB") * 1024^3 /
 * (8 n))) with the default option value 1 [bigstatsr R/utils.R, third-party].  Default
 * block.size of every driver: R/pairwise_king.R:36, R/snp_ibs.R:47,
 * R/snp_allele_sharing.R:37. */
int64_t orc_block_size(int64_t n) {
  double b = floor(1073741824.0 / (8.0 * (double)n));
  return b < 1.0 ? 1 : (int64_t)b;
}

/* CutBySize(m, block.size, nb = ceiling(m / block.size)) -> bigparallelr::split_len(m,
 * nb_split = nb)  (R/local_reimplementations.R:13-15).  split_len [bigparallelr
 * R/split.R, third-party]: nb_split clamped to [1, total_len]; int <- total_len /
 * nb_split; upper <- round(1:nb_split * int); lower <- c(1, upper[-nb_split] + 1);
 * size <- c(upper[1], diff(upper)).  R's round() is IEC 60559 half-to-even. */
int64_t orc_split_len(int64_t m, int64_t block_size, int64_t* lower, int64_t* upper,
                      int64_t* size, int64_t cap) {
  if (m <= 0 || block_size <= 0) return -1;
  int64_t nb = (int64_t)ceil((double)m / (double)block_size);
  if (nb > m) nb = m;
  if (nb == 0) nb = 1;
  double step = (double)m / (double)nb;
  int64_t prev = 0;
  for (int64_t k = 1; k <= nb; ++k) {
    int64_t up = (int64_t)nearbyint((double)k * step);
    if (k - 1 < cap) {
      if (lower) lower[k - 1] = prev + 1;
      if (upper) upper[k - 1] = up;
      if (size) size[k - 1] = up - prev;
    }
    prev = up;
  }
  return nb;
}

static int64_t max_block(int64_t m, int64_t block_size, int64_t** lo, int64_t** up, int64_t** sz,
                         int64_t* nb_out) {
  int64_t nb = orc_split_len(m, block_size, NULL, NULL, NULL, 0);
  *lo = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
  *up = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
  *sz = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
  orc_split_len(m, block_size, *lo, *up, *sz, nb);
  int64_t mx = 0;
  for (int64_t k = 0; k < nb; ++k)
    if ((*sz)[k] > mx) mx = (*sz)[k];
  *nb_out = nb;
  return mx;
}

/* SubBMAcc<unsigned char> macc(xpBM, rowInd, colInd, 1): macc(i, j) = the RAW byte at
 * bk[(rowInd[i]-1) + (colInd[j]-1) * nrow]  (src/snp_king.cpp:35; never code256). */
static inline int macc(const uint8_t* bk, int64_t nrow, const int32_t* row, const int32_t* col,
                       int64_t i, int64_t j) {
  return bk[(int64_t)(row[i] - 1) + (int64_t)(col[j] - 1) * nrow];
}

/* C (n x n, column-major) += alpha * A * B'  with A, B n x b column-major: the dgemm/dsyrk
 * Armadillo emits for X * Y.t(). */
static void add_abt(double* C, const double* A, const double* B, int64_t n, int64_t b,
                    double alpha) {
  for (int64_t k = 0; k < b; ++k) {
    const double* a = A + k * n;
    const double* bb = B + k * n;
    for (int64_t j = 0; j < n; ++j) {
      const double s = alpha * bb[j];
      if (s == 0.0) continue;
      double* c = C + j * n;
      for (int64_t i = 0; i < n; ++i) c[i] += a[i] * s;
    }
  }
}

/* ------------------------------------------------------------------ KING */

/* increment_king_numerator (src/snp_king.cpp:21-74) driven by snp_king (R/snp_king.R:32-103).
 * Outputs (n x n column-major, any may be NULL): king, K (numerator), N_Aa_i. */
int orc_snp_king(const uint8_t* bk, int64_t nrow, const int32_t* ind_row, int64_t n,
                 const int32_t* ind_col, int64_t m, int64_t block_size, double* king, double* K_out,
                 double* N_out) {
  int64_t *lo, *up, *sz, nb;
  const int64_t bmax = max_block(m, block_size, &lo, &up, &sz, &nb); /* R/snp_king.R:51,57-61 */
  const size_t nn = (size_t)n * (size_t)n, nbm = (size_t)n * (size_t)bmax;
  double* K = (double*)calloc(nn, sizeof(double));  /* k_numerator FBM(n, n, init = 0) :45 */
  double* N = (double*)calloc(nn, sizeof(double));  /* N_Aa_i :48 */
  double* g0 = (double*)malloc(nbm * sizeof(double));
  double* g1 = (double*)malloc(nbm * sizeof(double));
  double* g2 = (double*)malloc(nbm * sizeof(double));
  double* gv = (double*)malloc(nbm * sizeof(double));
  for (int64_t blk = 0; blk < nb; ++blk) { /* R/snp_king.R:63-77 */
    const int32_t* cols = ind_col + (lo[blk] - 1);
    const int64_t b = sz[blk];
    memset(g0, 0, nbm * sizeof(double)); /* src/snp_king.cpp:37-40 */
    memset(g1, 0, nbm * sizeof(double));
    memset(g2, 0, nbm * sizeof(double));
    memset(gv, 0, nbm * sizeof(double));
    for (int64_t j = 0; j < b; ++j) /* :45-58 */
      for (int64_t i = 0; i < n; ++i) {
        const int value = macc(bk, nrow, ind_row, cols, i, j);
        if (value == 0) { g0[i + j * n] = 1; gv[i + j * n] += 1; }
        else if (value == 1) { g1[i + j * n] = 1; gv[i + j * n] += 1; }
        else if (value == 2) { g2[i + j * n] = 1; gv[i + j * n] += 1; }
      }
    /* :61-68 pad column (at most one) is zero already */
    /* :70  K += g1 g1' - 2 (g0 g2' + g2 g0') */
    add_abt(K, g1, g1, n, bmax, 1.0);
    add_abt(K, g0, g2, n, bmax, -2.0);
    add_abt(K, g2, g0, n, bmax, -2.0);
    /* :72  N_Aa_i += g1 valid' */
    add_abt(N, g1, gv, n, bmax, 1.0);
  }
  if (king) { /* divide_king_sub, R/snp_king.R:82-91; N_Aa_j = t(N_Aa_i) :79 */
    for (int64_t j = 0; j < n; ++j)
      for (int64_t i = 0; i < n; ++i) {
        const double Ni = N[i + j * n], Nj = N[j + i * n];
        const double pm = Ni < Nj ? Ni : Nj; /* pmin */
        const double t1 = K[i + j * n] / (2 * pm);
        const double t2 = t1 + 0.5;
        const double t3 = 0.25 * (Ni + Nj);
        king[i + j * n] = t2 - t3 / pm;
      }
  }
  if (K_out) memcpy(K_out, K, nn * sizeof(double));
  if (N_out) memcpy(N_out, N, nn * sizeof(double));
  free(K); free(N); free(g0); free(g1); free(g2); free(gv); free(lo); free(up); free(sz);
  return 0;
}

/* ------------------------------------------------------------------ IBS */

/* increment_ibs_counts (src/snp_ibs.cpp:22-74) driven by snp_ibs (R/snp_ibs.R:42-104).
 * type: 0 = proportion, 1 = adjusted_counts, 2 = raw_counts.  out = the returned matrix
 * (ibs counts for raw_counts); ibs_out / valid_out (optional) always get the raw counts. */
int orc_snp_ibs(const uint8_t* bk, int64_t nrow, const int32_t* ind_row, int64_t n,
                const int32_t* ind_col, int64_t m, int64_t block_size, int type, double* out,
                double* ibs_out, double* valid_out) {
  int64_t *lo, *up, *sz, nb;
  const int64_t bmax = max_block(m, block_size, &lo, &up, &sz, &nb);
  const size_t nn = (size_t)n * (size_t)n, nbm = (size_t)n * (size_t)bmax;
  double* K = (double*)calloc(nn, sizeof(double));  /* IBS :54 */
  double* K2 = (double*)calloc(nn, sizeof(double)); /* IBS_valid_loci :56 */
  double* g0 = (double*)malloc(nbm * sizeof(double));
  double* g1 = (double*)malloc(nbm * sizeof(double));
  double* g2 = (double*)malloc(nbm * sizeof(double));
  double* g02 = (double*)malloc(nbm * sizeof(double));
  for (int64_t blk = 0; blk < nb; ++blk) {
    const int32_t* cols = ind_col + (lo[blk] - 1);
    const int64_t b = sz[blk];
    memset(g0, 0, nbm * sizeof(double)); /* src/snp_ibs.cpp:38-40 */
    memset(g1, 0, nbm * sizeof(double));
    memset(g2, 0, nbm * sizeof(double));
    for (int64_t j = 0; j < b; ++j) /* :45-56 */
      for (int64_t i = 0; i < n; ++i) {
        const int value = macc(bk, nrow, ind_row, cols, i, j);
        if (value == 0) g0[i + j * n] = 1;
        else if (value == 1) g1[i + j * n] = 1;
        else if (value == 2) g2[i + j * n] = 1;
      }
    /* :67-68  K += 2 (g2 g2' + g1 g1' + g0 g0') + g1 (g0+g2)' + (g0+g2) g1' */
    for (size_t e = 0; e < nbm; ++e) g02[e] = g0[e] + g2[e];
    add_abt(K, g2, g2, n, bmax, 2.0);
    add_abt(K, g1, g1, n, bmax, 2.0);
    add_abt(K, g0, g0, n, bmax, 2.0);
    add_abt(K, g1, g02, n, bmax, 1.0);
    add_abt(K, g02, g1, n, bmax, 1.0);
    /* :71-72  g0 <- g0 + g1 + g2;  K2 += 2 g0 g0' */
    for (size_t e = 0; e < nbm; ++e) g0[e] = g0[e] + g1[e] + g2[e];
    add_abt(K2, g0, g0, n, bmax, 2.0);
  }
  if (out) {
    for (size_t e = 0; e < nn; ++e) {
      if (type == 2) out[e] = K[e];                       /* R/snp_ibs.R:84-85 */
      else {
        const double prop = K[e] / K2[e];                 /* :89 */
        out[e] = type == 1 ? prop * (double)m : prop;     /* :100 */
      }
    }
  }
  if (ibs_out) memcpy(ibs_out, K, nn * sizeof(double));
  if (valid_out) memcpy(valid_out, K2, nn * sizeof(double));
  free(K); free(K2); free(g0); free(g1); free(g2); free(g02); free(lo); free(up); free(sz);
  return 0;
}

/* ------------------------------------------------------------------ allele sharing */

static double na_real(void) { /* R's NA_real_: NaN with low word 1954 */
  union { double d; uint64_t u; } x;
  x.u = 0x7FF00000000007A2ull;
  return x.d;
}

/* increment_as_counts (src/snp_as.cpp:22-67) driven by snp_allele_sharing
 * (R/snp_allele_sharing.R:33-82), INCLUDING the padding quirk: the scratch has
 * max(block size) columns (R :55-56); for a shorter block the spare column is set to
 * dos = 1, na = 0 (src/snp_as.cpp:56-62), so K gains +1 everywhere for that block. */
int orc_snp_as(const uint8_t* bk, int64_t nrow, const int32_t* ind_row, int64_t n,
               const int32_t* ind_col, int64_t m, int64_t block_size, double* as_out, double* num_out,
               double* den_out) {
  int64_t *lo, *up, *sz, nb;
  const int64_t bmax = max_block(m, block_size, &lo, &up, &sz, &nb);
  const size_t nn = (size_t)n * (size_t)n, nbm = (size_t)n * (size_t)bmax;
  double* K = (double*)calloc(nn, sizeof(double));  /* dos_tcross :43 */
  double* K2 = (double*)calloc(nn, sizeof(double)); /* na_tcross :46 */
  double* dos = (double*)malloc(nbm * sizeof(double));
  double* na = (double*)malloc(nbm * sizeof(double));
  for (int64_t blk = 0; blk < nb; ++blk) {
    const int32_t* cols = ind_col + (lo[blk] - 1);
    const int64_t b = sz[blk];
    for (size_t e = 0; e < nbm; ++e) { dos[e] = 0; na[e] = 1; } /* src/snp_as.cpp:36-38 */
    for (int64_t j = 0; j < b; ++j) /* :44-53 */
      for (int64_t i = 0; i < n; ++i) {
        const int value = macc(bk, nrow, ind_row, cols, i, j);
        if (value < 3) dos[i + j * n] = value - 1;
        else { dos[i + j * n] = 0; na[i + j * n] = 0; }
      }
    if (bmax > b) { /* :56-62 (myassert: exactly one spare column) */
      if (bmax != b + 1) return 1;
      for (int64_t i = 0; i < n; ++i) { dos[i + b * n] = 1; na[i + b * n] = 0; }
    }
    add_abt(K, dos, dos, n, bmax, 1.0); /* :64 */
    add_abt(K2, na, na, n, bmax, 1.0);  /* :65 */
  }
  if (as_out) { /* R/snp_allele_sharing.R:77-80 */
    for (size_t e = 0; e < nn; ++e) {
      const double q = K[e] / K2[e];
      const double r = 0.5 * (1 + q);
      as_out[e] = K2[e] == 0 ? na_real() : r;
    }
  }
  if (num_out) memcpy(num_out, K, nn * sizeof(double));
  if (den_out) memcpy(den_out, K2, nn * sizeof(double));
  free(K); free(K2); free(dos); free(na); free(lo); free(up); free(sz);
  return 0;
}

/* ------------------------------------------------------------------ GRM */

/* pairwise_grm (R/pairwise_grm.R:42-50): mii <- diag(AS); diag(AS) <- NA;
 * mb <- mean(AS, na.rm = TRUE); diag(AS) <- mii; ((AS - mb) / (1 - mb)) * 2.
 * R's mean() of doubles: long double sum / n, then one refinement pass
 * (R src/main/summary.c, real_mean). */
int orc_grm(const double* as, int64_t n, double* grm, double* mb_out) {
  long double s = 0.0L;
  int64_t cnt = 0;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const double x = as[i + j * n];
      if (i == j || isnan(x)) continue;
      s += x; ++cnt;
    }
  long double mean = s / (long double)cnt;
  long double t = 0.0L;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const double x = as[i + j * n];
      if (i == j || isnan(x)) continue;
      t += (x - mean);
    }
  mean += t / (long double)cnt;
  const double mb = (double)mean;
  if (mb_out) *mb_out = mb;
  for (int64_t e = 0; e < n * n; ++e) {
    const double coa = (as[e] - mb) / (1 - mb);
    grm[e] = coa * 2;
  }
  return 0;
}

/* ================================================================== consumers of the path
 * (SURVEY.md §8f): what the reference does with the pairwise matrices, restated literally. */

/* R's mean(x) for doubles (src/main/summary.c, real_mean): long double sum / n, then one
 * refinement pass when the first estimate is finite.  mean(numeric(0)) is NaN. */
static double r_mean(const double* x, int64_t n) {
  if (n == 0) return NAN;
  long double s = 0.0L;
  for (int64_t k = 0; k < n; ++k) s += x[k];
  s /= (long double)n;
  if (isfinite((double)s)) {
    long double t = 0.0L;
    for (int64_t k = 0; k < n; ++k) t += (x[k] - s);
    s += t / (long double)n;
  }
  return (double)s;
}

/* filter_high_relatedness (R/filter_high_relatedness.R:26-145), the caret::findCorrelation
 * style greedy filter.  mat: n x n column-major.  keep[k] (original order) = 1 if individual k
 * passes; order[r] = 0-based original index at rank r (max_abs_cor_order - 1); deleted_rank[r]
 * = col_to_delete in the re-ordered space.  Returns 0, or 1 where R would stop with
 * "missing value where TRUE/FALSE needed" (a NaN relatedness reached by the comparison
 * at :103). */
int orc_filter_high_relatedness(const double* mat, int64_t n, double thr, int32_t* keep,
                                int32_t* order, int32_t* deleted_rank) {
  if (n == 1) { keep[0] = 1; order[0] = 0; deleted_rank[0] = 0; return 0; }   /* :46-53 */
  const size_t nn = (size_t)n * (size_t)n;
  double* A = (double*)malloc(nn * sizeof(double));
  double* M = (double*)malloc(nn * sizeof(double));
  double* M2 = (double*)malloc(nn * sizeof(double));
  double* buf = (double*)malloc(nn * sizeof(double));
  double* cm = (double*)malloc((size_t)n * sizeof(double));
  for (size_t e = 0; e < nn; ++e) A[e] = fabs(mat[e]);                        /* :56 */
  /* :67-72 column means with the diagonal NA, na.rm = TRUE */
  for (int64_t j = 0; j < n; ++j) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i)
      if (i != j && !isnan(A[i + j * n])) buf[c++] = A[i + j * n];
    cm[j] = r_mean(buf, c);
  }
  /* order(cm, decreasing = TRUE): stable (radix), NaN last */
  for (int64_t r = 0; r < n; ++r) order[r] = (int32_t)r;
  for (int64_t a = 1; a < n; ++a) {  /* stable insertion sort */
    const int32_t v = order[a];
    int64_t b = a - 1;
    while (b >= 0) {
      const double x = cm[order[b]], y = cm[v];
      const int before = isnan(x) ? !isnan(y) : (!isnan(y) && y > x);  /* v strictly precedes order[b] */
      if (!before) break;
      order[b + 1] = order[b];
      --b;
    }
    order[b + 1] = v;
  }
  for (int64_t j = 0; j < n; ++j)                                             /* :75 */
    for (int64_t i = 0; i < n; ++i) M[i + j * n] = A[order[i] + (int64_t)order[j] * n];
  memcpy(M2, M, nn * sizeof(double));                                         /* :83-84 */
  for (int64_t i = 0; i < n; ++i) M2[i + i * n] = NAN;
  memset(deleted_rank, 0, (size_t)n * sizeof(int32_t));
  int status = 0;
  for (int64_t i = 0; i < n - 1 && !status; ++i) {                            /* :89 */
    int any = 0;
    for (size_t e = 0; e < nn && !any; ++e) any = !isnan(M2[e]) && M2[e] > thr;
    if (!any) break;                                                          /* :90-95 */
    if (deleted_rank[i]) continue;                                            /* :96-98 */
    for (int64_t j = i + 1; j < n; ++j) {
      if (deleted_rank[i] || deleted_rank[j]) continue;                       /* :100 */
      const double v = M[i + j * n];
      if (isnan(v)) { status = 1; break; }                                    /* if (NA) -> error */
      if (!(v > thr)) continue;                                               /* :101 */
      int64_t c = 0;
      for (int64_t b = 0; b < n; ++b)
        if (!isnan(M2[i + b * n])) buf[c++] = M2[i + b * n];
      const double mn1 = r_mean(buf, c);                                      /* :102 */
      c = 0;
      for (int64_t b = 0; b < n; ++b)                                         /* matrix2[-j, ] in column-major order */
        for (int64_t a = 0; a < n; ++a)
          if (a != j && !isnan(M2[a + b * n])) buf[c++] = M2[a + b * n];
      const double mn2 = r_mean(buf, c);                                      /* :103 */
      const int64_t k = (mn1 > mn2) ? i : j;                                  /* :117-131 */
      deleted_rank[k] = 1;
      for (int64_t b = 0; b < n; ++b) { M2[k + b * n] = NAN; M2[b + k * n] = NAN; }
    }
  }
  for (int64_t r = 0; r < n; ++r) keep[order[r]] = !deleted_rank[r];          /* :138-142 */
  free(A); free(M); free(M2); free(buf); free(cm);
  return status;
}

/* indiv_inbreeding (R/indiv_inbreeding.R:56-81): mii <- diag(AS); diag(AS) <- NA;
 * mb <- mean(AS, na.rm = TRUE); ((mii * 2 - 1) - mb) / (1 - mb). */
int orc_indiv_inbreeding(const double* as, int64_t n, double* out, double* mb_out) {
  double* buf = (double*)malloc((size_t)n * (size_t)n * sizeof(double));
  int64_t c = 0;
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i)
      if (i != j && !isnan(as[i + j * n])) buf[c++] = as[i + j * n];
  const double mb = r_mean(buf, c);
  if (mb_out) *mb_out = mb;
  for (int64_t i = 0; i < n; ++i) out[i] = ((as[i + i * n] * 2 - 1) - mb) / (1 - mb);
  free(buf);
  return 0;
}

/* Shared by pop_fst (R/pop_fst.R:31-76) and pop_fis_wg17 (R/pop_fis.R:137-191):
 * group = 0-based population of each individual (levels in increasing order, as
 * factor(group_indices(.x)) gives them), npop populations.
 * Fsts[p] = mean(Mij[pop p, pop p]) with diag NA; mMij[i,j] = mean(Mij[pop i, pop j]);
 * Mb = sum_{j<i} mMij[i,j] * 2 / (npop (npop - 1)). */
static void wg17_means(const double* as, int64_t n, const int32_t* group, int64_t npop,
                       double* Fsts, double* Mb_out) {
  double* buf = (double*)malloc((size_t)n * (size_t)n * sizeof(double));
  for (int64_t p = 0; p < npop; ++p) {
    int64_t c = 0;
    for (int64_t j = 0; j < n; ++j) {           /* Mij[wil, wil] in column-major order */
      if (group[j] != p) continue;
      for (int64_t i = 0; i < n; ++i)
        if (group[i] == p && i != j && !isnan(as[i + j * n])) buf[c++] = as[i + j * n];
    }
    Fsts[p] = r_mean(buf, c);
  }
  double Mb = 0;
  for (int64_t pi = 1; pi < npop; ++pi)
    for (int64_t pj = 0; pj < pi; ++pj) {
      int64_t c = 0;
      for (int64_t j = 0; j < n; ++j) {         /* Mij[p1, p2]: rows pop pi, columns pop pj */
        if (group[j] != pj) continue;
        for (int64_t i = 0; i < n; ++i)
          if (group[i] == pi && !isnan(as[i + j * n])) buf[c++] = as[i + j * n];
      }
      Mb = Mb + r_mean(buf, c);
    }
  *Mb_out = Mb * 2 / (double)(npop * (npop - 1));
  free(buf);
}

/* pop_fst: fst[p] = (Fsts[p] - Mb) / (1 - Mb); fst[npop] = mean(fst, na.rm = TRUE) (global). */
int orc_pop_fst(const double* as, int64_t n, const int32_t* group, int64_t npop, double* fst) {
  double Mb;
  wg17_means(as, n, group, npop, fst, &Mb);
  double* buf = (double*)malloc((size_t)npop * sizeof(double));
  int64_t c = 0;
  for (int64_t p = 0; p < npop; ++p) {
    fst[p] = (fst[p] - Mb) / (1 - Mb);
    if (!isnan(fst[p])) buf[c++] = fst[p];
  }
  fst[npop] = r_mean(buf, c);
  free(buf);
  return 0;
}

/* pop_fis_wg17: Mii = diag(AS) * 2 - 1; Fi = (Mii[pop] - Fsts[pop]) / (1 - Fsts[pop]);
 * fis[p] = mean(Fi, na.rm = TRUE); fis[npop] = mean(fis, na.rm = TRUE) (global). */
int orc_pop_fis_wg17(const double* as, int64_t n, const int32_t* group, int64_t npop, double* fis) {
  double Mb;
  double* Fsts = (double*)malloc((size_t)npop * sizeof(double));
  wg17_means(as, n, group, npop, Fsts, &Mb);
  double* buf = (double*)malloc((size_t)(n > npop ? n : npop) * sizeof(double));
  for (int64_t p = 0; p < npop; ++p) {
    int64_t c = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (group[i] != p) continue;
      const double mii = as[i + i * n] * 2 - 1;
      const double fi = (mii - Fsts[p]) / (1 - Fsts[p]);
      if (!isnan(fi)) buf[c++] = fi;
    }
    fis[p] = r_mean(buf, c);
  }
  int64_t c = 0;
  for (int64_t p = 0; p < npop; ++p)
    if (!isnan(fis[p])) buf[c++] = fis[p];
  fis[npop] = r_mean(buf, c);
  free(Fsts); free(buf);
  return 0;
}

/* gt_ind_hetero (src/gt_ind_hetero.cpp:12-42) through SubBMCode256Acc: x = code256[byte];
 * x > -1 ? (x == 1 -> het) : missing.  out: het_n[n], na_n[n]. */
int orc_gt_ind_hetero(const uint8_t* bk, int64_t nrow, const int32_t* ind_row, int64_t n,
                      const int32_t* ind_col, int64_t m, const double* code256,
                      int32_t* het_n, int32_t* na_n) {
  memset(het_n, 0, (size_t)n * sizeof(int32_t));
  memset(na_n, 0, (size_t)n * sizeof(int32_t));
  for (int64_t j = 0; j < m; ++j)
    for (int64_t i = 0; i < n; ++i) {
      const double x = code256[macc(bk, nrow, ind_row, ind_col, i, j)];
      if (x > -1) { if (x == 1) het_n[i] += 1; }
      else na_n[i] += 1;
    }
  return 0;
}

/* alt_freq_dip_pseudo_cpp (src/alt_freq_dip_pseudo_cpp.cpp:8-58): per locus
 * n_alt = sum x / (3 - ploidy_i), n_valid = sum ploidy_i over non-missing; freq = n_alt / n_valid
 * (NA_real_ when n_valid == 0).  out0[m] = n_alt or freq, out1[m] = n_valid. */
int orc_alt_freq(const uint8_t* bk, int64_t nrow, const int32_t* ind_row, int64_t n,
                 const int32_t* ind_col, int64_t m, const double* code256, const double* ploidy,
                 int as_counts, double* out0, double* out1) {
  double* mult = (double*)malloc((size_t)n * sizeof(double));
  for (int64_t i = 0; i < n; ++i) mult[i] = 1 / (3 - ploidy[i]);
  for (int64_t j = 0; j < m; ++j) {
    double a = 0, v = 0;
    for (int64_t i = 0; i < n; ++i) {
      const double x = code256[macc(bk, nrow, ind_row, ind_col, i, j)];
      if (x > -1) { a += x * mult[i]; v += ploidy[i]; }
    }
    out0[j] = a; out1[j] = v;
  }
  if (!as_counts)
    for (int64_t j = 0; j < m; ++j) out0[j] = out1[j] > 0 ? out0[j] / out1[j] : na_real();
  free(mult);
  return 0;
}
