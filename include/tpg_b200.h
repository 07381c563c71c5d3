/*
 * tpg_b200.h — C ABI of the B200-native pairwise-individual similarity path
 * (KING-robust kinship, IBS, allele sharing, GRM) of tidypopgen.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no R / Rcpp /
 * torch types.  Every entry point names the reference interface it replaces
 * (paths relative to the tidypopgen source tree).
 *
 * Reference today                                   Replacement
 * ------------------------------------------------  -----------------------------
 * .Call(_tidypopgen_increment_king_numerator) once  ONE tpg_pairwise() per job
 *   per loci block   src/RcppExports.cpp:289-306    (stat = TPG_STAT_KING)
 *   kernel           src/snp_king.cpp:21-74
 *   driver + finish  R/snp_king.R:32-103
 * .Call(_tidypopgen_increment_ibs_counts)           tpg_pairwise(TPG_STAT_IBS_*)
 *   src/RcppExports.cpp:272-288, src/snp_ibs.cpp:22-74, R/snp_ibs.R:42-104
 * .Call(_tidypopgen_increment_as_counts)            tpg_pairwise(TPG_STAT_AS)
 *   src/RcppExports.cpp:256-271, src/snp_as.cpp:22-67, R/snp_allele_sharing.R:33-82
 * pairwise_grm post-processing R/pairwise_grm.R:42-50   tpg_pairwise(TPG_STAT_GRM),
 *                                                       tpg_grm_from_as()
 * CutBySize -> bigparallelr::split_len              tpg_split_len(), tpg_as_pad_count()
 *   R/local_reimplementations.R:13-15
 * bigstatsr::block_size(n)                          tpg_block_size()
 *
 * Index conventions are R's: ind_row / ind_col are 1-based, any order, any subset
 * (R/snp_king.R:12-15); outputs are n x n column-major doubles where row/col k
 * corresponds to ind_row[k].
 *
 * There is no CPU fallback: every compute entry point needs a B200 (sm_100a) and
 * returns TPG_ERR_CUDA otherwise.
 */
#ifndef TPG_B200_H
#define TPG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPG_ABI_VERSION 2

/* ---- error codes (0 = success).  The message goes to the caller's err buffer;
 * the R glue turns a non-zero code into Rcpp::stop(msg), mirroring myassert()
 * -> Rcpp::exception -> END_RCPP in the reference (src/snp_king.cpp:11,63). */
enum {
  TPG_OK = 0,
  TPG_ERR_ARG = 1,   /* bad argument / index out of range / short .bk          */
  TPG_ERR_IO = 2,    /* cannot open / read the backing file                    */
  TPG_ERR_CUDA = 3,  /* CUDA error, no device, or not an sm_100 device         */
  TPG_ERR_NOMEM = 4, /* host or device allocation failed                       */
  TPG_ERR_KERNEL = 5 /* device-side watchdog tripped (pipeline stall)          */
};

/* ---- statistics.  One job computes one of these. */
enum {
  TPG_STAT_KING = 0,           /* snp_king: out0 = KING-robust kinship           */
  TPG_STAT_IBS_RAW = 1,        /* snp_ibs(type="raw_counts"): out0 = ibs, out1 = valid_n */
  TPG_STAT_IBS_PROPORTION = 2, /* snp_ibs(type="proportion"):  out0 = ibs/valid_n */
  TPG_STAT_IBS_ADJUSTED = 3,   /* snp_ibs(type="adjusted_counts"): out0 = prop*m  */
  TPG_STAT_AS = 4,             /* snp_allele_sharing: out0 = allele sharing       */
  TPG_STAT_GRM = 5             /* pairwise_grm: out0 = 2*(AS-mb)/(1-mb)           */
};

/* ---- where the FBM.code256 bytes live. */
enum {
  TPG_SRC_FILE = 0,  /* bk_path: the .bk backing file (R: X$backingfile)       */
  TPG_SRC_HOST = 1,  /* bk_data: host pointer (e.g. the mmap bigstatsr holds)  */
  TPG_SRC_DEVICE = 2 /* bk_data: device pointer, bytes already resident in HBM */
};

/* ---- how a multi-GPU job is partitioned ("Multi-GPU"). */
enum {
  TPG_SHARD_NONE = 0,
  TPG_SHARD_LOCI = 1, /* this call sees loci shard_rank of shard_world (contiguous slices of ind_col) */
  TPG_SHARD_TILES = 2 /* this call computes output blocks owned by shard_rank (all loci) */
};

/*
 * One pairwise job.  Mirrors the arguments of snp_king / snp_ibs /
 * snp_allele_sharing (R/snp_king.R:32-37, R/snp_ibs.R:42-48,
 * R/snp_allele_sharing.R:33-38): X -> (src_kind, bk_*, nrow_total, ncol_total),
 * ind.row -> ind_row, ind.col -> ind_col, block.size -> block_size,
 * type -> stat.
 */
typedef struct tpg_job {
  int32_t abi_version;   /* TPG_ABI_VERSION */
  int32_t stat;          /* TPG_STAT_*      */
  int32_t src_kind;      /* TPG_SRC_*       */
  int32_t device;        /* CUDA device ordinal */
  const char* bk_path;   /* TPG_SRC_FILE */
  const uint8_t* bk_data; /* TPG_SRC_HOST / TPG_SRC_DEVICE: column-major nrow_total x ncol_total, 1 byte/genotype */
  int64_t nrow_total;    /* nrow(X): individuals stored in the .bk */
  int64_t ncol_total;    /* ncol(X): loci stored in the .bk        */
  const int32_t* ind_row; /* 1-based; NULL = all rows in order     */
  int64_t n;             /* length(ind.row)                        */
  const int32_t* ind_col; /* 1-based; NULL = all columns in order  */
  int64_t m;             /* length(ind.col)                        */
  int64_t block_size;    /* R block.size; only affects allele sharing / GRM through the
                            reference's padding quirk (src/snp_as.cpp:56-62); <=0 = bigstatsr default */
  int32_t pad_quirk;     /* 1 = reproduce that quirk (reference-exact, default), 0 = mathematically clean */
  int32_t shard_mode;    /* TPG_SHARD_* */
  int32_t shard_rank;
  int32_t shard_world;
  int64_t gpu_chunk_loci; /* loci per GPU chunk (0 = auto).  Decoupled from block_size. */
  int32_t flags;         /* TPG_FLAG_* */
  int32_t reserved;
} tpg_job;

#define TPG_FLAG_PLANES_K_MAJOR 1 /* transpose in decode, K-major UMMA operands (default: MN-major) */
#define TPG_FLAG_NO_OVERLAP 2     /* serialise copy / decode / gram (debugging)                      */
#define TPG_FLAG_GRAM_1CTA 4      /* single-CTA gram kernel (default: CTA pairs, cta_group::2)       */
#define TPG_FLAG_NO_PACK 8        /* host sources: ship raw bytes, no 2-bit packing while staging     */
#define TPG_FLAG_GRAM_I8 32        /* int8 Gram kernels only (default: chunks, raw or 2-bit packed, take *
                                   * the 4-bit kind::mxf4 kernel — exact, twice the MAC rate)           */
#define TPG_FLAG_ASYNC 16         /* tpg_accumulate: enqueue and return; tpg_finish on the same ctx   *
                                   * waits (source bytes must stay valid until then).  Work enqueued  *
                                   * on tpg_ctx_stream in between (an all-reduce) is ordered after it */

/* Per-stage figures for the caller (R glue prints them when verbose). */
typedef struct tpg_stats {
  double ms_total;   /* wall time of the call                                       */
  double ms_read;    /* host: disk/mmap -> pinned staging (summed over chunks)      */
  double ms_h2d;     /* device: H2D copies  (CUDA events, summed)                   */
  double ms_decode;  /* device: decode_planes kernels (CUDA events, summed)         */
  double ms_gram;    /* device: gram_i8 kernels (CUDA events, summed)               */
  double ms_finish;  /* device: finish kernels                                      */
  double ms_d2h;     /* device: D2H of results                                      */
  int64_t bytes_h2d;
  int64_t bytes_d2h;
  int64_t n_chunks;
  int64_t n_launches; /* kernels launched by this library in the call               */
  int64_t mma_macs;   /* int8 MACs issued to the tensor pipe (incl. padding)        */
  int64_t n_short;    /* allele-sharing padding-quirk constant that was applied     */
  int64_t n_packed_chunks; /* chunks that crossed PCIe 2-bit packed (4 genotypes / byte)  */
  int64_t n_mxf4_chunks;   /* chunks whose products ran on the 4-bit kind::mxf4 kernel   */
  int64_t n_i8_chunks;     /* chunks whose products ran on the int8 kind::i8 kernels     */
  int64_t max_chain_loci;  /* longest accumulation chain (loci) any 4-bit work item ran: the FP32 *
                            * accumulator of kind::mxf4 is exact for integer sums below 2^24, so  *
                            * the library keeps this <= TPG_MXF4_MAX_CHAIN_LOCI whatever the      *
                            * chunk size or k-split override (longer chunks are split)            */
  int64_t bytes_file;      /* bytes read from the backing file                             */
  int64_t n_direct_io;     /* chunks read with O_DIRECT (page cache bypassed)              */
  int64_t n_cached_jobs;   /* 1 if the count matrices came from the context's cache       */
  double ms_reduce;        /* multi-GPU: summing the count matrices over NVLink            */
  int64_t n_geno_cached;   /* 1 if the genotypes came from the on-device 2-bit cache (tpg_config.cache_genotypes) */
  int64_t n_result_shells; /* shells the result came home in (tpg_config.result_shells), 0 = one piece            */
  int64_t n_wide_launches; /* Gram launches that ran a wide-tile kernel (256 x 480 4-bit / 256 x 512 int8, *
                            * TMA-reduction epilogue) rather than a ping-pong pair kernel                */
} tpg_stats;

/* Longest accumulation chain a 4-bit (kind::mxf4, FP32 accumulator) work item may run. */
#define TPG_MXF4_MAX_CHAIN_LOCI 8388608 /* 2^23: every partial sum stays an exact integer */

/*
 * Results.  All matrices are caller-allocated, n x n, column-major (an R REALSXP,
 * or the mmapped .bk of a bigstatsr::FBM(n, n) for raw_counts — R/snp_ibs.R:84-85).
 * out0 is required; the others may be NULL.
 *
 *   stat            out0              out1       aux0 (optional)     aux1 (optional)
 *   KING            kinship           -          KING numerator K    N_Aa_i (non-symmetric)
 *   IBS_RAW         ibs counts        valid_n    -                   -
 *   IBS_PROPORTION  ibs/valid_n       -          ibs counts          valid_n
 *   IBS_ADJUSTED    ibs/valid_n * m   -          ibs counts          valid_n
 *   AS              allele sharing    -          dos_tcross (num)    na_tcross (denom)
 *   GRM             GRM               -          allele sharing      -
 */
typedef struct tpg_result {
  double* out0;
  double* out1;
  double* aux0;
  double* aux1;
  int32_t out_on_device; /* 1 = the four pointers are device pointers (no D2H) */
  int32_t reserved;
  /* File-backed outputs: when out0 (out1) is NULL and out0_path (out1_path) is not, the library
   * maps that file read-write and writes the n x n column-major doubles into it — the .bk of a
   * bigstatsr::FBM(n, n) (type double has no header), which is what snp_ibs(type = "raw_counts")
   * returns (R/snp_ibs.R:54-56,84-85; the reference aliases the same mapping through
   * FBM_RW2arma, src/snp_ibs.cpp:8-15).  The file must exist and hold >= 8 n^2 bytes. */
  const char* out0_path;
  const char* out1_path;
  tpg_stats stats;
} tpg_result;

/* ------------------------------------------------------------------ configuration
 * One documented struct instead of environment switches.  The R glue fills it from
 * options(tidypopgen.gpu.<field>) (r/R/snp_king_gpu.R: tpg_gpu_config()); the Python mirror
 * from tidypopgen_b200.config(**fields).  Process-wide; read when a job starts.
 * tpg_config_init() gives the built-in defaults and then applies the TPG_<FIELD> environment
 * variables named below (read there and nowhere else — a debugging convenience). */
typedef struct tpg_config {
  int32_t abi_version;        /* TPG_ABI_VERSION */
  int32_t gram_route;         /* TPG_GRAM_I8: 0 = 4-bit kind::mxf4 where it is exact (default), 1 = int8 kind::i8 only */
  int32_t gram_ctas;          /* TPG_GRAM_CTAS: 0/2 = CTA pairs (default), 1 = single-CTA int8 kernel */
  int32_t gram_ksplit;        /* TPG_GRAM_KS: k-splits per tile, 0 = auto */
  int32_t gram_max_kblocks;   /* TPG_GRAM_MAXKB: k-blocks per work item in the many-tiles regime, 0 = 128 */
  int32_t gram_clusters;      /* TPG_GRAM_CLUSTERS: CTA pairs to launch, 0 = all resident */
  int32_t gram_no_trim;       /* TPG_GRAM_NO_TRIM: 1 = do not trim the last tile column to the valid individuals */
  int32_t gram_wide;          /* TPG_GRAM_WIDE: 4-bit route, 256 x 480 tiles (one A stream feeds two B tiles: a quarter  *
                               * less L2 traffic per MAC). 0 = where tiles are plentiful (default), 1 = never, 2 = always */
  int32_t gram_l2_rows;       /* TPG_GRAM_L2_ROWS / TPG_GRAM_L2_COLS: the L2 neighbourhood of the Gram work list, in 256-row and   *
                               * 240- (int8: 256-) column tiles; 0 = auto: 8 x 8 (2048 x 1920) for biobank-sized 4-bit jobs, else 4 x 4 */
  int32_t gram_l2_cols;
  int32_t gram_order;         /* TPG_GRAM_ORDER: work-list order inside a neighbourhood. 0 = auto (product-major for the int8     *
                               * kernels and for biobank-sized 4-bit jobs, tile-major for mid-sized 4-bit jobs), 1 = product-major, *
                               * 2 = tile-major (round 1's order)                                                                  */
  int32_t pack_mode;          /* TPG_PACK_MODE: 0 = hybrid (pack or raw DMA per chunk), 1 = never pack (TPG_NO_PACK), *
                               * 2 = always pack (TPG_NO_HYBRID) */
  int32_t pk2_i8;             /* TPG_PK2_I8: 1 = packed chunks take the int8 kernels */
  int32_t no_compact;         /* TPG_NO_COMPACT: 1 = gathered rows are gathered inside the decode kernel */
  int32_t gather_smem_kb;     /* TPG_GATHER_SMEM_KB: shared memory per gather_pack CTA, 0 = 52 (four CTAs per SM) */
  int32_t gather_layout;      /* TPG_GATHER_LAYOUT: 0 = column-major staging (default), 1 = transposed staging, four loci *
                               * per shared-memory word (fewer bank conflicts, but measured slower: DESIGN.md section 4)   */
  int32_t io_mode;            /* TPG_IO_MODE: file sources. 0 = auto (pread through the page cache), 1 = O_DIRECT *
                               * (cold files, bypasses the page cache), 2 = mmap (TPG_NO_PREAD) */
  int32_t no_thp;             /* TPG_NO_THP: 1 = no MADV_HUGEPAGE on result matrices */
  int32_t decode_k4_v1;       /* TPG_DECODE_K4_V1: 1 = first build of the 4-bit decode kernel (cross-check) */
  int32_t host_threads;       /* TPG_HOST_THREADS: staging worker threads, 0 = cores / local ranks (max 32) */
  int32_t scan_loci_per_cta;  /* TPG_SCAN_LOCI: 0 = auto */
  int32_t scan_no_indiv;      /* TPG_SCAN_NO_INDIV: debugging */
  int32_t scan_table;         /* TPG_SCAN_TABLE: scan kernel, 0 = auto, 1 = lookup table (any code table), 2 = table-free, *
                               * 16-locus rounds, 3 = table-free, warp per locus (auto: n >= 512)                        */
  int32_t cache_counts;       /* TPG_CACHE_COUNTS: 1 = a file-sourced job computes all four count matrices and keeps *
                               * them in the context; the next job on the same file (size, mtime), ind.row and     *
                               * ind.col finishes from them without touching the genotypes (pairwise_king, then    *
                               * pairwise_ibs, then pairwise_allele_sharing on one gen_tibble = one pass)           */
  int32_t cache_genotypes;    /* TPG_CACHE_GENOTYPES: 1 = a file-sourced job that reads the WHOLE .bk (all rows, all   *
                               * loci) leaves its 2-bit packed image in HBM (nrow x ncol / 4 bytes); later jobs on  *
                               * that file (same size, mtime) — any ind.row / ind.col, any statistic — gather from *
                               * it on the device: no disk read, no host packing, no PCIe (tpg_stats.n_geno_cached) */
  int32_t result_shells;      /* TPG_RESULT_SHELLS: whole-file jobs with >= ~27,000 individuals compute and hand back  *
                               * the result in nested shells (top-left square first, then L-shaped rims), so the D2H *
                               * of one shell overlaps the products of the next: 0 = auto (6, 4 or 3 shells), 1 = off, *
                               * 2..8 = that many shells                                                            */
  int64_t device_chunk_loci;  /* TPG_DEVICE_CHUNK_LOCI: loci per chunk of a device-resident source, 0 = 2^20 */
  int64_t host_chunk_mb;      /* TPG_HOST_CHUNK_MB: raw MB per chunk of a host / file source, 0 = 48 */
  double link_gbps;           /* TPG_LINK_GBPS: PCIe model of the hybrid schedule, 0 = 45 */
  double hybrid_factor;       /* TPG_HYBRID_FACTOR: 0 = 0.9 */
} tpg_config;

void tpg_config_init(tpg_config* cfg);                             /* defaults + TPG_* environment   */
int tpg_config_set(const tpg_config* cfg, char* err, size_t errlen); /* process-wide, validated       */
void tpg_config_get(tpg_config* cfg);                              /* the configuration in force     */

/* ------------------------------------------------------------------ host-only helpers */

const char* tpg_version(void);

/* Number of CUDA devices visible (0 on a CPU-only host; never fails). */
int tpg_device_count(void);

/* bigstatsr::block_size(n) with the default options(bigstatsr.block.sizeGB = 1):
 * max(1, floor(2^30 / (8 n))).  Used as the default block.size by every
 * wrapper (R/pairwise_king.R:36, R/snp_ibs.R:47, R/snp_allele_sharing.R:37). */
int64_t tpg_block_size(int64_t n);

/* CutBySize(m, block.size) = bigparallelr::split_len(m, nb_split = ceiling(m / block.size))
 * (R/local_reimplementations.R:13-15).  Fills lower/upper/size (1-based, inclusive)
 * for up to cap intervals; returns the number of intervals, or -1 on bad arguments. */
int64_t tpg_split_len(int64_t m, int64_t block_size, int64_t* lower, int64_t* upper,
                      int64_t* size, int64_t cap);

/* Number of loci blocks shorter than the longest one = the constant the reference
 * adds to every allele-sharing numerator (src/snp_as.cpp:56-64 with the scratch
 * sized by R/snp_allele_sharing.R:55-56). */
int64_t tpg_as_pad_count(int64_t m, int64_t block_size);

/* ------------------------------------------------------------------ one call per job */

/* Runs a whole job on job->device and writes the result matrices.
 * Returns TPG_OK or an error code with a message in err (never partial results
 * reported as success). */
int tpg_pairwise(const tpg_job* job, tpg_result* res, char* err, size_t errlen);

/* The same job spread over several GPUs of one box from ONE process (what an R session
 * needs): one host thread per device accumulates its shard — shard_mode TPG_SHARD_LOCI
 * (contiguous slices of ind_col) or TPG_SHARD_TILES (owned 256x256 output blocks, all loci);
 * TPG_SHARD_NONE picks loci while a device's count matrices stay within 64 GB — the int32 count
 * matrices are summed onto devices[0] by a pairwise tree of peer-reading add kernels over
 * NVLink and finished there.  Results are identical to the single-GPU job. */
int tpg_pairwise_multi(const tpg_job* job, tpg_result* res, const int32_t* devices, int32_t ndev,
                       int32_t shard_mode, char* err, size_t errlen);

/* Several statistics of the SAME (X, ind.row, ind.col) in one pass over the genotypes: the
 * four integer products DD, NN, HH, HN are computed once (2.5 MAC per pair-SNP, one decode)
 * and each statistic is finished from them — instead of one decode and 2.0 + 1.5 + 1.0 MAC for
 * snp_king (R/snp_king.R:32), snp_ibs (R/snp_ibs.R:42) and snp_allele_sharing
 * (R/snp_allele_sharing.R:33) run one after the other.  job->stat is ignored; stats[k] is a
 * TPG_STAT_* and results[k] its outputs (same conventions as tpg_pairwise).  devices / ndev as
 * in tpg_pairwise_multi (NULL / <= 1: job->device only).  results[0].stats carries the
 * stage times of the shared pass. */
int tpg_pairwise_many(const tpg_job* job, const int32_t* stats, int32_t nstats, tpg_result* results,
                      const int32_t* devices, int32_t ndev, int32_t shard_mode, char* err, size_t errlen);

/* pairwise_grm(x, allele_sharing_mat = <precomputed>) — R/pairwise_grm.R:42-50.
 * as_in / grm_out: n x n column-major host doubles (may alias). */
int tpg_grm_from_as(int device, const double* as_in, int64_t n, double* grm_out, double* mb_out,
                    char* err, size_t errlen);

/* ------------------------------------------------------------------ consumers of the matrices
 * (SURVEY.md section 8f): what the reference does next with an n x n result, done where the
 * matrix already is.  `matrix` / `allele_sharing` are n x n column-major doubles, on the host
 * (on_device = 0) or already in HBM (on_device = 1, e.g. the out0 of a job run with
 * out_on_device = 1). */

/* filter_high_relatedness(matrix, kings_threshold) — R/filter_high_relatedness.R:26-145.
 *   keep[k]  (n, original order) = 1 if individual k passes the filter (the reference's [[3]])
 *   order[r] (n, may be NULL)    = 0-based original index visited r-th (max_abs_cor_order - 1);
 *                                  passed_filter ([[1]]) is order[] restricted to kept individuals
 * Returns TPG_ERR_ARG with R's message "missing value where TRUE/FALSE needed" where the
 * reference stops on a NaN relatedness. */
int tpg_filter_high_relatedness(int device, const double* matrix, int32_t on_device, int64_t n,
                                double kings_threshold, int32_t* keep, int32_t* order, int64_t* n_removed,
                                char* err, size_t errlen);

/* qc_report_indiv(kings_threshold = ...) without shipping the matrix (R/qc_report_indiv.R:85-99):
 * runs the KING job of `job` (stat is forced to TPG_STAT_KING), keeps the n x n matrix in HBM,
 * filters it there and returns keep / order as tpg_filter_high_relatedness does.  king_host
 * (n x n, may be NULL) receives the kinship matrix only if the caller wants it. */
int tpg_king_filter(const tpg_job* job, double kings_threshold, int32_t* keep, int32_t* order,
                    int64_t* n_removed, double* king_host, tpg_stats* stats, char* err, size_t errlen);

/* qc_report_indiv on a GROUPED gen_tibble (R/qc_report_indiv.R:110-181): the reference runs one
 * pairwise_king and one filter_high_relatedness per group (:161, :169-180) — many small jobs.  Here the
 * individuals are sorted by group inside the library, ONE KING job computes only the tiles that hold a
 * same-group pair (the groups are the diagonal blocks), and each block is filtered where it is.
 *   group     (n) 0-based group of each individual of ind_row (job order), ngroups groups
 *   keep      (n, job order) 1 = passes its group's filter (the concatenated [[3]] of the reference)
 *   n_removed (ngroups, or NULL)
 *   king_host (sum of n_g^2 doubles, or NULL) the per-group KING matrices, group after group, each
 *             column-major n_g x n_g in the job's within-group order (attr(to_keep, "king")) */
int tpg_king_filter_groups(const tpg_job* job, const int32_t* group, int32_t ngroups, double kings_threshold,
                           int32_t* keep, int64_t* n_removed, double* king_host, tpg_stats* stats, char* err,
                           size_t errlen);

/* Weir & Goudet (2017) statistics from an allele-sharing matrix:
 *   inbreeding (n, or NULL)       indiv_inbreeding — R/indiv_inbreeding.R:56-81
 *   fst  (npop + 1, or NULL)      pop_fst, last entry = global — R/pop_fst.R:31-76
 *   fis  (npop + 1, or NULL)      pop_fis(method = "WG17"), last = global — R/pop_fis.R:137-191
 *   mb_out (2, or NULL)           [mean off-diagonal allele sharing, Mb between populations]
 * group: 0-based population of each individual (NULL = one population; fst / fis need >= 2). */
int tpg_wg17(int device, const double* allele_sharing, int32_t on_device, int64_t n, const int32_t* group,
             int64_t npop, double* inbreeding, double* fst, double* fis, double* mb_out, char* err,
             size_t errlen);

/* One pass over the genotype bytes THROUGH code256 (these callers use SubBMCode256Acc, unlike the
 * pairwise kernels): per-individual heterozygote / missing counts and per-locus alternate-allele
 * and missing counts.  Replaces gt_ind_hetero (src/gt_ind_hetero.cpp:12-42; indiv_het_obs,
 * indiv_missingness, qc_report_indiv), alt_freq_dip_pseudo_cpp (src/alt_freq_dip_pseudo_cpp.cpp:8-58;
 * loci_alt_freq) and the NA row of big_counts (R/loci_missingness.R:49-60).
 *   job      source, ind_row / ind_col, device, gpu_chunk_loci (stat and sharding are ignored)
 *   code256  the FBM's 256 decoded values; NA / <= -1 = missing; typed values must be integers
 *   ploidy   n doubles, 2 (diploid) or 1 (pseudohaploid); NULL = all 2
 *   het_n, na_n (n)   counts per individual               (each may be NULL)
 *   loc0 (m)          as_counts ? n_alt : n_alt / n_valid (NA_real_ where n_valid == 0)
 *   loc1 (m)          n_valid = sum of ploidy over typed individuals
 *   loc_na (m)        missing individuals per locus
 * Every sum is an exact integer (n_alt in half-allele units), so results equal the reference's. */
int tpg_scan(const tpg_job* job, const double* code256, const double* ploidy, int32_t as_counts,
             int32_t* het_n, int32_t* na_n, double* loc0, double* loc1, int32_t* loc_na, tpg_stats* stats,
             char* err, size_t errlen);

/* The same statistics straight from the genotypes: runs the allele-sharing job of `job` (stat is
 * forced to TPG_STAT_AS, block_size / pad_quirk honoured), keeps the n x n matrix in HBM and
 * reduces it there.  as_host (n x n, may be NULL) receives the matrix only if wanted. */
int tpg_as_wg17(const tpg_job* job, const int32_t* group, int64_t npop, double* inbreeding, double* fst,
                double* fis, double* mb_out, double* as_host, tpg_stats* stats, char* err, size_t errlen);

/* ------------------------------------------------------------------ staged interface
 * Used by the sharding layer (one process per GPU): accumulate integer counts for a
 * shard, combine them outside (NCCL all-reduce on the device buffer), then finish. */

typedef struct tpg_ctx tpg_ctx;

tpg_ctx* tpg_ctx_create(int device, char* err, size_t errlen);
void tpg_ctx_destroy(tpg_ctx* ctx);
/* A context keeps its staging / plane / result buffers between jobs (grow-only);
 * tpg_ctx_trim frees them.  tpg_release does the same for the context tpg_pairwise
 * caches per device. */
void tpg_ctx_trim(tpg_ctx* ctx);
void tpg_release(int device);
/* The cudaStream_t the Gram / finish kernels of this context are launched on (for event timing). */
void* tpg_ctx_stream(tpg_ctx* ctx);
/* tpg_pairwise on a caller-owned context (so a caller can bracket whole jobs with CUDA
 * events on tpg_ctx_stream); tpg_pairwise itself uses a context cached per device. */
int tpg_pairwise_ctx(tpg_ctx* ctx, const tpg_job* job, tpg_result* res, char* err, size_t errlen);

/* Layout of the int32 count buffer for n individuals: TPG_NUM_COUNT_MATS matrices,
 * each column-major with leading dimension ld = n_pad (X(i,j) at X[i + j*ld]);
 * the symmetric ones (DD, NN, HH) only hold entries with i <= j (plus whatever else
 * falls in the same 128x256 tile). */
#define TPG_NUM_COUNT_MATS 4
enum { TPG_MAT_DD = 0, TPG_MAT_NN = 1, TPG_MAT_HH = 2, TPG_MAT_HN = 3 };
/* Tile sharding (TPG_SHARD_TILES): the rank of shard_world that owns the 256 x 256 output block (bi <= bj)
 * of a job on n individuals.  Blocks are grouped into square regions — 3840 x 3840 (whole 256-row and
 * 240- / 480-column tiles of the 4-bit Gram kernels) assigned by weight, largest first to the least loaded
 * rank, while a rank gets at least 8 of them; else 1024 x 1024 (512, 256) dealt round-robin.  Host only. */
int32_t tpg_tile_owner(int64_t n, int64_t bi, int64_t bj, int32_t shard_world);
/* Result shells (tpg_config.result_shells): how a whole-file job on n individuals would cut its n x n result —
 * shells = 0 asks for the automatic count (6, 4 or 3, as many as the 3840-wide regions allow).  Returns the number of
 * shells K (0 = the result comes home in one piece) and writes K + 1 bounds: shell k holds the entries (i, j) with
 * bounds[k] <= max(i, j) < bounds[k + 1] (a top-left square, then L-shaped rims).  Geometry only; whether a job
 * uses them also depends on its source, flags and the free HBM.  Host only. */
int32_t tpg_result_shell_bounds(int64_t n, int32_t shells, int64_t* bounds /* >= 9 entries */);
int64_t tpg_counts_ld(int64_t n);       /* n rounded up to the tile multiple       */
int64_t tpg_counts_elems(int64_t n);    /* TPG_NUM_COUNT_MATS * ld * ld            */

/* Adds the counts of job's loci (honouring shard_mode) into d_counts (device,
 * tpg_counts_elems(n) int32, zeroed by the caller before the first call). */
int tpg_accumulate(tpg_ctx* ctx, const tpg_job* job, int32_t* d_counts, tpg_stats* stats,
                   char* err, size_t errlen);

/* Turns counts into the job's statistic.  m = total loci over ALL shards,
 * n_short = tpg_as_pad_count(m, block_size) (0 when pad_quirk is off). */
int tpg_finish(tpg_ctx* ctx, const tpg_job* job, const int32_t* d_counts, tpg_result* res,
               char* err, size_t errlen);

/* ------------------------------------------------------------------ kernel-level entry
 * points (device pointers; used by the parity tests and the profiler harness). */

/* decode_planes: raw FBM bytes -> int8 indicator planes d (= hom-alt minus hom-ref,
 * -1/0/+1), NM (non-missing, 0/1), H1 (het, 0/1).  Replaces the decode loops
 * src/snp_king.cpp:45-58, src/snp_ibs.cpp:45-56, src/snp_as.cpp:44-53 and the
 * SubBMAcc gather (src/snp_king.cpp:35).
 *   d_raw      : device, column-major bytes, column stride raw_stride
 *   d_row_idx  : device, 0-based rows (NULL = rows row0 .. row0+n-1)
 *   d_col_idx  : device, 0-based raw columns per chunk locus (NULL = col0 .. col0+lc-1)
 *   d_planes   : device, 3 planes, plane p at d_planes + p*plane_stride
 *   layout     : MN-major (default): plane[l*ld + i], ld = n_pad, lc_pad rows
 *                K-major (k_major=1): plane[i*ld + l], ld = lc_pad, n_pad rows
 * k_major bit 2 (value 4): d_raw is a pk2 chunk (tpg_pk2_pack_host; raw_stride = tpg_pk2_stride,
 * no row/column gather, MN-major only).
 * Padding rows/loci are written as zeros. */
int tpg_decode_planes(tpg_ctx* ctx, const uint8_t* d_raw, int64_t raw_bytes, int64_t raw_stride,
                      const int32_t* d_row_idx, int64_t row0, int64_t n, int64_t n_pad,
                      const int32_t* d_col_idx, int64_t col0, int64_t lc, int64_t lc_pad,
                      int8_t* d_planes, int64_t plane_stride, int32_t k_major,
                      char* err, size_t errlen);

/* gram_i8: int8 tcgen05 Gram products of the planes into the int32 count buffer
 * (atomic adds, so calls accumulate).  Replaces the dsyrk/dgemm products
 * src/snp_king.cpp:70-72, src/snp_ibs.cpp:67-72, src/snp_as.cpp:64-65.
 * prod_mask: bit TPG_MAT_* set = compute that product.
 * k_major: bit 0 = K-major planes, bit 1 = single-CTA kernel (default: CTA pairs). */
int tpg_gram(tpg_ctx* ctx, const int8_t* d_planes, int64_t plane_stride, int64_t n_pad,
             int64_t lc_pad, int32_t k_major, int32_t prod_mask, int32_t shard_rank,
             int32_t shard_world, int32_t* d_counts, float* ms_out, char* err, size_t errlen);

/* Deterministic synthetic genotypes (counter-based hash, include/tpg_synth.h; tpg_synth_fill_host
 * below is the byte-identical host twin): fills column-major nrow x ncol bytes on the device. */
int tpg_synth_fill(tpg_ctx* ctx, uint8_t* d_raw, int64_t nrow, int64_t ncol, int64_t col_offset,
                   uint64_t seed, double missing_rate, int32_t mode, char* err, size_t errlen);

/* Measures the tensor-pipe issue rate of `device` for the MMA kind of one of the two Gram routes —
 * kind 0: tcgen05.mma.cta_group::2.kind::i8 (M256 N256 K32), kind 1: kind::mxf4.block_scale (M256 N256 K64)
 * — with the operands resident in shared memory (no TMA / L2 traffic, no epilogue), on every CTA pair
 * of the chip at once, for about `seconds`.  *tops = 2 x MACs per second / 1e12: the measured roofline
 * the Gram kernels are read against (bench.py).  The operands are pseudo-random -1 / 0 / +1 like the
 * indicator planes (kind + 2: all zeros — same issue rate, far less power); short runs give the burst
 * figure, runs of a second or more the power-limited sustained one. */
int tpg_mma_peak(int device, int32_t kind, double seconds, double* tops, double* ms, char* err, size_t errlen);

/* The 2-bit chunk format of the streamer ("pk2", csrc/pack.cpp): bytes per packed column for a
 * row window of span individuals, and the packer itself (runs on the library's worker pool):
 * column c is read at src + c*src_stride (span raw bytes) and written at dst + c*stride. */
int64_t tpg_pk2_stride(int64_t span);
void tpg_pk2_pack_host(const uint8_t* src, int64_t src_stride, int64_t span, int64_t ncol, uint8_t* dst);

/* Host twin of tpg_synth_fill (byte-identical; include/tpg_synth.h holds the hash). */
void tpg_synth_fill_host(uint8_t* raw, int64_t nrow, int64_t ncol, int64_t col_offset,
                         uint64_t seed, double missing_rate, int32_t mode);

#ifdef __cplusplus
}
#endif
#endif /* TPG_B200_H */
