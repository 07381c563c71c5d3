// Internal declarations shared by the translation units of libtpg_b200.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../include/tpg_b200.h"
#include "config.h"

namespace tpg {

// ---- tile geometry of the Gram kernel ("gram_i8")
constexpr int kTileM = 128;   // A rows per tile  (TMEM lanes)
constexpr int kTileN = 256;   // B rows per tile  (TMEM columns)
constexpr int kTileK = 128;   // loci per pipeline stage (= one 128B swizzle span)
constexpr int kRowPad = 256;  // individuals are padded to this multiple
constexpr int kLociPad = 128; // loci per chunk are padded to this multiple

struct __align__(16) WorkItem {
  int32_t prod_ti;  // prod in bits [0,4), ti in bits [4,32)
  int32_t tj;
  int32_t kb0;      // first k-block (units of kTileK loci)
  int32_t kb1;      // one past the last k-block
};

// Per-device state: streams, events, reusable buffers, the driver entry point for
// tensor-map encoding.
struct Ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t s_compute = nullptr;  // gram_i8, finish (the stream tpg_ctx_stream exposes)
  cudaStream_t s_decode = nullptr;   // decode_planes of chunk c+1 runs beside gram_i8 of chunk c
  cudaStream_t s_copy = nullptr;     // H2D
  int* d_err_flag = nullptr;       // device alias of the host-mapped watchdog word
  int* h_err_flag = nullptr;       // ... and its host address (read without a copy)
  WorkItem* d_items = nullptr;     // work list of the last gram launch
  size_t d_items_cap = 0;
  std::vector<WorkItem> h_items;   // host copy (kept alive for the async upload)
  std::string items_key;           // geometry the uploaded list was built for
  void* encode_tiled = nullptr;    // cuTensorMapEncodeTiled via cudaGetDriverEntryPoint
  int64_t launches = 0;
  bool gram_attr_set = false;
  bool mxf4_attr_set = false;
  bool i8wide_attr_set = false;
  bool gather_attr_set = false;
  int pair_clusters = 0;           // CTA pairs resident at once (cudaOccupancyMaxActiveClusters)
  // Grow-only buffer cache, reused across jobs (cudaMalloc / cudaMallocHost / cudaFree are
  // slow and synchronising; a job must not pay them every call).  tpg_ctx_trim() frees it.
  struct Slot { void* p = nullptr; size_t cap = 0; };
  Slot d_raw[2], d_gat[2], d_planes[2], h_pin[2], h_out[2], d_rows, d_cols, d_counts, d_out[4], d_red, d_owner;
  std::vector<uint8_t> h_owner;    // tile-sharding owner table of owner_key = "nblocks/world" (internal.h)
  std::string owner_key;
  // result shells (api.cu, pairwise_shells): > 0 makes owner_table() build the shell table — owner = the shell of
  // max(region of the block row, region of the block column) — instead of the weighted one; bounds in regions
  int owner_shells = 0;
  std::vector<int64_t> shell_bounds;
  bool keep_events = false;        // accumulate_impl appends to the event pools instead of recycling them
  // d_planes[0] still holds the decoded planes of planes_key (packed image + rows + loci + plane form): the next
  // accumulate with the same key — the next result shell — multiplies them again without gathering / decoding
  std::string planes_key;
  // count-matrix cache (tpg_config.cache_counts): d_counts holds all four products of the job
  // described by counts_key (file identity + indices); empty = nothing cached
  std::string counts_key;
  int64_t counts_key_n = 0;
  // genotype cache (tpg_config.cache_genotypes): the whole .bk of geno_key (file identity), 2-bit packed,
  // column k at k * geno_stride (pk2 layout of pack.cpp over all nrow_total rows); ev_geno = its last write
  Slot d_geno;
  std::string geno_key;
  int64_t geno_stride = 0;
  cudaEvent_t ev_geno = nullptr;
  std::vector<cudaEvent_t> ev_timing, ev_plain;
  // accumulate enqueued but not yet collected (tpg_pairwise defers the sync to the end of finish)
  struct ChunkEv { cudaEvent_t c0, c1, d0, d1, g0, g1; };
  std::vector<ChunkEv> pend_ev;
  bool pending = false;
  size_t ev_timing_used = 0, ev_plain_used = 0;
};
// Ensures slot holds >= bytes (device memory, or pinned host memory when host = true).
int ensure_slot(Ctx::Slot* s, size_t bytes, bool host, const char* what, const struct Err& E);
void trim_ctx(Ctx* c);
cudaEvent_t pooled_event(Ctx* c, bool timing);

struct Err {
  char* buf;
  size_t len;
  int fail(int code, const char* fmt, ...) const;
};

#define TPG_CUDA_OK(call, E)                                                              \
  do {                                                                                    \
    cudaError_t _e = (call);                                                              \
    if (_e != cudaSuccess)                                                                \
      return (E).fail(TPG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                                \
  } while (0)

// ---- kernels' host launchers (each returns TPG_OK or an error code)
struct DecodeArgs {
  const uint8_t* raw;
  int64_t raw_bytes;   // bytes addressable from raw (for guarded word loads)
  int64_t raw_stride;  // bytes between consecutive raw columns
  const int32_t* row_idx;  // device, 0-based, or nullptr
  int64_t row0;
  int64_t n, n_pad;
  const int32_t* col_idx;  // device, 0-based, or nullptr
  int64_t col0;
  int64_t lc, lc_pad;
  int8_t* planes;
  int64_t plane_stride;
  int nplanes;  // 2 (d, NM) or 3 (d, NM, H1)
  bool k_major;
  bool packed = false;  // raw is a pk2 chunk (pack.cpp): raw_stride = packed bytes per column
  bool k4 = false;      // write K-major 4-bit E2M1 planes (operands of the kind::mxf4 Gram kernel)
};
int launch_decode(Ctx* ctx, const DecodeArgs& a, cudaStream_t st, const Err& E);
// Gathers the rows (and loci) of a raw chunk into a dense column-major chunk: out[l * n_pad + i], padding
// individuals = 0x03 (decode.cu).  a.planes / layout fields are ignored.
int launch_compact_rows(Ctx* ctx, const DecodeArgs& a, uint8_t* out, cudaStream_t st, const Err& E);
// The same gather through shared memory (decode.cu, "row gather through shared memory"): stages rows
// [row_lo, row_lo + span_rows) of every locus with coalesced loads, gathers by a.row_idx (offsets relative
// to row_lo) and writes the dense chunk 2-bit packed (pk2: column l at out + l * out_stride).  packed_in:
// the chunk is itself pk2.  gather_pack_fits: a column span of that many BYTES fits a CTA's shared memory.
bool gather_pack_fits(int64_t span_bytes);
int launch_gather_pack(Ctx* ctx, const DecodeArgs& a, bool packed_in, int64_t row_lo, int64_t span_rows,
                       uint8_t* out, int64_t out_stride, cudaStream_t st, const Err& E);

// One-pass scan of a raw chunk through code256 (decode.cu): device pointers; the per-locus
// arrays are already offset to the chunk's first locus.  Uses a.raw / row / col addressing only.
struct ScanArgs {
  const uint8_t* lut;
  const uint8_t* weight;
  const uint8_t* ploidy;
  int32_t* het_n;
  int32_t* na_n;
  int32_t* alt2;
  int32_t* valid;
  int32_t* na_l;
  int std_table = 0;  // 1 = the table is bigsnpr CODE_012, 2 = CODE_IMPUTE_PRED (table-free kernel), 0 = anything else
};
int launch_scan(Ctx* ctx, const DecodeArgs& a, const ScanArgs& s, cudaStream_t st, const Err& E);

struct GramArgs {
  const int8_t* planes;
  int64_t plane_stride;
  int64_t n;       // valid individuals (entries beyond are skipped in the epilogue)
  int64_t n_pad;
  int64_t lc_pad;  // loci in this chunk (padded)
  bool k_major;
  int ctas = 0;    // 1 = single-CTA kernel, 2 = CTA-pair kernel, 0 = default (pair)
  bool mxf4 = false;  // planes are K-major 4-bit E2M1 (decode k4): kind::mxf4 pair kernel, 256 x 240 tiles
  int prod_mask;
  int shard_rank, shard_world;  // tile sharding (world <= 1: everything)
  const uint8_t* owner = nullptr;  // HOST copy of the owner table (or nullptr): the work list is built on the host
  int32_t* counts;
  int64_t ld;
  int64_t* max_chain_out = nullptr;  // 4-bit route: longest accumulation chain (loci) of the launch
  int64_t* wide_out = nullptr;       // incremented when the launch ran a wide-tile kernel
  // Block-diagonal jobs (one KING per group, individuals sorted by group): host array of n non-decreasing
  // group ids; only tiles that hold at least one same-group pair are computed.  nullptr = every tile.
  const int32_t* group_sorted = nullptr;
  uint64_t group_sig = 0;            // identifies the grouping in the cached work list's key
};
int launch_gram(Ctx* ctx, const GramArgs& a, cudaStream_t st, int64_t* macs_out, const Err& E);

// Tile sharding: the 256x256 output blocks (bi <= bj, nblocks per side) are grouped into square
// REGIONS of R x R blocks and the regions of the upper triangle are dealt round-robin in row-major
// order; identical on every rank, host and device.  A rank therefore owns whole neighbourhoods — the unit
// whose operand panels the Gram scheduler keeps in L2; owning scattered single blocks made every tile fetch
// its panels from HBM (C4 on 8 GPUs: 1752 TOP/s per GPU against 2945 for one GPU on the whole matrix).
// R = 15 (3840 x 3840: a multiple of the 256-row AND of the 240- / 480-column tiles of the 4-bit kernels,
// so a region is cut into whole tiles — with R = 4 every 1024-column region ends in a 64-column tile that
// moves as many operand bytes as a full one) while there are at least 8 regions per rank; R = 4, 2, 1 for
// smaller matrices.
#ifdef __CUDACC__
__host__ __device__
#endif
inline int64_t shard_region_blocks(int64_t nblocks, int world) {
  {
    const int64_t nr = (nblocks + 14) / 15;
    if (nr * (nr + 1) / 2 >= int64_t(8) * world) return 15;
  }
  int64_t r = 4;
  while (r > 1) {
    const int64_t nr = (nblocks + r - 1) / r;
    if (nr * (nr + 1) / 2 >= int64_t(8) * world) break;
    r >>= 1;
  }
  return r;
}
// owner: optional table (one rank per region, row-major over the upper triangle of regions) — the balanced
// assignment build_owner_table() makes for the large regions; nullptr = round-robin by region index.
#ifdef __CUDACC__
__host__ __device__
#endif
inline bool block_owned(int64_t bi, int64_t bj, int64_t nblocks, int rank, int world, const uint8_t* owner = nullptr) {
  if (world <= 1) return true;
  const int64_t r = shard_region_blocks(nblocks, world);
  const int64_t ri = bi / r, rj = bj / r, nr = (nblocks + r - 1) / r;
  const int64_t idx = ri * nr - ri * (ri - 1) / 2 + (rj - ri);
  return owner ? int(owner[idx]) == rank : int(idx % world) == rank;
}
// With 3840 x 3840 regions there are few of them (105 for 50,000 individuals) and they differ in weight —
// diagonal ones are half, the last row / column is a sliver — so dealing them round-robin leaves one of
// 8 ranks with 18 % more work than the mean.  For R = 15 the regions are therefore assigned by weight:
// largest first, each to the least loaded rank (ties: lower index, lower rank) — a pure function of
// (nblocks, world), identical on every rank, in dist.py and here.  Empty for smaller regions (round-robin).
void build_owner_table(int64_t nblocks, int world, std::vector<uint8_t>* out);
// The table for (n_pad, world) on ctx's device (cached): *host / *dev = nullptr where round-robin applies.
int owner_table(struct Ctx* ctx, int64_t n_pad, int world, const uint8_t** host, const uint8_t** dev, const struct Err& E);

struct FinishArgs {
  int stat;
  const int32_t* counts;
  int64_t ld;
  int64_t n;
  int64_t m;        // total loci (adjusted_counts multiplies by it)
  int64_t n_short;  // allele-sharing pad-quirk constant
  int shard_rank, shard_world;  // TILES mode: only owned blocks are written
  const uint8_t* owner = nullptr;  // device copy of the owner table (or nullptr)
  double* out0;     // device, n x n column-major
  double* out1;
  double* aux0;
  double* aux1;
};
int launch_finish(Ctx* ctx, const FinishArgs& a, cudaStream_t st, const Err& E);
// GRM from an allele-sharing matrix already on the device (in place allowed).
int launch_grm(Ctx* ctx, const double* d_as, int64_t n, double* d_grm, double* mb_host,
               cudaStream_t st, const Err& E);

int launch_synth(Ctx* ctx, uint8_t* d_raw, int64_t nrow, int64_t ncol, int64_t col_offset,
                 uint64_t seed, double missing_rate, int mode, cudaStream_t st, const Err& E);

// ---- consumers of the pairwise matrices (relate.cu, wg17.cu)
int run_filter_high_relatedness(Ctx* ctx, const double* mat, bool on_device, int64_t n, double thr, int32_t* keep,
                                int32_t* order, int64_t* n_removed, const Err& E);
int run_wg17(Ctx* ctx, const double* as_in, bool on_device, int64_t n, const int32_t* group, int64_t npop,
             double* inbreeding, double* fst, double* fis, double* mb_out, const Err& E);

int check_err_flag(Ctx* ctx, const Err& E);
int launch_add_counts(Ctx* ctx, int32_t* a, const int32_t* b, int64_t elems, cudaStream_t st, const Err& E);
// a += b over the 256-blocks a job accumulated (matrices of prod_mask, upper blocks of the symmetric ones); b may be a peer's buffer.
int launch_add_counts_blocks(Ctx* ctx, int32_t* a, const int32_t* b, int64_t n, int prod_mask, cudaStream_t st, const Err& E);
// Zeroes what a job will accumulate into: the matrices of prod_mask, upper 256-blocks of the symmetric ones,
// owned blocks only when tile-sharded (finish.cu).
int launch_zero_counts(Ctx* ctx, int32_t* counts, int64_t n, int prod_mask, int shard_rank, int shard_world,
                       const uint8_t* d_owner, cudaStream_t st, const Err& E);

int run_mma_peak(Ctx* ctx, int kind, double seconds, double* tops, double* ms_out, const Err& E);

// ---- host helpers (split.cpp)
int64_t block_size_default(int64_t n);
int64_t split_len(int64_t m, int64_t block_size, int64_t* lower, int64_t* upper, int64_t* size,
                  int64_t cap);
int64_t as_pad_count(int64_t m, int64_t block_size);

}  // namespace tpg
