// C-ABI of libtpg_b200.so (include/tpg_b200.h): job validation, the .bk chunk streamer
// (disk/mmap -> pinned staging -> HBM, double buffered against decode + gram), the staged
// accumulate/finish interface used by the sharding layer, and kernel-level entry points.
//
// Replaces the per-block R loop + .Call of the reference (R/snp_king.R:63-77,
// R/snp_ibs.R:69-83, R/snp_allele_sharing.R:58-70) by ONE call per job; the GPU chunk
// size is decoupled from R's block.size (which only survives as the allele-sharing
// padding constant, see tpg_as_pad_count).
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <mutex>
#include <thread>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "internal.h"
#include "hostpool.h"
#include "config.h"
#include "../../include/tpg_synth.h"

namespace tpg {

int Err::fail(int code, const char* fmt, ...) const {
  if (buf && len) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, len, fmt, ap);
    va_end(ap);
  }
  return code;
}

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ host helpers

int64_t block_size_default(int64_t n) {
  if (n <= 0) return 1;
  const double b = std::floor(1073741824.0 / (8.0 * double(n)));
  return b < 1.0 ? 1 : int64_t(b);
}

// bigparallelr::split_len(total_len = m, nb_split = ceiling(m / block.size)):
//   nb clamped to [1, m]; int = m / nb (double); upper = round(1:nb * int) with R's
//   half-to-even rounding; lower = c(1, upper[-nb] + 1); size = c(upper[1], diff(upper)).
int64_t split_len(int64_t m, int64_t block_size, int64_t* lower, int64_t* upper, int64_t* size,
                  int64_t cap) {
  if (m <= 0 || block_size <= 0) return -1;
  int64_t nb = int64_t(std::ceil(double(m) / double(block_size)));
  if (nb > m) nb = m;
  if (nb <= 0) nb = 1;
  const double step = double(m) / double(nb);
  int64_t prev = 0;
  for (int64_t k = 1; k <= nb; ++k) {
    const int64_t up = int64_t(std::nearbyint(double(k) * step));  // FE_TONEAREST = half-even
    if (k - 1 < cap) {
      if (lower) lower[k - 1] = prev + 1;
      if (upper) upper[k - 1] = up;
      if (size) size[k - 1] = up - prev;
    }
    prev = up;
  }
  return nb;
}

int64_t as_pad_count(int64_t m, int64_t block_size) {
  if (m <= 0 || block_size <= 0) return 0;
  int64_t nb = int64_t(std::ceil(double(m) / double(block_size)));
  if (nb > m) nb = m;
  if (nb <= 0) nb = 1;
  const double step = double(m) / double(nb);
  int64_t prev = 0, mx = 0;
  for (int64_t k = 1; k <= nb; ++k) {
    const int64_t up = int64_t(std::nearbyint(double(k) * step));
    mx = std::max(mx, up - prev);
    prev = up;
  }
  prev = 0;
  int64_t n_short = 0;
  for (int64_t k = 1; k <= nb; ++k) {
    const int64_t up = int64_t(std::nearbyint(double(k) * step));
    if (up - prev < mx) ++n_short;
    prev = up;
  }
  return n_short;
}

static int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }

// ------------------------------------------------------------------ context

static int ctx_init(Ctx* c, int device, const Err& E) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return E.fail(TPG_ERR_CUDA, "no CUDA device available (%s); this library has no CPU path",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return E.fail(TPG_ERR_ARG, "device %d out of range [0,%d)", device, ndev);
  TPG_CUDA_OK(cudaSetDevice(device), E);
  cudaDeviceProp prop;
  TPG_CUDA_OK(cudaGetDeviceProperties(&prop, device), E);
  if (prop.major != 10)
    return E.fail(TPG_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                  device, prop.major, prop.minor);
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  TPG_CUDA_OK(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking), E);
  TPG_CUDA_OK(cudaStreamCreateWithFlags(&c->s_decode, cudaStreamNonBlocking), E);
  TPG_CUDA_OK(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking), E);
  TPG_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&c->h_err_flag), sizeof(int), cudaHostAllocMapped), E);
  *c->h_err_flag = 0;
  TPG_CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_err_flag), c->h_err_flag, 0), E);
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
    return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  c->encode_tiled = fn;
  return TPG_OK;
}

int check_err_flag(Ctx* ctx, const Err& E) {
  const int flag = *reinterpret_cast<volatile int*>(ctx->h_err_flag);
  if (flag) return E.fail(TPG_ERR_KERNEL, "gram_i8 pipeline watchdog tripped (role %d)", flag);
  return TPG_OK;
}

int ensure_slot(Ctx::Slot* s, size_t bytes, bool host, const char* what, const Err& E) {
  if (s->cap >= bytes && s->p) return TPG_OK;
  if (s->p) { if (host) cudaFreeHost(s->p); else cudaFree(s->p); }
  s->p = nullptr;
  s->cap = 0;
  const size_t want = bytes + bytes / 16 + 256;  // a little headroom against near-miss regrowth
  cudaError_t e = host ? cudaMallocHost(&s->p, want) : cudaMalloc(&s->p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    s->p = nullptr;
    return E.fail(TPG_ERR_NOMEM, "cannot allocate %zu bytes of %s (%s)", want, what, cudaGetErrorString(e));
  }
  s->cap = want;
  return TPG_OK;
}

void build_owner_table(int64_t nblocks, int world, std::vector<uint8_t>* out) {
  out->clear();
  if (world <= 1 || world > 255) return;
  const int64_t r = shard_region_blocks(nblocks, world);
  if (r != 15) return;  // small regions are plentiful: round-robin balances them
  const int64_t nr = (nblocks + r - 1) / r, total = nr * (nr + 1) / 2;
  struct Reg { int64_t w, idx; };
  std::vector<Reg> regs;
  regs.reserve(size_t(total));
  for (int64_t ri = 0; ri < nr; ++ri)
    for (int64_t rj = ri; rj < nr; ++rj) {
      const int64_t h = std::min(r, nblocks - ri * r), w = std::min(r, nblocks - rj * r);
      regs.push_back({ri == rj ? h * w : 2 * h * w, ri * nr - ri * (ri - 1) / 2 + (rj - ri)});
    }
  std::stable_sort(regs.begin(), regs.end(), [](const Reg& a, const Reg& b) { return a.w != b.w ? a.w > b.w : a.idx < b.idx; });
  std::vector<int64_t> load(size_t(world), 0);
  out->assign(size_t(total), 0);
  for (const Reg& g : regs) {
    int best = 0;
    for (int k = 1; k < world; ++k) if (load[k] < load[best]) best = k;
    load[best] += g.w;
    (*out)[size_t(g.idx)] = uint8_t(best);
  }
}

// Result shells (pairwise_shells): the regions are cut into `shells` nested groups — the first one a top-left square
// holding `first_frac` of the area, the others L-shaped rims of equal area — and a block belongs to the shell of
// max(region of its row, region of its column).  Same table layout as build_owner_table.
constexpr double kShellFirstFrac = 0.45;   // shell 0's share of the area: about what the Gram kernels do while the host reads and packs the file
static void build_shell_table(int64_t nblocks, int shells, double first_frac, std::vector<uint8_t>* out, std::vector<int64_t>* bounds) {
  out->clear();
  bounds->clear();
  const int64_t r = 15, nr = (nblocks + r - 1) / r;
  if (shells < 2 || shells > nr) return;
  bounds->push_back(0);
  const double a0 = first_frac * double(nr) * double(nr);
  for (int k = 1; k < shells; ++k) {
    const double area = a0 + (double(nr) * double(nr) - a0) * double(k - 1) / double(shells - 1);
    int64_t b = int64_t(std::llround(std::sqrt(area)));
    b = std::max<int64_t>(bounds->back() + 1, std::min<int64_t>(b, nr - (shells - k)));
    bounds->push_back(b);
  }
  bounds->push_back(nr);
  out->assign(size_t(nr * (nr + 1) / 2), 0);
  for (int64_t ri = 0; ri < nr; ++ri)
    for (int64_t rj = ri; rj < nr; ++rj) {
      int k = 0;
      while (rj >= (*bounds)[size_t(k) + 1]) ++k;
      (*out)[size_t(ri * nr - ri * (ri - 1) / 2 + (rj - ri))] = uint8_t(k);
    }
}

int owner_table(Ctx* ctx, int64_t n_pad, int world, const uint8_t** host, const uint8_t** dev, const Err& E) {
  *host = nullptr;
  *dev = nullptr;
  if (world <= 1) return TPG_OK;
  const int64_t nblocks = n_pad / 256;
  char key[64];
  snprintf(key, sizeof key, "%s%lld/%d", ctx->owner_shells > 0 ? "shells:" : "", (long long)nblocks, world);
  if (ctx->owner_key != key) {
    if (ctx->owner_shells > 0) build_shell_table(nblocks, world, kShellFirstFrac, &ctx->h_owner, &ctx->shell_bounds);
    else build_owner_table(nblocks, world, &ctx->h_owner);
    if (!ctx->h_owner.empty()) {
      int rc = ensure_slot(&ctx->d_owner, ctx->h_owner.size(), false, "owner table", E);
      if (rc) return rc;
      // (synchronous: the table of a previous geometry may still be read by kernels in flight)
      TPG_CUDA_OK(cudaDeviceSynchronize(), E);
      TPG_CUDA_OK(cudaMemcpy(ctx->d_owner.p, ctx->h_owner.data(), ctx->h_owner.size(), cudaMemcpyHostToDevice), E);
    }
    ctx->owner_key = key;
  }
  if (!ctx->h_owner.empty()) {
    *host = ctx->h_owner.data();
    *dev = static_cast<const uint8_t*>(ctx->d_owner.p);
  }
  return TPG_OK;
}

void trim_ctx(Ctx* c) {
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  c->counts_key.clear();
  c->geno_key.clear();
  c->planes_key.clear();
  c->owner_key.clear();
  c->h_owner.clear();
  Ctx::Slot* dev[] = {&c->d_geno, &c->d_owner, &c->d_raw[0], &c->d_raw[1], &c->d_gat[0], &c->d_gat[1], &c->d_planes[0], &c->d_planes[1], &c->d_rows,
                      &c->d_cols, &c->d_counts, &c->d_out[0], &c->d_out[1], &c->d_out[2],
                      &c->d_out[3], &c->d_red};
  for (Ctx::Slot* s : dev) { if (s->p) cudaFree(s->p); s->p = nullptr; s->cap = 0; }
  for (Ctx::Slot& s : c->h_pin) { if (s.p) cudaFreeHost(s.p); s.p = nullptr; s.cap = 0; }
  for (Ctx::Slot& s : c->h_out) { if (s.p) cudaFreeHost(s.p); s.p = nullptr; s.cap = 0; }
}

cudaEvent_t pooled_event(Ctx* c, bool timing) {
  std::vector<cudaEvent_t>& pool = timing ? c->ev_timing : c->ev_plain;
  size_t& used = timing ? c->ev_timing_used : c->ev_plain_used;
  if (used == pool.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming);
    pool.push_back(e);
  }
  return pool[used++];
}

// ------------------------------------------------------------------ job plumbing

struct Plan {
  // rows
  bool rows_contig = true;
  int64_t row0 = 0;
  std::vector<int32_t> rows0;  // 0-based (gather mode)
  // loci of this shard, 0-based raw columns, in job order
  bool cols_contig = true;
  int64_t col0 = 0;
  std::vector<int32_t> cols0;
  int64_t n = 0, n_pad = 0, m_shard = 0;
  int prod_mask = 0;
  int nplanes = 3;
};

static int prod_mask_for(int stat) {
  switch (stat) {
    case TPG_STAT_KING: return (1 << TPG_MAT_DD) | (1 << TPG_MAT_NN) | (1 << TPG_MAT_HN);
    case TPG_STAT_IBS_RAW:
    case TPG_STAT_IBS_PROPORTION:
    case TPG_STAT_IBS_ADJUSTED: return (1 << TPG_MAT_DD) | (1 << TPG_MAT_NN) | (1 << TPG_MAT_HH);
    case TPG_STAT_AS:
    case TPG_STAT_GRM: return (1 << TPG_MAT_DD) | (1 << TPG_MAT_NN);
    default: return 0;
  }
}

static int make_plan(const tpg_job* job, Plan* pl, const Err& E) {
  if (!job) return E.fail(TPG_ERR_ARG, "job is NULL");
  if (job->abi_version != TPG_ABI_VERSION)
    return E.fail(TPG_ERR_ARG, "ABI version mismatch: job %d, library %d", job->abi_version, TPG_ABI_VERSION);
  if (job->nrow_total <= 0 || job->ncol_total <= 0)
    return E.fail(TPG_ERR_ARG, "X has no rows or columns (%lld x %lld)", (long long)job->nrow_total,
                  (long long)job->ncol_total);
  pl->prod_mask = prod_mask_for(job->stat);
  if (!pl->prod_mask) return E.fail(TPG_ERR_ARG, "unknown stat %d", job->stat);
  pl->nplanes = (pl->prod_mask & ((1 << TPG_MAT_HH) | (1 << TPG_MAT_HN))) ? 3 : 2;
  const int64_t n = job->ind_row ? job->n : job->nrow_total;
  const int64_t m = job->ind_col ? job->m : job->ncol_total;
  if (n <= 0 || m <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row or ind.col (n=%lld, m=%lld)", (long long)n, (long long)m);
  if (n > (int64_t(1) << 24)) return E.fail(TPG_ERR_ARG, "n=%lld exceeds the supported 2^24 individuals", (long long)n);
  // every count (DD, NN, HH, HN <= m) lives in an int32 matrix; finish combines them in FP64
  if (m >= (int64_t(1) << 31) - 1)
    return E.fail(TPG_ERR_ARG, "m=%lld loci exceed what the int32 count matrices hold (2^31 - 2); split ind.col", (long long)m);
  pl->n = n;
  pl->n_pad = round_up(n, kRowPad);
  // rows: R semantics, 1-based, no negatives (R/snp_king.R:12-15); bounds as SubBMAcc checks them
  if (job->ind_row) {
    bool contig = true;
    for (int64_t k = 0; k < n; ++k) {
      const int64_t r = job->ind_row[k];
      if (r < 1 || r > job->nrow_total)
        return E.fail(TPG_ERR_ARG, "ind.row[%lld] = %lld is out of bounds [1, %lld]", (long long)(k + 1),
                      (long long)r, (long long)job->nrow_total);
      if (r != int64_t(job->ind_row[0]) + k) contig = false;
    }
    pl->rows_contig = contig;
    if (contig) {
      pl->row0 = job->ind_row[0] - 1;
    } else {
      pl->rows0.resize(n);
      for (int64_t k = 0; k < n; ++k) pl->rows0[k] = job->ind_row[k] - 1;
    }
  } else {
    pl->rows_contig = true;
    pl->row0 = 0;
  }
  // loci of this shard
  int64_t c_lo = 0, c_hi = m;
  if (job->shard_mode == TPG_SHARD_LOCI && job->shard_world > 1) {
    if (job->shard_rank < 0 || job->shard_rank >= job->shard_world)
      return E.fail(TPG_ERR_ARG, "shard_rank %d out of range", job->shard_rank);
    c_lo = m * job->shard_rank / job->shard_world;
    c_hi = m * (job->shard_rank + 1) / job->shard_world;
  } else if (job->shard_mode == TPG_SHARD_TILES && job->shard_world > 1) {
    if (job->shard_rank < 0 || job->shard_rank >= job->shard_world)
      return E.fail(TPG_ERR_ARG, "shard_rank %d out of range", job->shard_rank);
  }
  pl->m_shard = c_hi - c_lo;
  if (job->ind_col) {
    bool contig = true;
    for (int64_t k = 0; k < m; ++k) {
      const int64_t c = job->ind_col[k];
      if (c < 1 || c > job->ncol_total)
        return E.fail(TPG_ERR_ARG, "ind.col[%lld] = %lld is out of bounds [1, %lld]", (long long)(k + 1),
                      (long long)c, (long long)job->ncol_total);
    }
    for (int64_t k = c_lo; k < c_hi; ++k)
      if (job->ind_col[k] != int64_t(job->ind_col[c_lo]) + (k - c_lo)) { contig = false; break; }
    pl->cols_contig = contig;
    if (pl->m_shard > 0) {
      if (contig) {
        pl->col0 = job->ind_col[c_lo] - 1;
      } else {
        pl->cols0.resize(pl->m_shard);
        for (int64_t k = c_lo; k < c_hi; ++k) pl->cols0[k - c_lo] = job->ind_col[k] - 1;
      }
    }
  } else {
    pl->cols_contig = true;
    pl->col0 = c_lo;
  }
  return TPG_OK;
}

struct Mapped {
  const uint8_t* base = nullptr;
  size_t bytes = 0;
  int fd = -1;
  int fd_direct = -1;  // the same file opened O_DIRECT (io_mode 1), -1 where the filesystem refuses
  bool owned = false;
  ~Mapped() {
    if (owned && base) munmap(const_cast<uint8_t*>(base), bytes);
    if (fd >= 0) close(fd);
    if (fd_direct >= 0) close(fd_direct);
  }
};

static int open_source(const tpg_job* job, Mapped* mp, int io_mode, const Err& E) {
  const int64_t need = job->nrow_total * job->ncol_total;
  if (job->src_kind == TPG_SRC_FILE) {
    if (!job->bk_path) return E.fail(TPG_ERR_ARG, "bk_path is NULL");
    mp->fd = open(job->bk_path, O_RDONLY);
    if (mp->fd < 0) return E.fail(TPG_ERR_IO, "cannot open backing file '%s': %s", job->bk_path, strerror(errno));
    struct stat sb;
    if (fstat(mp->fd, &sb) != 0) return E.fail(TPG_ERR_IO, "fstat('%s') failed: %s", job->bk_path, strerror(errno));
    if (int64_t(sb.st_size) < need)
      return E.fail(TPG_ERR_ARG, "backing file '%s' has %lld bytes, %lld x %lld needs %lld", job->bk_path,
                    (long long)sb.st_size, (long long)job->nrow_total, (long long)job->ncol_total, (long long)need);
    void* p = mmap(nullptr, size_t(need), PROT_READ, MAP_PRIVATE, mp->fd, 0);
    if (p == MAP_FAILED) return E.fail(TPG_ERR_IO, "mmap('%s') failed: %s", job->bk_path, strerror(errno));
    madvise(p, size_t(need), MADV_SEQUENTIAL);
    // io_mode 1: reads bypass the page cache (a cold 25 GB .bk should not evict everything else and
    // pays no copy through it); tmpfs and some network filesystems refuse O_DIRECT -> buffered reads
    if (io_mode == 1) mp->fd_direct = open(job->bk_path, O_RDONLY | O_DIRECT);
    mp->base = static_cast<const uint8_t*>(p);
    mp->bytes = size_t(need);
    mp->owned = true;
    return TPG_OK;
  }
  if (job->src_kind == TPG_SRC_HOST || job->src_kind == TPG_SRC_DEVICE) {
    if (!job->bk_data) return E.fail(TPG_ERR_ARG, "bk_data is NULL");
    mp->base = job->bk_data;
    mp->bytes = size_t(need);
    return TPG_OK;
  }
  return E.fail(TPG_ERR_ARG, "unknown src_kind %d", job->src_kind);
}

// Host staging is a memory-bandwidth job: it runs on the library's persistent worker pool
// (pack.cpp; one memcpy / pack stream tops out far below what PCIe 5 can take).
template <class F>
static void parallel_for(int64_t n, int64_t grain, F&& body) {
  const std::function<void(int64_t, int64_t)> fn = body;
  HostPool::instance().run(n, grain, fn);
}

static bool is_pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// Waits for an enqueued accumulate, checks the watchdog word and folds the per-chunk event
// timings into st.
static int collect_accumulate(Ctx* ctx, tpg_stats* st, const Err& E) {
  if (!ctx->pending) return TPG_OK;
  ctx->pending = false;
  cudaError_t e = cudaStreamSynchronize(ctx->s_compute);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_decode);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_copy);
  if (e != cudaSuccess) {
    Err quiet{nullptr, 0};
    if (check_err_flag(ctx, quiet) == TPG_ERR_KERNEL) return check_err_flag(ctx, E);
    return E.fail(TPG_ERR_CUDA, "accumulate failed on device: %s", cudaGetErrorString(e));
  }
  int rc = check_err_flag(ctx, E);
  if (rc) return rc;
  for (const Ctx::ChunkEv& t : ctx->pend_ev) {
    float ms = 0;
    cudaEventElapsedTime(&ms, t.c0, t.c1); st->ms_h2d += ms;
    cudaEventElapsedTime(&ms, t.d0, t.d1); st->ms_decode += ms;
    cudaEventElapsedTime(&ms, t.g0, t.g1); st->ms_gram += ms;
  }
  return TPG_OK;
}

// scan != nullptr: the chunks feed scan_counts (through code256) instead of decode + gram; the
// per-locus arrays of *scan are indexed by the job's locus position.
// prod_mask_override != 0: compute these products instead of the ones job->stat needs (one pass
// for several statistics: tpg_pairwise_many, the count-matrix cache).
constexpr uint64_t kFnvSeed = 1469598103934665603ull;
static uint64_t fnv1a(const void* p, size_t bytes, uint64_t h);

// Identity of a .bk for the on-device genotype cache (tpg_config.cache_genotypes): device, inode, size,
// mtime and the matrix shape.  Empty = not cacheable.
static std::string geno_cache_key(const tpg_job* job, bool force = false) {
  if ((!cfg().cache_genotypes && !force) || job->src_kind != TPG_SRC_FILE || !job->bk_path) return std::string();
  struct stat sb;
  if (stat(job->bk_path, &sb) != 0) return std::string();
  char key[256];
  snprintf(key, sizeof key, "%llu:%llu:%lld:%lld.%09ld|%lldx%lld", (unsigned long long)sb.st_dev,
           (unsigned long long)sb.st_ino, (long long)sb.st_size, (long long)sb.st_mtim.tv_sec,
           (long)sb.st_mtim.tv_nsec, (long long)job->nrow_total, (long long)job->ncol_total);
  return std::string(key);
}

static int accumulate_impl(Ctx* ctx, const tpg_job* job, int32_t* d_counts, tpg_stats* st,
                           const Err& E, bool defer = false, const ScanArgs* scan = nullptr,
                           int prod_mask_override = 0, const int32_t* group_sorted = nullptr, uint64_t group_sig = 0,
                           bool force_geno = false) {
  const double t_begin = now_ms();
  const tpg_config C = cfg();
  Plan pl;
  int rc = make_plan(job, &pl, E);
  if (rc) return rc;
  if (prod_mask_override) {
    pl.prod_mask = prod_mask_override;
    pl.nplanes = (pl.prod_mask & ((1 << TPG_MAT_HH) | (1 << TPG_MAT_HN))) ? 3 : 2;
  }
  TPG_CUDA_OK(cudaSetDevice(ctx->device), E);
  if (pl.m_shard <= 0) return TPG_OK;  // more ranks than loci: nothing to add
  Mapped src;
  rc = open_source(job, &src, C.io_mode, E);
  if (rc) return rc;

  bool on_device = job->src_kind == TPG_SRC_DEVICE;
  const bool k_major = !scan && (job->flags & TPG_FLAG_PLANES_K_MAJOR) != 0;
  const bool no_overlap = (job->flags & TPG_FLAG_NO_OVERLAP) != 0;
  const int64_t nrow = job->nrow_total;
  bool rows_g = !pl.rows_contig;

  // Rows actually needed from each raw column: [rlo, rhi).  Staging copies only that span.
  int64_t rlo = pl.row0, rhi = pl.row0 + pl.n;
  if (rows_g) {
    rlo = nrow; rhi = 0;
    for (int32_t r : pl.rows0) { rlo = std::min<int64_t>(rlo, r); rhi = std::max<int64_t>(rhi, r + 1); }
  }
  // ---- on-device genotype cache (tpg_config.cache_genotypes).  A hit turns the job into a device-resident one
  // whose source is the 2-bit packed image: every chunk is gathered from it (gather_pack_kernel, packed input;
  // ind.row — or the identity — and ind.col are applied there) and decoded as any packed chunk.
  const std::string geno_key = (scan || k_major || C.no_compact) ? std::string() : geno_cache_key(job, force_geno);
  const uint8_t* dev_base = src.base;   // device-resident source: bytes, extent, column stride
  int64_t dev_bytes = int64_t(src.bytes), dev_stride = nrow;
  bool from_cache = false;
  if (!geno_key.empty() && ctx->geno_key == geno_key && ctx->d_geno.p && gather_pack_fits((rhi - (rlo & ~int64_t(3)) + 3) / 4)) {
    from_cache = true;
    on_device = true;
    dev_base = static_cast<const uint8_t*>(ctx->d_geno.p);
    dev_stride = ctx->geno_stride;
    dev_bytes = dev_stride * job->ncol_total;
    if (!rows_g) {                       // contiguous rows: the gather takes the identity index
      pl.rows0.resize(size_t(pl.n));
      for (int64_t i = 0; i < pl.n; ++i) pl.rows0[size_t(i)] = int32_t(pl.row0 + i);
      rows_g = true;
    }
    rlo &= ~int64_t(3);                  // packed input: the staged span starts at a whole byte
    st->n_geno_cached = 1;
  }
  int64_t span = on_device ? nrow : rhi - rlo;  // bytes per staged column
  // Gathered rows (ind.row in arbitrary order): the chunk is first compacted by gather_pack_kernel
  // (decode.cu: the row span staged in shared memory, gathered there, written 2-bit packed), then
  // decoded on the contiguous packed-input path.  Needs the column span to fit shared memory —
  // as raw bytes, or as 2-bit codes when the chunk crosses PCIe packed.
  const bool gp_base = rows_g && !scan && !k_major && !C.no_compact;
  const bool gp_raw_fits = gp_base && gather_pack_fits(rhi - rlo);
  const bool gp_pk_fits = gp_base && gather_pack_fits((rhi - rlo + 3) / 4);
  // 2-bit packing while staging (pack.cpp): 4x fewer bytes over PCIe.  Contiguous row window (or a
  // gathered one that gather_pack can take), MN-major planes only; scans decode through code256
  // and need the whole byte.
  const bool pack_ok = !scan && !on_device && !k_major && !(job->flags & TPG_FLAG_NO_PACK) && C.pack_mode != 1 &&
                       (!rows_g || gp_pk_fits);
  // Pinned (or cudaHostRegister'ed) source with whole contiguous columns: can be DMA'd straight from it.
  // (gathered rows that can be packed always are: one chunk form per job keeps the row index simple)
  const bool direct_ok = !on_device && job->src_kind == TPG_SRC_HOST && pl.cols_contig && is_pinned_host(src.base) &&
                         !(rows_g && pack_ok);
  // The 4-bit route: decode_planes_k4 (raw bytes or 2-bit packed chunks) + the kind::mxf4 Gram kernel
  // (twice the int8 MAC rate, half the plane bytes, exact: gram.cu).  Tile-sharded jobs too small for
  // 1024-column ownership regions use the int8 kernels; both kinds add into the same int32 counts.
  const uint8_t *owner_h = nullptr, *owner_d = nullptr;
  if (job->shard_mode == TPG_SHARD_TILES && job->shard_world > 1) {
    rc = owner_table(ctx, pl.n_pad, job->shard_world, &owner_h, &owner_d, E);
    if (rc) return rc;
  }
  const bool q4_forced_off = (job->flags & (TPG_FLAG_GRAM_I8 | TPG_FLAG_GRAM_1CTA)) || C.gram_route == 1 || C.gram_ctas == 1;
  const bool q4_tiles_ok = !(job->shard_mode == TPG_SHARD_TILES && job->shard_world > 1) ||
                           shard_region_blocks(pl.n_pad / 256, job->shard_world) >= 4;
  const bool q4_ok = !scan && !k_major && !q4_forced_off && q4_tiles_ok;
  const bool direct_only = direct_ok && !pack_ok;
  if (direct_only) { span = nrow; rlo = 0; }
  const int64_t g_span = rhi - rlo;            // rows a gather has to stage (relative row index 0 = raw row rlo)
  const bool gp_staged = gp_base && !on_device && !pack_ok && gather_pack_fits(g_span);  // raw chunks of a host source
  const bool gp_dev = on_device && (from_cache || gp_raw_fits);
  const bool gp_any = gp_dev || gp_staged || (rows_g && pack_ok);
  // device-resident gathered chunk whose span is too long for shared memory: compact through L2 instead
  const bool compact_gather = on_device && rows_g && !gp_dev && !scan && !C.no_compact;
  const int64_t pk_stride = pk2_stride(span);
  const int64_t g_stride = pk2_stride(pl.n);  // bytes per locus of a gathered (dense, packed) chunk
  // A whole-file job (all rows, all loci, in order) whose chunks cross PCIe packed fills the cache as it goes: one
  // device-to-device copy per chunk behind its H2D.  Only where HBM has room to spare (a quarter of what is free).
  bool fill_cache = false;
  if (!geno_key.empty() && !from_cache && pack_ok && !rows_g && rlo == 0 && span == nrow && pl.cols_contig && pl.col0 == 0 &&
      pl.m_shard == job->ncol_total && !direct_ok && gather_pack_fits((nrow + 3) / 4)) {
    const size_t need = size_t(pk_stride) * size_t(job->ncol_total) + 16;
    bool room = ctx->d_geno.cap >= need;
    if (!room) {
      size_t free_b = 0, total_b = 0;
      TPG_CUDA_OK(cudaMemGetInfo(&free_b, &total_b), E);
      room = double(need) <= (force_geno ? 0.9 : 0.25) * double(free_b + ctx->d_geno.cap);
    }
    if (room) {
      ctx->geno_key.clear();             // being rewritten: not valid until this job has enqueued every chunk
      rc = ensure_slot(&ctx->d_geno, need, false, "genotype cache", E);
      if (rc) return rc;
      ctx->geno_stride = pk_stride;
      fill_cache = true;
    }
  }

  // ---- chunk size (loci per GPU chunk), decoupled from R's block.size
  int64_t lc = job->gpu_chunk_loci;
  // device bytes per locus: raw double buffer (largest form in use) + gathered copy + plane double buffer.
  // Planes are 4-bit (half a byte per genotype and plane) when every chunk of the job takes the 4-bit route.
  const int64_t raw_per_locus = on_device ? 0 : std::max(direct_ok ? nrow : 0, pack_ok ? pk_stride : span);
  const bool all_q4 = q4_ok && !(C.pk2_i8 && (pack_ok || gp_any));
  auto plane_bytes = [&](int64_t loci) {
    return all_q4 ? size_t(pl.n_pad) * size_t(round_up(loci, 256) / 2) * size_t(pl.nplanes)
                  : size_t(pl.n_pad) * size_t(loci) * size_t(pl.nplanes);
  };
  const int64_t per_locus = 2 * raw_per_locus + (gp_any ? 2 * g_stride : 0) +
                            (all_q4 ? int64_t(pl.nplanes) * pl.n_pad : 2 * int64_t(pl.nplanes) * pl.n_pad);
  if (lc <= 0) {
    // device-resident bytes: nothing to overlap with (decode and gram contend for HBM when run
    // side by side — measured), so one big chunk = fewest launches and atomics
    if (on_device) {
      lc = C.device_chunk_loci > 0 ? std::max<int64_t>(kLociPad, C.device_chunk_loci) : (int64_t(1) << 20);
    } else {
      // host sources: the first chunk's staging is exposed, so chunks are small (tens of MB)
      const int64_t mb = C.host_chunk_mb > 0 ? C.host_chunk_mb : 48;
      lc = std::min<int64_t>(262144, std::max<int64_t>(16384, (mb << 20) / std::max<int64_t>(span, 1)));
    }
  }
  lc = std::min(lc, round_up(pl.m_shard, kLociPad));
  lc = std::max<int64_t>(kLociPad, lc / kLociPad * kLociPad);
  {
    // Cap by free HBM — but cudaMemGetInfo is slow, so only ask when the cached buffers
    // would have to grow for this chunk size.
    const size_t need_planes = plane_bytes(lc);
    const size_t need_raw = on_device ? 0 : size_t(raw_per_locus) * size_t(lc) + 16;
    const size_t need_gat = gp_any ? size_t(g_stride) * size_t(lc) + 16 : 0;
    const bool two = pl.m_shard > lc;
    const bool fits = ctx->d_planes[0].cap >= need_planes && (!two || ctx->d_planes[1].cap >= need_planes) &&
                      (on_device || (ctx->d_raw[0].cap >= need_raw && (!two || ctx->d_raw[1].cap >= need_raw))) &&
                      (ctx->d_gat[0].cap >= need_gat && (!two || ctx->d_gat[1].cap >= need_gat));
    if (!fits) {
      size_t free_b = 0, total_b = 0;
      TPG_CUDA_OK(cudaMemGetInfo(&free_b, &total_b), E);
      const size_t held = ctx->d_planes[0].cap + ctx->d_planes[1].cap + ctx->d_raw[0].cap + ctx->d_raw[1].cap +
                          ctx->d_gat[0].cap + ctx->d_gat[1].cap;
      // a job that would be ONE chunk needs every buffer once, not twice: C4's 37.6 GB of 4-bit planes fit
      // beside 25 GB of bytes, 40 GB of counts and the 20 GB result, and the whole job is one decode + one Gram launch
      const size_t single = need_planes + need_raw + need_gat;
      if (two || double(single) > 0.85 * double(free_b + held)) {
        const int64_t mem_cap = std::max<int64_t>(kLociPad, int64_t(double(free_b + held) * 0.45 / double(per_locus)));
        lc = std::min(lc, mem_cap);
        lc = std::max<int64_t>(kLociPad, lc / kLociPad * kLociPad);
      }
    }
  }
  const int64_t n_chunks = (pl.m_shard + lc - 1) / lc;
  const int64_t plane_stride = pl.n_pad * lc;
  // One chunk out of the packed image: its planes stay valid in d_planes[0] until another job decodes there
  // (result shells run the same rows and loci K - 1 times).
  std::string planes_key;
  if (from_cache && n_chunks == 1 && !scan) {
    char pk[160];
    snprintf(pk, sizeof pk, "|n%lld|m%lld|r%016llx|c%016llx|p%d|q%d|lc%lld", (long long)pl.n, (long long)pl.m_shard,
             (unsigned long long)fnv1a(pl.rows0.data(), sizeof(int32_t) * pl.rows0.size(), kFnvSeed),
             (unsigned long long)(pl.cols_contig ? uint64_t(pl.col0) : fnv1a(pl.cols0.data(), sizeof(int32_t) * pl.cols0.size(), kFnvSeed)),
             pl.nplanes, int(q4_ok && !C.pk2_i8), (long long)lc);
    planes_key = geno_key + pk;
  }
  const bool reuse_planes = !planes_key.empty() && ctx->planes_key == planes_key && ctx->d_planes[0].p;
  if (!scan && !reuse_planes) ctx->planes_key.clear();

  struct { uint8_t* d_raw[2]; uint8_t* d_gat[2]; int8_t* d_planes[2]; uint8_t* h_pin[2]; int32_t* d_rows; int32_t* d_cols;
           Ctx* c; cudaEvent_t ev(bool t) { return pooled_event(c, t); } } B{};
  B.c = ctx;
  if (!ctx->keep_events) ctx->ev_timing_used = ctx->ev_plain_used = 0;
  const int64_t stage_per_locus = pack_ok ? pk_stride : span;  // pinned staging bytes per locus
  for (int b = 0; b < 2 && b < n_chunks; ++b) {
    if (!scan) {
      rc = ensure_slot(&ctx->d_planes[b], plane_bytes(lc), false, "plane storage", E);
      if (rc) return rc;
      B.d_planes[b] = static_cast<int8_t*>(ctx->d_planes[b].p);
    }
    if (gp_any) {
      rc = ensure_slot(&ctx->d_gat[b], size_t(g_stride * lc) + 16, false, "gathered chunk storage", E);
      if (rc) return rc;
      B.d_gat[b] = static_cast<uint8_t*>(ctx->d_gat[b].p);
    }
    if (!on_device) {
      rc = ensure_slot(&ctx->d_raw[b], size_t(raw_per_locus * lc) + 16, false, "raw chunk storage", E);
      if (rc) return rc;
      B.d_raw[b] = static_cast<uint8_t*>(ctx->d_raw[b].p);
      if (!direct_only) {
        rc = ensure_slot(&ctx->h_pin[b], size_t(stage_per_locus * lc), true, "pinned staging", E);
        if (rc) return rc;
        B.h_pin[b] = static_cast<uint8_t*>(ctx->h_pin[b].p);
      }
    }
  }
  if (rows_g) {
    // row offsets as the kernels see them: relative to raw row rlo for staged / packed chunks and for
    // every gather_pack input; absolute for device-resident chunks on the older paths
    std::vector<int32_t> rel(pl.rows0);
    if (!on_device || gp_dev) for (auto& r : rel) r -= int32_t(rlo);  // (direct-only: rlo = 0)
    rc = ensure_slot(&ctx->d_rows, sizeof(int32_t) * (pl.n + 16), false, "row index", E);
    if (rc) return rc;
    B.d_rows = static_cast<int32_t*>(ctx->d_rows.p);
    // (stream-ordered after the previous job's kernels that read the old index)
    TPG_CUDA_OK(cudaStreamSynchronize(ctx->s_decode), E);
    TPG_CUDA_OK(cudaStreamSynchronize(ctx->s_compute), E);
    TPG_CUDA_OK(cudaMemcpy(B.d_rows, rel.data(), sizeof(int32_t) * pl.n, cudaMemcpyHostToDevice), E);
  }
  if (on_device && !pl.cols_contig) {
    rc = ensure_slot(&ctx->d_cols, sizeof(int32_t) * pl.m_shard, false, "column index", E);
    if (rc) return rc;
    B.d_cols = static_cast<int32_t*>(ctx->d_cols.p);
    TPG_CUDA_OK(cudaStreamSynchronize(ctx->s_decode), E);
    TPG_CUDA_OK(cudaStreamSynchronize(ctx->s_compute), E);
    TPG_CUDA_OK(cudaMemcpy(B.d_cols, pl.cols0.data(), sizeof(int32_t) * pl.m_shard, cudaMemcpyHostToDevice), E);
  }

  cudaEvent_t ev_h2d_done[2] = {B.ev(false), B.ev(false)};
  cudaEvent_t ev_dec_done[2] = {B.ev(false), B.ev(false)};
  cudaEvent_t ev_gram_done[2] = {B.ev(false), B.ev(false)};
  // decode must not start before everything already queued on the compute stream (the
  // caller's zeroing of the counts, a previous job's finish) — cheap, and keeps jobs ordered.
  cudaEvent_t ev_start = B.ev(false);
  TPG_CUDA_OK(cudaEventRecord(ev_start, ctx->s_compute), E);
  // one device-resident chunk: nothing to overlap, so decode runs on the compute stream itself and
  // the step pays no cross-stream event hops
  cudaStream_t s_dec = (n_chunks == 1 && on_device) ? ctx->s_compute : ctx->s_decode;
  TPG_CUDA_OK(cudaStreamWaitEvent(s_dec, ev_start, 0), E);
  if (from_cache && ctx->ev_geno) TPG_CUDA_OK(cudaStreamWaitEvent(s_dec, ctx->ev_geno, 0), E);
  std::vector<Ctx::ChunkEv>& tm = ctx->pend_ev;
  tm.assign(n_chunks, Ctx::ChunkEv{});
  int64_t macs = 0, max_chain = 0, n_wide = 0;
  double ms_read = 0;
  // hybrid link/core scheduling state (see the chunk loop)
  cudaEvent_t ev_pin_free[2] = {B.ev(false), B.ev(false)};
  int64_t n_staged = 0, n_packed = 0, n_q4 = 0, n_i8 = 0, n_dio = 0;
  double link_free_at = 0;
  const double link_gbps = C.link_gbps > 0 ? std::max(1.0, C.link_gbps) : 45.0;
  double pack_gbps = 4.0 * double(HostPool::max_threads());
  const bool hybrid = C.pack_mode != 2;
  const double hybrid_factor = C.hybrid_factor > 0 ? C.hybrid_factor : 0.9;
  // .bk files are read with pread() — whole columns in ~1 MB blocks into per-thread buffers that stay
  // in L2 — unless io_mode asks for the mmap; with io_mode = 1 the reads bypass the page cache (O_DIRECT)
  const bool use_pread = job->src_kind == TPG_SRC_FILE && C.io_mode != 2;
  const bool direct_io = use_pread && src.fd_direct >= 0;
  std::atomic<int> io_err{0};        // errno of a failed read (captured in the worker), -1 = unexpected EOF
  std::atomic<int64_t> file_bytes{0};

  // From the first enqueue on, an early error return must not leave device work in flight on the
  // three streams: the next job on the cached context would reuse the staging and raw buffers.
  struct DrainGuard {
    Ctx* c; bool armed;
    ~DrainGuard() {
      if (!armed) return;
      cudaStreamSynchronize(c->s_copy);
      cudaStreamSynchronize(c->s_decode);
      cudaStreamSynchronize(c->s_compute);
      c->pending = false;
    }
  } drain{ctx, true};

  // ---- host-source staging, software-pipelined: the workers stage (pack / copy) chunk c+1 while
  // this thread enqueues the H2D, decode and Gram of chunk c.
  struct Stage {
    bool use_direct = false, packed = false, async = false;
    const uint8_t* hp = nullptr;
    int64_t h2d_bytes = 0, col_stride = 0;
    int sb = 0;  // pinned staging buffer
  };
  Stage stg[2];
  // Decides how chunk c crosses PCIe and starts its staging:
  //   direct  raw bytes DMA'd from the caller's pinned memory (no host work, 1 B / genotype)
  //   pk2     packed 4 genotypes / byte into pinned staging by the worker pool (0.25 B / genotype)
  //   staged  raw row window copied into pinned staging (scans, K-major planes, spans too long to gather packed)
  // With both direct and pk2 available the link and the cores work side by side: a chunk goes
  // raw when the copy queue would otherwise run dry before this chunk could be packed.
  auto start_stage = [&](int64_t c) -> int {
    Stage& S = stg[c & 1];
    S = Stage{};
    const int64_t k0 = c * lc, k1 = std::min(pl.m_shard, k0 + lc);
    const int64_t lcc = k1 - k0;
    const double t0 = now_ms();
    S.use_direct = direct_only;
    if (direct_ok && pack_ok) {
      const double backlog = std::max(0.0, link_free_at - t0);
      const double t_pack = double(lcc * span) / (pack_gbps * 1e6);
      S.use_direct = hybrid && backlog < hybrid_factor * t_pack;
    }
    if (S.use_direct) {
      S.hp = src.base + (pl.col0 + k0) * nrow;
      S.h2d_bytes = lcc * nrow;
      S.col_stride = nrow;
      return TPG_OK;
    }
    S.sb = int(n_staged & 1);
    if (n_staged >= 2) TPG_CUDA_OK(cudaEventSynchronize(ev_pin_free[S.sb]), E);
    ++n_staged;
    uint8_t* stage = B.h_pin[S.sb];
    const uint8_t* base = src.base;
    const Plan* plp = &pl;
    S.hp = stage;
    S.async = true;
    S.packed = pack_ok;
    const int64_t ostride = pack_ok ? pk_stride : span;  // staged bytes per locus
    S.h2d_bytes = lcc * ostride;
    S.col_stride = ostride;
    const bool do_pack = pack_ok;
    if (use_pread) {
      // Runs of consecutive raw columns are read with one pread() each (whole columns: the rows outside
      // [rlo, rhi) of a window are skipped only when the window is short) into a per-thread buffer and
      // packed / copied from there — a fresh mmap pays one page fault per 4 KiB.  O_DIRECT: offsets,
      // lengths and the buffer are 4 KiB aligned; the file's tail is a short read.
      const int fd = direct_io ? src.fd_direct : src.fd;
      const int64_t cpb = std::max<int64_t>(1, (int64_t(1) << 20) / nrow);  // columns per block
      std::atomic<int>* bad = &io_err;
      std::atomic<int64_t>* fbytes = &file_bytes;
      const bool dio = direct_io;
      const int64_t file_size = int64_t(src.bytes);
      HostPool::instance().begin(lcc, std::max<int64_t>(cpb, (int64_t(1) << 20) / std::max<int64_t>(span, 1)), [=](int64_t a, int64_t e) {
        static thread_local uint8_t* buf = nullptr;
        static thread_local size_t buf_cap = 0;
        const size_t need = size_t(cpb * nrow) + 2 * 4096;
        if (buf_cap < need) {
          free(buf);
          buf = static_cast<uint8_t*>(aligned_alloc(4096, (need + 4095) / 4096 * 4096));
          buf_cap = buf ? need : 0;
          if (!buf) { bad->store(ENOMEM); return; }
        }
        int64_t k = k0 + a;
        while (k < k0 + e) {
          const int64_t c_first = plp->cols_contig ? plp->col0 + k : plp->cols0[k];
          int64_t nc = 1;
          while (k + nc < k0 + e && nc < cpb &&
                 (plp->cols_contig || plp->cols0[k + nc] == plp->cols0[k + nc - 1] + 1)) ++nc;
          // a short row window of long columns: read just the window of each column
          const bool window = span * 4 < nrow && !dio;
          const int64_t reads = window ? nc : 1;
          for (int64_t q = 0; q < reads; ++q) {
            const int64_t off = window ? (c_first + q) * nrow + rlo : c_first * nrow;
            const int64_t len = window ? span : nc * nrow;
            int64_t a0 = off, a1 = off + len;
            if (dio) { a0 = off / 4096 * 4096; a1 = std::min<int64_t>((off + len + 4095) / 4096 * 4096, (file_size + 4095) / 4096 * 4096); }
            uint8_t* dst = window ? buf + q * span : buf;
            int64_t got = 0;
            while (a0 + got < off + len) {
              const ssize_t r = pread(fd, dst + got, size_t(a1 - a0 - got), off_t(a0 + got));
              if (r < 0) { if (errno == EINTR) continue; bad->store(errno ? errno : EIO); return; }
              if (r == 0) { bad->store(-1); return; }
              got += r;
            }
            fbytes->fetch_add(got);
            if (!window) {
              const uint8_t* cols = buf + (off - a0);
              for (int64_t j = 0; j < nc; ++j) {
                const uint8_t* s0 = cols + j * nrow + rlo;
                uint8_t* d0 = stage + (k - k0 + j) * ostride;
                if (do_pack) pk2_pack_column(s0, span, d0, ostride);
                else memcpy(d0, s0, size_t(span));
              }
            }
          }
          if (window)
            for (int64_t j = 0; j < nc; ++j) {
              uint8_t* d0 = stage + (k - k0 + j) * ostride;
              if (do_pack) pk2_pack_column(buf + j * span, span, d0, ostride);
              else memcpy(d0, buf + j * span, size_t(span));
            }
          k += nc;
        }
      });
    } else if (do_pack) {
      HostPool::instance().begin(lcc, std::max<int64_t>(1, (int64_t(1) << 20) / span), [=](int64_t a, int64_t e) {
        for (int64_t k = k0 + a; k < k0 + e; ++k) {
          const int64_t col = plp->cols_contig ? plp->col0 + k : plp->cols0[k];
          pk2_pack_column(base + col * nrow + rlo, span, stage + (k - k0) * pk_stride, pk_stride);
        }
      });
    } else if (pl.cols_contig && span == nrow) {
      const uint8_t* from = base + (pl.col0 + k0) * nrow;
      HostPool::instance().begin(lcc * nrow, int64_t(8) << 20, [=](int64_t a, int64_t e) { memcpy(stage + a, from + a, size_t(e - a)); });
    } else {
      HostPool::instance().begin(lcc, std::max<int64_t>(1, (int64_t(4) << 20) / span), [=](int64_t a, int64_t e) {
        for (int64_t k = k0 + a; k < k0 + e; ++k) {
          const int64_t col = plp->cols_contig ? plp->col0 + k : plp->cols0[k];
          memcpy(stage + (k - k0) * span, base + col * nrow + rlo, size_t(span));
        }
      });
    }
    return TPG_OK;
  };
  // Waits for chunk c's staging, books its transfer in the link model.
  auto finish_stage = [&](int64_t c) -> int {
    Stage& S = stg[c & 1];
    if (S.async) {
      const double ms = HostPool::instance().end();
      S.async = false;
      ms_read += ms;
      const int64_t lcc = std::min(pl.m_shard, (c + 1) * lc) - c * lc;
      if (S.packed && ms > 0) pack_gbps = 0.5 * pack_gbps + 0.5 * double(lcc * span) / (ms * 1e6);
      if (const int err = io_err.load())
        return E.fail(TPG_ERR_IO, "reading '%s' failed: %s", job->bk_path,
                      err < 0 ? "unexpected end of file (backing file shorter than nrow x ncol?)" : strerror(err));
      if (use_pread && direct_io) ++n_dio;
    }
    link_free_at = std::max(link_free_at, now_ms()) + double(S.h2d_bytes) / (link_gbps * 1e6);
    return TPG_OK;
  };
  // (an early error return must not leave the pool's job open)
  struct StageGuard {
    Stage* s;
    ~StageGuard() { for (int i = 0; i < 2; ++i) if (s[i].async) { HostPool::instance().end(); s[i].async = false; } }
  } stage_guard{stg};
  if (!on_device) { rc = start_stage(0); if (rc) return rc; }

  for (int64_t c = 0; c < n_chunks; ++c) {
    const int b = int(c & 1);
    const int64_t k0 = c * lc, k1 = std::min(pl.m_shard, k0 + lc);
    const int64_t lcc = k1 - k0;
    const int64_t lc_pad = round_up(lcc, kLociPad);
    tm[c] = {B.ev(true), B.ev(true), B.ev(true), B.ev(true), B.ev(true), B.ev(true)};

    DecodeArgs da;
    da.n = pl.n;
    da.n_pad = pl.n_pad;
    da.lc = lcc;
    da.lc_pad = lc_pad;
    da.planes = B.d_planes[b];
    // planes of a short last chunk keep the full-chunk plane stride (one allocation)
    da.plane_stride = plane_stride;
    da.nplanes = pl.nplanes;
    da.k_major = k_major;
    da.row_idx = B.d_rows;
    bool gather_now = false;     // this chunk goes through gather_pack before it is decoded
    bool gather_packed_in = false;
    int64_t gather_row_lo = 0;
    if (on_device) {
      da.raw = dev_base;
      da.raw_bytes = dev_bytes;
      da.raw_stride = dev_stride;
      da.row0 = pl.row0;
      da.col_idx = pl.cols_contig ? nullptr : B.d_cols + k0;
      da.col0 = pl.col0 + k0;
      gather_now = gp_dev;
      gather_packed_in = from_cache;
      gather_row_lo = rlo;
      TPG_CUDA_OK(cudaEventRecord(tm[c].c0, ctx->s_copy), E);
      TPG_CUDA_OK(cudaEventRecord(tm[c].c1, ctx->s_copy), E);
    } else {
      rc = finish_stage(c);
      if (rc) return rc;
      const Stage S = stg[c & 1];
      if (c + 1 < n_chunks) { rc = start_stage(c + 1); if (rc) return rc; }  // the workers go on while we enqueue
      const bool use_direct = S.use_direct, packed = S.packed;
      const uint8_t* hp = S.hp;
      const int64_t h2d_bytes = S.h2d_bytes, col_stride = S.col_stride;
      // ---- H2D on the copy stream, after the decode that last read this raw buffer
      if (c >= 2) TPG_CUDA_OK(cudaStreamWaitEvent(ctx->s_copy, ev_dec_done[b], 0), E);
      TPG_CUDA_OK(cudaEventRecord(tm[c].c0, ctx->s_copy), E);
      TPG_CUDA_OK(cudaMemcpyAsync(B.d_raw[b], hp, size_t(h2d_bytes), cudaMemcpyHostToDevice, ctx->s_copy), E);
      TPG_CUDA_OK(cudaEventRecord(tm[c].c1, ctx->s_copy), E);
      if (fill_cache && packed && col_stride == ctx->geno_stride)
        TPG_CUDA_OK(cudaMemcpyAsync(static_cast<uint8_t*>(ctx->d_geno.p) + k0 * col_stride, B.d_raw[b], size_t(h2d_bytes),
                                    cudaMemcpyDeviceToDevice, ctx->s_copy), E);
      else fill_cache = false;           // (never for a file source; a chunk that crossed raw would leave a hole)
      TPG_CUDA_OK(cudaEventRecord(ev_h2d_done[b], ctx->s_copy), E);
      if (!use_direct) TPG_CUDA_OK(cudaEventRecord(ev_pin_free[S.sb], ctx->s_copy), E);
      TPG_CUDA_OK(cudaStreamWaitEvent(s_dec, ev_h2d_done[b], 0), E);
      st->bytes_h2d += h2d_bytes;
      if (packed) n_packed++;
      da.raw = B.d_raw[b];
      da.raw_bytes = h2d_bytes;
      da.raw_stride = col_stride;
      da.packed = packed;
      // staged / packed chunks start at raw row rlo; direct chunks hold whole columns
      da.row0 = packed ? 0 : (pl.rows_contig ? pl.row0 - (use_direct ? 0 : rlo) : 0);
      da.col_idx = nullptr;
      da.col0 = 0;
      if (rows_g && (packed || gp_staged)) {
        gather_now = true;
        gather_packed_in = packed;
        gather_row_lo = use_direct ? rlo : 0;  // (a direct chunk of a gathered job only exists direct-only: rlo = 0)
      }
    }
    if (scan) {  // one pass over the chunk's bytes; no planes, no products
      ScanArgs sa = *scan;
      sa.alt2 += k0; sa.valid += k0; sa.na_l += k0;
      TPG_CUDA_OK(cudaEventRecord(tm[c].d0, s_dec), E);
      rc = launch_scan(ctx, da, sa, s_dec, E);
      if (rc) return rc;
      TPG_CUDA_OK(cudaEventRecord(tm[c].d1, s_dec), E);
      TPG_CUDA_OK(cudaEventRecord(ev_dec_done[b], s_dec), E);
      TPG_CUDA_OK(cudaStreamWaitEvent(ctx->s_compute, ev_dec_done[b], 0), E);
      TPG_CUDA_OK(cudaEventRecord(tm[c].g0, ctx->s_compute), E);
      TPG_CUDA_OK(cudaEventRecord(tm[c].g1, ctx->s_compute), E);
      continue;
    }
    // planes[b] (and the gathered copy b) were last read by the gram / decode of chunk c-2
    if (c >= 2) TPG_CUDA_OK(cudaStreamWaitEvent(s_dec, ev_gram_done[b], 0), E);
    TPG_CUDA_OK(cudaEventRecord(tm[c].d0, s_dec), E);
    if (gather_now && !reuse_planes) {
      // gathered rows: stage the row span in shared memory, gather there, write the dense chunk 2-bit
      // packed (decode.cu, "row gather through shared memory"); ind.col of a device source is applied too
      rc = launch_gather_pack(ctx, da, gather_packed_in, gather_row_lo, g_span, B.d_gat[b], g_stride, s_dec, E);
      if (rc) return rc;
      da.raw = B.d_gat[b];
      da.raw_bytes = g_stride * lcc;
      da.raw_stride = g_stride;
      da.packed = true;
      da.row_idx = nullptr;
      da.row0 = 0;
      da.col_idx = nullptr;
      da.col0 = 0;
    } else if (gather_now) {
      da.packed = true;                   // planes reused: the same plane form as the gathered chunk got
    }
    // (2-bit packed chunks decode to 4-bit planes too; pk2_i8 keeps them on the int8 kernels)
    const bool use_q4 = q4_ok && !(da.packed && C.pk2_i8);
    const int64_t q4_lc_pad = round_up(lcc, 256);
    const int64_t q4_stride = pl.n_pad * (round_up(lc, 256) / 2);  // <= plane_stride: fits the int8 plane buffer
    if (use_q4) {
      da.k4 = true;
      da.lc_pad = q4_lc_pad;
      da.plane_stride = q4_stride;
    }
    if (compact_gather && use_q4) {
      // span too long for shared memory: compact through L2 (compact_rows_kernel), then the 4-bit decode
      // takes its contiguous fast path on the dense copy (raw buffer b is free: device sources stage nothing)
      rc = ensure_slot(&ctx->d_raw[b], size_t(pl.n_pad) * size_t(lcc) + 64, false, "compacted chunk", E);
      if (rc) return rc;
      uint8_t* dense = static_cast<uint8_t*>(ctx->d_raw[b].p);
      rc = launch_compact_rows(ctx, da, dense, s_dec, E);
      if (rc) return rc;
      da.raw = dense;
      da.raw_bytes = pl.n_pad * lcc + 64;
      da.raw_stride = pl.n_pad;
      da.row_idx = nullptr;
      da.row0 = 0;
      da.col_idx = nullptr;
      da.col0 = 0;
    }
    if (!reuse_planes) {
      rc = launch_decode(ctx, da, s_dec, E);
      if (rc) return rc;
    }
    TPG_CUDA_OK(cudaEventRecord(tm[c].d1, s_dec), E);
    TPG_CUDA_OK(cudaEventRecord(ev_dec_done[b], s_dec), E);
    TPG_CUDA_OK(cudaStreamWaitEvent(ctx->s_compute, ev_dec_done[b], 0), E);

    GramArgs ga;
    ga.planes = B.d_planes[b];
    ga.plane_stride = use_q4 ? q4_stride : plane_stride;
    ga.n = pl.n;
    ga.n_pad = pl.n_pad;
    ga.lc_pad = use_q4 ? q4_lc_pad : lc_pad;
    ga.mxf4 = use_q4;
    if (use_q4) n_q4++; else n_i8++;
    ga.k_major = k_major;
    ga.ctas = (job->flags & TPG_FLAG_GRAM_1CTA) ? 1 : 0;
    ga.prod_mask = pl.prod_mask;
    ga.shard_rank = job->shard_mode == TPG_SHARD_TILES ? job->shard_rank : 0;
    ga.shard_world = job->shard_mode == TPG_SHARD_TILES ? job->shard_world : 1;
    ga.owner = owner_h;
    ga.counts = d_counts;
    ga.ld = pl.n_pad;
    ga.max_chain_out = &max_chain;
    ga.wide_out = &n_wide;
    ga.group_sorted = group_sorted;
    ga.group_sig = group_sig;
    int64_t mm = 0;
    TPG_CUDA_OK(cudaEventRecord(tm[c].g0, ctx->s_compute), E);
    rc = launch_gram(ctx, ga, ctx->s_compute, &mm, E);
    if (rc) return rc;
    macs += mm;
    TPG_CUDA_OK(cudaEventRecord(tm[c].g1, ctx->s_compute), E);
    TPG_CUDA_OK(cudaEventRecord(ev_gram_done[b], ctx->s_compute), E);
    if (no_overlap) TPG_CUDA_OK(cudaStreamSynchronize(ctx->s_compute), E);
  }
  if (!planes_key.empty()) ctx->planes_key = planes_key;
  if (fill_cache) {                      // every chunk is on its way: the image is valid once s_copy gets here
    if (!ctx->ev_geno) TPG_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_geno, cudaEventDisableTiming), E);
    TPG_CUDA_OK(cudaEventRecord(ctx->ev_geno, ctx->s_copy), E);
    ctx->geno_key = geno_key;
  }
  drain.armed = false;
  ctx->pending = true;
  if (!defer) {
    rc = collect_accumulate(ctx, st, E);
    if (rc) return rc;
  }
  st->ms_read += ms_read;
  st->n_chunks += n_chunks;
  st->n_packed_chunks += n_packed;
  st->n_mxf4_chunks += n_q4;
  st->n_i8_chunks += n_i8;
  st->n_wide_launches += n_wide;
  st->max_chain_loci = std::max(st->max_chain_loci, max_chain);
  st->bytes_file += file_bytes.load();
  st->n_direct_io += n_dio;
  st->mma_macs += macs;
  st->ms_total += now_ms() - t_begin;
  return TPG_OK;
}

// D2H of up to `count` result buffers of `bytes` each (null host pointers are skipped), enqueued
// on the compute stream: through two pinned bounce buffers, the pieces copied into the caller's
// (pageable, usually untouched) memory by the worker pool while the next piece is in flight — a
// cudaMemcpy straight into pageable memory runs at ~3 GB/s, first-touch faults included.
static int results_to_host(Ctx* ctx, const void* const* dev, void* const* host, int count, size_t bytes,
                           int64_t* bytes_d2h, const Err& E) {
  // (128 MB pieces: handing a piece to the worker pool costs a wake-up and a tail; at 32 MB that was a third of the
  //  0.9 ms a piece took)
  const size_t kPiece = size_t(128) << 20;
  const size_t piece = std::min(bytes, kPiece);
  int rc = TPG_OK;
  for (int b = 0; b < 2 && rc == TPG_OK; ++b)
    rc = ensure_slot(&ctx->h_out[b], piece, true, "pinned result staging", E);
  if (rc) return rc;
  struct Piece { uint8_t* dst; const uint8_t* from; size_t len; cudaEvent_t ev; };
  Piece prev{nullptr, nullptr, 0, nullptr};
  auto drain = [&](const Piece& q) {
    if (!q.dst) return cudaSuccess;
    cudaError_t e = cudaEventSynchronize(q.ev);
    if (e != cudaSuccess) return e;
    uint8_t* dst = q.dst;
    const uint8_t* from = q.from;
    parallel_for(int64_t(q.len), int64_t(256) << 10, [=](int64_t a, int64_t z) { copy_streaming(dst + a, from + a, size_t(z - a)); });
    return cudaSuccess;
  };
  cudaEvent_t evp[2] = {pooled_event(ctx, false), pooled_event(ctx, false)};
  int64_t idx = 0;
  for (int o = 0; o < count && rc == TPG_OK; ++o) {
    if (!host[o]) continue;
    // A fresh result matrix is untouched memory: the copy below is bound by first-touch page
    // faults (4 KiB each).  Ask for transparent huge pages on its 2 MiB-aligned interior; the
    // call is advisory and fails harmlessly where THP is off.
    if (bytes >= (size_t(8) << 20) && !cfg().no_thp) {
      const uintptr_t lo = (reinterpret_cast<uintptr_t>(host[o]) + (size_t(2) << 20) - 1) & ~((uintptr_t(2) << 20) - 1);
      const uintptr_t hi = (reinterpret_cast<uintptr_t>(host[o]) + bytes) & ~((uintptr_t(2) << 20) - 1);
      if (hi > lo) madvise(reinterpret_cast<void*>(lo), size_t(hi - lo), MADV_HUGEPAGE);
    }
    for (size_t off = 0; off < bytes && rc == TPG_OK; off += piece, ++idx) {
      const int b = int(idx & 1);
      const size_t len = std::min(piece, bytes - off);
      uint8_t* bounce = static_cast<uint8_t*>(ctx->h_out[b].p);
      cudaError_t e = cudaMemcpyAsync(bounce, static_cast<const uint8_t*>(dev[o]) + off, len,
                                      cudaMemcpyDeviceToHost, ctx->s_compute);
      if (e == cudaSuccess) e = cudaEventRecord(evp[b], ctx->s_compute);
      if (e == cudaSuccess) e = drain(prev);  // the other bounce buffer is free again after this
      if (e != cudaSuccess) rc = E.fail(TPG_ERR_CUDA, "D2H of result failed: %s", cudaGetErrorString(e));
      prev = Piece{static_cast<uint8_t*>(host[o]) + off, bounce, len, evp[b]};
      if (bytes_d2h) *bytes_d2h += int64_t(len);
    }
  }
  if (rc == TPG_OK && drain(prev) != cudaSuccess) rc = E.fail(TPG_ERR_CUDA, "D2H of result failed");
  return rc;
}

static int finish_impl(Ctx* ctx, const tpg_job* job, const int32_t* d_counts, tpg_result* res,
                       const Err& E) {
  if (!res || !res->out0) return E.fail(TPG_ERR_ARG, "result / out0 is NULL");
  const double t_begin = now_ms();
  TPG_CUDA_OK(cudaSetDevice(ctx->device), E);
  const int64_t n = job->ind_row ? job->n : job->nrow_total;
  const int64_t m = job->ind_col ? job->m : job->ncol_total;
  const int64_t n_pad = round_up(n, kRowPad);
  int64_t n_short = 0;
  if ((job->stat == TPG_STAT_AS || job->stat == TPG_STAT_GRM) && job->pad_quirk) {
    const int64_t bs = job->block_size > 0 ? job->block_size : block_size_default(job->nrow_total);
    n_short = as_pad_count(m, bs);
  }
  res->stats.n_short = n_short;
  FinishArgs fa;
  fa.stat = job->stat;
  fa.counts = d_counts;
  fa.ld = n_pad;
  fa.n = n;
  fa.m = m;
  fa.n_short = n_short;
  fa.shard_rank = job->shard_mode == TPG_SHARD_TILES ? job->shard_rank : 0;
  fa.shard_world = job->shard_mode == TPG_SHARD_TILES ? job->shard_world : 1;
  if (fa.shard_world > 1) {
    const uint8_t* owner_h = nullptr;
    int rc0 = owner_table(ctx, n_pad, fa.shard_world, &owner_h, &fa.owner, E);
    if (rc0) return rc0;
  }
  double* host[4] = {res->out0, res->out1, res->aux0, res->aux1};
  double* dev[4] = {nullptr, nullptr, nullptr, nullptr};
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  for (int o = 0; o < 4; ++o) {
    if (!host[o]) continue;
    if (res->out_on_device) { dev[o] = host[o]; continue; }
    int rc0 = ensure_slot(&ctx->d_out[o], bytes, false, "result matrix", E);
    if (rc0) return rc0;
    dev[o] = static_cast<double*>(ctx->d_out[o].p);
    if (fa.shard_world > 1)  // blocks of other ranks: leave the caller's values untouched
      TPG_CUDA_OK(cudaMemcpyAsync(dev[o], host[o], bytes, cudaMemcpyHostToDevice, ctx->s_compute), E);
  }
  fa.out0 = dev[0]; fa.out1 = dev[1]; fa.aux0 = dev[2]; fa.aux1 = dev[3];
  if (!ctx->pending) ctx->ev_timing_used = 0;
  cudaEvent_t e0 = pooled_event(ctx, true), e1 = pooled_event(ctx, true), e2 = pooled_event(ctx, true);
  cudaEventRecord(e0, ctx->s_compute);
  int rc = launch_finish(ctx, fa, ctx->s_compute, E);
  cudaEventRecord(e1, ctx->s_compute);
  if (rc == TPG_OK && !res->out_on_device) {
    const void* dv[4] = {dev[0], dev[1], dev[2], dev[3]};
    void* hv[4] = {host[0], host[1], host[2], host[3]};
    rc = results_to_host(ctx, dv, hv, 4, bytes, &res->stats.bytes_d2h, E);
  }
  cudaEventRecord(e2, ctx->s_compute);
  if (ctx->pending) {  // the accumulate of the same job was enqueued without a sync
    char keep[512];
    Err E2{keep, sizeof keep};
    keep[0] = 0;
    const int rc2 = collect_accumulate(ctx, &res->stats, E2);
    // a tripped pipeline watchdog explains a failed D2H, not the other way round
    if (rc2 != TPG_OK && (rc == TPG_OK || rc2 == TPG_ERR_KERNEL)) rc = E.fail(rc2, "%s", keep);
  }
  cudaError_t e = cudaStreamSynchronize(ctx->s_compute);
  if (rc == TPG_OK && e != cudaSuccess) rc = E.fail(TPG_ERR_CUDA, "finish failed on device: %s", cudaGetErrorString(e));
  if (rc == TPG_OK) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1); res->stats.ms_finish += ms;
    cudaEventElapsedTime(&ms, e1, e2); res->stats.ms_d2h += ms;  // DMA time; the host copy ran inside it
  }
  res->stats.ms_total += now_ms() - t_begin;
  res->stats.n_launches = ctx->launches;
  return rc;
}

// One cached context per device for the one-call-per-job entry point (R is single threaded;
// the mutex only protects the table).
static std::mutex g_mu;
static Ctx* g_ctx[64] = {nullptr};

static Ctx* cached_ctx(int device, const Err& E, int* rc) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (device < 0 || device >= 64) { *rc = E.fail(TPG_ERR_ARG, "device %d out of range", device); return nullptr; }
  if (g_ctx[device]) return g_ctx[device];
  Ctx* c = new Ctx();
  *rc = ctx_init(c, device, E);
  if (*rc) { delete c; return nullptr; }
  g_ctx[device] = c;
  return c;
}

// One context per (rank slot) for the in-process multi-GPU driver; a slot is rebuilt when
// the caller maps it to another device.
static Ctx* g_rank_ctx[64] = {nullptr};

static Ctx* rank_ctx(int slot, int device, const Err& E, int* rc) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot < 0 || slot >= 64) { *rc = E.fail(TPG_ERR_ARG, "too many devices"); return nullptr; }
  Ctx* c = g_rank_ctx[slot];
  if (c && c->device == device) return c;
  if (c) {
    trim_ctx(c);
    g_rank_ctx[slot] = nullptr;  // streams/events of the old slot are leaked on purpose (rare path)
  }
  c = new Ctx();
  *rc = ctx_init(c, device, E);
  if (*rc) { delete c; return nullptr; }
  g_rank_ctx[slot] = c;
  return c;
}

// a += b, launched on device 0's stream (finish.cu)
int launch_add_counts(Ctx* ctx, int32_t* a, const int32_t* b, int64_t elems, cudaStream_t st, const Err& E);

// ------------------------------------------------------------------ one pass, several statistics
static int union_mask(const int32_t* stats, int nstats) {
  int mask = 0;
  for (int k = 0; k < nstats; ++k) mask |= prod_mask_for(stats[k]);
  return mask;
}
static const int kAllProducts = (1 << TPG_MAT_DD) | (1 << TPG_MAT_NN) | (1 << TPG_MAT_HH) | (1 << TPG_MAT_HN);

static uint64_t fnv1a(const void* p, size_t bytes, uint64_t h) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

// Identity of (X, ind.row, ind.col) for the count-matrix cache (tpg_config.cache_counts): file-backed
// sources only — device, inode, size and mtime of the .bk say whether the bytes are still the same.
// Empty = not cacheable.
static std::string counts_cache_key(const tpg_job* job) {
  if (!cfg().cache_counts || job->src_kind != TPG_SRC_FILE || !job->bk_path) return std::string();
  if (job->shard_mode != TPG_SHARD_NONE && job->shard_world > 1) return std::string();
  struct stat sb;
  if (stat(job->bk_path, &sb) != 0) return std::string();
  const int64_t n = job->ind_row ? job->n : job->nrow_total, m = job->ind_col ? job->m : job->ncol_total;
  const uint64_t hr = job->ind_row ? fnv1a(job->ind_row, sizeof(int32_t) * size_t(n), kFnvSeed) : 0;
  const uint64_t hc = job->ind_col ? fnv1a(job->ind_col, sizeof(int32_t) * size_t(m), kFnvSeed) : 0;
  char key[512];
  snprintf(key, sizeof key, "%llu:%llu:%lld:%lld.%09ld|%lldx%lld|r%lld:%016llx|c%lld:%016llx",
           (unsigned long long)sb.st_dev, (unsigned long long)sb.st_ino, (long long)sb.st_size,
           (long long)sb.st_mtim.tv_sec, (long)sb.st_mtim.tv_nsec, (long long)job->nrow_total,
           (long long)job->ncol_total, (long long)(job->ind_row ? n : -1), (unsigned long long)hr,
           (long long)(job->ind_col ? m : -1), (unsigned long long)hc);
  return std::string(key);
}

// Finishes every requested statistic from one set of count matrices on ctx.  results[0].stats
// receives the stage times of the shared pass (the accumulate is collected by the first finish).
static int finish_many(Ctx* c, const tpg_job* job, const int32_t* stats, int nstats, const int32_t* d_counts,
                       tpg_result* results, const Err& E) {
  int rc = TPG_OK;
  for (int k = 0; k < nstats && rc == TPG_OK; ++k) {
    tpg_job jk = *job;
    jk.stat = stats[k];
    rc = finish_impl(c, &jk, d_counts, &results[k], E);
  }
  return rc;
}

// Multi-GPU hand-back of large results: every statistic is finished on the first device INTO device
// buffers, then each device pulls its slab of columns over NVLink and sends it home over ITS OWN PCIe link
// (own pinned bounce buffers; the copies into the caller's pageable matrix go through the shared worker pool).
// One link carries a 20 GB matrix in ~0.4 s at best; eight carry it as fast as the host can take it.
static int finish_many_multilink(const std::vector<Ctx*>& ctxs, const tpg_job* job, const int32_t* stats, int nstats,
                                 const int32_t* d_counts, tpg_result* results, const Err& E) {
  Ctx* c0 = ctxs[0];
  const int ndev = int(ctxs.size());
  const int64_t n = job->ind_row ? job->n : job->nrow_total;
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  for (int k = 0; k < nstats; ++k) {
    tpg_job jk = *job;
    jk.stat = stats[k];
    double* host[4] = {results[k].out0, results[k].out1, results[k].aux0, results[k].aux1};
    tpg_result rd = results[k];
    double* dev0[4] = {nullptr, nullptr, nullptr, nullptr};
    TPG_CUDA_OK(cudaSetDevice(c0->device), E);
    for (int o = 0; o < 4; ++o) {
      if (!host[o]) continue;
      int rc = ensure_slot(&c0->d_out[o], bytes, false, "result matrix", E);
      if (rc) return rc;
      dev0[o] = static_cast<double*>(c0->d_out[o].p);
    }
    rd.out0 = dev0[0]; rd.out1 = dev0[1]; rd.aux0 = dev0[2]; rd.aux1 = dev0[3];
    rd.out_on_device = 1;
    int rc = finish_impl(c0, &jk, d_counts, &rd, E);   // device -> device: no D2H here
    results[k].stats = rd.stats;
    if (rc) return rc;
    const double t0 = now_ms();
    std::vector<int> rcs(ndev, TPG_OK);
    std::vector<std::string> msgs(ndev);
    std::vector<int64_t> moved(ndev, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r) {
      th.emplace_back([&, r] {
        char buf[512] = {0};
        Err Er{buf, sizeof buf};
        Ctx* c = ctxs[r];
        int rcr = TPG_OK;
        const int64_t j0 = n * r / ndev, j1 = n * (r + 1) / ndev;
        const size_t off = size_t(j0) * size_t(n), len = sizeof(double) * size_t(j1 - j0) * size_t(n);
        if (len == 0) return;
        if (cudaSetDevice(c->device) != cudaSuccess) rcr = Er.fail(TPG_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
        for (int o = 0; o < 4 && rcr == TPG_OK; ++o) {
          if (!host[o]) continue;
          const void* src = dev0[o] + off;
          if (r != 0 && c->device != c0->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(c0->device, 0);
            if (e != cudaSuccess) cudaGetLastError();  // already enabled / not possible: the copy still works (staged)
            rcr = ensure_slot(&c->d_out[o], len, false, "result slab", Er);
            if (rcr) break;
            e = cudaMemcpyPeerAsync(c->d_out[o].p, c->device, src, c0->device, len, c->s_compute);
            if (e != cudaSuccess) { rcr = Er.fail(TPG_ERR_CUDA, "peer copy of a result slab failed: %s", cudaGetErrorString(e)); break; }
            src = c->d_out[o].p;
          }
          const void* dv[1] = {src};
          void* hv[1] = {host[o] + off};
          c->ev_plain_used = 0;
          rcr = results_to_host(c, dv, hv, 1, len, &moved[r], Er);
          if (rcr == TPG_OK && cudaStreamSynchronize(c->s_compute) != cudaSuccess) rcr = Er.fail(TPG_ERR_CUDA, "D2H of a result slab failed");
        }
        rcs[r] = rcr;
        msgs[r] = buf;
      });
    }
    for (auto& t : th) t.join();
    TPG_CUDA_OK(cudaSetDevice(c0->device), E);
    for (int r = 0; r < ndev; ++r) {
      if (rcs[r]) return E.fail(rcs[r], "device %d (rank %d): %s", ctxs[r]->device, r, msgs[r].c_str());
      results[k].stats.bytes_d2h += moved[r];
    }
    results[k].stats.ms_d2h += now_ms() - t0;
  }
  return TPG_OK;
}

// In-process multi-GPU job (the R process is one process): one host thread per device
// accumulates its shard (loci slice, or owned output blocks) into its own count matrices;
// the int32 matrices are then summed onto the first device over NVLink (peer copies) and
// finished there once.  Exact for both partitions, because counts are integer sums and the
// tile partition's per-rank matrices are disjoint.
static int pairwise_multi(const tpg_job* job, const int32_t* stats, int nstats, tpg_result* results,
                          const int32_t* devices, int ndev, int shard_mode, const Err& E) {
  if (ndev <= 0 || !devices) return E.fail(TPG_ERR_ARG, "no devices given");
  if (ndev > 64) return E.fail(TPG_ERR_ARG, "at most 64 devices");
  if (ndev > 1 && job->src_kind == TPG_SRC_DEVICE)
    return E.fail(TPG_ERR_ARG, "a device-resident source lives on one GPU; multi-GPU jobs take a file or host source");
  if (nstats <= 0 || !stats || !results) return E.fail(TPG_ERR_ARG, "no statistics requested");
  for (int k = 0; k < nstats; ++k) {
    if (!prod_mask_for(stats[k])) return E.fail(TPG_ERR_ARG, "unknown stat %d", stats[k]);
    if (!results[k].out0) return E.fail(TPG_ERR_ARG, "result / out0 is NULL");
    memset(&results[k].stats, 0, sizeof results[k].stats);
  }
  const int64_t n = job->ind_row ? job->n : job->nrow_total;
  if (n <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row");
  if (shard_mode != TPG_SHARD_LOCI && shard_mode != TPG_SHARD_TILES)
    // Host bytes are the scarce resource of a one-process job: loci sharding reads (and packs) every
    // byte once, tile sharding once per device.  So loci while the four int32 count matrices of a
    // device stay within 64 GB (n <= ~63,000), tiles beyond.
    shard_mode = tpg_counts_elems(n) * int64_t(sizeof(int32_t)) <= (int64_t(64) << 30) ? TPG_SHARD_LOCI : TPG_SHARD_TILES;
  const double t_begin = now_ms();
  const size_t celems = size_t(tpg_counts_elems(n));
  const size_t cbytes = sizeof(int32_t) * celems;
  const std::string key = counts_cache_key(job);
  const int mask = key.empty() ? union_mask(stats, nstats) : kAllProducts;

  std::vector<Ctx*> ctxs(ndev, nullptr);
  for (int r = 0; r < ndev; ++r) {
    int rc = TPG_OK;
    ctxs[r] = rank_ctx(r, devices[r], E, &rc);
    if (!ctxs[r]) return rc;
  }
  Ctx* c0 = ctxs[0];
  tpg_job j0 = *job;
  j0.device = c0->device;
  j0.shard_mode = TPG_SHARD_NONE;
  j0.shard_rank = 0;
  j0.shard_world = 1;
  if (!key.empty() && c0->counts_key == key && c0->counts_key_n == n && c0->d_counts.p) {
    // the count matrices of this (X, ind.row, ind.col) are still on the first device: finish from them
    TPG_CUDA_OK(cudaSetDevice(c0->device), E);
    int rc = finish_many(c0, &j0, stats, nstats, static_cast<int32_t*>(c0->d_counts.p), results, E);
    results[0].stats.n_cached_jobs = 1;
    results[0].stats.ms_total = now_ms() - t_begin;
    return rc;
  }
  c0->counts_key.clear();
  std::vector<int> rcs(ndev, TPG_OK);
  std::vector<std::string> msgs(ndev);
  std::vector<tpg_stats> sts(ndev);
  std::vector<std::thread> th;
  for (int r = 0; r < ndev; ++r) {
    th.emplace_back([&, r] {
      char buf[512] = {0};
      Err Er{buf, sizeof buf};
      Ctx* c = ctxs[r];
      memset(&sts[r], 0, sizeof(tpg_stats));
      tpg_job j = *job;
      j.stat = stats[0];
      j.device = devices[r];
      j.shard_mode = ndev > 1 ? shard_mode : TPG_SHARD_NONE;
      j.shard_rank = r;
      j.shard_world = ndev;
      int rc = TPG_OK;
      if (cudaSetDevice(c->device) != cudaSuccess) rc = Er.fail(TPG_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
      if (rc == TPG_OK) rc = ensure_slot(&c->d_counts, cbytes, false, "count matrices", Er);
      // loci-sharded: every device adds into all blocks; tile-sharded: its own blocks only — but the tree
      // reduction below sums WHOLE matrices, so everything a peer may read has to be zero
      if (rc == TPG_OK) rc = launch_zero_counts(c, static_cast<int32_t*>(c->d_counts.p), n, mask, 0, 1, nullptr, c->s_compute, Er);
      if (rc == TPG_OK) {
        rc = accumulate_impl(c, &j, static_cast<int32_t*>(c->d_counts.p), &sts[r], Er, false, nullptr, mask);
      }
      rcs[r] = rc;
      msgs[r] = buf;
    });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < ndev; ++r)
    if (rcs[r]) return E.fail(rcs[r], "device %d (rank %d): %s", devices[r], r, msgs[r].c_str());

  // ---- sum the count matrices onto rank 0's device: a pairwise tree over NVLink.  In step s the
  // device of rank r (r % 2s == 0) adds the matrices of rank r + s, reading them straight out of the
  // peer's HBM inside the add kernel (peer access; no staging copy) — the pairs of a step run side
  // by side on their own devices, so 8 ranks need 3 steps instead of 7 serial copy + add rounds.
  // Without peer access between a pair (e.g. the same device listed twice) the pair falls back to a
  // copy through a staging buffer.
  const int64_t ld = tpg_counts_ld(n), ms = ld * ld;
  const double t_red = now_ms();
  for (int step = 1; step < ndev; step <<= 1) {
    std::vector<int> receivers;
    for (int r = 0; r + step < ndev; r += 2 * step) receivers.push_back(r);
    for (int r : receivers) {
      Ctx* cr = ctxs[r];
      Ctx* cp = ctxs[r + step];
      TPG_CUDA_OK(cudaSetDevice(cr->device), E);
      bool direct = cr->device != cp->device;
      if (direct) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, cr->device, cp->device) != cudaSuccess || !can) direct = false;
        if (direct) {
          cudaError_t e = cudaDeviceEnablePeerAccess(cp->device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) direct = false;
          cudaGetLastError();
        }
      }
      if (cr->device == cp->device) direct = true;  // same device: the "peer" pointer is local
      int32_t* tmp = nullptr;
      if (!direct) {
        int rc = ensure_slot(&cr->d_out[3], sizeof(int32_t) * size_t(ms), false, "reduction staging", E);
        if (rc) return rc;
        tmp = static_cast<int32_t*>(cr->d_out[3].p);
      }
      if (direct) {
        // one launch: the blocks the job accumulated, read straight out of the peer's HBM
        int rc = launch_add_counts_blocks(cr, static_cast<int32_t*>(cr->d_counts.p), static_cast<const int32_t*>(cp->d_counts.p),
                                          n, mask, cr->s_compute, E);
        if (rc) return rc;
      } else {
        for (int k = 0; k < TPG_NUM_COUNT_MATS; ++k) {
          if (!(mask & (1 << k))) continue;
          int32_t* mine = static_cast<int32_t*>(cr->d_counts.p) + k * ms;
          const int32_t* theirs = static_cast<const int32_t*>(cp->d_counts.p) + k * ms;
          TPG_CUDA_OK(cudaMemcpyPeerAsync(tmp, cr->device, theirs, cp->device, sizeof(int32_t) * size_t(ms), cr->s_compute), E);
          int rc = launch_add_counts(cr, mine, tmp, ms, cr->s_compute, E);
          if (rc) return rc;
        }
      }
    }
    for (int r : receivers) {  // the step is done when every pair is
      TPG_CUDA_OK(cudaSetDevice(ctxs[r]->device), E);
      TPG_CUDA_OK(cudaStreamSynchronize(ctxs[r]->s_compute), E);
    }
  }
  TPG_CUDA_OK(cudaSetDevice(c0->device), E);
  const double ms_reduce = now_ms() - t_red;
  // large host-bound results go home over every device's PCIe link, small ones (or device-resident ones) from the first
  bool multilink = ndev > 1 && sizeof(double) * size_t(n) * size_t(n) >= (size_t(256) << 20);
  for (int k = 0; k < nstats; ++k) multilink = multilink && !results[k].out_on_device;
  int rc = multilink ? finish_many_multilink(ctxs, &j0, stats, nstats, static_cast<int32_t*>(c0->d_counts.p), results, E)
                     : finish_many(c0, &j0, stats, nstats, static_cast<int32_t*>(c0->d_counts.p), results, E);
  if (rc == TPG_OK && !key.empty()) { c0->counts_key = key; c0->counts_key_n = n; }
  // stats: stage times are the max over devices (they ran side by side); bytes and MACs add up
  tpg_stats& S = results[0].stats;
  for (int r = 0; r < ndev; ++r) {
    S.ms_read = std::max(S.ms_read, sts[r].ms_read);
    S.ms_h2d = std::max(S.ms_h2d, sts[r].ms_h2d);
    S.ms_decode = std::max(S.ms_decode, sts[r].ms_decode);
    S.ms_gram = std::max(S.ms_gram, sts[r].ms_gram);
    S.max_chain_loci = std::max(S.max_chain_loci, sts[r].max_chain_loci);
    S.bytes_h2d += sts[r].bytes_h2d;
    S.bytes_file += sts[r].bytes_file;
    S.n_chunks += sts[r].n_chunks;
    S.n_packed_chunks += sts[r].n_packed_chunks;
    S.n_mxf4_chunks += sts[r].n_mxf4_chunks;
    S.n_i8_chunks += sts[r].n_i8_chunks;
    S.n_wide_launches += sts[r].n_wide_launches;
    S.n_direct_io += sts[r].n_direct_io;
    S.mma_macs += sts[r].mma_macs;
    S.n_launches += sts[r].n_launches;
  }
  S.ms_reduce = ms_reduce;
  S.ms_total = now_ms() - t_begin;
  return rc;
}

// Caller-visible result matrices that live in files (tpg_result.out0_path / out1_path): mapped
// read-write for the duration of a job, synced and unmapped afterwards.
struct ResultMaps {
  struct M { void* p = nullptr; size_t bytes = 0; };
  std::vector<M> maps;
  ~ResultMaps() {
    for (M& m : maps)
      if (m.p) { msync(m.p, m.bytes, MS_SYNC); munmap(m.p, m.bytes); }
  }
  int map_one(const char* path, size_t bytes, double** out, const Err& E) {
    const int fd = open(path, O_RDWR);
    if (fd < 0) return E.fail(TPG_ERR_IO, "cannot open result backing file '%s': %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fd, &sb) != 0 || size_t(sb.st_size) < bytes) {
      close(fd);
      return E.fail(TPG_ERR_ARG, "result backing file '%s' is smaller than the n x n doubles it must hold (%zu bytes)", path, bytes);
    }
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return E.fail(TPG_ERR_IO, "mmap('%s') failed: %s", path, strerror(errno));
    maps.push_back(M{p, bytes});
    *out = static_cast<double*>(p);
    return TPG_OK;
  }
  // Fills out0 / out1 of a working copy of the caller's results from their paths.
  int apply(const tpg_job* job, tpg_result* res, int nres, const Err& E) {
    const int64_t n = job->ind_row ? job->n : job->nrow_total;
    if (n <= 0) return TPG_OK;  // reported by the job itself
    const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
    for (int k = 0; k < nres; ++k) {
      if ((res[k].out0_path || res[k].out1_path) && res[k].out_on_device)
        return E.fail(TPG_ERR_ARG, "file-backed results cannot be device pointers");
      if (!res[k].out0 && res[k].out0_path) { int rc = map_one(res[k].out0_path, bytes, &res[k].out0, E); if (rc) return rc; }
      if (!res[k].out1 && res[k].out1_path) { int rc = map_one(res[k].out1_path, bytes, &res[k].out1, E); if (rc) return rc; }
    }
    return TPG_OK;
  }
};

// ------------------------------------------------------------------ result shells
// A large result takes as long to come home as a third of the products take (n = 50,000: 20 GB, 0.4 s over one PCIe
// link against 0.75 s of Gram kernels), and a chunk-streamed job can finish nothing before its last chunk.  For a
// whole-file job the matrix is therefore produced in nested shells — a top-left square, then L-shaped rims
// (build_shell_table; tile sharding's machinery with the shells as "ranks" of one GPU):
//   phase A   the file is streamed ONCE: every chunk adds the products of shell 0 and leaves its 2-bit packed image
//             in HBM (the genotype cache's fill path) — as much Gram work as the host needs to read and pack the file;
//   phase B   shells 1 .. K-1 run from the resident image (gather + decode + one Gram launch over all loci each);
//   each shell is finished as soon as its products are done, and its D2H runs on the copy stream while the next
//   shell's kernels run.
// Exposed at the end: the last shell's hand-back only.

// D2H of a rectangle [r0, r1) x [c0, c1) of column-major n x n double matrices (device `dev`, host `host`) on
// `stream`: through the two pinned bounce buffers, columns copied into the caller's pageable matrix by the pool.
static int rect_to_host(Ctx* ctx, cudaStream_t stream, const double* const* dev, double* const* host, int count, int64_t n,
                        int64_t r0, int64_t r1, int64_t c0, int64_t c1, int64_t* bytes_d2h, const Err& E) {
  if (r1 <= r0 || c1 <= c0) return TPG_OK;
  const size_t kPiece = size_t(128) << 20;
  const size_t w = size_t(r1 - r0) * sizeof(double);            // bytes per column of the rectangle
  const int64_t cols_per_piece = std::max<int64_t>(1, int64_t(kPiece / w));
  int rc = TPG_OK;
  for (int b = 0; b < 2 && rc == TPG_OK; ++b)
    rc = ensure_slot(&ctx->h_out[b], std::max(kPiece, w), true, "pinned result staging", E);
  if (rc) return rc;
  struct Piece { double* dst; const uint8_t* from; int64_t ncols; cudaEvent_t ev; };
  Piece prev{nullptr, nullptr, 0, nullptr};
  auto drain = [&](const Piece& q) {
    if (!q.dst) return cudaSuccess;
    cudaError_t e = cudaEventSynchronize(q.ev);
    if (e != cudaSuccess) return e;
    double* dst = q.dst;
    const uint8_t* from = q.from;
    const int64_t grain = std::max<int64_t>(1, int64_t((size_t(256) << 10) / w));
    parallel_for(q.ncols, grain, [=](int64_t a, int64_t z) {
      for (int64_t j = a; j < z; ++j) copy_streaming(dst + j * n, from + size_t(j) * w, w);
    });
    return cudaSuccess;
  };
  cudaEvent_t evp[2];
  for (int b = 0; b < 2; ++b) TPG_CUDA_OK(cudaEventCreateWithFlags(&evp[b], cudaEventDisableTiming), E);
  int64_t idx = 0;
  cudaError_t e = cudaSuccess;
  for (int o = 0; o < count && e == cudaSuccess; ++o) {
    if (!host[o]) continue;
    for (int64_t c = c0; c < c1 && e == cudaSuccess; c += cols_per_piece, ++idx) {
      const int64_t nc = std::min(cols_per_piece, c1 - c);
      const int b = int(idx & 1);
      // bounce buffer b was last used two pieces ago: that piece has been drained (it was `prev` one step back)
      e = cudaMemcpy2DAsync(ctx->h_out[b].p, w, dev[o] + c * n + r0, size_t(n) * sizeof(double), w, size_t(nc),
                            cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaEventRecord(evp[b], stream);
      if (e == cudaSuccess) e = drain(prev);
      prev = Piece{host[o] + c * n + r0, static_cast<const uint8_t*>(ctx->h_out[b].p), nc, evp[b]};
      *bytes_d2h += int64_t(w) * nc;
    }
  }
  if (e == cudaSuccess) e = drain(prev);
  for (int b = 0; b < 2; ++b) cudaEventDestroy(evp[b]);
  if (e != cudaSuccess) return E.fail(TPG_ERR_CUDA, "result D2H failed: %s", cudaGetErrorString(e));
  return TPG_OK;
}

// Shells need the 3840-wide ownership regions (shard_region_blocks): as many as they allow, or exactly `asked`.
static int shell_count(int64_t nblocks, int asked) {
  for (int cand : {6, 4, 3}) {
    if (asked >= 2) cand = asked;
    if (shard_region_blocks(nblocks, cand) == 15 && (nblocks + 14) / 15 >= cand) return cand;
    if (asked >= 2) break;
  }
  return 0;
}

// How many shells this job's result comes home in (0 = the ordinary path).
static int shells_for(Ctx* c, const tpg_job* job, int stat, const tpg_result* res) {
  const tpg_config C = cfg();
  if (C.result_shells == 1) return 0;
  if (job->src_kind != TPG_SRC_FILE || !job->bk_path || job->ind_row || job->ind_col) return 0;
  if (job->shard_mode != TPG_SHARD_NONE && job->shard_world > 1) return 0;
  if (stat == TPG_STAT_GRM || res->out_on_device) return 0;               // GRM's mean needs the whole matrix
  if (job->flags & (TPG_FLAG_PLANES_K_MAJOR | TPG_FLAG_NO_PACK | TPG_FLAG_GRAM_1CTA | TPG_FLAG_NO_OVERLAP)) return 0;
  if (C.pack_mode == 1 || C.no_compact || C.cache_counts || C.gram_ctas == 1) return 0;
  const int64_t n = job->nrow_total, nblocks = round_up(n, kRowPad) / 256;
  const int K = shell_count(nblocks, C.result_shells);
  if (!K) return 0;
  if (!gather_pack_fits((n + 3) / 4)) return 0;
  // room for the packed image beside everything a phase-B shell holds at once
  size_t free_b = 0, total_b = 0;
  if (cudaSetDevice(c->device) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
  const size_t held = c->d_geno.cap + c->d_counts.cap + c->d_out[0].cap + c->d_out[1].cap + c->d_out[2].cap + c->d_out[3].cap +
                      c->d_planes[0].cap + c->d_planes[1].cap + c->d_gat[0].cap + c->d_gat[1].cap + c->d_raw[0].cap + c->d_raw[1].cap;
  const int nout = 1 + (res->out1 ? 1 : 0) + (res->aux0 ? 1 : 0) + (res->aux1 ? 1 : 0);
  const double n_pad = double(round_up(n, kRowPad)), m = double(job->ncol_total);
  const double need = 16.0 * n_pad * n_pad + 8.0 * double(n) * double(n) * nout + double(pk2_stride(n)) * m * 2.0 +
                      1.5 * n_pad * m + 2e9;   // counts, results, image + one gathered copy, 4-bit planes, staging
  if (need > 0.92 * double(free_b + held)) return 0;
  return K;
}

static int pairwise_shells(Ctx* c, const tpg_job* job, int stat, tpg_result* res, int K, const Err& E) {
  const double t_begin = now_ms();
  const int64_t l0 = c->launches;
  const int64_t n = job->nrow_total, n_pad = round_up(n, kRowPad);
  tpg_stats& st = res->stats;
  int rc = ensure_slot(&c->d_counts, sizeof(int32_t) * size_t(tpg_counts_elems(n)), false, "count matrices", E);
  if (rc) return rc;
  int32_t* d_counts = static_cast<int32_t*>(c->d_counts.p);
  double* host[4] = {res->out0, res->out1, res->aux0, res->aux1};
  const double* dev[4] = {nullptr, nullptr, nullptr, nullptr};
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  for (int o = 0; o < 4; ++o) {
    if (!host[o]) continue;
    rc = ensure_slot(&c->d_out[o], bytes, false, "result matrix", E);
    if (rc) return rc;
    dev[o] = static_cast<const double*>(c->d_out[o].p);
    if (bytes >= (size_t(8) << 20) && !cfg().no_thp) {   // a fresh result matrix: ask for huge pages (results_to_host)
      const uintptr_t lo = (reinterpret_cast<uintptr_t>(host[o]) + (size_t(2) << 20) - 1) & ~((uintptr_t(2) << 20) - 1);
      const uintptr_t hi = (reinterpret_cast<uintptr_t>(host[o]) + bytes) & ~((uintptr_t(2) << 20) - 1);
      if (hi > lo) madvise(reinterpret_cast<void*>(lo), size_t(hi - lo), MADV_HUGEPAGE);
    }
  }
  struct Mode {   // the shell table replaces the weighted owner table while this job runs
    Ctx* c;
    ~Mode() {
      cudaStreamSynchronize(c->s_compute); cudaStreamSynchronize(c->s_decode); cudaStreamSynchronize(c->s_copy);
      c->owner_shells = 0; c->keep_events = false; c->owner_key.clear(); c->pending = false;
      if (!cfg().cache_genotypes) c->geno_key.clear();
    }
  } mode{c};
  c->owner_shells = K;
  c->keep_events = true;
  c->counts_key.clear();
  c->ev_timing_used = c->ev_plain_used = 0;
  const uint8_t *owner_h = nullptr, *owner_d = nullptr;
  rc = owner_table(c, n_pad, K, &owner_h, &owner_d, E);
  if (rc) return rc;
  if (!owner_d || int(c->shell_bounds.size()) != K + 1) return E.fail(TPG_ERR_ARG, "result shells: no shell table for this shape");
  const int mask = prod_mask_for(stat);
  rc = launch_zero_counts(c, d_counts, n, mask, 0, 1, nullptr, c->s_compute, E);   // every shell's blocks, once
  if (rc) return rc;
  tpg_job jk = *job;
  jk.stat = stat;
  jk.shard_mode = TPG_SHARD_TILES;
  jk.shard_world = K;
  FinishArgs fa;
  fa.stat = stat;
  fa.counts = d_counts;
  fa.ld = n_pad;
  fa.n = n;
  fa.m = job->ncol_total;
  fa.n_short = 0;
  if (stat == TPG_STAT_AS && job->pad_quirk)
    fa.n_short = as_pad_count(job->ncol_total, job->block_size > 0 ? job->block_size : block_size_default(job->nrow_total));
  st.n_short = fa.n_short;
  fa.shard_world = K;
  fa.owner = owner_d;
  fa.out0 = const_cast<double*>(dev[0]); fa.out1 = const_cast<double*>(dev[1]);
  fa.aux0 = const_cast<double*>(dev[2]); fa.aux1 = const_cast<double*>(dev[3]);
  std::vector<std::vector<Ctx::ChunkEv>> evs(static_cast<size_t>(K));
  std::vector<cudaEvent_t> ev_fin(static_cast<size_t>(K)), ev_f0(static_cast<size_t>(K));
  const std::string gkey = geno_cache_key(job, true);
  for (int b = 0; b < 2; ++b) {           // the bounce buffers, before the hand-back thread needs them
    rc = ensure_slot(&c->h_out[b], size_t(128) << 20, true, "pinned result staging", E);
    if (rc) return rc;
  }
  // Hand-back of one shell at a time on a helper thread (DMA on the copy stream, the pool's copies into the caller's
  // matrix), so that this thread can enqueue the next shells meanwhile; phase B stages nothing, the pool is free.
  struct HandBack {
    std::thread th;
    int rc = TPG_OK;
    char err[512] = {0};
    int64_t bytes = 0;
    double ms = 0;
  } hb;
  auto join_hand_back = [&]() -> int {
    if (hb.th.joinable()) hb.th.join();
    if (hb.rc != TPG_OK) { const int r = E.fail(hb.rc, "%s", hb.err); hb.rc = TPG_OK; return r; }
    return TPG_OK;
  };
  struct Joiner { HandBack* h; ~Joiner() { if (h->th.joinable()) h->th.join(); } } joiner{&hb};
  auto start_hand_back = [&](int k) -> int {   // shell k: columns [ca, cb) down to row cb, and rows [ca, cb) of the columns before
    int r = join_hand_back();
    if (r) return r;
    const int64_t ca = std::min(n, c->shell_bounds[size_t(k)] * 15 * 256), cb = std::min(n, c->shell_bounds[size_t(k) + 1] * 15 * 256);
    const cudaEvent_t ready = ev_fin[size_t(k)];
    auto body = [c, &hb, dev, host, n, ca, cb, ready]() {
      Err Et{hb.err, sizeof hb.err};
      const double t0 = now_ms();
      if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamWaitEvent(c->s_copy, ready, 0) != cudaSuccess) {
        hb.rc = Et.fail(TPG_ERR_CUDA, "result shells: the copy stream could not wait for the shell");
        return;
      }
      int r2 = rect_to_host(c, c->s_copy, dev, host, 4, n, 0, cb, ca, cb, &hb.bytes, Et);
      if (r2 == TPG_OK) r2 = rect_to_host(c, c->s_copy, dev, host, 4, n, ca, cb, 0, ca, &hb.bytes, Et);
      hb.rc = r2;
      hb.ms += now_ms() - t0;
    };
    try {
      hb.th = std::thread(body);
    } catch (const std::exception&) {   // no thread to be had: hand back on this one
      body();
    }
    return TPG_OK;
  };
  for (int k = 0; k < K && rc == TPG_OK; ++k) {
    jk.shard_rank = k;
    tpg_stats sk;
    memset(&sk, 0, sizeof sk);
    // shell 0 streams the file and must leave the packed image behind; the others read it (and, from the second
    // one on, the planes the first of them decoded)
    rc = accumulate_impl(c, &jk, d_counts, &sk, E, /*defer=*/true, nullptr, 0, nullptr, 0, /*force_geno=*/true);
    if (rc) break;
    if (k == 0 && c->geno_key != gkey) { rc = E.fail(TPG_ERR_NOMEM, "result shells: the packed genotype image did not fit"); break; }
    if (k > 0 && !sk.n_geno_cached) { rc = E.fail(TPG_ERR_CUDA, "result shells: shell %d did not run from the packed image", k); break; }
    evs[size_t(k)] = c->pend_ev;
    c->pending = false;                   // collected below, after the last shell
    st.ms_read += sk.ms_read; st.n_chunks += sk.n_chunks; st.n_packed_chunks += sk.n_packed_chunks; st.bytes_h2d += sk.bytes_h2d;
    st.n_mxf4_chunks += sk.n_mxf4_chunks; st.n_i8_chunks += sk.n_i8_chunks; st.n_wide_launches += sk.n_wide_launches;
    st.max_chain_loci = std::max(st.max_chain_loci, sk.max_chain_loci); st.bytes_file += sk.bytes_file;
    st.n_direct_io += sk.n_direct_io; st.mma_macs += sk.mma_macs;
    fa.shard_rank = k;
    ev_f0[size_t(k)] = pooled_event(c, true);
    ev_fin[size_t(k)] = pooled_event(c, true);
    TPG_CUDA_OK(cudaEventRecord(ev_f0[size_t(k)], c->s_compute), E);
    rc = launch_finish(c, fa, c->s_compute, E);
    if (rc) break;
    TPG_CUDA_OK(cudaEventRecord(ev_fin[size_t(k)], c->s_compute), E);
    if (k > 0) rc = start_hand_back(k - 1);   // the previous shell comes home while this one's kernels run
  }
  if (rc == TPG_OK) rc = start_hand_back(K - 1);
  { const int r = join_hand_back(); if (rc == TPG_OK) rc = r; }
  st.bytes_d2h += hb.bytes;
  const double ms_host_d2h = hb.ms;
  cudaError_t e = cudaStreamSynchronize(c->s_compute);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_decode);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->s_copy);
  if (rc == TPG_OK) rc = check_err_flag(c, E);
  if (rc == TPG_OK && e != cudaSuccess) rc = E.fail(TPG_ERR_CUDA, "result shells failed on device: %s", cudaGetErrorString(e));
  if (rc == TPG_OK) {
    for (int k = 0; k < K; ++k) {
      float ms = 0;
      for (const Ctx::ChunkEv& t : evs[size_t(k)]) {
        cudaEventElapsedTime(&ms, t.c0, t.c1); st.ms_h2d += ms;
        cudaEventElapsedTime(&ms, t.d0, t.d1); st.ms_decode += ms;
        cudaEventElapsedTime(&ms, t.g0, t.g1); st.ms_gram += ms;
      }
      cudaEventElapsedTime(&ms, ev_f0[size_t(k)], ev_fin[size_t(k)]); st.ms_finish += ms;
    }
    st.ms_d2h = ms_host_d2h;              // host time spent in the hand-back loops (most of it beside the next shell's kernels)
    st.n_result_shells = K;
  }
  st.n_launches = c->launches - l0;
  st.ms_total = now_ms() - t_begin;
  return rc;
}

static int pairwise_many_on(Ctx* c, const tpg_job* job, const int32_t* stats, int nstats, tpg_result* results,
                            const Err& E, const int32_t* group_sorted = nullptr, uint64_t group_sig = 0) {
  if (job->shard_mode == TPG_SHARD_LOCI && job->shard_world > 1)
    return E.fail(TPG_ERR_ARG, "loci-sharded jobs need the staged interface (tpg_accumulate, "
                               "all-reduce of the counts, tpg_finish)");
  if (nstats <= 0 || !stats || !results) return E.fail(TPG_ERR_ARG, "no statistics requested");
  const int64_t l0 = c->launches;
  const double t_begin = now_ms();
  const int64_t n = job->ind_row ? job->n : job->nrow_total;
  if (n <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row");
  for (int k = 0; k < nstats; ++k) {
    if (!prod_mask_for(stats[k])) return E.fail(TPG_ERR_ARG, "unknown stat %d", stats[k]);
    if (!results[k].out0) return E.fail(TPG_ERR_ARG, "result / out0 is NULL");
    memset(&results[k].stats, 0, sizeof results[k].stats);
  }
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  if (nstats == 1 && !group_sorted) {
    if (const int K = shells_for(c, job, stats[0], &results[0])) return pairwise_shells(c, job, stats[0], &results[0], K, E);
  }
  const size_t cbytes = sizeof(int32_t) * size_t(tpg_counts_elems(n));
  const std::string key = group_sorted ? std::string() : counts_cache_key(job);
  int rc = TPG_OK;
  if (!key.empty() && c->counts_key == key && c->counts_key_n == n && c->d_counts.p && c->d_counts.cap >= cbytes) {
    // same .bk (size, mtime), same ind.row / ind.col as the job that left its four count matrices
    // here: nothing to read, decode or multiply
    rc = finish_many(c, job, stats, nstats, static_cast<int32_t*>(c->d_counts.p), results, E);
    results[0].stats.n_cached_jobs = 1;
  } else {
    c->counts_key.clear();
    rc = ensure_slot(&c->d_counts, cbytes, false, "count matrices", E);
    if (rc) return rc;
    int32_t* d_counts = static_cast<int32_t*>(c->d_counts.p);
    const int mask = !key.empty() ? kAllProducts : union_mask(stats, nstats);
    // zero -> decode -> gram -> finish are enqueued back to back; one sync at the very end
    const bool tiles = job->shard_mode == TPG_SHARD_TILES && job->shard_world > 1;
    const uint8_t *owner_h = nullptr, *owner_d = nullptr;
    if (tiles) { rc = owner_table(c, round_up(n, kRowPad), job->shard_world, &owner_h, &owner_d, E); if (rc) return rc; }
    rc = launch_zero_counts(c, d_counts, n, mask, tiles ? job->shard_rank : 0, tiles ? job->shard_world : 1, owner_d, c->s_compute, E);
    if (rc) return rc;
    tpg_job jw = *job;
    jw.stat = stats[0];  // (tpg_pairwise_many ignores job->stat; the plan only needs a valid one)
    rc = accumulate_impl(c, &jw, d_counts, &results[0].stats, E, /*defer=*/true, nullptr, mask, group_sorted, group_sig);
    if (rc == TPG_OK) rc = finish_many(c, job, stats, nstats, d_counts, results, E);
    if (c->pending) {  // finish bailed out early: still drain the device work
      tpg_stats scratch;
      memset(&scratch, 0, sizeof scratch);
      Err quiet{nullptr, 0};
      collect_accumulate(c, &scratch, quiet);
    }
    if (rc == TPG_OK && !key.empty()) { c->counts_key = key; c->counts_key_n = n; }
  }
  results[0].stats.n_launches = c->launches - l0;
  results[0].stats.ms_total = now_ms() - t_begin;
  return rc;
}

// Runs `body` on working copies of the caller's results whose file-backed outputs are mapped, then
// hands the stats back.
template <class F>
static int with_result_maps(const tpg_job* job, tpg_result* results, int nres, const Err& E, F&& body) {
  std::vector<tpg_result> work(results, results + nres);
  ResultMaps maps;
  int rc = maps.apply(job, work.data(), nres, E);
  if (rc == TPG_OK) rc = body(work.data());
  for (int k = 0; k < nres; ++k) results[k].stats = work[k].stats;
  return rc;
}

}  // namespace tpg

using namespace tpg;

extern "C" {

const char* tpg_version(void) { return "tidypopgen_b200 0.2.0 (sm_100a, abi 2)"; }

int tpg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int64_t tpg_block_size(int64_t n) { return block_size_default(n); }

int64_t tpg_split_len(int64_t m, int64_t block_size, int64_t* lower, int64_t* upper, int64_t* size,
                      int64_t cap) {
  return split_len(m, block_size, lower, upper, size, cap);
}

int64_t tpg_as_pad_count(int64_t m, int64_t block_size) { return as_pad_count(m, block_size); }

int32_t tpg_result_shell_bounds(int64_t n, int32_t shells, int64_t* bounds) {
  if (n <= 0 || !bounds || shells == 1 || shells < 0 || shells > 8) return 0;
  const int64_t nblocks = tpg::round_up(n, tpg::kRowPad) / 256;
  const int K = tpg::shell_count(nblocks, shells);
  if (!K) return 0;
  std::vector<uint8_t> table;
  std::vector<int64_t> b;
  tpg::build_shell_table(nblocks, K, tpg::kShellFirstFrac, &table, &b);
  if (int(b.size()) != K + 1) return 0;
  for (int k = 0; k <= K; ++k) bounds[k] = std::min<int64_t>(n, b[size_t(k)] * 15 * 256);
  return K;
}

int32_t tpg_tile_owner(int64_t n, int64_t bi, int64_t bj, int32_t world) {
  if (n <= 0 || world <= 1) return 0;
  const int64_t nblocks = round_up(n, kRowPad) / 256;
  if (bi > bj) std::swap(bi, bj);
  if (bi < 0 || bj >= nblocks) return -1;
  static std::mutex mu;
  static std::vector<uint8_t> table;
  static int64_t t_nb = -1;
  static int t_world = -1;
  std::lock_guard<std::mutex> lk(mu);
  if (t_nb != nblocks || t_world != world) { build_owner_table(nblocks, world, &table); t_nb = nblocks; t_world = world; }
  for (int r = 0; r < world; ++r)
    if (block_owned(bi, bj, nblocks, r, world, table.empty() ? nullptr : table.data())) return r;
  return -1;
}

int64_t tpg_counts_ld(int64_t n) { return round_up(n, kRowPad); }
int64_t tpg_counts_elems(int64_t n) {
  const int64_t ld = round_up(n, kRowPad);
  return int64_t(TPG_NUM_COUNT_MATS) * ld * ld;
}

tpg_ctx* tpg_ctx_create(int device, char* err, size_t errlen) {
  Err E{err, errlen};
  Ctx* c = new Ctx();
  if (ctx_init(c, device, E) != TPG_OK) { delete c; return nullptr; }
  return reinterpret_cast<tpg_ctx*>(c);
}

void tpg_ctx_destroy(tpg_ctx* h) {
  Ctx* c = reinterpret_cast<Ctx*>(h);
  if (!c) return;
  trim_ctx(c);
  for (cudaEvent_t e : c->ev_timing) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_plain) cudaEventDestroy(e);
  if (c->s_compute) cudaStreamDestroy(c->s_compute);
  if (c->s_decode) cudaStreamDestroy(c->s_decode);
  if (c->s_copy) cudaStreamDestroy(c->s_copy);
  if (c->h_err_flag) cudaFreeHost(c->h_err_flag);
  if (c->d_items) cudaFree(c->d_items);
  delete c;
}

void tpg_ctx_trim(tpg_ctx* h) {
  if (h) trim_ctx(reinterpret_cast<Ctx*>(h));
}

void* tpg_ctx_stream(tpg_ctx* h) {
  return h ? reinterpret_cast<void*>(reinterpret_cast<Ctx*>(h)->s_compute) : nullptr;
}

void tpg_release(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (device >= 0 && device < 64 && g_ctx[device]) trim_ctx(g_ctx[device]);
  // ... and the per-rank contexts tpg_pairwise_multi / tpg_pairwise_many keep on that device
  for (Ctx* c : g_rank_ctx)
    if (c && c->device == device) trim_ctx(c);
}

int tpg_accumulate(tpg_ctx* h, const tpg_job* job, int32_t* d_counts, tpg_stats* stats, char* err,
                   size_t errlen) {
  Err E{err, errlen};
  if (!h) return E.fail(TPG_ERR_ARG, "ctx is NULL");
  if (!d_counts) return E.fail(TPG_ERR_ARG, "d_counts is NULL");
  tpg_stats local;
  memset(&local, 0, sizeof local);
  tpg_stats* st = stats ? stats : &local;
  Ctx* c = reinterpret_cast<Ctx*>(h);
  const int64_t l0 = c->launches;
  // TPG_FLAG_ASYNC: enqueue only — the device work is collected (and its stage times folded into
  // the result's stats) by the tpg_finish that follows on the same context
  int rc = accumulate_impl(c, job, d_counts, st, E, /*defer=*/job && (job->flags & TPG_FLAG_ASYNC));
  st->n_launches += c->launches - l0;
  return rc;
}

int tpg_finish(tpg_ctx* h, const tpg_job* job, const int32_t* d_counts, tpg_result* res, char* err,
               size_t errlen) {
  Err E{err, errlen};
  if (!h) return E.fail(TPG_ERR_ARG, "ctx is NULL");
  if (!job) return E.fail(TPG_ERR_ARG, "job is NULL");
  if (!d_counts) return E.fail(TPG_ERR_ARG, "d_counts is NULL");
  Ctx* c = reinterpret_cast<Ctx*>(h);
  const int64_t l0 = c->launches;
  int rc = finish_impl(c, job, d_counts, res, E);
  if (res) res->stats.n_launches = c->launches - l0;
  return rc;
}

int tpg_pairwise(const tpg_job* job, tpg_result* res, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!job || !res) return E.fail(TPG_ERR_ARG, "job / result is NULL");
  int rc = TPG_OK;
  Ctx* c = cached_ctx(job->device, E, &rc);
  if (!c) return rc;
  return with_result_maps(job, res, 1, E, [&](tpg_result* w) { return pairwise_many_on(c, job, &job->stat, 1, w, E); });
}

int tpg_pairwise_ctx(tpg_ctx* h, const tpg_job* job, tpg_result* res, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!h || !job || !res) return E.fail(TPG_ERR_ARG, "ctx / job / result is NULL");
  Ctx* c = reinterpret_cast<Ctx*>(h);
  return with_result_maps(job, res, 1, E, [&](tpg_result* w) { return pairwise_many_on(c, job, &job->stat, 1, w, E); });
}

int tpg_pairwise_multi(const tpg_job* job, tpg_result* res, const int32_t* devices, int32_t ndev,
                       int32_t shard_mode, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!job || !res) return E.fail(TPG_ERR_ARG, "job / result is NULL");
  return with_result_maps(job, res, 1, E, [&](tpg_result* w) {
    return pairwise_multi(job, &job->stat, 1, w, devices, ndev, shard_mode, E);
  });
}

int tpg_pairwise_many(const tpg_job* job, const int32_t* stats, int32_t nstats, tpg_result* results,
                      const int32_t* devices, int32_t ndev, int32_t shard_mode, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!job || !stats || !results || nstats <= 0) return E.fail(TPG_ERR_ARG, "job / stats / results is NULL");
  if (nstats > 16) return E.fail(TPG_ERR_ARG, "at most 16 statistics per pass");
  if (devices && ndev > 1)
    return with_result_maps(job, results, nstats, E, [&](tpg_result* w) {
      return pairwise_multi(job, stats, nstats, w, devices, ndev, shard_mode, E);
    });
  int rc = TPG_OK;
  Ctx* c = cached_ctx((devices && ndev == 1) ? devices[0] : job->device, E, &rc);
  if (!c) return rc;
  return with_result_maps(job, results, nstats, E, [&](tpg_result* w) { return pairwise_many_on(c, job, stats, nstats, w, E); });
}

int tpg_grm_from_as(int device, const double* as_in, int64_t n, double* grm_out, double* mb_out,
                    char* err, size_t errlen) {
  Err E{err, errlen};
  if (!as_in || !grm_out || n <= 0) return E.fail(TPG_ERR_ARG, "bad arguments");
  int rc = TPG_OK;
  Ctx* c = cached_ctx(device, E, &rc);
  if (!c) return rc;
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  double* d = nullptr;
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  if (cudaMalloc(&d, bytes) != cudaSuccess) return E.fail(TPG_ERR_NOMEM, "cannot allocate %zu bytes", bytes);
  cudaError_t e = cudaMemcpyAsync(d, as_in, bytes, cudaMemcpyHostToDevice, c->s_compute);
  if (e == cudaSuccess) rc = launch_grm(c, d, n, d, mb_out, c->s_compute, E);
  else rc = E.fail(TPG_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e));
  if (rc == TPG_OK) {
    e = cudaMemcpy(grm_out, d, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = E.fail(TPG_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return rc;
}

int tpg_filter_high_relatedness(int device, const double* matrix, int32_t on_device, int64_t n,
                                double kings_threshold, int32_t* keep, int32_t* order, int64_t* n_removed,
                                char* err, size_t errlen) {
  Err E{err, errlen};
  int rc = TPG_OK;
  Ctx* c = cached_ctx(device, E, &rc);
  if (!c) return rc;
  return run_filter_high_relatedness(c, matrix, on_device != 0, n, kings_threshold, keep, order, n_removed, E);
}

int tpg_king_filter(const tpg_job* job_in, double kings_threshold, int32_t* keep, int32_t* order,
                    int64_t* n_removed, double* king_host, tpg_stats* stats, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!job_in || !keep) return E.fail(TPG_ERR_ARG, "job / keep is NULL");
  tpg_job job = *job_in;
  job.stat = TPG_STAT_KING;
  int rc = TPG_OK;
  Ctx* c = cached_ctx(job.device, E, &rc);
  if (!c) return rc;
  const int64_t n = job.ind_row ? job.n : job.nrow_total;
  if (n <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row");
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  rc = ensure_slot(&c->d_out[0], bytes, false, "KING matrix", E);
  if (rc) return rc;
  tpg_result res;
  memset(&res, 0, sizeof res);
  res.out0 = static_cast<double*>(c->d_out[0].p);
  res.out_on_device = 1;
  rc = pairwise_many_on(c, &job, &job.stat, 1, &res, E);   // KING stays in HBM
  if (rc) return rc;
  const double t0 = now_ms();
  rc = run_filter_high_relatedness(c, res.out0, true, n, kings_threshold, keep, order, n_removed, E);
  res.stats.ms_finish += now_ms() - t0;  // the filter is accounted with the finishing stage
  if (rc == TPG_OK && king_host) {
    const void* dv[1] = {res.out0};
    void* hv[1] = {king_host};
    rc = results_to_host(c, dv, hv, 1, bytes, &res.stats.bytes_d2h, E);
  }
  if (stats) *stats = res.stats;
  return rc;
}

int tpg_king_filter_groups(const tpg_job* job_in, const int32_t* group, int32_t ngroups, double kings_threshold,
                           int32_t* keep, int64_t* n_removed, double* king_host, tpg_stats* stats, char* err,
                           size_t errlen) {
  Err E{err, errlen};
  if (!job_in || !group || !keep || ngroups <= 0) return E.fail(TPG_ERR_ARG, "job / group / keep is NULL or no groups");
  tpg_job job = *job_in;
  job.stat = TPG_STAT_KING;
  job.shard_mode = TPG_SHARD_NONE;
  job.shard_world = 1;
  job.shard_rank = 0;
  const int64_t n = job.ind_row ? job.n : job.nrow_total;
  if (n <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row");
  if (n > INT32_MAX) return E.fail(TPG_ERR_ARG, "too many individuals");
  // individuals sorted by group (stable: the order within a group is the job's), so that every group is a
  // diagonal block of ONE n x n KING job whose Gram kernels skip the tiles that hold no same-group pair
  std::vector<int64_t> start(size_t(ngroups) + 1, 0);
  for (int64_t k = 0; k < n; ++k) {
    if (group[k] < 0 || group[k] >= ngroups)
      return E.fail(TPG_ERR_ARG, "group[%lld] = %d is out of range [0, %d)", (long long)k, group[k], ngroups);
    start[size_t(group[k]) + 1]++;
  }
  for (int32_t g = 0; g < ngroups; ++g) start[g + 1] += start[g];
  std::vector<int32_t> perm(n), rows(n), gsorted(n);
  {
    std::vector<int64_t> fill(start.begin(), start.end() - 1);
    for (int64_t k = 0; k < n; ++k) perm[size_t(fill[group[k]]++)] = int32_t(k);
  }
  for (int64_t s = 0; s < n; ++s) {
    rows[s] = job.ind_row ? job.ind_row[perm[s]] : perm[s] + 1;
    gsorted[s] = group[perm[s]];
  }
  job.ind_row = rows.data();
  job.n = n;
  int rc = TPG_OK;
  Ctx* c = cached_ctx(job.device, E, &rc);
  if (!c) return rc;
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  rc = ensure_slot(&c->d_out[0], bytes, false, "KING matrix", E);
  if (rc) return rc;
  tpg_result res;
  memset(&res, 0, sizeof res);
  res.out0 = static_cast<double*>(c->d_out[0].p);
  res.out_on_device = 1;
  uint64_t sig = 1469598103934665603ull;
  for (int32_t g = 0; g <= ngroups; ++g) { sig ^= uint64_t(start[g]); sig *= 1099511628211ull; }
  rc = pairwise_many_on(c, &job, &job.stat, 1, &res, E, gsorted.data(), sig);   // KING stays in HBM
  if (rc) return rc;
  const double t0 = now_ms();
  std::vector<int32_t> keep_g;
  int64_t king_off = 0;
  for (int32_t g = 0; g < ngroups && rc == TPG_OK; ++g) {
    const int64_t s = start[g], ng = start[g + 1] - start[g];
    if (n_removed) n_removed[g] = 0;
    if (ng <= 0) continue;
    // the group's diagonal block as a dense ng x ng matrix
    const size_t gbytes = sizeof(double) * size_t(ng) * size_t(ng);
    rc = ensure_slot(&c->d_out[1], gbytes, false, "group KING matrix", E);
    if (rc) break;
    double* dense = static_cast<double*>(c->d_out[1].p);
    TPG_CUDA_OK(cudaMemcpy2DAsync(dense, sizeof(double) * size_t(ng), res.out0 + s + s * n, sizeof(double) * size_t(n),
                                  sizeof(double) * size_t(ng), size_t(ng), cudaMemcpyDeviceToDevice, c->s_compute), E);
    keep_g.assign(size_t(ng), 1);
    int64_t removed = 0;
    rc = run_filter_high_relatedness(c, dense, true, ng, kings_threshold, keep_g.data(), nullptr, &removed, E);
    if (rc) break;
    for (int64_t k = 0; k < ng; ++k) keep[perm[size_t(s + k)]] = keep_g[size_t(k)];
    if (n_removed) n_removed[g] = removed;
    if (king_host) {
      TPG_CUDA_OK(cudaMemcpyAsync(king_host + king_off, dense, gbytes, cudaMemcpyDeviceToHost, c->s_compute), E);
      TPG_CUDA_OK(cudaStreamSynchronize(c->s_compute), E);
      res.stats.bytes_d2h += int64_t(gbytes);
      king_off += ng * ng;
    }
  }
  res.stats.ms_finish += now_ms() - t0;  // the filters are accounted with the finishing stage
  if (stats) *stats = res.stats;
  return rc;
}

int tpg_as_wg17(const tpg_job* job_in, const int32_t* group, int64_t npop, double* inbreeding, double* fst,
                double* fis, double* mb_out, double* as_host, tpg_stats* stats, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!job_in) return E.fail(TPG_ERR_ARG, "job is NULL");
  tpg_job job = *job_in;
  job.stat = TPG_STAT_AS;
  int rc = TPG_OK;
  Ctx* c = cached_ctx(job.device, E, &rc);
  if (!c) return rc;
  const int64_t n = job.ind_row ? job.n : job.nrow_total;
  if (n <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row");
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  const size_t bytes = sizeof(double) * size_t(n) * size_t(n);
  rc = ensure_slot(&c->d_out[0], bytes, false, "allele-sharing matrix", E);
  if (rc) return rc;
  tpg_result res;
  memset(&res, 0, sizeof res);
  res.out0 = static_cast<double*>(c->d_out[0].p);
  res.out_on_device = 1;
  rc = pairwise_many_on(c, &job, &job.stat, 1, &res, E);   // allele sharing stays in HBM
  if (rc) return rc;
  const double t0 = now_ms();
  rc = run_wg17(c, res.out0, true, n, group, npop, inbreeding, fst, fis, mb_out, E);
  res.stats.ms_finish += now_ms() - t0;
  if (rc == TPG_OK && as_host) {
    const void* dv[1] = {res.out0};
    void* hv[1] = {as_host};
    rc = results_to_host(c, dv, hv, 1, bytes, &res.stats.bytes_d2h, E);
  }
  if (stats) *stats = res.stats;
  return rc;
}

int tpg_wg17(int device, const double* allele_sharing, int32_t on_device, int64_t n, const int32_t* group,
             int64_t npop, double* inbreeding, double* fst, double* fis, double* mb_out, char* err,
             size_t errlen) {
  Err E{err, errlen};
  int rc = TPG_OK;
  Ctx* c = cached_ctx(device, E, &rc);
  if (!c) return rc;
  return run_wg17(c, allele_sharing, on_device != 0, n, group, npop, inbreeding, fst, fis, mb_out, E);
}

int tpg_scan(const tpg_job* job_in, const double* code256, const double* ploidy, int32_t as_counts,
             int32_t* het_n, int32_t* na_n, double* loc0, double* loc1, int32_t* loc_na, tpg_stats* stats,
             char* err, size_t errlen) {
  Err E{err, errlen};
  if (!job_in || !code256) return E.fail(TPG_ERR_ARG, "job / code256 is NULL");
  if (job_in->shard_mode != TPG_SHARD_NONE && job_in->shard_world > 1)
    return E.fail(TPG_ERR_ARG, "scans run on one device");
  tpg_job job = *job_in;
  job.stat = TPG_STAT_AS;  // only selects the plan's bookkeeping; no planes or products are made
  job.shard_mode = TPG_SHARD_NONE;
  job.shard_world = 1;
  job.shard_rank = 0;
  const int64_t n = job.ind_row ? job.n : job.nrow_total;
  const int64_t m = job.ind_col ? job.m : job.ncol_total;
  if (n <= 0 || m <= 0) return E.fail(TPG_ERR_ARG, "empty ind.row or ind.col");
  // code256 -> byte table.  x > -1 is "typed" (src/gt_ind_hetero.cpp:29, alt_freq_dip_pseudo_cpp.cpp:37)
  uint8_t lut[256];
  for (int b = 0; b < 256; ++b) {
    const double x = code256[b];
    if (!(x > -1)) { lut[b] = 0xFF; continue; }
    if (x != std::floor(x) || x < 0 || x > 127)
      return E.fail(TPG_ERR_ARG, "code256[%d] = %g: this path takes integer dosages (CODE_012, CODE_IMPUTE_PRED ...), "
                                 "not fractional ones", b, x);
    lut[b] = uint8_t(x);
  }
  std::vector<uint8_t> wp(2 * n);
  for (int64_t i = 0; i < n; ++i) {
    const double pl = ploidy ? ploidy[i] : 2.0;
    if (pl != 1.0 && pl != 2.0)
      return E.fail(TPG_ERR_ARG, "ploidy[%lld] = %g: diploid (2) or pseudohaploid (1) individuals only "
                                 "(alt_freq_dip_pseudo_cpp)", (long long)i, pl);
    wp[i] = uint8_t(pl == 2.0 ? 2 : 1);  // 2 / (3 - ploidy)
    wp[n + i] = uint8_t(pl);
  }
  int rc = TPG_OK;
  Ctx* c = cached_ctx(job.device, E, &rc);
  if (!c) return rc;
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  tpg_stats local;
  memset(&local, 0, sizeof local);
  tpg_stats* st = stats ? stats : &local;
  memset(st, 0, sizeof *st);
  const int64_t l0 = c->launches;
  const size_t bytes = 256 + 2 * size_t(n) + 16 + sizeof(int32_t) * (2 * size_t(n) + 3 * size_t(m));
  rc = ensure_slot(&c->d_red, bytes, false, "scan accumulators", E);
  if (rc) return rc;
  uint8_t* base = static_cast<uint8_t*>(c->d_red.p);
  const size_t small = (256 + 2 * size_t(n) + 15) / 16 * 16;
  int32_t* acc = reinterpret_cast<int32_t*>(base + small);
  TPG_CUDA_OK(cudaMemcpyAsync(base, lut, 256, cudaMemcpyHostToDevice, c->s_compute), E);
  TPG_CUDA_OK(cudaMemcpyAsync(base + 256, wp.data(), 2 * size_t(n), cudaMemcpyHostToDevice, c->s_compute), E);
  TPG_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(int32_t) * (2 * size_t(n) + 3 * size_t(m)), c->s_compute), E);
  c->launches++;
  ScanArgs sa;
  {
    // the two tables tidypopgen uses (bigsnpr CODE_012 / CODE_IMPUTE_PRED) decode without a lookup
    bool t012 = true, timp = true;
    for (int b = 0; b < 256; ++b) {
      const uint8_t e012 = b <= 2 ? uint8_t(b) : 0xFF;
      const uint8_t eimp = b <= 2 ? uint8_t(b) : (b >= 4 && b <= 6) ? uint8_t(b - 4) : 0xFF;
      t012 &= lut[b] == e012;
      timp &= lut[b] == eimp;
    }
    sa.std_table = t012 ? 1 : timp ? 2 : 0;
  }
  sa.lut = base; sa.weight = base + 256; sa.ploidy = base + 256 + n;
  sa.het_n = acc; sa.na_n = acc + n; sa.alt2 = acc + 2 * n; sa.valid = sa.alt2 + m; sa.na_l = sa.valid + m;
  rc = accumulate_impl(c, &job, nullptr, st, E, /*defer=*/false, &sa);
  if (rc) return rc;
  std::vector<int32_t> h(2 * size_t(n) + 3 * size_t(m));
  TPG_CUDA_OK(cudaMemcpyAsync(h.data(), acc, sizeof(int32_t) * h.size(), cudaMemcpyDeviceToHost, c->s_compute), E);
  TPG_CUDA_OK(cudaStreamSynchronize(c->s_compute), E);
  st->bytes_d2h += int64_t(sizeof(int32_t) * h.size());
  if (het_n) memcpy(het_n, h.data(), sizeof(int32_t) * size_t(n));
  if (na_n) memcpy(na_n, h.data() + n, sizeof(int32_t) * size_t(n));
  const int32_t* alt2 = h.data() + 2 * n;
  const int32_t* valid = alt2 + m;
  const int32_t* nal = valid + m;
  if (loc_na) memcpy(loc_na, nal, sizeof(int32_t) * size_t(m));
  for (int64_t l = 0; l < m; ++l) {
    const double a = double(alt2[l]) * 0.5, v = double(valid[l]);  // exact: half-allele units
    if (loc1) loc1[l] = v;
    if (loc0) {
      if (as_counts) loc0[l] = a;
      else if (v > 0) loc0[l] = a / v;
      else { const uint64_t na_bits = 0x7FF00000000007A2ull; memcpy(&loc0[l], &na_bits, 8); }  // NA_real_
    }
  }
  st->n_launches = c->launches - l0;
  return TPG_OK;
}

int tpg_decode_planes(tpg_ctx* h, const uint8_t* d_raw, int64_t raw_bytes, int64_t raw_stride,
                      const int32_t* d_row_idx, int64_t row0, int64_t n, int64_t n_pad,
                      const int32_t* d_col_idx, int64_t col0, int64_t lc, int64_t lc_pad,
                      int8_t* d_planes, int64_t plane_stride, int32_t k_major, char* err,
                      size_t errlen) {
  Err E{err, errlen};
  if (!h) return E.fail(TPG_ERR_ARG, "ctx is NULL");
  Ctx* c = reinterpret_cast<Ctx*>(h);
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  DecodeArgs a;
  a.raw = d_raw; a.raw_bytes = raw_bytes; a.raw_stride = raw_stride;
  a.row_idx = d_row_idx; a.row0 = row0; a.n = n; a.n_pad = n_pad;
  a.col_idx = d_col_idx; a.col0 = col0; a.lc = lc; a.lc_pad = lc_pad;
  a.planes = d_planes; a.plane_stride = plane_stride; a.nplanes = 3; a.k_major = (k_major & 1) != 0;
  a.packed = (k_major & 4) != 0;
  a.k4 = (k_major & 8) != 0;
  int rc = launch_decode(c, a, c->s_compute, E);
  if (rc) return rc;
  TPG_CUDA_OK(cudaStreamSynchronize(c->s_compute), E);
  return TPG_OK;
}

int tpg_gram(tpg_ctx* h, const int8_t* d_planes, int64_t plane_stride, int64_t n_pad,
             int64_t lc_pad, int32_t k_major, int32_t prod_mask, int32_t shard_rank,
             int32_t shard_world, int32_t* d_counts, float* ms_out, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!h) return E.fail(TPG_ERR_ARG, "ctx is NULL");
  Ctx* c = reinterpret_cast<Ctx*>(h);
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  GramArgs a;
  a.planes = d_planes; a.plane_stride = plane_stride; a.n = n_pad; a.n_pad = n_pad;
  a.lc_pad = lc_pad; a.k_major = (k_major & 1) != 0; a.ctas = (k_major & 2) ? 1 : 0; a.mxf4 = (k_major & 8) != 0;
  a.prod_mask = prod_mask;
  a.shard_rank = shard_rank; a.shard_world = shard_world; a.counts = d_counts; a.ld = n_pad;
  if (shard_world > 1) {
    const uint8_t* owner_d = nullptr;
    int rc0 = owner_table(c, n_pad, shard_world, &a.owner, &owner_d, E);
    if (rc0) return rc0;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, c->s_compute);
  int rc = launch_gram(c, a, c->s_compute, nullptr, E);
  cudaEventRecord(e1, c->s_compute);
  cudaError_t e = cudaStreamSynchronize(c->s_compute);
  if (rc == TPG_OK && e != cudaSuccess) {
    Err quiet{nullptr, 0};
    rc = check_err_flag(c, quiet) == TPG_ERR_KERNEL ? check_err_flag(c, E)
                                                     : E.fail(TPG_ERR_CUDA, "gram failed: %s", cudaGetErrorString(e));
  }
  if (rc == TPG_OK) rc = check_err_flag(c, E);
  if (rc == TPG_OK && ms_out) cudaEventElapsedTime(ms_out, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return rc;
}

int tpg_mma_peak(int device, int32_t kind, double seconds, double* tops, double* ms, char* err, size_t errlen) {
  Err E{err, errlen};
  int rc = TPG_OK;
  Ctx* c = cached_ctx(device, E, &rc);
  if (!c) return rc;
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  return run_mma_peak(c, kind, seconds, tops, ms, E);
}

int tpg_synth_fill(tpg_ctx* h, uint8_t* d_raw, int64_t nrow, int64_t ncol, int64_t col_offset,
                   uint64_t seed, double missing_rate, int32_t mode, char* err, size_t errlen) {
  Err E{err, errlen};
  if (!h) return E.fail(TPG_ERR_ARG, "ctx is NULL");
  Ctx* c = reinterpret_cast<Ctx*>(h);
  TPG_CUDA_OK(cudaSetDevice(c->device), E);
  int rc = launch_synth(c, d_raw, nrow, ncol, col_offset, seed, missing_rate, mode, c->s_compute, E);
  if (rc) return rc;
  TPG_CUDA_OK(cudaStreamSynchronize(c->s_compute), E);
  return TPG_OK;
}

/* Host twin of tpg_synth_fill (fixtures for the CPU-side checks and the CPU baseline). */
void tpg_synth_fill_host(uint8_t* raw, int64_t nrow, int64_t ncol, int64_t col_offset, uint64_t seed,
                         double missing_rate, int32_t mode) {
  const uint32_t thr = uint32_t(missing_rate * 4294967296.0);
  for (int64_t l = 0; l < ncol; ++l)
    for (int64_t i = 0; i < nrow; ++i)
      raw[l * nrow + i] = tpg_synth_byte(seed, uint64_t(l + col_offset), uint64_t(i), thr, mode);
}

}  // extern "C"
