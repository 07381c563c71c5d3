// tpg_config: the one place where the library's tuning switches live (include/tpg_b200.h).
// SURVEY.md section 5 asked for options(tidypopgen.gpu.*) rather than environment sprawl: the R
// glue fills this struct from R options, the Python mirror from tidypopgen_b200.config().
// tpg_config_init() is the ONLY reader of TPG_* environment variables.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "config.h"

namespace {

std::mutex g_mu;
tpg_config g_cfg;
bool g_init = false;

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}
long long env_ll(const char* name, long long dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoll(e) : dflt;
}
double env_f(const char* name, double dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atof(e) : dflt;
}
bool env_set(const char* name) {
  const char* e = getenv(name);
  return e && *e && strcmp(e, "0") != 0;
}

}  // namespace

extern "C" {

void tpg_config_init(tpg_config* c) {
  if (!c) return;
  memset(c, 0, sizeof *c);
  c->abi_version = TPG_ABI_VERSION;
  c->gram_route = env_set("TPG_GRAM_I8") ? 1 : 0;
  c->gram_ctas = env_int("TPG_GRAM_CTAS", 0);
  c->gram_ksplit = env_int("TPG_GRAM_KS", 0);
  c->gram_max_kblocks = env_int("TPG_GRAM_MAXKB", 0);
  c->gram_clusters = env_int("TPG_GRAM_CLUSTERS", 0);
  c->gram_no_trim = env_set("TPG_GRAM_NO_TRIM") ? 1 : 0;
  c->gram_wide = env_int("TPG_GRAM_WIDE", 0);
  c->gram_l2_rows = env_int("TPG_GRAM_L2_ROWS", 0);
  c->gram_l2_cols = env_int("TPG_GRAM_L2_COLS", 0);
  c->gram_order = env_int("TPG_GRAM_ORDER", 0);
  c->pack_mode = env_int("TPG_PACK_MODE", 0);
  if (env_set("TPG_NO_PACK")) c->pack_mode = 1;
  else if (env_set("TPG_NO_HYBRID")) c->pack_mode = 2;
  c->pk2_i8 = env_set("TPG_PK2_I8") ? 1 : 0;
  c->no_compact = env_set("TPG_NO_COMPACT") ? 1 : 0;
  c->gather_smem_kb = env_int("TPG_GATHER_SMEM_KB", 0);
  c->gather_layout = env_int("TPG_GATHER_LAYOUT", 0);
  c->io_mode = env_int("TPG_IO_MODE", 0);
  if (env_set("TPG_NO_PREAD")) c->io_mode = 2;
  c->no_thp = env_set("TPG_NO_THP") ? 1 : 0;
  c->decode_k4_v1 = env_set("TPG_DECODE_K4_V1") ? 1 : 0;
  c->host_threads = env_int("TPG_HOST_THREADS", 0);
  c->scan_loci_per_cta = env_int("TPG_SCAN_LOCI", 0);
  c->scan_no_indiv = env_set("TPG_SCAN_NO_INDIV") ? 1 : 0;
  c->scan_table = env_int("TPG_SCAN_TABLE", 0);
  c->cache_counts = env_set("TPG_CACHE_COUNTS") ? 1 : 0;
  c->cache_genotypes = env_set("TPG_CACHE_GENOTYPES") ? 1 : 0;
  c->result_shells = env_int("TPG_RESULT_SHELLS", 0);
  c->device_chunk_loci = env_ll("TPG_DEVICE_CHUNK_LOCI", 0);
  c->host_chunk_mb = env_ll("TPG_HOST_CHUNK_MB", 0);
  c->link_gbps = env_f("TPG_LINK_GBPS", 0.0);
  c->hybrid_factor = env_f("TPG_HYBRID_FACTOR", 0.0);
}

int tpg_config_set(const tpg_config* c, char* err, size_t errlen) {
  auto fail = [&](const char* msg) {
    if (err && errlen) snprintf(err, errlen, "tpg_config_set: %s", msg);
    return int(TPG_ERR_ARG);
  };
  if (!c) return fail("config is NULL");
  if (c->abi_version != TPG_ABI_VERSION) return fail("ABI version mismatch");
  if (c->gram_route < 0 || c->gram_route > 1) return fail("gram_route must be 0 (4-bit where exact) or 1 (int8)");
  if (c->gram_ctas < 0 || c->gram_ctas > 2) return fail("gram_ctas must be 0, 1 or 2");
  if (c->gram_ksplit < 0 || c->gram_max_kblocks < 0 || c->gram_clusters < 0) return fail("gram_* must be >= 0");
  if (c->gram_l2_rows < 0 || c->gram_l2_cols < 0 || c->gram_l2_rows > 4096 || c->gram_l2_cols > 4096) return fail("gram_l2_rows / gram_l2_cols out of range");
  if (c->gram_order < 0 || c->gram_order > 2) return fail("gram_order must be 0 (auto), 1 (product-major) or 2 (tile-major)");
  if (c->gram_wide < 0 || c->gram_wide > 2) return fail("gram_wide must be 0 (auto), 1 (never) or 2 (always)");
  if (c->pack_mode < 0 || c->pack_mode > 2) return fail("pack_mode must be 0 (hybrid), 1 (never) or 2 (always)");
  if (c->gather_layout < 0 || c->gather_layout > 1) return fail("gather_layout must be 0 or 1");
  if (c->gather_smem_kb < 0 || c->gather_smem_kb > 200) return fail("gather_smem_kb must be in [0, 200]");
  if (c->io_mode < 0 || c->io_mode > 2) return fail("io_mode must be 0 (page cache), 1 (O_DIRECT) or 2 (mmap)");
  if (c->host_threads < 0 || c->host_threads > 1024) return fail("host_threads out of range");
  if (c->scan_loci_per_cta < 0) return fail("scan_loci_per_cta must be >= 0");
  if (c->result_shells < 0 || c->result_shells > 8) return fail("result_shells must be 0 (auto), 1 (off) or 2..8");
  if (c->scan_table < 0 || c->scan_table > 3) return fail("scan_table must be 0 (auto), 1 (lookup table), 2 (16-locus rounds) or 3 (warp per locus)");
  if (c->device_chunk_loci < 0 || c->host_chunk_mb < 0) return fail("chunk sizes must be >= 0");
  if (c->link_gbps < 0 || c->hybrid_factor < 0) return fail("link_gbps / hybrid_factor must be >= 0");
  std::lock_guard<std::mutex> lk(g_mu);
  g_cfg = *c;
  g_init = true;
  return TPG_OK;
}

void tpg_config_get(tpg_config* c) {
  if (!c) return;
  *c = tpg::cfg();
}

}  // extern "C"

namespace tpg {

tpg_config cfg() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_init) {
    tpg_config_init(&g_cfg);
    g_init = true;
  }
  return g_cfg;
}

}  // namespace tpg
