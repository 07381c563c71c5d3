// gram_i8: exact int8 x int8 -> s32 Gram products of the indicator planes on the 5th-gen
// tensor cores (tcgen05.mma kind::i8, accumulators in TMEM, operands staged by TMA).
//
// Replaces the FP64 dsyrk/dgemm products of the reference:
//   KING  src/snp_king.cpp:70-72   K += g1 g1' - 2 (g0 g2' + g2 g0');  N_Aa_i += g1 valid'
//   IBS   src/snp_ibs.cpp:67-72    K += 2(g2g2'+g1g1'+g0g0') + g1(g0+g2)' + (g0+g2)g1';  K2 += 2 nm nm'
//   AS    src/snp_as.cpp:64-65     K += dos dos';  K2 += na na'
// With d = g2 - g0 (-1/0/+1), NM = g0+g1+g2, H1 = g1 all of these are linear
// combinations of four integer products ("Algebra"):
//   DD = d d'   NN = NM NM'   HH = H1 H1'   (symmetric: upper-triangle tiles only)
//   HN = H1 NM'                              (full)
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0   TMA producer       (one lane)    global planes -> 4-stage smem ring
//   warp 1   MMA issuer         (one lane)    4 x tcgen05.mma (M128 N256 K32) per stage
//   warps 2-5 epilogue          TMEM -> registers -> red.global.add.s32 (coalesced, exact)
// TMEM holds two 128x256 s32 accumulators, so the epilogue of one work item overlaps the
// MMAs of the next.  A work item is (product, tile, k-block range); the host orders the
// list so that CTAs running at the same time share operand panels in L2.
#include <algorithm>
#include <cstring>
#include <functional>
#include <queue>
#include "internal.h"
#include "ptx.cuh"

namespace tpg {

constexpr int kStages = 4;
constexpr int kABytes = kTileM * kTileK;            // 16 KiB
constexpr int kBBytes = kTileN * kTileK;            // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;      // 48 KiB
constexpr int kGramThreads = 192;
constexpr int kTmemCols = 512;                      // 2 accumulators x 256 columns
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

// product -> (A plane, B plane); planes: 0 = d, 1 = NM, 2 = H1
__device__ __forceinline__ void prod_planes(int prod, int& pa, int& pb) {
  pa = (prod == TPG_MAT_DD) ? 0 : (prod == TPG_MAT_NN) ? 1 : 2;
  pb = (prod == TPG_MAT_DD) ? 0 : (prod == TPG_MAT_HH) ? 2 : 1;
}

// WorkItem.prod_ti: product in bits [0,4), B columns / 16 in [4,9) (1..16), A row tile in [9,32).
__host__ __device__ __forceinline__ int item_prod(const WorkItem& w) { return w.prod_ti & 15; }
__host__ __device__ __forceinline__ int item_ncols(const WorkItem& w) { return ((w.prod_ti >> 4) & 31) * 16; }
__host__ __device__ __forceinline__ int item_ti(const WorkItem& w) { return w.prod_ti >> 9; }

struct GramParams {
  const WorkItem* items;   // grouped by CTA
  const int32_t* cta_off;  // [gridDim.x + 1] item ranges
  int32_t* counts;
  int64_t mat_stride;  // elements between count matrices
  int64_t ld;
  int n;               // valid individuals
  int* err_flag;
};

template <bool kMajorK>
__global__ void __launch_bounds__(kGramThreads, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_nm,
               const __grid_constant__ CUtensorMap tm_h1, const GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                    // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_d);
    tma_prefetch_desc(&tm_nm);
    tma_prefetch_desc(&tm_h1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int it0 = p.cta_off[blockIdx.x], it1 = p.cta_off[blockIdx.x + 1];

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        const WorkItem w = p.items[it];
        const int prod = item_prod(w), ti = item_ti(w), c0 = w.c0, nc = item_ncols(w);
        int pa, pb;
        prod_planes(prod, pa, pb);
        const CUtensorMap* ma = pa == 0 ? &tm_d : pa == 1 ? &tm_nm : &tm_h1;
        const CUtensorMap* mb = pb == 0 ? &tm_d : pb == 1 ? &tm_nm : &tm_h1;
        const bool two = nc > 128;  // the second 128-row box of B is only fetched when used
        const uint32_t bytes = two ? kStageBytes : kStageBytes - kABytes;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          if (kMajorK) {  // planes[row][locus]: coords (locus, row)
            tma_load_2d(sa, ma, &full_bar[stage], kb * kTileK, ti * kTileM);
            tma_load_2d(sb, mb, &full_bar[stage], kb * kTileK, c0);
            if (two) tma_load_2d(sb + kABytes, mb, &full_bar[stage], kb * kTileK, c0 + 128);
          } else {        // planes[locus][row]: coords (row, locus)
            tma_load_2d(sa, ma, &full_bar[stage], ti * kTileM, kb * kTileK);
            tma_load_2d(sb, mb, &full_bar[stage], c0, kb * kTileK);
            if (two) tma_load_2d(sb + kABytes, mb, &full_bar[stage], c0 + 128, kb * kTileK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc0 = make_idesc_i8(kTileM, 0, !kMajorK);  // N is per work item
      // K-major  SW128: 8-row groups 1024 B apart (SBO); one MMA consumes 32 B of K.
      // MN-major SW128: 8-locus groups 1024 B apart (SBO), 128-row blocks 16 KiB apart (LBO);
      //                 one MMA consumes 32 loci = 4096 B.
      constexpr uint32_t lbo = kMajorK ? 16u : uint32_t(kABytes);
      constexpr uint32_t sbo = 1024u;
      constexpr uint32_t kstep = kMajorK ? (32u >> 4) : (4096u >> 4);  // descriptor units of 16 B
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = it0; it < it1; ++it) {
        const WorkItem w = p.items[it];
        const uint32_t idesc = idesc0 | (uint32_t(item_ncols(w) >> 3) << 17);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * kTileN);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa, lbo, sbo);
          const uint64_t db = make_smem_desc_sw128(sa + kABytes, lbo, sbo);
#pragma unroll
          for (int k = 0; k < kTileK / 32; ++k) {
            umma_i8(tmem_d, da + uint64_t(k * kstep), db + uint64_t(k * kstep), idesc,
                    (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int it = it0; it < it1; ++it) {
      const WorkItem w = p.items[it];
      const int prod = item_prod(w), ti = item_ti(w), nc = item_ncols(w);
      const bool sym = prod != TPG_MAT_HN;
      mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int i = ti * kTileM + q * 32 + lane;  // A row  = count-matrix row
      const int j0 = w.c0;                        // B rows = count-matrix columns
      int32_t* cbase = p.counts + int64_t(prod) * p.mat_stride + i;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kTileN);
#pragma unroll 1
      for (int c = 0; c < nc; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(c), v);
        tmem_ld_wait();
        if (c + 32 >= nc) {  // accumulator drained: hand it back before the last adds
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (j0 + c < p.n && i < p.n) {
          const int lim = min(32, nc - c);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int j = j0 + c + jj;
            // X(i,j) at X[i + j*ld]: the 32 lanes of a warp hit 32 consecutive ints.
            if (jj < lim && j < p.n && (!sym || i <= j)) atomicAdd(cbase + int64_t(j) * p.ld, int32_t(v[jj]));
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_plane_map(Ctx* ctx, CUtensorMap* out, const int8_t* base, uint64_t inner,
                          uint64_t outer, const Err& E) {
  if (!ctx->encode_tiled) return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled entry point missing");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner};  // bytes, dim 1
  cuuint32_t box[2] = {128, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
  return TPG_OK;
}

// How many k-splits a tile gets: enough work items per CTA (>= 16) that the list schedule
// below ends within a few percent of the ideal makespan, while an item keeps >= 16 k-blocks
// so its epilogue (drained while the next item's MMAs run) stays amortised.
static int choose_ksplit(int64_t n_tiles, int64_t kblocks, int sms) {
  if (n_tiles >= int64_t(16) * sms) return 1;
  int64_t ks = (int64_t(16) * sms + n_tiles - 1) / n_tiles;
  ks = std::min<int64_t>(ks, std::max<int64_t>(1, kblocks / 16));
  return int(std::max<int64_t>(1, ks));
}

int launch_gram(Ctx* ctx, const GramArgs& a, cudaStream_t st, int64_t* macs_out, const Err& E) {
  if (a.n_pad % kRowPad || a.lc_pad % kLociPad || a.lc_pad <= 0)
    return E.fail(TPG_ERR_ARG, "gram: n_pad (%lld) / lc_pad (%lld) not tile multiples",
                  (long long)a.n_pad, (long long)a.lc_pad);
  const int64_t n = std::min(a.n, a.n_pad);
  const int64_t tiles_m = (n + kTileM - 1) / kTileM, tiles_n = (n + kTileN - 1) / kTileN;
  const int64_t nblocks = a.n_pad / kTileN;  // ownership grid (256 x 256 blocks)
  const int64_t kblocks = a.lc_pad / kTileK;

  // ---- tile list in L2-friendly order: 1024x1024 regions, then tiles, then products.
  // A tile is 128 rows x up to 256 columns; the column extent is trimmed to the valid
  // individuals (multiple of 16) and, for symmetric products, to start at the diagonal.
  struct T { int prod, ti, c0, nc; };
  std::vector<T> tiles;
  const int64_t RI = 8, RJ = 4;
  int64_t macs_per_locus = 0;
  for (int64_t rj = 0; rj < tiles_n; rj += RJ)
    for (int64_t ri = 0; ri < tiles_m; ri += RI)
      for (int64_t tj = rj; tj < std::min(rj + RJ, tiles_n); ++tj)
        for (int64_t ti = ri; ti < std::min(ri + RI, tiles_m); ++ti) {
          const int64_t col_lo = kTileN * tj;
          const int64_t col_hi = std::min<int64_t>(col_lo + kTileN, (n + 15) / 16 * 16);  // exclusive
          const int64_t bi = ti / 2, bj = tj;
          const int64_t lo = std::min(bi, bj), hi = std::max(bi, bj);
          if (!block_owned(lo, hi, nblocks, a.shard_rank, a.shard_world)) continue;
          for (int prod = 0; prod < TPG_NUM_COUNT_MATS; ++prod) {
            if (!(a.prod_mask & (1 << prod))) continue;
            int64_t c0 = col_lo;
            if (prod != TPG_MAT_HN) c0 = std::max<int64_t>(c0, kTileM * ti);  // nothing left of the diagonal
            if (c0 >= col_hi) continue;
            tiles.push_back({prod, int(ti), int(c0), int(col_hi - c0)});
            macs_per_locus += int64_t(kTileM) * (col_hi - c0);
          }
        }
  if (macs_out) *macs_out = macs_per_locus * a.lc_pad;
  if (tiles.empty()) return TPG_OK;

  const int sms = ctx->num_sms;
  const int ks = choose_ksplit(int64_t(tiles.size()), kblocks, sms);

  // The work list only depends on the geometry; chunks of equal size reuse the upload.
  char key[160];
  snprintf(key, sizeof key, "%lld/%lld/%lld/%d/%d/%d/%d", (long long)n, (long long)a.n_pad,
           (long long)a.lc_pad, a.prod_mask, a.shard_rank, a.shard_world, ks);
  const int grid = int(std::min<size_t>(tiles.size() * size_t(ks), size_t(sms)));
  if (ctx->items_key != key) {
    // List scheduling on the host (deterministic stand-in for a device-side work queue):
    // items in k-split-major order go to the CTA with the least accumulated MMA time.
    std::vector<std::vector<WorkItem>> per(grid);
    typedef std::pair<int64_t, int> Load;  // (accumulated cost, cta)
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int c = 0; c < grid; ++c) heap.push({0, c});
    for (int s = 0; s < ks; ++s) {
      const int32_t kb0 = int32_t(kblocks * s / ks), kb1 = int32_t(kblocks * (s + 1) / ks);
      if (kb1 <= kb0) continue;
      for (const T& t : tiles) {
        Load l = heap.top();
        heap.pop();
        per[l.second].push_back({int32_t(t.prod | ((t.nc / 16) << 4) | (t.ti << 9)), t.c0, kb0, kb1});
        // an MMA costs max(N, ~64) cycles-equivalents; the epilogue adds a little per item
        l.first += int64_t(std::max(t.nc, 64)) * (kb1 - kb0) + 64;
        heap.push(l);
      }
    }
    std::vector<WorkItem>& blob = ctx->h_items;  // [offsets (as WorkItem-sized words) | items]
    const size_t off_words = (size_t(grid) + 1 + 3) / 4;  // int32 offsets packed 4 per WorkItem
    size_t total = 0;
    for (auto& v : per) total += v.size();
    blob.assign(off_words + total, WorkItem{0, 0, 0, 0});
    int32_t* offs = reinterpret_cast<int32_t*>(blob.data());
    size_t pos = 0;
    for (int c = 0; c < grid; ++c) {
      offs[c] = int32_t(pos);
      for (const WorkItem& w : per[c]) blob[off_words + pos++] = w;
    }
    offs[grid] = int32_t(pos);
    if (blob.size() > ctx->d_items_cap) {
      TPG_CUDA_OK(cudaStreamSynchronize(st), E);
      if (ctx->d_items) cudaFree(ctx->d_items);
      ctx->d_items = nullptr;
      ctx->d_items_cap = 0;
      const size_t cap = blob.size() + blob.size() / 4 + 1024;
      TPG_CUDA_OK(cudaMalloc(&ctx->d_items, cap * sizeof(WorkItem)), E);
      ctx->d_items_cap = cap;
    }
    // Stream-ordered after any kernel still reading the previous list.
    TPG_CUDA_OK(cudaMemcpyAsync(ctx->d_items, blob.data(), blob.size() * sizeof(WorkItem),
                                cudaMemcpyHostToDevice, st), E);
    ctx->items_key = key;
    ctx->items_off_words = off_words;
  }

  CUtensorMap maps[3];
  for (int pl = 0; pl < 3; ++pl) {
    const int8_t* base = a.planes + int64_t(pl) * a.plane_stride;
    const uint64_t inner = a.k_major ? uint64_t(a.lc_pad) : uint64_t(a.n_pad);
    const uint64_t outer = a.k_major ? uint64_t(a.n_pad) : uint64_t(a.lc_pad);
    // AS / GRM jobs carry two planes only: alias the missing H1 map to plane 0 (never used).
    const bool present = pl < 2 || (a.prod_mask & ((1 << TPG_MAT_HH) | (1 << TPG_MAT_HN)));
    int rc = make_plane_map(ctx, &maps[pl], present ? base : a.planes, inner, outer, E);
    if (rc) return rc;
  }

  GramParams p;
  p.items = ctx->d_items + ctx->items_off_words;
  p.cta_off = reinterpret_cast<const int32_t*>(ctx->d_items);
  p.counts = a.counts;
  p.mat_stride = a.ld * a.ld;
  p.ld = a.ld;
  p.n = int(a.n);
  p.err_flag = ctx->d_err_flag;

  if (a.k_major) {
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), E);
    gram_i8_kernel<true><<<grid, kGramThreads, kSmemBytes, st>>>(maps[0], maps[1], maps[2], p);
  } else {
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), E);
    gram_i8_kernel<false><<<grid, kGramThreads, kSmemBytes, st>>>(maps[0], maps[1], maps[2], p);
  }
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

}  // namespace tpg
