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
// Five kernels share one skeleton — persistent CTAs, 192 threads, warp-specialised:
//   warp 0   TMA producer       (one lane)    global planes -> shared-memory ring
//   warp 1   MMA issuer         (one lane)    tcgen05.mma per stage, accumulators in TMEM
//   warps 2-5 epilogue          TMEM -> registers -> the int32 count matrices (exact, order-independent adds)
// and one host scheduler (launch_gram / launch_gram_mxf4: work lists ordered so that CTAs running at the same
// time share operand panels in L2, cut into runs of equal-cost items):
//   gram_i8_kernel          one CTA per SM, kind::i8 M128 x N256, red.global epilogue (round 1's first kernel; cross-check)
//   gram_i8_pair_kernel     CTA pairs (cta_group::2), kind::i8 256 x 256 tiles, two accumulators ping-pong
//   gram_i8_wide_kernel     CTA pairs, kind::i8 256 x 512 items (one A stage feeds two B stages)
//   gram_mxf4_pair_kernel   CTA pairs, kind::mxf4 on 4-bit E2M1 planes, 256 x 240 tiles ping-pong (the default route)
//   gram_mxf4_wide_kernel   CTA pairs, kind::mxf4 256 x 480 items (the default for large matrices)
// The pair and wide kernels' epilogues stage 32 x 32 boxes in shared memory and add them with TMA reductions
// (cp.reduce.async.bulk.tensor).  A work item is (product, tile, k-block range).
#include <algorithm>
#include <cstdlib>
#include <cstring>
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

struct GramParams {
  const WorkItem* items;
  int n_items;
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

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const WorkItem w = p.items[it];
        const int prod = w.prod_ti & 15, ti = w.prod_ti >> 4, tj = w.tj;
        int pa, pb;
        prod_planes(prod, pa, pb);
        const CUtensorMap* ma = pa == 0 ? &tm_d : pa == 1 ? &tm_nm : &tm_h1;
        const CUtensorMap* mb = pb == 0 ? &tm_d : pb == 1 ? &tm_nm : &tm_h1;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          if (kMajorK) {  // planes[row][locus]: coords (locus, row)
            tma_load_2d(sa, ma, &full_bar[stage], kb * kTileK, ti * kTileM);
            tma_load_2d(sb, mb, &full_bar[stage], kb * kTileK, tj * kTileN);
            tma_load_2d(sb + kABytes, mb, &full_bar[stage], kb * kTileK, tj * kTileN + 128);
          } else {        // planes[locus][row]: coords (row, locus)
            tma_load_2d(sa, ma, &full_bar[stage], ti * kTileM, kb * kTileK);
            tma_load_2d(sb, mb, &full_bar[stage], tj * kTileN, kb * kTileK);
            tma_load_2d(sb + kABytes, mb, &full_bar[stage], tj * kTileN + 128, kb * kTileK);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_i8(kTileM, kTileN, !kMajorK);
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
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const WorkItem w = p.items[it];
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
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const WorkItem w = p.items[it];
      const int prod = w.prod_ti & 15, ti = w.prod_ti >> 4, tj = w.tj;
      const bool sym = prod != TPG_MAT_HN;
      mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int i = ti * kTileM + q * 32 + lane;  // A row  = count-matrix row
      const int j0 = tj * kTileN;                 // B rows = count-matrix columns
      int32_t* cbase = p.counts + int64_t(prod) * p.mat_stride + i;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kTileN);
#pragma unroll 1
      for (int c = 0; c < kTileN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(c), v);
        tmem_ld_wait();
        if (c + 32 == kTileN) {  // accumulator drained: hand it back before the last adds
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        if (j0 + c < p.n && i < p.n) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int j = j0 + c + jj;
            // X(i,j) at X[i + j*ld]: the 32 lanes of a warp hit 32 consecutive ints.
            if (j < p.n && (!sym || i <= j)) atomicAdd(cbase + int64_t(j) * p.ld, int32_t(v[jj]));
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

// ------------------------------------------------------------------ CTA-pair variant
// Two CTAs of one cluster (a TPC's SM pair) share every MMA: tcgen05.mma.cta_group::2 with
// M = 256 (128 A rows from each CTA's smem) x N = 256 (128 B rows from each CTA's smem), so
// a CTA stages 32 KiB per 128 loci instead of 48 KiB — a third less L2 -> smem traffic per
// MAC, which is what bounds the single-CTA kernel ("gram_i8").  A work item is a
// 256 x 256 tile; CTA r of the pair owns its rows [128 r, 128 r + 128) (TMEM lanes) and
// loads B rows [128 r, 128 r + 128).  Only the leader CTA (rank 0) issues MMAs; TMA bytes of
// both CTAs are counted on the leader's full barrier; tcgen05.commit multicasts the "slot
// free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps hand the
// accumulator back with a remote arrive on the leader's barrier.
constexpr int kPairStages = 6;
constexpr int kPairTileM = 256;
constexpr int kHalfBytes = 128 * kTileK;              // one 128 x 128 box, 16 KiB
constexpr int kPairStageBytes = 2 * kHalfBytes;       // per CTA: A half + B half
constexpr int kPairEpiBytes = 4 * 2 * 4096;             // epilogue staging: per warp two 32 x 32 int32 boxes (TMA reductions)
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + kPairEpiBytes + 1024 + 256;

template <bool kMajorK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGramThreads, 1)
gram_i8_pair_kernel(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_nm,
                    const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_counts,
                    const GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* epi = smem + kPairStages * kPairStageBytes;   // [4 warps][2][32 x 32 int32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kPairEpiBytes);
  uint64_t* full_bar = bars;                          // [kPairStages] used in the leader only
  uint64_t* empty_bar = bars + kPairStages;           // [kPairStages] in each CTA
  uint64_t* tfull_bar = bars + 2 * kPairStages;       // [2] in each CTA
  uint64_t* tempty_bar = bars + 2 * kPairStages + 2;  // [2] used in the leader only (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_d);
    tma_prefetch_desc(&tm_nm);
    tma_prefetch_desc(&tm_h1);
    tma_prefetch_desc(&tm_counts);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const int prod = w.prod_ti & 15, ti = w.prod_ti >> 8, tj = w.tj;
        const int ncols = (((w.prod_ti >> 4) & 15) + 1) * 16;  // N of this tile's MMAs (trimmed last tile column)
        int pa, pb;
        prod_planes(prod, pa, pb);
        const CUtensorMap* ma = pa == 0 ? &tm_d : pa == 1 ? &tm_nm : &tm_h1;
        const CUtensorMap* mb = pb == 0 ? &tm_d : pb == 1 ? &tm_nm : &tm_h1;
        const int row_a = ti * kPairTileM + int(rank) * 128;
        // CTA r supplies B rows [r N/2, (r+1) N/2) of the tile; the box is 128 rows, the MMA reads N/2
        const int row_b = tj * kTileN + int(rank) * (ncols / 2);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          uint8_t* sa = smem + stage * kPairStageBytes;
          uint8_t* sb = sa + kHalfBytes;
          // the leader expects the bytes of BOTH CTAs on its barrier
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (kMajorK) {
            tma_load_2d_pair(sa, ma, bar0, kb * kTileK, row_a);
            tma_load_2d_pair(sb, mb, bar0, kb * kTileK, row_b);
          } else {
            tma_load_2d_pair(sa, ma, bar0, row_a, kb * kTileK);
            tma_load_2d_pair(sb, mb, bar0, row_b, kb * kTileK);
          }
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc0 = make_idesc_i8(kPairTileM, 0, !kMajorK);
      constexpr uint32_t lbo = kMajorK ? 16u : uint32_t(kHalfBytes);  // one 128-row block: unused
      constexpr uint32_t sbo = 1024u;
      constexpr uint32_t kstep = kMajorK ? (32u >> 4) : (4096u >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const uint32_t ncols = (((uint32_t(w.prod_ti) >> 4) & 15u) + 1u) * 16u;
        const uint32_t idesc = idesc0 | ((ncols >> 3) << 17);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * kTileN);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa, lbo, sbo);
          const uint64_t db = make_smem_desc_sw128(sa + kHalfBytes, lbo, sbo);
#pragma unroll
          for (int k = 0; k < kTileK / 32; ++k) {
            umma_i8_pair(tmem_d, da + uint64_t(k * kstep), db + uint64_t(k * kstep), idesc,
                         (kb > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 3);  // frees the slot in both CTAs
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[acc], 3);      // accumulators of both CTAs complete
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..5, both CTAs)
    // 32 x 32 boxes go through a staging buffer in shared memory and ONE TMA reduction each (as in the wide kernels):
    // with red.global — ~1.3 cycles per lane and value — draining a 128 x 256 accumulator took longer than the MMAs of a
    // short item (C2: 69 k-blocks), and the ping-pong hid nothing.  Whole boxes are added: entries below the diagonal of
    // the symmetric products receive values too (nothing reads them), padding rows / columns add zeros.
    const int q = warp & 3;
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    uint8_t* ebase = epi + q * 2 * 4096;
    for (int it = cluster; it < p.n_items; it += n_clusters) {
      const WorkItem w = p.items[it];
      const int prod = w.prod_ti & 15, ti = w.prod_ti >> 8, tj = w.tj;
      const int ncols = (((w.prod_ti >> 4) & 15) + 1) * 16;   // multiple of 32 (host)
      mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int i0 = ti * kPairTileM + int(rank) * 128 + q * 32;   // this warp's 32 rows
      const int j0 = tj * kTileN;
      const int jmat = prod * int(p.ld);                           // the count matrices are stacked along j
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * kTileN);
#pragma unroll 1
      for (int c = 0; c < ncols; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(c), v);
        tmem_ld_wait();
        if (c + 32 == ncols) {   // accumulator drained: hand it back (to the leader's barrier)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        }
        if (j0 + c < p.n && i0 < p.n) {                            // warp-uniform
          if (lane == 0) bulk_wait_group_read<1>();                // the staging box used two chunks ago has been read
          __syncwarp();
          int32_t* box = reinterpret_cast<int32_t*>(ebase + buf * 4096);   // [32 columns][32 rows]
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) box[jj * 32 + lane] = int32_t(v[jj]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tm_counts, box, i0, jmat + j0 + c);
            bulk_commit_group();
          }
          buf ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_group<0>();   // every reduction of this warp has landed before the CTA exits
    __syncwarp();
  }

  // No CTA may leave while its partner can still signal its barriers or read its smem.
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ 4-bit variant (kind::mxf4)
// Every plane value is -1, 0 or +1: exactly representable in E2M1, and the block-scaled 4-bit MMA
// (scales = 1) issues twice the MACs per cycle of kind::i8 with half the operand bytes.  Its FP32
// accumulator keeps integer sums exact below 2^24 (known-answer experiment: scripts/mxf4_probe.cu,
// profiles/r01_mxf4_probe.json); one accumulation chain here is at most a chunk (<= 2^20 loci).
// Same CTA-pair skeleton as gram_i8_pair_kernel with K-major 4-bit planes (decode_planes_k4): a
// stage is 128 bytes = 256 loci of 128 A rows + 128 B rows per CTA, four K = 64 MMAs per stage.
// TMEM also has to hold the scale factors, so a tile is 256 x 240 (two 240-column accumulators at
// columns 0 and 256; columns [240, 256) of every lane are filled with UE8M0 1.0 and serve as SFA /
// SFB whatever their layout).  The epilogue converts the (integer-valued) floats and adds them to
// the same int32 count matrices as the int8 kernels, so chunks of both kinds can be mixed.
constexpr int kQTileN = 240;
constexpr int kQStageLoci = 256;

template <int kDummy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGramThreads, 1)
gram_mxf4_pair_kernel(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_nm,
                      const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_counts,
                      const GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* epi = smem + kPairStages * kPairStageBytes;   // [4 warps][2][32 x 32 int32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kPairEpiBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kPairStages;
  uint64_t* tfull_bar = bars + 2 * kPairStages;
  uint64_t* tempty_bar = bars + 2 * kPairStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_d);
    tma_prefetch_desc(&tm_nm);
    tma_prefetch_desc(&tm_h1);
    tma_prefetch_desc(&tm_counts);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 2) {  // scale factors: UE8M0 1.0 (0x7F) in columns [224, 256) of this warp's lanes;
                    // [224, 240) belongs to accumulator 0 and is overwritten by its first MMA
    tmem_fill_32x32(tmem_base + (uint32_t((warp & 3) * 32) << 16) + 224u, 0x7F7F7F7Fu);
    tmem_st_wait();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const int prod = w.prod_ti & 15, ti = w.prod_ti >> 8;
        const int ncols = (((w.prod_ti >> 4) & 15) + 1) * 16;        // N of this tile (240, or 64 at a region's edge)
        int pa, pb;
        prod_planes(prod, pa, pb);
        const CUtensorMap* ma = pa == 0 ? &tm_d : pa == 1 ? &tm_nm : &tm_h1;
        const CUtensorMap* mb = pb == 0 ? &tm_d : pb == 1 ? &tm_nm : &tm_h1;
        const int row_a = ti * kPairTileM + int(rank) * 128;
        const int row_b = w.tj + int(rank) * (ncols / 2);            // w.tj = first column; N/2 of the 128 loaded rows are used
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          uint8_t* sa = smem + stage * kPairStageBytes;
          uint8_t* sb = sa + kHalfBytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPairStageBytes);
          const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair(sa, ma, bar0, kb * 128, row_a);   // (byte along K, row)
          tma_load_2d_pair(sb, mb, bar0, kb * 128, row_b);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc0 = make_idesc_mxf4(kPairTileM, 0);
      const uint32_t sfa = tmem_base + 240u, sfb = tmem_base + 244u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const uint32_t ncols = (((uint32_t(w.prod_ti) >> 4) & 15u) + 1u) * 16u;
        const uint32_t idesc = idesc0 | ((ncols >> 3) << 17);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, p.err_flag, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * 256);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa, 16u, 1024u);
          const uint64_t db = make_smem_desc_sw128(sa + kHalfBytes, 16u, 1024u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 32 bytes = 64 loci per MMA
            umma_mxf4_pair(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc,
                           (kb > w.kb0 || k > 0) ? 1u : 0u, sfa, sfb);
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..5, both CTAs)
    // TMA reductions of 32 x 32 boxes, as in gram_i8_pair_kernel (the values are converted to int32 first)
    const int q = warp & 3;
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    uint8_t* ebase = epi + q * 2 * 4096;
    for (int it = cluster; it < p.n_items; it += n_clusters) {
      const WorkItem w = p.items[it];
      const int prod = w.prod_ti & 15, ti = w.prod_ti >> 8;
      const int ncols = (((w.prod_ti >> 4) & 15) + 1) * 16;   // >= 32
      mbar_wait(&tfull_bar[acc], acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int i0 = ti * kPairTileM + int(rank) * 128 + q * 32;   // this warp's 32 rows
      const int j0 = w.tj;
      const int jmat = prod * int(p.ld);
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * 256);
#pragma unroll 1
      for (int c = 0; c < ncols; c += 32) {
        const bool last = c + 32 >= ncols;
        const int cc = last ? ncols - 32 : c;       // the last 32-column load overlaps the previous one
        const int first = last ? 32 - (ncols - c) : 0;
        uint32_t v[32];
        tmem_ld_32x32(taddr + uint32_t(cc), v);
        tmem_ld_wait();
        if (last) {  // accumulator drained: hand it back (to the leader's barrier)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        }
        if (j0 + cc < p.n && i0 < p.n) {                           // warp-uniform
          if (lane == 0) bulk_wait_group_read<1>();
          __syncwarp();
          int32_t* box = reinterpret_cast<int32_t*>(ebase + buf * 4096);   // [32 columns][32 rows]
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)   // columns already added by the previous (overlapping) chunk add zero
            box[jj * 32 + lane] = jj >= first ? __float2int_rn(__uint_as_float(v[jj])) : 0;
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tm_counts, box, i0, jmat + j0 + cc);
            bulk_commit_group();
          }
          buf ^= 1;
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ 4-bit variant, wide tiles
// gram_mxf4_pair_kernel is bound by the L2 -> shared-memory operand stream, not by the tensor pipe: a CTA
// stages 32 KiB per 256 loci for 128 x 240 x 256 MACs, 10 - 12 TB/s over the whole chip — the measured
// ceiling of the L2 slices (ncu: lts throughput at its cap, tensor pipe 75 % active).
// Fewer operand bytes per MAC is the only lever, and the bytes per MAC of an output-stationary tile are
// 1/M + 1/N.  This kernel doubles N: a work item is a 256 x 480 tile — BOTH 240-column accumulators of the
// CTA pair's TMEM belong to it — so one A stage (16 KiB per CTA) feeds two B stages: 48 KiB per 256 loci
// for twice the MACs, a quarter less L2 traffic per MAC.  The price is the epilogue, which no longer
// overlaps the next item's MMAs (there is no third accumulator to ping-pong with); with 128 k-blocks per
// item it is a few per cent, so the host uses this kernel where tiles are plentiful (launch_gram_mxf4) and
// the ping-pong kernel for small matrices.
// That only pays if the exposed epilogue is short, and red.global is not: ~1.3 cycles per lane, 80,000 cycles
// for a 128 x 480 accumulator against 123,000 of MMAs.  So this kernel's epilogue goes through the TMA instead:
// a warp converts 32 x 32 accumulator values, stores them to a 4 KiB staging box in shared memory (conflict
// free: lane = row) and ONE cp.reduce.async.bulk.tensor adds the box into the int32 count matrix in the
// memory system — 32 shared-memory stores per thread instead of 32 global atomics, two boxes per warp in flight.
// Whole boxes are added: entries below the diagonal of the symmetric products receive values too (nothing
// reads them), padding rows / columns add zeros (their planes are zero).
// Everything else — CTA pairs, barriers, scale-factor columns [240, 256), the chain-length guard — is as in
// gram_mxf4_pair_kernel.
constexpr int kWideStages = 4;
constexpr int kWideStageBytes = 3 * kHalfBytes;       // per CTA: A half + two B halves
constexpr int kWideBBoxBytes = 120 * 128;             // a B box: 120 rows x 128 bytes
constexpr int kWideEpiBytes = 4 * 2 * 4096;           // epilogue staging: per warp two 32 x 32 int32 boxes
constexpr int kWideSmemBytes = kWideStages * kWideStageBytes + kWideEpiBytes + 1024 + 256;

// work item: prod [0,4) | (n0 / 16 - 1) [4,8) | (n1 / 16) [8,12) | ti [12,32); tj = first column;
// half 0 = columns [tj, tj + n0), half 1 = columns [tj + 240, tj + 240 + n1) (n1 = 0: absent)
template <int kDummy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGramThreads, 1)
gram_mxf4_wide_kernel(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_nm,
                      const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tb_d,
                      const __grid_constant__ CUtensorMap tb_nm, const __grid_constant__ CUtensorMap tb_h1,
                      const __grid_constant__ CUtensorMap tm_counts, const GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* epi = smem + kWideStages * kWideStageBytes;  // [4 warps][2][32 x 32 int32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kWideEpiBytes);
  uint64_t* full_bar = bars;                          // [kWideStages] used in the leader only
  uint64_t* empty_bar = bars + kWideStages;           // [kWideStages] in each CTA
  uint64_t* tfull_bar = bars + 2 * kWideStages;       // [1] in each CTA
  uint64_t* tempty_bar = bars + 2 * kWideStages + 1;  // [1] used in the leader only (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWideStages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_d);
    tma_prefetch_desc(&tm_nm);
    tma_prefetch_desc(&tm_h1);
    tma_prefetch_desc(&tb_d);
    tma_prefetch_desc(&tb_nm);
    tma_prefetch_desc(&tb_h1);
    tma_prefetch_desc(&tm_counts);
    for (int s = 0; s < kWideStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 2) {  // scale factors: UE8M0 1.0 (0x7F) in columns [224, 256) of this warp's lanes;
                    // [224, 240) belongs to accumulator 0 and is overwritten by its first MMA
    tmem_fill_32x32(tmem_base + (uint32_t((warp & 3) * 32) << 16) + 224u, 0x7F7F7F7Fu);
    tmem_st_wait();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const int prod = w.prod_ti & 15, ti = int(uint32_t(w.prod_ti) >> 12);
        const int n0 = (((w.prod_ti >> 4) & 15) + 1) * 16, n1 = ((w.prod_ti >> 8) & 15) * 16;
        int pa, pb;
        prod_planes(prod, pa, pb);
        const CUtensorMap* ma = pa == 0 ? &tm_d : pa == 1 ? &tm_nm : &tm_h1;
        // B boxes are 120 rows (a CTA supplies N / 2 <= 120 rows of a tile), not 128: 6 % fewer B bytes over L2
        const CUtensorMap* mb = pb == 0 ? &tb_d : pb == 1 ? &tb_nm : &tb_h1;
        const int row_a = ti * kPairTileM + int(rank) * 128;
        const int row_b0 = w.tj + int(rank) * (n0 / 2);
        const int row_b1 = w.tj + kQTileN + int(rank) * (n1 / 2);
        const uint32_t bytes = 2u * uint32_t(kHalfBytes + (n1 ? 2 : 1) * kWideBBoxBytes);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          uint8_t* sa = smem + stage * kWideStageBytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], bytes);
          const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair(sa, ma, bar0, kb * 128, row_a);   // (byte along K, row)
          tma_load_2d_pair(sa + kHalfBytes, mb, bar0, kb * 128, row_b0);
          if (n1) tma_load_2d_pair(sa + 2 * kHalfBytes, mb, bar0, kb * 128, row_b1);
          if (++stage == kWideStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc0 = make_idesc_mxf4(kPairTileM, 0);
      const uint32_t sfa = tmem_base + 240u, sfb = tmem_base + 244u;
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const uint32_t n0 = (((uint32_t(w.prod_ti) >> 4) & 15u) + 1u) * 16u, n1 = ((uint32_t(w.prod_ti) >> 8) & 15u) * 16u;
        const uint32_t idesc_a = idesc0 | ((n0 >> 3) << 17), idesc_b = idesc0 | ((n1 >> 3) << 17);
        mbar_wait(tempty_bar, acc_phase ^ 1, p.err_flag, 2);   // both accumulators drained by the previous item
        tc_fence_after();
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kWideStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa, 16u, 1024u);
          const uint64_t db0 = make_smem_desc_sw128(sa + kHalfBytes, 16u, 1024u);
          const uint64_t db1 = make_smem_desc_sw128(sa + 2 * kHalfBytes, 16u, 1024u);
          const uint32_t accum0 = kb > w.kb0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 32 bytes = 64 loci per MMA; the A step is used twice in a row
            umma_mxf4_pair(tmem_base, da + uint64_t(k * 2), db0 + uint64_t(k * 2), idesc_a, (accum0 | uint32_t(k)) ? 1u : 0u, sfa, sfb);
            if (n1) umma_mxf4_pair(tmem_base + 256u, da + uint64_t(k * 2), db1 + uint64_t(k * 2), idesc_b, (accum0 | uint32_t(k)) ? 1u : 0u, sfa, sfb);
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == kWideStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tfull_bar, 3);
        acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..5, both CTAs)
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    int buf = 0;
    uint8_t* ebase = epi + q * 2 * 4096;
    for (int it = cluster; it < p.n_items; it += n_clusters) {
      const WorkItem w = p.items[it];
      const int prod = w.prod_ti & 15, ti = int(uint32_t(w.prod_ti) >> 12);
      const int n0 = (((w.prod_ti >> 4) & 15) + 1) * 16, n1 = ((w.prod_ti >> 8) & 15) * 16;   // each >= 32 or absent
      mbar_wait(tfull_bar, acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int i0 = ti * kPairTileM + int(rank) * 128 + q * 32;       // this warp's 32 rows
      const int jmat = prod * int(p.ld);                                // count matrices are stacked along j
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int ncols = half ? n1 : n0;
        if (ncols == 0) break;
        const int j0 = w.tj + half * kQTileN;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(half * 256);
        const bool last_half = half == 1 || n1 == 0;
#pragma unroll 1
        for (int c = 0; c < ncols; c += 32) {
          const bool last = c + 32 >= ncols;
          const int cc = last ? ncols - 32 : c;       // the last 32-column load overlaps the previous one
          const int first = last ? 32 - (ncols - c) : 0;
          uint32_t v[32];
          tmem_ld_32x32(taddr + uint32_t(cc), v);
          tmem_ld_wait();
          if (last && last_half) {  // both accumulators drained: hand them back (to the leader's barrier)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar), 0));
          }
          // the staging box this warp used two chunks ago must have been read by its reduction
          if (lane == 0) bulk_wait_group_read<1>();
          __syncwarp();
          int32_t* box = reinterpret_cast<int32_t*>(ebase + buf * 4096);   // [32 columns][32 rows]
#pragma unroll
          for (int jj = 0; jj < 32; ++jj)   // columns already added by the previous (overlapping) chunk add zero
            box[jj * 32 + lane] = jj >= first ? __float2int_rn(__uint_as_float(v[jj])) : 0;
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tm_counts, box, i0, jmat + j0 + cc);
            bulk_commit_group();
          }
          buf ^= 1;
        }
      }
      acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_group<0>();   // every reduction of this warp has landed before the CTA exits
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ int8 variant, wide tiles
// The int8 -> s32 twin of gram_mxf4_wide_kernel (north_star's literal route): a work item is a 256 x 512 tile — both
// 256-column s32 accumulators of the CTA pair's TMEM — so one A stage feeds two B stages (48 KiB per 128 loci and CTA
// for twice the MACs), and the epilogue adds 32 x 32 int32 boxes with TMA reductions.  The ping-pong pair kernel reads
// 10.9 TB from DRAM per C4 launch (5.6 TB/s: HBM-bound, ncu); per MAC this kernel moves a quarter less.
// work item: prod [0,4) | (n0 / 16 - 1) [4,8) | (n1 / 16) [8,13) | ti [13,32); tj = first 256-column tile;
// half 0 = columns [256 tj, 256 tj + n0), half 1 = columns [256 (tj + 1), 256 (tj + 1) + n1) (n1 = 0: absent)
template <bool kMajorK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGramThreads, 1)
gram_i8_wide_kernel(const __grid_constant__ CUtensorMap tm_d, const __grid_constant__ CUtensorMap tm_nm,
                    const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_counts,
                    const GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* epi = smem + kWideStages * kWideStageBytes;  // [4 warps][2][32 x 32 int32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kWideEpiBytes);
  uint64_t* full_bar = bars;                          // [kWideStages] used in the leader only
  uint64_t* empty_bar = bars + kWideStages;           // [kWideStages] in each CTA
  uint64_t* tfull_bar = bars + 2 * kWideStages;       // [1] in each CTA
  uint64_t* tempty_bar = bars + 2 * kWideStages + 1;  // [1] used in the leader only (8 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWideStages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_d);
    tma_prefetch_desc(&tm_nm);
    tma_prefetch_desc(&tm_h1);
    tma_prefetch_desc(&tm_counts);
    for (int s = 0; s < kWideStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const int prod = w.prod_ti & 15, ti = int(uint32_t(w.prod_ti) >> 13), tj = w.tj;
        const int n0 = (((w.prod_ti >> 4) & 15) + 1) * 16, n1 = ((w.prod_ti >> 8) & 31) * 16;
        int pa, pb;
        prod_planes(prod, pa, pb);
        const CUtensorMap* ma = pa == 0 ? &tm_d : pa == 1 ? &tm_nm : &tm_h1;
        const CUtensorMap* mb = pb == 0 ? &tm_d : pb == 1 ? &tm_nm : &tm_h1;
        const int row_a = ti * kPairTileM + int(rank) * 128;
        const int row_b0 = tj * kTileN + int(rank) * (n0 / 2);          // N/2 of the 128 loaded rows are used
        const int row_b1 = (tj + 1) * kTileN + int(rank) * (n1 / 2);
        const uint32_t bytes = 2u * uint32_t(n1 ? kWideStageBytes : kPairStageBytes);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err_flag, 1);
          uint8_t* sa = smem + stage * kWideStageBytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], bytes);
          const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (kMajorK) {
            tma_load_2d_pair(sa, ma, bar0, kb * kTileK, row_a);
            tma_load_2d_pair(sa + kHalfBytes, mb, bar0, kb * kTileK, row_b0);
            if (n1) tma_load_2d_pair(sa + 2 * kHalfBytes, mb, bar0, kb * kTileK, row_b1);
          } else {
            tma_load_2d_pair(sa, ma, bar0, row_a, kb * kTileK);
            tma_load_2d_pair(sa + kHalfBytes, mb, bar0, row_b0, kb * kTileK);
            if (n1) tma_load_2d_pair(sa + 2 * kHalfBytes, mb, bar0, row_b1, kb * kTileK);
          }
          if (++stage == kWideStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc0 = make_idesc_i8(kPairTileM, 0, !kMajorK);
      constexpr uint32_t lbo = kMajorK ? 16u : uint32_t(kHalfBytes);
      constexpr uint32_t sbo = 1024u;
      constexpr uint32_t kstep = kMajorK ? (32u >> 4) : (4096u >> 4);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int it = cluster; it < p.n_items; it += n_clusters) {
        const WorkItem w = p.items[it];
        const uint32_t n0 = (((uint32_t(w.prod_ti) >> 4) & 15u) + 1u) * 16u, n1 = ((uint32_t(w.prod_ti) >> 8) & 31u) * 16u;
        const uint32_t idesc_a = idesc0 | ((n0 >> 3) << 17), idesc_b = idesc0 | ((n1 >> 3) << 17);
        mbar_wait(tempty_bar, acc_phase ^ 1, p.err_flag, 2);   // both accumulators drained by the previous item
        tc_fence_after();
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kWideStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa, lbo, sbo);
          const uint64_t db0 = make_smem_desc_sw128(sa + kHalfBytes, lbo, sbo);
          const uint64_t db1 = make_smem_desc_sw128(sa + 2 * kHalfBytes, lbo, sbo);
          const uint32_t accum0 = kb > w.kb0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < kTileK / 32; ++k) {
            umma_i8_pair(tmem_base, da + uint64_t(k * kstep), db0 + uint64_t(k * kstep), idesc_a, (accum0 | uint32_t(k)) ? 1u : 0u);
            if (n1) umma_i8_pair(tmem_base + 256u, da + uint64_t(k * kstep), db1 + uint64_t(k * kstep), idesc_b, (accum0 | uint32_t(k)) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == kWideStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(tfull_bar, 3);
        acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (warps 2..5, both CTAs)
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    int buf = 0;
    uint8_t* ebase = epi + q * 2 * 4096;
    for (int it = cluster; it < p.n_items; it += n_clusters) {
      const WorkItem w = p.items[it];
      const int prod = w.prod_ti & 15, ti = int(uint32_t(w.prod_ti) >> 13);
      const int n0 = (((w.prod_ti >> 4) & 15) + 1) * 16, n1 = ((w.prod_ti >> 8) & 31) * 16;   // multiples of 32 (host)
      mbar_wait(tfull_bar, acc_phase, p.err_flag, 4);
      tc_fence_after();
      const int i0 = ti * kPairTileM + int(rank) * 128 + q * 32;       // this warp's 32 rows
      const int jmat = prod * int(p.ld);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int ncols = half ? n1 : n0;
        if (ncols == 0) break;
        const int j0 = (w.tj + half) * kTileN;
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(half * 256);
        const bool last_half = half == 1 || n1 == 0;
#pragma unroll 1
        for (int c = 0; c < ncols; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + uint32_t(c), v);
          tmem_ld_wait();
          if (c + 32 == ncols && last_half) {  // both accumulators drained: hand them back (to the leader's barrier)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar), 0));
          }
          if (lane == 0) bulk_wait_group_read<1>();   // the staging box used two chunks ago has been read
          __syncwarp();
          int32_t* box = reinterpret_cast<int32_t*>(ebase + buf * 4096);   // [32 columns][32 rows]
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) box[jj * 32 + lane] = int32_t(v[jj]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tm_counts, box, i0, jmat + j0 + c);
            bulk_commit_group();
          }
          buf ^= 1;
        }
      }
      acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_group<0>();
    __syncwarp();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_plane_map(Ctx* ctx, CUtensorMap* out, const int8_t* base, uint64_t inner,
                          uint64_t outer, const Err& E, uint32_t box_rows = 128) {
  if (!ctx->encode_tiled) return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled entry point missing");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {inner};  // bytes, dim 1
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<int8_t*>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
  return TPG_OK;
}

// The count matrices as one 2-D int32 tensor (rows i fastest, the four matrices stacked along j): the epilogues' TMA
// reductions add 32 x 32 boxes into it.
static int make_counts_map(Ctx* ctx, CUtensorMap* out, int32_t* counts, int64_t ld, const Err& E) {
  if (!ctx->encode_tiled) return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled entry point missing");
  cuuint64_t dims[2] = {cuuint64_t(ld), cuuint64_t(ld) * TPG_NUM_COUNT_MATS};
  cuuint64_t strides[1] = {cuuint64_t(ld) * sizeof(int32_t)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      out, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, counts, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return E.fail(TPG_ERR_CUDA, "cuTensorMapEncodeTiled(counts) failed (CUresult %d)", int(r));
  return TPG_OK;
}

// The persistent CTA pairs take work items round-robin (item = pair index + k * pairs), and they share operand panels
// through L2 only while they walk the list side by side.  Items of unequal cost break that: a pair that drew a few
// narrow items runs ahead, the resident items spread over more neighbourhoods than L2 holds.  (Measured, int8 wide
// kernel at C4, one of 2 tile-sharded ranks: a 3840-column region is 7 items of 512 columns + 1 of 256, and the rank
// ran 28 % slower per MAC than the unsharded job, whose items are all alike.)  So the list is cut into runs of equal
// cost: all full-width items first, in their L2 order, then the narrower ones by width — within a run every pair
// finishes its item at the same time again.
template <class T>
static void equal_cost_runs(std::vector<T>& tiles) {
  int full = 0;
  for (const T& t : tiles) full = std::max(full, t.n0 + t.n1);
  std::stable_sort(tiles.begin(), tiles.end(), [full](const T& x, const T& y) {
    const int wx = x.n0 + x.n1, wy = y.n0 + y.n1;
    return (wx == full) != (wy == full) ? wx == full : wx > wy;
  });
}

// Chooses how many k-splits a tile gets so that the persistent CTAs finish together.
static int choose_ksplit(int64_t n_tile_items, int64_t kblocks, int sms, int default_max_kblocks = 128) {
  const tpg_config C = cfg();
  if (C.gram_ksplit > 0) return int(std::max<int64_t>(1, std::min<int64_t>(kblocks, int64_t(C.gram_ksplit))));
  // Large problems: every CTA runs many items, balance is not the issue — L2 is.  CTAs that
  // run side by side share operand panels only while they stream the same loci; over a long
  // k range they drift apart and every CTA ends up fetching its panels from HBM by itself
  // (measured, KING: n = 20,000 x 100,000 loci in one item: 110 ms; items of 256 k-blocks: 50 ms;
  // n = 50,000 x 500,000: 256 k-blocks 3.25 s, 128 k-blocks 1.71 s).  With items of 128 k-blocks
  // (16,384 loci) the operand panels of a whole 1024 x 1024 tile region (~100 MB) fit in the
  // 126 MB L2, so the reuse no longer depends on the CTAs staying in lock step.
  const int max_per = C.gram_max_kblocks > 0 ? std::max(8, int(C.gram_max_kblocks)) : default_max_kblocks;
  const int ks_l2 = int((kblocks + max_per - 1) / max_per);
  if (n_tile_items >= int64_t(8) * sms) return std::max(1, ks_l2);
  int best = 1;
  double best_cost = 1e300;
  const int max_ks = int(std::min<int64_t>(kblocks, 512));
  for (int ks = 1; ks <= max_ks; ++ks) {
    const int64_t per = (kblocks + ks - 1) / ks;      // k-blocks per item
    if (ks > 1 && per < 8) break;                     // keep the epilogue amortised
    const int64_t waves = (n_tile_items * ks + sms - 1) / sms;
    const double cost = double(waves) * (double(per) + 2.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ks; }
  }
  return std::max(best, std::min(ks_l2, max_ks));
}

// Block-diagonal jobs: does the tile rows [r0, r1) x columns [c0, c1) hold a pair of individuals of the same
// group?  Group ids are non-decreasing along the (sorted) individuals, so it is an interval overlap.
static bool tile_wanted(const GramArgs& a, int64_t r0, int64_t r1, int64_t c0, int64_t c1) {
  if (!a.group_sorted) return true;
  const int64_t last = a.n - 1;
  if (r0 > last || c0 > last) return false;
  const int32_t gr0 = a.group_sorted[r0], gr1 = a.group_sorted[std::min(r1 - 1, last)];
  const int32_t gc0 = a.group_sorted[c0], gc1 = a.group_sorted[std::min(c1 - 1, last)];
  return gr0 <= gc1 && gc0 <= gr1;
}

// Host side of the 4-bit variant: 256 x 240 tiles over K-major 4-bit planes, k-blocks of 256 loci.
static int launch_gram_mxf4(Ctx* ctx, const GramArgs& a, cudaStream_t st, int64_t* macs_out, const Err& E) {
  if (a.n_pad % kRowPad || a.lc_pad % kQStageLoci || a.lc_pad <= 0)
    return E.fail(TPG_ERR_ARG, "gram(mxf4): n_pad (%lld) / lc_pad (%lld) not multiples of 256",
                  (long long)a.n_pad, (long long)a.lc_pad);
  const int64_t nblocks = a.n_pad / 256;
  const int64_t region_w = a.shard_world > 1 ? 256 * shard_region_blocks(nblocks, a.shard_world) : 0;
  if (a.shard_world > 1 && region_w < 1024)
    return E.fail(TPG_ERR_ARG, "gram(mxf4): tile sharding needs ownership regions of at least 1024 columns (int8 kernels otherwise)");
  if (!ctx->gram_attr_set) {
    GramArgs probe = a;   // sets the attributes of every gram kernel and counts the resident CTA pairs
    probe.mxf4 = false;
    probe.prod_mask = 0;
    int rc = launch_gram(ctx, probe, st, nullptr, E);
    if (rc) return rc;
  }
  if (!ctx->mxf4_attr_set) {
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_mxf4_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_mxf4_wide_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmemBytes), E);
    ctx->mxf4_attr_set = true;
  }
  if (ctx->pair_clusters <= 0) return E.fail(TPG_ERR_CUDA, "gram(mxf4): the device reports no resident CTA pairs");
  const int units = ctx->pair_clusters;
  const int64_t tiles_m = a.n_pad / kPairTileM;
  const int64_t kblocks = a.lc_pad / kQStageLoci;
  // Column tiles never straddle an ownership region: a 3840-column region is sixteen 240-column tiles, a
  // 1024-column one 240 + 240 + 240 + 240 + 64; unsharded jobs simply run 240-wide tiles across the whole width.
  // The last tile of the matrix is trimmed to the valid individuals (multiples of 16, at least 32 for the epilogue).
  std::vector<std::pair<int, int>> cols;  // (first column, N)
  size_t tiles_per_region = 0;
  if (a.shard_world > 1) {
    tiles_per_region = size_t((region_w + kQTileN - 1) / kQTileN);
    for (int64_t r0 = 0; r0 < a.n; r0 += region_w)
      for (int64_t c0 = r0; c0 < r0 + region_w && c0 < a.n; c0 += kQTileN) {
        const int64_t room = std::min<int64_t>(kQTileN, r0 + region_w - c0);
        const int64_t need = std::max<int64_t>(32, (a.n - c0 + 15) / 16 * 16);
        cols.push_back({int(c0), int(std::min(room, need))});
      }
  } else {
    // the last column tile is trimmed to the valid individuals (multiples of 16, at least 32 for the epilogue)
    const bool trim = !cfg().gram_no_trim;
    for (int64_t c0 = 0; c0 < a.n; c0 += kQTileN) {
      const int64_t need = std::max<int64_t>(32, (a.n - c0 + 15) / 16 * 16);
      cols.push_back({int(c0), trim ? int(std::min<int64_t>(kQTileN, need)) : kQTileN});
    }
  }
  // Wide tiles (gram_mxf4_wide_kernel): two neighbouring 240-column tiles share one A stream.  Worth it where
  // tiles are plentiful (the kernel is then bound by the L2 -> smem operand stream); small matrices keep the
  // ping-pong kernel, whose epilogue overlaps the next item.  tpg_config.gram_wide: 0 auto, 1 never, 2 always.
  struct WC { int c0, n0, n1; };
  std::vector<WC> wcols;
  auto count_tiles = [&](bool wide_cols) {
    int64_t cnt = 0;
    for (int64_t ti = 0; ti < tiles_m; ++ti) {
      if (kPairTileM * ti >= a.n) continue;
      if (wide_cols) { for (const WC& w : wcols) if (w.c0 + (w.n1 ? kQTileN + w.n1 : w.n0) - 1 >= kPairTileM * ti) ++cnt; }
      else { for (const auto& c : cols) if (c.first + c.second - 1 >= kPairTileM * ti) ++cnt; }
    }
    return cnt;
  };
  const tpg_config C = cfg();
  bool wide = C.gram_wide == 2 || (C.gram_wide == 0 && count_tiles(false) >= int64_t(4) * units);
  if (wide) {
    // pair up: unsharded 240 | 240 | 240 ...; sharded, per 1024-column region, (240 + 240) (240 + 240) (64)
    // (a pair never spans two ownership regions)
    for (size_t k = 0; k < cols.size();) {
        const bool same_region = !region_w || (k + 1 < cols.size() && cols[k].first / region_w == cols[k + 1].first / region_w);
        const bool pair2 = k + 1 < cols.size() && same_region && cols[k].second == kQTileN &&
                           cols[k + 1].first == cols[k].first + kQTileN;
        wcols.push_back({cols[k].first, cols[k].second, pair2 ? cols[k + 1].second : 0});
        k += pair2 ? 2 : 1;
    }
  } else {
    for (const auto& c : cols) wcols.push_back({c.first, c.second, 0});
  }
  // L2 neighbourhoods (tpg_config.gram_l2_rows x gram_l2_cols, in 256-row / 240-column tiles; default 2048 x 1920)
  // and, inside one, PRODUCT-major order: the CTA pairs resident together stream the panels of ONE pair of planes
  // (2048 + 1920 rows x 16 KiB per 128-k-block item = 65 MB: L2-resident) instead of all three planes of a
  // 1024 x 1024 neighbourhood (round 1's order; ncu at C4: 3.4 TB of DRAM reads for 37.6 GB of planes).  For this
  // kernel the order turned out not to matter (scripts/order_sweep.py: 798 - 806 ms for every variant at C4): it is
  // bound by the L2 -> shared-memory stream and the power cap, not by HBM.
  // Two regimes, told apart by how many waves of work items one k-split is (estimated before the list is built):
  // a biobank-sized matrix (C4: ~550 waves per k-split) runs best product-major in 2048 x 1920 neighbourhoods with
  // long items (256 k-blocks: the wide kernel's exposed epilogue is amortised; 803 -> 749 ms), a mid-sized one (C5,
  // 10,000 individuals: ~18 waves) with round 1's tile-major 1024 x 960 neighbourhoods and 128 k-blocks (46 vs 51 - 53
  // ms) — scripts/gram_sweep_c5.py, profiles/r02_gram_sweep_c5.jsonl.
  int nprod = 0;
  for (int prod = 0; prod < TPG_NUM_COUNT_MATS; ++prod) nprod += ((a.prod_mask >> prod) & 1) * (prod == TPG_MAT_HN ? 2 : 1);
  const int64_t waves_est = count_tiles(wide) * nprod / std::max(1, a.shard_world) / std::max(1, units);
  const bool big = waves_est >= 50;
  const int64_t RI = std::max<int64_t>(1, C.gram_l2_rows > 0 ? C.gram_l2_rows : (big ? 8 : 4));
  const size_t rj240 = size_t(std::max(1, C.gram_l2_cols > 0 ? int(C.gram_l2_cols) : (big ? 8 : 4)));
  const size_t RJ = wide ? std::max<size_t>(1, rj240 / 2) : rj240;
  const bool prod_major = C.gram_order == 0 ? big : C.gram_order == 1;
  struct T { int prod, ti, col0, n0, n1; };
  std::vector<T> tiles;
  for (size_t rj = 0; rj < wcols.size(); rj += RJ)
    for (int64_t ri = 0; ri < tiles_m; ri += RI)
      for (int pp = 0; pp < (prod_major ? TPG_NUM_COUNT_MATS : 1); ++pp)
      for (size_t cj = rj; cj < std::min(rj + RJ, wcols.size()); ++cj)
        for (int64_t ti = ri; ti < std::min(ri + RI, tiles_m); ++ti) {
          if (kPairTileM * ti >= a.n) continue;  // rows of padding individuals only
          const int c0 = wcols[cj].c0, n0 = wcols[cj].n0, n1 = wcols[cj].n1;
          const int c_end = c0 + (n1 ? kQTileN + n1 : n0);
          const bool upper = c_end - 1 >= kPairTileM * ti;
          if (!tile_wanted(a, kPairTileM * ti, kPairTileM * (ti + 1), c0, c_end)) continue;
          const int64_t bi = ti, bj = c0 / 256;  // 256-blocks; a tile never straddles an ownership region, so any block of it will do
          for (int prod = prod_major ? pp : 0; prod < (prod_major ? pp + 1 : TPG_NUM_COUNT_MATS); ++prod) {
            if (!(a.prod_mask & (1 << prod))) continue;
            if (prod != TPG_MAT_HN && !upper) continue;
            if (!block_owned(std::min(bi, bj), std::max(bi, bj), nblocks, a.shard_rank, a.shard_world, a.owner)) continue;
            tiles.push_back({prod, int(ti), c0, n0, n1});
          }
        }
  if (tiles.empty()) { if (macs_out) *macs_out = 0; return TPG_OK; }
  if (wide) equal_cost_runs(tiles);
  // (the wide kernel's epilogue is exposed, so its items are longer: 256 k-blocks = 65,536 loci; C4 sweep,
  //  profiles/r02_gram_sweep.jsonl: 96 -> 810 ms, 128 -> 789, 192 -> 780, 256 -> 772, 384 -> 769)
  int ks = choose_ksplit(int64_t(tiles.size()), kblocks, units, (wide && big) ? 256 : 128);
  // EXACTNESS GUARD.  The FP32 accumulator of kind::mxf4 holds integer sums exactly only below 2^24
  // (gram_mxf4_pair_kernel): one work item = one accumulation chain never runs more than
  // TPG_MXF4_MAX_CHAIN_LOCI (2^23) loci, whatever the chunk size, the auto heuristic or a gram_ksplit
  // override asked for — longer k ranges are split, the int32 atomics combine the parts exactly.
  constexpr int64_t kMaxChainKb = TPG_MXF4_MAX_CHAIN_LOCI / kQStageLoci;
  if ((kblocks + ks - 1) / ks > kMaxChainKb) ks = int((kblocks + kMaxChainKb - 1) / kMaxChainKb);
  if (a.max_chain_out) {
    int64_t longest = 0;
    for (int s2 = 0; s2 < ks; ++s2) longest = std::max<int64_t>(longest, kblocks * (s2 + 1) / ks - kblocks * s2 / ks);
    *a.max_chain_out = std::max(*a.max_chain_out, longest * kQStageLoci);
  }
  if (macs_out) {
    int64_t cells = 0;
    for (const T& t : tiles) cells += int64_t(kPairTileM) * (t.n0 + t.n1);
    *macs_out = cells * a.lc_pad;
  }
  char key[200];
  snprintf(key, sizeof key, "q4%c/%lld/%lld/%lld/%d/%d/%d/%d/%llx/%lld/%zu/%d", wide ? 'w' : 'p', (long long)a.n, (long long)a.n_pad, (long long)a.lc_pad,
           a.prod_mask, ks, a.shard_rank, a.shard_world, (unsigned long long)(a.group_sorted ? a.group_sig | 1u : 0u),
           (long long)RI, RJ, int(prod_major));
  if (ctx->items_key != key) {
    std::vector<WorkItem>& items = ctx->h_items;
    items.clear();
    items.reserve(tiles.size() * ks);
    for (int s = 0; s < ks; ++s) {
      const int32_t kb0 = int32_t(kblocks * s / ks), kb1 = int32_t(kblocks * (s + 1) / ks);
      if (kb1 <= kb0) continue;
      for (const T& t : tiles) {
        if (wide) items.push_back({int32_t(uint32_t(t.prod) | (uint32_t(t.n0 / 16 - 1) << 4) | (uint32_t(t.n1 / 16) << 8) | (uint32_t(t.ti) << 12)), t.col0, kb0, kb1});
        else items.push_back({int32_t(t.prod | ((t.n0 / 16 - 1) << 4) | (t.ti << 8)), t.col0, kb0, kb1});
      }
    }
    if (items.size() > ctx->d_items_cap) {
      TPG_CUDA_OK(cudaStreamSynchronize(st), E);
      if (ctx->d_items) cudaFree(ctx->d_items);
      ctx->d_items = nullptr;
      ctx->d_items_cap = 0;
      const size_t cap = items.size() + items.size() / 4 + 1024;
      TPG_CUDA_OK(cudaMalloc(&ctx->d_items, cap * sizeof(WorkItem)), E);
      ctx->d_items_cap = cap;
    }
    TPG_CUDA_OK(cudaMemcpyAsync(ctx->d_items, items.data(), items.size() * sizeof(WorkItem), cudaMemcpyHostToDevice, st), E);
    ctx->items_key = key;
  }
  const size_t n_items = ctx->h_items.size();
  CUtensorMap maps[3];
  for (int pl = 0; pl < 3; ++pl) {
    const int8_t* base = a.planes + int64_t(pl) * a.plane_stride;
    const bool present = pl < 2 || (a.prod_mask & ((1 << TPG_MAT_HH) | (1 << TPG_MAT_HN)));
    // K-major 4-bit planes: inner = bytes along K (lc_pad / 2), outer = rows; rows past n_pad are zero fill
    int rc = make_plane_map(ctx, &maps[pl], present ? base : a.planes, uint64_t(a.lc_pad / 2), uint64_t(a.n_pad), E);
    if (rc) return rc;
  }
  GramParams p;
  p.items = ctx->d_items;
  p.n_items = int(n_items);
  p.counts = a.counts;
  p.mat_stride = a.ld * a.ld;
  p.ld = a.ld;
  p.n = int(a.n);
  p.err_flag = ctx->d_err_flag;
  const int grid = 2 * int(std::min<size_t>(n_items, size_t(units)));
  if (wide && a.wide_out) ++*a.wide_out;
  CUtensorMap cmap;
  if (int rc = make_counts_map(ctx, &cmap, a.counts, a.ld, E)) return rc;
  if (wide) {
    CUtensorMap bmaps[3];
    for (int pl = 0; pl < 3; ++pl) {
      const bool present = pl < 2 || (a.prod_mask & ((1 << TPG_MAT_HH) | (1 << TPG_MAT_HN)));
      int rc = make_plane_map(ctx, &bmaps[pl], present ? a.planes + int64_t(pl) * a.plane_stride : a.planes,
                              uint64_t(a.lc_pad / 2), uint64_t(a.n_pad), E, 120);
      if (rc) return rc;
    }
    gram_mxf4_wide_kernel<0><<<grid, kGramThreads, kWideSmemBytes, st>>>(maps[0], maps[1], maps[2], bmaps[0], bmaps[1], bmaps[2], cmap, p);
  }
  else gram_mxf4_pair_kernel<0><<<grid, kGramThreads, kPairSmemBytes, st>>>(maps[0], maps[1], maps[2], cmap, p);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

static bool pair_mode(const GramArgs& a) {
  if (a.ctas == 1) return false;
  if (a.ctas == 2) return true;
  return cfg().gram_ctas != 1;
}

int launch_gram(Ctx* ctx, const GramArgs& a, cudaStream_t st, int64_t* macs_out, const Err& E) {
  if (a.n_pad % kRowPad || a.lc_pad % kLociPad || a.lc_pad <= 0)
    return E.fail(TPG_ERR_ARG, "gram: n_pad (%lld) / lc_pad (%lld) not tile multiples",
                  (long long)a.n_pad, (long long)a.lc_pad);
  if (a.mxf4) return launch_gram_mxf4(ctx, a, st, macs_out, E);
  const bool pair = pair_mode(a);
  const int tile_m = pair ? kPairTileM : kTileM;
  const int64_t tiles_m = a.n_pad / tile_m, tiles_n = a.n_pad / kTileN;
  const int64_t kblocks = a.lc_pad / kTileK;

  if (!ctx->gram_attr_set) {  // once per context (= per device)
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_pair_kernel<true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_pair_kernel<false>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes), E);
    // how many CTA pairs the device keeps resident at once (GPCs with an odd SM count lose one)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(ctx->num_sms / 2 * 2));
    cfg.blockDim = dim3(kGramThreads);
    cfg.dynamicSmemBytes = kPairSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, gram_i8_pair_kernel<false>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); nc = 0; }
    ctx->pair_clusters = std::min(nc, ctx->num_sms / 2);
    if (tpg::cfg().gram_clusters > 0) ctx->pair_clusters = std::max(1, std::min(int(tpg::cfg().gram_clusters), ctx->num_sms / 2));
    ctx->gram_attr_set = true;
  }
  if (pair && ctx->pair_clusters <= 0)
    return E.fail(TPG_ERR_CUDA, "gram: the device reports no resident CTA pairs for gram_i8_pair_kernel");
  // schedulable units: SMs (one CTA each) or resident CTA pairs
  const int units = pair ? ctx->pair_clusters : ctx->num_sms;

  // ---- tile list in L2-friendly order: neighbourhoods, then products, then tiles
  const tpg_config C2 = cfg();
  const bool trim_n = pair && !C2.gram_no_trim;
  // N of a tile's MMAs: the pair kernels trim the last tile column to its valid columns (multiple of 32)
  auto tile_ncols = [&](int64_t tj) {
    if (!trim_n) return int(kTileN);
    const int64_t valid = std::min<int64_t>(kTileN, a.n - tj * kTileN);
    return int(std::max<int64_t>(32, (valid + 31) / 32 * 32));
  };
  // Wide items (gram_i8_wide_kernel): two neighbouring 256-column tiles of one ownership region share an A stream.
  // Where tiles are plentiful, as for the 4-bit route (tpg_config.gram_wide: 0 auto, 1 never, 2 always).
  const int64_t upper_tiles = tiles_m * (tiles_n + 1) / 2;
  const bool wide = pair && (C2.gram_wide == 2 || (C2.gram_wide == 0 && upper_tiles >= int64_t(4) * units));
  if (wide && !ctx->i8wide_attr_set) {
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmemBytes), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gram_i8_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmemBytes), E);
    ctx->i8wide_attr_set = true;
  }
  struct JC { int tj, n0, n1; };   // column item: tile tj (n0 columns) and, when paired, tile tj + 1 (n1 columns)
  std::vector<JC> jcs;
  {
    const int64_t region = a.shard_world > 1 ? shard_region_blocks(tiles_n, a.shard_world) : tiles_n;
    for (int64_t tj = 0; tj < tiles_n;) {
      const bool pair2 = wide && tj + 1 < tiles_n && tj / region == (tj + 1) / region && tile_ncols(tj) == kTileN &&
                         kTileN * (tj + 1) < a.n;
      jcs.push_back({int(tj), tile_ncols(tj), pair2 ? tile_ncols(tj + 1) : 0});
      tj += pair2 ? 2 : 1;
    }
  }
  struct T { int prod, ti, tj, n0, n1; };
  std::vector<T> tiles;
  // L2 neighbourhoods (tpg_config.gram_l2_rows / gram_l2_cols, in 256-row / 256-column tiles; default 1024 x 1024)
  // and, inside one, PRODUCT-major order: CTAs resident together then stream the panels of ONE pair of planes —
  // 2 x 1024 rows x 16 KiB per 128-k-block item = 33 MB, L2-resident — instead of all three planes of the
  // neighbourhood at once (round 1's order: tile-major, products innermost; ncu at C4: 13 - 16 TB of DRAM reads for
  // 75 GB of int8 planes, the kernel ran at 5.4 - 6.3 TB/s of HBM).  Measured at C4 (scripts/order_sweep.py,
  // profiles/r02_order_sweep.jsonl): 2174 -> 1939 ms; larger neighbourhoods are slower again (2048 x 2048: 2302 ms).
  const int64_t RI = std::max<int64_t>(1, (C2.gram_l2_rows > 0 ? C2.gram_l2_rows : 4) * 256 / tile_m);
  const int64_t rj256 = std::max<int64_t>(1, C2.gram_l2_cols > 0 ? C2.gram_l2_cols : 4);
  const size_t RJ = size_t(wide ? std::max<int64_t>(1, rj256 / 2) : rj256);
  const bool prod_major = C2.gram_order != 2;
  for (size_t rj = 0; rj < jcs.size(); rj += RJ)
    for (int64_t ri = 0; ri < tiles_m; ri += RI)
      for (int pp = 0; pp < (prod_major ? TPG_NUM_COUNT_MATS : 1); ++pp)
      for (size_t cj = rj; cj < std::min(rj + RJ, jcs.size()); ++cj)
        for (int64_t ti = ri; ti < std::min(ri + RI, tiles_m); ++ti) {
          const JC& jc = jcs[cj];
          const int64_t tj = jc.tj;
          const int64_t c_end = jc.n1 ? kTileN * (tj + 1) + jc.n1 : kTileN * tj + kTileN;
          const bool upper = c_end - 1 >= tile_m * ti;
          if (!tile_wanted(a, tile_m * ti, tile_m * (ti + 1), kTileN * tj, c_end)) continue;
          // 256x256 ownership block of the tile (a pair lies inside one region); lower blocks mirror
          const int64_t bi = ti * tile_m / 256, bj = tj;
          for (int prod = prod_major ? pp : 0; prod < (prod_major ? pp + 1 : TPG_NUM_COUNT_MATS); ++prod) {
            if (!(a.prod_mask & (1 << prod))) continue;
            if (prod != TPG_MAT_HN && !upper) continue;
            const int64_t lo = std::min(bi, bj), hi = std::max(bi, bj);
            if (!block_owned(lo, hi, tiles_n, a.shard_rank, a.shard_world, a.owner)) continue;
            tiles.push_back({prod, int(ti), int(tj), jc.n0, jc.n1});
          }
        }
  if (tiles.empty()) { if (macs_out) *macs_out = 0; return TPG_OK; }
  if (wide) equal_cost_runs(tiles);

  const int ks = choose_ksplit(int64_t(tiles.size()), kblocks, units);

  // The work list only depends on the geometry; chunks of equal size reuse the upload.
  if (macs_out) {
    int64_t cells = 0;
    for (const T& t : tiles) cells += int64_t(tile_m) * (pair ? t.n0 + t.n1 : kTileN);
    *macs_out = cells * a.lc_pad;
  }
  char key[200];
  snprintf(key, sizeof key, "%lld/%lld/%lld/%d/%d/%d/%d/%d/%d/%llx/%lld/%zu/%d/%d", (long long)a.n, (long long)a.n_pad, (long long)a.lc_pad,
           a.prod_mask, a.shard_rank, a.shard_world, ks, tile_m, int(trim_n),
           (unsigned long long)(a.group_sorted ? a.group_sig | 1u : 0u), (long long)RI, RJ, int(prod_major), int(wide));
  if (ctx->items_key != key) {
    std::vector<WorkItem>& items = ctx->h_items;
    items.clear();
    items.reserve(tiles.size() * ks);
    for (int s = 0; s < ks; ++s) {
      const int32_t kb0 = int32_t(kblocks * s / ks), kb1 = int32_t(kblocks * (s + 1) / ks);
      if (kb1 <= kb0) continue;
      for (const T& t : tiles) {
        if (!pair) { items.push_back({int32_t(t.prod | (t.ti << 4)), t.tj, kb0, kb1}); continue; }
        if (wide) {
          items.push_back({int32_t(uint32_t(t.prod) | (uint32_t(t.n0 / 16 - 1) << 4) | (uint32_t(t.n1 / 16) << 8) | (uint32_t(t.ti) << 13)), t.tj, kb0, kb1});
          continue;
        }
        // pair kernel: N of the MMAs trimmed to the valid columns of the last tile column (multiple of 32)
        items.push_back({int32_t(t.prod | ((t.n0 / 16 - 1) << 4) | (t.ti << 8)), t.tj, kb0, kb1});
      }
    }
    if (items.size() > ctx->d_items_cap) {
      TPG_CUDA_OK(cudaStreamSynchronize(st), E);
      if (ctx->d_items) cudaFree(ctx->d_items);
      ctx->d_items = nullptr;
      ctx->d_items_cap = 0;
      const size_t cap = items.size() + items.size() / 4 + 1024;
      TPG_CUDA_OK(cudaMalloc(&ctx->d_items, cap * sizeof(WorkItem)), E);
      ctx->d_items_cap = cap;
    }
    // Stream-ordered after any kernel still reading the previous list.
    TPG_CUDA_OK(cudaMemcpyAsync(ctx->d_items, items.data(), items.size() * sizeof(WorkItem),
                                cudaMemcpyHostToDevice, st), E);
    ctx->items_key = key;
  }
  const size_t n_items = ctx->h_items.size();

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
  p.items = ctx->d_items;
  p.n_items = int(n_items);
  p.counts = a.counts;
  p.mat_stride = a.ld * a.ld;
  p.ld = a.ld;
  p.n = int(a.n);
  p.err_flag = ctx->d_err_flag;

  if (pair && wide && a.wide_out) ++*a.wide_out;
  CUtensorMap cmap;
  if (pair) if (int rc = make_counts_map(ctx, &cmap, a.counts, a.ld, E)) return rc;
  if (pair && wide) {
    const int grid = 2 * int(std::min<size_t>(n_items, size_t(units)));
    if (a.k_major) gram_i8_wide_kernel<true><<<grid, kGramThreads, kWideSmemBytes, st>>>(maps[0], maps[1], maps[2], cmap, p);
    else gram_i8_wide_kernel<false><<<grid, kGramThreads, kWideSmemBytes, st>>>(maps[0], maps[1], maps[2], cmap, p);
  } else if (pair) {
    const int grid = 2 * int(std::min<size_t>(n_items, size_t(units)));  // cluster dims are compiled in
    if (a.k_major) gram_i8_pair_kernel<true><<<grid, kGramThreads, kPairSmemBytes, st>>>(maps[0], maps[1], maps[2], cmap, p);
    else gram_i8_pair_kernel<false><<<grid, kGramThreads, kPairSmemBytes, st>>>(maps[0], maps[1], maps[2], cmap, p);
  } else {
    const int grid = int(std::min<size_t>(n_items, size_t(units)));
    if (a.k_major) gram_i8_kernel<true><<<grid, kGramThreads, kSmemBytes, st>>>(maps[0], maps[1], maps[2], p);
    else gram_i8_kernel<false><<<grid, kGramThreads, kSmemBytes, st>>>(maps[0], maps[1], maps[2], p);
  }
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

}  // namespace tpg
