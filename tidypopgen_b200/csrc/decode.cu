// decode_planes: raw FBM.code256 bytes -> int8 indicator planes, HBM-bandwidth bound.
//
// Replaces the serial decode loops of the reference (and the SubBMAcc row/column gather
// in front of them):
//   src/snp_king.cpp:45-58   value==0/1/2 -> genotype0/1/2, genotype_valid
//   src/snp_ibs.cpp:45-56    same three indicators
//   src/snp_as.cpp:44-53     value<3 -> dos = value-1, na = 1; else 0, 0
// All three read the RAW byte (SubBMAcc<unsigned char>, never code256), so the rule is
// "byte <= 2 is a dosage, anything else is missing" — imputed codes 4/5/6 included.
//
// Planes written (int8): d = [byte==2] - [byte==0]  (-1/0/+1; the reference's dos-1)
//                        NM = [byte<=2]             (0/1;  genotype_valid / na_mat)
//                        H1 = [byte==1]             (0/1;  genotype1)
// Algorithmic traffic: 1 byte read + 3 bytes written per genotype (2 written for AS/GRM).
//
// Default layout is MN-major: plane[l * n_pad + i] — the same orientation as the .bk
// (individuals contiguous), so the kernel is a pure stream: each lane owns 4 consecutive
// individuals (one 32-bit word), a warp reads/writes 128 contiguous bytes per locus, and
// tcgen05 consumes the planes as MN-major operands.  Column starts in a .bk are not
// aligned (nrow is arbitrary), so loads are aligned 32-bit words re-assembled with a
// funnel shift across neighbouring lanes.  The K-major variant transposes 4x4 byte blocks
// in registers (PRMT) and writes plane[i * lc_pad + l].
#include <algorithm>
#include <cstdlib>
#include "internal.h"
#include "swar.cuh"

namespace tpg {

struct DecodeParams {
  const uint8_t* raw;
  const uint8_t* raw_end;
  int64_t raw_stride;
  const int32_t* row_idx;
  int64_t row0;
  const int32_t* col_idx;
  int64_t col0;
  int n, n_pad, lc, lc_pad;
  int8_t* planes;
  int64_t plane_stride;
  int nplanes;
};

// 4 packed genotype bytes -> 4 packed plane bytes each (SWAR, no per-byte branches).
__device__ __forceinline__ void decode4(uint32_t g, uint32_t& d, uint32_t& nm, uint32_t& h1) {
  const uint32_t t = (g & 0x7F7F7F7Fu) + 0x7D7D7D7Du;  // bit 7 set iff (byte & 127) >= 3
  nm = (((t | g) >> 7) & 0x01010101u) ^ 0x01010101u;   // 1 iff byte <= 2
  h1 = g & nm;                                         // valid bytes are 0,1,2: bit 0 <=> het
  const uint32_t h2 = (g >> 1) & nm;                   // bit 1 <=> hom-alt
  const uint32_t h0 = nm ^ h1 ^ h2;                    // the remaining valid ones
  d = h2 | (h0 * 255u);                                // +1 -> 0x01, -1 -> 0xFF
}

// Aligned 32-bit load that never touches bytes outside [raw, raw_end).
__device__ __forceinline__ uint32_t guarded_ld32(const uint8_t* p, const uint8_t* lo,
                                                 const uint8_t* hi) {
  if (p >= lo && p + 4 <= hi) return __ldg(reinterpret_cast<const uint32_t*>(p));
  uint32_t w = 0x03030303u;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (p + b >= lo && p + b < hi) w = (w & ~(0xFFu << (8 * b))) | (uint32_t(__ldg(p + b)) << (8 * b));
  return w;
}

// Returns the word holding the genotypes of individuals i0..i0+3 (this lane's four) of one
// raw column.  Contiguous rows: one aligned load per lane + shuffle + funnel shift.
// Gathered rows: four byte loads.  Individuals >= n read as 0x03 (missing).
struct RowSel {
  bool gather;
  int32_t r[4];    // gathered raw rows (gather mode)
  int64_t off;     // row0 + i0 (contiguous mode)
  uint32_t keep;   // byte mask of individuals < n
  bool ld;         // this lane's aligned word is needed (by itself or by lane-1's funnel shift)
};

// kSafe = the caller proved that every word this block can touch lies inside the raw buffer,
// so the loads are unconditional and branch-free (8-16 of them stay in flight per lane).
template <bool kSafe>
__device__ __forceinline__ uint32_t load4(const DecodeParams& p, const RowSel& rs, int64_t col,
                                          int lane) {
  const uint8_t* cbase = p.raw + col * p.raw_stride;
  uint32_t w;
  if (rs.gather) {
    w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (rs.keep & (0xFFu << (8 * b))) w |= uint32_t(__ldg(cbase + rs.r[b])) << (8 * b);
  } else {
    const uint8_t* a = cbase + rs.off;
    const uint32_t sh = uint32_t(reinterpret_cast<uintptr_t>(a) & 3u);
    const uint8_t* al = a - sh;
    uint32_t w0, w1;
    if (kSafe) {
      w0 = __ldg(reinterpret_cast<const uint32_t*>(al));
      w1 = __shfl_down_sync(0xFFFFFFFFu, w0, 1);
      if (lane == 31) w1 = __ldg(reinterpret_cast<const uint32_t*>(al) + 1);
    } else {
      // Lanes far beyond the last individual skip the load (their bytes are masked anyway).
      w0 = rs.ld ? guarded_ld32(al, p.raw, p.raw_end) : 0u;
      w1 = __shfl_down_sync(0xFFFFFFFFu, w0, 1);
      if (lane == 31 && sh && rs.keep) w1 = guarded_ld32(al + 4, p.raw, p.raw_end);
    }
    w = __funnelshift_r(w0, w1, sh * 8);
  }
  return (w & rs.keep) | (0x03030303u & ~rs.keep);
}

// True when all word loads of a block covering individuals [i_lo, i_hi) and the loci
// [l0, l0 + nl) stay inside the raw buffer (3 bytes of alignment slack below, 8 above).
__device__ __forceinline__ bool block_is_safe(const DecodeParams& p, int l0, int nl, int i_lo, int i_hi) {
  if (p.row_idx) return true;  // gathered rows are validated indices: byte loads, always in range
  int64_t cmin = INT64_MAX, cmax = -1;
  for (int k = 0; k < nl; ++k) {
    const int l = l0 + k;
    if (l >= p.lc) break;
    const int64_t col = p.col_idx ? int64_t(p.col_idx[l]) : p.col0 + l;
    cmin = col < cmin ? col : cmin;
    cmax = col > cmax ? col : cmax;
  }
  if (cmax < 0) return true;
  const int64_t lo = cmin * p.raw_stride + p.row0 + i_lo - 3;
  const int64_t hi = cmax * p.raw_stride + p.row0 + i_hi + 8;
  return lo >= 0 && hi <= int64_t(p.raw_end - p.raw);
}

__device__ __forceinline__ RowSel make_rowsel(const DecodeParams& p, int i0) {
  RowSel rs;
  rs.gather = p.row_idx != nullptr;
  const int nv = min(max(p.n - i0, 0), 4);
  rs.keep = nv >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nv)) - 1u);
  rs.off = p.row0 + i0;
  rs.ld = i0 < p.n + 4;
#pragma unroll
  for (int b = 0; b < 4; ++b) rs.r[b] = (rs.gather && b < nv) ? p.row_idx[i0 + b] : 0;
  return rs;
}

// ------------------------------------------------------------------ MN-major (streaming)
constexpr int kDecThreads = 256;
constexpr int kDecLociPerBlock = 16;  // two batches of 8 independent loads per lane

template <bool kSafe>
__device__ __forceinline__ void decode_mn_body(const DecodeParams& p, const RowSel& rs, int i0, int l0,
                                               int lane) {
#pragma unroll
  for (int half = 0; half < kDecLociPerBlock / 8; ++half) {
    uint32_t g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int l = l0 + half * 8 + u;
      if (l < p.lc) {
        const int64_t col = p.col_idx ? int64_t(p.col_idx[l]) : p.col0 + l;
        g[u] = load4<kSafe>(p, rs, col, lane);
      } else {
        g[u] = 0x03030303u;  // padding loci decode to zeros
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int l = l0 + half * 8 + u;
      if (l >= p.lc_pad) break;
      uint32_t d, nm, h1;
      decode4(g[u], d, nm, h1);
      int8_t* o = p.planes + int64_t(l) * p.n_pad + i0;
      *reinterpret_cast<uint32_t*>(o) = d;
      *reinterpret_cast<uint32_t*>(o + p.plane_stride) = nm;
      if (p.nplanes > 2) *reinterpret_cast<uint32_t*>(o + 2 * p.plane_stride) = h1;
    }
  }
}

// Straight-line fast path for interior blocks (all 16 loci real, every load in range): the
// row-gather / column-index choices are template parameters so that no branch separates the
// loads and 8 (contiguous rows) or 32 (gathered rows) of them are in flight per lane.
template <bool kGather, bool kColIdx>
__device__ __forceinline__ void decode_mn_fast(const DecodeParams& p, const RowSel& rs, int i0, int l0,
                                               int lane) {
#pragma unroll
  for (int half = 0; half < kDecLociPerBlock / 8; ++half) {
    const uint8_t* cb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int l = l0 + half * 8 + u;
      const int64_t col = kColIdx ? int64_t(__ldg(p.col_idx + l)) : p.col0 + l;
      cb[u] = p.raw + col * p.raw_stride;
    }
    uint32_t g[8];
    if (kGather) {
      uint32_t b0[8], b1[8], b2[8], b3[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // rows of padding individuals are 0 (a valid address)
        b0[u] = __ldg(cb[u] + rs.r[0]);
        b1[u] = __ldg(cb[u] + rs.r[1]);
        b2[u] = __ldg(cb[u] + rs.r[2]);
        b3[u] = __ldg(cb[u] + rs.r[3]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) g[u] = b0[u] | (b1[u] << 8) | (b2[u] << 16) | (b3[u] << 24);
    } else {
      uint32_t w0[8], wx[8], sh[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint8_t* a = cb[u] + rs.off;
        sh[u] = uint32_t(reinterpret_cast<uintptr_t>(a) & 3u);
        const uint32_t* al = reinterpret_cast<const uint32_t*>(a - sh[u]);
        w0[u] = __ldg(al);
        wx[u] = 0;
        if (lane == 31) wx[u] = __ldg(al + 1);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint32_t w1 = __shfl_down_sync(0xFFFFFFFFu, w0[u], 1);
        if (lane == 31) w1 = wx[u];
        g[u] = __funnelshift_r(w0[u], w1, sh[u] * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int l = l0 + half * 8 + u;
      uint32_t d, nm, h1;
      decode4((g[u] & rs.keep) | (0x03030303u & ~rs.keep), d, nm, h1);
      int8_t* o = p.planes + int64_t(l) * p.n_pad + i0;
      *reinterpret_cast<uint32_t*>(o) = d;
      *reinterpret_cast<uint32_t*>(o + p.plane_stride) = nm;
      if (p.nplanes > 2) *reinterpret_cast<uint32_t*>(o + 2 * p.plane_stride) = h1;
    }
  }
}

__global__ void __launch_bounds__(kDecThreads)
decode_planes_mn_kernel(const DecodeParams p) {
  const int lane = threadIdx.x & 31;
  // 1-D grid, individual blocks fastest: CTAs resident together sweep whole raw columns.
  const int nib = (p.n_pad + kDecThreads * 4 - 1) / (kDecThreads * 4);
  const int ib = blockIdx.x % nib, lb = blockIdx.x / nib;
  const int i0 = (ib * kDecThreads + threadIdx.x) * 4;
  if (i0 >= p.n_pad) return;  // n_pad is a multiple of 128: whole warps leave together
  const RowSel rs = make_rowsel(p, i0);
  const int l0 = lb * kDecLociPerBlock;
  const bool safe = block_is_safe(p, l0, kDecLociPerBlock, ib * kDecThreads * 4, (ib + 1) * kDecThreads * 4);
  if (safe && l0 + kDecLociPerBlock <= p.lc) {
    if (p.row_idx) {
      if (p.col_idx) decode_mn_fast<true, true>(p, rs, i0, l0, lane);
      else decode_mn_fast<true, false>(p, rs, i0, l0, lane);
    } else {
      if (p.col_idx) decode_mn_fast<false, true>(p, rs, i0, l0, lane);
      else decode_mn_fast<false, false>(p, rs, i0, l0, lane);
    }
  } else if (safe) {
    decode_mn_body<true>(p, rs, i0, l0, lane);
  } else {
    decode_mn_body<false>(p, rs, i0, l0, lane);
  }
}

// ------------------------------------------------------------------ row / locus gather
// Gathered rows (ind.row in arbitrary order) read one byte per genotype from anywhere in a raw column:
// every byte is its own 32-byte sector request, and what makes that affordable is L1 reuse — many
// individuals of the SAME column read by the same SM.  The 4-bit decode kernel works on 128
// individuals x 256 loci and keeps 192 KiB of the SM as shared memory, the opposite of what a gather
// wants (measured: 10,000-row columns, 7,000 shuffled rows: 0.7 TB/s algorithmic).  So device-resident
// gathered jobs first COMPACT the chunk — this kernel: 1024 individuals x 16 loci per CTA, no shared
// memory, 32 byte loads in flight per lane, one coalesced 4-byte store per lane and locus — into a dense
// column-major chunk (stride n_pad, padding individuals = 0x03), and the decode kernel then takes its
// contiguous fast path.  Costs one extra byte written and read per genotype.
template <bool kColIdx>
__global__ void __launch_bounds__(kDecThreads)
compact_rows_kernel(const DecodeParams p, uint8_t* __restrict__ out) {
  const int nib = (p.n_pad + kDecThreads * 4 - 1) / (kDecThreads * 4);
  const int ib = blockIdx.x % nib, lb = blockIdx.x / nib;   // individual blocks fastest: CTAs resident together share columns
  const int i0 = (ib * kDecThreads + threadIdx.x) * 4;
  if (i0 >= p.n_pad) return;
  const RowSel rs = make_rowsel(p, i0);                      // rows of padding individuals are 0 (a valid address)
  const int l0 = lb * kDecLociPerBlock;
#pragma unroll
  for (int half = 0; half < kDecLociPerBlock / 8; ++half) {
    const uint8_t* cb[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int l = min(l0 + half * 8 + u, p.lc - 1);        // loci past the chunk re-read the last one (not stored)
      const int64_t col = kColIdx ? int64_t(__ldg(p.col_idx + l)) : p.col0 + l;
      cb[u] = p.raw + col * p.raw_stride;
    }
    uint32_t b0[8], b1[8], b2[8], b3[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      b0[u] = __ldg(cb[u] + rs.r[0]);
      b1[u] = __ldg(cb[u] + rs.r[1]);
      b2[u] = __ldg(cb[u] + rs.r[2]);
      b3[u] = __ldg(cb[u] + rs.r[3]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int l = l0 + half * 8 + u;
      if (l >= p.lc) break;
      const uint32_t g = b0[u] | (b1[u] << 8) | (b2[u] << 16) | (b3[u] << 24);
      *reinterpret_cast<uint32_t*>(out + int64_t(l) * p.n_pad + i0) = (g & rs.keep) | (0x03030303u & ~rs.keep);
    }
  }
}

int launch_compact_rows(Ctx* ctx, const DecodeArgs& a, uint8_t* out, cudaStream_t st, const Err& E) {
  if (!a.row_idx || a.n_pad % 128 || a.n > a.n_pad || a.n <= 0 || a.lc <= 0 || a.n_pad > INT32_MAX || a.lc > INT32_MAX ||
      (reinterpret_cast<uintptr_t>(out) & 3))
    return E.fail(TPG_ERR_ARG, "compact_rows: needs a row index, n_pad a multiple of 128, a 4-byte aligned target");
  DecodeParams p;
  p.raw = a.raw;
  p.raw_end = a.raw + a.raw_bytes;
  p.raw_stride = a.raw_stride;
  p.row_idx = a.row_idx;
  p.row0 = a.row0;
  p.col_idx = a.col_idx;
  p.col0 = a.col0;
  p.n = int(a.n);
  p.n_pad = int(a.n_pad);
  p.lc = int(a.lc);
  p.lc_pad = int(a.lc);
  p.planes = nullptr;
  p.plane_stride = 0;
  p.nplanes = 0;
  const int64_t blocks = ((a.n_pad + kDecThreads * 4 - 1) / (kDecThreads * 4)) * ((a.lc + kDecLociPerBlock - 1) / kDecLociPerBlock);
  if (blocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "compact_rows: chunk too large for one launch");
  if (a.col_idx) compact_rows_kernel<true><<<unsigned(blocks), kDecThreads, 0, st>>>(p, out);
  else compact_rows_kernel<false><<<unsigned(blocks), kDecThreads, 0, st>>>(p, out);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

// ------------------------------------------------------------------ row gather through shared memory
// SubBMAcc's row gather (src/snp_ibs.cpp:35: macc(i, j) = bk[(rowInd[i] - 1) + (colInd[j] - 1) * nrow]) for
// ind.row in arbitrary order.  A byte gather straight from global memory costs one 32-byte sector per
// genotype (compact_rows_kernel above: 0.13 of the HBM figure).  Here a CTA first stages the needed row
// span [row_lo, row_lo + span) of L loci in shared memory with coalesced 16-byte loads — every byte of the
// span crosses HBM once, in whole lines — and then gathers from shared memory by row index: a thread owns
// 16 consecutive OUTPUT individuals (their 16 row offsets stay in registers for all L loci) and emits one
// 32-bit word of 2-bit codes per locus, i.e. the dense chunk leaves in the pk2 layout (pack.cpp) that
// decode_planes_k4_kernel<true> / decode_planes_pk2_kernel read: a quarter of the bytes of a raw compaction.
// kPackedIn: the staged chunk is itself 2-bit packed (a host chunk packed by the worker pool while staging),
// so gathered host jobs cross PCIe at 0.25 B per genotype too.
// Traffic: span (or span / 4) bytes read per locus + n / 4 bytes written.
constexpr int kGatherThreads = 256;

struct GatherParams {
  const uint8_t* raw;        // chunk base
  const uint8_t* raw_end;
  int64_t raw_stride;        // bytes between consecutive columns of the chunk
  const int32_t* row_idx;    // n offsets, relative to the first row of the staged span
  const int32_t* col_idx;    // device sources with gathered loci (raw columns), or nullptr
  int64_t col0;              // first column (contiguous loci)
  int64_t row_lo;            // first row of the span within a column (raw input: bytes; packed input: individuals, multiple of 4)
  int span_bytes;            // bytes of a column to stage (raw: rows; packed: ceil(rows / 4))
  int n;                     // output individuals
  int lc;                    // loci in the chunk
  int loci_per_cta;
  int col_pitch;             // smem bytes per staged column (multiple of 16, >= span_bytes + 16)
  uint8_t* out;              // pk2 dense chunk: column l at out + l * out_stride
  int64_t out_stride;        // multiple of 16
};

__device__ __forceinline__ uint32_t smem_u32_(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <bool kPackedIn>
__global__ void __launch_bounds__(kGatherThreads)
gather_pack_kernel(const GatherParams p) {
  extern __shared__ uint4 gsm4[];
  uint8_t* sm = reinterpret_cast<uint8_t*>(gsm4);
  __shared__ int mis_s[64];
  __shared__ const uint8_t* col_s[64];
  const int l0 = blockIdx.x * p.loci_per_cta;
  const int nl = min(p.loci_per_cta, p.lc - l0);
  // ---- stage: column l lands at sm + l * col_pitch + (its global address & 15), in 16-byte chunks copied
  // asynchronously (cp.async: no register staging, every chunk of the CTA in flight at once)
  const int chunks = (p.span_bytes + 15 + 15) / 16;  // covers any misalignment
  if (threadIdx.x < nl) {
    const int l = threadIdx.x;
    const int64_t col = p.col_idx ? int64_t(__ldg(p.col_idx + l0 + l)) : p.col0 + l0 + l;
    const uint8_t* g = p.raw + col * p.raw_stride + (kPackedIn ? (p.row_lo >> 2) : p.row_lo);
    col_s[l] = g;
    mis_s[l] = int(reinterpret_cast<uintptr_t>(g) & 15u);
  }
  __syncthreads();
  const int total = nl * chunks;
  for (int f = threadIdx.x; f < total; f += kGatherThreads) {
    const int l = f / chunks, c = f - l * chunks;
    const uint8_t* g = col_s[l];
    const uint8_t* src = g - mis_s[l] + size_t(c) * 16;
    if (src >= g + p.span_bytes) continue;
    uint8_t* dst = sm + size_t(l) * p.col_pitch + size_t(c) * 16;
    if (src >= p.raw && src + 16 <= p.raw_end) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32_(dst)), "l"(src) : "memory");
    } else {  // first / last bytes of the allocation: byte-wise, never outside [raw, raw_end)
      uint32_t w[4] = {0, 0, 0, 0};
      for (int b = 0; b < 16; ++b)
        if (src + b >= p.raw && src + b < p.raw_end) w[b >> 2] |= uint32_t(__ldg(src + b)) << (8 * (b & 3));
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- gather: one 32-bit word (16 individuals) per thread and locus
  const int words = int(p.out_stride >> 2);
  for (int w = threadIdx.x; w < words; w += kGatherThreads) {
    const int i0 = w * 16;
    int32_t r[16];
    if (i0 + 16 <= p.n) {
      const int4* rp = reinterpret_cast<const int4*>(p.row_idx + i0);  // row_idx is 16-byte aligned (cudaMalloc)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 t = __ldg(rp + q);
        r[4 * q] = t.x; r[4 * q + 1] = t.y; r[4 * q + 2] = t.z; r[4 * q + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = (i0 + j < p.n) ? __ldg(p.row_idx + i0 + j) : -1;
    }
    const int rbase = kPackedIn ? int(p.row_lo & 3) : 0;  // (row_lo is a multiple of 4 for packed chunks: always 0)
    for (int l = 0; l < nl; ++l) {
      const uint8_t* scol = sm + size_t(l) * p.col_pitch + mis_s[l];
      uint32_t word = 0;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t c = 3u;
        if (r[j] >= 0) {
          if (kPackedIn) {
            const int rr = r[j] + rbase;
            c = (uint32_t(scol[rr >> 2]) >> ((rr & 3) * 2)) & 3u;
          } else {
            c = min(uint32_t(scol[r[j]]), 3u);
          }
        }
        word |= c << (2 * j);
      }
      *reinterpret_cast<uint32_t*>(p.out + int64_t(l0 + l) * p.out_stride + int64_t(w) * 4) = word;
    }
  }
}

// A second build (tpg_config.gather_layout = 1; measured SLOWER than the column-major one above, 2.7 vs 1.6 ms on
// the C5 subset, and therefore not the default — kept for the cross-check and the next look at it): the staged tile is TRANSPOSED — word r of a group
// holds row r (packed input: byte r) of FOUR loci — so one shared-memory read serves four loci and the gather
// phase's bank-conflicted lookups drop four-fold (ncu on the column-major build above: l1tex 90 % busy, 210 M bank
// conflicts, 15 instructions per genotype; C5 subset 1.61 ms).  Staging: a thread loads rows 4q..4q+3 of the four
// columns (two aligned words + a funnel shift each: the columns are unaligned), transposes the 4 x 4 bytes in
// registers (8 PRMT) and stores four words with one 16-byte store.  Gather: four individuals at a time — four
// lookups, clamp (raw) or shift + mask (packed), another 4 x 4 transpose, and each locus gets a byte of four codes.
template <bool kPackedIn>
__global__ void __launch_bounds__(kGatherThreads)
gather_pack_t4_kernel(const GatherParams p) {
  extern __shared__ uint4 gsm4[];
  uint32_t* tile = reinterpret_cast<uint32_t*>(gsm4);   // [groups][4 * nq] words
  __shared__ const uint8_t* col_s[64];
  __shared__ int safe_s;
  const int l0 = blockIdx.x * p.loci_per_cta;
  const int nl = min(p.loci_per_cta, p.lc - l0);
  const int ngroups = (nl + 3) >> 2;
  const int nq = (p.span_bytes + 3) >> 2;               // quads of bytes per column
  const int pitch = nq * 4;                             // words per group
  if (threadIdx.x == 0) safe_s = 1;
  __syncthreads();
  if (threadIdx.x < 4 * ngroups) {
    const int l = min(int(threadIdx.x), nl - 1);        // a short last group re-reads the last locus (not stored)
    const int64_t col = p.col_idx ? int64_t(__ldg(p.col_idx + l0 + l)) : p.col0 + l0 + l;
    const uint8_t* g = p.raw + col * p.raw_stride + (kPackedIn ? (p.row_lo >> 2) : p.row_lo);
    col_s[threadIdx.x] = g;
    // every aligned word a staging thread may touch lies inside the buffer?
    if (g - 3 < p.raw || g + int64_t(nq) * 4 + 8 > p.raw_end) safe_s = 0;
  }
  __syncthreads();
  const bool safe = safe_s != 0;
  for (int g = 0; g < ngroups; ++g) {
    const uint8_t* c[4] = {col_s[4 * g], col_s[4 * g + 1], col_s[4 * g + 2], col_s[4 * g + 3]};
    for (int q = threadIdx.x; q < nq; q += kGatherThreads) {
      uint32_t w[4], t[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const uint8_t* a = c[l] + 4 * q;
        const uint32_t sh = uint32_t(reinterpret_cast<uintptr_t>(a) & 3u);
        const uint8_t* al = a - sh;
        uint32_t w0, w1;
        if (safe) {
          w0 = __ldg(reinterpret_cast<const uint32_t*>(al));
          w1 = __ldg(reinterpret_cast<const uint32_t*>(al) + 1);
        } else {
          w0 = guarded_ld32(al, p.raw, p.raw_end);
          w1 = guarded_ld32(al + 4, p.raw, p.raw_end);
        }
        w[l] = __funnelshift_r(w0, w1, sh * 8);
      }
      transpose4x4(w, t);                               // t[r]: byte l = locus 4g + l at row 4q + r
      *reinterpret_cast<uint4*>(tile + g * pitch + 4 * q) = make_uint4(t[0], t[1], t[2], t[3]);
    }
  }
  __syncthreads();
  const int words = int(p.out_stride >> 2);
  for (int wi = threadIdx.x; wi < words; wi += kGatherThreads) {
    const int i0 = wi * 16;
    int32_t r[16];
    if (i0 + 16 <= p.n) {
      const int4* rp = reinterpret_cast<const int4*>(p.row_idx + i0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int4 v = __ldg(rp + q);
        r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = (i0 + j < p.n) ? __ldg(p.row_idx + i0 + j) : -1;
    }
    for (int g = 0; g < ngroups; ++g) {
      const uint32_t* tg = tile + g * pitch;
      uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        uint32_t cw[4], t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = r[4 * b + k];
          if (kPackedIn) {
            const uint32_t x = rr >= 0 ? (tg[rr >> 2] >> (2 * (rr & 3))) : 0xFFFFFFFFu;
            cw[k] = x & 0x03030303u;
          } else {
            cw[k] = clamp4(rr >= 0 ? tg[rr] : 0xFFFFFFFFu);
          }
        }
        transpose4x4(cw, t);                            // t[l]: byte k = code of individual 4b + k at locus 4g + l
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          const uint32_t pk = t[l] | (t[l] >> 6) | (t[l] >> 12) | (t[l] >> 18);
          acc[l] |= (pk & 0xFFu) << (8 * b);
        }
      }
#pragma unroll
      for (int l = 0; l < 4; ++l)
        if (4 * g + l < nl)
          *reinterpret_cast<uint32_t*>(p.out + int64_t(l0 + 4 * g + l) * p.out_stride + int64_t(wi) * 4) = acc[l];
    }
  }
}

// Shared memory a gather CTA may use.  A CTA alternates between two phases that want different resources —
// staging (HBM latency / bandwidth) and the gather (shared-memory lookups) — so several small CTAs per SM
// overlap better than two large ones: four per SM while a few columns fit 52 KiB, fewer for longer columns.
constexpr int kGatherSmemTwo = 52 * 1024, kGatherSmemMax = 200 * 1024;

bool gather_pack_fits(int64_t span_bytes) { return span_bytes + 32 <= kGatherSmemMax; }

int launch_gather_pack(Ctx* ctx, const DecodeArgs& a, bool packed_in, int64_t row_lo, int64_t span_rows,
                       uint8_t* out, int64_t out_stride, cudaStream_t st, const Err& E) {
  const int64_t span_bytes = packed_in ? (span_rows + 3) / 4 : span_rows;
  if (!a.row_idx || a.n <= 0 || a.lc <= 0 || a.n > INT32_MAX || a.lc > INT32_MAX || out_stride % 16 ||
      out_stride * 4 < a.n || (reinterpret_cast<uintptr_t>(out) & 15) || (packed_in && (row_lo & 3)))
    return E.fail(TPG_ERR_ARG, "gather_pack: needs a row index and a 16-byte aligned pk2 target");
  if (!gather_pack_fits(span_bytes))
    return E.fail(TPG_ERR_ARG, "gather_pack: a %lld-byte column span does not fit shared memory", (long long)span_bytes);
  GatherParams p;
  p.raw = a.raw;
  p.raw_end = a.raw + a.raw_bytes;
  p.raw_stride = a.raw_stride;
  p.row_idx = a.row_idx;
  p.col_idx = a.col_idx;
  p.col0 = a.col0;
  p.row_lo = row_lo;
  p.span_bytes = int(span_bytes);
  p.n = int(a.n);
  p.lc = int(a.lc);
  p.col_pitch = int((span_bytes + 16 + 15) / 16 * 16);
  int budget = kGatherSmemTwo;
  const tpg_config C = cfg();
  if (C.gather_smem_kb > 0) budget = std::min(kGatherSmemMax, std::max(8, int(C.gather_smem_kb)) * 1024);
  // the transposed build stages groups of four columns (16 bytes per quad of rows and group)
  const int64_t group_bytes = (span_bytes + 3) / 4 * 16;
  const bool t4 = group_bytes <= kGatherSmemMax && C.gather_layout == 1;
  const int64_t unit = t4 ? group_bytes : p.col_pitch;            // smallest piece of shared memory a CTA can work with
  while (budget < kGatherSmemMax && unit * (t4 ? 1 : 2) > budget) budget *= 2;   // long columns: fewer CTAs per SM
  budget = std::min(budget, kGatherSmemMax);
  if (t4) {
    p.loci_per_cta = int(std::max<int64_t>(1, std::min<int64_t>(16, budget / group_bytes))) * 4;
    while (p.loci_per_cta > 4 && (a.lc + p.loci_per_cta - 1) / p.loci_per_cta < 2 * ctx->num_sms) p.loci_per_cta -= 4;
  } else {
    p.loci_per_cta = int(std::max<int64_t>(1, std::min<int64_t>(64, budget / p.col_pitch)));
    // enough CTAs to fill the machine: shrink the loci per CTA for short chunks
    while (p.loci_per_cta > 1 && (a.lc + p.loci_per_cta - 1) / p.loci_per_cta < 2 * ctx->num_sms) p.loci_per_cta /= 2;
  }
  p.out = out;
  p.out_stride = out_stride;
  const size_t smem = t4 ? size_t(p.loci_per_cta / 4) * size_t(group_bytes) : size_t(p.loci_per_cta) * size_t(p.col_pitch);
  const unsigned blocks = unsigned((a.lc + p.loci_per_cta - 1) / p.loci_per_cta);
  if (!ctx->gather_attr_set) {  // function attributes are per device
    TPG_CUDA_OK(cudaFuncSetAttribute(gather_pack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmemMax), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gather_pack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmemMax), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gather_pack_t4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmemMax), E);
    TPG_CUDA_OK(cudaFuncSetAttribute(gather_pack_t4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGatherSmemMax), E);
    ctx->gather_attr_set = true;
  }
  if (t4 && packed_in) gather_pack_t4_kernel<true><<<blocks, kGatherThreads, smem, st>>>(p);
  else if (t4) gather_pack_t4_kernel<false><<<blocks, kGatherThreads, smem, st>>>(p);
  else
  if (packed_in) gather_pack_kernel<true><<<blocks, kGatherThreads, smem, st>>>(p);
  else gather_pack_kernel<false><<<blocks, kGatherThreads, smem, st>>>(p);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

// ------------------------------------------------------------------ K-major (transposing)
// Block = 128 individuals x 128 loci; warp w owns loci [16w, 16w+16), lane owns 4 individuals.
__global__ void __launch_bounds__(256)
decode_planes_k_kernel(const DecodeParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nib = p.n_pad / 128;
  const int i0 = int(blockIdx.x % nib) * 128 + lane * 4;
  const int l0 = int(blockIdx.x / nib) * 128 + warp * 16;
  if (l0 >= p.lc_pad) return;
  const RowSel rs = make_rowsel(p, i0);
  const int ib0 = int(blockIdx.x % nib) * 128;
  const bool safe = block_is_safe(p, l0, 16, ib0, ib0 + 128);
  uint32_t g[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int l = l0 + k;
    if (l < p.lc) {
      const int64_t col = p.col_idx ? int64_t(p.col_idx[l]) : p.col0 + l;
      g[k] = safe ? load4<true>(p, rs, col, lane) : load4<false>(p, rs, col, lane);
    } else {
      g[k] = 0x03030303u;
    }
  }
  // 4x4 byte transposes: t[r][q] = loci 4q..4q+3 of individual i0+r.
  uint32_t t[4][4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint32_t a0 = __byte_perm(g[4 * q], g[4 * q + 1], 0x5140);
    const uint32_t a1 = __byte_perm(g[4 * q + 2], g[4 * q + 3], 0x5140);
    const uint32_t b0 = __byte_perm(g[4 * q], g[4 * q + 1], 0x7362);
    const uint32_t b1 = __byte_perm(g[4 * q + 2], g[4 * q + 3], 0x7362);
    t[0][q] = __byte_perm(a0, a1, 0x5410);
    t[1][q] = __byte_perm(a0, a1, 0x7632);
    t[2][q] = __byte_perm(b0, b1, 0x5410);
    t[3][q] = __byte_perm(b0, b1, 0x7632);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    uint4 vd, vn, vh;
    decode4(t[r][0], vd.x, vn.x, vh.x);
    decode4(t[r][1], vd.y, vn.y, vh.y);
    decode4(t[r][2], vd.z, vn.z, vh.z);
    decode4(t[r][3], vd.w, vn.w, vh.w);
    int8_t* o = p.planes + int64_t(i0 + r) * p.lc_pad + l0;
    *reinterpret_cast<uint4*>(o) = vd;
    *reinterpret_cast<uint4*>(o + p.plane_stride) = vn;
    if (p.nplanes > 2) *reinterpret_cast<uint4*>(o + 2 * p.plane_stride) = vh;
  }
}

// ------------------------------------------------------------------ K-major 4-bit planes (E2M1)
// Operand planes of the kind::mxf4 Gram kernel (gram.cu): every plane value is -1, 0 or +1, which
// E2M1 holds exactly (+1 = 0x2, -1 = 0xA, 0 = 0x0), two per byte, K-major:
//   plane[i * pitch + byte], pitch = lc_pad / 2, lc_pad a multiple of 256 (128-byte rows for TMA).
// Within a group of 8 loci the byte k (0..3) holds locus k in its low and locus k + 4 in its high
// nibble — a fixed permutation of K that is the same for every row and plane, which a dot product
// over K does not see, and that makes packing two transposed words one shift and one OR.
// Block = 128 individuals x 256 loci; warp w owns loci [32w, 32w + 32), a lane 4 individuals:
// 32 loads in flight per lane, 4x4 byte transposes (PRMT), one 16-byte store per (individual, plane).
// Traffic: 1 byte read + 0.5 P bytes written per genotype.
// E2M1 nibble codes of 4 packed genotypes, straight from the raw bytes (one code per byte, in its
// low nibble): cn = 0x2 where byte <= 2; ch = 0x2 where byte == 1; cd = 0x2 where byte == 2,
// 0xA (-1) where byte == 0.
__device__ __forceinline__ uint32_t spread4(uint32_t t) {  // 4 codes in bits [0,8) -> 4 bytes
  return (t | (t << 6) | (t << 12) | (t << 18)) & 0x03030303u;
}

__device__ __forceinline__ void codes4(uint32_t g, uint32_t& cd, uint32_t& cn, uint32_t& ch) {
  const uint32_t t = (g & 0x7F7F7F7Fu) + 0x7D7D7D7Du;   // bit 7 set iff (byte & 127) >= 3
  cn = (~(t | g) >> 6) & 0x02020202u;                   // bit 7 clear <=> byte <= 2 -> bit 1
  ch = (g << 1) & cn;                                   // valid bytes are 0,1,2: bit 0 <=> het
  const uint32_t c2 = g & cn;                           // bit 1 <=> hom-alt
  const uint32_t c0 = cn ^ ch ^ c2;                     // the remaining valid ones: hom-ref
  cd = c0 * 5u + c2;                                    // 0x02 * 5 = 0x0A (-1), no carries between bytes
}

// First build of the kernel (byte-density codes4 after the transposes, shuffle-assembled loads):
// kept selectable (TPG_DECODE_K4_V1=1) as the cross-check of decode_planes_k4_kernel below.
__global__ void __launch_bounds__(256, 4)
decode_planes_k4_v1_kernel(const DecodeParams p) {
  __shared__ uint4 tile[3][128][8];  // output rows, 16-byte chunks XOR-swizzled by row
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nib = p.n_pad / 128;
  const int ib0 = int(blockIdx.x % nib) * 128;
  const int i0 = ib0 + lane * 4;
  const int lt = int(blockIdx.x / nib) * 256;               // lc_pad is a multiple of 256: always in range
  const RowSel rs = make_rowsel(p, i0);
  const bool full = rs.keep == 0xFFFFFFFFu;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {                     // 16 loci at a time keeps the registers down
    const int l0 = lt + warp * 32 + half * 16;
    uint32_t g[16];
    // Contiguous rows and columns, all 16 loci real and every word load inside the buffer: the
    // straight-line path.  Column u starts stride bytes after column u - 1, so addresses and the
    // byte misalignment of a column are running sums, not 64-bit multiplies.
    bool fast = !p.row_idx && !p.col_idx && l0 + 16 <= p.lc;
    if (fast) {
      const int64_t lo = (p.col0 + l0) * p.raw_stride + p.row0 + ib0 - 3;
      const int64_t hi = (p.col0 + l0 + 15) * p.raw_stride + p.row0 + ib0 + 128 + 8;
      fast = lo >= 0 && hi <= int64_t(p.raw_end - p.raw);
    }
    const bool safe = fast || block_is_safe(p, l0, 16, ib0, ib0 + 128);
    if (fast) {
      const uint8_t* a = p.raw + (p.col0 + l0) * p.raw_stride + rs.off;
      uint32_t wx[16];
      uint32_t shs = 0;  // 2 bits per locus
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const uint32_t sh = uint32_t(reinterpret_cast<uintptr_t>(a) & 3u);
        shs |= sh << (2 * u);
        const uint32_t* al = reinterpret_cast<const uint32_t*>(a - sh);
        g[u] = __ldg(al);
        wx[u] = 0;
        if (lane == 31) wx[u] = __ldg(al + 1);
        a += p.raw_stride;
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint32_t w1 = __shfl_down_sync(0xFFFFFFFFu, g[u], 1);
        if (lane == 31) w1 = wx[u];
        g[u] = __funnelshift_r(g[u], w1, ((shs >> (2 * u)) & 3u) * 8);
      }
      if (!full) {
#pragma unroll
        for (int u = 0; u < 16; ++u) g[u] = (g[u] & rs.keep) | (0x03030303u & ~rs.keep);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int l = l0 + k;
        g[k] = 0x03030303u;
        if (l < p.lc) {
          const int64_t col = p.col_idx ? int64_t(p.col_idx[l]) : p.col0 + l;
          g[k] = safe ? load4<true>(p, rs, col, lane) : load4<false>(p, rs, col, lane);
        }
      }
    }
    // out[r][w]: 8 loci per word, 2 words = 16 loci per individual r
    uint32_t od[4][2], on[4][2], oh[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // 4 loci at a time: t[r] = loci 4q..4q+3 of individual i0 + r
      const uint32_t a0 = __byte_perm(g[4 * q], g[4 * q + 1], 0x5140);
      const uint32_t a1 = __byte_perm(g[4 * q + 2], g[4 * q + 3], 0x5140);
      const uint32_t b0 = __byte_perm(g[4 * q], g[4 * q + 1], 0x7362);
      const uint32_t b1 = __byte_perm(g[4 * q + 2], g[4 * q + 3], 0x7362);
      uint32_t t[4];
      t[0] = __byte_perm(a0, a1, 0x5410);
      t[1] = __byte_perm(a0, a1, 0x7632);
      t[2] = __byte_perm(b0, b1, 0x5410);
      t[3] = __byte_perm(b0, b1, 0x7632);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        uint32_t cd, cn, ch;
        codes4(t[r], cd, cn, ch);
        const int w = q >> 1;
        if ((q & 1) == 0) { od[r][w] = cd; on[r][w] = cn; oh[r][w] = ch; }          // loci k     -> low nibbles
        else { od[r][w] |= cd << 4; on[r][w] |= cn << 4; oh[r][w] |= ch << 4; }     // loci k + 4 -> high nibbles
      }
    }
    // 8 bytes per (row, plane) into the row's chunk `warp`, half `half`; chunks swizzled by row
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = lane * 4 + r, ch = warp ^ ((row >> 2) & 7);
      reinterpret_cast<uint2*>(&tile[0][row][ch])[half] = make_uint2(od[r][0], od[r][1]);
      reinterpret_cast<uint2*>(&tile[1][row][ch])[half] = make_uint2(on[r][0], on[r][1]);
      if (p.nplanes > 2) reinterpret_cast<uint2*>(&tile[2][row][ch])[half] = make_uint2(oh[r][0], oh[r][1]);
    }
  }
  __syncthreads();
  // whole 128-byte lines out: a warp store covers 4 rows x 8 chunks
  const int64_t pitch = p.lc_pad >> 1;
  const int lbyte = lt >> 1;  // this tile's first byte within a plane row
  for (int pln = 0; pln < p.nplanes; ++pln)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = warp * 16 + k * 4 + (lane >> 3), c = lane & 7;
      const uint4 v = tile[pln][row][c ^ ((row >> 2) & 7)];
      *reinterpret_cast<uint4*>(p.planes + int64_t(pln) * p.plane_stride + int64_t(ib0 + row) * pitch + lbyte + c * 16) = v;
    }
}

// The default build.  Same block shape, layout and staging as above; what changed is the
// instruction count of the straight-line path (the first build was issue-bound: 648 instructions
// per 64 genotypes and lane, ncu issue slots 63 % busy at 0.58 of the HBM figure):
//   * loads: one IMAD.WIDE forms the address of column u (a0 + u * stride), both aligned words
//     are loaded by the lane itself (the second one is an L1 hit) — no shuffle, no select, no
//     packed shift table;
//   * compute at nibble density (swar.cuh): bytes are clamped to 2-bit codes first (5 ops per
//     4 genotypes), loci k and k + 4 are paired into nibbles BEFORE the 4x4 byte transposes
//     (half the PRMTs), and the three E2M1 codes are derived on 8 genotypes per word (6 ops).
// 0.376 -> 0.329 ms on the 929 x 650,000 chunk (ncu: 154 M warp instructions instead of 278 M, issue
// slots 43 % busy, DRAM 4.75 TB/s).  Measured beside it and rejected: writing 16 bytes per
// (individual, plane) straight from registers, no shared memory and no barrier — 0.81 ms when the
// two halves of a 32-byte sector come from different warps (partial-sector writes), 0.40 ms when
// one store instruction completes the sector; issuing all 32 loci of loads before decoding
// (80 registers, 3 CTAs per SM) changed nothing.
// kPacked: the chunk crossed PCIe 2-bit packed (pk2, below) — a lane's four individuals are ONE byte
// per locus (a warp reads one 32-byte sector per locus) and already hold the clamped codes.
template <bool kPacked>
__global__ void __launch_bounds__(256, 4)
decode_planes_k4_kernel(const DecodeParams p) {
  __shared__ uint4 tile[3][128][8];  // output rows, 16-byte chunks XOR-swizzled by row
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nib = p.n_pad / 128;
  const int ib0 = int(blockIdx.x % nib) * 128;
  const int i0 = ib0 + lane * 4;
  const int lt = int(blockIdx.x / nib) * 256;               // lc_pad is a multiple of 256: always in range
  const RowSel rs = make_rowsel(p, i0);
  const bool full = rs.keep == 0xFFFFFFFFu;
  const bool plain = !p.row_idx && !p.col_idx && p.raw_stride < (int64_t(1) << 27);
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {                     // 16 loci at a time keeps the registers down
    const int l0 = lt + warp * 32 + half * 16;
    uint32_t q[16];                                          // clamped codes, one per byte: 0 / 1 / 2, 3 = missing
    if (kPacked) {
      // column l at l * raw_stride, individuals i0 .. i0 + 3 in byte i0 >> 2 (pack.cpp); bytes beyond the
      // packed column and loci beyond the chunk read as missing
      const int64_t boff = i0 >> 2;
      const uint8_t* a = p.raw + int64_t(l0) * p.raw_stride + boff;
      uint32_t pk[16];
      if (boff < p.raw_stride && l0 + 16 <= p.lc) {
#pragma unroll
        for (int u = 0; u < 16; ++u) pk[u] = __ldg(a + int64_t(u) * p.raw_stride);
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          pk[u] = 0xFFu;
          if (boff < p.raw_stride && l0 + u < p.lc) pk[u] = __ldg(a + int64_t(u) * p.raw_stride);
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) q[u] = spread4(pk[u]) | (0x03030303u & ~rs.keep);
    } else {
      uint32_t g[16];
      bool fast = plain && l0 + 16 <= p.lc;
      if (fast) {
        const int64_t lo = (p.col0 + l0) * p.raw_stride + p.row0 + ib0 - 3;
        const int64_t hi = (p.col0 + l0 + 15) * p.raw_stride + p.row0 + ib0 + 128 + 8;
        fast = lo >= 0 && hi <= int64_t(p.raw_end - p.raw);
      }
      if (fast) {
        const uint64_t a0 = uint64_t(reinterpret_cast<uintptr_t>(p.raw + (p.col0 + l0) * p.raw_stride + rs.off));
        const uint32_t stride = uint32_t(p.raw_stride);        // < 2^27: 15 * stride fits 32 bits
        uint32_t w1[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          uint64_t au;
          asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(au) : "r"(stride), "r"(uint32_t(u)), "l"(a0));
          const uint32_t* al = reinterpret_cast<const uint32_t*>(uintptr_t(au & ~uint64_t(3)));
          g[u] = __ldg(al);
          w1[u] = __ldg(al + 1);
        }
        const uint32_t a8 = uint32_t(a0) << 3, s8 = stride << 3;  // the funnel shift takes its amount mod 32
#pragma unroll
        for (int u = 0; u < 16; ++u) g[u] = __funnelshift_r(g[u], w1[u], a8 + uint32_t(u) * s8);
      } else {
        const bool safe = block_is_safe(p, l0, 16, ib0, ib0 + 128);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int l = l0 + k;
          g[k] = 0x03030303u;
          if (l < p.lc) {
            const int64_t col = p.col_idx ? int64_t(p.col_idx[l]) : p.col0 + l;
            g[k] = safe ? load4<true>(p, rs, col, lane) : load4<false>(p, rs, col, lane);
          }
        }
      }
      if (fast && !full) {
#pragma unroll
        for (int u = 0; u < 16; ++u) g[u] |= ~rs.keep;         // individuals >= n: any byte above 2 is missing
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) q[u] = clamp4(g[u]);
    }
    // 2-bit codes; loci k (low nibble) and k + 4 (high nibble) of each group of 8 share a byte
    uint32_t od[4][2], on[4][2], oh[4][2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      uint32_t pr[4], t[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) pr[k] = nibble_pair(q[8 * w + k], q[8 * w + 4 + k]);
      transpose4x4(pr, t);                                    // t[r]: byte k = loci (k, k + 4) of individual i0 + r
#pragma unroll
      for (int r = 0; r < 4; ++r) codes8(t[r], od[r][w], on[r][w], oh[r][w]);
    }
    // 8 bytes per (row, plane) into the row's chunk `warp`, half `half`; chunks swizzled by row, and
    // the two halves of a chunk swapped for lanes 8-15 and 24-31, so that the 16 lanes of a half-warp
    // hit 16 different 8-byte bank pairs (without the swap lanes L and L + 8 collide).
    const int hs = half ^ ((lane >> 3) & 1);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = lane * 4 + r, ch = warp ^ ((row >> 2) & 7);
      reinterpret_cast<uint2*>(&tile[0][row][ch])[hs] = make_uint2(od[r][0], od[r][1]);
      reinterpret_cast<uint2*>(&tile[1][row][ch])[hs] = make_uint2(on[r][0], on[r][1]);
      if (p.nplanes > 2) reinterpret_cast<uint2*>(&tile[2][row][ch])[hs] = make_uint2(oh[r][0], oh[r][1]);
    }
  }
  __syncthreads();
  // whole 128-byte lines out: a warp store covers 4 rows x 8 chunks
  const int64_t pitch = p.lc_pad >> 1;
  const int lbyte = lt >> 1;  // this tile's first byte within a plane row
  for (int pln = 0; pln < p.nplanes; ++pln)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = warp * 16 + k * 4 + (lane >> 3), c = lane & 7;
      uint4 v = tile[pln][row][c ^ ((row >> 2) & 7)];
      if ((row >> 5) & 1) v = make_uint4(v.z, v.w, v.x, v.y);  // undo the half swap (row >> 2 = the writing lane)
      // (streaming __stcs stores measured here: no gain, 0.341 vs 0.330 ms at 929 x 650,000)
      *reinterpret_cast<uint4*>(p.planes + int64_t(pln) * p.plane_stride + int64_t(ib0 + row) * pitch + lbyte + c * 16) = v;
    }
}

// ------------------------------------------------------------------ pk2 input (2-bit packed)
// Host-sourced chunks arrive packed four genotypes per byte (pack.cpp: column k at
// k * raw_stride, individual i of the staged window in bits [2(i&3), 2(i&3)+2) of byte i>>2,
// code 3 = missing, columns padded with 0xFF to a multiple of 16 bytes).  A lane owns 16
// consecutive individuals = one aligned 32-bit word; a warp reads 128 contiguous bytes and
// writes 512 contiguous bytes per plane and locus; 8 loci (8 independent loads) per warp.
// Algorithmic traffic: 0.25 byte read + 3 bytes written per genotype.
constexpr int kPkLociPerWarp = 8;


__global__ void __launch_bounds__(kDecThreads)
decode_planes_pk2_kernel(const DecodeParams p) {
  const int lane = threadIdx.x & 31;
  const int nwi = (p.n_pad + 511) / 512;  // warps across the individuals
  const int64_t gw = (int64_t(blockIdx.x) * kDecThreads + threadIdx.x) >> 5;
  const int wi = int(gw % nwi);
  const int l0 = int(gw / nwi) * kPkLociPerWarp;
  const int i0 = wi * 512 + lane * 16;
  if (l0 >= p.lc_pad || i0 >= p.n_pad) return;
  const int64_t boff = int64_t(i0 >> 2);
  const bool in_col = boff + 4 <= p.raw_stride;  // else: beyond the packed column = missing
  uint32_t w[kPkLociPerWarp];
#pragma unroll
  for (int u = 0; u < kPkLociPerWarp; ++u) {
    const int l = l0 + u;
    w[u] = 0xFFFFFFFFu;
    if (l < p.lc && in_col)
      w[u] = __ldg(reinterpret_cast<const uint32_t*>(p.raw + int64_t(l) * p.raw_stride + boff));
  }
#pragma unroll
  for (int u = 0; u < kPkLociPerWarp; ++u) {
    const int l = l0 + u;
    if (l >= p.lc_pad) break;
    uint4 vd, vn, vh;
    decode4(spread4(w[u] & 0xFFu), vd.x, vn.x, vh.x);
    decode4(spread4((w[u] >> 8) & 0xFFu), vd.y, vn.y, vh.y);
    decode4(spread4((w[u] >> 16) & 0xFFu), vd.z, vn.z, vh.z);
    decode4(spread4(w[u] >> 24), vd.w, vn.w, vh.w);
    int8_t* o = p.planes + int64_t(l) * p.n_pad + i0;
    *reinterpret_cast<uint4*>(o) = vd;
    *reinterpret_cast<uint4*>(o + p.plane_stride) = vn;
    if (p.nplanes > 2) *reinterpret_cast<uint4*>(o + 2 * p.plane_stride) = vh;
  }
}

// ------------------------------------------------------------------ scans through code256
// The O(n m) statistics that share this kernel's access pattern:
//   gt_ind_hetero            src/gt_ind_hetero.cpp:12-42           het_n, na_n per individual
//   alt_freq_dip_pseudo_cpp  src/alt_freq_dip_pseudo_cpp.cpp:8-58  n_alt, n_valid per locus
//   big_counts' NA row       R/loci_missingness.R:49-60            n_na per locus
// Unlike the pairwise kernels these DO decode through code256 (SubBMCode256Acc): the host turns
// the 256 doubles into a byte table (0xFF = missing, else the integer dosage) and per-individual
// weights (2 / (3 - ploidy): 2 for diploids, 1 for pseudohaploids), so every sum is an exact
// integer — n_alt is kept in half-allele units.  One pass, 1 byte read per genotype.
constexpr int kScanLociDefault = 256;  // loci per CTA (rounds of 16)

struct ScanParams {
  const uint8_t* lut;     // 256 entries
  const uint8_t* weight;  // n: 2 / (3 - ploidy)
  const uint8_t* ploidy;  // n
  int32_t* het_n;         // n   (atomic)
  int32_t* na_n;          // n   (atomic)
  int32_t* alt2;          // this chunk's loci (atomic): sum dosage * weight
  int32_t* valid;         //                           sum ploidy over non-missing
  int32_t* na_l;          //                           missing count
  int loci_per_cta;       // multiple of 16
  int flags;              // debugging: bit 0 = skip the per-individual atomics
};

// Sums 16 per-lane values over the 32 lanes of a warp with 16 shuffles (instead of 16 x 5): at
// every step a lane hands half of its values to its partner and keeps the other half.  On
// return, lane L (L even) holds in v[0] the warp total of locus (L >> 1).
__device__ __forceinline__ void warp_multi_reduce16(uint32_t (&v)[16], int lane) {
#pragma unroll
  for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < h; ++k) {
      const uint32_t send = up ? v[k] : v[k + h];
      const uint32_t keep = up ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, o);
    }
  }
  v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], 1);
}

__global__ void __launch_bounds__(kDecThreads)
scan_counts_kernel(const DecodeParams p, const ScanParams sp) {
  __shared__ uint8_t lut[256];
  __shared__ uint32_t part[kDecThreads / 32][16];
  lut[threadIdx.x] = sp.lut[threadIdx.x];  // kDecThreads == 256
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nib = (p.n + kDecThreads * 4 - 1) / (kDecThreads * 4);
  const int ib = blockIdx.x % nib, lb = blockIdx.x / nib;
  const int i0 = (ib * kDecThreads + threadIdx.x) * 4;
  const RowSel rs = make_rowsel(p, i0);
  const int nv = min(max(p.n - i0, 0), 4);
  const int n_here = min(kDecThreads * 4, p.n - ib * kDecThreads * 4);  // individuals of this CTA
  // per-thread byte vectors over its 4 individuals: weight 2 / (3 - ploidy), diploid flag, validity
  uint32_t w4 = 0, dip4 = 0, keep01 = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (b < nv) {
      w4 |= uint32_t(sp.weight[i0 + b]) << (8 * b);
      dip4 |= (sp.ploidy[i0 + b] == 2 ? 1u : 0u) << (8 * b);
      keep01 |= 1u << (8 * b);
    }
  __syncthreads();
  int het[4] = {0, 0, 0, 0}, na[4] = {0, 0, 0, 0};
  const int lbase = lb * sp.loci_per_cta;
  for (int round = 0; round < sp.loci_per_cta / 16; ++round) {
    const int l0 = lbase + round * 16;
    if (l0 >= p.lc) break;  // uniform over the CTA
    const bool safe = block_is_safe(p, l0, 16, ib * kDecThreads * 4, (ib + 1) * kDecThreads * 4);
    uint32_t g[16];
    if (safe && l0 + 16 <= p.lc && !p.row_idx) {
      // straight line: no branch between the 16 loads, so all of them are in flight together
      uint32_t w0[16], wx[16], sh[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int64_t col = p.col_idx ? int64_t(__ldg(p.col_idx + l0 + u)) : p.col0 + l0 + u;
        const uint8_t* a = p.raw + col * p.raw_stride + rs.off;
        sh[u] = uint32_t(reinterpret_cast<uintptr_t>(a) & 3u);
        const uint32_t* al = reinterpret_cast<const uint32_t*>(a - sh[u]);
        w0[u] = __ldg(al);
        wx[u] = 0;
        if (lane == 31) wx[u] = __ldg(al + 1);
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint32_t w1 = __shfl_down_sync(0xFFFFFFFFu, w0[u], 1);
        if (lane == 31) w1 = wx[u];
        g[u] = __funnelshift_r(w0[u], w1, sh[u] * 8);
      }
    } else {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int l = l0 + u;
        g[u] = 0x03030303u;
        if (l < p.lc) {
          const int64_t col = p.col_idx ? int64_t(p.col_idx[l]) : p.col0 + l;
          g[u] = safe ? load4<true>(p, rs, col, lane) : load4<false>(p, rs, col, lane);
        }
      }
    }
    // per locus and lane: alt [0,15) | typed [15,23) | typed diploids [23,31) — the warp totals
    // (<= 32512, 128, 128) cannot carry into the next field.  Four individuals at a time (SWAR):
    // table bytes are 0xFF (missing) or <= 127, so bit 7 is the missing flag and byte sums
    // never carry; the weighted sums are 4-way byte dot products.
    uint32_t v[16];
    uint32_t het4 = 0, na4 = 0;  // byte counters, at most 16 per round
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const uint32_t gw = g[u];
      const uint32_t c4 = uint32_t(lut[gw & 0xFFu]) | (uint32_t(lut[(gw >> 8) & 0xFFu]) << 8) |
                          (uint32_t(lut[(gw >> 16) & 0xFFu]) << 16) | (uint32_t(lut[gw >> 24]) << 24);
      const uint32_t real01 = (l0 + u < p.lc) ? keep01 : 0u;  // padding loci / individuals count nothing
      const uint32_t miss01 = (c4 >> 7) & 0x01010101u;
      const uint32_t typed01 = (miss01 ^ 0x01010101u) & real01;
      const uint32_t cm = c4 & (typed01 * 0x7Fu);              // dosage of the typed, 0 elsewhere
      const uint32_t alt = __dp4a(cm, w4, 0u);
      const uint32_t typed = __dp4a(typed01, 0x01010101u, 0u);
      const uint32_t td = __dp4a(typed01, dip4, 0u);
      v[u] = alt | (typed << 15) | (td << 23);
      const uint32_t ne1 = ((cm ^ 0x01010101u) + 0x7F7F7F7Fu) >> 7;  // bit 0 of a byte: dosage != 1
      het4 += (ne1 ^ 0x01010101u) & typed01;                   // (bytes of cm ^ 1 are <= 0x7F: no carries)
      na4 += miss01 & real01;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) { het[b] += (het4 >> (8 * b)) & 0xFFu; na[b] += (na4 >> (8 * b)) & 0xFFu; }
    warp_multi_reduce16(v, lane);
    if ((lane & 1) == 0) part[warp][lane >> 1] = v[0];
    __syncthreads();
    if (threadIdx.x < 16 && l0 + int(threadIdx.x) < p.lc) {
      uint32_t alt = 0, typed = 0, td = 0;
#pragma unroll
      for (int w = 0; w < kDecThreads / 32; ++w) {
        const uint32_t x = part[w][threadIdx.x];
        alt += x & 0x7FFFu; typed += (x >> 15) & 0xFFu; td += x >> 23;
      }
      const int l = l0 + threadIdx.x;
      const int32_t val = int32_t(typed + td);  // sum of ploidy over the typed: 1 each + 1 more per diploid
      const int32_t nal = n_here - int32_t(typed);
      if (nib == 1) { sp.alt2[l] += int32_t(alt); sp.valid[l] += val; sp.na_l[l] += nal; }
      else { atomicAdd(sp.alt2 + l, int32_t(alt)); atomicAdd(sp.valid + l, val); atomicAdd(sp.na_l + l, nal); }
    }
    __syncthreads();
  }
  if (!(sp.flags & 1))
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (b < nv) {
      if (het[b]) atomicAdd(sp.het_n + i0 + b, het[b]);
      if (na[b]) atomicAdd(sp.na_n + i0 + b, na[b]);
    }
}

// Table-free build for the two standard code tables (bigsnpr CODE_012: 0 / 1 / 2, everything else NA;
// CODE_IMPUTE_PRED: 4 / 5 / 6 decode to 0 / 1 / 2 as well).  The per-byte shared-memory lookups and the
// shuffle-assembled loads kept scan_counts_kernel at 0.18 of the HBM figure (instruction bound); here
//   * the table is arithmetic: clamp4 (swar.cuh) maps every byte above 2 to code 3, and for the imputed
//     table clamp4(g) & clamp4(g ^ 0x04040404) folds 4 / 5 / 6 onto 0 / 1 / 2;
//   * a lane loads both aligned words of its four genotypes itself (the second one is an L1 hit), as
//     decode_planes_k4_kernel does: no shuffle, no select;
//   * the 16 per-locus sums of a round are 16 REDUX.SUM (one instruction per locus and warp) instead of a
//     16-shuffle butterfly with its selects.
// Contiguous rows and loci only (launch_scan routes everything else to scan_counts_kernel); same outputs,
// bit for bit.
template <bool kImpute>
__global__ void __launch_bounds__(kDecThreads, 4)
scan_counts_fast_kernel(const DecodeParams p, const ScanParams sp) {
  __shared__ uint32_t part[kDecThreads / 32][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nib = (p.n + kDecThreads * 4 - 1) / (kDecThreads * 4);
  const int ib = blockIdx.x % nib, lb = blockIdx.x / nib;
  const int i0 = (ib * kDecThreads + threadIdx.x) * 4;
  const int nv = min(max(p.n - i0, 0), 4);
  const uint32_t keep = nv >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nv)) - 1u);
  const int n_here = min(kDecThreads * 4, p.n - ib * kDecThreads * 4);
  uint32_t w4 = 0, dip4 = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (b < nv) {
      w4 |= uint32_t(sp.weight[i0 + b]) << (8 * b);
      dip4 |= (sp.ploidy[i0 + b] == 2 ? 1u : 0u) << (8 * b);
    }
  int het[4] = {0, 0, 0, 0}, na[4] = {0, 0, 0, 0};
  const int lbase = lb * sp.loci_per_cta;
  const uint32_t stride = uint32_t(p.raw_stride);
  for (int round = 0; round < sp.loci_per_cta / 16; ++round) {
    const int l0 = lbase + round * 16;
    if (l0 >= p.lc) break;  // uniform over the CTA
    uint32_t g[16];
    // every word this CTA touches in the round lies inside the buffer (3 bytes of slack below, 8 above)?
    const int64_t lo = (p.col0 + l0) * p.raw_stride + p.row0 + int64_t(ib) * kDecThreads * 4 - 3;
    const int64_t hi = (p.col0 + l0 + 15) * p.raw_stride + p.row0 + int64_t(ib + 1) * kDecThreads * 4 + 8;
    const bool full_round = l0 + 16 <= p.lc;
    if (full_round && lo >= 0 && hi <= int64_t(p.raw_end - p.raw)) {
      const uint64_t a0 = uint64_t(reinterpret_cast<uintptr_t>(p.raw + (p.col0 + l0) * p.raw_stride + p.row0 + i0));
      uint32_t w1[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint64_t au;
        asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(au) : "r"(stride), "r"(uint32_t(u)), "l"(a0));
        const uint32_t* al = reinterpret_cast<const uint32_t*>(uintptr_t(au & ~uint64_t(3)));
        g[u] = __ldg(al);
        w1[u] = __ldg(al + 1);
      }
      const uint32_t a8 = uint32_t(a0) << 3, s8 = stride << 3;  // the funnel shift takes its amount mod 32
#pragma unroll
      for (int u = 0; u < 16; ++u) g[u] = __funnelshift_r(g[u], w1[u], a8 + uint32_t(u) * s8) | ~keep;
    } else {
      RowSel rs;
      rs.gather = false;
      rs.keep = keep;
      rs.off = p.row0 + i0;
      rs.ld = i0 < p.n + 4;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        g[u] = 0xFFFFFFFFu;
        if (l0 + u < p.lc) g[u] = load4<false>(p, rs, p.col0 + l0 + u, lane) | ~keep;
      }
    }
    // codes: 0 / 1 / 2 = dosage, 3 = missing (padding individuals and loci were forced to 0xFF above).
    uint32_t mine = 0;           // lane u keeps the warp total of locus u
    uint32_t het4 = 0, na4 = 0;  // byte counters, at most 16 per round
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      uint32_t c = clamp4(g[u]);
      if (kImpute) c &= clamp4(g[u] ^ 0x04040404u);
      const uint32_t hi1 = (c >> 1) & 0x01010101u;           // bit 1 of each code
      const uint32_t na01 = hi1 & c;                          // 3
      const uint32_t het01 = (hi1 ^ 0x01010101u) & c & 0x01010101u;  // 1
      const uint32_t typed01 = na01 ^ 0x01010101u;
      const uint32_t dos = c ^ (na01 * 3u);                   // missing -> 0
      // alt [0,15) | typed [15,23) | typed diploids [23,31): warp totals (<= 32512, 128, 128) cannot carry
      const uint32_t v = __dp4a(dos, w4, 0u) | (__dp4a(typed01, 0x01010101u, 0u) << 15) | (__dp4a(typed01, dip4, 0u) << 23);
      const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, v);
      if (lane == u) mine = tot;
      het4 += het01;
      na4 += na01 & ((full_round || l0 + u < p.lc) ? (keep & 0x01010101u) : 0u);  // padding loci / individuals are not NAs
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) { het[b] += (het4 >> (8 * b)) & 0xFFu; na[b] += (na4 >> (8 * b)) & 0xFFu; }
    if (lane < 16) part[warp][lane] = mine;
    __syncthreads();
    if (threadIdx.x < 16 && l0 + int(threadIdx.x) < p.lc) {
      uint32_t alt = 0, typed = 0, td = 0;
#pragma unroll
      for (int w = 0; w < kDecThreads / 32; ++w) {
        const uint32_t x = part[w][threadIdx.x];
        alt += x & 0x7FFFu; typed += (x >> 15) & 0xFFu; td += x >> 23;
      }
      const int l = l0 + threadIdx.x;
      const int32_t val = int32_t(typed + td);  // sum of ploidy over the typed: 1 each + 1 more per diploid
      const int32_t nal = n_here - int32_t(typed);
      if (nib == 1) { sp.alt2[l] += int32_t(alt); sp.valid[l] += val; sp.na_l[l] += nal; }
      else { atomicAdd(sp.alt2 + l, int32_t(alt)); atomicAdd(sp.valid + l, val); atomicAdd(sp.na_l + l, nal); }
    }
    __syncthreads();
  }
  if (!(sp.flags & 1))
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (b < nv) {
        if (het[b]) atomicAdd(sp.het_n + i0 + b, het[b]);
        if (na[b]) atomicAdd(sp.na_n + i0 + b, na[b]);
      }
}

// Warp-per-locus build of the table-free scan.  scan_counts_fast_kernel spends 45 instructions per word of four
// genotypes (ncu: 0.36 warp instructions per byte, issue-bound at 0.29 of the HBM figure): every word pays for its own
// horizontal reduction (three dp4a, packing, one REDUX) and for two loads and a funnel shift, and every round of
// 16 loci for two block barriers.  Here a WARP owns a locus: a lane takes 8 words of it (32 individuals), adds their
// contributions vertically — the dp4a accumulates — and the warp reduces once per locus (two REDUX for 256 words).
// The words are the ALIGNED words of the buffer: a locus that starts at byte `a` (mod 4) of its first word simply has
// its individuals shifted by `a` byte lanes, and since the 8 warps of a CTA take every 8th locus, `a` is the same for
// all loci of a warp (8 x stride = 0 mod 4) — the lane's weights, ploidies and per-individual counters are laid out
// once for that shift.  Byte lanes outside the locus (the neighbours' genotypes in the first and last word, the
// over-read of the last word block) carry weight 0 and their counters are dropped, so nothing is masked in the
// loop; only words that would leave the buffer take a byte-wise loader.  No shared memory, no block barrier.
// ~17 instructions per word.  Contiguous rows and loci (launch_scan); outputs bit for bit those of the other builds.
constexpr int kScanWarpWords = 8;                        // words per lane and locus
constexpr int kScanWordsPerCta = 32 * kScanWarpWords;   // 256 words = 1 KiB of a locus per CTA

__device__ __forceinline__ uint32_t scan_word_checked(const uint8_t* q, const uint8_t* lo, const uint8_t* hi) {
  uint32_t g = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b) g |= uint32_t((q + b >= lo && q + b < hi) ? q[b] : 0xFFu) << (8 * b);
  return g;
}

template <bool kImpute>
__global__ void __launch_bounds__(kDecThreads, 3)
scan_counts_warp_kernel(const DecodeParams p, const ScanParams sp) {
  constexpr int R = kScanWarpWords;
  constexpr int kWarps = kDecThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwb = ((p.n + 6) / 4 + kScanWordsPerCta - 1) / kScanWordsPerCta;  // a locus touches <= ceil((3 + n) / 4) words
  const int wb = blockIdx.x % nwb, lb = blockIdx.x / nwb;
  __shared__ int32_t tot_s[2][4 * kScanWordsPerCta + 8];  // per-individual totals of the CTA: [het | missing][4 (j - first word) - a + b + 4]
  const int lend = min((lb + 1) * sp.loci_per_cta, p.lc);
  int l = lb * sp.loci_per_cta + warp;
  for (int k = threadIdx.x; k < 2 * (4 * kScanWordsPerCta + 8); k += kDecThreads) (&tot_s[0][0])[k] = 0;
  __syncthreads();
  const uint8_t* col = p.raw + (p.col0 + l) * p.raw_stride + p.row0;
  const int a = int(reinterpret_cast<uintptr_t>(col) & 3u);
  const int j0 = wb * kScanWordsPerCta + lane;           // word r of this lane: j0 + 32 r; individuals 4 j - a + b
  uint32_t w4[R], f4[R], het4[R], na4[R];
  uint32_t pl_lane = 0;                                  // sum of ploidy over this lane's individuals
#pragma unroll
  for (int r = 0; r < R; ++r) {
    w4[r] = f4[r] = het4[r] = na4[r] = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int ind = 4 * (j0 + 32 * r) - a + b;
      if (ind >= 0 && ind < p.n) {
        const uint32_t pl = sp.ploidy[ind];
        w4[r] |= uint32_t(sp.weight[ind]) << (8 * b);
        f4[r] |= (1u + 64u * pl) << (8 * b);             // one dp4a: missing count + 64 x missing ploidy
        pl_lane += pl;
      }
    }
  }
  const uint8_t* wp = col - a + 4 * int64_t(j0);         // this lane's first (aligned) word of the locus
  const int64_t step = int64_t(kWarps) * p.raw_stride;
  auto load = [&](const uint8_t* q, uint32_t (&g)[R]) {
    const uint8_t* q0 = q - 4 * lane;                    // the warp's first byte (uniform)
    if (q0 >= p.raw && q0 + 4 * kScanWordsPerCta <= p.raw_end) {
#pragma unroll
      for (int r = 0; r < R; ++r) g[r] = __ldg(reinterpret_cast<const uint32_t*>(q + 128 * r));
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) g[r] = scan_word_checked(q + 128 * r, p.raw, p.raw_end);
    }
  };
  auto consume = [&](const uint32_t (&gin)[R], int locus) {
    uint32_t alt = 0, nn = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint32_t g = gin[r];
      uint32_t t = (g & 0x7F7F7F7Fu) + 0x7D7D7D7Du;      // bit 7 set iff (byte & 127) >= 3
      uint32_t m7 = (t | g) & 0x80808080u;               // bit 7 set iff byte > 2
      if (kImpute) {                                     // 4 / 5 / 6 -> 0 / 1 / 2, then classify again
        g ^= m7 >> 5;
        t = (g & 0x7F7F7F7Fu) + 0x7D7D7D7Du;
        m7 = (t | g) & 0x80808080u;
      }
      const uint32_t na01 = m7 >> 7;
      const uint32_t dos = g & ~((m7 - na01) | m7);      // missing bytes -> 0; the typed ones are their dosage
      alt = __dp4a(dos, w4[r], alt);                     // x weight
      nn = __dp4a(na01, f4[r], nn);
      het4[r] += dos & 0x01010101u;                      // dosage 1
      na4[r] += na01;
    }
    // per lane: nn & 63 = missing individuals (<= 32), nn >> 6 = their ploidy (<= 64)
    const uint32_t alt_t = __reduce_add_sync(0xFFFFFFFFu, alt);
    const uint32_t t2 = __reduce_add_sync(0xFFFFFFFFu, (nn & 63u) | ((pl_lane - (nn >> 6)) << 16));
    if (lane == 0) {
      const int32_t val = int32_t(t2 >> 16), nal = int32_t(t2 & 0xFFFFu);
      // reductions, also where this warp is the only writer (nwb == 1): a load - add - store would stall it three times
      atomicAdd(sp.alt2 + locus, int32_t(alt_t)); atomicAdd(sp.valid + locus, val); atomicAdd(sp.na_l + locus, nal);
    }
  };
  // two loci in flight per warp: 16 independent loads per lane
  for (; l < lend; l += 2 * kWarps, wp += 2 * step) {
    uint32_t g0[R], g1[R];
    const bool two = l + kWarps < lend;
    load(wp, g0);
    if (two) load(wp + step, g1);
    consume(g0, l);
    if (two) consume(g1, l + kWarps);
  }
  // per-individual counters: summed over the CTA's warps in shared memory (the warps' byte lanes are shifted against
  // each other by their `a`), then one global reduction per individual and CTA — with every warp reducing for itself
  // the 929 counters of C2 took 14.5 M same-address atomics and the kernel 0.57 ms
  if (!(sp.flags & 1)) {
    if (lb * sp.loci_per_cta + warp < lend) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int k = 4 * (lane + 32 * r) - a + b + 4;
          const int h = int((het4[r] >> (8 * b)) & 0xFFu), m = int((na4[r] >> (8 * b)) & 0xFFu);
          if (h) atomicAdd(&tot_s[0][k], h);
          if (m) atomicAdd(&tot_s[1][k], m);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 4 * kScanWordsPerCta + 8; k += kDecThreads) {
      const int ind = 4 * wb * kScanWordsPerCta + k - 4;
      if (ind >= 0 && ind < p.n) {
        const int h = tot_s[0][k], m = tot_s[1][k];
        if (h) atomicAdd(sp.het_n + ind, h);
        if (m) atomicAdd(sp.na_n + ind, m);
      }
    }
  }
}

int launch_scan(Ctx* ctx, const DecodeArgs& a, const ScanArgs& s, cudaStream_t st, const Err& E) {
  if (a.n <= 0 || a.lc <= 0 || a.n > INT32_MAX / 2 || a.lc > INT32_MAX / 2)
    return E.fail(TPG_ERR_ARG, "scan: bad geometry n=%lld lc=%lld", (long long)a.n, (long long)a.lc);
  DecodeParams p;
  p.raw = a.raw;
  p.raw_end = a.raw + a.raw_bytes;
  p.raw_stride = a.raw_stride;
  p.row_idx = a.row_idx;
  p.row0 = a.row0;
  p.col_idx = a.col_idx;
  p.col0 = a.col0;
  p.n = int(a.n);
  p.n_pad = int(a.n);
  p.lc = int(a.lc);
  p.lc_pad = int(a.lc);
  p.planes = nullptr;
  p.plane_stride = 0;
  p.nplanes = 0;
  ScanParams sp;
  sp.lut = s.lut; sp.weight = s.weight; sp.ploidy = s.ploidy;
  sp.het_n = s.het_n; sp.na_n = s.na_n; sp.alt2 = s.alt2; sp.valid = s.valid; sp.na_l = s.na_l;
  sp.loci_per_cta = kScanLociDefault;
  const tpg_config C = cfg();
  if (C.scan_loci_per_cta > 0) sp.loci_per_cta = std::max(16, C.scan_loci_per_cta / 16 * 16);
  sp.flags = C.scan_no_indiv ? 1 : 0;
  const int64_t blocks = ((a.n + kDecThreads * 4 - 1) / (kDecThreads * 4)) * ((a.lc + sp.loci_per_cta - 1) / sp.loci_per_cta);
  if (blocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "scan: chunk too large for one launch");
  // the table-free build: a standard code table, contiguous rows and loci, a stride the 32-bit address
  // arithmetic of the fast loads can take
  // tpg_config.scan_table: 0 = auto, 1 = the lookup-table kernel, 2 = the 16-locus-round kernel, 3 = the warp-per-locus kernel
  const bool contiguous = s.std_table != 0 && C.scan_table != 1 && !a.row_idx && !a.col_idx && !a.packed;
  // warp per locus: worth it once a locus fills most of a CTA's 256 words (n >= 512); its byte counters hold 255 loci per warp
  if (contiguous && (C.scan_table == 3 || (C.scan_table == 0 && a.n >= 512))) {
    const int64_t nwb = ((a.n + 6) / 4 + kScanWordsPerCta - 1) / kScanWordsPerCta;
    int lpc;
    if (C.scan_loci_per_cta > 0) lpc = std::min(2032, std::max(8, C.scan_loci_per_cta / 8 * 8));
    else {  // whole waves of 3 CTAs per SM, at most 1024 loci per CTA (the per-individual atomics are paid once per CTA)
      const int64_t slots = int64_t(ctx->num_sms) * 3;
      const int64_t waves = std::max<int64_t>(1, (nwb * a.lc + 1024 * slots - 1) / (1024 * slots));
      const int64_t groups = std::max<int64_t>(1, waves * slots / nwb);
      lpc = int(std::min<int64_t>(2032, std::max<int64_t>(64, ((a.lc + groups - 1) / groups + 7) / 8 * 8)));
    }
    sp.loci_per_cta = lpc;
    const int64_t wblocks = nwb * ((a.lc + lpc - 1) / lpc);
    if (wblocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "scan: chunk too large for one launch");
    if (s.std_table == 2) scan_counts_warp_kernel<true><<<unsigned(wblocks), kDecThreads, 0, st>>>(p, sp);
    else scan_counts_warp_kernel<false><<<unsigned(wblocks), kDecThreads, 0, st>>>(p, sp);
    TPG_CUDA_OK(cudaGetLastError(), E);
    ctx->launches++;
    return TPG_OK;
  }
  const bool fast = contiguous && a.raw_stride < (int64_t(1) << 27);
  if (fast && s.std_table == 2) scan_counts_fast_kernel<true><<<unsigned(blocks), kDecThreads, 0, st>>>(p, sp);
  else if (fast) scan_counts_fast_kernel<false><<<unsigned(blocks), kDecThreads, 0, st>>>(p, sp);
  else scan_counts_kernel<<<unsigned(blocks), kDecThreads, 0, st>>>(p, sp);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

int launch_decode(Ctx* ctx, const DecodeArgs& a, cudaStream_t st, const Err& E) {
  if (a.n_pad % 128 || a.lc_pad % 16 || a.n > a.n_pad || a.lc > a.lc_pad || a.n <= 0 || a.lc <= 0)
    return E.fail(TPG_ERR_ARG, "decode: bad geometry n=%lld n_pad=%lld lc=%lld lc_pad=%lld",
                  (long long)a.n, (long long)a.n_pad, (long long)a.lc, (long long)a.lc_pad);
  if (a.n_pad > INT32_MAX || a.lc_pad > INT32_MAX)
    return E.fail(TPG_ERR_ARG, "decode: chunk dimensions exceed int32");
  DecodeParams p;
  p.raw = a.raw;
  p.raw_end = a.raw + a.raw_bytes;
  p.raw_stride = a.raw_stride;
  p.row_idx = a.row_idx;
  p.row0 = a.row0;
  p.col_idx = a.col_idx;
  p.col0 = a.col0;
  p.n = int(a.n);
  p.n_pad = int(a.n_pad);
  p.lc = int(a.lc);
  p.lc_pad = int(a.lc_pad);
  p.planes = a.planes;
  p.plane_stride = a.plane_stride;
  p.nplanes = a.nplanes;
  if (a.k4) {
    if (a.lc_pad % 256 || a.plane_stride % 16 || (reinterpret_cast<uintptr_t>(a.planes) & 15))
      return E.fail(TPG_ERR_ARG, "decode(k4): lc_pad must be a multiple of 256, planes 16-byte aligned");
    if (a.packed && (a.row_idx || a.col_idx || a.row0 || a.col0))
      return E.fail(TPG_ERR_ARG, "decode(k4, pk2): packed chunks are ungathered");
    const int64_t blocks = (a.n_pad / 128) * (a.lc_pad / 256);
    if (blocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "decode: chunk too large for one launch");
    const bool v1 = cfg().decode_k4_v1 != 0;  // the first build, kept as a cross-check
    if (a.packed) decode_planes_k4_kernel<true><<<unsigned(blocks), 256, 0, st>>>(p);
    else if (v1) decode_planes_k4_v1_kernel<<<unsigned(blocks), 256, 0, st>>>(p);
    else decode_planes_k4_kernel<false><<<unsigned(blocks), 256, 0, st>>>(p);
  } else if (a.packed) {
    if (a.k_major || a.row_idx || a.col_idx || a.row0 || a.col0 || a.raw_stride % 16 || a.plane_stride % 16 ||
        (reinterpret_cast<uintptr_t>(a.raw) & 15) || (reinterpret_cast<uintptr_t>(a.planes) & 15))
      return E.fail(TPG_ERR_ARG, "decode(pk2): packed chunks are MN-major, ungathered and 16-byte aligned");
    const int64_t warps = ((a.n_pad + 511) / 512) * ((a.lc_pad + kPkLociPerWarp - 1) / kPkLociPerWarp);
    const int64_t blocks = (warps * 32 + kDecThreads - 1) / kDecThreads;
    if (blocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "decode: chunk too large for one launch");
    decode_planes_pk2_kernel<<<unsigned(blocks), kDecThreads, 0, st>>>(p);
  } else if (a.k_major) {
    if (a.lc_pad % 128) return E.fail(TPG_ERR_ARG, "decode(K-major): lc_pad must be a multiple of 128");
    const int64_t blocks = (a.n_pad / 128) * (a.lc_pad / 128);
    if (blocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "decode: chunk too large for one launch");
    decode_planes_k_kernel<<<unsigned(blocks), 256, 0, st>>>(p);
  } else {
    const int64_t blocks = ((a.n_pad + kDecThreads * 4 - 1) / (kDecThreads * 4)) *
                           ((a.lc_pad + kDecLociPerBlock - 1) / kDecLociPerBlock);
    if (blocks > INT32_MAX) return E.fail(TPG_ERR_ARG, "decode: chunk too large for one launch");
    decode_planes_mn_kernel<<<unsigned(blocks), kDecThreads, 0, st>>>(p);
  }
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

}  // namespace tpg
