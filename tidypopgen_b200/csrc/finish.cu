// finish_*: int32 count matrices -> the reference's FP64 statistics, in the reference's
// operation order (no FMA contraction: every step is an explicit _rn intrinsic).
//
//   KING  R/snp_king.R:82-91   K/(2*pmin(Ni,Nj)) + 0.5 - 0.25*(Ni+Nj)/pmin(Ni,Nj)
//         evaluated left to right as ((K/(2p)) + 0.5) - ((0.25*(Ni+Nj))/p); p = 0 gives
//         NaN / +-Inf by IEEE-754 exactly as R does (pinned by test_autoplot_pairwise.R:17-24)
//   IBS   R/snp_ibs.R:84-101   raw counts | IBS/valid | (IBS/valid)*m
//   AS    R/snp_allele_sharing.R:77-80   0.5*(1 + num/denom); denom==0 -> NA_real_
//   GRM   R/pairwise_grm.R:42-50         mb = mean(offdiag(AS), na.rm=TRUE); ((AS-mb)/(1-mb))*2
//
// Count identities (d = g2-g0, NM = valid, H1 = het; DESIGN.md "Algebra"):
//   KING numerator  K  = DD - NN + HN(i,j) + HN(j,i)        N_Aa_i = HN(i,j)
//   IBS             ibs = NN + HH + DD                       valid_n = 2 NN
//   allele sharing  num = DD (+ n_short, the pad quirk of src/snp_as.cpp:56-64), denom = NN
//
// One CTA per 32x32 upper-triangle tile: reads are coalesced along i (the count matrices
// are column-major), the mirror tile goes through shared memory so both output tiles are
// written coalesced.  HBM-bound: 4 B per count matrix read + 8 B per output written.
#include "internal.h"
#include "ddsum.cuh"
#include "../../include/tpg_synth.h"

namespace tpg {

struct FinishParams {
  int stat;
  const int32_t* counts;
  int64_t ld;
  int n;
  double m;
  int n_short;
  int shard_rank, shard_world;
  const uint8_t* owner;
  double* out[4];  // out0, out1, aux0, aux1 (nullable)
};

__device__ __forceinline__ double na_real() {
  return __longlong_as_double(0x7FF00000000007A2ll);  // R's NA_real_ (payload 1954)
}

__global__ void __launch_bounds__(256) finish_kernel(const FinishParams p) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi > bj) return;
  if (p.shard_world > 1) {
    const int64_t nb = (p.ld + 255) / 256;
    if (!block_owned((int64_t(bi) * 32) / 256, (int64_t(bj) * 32) / 256, nb, p.shard_rank,
                     p.shard_world, p.owner))
      return;
  }
  __shared__ int32_t tN[32][33];
  __shared__ double tR[4][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int I0 = bi * 32, J0 = bj * 32;
  const int64_t ms = p.ld * p.ld;
  const int32_t* DD = p.counts + TPG_MAT_DD * ms;
  const int32_t* NN = p.counts + TPG_MAT_NN * ms;
  const int32_t* HH = p.counts + TPG_MAT_HH * ms;
  const int32_t* HN = p.counts + TPG_MAT_HN * ms;
  const bool king = p.stat == TPG_STAT_KING;

  if (king) {
    // HN(j, i) for the tile, read coalesced along j and transposed through smem.
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int a = ty + 8 * k;
      tN[a][tx] = HN[int64_t(J0 + tx) + int64_t(I0 + a) * p.ld];
    }
  }
  __syncthreads();

#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = I0 + tx, jl = ty + 8 * k, j = J0 + jl;
    double v[4] = {0, 0, 0, 0};  // values for (i, j)
    double w3 = 0;               // aux1 for (j, i) where it differs (KING N_Aa)
    if (i < p.n && j < p.n && i <= j) {
      const int64_t e = int64_t(i) + int64_t(j) * p.ld;
      const int32_t dd = DD[e], nn = NN[e];
      if (king) {
        const int32_t ni = HN[e], nj = tN[tx][jl];
        // (count matrices are int32, each <= m; their combinations are formed in FP64: exact, no wrap)
        const double K = ((double(dd) - double(nn)) + double(ni)) + double(nj);
        const double Ni = double(ni), Nj = double(nj);
        const double pm = fmin(Ni, Nj);
        const double t1 = __ddiv_rn(K, __dmul_rn(2.0, pm));
        const double t2 = __dadd_rn(t1, 0.5);
        const double t3 = __dmul_rn(0.25, __dadd_rn(Ni, Nj));
        v[0] = __dsub_rn(t2, __ddiv_rn(t3, pm));
        v[2] = K;
        v[3] = Ni;
        w3 = Nj;
      } else if (p.stat == TPG_STAT_AS || p.stat == TPG_STAT_GRM) {
        const double num = double(dd) + double(p.n_short), den = double(nn);
        v[0] = nn == 0 ? na_real() : __dmul_rn(0.5, __dadd_rn(1.0, __ddiv_rn(num, den)));
        v[2] = num;
        v[3] = den;
      } else {
        const double ibs = (double(nn) + double(HH[e])) + double(dd), valid = 2.0 * double(nn);
        if (p.stat == TPG_STAT_IBS_RAW) {
          v[0] = ibs;
          v[1] = valid;
        } else {
          const double prop = __ddiv_rn(ibs, valid);
          v[0] = p.stat == TPG_STAT_IBS_ADJUSTED ? __dmul_rn(prop, p.m) : prop;
          v[2] = ibs;
          v[3] = valid;
        }
      }
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (p.out[o]) p.out[o][int64_t(i) + int64_t(j) * p.n] = v[o];
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) tR[o][jl][tx] = (o == 3 && king) ? w3 : v[o];
  }
  __syncthreads();
  // mirror: element (row = J0+tx, col = I0+a) takes the value computed for (i = I0+a, j = J0+tx)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int a = ty + 8 * k;
    const int i = I0 + a, j = J0 + tx;
    if (i < p.n && j < p.n && i < j) {
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (p.out[o]) p.out[o][int64_t(j) + int64_t(i) * p.n] = tR[o][tx][a];
    }
  }
}

// ------------------------------------------------------------------ GRM (Weir & Goudet 2017)
constexpr int kRedBlocks = 1024;

// mb = mean of the off-diagonal, non-NA allele sharing (R: mean(x[off-diag], na.rm = TRUE), a
// long-double two-pass mean).  Accumulated in double-double (ddsum.cuh) over a fixed tree — the
// same arithmetic as wg17.cu / relate.cu, so GRM, indiv_inbreeding and pop_fst see ONE mb, correct
// to the last double and identical from run to run.
__global__ void __launch_bounds__(256)
grm_partial_kernel(const double* as, int64_t n, double* p_hi, double* p_lo, unsigned long long* pcnt) {
  __shared__ double s_hi[8], s_lo[8];
  __shared__ unsigned long long sc[8];
  dd acc = dd_zero();
  unsigned long long c = 0;
  const int64_t total = n * n;
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int64_t i = e % n, j = e / n;
    const double x = as[e];
    if (i != j && !isnan(x)) { acc = dd_add_d(acc, x); ++c; }
  }
  acc = dd_warp_sum(acc);
  c = (unsigned long long)i64_warp_sum((long long)c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_hi[warp] = acc.hi; s_lo[warp] = acc.lo; sc[warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    dd t = dd_zero();
    unsigned long long cc = 0;
    for (int w = 0; w < 8; ++w) { t = dd_add(t, dd{s_hi[w], s_lo[w]}); cc += sc[w]; }
    p_hi[blockIdx.x] = t.hi; p_lo[blockIdx.x] = t.lo; pcnt[blockIdx.x] = cc;
  }
}

__global__ void __launch_bounds__(32)
grm_mean_kernel(const double* p_hi, const double* p_lo, const unsigned long long* pcnt, int nb, double* mb) {
  dd acc = dd_zero();
  long long c = 0;
  for (int b = threadIdx.x; b < nb; b += 32) { acc = dd_add(acc, dd{p_hi[b], p_lo[b]}); c += (long long)pcnt[b]; }
  acc = dd_warp_sum(acc);
  c = i64_warp_sum(c);
  if (threadIdx.x == 0) *mb = dd_mean(acc, double(c));  // 0 entries -> NaN like R's mean(numeric(0))
}

__global__ void __launch_bounds__(256)
grm_apply_kernel(const double* as, int64_t total, const double* mbp, double* out) {
  const double mb = *mbp;
  const double den = __dsub_rn(1.0, mb);
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const double x = as[e];
    // NA_real_ stays NA_real_ (payload kept), as NA - mb does in R.
    out[e] = isnan(x) ? x : __dmul_rn(__ddiv_rn(__dsub_rn(x, mb), den), 2.0);
  }
}

int launch_grm(Ctx* ctx, const double* d_as, int64_t n, double* d_grm, double* mb_host,
               cudaStream_t st, const Err& E) {
  int rc = ensure_slot(&ctx->d_red, kRedBlocks * (2 * sizeof(double) + sizeof(unsigned long long)) + 16,
                       false, "reduction scratch", E);
  if (rc) return rc;
  double* d_ph = static_cast<double*>(ctx->d_red.p);
  double* d_pl = d_ph + kRedBlocks;
  unsigned long long* d_pc = reinterpret_cast<unsigned long long*>(d_pl + kRedBlocks);
  double* d_mb = reinterpret_cast<double*>(d_pc + kRedBlocks);
  const int64_t total = n * n;
  const int nb = int(std::min<int64_t>(kRedBlocks, (total + 255) / 256));
  grm_partial_kernel<<<nb, 256, 0, st>>>(d_as, n, d_ph, d_pl, d_pc);
  grm_mean_kernel<<<1, 32, 0, st>>>(d_ph, d_pl, d_pc, nb, d_mb);
  const int gb = int(std::min<int64_t>(148 * 16, (total + 255) / 256));
  grm_apply_kernel<<<gb, 256, 0, st>>>(d_as, total, d_mb, d_grm);
  ctx->launches += 3;
  cudaError_t e = cudaGetLastError();
  double mb = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&mb, d_mb, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return E.fail(TPG_ERR_CUDA, "grm kernels failed: %s", cudaGetErrorString(e));
  if (mb_host) *mb_host = mb;
  return TPG_OK;
}

int launch_finish(Ctx* ctx, const FinishArgs& a, cudaStream_t st, const Err& E) {
  if (a.n <= 0 || a.n > a.ld) return E.fail(TPG_ERR_ARG, "finish: bad n / ld");
  if (!a.out0) return E.fail(TPG_ERR_ARG, "finish: out0 is required");
  if (a.stat == TPG_STAT_IBS_RAW && !a.out1) return E.fail(TPG_ERR_ARG, "finish: raw_counts needs out1");
  if (a.stat == TPG_STAT_GRM && a.shard_world > 1)
    return E.fail(TPG_ERR_ARG, "finish: GRM needs the whole allele-sharing matrix; finish AS per "
                               "tile shard, gather, then call tpg_grm_from_as");
  FinishParams p;
  p.stat = a.stat;
  p.counts = a.counts;
  p.ld = a.ld;
  p.n = int(a.n);
  p.m = double(a.m);
  p.n_short = int(a.n_short);
  p.shard_rank = a.shard_rank;
  p.shard_world = a.shard_world;
  p.owner = a.owner;
  p.out[0] = a.out0;
  p.out[1] = a.out1;
  p.out[2] = a.aux0;
  p.out[3] = a.aux1;
  if (a.stat == TPG_STAT_GRM) {  // AS first (aux0 may keep it), then the scalar correction
    p.out[1] = nullptr;
    p.out[2] = nullptr;
    p.out[3] = nullptr;
  }
  const unsigned nt = unsigned((a.n + 31) / 32);
  finish_kernel<<<dim3(nt, nt), 256, 0, st>>>(p);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  if (a.stat == TPG_STAT_GRM) {
    if (a.aux0)
      TPG_CUDA_OK(cudaMemcpyAsync(a.aux0, a.out0, sizeof(double) * a.n * a.n,
                                  cudaMemcpyDeviceToDevice, st), E);
    return launch_grm(ctx, a.out0, a.n, a.out0, nullptr, st, E);
  }
  return TPG_OK;
}

// ------------------------------------------------------------------ count reduction (multi-GPU)
__global__ void __launch_bounds__(256) add_counts_kernel(int4* a, const int4* b, int64_t n4) {
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < n4; e += int64_t(gridDim.x) * 256) {
    int4 x = a[e];
    const int4 y = b[e];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    a[e] = x;
  }
}

int launch_add_counts(Ctx* ctx, int32_t* a, const int32_t* b, int64_t elems, cudaStream_t st, const Err& E) {
  const int64_t n4 = elems / 4;  // ld is a multiple of 256: always divisible
  const int gb = int(std::min<int64_t>(148 * 16, (n4 + 255) / 256));
  add_counts_kernel<<<gb, 256, 0, st>>>(reinterpret_cast<int4*>(a), reinterpret_cast<const int4*>(b), n4);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

// Block-wise a += b over what a job accumulated: the matrices of prod_mask, and of the symmetric ones only the
// 256-blocks on or above the diagonal — a third less NVLink traffic than whole matrices for KING (b is a
// peer's buffer, read straight out of its HBM).  One CTA per (matrix, block).
__global__ void __launch_bounds__(256)
add_counts_blocks_kernel(int32_t* a, const int32_t* b, int64_t ld, int prod_mask) {
  const int mat = blockIdx.z;
  if (!(prod_mask & (1 << mat))) return;
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (mat != TPG_MAT_HN && bi > bj) return;
  const int64_t off = int64_t(mat) * ld * ld + int64_t(bj) * 256 * ld + int64_t(bi) * 256;
  int4* pa = reinterpret_cast<int4*>(a + off);
  const int4* pb = reinterpret_cast<const int4*>(b + off);
  const int64_t ld4 = ld / 4;
  const int q = threadIdx.x & 63, c0 = threadIdx.x >> 6;
#pragma unroll 4
  for (int c = c0; c < 256; c += 4) {
    int4 x = pa[int64_t(c) * ld4 + q];
    const int4 y = pb[int64_t(c) * ld4 + q];
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    pa[int64_t(c) * ld4 + q] = x;
  }
}

int launch_add_counts_blocks(Ctx* ctx, int32_t* a, const int32_t* b, int64_t n, int prod_mask, cudaStream_t st, const Err& E) {
  const int64_t ld = (n + 255) / 256 * 256;
  const int nb = int(ld / 256);
  if (nb > 65535) return E.fail(TPG_ERR_ARG, "add_counts: too many blocks");
  add_counts_blocks_kernel<<<dim3(nb, nb, TPG_NUM_COUNT_MATS), 256, 0, st>>>(a, b, ld, prod_mask);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

// ------------------------------------------------------------------ zeroing the count matrices
// A job's accumulators start at zero.  A plain memset of the whole buffer (16 n_pad^2 bytes: 40 GB at
// n = 50,000) writes four full matrices; what the Gram epilogue and finish ever touch is less: only the
// products of the job (prod_mask), only 256 x 256 blocks on or above the diagonal of the symmetric ones,
// and — tile-sharded — only the blocks this rank owns (an eighth of them on 8 GPUs).  One CTA per
// (matrix, block): 256 columns x 1 KiB.
__global__ void __launch_bounds__(256)
zero_counts_kernel(int32_t* counts, int64_t ld, int nblocks, int prod_mask, int rank, int world, const uint8_t* owner) {
  const int mat = blockIdx.z;
  if (!(prod_mask & (1 << mat))) return;
  const int bi = blockIdx.y, bj = blockIdx.x;   // row block, column block
  if (mat != TPG_MAT_HN && bi > bj) return;      // symmetric products live in the upper blocks
  if (!block_owned(min(bi, bj), max(bi, bj), nblocks, rank, world, owner)) return;
  int4* base = reinterpret_cast<int4*>(counts + int64_t(mat) * ld * ld + int64_t(bj) * 256 * ld + int64_t(bi) * 256);
  const int64_t ld4 = ld / 4;
  const int4 z = make_int4(0, 0, 0, 0);
  // 256 columns x 64 int4: thread t writes int4 (t & 63) of columns (t >> 6) + 4 k
  const int q = threadIdx.x & 63, c0 = threadIdx.x >> 6;
#pragma unroll 8
  for (int c = c0; c < 256; c += 4) base[int64_t(c) * ld4 + q] = z;
}

int launch_zero_counts(Ctx* ctx, int32_t* counts, int64_t n, int prod_mask, int shard_rank, int shard_world,
                       const uint8_t* d_owner, cudaStream_t st, const Err& E) {
  const int64_t ld = (n + 255) / 256 * 256;
  const int nb = int(ld / 256);
  if (nb > 65535) return E.fail(TPG_ERR_ARG, "zero_counts: too many blocks");
  zero_counts_kernel<<<dim3(nb, nb, TPG_NUM_COUNT_MATS), 256, 0, st>>>(counts, ld, nb, prod_mask, shard_rank,
                                                                      shard_world > 1 ? shard_world : 1, d_owner);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

// ------------------------------------------------------------------ synthetic genotypes
__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* raw, int64_t nrow, int64_t ncol, int64_t col_offset, uint64_t seed,
             uint32_t miss_thr, int mode) {
  const int64_t total = nrow * ncol;
  for (int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x; e < total; e += int64_t(gridDim.x) * 256) {
    const int64_t i = e % nrow, l = e / nrow;
    raw[e] = tpg_synth_byte(seed, uint64_t(l + col_offset), uint64_t(i), miss_thr, mode);
  }
}

int launch_synth(Ctx* ctx, uint8_t* d_raw, int64_t nrow, int64_t ncol, int64_t col_offset,
                 uint64_t seed, double missing_rate, int mode, cudaStream_t st, const Err& E) {
  if (missing_rate < 0 || missing_rate >= 1) return E.fail(TPG_ERR_ARG, "synth: missing_rate out of range");
  const uint32_t thr = uint32_t(missing_rate * 4294967296.0);
  const int64_t total = nrow * ncol;
  const int gb = int(std::min<int64_t>(148 * 32, (total + 255) / 256));
  synth_kernel<<<gb, 256, 0, st>>>(d_raw, nrow, ncol, col_offset, seed, thr, mode);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches++;
  return TPG_OK;
}

}  // namespace tpg
