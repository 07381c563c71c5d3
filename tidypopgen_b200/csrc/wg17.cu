// Weir & Goudet (2017) consumers of the allele-sharing matrix, computed where the matrix
// already is (HBM) instead of pulling n x n doubles into R:
//   indiv_inbreeding   R/indiv_inbreeding.R:56-81   ((Mii*2 - 1) - mb) / (1 - mb)
//   pop_fst            R/pop_fst.R:31-76            (Fsts - Mb) / (1 - Mb)
//   pop_fis_wg17       R/pop_fis.R:137-191          mean_pop((Mii*2-1 - Fsts) / (1 - Fsts))
// All three only need (a) diag(AS) and (b) the mean of AS over each (population, population)
// block with the diagonal and NAs left out — an npop x npop table of (sum, count).
//
// Two HBM-bound passes, both in fixed order (run-to-run identical):
//   1. group_colsums: one warp per column j; for every population g the lanes stride over
//      the rows of g (individuals are visited through a population-sorted permutation) and
//      the 32 partial double-double sums are combined by a butterfly -> S[g][j], C[g][j].
//   2. group_pairs:   one warp per (gi, gj); lanes stride over the columns of gj.
// then a single-warp finish evaluates the reference formulas in R's operator order.
// Traffic: 8 B read per matrix element, once.
#include <algorithm>
#include <cmath>
#include <vector>
#include "ddsum.cuh"
#include "internal.h"

namespace tpg {

__global__ void __launch_bounds__(256)
group_colsums_kernel(const double* __restrict__ as, int64_t ld, int n, int npop,
                     const int32_t* __restrict__ perm, const int32_t* __restrict__ seg,
                     double* __restrict__ s_hi, double* __restrict__ s_lo, int32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const double* col = as + int64_t(j) * ld;
  for (int g = 0; g < npop; ++g) {
    dd acc = dd_zero();
    long long c = 0;
    for (int r = seg[g] + lane; r < seg[g + 1]; r += 32) {
      const int i = perm[r];
      if (i == j) continue;                 // diag(Mij) <- NA
      const double v = col[i];
      if (v != v) continue;                 // na.rm = TRUE
      acc = dd_add_d(acc, v);
      ++c;
    }
    acc = dd_warp_sum(acc);
    c = i64_warp_sum(c);
    if (lane == 0) {
      s_hi[int64_t(g) * n + j] = acc.hi;
      s_lo[int64_t(g) * n + j] = acc.lo;
      cnt[int64_t(g) * n + j] = int32_t(c);
    }
  }
}

__global__ void __launch_bounds__(256)
group_pairs_kernel(int n, int npop, const int32_t* __restrict__ perm, const int32_t* __restrict__ seg,
                   const double* __restrict__ s_hi, const double* __restrict__ s_lo,
                   const int32_t* __restrict__ cnt, double* __restrict__ p_hi, double* __restrict__ p_lo,
                   double* __restrict__ p_cnt) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= npop * npop) return;
  const int gi = w / npop, gj = w % npop;   // rows of population gi, columns of population gj
  dd acc = dd_zero();
  long long c = 0;
  for (int r = seg[gj] + lane; r < seg[gj + 1]; r += 32) {
    const int j = perm[r];
    acc = dd_add(acc, dd{s_hi[int64_t(gi) * n + j], s_lo[int64_t(gi) * n + j]});
    c += cnt[int64_t(gi) * n + j];
  }
  acc = dd_warp_sum(acc);
  c = i64_warp_sum(c);
  if (lane == 0) { p_hi[w] = acc.hi; p_lo[w] = acc.lo; p_cnt[w] = double(c); }
}

// mean(x, na.rm = TRUE) of a short vector held by one warp (lanes stride), in double-double.
__device__ double warp_mean_narm(const double* x, int len, int lane) {
  dd acc = dd_zero();
  long long c = 0;
  for (int k = lane; k < len; k += 32) {
    const double v = x[k];
    if (v != v) continue;
    acc = dd_add_d(acc, v);
    ++c;
  }
  acc = dd_warp_sum(acc);
  c = i64_warp_sum(c);
  return dd_mean(acc, double(c));
}

struct Wg17Out {
  double* inbreeding;  // n or null
  double* fst;         // npop + 1 or null
  double* fis;         // npop + 1 or null
  double* mb;          // [0] = mean off-diagonal AS over everybody, [1] = Mb between populations
  double* scratch;     // npop (Fsts) + n (Fi)
};

__global__ void __launch_bounds__(32)
wg17_finish_kernel(const double* __restrict__ as, int64_t ld, int n, int npop,
                   const int32_t* __restrict__ perm, const int32_t* __restrict__ seg,
                   const double* __restrict__ p_hi, const double* __restrict__ p_lo,
                   const double* __restrict__ p_cnt, Wg17Out o) {
  const int lane = threadIdx.x;
  double* Fsts = o.scratch;       // mean within population p (diagonal out)
  double* Fi = o.scratch + npop;  // per-individual values of one population at a time
  // ---- mb over all individuals (indiv_inbreeding.R:76-78): all blocks together
  dd tot = dd_zero();
  double tc = 0;
  for (int w = 0; w < npop * npop; ++w) {  // fixed order, every lane the same
    tot = dd_add(tot, dd{p_hi[w], p_lo[w]});
    tc += p_cnt[w];
  }
  const double mb_all = dd_mean(tot, tc);
  for (int p = lane; p < npop; p += 32)
    Fsts[p] = dd_mean(dd{p_hi[p * npop + p], p_lo[p * npop + p]}, p_cnt[p * npop + p]);
  __syncwarp();
  // ---- Mb (pop_fst.R:52-63): Mb <- Mb + mMij[i, j] for i in 2..npop, j in 1..i-1, in that order
  double Mb = 0;
  for (int i = 1; i < npop; ++i)
    for (int j = 0; j < i; ++j)
      Mb = Mb + dd_mean(dd{p_hi[i * npop + j], p_lo[i * npop + j]}, p_cnt[i * npop + j]);
  Mb = Mb * 2 / double(int64_t(npop) * (npop - 1));
  if (lane == 0 && o.mb) { o.mb[0] = mb_all; o.mb[1] = Mb; }
  if (o.inbreeding)
    for (int i = lane; i < n; i += 32) {
      const double mii = as[i + int64_t(i) * ld];
      o.inbreeding[i] = ((mii * 2 - 1) - mb_all) / (1 - mb_all);
    }
  if (o.fst) {
    for (int p = lane; p < npop; p += 32) o.fst[p] = (Fsts[p] - Mb) / (1 - Mb);
    __syncwarp();
    const double g = warp_mean_narm(o.fst, npop, lane);
    if (lane == 0) o.fst[npop] = g;
  }
  if (o.fis) {
    for (int p = 0; p < npop; ++p) {
      const int len = seg[p + 1] - seg[p];
      for (int r = lane; r < len; r += 32) {
        const int i = perm[seg[p] + r];
        const double mii = as[i + int64_t(i) * ld] * 2 - 1;
        Fi[r] = (mii - Fsts[p]) / (1 - Fsts[p]);
      }
      __syncwarp();
      const double f = warp_mean_narm(Fi, len, lane);
      if (lane == 0) o.fis[p] = f;
      __syncwarp();
    }
    const double g = warp_mean_narm(o.fis, npop, lane);
    if (lane == 0) o.fis[npop] = g;
  }
}

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes, const Err& E) {
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
      cudaGetLastError();
      p = nullptr;
      return E.fail(TPG_ERR_NOMEM, "cannot allocate %zu bytes on the device", bytes);
    }
    return TPG_OK;
  }
  template <class T> T* as() { return static_cast<T*>(p); }
};

int run_wg17(Ctx* ctx, const double* as_in, bool on_device, int64_t n, const int32_t* group, int64_t npop,
             double* inbreeding, double* fst, double* fis, double* mb_out, const Err& E) {
  if (!as_in || n <= 0) return E.fail(TPG_ERR_ARG, "allele_sharing_mat is NULL or empty");
  if (n > INT32_MAX / 2) return E.fail(TPG_ERR_ARG, "n too large");
  if (!group) npop = 1;
  if (npop <= 0 || npop > n) return E.fail(TPG_ERR_ARG, "bad number of populations %lld", (long long)npop);
  if ((fst || fis) && npop < 2) return E.fail(TPG_ERR_ARG, ".x should be a grouped gen_tibble");
  TPG_CUDA_OK(cudaSetDevice(ctx->device), E);
  cudaStream_t st = ctx->s_compute;
  // population-sorted permutation (stable: which(pop == z) order, R/pop_fst.R:44)
  std::vector<int32_t> perm(n), seg(npop + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t g = group ? group[i] : 0;
    if (g < 0 || g >= npop) return E.fail(TPG_ERR_ARG, "group[%lld] = %d is out of range [0, %lld)", (long long)i, g, (long long)npop);
    seg[g + 1]++;
  }
  for (int64_t g = 0; g < npop; ++g) {
    if (seg[g + 1] == 0) return E.fail(TPG_ERR_ARG, "population %lld has no individuals", (long long)g);
    seg[g + 1] += seg[g];
  }
  {
    std::vector<int32_t> fill(seg.begin(), seg.end() - 1);
    for (int64_t i = 0; i < n; ++i) perm[fill[group ? group[i] : 0]++] = int32_t(i);
  }
  DevBuf d_as, d_perm, d_seg, d_shi, d_slo, d_cnt, d_p, d_out;
  const double* as = as_in;
  int rc;
  if (!on_device) {
    if ((rc = d_as.alloc(sizeof(double) * size_t(n) * size_t(n), E))) return rc;
    TPG_CUDA_OK(cudaMemcpyAsync(d_as.p, as_in, sizeof(double) * size_t(n) * size_t(n), cudaMemcpyHostToDevice, st), E);
    as = d_as.as<double>();
  }
  if ((rc = d_perm.alloc(sizeof(int32_t) * n, E)) || (rc = d_seg.alloc(sizeof(int32_t) * (npop + 1), E)) ||
      (rc = d_shi.alloc(sizeof(double) * n * npop, E)) || (rc = d_slo.alloc(sizeof(double) * n * npop, E)) ||
      (rc = d_cnt.alloc(sizeof(int32_t) * n * npop, E)) || (rc = d_p.alloc(sizeof(double) * 3 * npop * npop, E)) ||
      (rc = d_out.alloc(sizeof(double) * (2 * n + 3 * (npop + 1) + 2), E)))
    return rc;
  TPG_CUDA_OK(cudaMemcpyAsync(d_perm.p, perm.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st), E);
  TPG_CUDA_OK(cudaMemcpyAsync(d_seg.p, seg.data(), sizeof(int32_t) * (npop + 1), cudaMemcpyHostToDevice, st), E);
  const int wpb = 8;
  group_colsums_kernel<<<unsigned((n + wpb - 1) / wpb), 32 * wpb, 0, st>>>(
      as, n, int(n), int(npop), d_perm.as<int32_t>(), d_seg.as<int32_t>(), d_shi.as<double>(), d_slo.as<double>(),
      d_cnt.as<int32_t>());
  double* p_hi = d_p.as<double>();
  double* p_lo = p_hi + npop * npop;
  double* p_cnt = p_lo + npop * npop;
  group_pairs_kernel<<<unsigned((npop * npop + wpb - 1) / wpb), 32 * wpb, 0, st>>>(
      int(n), int(npop), d_perm.as<int32_t>(), d_seg.as<int32_t>(), d_shi.as<double>(), d_slo.as<double>(),
      d_cnt.as<int32_t>(), p_hi, p_lo, p_cnt);
  double* o = d_out.as<double>();
  Wg17Out w;
  w.inbreeding = inbreeding ? o : nullptr;
  w.fst = fst ? o + n : nullptr;
  w.fis = fis ? o + n + (npop + 1) : nullptr;
  w.mb = o + n + 2 * (npop + 1);
  w.scratch = o + n + 2 * (npop + 1) + 2;  // npop + n doubles
  wg17_finish_kernel<<<1, 32, 0, st>>>(as, n, int(n), int(npop), d_perm.as<int32_t>(), d_seg.as<int32_t>(), p_hi, p_lo,
                                       p_cnt, w);
  TPG_CUDA_OK(cudaGetLastError(), E);
  ctx->launches += 3;
  if (inbreeding) TPG_CUDA_OK(cudaMemcpyAsync(inbreeding, w.inbreeding, sizeof(double) * n, cudaMemcpyDeviceToHost, st), E);
  if (fst) TPG_CUDA_OK(cudaMemcpyAsync(fst, w.fst, sizeof(double) * (npop + 1), cudaMemcpyDeviceToHost, st), E);
  if (fis) TPG_CUDA_OK(cudaMemcpyAsync(fis, w.fis, sizeof(double) * (npop + 1), cudaMemcpyDeviceToHost, st), E);
  if (mb_out) TPG_CUDA_OK(cudaMemcpyAsync(mb_out, w.mb, sizeof(double) * 2, cudaMemcpyDeviceToHost, st), E);
  TPG_CUDA_OK(cudaStreamSynchronize(st), E);
  return TPG_OK;
}

}  // namespace tpg
