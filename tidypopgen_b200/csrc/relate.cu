// filter_high_relatedness on the device — the main consumer of the KING matrix
// (R/filter_high_relatedness.R:26-145, called by qc_report_indiv, R/qc_report_indiv.R:85-99):
// a caret::findCorrelation-style greedy filter that, for every pair above the threshold,
// drops the member with the higher mean relatedness.  The reference recomputes two means
// over the whole n x n matrix for every such pair (O(n^2) each, in R); here
//   1. col_stats      column means of |M| (diagonal and NaN out)  -> the visiting order
//   2. row_stats      per-row double-double sums / counts of |M|
//   3. count / fill   the ordered list of "events": pairs (a < b in visiting order) whose
//                     value is above the threshold, or NaN (where R stops with an error)
//   4. walk           ONE persistent CTA replays the reference's double loop over the events;
//                     a decision is O(1) from the running sums, a deletion updates the row
//                     sums from one matrix column (O(n), all threads) and re-reduces the totals.
// The matrix is read ~3 times in total and never leaves HBM.
//
// Means are accumulated in double-double and rounded once (ddsum.cuh); the reference's are
// R's long-double two-pass mean().  Both are the exact mean to within an ulp of their
// working precision, so the comparisons mn1 > mn2 and the ordering agree unless two means
// coincide to ~1e-19 relative — exact ties (identical value multisets, dyadic toy data)
// are reproduced because both sides are then exact.
#include <algorithm>
#include <cmath>
#include <vector>
#include "ddsum.cuh"
#include "internal.h"

namespace tpg {

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

// ---- 1. column means of |M| with the diagonal and NaNs left out (:56-72)
__global__ void __launch_bounds__(256)
relate_col_stats_kernel(const double* __restrict__ m, int64_t ld, int n, double* __restrict__ colmean) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const double* col = m + int64_t(j) * ld;
  dd acc = dd_zero();
  long long c = 0, ninf = 0;
  for (int i = lane; i < n; i += 32) {
    if (i == j) continue;
    const double v = fabs(col[i]);
    if (v != v) continue;
    ++c;
    if (isinf(v)) ++ninf; else acc = dd_add_d(acc, v);
  }
  acc = dd_warp_sum(acc);
  c = i64_warp_sum(c);
  ninf = i64_warp_sum(ninf);
  if (lane == 0) colmean[j] = (ninf && c) ? double(INFINITY) : dd_mean(acc, double(c));
}

// ---- 2. per-row sums of |M| (row i = matrix2[i, ]): 32 rows per CTA, coalesced along rows
__global__ void __launch_bounds__(256)
relate_row_stats_kernel(const double* __restrict__ m, int64_t ld, int n, double* __restrict__ rs_hi,
                        double* __restrict__ rs_lo, int32_t* __restrict__ rcnt, int32_t* __restrict__ rinf) {
  __shared__ double sh_hi[8][33], sh_lo[8][33];
  __shared__ int sh_c[8][33], sh_i[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  dd acc = dd_zero();
  int c = 0, ninf = 0;
  if (i < n)
    for (int j = ty; j < n; j += 8) {
      if (j == i) continue;
      const double v = fabs(m[i + int64_t(j) * ld]);
      if (v != v) continue;
      ++c;
      if (isinf(v)) ++ninf; else acc = dd_add_d(acc, v);
    }
  sh_hi[ty][tx] = acc.hi; sh_lo[ty][tx] = acc.lo; sh_c[ty][tx] = c; sh_i[ty][tx] = ninf;
  __syncthreads();
  if (ty == 0 && i < n) {
    dd t = dd{sh_hi[0][tx], sh_lo[0][tx]};
    int cc = sh_c[0][tx], ii = sh_i[0][tx];
    for (int k = 1; k < 8; ++k) { t = dd_add(t, dd{sh_hi[k][tx], sh_lo[k][tx]}); cc += sh_c[k][tx]; ii += sh_i[k][tx]; }
    rs_hi[i] = t.hi; rs_lo[i] = t.lo; rcnt[i] = cc; rinf[i] = ii;
  }
}

// ---- 3. events of visiting rank a: b > a with |M[P[a], P[b]]| > thr or NaN, in ascending b.
// pass 0 counts (and counts every ordered pair above the threshold for the `any` test, :90),
// pass 1 fills ev[eoff[a] ..) with b | (isnan << 31).
template <bool kFill>
__global__ void __launch_bounds__(256)
relate_events_kernel(const double* __restrict__ m, int64_t ld, int n, double thr, const int32_t* __restrict__ P,
                     int32_t* __restrict__ ecount, int32_t* __restrict__ gt_all, const int64_t* __restrict__ eoff,
                     uint32_t* __restrict__ ev) {
  const int lane = threadIdx.x & 31;
  const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= n) return;
  const int64_t ia = P[a];
  int cnt = 0, gt = 0;
  int64_t pos = kFill ? eoff[a] : 0;
  for (int b0 = kFill ? (a + 1) & ~31 : 0; b0 < n; b0 += 32) {
    const int b = b0 + lane;
    bool is_ev = false, is_nan = false;
    if (b < n && b != a) {
      const double v = fabs(m[ia + int64_t(P[b]) * ld]);
      is_nan = v != v;
      const bool above = !is_nan && v > thr;
      if (above) ++gt;
      is_ev = b > a && (above || is_nan);
    }
    if (kFill) {
      const uint32_t mask = __ballot_sync(0xFFFFFFFFu, is_ev);
      if (is_ev) ev[pos + __popc(mask & ((1u << lane) - 1u))] = uint32_t(b) | (is_nan ? 0x80000000u : 0u);
      pos += __popc(mask);
    } else if (is_ev) {
      ++cnt;
    }
  }
  if (!kFill) {
    cnt = int(i64_warp_sum(cnt));
    gt = int(i64_warp_sum(gt));
    if (lane == 0) { ecount[a] = cnt; gt_all[a] = gt; }
  }
}

// ---- 4. the greedy walk (:89-135), one CTA
constexpr int kWalkThreads = 1024;

struct WalkParams {
  const double* m;
  int64_t ld;
  int n;
  double thr;
  const int32_t* P;        // rank -> original index
  const int32_t* erows;    // ranks that have events, ascending
  int n_erows;
  const int64_t* eoff;     // per rank: start of its events (n + 1 entries)
  const uint32_t* ev;
  double* rs_hi;           // row state, by rank
  double* rs_lo;
  int32_t* rcnt;
  int32_t* rinf;
  int32_t* rgt;            // living partners above the threshold in this row
  uint8_t* dead;           // by rank (zeroed by the host)
  int32_t* status;         // [0] = 0 ok / 1 NaN reached, [1] = number deleted
};

// Block-wide sums in a fixed order (identical in every thread): a double-double and three counts.
__device__ void block_sums(dd* v, long long* c, long long* i, long long* g, double* sh_hi, double* sh_lo,
                           long long* sh_c, long long* sh_i, long long* sh_g) {
  dd acc = dd_warp_sum(*v);
  const long long cc = i64_warp_sum(*c), ii = i64_warp_sum(*i), gg = i64_warp_sum(*g);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh_hi[w] = acc.hi; sh_lo[w] = acc.lo; sh_c[w] = cc; sh_i[w] = ii; sh_g[w] = gg; }
  __syncthreads();
  dd t = dd_zero();
  long long tc = 0, ti = 0, tg = 0;
  for (int k = 0; k < kWalkThreads / 32; ++k) {
    t = dd_add(t, dd{sh_hi[k], sh_lo[k]});
    tc += sh_c[k]; ti += sh_i[k]; tg += sh_g[k];
  }
  *v = t; *c = tc; *i = ti; *g = tg;
  __syncthreads();
}

// Row state is indexed by RANK here (the host permutes it once): rs = sum of the row's living
// entries, rcnt = how many (Inf included), rinf = infinities, rgt = entries above the threshold.
__global__ void __launch_bounds__(kWalkThreads, 1)
relate_walk_kernel(const WalkParams p) {
  __shared__ double sh_hi[32], sh_lo[32];
  __shared__ long long sh_c[32], sh_i[32], sh_g[32];
  __shared__ uint32_t sh_ev[kWalkThreads];
  __shared__ uint32_t sh_mask[kWalkThreads / 32];
  // totals over the living rows: S, C (count), I (infinities), live (ordered pairs above the
  // threshold: the `any` test of :90) — summed once, then kept up to date per deletion
  dd S = dd_zero();
  long long C = 0, I = 0, live = 0;
  for (int r = threadIdx.x; r < p.n; r += kWalkThreads) {
    S = dd_add(S, dd{p.rs_hi[r], p.rs_lo[r]});
    C += p.rcnt[r];
    I += p.rinf[r];
    live += p.rgt[r];
  }
  block_sums(&S, &C, &I, &live, sh_hi, sh_lo, sh_c, sh_i, sh_g);
  int deleted = 0;
  for (int e = 0; e < p.n_erows; ++e) {
    if (live == 0) break;                       // :90 — nothing above the threshold is left
    const int a = p.erows[e];
    if (p.dead[a]) continue;                    // :96
    bool a_dead = false;
    for (int64_t base = p.eoff[a]; base < p.eoff[a + 1] && !a_dead; base += kWalkThreads) {
      const int64_t k = base + threadIdx.x;
      uint32_t w = 0;
      bool alive = false;
      if (k < p.eoff[a + 1]) { w = p.ev[k]; alive = !p.dead[w & 0x7FFFFFFFu]; }
      sh_ev[threadIdx.x] = w;
      // Partners alive NOW stay alive while this row is walked, except the one a decision removes
      // (each partner occurs once per row), so the snapshot can be consumed bit by bit.
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, alive);
      if ((threadIdx.x & 31) == 0) sh_mask[threadIdx.x >> 5] = m;
      __syncthreads();
      for (int wq = 0; wq < kWalkThreads / 32 && !a_dead; ++wq) {
        uint32_t bits = sh_mask[wq];
        while (bits && !a_dead) {
          const int q = wq * 32 + __ffs(bits) - 1;
          bits &= bits - 1;
          const uint32_t evq = sh_ev[q];
          if (evq & 0x80000000u) {              // if (NA > threshold): R stops here
            if (threadIdx.x == 0) { p.status[0] = 1; p.status[1] = deleted; }
            return;
          }
          const int b = int(evq);
          // mn1 = mean(matrix2[a, ]), mn2 = mean(matrix2[-b, ]), na.rm = TRUE (:102-103)
          const double mn1 = (p.rinf[a] && p.rcnt[a]) ? double(INFINITY)
                                                      : dd_mean(dd{p.rs_hi[a], p.rs_lo[a]}, double(p.rcnt[a]));
          const long long c2 = C - p.rcnt[b], i2 = I - p.rinf[b];
          const double mn2 = (i2 && c2) ? double(INFINITY)
                                        : dd_mean(dd_add(S, dd_neg(dd{p.rs_hi[b], p.rs_lo[b]})), double(c2));
          const int kill = (mn1 > mn2) ? a : b; // :117-131
          const int64_t colk = int64_t(p.P[kill]) * p.ld;
          // the victim's own row leaves the totals ...
          dd rem = dd{p.rs_hi[kill], p.rs_lo[kill]};
          long long remc = p.rcnt[kill], remi = p.rinf[kill], remg = p.rgt[kill];
          __syncthreads();                      // everyone has read the sums of the victim
          // ... and so does its column: matrix2[r, kill] <- NA for every living r
          dd part = dd_zero();
          long long pc = 0, pi = 0, pg = 0;
          for (int r = threadIdx.x; r < p.n; r += kWalkThreads) {
            if (r == kill || p.dead[r]) continue;
            const double v = fabs(p.m[p.P[r] + colk]);
            if (v != v) continue;
            p.rcnt[r] -= 1; ++pc;
            if (isinf(v)) { p.rinf[r] -= 1; ++pi; }
            else {
              const dd t = dd_add_d(dd{p.rs_hi[r], p.rs_lo[r]}, -v);
              p.rs_hi[r] = t.hi; p.rs_lo[r] = t.lo;
              part = dd_add_d(part, v);
            }
            if (v > p.thr) { p.rgt[r] -= 1; ++pg; }
          }
          if (threadIdx.x == 0) p.dead[kill] = 1;                  // matrix2[kill, ] <- NA
          block_sums(&part, &pc, &pi, &pg, sh_hi, sh_lo, sh_c, sh_i, sh_g);
          S = dd_add(S, dd_neg(dd_add(rem, part)));
          C -= remc + pc; I -= remi + pi; live -= remg + pg;
          ++deleted;
          if (kill == a) a_dead = true;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) { p.status[0] = 0; p.status[1] = deleted; }
}

int run_filter_high_relatedness(Ctx* ctx, const double* mat, bool on_device, int64_t n, double thr, int32_t* keep,
                                int32_t* order, int64_t* n_removed, const Err& E) {
  if (!mat || n <= 0 || !keep) return E.fail(TPG_ERR_ARG, "matrix / keep is NULL or empty");
  if (n > (int64_t(1) << 30)) return E.fail(TPG_ERR_ARG, "n too large");
  if (thr != thr) return E.fail(TPG_ERR_ARG, "kings_threshold is NA");
  if (n_removed) *n_removed = 0;
  if (n == 1) {  // :46-53
    keep[0] = 1;
    if (order) order[0] = 0;
    return TPG_OK;
  }
  TPG_CUDA_OK(cudaSetDevice(ctx->device), E);
  cudaStream_t st = ctx->s_compute;
  DevBuf d_m, d_cm, d_rs, d_ri, d_rgt, d_P, d_ec, d_eoff, d_ev, d_erows, d_dead, d_status;
  const double* m = mat;
  int rc;
  const size_t nn = size_t(n) * size_t(n);
  if (!on_device) {
    if ((rc = d_m.alloc(sizeof(double) * nn, E))) return rc;
    TPG_CUDA_OK(cudaMemcpyAsync(d_m.p, mat, sizeof(double) * nn, cudaMemcpyHostToDevice, st), E);
    m = d_m.as<double>();
  }
  if ((rc = d_cm.alloc(sizeof(double) * n, E)) || (rc = d_rs.alloc(sizeof(double) * 2 * n, E)) ||
      (rc = d_ri.alloc(sizeof(int32_t) * 2 * n, E)) || (rc = d_P.alloc(sizeof(int32_t) * n, E)) ||
      (rc = d_ec.alloc(sizeof(int32_t) * 2 * n, E)) || (rc = d_eoff.alloc(sizeof(int64_t) * (n + 1), E)) ||
      (rc = d_dead.alloc(size_t(n), E)) || (rc = d_status.alloc(sizeof(int32_t) * 2, E)))
    return rc;
  const int wpb = 8;
  const unsigned wblocks = unsigned((n + wpb - 1) / wpb);
  relate_col_stats_kernel<<<wblocks, 32 * wpb, 0, st>>>(m, n, int(n), d_cm.as<double>());
  relate_row_stats_kernel<<<unsigned((n + 31) / 32), 256, 0, st>>>(m, n, int(n), d_rs.as<double>(), d_rs.as<double>() + n,
                                                                    d_ri.as<int32_t>(), d_ri.as<int32_t>() + n);
  std::vector<double> cm(n);
  TPG_CUDA_OK(cudaMemcpyAsync(cm.data(), d_cm.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st), E);
  TPG_CUDA_OK(cudaStreamSynchronize(st), E);
  // order(colmeans, decreasing = TRUE): stable, NaN last (:71-72)
  std::vector<int32_t> P(n);
  for (int64_t i = 0; i < n; ++i) P[i] = int32_t(i);
  std::stable_sort(P.begin(), P.end(), [&](int32_t x, int32_t y) {
    const double a = cm[x], b = cm[y];
    if (a != a) return false;
    if (b != b) return true;
    return a > b;
  });
  TPG_CUDA_OK(cudaMemcpyAsync(d_P.p, P.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st), E);
  relate_events_kernel<false><<<wblocks, 32 * wpb, 0, st>>>(m, n, int(n), thr, d_P.as<int32_t>(), d_ec.as<int32_t>(),
                                                           d_ec.as<int32_t>() + n, nullptr, nullptr);
  std::vector<int32_t> ec(2 * n);
  TPG_CUDA_OK(cudaMemcpyAsync(ec.data(), d_ec.p, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, st), E);
  TPG_CUDA_OK(cudaStreamSynchronize(st), E);
  std::vector<int64_t> eoff(n + 1, 0);
  std::vector<int32_t> erows;
  long long live = 0;
  for (int64_t a = 0; a < n; ++a) {
    eoff[a + 1] = eoff[a] + ec[a];
    if (ec[a]) erows.push_back(int32_t(a));
    live += ec[n + a];
  }
  ctx->launches += 3;
  std::vector<uint8_t> dead(n, 0);
  int32_t status[2] = {0, 0};
  if (!erows.empty() && live > 0) {
    if ((rc = d_ev.alloc(sizeof(uint32_t) * size_t(eoff[n]), E)) || (rc = d_erows.alloc(sizeof(int32_t) * erows.size(), E)) ||
        (rc = d_rgt.alloc(sizeof(int32_t) * n, E)))
      return rc;
    // row state in visiting (rank) order
    std::vector<double> rs(2 * n), rs_r(2 * n);
    std::vector<int32_t> ri(2 * n), ri_r(2 * n), rgt(n);
    TPG_CUDA_OK(cudaMemcpyAsync(rs.data(), d_rs.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, st), E);
    TPG_CUDA_OK(cudaMemcpyAsync(ri.data(), d_ri.p, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost, st), E);
    TPG_CUDA_OK(cudaStreamSynchronize(st), E);
    for (int64_t a = 0; a < n; ++a) {
      rs_r[a] = rs[P[a]]; rs_r[n + a] = rs[n + P[a]];
      ri_r[a] = ri[P[a]]; ri_r[n + a] = ri[n + P[a]];
      rgt[a] = ec[n + a];
    }
    TPG_CUDA_OK(cudaMemcpyAsync(d_rs.p, rs_r.data(), sizeof(double) * 2 * n, cudaMemcpyHostToDevice, st), E);
    TPG_CUDA_OK(cudaMemcpyAsync(d_ri.p, ri_r.data(), sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, st), E);
    TPG_CUDA_OK(cudaMemcpyAsync(d_rgt.p, rgt.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st), E);
    TPG_CUDA_OK(cudaMemcpyAsync(d_eoff.p, eoff.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, st), E);
    TPG_CUDA_OK(cudaMemcpyAsync(d_erows.p, erows.data(), sizeof(int32_t) * erows.size(), cudaMemcpyHostToDevice, st), E);
    TPG_CUDA_OK(cudaMemsetAsync(d_dead.p, 0, size_t(n), st), E);
    TPG_CUDA_OK(cudaStreamSynchronize(st), E);  // the host vectors above go out of scope
    relate_events_kernel<true><<<wblocks, 32 * wpb, 0, st>>>(m, n, int(n), thr, d_P.as<int32_t>(), nullptr, nullptr,
                                                            d_eoff.as<int64_t>(), d_ev.as<uint32_t>());
    WalkParams w;
    w.m = m; w.ld = n; w.n = int(n); w.thr = thr; w.P = d_P.as<int32_t>();
    w.erows = d_erows.as<int32_t>(); w.n_erows = int(erows.size());
    w.eoff = d_eoff.as<int64_t>(); w.ev = d_ev.as<uint32_t>();
    w.rs_hi = d_rs.as<double>(); w.rs_lo = d_rs.as<double>() + n;
    w.rcnt = d_ri.as<int32_t>(); w.rinf = d_ri.as<int32_t>() + n;
    w.rgt = d_rgt.as<int32_t>();
    w.dead = d_dead.as<uint8_t>(); w.status = d_status.as<int32_t>();
    relate_walk_kernel<<<1, kWalkThreads, 0, st>>>(w);
    TPG_CUDA_OK(cudaGetLastError(), E);
    ctx->launches += 2;
    TPG_CUDA_OK(cudaMemcpyAsync(dead.data(), d_dead.p, size_t(n), cudaMemcpyDeviceToHost, st), E);
    TPG_CUDA_OK(cudaMemcpyAsync(status, d_status.p, sizeof status, cudaMemcpyDeviceToHost, st), E);
    TPG_CUDA_OK(cudaStreamSynchronize(st), E);
    if (status[0] == 1)  // what R reports when the comparison at :101 meets a NaN
      return E.fail(TPG_ERR_ARG, "missing value where TRUE/FALSE needed (the relatedness matrix holds NaN for a "
                                 "pair the filter has to compare)");
  }
  for (int64_t r = 0; r < n; ++r) {
    keep[P[r]] = dead[r] ? 0 : 1;
    if (order) order[r] = P[r];
  }
  if (n_removed) *n_removed = status[1];
  return TPG_OK;
}

}  // namespace tpg
