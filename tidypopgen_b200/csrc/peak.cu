// tpg_mma_peak: the tensor-pipe issue rate of THIS GPU for the two MMA kinds the Gram kernels use,
// measured the way the Gram kernels issue them — tcgen05.mma.cta_group::2, M = 256 over a CTA pair,
// N = 256, one elected thread, accumulating into one TMEM tile — but with the operands resident in
// shared memory: no TMA, no L2 traffic, no epilogue.  What is left is the rate at which the tensor
// cores retire MMAs, i.e. the roofline the Gram kernels are read against (bench.py `roofline.peak`).
// MEASURED_PEAKS.json only has a bf16 cuBLAS figure; 2 x / 4 x that is an inference, this is a measurement.
//
// Every resident CTA pair of the device runs the same loop (a full-chip, power-limited figure, not one
// SM's): groups of four MMAs (the four 32-byte K steps of a 128-byte swizzled row: K = 4 x 32 int8 or
// 4 x 64 E2M1), a commit every 16 groups, the issuer waiting two commits behind so the queue never drains.
#include "internal.h"
#include "ptx.cuh"

namespace tpg {

constexpr int kPeakThreads = 128;
constexpr int kPeakHalf = 128 * 128;                      // one 128-row x 128-byte operand box
constexpr int kPeakSmem = 2 * kPeakHalf + 1024 + 64;      // A half + B half per CTA

template <bool kMxf4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPeakThreads, 1)
mma_peak_kernel(int groups, int random_operands, int* err_flag) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * kPeakHalf);  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  // operands: pseudo-random -1 / 0 / +1 like the indicator planes (E2M1 nibbles 0xA / 0x0 / 0x2, int8 bytes
  // 0xFF / 0x00 / 0x01), or all zeros.  The values do not change the issue rate, but they do change the power
  // the tensor cores draw: a run of a second on zeros stays at the maximum clock, real data meets the power cap.
  for (int e = threadIdx.x; e < 2 * kPeakHalf / 4; e += kPeakThreads) {
    uint32_t w = 0;
    if (random_operands) {
      uint32_t h = (uint32_t(e) * 2654435761u) ^ (blockIdx.x * 0x9E3779B9u);
      for (int b = 0; b < (kMxf4 ? 8 : 4); ++b) {
        h = h * 1664525u + 1013904223u;
        const uint32_t t = (h >> 24) % 3u;  // 0, 1, 2 -> 0, +1, -1
        if (kMxf4) w |= (t == 0 ? 0x0u : t == 1 ? 0x2u : 0xAu) << (4 * b);
        else w |= (t == 0 ? 0x00u : t == 1 ? 0x01u : 0xFFu) << (8 * b);
      }
    }
    reinterpret_cast<uint32_t*>(smem)[e] = w;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (kMxf4) {  // UE8M0 1.0 in columns [256, 288) of every lane: the scale factors
    tmem_fill_32x32(tmem_base + (uint32_t(warp * 32) << 16) + 256u, 0x7F7F7F7Fu);
    tmem_st_wait();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem);
    const uint64_t da = make_smem_desc_sw128(sa, 16u, 1024u);
    const uint64_t db = make_smem_desc_sw128(sa + kPeakHalf, 16u, 1024u);
    const uint32_t idesc = kMxf4 ? make_idesc_mxf4(256, 256) : make_idesc_i8(256, 256, false);
    const uint32_t sfa = tmem_base + 256u, sfb = tmem_base + 260u;
    uint32_t phase[2] = {0, 0};
    int commits = 0;
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kMxf4) umma_mxf4_pair(tmem_base, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (g | k) ? 1u : 0u, sfa, sfb);
        else umma_i8_pair(tmem_base, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (g | k) ? 1u : 0u);
      }
      if ((g & 15) == 15 || g + 1 == groups) {
        const int b = commits & 1;
        if (commits >= 2) { mbar_wait(&bar[b], phase[b], err_flag, 7); phase[b] ^= 1; }
        umma_commit_pair(&bar[b], 1);
        ++commits;
      }
    }
    for (int c = commits >= 2 ? commits - 2 : 0; c < commits; ++c) {  // drain the last commits
      const int b = c & 1;
      mbar_wait(&bar[b], phase[b], err_flag, 8);
      phase[b] ^= 1;
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// Runs the probe for about `seconds` (one calibration launch first); *tops = 2 * MACs / s / 1e12.
int run_mma_peak(Ctx* ctx, int kind_in, double seconds, double* tops, double* ms_out, const Err& E) {
  const int kind = kind_in & 1, random_ops = (kind_in & 2) ? 0 : 1;   // bit 1: all-zero operands
  if (kind_in < 0 || kind_in > 3) return E.fail(TPG_ERR_ARG, "mma_peak: kind must be 0 (int8) or 1 (4-bit mxf4), + 2 for zero operands");
  if (!(seconds > 0) || seconds > 5.0) return E.fail(TPG_ERR_ARG, "mma_peak: seconds must be in (0, 5]");
  const int pairs = ctx->num_sms / 2;
  cudaStream_t st = ctx->s_compute;
  auto launch = [&](int groups) -> cudaError_t {
    if (kind == 1) {
      cudaFuncSetAttribute(mma_peak_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeakSmem);
      mma_peak_kernel<true><<<2 * pairs, kPeakThreads, kPeakSmem, st>>>(groups, random_ops, ctx->d_err_flag);
    } else {
      cudaFuncSetAttribute(mma_peak_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeakSmem);
      mma_peak_kernel<false><<<2 * pairs, kPeakThreads, kPeakSmem, st>>>(groups, random_ops, ctx->d_err_flag);
    }
    return cudaGetLastError();
  };
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = TPG_OK;
  float ms = 0;
  int groups = 4096;
  for (int pass = 0; pass < 2 && rc == TPG_OK; ++pass) {
    cudaEventRecord(e0, st);
    cudaError_t e = launch(groups);
    cudaEventRecord(e1, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      Err quiet{nullptr, 0};
      rc = check_err_flag(ctx, quiet) == TPG_ERR_KERNEL ? check_err_flag(ctx, E)
                                                       : E.fail(TPG_ERR_CUDA, "mma_peak failed: %s", cudaGetErrorString(e));
      break;
    }
    ctx->launches++;
    cudaEventElapsedTime(&ms, e0, e1);
    if (pass == 0) {  // scale to the requested duration
      const double per_group_ms = double(ms) / groups;
      groups = int(std::min(2.0e9, std::max(4096.0, seconds * 1e3 / std::max(per_group_ms, 1e-7))));
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (rc) return rc;
  const double k_per_group = kind == 1 ? 256.0 : 128.0;  // loci (K) per group of four MMAs
  const double macs = double(pairs) * double(groups) * 256.0 * 256.0 * k_per_group;
  if (tops) *tops = 2.0 * macs / (double(ms) * 1e-3) / 1e12;
  if (ms_out) *ms_out = ms;
  return TPG_OK;
}

}  // namespace tpg
