// Known-answer experiment for DESIGN.md "what comes next" item 1: does the block-scaled 4-bit MMA
// (tcgen05.mma kind::mxf4, E2M1 operands, UE8M0 scales = 1) keep INTEGER sums exact in its FP32
// accumulator, and how fast is it next to kind::i8?  Not part of the library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I tidypopgen_b200/csrc scripts/mxf4_probe.cu -o gpurun_out/mxf4_probe
//
// One CTA.  A (128 x 256) and B (N x 256) tiles of E2M1 values in {-1, 0, +1}, K-major SWIZZLE_128B;
// every 16-byte chunk of a row holds the same 32-element pattern, which makes the tile invariant
// under the swizzle (so it can be written with plain stores).  One MMA (K = 64) adds
// 2 * dot(u_i, v_j) to D[i][j]; after T MMAs D must equal 2 T dot(u_i, v_j) EXACTLY while that is
// below 2^24.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"

using namespace tpg;

__device__ __forceinline__ void umma_mxf4(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc,
                                          uint32_t sfa, uint32_t sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, uint32_t v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kN = 256;

// mode 0: mxf4 exactness / timing; mode 1: i8 timing (same tile bytes interpreted as int8)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ a_rows, const uint8_t* __restrict__ b_rows, int T, int mode, float* d_out,
             long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                // 128 rows x 128 B
  uint8_t* sb = smem + 16384;        // 256 rows x 128 B (two 128-row blocks for i8 MN... K-major: 8-row groups)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // rows: 16-byte pattern per row, replicated over the 8 chunks of the 128-byte row
  for (int e = tid; e < 128 * 8; e += 128) {
    const int r = e >> 3, c = e & 7;
    reinterpret_cast<uint4*>(sa + r * 128)[c] = reinterpret_cast<const uint4*>(a_rows)[r];
  }
  for (int e = tid; e < kN * 8; e += 128) {
    const int r = e >> 3, c = e & 7;
    reinterpret_cast<uint4*>(sb + r * 128)[c] = reinterpret_cast<const uint4*>(b_rows)[r];
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  // scale factors = 1.0 (UE8M0 0x7F) everywhere in columns [256, 384) of all 128 lanes
  for (int c = 256; c < 384; c += 32) tmem_st_32x32(tb + (uint32_t(warp * 32) << 16) + uint32_t(c), 0x7F7F7F7Fu);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    // K-major SW128: 8-row groups 1024 B apart; one MMA consumes 32 B of a row (64 E2M1 / 32 int8)
    const uint64_t da = make_smem_desc_sw128(smem_u32(sa), 16u, 1024u);
    const uint64_t db = make_smem_desc_sw128(smem_u32(sb), 16u, 1024u);
    // block-scaled descriptor: a/b format E2M1 (1) at [7,10) / [10,13), K-major, N>>3 at [17,23),
    // scale format UE8M0 (1) at [23], M>>4 at [24,29)
    const uint32_t idesc_mx = (1u << 7) | (1u << 10) | (uint32_t(kN >> 3) << 17) | (1u << 23) | (uint32_t(128 >> 4) << 24);
    const uint32_t idesc_i8 = make_idesc_i8(128, kN, false);
    t0 = clock64();
    for (int t = 0; t < T; ++t) {
      const uint64_t off = uint64_t((t & 3) * 2);
      if (mode == 0) umma_mxf4(tb, da + off, db + off, idesc_mx, t > 0 ? 1u : 0u, tb + 256, tb + 320);
      else umma_i8(tb, da + off, db + off, idesc_i8, t > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0, nullptr, 1);
  if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
  tc_fence_after();
  for (int c = 0; c < kN; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tb + (uint32_t(warp * 32) << 16) + uint32_t(c), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d_out[(warp * 32 + (tid & 31)) * kN + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

static int e2m1(int v) { return v == 0 ? 0x0 : v > 0 ? 0x2 : 0xA; }  // 0, +1.0, -1.0

int main() {
  const int rowsA = 128, rowsB = kN;
  std::vector<int> ua(rowsA * 32), vb(rowsB * 32);
  std::vector<uint8_t> ha(rowsA * 16), hb(rowsB * 16);
  srand(7);
  for (int r = 0; r < rowsA; ++r)
    for (int k = 0; k < 32; ++k) ua[r * 32 + k] = r == 0 ? 1 : (rand() % 3) - 1;        // row 0: all +1
  for (int r = 0; r < rowsB; ++r)
    for (int k = 0; k < 32; ++k) vb[r * 32 + k] = r == 0 ? 1 : r == 1 ? -1 : (rand() % 3) - 1;
  for (int r = 0; r < rowsA; ++r)
    for (int k = 0; k < 16; ++k) ha[r * 16 + k] = uint8_t(e2m1(ua[r * 32 + 2 * k]) | (e2m1(ua[r * 32 + 2 * k + 1]) << 4));
  for (int r = 0; r < rowsB; ++r)
    for (int k = 0; k < 16; ++k) hb[r * 16 + k] = uint8_t(e2m1(vb[r * 32 + 2 * k]) | (e2m1(vb[r * 32 + 2 * k + 1]) << 4));
  uint8_t *da, *db;
  float* dd;
  long long* dc;
  cudaMalloc(&da, ha.size()); cudaMalloc(&db, hb.size()); cudaMalloc(&dd, sizeof(float) * rowsA * kN); cudaMalloc(&dc, 8);
  cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice);
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hd(rowsA * kN);
  printf("{\"probe\": \"tcgen05.mma kind::mxf4 (E2M1, UE8M0 scales = 1), M128 N%d K64, integer known-answer\", \"runs\": [\n", kN);
  const int Ts[] = {1, 16, 1024, 8192, 16384, 65536, 131072, 262144, 524288};
  for (int ti = 0; ti < 9; ++ti) {
    const int T = Ts[ti];
    probe_kernel<<<1, 128, smem>>>(da, db, T, 0, dd, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hd.data(), dd, sizeof(float) * hd.size(), cudaMemcpyDeviceToHost);
    long long cyc = 0;
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    long long bad = 0, maxabs = 0;
    double maxerr = 0;
    for (int i = 0; i < rowsA; ++i)
      for (int j = 0; j < rowsB; ++j) {
        long long dot = 0;
        for (int k = 0; k < 32; ++k) dot += ua[i * 32 + k] * vb[j * 32 + k];
        const long long want = 2LL * T * dot;
        const double got = hd[i * kN + j];
        if (llabs(want) < (1LL << 24)) {
          if (got != double(want)) { ++bad; if (fabs(got - want) > maxerr) maxerr = fabs(got - want); }
          if (llabs(want) > maxabs) maxabs = llabs(want);
        }
      }
    printf(" {\"T\": %d, \"loci\": %lld, \"D[0][0]\": %.1f, \"want[0][0]\": %lld, \"entries_below_2^24_wrong\": %lld, \"max_abs_err\": %.1f, "
           "\"largest_checked\": %lld, \"cycles\": %lld, \"cycles_per_mma\": %.1f}%s\n",
           T, 64LL * T, hd[0], 2LL * T * 32, bad, maxerr, maxabs, cyc, double(cyc) / T, ti == 8 ? "" : ",");
  }
  // timing: mxf4 (K = 64) against i8 (K = 32) on one SM, same N
  long long c_mx = 0, c_i8 = 0;
  probe_kernel<<<1, 128, smem>>>(da, db, 262144, 0, dd, dc); cudaDeviceSynchronize(); cudaMemcpy(&c_mx, dc, 8, cudaMemcpyDeviceToHost);
  probe_kernel<<<1, 128, smem>>>(da, db, 262144, 1, dd, dc); cudaDeviceSynchronize(); cudaMemcpy(&c_i8, dc, 8, cudaMemcpyDeviceToHost);
  printf("], \"timing_T262144\": {\"mxf4_cycles_per_mma\": %.1f, \"i8_cycles_per_mma\": %.1f, \"mxf4_MAC_per_cycle\": %.0f, \"i8_MAC_per_cycle\": %.0f}}\n",
         c_mx / 262144.0, c_i8 / 262144.0, 128.0 * kN * 64 / (c_mx / 262144.0), 128.0 * kN * 32 / (c_i8 / 262144.0));
  return 0;
}
