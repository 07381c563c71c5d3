// Word-parallel (SWAR) genotype arithmetic of the 4-bit decode kernels: plain 32-bit integer
// operations that compile for the device and — with PRMT emulated — for the host, so that
// tests/test_host_logic.py can check every byte value and layout exhaustively without a GPU
// (tests/swar_check.cpp).  The rule restated here is the reference's raw-byte rule:
//   src/snp_king.cpp:45-58, src/snp_ibs.cpp:45-56   value == 0 / 1 / 2 -> hom-ref / het / hom-alt
//   src/snp_as.cpp:44-53                            value < 3 -> typed, anything else missing
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TPG_HD __host__ __device__ __forceinline__
#else
#define TPG_HD inline
#endif

namespace tpg {

// prmt.b32 in its generic form: selector nibble s picks byte (s & 7) of {x (0-3), y (4-7)};
// with bit 3 set the byte's sign bit is replicated over the whole result byte instead.
TPG_HD uint32_t prmt(uint32_t x, uint32_t y, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(y), "r"(sel));
  return d;
#else
  const uint64_t pool = (uint64_t(y) << 32) | x;
  uint32_t d = 0;
  for (int b = 0; b < 4; ++b) {
    const uint32_t s = (sel >> (4 * b)) & 0xFu;
    uint32_t v = uint32_t(pool >> (8 * (s & 7u))) & 0xFFu;
    if (s & 8u) v = (v & 0x80u) ? 0xFFu : 0x00u;
    d |= v << (8 * b);
  }
  return d;
#endif
}

// 4 raw genotype bytes -> 4 two-bit codes, one per byte: 0 / 1 / 2 = the dosage, 3 = missing
// (every byte value above 2, imputed codes and polyploid dosages included).
TPG_HD uint32_t clamp4(uint32_t g) {
  const uint32_t t = (g & 0x7F7F7F7Fu) + 0x7D7D7D7Du;  // bit 7 set iff (byte & 127) >= 3, no carries
  const uint32_t m = prmt(t | g, 0u, 0xBA98u);          // 0xFF where byte >= 3
  return (g | m) & 0x03030303u;
}

// Two words of clamped codes -> one word of 8 nibbles: `lo` in the low, `hi` in the high nibble
// of each byte.
TPG_HD uint32_t nibble_pair(uint32_t lo, uint32_t hi) { return hi * 16u + lo; }

// 4x4 byte transpose: w[k] = byte r holds item (row r, column k)  ->  t[r] = byte k holds it.
TPG_HD void transpose4x4(const uint32_t (&w)[4], uint32_t (&t)[4]) {
  const uint32_t a0 = prmt(w[0], w[1], 0x5140u);
  const uint32_t a1 = prmt(w[2], w[3], 0x5140u);
  const uint32_t b0 = prmt(w[0], w[1], 0x7362u);
  const uint32_t b1 = prmt(w[2], w[3], 0x7362u);
  t[0] = prmt(a0, a1, 0x5410u);
  t[1] = prmt(a0, a1, 0x7632u);
  t[2] = prmt(b0, b1, 0x5410u);
  t[3] = prmt(b0, b1, 0x7632u);
}

// 8 clamped codes (one per nibble) -> the E2M1 nibbles of the three planes:
//   cn (NM): 0x2 (+1) where typed;  ch (H1): 0x2 where het;
//   cd (d):  0x2 (+1) where hom-alt, 0xA (-1) where hom-ref.
TPG_HD void codes8(uint32_t q, uint32_t& cd, uint32_t& cn, uint32_t& ch) {
  const uint32_t s = q << 1;                       // bit 1 of a nibble: q's bit 0 (s), q's bit 1 (q)
  cn = ~(q & s) & 0x22222222u;                     // not 3
  ch = (s & ~q) & 0x22222222u;                     // 1
  const uint32_t c2 = (q & ~s) & 0x22222222u;      // 2
  const uint32_t c0 = ~(q | s) & 0x22222222u;      // 0
  cd = c0 * 5u + c2;                               // 0x2 * 5 = 0xA, nibbles never carry
}

}  // namespace tpg
