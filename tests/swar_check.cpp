// Host-side check of tidypopgen_b200/csrc/swar.cuh (the word-parallel arithmetic of
// decode_planes_k4_kernel), built and run by tests/test_host_logic.py with g++ — no GPU.
// The expectation is the reference's raw-byte rule, one genotype at a time
// (src/snp_king.cpp:45-58, src/snp_as.cpp:44-53): byte 0 / 1 / 2 = hom-ref / het / hom-alt,
// anything else missing; planes d = [2] - [0], NM = [<= 2], H1 = [1]; E2M1: +1 = 0x2, -1 = 0xA.
#include <cstdio>
#include <cstdlib>
#include "../tidypopgen_b200/csrc/swar.cuh"

static uint32_t rng_state = 0x2545F491u;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 17; rng_state ^= rng_state << 5; return rng_state; }

static void expect(uint8_t b, unsigned& d, unsigned& nm, unsigned& h1) {
  nm = b <= 2 ? 0x2u : 0u;
  h1 = b == 1 ? 0x2u : 0u;
  d = b == 2 ? 0x2u : (b == 0 ? 0xAu : 0u);
}

int main() {
  long checks = 0;
  // clamp4: every byte value in every byte position, neighbours arbitrary (no carries between bytes)
  for (int pos = 0; pos < 4; ++pos)
    for (unsigned b = 0; b < 256; ++b)
      for (int rep = 0; rep < 64; ++rep) {
        uint32_t g = rnd();
        if (rep == 0) g = 0; else if (rep == 1) g = 0xFFFFFFFFu; else if (rep == 2) g = 0x02020202u; else if (rep == 3) g = 0x03030303u;
        g = (g & ~(0xFFu << (8 * pos))) | (b << (8 * pos));
        const uint32_t q = tpg::clamp4(g);
        for (int k = 0; k < 4; ++k) {
          const unsigned by = (g >> (8 * k)) & 0xFFu, want = by <= 2 ? by : 3u;
          if (((q >> (8 * k)) & 0xFFu) != want) { printf("clamp4 fails: g=%08x q=%08x byte %d\n", g, q, k); return 1; }
          ++checks;
        }
      }
  // the whole lane pipeline: 16 loci x 4 individuals -> 2 words per individual and plane
  for (int it = 0; it < 200000; ++it) {
    uint32_t g[16];
    for (int u = 0; u < 16; ++u) {
      uint32_t w = 0;
      for (int r = 0; r < 4; ++r) {
        const uint32_t x = rnd();
        uint32_t b = (x & 0xF) < 11 ? (x >> 8) % 3 : ((x & 0xF) < 14 ? 3 + (x >> 8) % 5 : (x >> 8) & 0xFF);  // mostly dosages, some NA / imputed / junk
        w |= b << (8 * r);
      }
      g[u] = w;
    }
    for (int w = 0; w < 2; ++w) {
      uint32_t pr[4], t[4];
      for (int k = 0; k < 4; ++k) pr[k] = tpg::nibble_pair(tpg::clamp4(g[8 * w + k]), tpg::clamp4(g[8 * w + 4 + k]));
      tpg::transpose4x4(pr, t);
      for (int r = 0; r < 4; ++r) {
        uint32_t cd, cn, ch;
        tpg::codes8(t[r], cd, cn, ch);
        for (int k = 0; k < 8; ++k) {  // locus 8w + k of individual r: byte k & 3, low nibble for k < 4
          const uint8_t b = (g[8 * w + k] >> (8 * r)) & 0xFFu;
          unsigned d, nm, h1;
          expect(b, d, nm, h1);
          const int sh = 8 * (k & 3) + (k >= 4 ? 4 : 0);
          if (((cd >> sh) & 0xFu) != d || ((cn >> sh) & 0xFu) != nm || ((ch >> sh) & 0xFu) != h1) {
            printf("codes fail: byte %u -> d %x nm %x h1 %x (want %x %x %x)\n", b, (cd >> sh) & 0xFu, (cn >> sh) & 0xFu, (ch >> sh) & 0xFu, d, nm, h1);
            return 1;
          }
          ++checks;
        }
      }
    }
  }
  // prmt emulation against its definition on a few fixed vectors (sign-replicate mode included)
  if (tpg::prmt(0x80017F00u, 0u, 0xBA98u) != 0xFF000000u || tpg::prmt(0x33221100u, 0x77665544u, 0x7531u) != 0x77553311u) {
    printf("prmt emulation fails\n");
    return 1;
  }
  printf("ok %ld\n", checks);
  return 0;
}
