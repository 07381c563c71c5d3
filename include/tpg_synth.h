/*
 * tpg_synth.h — deterministic synthetic genotypes for the benchmark shapes.
 *
 * A counter-based integer hash (splitmix64 finaliser) keyed by (seed, locus, individual),
 * so any shard can be generated independently and the CPU and GPU generators are
 * byte-identical (integer arithmetic only).  The law follows the reference's own fixture
 * builder data-raw/create_ibs_dataset.R:48-79: per-locus allele frequency ~ U(0.05, 0.5),
 * genotype ~ Binomial(2, p), missingness iid Bernoulli(rate).
 *
 * mode 0: diploid, missing -> byte 3 (the FBM.code256 NA code, R/gen_tibble_fbm.R:193)
 * mode 1: "imputed" file: missing -> byte 4 + g (codes 4/5/6 as bigsnpr::snp_fastImputeSimple
 *         leaves them; the pairwise path still treats them as missing — SURVEY.md Q2)
 * mode 2: mixed ploidy: 30% of individuals pseudohaploid (bytes 0/2 only), 5% tetraploid-coded
 *         (bytes 0..4, NA = 5), the rest diploid (NA = 3)
 */
#ifndef TPG_SYNTH_H
#define TPG_SYNTH_H
#include <stdint.h>

#if defined(__CUDACC__)
#define TPG_HD __host__ __device__ __forceinline__
#else
#define TPG_HD static inline
#endif

TPG_HD uint64_t tpg_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

/* miss_thr = floor(missing_rate * 2^32), computed once on the host. */
TPG_HD uint8_t tpg_synth_byte(uint64_t seed, uint64_t locus, uint64_t indiv, uint32_t miss_thr,
                              int mode) {
  const uint64_t hl = tpg_mix64(seed ^ tpg_mix64(locus * 2u + 1u));
  /* allele frequency threshold: 0.05 + 0.45 u, in units of 2^-32 */
  const uint32_t thr = 214748365u + (uint32_t)(((hl >> 32) * 1932735283ull) >> 32);
  const uint64_t hc = tpg_mix64(hl ^ (indiv * 0x9E3779B97F4A7C15ull + 0x1234567ull));
  const uint64_t hm = tpg_mix64(hc + 0xABCDEFull);
  const uint32_t a1 = (uint32_t)hc < thr, a2 = (uint32_t)(hc >> 32) < thr;
  const int missing = (uint32_t)hm < miss_thr;
  uint32_t g = a1 + a2;
  uint32_t na = 3;
  if (mode == 2) {
    const uint32_t cls = (uint32_t)(tpg_mix64(seed * 31u + indiv) % 100u);
    if (cls < 30u) {
      g = 2u * a1; /* pseudohaploid: homozygous calls only */
    } else if (cls < 35u) {
      const uint64_t h2 = tpg_mix64(hc ^ 0x5555ull);
      g += ((uint32_t)h2 < thr) + ((uint32_t)(h2 >> 32) < thr); /* tetraploid dosage 0..4 */
      na = 5;
    }
  }
  if (missing) return (uint8_t)(mode == 1 ? 4u + g : na);
  return (uint8_t)g;
}

#endif /* TPG_SYNTH_H */
