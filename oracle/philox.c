/* oracle/philox.c -- Philox4x32-10 (Salmon et al., SC'11) and the uniform conversion of DESIGN.md "RNG".
 * TEST INFRASTRUCTURE (see oracle.h).  The reference draws from R's global Mersenne-Twister stream
 * (e.g. src/multinomial.cpp:8,18), which cannot be reproduced on a GPU; both this oracle and the CUDA
 * library therefore use the same counter-based streams so whole filters can be compared seed-for-seed. */
#include "oracle.h"

static inline void mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  *lo = (uint32_t)p;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo(0xD2511F53u, c0, &hi0, &lo0);
    mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* first uniform (52 bits) + the raw words z, w of the same block: models that split the second half into two 32-bit uniforms */
void orc_philox_uniform1_raw(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double* ua, uint32_t* z, uint32_t* w) {
  uint32_t ctr[4] = {c0, c1, c2, c3}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  orc_philox4x32_10(ctr, key, r);
  uint64_t xa = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
  *ua = ((double)(xa >> 12) + 0.5) * 0x1p-52;
  *z = r[2]; *w = r[3];
}
/* two uniforms in the OPEN interval (0,1): ((x >> 12) + 0.5) * 2^-52, exact in fp64 */
void orc_philox_uniform2(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, double* ua, double* ub) {
  uint32_t ctr[4] = {c0, c1, c2, c3}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, r[4];
  orc_philox4x32_10(ctr, key, r);
  uint64_t xa = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
  uint64_t xb = (uint64_t)r[2] | ((uint64_t)r[3] << 32);
  *ua = ((double)(xa >> 12) + 0.5) * 0x1p-52;
  *ub = ((double)(xb >> 12) + 0.5) * 0x1p-52;
}

const char* orc_version(void) { return "cpimh-oracle 0.2 (resamplers, tree, create_A_, SVLevy pinned to the reference's own sources: oracle/ref_build; mvnorm / lorenz96 unpinned)"; }

/* elementwise rng_log / rng_sincos2pi (rngmath.h) for the tests */
#include "rngmath.h"
void orc_rng_transform(const double* u, int64_t n, double* lg, double* sn, double* cs) {
  for (int64_t i = 0; i < n; i++) { lg[i] = orc_rng_log(u[i]); orc_rng_sincos2pi(u[i], &sn[i], &cs[i]); }
}
