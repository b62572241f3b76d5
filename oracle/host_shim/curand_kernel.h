/* oracle/host_shim/curand_kernel.h -- TEST INFRASTRUCTURE ONLY.
 * Host stand-in for the three cuRAND device-API names the reference uses (curandState, curand_init,
 * curand_uniform). The generator is Philox4x32-10 keyed by the seed with (subsequence, draw index) as the
 * counter: statistically equivalent to XORWOW, not the same stream (the reference's results are compared
 * statistically, never per random number). curand_uniform keeps cuRAND's (0,1] convention. */
#ifndef RLPT_ORACLE_CURAND_SHIM_H
#define RLPT_ORACLE_CURAND_SHIM_H
#include <cstdint>
struct curandStateXORWOW { uint64_t seed; uint64_t subsequence; uint64_t draw; uint32_t buf[4]; int have; };
typedef curandStateXORWOW curandState;
static inline void rlpt_shim_philox(uint32_t k0, uint32_t k1, uint32_t c[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static inline void curand_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset, curandState* s) {
    s->seed = seed; s->subsequence = subsequence; s->draw = offset; s->have = 0;
}
static inline float curand_uniform(curandState* s) {
    if (s->have == 0) {
        uint32_t c[4] = { (uint32_t)s->draw, (uint32_t)(s->draw >> 32), (uint32_t)s->subsequence, (uint32_t)(s->subsequence >> 32) };
        rlpt_shim_philox((uint32_t)s->seed, (uint32_t)(s->seed >> 32), c);
        for (int i = 0; i < 4; ++i) s->buf[i] = c[i];
        s->have = 4; s->draw++;
    }
    uint32_t x = s->buf[4 - s->have]; s->have--;
    return x * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}
#endif
