/*
 * pmvs_rng.h — counter-based replacement of the reference PSO's srand(time)+rand() (TMVS/pso/psosolver.cpp:60-68).
 *
 * The reference seeds every solver with the wall clock and draws rand() inside OpenMP loops, so it is not
 * reproducible even against itself. Here draw number `ctr` of solver run `run` of patch `patchId` is a pure
 * function of (run seed, patchId, run, ctr): splitmix64 finaliser over a Weyl sequence, top 31 bits, which plays the
 * role of glibc's rand() (RAND_MAX = 2^31-1). random() = rand31 / RAND_MAX lies in [0,1] like psosolver.cpp:66-68.
 *
 * Draw order of one solver (single-thread order of the reference):
 *   ctor initParticles   psosolver.cpp:100-108   for d<3: for i<P: pos -> ctr 2*(d*P+i), vel -> ctr 2*(d*P+i)+1
 *   setParticle(init)    psosolver.cpp:273-281   vel of particle 0, d = 0..2      -> ctr 6P+d
 *   iteration `it`       psosolver.cpp:232-237   particle i: pVecW,gVecW,lVecW,nVecW -> ctr base+4*(it*P+i)+{0..3},
 *                                                base = 6P+3 after setParticle, 6P without it
 */
#ifndef PMVS_RNG_H
#define PMVS_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define PMVS_HD __host__ __device__ __forceinline__
#else
#define PMVS_HD inline
#endif

#define PMVS_GOLD 0x9E3779B97F4A7C15ULL

PMVS_HD uint64_t pmvs_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
PMVS_HD uint64_t pmvs_stream_key(uint64_t seed, int patchId, int run) {
    uint64_t k = pmvs_mix64(seed + PMVS_GOLD * (uint64_t)(uint32_t)(patchId + 1));
    return pmvs_mix64(k + PMVS_GOLD * (uint64_t)(uint32_t)(run + 1));
}
PMVS_HD uint32_t pmvs_rand31(uint64_t key, uint64_t ctr) { return (uint32_t)(pmvs_mix64(key + PMVS_GOLD * (ctr + 1)) >> 33); }
PMVS_HD double pmvs_random(uint64_t key, uint64_t ctr) { return ((double)pmvs_rand31(key, ctr)) / 2147483647.0; }

#endif
