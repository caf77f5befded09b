/* TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
 *
 * Philox4x32-10 (Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy
 * as 1, 2, 3", SC'11), written from the published algorithm.  This is the
 * counter-based generator that replaces the reference's bulk cuRAND XORWOW
 * arrays (reference call sites: src/ising3d_gpu_m.f90:64-65,179;
 * src/ising2d_gpu_m.f90:56-57,138; src/clock_gpu_m.f90:73-74,188-189;
 * src/xy2d_periodic_gpu_m.f90:74-75,355-356).  The product has its own,
 * independently written device copy (csrc/philox.cuh); both are pinned to the
 * Random123 known-answer vectors in tests/test_oracle_rng.py.
 */
#ifndef ORC_PHILOX_H
#define ORC_PHILOX_H
#include <stdint.h>

#define ORC_PHILOX_M0 0xD2511F53u
#define ORC_PHILOX_M1 0xCD9E8D57u
#define ORC_PHILOX_W0 0x9E3779B9u
#define ORC_PHILOX_W1 0xBB67AE85u

static inline void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                                     uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)ORC_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)ORC_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += ORC_PHILOX_W0;
        k1 += ORC_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#endif
