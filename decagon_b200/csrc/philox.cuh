// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// The reference's randomness (TF's unseeded dropout and fixed_unigram_candidate_sampler,
// layers.py:23-31,112 and optimizer.py:40-47) is not reproducible; the product therefore owns
// counter-based streams that the oracle restates bit-exactly:
//   element e of (relation, stream, step) = word (e & 3) of
//   Philox(counter = (e >> 2, relation, stream, step), key = (seed_lo, seed_hi)).
#pragma once
#include <stdint.h>

namespace dgn {

constexpr uint32_t kStreamDropout1 = 1, kStreamDropout2 = 2, kStreamNegatives = 3;

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

}  // namespace dgn
