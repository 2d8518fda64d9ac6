// philox.cuh — Philox4x32-10 counter-based uniforms (Salmon et al., Random123), host + device.
// The reference draws from the global numpy MT19937 (np.random.choice, mcts_self_play.py:106); a batched
// engine needs a stream that does not depend on scheduling, so games are keyed by their global id:
//   key = seed (lo, hi), counter = (game_lo, game_hi, draw index >> 2, stream id)
//   One Philox block serves four consecutive draws of a game: draw d takes output word d & 3, u = word / 2^32
//   (m53 = word << 21, so that u = m53 / 2^53 like a replayed double).  2^-32 is finer than the fp32 rounding of the
//   reference's own softmax, and the rollout kernel runs the 10 rounds once per four stones instead of once per stone.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IAGO_PHD __host__ __device__ __forceinline__
#else
#define IAGO_PHD inline
#endif

namespace iago {

IAGO_PHD void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__)
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
    const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
}

// The four output words of the block that serves draws 4*block .. 4*block + 3 of `game`.
IAGO_PHD void philox_block(uint64_t seed, uint64_t game, uint32_t block, uint32_t stream, uint32_t (&out)[4]) {
    uint32_t c0 = (uint32_t)game, c1 = (uint32_t)(game >> 32), c2 = block, c3 = stream;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit integer m with u = m / 2^53 for draw number `draw` of `game`.
IAGO_PHD uint64_t philox_m53(uint64_t seed, uint64_t game, uint32_t draw, uint32_t stream) {
    uint32_t o[4];
    philox_block(seed, game, draw >> 2, stream, o);
    const uint32_t w = (draw & 2) ? ((draw & 1) ? o[3] : o[2]) : ((draw & 1) ? o[1] : o[0]);
    return (uint64_t)w << 21;
}

}  // namespace iago
