// bitboard.cuh — Othello rules on 64-bit bitboard pairs (device + host).
//
// Replaces the reference's per-cell ray walks (GameFunctions.legal_actions game.py:209-235 and
// place_stone game.py:179-207, copy-pasted into mcts_self_play.py:36-89, src/rl_self_play.py:36-88,
// rl_env.py:88-138) with shift-and-mask floods (parallel-prefix form).  Bit k <-> action k = row*8+col, so ascending bit
// order is the reference's ascending action list.
//
//   own = stones of the side to move, opp = the other side.
//   Direction d moves a stone set by +1 (east), +8 (south), +9 (south-east), +7 (south-west) or the
//   negatives.  Horizontal and diagonal floods run on opp & 0x7E7E.. so a run can never wrap around a
//   board edge (an edge-column stone cannot be in the interior of a horizontal/diagonal bracket).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define IAGO_HD __host__ __device__ __forceinline__
#else
#define IAGO_HD inline
#endif

namespace iago {

typedef unsigned long long u64;

constexpr u64 kInnerCols = 0x7E7E7E7E7E7E7E7EULL;

template <int S>
IAGO_HD u64 shl(u64 x) { return x << S; }
template <int S>
IAGO_HD u64 shr(u64 x) { return x >> S; }

// Flood of `gen` through `mask` along +S (left shift) / -S, Kogge-Stone style: after the first two steps a run of up to 2
// stones is covered, each doubling step (through `pre` = stones whose neighbour in the direction of travel is also in the mask)
// adds 2 more, so four dependent steps cover the longest possible run on an 8-wide board (6) where a step-by-step flood needs six.
template <int S>
IAGO_HD u64 fill_up(u64 gen, u64 mask) {
    u64 t = mask & (gen << S);
    t |= mask & (t << S);
    const u64 pre = mask & (mask << S);
    t |= pre & (t << (2 * S));
    t |= pre & (t << (2 * S));
    return t;
}
template <int S>
IAGO_HD u64 fill_dn(u64 gen, u64 mask) {
    u64 t = mask & (gen >> S);
    t |= mask & (t >> S);
    const u64 pre = mask & (mask >> S);
    t |= pre & (t >> (2 * S));
    t |= pre & (t >> (2 * S));
    return t;
}
template <int S>
IAGO_HD u64 moves_up(u64 own, u64 mask) { return fill_up<S>(own, mask) << S; }   // the bracket cell beyond the run
template <int S>
IAGO_HD u64 moves_dn(u64 own, u64 mask) { return fill_dn<S>(own, mask) >> S; }

// legal_actions: empty cells from which some direction has >= 1 opponent stone and then an own stone.
IAGO_HD u64 legal_moves(u64 own, u64 opp) {
    const u64 mo = opp & kInnerCols;
    u64 m = moves_up<1>(own, mo) | moves_dn<1>(own, mo);
    m |= moves_up<8>(own, opp) | moves_dn<8>(own, opp);
    m |= moves_up<7>(own, mo) | moves_dn<7>(own, mo);
    m |= moves_up<9>(own, mo) | moves_dn<9>(own, mo);
    return m & ~(own | opp);
}

template <int S>
IAGO_HD u64 flips_up(u64 mv, u64 own, u64 mask) {
    const u64 t = fill_up<S>(mv, mask);
    return ((t << S) & own) ? t : 0ULL;
}
template <int S>
IAGO_HD u64 flips_dn(u64 mv, u64 own, u64 mask) {
    const u64 t = fill_dn<S>(mv, mask);
    return ((t >> S) & own) ? t : 0ULL;
}

// Stones of `opp` bracketed by placing `mv` (single bit) for `own`.  No legality check, like the reference.
IAGO_HD u64 flips_for(u64 mv, u64 own, u64 opp) {
    const u64 mo = opp & kInnerCols;
    u64 f = flips_up<1>(mv, own, mo) | flips_dn<1>(mv, own, mo);
    f |= flips_up<8>(mv, own, opp) | flips_dn<8>(mv, own, opp);
    f |= flips_up<7>(mv, own, mo) | flips_dn<7>(mv, own, mo);
    f |= flips_up<9>(mv, own, mo) | flips_dn<9>(mv, own, mo);
    return f;
}

// place_stone(state, action, color): the cell becomes `own` whatever it held (game.py:185), then flips.
IAGO_HD void place(u64 mv, u64 &own, u64 &opp) {
    own |= mv;
    opp &= ~mv;
    const u64 f = flips_for(mv, own, opp);
    own |= f;
    opp &= ~f;
}

}  // namespace iago
