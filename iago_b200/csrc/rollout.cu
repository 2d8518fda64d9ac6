// rollout.cu — K1+K2: the lockstep rollout kernels and the batched rules entry points.
//
// Main path (every weight set that passes the FAST test below): rollout_pair_kernel — TWO lanes own one game, lane 1 on the board
// turned by 180 degrees, so both run the same four flood directions and each owns the legal cells of one half of the board; see the
// comment above that kernel.  rollout_kernel (one thread per game, the first version of this file) remains for weight sets that
// need the SAFE sampler, and rollout_sample_kernel serves single draws.  A turn is
//   legal moves (shift-and-mask floods / carry propagation, bitboard.cuh)      <- GameFunctions.legal_actions
//   rollout policy at the legal cells only: two 512-entry table reads (9-bit 3x3 neighbourhood of each plane) and the cell's
//   bias term from shared memory                                                  <- RolloutPolicy.__call__ network.py:59-64
//   inverse cdf against one Philox / replayed uniform                            <- Simulate.get_action mcts_self_play.py:100-110
//   flips + board update                                                         <- place_stone
//   pass / terminal bookkeeping exactly as Simulate.turn / __call__ (mcts_self_play.py:25-29,124-134)
// and the game ends with judge (mcts_self_play.py:113-121).  Nothing goes to global memory during the game except the optional
// move log; algorithmic HBM traffic is 17 B in + 17-21 B out per game.
//
// Canonical rollout arithmetic (identical, bit for bit, in oracle/othello_ref.c).  S_c = sum of the set taps of plane c in
// ascending tap order, logit = (S0 + S1) + bias[k].
//   FAST (max|S0| + max|S1| + max|bias| <= 300, i.e. any finite sanely trained rollout net; decided when the weights are loaded):
//     w_k = (E0 * E1) * EB in double with E_c = canon_exp(S_c), EB = canon_exp(bias[k]) from look-up tables — the softmax
//     numerator as a product of exponentials, no exp and no max pass in the kernel.  Two-sided cdf: A_i = double running sum over
//     the legal cells 0..31 ascending, D_j = over the legal cells 63..32 DESCENDING, total = A_last + D_last, T = u * total with
//     u = m53 / 2^53; if T < A_last (or no legal cell lies above 31) the move is the first cell with A_i > T, otherwise the cell at
//     position #{j <= n_hi - 2 : D_j < total - T} of the descending list.  In exact arithmetic this is np.random.choice's
//     searchsorted(cdf, u, 'right').
//   SAFE (any other weights): e = exp32(logit - max_legal); q = floor(e * 2^26) (uint32); pick the first legal k with
//     cum_q > floor((m53 >> 21) * total / 2^32).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "bitboard.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace iago {

constexpr int kBlock = 64;     // threads (= games) per CTA
constexpr int kMaxLegal = 32;  // scratch slots per game; more legal moves than this (hill-climbed maximum is 34) take the recompute path

// ---------------------------------------------------------------- device helpers

// A board plane pre-shifted left by 9 as three 32-bit words: the 19-bit window starting at bit k then holds columns j-1..j+1 of
// rows i-1, i, i+1 of cell k = i*8 + j at bits 0-2, 8-10, 16-18.  Rows outside the board read 0 (the conv's zero padding); the
// column that wraps around at j = 0 / j = 7 is removed by colmask[j].
struct Plane3 {
    uint32_t w0, w1, w2;
};
__device__ __forceinline__ Plane3 plane3(u64 bb) {
    const uint32_t lo = (uint32_t)bb, hi = (uint32_t)(bb >> 32);
    Plane3 p;
    p.w0 = lo << 9;
    p.w1 = __funnelshift_l(lo, hi, 9);
    p.w2 = hi >> 23;
    return p;
}
// Byte offset (index * 4) into a 512-entry pattern table: the multiply gathers the three 3-bit fields into 9 contiguous bits
// (row i+1 -> bits 0-2, row i -> 3-5, row i-1 -> 6-8; all partial products land on disjoint bits, so there are no carries).
__device__ __forceinline__ uint32_t pat_offset(const Plane3 &p, int k, uint32_t cm) {
    const bool up = k >= 32;
    const uint32_t a = up ? p.w1 : p.w0, b = up ? p.w2 : p.w1;
    const uint32_t t = __funnelshift_r(a, b, (uint32_t)k) & cm;  // shift amount is taken mod 32
    return ((t * 0x400801u) >> 14) & 0x7FCu;
}
__device__ __forceinline__ float table_at(const float *table, uint32_t byte_off) {
    return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(table) + byte_off);
}

// exp(x) for x <= 0, every step a single IEEE rounding (same sequence as exp32_neg in the oracle).
__device__ __forceinline__ float exp32_neg(float x) {
    if (x < -80.0f) return 0.0f;
    const float z = __fmul_rn(x, 1.44269504088896341f);
    const float n = rintf(z);
    float r = __fmaf_rn(n, -0.693145751953125f, x);
    r = __fmaf_rn(n, -1.42860682030941723e-6f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float r2 = __fmul_rn(r, r);
    float y = __fmaf_rn(p, r2, r);
    y = __fadd_rn(y, 1.0f);
    const float s = __int_as_float(((int)n + 127) << 23);
    return __fmul_rn(y, s);
}

// Shared-memory copy of the tables one kernel needs: float (lut, bias) for logits and the SAFE sampler, double (elut, ebias)
// for FAST.  t0 = channel 0 = opponent stones, t1 = channel 1 = mover's stones (game.py:167-174).
template <class T>
struct PolicySmemT {
    T t0[512], t1[512];
    T tb[64];
    uint32_t colmask[8];
};
typedef PolicySmemT<float> PolicySmem;
typedef PolicySmemT<double> PolicySmemD;

__device__ __forceinline__ void load_policy(PolicySmem &w, const RolloutWeights *__restrict__ gw, int tid, int nthreads) {
    const float *src = &gw->lut[0][0];                                   // [2][512] then [64], contiguous
    for (int i = tid; i < 1024 + 64; i += nthreads) w.t0[i] = src[i];   // t0, t1, tb are contiguous too
    if (tid < 8) w.colmask[tid] = gw->colmask[tid];
}
__device__ __forceinline__ void load_policy(PolicySmemD &w, const RolloutWeights *__restrict__ gw, int tid, int nthreads) {
    const double *src = &gw->elut[0][0];
    for (int i = tid; i < 1024 + 64; i += nthreads) w.t0[i] = src[i];
    if (tid < 8) w.colmask[tid] = gw->colmask[tid];
}

__device__ __forceinline__ float logit_at(const PolicySmem &w, const Plane3 &pm, const Plane3 &po, int k) {
    const uint32_t cm = w.colmask[k & 7];
    const float s0 = table_at(w.t0, pat_offset(po, k, cm));
    const float s1 = table_at(w.t1, pat_offset(pm, k, cm));
    return __fadd_rn(__fadd_rn(s0, s1), w.tb[k]);
}

__device__ __forceinline__ uint32_t q_of(float e) { return __float2uint_rz(__fmul_rn(e, 67108864.0f)); }  // floor(e * 2^26)

// SAFE sampler. sa / sc: this thread's columns of the shared scratch (stride kBlock).
__device__ __noinline__ int sample_move_safe(const PolicySmem &w, u64 own, u64 opp, u64 legal, u64 m53, uint32_t *sa, uint8_t *sc) {
    const int n = __popcll(legal);
    if (n == 1) return __ffsll((long long)legal) - 1;  // same answer as the general path, no arithmetic needed
    const Plane3 pm = plane3(own), po = plane3(opp);
    const uint32_t u32 = (uint32_t)(m53 >> 21);
    if (n <= kMaxLegal) {
        float mx = -3.0e38f;
        {
            u64 m = legal;
            for (int i = 0; i < n; i++) {
                const int k = __ffsll((long long)m) - 1;
                m &= m - 1;
                const float l = logit_at(w, pm, po, k);
                mx = fmaxf(mx, l);
                sa[i * kBlock] = __float_as_uint(l);
                sc[i * kBlock] = (uint8_t)k;
            }
        }
        uint32_t cum = 0;
        for (int i = 0; i < n; i++) {
            cum += q_of(exp32_neg(__fsub_rn(__uint_as_float(sa[i * kBlock]), mx)));  // at most 63 terms of at most 2^26: no overflow
            sa[i * kBlock] = cum;
        }
        const uint32_t T = __umulhi(u32, cum);
        int idx = 0;
        for (int i = 0; i < n; i++) idx += (sa[i * kBlock] <= T) ? 1 : 0;
        idx = min(idx, n - 1);
        return (int)sc[idx * kBlock];
    }
    // > kMaxLegal legal moves: unreachable in real play, possible on arbitrary boards. Recompute instead of storing.
    float mx = -3.0e38f;
    for (u64 m = legal; m; m &= m - 1) mx = fmaxf(mx, logit_at(w, pm, po, __ffsll((long long)m) - 1));
    uint32_t total = 0;
    for (u64 m = legal; m; m &= m - 1)
        total += q_of(exp32_neg(__fsub_rn(logit_at(w, pm, po, __ffsll((long long)m) - 1), mx)));
    const uint32_t T = __umulhi(u32, total);
    uint32_t cum = 0;
    int last = 0;
    for (u64 m = legal; m; m &= m - 1) {
        last = __ffsll((long long)m) - 1;
        cum += q_of(exp32_neg(__fsub_rn(logit_at(w, pm, po, last), mx)));
        if (cum > T) return last;
    }
    return last;
}

// FAST sampler: softmax numerators as table products, double running sum, no exp and no max pass.
__device__ __forceinline__ double weight_at(const PolicySmemD &w, const Plane3 &pm, const Plane3 &po, int k) {
    const uint32_t cm = w.colmask[k & 7];
    const double e0 = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(w.t0) + 2 * pat_offset(po, k, cm));
    const double e1 = *reinterpret_cast<const double *>(reinterpret_cast<const char *>(w.t1) + 2 * pat_offset(pm, k, cm));
    return __dmul_rn(__dmul_rn(e0, e1), w.tb[k]);
}

// One thread, whole board (rollout_sample_kernel, and the rollout kernel of last resort): the two-sided cdf of the canonical rule —
// A = running sum over the legal cells 0..31 ascending, D = over the legal cells 63..32 descending (see the header of this file).
__device__ __noinline__ int sample_move_fast(const PolicySmemD &w, u64 own, u64 opp, u64 legal, u64 m53, double *sa, uint8_t *sc) {
    const int n = __popcll(legal);
    if (n == 1) return __ffsll((long long)legal) - 1;
    const Plane3 pm = plane3(own), po = plane3(opp);
    const double u = __dmul_rn((double)(long long)m53, 1.1102230246251565e-16);  // m53 / 2^53, exact
    const uint32_t llo = (uint32_t)legal, lhi = (uint32_t)(legal >> 32);
    const int nl = __popc(llo), nh = __popc(lhi);
    double lo = 0.0, hi = 0.0;
    for (uint32_t m = llo; m; m &= m - 1) lo = __dadd_rn(lo, weight_at(w, pm, po, __ffs((int)m) - 1));
    for (uint32_t m = lhi; m; m &= ~(0x80000000u >> __clz((int)m))) hi = __dadd_rn(hi, weight_at(w, pm, po, 63 - __clz((int)m)));
    const double total = __dadd_rn(lo, hi), T = __dmul_rn(u, total);
    (void)sa; (void)sc;
    if (T < lo || nh == 0) {
        double cum = 0.0;
        int last = 0;
        for (uint32_t m = llo; m; m &= m - 1) {
            last = __ffs((int)m) - 1;
            cum = __dadd_rn(cum, weight_at(w, pm, po, last));
            if (cum > T) return last;
        }
        return last;
    }
    (void)nl;
    const double R = __dsub_rn(total, T);
    double cum = 0.0;
    int last = 63;
    for (uint32_t m = lhi; m; m &= ~(0x80000000u >> __clz((int)m))) {
        last = 63 - __clz((int)m);
        cum = __dadd_rn(cum, weight_at(w, pm, po, last));
        if (!(cum < R)) return last;
    }
    return last;
}

// The sampler a kernel instantiation uses, with its shared-memory table and scratch types.
template <bool FAST>
struct Sampler;
template <>
struct Sampler<true> {
    typedef PolicySmemD Tables;
    typedef double Scratch;
    static __device__ __forceinline__ int pick(const Tables &w, u64 own, u64 opp, u64 legal, u64 m53, Scratch *sa, uint8_t *sc) {
        return sample_move_fast(w, own, opp, legal, m53, sa, sc);
    }
};
template <>
struct Sampler<false> {
    typedef PolicySmem Tables;
    typedef uint32_t Scratch;
    static __device__ __forceinline__ int pick(const Tables &w, u64 own, u64 opp, u64 legal, u64 m53, Scratch *sa, uint8_t *sc) {
        return sample_move_safe(w, own, opp, legal, m53, sa, sc);
    }
};

// ---------------------------------------------------------------- kernels

struct RolloutArgs {
    const u64 *p1, *p2;
    const uint8_t *color;
    long long n;
    uint32_t stream_id;
    u64 seed, game_id0;
    const double *uniforms;
    long long u_stride;
    const int8_t *forced;
    long long f_stride;
    int8_t *result;
    u64 *final_p1, *final_p2;
    int32_t *n_moves;
    int8_t *move_log;
    u64 *counters;
    const u64 *game_ids;  // nullable: Philox game id of game g (default game_id0 + g)
    int game_threads;     // paired kernel: threads of a CTA that own games (the rest only help to fill the tables); 0 = all
    u64 *counters_out;    // paired kernel, nullable: the CTA that finishes last writes the two totals here (host-mapped memory: no
    unsigned *ticket;     //   copy behind the kernel) and leaves counters[] and the ticket at zero for the next launch
};

template <int MODE, bool LOG, bool FAST>
__global__ void __launch_bounds__(kBlock) rollout_kernel(RolloutArgs a, const RolloutWeights *__restrict__ gw) {
    __shared__ typename Sampler<FAST>::Tables w;
    __shared__ typename Sampler<FAST>::Scratch scratch_a[kMaxLegal * kBlock];
    __shared__ uint8_t scratch_b[kMaxLegal * kBlock];
    if (MODE != IAGO_RNG_FORCED) load_policy(w, gw, threadIdx.x, kBlock);
    __syncthreads();

    const long long g = (long long)blockIdx.x * kBlock + threadIdx.x;
    int placed = 0, turns = 0;
    if (g < a.n) {
        const int color = a.color[g];
        const u64 gid = a.game_ids ? a.game_ids[g] : a.game_id0 + (u64)g;
        u64 own = (color == 1) ? a.p1[g] : a.p2[g];
        u64 opp = (color == 1) ? a.p2[g] : a.p1[g];
        int stone_num = __popcll(own | opp);  // 64 - sum(state == 0), mcts_self_play.py:15
        bool pass_flg = false;
        typename Sampler<FAST>::Scratch *sa = scratch_a + threadIdx.x;
        uint8_t *sb = scratch_b + threadIdx.x;
        uint32_t rnd[4] = {0, 0, 0, 0};  // the Philox block serving draws 4*(placed >> 2) .. + 3
        while (stone_num < 64) {  // mcts_self_play.py:26 — the end test runs once per PAIR of turns
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                const u64 legal = legal_moves(own, opp);
                turns++;
                if (legal) {
                    int k;
                    if (MODE == IAGO_RNG_FORCED) {
                        k = (placed < a.f_stride) ? a.forced[g * a.f_stride + placed] : -1;
                    } else {
                        u64 m53;
                        if (MODE == IAGO_RNG_UNIFORMS)
                            m53 = __double2ull_rz((placed < a.u_stride ? a.uniforms[g * a.u_stride + placed] : 0.5) * 9007199254740992.0);   // never past a short replay stream
                        else {
                            if ((placed & 3) == 0) philox_block(a.seed, gid, (uint32_t)placed >> 2, a.stream_id, rnd);
                            const uint32_t wd = (placed & 2) ? ((placed & 1) ? rnd[3] : rnd[2]) : ((placed & 1) ? rnd[1] : rnd[0]);
                            m53 = (u64)wd << 21;
                        }
                        k = Sampler<FAST>::pick(w, own, opp, legal, m53, sa, sb);
                    }
                    if (MODE == IAGO_RNG_FORCED && (k < 0 || k > 63)) {
                        stone_num = 64;  // replay stream exhausted: stop this game where it stands
                    } else {
                        place(1ULL << k, own, opp);
                        if (LOG) a.move_log[g * 64 + placed] = (int8_t)k;
                        placed++;
                        pass_flg = false;
                        stone_num++;
                    }
                } else {
                    if (pass_flg) stone_num = 64;  // two consecutive passes end the game
                    pass_flg = true;
                }
                const u64 t = own; own = opp; opp = t;
            }
        }
        // an even number of swaps happened: own = stones of `color` again
        const int me = __popcll(own), op = __popcll(opp);
        a.result[g] = (int8_t)((me > op) - (me < op));
        if (a.final_p1) {
            a.final_p1[g] = (color == 1) ? own : opp;
            a.final_p2[g] = (color == 1) ? opp : own;
        }
        if (a.n_moves) a.n_moves[g] = placed;
        if (LOG)
            for (int i = placed; i < 64; i++) a.move_log[g * 64 + i] = -1;
    }
    if (a.counters) {
        unsigned p = (unsigned)placed, t = (unsigned)turns;
        p = __reduce_add_sync(0xFFFFFFFFu, p);
        t = __reduce_add_sync(0xFFFFFFFFu, t);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(a.counters + 0, (u64)p);
            atomicAdd(a.counters + 1, (u64)t);
        }
    }
}

// ---------------------------------------------------------------- the paired rollout kernel (FAST weights)
//
// TWO lanes own one game.  Lane h = 0 holds the board as it is, lane h = 1 holds it turned by 180 degrees (bit k -> 63 - k, one
// BREV per word).  Othello's rules do not change under that turn, and it maps the four "right shift" flood directions onto the
// four "left shift" ones — so both lanes run the SAME instructions: each floods only the directions +1, +7, +8, +9 of its own
// frame (half of movegen, half of the flips) and handles the legal cells 0..31 of its own frame (board rows 0-3 in lane 0, rows
// 7-4 in lane 1) with policy tables built for its frame (elut / elut_r).  Per turn the pair exchanges through shuffles: the half
// move sets, the half flip sets (each 2 x 32 bits + BREV), the two half sums of the softmax numerators and the candidate cell.
// Against one thread per game this doubles the warps per SM (65,536 games = 27.7 warps per SM instead of 13.8), removes the
// word selects from the per-cell loop and shortens the divergent loop from max(n) over 32 games to max(n_half) over 32 halves.
// One CTA per SM: the tables (84 KB in bank strips, see PairLayout below) are held once per SM, next to 16 scratch slots per lane.
// 65,536 games = 4,096 warps = 147 CTAs of 28 warps; small batches use smaller CTAs (host side).
constexpr int kPairMaxWarps = 28;
constexpr int kPairSlots = 16;    // running sums kept per lane; a half board with more legal cells (never seen in play) recomputes

struct __align__(16) PairCell {   // what the per-cell loop needs about cell c of a frame, one 16-byte load
    double eb;                    // exp(bias) of the cell
    uint32_t cm, pad;             // 3x3 window mask of its column
};
// Shared-memory layout of the paired kernel.  The look-up tables are laid out in bank STRIPS so that the per-cell gathers never
// conflict, whatever the lanes ask for:
//   * pattern tables (64 KB at a 64 KB boundary of the shared window): row = 9-bit pattern index (128 B per row), strip = 8 bytes
//     = one bank pair, lane l reads strip l & 15 — lanes l and l + 16 share a strip, so a 64-bit gather of a whole warp is two
//     wavefronts, the minimum for 256 B.  A legal cell is empty, so the centre bit of its pattern (index bit 4) is 0 in both planes:
//     plane 0 (opponent) lives in the rows with bit 4 clear and plane 1 (mover) in the rows with bit 4 set.  Strip parity = lane
//     parity = frame (odd lanes hold the board turned by 180 degrees), so a strip holds its own frame's table.
//     address = TB | (index << 7) | (plane << 11) | ((lane & 15) << 3): the same three instructions as a plain table.
//   * cell records (4 KB): row = cell, strip = 16 bytes, lane l reads strip l & 7 (four lanes per strip = the four wavefronts a
//     512 B gather needs anyway).
//   * line masks of the flips (16 KB): row = (cell, direction pair), strip = 16 bytes = line[d], line[d + 1]; two 16-byte gathers.
// The scratch columns of the warps fill the space below and above the pattern tables.  The host asks for the largest dynamic
// window the device offers (one CTA per SM whatever its size); the kernel lays the pieces out around the 64 KB boundary inside
// that window and traps if they do not fit.
struct PairScratch {              // per warp
    double cum[kPairSlots][32];
    uint8_t cell[kPairSlots][32];
    uint32_t rnd[4][32];          // each lane's current Philox block (kept out of the register file)
};
constexpr uint32_t kPairPatBytes = 0x10000u, kPairCellBytes = 32u * 128u, kPairLineBytes = 2u * 64u * 128u;
struct PairLayout {               // shared-window addresses
    uint32_t pat, cells, line, scratch;
};
__device__ __forceinline__ PairLayout pair_layout(uint32_t w0, uint32_t dyn_bytes, int warp, int nwarps) {
    PairLayout L;
    L.pat = (w0 + 0xFFFFu) & ~0xFFFFu;
    const uint32_t S = (uint32_t)sizeof(PairScratch);
    const uint32_t below = (L.pat - ((w0 + 15u) & ~15u)) / S;          // warps whose scratch fits below the pattern tables
    const uint32_t above = (uint32_t)nwarps > below ? (uint32_t)nwarps - below : 0u;
    L.scratch = (uint32_t)warp < below ? ((w0 + 15u) & ~15u) + (uint32_t)warp * S : L.pat + kPairPatBytes + ((uint32_t)warp - below) * S;
    L.cells = (L.pat + kPairPatBytes + above * S + 127u) & ~127u;
    L.line = L.cells + kPairCellBytes;
    if (L.line + kPairLineBytes > w0 + dyn_bytes) __trap();            // the window is too small: fail loudly
    return L;
}

__device__ __forceinline__ u64 rev64(u64 x) { return ((u64)__brev((uint32_t)x) << 32) | (u64)__brev((uint32_t)(x >> 32)); }
__device__ __forceinline__ u64 pair_xchg(unsigned mask, u64 x) {   // the partner lane's x, turned into this lane's frame
    const uint32_t lo = __shfl_xor_sync(mask, (uint32_t)x, 1), hi = __shfl_xor_sync(mask, (uint32_t)(x >> 32), 1);
    return ((u64)__brev(lo) << 32) | (u64)__brev(hi);
}
// Moves reached by walking towards higher bits.  Along +1 the walk is an addition: a carry injected at the first opponent stone
// east of an own stone runs through the run and leaves a 1 on the cell behind it (edge columns are not in `mo`, so a carry never
// leaves its row); the other three directions flood.  Cells that are not empty are removed by the caller.
__device__ __forceinline__ u64 half_moves(u64 own, u64 opp) {
    const u64 mo = opp & kInnerCols;
    const u64 east = (mo + ((own << 1) & mo)) & ~mo;
    return east | moves_up<8>(own, opp) | moves_up<7>(own, mo) | moves_up<9>(own, mo);
}
// Stones bracketed by the move `mv` = bit k along the four directions towards higher bits, by carry propagation: with every bit
// outside the line L = line[d][k] set, adding mv ripples from k through the opponent run on the line and stops on the first line
// cell that holds no opponent stone; if that cell is own, everything on the line below it is flipped.
__device__ __forceinline__ u64 flip_line(u64 L, u64 mv, u64 own, u64 opp) {
    const u64 out = ((opp | ~L) + mv) & L & own;   // the bracketing stone (one bit) or 0
    const u64 m = out - 1;                          // the bits below it; all ones when there is none: removed by the sign word
    const uint32_t sg = (uint32_t)((int32_t)(uint32_t)(m >> 32) >> 31);
    return m & L & ~(((u64)sg << 32) | sg);
}
// line_base = shared-window address of this lane's 16-byte strip of the line-mask table: row 2k holds line[0][k], line[1][k],
// row 2k + 1 holds line[2][k], line[3][k] (line[d][k] = the cells beyond k along direction +1, +7, +8, +9 up to the board edge)
__device__ __forceinline__ u64 half_flips(uint32_t line_base, int k, u64 mv, u64 own, u64 opp) {
    uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
    const uint32_t addr = line_base + ((uint32_t)k << 8);
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+128];" : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(addr));
    return flip_line(((u64)a1 << 32) | a0, mv, own, opp) | flip_line(((u64)a3 << 32) | a2, mv, own, opp) |
           flip_line(((u64)b1 << 32) | b0, mv, own, opp) | flip_line(((u64)b3 << 32) | b2, mv, own, opp);
}

struct PairPlanes {
    uint32_t o0, o1, m0, m1;   // opponent / mover planes << 10, words 0 and 1: all a window starting at a cell below 32 can reach
};
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// e_base = shared-window address of this lane's 8-byte strip of the pattern tables (64 KB aligned + strip offset), e_base1 =
// e_base | 0x800 (the mover's rows), cell_base = of its 16-byte strip of the cell records
// The cell is given as q = 31 - cell (the per-cell loop walks the bit-reversed legal word from the top, and q is what FLO returns):
// the cell records are stored in that order and the planes are pre-shifted by 10, so that a LEFT funnel shift by q leaves the
// 19-bit window that starts at the cell in the low bits (bits cell .. cell + 31 of plane << 9).
__device__ __forceinline__ double pair_weight(uint32_t e_base, uint32_t e_base1, uint32_t cell_base, const PairPlanes &p, int q) {
    uint32_t ebl, ebh, cm, pad;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ebl), "=r"(ebh), "=r"(cm), "=r"(pad) : "r"(cell_base + ((uint32_t)q << 7)));
    (void)pad;
    const uint32_t t0 = __funnelshift_l(p.o0, p.o1, (uint32_t)q) & cm;
    const uint32_t t1 = __funnelshift_l(p.m0, p.m1, (uint32_t)q) & cm;
    const double e0 = lds_f64((((t0 * 0x400801u) >> 9) & 0xFF80u) | e_base);
    const double e1 = lds_f64((((t1 * 0x400801u) >> 9) & 0xFF80u) | e_base1);
    return __dmul_rn(__dmul_rn(e0, e1), __hiloint2double((int)ebh, (int)ebl));
}
// More legal cells in this half than scratch slots: nothing was stored, walk again.  First cell (own-frame ascending) whose
// running sum fails `pred`, the last cell when none does.
__device__ __noinline__ int pair_pick_slow(uint32_t e_base, uint32_t cells, PairPlanes p, uint32_t wd, int limit, double thr, int h) {
    double cum = 0.0;
    int last = 0, j = 0;
    for (; wd; wd &= wd - 1, j++) {
        last = __ffs((int)wd) - 1;
        cum = __dadd_rn(cum, pair_weight(e_base, e_base | 0x800u, cells, p, 31 - last));
        const bool pred = h ? (cum < thr) : (cum <= thr);
        if (!(j < limit && pred)) break;
    }
    return last;
}

template <int MODE, bool LOG>
__global__ void __launch_bounds__(kPairMaxWarps * 32, 1) rollout_pair_kernel(RolloutArgs a, const RolloutWeights *__restrict__ gw) {
    extern __shared__ __align__(16) unsigned char pair_smem_raw[];
    const uint32_t w0 = (uint32_t)__cvta_generic_to_shared(pair_smem_raw);
    uint32_t dyn_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
    const PairLayout lay = pair_layout(w0, dyn_bytes, threadIdx.x >> 5, (blockDim.x + 31) >> 5);
    if (MODE != IAGO_RNG_FORCED) {
        double *pat = reinterpret_cast<double *>(pair_smem_raw + (lay.pat - w0));
        for (int i = threadIdx.x; i < 8192; i += blockDim.x) {   // entry i = row i >> 4 (pattern, bit 4 = plane), strip i & 15 (parity = frame)
            const int row = i >> 4, plane = (row >> 4) & 1, idx = row & ~0x10;
            pat[i] = (i & 1) ? gw->elut_r[plane][idx] : gw->elut[plane][idx];
        }
        PairCell *cl = reinterpret_cast<PairCell *>(pair_smem_raw + (lay.cells - w0));
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {    // entry i = row i >> 3 (cell), strip i & 7
            const int c = 31 - (i >> 3);                         // row q holds cell 31 - q
            PairCell ci;
            ci.eb = (i & 1) ? gw->ebias_r[c] : gw->ebias[c];
            ci.cm = gw->colmask[c & 7];
            ci.pad = 0;
            cl[i] = ci;
        }
    }
    {
        u64 *ln = reinterpret_cast<u64 *>(pair_smem_raw + (lay.line - w0));
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) {   // entry i = row i >> 4 = 2k + (d >> 1), strip (i >> 1) & 7, word d & 1
            const int k = i >> 5, d = ((i >> 4) & 1) * 2 + (i & 1), r = k >> 3, c = k & 7;
            const int steps = d == 0 ? 7 - c : d == 1 ? min(7 - r, c) : d == 2 ? 7 - r : min(7 - r, 7 - c);
            const int S = d == 0 ? 1 : d + 6;
            u64 L = 0;
            for (int t = 1; t <= steps; t++) L |= 1ULL << (k + t * S);
            ln[i] = L;
        }
    }
    __syncthreads();

    // Control flow is WARP-UNIFORM from here on: every lane runs every turn until the last game of its warp has ended, games that
    // have ended (or never existed: g >= n) and turns without a legal move go through the same instructions with an empty move
    // (no legal cell -> no loop trip, mv = 0 -> no flips).  All shuffles therefore name the full warp — a pair mask would make the
    // compiler guard each of them with a MATCH / vote sequence — and only the per-cell loop diverges.
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const int h = threadIdx.x & 1;
    const unsigned pmask = 3u << (threadIdx.x & 30);
    // small batches run in CTAs of at least 8 warps so that the 84 KB of tables are filled by 256 threads; only the first
    // a.game_threads threads of a CTA own games, the others leave after the fill
    const int gthreads = a.game_threads ? a.game_threads : (int)blockDim.x;
    const long long g = ((long long)blockIdx.x * gthreads + threadIdx.x) >> 1;
    const bool valid = g < a.n && (int)threadIdx.x < gthreads;   // the same for both lanes of a pair, as is every predicate below
    int placed = 0, turns = 0, color = 1, stone_num = 64;
    u64 gid = 0, own = 0, opp = 0;
    if (valid) {
        color = a.color[g];
        gid = a.game_ids ? a.game_ids[g] : a.game_id0 + (u64)g;
        own = (color == 1) ? a.p1[g] : a.p2[g];
        opp = (color == 1) ? a.p2[g] : a.p1[g];
        if (h) { own = rev64(own); opp = rev64(opp); }
        stone_num = __popcll(own | opp);  // 64 - sum(state == 0), mcts_self_play.py:15
    }
    bool pass_flg = false;
    const uint32_t e_base = lay.pat + ((threadIdx.x & 15) << 3);
    uint32_t e_base1;
    asm("or.b32 %0, %1, 0x800;" : "=r"(e_base1) : "r"(e_base));   // opaque: otherwise the OR is redone per cell
    const uint32_t cells = lay.cells + ((threadIdx.x & 7) << 4);
    const uint32_t line_base = lay.line + ((threadIdx.x & 7) << 4);
    PairScratch &scr = *reinterpret_cast<PairScratch *>(pair_smem_raw + (lay.scratch - w0));
    double *sa = &scr.cum[0][threadIdx.x & 31];
    uint8_t *sc = &scr.cell[0][threadIdx.x & 31];
    uint32_t *srnd = &scr.rnd[0][threadIdx.x & 31];
    while (__any_sync(kFull, stone_num < 64)) {
        const bool alive = stone_num < 64;  // mcts_self_play.py:26 — the end test runs once per PAIR of turns
#pragma unroll 1
        for (int half = 0; half < 2; half++) {
            const u64 hm = half_moves(own, opp);
            const u64 legal = (hm | pair_xchg(kFull, hm)) & ~(own | opp);
            bool has = alive && legal != 0;   // this game places a stone in this turn
            turns += alive ? 1 : 0;
            int k;   // the move, in this lane's frame
            if (MODE == IAGO_RNG_FORCED) {
                k = (has && placed < a.f_stride) ? a.forced[g * a.f_stride + placed] : -1;
                if (has && (k < 0 || k > 63)) {
                    stone_num = 64;  // replay stream exhausted: stop this game where it stands
                    has = false;
                }
                if (h) k = 63 - k;
            } else {
                u64 m53;
                if (MODE == IAGO_RNG_UNIFORMS)
                    m53 = __double2ull_rz(((has && placed < a.u_stride) ? a.uniforms[g * a.u_stride + placed] : 0.5) * 9007199254740992.0);   // never past a short replay stream
                else {
                    // lane h computes Philox block 2 * (placed >> 3) + h: the ten rounds run once per EIGHT stones of a
                    // game; the words wait in shared memory, draw d = word d & 3 of lane (d >> 2) & 1 of the pair
                    if (has && (placed & 7) == 0) {
                        uint32_t o[4];
                        philox_block(a.seed, gid, (((uint32_t)placed >> 3) << 1) + h, a.stream_id, o);
                        __syncwarp(pmask);   // the partner has read the last word of the block these stores replace
#pragma unroll
                        for (int j = 0; j < 4; j++) srnd[j * 32] = o[j];
                        __syncwarp(pmask);
                    }
                    const uint32_t wd = srnd[(placed & 3) * 32 + ((placed >> 2) & 1) - h];
                    m53 = (u64)wd << 21;
                }
                const double u = __dmul_rn((double)(long long)m53, 1.1102230246251565e-16);  // m53 / 2^53, exact
                // this lane's half of the softmax numerators: legal cells 0..31 of its frame, ascending
                const uint32_t lw = has ? (uint32_t)legal : 0u;
                const int n = __popc(lw);
                PairPlanes pl;
                pl.o0 = (uint32_t)opp << 10;
                pl.o1 = __funnelshift_l((uint32_t)opp, (uint32_t)(opp >> 32), 10);
                pl.m0 = (uint32_t)own << 10;
                pl.m1 = __funnelshift_l((uint32_t)own, (uint32_t)(own >> 32), 10);
                const bool stored = n <= kPairSlots;
                double cum = 0.0;
                if (stored) {
                    // ascending cells = descending bits of the reversed word: one FLO per cell, pointers instead of an index
                    uint32_t wr = __brev(lw);
                    double *sap = sa;
                    uint8_t *scp = sc;
#pragma unroll 1
                    while (wr) {   // (unrolling makes the warp run the remainders of all its lanes, and requesting the next cell's record
                                   // one trip ahead costs more in register moves than the latency it hides: both measured slower)
                        uint32_t q, below;
                        asm("bfind.u32 %0, %1;" : "=r"(q) : "r"(wr));              // top set bit: q = 31 - cell, wr != 0
                        asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(below) : "r"(q));  // the q bits below it
                        cum = __dadd_rn(cum, pair_weight(e_base, e_base1, cells, pl, (int)q));
                        *sap = cum;
                        *scp = (uint8_t)q;
                        sap += 32;
                        scp += 32;
                        wr &= below;
                    }
                } else {
                    for (uint32_t wd = lw; wd; wd &= wd - 1) cum = __dadd_rn(cum, pair_weight(e_base, e_base1, cells, pl, 32 - __ffs((int)wd)));
                }
                // A_last (lane 0) and D_last (lane 1) -> total, T, and which half holds the answer
                const double oth = __hiloint2double(__shfl_xor_sync(kFull, __double2hiint(cum), 1),
                                                    __shfl_xor_sync(kFull, __double2loint(cum), 1));
                const double lo_t = h ? oth : cum, hi_t = h ? cum : oth;
                const double total = __dadd_rn(lo_t, hi_t), T = __dmul_rn(u, total);
                const bool in_hi = !(T < lo_t) && hi_t > 0.0;
                // lane 0: number of A_i <= T (at most n - 1); lane 1: number of D_j < total - T among j <= n - 2
                const double thr = h ? __dsub_rn(total, T) : T;
                int c;
                if (stored) {
                    // the sums ascend, so the count is found by four fixed steps; v <= thr is v < nextup(thr) (thr >= 0)
                    const double thr2 = __longlong_as_double(__double_as_longlong(thr) + (long long)(1 - h));
                    const int end = n - h;
                    int idx = 0;
#pragma unroll
                    for (int s = 8; s; s >>= 1) {
                        const int t = idx + s;
                        if (t <= end && sa[(t - 1) * 32] < thr2) idx = t;   // slot t - 1 <= 14 exists whatever n is
                    }
                    idx = max(min(idx, n - 1), 0);
                    c = 31 - (int)sc[idx * 32];
                } else {
                    c = pair_pick_slow(e_base, cells, pl, lw, n - h, thr, h);
                }
                const int ct = h ? 63 - c : c;   // this lane's candidate as a cell of the real board
                const int octv = __shfl_xor_sync(kFull, ct, 1);
                const int kt = (in_hi == (h != 0)) ? ct : octv;
                k = h ? 63 - kt : kt;
            }
            k &= 63;   // (a game without a move carries a meaningless k)
            const u64 mv = has ? 1ULL << k : 0ULL;
            own |= mv;
            opp &= ~mv;
            const u64 hf = half_flips(line_base, k, mv, own, opp);   // mv = 0: nothing is bracketed
            const u64 f = hf | pair_xchg(kFull, hf);
            own |= f;
            opp &= ~f;
            if (LOG && has && !h) a.move_log[g * 64 + placed] = (int8_t)k;
            if (has) {
                placed++;
                pass_flg = false;
                stone_num++;
            } else if (alive && (MODE != IAGO_RNG_FORCED || legal == 0)) {
                if (pass_flg) stone_num = 64;  // two consecutive passes end the game
                pass_flg = true;
            }
            const u64 t = own; own = opp; opp = t;
        }
    }
    if (valid && !h) {
        // an even number of swaps happened: own = stones of `color` again
        const int me = __popcll(own), op = __popcll(opp);
        a.result[g] = (int8_t)((me > op) - (me < op));
        if (a.final_p1) {   // nullable for callers that only want the result (mcts.cu)
            a.final_p1[g] = (color == 1) ? own : opp;
            a.final_p2[g] = (color == 1) ? opp : own;
        }
        if (a.n_moves) a.n_moves[g] = placed;
        if (LOG)
            for (int i = placed; i < 64; i++) a.move_log[g * 64 + i] = -1;
    }
    if (a.counters) {
        unsigned p = h ? 0u : (unsigned)placed, t = h ? 0u : (unsigned)turns;
        p = __reduce_add_sync(0xFFFFFFFFu, p);
        t = __reduce_add_sync(0xFFFFFFFFu, t);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(a.counters + 0, (u64)p);
            atomicAdd(a.counters + 1, (u64)t);
        }
        if (a.counters_out) {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) {   // every other CTA fenced its adds before it took its ticket
                    __threadfence();
                    a.counters_out[0] = atomicExch(a.counters + 0, 0ULL);
                    a.counters_out[1] = atomicExch(a.counters + 1, 0ULL);
                    *a.ticket = 0u;
                    __threadfence_system();
                }
            }
        }
    }
}

// One policy draw per board, no board update: Simulate.get_action (mcts_self_play.py:100-110). action -1 = no legal move.
template <int MODE, bool FAST>
__global__ void __launch_bounds__(kBlock) rollout_sample_kernel(const u64 *__restrict__ p1, const u64 *__restrict__ p2,
                                                                const uint8_t *__restrict__ color, long long n,
                                                                uint32_t stream_id, u64 seed, u64 game_id0, uint32_t draw,
                                                                const double *__restrict__ uniforms,
                                                                int8_t *__restrict__ action,
                                                                const RolloutWeights *__restrict__ gw) {
    __shared__ typename Sampler<FAST>::Tables w;
    __shared__ typename Sampler<FAST>::Scratch scratch_a[kMaxLegal * kBlock];
    __shared__ uint8_t scratch_b[kMaxLegal * kBlock];
    load_policy(w, gw, threadIdx.x, kBlock);
    __syncthreads();
    const long long g = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (g >= n) return;
    const bool first = color[g] == 1;
    const u64 own = first ? p1[g] : p2[g], opp = first ? p2[g] : p1[g];
    const u64 legal = legal_moves(own, opp);
    if (!legal) { action[g] = -1; return; }
    const u64 m53 = (MODE == IAGO_RNG_UNIFORMS) ? __double2ull_rz(uniforms[g] * 9007199254740992.0)
                                                 : philox_m53(seed, game_id0 + (u64)g, draw, stream_id);
    action[g] = (int8_t)Sampler<FAST>::pick(w, own, opp, legal, m53, scratch_a + threadIdx.x, scratch_b + threadIdx.x);
}

__global__ void legal_actions_kernel(const u64 *__restrict__ p1, const u64 *__restrict__ p2,
                                     const uint8_t *__restrict__ color, u64 *__restrict__ moves, long long n) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const bool first = color[g] == 1;
    const u64 a = p1[g], b = p2[g];
    moves[g] = legal_moves(first ? a : b, first ? b : a);
}

__global__ void place_stone_kernel(u64 *__restrict__ p1, u64 *__restrict__ p2, const int8_t *__restrict__ action,
                                   const uint8_t *__restrict__ color, long long n) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int k = action[g];
    if (k < 0 || k > 63) return;  // -1 = pass = no-op (game.py:181-182)
    const bool first = color[g] == 1;
    u64 own = first ? p1[g] : p2[g], opp = first ? p2[g] : p1[g];
    place(1ULL << k, own, opp);
    p1[g] = first ? own : opp;
    p2[g] = first ? opp : own;
}

__global__ void rollout_logits_kernel(const u64 *__restrict__ p1, const u64 *__restrict__ p2,
                                      const uint8_t *__restrict__ color, float *__restrict__ logits, long long n,
                                      const RolloutWeights *__restrict__ gw) {
    __shared__ PolicySmem w;
    load_policy(w, gw, threadIdx.x, blockDim.x);
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (game, cell)
    const long long g = idx >> 6;
    if (g >= n) return;
    const bool first = color[g] == 1;
    const u64 a = p1[g], b = p2[g];
    logits[idx] = logit_at(w, plane3(first ? a : b), plane3(first ? b : a), (int)(idx & 63));
}

// Integer-issue roofline denominator: 8 independent rotate+xor chains per thread (SHF + LOP3 on the alu pipe).
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t *out, int iters, uint32_t salt) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0x9E3779B9u + blockIdx.x + i * 0x85EBCA6Bu + salt;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                uint32_t r;
                asm volatile("shf.l.wrap.b32 %0, %1, %1, 5;" : "=r"(r) : "r"(x[i]));
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(x[i]) : "r"(r), "r"(x[(i + 1) & 7]), "r"(salt));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i];
    if (acc == 0x12345678u) out[0] = acc;  // practically never; keeps the chains alive
}

// ---------------------------------------------------------------- host side

// exp(x) in double with a fixed operation sequence (Cody-Waite reduction + degree-13 Taylor, Horner), so that this library and
// the C oracle compute bit-identical tables whatever libm is installed.  Relative error ~2e-16, far below the float rounding.
static double canon_exp(double x) {
    const double n = nearbyint(x * 1.4426950408889634);
    double r = x - n * 0.693147180369123816490;   // ln2 high part (exact product for |n| < 2^20)
    r = r - n * 1.90821492927058770002e-10;        // ln2 low part
    double p = 1.0 / 6227020800.0;                 // 1/13!
    const double inv[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                            1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
    for (int i = 0; i < 13; i++) p = p * r + inv[i];
    return ldexp(p, (int)n);
}

// Returns true when the FAST sampler applies: every partial product of exp tables lies within e^-300 .. e^300 (no double
// overflow / underflow; 63 terms of at most e^300 sum to a finite double).
static bool build_rollout_weights(const float *W, const float *b, RolloutWeights *out) {
    float amax[2] = {0.0f, 0.0f}, bmax = 0.0f;
    int finite = 1;
    for (int c = 0; c < 2; c++)
        for (int idx = 0; idx < 512; idx++) {
            float acc = 0.0f;
            for (int t = 0; t < 9; t++)                       // taps ascending: the canonical summation order
                if (idx >> (6 - 3 * (t / 3) + t % 3) & 1) acc = acc + W[c * 9 + t];
            out->lut[c][idx] = acc;
            out->elut[c][idx] = canon_exp((double)acc);
            if (fabsf(acc) > amax[c]) amax[c] = fabsf(acc);
            if (!(fabsf(acc) <= 300.0f)) finite = 0;   /* also catches NaN */
        }
    memcpy(out->bias, b, 64 * sizeof(float));
    for (int k = 0; k < 64; k++) {
        out->ebias[k] = canon_exp((double)b[k]);
        if (fabsf(b[k]) > bmax) bmax = fabsf(b[k]);
        if (!(fabsf(b[k]) <= 300.0f)) finite = 0;
    }
    for (int c = 0; c < 2; c++)
        for (int idx = 0; idx < 512; idx++) {
            int r = 0;
            for (int bit = 0; bit < 9; bit++) r |= ((idx >> bit) & 1) << (8 - bit);
            out->elut_r[c][idx] = out->elut[c][r];
        }
    for (int k = 0; k < 64; k++) out->ebias_r[k] = out->ebias[63 - k];
    for (int j = 0; j < 8; j++) out->colmask[j] = 0x070707u & ~(j == 0 ? 0x010101u : 0u) & ~(j == 7 ? 0x040404u : 0u);
    const float span = amax[0] + amax[1] + bmax;
    return finite && span <= 300.0f;   // NaN / inf weights take the SAFE path
}

template <int MODE, bool FAST>
static void launch_rollout(const RolloutArgs &a, const RolloutWeights *w, cudaStream_t s) {
#ifndef IAGO_ROLLOUT_SINGLE
    if (FAST) {   // two lanes per game; CTAs as large as it takes to cover the batch with one CTA per SM, at most 28 warps
        static int sms = 0;
        static bool attr_set[2] = {false, false};   // per instantiation and LOG variant; the attribute is per function and sticky
        if (!sms) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (sms <= 0) sms = 148;
        }
        const long long warps = (2 * a.n + 31) / 32;
        long long wpc = (warps + sms - 1) / sms;
        wpc = wpc < 2 ? 2 : wpc > kPairMaxWarps ? kPairMaxWarps : wpc;
        const unsigned grid = (unsigned)((warps + wpc - 1) / wpc);
        RolloutArgs ah = a;
        ah.game_threads = (int)wpc * 32;
        const unsigned cta_threads = (unsigned)(wpc < 8 ? 8 : wpc) * 32;   // helper warps for the table fill of a small CTA
        static int smem = 0;                          // the whole opt-in window (the kernel's layout is built around its 64 KB boundary)
        if (!smem) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        }
        if (a.move_log) {
            if (!attr_set[1]) cudaFuncSetAttribute(rollout_pair_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr_set[1] = true;
            rollout_pair_kernel<MODE, true><<<grid, cta_threads, (size_t)smem, s>>>(ah, w);
        } else {
            if (!attr_set[0]) cudaFuncSetAttribute(rollout_pair_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            attr_set[0] = true;
            rollout_pair_kernel<MODE, false><<<grid, cta_threads, (size_t)smem, s>>>(ah, w);
        }
        return;
    }
#endif
    const unsigned grid = (unsigned)((a.n + kBlock - 1) / kBlock);
    if (a.move_log)
        rollout_kernel<MODE, true, FAST><<<grid, kBlock, 0, s>>>(a, w);
    else
        rollout_kernel<MODE, false, FAST><<<grid, kBlock, 0, s>>>(a, w);
}

static void launch_rollout_mode(const RolloutArgs &a, int mode, bool fast, const RolloutWeights *w, cudaStream_t s) {
    switch (mode) {
        case IAGO_RNG_PHILOX:
            fast ? launch_rollout<IAGO_RNG_PHILOX, true>(a, w, s) : launch_rollout<IAGO_RNG_PHILOX, false>(a, w, s);
            break;
        case IAGO_RNG_UNIFORMS:
            fast ? launch_rollout<IAGO_RNG_UNIFORMS, true>(a, w, s) : launch_rollout<IAGO_RNG_UNIFORMS, false>(a, w, s);
            break;
        default: launch_rollout<IAGO_RNG_FORCED, true>(a, w, s); break;   // no sampling: FAST is irrelevant
    }
}

static int rollout_launch(iago_ctx *ctx, const RolloutArgs &a, int mode, cudaStream_t s) {
    IAGO_CUDA(cudaEventRecord(ctx->ev0, s));
    launch_rollout_mode(a, mode, ctx->rollout_fast, ctx->d_rollout, s);
    IAGO_CUDA(cudaGetLastError());
    IAGO_CUDA(cudaEventRecord(ctx->ev1, s));
    ctx->timed = true;
    return IAGO_OK;
}

static int check_rng(const iago_rng *rng, bool need_weights, const iago_ctx *ctx) {
    IAGO_REQUIRE(rng != nullptr, "rng is NULL");
    IAGO_REQUIRE(rng->mode >= IAGO_RNG_PHILOX && rng->mode <= IAGO_RNG_FORCED, "rng.mode");
    if (rng->mode == IAGO_RNG_UNIFORMS) IAGO_REQUIRE(rng->uniforms && rng->u_stride > 0, "rng.uniforms / u_stride");
    if (rng->mode == IAGO_RNG_FORCED) IAGO_REQUIRE(rng->forced && rng->f_stride > 0, "rng.forced / f_stride");
    if (need_weights && rng->mode != IAGO_RNG_FORCED && !ctx->rollout_loaded) {
        set_error("rollout weights not loaded (call iago_load_rollout first)");
        return IAGO_E_STATE;
    }
    return IAGO_OK;
}

int rollout_launch_ids(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                       uint64_t seed, uint32_t stream_id, const uint64_t *game_ids, int8_t *result, uint64_t *final_p1,
                       uint64_t *final_p2, void *stream) {
    if (!ctx->rollout_loaded) {
        set_error("rollout weights not loaded (call iago_load_rollout first)");
        return IAGO_E_STATE;
    }
    if (n == 0) return IAGO_OK;
    RolloutArgs a{(const u64 *)p1, (const u64 *)p2, color, n, stream_id, seed, 0, nullptr, 0, nullptr, 0, result,
                  (u64 *)final_p1, (u64 *)final_p2, nullptr, nullptr, nullptr, (const u64 *)game_ids};
    launch_rollout_mode(a, IAGO_RNG_PHILOX, ctx->rollout_fast, ctx->d_rollout, (cudaStream_t)stream);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

}  // namespace iago

using namespace iago;

extern "C" {

int iago_load_rollout(iago_ctx *ctx, const float *conv1_W, const float *bias2_b) {
    IAGO_REQUIRE(ctx && conv1_W && bias2_b, "NULL argument");
    DeviceGuard guard(ctx->device);
    RolloutWeights *h = new RolloutWeights();
    const bool fast = build_rollout_weights(conv1_W, bias2_b, h);
    cudaError_t e = cudaMemcpyAsync(ctx->d_rollout, h, sizeof *h, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    delete h;
    IAGO_CUDA(e);
    ctx->rollout_loaded = true;
    ctx->rollout_fast = fast;
    return IAGO_OK;
}

int iago_legal_actions(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                       uint64_t *moves, int64_t n, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && moves, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    legal_actions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const u64 *)p1, (const u64 *)p2, color,
                                                                      (u64 *)moves, n);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_place_stone(iago_ctx *ctx, uint64_t *p1, uint64_t *p2, const int8_t *action, const uint8_t *color,
                     int64_t n, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && action && color, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    place_stone_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((u64 *)p1, (u64 *)p2, action, color, n);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_rollout_logits(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                        float *logits, int64_t n, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && logits, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    if (!ctx->rollout_loaded) {
        set_error("rollout weights not loaded (call iago_load_rollout first)");
        return IAGO_E_STATE;
    }
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    rollout_logits_kernel<<<(unsigned)((n * 64 + 255) / 256), 256, 0, s>>>((const u64 *)p1, (const u64 *)p2, color,
                                                                           logits, n, ctx->d_rollout);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_rollout_sample(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                        const iago_rng *rng, uint32_t draw, int8_t *action, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && action, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    int rc = check_rng(rng, true, ctx);
    if (rc) return rc;
    IAGO_REQUIRE(rng->mode != IAGO_RNG_FORCED, "rng.mode FORCED is meaningless for a single draw");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
#define IAGO_SAMPLE(MODE_, FAST_, U_)                                                                                          \
    rollout_sample_kernel<MODE_, FAST_><<<grid, kBlock, 0, s>>>((const u64 *)p1, (const u64 *)p2, color, n, rng->stream_id, \
                                                                  rng->seed, rng->game_id0, draw, U_, action, ctx->d_rollout)
    if (rng->mode == IAGO_RNG_UNIFORMS) {
        if (ctx->rollout_fast) IAGO_SAMPLE(IAGO_RNG_UNIFORMS, true, rng->uniforms); else IAGO_SAMPLE(IAGO_RNG_UNIFORMS, false, rng->uniforms);
    } else {
        if (ctx->rollout_fast) IAGO_SAMPLE(IAGO_RNG_PHILOX, true, nullptr); else IAGO_SAMPLE(IAGO_RNG_PHILOX, false, nullptr);
    }
#undef IAGO_SAMPLE
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_rollout(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                 const iago_rng *rng, int8_t *result, uint64_t *final_p1, uint64_t *final_p2,
                 int32_t *n_moves, int8_t *move_log, uint64_t *counters, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && result && final_p1 && final_p2, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    int rc = check_rng(rng, true, ctx);
    if (rc) return rc;
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    RolloutArgs a{(const u64 *)p1, (const u64 *)p2, color, n, rng->stream_id, rng->seed, rng->game_id0,
                  rng->uniforms, rng->u_stride, rng->forced, rng->f_stride, result, (u64 *)final_p1,
                  (u64 *)final_p2, n_moves, move_log, (u64 *)counters, nullptr};
    return rollout_launch(ctx, a, rng->mode, s);
}

// Host-buffer entry point.  The batch is cut into up to kHostChunks chunks, each with its own region of the device staging
// buffer and its own stream: while chunk c runs, chunk c+1 is being copied in and chunk c-1 copied out.  Philox game ids stay
// global (game_id0 + index), so the results do not depend on the chunking.
//   pageable caller buffers: packed into / unpacked from the context's pinned staging buffer on the host (that memcpy overlaps
//                            with the copies and kernels of the other chunks);
//   pinned caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory — every buffer of the call): the async copies
//                            read and write the caller's memory directly, no host-side packing at all; and when the buffers are
//                            mapped into the device's address space and the uniforms come from Philox, the kernel itself reads
//                            and writes them (one launch, see below).
constexpr int kHostChunks = 4;

// True when p is NULL or page-locked host memory known to CUDA; *dev (optional) receives the address the device can use for it
// (NULL when the allocation is not mapped into the device's address space).
static bool is_pinned(const void *p, void **dev = nullptr) {
    if (dev) *dev = nullptr;
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (dev) *dev = at.devicePointer;
    return at.type == cudaMemoryTypeHost;
}

int iago_rollout_host(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                      const iago_rng *rng, int8_t *result, uint64_t *final_p1, uint64_t *final_p2,
                      int32_t *n_moves, int8_t *move_log, uint64_t *counters_host) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && result && final_p1 && final_p2, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    int rc = check_rng(rng, true, ctx);
    if (rc) return rc;
    if (counters_host) counters_host[0] = counters_host[1] = 0;
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    auto up8 = [](size_t x) { return (x + 7) & ~(size_t)7; };
    const size_t u_stride = rng->mode == IAGO_RNG_UNIFORMS ? (size_t)rng->u_stride : 0;
    const size_t f_stride = rng->mode == IAGO_RNG_FORCED ? (size_t)rng->f_stride : 0;
    void *m_p1, *m_p2, *m_col, *m_res, *m_f1, *m_f2, *m_nm, *m_log;
    const bool direct = is_pinned(p1, &m_p1) & is_pinned(p2, &m_p2) & is_pinned(color, &m_col) & is_pinned(result, &m_res) &
                        is_pinned(final_p1, &m_f1) & is_pinned(final_p2, &m_f2) & is_pinned(n_moves, &m_nm) &
                        is_pinned(move_log, &m_log) && (!u_stride || is_pinned(rng->uniforms)) && (!f_stride || is_pinned(rng->forced));
    const bool mapped = direct && rng->mode == IAGO_RNG_PHILOX && m_p1 && m_p2 && m_col && m_res && m_f1 && m_f2 &&
                        (m_nm || !n_moves) && (m_log || !move_log);
    // measured at 65,536 games (tools/e2e_probe.py): staged 4 / 2 / 1 chunks 0.43 / 0.46 / 0.52 ms, direct 0.37 / 0.37 / 0.38, mapped 0.33
    const int chunks = mapped ? 1 : n >= 4 * 4096 ? (direct ? 2 : kHostChunks) : 1;
    const size_t per = (((size_t)n + chunks - 1) / chunks + 63) & ~(size_t)63;   // games per chunk, a whole number of CTAs
    // per-chunk region: input block p1 | p2 | color | replay stream ; output block fp1 | fp2 | n_moves | counters | result | log
    const size_t o_p1 = 0, o_p2 = o_p1 + 8 * per, o_col = o_p2 + 8 * per, o_rep = o_col + up8(per);
    const size_t in_bytes = o_rep + 8 * per * u_stride + up8(per * f_stride);
    const size_t o_f1 = in_bytes, o_f2 = o_f1 + 8 * per, o_nm = o_f2 + 8 * per, o_cnt = o_nm + up8(4 * per);
    const size_t o_res = o_cnt + 16, o_log = o_res + up8(per);
    const size_t region = up8(o_log + (move_log ? 64 * per : 0)) + 64;
    rc = ensure_staging(ctx, region * chunks);
    if (rc) return rc;
    if (!ctx->host_streams[0]) {
        ctx->host_streams[0] = ctx->stream;
        for (int c = 1; c < kHostChunks; c++) IAGO_CUDA(cudaStreamCreateWithFlags(&ctx->host_streams[c], cudaStreamNonBlocking));
    }
    const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
    if (mapped) {
        // Pinned, device-mapped caller buffers and no replay stream: ONE launch over the whole batch whose loads and stores cross
        // PCIe themselves (17 B in / 21 B out per game, coalesced, once per game) — no staging, no copy engines, full occupancy.
        // The plies / turns totals come back the same way: the last CTA writes them into the pinned staging block (no memset in
        // front of the launch, no copy behind it: ctx->d_counters[0..1] and the ticket in [2] are zero between launches).
        cudaStream_t s = ctx->stream;
        char *h = (char *)ctx->stage.host;
        void *h_cnt_dev = nullptr;
        const bool publish = ctx->rollout_fast;   // the paired kernel; the one-thread kernel of unusual weight sets keeps the copy
        if (publish) IAGO_CUDA(cudaHostGetDevicePointer(&h_cnt_dev, h + o_cnt, 0));
        else IAGO_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 16, s));
        RolloutArgs a{(const u64 *)m_p1, (const u64 *)m_p2, (const uint8_t *)m_col, (long long)n, rng->stream_id, rng->seed,
                      rng->game_id0, nullptr, 0, nullptr, 0, (int8_t *)m_res, (u64 *)m_f1, (u64 *)m_f2, (int32_t *)m_nm,
                      (int8_t *)m_log, (u64 *)ctx->d_counters, nullptr, 0, (u64 *)h_cnt_dev, (unsigned *)(ctx->d_counters + 2)};
        IAGO_CUDA(cudaEventRecord(ctx->ev0, s));
        launch_rollout_mode(a, rng->mode, ctx->rollout_fast, ctx->d_rollout, s);
        IAGO_CUDA(cudaGetLastError());
        IAGO_CUDA(cudaEventRecord(ctx->ev1, s));
        if (!publish) IAGO_CUDA(cudaMemcpyAsync(h + o_cnt, ctx->d_counters, 16, cudaMemcpyDeviceToHost, s));
        ctx->timed = true;
        IAGO_CUDA(cudaStreamSynchronize(s));
        if (counters_host) memcpy(counters_host, h + o_cnt, 16);
        return IAGO_OK;
    }
    for (int c = 0; c < chunks; c++) {
        const size_t g0 = (size_t)c * per;
        if (g0 >= (size_t)n) break;
        const size_t N = ((size_t)n - g0 < per) ? (size_t)n - g0 : per;
        char *h = (char *)ctx->stage.host + region * c, *d = (char *)ctx->stage.dev + region * c;
        cudaStream_t s = ctx->host_streams[c];
        if (direct) {
            IAGO_CUDA(cudaMemcpyAsync(d + o_p1, p1 + g0, 8 * N, H2D, s));
            IAGO_CUDA(cudaMemcpyAsync(d + o_p2, p2 + g0, 8 * N, H2D, s));
            IAGO_CUDA(cudaMemcpyAsync(d + o_col, color + g0, N, H2D, s));
            if (u_stride) IAGO_CUDA(cudaMemcpyAsync(d + o_rep, rng->uniforms + g0 * u_stride, 8 * N * u_stride, H2D, s));
            if (f_stride) IAGO_CUDA(cudaMemcpyAsync(d + o_rep, rng->forced + g0 * f_stride, N * f_stride, H2D, s));
        } else {
            memcpy(h + o_p1, p1 + g0, 8 * N);
            memcpy(h + o_p2, p2 + g0, 8 * N);
            memcpy(h + o_col, color + g0, N);
            if (u_stride) memcpy(h + o_rep, rng->uniforms + g0 * u_stride, 8 * N * u_stride);
            if (f_stride) memcpy(h + o_rep, rng->forced + g0 * f_stride, N * f_stride);
            IAGO_CUDA(cudaMemcpyAsync(d, h, in_bytes, H2D, s));
        }
        IAGO_CUDA(cudaMemsetAsync(d + o_cnt, 0, 16, s));
        RolloutArgs a{(const u64 *)(d + o_p1), (const u64 *)(d + o_p2), (const uint8_t *)(d + o_col), (long long)N,
                      rng->stream_id, rng->seed, rng->game_id0 + g0, (const double *)(d + o_rep), rng->u_stride,
                      (const int8_t *)(d + o_rep), rng->f_stride, (int8_t *)(d + o_res), (u64 *)(d + o_f1),
                      (u64 *)(d + o_f2), (int32_t *)(d + o_nm), move_log ? (int8_t *)(d + o_log) : nullptr,
                      (u64 *)(d + o_cnt), nullptr};
        if (c == 0) IAGO_CUDA(cudaEventRecord(ctx->ev0, s));
        launch_rollout_mode(a, rng->mode, ctx->rollout_fast, ctx->d_rollout, s);
        IAGO_CUDA(cudaGetLastError());
        if (c == 0) IAGO_CUDA(cudaEventRecord(ctx->ev1, s));   // iago_last_kernel_ms: the first chunk's launch (the whole batch when n < 16,384)
        if (direct) {
            IAGO_CUDA(cudaMemcpyAsync(final_p1 + g0, d + o_f1, 8 * N, D2H, s));
            IAGO_CUDA(cudaMemcpyAsync(final_p2 + g0, d + o_f2, 8 * N, D2H, s));
            IAGO_CUDA(cudaMemcpyAsync(result + g0, d + o_res, N, D2H, s));
            if (n_moves) IAGO_CUDA(cudaMemcpyAsync(n_moves + g0, d + o_nm, 4 * N, D2H, s));
            if (move_log) IAGO_CUDA(cudaMemcpyAsync(move_log + 64 * g0, d + o_log, 64 * N, D2H, s));
            IAGO_CUDA(cudaMemcpyAsync(h + o_cnt, d + o_cnt, 16, D2H, s));
        } else {
            IAGO_CUDA(cudaMemcpyAsync(h + o_f1, d + o_f1, (move_log ? o_log + 64 * N : o_res + N) - o_f1, D2H, s));
        }
    }
    ctx->timed = true;
    for (int c = 0; c < chunks; c++) {
        const size_t g0 = (size_t)c * per;
        if (g0 >= (size_t)n) break;
        const size_t N = ((size_t)n - g0 < per) ? (size_t)n - g0 : per;
        const char *h = (const char *)ctx->stage.host + region * c;
        IAGO_CUDA(cudaStreamSynchronize(ctx->host_streams[c]));
        if (!direct) {
            memcpy(final_p1 + g0, h + o_f1, 8 * N);
            memcpy(final_p2 + g0, h + o_f2, 8 * N);
            memcpy(result + g0, h + o_res, N);
            if (n_moves) memcpy(n_moves + g0, h + o_nm, 4 * N);
            if (move_log) memcpy(move_log + 64 * g0, h + o_log, 64 * N);
        }
        if (counters_host) {
            uint64_t cnt[2];
            memcpy(cnt, h + o_cnt, 16);
            counters_host[0] += cnt[0];
            counters_host[1] += cnt[1];
        }
    }
    return IAGO_OK;
}

// Asynchronous form of iago_rollout_host for a caller that streams batches: up to kHostLanes submissions are in flight, each on its
// own stream — H2D copies of the three input arrays, the kernel on the lane's device block, D2H copies of the results — so the
// copies of one batch run on the copy engines while the kernel of another has the SMs.  Page-locked buffers and Philox uniforms only.
int iago_rollout_host_submit(iago_ctx *ctx, int lane, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                             const iago_rng *rng, int8_t *result, uint64_t *final_p1, uint64_t *final_p2, int32_t *n_moves,
                             int8_t *move_log) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && result && final_p1 && final_p2, "NULL argument");
    IAGO_REQUIRE(lane >= 0 && lane < kHostLanes, "lane out of range");
    IAGO_REQUIRE(n >= 0, "n < 0");
    int rc = check_rng(rng, true, ctx);
    if (rc) return rc;
    IAGO_REQUIRE(rng->mode == IAGO_RNG_PHILOX, "iago_rollout_host_submit draws Philox uniforms only (replay streams: iago_rollout_host)");
    HostLane &l = ctx->lanes[lane];
    if (l.busy) {
        set_error("iago_rollout_host_submit: lane %d has a submission in flight (call iago_rollout_host_wait first)", lane);
        return IAGO_E_STATE;
    }
    IAGO_REQUIRE(is_pinned(p1) && is_pinned(p2) && is_pinned(color) && is_pinned(result) && is_pinned(final_p1) && is_pinned(final_p2) &&
                     is_pinned(n_moves) && is_pinned(move_log),
                 "iago_rollout_host_submit needs page-locked buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory)");
    DeviceGuard guard(ctx->device);
    if (!l.stream) {
        IAGO_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        IAGO_CUDA(cudaMalloc(&l.d_cnt, 4 * sizeof(uint64_t)));
        IAGO_CUDA(cudaMemset(l.d_cnt, 0, 4 * sizeof(uint64_t)));
        IAGO_CUDA(cudaHostAlloc(&l.h_cnt, 2 * sizeof(uint64_t), cudaHostAllocMapped));
    }
    l.h_cnt[0] = l.h_cnt[1] = 0;
    l.busy = true;
    if (n == 0) return IAGO_OK;
    auto up8 = [](size_t x) { return (x + 7) & ~(size_t)7; };
    const size_t N = (size_t)n;
    const size_t o_p1 = 0, o_p2 = o_p1 + 8 * N, o_col = o_p2 + 8 * N, o_f1 = o_col + up8(N), o_f2 = o_f1 + 8 * N, o_nm = o_f2 + 8 * N;
    const size_t o_res = o_nm + up8(4 * N), o_log = o_res + up8(N), bytes = o_log + (move_log ? 64 * N : 0);
    if (l.bytes < bytes) {
        if (l.dev) cudaFree(l.dev);
        l.dev = nullptr;
        l.bytes = 0;
        IAGO_CUDA(cudaMalloc(&l.dev, bytes + bytes / 4));
        l.bytes = bytes + bytes / 4;
    }
    cudaStream_t s = l.stream;
    char *d = l.dev;
    const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
    IAGO_CUDA(cudaMemcpyAsync(d + o_p1, p1, 8 * N, H2D, s));
    IAGO_CUDA(cudaMemcpyAsync(d + o_p2, p2, 8 * N, H2D, s));
    IAGO_CUDA(cudaMemcpyAsync(d + o_col, color, N, H2D, s));
    const bool publish = ctx->rollout_fast;   // the paired kernel's last CTA writes the totals to h_cnt and re-zeroes d_cnt
    void *h_cnt_dev = nullptr;
    if (publish) IAGO_CUDA(cudaHostGetDevicePointer(&h_cnt_dev, l.h_cnt, 0));
    else IAGO_CUDA(cudaMemsetAsync(l.d_cnt, 0, 16, s));
    RolloutArgs a{(const u64 *)(d + o_p1), (const u64 *)(d + o_p2), (const uint8_t *)(d + o_col), (long long)n, rng->stream_id, rng->seed,
                  rng->game_id0, nullptr, 0, nullptr, 0, (int8_t *)(d + o_res), (u64 *)(d + o_f1), (u64 *)(d + o_f2),
                  n_moves ? (int32_t *)(d + o_nm) : nullptr, move_log ? (int8_t *)(d + o_log) : nullptr, (u64 *)l.d_cnt, nullptr, 0,
                  (u64 *)h_cnt_dev, (unsigned *)(l.d_cnt + 2)};
    launch_rollout_mode(a, rng->mode, ctx->rollout_fast, ctx->d_rollout, s);
    IAGO_CUDA(cudaGetLastError());
    IAGO_CUDA(cudaMemcpyAsync(final_p1, d + o_f1, 8 * N, D2H, s));
    IAGO_CUDA(cudaMemcpyAsync(final_p2, d + o_f2, 8 * N, D2H, s));
    IAGO_CUDA(cudaMemcpyAsync(result, d + o_res, N, D2H, s));
    if (n_moves) IAGO_CUDA(cudaMemcpyAsync(n_moves, d + o_nm, 4 * N, D2H, s));
    if (move_log) IAGO_CUDA(cudaMemcpyAsync(move_log, d + o_log, 64 * N, D2H, s));
    if (!publish) IAGO_CUDA(cudaMemcpyAsync(l.h_cnt, l.d_cnt, 16, D2H, s));
    return IAGO_OK;
}

// Blocks until the submission of `lane` has delivered its results; counters_host[2] (nullable) = {stones placed, turns taken}.
int iago_rollout_host_wait(iago_ctx *ctx, int lane, uint64_t *counters_host) {
    IAGO_REQUIRE(ctx != nullptr, "ctx is NULL");
    IAGO_REQUIRE(lane >= 0 && lane < kHostLanes, "lane out of range");
    HostLane &l = ctx->lanes[lane];
    if (!l.busy) {
        set_error("iago_rollout_host_wait: lane %d has nothing in flight", lane);
        return IAGO_E_STATE;
    }
    DeviceGuard guard(ctx->device);
    l.busy = false;
    IAGO_CUDA(cudaStreamSynchronize(l.stream));
    if (counters_host) {
        counters_host[0] = l.h_cnt[0];
        counters_host[1] = l.h_cnt[1];
    }
    return IAGO_OK;
}

int iago_measure_int_peak(iago_ctx *ctx, int iters, double *lane_ops_per_s) {
    IAGO_REQUIRE(ctx && lane_ops_per_s && iters > 0, "bad argument");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    const int blocks = ctx->sm_count * 8, threads = 256;  // 2048 threads / SM: full occupancy
    uint32_t *d_out = (uint32_t *)ctx->d_counters + 8;    // scratch word inside the counters allocation
    int_peak_kernel<<<blocks, threads, 0, s>>>(d_out, 64, 1u);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        IAGO_CUDA(cudaEventRecord(ctx->ev0, s));
        int_peak_kernel<<<blocks, threads, 0, s>>>(d_out, iters, 2u + rep);
        IAGO_CUDA(cudaEventRecord(ctx->ev1, s));
        IAGO_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        IAGO_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (ms < best) best = ms;
    }
    IAGO_CUDA(cudaGetLastError());
    ctx->timed = false;
    const double ops = (double)blocks * threads * (double)iters * 4.0 * 8.0 * 2.0;  // SHF + LOP3 per chain step
    *lane_ops_per_s = ops / (best * 1e-3);
    return IAGO_OK;
}

}  // extern "C"
