// mcts.cu — K5: PV-MCTS on a GPU-resident node pool, many trees per GPU, leaf-parallel with virtual loss.
//
// Reference: MCTS.py:10-76 (Node: P = u = prior + 0.1, select = arg-max of Q + u, U = c_puct * P * sqrt(N_parent) /
// (0.01 + n), update = running mean, update_recursive = the SAME value with the SAME sign all the way to the root)
// and MCTS.py:78-154 (playout: descend; a leaf with n_visits >= n_thr is expanded — no legal move: one pass child
// with prior 1, one legal move: one child with prior 1, otherwise SLPolicy probabilities — and the descent goes on;
// a leaf with n_visits < n_thr is evaluated: v = Value(state, mover), z = Simulate(state)(mover),
// leaf_value = (1 - lambda) v + lambda z; get_move = most visited child, lowest action on ties; update_with_move =
// re-root on the played child or start a fresh tree).
//
// Design (DESIGN.md "PV-MCTS"):
//   * T independent trees (one per game) live in one pool, `cap` 48-byte nodes per tree; children of a node are a
//     contiguous block in ascending action order (= the reference's dict insertion order, so "first maximum" is the
//     same child).  Boards are not stored: a descent replays the moves on a bitboard pair in registers.
//   * One wave = up to B playouts per tree:
//       select (pass 1)  the B descents of a tree run in a fixed order, each seeing the virtual visits of the ones before it
//                        (deterministic) — executed as a software pipeline over 12 warps per tree that preserves exactly that
//                        order (mcts_select_pipe_kernel; mcts_select_kernel is the one-warp form, used for B < 8 and the
//                        exact mode); lanes score the children in parallel (pipelined form: an fp32 screen, float64 only for
//                        near-ties — the arg-max is the float64 one either way); a leaf that must be expanded and needs priors
//                        is parked and queued
//       policy           ONE fused trunk launch over every queued position of every tree (trunk.cu, count on device)
//       expand           priors written, children become visible
//       select (pass 2)  the parked descents continue into the fresh children
//       value            ONE trunk launch over the leaves whose value is not cached yet (Value is a pure function of
//                        the position; the reference re-runs it up to n_thr times on the same leaf)
//       rollout          ONE lockstep rollout launch over all T*B leaves (rollout.cu), Philox keyed by
//                        (global tree id, playout index); it runs on a second stream beside the value launch
//     and the waves of a search after the first replay a captured CUDA graph (iago_mcts_search)
//       backup           exact mode (B = 1): the reference's float32 / float64 running mean, applied in order;
//                        batched mode: atomicAdd on n and on a 2^-40 fixed-point value sum (order-free, so the
//                        result does not depend on scheduling), virtual visits removed
//   * With B = 1 the search is the reference's sequential algorithm, arithmetic type for arithmetic type
//     (numpy scalar rules: P is float32 unless it is the literal 1.1; Q is float32 when lambda < 1, float64 when
//     lambda >= 1; U is float64), which is what the parity tests pin against the reference's own trees.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "bitboard.cuh"
#include "common.cuh"

namespace iago {

constexpr double kFix = 1099511627776.0;  // 2^40: fixed-point scale of the value sum W

enum : uint8_t { F_P_F64 = 1, F_PENDING = 2, F_V_CLAIMED = 4, F_V_VALID = 8 };
enum : uint8_t { S_INACTIVE = 0, S_EVAL = 1, S_PARKED = 2 };

struct __align__(16) MctsNode {
    double P;          // prior + 0.1 (MCTS.py:18-19)
    double Q;          // running mean (exact mode); batched mode reads W / n instead
    long long W;       // sum of leaf values, 2^-40 fixed point
    float v;           // cached Value output for this position (F_V_VALID)
    int n;             // n_visits
    int vn;            // virtual visits of in-flight descents
    int parent;        // -1 at the root
    int first_child;
    int8_t action;     // move that leads here (-1 = pass)
    uint8_t nch;       // number of children (0 = leaf)
    uint8_t flags;
    uint8_t nch_pending;
};
static_assert(sizeof(MctsNode) == 48, "node layout");

struct MctsParams {
    double lambda, c_puct, vloss;
    int n_thr, B, exact, cache_v, need_v, need_z;
    long long target;  // playouts per tree at which this search stops
};

struct MctsDev {
    MctsNode *nodes;   // [T][cap]
    int cap, T;
    int *n_nodes;      // [T]
    u64 *root_p1, *root_p2;
    uint8_t *root_color;
    long long *done;   // [T] playouts finished since the tree was created (indexes rng ids and forced streams)
    u64 *tree_gid;     // [T] global tree id (rng key)
    // per slot [T*B]
    int *leaf_node;
    uint8_t *status, *leaf_color;
    u64 *leaf_p1, *leaf_p2, *game_ids;
    int8_t *z;
    float *v_slot;
    // request lists
    int *counts;       // [0] policy requests, [1] value requests, [2] pool overflows, [3] parked after pass 2
    u64 *pol_p1, *pol_p2;
    uint8_t *pol_color;
    int *pol_node;     // global node index t*cap + i
    float *probs;      // [T*B][64]
    u64 *val_p1, *val_p2;
    uint8_t *val_color;
    int *val_node, *val_slot;
    float *vals;
    const float *forced_v;  // [T][forced_stride] nullable (test hook: replay the reference's leaf evaluations)
    const int8_t *forced_z;
    long long forced_stride;
};

__device__ __forceinline__ double shfl_down_d(double x, int o) {
    int lo = __double2loint(x), hi = __double2hiint(x);
    lo = __shfl_down_sync(0xFFFFFFFFu, lo, o);
    hi = __shfl_down_sync(0xFFFFFFFFu, hi, o);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ int nth_set_bit(u64 m, int i) {
    for (int j = 0; j < i; j++) m &= m - 1;
    return __ffsll((long long)m) - 1;
}

// One warp per tree; the B descents of the tree run one after the other.
__global__ void __launch_bounds__(32) mcts_select_kernel(MctsDev d, MctsParams p, int pass) {
    const int t = blockIdx.x, lane = threadIdx.x;
    MctsNode *nodes = d.nodes + (size_t)t * d.cap;
    int n_nodes = d.n_nodes[t];
    const long long done = d.done[t];
    for (int s = 0; s < p.B; s++) {
        const int slot = t * p.B + s;
        int node, color;
        u64 own, opp;
        if (pass == 1) {
            if (done + s >= p.target) {
                if (lane == 0) {
                    d.status[slot] = S_INACTIVE;
                    d.leaf_p1[slot] = 0; d.leaf_p2[slot] = 0; d.leaf_color[slot] = 1;  // an empty board: the rollout ends at once
                    d.game_ids[slot] = 0;
                }
                continue;
            }
            node = 0;
            color = d.root_color[t];
            own = color == 1 ? d.root_p1[t] : d.root_p2[t];
            opp = color == 1 ? d.root_p2[t] : d.root_p1[t];
            if (lane == 0) {
                nodes[0].vn += 1;
                d.game_ids[slot] = (d.tree_gid[t] << 32) | (u64)(uint32_t)(done + s);
            }
            __syncwarp();
        } else {
            if (d.status[slot] != S_PARKED) continue;
            node = d.leaf_node[slot];
            color = d.leaf_color[slot];
            own = color == 1 ? d.leaf_p1[slot] : d.leaf_p2[slot];
            opp = color == 1 ? d.leaf_p2[slot] : d.leaf_p1[slot];
        }
        // Header of the node the descent stands on.  Loaded from memory for the starting node only: while the children are scored
        // every lane already holds its child's whole node, so the winner's header and action come by shuffle and each level costs
        // ONE dependent round of loads (the children of the chosen node) instead of three.
        int nch, flags, n_here, vn_here, fc;
        {
            const MctsNode *nd = nodes + node;
            nch = nd->nch; flags = nd->flags; n_here = nd->n; vn_here = nd->vn; fc = nd->first_child;
        }
        for (;;) {  // ends at a leaf: every step goes one level down a finite pool
            if (nch == 0) {
                bool park = (flags & F_PENDING) != 0;
                if (!park && n_here >= p.n_thr) {  // MCTS.py:108-121 expand
                    const u64 legal = legal_moves(own, opp);
                    const int k = __popcll(legal), c = k > 0 ? k : 1;
                    if (n_nodes + c <= d.cap) {
                        const int base = n_nodes;
                        n_nodes += c;
                        for (int i = lane; i < c; i += 32) {
                            MctsNode ch;
                            ch.P = k <= 1 ? 1.1 : 0.0;   // Node(node, 1): the literal 1 + 0.1 in float64 (MCTS.py:114,117)
                            ch.Q = 0.0; ch.W = 0; ch.v = 0.0f; ch.n = 0; ch.vn = 0; ch.parent = node; ch.first_child = -1;
                            ch.action = (int8_t)(k == 0 ? -1 : nth_set_bit(legal, i));
                            ch.nch = 0; ch.flags = k <= 1 ? F_P_F64 : 0; ch.nch_pending = 0;
                            nodes[base + i] = ch;
                        }
                        __syncwarp();
                        if (k <= 1) {
                            if (lane == 0) { nodes[node].first_child = base; nodes[node].nch = (uint8_t)c; }
                            __syncwarp();
                            fc = base;
                            nch = c;
                            continue;  // the playout goes on through the only child (MCTS.py:121)
                        }
                        if (lane == 0) {
                            nodes[node].first_child = base;
                            nodes[node].nch_pending = (uint8_t)c;
                            nodes[node].flags = (uint8_t)(flags | F_PENDING);
                            const int j = atomicAdd(d.counts + 0, 1);
                            d.pol_p1[j] = color == 1 ? own : opp;
                            d.pol_p2[j] = color == 1 ? opp : own;
                            d.pol_color[j] = (uint8_t)color;
                            d.pol_node[j] = t * d.cap + node;
                        }
                        park = true;
                    } else if (lane == 0) {
                        atomicAdd(d.counts + 2, 1);  // pool full: evaluate instead of expanding (reported to the host)
                    }
                }
                if (lane == 0) {
                    d.leaf_node[slot] = node;
                    d.leaf_p1[slot] = color == 1 ? own : opp;
                    d.leaf_p2[slot] = color == 1 ? opp : own;
                    d.leaf_color[slot] = (uint8_t)color;
                    if (park) {
                        d.status[slot] = S_PARKED;
                        if (pass == 2) atomicAdd(d.counts + 3, 1);
                    } else {
                        d.status[slot] = S_EVAL;
                        if (p.need_v) {
                            const bool have = p.cache_v && (nodes[node].flags & (F_V_VALID | F_V_CLAIMED));
                            if (!have) {
                                if (p.cache_v) nodes[node].flags |= F_V_CLAIMED;
                                const int j = atomicAdd(d.counts + 1, 1);
                                d.val_p1[j] = color == 1 ? own : opp;
                                d.val_p2[j] = color == 1 ? opp : own;
                                d.val_color[j] = (uint8_t)color;
                                d.val_node[j] = t * d.cap + node;
                                d.val_slot[j] = slot;
                            }
                        }
                    }
                }
                __syncwarp();
                break;
            }
            // ---- select (MCTS.py:39-49): arg-max over children of Q + u, first maximum wins
            const double n_parent = p.exact ? (double)n_here : (double)(n_here + vn_here - 1);  // without this descent's own virtual visit
            const double sq = __dsqrt_rn(n_parent);
            double best = -1.0e300;
            int best_i = 1 << 30;
            int b_n = 0, b_vn = 0, b_fc = -1, b_meta = 0;   // the best child of THIS lane: n, vn, first_child, action | nch | flags | nch_pending
            for (int i = lane; i < nch; i += 32) {
                const uint4 *raw = reinterpret_cast<const uint4 *>(nodes + fc + i);
                const uint4 w0 = raw[0], w1 = raw[1], w2 = raw[2];       // P Q | W v n | vn parent first_child meta
                const double cP = __hiloint2double((int)w0.y, (int)w0.x), cQ = __hiloint2double((int)w0.w, (int)w0.z);
                const long long cW = (long long)(((u64)w1.y << 32) | (u64)w1.x);
                const int cn = (int)w1.w, cvn = (int)w2.x, cflags = (int)((w2.w >> 16) & 0xFFu);
                const double cp = (cflags & F_P_F64) ? __dmul_rn(p.c_puct, cP) : (double)__fmul_rn((float)p.c_puct, (float)cP);
                double q, den;
                if (p.exact) {
                    q = cQ;
                    den = __dadd_rn(0.01, (double)cn);
                } else {
                    const int tot = cn + cvn;
                    q = tot > 0 ? __ddiv_rn(__dsub_rn((double)cW * (1.0 / kFix), __dmul_rn(p.vloss, (double)cvn)), (double)tot) : 0.0;
                    den = __dadd_rn(0.01, (double)tot);
                }
                const double u = __ddiv_rn(__dmul_rn(cp, sq), den);
                const double val = __dadd_rn(q, u);
                if (val > best) {  // ascending i: the first maximum stays
                    best = val; best_i = i;
                    b_n = cn; b_vn = cvn; b_fc = (int)w2.z; b_meta = (int)w2.w;
                }
            }
            int win = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = shfl_down_d(best, o);
                const int oi = __shfl_down_sync(0xFFFFFFFFu, best_i, o);
                const int ow = __shfl_down_sync(0xFFFFFFFFu, win, o);
                if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; win = ow; }
            }
            win = __shfl_sync(0xFFFFFFFFu, win, 0);
            best_i = __shfl_sync(0xFFFFFFFFu, best_i, 0);
            const int child = fc + best_i;
            n_here = __shfl_sync(0xFFFFFFFFu, b_n, win);
            vn_here = __shfl_sync(0xFFFFFFFFu, b_vn, win) + 1;   // with this descent's own virtual visit, stored below
            fc = __shfl_sync(0xFFFFFFFFu, b_fc, win);
            const int meta = __shfl_sync(0xFFFFFFFFu, b_meta, win);
            const int act = (int)(int8_t)(meta & 0xFF);
            nch = (meta >> 8) & 0xFF;
            flags = (meta >> 16) & 0xFF;
            if (lane == 0) nodes[child].vn = vn_here;
            if (act >= 0) place(1ULL << act, own, opp);   // action -1 = pass = no-op (game.py:181-182)
            { const u64 tmp = own; own = opp; opp = tmp; }
            color = 3 - color;
            node = child;
            __syncwarp();
        }
    }
    if (lane == 0) d.n_nodes[t] = n_nodes;
}

// priors: child P = float32(prob[action] + 0.1) (MCTS.py:93-99, :18-19); the children become visible.
__device__ __forceinline__ void expand_one(const MctsDev &d, int j) {
    MctsNode *nd = d.nodes + d.pol_node[j];
    MctsNode *tree = d.nodes + (size_t)(d.pol_node[j] / d.cap) * d.cap;
    const int c = nd->nch_pending, fc = nd->first_child;
    const float *pr = d.probs + (size_t)j * 64;
    for (int i = 0; i < c; i++) {
        MctsNode *ch = tree + fc + i;
        ch->P = (double)__fadd_rn(pr[ch->action], 0.1f);
    }
    nd->nch = (uint8_t)c;
    nd->nch_pending = 0;
    nd->flags &= (uint8_t)~F_PENDING;
}
__global__ void mcts_expand_kernel(MctsDev d) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < d.counts[0]) expand_one(d, j);
}

// ---------------------------------------------------------------- pipelined select
// The same descents in the same order, as a software pipeline: kPipeWarps warps per tree, warp w runs the tree's active descents
// w, w + W, w + 2W, ... .  Descent a may work on stage k (stage 0 = root bookkeeping, stage L + 1 = choosing a child at depth L)
// only after descent a - 1 has finished stage k or ended above it, and may do its leaf work (expansion, request lists, node
// allocation) only after descent a - 1 has ended.  By induction every descent then sees exactly the virtual visits, expansions and
// node indices it would see if the descents ran one after the other, so the tree is bit-identical to mcts_select_kernel's — but a
// level of the tree costs one load round per descent in flight instead of one per descent.
// Progress flags live in shared memory (prog[a] = stages finished, INT_MAX = ended); node data written by one warp is published to
// the other warps of the CTA by __threadfence_block before the flag is raised and read after a fence behind the flag (release /
// acquire at CTA scope; all warps of a tree share one SM and its L1).  In pass 2 stages count from the node a descent was parked
// on: only descents parked on the same node meet in the tree (a pending node has no visible children, so no parked node lies
// below another), those start at the same depth, and the wait chain is transitive.
#ifndef IAGO_PIPE_WARPS
#define IAGO_PIPE_WARPS 12   // 8 / 12 / 16 measured on the final kernel: 24.5 / 24.2 / 24.3 ms per 16,384-playout move of one tree
#endif
constexpr int kPipeWarps = IAGO_PIPE_WARPS;
constexpr int kPipeMaxB = 1024;
constexpr float kScreenEps = 8.0e-6f;   // fp32 screen of the select scores: see mcts_select_pipe_kernel

// place_stone by the whole warp: lanes 0-3 find the stones bracketed along +1, +7, +8, +9, lanes 4-7 along the opposite directions on
// the board turned by 180 degrees (one bit reversal), each by carry propagation along its line (see half_flips in rollout.cu: with
// every bit outside the line set, adding the move ripples through the opponent run and stops on the first line cell that holds no
// opponent stone; if that cell is the mover's, everything on the line below it is flipped) — about 40 instructions and two REDUX.OR
// per level of a descent where every lane used to run all eight floods.  line[d][k] = the cells beyond k along direction d.
__device__ __forceinline__ u64 rev64(u64 x) { return ((u64)__brev((uint32_t)x) << 32) | (u64)__brev((uint32_t)(x >> 32)); }
__device__ __forceinline__ void place_warp(const u64 *line, int act, u64 &own, u64 &opp, int lane) {
    const u64 mv = 1ULL << act;
    own |= mv;          // the cell becomes the mover's whatever it held (game.py:185)
    opp &= ~mv;
    u64 f = 0;
    if (lane < 8) {
        const bool r = lane >= 4;
        const u64 o = r ? rev64(own) : own, q = r ? rev64(opp) : opp;
        const int k = r ? 63 - act : act;
        const u64 L = line[(lane & 3) * 64 + k];
        const u64 out = ((q | ~L) + (1ULL << k)) & L & o;   // the bracketing stone (one bit) or 0
        const u64 ff = out ? (out - 1) & L : 0ULL;
        f = r ? rev64(ff) : ff;
    }
    const uint32_t flo = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)f), fhi = __reduce_or_sync(0xFFFFFFFFu, (uint32_t)(f >> 32));
    f = ((u64)fhi << 32) | flo;
    own |= f;
    opp &= ~f;
}
__device__ __forceinline__ void fill_line_table(u64 *line, int tid, int nthreads) {
    for (int i = tid; i < 256; i += nthreads) {
        const int d = i >> 6, k = i & 63, r = k >> 3, c = k & 7;
        const int steps = d == 0 ? 7 - c : d == 1 ? min(7 - r, c) : d == 2 ? 7 - r : min(7 - r, 7 - c);
        const int S = d == 0 ? 1 : d + 6;
        u64 L = 0;
        for (int t = 1; t <= steps; t++) L |= 1ULL << (k + t * S);
        line[i] = L;
    }
}

__device__ __forceinline__ void pipe_fence() {
    asm volatile("fence.acq_rel.cta;" ::: "memory");   // release before the flag is raised / acquire behind it: all the hand-over needs
}
__device__ __forceinline__ void pipe_wait(const volatile int *prog, int a, int need) {   // until descent a - 1 has progress >= need
    if (a == 0) return;
    while (prog[a - 1] < need) __nanosleep(20);   // (a pure spin measured the same)
    pipe_fence();
}
__device__ __forceinline__ void pipe_signal(volatile int *prog, int a, int value, int lane) {
    __syncwarp();
    pipe_fence();
    if (lane == 0) prog[a] = value;
}

// expand_first (pass 2 of a search with a few trees): the priors of this tree's pending nodes are written by this launch instead of
// a launch of their own in front of it.
__global__ void __launch_bounds__(kPipeWarps * 32) mcts_select_pipe_kernel(MctsDev d, MctsParams p, int pass, int expand_first) {
    __shared__ volatile int prog[kPipeMaxB];
    __shared__ short active[kPipeMaxB];
    __shared__ int n_active;
    __shared__ volatile int s_n_nodes;
    // Evaluation requests are staged per tree (the slot of the asking descent, in descent order) and copied to the global lists when
    // the launch ends, with ONE atomicAdd per list and tree: an atomic with a return value inside the leaf work — which the descents
    // of a tree do strictly one after the other — is an L2 round trip on the critical path of every descent.
    __shared__ short s_pol[kPipeMaxB], s_val[kPipeMaxB];
    __shared__ volatile int s_npol, s_nval;
    __shared__ int s_base_pol, s_base_val;
    __shared__ u64 s_line[256];
    fill_line_table(s_line, threadIdx.x, blockDim.x);
    const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MctsNode *nodes = d.nodes + (size_t)t * d.cap;
    const long long done = d.done[t];
    if (expand_first) {
        const int nreq = d.counts[0];
        for (int j = threadIdx.x; j < nreq; j += blockDim.x)
            if (d.pol_node[j] / d.cap == t) expand_one(d, j);
        __syncthreads();   // (block-wide visibility of the children for the descents below)
    }
    // the descents that run in this pass, in slot order (ballot + per-warp offsets: one thread walking the B status bytes of a parked
    // pass cost 5 us per launch)
    {
        __shared__ int s_wcount[kPipeWarps];
        int base = 0;
        for (int s0 = 0; s0 < p.B; s0 += (int)blockDim.x) {
            const int s = s0 + (int)threadIdx.x;
            const bool on = s < p.B && (pass == 1 ? (done + s < p.target) : (d.status[t * p.B + s] == S_PARKED));
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, on);
            if (lane == 0) s_wcount[warp] = __popc(bal);
            __syncthreads();
            int off = base, tot = 0;
            for (int w = 0; w < kPipeWarps; w++) {
                if (w < warp) off += s_wcount[w];
                tot += s_wcount[w];
            }
            if (on) active[off + __popc(bal & ((1u << lane) - 1u))] = (short)s;
            base += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            n_active = base;
            s_n_nodes = d.n_nodes[t];
            s_npol = 0;
            s_nval = 0;
        }
    }
    for (int i = threadIdx.x; i < p.B; i += blockDim.x) {
        prog[i] = 0;
        if (pass == 1 && done + i >= p.target && i < p.B) {   // idle slot: an empty board, the rollout ends at once
            const int slot = t * p.B + i;
            d.status[slot] = S_INACTIVE;
            d.leaf_p1[slot] = 0; d.leaf_p2[slot] = 0; d.leaf_color[slot] = 1;
            d.game_ids[slot] = 0;
        }
    }
    __syncthreads();
    const int na = n_active;
    for (int a = warp; a < na; a += kPipeWarps) {
        const int s = active[a], slot = t * p.B + s;
        int node, color, stage = 0;
        u64 own, opp;
        int nch, flags, n_here, vn_here, fc;
        pipe_wait(prog, a, 1);
        if (pass == 1) {
            node = 0;
            color = d.root_color[t];
            own = color == 1 ? d.root_p1[t] : d.root_p2[t];
            opp = color == 1 ? d.root_p2[t] : d.root_p1[t];
            const uint4 w1 = reinterpret_cast<const uint4 *>(nodes)[1], w2 = reinterpret_cast<const uint4 *>(nodes)[2];
            n_here = (int)w1.w; vn_here = (int)w2.x + 1; fc = (int)w2.z;
            nch = (w2.w >> 8) & 0xFF; flags = (w2.w >> 16) & 0xFF;
            if (lane == 0) {
                nodes[0].vn = vn_here;
                d.game_ids[slot] = (d.tree_gid[t] << 32) | (u64)(uint32_t)(done + s);
            }
        } else {
            node = d.leaf_node[slot];
            color = d.leaf_color[slot];
            own = color == 1 ? d.leaf_p1[slot] : d.leaf_p2[slot];
            opp = color == 1 ? d.leaf_p2[slot] : d.leaf_p1[slot];
            const uint4 w1 = reinterpret_cast<const uint4 *>(nodes + node)[1], w2 = reinterpret_cast<const uint4 *>(nodes + node)[2];
            n_here = (int)w1.w; vn_here = (int)w2.x; fc = (int)w2.z;
            nch = (w2.w >> 8) & 0xFF; flags = (w2.w >> 16) & 0xFF;
        }
        stage = 1;
        pipe_signal(prog, a, stage, lane);
        bool exclusive = false;   // true once descent a - 1 has ended: nothing this descent reads can change any more
        for (;;) {
            if (nch == 0) {
                if (!exclusive) {
                    pipe_wait(prog, a, 0x7FFFFFFF);
                    exclusive = true;
                    // an earlier descent may have expanded or parked this node after its header was read: read it again
                    const uint4 w1 = reinterpret_cast<const uint4 *>(nodes + node)[1], w2 = reinterpret_cast<const uint4 *>(nodes + node)[2];
                    n_here = (int)w1.w; fc = (int)w2.z;
                    nch = (w2.w >> 8) & 0xFF; flags = (w2.w >> 16) & 0xFF;
                    if (nch != 0) continue;
                }
                int n_nodes = s_n_nodes;
                bool park = (flags & F_PENDING) != 0;
                if (!park && n_here >= p.n_thr) {  // MCTS.py:108-121 expand
                    const u64 legal = legal_moves(own, opp);
                    const int k = __popcll(legal), c = k > 0 ? k : 1;
                    if (n_nodes + c <= d.cap) {
                        const int base = n_nodes;
                        n_nodes += c;
                        for (int i = lane; i < c; i += 32) {
                            MctsNode ch;
                            ch.P = k <= 1 ? 1.1 : 0.0;   // Node(node, 1): the literal 1 + 0.1 in float64 (MCTS.py:114,117)
                            ch.Q = 0.0; ch.W = 0; ch.v = 0.0f; ch.n = 0; ch.vn = 0; ch.parent = node; ch.first_child = -1;
                            ch.action = (int8_t)(k == 0 ? -1 : nth_set_bit(legal, i));
                            ch.nch = 0; ch.flags = k <= 1 ? F_P_F64 : 0; ch.nch_pending = 0;
                            nodes[base + i] = ch;
                        }
                        __syncwarp();
                        if (lane == 0) s_n_nodes = n_nodes;
                        if (k <= 1) {
                            if (lane == 0) { nodes[node].first_child = base; nodes[node].nch = (uint8_t)c; }
                            __syncwarp();
                            __threadfence_block();
                            fc = base;
                            nch = c;
                            continue;  // the playout goes on through the only child (MCTS.py:121)
                        }
                        if (lane == 0) {
                            nodes[node].first_child = base;
                            nodes[node].nch_pending = (uint8_t)c;
                            nodes[node].flags = (uint8_t)(flags | F_PENDING);
                            const int j = s_npol;
                            s_pol[j] = (short)s;
                            s_npol = j + 1;
                        }
                        park = true;
                    } else if (lane == 0) {
                        atomicAdd(d.counts + 2, 1);  // pool full: evaluate instead of expanding (reported to the host)
                    }
                }
                if (lane == 0 && !park && p.need_v) {   // the value of a node is asked for once (cache_v): claimed in descent order
                    const uint8_t fl = (uint8_t)flags;   // read after descent a - 1 had ended (the header re-read above, or a level scored since)
                    const bool have = p.cache_v && (fl & (F_V_VALID | F_V_CLAIMED));
                    if (!have) {
                        if (p.cache_v) nodes[node].flags = (uint8_t)(fl | F_V_CLAIMED);
                        const int j = s_nval;
                        s_val[j] = (short)s;
                        s_nval = j + 1;
                    }
                }
                // everything a later descent may read is written: let it go on, then record the leaf for the kernels behind this one
                pipe_signal(prog, a, 0x7FFFFFFF, lane);
                if (lane == 0) {
                    d.leaf_node[slot] = node;
                    d.leaf_p1[slot] = color == 1 ? own : opp;
                    d.leaf_p2[slot] = color == 1 ? opp : own;
                    d.leaf_color[slot] = (uint8_t)color;
                    d.status[slot] = park ? S_PARKED : S_EVAL;
                    if (park && pass == 2) atomicAdd(d.counts + 3, 1);
                }
                break;
            }
            // ---- select (MCTS.py:39-49): arg-max over children of Q + u, first maximum wins
            // The arg-max is that of the reference's float64 scores, found in two steps.  (1) An fp32 SCREEN: every lane scores its
            // child in single precision (a dozen 4-cycle operations instead of a float64 square root and two float64 divisions: the
            // time a warp spends on a level is what paces the pipeline of descents); one REDUX.MAX gives the best score, and every
            // child within kScreenEps * (2 + |best|) of it — eight times the worst-case distance between an fp32 score and its float64
            // value — is a candidate.  (2) Only when there is more than one candidate (near-ties, or more than 32 children) the
            // float64 scores are computed and compared exactly as before.  A single candidate IS the float64 arg-max, so the tree
            // does not change.  What does not depend on the earlier descents (the node's visit count, the priors) is read before the wait.
            const int n_par = n_here + vn_here - 1;   // without this descent's own virtual visit
            const float sqf = sqrtf((float)n_par);
            uint4 w0 = make_uint4(0, 0, 0, 0);
            float numf = 0.0f;
            if (lane < nch) {
                w0 = reinterpret_cast<const uint4 *>(nodes + fc + lane)[0];
                numf = (float)p.c_puct * (float)__hiloint2double((int)w0.y, (int)w0.x) * sqf;
            }
            if (!exclusive) pipe_wait(prog, a, stage + 1);
            uint4 w1 = make_uint4(0, 0, 0, 0), w2 = make_uint4(0, 0, 0, 0);   // W v n | vn parent first_child meta of the lane's first child
            float valf = -3.0e38f;
            if (lane < nch) {
                const uint4 *raw = reinterpret_cast<const uint4 *>(nodes + fc + lane);
                w1 = raw[1]; w2 = raw[2];
                const long long cW = (long long)(((u64)w1.y << 32) | (u64)w1.x);
                const int cvn = (int)w2.x, tot = (int)w1.w + cvn;
                const float totf = (float)tot;
                const float qf = tot > 0 ? __fdividef((float)cW * (float)(1.0 / kFix) - (float)p.vloss * (float)cvn, totf) : 0.0f;
                valf = qf + __fdividef(numf, 0.01f + totf);
            }
            int b_n = (int)w1.w, b_vn = (int)w2.x, b_fc = (int)w2.z, b_meta = (int)w2.w;   // the best child of THIS lane
            int best_i, win;
            {
                const uint32_t fb = __float_as_uint(valf);
                const uint32_t fkey = fb ^ ((uint32_t)((int32_t)fb >> 31) | 0x80000000u);   // order-preserving
                const uint32_t mkey = __reduce_max_sync(0xFFFFFFFFu, fkey);
                const float mf = __uint_as_float(mkey ^ (((mkey >> 31) - 1u) | 0x80000000u));
                const unsigned cand = __ballot_sync(0xFFFFFFFFu, valf >= mf - kScreenEps * (2.0f + fabsf(mf)));
                win = __ffs((int)cand) - 1;
                best_i = win;
                if (nch > 32 || (cand & (cand - 1)) != 0) {   // warp-uniform: near-ties are settled in float64
                    const double sq = __dsqrt_rn((double)n_par);
                    double best = -1.0e300;
                    best_i = 1 << 30;
                    for (int i = lane; i < nch; i += 32) {
                        uint4 x0 = w0, x1 = w1, x2 = w2;
                        if (i >= 32) {
                            const uint4 *raw = reinterpret_cast<const uint4 *>(nodes + fc + i);
                            x0 = raw[0]; x1 = raw[1]; x2 = raw[2];
                        }
                        const int cflags = (int)((x2.w >> 16) & 0xFFu);
                        const double cP = __hiloint2double((int)x0.y, (int)x0.x);
                        const double cp = (cflags & F_P_F64) ? __dmul_rn(p.c_puct, cP) : (double)__fmul_rn((float)p.c_puct, (float)cP);
                        const long long cW = (long long)(((u64)x1.y << 32) | (u64)x1.x);
                        const int cn = (int)x1.w, cvn = (int)x2.x;
                        const int tot = cn + cvn;
                        const double q = tot > 0 ? __ddiv_rn(__dsub_rn((double)cW * (1.0 / kFix), __dmul_rn(p.vloss, (double)cvn)), (double)tot) : 0.0;
                        const double den = __dadd_rn(0.01, (double)tot);
                        const double u = __ddiv_rn(__dmul_rn(cp, sq), den);
                        const double val = __dadd_rn(q, u);
                        if (val > best) {  // ascending i: the first maximum stays
                            best = val; best_i = i;
                            b_n = cn; b_vn = cvn; b_fc = (int)x2.z; b_meta = (int)x2.w;
                        }
                    }
                    // warp arg-max on an order-preserving integer key: two 32-bit REDUX.MAX (high word, then low word among the lanes
                    // that hold the high maximum) and a REDUX.MIN of the child index among the lanes that hold the maximum — the first
                    // maximum wins, as in the scalar loop (every val is a finite double; -0.0 cannot come out of q + u with u >= +0)
                    const long long bits = __double_as_longlong(best);
                    const u64 key = (u64)bits ^ (u64)((bits >> 63) | (long long)0x8000000000000000LL);
                    const uint32_t khi = (uint32_t)(key >> 32), klo = (uint32_t)key;
                    const uint32_t mhi = __reduce_max_sync(0xFFFFFFFFu, khi);
                    const uint32_t mlo = __reduce_max_sync(0xFFFFFFFFu, khi == mhi ? klo : 0u);
                    const bool top = khi == mhi && klo == mlo;
                    best_i = (int)__reduce_min_sync(0xFFFFFFFFu, top ? (uint32_t)best_i : 0xFFFFFFFFu);
                    win = best_i & 31;   // child i is held by lane i & 31
                }
            }
            const int child = fc + best_i;
            n_here = __shfl_sync(0xFFFFFFFFu, b_n, win);
            vn_here = __shfl_sync(0xFFFFFFFFu, b_vn, win) + 1;   // with this descent's own virtual visit, stored below
            fc = __shfl_sync(0xFFFFFFFFu, b_fc, win);
            const int meta = __shfl_sync(0xFFFFFFFFu, b_meta, win);
            const int act = (int)(int8_t)(meta & 0xFF);
            nch = (meta >> 8) & 0xFF;
            flags = (meta >> 16) & 0xFF;
            if (lane == 0) nodes[child].vn = vn_here;
            stage++;
            if (!exclusive) pipe_signal(prog, a, stage, lane);
            if (act >= 0) place_warp(s_line, act, own, opp, lane);   // action -1 = pass = no-op (game.py:181-182)
            { const u64 tmp = own; own = opp; opp = tmp; }
            color = 3 - color;
            node = child;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        d.n_nodes[t] = s_n_nodes;
        s_base_pol = s_npol ? atomicAdd(d.counts + 0, s_npol) : 0;
        s_base_val = s_nval ? atomicAdd(d.counts + 1, s_nval) : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s_npol; i += blockDim.x) {
        const int slot = t * p.B + s_pol[i], j = s_base_pol + i;
        d.pol_p1[j] = d.leaf_p1[slot];
        d.pol_p2[j] = d.leaf_p2[slot];
        d.pol_color[j] = d.leaf_color[slot];
        d.pol_node[j] = t * d.cap + d.leaf_node[slot];
    }
    for (int i = threadIdx.x; i < s_nval; i += blockDim.x) {
        const int slot = t * p.B + s_val[i], j = s_base_val + i;
        d.val_p1[j] = d.leaf_p1[slot];
        d.val_p2[j] = d.leaf_p2[slot];
        d.val_color[j] = d.leaf_color[slot];
        d.val_node[j] = t * d.cap + d.leaf_node[slot];
        d.val_slot[j] = slot;
    }
}

__device__ __forceinline__ void scatter_value_one(const MctsDev &d, int cache_v, int j) {
    if (j >= d.counts[1]) return;
    const float v = d.vals[j];
    d.v_slot[d.val_slot[j]] = v;
    if (cache_v) {
        MctsNode *nd = d.nodes + d.val_node[j];
        nd->v = v;
        nd->flags = (uint8_t)((nd->flags & ~F_V_CLAIMED) | F_V_VALID);
    }
}
__global__ void mcts_scatter_value_kernel(MctsDev d, int cache_v) { scatter_value_one(d, cache_v, blockIdx.x * blockDim.x + threadIdx.x); }

// leaf_value = (1 - lambda) * v + lambda * z with numpy's scalar types (MCTS.py:123-125): float32 when lambda < 1
// (v is np.float32, Python scalars are weak), float64 when lambda >= 1 (v is the int 0).
__device__ __forceinline__ double leaf_value(const MctsParams &p, float v, int z) {
    if (p.lambda >= 1.0) return (1.0 - p.lambda) * 0.0 + p.lambda * (double)z;
    const float a = __fmul_rn((float)(1.0 - p.lambda), v);
    const float b = (float)(p.lambda * (double)z);
    return (double)__fadd_rn(a, b);
}

__device__ __forceinline__ float slot_v(const MctsDev &d, const MctsParams &p, int t, int slot, int s, const MctsNode *leaf) {
    if (!p.need_v) return 0.0f;
    if (d.forced_v) return d.forced_v[(size_t)t * d.forced_stride + d.done[t] + s];
    return p.cache_v ? leaf->v : d.v_slot[slot];
}
__device__ __forceinline__ int slot_z(const MctsDev &d, const MctsParams &p, int t, int slot, int s) {
    if (!p.need_z) return 0;
    if (d.forced_z) return d.forced_z[(size_t)t * d.forced_stride + d.done[t] + s];
    return d.z[slot];
}

// exact mode: one thread per tree, slots in order, the reference's running mean (MCTS.py:61-72).
// The last kernel of a wave leaves the two request counters at zero for the next wave (one memset node less per wave).
__device__ __forceinline__ void reset_request_counts(const MctsDev &d) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { d.counts[0] = 0; d.counts[1] = 0; }
}

__global__ void mcts_backup_exact_kernel(MctsDev d, MctsParams p) {
    reset_request_counts(d);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T) return;
    MctsNode *nodes = d.nodes + (size_t)t * d.cap;
    int n_done = 0;
    for (int s = 0; s < p.B; s++) {
        const int slot = t * p.B + s;
        if (d.status[slot] == S_INACTIVE) continue;
        n_done++;
        int node = d.leaf_node[slot];
        if (d.status[slot] != S_EVAL) {  // still parked (n_thr < 1 or pool trouble): drop the virtual visits only
            for (; node >= 0; node = nodes[node].parent) nodes[node].vn -= 1;
            continue;
        }
        const double lv = leaf_value(p, slot_v(d, p, t, slot, s, nodes + node), slot_z(d, p, t, slot, s));
        const long long fix = llrint(lv * kFix);
        for (; node >= 0; node = nodes[node].parent) {
            MctsNode *nd = nodes + node;
            const int n = nd->n + 1;
            nd->n = n;
            if (p.lambda >= 1.0) {
                nd->Q = __dadd_rn(nd->Q, __ddiv_rn(__dsub_rn(lv, nd->Q), (double)n));
            } else {
                const float q = (float)nd->Q;
                nd->Q = (double)__fadd_rn(q, __fdiv_rn(__fsub_rn((float)lv, q), (float)n));
            }
            nd->W += fix;
            nd->vn -= 1;
        }
    }
    d.done[t] += n_done;
}

// batched mode: one thread per slot, order-free integer atomics.
__device__ __forceinline__ void backup_atomic_one(const MctsDev &d, const MctsParams &p, int slot) {
    if (slot >= d.T * p.B) return;
    const int t = slot / p.B, s = slot - t * p.B;
    if (d.status[slot] == S_INACTIVE) return;
    MctsNode *nodes = d.nodes + (size_t)t * d.cap;
    int node = d.leaf_node[slot];
    if (d.status[slot] != S_EVAL) {
        for (; node >= 0; node = nodes[node].parent) atomicSub(&nodes[node].vn, 1);
        return;
    }
    const double lv = leaf_value(p, slot_v(d, p, t, slot, s, nodes + node), slot_z(d, p, t, slot, s));
    const long long fix = llrint(lv * kFix);
    for (; node >= 0; node = nodes[node].parent) {
        MctsNode *nd = nodes + node;
        atomicAdd(&nd->n, 1);
        atomicAdd(reinterpret_cast<unsigned long long *>(&nd->W), (unsigned long long)fix);
        atomicSub(&nd->vn, 1);
    }
}
__global__ void mcts_backup_atomic_kernel(MctsDev d, MctsParams p) { backup_atomic_one(d, p, blockIdx.x * blockDim.x + threadIdx.x); }

__device__ __forceinline__ void advance_done_one(const MctsDev &d, const MctsParams &p, int t) {
    if (t >= d.T) return;
    const long long left = p.target - d.done[t];
    d.done[t] += left < p.B ? (left > 0 ? left : 0) : p.B;
}
__global__ void mcts_advance_done_kernel(MctsDev d, MctsParams p) {
    reset_request_counts(d);
    advance_done_one(d, p, blockIdx.x * blockDim.x + threadIdx.x);
}

// The end of a wave as ONE launch when all slots fit one CTA (T * B <= 1,024: a single tree, a few root-parallel trees): every
// playout backed up, then the playout counters advanced (a search of one tree is bound by dependent launches).
__global__ void __launch_bounds__(1024) mcts_tail_kernel(MctsDev d, MctsParams p) {
    reset_request_counts(d);
    backup_atomic_one(d, p, threadIdx.x);
    __syncthreads();
    advance_done_one(d, p, threadIdx.x);
}

__global__ void mcts_init_roots_kernel(MctsDev d, const u64 *p1, const u64 *p2, const uint8_t *color, int reset_tree) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T) return;
    d.root_p1[t] = p1[t]; d.root_p2[t] = p2[t]; d.root_color[t] = color[t];
    if (reset_tree) {
        MctsNode r;
        r.P = 1.1; r.Q = 0.0; r.W = 0; r.v = 0.0f; r.n = 0; r.vn = 0; r.parent = -1; r.first_child = -1;
        r.action = 0; r.nch = 0; r.flags = F_P_F64; r.nch_pending = 0;  // Node(None, 1.0) (MCTS.py:81)
        d.nodes[(size_t)t * d.cap] = r;
        d.n_nodes[t] = 1;
    }
}

// update_with_move (MCTS.py:149-154): re-root on the played child, compacting its subtree (breadth first, children stay
// contiguous and in order) into the other half of the pool; an unknown move starts a fresh tree.  The root position moves too.
__global__ void mcts_advance_kernel(MctsDev d, MctsNode *dst_all, const int8_t *action, const uint8_t *mask) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T) return;
    const MctsNode *src = d.nodes + (size_t)t * d.cap;
    MctsNode *dst = dst_all + (size_t)t * d.cap;
    if (mask && !mask[t]) {  // this tree does not move: carry it over as it is
        const int n = d.n_nodes[t];
        for (int i = 0; i < n; i++) dst[i] = src[i];
        return;
    }
    const int a = action[t];
    int found = -1;
    for (int i = 0; i < src[0].nch; i++)
        if (src[src[0].first_child + i].action == a) { found = src[0].first_child + i; break; }
    int count = 1;
    if (found >= 0) {
        dst[0] = src[found];
        dst[0].parent = -1;
        int head = 0;
        while (head < count) {
            const int nch = dst[head].nch;
            if (nch > 0) {
                const int ofc = dst[head].first_child;
                for (int i = 0; i < nch; i++) {
                    dst[count + i] = src[ofc + i];
                    dst[count + i].parent = head;
                }
                dst[head].first_child = count;
                count += nch;
            }
            head++;
        }
    } else {
        MctsNode r;
        r.P = 1.1; r.Q = 0.0; r.W = 0; r.v = 0.0f; r.n = 0; r.vn = 0; r.parent = -1; r.first_child = -1;
        r.action = 0; r.nch = 0; r.flags = F_P_F64; r.nch_pending = 0;
        dst[0] = r;
    }
    d.n_nodes[t] = count;
    // the position follows the move (game.py:128; a pass only changes the side to move)
    const int color = d.root_color[t];
    u64 own = color == 1 ? d.root_p1[t] : d.root_p2[t], opp = color == 1 ? d.root_p2[t] : d.root_p1[t];
    if (a >= 0 && a < 64) place(1ULL << a, own, opp);
    d.root_p1[t] = color == 1 ? own : opp;
    d.root_p2[t] = color == 1 ? opp : own;
    d.root_color[t] = (uint8_t)(3 - color);
}

// get_move (MCTS.py:147): visits / Q per action (index 64 = pass), most visited child, first maximum.
__global__ void mcts_root_stats_kernel(MctsDev d, int exact, int32_t *visits, float *q, int8_t *best) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.T) return;
    const MctsNode *nodes = d.nodes + (size_t)t * d.cap;
    for (int i = 0; i < 65; i++) { visits[t * 65 + i] = 0; q[t * 65 + i] = 0.0f; }
    int bn = -1, ba = -2;  // -2: the root has no children (the reference raises ValueError on max() of an empty dict)
    for (int i = 0; i < nodes[0].nch; i++) {
        const MctsNode *ch = nodes + nodes[0].first_child + i;
        const int idx = ch->action < 0 ? 64 : ch->action;
        visits[t * 65 + idx] = ch->n;
        q[t * 65 + idx] = exact ? (float)ch->Q : (ch->n > 0 ? (float)((double)ch->W / kFix / (double)ch->n) : 0.0f);
        if (ch->n > bn) { bn = ch->n; ba = ch->action; }
    }
    best[t] = (int8_t)ba;
}

}  // namespace iago

using namespace iago;

struct iago_mcts {
    iago_ctx *ctx = nullptr;
    int T = 0, cap = 0, Bmax = 0;
    MctsNode *pool[2] = {nullptr, nullptr};
    int cur = 0;
    MctsDev d{};
    std::vector<void *> allocs;
    float *d_forced_v = nullptr;
    int8_t *d_forced_z = nullptr;
    int *h_counts = nullptr;  // pinned
    int32_t *d_stats = nullptr;  // visits [T][65] | q [T][65] | best [T]
    int last_exact = 1;
    long long overflows = 0;
    cudaStream_t side = nullptr;            // the lockstep rollouts of a wave run here, beside the value-net launch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_in = nullptr;
};

template <class T_>
static int dev_alloc(iago_mcts *m, T_ **p, size_t count) {
    IAGO_CUDA(cudaMalloc((void **)p, count * sizeof(T_)));
    IAGO_CUDA(cudaMemset(*p, 0, count * sizeof(T_)));
    m->allocs.push_back(*p);
    return IAGO_OK;
}

extern "C" {

int iago_mcts_create(iago_ctx *ctx, int n_trees, int max_nodes, int max_leaf_batch, uint64_t tree_id0, iago_mcts **out) {
    IAGO_REQUIRE(ctx && out, "NULL argument");
    IAGO_REQUIRE(n_trees > 0 && max_nodes >= 64 && max_leaf_batch > 0, "n_trees, max_nodes >= 64, max_leaf_batch must be positive");
    IAGO_REQUIRE((long long)n_trees * max_nodes < (1LL << 31), "n_trees * max_nodes must fit an int32 node index");
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    iago_mcts *m = new iago_mcts();
    m->ctx = ctx; m->T = n_trees; m->cap = max_nodes; m->Bmax = max_leaf_batch;
    const size_t T = n_trees, S = T * max_leaf_batch;
    int rc = 0;
#define A(ptr, cnt) if (!rc) rc = dev_alloc(m, &(ptr), (cnt))
    A(m->pool[0], T * max_nodes); A(m->pool[1], T * max_nodes);
    MctsDev &d = m->d;
    d.cap = max_nodes; d.T = n_trees;
    A(d.n_nodes, T); A(d.root_p1, T); A(d.root_p2, T); A(d.root_color, T); A(d.done, T); A(d.tree_gid, T);
    A(d.leaf_node, S); A(d.status, S); A(d.leaf_color, S); A(d.leaf_p1, S); A(d.leaf_p2, S); A(d.game_ids, S);
    A(d.z, S); A(d.v_slot, S); A(d.counts, 8);
    A(d.pol_p1, S); A(d.pol_p2, S); A(d.pol_color, S); A(d.pol_node, S); A(d.probs, S * 64);
    A(m->d_stats, T * (65 * 2 + 1));
    A(d.val_p1, S); A(d.val_p2, S); A(d.val_color, S); A(d.val_node, S); A(d.val_slot, S); A(d.vals, S);
#undef A
    if (!rc && cudaMallocHost((void **)&m->h_counts, 8 * sizeof(int)) != cudaSuccess) {
        set_error("cudaMallocHost failed");
        rc = IAGO_E_CUDA;
    }
    if (rc) {
        for (void *p : m->allocs) cudaFree(p);
        delete m;
        return rc;
    }
    std::vector<u64> gid(T);
    for (size_t i = 0; i < T; i++) gid[i] = tree_id0 + i;
    IAGO_CUDA(cudaMemcpy(d.tree_gid, gid.data(), T * 8, cudaMemcpyHostToDevice));
    d.nodes = m->pool[0];
    *out = m;
    return IAGO_OK;
}

int iago_mcts_destroy(iago_mcts *m) {
    if (!m) return IAGO_OK;
    DeviceGuard guard(m->ctx->device);
    cudaDeviceSynchronize();
    for (void *p : m->allocs) cudaFree(p);
    cudaFree(m->d_forced_v);
    cudaFree(m->d_forced_z);
    if (m->h_counts) cudaFreeHost(m->h_counts);
    if (m->side) cudaStreamDestroy(m->side);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    if (m->ev_in) cudaEventDestroy(m->ev_in);
    delete m;
    return IAGO_OK;
}

int iago_mcts_set_roots(iago_mcts *m, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int reset_tree, void *stream) {
    IAGO_REQUIRE(m && p1 && p2 && color, "NULL argument");
    DeviceGuard guard(m->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t T = m->T;
    // stage through the (idle) request buffers
    IAGO_CUDA(cudaMemcpyAsync(m->d.pol_p1, p1, T * 8, cudaMemcpyHostToDevice, s));
    IAGO_CUDA(cudaMemcpyAsync(m->d.pol_p2, p2, T * 8, cudaMemcpyHostToDevice, s));
    IAGO_CUDA(cudaMemcpyAsync(m->d.pol_color, color, T, cudaMemcpyHostToDevice, s));
    mcts_init_roots_kernel<<<(m->T + 127) / 128, 128, 0, s>>>(m->d, m->d.pol_p1, m->d.pol_p2, m->d.pol_color, reset_tree);
    IAGO_CUDA(cudaGetLastError());
    if (reset_tree) IAGO_CUDA(cudaMemsetAsync(m->d.done, 0, T * 8, s));
    IAGO_CUDA(cudaStreamSynchronize(s));
    return IAGO_OK;
}

int iago_mcts_search(iago_mcts *m, const iago_mcts_params *pp, void *stream) {
    IAGO_REQUIRE(m && pp, "NULL argument");
    IAGO_REQUIRE(pp->leaf_batch >= 1 && pp->leaf_batch <= m->Bmax, "leaf_batch out of range (1..max_leaf_batch)");
    IAGO_REQUIRE(pp->n_playouts >= 0 && pp->n_thr >= 1, "n_playouts >= 0 and n_thr >= 1 required");
    IAGO_REQUIRE(pp->precision >= 1 && pp->precision <= 3, "precision must be 1, 2 or 3");
    iago_ctx *ctx = m->ctx;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    MctsParams p{};
    p.lambda = pp->lmbda; p.c_puct = pp->c_puct; p.vloss = pp->virtual_loss; p.n_thr = pp->n_thr; p.B = pp->leaf_batch;
    p.exact = pp->leaf_batch == 1; p.cache_v = pp->cache_value != 0;
    p.need_v = pp->lmbda < 1.0; p.need_z = pp->lmbda > 0.0;
    const bool forced = pp->forced_v || pp->forced_z;
    if (p.need_v && !pp->forced_v && !trunk_slot_holds(ctx, pp->slot_value, 1)) {
        set_error("iago_mcts_search: slot %d holds no value network", pp->slot_value);
        return IAGO_E_STATE;
    }
    if (!trunk_slot_holds(ctx, pp->slot_policy, 0)) {
        set_error("iago_mcts_search: slot %d holds no policy network", pp->slot_policy);
        return IAGO_E_STATE;
    }
    if (p.need_z && !pp->forced_z && !ctx->rollout_loaded) {
        set_error("iago_mcts_search: rollout weights not loaded");
        return IAGO_E_STATE;
    }
    m->last_exact = p.exact;
    MctsDev d = m->d;
    const size_t T = m->T;
    // all trees stop at done + n_playouts; trees advance in lockstep so one host-side copy of `done` is enough
    long long done0 = 0;
    IAGO_CUDA(cudaMemcpyAsync(&done0, d.done, 8, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaStreamSynchronize(s));
    p.target = done0 + pp->n_playouts;
    d.forced_v = nullptr; d.forced_z = nullptr; d.forced_stride = 0;
    if (forced) {
        IAGO_REQUIRE(pp->forced_stride >= p.target, "forced_stride must cover every playout of the tree so far");
        cudaFree(m->d_forced_v); cudaFree(m->d_forced_z);
        m->d_forced_v = nullptr; m->d_forced_z = nullptr;
        if (pp->forced_v) {
            IAGO_CUDA(cudaMalloc((void **)&m->d_forced_v, T * pp->forced_stride * 4));
            IAGO_CUDA(cudaMemcpyAsync(m->d_forced_v, pp->forced_v, T * pp->forced_stride * 4, cudaMemcpyHostToDevice, s));
        }
        if (pp->forced_z) {
            IAGO_CUDA(cudaMalloc((void **)&m->d_forced_z, T * pp->forced_stride));
            IAGO_CUDA(cudaMemcpyAsync(m->d_forced_z, pp->forced_z, T * pp->forced_stride, cudaMemcpyHostToDevice, s));
        }
        d.forced_v = m->d_forced_v; d.forced_z = m->d_forced_z; d.forced_stride = pp->forced_stride;
    }
    const long long S = (long long)T * p.B;
    const int waves = (int)((pp->n_playouts + p.B - 1) / p.B);
    const bool run_v = p.need_v && !d.forced_v, run_z = p.need_z && !d.forced_z;
    // One wave = the kernel sequence of the file header.  Everything a wave needs lives on the device (request counts, playout
    // counters), so the waves of a search are identical launches: the first one runs directly, the second is captured into a CUDA
    // graph and replayed for the rest — a single-tree search is bound by dependent launches, not by work.  The legacy default stream
    // cannot be captured, so the search runs on the context's own stream, ordered behind the caller's stream by an event.
    auto wave = [&](cudaStream_t ws) -> int {
        const bool pipe = !p.exact && p.B >= 8 && p.B <= kPipeMaxB;   // descents of a tree as a software pipeline over kPipeWarps warps
        if (pipe) mcts_select_pipe_kernel<<<m->T, kPipeWarps * 32, 0, ws>>>(d, p, 1, 0);
        else mcts_select_kernel<<<m->T, 32, 0, ws>>>(d, p, 1);
        IAGO_CUDA(cudaGetLastError());
        int rc = trunk_launch(ctx, pp->slot_policy, 0, (const uint64_t *)d.pol_p1, (const uint64_t *)d.pol_p2, d.pol_color, S, d.probs, 1, pp->precision, ws, d.counts + 0);
        if (rc) return rc;
        const bool expand_in_select = pipe && m->T <= 8;   // a few trees: every CTA can afford to look through the request list
        if (!expand_in_select) mcts_expand_kernel<<<(unsigned)((S + 127) / 128), 128, 0, ws>>>(d);
        if (pipe) mcts_select_pipe_kernel<<<m->T, kPipeWarps * 32, 0, ws>>>(d, p, 2, expand_in_select ? 1 : 0);
        else mcts_select_kernel<<<m->T, 32, 0, ws>>>(d, p, 2);
        IAGO_CUDA(cudaGetLastError());
        // value net and rollouts read the same leaves and write different outputs: the rollouts go to a second stream
        const bool fork = run_v && run_z;
        if (fork) {
            if (!m->side) {
                IAGO_CUDA(cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking));
                IAGO_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
                IAGO_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
            }
            IAGO_CUDA(cudaEventRecord(m->ev_fork, ws));
            IAGO_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
        }
        if (run_z) {
            rc = rollout_launch_ids(ctx, (const uint64_t *)d.leaf_p1, (const uint64_t *)d.leaf_p2, d.leaf_color, S, pp->seed, 2u,
                                    (const uint64_t *)d.game_ids, d.z, nullptr, nullptr, fork ? m->side : ws);
            if (rc) return rc;
            if (fork) IAGO_CUDA(cudaEventRecord(m->ev_join, m->side));
        }
        const bool one_cta_tail = !p.exact && S <= 1024;
        if (run_v) {
            rc = trunk_launch(ctx, pp->slot_value, 1, (const uint64_t *)d.val_p1, (const uint64_t *)d.val_p2, d.val_color, S, d.vals, 0, pp->precision, ws, d.counts + 1);
            if (rc) return rc;
            mcts_scatter_value_kernel<<<(unsigned)((S + 127) / 128), 128, 0, ws>>>(d, p.cache_v);
        }
        if (fork) IAGO_CUDA(cudaStreamWaitEvent(ws, m->ev_join, 0));
        if (one_cta_tail) {
            mcts_tail_kernel<<<1, 1024, 0, ws>>>(d, p);
        } else if (p.exact) {
            mcts_backup_exact_kernel<<<(m->T + 63) / 64, 64, 0, ws>>>(d, p);
        } else {
            mcts_backup_atomic_kernel<<<(unsigned)((S + 127) / 128), 128, 0, ws>>>(d, p);
            mcts_advance_done_kernel<<<(m->T + 127) / 128, 128, 0, ws>>>(d, p);
        }
        IAGO_CUDA(cudaGetLastError());
        return IAGO_OK;
    };
    cudaStream_t ws = s;
    const bool use_graph = waves > 2 && !getenv("IAGO_MCTS_NO_GRAPH");
    if (use_graph && ctx->stream != s) {
        if (!m->ev_in) IAGO_CUDA(cudaEventCreateWithFlags(&m->ev_in, cudaEventDisableTiming));
        IAGO_CUDA(cudaEventRecord(m->ev_in, s));
        ws = ctx->stream;
        IAGO_CUDA(cudaStreamWaitEvent(ws, m->ev_in, 0));
    }
    IAGO_CUDA(cudaMemsetAsync(d.counts, 0, 8 * sizeof(int), ws));
    int w0 = 0;
    if (use_graph) {
        int rc = wave(ws);   // also sets function attributes and creates the side stream: nothing of that kind happens under capture
        if (rc) return rc;
        w0 = 1;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        if (cudaStreamBeginCapture(ws, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            rc = wave(ws);
            const cudaError_t e = cudaStreamEndCapture(ws, &graph);
            if (rc == IAGO_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                for (; w0 < waves; w0++) {
                    if (cudaGraphLaunch(exec, ws) != cudaSuccess) break;
                }
            }
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
        }
        cudaGetLastError();   // a failed capture falls back to direct launches for the remaining waves
    }
    for (; w0 < waves; w0++) {
        int rc = wave(ws);
        if (rc) return rc;
    }
    IAGO_CUDA(cudaMemcpyAsync(m->h_counts, d.counts, 8 * sizeof(int), cudaMemcpyDeviceToHost, ws));
    IAGO_CUDA(cudaStreamSynchronize(ws));
    m->overflows = m->h_counts[2];
    return IAGO_OK;
}

int iago_mcts_root_stats(iago_mcts *m, int32_t *visits, float *q, int8_t *best, void *stream) {
    IAGO_REQUIRE(m && visits && q && best, "NULL argument");
    DeviceGuard guard(m->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t T = m->T;
    int32_t *dv = m->d_stats;
    float *dq = reinterpret_cast<float *>(m->d_stats + T * 65);
    int8_t *db = reinterpret_cast<int8_t *>(m->d_stats + 2 * T * 65);
    mcts_root_stats_kernel<<<(m->T + 63) / 64, 64, 0, s>>>(m->d, m->last_exact, dv, dq, db);
    IAGO_CUDA(cudaGetLastError());
    IAGO_CUDA(cudaMemcpyAsync(visits, dv, T * 65 * 4, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaMemcpyAsync(q, dq, T * 65 * 4, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaMemcpyAsync(best, db, T, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaStreamSynchronize(s));
    return IAGO_OK;
}

int iago_mcts_advance(iago_mcts *m, const int8_t *action, const uint8_t *mask, void *stream) {
    IAGO_REQUIRE(m && action, "NULL argument");
    DeviceGuard guard(m->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t T = m->T;
    int8_t *da = reinterpret_cast<int8_t *>(m->d.pol_color);
    uint8_t *dm = m->d.val_color;
    IAGO_CUDA(cudaMemcpyAsync(da, action, T, cudaMemcpyHostToDevice, s));
    if (mask) IAGO_CUDA(cudaMemcpyAsync(dm, mask, T, cudaMemcpyHostToDevice, s));
    MctsNode *dst = m->pool[m->cur ^ 1];
    mcts_advance_kernel<<<(m->T + 31) / 32, 32, 0, s>>>(m->d, dst, da, mask ? dm : nullptr);
    IAGO_CUDA(cudaGetLastError());
    IAGO_CUDA(cudaStreamSynchronize(s));
    m->cur ^= 1;
    m->d.nodes = dst;
    return IAGO_OK;
}

int iago_mcts_get_roots(iago_mcts *m, uint64_t *p1, uint64_t *p2, uint8_t *color, int64_t *playouts_done, void *stream) {
    IAGO_REQUIRE(m && p1 && p2 && color, "NULL argument");
    DeviceGuard guard(m->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t T = m->T;
    IAGO_CUDA(cudaMemcpyAsync(p1, m->d.root_p1, T * 8, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaMemcpyAsync(p2, m->d.root_p2, T * 8, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaMemcpyAsync(color, m->d.root_color, T, cudaMemcpyDeviceToHost, s));
    if (playouts_done) IAGO_CUDA(cudaMemcpyAsync(playouts_done, m->d.done, T * 8, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaStreamSynchronize(s));
    return IAGO_OK;
}

int iago_mcts_export_tree(iago_mcts *m, int tree, int32_t capacity, int32_t *parent, int8_t *action, int32_t *n_visits,
                          double *Q, double *P, int32_t *first_child, int32_t *n_children, int32_t *count, void *stream) {
    IAGO_REQUIRE(m && parent && action && n_visits && Q && P && first_child && n_children && count, "NULL argument");
    IAGO_REQUIRE(tree >= 0 && tree < m->T, "tree index out of range");
    DeviceGuard guard(m->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    int n = 0;
    IAGO_CUDA(cudaMemcpyAsync(&n, m->d.n_nodes + tree, 4, cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaStreamSynchronize(s));
    *count = n;
    if (n > capacity) {
        set_error("iago_mcts_export_tree: tree has %d nodes, capacity %d", n, capacity);
        return IAGO_E_INVALID;
    }
    std::vector<MctsNode> h(n);
    IAGO_CUDA(cudaMemcpyAsync(h.data(), m->d.nodes + (size_t)tree * m->cap, (size_t)n * sizeof(MctsNode), cudaMemcpyDeviceToHost, s));
    IAGO_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; i++) {
        parent[i] = h[i].parent; action[i] = h[i].action; n_visits[i] = h[i].n;
        Q[i] = m->last_exact ? h[i].Q : (h[i].n > 0 ? (double)h[i].W / kFix / (double)h[i].n : 0.0);
        P[i] = h[i].P; first_child[i] = h[i].first_child; n_children[i] = h[i].nch;
    }
    return IAGO_OK;
}

int iago_mcts_overflows(iago_mcts *m, int64_t *count) {
    IAGO_REQUIRE(m && count, "NULL argument");
    *count = m->overflows;
    return IAGO_OK;
}

}  // extern "C"
