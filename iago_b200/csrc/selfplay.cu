// selfplay.cu — batched learner-vs-opponent self-play with the SL-size policy nets.
//
// Reference: src/rl_self_play.py:8-145.  Game(model1, model2)() plays colour 1 (learner, model1) against colour 2
// (opponent, model2) with `while stone_num < 64: turn(1); turn(2)`; turn (:130-145) = legal_actions, get_action
// (:111-127: SLPolicy forward -> float32 probabilities * float64 validity mask, renormalise, np.random.choice = one
// uniform per move), record the learner's swapped pre-move board + action (:134-138), place_stone; two consecutive
// passes set stone_num = 64; judge (:91-100) is from colour 1's view.  stone_num starts at 4 whatever the caller
// poked into `state` (src/train_rl.py:43-46 drops an extra un-flipped colour-2 stone on odd games), so it is a
// counter here too, not a popcount.
//
// All n games advance in lockstep: one fused trunk forward (trunk.cu) for every game's position, then one light
// kernel (one thread per game) that masks, samples / arg-maxes, flips and does the pass / terminal bookkeeping.
#include <vector>

#include "bitboard.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace iago {

enum { SEL_SAMPLE = 0, SEL_GREEDY = 1 };

struct SelfplayArgs {
    u64 *p1, *p2;
    int32_t *stone_num;
    uint8_t *pass_flg;
    int32_t *placed;     // stones placed so far = index of the next uniform
    long long n;
    int color;           // side to move in this half-step: 1 = learner, 2 = opponent
    int select;          // SEL_SAMPLE / SEL_GREEDY
    int rng_mode;
    uint32_t stream_id;
    u64 seed, game_id0;
    const double *uniforms;
    long long u_stride;
    const int8_t *forced;
    long long f_stride;
    const float *logits;  // [requests][64] from the trunk for the side to move
    const int32_t *req_index;   // [n] row of `logits` that holds game g's position, -1 = no evaluation was requested (see selfplay_request_kernel)
    // records of the learner's decisions (rl_self_play.py:134-138)
    u64 *rec_own, *rec_opp;
    int8_t *rec_action;
    int32_t *n_rec;
    int rec_cap;
    int8_t *move_log;     // [n][64] nullable
    int32_t *active;      // device counter: games with stone_num < 64 after this half-step
};

// The positions the side to move needs its net for: games that are still running and have a CHOICE (two or more legal moves).
// The reference evaluates the net at batch size 1 whenever a move exists (rl_self_play.py:111-127); for a single legal move the
// outcome does not depend on it, and a finished game or a pass never reaches get_action — so those lanes cost no trunk work here.
// Writes the compact request list (boards + colour), the game -> row map and the list length; `evaluated` accumulates the lengths.
__global__ void __launch_bounds__(128) selfplay_request_kernel(const u64 *__restrict__ p1, const u64 *__restrict__ p2, const int32_t *__restrict__ stone_num,
                                                               long long n, int color, u64 *__restrict__ req_p1, u64 *__restrict__ req_p2,
                                                               uint8_t *__restrict__ req_color, int32_t *__restrict__ req_index, int32_t *__restrict__ count,
                                                               unsigned long long *__restrict__ evaluated) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool want = false;
    u64 b1 = 0, b2 = 0;
    if (g < n && stone_num[g] < 64) {
        b1 = p1[g]; b2 = p2[g];
        const u64 legal = color == 1 ? legal_moves(b1, b2) : legal_moves(b2, b1);
        want = (legal & (legal - 1)) != 0;   // at least two bits set
    }
    // one atomic per warp: rows of a warp's games are consecutive
    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, want);
    int base = 0;
    if ((threadIdx.x & 31) == 0 && ballot) {
        base = atomicAdd(count, __popc(ballot));
        atomicAdd(evaluated, (unsigned long long)__popc(ballot));
    }
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (g < n) {
        int row = -1;
        if (want) {
            row = base + __popc(ballot & ((1u << (threadIdx.x & 31)) - 1u));
            req_p1[row] = b1;
            req_p2[row] = b2;
            req_color[row] = (uint8_t)color;
        }
        req_index[g] = row;
    }
}

__global__ void __launch_bounds__(128) selfplay_turn_kernel(SelfplayArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    int stone_num = a.stone_num[g];
    if (stone_num >= 64) return;  // finished games take no more turns (the reference loop has exited)
    const u64 b1 = a.p1[g], b2 = a.p2[g];
    u64 own = a.color == 1 ? b1 : b2, opp = a.color == 1 ? b2 : b1;
    const u64 legal = legal_moves(own, opp);
    if (legal) {
        const int placed = a.placed[g];
        int k = -1;
        if (a.rng_mode == IAGO_RNG_FORCED) {
            k = placed < a.f_stride ? a.forced[g * a.f_stride + placed] : -1;
        } else {
            const int ri = a.req_index[g];
            const float *lg = a.logits + (long long)(ri < 0 ? 0 : ri) * 64;
            if (ri < 0) {
                // one legal move: the choice does not depend on the net (arg-max over one cell; np.random.choice over one non-zero
                // probability), so no forward was requested for this game.  The sampler still consumes its uniform: `placed` advances.
                k = __ffsll((long long)legal) - 1;
            } else if (a.select == SEL_GREEDY) {
                // arg-max of the logits over legal moves, lowest index on ties (BASELINE configs[2])
                float best = -3.0e38f;
                for (u64 m = legal; m; m &= m - 1) {
                    const int c = __ffsll((long long)m) - 1;
                    const float v = lg[c];
                    if (v > best) { best = v; k = c; }
                }
            } else {
                // float32 softmax over all 64 cells, then float64 masked renormalisation and inverse cdf
                float mx = -3.0e38f;
                for (int i = 0; i < 64; i++) mx = fmaxf(mx, lg[i]);
                float sum = 0.0f;
                for (int i = 0; i < 64; i++) sum += expf(lg[i] - mx);
                double total = 0.0;
                for (u64 m = legal; m; m &= m - 1) total += (double)(expf(lg[__ffsll((long long)m) - 1] - mx) / sum);
                double u;
                if (a.rng_mode == IAGO_RNG_UNIFORMS)
                    u = placed < a.u_stride ? a.uniforms[g * a.u_stride + placed] : 0.5;   // (a replay stream shorter than the game: never read past it)
                else
                    u = (double)philox_m53(a.seed, a.game_id0 + (u64)g, (uint32_t)placed, a.stream_id) * (1.0 / 9007199254740992.0);
                const double t = u * total;
                double cum = 0.0;
                for (u64 m = legal; m; m &= m - 1) {
                    k = __ffsll((long long)m) - 1;
                    cum += (double)(expf(lg[k] - mx) / sum);
                    if (cum > t) break;
                }
            }
        }
        if (k < 0 || k > 63) {
            stone_num = 64;  // replay stream exhausted
        } else {
            if (a.color == 1) {
                const int r = a.n_rec[g];
                if (r < a.rec_cap) {
                    a.rec_own[g * a.rec_cap + r] = own;
                    a.rec_opp[g * a.rec_cap + r] = opp;
                    a.rec_action[g * a.rec_cap + r] = (int8_t)k;
                }
                a.n_rec[g] = r + 1;
            }
            place(1ULL << k, own, opp);
            if (a.move_log) a.move_log[g * 64 + placed] = (int8_t)k;
            a.placed[g] = placed + 1;
            a.pass_flg[g] = 0;
            stone_num += 1;
            a.p1[g] = a.color == 1 ? own : opp;
            a.p2[g] = a.color == 1 ? opp : own;
        }
    } else {
        if (a.pass_flg[g]) stone_num = 64;
        a.pass_flg[g] = 1;
    }
    a.stone_num[g] = stone_num;
    if (a.color == 2 && stone_num < 64) atomicAdd(a.active, 1);
}

__global__ void selfplay_init_kernel(u64 *p1, u64 *p2, const u64 *in1, const u64 *in2, int32_t *stone_num, uint8_t *pass_flg,
                                     int32_t *placed, int32_t *n_rec, int8_t *move_log, uint8_t *c1, uint8_t *c2, long long n) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    p1[g] = in1 ? in1[g] : ((1ULL << 35) | (1ULL << 28));  // game.py:26-30 / rl_self_play.py:12-16
    p2[g] = in2 ? in2[g] : ((1ULL << 27) | (1ULL << 36));
    stone_num[g] = 4;  // rl_self_play.py:20 — a counter, not a popcount
    pass_flg[g] = 0;
    placed[g] = 0;
    n_rec[g] = 0;
    c1[g] = 1;
    c2[g] = 2;
    if (move_log)
        for (int i = 0; i < 64; i++) move_log[g * 64 + i] = -1;
}

__global__ void selfplay_judge_kernel(const u64 *p1, const u64 *p2, int8_t *result, long long n) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const int a = __popcll(p1[g]), b = __popcll(p2[g]);
    result[g] = (int8_t)((a > b) - (a < b));
}

// ---------------------------------------------------------------- rl_env.GameEnv.step (rl_env.py:41-74) for n environments
struct EnvArgs {
    u64 *p1, *p2;
    int32_t *stone_num;
    uint8_t *pass_flg, *done;
    int32_t *draws;          // uniforms consumed so far per environment (index of the next one)
    const int8_t *action;    // learner's action (phase 1)
    int8_t *opp_action;      // opponent's answer, -1 = pass (phase 2, nullable)
    const float *probs;      // [n][64] SLPolicy output for colour 2 to move (phase 2)
    long long n;
    int rng_mode;
    uint32_t stream_id;
    u64 seed, game_id0;
    const double *uniforms;
    long long u_stride;
    int32_t *errors;         // [0] rejection loops that hit the reference's recursion limit
};

__device__ __forceinline__ double env_uniform(const EnvArgs &a, long long g, int k) {
    if (a.rng_mode == IAGO_RNG_UNIFORMS) return k < a.u_stride ? a.uniforms[g * a.u_stride + k] : 0.5;
    return (double)philox_m53(a.seed, a.game_id0 + (u64)g, (uint32_t)k, a.stream_id) * (1.0 / 9007199254740992.0);
}

// Phase 1: the learner (colour 1) plays `action`; an illegal action is replaced by positions[floor(u * len)] (the reference
// calls Python's random.choice there, rl_env.py:46-48); no legal move = pass, a second consecutive pass ends the game.
__global__ void __launch_bounds__(128) env_learner_kernel(EnvArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    a.done[g] = 0;
    u64 own = a.p1[g], opp = a.p2[g];
    const u64 legal = legal_moves(own, opp);
    if (legal) {
        int k = a.action[g];
        if (k < 0 || k > 63 || !((legal >> k) & 1)) {
            const int cnt = __popcll(legal);
            const int d = a.draws[g];
            int idx = (int)(env_uniform(a, g, d) * (double)cnt);
            idx = idx < cnt ? idx : cnt - 1;
            a.draws[g] = d + 1;
            u64 m = legal;
            for (int j = 0; j < idx; j++) m &= m - 1;
            k = __ffsll((long long)m) - 1;
        }
        place(1ULL << k, own, opp);
        a.p1[g] = own;
        a.p2[g] = opp;
        a.stone_num[g] += 1;
        a.pass_flg[g] = 0;
    } else {
        if (a.pass_flg[g]) a.done[g] = 1;
        a.pass_flg[g] = 1;
    }
}

// get_position (rl_env.py:152-172, self_play.py:8-30): p = out - min(out) over ALL 64 cells in float32 (numpy's pairwise sum
// order), np.random.choice = float64 cdf + searchsorted(right), re-drawn with the next uniform until the cell is legal.
// Returns the cell or -1 when the reference's recursion limit (10,000, rl_env.py:8) is hit; d = index of the next uniform.
__device__ __forceinline__ int unmasked_draw(const EnvArgs &a, long long g, const float *pr, u64 legal, int &d) {
    float mn = pr[0];
    for (int i = 1; i < 64; i++) mn = fminf(mn, pr[i]);
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = __fsub_rn(pr[j], mn);
    for (int i = 8; i < 64; i += 8)
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], __fsub_rn(pr[i + j], mn));
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                                __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    double total = 0.0;
    for (int i = 0; i < 64; i++) total = __dadd_rn(total, (double)__fdiv_rn(__fsub_rn(pr[i], mn), sum));
    for (int attempt = 0; attempt < 10000; attempt++) {
        const double u = env_uniform(a, g, d++);
        double cum = 0.0;
        int idx = 64;
        for (int i = 0; i < 64; i++) {
            cum = __dadd_rn(cum, (double)__fdiv_rn(__fsub_rn(pr[i], mn), sum));
            if (__ddiv_rn(cum, total) > u) { idx = i; break; }
        }
        if (idx < 64 && ((legal >> idx) & 1)) return idx;
    }
    return -1;
}

// Phase 2: the opponent (colour 2) answers.
__global__ void __launch_bounds__(128) env_opponent_kernel(EnvArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    u64 own = a.p2[g], opp = a.p1[g];
    const u64 legal = legal_moves(own, opp);
    int chosen = -1;
    if (legal) {
        int d = a.draws[g];
        chosen = unmasked_draw(a, g, a.probs + g * 64, legal, d);
        a.draws[g] = d;
        if (chosen < 0) {
            atomicAdd(a.errors, 1);
            a.done[g] = 1;
        } else {
            place(1ULL << chosen, own, opp);
            a.p2[g] = own;
            a.p1[g] = opp;
            a.stone_num[g] += 1;
            a.pass_flg[g] = 0;
        }
    } else {
        if (a.pass_flg[g]) a.done[g] = 1;
        a.pass_flg[g] = 1;
    }
    if (a.stone_num[g] >= 64) a.done[g] = 1;
    if (a.opp_action) a.opp_action[g] = (int8_t)chosen;
}

// The sampler alone: p1 = mover's stones, p2 = the other side's; action -1 = no legal move, -2 = recursion limit.
__global__ void __launch_bounds__(128) sample_unmasked_kernel(EnvArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    const u64 legal = legal_moves(a.p1[g], a.p2[g]);
    int chosen = -1;
    if (legal) {
        int d = a.draws[g];
        chosen = unmasked_draw(a, g, a.probs + g * 64, legal, d);
        a.draws[g] = d;
        if (chosen < 0) { atomicAdd(a.errors, 1); chosen = -2; }
    }
    a.opp_action[g] = (int8_t)chosen;
}

// get_action_auto / get_action (game.py:101-108, rl_self_play.py:111-127) for given probabilities: p = prob (float32) * validity
// mask (float64), renormalised, np.random.choice with ONE uniform.  action -1 = no legal move.
__global__ void __launch_bounds__(128) sample_masked_kernel(EnvArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    const u64 legal = legal_moves(a.p1[g], a.p2[g]);
    int chosen = -1;
    if (legal) {
        const float *pr = a.probs + g * 64;
        double total = 0.0;
        for (u64 m = legal; m; m &= m - 1) total = __dadd_rn(total, (double)pr[__ffsll((long long)m) - 1]);
        const int d = a.draws[g];
        const double t = __dmul_rn(env_uniform(a, g, d), total);
        a.draws[g] = d + 1;
        double cum = 0.0;
        for (u64 m = legal; m; m &= m - 1) {
            chosen = __ffsll((long long)m) - 1;
            cum = __dadd_rn(cum, (double)pr[chosen]);
            if (cum > t) break;
        }
    }
    a.opp_action[g] = (int8_t)chosen;
}

struct SelfplayWs {
    long long cap = 0;
    int32_t *stone_num = nullptr, *placed = nullptr, *active = nullptr;
    uint8_t *pass_flg = nullptr, *c1 = nullptr, *c2 = nullptr;
    float *logits = nullptr;
    int32_t *h_active = nullptr;  // pinned: [0] games still running, [1..2] = evaluated positions (64-bit)
    u64 *req_p1 = nullptr, *req_p2 = nullptr;
    uint8_t *req_color = nullptr;
    int32_t *req_index = nullptr, *req_count = nullptr;
    unsigned long long *evaluated = nullptr;
    int32_t *env_err = nullptr;
};

static int ws_ensure(iago_ctx *ctx, long long n, SelfplayWs **out) {
    if (!ctx->selfplay) ctx->selfplay = new SelfplayWs();
    SelfplayWs *w = static_cast<SelfplayWs *>(ctx->selfplay);
    if (w->cap < n) {
        cudaFree(w->stone_num); cudaFree(w->placed); cudaFree(w->pass_flg); cudaFree(w->c1); cudaFree(w->c2); cudaFree(w->logits);
        cudaFree(w->req_p1); cudaFree(w->req_p2); cudaFree(w->req_color); cudaFree(w->req_index);
        IAGO_CUDA(cudaMalloc(&w->req_p1, n * 8));
        IAGO_CUDA(cudaMalloc(&w->req_p2, n * 8));
        IAGO_CUDA(cudaMalloc(&w->req_color, n));
        IAGO_CUDA(cudaMalloc(&w->req_index, n * 4));
        IAGO_CUDA(cudaMalloc(&w->stone_num, n * 4));
        IAGO_CUDA(cudaMalloc(&w->placed, n * 4));
        IAGO_CUDA(cudaMalloc(&w->pass_flg, n));
        IAGO_CUDA(cudaMalloc(&w->c1, n));
        IAGO_CUDA(cudaMalloc(&w->c2, n));
        IAGO_CUDA(cudaMalloc(&w->logits, n * 64 * 4));
        w->cap = n;
    }
    if (!w->active) {
        IAGO_CUDA(cudaMalloc(&w->active, 4));
        IAGO_CUDA(cudaMalloc(&w->req_count, 4));
        IAGO_CUDA(cudaMalloc(&w->evaluated, 8));
        IAGO_CUDA(cudaMallocHost(&w->h_active, 16));
    }
    *out = w;
    return IAGO_OK;
}

void selfplay_destroy(iago_ctx *ctx) {
    if (!ctx->selfplay) return;
    SelfplayWs *w = static_cast<SelfplayWs *>(ctx->selfplay);
    cudaFree(w->stone_num); cudaFree(w->placed); cudaFree(w->pass_flg); cudaFree(w->c1); cudaFree(w->c2); cudaFree(w->logits);
    cudaFree(w->active);
    cudaFree(w->req_p1); cudaFree(w->req_p2); cudaFree(w->req_color); cudaFree(w->req_index); cudaFree(w->req_count); cudaFree(w->evaluated);
    cudaFree(w->env_err);
    if (w->h_active) cudaFreeHost(w->h_active);
    delete w;
    ctx->selfplay = nullptr;
}

}  // namespace iago

using namespace iago;

extern "C" {

int iago_selfplay(iago_ctx *ctx, int slot_learner, int slot_opponent, int64_t n, const uint64_t *init_p1,
                  const uint64_t *init_p2, int select, int precision, const iago_rng *rng, uint64_t *final_p1,
                  uint64_t *final_p2, int8_t *result, uint64_t *rec_own, uint64_t *rec_opp, int8_t *rec_action,
                  int32_t *n_rec, int rec_cap, int8_t *move_log, int64_t *stats, void *stream) {
    IAGO_REQUIRE(ctx && rng && final_p1 && final_p2 && result && rec_own && rec_opp && rec_action && n_rec, "NULL argument");
    IAGO_REQUIRE(n >= 0 && rec_cap > 0, "n < 0 or rec_cap <= 0");
    IAGO_REQUIRE(select == SEL_SAMPLE || select == SEL_GREEDY, "select must be 0 (sample) or 1 (greedy)");
    IAGO_REQUIRE(rng->mode >= IAGO_RNG_PHILOX && rng->mode <= IAGO_RNG_FORCED, "rng.mode");
    if (rng->mode == IAGO_RNG_UNIFORMS) IAGO_REQUIRE(rng->uniforms && rng->u_stride > 0, "rng.uniforms / u_stride");
    if (rng->mode == IAGO_RNG_FORCED) IAGO_REQUIRE(rng->forced && rng->f_stride > 0, "rng.forced / f_stride");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    SelfplayWs *w = nullptr;
    int rc = ws_ensure(ctx, n, &w);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + 127) / 128);
    u64 *p1 = (u64 *)final_p1, *p2 = (u64 *)final_p2;  // the game state lives in the caller's output arrays
    selfplay_init_kernel<<<grid, 128, 0, s>>>(p1, p2, (const u64 *)init_p1, (const u64 *)init_p2, w->stone_num, w->pass_flg,
                                               w->placed, n_rec, move_log, w->c1, w->c2, n);
    IAGO_CUDA(cudaGetLastError());
    SelfplayArgs a{p1, p2, w->stone_num, w->pass_flg, w->placed, n, 1, select, rng->mode, rng->stream_id, rng->seed,
                   rng->game_id0, rng->uniforms, rng->u_stride, rng->forced, rng->f_stride, w->logits, w->req_index, (u64 *)rec_own,
                   (u64 *)rec_opp, rec_action, n_rec, rec_cap, move_log, w->active};
    const bool need_net = rng->mode != IAGO_RNG_FORCED;
    long long pairs = 0, forwards = 0;
    IAGO_CUDA(cudaMemsetAsync(w->evaluated, 0, 8, s));
    for (;;) {
        IAGO_CUDA(cudaMemsetAsync(w->active, 0, 4, s));
        for (int color = 1; color <= 2; color++) {
            if (need_net) {
                // the trunk runs on the games that are alive and have a choice (a device-built list with a device-side length)
                IAGO_CUDA(cudaMemsetAsync(w->req_count, 0, 4, s));
                selfplay_request_kernel<<<grid, 128, 0, s>>>(p1, p2, w->stone_num, n, color, w->req_p1, w->req_p2, w->req_color, w->req_index,
                                                            w->req_count, w->evaluated);
                IAGO_CUDA(cudaGetLastError());
                rc = trunk_launch(ctx, color == 1 ? slot_learner : slot_opponent, 0, (const uint64_t *)w->req_p1, (const uint64_t *)w->req_p2,
                                  w->req_color, n, w->logits, 0, precision, s, w->req_count);
                if (rc) return rc;
                forwards++;
            }
            a.color = color;
            selfplay_turn_kernel<<<grid, 128, 0, s>>>(a);
            IAGO_CUDA(cudaGetLastError());
        }
        pairs++;
        IAGO_CUDA(cudaMemcpyAsync(w->h_active, w->active, 4, cudaMemcpyDeviceToHost, s));
        IAGO_CUDA(cudaStreamSynchronize(s));
        if (*w->h_active == 0) break;
        if (pairs > 70) {  // every pair of turns places a stone or ends the game; 60 empties bound the loop
            set_error("iago_selfplay: games did not terminate after %lld turn pairs", pairs);
            return IAGO_E_STATE;
        }
    }
    selfplay_judge_kernel<<<grid, 128, 0, s>>>(p1, p2, result, n);
    IAGO_CUDA(cudaGetLastError());
    if (stats) {
        stats[0] = pairs;
        stats[1] = forwards;
        IAGO_CUDA(cudaMemcpyAsync(w->h_active + 2, w->evaluated, 8, cudaMemcpyDeviceToHost, s));
        IAGO_CUDA(cudaStreamSynchronize(s));
        stats[2] = need_net ? (int64_t)*reinterpret_cast<unsigned long long *>(w->h_active + 2) : 0;   // positions the nets evaluated
    }
    return IAGO_OK;
}

int iago_env_step(iago_ctx *ctx, int slot_opponent, int precision, int64_t n, uint64_t *p1, uint64_t *p2, int32_t *stone_num,
                  uint8_t *pass_flg, const int8_t *action, const iago_rng *rng, int32_t *draws, uint8_t *done,
                  int8_t *opp_action, int32_t *errors_host, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && stone_num && pass_flg && action && rng && draws && done, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    IAGO_REQUIRE(rng->mode == IAGO_RNG_PHILOX || rng->mode == IAGO_RNG_UNIFORMS, "rng.mode must be PHILOX or UNIFORMS");
    if (rng->mode == IAGO_RNG_UNIFORMS) IAGO_REQUIRE(rng->uniforms && rng->u_stride > 0, "rng.uniforms / u_stride");
    if (errors_host) *errors_host = 0;
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    SelfplayWs *w = nullptr;
    int rc = ws_ensure(ctx, n, &w);
    if (rc) return rc;
    if (!w->env_err) IAGO_CUDA(cudaMalloc(&w->env_err, 4));
    cudaStream_t s = (cudaStream_t)stream;
    IAGO_CUDA(cudaMemsetAsync(w->env_err, 0, 4, s));
    const unsigned grid = (unsigned)((n + 127) / 128);
    EnvArgs a{(u64 *)p1, (u64 *)p2, stone_num, pass_flg, done, draws, action, opp_action, w->logits, n, rng->mode,
              rng->stream_id, rng->seed, rng->game_id0, rng->uniforms, rng->u_stride, w->env_err};
    env_learner_kernel<<<grid, 128, 0, s>>>(a);
    IAGO_CUDA(cudaGetLastError());
    IAGO_CUDA(cudaMemsetAsync(w->c2, 2, (size_t)n, s));
    rc = trunk_launch(ctx, slot_opponent, 0, p1, p2, w->c2, n, w->logits, 1, precision, s);  // probabilities, colour 2 to move
    if (rc) return rc;
    env_opponent_kernel<<<grid, 128, 0, s>>>(a);
    IAGO_CUDA(cudaGetLastError());
    if (errors_host) {
        IAGO_CUDA(cudaMemcpyAsync(w->h_active, w->env_err, 4, cudaMemcpyDeviceToHost, s));
        IAGO_CUDA(cudaStreamSynchronize(s));
        *errors_host = *w->h_active;
    }
    return IAGO_OK;
}

int iago_sample_unmasked(iago_ctx *ctx, const float *probs, const uint64_t *own, const uint64_t *opp, int64_t n,
                         const iago_rng *rng, int32_t *draws, int8_t *action, int32_t *errors_host, void *stream) {
    IAGO_REQUIRE(ctx && probs && own && opp && rng && draws && action, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    IAGO_REQUIRE(rng->mode == IAGO_RNG_PHILOX || rng->mode == IAGO_RNG_UNIFORMS, "rng.mode must be PHILOX or UNIFORMS");
    if (rng->mode == IAGO_RNG_UNIFORMS) IAGO_REQUIRE(rng->uniforms && rng->u_stride > 0, "rng.uniforms / u_stride");
    if (errors_host) *errors_host = 0;
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    SelfplayWs *w = nullptr;
    int rc = ws_ensure(ctx, 1, &w);
    if (rc) return rc;
    if (!w->env_err) IAGO_CUDA(cudaMalloc(&w->env_err, 4));
    cudaStream_t s = (cudaStream_t)stream;
    IAGO_CUDA(cudaMemsetAsync(w->env_err, 0, 4, s));
    EnvArgs a{(u64 *)own, (u64 *)opp, nullptr, nullptr, nullptr, draws, nullptr, action, probs, n, rng->mode,
              rng->stream_id, rng->seed, rng->game_id0, rng->uniforms, rng->u_stride, w->env_err};
    sample_unmasked_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a);
    IAGO_CUDA(cudaGetLastError());
    if (errors_host) {
        IAGO_CUDA(cudaMemcpyAsync(w->h_active, w->env_err, 4, cudaMemcpyDeviceToHost, s));
        IAGO_CUDA(cudaStreamSynchronize(s));
        *errors_host = *w->h_active;
    }
    return IAGO_OK;
}

int iago_sample_masked(iago_ctx *ctx, const float *probs, const uint64_t *own, const uint64_t *opp, int64_t n,
                       const iago_rng *rng, int32_t *draws, int8_t *action, void *stream) {
    IAGO_REQUIRE(ctx && probs && own && opp && rng && draws && action, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    IAGO_REQUIRE(rng->mode == IAGO_RNG_PHILOX || rng->mode == IAGO_RNG_UNIFORMS, "rng.mode must be PHILOX or UNIFORMS");
    if (rng->mode == IAGO_RNG_UNIFORMS) IAGO_REQUIRE(rng->uniforms && rng->u_stride > 0, "rng.uniforms / u_stride");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    EnvArgs a{(u64 *)own, (u64 *)opp, nullptr, nullptr, nullptr, draws, nullptr, action, probs, n, rng->mode,
              rng->stream_id, rng->seed, rng->game_id0, rng->uniforms, rng->u_stride, nullptr};
    sample_masked_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(a);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

}  // extern "C"

namespace iago {
// ---------------------------------------------------------------- REINFORCE set plumbing on the device (src/train_rl.py:41-53)
// Openings of a set: game i starts from the standard position; every odd game gets an extra colour-2 stone on one of
// (2,4), (3,5), (4,2), (5,3) WITHOUT flipping ("switch head and tail", src/train_rl.py:43-46).  The reference picks the cell with
// Python's random.choice; here it is word 0 of the Philox block (seed, global game id, 0, stream 6) & 3.
__global__ void __launch_bounds__(256) openings_kernel(long long n, u64 seed, u64 game_id0, u64 *__restrict__ p1, u64 *__restrict__ p2) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    u64 b1 = (1ULL << 28) | (1ULL << 35), b2 = (1ULL << 27) | (1ULL << 36);   // game.py:26-30: (4,3),(3,4) = 1; (3,3),(4,4) = 2
    if (g & 1) {
        uint32_t o[4];
        philox_block(seed, game_id0 + (u64)g, 0, 6u, o);
        const int cells[4] = {2 * 8 + 4, 3 * 8 + 5, 4 * 8 + 2, 5 * 8 + 3};
        b2 |= 1ULL << cells[o[0] & 3u];
    }
    p1[g] = b1;
    p2[g] = b2;
}

// The learner's recorded decisions of n games, [n][rec_cap] with n_rec[g] valid entries, flattened game by game (the order
// src/train_rl.py:47-50 stacks them in) with the game's result as the reward of each of its records.  One CTA: per-thread counts,
// block scan, scatter.  out_count[0] = records, out_count[1] = games won by the learner, out_count[2] = largest n_rec (> rec_cap
// means records were lost: the caller raises).
__global__ void __launch_bounds__(1024) compact_records_kernel(long long n, int rec_cap, const u64 *__restrict__ rec_own, const u64 *__restrict__ rec_opp,
                                                               const int8_t *__restrict__ rec_action, const int32_t *__restrict__ n_rec,
                                                               const int8_t *__restrict__ result, u64 *__restrict__ out_own, u64 *__restrict__ out_opp,
                                                               int8_t *__restrict__ out_action, float *__restrict__ out_reward, int32_t *__restrict__ out_count) {
    __shared__ int part[1024];
    __shared__ int s_wins, s_max;
    const int t = threadIdx.x;
    const long long per = (n + 1023) / 1024, lo = (long long)t * per, hi = lo + per < n ? lo + per : n;
    if (t == 0) { s_wins = 0; s_max = 0; }
    __syncthreads();
    int cnt = 0, wins = 0, mx = 0;
    for (long long g = lo; g < hi; g++) {
        const int k = n_rec[g];
        cnt += k < rec_cap ? k : rec_cap;
        wins += result[g] == 1;
        mx = k > mx ? k : mx;
    }
    part[t] = cnt;
    if (wins) atomicAdd(&s_wins, wins);
    if (mx) atomicMax(&s_max, mx);
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {   // inclusive scan
        const int v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long o = part[t] - cnt;
    for (long long g = lo; g < hi; g++) {
        const int k = n_rec[g] < rec_cap ? n_rec[g] : rec_cap;
        const float r = (float)result[g];
        for (int i = 0; i < k; i++, o++) {
            out_own[o] = rec_own[g * rec_cap + i];
            out_opp[o] = rec_opp[g * rec_cap + i];
            out_action[o] = rec_action[g * rec_cap + i];
            out_reward[o] = r;
        }
    }
    if (t == 1023) {
        out_count[0] = part[1023];
        out_count[1] = s_wins;
        out_count[2] = s_max;
    }
}

}  // namespace iago

using namespace iago;

extern "C" {

int iago_reinforce_openings(iago_ctx *ctx, int64_t n, uint64_t seed, uint64_t game_id0, uint64_t *p1, uint64_t *p2, void *stream) {
    IAGO_REQUIRE(ctx && p1 && p2 && n >= 0, "NULL argument");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    openings_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, seed, game_id0, (u64 *)p1, (u64 *)p2);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_reinforce_compact(iago_ctx *ctx, int64_t n, int rec_cap, const uint64_t *rec_own, const uint64_t *rec_opp, const int8_t *rec_action,
                           const int32_t *n_rec, const int8_t *result, uint64_t *out_own, uint64_t *out_opp, int8_t *out_action,
                           float *out_reward, int32_t *out_count, void *stream) {
    IAGO_REQUIRE(ctx && rec_own && rec_opp && rec_action && n_rec && result && out_own && out_opp && out_action && out_reward && out_count,
                 "NULL argument");
    IAGO_REQUIRE(n >= 0 && rec_cap > 0, "n / rec_cap");
    DeviceGuard guard(ctx->device);
    compact_records_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, rec_cap, (const u64 *)rec_own, (const u64 *)rec_opp, rec_action, n_rec, result,
                                                                (u64 *)out_own, (u64 *)out_opp, out_action, out_reward, out_count);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

}  // extern "C"
