// valuegen.cu — batched value-network training-data generation.
//
// Reference: value_self_play.SelfPlay(stop_num)() (value_self_play.py:12-59) driven by gen_value_data.py:12-19.
//   phase 0  `while stone_num < stop_num: turn(cl, model0); cl = order(cl)`   — the SL policy plays both colours
//   phase 1  color = cl; the board is recorded from the mover's view (value_self_play.py:39-44: mover's stones 2, the
//            other side's 1); valid_pos(cl) empty -> return (state, -1); else ONE uniformly random legal move
//            (random.choice), pass_flg = False, stone_num += 1, cl = order(cl)                      (:46-53)
//   phase 2  `while stone_num < 64: turn(cl, model1); cl = order(cl)`          — the RL policy plays the game out
//   result   judge(color) from the recorded mover's view                                            (:59, :120-128)
// turn (:152-162): no legal move = pass, two consecutive passes set stone_num = 64.  get_position (:131-149): softmax over
// ALL 64 outputs of the net (the old SLPolicyNet emitted logits), np.random.choice = float64 cdf + searchsorted(right) with
// one uniform; an illegal cell is replaced by random.choice(positions) — here positions[floor(u' * len)] with the next uniform
// of the game's stream, like iago_env_step.  The reference's softmax does not subtract the maximum (it overflows to NaN and
// raises for logits above 88.7); this one does, which is the same distribution wherever the reference's is finite.
//
// The side to move is the same in every game of the batch (cl flips once per turn, the random move included), but which
// net a game needs differs, so each turn builds two request lists on the device (games in phase 0 / phase 2 that have a legal
// move), runs the fused trunk once per list with a device-side count, and then one thread per game samples, flips and does the
// bookkeeping.  The host reads one counter every four turns.
#include <vector>

#include "bitboard.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace iago {

enum { VG_SL = 0, VG_RANDOM = 1, VG_RL = 2, VG_DONE = 3 };

struct ValueGenWs {
    long long cap = 0;
    u64 *req_p1[2] = {}, *req_p2[2] = {};
    uint8_t *req_color[2] = {};
    float *logits[2] = {};
    int32_t *req_index = nullptr, *stone_num = nullptr;
    uint8_t *pass_flg = nullptr, *phase = nullptr;
    int32_t *counts = nullptr;      // [0], [1]: request list lengths; [2]: games not done
    int32_t *h_counts = nullptr;    // pinned
};

struct ValueGenArgs {
    u64 *p1, *p2;
    const int32_t *stop_num;
    int32_t *stone_num, *draws, *req_index, *counts;
    uint8_t *pass_flg, *phase;
    u64 *req_p1[2], *req_p2[2];
    uint8_t *req_color[2];
    const float *logits[2];
    long long n;
    int color;
    int rng_mode;
    uint32_t stream_id;
    u64 seed, game_id0;
    const double *uniforms;
    long long u_stride;
    u64 *rec_own, *rec_opp;
    uint8_t *rec_color;
    int8_t *rec_action, *result;
};

__device__ __forceinline__ double vg_uniform(const ValueGenArgs &a, long long g, int k) {
    if (a.rng_mode == IAGO_RNG_UNIFORMS) return k < a.u_stride ? a.uniforms[g * a.u_stride + k] : 0.5;
    return (double)philox_m53(a.seed, a.game_id0 + (u64)g, (uint32_t)k, a.stream_id) * (1.0 / 9007199254740992.0);
}

__global__ void valuegen_init_kernel(ValueGenArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    a.p1[g] = (1ULL << 35) | (1ULL << 28);   // value_self_play.py:16-20
    a.p2[g] = (1ULL << 27) | (1ULL << 36);
    a.stone_num[g] = 4;
    a.pass_flg[g] = 0;
    a.phase[g] = VG_SL;
    a.draws[g] = 0;
    a.rec_own[g] = 0;
    a.rec_opp[g] = 0;
    a.rec_color[g] = 0;
    a.rec_action[g] = -1;
    a.result[g] = -1;
}

// Loop tests of this turn (value_self_play.py:35, :55) and the request lists for the two nets.
__global__ void __launch_bounds__(128) valuegen_prepare_kernel(ValueGenArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    int phase = a.phase[g];
    if (phase == VG_DONE) return;
    const int stone_num = a.stone_num[g];
    if (phase == VG_SL && stone_num >= a.stop_num[g]) phase = VG_RANDOM;
    if (phase == VG_RL && stone_num >= 64) phase = VG_DONE;
    a.phase[g] = (uint8_t)phase;
    if (phase == VG_DONE) {
        // judge(color) for the recorded mover (value_self_play.py:120-128)
        const int c = a.rec_color[g];
        const int me = __popcll(c == 1 ? a.p1[g] : a.p2[g]), op = __popcll(c == 1 ? a.p2[g] : a.p1[g]);
        a.result[g] = (int8_t)((me > op) - (me < op));
        return;
    }
    atomicAdd(a.counts + 2, 1);
    if (phase == VG_RANDOM) return;
    const u64 b1 = a.p1[g], b2 = a.p2[g];
    const u64 own = a.color == 1 ? b1 : b2, opp = a.color == 1 ? b2 : b1;
    if (!legal_moves(own, opp)) return;
    const int list = phase == VG_SL ? 0 : 1;
    const int i = atomicAdd(a.counts + list, 1);
    a.req_p1[list][i] = b1;
    a.req_p2[list][i] = b2;
    a.req_color[list][i] = (uint8_t)a.color;
    a.req_index[g] = i;
}

// get_position's draw (value_self_play.py:131-149): float32 softmax over all 64 cells (numpy's pairwise sum order for the
// denominator), float64 cdf, searchsorted(right).
__device__ __forceinline__ int softmax_draw(const float *lg, double u) {
    float mx = lg[0];
    for (int i = 1; i < 64; i++) mx = fmaxf(mx, lg[i]);
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = expf(__fsub_rn(lg[j], mx));
    for (int i = 8; i < 64; i += 8)
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __fadd_rn(r[j], expf(__fsub_rn(lg[i + j], mx)));
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                                __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    double total = 0.0;
    for (int i = 0; i < 64; i++) total = __dadd_rn(total, (double)__fdiv_rn(expf(__fsub_rn(lg[i], mx)), sum));
    double cum = 0.0;
    for (int i = 0; i < 64; i++) {
        cum = __dadd_rn(cum, (double)__fdiv_rn(expf(__fsub_rn(lg[i], mx)), sum));
        if (__ddiv_rn(cum, total) > u) return i;
    }
    return 64;
}

__device__ __forceinline__ int nth_legal(u64 legal, double u) {
    const int cnt = __popcll(legal);
    int idx = (int)(u * (double)cnt);
    idx = idx < cnt ? idx : cnt - 1;
    u64 m = legal;
    for (int j = 0; j < idx; j++) m &= m - 1;
    return __ffsll((long long)m) - 1;
}

__global__ void __launch_bounds__(128) valuegen_turn_kernel(ValueGenArgs a) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n) return;
    const int phase = a.phase[g];
    if (phase == VG_DONE) return;
    const u64 b1 = a.p1[g], b2 = a.p2[g];
    u64 own = a.color == 1 ? b1 : b2, opp = a.color == 1 ? b2 : b1;
    const u64 legal = legal_moves(own, opp);
    int k = -1;
    if (phase == VG_RANDOM) {
        a.rec_own[g] = own;                 // the mover's view of the position before the random move
        a.rec_opp[g] = opp;
        a.rec_color[g] = (uint8_t)a.color;
        if (!legal) {                       // value_self_play.py:47-48: return state, -1
            a.result[g] = -1;
            a.phase[g] = VG_DONE;
            return;
        }
        const int d = a.draws[g];
        k = nth_legal(legal, vg_uniform(a, g, d));
        a.draws[g] = d + 1;
        a.rec_action[g] = (int8_t)k;
        a.phase[g] = VG_RL;
    } else if (legal) {
        const float *lg = a.logits[phase == VG_SL ? 0 : 1] + (size_t)a.req_index[g] * 64;
        int d = a.draws[g];
        k = softmax_draw(lg, vg_uniform(a, g, d++));
        if (k > 63 || !((legal >> k) & 1)) k = nth_legal(legal, vg_uniform(a, g, d++));   // value_self_play.py:144-148
        a.draws[g] = d;
    }
    if (k >= 0) {
        place(1ULL << k, own, opp);
        a.p1[g] = a.color == 1 ? own : opp;
        a.p2[g] = a.color == 1 ? opp : own;
        a.pass_flg[g] = 0;
        a.stone_num[g] += 1;
    } else {
        if (a.pass_flg[g]) a.stone_num[g] = 64;   // game over when two players pass consecutively
        a.pass_flg[g] = 1;
    }
}

static int vg_ensure(iago_ctx *ctx, long long n, ValueGenWs **out) {
    if (!ctx->valuegen) ctx->valuegen = new ValueGenWs();
    ValueGenWs *w = static_cast<ValueGenWs *>(ctx->valuegen);
    if (w->cap < n) {
        for (int l = 0; l < 2; l++) {
            cudaFree(w->req_p1[l]); cudaFree(w->req_p2[l]); cudaFree(w->req_color[l]); cudaFree(w->logits[l]);
            IAGO_CUDA(cudaMalloc(&w->req_p1[l], n * 8));
            IAGO_CUDA(cudaMalloc(&w->req_p2[l], n * 8));
            IAGO_CUDA(cudaMalloc(&w->req_color[l], n));
            IAGO_CUDA(cudaMalloc(&w->logits[l], n * 64 * 4));
        }
        cudaFree(w->req_index); cudaFree(w->stone_num); cudaFree(w->pass_flg); cudaFree(w->phase);
        IAGO_CUDA(cudaMalloc(&w->req_index, n * 4));
        IAGO_CUDA(cudaMalloc(&w->stone_num, n * 4));
        IAGO_CUDA(cudaMalloc(&w->pass_flg, n));
        IAGO_CUDA(cudaMalloc(&w->phase, n));
        w->cap = n;
    }
    if (!w->counts) {
        IAGO_CUDA(cudaMalloc(&w->counts, 16));
        IAGO_CUDA(cudaMallocHost(&w->h_counts, 16));
    }
    *out = w;
    return IAGO_OK;
}

void valuegen_destroy(iago_ctx *ctx) {
    if (!ctx->valuegen) return;
    ValueGenWs *w = static_cast<ValueGenWs *>(ctx->valuegen);
    for (int l = 0; l < 2; l++) {
        cudaFree(w->req_p1[l]); cudaFree(w->req_p2[l]); cudaFree(w->req_color[l]); cudaFree(w->logits[l]);
    }
    cudaFree(w->req_index); cudaFree(w->stone_num); cudaFree(w->pass_flg); cudaFree(w->phase); cudaFree(w->counts);
    if (w->h_counts) cudaFreeHost(w->h_counts);
    delete w;
    ctx->valuegen = nullptr;
}

}  // namespace iago

using namespace iago;

extern "C" {

int iago_value_selfplay(iago_ctx *ctx, int slot_sl, int slot_rl, int64_t n, const int32_t *stop_num, int precision,
                        const iago_rng *rng, uint64_t *rec_own, uint64_t *rec_opp, uint8_t *rec_color, int8_t *rec_action,
                        int8_t *result, uint64_t *final_p1, uint64_t *final_p2, int32_t *draws, int64_t *stats, void *stream) {
    IAGO_REQUIRE(ctx && stop_num && rng && rec_own && rec_opp && rec_color && rec_action && result && final_p1 && final_p2 && draws,
                 "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    IAGO_REQUIRE(rng->mode == IAGO_RNG_PHILOX || rng->mode == IAGO_RNG_UNIFORMS, "rng.mode must be PHILOX or UNIFORMS");
    if (rng->mode == IAGO_RNG_UNIFORMS) IAGO_REQUIRE(rng->uniforms && rng->u_stride > 0, "rng.uniforms / u_stride");
    if (stats) stats[0] = stats[1] = 0;
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    ValueGenWs *w = nullptr;
    int rc = vg_ensure(ctx, n, &w);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + 127) / 128);
    ValueGenArgs a{};
    a.p1 = (u64 *)final_p1; a.p2 = (u64 *)final_p2;   // the game state lives in the caller's output arrays
    a.stop_num = stop_num; a.stone_num = w->stone_num; a.draws = draws; a.req_index = w->req_index; a.counts = w->counts;
    a.pass_flg = w->pass_flg; a.phase = w->phase;
    for (int l = 0; l < 2; l++) {
        a.req_p1[l] = w->req_p1[l]; a.req_p2[l] = w->req_p2[l]; a.req_color[l] = w->req_color[l]; a.logits[l] = w->logits[l];
    }
    a.n = n; a.rng_mode = rng->mode; a.stream_id = rng->stream_id; a.seed = rng->seed; a.game_id0 = rng->game_id0;
    a.uniforms = rng->uniforms; a.u_stride = rng->u_stride;
    a.rec_own = (u64 *)rec_own; a.rec_opp = (u64 *)rec_opp; a.rec_color = rec_color; a.rec_action = rec_action; a.result = result;
    valuegen_init_kernel<<<grid, 128, 0, s>>>(a);
    IAGO_CUDA(cudaGetLastError());
    long long turns = 0, forwards = 0;
    for (;;) {
        a.color = 1 + (int)(turns & 1);   // cl starts at 1 and flips once per turn in all three phases
        IAGO_CUDA(cudaMemsetAsync(w->counts, 0, 12, s));
        valuegen_prepare_kernel<<<grid, 128, 0, s>>>(a);
        IAGO_CUDA(cudaGetLastError());
        const int slots[2] = {slot_sl, slot_rl};
        for (int l = 0; l < 2; l++) {
            rc = trunk_launch(ctx, slots[l], 0, (const uint64_t *)w->req_p1[l], (const uint64_t *)w->req_p2[l], w->req_color[l], n,
                              w->logits[l], 0, precision, s, w->counts + l);
            if (rc) return rc;
            forwards++;
        }
        valuegen_turn_kernel<<<grid, 128, 0, s>>>(a);
        IAGO_CUDA(cudaGetLastError());
        turns++;
        if ((turns & 3) == 0) {
            IAGO_CUDA(cudaMemcpyAsync(w->h_counts, w->counts, 12, cudaMemcpyDeviceToHost, s));
            IAGO_CUDA(cudaStreamSynchronize(s));
            if (w->h_counts[2] == 0) break;   // no game was alive at the start of the last turn
        }
        if (turns > 400) {
            set_error("iago_value_selfplay: games did not terminate after %lld turns", turns);
            return IAGO_E_STATE;
        }
    }
    if (stats) {
        stats[0] = turns;
        stats[1] = forwards;
    }
    return IAGO_OK;
}

}  // extern "C"
