/*
 * iago_b200.h — C ABI of libiago_b200.so, the B200-native drop-in for IaGo's rollout / self-play /
 * PV-MCTS hot path.
 *
 * The reference (/root/reference) is pure Python and has no FFI of its own; its "plugin API" for this
 * path is the importable Python surface.  This header is what a ctypes stub on the
 * reference side binds; every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - Every function returns int: 0 = ok, < 0 = error (IAGO_E_*); iago_last_error() gives a thread-local
 *     message.  No C++ exception crosses this boundary.
 *   - Plain pointers and sizes only.  Unless a function name ends in _host, data pointers are DEVICE
 *     pointers owned by the caller (e.g. torch tensors); they are never freed or retained past the call.
 *     Small model parameters (weights) are HOST pointers and are copied.
 *   - Launches are asynchronous on the cudaStream_t passed as `void *stream`; NULL is CUDA's legacy default
 *     stream, exactly as in the runtime API (pass iago_ctx_stream(ctx) for the context's own non-blocking
 *     stream).  *_host entry points take HOST buffers, do H2D + kernel + D2H on the context's stream, and
 *     synchronise before return.
 *   - Boards are bitboard pairs: bit k <-> action k = row*8+col of the reference's 8x8 array
 *     (game.py:126,184).  p1 = stones of colour 1 (moves first), p2 = colour 2.  colour in {1,2}.
 *   - One context per device; a context is not re-entrant.  Different contexts may be used from
 *     different host threads.
 *   - There is NO CPU fallback: every compute entry point fails with IAGO_E_CUDA if no sm_100 device
 *     is usable.
 */
#ifndef IAGO_B200_H
#define IAGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IAGO_ABI_VERSION 1

#if defined(__GNUC__)
#define IAGO_API __attribute__((visibility("default")))
#else
#define IAGO_API
#endif

enum {
    IAGO_OK = 0,
    IAGO_E_INVALID = -1,   /* bad argument */
    IAGO_E_CUDA = -2,      /* CUDA runtime / driver error, or no usable device */
    IAGO_E_STATE = -3,     /* call order (e.g. weights not loaded) */
    IAGO_E_NOMEM = -4
};

typedef struct iago_ctx iago_ctx;

/* Resident SLPolicy / Value weight sets per context (net "slots"). */
#define IAGO_NET_SLOTS 32

/* How a game draws its moves.  Exactly one uniform is consumed per stone placed and none per pass
 * (mcts_self_play.py:100-110 -> np.random.choice -> one random_sample()). */
enum {
    IAGO_RNG_PHILOX = 0,    /* u = Philox4x32-10(key = seed, ctr = (game_lo, game_hi, draw, stream_id));
                               u = ((a>>5)*2^26 + (b>>6)) / 2^53, game = game_id0 + index              */
    IAGO_RNG_UNIFORMS = 1,  /* replay: the k-th stone of game g uses uniforms[g*u_stride + k]           */
    IAGO_RNG_FORCED = 2     /* replay: the k-th stone of game g is forced[g*f_stride + k] (no policy)   */
};

typedef struct iago_rng {
    int32_t mode;
    uint32_t stream_id;
    uint64_t seed;
    uint64_t game_id0;
    const double *uniforms; /* device (or host for *_host calls) */
    int64_t u_stride;
    const int8_t *forced;   /* device (or host for *_host calls) */
    int64_t f_stride;
} iago_rng;

IAGO_API int iago_abi_version(void);
IAGO_API const char *iago_last_error(void);

/* Device + stream + weight slots.  Replaces the per-object model construction of the reference
 * (mcts_self_play.py:18-19 reloads rollout_model.npz for every rollout). */
IAGO_API int iago_ctx_create(int device, iago_ctx **out);
IAGO_API int iago_ctx_destroy(iago_ctx *ctx);
IAGO_API int iago_ctx_sync(iago_ctx *ctx);
/* cudaStream_t of the context (for callers that want to order their own work after ours). */
IAGO_API void *iago_ctx_stream(iago_ctx *ctx);

/* RolloutPolicy parameters: conv1/W [1][2][3][3] and bias2/b [64] (network.py:49-64). HOST pointers. */
IAGO_API int iago_load_rollout(iago_ctx *ctx, const float *conv1_W, const float *bias2_b);

/* GameFunctions.legal_actions(state, color) (game.py:209-235; clones mcts_self_play.py:64-89,
 * src/rl_self_play.py:63-88, rl_env.py:114-138).  moves[i] = bit mask of legal actions (ascending bit
 * order = the reference's ascending list). */
IAGO_API int iago_legal_actions(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                       uint64_t *moves, int64_t n, void *stream);

/* GameFunctions.place_stone(state, action, color) (game.py:179-207), in place, no legality check,
 * action -1 = no-op (game.py:181-182). */
IAGO_API int iago_place_stone(iago_ctx *ctx, uint64_t *p1, uint64_t *p2, const int8_t *action, const uint8_t *color,
                     int64_t n, void *stream);

/* Simulate(state)(color) for n independent games (mcts_self_play.py:9-134), one lockstep kernel.
 *   in : p1, p2 [n] start boards; color [n] = the side that moves first AND the side the result is for
 *   out: result [n] in {+1,0,-1} (judge, mcts_self_play.py:113-121); final_p1/final_p2 [n];
 *        n_moves [n] stones placed (nullable); move_log [n][64] actions in order, -1 padded (nullable);
 *        counters[2] += {stones placed, turns taken} over the batch (nullable, device uint64)          */
IAGO_API int iago_rollout(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                 const iago_rng *rng, int8_t *result, uint64_t *final_p1, uint64_t *final_p2,
                 int32_t *n_moves, int8_t *move_log, uint64_t *counters, void *stream);

/* Same, HOST buffers in and out; synchronous.  This is the call the Python facade `Simulate` / `simulate_batch` makes and the
 * one bench.py's e2e figure times.  Pageable buffers are packed into pinned staging inside the context and pipelined in four
 * chunks (H2D / kernel / D2H overlap).  When EVERY buffer of the call is page-locked (cudaHostAlloc, cudaHostRegister, torch
 * pin_memory) the library uses them in place: with Philox uniforms one launch reads and writes the device-mapped buffers
 * itself (17 B in, 21 B out per game across PCIe), with a replay stream the async copies go straight from / to them. */
IAGO_API int iago_rollout_host(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                      const iago_rng *rng, int8_t *result, uint64_t *final_p1, uint64_t *final_p2,
                      int32_t *n_moves, int8_t *move_log, uint64_t *counters_host);

/* The same work for a caller that streams batches (mcts_self_play.py:137-150 calls Simulate once per game in a loop; a batched
 * caller keeps several batches in flight): submit returns at once, wait blocks until that lane's results are in the caller's
 * buffers.  lane in [0, IAGO_HOST_LANES): each lane has its own stream and device block, so the H2D / D2H copies of one batch run
 * on the copy engines while the kernel of another batch has the SMs.  Buffers must be page-locked and stay untouched until
 * the wait; Philox uniforms only (replay streams go through iago_rollout_host).  Errors: IAGO_E_STATE when the lane is still
 * in flight (submit) or idle (wait); IAGO_E_INVALID for pageable buffers. */
#define IAGO_HOST_LANES 4
IAGO_API int iago_rollout_host_submit(iago_ctx *ctx, int lane, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                             int64_t n, const iago_rng *rng, int8_t *result, uint64_t *final_p1, uint64_t *final_p2,
                             int32_t *n_moves, int8_t *move_log);
IAGO_API int iago_rollout_host_wait(iago_ctx *ctx, int lane, uint64_t *counters_host);

/* One Simulate.get_action draw per board (mcts_self_play.py:100-110): legal moves, rollout policy, masked
 * renormalised inverse-cdf sample.  No board update.  action[i] = -1 when the side to move has no legal move.
 * PHILOX: the draw-th uniform of game game_id0+i; UNIFORMS: uniforms[i] (u_stride ignored). */
IAGO_API int iago_rollout_sample(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                        const iago_rng *rng, uint32_t draw, int8_t *action, void *stream);

/* RolloutPolicy forward on make_state_var(state, color) (network.py:59-64, game.py:167-174):
 * logits [n][64] (pre-softmax) in the canonical summation order of DESIGN.md. */
IAGO_API int iago_rollout_logits(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                        float *logits, int64_t n, void *stream);

/* ---- SLPolicy / Value networks (network.py:15-47, 66-96) on the fused tcgen05 trunk kernel ----
 *
 * iago_load_net: replaces serializers.load_npz(path, model) (MCTS.py:83,85, game.py:20, src/train_rl.py:23,37).
 * `params` is a HOST array of fp32 in this order (the npz key layout of the reference, 'predictor/' prefix stripped):
 *   block1/conv/W [64][2][3][3], block1/conv/b [64], block2/conv/W [128][64][3][3], b [128],
 *   block3..8/conv/W [128][128][3][3], b [128],
 *   kind 0 (SLPolicy, 960,768 floats): conv9/W [1][128][1][1], bias10/b [64]
 *   kind 1 (Value,    970,049 floats): block9/conv/W [1][128][3][3], block9/conv/b [1], fc10/W [128][64], fc11/W [1][128]
 * The fp16 hi/lo split and the tensor-core operand layout are produced inside.  slot in 0..IAGO_NET_SLOTS-1
 * (a loaded slot holds about 15 MB of device memory: four operand layouts of the 3.8 MB weight set). */
IAGO_API int iago_load_net(iago_ctx *ctx, int slot, int kind, const float *params, int64_t n_floats);

/* SLPolicy.__call__ on make_state_var(state, color) for n positions given as bitboards.
 * out [n][64]: out_kind 0 = pre-softmax logits, 1 = softmax probabilities (what the reference returns).
 * precision 3 = error-compensated fp16 hi/lo split (3 MMAs, ~fp32 accuracy, max-abs logit error ~1e-4), 1 = single-pass fp16 (~0.1),
 * 2 = fp16 main product + the two cross terms in FP8 / E4M3 on a second accumulator (2 MMA units, ~3e-3; north-star bar 1e-2;
 * slots refreshed by a trainer run it as 3). */
IAGO_API int iago_policy_forward(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                                 int64_t n, float *out, int out_kind, int precision, void *stream);

/* SLPolicy forward that also keeps every block's output: acts[l] (DEVICE, nullable each) receives the post-ReLU output of
 * block l+1 (network.py:36-43) as fp32 [n][channels][64] (channels = 64 for l = 0, else 128); logits [n][64].  This is the
 * forward pass of the REINFORCE update (the backward needs the activations) and a test hook for layer-by-layer parity. */
IAGO_API int iago_policy_forward_acts(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                                      int64_t n, float *logits, float *const *acts, int precision, void *stream);
/* The same for the Value trunk (slot of kind 1): values [n] and the 8 blocks' outputs. */
IAGO_API int iago_value_forward_acts(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                                     int64_t n, float *values, float *const *acts, int precision, void *stream);

/* Value.__call__ (inference: dropout off, MCTS.py:86) -> out [n]. */
IAGO_API int iago_value_forward(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                                int64_t n, float *out, int precision, void *stream);

/* ---- batched learner-vs-opponent self-play: rl_self_play.Game(model1, model2)() (src/rl_self_play.py:8-145) ----
 * n games in lockstep; colour 1 = learner (net in slot_learner), colour 2 = opponent (slot_opponent).
 *   init_p1/init_p2 : start boards (NULL = the opening, rl_self_play.py:12-16).  stone_num starts at 4 whatever the
 *                     boards hold, exactly like the reference (src/train_rl.py:43-46 pokes an extra stone in).
 *   select          : 0 = sample from the masked, renormalised policy (rl_self_play.py:111-127; one uniform per move,
 *                     rng as for iago_rollout), 1 = greedy arg-max of the logits over legal moves (lowest index on ties)
 *   final_p1/final_p2 [n] : in/out game state, final boards on return;  result [n] = judge from colour 1's view
 *   rec_own/rec_opp [n][rec_cap], rec_action [n][rec_cap], n_rec [n] : the learner's pre-move positions (own =
 *                     learner's stones; the reference stores the same board with colours swapped,
 *                     rl_self_play.py:134-138) and chosen actions
 *   move_log [n][64] nullable; stats (HOST int64[3], nullable) = {turn pairs executed, trunk launches, positions the nets evaluated}
 * A trunk launch covers only the games that are still running and have a choice (>= 2 legal moves): a device-built request list with
 * a device-side length; a single legal move is played without the net (its uniform is still consumed).
 * Synchronises the stream once per pair of turns (termination test).  All other pointers are device pointers. */
IAGO_API int iago_selfplay(iago_ctx *ctx, int slot_learner, int slot_opponent, int64_t n, const uint64_t *init_p1,
                           const uint64_t *init_p2, int select, int precision, const iago_rng *rng, uint64_t *final_p1,
                           uint64_t *final_p2, int8_t *result, uint64_t *rec_own, uint64_t *rec_opp, int8_t *rec_action,
                           int32_t *n_rec, int rec_cap, int8_t *move_log, int64_t *stats, void *stream);

/* ---- gym-style environment: rl_env.GameEnv.step(action) for n environments (rl_env.py:41-74) ----
 * In/out DEVICE state per environment: p1/p2 boards, stone_num (starts at 4, rl_env.py:34), pass_flg, draws (uniforms consumed
 * so far).  The learner (colour 1) plays action[i]; an illegal action is replaced by positions[floor(u * len)] with the next
 * uniform (the reference calls Python's random.choice there, rl_env.py:46-48, a stream that cannot be replayed from doubles).
 * The opponent (colour 2, SLPolicy in slot_opponent) answers through get_position (rl_env.py:152-172): p = out - min(out)
 * over ALL 64 cells, renormalised, np.random.choice, re-drawn (one more uniform) until the cell is legal.
 * done[i] = two consecutive passes or stone_num >= 64.  opp_action (nullable) = the opponent's move, -1 = pass.
 * errors_host (HOST, nullable; reading it synchronises) = rejection loops that exceeded the reference's recursion limit
 * of 10,000 (rl_env.py:8; the reference dies with RecursionError there).  rng.mode PHILOX or UNIFORMS. */
IAGO_API int iago_env_step(iago_ctx *ctx, int slot_opponent, int precision, int64_t n, uint64_t *p1, uint64_t *p2,
                           int32_t *stone_num, uint8_t *pass_flg, const int8_t *action, const iago_rng *rng,
                           int32_t *draws, uint8_t *done, int8_t *opp_action, int32_t *errors_host, void *stream);

/* The sampler of get_position / get_position_self alone (rl_env.py:152-172, self_play.py:8-30): probs [n][64] DEVICE (SLPolicy
 * output for the mover), own/opp = mover's / other side's stones; action[i] = sampled legal cell, -1 = no legal move,
 * -2 = recursion limit hit.  draws in/out as for iago_env_step. */
IAGO_API int iago_sample_unmasked(iago_ctx *ctx, const float *probs, const uint64_t *own, const uint64_t *opp, int64_t n,
                                  const iago_rng *rng, int32_t *draws, int8_t *action, int32_t *errors_host, void *stream);

/* The masked sampler of get_action_auto / rl_self_play.get_action (game.py:101-108, rl_self_play.py:111-127) for given
 * probabilities: p = prob * validity mask, renormalised, np.random.choice with ONE uniform.  Arguments as for
 * iago_sample_unmasked; action -1 = no legal move. */
IAGO_API int iago_sample_masked(iago_ctx *ctx, const float *probs, const uint64_t *own, const uint64_t *opp, int64_t n,
                                const iago_rng *rng, int32_t *draws, int8_t *action, void *stream);

/* ---- PV-MCTS: MCTS.py:10-154 on a GPU-resident node pool, n_trees independent searches in lockstep ----
 * One iago_mcts holds n_trees trees (one per game) of at most max_nodes nodes each.  A search runs waves of up to
 * leaf_batch playouts per tree: leaf-parallel selection with virtual visits / virtual loss, ONE batched SLPolicy launch
 * for the expansions, ONE batched Value launch and ONE lockstep rollout launch for the leaves, then backup.
 * With leaf_batch = 1 the search is the reference's sequential playout loop, numpy scalar type for scalar type
 * (DESIGN.md "PV-MCTS"); that mode is what the parity tests compare with the reference's own trees. */
typedef struct iago_mcts iago_mcts;

typedef struct iago_mcts_params {
    double lmbda;          /* MCTS(lmbda=0.5): leaf_value = (1-lmbda) v + lmbda z        (MCTS.py:80,123-125) */
    double c_puct;         /* MCTS(c_puct=1)                                              (MCTS.py:48-49)     */
    double virtual_loss;   /* value subtracted per in-flight descent through a child (batched mode only)       */
    int32_t n_thr;         /* MCTS(n_thr=15): a leaf is expanded once it has n_thr visits (MCTS.py:108)        */
    int32_t leaf_batch;    /* playouts per tree per wave, 1 = the reference's sequential algorithm             */
    int32_t n_playouts;    /* playouts per tree in this call (the reference runs until time_limit, MCTS.py:139) */
    int32_t slot_policy;   /* net slots loaded with iago_load_net ('./models/sl_model.npz', value_model.npz)    */
    int32_t slot_value;
    int32_t precision;     /* 1, 2 or 3, as for iago_policy_forward                                              */
    int32_t cache_value;   /* 1 = keep Value(position) in the node instead of re-running it (same outputs)     */
    int32_t reserved;
    uint64_t seed;         /* Philox key of the rollouts: game id = (tree id << 32) | playout index, stream 2  */
    const float *forced_v; /* HOST, nullable: [n_trees][forced_stride] leaf values v by playout index (replay) */
    const int8_t *forced_z;/* HOST, nullable: rollout results z by playout index                                */
    int64_t forced_stride;
} iago_mcts_params;

IAGO_API int iago_mcts_create(iago_ctx *ctx, int n_trees, int max_nodes, int max_leaf_batch, uint64_t tree_id0,
                              iago_mcts **out);
IAGO_API int iago_mcts_destroy(iago_mcts *m);
/* Root positions (HOST arrays [n_trees]); reset_tree != 0 also starts fresh trees, root = Node(None, 1.0) (MCTS.py:81). */
IAGO_API int iago_mcts_set_roots(iago_mcts *m, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                                 int reset_tree, void *stream);
IAGO_API int iago_mcts_get_roots(iago_mcts *m, uint64_t *p1, uint64_t *p2, uint8_t *color, int64_t *playouts_done,
                                 void *stream);
/* MCTS.playout x n_playouts on every tree (MCTS.py:105-133, the loop of get_move :139-145). Synchronises at the end. */
IAGO_API int iago_mcts_search(iago_mcts *m, const iago_mcts_params *params, void *stream);
/* get_move's answer (MCTS.py:147): HOST outputs visits [n_trees][65], q [n_trees][65] (index = action, 64 = pass),
 * best [n_trees] = most visited child, lowest action on ties; -2 if the root has no children yet (the reference
 * raises ValueError there). */
IAGO_API int iago_mcts_root_stats(iago_mcts *m, int32_t *visits, float *q, int8_t *best, void *stream);
/* update_with_move(last_move) (MCTS.py:149-154) per tree: re-root on the child (its subtree is kept) or start a fresh
 * tree if the move is not a child; the root position follows the move (-1 = pass).  action, mask: HOST [n_trees];
 * mask nullable, trees with mask 0 do not move. */
IAGO_API int iago_mcts_advance(iago_mcts *m, const int8_t *action, const uint8_t *mask, void *stream);
/* Tree dump for tests: nodes in pool order (children of a node contiguous, ascending action). HOST outputs. */
IAGO_API int iago_mcts_export_tree(iago_mcts *m, int tree, int32_t capacity, int32_t *parent, int8_t *action,
                                   int32_t *n_visits, double *Q, double *P, int32_t *first_child, int32_t *n_children,
                                   int32_t *count, void *stream);
/* Expansions skipped because a tree's pool was full during the last search (0 in a correctly sized pool). */
IAGO_API int iago_mcts_overflows(iago_mcts *m, int64_t *count);

/* ---- value-network training data: value_self_play.py:12-59 driven by gen_value_data.py:12-19 ----
 * n lockstep SelfPlay(stop_num[g])() games: the policy in slot_sl plays both colours while stone_num < stop_num[g]; the
 * side to move then records the board from its own view (rec_own = its stones, rec_opp = the other side's; the reference
 * stores them as 2 and 1) and plays ONE uniformly random legal move; the policy in slot_rl plays the game out; result[g] =
 * judge(mover) in {1, 0, -1}, or -1 with rec_action -1 when the mover had no legal move at that point (value_self_play.py:47-48).
 * Sampling is get_position's (:131-149): softmax over all 64 net outputs, one uniform, an illegal cell replaced by
 * positions[floor(u' * len)] with the game's next uniform.  Uniform k of game g: Philox (seed, game_id0 + g, k, stream_id) or
 * uniforms[g * u_stride + k].  All arrays are DEVICE arrays [n]; draws[g] = uniforms consumed; stats (HOST, nullable) =
 * {turns, trunk launches}.  Synchronises the stream. */
IAGO_API int iago_value_selfplay(iago_ctx *ctx, int slot_sl, int slot_rl, int64_t n, const int32_t *stop_num, int precision,
                                 const iago_rng *rng, uint64_t *rec_own, uint64_t *rec_opp, uint8_t *rec_color,
                                 int8_t *rec_action, int8_t *result, uint64_t *final_p1, uint64_t *final_p2, int32_t *draws,
                                 int64_t *stats, void *stream);

/* ---- REINFORCE update of the SL-size policy: src/train_rl.py:55-66 (K6) ----
 * A trainer owns the learner's fp32 parameters (flat, iago_load_net order, kind 0), Adam moments and the activation
 * workspace for up to max_positions positions per call. */
typedef struct iago_trainer iago_trainer;
IAGO_API int iago_reinforce_create(iago_ctx *ctx, const float *params, int64_t n_floats, int max_positions, iago_trainer **out);
IAGO_API int iago_reinforce_destroy(iago_trainer *t);
/* Gradient of SUM_i c_i * r_i, c = softmax_cross_entropy(SLPolicy(x), y, reduce='no') on the PROBABILITIES the net returns
 * (the reference's double softmax, src/train_rl.py:61-64), for m recorded learner decisions: own/opp = learner's / opponent's
 * stones before the move (rl_self_play.py:134-138), action = y, reward = the game's judge per position.  DEVICE pointers.
 * grad: DEVICE float[960768 + 2]: the gradient in parameter order, then SUM c*r, then the position count — divide by the
 * count (after an all-reduce across ranks, if any) to get the reference's mean.  accumulate != 0 adds to grad.
 * probs_out (nullable, DEVICE [m][64]) receives pred. */
IAGO_API int iago_reinforce_grad(iago_trainer *t, const uint64_t *own, const uint64_t *opp, const int8_t *action,
                                 const float *reward, int64_t m, float *grad, int accumulate, float *probs_out, void *stream);
/* optimizer.update() with Chainer's Adam + WeightDecay hook (src/train_rl.py:24-26,66): g = grad / count + weight_decay * w;
 * m += (1-b1)(g-m); v += (1-b2)(g*g-v); w -= alpha*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps).  grad: DEVICE. */
IAGO_API int iago_reinforce_adam_step(iago_trainer *t, const float *grad, double count, double alpha, double beta1, double beta2,
                                      double eps, double weight_decay, void *stream);
/* Parameters / Adam state to and from HOST arrays (checkpoints: serializers.save_npz of model and optimizer,
 * src/train_rl.py:76-77); any pointer may be NULL; step < 0 keeps t. */
IAGO_API int iago_reinforce_get_state(iago_trainer *t, float *params, float *adam_m, float *adam_v, int64_t *step);
IAGO_API int iago_reinforce_set_state(iago_trainer *t, const float *params, const float *adam_m, const float *adam_v, int64_t step);
/* use_tensor_cores != 0 (default 1): forward on the fused tcgen05 trunk, the data gradients of blocks 8..2 as one fused
 * tcgen05 launch (bf16 hi/lo split, 3 MMAs; a single bf16 pass was measured at up to 4e-2 of max|g| and is not offered), weight gradients of the 128-output-channel layers as
 * fp16 tcgen05 GEMMs (dY scaled by a power of two and split hi/lo), all with fp32 accumulation; 0: every kernel in fp32 on the CUDA cores (the checker for that path). */
IAGO_API int iago_reinforce_set_option(iago_trainer *t, int use_tensor_cores);
/* Makes the trainer's current parameters the policy in net slot `slot` (what self-play then plays with). */
IAGO_API int iago_reinforce_sync_slot(iago_trainer *t, int slot);
/* The same, enqueued on `stream` behind the update that produced the parameters, without host synchronisation (a slot that holds no
 * net of this kind yet falls back to the synchronous path once). */
IAGO_API int iago_reinforce_sync_slot_async(iago_trainer *t, int slot, void *stream);
/* iago_reinforce_adam_step with the divisor read ON THE DEVICE: count = grad[n_params + 1] of the (all-reduced) vector
 * [gradient | loss numerator | count] — the update follows the all-reduce on the stream with no host round trip
 * (src/train_rl.py:66 on R ranks).  count = 0 leaves the parameters unchanged. */
IAGO_API int iago_reinforce_adam_step_dev(iago_trainer *t, const float *grad, double alpha, double beta1, double beta2, double eps,
                                          double weight_decay, void *stream);
/* Set plumbing of src/train_rl.py:41-53 on the device.  openings: p1 / p2 (DEVICE [n]) = the start position, odd games with the
 * extra un-flipped colour-2 stone on one of (2,4),(3,5),(4,2),(5,3) ("switch head and tail", :43-46; the cell = Philox(seed, global
 * game id, stream 6) & 3 where the reference uses random.choice).  compact: the learner records [n][rec_cap] of iago_selfplay
 * flattened game by game into out_* (DEVICE, capacity n * rec_cap) with reward = the game's result; out_count (DEVICE int32[3]) =
 * records, games won by the learner, largest n_rec (larger than rec_cap = records were lost). */
IAGO_API int iago_reinforce_openings(iago_ctx *ctx, int64_t n, uint64_t seed, uint64_t game_id0, uint64_t *p1, uint64_t *p2, void *stream);
IAGO_API int iago_reinforce_compact(iago_ctx *ctx, int64_t n, int rec_cap, const uint64_t *rec_own, const uint64_t *rec_opp,
                                    const int8_t *rec_action, const int32_t *n_rec, const int8_t *result, uint64_t *out_own,
                                    uint64_t *out_opp, int8_t *out_action, float *out_reward, int32_t *out_count, void *stream);

/* ---- collectives: sum all-reduces over NCCL, enqueued on the caller's stream ----
 * Games and trees are independent; the only exchanges are the fp32 vector [gradient | loss numerator | count] once per REINFORCE
 * update (src/train_rl.py:55-66 on R ranks) and int64 result counters at report time.  An iago_comm wraps an ncclComm_t: adopt one
 * the host created (iago_comm_from_nccl; the caller keeps ownership), or create one: rank 0 calls iago_comm_unique_id and sends the
 * 128 bytes to the other ranks by any host channel, then every rank calls iago_comm_create.  NCCL is bound at run time
 * (libnccl.so.2); without it these entry points return IAGO_E_STATE. */
typedef struct iago_comm iago_comm;
IAGO_API int iago_comm_unique_id(char *id, int64_t bytes);
IAGO_API int iago_comm_create(iago_ctx *ctx, const char *id, int rank, int world, iago_comm **out);
IAGO_API int iago_comm_from_nccl(iago_ctx *ctx, void *nccl_comm, int rank, int world, iago_comm **out);
IAGO_API int iago_comm_destroy(iago_comm *c);
IAGO_API int iago_comm_rank(iago_comm *c, int *rank, int *world);
IAGO_API int iago_comm_allreduce_sum_f32(iago_comm *c, float *buf, int64_t count, void *stream);
IAGO_API int iago_comm_allreduce_sum_i64(iago_comm *c, int64_t *buf, int64_t count, void *stream);

/* ---- supervised trainers: train_policy.py:16-84 (SL policy / rollout policy), train_value.py:9-70 ----
 * The SL policy is trained with the REINFORCE entry points above (reward 1 for every record gives exactly
 * F.softmax_cross_entropy(model(x), y), train_policy.py:61-62).  A Value trainer is created with kind 1 (970,049 floats in
 * iago_load_net order) and shares iago_reinforce_{destroy,adam_step,get_state,set_state,set_option,sync_slot}. */
IAGO_API int iago_trainer_create(iago_ctx *ctx, int kind, const float *params, int64_t n_floats, int max_positions, iago_trainer **out);
/* train_value.py:50-56 for m records: own / opp = stones of "2" / "1" (channel 1 / channel 0), target = game results.
 * grad (DEVICE float[970,049 + 2]) = gradient of SUM (v - y)^2 | that sum | m (mean_squared_error = the sum / m; the Adam step
 * divides once).  Training-mode dropout between fc10 and fc11 (network.py:94, ratio 0.4): unit i of record r is kept when
 * Philox(dropout_seed, position_id0 + r, i, stream 5) >= ratio, kept units scaled by 1 / (1 - ratio); ratio 0 = evaluation.
 * pred_out (nullable, DEVICE float[m]) = the forward's outputs; mask_out (nullable, DEVICE u8[m][128]) = the dropout mask. */
IAGO_API int iago_value_grad(iago_trainer *t, const uint64_t *own, const uint64_t *opp, const float *target, int64_t m, float *grad,
                             int accumulate, double dropout_ratio, uint64_t dropout_seed, uint64_t position_id0, float *pred_out,
                             uint8_t *mask_out, void *stream);
/* Test-set metrics (train_policy.py:66-70, train_value.py:58-62), ADDED into DEVICE float out[2] (zero it first):
 * policy: values = probabilities [n][64] (or logits when is_logits != 0): out[0] += sum softmax_cross_entropy(pred, y) with pred the
 * probabilities (the reference applies log-softmax to them again), out[1] += correct arg-maxes (F.accuracy);
 * value: out[0] += sum (pred - target)^2. */
IAGO_API int iago_policy_eval(iago_ctx *ctx, const float *values, int is_logits, const int8_t *action, int64_t n, float *out, void *stream);
IAGO_API int iago_value_eval(iago_ctx *ctx, const float *pred, const float *target, int64_t n, float *out, void *stream);
/* RolloutPolicy trainer (train_policy.py --policy rollout): 82 parameters = conv1/W [1][2][3][3] then bias2/b [64].
 * grad (DEVICE float[84]) = gradient of SUM softmax_cross_entropy | that sum | m; state (HOST float[3][82]) = params | Adam m | v. */
typedef struct iago_rollout_trainer iago_rollout_trainer;
IAGO_API int iago_rollout_trainer_create(iago_ctx *ctx, const float *conv1_W, const float *bias2_b, iago_rollout_trainer **out);
IAGO_API int iago_rollout_trainer_destroy(iago_rollout_trainer *t);
IAGO_API int iago_rollout_trainer_grad(iago_rollout_trainer *t, const uint64_t *own, const uint64_t *opp, const int8_t *action,
                                       int64_t m, float *grad, int accumulate, void *stream);
IAGO_API int iago_rollout_trainer_adam_step(iago_rollout_trainer *t, const float *grad, double count, double alpha, double beta1,
                                            double beta2, double eps, double weight_decay, void *stream);
IAGO_API int iago_rollout_trainer_get_state(iago_rollout_trainer *t, float *state, int64_t *step);
IAGO_API int iago_rollout_trainer_set_state(iago_rollout_trainer *t, const float *state, int64_t step);


/* Integer-issue micro-benchmark used as the roofline denominator of the rollout kernel:
 * runs `iters` rounds of dependent LOP3/SHF chains on every SM and returns int32 lane-ops/s. */
IAGO_API int iago_measure_int_peak(iago_ctx *ctx, int iters, double *lane_ops_per_s);

/* Timing helper for callers that cannot create CUDA events themselves (ctypes): elapsed ms of the last
 * iago_rollout* launch measured with CUDA events on the launching stream (kernel only). */
IAGO_API int iago_last_kernel_ms(iago_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* IAGO_B200_H */
