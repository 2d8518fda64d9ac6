/*
 * othello_ref.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's rollout hot path, written the way the
 * reference writes it: an 8x8 float board (0 empty / 1 / 2), ray walks over the
 * 8 directions, one Python-style turn loop.  It deliberately shares NO code and NO
 * data structure with the CUDA product path (which is bitboards), so that agreement
 * between the two is evidence.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Follows (all paths under /root/reference):
 *   legal_actions   game.py:209-235      (clones mcts_self_play.py:64-89, src/rl_self_play.py:63-88)
 *   place_stone     game.py:179-207      (clones mcts_self_play.py:36-62, src/rl_self_play.py:36-61)
 *   make_state_var  game.py:167-174      (mcts_self_play.py:91-97)
 *   RolloutPolicy   network.py:49-64     conv1 2->1 3x3 pad 1 no bias, + bias2.b[64], softmax
 *   get_action      mcts_self_play.py:100-110   p = softmax(f32) * valid(f64); np.random.choice(64, p/sum)
 *   turn / __call__ / judge   mcts_self_play.py:124-134 / 25-29 / 113-121
 *
 * Pinning: tests/test_oracle_golden.py checks this file against golden vectors produced
 * by running the UNMODIFIED reference Python (oracle/gen_golden.py) — legal sets, boards
 * after place_stone, perft 1..6, and full Simulate trajectories under the same uniform stream.
 *
 * Floating point (shared, bit for bit, with the CUDA kernel — see DESIGN.md "canonical
 * rollout arithmetic"):
 *   logit[k] = (S0 + S1) + b[k];  S_c = sum over taps t = ky*3+kx ascending of W[c][t]*x[c][..]
 *              (x is 0/1 so every product is exact; zero taps are exact no-ops)
 *   Two samplers, chosen per weight set (the same rule as the product, decided from the weights alone):
 *   FAST  when max|S0| + max|S1| + max|b| <= 300 over all 512 tap patterns (any finite, sanely trained rollout net):
 *     w[k]   = (E0 * E1) * EB in double,  E_c = canon_exp((double)S_c), EB = canon_exp((double)b[k])
 *              — the softmax numerator as a product of exponentials; the reference's max subtraction, denominator
 *              and fp32 division cancel in p/sum(p), and its float64 cdf is float64 here too
 *     choice = first legal k (ascending) whose cdf exceeds T = u * total, u = m53 / 2^53, with a two-sided double cdf:
 *              A_k = running sum over the legal cells 0..31 ascending, D_k = running sum over the legal cells 63..32
 *              descending, total = A_last + D_last; k < 32 (taken when T < A_last): first k with A_k > T; k >= 32: first k
 *              with D_(legal cells above k) < total - T.  In exact arithmetic this IS
 *              searchsorted(cumsum(p)/cumsum(p)[-1], u, 'right') of np.random.choice; in double it differs from the
 *              reference's cdf by ~1e-16, far below the rounding error of the reference's own fp32 softmax (~1e-9 in cdf
 *              units).  (Two-sided because the kernel gives each half of the board to its own lane.)
 *     canon_exp = Cody-Waite reduction + degree-13 Taylor in double, fixed operation order (no libm dependence)
 *   SAFE  otherwise:
 *     e[k]   = exp32(logit[k] - max over LEGAL k)   only at legal cells
 *     choice = fixed-point inverse cdf: q_k = floor(e_k * 2^26) (uint32), first legal k (ascending) with
 *              cum_k > floor(u32 * total / 2^32), u32 = m53 >> 21, m53 = floor(u * 2^53)
 *   exp32    = Cephes-style range reduction + degree-5 polynomial, every step an explicit
 *              fmaf / single rounding, so gcc and nvcc produce identical bits.
 *   uniforms = Philox4x32-10, key = seed, counter = (game_lo, game_hi, draw >> 2, stream);
 *              draw d uses output word d & 3: u = word / 2^32 (one block serves four consecutive draws)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ rules */

static const int DYS[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
static const int DXS[8] = {-1, 0, 1, -1, 1, -1, 0, 1};

static inline int is_outside(int y, int x) { return y < 0 || y > 7 || x < 0 || x > 7; }

/* game.py:209-235 — returns count; out[] ascending (row-major scan order) */
EXPORT int oracle_legal_actions(const float *state, int color, int *out) {
    int n = 0;
    for (int i = 0; i < 8; i++) {
        for (int j = 0; j < 8; j++) {
            if (state[i * 8 + j] != 0.0f) continue;
            for (int d = 0; d < 8; d++) {
                int dy = DYS[d], dx = DXS[d];
                if (is_outside(i + dy, j + dx)) continue;
                if (state[(i + dy) * 8 + (j + dx)] + (float)color != 3.0f) continue;
                int ry = i + dy, rx = j + dx, out_flg = 0;
                while (state[ry * 8 + rx] + (float)color == 3.0f) {
                    ry += dy; rx += dx;
                    out_flg = is_outside(ry, rx);
                    if (out_flg) break;
                }
                if (out_flg) continue;
                if (state[ry * 8 + rx] == (float)color) { out[n++] = i * 8 + j; break; }
            }
        }
    }
    return n;
}

/* game.py:179-207 — in place; action -1 is a no-op (game.py:181-182); no legality check */
EXPORT void oracle_place_stone(float *state, int action, int color) {
    if (action == -1) return;
    int py = action / 8, px = action % 8;
    state[py * 8 + px] = (float)color;
    for (int d = 0; d < 8; d++) {
        int dy = DYS[d], dx = DXS[d];
        if (is_outside(py + dy, px + dx)) continue;
        if (state[(py + dy) * 8 + (px + dx)] + (float)color != 3.0f) continue;
        int ry = py + dy, rx = px + dx;
        while (state[ry * 8 + rx] + (float)color == 3.0f) {
            ry += dy; rx += dx;
            if (is_outside(ry, rx)) break;
        }
        if (is_outside(ry, rx)) continue;
        if (state[ry * 8 + rx] == (float)color) {
            ry -= dy; rx -= dx;
            while (state[ry * 8 + rx] + (float)color == 3.0f) {
                state[ry * 8 + rx] = (float)color;
                ry -= dy; rx -= dx;
            }
        }
    }
}

/* perft with pass handling (a pass is a ply only if the opponent can move) — used to pin the
 * rules against the known series 4, 12, 56, 244, 1396, 8200, 55092, 390216. */
static uint64_t perft_rec(const float *state, int color, int depth, int passed) {
    if (depth == 0) return 1;
    int acts[64];
    int n = oracle_legal_actions(state, color, acts);
    if (n == 0) {
        if (passed) return 1;
        return perft_rec(state, 3 - color, depth - 1, 1);
    }
    uint64_t total = 0;
    for (int a = 0; a < n; a++) {
        float s[64];
        memcpy(s, state, sizeof s);
        oracle_place_stone(s, acts[a], color);
        total += perft_rec(s, 3 - color, depth - 1, 0);
    }
    return total;
}

EXPORT uint64_t oracle_perft(const float *state, int color, int depth) {
    return perft_rec(state, color, depth, 0);
}

/* ------------------------------------------------------------------ Philox4x32-10 */

static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

EXPORT void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

static inline uint64_t philox_m53(uint64_t seed, uint64_t game, uint32_t draw, uint32_t stream) {
    uint32_t o[4];
    /* one block serves four consecutive draws: draw d takes word d & 3 of block d >> 2; u = word / 2^32 */
    philox4x32_10((uint32_t)game, (uint32_t)(game >> 32), draw >> 2, stream,
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return (uint64_t)o[draw & 3] << 21;
}

static inline double philox_uniform(uint64_t seed, uint64_t game, uint32_t draw, uint32_t stream) {
    return (double)philox_m53(seed, game, draw, stream) * (1.0 / 9007199254740992.0);
}

EXPORT double oracle_philox_uniform(uint64_t seed, uint64_t game, uint32_t draw, uint32_t stream) {
    return philox_uniform(seed, game, draw, stream);
}

/* ------------------------------------------------------------------ rollout policy */

/* exp32 for x <= 0; every operation is a single IEEE rounding (see header). */
static inline float exp32_neg(float x) {
    if (x < -80.0f) return 0.0f;
    float z = x * 1.44269504088896341f;
    float n = rintf(z);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.42860682030941723e-6f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r);
    y = y + 1.0f;
    union { uint32_t u; float f; } s;
    s.u = (uint32_t)((int)n + 127) << 23;
    return y * s.f;
}

EXPORT float oracle_exp32_neg(float x) { return exp32_neg(x); }

/* network.py:59-64 on make_state_var(state, color): channel 0 = opponent, channel 1 = mover
 * (game.py:169-171 swaps 1<->2 for colour 1 via state*(3-state)*(3-state)/2, exact in fp32, then
 * channel c = (state == c+1); the net effect is channel 0 = (cell == 3-color), channel 1 = (cell == color)).
 * One output cell, canonical summation order. */
static inline float rollout_logit_at(const float *state, int color, const float *W, const float *b, int i, int j) {
    float S[2];
    for (int c = 0; c < 2; c++) {
        const float who = (c == 0) ? (float)(3 - color) : (float)color;
        float acc = 0.0f;
        for (int ky = 0; ky < 3; ky++)
            for (int kx = 0; kx < 3; kx++) {
                int y = i + ky - 1, x = j + kx - 1;
                if (is_outside(y, x)) continue;
                if (state[y * 8 + x] == who) acc = acc + W[c * 9 + ky * 3 + kx];
            }
        S[c] = acc;
    }
    return (S[0] + S[1]) + b[i * 8 + j];
}

/* logits[64] before softmax */
EXPORT void oracle_rollout_logits(const float *state, int color, const float *W /*[2][3][3]*/,
                                  const float *b /*[64]*/, float *logits) {
    for (int k = 0; k < 64; k++) logits[k] = rollout_logit_at(state, color, W, b, k / 8, k % 8);
}

/* exp(x) in double, fixed operation sequence (see header). */
static double canon_exp(double x) {
    const double n = nearbyint(x * 1.4426950408889634);
    double r = x - n * 0.693147180369123816490;
    r = r - n * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;
    const double inv[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
                            1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
    for (int i = 0; i < 13; i++) p = p * r + inv[i];
    return ldexp(p, (int)n);
}
EXPORT double oracle_canon_exp(double x) { return canon_exp(x); }

/* Per weight set: which sampler applies, and (FAST) the exponentials of every tap-pattern sum, indexed by the pattern with
 * bit t = tap t = ky*3+kx set when that neighbour holds the plane's stone. */
typedef struct {
    int fast;
    double E[2][512];
    double EB[64];
} policy_tables;

static void build_tables(const float *W, const float *b, policy_tables *t) {
    float amax[2] = {0.0f, 0.0f}, bmax = 0.0f;
    int finite = 1;
    for (int c = 0; c < 2; c++)
        for (int pat = 0; pat < 512; pat++) {
            float acc = 0.0f;
            for (int k = 0; k < 9; k++) if (pat >> k & 1) acc = acc + W[c * 9 + k];
            t->E[c][pat] = canon_exp((double)acc);
            if (fabsf(acc) > amax[c]) amax[c] = fabsf(acc);
            if (!(fabsf(acc) <= 300.0f)) finite = 0;   /* also catches NaN */
        }
    for (int k = 0; k < 64; k++) {
        t->EB[k] = canon_exp((double)b[k]);
        if (fabsf(b[k]) > bmax) bmax = fabsf(b[k]);
        if (!(fabsf(b[k]) <= 300.0f)) finite = 0;
    }
    t->fast = finite && (amax[0] + amax[1] + bmax) <= 300.0f;
}

EXPORT int oracle_policy_is_fast(const float *W, const float *b) {
    policy_tables t;
    build_tables(W, b, &t);
    return t.fast;
}

/* FAST weight of one cell: the tap patterns are read off the board exactly as rollout_logit_at walks it. */
static inline double rollout_weight_at(const float *state, int color, const policy_tables *t, int i, int j) {
    int pat[2] = {0, 0};
    for (int c = 0; c < 2; c++) {
        const float who = (c == 0) ? (float)(3 - color) : (float)color;
        for (int ky = 0; ky < 3; ky++)
            for (int kx = 0; kx < 3; kx++) {
                int y = i + ky - 1, x = j + kx - 1;
                if (is_outside(y, x)) continue;
                if (state[y * 8 + x] == who) pat[c] |= 1 << (ky * 3 + kx);
            }
    }
    return (t->E[0][pat[0]] * t->E[1][pat[1]]) * t->EB[i * 8 + j];
}

/* mcts_self_play.py:100-110 with the uniform supplied by the caller as m53 = floor(u * 2^53). */
static int sample_action(const float *state, int color, const int *actions, int n,
                         const float *W, const float *b, const policy_tables *t, uint64_t m53) {
    if (n == 1) return actions[0];
    if (t->fast) {
        /* only the legal cells survive the mask (mcts_self_play.py:103-105).  Two-sided cdf: the cells of board rows 0-3 are
         * summed ascending (A), those of rows 4-7 descending (D); cdf(k) = A_k for k < 32 and total - D_(cells above k) else. */
        double A[32], D[32], lo = 0.0, hi = 0.0;
        int cl[32], ch[32], nl = 0, nh = 0;
        for (int a = 0; a < n; a++)
            if (actions[a] < 32) {
                lo = lo + rollout_weight_at(state, color, t, actions[a] / 8, actions[a] % 8);
                A[nl] = lo;
                cl[nl++] = actions[a];
            }
        for (int a = n - 1; a >= 0; a--)
            if (actions[a] >= 32) {
                hi = hi + rollout_weight_at(state, color, t, actions[a] / 8, actions[a] % 8);
                D[nh] = hi;
                ch[nh++] = actions[a];
            }
        const double u = (double)(int64_t)m53 * 1.1102230246251565e-16; /* m53 / 2^53, exact */
        const double total = lo + hi;
        const double T = u * total;
        if (T < lo || nh == 0) {            /* first k with A_k > T */
            int idx = 0;
            for (int i = 0; i < nl; i++) idx += (A[i] <= T);
            return cl[idx < nl - 1 ? idx : nl - 1];
        }
        const double R = total - T;          /* first k (ascending) with total - D_(above k) > T  <=>  D_(above k) < R */
        int idx = 0;
        for (int j = 0; j + 1 < nh; j++) idx += (D[j] < R);
        return ch[idx];
    }
    float logits[64];
    for (int a = 0; a < n; a++) logits[actions[a]] = rollout_logit_at(state, color, W, b, actions[a] / 8, actions[a] % 8);
    float m = logits[actions[0]];
    for (int a = 1; a < n; a++) if (logits[actions[a]] > m) m = logits[actions[a]];
    uint32_t cum[64], total = 0;
    for (int a = 0; a < n; a++) {
        float e = exp32_neg(logits[actions[a]] - m);
        total += (uint32_t)(e * 67108864.0f); /* 2^26: exact scaling, truncating convert */
        cum[a] = total;
    }
    uint32_t T = (uint32_t)(((uint64_t)(uint32_t)(m53 >> 21) * total) >> 32);
    for (int a = 0; a < n; a++) if (cum[a] > T) return actions[a];
    return actions[n - 1];
}

static inline uint64_t m53_of_double(double u) { return (uint64_t)(u * 9007199254740992.0); }

EXPORT int oracle_rollout_sample(const float *state, int color, const float *W, const float *b, double u) {
    int acts[64];
    int n = oracle_legal_actions(state, color, acts);
    if (n == 0) return -1;
    policy_tables t;
    build_tables(W, b, &t);
    return sample_action(state, color, acts, n, W, b, &t, m53_of_double(u));
}

/* ------------------------------------------------------------------ Simulate */

enum { RNG_PHILOX = 0, RNG_UNIFORMS = 1, RNG_FORCED = 2 };

typedef struct {
    int mode;
    uint64_t seed;
    uint32_t stream;
    const double *uniforms; /* [n][u_stride]  : k-th stone placed in game g uses uniforms[g*u_stride+k] */
    int64_t u_stride;
    const int8_t *forced;   /* [n][f_stride]  : k-th stone placed in game g is forced[g*f_stride+k]      */
    int64_t f_stride;
} rng_spec;

/* One Simulate(state)(color): mcts_self_play.py:11-29,113-134. Returns result for `color`. */
static int simulate_one(float *state, int color, uint64_t game_id, int64_t g, const rng_spec *rng,
                        const float *W, const float *b, const policy_tables *tab, int8_t *moves /*[64] or NULL*/,
                        int *n_moves, int *n_turns) {
    int stone_num = 0;
    for (int k = 0; k < 64; k++) stone_num += (state[k] != 0.0f); /* 64 - sum(state==0) */
    int pass_flg = 0, placed = 0, turns = 0;
    while (stone_num < 64) {
        for (int half = 0; half < 2; half++) {
            int c = half == 0 ? color : 3 - color;
            int acts[64];
            int n = oracle_legal_actions(state, c, acts);
            turns++;
            if (n > 0) {
                int action;
                if (rng->mode == RNG_FORCED) {
                    action = rng->forced[g * rng->f_stride + placed];
                } else {
                    uint64_t m = rng->mode == RNG_UNIFORMS
                                     ? m53_of_double(rng->uniforms[g * rng->u_stride + placed])
                                     : philox_m53(rng->seed, game_id, (uint32_t)placed, rng->stream);
                    action = sample_action(state, c, acts, n, W, b, tab, m);
                }
                oracle_place_stone(state, action, c);
                if (moves) moves[placed] = (int8_t)action;
                placed++;
                pass_flg = 0;
                stone_num += 1;
            } else {
                if (pass_flg) stone_num = 64;
                pass_flg = 1;
            }
        }
    }
    if (moves) for (int k = placed; k < 64; k++) moves[k] = -1;
    *n_moves = placed;
    *n_turns = turns;
    int me = 0, op = 0;
    for (int k = 0; k < 64; k++) {
        me += (state[k] == (float)color);
        op += (state[k] == (float)(3 - color));
    }
    return me > op ? 1 : (me < op ? -1 : 0);
}

typedef struct {
    float *states; const int *colors; int64_t n; const float *W; const float *b; const policy_tables *tab;
    const rng_spec *rng; uint64_t game_id0;
    int8_t *results; int8_t *moves; int32_t *n_moves; int32_t *n_turns;
    atomic_llong *next;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (;;) {
        int64_t g0 = atomic_fetch_add(j->next, 64);
        if (g0 >= j->n) break;
        int64_t g1 = g0 + 64 < j->n ? g0 + 64 : j->n;
        for (int64_t g = g0; g < g1; g++) {
            int nm = 0, nt = 0;
            int r = simulate_one(j->states + g * 64, j->colors[g], j->game_id0 + (uint64_t)g, g, j->rng,
                                 j->W, j->b, j->tab, j->moves ? j->moves + g * 64 : 0, &nm, &nt);
            j->results[g] = (int8_t)r;
            if (j->n_moves) j->n_moves[g] = nm;
            if (j->n_turns) j->n_turns[g] = nt;
        }
    }
    return 0;
}

EXPORT int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}

/*
 * Batch of independent Simulate runs.  states: [n][64] float (0/1/2), modified in place to the
 * final boards.  colors[n] in {1,2} = the `color` argument of Simulate.__call__ (moves first).
 * game_id0 + g is the Philox game id.  moves may be NULL.  threads <= 0 -> all online cores
 * (pthreads, dynamic chunks of 64 games).  Returns the number of threads used.
 */
EXPORT int oracle_simulate_batch(float *states, const int *colors, int64_t n, const float *W, const float *b,
                                 int mode, uint64_t seed, uint32_t stream, uint64_t game_id0,
                                 const double *uniforms, int64_t u_stride,
                                 const int8_t *forced, int64_t f_stride,
                                 int8_t *results, int8_t *moves, int32_t *n_moves, int32_t *n_turns,
                                 int threads) {
    rng_spec rng = {mode, seed, stream, uniforms, u_stride, forced, f_stride};
    if (threads <= 0) threads = oracle_max_threads();
    if (threads > 256) threads = 256;
    atomic_llong next = 0;
    policy_tables tab;
    if (mode != RNG_FORCED) build_tables(W, b, &tab); else tab.fast = 0;
    batch_job job = {states, colors, n, W, b, &tab, &rng, game_id0, results, moves, n_moves, n_turns, &next};
    if (threads == 1) { batch_worker(&job); return 1; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < threads - 1; t++)
        if (pthread_create(&th[started], 0, batch_worker, &job) == 0) started++;
    batch_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], 0);
    return started + 1;
}
