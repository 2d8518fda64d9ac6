// common.cuh — context object and error plumbing shared by the translation units of libiago_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/iago_b200.h"

namespace iago {

void set_error(const char *fmt, ...);

// Rollout policy in the form the kernels consume (built on the host in the canonical summation order).
// LUT index of a 3x3 neighbourhood: tap t = ky*3 + kx (cell (i+ky-1, j+kx-1)) sits at bit 6 - 3*ky + kx, i.e. row i+1 in bits
// 0-2, row i in bits 3-5, row i-1 in bits 6-8 (the order the kernel's multiply-gather produces, rollout.cu pat_index).
struct RolloutWeights {
    float lut[2][512];   // lut[c][index]  = sum over taps t ascending of W[c][t] for the set taps of the pattern
    float bias[64];
    uint32_t colmask[8]; // 3x3 window mask for column j: 0x070707 without the wrapped column at j = 0 / j = 7
    double elut[2][512]; // elut[c][index] = canon_exp(lut[c][index])                  (fast sampler: e = elut0 * elut1 * ebias)
    double ebias[64];    // ebias[k]       = canon_exp(bias[k])
    // the same tables for the board turned by 180 degrees (cell k -> 63 - k, tap t -> 8 - t, i.e. the 9-bit index reversed): the
    // lane of a rollout pair that owns board rows 4-7 plays the turned game (rollout.cu, rollout_pair_kernel)
    double elut_r[2][512];
    double ebias_r[64];
};

struct Staging {
    void *host = nullptr;  // pinned
    void *dev = nullptr;
    size_t bytes = 0;
};

// One in-flight iago_rollout_host_submit: its own stream, device block and counters (rollout.cu)
struct HostLane {
    cudaStream_t stream = nullptr;
    char *dev = nullptr;          // p1 | p2 | color | final_p1 | final_p2 | n_moves | result | move log
    size_t bytes = 0;
    uint64_t *d_cnt = nullptr;    // {stones, turns, ticket, -}: zero between launches
    uint64_t *h_cnt = nullptr;    // page-locked, device-mapped: the kernel's last CTA publishes the totals here
    bool busy = false;
};
constexpr int kHostLanes = 4;

}  // namespace iago

struct iago_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t host_streams[4] = {};   // iago_rollout_host's chunk pipeline ([0] = stream, the others are created on first use)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    iago::RolloutWeights *d_rollout = nullptr;
    bool rollout_loaded = false;
    bool rollout_fast = false;   // every possible |logit| <= 300: the product-of-exponentials sampler cannot over/underflow in double
    iago::Staging stage;
    iago::HostLane lanes[iago::kHostLanes];   // iago_rollout_host_submit / _wait
    uint64_t *d_counters = nullptr;
    void *trunk = nullptr;     // conv-net state (trunk.cu)
    void *selfplay = nullptr;  // self-play workspace (selfplay.cu)
    void *mcts = nullptr;      // search trees (mcts.cu)
    void *valuegen = nullptr;  // value-data generation workspace (valuegen.cu)
};

#define IAGO_CUDA(expr)                                                                      \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            iago::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return IAGO_E_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define IAGO_REQUIRE(cond, msg)                     \
    do {                                            \
        if (!(cond)) {                              \
            iago::set_error("invalid argument: %s", msg); \
            return IAGO_E_INVALID;                  \
        }                                           \
    } while (0)

namespace iago {
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
int ensure_staging(iago_ctx *ctx, size_t bytes);
// trunk.cu: SLPolicy (want_kind 0) / Value (1) forward on device bitboards; out_kind 0 = logits, 1 = probabilities.
// n_dev (nullable, device): the live count is min(n, *n_dev) — for request lists whose length only the device knows.
int trunk_launch(iago_ctx *ctx, int slot, int want_kind, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                 int64_t n, float *out, int out_kind, int precision, void *stream, const int *n_dev = nullptr,
                 float *const *dump = nullptr);  // dump[l] (nullable): fp32 [n][channels][64] output of block l+1
// rollout.cu: Simulate for n device-resident games whose Philox game ids are given one by one (mcts.cu leaf batches).
int rollout_launch_ids(iago_ctx *ctx, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                       uint64_t seed, uint32_t stream_id, const uint64_t *game_ids, int8_t *result, uint64_t *final_p1,
                       uint64_t *final_p2, void *stream);
bool trunk_slot_holds(iago_ctx *ctx, int slot, int kind);
void valuegen_destroy(iago_ctx *ctx);
// trunk.cu, backward data-gradient chain of the SLPolicy trunk on the tensor cores (used by reinforce.cu):
//   pack: W[l] (l = 1..7, device fp32 [128][cin_l][3][3]) -> bf16 hi/lo weight units in `blob`
//   launch: dy_in = gradient w.r.t. block 8's output [n][128][64]; chain layer i (i = 0..6) writes the gradient w.r.t. the output
//   of block 7-i to dx_out[i] ([n][cin][64], cin = 64 for i = 6), gated by mask[i] > 0 (that block's forward output).
// trunk.cu: rewrite the weight blob of an already loaded slot of the same kind from DEVICE fp32 parameters (iago_load_net order).
int trunk_refresh_slot(iago_ctx *ctx, int slot, int kind, const float *d_params, void *stream);
size_t trunk_backward_blob_bytes();
int trunk_backward_pack(iago_ctx *ctx, const float *const *W, uint8_t *blob, void *stream);
int trunk_backward_launch(iago_ctx *ctx, const uint8_t *blob, const float *dy_in, const float *const *mask,
                          float *const *dx_out, int64_t n, int precision, void *stream,
                          unsigned *const *dymax = nullptr);   // dymax[i] (nullable): atomicMax of the bit pattern of |dx_out[i]|
}  // namespace iago
