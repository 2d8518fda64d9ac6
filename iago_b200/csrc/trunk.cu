// trunk.cu — K3/K4: SLPolicy / Value forward as ONE fused, persistent tcgen05 kernel per batch.
//
// Reference: network.py:5-13 (Block = 3x3 conv pad 1 + bias + ReLU), network.py:15-47 (SLPolicy: 8 blocks,
// conv9 1x1 128->1 no bias, +bias10.b[64], softmax), network.py:66-96 (Value: same trunk, block9 3x3 128->1
// + bias + ReLU, fc10 64->128 no bias, dropout off at inference, fc11 128->1 no bias), input encoding
// game.py:167-174 (channel 0 = opponent stones, channel 1 = mover's stones).
//
// Design (DESIGN.md "conv trunk"):
//   * One CTA owns a tile of 2 boards = 128 positions = the M of one tcgen05.mma (cta_group::1, M=128).
//     Row m of the tile is (board-row r, board b, column c) with m = (r*2+b)*8 + c, so that an 8-row UMMA core
//     matrix is one board row and consecutive core matrices are a constant 160 B apart.
//   * The whole trunk stays on chip.  The activation tile lives in shared memory as
//     [channel-group of 8][padded row 0..9][board 0..1][padded col 0..9][8 x fp16] (no-swizzle K-major canonical
//     layout).  A 3x3 tap (dy,dx) is then just a different descriptor START ADDRESS on the same tile — implicit
//     GEMM with the zero halo supplying the conv padding; nothing is ever re-laid-out or written to HBM.
//   * Weights are pre-packed on the host into the B-operand core-matrix layout, one "unit" per
//     (layer, tap, 64-channel chunk), and streamed L2 -> smem with cp.async.bulk (TMA 1D) through a 3-stage
//     mbarrier ring by a producer thread that runs ahead across layer and tile boundaries.
//   * Accumulators live in TMEM (128 fp32 columns).  The epilogue warps read them with tcgen05.ld, add bias,
//     ReLU, and write the next layer's activation tile in place (the MMAs of the layer are complete by then).
//   * Precision: fp16 operands, fp32 accumulate.  precision=3 splits both operands into hi + lo fp16 parts and
//     issues three MMAs per step (hi*hi + lo*hi + hi*lo), which recovers ~fp32 accuracy (max-abs logit error
//     ~1e-4 on sl_model.npz); precision=1 is the single-pass fp16 path (max-abs ~0.1).  SURVEY.md §0.5.
//     precision=2 keeps hi*hi in fp16 and computes the two cross terms, which are 2^-11 of the product, in FP8 (E4M3, kind::f8f6f4,
//     twice the fp16 rate): cross = fp8(a_lo * 2^11) * fp8(w_hi * sw) + fp8(a_hi) * fp8(w_lo * 2^11 * sw) into a second TMEM
//     accumulator that the epilogue folds in with the factor 1 / (2^11 * sw) (sw = a power of two per layer from max|w|).  Two MMA
//     units per K step instead of three; max-abs logit error ~3e-3 on sl_model.npz (north-star bar 1e-2), arg-max unchanged.
//   * Heads: policy = per-row dot with conv9 (fp32, CUDA cores) in the layer-8 epilogue (+bias10, optional
//     softmax); value = block9 as a ninth MMA layer with N padded to 16, then relu and the collapsed
//     fc11*fc10 64-vector.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "bitboard.cuh"
#include "common.cuh"
#include "tc.cuh"

namespace iago {

#ifdef IAGO_TRUNK_TRACE
// debug build only (tools/trace_trunk.py): SM clock at the pipeline's hand-over points of CTA 0, first tiles
__device__ unsigned long long g_trace[4096];
#define TRACE(it_, l_, ev_) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it_) < 4) g_trace[((it_) * 9 + (l_)) * 8 + (ev_)] = clock64(); } while (0)
#define TRACEE(it_, l_, k_) do { if (blockIdx.x == 0 && threadIdx.x == 0 && (it_) == 1) g_trace[3072 + (l_) * 8 + (k_)] = clock64(); } while (0)
#define TRACEU(it_, l_, u_, ev_) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it_) == 1 && (l_) == 2) g_trace[2048 + (u_) * 8 + (ev_)] = clock64(); } while (0)
#else
#define TRACE(tile_, l_, ev_) do { } while (0)
#define TRACEU(tile_, l_, u_, ev_) do { } while (0)
#define TRACEE(tile_, l_, k_) do { } while (0)
#endif

// ---------------------------------------------------------------- geometry
constexpr int kTileRows = 128;                    // M of the MMA = 2 boards
constexpr int kGroupBytes = 10 * 2 * 10 * 16;     // one 8-channel group of the padded tile: 3200 B  (= LBO of A)
constexpr int kRowPitch = 10 * 16;                // 160 B between consecutive 8-row core matrices     (= SBO of A)
constexpr int kActBytes = 16 * kGroupBytes;       // 128 channels: 51,200 B per precision part
constexpr int kA1Bytes = 4 * kTileRows * 16;      // layer-1 explicit im2col tile, K = 32: 8,192 B
constexpr int kStageBytes = 32768;                // one weight unit: hi [8][128][8] + lo
constexpr int kStages = 3;                        // one producer warp (9, 10, 11) per ring stage
static_assert(kStages == 3, "the producer warps are numbered by ring stage");
// 2-D tensor maps (256-byte rows) over a pair-layout weight blob, one per half-unit size: 64 rows (16 KB: the 128-channel layers),
// 16 rows (4 KB: layer 1), 8 rows (2 KB: the value head's block9), 32 rows (8 KB: the 64-column last layer of the backward chain)
struct alignas(64) TrunkMaps {
    CUtensorMap m[4];
};
constexpr int kMaxLayers = 9;
constexpr int kEpiThreads = 512;                  // warps 0-15: four threads per tile row, 16 of a pass's 64 output channels each
constexpr int kIssuerWarp = kEpiThreads / 32;     // warp 16
constexpr int kThreads = kEpiThreads + 128;        // + warp 16 MMA issuer, warps 17-19 weight producers
constexpr int kTmemCols = 512;                    // two 128-column fp32 accumulators + two for the FP8 cross terms (precision 2)
constexpr int kCrossCol = 256;                    // first TMEM column of the cross-term accumulators
constexpr int kA8Bytes = 8 * kGroupBytes;         // an FP8 activation tile: 8 groups of 16 channels = 25,600 B; A8 | AL8 share OFF_ALO

constexpr int OFF_AHI = 0;
constexpr int OFF_ALO = OFF_AHI + kActBytes;
constexpr int OFF_A1 = OFF_ALO + kActBytes;
constexpr int OFF_STAGE = OFF_A1 + kA1Bytes;
constexpr int OFF_BIAS = OFF_STAGE + kStages * kStageBytes;   // float [9][128]
constexpr int OFF_HEAD = OFF_BIAS + kMaxLayers * 128 * 4;      // float w9[128], b10[64], wfc[64]
constexpr int OFF_SCRATCH = OFF_HEAD + 256 * 4;                // float [4][128] partial dots + [2][64] logits
constexpr int OFF_BAR = OFF_SCRATCH + 768 * 4;                 // mbarriers + tmem pointer
constexpr int kSmemBytes = OFF_BAR + 256;
constexpr int kMaxStages = 6;                      // barrier slots: a CTA pair runs 6 stages of 16 KB (its half of a unit), a single CTA 3 of 32 KB

struct LayerDesc {
    int n_units;     // weight units (tap x 64-channel chunk), consumed in order
    int ksteps;      // K=16 MMA steps per unit
    int n;           // N of the MMA (output channels, padded)
    int unit_bytes;  // hi + lo
    int lo_off;      // byte offset of the lo block inside a unit
    int b_lbo;       // B operand: bytes between K-adjacent core matrices (= n * 16)
    int chunks;      // 64-channel chunks per tap (1 or 2); 0 for the explicit layer 1
    float cscale;    // precision 2: 1 / (2^11 * sw) as computed at load time (the kernel reads the slot's device copy, TrunkArgs::cscale)
};

struct NetDesc {
    int n_layers;    // 8 (policy) or 9 (value)
    int kind;        // 0 = SL policy, 1 = value
    LayerDesc layer[kMaxLayers];
    long long unit_base[kMaxLayers];  // byte offset of the layer's first unit in the weight blob
};

struct TrunkArgs {
    const u64 *p1, *p2;
    const uint8_t *color;
    long long n;          // positions
    float *out;           // policy: [n][64]; value: [n]
    int out_kind;         // policy: 0 = logits, 1 = softmax probabilities
    int precision;        // 1 = fp16 single pass, 3 = hi/lo split (3 MMAs), 2 = fp16 + FP8 cross terms (blob = the slot's second blob)
    const uint8_t *blob;  // packed weight units
    const float *bias;    // [n_layers][128]
    const float *head;    // policy: w9[128], b10[64]; value: b9 at [0], wfc[64] at [128..192)
    const int *n_dev;     // nullable: the live position count is min(n, *n_dev) (request lists built on the device, mcts.cu)
    const float *cscale;  // [n_layers] precision 2: factor that folds the FP8 cross-term accumulator in, 1 / (2^11 * sw_l) (device: a slot refresh rewrites it)
    float *dump[8];       // nullable each: post-ReLU output of block l+1 as fp32 [n][channels][64] (kept for the backward pass, reinforce.cu)
    // backward mode only (trunk_kernel<1>, the data-gradient chain of reinforce.cu):
    const float *dy_in;   // [n][128][64] gradient w.r.t. the output of block 8, entering the chain
    const float *mask[8]; // mask[i]: [n][N_i][64] forward activation whose sign gates the output of chain layer i (ReLU backward)
    unsigned *dymax[8];   // backward mode, nullable each: bit pattern of max |value written to dump[i]| (atomicMax; the weight-gradient
                          // kernel scales its fp16 operands by it — reinforce.cu used to read every gradient tensor again for this)
};

// hi/lo split (fp16 in the forward, bf16 in the backward chain: gradients need the exponent range) of 16 values of this
// thread's tile row, written as 2 channel groups of the activation tile.
template <bool BF16>
__device__ __forceinline__ void store_act16(uint8_t *smem, const float (&x)[16], uint32_t group0_off, bool split) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float f0 = x[q * 8 + 2 * e], f1 = x[q * 8 + 2 * e + 1];
            if (BF16) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
                const float2 hf = __bfloat1622float2(h);
                const __nv_bfloat162 l = __floats2bfloat162_rn(f0 - hf.x, f1 - hf.y);
                hw[e] = *reinterpret_cast<const uint32_t *>(&h);
                lw[e] = *reinterpret_cast<const uint32_t *>(&l);
            } else {
                const __half2 h = __floats2half2_rn(f0, f1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(f0 - hf.x, f1 - hf.y);
                hw[e] = *reinterpret_cast<const uint32_t *>(&h);
                lw[e] = *reinterpret_cast<const uint32_t *>(&l);
            }
        }
        const uint32_t off = group0_off + (uint32_t)q * kGroupBytes;
        *reinterpret_cast<uint4 *>(smem + OFF_AHI + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        if (split) *reinterpret_cast<uint4 *>(smem + OFF_ALO + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}

// precision 2: 16 values of this thread's tile row (as 8 packed pairs) -> fp16 hi parts (2 groups of 8 channels) and one 16-channel row
// of each FP8 tile: A8 = e4m3(hi), AL8 = e4m3((x - hi) * 2^11).  Two steps, so that the hi tile — all the next layer's fp16 MMAs need —
// can be published before the FP8 conversions are done.  Packed fp32x2 arithmetic and the fp16x2 -> e4m3x2 conversion keep the whole
// at 7 instructions per pair (the pass is issue-bound: it is what a layer boundary waits for).
__device__ __forceinline__ void store_act16_p2_hi(uint8_t *smem, const uint64_t (&x2)[8], int col0, uint32_t row_off, uint32_t (&hw)[8]) {
#pragma unroll
    for (int e = 0; e < 8; e++) {
        float f0, f1;
        f32x2_unpack(x2[e], f0, f1);
        const __half2 h = __floats2half2_rn(f0, f1);
        hw[e] = *reinterpret_cast<const uint32_t *>(&h);
    }
    *reinterpret_cast<uint4 *>(smem + OFF_AHI + (uint32_t)(col0 >> 3) * kGroupBytes + row_off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4 *>(smem + OFF_AHI + (uint32_t)((col0 >> 3) + 1) * kGroupBytes + row_off) = make_uint4(hw[4], hw[5], hw[6], hw[7]);
}
__device__ __forceinline__ void store_act16_p2_f8(uint8_t *smem, const uint64_t (&x2)[8], const uint32_t (&hw)[8], int col0, uint32_t row_off) {
    uint32_t h8[4], l8[4];
    const uint64_t k2048 = f32x2(2048.0f, 2048.0f), kneg = f32x2(-1.0f, -1.0f);
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hw[e]));
        const uint64_t lo = mul2(fma2(f32x2(hf.x, hf.y), kneg, x2[e]), k2048);   // (x - hi) * 2^11, exact until the FP8 rounding
        float l0, l1;
        f32x2_unpack(lo, l0, l1);
        const uint32_t a8 = e4m3x2_from_half2(hw[e]), b8 = e4m3x2_from_floats(l0, l1);
        if ((e & 1) == 0) { h8[e >> 1] = a8; l8[e >> 1] = b8; } else { h8[e >> 1] |= a8 << 16; l8[e >> 1] |= b8 << 16; }
    }
    const uint32_t off = (uint32_t)(col0 >> 4) * kGroupBytes + row_off;
    *reinterpret_cast<uint4 *>(smem + OFF_ALO + off) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
    *reinterpret_cast<uint4 *>(smem + OFF_ALO + kA8Bytes + off) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
}

template <int CG>
__device__ __forceinline__ void arrive_leader(uint32_t bar) {
    if (CG == 2) mbar_arrive_cluster(bar); else mbar_arrive(bar);
}

// Thread layout: warps 0-15 epilogue (warp w reads TMEM lanes 32*(w%4)..; quarter q = w / 4 takes columns [16q, 16q+16) of each
// 64-column pass), warp 16 = MMA issuer (also allocates TMEM), warps 17-19 = weight producers.
//
// Pipeline per tile (DESIGN.md "trunk pipeline"): the accumulator is double-buffered in TMEM (layer l uses buffer
// l&1) and weight units are ordered chunk-major, so the MMAs of layer l+1 on input channels 0..63 start as soon as
// the epilogue of layer l has written those channels, while it is still converting channels 64..127.
//
// MODE 0 = forward (SLPolicy / Value).  MODE 1 = the backward data-gradient chain of the REINFORCE update (reinforce.cu):
// chain layer i is the transposed, tap-flipped conv of block 8-i in bf16 hi/lo; the tile entering the chain is read from
// HBM (dy_in), every layer's output is gated by the sign of the forward activation (mask[i]) and written to HBM (dump[i])
// for the weight-gradient kernel as well as to the activation tile for the next chain layer.
//
// CG = 2: two CTAs of a cluster (one TPC) work as a pair on 4 boards: every MMA is cta_group::2 with M = 256 (128 rows of A = 2 boards per
// CTA) and each CTA holds only ITS HALF of a weight unit's output channels (rank 0: 0..N/2-1), loaded by 2-D tensor-map TMA whose bytes
// complete on the leader's "full" barrier.  The same 96 KB ring is then SIX units deep instead of three and each SM ingests half the
// weight bytes — the L2 -> shared-memory latency of a unit (≈ 1,000 cycles end to end) is what paced precision 2 with a 3-deep ring.
// The leader's warp 8 issues for the pair; commits are multicast to both CTAs; the epilogue threads of both CTAs arrive on the leader's
// "activations written" barriers (tools/micro/umma2_check.cu is the known-answer test of these mechanisms).
template <int MODE, int CG>
__global__ void __launch_bounds__(kThreads, 1) trunk_kernel(const __grid_constant__ TrunkArgs a, const __grid_constant__ NetDesc net,
                                                            const __grid_constant__ TrunkMaps maps) {
    static_assert(CG == 1 || CG == 2, "one CTA or a CTA pair per tile");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;          // 0 = leader
    constexpr int kRing = CG == 2 ? 6 : kStages;                      // ring stages
    constexpr int kRingBytes = kStageBytes / CG;                      // bytes per stage
    // The producer and issuer warps run their loops with all 32 lanes (warp-uniform control flow, one lane elected for the TMA / tcgen05
    // instructions): the compiler then keeps descriptors and barrier addresses in uniform registers.  With the role branch on a single
    // lane ((tid & 31) == 0) every UTCHMMA was wrapped in an R2UR + ELECT + BRA.U.ANY waterfall and the issuing thread, not the tensor
    // pipe, paced the layer (700 cycles per precision-2 unit against 512 of MMA time).  The shuffle makes `warp` provably uniform.
    const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * kMaxStages;
    const uint32_t bar_acc = bar_empty + 8 * kMaxStages;  // [2] accumulator buffer complete
    const uint32_t bar_act = bar_acc + 16;              // [2] input channels 0..63 / 64..127 of the next layer written
    const uint32_t bar_a1 = bar_act + 16;               // layer-1 im2col tile written, TMEM buffer 0 drained
    const uint32_t bar_hi = bar_a1 + 8;                 // precision 2: the fp16 hi tile of input channels 0..63 written (its FP8 tiles follow: bar_act[0])
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 8 * (2 * kMaxStages + 6));

    // ---- one-time setup
    for (int i = tid; i < (OFF_A1 + kA1Bytes) / 16; i += kThreads) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (MODE == 0) {
        float *sb = reinterpret_cast<float *>(smem + OFF_BIAS);
        for (int i = tid; i < net.n_layers * 128; i += kThreads) sb[i] = a.bias[i];
        float *sh = reinterpret_cast<float *>(smem + OFF_HEAD);
        for (int i = tid; i < 256; i += kThreads) sh[i] = a.head[i];
    }
    if (tid == 0) {
        for (int s = 0; s < kRing; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        mbar_init(bar_acc + 8, 1);
        // one arrival per epilogue warp (each thread fences its own writes, the warp synchronises, lane 0 arrives); pair: the warps of
        // both CTAs arrive on the leader's barriers
        mbar_init(bar_act, CG * (kEpiThreads / 32));
        mbar_init(bar_act + 8, CG * (kEpiThreads / 32));
        mbar_init(bar_a1, CG * (kEpiThreads / 32));
        mbar_init(bar_hi, CG * (kEpiThreads / 32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kIssuerWarp) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    fence_async_smem();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    long long n_pos = a.n;
    if (a.n_dev) n_pos = min(n_pos, (long long)__shfl_sync(0xFFFFFFFFu, *a.n_dev, 0));   // (the shuffle tells the compiler the bound is warp-uniform)
    const long long n_tiles = (n_pos + 2 * CG - 1) / (2 * CG);   // a tile = 2 boards per CTA (4 per pair)
    const long long tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;   // CTAs 2i, 2i+1 form cluster i
    const int L = net.n_layers;
    const bool split = a.precision >= 3;
    const bool p2 = MODE == 0 && a.precision == 2;   // fp16 main product + FP8 cross terms in the second accumulator

    if (warp > kIssuerWarp) {
        // ================= producers: stream weight units L2 -> smem ring, one warp per ring stage =================
        // A thread's wait-empty / expect_tx / cp.async.bulk round takes ≈ 0.3 µs whatever the copy size, and the rounds of ONE thread do
        // not overlap (tools/micro/tma_stream.cu: 74-100 GB/s per SM with one producer thread at any ring depth, 145 with two, 215 with
        // three) — so a single producer paced precision 1 / 2 at one unit per 0.3 µs.  Producer w owns ring stage w: it copies the units
        // g = w, w + 3, ... of the kernel-wide unit sequence (all 32 lanes run the loop, one elected lane issues).
        // Pair: producer w copies the units g = w, w + 3, ... into ring stage g mod 6 — this CTA's half of the unit, by tensor-map TMA.
        const uint32_t my_turn = (uint32_t)(warp - kIssuerWarp - 1);
        uint32_t stage = 0, phase = 0, turn = 0;
        for (long long tile = tile0, it = 0; tile < n_tiles; tile += tile_step, it++) {
            for (int l = 0; l < L; l++) {
                const uint8_t *src = a.blob + net.unit_base[l];
                const uint32_t unit_bytes = (uint32_t)net.layer[l].unit_bytes;
                const uint32_t bytes = (split || p2 || CG == 2) ? unit_bytes : (uint32_t)net.layer[l].lo_off;  // single pass needs hi only
                const int nu = net.layer[l].n_units;
                // pair: row (256 B) of this CTA's half of the layer's first unit, and the tensor map whose box is one half-unit
                int row = (int)((net.unit_base[l] + (long long)rank * (unit_bytes / 2)) >> 8);
                const void *map = &maps.m[net.layer[l].chunks == 0 ? 1 : net.layer[l].n == 16 ? 2 : net.layer[l].n == 64 ? 3 : 0];
                for (int u = 0; u < nu; u++) {
                    if (turn == my_turn) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        TRACEU(it, l, u, 0);
                        if (elect_one()) {
                            if (CG == 2) {
                                if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, bytes);   // both halves complete on the leader's barrier
                                tma2d_pair(sbase + OFF_STAGE + stage * kRingBytes, map, 0, row, bar_full + 8 * stage);
                            } else {
                                mbar_expect_tx(bar_full + 8 * stage, bytes);
                                bulk_g2s(sbase + OFF_STAGE + stage * kRingBytes, src, bytes, bar_full + 8 * stage);
                            }
                        }
                        TRACEU(it, l, u, 1);
                    }
                    src += unit_bytes;
                    row += (int)(unit_bytes >> 8);
                    if (++turn == 3) turn = 0;
                    if (++stage == kRing) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kIssuerWarp) {
        // ================= MMA issuer (pair: the leader's warp issues for both CTAs, the peer's warp 8 has nothing to do)
        // the whole warp runs the loop, one elected lane issues the MMAs and commits
        if (rank == 0) {
        uint32_t stage = 0, phase = 0, act_phase0 = 0, act_phase1 = 0, a1_phase = 0, hi_phase = 0;
        const uint32_t hiA = (uint32_t)(kRowPitch >> 4) | (1u << 14), hiB = (uint32_t)(128 >> 4) | (1u << 14);   // high words: SBO + version
        for (long long tile = tile0, it = 0; tile < n_tiles; tile += tile_step, it++) {
            for (int l = 0; l < L; l++) {
                const int ld_n = net.layer[l].n, ld_chunks = net.layer[l].chunks;
                const int ld_lo_off = net.layer[l].lo_off / CG, ld_b_lbo = net.layer[l].b_lbo / CG;   // pair: this CTA's half of the unit's N
                const uint32_t idesc = MODE == 0 ? instr_desc(kTileRows * CG, ld_n) : instr_desc_bf16(kTileRows * CG, ld_n);
                const uint32_t d_tmem = tmem + (uint32_t)(l & 1) * 128u;
                const uint32_t b_step = (uint32_t)(2 * ld_b_lbo) >> 4;              // descriptor units (16 B) per K=16 step
                const uint32_t b_lo_word = ((uint32_t)(ld_b_lbo >> 4) << 16);
                if (MODE == 0 && ld_chunks == 0) {
                    // layer 1: explicit im2col tile [4][128][16 B], LBO = 2048, SBO = 128; K = 32 = two steps
                    mbar_wait_cg<CG>(bar_a1, a1_phase);
                    a1_phase ^= 1;
                    mbar_wait_cg<CG>(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t bst = sbase + OFF_STAGE + stage * kRingBytes;
                    const uint32_t a1_hiw = (uint32_t)(128 >> 4) | (1u << 14);
                    const uint32_t aw = ((sbase + OFF_A1) >> 4) | ((uint32_t)((kTileRows * 16) >> 4) << 16);
                    const uint32_t bw = (bst >> 4) | b_lo_word, blw = ((bst + ld_lo_off) >> 4) | b_lo_word;
                    constexpr uint32_t dA1 = (2 * kTileRows * 16) >> 4;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 2; ks++) {
                            const uint64_t da = pack64(aw + ks * dA1, a1_hiw);
                            if (ks == 0) umma_f16_cg<CG>(d_tmem, da, pack64(bw, hiB), idesc, 0); else umma_f16_cg<CG>(d_tmem, da, pack64(bw + b_step, hiB), idesc, 1);
                            if (split || p2) umma_f16_cg<CG>(d_tmem, da, pack64(blw + ks * b_step, hiB), idesc, 1);  // inputs are exactly 0/1: no lo part
                        }
                        umma_commit_cg<CG>(bar_empty + 8 * stage);
                    }
                    if (++stage == kRing) { stage = 0; phase ^= 1; }
                } else {
                    // The units of the layer as one software-pipelined stream: the wait for the NEXT unit's weights (and, at a chunk
                    // boundary, for the next 64 input channels) is taken while this unit's last MMA group is still to be issued and
                    // the earlier ones are queued, and the commit that frees a ring stage is issued after the first MMA of the
                    // following unit — so the tensor pipe sees no gap between units.
                    const uint32_t a_lo_word = ((uint32_t)(kGroupBytes >> 4) << 16);
                    if (p2) {
                        mbar_wait_cg<CG>(bar_hi, hi_phase);   // the fp16 hi tile of input channels 0..63 is written
                        hi_phase ^= 1;
                    } else {
                        mbar_wait_cg<CG>(bar_act, act_phase0);  // input channels 0..63 are written
                        act_phase0 ^= 1;
                    }
                    tc_fence_after();
                    TRACE(it, l, 0);
                    mbar_wait_cg<CG>(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    uint32_t prev_bar = 0;   // ring stage barrier whose commit is still to be issued (0 = none)
                    uint32_t acc = 0;
                    int tap_start = 0;
                    if (p2) {
                        // Precision 2 at a layer boundary: the epilogue publishes the fp16 hi tile of channels 0..63 BEFORE it converts them
                        // to FP8, and the first two units are issued as  f16(u0) f16(u1) | FP8 tiles ready | f8(u0) f8(u1):  eight MMAs
                        // (512 cycles) cover the FP8 conversion.  Each accumulator still receives its products in the usual order.
                        const uint32_t st0 = stage;
                        uint32_t st1 = stage + 1, ph1 = phase;
                        if (st1 == kRing) { st1 = 0; ph1 ^= 1; }
                        uint32_t st2 = st1 + 1, ph2 = ph1;
                        if (st2 == kRing) { st2 = 0; ph2 ^= 1; }
                        const uint32_t b0 = sbase + OFF_STAGE + st0 * kRingBytes, b1 = sbase + OFF_STAGE + st1 * kRingBytes;
                        constexpr uint32_t dAq = (2 * kGroupBytes) >> 4;
                        const uint32_t ahw0 = ((sbase + OFF_AHI) >> 4) | a_lo_word, ahw1 = ((sbase + OFF_AHI + 16) >> 4) | a_lo_word;   // taps 0, 1
                        const uint32_t a8w0 = ((sbase + OFF_ALO) >> 4) | a_lo_word, al8w0 = ((sbase + OFF_ALO + kA8Bytes) >> 4) | a_lo_word;
                        const uint32_t a8w1 = ((sbase + OFF_ALO + 16) >> 4) | a_lo_word, al8w1 = ((sbase + OFF_ALO + kA8Bytes + 16) >> 4) | a_lo_word;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 4; ks++)
                                umma_f16_cg<CG>(d_tmem, pack64(ahw0 + ks * dAq, hiA), pack64(((b0 >> 4) | b_lo_word) + ks * b_step, hiB), idesc, ks > 0);
                            mbar_wait_cg<CG>(bar_full + 8 * st1, ph1);
                            tc_fence_after();
#pragma unroll
                            for (int ks = 0; ks < 4; ks++)
                                umma_f16_cg<CG>(d_tmem, pack64(ahw1 + ks * dAq, hiA), pack64(((b1 >> 4) | b_lo_word) + ks * b_step, hiB), idesc, 1);
                            mbar_wait_cg<CG>(bar_act, act_phase0);   // the FP8 tiles of channels 0..63
                            tc_fence_after();
#pragma unroll
                            for (int uu = 0; uu < 2; uu++) {
                                const uint32_t bb = uu ? b1 : b0, a8w = uu ? a8w1 : a8w0, al8w = uu ? al8w1 : al8w0;
                                const uint32_t w8 = ((bb + ld_lo_off) >> 4) | b_lo_word, wl8 = ((bb + ld_lo_off + (ld_lo_off >> 1)) >> 4) | b_lo_word;
#pragma unroll
                                for (int ks = 0; ks < 2; ks++) {
                                    umma_f8_cg<CG>(d_tmem + kCrossCol, pack64(al8w + ks * dAq, hiA), pack64(w8 + ks * b_step, hiB), idesc, (uu | ks) ? 1u : 0u);
                                    umma_f8_cg<CG>(d_tmem + kCrossCol, pack64(a8w + ks * dAq, hiA), pack64(wl8 + ks * b_step, hiB), idesc, 1);
                                }
                                if (uu == 0) {
                                    umma_commit_cg<CG>(bar_empty + 8 * st0);   // unit 0's ring stage
                                    mbar_wait_cg<CG>(bar_full + 8 * st2, ph2); // the weights of unit 2, for the loop below
                                    tc_fence_after();
                                }
                            }
                        }
                        act_phase0 ^= 1;
                        prev_bar = bar_empty + 8 * st1;
                        stage = st2; phase = ph2;
                        acc = 1;
                        tap_start = 2;
                    }
                    // Straight-line issue per unit: every MMA's descriptor is (a uniform-register low word + a constant, constant high word), and
                    // a unit is ONE elected region: a non-blocking test of the next unit's "full" barrier, the MMAs (with the deferred commit of the
                    // previous unit's ring stage after the first one) up to the last group, then the test's answer — a blocking wait only if the weights
                    // have not landed —, then the last MMA group.  (Two elected regions per unit with blocking waits between them cost ≈ 110 + 100 cycles more per unit.)
                    constexpr uint32_t dA = (2 * kGroupBytes) >> 4;
                    for (int chunk = 0; chunk < ld_chunks; chunk++) {
                        const uint32_t a_hi_base = sbase + OFF_AHI + (uint32_t)chunk * 8 * kGroupBytes, a_lo_base = sbase + OFF_ALO + (uint32_t)chunk * 8 * kGroupBytes;
                        const uint32_t a8_base = sbase + OFF_ALO + (uint32_t)chunk * 4 * kGroupBytes;   // FP8 tiles: 16 channels per group
                        const bool last_chunk = chunk + 1 == ld_chunks;
#pragma unroll 1
                        for (int tap = (chunk == 0 ? tap_start : 0); tap < 9; tap++) {   // (unrolling the taps made the issuer's loop ≈ 40 KB of code and 10-25 % slower: instruction fetch)
                            const int ky = tap / 3, kx = tap - ky * 3;  // (dy,dx) = (ky-1,kx-1); the halo sits at padded index 0
                            const uint32_t off = (uint32_t)(ky * 20 + kx) * 16;
                            const int u = chunk * 9 + tap;
                            (void)u;
                            TRACEU(it, l, u, 2);
                            const uint32_t bst = sbase + OFF_STAGE + stage * kRingBytes;
                            const uint32_t ahw = ((a_hi_base + off) >> 4) | a_lo_word, alw = ((a_lo_base + off) >> 4) | a_lo_word;
                            const uint32_t a8w = ((a8_base + off) >> 4) | a_lo_word, al8w = ((a8_base + kA8Bytes + off) >> 4) | a_lo_word;
                            const uint32_t bw = (bst >> 4) | b_lo_word, blw = ((bst + ld_lo_off) >> 4) | b_lo_word;
                            const uint32_t wl8w = ((bst + ld_lo_off + (ld_lo_off >> 1)) >> 4) | b_lo_word;   // precision 2: W8 sits at blw, WL8 after it
                            uint32_t nstage = stage + 1, nphase = phase;
                            if (nstage == kRing) { nstage = 0; nphase ^= 1; }
                            const bool more = !(tap == 8 && last_chunk);
                            if (elect_one()) {
                                // the next unit's weights: test now, look at the answer after this unit's MMAs are issued
                                const uint32_t next_ready = (CG == 2 && more) ? mbar_test_cg<CG>(bar_full + 8 * nstage, nphase) : 0u;   // (single CTA, 3-deep ring: the copy has rarely landed this early and the extra test measured 8 % slower)
                                // ---- first MMA group
                                if (split) {
#pragma unroll
                                    for (int ks = 0; ks < 3; ks++) {
                                        const uint64_t da = pack64(ahw + ks * dA, hiA), dal = pack64(alw + ks * dA, hiA);
                                        const uint64_t db = pack64(bw + ks * b_step, hiB), dbl = pack64(blw + ks * b_step, hiB);
                                        if (ks == 0) umma_f16_cg<CG>(d_tmem, da, db, idesc, acc); else umma_f16_cg<CG>(d_tmem, da, db, idesc, 1);
                                        umma_f16_cg<CG>(d_tmem, da, dbl, idesc, 1);
                                        umma_f16_cg<CG>(d_tmem, dal, db, idesc, 1);
                                        if (ks == 0 && prev_bar) umma_commit_cg<CG>(prev_bar);  // the previous unit's stage (covers these MMAs too: harmless)
                                    }
                                } else if (p2) {
                                    // FP8 tiles: 16 channels per 16-byte row, so a 64-channel chunk is 4 groups and one K = 32 MMA spans two of them —
                                    // the same descriptor stepping as fp16
#pragma unroll
                                    for (int ks = 0; ks < 4; ks++) {
                                        const uint64_t da = pack64(ahw + ks * dA, hiA), db = pack64(bw + ks * b_step, hiB);
                                        if (ks == 0) umma_f16_cg<CG>(d_tmem, da, db, idesc, acc); else umma_f16_cg<CG>(d_tmem, da, db, idesc, 1);
                                        if (ks == 0 && prev_bar) umma_commit_cg<CG>(prev_bar);
                                    }
                                    umma_f8_cg<CG>(d_tmem + kCrossCol, pack64(al8w, hiA), pack64(blw, hiB), idesc, acc);
                                    umma_f8_cg<CG>(d_tmem + kCrossCol, pack64(a8w, hiA), pack64(wl8w, hiB), idesc, 1);
                                } else {
#pragma unroll
                                    for (int ks = 0; ks < 3; ks++) {
                                        const uint64_t da = pack64(ahw + ks * dA, hiA), db = pack64(bw + ks * b_step, hiB);
                                        if (ks == 0) umma_f16_cg<CG>(d_tmem, da, db, idesc, acc); else umma_f16_cg<CG>(d_tmem, da, db, idesc, 1);
                                        if (ks == 0 && prev_bar) umma_commit_cg<CG>(prev_bar);
                                    }
                                }
                                // ---- the waits, while the MMAs above are queued: at the chunk boundary the next 64 input channels; the next unit's
                                // weights only if the early test failed.  The last MMA group is issued after them: it keeps the tensor pipe busy while
                                // this thread goes round the loop (checking after ALL the unit's MMAs measured 10 % slower with the 3-deep ring).
                                if (more) {
                                    if (tap == 8) mbar_wait_cg<CG>(bar_act + 8, act_phase1);
                                    if (!next_ready) mbar_wait_cg<CG>(bar_full + 8 * nstage, nphase);
                                    tc_fence_after();
                                }
                                // ---- last MMA group
                                if (split) {
                                    const uint64_t da = pack64(ahw + 3 * dA, hiA), dal = pack64(alw + 3 * dA, hiA);
                                    const uint64_t db = pack64(bw + 3 * b_step, hiB), dbl = pack64(blw + 3 * b_step, hiB);
                                    umma_f16_cg<CG>(d_tmem, da, db, idesc, 1);
                                    umma_f16_cg<CG>(d_tmem, da, dbl, idesc, 1);
                                    umma_f16_cg<CG>(d_tmem, dal, db, idesc, 1);
                                } else if (p2) {
                                    umma_f8_cg<CG>(d_tmem + kCrossCol, pack64(al8w + dA, hiA), pack64(blw + b_step, hiB), idesc, 1);
                                    umma_f8_cg<CG>(d_tmem + kCrossCol, pack64(a8w + dA, hiA), pack64(wl8w + b_step, hiB), idesc, 1);
                                } else {
                                    umma_f16_cg<CG>(d_tmem, pack64(ahw + 3 * dA, hiA), pack64(bw + 3 * b_step, hiB), idesc, 1);
                                }
                            }
                            if (tap == 8 && more) {
                                act_phase1 ^= 1;
                                TRACE(it, l, 1);
                            }
                            acc = 1;
                            TRACEU(it, l, u, 3);
                            prev_bar = bar_empty + 8 * stage;
                            stage = nstage; phase = nphase;
                        }
                    }
                    if (elect_one()) umma_commit_cg<CG>(prev_bar);  // frees the last weight stage of the layer when its MMAs retire
                }
                if (elect_one()) umma_commit_cg<CG>(bar_acc + 8 * (l & 1));  // accumulator of this layer complete
                TRACE(it, l, 2);
            }
        }
        }
    } else {
        // ================= epilogue warps: thread pair (m, m+128) owns tile row m = TMEM lane m =================
        const int m = tid & 127, qt = tid >> 7;    // quarter qt: columns [16 qt, 16 qt + 16) of each 64-column pass
        const int g = m >> 3, c = m & 7, r = g >> 1, b = g & 1;
        const uint32_t row_off = (uint32_t)(((r + 1) * 2 + b) * 10 + (c + 1)) * 16;  // interior cell of the padded tile
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const float *sbias = reinterpret_cast<const float *>(smem + OFF_BIAS);
        const float *shead = reinterpret_cast<const float *>(smem + OFF_HEAD);
        float *scratch = reinterpret_cast<float *>(smem + OFF_SCRATCH);   // [4][128] partial dots / [2][64] logits
        const int cell = r * 8 + c;
        uint32_t acc_phase = 0;   // bit b: parity of accumulator buffer b's barrier
        // the barriers the issuer waits on are the leader's: a pair's peer CTA arrives on them through the cluster address
        const uint32_t act_leader = CG == 2 ? mapa_rank(bar_act, 0) : bar_act, a1_leader = CG == 2 ? mapa_rank(bar_a1, 0) : bar_a1;
        const uint32_t hi_leader = CG == 2 ? mapa_rank(bar_hi, 0) : bar_hi;
        // layer-1 input of position px: explicit im2col of the two bit planes, k = tap*2 + channel (0 = opponent, 1 = mover).  The tile
        // lives in its own region (OFF_A1), free again as soon as the layer-1 MMAs that read it have retired, so the NEXT tile's boards
        // are fetched and expanded while the current tile's last layer is on the tensor pipe: the issuer can put the next tile's layer 1
        // behind that layer without waiting for an epilogue (≈ 5,000 cycles stood between two tiles before).
        auto build_a1 = [&](long long px) {
            if (qt == 0) {
                u64 own = 0, opp = 0;
                if (px < n_pos) {
                    const bool first = a.color[px] == 1;
                    const u64 x1 = a.p1[px], x2 = a.p2[px];
                    own = first ? x1 : x2;
                    opp = first ? x2 : x1;
                }
                uint32_t bits = 0;
#pragma unroll
                for (int t = 0; t < 9; t++) {
                    const int y = r + t / 3 - 1, x = c + t % 3 - 1;
                    if (y >= 0 && y < 8 && x >= 0 && x < 8) {
                        const int k = y * 8 + x;
                        bits |= (uint32_t)((opp >> k) & 1) << (2 * t);
                        bits |= (uint32_t)((own >> k) & 1) << (2 * t + 1);
                    }
                }
#pragma unroll
                for (int kg = 0; kg < 4; kg++) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const uint32_t lo = (bits >> (kg * 8 + 2 * e)) & 1, hi = (bits >> (kg * 8 + 2 * e + 1)) & 1;
                        w[e] = (lo ? 0x3C00u : 0u) | (hi ? 0x3C000000u : 0u);  // fp16 1.0 = 0x3C00
                    }
                    *reinterpret_cast<uint4 *>(smem + OFF_A1 + kg * (kTileRows * 16) + m * 16) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            fence_async_smem();
        };
        // "im2col tile written AND accumulator buffer 0 drained by this warp": buffer 0 holds the layers with even index, so when the last
        // layer's index is odd the arrival can precede its epilogue (the last even layer's accumulator was read before)
        auto arrive_a1 = [&]() {
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) arrive_leader<CG>(a1_leader);
        };
        const bool a1_early = ((L - 1) & 1) == 1;
        for (long long tile = tile0, it = 0; tile < n_tiles; tile += tile_step, it++) {
            const long long pos = (tile * CG + rank) * 2 + b;
            const bool valid = pos < n_pos;
            const bool has_next = tile + tile_step < n_tiles;
            if (MODE == 1) {
                // ---- chain entry: the gradient tile [128 channels][cell] of this row's position -> bf16 hi/lo activation tile
                const float *src = a.dy_in + (size_t)pos * 128 * 64 + cell;
#pragma unroll 1
                for (int ps = 0; ps < 2; ps++) {
                    const int col0 = ps * 64 + qt * 16;
                    float x[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) x[j] = valid ? __ldg(src + (size_t)(col0 + j) * 64) : 0.0f;
                    store_act16<true>(smem, x, (uint32_t)(col0 >> 3) * kGroupBytes + row_off, split);
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if ((tid & 31) == 0) arrive_leader<CG>(act_leader + 8 * ps);
                }
            } else {
            // ---- layer-1 input: the first tile's im2col tile is built here; every later one was built while the previous tile's last
            //      layer ran (build_a1 / arrive_a1 below)
                if (it == 0) {
                    build_a1(pos);
                    arrive_a1();
                }
            }

            for (int l = 0; l < L; l++) {
                // the layer's constants are fetched BEFORE the wait (indexed constant loads: ≈ 500 cycles after it otherwise, on the path the
                // next layer's first MMA waits for)
                LayerDesc ld;
                ld.n = net.layer[l].n;
                ld.cscale = (MODE == 0 && a.cscale) ? __ldg(a.cscale + l) : 0.0f;
                float *const dump_l = MODE == 0 ? a.dump[l] : nullptr;
                asm volatile("" ::"r"(ld.n), "f"(ld.cscale), "l"(dump_l));
                if (MODE == 0 && l == L - 1 && has_next) {
                    build_a1(((tile + tile_step) * CG + rank) * 2 + b);
                    if (a1_early) arrive_a1();
                }
                mbar_wait(bar_acc + 8 * (l & 1), (acc_phase >> (l & 1)) & 1u);
                acc_phase ^= 1u << (l & 1);
                tc_fence_after();
                if (tid == 0) TRACE(it, l, 3);
                const uint32_t t_addr = lane_addr + (uint32_t)(l & 1) * 128u;
                const bool policy_head = (l == 7 && net.kind == 0);
                const bool writes_act = (l + 1 < L);
                if (MODE == 1) {
                    const int passes = ld.n / 64;
                    const float *mk = a.mask[l] + ((size_t)pos * ld.n) * 64 + cell;
                    float *dst = a.dump[l] + ((size_t)pos * ld.n) * 64 + cell;
#pragma unroll 1
                    for (int ps = 0; ps < passes; ps++) {
                        const int col0 = ps * 64 + qt * 16;
                        float g[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) g[j] = valid ? __ldg(mk + (size_t)(col0 + j) * 64) : 0.0f;   // in flight during the TMEM read
                        uint32_t v[16];
                        tmem_ld16(t_addr + col0, v);
                        tmem_wait_ld();
                        float x[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) x[j] = g[j] > 0.0f ? __uint_as_float(v[j]) : 0.0f;
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; j++) dst[(size_t)(col0 + j) * 64] = x[j];
                        }
                        if (a.dymax[l] != nullptr) {   // (x = 0 for rows beyond the batch)
                            float mx = 0.0f;
#pragma unroll
                            for (int j = 0; j < 16; j++) mx = fmaxf(mx, fabsf(x[j]));
                            const unsigned wmx = __reduce_max_sync(0xFFFFFFFFu, mx < 3.0e38f ? __float_as_uint(mx) : 0u);   // NaN / inf stay out
                            if ((tid & 31) == 0 && wmx != 0u) atomicMax(a.dymax[l], wmx);
                        }
                        if (writes_act) {
                            store_act16<true>(smem, x, (uint32_t)(col0 >> 3) * kGroupBytes + row_off, split);
                            fence_async_smem();
                            tc_fence_before();
                            __syncwarp();
                            if ((tid & 31) == 0) arrive_leader<CG>(act_leader + 8 * ps);
                        }
                    }
                } else if (l < 8) {
                    float dot = 0.0f;
                    const int passes = ld.n / 64;  // 1 for the 64-channel layer, else 2
                    for (int ps = 0; ps < passes; ps++) {
                        const int col0 = ps * 64 + qt * 16;
                        uint32_t v[16];
                        if (ps == 0) TRACEE(it, l, 0);
                        tmem_ld16(t_addr + col0, v);
                        float x[16];
                        if (p2 && l > 0) {   // fold the FP8 cross-term accumulator in (layer 1 has none: its inputs are exactly 0 / 1)
                            uint32_t vx[16];
                            tmem_ld16(t_addr + kCrossCol + col0, vx);
                            tmem_wait_ld();
                            const uint64_t cs2 = f32x2(ld.cscale, ld.cscale);
                            const float2 *b2 = reinterpret_cast<const float2 *>(sbias + l * 128 + col0);
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const float2 bb = b2[j];
                                const uint64_t y = add2(fma2(f32x2(__uint_as_float(vx[2 * j]), __uint_as_float(vx[2 * j + 1])), cs2,
                                                             f32x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]))), f32x2(bb.x, bb.y));
                                float y0, y1;
                                f32x2_unpack(y, y0, y1);
                                x[2 * j] = fmaxf(y0, 0.0f);
                                x[2 * j + 1] = fmaxf(y1, 0.0f);
                            }
                        } else {
                            tmem_wait_ld();
#pragma unroll
                            for (int j = 0; j < 16; j++) x[j] = fmaxf(__uint_as_float(v[j]) + sbias[l * 128 + col0 + j], 0.0f);
                        }
                        if (ps == 0) TRACEE(it, l, 1);
                        if (policy_head) {
#pragma unroll
                            for (int j = 0; j < 16; j++) dot = fmaf(x[j], shead[col0 + j], dot);
                        }
                        if (dump_l != nullptr && valid) {
                            float *dst = dump_l + ((size_t)pos * ld.n + col0) * 64 + cell;
#pragma unroll
                            for (int j = 0; j < 16; j++) dst[(size_t)j * 64] = x[j];
                        }
                        if (writes_act) {
                            if (p2) {
                                uint64_t x2[8];
                                uint32_t hw[8];
#pragma unroll
                                for (int j = 0; j < 8; j++) x2[j] = f32x2(x[2 * j], x[2 * j + 1]);
                                store_act16_p2_hi(smem, x2, col0, row_off, hw);
                                if (ps == 0) {   // channels 0..63: the hi tile is all the next layer's first fp16 MMAs need
                                    fence_async_smem();
                                    __syncwarp();
                                    if ((tid & 31) == 0) arrive_leader<CG>(hi_leader);
                                }
                                store_act16_p2_f8(smem, x2, hw, col0, row_off);
                            } else store_act16<false>(smem, x, (uint32_t)(col0 >> 3) * kGroupBytes + row_off, split);
                            if (ps == 0) TRACEE(it, l, 2);
                            fence_async_smem();
                            if (ps == 0) TRACEE(it, l, 3);
                            tc_fence_before();
                            __syncwarp();
                            if ((tid & 31) == 0) arrive_leader<CG>(act_leader + 8 * ps);  // input channels [64*ps, 64*ps+64) of the next layer are in place
                            if (ps == 0) TRACEE(it, l, 4);
                            if (tid == 0) TRACE(it, l, 4 + ps);
                        }
                    }
                    if (policy_head) {
                        // policy head: conv9 (1x1, no bias) + bias10[cell] (network.py:44-46); the row's four threads hold a quarter each
                        scratch[qt * 128 + m] = dot;
                        asm volatile("bar.sync 1, 512;" ::: "memory");
                        if (qt == 0) {
                            const float logit = ((scratch[m] + scratch[128 + m]) + (scratch[256 + m] + scratch[384 + m])) + shead[128 + cell];
                            if (a.out_kind == 0) {
                                if (valid) a.out[pos * 64 + cell] = logit;
                            } else {
                                float *lg = scratch + 512;
                                lg[b * 64 + cell] = logit;
                                asm volatile("bar.sync 2, 128;" ::: "memory");
                                float mx = -3.0e38f;
                                for (int i = 0; i < 64; i++) mx = fmaxf(mx, lg[b * 64 + i]);
                                float sum = 0.0f;
                                for (int i = 0; i < 64; i++) sum += __expf(lg[b * 64 + i] - mx);
                                if (valid) a.out[pos * 64 + cell] = __expf(logit - mx) / sum;
                            }
                        }
                        asm volatile("bar.sync 1, 512;" ::: "memory");  // scratch is reused by the next tile
                    }
                } else {
                    // value head: relu(block9 + b9) . (fc11 * fc10)  (network.py:92-95, dropout off)
                    if (qt == 0) {
                        uint32_t v, vx = 0;
                        tmem_ld1(t_addr, v);
                        if (p2) tmem_ld1(t_addr + kCrossCol, vx);
                        tmem_wait_ld();
                        const float h = fmaxf(fmaf(__uint_as_float(vx), p2 ? ld.cscale : 0.0f, __uint_as_float(v)) + shead[0], 0.0f);
                        scratch[b * 64 + cell] = h * shead[128 + cell];
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                        if (warp < 2) {
                            float s = scratch[warp * 64 + (tid & 31)] + scratch[warp * 64 + 32 + (tid & 31)];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
                            const long long p = (tile * CG + rank) * 2 + warp;
                            if ((tid & 31) == 0 && p < n_pos) a.out[p] = s;
                        }
                        asm volatile("bar.sync 2, 128;" ::: "memory");
                    }
                }
                tc_fence_before();
            }
            if (MODE == 0 && has_next && !a1_early) arrive_a1();   // an even last layer sits in buffer 0: drained only now
        }
    }

    // ---- teardown
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: no CTA leaves (or frees TMEM) while the other may still be addressed
    if (warp == kIssuerWarp) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------- host side: packing + launch

struct NetSlot {
    bool loaded = false;
    NetDesc desc;
    uint8_t *d_blob = nullptr;
    uint8_t *d_blob2 = nullptr;   // precision 2: fp16 hi block | W8 | WL8 per unit (nullptr after a device-side refresh: precision 2 then runs as 3)
    // CTA-pair layout of the same two blobs (every unit as [rank 0's half of the output channels | rank 1's half]) and their tensor maps;
    // nullptr after a device-side refresh: the slot then runs on the single-CTA kernel
    uint8_t *d_pair = nullptr, *d_pair2 = nullptr;
    TrunkMaps maps, maps2;
    float *d_cscale = nullptr;    // [kMaxLayers]
    float *d_bias = nullptr;
    float *d_head = nullptr;
};

struct TrunkState {
    NetSlot slot[IAGO_NET_SLOTS];
    bool attr_set = false;
    const uint8_t *bwd_pair = nullptr;   // the pair-layout backward blob bwd_maps was built for
    TrunkMaps bwd_maps;
    float *d_sw = nullptr;   // [kMaxLayers] scratch of trunk_refresh_slot: the FP8 weight scale per layer
};

static TrunkState *state(iago_ctx *ctx) {
    if (!ctx->trunk) ctx->trunk = new TrunkState();
    return static_cast<TrunkState *>(ctx->trunk);
}
static float *st_scratch(iago_ctx *ctx) {
    TrunkState *st = state(ctx);
    if (!st->d_sw && cudaMalloc(&st->d_sw, kMaxLayers * sizeof(float)) != cudaSuccess) {
        set_error("cudaMalloc failed for the weight-scale scratch");
        return nullptr;
    }
    return st->d_sw;
}

bool trunk_slot_holds(iago_ctx *ctx, int slot, int kind) {
    if (slot < 0 || slot >= IAGO_NET_SLOTS || !ctx->trunk) return false;
    const NetSlot &s = static_cast<TrunkState *>(ctx->trunk)->slot[slot];
    return s.loaded && s.desc.kind == kind;
}

void trunk_destroy(iago_ctx *ctx) {
    if (!ctx->trunk) return;
    TrunkState *st = static_cast<TrunkState *>(ctx->trunk);
    cudaFree(st->d_sw);
    for (auto &s : st->slot) {
        cudaFree(s.d_blob);
        cudaFree(s.d_blob2);
        cudaFree(s.d_pair);
        cudaFree(s.d_pair2);
        cudaFree(s.d_cscale);
        cudaFree(s.d_bias);
        cudaFree(s.d_head);
    }
    delete st;
    ctx->trunk = nullptr;
}

static inline void split_half(float w, uint16_t &hi, uint16_t &lo) {
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
}

// One unit = hi block [kgroups][n_pad][8] followed by the lo block.  get(n, k) returns the fp32 weight of output n,
// reduction index k (0 <= k < kgroups*8) or 0 for padding.
template <class F>
static void pack_unit(std::vector<uint8_t> &blob, int kgroups, int n_pad, F get) {
    const size_t half_elems = (size_t)kgroups * n_pad * 8;
    const size_t base = blob.size();
    blob.resize(base + half_elems * 4);
    uint16_t *hi = reinterpret_cast<uint16_t *>(blob.data() + base);
    uint16_t *lo = hi + half_elems;
    for (int kg = 0; kg < kgroups; kg++)
        for (int n = 0; n < n_pad; n++)
            for (int e = 0; e < 8; e++) split_half(get(n, kg * 8 + e), hi[((size_t)kg * n_pad + n) * 8 + e], lo[((size_t)kg * n_pad + n) * 8 + e]);
}

// Precision-2 unit: the same fp16 hi block, then W8 = e4m3(w_hi * sw) and WL8 = e4m3(w_lo * 2^11 * sw), both [kgroups / 2][n_pad][16]
// (16 reduction indices per 16-byte row), each half the size of the hi block.  w_hi / w_lo are the fp16 parts of pack_unit.
template <class F>
static void pack_unit_p2(std::vector<uint8_t> &blob, int kgroups, int n_pad, float sw, F get) {
    const size_t half_elems = (size_t)kgroups * n_pad * 8;
    const size_t base = blob.size();
    blob.resize(base + half_elems * 4);
    uint16_t *hi = reinterpret_cast<uint16_t *>(blob.data() + base);
    uint8_t *w8 = blob.data() + base + half_elems * 2, *wl8 = w8 + half_elems;
    for (int kg = 0; kg < kgroups; kg++)
        for (int n = 0; n < n_pad; n++)
            for (int e = 0; e < 8; e++) {
                uint16_t h, l;
                split_half(get(n, kg * 8 + e), h, l);
                hi[((size_t)kg * n_pad + n) * 8 + e] = h;
                const int k = kg * 8 + e;
                const size_t o8 = ((size_t)(k >> 4) * n_pad + n) * 16 + (k & 15);
                w8[o8] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(__ushort_as_half(h)) * sw, __NV_SATFINITE, __NV_E4M3);
                wl8[o8] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(__ushort_as_half(l)) * 2048.0f * sw, __NV_SATFINITE, __NV_E4M3);
            }
}

// Pair layout: the unit as two halves of its output channels, each a complete (hi | lo) or (hi | W8 | WL8) unit of n_pad / 2 columns.
template <class F>
static void pack_unit_pair(std::vector<uint8_t> &blob, int kgroups, int n_pad, F get) {
    for (int h = 0; h < 2; h++) pack_unit(blob, kgroups, n_pad / 2, [&](int n, int k) { return get(h * (n_pad / 2) + n, k); });
}
template <class F>
static void pack_unit_p2_pair(std::vector<uint8_t> &blob, int kgroups, int n_pad, float sw, F get) {
    for (int h = 0; h < 2; h++) pack_unit_p2(blob, kgroups, n_pad / 2, sw, [&](int n, int k) { return get(h * (n_pad / 2) + n, k); });
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Tensor maps over a pair-layout blob seen as rows of 256 bytes; box = one half-unit (64, 16 or 8 rows).
static int make_maps(uint8_t *d_blob, size_t bytes, TrunkMaps &out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return IAGO_E_CUDA;
        }
    }
    const cuuint32_t rows[4] = {64, 16, 8, 32};
    for (int i = 0; i < 4; i++) {
        const cuuint64_t gdim[2] = {256, (cuuint64_t)(bytes / 256)}, gstride[1] = {256};
        const cuuint32_t box[2] = {256, rows[i]}, estr[2] = {1, 1};
        const CUresult r = encode(&out.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_blob, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
            return IAGO_E_CUDA;
        }
    }
    return IAGO_OK;
}

// sw: the power of two that brings the layer's largest |weight| into [128, 256) (E4M3 tops out at 448)
static float fp8_weight_scale(const float *w, size_t count) {
    float mx = 0.0f;
    for (size_t i = 0; i < count; i++) mx = fmaxf(mx, fabsf(w[i]));
    if (!(mx > 0.0f) || !isfinite(mx)) return 1.0f;
    return ldexpf(1.0f, 7 - ilogbf(mx));
}

}  // namespace iago

using namespace iago;

#ifdef IAGO_TRUNK_TRACE
extern "C" __attribute__((visibility("default"))) int iago_debug_trace(unsigned long long *out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_trace, (size_t)n * 8);
}
#endif

extern "C" {

int iago_load_net(iago_ctx *ctx, int slot, int kind, const float *params, int64_t n_floats) {
    IAGO_REQUIRE(ctx && params, "NULL argument");
    IAGO_REQUIRE(slot >= 0 && slot < IAGO_NET_SLOTS, "slot out of range (0..IAGO_NET_SLOTS-1)");
    IAGO_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (SL policy) or 1 (value)");
    const int cin[8] = {2, 64, 128, 128, 128, 128, 128, 128}, cout[8] = {64, 128, 128, 128, 128, 128, 128, 128};
    int64_t need = 0;
    for (int l = 0; l < 8; l++) need += (int64_t)cout[l] * cin[l] * 9 + cout[l];
    need += kind == 0 ? 128 + 64 : 1152 + 1 + 128 * 64 + 128;
    if (n_floats != need) {
        set_error("iago_load_net: expected %lld floats for kind %d, got %lld", (long long)need, kind, (long long)n_floats);
        return IAGO_E_INVALID;
    }
    DeviceGuard guard(ctx->device);
    NetSlot &s = state(ctx)->slot[slot];
    NetDesc d;
    memset(&d, 0, sizeof d);
    d.kind = kind;
    d.n_layers = kind == 0 ? 8 : 9;
    std::vector<uint8_t> blob, blob2, pair, pair2;
    std::vector<float> bias((size_t)kMaxLayers * 128, 0.0f), head(256, 0.0f);
    const float *p = params;
    for (int l = 0; l < 8; l++) {
        const float *W = p, *B = p + (size_t)cout[l] * cin[l] * 9;
        p = B + cout[l];
        for (int n = 0; n < cout[l]; n++) bias[(size_t)l * 128 + n] = B[n];
        LayerDesc &ld = d.layer[l];
        d.unit_base[l] = (long long)blob.size();
        ld.n = cout[l];
        ld.b_lbo = ld.n * 16;
        if (l == 0) {
            // explicit im2col: k = tap*2 + channel, K padded 18 -> 32
            ld.n_units = 1; ld.ksteps = 2; ld.chunks = 0;
            auto w1 = [&](int n, int k) { return k < 18 ? W[((size_t)n * 2 + (k & 1)) * 9 + (k >> 1)] : 0.0f; };
            pack_unit(blob, 4, ld.n, w1);
            pack_unit(blob2, 4, ld.n, w1);  // layer 1: fp16 hi / lo in both blobs
            pack_unit_pair(pair, 4, ld.n, w1);
            pack_unit_pair(pair2, 4, ld.n, w1);
            ld.cscale = 0.0f;
            ld.lo_off = 4 * ld.n * 16;
        } else {
            ld.chunks = cin[l] / 64; ld.n_units = 9 * ld.chunks; ld.ksteps = 4;
            const float sw = fp8_weight_scale(W, (size_t)cout[l] * cin[l] * 9);
            ld.cscale = 1.0f / (2048.0f * sw);
            for (int ch = 0; ch < ld.chunks; ch++)  // chunk-major: all taps of input channels 0..63 first
                for (int tap = 0; tap < 9; tap++) {
                    auto wl = [&](int n, int k) { return W[((size_t)n * cin[l] + ch * 64 + k) * 9 + tap]; };
                    pack_unit(blob, 8, ld.n, wl);
                    pack_unit_p2(blob2, 8, ld.n, sw, wl);
                    pack_unit_pair(pair, 8, ld.n, wl);
                    pack_unit_p2_pair(pair2, 8, ld.n, sw, wl);
                }
            ld.lo_off = 8 * ld.n * 16;
        }
        ld.unit_bytes = 2 * ld.lo_off;
    }
    if (kind == 0) {
        for (int i = 0; i < 128; i++) head[i] = p[i];          // conv9/W [1][128][1][1]
        for (int i = 0; i < 64; i++) head[128 + i] = p[128 + i];  // bias10/b
    } else {
        const float *W9 = p, *b9 = p + 1152, *fc10 = b9 + 1, *fc11 = fc10 + 128 * 64;
        LayerDesc &ld = d.layer[8];
        d.unit_base[8] = (long long)blob.size();
        ld.n = 16; ld.b_lbo = 16 * 16; ld.chunks = 2; ld.n_units = 18; ld.ksteps = 4;
        const float sw9 = fp8_weight_scale(W9, 1152);
        ld.cscale = 1.0f / (2048.0f * sw9);
        for (int ch = 0; ch < 2; ch++)
            for (int tap = 0; tap < 9; tap++) {
                auto w9 = [&](int n, int k) { return n == 0 ? W9[(size_t)(ch * 64 + k) * 9 + tap] : 0.0f; };
                pack_unit(blob, 8, 16, w9);
                pack_unit_p2(blob2, 8, 16, sw9, w9);
                pack_unit_pair(pair, 8, 16, w9);
                pack_unit_p2_pair(pair2, 8, 16, sw9, w9);
            }
        ld.lo_off = 8 * 16 * 16;
        ld.unit_bytes = 2 * ld.lo_off;
        head[0] = b9[0];
        // fc11 (1x128) * fc10 (128x64): no nonlinearity between them at inference (dropout is identity), collapse in fp64
        for (int j = 0; j < 64; j++) {
            double acc = 0.0;
            for (int i = 0; i < 128; i++) acc += (double)fc11[i] * (double)fc10[(size_t)i * 64 + j];
            head[128 + j] = (float)acc;
        }
    }
    if (blob2.size() != blob.size() || pair.size() != blob.size() || pair2.size() != blob.size()) {
        set_error("iago_load_net: internal error, the two weight blobs differ in size");
        return IAGO_E_STATE;
    }
    cudaFree(s.d_blob); cudaFree(s.d_blob2); cudaFree(s.d_pair); cudaFree(s.d_pair2); cudaFree(s.d_cscale); cudaFree(s.d_bias); cudaFree(s.d_head);
    s = NetSlot();
    {
        float cs[kMaxLayers];
        for (int l = 0; l < kMaxLayers; l++) cs[l] = d.layer[l].cscale;
        IAGO_CUDA(cudaMalloc(&s.d_cscale, sizeof cs));
        IAGO_CUDA(cudaMemcpy(s.d_cscale, cs, sizeof cs, cudaMemcpyHostToDevice));
    }
    IAGO_CUDA(cudaMalloc(&s.d_pair, pair.size()));
    IAGO_CUDA(cudaMalloc(&s.d_pair2, pair2.size()));
    IAGO_CUDA(cudaMemcpy(s.d_pair, pair.data(), pair.size(), cudaMemcpyHostToDevice));
    IAGO_CUDA(cudaMemcpy(s.d_pair2, pair2.data(), pair2.size(), cudaMemcpyHostToDevice));
    if (int rc = make_maps(s.d_pair, pair.size(), s.maps)) return rc;
    if (int rc = make_maps(s.d_pair2, pair2.size(), s.maps2)) return rc;
    IAGO_CUDA(cudaMalloc(&s.d_blob, blob.size()));
    IAGO_CUDA(cudaMalloc(&s.d_blob2, blob2.size()));
    IAGO_CUDA(cudaMemcpy(s.d_blob2, blob2.data(), blob2.size(), cudaMemcpyHostToDevice));
    IAGO_CUDA(cudaMalloc(&s.d_bias, bias.size() * 4));
    IAGO_CUDA(cudaMalloc(&s.d_head, head.size() * 4));
    IAGO_CUDA(cudaMemcpy(s.d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    IAGO_CUDA(cudaMemcpy(s.d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    IAGO_CUDA(cudaMemcpy(s.d_head, head.data(), head.size() * 4, cudaMemcpyHostToDevice));
    s.desc = d;
    s.loaded = true;
    return IAGO_OK;
}

}  // extern "C"

namespace iago {
static int set_trunk_attrs(TrunkState *st) {
    if (st->attr_set) return IAGO_OK;
    IAGO_CUDA(cudaFuncSetAttribute(trunk_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    IAGO_CUDA(cudaFuncSetAttribute(trunk_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    IAGO_CUDA(cudaFuncSetAttribute(trunk_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    IAGO_CUDA(cudaFuncSetAttribute(trunk_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    st->attr_set = true;
    return IAGO_OK;
}

int trunk_launch(iago_ctx *ctx, int slot, int want_kind, const uint64_t *p1, const uint64_t *p2, const uint8_t *color,
                 int64_t n, float *out, int out_kind, int precision, void *stream, const int *n_dev, float *const *dump) {
    IAGO_REQUIRE(ctx && p1 && p2 && color && out, "NULL argument");
    IAGO_REQUIRE(slot >= 0 && slot < IAGO_NET_SLOTS, "slot out of range (0..IAGO_NET_SLOTS-1)");
    IAGO_REQUIRE(n >= 0, "n < 0");
    IAGO_REQUIRE(precision >= 1 && precision <= 3, "precision must be 1 (fp16), 2 (fp16 + FP8 cross terms) or 3 (fp16 hi/lo split)");
    TrunkState *st = state(ctx);
    NetSlot &s = st->slot[slot];
    if (!s.loaded || s.desc.kind != want_kind) {
        set_error("net slot %d holds no %s network (call iago_load_net)", slot, want_kind == 0 ? "policy" : "value");
        return IAGO_E_STATE;
    }
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    if (int rc = set_trunk_attrs(st)) return rc;
    // CTA pairs (4 boards per cluster) whenever the slot has the pair-layout blobs — the trainer's forward with its activation dumps
    // included (each CTA streams half of a weight unit from L2: 52.6 against 53.5 ms of gradient time per 2,048-game set);
    // IAGO_TRUNK_CG1=1 forces the single-CTA kernel, IAGO_TRUNK_DUMP_CG1=1 only for launches that dump (A/B measurements)
    static const bool force_cg1 = getenv("IAGO_TRUNK_CG1") != nullptr;
    static const bool dump_cg1 = getenv("IAGO_TRUNK_DUMP_CG1") != nullptr;   // A/B: the trainer's forward on single CTAs
    const bool pairs = s.d_pair && !(dump && dump_cg1) && !force_cg1;
    const long long tiles = pairs ? (n + 3) / 4 : (n + 1) / 2;
    const long long max_groups = pairs ? ctx->sm_count / 2 : ctx->sm_count;
    const int grid = (int)(tiles < max_groups ? tiles : max_groups) * (pairs ? 2 : 1);
    if (precision == 2 && dump) precision = 3;   // the trainer's forward (activations dumped for the backward pass) keeps full accuracy
    TrunkArgs a{(const u64 *)p1, (const u64 *)p2, color, n, out, out_kind, precision, precision == 2 ? s.d_blob2 : s.d_blob, s.d_bias, s.d_head, n_dev, s.d_cscale, {}, nullptr, {}};
    if (pairs) a.blob = precision == 2 ? s.d_pair2 : s.d_pair;
    for (int l = 0; l < 8; l++) a.dump[l] = dump ? dump[l] : nullptr;
    cudaStream_t cs = (cudaStream_t)stream;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(cs, &cap);
    const bool timed = cap == cudaStreamCaptureStatusNone;   // the timing events of iago_last_kernel_ms stay out of captured graphs (mcts.cu)
    if (timed) IAGO_CUDA(cudaEventRecord(ctx->ev0, cs));
    if (pairs) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = kSmemBytes;
        cfg.stream = cs;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        IAGO_CUDA(cudaLaunchKernelEx(&cfg, trunk_kernel<0, 2>, a, s.desc, precision == 2 ? s.maps2 : s.maps));
    } else {
        trunk_kernel<0, 1><<<grid, kThreads, kSmemBytes, cs>>>(a, s.desc, TrunkMaps{});
    }
    IAGO_CUDA(cudaGetLastError());
    if (timed) {
        IAGO_CUDA(cudaEventRecord(ctx->ev1, cs));
        ctx->timed = true;
    }
    return IAGO_OK;
}

// ---------------------------------------------------------------- refresh a policy slot from DEVICE parameters
// The same blob iago_load_net builds on the host (same roundings, bit-identical), written by kernels: the REINFORCE
// trainer refreshes its playing slot after every Adam step without a device -> host -> device round trip.
// One thread per weight element of a layer: the element's fp16 hi / lo parts and FP8 forms go to their places in the four operand
// layouts (single-CTA and pair, precision 3 and precision 2).  mode 0 = a trunk layer (chunk-major units of 64 input channels),
// 1 = layer 1 (explicit im2col, K = 18 padded to 32; fp16 hi / lo in every blob), 2 = the value head's block9 (N padded to 16).
__global__ void pack_all_kernel(const float *__restrict__ W, int cin, int n_pad, int mode, const float *__restrict__ sw_ptr,
                                uint8_t *__restrict__ blob, uint8_t *__restrict__ blob2, uint8_t *__restrict__ pair, uint8_t *__restrict__ pair2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int kgroups = mode == 1 ? 4 : 8, n_units = mode == 1 ? 1 : 9 * (cin / 64);
    if (idx >= n_units * kgroups * n_pad * 8) return;
    const int e = idx & 7, n = (idx >> 3) % n_pad, kg = ((idx >> 3) / n_pad) % kgroups, ut = (idx >> 3) / n_pad / kgroups;
    const int k = kg * 8 + e;
    float w;
    if (mode == 1) {
        w = k < 18 ? W[((size_t)n * 2 + (k & 1)) * 9 + (k >> 1)] : 0.0f;       // k = tap*2 + channel
    } else {
        const int ch = ut / 9, tap = ut % 9;                                   // chunk-major units
        w = mode == 2 ? (n == 0 ? W[(size_t)(ch * 64 + k) * 9 + tap] : 0.0f) : W[((size_t)n * cin + ch * 64 + k) * 9 + tap];
    }
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const size_t half_elems = (size_t)kgroups * n_pad * 8, unit_bytes = half_elems * 4;
    // single-CTA layouts
    {
        const size_t off = ((size_t)kg * n_pad + n) * 8 + e;
        __half *u = reinterpret_cast<__half *>(blob + (size_t)ut * unit_bytes);
        u[off] = h;
        u[half_elems + off] = l;
        __half *u2 = reinterpret_cast<__half *>(blob2 + (size_t)ut * unit_bytes);
        u2[off] = h;
        if (mode == 1) {
            u2[half_elems + off] = l;
        } else {
            const float sw = *sw_ptr;
            uint8_t *w8 = blob2 + (size_t)ut * unit_bytes + half_elems * 2, *wl8 = w8 + half_elems;
            const size_t o8 = ((size_t)(k >> 4) * n_pad + n) * 16 + (k & 15);
            w8[o8] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(h) * sw, __NV_SATFINITE, __NV_E4M3);
            wl8[o8] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(l) * 2048.0f * sw, __NV_SATFINITE, __NV_E4M3);
        }
    }
    // pair layouts: the unit as [rank 0's half of the output channels | rank 1's half], each a complete unit of n_pad / 2 columns
    {
        const int nh = n_pad / 2, hh = n / nh, nn = n % nh;
        const size_t he2 = half_elems / 2, base = (size_t)ut * unit_bytes + (size_t)hh * (unit_bytes / 2);
        const size_t off = ((size_t)kg * nh + nn) * 8 + e;
        __half *u = reinterpret_cast<__half *>(pair + base);
        u[off] = h;
        u[he2 + off] = l;
        __half *u2 = reinterpret_cast<__half *>(pair2 + base);
        u2[off] = h;
        if (mode == 1) {
            u2[he2 + off] = l;
        } else {
            const float sw = *sw_ptr;
            uint8_t *w8 = pair2 + base + he2 * 2, *wl8 = w8 + he2;
            const size_t o8 = ((size_t)(k >> 4) * nh + nn) * 16 + (k & 15);
            w8[o8] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(h) * sw, __NV_SATFINITE, __NV_E4M3);
            wl8[o8] = (uint8_t)__nv_cvt_float_to_fp8(__half2float(l) * 2048.0f * sw, __NV_SATFINITE, __NV_E4M3);
        }
    }
}

// Per-layer FP8 weight scale (the rule of fp8_weight_scale on the host): sw = the power of two that brings max |w| into [128, 256);
// cscale = 1 / (2^11 * sw).  One block per layer; layer 0 has no FP8 part.
__global__ void layer_scale_kernel(const float *__restrict__ params, int kind, float *__restrict__ sw_out, float *__restrict__ cscale_out) {
    const int cin[8] = {2, 64, 128, 128, 128, 128, 128, 128}, cout[8] = {64, 128, 128, 128, 128, 128, 128, 128};
    const int l = blockIdx.x;
    size_t off = 0, count = 0;
    for (int i = 0; i < 8; i++) {
        const size_t nw = (size_t)cout[i] * cin[i] * 9;
        if (i == l) count = nw;
        if (i < l || l == 8) off += nw + cout[i];
    }
    if (l == 8) count = kind == 1 ? 1152 : 0;
    __shared__ float red[256];
    float mx = 0.0f;
    for (size_t i = threadIdx.x; i < count; i += blockDim.x) mx = fmaxf(mx, fabsf(params[off + i]));
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
        if ((int)threadIdx.x < sft) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + sft]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        mx = red[0];
        const float sw = (!(mx > 0.0f) || !isfinite(mx)) ? 1.0f : ldexpf(1.0f, 7 - ilogbf(mx));
        sw_out[l] = sw;
        cscale_out[l] = (l == 0 || count == 0) ? 0.0f : 1.0f / (2048.0f * sw);
    }
}

__global__ void pack_bias_head_kernel(const float *__restrict__ params, float *__restrict__ bias, float *__restrict__ head, int kind) {
    // params: iago_load_net order.  bias[l][128] (zero padded); policy head = w9[128] | b10[64] (the value head is packed separately)
    const int cin[8] = {2, 64, 128, 128, 128, 128, 128, 128}, cout[8] = {64, 128, 128, 128, 128, 128, 128, 128};
    size_t off = 0;
    for (int l = 0; l < 8; l++) {
        off += (size_t)cout[l] * cin[l] * 9;
        for (int i = threadIdx.x; i < 128; i += blockDim.x) bias[l * 128 + i] = i < cout[l] ? params[off + i] : 0.0f;
        off += cout[l];
    }
    if (kind == 0)
        for (int i = threadIdx.x; i < 192; i += blockDim.x) head[i] = params[off + i];
}

// Value head scalars: b9 and the collapsed fc11 * fc10 64-vector (fp64 accumulation, as iago_load_net does on the host).
__global__ void pack_value_head_kernel(const float *__restrict__ hp /* W9[1152] | b9 | fc10[128][64] | fc11[128] */, float *__restrict__ head) {
    const int idx = threadIdx.x;
    if (idx < 64) {
        const float *fc10 = hp + 1153, *fc11 = fc10 + 128 * 64;
        double acc = 0.0;
        for (int i = 0; i < 128; i++) acc += (double)fc11[i] * (double)fc10[(size_t)i * 64 + idx];
        head[128 + idx] = (float)acc;
    }
    if (idx == 64) head[0] = hp[1152];
}

int trunk_refresh_slot(iago_ctx *ctx, int slot, int kind, const float *d_params, void *stream) {
    IAGO_REQUIRE(ctx && d_params, "NULL argument");
    IAGO_REQUIRE(slot >= 0 && slot < IAGO_NET_SLOTS, "slot out of range (0..IAGO_NET_SLOTS-1)");
    NetSlot &s = state(ctx)->slot[slot];
    if (!s.loaded || s.desc.kind != kind) {
        set_error("net slot %d holds no network of kind %d to refresh (call iago_load_net once first)", slot, kind);
        return IAGO_E_STATE;
    }
    const int cin[8] = {2, 64, 128, 128, 128, 128, 128, 128}, cout[8] = {64, 128, 128, 128, 128, 128, 128, 128};
    cudaStream_t cs = (cudaStream_t)stream;
    if (!st_scratch(ctx)) return IAGO_E_CUDA;
    float *d_sw = st_scratch(ctx);
    layer_scale_kernel<<<kMaxLayers, 256, 0, cs>>>(d_params, kind, d_sw, s.d_cscale);
    size_t off = 0;
    for (int l = 0; l < 8; l++) {
        const int total = (l == 0 ? 4 * cout[l] * 8 : 9 * (cin[l] / 64) * 8 * cout[l] * 8);
        const size_t ub = (size_t)s.desc.unit_base[l];
        pack_all_kernel<<<(total + 255) / 256, 256, 0, cs>>>(d_params + off, cin[l], cout[l], l == 0 ? 1 : 0, d_sw + l, s.d_blob + ub, s.d_blob2 + ub,
                                                             s.d_pair + ub, s.d_pair2 + ub);
        off += (size_t)cout[l] * cin[l] * 9 + cout[l];
    }
    pack_bias_head_kernel<<<1, 128, 0, cs>>>(d_params, s.d_bias, s.d_head, kind);
    if (kind == 1) {
        const size_t ub = (size_t)s.desc.unit_base[8];
        pack_all_kernel<<<(18 * 8 * 16 * 8 + 255) / 256, 256, 0, cs>>>(d_params + off, 128, 16, 2, d_sw + 8, s.d_blob + ub, s.d_blob2 + ub, s.d_pair + ub,
                                                                      s.d_pair2 + ub);
        pack_value_head_kernel<<<1, 128, 0, cs>>>(d_params + off, s.d_head);
    }
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

// ---------------------------------------------------------------- backward data-gradient chain (reinforce.cu)
// Chain layer i = dgrad of block 8-i (i = 0..6): dX[c] = sum_{o,tap} dY[o] (shifted by tap) * W[o][c][8-tap], K = 128 output
// channels of the block in two 64-channel chunks, N = its input channels (128, or 64 for block 2).
static size_t backward_layout_bytes() {
    size_t b = 0;
    for (int i = 0; i < 7; i++) b += (size_t)18 * 2 * 8 * (i == 6 ? 64 : 128) * 16;
    return b;
}
size_t trunk_backward_blob_bytes() { return 2 * backward_layout_bytes(); }   // the single-CTA layout, then the pair layout

// blob unit (chunk ch, tap): hi [8 k-groups][N][8] bf16 then lo; element (kg, n, e) = W_l[o = ch*64 + kg*8 + e][c = n][8 - tap].
__global__ void pack_dgrad_kernel(const float *__restrict__ W, uint8_t *__restrict__ units, uint8_t *__restrict__ pair_units, int cin) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over [ch 2][tap 9][kg 8][n cin][e 8]
    const int total = 2 * 9 * 8 * cin * 8;
    if (idx >= total) return;
    const int e = idx & 7, n = (idx >> 3) % cin, kg = ((idx >> 3) / cin) & 7, ut = (idx >> 3) / cin / 8;   // ut = ch*9 + tap
    const int ch = ut / 9, tap = ut % 9;
    const int o = ch * 64 + kg * 8 + e;
    const float w = W[((size_t)o * cin + n) * 9 + (8 - tap)];
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    const size_t half_elems = (size_t)8 * cin * 8;
    __nv_bfloat16 *unit = reinterpret_cast<__nv_bfloat16 *>(units) + (size_t)ut * 2 * half_elems;
    const size_t off = ((size_t)kg * cin + n) * 8 + e;
    unit[off] = h;
    unit[half_elems + off] = l;
    // pair layout (as pack_all_kernel's): [rank 0's half of the N columns | rank 1's half], each a complete unit of cin / 2 columns
    const int nh = cin / 2, hh = n / nh, nn = n % nh;
    __nv_bfloat16 *pu = reinterpret_cast<__nv_bfloat16 *>(pair_units) + (size_t)ut * 2 * half_elems + (size_t)hh * half_elems;
    const size_t poff = ((size_t)kg * nh + nn) * 8 + e;
    pu[poff] = h;
    pu[half_elems / 2 + poff] = l;
}

// The chain's layer table (a function of the architecture alone; passed to the kernel by value).
static NetDesc backward_desc() {
    NetDesc d;
    memset(&d, 0, sizeof d);
    d.n_layers = 7;
    d.kind = 2;
    size_t off = 0;
    for (int i = 0; i < 7; i++) {
        const int l = 7 - i, cin = l == 1 ? 64 : 128;
        LayerDesc &ld = d.layer[i];
        ld.n_units = 18; ld.ksteps = 4; ld.n = cin; ld.lo_off = 8 * cin * 16; ld.unit_bytes = 2 * ld.lo_off; ld.b_lbo = cin * 16; ld.chunks = 2;
        d.unit_base[i] = (long long)off;
        off += (size_t)18 * ld.unit_bytes;
    }
    return d;
}

int trunk_backward_pack(iago_ctx *ctx, const float *const *W /* W[l], l = 1..7: [128][cin_l][3][3] on the device */, uint8_t *blob,
                        void *stream) {
    const NetDesc d = backward_desc();
    cudaStream_t cs = (cudaStream_t)stream;
    for (int i = 0; i < 7; i++) {
        const int l = 7 - i, cin = l == 1 ? 64 : 128;
        const int total = 2 * 9 * 8 * cin * 8;
        pack_dgrad_kernel<<<(total + 255) / 256, 256, 0, cs>>>(W[l], blob + d.unit_base[i], blob + backward_layout_bytes() + d.unit_base[i], cin);
    }
    IAGO_CUDA(cudaGetLastError());
    (void)ctx;
    return IAGO_OK;
}

int trunk_backward_launch(iago_ctx *ctx, const uint8_t *blob, const float *dy_in, const float *const *mask,
                          float *const *dx_out, int64_t n, int precision, void *stream, unsigned *const *dymax) {
    IAGO_REQUIRE(ctx && blob && dy_in && mask && dx_out, "NULL argument");
    IAGO_REQUIRE(precision == 1 || precision == 3, "precision must be 1 (bf16) or 3 (bf16 hi/lo split)");
    if (n <= 0) return IAGO_OK;
    TrunkState *st = state(ctx);
    if (int rc = set_trunk_attrs(st)) return rc;
    // CTA pairs like the forward kernel (each CTA streams half of a weight unit from L2: a single-CTA chain, 148 SMs x 32 KB per unit,
    // sits on the chip's L2 bandwidth together with its mask reads and gradient writes); IAGO_TRUNK_CG1=1 forces single CTAs
    static const bool force_cg1 = getenv("IAGO_TRUNK_CG1") != nullptr || getenv("IAGO_TRUNK_BWD_CG1") != nullptr;
    const bool pairs = !force_cg1;
    const long long tiles = pairs ? (n + 3) / 4 : (n + 1) / 2;
    const long long max_groups = pairs ? ctx->sm_count / 2 : ctx->sm_count;
    const int grid = (int)(tiles < max_groups ? tiles : max_groups) * (pairs ? 2 : 1);
    TrunkArgs a{nullptr, nullptr, nullptr, n, nullptr, 0, precision, blob, nullptr, nullptr, nullptr, nullptr, {}, dy_in, {}};
    for (int i = 0; i < 7; i++) {
        a.dump[i] = dx_out[i];
        a.mask[i] = mask[i];
        a.dymax[i] = dymax ? dymax[i] : nullptr;
    }
    a.dump[7] = nullptr;
    a.mask[7] = nullptr;
    a.dymax[7] = nullptr;
    if (pairs) {
        const uint8_t *pair_blob = blob + backward_layout_bytes();
        if (st->bwd_pair != pair_blob) {
            if (int rc = make_maps(const_cast<uint8_t *>(pair_blob), backward_layout_bytes(), st->bwd_maps)) return rc;
            st->bwd_pair = pair_blob;
        }
        a.blob = pair_blob;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = kSmemBytes;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        IAGO_CUDA(cudaLaunchKernelEx(&cfg, trunk_kernel<1, 2>, a, backward_desc(), st->bwd_maps));
    } else {
        trunk_kernel<1, 1><<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(a, backward_desc(), TrunkMaps{});
    }
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}
}  // namespace iago

extern "C" {

int iago_policy_forward(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                        float *out, int out_kind, int precision, void *stream) {
    IAGO_REQUIRE(out_kind == 0 || out_kind == 1, "out_kind must be 0 (logits) or 1 (probabilities)");
    return trunk_launch(ctx, slot, 0, p1, p2, color, n, out, out_kind, precision, stream);
}

int iago_value_forward(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                       float *out, int precision, void *stream) {
    return trunk_launch(ctx, slot, 1, p1, p2, color, n, out, 0, precision, stream);
}

int iago_policy_forward_acts(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                             float *logits, float *const *acts, int precision, void *stream) {
    IAGO_REQUIRE(acts != nullptr, "acts is NULL");
    return trunk_launch(ctx, slot, 0, p1, p2, color, n, logits, 0, precision, stream, nullptr, acts);
}

int iago_value_forward_acts(iago_ctx *ctx, int slot, const uint64_t *p1, const uint64_t *p2, const uint8_t *color, int64_t n,
                            float *values, float *const *acts, int precision, void *stream) {
    IAGO_REQUIRE(acts != nullptr, "acts is NULL");
    return trunk_launch(ctx, slot, 1, p1, p2, color, n, values, 0, precision, stream, nullptr, acts);
}

}  // extern "C"
