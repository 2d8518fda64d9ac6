// reinforce.cu — K6: the REINFORCE / supervised update of src/train_rl.py:55-66 for the SL-size nets: heads, weight gradients (tcgen05
// GEMMs; an all-fp32 CUDA-core version of every kernel is kept as the checker, iago_reinforce_set_option(0)), bias gradients, Adam.
//
// Reference: x = stack([states==1, states==2]) (channel 0 = opponent, channel 1 = learner; the recorded states are
// colour-swapped, rl_self_play.py:134-138), pred = SLPolicy(x) — already softmax PROBABILITIES (network.py:47) —
// c = softmax_cross_entropy(pred, y, reduce='no') (a second log-softmax on the probabilities: reference quirk, kept),
// loss = mean(c * r), backward, Adam (Chainer defaults alpha 1e-3 / 5e-4, beta 0.9 / 0.999, eps 1e-8) with the
// WeightDecay(5e-4) hook (grad += 5e-4 * w) (src/train_rl.py:24-26).
//
// Layout: activations [M][C][64] fp32, kept for every layer (the backward needs them); parameters, gradients and Adam
// moments are flat fp32 arrays in the reference's npz key order (include/iago_b200.h iago_load_net).  The gradient
// returned is that of SUM_i c_i r_i (not the mean) plus the loss numerator and the position count, so that data-parallel
// ranks can all-reduce sums and divide once (DESIGN.md "multi-GPU").
//
// Kernels: implicit-GEMM 3x3 conv on a 2-board x BN-channel CTA tile with 8 x (BN/16) register tiles (forward with
// bias + ReLU; dgrad = the same kernel on transposed, tap-flipped weights with a ReLU-mask epilogue), a weight-gradient
// kernel (o x c x 9 register tiles, positions split over CTAs, deterministic two-stage reduction), the head (1x1 conv,
// per-cell bias, double softmax loss) and a fused WeightDecay + Adam step.  Every reduction runs in a fixed order: the
// gradient is bit-reproducible run to run.
#include <math.h>
#include <string.h>

#include <vector>

#include "bitboard.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc.cuh"

namespace iago {

constexpr int kNP = 960768;  // SLPolicy parameters
static const int kCin[8] = {2, 64, 128, 128, 128, 128, 128, 128};
static const int kCout[8] = {64, 128, 128, 128, 128, 128, 128, 128};

// ---------------------------------------------------------------- input planes
__global__ void planes_kernel(const u64 *__restrict__ own, const u64 *__restrict__ opp, float *__restrict__ x, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (position, cell)
    if (i >= m * 64) return;
    const long long p = i >> 6;
    const int cell = (int)(i & 63);
    x[(p * 2 + 0) * 64 + cell] = (float)((opp[p] >> cell) & 1);  // channel 0 = opponent stones (game.py:167-174)
    x[(p * 2 + 1) * 64 + cell] = (float)((own[p] >> cell) & 1);  // channel 1 = mover's (learner's) stones
}

// ---------------------------------------------------------------- weight re-layout
// forward : Wk[(c*9 + tap)][o]      = W[o][c][tap]
// dgrad   : Wk[(o*9 + tap)][c]      = W[o][c][8 - tap]      (inputs of the dgrad conv are the o channels)
__global__ void relayout_kernel(const float *__restrict__ W, float *__restrict__ Wf, float *__restrict__ Wd, int cin, int cout, int cin_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * cin * 9) return;
    const int tap = i % 9, c = (i / 9) % cin, o = i / (9 * cin);
    const float w = W[i];
    Wf[(size_t)(c * 9 + tap) * cout + o] = w;
    if (Wd) Wd[(size_t)(o * 9 + (8 - tap)) * cin + c] = w;
    (void)cin_pad;
}

// ---------------------------------------------------------------- 3x3 conv, implicit GEMM, fp32
// CTA tile: 2 boards (128 rows) x BN output channels; thread (ty, tx): ty = board row (b = ty / 8, r = ty % 8) -> the 8
// cells of that row, tx -> BN/16 consecutive output channels.
enum { EPI_FWD = 0, EPI_DGRAD = 1 };

template <int BN, int EPI>
__global__ void __launch_bounds__(256) conv3x3_kernel(const float *__restrict__ in, const float *__restrict__ Wk,
                                                      const float *__restrict__ bias, const float *__restrict__ mask_src,
                                                      float *__restrict__ out, long long m, int cin, int cout) {
    constexpr int TN = BN / 16;
    __shared__ float in_s[8][2][10][10];
    __shared__ __align__(16) float w_s[72][BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b = ty >> 3, r = ty & 7;
    const long long pos0 = (long long)blockIdx.x * 2;
    const int n0 = blockIdx.y * BN + tx * TN;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;
    for (int i = tid; i < 8 * 2 * 100; i += 256) (&in_s[0][0][0][0])[i] = 0.0f;  // halo stays zero
    const int chunks = (cin + 7) / 8;
    for (int ch0 = 0; ch0 < chunks * 8; ch0 += 8) {
        __syncthreads();
        // input chunk: 8 channels x 2 boards x 64 cells
        for (int i = tid; i < 8 * 2 * 64; i += 256) {
            const int cell = i & 63, bb = (i >> 6) & 1, c = i >> 7;
            const long long p = pos0 + bb;
            float v = 0.0f;
            if (p < m && ch0 + c < cin) v = in[(p * cin + ch0 + c) * 64 + cell];
            in_s[c][bb][(cell >> 3) + 1][(cell & 7) + 1] = v;
        }
        // weight chunk: rows (ch0*9 .. ch0*9+71) x BN columns of this CTA
        for (int i = tid; i < 72 * (BN / 4); i += 256) {
            const int row = i / (BN / 4), c4 = i % (BN / 4);
            const int krow = ch0 * 9 + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (krow < cin * 9) v = *reinterpret_cast<const float4 *>(Wk + (size_t)krow * cout + blockIdx.y * BN + c4 * 4);
            *reinterpret_cast<float4 *>(&w_s[row][c4 * 4]) = v;
        }
        __syncthreads();
#pragma unroll 2
        for (int c = 0; c < 8; c++) {
#pragma unroll
            for (int dy = 0; dy < 3; dy++) {
                float a[10];
#pragma unroll
                for (int i = 0; i < 10; i++) a[i] = in_s[c][b][r + dy][i];
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    float w[TN];
#pragma unroll
                    for (int j = 0; j < TN; j += 4) {
                        const float4 v = *reinterpret_cast<const float4 *>(&w_s[c * 9 + dy * 3 + dx][tx * TN + j]);
                        w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i + dx], w[j], acc[i][j]);
                }
            }
        }
    }
    const long long p = pos0 + b;
    if (p >= m) return;
#pragma unroll
    for (int j = 0; j < TN; j++) {
        const int n = n0 + j;
        if (n >= cout) continue;
        float v[8];
        const size_t off = ((size_t)p * cout + n) * 64 + r * 8;
        if (EPI == EPI_FWD) {
            const float bj = bias[n];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = fmaxf(acc[i][j] + bj, 0.0f);
        } else {
            const float4 m0 = *reinterpret_cast<const float4 *>(mask_src + off), m1 = *reinterpret_cast<const float4 *>(mask_src + off + 4);
            const float mk[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = mk[i] > 0.0f ? acc[i][j] : 0.0f;
        }
        *reinterpret_cast<float4 *>(out + off) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(out + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

// ---------------------------------------------------------------- weight gradient
// dW[o][c][tap] = sum over positions and cells of dY[p][o][cell] * X[p][c][cell + tap]; db[o] = sum dY[p][o][cell].
// grid = (c tiles of 16, position slices); thread (og, ci): 8 output channels x 1 input channel x 9 taps.
// partial[slice] holds [cout][cin][9] then [cout] bias sums (written by the c tile 0 CTAs).
__global__ void __launch_bounds__(256) wgrad3x3_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                       float *__restrict__ partial, long long m, int cin, int cout,
                                                       int pos_per_slice, size_t partial_stride) {
    __shared__ float dy_s[128][65];
    __shared__ float x_s[16][101];
    const int tid = threadIdx.x, og = tid >> 4, ci = tid & 15;
    const int c = blockIdx.x * 16 + ci;
    const long long p_begin = (long long)blockIdx.y * pos_per_slice;
    const long long p_end = min(m, p_begin + pos_per_slice);
    float acc[8][9], bsum[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        bsum[j] = 0.0f;
#pragma unroll
        for (int t = 0; t < 9; t++) acc[j][t] = 0.0f;
    }
    for (int i = tid; i < 16 * 101; i += 256) (&x_s[0][0])[i] = 0.0f;
    const int o_tiles = (cout + 127) / 128;  // cout <= 128 here
    (void)o_tiles;
    for (long long p = p_begin; p < p_end; p++) {
        __syncthreads();
        for (int i = tid; i < cout * 64; i += 256) dy_s[i >> 6][i & 63] = dy[(size_t)p * cout * 64 + i];
        for (int i = tid; i < 16 * 64; i += 256) {
            const int cc = i >> 6, cell = i & 63, cg = blockIdx.x * 16 + cc;
            x_s[cc][((cell >> 3) + 1) * 10 + (cell & 7) + 1] = cg < cin ? x[((size_t)p * cin + cg) * 64 + cell] : 0.0f;
        }
        __syncthreads();
        if (og * 8 < cout) {
#pragma unroll 1
            for (int row = 0; row < 8; row++) {
                float xv[3][10];
#pragma unroll
                for (int dyy = 0; dyy < 3; dyy++)
#pragma unroll
                    for (int i = 0; i < 10; i++) xv[dyy][i] = x_s[ci][(row + dyy) * 10 + i];
#pragma unroll
                for (int col = 0; col < 8; col++) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float d = dy_s[og * 8 + j][row * 8 + col];
                        bsum[j] += d;
#pragma unroll
                        for (int t = 0; t < 9; t++) acc[j][t] = fmaf(d, xv[t / 3][col + t % 3], acc[j][t]);
                    }
                }
            }
        }
    }
    float *dst = partial + (size_t)blockIdx.y * partial_stride;
    if (og * 8 < cout && c < cin) {
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int t = 0; t < 9; t++) dst[((size_t)(og * 8 + j) * cin + c) * 9 + t] = acc[j][t];
    }
    if (blockIdx.x == 0 && ci == 0 && og * 8 < cout) {
#pragma unroll
        for (int j = 0; j < 8; j++) dst[(size_t)cout * cin * 9 + og * 8 + j] = bsum[j];
    }
}

// Block 1 (2 input planes -> 64 channels): the generic kernel above keeps 16 input channels x 16 groups of 8 outputs busy, of which
// this layer uses 2 x 8 (6 % of its FMAs).  Here a thread owns one output channel and a quarter of the board rows for both planes and
// all nine taps (18 accumulators + the bias sum), the quarters are added in a fixed order at the end.  Same output layout:
// partial[slice][(o * 2 + c) * 9 + tap], then the 64 bias sums.
__global__ void __launch_bounds__(256) wgrad_block1_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                           float *__restrict__ partial, long long m, int pos_per_slice,
                                                           size_t partial_stride) {
    __shared__ float dy_s[64][65];
    __shared__ float x_s[2][104];
    __shared__ float red[4][64][19];
    const int tid = threadIdx.x, o = tid & 63, q = tid >> 6;
    const long long p_begin = (long long)blockIdx.x * pos_per_slice;
    const long long p_end = min(m, p_begin + pos_per_slice);
    float acc[2][9], bsum = 0.0f;
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int t = 0; t < 9; t++) acc[c][t] = 0.0f;
    for (int i = tid; i < 2 * 104; i += 256) (&x_s[0][0])[i] = 0.0f;   // the halo stays zero
    for (long long p = p_begin; p < p_end; p++) {
        __syncthreads();
        for (int i = tid; i < 64 * 64; i += 256) dy_s[i >> 6][i & 63] = dy[(size_t)p * 64 * 64 + i];
        if (tid < 128) {
            const int c = tid >> 6, cell = tid & 63;
            x_s[c][((cell >> 3) + 1) * 10 + (cell & 7) + 1] = x[((size_t)p * 2 + c) * 64 + cell];
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const int row = q * 2 + r;
#pragma unroll
            for (int col = 0; col < 8; col++) {
                const float d = dy_s[o][row * 8 + col];
                bsum += d;
#pragma unroll
                for (int c = 0; c < 2; c++)
#pragma unroll
                    for (int t = 0; t < 9; t++) acc[c][t] = fmaf(d, x_s[c][(row + t / 3) * 10 + col + t % 3], acc[c][t]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 2; c++)
#pragma unroll
        for (int t = 0; t < 9; t++) red[q][o][c * 9 + t] = acc[c][t];
    red[q][o][18] = bsum;
    __syncthreads();
    float *dst = partial + (size_t)blockIdx.x * partial_stride;
    for (int i = tid; i < 64 * 19; i += 256) {
        const int oo = i / 19, k = i - oo * 19;
        const float v = (red[0][oo][k] + red[1][oo][k]) + (red[2][oo][k] + red[3][oo][k]);
        if (k < 18) dst[(size_t)oo * 18 + k] = v;        // (o * 2 + c) * 9 + tap = o * 18 + c * 9 + tap
        else dst[(size_t)64 * 2 * 9 + oo] = v;
    }
}

// ---------------------------------------------------------------- weight gradient on the tensor cores (layers with Cout = 128)
// dW[o][c][tap] = sum_{p, cell} dY[p][o][cell] * X[p][c][cell + tap] as 9 GEMMs D_tap[o][c] += A[o][k] * B_tap[c][k], k = (p, cell).
// One CTA owns one kernel row ky (3 taps, 3 x N fp32 TMEM columns) and one slice of positions; per position (one pipeline
// stage, K = 64) the producer warps read dY and X (fp32, [p][ch][64]) once, convert to fp16 (dY scaled by a power of two taken from the
// layer's max |dY| and split into hi + lo parts; X as is: activations are bounded) — bf16 operands were measured at 1e-2 of max|g| — and lay them out as no-swizzle
// K-major core matrices: A = [o group][board row][o % 8][8 cells]; B = three column-shifted copies (kx = 0, 1, 2 <-> dx = -1, 0, +1,
// zero filled) of [c group][padded row 0..9][c % 8][8 cells], so that a tap is just a descriptor start address: copy kx, padded
// row y + ky.  One thread issues 4 K=16 MMAs per tap and stage; accumulators stay in TMEM until the slice is done.
// Output: partial[slice][tap][o][c] (coalesced); reduce_taps_kernel sums the slices in order and writes [o][c][tap].
#ifndef IAGO_WG_PRODUCERS
#define IAGO_WG_PRODUCERS 16
#endif
constexpr int kWgProducers = IAGO_WG_PRODUCERS;          // producer warps (8 or 16; warps 0-3 also run the epilogue), then one MMA issuer warp
constexpr int kWgItems = 32 / kWgProducers;               // (channel group, row half) items of the dY tile per producer thread and position
constexpr int kWgThreads = kWgProducers * 32 + 32;
// Core matrices (8 channels x 16 B) sit 144 B apart along K (board rows), not 128: a quarter warp that stores the 8 rows of ONE channel
// then hits 8 different 16-byte bank groups, so the producers can read global memory row-fastest — a warp's LDG.256 covers 1 KB of
// consecutive addresses.  (With 128 B the conflict-free store mapping was channel-fastest, whose loads touch 8 lines in 32 separate
// sectors: ncu counted 39 L1 wavefronts per load instruction and the LSU data pipe at 87 % of peak, the kernel's limiter.)
constexpr int kWgLbo = 144;                       // bytes between K-adjacent core matrices (descriptor LBO)
constexpr int kWgSboA = 8 * kWgLbo;               // dY: bytes between o groups (8 board rows each)
constexpr int kWgSboB = 10 * kWgLbo;              // X: bytes between c groups (10 padded rows each)
constexpr int kWgATile = 16 * kWgSboA;            // dY tile: 16 o-groups = 18,432
constexpr int kWgXCopy = 16 * kWgSboB;            // one shifted copy of the X tile: 23,040
constexpr int kWgXBase = 2 * kWgATile;            // dY hi tile, dY lo tile, then the three X copies
constexpr int kWgStage = kWgXBase + 3 * kWgXCopy;   // 105,984
constexpr int kWgStages = 2;
constexpr int kWgSmem = kWgStages * kWgStage + 64;

// One 32-byte sector per lane in ONE request (LDG.256, sm_100): with two 16-byte loads every sector was requested twice and the
// second half had to survive in an L1 that the kernel's 207 KB of shared memory leaves almost no room for.
__device__ __forceinline__ void ldg256(const float *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"   // read once: no L1 line (6.12 against 6.18 ms per gradient)
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// 8 floats, scaled by a power of two into fp16's range -> fp16 hi parts and fp16 lo parts (x - hi)
__device__ __forceinline__ void pack_f16x8_split(const float4 a, const float4 b, float scale, uint4 &hi, uint4 &lo) {
    const float f[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const __half2 ph = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        const float2 back = __half22float2(ph);
        const __half2 pl = __floats2half2_rn(f[2 * i] - back.x, f[2 * i + 1] - back.y);
        h[i] = *reinterpret_cast<const uint32_t *>(&ph);
        l[i] = *reinterpret_cast<const uint32_t *>(&pl);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// 8 floats -> fp16 (activations: bounded range, 11-bit significand)
__device__ __forceinline__ uint4 pack_f16x8(const float4 a, const float4 b) {
    const __half2 p0 = __floats2half2_rn(a.x, a.y), p1 = __floats2half2_rn(a.z, a.w);
    const __half2 p2 = __floats2half2_rn(b.x, b.y), p3 = __floats2half2_rn(b.z, b.w);
    return make_uint4(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1),
                      *reinterpret_cast<const uint32_t *>(&p2), *reinterpret_cast<const uint32_t *>(&p3));
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                                  float *__restrict__ partial, long long m, int cin,
                                                                  int pos_per_slice, size_t partial_stride,
                                                                  const unsigned *__restrict__ dymax, float *__restrict__ bias_partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int ky = blockIdx.y;                       // kernel row of this CTA: taps ky*3 + {0,1,2}
    const int N = cin;                               // 64 or 128 input channels = N of the MMA
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + kWgStages * kWgStage, bar_empty = bar_full + 8 * kWgStages, bar_acc = bar_empty + 8 * kWgStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kWgStages * kWgStage + 8 * (2 * kWgStages + 1));
    const long long p_begin = (long long)blockIdx.x * pos_per_slice;
    const long long p_end = min(m, p_begin + pos_per_slice);
    const int n_pos = (int)max(0LL, p_end - p_begin);
    // dY is scaled so that its largest magnitude lands in [1024, 2048): far from fp16 overflow, and everything within 2^-24 of the
    // maximum keeps at least subnormal precision in the hi part (the lo part extends that by 11 bits)
    float scale = 1.0f;
    {
        const float mx = __uint_as_float(*dymax);
        if (mx > 0.0f) scale = exp2f((float)(10 - ilogbf(mx)));
    }
    const float inv_scale = 1.0f / scale;

    for (int i = tid; i < kWgStages * kWgStage / 16; i += kWgThreads) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);  // halo rows stay zero
    if (tid == 0) {
        for (int s = 0; s < kWgStages; s++) {
            mbar_init(bar_full + 8 * s, kWgProducers * 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWgProducers) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < kWgProducers) {
        // ================= producers: fp32 global -> 16-bit core matrices in shared memory =================
        // Software-pipelined through registers: the 16 x 16-byte loads of position p+1 are in flight while position p is
        // converted and stored, so HBM/L2 latency is covered without a third shared-memory stage.
        // Work item = one (channel, board row) chunk of 8 cells = 32 B of fp32 in, one 16-byte core-matrix row out.  Lane bits: board row in
        // bits 0-2, channel % 4 in bits 3-4: a warp reads 1 KB of consecutive bytes per LDG.256, and the 8 lanes of a quarter warp store
        // the 8 rows of one channel, kWgLbo = 144 B apart (8 different bank groups).
        uint32_t stage = 0, phase = 0;
        const int xchunks = cin / (4 * kWgProducers);   // (cin * 8 rows) / producer threads: 4 or 2 with 8 producer warps, 2 or 1 with 16
        const int lane = tid & 31, row_of_lane = lane & 7, ch_of_lane = lane >> 3;
        int a_ch[kWgItems], a_row[kWgItems], x_ch[kWgItems], x_row[kWgItems];
#pragma unroll
        for (int it = 0; it < kWgItems; it++) {
            a_ch[it] = ((tid >> 5) * kWgItems + it) * 4 + ch_of_lane;   // a warp's load = 4 whole channels = 1 KB of consecutive bytes
            a_row[it] = row_of_lane;
            x_ch[it] = ((tid >> 5) * xchunks + it) * 4 + ch_of_lane;
            x_row[it] = row_of_lane;
        }
        float4 cur[4 * kWgItems], nxt[4 * kWgItems];
        auto load = [&](float4 (&r)[4 * kWgItems], long long p) {
            const float *dsrc = dy + (size_t)p * 128 * 64;
            const float *xsrc = x + (size_t)p * cin * 64;
#pragma unroll
            for (int it = 0; it < kWgItems; it++) {
                ldg256(dsrc + a_ch[it] * 64 + a_row[it] * 8, r[2 * it], r[2 * it + 1]);
            }
#pragma unroll
            for (int it = 0; it < kWgItems; it++) {
                if (it < xchunks) {
                    ldg256(xsrc + x_ch[it] * 64 + x_row[it] * 8, r[2 * kWgItems + 2 * it], r[2 * kWgItems + 2 * it + 1]);
                }
            }
        };
        // Bias gradient = sum of dY over positions and cells, taken by the ky = 0 CTA of each slice from the registers it converts anyway
        // (a separate kernel used to read every dY tensor once more for it: 0.39 ms per 8,192 positions).  Both items of a thread are the
        // two row halves of ONE channel; the four row_low lanes are folded by shuffles after the loop.  Fixed order throughout.
        float bsum[kWgItems];
#pragma unroll
        for (int it = 0; it < kWgItems; it++) bsum[it] = 0.0f;
        if (n_pos > 0) load(cur, p_begin);
        for (int ip = 0; ip < n_pos; ip++) {
            if (ip + 1 < n_pos) load(nxt, p_begin + ip + 1);
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            uint8_t *st = smem + stage * kWgStage;
            // dY: 128 o x 8 rows of 8 cells
#pragma unroll
            for (int it = 0; it < kWgItems; it++) {
                const int o = a_ch[it], row = a_row[it];
                uint4 hi, lo;
                pack_f16x8_split(cur[2 * it], cur[2 * it + 1], scale, hi, lo);
                if (ky == 0) {
                    const float4 u = cur[2 * it], v = cur[2 * it + 1];
                    bsum[it] += ((u.x + u.y) + (u.z + u.w)) + ((v.x + v.y) + (v.z + v.w));
                }
                const uint32_t off = (o >> 3) * kWgSboA + row * kWgLbo + (o & 7) * 16;
                *reinterpret_cast<uint4 *>(st + off) = hi;
                *reinterpret_cast<uint4 *>(st + kWgATile + off) = lo;
            }
            // X: cin channels x 8 rows, three column-shifted copies
#pragma unroll
            for (int it = 0; it < kWgItems; it++) {
                if (it < xchunks) {
                    const int c = x_ch[it], row = x_row[it];
                    const uint4 v = pack_f16x8(cur[2 * kWgItems + 2 * it], cur[2 * kWgItems + 2 * it + 1]);
                    const uint32_t off = kWgXBase + (c >> 3) * kWgSboB + (row + 1) * kWgLbo + (c & 7) * 16;
                    // kx = 0: out[x] = in[x - 1]; kx = 1: in[x]; kx = 2: out[x] = in[x + 1]   (zero beyond the board edge)
                    const uint4 left = make_uint4(v.x << 16, __funnelshift_l(v.x, v.y, 16), __funnelshift_l(v.y, v.z, 16), __funnelshift_l(v.z, v.w, 16));
                    const uint4 right = make_uint4(__funnelshift_r(v.x, v.y, 16), __funnelshift_r(v.y, v.z, 16), __funnelshift_r(v.z, v.w, 16), v.w >> 16);
                    *reinterpret_cast<uint4 *>(st + off) = left;
                    *reinterpret_cast<uint4 *>(st + off + kWgXCopy) = v;
                    *reinterpret_cast<uint4 *>(st + off + 2 * kWgXCopy) = right;
                }
            }
            fence_async_smem();
            mbar_arrive(bar_full + 8 * stage);
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
#pragma unroll
            for (int i = 0; i < 4 * kWgItems; i++) cur[i] = nxt[i];
        }
        if (ky == 0) {
#pragma unroll
            for (int it = 0; it < kWgItems; it++) {
                float b = bsum[it];
                b += __shfl_xor_sync(0xFFFFFFFFu, b, 1);
                b += __shfl_xor_sync(0xFFFFFFFFu, b, 2);
                b += __shfl_xor_sync(0xFFFFFFFFu, b, 4);
                if (row_of_lane == 0) bias_partial[(size_t)blockIdx.x * 128 + a_ch[it]] = b;
            }
        }
    } else if ((tid & 31) == 0) {
        // ================= MMA issuer =================
        const uint32_t idesc = instr_desc(128, N), idesc256 = instr_desc(128, 256);   // fp16 operands, fp32 accumulate
        const uint64_t hi_a = ((uint64_t)(kWgSboA >> 4) << 32) | (1ULL << 46);   // SBO: bytes between o groups
        const uint64_t hi_b = ((uint64_t)(kWgSboB >> 4) << 32) | (1ULL << 46);   // SBO: bytes between c groups (10 padded rows)
        const uint32_t lbo_word = (uint32_t)(kWgLbo >> 4) << 16;                 // LBO: bytes between K-adjacent core matrices (board rows)
        uint32_t stage = 0, phase = 0;
        for (int ip = 0; ip < n_pos; ip++) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t st = sbase + stage * kWgStage;
            if (N == 128) {
                // The X copies for kx = 0 and kx = 1 are adjacent in units of the c-group stride (16 groups x kWgSboB = one copy), so
                // one N = 256 MMA serves both taps and reads the dY tile once for the two of them; accumulator columns 0-255 are
                // exactly taps kx = 0 | kx = 1.  (Shared-memory throughput is what bounds this kernel: 32 KB less per position.)
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {   // board rows (2 ks, 2 ks + 1) of dY against padded rows (2 ks + ky, 2 ks + ky + 1) of X
                    const uint32_t aw = ((st + ks * 2 * kWgLbo) >> 4) | lbo_word, alw = ((st + kWgATile + ks * 2 * kWgLbo) >> 4) | lbo_word;
                    const uint32_t bw01 = ((st + kWgXBase + (2 * ks + ky) * kWgLbo) >> 4) | lbo_word;
                    const uint32_t bw2 = ((st + kWgXBase + 2 * kWgXCopy + (2 * ks + ky) * kWgLbo) >> 4) | lbo_word;
                    const uint32_t acc = (ip > 0 || ks > 0) ? 1u : 0u;
                    umma_f16(tmem, hi_a | aw, hi_b | bw01, idesc256, acc);
                    umma_f16(tmem, hi_a | alw, hi_b | bw01, idesc256, 1u);
                    umma_f16(tmem + 256, hi_a | aw, hi_b | bw2, idesc, acc);
                    umma_f16(tmem + 256, hi_a | alw, hi_b | bw2, idesc, 1u);
                }
            } else {
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        const uint32_t aw = ((st + ks * 2 * kWgLbo) >> 4) | lbo_word, alw = ((st + kWgATile + ks * 2 * kWgLbo) >> 4) | lbo_word;
                        const uint32_t bw = ((st + kWgXBase + kx * kWgXCopy + (2 * ks + ky) * kWgLbo) >> 4) | lbo_word;
                        umma_f16(tmem + kx * 128, hi_a | aw, hi_b | bw, idesc, (ip > 0 || ks > 0) ? 1u : 0u);
                        umma_f16(tmem + kx * 128, hi_a | alw, hi_b | bw, idesc, 1u);
                    }
                }
            }
            umma_commit(bar_empty + 8 * stage);
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_acc);
    }

    // ================= epilogue: TMEM -> partial[slice][tap][o][c] =================
    if (warp < 4) {
        float *dst = partial + (size_t)blockIdx.x * partial_stride;
        const int o = tid;   // TMEM lane = output channel
        if (n_pos > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
        }
        for (int kx = 0; kx < 3; kx++) {
            const int tap = ky * 3 + kx;
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t v[32];
                if (n_pos > 0) {
                    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + kx * 128 + c0, v);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = 0;
                }
                float *row = dst + ((size_t)tap * 128 + o) * N + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(row + j) = make_float4(__uint_as_float(v[j]) * inv_scale, __uint_as_float(v[j + 1]) * inv_scale,
                                                                       __uint_as_float(v[j + 2]) * inv_scale, __uint_as_float(v[j + 3]) * inv_scale);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWgProducers) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// grad[(o*cin + c)*9 + tap] (+)= sum over slices (in order) of partial[s][tap][o][c]
__global__ void reduce_taps_kernel(const float *__restrict__ partial, float *__restrict__ out, int cin, int slices, size_t stride, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // index into [tap][o][c]
    if (i >= 9 * 128 * cin) return;
    const int c = i % cin, o = (i / cin) % 128, tap = i / (cin * 128);
    float s = 0.0f;
    for (int k = 0; k < slices; k++) s += partial[(size_t)k * stride + i];
    float *dst = out + ((size_t)o * cin + c) * 9 + tap;
    *dst = accumulate ? *dst + s : s;
}

// out[i] (+)= sum over slices of partial[s][i], slices in order.
__global__ void reduce_slices_kernel(const float *__restrict__ partial, float *__restrict__ out, int count, int slices,
                                     size_t stride, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.0f;
    for (int k = 0; k < slices; k++) s += partial[(size_t)k * stride + i];
    out[i] = accumulate ? out[i] + s : s;
}

// ---------------------------------------------------------------- head: conv9 (1x1) + bias10 + softmax + CE(softmax) * r
// one CTA of 64 threads (= cells) per position.
__device__ __forceinline__ float block64_max(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const float r = fmaxf(red[0], red[1]);
    __syncthreads();
    return r;
}
__device__ __forceinline__ float block64_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const float r = red[0] + red[1];
    __syncthreads();
    return r;
}

// max |dY| of a layer as a bit pattern, collected where dY is written (the weight-gradient kernel scales its fp16 operands by it):
// one atomicMax per warp; NaN / inf stay out, the scale then defaults to 1
__device__ __forceinline__ void publish_absmax(float mx, unsigned *dymax) {
    const unsigned w = __reduce_max_sync(0xFFFFFFFFu, mx < 3.0e38f ? __float_as_uint(mx) : 0u);
    if ((threadIdx.x & 31) == 0 && w != 0u) atomicMax(dymax, w);
}

__global__ void __launch_bounds__(64) head_kernel(const float *__restrict__ act8, const float *__restrict__ w9, const float *__restrict__ b10,
                                                  const int8_t *__restrict__ action, const float *__restrict__ reward,
                                                  float *__restrict__ dlogit, float *__restrict__ dact8, float *__restrict__ loss_terms,
                                                  float *__restrict__ probs_out, long long m, unsigned *__restrict__ dymax) {
    __shared__ float red[2];
    const long long p = blockIdx.x;
    const int cell = threadIdx.x;
    const float *a = act8 + (size_t)p * 128 * 64;
    float logit = 0.0f;
    for (int c = 0; c < 128; c++) logit = fmaf(w9[c], a[c * 64 + cell], logit);
    logit += b10[cell];
    // pred = softmax(logits)                                             network.py:47
    const float mx = block64_max(logit, red);
    const float e = expf(logit - mx);
    const float pred = e / block64_sum(e, red);
    if (probs_out) probs_out[p * 64 + cell] = pred;
    // c = -log_softmax(pred)[y]; loss term = c * r                       src/train_rl.py:62-64
    const float mx2 = block64_max(pred, red);
    const float e2 = expf(pred - mx2);
    const float s2 = block64_sum(e2, red);
    const float q = e2 / s2;                      // softmax(pred)
    const int y = action[p];
    const float r = reward[p];
    if (cell == y) loss_terms[p] = -(pred - mx2 - logf(s2)) * r;
    const float dpred = r * (q - (cell == y ? 1.0f : 0.0f));
    // back through pred = softmax(logits): dlogit = pred * (dpred - sum_k dpred_k pred_k)
    const float dot = block64_sum(dpred * pred, red);
    const float dl = pred * (dpred - dot);
    dlogit[p * 64 + cell] = dl;
    float *da = dact8 + (size_t)p * 128 * 64;
    float amax = 0.0f;
    for (int c = 0; c < 128; c++) {
        const float g = a[c * 64 + cell] > 0.0f ? w9[c] * dl : 0.0f;
        da[c * 64 + cell] = g;
        amax = fmaxf(amax, fabsf(g));
    }
    publish_absmax(amax, dymax);
}

// dw9[c] = sum_{p,cell} dlogit[p][cell] * act8[p][c][cell]  (block c < 128); db10[cell] = sum_p dlogit[p][cell] (block 128 + cell);
// block 192: loss numerator = sum_p loss_terms[p].  Fixed-order tree per block.
// grid (193, kHeadSlices): block (b, s) sums its quantity over the s-th slice of the positions into partial[s * 193 + b] (one CTA per
// quantity streamed the whole batch through one dependent accumulation chain: 0.31 ms per 8,192 positions); head_grad_reduce_kernel
// adds the slices in order.
constexpr int kHeadSlices = 16;
__global__ void __launch_bounds__(256) head_grad_kernel(const float *__restrict__ act8, const float *__restrict__ dlogit,
                                                        const float *__restrict__ loss_terms, float *__restrict__ partial, long long m) {
    __shared__ float red[256];
    const int blk = blockIdx.x, tid = threadIdx.x;
    const long long per = (m + kHeadSlices - 1) / kHeadSlices;
    const long long p0 = (long long)blockIdx.y * per, p1 = min(m, p0 + per);
    float s = 0.0f;
    if (blk < 128) {
        for (long long i = p0 * 64 + tid; i < p1 * 64; i += 256) s += dlogit[i] * act8[((i >> 6) * 128 + blk) * 64 + (i & 63)];
    } else if (blk < 192) {
        for (long long p = p0 + tid; p < p1; p += 256) s += dlogit[p * 64 + (blk - 128)];
    } else {
        for (long long p = p0 + tid; p < p1; p += 256) s += loss_terms[p];
    }
    red[tid] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) partial[blockIdx.y * 193 + blk] = red[0];
}
__global__ void head_grad_reduce_kernel(const float *__restrict__ partial, float *__restrict__ gw9, float *__restrict__ gb10,
                                        float *__restrict__ gloss, int accumulate) {
    const int blk = threadIdx.x;
    if (blk >= 193) return;
    float v = 0.0f;
    for (int s = 0; s < kHeadSlices; s++) v += partial[s * 193 + blk];
    float *dst = blk < 128 ? gw9 + blk : (blk < 192 ? gb10 + (blk - 128) : gloss);
    *dst = accumulate ? *dst + v : v;
}

// ---------------------------------------------------------------- Value head (network.py:92-95) forward + backward, train_value.py:50-56
// One CTA of 128 threads per position.  pre9 = block9 (3x3, 128 -> 1) + b9, h9 = relu(pre9), u = fc10 h9, z = dropout(u, ratio)
// (Chainer: mask = rand >= ratio, scale 1/(1 - ratio)), v = fc11 z; loss term = (v - y)^2, dv = 2 (v - y)  (mean_squared_error is the
// mean over the minibatch; the caller divides the summed gradient by the position count once, like the policy trainer).
// Kept per position for the weight-gradient kernels: h9[64], du[128] (d loss / d u), z[128], dpre9[64], dv.
// dact8 = gradient w.r.t. block 8's pre-activation (the ReLU mask of act8 applied), the entry of the trunk's backward chain.
struct ValueHeadArgs {
    const float *act8, *w9, *b9, *fc10, *fc11, *target;
    float *pred, *loss_terms, *h9, *du, *z, *dpre9, *dv, *dact8;
    uint8_t *mask_out;   // nullable [m][128]: 1 = kept
    unsigned *dymax;     // bit pattern of max |dact8| (atomicMax)
    long long m;
    float ratio;
    u64 seed, pos_id0;
};

__device__ __forceinline__ float block128_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const float r = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(128) value_head_kernel(ValueHeadArgs a) {
    __shared__ float a8[128][65];
    __shared__ float w9s[1152];
    __shared__ float h9s[64], dps[64], dus[128];
    __shared__ float red[4];
    const long long p = blockIdx.x;
    const int tid = threadIdx.x;
    const float *src = a.act8 + (size_t)p * 128 * 64;
    for (int i = tid; i < 128 * 64; i += 128) a8[i >> 6][i & 63] = src[i];
    for (int i = tid; i < 1152; i += 128) w9s[i] = a.w9[i];
    __syncthreads();
    if (tid < 64) {
        const int y = tid >> 3, x = tid & 7;
        float acc = 0.0f;
        for (int c = 0; c < 128; c++)
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) acc = fmaf(w9s[c * 9 + t], a8[c][yy * 8 + xx], acc);
            }
        acc += a.b9[0];
        h9s[tid] = fmaxf(acc, 0.0f);
        a.h9[p * 64 + tid] = h9s[tid];
    }
    __syncthreads();
    float u = 0.0f;
    {
        const float *row = a.fc10 + (size_t)tid * 64;
        for (int j = 0; j < 64; j++) u = fmaf(row[j], h9s[j], u);
    }
    float keep = 1.0f, scale = 1.0f;
    if (a.ratio > 0.0f) {
        const double r = (double)philox_m53(a.seed, a.pos_id0 + (u64)p, (uint32_t)tid, 5u) * (1.0 / 9007199254740992.0);
        keep = r >= (double)a.ratio ? 1.0f : 0.0f;      // chainer.functions.dropout: mask = rand >= ratio
        scale = 1.0f / (1.0f - a.ratio);
    }
    if (a.mask_out) a.mask_out[p * 128 + tid] = keep != 0.0f;
    const float z = u * keep * scale;
    const float w11 = a.fc11[tid];
    const float v = block128_sum(w11 * z, red);
    const float dv = 2.0f * (v - a.target[p]);
    if (tid == 0) {
        a.pred[p] = v;
        a.loss_terms[p] = (v - a.target[p]) * (v - a.target[p]);
        a.dv[p] = dv;
    }
    a.z[p * 128 + tid] = z;
    const float du = dv * w11 * keep * scale;
    a.du[p * 128 + tid] = du;
    dus[tid] = du;
    __syncthreads();
    if (tid < 64) {
        float d = 0.0f;
        for (int i = 0; i < 128; i++) d = fmaf(a.fc10[(size_t)i * 64 + tid], dus[i], d);
        d = h9s[tid] > 0.0f ? d : 0.0f;
        dps[tid] = d;
        a.dpre9[p * 64 + tid] = d;
    }
    __syncthreads();
    float *dst = a.dact8 + (size_t)p * 128 * 64;
    float amax = 0.0f;
    for (int i = tid; i < 128 * 64; i += 128) {
        const int c = i >> 6, cell = i & 63, y = cell >> 3, x = cell & 7;
        float g = 0.0f;
        // act8[c][cell] feeds pre9[cell'] with cell' = cell - (tap offset): tap (ky,kx) of output (y - ky + 1, x - kx + 1)
#pragma unroll
        for (int t = 0; t < 9; t++) {
            const int yy = y - (t / 3 - 1), xx = x - (t % 3 - 1);
            if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) g = fmaf(w9s[c * 9 + t], dps[yy * 8 + xx], g);
        }
        g = a8[c][cell] > 0.0f ? g : 0.0f;
        dst[i] = g;
        amax = fmaxf(amax, fabsf(g));
    }
    publish_absmax(amax, a.dymax);
}

// Weight gradients of the value head, every sum in a fixed order.  Blocks [0,128): dW9[c][9 taps]; block 128: db9 and the loss
// numerator; blocks [129, 129+32): dfc10 rows 4 per block; block 161: dfc11.
__global__ void __launch_bounds__(256) value_head_grad_kernel(const float *__restrict__ act8, const float *__restrict__ h9,
                                                              const float *__restrict__ du, const float *__restrict__ z,
                                                              const float *__restrict__ dpre9, const float *__restrict__ dv,
                                                              const float *__restrict__ loss_terms, float *__restrict__ gw9,
                                                              float *__restrict__ gb9, float *__restrict__ gfc10, float *__restrict__ gfc11,
                                                              float *__restrict__ gloss, long long m, int accumulate) {
    __shared__ float red[256];
    const int blk = blockIdx.x, tid = threadIdx.x;
    auto reduce_store = [&](float v, float *dst) {
        red[tid] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) *dst = accumulate ? *dst + red[0] : red[0];
        __syncthreads();
    };
    if (blk < 128) {
        float acc[9];
#pragma unroll
        for (int t = 0; t < 9; t++) acc[t] = 0.0f;
        for (long long i = tid; i < m * 64; i += 256) {           // (position, cell)
            const long long p = i >> 6;
            const int cell = (int)(i & 63), y = cell >> 3, x = cell & 7;
            const float d = dpre9[i];
            const float *a = act8 + ((size_t)p * 128 + blk) * 64;
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) acc[t] = fmaf(d, a[yy * 8 + xx], acc[t]);
            }
        }
        for (int t = 0; t < 9; t++) reduce_store(acc[t], gw9 + blk * 9 + t);
    } else if (blk == 128) {
        float s = 0.0f, l = 0.0f;
        for (long long i = tid; i < m * 64; i += 256) s += dpre9[i];
        for (long long p = tid; p < m; p += 256) l += loss_terms[p];
        reduce_store(s, gb9);
        reduce_store(l, gloss);
    } else if (blk < 161) {
        // dfc10[i][j] = sum_p du[p][i] * h9[p][j]: 4 rows i per block, thread = (row, j)
        const int i = (blk - 129) * 4 + (tid >> 6), j = tid & 63;
        float s = 0.0f;
        for (long long p = 0; p < m; p++) s = fmaf(du[p * 128 + i], h9[p * 64 + j], s);
        float *dst = gfc10 + (size_t)i * 64 + j;
        *dst = accumulate ? *dst + s : s;
    } else {
        if (tid < 128) {
            float s = 0.0f;
            for (long long p = 0; p < m; p++) s = fmaf(dv[p], z[p * 128 + tid], s);
            gfc11[tid] = accumulate ? gfc11[tid] + s : s;
        }
    }
}

// Evaluation helpers.  policy: values = probabilities (or logits, is_logits) [n][64]; out[0] += sum of softmax_cross_entropy(pred, y)
// with pred the PROBABILITIES (the reference applies log-softmax to them again, train_policy.py:62,69), out[1] += correct arg-maxes
// (F.accuracy).  value: out[0] += sum (v - y)^2.  One block, fixed order.
__global__ void __launch_bounds__(256) policy_eval_kernel(const float *__restrict__ values, const int8_t *__restrict__ action, long long n,
                                                          int is_logits, float *__restrict__ out) {
    __shared__ float red[2][256];
    const int tid = threadIdx.x;
    float loss = 0.0f, hits = 0.0f;
    for (long long p = tid; p < n; p += 256) {
        const float *v = values + p * 64;
        float pr[64];
        float mx = v[0];
        int arg = 0;
        for (int i = 1; i < 64; i++) if (v[i] > mx) { mx = v[i]; arg = i; }
        if (is_logits) {
            float sum = 0.0f;
            for (int i = 0; i < 64; i++) { pr[i] = expf(v[i] - mx); sum += pr[i]; }
            for (int i = 0; i < 64; i++) pr[i] /= sum;
        } else {
            for (int i = 0; i < 64; i++) pr[i] = v[i];
        }
        float mx2 = pr[0];
        for (int i = 1; i < 64; i++) mx2 = fmaxf(mx2, pr[i]);
        float s2 = 0.0f;
        for (int i = 0; i < 64; i++) s2 += expf(pr[i] - mx2);
        const int y = action[p];
        loss += -(pr[y] - mx2 - logf(s2));
        hits += arg == y ? 1.0f : 0.0f;
    }
    red[0][tid] = loss; red[1][tid] = hits;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) { red[0][tid] += red[0][tid + o]; red[1][tid] += red[1][tid + o]; }
        __syncthreads();
    }
    if (tid == 0) { out[0] += red[0][0]; out[1] += red[1][0]; }
}

__global__ void __launch_bounds__(256) value_eval_kernel(const float *__restrict__ pred, const float *__restrict__ target, long long n,
                                                         float *__restrict__ out) {
    __shared__ float red[256];
    const int tid = threadIdx.x;
    float s = 0.0f;
    for (long long p = tid; p < n; p += 256) s += (pred[p] - target[p]) * (pred[p] - target[p]);
    red[tid] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) out[0] += red[0];
}

// ---------------------------------------------------------------- RolloutPolicy (network.py:49-64) supervised gradient, train_policy.py --policy rollout
// One CTA of 64 threads (= cells) walks positions blockIdx.x, blockIdx.x + gridDim.x, ...; logits = conv1 (2 -> 1, 3x3, no bias) +
// bias2[cell] straight from the bitboards, pred = softmax, loss = softmax_cross_entropy(pred, y) (log-softmax applied to the
// probabilities again, as for the SL policy), gradients accumulated per thread and reduced once per CTA in a fixed order.
// partial[cta] = dW[18] | db[64] | loss numerator.
__global__ void __launch_bounds__(64) rollout_grad_kernel(const u64 *__restrict__ own, const u64 *__restrict__ opp, const int8_t *__restrict__ action,
                                                          const float *__restrict__ params, float *__restrict__ partial, long long m) {
    __shared__ float red[2];
    __shared__ float wred[64];
    const int cell = threadIdx.x, y = cell >> 3, x = cell & 7;
    float W[18];
#pragma unroll
    for (int i = 0; i < 18; i++) W[i] = params[i];
    const float b = params[18 + cell];
    float gw[18], gb = 0.0f, gl = 0.0f;
#pragma unroll
    for (int i = 0; i < 18; i++) gw[i] = 0.0f;
    for (long long p = blockIdx.x; p < m; p += gridDim.x) {
        const u64 pl[2] = {opp[p], own[p]};   // channel 0 = opponent stones, channel 1 = mover's stones
        float xs[18];
        float s[2] = {0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int t = 0; t < 9; t++) {
                const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
                const bool in = yy >= 0 && yy < 8 && xx >= 0 && xx < 8;
                const float v = in ? (float)((pl[c] >> (yy * 8 + xx)) & 1) : 0.0f;
                xs[c * 9 + t] = v;
                if (v != 0.0f) s[c] = __fadd_rn(s[c], W[c * 9 + t]);
            }
        const float logit = __fadd_rn(__fadd_rn(s[0], s[1]), b);
        const float mx = block64_max(logit, red);
        const float e = expf(logit - mx);
        const float pred = e / block64_sum(e, red);
        const float mx2 = block64_max(pred, red);
        const float e2 = expf(pred - mx2);
        const float s2 = block64_sum(e2, red);
        const float q = e2 / s2;
        const int ya = action[p];
        if (cell == ya) gl += -(pred - mx2 - logf(s2));
        const float dpred = q - (cell == ya ? 1.0f : 0.0f);
        const float dot = block64_sum(dpred * pred, red);
        const float dl = pred * (dpred - dot);
        gb += dl;
#pragma unroll
        for (int i = 0; i < 18; i++) gw[i] = fmaf(dl, xs[i], gw[i]);
    }
    float *dst = partial + (size_t)blockIdx.x * 83;
    dst[18 + cell] = gb;
    for (int i = 0; i < 19; i++) {
        wred[cell] = i < 18 ? gw[i] : gl;
        __syncthreads();
        for (int o = 32; o > 0; o >>= 1) {
            if (cell < o) wred[cell] += wred[cell + o];
            __syncthreads();
        }
        if (cell == 0) dst[i < 18 ? i : 82] = wred[0];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- WeightDecay hook + Adam (Chainer: optimizers.Adam, optimizer_hooks.WeightDecay)
__global__ void adam_kernel(float *__restrict__ p, float *__restrict__ mo, float *__restrict__ ve, const float *__restrict__ g,
                            int n, float inv_count, float wd, float beta1, float beta2, float eps, float lr_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w = p[i];
    const float grad = fmaf(wd, w, g[i] * inv_count);       // hook: grad += rate * param
    const float m1 = mo[i] + (1.0f - beta1) * (grad - mo[i]);
    const float v1 = ve[i] + (1.0f - beta2) * (grad * grad - ve[i]);
    mo[i] = m1;
    ve[i] = v1;
    p[i] = w - lr_t * m1 / (sqrtf(v1) + eps);
}

// The same update with the divisor read on the device: count = g[n + 1] of the all-reduced vector [gradient | loss numerator | count]
// (no host round trip between the all-reduce and the update).  A count of zero (no position on any rank) leaves the parameters alone.
__global__ void adam_dev_kernel(float *__restrict__ p, float *__restrict__ mo, float *__restrict__ ve, const float *__restrict__ g,
                                int n, float wd, float beta1, float beta2, float eps, float lr_t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float count = g[n + 1];
    if (!(count > 0.0f)) return;
    const float w = p[i];
    const float grad = fmaf(wd, w, g[i] * (1.0f / count));
    const float m1 = mo[i] + (1.0f - beta1) * (grad - mo[i]);
    const float v1 = ve[i] + (1.0f - beta2) * (grad * grad - ve[i]);
    mo[i] = m1;
    ve[i] = v1;
    p[i] = w - lr_t * m1 / (sqrtf(v1) + eps);
}

__global__ void set_count_kernel(float *dst, float m, int accumulate) { *dst = accumulate ? *dst + m : m; }

}  // namespace iago

using namespace iago;

constexpr int kNPValue = 970049;  // Value parameters

struct iago_trainer {
    iago_ctx *ctx = nullptr;
    int kind = 0;                     // 0 = SLPolicy (REINFORCE / supervised policy), 1 = Value (train_value.py)
    int np = kNP;                     // parameter count
    int max_pos = 0;
    float *params = nullptr, *adam_m = nullptr, *adam_v = nullptr;  // [kNP]
    long long t = 0;
    float *wf[8] = {}, *wd[8] = {};   // conv-kernel weight layouts
    float *act[9] = {};               // act[0] = input planes, act[l] = output of block l
    float *dyb[8] = {};               // dyb[l]: gradient w.r.t. the pre-activation output of block l+1, [max_pos][cout_l][64]
    uint8_t *bwd_blob = nullptr;      // bf16 hi/lo weight units of the data-gradient chain (trunk.cu, backward mode)
    int bwd_precision = 3;
    float *dlogit = nullptr, *loss_terms = nullptr, *partial = nullptr, *bias_partial = nullptr;
    unsigned *dymax = nullptr;        // [8] bit pattern of max |dY| per layer, collected by the kernels that write dyb[l]
    size_t partial_stride = 0;
    int slices = 0;
    int tc_slices = 0, slices0 = 0;
    bool use_tc = true, tc_attr = false;
    int synced_slot = -1;             // trunk slot that holds the CURRENT parameters (-1: stale)
    uint8_t *ones = nullptr;          // colour array (all 1) for the trunk launch
    float *logits_scratch = nullptr;
    size_t w_off[8], b_off[8], w9_off, b10_off;      // policy head: conv9/W [128], bias10/b [64]
    size_t vw9_off = 0, vb9_off = 0, fc10_off = 0, fc11_off = 0;   // value head: block9/conv/W [1152], /b [1], fc10/W [128][64], fc11/W [128]
    float *h9 = nullptr, *du = nullptr, *zbuf = nullptr, *dpre9 = nullptr, *dv = nullptr, *vpred = nullptr;
    std::vector<void *> allocs;
};

template <class T_>
static int tr_alloc(iago_trainer *t, T_ **p, size_t count) {
    IAGO_CUDA(cudaMalloc((void **)p, count * sizeof(T_)));
    IAGO_CUDA(cudaMemset(*p, 0, count * sizeof(T_)));
    t->allocs.push_back(*p);
    return IAGO_OK;
}

static void relayout_all(iago_trainer *t, cudaStream_t s) {
    for (int l = 0; l < 8; l++) {
        const int n = kCout[l] * kCin[l] * 9;
        relayout_kernel<<<(n + 255) / 256, 256, 0, s>>>(t->params + t->w_off[l], t->wf[l], l > 0 ? t->wd[l] : nullptr, kCin[l], kCout[l], 0);
    }
}

extern "C" {

int iago_trainer_create(iago_ctx *ctx, int kind, const float *params, int64_t n_floats, int max_positions, iago_trainer **out) {
    IAGO_REQUIRE(ctx && params && out, "NULL argument");
    IAGO_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (SL policy) or 1 (value)");
    IAGO_REQUIRE(n_floats == (kind == 0 ? kNP : kNPValue), "parameter vector must hold 960,768 (policy) / 970,049 (value) floats (iago_load_net order)");
    IAGO_REQUIRE(max_positions > 0 && max_positions <= (1 << 20), "max_positions out of range");
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    iago_trainer *t = new iago_trainer();
    t->ctx = ctx;
    t->kind = kind;
    t->np = (int)n_floats;
    t->max_pos = max_positions;
    size_t off = 0;
    for (int l = 0; l < 8; l++) {
        t->w_off[l] = off; off += (size_t)kCout[l] * kCin[l] * 9;
        t->b_off[l] = off; off += kCout[l];
    }
    if (kind == 0) {
        t->w9_off = off; off += 128;
        t->b10_off = off; off += 64;
    } else {
        t->w9_off = t->b10_off = 0;
        t->vw9_off = off; off += 1152;
        t->vb9_off = off; off += 1;
        t->fc10_off = off; off += 128 * 64;
        t->fc11_off = off; off += 128;
    }
    int rc = 0;
    const size_t M = max_positions;
#define A(ptr, cnt) if (!rc) rc = tr_alloc(t, &(ptr), (cnt))
    A(t->params, t->np); A(t->adam_m, t->np); A(t->adam_v, t->np);
    for (int l = 0; l < 8; l++) {
        A(t->wf[l], (size_t)kCout[l] * kCin[l] * 9);
        if (l > 0) A(t->wd[l], (size_t)kCout[l] * kCin[l] * 9);
    }
    A(t->act[0], M * 2 * 64);
    for (int l = 0; l < 8; l++) A(t->act[l + 1], M * kCout[l] * 64);
    for (int l = 0; l < 8; l++) A(t->dyb[l], M * kCout[l] * 64);
    A(t->bwd_blob, trunk_backward_blob_bytes());
    A(t->dlogit, M * 64); A(t->loss_terms, M); A(t->ones, M); A(t->logits_scratch, M * 64);
    if (kind == 1) {
        A(t->h9, M * 64); A(t->du, M * 128); A(t->zbuf, M * 128); A(t->dpre9, M * 64); A(t->dv, M); A(t->vpred, M);
    }
    A(t->dymax, 8);
    t->slices = 24;                                  // fp32 weight-gradient kernel: position slices of the 64/128-input-channel layers
    t->slices0 = 2 * ctx->sm_count;                  // ... and of block 1 (2 input channels: one c tile, so the slices are the whole grid)
    t->tc_slices = ctx->sm_count >= 3 ? ctx->sm_count / 3 : 1;   // 3 kernel rows x slices <= one CTA per SM: the kernel's 207 KB of shared
                                                                 // memory allow one CTA per SM, so rounding UP (150 CTAs on 148 SMs) ran
                                                                 // every launch as two waves — twice the time of 147 CTAs
    t->partial_stride = (size_t)128 * 128 * 9 + 128;
    A(t->partial, (size_t)(t->tc_slices > t->slices ? t->tc_slices : t->slices) * t->partial_stride);
    A(t->bias_partial, (size_t)t->tc_slices * 128);
#undef A
    if (rc) {
        for (void *p : t->allocs) cudaFree(p);
        delete t;
        return rc;
    }
    IAGO_CUDA(cudaMemcpy(t->params, params, (size_t)t->np * 4, cudaMemcpyHostToDevice));
    IAGO_CUDA(cudaMemset(t->ones, 1, M));
    *out = t;
    return IAGO_OK;
}

int iago_reinforce_create(iago_ctx *ctx, const float *params, int64_t n_floats, int max_positions, iago_trainer **out) {
    return iago_trainer_create(ctx, 0, params, n_floats, max_positions, out);
}

int iago_reinforce_destroy(iago_trainer *t) {
    if (!t) return IAGO_OK;
    DeviceGuard guard(t->ctx->device);
    cudaDeviceSynchronize();
    for (void *p : t->allocs) cudaFree(p);
    delete t;
    return IAGO_OK;
}

// Forward through the 8 blocks with every block's output kept in act[1..8] (fp32 [m][C][64]).
static int forward_acts(iago_trainer *t, const uint64_t *own, const uint64_t *opp, int64_t m, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (!(t->use_tc && t->synced_slot >= 0)) relayout_all(t, s);   // Wf / Wd feed the fp32 conv kernels only
    const unsigned tiles = (unsigned)((m + 1) / 2);
    planes_kernel<<<(unsigned)((m * 64 + 255) / 256), 256, 0, s>>>((const u64 *)own, (const u64 *)opp, t->act[0], m);
    if (t->use_tc && t->synced_slot >= 0) {
        // the fused tcgen05 trunk (trunk.cu) on the slot that holds these parameters, dumping every block's output in fp32
        float *out = t->kind == 0 ? t->logits_scratch : t->vpred;
        int rc = trunk_launch(t->ctx, t->synced_slot, t->kind, own, opp, t->ones, m, out, 0, 3, stream, nullptr, t->act + 1);
        if (rc) return rc;
    } else {
        for (int l = 0; l < 8; l++) {
            const float *b = t->params + t->b_off[l];
            if (kCout[l] == 64)
                conv3x3_kernel<64, EPI_FWD><<<dim3(tiles, 1), 256, 0, s>>>(t->act[l], t->wf[l], b, nullptr, t->act[l + 1], m, kCin[l], kCout[l]);
            else
                conv3x3_kernel<128, EPI_FWD><<<dim3(tiles, 1), 256, 0, s>>>(t->act[l], t->wf[l], b, nullptr, t->act[l + 1], m, kCin[l], kCout[l]);
        }
    }
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

// Backward through the 8 blocks from dyb[7] (gradient w.r.t. block 8's pre-activation): weight / bias gradients into `grad`.
static int backward_trunk(iago_trainer *t, int64_t m, float *grad, int accumulate, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned tiles = (unsigned)((m + 1) / 2);
    if (t->use_tc) {
        // data gradients of blocks 8..2 as ONE fused tcgen05 launch (bf16 hi/lo, the tile stays on chip between layers)
        const float *W[8], *mask[7];
        float *dx[7];
        for (int l = 0; l < 8; l++) W[l] = t->params + t->w_off[l];
        for (int i = 0; i < 7; i++) {
            mask[i] = t->act[7 - i];
            dx[i] = t->dyb[6 - i];
        }
        int rc = trunk_backward_pack(t->ctx, W, t->bwd_blob, stream);
        if (rc) return rc;
        unsigned *dymax[7];
        for (int i = 0; i < 7; i++) dymax[i] = kCout[6 - i] == 128 ? t->dymax + (6 - i) : nullptr;   // dx[i] = dY of block 7 - i (index 6 - i)
        rc = trunk_backward_launch(t->ctx, t->bwd_blob, t->dyb[7], mask, dx, m, t->bwd_precision, stream, dymax);
        if (rc) return rc;
    }
    for (int l = 7; l >= 0; l--) {
        const int want = l == 0 ? t->slices0 : t->slices;
        const int pos_per_slice = (int)((m + want - 1) / want);
        const int slices = (int)((m + pos_per_slice - 1) / pos_per_slice);
        const float *dy = t->dyb[l];
        if (t->use_tc && kCout[l] == 128) {
            if (!t->tc_attr) {
                IAGO_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
                t->tc_attr = true;
            }
            const int pps = (int)((m + t->tc_slices - 1) / t->tc_slices);
            const int sl = (int)((m + pps - 1) / pps);
            // dymax[l] was collected by the kernel that wrote dyb[l] (the head for l = 7, the fused backward chain below it)
            wgrad_tc_kernel<<<dim3(sl, 3), kWgThreads, kWgSmem, s>>>(t->act[l], dy, t->partial, m, kCin[l], pps, t->partial_stride, t->dymax + l,
                                                                     t->bias_partial);
            reduce_slices_kernel<<<1, 256, 0, s>>>(t->bias_partial, grad + t->b_off[l], kCout[l], sl, (size_t)kCout[l], accumulate);
            reduce_taps_kernel<<<(9 * 128 * kCin[l] + 255) / 256, 256, 0, s>>>(t->partial, grad + t->w_off[l], kCin[l], sl, t->partial_stride, accumulate);
        } else {
            const int count = kCout[l] * kCin[l] * 9 + kCout[l];  // W then b are adjacent in the flat layout as well
            const size_t stride = l == 0 ? (size_t)count : t->partial_stride;
            if (l == 0 && kCin[0] == 2 && kCout[0] == 64)
                wgrad_block1_kernel<<<slices, 256, 0, s>>>(t->act[l], dy, t->partial, m, pos_per_slice, stride);
            else
                wgrad3x3_kernel<<<dim3((kCin[l] + 15) / 16, slices), 256, 0, s>>>(t->act[l], dy, t->partial, m, kCin[l], kCout[l], pos_per_slice, stride);
            reduce_slices_kernel<<<(count + 255) / 256, 256, 0, s>>>(t->partial, grad + t->w_off[l], count, slices, stride, accumulate);
        }
        if (!t->use_tc && l > 0) {
            float *dx = t->dyb[l - 1];
            if (kCin[l] == 64)
                conv3x3_kernel<64, EPI_DGRAD><<<dim3(tiles, 1), 256, 0, s>>>(dy, t->wd[l], nullptr, t->act[l], dx, m, kCout[l], kCin[l]);
            else
                conv3x3_kernel<128, EPI_DGRAD><<<dim3(tiles, 1), 256, 0, s>>>(dy, t->wd[l], nullptr, t->act[l], dx, m, kCout[l], kCin[l]);
        }
    }
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_reinforce_grad(iago_trainer *t, const uint64_t *own, const uint64_t *opp, const int8_t *action, const float *reward,
                        int64_t m, float *grad, int accumulate, float *probs_out, void *stream) {
    IAGO_REQUIRE(t && own && opp && action && reward && grad, "NULL argument");
    IAGO_REQUIRE(t->kind == 0, "this trainer holds a Value network (use iago_value_grad)");
    IAGO_REQUIRE(m >= 0 && m <= t->max_pos, "m exceeds the trainer's max_positions");
    if (m == 0) return IAGO_OK;
    DeviceGuard guard(t->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = forward_acts(t, own, opp, m, stream);
    if (rc) return rc;
    // ---- head forward + backward
    IAGO_CUDA(cudaMemsetAsync(t->dymax, 0, 8 * sizeof(unsigned), s));
    head_kernel<<<(unsigned)m, 64, 0, s>>>(t->act[8], t->params + t->w9_off, t->params + t->b10_off, action, reward, t->dlogit,
                                          t->dyb[7], t->loss_terms, probs_out, m, t->dymax + 7);
    head_grad_kernel<<<dim3(193, kHeadSlices), 256, 0, s>>>(t->act[8], t->dlogit, t->loss_terms, t->partial, m);
    head_grad_reduce_kernel<<<1, 256, 0, s>>>(t->partial, grad + t->w9_off, grad + t->b10_off, grad + kNP, accumulate);
    IAGO_CUDA(cudaGetLastError());
    rc = backward_trunk(t, m, grad, accumulate, stream);
    if (rc) return rc;
    // the position count rides along with the gradient (element kNP + 1), so ranks all-reduce numerator and count together
    set_count_kernel<<<1, 1, 0, s>>>(grad + kNP + 1, (float)m, accumulate);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_value_grad(iago_trainer *t, const uint64_t *own, const uint64_t *opp, const float *target, int64_t m, float *grad,
                    int accumulate, double dropout_ratio, uint64_t dropout_seed, uint64_t position_id0, float *pred_out,
                    uint8_t *mask_out, void *stream) {
    IAGO_REQUIRE(t && own && opp && target && grad, "NULL argument");
    IAGO_REQUIRE(t->kind == 1, "this trainer holds an SL policy network (use iago_reinforce_grad)");
    IAGO_REQUIRE(m >= 0 && m <= t->max_pos, "m exceeds the trainer's max_positions");
    IAGO_REQUIRE(dropout_ratio >= 0.0 && dropout_ratio < 1.0, "dropout_ratio must lie in [0, 1)");
    if (m == 0) return IAGO_OK;
    DeviceGuard guard(t->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc = forward_acts(t, own, opp, m, stream);
    if (rc) return rc;
    const float *P = t->params;
    ValueHeadArgs a{t->act[8], P + t->vw9_off, P + t->vb9_off, P + t->fc10_off, P + t->fc11_off, target,
                    t->vpred, t->loss_terms, t->h9, t->du, t->zbuf, t->dpre9, t->dv, t->dyb[7], mask_out, t->dymax + 7, m,
                    (float)dropout_ratio, dropout_seed, position_id0};
    IAGO_CUDA(cudaMemsetAsync(t->dymax, 0, 8 * sizeof(unsigned), s));
    value_head_kernel<<<(unsigned)m, 128, 0, s>>>(a);
    value_head_grad_kernel<<<162, 256, 0, s>>>(t->act[8], t->h9, t->du, t->zbuf, t->dpre9, t->dv, t->loss_terms, grad + t->vw9_off,
                                               grad + t->vb9_off, grad + t->fc10_off, grad + t->fc11_off, grad + t->np, m, accumulate);
    IAGO_CUDA(cudaGetLastError());
    if (pred_out) IAGO_CUDA(cudaMemcpyAsync(pred_out, t->vpred, (size_t)m * 4, cudaMemcpyDeviceToDevice, s));
    rc = backward_trunk(t, m, grad, accumulate, stream);
    if (rc) return rc;
    set_count_kernel<<<1, 1, 0, s>>>(grad + t->np + 1, (float)m, accumulate);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

/* sums into out[0] (loss) and out[1] (correct arg-maxes); the caller zeroes `out` (device float[2]) */
int iago_policy_eval(iago_ctx *ctx, const float *values, int is_logits, const int8_t *action, int64_t n, float *out, void *stream) {
    IAGO_REQUIRE(ctx && values && action && out, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    policy_eval_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(values, action, n, is_logits, out);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_value_eval(iago_ctx *ctx, const float *pred, const float *target, int64_t n, float *out, void *stream) {
    IAGO_REQUIRE(ctx && pred && target && out, "NULL argument");
    IAGO_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return IAGO_OK;
    DeviceGuard guard(ctx->device);
    value_eval_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pred, target, n, out);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_reinforce_adam_step(iago_trainer *t, const float *grad, double count, double alpha, double beta1, double beta2, double eps,
                             double weight_decay, void *stream) {
    IAGO_REQUIRE(t && grad, "NULL argument");
    IAGO_REQUIRE(count > 0, "count must be positive");
    DeviceGuard guard(t->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    t->t += 1;
    t->synced_slot = -1;   // the slot no longer holds these parameters until iago_reinforce_sync_slot
    const double fix1 = 1.0 - pow(beta1, (double)t->t), fix2 = 1.0 - pow(beta2, (double)t->t);
    const float lr_t = (float)(alpha * sqrt(fix2) / fix1);   // Chainer AdamRule.lr
    adam_kernel<<<(t->np + 255) / 256, 256, 0, s>>>(t->params, t->adam_m, t->adam_v, grad, t->np, (float)(1.0 / count), (float)weight_decay,
                                                  (float)beta1, (float)beta2, (float)eps, lr_t);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_reinforce_adam_step_dev(iago_trainer *t, const float *grad, double alpha, double beta1, double beta2, double eps, double weight_decay,
                                 void *stream) {
    IAGO_REQUIRE(t && grad, "NULL argument");
    DeviceGuard guard(t->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    t->t += 1;
    t->synced_slot = -1;
    const double fix1 = 1.0 - pow(beta1, (double)t->t), fix2 = 1.0 - pow(beta2, (double)t->t);
    const float lr_t = (float)(alpha * sqrt(fix2) / fix1);   // Chainer AdamRule.lr
    adam_dev_kernel<<<(t->np + 255) / 256, 256, 0, s>>>(t->params, t->adam_m, t->adam_v, grad, t->np, (float)weight_decay, (float)beta1, (float)beta2,
                                                      (float)eps, lr_t);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_reinforce_get_state(iago_trainer *t, float *params, float *adam_m, float *adam_v, int64_t *step) {
    IAGO_REQUIRE(t, "NULL argument");
    DeviceGuard guard(t->ctx->device);
    IAGO_CUDA(cudaDeviceSynchronize());
    if (params) IAGO_CUDA(cudaMemcpy(params, t->params, (size_t)t->np * 4, cudaMemcpyDeviceToHost));
    if (adam_m) IAGO_CUDA(cudaMemcpy(adam_m, t->adam_m, (size_t)t->np * 4, cudaMemcpyDeviceToHost));
    if (adam_v) IAGO_CUDA(cudaMemcpy(adam_v, t->adam_v, (size_t)t->np * 4, cudaMemcpyDeviceToHost));
    if (step) *step = t->t;
    return IAGO_OK;
}

int iago_reinforce_set_state(iago_trainer *t, const float *params, const float *adam_m, const float *adam_v, int64_t step) {
    IAGO_REQUIRE(t, "NULL argument");
    DeviceGuard guard(t->ctx->device);
    IAGO_CUDA(cudaDeviceSynchronize());
    if (params) {
        IAGO_CUDA(cudaMemcpy(t->params, params, (size_t)t->np * 4, cudaMemcpyHostToDevice));
        t->synced_slot = -1;
    }
    if (adam_m) IAGO_CUDA(cudaMemcpy(t->adam_m, adam_m, (size_t)t->np * 4, cudaMemcpyHostToDevice));
    if (adam_v) IAGO_CUDA(cudaMemcpy(t->adam_v, adam_v, (size_t)t->np * 4, cudaMemcpyHostToDevice));
    if (step >= 0) t->t = step;
    return IAGO_OK;
}

/* ---- RolloutPolicy trainer: 82 parameters (conv1/W [1][2][3][3] then bias2/b [64]); gradient vector = 82 | loss numerator | count */
struct iago_rollout_trainer {
    iago_ctx *ctx = nullptr;
    float *params = nullptr, *adam_m = nullptr, *adam_v = nullptr, *partial = nullptr;
    int ctas = 0;
    long long t = 0;
};

int iago_rollout_trainer_create(iago_ctx *ctx, const float *conv1_W, const float *bias2_b, iago_rollout_trainer **out) {
    IAGO_REQUIRE(ctx && conv1_W && bias2_b && out, "NULL argument");
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    iago_rollout_trainer *t = new iago_rollout_trainer();
    t->ctx = ctx;
    t->ctas = 8 * ctx->sm_count;
    cudaError_t e = cudaMalloc(&t->params, 3 * 82 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&t->partial, (size_t)t->ctas * 83 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(t->params, 0, 3 * 82 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(t->params, conv1_W, 18 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t->params + 18, bias2_b, 64 * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(t->params); cudaFree(t->partial);
        delete t;
        IAGO_CUDA(e);
    }
    t->adam_m = t->params + 82;
    t->adam_v = t->params + 164;
    *out = t;
    return IAGO_OK;
}

int iago_rollout_trainer_destroy(iago_rollout_trainer *t) {
    if (!t) return IAGO_OK;
    DeviceGuard guard(t->ctx->device);
    cudaDeviceSynchronize();
    cudaFree(t->params); cudaFree(t->partial);
    delete t;
    return IAGO_OK;
}

int iago_rollout_trainer_grad(iago_rollout_trainer *t, const uint64_t *own, const uint64_t *opp, const int8_t *action, int64_t m,
                              float *grad, int accumulate, void *stream) {
    IAGO_REQUIRE(t && own && opp && action && grad, "NULL argument");
    IAGO_REQUIRE(m >= 0, "m < 0");
    if (m == 0) return IAGO_OK;
    DeviceGuard guard(t->ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int ctas = (int)(m < t->ctas ? m : t->ctas);
    rollout_grad_kernel<<<ctas, 64, 0, s>>>((const u64 *)own, (const u64 *)opp, action, t->params, t->partial, m);
    reduce_slices_kernel<<<1, 128, 0, s>>>(t->partial, grad, 83, ctas, 83, accumulate);
    set_count_kernel<<<1, 1, 0, s>>>(grad + 83, (float)m, accumulate);
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

int iago_rollout_trainer_adam_step(iago_rollout_trainer *t, const float *grad, double count, double alpha, double beta1, double beta2,
                                   double eps, double weight_decay, void *stream) {
    IAGO_REQUIRE(t && grad, "NULL argument");
    IAGO_REQUIRE(count > 0, "count must be positive");
    DeviceGuard guard(t->ctx->device);
    t->t += 1;
    const double fix1 = 1.0 - pow(beta1, (double)t->t), fix2 = 1.0 - pow(beta2, (double)t->t);
    adam_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(t->params, t->adam_m, t->adam_v, grad, 82, (float)(1.0 / count), (float)weight_decay,
                                                    (float)beta1, (float)beta2, (float)eps, (float)(alpha * sqrt(fix2) / fix1));
    IAGO_CUDA(cudaGetLastError());
    return IAGO_OK;
}

/* state: HOST float[3][82] = params | adam m | adam v */
int iago_rollout_trainer_get_state(iago_rollout_trainer *t, float *state, int64_t *step) {
    IAGO_REQUIRE(t && state, "NULL argument");
    DeviceGuard guard(t->ctx->device);
    IAGO_CUDA(cudaDeviceSynchronize());
    IAGO_CUDA(cudaMemcpy(state, t->params, 3 * 82 * sizeof(float), cudaMemcpyDeviceToHost));
    if (step) *step = t->t;
    return IAGO_OK;
}

int iago_rollout_trainer_set_state(iago_rollout_trainer *t, const float *state, int64_t step) {
    IAGO_REQUIRE(t && state, "NULL argument");
    DeviceGuard guard(t->ctx->device);
    IAGO_CUDA(cudaDeviceSynchronize());
    IAGO_CUDA(cudaMemcpy(t->params, state, 3 * 82 * sizeof(float), cudaMemcpyHostToDevice));
    if (step >= 0) t->t = step;
    return IAGO_OK;
}

int iago_reinforce_set_option(iago_trainer *t, int use_tensor_cores) {
    IAGO_REQUIRE(t, "NULL argument");
    t->use_tc = use_tensor_cores != 0;
    return IAGO_OK;
}

int iago_reinforce_sync_slot(iago_trainer *t, int slot) {
    IAGO_REQUIRE(t, "NULL argument");
    int rc;
    if (trunk_slot_holds(t->ctx, slot, t->kind)) {
        // the slot's buffers exist: repack on the device (after whatever stream the Adam step ran on has drained)
        DeviceGuard guard(t->ctx->device);
        IAGO_CUDA(cudaDeviceSynchronize());
        rc = trunk_refresh_slot(t->ctx, slot, t->kind, t->params, t->ctx->stream);
        if (rc == IAGO_OK) IAGO_CUDA(cudaStreamSynchronize(t->ctx->stream));
    } else {
        std::vector<float> h(t->np);
        rc = iago_reinforce_get_state(t, h.data(), nullptr, nullptr, nullptr);
        if (rc) return rc;
        rc = iago_load_net(t->ctx, slot, t->kind, h.data(), t->np);
    }
    if (rc == IAGO_OK) t->synced_slot = slot;
    return rc;
}

int iago_reinforce_sync_slot_async(iago_trainer *t, int slot, void *stream) {
    IAGO_REQUIRE(t, "NULL argument");
    if (!trunk_slot_holds(t->ctx, slot, t->kind)) return iago_reinforce_sync_slot(t, slot);   // first use: the slot's buffers do not exist yet
    // the pack kernels follow the Adam step on the same stream: no host synchronisation
    const int rc = trunk_refresh_slot(t->ctx, slot, t->kind, t->params, stream);
    if (rc == IAGO_OK) t->synced_slot = slot;
    return rc;
}

}  // extern "C"
