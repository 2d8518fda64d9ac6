// Micro-benchmark + known-answer check for the CTA-pair tensor path used by trunk.cu's paired kernel:
//   tcgen05.alloc / mma / commit with cta_group::2 (M = 256 over two CTAs, each CTA holds 128 rows of A and HALF of B's N columns),
//   B halves loaded by 2-D TMA (cp.async.bulk.tensor.2d.cta_group::2) that signals the LEADER CTA's mbarrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_check umma2_check.cu && ./umma2_check
// Part 1 (check): D[256][128] = A[256][64] * B[128][64]^T on small integers, compared with the host: proves which CTA's B half
//   feeds which N columns, the descriptor strides and the multicast commit.
// Part 2 (rate): cycles per "precision-2 unit" (4 fp16 K=16 MMAs + 4 FP8 K=32 MMAs) for cta_group::1 and cta_group::2, alone and
//   with the weight stream of the trunk running into the same shared memory (32 KB per unit and CTA for cta_group::1, 16 KB for
//   cta_group::2), on one cluster and on all SMs.  Tests whether shared-memory bandwidth (MMA operand reads + TMA writes) paces the trunk.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../iago_b200/csrc/tc.cuh"
using namespace iago;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tma2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    // the barrier is the LEADER's (peer bit cleared): both CTAs' copies complete their bytes on it
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu) : "memory");
}

constexpr int K = 64, N = 128;
constexpr int OFF_A = 0, OFF_B = 16384, OFF_BAR = 16384 + 8192;
constexpr int kCheckSmem = OFF_BAR + 64;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
check_kernel(const __half *A /* [256][64] row-major */, const __grid_constant__ CUtensorMap bmap, float *D /* [256][128] */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem), rank = cluster_rank();
    const uint32_t bar_full = sbase + OFF_BAR, bar_acc = bar_full + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BAR + 32);
    const int tid = threadIdx.x;
    // A tile: canonical no-swizzle K-major [kgroup 8][row 128][8 fp16]
    for (int i = tid; i < 128 * 8; i += 128) {
        const int row = i & 127, kg = i >> 7;
        *reinterpret_cast<uint4 *>(smem + OFF_A + (kg * 128 + row) * 16) = *reinterpret_cast<const uint4 *>(A + (size_t)(rank * 128 + row) * K + kg * 8);
    }
    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 32) {
        if (rank == 0) mbar_expect_tx(bar_full, 2 * 8192);
        tma2d_pair(sbase + OFF_B, &bmap, 0, (int)rank * 32, bar_full);   // this CTA's half: 32 rows of 256 B
    }
    if (tid == 64 && rank == 0) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
        const uint32_t idesc = instr_desc(256, N);
        for (int ks = 0; ks < 4; ks++) {
            const uint64_t da = smem_desc(sbase + OFF_A + ks * 2 * 2048, 2048, 128);
            const uint64_t db = smem_desc(sbase + OFF_B + ks * 2 * 1024, 1024, 128);   // half of B: 64 rows per K group
            umma2_f16(tmem, da, db, idesc, ks > 0);
        }
        umma2_commit_mc(bar_acc, 3);
    }
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 32; j++) D[(size_t)(rank * 128 + tid) * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    cluster_sync_all();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
}

// ---------------------------------------------------------------- rate
constexpr int kGroup = 3200, kRowPitch = 160;                 // the trunk's activation tile geometry
constexpr int R_A = 0, R_A2 = 51200, R_RING = 102400;         // A hi tile, A8 | AL8 tiles, 3 x 32 KB (cta_group::1) or 6 x 16 KB (cta_group::2) ring
constexpr int R_B = R_RING;                                   // the MMAs read the unit in ring stage 0 (overwritten by the stream: timing only)
constexpr int R_BAR = R_RING + 98304;
constexpr int kRateSmem = R_BAR + 256;

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(const uint8_t *blob, int units, int with_stream, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = CG == 2 ? cluster_rank() : 0;
    const uint32_t bar_done = sbase + R_BAR, bar_ring = bar_done + 8;   // ring: up to 6 "full" barriers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + R_BAR + 128);
    volatile int *progress = reinterpret_cast<volatile int *>(smem + R_BAR + 160);
    const int tid = threadIdx.x;
    for (int i = tid; i < R_RING / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar_done, 1);
        for (int s = 0; s < 6; s++) mbar_init(bar_ring + 8 * s, 1);
        *progress = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    fence_async_smem();
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int stages = CG == 2 ? 6 : 3, stage_bytes = CG == 2 ? 16384 : 32768;
    if (tid == 32 && with_stream) {
        // the weight stream: one bulk copy per unit into the ring, at most `stages` ahead of the MMA issuer
        for (int u = 0; u < units; u++) {
            const int s = u % stages;
            if (u >= stages) mbar_wait(bar_ring + 8 * s, ((u / stages) - 1) & 1);   // the copy that used this stage has landed
            if (CG == 2 && rank == 1) {   // the leader's progress counter, read through distributed shared memory
                uint32_t remote, v;
                asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(sbase + R_BAR + 160));
                do { asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory"); } while (u >= (int)v + stages);
            } else {
                while (u >= *progress + stages) { }
            }
            mbar_expect_tx(bar_ring + 8 * s, stage_bytes);
            bulk_g2s(sbase + R_RING + s * stage_bytes, blob + ((size_t)(u % 100) * 32768 + rank * 16384), stage_bytes, bar_ring + 8 * s);
        }
        for (int u = units > stages ? units - stages : 0; u < units; u++) mbar_wait(bar_ring + 8 * (u % stages), (u / stages) & 1);
    }
    if (tid == 64 && rank == 0) {
        const uint32_t hiA = (uint32_t)(kRowPitch >> 4) | (1u << 14), hiB = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lbo = (uint32_t)(kGroup >> 4) << 16;
        const uint32_t b_lbo = (uint32_t)((CG == 2 ? 1024 : 2048) >> 4) << 16;
        const uint32_t idesc = instr_desc(CG == 2 ? 256 : 128, 128);
        const uint32_t aw = ((sbase + R_A) >> 4) | a_lbo, a8w = ((sbase + R_A2) >> 4) | a_lbo, al8w = ((sbase + R_A2 + 25600) >> 4) | a_lbo;
        const uint32_t bw = ((sbase + R_B) >> 4) | b_lbo, w8 = ((sbase + R_B + (CG == 2 ? 8192 : 16384)) >> 4) | b_lbo;
        const uint32_t wl8 = ((sbase + R_B + (CG == 2 ? 12288 : 24576)) >> 4) | b_lbo;
        constexpr uint32_t dA = (2 * kGroup) >> 4;
        const uint32_t b_step = (CG == 2 ? 2048u : 4096u) >> 4;
        const long long t0 = clock64();
#pragma unroll 1
        for (int u = 0; u < units; u++) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
                const uint64_t da = pack64(aw + ks * dA, hiA), db = pack64(bw + ks * b_step, hiB);
                if (CG == 2) umma2_f16(tmem, da, db, idesc, 1); else umma_f16(tmem, da, db, idesc, 1);
            }
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                const uint64_t da8 = pack64(a8w + ks * dA, hiA), dal8 = pack64(al8w + ks * dA, hiA);
                const uint64_t dw8 = pack64(w8 + ks * b_step, hiB), dwl8 = pack64(wl8 + ks * b_step, hiB);
                if (CG == 2) { umma2_f8(tmem + 256, dal8, dw8, idesc, 1); umma2_f8(tmem + 256, da8, dwl8, idesc, 1); }
                else { umma_f8(tmem + 256, dal8, dw8, idesc, 1); umma_f8(tmem + 256, da8, dwl8, idesc, 1); }
            }
            *progress = u + 1;
        }
        const long long t1 = clock64();
        if (CG == 2) umma2_commit_mc(bar_done, 1); else umma_commit(bar_done);
        mbar_wait(bar_done, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (tid < 32) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    // ---------------- part 1: known-answer check
    std::vector<__half> hA(256 * K), hBp(2 * 8 * 64 * 8);
    std::vector<float> fA(256 * K), fB(N * K);
    for (int m = 0; m < 256; m++)
        for (int k = 0; k < K; k++) { fA[m * K + k] = (float)(((m * 7 + k * 3) % 5) - 2); hA[m * K + k] = __float2half(fA[m * K + k]); }
    for (int n = 0; n < N; n++)
        for (int k = 0; k < K; k++) fB[n * K + k] = (float)(((n * 5 + k) % 7) - 3);
    // packed B: [half h][kgroup 8][n 64][8]: half h holds output columns h*64 .. h*64+63
    for (int h = 0; h < 2; h++)
        for (int kg = 0; kg < 8; kg++)
            for (int n = 0; n < 64; n++)
                for (int e = 0; e < 8; e++) hBp[((h * 8 + kg) * 64 + n) * 8 + e] = __float2half(fB[(h * 64 + n) * K + kg * 8 + e]);
    __half *dA, *dB; float *dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hBp.size() * 2); cudaMalloc(&dD, 256 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hBp.data(), hBp.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, 256 * N * 4);
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
    if (ge != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled: %s\n", cudaGetErrorString(ge)); return 1; }
    CUtensorMap map;
    const cuuint64_t gdim[2] = {256, 64}, gstride[1] = {256};
    const cuuint32_t box[2] = {256, 32}, estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("cuTensorMapEncodeTiled -> %d\n", (int)cr);
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCheckSmem);
    check_kernel<<<2, 128, kCheckSmem>>>(dA, map, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("check kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> hD(256 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_swapped = 0;
    for (int m = 0; m < 256; m++)
        for (int n = 0; n < N; n++) {
            float ref = 0, ref_sw = 0;
            for (int k = 0; k < K; k++) { ref += fA[m * K + k] * fB[n * K + k]; ref_sw += fA[m * K + k] * fB[(n ^ 64) * K + k]; }
            bad += hD[m * N + n] != ref;
            bad_swapped += hD[m * N + n] != ref_sw;
        }
    printf("cta_group::2 check: %d of %d outputs differ from the host (with the B halves swapped: %d)  D[0][0..3] = %g %g %g %g, D[128][64] = %g\n", bad, 256 * N,
           bad_swapped, hD[0], hD[1], hD[2], hD[3], hD[128 * N + 64]);

    // ---------------- part 2: rates
    uint8_t *blob; long long *d_out, h[2];
    cudaMalloc(&blob, 100 * 32768 + 32768); cudaMemset(blob, 0, 100 * 32768 + 32768); cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRateSmem);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRateSmem);
    const int units = 2048;
    for (int grid : {2, 148})
        for (int cg = 1; cg <= 2; cg++)
            for (int ws = 0; ws <= 1; ws++) {
                for (int rep = 0; rep < 2; rep++) {
                    if (cg == 1) rate_kernel<1><<<grid, 128, kRateSmem>>>(blob, units, ws, d_out);
                    else {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = kRateSmem;
                        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
                        cfg.attrs = &at; cfg.numAttrs = 1;
                        cudaLaunchKernelEx(&cfg, rate_kernel<2>, (const uint8_t *)blob, units, ws, d_out);
                    }
                }
                e = cudaDeviceSynchronize();
                cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
                printf("grid %3d cta_group::%d %s: %7.1f cycles per p2 unit issued, %7.1f to completion  %s\n", grid, cg, ws ? "with weight stream" : "MMAs alone        ",
                       (double)h[0] / units, (double)h[1] / units, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
