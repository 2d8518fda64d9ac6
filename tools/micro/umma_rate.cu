// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, N = 128, K = 16, fp32 accumulate, operands in
// shared memory, no-swizzle K-major) for the issue patterns of trunk.cu.  One thread issues, one accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../iago_b200/csrc/tc.cuh"
using namespace iago;

constexpr int kGroup = 3200, kRowPitch = 160;   // the trunk's activation tile geometry
constexpr int kARegion = 16 * kGroup * 2, kBRegion = 65536;

// pattern 0: the same A / B slice every time            1: K advances every MMA (precision 1)
//         2: three MMAs per K step (precision 3)        3: K advances, all-new A and B regions each MMA (far apart)
//         4: like 1 but two accumulators alternating     5: like 2 but N = 256 for the first MMA of a step (2 MMAs per step)
__global__ void __launch_bounds__(64, 1) rate_kernel(int pattern, int steps, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sbase = smem_u32(smem), bar_a = smem_u32(&bar);
    for (int i = threadIdx.x; i < (kARegion + kBRegion) / 16; i += 64) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(bar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint64_t hi_a = ((uint64_t)(kRowPitch >> 4) << 32) | (1ULL << 46), hi_b = ((uint64_t)(128 >> 4) << 32) | (1ULL << 46);
        const uint32_t a_lbo = (uint32_t)(kGroup >> 4) << 16, b_lbo = (uint32_t)(2048 >> 4) << 16;
        const uint32_t idesc = instr_desc(128, 128), idesc256 = instr_desc(128, 256);
        const uint32_t a0 = sbase, a1 = sbase + 16 * kGroup, b0 = sbase + kARegion, b1 = b0 + 16384;
        long long t0 = clock64();
        long long n_mma = 0;
        if (pattern >= 6) {
            const uint64_t da0 = hi_a | ((a0 >> 4) | a_lbo), db0 = hi_b | ((b0 >> 4) | b_lbo);
            uint64_t da[4], db[4], dal[4], dbl[4];
            for (int ks = 0; ks < 4; ks++) {
                da[ks] = hi_a | (((a0 + ks * 2 * kGroup) >> 4) | a_lbo); dal[ks] = hi_a | (((a1 + ks * 2 * kGroup) >> 4) | a_lbo);
                db[ks] = hi_b | (((b0 + ks * 4096) >> 4) | b_lbo); dbl[ks] = hi_b | (((b1 + ks * 4096) >> 4) | b_lbo);
            }
            t0 = clock64();
            if (pattern == 6) {          // one MMA per trip of a tight loop, fixed descriptors
#pragma unroll 1
                for (int s = 0; s < steps; s++) umma_f16(tmem, da0, db0, idesc, 1);
                n_mma = steps;
            } else if (pattern == 7) {   // twelve MMAs straight per trip, fixed descriptors
#pragma unroll 1
                for (int s = 0; s < steps; s++) {
#pragma unroll
                    for (int j = 0; j < 12; j++) umma_f16(tmem, da0, db0, idesc, 1);
                }
                n_mma = 12LL * steps;
            } else if (pattern == 8) {   // the trunk's unit: 4 K steps x 3 MMAs, descriptors precomputed in registers
#pragma unroll 1
                for (int s = 0; s < steps; s++) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        umma_f16(tmem, da[ks], db[ks], idesc, 1);
                        umma_f16(tmem, da[ks], dbl[ks], idesc, 1);
                        umma_f16(tmem, dal[ks], db[ks], idesc, 1);
                    }
                }
                n_mma = 12LL * steps;
            } else if (pattern == 10) {  // FP8 (kind::f8f6f4, K = 32) MMAs straight, fixed descriptors
#pragma unroll 1
                for (int s = 0; s < steps; s++) {
#pragma unroll
                    for (int j = 0; j < 12; j++) umma_f8(tmem, da0, db0, idesc, 1);
                }
                n_mma = 12LL * steps;
            } else if (pattern == 11) {  // precision 2's unit: 4 fp16 MMAs then 4 FP8 MMAs into a second accumulator
#pragma unroll 1
                for (int s = 0; s < steps; s++) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) umma_f16(tmem, da[ks], db[ks], idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) {
                        umma_f8(tmem + 256, dal[ks], dbl[ks], idesc, 1);
                        umma_f8(tmem + 256, da[ks], dbl[ks + 2], idesc, 1);
                    }
                }
                n_mma = 8LL * steps;
            } else if (pattern == 12) {  // the same 8 MMAs, all kind::f16 (what a kind switch costs)
#pragma unroll 1
                for (int s = 0; s < steps; s++) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) umma_f16(tmem, da[ks], db[ks], idesc, 1);
#pragma unroll
                    for (int ks = 0; ks < 2; ks++) {
                        umma_f16(tmem + 256, dal[ks], dbl[ks], idesc, 1);
                        umma_f16(tmem + 256, da[ks], dbl[ks + 2], idesc, 1);
                    }
                }
                n_mma = 8LL * steps;
            } else {                     // precision 1's unit: 4 K steps x 1 MMA, precomputed
#pragma unroll 1
                for (int s = 0; s < steps; s++) {
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) umma_f16(tmem, da[ks], db[ks], idesc, 1);
                }
                n_mma = 4LL * steps;
            }
        } else
        for (int s = 0; s < steps; s++) {
            const int ks = s & 3;
            const uint32_t aw = ((a0 + ks * 2 * kGroup) >> 4) | a_lbo, alw = ((a1 + ks * 2 * kGroup) >> 4) | a_lbo;
            const uint32_t bw = ((b0 + ks * 4096) >> 4) | b_lbo, blw = ((b1 + ks * 4096) >> 4) | b_lbo;
            if (pattern == 0) { umma_f16(tmem, hi_a | ((a0 >> 4) | a_lbo), hi_b | ((b0 >> 4) | b_lbo), idesc, 1); n_mma++; }
            else if (pattern == 1) { umma_f16(tmem, hi_a | aw, hi_b | bw, idesc, 1); n_mma++; }
            else if (pattern == 2) {
                umma_f16(tmem, hi_a | aw, hi_b | bw, idesc, 1);
                umma_f16(tmem, hi_a | aw, hi_b | blw, idesc, 1);
                umma_f16(tmem, hi_a | alw, hi_b | bw, idesc, 1);
                n_mma += 3;
            } else if (pattern == 3) {
                const uint32_t off = (uint32_t)(s % 7) * 7 * 2048;   // scattered over the regions
                umma_f16(tmem, hi_a | (((a0 + (off % (kARegion - 8 * kGroup)) / 16 * 16) >> 4) | a_lbo), hi_b | (((b0 + off % 49152) >> 4) | b_lbo), idesc, 1);
                n_mma++;
            } else if (pattern == 4) { umma_f16(tmem + (s & 1) * 128, hi_a | aw, hi_b | bw, idesc, 1); n_mma++; }
            else {
                umma_f16(tmem, hi_a | aw, hi_b | (((b0 + ks * 8192) >> 4) | ((uint32_t)(4096 >> 4) << 16)), idesc256, 1);   // B = [hi | lo] rows, LBO 4,096
                umma_f16(tmem, hi_a | alw, hi_b | (((b0 + ks * 8192) >> 4) | ((uint32_t)(4096 >> 4) << 16)), idesc, 1);
                n_mma += 2;
            }
        }
        const long long t1 = clock64();
        umma_commit(bar_a);
        mbar_wait(bar_a, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0; out[2] = n_mma;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

int main() {
    long long *d_out, h[3];
    cudaMalloc(&d_out, 24);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kARegion + kBRegion);
    const char *names[] = {"same A/B slice every MMA", "K advances every MMA (p1)", "3 MMAs per K step (p3)", "scattered slices", "K advances, 2 accumulators", "N=256 + N=128 per K step", "tight loop, 1 MMA per trip", "12 MMAs straight, fixed desc", "unit of 4x3 MMAs, precomputed", "unit of 4x1 MMAs, precomputed", "12 FP8 K=32 MMAs straight", "p2 unit: 4 f16 + 4 f8", "same 8 MMAs all f16"};
    for (int grid : {1})
        for (int p = 6; p < 13; p++) {
            rate_kernel<<<grid, 64, kARegion + kBRegion>>>(p, 64, d_out);
            rate_kernel<<<grid, 64, kARegion + kBRegion>>>(p, 4096, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
            printf("grid %3d  %-30s: %6.1f cycles per MMA issued, %6.1f to completion (%lld MMAs)  %s\n", grid, names[p], (double)h[0] / h[2], (double)h[1] / h[2], h[2],
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
