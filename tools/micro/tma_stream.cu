// Micro-benchmark: how fast can one SM stream an L2-resident blob into a shared-memory ring with cp.async.bulk (1-D TMA)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
// One producer thread per CTA issues the copies, one consumer thread frees a stage as soon as it is full (no math): the figure is
// the ceiling of the weight stream of trunk.cu for a given ring shape.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../iago_b200/csrc/tc.cuh"
using namespace iago;

__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {   // busy test_wait: no hardware suspend
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
template <bool SPIN>
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t *blob, size_t blob_bytes, int stages, int stage_bytes, int rounds, unsigned long long *sink, int producers, int split = 1) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + stages * stage_bytes, bar_empty = bar_full + 8 * stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long units = (long long)(blob_bytes / stage_bytes) * rounds;
    if ((threadIdx.x & 31) == 0 && threadIdx.x != 32 && (threadIdx.x == 0 ? 0 : (threadIdx.x >> 5) - 1) < producers) {
        const int me = threadIdx.x == 0 ? 0 : (threadIdx.x >> 5) - 1;   // producers on warps 0, 2, 3 (warp 1 = consumer)
        for (long long u = me; u < units; u += producers) {
            const uint32_t stage = (uint32_t)(u % stages), phase = (uint32_t)((u / stages) & 1);
            if (SPIN) mbar_spin(bar_empty + 8 * stage, phase ^ 1); else mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_expect_tx(bar_full + 8 * stage, stage_bytes);
            const int part = stage_bytes / split;
            for (int q = 0; q < split; q++)
                bulk_g2s(sbase + stage * stage_bytes + q * part, blob + (size_t)(u % (blob_bytes / stage_bytes)) * stage_bytes + q * part, part, bar_full + 8 * stage);
        }
    } else if (threadIdx.x == 32) {
        uint32_t stage = 0, phase = 0;
        unsigned long long acc = 0;
        for (long long u = 0; u < units; u++) {
            if (SPIN) mbar_spin(bar_full + 8 * stage, phase); else mbar_wait(bar_full + 8 * stage, phase);
            acc += smem[stage * stage_bytes];
            mbar_arrive(bar_empty + 8 * stage);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x123456789ULL) *sink = acc;
    }
}

// Variant T: the same ring fed by 2-D tensor-map TMA (cp.async.bulk.tensor.2d, box = 256 B x stage_bytes/256 rows).
__global__ void __launch_bounds__(128, 1) stream_tensor_kernel(const __grid_constant__ CUtensorMap map, size_t blob_bytes, int stages, int stage_bytes, int rounds, unsigned long long *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + stages * stage_bytes, bar_empty = bar_full + 8 * stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long per = (long long)(blob_bytes / stage_bytes), units = per * rounds;
    const int rows = stage_bytes / 256;
    if (threadIdx.x == 0) {
        for (long long u = 0; u < units; u++) {
            const uint32_t stage = (uint32_t)(u % stages), phase = (uint32_t)((u / stages) & 1);
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_expect_tx(bar_full + 8 * stage, stage_bytes);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(sbase + stage * stage_bytes), "l"(&map), "r"(0), "r"((int)(u % per) * rows), "r"(bar_full + 8 * stage) : "memory");
        }
    } else if (threadIdx.x == 32) {
        uint32_t stage = 0, phase = 0;
        unsigned long long acc = 0;
        for (long long u = 0; u < units; u++) {
            mbar_wait(bar_full + 8 * stage, phase);
            acc += smem[stage * stage_bytes];
            mbar_arrive(bar_empty + 8 * stage);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x123456789ULL) *sink = acc;
    }
}

// Variant L: the ring fed by cp.async (LDGSTS, 16 B per thread) from two producer warps; cp.async.mbarrier.arrive.noinc signals "full".
__global__ void __launch_bounds__(128, 1) stream_ldgsts_kernel(const uint8_t *blob, size_t blob_bytes, int stages, int stage_bytes, int rounds, unsigned long long *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + stages * stage_bytes, bar_empty = bar_full + 8 * stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(bar_full + 8 * s, 64); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long per = (long long)(blob_bytes / stage_bytes), units = per * rounds;
    if (threadIdx.x < 64) {
        for (long long u = 0; u < units; u++) {
            const uint32_t stage = (uint32_t)(u % stages), phase = (uint32_t)((u / stages) & 1);
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint8_t *src = blob + (size_t)(u % per) * stage_bytes;
            const uint32_t dst = sbase + stage * stage_bytes;
            for (int o = threadIdx.x * 16; o < stage_bytes; o += 64 * 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + o), "l"(src + o) : "memory");
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_full + 8 * stage) : "memory");
        }
    } else if (threadIdx.x == 96) {
        uint32_t stage = 0, phase = 0;
        unsigned long long acc = 0;
        for (long long u = 0; u < units; u++) {
            mbar_wait(bar_full + 8 * stage, phase);
            acc += smem[stage * stage_bytes];
            mbar_arrive(bar_empty + 8 * stage);
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x123456789ULL) *sink = acc;
    }
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const size_t blob_bytes = 7 << 20;
    uint8_t *blob; unsigned long long *sink;
    cudaMalloc(&blob, blob_bytes); cudaMemset(blob, 1, blob_bytes); cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int shapes[][2] = {{3, 32768}, {6, 16384}, {12, 8192}, {3, 16384}, {2, 32768}, {6, 32768}, {4, 49152}, {24, 4096}};
    for (int prod = 1; prod <= 3; prod++)
    for (int spin = 0; spin < 1; spin++)
    for (int grid : {148}) {
        for (auto &sh : shapes) {
            const int stages = sh[0], sb = sh[1], rounds = grid == 1 ? 8 : 8;
            if (stages < prod) continue;
            const size_t smem = (size_t)stages * sb + 16 * stages + 64;
            if (spin) stream_kernel<true><<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, 1, sink, prod); else stream_kernel<false><<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, 1, sink, prod);
            cudaEventRecord(a);
            if (spin) stream_kernel<true><<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, rounds, sink, prod); else stream_kernel<false><<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, rounds, sink, prod);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            const double per_sm = (double)blob_bytes * rounds / (ms * 1e-3) / 1e9;
            printf("producers %d %s grid %3d  ring %2d x %5d B (%3d KB in flight): %7.3f ms  %6.1f GB/s per SM  %7.2f TB/s total  %s\n", prod, spin ? "spin" : "wait", grid, stages, sb, stages * sb / 1024, ms, per_sm,
                   per_sm * grid / 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    // tensor-map TMA and LDGSTS variants of the same rings
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres);
    cudaFuncSetAttribute(stream_tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(stream_ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int variant = 0; variant < 2; variant++)
    for (int grid : {1, 148}) {
        for (auto &sh : shapes) {
            const int stages = sh[0], sb = sh[1], rounds = 8;
            const size_t smem = (size_t)stages * sb + 16 * stages + 64;
            CUtensorMap map;
            const cuuint64_t gdim[2] = {256, blob_bytes / 256}, gstride[1] = {256};
            const cuuint32_t box[2] = {256, (cuuint32_t)(sb / 256)}, estr[2] = {1, 1};
            if (variant == 0 && encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, blob, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
            for (int rep = 0; rep < 2; rep++) {
                if (rep == 1) cudaEventRecord(a);
                if (variant == 0) stream_tensor_kernel<<<grid, 128, smem>>>(map, blob_bytes, stages, sb, rep ? rounds : 1, sink);
                else stream_ldgsts_kernel<<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, rep ? rounds : 1, sink);
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            const double per_sm = (double)blob_bytes * rounds / (ms * 1e-3) / 1e9;
            printf("%s grid %3d  ring %2d x %5d B (%3d KB in flight): %7.3f ms  %6.1f GB/s per SM  %7.2f TB/s total  %s\n", variant == 0 ? "tensor-map TMA" : "cp.async 16 B  ", grid, stages, sb,
                   stages * sb / 1024, ms, per_sm, per_sm * grid / 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    // (a) does the fixed time per copy hold for larger copies?  (b) a unit issued as 2 / 4 / 8 copies under one expect_tx
    const int shapes2[][3] = {{2, 65536, 1}, {2, 98304, 1}, {3, 32768, 2}, {3, 32768, 4}, {3, 32768, 8}, {6, 32768, 8}, {2, 98304, 6}, {3, 65536, 16}};
    for (int grid : {1, 148})
        for (auto &sh : shapes2) {
            const int stages = sh[0], sb = sh[1], split = sh[2], rounds = 8;
            const size_t smem = (size_t)stages * sb + 16 * stages + 64;
            cudaFuncSetAttribute(stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            stream_kernel<false><<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, 1, sink, 1, split);
            cudaEventRecord(a);
            stream_kernel<false><<<grid, 128, smem>>>(blob, blob_bytes, stages, sb, rounds, sink, 1, split);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            const double per_sm = (double)(blob_bytes / sb) * sb * rounds / (ms * 1e-3) / 1e9;
            printf("bulk, unit = %d copies  grid %3d  ring %2d x %5d B: %7.3f ms  %6.1f GB/s per SM  %7.2f TB/s total  %6.3f us per unit  %s\n", split, grid, stages, sb, ms, per_sm,
                   per_sm * grid / 1e3, ms * 1e3 / ((double)(blob_bytes / sb) * rounds), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
