// Micro-benchmark: how fast can one SM stream an L2-resident blob into a shared-memory ring with cp.async.bulk (1-D TMA)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
// One producer thread per CTA issues the copies, one consumer thread frees a stage as soon as it is full (no math): the figure is
// the ceiling of the weight stream of trunk.cu for a given ring shape.
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
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t *blob, size_t blob_bytes, int stages, int stage_bytes, int rounds, unsigned long long *sink, int producers) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_full = sbase + stages * stage_bytes, bar_empty = bar_full + 8 * stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long units = (long long)(blob_bytes / stage_bytes) * rounds;
    if (threadIdx.x == 0 || (threadIdx.x == 64 && producers == 2)) {
        const int me = threadIdx.x == 0 ? 0 : 1;
        for (long long u = me; u < units; u += producers) {
            const uint32_t stage = (uint32_t)(u % stages), phase = (uint32_t)((u / stages) & 1);
            if (SPIN) mbar_spin(bar_empty + 8 * stage, phase ^ 1); else mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            mbar_expect_tx(bar_full + 8 * stage, stage_bytes);
            bulk_g2s(sbase + stage * stage_bytes, blob + (size_t)(u % (blob_bytes / stage_bytes)) * stage_bytes, stage_bytes, bar_full + 8 * stage);
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

int main() {
    const size_t blob_bytes = 7 << 20;
    uint8_t *blob; unsigned long long *sink;
    cudaMalloc(&blob, blob_bytes); cudaMemset(blob, 1, blob_bytes); cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int shapes[][2] = {{3, 32768}, {6, 16384}, {12, 8192}, {3, 16384}, {2, 32768}, {6, 32768}, {4, 49152}, {24, 4096}};
    for (int prod = 1; prod <= 2; prod++)
    for (int spin = 0; spin < 1; spin++)
    for (int grid : {1}) {
        for (auto &sh : shapes) {
            const int stages = sh[0], sb = sh[1], rounds = grid == 1 ? 8 : 8;
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
    return 0;
}
