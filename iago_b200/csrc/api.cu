// api.cu — context lifetime, error reporting, staging buffers (C ABI of include/iago_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace iago {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int ensure_staging(iago_ctx *ctx, size_t bytes) {
    if (ctx->stage.bytes >= bytes) return IAGO_OK;
    size_t want = bytes + bytes / 4 + 4096;
    if (ctx->stage.host) cudaFreeHost(ctx->stage.host);
    if (ctx->stage.dev) cudaFree(ctx->stage.dev);
    ctx->stage = Staging();
    IAGO_CUDA(cudaMallocHost(&ctx->stage.host, want));
    IAGO_CUDA(cudaMalloc(&ctx->stage.dev, want));
    ctx->stage.bytes = want;
    return IAGO_OK;
}

void trunk_destroy(iago_ctx *ctx);
void selfplay_destroy(iago_ctx *ctx);
void valuegen_destroy(iago_ctx *ctx);

}  // namespace iago

extern "C" {

int iago_abi_version(void) { return IAGO_ABI_VERSION; }

const char *iago_last_error(void) { return iago::g_err; }

int iago_ctx_create(int device, iago_ctx **out) {
    IAGO_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    int count = 0;
    IAGO_CUDA(cudaGetDeviceCount(&count));
    if (count <= 0) {
        iago::set_error("no CUDA device visible; libiago_b200 has no CPU fallback");
        return IAGO_E_CUDA;
    }
    IAGO_REQUIRE(device >= 0 && device < count, "device index out of range");
    cudaDeviceProp prop;
    IAGO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        iago::set_error("device %d is sm_%d%d; libiago_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return IAGO_E_CUDA;
    }
    iago::DeviceGuard guard(device);
    iago_ctx *ctx = new iago_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    IAGO_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    IAGO_CUDA(cudaEventCreate(&ctx->ev0));
    IAGO_CUDA(cudaEventCreate(&ctx->ev1));
    IAGO_CUDA(cudaMalloc(&ctx->d_rollout, sizeof(iago::RolloutWeights)));
    IAGO_CUDA(cudaMalloc(&ctx->d_counters, 8 * sizeof(uint64_t)));
    IAGO_CUDA(cudaMemset(ctx->d_counters, 0, 8 * sizeof(uint64_t)));
    *out = ctx;
    return IAGO_OK;
}

int iago_ctx_destroy(iago_ctx *ctx) {
    if (!ctx) return IAGO_OK;
    iago::DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    iago::trunk_destroy(ctx);
    iago::selfplay_destroy(ctx);
    iago::valuegen_destroy(ctx);
    if (ctx->stage.host) cudaFreeHost(ctx->stage.host);
    if (ctx->stage.dev) cudaFree(ctx->stage.dev);
    cudaFree(ctx->d_rollout);
    cudaFree(ctx->d_counters);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    for (int c = 1; c < 4; c++)
        if (ctx->host_streams[c]) cudaStreamDestroy(ctx->host_streams[c]);
    for (iago::HostLane &l : ctx->lanes) {
        if (l.stream) {
            cudaStreamSynchronize(l.stream);
            cudaStreamDestroy(l.stream);
        }
        if (l.dev) cudaFree(l.dev);
        if (l.d_cnt) cudaFree(l.d_cnt);
        if (l.h_cnt) cudaFreeHost(l.h_cnt);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return IAGO_OK;
}

int iago_ctx_sync(iago_ctx *ctx) {
    IAGO_REQUIRE(ctx != nullptr, "ctx is NULL");
    iago::DeviceGuard guard(ctx->device);
    IAGO_CUDA(cudaStreamSynchronize(ctx->stream));
    return IAGO_OK;
}

void *iago_ctx_stream(iago_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int iago_last_kernel_ms(iago_ctx *ctx, float *ms) {
    IAGO_REQUIRE(ctx != nullptr && ms != nullptr, "NULL argument");
    if (!ctx->timed) {
        iago::set_error("no timed launch recorded yet");
        return IAGO_E_STATE;
    }
    iago::DeviceGuard guard(ctx->device);
    IAGO_CUDA(cudaEventSynchronize(ctx->ev1));
    IAGO_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return IAGO_OK;
}

}  // extern "C"
