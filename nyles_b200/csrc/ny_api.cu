// Context management and error reporting of libnyles_b200.so.
#include "ny_common.cuh"

static thread_local char g_err[512] = "";

void ny_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ny_last_error(void) { return g_err; }
extern "C" int ny_version(void) { return 100; }

extern "C" int ny_init(int device, ny_ctx** out)
{
    if (!out) { ny_set_error("ny_init: null out pointer"); return NY_ERR_ARG; }
    int count = 0;
    NY_CUDA(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) {
        ny_set_error("ny_init: device %d not present (%d CUDA devices visible)", device, count);
        return NY_ERR_ARG;
    }
    NY_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    NY_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        ny_set_error("ny_init: libnyles_b200 carries sm_100a code only; device %d is sm_%d%d",
                     device, prop.major, prop.minor);
        return NY_ERR_CUDA;
    }
    ny_ctx* ctx = new ny_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->launches = 0;
    ctx->scratch_doubles = 1 << 16;
    ctx->d_scratch = nullptr;
    ctx->h_pinned = nullptr;
    if (cudaMalloc(&ctx->d_scratch, ctx->scratch_doubles * sizeof(double)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_pinned, 64 * sizeof(double)) != cudaSuccess) {
        ny_set_error("ny_init: scratch allocation failed");
        ny_free(ctx);
        return NY_ERR_CUDA;
    }
    *out = ctx;
    return NY_OK;
}

extern "C" void ny_free(ny_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    delete ctx;
}

extern "C" long long ny_launch_count(ny_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void ny_launch_count_reset(ny_ctx* ctx) { if (ctx) ctx->launches = 0; }
