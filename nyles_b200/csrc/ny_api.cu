// Context management and error reporting of libnyles_b200.so.
#include "ny_common.cuh"
#include <cstdlib>

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
    ctx->fast_arith = 0;
    ctx->mom_variant = 0;
    { const char* v = getenv("NY_MOM_VARIANT"); if (v && *v >= '0' && *v <= '2') ctx->mom_variant = *v - '0'; }
    ctx->scratch_doubles = 1 << 16;
    ctx->d_scratch = nullptr;
    ctx->h_pinned = nullptr;
    ctx->prof_mask = 0;
    for (int t = 0; t < NY_PROF_NTAGS; t++) { ctx->prof_ms[t] = 0.0; ctx->prof_n[t] = 0; }
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
    for (ny_prof_rec& r : ctx->prof_recs) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    delete ctx;
}

extern "C" long long ny_launch_count(ny_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" void ny_launch_count_reset(ny_ctx* ctx) { if (ctx) ctx->launches = 0; }

// ---- event profiling of kernel families ----------------------------------------------------
static const char* const g_prof_names[NY_PROF_NTAGS] = {
    "rhs_tracer", "rhs_momentum", "vorticity_ke", "div", "gradp", "U_from_u", "timescheme", "maxspeed",
    "halo", "mg_smooth_fine", "mg_residual_fine", "mg_restrict_fine", "mg_prolong_fine", "mg_norm",
    "mg_coarse_levels", "mg_embed_extract", "mg_down_fine", "mg_up_fine"};

extern "C" const char* ny_prof_name(int tag)
{
    return (tag >= 0 && tag < NY_PROF_NTAGS) ? g_prof_names[tag] : "";
}

extern "C" int ny_prof_collect(ny_ctx* ctx, double* ms_host, long long* n_host)
{
    NY_REQUIRE(ctx, "null argument");
    NY_CUDA(cudaDeviceSynchronize());
    for (ny_prof_rec& r : ctx->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.t0, r.t1) == cudaSuccess) {
            ctx->prof_ms[r.tag] += ms;
            ctx->prof_n[r.tag] += 1;
        }
        ctx->prof_pool.push_back(r.t0);
        ctx->prof_pool.push_back(r.t1);
    }
    ctx->prof_recs.clear();
    for (int t = 0; t < NY_PROF_NTAGS; t++) {
        if (ms_host) ms_host[t] = ctx->prof_ms[t];
        if (n_host) n_host[t] = ctx->prof_n[t];
    }
    return NY_OK;
}

extern "C" int ny_prof_start(ny_ctx* ctx, unsigned long long mask)
{
    NY_REQUIRE(ctx, "null argument");
    int r = ny_prof_collect(ctx, nullptr, nullptr);
    if (r != NY_OK) return r;
    for (int t = 0; t < NY_PROF_NTAGS; t++) { ctx->prof_ms[t] = 0.0; ctx->prof_n[t] = 0; }
    ctx->prof_mask = mask;
    return NY_OK;
}

// ---- TMA tensor maps -------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is resolved through the runtime so that the
// library needs no link-time libcuda (the CPU build box has none) and loads without a GPU.
#include "ny_tma.cuh"

typedef CUresult (*ny_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ny_tma_encode_3d(CUtensorMap* map, const double* base, int nx, int ny, int nz, int bx, int by, int bz)
{
    static ny_encode_tiled_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
            ny_set_error("cuTensorMapEncodeTiled is not available from this driver (%s)", cudaGetErrorString(e));
            return NY_ERR_CUDA;
        }
        encode = reinterpret_cast<ny_encode_tiled_fn>(fn);
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (nx & 1) || (bx & 1) || bx > 256 || by > 256 || bz > 256 ||
        bx < 1 || by < 1 || bz < 1) {
        ny_set_error("ny_tma_encode_3d: array %p (%d,%d,%d) box (%d,%d,%d) violates the TMA alignment rules",
                     (const void*)base, nz, ny, nx, bz, by, bx);
        return NY_ERR_ARG;
    }
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    const cuuint64_t strides[2] = {(cuuint64_t)nx * 8, (cuuint64_t)nx * (cuuint64_t)ny * 8};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ny_set_error("cuTensorMapEncodeTiled failed with CUresult %d for array (%d,%d,%d) box (%d,%d,%d)", (int)r, nz, ny,
                     nx, bz, by, bx);
        return NY_ERR_CUDA;
    }
    return NY_OK;
}
