// Slab communicator of libnyles_b200.so: NCCL over NVLink, one rank per GPU, ranks ordered along z.
//
// Replaces the mpi4py / mpi_f08 plumbing of the reference for a z-slab decomposition:
//   core/mpi/halo.py:140-178 (persistent-request halo fill)      -> ny_halo_exchange
//   core/mpi/mpitools.py:28-32, mgfor/operators.f90:122 (allreduce) -> comm_allreduce
//   mgfor/mod_gluesplit.f90:142-207 (Allgather of coarse tiles)  -> comm_allgather
// NCCL is resolved with dlsym at first use: a process that already loaded libnccl (PyTorch) shares
// that copy; a single-GPU user never needs it.
#include "ny_common.cuh"
#include "ny_comm.cuh"
#include <dlfcn.h>

namespace {

struct NcclApi {
    bool ok = false;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
} g_nccl;

bool load_nccl()
{
    if (g_nccl.ok) return true;
    void* h = RTLD_DEFAULT;
    if (!dlsym(h, "ncclCommInitRank")) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { ny_set_error("NCCL is not loaded in this process and libnccl.so.2 cannot be opened: %s", dlerror()); return false; }
    }
#define NY_SYM(name)                                                                     \
    *(void**)(&g_nccl.name) = dlsym(h, "nccl" #name);                                    \
    if (!g_nccl.name) { ny_set_error("libnccl lacks nccl" #name); return false; }
    NY_SYM(GetUniqueId) NY_SYM(CommInitRank) NY_SYM(CommDestroy) NY_SYM(AllReduce) NY_SYM(AllGather)
    NY_SYM(Send) NY_SYM(Recv) NY_SYM(GroupStart) NY_SYM(GroupEnd) NY_SYM(GetErrorString)
#undef NY_SYM
    g_nccl.ok = true;
    return true;
}

#define NY_NCCL(call)                                                                    \
    do {                                                                                 \
        ncclResult_t _r = (call);                                                        \
        if (_r != ncclSuccess) {                                                         \
            ny_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r)); \
            return NY_ERR_COMM;                                                          \
        }                                                                                \
    } while (0)

// periodic wrap of the x and / or y halos of a model array over ALL planes (halo planes included,
// which reproduces the edge and corner boxes of the 26-neighbour exchange, core/mpi/halo.py:93-120)
__global__ void __launch_bounds__(256)
k_wrap_xy(double* __restrict__ f, int nz, int ny, int nx, int nh, int yper, int xper)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= nx || j >= ny || k >= nz) return;
    int si = i, sj = j;
    bool halo = false;
    if (xper) {
        int n = nx - 2 * nh;
        if (i < nh) { si = i + n; halo = true; } else if (i >= nh + n) { si = i - n; halo = true; }
    }
    if (yper) {
        int n = ny - 2 * nh;
        if (j < nh) { sj = j + n; halo = true; } else if (j >= nh + n) { sj = j - n; halo = true; }
    }
    if (!halo) return;
    long long plane = (long long)k * ny * nx;
    f[plane + (long long)j * nx + i] = f[plane + (long long)sj * nx + si];
}

}  // namespace

extern "C" int ny_comm_unique_id(char* id128)
{
    NY_REQUIRE(id128, "null argument");
    if (!load_nccl()) return NY_ERR_COMM;
    ncclUniqueId id;
    NY_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, NCCL_UNIQUE_ID_BYTES);
    return NY_OK;
}

extern "C" int ny_comm_init(ny_ctx* ctx, int nranks, int rank, const char* id128, ny_comm** out)
{
    NY_REQUIRE(ctx && out && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "bad argument");
    if (!load_nccl()) return NY_ERR_COMM;
    NY_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(id.internal, id128, NCCL_UNIQUE_ID_BYTES);
    ny_comm* c = new ny_comm();
    c->ctx = ctx; c->nranks = nranks; c->rank = rank; c->nccl = nullptr; c->d_red = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
    if (r != ncclSuccess) {
        ny_set_error("ncclCommInitRank(%d of %d) -> %s", rank, nranks, g_nccl.GetErrorString(r));
        delete c;
        return NY_ERR_COMM;
    }
    c->xstream = nullptr; c->ev_ready = nullptr; c->ev_done = nullptr;
    int lo_prio = 0, hi_prio = 0;
    cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio);
    if (cudaStreamCreateWithPriority(&c->xstream, cudaStreamNonBlocking, hi_prio) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaMalloc(&c->d_red, 8 * sizeof(double)) != cudaSuccess) {
        ny_set_error("ny_comm_init: cudaMalloc failed");
        g_nccl.CommDestroy(c->nccl);
        delete c;
        return NY_ERR_CUDA;
    }
    *out = c;
    return NY_OK;
}

extern "C" void ny_comm_free(ny_comm* c)
{
    if (!c) return;
    if (c->nccl && g_nccl.ok) g_nccl.CommDestroy(c->nccl);
    if (c->d_red) cudaFree(c->d_red);
    if (c->xstream) cudaStreamDestroy(c->xstream);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    delete c;
}

extern "C" int ny_comm_size(ny_comm* c) { return c ? c->nranks : 1; }
extern "C" int ny_comm_rank(ny_comm* c) { return c ? c->rank : 0; }

// ---- internal helpers used by the multigrid --------------------------------------------------
int ny_comm_allreduce(ny_comm* c, double* d_buf, int n, int op_max, cudaStream_t st)
{
    if (!c || c->nranks == 1) return NY_OK;
    NY_NCCL(g_nccl.AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, op_max ? ncclMax : ncclSum, c->nccl, st));
    return NY_OK;
}

int ny_comm_allgather_inplace(ny_comm* c, double* d_recv, size_t count_per_rank, cudaStream_t st)
{
    if (!c || c->nranks == 1) return NY_OK;
    NY_NCCL(g_nccl.AllGather(d_recv + (size_t)c->rank * count_per_rank, d_recv, count_per_rank, ncclDouble, c->nccl, st));
    return NY_OK;
}

// Exchange the nh z-faces of nf slab arrays with the ranks below / above (-1: none).
// Array a: planes of `plane` doubles; lo = index of the first interior plane, nint = interior planes.
int ny_comm_exchange_z(ny_comm* c, double* const* arrays, int nf, size_t plane, int lo, int nint, int nh,
                       int below, int above, cudaStream_t st)
{
    if (!c || (below < 0 && above < 0)) return NY_OK;
    const size_t cnt = (size_t)nh * plane;
    NY_NCCL(g_nccl.GroupStart());
    // sends first, then receives in the opposite neighbour order: when below == above (two ranks,
    // periodic z) NCCL matches per peer in posting order, and my low face is the peer's HIGH halo.
    for (int f = 0; f < nf; f++) {
        double* a = arrays[f];
        if (below >= 0) NY_NCCL(g_nccl.Send(a + (size_t)lo * plane, cnt, ncclDouble, below, c->nccl, st));
        if (above >= 0) NY_NCCL(g_nccl.Send(a + (size_t)(lo + nint - nh) * plane, cnt, ncclDouble, above, c->nccl, st));
        if (above >= 0) NY_NCCL(g_nccl.Recv(a + (size_t)(lo + nint) * plane, cnt, ncclDouble, above, c->nccl, st));
        if (below >= 0) NY_NCCL(g_nccl.Recv(a + (size_t)(lo - nh) * plane, cnt, ncclDouble, below, c->nccl, st));
    }
    NY_NCCL(g_nccl.GroupEnd());
    return NY_OK;
}

int ny_comm_exchange_z2(ny_comm* c, double* a0, size_t plane0, int nint0, double* a1, size_t plane1, int nint1, int nh,
                        int below, int above, cudaStream_t st)
{
    if (!c || (below < 0 && above < 0)) return NY_OK;
    double* arr[2] = {a0, a1};
    const size_t plane[2] = {plane0, plane1};
    const int nint[2] = {nint0, nint1};
    NY_NCCL(g_nccl.GroupStart());
    for (int f = 0; f < 2; f++) {                 // same posting order as ny_comm_exchange_z
        double* a = arr[f];
        const size_t cnt = (size_t)nh * plane[f];
        if (below >= 0) NY_NCCL(g_nccl.Send(a + (size_t)nh * plane[f], cnt, ncclDouble, below, c->nccl, st));
        if (above >= 0) NY_NCCL(g_nccl.Send(a + (size_t)nint[f] * plane[f], cnt, ncclDouble, above, c->nccl, st));
        if (above >= 0) NY_NCCL(g_nccl.Recv(a + (size_t)(nh + nint[f]) * plane[f], cnt, ncclDouble, above, c->nccl, st));
        if (below >= 0) NY_NCCL(g_nccl.Recv(a, cnt, ncclDouble, below, c->nccl, st));
    }
    NY_NCCL(g_nccl.GroupEnd());
    return NY_OK;
}

// ---- model halo fill for z slabs (core/mpi/halo.py:140-178) ---------------------------------------
extern "C" int ny_halo_exchange(ny_ctx* ctx, ny_comm* comm, double* const* fields_host, int nfields, ny_ext e, int nh,
                                int below, int above, int yper, int xper, void* stream)
{
    NY_REQUIRE(ctx && fields_host && nfields > 0 && nfields <= 16, "bad argument");
    NY_REQUIRE((below < 0 && above < 0) || comm, "a communicator is required for slab neighbours");
    cudaStream_t st = ny_stream(stream);
    ny_prof_scope ps(ctx, NY_PROF_HALO, st);
    const size_t plane = (size_t)e.ny * e.nx;
    const int lo = below >= 0 ? nh : 0;
    const int nint = e.nz - lo - (above >= 0 ? nh : 0);
    NY_REQUIRE(nint >= nh || (below < 0 && above < 0), "slab thinner than the halo");
    int r = ny_comm_exchange_z(comm, fields_host, nfields, plane, lo, nint, nh, below, above, st);
    if (r != NY_OK) return r;
    if (xper || yper) {
        ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
        for (int f = 0; f < nfields; f++) {
            k_wrap_xy<<<g.grid, g.block, 0, st>>>(fields_host[f], e.nz, e.ny, e.nx, nh, yper, xper);
            NY_CHECK_LAUNCH(ctx);
        }
    }
    return NY_OK;
}

extern "C" int ny_comm_allreduce_host(ny_comm* c, double* values_host, int n, int op_max, void* stream)
{
    NY_REQUIRE(values_host && n >= 1 && n <= 8, "bad argument");
    if (!c || c->nranks == 1) return NY_OK;
    cudaStream_t st = ny_stream(stream);
    NY_CUDA(cudaMemcpyAsync(c->d_red, values_host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    int r = ny_comm_allreduce(c, c->d_red, n, op_max, st);
    if (r != NY_OK) return r;
    NY_CUDA(cudaMemcpyAsync(values_host, c->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    NY_CUDA(cudaStreamSynchronize(st));
    return NY_OK;
}
