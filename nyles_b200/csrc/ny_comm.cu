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
#include <cstdlib>
#include <vector>

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
    memset(&c->p2p, 0, sizeof(c->p2p));
    c->n_exchanges = 0; c->bytes_sent = 0;
    c->p2p.peer_rank[0] = c->p2p.peer_rank[1] = -1;
    { const char* e = getenv("NY_COMM_P2P"); if (e && e[0] == '0') c->p2p.state = -1; }
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

static void p2p_release(ny_comm* c);

extern "C" void ny_comm_free(ny_comm* c)
{
    if (!c) return;
    p2p_release(c);
    if (c->nccl && g_nccl.ok) g_nccl.CommDestroy(c->nccl);
    if (c->d_red) cudaFree(c->d_red);
    if (c->xstream) cudaStreamDestroy(c->xstream);
    if (c->ev_ready) cudaEventDestroy(c->ev_ready);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    delete c;
}

extern "C" int ny_comm_stats(ny_comm* c, long long* exchanges, long long* bytes_sent, int reset)
{
    NY_REQUIRE(exchanges && bytes_sent, "null argument");
    *exchanges = c ? c->n_exchanges : 0;
    *bytes_sent = c ? c->bytes_sent : 0;
    if (c && reset) { c->n_exchanges = 0; c->bytes_sent = 0; }
    return NY_OK;
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

// ---- P2P halo exchange ---------------------------------------------------------------------------
// ncclSend/ncclRecv moves a 3-plane face at ~90 GB/s per direction (measured, profiles/), a fraction of
// what NVLink 5 carries, and every group is a rendezvous of both ranks inside NCCL.  Here a rank PUSHES its
// boundary planes straight into a receive slot in the neighbour's memory (IPC-mapped, plain 16-byte
// stores over NVLink from a grid of CTAs), raises a flag there once all its stores are globally visible,
// then waits for the flag the neighbour raises in ITS slot and copies the planes into its halo -- all in one
// launch (k_p2p_exchange).
// Slot reuse needs no acknowledgement: exchanges are issued in the same order on every rank and a rank
// cannot run more than one exchange ahead of a neighbour whose data it waits for, so with
// NY_P2P_SLOTS >= 2 (4 here: one exchange may be in flight on the overlap stream as well) a slot is never
// overwritten before its previous content has been unpacked.
struct P2PSeg { const double* src; double* dst; unsigned long long count; };
struct P2PArgs {
    int nseg;
    P2PSeg seg[16];
    unsigned long long* flag[2];      // push: the neighbours' flags to raise; unpack: the local flags to wait for
    unsigned int* counter;
    unsigned long long seq;
};

__device__ __forceinline__ void p2p_copy(const P2PSeg& s, long long first, long long stride)
{
    const bool vec = ((((unsigned long long)s.src) | ((unsigned long long)s.dst)) & 15ull) == 0 && (s.count & 1ull) == 0;
    if (vec) {
        const double2* a = reinterpret_cast<const double2*>(s.src);
        double2* b = reinterpret_cast<double2*>(s.dst);
        const long long n = (long long)(s.count >> 1);
        long long t = first;
        for (; t + 3 * stride < n; t += 4 * stride) {          // four independent 16-byte transfers in flight per thread
            const double2 v0 = a[t], v1 = a[t + stride], v2 = a[t + 2 * stride], v3 = a[t + 3 * stride];
            b[t] = v0; b[t + stride] = v1; b[t + 2 * stride] = v2; b[t + 3 * stride] = v3;
        }
        for (; t < n; t += stride) b[t] = a[t];
    } else {
        for (long long t = first; t < (long long)s.count; t += stride) s.dst[t] = s.src[t];
    }
}

__device__ __forceinline__ void p2p_wait_flags(const P2PArgs& a)
{
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int d = 0; d < 2; d++) {
        if (!a.flag[d]) continue;
        const volatile unsigned long long* f = a.flag[d];
        while (*f < a.seq) {
            __nanosleep(100);
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 60000000000ull) {     // a neighbour died: fail loudly instead of hanging the GPU
                printf("libnyles_b200: halo exchange %llu timed out waiting for a slab neighbour\n", a.seq);
                __trap();
            }
        }
    }
    __threadfence_system();
}

// ONE launch per exchange: every CTA pushes its share of the boundary planes into the neighbours' slots; the last
// one to finish raises the arrival flags over there; then every CTA waits for the flags the neighbours raise HERE
// and unpacks its share of the received planes into the halo.  The grid is never larger than what is resident at
// once (2 CTAs of 512 threads per SM), so CTAs that spin on a flag cannot keep CTAs that still have to push off the
// machine.
__global__ void __launch_bounds__(512)
k_p2p_exchange(P2PArgs push, P2PArgs pull)
{
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    for (int s = 0; s < push.nseg; s++) p2p_copy(push.seg[s], first, stride);
    __threadfence_system();                         // my stores are visible to the neighbour's GPU ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(push.counter, 1u);
        if (done == gridDim.x - 1) {                // ... and so are those of every other CTA: raise the flags
            *push.counter = 0u;
            __threadfence_system();
            if (push.flag[0]) *reinterpret_cast<volatile unsigned long long*>(push.flag[0]) = push.seq;
            if (push.flag[1]) *reinterpret_cast<volatile unsigned long long*>(push.flag[1]) = push.seq;
        }
        p2p_wait_flags(pull);
    }
    __syncthreads();
    for (int s = 0; s < pull.nseg; s++) p2p_copy(pull.seg[s], first, stride);
}

static void p2p_release(ny_comm* c)
{
    ny_p2p& p = c->p2p;
    if (p.peer[0]) cudaIpcCloseMemHandle(p.peer[0]);
    if (p.peer[1] && p.peer[1] != p.peer[0]) cudaIpcCloseMemHandle(p.peer[1]);
    if (p.local) cudaFree(p.local);
    p.peer[0] = p.peer[1] = nullptr; p.local = nullptr; p.flags = nullptr; p.counters = nullptr;
    p.slot_bytes = 0;
    if (p.state == 1) p.state = 0;
}

// Collective: (re)allocate the receive slots for faces of up to need_bytes per direction and map the
// neighbours' buffers.  Every rank calls it at the same point with the same size (slabs are equal).
static int p2p_setup(ny_comm* c, size_t need_bytes)
{
    ny_p2p& p = c->p2p;
    // Always the two RING neighbours, whether or not the current geometry wraps in z: which exchanges have a partner
    // below rank 0 / above rank P-1 changes with the topology of the model that is running, and a re-mapping decided
    // from that would be entered by the edge ranks only -- a collective that the middle ranks never join (found at
    // 8 ranks: closed -> perio_xyz in one process hung ranks 0 and 7 in the all-reduce below).
    const int below = (c->rank + c->nranks - 1) % c->nranks, above = (c->rank + 1) % c->nranks;
    NY_CUDA(cudaDeviceSynchronize());
    {   // nobody may still be using the old mapping
        double* z = c->d_red;
        NY_CUDA(cudaMemset(z, 0, sizeof(double)));
        NY_NCCL(g_nccl.AllReduce(z, z, 1, ncclDouble, ncclSum, c->nccl, 0));
        NY_CUDA(cudaDeviceSynchronize());
    }
    p2p_release(c);
    const size_t slot = ((need_bytes + need_bytes / 4 + (1u << 20) - 1) >> 20) << 20;       // 25 % head room, 1 MiB granules
    const size_t data = 2 * (size_t)NY_P2P_SLOTS * slot, total = data + 4096;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&p.local, total) != cudaSuccess) { cudaGetLastError(); p.local = nullptr; ok = 0; }
    if (ok) {
        cudaMemset(p.local + data, 0, 4096);
        if (cudaIpcGetMemHandle(&mine, p.local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    }
    // all-gather the 64-byte handles through NCCL
    unsigned char* d_h = nullptr;
    NY_CUDA(cudaMalloc(&d_h, (size_t)c->nranks * sizeof(cudaIpcMemHandle_t)));
    NY_CUDA(cudaMemcpy(d_h + (size_t)c->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NY_NCCL(g_nccl.AllGather(d_h + (size_t)c->rank * sizeof(mine), d_h, sizeof(mine), ncclChar, c->nccl, 0));
    std::vector<cudaIpcMemHandle_t> all(c->nranks);
    NY_CUDA(cudaMemcpy(all.data(), d_h, (size_t)c->nranks * sizeof(mine), cudaMemcpyDeviceToHost));
    cudaFree(d_h);
    const int nb[2] = {below, above};
    for (int d = 0; d < 2 && ok; d++) {
        if (nb[d] < 0) continue;
        if (d == 1 && above == below) { p.peer[1] = p.peer[0]; continue; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[nb[d]], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        p.peer[d] = static_cast<unsigned char*>(q);
    }
    // everybody or nobody
    double flag = ok ? 0.0 : 1.0;
    NY_CUDA(cudaMemcpy(c->d_red, &flag, sizeof(double), cudaMemcpyHostToDevice));
    NY_NCCL(g_nccl.AllReduce(c->d_red, c->d_red, 1, ncclDouble, ncclSum, c->nccl, 0));
    NY_CUDA(cudaMemcpy(&flag, c->d_red, sizeof(double), cudaMemcpyDeviceToHost));
    if (flag != 0.0) {
        p2p_release(c);
        p.state = -1;
        if (c->rank == 0)
            fprintf(stderr, "libnyles_b200: CUDA IPC peer mapping is not available on %d rank(s); halo faces go through "
                            "ncclSend/ncclRecv\n", (int)flag);
        return NY_OK;
    }
    p.slot_bytes = slot;
    p.flags = reinterpret_cast<unsigned long long*>(p.local + data);
    p.counters = reinterpret_cast<unsigned int*>(p.local + data + 2048);
    p.peer_rank[0] = below; p.peer_rank[1] = above;
    p.state = 1;
    return NY_OK;
}

// faces of nf arrays (plane sizes / interior thicknesses may differ) through the peer slots; returns
// NY_OK with *done = false when P2P is not available
static int p2p_exchange(ny_comm* c, double* const* arrays, const size_t* plane, const int* lo, const int* nint, int nf,
                        int nh, int below, int above, cudaStream_t st, bool* done)
{
    ny_p2p& p = c->p2p;
    *done = false;
    if (p.state < 0 || nf > 8) return NY_OK;
    size_t bytes = 0;
    for (int f = 0; f < nf; f++) bytes += (size_t)nh * plane[f] * sizeof(double);
    // (re)allocation is decided from values that are identical on every rank: the state and the face size
    if (p.state == 0 || bytes > p.slot_bytes) {
        // room for four such faces: the first exchange of a run is a single field, vectors and the
        // multigrid's padded planes follow
        int r = p2p_setup(c, p.state == 0 ? 4 * bytes : bytes);
        if (r != NY_OK) return r;
        if (p.state != 1) return NY_OK;
    }
    if ((below >= 0 && below != p.peer_rank[0]) || (above >= 0 && above != p.peer_rank[1])) {
        ny_set_error("p2p_exchange: neighbours (%d, %d) are not the ring neighbours (%d, %d) of rank %d", below, above,
                     p.peer_rank[0], p.peer_rank[1], c->rank);
        return NY_ERR_ARG;
    }
    const unsigned long long seq = ++p.seq;
    const int slot = (int)(seq % NY_P2P_SLOTS);
    const size_t data = 2 * (size_t)NY_P2P_SLOTS * p.slot_bytes;
    // region 0 of a buffer receives from the rank below, region 1 from the rank above
    auto slot_of = [&](unsigned char* base, int region) {
        return reinterpret_cast<double*>(base + ((size_t)region * NY_P2P_SLOTS + slot) * p.slot_bytes);
    };
    auto flag_of = [&](unsigned char* base, int region) {
        return reinterpret_cast<unsigned long long*>(base + data) + region * NY_P2P_SLOTS + slot;
    };
    P2PArgs push, pull;
    memset(&push, 0, sizeof(push));
    memset(&pull, 0, sizeof(pull));
    push.seq = pull.seq = seq;
    push.counter = p.counters + slot;
    size_t off = 0;
    for (int f = 0; f < nf; f++) {
        double* a = arrays[f];
        const size_t cnt = (size_t)nh * plane[f];
        if (below >= 0) {      // my lowest interior planes -> the "from above" slot of the rank below
            push.seg[push.nseg++] = {a + (size_t)lo[f] * plane[f], slot_of(p.peer[0], 1) + off, cnt};
            pull.seg[pull.nseg++] = {slot_of(p.local, 0) + off, a + (size_t)(lo[f] - nh) * plane[f], cnt};
        }
        if (above >= 0) {      // my highest interior planes -> the "from below" slot of the rank above
            push.seg[push.nseg++] = {a + (size_t)(lo[f] + nint[f] - nh) * plane[f], slot_of(p.peer[1], 0) + off, cnt};
            pull.seg[pull.nseg++] = {slot_of(p.local, 1) + off, a + (size_t)(lo[f] + nint[f]) * plane[f], cnt};
        }
        off += cnt;
    }
    if (below >= 0) { push.flag[0] = flag_of(p.peer[0], 1); pull.flag[0] = flag_of(p.local, 0); }
    if (above >= 0) { push.flag[1] = flag_of(p.peer[1], 0); pull.flag[1] = flag_of(p.local, 1); }
    const size_t moved = bytes * ((below >= 0) + (above >= 0));
    c->n_exchanges++; c->bytes_sent += (long long)moved;
    const int sms = c->ctx->num_sms > 0 ? c->ctx->num_sms : 148;
    int nblk = (int)(moved >> 16);                  // one CTA per 64 KiB, 4 .. 2 per SM
    nblk = nblk < 4 ? 4 : (nblk > 2 * sms ? 2 * sms : nblk);
    k_p2p_exchange<<<nblk, 512, 0, st>>>(push, pull);
    NY_CHECK_LAUNCH(c->ctx);
    *done = true;
    return NY_OK;
}

// Exchange the nh z-faces of nf slab arrays with the ranks below / above (-1: none).
// Array a: planes of `plane` doubles; lo = index of the first interior plane, nint = interior planes.
int ny_comm_exchange_z(ny_comm* c, double* const* arrays, int nf, size_t plane, int lo, int nint, int nh,
                       int below, int above, cudaStream_t st)
{
    if (!c || (below < 0 && above < 0)) return NY_OK;
    if (nf <= 8) {
        size_t planes[8]; int los[8], nints[8];
        for (int f = 0; f < nf; f++) { planes[f] = plane; los[f] = lo; nints[f] = nint; }
        bool done = false;
        int r = p2p_exchange(c, arrays, planes, los, nints, nf, nh, below, above, st, &done);
        if (r != NY_OK || done) return r;
    }
    const size_t cnt = (size_t)nh * plane;
    c->n_exchanges++; c->bytes_sent += (long long)(cnt * sizeof(double) * nf * ((below >= 0) + (above >= 0)));
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
    {
        const int los[2] = {nh, nh};
        bool done = false;
        int r = p2p_exchange(c, arr, plane, los, nint, 2, nh, below, above, st, &done);
        if (r != NY_OK || done) return r;
    }
    c->n_exchanges++;
    c->bytes_sent += (long long)((size_t)nh * (plane0 + plane1) * sizeof(double) * ((below >= 0) + (above >= 0)));
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
