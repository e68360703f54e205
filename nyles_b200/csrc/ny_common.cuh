// Shared host/device helpers of libnyles_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include "../../include/nyles_b200.h"

#include <vector>

// kernel families that can be timed individually with CUDA events (ny_prof_*), see nyles_b200.h
struct ny_prof_rec { int tag; cudaEvent_t t0, t1; };

struct ny_ctx {
    int device;
    int num_sms;
    long long launches;
    int fast_arith;           // WENO arithmetic mode, see ny_weno.cuh (0 = strict / bit-exact)
    int mom_variant;          // momentum kernel choice, see ny_set_momentum_variant (0 = by size)
    double* d_scratch;        // reduction partials (device)
    size_t scratch_doubles;
    double* h_pinned;         // small pinned mailbox for scalars
    // event profiling of tagged launch groups
    unsigned long long prof_mask;
    std::vector<ny_prof_rec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[NY_PROF_NTAGS];
    long long prof_n[NY_PROF_NTAGS];
};

// RAII: brackets the launches issued in its scope with two events on `st` when the tag is enabled
struct ny_prof_scope {
    ny_ctx* ctx; int slot; cudaStream_t st;
    ny_prof_scope(ny_ctx* c, int tag, cudaStream_t s) : ctx(c), slot(-1), st(s)
    {
        if (!(c->prof_mask >> tag & 1ull)) return;
        ny_prof_rec r; r.tag = tag;
        for (cudaEvent_t* e : {&r.t0, &r.t1}) {
            if (!c->prof_pool.empty()) { *e = c->prof_pool.back(); c->prof_pool.pop_back(); }
            else if (cudaEventCreate(e) != cudaSuccess) return;
        }
        cudaEventRecord(r.t0, s);
        slot = (int)c->prof_recs.size();
        c->prof_recs.push_back(r);
    }
    ~ny_prof_scope() { if (slot >= 0) cudaEventRecord(ctx->prof_recs[slot].t1, st); }
};

void ny_set_error(const char* fmt, ...);

#define NY_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ny_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return NY_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define NY_CHECK_LAUNCH(ctx)                                                             \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        cudaError_t _e = cudaPeekAtLastError();                                          \
        if (_e != cudaSuccess) {                                                         \
            ny_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return NY_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define NY_REQUIRE(cond, msg)                                                            \
    do {                                                                                 \
        if (!(cond)) { ny_set_error("%s: %s", __func__, msg); return NY_ERR_ARG; }       \
    } while (0)

static inline cudaStream_t ny_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// 3-D launch geometry for cell-parallel kernels: x covers i (coalesced), y covers j, z covers k.
struct ny_grid3 { dim3 grid, block; };
static inline ny_grid3 ny_cells_launch(int nz, int ny, int nx, int bx = 32, int by = 4, int bz = 2)
{
    ny_grid3 g;
    g.block = dim3(bx, by, bz);
    g.grid = dim3((nx + bx - 1) / bx, (ny + by - 1) / by, (nz + bz - 1) / bz);
    return g;
}

