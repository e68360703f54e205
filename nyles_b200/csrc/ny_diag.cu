// Diagnostic / projection-side kernels, time-scheme updates, reductions and the one-process halo.
//
// Replaces core/fortran_vorticity.f90:2-28 (driver core/vorticity.py:7-34),
// core/fortran_kinenergy.f90:3-56 (core/kinenergy.py:7-24), core/fortran_bernoulli.f90:2-26,61-97
// (core/projection.py:16-29,84-87), core/cov_to_contra.py:4-20, the NumPy updates of
// core/timescheme.py:113-221, core/nyles.py:244-250 and core/mpi/halo.py:140-178.
#include "ny_common.cuh"
#include <cstdlib>

namespace {

struct Ext { int nz, ny, nx; long long sj, sk; };
inline Ext make_ext(ny_ext e)
{
    Ext x; x.nz = e.nz; x.ny = e.ny; x.nx = e.nx; x.sj = e.nx; x.sk = (long long)e.nx * e.ny; return x;
}

#define CELL_INDEX                                                    \
    int i = blockIdx.x * blockDim.x + threadIdx.x;                    \
    int j = blockIdx.y * blockDim.y + threadIdx.y;                    \
    int k = blockIdx.z * blockDim.z + threadIdx.z;                    \
    if (i >= e.nx || j >= e.ny || k >= e.nz) return;                  \
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;

// omega_z: rows j<ny-1: cols i<nx-1 computed, col nx-1 zero; row ny-1 untouched.  Cyclic for x,y.
__global__ void __launch_bounds__(256)
k_vorticity(const double* __restrict__ ux, const double* __restrict__ uy, const double* __restrict__ uz,
            double* __restrict__ wx, double* __restrict__ wy, double* __restrict__ wz, double fparam, Ext e)
{
    CELL_INDEX
    // dirk='i' (arrays [i,k,j]): omega_x, valid k<nz-1; j<ny-1 computed, j=ny-1 zero
    if (k < e.nz - 1) {
        if (j < e.ny - 1) wx[c] = uz[c + e.sj] - uz[c] - uy[c + e.sk] + uy[c];
        else wx[c] = 0.0;
    }
    // dirk='j' (arrays [j,i,k]): omega_y, valid i<nx-1; k<nz-1 computed, k=nz-1 zero
    if (i < e.nx - 1) {
        if (k < e.nz - 1) wy[c] = ux[c + e.sk] - ux[c] - uz[c + 1] + uz[c];
        else wy[c] = 0.0;
    }
    // dirk='k' (arrays [k,j,i]): omega_z, valid j<ny-1; i<nx-1 computed, i=nx-1 zero
    if (j < e.ny - 1) {
        if (i < e.nx - 1) {
            double w = uy[c + 1] - uy[c] - ux[c + e.sj] + ux[c];
            if (fparam > 0.0) w = w + fparam;                        // vorticity.py:33-34
            wz[c] = w;
        } else wz[c] = 0.0;
    }
}

__global__ void __launch_bounds__(256)
k_kin(const double* __restrict__ ux, const double* __restrict__ uy, const double* __restrict__ uz,
      double* __restrict__ ke, double cx, double cy, double cz, Ext e)
{
    CELL_INDEX
    double acc = 0.0;                                                // kinenergy.py:19
    if (i > 0) { double a = ux[c], b = ux[c - 1]; acc = acc + cx * (a * a + b * b); }
    if (j > 0) { double a = uy[c], b = uy[c - e.sj]; acc = acc + cy * (a * a + b * b); }
    if (k > 0) { double a = uz[c], b = uz[c - e.sk]; acc = acc + cz * (a * a + b * b); }
    ke[c] = acc;
}

constexpr unsigned NY_MAX_SLOTS = 4096;   // atomicMax targets of k_diag_post (spread to avoid contention)

// U_from_u + vorticity + kinenergy in one pass over u (same statements as k_scale3, k_vorticity, k_kin)
__global__ void __launch_bounds__(256)
k_diag_post(const double* __restrict__ ux, const double* __restrict__ uy, const double* __restrict__ uz,
            double* __restrict__ Ux, double* __restrict__ Uy, double* __restrict__ Uz,
            double* __restrict__ wx, double* __restrict__ wy, double* __restrict__ wz, double* __restrict__ ke,
            double idx2, double idy2, double idz2, double cx, double cy, double cz, double fparam, Ext e,
            unsigned long long* __restrict__ maxbits)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    const bool inside = i < e.nx && j < e.ny && k < e.nz;
    // max(U^2+V^2+W^2) of core/nyles.py:244-250 on the fly: non-negative doubles (and +NaN, which then
    // wins) order like their bit patterns, so an integer atomicMax gives the exact maximum
    unsigned long long qb = 0ull;
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    double u0 = 0.0, v0 = 0.0, w0 = 0.0;
    if (inside) {
        u0 = ux[c]; v0 = uy[c]; w0 = uz[c];
        const double U0 = u0 * idx2, V0 = v0 * idy2, W0 = w0 * idz2;
        Ux[c] = U0; Uy[c] = V0; Uz[c] = W0;
        const double q = U0 * U0 + V0 * V0 + W0 * W0;
        qb = (q == q) ? (unsigned long long)__double_as_longlong(q) : 0x7ff8000000000000ull;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, qb, o);
        qb = other > qb ? other : qb;
    }
    if ((threadIdx.x & 31) == 0) {
        const unsigned slot = ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8u +
                              ((threadIdx.z * blockDim.y + threadIdx.y) * blockDim.x + threadIdx.x) / 32u;
        atomicMax(maxbits + (slot & (NY_MAX_SLOTS - 1)), qb);
    }
    if (!inside) return;
    const bool ip = i < e.nx - 1, jp = j < e.ny - 1, kp = k < e.nz - 1;
    if (kp) wx[c] = jp ? uz[c + e.sj] - w0 - uy[c + e.sk] + v0 : 0.0;
    if (ip) wy[c] = kp ? ux[c + e.sk] - u0 - uz[c + 1] + w0 : 0.0;
    if (jp) {
        if (ip) {
            double w = uy[c + 1] - v0 - ux[c + e.sj] + u0;
            if (fparam > 0.0) w = w + fparam;
            wz[c] = w;
        } else wz[c] = 0.0;
    }
    double acc = 0.0;
    if (i > 0) { double q = ux[c - 1]; acc = acc + cx * (u0 * u0 + q * q); }
    if (j > 0) { double q = uy[c - e.sj]; acc = acc + cy * (v0 * v0 + q * q); }
    if (k > 0) { double q = uz[c - e.sk]; acc = acc + cz * (w0 * w0 + q * q); }
    ke[c] = acc;
}

__global__ void __launch_bounds__(256)
k_div(const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
      double* __restrict__ div, Ext e)
{
    CELL_INDEX
    double d = (i > 0) ? (Ux[c] - Ux[c - 1]) : Ux[c];                // iflag = 0: overwrite
    d = (j > 0) ? d + (Uy[c] - Uy[c - e.sj]) : d + Uy[c];            // iflag > 0: accumulate
    d = (k > 0) ? d + (Uz[c] - Uz[c - e.sk]) : d + Uz[c];
    div[c] = d;
}

__global__ void __launch_bounds__(256)
k_gradp(const double* __restrict__ p, double* __restrict__ ux, double* __restrict__ uy,
        double* __restrict__ uz, Ext e)
{
    CELL_INDEX
    const double p0 = p[c];
    if (i < e.nx - 1) ux[c] = ux[c] - (p[c + 1] - p0);
    if (j < e.ny - 1) uy[c] = uy[c] - (p[c + e.sj] - p0);
    if (k < e.nz - 1) uz[c] = uz[c] - (p[c + e.sk] - p0);
}

__global__ void __launch_bounds__(256)
k_scale3(const double* __restrict__ a0, const double* __restrict__ a1, const double* __restrict__ a2,
         double* __restrict__ o0, double* __restrict__ o1, double* __restrict__ o2,
         double s0, double s1, double s2, long long n)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride) { o0[t] = a0[t] * s0; o1[t] = a1[t] * s1; o2[t] = a2[t] * s2; }
}

// ---- time schemes: elementwise, grid-stride --------------------------------------------
enum { TS_AXPY, TS_LF_FIRST, TS_LF_PRED, TS_LF_CORR, TS_RK2, TS_RK3 };
template <int MODE>
__global__ void __launch_bounds__(256)
k_ts(double* __restrict__ s, const double* __restrict__ d0, const double* __restrict__ d1,
     const double* __restrict__ d2, double* __restrict__ sb, double* __restrict__ sn,
     double a, long long n)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; t < n; t += stride) {
        if (MODE == TS_AXPY) {                       // s += a*ds
            s[t] = s[t] + a * d0[t];
        } else if (MODE == TS_LF_FIRST) {            // timescheme.py:131-139, a = dt
            double v = s[t];
            sn[t] = v; sb[t] = v;
            s[t] = v + a * d0[t];
        } else if (MODE == TS_LF_PRED) {             // timescheme.py:144-162, a = dt
            double v = s[t], vb = sb[t];
            double lf = vb + (2. * a) * d0[t];
            s[t] = (1. / 12.) * (5. * lf + 8. * v - vb);
            sn[t] = v; sb[t] = v;
        } else if (MODE == TS_LF_CORR) {             // timescheme.py:170-175
            s[t] = sn[t] + a * d0[t];
        } else if (MODE == TS_RK2) {                 // timescheme.py:206-211, a = dt
            s[t] = s[t] + (a / 4.) * (d1[t] - 3 * d0[t]);
        } else {                                     // timescheme.py:214-220
            s[t] = s[t] + (a / 12.) * (8 * d2[t] - d0[t] - d1[t]);
        }
    }
}

// ---- max(U^2+V^2+W^2): two-stage deterministic reduction ---------------------------------
__global__ void __launch_bounds__(256)
k_maxspeed_partial(const double* __restrict__ U, const double* __restrict__ V, const double* __restrict__ W,
                   long long n, double* __restrict__ partial)
{
    __shared__ double sh[8];
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    double m = 0.0;
    bool bad = false;
    for (; t < n; t += stride) {
        double u = U[t], v = V[t], w = W[t];
        double q = u * u + v * v + w * w;
        if (!(q == q)) bad = true;
        m = fmax(m, q);
    }
    if (bad) m = nan("");
    // NaN-propagating max inside the block
    for (int o = 16; o > 0; o >>= 1) {
        double other = __shfl_xor_sync(0xffffffffu, m, o);
        m = (m != m || other != other) ? nan("") : fmax(m, other);
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = sh[0];
        for (int w = 1; w < (blockDim.x >> 5); w++) r = (r != r || sh[w] != sh[w]) ? nan("") : fmax(r, sh[w]);
        partial[blockIdx.x] = r;
    }
}
__global__ void __launch_bounds__(256)
k_maxbits_final(const unsigned long long* __restrict__ bits, double* __restrict__ out)
{
    __shared__ unsigned long long sh[256];
    unsigned long long m = 0ull;
    for (unsigned t = threadIdx.x; t < NY_MAX_SLOTS; t += blockDim.x) m = bits[t] > m ? bits[t] : m;
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x + o] > sh[threadIdx.x] ? sh[threadIdx.x + o] : sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = __longlong_as_double((long long)sh[0]);
}
__global__ void k_max_final(const double* __restrict__ partial, int nb, double* __restrict__ out)
{
    double r = partial[0];
    for (int b = 1; b < nb; b++) r = (r != r || partial[b] != partial[b]) ? nan("") : fmax(r, partial[b]);
    out[0] = r;
}

// ---- periodic halo of a model array: 26 disjoint halo boxes, sources in the interior ------
// per-axis geometry: lo = first interior index (nh or 0), n = interior extent, tot = array extent
struct HaloGeom { int lo[3], n[3], tot[3], per[3], nh; };
__global__ void __launch_bounds__(256)
k_halo_self(double* __restrict__ f, HaloGeom g)
{
    // every thread owns one array cell; cells in a halo slab of a periodic axis copy from the
    // wrapped interior position (core/mpi/halo.py:93-120: interior strips only are ever sent).
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= g.tot[2] || j >= g.tot[1] || k >= g.tot[0]) return;
    int idx[3] = {k, j, i};
    int src[3];
    bool halo = false;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        src[a] = idx[a];
        if (g.per[a]) {
            if (idx[a] < g.lo[a]) { src[a] = idx[a] + g.n[a]; halo = true; }
            else if (idx[a] >= g.lo[a] + g.n[a]) { src[a] = idx[a] - g.n[a]; halo = true; }
        }
    }
    if (!halo) return;
    long long sj = g.tot[2], sk = (long long)g.tot[2] * g.tot[1];
    f[(long long)k * sk + (long long)j * sj + i] = f[(long long)src[0] * sk + (long long)src[1] * sj + src[2]];
}

inline int ts_blocks(ny_ctx* ctx, long long n)
{
    long long b = (n + 255) / 256;
    long long cap = (long long)ctx->num_sms * 16;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

extern "C" int ny_vorticity(ny_ctx* ctx, const double* ux, const double* uy, const double* uz,
                            double* wx, double* wy, double* wz, ny_ext e, double fparam, void* stream)
{
    NY_REQUIRE(ctx && ux && uy && uz && wx && wy && wz, "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_VORT_KE, ny_stream(stream));
    k_vorticity<<<g.grid, g.block, 0, ny_stream(stream)>>>(ux, uy, uz, wx, wy, wz, fparam, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_kin(ny_ctx* ctx, const double* ux, const double* uy, const double* uz, double* ke,
                      double idx2, double idy2, double idz2, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && ux && uy && uz && ke, "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    // fortran_kinenergy.f90:43,48: cff2 = 0.5*ds2; ke += cff2*0.5*(...)
    ny_prof_scope ps(ctx, NY_PROF_VORT_KE, ny_stream(stream));
    k_kin<<<g.grid, g.block, 0, ny_stream(stream)>>>(ux, uy, uz, ke, (0.5 * idx2) * 0.5, (0.5 * idy2) * 0.5,
                                                      (0.5 * idz2) * 0.5, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_diag_post(ny_ctx* ctx, const double* ux, const double* uy, const double* uz,
                            double* Ux, double* Uy, double* Uz, double* wx, double* wy, double* wz, double* ke,
                            double idx2, double idy2, double idz2, double fparam, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && ux && uy && uz && Ux && Uy && Uz && wx && wy && wz && ke, "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_VORT_KE, ny_stream(stream));
    NY_REQUIRE(ctx->scratch_doubles >= 32768 + NY_MAX_SLOTS, "scratch too small");
    unsigned long long* maxbits = reinterpret_cast<unsigned long long*>(ctx->d_scratch + 32768);
    NY_CUDA(cudaMemsetAsync(maxbits, 0, NY_MAX_SLOTS * sizeof(unsigned long long), ny_stream(stream)));
    k_diag_post<<<g.grid, g.block, 0, ny_stream(stream)>>>(ux, uy, uz, Ux, Uy, Uz, wx, wy, wz, ke, idx2, idy2, idz2,
                                                           (0.5 * idx2) * 0.5, (0.5 * idy2) * 0.5, (0.5 * idz2) * 0.5,
                                                           fparam, make_ext(e), maxbits);
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_diag_post_max_speed2(ny_ctx* ctx, double* out_host, void* stream)
{
    NY_REQUIRE(ctx && out_host, "null argument");
    cudaStream_t st = ny_stream(stream);
    ny_prof_scope ps(ctx, NY_PROF_MAXSPEED, st);
    k_maxbits_final<<<1, 256, 0, st>>>(reinterpret_cast<unsigned long long*>(ctx->d_scratch + 32768), ctx->d_scratch);
    NY_CHECK_LAUNCH(ctx);
    NY_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
    NY_CUDA(cudaStreamSynchronize(st));
    *out_host = ctx->h_pinned[0];
    return NY_OK;
}

extern "C" int ny_div(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz, double* div,
                      ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && Ux && Uy && Uz && div, "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_DIV, ny_stream(stream));
    k_div<<<g.grid, g.block, 0, ny_stream(stream)>>>(Ux, Uy, Uz, div, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_gradp(ny_ctx* ctx, const double* p, double* ux, double* uy, double* uz, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && p && ux && uy && uz, "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_GRADP, ny_stream(stream));
    k_gradp<<<g.grid, g.block, 0, ny_stream(stream)>>>(p, ux, uy, uz, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_U_from_u(ny_ctx* ctx, const double* ux, const double* uy, const double* uz,
                           double* Ux, double* Uy, double* Uz, double idx2, double idy2, double idz2,
                           ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && ux && uy && uz && Ux && Uy && Uz, "null argument");
    long long n = (long long)e.nz * e.ny * e.nx;
    ny_prof_scope ps(ctx, NY_PROF_U_FROM_U, ny_stream(stream));
    k_scale3<<<ts_blocks(ctx, n), 256, 0, ny_stream(stream)>>>(ux, uy, uz, Ux, Uy, Uz, idx2, idy2, idz2, n);
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

#define TS_LAUNCH(MODE, ...)                                                                     \
    ny_prof_scope ps(ctx, NY_PROF_TIMESCHEME, ny_stream(stream));                                \
    k_ts<MODE><<<ts_blocks(ctx, n), 256, 0, ny_stream(stream)>>>(__VA_ARGS__);                   \
    NY_CHECK_LAUNCH(ctx);                                                                        \
    return NY_OK;

extern "C" int ny_ts_axpy(ny_ctx* ctx, double* s, const double* ds, double a, long long n, void* stream)
{
    NY_REQUIRE(ctx && s && ds, "null argument");
    TS_LAUNCH(TS_AXPY, s, ds, nullptr, nullptr, nullptr, nullptr, a, n)
}
extern "C" int ny_ts_lfam3_first(ny_ctx* ctx, double* s, const double* ds, double* sb, double* sn, double dt,
                                 long long n, void* stream)
{
    NY_REQUIRE(ctx && s && ds && sb && sn, "null argument");
    TS_LAUNCH(TS_LF_FIRST, s, ds, nullptr, nullptr, sb, sn, dt, n)
}
extern "C" int ny_ts_lfam3_pred(ny_ctx* ctx, double* s, const double* ds, double* sb, double* sn, double dt,
                                long long n, void* stream)
{
    NY_REQUIRE(ctx && s && ds && sb && sn, "null argument");
    TS_LAUNCH(TS_LF_PRED, s, ds, nullptr, nullptr, sb, sn, dt, n)
}
extern "C" int ny_ts_lfam3_corr(ny_ctx* ctx, double* s, const double* ds, const double* sn, double dt,
                                long long n, void* stream)
{
    NY_REQUIRE(ctx && s && ds && sn, "null argument");
    TS_LAUNCH(TS_LF_CORR, s, ds, nullptr, nullptr, nullptr, const_cast<double*>(sn), dt, n)
}
extern "C" int ny_ts_rk3_stage2(ny_ctx* ctx, double* s, const double* ds0, const double* ds1, double dt,
                                long long n, void* stream)
{
    NY_REQUIRE(ctx && s && ds0 && ds1, "null argument");
    TS_LAUNCH(TS_RK2, s, ds0, ds1, nullptr, nullptr, nullptr, dt, n)
}
extern "C" int ny_ts_rk3_stage3(ny_ctx* ctx, double* s, const double* ds0, const double* ds1, const double* ds2,
                                double dt, long long n, void* stream)
{
    NY_REQUIRE(ctx && s && ds0 && ds1 && ds2, "null argument");
    TS_LAUNCH(TS_RK3, s, ds0, ds1, ds2, nullptr, nullptr, dt, n)
}

extern "C" int ny_max_speed2(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz, long long n,
                             double* out_host, void* stream)
{
    NY_REQUIRE(ctx && Ux && Uy && Uz && out_host && n > 0, "null argument");
    int nb = ts_blocks(ctx, n);
    NY_REQUIRE((size_t)nb + 1 <= ctx->scratch_doubles, "scratch too small");
    cudaStream_t st = ny_stream(stream);
    ny_prof_scope ps(ctx, NY_PROF_MAXSPEED, ny_stream(stream));
    k_maxspeed_partial<<<nb, 256, 0, st>>>(Ux, Uy, Uz, n, ctx->d_scratch + 1);
    NY_CHECK_LAUNCH(ctx);
    k_max_final<<<1, 1, 0, st>>>(ctx->d_scratch + 1, nb, ctx->d_scratch);
    NY_CHECK_LAUNCH(ctx);
    NY_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
    NY_CUDA(cudaStreamSynchronize(st));
    *out_host = ctx->h_pinned[0];
    return NY_OK;
}

extern "C" int ny_halo_fill_self(ny_ctx* ctx, double* f, ny_ext e, int nh, const int per[3], void* stream)
{
    NY_REQUIRE(ctx && f && per, "null argument");
    if (!per[0] && !per[1] && !per[2]) return NY_OK;
    HaloGeom g;
    int tot[3] = {e.nz, e.ny, e.nx};
    for (int a = 0; a < 3; a++) {
        g.per[a] = per[a] ? 1 : 0;
        g.tot[a] = tot[a];
        g.lo[a] = per[a] ? nh : 0;
        g.n[a] = tot[a] - (per[a] ? 2 * nh : 0);
        NY_REQUIRE(g.n[a] >= nh || !per[a], "interior narrower than the halo (core/mpi/halo.py check_halo_width)");
    }
    g.nh = nh;
    ny_grid3 l = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_HALO, ny_stream(stream));
    k_halo_self<<<l.grid, l.block, 0, ny_stream(stream)>>>(f, g);
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}
