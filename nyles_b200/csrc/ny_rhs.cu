// Right-hand-side kernels of the Nyles LES step: tracer advection, vortex force, Bernoulli term.
//
// Replaces core/fortran_upwind.f90:3-87, core/fortran_vortex_force.f90:10-165,
// core/fortran_bernoulli.f90:2-58 as driven by core/tracer.py:44-72, core/vortex_force.py:69-81,
// core/bernoulli.py:14-32.  One canonical (k,j,i) layout; every line sweep of the Fortran is
// evaluated per cell so that all global loads are coalesced along i whatever the sweep axis.
// Accumulation order per output cell equals the order of the reference's passes.
#include "ny_common.cuh"
#include "ny_weno.cuh"

namespace {

struct Ext { int nz, ny, nx; long long sj, sk; };
__host__ __device__ inline Ext make_ext(ny_ext e)
{
    Ext x; x.nz = e.nz; x.ny = e.ny; x.nx = e.nx; x.sj = e.nx; x.sk = (long long)e.nx * e.ny; return x;
}

// ---- tracer: flux through the face on the + side of line position s ----------------------
// q = tracer, u = contravariant velocity of the sweep axis, both sampled along the line.
__device__ __forceinline__ double tracer_flux(const double* __restrict__ trac, const double* __restrict__ U,
                                              long long base, long long stride, int s, int n)
{
    double u = U[base + (long long)s * stride];
    return nyw::line_flux(s, n, u, [&](int t) { return trac[base + (long long)t * stride]; });
}

// one pass of fortran_upwind along one axis for the cell at line position s
__device__ __forceinline__ double upwind_axis(double acc, const double* __restrict__ trac,
                                              const double* __restrict__ U, long long base,
                                              long long stride, int s, int n)
{
    double fp = tracer_flux(trac, U, base, stride, s, n);
    if (s == 0) return acc - fp;                                  // fortran_upwind.f90:75-76
    double fm = tracer_flux(trac, U, base, stride, s - 1, n);
    return acc + fm - fp;                                         // :77-79
}

// zero-flux Laplacian increment of fortran_dissipation.f90:2-35 for the cell at line position s
__device__ __forceinline__ double lap_axis(double acc, const double* __restrict__ phi, long long c,
                                           long long stride, int s, int n, double coef)
{
    const double p0 = phi[c];
    double fx = (s < n - 1) ? phi[c + stride] - p0 : 0.0;
    double fxm = (s > 0) ? p0 - phi[c - stride] : 0.0;
    return acc + coef * (fx - fxm);
}

// DIFF: tracer.py:72-77 interleaves add_laplacian after the upwind pass of each direction
template <bool DIFF>
__global__ void __launch_bounds__(256)
k_upwind(const double* __restrict__ trac, const double* __restrict__ Ux, const double* __restrict__ Uy,
         const double* __restrict__ Uz, double* __restrict__ dtrac, double cx, double cy, double cz, Ext e)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= e.nx || j >= e.ny || k >= e.nz) return;
    long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    double acc = 0.0;                                             // tracer.py:70-71
    acc = upwind_axis(acc, trac, Ux, c - i, 1, i, e.nx);
    if (DIFF) acc = lap_axis(acc, trac, c, 1, i, e.nx, cx);
    acc = upwind_axis(acc, trac, Uy, c - (long long)j * e.sj, e.sj, j, e.ny);
    if (DIFF) acc = lap_axis(acc, trac, c, e.sj, j, e.ny, cy);
    acc = upwind_axis(acc, trac, Uz, c - (long long)k * e.sk, e.sk, k, e.nz);
    if (DIFF) acc = lap_axis(acc, trac, c, e.sk, k, e.nz, cz);
    dtrac[c] = acc;
}

// ---- vortex force: flux of one sweep for the cell at line position s ------------------------
// US: velocity along the sweep axis; tstride: stride of the target component's own axis (the
// averaging axis); W: the vorticity component normal to (sweep, target).
__device__ __forceinline__ double vf_flux(const double* __restrict__ US, const double* __restrict__ W,
                                          long long base, long long stride, long long tstride, int s, int n)
{
    long long cs = base + (long long)s * stride;
    double UU_1 = 0.5 * (US[cs] + US[cs + tstride]);
    double UU_0 = (s > 0) ? 0.5 * (US[cs - stride] + US[cs - stride + tstride]) : 0.0;
    double u1d = 0.5 * (UU_0 + UU_1);                             // fortran_vortex_force.f90:68-72
    return nyw::line_flux(s, n, u1d, [&](int t) {                 // q(1)=0, q(k)=vort(k-1), :73-76
        return (t > 0) ? W[base + (long long)(t - 1) * stride] : 0.0;
    });
}

template <bool ACCUM, bool VORTEX, bool BERN>
__global__ void __launch_bounds__(256)
k_momentum(const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
           const double* __restrict__ wx, const double* __restrict__ wy, const double* __restrict__ wz,
           const double* __restrict__ ke, const double* __restrict__ b,
           double* __restrict__ dux, double* __restrict__ duy, double* __restrict__ duz,
           double cff, int with_b, Ext e)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= e.nx || j >= e.ny || k >= e.nz) return;
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    const long long bx = c - i, by = c - (long long)j * e.sj, bz = c - (long long)k * e.sk;

    double ax = ACCUM ? dux[c] : 0.0, ay = ACCUM ? duy[c] : 0.0, az = ACCUM ? duz[c] : 0.0;
    if (VORTEX) {
        // order of the three passes "ikj","jik","kji" of vortex_force.py:69-81
        if (i < e.nx - 1) {
            ax = ax + vf_flux(Uy, wz, by, e.sj, 1, j, e.ny);          // pass 1 flip  : +F_y(omega_z)
            ax = ax - vf_flux(Uz, wy, bz, e.sk, 1, k, e.nz);          // pass 3 direc : -F_z(omega_y)
        }
        if (j < e.ny - 1) {
            ay = ay - vf_flux(Ux, wz, bx, 1, e.sj, i, e.nx);          // pass 1 direc : -F_x(omega_z)
            ay = ay + vf_flux(Uz, wx, bz, e.sk, e.sj, k, e.nz);       // pass 2 flip  : +F_z(omega_x)
        }
        if (k < e.nz - 1) {
            az = az - vf_flux(Uy, wx, by, e.sj, e.sk, j, e.ny);       // pass 2 direc : -F_y(omega_x)
            az = az + vf_flux(Ux, wy, bx, 1, e.sk, i, e.nx);          // pass 3 flip  : +F_x(omega_y)
        }
    }
    if (BERN) {
        const double k0 = ke[c];
        if (i < e.nx - 1) ax = ax - (ke[c + 1] - k0);                 // gradke, fortran_bernoulli.f90:20
        if (j < e.ny - 1) ay = ay - (ke[c + e.sj] - k0);
        if (k < e.nz - 1) {
            az = az - (ke[c + e.sk] - k0);
            if (with_b) az = az + cff * (b[c + e.sk] + b[c]);         // gradkeandb, :50-51
        }
    }
    dux[c] = ax; duy[c] = ay; duz[c] = az;
}

__global__ void __launch_bounds__(256)
k_add_laplacian(const double* __restrict__ phi, double* __restrict__ dphi, double cx, double cy, double cz, Ext e)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= e.nx || j >= e.ny || k >= e.nz) return;
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    double acc = dphi[c];
    // fortran_dissipation.f90:2-35 along i, then j, then k (core/viscosity.py:3-9)
    acc = lap_axis(acc, phi, c, 1, i, e.nx, cx);
    acc = lap_axis(acc, phi, c, e.sj, j, e.ny, cy);
    acc = lap_axis(acc, phi, c, e.sk, k, e.nz, cz);
    dphi[c] = acc;
}

inline bool ext_ok(ny_ext e) { return e.nx >= 5 && e.ny >= 5 && e.nz >= 5; }

}  // namespace

extern "C" int ny_upwind(ny_ctx* ctx, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                         double* dtrac, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && trac && Ux && Uy && Uz && dtrac, "null argument");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_RHS_TRACER, ny_stream(stream));
    k_upwind<false><<<g.grid, g.block, 0, ny_stream(stream)>>>(trac, Ux, Uy, Uz, dtrac, 0.0, 0.0, 0.0, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_upwind_diff(ny_ctx* ctx, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                              double* dtrac, double cx, double cy, double cz, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && trac && Ux && Uy && Uz && dtrac, "null argument");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_RHS_TRACER, ny_stream(stream));
    k_upwind<true><<<g.grid, g.block, 0, ny_stream(stream)>>>(trac, Ux, Uy, Uz, dtrac, cx, cy, cz, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_vortex_force(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz,
                               const double* wx, const double* wy, const double* wz,
                               double* dux, double* duy, double* duz, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && Ux && Uy && Uz && wx && wy && wz && dux && duy && duz, "null argument");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_RHS_MOMENTUM, ny_stream(stream));
    k_momentum<true, true, false><<<g.grid, g.block, 0, ny_stream(stream)>>>(
        Ux, Uy, Uz, wx, wy, wz, nullptr, nullptr, dux, duy, duz, 0.0, 0, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_bernoulli(ny_ctx* ctx, const double* ke, const double* b, double* dux, double* duy, double* duz,
                            double dz, int euler, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && ke && dux && duy && duz && (euler || b), "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_RHS_MOMENTUM, ny_stream(stream));
    k_momentum<true, false, true><<<g.grid, g.block, 0, ny_stream(stream)>>>(
        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ke, b, dux, duy, duz, 0.5 * dz, euler ? 0 : 1, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_add_laplacian(ny_ctx* ctx, const double* phi, double* dphi, double cx, double cy, double cz,
                                ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && phi && dphi, "null argument");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    k_add_laplacian<<<g.grid, g.block, 0, ny_stream(stream)>>>(phi, dphi, cx, cy, cz, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_rhs(ny_ctx* ctx, const double* b, const double* Ux, const double* Uy, const double* Uz,
                      const double* wx, const double* wy, const double* wz, const double* ke,
                      double* db, double* dux, double* duy, double* duz,
                      double dz, int flags, ny_ext e, void* stream)
{
    const bool euler = flags & 1, linear = flags & 2;
    NY_REQUIRE(ctx && Ux && Uy && Uz && ke && dux && duy && duz, "null argument");
    NY_REQUIRE(euler || (b && db), "b and db are required unless the Euler flag is set");
    NY_REQUIRE(linear || (wx && wy && wz), "vorticity is required unless the linear flag is set");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    Ext x = make_ext(e);
    if (!euler) {
        ny_prof_scope ps(ctx, NY_PROF_RHS_TRACER, ny_stream(stream));
        k_upwind<false><<<g.grid, g.block, 0, ny_stream(stream)>>>(b, Ux, Uy, Uz, db, 0.0, 0.0, 0.0, x);
        NY_CHECK_LAUNCH(ctx);
    }
    ny_prof_scope ps(ctx, NY_PROF_RHS_MOMENTUM, ny_stream(stream));
    if (linear)
        k_momentum<false, false, true><<<g.grid, g.block, 0, ny_stream(stream)>>>(
            Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, 0.5 * dz, euler ? 0 : 1, x);
    else
        k_momentum<false, true, true><<<g.grid, g.block, 0, ny_stream(stream)>>>(
            Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, 0.5 * dz, euler ? 0 : 1, x);
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}
