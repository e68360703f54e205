// Right-hand-side kernels of the Nyles LES step: tracer advection, vortex force, Bernoulli term.
//
// Replaces core/fortran_upwind.f90:3-87, core/fortran_vortex_force.f90:10-165,
// core/fortran_bernoulli.f90:2-58 as driven by core/tracer.py:44-72, core/vortex_force.py:69-81,
// core/bernoulli.py:14-32.  One canonical (k,j,i) layout; every line sweep of the Fortran is
// evaluated per cell so that all global loads are coalesced along i whatever the sweep axis.
// Accumulation order per output cell equals the order of the reference's passes.
#include "ny_common.cuh"
#include <cstdlib>
#include "ny_weno.cuh"
#include "ny_tma.cuh"

namespace {

struct Ext { int nz, ny, nx; long long sj, sk; };
__host__ __device__ inline Ext make_ext(ny_ext e)
{
    Ext x; x.nz = e.nz; x.ny = e.ny; x.nx = e.nx; x.sj = e.nx; x.sk = (long long)e.nx * e.ny; return x;
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

// ---- tracer advection: dtrac = -div(U trac), every face flux evaluated once -----------------------
// fortran_upwind.f90:66-82 per direction: dtrac(1) -= flux(1); dtrac(s) += flux(s-1) - flux(s).
// The flux through a face is shared by the two cells next to it, so a thread evaluates ONE flux
// per axis and plane (the face on its + side) and obtains the face on its - side
//   x : from the lane to its left (warp shuffle; a warp covers 32 faces and outputs 31 cells),
//   y : from the warp that owns the row below (double-buffered shared memory; warp 0 of a CTA
//       only produces y fluxes for warp 1),
//   z : from its own previous plane (the CTA marches along k; the line values q[k-2..k+3] live
//       in a register queue, one new load per plane).
// DIFF: tracer.py:72-77 interleaves add_laplacian after the upwind pass of each direction.
// Time-scheme update of the tracer applied by the kernel itself (same statements as k_ts in ny_diag.cu,
// core/timescheme.py:131-175).  The kernel reads the tracer with a stencil, so the new value goes to a
// separate array `out` and the caller rotates its buffers: mode 0: off (dtrac is stored), 1: Euler
// start-up (out = s + dt ds), 2: LFAM3 predictor (reads sb), 3: LFAM3 corrector (out = sn + dt ds).
// add != nullptr: a user tendency (the array a forcing object adds to dtrac, core/model_les.py:143-144) is added to the
// finished tendency before the update, as `forcing.add` does after the RHS.
struct TrUpd { int mode; double dt; const double* sb; const double* sn; double* out; const double* add; };

constexpr int UP_NW = 8;             // warps (rows) per CTA: UP_NW-1 output rows of 31 cells
#ifndef UP_MINB
#define UP_MINB 3
#endif
template <bool FAST, bool DIFF>
__global__ void __launch_bounds__(UP_NW * 32, UP_MINB)
k_upwind2(const double* __restrict__ trac, const double* __restrict__ Ux, const double* __restrict__ Uy,
          const double* __restrict__ Uz, double* __restrict__ dtrac, double cx, double cy, double cz, Ext e, int kchunk,
          TrUpd upd)
{
    __shared__ double sFy[2][UP_NW][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = (int)blockIdx.x * 31 - 1 + lane;
    const int j = (int)blockIdx.y * (UP_NW - 1) - 1 + warp;
    const int k0 = (int)blockIdx.z * kchunk, k1 = min(k0 + kchunk, e.nz);
    const bool col = i >= 0 && i < e.nx && j >= 0 && j < e.ny;       // this thread owns a column of cells
    const bool outp = col && lane >= 1 && warp >= 1;
    const long long o = col ? (long long)j * e.sj + i : 0;
    const int ks = k0 > 0 ? k0 - 1 : 0;                               // first plane whose z flux is needed
    // every face of the CTA's tile lies in the interior range of flux1d (Fortran i = 3..n-3) in x and y
    const bool xy_hot = (int)blockIdx.x * 31 - 1 >= 2 && (int)blockIdx.x * 31 + 30 <= e.nx - 4 &&
                        (int)blockIdx.y * (UP_NW - 1) - 1 >= 2 && (int)blockIdx.y * (UP_NW - 1) + UP_NW - 2 <= e.ny - 4;

    // register queue zq[d+2] = trac[k+d], d = -2..3
    double zq[6];
#pragma unroll
    for (int d = -2; d <= 3; d++) {
        const int kk = ks + d;
        zq[d + 2] = (col && kk >= 0 && kk < e.nz) ? trac[(long long)kk * e.sk + o] : 0.0;
    }
    double Fz_prev = 0.0;
    double ux = 0.0, uy = 0.0, uz = 0.0;
    if (col) { const long long c = (long long)ks * e.sk + o; ux = Ux[c]; uy = Uy[c]; uz = Uz[c]; }
    for (int k = ks; k < k1; k++) {
        const long long c = (long long)k * e.sk + o;
        const bool full = k >= k0;                                    // k0-1 only supplies the z flux below the chunk
        // prefetch the next plane: face velocities and the new end of the z queue
        double nux = 0.0, nuy = 0.0, nuz = 0.0, nq = 0.0;
        if (col && k + 1 < k1) { nux = Ux[c + e.sk]; nuy = Uy[c + e.sk]; nuz = Uz[c + e.sk]; }
        if (col && k + 4 < e.nz) nq = trac[c + 4 * e.sk];
        double told = 0.0;                                            // the other time level the update reads
        if (upd.mode >= 2 && outp && full) told = upd.mode == 2 ? upd.sb[c] : upd.sn[c];
        double Fx = 0.0, Fy = 0.0, Fz = 0.0;
        if (xy_hot && full && k >= 2 && k <= e.nz - 4) {             // CTA-uniform: branch-free interior path
            Fz = nyw::hot_flux<FAST>(uz, [&](int d) { return zq[d + 2]; });
            Fy = nyw::hot_flux<FAST>(uy, [&](int d) { return trac[c + (long long)d * e.sj]; });
            if (warp >= 1) Fx = nyw::hot_flux<FAST>(ux, [&](int d) { return trac[c + d]; });
        } else if (col) {
            Fz = nyw::line_flux<FAST>(k, e.nz, uz, [&](int d) { return zq[d + 2]; });
            if (full) {
                Fy = nyw::line_flux<FAST>(j, e.ny, uy, [&](int d) { return trac[c + (long long)d * e.sj]; });
                if (warp >= 1) Fx = nyw::line_flux<FAST>(i, e.nx, ux, [&](int d) { return trac[c + d]; });
            }
        }
        if (full) {
            const int buf = k & 1;
            sFy[buf][warp][lane] = Fy;
            const double Fxm = __shfl_up_sync(0xffffffffu, Fx, 1);
            __syncthreads();
            if (outp) {
                const double Fym = sFy[buf][warp - 1][lane];
                double acc = 0.0;                                     // tracer.py:70-71
                acc = (i == 0) ? acc - Fx : acc + Fxm - Fx;
                if (DIFF) acc = lap_axis(acc, trac, c, 1, i, e.nx, cx);
                acc = (j == 0) ? acc - Fy : acc + Fym - Fy;
                if (DIFF) acc = lap_axis(acc, trac, c, e.sj, j, e.ny, cy);
                acc = (k == 0) ? acc - Fz : acc + Fz_prev - Fz;
                if (DIFF) acc = lap_axis(acc, trac, c, e.sk, k, e.nz, cz);
                if (upd.add) acc = acc + upd.add[c];                  // model_les.py:143-144
                if (upd.mode == 0) dtrac[c] = acc;
                else if (upd.mode == 2) {                             // timescheme.py:144-162
                    const double v = zq[2], lf = told + (2. * upd.dt) * acc;
                    upd.out[c] = (1. / 12.) * (5. * lf + 8. * v - told);
                } else if (upd.mode == 3) upd.out[c] = told + upd.dt * acc;      // timescheme.py:170-175
                else upd.out[c] = zq[2] + upd.dt * acc;               // timescheme.py:131-139
            }
        }
        Fz_prev = Fz;
#pragma unroll
        for (int d = 0; d < 5; d++) zq[d] = zq[d + 1];
        zq[5] = nq;
        ux = nux; uy = nuy; uz = nuz;
    }
}

// ---- vortex force: flux of one sweep for the cell at line position s ------------------------
// US: velocity along the sweep axis; tstride: stride of the target component's own axis (the
// averaging axis); W: the vorticity component normal to (sweep, target).
// INTERIOR: the caller guarantees 3 <= s <= n-4, so no closure of flux1d and no q(1)=0 applies.
template <bool FAST, bool INTERIOR>
__device__ __forceinline__ double vf_flux(const double* __restrict__ US, const double* __restrict__ W,
                                          long long cs, long long stride, long long tstride, int s, int n)
{
    // UU(s) = 0.5*(U + U'), u1d = 0.5*(UU(s-1) + UU(s)) (fortran_vortex_force.f90:68-72): the halvings are
    // exact, so the two rounded sums added and scaled once give the same bits with two multiplies fewer
    const double s1 = US[cs] + US[cs + tstride];
    const double s0 = (INTERIOR || s > 0) ? US[cs - stride] + US[cs - stride + tstride] : 0.0;
    const double u1d = 0.25 * (s0 + s1);
    auto q = [&](int d) {                                         // q(1)=0, q(k)=vort(k-1), :73-76
        return (INTERIOR || s + d > 0) ? W[cs + (long long)(d - 1) * stride] : 0.0;
    };
    if (INTERIOR) return nyw::hot_flux<FAST>(u1d, q);
    return nyw::line_flux<FAST>(s, n, u1d, q);
}

// LFAM3 / Euler-forward update of the three velocity components applied by the momentum kernel itself
// (core/timescheme.py:131-175; same statements as k_ts in ny_diag.cu): the tendencies never go to memory.
// mode 0: off (du is stored), 1: Euler start-up, 2: predictor, 3: corrector.  The kernel reads U, vor, ke
// and b only, so updating u in place is race-free.
// out[] != nullptr: rotating form -- only the new value is written, to out[] (the caller rotates its buffers:
// the old u array becomes un, see ny_rhs_step); s, sb, sn are then read-only.
struct TsUpd { int mode; double dt; double *s[3], *sb[3], *sn[3], *out[3]; const double* add[3]; };

__device__ __forceinline__ void ts_apply(const TsUpd& u, int comp, long long c, double ds)
{
    double* __restrict__ s = u.s[comp];
    if (u.add[comp]) ds = ds + u.add[comp][c];       // user tendency added after the RHS (model_les.py:143-144)
    if (u.out[comp]) {
        double* __restrict__ o = u.out[comp];
        if (u.mode == 2) {
            const double v = s[c], vb = u.sb[comp][c];
            const double lf = vb + (2. * u.dt) * ds;
            o[c] = (1. / 12.) * (5. * lf + 8. * v - vb);
        } else if (u.mode == 3) o[c] = u.sn[comp][c] + u.dt * ds;
        else o[c] = s[c] + u.dt * ds;
        return;
    }
    if (u.mode == 2) {                               // timescheme.py:144-162
        const double v = s[c], vb = u.sb[comp][c];
        const double lf = vb + (2. * u.dt) * ds;
        s[c] = (1. / 12.) * (5. * lf + 8. * v - vb);
        u.sn[comp][c] = v; u.sb[comp][c] = v;
    } else if (u.mode == 3) {                        // timescheme.py:170-175
        s[c] = u.sn[comp][c] + u.dt * ds;
    } else {                                         // timescheme.py:131-139
        const double v = s[c];
        u.sn[comp][c] = v; u.sb[comp][c] = v;
        s[c] = v + u.dt * ds;
    }
}

template <bool FAST, bool INTERIOR, bool ACCUM, bool VORTEX, bool BERN>
__device__ __forceinline__ void momentum_cell(
    const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
    const double* __restrict__ wx, const double* __restrict__ wy, const double* __restrict__ wz,
    const double* __restrict__ ke, const double* __restrict__ b,
    double* __restrict__ dux, double* __restrict__ duy, double* __restrict__ duz,
    double cff, int with_b, const Ext& e, int i, int j, int k, const TsUpd& upd)
{
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    double ax = ACCUM ? dux[c] : 0.0, ay = ACCUM ? duy[c] : 0.0, az = ACCUM ? duz[c] : 0.0;
    if (VORTEX) {
        // order of the three passes "ikj","jik","kji" of vortex_force.py:69-81
        if (INTERIOR || i < e.nx - 1) {
            ax = ax + vf_flux<FAST, INTERIOR>(Uy, wz, c, e.sj, 1, j, e.ny);          // pass 1 flip  : +F_y(omega_z)
            ax = ax - vf_flux<FAST, INTERIOR>(Uz, wy, c, e.sk, 1, k, e.nz);          // pass 3 direc : -F_z(omega_y)
        }
        if (INTERIOR || j < e.ny - 1) {
            ay = ay - vf_flux<FAST, INTERIOR>(Ux, wz, c, 1, e.sj, i, e.nx);          // pass 1 direc : -F_x(omega_z)
            ay = ay + vf_flux<FAST, INTERIOR>(Uz, wx, c, e.sk, e.sj, k, e.nz);       // pass 2 flip  : +F_z(omega_x)
        }
        if (INTERIOR || k < e.nz - 1) {
            az = az - vf_flux<FAST, INTERIOR>(Uy, wx, c, e.sj, e.sk, j, e.ny);       // pass 2 direc : -F_y(omega_x)
            az = az + vf_flux<FAST, INTERIOR>(Ux, wy, c, 1, e.sk, i, e.nx);          // pass 3 flip  : +F_x(omega_y)
        }
    }
    if (BERN) {
        const double k0 = ke[c];
        if (INTERIOR || i < e.nx - 1) ax = ax - (ke[c + 1] - k0);                 // gradke, fortran_bernoulli.f90:20
        if (INTERIOR || j < e.ny - 1) ay = ay - (ke[c + e.sj] - k0);
        if (INTERIOR || k < e.nz - 1) {
            az = az - (ke[c + e.sk] - k0);
            if (with_b) az = az + cff * (b[c + e.sk] + b[c]);         // gradkeandb, :50-51
        }
    }
    if (upd.mode == 0) { dux[c] = ax; duy[c] = ay; duz[c] = az; }
    else { ts_apply(upd, 0, c, ax); ts_apply(upd, 1, c, ay); ts_apply(upd, 2, c, az); }
}

// occupancy beats registers here: the kernel waits on loads and on the fp64 division chains, and more
// resident warps hide both (measured at 512^3, strict arithmetic: 3 CTAs/SM 8.4 ms, 4: 7.1, 5: 6.8, 6: 6.8,
// 8: 7.0 -- a few spilled registers cost less than the lost warps)
#ifndef MOM_MINB
#define MOM_MINB 5
#endif
// The boxes of cells [i0,i1) x [j0,j1) x [k0,k1) that ONE launch of k_momentum covers: the whole array, or the six
// slabs of the 3-cell frame around the cells that k_mom3 computes.  Each box has its own block shape (bx, by, bz),
// 256 threads, and its own grid (gx, gy, .) of CTAs; the CTAs of all boxes are enumerated along blockIdx.x.
struct CellBox { int i0, i1, j0, j1, k0, k1, bx, by, bz, gx, gy, start; };
struct CellBoxes { int n, total; CellBox box[6]; };

template <bool FAST, bool ACCUM, bool VORTEX, bool BERN>
__global__ void __launch_bounds__(256, MOM_MINB)
k_momentum(const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
           const double* __restrict__ wx, const double* __restrict__ wy, const double* __restrict__ wz,
           const double* __restrict__ ke, const double* __restrict__ b,
           double* __restrict__ dux, double* __restrict__ duy, double* __restrict__ duz,
           double cff, int with_b, Ext e, TsUpd upd, const __grid_constant__ CellBoxes boxes)
{
    int nb = 0;
    while (nb + 1 < boxes.n && (int)blockIdx.x >= boxes.box[nb + 1].start) nb++;
    const CellBox& cb = boxes.box[nb];
    const int local = (int)blockIdx.x - cb.start;
    const int cx = local % cb.gx, cy = (local / cb.gx) % cb.gy, cz = local / (cb.gx * cb.gy);
    const int tid = threadIdx.x, tx = tid % cb.bx, ty = (tid / cb.bx) % cb.by, tz = tid / (cb.bx * cb.by);
    const int i0 = cb.i0 + cx * cb.bx, j0 = cb.j0 + cy * cb.by, k0 = cb.k0 + cz * cb.bz;
    const int i = i0 + tx, j = j0 + ty, k = k0 + tz;
    // CTA-uniform: every cell of the tile has all six sweeps in the interior range of flux1d
    const bool interior = VORTEX && i0 >= 3 && i0 + cb.bx - 1 <= e.nx - 4 && j0 >= 3 && j0 + cb.by - 1 <= e.ny - 4 &&
                          k0 >= 3 && k0 + cb.bz - 1 <= e.nz - 4 && i0 + cb.bx <= cb.i1 && j0 + cb.by <= cb.j1 &&
                          k0 + cb.bz <= cb.k1;
    if (interior) {
        momentum_cell<FAST, VORTEX, ACCUM, VORTEX, BERN>(Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, cff, with_b, e, i, j, k, upd);
        return;
    }
    if (i >= cb.i1 || j >= cb.j1 || k >= cb.k1) return;
    momentum_cell<FAST, false, ACCUM, VORTEX, BERN>(Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, cff, with_b, e, i, j, k, upd);
}

#include "ny_mom3.cuh"

// ================================================================================================
//  The linear (non-WENO) upwind branch of the reference: `linear = .true.` in fortran_upwind.f90:31 and
//  fortran_vortex_force.f90:28,108 (a local flag, .false. as shipped, so dormant there), with the interpolations
//  of core/interpolate_tracer.f90 / core/interpolate.f90, orders 1..5.  One thread per cell, every value the
//  cell needs recomputed from its line neighbours; coefficients are the REAL(4) constant expressions promoted to
//  double.  Not a hot path: written for parity, not for speed.
// ================================================================================================
namespace lin {
constexpr double C1 = (double)(-(1.0f / 6.0f)), C2 = (double)(5.0f / 6.0f), C3 = (double)(2.0f / 6.0f);
constexpr double E1 = (double)(-(1.0f / 12.0f)), E2 = (double)(7.0f / 12.0f);
constexpr double B1 = (double)(2.0f / 60.0f), B2 = (double)(-(13.0f / 60.0f)), B3 = (double)(47.0f / 60.0f),
                 B4 = (double)(27.0f / 60.0f), B5 = (double)(-(3.0f / 60.0f));

// v(i): value at 1-based line position i.  P = qp flavour (coefficients c1 c2 c3 / b1..b5), M = mirrored.
template <class V> __device__ __forceinline__ double third(bool plus, V v, int i)
{
    return plus ? C1 * v(i - 1) + C2 * v(i) + C3 * v(i + 1) : C3 * v(i - 1) + C2 * v(i) + C1 * v(i + 1);
}
template <class V> __device__ __forceinline__ double fifth(bool plus, V v, int i)
{
    return plus ? B1 * v(i - 2) + B2 * v(i - 1) + B3 * v(i) + B4 * v(i + 1) + B5 * v(i + 2)
                : B5 * v(i - 2) + B4 * v(i - 1) + B3 * v(i) + B2 * v(i + 1) + B1 * v(i + 2);
}

// interpolate_tracer.f90:45-110: qp(i) (plus) or qm(i) at 1 <= i <= n for the odd orders, the centred qp(i)
// (1 <= i <= n-1) for the even ones
template <class V> __device__ __forceinline__ double interp_tr(int order, int n, bool plus, V v, int i)
{
    if (order == 5) {
        if (i == 1 || i == n) return v(i);
        if (i == 2 || i == n - 1) return third(plus, v, i);
        return fifth(plus, v, i);
    }
    if (order == 3) return (i == 1 || i == n) ? v(i) : third(plus, v, i);
    if (order == 1) return v(i);
    if (order == 2) return 0.5 * (v(i) + v(i + 1));
    /* order 4 */
    if (i == 1) return E2 * (v(i) + v(i + 1)) + E1 * (v(i + 2));
    if (i == n - 1) return E2 * (v(i) + v(i + 1)) + E1 * (v(i - 1));
    return E2 * (v(i) + v(i + 1)) + E1 * (v(i - 1) + v(i + 2));
}

// interpolate.f90:34-110 (the flavour of the vortex force): qp(i) for 0 <= i <= n-1 (plus) or qm(i) for 1 <= i <= n
// for the odd orders, qm(i) (1 <= i <= n) for the even ones.  Later assignments of the Fortran win.
template <class V> __device__ __forceinline__ double interp_vf(int order, int n, bool plus, V v, int i)
{
    if (order == 5) {
        if (i == 0) return 0.0;
        if (i == n || i == n - 1 || i == 1) return v(i);
        if (i == n - 2 || i == 2) return third(plus, v, i);
        return fifth(plus, v, i);
    }
    if (order == 3) {
        if (i == 0) return 0.0;
        if (i == n || i == n - 1 || i == 1) return v(i);
        return third(plus, v, i);
    }
    if (order == 1) return i == 0 ? 0.0 : v(i);
    if (order == 2) return i == 1 ? 0.5 * v(1) : 0.5 * (v(i - 1) + v(i));
    /* order 4 */
    if (i == 1) return E2 * (v(i)) + E1 * (v(i + 1));
    if (i == 2) return E2 * (v(i - 1) + v(i)) + E1 * (v(i + 1));
    if (i == n) return E2 * (v(i - 1) + v(i)) + E1 * (v(i - 2));
    return E2 * (v(i - 1) + v(i)) + E1 * (v(i - 2) + v(i + 1));
}

// flux through the face between line cells s and s+1 (0-based s), fortran_upwind.f90:43-63; the last face is closed
template <class Q, class U> __device__ __forceinline__ double tracer_flux(int order, int n, int s, Q q, U u)
{
    if (s < 0 || s >= n - 1) return 0.0;
    const int i = s + 1;                                  // Fortran position
    auto v = [&](int p) { return q(p - 1); };
    const double ui = u(s);
    if ((order & 1) == 0) return ui * interp_tr(order, n, true, v, i);
    const double UU = fabs(ui), up = 0.5 * (ui + UU), um = 0.5 * (ui - UU);
    return up * interp_tr(order, n, true, v, i) + um * interp_tr(order, n, false, v, i + 1);
}
}  // namespace lin

__global__ void __launch_bounds__(256)
k_upwind_linear(const double* __restrict__ trac, const double* __restrict__ Ux, const double* __restrict__ Uy,
                const double* __restrict__ Uz, double* __restrict__ dtrac, int order, Ext e)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= e.nx || j >= e.ny || k >= e.nz) return;
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    double acc = 0.0;                                     // tracer.py:70-71
    {
        auto q = [&](int p) { return trac[c + (p - i)]; };
        auto u = [&](int p) { return Ux[c + (p - i)]; };
        acc = acc + lin::tracer_flux(order, e.nx, i - 1, q, u) - lin::tracer_flux(order, e.nx, i, q, u);
    }
    {
        auto q = [&](int p) { return trac[c + (long long)(p - j) * e.sj]; };
        auto u = [&](int p) { return Uy[c + (long long)(p - j) * e.sj]; };
        acc = acc + lin::tracer_flux(order, e.ny, j - 1, q, u) - lin::tracer_flux(order, e.ny, j, q, u);
    }
    {
        auto q = [&](int p) { return trac[c + (long long)(p - k) * e.sk]; };
        auto u = [&](int p) { return Uz[c + (long long)(p - k) * e.sk]; };
        acc = acc + lin::tracer_flux(order, e.nz, k - 1, q, u) - lin::tracer_flux(order, e.nz, k, q, u);
    }
    dtrac[c] = acc;
}

// one sweep of the linear vortex force for the cell at 0-based line position s: res is updated in place with the
// statements of fortran_vortex_force.f90:52-63 (direc, sign = -1) or :131-142 (flip, sign = +1)
__device__ __forceinline__ double vf_linear(double res, int order, int sign, const double* __restrict__ US,
                                            const double* __restrict__ W, long long cs, long long stride, long long tstride,
                                            int s, int n)
{
    auto v = [&](int p) { return W[cs + (long long)(p - 1 - s) * stride]; };       // vU(p) = vort at line position p
    if ((order & 1) == 0) {
        const double qm = lin::interp_vf(order, n, false, v, s + 1);
        return sign < 0 ? res - qm : res + qm;
    }
    const double s1 = 0.5 * (US[cs] + US[cs + tstride]);
    const double s0 = s > 0 ? 0.5 * (US[cs - stride] + US[cs - stride + tstride]) : 0.0;
    const double Ui = 0.5 * (s0 + s1), UU = fabs(Ui), up = 0.5 * (Ui + UU), um = 0.5 * (Ui - UU);
    const double qp = lin::interp_vf(order, n, true, v, s), qm = lin::interp_vf(order, n, false, v, s + 1);
    return sign < 0 ? res - qp * up - qm * um : res + qp * up + qm * um;
}

__global__ void __launch_bounds__(256)
k_vortex_force_linear(const double* __restrict__ Ux, const double* __restrict__ Uy, const double* __restrict__ Uz,
                      const double* __restrict__ wx, const double* __restrict__ wy, const double* __restrict__ wz,
                      double* __restrict__ dux, double* __restrict__ duy, double* __restrict__ duz, int order, Ext e)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= e.nx || j >= e.ny || k >= e.nz) return;
    const long long c = (long long)k * e.sk + (long long)j * e.sj + i;
    double ax = dux[c], ay = duy[c], az = duz[c];
    // the three passes "ikj", "jik", "kji" of vortex_force.py:69-81, as in momentum_cell
    if (i < e.nx - 1) {
        ax = vf_linear(ax, order, +1, Uy, wz, c, e.sj, 1, j, e.ny);
        ax = vf_linear(ax, order, -1, Uz, wy, c, e.sk, 1, k, e.nz);
    }
    if (j < e.ny - 1) {
        ay = vf_linear(ay, order, -1, Ux, wz, c, 1, e.sj, i, e.nx);
        ay = vf_linear(ay, order, +1, Uz, wx, c, e.sk, e.sj, k, e.nz);
    }
    if (k < e.nz - 1) {
        az = vf_linear(az, order, -1, Uy, wx, c, e.sj, e.sk, j, e.ny);
        az = vf_linear(az, order, +1, Ux, wy, c, 1, e.sk, i, e.nx);
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

__global__ void __launch_bounds__(256)
k_debug_weno5(const double* __restrict__ q, double* __restrict__ out, long long n, int fast)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double a = q[t], b = q[n + t], c = q[2 * n + t], d = q[3 * n + t], e = q[4 * n + t];
    out[t] = fast ? nyw::weno5<true>(a, b, c, d, e) : nyw::weno5<false>(a, b, c, d, e);
}

// fp64 pipe peak: eight independent DFMA chains per thread, no memory traffic
__global__ void __launch_bounds__(512)
k_fp64_peak(double* __restrict__ out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll 4
        for (int u = 0; u < 4; u++) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;               // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256)
k_debug_weno3(const double* __restrict__ q, double* __restrict__ out, long long n)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    out[t] = nyw::weno3(q[t], q[n + t], q[2 * n + t]);
}

__global__ void __launch_bounds__(256)
k_debug_div(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, long long n,
            unsigned long long* __restrict__ mismatch)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double q = nyw::div_inrange(a[t], b[t]), ref = a[t] / b[t];
    out[t] = q;
    if (__double_as_longlong(q) != __double_as_longlong(ref)) atomicAdd(mismatch, 1ull);
}

inline bool ext_ok(ny_ext e) { return e.nx >= 5 && e.ny >= 5 && e.nz >= 5; }

}  // namespace

// launch helpers ----------------------------------------------------------------------------------
static int launch_upwind(ny_ctx* ctx, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                         double* dtrac, bool diff, double cx, double cy, double cz, ny_ext e, cudaStream_t st,
                         const TrUpd* updp = nullptr)
{
    TrUpd upd;
    memset(&upd, 0, sizeof(upd));
    if (updp) upd = *updp;
    Ext x = make_ext(e);
    ny_prof_scope ps(ctx, NY_PROF_RHS_TRACER, st);
    // (A plane-marching TMA variant of this kernel -- helper warp for the extra faces, pairwise mbarriers, cells
    // finished one iteration late -- was built, verified bit-identical and measured at 4.07 ms against the 4.05 ms of
    // this one per 512^3 launch: both sit at 1965 MHz and 960-990 W, next to the 1 kW cap, and the exchange of y fluxes
    // between the row warps costs what the staging saves.  It was dropped; profiles/r2_d_* keep the record.)
    const int gx = (e.nx + 30) / 31, gy = (e.ny + UP_NW - 2) / (UP_NW - 1);
    // split k so that the launch has ~32 CTAs per SM; each chunk pays one extra plane of z fluxes
    long long want = ((long long)ctx->num_sms * 32 + (long long)gx * gy - 1) / ((long long)gx * gy);
    int nchunk = (int)(want < 1 ? 1 : want);
    int kchunk = (e.nz + nchunk - 1) / nchunk;
    if (kchunk < 16) kchunk = e.nz < 16 ? e.nz : 16;
    dim3 grid(gx, gy, (e.nz + kchunk - 1) / kchunk);
    if (ctx->fast_arith) {
        if (diff) k_upwind2<true, true><<<grid, UP_NW * 32, 0, st>>>(trac, Ux, Uy, Uz, dtrac, cx, cy, cz, x, kchunk, upd);
        else k_upwind2<true, false><<<grid, UP_NW * 32, 0, st>>>(trac, Ux, Uy, Uz, dtrac, cx, cy, cz, x, kchunk, upd);
    } else {
        if (diff) k_upwind2<false, true><<<grid, UP_NW * 32, 0, st>>>(trac, Ux, Uy, Uz, dtrac, cx, cy, cz, x, kchunk, upd);
        else k_upwind2<false, false><<<grid, UP_NW * 32, 0, st>>>(trac, Ux, Uy, Uz, dtrac, cx, cy, cz, x, kchunk, upd);
    }
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

template <bool ACCUM, bool VORTEX, bool BERN>
static int launch_momentum(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz,
                           const double* wx, const double* wy, const double* wz, const double* ke, const double* b,
                           double* dux, double* duy, double* duz, double cff, int with_b, ny_ext e, cudaStream_t st,
                           const TsUpd* updp = nullptr)
{
    TsUpd upd;
    memset(&upd, 0, sizeof(upd));
    if (updp) upd = *updp;
    ny_prof_scope ps(ctx, NY_PROF_RHS_MOMENTUM, st);
    // one launch of the cell-parallel kernel over a list of boxes, each with a block shape that suits it
    CellBoxes boxes;
    boxes.n = 0; boxes.total = 0;
    auto add_box = [&](int i0, int i1, int j0, int j1, int k0, int k1, int bx, int by, int bz) {
        if (i0 >= i1 || j0 >= j1 || k0 >= k1) return;
        CellBox& cb = boxes.box[boxes.n++];
        cb.i0 = i0; cb.i1 = i1; cb.j0 = j0; cb.j1 = j1; cb.k0 = k0; cb.k1 = k1; cb.bx = bx; cb.by = by; cb.bz = bz;
        cb.gx = (i1 - i0 + bx - 1) / bx; cb.gy = (j1 - j0 + by - 1) / by;
        cb.start = boxes.total;
        boxes.total += cb.gx * cb.gy * ((k1 - k0 + bz - 1) / bz);
    };
    auto launch_boxes = [&]() -> int {
        if (boxes.total == 0) return NY_OK;
        if (ctx->fast_arith && VORTEX)
            k_momentum<true, ACCUM, VORTEX, BERN><<<boxes.total, 256, 0, st>>>(Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, cff,
                                                                              with_b, make_ext(e), upd, boxes);
        else
            k_momentum<false, ACCUM, VORTEX, BERN><<<boxes.total, 256, 0, st>>>(Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, cff,
                                                                               with_b, make_ext(e), upd, boxes);
        NY_CHECK_LAUNCH(ctx);
        return NY_OK;
    };
    // The cells whose six sweeps are all in the interior range of flux1d go to the plane-marching TMA kernel
    // (ny_mom3.cuh); k_momentum keeps the 3-cell frame around them.  TMA needs an even row length and 16-byte
    // aligned arrays; small grids stay on one launch (mom_variant: 0 = by size, 1 = always k_momentum, 2 = k_mom3
    // wherever it is legal).
    const bool fused = VORTEX && BERN && !ACCUM;
    const long long cells = (long long)e.nx * e.ny * e.nz;
    bool use3 = fused && ctx->mom_variant != 1 && (ctx->mom_variant == 2 || cells >= (1LL << 18)) && (e.nx & 1) == 0 &&
                e.nx >= 7 && e.ny >= 7 && e.nz >= 7;
    if (use3) {
        const void* ptrs[] = {Ux, Uy, Uz, wx, wy, wz, ke, with_b ? b : ke};
        for (const void* p : ptrs) use3 = use3 && (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    }
    if (!use3) {
        add_box(0, e.nx, 0, e.ny, 0, e.nz, 32, 4, 2);
        return launch_boxes();
    }
    {
        using namespace m3;
        Maps tm;
        int r = ny_tma_encode_3d(&tm.wz, wz, e.nx, e.ny, e.nz, PW, TY + 6, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.wx, wx, e.nx, e.ny, e.nz, PN, TY + 6, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.wy, wy, e.nx, e.ny, e.nz, PW, TY, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.Ux, Ux, e.nx, e.ny, e.nz, PN, TY + 1, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.Uy, Uy, e.nx, e.ny, e.nz, PN, TY + 1, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.Uz, Uz, e.nx, e.ny, e.nz, PN, TY + 1, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.ke, ke, e.nx, e.ny, e.nz, PN, TY + 1, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&tm.b, with_b ? b : ke, e.nx, e.ny, e.nz, PN, TY, 1);
        if (r != NY_OK) return r;
        const int ni = e.nx - 6, nj = e.ny - 6, nk = e.nz - 6;
        const int gx = (ni + TX - 1) / TX, gy = (nj + TY - 1) / TY;
        // Chunks of planes.  A chunk re-reads two planes and refills its queues, so chunks should not be short; but the
        // last CTAs of an SM run below full occupancy for about half a chunk, so they should not be long either
        // (measured at 512^3: 169 planes per chunk 6.2 ms, 127 planes 5.5 ms; profiles/r2_b_*).
        int kchunk = 32;
        { const char* v = getenv("NY_MOM3_KCHUNK"); if (v && atoi(v) > 0) kchunk = atoi(v); }
        if (kchunk > nk) kchunk = nk;
        dim3 grid(gx, gy, (nk + kchunk - 1) / kchunk);
        static bool attr_set[2] = {false, false};
        const int fa = ctx->fast_arith ? 1 : 0;
        if (!attr_set[fa]) {
            cudaError_t ce = fa ? cudaFuncSetAttribute(k_mom3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)
                                : cudaFuncSetAttribute(k_mom3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
            if (ce != cudaSuccess) { ny_set_error("k_mom3: cannot reserve %d bytes of shared memory", SMEM); return NY_ERR_CUDA; }
            attr_set[fa] = true;
        }
        if (fa) k_mom3<true><<<grid, TX * TY, SMEM, st>>>(tm, Uz, wx, wy, dux, duy, duz, cff, with_b, make_ext(e), kchunk, upd);
        else k_mom3<false><<<grid, TX * TY, SMEM, st>>>(tm, Uz, wx, wy, dux, duy, duz, cff, with_b, make_ext(e), kchunk, upd);
        NY_CHECK_LAUNCH(ctx);
    }
    // the frame in one launch: two z slabs over the whole plane, two y slabs between them, two x slabs between those
    const int x1 = e.nx - 3, y1 = e.ny - 3, z1 = e.nz - 3;
    add_box(0, e.nx, 0, e.ny, 0, 3, 32, 8, 1);
    add_box(0, e.nx, 0, e.ny, z1, e.nz, 32, 8, 1);
    add_box(0, e.nx, 0, 3, 3, z1, 32, 1, 8);
    add_box(0, e.nx, y1, e.ny, 3, z1, 32, 1, 8);
    add_box(0, 3, 3, y1, 3, z1, 4, 8, 8);
    add_box(x1, e.nx, 3, y1, 3, z1, 4, 8, 8);
    return launch_boxes();
}

extern "C" int ny_set_arith(ny_ctx* ctx, int fast)
{
    NY_REQUIRE(ctx, "null argument");
    ctx->fast_arith = fast ? 1 : 0;
    return NY_OK;
}
extern "C" int ny_get_arith(ny_ctx* ctx) { return ctx ? ctx->fast_arith : 0; }

extern "C" int ny_set_momentum_variant(ny_ctx* ctx, int variant)
{
    NY_REQUIRE(ctx && variant >= 0 && variant <= 2, "bad argument");
    ctx->mom_variant = variant;
    return NY_OK;
}

extern "C" int ny_debug_weno5(ny_ctx* ctx, const double* q, double* out, long long n, void* stream)
{
    NY_REQUIRE(ctx && q && out && n > 0, "bad argument");
    k_debug_weno5<<<(unsigned)((n + 255) / 256), 256, 0, ny_stream(stream)>>>(q, out, n, ctx->fast_arith);
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

// Sustained rate of the fp64 pipe under this board's power cap: runs DFMA-only launches back to back for about
// `seconds` and returns fp64 (warp-wide) instructions issued per second summed over the device, with the elapsed
// time measured by CUDA events on `stream`.  The denominator of the WENO kernels' pipe fraction (bench.py).
extern "C" int ny_debug_fp64_peak(ny_ctx* ctx, double seconds, double* dp_warp_instr_per_s, void* stream)
{
    NY_REQUIRE(ctx && dp_warp_instr_per_s && seconds > 0.0 && seconds <= 10.0, "bad argument");
    cudaStream_t st = ny_stream(stream);
    const int iters = 1 << 14, blocks = 4 * (ctx->num_sms > 0 ? ctx->num_sms : 148);
    cudaEvent_t e0, e1;
    NY_CUDA(cudaEventCreate(&e0));
    NY_CUDA(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, 512, 0, st>>>(ctx->d_scratch, iters, 1.0);           // warm-up, also times one launch
    NY_CUDA(cudaEventRecord(e0, st));
    k_fp64_peak<<<blocks, 512, 0, st>>>(ctx->d_scratch, iters, 1.0);
    NY_CUDA(cudaEventRecord(e1, st));
    NY_CUDA(cudaEventSynchronize(e1));
    float ms1 = 0.f;
    NY_CUDA(cudaEventElapsedTime(&ms1, e0, e1));
    int n = (int)(seconds * 1e3 / (ms1 > 0.01f ? ms1 : 0.01f));
    n = n < 1 ? 1 : (n > 2000 ? 2000 : n);
    NY_CUDA(cudaEventRecord(e0, st));
    for (int r = 0; r < n; r++) k_fp64_peak<<<blocks, 512, 0, st>>>(ctx->d_scratch, iters, 1.0);
    NY_CUDA(cudaEventRecord(e1, st));
    NY_CUDA(cudaEventSynchronize(e1));
    NY_CHECK_LAUNCH(ctx);
    float ms = 0.f;
    NY_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // per launch: blocks * 16 warps * iters * 32 DFMA
    *dp_warp_instr_per_s = (double)n * blocks * 16.0 * iters * 32.0 / (ms * 1e-3);
    return NY_OK;
}

extern "C" int ny_debug_weno3(ny_ctx* ctx, const double* q, double* out, long long n, void* stream)
{
    NY_REQUIRE(ctx && q && out && n > 0, "bad argument");
    k_debug_weno3<<<(unsigned)((n + 255) / 256), 256, 0, ny_stream(stream)>>>(q, out, n);
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

extern "C" int ny_debug_div(ny_ctx* ctx, const double* a, const double* b, double* out, long long n,
                            long long* mismatch_host, void* stream)
{
    NY_REQUIRE(ctx && a && b && out && mismatch_host && n > 0, "bad argument");
    cudaStream_t st = ny_stream(stream);
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(ctx->d_scratch);
    NY_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));
    k_debug_div<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, out, n, d_cnt);
    NY_CHECK_LAUNCH(ctx);
    NY_CUDA(cudaMemcpyAsync(mismatch_host, d_cnt, sizeof(long long), cudaMemcpyDeviceToHost, st));
    NY_CUDA(cudaStreamSynchronize(st));
    return NY_OK;
}

extern "C" int ny_upwind(ny_ctx* ctx, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                         double* dtrac, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && trac && Ux && Uy && Uz && dtrac, "null argument");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    return launch_upwind(ctx, trac, Ux, Uy, Uz, dtrac, false, 0.0, 0.0, 0.0, e, ny_stream(stream));
}

extern "C" int ny_upwind_diff(ny_ctx* ctx, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                              double* dtrac, double cx, double cy, double cz, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && trac && Ux && Uy && Uz && dtrac, "null argument");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    return launch_upwind(ctx, trac, Ux, Uy, Uz, dtrac, true, cx, cy, cz, e, ny_stream(stream));
}

extern "C" int ny_vortex_force(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz,
                               const double* wx, const double* wy, const double* wz,
                               double* dux, double* duy, double* duz, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && Ux && Uy && Uz && wx && wy && wz && dux && duy && duz, "null argument");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    return launch_momentum<true, true, false>(ctx, Ux, Uy, Uz, wx, wy, wz, nullptr, nullptr, dux, duy, duz, 0.0, 0, e,
                                              ny_stream(stream));
}

extern "C" int ny_bernoulli(ny_ctx* ctx, const double* ke, const double* b, double* dux, double* duy, double* duz,
                            double dz, int euler, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && ke && dux && duy && duz && (euler || b), "null argument");
    return launch_momentum<true, false, true>(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ke, b,
                                              dux, duy, duz, 0.5 * dz, euler ? 0 : 1, e, ny_stream(stream));
}

// fortran_upwind.upwind with the Fortran's `linear` flag set, for the three directions (tracer.py:44-72):
// dtrac = -div(U trac) with the linear interpolation of the given order (1..5)
extern "C" int ny_upwind_linear(ny_ctx* ctx, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                                double* dtrac, int order, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && trac && Ux && Uy && Uz && dtrac, "null argument");
    NY_REQUIRE(order >= 1 && order <= 5, "order must be 1..5 (core/interpolate_tracer.f90)");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_RHS_TRACER, ny_stream(stream));
    k_upwind_linear<<<g.grid, g.block, 0, ny_stream(stream)>>>(trac, Ux, Uy, Uz, dtrac, order, make_ext(e));
    NY_CHECK_LAUNCH(ctx);
    return NY_OK;
}

// vortex_force.vortex_force with the Fortran's `linear` flag set (accumulates into du, like ny_vortex_force)
extern "C" int ny_vortex_force_linear(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz,
                                      const double* wx, const double* wy, const double* wz,
                                      double* dux, double* duy, double* duz, int order, ny_ext e, void* stream)
{
    NY_REQUIRE(ctx && Ux && Uy && Uz && wx && wy && wz && dux && duy && duz, "null argument");
    NY_REQUIRE(order >= 1 && order <= 5, "order must be 1..5 (core/interpolate.f90)");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5");
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    ny_prof_scope ps(ctx, NY_PROF_RHS_MOMENTUM, ny_stream(stream));
    k_vortex_force_linear<<<g.grid, g.block, 0, ny_stream(stream)>>>(Ux, Uy, Uz, wx, wy, wz, dux, duy, duz, order, make_ext(e));
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

static int rhs_impl(ny_ctx* ctx, const double* b, const double* Ux, const double* Uy, const double* Uz,
                    const double* wx, const double* wy, const double* wz, const double* ke,
                    double* db, double* dux, double* duy, double* duz,
                    double dz, int flags, ny_ext e, void* stream, const TsUpd* upd, const TrUpd* tupd = nullptr)
{
    const bool euler = flags & 1, linear = flags & 2;
    NY_REQUIRE(ctx && Ux && Uy && Uz && ke && (upd || (dux && duy && duz)), "null argument");
    NY_REQUIRE(euler || (b && (db || tupd)), "b and db are required unless the Euler flag is set");
    NY_REQUIRE(linear || (wx && wy && wz), "vorticity is required unless the linear flag is set");
    NY_REQUIRE(ext_ok(e), "every extent must be >= 5 (flux1d closure, core/weno.f90:106-153)");
    cudaStream_t st = ny_stream(stream);
    if (!euler) {
        int r = launch_upwind(ctx, b, Ux, Uy, Uz, db, false, 0.0, 0.0, 0.0, e, st, tupd);
        if (r != NY_OK) return r;
    }
    if (linear)
        return launch_momentum<false, false, true>(ctx, Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, 0.5 * dz,
                                                   euler ? 0 : 1, e, st, upd);
    return launch_momentum<false, true, true>(ctx, Ux, Uy, Uz, wx, wy, wz, ke, b, dux, duy, duz, 0.5 * dz,
                                              euler ? 0 : 1, e, st, upd);
}

extern "C" int ny_rhs(ny_ctx* ctx, const double* b, const double* Ux, const double* Uy, const double* Uz,
                      const double* wx, const double* wy, const double* wz, const double* ke,
                      double* db, double* dux, double* duy, double* duz,
                      double dz, int flags, ny_ext e, void* stream)
{
    return rhs_impl(ctx, b, Ux, Uy, Uz, wx, wy, wz, ke, db, dux, duy, duz, dz, flags, e, stream, nullptr);
}

extern "C" int ny_rhs_update_u(ny_ctx* ctx, const double* b, const double* Ux, const double* Uy, const double* Uz,
                               const double* wx, const double* wy, const double* wz, const double* ke,
                               double* db, double* const u[3], double* const ub[3], double* const un[3],
                               int mode, double dt, double dz, int flags, ny_ext e, void* stream)
{
    NY_REQUIRE(u && ub && un && mode >= 1 && mode <= 3, "bad argument");
    TsUpd upd;
    memset(&upd, 0, sizeof(upd));
    upd.mode = mode; upd.dt = dt;
    for (int a = 0; a < 3; a++) {
        NY_REQUIRE(u[a] && un[a] && (mode == 3 || ub[a]), "null field");
        upd.s[a] = u[a]; upd.sb[a] = ub[a]; upd.sn[a] = un[a];
    }
    return rhs_impl(ctx, b, Ux, Uy, Uz, wx, wy, wz, ke, db, nullptr, nullptr, nullptr, dz, flags, e, stream, &upd);
}

// Fields in s / sb / sn / out: 0 = b (ignored when the Euler flag is set), 1..3 = u components.
extern "C" int ny_rhs_step(ny_ctx* ctx, const double* Ux, const double* Uy, const double* Uz,
                           const double* wx, const double* wy, const double* wz, const double* ke,
                           const double* const s[4], const double* const sb[4], const double* const sn[4],
                           double* const out[4], const double* const add[4], int mode, double dt, double dz, int flags,
                           ny_ext e, void* stream)
{
    NY_REQUIRE(s && sb && sn && out && mode >= 1 && mode <= 3, "bad argument");
    const bool euler = flags & 1;
    TsUpd upd;
    memset(&upd, 0, sizeof(upd));
    upd.mode = mode; upd.dt = dt;
    for (int f = euler ? 1 : 0; f < 4; f++) {
        NY_REQUIRE(out[f] && s[f] && (mode != 3 || sn[f]) && (mode != 2 || sb[f]), "null field");
        NY_REQUIRE(out[f] != s[f] && (mode != 3 || out[f] != sn[f]) && (mode != 2 || out[f] != sb[f]),
                   "out must not be an array the launch reads (the caller rotates its buffers)");
    }
    for (int a = 0; a < 3; a++) {
        upd.s[a] = const_cast<double*>(s[a + 1]); upd.sb[a] = const_cast<double*>(sb[a + 1]);
        upd.sn[a] = const_cast<double*>(sn[a + 1]); upd.out[a] = out[a + 1];
        upd.add[a] = add ? add[a + 1] : nullptr;
    }
    TrUpd tupd;
    memset(&tupd, 0, sizeof(tupd));
    if (!euler) { tupd.mode = mode; tupd.dt = dt; tupd.sb = sb[0]; tupd.sn = sn[0]; tupd.out = out[0]; tupd.add = add ? add[0] : nullptr; }
    return rhs_impl(ctx, euler ? nullptr : s[0], Ux, Uy, Uz, wx, wy, wz, ke, nullptr, nullptr, nullptr, nullptr, dz, flags,
                    e, stream, &upd, euler ? nullptr : &tupd);
}
