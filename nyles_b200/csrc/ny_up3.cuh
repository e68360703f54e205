// Tracer right-hand side -div(U trac) (+ time-scheme update) for the WHOLE array as a plane-marching kernel with
// TMA-staged tiles; included inside ny_rhs.cu's anonymous namespace.  Same arithmetic and the same order of
// accumulation as k_upwind2 (core/fortran_upwind.f90:66-82 per direction, core/weno.f90:106-153, tracer.py:44-72,
// timescheme.py:131-175), so the two are bit-identical; k_upwind2 stays for odd nx, for the diffusion variant and for
// small grids.
//
// A CTA owns a tile of 32 x 8 columns (i, j) and a chunk of planes.  Warps 0..7 own one row each: a thread evaluates
// ONE flux per axis and plane -- through the face on the + side of its cell -- and needs the flux through the face
// on its - side from somebody else:
//   x : the lane to its left (shuffle); lane 0 takes it from the helper warp
//   y : the warp that owns the row below, through shared memory; row 0 takes it from the helper warp
//   z : its own previous plane.
// Warp 8, the helper, evaluates the 32 y faces below the tile and the 8 x faces left of it, so that all 256 cells of
// the tile are produced (k_upwind2's overlapping tiles produce 31 x 7 of 32 x 8).  Nobody waits for the whole block:
// a row warp publishes its y fluxes and arrives on an mbarrier that only the warp above waits for; the helper has its
// own; a plane's stage is released through an "empty" mbarrier that only the helper, which issues the TMA copies,
// waits for.  x / y stencils are shared-memory loads at compile-time offsets with the upwind side chosen
// by address (ny_mom3.cuh's flux_smem, whose q(d) = P[(d-1) s]: the tracer's q(d) = trac[s+d] is P = cell s+1);
// tiles that touch a wall evaluate that axis with the closures of flux1d (line_flux), which
// is CTA-uniform; planes next to the bottom / top do the same for z.  The z stencil lives in a register queue whose
// new end (plane k+4) is loaded one iteration ahead.
// A cell is FINISHED one iteration late: at iteration t a warp evaluates the three fluxes of plane t and publishes
// its y fluxes, and only then combines plane t-1 with the y fluxes its neighbour published a whole iteration ago --
// so a warp practically never waits for another one (with same-plane combination the eight rows of a tile moved in
// lockstep and the fp64 pipe idled 46 % of the time; profiles/r2_d_*).
namespace u3 {
constexpr int TX = 32, TY = 8, NS = 3, NW = TY + 1, NF = 4;   // NF: ring of published fluxes
constexpr int PT = 40, PU = 34;                    // row pitches: tracer tile (origin i0-4), velocity tiles (origin i0-2)
constexpr int O_T = 0;                             // (TY+6) x PT, box origin (i0-4, j0-3)
constexpr int O_UX = O_T + (TY + 6) * PT;          // TY x PU,     origin (i0-2, j0)
constexpr int O_UY = O_UX + 272;                   // (TY+1) x PU, origin (i0-2, j0-1)
constexpr int O_UZ = O_UY + 320;                   // TY x PU,     origin (i0-2, j0)
constexpr int STAGE = O_UZ + 272;                  // doubles per stage
static_assert(((TY + 6) * PT) % 16 == 0 && TY * PU <= 272 && (TY + 1) * PU <= 320 && STAGE % 16 == 0, "stage layout");
constexpr int BYTES = ((TY + 6) * PT + 2 * TY * PU + (TY + 1) * PU) * 8;
constexpr int O_FY = NS * STAGE;                   // y fluxes [NF][TY+1][32]: row r = the face below tile row r
constexpr int O_FX0 = O_FY + NF * (TY + 1) * 32;   // x fluxes through the faces left of the tile [NF][TY]
constexpr int O_BAR = O_FX0 + NF * TY;             // full[NS], empty[NS], hdone[NF], yrow[NF][TY]
constexpr int SMEM = (O_BAR + 2 * NS + NF + NF * TY) * 8;
struct Maps { CUtensorMap t, Ux, Uy, Uz; };
}  // namespace u3

template <bool FAST>
__global__ void __launch_bounds__(u3::NW * 32, 3)
k_up3(const __grid_constant__ u3::Maps tm, const double* __restrict__ trac, const double* __restrict__ gUz,
      double* __restrict__ dtrac, Ext e, int kchunk, TrUpd upd)
{
    using namespace u3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* const sm = reinterpret_cast<double*>(smem_raw);
    uint64_t* const full = reinterpret_cast<uint64_t*>(sm + O_BAR);
    uint64_t* const empty = full + NS;
    uint64_t* const hdone = empty + NS;
    uint64_t* const yrow = hdone + NF;
    const int tx = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool helper = w == TY;
    const int i0 = (int)blockIdx.x * TX, j0 = (int)blockIdx.y * TY;
    const int k0 = (int)blockIdx.z * kchunk, k1 = min(k0 + kchunk, e.nz);            // planes [k0, k1)
    const int i = i0 + tx, j = j0 + (helper ? 0 : w);
    const bool active = !helper && i < e.nx && j < e.ny;
    const long long col = active ? (long long)j * e.sj + i : 0;
    // every face the tile evaluates along x (i0-1 .. i0+31) / y (j0-1 .. j0+7) is in the interior range of flux1d
    const bool x_hot = i0 - 1 >= 2 && i0 + TX - 1 <= e.nx - 4, y_hot = j0 - 1 >= 2 && j0 + TY - 1 <= e.ny - 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            nytma::mbar_init(&full[s], 1);
            nytma::mbar_init(&empty[s], NW);
        }
        for (int s = 0; s < NF; s++) {
            nytma::mbar_init(&hdone[s], 1);
            for (int r = 0; r < TY; r++) nytma::mbar_init(&yrow[s * TY + r], 1);
        }
        nytma::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int p) {                      // plane p -> stage (p - k0) % NS
        const int st = (p - k0) % NS;
        double* d = sm + st * STAGE;
        nytma::mbar_expect_tx(&full[st], BYTES);
        nytma::load_3d(d + O_T, &tm.t, i0 - 4, j0 - 3, p, &full[st]);
        nytma::load_3d(d + O_UX, &tm.Ux, i0 - 2, j0, p, &full[st]);
        nytma::load_3d(d + O_UY, &tm.Uy, i0 - 2, j0 - 1, p, &full[st]);
        nytma::load_3d(d + O_UZ, &tm.Uz, i0 - 2, j0, p, &full[st]);
    };
    if (threadIdx.x == 0) {
        issue(k0);
        if (k0 + 1 < k1) issue(k0 + 1);
    }

    // z queue zq[d+2] = trac[k+d], d = -2..3, and the flux through the face below the chunk
    double zq[6];
    double Fz_prev = 0.0;
    if (active) {
#pragma unroll
        for (int d = -2; d <= 3; d++) {
            const int kk = k0 - 1 + d;
            zq[d + 2] = (kk >= 0 && kk < e.nz) ? trac[(long long)kk * e.sk + col] : 0.0;
        }
        if (k0 > 0) {
            const double uz = gUz[(long long)(k0 - 1) * e.sk + col];
            Fz_prev = nyw::line_flux<FAST>(k0 - 1, e.nz, uz, [&](int d) { return zq[d + 2]; });
        }
#pragma unroll
        for (int d = 0; d < 5; d++) zq[d] = zq[d + 1];
        zq[5] = (k0 + 3 < e.nz) ? trac[(long long)(k0 + 3) * e.sk + col] : 0.0;
    }

    const int oT = ((helper ? 0 : w) + 3) * PT + tx + 4, oU = (helper ? 0 : w) * PU + tx + 2;
    // plane t-1 waiting to be finished: its x flux pair, its own y flux, its z flux pair
    double pFx = 0.0, pFxm = 0.0, pFy = 0.0, pFz = 0.0, pFzp = 0.0;
    // combine and store plane kk = k0 + tt (tt = its iteration); v = trac there
    auto finish = [&](int tt, double v) {
        const int fs = tt % NF, kk = k0 + tt;
        const uint32_t fpar = (uint32_t)(tt / NF) & 1u;
        double Fxm = pFxm;
        if (tx == 0 && i0 > 0) {
            nytma::mbar_wait(&hdone[fs], fpar);
            Fxm = sm[O_FX0 + fs * TY + w];
        }
        // the y fluxes of the row below: from the helper (row 0) or from the warp that owns that row
        if (w == 0) { if (j0 > 0) nytma::mbar_wait(&hdone[fs], fpar); }
        else nytma::mbar_wait(&yrow[fs * TY + w - 1], fpar);
        if (active) {
            const long long c = (long long)kk * e.sk + col;
            const double Fym = sm[O_FY + (fs * (TY + 1) + w) * 32 + tx];
            double acc = 0.0;                                             // tracer.py:70-71
            acc = (i == 0) ? acc - pFx : acc + Fxm - pFx;
            acc = (j == 0) ? acc - pFy : acc + Fym - pFy;
            acc = (kk == 0) ? acc - pFz : acc + pFzp - pFz;
            if (upd.add) acc = acc + upd.add[c];                          // model_les.py:143-144
            if (upd.mode == 0) dtrac[c] = acc;
            else if (upd.mode == 2) {                                     // timescheme.py:144-162
                const double told = upd.sb[c], lf = told + (2. * upd.dt) * acc;
                upd.out[c] = (1. / 12.) * (5. * lf + 8. * v - told);
            } else if (upd.mode == 3) upd.out[c] = upd.sn[c] + upd.dt * acc;   // timescheme.py:170-175
            else upd.out[c] = v + upd.dt * acc;                           // timescheme.py:131-139
        }
    };
    for (int k = k0; k < k1; k++) {
        const int t = k - k0, st = t % NS, fs = t % NF;
        const uint32_t par = (uint32_t)(t / NS) & 1u;
        // plane k+2 goes into the stage of plane k-1 once all nine warps have released it.  The helper is the
        // producer: it has a third less work than a row warp and a whole iteration of slack before its fluxes are
        // read, so the wait for the slowest row warp costs nothing (with the duty rotating over the row warps, the
        // warp whose turn it was stalled until everybody had finished the previous plane -- 23 % of all stall samples)
        if (helper && tx == 0 && k + 2 < k1) {
            const int t2 = t + 2;
            if (t >= 1) nytma::mbar_wait(&empty[t2 % NS], (uint32_t)((t - 1) / NS) & 1u);
            issue(k + 2);
        }
        nytma::mbar_wait(&full[st], par);
        const double* const A = sm + st * STAGE;
        double* const sFy = sm + O_FY + fs * (TY + 1) * 32;
        double* const sFx0 = sm + O_FX0 + fs * TY;
        if (helper) {
            // the x faces left of the tile (column i0-1, row j0+tx) and the y faces below it (row j0-1, column i0+tx)
            if (i0 > 0 && tx < TY && j0 + tx < e.ny) {
                const double ux = A[O_UX + tx * PU + 1];                               // Ux[j0+tx][i0-1]
                const double* P = A + O_T + (tx + 3) * PT + 3;                         // trac[j0+tx][i0-1]
                sFx0[tx] = x_hot ? m3::flux_smem<FAST>(ux, P + 1, 1)
                                 : nyw::line_flux<FAST>(i0 - 1, e.nx, ux, [&](int d) { return P[d]; });
            }
            if (j0 > 0 && i < e.nx) {
                const double uy = A[O_UY + tx + 2];                                    // Uy[j0-1][i]
                const double* P = A + O_T + 2 * PT + tx + 4;                           // trac[j0-1][i]
                sFy[tx] = y_hot ? m3::flux_smem<FAST>(uy, P + PT, PT)
                                : nyw::line_flux<FAST>(j0 - 1, e.ny, uy, [&](int d) { return P[d * PT]; });
            }
            __syncwarp();
            if (tx == 0) { nytma::mbar_arrive(&hdone[fs]); nytma::mbar_arrive(&empty[st]); }
            continue;
        }
        const long long c = (long long)k * e.sk + col;
        double nq = 0.0;                                   // the new end of the z queue for the next plane
        if (active && k + 4 < e.nz) nq = trac[c + 4 * e.sk];
        const double* const P = A + O_T + oT;
        double Fy = 0.0, Fx = 0.0, Fz = 0.0;
        if (active) {
            const double uy = A[O_UY + oU + PU];
            Fy = y_hot ? m3::flux_smem<FAST>(uy, P + PT, PT) : nyw::line_flux<FAST>(j, e.ny, uy, [&](int d) { return P[d * PT]; });
        }
        sFy[(w + 1) * 32 + tx] = Fy;
        __syncwarp();
        if (tx == 0) nytma::mbar_arrive(&yrow[fs * TY + w]);
        if (active) {
            const double ux = A[O_UX + oU], uz = A[O_UZ + oU];
            Fx = x_hot ? m3::flux_smem<FAST>(ux, P + 1, 1) : nyw::line_flux<FAST>(i, e.nx, ux, [&](int d) { return P[d]; });
            Fz = (k >= 2 && k <= e.nz - 4) ? nyw::hot_flux<FAST>(uz, [&](int d) { return zq[d + 2]; })
                                            : nyw::line_flux<FAST>(k, e.nz, uz, [&](int d) { return zq[d + 2]; });
        }
        const double Fxm = __shfl_up_sync(0xffffffffu, Fx, 1);
        if (t >= 1) finish(t - 1, zq[1]);
        // this warp is done with the tile of plane k AND with the fluxes published for plane k-1 (whose ring slot is
        // only rewritten after plane k+3 has landed, which this arrival gates)
        __syncwarp();
        if (tx == 0) nytma::mbar_arrive(&empty[st]);
        pFx = Fx; pFxm = Fxm; pFy = Fy; pFz = Fz; pFzp = Fz_prev;
        Fz_prev = Fz;
#pragma unroll
        for (int d = 0; d < 5; d++) zq[d] = zq[d + 1];
        zq[5] = nq;
    }
    if (!helper) finish(k1 - 1 - k0, zq[1]);
}
