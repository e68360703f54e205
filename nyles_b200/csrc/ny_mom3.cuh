// Momentum right-hand side (vortex force + Bernoulli gradient + time-scheme update of u) for the cells
// whose six WENO sweeps lie in the interior range of flux1d (3 <= s <= n-4 along every axis): a plane-marching
// kernel with TMA-staged tiles.  Included inside ny_rhs.cu's anonymous namespace; same arithmetic, statement
// for statement, as momentum_cell<.., INTERIOR = true, ..> (core/fortran_vortex_force.f90:66-80,145-159,
// core/weno.f90:25-54,106-153, core/fortran_bernoulli.f90:2-58), so the two are bit-identical.
//
// A CTA of 8 warps owns a tile of 32 x 8 columns (i, j) and a chunk of planes and marches along k.  Per plane,
// ONE thread issues eight bulk-tensor copies (wz, wx, wy, Ux, Uy, Uz, ke, b tiles with exactly the aprons the
// stencils reach: 3 cells for the vorticity along its sweep axes, 1 for the velocities and ke) into a ring of
// three stages; an mbarrier per stage tells the warps when a plane has landed.  Every x / y stencil value is then
// a shared-memory load at a compile-time offset from the thread's cell.  The upwind side of a WENO5 stencil is
// chosen by ADDRESS (centre and step of the five loads follow the sign of the face velocity) instead of loading
// six values and selecting five, which removes ten selects and one load per flux.  Along z the stencil values of
// wx, wy live in register queues, one new element per plane, and the two z-sweep fluxes of a cell are evaluated
// one iteration late (when plane k+2 of the queue has arrived), so that three stages suffice: the partial sums of
// plane k wait in four registers.  The LFAM3 / Euler update is applied on the fly as in k_momentum (ts_apply).
// There is no block-wide barrier in the plane loop: a warp that has finished with a plane arrives on the stage's
// "empty" mbarrier, and the warp whose turn it is to issue the next copies (the duty rotates over the eight warps)
// waits for those eight arrivals before it overwrites the stage; everybody else only waits for data.


namespace m3 {
constexpr int TX = 32, TY = 8, NS = 3;
constexpr int PW = 38, PN = 34;                    // row pitches of the wide (3-cell apron in x) and narrow tiles;
                                                   // a box must start on an EVEN column (16-byte boundary): i0-3 and i0-1 are
constexpr int O_WZ = 0;                            // (TY+6) x PW, box origin (i0-3, j0-3)
constexpr int O_WX = O_WZ + 544;                   // (TY+6) x PN, origin (i0-1, j0-3)
constexpr int O_WY = O_WX + 480;                   // TY x PW,     origin (i0-3, j0)
constexpr int O_UX = O_WY + TY * PW;               // (TY+1) x PN, origin (i0-1, j0)
constexpr int O_UY = O_UX + 320;                   // (TY+1) x PN, origin (i0-1, j0-1)
constexpr int O_UZ = O_UY + 320;                   // (TY+1) x PN, origin (i0-1, j0)
constexpr int O_KE = O_UZ + 320;                   // (TY+1) x PN, origin (i0-1, j0)
constexpr int O_B = O_KE + 320;                    // TY x PN,     origin (i0-1, j0)
constexpr int STAGE = O_B + 272;                   // doubles per stage (every offset is a multiple of 128 bytes)
static_assert((TY + 6) * PW <= 544 && (TY * PW) % 16 == 0 && (TY + 6) * PN <= 480 && (TY + 1) * PN <= 320 && TY * PN <= 272 && STAGE % 16 == 0, "stage layout");
constexpr int BYTES_NOB = ((TY + 6) * (PW + PN) + TY * PW + 4 * (TY + 1) * PN) * 8;
constexpr int BYTES_B = TY * PN * 8;
constexpr int SMEM = NS * STAGE * 8 + 64;            // + full[NS], empty[NS]

struct Maps { CUtensorMap wz, wx, wy, Ux, Uy, Uz, ke, b; };

// flux through the face of the thread's cell for a sweep whose stencil lies in shared memory: P points to the
// cell's own value of the vorticity component, `s` is the element stride of the sweep axis.  q(d) = P[(d-1) s]
// (fortran_vortex_force.f90:73-76: q(k) = vort(k-1)); u > 0 takes q(-2..2), else q(3..-1) (weno.f90:126-133).
template <bool FAST>
__device__ __forceinline__ double flux_smem(double u, const double* P, int s)
{
    const bool up = u > 0.0;
    int t = up ? s : -s, o = up ? -s : 0;
    // The step is made opaque to the compiler.  With a literal stride of 1, nvcc 12.9 turns the five indexed loads into
    // loads of both candidates plus selects and gets one select wrong (C[t] read C[+1] for t = -1: measured in
    // k_up3, tools/dbg_up3c.py); a step it cannot see through is a plain register offset.
    asm volatile("" : "+r"(t), "+r"(o));
    const double* C = P + o;
    return u * nyw::weno5<FAST>(C[-2 * t], C[-t], C[0], C[t], C[2 * t]);
}
}  // namespace m3

template <bool FAST>
__global__ void __launch_bounds__(m3::TX * m3::TY, 3)
k_mom3(const __grid_constant__ m3::Maps tm, const double* __restrict__ gUz, const double* __restrict__ gwx,
       const double* __restrict__ gwy, double* __restrict__ dux, double* __restrict__ duy, double* __restrict__ duz,
       double cff, int with_b, Ext e, int kchunk, TsUpd upd)
{
    using namespace m3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double* const sm = reinterpret_cast<double*>(smem_raw);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + NS * STAGE * 8);
    uint64_t* const empty = full + NS;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int i0 = 3 + (int)blockIdx.x * TX, j0 = 3 + (int)blockIdx.y * TY;
    const int k0 = 3 + (int)blockIdx.z * kchunk, k1 = min(k0 + kchunk, e.nz - 3);     // planes [k0, k1)
    const int i = i0 + tx, j = j0 + ty;
    const bool active = i <= e.nx - 4 && j <= e.ny - 4;
    const long long col = active ? (long long)j * e.sj + i : (long long)j0 * e.sj + i0;
    const uint32_t bytes = BYTES_NOB + (with_b ? BYTES_B : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { nytma::mbar_init(&full[s], 1); nytma::mbar_init(&empty[s], TY); }
        nytma::fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int p) {                      // plane p -> stage (p - k0) % NS
        const int st = (p - k0) % NS;
        double* d = sm + st * STAGE;
        nytma::mbar_expect_tx(&full[st], bytes);
        nytma::load_3d(d + O_WZ, &tm.wz, i0 - 3, j0 - 3, p, &full[st]);
        nytma::load_3d(d + O_WX, &tm.wx, i0 - 1, j0 - 3, p, &full[st]);
        nytma::load_3d(d + O_WY, &tm.wy, i0 - 3, j0, p, &full[st]);
        nytma::load_3d(d + O_UX, &tm.Ux, i0 - 1, j0, p, &full[st]);
        nytma::load_3d(d + O_UY, &tm.Uy, i0 - 1, j0 - 1, p, &full[st]);
        nytma::load_3d(d + O_UZ, &tm.Uz, i0 - 1, j0, p, &full[st]);
        nytma::load_3d(d + O_KE, &tm.ke, i0 - 1, j0, p, &full[st]);
        if (with_b) nytma::load_3d(d + O_B, &tm.b, i0 - 1, j0, p, &full[st]);
    };
    const int plast = k1 + 1;                      // last plane the chunk reads (top of the z queues)
    if (threadIdx.x == 0) {
        issue(k0);
        issue(k0 + 1);
    }

    // z queues: at the top of iteration k they hold w[k-5 .. k]; only w[k0-3 ..] is ever used
    double qx[6], qy[6];
    qx[0] = qx[1] = qy[0] = qy[1] = 0.0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const long long c = (long long)(k0 - 3 + d) * e.sk + col;
        qx[d + 2] = gwx[c];
        qy[d + 2] = gwy[c];
    }
    // sums of Uz over the averaging axis of the target component (fortran_vortex_force.f90:68-70) at the planes
    // k-1 (z?1) and k-2 (z?2): a = across x (for du_x), b = across y (for du_y)
    double za1, zb1, za2 = 0.0, zb2 = 0.0;
    {
        const long long c = (long long)(k0 - 1) * e.sk + col;
        const double u0 = gUz[c];
        za1 = u0 + gUz[c + 1];
        zb1 = u0 + gUz[c + e.sj];
    }
    double pax = 0.0, pay = 0.0, gx = 0.0, gy = 0.0;     // plane k-1: partial du_x, du_y and the two ke differences

    const int oW = (ty + 3) * PW + tx + 3, oX = (ty + 3) * PN + tx + 1, oY = ty * PW + tx + 3, oN = ty * PN + tx + 1;
    nytma::mbar_wait(&full[0], 0);
    for (int k = k0; k <= k1; k++) {
        const int t = k - k0;
        {   // plane k+2 goes into the stage of plane k-1 once all eight warps have released it; this plane it is the
            // turn of warp (t+2) % 8 to wait for that and to issue the copies
            const int t2 = t + 2;
            if (ty == (t2 & (TY - 1)) && tx == 0 && k + 2 <= plast) {
                if (t >= 1) nytma::mbar_wait(&empty[t2 % NS], (uint32_t)((t - 1) / NS) & 1u);
                issue(k + 2);
            }
            const int t1 = t + 1;                                   // plane k+1 has landed
            nytma::mbar_wait(&full[t1 % NS], (uint32_t)(t1 / NS) & 1u);
        }
        const double* const A = sm + (t % NS) * STAGE;              // plane k
        const double* const B = sm + ((t + 1) % NS) * STAGE;        // plane k+1
        const long long c = (long long)k * e.sk + col;
        // ---- the z sweeps of plane k-1, now that w[k+1] is there (vortex_force.py:69-81, passes 3 and 2)
#pragma unroll
        for (int d = 0; d < 5; d++) { qx[d] = qx[d + 1]; qy[d] = qy[d + 1]; }
        qx[5] = B[O_WX + oX];
        qy[5] = B[O_WY + oY];
        if (k > k0 && active) {
            const double ua = 0.25 * (za2 + za1), ub = 0.25 * (zb2 + zb1);
            // q(d) = w[(k-1) + d - 1] = queue[d + 2]  (the queue holds w[k-4 .. k+1])
            double ax = pax - nyw::hot_flux<FAST>(ua, [&](int d) { return qy[d + 2]; });      // pass 3 direc: -F_z(omega_y)
            double ay = pay + nyw::hot_flux<FAST>(ub, [&](int d) { return qx[d + 2]; });      // pass 2 flip : +F_z(omega_x)
            ax = ax - gx;                                                                     // gradke, fortran_bernoulli.f90:20
            ay = ay - gy;
            const long long cm = c - e.sk;
            if (upd.mode == 0) { dux[cm] = ax; duy[cm] = ay; }
            else { ts_apply(upd, 0, cm, ax); ts_apply(upd, 1, cm, ay); }
        }
        if (k == k1) break;
        // ---- the x and y sweeps of plane k
        const double ux00 = A[O_UX + oN], ux0m = A[O_UX + oN - 1];                 // Ux[j][i], Ux[j][i-1]
        const double uy00 = A[O_UY + oN + PN], uym0 = A[O_UY + oN];                // Uy[j][i], Uy[j-1][i]
        const double uz00 = A[O_UZ + oN];
        za2 = za1; zb2 = zb1;
        za1 = uz00 + A[O_UZ + oN + 1];
        zb1 = uz00 + A[O_UZ + oN + PN];
        if (active) {
            // pass 1 flip: +F_y(omega_z), U = Uy averaged across x
            const double u1 = 0.25 * ((uym0 + A[O_UY + oN + 1]) + (uy00 + A[O_UY + oN + PN + 1]));
            pax = 0.0 + m3::flux_smem<FAST>(u1, A + O_WZ + oW, PW);
            // pass 1 direc: -F_x(omega_z), U = Ux averaged across y
            const double u2 = 0.25 * ((ux0m + A[O_UX + oN + PN - 1]) + (ux00 + A[O_UX + oN + PN]));
            pay = 0.0 - m3::flux_smem<FAST>(u2, A + O_WZ + oW, 1);
            // pass 2 direc: -F_y(omega_x), U = Uy averaged across z
            const double u3 = 0.25 * ((uym0 + B[O_UY + oN]) + (uy00 + B[O_UY + oN + PN]));
            double az = 0.0 - m3::flux_smem<FAST>(u3, A + O_WX + oX, PN);
            // pass 3 flip: +F_x(omega_y), U = Ux averaged across z
            const double u4 = 0.25 * ((ux0m + B[O_UX + oN - 1]) + (ux00 + B[O_UX + oN]));
            az = az + m3::flux_smem<FAST>(u4, A + O_WY + oY, 1);
            // Bernoulli (fortran_bernoulli.f90:20,50-51)
            const double ke0 = A[O_KE + oN];
            gx = A[O_KE + oN + 1] - ke0;
            gy = A[O_KE + oN + PN] - ke0;
            az = az - (B[O_KE + oN] - ke0);
            if (with_b) az = az + cff * (B[O_B + oN] + A[O_B + oN]);
            if (upd.mode == 0) duz[c] = az;
            else ts_apply(upd, 2, c, az);
        }
        // this warp is done with plane k (its stage is only re-used for plane k+3)
        __syncwarp();
        if (tx == 0) nytma::mbar_arrive(&empty[t % NS]);
    }
}
