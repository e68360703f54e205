// Fused legs of the multigrid V-cycle for box domains (included inside ny_mg.cu's anonymous namespace).
//
//   down leg  : smooth (two Jacobi sweeps) + residual + restriction   x, b -> x', b_coarse     (r, y never stored)
//   up leg    : prolongation + smooth                                 x, x_coarse, b -> x'
//   last leg  : prolongation + smooth + residual norm (finest level)  x, x_coarse, b -> x', sum r^2
// which replace the sequences smooth; residual; restriction / prolongation; smooth [; residual; norm] of
// solvers.f90:35-55 and operators.f90:81-244 with the same per-cell arithmetic in the same order, so
// the iterates are bit-identical to the one-kernel-per-operator path.
//
// One CTA (16 warps) owns a tile of the plane and a chunk of planes and marches along k.  Planes of x
// and b (and the coarse x tile) are staged in shared memory by TMA bulk-tensor copies issued two planes
// ahead by one thread and tracked with mbarriers; tiles that stick out of the array are zero-filled by
// the hardware, so the plane loop has no bounds tests.  A thread owns a 2 x 2 patch of columns and keeps
// the k-neighbours of every intermediate field in registers; of the in-plane neighbours, the columns come
// from the neighbouring lanes (warp shuffles), the rows above / below the patch from shared memory
// (x: the TMA tile itself; y and x': double-buffered tiles, one __syncthreads per plane).
//
// Pipeline at iteration p (array plane indices):  [prolong x[p+1]]  ->  y[p] = sweep1(x[p-1..p+1])
//   ->  x'[p-1] = sweep2(y[p-2..p])  ->  r[p-2] = residual(x'[p-3..p-1])  -> restriction / norm.
// Every stage is evaluated on the whole 32 x 64 region; its valid part shrinks by one ring per stage, the
// tile that is stored is the region minus a 3-wide apron (4 rows on top for the down leg, which keeps
// the patch rows aligned with the coarse rows).  Cells outside the domain (wall halos) keep x, cells of
// periodic / slab halos are recomputed from the (consistent) 3-wide halo of x and b instead of being
// exchanged between the fused operators; the caller fills the halo of the result afterwards.
#include "ny_tma.cuh"

constexpr int VL_NW = 16;                       // warps per CTA
constexpr int VL_RJ = 2 * VL_NW;                // region rows
constexpr int VL_RI = 64;                       // region columns
constexpr int VL_TI = VL_RI - 6;                // stored columns per tile
constexpr int VL_S = 4;                         // ring of plane stages (x and b)
constexpr int VL_SC = 6;                        // ring of coarse planes
constexpr int VL_L2AHEAD = 6;                   // planes ahead of the one being issued that are prefetched into L2
constexpr int VL_CJ = VL_NW + 2, VL_CI = 34;    // coarse tile (rows, columns)
constexpr int VL_PLANE = VL_RJ * VL_RI;         // doubles per plane tile
constexpr int VL_CBYTES = VL_CJ * VL_CI * 8;    // bytes of one coarse tile
constexpr int VL_CSLOT = ((VL_CBYTES + 127) / 128) * 128;
enum { POST_NONE = 0, POST_RESTRICT = 1, POST_NORM = 2 };

template <bool PRO, int POST>
struct VlegLayout {
    static constexpr int apron_top = POST == POST_RESTRICT ? 4 : 3;
    static constexpr int tj = VL_RJ - 2 * apron_top;          // stored rows per tile: 24 (down leg) or 26
    static constexpr int off_x = 0;
    static constexpr int off_b = off_x + VL_S * VL_PLANE * 8;
    static constexpr int off_y = off_b + VL_S * VL_PLANE * 8;
    static constexpr int off_z = off_y + 2 * VL_PLANE * 8;
    static constexpr int off_c = off_z + (POST != POST_NONE ? 2 * VL_PLANE * 8 : 0);
    static constexpr int off_misc = off_c + (PRO ? VL_SC * VL_CSLOT : 0);
    static constexpr int bytes = off_misc + 320;               // recip[8], pcoef[4], red[16], full[VL_S]
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

struct P4 { double2 A, B; };           // the 2 x 2 patch of a thread: rows A, B, columns (.x, .y)

// per-thread constants of a leg (live in registers after inlining)
struct VlegCtx {
    double *sx, *sb, *sy, *sz;
    const unsigned char* sc;
    const double *s_recip, *s_pcoef;
    double* xo; double* bc;
    Box g, gc;
    double omega, cff1;
    int cA0, cA1, cB0, cB1;            // in-plane neighbour counts (negative: cell outside the domain)
    double rA0, rA1, rB0, rB1;         // 1 / (count + 2)
    // which cells of the patch belong to the stored tile and to the interior, as ONE register of flags (made opaque
    // to the compiler, which otherwise re-derives every flag from threadIdx in every plane to save registers):
    // bits 0-2: row A stores both / only .x / only .y; bits 3-5: the same for row B; bits 6-9: cell A.x, A.y, B.x,
    // B.y counts in the residual norm; bit 7 also guards the restriction store
    unsigned sf;
    int oA, oUp, oDn, oC;              // shared-memory offsets (doubles) inside a plane / coarse tile
    unsigned gA, sk32;                 // element offset of (row A, column .x) inside a plane; plane stride (a level has < 2^32 cells)
    unsigned bcA, csk32;               // coarse cell under the patch (row, column part of its offset); coarse plane stride
    int exA0, exA1, exB0, exB1;
    int ai, ajA;
    int k0, k1, pb1, pb2, pb3;
};

// One plane of the pipeline.  FAST: every stage runs, the three planes involved have both k-neighbours
// inside the domain and the plane written lies in the chunk -- no tests left but the per-cell domain
// flags.  The roles (minus, centre, plus) of the register sets rotate statically in the caller.
// INT: no cell of the CTA's region lies on or next to an x / y wall, and every coarse cell the prolongation
// reads is inside the domain: each in-plane neighbour count is 4, so 1/diag, diag and Pcoef depend on the
// plane only (CTA-uniform) and no per-cell domain flag is left.  These tiles -- the bulk of a large level --
// are launched as their own kernel instance, whose register allocation is free of the per-cell constants.
template <bool PRO, int POST, bool FAST, bool INT>
__device__ __forceinline__ void vleg_plane(const VlegCtx& c, int p, int t, const P4& xm, const P4& xc, P4& xp,
                                           const P4& ym, const P4& yc, P4& yp, const P4& zm, const P4& zc, P4& zp,
                                           P4& bp, const P4& b1, const P4& b2, double& acc, double& rsum)
{
    const Box& g = c.g;
    double* const px = c.sx + (t & (VL_S - 1)) * VL_PLANE;                          // x[p+1]
    const int sprev = ((t + VL_S - 1) & (VL_S - 1)) * VL_PLANE;
    const double* const pxc = c.sx + sprev;                                          // x[p]
    const double* const pbc = c.sb + sprev;                                          // b[p]
    const int oA = c.oA, oB = c.oA + VL_RI;
    xp.A = ld2(px + oA); xp.B = ld2(px + oB);
    if (PRO) {
        // x[p+1] += Pcoef * (weights 27 9 9 3 / 9 3 3 1 of the eight nearest coarse cells), basicoperators.f90:173-231
        const int f = p + 1, kf = f - NH, akc = NH + (kf >> 1), ako = (kf & 1) ? akc + 1 : akc - 1;
        const double* cb = reinterpret_cast<const double*>(c.sc + (akc % VL_SC) * VL_CSLOT) + c.oC;
        const double* co = reinterpret_cast<const double*>(c.sc + (ako % VL_SC) * VL_CSLOT) + c.oC;
        const double b00 = cb[0], b01 = cb[1], b10 = cb[VL_CI], b11 = cb[VL_CI + 1];
        const double o00 = co[0], o01 = co[1], o10 = co[VL_CI], o11 = co[VL_CI + 1];
        const int ez = FAST ? 1 : (int)in_z(c.gc, ako);
        const bool fz = FAST ? true : in_z(g, f);
        const double pbA0 = 9 * b00 + 3 * b01 + 3 * b10 + b11, poA0 = 9 * o00 + 3 * o01 + 3 * o10 + o11;
        const double pbA1 = 9 * b01 + 3 * b00 + 3 * b11 + b10, poA1 = 9 * o01 + 3 * o00 + 3 * o11 + o10;
        const double pbB0 = 9 * b10 + 3 * b11 + 3 * b00 + b01, poB0 = 9 * o10 + 3 * o11 + 3 * o00 + o01;
        const double pbB1 = 9 * b11 + 3 * b10 + 3 * b01 + b00, poB1 = 9 * o11 + 3 * o10 + 3 * o01 + o00;
        if (INT) {
            if (fz) {
                const double pc = FAST ? 1.0 / 64.0 : c.s_pcoef[2 + ez];
                xp.A.x = xp.A.x + pc * (3 * pbA0 + poA0);
                xp.A.y = xp.A.y + pc * (3 * pbA1 + poA1);
                xp.B.x = xp.B.x + pc * (3 * pbB0 + poB0);
                xp.B.y = xp.B.y + pc * (3 * pbB1 + poB1);
            }
        } else {
            if (fz && c.cA0 >= 0) xp.A.x = xp.A.x + c.s_pcoef[c.exA0 + ez] * (3 * pbA0 + poA0);
            if (fz && c.cA1 >= 0) xp.A.y = xp.A.y + c.s_pcoef[c.exA1 + ez] * (3 * pbA1 + poA1);
            if (fz && c.cB0 >= 0) xp.B.x = xp.B.x + c.s_pcoef[c.exB0 + ez] * (3 * pbB0 + poB0);
            if (fz && c.cB1 >= 0) xp.B.y = xp.B.y + c.s_pcoef[c.exB1 + ez] * (3 * pbB1 + poB1);
        }
        st2(px + oA, xp.A);             // the rows above / below read the prolonged plane in the next iteration
        st2(px + oB, xp.B);
        nytma::fence_proxy_async();
    }
    if (FAST || p >= c.pb1) {           // ---- sweep 1 at plane p (fsmoother3d, basicoperators.f90:363-400)
        bp.A = ld2(pbc + oA); bp.B = ld2(pbc + oB);
        const double2 up = ld2(pxc + c.oUp), dn = ld2(pxc + c.oDn);
        const int cz = FAST ? 2 : cnt_z(g, p);
        if (INT) {
            const double rc = cz == 2 ? 1.0 / 6.0 : c.s_recip[max(4 + cz, 0)];
            sweep_patch(xc.A, xc.B, xm.A, xm.B, xp.A, xp.B, up, dn, bp.A, bp.B, rc, rc, rc, rc, c.omega, c.cff1, yp.A, yp.B);
        } else if (cz == 2)
            sweep_patch(xc.A, xc.B, xm.A, xm.B, xp.A, xp.B, up, dn, bp.A, bp.B, c.rA0, c.rA1, c.rB0, c.rB1,
                        c.omega, c.cff1, yp.A, yp.B);
        else                            // planes next to the k ends (CTA-uniform)
            sweep_patch(xc.A, xc.B, xm.A, xm.B, xp.A, xp.B, up, dn, bp.A, bp.B, c.s_recip[max(c.cA0 + cz, 0)],
                        c.s_recip[max(c.cA1 + cz, 0)], c.s_recip[max(c.cB0 + cz, 0)], c.s_recip[max(c.cB1 + cz, 0)],
                        c.omega, c.cff1, yp.A, yp.B);
        double* const py = c.sy + (p & 1) * VL_PLANE;
        st2(py + oA, yp.A);
        st2(py + oB, yp.B);
    }
    if (FAST || p >= c.pb2) {           // ---- sweep 2 at plane q = p-1
        const int q = p - 1;
        const double* const py = c.sy + (q & 1) * VL_PLANE;
        const double2 up = ld2(py + c.oUp), dn = ld2(py + c.oDn);
        const int cz = FAST ? 2 : cnt_z(g, q);
        if (INT) {
            const double rc = cz == 2 ? 1.0 / 6.0 : c.s_recip[max(4 + cz, 0)];
            sweep_patch(yc.A, yc.B, ym.A, ym.B, yp.A, yp.B, up, dn, b1.A, b1.B, rc, rc, rc, rc, c.omega, c.cff1, zp.A, zp.B);
        } else if (cz == 2)
            sweep_patch(yc.A, yc.B, ym.A, ym.B, yp.A, yp.B, up, dn, b1.A, b1.B, c.rA0, c.rA1, c.rB0, c.rB1,
                        c.omega, c.cff1, zp.A, zp.B);
        else
            sweep_patch(yc.A, yc.B, ym.A, ym.B, yp.A, yp.B, up, dn, b1.A, b1.B, c.s_recip[max(c.cA0 + cz, 0)],
                        c.s_recip[max(c.cA1 + cz, 0)], c.s_recip[max(c.cB0 + cz, 0)], c.s_recip[max(c.cB1 + cz, 0)],
                        c.omega, c.cff1, zp.A, zp.B);
        const bool qz = FAST ? true : in_z(g, q);
        if (INT) {
            if (!qz) { zp.A = xm.A; zp.B = xm.B; }           // a plane outside the domain: x is kept
        } else {
            if (!(qz && c.cA0 >= 0)) zp.A.x = xm.A.x;        // outside the domain: x is kept
            if (!(qz && c.cA1 >= 0)) zp.A.y = xm.A.y;
            if (!(qz && c.cB0 >= 0)) zp.B.x = xm.B.x;
            if (!(qz && c.cB1 >= 0)) zp.B.y = xm.B.y;
        }
        if (POST != POST_NONE) {
            double* const pz = c.sz + (q & 1) * VL_PLANE;
            st2(pz + oA, zp.A);
            st2(pz + oB, zp.B);
        }
        if (FAST || (q >= c.k0 && q < c.k1)) {
            // independent single-instruction bodies: predicated stores, no branches (lanes 1 and 30 of every warp
            // hold a column pair that straddles the edge of the stored tile)
            double* const dA = c.xo + ((unsigned)q * c.sk32 + c.gA);
            double* const dB = dA + g.sj;
            const unsigned sf = c.sf;
            if (sf & 1u) st2(dA, zp.A);
            if (sf & 2u) dA[0] = zp.A.x;
            if (sf & 4u) dA[1] = zp.A.y;
            if (sf & 8u) st2(dB, zp.B);
            if (sf & 16u) dB[0] = zp.B.x;
            if (sf & 32u) dB[1] = zp.B.y;
        }
    }
    if (POST != POST_NONE && (FAST || p >= c.pb3)) {     // ---- residual at plane q = p-2 (fresidual3d, :300-323)
        const int q = p - 2;
        const double* const pz = c.sz + (q & 1) * VL_PLANE;
        const double2 up = ld2(pz + c.oUp), dn = ld2(pz + c.oDn);
        const double lA = shfl_up1(zc.A.y), rgA = shfl_dn1(zc.A.x), lB = shfl_up1(zc.B.y), rgB = shfl_dn1(zc.B.x);
        const double sA0 = lA + zc.A.y + up.x + zc.B.x + zm.A.x + zp.A.x;
        const double sA1 = zc.A.x + rgA + up.y + zc.B.y + zm.A.y + zp.A.y;
        const double sB0 = lB + zc.B.y + zc.A.x + dn.x + zm.B.x + zp.B.x;
        const double sB1 = zc.B.x + rgB + zc.A.y + dn.y + zm.B.y + zp.B.y;
        const int cz = FAST ? 2 : cnt_z(g, q);
        const double r0 = b2.A.x + (INT ? (double)(4 + cz) : (double)(c.cA0 + cz)) * zc.A.x - sA0;
        const double r1 = b2.A.y + (INT ? (double)(4 + cz) : (double)(c.cA1 + cz)) * zc.A.y - sA1;
        const double r2 = b2.B.x + (INT ? (double)(4 + cz) : (double)(c.cB0 + cz)) * zc.B.x - sB0;
        const double r3 = b2.B.y + (INT ? (double)(4 + cz) : (double)(c.cB1 + cz)) * zc.B.y - sB1;
        if (POST == POST_NORM) {
            if (FAST || q < c.k1) {     // fnorm, basicoperators.f90:422-440 (interior cells, msk = 1)
                double a = 0.0;
                const unsigned sf = c.sf;
                if (sf & 64u) a = a + r0 * r0;
                if (sf & 128u) a = a + r1 * r1;
                if (sf & 256u) a = a + r2 * r2;
                if (sf & 512u) a = a + r3 * r3;
                acc = acc + a;
            }
        } else {
            // frestrict_centers3d (:32-60): the coarse cell under columns (e+1, e+2), rows (A, B), planes
            // (q, q+1); the eight residuals are added in the order of the Fortran loop
            const double r0n = shfl_dn1(r0), r2n = shfl_dn1(r2);
            if (FAST || q < c.k1) {
                if (((q - NH) & 1) == 0) {
                    rsum = r1 + r0n; rsum = rsum + r3; rsum = rsum + r2n;
                } else {
                    rsum = rsum + r1; rsum = rsum + r0n; rsum = rsum + r3; rsum = rsum + r2n;
                    if (c.sf & 128u) c.bc[(unsigned)(NH + ((q - 1 - NH) >> 1)) * c.csk32 + c.bcA] = 0.5 * rsum;
                }
            }
        }
    }
}

// Which tiles of the plane a launch covers.  mode 0: all of them, (blockIdx.x, blockIdx.y) is the tile.
// mode 1: the rectangle [bx0, bx0+nbx) x [by0, by0+nby), grid (nbx, nby, .).  mode 2: the frame around that
// rectangle, tiles enumerated along blockIdx.x (rows above, the two side strips, rows below), grid (n, 1, .).
struct TileMap { int mode, bx0, by0, nbx, nby, gx; };

__device__ __forceinline__ void tile_of(const TileMap& m, int& bx, int& by)
{
    if (m.mode == 0) { bx = blockIdx.x; by = blockIdx.y; return; }
    if (m.mode == 1) { bx = m.bx0 + blockIdx.x; by = m.by0 + blockIdx.y; return; }
    int idx = blockIdx.x;
    const int top = m.by0 * m.gx, side = m.gx - m.nbx;
    if (idx < top) { by = idx / m.gx; bx = idx - by * m.gx; return; }
    idx -= top;
    if (idx < m.nby * side) {
        const int row = idx / side, r = idx - row * side;
        by = m.by0 + row; bx = r < m.bx0 ? r : r + m.nbx;
        return;
    }
    idx -= m.nby * side;
    by = m.by0 + m.nby + idx / m.gx; bx = idx % m.gx;
}

template <bool PRO, int POST, bool INT>
__global__ void __launch_bounds__(VL_NW * 32, 1)
k_vleg(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmb,
       const __grid_constant__ CUtensorMap tmc, double* __restrict__ xo, double* __restrict__ bc,
       double* __restrict__ partial, Box g, Box gc, double omega, double cff1, int kchunk, int kz0, int kz1,
       TileMap tm)
{
    // the launch covers the interior planes [kz0, kz1) (0-based) of the level in chunks of kchunk planes
    using LY = VlegLayout<PRO, POST>;
    constexpr int APT = LY::apron_top, TJ = LY::tj;
    constexpr int NPOST = POST != POST_NONE ? 1 : 0;
    extern __shared__ __align__(1024) unsigned char smem[];
    double* const s_recip = reinterpret_cast<double*>(smem + LY::off_misc);
    double* const s_pcoef = s_recip + 8;
    double* const s_red = s_pcoef + 4;
    uint64_t* const full = reinterpret_cast<uint64_t*>(s_red + 16);
    double* const sx = reinterpret_cast<double*>(smem + LY::off_x);
    double* const sb = reinterpret_cast<double*>(smem + LY::off_b);
    unsigned char* const sc = smem + LY::off_c;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = 2 * lane, rA = 2 * warp;
    int tbx, tby;
    tile_of(tm, tbx, tby);
    const int RX0 = tbx * VL_TI;                                        // array column of region column 0 (even)
    const int RY0 = NH + tby * TJ - APT;                                // array row of region row 0
    const int k0 = NH + kz0 + (int)blockIdx.z * kchunk, k1 = min(k0 + kchunk, NH + kz1);   // stored planes [k0, k1)
    // first iteration that runs sweep 1 / sweep 2 / the residual, first and last iteration, first plane loaded
    const int pb1 = k0 - 1 - NPOST, pb2 = k0 + 1 - NPOST, pb3 = k0 + 2;
    const int pstart = pb1 - 2, pend = k1 + NPOST;
    const int pl0 = pstart + 1, nplanes = pend - pstart + 1;

    if (threadIdx.x < 8) s_recip[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    if (threadIdx.x < 4) s_pcoef[threadIdx.x] = pcoef_of(threadIdx.x);
    if (threadIdx.x == 0) {
        for (int s = 0; s < VL_S; s++) nytma::mbar_init(&full[s], 1);
        nytma::fence_barrier_init();
    }
    __syncthreads();

    // first coarse column under the region; a box must start on a 16-byte boundary, i.e. on an even column
    const int CX0 = RX0 / 2 + 1;
    // ---- producer: plane t of this chunk = array plane pl0 + t -> stage t % VL_S.  Stateless (the coarse planes a
    // fine plane brings along follow from t alone), so that the duty can rotate over the warps: whoever issues is
    // late for the plane it is working on, and a fixed producer would hold up the block barrier of every plane.
    const int cfirst = PRO ? NH + ((pl0 - NH) >> 1) - 1 : 0;            // first coarse plane needed
    auto issue = [&](int t) {
        const int f = pl0 + t, s = t & (VL_S - 1);
        uint32_t bytes = 2 * VL_PLANE * 8;
        int c_lo = 0, c_hi = -1;
        if (PRO) {
            c_hi = NH + ((f - NH) >> 1) + 1;
            c_lo = t == 0 ? cfirst : NH + ((f - 1 - NH) >> 1) + 2;
            if (c_hi >= c_lo) bytes += (uint32_t)(c_hi - c_lo + 1) * VL_CBYTES;
        }
        nytma::mbar_expect_tx(&full[s], bytes);
        nytma::load_3d(sx + s * VL_PLANE, &tmx, RX0, RY0, f, &full[s]);
        nytma::load_3d(sb + s * VL_PLANE, &tmb, RX0, RY0, f, &full[s]);
        if (PRO) {
            for (int cp = c_lo; cp <= c_hi; cp++)
                nytma::load_3d(sc + (cp % VL_SC) * VL_CSLOT, &tmc, CX0 & ~1, RY0 / 2 + 1, cp, &full[s]);
        }
        // the planes after the next ones are pulled into L2 meanwhile
        if (t + VL_L2AHEAD < nplanes) {
            nytma::prefetch_3d(&tmx, RX0, RY0, f + VL_L2AHEAD);
            nytma::prefetch_3d(&tmb, RX0, RY0, f + VL_L2AHEAD);
        }
    };
    if (threadIdx.x == 0) {
        nytma::prefetch_map(&tmx);
        nytma::prefetch_map(&tmb);
        if (PRO) nytma::prefetch_map(&tmc);
        issue(0);
        if (nplanes > 1) issue(1);
    }

    // ---- per-thread constants ------------------------------------------------------------------------
    VlegCtx c;
    c.sx = sx; c.sb = sb; c.sc = sc; c.s_recip = s_recip; c.s_pcoef = s_pcoef;
    c.sy = reinterpret_cast<double*>(smem + LY::off_y);
    c.sz = reinterpret_cast<double*>(smem + LY::off_z);
    c.xo = xo; c.bc = bc; c.g = g; c.gc = gc; c.omega = omega; c.cff1 = cff1;
    c.ai = RX0 + e; c.ajA = RY0 + rA;
    const int ai = c.ai, ajA = c.ajA, ajB = ajA + 1;
    c.cA0 = cnt_xy(g, ai, ajA); c.cA1 = cnt_xy(g, ai + 1, ajA); c.cB0 = cnt_xy(g, ai, ajB); c.cB1 = cnt_xy(g, ai + 1, ajB);
    c.rA0 = c.cA0 >= 0 ? 1.0 / (double)(c.cA0 + 2) : 0.0; c.rA1 = c.cA1 >= 0 ? 1.0 / (double)(c.cA1 + 2) : 0.0;
    c.rB0 = c.cB0 >= 0 ? 1.0 / (double)(c.cB0 + 2) : 0.0; c.rB1 = c.cB1 >= 0 ? 1.0 / (double)(c.cB1 + 2) : 0.0;
    {
        const bool oc0 = e >= 3 && e < 3 + VL_TI && ai < g.nx + NH, oc1 = e + 1 >= 3 && e + 1 < 3 + VL_TI && ai + 1 < g.nx + NH;
        const bool orA = rA >= APT && rA < APT + TJ && ajA < g.ny + NH, orB = rA + 1 >= APT && rA + 1 < APT + TJ && ajB < g.ny + NH;
        unsigned sf = 0;
        if (orA) sf |= (oc0 && oc1) ? 1u : oc0 ? 2u : oc1 ? 4u : 0u;
        if (orB) sf |= (oc0 && oc1) ? 8u : oc0 ? 16u : oc1 ? 32u : 0u;
        if (orA && oc0) sf |= 64u;
        if (orA && oc1) sf |= 128u;
        if (orB && oc0) sf |= 256u;
        if (orB && oc1) sf |= 512u;
        unsigned gA = (unsigned)((long long)ajA * g.sj + ai), bcA = 0;
        // the coarse cell under columns (e+1, e+2) and rows (A, B); only threads with bit 7 set ever use it
        if (POST == POST_RESTRICT && (sf & 128u))
            bcA = (unsigned)((long long)(NH + ((ajA - NH) >> 1)) * gc.sj + (NH + ((ai + 1 - NH) >> 1)));
        asm volatile("mov.b32 %0, %0;" : "+r"(sf));
        asm volatile("mov.b32 %0, %0;" : "+r"(gA));
        asm volatile("mov.b32 %0, %0;" : "+r"(bcA));
        c.sf = sf; c.gA = gA; c.bcA = bcA;
        c.sk32 = (unsigned)g.sk; c.csk32 = (unsigned)gc.sk;
    }
    c.oA = rA * VL_RI + e;
    c.oUp = (rA > 0 ? rA - 1 : 0) * VL_RI + e; c.oDn = (rA + 2 < VL_RJ ? rA + 2 : VL_RJ - 1) * VL_RI + e;
    c.exA0 = c.exA1 = c.exB0 = c.exB1 = 0;
    if (PRO) {
        // coarse tile rows (warp, warp+1) x columns (lane, lane+1); in-domain flags of the "other" coarse
        // column / row of each fine cell (operators.f90:395-424 for the default mask)
        const int cx = CX0 + lane, cy = RY0 / 2 + 1 + warp;
        const int x0 = (int)in_x(gc, cx), x1 = (int)in_x(gc, cx + 1), y0 = (int)in_y(gc, cy), y1 = (int)in_y(gc, cy + 1);
        c.exA0 = x1 + y1; c.exA1 = x0 + y1; c.exB0 = x1 + y0; c.exB1 = x0 + y0;
    }
    c.oC = warp * VL_CI + lane + (CX0 & 1);
    c.k0 = k0; c.k1 = k1; c.pb1 = pb1; c.pb2 = pb2; c.pb3 = pb3;

    // planes whose two k-neighbours are inside the domain: [zlo1, zhi1]
    const int zlo1 = g.zlo ? -(1 << 20) : NH + 1, zhi1 = g.zhi ? (1 << 20) : g.nz - NH - 2;
    // all stages active, planes p-2 .. p (and p+1 and its coarse partner for the prolongation) regular,
    // stored plane inside the chunk
    const int fast_lo = max(POST != POST_NONE ? pb3 : pb2, zlo1 + 2), fast_hi = min(k1, zhi1 - (PRO ? 2 : 0));

    const double2 zero2 = make_double2(0.0, 0.0);
    P4 X0 = {zero2, zero2}, X1 = X0, X2 = X0, Y0 = X0, Y1 = X0, Y2 = X0, Z0 = X0, Z1 = X0, Z2 = X0, B0 = X0, B1 = X0, B2 = X0;
    double acc = 0.0, rsum = 0.0;
    int p = pstart;

#define VLEG_STEP(m, c_, n)                                                                                         \
    {                                                                                                               \
        if (p > pend) break;                                                                                        \
        const int t = p - pstart;                                                                                   \
        nytma::mbar_wait(&full[t & (VL_S - 1)], (uint32_t)(t / VL_S) & 1u);                                         \
        __syncthreads(); /* plane t landed; everything written in the last iteration is visible */                  \
        if (lane == 0 && warp == (t & (VL_NW - 1)) && t + 2 < nplanes) issue(t + 2);                                \
        if (p >= fast_lo && p <= fast_hi)                                                                           \
            vleg_plane<PRO, POST, true, INT>(c, p, t, X##m, X##c_, X##n, Y##m, Y##c_, Y##n, Z##m, Z##c_, Z##n, B##n, B##c_, B##m, acc, rsum); \
        else                                                                                                        \
            vleg_plane<PRO, POST, false, INT>(c, p, t, X##m, X##c_, X##n, Y##m, Y##c_, Y##n, Z##m, Z##c_, Z##n, B##n, B##c_, B##m, acc, rsum); \
        p++;                                                                                                        \
    }
    for (;;) {
        VLEG_STEP(0, 1, 2)
        VLEG_STEP(1, 2, 0)
        VLEG_STEP(2, 0, 1)
    }
#undef VLEG_STEP

    if (POST == POST_NORM) {
        for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tsum = 0.0;
            for (int w = 0; w < VL_NW; w++) tsum = tsum + s_red[w];
            partial[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tsum;
        }
    }
}
