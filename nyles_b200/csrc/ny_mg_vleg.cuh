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
    static constexpr int bytes = off_misc + 256;               // recip[8], red[16], full[VL_S]
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

template <bool PRO, int POST>
__global__ void __launch_bounds__(VL_NW * 32, 1)
k_vleg(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmb,
       const __grid_constant__ CUtensorMap tmc, double* __restrict__ xo, double* __restrict__ bc,
       double* __restrict__ partial, Box g, Box gc, double omega, double cff1, int kchunk)
{
    using LY = VlegLayout<PRO, POST>;
    constexpr int APT = LY::apron_top, TJ = LY::tj;
    constexpr int NPOST = POST != POST_NONE ? 1 : 0;
    extern __shared__ __align__(1024) unsigned char smem[];
    double* const sx = reinterpret_cast<double*>(smem + LY::off_x);
    double* const sb = reinterpret_cast<double*>(smem + LY::off_b);
    double* const sy = reinterpret_cast<double*>(smem + LY::off_y);
    double* const sz = reinterpret_cast<double*>(smem + LY::off_z);
    unsigned char* const sc = smem + LY::off_c;
    double* const s_recip = reinterpret_cast<double*>(smem + LY::off_misc);
    double* const s_red = s_recip + 8;
    uint64_t* const full = reinterpret_cast<uint64_t*>(s_red + 16);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = 2 * lane, rA = 2 * warp;
    const int RX0 = (int)blockIdx.x * VL_TI;                            // array column of region column 0 (even)
    const int RY0 = NH + (int)blockIdx.y * TJ - APT;                    // array row of region row 0
    const int ai = RX0 + e, ajA = RY0 + rA, ajB = ajA + 1;
    const int k0 = NH + (int)blockIdx.z * kchunk, k1 = min(k0 + kchunk, g.nz - NH);   // stored planes [k0, k1)
    // first iteration that runs sweep 1 / sweep 2 / the residual, first and last iteration, first plane loaded
    const int pb1 = k0 - 1 - NPOST, pb2 = k0 + 1 - NPOST, pb3 = k0 + 2;
    const int pstart = pb1 - 2, pend = k1 + NPOST;
    const int pl0 = pstart + 1, nplanes = pend - pstart + 1;

    if (threadIdx.x < 8) s_recip[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < VL_S; s++) nytma::mbar_init(&full[s], 1);
        nytma::fence_barrier_init();
    }
    __syncthreads();

    // first coarse column under the region; a box must start on a 16-byte boundary, i.e. on an even column
    const int CX0 = RX0 / 2 + 1;
    // ---- producer state (thread 0 only): plane t of this chunk = array plane pl0 + t -> stage t % VL_S
    int cnext = 0;
    if (PRO) cnext = NH + ((pl0 - NH) >> 1) - 1;                         // first coarse plane needed
    auto issue = [&](int t) {
        const int f = pl0 + t, s = t & (VL_S - 1);
        uint32_t bytes = 2 * VL_PLANE * 8;
        int mneed = -1;
        if (PRO) {
            mneed = NH + ((f - NH) >> 1) + 1;
            if (mneed >= cnext) bytes += (uint32_t)(mneed - cnext + 1) * VL_CBYTES;
        }
        nytma::mbar_expect_tx(&full[s], bytes);
        nytma::load_3d(sx + s * VL_PLANE, &tmx, RX0, RY0, f, &full[s]);
        nytma::load_3d(sb + s * VL_PLANE, &tmb, RX0, RY0, f, &full[s]);
        if (PRO) {
            for (; cnext <= mneed; cnext++)
                nytma::load_3d(sc + (cnext % VL_SC) * VL_CSLOT, &tmc, CX0 & ~1, RY0 / 2 + 1, cnext, &full[s]);
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
    // in-plane neighbour counts of the four cells (negative outside the domain), 1/diag for two k-neighbours
    const int cA0 = cnt_xy(g, ai, ajA), cA1 = cnt_xy(g, ai + 1, ajA), cB0 = cnt_xy(g, ai, ajB), cB1 = cnt_xy(g, ai + 1, ajB);
    const double rA0 = cA0 >= 0 ? 1.0 / (double)(cA0 + 2) : 0.0, rA1 = cA1 >= 0 ? 1.0 / (double)(cA1 + 2) : 0.0;
    const double rB0 = cB0 >= 0 ? 1.0 / (double)(cB0 + 2) : 0.0, rB1 = cB1 >= 0 ? 1.0 / (double)(cB1 + 2) : 0.0;
    // cells of the stored tile that belong to the interior of the array
    const bool oc0 = e >= 3 && e < 3 + VL_TI && ai < g.nx + NH, oc1 = e + 1 >= 3 && e + 1 < 3 + VL_TI && ai + 1 < g.nx + NH;
    const bool orA = rA >= APT && rA < APT + TJ && ajA < g.ny + NH, orB = rA + 1 >= APT && rA + 1 < APT + TJ && ajB < g.ny + NH;
    // shared-memory offsets inside a plane tile
    const int oA = rA * VL_RI + e, oB = oA + VL_RI;
    const int oUp = (rA > 0 ? rA - 1 : 0) * VL_RI + e, oDn = (rA + 2 < VL_RJ ? rA + 2 : VL_RJ - 1) * VL_RI + e;
    const long long gA = (long long)ajA * g.sj + ai, gB = gA + g.sj;
    // prolongation: coarse tile rows (warp, warp+1) x columns (lane, lane+1); in-domain flags of the
    // "other" coarse column / row of each fine cell (operators.f90:395-424 for the default mask)
    int exA0 = 0, exA1 = 0, exB0 = 0, exB1 = 0;
    if (PRO) {
        const int cx = CX0 + lane, cy = RY0 / 2 + 1 + warp;
        const int x0 = (int)in_x(gc, cx), x1 = (int)in_x(gc, cx + 1), y0 = (int)in_y(gc, cy), y1 = (int)in_y(gc, cy + 1);
        exA0 = x1 + y1; exA1 = x0 + y1; exB0 = x1 + y0; exB1 = x0 + y0;
    }
    const int oC = warp * VL_CI + lane + (CX0 & 1);

    const double2 zero2 = make_double2(0.0, 0.0);
    double2 xmA = zero2, xmB = zero2, xcA = zero2, xcB = zero2;
    double2 ymA = zero2, ymB = zero2, ycA = zero2, ycB = zero2;
    double2 zmA = zero2, zmB = zero2, zcA = zero2, zcB = zero2;
    double2 b1A = zero2, b1B = zero2, b2A = zero2, b2B = zero2;
    double acc = 0.0, rsum = 0.0;

    auto coef = [&](int cz, double& a0, double& a1, double& b0, double& b1) {
        if (cz == 2) { a0 = rA0; a1 = rA1; b0 = rB0; b1 = rB1; }
        else {                                                  // planes next to the k ends (CTA-uniform)
            a0 = s_recip[max(cA0 + cz, 0)]; a1 = s_recip[max(cA1 + cz, 0)];
            b0 = s_recip[max(cB0 + cz, 0)]; b1 = s_recip[max(cB1 + cz, 0)];
        }
    };

    for (int p = pstart; p <= pend; p++) {
        const int t = p - pstart;
        nytma::mbar_wait(&full[t & (VL_S - 1)], (uint32_t)(t / VL_S) & 1u);
        __syncthreads();                // plane t landed; everything written in the last iteration is visible
        if (threadIdx.x == 0 && t + 2 < nplanes) issue(t + 2);
        double* const px = sx + (t & (VL_S - 1)) * VL_PLANE;                       // x[p+1]
        const double* const pxc = sx + ((t + VL_S - 1) & (VL_S - 1)) * VL_PLANE;   // x[p]
        const double* const pbc = sb + ((t + VL_S - 1) & (VL_S - 1)) * VL_PLANE;   // b[p]
        double2 xpA = ld2(px + oA), xpB = ld2(px + oB);
        if (PRO) {
            // x[p+1] += Pcoef * (27 9 9 3 / 9 3 3 1 weights of the eight nearest coarse cells), basicoperators.f90:173-231
            const int f = p + 1, kf = f - NH, akc = NH + (kf >> 1), ako = (kf & 1) ? akc + 1 : akc - 1;
            const double* cb = reinterpret_cast<const double*>(sc + (akc % VL_SC) * VL_CSLOT) + oC;
            const double* co = reinterpret_cast<const double*>(sc + (ako % VL_SC) * VL_CSLOT) + oC;
            const double b00 = cb[0], b01 = cb[1], b10 = cb[VL_CI], b11 = cb[VL_CI + 1];
            const double o00 = co[0], o01 = co[1], o10 = co[VL_CI], o11 = co[VL_CI + 1];
            const int ez = (int)in_z(gc, ako);
            const bool fz = in_z(g, f);
            const double pbA0 = 9 * b00 + 3 * b01 + 3 * b10 + b11, poA0 = 9 * o00 + 3 * o01 + 3 * o10 + o11;
            const double pbA1 = 9 * b01 + 3 * b00 + 3 * b11 + b10, poA1 = 9 * o01 + 3 * o00 + 3 * o11 + o10;
            const double pbB0 = 9 * b10 + 3 * b11 + 3 * b00 + b01, poB0 = 9 * o10 + 3 * o11 + 3 * o00 + o01;
            const double pbB1 = 9 * b11 + 3 * b10 + 3 * b01 + b00, poB1 = 9 * o11 + 3 * o10 + 3 * o01 + o00;
            if (fz && cA0 >= 0) xpA.x = xpA.x + pcoef_of(exA0 + ez) * (3 * pbA0 + poA0);
            if (fz && cA1 >= 0) xpA.y = xpA.y + pcoef_of(exA1 + ez) * (3 * pbA1 + poA1);
            if (fz && cB0 >= 0) xpB.x = xpB.x + pcoef_of(exB0 + ez) * (3 * pbB0 + poB0);
            if (fz && cB1 >= 0) xpB.y = xpB.y + pcoef_of(exB1 + ez) * (3 * pbB1 + poB1);
            st2(px + oA, xpA);          // the rows above / below read the prolonged plane in the next iteration
            st2(px + oB, xpB);
            nytma::fence_proxy_async();
        }
        double2 ypA = zero2, ypB = zero2, zpA = zero2, zpB = zero2, bpA = zero2, bpB = zero2;
        if (p >= pb1) {                 // ---- sweep 1 at plane p (fsmoother3d, basicoperators.f90:363-400)
            bpA = ld2(pbc + oA); bpB = ld2(pbc + oB);
            double iA0, iA1, iB0, iB1;
            coef(cnt_z(g, p), iA0, iA1, iB0, iB1);
            sweep_patch(xcA, xcB, xmA, xmB, xpA, xpB, ld2(pxc + oUp), ld2(pxc + oDn), bpA, bpB,
                        iA0, iA1, iB0, iB1, omega, cff1, ypA, ypB);
            double* const py = sy + (p & 1) * VL_PLANE;
            st2(py + oA, ypA);
            st2(py + oB, ypB);
        }
        if (p >= pb2) {                 // ---- sweep 2 at plane q = p-1
            const int q = p - 1;
            const double* const py = sy + (q & 1) * VL_PLANE;
            double iA0, iA1, iB0, iB1;
            coef(cnt_z(g, q), iA0, iA1, iB0, iB1);
            sweep_patch(ycA, ycB, ymA, ymB, ypA, ypB, ld2(py + oUp), ld2(py + oDn), b1A, b1B,
                        iA0, iA1, iB0, iB1, omega, cff1, zpA, zpB);
            const bool qz = in_z(g, q);
            if (!(qz && cA0 >= 0)) zpA.x = xmA.x;            // outside the domain: x is kept
            if (!(qz && cA1 >= 0)) zpA.y = xmA.y;
            if (!(qz && cB0 >= 0)) zpB.x = xmB.x;
            if (!(qz && cB1 >= 0)) zpB.y = xmB.y;
            if (POST != POST_NONE) {
                double* const pz = sz + (q & 1) * VL_PLANE;
                st2(pz + oA, zpA);
                st2(pz + oB, zpB);
            }
            if (q >= k0 && q < k1) {
                double* const dA = xo + (long long)q * g.sk + gA;
                double* const dB = xo + (long long)q * g.sk + gB;
                if (orA) {
                    if (oc0 && oc1) st2(dA, zpA);
                    else if (oc0) dA[0] = zpA.x;
                    else if (oc1) dA[1] = zpA.y;
                }
                if (orB) {
                    if (oc0 && oc1) st2(dB, zpB);
                    else if (oc0) dB[0] = zpB.x;
                    else if (oc1) dB[1] = zpB.y;
                }
            }
        }
        if (POST != POST_NONE && p >= pb3) {     // ---- residual at plane q = p-2 (fresidual3d, :300-323)
            const int q = p - 2;
            const double* const pz = sz + (q & 1) * VL_PLANE;
            const double2 up = ld2(pz + oUp), dn = ld2(pz + oDn);
            const double lA = shfl_up1(zcA.y), rgA = shfl_dn1(zcA.x), lB = shfl_up1(zcB.y), rgB = shfl_dn1(zcB.x);
            const double sA0 = lA + zcA.y + up.x + zcB.x + zmA.x + zpA.x;
            const double sA1 = zcA.x + rgA + up.y + zcB.y + zmA.y + zpA.y;
            const double sB0 = lB + zcB.y + zcA.x + dn.x + zmB.x + zpB.x;
            const double sB1 = zcB.x + rgB + zcA.y + dn.y + zmB.y + zpB.y;
            const int cz = cnt_z(g, q);
            const double r0 = b2A.x + (double)(cA0 + cz) * zcA.x - sA0;
            const double r1 = b2A.y + (double)(cA1 + cz) * zcA.y - sA1;
            const double r2 = b2B.x + (double)(cB0 + cz) * zcB.x - sB0;
            const double r3 = b2B.y + (double)(cB1 + cz) * zcB.y - sB1;
            if (POST == POST_NORM) {
                if (q < k1) {           // fnorm, basicoperators.f90:422-440 (interior cells, msk = 1)
                    double a = 0.0;
                    if (orA && oc0) a = a + r0 * r0;
                    if (orA && oc1) a = a + r1 * r1;
                    if (orB && oc0) a = a + r2 * r2;
                    if (orB && oc1) a = a + r3 * r3;
                    acc = acc + a;
                }
            } else {
                // frestrict_centers3d (:32-60): the coarse cell under columns (e+1, e+2), rows (A, B), planes
                // (q, q+1); the eight residuals are added in the order of the Fortran loop
                const double r0n = shfl_dn1(r0), r2n = shfl_dn1(r2);
                if (q < k1) {
                    if (((q - NH) & 1) == 0) {
                        rsum = r1 + r0n; rsum = rsum + r3; rsum = rsum + r2n;
                    } else {
                        rsum = rsum + r1; rsum = rsum + r0n; rsum = rsum + r3; rsum = rsum + r2n;
                        if (orA && oc1)
                            bc[(long long)(NH + ((q - 1 - NH) >> 1)) * gc.sk + (long long)(NH + ((ajA - NH) >> 1)) * gc.sj +
                               (NH + ((ai + 1 - NH) >> 1))] = 0.5 * rsum;
                    }
                }
            }
        }
        xmA = xcA; xmB = xcB; xcA = xpA; xcB = xpB;
        ymA = ycA; ymB = ycB; ycA = ypA; ycB = ypB;
        zmA = zcA; zmB = zcB; zcA = zpA; zcB = zpB;
        b2A = b1A; b2B = b1B; b1A = bpA; b1B = bpB;
    }
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
