// Geometric multigrid pressure solver: sm_100a replacement of libmgmod64.so (core/mgfor/*.f90).
//
// Same algorithm, same iterates as the reference (required for field parity at the loose
// stopping tolerance, SURVEY.md section 7 hard part 1):
//   hierarchy          mg_setup.f90:225-307      masks/Rcoef/Pcoef/diag   operators.f90:246-505
//   smoother           basicoperators.f90:363-400 (two damped-Jacobi sweeps, omega = 0.9)
//   residual           basicoperators.f90:300-323  restriction  :32-60   prolongation :173-231
//   norm               basicoperators.f90:422-440  V-cycle / solve  solvers.f90:8-55
//   halo fill          mod_halo.f90:200-262
// Level arrays are (nz+2nh, ny+2nh, nx+2nh), i fastest, exactly the reference's padded layout.
// Arithmetic: source order, no FMA (-fmad=false) -> bit-identical to the oracle except for the
// summation order of the two norms, which only feed the stopping test.
//
// Two code paths, both CUDA:
//   * generic : one kernel per Fortran loop nest, coefficient arrays (msk, diag, idiag, Rcoef,
//               Pcoef) read from memory.  Used for the one-time operator setup (which follows
//               operators.f90:299-457 literally) and whenever the mask is not the default box.
//   * box     : the hot path.  For box domains every coefficient is a function of the cell
//               position, so nothing but x and b is read: the two Jacobi sweeps are fused in one
//               pass (k-streaming tiles, z neighbours in registers, in-plane neighbours through
//               shared memory), residual+restriction and residual+norm never materialise r.
//               The analytic coefficients are verified against the arrays at creation.
// Multi-GPU: z slabs (one rank per GPU).  Distributed levels exchange 3-plane z faces through
// NCCL; once a level is small it is gathered (ncclAllGather) and all coarser levels are solved
// redundantly on every rank (mirror of mg_setup.f90:275-293 / mod_gluesplit.f90).
#include "ny_common.cuh"
#include <cstdlib>
#include "ny_comm.cuh"
#include "ny_tma.cuh"

namespace {

constexpr int MAXLEV = 50;
constexpr int NH = 3;
constexpr int MAX_PARTIALS = 1 << 15;
// levels with at most this many cells are gathered: 128^3 (and a little: the test is <=).  64^3 in round 1; at 8 B200
// the two face exchanges per V-cycle of a distributed 128^3 level cost more than solving it redundantly
// (64.6 -> 64.1 ms per 1024^3 step, profiles/r2_f_*)
long long g_gather_cells = 2200000;
long long g_tail_cells = 2048;            // levels with at most this many cells form the one-launch tail of the V-cycle (one CTA)
// ... and the replicated levels above them with at most this many cells join that launch as its "wide" levels, run
// by a co-resident grid with grid barriers between the operators (0: none): 64^3 and a little.  Measured
// (profiles/r2_l_*): 64^3 .. 16^3 levels and the whole 128 x 32 x 32 lock-exchange grid are faster wide than through
// the fused legs, which stay with the levels from 128^3 up.
long long g_wide_cells = 300000;
// levels with fewer cells than this run one kernel per operator instead of the fused plane-marching legs, whose
// pipeline (6 planes of fill per chunk, ~1.4 us per plane) is latency bound on small levels
long long g_leg_min_cells = 0;
long long g_split_tiles = 148;            // levels with at least this many 58 x 24 tiles launch wall-free tiles separately
// slab levels with at least this many local cells compute the planes next to their neighbours first and overlap
// the exchange with the rest.  Off by default: with the peer-memory exchange a 1024^2 x 3 face takes 60 us, less
// than the two extra pipeline fills of the split launches cost (8 B200, 1024^3: 66.2 ms per step with the
// overlap, 64.4 without; profiles/r1_p_*).  ny_mg_set_overlap_cells / NY_MG_OVERLAP_CELLS turn it on.
long long g_overlap_cells = 1LL << 62;

struct Level {
    int nx, ny, nz;                        // nz = local planes including the 2*nh halo planes
    long long sj, sk;
    size_t n;
    int zlo, zhi;                          // the domain continues below / above the local array (periodic or slab neighbour)
    int gathered;                          // replicated on every rank (always true on one rank)
    double *x, *b, *r, *y, *diag, *idiag, *Rcoef, *Pcoef, *msk;
};

// what the box kernels need to know about a level array (or a window of one)
struct Box {
    int nx, ny, nz;                        // interior extents in x, y; padded plane count in z
    long long sj, sk;
    int xper, yper, zlo, zhi;
};

}  // namespace

// TMA tensor maps of a level: plane tiles of x, y (the two smoother buffers, swapped together with the
// pointers) and b; cx / cy: the coarse tile of x / y as the next finer level sees this level (the
// whole array, or the window of a gathered level that lies under the rank's slab)
struct LevelMaps { CUtensorMap x, y, b, cx, cy; int ok; };

struct ny_mg {
    ny_ctx* ctx;
    ny_comm* comm;
    LevelMaps maps[50];
    int fused;                             // use the fused V-cycle legs where a level allows them
    int split_tiles;                       // fused legs: wall-free tiles run the specialised kernel instance
    int tail;                              // closed boxes: the smallest levels of a V-cycle run in one launch
    int halo_ok;                           // periodic / slab halos of x and b of level 1 are consistent
    int ysync;                             // wall-halo cells of y equal those of x on every level
    int nranks, rank;
    int nlevels, nh, topology, maxite;
    int xper, yper, zper;
    int box;                               // default mask: analytic coefficients are valid
    int glev;                              // first gathered level (0-based); 0 on one rank
    // tuning, fixed when the multigrid is created (copied from the process-wide defaults that the ny_mg_set_* calls
    // and the NY_MG_* environment variables set; a setter never changes an existing multigrid)
    long long tail_cells, wide_cells, split_tiles_min, overlap_cells, leg_min_cells;
    int wide_max_blocks;                   // CTAs of k_vcycle_tail that a cooperative launch can hold (0: no wide levels)
    unsigned* d_bar;                       // grid barrier of the wide levels: arrivals, generation
    int below, above;                      // slab neighbours (-1: none)
    double tol, omega;
    Level lev[MAXLEV];
    double* tmp;                           // staging for overlapping halo sections
    size_t tmp_doubles;
    double* d_red;                         // MAX_PARTIALS partials + 4 results
    int* d_flag;
};

namespace {

__host__ __device__ inline long long IX(const Level& L, int nh, int i, int j, int k)
{   // Fortran indices (i in 1-nh..nx+nh, j likewise, k in 1..nz) -> linear offset
    return (long long)(k - 1) * L.sk + (long long)(j - 1 + nh) * L.sj + (i - 1 + nh);
}

inline Box box_of(const ny_mg* mg, const Level& L)
{
    Box g; g.nx = L.nx; g.ny = L.ny; g.nz = L.nz; g.sj = L.sj; g.sk = L.sk;
    g.xper = mg->xper; g.yper = mg->yper; g.zlo = L.zlo; g.zhi = L.zhi;
    return g;
}

// ---- analytic coefficients of a box domain (array indices, 0-based) --------------------------
__device__ __forceinline__ bool in_x(const Box& g, int ai) { return g.xper || (ai >= NH && ai < g.nx + NH); }
__device__ __forceinline__ bool in_y(const Box& g, int aj) { return g.yper || (aj >= NH && aj < g.ny + NH); }
__device__ __forceinline__ bool in_z(const Box& g, int ak) { return (ak >= NH || g.zlo) && (ak < g.nz - NH || g.zhi); }
// number of in-domain neighbours in the plane (the x,y part of diag = msk * sum_6 msk), or a large
// negative number if the column itself is outside the domain
__device__ __forceinline__ int cnt_xy(const Box& g, int ai, int aj)
{
    if (!(in_x(g, ai) && in_y(g, aj))) return -100;
    return (int)in_x(g, ai - 1) + (int)in_x(g, ai + 1) + (int)in_y(g, aj - 1) + (int)in_y(g, aj + 1);
}
__device__ __forceinline__ int cnt_z(const Box& g, int ak)
{
    if (!in_z(g, ak)) return -100;
    return (int)in_z(g, ak - 1) + (int)in_z(g, ak + 1);
}
// Pcoef of a fine cell whose three "other" coarse neighbours have e in-domain flags set
// (operators.f90:395-424 evaluated for the default mask: 1 / ((3+ex)(3+ey)(3+ez)))
__device__ __forceinline__ double pcoef_of(int e)
{
    return e == 3 ? 1.0 / 64.0 : (e == 2 ? 1.0 / 48.0 : (e == 1 ? 1.0 / 36.0 : 1.0 / 27.0));
}

// =================================================================================================
//  generic kernels (coefficient arrays)
// =================================================================================================
__global__ void __launch_bounds__(256)
k_sweep(const double* __restrict__ src, double* __restrict__ dst, const double* __restrict__ b,
        const double* __restrict__ idiag, double omega, double cff1, Level L, int nh,
        int i0, int i1, int j0, int j1, int k0, int k1)
{
    int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = j0 + blockIdx.y * blockDim.y + threadIdx.y;
    int k = k0 + blockIdx.z * blockDim.z + threadIdx.z;
    if (i > i1 || j > j1 || k > k1) return;
    long long c = IX(L, nh, i, j, k);
    double s = src[c - 1] + src[c + 1] + src[c - L.sj] + src[c + L.sj] + src[c - L.sk] + src[c + L.sk];
    dst[c] = cff1 * src[c] + omega * (s - b[c]) * idiag[c];
}

__global__ void __launch_bounds__(256)
k_residual(const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ r,
           const double* __restrict__ msk, const double* __restrict__ diag, Level L, int nh)
{
    int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    int k = 1 + nh + blockIdx.z * blockDim.z + threadIdx.z;
    if (i > L.nx || j > L.ny || k > L.nz - nh) return;
    long long c = IX(L, nh, i, j, k);
    double s = x[c - 1] + x[c + 1] + x[c - L.sj] + x[c + L.sj] + x[c - L.sk] + x[c + L.sk];
    r[c] = msk[c] * (b[c] + diag[c] * x[c] - s);
}

// coef == nullptr: box domain, Rcoef = 0.5 on every interior coarse cell
__global__ void __launch_bounds__(256)
k_restrict(const double* __restrict__ xf, double* __restrict__ xc, const double* __restrict__ coef,
           Level F, Level C, int nh)
{
    int ic = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int jc = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    int kc = 1 + nh + blockIdx.z * blockDim.z + threadIdx.z;
    if (ic > C.nx || jc > C.ny || kc > C.nz - nh) return;
    int i = 1 + (ic - 1) * 2, j = 1 + (jc - 1) * 2, k = 1 + nh + (kc - 1 - nh) * 2;
    long long f = IX(F, nh, i, j, k);
    double s = xf[f] + xf[f + 1] + xf[f + F.sj] + xf[f + F.sj + 1]
             + xf[f + F.sk] + xf[f + F.sk + 1] + xf[f + F.sk + F.sj] + xf[f + F.sk + F.sj + 1];
    long long c = IX(C, nh, ic, jc, kc);
    xc[c] = (coef ? coef[c] : 0.5) * s;
}

// one thread per fine cell; (i,j) parity picks the coarse neighbours, k parity picks plane a or c.
// coef == nullptr: box domain, analytic Pcoef.
__global__ void __launch_bounds__(256)
k_prolong(double* __restrict__ xf, const double* __restrict__ xc, const double* __restrict__ coef,
          Level F, Level C, Box gc, int nh)
{
    int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    int k = 1 + nh + blockIdx.z * blockDim.z + threadIdx.z;
    if (i > F.nx || j > F.ny || k > F.nz - nh) return;
    int ic = (i + 1) / 2, jc = (j + 1) / 2;
    int di = (i & 1) ? -1 : 1, dj = (j & 1) ? -1 : 1;
    int upper = (k - (1 + nh)) & 1;                // 0: cell k of the pair, 1: cell k+1
    int kb = k - upper;
    int kc = nh + (kb + 1 - nh) / 2;
    int ko = upper ? kc + 1 : kc - 1;
    long long cb = IX(C, nh, ic, jc, kc), co = IX(C, nh, ic, jc, ko);
    long long ox = di, oy = (long long)dj * C.sj;
    double pb = 9 * xc[cb] + 3 * xc[cb + ox] + 3 * xc[cb + oy] + xc[cb + ox + oy];
    double po = 9 * xc[co] + 3 * xc[co + ox] + 3 * xc[co + oy] + xc[co + ox + oy];
    long long f = IX(F, nh, i, j, k);
    double cf;
    if (coef) cf = coef[f];
    else cf = pcoef_of((int)in_x(gc, ic + di - 1 + nh) + (int)in_y(gc, jc + dj - 1 + nh) + (int)in_z(gc, ko - 1));
    xf[f] = xf[f] + cf * (3 * pb + po);
}

// sum(msk * x*x) over the interior: fixed grid-stride order => deterministic
__global__ void __launch_bounds__(256)
k_norm_partial(const double* __restrict__ msk, const double* __restrict__ x, Level L, int nh,
               double* __restrict__ partial)
{
    __shared__ double sh[8];
    const long long nzi = L.nz - 2 * nh;
    const long long total = nzi * L.ny * L.nx;
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t % L.nx) + 1;
        long long q = t / L.nx;
        int j = (int)(q % L.ny) + 1;
        int k = (int)(q / L.ny) + 1 + nh;
        long long c = IX(L, nh, i, j, k);
        double v = x[c];
        acc = acc + msk[c] * (v * v);
    }
    for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int w = 0; w < (blockDim.x >> 5); w++) r = r + sh[w];
        partial[blockIdx.x] = r;
    }
}
__global__ void __launch_bounds__(256)
k_sum_final(const double* __restrict__ partial, int nb, double* __restrict__ out)
{
    __shared__ double sh[256];
    double acc = 0.0;
    for (int t = threadIdx.x; t < nb; t += blockDim.x) acc = acc + partial[t];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// k_sum_final for gridDim.x sums at once (block s: partials [s nb, (s + 1) nb) -> out[gridDim.x - 1 - s]), each also
// stored into the host's pinned mailbox: the two norms that open a solve, ready for the stop test after one launch
__global__ void __launch_bounds__(256)
k_sum_final_store(const double* __restrict__ partial, int nb, double* __restrict__ out, double* host)
{
    __shared__ double sh[256];
    const double* p = partial + (long long)blockIdx.x * nb;
    double acc = 0.0;
    for (int t = threadIdx.x; t < nb; t += blockDim.x) acc = acc + p[t];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int slot = gridDim.x - 1 - blockIdx.x;
        out[slot] = sh[0];
        host[slot] = sh[0];
        __threadfence_system();
    }
}

// ---- array-section copy  dst = src  (Fortran indices), optionally staged through tmp ----------
__global__ void __launch_bounds__(256)
k_box_copy(const double* src, double* dst, Level L, int nh,
           int di0, int dj0, int dk0, int si0, int sj0, int sk0, int ni, int nj, int nk,
           int src_packed, int dst_packed)
{
    long long total = (long long)ni * nj * nk;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int i = (int)(t % ni);
        long long q = t / ni;
        int j = (int)(q % nj), k = (int)(q / nj);
        double v = src_packed ? src[t] : src[IX(L, nh, si0 + i, sj0 + j, sk0 + k)];
        if (dst_packed) dst[t] = v; else dst[IX(L, nh, di0 + i, dj0 + j, dk0 + k)] = v;
    }
}

// Whole periodic halo fill in one launch, valid when every wrapped axis is at least nh wide (then
// the statement sequence of mod_halo.f90:235-262 equals the rule below).  wz: wrap z too.
// Threads: segment 0 = the 2*nh z-halo planes, segment 1 = y-halo rows of the other planes,
// segment 2 = x-halo columns of the remaining rows.  x and y sections of the Fortran cover all k,
// the corner sections exist only when both x and y are periodic, the z sections copy full planes.
__global__ void __launch_bounds__(256)
k_fill_periodic(double* __restrict__ a, int nx, int ny, int nz, long long sj, long long sk,
                int xper, int yper, int wz, long long n0, long long n1, long long n2)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int ai, aj, ak;
    const int tx = nx + 2 * NH, ty = ny + 2 * NH;
    if (t < n0) {
        ai = (int)(t % tx); long long q = t / tx; aj = (int)(q % ty); int p = (int)(q / ty);
        ak = p < NH ? p : nz - 2 * NH + p;
    } else if (t < n0 + n1) {
        t -= n0;
        ai = (int)(t % tx); long long q = t / tx; int r = (int)(q % (2 * NH)); ak = NH + (int)(q / (2 * NH));
        aj = r < NH ? r : ny + r;
    } else if (t < n0 + n1 + n2) {
        t -= n0 + n1;
        int c = (int)(t % (2 * NH)); long long q = t / (2 * NH); aj = NH + (int)(q % ny); ak = NH + (int)(q / ny);
        ai = c < NH ? c : nx + c;
    } else return;
    const bool hx = ai < NH || ai >= nx + NH, hy = aj < NH || aj >= ny + NH, hz = ak < NH || ak >= nz - NH;
    int si = ai, sjj = aj, skk = ak;
    if (hz && wz) skk = ak < NH ? ak + (nz - 2 * NH) : ak - (nz - 2 * NH);
    if (hx && hy) {
        if (xper && yper) { si = ai < NH ? ai + nx : ai - nx; sjj = aj < NH ? aj + ny : aj - ny; }
    } else if (hx) {
        if (xper) si = ai < NH ? ai + nx : ai - nx;
    } else if (hy) {
        if (yper) sjj = aj < NH ? aj + ny : aj - ny;
    }
    if (si == ai && sjj == aj && skk == ak) return;
    a[(long long)ak * sk + (long long)aj * sj + ai] = a[(long long)skk * sk + (long long)sjj * sj + si];
}

// ---- elementwise helpers for the one-time operator setup -------------------------------------
enum { EW_SET, EW_COPY, EW_GT0_ONE, EW_MUL, EW_NEG, EW_RECIP_GT0 };
template <int OP>
__global__ void __launch_bounds__(256)
k_ew(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b, double val, long long n)
{
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (long long)gridDim.x * blockDim.x) {
        if (OP == EW_SET) dst[t] = val;
        else if (OP == EW_COPY) dst[t] = a[t];
        else if (OP == EW_GT0_ONE) dst[t] = (a[t] > 0.0) ? 1.0 : 0.0;
        else if (OP == EW_MUL) dst[t] = a[t] * b[t];
        else if (OP == EW_NEG) dst[t] = -a[t];
        else dst[t] = (a[t] > 0.0) ? val / a[t] : 0.0;          // where(a>0) dst = val/a, else 0
    }
}

// set_default_msk (mg_setup.f90:180-211): 1 on interior i,j and interior k (all k if z-periodic)
__global__ void __launch_bounds__(256)
k_default_msk(double* __restrict__ msk, Level L, int nh, int zper)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)L.n) return;
    int i = (int)(t % L.sj) + 1 - nh;
    long long q = t / L.sj;
    int j = (int)(q % (L.ny + 2 * nh)) + 1 - nh;
    int k = (int)(q / (L.ny + 2 * nh)) + 1;
    bool in = i >= 1 && i <= L.nx && j >= 1 && j <= L.ny && (zper || (k >= 1 + nh && k <= L.nz - nh));
    msk[t] = in ? 1.0 : 0.0;
}
// apply_default_msk (operators.f90:246-297): zero the mask on x/y sides that have no neighbour
__global__ void __launch_bounds__(256)
k_apply_default_msk(double* __restrict__ msk, Level L, int nh, int xper, int yper)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)L.n) return;
    int i = (int)(t % L.sj) + 1 - nh;
    int j = (int)((t / L.sj) % (L.ny + 2 * nh)) + 1 - nh;
    if ((!xper && (i <= 0 || i >= L.nx + 1)) || (!yper && (j <= 0 || j >= L.ny + 1))) msk[t] = 0.0;
}

// compare the analytic box coefficients with the arrays produced by the reference's setup
// algorithm; any mismatch raises the flag and the generic path stays in use
__global__ void __launch_bounds__(256)
k_check_box(Level L, Box g, int has_coarse, int is_coarse, Box gc, int* __restrict__ flag)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)L.n) return;
    int ai = (int)(t % L.sj);
    long long q = t / L.sj;
    int aj = (int)(q % (L.ny + 2 * NH)), ak = (int)(q / (L.ny + 2 * NH));
    // only what the box kernels rely on is compared: idiag on the sweep-1 domain (interior + 1 ring),
    // msk, diag, Rcoef, Pcoef on the interior (beyond ring 2 the halo of levels narrower than nh
    // holds stale copies, mod_halo.f90:235-262 with overlapping sections)
    int c = cnt_xy(g, ai, aj) + cnt_z(g, ak);
    const bool ring1 = ai >= NH - 1 && ai <= L.nx + NH && aj >= NH - 1 && aj <= L.ny + NH && ak >= NH - 1 && ak <= L.nz - NH;
    const bool interior = ai >= NH && ai < L.nx + NH && aj >= NH && aj < L.ny + NH && ak >= NH && ak < L.nz - NH;
    bool bad = ring1 && L.idiag[t] != (c > 0 ? 1.0 / (double)c : 0.0);
    if (interior && (L.msk[t] != 1.0 || L.diag[t] != (double)c)) bad = true;
    if (interior && is_coarse && L.Rcoef[t] != 0.5) bad = true;
    if (interior && has_coarse) {
        int i = ai - NH + 1, j = aj - NH + 1, k = ak + 1;
        int ic = (i + 1) / 2, jc = (j + 1) / 2;
        int di = (i & 1) ? -1 : 1, dj = (j & 1) ? -1 : 1;
        int upper = (k - (1 + NH)) & 1;
        int kc = NH + (k - upper + 1 - NH) / 2;
        int ko = upper ? kc + 1 : kc - 1;
        int e = (int)in_x(gc, ic + di - 1 + NH) + (int)in_y(gc, jc + dj - 1 + NH) + (int)in_z(gc, ko - 1);
        if (L.Pcoef[t] != pcoef_of(e)) bad = true;
    }
    if (bad) atomicOr(flag, 1);
}

// b_mg[idx] = div  /  p = x_mg[idx]*scale   (core/mgfordriver.py:72,78)
__global__ void __launch_bounds__(256)
k_embed(double* __restrict__ bmg, const double* __restrict__ div, Level L, int nz, int ny, int nx,
        int k0, int j0, int i0)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= nx || j >= ny || k >= nz) return;
    bmg[(long long)(k + k0) * L.sk + (long long)(j + j0) * L.sj + (i + i0)] =
        div[((long long)k * ny + j) * nx + i];
}
__global__ void __launch_bounds__(256)
k_extract(const double* __restrict__ xmg, double* __restrict__ p, Level L, int nz, int ny, int nx,
          int k0, int j0, int i0, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= nx || j >= ny || k >= nz) return;
    p[((long long)k * ny + j) * nx + i] =
        xmg[(long long)(k + k0) * L.sk + (long long)(j + j0) * L.sj + (i + i0)] * scale;
}

// fused forms of the projection glue (core/projection.py:39-87 + core/mgfordriver.py:66-78):
// U = u*ids2 on the fly, div = delta U (fortran_bernoulli.f90:61-97), stored in the model's div array
// and embedded in the right-hand side of the finest level
__global__ void __launch_bounds__(256)
k_div_embed(const double* __restrict__ ux, const double* __restrict__ uy, const double* __restrict__ uz,
            double* __restrict__ div, double* __restrict__ bmg, double idx2, double idy2, double idz2,
            Level L, int nz, int ny, int nx, int k0, int j0, int i0)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= nx || j >= ny || k >= nz) return;
    const long long sj = nx, sk = (long long)nx * ny, c = (long long)k * sk + (long long)j * sj + i;
    const double U0 = ux[c] * idx2, V0 = uy[c] * idy2, W0 = uz[c] * idz2;
    double d = (i > 0) ? (U0 - ux[c - 1] * idx2) : U0;
    d = (j > 0) ? d + (V0 - uy[c - sj] * idy2) : d + V0;
    d = (k > 0) ? d + (W0 - uz[c - sk] * idz2) : d + W0;
    div[c] = d;
    bmg[(long long)(k + k0) * L.sk + (long long)(j + j0) * L.sj + (i + i0)] = d;
}
// p = x*scale over the model array, u -= delta p (fortran_bernoulli.f90:2-26 as called by projection.py:84-87)
__global__ void __launch_bounds__(256)
k_extract_gradp(const double* __restrict__ xmg, double* __restrict__ p, double* __restrict__ ux,
                double* __restrict__ uy, double* __restrict__ uz, Level L, int nz, int ny, int nx,
                int k0, int j0, int i0, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i >= nx || j >= ny || k >= nz) return;
    const long long c = ((long long)k * ny + j) * nx + i;
    const long long m = (long long)(k + k0) * L.sk + (long long)(j + j0) * L.sj + (i + i0);
    const double p0 = xmg[m] * scale;
    p[c] = p0;
    if (i < nx - 1) ux[c] = ux[c] - (xmg[m + 1] * scale - p0);
    if (j < ny - 1) uy[c] = uy[c] - (xmg[m + L.sj] * scale - p0);
    if (k < nz - 1) uz[c] = uz[c] - (xmg[m + L.sk] * scale - p0);
}

// =================================================================================================
//  box kernels (analytic coefficients): the hot path
// =================================================================================================

// ---- fused smoother: both Jacobi sweeps of fsmoother3d in one pass over x and b ----------------
// A CTA owns a 12 x 60 output tile of the padded plane (16 x 64 with the 2-cell apron the second
// sweep needs) and a chunk of planes; it streams along k.  A thread owns a 2 x 2 patch of columns
// (two rows, one double2 each: row loads are 512-byte coalesced vector loads) and keeps, per
// column, x[p-1..p+2], y[p-2..p] and b[p-1..p+1] in registers.  Of the in-plane neighbours, half
// the rows come from the thread's own registers, the columns from the neighbouring lanes (warp
// shuffles), and only the row above / below the patch goes through shared memory (double
// buffered, one __syncthreads per plane).
// Iteration p: stage 1 forms y[p] (sweep 1) from x[p-1], x[p], x[p+1]; stage 2 forms x'[p-1]
// (sweep 2) from y[p-2], y[p-1], y[p] and stores it.  Halo cells of the output buffer receive the
// old x, so the caller can swap x and y afterwards.
constexpr int S2_NW = 8;                 // warps per CTA
constexpr int S2_RJ = 2 * S2_NW;         // region rows
constexpr int S2_RI = 64;                // region columns
constexpr int S2_TJ = S2_RJ - 4, S2_TI = S2_RI - 4;

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_dn1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// one Jacobi sweep on the thread's 2 x 2 patch: centre rows cA, cB (this plane), the planes below /
// above (mA, mB / pA, pB), the rows above / below the patch (up, dn), right-hand side and 1/diag
__device__ __forceinline__ void sweep_patch(double2 cA, double2 cB, double2 mA, double2 mB, double2 pA, double2 pB,
                                            double2 up, double2 dn, double2 bA, double2 bB,
                                            double iA0, double iA1, double iB0, double iB1,
                                            double omega, double cff1, double2& oA, double2& oB)
{
    const double lA = shfl_up1(cA.y), rA = shfl_dn1(cA.x), lB = shfl_up1(cB.y), rB = shfl_dn1(cB.x);
    // sum order of the Fortran: x(i-1) + x(i+1) + x(j-1) + x(j+1) + x(k-1) + x(k+1)
    const double sA0 = lA + cA.y + up.x + cB.x + mA.x + pA.x;
    const double sA1 = cA.x + rA + up.y + cB.y + mA.y + pA.y;
    const double sB0 = lB + cB.y + cA.x + dn.x + mB.x + pB.x;
    const double sB1 = cB.x + rB + cA.y + dn.y + mB.y + pB.y;
    oA.x = cff1 * cA.x + omega * (sA0 - bA.x) * iA0;
    oA.y = cff1 * cA.y + omega * (sA1 - bA.y) * iA1;
    oB.x = cff1 * cB.x + omega * (sB0 - bB.x) * iB0;
    oB.y = cff1 * cB.y + omega * (sB1 - bB.y) * iB1;
}

__global__ void __launch_bounds__(S2_NW * 32, 2)
k_smooth2(const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ xo,
          Box g, double omega, double cff1, int kchunk)
{
    __shared__ double2 sx[2][S2_RJ][32];
    __shared__ double2 sy[2][S2_RJ][32];
    __shared__ double s_recip[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) s_recip[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;

    const int ai = (int)blockIdx.x * S2_TI - 2 + 2 * lane;          // array column of element .x (even)
    const int ljA = 2 * warp, ljB = ljA + 1;                         // rows of the patch inside the region
    const int ajA = (int)blockIdx.y * S2_TJ - 2 + ljA, ajB = ajA + 1;
    const int ko0 = (int)blockIdx.z * kchunk;
    const int ko1 = min(ko0 + kchunk, g.nz);
    const int tx = g.nx + 2 * NH, ty = g.ny + 2 * NH;

    const bool colok = ai >= 0 && ai < tx;
    const bool inA = colok && ajA >= 0 && ajA < ty, inB = colok && ajB >= 0 && ajB < ty;
    const long long offA = inA ? (long long)ajA * g.sj + ai : -1, offB = inB ? (long long)ajB * g.sj + ai : -1;
    // in-plane neighbour counts of the four cells (negative outside the domain) and 1/diag for cz = 2
    const int cA0 = cnt_xy(g, ai, ajA), cA1 = cnt_xy(g, ai + 1, ajA), cB0 = cnt_xy(g, ai, ajB), cB1 = cnt_xy(g, ai + 1, ajB);
    const double rA0 = cA0 >= 0 ? 1.0 / (double)(cA0 + 2) : 0.0, rA1 = cA1 >= 0 ? 1.0 / (double)(cA1 + 2) : 0.0;
    const double rB0 = cB0 >= 0 ? 1.0 / (double)(cB0 + 2) : 0.0, rB1 = cB1 >= 0 ? 1.0 / (double)(cB1 + 2) : 0.0;
    const bool i0in = ai >= NH && ai < g.nx + NH, i1in = ai + 1 >= NH && ai + 1 < g.nx + NH;
    const bool jAin = ajA >= NH && ajA < g.ny + NH, jBin = ajB >= NH && ajB < g.ny + NH;
    const bool lane_out = lane >= 1 && lane <= 30;
    const bool outA = inA && lane_out && ljA >= 2 && ljA < S2_RJ - 2, outB = inB && lane_out && ljB >= 2 && ljB < S2_RJ - 2;
    const int jup = ljA > 0 ? ljA - 1 : 0, jdn = ljB < S2_RJ - 1 ? ljB + 1 : S2_RJ - 1;

    auto ld = [&](const double* __restrict__ a, long long off, int p) -> double2 {
        if (off < 0 || p < 0 || p >= g.nz) return make_double2(0.0, 0.0);
        return *reinterpret_cast<const double2*>(a + (long long)p * g.sk + off);
    };
    const int p0 = ko0 - 1;
    double2 xmA = ld(x, offA, p0 - 1), xcA = ld(x, offA, p0), xpA = ld(x, offA, p0 + 1);
    double2 xmB = ld(x, offB, p0 - 1), xcB = ld(x, offB, p0), xpB = ld(x, offB, p0 + 1);
    double2 bmA = ld(b, offA, p0 - 1), bcA = ld(b, offA, p0), bmB = ld(b, offB, p0 - 1), bcB = ld(b, offB, p0);
    double2 ymA = make_double2(0.0, 0.0), ycA = ymA, ymB = ymA, ycB = ymA;
    __syncthreads();

    for (int p = p0; p <= ko1; p++) {
        const int buf = p & 1, q = p - 1;
        // prefetch one plane ahead
        const double2 xnA = ld(x, offA, p + 2), xnB = ld(x, offB, p + 2), bnA = ld(b, offA, p + 1), bnB = ld(b, offB, p + 1);
        sx[buf][ljA][lane] = xcA;
        sx[buf][ljB][lane] = xcB;
        __syncthreads();                                  // x[p] (and y[p-1] from the last iteration) visible
        const int cz1 = cnt_z(g, p), cz2 = cnt_z(g, q);
        double2 ypA, ypB;
        {   // ---- stage 1: sweep 1 at plane p
            double iA0 = rA0, iA1 = rA1, iB0 = rB0, iB1 = rB1;
            if (cz1 != 2) {                               // planes next to the z ends (CTA-uniform)
                iA0 = s_recip[max(cA0 + cz1, 0)]; iA1 = s_recip[max(cA1 + cz1, 0)];
                iB0 = s_recip[max(cB0 + cz1, 0)]; iB1 = s_recip[max(cB1 + cz1, 0)];
            }
            sweep_patch(xcA, xcB, xmA, xmB, xpA, xpB, sx[buf][jup][lane], sx[buf][jdn][lane], bcA, bcB,
                        iA0, iA1, iB0, iB1, omega, cff1, ypA, ypB);
        }
        {   // ---- stage 2: sweep 2 at plane q = p-1
            double iA0 = rA0, iA1 = rA1, iB0 = rB0, iB1 = rB1;
            if (cz2 != 2) {
                iA0 = s_recip[max(cA0 + cz2, 0)]; iA1 = s_recip[max(cA1 + cz2, 0)];
                iB0 = s_recip[max(cB0 + cz2, 0)]; iB1 = s_recip[max(cB1 + cz2, 0)];
            }
            double2 oA, oB;
            sweep_patch(ycA, ycB, ymA, ymB, ypA, ypB, sy[buf ^ 1][jup][lane], sy[buf ^ 1][jdn][lane], bmA, bmB,
                        iA0, iA1, iB0, iB1, omega, cff1, oA, oB);
            const bool qint = q >= NH && q < g.nz - NH;
            if (!(qint && jAin && i0in)) oA.x = xmA.x;    // outside the interior: keep x
            if (!(qint && jAin && i1in)) oA.y = xmA.y;
            if (!(qint && jBin && i0in)) oB.x = xmB.x;
            if (!(qint && jBin && i1in)) oB.y = xmB.y;
            if (q >= ko0 && q < ko1) {
                if (outA) *reinterpret_cast<double2*>(xo + (long long)q * g.sk + offA) = oA;
                if (outB) *reinterpret_cast<double2*>(xo + (long long)q * g.sk + offB) = oB;
            }
        }
        sy[buf][ljA][lane] = ypA;
        sy[buf][ljB][lane] = ypB;
        xmA = xcA; xcA = xpA; xpA = xnA; xmB = xcB; xcB = xpB; xpB = xnB;
        bmA = bcA; bcA = bnA; bmB = bcB; bcB = bnB;
        ymA = ycA; ycA = ypA; ymB = ycB; ycB = ypB;
    }
}

// ---- residual on the interior with analytic diag; what happens to r depends on MODE ------------
enum { RS_STORE, RS_NORM, RS_NORMB, RS_NORM2 };   // RS_NORM2: sum r^2 and sum b^2 in one pass over b
// block (32,8): tile 32 x 8 of the interior, marching over a chunk of planes
template <int MODE>
__global__ void __launch_bounds__(256)
k_resid(const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ r,
        Box g, int kchunk, double* __restrict__ partial)
{
    __shared__ double sh[8];
    const int ai = NH + blockIdx.x * 32 + threadIdx.x;
    const int aj = NH + blockIdx.y * 8 + threadIdx.y;
    const int k0 = NH + blockIdx.z * kchunk, k1 = min(k0 + kchunk, g.nz - NH);
    const bool act = ai < g.nx + NH && aj < g.ny + NH;
    double acc = 0.0, accb = 0.0;
    if (act) {
        const int cxy = cnt_xy(g, ai, aj);
        const long long o = (long long)aj * g.sj + ai;
        if (MODE == RS_NORMB) {
            for (int k = k0; k < k1; k++) { double v = b[(long long)k * g.sk + o]; acc = acc + v * v; }
        } else {
            double xm = x[(long long)(k0 - 1) * g.sk + o], xc = x[(long long)k0 * g.sk + o];
            for (int k = k0; k < k1; k++) {
                const long long c = (long long)k * g.sk + o;
                const double xp = x[c + g.sk];
                const double s = x[c - 1] + x[c + 1] + x[c - g.sj] + x[c + g.sj] + xm + xp;
                const double diag = (double)(cxy + cnt_z(g, k));
                const double bv = b[c];
                const double rv = bv + diag * xc - s;
                if (MODE == RS_STORE) r[c] = rv; else acc = acc + rv * rv;
                if (MODE == RS_NORM2) accb = accb + bv * bv;
                xm = xc; xc = xp;
            }
        }
    }
    if (MODE != RS_STORE) {
        const int tid = threadIdx.y * 32 + threadIdx.x;
        for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) sh[tid >> 5] = acc;
        __syncthreads();
        const long long slot = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; w++) t = t + sh[w];
            partial[slot] = t;
        }
        if (MODE == RS_NORM2) {          // the same tree for sum b^2; its partials follow those of sum r^2
            __syncthreads();
            for (int o = 16; o > 0; o >>= 1) accb = accb + __shfl_xor_sync(0xffffffffu, accb, o);
            if ((tid & 31) == 0) sh[tid >> 5] = accb;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < 8; w++) t = t + sh[w];
                partial[(long long)gridDim.x * gridDim.y * gridDim.z + slot] = t;
            }
        }
    }
}

// ---- residual + restriction: b_c = 0.5 * sum_8 r_f, r never stored --------------------------------
// block (32,8): a thread owns one fine column i and the two fine rows of a coarse row, and marches
// over fine plane pairs.  All loads are row-coalesced; z neighbours and the row partner come from
// registers; the i-pair partner's residuals arrive by shuffle, and the even lane adds the eight
// residuals in the order of frestrict_centers3d.
__global__ void __launch_bounds__(256)
k_resid_restrict(const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ bc,
                 Box g, Box gc, int kchunk)
{
    const int fi = (int)blockIdx.x * 32 + threadIdx.x;              // 0-based interior fine column
    const int jc = (int)blockIdx.y * 8 + threadIdx.y;               // 0-based interior coarse row
    const int kc0 = (int)blockIdx.z * kchunk, kc1 = min(kc0 + kchunk, gc.nz - 2 * NH);
    const bool act = fi < g.nx && jc < gc.ny;
    const int ai = NH + (act ? fi : 0), ajA = NH + 2 * (act ? jc : 0), ajB = ajA + 1;
    const int cA = cnt_xy(g, ai, ajA), cB = cnt_xy(g, ai, ajB);
    const long long oA = (long long)ajA * g.sj + ai, oB = oA + g.sj;
    int k = NH + 2 * kc0;
    double xmA = x[(long long)(k - 1) * g.sk + oA], xcA = x[(long long)k * g.sk + oA];
    double xmB = x[(long long)(k - 1) * g.sk + oB], xcB = x[(long long)k * g.sk + oB];
    for (int kc = kc0; kc < kc1; kc++) {
        double s = 0.0;
#pragma unroll
        for (int dk = 0; dk < 2; dk++, k++) {
            const long long pA = (long long)k * g.sk + oA, pB = pA + g.sj;
            const double xpA = x[pA + g.sk], xpB = x[pB + g.sk];
            const int czk = cnt_z(g, k);
            // x(i-1) + x(i+1) + x(j-1) + x(j+1) + x(k-1) + x(k+1)
            const double sA = x[pA - 1] + x[pA + 1] + x[pA - g.sj] + xcB + xmA + xpA;
            const double sB = x[pB - 1] + x[pB + 1] + xcA + x[pB + g.sj] + xmB + xpB;
            const double rA = b[pA] + (double)(cA + czk) * xcA - sA;
            const double rB = b[pB] + (double)(cB + czk) * xcB - sB;
            const double rA1 = __shfl_down_sync(0xffffffffu, rA, 1), rB1 = __shfl_down_sync(0xffffffffu, rB, 1);
            s = dk == 0 ? rA : s + rA;
            s = s + rA1;
            s = s + rB;
            s = s + rB1;
            xmA = xcA; xcA = xpA; xmB = xcB; xcB = xpB;
        }
        if (act && !(fi & 1))
            bc[(long long)(NH + kc) * gc.sk + (long long)(NH + jc) * gc.sj + (NH + (fi >> 1))] = 0.5 * s;
    }
}

// ---- prolongation with the analytic Pcoef: x_f += Pcoef * (trilinear weights) x_c -------------------
// block (32,8): a thread owns one fine column and marches over coarse planes; the plane sums
// 9 xc + 3 xc(i') + 3 xc(j') + xc(i',j') of three consecutive coarse planes sit in registers, so a
// fine cell costs two coarse loads (L1 hits: four fine columns share them), one load and one store.
__global__ void __launch_bounds__(256)
k_prolong_box(double* __restrict__ xf, const double* __restrict__ xc, Box g, Box gc, int kchunk)
{
    const int fi = (int)blockIdx.x * 32 + threadIdx.x, fj = (int)blockIdx.y * 8 + threadIdx.y;   // 0-based interior fine
    if (fi >= g.nx || fj >= g.ny) return;
    const int kc0 = (int)blockIdx.z * kchunk, kc1 = min(kc0 + kchunk, gc.nz - 2 * NH);          // 0-based interior coarse planes
    // Fortran: ic = (i+1)/2, di = -1 for odd i (1-based)  <=>  0-based fine fi even -> neighbour ic-1
    const int aic = NH + (fi >> 1), ajc = NH + (fj >> 1);
    const int di = (fi & 1) ? 1 : -1, dj = (fj & 1) ? 1 : -1;
    const long long oc = (long long)ajc * gc.sj + aic, ox = di, oy = (long long)dj * gc.sj;
    const int exy = (int)in_x(gc, aic + di) + (int)in_y(gc, ajc + dj);
    const long long of = (long long)(NH + fj) * g.sj + (NH + fi);
    auto plane = [&](int akc) -> double {
        const long long c = (long long)akc * gc.sk + oc;
        return 9 * xc[c] + 3 * xc[c + ox] + 3 * xc[c + oy] + xc[c + ox + oy];
    };
    double pa = plane(NH + kc0 - 1), pb = plane(NH + kc0);
    for (int kc = kc0; kc < kc1; kc++) {
        const int akc = NH + kc;
        const double pc = plane(akc + 1);
        const long long f0 = (long long)(NH + 2 * kc) * g.sk + of, f1 = f0 + g.sk;
        const double c0 = pcoef_of(exy + (int)in_z(gc, akc - 1)), c1 = pcoef_of(exy + (int)in_z(gc, akc + 1));
        xf[f0] = xf[f0] + c0 * (3 * pb + pa);
        xf[f1] = xf[f1] + c1 * (3 * pb + pc);
        pa = pb; pb = pc;
    }
}

#include "ny_mg_vleg.cuh"

// ---- the tail of the V-cycle in ONE launch -------------------------------------------------------
// Levels of a few thousand cells cost ~20 us per leg as separate launches (pipeline fill, launch and
// drain dominate) and there are a dozen of them per V-cycle.  One launch runs the whole tail -- smooth,
// residual + restriction down to the coarsest level, its smoothing, prolongation + smooth back up -- with a
// barrier between the operators.  Two kinds of level:
//   narrow (at most tail_cells cells): one CTA, __syncthreads between the operators;
//   wide   (at most wide_cells cells): every CTA of a co-resident grid (cooperative launch, one CTA per SM),
//          a grid barrier between the operators.  The plane-marching legs have too few tiles on such levels
//          (a 128 x 32 plane is four tiles), so that a leg costs 25-45 us whatever the level holds; one cell per
//          thread and ~3 us per barrier is several times faster there.
// The wide levels are the leading ones; CTA 0 runs the narrow levels while the others wait at the barrier
// that ends them.  The arrays stay in global memory (the barrier fences; no read-only-path loads).
// Per-cell arithmetic and operation order are those of k_sweep / k_residual / k_restrict / k_prolong with the
// analytic box coefficients, hence of the reference (basicoperators.f90:32-60,173-231,300-323,363-400).
constexpr int TAIL_MAX = 10;
constexpr int TAIL_THREADS = 1024;
// n / d for 0 <= n < 2^31 and a divisor known on the host, as a multiplication (Granlund & Montgomery 1994, N = 32:
// l = ceil(log2 d), m = floor(2^32 (2^l - d) / d) + 1, n / d = (mulhi(m, n) + n) >> l): the tail deals cells to threads
// by their linear index, and a hardware-less 32-bit division per index costs more than the stencil itself
struct FastDiv { unsigned m, l; };
inline FastDiv fastdiv_of(int d)
{
    FastDiv f;
    f.l = 0;
    while ((1ull << f.l) < (unsigned long long)d) f.l++;
    f.m = (unsigned)(((1ull << 32) * ((1ull << f.l) - (unsigned long long)d)) / (unsigned long long)d + 1ull);
    return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f)
{
    return (int)((__umulhi(f.m, (unsigned)n) + (unsigned)n) >> f.l);
}
// fx / fy: divisions by nx / ny + {0, 2, 2 NH}: the interior, the interior widened by one ring, the whole plane
struct TailLevel { double *x, *y, *b; Box g; FastDiv fx[3], fy[3]; };
struct TailArgs {
    int n, nwide;
    unsigned* bar;
    double* norm_partial;                    // not null: sum r^2 of the first level after the cycle: one partial per CTA,
    double *norm_out, *norm_host;            //   their sum to the device slot and to the host's pinned mailbox
    double omega, cff1;
    TailLevel lev[TAIL_MAX];
};

// Sense-reversing barrier over the CTAs of a cooperative launch: bar[0] counts arrivals, bar[1] is the generation.
// Thread 0 arrives with an acq_rel atomic (release: what its CTA wrote before the block barrier; acquire: what the
// CTAs that arrived earlier wrote), the last one opens the next generation with a release store, the others poll it
// with acquire loads; the block barriers extend the ordering to the rest of the CTA.
__device__ __forceinline__ void tail_grid_barrier(unsigned* bar)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned gen, now, prev;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(bar) : "memory");
        if (prev == gridDim.x - 1) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(bar), "r"(0u) : "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bar + 1), "r"(gen + 1) : "memory");
        } else {
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(bar + 1) : "memory");
            } while (now == gen);
        }
    }
    __syncthreads();
}

// the threads that share an operator: index t of T, and their barrier
struct TailTeam {
    int t, T;
    unsigned* bar;                           // nullptr: the threads of one CTA
    bool on;                                 // this thread belongs to the team
    __device__ __forceinline__ void sync() const { if (bar) tail_grid_barrier(bar); else __syncthreads(); }
};

// cnt_xy / cnt_z without branches (the tail's phases are issue bound: a thread has one or two cells per operator)
__device__ __forceinline__ int tin_x(const Box& g, int ai) { return g.xper | (int)((unsigned)(ai - NH) < (unsigned)g.nx); }
__device__ __forceinline__ int tin_y(const Box& g, int aj) { return g.yper | (int)((unsigned)(aj - NH) < (unsigned)g.ny); }
__device__ __forceinline__ int tin_z(const Box& g, int ak) { return ((int)(ak >= NH) | g.zlo) & ((int)(ak < g.nz - NH) | g.zhi); }
__device__ __forceinline__ int tail_cnt(const Box& g, int ai, int aj, int ak)
{
    const int inside = tin_x(g, ai) & tin_y(g, aj) & tin_z(g, ak);
    const int n = tin_x(g, ai - 1) + tin_x(g, ai + 1) + tin_y(g, aj - 1) + tin_y(g, aj + 1) + tin_z(g, ak - 1) + tin_z(g, ak + 1);
    return inside ? n : -100;              // cnt_xy + cnt_z wherever that sum is positive; negative elsewhere, as it is
}

// (i, j, k) of cell c of an ni x nj x nk box in row-major order; the threads of a team take cells t, t + T, ...
struct TailCell {
    int i, j, k;
    __device__ __forceinline__ TailCell(int c, int ni, int nj, const FastDiv& fi, const FastDiv& fj)
    {
        const int q = fdiv(c, fi);
        i = c - q * ni;
        k = fdiv(q, fj);
        j = q - k * nj;
    }
};

// one Jacobi sweep src -> dst on the interior widened by `ring` cells (fsmoother3d: ring 1, then ring 0)
__device__ __forceinline__ void tail_sweep(const double* src, double* dst, const TailLevel& L, double omega, double cff1,
                                           int ring, const TailTeam& tm, const double* recip)
{
    const Box& g = L.g;
    const double* b = L.b;
    const int ni = g.nx + 2 * ring, nj = g.ny + 2 * ring, total = ni * nj * (g.nz - 2 * NH + 2 * ring);
    const FastDiv fi = L.fx[ring], fj = L.fy[ring];
#pragma unroll 2
    for (int t = tm.t; t < total; t += tm.T) {
        const TailCell it(t, ni, nj, fi, fj);
        const int ai = NH - ring + it.i, aj = NH - ring + it.j, ak = NH - ring + it.k;
        const long long c = (long long)ak * g.sk + (long long)aj * g.sj + ai;
        const double s = src[c - 1] + src[c + 1] + src[c - g.sj] + src[c + g.sj] + src[c - g.sk] + src[c + g.sk];
        const int cnt = tail_cnt(g, ai, aj, ak);
        const double idiag = cnt > 0 ? recip[cnt] : 0.0;               // 1 / cnt, correctly rounded
        dst[c] = cff1 * src[c] + omega * (s - b[c]) * idiag;
    }
}

__device__ __forceinline__ double tail_resid(const double* x, const double* b, const Box& g, int ai, int aj, int ak)
{
    const long long c = (long long)ak * g.sk + (long long)aj * g.sj + ai;
    const double s = x[c - 1] + x[c + 1] + x[c - g.sj] + x[c + g.sj] + x[c - g.sk] + x[c + g.sk];
    const double diag = (double)tail_cnt(g, ai, aj, ak);            // interior cells only: cnt_xy + cnt_z
    return b[c] + diag * x[c] - s;
}

constexpr int TAIL_STAGE = 2048;          // doubles of staging for array sections that overlap themselves

// array section dst = src with Fortran semantics (right-hand side evaluated first), Fortran indices as in
// k_box_copy; every thread of ONE CTA takes part (narrow levels only), the section is complete when the call returns
__device__ __forceinline__ void tail_assign_box(double* a, const Box& g, int di0, int dj0, int dk0, int si0, int sj0, int sk0,
                                                int ni, int nj, int nk, bool staged, double* stage)
{
    const int total = ni * nj * nk;
    auto at = [&](int i, int j, int k) { return (long long)(k - 1) * g.sk + (long long)(j - 1 + NH) * g.sj + (i - 1 + NH); };
    if (total <= 0) return;
    if (staged) {
        for (int t = threadIdx.x; t < total; t += TAIL_THREADS) {
            const int i = t % ni, q = t / ni, j = q % nj, k = q / nj;
            stage[t] = a[at(si0 + i, sj0 + j, sk0 + k)];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < total; t += TAIL_THREADS) {
            const int i = t % ni, q = t / ni, j = q % nj, k = q / nj;
            a[at(di0 + i, dj0 + j, dk0 + k)] = stage[t];
        }
    } else {
        for (int t = threadIdx.x; t < total; t += TAIL_THREADS) {
            const int i = t % ni, q = t / ni, j = q % nj, k = q / nj;
            a[at(di0 + i, dj0 + j, dk0 + k)] = a[at(si0 + i, sj0 + j, sk0 + k)];
        }
    }
    __syncthreads();
}

// the statement sequence of mod_halo.f90:235-262 (fill_sequential) for the tiny periodic levels whose halo sections
// overlap themselves (an axis narrower than the halo); one CTA, kept out of line: it is rare and large
__device__ __noinline__ void tail_fill_sequential(double* a, int nx, int ny, int nz, long long sj, long long sk,
                                                  int xper, int yper, int wz, double* stage)
{
    Box g;                                   // (by value: a reference into the kernel's parameters would drag them all
    g.nx = nx; g.ny = ny; g.nz = nz; g.sj = sj; g.sk = sk;                   //  into local memory)
    g.xper = xper; g.yper = yper; g.zlo = wz; g.zhi = wz;
    const int nzi = nz - 2 * NH;
    const int nh = NH;
    if (g.xper) {
        const bool ov = nx < nh;
        tail_assign_box(a, g, nx + 1, 1, 1, 1, 1, 1, nh, ny, nz, ov, stage);
        tail_assign_box(a, g, 1 - nh, 1, 1, nx - nh + 1, 1, 1, nh, ny, nz, ov, stage);
    }
    if (g.yper) {
        const bool ov = ny < nh;
        tail_assign_box(a, g, 1, ny + 1, 1, 1, 1, 1, nx, nh, nz, ov, stage);
        tail_assign_box(a, g, 1, 1 - nh, 1, 1, ny - nh + 1, 1, nx, nh, nz, ov, stage);
    }
    if (g.xper && g.yper) {
        const bool ov = nx < nh || ny < nh;
        tail_assign_box(a, g, nx + 1, ny + 1, 1, 1, 1, 1, nh, nh, nz, ov, stage);
        tail_assign_box(a, g, 1 - nh, ny + 1, 1, nx - nh + 1, 1, 1, nh, nh, nz, ov, stage);
        tail_assign_box(a, g, nx + 1, 1 - nh, 1, 1, ny - nh + 1, 1, nh, nh, nz, ov, stage);
        tail_assign_box(a, g, 1 - nh, 1 - nh, 1, nx - nh + 1, ny - nh + 1, 1, nh, nh, nz, ov, stage);
    }
    if (wz) {
        const bool ov = nzi < nh;
        tail_assign_box(a, g, 1 - nh, 1 - nh, 1, 1 - nh, 1 - nh, nzi + 1, nx + 2 * nh, ny + 2 * nh, nh, ov, stage);
        tail_assign_box(a, g, 1 - nh, 1 - nh, nz - nh + 1, 1 - nh, 1 - nh, nh + 1, nx + 2 * nh, ny + 2 * nh, nh, ov, stage);
    }
}


// periodic halo fill of a replicated level inside the tail: the rule of k_fill_periodic where every wrapped axis is
// at least nh wide, the statement sequence of mod_halo.f90:235-262 (fill_sequential) on the tiny levels below that
// (those are narrow levels: tail_first sees to it)
__device__ __forceinline__ void tail_fill(double* a, const TailLevel& L, double* stage, const TailTeam& tm)
{
    const Box& g = L.g;
    const FastDiv fx = L.fx[2], fy = L.fy[2];
    const bool wz = g.zlo && g.zhi;
    if (!(g.xper || g.yper || wz)) return;
    const int nx = g.nx, ny = g.ny, nz = g.nz, nzi = nz - 2 * NH;
    const bool simple = (!g.xper || nx >= NH) && (!g.yper || ny >= NH) && (!wz || nzi >= NH);
    if (simple) {
        const int tx = nx + 2 * NH, ty = ny + 2 * NH, n = tx * ty * nz;
        for (int t = tm.t; t < n; t += tm.T) {
            const TailCell it(t, tx, ty, fx, fy);
            const int ai = it.i, aj = it.j, ak = it.k;
            const bool hx = ai < NH || ai >= nx + NH, hy = aj < NH || aj >= ny + NH, hz = ak < NH || ak >= nz - NH;
            int si = ai, sjj = aj, skk = ak;
            if (hz && wz) skk = ak < NH ? ak + nzi : ak - nzi;
            if (hx && hy) {
                if (g.xper && g.yper) { si = ai < NH ? ai + nx : ai - nx; sjj = aj < NH ? aj + ny : aj - ny; }
            } else if (hx) {
                if (g.xper) si = ai < NH ? ai + nx : ai - nx;
            } else if (hy) {
                if (g.yper) sjj = aj < NH ? aj + ny : aj - ny;
            }
            if (si != ai || sjj != aj || skk != ak)
                a[(long long)ak * g.sk + (long long)aj * g.sj + ai] = a[(long long)skk * g.sk + (long long)sjj * g.sj + si];
        }
        tm.sync();
        return;
    }
    tail_fill_sequential(a, nx, ny, nz, g.sj, g.sk, g.xper, g.yper, (int)wz, stage);
}

__global__ void __launch_bounds__(TAIL_THREADS, 1)
k_vcycle_tail(TailArgs a)
{
    __shared__ double stage[TAIL_STAGE];
    __shared__ double recip[8];
    if (threadIdx.x < 8) recip[threadIdx.x] = threadIdx.x ? 1.0 / (double)threadIdx.x : 0.0;
    __syncthreads();
    const double omega = a.omega, cff1 = a.cff1;
    const int nwide = a.nwide;
    // levels [0, nwide): all CTAs; levels [nwide, n): CTA 0
    const TailTeam wide = {(int)(blockIdx.x * TAIL_THREADS + threadIdx.x), (int)(gridDim.x * TAIL_THREADS), a.bar, true};
    const TailTeam narrow = {(int)threadIdx.x, TAIL_THREADS, nullptr, blockIdx.x == 0};
    auto team = [&](int l) -> const TailTeam& { return l < nwide ? wide : narrow; };
    auto smooth = [&](const TailLevel& L, const TailTeam& tm) {
        tail_sweep(L.x, L.y, L, omega, cff1, 1, tm, recip);
        tm.sync();
        tail_sweep(L.y, L.x, L, omega, cff1, 0, tm, recip);
        tm.sync();
        tail_fill(L.x, L, stage, tm);                    // operators.f90:168
    };
    for (int l = 0; l + 1 < a.n; l++) {                 // solvers.f90:41-46
        const TailLevel& F = a.lev[l];
        const TailLevel& C = a.lev[l + 1];
        const TailTeam& tf = team(l);
        const TailTeam& tc = team(l + 1);
        const Box& g = F.g;
        const Box& gc = C.g;
        if (tf.on) {
            smooth(F, tf);
            // eight threads per coarse cell, one fine residual each; the first adds them in the order of
            // frestrict_centers3d (i fastest, then j, then k).  T is a multiple of 32: the eight are lanes 8m .. 8m + 7.
            const int ncells = gc.nx * gc.ny * (gc.nz - 2 * NH);
            const int lane8 = threadIdx.x & 24, sub = threadIdx.x & 7;
            for (int t = tf.t; (t & ~31) < 8 * ncells; t += tf.T) {
                const int cell = t >> 3;
                const bool live = cell < ncells;
                const TailCell it(live ? cell : 0, gc.nx, gc.ny, C.fx[0], C.fy[0]);
                const int ic = it.i, jc = it.j, kc = it.k;
                const int ai = NH + 2 * ic + (sub & 1), aj = NH + 2 * jc + ((sub >> 1) & 1), ak = NH + 2 * kc + (sub >> 2);
                const double rv = live ? tail_resid(F.x, F.b, g, ai, aj, ak) : 0.0;
                double r = rv;
#pragma unroll
                for (int o = 1; o < 8; o++) r = r + __shfl_sync(0xffffffffu, rv, lane8 + o);
                if (live && sub == 0)
                    C.b[(long long)(NH + kc) * gc.sk + (long long)(NH + jc) * gc.sj + (NH + ic)] = 0.5 * r;
            }
            const long long nc = gc.sk * gc.nz;
            for (long long t = tf.t; t < nc; t += tf.T) C.x[t] = 0.0;               // operators.f90:209
            tf.sync();
        }
        if (tc.on) tail_fill(C.b, C, stage, tc);         // operators.f90:211
    }
    if (team(a.n - 1).on) smooth(a.lev[a.n - 1], team(a.n - 1));
    for (int l = a.n - 2; l >= 0; l--) {                // solvers.f90:51-54
        const TailLevel& F = a.lev[l];
        const TailLevel& C = a.lev[l + 1];
        const TailTeam& tf = team(l);
        const Box& g = F.g;
        const Box& gc = C.g;
        if (l == nwide - 1 && nwide < a.n) wide.sync();   // what CTA 0 did on the narrow levels becomes visible
        if (!tf.on) continue;
        const int nfine = g.nx * g.ny * (g.nz - 2 * NH);
#pragma unroll 2
        for (int t = tf.t; t < nfine; t += tf.T) {
            const TailCell it(t, g.nx, g.ny, F.fx[0], F.fy[0]);
            const int fi = it.i, fj = it.j, fk = it.k;                              // 0-based interior fine cell
            const int aic = NH + (fi >> 1), ajc = NH + (fj >> 1), akc = NH + (fk >> 1);
            const int di = (fi & 1) ? 1 : -1, dj = (fj & 1) ? 1 : -1, ako = (fk & 1) ? akc + 1 : akc - 1;
            const long long ox = di, oy = (long long)dj * gc.sj;
            const long long cb = (long long)akc * gc.sk + (long long)ajc * gc.sj + aic;
            const long long co = (long long)ako * gc.sk + (long long)ajc * gc.sj + aic;
            const double pb = 9 * C.x[cb] + 3 * C.x[cb + ox] + 3 * C.x[cb + oy] + C.x[cb + ox + oy];
            const double po = 9 * C.x[co] + 3 * C.x[co + ox] + 3 * C.x[co + oy] + C.x[co + ox + oy];
            const double cf = pcoef_of(tin_x(gc, aic + di) + tin_y(gc, ajc + dj) + tin_z(gc, ako));
            const long long f = (long long)(NH + fk) * g.sk + (long long)(NH + fj) * g.sj + (NH + fi);
            F.x[f] = F.x[f] + cf * (3 * pb + po);
        }
        tf.sync();
        tail_fill(F.x, F, stage, tf);                    // operators.f90:242
        smooth(F, tf);
    }
    // the fused legs of the finer levels rely on y == x on the wall halos of every level
    const TailTeam& te = nwide > 0 ? wide : narrow;
    if (!te.on) return;
    for (int l = 0; l < a.n; l++) {
        const TailLevel& L = a.lev[l];
        const long long n = L.g.sk * L.g.nz;
        for (long long t = te.t; t < n; t += te.T) L.y[t] = L.x[t];
    }
    if (a.norm_partial) {                                // the tail is the whole cycle: sum r^2 for the stop test
        const TailLevel& L = a.lev[0];                   // (operators.f90:81-125; x is final and fenced: the smoothing
        double acc = 0.0;                                //  that wrote it ended with the team's barrier)
        const int ncells = L.g.nx * L.g.ny * (L.g.nz - 2 * NH);
        for (int t = te.t; t < ncells; t += te.T) {
            const TailCell it(t, L.g.nx, L.g.ny, L.fx[0], L.fy[0]);
            const double rv = tail_resid(L.x, L.b, L.g, NH + it.i, NH + it.j, NH + it.k);
            acc = acc + rv * rv;
        }
        for (int o = 16; o > 0; o >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, o);
        __syncthreads();                                 // `stage` may still be in use by a halo section
        if ((threadIdx.x & 31) == 0) stage[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < TAIL_THREADS / 32; w++) t = t + stage[w];
            a.norm_partial[blockIdx.x] = t;
        }
        te.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) {       // the CTAs' partial sums in CTA order
            double t = 0.0;
            const int nb = nwide > 0 ? (int)gridDim.x : 1;
            for (int w = 0; w < nb; w++) t = t + a.norm_partial[w];
            *a.norm_out = t;
            *a.norm_host = t;
            __threadfence_system();
        }
    }
}

inline int ew_blocks(long long n)
{
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 4736 ? 4736 : b));
}

#define LAUNCH_OK(mg) NY_CHECK_LAUNCH((mg)->ctx)
#define TRY(call) do { int _r = (call); if (_r != NY_OK) return _r; } while (0)

template <int OP>
int ew(ny_mg* mg, cudaStream_t st, double* dst, const double* a, const double* b, double val, size_t n)
{
    k_ew<OP><<<ew_blocks((long long)n), 256, 0, st>>>(dst, a, b, val, (long long)n);
    LAUNCH_OK(mg);
    return NY_OK;
}

// dst section = src section with Fortran assignment semantics (rhs evaluated first)
int assign_box(ny_mg* mg, cudaStream_t st, const Level& L, double* a, int di0, int dj0, int dk0,
               int si0, int sj0, int sk0, int ni, int nj, int nk, bool may_overlap)
{
    long long total = (long long)ni * nj * nk;
    if (total <= 0) return NY_OK;
    int nb = ew_blocks(total);
    if (may_overlap) {
        if ((size_t)total > mg->tmp_doubles) { ny_set_error("halo staging buffer too small"); return NY_ERR_ARG; }
        k_box_copy<<<nb, 256, 0, st>>>(a, mg->tmp, L, mg->nh, 0, 0, 0, si0, sj0, sk0, ni, nj, nk, 0, 1);
        LAUNCH_OK(mg);
        k_box_copy<<<nb, 256, 0, st>>>(mg->tmp, a, L, mg->nh, di0, dj0, dk0, 0, 0, 0, ni, nj, nk, 1, 0);
        LAUNCH_OK(mg);
    } else {
        k_box_copy<<<nb, 256, 0, st>>>(a, a, L, mg->nh, di0, dj0, dk0, si0, sj0, sk0, ni, nj, nk, 0, 0);
        LAUNCH_OK(mg);
    }
    return NY_OK;
}

// mod_halo.f90:235-262 exchange_with_myself, statement by statement (tiny levels)
int fill_sequential(ny_mg* mg, cudaStream_t st, const Level& L, double* a, bool wrap_z)
{
    const int nh = mg->nh, nx = L.nx, ny = L.ny, nz = L.nz;
    if (mg->xper) {
        bool ov = nx < nh;
        TRY(assign_box(mg, st, L, a, nx + 1, 1, 1, 1, 1, 1, nh, ny, nz, ov));
        TRY(assign_box(mg, st, L, a, 1 - nh, 1, 1, nx - nh + 1, 1, 1, nh, ny, nz, ov));
    }
    if (mg->yper) {
        bool ov = ny < nh;
        TRY(assign_box(mg, st, L, a, 1, ny + 1, 1, 1, 1, 1, nx, nh, nz, ov));
        TRY(assign_box(mg, st, L, a, 1, 1 - nh, 1, 1, ny - nh + 1, 1, nx, nh, nz, ov));
    }
    if (mg->xper && mg->yper) {
        bool ov = nx < nh || ny < nh;
        TRY(assign_box(mg, st, L, a, nx + 1, ny + 1, 1, 1, 1, 1, nh, nh, nz, ov));
        TRY(assign_box(mg, st, L, a, 1 - nh, ny + 1, 1, nx - nh + 1, 1, 1, nh, nh, nz, ov));
        TRY(assign_box(mg, st, L, a, nx + 1, 1 - nh, 1, 1, ny - nh + 1, 1, nh, nh, nz, ov));
        TRY(assign_box(mg, st, L, a, 1 - nh, 1 - nh, 1, nx - nh + 1, ny - nh + 1, 1, nh, nh, nz, ov));
    }
    if (wrap_z) {
        bool ov = (nz - 2 * nh) < nh;
        TRY(assign_box(mg, st, L, a, 1 - nh, 1 - nh, 1, 1 - nh, 1 - nh, nz - 2 * nh + 1, nx + 2 * nh, ny + 2 * nh, nh, ov));
        TRY(assign_box(mg, st, L, a, 1 - nh, 1 - nh, nz - nh + 1, 1 - nh, 1 - nh, nh + 1, nx + 2 * nh, ny + 2 * nh, nh, ov));
    }
    return NY_OK;
}

// halo fill of a level array: periodic wraps locally, z faces of distributed levels through NCCL
int fill(ny_mg* mg, cudaStream_t st, const Level& L, double* a)
{
    const int nh = mg->nh;
    const bool dist = !L.gathered;
    const bool wrap_z = mg->zper && !dist;
    if (mg->xper || mg->yper || wrap_z) {
        const int nzi = L.nz - 2 * nh;
        const bool simple = (!mg->xper || L.nx >= nh) && (!mg->yper || L.ny >= nh) && (!wrap_z || nzi >= nh);
        if (simple) {
            const long long n0 = 2LL * nh * L.sk;
            const long long n1 = mg->yper ? (long long)nzi * 2 * nh * L.sj : 0;
            const long long n2 = mg->xper ? (long long)nzi * L.ny * 2 * nh : 0;
            const long long tot = n0 + n1 + n2;
            k_fill_periodic<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, L.nx, L.ny, L.nz, L.sj, L.sk, mg->xper,
                                                                            mg->yper, wrap_z ? 1 : 0, n0, n1, n2);
            LAUNCH_OK(mg);
        } else {
            TRY(fill_sequential(mg, st, L, a, wrap_z));
        }
    }
    if (dist && (mg->below >= 0 || mg->above >= 0)) {
        double* arr[1] = {a};
        ny_prof_scope ps(mg->ctx, NY_PROF_HALO, st);
        TRY(ny_comm_exchange_z(mg->comm, arr, 1, (size_t)L.sk, nh, L.nz - 2 * nh, nh, mg->below, mg->above, st));
    }
    return NY_OK;
}

inline dim3 box_grid(int ni, int nj, int nk, dim3 b)
{
    return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, (nk + b.z - 1) / b.z);
}

inline int chunk_for(ny_mg* mg, int planes, long long tiles, int min_chunk)
{   // split the planes so that the launch has a few CTAs per SM, but never below min_chunk planes
    long long target = (long long)mg->ctx->num_sms * 32;
    long long nchunks = (target + tiles - 1) / tiles;
    if (nchunks < 1) nchunks = 1;
    int chunk = (int)((planes + nchunks - 1) / nchunks);
    if (chunk < min_chunk) chunk = min_chunk;
    if (chunk > planes) chunk = planes > 0 ? planes : 1;
    return chunk;
}

void swap_xy(ny_mg* mg, int lev)
{
    Level& L = mg->lev[lev - 1];
    LevelMaps& M = mg->maps[lev - 1];
    double* t = L.x; L.x = L.y; L.y = t;
    CUtensorMap m = M.x; M.x = M.y; M.y = m;
    m = M.cx; M.cx = M.cy; M.cy = m;
}

int smooth(ny_mg* mg, cudaStream_t st, int lev)
{
    Level& L = mg->lev[lev - 1];
    const int nh = mg->nh;
    const double omega = mg->omega, cff1 = 1.0 - omega;
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_SMOOTH_FINE : NY_PROF_MG_COARSE, st);
        if (mg->box) {
            const int gx = (L.nx + 2 * nh + S2_TI - 1) / S2_TI, gy = (L.ny + 2 * nh + S2_TJ - 1) / S2_TJ;
            const int chunk = chunk_for(mg, L.nz, (long long)gx * gy, 16);
            dim3 grid(gx, gy, (L.nz + chunk - 1) / chunk);
            k_smooth2<<<grid, S2_NW * 32, 0, st>>>(L.x, L.b, L.y, box_of(mg, L), omega, cff1, chunk);
            LAUNCH_OK(mg);
            swap_xy(mg, lev);
        } else {
            mg->ysync = 0;                 // sweep 1 writes ring 1 of y
            dim3 b(32, 4, 2);
            k_sweep<<<box_grid(L.nx + 2, L.ny + 2, L.nz - 2 * nh + 2, b), b, 0, st>>>(
                L.x, L.y, L.b, L.idiag, omega, cff1, L, nh, 0, L.nx + 1, 0, L.ny + 1, nh, L.nz + 1 - nh);
            LAUNCH_OK(mg);
            k_sweep<<<box_grid(L.nx, L.ny, L.nz - 2 * nh, b), b, 0, st>>>(
                L.y, L.x, L.b, L.idiag, omega, cff1, L, nh, 1, L.nx, 1, L.ny, nh + 1, L.nz - nh);
            LAUNCH_OK(mg);
        }
    }
    return fill(mg, st, L, L.x);
}

// launch geometry of the interior-marching kernels (block 32 x 8)
struct March { dim3 grid; int chunk; int nparts; };
inline March march_geom(ny_mg* mg, int nx, int ny, int nzi)
{
    March m;
    const int gx = (nx + 31) / 32, gy = (ny + 7) / 8;
    m.chunk = chunk_for(mg, nzi, (long long)gx * gy, 8);
    int gz = (nzi + m.chunk - 1) / m.chunk;
    while ((long long)gx * gy * gz > MAX_PARTIALS) { m.chunk *= 2; gz = (nzi + m.chunk - 1) / m.chunk; }
    m.grid = dim3(gx, gy, gz);
    m.nparts = gx * gy * gz;
    return m;
}

int residual(ny_mg* mg, cudaStream_t st, int lev)
{
    Level& L = mg->lev[lev - 1];
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_RESIDUAL_FINE : NY_PROF_MG_COARSE, st);
        if (mg->box) {
            March m = march_geom(mg, L.nx, L.ny, L.nz - 2 * mg->nh);
            k_resid<RS_STORE><<<m.grid, dim3(32, 8), 0, st>>>(L.x, L.b, L.r, box_of(mg, L), m.chunk, nullptr);
        } else {
            dim3 b(32, 4, 2);
            k_residual<<<box_grid(L.nx, L.ny, L.nz - 2 * mg->nh, b), b, 0, st>>>(L.x, L.b, L.r, L.msk, L.diag, L, mg->nh);
        }
        LAUNCH_OK(mg);
    }
    return fill(mg, st, L, L.r);
}

// the coarse level as seen from the slab of the finer level: the level itself, or the window of a
// gathered level that lies under this rank's slab
Level coarse_view(ny_mg* mg, int lev)
{
    const Level& F = mg->lev[lev - 1];
    Level C = mg->lev[lev];
    if (!F.gathered && C.gathered) {
        const int nzc = (F.nz - 2 * mg->nh) / 2;                       // coarse planes under this slab
        const size_t off = (size_t)mg->rank * nzc * C.sk;
        C.nz = nzc + 2 * mg->nh;
        C.n = (size_t)C.sk * C.nz;
        C.zlo = F.zlo; C.zhi = F.zhi;
        C.x += off; C.b += off; C.r += off; C.y += off;
    }
    return C;
}

// after the restriction wrote this rank's part of a gathered level: collect the other parts
int gather_after_restriction(ny_mg* mg, cudaStream_t st, int lev)
{
    const Level& F = mg->lev[lev - 1];
    Level& C = mg->lev[lev];
    if (F.gathered || !C.gathered) return NY_OK;
    const size_t cnt = (size_t)((F.nz - 2 * mg->nh) / 2) * C.sk;
    ny_prof_scope ps(mg->ctx, NY_PROF_HALO, st);
    return ny_comm_allgather_inplace(mg->comm, C.b + (size_t)mg->nh * C.sk, cnt, st);
}

int restriction(ny_mg* mg, cudaStream_t st, int lev, bool from_b)
{
    Level& F = mg->lev[lev - 1];
    Level& C = mg->lev[lev];
    Level V = coarse_view(mg, lev);
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_RESTRICT_FINE : NY_PROF_MG_COARSE, st);
        dim3 b(32, 4, 2);
        k_restrict<<<box_grid(V.nx, V.ny, V.nz - 2 * mg->nh, b), b, 0, st>>>(
            from_b ? F.b : F.r, V.b, mg->box ? nullptr : C.Rcoef, F, V, mg->nh);
        LAUNCH_OK(mg);
        NY_CUDA(cudaMemsetAsync(C.x, 0, C.n * sizeof(double), st));           // operators.f90:209
    }
    TRY(gather_after_restriction(mg, st, lev));
    return fill(mg, st, C, C.b);
}

// box path of "residual(lev); restriction(lev)" without storing r
int residual_restriction(ny_mg* mg, cudaStream_t st, int lev)
{
    if (!mg->box) { TRY(residual(mg, st, lev)); return restriction(mg, st, lev, false); }
    Level& F = mg->lev[lev - 1];
    Level& C = mg->lev[lev];
    Level V = coarse_view(mg, lev);
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_RESIDUAL_FINE : NY_PROF_MG_COARSE, st);
        March m = march_geom(mg, F.nx, V.ny, V.nz - 2 * mg->nh);       // fine columns x coarse rows x coarse planes
        k_resid_restrict<<<m.grid, dim3(32, 8), 0, st>>>(F.x, F.b, V.b, box_of(mg, F), box_of(mg, V), m.chunk);
        LAUNCH_OK(mg);
        NY_CUDA(cudaMemsetAsync(C.x, 0, C.n * sizeof(double), st));
    }
    TRY(gather_after_restriction(mg, st, lev));
    return fill(mg, st, C, C.b);
}

int prolongation(ny_mg* mg, cudaStream_t st, int lev)
{
    Level& F = mg->lev[lev - 1];
    Level V = coarse_view(mg, lev);
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_PROLONG_FINE : NY_PROF_MG_COARSE, st);
        if (mg->box) {
            March m = march_geom(mg, F.nx, F.ny, V.nz - 2 * mg->nh);   // fine columns x fine rows x coarse planes
            k_prolong_box<<<m.grid, dim3(32, 8), 0, st>>>(F.x, V.x, box_of(mg, F), box_of(mg, V), m.chunk);
        } else {
            dim3 b(32, 4, 2);
            k_prolong<<<box_grid(F.nx, F.ny, F.nz - 2 * mg->nh, b), b, 0, st>>>(F.x, V.x, F.Pcoef, F, V, box_of(mg, V), mg->nh);
        }
        LAUNCH_OK(mg);
    }
    return fill(mg, st, F, F.x);
}

// enqueue the global sum of the per-block partials into d_red[MAX_PARTIALS + slot]
int finish_sum(ny_mg* mg, cudaStream_t st, int nparts, int slot, int first = 0)
{
    k_sum_final<<<1, 256, 0, st>>>(mg->d_red + first, nparts, mg->d_red + MAX_PARTIALS + slot);
    LAUNCH_OK(mg);
    return ny_comm_allreduce(mg->comm, mg->d_red + MAX_PARTIALS + slot, 1, 0, st);
}

// sum(msk*b^2) -> slot 0 and, after residual(1), sum(msk*r^2) -> slot 1, from one pass over x and b
// (same per-thread order and reduction trees as the two separate passes: same bits)
// stored != nullptr: the caller accepts the results in the pinned mailbox (*stored = 1) instead of read_scalars
int norm_b_and_r_async(ny_mg* mg, cudaStream_t st, int* stored = nullptr);

// sum(msk*b^2) of level 1 -> slot 0
int norm_b_async(ny_mg* mg, cudaStream_t st)
{
    Level& L = mg->lev[0];
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
    int nparts;
    if (mg->box) {
        March m = march_geom(mg, L.nx, L.ny, L.nz - 2 * mg->nh);
        k_resid<RS_NORMB><<<m.grid, dim3(32, 8), 0, st>>>(nullptr, L.b, nullptr, box_of(mg, L), m.chunk, mg->d_red);
        nparts = m.nparts;
    } else {
        nparts = 1184;
        k_norm_partial<<<nparts, 256, 0, st>>>(L.msk, L.b, L, mg->nh, mg->d_red);
    }
    LAUNCH_OK(mg);
    return finish_sum(mg, st, nparts, 0);
}

// residual(1) followed by sum(msk*r^2) -> slot 1
int norm_r_async(ny_mg* mg, cudaStream_t st)
{
    Level& L = mg->lev[0];
    int nparts;
    if (mg->box) {
        ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
        March m = march_geom(mg, L.nx, L.ny, L.nz - 2 * mg->nh);
        k_resid<RS_NORM><<<m.grid, dim3(32, 8), 0, st>>>(L.x, L.b, nullptr, box_of(mg, L), m.chunk, mg->d_red);
        LAUNCH_OK(mg);
        nparts = m.nparts;
        return finish_sum(mg, st, nparts, 1);
    }
    TRY(residual(mg, st, 1));
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
    nparts = 1184;
    k_norm_partial<<<nparts, 256, 0, st>>>(L.msk, L.r, L, mg->nh, mg->d_red);
    LAUNCH_OK(mg);
    return finish_sum(mg, st, nparts, 1);
}

int norm_b_and_r_async(ny_mg* mg, cudaStream_t st, int* stored)
{
    Level& L = mg->lev[0];
    March m = march_geom(mg, L.nx, L.ny, L.nz - 2 * mg->nh);
    if (!mg->box || 2 * m.nparts > MAX_PARTIALS) {
        TRY(norm_b_async(mg, st));
        return norm_r_async(mg, st);
    }
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
    k_resid<RS_NORM2><<<m.grid, dim3(32, 8), 0, st>>>(L.x, L.b, nullptr, box_of(mg, L), m.chunk, mg->d_red);
    LAUNCH_OK(mg);
    if (!mg->comm && stored) {                 // one rank: both sums and their way to the host in one launch
        // (the partials of sum r^2 come first, those of sum b^2 follow: result slots 1 and 0)
        k_sum_final_store<<<2, 256, 0, st>>>(mg->d_red, m.nparts, mg->d_red + MAX_PARTIALS, mg->ctx->h_pinned);
        LAUNCH_OK(mg);
        *stored = 1;
        return NY_OK;
    }
    TRY(finish_sum(mg, st, m.nparts, 0, m.nparts));
    return finish_sum(mg, st, m.nparts, 1);
}

// ---- fused V-cycle legs (ny_mg_vleg.cuh) ---------------------------------------------------------
// A level runs the fused legs when the analytic box coefficients are valid, its interior is at least
// as wide as the halo in every direction (so that all three halo rings are plain copies) and -- for
// level 1, whose x and b come from the caller -- its periodic / slab halos are known to be consistent.
bool leg_ok(const ny_mg* mg, int lev)
{
    if (!mg->box || !mg->fused || lev >= mg->nlevels) return false;
    const Level& L = mg->lev[lev - 1];
    if (!mg->maps[lev - 1].ok || !mg->maps[lev].ok) return false;
    if (L.nx < 4 || L.ny < 4 || L.nz - 2 * NH < 4 || ((L.nz - 2 * NH) & 1)) return false;
    if ((long long)L.nx * L.ny * (L.nz - 2 * NH) < mg->leg_min_cells) return false;
    const bool walls_only = !mg->xper && !mg->yper && !L.zlo && !L.zhi;
    return lev > 1 || walls_only || mg->halo_ok;
}

// y must agree with x on the wall halos before a leg writes its result into y and swaps
int sync_y(ny_mg* mg, cudaStream_t st)
{
    if (mg->ysync) return NY_OK;
    for (int l = 0; l < mg->nlevels; l++)
        NY_CUDA(cudaMemcpyAsync(mg->lev[l].y, mg->lev[l].x, mg->lev[l].n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    mg->ysync = 1;
    return NY_OK;
}

struct LegGeom { dim3 grid; int chunk; int nparts; };
// launch geometry for the interior planes [0, nzi) of a level cut into chunks
inline LegGeom leg_geom(ny_mg* mg, const Level& L, int nzi, int tj, int extra, bool even, long long ntiles = 0)
{
    LegGeom q;
    const int gx = (L.nx + VL_TI - 1) / VL_TI, gy = (L.ny + tj - 1) / tj;
    const long long tiles = ntiles > 0 ? ntiles : (long long)gx * gy, sms = mg->ctx->num_sms;
    // chunks of planes: minimise (waves of CTAs) x (planes marched per CTA, including the pipeline fill)
    int best = 1;
    long long best_cost = -1;
    for (int nc = 1; nc <= 64 && nc * 8 <= (nzi > 8 ? nzi : 8); nc++) {
        int chunk = (nzi + nc - 1) / nc;
        if (even && (chunk & 1)) chunk++;
        const int ncc = (nzi + chunk - 1) / chunk;
        const long long waves = (tiles * ncc + sms - 1) / sms;
        const long long cost = waves * (chunk + extra);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = chunk; }
    }
    q.chunk = best;
    const int gz = (nzi + best - 1) / best;
    q.grid = dim3(gx, gy, gz);
    q.nparts = (int)(tiles * gz);
    return q;
}

// Tiles of a level whose 64 x 32 region keeps clear of the x / y walls and whose coarse tile lies inside the
// coarse domain (the INT instance of k_vleg): the rectangle [bx0, bx0+nbx) x [by0, by0+nby) of the tile grid.
inline TileMap interior_tiles(const ny_mg* mg, const Level& L, int tj, int apron_top)
{
    TileMap m;
    const int gx = (L.nx + VL_TI - 1) / VL_TI, gy = (L.ny + tj - 1) / tj;
    m.mode = 1; m.gx = gx;
    int x0 = 0, x1 = gx, y0 = 0, y1 = gy;
    if (!mg->xper) {
        // columns RX0 .. RX0+VL_RI-1 and both neighbours inside [NH, nx+NH): RX0 >= NH+1, RX0+VL_RI <= nx+NH-1
        x0 = gx; x1 = 0;
        for (int b = 0; b < gx; b++) {
            const int RX0 = b * VL_TI;
            if (RX0 >= NH + 1 && RX0 + VL_RI <= L.nx + NH - 1) { if (b < x0) x0 = b; x1 = b + 1; }
        }
    }
    if (!mg->yper) {
        y0 = gy; y1 = 0;
        for (int b = 0; b < gy; b++) {
            const int RY0 = NH + b * tj - apron_top;
            if (RY0 >= NH + 1 && RY0 + VL_RJ <= L.ny + NH - 1) { if (b < y0) y0 = b; y1 = b + 1; }
        }
    }
    m.bx0 = x0; m.nbx = x1 > x0 ? x1 - x0 : 0;
    m.by0 = y0; m.nby = y1 > y0 ? y1 - y0 : 0;
    return m;
}

constexpr int LEG_EDGE = 8;      // planes next to a slab neighbour that are computed first (even, >= 2 * NH)

// One leg of level lev.  On slabs without periodic wraps the planes next to the slab neighbours are
// computed first; their exchange (x' and, after a down leg, the coarse b) then runs on the communicator's
// own stream while the rest of the slab is computed.  *exchanged tells the caller that the halos are done.
template <bool PRO, int POST>
int launch_leg(ny_mg* mg, cudaStream_t st, int lev, const Level& V, bool* exchanged)
{
    using LY = VlegLayout<PRO, POST>;
    static bool attr_set = false;
    if (!attr_set) {
        NY_CUDA(cudaFuncSetAttribute(k_vleg<PRO, POST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::bytes));
        NY_CUDA(cudaFuncSetAttribute(k_vleg<PRO, POST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::bytes));
        attr_set = true;
    }
    Level& F = mg->lev[lev - 1];
    Level& C = mg->lev[lev];
    LevelMaps& M = mg->maps[lev - 1];
    const int nzi = F.nz - 2 * NH;
    const int extra = POST != POST_NONE ? 6 : 4;
    const double omega = mg->omega, cff1 = 1.0 - omega;
    const Box gf = box_of(mg, F), gv = box_of(mg, V);
    const CUtensorMap& tmc = PRO ? mg->maps[lev].cx : M.x;
    int nparts = 0;
    // tiles away from the x / y walls run the INT instance, the frame around them the general one
    TileMap inner = interior_tiles(mg, F, LY::tj, LY::apron_top);
    const int tgx = (F.nx + VL_TI - 1) / VL_TI, tgy = (F.ny + LY::tj - 1) / LY::tj;
    const long long n_inner = (long long)inner.nbx * inner.nby, n_frame = (long long)tgx * tgy - n_inner;
    auto launch = [&](int kz0, int kz1, cudaStream_t s) -> int {
        // (two launches only pay where each fills the machine several times: the large levels)
        if (n_inner > 0 && mg->split_tiles && n_inner + n_frame >= mg->split_tiles_min) {
            LegGeom q = leg_geom(mg, F, kz1 - kz0, LY::tj, extra, POST == POST_RESTRICT, n_inner);
            LegGeom qf = leg_geom(mg, F, kz1 - kz0, LY::tj, extra, POST == POST_RESTRICT, n_frame > 0 ? n_frame : 1);
            if (POST == POST_NORM && nparts + q.nparts + qf.nparts > MAX_PARTIALS) { ny_set_error("too many partial sums"); return NY_ERR_ARG; }
            k_vleg<PRO, POST, true><<<dim3(inner.nbx, inner.nby, q.grid.z), VL_NW * 32, LY::bytes, s>>>(
                M.x, M.b, tmc, F.y, V.b, mg->d_red + nparts, gf, gv, omega, cff1, q.chunk, kz0, kz1, inner);
            LAUNCH_OK(mg);
            nparts += q.nparts;
            if (n_frame > 0) {
                TileMap fr = inner;
                fr.mode = 2;
                k_vleg<PRO, POST, false><<<dim3((unsigned)n_frame, 1, qf.grid.z), VL_NW * 32, LY::bytes, s>>>(
                    M.x, M.b, tmc, F.y, V.b, mg->d_red + nparts, gf, gv, omega, cff1, qf.chunk, kz0, kz1, fr);
                LAUNCH_OK(mg);
                nparts += qf.nparts;
            }
            return NY_OK;
        }
        const LegGeom q = leg_geom(mg, F, kz1 - kz0, LY::tj, extra, POST == POST_RESTRICT);
        if (POST == POST_NORM && nparts + q.nparts > MAX_PARTIALS) { ny_set_error("too many partial sums"); return NY_ERR_ARG; }
        TileMap all = inner;
        all.mode = 0;
        k_vleg<PRO, POST, false><<<q.grid, VL_NW * 32, LY::bytes, s>>>(M.x, M.b, tmc, F.y, V.b, mg->d_red + nparts, gf, gv, omega,
                                                                       cff1, q.chunk, kz0, kz1, all);
        LAUNCH_OK(mg);
        nparts += q.nparts;
        return NY_OK;
    };
    *exchanged = false;
    const bool lo = mg->below >= 0, hi = mg->above >= 0;
    const bool coarse_dist = POST != POST_RESTRICT || !C.gathered;
    // worth it only where the interior part runs much longer than the exchange (measured: level 1 of a 512^3
    // slab yes, its 256^3-per-8 coarser levels no)
    // (and the faces are large: 1024^2 planes gain 3 ms per step at 8 GPUs, 512^2 planes lose 0.5 ms at 2)
    const bool big = (long long)F.nx * F.ny * nzi >= mg->overlap_cells && (long long)F.nx * F.ny * 32 >= mg->overlap_cells;
    if (mg->comm && !F.gathered && (lo || hi) && !mg->xper && !mg->yper && coarse_dist && nzi >= 4 * LEG_EDGE && big) {
        ny_comm* cm = mg->comm;
        if (lo) TRY(launch(0, LEG_EDGE, st));
        if (hi) TRY(launch(nzi - LEG_EDGE, nzi, st));
        NY_CUDA(cudaEventRecord(cm->ev_ready, st));
        NY_CUDA(cudaStreamWaitEvent(cm->xstream, cm->ev_ready, 0));
        {
            ny_prof_scope ps(mg->ctx, NY_PROF_HALO, cm->xstream);
            if (POST == POST_RESTRICT)
                TRY(ny_comm_exchange_z2(cm, F.y, (size_t)F.sk, nzi, C.b, (size_t)C.sk, C.nz - 2 * NH, NH, mg->below, mg->above,
                                        cm->xstream));
            else {
                double* arr[1] = {F.y};
                TRY(ny_comm_exchange_z(cm, arr, 1, (size_t)F.sk, NH, nzi, NH, mg->below, mg->above, cm->xstream));
            }
        }
        NY_CUDA(cudaEventRecord(cm->ev_done, cm->xstream));
        TRY(launch(lo ? LEG_EDGE : 0, hi ? nzi - LEG_EDGE : nzi, st));
        NY_CUDA(cudaStreamWaitEvent(st, cm->ev_done, 0));
        *exchanged = true;
    } else {
        TRY(launch(0, nzi, st));
    }
    swap_xy(mg, lev);
    return POST == POST_NORM ? nparts : 0;
}

// smooth(lev); residual(lev); restriction(lev)   (solvers.f90:41-46)
int down_leg(ny_mg* mg, cudaStream_t st, int lev)
{
    Level& F = mg->lev[lev - 1];
    Level& C = mg->lev[lev];
    Level V = coarse_view(mg, lev);
    TRY(sync_y(mg, st));
    bool exchanged = false;
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_DOWN_FINE : NY_PROF_MG_COARSE, st);
        NY_CUDA(cudaMemsetAsync(C.x, 0, C.n * sizeof(double), st));           // operators.f90:209
        int r = launch_leg<false, POST_RESTRICT>(mg, st, lev, V, &exchanged);
        if (r < 0) return r;
    }
    if (exchanged) return NY_OK;
    const bool slabs = mg->below >= 0 || mg->above >= 0;
    if (slabs && !F.gathered && !C.gathered && !mg->xper && !mg->yper) {
        // both levels are distributed and there is nothing to wrap locally: one exchange for x and b_coarse
        ny_prof_scope ps(mg->ctx, NY_PROF_HALO, st);
        return ny_comm_exchange_z2(mg->comm, F.x, (size_t)F.sk, F.nz - 2 * NH, C.b, (size_t)C.sk, C.nz - 2 * NH, NH,
                                   mg->below, mg->above, st);
    }
    TRY(fill(mg, st, F, F.x));
    TRY(gather_after_restriction(mg, st, lev));
    return fill(mg, st, C, C.b);
}

// prolongation(lev); smooth(lev)   (solvers.f90:51-54); with_norm: also sum r^2 of the result -> partials
int up_leg(ny_mg* mg, cudaStream_t st, int lev, bool with_norm, int* nparts)
{
    Level& F = mg->lev[lev - 1];
    Level V = coarse_view(mg, lev);
    TRY(sync_y(mg, st));
    bool exchanged = false;
    {
        ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_UP_FINE : NY_PROF_MG_COARSE, st);
        int r = with_norm ? launch_leg<true, POST_NORM>(mg, st, lev, V, &exchanged)
                          : launch_leg<true, POST_NONE>(mg, st, lev, V, &exchanged);
        if (r < 0) return r;
        if (nparts) *nparts = r;
    }
    if (exchanged) return NY_OK;
    return fill(mg, st, F, F.x);
}

// first level (1-based) of the V-cycle tail that k_vcycle_tail runs, or nlevels + 1 when there is none:
// box on analytic coefficients (closed or periodic), levels replicated on every rank, at most tail_cells cells each
// (narrow levels, one CTA) or, above those, at most wide_cells cells (wide levels, the whole co-resident grid)
inline long long level_cells(const Level& L) { return (long long)L.nx * L.ny * (L.nz - 2 * NH); }

int tail_first(const ny_mg* mg)
{
    const int none = mg->nlevels + 1;
    if (!mg->box || !mg->fused || !mg->tail || mg->tail_cells <= 0) return none;
    const long long wide = mg->wide_max_blocks > 0 ? mg->wide_cells : 0;
    int lt = mg->nlevels;
    while (lt >= 1) {
        const Level& L = mg->lev[lt - 1];
        // replicated levels only (their z ends are walls or a local periodic wrap, never a slab neighbour)
        if (!L.gathered || L.zlo != mg->zper || L.zhi != mg->zper) break;
        const bool narrow = level_cells(L) <= mg->tail_cells;
        if (!narrow && level_cells(L) > wide) break;
        // tiny periodic levels stage their self-overlapping halo sections through the shared memory of one CTA
        const bool wz = mg->zper;
        const bool simple = (!mg->xper || L.nx >= NH) && (!mg->yper || L.ny >= NH) && (!wz || L.nz - 2 * NH >= NH);
        if (!simple && (!narrow || L.n > (size_t)TAIL_STAGE)) break;
        lt--;
    }
    lt++;
    if (lt > mg->nlevels) return none;
    if (mg->nlevels - lt + 1 > TAIL_MAX) lt = mg->nlevels - TAIL_MAX + 1;
    return lt;
}

// norm_parts != nullptr and the tail starts at level 1: it also leaves sum r^2 of level 1 as *norm_parts partial sums
int vcycle_tail(ny_mg* mg, cudaStream_t st, int lt, int* norm_parts)
{
    TailArgs a;
    a.n = mg->nlevels - lt + 1;
    a.nwide = 0;
    a.bar = mg->d_bar;
    // (one rank: a multigrid whose first level is replicated has no other slabs to sum over)
    a.norm_partial = norm_parts && lt == 1 && !mg->comm ? mg->d_red : nullptr;
    a.norm_out = mg->d_red + MAX_PARTIALS + 1;
    a.norm_host = mg->ctx->h_pinned + 1;
    a.omega = mg->omega; a.cff1 = 1.0 - mg->omega;
    long long most = 0;
    for (int l = 0; l < a.n; l++) {
        Level& L = mg->lev[lt - 1 + l];
        a.lev[l].x = L.x; a.lev[l].y = L.y; a.lev[l].b = L.b; a.lev[l].g = box_of(mg, L);
        const int widen[3] = {0, 2, 2 * NH};
        for (int r = 0; r < 3; r++) {
            a.lev[l].fx[r] = fastdiv_of(L.nx + widen[r]);
            a.lev[l].fy[r] = fastdiv_of(L.ny + widen[r]);
        }
        if (level_cells(L) > mg->tail_cells) {              // level sizes fall with l: the wide levels lead
            a.nwide = l + 1;
            const long long ring = (long long)(L.nx + 2) * (L.ny + 2) * (L.nz - 2 * NH + 2);
            if (ring > most) most = ring;
        }
    }
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_COARSE, st);
    if (a.nwide == 0) {
        k_vcycle_tail<<<1, TAIL_THREADS, 0, st>>>(a);
        LAUNCH_OK(mg);
        if (a.norm_partial) *norm_parts = -1;
        return NY_OK;
    }
    long long blocks = (most + TAIL_THREADS - 1) / TAIL_THREADS;
    if (blocks > mg->wide_max_blocks) blocks = mg->wide_max_blocks;
    if (blocks < 2) blocks = 2;
    void* args[] = {&a};
    NY_CUDA(cudaLaunchCooperativeKernel((const void*)k_vcycle_tail, dim3((unsigned)blocks), dim3(TAIL_THREADS), args, 0, st));
    mg->ctx->launches++;
    if (a.norm_partial) *norm_parts = -1;
    return NY_OK;
}

// solvers.f90:35-55.  norm_parts != nullptr: the caller wants sum r^2 of level 1 after the cycle; if the
// last leg could provide it, *norm_parts = number of partial sums waiting in d_red; -1: the one-launch tail was the
// whole cycle and has left the finished sum in result slot 1 and in the pinned mailbox; else 0.
int vcycle(ny_mg* mg, cudaStream_t st, int* norm_parts = nullptr)
{
    int lev1 = mg->nlevels - 1;
    if (norm_parts) *norm_parts = 0;
    const int lt = tail_first(mg);
    if (lt <= mg->nlevels) lev1 = lt - 1;                 // levels lt .. nlevels: one launch
    for (int lev = 1; lev <= lev1; lev++) {
        if (leg_ok(mg, lev)) { TRY(down_leg(mg, st, lev)); continue; }
        TRY(smooth(mg, st, lev));
        TRY(residual_restriction(mg, st, lev));
    }
    if (lt <= mg->nlevels) TRY(vcycle_tail(mg, st, lt, norm_parts));
    else TRY(smooth(mg, st, lev1 + 1));
    for (int lev = lev1; lev >= 1; lev--) {
        if (leg_ok(mg, lev)) { TRY(up_leg(mg, st, lev, lev == 1 && norm_parts, norm_parts)); continue; }
        TRY(prolongation(mg, st, lev));
        TRY(smooth(mg, st, lev));
    }
    return NY_OK;
}

// operators.f90:461-505 on one rank, through the generic kernels
int setup_operators(ny_mg* mg, cudaStream_t st)
{
    const int nl = mg->nlevels;
    TRY(fill(mg, st, mg->lev[0], mg->lev[0].msk));               // setup_fine_msk, mg_setup.f90:213-223
    for (int lev = 1; lev <= nl - 1; lev++) {                    // compute_msk, operators.f90:299-335
        Level &F = mg->lev[lev - 1], &C = mg->lev[lev];
        TRY(ew<EW_SET>(mg, st, C.msk, nullptr, nullptr, 1.0, C.n));
        TRY(ew<EW_SET>(mg, st, C.Rcoef, nullptr, nullptr, 1.0, C.n));
        TRY(ew<EW_COPY>(mg, st, F.b, F.msk, nullptr, 0.0, F.n));
        TRY(restriction(mg, st, lev, true));
        TRY(ew<EW_GT0_ONE>(mg, st, C.msk, C.b, nullptr, 0.0, C.n));
        k_apply_default_msk<<<(unsigned)((C.n + 255) / 256), 256, 0, st>>>(C.msk, C, mg->nh, mg->xper, mg->yper);
        LAUNCH_OK(mg);
        TRY(ew<EW_SET>(mg, st, C.y, nullptr, nullptr, 0.0, C.n));
    }
    for (int lev = 1; lev <= nl - 1; lev++) {
        Level &F = mg->lev[lev - 1], &C = mg->lev[lev];
        // compute_Rcoef, operators.f90:337-393
        TRY(ew<EW_COPY>(mg, st, C.Rcoef, C.msk, nullptr, 0.0, C.n));
        TRY(ew<EW_SET>(mg, st, F.b, nullptr, nullptr, 1.0, F.n));
        TRY(restriction(mg, st, lev, true));
        TRY(ew<EW_RECIP_GT0>(mg, st, C.y, C.b, nullptr, 4.0, C.n));
        TRY(ew<EW_MUL>(mg, st, C.Rcoef, C.msk, C.y, 0.0, C.n));
        TRY(ew<EW_COPY>(mg, st, C.y, C.Rcoef, nullptr, 0.0, C.n));
        // compute_Pcoef, operators.f90:395-424
        TRY(ew<EW_SET>(mg, st, F.x, nullptr, nullptr, 0.0, F.n));
        TRY(ew<EW_COPY>(mg, st, C.x, C.msk, nullptr, 0.0, C.n));
        TRY(ew<EW_COPY>(mg, st, F.Pcoef, F.msk, nullptr, 0.0, F.n));
        TRY(prolongation(mg, st, lev));
        TRY(ew<EW_RECIP_GT0>(mg, st, F.y, F.x, nullptr, 1.0, F.n));
        TRY(ew<EW_MUL>(mg, st, F.Pcoef, F.msk, F.y, 0.0, F.n));
    }
    for (int lev = 1; lev <= nl; lev++) {                        // compute_diag, operators.f90:426-457
        Level& L = mg->lev[lev - 1];
        TRY(ew<EW_COPY>(mg, st, L.x, L.msk, nullptr, 0.0, L.n));
        TRY(ew<EW_SET>(mg, st, L.b, nullptr, nullptr, 0.0, L.n));
        TRY(ew<EW_SET>(mg, st, L.diag, nullptr, nullptr, 0.0, L.n));
        TRY(residual(mg, st, lev));
        TRY(ew<EW_NEG>(mg, st, L.diag, L.r, nullptr, 0.0, L.n));
        TRY(ew<EW_RECIP_GT0>(mg, st, L.idiag, L.diag, nullptr, 1.0, L.n));
        TRY(ew<EW_SET>(mg, st, L.x, nullptr, nullptr, 0.0, L.n));
    }
    return NY_OK;
}

// decide whether the analytic box coefficients reproduce the arrays exactly
int verify_box(ny_mg* mg, cudaStream_t st)
{
    NY_CUDA(cudaMemsetAsync(mg->d_flag, 0, sizeof(int), st));
    for (int l = 0; l < mg->nlevels; l++) {
        Level& L = mg->lev[l];
        const bool has_c = l + 1 < mg->nlevels;
        Box gc = box_of(mg, mg->lev[has_c ? l + 1 : l]);
        k_check_box<<<(unsigned)((L.n + 255) / 256), 256, 0, st>>>(L, box_of(mg, L), has_c ? 1 : 0, l > 0 ? 1 : 0, gc, mg->d_flag);
        LAUNCH_OK(mg);
    }
    int flag = 1;
    NY_CUDA(cudaMemcpyAsync(&flag, mg->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    NY_CUDA(cudaStreamSynchronize(st));
    mg->box = flag == 0;
    return NY_OK;
}

double* var_ptr(ny_mg* mg, int lev, int ivar)
{
    Level& L = mg->lev[lev - 1];
    switch (ivar) {
    case NY_MG_X: return L.x; case NY_MG_B: return L.b; case NY_MG_R: return L.r; case NY_MG_Y: return L.y;
    case NY_MG_DIAG: return L.diag; case NY_MG_IDIAG: return L.idiag; case NY_MG_MSK: return L.msk;
    case NY_MG_RCOEF: return L.Rcoef; case NY_MG_PCOEF: return L.Pcoef; default: return nullptr;
    }
}

int create(ny_ctx* ctx, ny_comm* comm, int nx, int ny, int nz_global, int topology, ny_mg** out)
{
    NY_REQUIRE(ctx && out, "null argument");
    NY_REQUIRE(nx >= 2 && ny >= 2 && nz_global >= 1, "grid too small");
    NY_REQUIRE(topology >= NY_TOPO_CLOSED && topology <= NY_TOPO_XYZPERIO, "unknown topology");
    const int P = comm ? comm->nranks : 1, rank = comm ? comm->rank : 0;
    NY_REQUIRE(nz_global % P == 0, "global nz must be a multiple of the number of slabs");
    {   // experiment switches (bench runs under torchrun): NY_MG_OVERLAP_CELLS, NY_MG_TAIL_CELLS, NY_MG_GATHER_CELLS
        const char* e = getenv("NY_MG_OVERLAP_CELLS");
        if (e && *e) g_overlap_cells = atoll(e);
        e = getenv("NY_MG_TAIL_CELLS");
        if (e && *e) g_tail_cells = atoll(e);
        e = getenv("NY_MG_WIDE_CELLS");
        if (e && *e) g_wide_cells = atoll(e);
        e = getenv("NY_MG_GATHER_CELLS");
        if (e && *e) g_gather_cells = atoll(e);
        e = getenv("NY_MG_LEG_MIN_CELLS");
        if (e && *e) g_leg_min_cells = atoll(e);
    }
    ny_mg* mg = new ny_mg();
    memset(mg, 0, sizeof(ny_mg));
    mg->ctx = ctx; mg->comm = P > 1 ? comm : nullptr; mg->nranks = P; mg->rank = rank;
    mg->tail_cells = g_tail_cells; mg->split_tiles_min = g_split_tiles; mg->overlap_cells = g_overlap_cells;
    mg->leg_min_cells = g_leg_min_cells; mg->wide_cells = g_wide_cells;
    mg->nh = NH; mg->maxite = 20; mg->tol = 1e-6; mg->omega = 0.9;          // mg_types.f90:15-26
    mg->topology = topology;
    mg->xper = topology == NY_TOPO_XPERIO || topology == NY_TOPO_XYPERIO || topology == NY_TOPO_XYZPERIO;
    mg->yper = topology == NY_TOPO_YPERIO || topology == NY_TOPO_XYPERIO || topology == NY_TOPO_XYZPERIO;
    mg->zper = topology == NY_TOPO_ZPERIO || topology == NY_TOPO_XYZPERIO;
    mg->below = rank > 0 ? rank - 1 : (mg->zper && P > 1 ? P - 1 : -1);
    mg->above = rank < P - 1 ? rank + 1 : (mg->zper && P > 1 ? 0 : -1);
    const int nh = mg->nh;
    // create_hierarchy, mg_setup.f90:225-307, on the GLOBAL grid.  The Fortran loops for ever (and
    // overruns its level table) when neither nx nor ny passes through 2; refuse instead.
    int gx[MAXLEV], gy[MAXLEV], gz[MAXLEV];
    int x = nx, y = ny, z = nz_global, i = 0;
    gx[0] = x; gy[0] = y; gz[0] = z;
    while (!(x == 2 || y == 2)) {
        if ((x & 1) || (y & 1) || x < 2 || y < 2 || (z & 1) || z < 2 || i + 1 >= MAXLEV) {
            ny_set_error("ny_mg_create: %dx%dx%d cannot be halved down to nx==2 or ny==2 with even extents "
                         "(core/mgfor/mg_setup.f90:262-303)", nx, ny, nz_global);
            delete mg;
            return NY_ERR_GRID;
        }
        x /= 2; y /= 2; z /= 2;
        i++;
        gx[i] = x; gy[i] = y; gz[i] = z;
    }
    mg->nlevels = i + 1;
    // slabs: a level stays distributed while its slab is at least 4 planes thick and the level has
    // more than g_gather_cells cells (the finest level is distributed whatever its size); from the first level that
    // is not, everything is gathered
    mg->glev = 0;
    if (P > 1) {
        int l = 0;
        while (l < mg->nlevels && gz[l] % P == 0 && gz[l] / P >= 4 && (l == 0 || (long long)gx[l] * gy[l] * gz[l] > g_gather_cells)) l++;
        mg->glev = l;
        if (l == 0) {
            ny_set_error("ny_mg_create: the finest level (%dx%dx%d on %d slabs) is too small to be distributed",
                         nx, ny, nz_global, P);
            delete mg;
            return NY_ERR_GRID;
        }
    }
    size_t tmp_need = 1;
    for (int l = 0; l < mg->nlevels; l++) {
        Level& L = mg->lev[l];
        L.gathered = l >= mg->glev;
        L.nx = gx[l]; L.ny = gy[l];
        L.nz = (L.gathered ? gz[l] : gz[l] / P) + 2 * nh;
        L.zlo = L.gathered ? mg->zper : (rank > 0 || mg->zper);
        L.zhi = L.gathered ? mg->zper : (rank < P - 1 || mg->zper);
        L.sj = L.nx + 2 * nh;
        L.sk = L.sj * (L.ny + 2 * nh);
        L.n = (size_t)L.sk * (size_t)L.nz;
        double** all[] = {&L.x, &L.b, &L.r, &L.y, &L.diag, &L.idiag, &L.Rcoef, &L.Pcoef, &L.msk};
        const int nalloc = P > 1 ? 4 : 9;                       // slabs use the analytic coefficients only
        for (int a = 0; a < nalloc; a++) {
            cudaError_t e = cudaMalloc(all[a], L.n * sizeof(double));
            if (e != cudaSuccess) {
                ny_set_error("ny_mg_create: cudaMalloc of level %d failed: %s", l + 1, cudaGetErrorString(e));
                ny_mg_destroy(mg);
                return NY_ERR_CUDA;
            }
            cudaMemset(*all[a], 0, L.n * sizeof(double));       // Fortran leaves these uninitialised
        }
        // staging is only needed where a periodic section can overlap itself (tiny levels)
        if (L.nx < nh || L.ny < nh || (L.nz - 2 * nh) < nh) {
            size_t need = (size_t)L.sk * nh;
            size_t need2 = (size_t)nh * (L.ny + 2 * nh) * L.nz;
            size_t need3 = (size_t)nh * (L.nx + 2 * nh) * L.nz;
            if (need > tmp_need) tmp_need = need;
            if (need2 > tmp_need) tmp_need = need2;
            if (need3 > tmp_need) tmp_need = need3;
        }
    }
    mg->fused = 1;
    mg->split_tiles = 1;
    mg->tail = 1;
    mg->halo_ok = 0;
    mg->ysync = 1;                                              // x and y are both zero
    for (int l = 0; l < mg->nlevels; l++) {
        Level& L = mg->lev[l];
        LevelMaps& M = mg->maps[l];
        const int tx = L.nx + 2 * nh, ty = L.ny + 2 * nh;
        int r = ny_tma_encode_3d(&M.x, L.x, tx, ty, L.nz, VL_RI, VL_RJ, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&M.y, L.y, tx, ty, L.nz, VL_RI, VL_RJ, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&M.b, L.b, tx, ty, L.nz, VL_RI, VL_RJ, 1);
        // the coarse tile of this level's x as the finer level sees it
        int wnz = L.nz;
        size_t woff = 0;
        if (l > 0 && !mg->lev[l - 1].gathered && L.gathered) {
            const int nzc = (mg->lev[l - 1].nz - 2 * nh) / 2;
            wnz = nzc + 2 * nh;
            woff = (size_t)rank * nzc * L.sk;
        }
        if (r == NY_OK) r = ny_tma_encode_3d(&M.cx, L.x + woff, tx, ty, wnz, VL_CI, VL_CJ, 1);
        if (r == NY_OK) r = ny_tma_encode_3d(&M.cy, L.y + woff, tx, ty, wnz, VL_CI, VL_CJ, 1);
        M.ok = r == NY_OK;
    }
    mg->tmp_doubles = tmp_need;
    if (cudaMalloc(&mg->tmp, tmp_need * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&mg->d_red, (MAX_PARTIALS + 4) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&mg->d_flag, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&mg->d_bar, 2 * sizeof(unsigned)) != cudaSuccess ||
        cudaMemset(mg->d_bar, 0, 2 * sizeof(unsigned)) != cudaSuccess) {
        ny_set_error("ny_mg_create: cudaMalloc failed");
        ny_mg_destroy(mg);
        return NY_ERR_CUDA;
    }
    {   // how many CTAs of the tail kernel are resident at once: the grid barrier of its wide levels needs them all
        int dev = 0, coop = 0, per_sm = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_vcycle_tail, TAIL_THREADS, 0) == cudaSuccess)
            mg->wide_max_blocks = per_sm > 0 ? ctx->num_sms : 0;              // one CTA per SM
        cudaGetLastError();
    }
    cudaStream_t st = 0;
    int r = NY_OK;
    if (P == 1) {
        for (int l = 0; l < mg->nlevels; l++) {
            Level& L = mg->lev[l];
            k_default_msk<<<(unsigned)((L.n + 255) / 256), 256, 0, st>>>(L.msk, L, nh, mg->zper);
            ctx->launches++;
        }
        mg->box = 0;                                            // the setup itself runs on the generic kernels
        r = setup_operators(mg, st);
        if (r == NY_OK) r = verify_box(mg, st);
    } else {
        mg->box = 1;
    }
    if (r == NY_OK && cudaStreamSynchronize(st) != cudaSuccess) { ny_set_error("ny_mg_create: setup failed"); r = NY_ERR_CUDA; }
    if (r != NY_OK) { ny_mg_destroy(mg); return r; }
    *out = mg;
    return NY_OK;
}

}  // namespace

extern "C" void ny_mg_destroy(ny_mg* mg)
{
    if (!mg) return;
    for (int l = 0; l < mg->nlevels; l++) {
        Level& L = mg->lev[l];
        double* ptrs[] = {L.x, L.b, L.r, L.y, L.diag, L.idiag, L.Rcoef, L.Pcoef, L.msk};
        for (double* p : ptrs) if (p) cudaFree(p);
    }
    if (mg->tmp) cudaFree(mg->tmp);
    if (mg->d_red) cudaFree(mg->d_red);
    if (mg->d_flag) cudaFree(mg->d_flag);
    if (mg->d_bar) cudaFree(mg->d_bar);
    delete mg;
}

extern "C" int ny_mg_create(ny_ctx* ctx, int nx, int ny, int nz, int topology, ny_mg** out)
{
    return create(ctx, nullptr, nx, ny, nz, topology, out);
}

extern "C" int ny_mg_create_slab(ny_ctx* ctx, ny_comm* comm, int nx, int ny, int nz_global, int topology, ny_mg** out)
{
    return create(ctx, comm, nx, ny, nz_global, topology, out);
}

extern "C" void ny_mg_set_gather_cells(long long cells) { g_gather_cells = cells; }
extern "C" void ny_mg_set_overlap_cells(long long cells) { g_overlap_cells = cells; }
extern "C" void ny_mg_set_split_tiles(long long tiles) { g_split_tiles = tiles; }
extern "C" void ny_mg_set_tail_cells(long long cells) { g_tail_cells = cells; }
extern "C" void ny_mg_set_wide_cells(long long cells) { g_wide_cells = cells; }

extern "C" int ny_mg_nlevels(ny_mg* mg) { return mg ? mg->nlevels : 0; }
extern "C" int ny_mg_is_box(ny_mg* mg) { return mg ? mg->box : 0; }
extern "C" int ny_mg_first_gathered_level(ny_mg* mg) { return mg ? mg->glev + 1 : 0; }

extern "C" int ny_mg_set_fast_path(ny_mg* mg, int on)
{
    NY_REQUIRE(mg, "null argument");
    NY_REQUIRE(mg->nranks == 1, "slab multigrids always run the box kernels (they hold no coefficient arrays)");
    mg->ysync = 0;
    if (on) return verify_box(mg, 0);
    mg->box = 0;
    return NY_OK;
}

extern "C" int ny_mg_set_fused_legs(ny_mg* mg, int on)
{
    NY_REQUIRE(mg, "null argument");
    mg->fused = on ? 1 : 0;
    mg->split_tiles = on == 2 ? 0 : 1;      // 2: fused legs, every tile through the general kernel instance,
    mg->tail = on == 2 ? 0 : 1;             //    and no one-launch tail
    return NY_OK;
}

extern "C" int ny_mg_shape(ny_mg* mg, int lev, int shape[3])
{
    NY_REQUIRE(mg && shape && lev >= 1 && lev <= mg->nlevels, "bad level");
    Level& L = mg->lev[lev - 1];
    shape[0] = L.nz; shape[1] = L.ny + 2 * mg->nh; shape[2] = L.nx + 2 * mg->nh;
    return NY_OK;
}

extern "C" int ny_mg_set_param(ny_mg* mg, int maxite, double tol, double omega)
{
    NY_REQUIRE(mg && maxite >= 0 && maxite < 31, "bad parameter (maxite must be < 31)");
    mg->maxite = maxite; mg->tol = tol; mg->omega = omega;
    return NY_OK;
}

extern "C" int ny_mg_set_array(ny_mg* mg, int lev, int ivar, const double* src, void* stream)
{
    NY_REQUIRE(mg && src && lev >= 1 && lev <= mg->nlevels, "bad argument");
    NY_REQUIRE((ivar >= NY_MG_X && ivar <= NY_MG_Y) || ivar == NY_MG_MSK, "only x,b,r,y,msk can be set (pytools.f90:45-62)");
    NY_REQUIRE(var_ptr(mg, lev, ivar), "this multigrid holds no such array (slab multigrids keep x, b, r, y only)");
    NY_CUDA(cudaMemcpyAsync(var_ptr(mg, lev, ivar), src, mg->lev[lev - 1].n * sizeof(double),
                            cudaMemcpyDeviceToDevice, ny_stream(stream)));
    if (ivar == NY_MG_MSK) mg->box = 0;          // a user mask: coefficients are no longer those of a box
    if (ivar == NY_MG_X || ivar == NY_MG_Y) mg->ysync = 0;
    if (lev == 1 && (ivar == NY_MG_X || ivar == NY_MG_B)) mg->halo_ok = 0;
    return NY_OK;
}

// setup_fine_msk + setup_operators (mg_setup.f90:213-223, operators.f90:461-505) after the caller changed
// the mask of level 1 with ny_mg_set_array(.., NY_MG_MSK, ..): coarse masks, Rcoef, Pcoef, diag, idiag of
// every level are rebuilt by the reference's own algorithm; the analytic box kernels stay in use only if
// the result is still the default box (it is not, with an obstacle: the generic kernels then run).
extern "C" int ny_mg_setup_operators(ny_mg* mg, void* stream)
{
    NY_REQUIRE(mg, "null argument");
    NY_REQUIRE(mg->nranks == 1, "user masks need the coefficient arrays, which slab multigrids do not hold");
    cudaStream_t st = ny_stream(stream);
    mg->box = 0;
    int r = setup_operators(mg, st);
    if (r == NY_OK) r = verify_box(mg, st);
    mg->ysync = 0;
    mg->halo_ok = 0;
    return r;
}

extern "C" int ny_mg_get_array(ny_mg* mg, int lev, int ivar, double* dst, void* stream)
{
    NY_REQUIRE(mg && dst && lev >= 1 && lev <= mg->nlevels && var_ptr(mg, lev, ivar), "bad argument");
    NY_CUDA(cudaMemcpyAsync(dst, var_ptr(mg, lev, ivar), mg->lev[lev - 1].n * sizeof(double),
                            cudaMemcpyDeviceToDevice, ny_stream(stream)));
    return NY_OK;
}

// The stopping test reads two scalars per V-cycle.  They are stored into the pinned mailbox by a kernel
// (pinned host memory is device-addressable under unified addressing) instead of a cudaMemcpyAsync: a copy
// would queue in the device-to-host copy engine behind any bulk download that overlaps the solve
// (Nyles.step_host returns finished fields while the last projection runs) and stall every V-cycle.
__global__ void k_store_scalars(const double* __restrict__ src, double* dst_host, int n)
{
    if ((int)threadIdx.x < n) dst_host[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

static int read_scalars(ny_mg* mg, cudaStream_t st, int n)
{
    k_store_scalars<<<1, 32, 0, st>>>(mg->d_red + MAX_PARTIALS, mg->ctx->h_pinned, n);
    NY_CHECK_LAUNCH(mg->ctx);
    NY_CUDA(cudaStreamSynchronize(st));
    return NY_OK;
}

// solvers.f90:8-33
extern "C" int ny_mg_solve(ny_mg* mg, ny_mg_stats* stats, void* stream)
{
    NY_REQUIRE(mg, "null argument");
    cudaStream_t st = ny_stream(stream);
    int nite = 0, nres = 0;
    double hist[32];
    // normb = sum(msk b^2); res = sum(msk r^2)/normb after residual(1)   (operators.f90:81-125)
    // both are enqueued before the first host read: one synchronisation instead of two
    int stored = 0;
    TRY(norm_b_and_r_async(mg, st, &stored));
    if (stored) NY_CUDA(cudaStreamSynchronize(st));
    else TRY(read_scalars(mg, st, 2));
    const double normb = mg->ctx->h_pinned[0];
    double res = normb > 0.0 ? mg->ctx->h_pinned[1] / normb : 0.0;
    hist[nres++] = res;
    for (;;) {
        if (res < mg->tol) break;
        int parts = 0;
        const bool last = nite + 1 >= mg->maxite;            // no residual is evaluated after the last cycle
        TRY(vcycle(mg, st, last ? nullptr : &parts));
        nite++;
        if (nite >= mg->maxite) break;
        if (parts < 0) {                                      // ... finished, out of the one-launch cycle
            NY_CUDA(cudaStreamSynchronize(st));
        } else if (parts > 0 && !mg->comm) {                  // one rank: final sum and its way to the host in one launch
            {
                ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
                k_sum_final_store<<<1, 256, 0, st>>>(mg->d_red, parts, mg->d_red + MAX_PARTIALS + 1, mg->ctx->h_pinned + 1);
                LAUNCH_OK(mg);
            }
            NY_CUDA(cudaStreamSynchronize(st));
        } else {
            if (parts > 0) {                                  // sum r^2 came out of the last leg of the cycle
                ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
                TRY(finish_sum(mg, st, parts, 1));
            } else {
                TRY(norm_r_async(mg, st));
            }
            TRY(read_scalars(mg, st, 2));
        }
        res = mg->ctx->h_pinned[1] / normb;
        if (nres < 32) hist[nres++] = res;
    }
    if (stats) {
        stats->nite = nite; stats->nres = nres; stats->res = res; stats->normb = normb;
        for (int t = 0; t < 32; t++) stats->reshist[t] = t < nres ? hist[t] : 0.0;
    }
    return NY_OK;
}

extern "C" int ny_mg_solve_directly(ny_mg* mg, double* p, const double* div, ny_ext e, const int lo[3],
                                    double scale, ny_mg_stats* stats, void* stream)
{
    NY_REQUIRE(mg && p && div && lo, "null argument");
    Level& L = mg->lev[0];
    const int nh = mg->nh;
    NY_REQUIRE(lo[0] + e.nz <= L.nz && lo[1] + e.ny <= L.ny + 2 * nh && lo[2] + e.nx <= L.nx + 2 * nh &&
               lo[0] >= 0 && lo[1] >= 0 && lo[2] >= 0, "model array does not fit the multigrid array");
    cudaStream_t st = ny_stream(stream);
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    {
        ny_prof_scope ps(mg->ctx, NY_PROF_MG_EMBED, st);
        k_embed<<<g.grid, g.block, 0, st>>>(L.b, div, L, e.nz, e.ny, e.nx, lo[0], lo[1], lo[2]);
        LAUNCH_OK(mg);
    }
    // the caller filled the halo of div (mgfordriver.py:70), x is the halo-filled result of the last solve
    mg->halo_ok = 1;
    TRY(ny_mg_solve(mg, stats, stream));
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_EMBED, st);
    k_extract<<<g.grid, g.block, 0, st>>>(mg->lev[0].x, p, L, e.nz, e.ny, e.nx, lo[0], lo[1], lo[2], scale);
    LAUNCH_OK(mg);
    return NY_OK;
}

// compute_p (core/projection.py:39-87) in three launches around the solve: div from u (U = u*ids2 on
// the fly) written to `div` and embedded in b, halo fill of b, solve, p = x*scale and u -= delta p.
extern "C" int ny_mg_project(ny_mg* mg, double* ux, double* uy, double* uz, double* div, double* p,
                             double idx2, double idy2, double idz2, ny_ext e, const int lo[3], double scale,
                             ny_mg_stats* stats, void* stream)
{
    NY_REQUIRE(mg && ux && uy && uz && div && p && lo, "null argument");
    Level& L = mg->lev[0];
    const int nh = mg->nh;
    NY_REQUIRE(lo[0] + e.nz <= L.nz && lo[1] + e.ny <= L.ny + 2 * nh && lo[2] + e.nx <= L.nx + 2 * nh &&
               lo[0] >= 0 && lo[1] >= 0 && lo[2] >= 0, "model array does not fit the multigrid array");
    cudaStream_t st = ny_stream(stream);
    ny_grid3 g = ny_cells_launch(e.nz, e.ny, e.nx);
    {
        ny_prof_scope ps(mg->ctx, NY_PROF_DIV, st);
        k_div_embed<<<g.grid, g.block, 0, st>>>(ux, uy, uz, div, L.b, idx2, idy2, idz2, L, e.nz, e.ny, e.nx, lo[0], lo[1], lo[2]);
        LAUNCH_OK(mg);
    }
    {   // what halo.fill(div) before the embedding does in the reference (mgfordriver.py:70-72)
        ny_prof_scope ps(mg->ctx, NY_PROF_HALO, st);
        TRY(fill(mg, st, L, L.b));
    }
    mg->halo_ok = 1;
    TRY(ny_mg_solve(mg, stats, stream));
    ny_prof_scope ps(mg->ctx, NY_PROF_GRADP, st);
    k_extract_gradp<<<g.grid, g.block, 0, st>>>(mg->lev[0].x, p, ux, uy, uz, L, e.nz, e.ny, e.nx, lo[0], lo[1], lo[2], scale);
    LAUNCH_OK(mg);
    return NY_OK;
}

// (Merging "u -= grad p" with the diagnostics that follow it -- 15 arrays through HBM instead of 18 -- was built as a
// one-thread-per-cell kernel that recomputes the projected velocity of the 12 neighbour components it needs, verified
// bit-identical and measured: 3.93 ms per 512^3 launch against 1.59 + 1.83 ms for k_extract_gradp + k_diag_post, which
// run at 5.4-5.8 TB/s; the 22 neighbour loads per cell cost more in L1/L2 than the 24 B/cell saved in HBM.  Dropped.)
extern "C" int ny_mg_op(ny_mg* mg, int op, int lev, void* stream)
{
    NY_REQUIRE(mg && lev >= 1 && lev <= mg->nlevels, "bad level");
    cudaStream_t st = ny_stream(stream);
    switch (op) {
    case NY_MG_OP_SMOOTH: return smooth(mg, st, lev);
    case NY_MG_OP_RESIDUAL: return residual(mg, st, lev);
    case NY_MG_OP_RESTRICTION: NY_REQUIRE(lev < mg->nlevels, "no coarser level"); return restriction(mg, st, lev, false);
    case NY_MG_OP_PROLONGATION: NY_REQUIRE(lev < mg->nlevels, "no coarser level"); return prolongation(mg, st, lev);
    case NY_MG_OP_VCYCLE: return vcycle(mg, st);
    case NY_MG_OP_FILL: {
        Level& L = mg->lev[lev - 1];
        TRY(fill(mg, st, L, L.x));
        TRY(fill(mg, st, L, L.b));
        if (lev == 1) mg->halo_ok = 1;
        return NY_OK;
    }
    default: ny_set_error("ny_mg_op: unknown op %d", op); return NY_ERR_ARG;
    }
}
