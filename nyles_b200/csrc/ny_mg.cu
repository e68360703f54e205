// Geometric multigrid pressure solver: sm_100a replacement of libmgmod64.so (core/mgfor/*.f90).
//
// Same algorithm, same iterates as the reference (required for field parity at the loose
// stopping tolerance, SURVEY.md section 7 hard part 1):
//   hierarchy          mg_setup.f90:225-307      masks/Rcoef/Pcoef/diag   operators.f90:246-505
//   smoother           basicoperators.f90:363-400 (two damped-Jacobi sweeps, omega = 0.9)
//   residual           basicoperators.f90:300-323  restriction  :32-60   prolongation :173-231
//   norm               basicoperators.f90:422-440  V-cycle / solve  solvers.f90:8-55
//   halo fill          mod_halo.f90:235-262 (one process: periodic self exchange)
// Level arrays are (nz+2nh, ny+2nh, nx+2nh), i fastest, exactly the reference's padded layout.
// Arithmetic: source order, no FMA (-fmad=false) -> bit-identical to the oracle except for the
// summation order of the two norms, which only feed the stopping test.
#include "ny_common.cuh"

namespace {

constexpr int MAXLEV = 50;
constexpr int NORM_BLOCKS = 1184;          // 8 x 148 SMs

struct Level {
    int nx, ny, nz;                        // nz includes the 2*nh halo planes
    long long sj, sk;
    size_t n;
    double *x, *b, *r, *y, *diag, *idiag, *Rcoef, *Pcoef, *msk;
};

}  // namespace

struct ny_mg {
    ny_ctx* ctx;
    int nlevels, nh, topology, maxite;
    int xper, yper, zper;
    double tol, omega;
    Level lev[MAXLEV];
    double* tmp;                           // staging for overlapping halo sections
    size_t tmp_doubles;
    double* d_red;                         // NORM_BLOCKS partials + 2 results
};

namespace {

// Fortran indices (i in 1-nh..nx+nh, j likewise, k in 1..nz) -> linear offset
__host__ __device__ inline long long IX(const Level& L, int nh, int i, int j, int k)
{
    return (long long)(k - 1) * L.sk + (long long)(j - 1 + nh) * L.sj + (i - 1 + nh);
}

// ---- stencil kernels ---------------------------------------------------------------------
// one damped-Jacobi sweep over the Fortran index box [i0,i1]x[j0,j1]x[k0,k1]
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
    xc[c] = coef[c] * s;
}

// one thread per fine cell; (i,j) parity picks the coarse neighbours, k parity picks plane a or c
__global__ void __launch_bounds__(256)
k_prolong(double* __restrict__ xf, const double* __restrict__ xc, const double* __restrict__ coef,
          Level F, Level C, int nh)
{
    int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    int k = 1 + nh + blockIdx.z * blockDim.z + threadIdx.z;
    // source loops run over 2x2x2 fine blocks; extents are even on every level that is refined
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
    xf[f] = xf[f] + coef[f] * (3 * pb + po);
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
k_norm_final(const double* __restrict__ partial, int nb, double* __restrict__ out)
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

inline int ew_blocks(long long n)
{
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 4736 ? 4736 : b));
}

#define LAUNCH_OK(mg) NY_CHECK_LAUNCH((mg)->ctx)

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

#define TRY(call) do { int _r = (call); if (_r != NY_OK) return _r; } while (0)

// mod_halo.f90:235-262 exchange_with_myself, statement by statement
int fill(ny_mg* mg, cudaStream_t st, const Level& L, double* a)
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
    if (mg->zper) {
        bool ov = (nz - 2 * nh) < nh;
        TRY(assign_box(mg, st, L, a, 1 - nh, 1 - nh, 1, 1 - nh, 1 - nh, nz - 2 * nh + 1, nx + 2 * nh, ny + 2 * nh, nh, ov));
        TRY(assign_box(mg, st, L, a, 1 - nh, 1 - nh, nz - nh + 1, 1 - nh, 1 - nh, nh + 1, nx + 2 * nh, ny + 2 * nh, nh, ov));
    }
    return NY_OK;
}

inline dim3 box_grid(int ni, int nj, int nk, dim3 b)
{
    return dim3((ni + b.x - 1) / b.x, (nj + b.y - 1) / b.y, (nk + b.z - 1) / b.z);
}

int smooth(ny_mg* mg, cudaStream_t st, int lev)
{
    Level& L = mg->lev[lev - 1];
    const int nh = mg->nh;
    const double omega = mg->omega, cff1 = 1.0 - omega;
    ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_SMOOTH_FINE : NY_PROF_MG_COARSE, st);
    dim3 b(32, 4, 2);
    k_sweep<<<box_grid(L.nx + 2, L.ny + 2, L.nz - 2 * nh + 2, b), b, 0, st>>>(
        L.x, L.y, L.b, L.idiag, omega, cff1, L, nh, 0, L.nx + 1, 0, L.ny + 1, nh, L.nz + 1 - nh);
    LAUNCH_OK(mg);
    k_sweep<<<box_grid(L.nx, L.ny, L.nz - 2 * nh, b), b, 0, st>>>(
        L.y, L.x, L.b, L.idiag, omega, cff1, L, nh, 1, L.nx, 1, L.ny, nh + 1, L.nz - nh);
    LAUNCH_OK(mg);
    return fill(mg, st, L, L.x);
}

int residual(ny_mg* mg, cudaStream_t st, int lev)
{
    Level& L = mg->lev[lev - 1];
    ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_RESIDUAL_FINE : NY_PROF_MG_COARSE, st);
    dim3 b(32, 4, 2);
    k_residual<<<box_grid(L.nx, L.ny, L.nz - 2 * mg->nh, b), b, 0, st>>>(L.x, L.b, L.r, L.msk, L.diag, L, mg->nh);
    LAUNCH_OK(mg);
    return fill(mg, st, L, L.r);
}

int restriction(ny_mg* mg, cudaStream_t st, int lev, bool from_b)
{
    Level &F = mg->lev[lev - 1], &C = mg->lev[lev];
    ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_RESTRICT_FINE : NY_PROF_MG_COARSE, st);
    dim3 b(32, 4, 2);
    k_restrict<<<box_grid(C.nx, C.ny, C.nz - 2 * mg->nh, b), b, 0, st>>>(from_b ? F.b : F.r, C.b, C.Rcoef, F, C, mg->nh);
    LAUNCH_OK(mg);
    NY_CUDA(cudaMemsetAsync(C.x, 0, C.n * sizeof(double), st));           // operators.f90:209
    return fill(mg, st, C, C.b);
}

int prolongation(ny_mg* mg, cudaStream_t st, int lev)
{
    Level &F = mg->lev[lev - 1], &C = mg->lev[lev];
    ny_prof_scope ps(mg->ctx, lev == 1 ? NY_PROF_MG_PROLONG_FINE : NY_PROF_MG_COARSE, st);
    dim3 b(32, 4, 2);
    k_prolong<<<box_grid(F.nx, F.ny, F.nz - 2 * mg->nh, b), b, 0, st>>>(F.x, C.x, F.Pcoef, F, C, mg->nh);
    LAUNCH_OK(mg);
    return fill(mg, st, F, F.x);
}

// enqueue sum(msk*v^2) of level 1 into d_red[slot]
int norm_async(ny_mg* mg, cudaStream_t st, const double* v, int slot)
{
    Level& L = mg->lev[0];
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_NORM, st);
    k_norm_partial<<<NORM_BLOCKS, 256, 0, st>>>(L.msk, v, L, mg->nh, mg->d_red + 2);
    LAUNCH_OK(mg);
    k_norm_final<<<1, 256, 0, st>>>(mg->d_red + 2, NORM_BLOCKS, mg->d_red + slot);
    LAUNCH_OK(mg);
    return NY_OK;
}

int vcycle(ny_mg* mg, cudaStream_t st)
{
    const int lev1 = mg->nlevels - 1;
    for (int lev = 1; lev <= lev1; lev++) {
        TRY(smooth(mg, st, lev));
        TRY(residual(mg, st, lev));
        TRY(restriction(mg, st, lev, false));
    }
    TRY(smooth(mg, st, lev1 + 1));
    for (int lev = lev1; lev >= 1; lev--) {
        TRY(prolongation(mg, st, lev));
        TRY(smooth(mg, st, lev));
    }
    return NY_OK;
}

int setup_operators(ny_mg* mg, cudaStream_t st)
{
    const int nl = mg->nlevels;
    // setup_fine_msk (mg_setup.f90:213-223): halo-fill the finest mask
    TRY(fill(mg, st, mg->lev[0], mg->lev[0].msk));
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

double* var_ptr(ny_mg* mg, int lev, int ivar)
{
    Level& L = mg->lev[lev - 1];
    switch (ivar) {
    case NY_MG_X: return L.x; case NY_MG_B: return L.b; case NY_MG_R: return L.r; case NY_MG_Y: return L.y;
    case NY_MG_DIAG: return L.diag; case NY_MG_IDIAG: return L.idiag; case NY_MG_MSK: return L.msk;
    case NY_MG_RCOEF: return L.Rcoef; case NY_MG_PCOEF: return L.Pcoef; default: return nullptr;
    }
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
    delete mg;
}

extern "C" int ny_mg_create(ny_ctx* ctx, int nx, int ny, int nz, int topology, ny_mg** out)
{
    NY_REQUIRE(ctx && out, "null argument");
    NY_REQUIRE(nx >= 2 && ny >= 2 && nz >= 1, "grid too small");
    NY_REQUIRE(topology >= NY_TOPO_CLOSED && topology <= NY_TOPO_XYZPERIO, "unknown topology");
    ny_mg* mg = new ny_mg();
    memset(mg, 0, sizeof(ny_mg));
    mg->ctx = ctx;
    mg->nh = 3; mg->maxite = 20; mg->tol = 1e-6; mg->omega = 0.9;          // mg_types.f90:15-26
    mg->topology = topology;
    mg->xper = topology == NY_TOPO_XPERIO || topology == NY_TOPO_XYPERIO || topology == NY_TOPO_XYZPERIO;
    mg->yper = topology == NY_TOPO_YPERIO || topology == NY_TOPO_XYPERIO || topology == NY_TOPO_XYZPERIO;
    mg->zper = topology == NY_TOPO_ZPERIO || topology == NY_TOPO_XYZPERIO;
    const int nh = mg->nh;
    // create_hierarchy, mg_setup.f90:225-307 (one process: never glued).  The Fortran loops for
    // ever (and overruns its level table) when neither nx nor ny passes through 2; refuse instead.
    int x = nx, y = ny, z = nz + 2 * nh, i = 0;
    mg->lev[0].nx = x; mg->lev[0].ny = y; mg->lev[0].nz = z;
    while (!(x == 2 || y == 2)) {
        if ((x & 1) || (y & 1) || x < 2 || y < 2 || ((z - 2 * nh) & 1) || (z - 2 * nh) < 2 || i + 1 >= MAXLEV) {
            ny_set_error("ny_mg_create: %dx%dx%d cannot be halved down to nx==2 or ny==2 with even extents "
                         "(core/mgfor/mg_setup.f90:262-303)", nx, ny, nz);
            delete mg;
            return NY_ERR_GRID;
        }
        x /= 2; y /= 2; z = z / 2 + nh;
        i++;
        mg->lev[i].nx = x; mg->lev[i].ny = y; mg->lev[i].nz = z;
    }
    mg->nlevels = i + 1;
    size_t tmp_need = 1;
    for (int l = 0; l < mg->nlevels; l++) {
        Level& L = mg->lev[l];
        L.sj = L.nx + 2 * nh;
        L.sk = L.sj * (L.ny + 2 * nh);
        L.n = (size_t)L.sk * (size_t)L.nz;
        double** ptrs[] = {&L.x, &L.b, &L.r, &L.y, &L.diag, &L.idiag, &L.Rcoef, &L.Pcoef, &L.msk};
        for (double** p : ptrs) {
            cudaError_t e = cudaMalloc(p, L.n * sizeof(double));
            if (e != cudaSuccess) {
                ny_set_error("ny_mg_create: cudaMalloc of level %d failed: %s", l + 1, cudaGetErrorString(e));
                ny_mg_destroy(mg);
                return NY_ERR_CUDA;
            }
            cudaMemset(*p, 0, L.n * sizeof(double));             // Fortran leaves these uninitialised
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
    mg->tmp_doubles = tmp_need;
    if (cudaMalloc(&mg->tmp, tmp_need * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&mg->d_red, (NORM_BLOCKS + 2) * sizeof(double)) != cudaSuccess) {
        ny_set_error("ny_mg_create: cudaMalloc failed");
        ny_mg_destroy(mg);
        return NY_ERR_CUDA;
    }
    cudaStream_t st = 0;
    for (int l = 0; l < mg->nlevels; l++) {
        Level& L = mg->lev[l];
        k_default_msk<<<(unsigned)((L.n + 255) / 256), 256, 0, st>>>(L.msk, L, nh, mg->zper);
        ctx->launches++;
    }
    int r = setup_operators(mg, st);
    if (r == NY_OK && cudaStreamSynchronize(st) != cudaSuccess) { ny_set_error("ny_mg_create: setup failed"); r = NY_ERR_CUDA; }
    if (r != NY_OK) { ny_mg_destroy(mg); return r; }
    *out = mg;
    return NY_OK;
}

extern "C" int ny_mg_nlevels(ny_mg* mg) { return mg ? mg->nlevels : 0; }

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
    NY_CUDA(cudaMemcpyAsync(var_ptr(mg, lev, ivar), src, mg->lev[lev - 1].n * sizeof(double),
                            cudaMemcpyDeviceToDevice, ny_stream(stream)));
    return NY_OK;
}

extern "C" int ny_mg_get_array(ny_mg* mg, int lev, int ivar, double* dst, void* stream)
{
    NY_REQUIRE(mg && dst && lev >= 1 && lev <= mg->nlevels && var_ptr(mg, lev, ivar), "bad argument");
    NY_CUDA(cudaMemcpyAsync(dst, var_ptr(mg, lev, ivar), mg->lev[lev - 1].n * sizeof(double),
                            cudaMemcpyDeviceToDevice, ny_stream(stream)));
    return NY_OK;
}

static int read_scalars(ny_mg* mg, cudaStream_t st, int n)
{
    NY_CUDA(cudaMemcpyAsync(mg->ctx->h_pinned, mg->d_red, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    NY_CUDA(cudaStreamSynchronize(st));
    return NY_OK;
}

// solvers.f90:8-33
extern "C" int ny_mg_solve(ny_mg* mg, ny_mg_stats* stats, void* stream)
{
    NY_REQUIRE(mg, "null argument");
    cudaStream_t st = ny_stream(stream);
    Level& L = mg->lev[0];
    int nite = 0, nres = 0;
    double hist[32];
    // normb = sum(msk b^2); res = sum(msk r^2)/normb after residual(1)   (operators.f90:81-125)
    TRY(norm_async(mg, st, L.b, 0));
    TRY(read_scalars(mg, st, 1));
    const double normb = mg->ctx->h_pinned[0];
    double res = 0.0;
    auto normresidual = [&]() -> int {
        if (normb > 0.0) {
            TRY(residual(mg, st, 1));
            TRY(norm_async(mg, st, L.r, 1));
            TRY(read_scalars(mg, st, 2));
            res = mg->ctx->h_pinned[1] / normb;
        } else res = 0.0;
        return NY_OK;
    };
    TRY(normresidual());
    hist[nres++] = res;
    for (;;) {
        if (res < mg->tol) break;
        TRY(vcycle(mg, st));
        nite++;
        if (nite >= mg->maxite) break;
        TRY(normresidual());
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
    TRY(ny_mg_solve(mg, stats, stream));
    ny_prof_scope ps(mg->ctx, NY_PROF_MG_EMBED, st);
    k_extract<<<g.grid, g.block, 0, st>>>(L.x, p, L, e.nz, e.ny, e.nx, lo[0], lo[1], lo[2], scale);
    LAUNCH_OK(mg);
    return NY_OK;
}

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
    default: ny_set_error("ny_mg_op: unknown op %d", op); return NY_ERR_ARG;
    }
}
