/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement of the Nyles geometric multigrid `mgfor` for the only
 * configuration Nyles uses (core/mgfordriver.py:10-24): 3-D, cell centres,
 * one process (npx = npy = 1), topology closed / perio_xy / perio_xyz (the other
 * enum values of core/mgfor/mg_enums.f90:5-7 are handled too).
 *
 * PARITY UNPINNED: the reference stores no expected residuals and mgfor cannot
 * be compiled here (no Fortran compiler / MPI).  mgfor is built with -Ofast
 * (core/build.py:167-174), so the reference itself is defined only up to
 * re-association; this file fixes the source order and no FMA.
 *
 * Storage: every level array is (1-nh:nx+nh, 1-nh:ny+nh, 1:nz), i fastest, where
 * nz already contains the 2*nh vertical halo rows (mg_setup.f90:15-27,239).
 * That is a C-contiguous (nz, ny+2nh, nx+2nh) buffer, the canonical (k,j,i) order.
 *
 * Memory that the Fortran leaves uninitialised after `allocate` is zero here
 * (hard part 4).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAXLEV 50
enum { T_CLOSED = 1, T_XPERIO, T_YPERIO, T_ZPERIO, T_XYPERIO, T_XYZPERIO };   /* mg_enums.f90:5-7 */
enum { V_X = 1, V_B, V_R, V_Y, V_DIAG, V_IDIAG, V_MSK, V_RCOEF, V_PCOEF };    /* pytools.f90:17-62 */

typedef struct {
    int nx, ny, nz;              /* nz includes the 2*nh halo planes */
    size_t n;
    double *x, *b, *r, *y, *diag, *idiag, *Rcoef, *Pcoef;
    int *msk;
} orc_level;

typedef struct {
    int nlevels, nh, topology, maxite;
    int xper, yper, zper;
    double tol, omega;
    orc_level lev[ORC_MAXLEV];
    /* stats of the last solve */
    int nite;
    double res, normb;
    double reshist[64];
    int nres;
} orc_mg;

static inline size_t IX(const orc_level *L, int nh, int i, int j, int k)
{   /* Fortran indices: i in 1-nh..nx+nh, j in 1-nh..ny+nh, k in 1..nz */
    return ((size_t)(k - 1) * (size_t)(L->ny + 2 * nh) + (size_t)(j - 1 + nh)) * (size_t)(L->nx + 2 * nh)
           + (size_t)(i - 1 + nh);
}

/* Fortran array-section assignment dst = src: the right-hand side is evaluated
 * completely before anything is stored, which matters when the two sections
 * overlap (coarse levels with nx, ny or the interior nz smaller than nh). */
static void assign_box(const orc_level *L, int nh, double *a,
                       int di0, int dj0, int dk0, int si0, int sj0, int sk0,
                       int ni, int nj, int nk)
{
    double *tmp = malloc(sizeof(double) * (size_t)ni * (size_t)nj * (size_t)nk);
    size_t t = 0;
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++)
        tmp[t++] = a[IX(L, nh, si0 + i, sj0 + j, sk0 + k)];
    t = 0;
    for (int k = 0; k < nk; k++) for (int j = 0; j < nj; j++) for (int i = 0; i < ni; i++)
        a[IX(L, nh, di0 + i, dj0 + j, dk0 + k)] = tmp[t++];
    free(tmp);
}

/* ---- mod_halo.f90:200-262: with one rank, fill == exchange_with_myself ---- */
static void fill(const orc_mg *mg, const orc_level *L, double *a)
{
    const int nh = mg->nh, nx = L->nx, ny = L->ny, nz = L->nz;
    if (mg->xper) {
        assign_box(L, nh, a, nx + 1, 1, 1, 1, 1, 1, nh, ny, nz);                 /* :246 */
        assign_box(L, nh, a, 1 - nh, 1, 1, nx - nh + 1, 1, 1, nh, ny, nz);       /* :247 */
    }
    if (mg->yper) {
        assign_box(L, nh, a, 1, ny + 1, 1, 1, 1, 1, nx, nh, nz);                 /* :249 */
        assign_box(L, nh, a, 1, 1 - nh, 1, 1, ny - nh + 1, 1, nx, nh, nz);       /* :250 */
    }
    if (mg->xper && mg->yper) {                                                  /* corners :252-255 */
        assign_box(L, nh, a, nx + 1, ny + 1, 1, 1, 1, 1, nh, nh, nz);
        assign_box(L, nh, a, 1 - nh, ny + 1, 1, nx - nh + 1, 1, 1, nh, nh, nz);
        assign_box(L, nh, a, nx + 1, 1 - nh, 1, 1, ny - nh + 1, 1, nh, nh, nz);
        assign_box(L, nh, a, 1 - nh, 1 - nh, 1, nx - nh + 1, ny - nh + 1, 1, nh, nh, nz);
    }
    if (mg->zper) {                                                              /* :257-260 */
        assign_box(L, nh, a, 1 - nh, 1 - nh, 1, 1 - nh, 1 - nh, nz - 2 * nh + 1, nx + 2 * nh, ny + 2 * nh, nh);
        assign_box(L, nh, a, 1 - nh, 1 - nh, nz - nh + 1, 1 - nh, 1 - nh, nh + 1, nx + 2 * nh, ny + 2 * nh, nh);
    }
}

/* ---- basicoperators.f90:363-400 fsmoother3d -------------------------------- */
static void fsmoother3d(const orc_mg *mg, const orc_level *L)
{
    const int nh = mg->nh, nx = L->nx, ny = L->ny, nz = L->nz;
    const double omega = mg->omega, cff1 = 1.0 - omega;
    double *x = L->x, *y = L->y; const double *b = L->b, *idiag = L->idiag;
    const ptrdiff_t si = 1, sj = nx + 2 * nh, sk = (ptrdiff_t)sj * (ny + 2 * nh);
#pragma omp parallel for collapse(2)
    for (int k = nh; k <= nz + 1 - nh; k++)
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) {
                size_t c = IX(L, nh, i, j, k);
                y[c] = cff1 * x[c] + omega * ((x[c - si] + x[c + si] + x[c - sj] + x[c + sj]
                                               + x[c - sk] + x[c + sk]) - b[c]) * idiag[c];
            }
#pragma omp parallel for collapse(2)
    for (int k = nh + 1; k <= nz - nh; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                size_t c = IX(L, nh, i, j, k);
                x[c] = cff1 * y[c] + omega * ((y[c - si] + y[c + si] + y[c - sj] + y[c + sj]
                                               + y[c - sk] + y[c + sk]) - b[c]) * idiag[c];
            }
}

/* ---- basicoperators.f90:300-323 fresidual3d --------------------------------- */
static void fresidual3d(const orc_mg *mg, const orc_level *L)
{
    const int nh = mg->nh, nx = L->nx, ny = L->ny, nz = L->nz;
    const double *x = L->x, *b = L->b, *diag = L->diag; double *r = L->r; const int *msk = L->msk;
    const ptrdiff_t si = 1, sj = nx + 2 * nh, sk = (ptrdiff_t)sj * (ny + 2 * nh);
#pragma omp parallel for collapse(2)
    for (int k = 1 + nh; k <= nz - nh; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                size_t c = IX(L, nh, i, j, k);
                r[c] = msk[c] * (b[c] + diag[c] * x[c]
                                 - (x[c - si] + x[c + si] + x[c - sj] + x[c + sj] + x[c - sk] + x[c + sk]));
            }
}

/* ---- basicoperators.f90:32-60 frestrict_centers3d --------------------------- */
static void frestrict3d(const orc_mg *mg, const orc_level *F, const orc_level *C,
                        const double *xf, double *xc, const double *coef)
{
    const int nh = mg->nh;
#pragma omp parallel for collapse(2)
    for (int kc = 1 + nh; kc <= C->nz - nh; kc++)
        for (int jc = 1; jc <= C->ny; jc++) {
            int k = 1 + nh + (kc - 1 - nh) * 2, j = 1 + (jc - 1) * 2;
            for (int ic = 1; ic <= C->nx; ic++) {
                int i = 1 + (ic - 1) * 2;
                xc[IX(C, nh, ic, jc, kc)] = coef[IX(C, nh, ic, jc, kc)] * (
                    xf[IX(F, nh, i, j, k)] + xf[IX(F, nh, i + 1, j, k)]
                    + xf[IX(F, nh, i, j + 1, k)] + xf[IX(F, nh, i + 1, j + 1, k)]
                    + xf[IX(F, nh, i, j, k + 1)] + xf[IX(F, nh, i + 1, j, k + 1)]
                    + xf[IX(F, nh, i, j + 1, k + 1)] + xf[IX(F, nh, i + 1, j + 1, k + 1)]);
            }
        }
}

/* ---- basicoperators.f90:173-231 fprolongation_centers3d --------------------- */
static void fprolong3d(const orc_mg *mg, const orc_level *F, const orc_level *C,
                       double *xf, const double *xc, const double *coef)
{
    const int nh = mg->nh;
#define XC(I, J, K) xc[IX(C, nh, (I), (J), (K))]
#define PLANE(di, dj, K) (9 * XC(ic, jc, K) + 3 * XC(ic + (di), jc, K) + 3 * XC(ic, jc + (dj), K) + XC(ic + (di), jc + (dj), K))
#pragma omp parallel for collapse(2)
    for (int k = 1 + nh; k <= F->nz - nh; k += 2)
        for (int j = 1; j <= F->ny - 1; j += 2) {
            int kc = nh + (k + 1 - nh) / 2, jc = (j + 1) / 2;
            for (int i = 1; i <= F->nx - 1; i += 2) {
                int ic = (i + 1) / 2;
                for (int dj = 0; dj <= 1; dj++)
                    for (int di = 0; di <= 1; di++) {
                        int sdi = di ? 1 : -1, sdj = dj ? 1 : -1;
                        double a = PLANE(sdi, sdj, kc - 1);
                        double b = PLANE(sdi, sdj, kc);
                        double c = PLANE(sdi, sdj, kc + 1);
                        size_t f0 = IX(F, nh, i + di, j + dj, k), f1 = IX(F, nh, i + di, j + dj, k + 1);
                        xf[f0] = xf[f0] + coef[f0] * (3 * b + a);
                        xf[f1] = xf[f1] + coef[f1] * (3 * b + c);
                    }
            }
        }
#undef PLANE
#undef XC
}

/* ---- basicoperators.f90:422-440 fnorm3d -------------------------------------- */
static double fnorm3d(const orc_mg *mg, const orc_level *L, const double *x)
{
    const int nh = mg->nh;
    double sum = 0.0;
#pragma omp parallel for collapse(2) reduction(+ : sum)
    for (int k = 1 + nh; k <= L->nz - nh; k++)
        for (int j = 1; j <= L->ny; j++)
            for (int i = 1; i <= L->nx; i++) {
                size_t c = IX(L, nh, i, j, k);
                sum = sum + L->msk[c] * (x[c] * x[c]);
            }
    return sum;
}

/* ---- operators.f90:127-244 level operations (each ends with a halo fill) ---- */
static void residual(orc_mg *mg, int lev)                  /* lev is 1-based as in the source */
{
    orc_level *L = &mg->lev[lev - 1];
    fresidual3d(mg, L);
    fill(mg, L, L->r);
}
static void smooth(orc_mg *mg, int lev)                    /* nite is ignored by the source (:151-170) */
{
    orc_level *L = &mg->lev[lev - 1];
    fsmoother3d(mg, L);
    fill(mg, L, L->x);
}
static void restriction(orc_mg *mg, int lev, int from_b)   /* which = rb (r->b) or bb (b->b) */
{
    orc_level *F = &mg->lev[lev - 1], *C = &mg->lev[lev];
    frestrict3d(mg, F, C, from_b ? F->b : F->r, C->b, C->Rcoef);
    memset(C->x, 0, C->n * sizeof(double));                /* :209 */
    fill(mg, C, C->b);
}
static void prolongation(orc_mg *mg, int lev)
{
    orc_level *F = &mg->lev[lev - 1], *C = &mg->lev[lev];
    fprolong3d(mg, F, C, F->x, C->x, F->Pcoef);
    fill(mg, F, F->x);
}

/* ---- mg_setup.f90:225-307 create_hierarchy (npx*npy == 1: never glued) ------ */
static int create_hierarchy(orc_mg *mg, int nx, int ny, int nz)
{
    int x = nx, y = ny, z = nz + 2 * mg->nh, i = 0;
    mg->lev[0].nx = x; mg->lev[0].ny = y; mg->lev[0].nz = z;
    for (;;) {
        if (x == 2 || y == 2) break;
        x /= 2; y /= 2; z = z / 2 + mg->nh;
        i++;
        if (i >= ORC_MAXLEV) return -1;
        mg->lev[i].nx = x; mg->lev[i].ny = y; mg->lev[i].nz = z;
    }
    mg->nlevels = i + 1;
    return 0;
}

static void *zalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

/* ---- operators.f90:246-297 apply_default_msk (one rank) ---------------------- */
static void apply_default_msk(orc_mg *mg, int lev)
{
    orc_level *L = &mg->lev[lev - 1];
    const int nh = mg->nh, nx = L->nx, ny = L->ny, nz = L->nz;
    for (int k = 1; k <= nz; k++)
        for (int j = 1 - nh; j <= ny + nh; j++)
            for (int i = 1 - nh; i <= nx + nh; i++) {
                int keep = 1;
                if (!mg->xper && (i <= 0 || i >= nx + 1)) keep = 0;   /* west/east == -1 */
                if (!mg->yper && (j <= 0 || j >= ny + 1)) keep = 0;   /* south/north == -1 */
                if (!keep) L->msk[IX(L, nh, i, j, k)] = 0;
            }
    memset(L->y, 0, L->n * sizeof(double));                            /* :296 */
}

static void set_all(double *a, size_t n, double v) { for (size_t i = 0; i < n; i++) a[i] = v; }

/* ---- operators.f90:299-335 compute_msk --------------------------------------- */
static void compute_msk(orc_mg *mg, int lev)
{
    orc_level *F = &mg->lev[lev - 1], *C = &mg->lev[lev];
    for (size_t i = 0; i < C->n; i++) C->msk[i] = 1;
    set_all(C->Rcoef, C->n, 1.0);
    for (size_t i = 0; i < F->n; i++) F->b[i] = F->msk[i];
    restriction(mg, lev, 1);
    for (size_t i = 0; i < C->n; i++) C->msk[i] = (C->b[i] > 0.0) ? 1 : 0;   /* threshold 0 for centres */
    apply_default_msk(mg, lev + 1);
}

/* ---- operators.f90:337-393 compute_Rcoef -------------------------------------- */
static void compute_Rcoef(orc_mg *mg, int lev)
{
    orc_level *F = &mg->lev[lev - 1], *C = &mg->lev[lev];
    for (size_t i = 0; i < C->n; i++) C->Rcoef[i] = C->msk[i];
    set_all(F->b, F->n, 1.0);
    restriction(mg, lev, 1);
    set_all(C->Rcoef, C->n, 0.0);
    for (size_t i = 0; i < C->n; i++) if (C->b[i] > 0.0) C->y[i] = 4.0 / C->b[i];
    for (size_t i = 0; i < C->n; i++) C->y[i] = C->msk[i] * C->y[i];
    memcpy(C->Rcoef, C->y, C->n * sizeof(double));
}

/* ---- operators.f90:395-424 compute_Pcoef -------------------------------------- */
static void compute_Pcoef(orc_mg *mg, int lev)
{
    orc_level *F = &mg->lev[lev - 1], *C = &mg->lev[lev];
    set_all(F->x, F->n, 0.0);
    for (size_t i = 0; i < C->n; i++) C->x[i] = C->msk[i];
    for (size_t i = 0; i < F->n; i++) F->Pcoef[i] = F->msk[i];
    prolongation(mg, lev);
    set_all(F->Pcoef, F->n, 0.0);
    for (size_t i = 0; i < F->n; i++) if (F->x[i] > 0.0) F->Pcoef[i] = 1.0 / F->x[i];
    for (size_t i = 0; i < F->n; i++) F->Pcoef[i] = F->msk[i] * F->Pcoef[i];
}

/* ---- operators.f90:426-457 compute_diag ---------------------------------------- */
static void compute_diag(orc_mg *mg, int lev)
{
    orc_level *L = &mg->lev[lev - 1];
    for (size_t i = 0; i < L->n; i++) L->x[i] = L->msk[i];
    set_all(L->b, L->n, 0.0);
    set_all(L->diag, L->n, 0.0);
    residual(mg, lev);
    for (size_t i = 0; i < L->n; i++) L->diag[i] = -L->r[i];
    set_all(L->idiag, L->n, 0.0);
    for (size_t i = 0; i < L->n; i++) if (L->diag[i] > 0.0) L->idiag[i] = 1.0 / L->diag[i];
    set_all(L->x, L->n, 0.0);
}

/* ---- operators.f90:461-505 setup_operators (3-D centres branch) ---------------- */
void orc_mg_setup_operators(orc_mg *mg)
{
    int lev;
    for (lev = 1; lev <= mg->nlevels - 1; lev++) compute_msk(mg, lev);
    for (lev = 1; lev <= mg->nlevels - 1; lev++) { compute_Rcoef(mg, lev); compute_Pcoef(mg, lev); }
    for (lev = 1; lev <= mg->nlevels; lev++) compute_diag(mg, lev);
}

/* ---- mg_setup.f90:213-223 setup_fine_msk: halo-fill the finest mask through y ---- */
void orc_mg_setup_fine_msk(orc_mg *mg)
{
    orc_level *L = &mg->lev[0];
    for (size_t i = 0; i < L->n; i++) L->y[i] = L->msk[i];
    fill(mg, L, L->y);
    for (size_t i = 0; i < L->n; i++) L->msk[i] = (int)L->y[i];
}

void orc_mg_free(orc_mg *mg)
{
    if (!mg) return;
    for (int l = 0; l < mg->nlevels; l++) {
        orc_level *L = &mg->lev[l];
        free(L->x); free(L->b); free(L->r); free(L->y); free(L->diag); free(L->idiag);
        free(L->Rcoef); free(L->Pcoef); free(L->msk);
    }
    free(mg);
}

/* ---- mg_setup.f90:318-437 get_ptrmg (vertices=F, short=F, is3d=T) --------------- */
orc_mg *orc_mg_create(int nx, int ny, int nz, int topology)
{
    orc_mg *mg = calloc(1, sizeof(orc_mg));
    if (!mg) return NULL;
    mg->nh = 3; mg->maxite = 20; mg->tol = 1e-6; mg->omega = 0.9;      /* mg_types.f90:15-26 */
    mg->topology = topology;
    mg->xper = (topology == T_XPERIO || topology == T_XYPERIO || topology == T_XYZPERIO);
    mg->yper = (topology == T_YPERIO || topology == T_XYPERIO || topology == T_XYZPERIO);
    mg->zper = (topology == T_ZPERIO || topology == T_XYZPERIO);
    if (create_hierarchy(mg, nx, ny, nz)) { free(mg); return NULL; }
    const int nh = mg->nh;
    for (int l = 0; l < mg->nlevels; l++) {                            /* allocate_mg :117-178 */
        orc_level *L = &mg->lev[l];
        L->n = (size_t)(L->nx + 2 * nh) * (size_t)(L->ny + 2 * nh) * (size_t)L->nz;
        L->x = zalloc(L->n, sizeof(double)); L->b = zalloc(L->n, sizeof(double));
        L->r = zalloc(L->n, sizeof(double)); L->y = zalloc(L->n, sizeof(double));
        L->diag = zalloc(L->n, sizeof(double)); L->idiag = zalloc(L->n, sizeof(double));
        L->Rcoef = zalloc(L->n, sizeof(double)); L->Pcoef = zalloc(L->n, sizeof(double));
        L->msk = zalloc(L->n, sizeof(int));
        /* set_default_msk :180-211 */
        int k0 = mg->zper ? 1 : 1 + nh, k1 = mg->zper ? L->nz : L->nz - nh;
        for (int k = k0; k <= k1; k++)
            for (int j = 1; j <= L->ny; j++)
                for (int i = 1; i <= L->nx; i++) L->msk[IX(L, nh, i, j, k)] = 1;
    }
    orc_mg_setup_fine_msk(mg);
    orc_mg_setup_operators(mg);
    return mg;
}

int orc_mg_nlevels(const orc_mg *mg) { return mg->nlevels; }
void orc_mg_shape(const orc_mg *mg, int lev, int *shape)   /* pytools.f90:8-15, numpy order */
{
    const orc_level *L = &mg->lev[lev - 1];
    shape[0] = L->nz; shape[1] = L->ny + 2 * mg->nh; shape[2] = L->nx + 2 * mg->nh;
}
void orc_mg_set_param(orc_mg *mg, int maxite, double tol, double omega)
{
    mg->maxite = maxite; mg->tol = tol; mg->omega = omega;
}

static double *var_ptr(orc_mg *mg, int lev, int ivar)
{
    orc_level *L = &mg->lev[lev - 1];
    switch (ivar) {
    case V_X: return L->x; case V_B: return L->b; case V_R: return L->r; case V_Y: return L->y;
    case V_DIAG: return L->diag; case V_IDIAG: return L->idiag; case V_RCOEF: return L->Rcoef;
    case V_PCOEF: return L->Pcoef; default: return NULL;
    }
}
/* pytools.f90:17-62 */
int orc_mg_get_array(orc_mg *mg, int lev, int ivar, double *out)
{
    orc_level *L = &mg->lev[lev - 1];
    if (ivar == V_MSK) { for (size_t i = 0; i < L->n; i++) out[i] = L->msk[i]; return 0; }
    double *p = var_ptr(mg, lev, ivar);
    if (!p) return -1;
    memcpy(out, p, L->n * sizeof(double));
    return 0;
}
int orc_mg_set_array(orc_mg *mg, int lev, int ivar, const double *in)
{
    orc_level *L = &mg->lev[lev - 1];
    if (ivar == V_MSK) { for (size_t i = 0; i < L->n; i++) L->msk[i] = (int)in[i]; return 0; }
    if (ivar < V_X || ivar > V_Y) return -1;               /* only x,b,r,y,msk are settable */
    memcpy(var_ptr(mg, lev, ivar), in, L->n * sizeof(double));
    return 0;
}

/* ---- operators.f90:98-125 norm, :81-96 normresidual (one rank: no allreduce) ---- */
double orc_mg_norm(orc_mg *mg, int lev, int which_b)
{
    orc_level *L = &mg->lev[lev - 1];
    return fnorm3d(mg, L, which_b ? L->b : L->r);
}
static double normresidual(orc_mg *mg, double normb)
{
    if (normb > 0.0) { residual(mg, 1); return orc_mg_norm(mg, 1, 0) / normb; }
    return 0.0;
}

/* ---- solvers.f90:35-55 vcycle ----------------------------------------------------- */
void orc_mg_vcycle(orc_mg *mg)
{
    int lev, lev1 = mg->nlevels - 1;
    for (lev = 1; lev <= lev1; lev++) { smooth(mg, lev); residual(mg, lev); restriction(mg, lev, 0); }
    smooth(mg, lev1 + 1);
    for (lev = lev1; lev >= 1; lev--) { prolongation(mg, lev); smooth(mg, lev); }
}

/* ---- solvers.f90:8-33 solve -------------------------------------------------------- */
void orc_mg_solve(orc_mg *mg)
{
    int nite = 0;
    double normb = orc_mg_norm(mg, 1, 1);
    double res = normresidual(mg, normb);
    mg->nres = 0;
    mg->reshist[mg->nres++] = res;
    for (;;) {
        if (res < mg->tol) break;
        orc_mg_vcycle(mg);
        nite++;
        if (nite >= mg->maxite) break;
        res = normresidual(mg, normb);
        if (mg->nres < 64) mg->reshist[mg->nres++] = res;
    }
    mg->nite = nite; mg->res = res; mg->normb = normb;
}

void orc_mg_stats(const orc_mg *mg, int *nite, double *res, double *normb)
{
    *nite = mg->nite; *res = mg->res; *normb = mg->normb;
}
int orc_mg_reshist(const orc_mg *mg, double *out) { memcpy(out, mg->reshist, mg->nres * sizeof(double)); return mg->nres; }

/* single operations, exported for operator-level parity tests */
void orc_mg_smooth(orc_mg *mg, int lev) { smooth(mg, lev); }
void orc_mg_residual(orc_mg *mg, int lev) { residual(mg, lev); }
void orc_mg_restriction(orc_mg *mg, int lev) { restriction(mg, lev, 0); }
void orc_mg_prolongation(orc_mg *mg, int lev) { prolongation(mg, lev); }
void orc_mg_fill(orc_mg *mg, int lev, int ivar) { orc_level *L = &mg->lev[lev - 1]; fill(mg, L, var_ptr(mg, lev, ivar)); }
