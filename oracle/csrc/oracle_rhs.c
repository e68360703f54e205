/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement of the Nyles Fortran right-hand-side kernels.  It exists only
 * so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check / time the CUDA path against it.  Nothing under nyles_b200/ may call it.
 *
 * PARITY UNPINNED by the reference's own tests: the reference ships no golden
 * vectors and its Fortran cannot be built in this image (no gfortran/f2py
 * backend).  The arithmetic below follows the Fortran source statement by
 * statement; the Python orchestration around it is pinned separately by running
 * the reference's own Python drivers on top of these kernels
 * (oracle/gen_golden.py -> tests/golden/).
 *
 * Build flavours (oracle/Makefile):
 *   strict : -O2 -ffp-contract=off            bit-reproducible, mirrors gfortran -O3
 *            on baseline x86-64 (no FMA, no re-association)
 *   fast   : -O3 -march=native -fopenmp       CPU-baseline timing only
 *
 * Array convention: every routine receives f2py-style logical 3-D arrays
 * A(a,b,c), 0-based here, addressed through explicit element strides so that the
 * caller may pass permuted views of one canonical (k,j,i) buffer.  f2py makes
 * memory layout invisible to the Fortran side (intent(inplace) re-strides), so
 * only logical indexing is part of the contract.
 *
 * Quirk that matters for parity (core/Makefile:1-2: no -fdefault-real-8):
 * un-suffixed real literals in the Fortran are REAL(4) constants promoted to
 * double, and `tau5` in weno5 is an implicitly typed REAL(4) variable
 * (core/weno.f90:45).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

#define AT(p, s, a, b, c) ((p)[(ptrdiff_t)(a) * (s)[0] + (ptrdiff_t)(b) * (s)[1] + (ptrdiff_t)(c) * (s)[2]])

/* ---- core/weno.f90:1-22 ------------------------------------------------ */
static inline double weno3(double qm, double q0, double qp)
{
    const double eps = (double)1e-14f;            /* weno.f90:8, REAL(4) literal */
    double qi1 = (-qm + 3 * q0) * 0.5;            /* :10 */
    double qi2 = (q0 + qp) * 0.5;                 /* :11 */
    double d1 = q0 - qm, d2 = qp - q0;
    double beta1 = d1 * d1;                       /* :13 */
    double beta2 = d2 * d2;                       /* :14 */
    double tau = fabs(beta2 - beta1);             /* :15 (declared real*8) */
    double w1 = 1.0 + tau / (beta1 + eps);        /* :17 */
    double w2 = (1.0 + tau / (beta2 + eps)) * 2;  /* :18 */
    return (w1 * qi1 + w2 * qi2) / (w1 + w2);     /* :20 */
}

/* ---- core/weno.f90:25-54 ----------------------------------------------- */
static inline double weno5(double qmm, double qm, double q0, double qp, double qpp)
{
    /* REAL(4) constant expressions, folded in single precision (:36-38,40) */
    const double c13 = (double)(1.0f / 3.0f);
    const double c76 = (double)(7.0f / 6.0f);
    const double c116 = (double)(11.0f / 6.0f);
    const double c16 = (double)(1.0f / 6.0f);
    const double c56 = (double)(5.0f / 6.0f);
    const double k1 = (double)(13.0f / 12.0f);
    const double k2 = 0.25;
    const double eps = (double)1e-16f;            /* :34 */

    double qi1 = c13 * qmm - c76 * qm + c116 * q0;     /* :36 */
    double qi2 = -(c16 * qm) + c56 * q0 + c13 * qp;    /* :37 */
    double qi3 = c13 * q0 + c56 * qp - c16 * qpp;      /* :38 */

    double a1 = qmm - 2 * qm + q0, a2 = qmm - 4 * qm + 3 * q0;
    double b1 = qm - 2 * q0 + qp, b2 = qm - qp;
    double g1 = q0 - 2 * qp + qpp, g2 = 3 * q0 - 4 * qp + qpp;
    double beta1 = k1 * (a1 * a1) + k2 * (a2 * a2);    /* :42 */
    double beta2 = k1 * (b1 * b1) + k2 * (b2 * b2);    /* :43 */
    double beta3 = k1 * (g1 * g1) + k2 * (g2 * g2);    /* :44 */

    /* :46 tau5 is not declared and there is no `implicit none` in this
       function -> REAL(4): the difference is rounded to single precision. */
    float tau5f = (float)fabs(beta1 - beta3);
    double tau5 = (double)tau5f;

    double w1 = 1.0 + tau5 / (beta1 + eps);            /* :49 */
    double w2 = 6 * (1.0 + tau5 / (beta2 + eps));      /* :50 */
    double w3 = 3 * (1.0 + tau5 / (beta3 + eps));      /* :51 */

    return (w1 * qi1 + w2 * qi2 + w3 * qi3) / (w1 + w2 + w3);   /* :53 */
}

/* exported for unit tests */
double orc_weno3(double qm, double q0, double qp) { return weno3(qm, q0, qp); }
double orc_weno5(double a, double b, double c, double d, double e) { return weno5(a, b, c, d, e); }

/* ---- core/weno.f90:106-153  flux1d -------------------------------------
 * q,u,flux are contiguous lines of length n, 0-based here (Fortran index = c+1).
 * Face c sits between cells c and c+1.  Requires n >= 5. */
static void flux1d(const double *u, const double *q, double *flux, int n)
{
    int i;
    /* Fortran i=1 */
    flux[0] = (u[0] > 0) ? u[0] * q[0] : u[0] * weno3(q[2], q[1], q[0]);
    /* i=2 */
    flux[1] = (u[1] > 0) ? u[1] * weno3(q[0], q[1], q[2])
                         : u[1] * weno5(q[4], q[3], q[2], q[1], q[0]);
    /* i=3..n-3 */
    for (i = 2; i <= n - 4; i++) {
        if (u[i] > 0)
            flux[i] = u[i] * weno5(q[i - 2], q[i - 1], q[i], q[i + 1], q[i + 2]);
        else
            flux[i] = u[i] * weno5(q[i + 3], q[i + 2], q[i + 1], q[i], q[i - 1]);
    }
    /* i=n-2 */
    i = n - 3;
    flux[i] = (u[i] > 0) ? u[i] * weno5(q[i - 2], q[i - 1], q[i], q[i + 1], q[i + 2])
                         : u[i] * weno3(q[i + 2], q[i + 1], q[i]);
    /* i=n-1 */
    i = n - 2;
    flux[i] = (u[i] > 0) ? u[i] * weno3(q[i - 1], q[i], q[i + 1]) : u[i] * q[i + 1];
    flux[n - 1] = 0.0;
}

void orc_flux1d(const double *u, const double *q, double *flux, int n) { flux1d(u, q, flux, n); }

/* ======================================================================
 * The linear (non-WENO) upwind branch: `linear = .true.` in fortran_upwind.f90:31 and
 * fortran_vortex_force.f90:28,108.  The reference ships with linear = .false. (the flag is a local
 * variable, not an argument), so this branch is dormant there; it is restated for SURVEY.md 8(f).4.
 * Coefficients are REAL(4) constant expressions promoted to double (interpolate.f90:19-33).
 * ====================================================================== */
#define LC1 ((double)(-(1.0f / 6.0f)))
#define LC2 ((double)(5.0f / 6.0f))
#define LC3 ((double)(2.0f / 6.0f))
#define LE1 ((double)(-(1.0f / 12.0f)))
#define LE2 ((double)(7.0f / 12.0f))
#define LB1 ((double)(2.0f / 60.0f))
#define LB2 ((double)(-(13.0f / 60.0f)))
#define LB3 ((double)(47.0f / 60.0f))
#define LB4 ((double)(27.0f / 60.0f))
#define LB5 ((double)(-(3.0f / 60.0f)))
/* 1-based access into 0-based C arrays, as the Fortran writes it */
#define V1(a, i) ((a)[(i) - 1])
#define THIRD_P(v, i) (LC1 * V1(v, (i) - 1) + LC2 * V1(v, i) + LC3 * V1(v, (i) + 1))
#define THIRD_M(v, i) (LC3 * V1(v, (i) - 1) + LC2 * V1(v, i) + LC1 * V1(v, (i) + 1))
#define FIFTH_P(v, i) (LB1 * V1(v, (i) - 2) + LB2 * V1(v, (i) - 1) + LB3 * V1(v, i) + LB4 * V1(v, (i) + 1) + LB5 * V1(v, (i) + 2))
#define FIFTH_M(v, i) (LB5 * V1(v, (i) - 2) + LB4 * V1(v, (i) - 1) + LB3 * V1(v, i) + LB2 * V1(v, (i) + 1) + LB1 * V1(v, (i) + 2))

/* core/interpolate.f90:3-112, the flavour included by fortran_vortex_force.f90: qp(0:n-1), qm(1:n).
 * qp0 points to qp(0); entries the Fortran leaves unassigned keep what the caller put there. */
static void interpolate_vf(const double *vU, double *qp0, double *qm, int order, int n)
{
    int i;
#define QP(i) qp0[(i)]
#define QM(i) qm[(i) - 1]
    if (order == 5) {
        QP(0) = 0.0;
        i = 1; QP(i) = V1(vU, i); QM(i) = V1(vU, i);
        i = 2; QP(i) = THIRD_P(vU, i); QM(i) = THIRD_M(vU, i);
        for (i = 3; i <= n - 3; i++) { QP(i) = FIFTH_P(vU, i); QM(i) = FIFTH_M(vU, i); }
        i = n - 2; QP(i) = THIRD_P(vU, i); QM(i) = THIRD_M(vU, i);
        i = n - 1; QP(i) = V1(vU, i); QM(i) = V1(vU, i);
        QM(n) = V1(vU, n);
    } else if (order == 3) {
        QP(0) = 0.0;
        i = 1; QP(i) = V1(vU, i); QM(i) = V1(vU, i);
        for (i = 2; i <= n - 2; i++) { QP(i) = THIRD_P(vU, i); QM(i) = THIRD_M(vU, i); }
        i = n - 1; QP(i) = V1(vU, i); QM(i) = V1(vU, i);
        QM(n) = V1(vU, n);
    } else if (order == 1) {
        QP(0) = 0.0;
        for (i = 1; i <= n - 1; i++) { QP(i) = V1(vU, i); QM(i) = V1(vU, i); }
        QM(n) = V1(vU, n);
    } else if (order == 2) {
        QM(1) = 0.5 * V1(vU, 1);
        for (i = 2; i <= n; i++) QM(i) = 0.5 * (V1(vU, i - 1) + V1(vU, i));
    } else if (order == 4) {
        i = 1; QM(i) = LE2 * (V1(vU, i)) + LE1 * (V1(vU, i + 1));
        i = 2; QM(i) = LE2 * (V1(vU, i - 1) + V1(vU, i)) + LE1 * (V1(vU, i + 1));
        for (i = 3; i <= n - 1; i++) QM(i) = LE2 * (V1(vU, i - 1) + V1(vU, i)) + LE1 * (V1(vU, i - 2) + V1(vU, i + 1));
        i = n; QM(i) = LE2 * (V1(vU, i - 1) + V1(vU, i)) + LE1 * (V1(vU, i - 2));
    }
#undef QP
#undef QM
}

/* core/interpolate_tracer.f90:15-113, the flavour included by fortran_upwind.f90: qp(1:n), qm(1:n) */
static void interpolate_tr(const double *q, double *qp, double *qm, int order, int n)
{
    int i;
#define QP(i) qp[(i) - 1]
#define QM(i) qm[(i) - 1]
    if (order == 5) {
        i = 1; QP(i) = V1(q, i); QM(i) = V1(q, i);
        i = 2; QP(i) = THIRD_P(q, i); QM(i) = THIRD_M(q, i);
        for (i = 3; i <= n - 2; i++) { QP(i) = FIFTH_P(q, i); QM(i) = FIFTH_M(q, i); }
        i = n - 1; QP(i) = THIRD_P(q, i); QM(i) = THIRD_M(q, i);
        i = n; QP(i) = V1(q, i); QM(i) = V1(q, i);
    } else if (order == 3) {
        i = 1; QP(i) = V1(q, i); QM(i) = V1(q, i);
        for (i = 2; i <= n - 1; i++) { QP(i) = THIRD_P(q, i); QM(i) = THIRD_M(q, i); }
        i = n; QP(i) = V1(q, i); QM(i) = V1(q, i);
    } else if (order == 1) {
        for (i = 1; i <= n; i++) { QP(i) = V1(q, i); QM(i) = V1(q, i); }
    } else if (order == 2) {
        for (i = 1; i <= n - 1; i++) QP(i) = 0.5 * (V1(q, i) + V1(q, i + 1));
    } else if (order == 4) {
        i = 1; QP(i) = LE2 * (V1(q, i) + V1(q, i + 1)) + LE1 * (V1(q, i + 2));
        for (i = 2; i <= n - 2; i++) QP(i) = LE2 * (V1(q, i) + V1(q, i + 1)) + LE1 * (V1(q, i - 1) + V1(q, i + 2));
        i = n - 1; QP(i) = LE2 * (V1(q, i) + V1(q, i + 1)) + LE1 * (V1(q, i - 1));
    }
#undef QP
#undef QM
}

void orc_interpolate_vf(const double *vU, double *qp0, double *qm, int order, int n) { interpolate_vf(vU, qp0, qm, order, n); }
void orc_interpolate_tr(const double *q, double *qp, double *qm, int order, int n) { interpolate_tr(q, qp, qm, order, n); }

/* fortran_upwind.f90:33-64, the branch `if (linear)` */
void orc_upwind_linear(const double *trac, const double *u, double *dtrac, int order,
                       int l, int m, int n, const ptrdiff_t *s)
{
#pragma omp parallel
    {
        double *up = calloc(5 * (size_t)n, sizeof(double));
        double *um = up + n, *phi = up + 2 * n, *qp = up + 3 * n, *qm = up + 4 * n;
#pragma omp for collapse(2)
        for (int k = 0; k < l; k++)
            for (int j = 0; j < m; j++) {
                for (int i = 0; i < n; i++) {
                    const double ui = AT(u, s, k, j, i), UU = fabs(ui);
                    up[i] = 0.5 * (ui + UU);
                    um[i] = 0.5 * (ui - UU);
                    phi[i] = AT(trac, s, k, j, i);
                }
                interpolate_tr(phi, qp, qm, order, n);
                double fxm = 0.0, fx;
                for (int i = 0; i < n - 1; i++) {
                    if (order % 2 == 0) fx = AT(u, s, k, j, i) * qp[i];
                    else fx = up[i] * qp[i] + um[i] * qm[i + 1];
                    AT(dtrac, s, k, j, i) = AT(dtrac, s, k, j, i) + fxm - fx;
                    fxm = fx;
                }
                fx = 0.0;
                AT(dtrac, s, k, j, n - 1) = AT(dtrac, s, k, j, n - 1) + fxm - fx;
            }
        free(up);
    }
}

/* fortran_vortex_force.f90:39-64, the branch `if (linear)` of vortex_force_direc */
void orc_vortex_force_direc_linear(const double *U, const double *vort, double *res, int order,
                                   int m, int n, int l, const ptrdiff_t *s)
{
#pragma omp parallel
    {
        double *vU = calloc(5 * (size_t)l + 1, sizeof(double));
        double *up = vU + l, *um = vU + 2 * l, *qm = vU + 3 * l, *qp0 = vU + 4 * l;   /* qp(0:l-1) */
#pragma omp for collapse(2)
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n - 1; i++) {
                double UU_0 = 0.0;
                for (int k = 0; k < l; k++) {
                    const double UU_1 = 0.5 * (AT(U, s, j, i, k) + AT(U, s, j, i + 1, k));
                    vU[k] = AT(vort, s, j, i, k);
                    const double Ui = 0.5 * (UU_0 + UU_1), UU = fabs(Ui);
                    up[k] = 0.5 * (Ui + UU);
                    um[k] = 0.5 * (Ui - UU);
                    UU_0 = UU_1;
                }
                interpolate_vf(vU, qp0, qm, order, l);
                for (int k = 0; k < l; k++) {            /* Fortran k = k+1: qp(k-1) -> qp0[k], qm(k) -> qm[k] */
                    if (order % 2 == 0) AT(res, s, j, i, k) = AT(res, s, j, i, k) - qm[k];
                    else AT(res, s, j, i, k) = AT(res, s, j, i, k) - qp0[k] * up[k] - qm[k] * um[k];
                }
            }
        free(vU);
    }
}

/* fortran_vortex_force.f90:118-143, the branch `if (linear)` of vortex_force_flip */
void orc_vortex_force_flip_linear(const double *U, const double *vort, double *res, int order,
                                  int m, int n, int l, const ptrdiff_t *s)
{
#pragma omp parallel
    {
        double *vU = calloc(5 * (size_t)n + 1, sizeof(double));
        double *up = vU + n, *um = vU + 2 * n, *qm = vU + 3 * n, *qp0 = vU + 4 * n;
#pragma omp for collapse(2)
        for (int j = 0; j < m; j++)
            for (int k = 0; k < l - 1; k++) {
                double UU_0 = 0.0;
                for (int i = 0; i < n; i++) {
                    const double UU_1 = 0.5 * (AT(U, s, j, i, k) + AT(U, s, j, i, k + 1));
                    vU[i] = AT(vort, s, j, i, k);
                    const double Ui = 0.5 * (UU_0 + UU_1), UU = fabs(Ui);
                    up[i] = 0.5 * (Ui + UU);
                    um[i] = 0.5 * (Ui - UU);
                    UU_0 = UU_1;
                }
                interpolate_vf(vU, qp0, qm, order, n);
                for (int i = 0; i < n; i++) {
                    if (order % 2 == 0) AT(res, s, j, i, k) = AT(res, s, j, i, k) + qm[i];
                    else AT(res, s, j, i, k) = AT(res, s, j, i, k) + qp0[i] * up[i] + qm[i] * um[i];
                }
            }
        free(vU);
    }
}

/* ---- core/fortran_vorticity.f90:2-28 ----------------------------------- */
void orc_vorticity(const double *ui, const double *uj, double *wk,
                   int l, int m, int n, const ptrdiff_t *s)
{
#pragma omp parallel for
    for (int k = 0; k < l; k++)
        for (int j = 0; j < m - 1; j++) {
            for (int i = 0; i < n - 1; i++)
                AT(wk, s, k, j, i) = AT(uj, s, k, j, i + 1) - AT(uj, s, k, j, i)
                                     - AT(ui, s, k, j + 1, i) + AT(ui, s, k, j, i);
            AT(wk, s, k, j, n - 1) = 0.0;
        }
}

/* ---- core/fortran_upwind.f90:3-87 (WENO branch :66-82; linear=.false.) -- */
void orc_upwind(const double *trac, const double *u, double *dtrac,
                int l, int m, int n, const ptrdiff_t *s)
{
#pragma omp parallel
    {
        double *up = malloc(sizeof(double) * 3 * (size_t)n);
        double *phi = up + n, *flux = up + 2 * n;
#pragma omp for collapse(2)
        for (int k = 0; k < l; k++)
            for (int j = 0; j < m; j++) {
                for (int i = 0; i < n; i++) {
                    up[i] = AT(u, s, k, j, i);
                    phi[i] = AT(trac, s, k, j, i);
                }
                flux1d(up, phi, flux, n);
                AT(dtrac, s, k, j, 0) = AT(dtrac, s, k, j, 0) - flux[0];
                for (int i = 1; i < n; i++)
                    AT(dtrac, s, k, j, i) = AT(dtrac, s, k, j, i) + flux[i - 1] - flux[i];
            }
        free(up);
    }
}

/* ---- core/fortran_vortex_force.f90:10-85 (WENO branch :66-80) -----------
 * arrays are A(j,i,k) with extents (m,n,l); sweep axis = last. */
void orc_vortex_force_direc(const double *U, const double *vort, double *res,
                            int m, int n, int l, const ptrdiff_t *s)
{
#pragma omp parallel
    {
        double *u1d = malloc(sizeof(double) * 3 * (size_t)l);
        double *q = u1d + l, *flux = u1d + 2 * l;
#pragma omp for collapse(2)
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n - 1; i++) {
                double UU_0 = 0.0;
                for (int k = 0; k < l; k++) {
                    double UU_1 = 0.5 * (AT(U, s, j, i, k) + AT(U, s, j, i + 1, k));
                    u1d[k] = 0.5 * (UU_0 + UU_1);
                    UU_0 = UU_1;
                }
                q[0] = 0.0;
                for (int k = 1; k < l; k++) q[k] = AT(vort, s, j, i, k - 1);
                flux1d(u1d, q, flux, l);
                for (int k = 0; k < l; k++)
                    AT(res, s, j, i, k) = AT(res, s, j, i, k) - flux[k];
            }
        free(u1d);
    }
}

/* ---- core/fortran_vortex_force.f90:87-165 (WENO branch :145-159) --------
 * sweep axis = middle index i (extent n); averaging across the last index. */
void orc_vortex_force_flip(const double *U, const double *vort, double *res,
                           int m, int n, int l, const ptrdiff_t *s)
{
#pragma omp parallel
    {
        double *u1d = malloc(sizeof(double) * 3 * (size_t)n);
        double *q = u1d + n, *flux = u1d + 2 * n;
#pragma omp for collapse(2)
        for (int j = 0; j < m; j++)
            for (int k = 0; k < l - 1; k++) {
                double UU_0 = 0.0;
                for (int i = 0; i < n; i++) {
                    double UU_1 = 0.5 * (AT(U, s, j, i, k) + AT(U, s, j, i, k + 1));
                    u1d[i] = 0.5 * (UU_0 + UU_1);
                    UU_0 = UU_1;
                }
                q[0] = 0.0;
                for (int i = 1; i < n; i++) q[i] = AT(vort, s, j, i - 1, k);
                flux1d(u1d, q, flux, n);
                for (int i = 0; i < n; i++)
                    AT(res, s, j, i, k) = AT(res, s, j, i, k) + flux[i];
            }
        free(u1d);
    }
}

/* ---- core/fortran_kinenergy.f90:3-56 ------------------------------------ */
void orc_kin(const double *u, double *ke, double ds2, int l, int m, int n, const ptrdiff_t *s)
{
    const double cff2 = 0.5 * ds2;       /* :43 */
    const double c = cff2 * 0.5;         /* :48, evaluated left to right */
#pragma omp parallel for collapse(2)
    for (int k = 0; k < l; k++)
        for (int j = 0; j < m; j++)
            for (int i = 1; i < n; i++) {
                double a = AT(u, s, k, j, i), b = AT(u, s, k, j, i - 1);
                AT(ke, s, k, j, i) = AT(ke, s, k, j, i) + c * (a * a + b * b);
            }
}

/* ---- core/fortran_bernoulli.f90:2-26 ------------------------------------ */
void orc_gradke(const double *ke, double *du, int l, int m, int n, const ptrdiff_t *s)
{
#pragma omp parallel for collapse(2)
    for (int k = 0; k < l; k++)
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n - 1; i++)
                AT(du, s, k, j, i) = AT(du, s, k, j, i) - (AT(ke, s, k, j, i + 1) - AT(ke, s, k, j, i));
}

/* ---- core/fortran_bernoulli.f90:29-58 ----------------------------------- */
void orc_gradkeandb(const double *ke, const double *b, double *du, double dz,
                    int l, int m, int n, const ptrdiff_t *s)
{
    const double cff = 0.5 * dz;
#pragma omp parallel for collapse(2)
    for (int k = 0; k < l; k++)
        for (int j = 0; j < m; j++)
            for (int i = 0; i < n - 1; i++)
                AT(du, s, k, j, i) = AT(du, s, k, j, i) - (AT(ke, s, k, j, i + 1) - AT(ke, s, k, j, i))
                                     + cff * (AT(b, s, k, j, i + 1) + AT(b, s, k, j, i));
}

/* ---- core/fortran_bernoulli.f90:61-97 ----------------------------------- */
void orc_div(double *d, const double *u, int iflag, int l, int m, int n, const ptrdiff_t *s)
{
#pragma omp parallel for collapse(2)
    for (int k = 0; k < l; k++)
        for (int j = 0; j < m; j++) {
            if (iflag > 0) {
                AT(d, s, k, j, 0) = AT(d, s, k, j, 0) + AT(u, s, k, j, 0);
                for (int i = 1; i < n; i++)
                    AT(d, s, k, j, i) = AT(d, s, k, j, i) + (AT(u, s, k, j, i) - AT(u, s, k, j, i - 1));
            } else {
                AT(d, s, k, j, 0) = AT(u, s, k, j, 0);
                for (int i = 1; i < n; i++)
                    AT(d, s, k, j, i) = AT(u, s, k, j, i) - AT(u, s, k, j, i - 1);
            }
        }
}

/* ---- core/fortran_dissipation.f90:2-35  add_laplacian ("next" row f1) ---- */
void orc_add_laplacian(const double *trac, double *dtrac, double coef,
                       int l, int m, int n, const ptrdiff_t *s)
{
#pragma omp parallel for collapse(2)
    for (int k = 0; k < l; k++)
        for (int j = 0; j < m; j++) {
            double fxm = 0.0;                  /* no flux through the left end */
            for (int i = 0; i < n - 1; i++) {
                double fx = AT(trac, s, k, j, i + 1) - AT(trac, s, k, j, i);
                AT(dtrac, s, k, j, i) = AT(dtrac, s, k, j, i) + coef * (fx - fxm);
                fxm = fx;
            }
            /* no flux through the right end */
            AT(dtrac, s, k, j, n - 1) = AT(dtrac, s, k, j, n - 1) + coef * (0.0 - fxm);
        }
}
