/*
 * nyles_b200.h -- C ABI of libnyles_b200.so
 *
 * B200 (sm_100a) replacement for the two native layers under Nyles' LES time step:
 *   (1) the f2py Fortran kernels  core/fortran_{vorticity,vortex_force,upwind,kinenergy,
 *       bernoulli}.f90 + core/weno.f90          (reference interface: core/Makefile:1-2)
 *   (2) the ctypes-loaded multigrid library libmgmod64.so built from core/mgfor/ (*.f90)
 *       (reference interface: core/mgfordriver.py:14-24, core/build.py:107-121,261-284)
 *
 * Conventions
 *   - plain C: pointers, ints, doubles; no torch / C++ types.
 *   - every `double*` is a DEVICE pointer unless its name ends in `_host`.
 *   - a field is a dense fp64 array F[k][j][i] with extents ny_ext{nz,ny,nx} (i fastest):
 *     the reference's `view('i')` (core/variables.py:146-148).  Extents are LOCAL array
 *     extents, halo rows included on the sides that have a neighbour
 *     (core/mpi/topology.py:253-304).  As in the Fortran, boundary closures are keyed to the
 *     array ends.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are
 *     asynchronous except where stated.  One ny_ctx per GPU, not thread-safe per context.
 *   - return value: 0 on success, negative ny_status on failure; ny_last_error() gives text.
 *     The library never calls exit() (the Fortran `stop`s, mg_setup.f90:337-340,377-381).
 *   - arithmetic: fp64, no FMA contraction, operation order of the Fortran source, including
 *     its REAL(4) literal/tau5 roundings (core/weno.f90:34-46) => bit-identical to the oracle.
 */
#ifndef NYLES_B200_H
#define NYLES_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ny_ctx ny_ctx;
typedef struct ny_mg ny_mg;
typedef struct ny_comm ny_comm;     /* slab communicator: NCCL over NVLink, one rank per GPU */

typedef struct ny_ext { int nz, ny, nx; } ny_ext;

typedef enum ny_status {
    NY_OK = 0,
    NY_ERR_ARG = -1,       /* bad argument (null pointer, extent < 5 on a WENO axis, ...) */
    NY_ERR_CUDA = -2,      /* CUDA runtime error */
    NY_ERR_GRID = -3,      /* grid cannot be coarsened by the multigrid hierarchy */
    NY_ERR_COMM = -4       /* communicator error */
} ny_status;

/* mgfor topology enum, core/mgfor/mg_enums.f90:5-7 */
enum { NY_TOPO_CLOSED = 1, NY_TOPO_XPERIO = 2, NY_TOPO_YPERIO = 3, NY_TOPO_ZPERIO = 4,
       NY_TOPO_XYPERIO = 5, NY_TOPO_XYZPERIO = 6 };

/* level-array selector of set/get_pyarray, core/mgfor/pytools.f90:17-62 */
enum { NY_MG_X = 1, NY_MG_B = 2, NY_MG_R = 3, NY_MG_Y = 4, NY_MG_DIAG = 5, NY_MG_IDIAG = 6,
       NY_MG_MSK = 7, NY_MG_RCOEF = 8, NY_MG_PCOEF = 9 };

/* single multigrid operations (core/mgfor/operators.f90:127-244), exposed for parity tests */
enum { NY_MG_OP_SMOOTH = 1, NY_MG_OP_RESIDUAL = 2, NY_MG_OP_RESTRICTION = 3,
       NY_MG_OP_PROLONGATION = 4, NY_MG_OP_VCYCLE = 5,
       NY_MG_OP_FILL = 6 };  /* halo fill of x and b of a level (mod_halo.f90:200-262); tells the solver
                                that the periodic / slab halos of level 1 are consistent */

typedef struct ny_mg_stats {          /* MG_Stats, core/mgfor/mg_types.f90:45-48 (+normb, history) */
    int nite;                         /* V-cycles done by the last solve */
    int nres;                         /* entries used in reshist */
    double res;                       /* sum(msk r^2)/sum(msk b^2) when the loop exited */
    double normb;                     /* sum(msk b^2) */
    double reshist[32];               /* res before the 1st cycle, after cycle 1, ... */
} ny_mg_stats;

/* kernel families that ny_prof_* can time (CUDA events on the launching stream) */
enum { NY_PROF_RHS_TRACER = 0, NY_PROF_RHS_MOMENTUM, NY_PROF_VORT_KE, NY_PROF_DIV, NY_PROF_GRADP,
       NY_PROF_U_FROM_U, NY_PROF_TIMESCHEME, NY_PROF_MAXSPEED, NY_PROF_HALO,
       NY_PROF_MG_SMOOTH_FINE, NY_PROF_MG_RESIDUAL_FINE, NY_PROF_MG_RESTRICT_FINE,
       NY_PROF_MG_PROLONG_FINE, NY_PROF_MG_NORM, NY_PROF_MG_COARSE, NY_PROF_MG_EMBED,
       NY_PROF_MG_DOWN_FINE,   /* fused smooth + residual + restriction of level 1 */
       NY_PROF_MG_UP_FINE,     /* fused prolongation + smooth (+ residual norm) of level 1 */
       NY_PROF_NTAGS };

/* ---- context ---------------------------------------------------------------------- */
int  ny_init(int device, ny_ctx** out);
void ny_free(ny_ctx* ctx);
const char* ny_last_error(void);
int  ny_version(void);
/* number of kernel launches issued through this context since creation / last reset */
long long ny_launch_count(ny_ctx* ctx);
void ny_launch_count_reset(ny_ctx* ctx);

/* Event timing of kernel families, used by bench.py for the per-kernel roofline.  mask: bit t
 * enables family t.  ny_prof_collect synchronises the device, folds all finished event pairs
 * into per-family totals and returns them: ms[t] (milliseconds) and n[t] (timed groups), arrays
 * of NY_PROF_NTAGS entries.  ny_prof_start resets the totals. */
int  ny_prof_start(ny_ctx* ctx, unsigned long long mask);
int  ny_prof_collect(ny_ctx* ctx, double* ms_host, long long* n_host);
const char* ny_prof_name(int tag);

/* Arithmetic of the WENO kernels.  fast = 0 (default): every operation of the Fortran in source
 * order, results bit-identical to the reference restatement.  fast = 1: beta1, beta3 and the
 * REAL(4) rounding of tau5 are still evaluated exactly (that rounding is discontinuous); the smooth
 * remainder of weno5 is re-associated (one reciprocal instead of four divisions, explicit FMAs).
 * Differences are a few ulp, far inside the 1e-12 parity bar (tests/test_gpu_operators.py). */
int  ny_set_arith(ny_ctx* ctx, int fast);
int  ny_get_arith(ny_ctx* ctx);
/* Which kernel evaluates the fused momentum right-hand side on the cells whose six WENO sweeps are interior
 * (3 <= s <= n-4 on every axis): 0 (default) = the plane-marching TMA kernel (k_mom3) for grids of at least 2^18
 * cells, the cell-parallel kernel below that; 1 = always the cell-parallel kernel; 2 = the TMA kernel wherever it
 * is legal (even nx, 16-byte aligned arrays).  The 3-cell frame around those cells always runs the cell-parallel
 * kernel.  All variants are bit-identical; the switch exists for tests and timing. */
int  ny_set_momentum_variant(ny_ctx* ctx, int variant);

/* ---- f2py kernel replacements ------------------------------------------------------ */
/* fortran_vorticity.vorticity x3 as driven by core/vorticity.py:7-34 (fparam = f*dx*dy, 0 = off) */
int ny_vorticity(ny_ctx*, const double* ux, const double* uy, const double* uz,
                 double* wx, double* wy, double* wz, ny_ext e, double fparam, void* stream);

/* fortran_upwind.upwind x3 as driven by core/tracer.py:44-72: dtrac = -div(U trac) (WENO) */
int ny_upwind(ny_ctx*, const double* trac, const double* Ux, const double* Uy, const double* Uz,
              double* dtrac, ny_ext e, void* stream);

/* same with tracer diffusion interleaved per direction as core/tracer.py:72-77 does when
 * last=True (coefficients = diff_coef*ids2 per axis) */
int ny_upwind_diff(ny_ctx*, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                   double* dtrac, double cx, double cy, double cz, ny_ext e, void* stream);

/* fortran_vortex_force.vortex_force_direc/_flip x3 as driven by core/vortex_force.py:69-81:
 * du += vortex force (six WENO line sweeps) */
int ny_vortex_force(ny_ctx*, const double* Ux, const double* Uy, const double* Uz,
                    const double* wx, const double* wy, const double* wz,
                    double* dux, double* duy, double* duz, ny_ext e, void* stream);

/* fortran_kinenergy.kin x3 as driven by core/kinenergy.py:7-24 */
int ny_kin(ny_ctx*, const double* ux, const double* uy, const double* uz, double* ke,
           double idx2, double idy2, double idz2, ny_ext e, void* stream);

/* fortran_bernoulli.gradke/gradkeandb as driven by core/bernoulli.py:14-32 (euler!=0: no b term) */
int ny_bernoulli(ny_ctx*, const double* ke, const double* b, double* dux, double* duy, double* duz,
                 double dz, int euler, ny_ext e, void* stream);

/* fortran_bernoulli.div x3 as driven by core/projection.py:16-29 */
int ny_div(ny_ctx*, const double* Ux, const double* Uy, const double* Uz, double* div,
           ny_ext e, void* stream);

/* fortran_bernoulli.gradke(p,u_d) x3, core/projection.py:84-87: u -= delta p */
int ny_gradp(ny_ctx*, const double* p, double* ux, double* uy, double* uz, ny_ext e, void* stream);

/* core/cov_to_contra.py:4-20: U = u*idx2, V = v*idy2, W = w*idz2 */
int ny_U_from_u(ny_ctx*, const double* ux, const double* uy, const double* uz,
                double* Ux, double* Uy, double* Uz, double idx2, double idy2, double idz2,
                ny_ext e, void* stream);

/* fortran_dissipation.add_laplacian along all three axes (core/viscosity.py:3-9, tracer.py:74-77):
 * dphi += sum_d coef_d * delta_d delta_d phi with zero-flux ends */
int ny_add_laplacian(ny_ctx*, const double* phi, double* dphi, double cx, double cy, double cz,
                     ny_ext e, void* stream);

/* ---- fused forms of the same arithmetic (identical results, fewer passes over HBM) --- */
/* whole right-hand side of model_les.LES.rhs / model_les_euler (core/model_les.py:130-144):
 * db, du are OVERWRITTEN (the reference zeroes them first).  flags: bit0 euler (no tracer, no b),
 * bit1 linear (no vortex force).  db may be null when euler. */
int ny_rhs(ny_ctx*, const double* b, const double* Ux, const double* Uy, const double* Uz,
           const double* wx, const double* wy, const double* wz, const double* ke,
           double* db, double* dux, double* duy, double* duz,
           double dz, int flags, ny_ext e, void* stream);
/* ny_rhs whose momentum kernel applies the time-scheme update of the three velocity components itself
 * instead of storing du (core/timescheme.py:131-175; u, ub, un = state, stateb, "state copy" arrays of
 * Timescheme): mode 1 = Euler start-up of LFAM3 (un = ub = u; u += dt du), 2 = LFAM3 predictor,
 * 3 = LFAM3 corrector (u = un + dt du).  db is still written; the tracer update stays with ny_ts_*. */
int ny_rhs_update_u(ny_ctx*, const double* b, const double* Ux, const double* Uy, const double* Uz,
                    const double* wx, const double* wy, const double* wz, const double* ke,
                    double* db, double* const u[3], double* const ub[3], double* const un[3],
                    int mode, double dt, double dz, int flags, ny_ext e, void* stream);

/* The right-hand side with the time-scheme update of ALL four prognostic fields applied by the tracer and
 * momentum kernels themselves (core/timescheme.py:131-175): no tendency array is written.  Field index
 * 0 = b (ignored with the Euler flag), 1..3 = u components.  The new value of every field goes to out[f],
 * which must be a separate array: the caller rotates its buffers afterwards --
 *   mode 1 (Euler start-up) and 2 (LFAM3 predictor): out becomes the state, the old state array IS the
 *     "state copy" sn (and, for mode 1, is copied to sb);
 *   mode 3 (LFAM3 corrector, out = sn + dt ds): out becomes the state, sn becomes sb.
 * Mode 2 reads s and sb, mode 3 reads sn (s is still the array the stencils read), mode 1 reads s.
 * add (NULL, or four pointers each of which may be NULL): user tendencies -- what a forcing object adds to
 * dstate after the RHS (core/model_les.py:143-144, experiments/forced_convection/forced_plume.py:76-80) --
 * added to the finished tendency of field f before its update, so a forced model keeps the fused path. */
int ny_rhs_step(ny_ctx*, const double* Ux, const double* Uy, const double* Uz,
                const double* wx, const double* wy, const double* wz, const double* ke,
                const double* const s[4], const double* const sb[4], const double* const sn[4],
                double* const out[4], const double* const add[4], int mode, double dt, double dz, int flags,
                ny_ext e, void* stream);

/* U_from_u + vorticity + kinenergy (core/model_les.py:112-123) in one pass over u */
int ny_diag_post(ny_ctx*, const double* ux, const double* uy, const double* uz,
                 double* Ux, double* Uy, double* Uz, double* wx, double* wy, double* wz, double* ke,
                 double idx2, double idy2, double idz2, double fparam, ny_ext e, void* stream);
/* max(U^2+V^2+W^2) over the whole arrays (core/nyles.py:244-250) as accumulated by the LAST ny_diag_post
 * of this context, which computes it while it writes U; synchronises the stream.  NaN if any entry is. */
int ny_diag_post_max_speed2(ny_ctx*, double* out_host, void* stream);

/* ---- time schemes (core/timescheme.py:113-221), n = number of doubles ------------------ */
int ny_ts_axpy(ny_ctx*, double* s, const double* ds, double a, long long n, void* stream);     /* s += a*ds */
int ny_ts_lfam3_first(ny_ctx*, double* s, const double* ds, double* sb, double* sn, double dt,
                      long long n, void* stream);
int ny_ts_lfam3_pred(ny_ctx*, double* s, const double* ds, double* sb, double* sn, double dt,
                     long long n, void* stream);
int ny_ts_lfam3_corr(ny_ctx*, double* s, const double* ds, const double* sn, double dt,
                     long long n, void* stream);
int ny_ts_rk3_stage2(ny_ctx*, double* s, const double* ds0, const double* ds1, double dt,
                     long long n, void* stream);
int ny_ts_rk3_stage3(ny_ctx*, double* s, const double* ds0, const double* ds1, const double* ds2,
                     double dt, long long n, void* stream);

/* core/nyles.py:244-250: max over the whole arrays of U^2+V^2+W^2.  Synchronises `stream`
 * and writes the scalar to *out_host. */
int ny_max_speed2(ny_ctx*, const double* Ux, const double* Uy, const double* Uz, long long n,
                  double* out_host, void* stream);

/* ---- halo (core/mpi/halo.py:140-178 for one process: periodic wrap of nh-wide strips) ---
 * per[3] = {z,y,x}: nonzero if that axis carries halos that wrap onto this same array. */
int ny_halo_fill_self(ny_ctx*, double* f, ny_ext e, int nh, const int per[3], void* stream);

/* ---- slab communicator (replaces mpi4py / mpi_f08: core/mpi/halo.py, mpitools.py, mgfor/mod_halo.f90,
 * mod_gluesplit.f90).  Ranks are ordered along z; rank r owns the r-th slab.  NCCL is looked up at
 * run time (dlsym), so single-GPU users do not need it.  Bootstrap: rank 0 calls ny_comm_unique_id,
 * the 128 bytes are broadcast by the host program (torch.distributed, MPI, a file ...), every rank
 * calls ny_comm_init. */
int  ny_comm_unique_id(char* id128);
int  ny_comm_init(ny_ctx* ctx, int nranks, int rank, const char* id128, ny_comm** out);
void ny_comm_free(ny_comm* comm);
int  ny_comm_size(ny_comm* comm);
int  ny_comm_rank(ny_comm* comm);
/* face exchanges issued on this communicator and bytes this rank has sent to its slab neighbours (over NVLink
 * peer memory, or through ncclSend on the fallback) since the last reset; a null comm reports zeros */
int  ny_comm_stats(ny_comm* comm, long long* exchanges, long long* bytes_sent, int reset);
/* sum (op_max = 0) or max (op_max != 0) of n <= 8 host doubles over all ranks; synchronises `stream`
 * (core/mpi/mpitools.py:28-32) */
int  ny_comm_allreduce_host(ny_comm* comm, double* values_host, int n, int op_max, void* stream);
/* Halo.fill for z slabs (core/mpi/halo.py:140-178): exchange the nh-plane z faces of `nfields`
 * arrays (fields_host: HOST array of device pointers) with ranks `below` / `above` (-1 = wall or
 * none), then wrap periodic x / y halos locally over all planes.
 * COLLECTIVE, like an MPI halo exchange: every rank of `comm` must issue the same sequence of exchanges (this call,
 * the projections and every multigrid operation on a slab multigrid) in the same order.  The peer-memory path keeps
 * a sequence number per communicator and writes into one of a few rotating slots of the neighbour without waiting
 * for an acknowledgement; a rank that skips an exchange leaves its neighbours spinning until their 60 s watchdog
 * traps ("halo exchange N timed out").  Exchanges of one communicator must not be issued from two streams at once. */
int  ny_halo_exchange(ny_ctx* ctx, ny_comm* comm, double* const* fields_host, int nfields, ny_ext e, int nh,
                      int below, int above, int yper, int xper, void* stream);

/* ---- multigrid (libmgmod64.so replacement) --------------------------------------------- */
/* get_ptrmg(npx=1,npy=1,nx,ny,nz,vertices=F,short=F,is3d=T,topology), mg_setup.f90:318-437 */
int  ny_mg_create(ny_ctx*, int nx, int ny, int nz, int topology, ny_mg** out);
/* the same solver on z slabs, one rank per GPU: nx, ny, nz_global are GLOBAL extents, rank r of
 * `comm` owns planes [r*nz_global/P, (r+1)*nz_global/P).  Level arrays are local slabs padded by nh
 * (z halos filled from the neighbours through NCCL); small levels are gathered and solved
 * redundantly (mg_setup.f90:275-293).  Collective: every rank must make the same calls. */
int  ny_mg_create_slab(ny_ctx*, ny_comm* comm, int nx, int ny, int nz_global, int topology, ny_mg** out);
/* Tuning defaults.  The five setters below change process-wide DEFAULTS that a multigrid copies when it is created
 * (ny_mg_create / ny_mg_create_slab); an existing multigrid keeps the values it was born with.
 * slab multigrids created afterwards gather every level with at most `cells` global cells
 * (default 2 200 000: 128^3 and below; the finest level always stays distributed); a level whose slab is thinner
 * than 4 planes is gathered in any case */
void ny_mg_set_gather_cells(long long cells);
/* slab levels with at least `cells` local cells compute the planes next to their slab neighbours first and
 * exchange them on a second stream while the rest of the slab is computed (default: never -- the peer-memory
 * exchange is cheaper than the split launches; 2^25 is the setting that paid with ncclSend/ncclRecv) */
void ny_mg_set_overlap_cells(long long cells);
/* fused legs: levels whose plane holds at least `tiles` 58 x 24 tiles launch the tiles that keep clear of the
 * x / y walls as a separate, wall-free kernel instance (default 148 = one per SM; tests lower it) */
void ny_mg_set_split_tiles(long long tiles);
/* the levels with at most `cells` cells (default 2048) form the tail of the V-cycle that one single-CTA launch
 * runs from the first smoothing down to the coarsest level and back (0: off) */
void ny_mg_set_tail_cells(long long cells);
/* ... and the replicated levels above those with at most `cells` cells (default 300 000: 64^3 and below, or all of
 * a 128 x 32 x 32 grid) join that launch as its "wide" levels: a cooperative launch of one CTA per SM runs them with a
 * grid barrier between the operators, CTA 0 runs the single-CTA levels (0: no wide levels).  Same arithmetic, same
 * results as the fused legs and the per-operator kernels. */
void ny_mg_set_wide_cells(long long cells);
void ny_mg_destroy(ny_mg*);
int  ny_mg_nlevels(ny_mg*);
/* 1 if the mask is the default box, so that the fused analytic-coefficient kernels are in use */
int  ny_mg_is_box(ny_mg*);
/* on = 0 forces the generic kernels that read the coefficient arrays (one rank only; used by the
 * tests to cross-check the two paths), on = 1 re-verifies and re-enables the box kernels */
int  ny_mg_set_fast_path(ny_mg*, int on);
/* on = 0: V-cycles run one box kernel per operator; on = 1 (default): the fused TMA-staged legs
 * (smooth+residual+restriction, prolongation+smooth[+norm]) wherever a level allows them, tiles away from
 * the x / y walls through the wall-free kernel instance and the smallest levels of a closed box in one launch;
 * on = 2: fused legs, general instance for every tile, no one-launch tail */
int  ny_mg_set_fused_legs(ny_mg*, int on);
/* 1-based index of the first level that is replicated on every rank (1 on a single rank) */
int  ny_mg_first_gathered_level(ny_mg*);
/* get_pyshape: shape[0..2] = (nz+2nh, ny+2nh, nx+2nh) of level lev (1-based), numpy order */
int  ny_mg_shape(ny_mg*, int lev, int shape[3]);
/* MG_Param fields that matter (mg_types.f90:15-26); defaults maxite=20, tol=1e-6, omega=0.9 */
int  ny_mg_set_param(ny_mg*, int maxite, double tol, double omega);
/* set_pyarray / get_pyarray: whole level arrays, device to device */
int  ny_mg_set_array(ny_mg*, int lev, int ivar, const double* src, void* stream);
int  ny_mg_get_array(ny_mg*, int lev, int ivar, double* dst, void* stream);
/* setup_fine_msk + setup_operators (mg_setup.f90:213-223, operators.f90:461-505): rebuild the coarse masks
 * and Rcoef, Pcoef, diag, idiag of every level after the mask of level 1 was changed with
 * ny_mg_set_array(.., ivar = 7, ..) -- obstacles, as in core/mgfor/tests.f90:207-212.  Single rank only. */
int  ny_mg_setup_operators(ny_mg*, void* stream);
/* solvers.f90:8-33.  Synchronises `stream` once per V-cycle for the stopping test. */
int  ny_mg_solve(ny_mg*, ny_mg_stats* stats_host, void* stream);
/* mgfordriver.MG.solve_directly (core/mgfordriver.py:66-78) without the two host copies:
 * b_mg[idx] = div; solve; p = x_mg[idx]*scale.  lo[3] = {k0,j0,i0}: offset of the model array
 * inside the padded MG array (nh on sides without neighbour, 0 on sides with). */
int  ny_mg_solve_directly(ny_mg*, double* p, const double* div, ny_ext e, const int lo[3],
                          double scale, ny_mg_stats* stats_host, void* stream);
/* projection.compute_p (core/projection.py:39-87) around the in-place solve: div = delta(u*ids2) is
 * written to `div` and embedded in b (then halo-filled there, which is what halo.fill(div) achieves in
 * the reference), solve, p = x*scale, u -= delta p.  The halo cells of the model's `div` array are not
 * refreshed. */
int  ny_mg_project(ny_mg*, double* ux, double* uy, double* uz, double* div, double* p,
                   double idx2, double idy2, double idz2, ny_ext e, const int lo[3], double scale,
                   ny_mg_stats* stats_host, void* stream);
int  ny_mg_op(ny_mg*, int op, int lev, void* stream);

/* ---- the linear (non-WENO) upwind branch ---------------------------------------------------------
 * What fortran_upwind.upwind and fortran_vortex_force.vortex_force_direc / _flip compute when their local flag
 * `linear` is .true. (core/fortran_upwind.f90:33-64, core/fortran_vortex_force.f90:39-64,118-143, with
 * core/interpolate_tracer.f90 / core/interpolate.f90; orders 1..5, REAL(4) coefficients).  The reference ships with
 * linear = .false., so these are never reached there; nyles_b200.LINEAR_UPWIND switches the models onto them. */
int  ny_upwind_linear(ny_ctx*, const double* trac, const double* Ux, const double* Uy, const double* Uz,
                      double* dtrac, int order, ny_ext e, void* stream);
int  ny_vortex_force_linear(ny_ctx*, const double* Ux, const double* Uy, const double* Uz,
                            const double* wx, const double* wy, const double* wz,
                            double* dux, double* duy, double* duz, int order, ny_ext e, void* stream);

/* ---- arithmetic primitives, exposed for the parity tests -------------------------------------
 * ny_debug_weno5: out[t] = weno5(q0[t], q1[t], q2[t], q3[t], q4[t]) (core/weno.f90:25-54) in the
 * context's arithmetic mode; q is 5 device arrays of n doubles stored back to back.
 * ny_debug_weno3: out[t] = weno3(q0[t], q1[t], q2[t]) (core/weno.f90:1-22); q is 3 arrays back to back.
 * ny_debug_div: out[t] = the in-range division of ny_weno.cuh applied to a[t] / b[t]; mismatch_host
 * receives the number of t for which it differs (bitwise) from the IEEE quotient. */
int  ny_debug_weno5(ny_ctx*, const double* q, double* out, long long n, void* stream);
int  ny_debug_weno3(ny_ctx*, const double* q, double* out, long long n, void* stream);
/* measurement aid: sustained issue rate of the fp64 pipe (DFMA only, ~`seconds` of back-to-back launches under the
 * board's power cap), in warp-wide fp64 instructions per second over the whole device; synchronises `stream` */
int  ny_debug_fp64_peak(ny_ctx*, double seconds, double* dp_warp_instr_per_s, void* stream);
int  ny_debug_div(ny_ctx*, const double* a, const double* b, double* out, long long n,
                  long long* mismatch_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NYLES_B200_H */
