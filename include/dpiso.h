/*
 * dpiso.h -- C ABI of libdpiso (sm_100a).  The drop-in boundary for the PISO-step hot path of
 * tum-pbs/differentiable-piso.  Every entry point replaces one launcher behind a TF-1.14 custom op of
 * the reference (paths relative to the reference checkout):
 *
 *   dpiso_csr_structure / dpiso_assemble   CentralDifferenceMatrixCsrKernelLauncher
 *                                          CUDAsrc/central_difference_csr_op.cc:33-36, .cu.cc:543-664
 *   dpiso_bicgstab_*                       MultiBicgstabIluLinearSolveLauncher
 *                                          CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cc:50-58, .cu.cc:85-531
 *   dpiso_laplace_*                        LaplaceMatrixKernelLauncher
 *                                          CUDAsrc/pressure_solve_op.cc:78-84, laplace_op.cu.cc:191-239
 *   dpiso_pressure_cg_*                    LaunchPressureKernel
 *                                          CUDAsrc/pressure_solve_op.cc:48-76, .cu.cc:140-696
 *   dpiso_predictor_rhs / fv_* / corrector_* / h_apply   the TF graph ops of diffpiso/piso_tf.py:36-75
 *                                          and diffpiso/piso_helpers.py:169-310 (and their gradients)
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name starts with h_;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - every function returns 0 on success, a negative DPISO_E* code otherwise (never exit()/assert());
 *     dpiso_last_error() returns a thread-local message for the last failure;
 *   - batch-major contiguous arrays: face vectors [batch][n_u+n_v] = [u rows..., v rows...] (x fastest),
 *     cell vectors [batch][ny*nx], CSR values [batch][nnz_u+nnz_v]; masks are shared by the batch and
 *     are padded-centred [(ny+2)*(nx+2)];
 *   - "ny,nx" is the centred resolution; u lives on ny x (nx+1) faces, v on (ny+1) x nx faces.
 */
#ifndef DPISO_H
#define DPISO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPISO_OK 0
#define DPISO_EINVAL (-1)      /* bad argument (grid too small, null pointer, ...) */
#define DPISO_ECUDA (-2)       /* CUDA runtime error, see dpiso_last_error() */
#define DPISO_EUNSUPPORTED (-3) /* configuration outside what the kernels cover */

/* pressure ghost-cell rule per side (PhiFlow extrapolation of the centred field) */
#define DPISO_PBC_REPLICATE 0
#define DPISO_PBC_ZERO 1
#define DPISO_PBC_PERIODIC 2

int dpiso_version(void);
const char *dpiso_last_error(void);

/* ---- sizes (host only; diffpiso/piso_tf.py:99-106) ---------------------------------------------- */
/* h_n[2] = rows of (u, v); h_nnz[2] = stored entries of (u, v) */
int dpiso_sizes(int ny, int nx, int per_x, int per_y, int *h_n, int *h_nnz);

/* ---- CSR structure: row_ptr = two 0-based arrays back to back (n_u+1, n_v+1), col_ind 0-based per
 *      component (calcCsrRowPtrGpu + the colInd stores of calcAdvetionMatrixX/Y). ------------------- */
int dpiso_csr_structure(int ny, int nx, int per_x, int per_y, int *row_ptr, int *col_ind, void *stream);

/* ---- advection-diffusion matrices (custom_padded + CentralDifferenceMatrixCsr) -------------------
 * vel [batch][nf] unpadded; the padding of piso_helpers.py:35-55 is applied on the fly.
 * dy, dx = grid_spacing (y, x); area_x, area_y = the op's cell_area input = prod(dx)/(dx, dy) as fp32 (piso_tf.py:96-97),
 * i.e. the area of a face normal to x (~dy) and normal to y (~dx).
 * dirichlet uint8 [nf]; active float [(ny+2)(nx+2)]; noslip uint8 [(ny+2)(nx+2)];
 * visc: visc_mode 0 = scalar (1 float), 1 = face field [nf] shared, 2 = face field [batch][nf].
 * per_x / per_y: bit 0 = periodic axis; bit 1 (only with bit 0) = pad the velocity of that axis by replication
 * although the matrix is periodic -- what the reference computes from the second step of run_piso_steps on, where the
 * re-wrapped state loses its extrapolation (combined_training_integrated.py:431-432,473-474).
 * outputs: values [batch][nnz] (centre = diag - beta), a_diag [batch][nf]. */
int dpiso_assemble(int batch, int ny, int nx, int per_x, int per_y, float dy, float dx, float area_x, float area_y,
                   float beta, const float *vel, const uint8_t *dirichlet, const float *active, const uint8_t *noslip,
                   const float *visc, int visc_mode, float *values, float *a_diag, void *stream);

/* ---- pointwise pieces of piso_step (forward) ------------------------------------------------------ */
/* rhs = vel*beta - G(p) [+ forcing*dx*dy]; Dirichlet faces <- -dirichlet_value   (piso_tf.py:36-40)
 * h_pbc[4] = ghost rule at y_lo,y_hi,x_lo,x_hi; dvals [dvals_batch ? batch : 1][nf]; forcing may be NULL */
int dpiso_predictor_rhs(int batch, int ny, int nx, float dy, float dx, float beta, const int *h_pbc,
                        const float *vel, const float *pres, const float *access, const uint8_t *dirichlet,
                        const float *dvals, int dvals_batch, const float *forcing, float *rhs, void *stream);
/* g = G(p) * min(accessible+, accessible-)   (piso_helpers.py:236-274) */
int dpiso_fv_gradient(int batch, int ny, int nx, float dy, float dx, const int *h_pbc, const float *access,
                      const float *p, float *g, void *stream);
/* div = D(vel)  or, with a_diag != NULL,  D(vel / (beta - a_diag))   (piso_helpers.py:285-289, piso_tf.py:66) */
int dpiso_fv_divergence(int batch, int ny, int nx, float dy, float dx, const float *vel, const float *a_diag,
                        float beta, float *div, void *stream);
/* u** = u* - G(p1)/(beta-A)/(dx*dy)   (piso_tf.py:58) */
int dpiso_corrector1(int batch, int ny, int nx, float dy, float dx, float beta, const int *h_pbc,
                     const float *access, const float *u_star, const float *p1, const float *a_diag,
                     float *u_s2, void *stream);
/* h = M (u**-u*) - (A-beta)(u**-u*)    (piso_helpers.py:209-223) */
int dpiso_h_apply(int batch, int ny, int nx, int per_x, int per_y, float beta, const float *values,
                  const float *a_diag, const float *u_star, const float *u_s2, float *h, void *stream);
/* u_next = u** + (h - G(p2)/(dx*dy))/(beta-A);  p_next = p + p1 + p2     (piso_tf.py:71-75) */
int dpiso_corrector2(int batch, int ny, int nx, float dy, float dx, float beta, const int *h_pbc,
                     const float *access, const float *u_s2, const float *h, const float *p2,
                     const float *a_diag, const float *p, const float *p1, float *u_next, float *p_next,
                     void *stream);

/* ---- pointwise adjoints (frozen coefficients) --------------------------------------
 * These are the gradients TF-1.14 autodiff assembles from the reference's registrations; on periodic axes the
 * registered gradients of the FV gradient / divergence are NOT the exact transposes and are
 * reproduced as registered. */
/* gp = [base] + G^T(t),  t = gs [/(beta - a_diag)] [/divisor] [negated] on the faces; gradient of
 * finite_volume_gradient_tensor incl. the accessible-mask multiply (piso_helpers.py:226-266).
 * a_diag and base may be NULL; divisor = 1 disables the division. */
int dpiso_fv_gradient_adj(int batch, int ny, int nx, float dy, float dx, const int *h_pbc, const float *access,
                          const float *gs, const float *a_diag, float beta, float divisor, int negate,
                          const float *base, float *gp, void *stream);
/* gv = ([base [- base_sub]] + D^T gc) [/(beta - a_diag)]: registered gradient of finite_volume_divergence
 * (piso_helpers.py:291-305).  base, base_sub [batch][nf] and a_diag may be NULL (base_sub needs base). */
int dpiso_fv_divergence_adj(int batch, int ny, int nx, int per_x, int per_y, float dy, float dx, const float *gc,
                            const float *base, const float *base_sub, const float *a_diag, float beta, float *gv,
                            void *stream);
/* gfree = (1-m) grhs, gvel = gfree*beta, gforce = gfree*dx*dy, gdvals = -m grhs  (m = Dirichlet mask): adjoint of the
 * rhs assembly (piso_tf.py:36-40); gforce / gdvals may be NULL.  The pressure part is -G^T(gfree).
 * solve_stats (optional): the stats [batch][2][4] of the transposed predictor solve that produced grhs; samples whose
 * solve raised the NaN warning get grhs * (1 - warn) = 0 (linear_solver.py:169-173), sample by sample. */
int dpiso_predictor_rhs_adj(int batch, int ny, int nx, float dy, float dx, float beta, const uint8_t *dirichlet,
                            const float *grhs, const int *solve_stats, float *gvel, float *gforce, float *gdvals,
                            float *gfree, void *stream);

/* ---- ILU0-preconditioned BiCGStab, batched over samples and the two components --------------------
 * Structure tables are built once per grid by dpiso_bicg_tables_create (below) -- or by the caller -- and stay on the
 * device (all pointers inside dpiso_bicg_tables are DEVICE pointers; the struct itself lives on the host).
 * M = A for the forward solve, M = A^T for the adjoint solve -- only the tables differ.
 * One persistent CTA per (sample, component) system. */
typedef struct {
    int n;                /* rows of this component */
    int n_levels;         /* wavefront levels (lx+ly hyperplanes) */
    int wa;               /* ELL width: max entries per row of the (possibly transposed) matrix, <= 6 */
    int max_level;        /* rows in the largest level */
    int wl, wu;           /* max number of lower / upper entries in a row */
    int dx;               /* faces per grid row of this component (original row index = ly * dx + lx) */
    int rows_ok;          /* 1: every row fits the canonical slot layout of the row-major solver kernel */
    const int *level_ptr; /* [n_levels+1] first position of each level (level-major numbering) */
    const int *perm;      /* [n] level-major position -> original row */
    const int *a_col;     /* [wa][n] column (as level-major position) per entry, in ascending ORIGINAL column
                             order; padding entries come last and point at the row itself */
    const int *a_src;     /* [wa][n] index into this component's CSR values holding M(row, col); -1 = padding */
    const int *a_rev;     /* [wa][n] index into this component's CSR values holding M(col, row); -1 = absent */
    const int *r_col;     /* [wa][n] the same three tables in ORIGINAL row order with original column indices (padding */
    const int *r_src;     /*         entries point at the row itself / -1): used by the row-major solver kernel, which  */
    const int *r_rev;     /*         needs no permutation; may be NULL (then rows_ok must be 0)                          */
    /* canonical slots of the row-major kernel (lower: [far below the y-neighbour, y-neighbour, far above it, x-neighbour],
     * upper: [x-neighbour, far below the y-neighbour, y-neighbour, far above it]); all [n] in original row order */
    const int *c_lsrc;    /* int[n][4] CSR value index of the lower slots, -1 = absent */
    const int *c_lrev;    /* int[n][4] CSR value index of their reverse entries M(col, row) */
    const int *c_usrc;    /* int[n][4] CSR value index of the upper slots */
    const int *c_lfar;    /* int[n][2] column of the two far lower slots, -1 = absent */
    const int *c_ufar;    /* int[n][2] column of the two far upper slots */
    const int *c_dsrc;    /* int[n]    CSR value index of the diagonal entry */
    /* level-major positions (the row-major kernel stores its planes and vectors level by level so that the wavefront
     * sweeps read consecutive addresses): all [n] indexed by level-major position q */
    const int *m_nbr;     /* int[n][4] positions of the regular neighbours (x-1, y-1, x+1, y+1), q itself where absent */
    const int *m_lfar;    /* int[n][2] positions of the two far lower slots, -1 = absent */
    const int *m_ufar;    /* int[n][2] positions of the two far upper slots, -1 = absent */
    void *owner;          /* allocation behind the arrays when built by dpiso_bicg_tables_create*, else NULL */
    int owner_is_host;    /* 1: `owner` is host memory (dpiso_bicg_tables_create_host) */
    int sym;              /* 1: the pattern is structurally symmetric (every entry has its reverse entry): ILU(0)(M^T)
                             is then exactly the transposed ILU(0)(M) and factor reuse is equivalent;
                             0 for components that are periodic along their staggered axis */
    int band_ok;          /* 1: every far (periodic wrap) entry is described by `far` below: the cluster-per-system kernel
                             for large grids applies */
    int far[8];           /* closed form of the far operands, lower entries then upper entries: {xa, xb, ya, yb} -- a row
                             takes, at column xa, its own value of column xb (in-row wrap, canonical slot 2 / 1), and
                             grid row ya takes the value of grid row yb in the same column (in-column wrap, slot 0 / 3);
                             -1 = no such entry */
} dpiso_bicg_tables;

/* Builds the tables of one component (comp 0 = u, 1 = v) for M = A (transpose 0) or M = A^T (transpose 1) from the
 * grid alone and uploads them to the current device (one allocation, stream-ordered copy).  The reference launcher
 * takes just the CSR arrays and transpose_op (multi_bicgstab_ilu_linear_solve_op.cc:50-58) and re-analyses the pattern
 * with cuSPARSE on every call (.cu.cc:113-134, 181-228); here the analysis is done once per grid.  Returns
 * DPISO_EUNSUPPORTED when lx+ly is not a valid level schedule or ILU(0) would interact with fill on this grid.
 * dpiso_bicg_tables_create_host builds the same tables in host memory (tests, inspection). */
int dpiso_bicg_tables_create(int ny, int nx, int per_x, int per_y, int comp, int transpose, dpiso_bicg_tables *out,
                             void *stream);
int dpiso_bicg_tables_create_host(int ny, int nx, int per_x, int per_y, int comp, int transpose, dpiso_bicg_tables *out);
int dpiso_bicg_tables_destroy(dpiso_bicg_tables *tab);

/* gd = M^T gh - (A - beta) gh: adjoint of dpiso_h_apply w.r.t. (u** - u*); takes the tables of M = A^T.
 * sum (optional, may be NULL): sum = base + gd, the accumulated gradient of u** (base [batch][nf] required then) */
int dpiso_h_apply_adj(int batch, const dpiso_bicg_tables *h_tabT_u, const dpiso_bicg_tables *h_tabT_v, int nnz_u,
                      int nnz_v, float beta, const float *values, const float *a_diag, const float *gh, float *gd,
                      const float *base, float *sum, void *stream);

/* workspace (floats) needed per (sample, component) system for the given tables */
size_t dpiso_bicgstab_workspace_floats(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v);

/* Solves (sign * values) x = rhs for every sample and both components; negate != 0 means sign = -1, i.e. the caller
 * passes the assembled M and the solve runs on -M as piso_tf.py:42 does, without materialising the negated copy.
 * values [batch][nnz_u+nnz_v] (nnz_* = CSR entries of the NON-transposed pattern; transposition is expressed by
 * the tables), rhs/x0/x [batch][n_u+n_v].  stats int [batch][2][4] = iterations, restarts, warn, exit kind (every
 * entry is written).  warn float [1]: reset to 0 by the call, set to 1 by any NaN-norm warning
 * (multi_bicgstab...cu.cc:251-256; the float the Python class returns, linear_solver.py:175).
 * workspace: float [batch*2*dpiso_bicgstab_workspace_floats()], caller-owned, contents undefined afterwards.
 * Factor reuse (north_star: "the adjoint solves reuse the forward factorisation"):
 *   pivots_out (optional) [batch][n_u+n_v]: receives the ILU(0) pivots u_ii of this solve;
 *   pivots_in  (optional) [batch][n_u+n_v]: pivots of the solve with the OTHER orientation of the same matrices.  The
 *   wavefront factorisation is skipped: ILU(0)(M^T) = (U^T D^-1)(D L^T), i.e. l'_ik = m_ik / d_k, upper entries
 *   unchanged.  Exact (to rounding) for structurally symmetric patterns; for components that are
 *   periodic along their staggered axis it is a different -- still valid -- preconditioner.
 *   By default pivots_in is honoured per component only where the tables say `sym` (exact reuse); the other
 *   component factorises as usual.  dpiso_bicgstab_set_reuse_policy(1) extends it to every component.
 * Both need dpiso_bicgstab_supports_factor_reuse() != 0 (the row-major kernel); otherwise they are ignored /
 * left untouched and the solve factorises as usual. */
int dpiso_bicgstab_ilu(int batch, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int nnz_u,
                       int nnz_v, const float *values, int negate, const float *rhs, const float *x0, float tol,
                       int max_it, float *x, int *stats, float *warn, float *pivots_out, const float *pivots_in,
                       float *workspace, void *stream);
/* The fp64 variant, LinearSolverCudaMultiBicgstabILU(cast_to_double=True) (diffpiso/linear_solver.py:130-133, launcher
 * multi_bicgstab_ilu_linear_solve_op.cu.cc:540-988): fp32 values / rhs / x0 in, fp64 solve, fp32 solution out (the casts of
 * linear_solver.py:131-133,171 are fused).  Same tables, stats and warn as above; workspace: caller-owned,
 * batch * 2 * dpiso_bicgstab_f64_workspace_bytes() bytes.  No shipped script of the reference uses this path: one CTA per
 * system, one barrier per wavefront level. */
size_t dpiso_bicgstab_f64_workspace_bytes(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v);
int dpiso_bicgstab_ilu_f64(int batch, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int nnz_u, int nnz_v,
                           const float *values, int negate, const float *rhs, const float *x0, float tol, int max_it,
                           float *x, int *stats, float *warn, void *workspace, void *stream);
int dpiso_bicgstab_supports_factor_reuse(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v);
/* profiling / test switch: dbg >= 0 overrides the DPISO_BICG_DBG environment variable, -1 restores it.  Bits: 8 force the
 * level-major fallback kernel, 64 the cluster-per-system kernel, 128 the row-per-thread kernel, 256 the tile kernel,
 * 1024 tile kernel with TMA bulk copies instead of per-lane cp.async, 16 / 32 / 512 A/B variants of the sweeps */
int dpiso_bicgstab_set_debug(int dbg);
/* 0 (default): reuse pivots only for structurally symmetric components; 1: for every component */
int dpiso_bicgstab_set_reuse_policy(int always);
/* tuning hook of the cluster-per-system kernel (grids with more than 512 rows per component): CTAs per system, 0 = heuristic */
int dpiso_bicgstab_set_band_cluster(int cluster);
/* tuning hook of the default (register-tiled) kernel: CTAs per system, 0 = heuristic; values below the minimum the grid
 * needs are ignored */
int dpiso_bicgstab_set_tile_cluster(int cluster);

/* profiling hook: dev_counters = device buffer of 8 int64 SM-cycle counters accumulated by system 0 of every following
 * solve ([0] setup, [1] ILU(0), [2] triangular sweeps, [4] SpMV / vector phases); NULL disables */
int dpiso_bicgstab_set_timing(long long *dev_counters);

/* ---- pressure matrix (calcPISOLaplaceMatrix) ----------------------------------------------------------
 * mode 0: k_faces [batch][n_v+n_u] is the scaling field flattened [v,u] (piso_cuda_pressure_solver.py:70)
 * mode 1: k_faces is a_diag [batch][n_u+n_v] ([u,v]); k = (1/(beta-a))*dx_factor is formed on the fly.
 * lap [batch][ny*nx][5] = [y-, x-, diag, x+, y+] */
int dpiso_laplace_f64(int batch, int ny, int nx, const float *active, const float *fluid, const float *k_faces,
                      int mode, float beta, float dx_factor, double *lap, void *stream);
int dpiso_laplace_f32(int batch, int ny, int nx, const float *active, const float *fluid, const float *k_faces,
                      int mode, float beta, float dx_factor, float *lap, void *stream);

/* ---- pressure CG (LaunchPressureKernel, init_with_zeros = true, randomized_restarts = 0) ------------
 * One thread-block cluster per sample, state resident in registers/shared memory; every sample follows the
 * reference's B=1 control flow (check cadence, residual resets) on its own.
 * lap [batch][n_c][5], div [batch][n_c], x [batch][n_c] (output, T), iterations int [batch].
 * x32 (optional, may be NULL): float copy of x (the reference casts the result to fp32). */
/* workspace: grids whose state fits on chip (BASELINE configs 1-4) need none; larger grids keep p, r, z (and x when
 * the caller passes no T-typed x) in a caller-owned scratch of dpiso_pressure_cg_workspace_bytes() bytes (0 = none
 * needed; `elem_size` 8 = fp64 / mixed, 4 = fp32; `have_x` = the caller passes a T-typed x).  Nothing is allocated
 * inside the calls. */
size_t dpiso_pressure_cg_workspace_bytes(int batch, int ny, int nx, int elem_size, int have_x);
int dpiso_pressure_cg_f64(int batch, int ny, int nx, int per_x, int per_y, const double *lap, const double *div,
                          float accuracy, int max_it, int residual_reset, int rank_deficient, double *x,
                          float *x32, int *iterations, void *workspace, void *stream);
int dpiso_pressure_cg_f32(int batch, int ny, int nx, int per_x, int per_y, const float *lap, const float *div,
                          float accuracy, int max_it, int residual_reset, int rank_deficient, float *x,
                          float *x32, int *iterations, void *workspace, void *stream);
/* fp32 divergence in, fp64 solve, fp32 pressure out: the cast_to_double=True path of
 * PisoPressureSolverCudaCustom.solve (piso_cuda_pressure_solver.py:55-58,111) without materialising the casts */
int dpiso_pressure_cg_mixed(int batch, int ny, int nx, int per_x, int per_y, const double *lap, const float *div32,
                            float accuracy, int max_it, int residual_reset, int rank_deficient, float *x32,
                            int *iterations, void *workspace, void *stream);

/* launch-configuration report for the last pressure CG call on this thread: h_out[0]=cluster size,
 * [1]=threads per CTA, [2]=cells per thread, [3]=dynamic smem bytes, [4]=kernel variant */
int dpiso_pressure_cg_last_config(int *h_out);
/* tuning hook for benchmarks / tests: force the cluster size (0 = heuristic) and the kernel variant (-1 = heuristic;
 * 0: 1024 threads, coefficients in smem; 1: 512 threads, coefficients in registers; 2: 512 threads, 2 CTAs/SM) */
int dpiso_pressure_cg_set_tuning(int cluster, int variant);
/* parity-measurement switch: 1 = the reference's reduction order of pressure_solve_op.cu.cc:571-634 ({p.r, p.z} ->
 * alpha -> update x, r -> {r.z} -> beta: two cluster-wide reductions per iteration) instead of the default merged
 * single reduction (r_new.z = r.z - alpha z.z, DESIGN.md deviation D2).  Cluster-resident kernel only. */
int dpiso_pressure_cg_set_reduction_order(int two_reductions);
/* A/B switch: 0 = never use the instantiations with a compile-time row length (nx = 128), 1 (default) = use them */
int dpiso_pressure_cg_set_static_nx(int enable);

#ifdef __cplusplus
}
#endif
#endif /* DPISO_H */
