/*
 * oracle/piso_oracle.c -- CPU restatement of the diffpiso PISO-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under differentiable-piso_b200/ may link,
 * import or call this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, and only as the checker / the
 * reported CPU baseline.
 *
 * Parity status: the reference ships no golden vectors for this path; the pins are made from the
 * reference itself:
 *   - assembly, CSR structure, Laplace matrix, pressure CG: the reference's own kernels compiled from
 *     /root/reference/CUDAsrc into oracle/_ref and run on a B200 (tests/test_gpu_reference_pin.py), frozen as
 *     tests/golden/ref_kernels/*.npz and checked on every CPU run (tests/test_cpu_golden.py): bit-exact;
 *   - the step's glue (padding, constants, rhs, correctors, H, pressure accumulation) and the backward pass of the
 *     registered gradients: the reference's own Python executed from source (tests/golden/reference_runner.py), frozen as
 *     tests/golden/ref_python/step_*.npz (tests/test_cpu_reference_python.py): forward bit-identical, gradients to
 *     solver tolerance;
 *   - BiCGStab+ILU0: solutions pinned against the reference's CPU solver path (spsolve on CSR built by the reference's
 *     convert_to_scipy_csr).  The cuSPARSE-10 arithmetic itself (csrilu02 / csrsv2 / CsrmvEx / csr2csc no longer exist
 *     in CUDA 12.9) cannot be rebuilt, so iteration counts and ILU(0) rounding are UNPINNED there: the algorithm below
 *     is the textbook one restated from the call sequence in
 *     CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:233-411.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Layout conventions:
 *   centred resolution (ny, nx), x fastest; u lives on ny x (nx+1) faces, v on
 *   (ny+1) x nx faces; flat face vectors are [u rows..., v rows...].
 *   padded-centred masks are (ny+2) x (nx+2).
 *
 * Floating point: compiled with -ffp-contract=off; every fused multiply-add is an
 * explicit fmaf()/fma() so that the rounding sequence is the one the CUDA kernels
 * use (nvcc contracts a*b+c by default, as it did for the reference's kernels).
 * Dot products / norms of fp32 vectors accumulate in fp64 and round once.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_EBADGRID 1

/* ------------------------------------------------------------------------------------------
 * sizes and CSR structure
 * ---------------------------------------------------------------------------------------- */

static void comp_dims(int ny, int nx, int comp, int *Dx, int *Dy) {
    if (comp == 0) { *Dx = nx + 1; *Dy = ny; } else { *Dx = nx; *Dy = ny + 1; }
}

/* diffpiso/piso_tf.py:99-106 (dimensions, dim_product, matrix_nnz) */
int orc_sizes(int ny, int nx, int per_x, int per_y, int *n_rows, int *nnz) {
    if (ny < 3 || nx < 3) return ORC_EBADGRID;
    for (int c = 0; c < 2; c++) {
        int Dx, Dy; comp_dims(ny, nx, c, &Dx, &Dy);
        int n = Dx * Dy;
        n_rows[c] = n;
        nnz[c] = 5 * n - 2 * (n / Dx) * (1 - per_x) - 2 * (n / Dy) * (1 - per_y);
    }
    return ORC_OK;
}

/* CUDAsrc/central_difference_csr_op.cu.cc:472-505 (calcCsrRowPtrGpu, 2-D branch):
 * closed-form row pointer = 5*(row+1) minus the neighbours that are missing up to and
 * including this row. */
static int rowptr_closed_form(int lx, int ly, int Dx, int Dy, int per_x, int per_y) {
    int row = lx + Dx * ly;
    int rp = (row + 1) * 5;
    int ly_pos = ly < 1 ? ly : 1;                                   /* min(ly,1) */
    int on_top = 1 + ((ly + 1 - Dy) > -1 ? (ly + 1 - Dy) : -1);      /* 1 iff ly == Dy-1 */
    int on_right = 1 + ((lx + 1 - Dx) > -1 ? (lx + 1 - Dx) : -1);    /* 1 iff lx == Dx-1 */
    rp -= ly_pos * (Dx * (1 - per_y));
    rp -= ((1 - ly_pos) + on_top) * (lx + 1) * (1 - per_y);
    rp -= (ly * 2 + 1 + on_right) * (1 - per_x);
    return rp;
}

/* CUDAsrc/central_difference_csr_op.cu.cc:166-210: position of each neighbour inside
 * the CSR row.  slot[k], k = 0:x-  1:x+  2:y-  3:y+  4:centre, relative to the row start;
 * has[k] tells whether the entry exists structurally. */
static void row_slots(int lx, int ly, int Dx, int Dy, int per_x, int per_y, int slot[5], int has[4]) {
    const int per[2] = { per_x, per_y };
    int reg[4] = { lx > 0, lx < Dx - 1, ly > 0, ly < Dy - 1 };  /* regular (non-wrapped) neighbour */
    /* start from the interior ordering  y-, x-, c, x+, y+  (":176-180") */
    slot[0] = 1; slot[1] = 3; slot[2] = 0; slot[3] = 4; slot[4] = 2;
    /* periodic wrap re-ordering (":185-196") */
    for (int d = 1; d >= 0; d--) {
        int lo = 2 * d, hi = 2 * d + 1;
        slot[lo] += (d + 1) * 2 * (1 - reg[lo]) * per[d];
        slot[lo] += (1 - reg[hi]) * per[d];
        slot[hi] -= (d + 1) * 2 * (1 - reg[hi]) * per[d];
        slot[hi] -= (1 - reg[lo]) * per[d];
        int shift = (reg[lo] - reg[hi]) * per[d];
        for (int e = 0; e < d; e++) { slot[2 * e] += shift; slot[2 * e + 1] += shift; }
        slot[4] += shift;
    }
    /* remove entries across non-periodic boundaries (":199-210") */
    int ord[5];
    for (int k = 0; k < 5; k++) ord[slot[k]] = k;
    for (int pos = 0; pos < 4; pos++) {
        int k = ord[pos];
        if (k == 4) continue;
        int drop = (1 - reg[k]) * (1 - per[k / 2]);
        for (int q = pos + 1; q < 5; q++) slot[ord[q]] -= drop;
    }
    for (int k = 0; k < 4; k++) has[k] = reg[k] || per[k / 2];
}

/* neighbour column of direction k (":219-231, 259-286"); stag_x/stag_y mark the
 * component's own staggered axis (wrap skips the duplicated face). */
static int nb_col(int k, int row, int lx, int ly, int Dx, int Dy, int stag_x, int stag_y) {
    switch (k) {
    case 0: return lx > 0 ? row - 1 : row + (Dx - 1 - stag_x);
    case 1: return lx < Dx - 1 ? row + 1 : row - (Dx - 1 - stag_x);
    case 2: return ly > 0 ? row - Dx : row + Dx * (Dy - 1 - stag_y);
    default: return ly < Dy - 1 ? row + Dx : row - Dx * (Dy - 1 - stag_y);
    }
}

/* row_ptr: two 0-based arrays back to back (n_u+1, n_v+1); col_ind: 0-based per component,
 * u block then v block.  Follows calcCsrRowPtrGpu + the colInd stores of calcAdvetionMatrixX/Y. */
int orc_csr_structure(int ny, int nx, int per_x, int per_y, int *row_ptr, int *col_ind) {
    int n[2], nnz[2];
    if (orc_sizes(ny, nx, per_x, per_y, n, nnz)) return ORC_EBADGRID;
    int rp_off = 0, ci_off = 0;
    for (int c = 0; c < 2; c++) {
        int Dx, Dy; comp_dims(ny, nx, c, &Dx, &Dy);
        int *rp = row_ptr + rp_off, *ci = col_ind + ci_off;
        rp[0] = 0;
        for (int ly = 0; ly < Dy; ly++)
            for (int lx = 0; lx < Dx; lx++) {
                int row = lx + Dx * ly;
                rp[row + 1] = rowptr_closed_form(lx, ly, Dx, Dy, per_x, per_y);
            }
        for (int ly = 0; ly < Dy; ly++)
            for (int lx = 0; lx < Dx; lx++) {
                int row = lx + Dx * ly, slot[5], has[4];
                row_slots(lx, ly, Dx, Dy, per_x, per_y, slot, has);
                for (int k = 0; k < 4; k++)
                    if (has[k]) ci[rp[row] + slot[k]] = nb_col(k, row, lx, ly, Dx, Dy, c == 0, c == 1);
                ci[rp[row] + slot[4]] = row;
            }
        rp_off += n[c] + 1;
        ci_off += nnz[c];
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * velocity padding -- diffpiso/piso_helpers.py:35-55 (custom_padded, width 1)
 * up: (ny+2) x (nx+3), vp: (ny+3) x (nx+2)
 * ---------------------------------------------------------------------------------------- */
static int wrap(int i, int n) { i %= n; return i < 0 ? i + n : i; }
static int clampi(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

void orc_pad_velocity(int ny, int nx, int per_x, int per_y, const float *u, const float *v, float *up, float *vp) {
    int wu = nx + 3, wv = nx + 2;
    for (int i = 0; i < ny + 2; i++) {
        int sy = per_y ? wrap(i - 1, ny) : clampi(i - 1, 0, ny - 1);
        for (int j = 0; j < wu; j++) {
            /* own axis, periodic: duplicated last face dropped, then padded (1,2) (":47-50") */
            int sx = per_x ? wrap(j - 1, nx) : clampi(j - 1, 0, nx);
            up[i * wu + j] = u[sy * (nx + 1) + sx];
        }
    }
    for (int i = 0; i < ny + 3; i++) {
        int sy = per_y ? wrap(i - 1, ny) : clampi(i - 1, 0, ny);
        for (int j = 0; j < wv; j++) {
            int sx = per_x ? wrap(j - 1, nx) : clampi(j - 1, 0, nx - 1);
            vp[i * wv + j] = v[sy * nx + sx];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * advection-diffusion matrix values -- CUDAsrc/central_difference_csr_op.cu.cc:148-453
 * ---------------------------------------------------------------------------------------- */
int orc_assemble(int ny, int nx, int per_x, int per_y, float dy, float dx, float area_x, float area_y, float beta,
                 const float *up, const float *vp, const uint8_t *dirichlet /* n_u+n_v */,
                 const float *active /* (ny+2)(nx+2) */, const uint8_t *noslip /* (ny+2)(nx+2) */,
                 const float *visc, int visc_is_field,
                 const int *row_ptr, float *values, float *a_diag) {
    int n[2], nnz[2];
    if (orc_sizes(ny, nx, per_x, per_y, n, nnz)) return ORC_EBADGRID;
    const int per[2] = { per_x, per_y };
    const float cell_area[2] = { area_x, area_y };   /* piso_tf.py:97: prod(dx)/(dx, dy) as fp32, ~ (dy, dx) */
    const float spacing[2] = { dx, dy };       /* piso_tf.py:96  */
    const int wu = nx + 3, wv = nx + 2, wm = nx + 2;
    memset(values, 0, sizeof(float) * (size_t)(nnz[0] + nnz[1]));  /* initWithZeros, ":627" */
    int rp_off = 0, val_off = 0, row_off = 0;
    for (int c = 0; c < 2; c++) {
        int Dx, Dy; comp_dims(ny, nx, c, &Dx, &Dy);
        const int *rp = row_ptr + rp_off;
        float *val = values + val_off;
        for (int ly = 0; ly < Dy; ly++)
            for (int lx = 0; lx < Dx; lx++) {
                int row = lx + Dx * ly, slot[5], has[4];
                row_slots(lx, ly, Dx, Dy, per_x, per_y, slot, has);
                int reg[4] = { lx > 0, lx < Dx - 1, ly > 0, ly < Dy - 1 };
                if (dirichlet[row_off + row]) {            /* ":214-238" */
                    val[rp[row] + slot[4]] = 1.0f;
                    a_diag[row_off + row] = 0.0f;
                    continue;
                }
                /* face fluxes F[k], k = x-, x+, y-, y+  (calcCellFluxesX/Y, ":35-101") */
                float F[4];
                if (c == 0) {
                    const float *a = up + (ly + 1) * wu + (lx + 1);
                    F[0] = (float)(.5 * (a[0] + a[-1]) * cell_area[0]);
                    F[1] = (float)(.5 * (a[1] + a[0]) * cell_area[0]);
                    const float *b = vp + (ly + 1) * wv + (lx + 1);
                    F[2] = (float)(.5 * (b[0] + b[-1]) * cell_area[1]);
                    F[3] = (float)(.5 * (b[wv] + b[wv - 1]) * cell_area[1]);
                } else {
                    const float *a = up + (ly + 1) * wu + (lx + 1);
                    F[0] = (float)(.5 * (a[0] + a[-wu]) * cell_area[0]);
                    F[1] = (float)(.5 * (a[1] + a[1 - wu]) * cell_area[0]);
                    const float *b = vp + (ly + 1) * wv + (lx + 1);
                    F[2] = (float)(.5 * (b[0] + b[-wv]) * cell_area[1]);
                    F[3] = (float)(.5 * (b[wv] + b[0]) * cell_area[1]);
                }
                /* padded-centred mask cell consulted per direction (gridIDXpaddedCenteredMasks, ":132-146") */
                int m[4];
                m[0] = (ly + 1) * wm + lx;
                m[1] = (ly + 1) * wm + lx + 1 + (c == 1);
                m[2] = ly * wm + lx + 1;
                m[3] = (ly + 1 + (c == 0)) * wm + lx + 1;
                float nu = visc[visc_is_field ? row_off + row : 0];
                float diag = 0.0f;
                for (int d = 1; d >= 0; d--) {            /* y first, then x (":248") */
                    float D = nu * cell_area[d] / spacing[d];
                    int not_stag = (d != c);
                    for (int side = 0; side < 2; side++) {
                        int k = 2 * d + side;
                        int ns = noslip[m[k]] ? 1 : 0;
                        int t = (active[m[k]] == 1.0f) || (reg[k] && ns);
                        float sgn_flux = side == 0 ? F[k] : -F[k];
                        if (t && has[k])
                            val[rp[row] + slot[k]] = (float)(sgn_flux * .5 + D);
                        int kfac = t + not_stag * (1 - t) * ns * 2;
                        diag = (float)(diag + ((sgn_flux * (float)(2 - t)) * .5 - D * (float)kfac));
                    }
                }
                (void)per;
                val[rp[row] + slot[4]] = diag - beta;       /* ":294" */
                a_diag[row_off + row] = diag;
            }
        rp_off += n[c] + 1;
        val_off += nnz[c];
        row_off += n[c];
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * finite-volume pressure gradient on faces -- diffpiso/piso_helpers.py:236-274
 * pbc[4] = pressure extrapolation at y_lo, y_hi, x_lo, x_hi: 0 replicate ('boundary'),
 * 1 zero ('constant'), 2 periodic.  access = accessible mask (ny+2)(nx+2).  Output flat [u, v].
 * ---------------------------------------------------------------------------------------- */
static float p_ghost(const float *p, int ny, int nx, int cy, int cx, const int *pbc) {
    if (cy < 0) { if (pbc[0] == 1) return 0.0f; cy = pbc[0] == 2 ? ny - 1 : 0; }
    if (cy >= ny) { if (pbc[1] == 1) return 0.0f; cy = pbc[1] == 2 ? 0 : ny - 1; }
    if (cx < 0) { if (pbc[2] == 1) return 0.0f; cx = pbc[2] == 2 ? nx - 1 : 0; }
    if (cx >= nx) { if (pbc[3] == 1) return 0.0f; cx = pbc[3] == 2 ? 0 : nx - 1; }
    return p[cy * nx + cx];
}

void orc_fv_gradient(int ny, int nx, float dy, float dx, const int *pbc, const float *access,
                     const float *p, float *g) {
    const float prod = (float)((double)dy * (double)dx);
    const int wm = nx + 2;
    float *gu = g, *gv = g + ny * (nx + 1);
    for (int cy = 0; cy < ny; cy++)
        for (int i = 0; i <= nx; i++) {
            float diff = p_ghost(p, ny, nx, cy, i, pbc) - p_ghost(p, ny, nx, cy, i - 1, pbc);
            float mk = fminf(access[(cy + 1) * wm + i], access[(cy + 1) * wm + i + 1]);
            gu[cy * (nx + 1) + i] = ((diff * prod) / dx) * mk;
        }
    for (int j = 0; j <= ny; j++)
        for (int cx = 0; cx < nx; cx++) {
            float diff = p_ghost(p, ny, nx, j, cx, pbc) - p_ghost(p, ny, nx, j - 1, cx, pbc);
            float mk = fminf(access[j * wm + cx + 1], access[(j + 1) * wm + cx + 1]);
            gv[j * nx + cx] = ((diff * prod) / dy) * mk;
        }
}

/* finite-volume divergence -- diffpiso/piso_helpers.py:285-289 */
void orc_fv_divergence(int ny, int nx, float dy, float dx, const float *vel, float *div) {
    const float prod = (float)((double)dy * (double)dx);
    const float *u = vel, *v = vel + ny * (nx + 1);
    for (int cy = 0; cy < ny; cy++)
        for (int cx = 0; cx < nx; cx++) {
            float ty = ((v[(cy + 1) * nx + cx] - v[cy * nx + cx]) * prod) / dy;
            float tx = ((u[cy * (nx + 1) + cx + 1] - u[cy * (nx + 1) + cx]) * prod) / dx;
            div[cy * nx + cx] = ty + tx;   /* math.sum over [y-term, x-term] */
        }
}

/* ------------------------------------------------------------------------------------------
 * generic CSR kernels used by the predictor (ILU0, triangular solves, SpMV, transpose)
 * ---------------------------------------------------------------------------------------- */
static double dotd(int n, const float *a, const float *b) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += (double)a[i] * (double)b[i];
    return s;
}
static float dotf(int n, const float *a, const float *b) { return (float)dotd(n, a, b); }
static float nrm2f(int n, const float *a) { return (float)sqrt(dotd(n, a, a)); }

void orc_spmv_f32(int n, const int *rp, const int *ci, const float *val, const float *x, float *y) {
    for (int i = 0; i < n; i++) {
        float acc = 0.0f;
        for (int k = rp[i]; k < rp[i + 1]; k++) acc = fmaf(val[k], x[ci[k]], acc);
        y[i] = acc;
    }
}

/* csr2csc (multi_bicgstab...cu.cc:117-119): transpose with ascending columns per row */
void orc_csr_transpose_f32(int n, const int *rp, const int *ci, const float *val, int *trp, int *tci, float *tval) {
    int nnz = rp[n];
    memset(trp, 0, sizeof(int) * (size_t)(n + 1));
    for (int k = 0; k < nnz; k++) trp[ci[k] + 1]++;
    for (int i = 0; i < n; i++) trp[i + 1] += trp[i];
    int *fill = (int *)malloc(sizeof(int) * (size_t)n);
    memcpy(fill, trp, sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++)
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            int q = fill[ci[k]]++;
            tci[q] = i;
            tval[q] = val[k];
        }
    free(fill);
}

/* ILU(0), IKJ on the CSR pattern, unit lower factor stored in the strict lower part, no
 * pivoting (what cusparseScsrilu02 computes; multi_bicgstab...cu.cc:181-218).  Returns the index
 * of the first zero pivot or -1. */
int orc_ilu0_f32(int n, const int *rp, const int *ci, const float *val, float *lu) {
    int zero_pivot = -1;
    memcpy(lu, val, sizeof(float) * (size_t)rp[n]);
    int *dpos = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) {
        dpos[i] = -1;
        for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) dpos[i] = k;
    }
    for (int i = 0; i < n; i++) {
        for (int kk = rp[i]; kk < rp[i + 1] && ci[kk] < i; kk++) {
            int k = ci[kk];
            float lik = lu[kk] / lu[dpos[k]];
            lu[kk] = lik;
            for (int jj = kk + 1; jj < rp[i + 1]; jj++) {
                int j = ci[jj];
                for (int mm = dpos[k] + 1; mm < rp[k + 1]; mm++)
                    if (ci[mm] == j) { lu[jj] = fmaf(-lik, lu[mm], lu[jj]); break; }
            }
        }
        if (lu[dpos[i]] == 0.0f && zero_pivot < 0) zero_pivot = i;
    }
    free(dpos);
    return zero_pivot;
}

/* unit-lower solve L y = b, then upper solve U x = y (csrsv2, ":321-327") */
void orc_lu_solve_f32(int n, const int *rp, const int *ci, const float *lu, const float *b, float *y, float *x) {
    for (int i = 0; i < n; i++) {
        float acc = b[i];
        for (int k = rp[i]; k < rp[i + 1] && ci[k] < i; k++) acc = fmaf(-lu[k], y[ci[k]], acc);
        y[i] = acc;
    }
    for (int i = n - 1; i >= 0; i--) {
        float acc = y[i], d = 1.0f;
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            if (ci[k] == i) d = lu[k];
            else if (ci[k] > i) acc = fmaf(-lu[k], x[ci[k]], acc);
        }
        x[i] = acc / d;
    }
}

/* ------------------------------------------------------------------------------------------
 * ILU0-preconditioned BiCGStab -- CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:233-411
 * One component.  stats[0] = iterations (it_count), stats[1] = restarts used, stats[2] = warn,
 * stats[3] = exit kind (0 lucky guess, 1 first test, 2 second test, 3 max_it)
 * ---------------------------------------------------------------------------------------- */
int orc_bicgstab_ilu_f32(int n, const int *rp_in, const int *ci_in, const float *val_in,
                         const float *rhs, const float *x0, float tol, int max_it, int transpose,
                         float *x, int *stats, float *final_res) {
    int nnz = rp_in[n];
    const int *rp = rp_in, *ci = ci_in; const float *val = val_in;
    int *trp = NULL, *tci = NULL; float *tval = NULL;
    if (transpose) {                                   /* ":113-134" */
        trp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
        tci = (int *)malloc(sizeof(int) * (size_t)nnz);
        tval = (float *)malloc(sizeof(float) * (size_t)nnz);
        orc_csr_transpose_f32(n, rp_in, ci_in, val_in, trp, tci, tval);
        rp = trp; ci = tci; val = tval;
    }
    float *lu = (float *)malloc(sizeof(float) * (size_t)nnz);
    float *w = (float *)calloc((size_t)n * 8, sizeof(float));
    float *r = w, *rh = w + n, *p = w + 2 * n, *v = w + 3 * n, *t = w + 4 * n, *z = w + 5 * n,
          *ph = w + 6 * n, *sh = w + 7 * n;
    orc_ilu0_f32(n, rp, ci, val, lu);

    float alpha = 1.f, rho = 1.f, rhop = 1.f, omega = 1.f, beta, nrm_r = 0.f;
    int it_count = 0, restarts = 0, exit_kind = 3;
    int warn = isnan(nrm2f(nnz, val)) || isnan(nrm2f(n, rhs));     /* ":245-256" */
    memcpy(x, x0, sizeof(float) * (size_t)n);                          /* ":261" */
    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        orc_spmv_f32(n, rp, ci, val, x, r);                            /* r = b - A x, ":275-282" */
        for (int i = 0; i < n; i++) r[i] = rhs[i] - r[i];
        nrm_r = nrm2f(n, r);
        if (nrm_r < tol) { exit_kind = 0; break; }                     /* ":287-289" */
        memcpy(rh, r, sizeof(float) * (size_t)n);
        memset(p, 0, sizeof(float) * (size_t)n);
        memset(v, 0, sizeof(float) * (size_t)n);
        exit_kind = 3;
        for (int it = 0; it < max_it; it++) {
            it_count++;
            rhop = rho;
            rho = dotf(n, r, rh);
            beta = (rho / rhop) * (alpha / omega);
            for (int i = 0; i < n; i++) {                              /* ":315-317" */
                float pi = fmaf(-omega, v[i], p[i]);
                pi = beta * pi;
                p[i] = pi + r[i];
            }
            orc_lu_solve_f32(n, rp, ci, lu, p, z, ph);
            orc_spmv_f32(n, rp, ci, val, ph, v);
            alpha = rho / dotf(n, rh, v);
            for (int i = 0; i < n; i++) { x[i] = fmaf(alpha, ph[i], x[i]); r[i] = fmaf(-alpha, v[i], r[i]); }
            nrm_r = nrm2f(n, r);
            if (nrm_r < tol) { exit_kind = 1; break; }
            orc_lu_solve_f32(n, rp, ci, lu, r, z, sh);
            orc_spmv_f32(n, rp, ci, val, sh, t);
            omega = dotf(n, t, r) / dotf(n, t, t);
            for (int i = 0; i < n; i++) { x[i] = fmaf(omega, sh[i], x[i]); r[i] = fmaf(-omega, t[i], r[i]); }
            nrm_r = nrm2f(n, r);
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > tol * 100 || isnan(nrm_r)) {                       /* ":392-404" */
            memset(x, 0, sizeof(float) * (size_t)n);
            if (restart == 1) restarts = 2;
        } else break;
    }
    stats[0] = it_count; stats[1] = restarts; stats[2] = warn; stats[3] = exit_kind;
    if (final_res) *final_res = nrm_r;
    free(lu); free(w); free(trp); free(tci); free(tval);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * The fp64 variant of the predictor solve -- CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:540-988, selected by
 * LinearSolverCudaMultiBicgstabILU(cast_to_double=True) (diffpiso/linear_solver.py:130-133, which casts the fp32
 * matrix values and right-hand side to fp64 and the solution back to fp32, ":171").  Same sequence as the fp32
 * launcher with cusparseD / cublasD calls; the tolerance stays a float.  Inputs and output are fp32 here as well.
 * ---------------------------------------------------------------------------------------- */
static double dotdd(int n, const double *a, const double *b) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
static void spmv_f64(int n, const int *rp, const int *ci, const double *val, const double *x, double *y) {
    for (int i = 0; i < n; i++) {
        double acc = 0.0;
        for (int k = rp[i]; k < rp[i + 1]; k++) acc = fma(val[k], x[ci[k]], acc);
        y[i] = acc;
    }
}
static void ilu0_f64(int n, const int *rp, const int *ci, const double *val, double *lu) {
    memcpy(lu, val, sizeof(double) * (size_t)rp[n]);
    int *dpos = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) {
        dpos[i] = -1;
        for (int k = rp[i]; k < rp[i + 1]; k++) if (ci[k] == i) dpos[i] = k;
    }
    for (int i = 0; i < n; i++)
        for (int kk = rp[i]; kk < rp[i + 1] && ci[kk] < i; kk++) {
            int k = ci[kk];
            double lik = lu[kk] / lu[dpos[k]];
            lu[kk] = lik;
            for (int jj = kk + 1; jj < rp[i + 1]; jj++) {
                int j = ci[jj];
                for (int mm = dpos[k] + 1; mm < rp[k + 1]; mm++)
                    if (ci[mm] == j) { lu[jj] = fma(-lik, lu[mm], lu[jj]); break; }
            }
        }
    free(dpos);
}
static void lu_solve_f64(int n, const int *rp, const int *ci, const double *lu, const double *b, double *y, double *x) {
    for (int i = 0; i < n; i++) {
        double acc = b[i];
        for (int k = rp[i]; k < rp[i + 1] && ci[k] < i; k++) acc = fma(-lu[k], y[ci[k]], acc);
        y[i] = acc;
    }
    for (int i = n - 1; i >= 0; i--) {
        double acc = y[i], d = 1.0;
        for (int k = rp[i]; k < rp[i + 1]; k++) {
            if (ci[k] == i) d = lu[k];
            else if (ci[k] > i) acc = fma(-lu[k], x[ci[k]], acc);
        }
        x[i] = acc / d;
    }
}
int orc_bicgstab_ilu_f64(int n, const int *rp_in, const int *ci_in, const float *val_in,
                         const float *rhs_in, const float *x0, float tol_f, int max_it, int transpose,
                         float *x_out, int *stats, float *final_res) {
    int nnz = rp_in[n];
    const int *rp = rp_in, *ci = ci_in;
    int *trp = NULL, *tci = NULL; float *tvalf = NULL;
    const float *valf = val_in;
    if (transpose) {
        trp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
        tci = (int *)malloc(sizeof(int) * (size_t)nnz);
        tvalf = (float *)malloc(sizeof(float) * (size_t)nnz);
        orc_csr_transpose_f32(n, rp_in, ci_in, val_in, trp, tci, tvalf);
        rp = trp; ci = tci; valf = tvalf;
    }
    double *val = (double *)malloc(sizeof(double) * (size_t)nnz), *lu = (double *)malloc(sizeof(double) * (size_t)nnz);
    for (int k = 0; k < nnz; k++) val[k] = (double)valf[k];                /* tf.cast(matrix_values, tf.float64) */
    double *w = (double *)calloc((size_t)n * 10, sizeof(double));
    double *r = w, *rh = w + n, *p = w + 2 * n, *v = w + 3 * n, *t = w + 4 * n, *z = w + 5 * n,
           *ph = w + 6 * n, *sh = w + 7 * n, *x = w + 8 * n, *rhs = w + 9 * n;
    for (int i = 0; i < n; i++) { rhs[i] = (double)rhs_in[i]; x[i] = (double)x0[i]; }
    ilu0_f64(n, rp, ci, val, lu);
    const double tol = (double)tol_f;
    double alpha = 1., rho = 1., rhop = 1., omega = 1., beta, nrm_r = 0.;
    int it_count = 0, restarts = 0, exit_kind = 3;
    int warn = isnan(sqrt(dotdd(nnz, val, val))) || isnan(sqrt(dotdd(n, rhs, rhs)));
    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        spmv_f64(n, rp, ci, val, x, r);
        for (int i = 0; i < n; i++) r[i] = rhs[i] - r[i];
        nrm_r = sqrt(dotdd(n, r, r));
        if (nrm_r < tol) { exit_kind = 0; break; }
        memcpy(rh, r, sizeof(double) * (size_t)n);
        memset(p, 0, sizeof(double) * (size_t)n);
        memset(v, 0, sizeof(double) * (size_t)n);
        exit_kind = 3;
        for (int it = 0; it < max_it; it++) {
            it_count++;
            rhop = rho;
            rho = dotdd(n, r, rh);
            beta = (rho / rhop) * (alpha / omega);
            for (int i = 0; i < n; i++) {
                double pi = fma(-omega, v[i], p[i]);
                pi = beta * pi;
                p[i] = pi + r[i];
            }
            lu_solve_f64(n, rp, ci, lu, p, z, ph);
            spmv_f64(n, rp, ci, val, ph, v);
            alpha = rho / dotdd(n, rh, v);
            for (int i = 0; i < n; i++) { x[i] = fma(alpha, ph[i], x[i]); r[i] = fma(-alpha, v[i], r[i]); }
            nrm_r = sqrt(dotdd(n, r, r));
            if (nrm_r < tol) { exit_kind = 1; break; }
            lu_solve_f64(n, rp, ci, lu, r, z, sh);
            spmv_f64(n, rp, ci, val, sh, t);
            omega = dotdd(n, t, r) / dotdd(n, t, t);
            for (int i = 0; i < n; i++) { x[i] = fma(omega, sh[i], x[i]); r[i] = fma(-omega, t[i], r[i]); }
            nrm_r = sqrt(dotdd(n, r, r));
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > tol * 100 || isnan(nrm_r)) {
            memset(x, 0, sizeof(double) * (size_t)n);
            if (restart == 1) restarts = 2;
        } else break;
    }
    for (int i = 0; i < n; i++) x_out[i] = (float)x[i];                    /* tf.cast(sol[3], tf.float32) */
    stats[0] = it_count; stats[1] = restarts; stats[2] = warn; stats[3] = exit_kind;
    if (final_res) *final_res = (float)nrm_r;
    free(val); free(lu); free(w); free(trp); free(tci); free(tvalf);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * PISO pressure matrix -- CUDAsrc/laplace_op.cu.cc:79-179 (calcPISOLaplaceMatrix)
 * k_faces is flattened [v, u] (piso_cuda_pressure_solver.py:70); output 5 per cell
 * [y-, x-, diag, x+, y+].
 * ---------------------------------------------------------------------------------------- */
#define LAPLACE_IMPL(NAME, T)                                                                       \
void NAME(int ny, int nx, const float *active, const float *fluid, const float *k_faces, T *lap) { \
    const int wm = nx + 2;                                                                          \
    const float *kv = k_faces, *ku = k_faces + (ny + 1) * nx;                                       \
    for (int cy = 0; cy < ny; cy++)                                                                 \
        for (int cx = 0; cx < nx; cx++) {                                                           \
            int row = cy * nx + cx, me = (cy + 1) * wm + cx + 1;                                    \
            int mnb[4] = { me - wm, me - 1, me + 1, me + wm };           /* y-, x-, x+, y+ */      \
            float kf[4] = { kv[cy * nx + cx], ku[cy * (nx + 1) + cx],                               \
                            ku[cy * (nx + 1) + cx + 1], kv[(cy + 1) * nx + cx] };                   \
            int self_active = active[me] != 0.0f;                                                   \
            int self_solid = (active[me] == 0.0f && fluid[me] == 0.0f);                             \
            T diag = 0;                                                                             \
            /* diagonal accumulation order: y-, y+, x-, x+ (":118-135") */                         \
            const int order[4] = { 0, 3, 1, 2 };                                                    \
            for (int q = 0; q < 4; q++) {                                                           \
                int k = order[q];                                                                   \
                int nb_solid = (active[mnb[k]] == 0.0f && fluid[mnb[k]] == 0.0f);                   \
                if (!nb_solid && self_active) diag -= kf[k];                                        \
            }                                                                                       \
            T off[4];                                                                               \
            for (int k = 0; k < 4; k++) {                                                           \
                int nb_fluid = (active[mnb[k]] == 1.0f && fluid[mnb[k]] == 1.0f);                   \
                off[k] = (nb_fluid && !self_solid) ? (T)kf[k] : (T)0;                               \
            }                                                                                       \
            T *o = lap + 5 * (size_t)row;                                                           \
            o[0] = off[0]; o[1] = off[1]; o[2] = diag; o[3] = off[2]; o[4] = off[3];                \
        }                                                                                           \
}
LAPLACE_IMPL(orc_laplace_f64, double)
LAPLACE_IMPL(orc_laplace_f32, float)

/* ------------------------------------------------------------------------------------------
 * pressure CG -- CUDAsrc/pressure_solve_op.cu.cc:57-133 (calcZ_v4, offsets) and 421-696 / 140-418
 * (LaunchPressureKernel), batch of one, init_with_zeros = true, randomized_restarts = 0.
 * ---------------------------------------------------------------------------------------- */
#define CG_IMPL(NAME, T, FMA, FABS)                                                                 \
static void NAME##_apply(int ny, int nx, int per_x, int per_y, const T *lap, const T *p, T shift, T *z) { \
    int n = ny * nx;                                                                                \
    for (int cy = 0; cy < ny; cy++)                                                                 \
        for (int cx = 0; cx < nx; cx++) {                                                           \
            int row = cy * nx + cx;                                                                 \
            /* neighbour rows with the periodic wrap offsets of calcDiagonalOffsets (":117-133") */ \
            int nb[5];                                                                              \
            nb[0] = row - nx + ((cy == 0) ? n * per_y : 0);                                         \
            nb[1] = row - 1 + ((cx == 0) ? nx * per_x : 0);                                         \
            nb[2] = row;                                                                            \
            nb[3] = row + 1 - ((cx == nx - 1) ? nx * per_x : 0);                                    \
            nb[4] = row + nx - ((cy == ny - 1) ? n * per_y : 0);                                    \
            const T *l = lap + 5 * (size_t)row;                                                     \
            T acc = 0;                                                                              \
            for (int k = 0; k < 5; k++)                                                             \
                if (l[k] != 0) acc = FMA(l[k], p[nb[k]], acc);       /* zero coefficients read p[0] (":88") */ \
            z[row] = acc + shift;                                                                   \
        }                                                                                           \
}                                                                                                   \
int NAME(int ny, int nx, int per_x, int per_y, const T *lap, const T *div, float accuracy,          \
         int max_it, int residual_reset, int rank_deficient, T *x, int *iterations) {               \
    int n = ny * nx;                                                                                \
    T *p = (T *)malloc(sizeof(T) * (size_t)n * 3), *r = p + n, *z = p + 2 * n;                      \
    T scale = 0;                                                                                    \
    if (rank_deficient) {                                            /* ":444-450" */              \
        for (int i = 0; i < n; i++) scale += FABS(lap[5 * (size_t)i + 2]);                          \
        scale *= .1 / n;                                                                            \
    }                                                                                               \
    for (int i = 0; i < n; i++) x[i] = 0;                            /* init_with_zeros */          \
    T sum = 0;                                                                                      \
    NAME##_apply(ny, nx, per_x, per_y, lap, x, scale * sum, z);                                     \
    for (int i = 0; i < n; i++) p[i] = r[i] = div[i] - z[i];         /* initVariablesWithGuess */   \
    int flag = 0, checker = 1, it = 0;                                                              \
    for (; it < max_it; it++) {                                                                     \
        if ((it + 1) % residual_reset == 0) {                        /* ":539-553" */              \
            sum = 0; for (int i = 0; i < n; i++) sum += x[i];                                       \
            NAME##_apply(ny, nx, per_x, per_y, lap, x, rank_deficient ? scale * sum : 0, z);        \
            for (int i = 0; i < n; i++) p[i] = r[i] = div[i] - z[i];                                \
            flag = 0;                                                                               \
        }                                                                                           \
        sum = 0; if (rank_deficient) for (int i = 0; i < n; i++) sum += p[i];                       \
        NAME##_apply(ny, nx, per_x, per_y, lap, p, rank_deficient ? scale * sum : 0, z);            \
        T p_r = 0, p_z = 0;                                                                         \
        for (int i = 0; i < n; i++) { p_r += p[i] * r[i]; p_z += p[i] * z[i]; }                     \
        T alpha = 0;                                                                                \
        if (FABS(p_z) > 0) alpha = p_r / p_z;                                                       \
        for (int i = 0; i < n; i++) { x[i] = FMA(alpha, p[i], x[i]); r[i] = FMA(-alpha, z[i], r[i]); } \
        if (checker % 5 == 0) {                                      /* ":591-614" */              \
            for (int i = 0; i < n; i++) if (FABS(r[i]) >= accuracy) { flag = 0; break; }            \
            if (flag) { it++; break; }                                                              \
            flag = 1;                                                                               \
        }                                                                                           \
        checker++;                                                                                  \
        T r_z = 0;                                                                                  \
        for (int i = 0; i < n; i++) r_z += r[i] * z[i];                                             \
        T beta = -r_z / p_z;                                                                        \
        for (int i = 0; i < n; i++) { T bp = beta * p[i]; p[i] = bp + r[i]; }                       \
    }                                                                                               \
    *iterations = it;                                                                               \
    free(p);                                                                                        \
    return ORC_OK;                                                                                  \
}
CG_IMPL(orc_pressure_cg_f64, double, fma, fabs)
CG_IMPL(orc_pressure_cg_f32, float, fmaf, fabsf)

/* ------------------------------------------------------------------------------------------
 * explicit H product -- diffpiso/piso_helpers.py:209-223: H d = M d - (A - beta) d
 * (un-negated values; one component)
 * ---------------------------------------------------------------------------------------- */
void orc_h_apply(int n, const int *rp, const int *ci, const float *val, const float *a_diag,
                 float beta, const float *d, float *h) {
    for (int i = 0; i < n; i++) {
        float acc = 0.0f;
        for (int k = rp[i]; k < rp[i + 1]; k++) acc += d[ci[k]] * val[k];     /* gather * values, segment_sum */
        h[i] = acc - (a_diag[i] - beta) * d[i];
    }
}

/* ------------------------------------------------------------------------------------------
 * one forward PISO step -- diffpiso/piso_tf.py:11-81.  Single sample.
 *
 * iparams: [0]=ny [1]=nx [2]=per_y [3]=per_x [4..7]=pbc of pressure (y_lo,y_hi,x_lo,x_hi)
 *          [8..11]=pbc of the pressure increments [12]=visc_is_field [13]=bicg max_it
 *          [14]=cg max_it [15]=cg residual_reset [16]=rank_deficient [17]=cg fp64 (1) / fp32 (0)
 *          [18],[19]=velocity padding periodic in y, x (1 = as the domain; 0 = replicate although the domain is periodic)
 * fparams: [0]=dy [1]=dx [2]=dt [3]=bicg tol [4]=cg accuracy [5]=beta [6]=unused [7]=dx_factor [8]=cell_area x [9]=cell_area y
 *          ([5..7] are the fp32 graph constants the Python side forms in fp64, piso_tf.py:26,53; 0 = derive here)
 * out_stats (int[12]): [0..3] u solve stats, [4..7] v solve stats, [8] cg1 its, [9] cg2 its
 * optional outputs may be NULL.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float *values, *a_diag, *rhs, *u_star, *div1, *p1, *u_s2, *h, *div2, *p2;   /* all optional */
    double *lap64; float *lap32;
} orc_step_extra;

int orc_piso_step(const int *ip, const float *fp, const float *vel, const float *pres,
                  const uint8_t *dirichlet, const float *dirichlet_values,
                  const float *active, const float *access, const uint8_t *noslip,
                  const float *visc, const float *forcing /* may be NULL */,
                  float *vel_out, float *pres_out, int *out_stats, orc_step_extra *ex) {
    const int ny = ip[0], nx = ip[1], per_y = ip[2], per_x = ip[3];
    const int *pbc = ip + 4, *pbc_inc = ip + 8;
    const float dy = fp[0], dx = fp[1], dt = fp[2];
    int n[2], nnz[2];
    if (orc_sizes(ny, nx, per_x, per_y, n, nnz)) return ORC_EBADGRID;
    const int nf = n[0] + n[1], nc = ny * nx, nz = nnz[0] + nnz[1];
    const double prod_d = (double)dy * (double)dx;
    const float prod = (float)prod_d;    /* as inside orc_fv_gradient / orc_fv_divergence */
    const float beta = fp[5] != 0.0f ? fp[5] : (float)(prod_d / (double)dt);  /* piso_tf.py:26 */
    const float dx_factor = fp[5] != 0.0f ? fp[7] : (float)(prod_d / ((double)dy * (double)dy)); /* piso_tf.py:53 (dx[0] = dy) */

    int *row_ptr = (int *)malloc(sizeof(int) * (size_t)(nf + 2));
    int *col_ind = (int *)malloc(sizeof(int) * (size_t)nz);
    float *values = (float *)malloc(sizeof(float) * (size_t)nz);
    float *neg = (float *)malloc(sizeof(float) * (size_t)nz);
    float *wf = (float *)calloc((size_t)nf * 8, sizeof(float));
    float *a_diag = wf, *g = wf + nf, *rhs = wf + 2 * nf, *ustar = wf + 3 * nf, *kf = wf + 4 * nf,
          *us2 = wf + 5 * nf, *h = wf + 6 * nf, *tmp = wf + 7 * nf;
    float *wc = (float *)calloc((size_t)nc * 4, sizeof(float));
    float *div = wc, *p1 = wc + nc, *p2 = wc + 2 * nc;
    float *up = (float *)malloc(sizeof(float) * (size_t)(ny + 2) * (nx + 3));
    float *vp = (float *)malloc(sizeof(float) * (size_t)(ny + 3) * (nx + 2));

    /* advection matrices (piso_tf.py:29-33) */
    orc_csr_structure(ny, nx, per_x, per_y, row_ptr, col_ind);
    /* ip[18], ip[19]: the velocity GRID's periodic flags (y, x) as custom_padded sees them; they differ from the domain's from the
     * second step of run_piso_steps on, where the re-wrapped state has the default 'boundary' extrapolation
     * (combined_training_integrated.py:431-432,473-474: the extrapolation is passed in the position of `name`) */
    orc_pad_velocity(ny, nx, per_x && ip[19], per_y && ip[18], vel, vel + n[0], up, vp);
    orc_assemble(ny, nx, per_x, per_y, dy, dx, fp[8] != 0.0f ? fp[8] : dy, fp[8] != 0.0f ? fp[9] : dx, beta, up, vp,
                 dirichlet, active, noslip, visc, ip[12],
                 row_ptr, values, a_diag);

    /* predictor rhs (piso_tf.py:36-40, piso_helpers.py:169-172) */
    orc_fv_gradient(ny, nx, dy, dx, pbc, access, pres, g);
    for (int i = 0; i < nf; i++) {
        float v = vel[i] * beta - g[i];
        if (forcing) v += forcing[i] * prod;
        rhs[i] = dirichlet[i] ? dirichlet_values[i] * -1.0f : v;
    }
    /* predictor solve with -M (piso_tf.py:42-43) */
    for (int i = 0; i < nz; i++) neg[i] = -values[i];
    orc_bicgstab_ilu_f32(n[0], row_ptr, col_ind, neg, rhs, vel, fp[3], ip[13], 0, ustar, out_stats, NULL);
    orc_bicgstab_ilu_f32(n[1], row_ptr + n[0] + 1, col_ind + nnz[0], neg + nnz[0], rhs + n[0], vel + n[0],
                         fp[3], ip[13], 0, ustar + n[0], out_stats + 4, NULL);

    /* corrector 1 (piso_tf.py:51-58) */
    orc_fv_divergence(ny, nx, dy, dx, ustar, div);
    /* scaling field 1/(beta-A)*dx_factor flattened [v,u] (piso_cuda_pressure_solver.py:70) */
    for (int i = 0; i < n[1]; i++) kf[i] = (1.0f / (beta - a_diag[n[0] + i])) * dx_factor;
    for (int i = 0; i < n[0]; i++) kf[n[1] + i] = (1.0f / (beta - a_diag[i])) * dx_factor;
    void *lap = malloc((ip[17] ? sizeof(double) : sizeof(float)) * 5 * (size_t)nc);
    void *cgx = malloc((ip[17] ? sizeof(double) : sizeof(float)) * (size_t)nc * 2);
    if (ip[17]) {
        double *d64 = (double *)cgx + nc;
        orc_laplace_f64(ny, nx, active, access, kf, (double *)lap);
        for (int i = 0; i < nc; i++) d64[i] = div[i];
        orc_pressure_cg_f64(ny, nx, per_x, per_y, (double *)lap, d64, fp[4], ip[14], ip[15], ip[16], (double *)cgx, out_stats + 8);
        for (int i = 0; i < nc; i++) p1[i] = (float)((double *)cgx)[i];
    } else {
        orc_laplace_f32(ny, nx, active, access, kf, (float *)lap);
        orc_pressure_cg_f32(ny, nx, per_x, per_y, (float *)lap, div, fp[4], ip[14], ip[15], ip[16], (float *)cgx, out_stats + 8);
        for (int i = 0; i < nc; i++) p1[i] = ((float *)cgx)[i];
    }
    if (ex && ex->div1) memcpy(ex->div1, div, sizeof(float) * (size_t)nc);
    orc_fv_gradient(ny, nx, dy, dx, pbc_inc, access, p1, g);
    for (int i = 0; i < nf; i++) us2[i] = ustar[i] - (g[i] / (beta - a_diag[i])) / prod;

    /* corrector 2 (piso_tf.py:61-73) */
    for (int i = 0; i < nf; i++) tmp[i] = us2[i] - ustar[i];
    orc_h_apply(n[0], row_ptr, col_ind, values, a_diag, beta, tmp, h);
    orc_h_apply(n[1], row_ptr + n[0] + 1, col_ind + nnz[0], values + nnz[0], a_diag + n[0], beta, tmp + n[0], h + n[0]);
    for (int i = 0; i < nf; i++) tmp[i] = h[i] / (beta - a_diag[i]);
    orc_fv_divergence(ny, nx, dy, dx, tmp, div);
    if (ip[17]) {
        double *d64 = (double *)cgx + nc;
        for (int i = 0; i < nc; i++) d64[i] = div[i];
        orc_pressure_cg_f64(ny, nx, per_x, per_y, (double *)lap, d64, fp[4], ip[14], ip[15], ip[16], (double *)cgx, out_stats + 9);
        for (int i = 0; i < nc; i++) p2[i] = (float)((double *)cgx)[i];
    } else {
        orc_pressure_cg_f32(ny, nx, per_x, per_y, (float *)lap, div, fp[4], ip[14], ip[15], ip[16], (float *)cgx, out_stats + 9);
        for (int i = 0; i < nc; i++) p2[i] = ((float *)cgx)[i];
    }
    orc_fv_gradient(ny, nx, dy, dx, pbc_inc, access, p2, g);
    for (int i = 0; i < nf; i++) vel_out[i] = us2[i] + (h[i] - g[i] / prod) / (beta - a_diag[i]);
    for (int i = 0; i < nc; i++) pres_out[i] = (pres[i] + p1[i]) + p2[i];               /* piso_tf.py:75 */

    if (ex) {
        if (ex->values) memcpy(ex->values, values, sizeof(float) * (size_t)nz);
        if (ex->a_diag) memcpy(ex->a_diag, a_diag, sizeof(float) * (size_t)nf);
        if (ex->rhs) memcpy(ex->rhs, rhs, sizeof(float) * (size_t)nf);
        if (ex->u_star) memcpy(ex->u_star, ustar, sizeof(float) * (size_t)nf);
        if (ex->p1) memcpy(ex->p1, p1, sizeof(float) * (size_t)nc);
        if (ex->u_s2) memcpy(ex->u_s2, us2, sizeof(float) * (size_t)nf);
        if (ex->h) memcpy(ex->h, h, sizeof(float) * (size_t)nf);
        if (ex->div2) memcpy(ex->div2, div, sizeof(float) * (size_t)nc);
        if (ex->p2) memcpy(ex->p2, p2, sizeof(float) * (size_t)nc);
        if (ex->lap64 && ip[17]) memcpy(ex->lap64, lap, sizeof(double) * 5 * (size_t)nc);
        if (ex->lap32 && !ip[17]) memcpy(ex->lap32, lap, sizeof(float) * 5 * (size_t)nc);
    }
    free(row_ptr); free(col_ind); free(values); free(neg); free(wf); free(wc); free(up); free(vp);
    free(lap); free(cgx);
    return ORC_OK;
}
