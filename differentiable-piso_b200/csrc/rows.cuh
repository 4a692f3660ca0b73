// rows.cuh -- per-row / per-cell arithmetic of the PISO step, shared by the CUDA kernels (device) and by the
// host emulation used in the CPU test-suite (tests/host_shim.cpp).  Every function handles ONE row of ONE sample
// and reproduces the reference's floating-point sequence (float vs double promotion, no re-association).
#pragma once
#include "structure.cuh"

namespace dpiso {

// ---- explicit-rounding helpers: the reference kernels were compiled by nvcc with default contraction; where the
// reference expression can not contract (separate TF ops, cuBLAS calls) we must not contract either.
#ifdef __CUDA_ARCH__
DPISO_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
DPISO_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
DPISO_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
DPISO_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
DPISO_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
DPISO_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
DPISO_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
// host build uses -ffp-contract=off
DPISO_HD float fmul(float a, float b) { return a * b; }
DPISO_HD float fadd(float a, float b) { return a + b; }
DPISO_HD float fsub(float a, float b) { return a - b; }
DPISO_HD float fdiv(float a, float b) { return a / b; }
DPISO_HD double dmul(double a, double b) { return a * b; }
DPISO_HD double dadd(double a, double b) { return a + b; }
DPISO_HD double dsub(double a, double b) { return a - b; }
#endif

// ---- padded velocity access (piso_helpers.py:35-55, custom_padded with width 1) -------------------------------
// up is (ny+2) x (nx+3), vp is (ny+3) x (nx+2); (i, j) index the padded arrays.
DPISO_HD float u_padded(const float *u, int ny, int nx, int per_x, int per_y, int i, int j) {
    const int sy = per_y ? wrapi(i - 1, ny) : clampi(i - 1, 0, ny - 1);
    const int sx = per_x ? wrapi(j - 1, nx) : clampi(j - 1, 0, nx);  // periodic: duplicated last face dropped
    return u[sy * (nx + 1) + sx];
}
DPISO_HD float v_padded(const float *v, int ny, int nx, int per_x, int per_y, int i, int j) {
    const int sy = per_y ? wrapi(i - 1, ny) : clampi(i - 1, 0, ny);
    const int sx = per_x ? wrapi(j - 1, nx) : clampi(j - 1, 0, nx - 1);
    return v[sy * nx + sx];
}

DPISO_HD float flux(float a, float b, float cell_area) {
    // ".5 * (a + b) * cellArea": float sum, then double products, rounded once to float (":60,66")
    return (float)(dmul(dmul(.5, (double)fadd(a, b)), (double)cell_area));
}

// ---- one row of the advection-diffusion matrix (calcAdvetionMatrixX/Y, ":148-453") ----------------------------
// vel: this sample's flat [u, v]; values/a_diag: this component's output blocks of this sample.
// dirichlet/visc point at this component's block (visc_stride 0 = scalar).
DPISO_HD void assemble_row(int comp, int row, int ny, int nx, int per_x, int per_y, float dy, float dx, float area_x,
                           float area_y, float beta, const float *vel, const uint8_t *dirichlet_c, const float *active,
                           const uint8_t *noslip, const float *visc_c, int visc_is_field, float *values_c,
                           float *a_diag_c) {
    // per_x / per_y: bit 0 = the axis is periodic (matrix structure, wrap entries); bit 1 = the VELOCITY is nevertheless
    // padded by replication on that axis.  That combination is what the reference computes from the second unrolled step
    // on: run_piso_steps re-wraps the state as StaggeredGrid(array, box, extrapolation), whose third positional parameter is
    // `name`, so the grid falls back to the default 'boundary' extrapolation and custom_padded replicates instead of
    // wrapping (combined_training_integrated.py:431-432,473-474; SURVEY quirk list, Q21).
    const int pad_per_x = (per_x & 1) && !(per_x & 2), pad_per_y = (per_y & 1) && !(per_y & 2);
    per_x &= 1; per_y &= 1;
    const CompDims cd = comp_dims(ny, nx, comp);
    const int lx = row % cd.Dx, ly = row / cd.Dx;
    const RowLayout L = row_layout(lx, ly, cd, per_x, per_y);
    float *val = values_c + L.rp;
    for (int k = 0; k < L.len; k++) val[k] = 0.0f;          // initWithZeros (":627")
    if (dirichlet_c[row]) {                                  // ":214-238"
        val[L.slot[4]] = 1.0f;
        a_diag_c[row] = 0.0f;
        return;
    }
    const float *u = vel, *v = vel + ny * (nx + 1);
    const float cell_area[2] = {area_x, area_y};             // piso_tf.py:97: prod(dx) / (dx, dy) ~ (dy, dx)
    const float spacing[2] = {dx, dy};                       // piso_tf.py:96
    float F[4];
    if (comp == 0) {       // calcCellFluxesX (":35-69"); padded location (ly+1, lx+1)
        const float c = u_padded(u, ny, nx, pad_per_x, pad_per_y, ly + 1, lx + 1);
        F[0] = flux(c, u_padded(u, ny, nx, pad_per_x, pad_per_y, ly + 1, lx), cell_area[0]);
        F[1] = flux(u_padded(u, ny, nx, pad_per_x, pad_per_y, ly + 1, lx + 2), c, cell_area[0]);
        F[2] = flux(v_padded(v, ny, nx, pad_per_x, pad_per_y, ly + 1, lx + 1), v_padded(v, ny, nx, pad_per_x, pad_per_y, ly + 1, lx),
                    cell_area[1]);
        F[3] = flux(v_padded(v, ny, nx, pad_per_x, pad_per_y, ly + 2, lx + 1), v_padded(v, ny, nx, pad_per_x, pad_per_y, ly + 2, lx),
                    cell_area[1]);
    } else {               // calcCellFluxesY (":73-101")
        F[0] = flux(u_padded(u, ny, nx, pad_per_x, pad_per_y, ly + 1, lx + 1), u_padded(u, ny, nx, pad_per_x, pad_per_y, ly, lx + 1),
                    cell_area[0]);
        F[1] = flux(u_padded(u, ny, nx, pad_per_x, pad_per_y, ly + 1, lx + 2), u_padded(u, ny, nx, pad_per_x, pad_per_y, ly, lx + 2),
                    cell_area[0]);
        const float c = v_padded(v, ny, nx, pad_per_x, pad_per_y, ly + 1, lx + 1);
        F[2] = flux(c, v_padded(v, ny, nx, pad_per_x, pad_per_y, ly, lx + 1), cell_area[1]);
        F[3] = flux(v_padded(v, ny, nx, pad_per_x, pad_per_y, ly + 2, lx + 1), c, cell_area[1]);
    }
    // padded-centred mask cell consulted per direction (gridIDXpaddedCenteredMasks, ":132-146")
    const int wm = nx + 2;
    int m[4];
    m[0] = (ly + 1) * wm + lx;
    m[1] = (ly + 1) * wm + lx + 1 + (comp == 1);
    m[2] = ly * wm + lx + 1;
    m[3] = (ly + 1 + (comp == 0)) * wm + lx + 1;
    const float nu = visc_c[visc_is_field ? row : 0];
    float diag = 0.0f;
    for (int d = 1; d >= 0; d--) {                           // y first, then x (":248")
        const float D = fdiv(fmul(nu, cell_area[d]), spacing[d]);
        const int not_stag = (d != comp);
        for (int side = 0; side < 2; side++) {
            const int k = 2 * d + side;
            const int ns = noslip[m[k]] ? 1 : 0;
            const int t = (active[m[k]] == 1.0f) || (L.reg[k] && ns);
            const float sf = side == 0 ? F[k] : -F[k];
            if (t && L.has[k]) val[L.slot[k]] = (float)dadd(dmul((double)sf, .5), (double)D);
            const int kfac = t + not_stag * (1 - t) * ns * 2;
            const double term = dsub(dmul((double)fmul(sf, (float)(2 - t)), .5), (double)fmul(D, (float)kfac));
            diag = (float)dadd((double)diag, term);
        }
    }
    val[L.slot[4]] = fsub(diag, beta);                       // ":294"
    a_diag_c[row] = diag;
}

// ---- ghost pressure + finite-volume gradient on one face (piso_helpers.py:236-274) ---------------------------
DPISO_HD float p_ghost(const float *p, int ny, int nx, int cy, int cx, const int *pbc) {
    if (cy < 0) { if (pbc[0] == DPISO_PBC_ZERO) return 0.0f; cy = pbc[0] == DPISO_PBC_PERIODIC ? ny - 1 : 0; }
    if (cy >= ny) { if (pbc[1] == DPISO_PBC_ZERO) return 0.0f; cy = pbc[1] == DPISO_PBC_PERIODIC ? 0 : ny - 1; }
    if (cx < 0) { if (pbc[2] == DPISO_PBC_ZERO) return 0.0f; cx = pbc[2] == DPISO_PBC_PERIODIC ? nx - 1 : 0; }
    if (cx >= nx) { if (pbc[3] == DPISO_PBC_ZERO) return 0.0f; cx = pbc[3] == DPISO_PBC_PERIODIC ? 0 : nx - 1; }
    return p[cy * nx + cx];
}

// face index i in the flat [u, v] numbering of one sample -> gradient value
DPISO_HD float fv_gradient_face(int i, int ny, int nx, float dy, float dx, float prod, const int *pbc,
                                const float *access, const float *p) {
    const int n_u = ny * (nx + 1), wm = nx + 2;
    if (i < n_u) {
        const int cy = i / (nx + 1), fx = i % (nx + 1);
        const float diff = fsub(p_ghost(p, ny, nx, cy, fx, pbc), p_ghost(p, ny, nx, cy, fx - 1, pbc));
        const float mk = fminf(access[(cy + 1) * wm + fx], access[(cy + 1) * wm + fx + 1]);
        return fmul(fdiv(fmul(diff, prod), dx), mk);
    }
    const int j = i - n_u;
    const int fy = j / nx, cx = j % nx;
    const float diff = fsub(p_ghost(p, ny, nx, fy, cx, pbc), p_ghost(p, ny, nx, fy - 1, cx, pbc));
    const float mk = fminf(access[fy * wm + cx + 1], access[(fy + 1) * wm + cx + 1]);
    return fmul(fdiv(fmul(diff, prod), dy), mk);
}

// ---- divergence of one cell (piso_helpers.py:285-289); optional division of the faces by (beta - a_diag) ------
DPISO_HD float face_scaled(const float *vel, const float *a_diag, float beta, int i) {
    return a_diag ? fdiv(vel[i], fsub(beta, a_diag[i])) : vel[i];
}
DPISO_HD float fv_divergence_cell(int c, int ny, int nx, float dy, float dx, float prod, const float *vel,
                                  const float *a_diag, float beta) {
    const int cy = c / nx, cx = c % nx, n_u = ny * (nx + 1);
    const float vt = face_scaled(vel, a_diag, beta, n_u + (cy + 1) * nx + cx);
    const float vb = face_scaled(vel, a_diag, beta, n_u + cy * nx + cx);
    const float ur = face_scaled(vel, a_diag, beta, cy * (nx + 1) + cx + 1);
    const float ul = face_scaled(vel, a_diag, beta, cy * (nx + 1) + cx);
    const float ty = fdiv(fmul(fsub(vt, vb), prod), dy);
    const float tx = fdiv(fmul(fsub(ur, ul), prod), dx);
    return fadd(ty, tx);                                    // math.sum over [y-term, x-term]
}

// ---- one row of the PISO pressure matrix (calcPISOLaplaceMatrix, laplace_op.cu.cc:79-179) --------------------
// kf[4] = face coefficients y-, x-, x+, y+ ; out[5] = [y-, x-, diag, x+, y+]
template <typename T>
DPISO_HD void laplace_row(int cy, int cx, int nx, const float *active, const float *fluid, const float kf[4],
                          T out[5]) {
    const int wm = nx + 2, me = (cy + 1) * wm + cx + 1;
    const int mnb[4] = {me - wm, me - 1, me + 1, me + wm};
    const bool self_active = active[me] != 0.0f;
    const bool self_solid = (active[me] == 0.0f && fluid[me] == 0.0f);
    T diag = 0;
    const int order[4] = {0, 3, 1, 2};                      // y-, y+, x-, x+ (":118-135")
    for (int q = 0; q < 4; q++) {
        const int k = order[q];
        const bool nb_solid = (active[mnb[k]] == 0.0f && fluid[mnb[k]] == 0.0f);
        if (!nb_solid && self_active) diag -= (T)kf[k];
    }
    T off[4];
    for (int k = 0; k < 4; k++) {
        const bool nb_fluid = (active[mnb[k]] == 1.0f && fluid[mnb[k]] == 1.0f);
        off[k] = (nb_fluid && !self_solid) ? (T)kf[k] : (T)0;
    }
    out[0] = off[0]; out[1] = off[1]; out[2] = diag; out[3] = off[2]; out[4] = off[3];
}

// scaling field from the matrix diagonal: (1/(beta - A)) * dx_factor   (piso_tf.py:53-54)
DPISO_HD float k_from_adiag(float a, float beta, float dx_factor) { return fmul(fdiv(1.0f, fsub(beta, a)), dx_factor); }

}  // namespace dpiso
