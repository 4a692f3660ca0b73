// tests/host_shim.cpp -- TEST INFRASTRUCTURE.  Compiles the per-row arithmetic of the CUDA kernels
// (differentiable-piso_b200/csrc/rows.cuh, structure.cuh) for the HOST so that the CPU test-suite can check the very
// code the device executes against the oracle without a GPU.  Never loaded by the product.
#include <cstring>

#include "../differentiable-piso_b200/csrc/rows.cuh"

namespace dpiso {
void set_error(const char *, ...) {}
}
using namespace dpiso;

extern "C" {

void hs_csr_structure(int ny, int nx, int per_x, int per_y, int *row_ptr, int *col_ind) {
    const Grid g = make_grid(ny, nx, per_x, per_y);
    for (int comp = 0; comp < 2; comp++) {
        const CompDims cd = comp_dims(ny, nx, comp);
        const int n = comp ? g.n_v : g.n_u;
        int *rp = row_ptr + (comp ? g.n_u + 1 : 0);
        int *ci = col_ind + (comp ? g.nnz_u : 0);
        for (int row = 0; row < n; row++) {
            const RowLayout L = row_layout(row % cd.Dx, row / cd.Dx, cd, per_x, per_y);
            rp[row] = L.rp;
            if (row == n - 1) rp[row + 1] = L.rp + L.len;
            for (int k = 0; k < 5; k++)
                if (k == 4 || L.has[k]) ci[L.rp + L.slot[k]] = L.col[k];
        }
    }
}

void hs_assemble(int ny, int nx, int per_x, int per_y, float dy, float dx, float area_x, float area_y, float beta,
                 const float *vel,
                 const uint8_t *dirichlet, const float *active, const uint8_t *noslip, const float *visc,
                 int visc_is_field, float *values, float *a_diag) {
    const Grid g = make_grid(ny, nx, per_x, per_y);
    for (int comp = 0; comp < 2; comp++) {
        const int n = comp ? g.n_v : g.n_u, fo = comp ? g.n_u : 0;
        for (int row = 0; row < n; row++)
            assemble_row(comp, row, ny, nx, per_x, per_y, dy, dx, area_x, area_y, beta, vel, dirichlet + fo, active, noslip,
                         visc_is_field ? visc + fo : visc, visc_is_field, values + (comp ? g.nnz_u : 0), a_diag + fo);
    }
}

void hs_fv_gradient(int ny, int nx, float dy, float dx, const int *pbc, const float *access, const float *p, float *g) {
    const float prod = (float)((double)dy * (double)dx);
    const int nf = ny * (nx + 1) + (ny + 1) * nx;
    for (int i = 0; i < nf; i++) g[i] = fv_gradient_face(i, ny, nx, dy, dx, prod, pbc, access, p);
}

void hs_fv_divergence(int ny, int nx, float dy, float dx, const float *vel, const float *a_diag, float beta, float *div) {
    const float prod = (float)((double)dy * (double)dx);
    for (int c = 0; c < ny * nx; c++) div[c] = fv_divergence_cell(c, ny, nx, dy, dx, prod, vel, a_diag, beta);
}

void hs_laplace_f64(int ny, int nx, const float *active, const float *fluid, const float *k_faces, double *lap) {
    const float *kv = k_faces, *ku = k_faces + (ny + 1) * nx;
    for (int cy = 0; cy < ny; cy++)
        for (int cx = 0; cx < nx; cx++) {
            const float kf[4] = {kv[cy * nx + cx], ku[cy * (nx + 1) + cx], ku[cy * (nx + 1) + cx + 1], kv[(cy + 1) * nx + cx]};
            laplace_row<double>(cy, cx, nx, active, fluid, kf, lap + 5 * (size_t)(cy * nx + cx));
        }
}

void hs_laplace_f32(int ny, int nx, const float *active, const float *fluid, const float *k_faces, float *lap) {
    const float *kv = k_faces, *ku = k_faces + (ny + 1) * nx;
    for (int cy = 0; cy < ny; cy++)
        for (int cx = 0; cx < nx; cx++) {
            const float kf[4] = {kv[cy * nx + cx], ku[cy * (nx + 1) + cx], ku[cy * (nx + 1) + cx + 1], kv[(cy + 1) * nx + cx]};
            laplace_row<float>(cy, cx, nx, active, fluid, kf, lap + 5 * (size_t)(cy * nx + cx));
        }
}

}  // extern "C"
