// piso_ops.cu -- assembly, pressure-matrix and pointwise kernels of the PISO step (sm_100a).
//
// All of these are single-pass streaming kernels (one thread per face row / cell, batch folded into the grid);
// they account for ~2% of the algorithmic traffic of a step, the solvers for the rest.
// Arithmetic lives in rows.cuh so that the CPU test-suite can run the very same per-row code on the host.
#include <mutex>
#include <string>

#include "rows.cuh"

namespace dpiso {

static thread_local std::string g_last_error;
void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

constexpr int kThreads = 256;
static inline unsigned blocks_for(long long n) { return (unsigned)((n + kThreads - 1) / kThreads); }

struct Pbc { int v[4]; };

// ---------------------------------------------------------------------------------------------------------------
__global__ void csr_structure_kernel(int ny, int nx, int per_x, int per_y, int n_u, int n_v, int nnz_u,
                                     int *__restrict__ row_ptr, int *__restrict__ col_ind) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_u + n_v) return;
    const int comp = i >= n_u;
    const int row = comp ? i - n_u : i;
    const CompDims cd = comp_dims(ny, nx, comp);
    const RowLayout L = row_layout(row % cd.Dx, row / cd.Dx, cd, per_x, per_y);
    int *rp = row_ptr + (comp ? n_u + 1 : 0);
    int *ci = col_ind + (comp ? nnz_u : 0);
    rp[row] = L.rp;
    if (row == (comp ? n_v : n_u) - 1) rp[row + 1] = L.rp + L.len;
#pragma unroll
    for (int k = 0; k < 5; k++)
        if (k == 4 || L.has[k]) ci[L.rp + L.slot[k]] = L.col[k];
}

__global__ void assemble_kernel(int batch, Grid g, float dy, float dx, float area_x, float area_y, float beta,
                                const float *__restrict__ vel,
                                const uint8_t *__restrict__ dirichlet, const float *__restrict__ active,
                                const uint8_t *__restrict__ noslip, const float *__restrict__ visc, int visc_mode,
                                float *__restrict__ values, float *__restrict__ a_diag) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = g.nf();
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    const int comp = i >= g.n_u;
    const int row = comp ? i - g.n_u : i;
    const int face_off = comp ? g.n_u : 0;
    const float *visc_c = visc;
    if (visc_mode == 1) visc_c = visc + face_off;
    else if (visc_mode == 2) visc_c = visc + (size_t)b * nf + face_off;
    assemble_row(comp, row, g.ny, g.nx, g.per_x, g.per_y, dy, dx, area_x, area_y, beta, vel + (size_t)b * nf,
                 dirichlet + face_off,
                 active, noslip, visc_c, visc_mode != 0, values + (size_t)b * g.nnz() + (comp ? g.nnz_u : 0),
                 a_diag + (size_t)b * nf + face_off);
}

__global__ void predictor_rhs_kernel(int batch, int ny, int nx, float dy, float dx, float prod, float beta, Pbc pbc,
                                     const float *__restrict__ vel, const float *__restrict__ pres,
                                     const float *__restrict__ access, const uint8_t *__restrict__ dirichlet,
                                     const float *__restrict__ dvals, int dvals_batch,
                                     const float *__restrict__ forcing, float *__restrict__ rhs) {
    const int nf = ny * (nx + 1) + (ny + 1) * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    if (dirichlet[i]) {                                       // piso_helpers.py:170
        rhs[t] = fmul(dvals[dvals_batch ? t : i], -1.0f);
        return;
    }
    const float gval = fv_gradient_face(i, ny, nx, dy, dx, prod, pbc.v, access, pres + (size_t)b * ny * nx);
    float r = fsub(fmul(vel[t], beta), gval);                 // piso_tf.py:36
    if (forcing) r = fadd(r, fmul(forcing[t], prod));         // piso_tf.py:38
    rhs[t] = r;
}

__global__ void fv_gradient_kernel(int batch, int ny, int nx, float dy, float dx, float prod, Pbc pbc,
                                   const float *__restrict__ access, const float *__restrict__ p,
                                   float *__restrict__ gout) {
    const int nf = ny * (nx + 1) + (ny + 1) * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    gout[t] = fv_gradient_face(i, ny, nx, dy, dx, prod, pbc.v, access, p + (size_t)b * ny * nx);
}

__global__ void fv_divergence_kernel(int batch, int ny, int nx, float dy, float dx, float prod,
                                     const float *__restrict__ vel, const float *__restrict__ a_diag, float beta,
                                     float *__restrict__ div) {
    const int nc = ny * nx, nf = ny * (nx + 1) + (ny + 1) * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nc) return;
    const int b = (int)(t / nc), c = (int)(t % nc);
    div[t] = fv_divergence_cell(c, ny, nx, dy, dx, prod, vel + (size_t)b * nf,
                                a_diag ? a_diag + (size_t)b * nf : nullptr, beta);
}

__global__ void corrector1_kernel(int batch, int ny, int nx, float dy, float dx, float prod, float beta, Pbc pbc,
                                  const float *__restrict__ access, const float *__restrict__ u_star,
                                  const float *__restrict__ p1, const float *__restrict__ a_diag,
                                  float *__restrict__ u_s2) {
    const int nf = ny * (nx + 1) + (ny + 1) * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    const float gval = fv_gradient_face(i, ny, nx, dy, dx, prod, pbc.v, access, p1 + (size_t)b * ny * nx);
    u_s2[t] = fsub(u_star[t], fdiv(fdiv(gval, fsub(beta, a_diag[t])), prod));   // piso_tf.py:58
}

__global__ void h_apply_kernel(int batch, Grid g, float beta, const float *__restrict__ values,
                               const float *__restrict__ a_diag, const float *__restrict__ u_star,
                               const float *__restrict__ u_s2, float *__restrict__ h) {
    const int nf = g.nf();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    const int comp = i >= g.n_u;
    const int row = comp ? i - g.n_u : i;
    const size_t fo = (size_t)b * nf + (comp ? g.n_u : 0);
    const CompDims cd = comp_dims(g.ny, g.nx, comp);
    const RowLayout L = row_layout(row % cd.Dx, row / cd.Dx, cd, g.per_x, g.per_y);
    const float *val = values + (size_t)b * g.nnz() + (comp ? g.nnz_u : 0) + L.rp;
    // gather * values, segment_sum in CSR order (piso_helpers.py:217-222): separate multiply and add
    int ord[5];
#pragma unroll
    for (int k = 0; k < 5; k++) ord[k] = -1;
#pragma unroll
    for (int k = 0; k < 5; k++)
        if (k == 4 || L.has[k]) ord[L.slot[k]] = k;
    float acc = 0.0f;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        if (s < L.len) {
            const int col = L.col[ord[s]];
            const float d = fsub(u_s2[fo + col], u_star[fo + col]);
            acc = fadd(acc, fmul(d, val[s]));
        }
    }
    const float di = fsub(u_s2[t], u_star[t]);
    h[t] = fsub(acc, fmul(fsub(a_diag[t], beta), di));       // piso_helpers.py:223
}

__global__ void corrector2_kernel(int batch, int ny, int nx, float dy, float dx, float prod, float beta, Pbc pbc,
                                  const float *__restrict__ access, const float *__restrict__ u_s2,
                                  const float *__restrict__ h, const float *__restrict__ p2,
                                  const float *__restrict__ a_diag, const float *__restrict__ p,
                                  const float *__restrict__ p1, float *__restrict__ u_next,
                                  float *__restrict__ p_next) {
    const int nf = ny * (nx + 1) + (ny + 1) * nx, nc = ny * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)batch * nf) {
        const int b = (int)(t / nf), i = (int)(t % nf);
        const float gval = fv_gradient_face(i, ny, nx, dy, dx, prod, pbc.v, access, p2 + (size_t)b * nc);
        // velocity_s2 + (H - G(p2)/prod(dx)) / (beta - A)     (piso_tf.py:71-72)
        u_next[t] = fadd(u_s2[t], fdiv(fsub(h[t], fdiv(gval, prod)), fsub(beta, a_diag[t])));
    }
    if (t < (long long)batch * nc) p_next[t] = fadd(fadd(p[t], p1[t]), p2[t]);   // piso_tf.py:75
}

template <typename T>
__global__ void laplace_kernel(int batch, int ny, int nx, const float *__restrict__ active,
                               const float *__restrict__ fluid, const float *__restrict__ k_faces, int mode,
                               float beta, float dx_factor, T *__restrict__ lap) {
    const int nc = ny * nx, n_u = ny * (nx + 1), n_v = (ny + 1) * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nc) return;
    const int b = (int)(t / nc), c = (int)(t % nc);
    const int cy = c / nx, cx = c % nx;
    const float *kb = k_faces + (size_t)b * (n_u + n_v);
    float kf[4];
    if (mode == 0) {      // [v, u] scaling field (piso_cuda_pressure_solver.py:70; gridIDXForStaggered)
        const float *kv = kb, *ku = kb + n_v;
        kf[0] = kv[cy * nx + cx]; kf[1] = ku[cy * (nx + 1) + cx];
        kf[2] = ku[cy * (nx + 1) + cx + 1]; kf[3] = kv[(cy + 1) * nx + cx];
    } else {              // a_diag in [u, v] order, scaling formed on the fly
        const float *au = kb, *av = kb + n_u;
        kf[0] = k_from_adiag(av[cy * nx + cx], beta, dx_factor);
        kf[1] = k_from_adiag(au[cy * (nx + 1) + cx], beta, dx_factor);
        kf[2] = k_from_adiag(au[cy * (nx + 1) + cx + 1], beta, dx_factor);
        kf[3] = k_from_adiag(av[(cy + 1) * nx + cx], beta, dx_factor);
    }
    T out[5];
    laplace_row<T>(cy, cx, nx, active, fluid, kf, out);
    T *o = lap + (size_t)t * 5;
#pragma unroll
    for (int k = 0; k < 5; k++) o[k] = out[k];
}

}  // namespace dpiso

using namespace dpiso;

static int check_grid(int batch, int ny, int nx) {
    DPISO_REQUIRE(batch >= 1, "batch must be >= 1 (got %d)", batch);
    DPISO_REQUIRE(ny >= 3 && nx >= 3, "grid too small for the 5-point pattern (need ny,nx >= 3, got %d x %d)", ny, nx);
    DPISO_REQUIRE((long long)batch * (5LL * ((long long)ny * (nx + 1) + (long long)(ny + 1) * nx)) < (1LL << 40),
                  "problem too large");
    return DPISO_OK;
}

static Pbc load_pbc(const int *h_pbc) {
    Pbc p;
    for (int i = 0; i < 4; i++) p.v[i] = h_pbc[i];
    return p;
}

extern "C" {

int dpiso_version(void) { return 100; }
const char *dpiso_last_error(void) { return g_last_error.c_str(); }

int dpiso_sizes(int ny, int nx, int per_x, int per_y, int *h_n, int *h_nnz) {
    if (int rc = check_grid(1, ny, nx)) return rc;
    const Grid g = make_grid(ny, nx, per_x, per_y);
    h_n[0] = g.n_u; h_n[1] = g.n_v; h_nnz[0] = g.nnz_u; h_nnz[1] = g.nnz_v;
    return DPISO_OK;
}

int dpiso_csr_structure(int ny, int nx, int per_x, int per_y, int *row_ptr, int *col_ind, void *stream) {
    if (int rc = check_grid(1, ny, nx)) return rc;
    DPISO_REQUIRE(row_ptr && col_ind, "null output pointer");
    const Grid g = make_grid(ny, nx, per_x, per_y);
    csr_structure_kernel<<<blocks_for(g.nf()), kThreads, 0, (cudaStream_t)stream>>>(ny, nx, g.per_x, g.per_y, g.n_u,
                                                                                   g.n_v, g.nnz_u, row_ptr, col_ind);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_assemble(int batch, int ny, int nx, int per_x, int per_y, float dy, float dx, float area_x, float area_y,
                   float beta, const float *vel,
                   const uint8_t *dirichlet, const float *active, const uint8_t *noslip, const float *visc,
                   int visc_mode, float *values, float *a_diag, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(vel && dirichlet && active && noslip && visc && values && a_diag, "null pointer");
    DPISO_REQUIRE(visc_mode >= 0 && visc_mode <= 2, "visc_mode must be 0, 1 or 2");
    Grid g = make_grid(ny, nx, per_x & 1, per_y & 1);
    g.per_x = per_x & 3; g.per_y = per_y & 3;                     // bit 1 travels to assemble_row (replicated velocity padding)
    assemble_kernel<<<blocks_for((long long)batch * g.nf()), kThreads, 0, (cudaStream_t)stream>>>(
        batch, g, dy, dx, area_x, area_y, beta, vel, dirichlet, active, noslip, visc, visc_mode, values, a_diag);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

static inline float cell_prod(float dy, float dx) { return (float)((double)dy * (double)dx); }

int dpiso_predictor_rhs(int batch, int ny, int nx, float dy, float dx, float beta, const int *h_pbc, const float *vel,
                        const float *pres, const float *access, const uint8_t *dirichlet, const float *dvals,
                        int dvals_batch, const float *forcing, float *rhs, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(h_pbc && vel && pres && access && dirichlet && dvals && rhs, "null pointer");
    const long long n = (long long)batch * (ny * (nx + 1) + (ny + 1) * nx);
    predictor_rhs_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, dy, dx, cell_prod(dy, dx), beta, load_pbc(h_pbc), vel, pres, access, dirichlet, dvals,
        dvals_batch, forcing, rhs);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_fv_gradient(int batch, int ny, int nx, float dy, float dx, const int *h_pbc, const float *access,
                      const float *p, float *g, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(h_pbc && access && p && g, "null pointer");
    const long long n = (long long)batch * (ny * (nx + 1) + (ny + 1) * nx);
    fv_gradient_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(batch, ny, nx, dy, dx, cell_prod(dy, dx),
                                                                            load_pbc(h_pbc), access, p, g);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_fv_divergence(int batch, int ny, int nx, float dy, float dx, const float *vel, const float *a_diag,
                        float beta, float *div, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(vel && div, "null pointer");
    fv_divergence_kernel<<<blocks_for((long long)batch * ny * nx), kThreads, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, dy, dx, cell_prod(dy, dx), vel, a_diag, beta, div);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_corrector1(int batch, int ny, int nx, float dy, float dx, float beta, const int *h_pbc, const float *access,
                     const float *u_star, const float *p1, const float *a_diag, float *u_s2, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(h_pbc && access && u_star && p1 && a_diag && u_s2, "null pointer");
    const long long n = (long long)batch * (ny * (nx + 1) + (ny + 1) * nx);
    corrector1_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, dy, dx, cell_prod(dy, dx), beta, load_pbc(h_pbc), access, u_star, p1, a_diag, u_s2);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_h_apply(int batch, int ny, int nx, int per_x, int per_y, float beta, const float *values,
                  const float *a_diag, const float *u_star, const float *u_s2, float *h, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(values && a_diag && u_star && u_s2 && h, "null pointer");
    const Grid g = make_grid(ny, nx, per_x, per_y);
    h_apply_kernel<<<blocks_for((long long)batch * g.nf()), kThreads, 0, (cudaStream_t)stream>>>(
        batch, g, beta, values, a_diag, u_star, u_s2, h);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_corrector2(int batch, int ny, int nx, float dy, float dx, float beta, const int *h_pbc, const float *access,
                     const float *u_s2, const float *h, const float *p2, const float *a_diag, const float *p,
                     const float *p1, float *u_next, float *p_next, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(h_pbc && access && u_s2 && h && p2 && a_diag && p && p1 && u_next && p_next, "null pointer");
    const long long n = (long long)batch * (ny * (nx + 1) + (ny + 1) * nx);
    corrector2_kernel<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, dy, dx, cell_prod(dy, dx), beta, load_pbc(h_pbc), access, u_s2, h, p2, a_diag, p, p1, u_next,
        p_next);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_laplace_f64(int batch, int ny, int nx, const float *active, const float *fluid, const float *k_faces,
                      int mode, float beta, float dx_factor, double *lap, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(active && fluid && k_faces && lap, "null pointer");
    laplace_kernel<double><<<blocks_for((long long)batch * ny * nx), kThreads, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, active, fluid, k_faces, mode, beta, dx_factor, lap);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_laplace_f32(int batch, int ny, int nx, const float *active, const float *fluid, const float *k_faces,
                      int mode, float beta, float dx_factor, float *lap, void *stream) {
    if (int rc = check_grid(batch, ny, nx)) return rc;
    DPISO_REQUIRE(active && fluid && k_faces && lap, "null pointer");
    laplace_kernel<float><<<blocks_for((long long)batch * ny * nx), kThreads, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, active, fluid, k_faces, mode, beta, dx_factor, lap);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

}  // extern "C"
