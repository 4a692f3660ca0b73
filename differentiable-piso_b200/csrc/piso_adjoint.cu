// piso_adjoint.cu -- transposed pointwise operators used by the backward pass of one PISO step.
//
// The reference's backward is what TF-1.14 autodiff assembles from its gradient registrations:
// matrices, 1/(beta-A), masks and Dirichlet values are constants; the FV gradient / divergence use the REGISTERED
// gradients, which on periodic axes are not the exact transposes (quirks Q19, Q20) -- reproduced here on purpose.
#include "rows.cuh"

namespace dpiso {

constexpr int kThreadsAdj = 256;
static inline unsigned blocks_adj(long long n) { return (unsigned)((n + kThreadsAdj - 1) / kThreadsAdj); }
struct PbcA { int v[4]; };

// upstream value on one face, brought to "gradient of the raw difference" form:
//   t = gs [/ (beta - a_diag)] [/ divisor] [negated];   s = ((t * mask) / d_dim) * prod
__device__ __forceinline__ float face_s(const float *gs, const float *a_diag, float beta, float divisor, int negate,
                                        int f, float mk, float d_dim, float prod) {
    float t = gs[f];
    if (a_diag) t = fdiv(t, fsub(beta, a_diag[f]));
    if (divisor != 1.0f) t = fdiv(t, divisor);
    if (negate) t = -t;
    return fmul(fdiv(fmul(t, mk), d_dim), prod);
}

// gp = [base] + G^T(...)      finite_volume_gradient_tensor backward (piso_helpers.py:236-266) with
// circular_padded_gradient's registered gradient on periodic axes (":230-232", Q20): the two end faces behave like
// faces next to an independent ghost cell; replicate ghosts cancel the end-face contribution, zero ghosts keep it.
__global__ void fv_gradient_adj_kernel(int batch, int ny, int nx, float dy, float dx, float prod, PbcA pbc,
                                       const float *__restrict__ access, const float *__restrict__ gs,
                                       const float *__restrict__ a_diag, float beta, float divisor, int negate,
                                       const float *__restrict__ base, float *__restrict__ gp) {
    const int nc = ny * nx, n_u = ny * (nx + 1), nf = n_u + (ny + 1) * nx, wm = nx + 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nc) return;
    const int b = (int)(t / nc), c = (int)(t % nc);
    const int cy = c / nx, cx = c % nx;
    const float *g = gs + (size_t)b * nf;
    const float *a = a_diag ? a_diag + (size_t)b * nf : nullptr;
    // u faces (cy, cx) and (cy, cx+1)
    const int fl = cy * (nx + 1) + cx, fr = fl + 1;
    const float mk_l = fminf(access[(cy + 1) * wm + cx], access[(cy + 1) * wm + cx + 1]);
    const float mk_r = fminf(access[(cy + 1) * wm + cx + 1], access[(cy + 1) * wm + cx + 2]);
    const float wl = (cx == 0 && pbc.v[2] == DPISO_PBC_REPLICATE) ? 0.0f : 1.0f;
    const float wr = (cx == nx - 1 && pbc.v[3] == DPISO_PBC_REPLICATE) ? 0.0f : 1.0f;
    // v faces (cy, cx) and (cy+1, cx)
    const int fb = n_u + cy * nx + cx, ft = fb + nx;
    const float mk_b = fminf(access[cy * wm + cx + 1], access[(cy + 1) * wm + cx + 1]);
    const float mk_t = fminf(access[(cy + 1) * wm + cx + 1], access[(cy + 2) * wm + cx + 1]);
    const float wb = (cy == 0 && pbc.v[0] == DPISO_PBC_REPLICATE) ? 0.0f : 1.0f;
    const float wt = (cy == ny - 1 && pbc.v[1] == DPISO_PBC_REPLICATE) ? 0.0f : 1.0f;
    float acc = base ? base[t] : 0.0f;
    acc = fadd(acc, fsub(fmul(wb, face_s(g, a, beta, divisor, negate, fb, mk_b, dy, prod)),
                         fmul(wt, face_s(g, a, beta, divisor, negate, ft, mk_t, dy, prod))));
    acc = fadd(acc, fsub(fmul(wl, face_s(g, a, beta, divisor, negate, fl, mk_l, dx, prod)),
                         fmul(wr, face_s(g, a, beta, divisor, negate, fr, mk_r, dx, prod))));
    gp[t] = acc;
}

// gv = ([base] + D^T gc) [/ (beta - a_diag)]     registered gradient of finite_volume_divergence
// (piso_helpers.py:291-305): non-periodic axis  v[j] = -(g[j])*c + (g[j-1])*c with zero ghosts; periodic axis (Q19)
// v[0] = g[N-2]*c - g[0]*c,  v[N] = g[N-1]*c - g[0]*c.
__global__ void fv_divergence_adj_kernel(int batch, int ny, int nx, int per_x, int per_y, float dy, float dx, float prod,
                                         const float *__restrict__ gc, const float *__restrict__ base,
                                         const float *__restrict__ base_sub, const float *__restrict__ a_diag,
                                         float beta, float *__restrict__ gv) {
    const int nc = ny * nx, n_u = ny * (nx + 1), nf = n_u + (ny + 1) * nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    const float *g = gc + (size_t)b * nc;
    float lo, hi, d;     // result = -(hi)*c + (lo)*c
    if (i < n_u) {
        const int cy = i / (nx + 1), j = i % (nx + 1);
        d = dx;
        if (per_x) {
            hi = g[cy * nx + (j == nx ? 0 : j)];
            lo = g[cy * nx + (j == 0 ? nx - 2 : j - 1)];
        } else {
            hi = j < nx ? g[cy * nx + j] : 0.0f;
            lo = j > 0 ? g[cy * nx + j - 1] : 0.0f;
        }
    } else {
        const int k = i - n_u;
        const int j = k / nx, cx = k % nx;
        d = dy;
        if (per_y) {
            hi = g[(j == ny ? 0 : j) * nx + cx];
            lo = g[(j == 0 ? ny - 2 : j - 1) * nx + cx];
        } else {
            hi = j < ny ? g[j * nx + cx] : 0.0f;
            lo = j > 0 ? g[(j - 1) * nx + cx] : 0.0f;
        }
    }
    float v = fadd(fdiv(fmul(-hi, prod), d), fdiv(fmul(lo, prod), d));
    if (base) v = fadd(base_sub ? fsub(base[t], base_sub[t]) : base[t], v);
    if (a_diag) v = fdiv(v, fsub(beta, a_diag[t]));
    gv[t] = v;
}

struct AdjTab { int n, wa; const int *perm, *a_col, *a_src, *r_col, *r_src; };

// gd = M^T gh - (A - beta) gh : adjoint of explicit_H_csr w.r.t. its vector argument.  Uses the level-major tables of
// M^T (built for the adjoint BiCGStab): row q of M^T lists the entries M(col, row) with their CSR positions.
__global__ void h_apply_adj_kernel(int batch, AdjTab tu, AdjTab tv, int nnz_u, int nnz_v, float beta,
                                   const float *__restrict__ values, const float *__restrict__ a_diag,
                                   const float *__restrict__ gh, float *__restrict__ gd,
                                   const float *__restrict__ base, float *__restrict__ sum) {
    const int nf = tu.n + tv.n;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int b = (int)(t / nf), i = (int)(t % nf);
    const int comp = i >= tu.n;
    const AdjTab &T = comp ? tv : tu;
    const int q = comp ? i - tu.n : i;
    const size_t fo = (size_t)b * nf + (comp ? tu.n : 0);
    const float *val = values + (size_t)b * (nnz_u + nnz_v) + (comp ? nnz_u : 0);
    if (T.r_col) {
        // row-major tables (original row order, original column indices, same ascending-column entry order): coalesced
        // output, neighbour gathers; identical sums
        float acc = 0.0f;
        for (int k = 0; k < T.wa; k++) {
            const int src = T.r_src[k * T.n + q];
            if (src >= 0) acc = fadd(acc, fmul(val[src], gh[fo + T.r_col[k * T.n + q]]));
        }
        const float out = fsub(acc, fmul(fsub(a_diag[fo + q], beta), gh[fo + q]));
        gd[fo + q] = out;
        if (sum) sum[fo + q] = fadd(base[fo + q], out);
        return;
    }
    const int row = T.perm[q];
    float acc = 0.0f;
    for (int k = 0; k < T.wa; k++) {
        const int src = T.a_src[k * T.n + q];
        if (src >= 0) acc = fadd(acc, fmul(val[src], gh[fo + T.perm[T.a_col[k * T.n + q]]]));
    }
    const float out = fsub(acc, fmul(fsub(a_diag[fo + row], beta), gh[fo + row]));
    gd[fo + row] = out;
    if (sum) sum[fo + row] = fadd(base[fo + row], out);
}

// adjoint of the rhs assembly (piso_tf.py:36-40, piso_helpers.py:170): with m = dirichlet mask,
//   gvel = (1-m)*grhs*beta,  gforce = (1-m)*grhs*prod,  gdvals = -m*grhs,  gfree = (1-m)*grhs (input of -G^T)
__global__ void predictor_rhs_adj_kernel(int batch, int nf, float prod, float beta, const uint8_t *__restrict__ dirichlet,
                                         const float *__restrict__ grhs, const int *__restrict__ solve_stats,
                                         float *__restrict__ gvel,
                                         float *__restrict__ gforce, float *__restrict__ gdvals,
                                         float *__restrict__ gfree) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)batch * nf) return;
    const int i = (int)(t % nf);
    float g = grhs[t];
    if (solve_stats) {          // df * (1 - warn) of linear_solver.py:169-173, per sample
        const int b = (int)(t / nf);
        const float keep = (solve_stats[b * 8 + 2] | solve_stats[b * 8 + 6]) ? 0.0f : 1.0f;
        g = fmul(g, keep);
    }
    const bool m = dirichlet[i] != 0;
    const float gf = m ? 0.0f : g;
    gvel[t] = fmul(gf, beta);
    if (gforce) gforce[t] = fmul(gf, prod);
    if (gdvals) gdvals[t] = m ? -g : 0.0f;
    gfree[t] = gf;
}

}  // namespace dpiso

using namespace dpiso;

static inline float cell_prod_adj(float dy, float dx) { return (float)((double)dy * (double)dx); }

extern "C" {

int dpiso_fv_gradient_adj(int batch, int ny, int nx, float dy, float dx, const int *h_pbc, const float *access,
                          const float *gs, const float *a_diag, float beta, float divisor, int negate,
                          const float *base, float *gp, void *stream) {
    DPISO_REQUIRE(batch >= 1 && ny >= 3 && nx >= 3, "bad sizes");
    DPISO_REQUIRE(h_pbc && access && gs && gp, "null pointer");
    PbcA pbc;
    for (int i = 0; i < 4; i++) pbc.v[i] = h_pbc[i];
    fv_gradient_adj_kernel<<<blocks_adj((long long)batch * ny * nx), kThreadsAdj, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, dy, dx, cell_prod_adj(dy, dx), pbc, access, gs, a_diag, beta, divisor, negate, base, gp);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_fv_divergence_adj(int batch, int ny, int nx, int per_x, int per_y, float dy, float dx, const float *gc,
                            const float *base, const float *base_sub, const float *a_diag, float beta, float *gv,
                            void *stream) {
    DPISO_REQUIRE(batch >= 1 && ny >= 3 && nx >= 3, "bad sizes");
    DPISO_REQUIRE(gc && gv && (base || !base_sub), "null pointer");
    const long long n = (long long)batch * (ny * (nx + 1) + (ny + 1) * nx);
    fv_divergence_adj_kernel<<<blocks_adj(n), kThreadsAdj, 0, (cudaStream_t)stream>>>(
        batch, ny, nx, per_x ? 1 : 0, per_y ? 1 : 0, dy, dx, cell_prod_adj(dy, dx), gc, base, base_sub, a_diag, beta, gv);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_h_apply_adj(int batch, const dpiso_bicg_tables *h_tabT_u, const dpiso_bicg_tables *h_tabT_v, int nnz_u,
                      int nnz_v, float beta, const float *values, const float *a_diag, const float *gh, float *gd,
                      const float *base, float *sum, void *stream) {
    DPISO_REQUIRE(batch >= 1 && h_tabT_u && h_tabT_v && values && a_diag && gh && gd && (base || !sum), "bad arguments");
    AdjTab tu = {h_tabT_u->n, h_tabT_u->wa, h_tabT_u->perm, h_tabT_u->a_col, h_tabT_u->a_src, h_tabT_u->r_col, h_tabT_u->r_src};
    AdjTab tv = {h_tabT_v->n, h_tabT_v->wa, h_tabT_v->perm, h_tabT_v->a_col, h_tabT_v->a_src, h_tabT_v->r_col, h_tabT_v->r_src};
    if (!tu.r_col || !tv.r_col || !tu.r_src || !tv.r_src) { tu.r_col = tv.r_col = nullptr; }
    const long long n = (long long)batch * (tu.n + tv.n);
    h_apply_adj_kernel<<<blocks_adj(n), kThreadsAdj, 0, (cudaStream_t)stream>>>(batch, tu, tv, nnz_u, nnz_v, beta,
                                                                                 values, a_diag, gh, gd, base, sum);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

int dpiso_predictor_rhs_adj(int batch, int ny, int nx, float dy, float dx, float beta, const uint8_t *dirichlet,
                            const float *grhs, const int *solve_stats, float *gvel, float *gforce, float *gdvals,
                            float *gfree, void *stream) {
    DPISO_REQUIRE(batch >= 1 && ny >= 3 && nx >= 3, "bad sizes");
    DPISO_REQUIRE(dirichlet && grhs && gvel && gfree, "null pointer");
    const int nf = ny * (nx + 1) + (ny + 1) * nx;
    predictor_rhs_adj_kernel<<<blocks_adj((long long)batch * nf), kThreadsAdj, 0, (cudaStream_t)stream>>>(
        batch, nf, cell_prod_adj(dy, dx), beta, dirichlet, grhs, solve_stats, gvel, gforce, gdvals, gfree);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

}  // extern "C"
