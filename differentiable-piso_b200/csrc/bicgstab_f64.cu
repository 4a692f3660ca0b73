// bicgstab_f64.cu -- the fp64 variant of the predictor solve: LinearSolverCudaMultiBicgstabILU(cast_to_double=True)
// (diffpiso/linear_solver.py:130-133 casts the fp32 matrix values and right-hand side to fp64, ":171" casts the solution
// back; launcher CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:540-988, the cusparseD / cublasD twin of the fp32 one).
//
// No shipped script of the reference selects it, so this is the plain design: one persistent CTA per (sample,
// component), level-major ELL tables (dpiso_bicg_tables: level_ptr, perm, a_col, a_src, a_rev), one named barrier per
// wavefront level, every vector and both factor planes in a caller-owned fp64 workspace.  The casts are fused: inputs and
// output stay fp32.  Same control flow as the fp32 kernels (lucky-guess test, two convergence tests per iteration, one
// restart from zero, NaN warning); sums accumulate in CSR (ascending column) order like the oracle's.
#include "bicgstab.cuh"

namespace dpiso {

struct F64Params {
    BicgTab tab[2];
    int nnz[2], n_face, nnz_total, n_max;
    size_t ws_doubles;     // per system
    const float *values, *rhs, *x0;
    float sign;
    float *x;
    int *stats;
    float *warn;
    double *workspace;
    float tol;
    int max_it;
};

// MODE 0: ILU(0) (a_val -> lu, pivots -> zs), 1: L solve zs = in - sum l * zs, 2: U solve zs = (zs - sum u * zs) / u_ii
template <int MODE>
__device__ void wavefront_f64(const BicgTab &T, int n_max, const float *__restrict__ values_c, double sign, const double *a_val,
                              double *lu, const double *in, double *zs) {
    const int wa = T.wa, n = T.n, nl = T.n_levels;
    __syncthreads();
    for (int s = 0; s < nl; s++) {
        const int d = MODE == 2 ? nl - 1 - s : s;
        const int q0 = T.level_ptr[d], q1 = T.level_ptr[d + 1];
        for (int q = q0 + (int)threadIdx.x; q < q1; q += (int)blockDim.x) {
            if (MODE == 0) {
                double diag = 0.0;
                int dslot = 0;
                for (int k = 0; k < wa; k++)
                    if (T.a_col[k * n + q] == q && T.a_src[k * n + q] >= 0) { dslot = k; diag = a_val[(size_t)k * n_max + q]; }
                for (int k = 0; k < wa; k++) {
                    const int col = T.a_col[k * n + q];
                    const double a = a_val[(size_t)k * n_max + q];
                    if (col < q) {
                        const double lik = a / zs[col];
                        lu[(size_t)k * n_max + q] = lik;
                        const int rev = T.a_rev[k * n + q];
                        if (rev >= 0) diag = fma(-lik, sign * (double)values_c[rev], diag);
                    } else if (k != dslot) {
                        lu[(size_t)k * n_max + q] = a;                 // U entries are unchanged by ILU(0) on this pattern
                    }
                }
                lu[(size_t)dslot * n_max + q] = diag;
                zs[q] = diag;
            } else if (MODE == 1) {
                double acc = in[q];
                for (int k = 0; k < wa; k++) {
                    const int col = T.a_col[k * n + q];
                    if (col < q) acc = fma(-lu[(size_t)k * n_max + q], zs[col], acc);
                }
                zs[q] = acc;
            } else {
                double acc = zs[q], dg = 1.0;
                for (int k = 0; k < wa; k++) {
                    const int col = T.a_col[k * n + q];
                    const double l = lu[(size_t)k * n_max + q];
                    if (col > q) acc = fma(-l, zs[col], acc);
                    else if (col == q && T.a_src[k * n + q] >= 0) dg = l;
                }
                zs[q] = acc / dg;
            }
        }
        __syncthreads();                                             // one level per barrier (global-memory writes included)
    }
}

__global__ void __launch_bounds__(kBicgThreads, 1) bicgstab_f64_kernel(const F64Params prm) {
    __shared__ double red[64];
    const int sys = blockIdx.x, sample = sys >> 1, comp = sys & 1;
    const BicgTab &T = prm.tab[comp];
    const int n = T.n, wa = T.wa, n_max = prm.n_max;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int face_off = comp ? prm.tab[0].n : 0;
    const float *values_c = prm.values + (size_t)sample * prm.nnz_total + (comp ? prm.nnz[0] : 0);
    const int nnz_c = prm.nnz[comp];
    const float *rhs_g = prm.rhs + (size_t)sample * prm.n_face + face_off;
    const float *x0_g = prm.x0 + (size_t)sample * prm.n_face + face_off;
    float *x_g = prm.x + (size_t)sample * prm.n_face + face_off;
    const double sign = (double)prm.sign;

    double *ws = prm.workspace + (size_t)sys * prm.ws_doubles;
    double *a_val = ws, *lu = a_val + (size_t)kMaxWa * n_max;
    double *b = lu + (size_t)kMaxWa * n_max, *x = b + n_max, *r = x + n_max, *rh = r + n_max, *p = rh + n_max, *v = p + n_max,
           *tt = v + n_max, *zs = tt + n_max;

    double nv = 0.0, nb = 0.0;
    for (int i = tid; i < nnz_c; i += NT) { const double a = (double)values_c[i]; nv += a * a; }
    for (int q = tid; q < n; q += NT) {
        const int orig = T.perm[q];
        const double bq = (double)rhs_g[orig];                       // tf.cast(rhs, tf.float64)
        b[q] = bq; nb += bq * bq;
        x[q] = (double)x0_g[orig];
        for (int k = 0; k < wa; k++) {
            const int src = T.a_src[k * n + q];
            a_val[(size_t)k * n_max + q] = src >= 0 ? sign * (double)values_c[src] : 0.0;
        }
    }
    block_sum2(nv, nb, red);
    const int warn = (isnan(sqrt(nv)) || isnan(sqrt(nb))) ? 1 : 0;
    wavefront_f64<0>(T, n_max, values_c, sign, a_val, lu, nullptr, zs);

    auto precondition = [&](const double *src) {
        wavefront_f64<1>(T, n_max, nullptr, 1.0, nullptr, lu, src, zs);
        wavefront_f64<2>(T, n_max, nullptr, 1.0, nullptr, lu, nullptr, zs);
    };
    auto spmv_row = [&](const double *vec, int q) {                  // CsrmvEx row: fma in ascending column order
        double acc = 0.0;
        for (int k = 0; k < wa; k++)
            if (T.a_src[k * n + q] >= 0) acc = fma(a_val[(size_t)k * n_max + q], vec[T.a_col[k * n + q]], acc);
        return acc;
    };

    double alpha = 1., rho = 1., rhop = 1., omega = 1., beta, nrm_r = 0.;
    int it_count = 0, restarts = 0, exit_kind = 3;
    const double tol = (double)prm.tol;
    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
        for (int q = tid; q < n; q += NT) { const double rq = b[q] - spmv_row(x, q); r[q] = rq; s0 += rq * rq; }
        block_sum2(s0, s1, red);
        nrm_r = sqrt(s0);
        if (nrm_r < tol) { exit_kind = 0; break; }
        for (int q = tid; q < n; q += NT) { rh[q] = r[q]; p[q] = 0.0; v[q] = 0.0; }
        exit_kind = 3;
        double rho_next = s0;
        for (int it = 0; it < prm.max_it; it++) {
            it_count++;
            rhop = rho; rho = rho_next;
            beta = (rho / rhop) * (alpha / omega);
            for (int q = tid; q < n; q += NT) p[q] = __dadd_rn(__dmul_rn(beta, fma(-omega, v[q], p[q])), r[q]);
            precondition(p);
            s0 = 0.0; s1 = 0.0;
            for (int q = tid; q < n; q += NT) { const double vq = spmv_row(zs, q); v[q] = vq; s0 += rh[q] * vq; }
            block_sum2(s0, s1, red);
            alpha = rho / s0;
            s0 = 0.0; s1 = 0.0;
            for (int q = tid; q < n; q += NT) {
                x[q] = fma(alpha, zs[q], x[q]);
                const double rq = fma(-alpha, v[q], r[q]);
                r[q] = rq; s0 += rq * rq;
            }
            block_sum2(s0, s1, red);
            nrm_r = sqrt(s0);
            if (nrm_r < tol) { exit_kind = 1; break; }
            precondition(r);
            s0 = 0.0; s1 = 0.0;
            for (int q = tid; q < n; q += NT) { const double tq = spmv_row(zs, q); tt[q] = tq; s0 += tq * r[q]; s1 += tq * tq; }
            block_sum2(s0, s1, red);
            omega = s0 / s1;
            s0 = 0.0; s1 = 0.0;
            for (int q = tid; q < n; q += NT) {
                x[q] = fma(omega, zs[q], x[q]);
                const double rq = fma(-omega, tt[q], r[q]);
                r[q] = rq; s0 += rq * rq; s1 += rq * rh[q];
            }
            block_sum2(s0, s1, red);
            nrm_r = sqrt(s0);
            rho_next = s1;
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > tol * 100 || isnan(nrm_r)) {
            __syncthreads();
            for (int q = tid; q < n; q += NT) x[q] = 0.0;
            if (restart == 1) restarts = 2;
        } else break;
    }
    __syncthreads();
    for (int q = tid; q < n; q += NT) x_g[T.perm[q]] = (float)x[q];  // tf.cast(sol[3], tf.float32)
    if (tid == 0) {
        int *st = prm.stats + (size_t)sys * 4;
        st[0] = it_count; st[1] = restarts; st[2] = warn; st[3] = exit_kind;
        if (warn) *prm.warn = 1.0f;
    }
}

static void to_tab_f64(const dpiso_bicg_tables *h, BicgTab &t) {
    t.n = h->n; t.n_levels = h->n_levels; t.wa = h->wa; t.max_level = h->max_level; t.wl = h->wl; t.wu = h->wu;
    t.dx = h->dx; t.rows_ok = h->rows_ok;
    t.level_ptr = h->level_ptr; t.perm = h->perm; t.a_col = h->a_col; t.a_src = h->a_src; t.a_rev = h->a_rev;
    t.r_col = nullptr; t.r_src = nullptr; t.r_rev = nullptr; t.c_lsrc = nullptr; t.c_lrev = nullptr; t.c_usrc = nullptr;
    t.c_lfar = nullptr; t.c_ufar = nullptr; t.c_dsrc = nullptr; t.m_nbr = nullptr; t.m_lfar = nullptr; t.m_ufar = nullptr;
}

}  // namespace dpiso

using namespace dpiso;

extern "C" {

size_t dpiso_bicgstab_f64_workspace_bytes(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v) {
    const size_t n_max = (size_t)(h_tab_u->n > h_tab_v->n ? h_tab_u->n : h_tab_v->n);
    return (2 * (size_t)kMaxWa + 8) * n_max * sizeof(double);
}

int dpiso_bicgstab_ilu_f64(int batch, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int nnz_u, int nnz_v,
                           const float *values, int negate, const float *rhs, const float *x0, float tol, int max_it,
                           float *x, int *stats, float *warn, void *workspace, void *stream) {
    DPISO_REQUIRE(batch >= 1 && h_tab_u && h_tab_v, "bad arguments");
    DPISO_REQUIRE(values && rhs && x0 && x && stats && warn && workspace, "null pointer");
    DPISO_REQUIRE(h_tab_u->wa >= 1 && h_tab_u->wa <= kMaxWa && h_tab_v->wa >= 1 && h_tab_v->wa <= kMaxWa, "ELL width out of range");
    F64Params prm;
    to_tab_f64(h_tab_u, prm.tab[0]);
    to_tab_f64(h_tab_v, prm.tab[1]);
    prm.nnz[0] = nnz_u; prm.nnz[1] = nnz_v; prm.nnz_total = nnz_u + nnz_v;
    prm.n_face = h_tab_u->n + h_tab_v->n;
    prm.n_max = h_tab_u->n > h_tab_v->n ? h_tab_u->n : h_tab_v->n;
    prm.ws_doubles = dpiso_bicgstab_f64_workspace_bytes(h_tab_u, h_tab_v) / sizeof(double);
    prm.values = values; prm.rhs = rhs; prm.x0 = x0; prm.x = x; prm.stats = stats; prm.warn = warn;
    prm.sign = negate ? -1.0f : 1.0f;
    prm.workspace = (double *)workspace; prm.tol = tol; prm.max_it = max_it;
    DPISO_CUDA_TRY(cudaMemsetAsync(warn, 0, sizeof(float), (cudaStream_t)stream));
    bicgstab_f64_kernel<<<batch * 2, kBicgThreads, 0, (cudaStream_t)stream>>>(prm);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

}  // extern "C"
