// bicgstab.cuh -- declarations shared by the batched ILU(0)-BiCGStab kernels (bicgstab.cu: one CTA per system;
// bicgstab_band.cu: one thread-block cluster per system for large grids).
#pragma once
#include "rows.cuh"

namespace dpiso {

constexpr int kBicgThreads = 512;
constexpr int kMaxWa = 6;

struct BicgTab {
    int n, n_levels, wa, max_level, wl, wu, dx, rows_ok;
    const int *level_ptr, *perm, *a_col, *a_src, *a_rev, *r_col, *r_src, *r_rev;
    const int4 *c_lsrc, *c_lrev, *c_usrc;
    const int2 *c_lfar, *c_ufar;
    const int *c_dsrc;
    const int4 *m_nbr;
    const int2 *m_lfar, *m_ufar;
};

struct BicgParams {
    BicgTab tab[2];
    int nnz[2];            // CSR entries of component 0, 1
    int n_face;            // n_u + n_v
    int nnz_total;
    int n_max;             // max(n_u, n_v): plane stride inside the workspace
    int zs_in_smem;
    size_t ws_floats;      // per system
    int stage_rows;        // rows per stage buffer (0 = staged fast path disabled)
    int lp_cap;            // ints reserved for the level_ptr copy in smem
    int compact;           // rows have <= 4 lower and <= 4 upper entries: compact triangular sweeps
    int ring_depth;        // levels in flight in the cp.async ring (16, 8 or 2)
    int rows_kernel;       // 1: bicgstab_rows_kernel (row-major layout, one thread per grid row in the sweeps)
    int rows_threads;      // its sweep threads P = roundup32(max dy)
    int rows_lp_cap;       // ints reserved for its level_ptr copy in shared memory
    int dbg;               // profiling experiments only (DPISO_BICG_DBG): 1 skip level barrier, 2 skip refill, 4 skip recurrence
                           // (level-major kernel); 8 force the level-major kernel
    const float *values, *rhs, *x0;
    float sign;            // +1 / -1: the solve runs on sign * values (piso_tf.py:42 passes -M)
    float *x;
    int *stats;
    float *warn;
    float *pivots_out;     // optional [batch][n_face]: ILU(0) pivots of this solve (row-major kernel)
    const float *pivots_in; // optional [batch][n_face]: pivots of the other orientation -> no factorisation sweep
    int reuse_mask;        // bit c: component c takes pivots_in
    float *workspace;
    float tol;
    int max_it;
    long long *timing;     // optional [8] cycle counters of system 0 (debug / profiling), may be NULL
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two doubles, result broadcast to every thread (2 barriers)
__device__ __forceinline__ void block_sum2(double &a, double &b, double *scratch /* [64] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    a = warp_sum_d(a); b = warp_sum_d(b);
    __syncthreads();                      // scratch free (previous result consumed)
    if (lane == 0) { scratch[warp] = a; scratch[32 + warp] = b; }
    __syncthreads();
    double va = lane < nw ? scratch[lane] : 0.0, vb = lane < nw ? scratch[32 + lane] : 0.0;
    a = warp_sum_d(va); b = warp_sum_d(vb);
}

// a / b as the ILU(0) sweeps need it: most lower slots of a row are absent (zero coefficient), and __fdiv_rn sends a zero
// numerator down its slow path (FCHK flags the exponent; measured: 64 such divisions made an ILU step 3x longer than the
// rest of it).  0 * b has the sign of 0 / b, so the product is bit-identical for every finite non-zero b.
__device__ __forceinline__ float ilu_div(float a, float b) { return a == 0.0f ? __fmul_rn(a, b) : __fdiv_rn(a, b); }

__device__ __forceinline__ void named_bar(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct RowsPlanes {
    float4 *alow;      // [n] A lower values in canonical slots (kept: the SpMV reads them)
    float *adiag;      // [n] A diagonal (kept)
    float4 *lval;      // [n] l_ik (written by the ILU sweep)
    float4 *arv;       // [n] ILU only: reverse entries u_ki = A(k, i) of the lower slots (aliases rh, p, v, tt)
    const int2 *lfar;  // [n] positions of the two far lower slots, -1 = absent (static table, shared by all systems)
    float4 *uval;      // [n] upper values (unchanged by ILU(0) on this pattern)
    const int2 *ufar;  // [n]
    float *udiag;      // [n] pivots (written by the ILU sweep)
};

// level-major position of grid point (row t, column x); lp = level_ptr (shared-memory copy)
__device__ __forceinline__ int lm_pos(const int *lp, int dx, int t, int x) {
    const int L = x + t;
    return lp[L] + t - max(0, L - dx + 1);
}


// bicgstab_band.cu: cluster-per-system kernel (large grids).  Returns DPISO_OK after enqueueing the solve, DPISO_EUNSUPPORTED
// when the grid does not qualify (the caller then falls back to the one-CTA kernels).
int launch_bicgstab_band(BicgParams &prm, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int batch,
                         void *stream);

// bicgstab_tile.cu: register-tiled wavefront sweeps, one cluster per system (default kernel); same return convention.
int launch_bicgstab_tile(BicgParams &prm, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int batch,
                         void *stream);
size_t tile_workspace_floats(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v);

}  // namespace dpiso
