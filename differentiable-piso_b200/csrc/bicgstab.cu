// bicgstab.cu -- batched ILU(0)-preconditioned BiCGStab for the velocity predictor and its adjoint.
//
// Re-design of BicgstabIluLinearSolveLauncher (CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:85-453), which drives
// cuSPARSE csrilu02 / csrsv2 / CsrmvEx and cuBLAS level-1 calls from the host (19 library launches and 6 blocking
// reductions per iteration, one host thread per velocity component).  Here ONE persistent CTA owns one
// (sample, component) system for the whole solve.  Two kernels implement the same sequence:
//   bicgstab_rows_kernel  row-major layout, one sweep thread per grid row, canonical entry slots, float4 vector phases;
//                         default whenever the solve vector + cp.async ring fit shared memory (see its header below)
//   bicgstab_kernel       level-major layout (described next); fallback for larger grids
// bicgstab_kernel:
//   * the matrix is re-packed once into a level-major ELL layout (rows ordered by the lx+ly wavefront, <= 6 entries
//     per row, column order = ascending original column so that every sum is accumulated in CSR order);
//   * ILU(0) and the four triangular solves per iteration sweep the wavefront levels inside the kernel; the vector
//     being solved for lives in shared memory, so the dependent chain of a level is smem -> FMA -> smem, and the
//     next level's coefficients are prefetched into registers before the level barrier (a named barrier over the
//     participating warps only);
//   * SpMVs read the preconditioned vector straight from shared memory; dot products / norms are accumulated in
//     fp64 per thread, reduced with warp shuffles and a shared-memory stage, and consumed on the device -- no scalar
//     ever returns to the host; convergence tests, the restart rule and the NaN warning follow the reference.
// The transposed solve of the adjoint uses tables built for A^T (the reference transposes with csr2csc and
// factorises again, ":113-134"); the kernels are identical.
#include <cstdlib>

#include "bicgstab.cuh"

namespace dpiso {

struct RowRegs {
    int q;                 // row (level-major position) handled by this thread in this level, -1 = none
    int col[kMaxWa];
    float val[kMaxWa];
    float aux[kMaxWa];     // MODE 0: upper entry u(col, q) paired with the k-th entry (0 when absent)
    float rhs;
};

// One wavefront sweep over the level-major rows.
//   MODE 0: ILU(0) factorisation  (in: a_val planes, out: lu planes, zs = pivots)
//   MODE 1: L solve   zs[q] = in[q] - sum_{col<q} lu*zs[col]
//   MODE 2: U solve   zs[q] = (zs[q] - sum_{col>q} lu*zs[col]) / lu_diag
template <int MODE>
__device__ __noinline__ void wavefront(const BicgTab &T, int n_max, const float *__restrict__ values_c, float sign, const float *a_val,
                          float *lu, const float *in, float *zs) {
    const int wa = T.wa, n = T.n;
    const int P = min((int)blockDim.x, (T.max_level + 31) & ~31);
    const int t = threadIdx.x;
    __syncthreads();
    if (t < P) {
        const int nl = T.n_levels;
        for (int s = 0; s < nl; s++) {
            const int d = MODE == 2 ? nl - 1 - s : s;
            const int q0 = T.level_ptr[d], q1 = T.level_ptr[d + 1];
            for (int q = q0 + t; q < q1; q += P) {
                if (MODE == 0) {
                    float diag = 0.0f;
                    int dslot = 0;
                    // pivot of this row first (needed as the fma accumulator), then the L entries in column order
                    for (int k = 0; k < wa; k++)
                        if (T.a_col[k * n + q] == q && T.a_src[k * n + q] >= 0) { dslot = k; diag = a_val[k * n_max + q]; }
                    for (int k = 0; k < wa; k++) {
                        const int col = T.a_col[k * n + q];
                        const float a = a_val[k * n_max + q];
                        if (col < q) {
                            const float lik = __fdiv_rn(a, zs[col]);
                            lu[k * n_max + q] = lik;
                            const int rev = T.a_rev[k * n + q];
                            if (rev >= 0) diag = fmaf(-lik, sign * values_c[rev], diag);
                        } else if (k != dslot) {
                            lu[k * n_max + q] = a;                    // U entries are unchanged by ILU(0) on this pattern
                        }
                    }
                    lu[dslot * n_max + q] = diag;
                    zs[q] = diag;
                } else if (MODE == 1) {
                    float acc = in[q];
                    for (int k = 0; k < wa; k++) {
                        const int col = T.a_col[k * n + q];
                        if (col < q) acc = fmaf(-lu[k * n_max + q], zs[col], acc);
                    }
                    zs[q] = acc;
                } else {
                    float acc = zs[q], dg = 1.0f;
                    for (int k = 0; k < wa; k++) {
                        const int col = T.a_col[k * n + q];
                        const float l = lu[k * n_max + q];
                        if (col > q) acc = fmaf(-l, zs[col], acc);
                        else if (col == q && T.a_src[k * n + q] >= 0) dg = l;
                    }
                    zs[q] = __fdiv_rn(acc, dg);
                }
            }
            named_bar(1, P);
        }
    }
    __syncthreads();
}

// Fast path: staged wavefront.  Levels are processed in chunks of K consecutive levels.  While the solver threads
// (the first P = roundup32(max_level) threads) sweep the levels of chunk c out of a shared-memory stage buffer, ALL
// threads of the CTA prefetch the coefficient rows of chunk c+1 from global memory into registers and drop them into the
// other stage buffer -- the L2 latency of the coefficient stream is hidden behind the recurrence, and the critical path
// of a level is  LDS (stage) -> LDS (zs gather) -> fma chain -> STS -> named barrier over the solver warps only.
// Requires max_level <= blockDim.x and stage_rows >= max_level.
// Stage buffer b (0/1) holds, for up to stage_rows rows: col[kMaxWa][rows] (int), val[kMaxWa][rows], aux[kMaxWa][rows]
// (MODE 0: u(col,row); MODE 1: aux[0][.] = right-hand side).  All shared-memory pointers are derived from the extern
// array inside this function so that the compiler keeps them in the shared address space (LDS/STS, no generic or
// local-memory traffic on the per-level critical path).
template <int MODE, bool kZsSmem>
__device__ __noinline__ void wavefront_staged(const BicgTab &T, int lp_cap, int n_max,
                                 const float *__restrict__ plane_a /* a_val (MODE 0) or lu */,
                                 const float *__restrict__ plane_rv /* MODE 0: M(col,row) planes */, float *lu,
                                 const float *in, float *zs_global, int stage_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int *lp = (const int *)smem_raw;
    int *const stage0 = (int *)smem_raw + lp_cap;
    const int sbuf = stage_rows * (3 * kMaxWa);          // ints per stage buffer
    float *const zs = kZsSmem ? (float *)(stage0 + 2 * sbuf) : zs_global;
    const int wa = T.wa, n = T.n, nl = T.n_levels;
    const int t = threadIdx.x;
    const int P = (T.max_level + 31) & ~31;
    const int K = max(1, stage_rows / T.max_level);
    const int nchunks = (nl + K - 1) / K;
    // chunk c covers levels [la, lb) (ascending sweeps) or levels lb-1 .. la (descending sweep, MODE 2)
    auto chunk_levels = [&](int c, int &la, int &lb) {
        if (MODE != 2) { la = c * K; lb = min(nl, la + K); }
        else { lb = nl - c * K; la = max(0, lb - K); }
    };
    int reg_col[kMaxWa];
    float reg_val[kMaxWa], reg_aux[kMaxWa];
    int reg_q = -1;
    auto fetch = [&](int c) {            // global -> registers, one row per thread (rows of a chunk <= stage_rows <= NT)
        reg_q = -1;
        if (c < nchunks) {
            int la, lb;
            chunk_levels(c, la, lb);
            const int q = lp[la] + t;
            if (q < lp[lb]) {
                reg_q = q;
#pragma unroll
                for (int k = 0; k < kMaxWa; k++) {
                    const bool on = k < wa;
                    reg_col[k] = on ? T.a_col[k * n + q] : -1;
                    reg_val[k] = on ? plane_a[k * n_max + q] : 0.0f;
                    if (MODE == 0) reg_aux[k] = on ? plane_rv[k * n_max + q] : 0.0f;
                }
                if (MODE == 1) reg_aux[0] = in[q];
            }
        }
    };
    auto stash = [&](int c) {            // registers -> stage buffer of chunk c
        if (reg_q >= 0) {
            int la, lb;
            chunk_levels(c, la, lb);
            const int i = reg_q - lp[la];
            int *sc = stage0 + (c & 1) * sbuf;
            float *sv = (float *)(sc + stage_rows * kMaxWa), *sa = (float *)(sc + 2 * stage_rows * kMaxWa);
#pragma unroll
            for (int k = 0; k < kMaxWa; k++) {
                sc[k * stage_rows + i] = reg_col[k];
                sv[k * stage_rows + i] = reg_val[k];
                if (MODE == 0) sa[k * stage_rows + i] = reg_aux[k];
            }
            if (MODE == 1) sa[i] = reg_aux[0];
        }
    };
    __syncthreads();
    fetch(0);
    stash(0);
    __syncthreads();
    for (int c = 0; c < nchunks; c++) {
        fetch(c + 1);                                     // loads in flight while the solver threads sweep chunk c
        if (t < P) {
            int la, lb;
            chunk_levels(c, la, lb);
            const int q_base = lp[la];
            const int *sc = stage0 + (c & 1) * sbuf;
            const float *sv = (const float *)(sc + stage_rows * kMaxWa), *sa = (const float *)(sc + 2 * stage_rows * kMaxWa);
            for (int s = 0; s < lb - la; s++) {
                const int d = MODE == 2 ? lb - 1 - s : la + s;
                const int q = lp[d] + t;
                if (q < lp[d + 1]) {
                    const int i = q - q_base;
                    int col[kMaxWa];
                    float val[kMaxWa];
#pragma unroll
                    for (int k = 0; k < kMaxWa; k++) { col[k] = sc[k * stage_rows + i]; val[k] = sv[k * stage_rows + i]; }
                    if (MODE == 0) {
                        // pivot = first self entry in column order (padding, also col == q, sits behind the real entries)
                        float diag = 0.0f;
                        int dslot = -1;
#pragma unroll
                        for (int k = 0; k < kMaxWa; k++)
                            if (col[k] == q && dslot < 0) { dslot = k; diag = val[k]; }
#pragma unroll
                        for (int k = 0; k < kMaxWa; k++) {
                            if (k < wa) {
                                if (col[k] >= 0 && col[k] < q) {
                                    const float lik = __fdiv_rn(val[k], zs[col[k]]);
                                    lu[k * n_max + q] = lik;
                                    diag = fmaf(-lik, sa[k * stage_rows + i], diag);
                                } else if (k != dslot) {
                                    lu[k * n_max + q] = val[k];   // U entries are unchanged by ILU(0) on this pattern
                                }
                            }
                        }
                        lu[dslot * n_max + q] = diag;
                        zs[q] = diag;
                    } else if (MODE == 1) {
                        float acc = sa[i];
#pragma unroll
                        for (int k = 0; k < kMaxWa; k++)
                            if (col[k] >= 0 && col[k] < q) acc = fmaf(-val[k], zs[col[k]], acc);
                        zs[q] = acc;
                    } else {
                        float acc = zs[q], dg = 1.0f;
                        bool seen_diag = false;
#pragma unroll
                        for (int k = 0; k < kMaxWa; k++) {
                            if (col[k] > q) acc = fmaf(-val[k], zs[col[k]], acc);
                            else if (col[k] == q && !seen_diag) { dg = val[k]; seen_diag = true; }
                        }
                        zs[q] = __fdiv_rn(acc, dg);
                    }
                }
                named_bar(1, P);
            }
        }
        stash(c + 1);
        __syncthreads();
    }
}

// Triangular sweeps (MODE 1: L, MODE 2: U) over COMPACT rows with a per-thread cp.async ring.
//
// After ILU(0) every row is compacted once to the <= 4 entries each sweep needs (CompactPlanes, level-major, 16-byte
// vectors).  Only the solver warps (first P = roundup32(max_level) threads) run a sweep; thread t owns row lp[d] + t of
// level d.  Each thread streams ITS OWN rows D levels ahead of the recurrence with cp.async into a private shared-memory
// ring slot, so the L2 latency of the coefficient stream is hidden D levels deep without staging threads, chunk
// barriers or registers.  A level is then:  LDS ring slot -> LDS zs[col] x4 -> 4 fma (-> div) -> STS zs -> bar.sync
// over the solver warps.  Requires wl, wu <= 4 (host-checked) and max_level <= blockDim.x.
struct CompactPlanes {
    const int4 *lcol, *ucol;        // [n]
    const float4 *lval, *uval;      // [n]
    const float *udiag;             // [n]
};

template <int MODE, bool kZsSmem, int D>
__device__ __noinline__ void wavefront_ring(const BicgTab &T, int lp_cap, const CompactPlanes cp, const float *in,
                                            float *zs_global, int stage_rows, int dbg = 0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int *lp = (const int *)smem_raw;
    int *const stage0 = (int *)smem_raw + lp_cap;                      // ring lives in the stage-buffer region
    const int sbuf_full = stage_rows * (3 * kMaxWa);
    float *const zs = kZsSmem ? (float *)(stage0 + 2 * sbuf_full) : zs_global;
    const int nl = T.n_levels;
    const int t = threadIdx.x;
    const int P = (T.max_level + 31) & ~31;
    int4 *const rc = (int4 *)stage0;                                   // [D][P]
    float4 *const rv = (float4 *)(stage0 + 4 * D * P);                 // [D][P]
    float *const re = (float *)(stage0 + 8 * D * P);                   // [D][P]
    __syncthreads();
    if (t < P) {
        const int4 *gcol = MODE == 1 ? cp.lcol : cp.ucol;
        const float4 *gval = MODE == 1 ? cp.lval : cp.uval;
        const float *gext = MODE == 1 ? in : cp.udiag;
        auto row_of = [&](int s) {                                     // this thread's row in sweep step s, or -1
            if (s >= nl) return -1;
            const int d = MODE == 2 ? nl - 1 - s : s;
            const int q = lp[d] + t;
            return q < lp[d + 1] ? q : -1;
        };
        auto issue = [&](int s) {                                      // sweep step s -> ring slot s % D
            const int q = row_of(s);
            if (q >= 0) {
                const int slot = (s & (D - 1)) * P + t;
                cp_async16(rc + slot, gcol + q);
                cp_async16(rv + slot, gval + q);
                cp_async4(re + slot, gext + q);
            }
            cp_async_commit();
        };
        // D-1 steps in flight; at step s the slot consumed at step s-1 is refilled FIRST (its latency and the row lookup
        // of the next step overlap the recurrence), then the dependent chain of the level runs:
        //     LDS ring slot -> LDS zs[col] x4 -> 4 fma (-> div) -> STS -> bar.sync
#pragma unroll 1
        for (int s = 0; s < D - 1; s++) issue(s);
        int q = row_of(0);
#pragma unroll 1
        for (int s = 0; s < nl; s++) {
            if (!(dbg & 2)) issue(s + D - 1); else cp_async_commit();
            const int q_next = row_of(s + 1);
            cp_async_wait<D - 1>();                                    // this thread's row of step s has landed
            if (q >= 0 && !(dbg & 4)) {
                const int slot = (s & (D - 1)) * P + t;
                const int4 c4 = rc[slot];
                const float4 v4 = rv[slot];
                const float e = re[slot];
                float acc = MODE == 1 ? e : zs[q];
                acc = fmaf(-v4.x, zs[c4.x], acc);
                acc = fmaf(-v4.y, zs[c4.y], acc);
                acc = fmaf(-v4.z, zs[c4.z], acc);
                acc = fmaf(-v4.w, zs[c4.w], acc);
                zs[q] = MODE == 1 ? acc : __fdiv_rn(acc, e);
            }
            if (!(dbg & 1)) named_bar(1, P);
            q = q_next;
        }
        cp_async_wait<0>();
    }
    __syncthreads();
}

#define DPISO_TICK(slot)                                                             \
    do {                                                                             \
        if (prm.timing && blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long _now = clock64();                                        \
            atomicAdd((unsigned long long *)&prm.timing[slot], (unsigned long long)(_now - tick)); \
            tick = _now;                                                             \
        }                                                                            \
    } while (0)

__global__ void __launch_bounds__(kBicgThreads, 1) bicgstab_kernel(const BicgParams prm) {
    long long tick = clock64();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[64];
    const int sys = blockIdx.x;
    const int sample = sys >> 1, comp = sys & 1;
    const BicgTab &T = prm.tab[comp];
    const int n = T.n, wa = T.wa, n_max = prm.n_max;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int face_off = comp ? prm.tab[0].n : 0;
    const float *values_c = prm.values + (size_t)sample * prm.nnz_total + (comp ? prm.nnz[0] : 0);
    const int nnz_c = prm.nnz[comp];
    const float *rhs_g = prm.rhs + (size_t)sample * prm.n_face + face_off;
    const float *x0_g = prm.x0 + (size_t)sample * prm.n_face + face_off;
    float *x_g = prm.x + (size_t)sample * prm.n_face + face_off;

    float *ws = prm.workspace + (size_t)sys * prm.ws_floats;
    float *__restrict__ a_val = ws;                       // [kMaxWa][n_max]  M(row, col)
    float *__restrict__ lu = a_val + (size_t)kMaxWa * n_max;
    float *__restrict__ a_rv = lu + (size_t)kMaxWa * n_max;   // [kMaxWa][n_max]  M(col, row) (0 where absent): ILU(0) pivot updates
    float *__restrict__ b = a_rv + (size_t)kMaxWa * n_max;
    float *__restrict__ x = b + n_max;
    float *__restrict__ r = x + n_max;
    float *__restrict__ rh = r + n_max;
    float *__restrict__ p = rh + n_max;
    float *__restrict__ v = p + n_max;
    float *__restrict__ tt = v + n_max;
    const int *__restrict__ t_col = T.a_col;
    // compact L / U planes for the ring sweeps (16-byte aligned: n_max is padded to a multiple of 4 by the host)
    CompactPlanes cpl;
    {
        float *cbase = tt + 2 * (size_t)n_max;
        cpl.lcol = (const int4 *)cbase;
        cpl.lval = (const float4 *)(cbase + 4 * (size_t)n_max);
        cpl.ucol = (const int4 *)(cbase + 8 * (size_t)n_max);
        cpl.uval = (const float4 *)(cbase + 12 * (size_t)n_max);
        cpl.udiag = cbase + 16 * (size_t)n_max;
    }
    // dynamic smem: [level_ptr copy][stage buffers][zs]
    int *lp_s = (int *)smem_raw;
    float *const zs_smem = (float *)((int *)smem_raw + prm.lp_cap + (size_t)2 * prm.stage_rows * 3 * kMaxWa);
    float *const zs_glob = tt + n_max;
    float *zs = prm.zs_in_smem ? zs_smem : zs_glob;
    const bool fast = prm.stage_rows > 0;
    if (fast) for (int i = tid; i <= T.n_levels; i += NT) lp_s[i] = T.level_ptr[i];

    // ---- setup: permuted copies, NaN guard (":245-256") ------------------------------------------------------
    double nv = 0.0, nb = 0.0;
#pragma unroll 8
    for (int i = tid; i < nnz_c; i += NT) { const double a = values_c[i]; nv += a * a; }
    for (int q = tid; q < n; q += NT) {
        const int orig = T.perm[q];
        const float bq = rhs_g[orig];
        b[q] = bq; nb += (double)bq * bq;
        x[q] = x0_g[orig];                                           // cublasScopy(x_old -> x) (":261")
        for (int k = 0; k < wa; k++) {
            const int src = T.a_src[k * n + q], rev = T.a_rev[k * n + q];
            a_val[k * n_max + q] = src >= 0 ? prm.sign * values_c[src] : 0.0f;
            a_rv[k * n_max + q] = rev >= 0 ? prm.sign * values_c[rev] : 0.0f;
        }
    }
    block_sum2(nv, nb, red);
    int warn = (isnan((float)sqrt(nv)) || isnan((float)sqrt(nb))) ? 1 : 0;
    DPISO_TICK(0);

    // ---- ILU(0) (csrilu02, ":181-218") ---------------------------------------------------------------------
    const bool zsm = prm.zs_in_smem != 0;
    if (fast) {
        if (zsm) wavefront_staged<0, true>(T, prm.lp_cap, n_max, a_val, a_rv, lu, nullptr, zs_glob, prm.stage_rows);
        else wavefront_staged<0, false>(T, prm.lp_cap, n_max, a_val, a_rv, lu, nullptr, zs_glob, prm.stage_rows);
    } else wavefront<0>(T, n_max, values_c, prm.sign, a_val, lu, nullptr, zs);
    __syncthreads();
    DPISO_TICK(1);
    if (fast && zsm && prm.compact) {            // compact every row once: <= 4 lower entries, <= 4 upper entries + pivot
        for (int q = tid; q < n; q += NT) {
            int lc[4] = {q, q, q, q}, uc[4] = {q, q, q, q};
            float lv[4] = {0.f, 0.f, 0.f, 0.f}, uv[4] = {0.f, 0.f, 0.f, 0.f};
            float dg = 1.0f;
            int nl_ = 0, nu_ = 0;
            bool seen_diag = false;
            for (int k = 0; k < wa; k++) {
                const int col = T.a_col[k * n + q];
                const float val = lu[k * n_max + q];
                if (col < q) {
#pragma unroll
                    for (int m = 0; m < 4; m++) if (nl_ == m) { lc[m] = col; lv[m] = val; }
                    nl_++;
                } else if (col > q) {
#pragma unroll
                    for (int m = 0; m < 4; m++) if (nu_ == m) { uc[m] = col; uv[m] = val; }
                    nu_++;
                } else if (!seen_diag) { dg = val; seen_diag = true; }   // padding (also col == q) sits behind the pivot
            }
            ((int4 *)cpl.lcol)[q] = make_int4(lc[0], lc[1], lc[2], lc[3]);
            ((float4 *)cpl.lval)[q] = make_float4(lv[0], lv[1], lv[2], lv[3]);
            ((int4 *)cpl.ucol)[q] = make_int4(uc[0], uc[1], uc[2], uc[3]);
            ((float4 *)cpl.uval)[q] = make_float4(uv[0], uv[1], uv[2], uv[3]);
            ((float *)cpl.udiag)[q] = dg;
        }
        __syncthreads();
    }

    auto precondition = [&](const float *src) {                      // zs = U^-1 L^-1 src   (csrsv2 x2, ":321-327")
        if (fast && zsm && prm.compact) {
            if (prm.ring_depth == 16) {
                wavefront_ring<1, true, 16>(T, prm.lp_cap, cpl, src, zs_glob, prm.stage_rows, prm.dbg);
                wavefront_ring<2, true, 16>(T, prm.lp_cap, cpl, nullptr, zs_glob, prm.stage_rows, prm.dbg);
            } else if (prm.ring_depth == 8) {
                wavefront_ring<1, true, 8>(T, prm.lp_cap, cpl, src, zs_glob, prm.stage_rows, prm.dbg);
                wavefront_ring<2, true, 8>(T, prm.lp_cap, cpl, nullptr, zs_glob, prm.stage_rows, prm.dbg);
            } else {
                wavefront_ring<1, true, 2>(T, prm.lp_cap, cpl, src, zs_glob, prm.stage_rows, prm.dbg);
                wavefront_ring<2, true, 2>(T, prm.lp_cap, cpl, nullptr, zs_glob, prm.stage_rows, prm.dbg);
            }
        } else if (fast && zsm) {
            wavefront_staged<1, true>(T, prm.lp_cap, n_max, lu, nullptr, nullptr, src, zs_glob, prm.stage_rows);
            wavefront_staged<2, true>(T, prm.lp_cap, n_max, lu, nullptr, nullptr, nullptr, zs_glob, prm.stage_rows);
        } else if (fast) {
            wavefront_staged<1, false>(T, prm.lp_cap, n_max, lu, nullptr, nullptr, src, zs_glob, prm.stage_rows);
            wavefront_staged<2, false>(T, prm.lp_cap, n_max, lu, nullptr, nullptr, nullptr, zs_glob, prm.stage_rows);
        } else {
            wavefront<1>(T, n_max, nullptr, 1.0f, nullptr, lu, src, zs);
            wavefront<2>(T, n_max, nullptr, 1.0f, nullptr, lu, nullptr, zs);
        }
    };
    auto spmv_row = [&](const float *vec, int q) {                   // CsrmvEx row: fma in ascending column order
        float av[kMaxWa];
        int ac[kMaxWa];
#pragma unroll
        for (int k = 0; k < kMaxWa; k++) {                           // all loads first (memory-level parallelism)
            av[k] = k < wa ? a_val[k * n_max + q] : 0.0f;
            ac[k] = k < wa ? t_col[k * n + q] : q;
        }
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < kMaxWa; k++)
            if (k < wa) acc = fmaf(av[k], vec[ac[k]], acc);
        return acc;
    };

    float alpha = 1.f, rho = 1.f, rhop = 1.f, omega = 1.f, beta, nrm_r = 0.f;
    int it_count = 0, restarts = 0, exit_kind = 3;
    const float tol = prm.tol;

    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        // r = b - A x  (":275-282"); x is in global memory -> stage it through zs for the gather
        for (int q = tid; q < n; q += NT) zs[q] = x[q];
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int q = tid; q < n; q += NT) {
            const float rq = __fsub_rn(b[q], spmv_row(zs, q));
            r[q] = rq; s0 += (double)rq * rq;
        }
        block_sum2(s0, s1, red);
        nrm_r = (float)sqrt(s0);
        if (nrm_r < tol) { exit_kind = 0; break; }                   // lucky guess (":287-289")
        for (int q = tid; q < n; q += NT) { rh[q] = r[q]; p[q] = 0.0f; v[q] = 0.0f; }
        exit_kind = 3;
        float rho_next = (float)s0;                                  // r.rh with rh = r
        for (int it = 0; it < prm.max_it; it++) {
            it_count++;
            rhop = rho;
            rho = rho_next;
            beta = __fmul_rn(__fdiv_rn(rho, rhop), __fdiv_rn(alpha, omega));
            __syncthreads();
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {                      // p = r + beta (p - omega v)  (":315-317")
                float pq = fmaf(-omega, v[q], p[q]);
                pq = __fmul_rn(beta, pq);
                p[q] = __fadd_rn(pq, r[q]);
            }
            DPISO_TICK(4);
            precondition(p);                                         // zs = p_hat
            __syncthreads();
            DPISO_TICK(2);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {                      // v = A p_hat ; rh.v
                const float vq = spmv_row(zs, q);
                v[q] = vq; s0 += (double)rh[q] * vq;
            }
            block_sum2(s0, s1, red);
            alpha = __fdiv_rn(rho, (float)s0);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {                      // x += alpha p_hat ; r -= alpha v ; |r|
                x[q] = fmaf(alpha, zs[q], x[q]);
                const float rq = fmaf(-alpha, v[q], r[q]);
                r[q] = rq; s0 += (double)rq * rq;
            }
            block_sum2(s0, s1, red);
            nrm_r = (float)sqrt(s0);
            if (nrm_r < tol) { exit_kind = 1; break; }
            DPISO_TICK(4);
            precondition(r);                                         // zs = s_hat
            __syncthreads();
            DPISO_TICK(2);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {                      // t = A s_hat ; t.r ; t.t
                const float tq = spmv_row(zs, q);
                tt[q] = tq; s0 += (double)tq * r[q]; s1 += (double)tq * tq;
            }
            block_sum2(s0, s1, red);
            omega = __fdiv_rn((float)s0, (float)s1);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {                      // x += omega s_hat ; r -= omega t ; |r| ; r.rh
                x[q] = fmaf(omega, zs[q], x[q]);
                const float rq = fmaf(-omega, tt[q], r[q]);
                r[q] = rq; s0 += (double)rq * rq; s1 += (double)rq * rh[q];
            }
            block_sum2(s0, s1, red);
            nrm_r = (float)sqrt(s0);
            rho_next = (float)s1;
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > __fmul_rn(tol, 100.0f) || isnan(nrm_r)) {        // ":392-404"
            __syncthreads();
            for (int q = tid; q < n; q += NT) x[q] = 0.0f;
            if (restart == 1) restarts = 2;
        } else break;
    }
    __syncthreads();
    DPISO_TICK(4);
    for (int q = tid; q < n; q += NT) x_g[T.perm[q]] = x[q];
    if (tid == 0) {
        int *st = prm.stats + (size_t)sys * 4;
        st[0] = it_count; st[1] = restarts; st[2] = warn; st[3] = exit_kind;
        if (warn) *prm.warn = 1.0f;
    }
}


// ===============================================================================================================
// Row-sweep variant (default where the grid allows it).  Same algorithm and the same per-row arithmetic as
// bicgstab_kernel.  In the three kinds of sweeps (ILU(0), L solve, U solve) thread j owns grid row ly = j and walks
// along x as the level d = lx + ly advances, so
//   * the x-neighbour operand is the thread's own previous result (a register), the y-neighbour operand is the
//     neighbouring lane's previous result (one shuffle); only a warp-edge lane and the periodic wrap entries ("far") read
//     the solve vector in shared memory;
//   * entries sit in canonical slots by kind (lower: [far below the y-neighbour, y-neighbour, far above it,
//     x-neighbour], upper: [x-neighbour, far below the y-neighbour, y-neighbour, far above it]) -- the ascending-column
//     order of the row, hence the same fma chain as the CSR formulation (structure.py checks that every row fits);
//   * every plane and vector is stored LEVEL BY LEVEL (position q = level_ptr[lx + ly] + ly - ly_min(level), closed
//     form): at a given step the lanes of a warp touch consecutive positions, so the per-thread cp.async ring reads
//     coalesced 128-byte lines and the solve vector in shared memory is conflict-free.  (Round 1 kept the original row
//     order: each lane streamed its own grid row, i.e. 32 different lines per cp.async -- 3 x 32 L1 wavefronts per warp
//     and level, which is what held a level at ~380 cycles whatever was done to the barrier: profiles/r02_bicgstab.md.)
//     The caller's row-major vectors are permuted once on the way in and once on the way out.
// ===============================================================================================================
constexpr int kRowsRing = 8;
constexpr int kRowsSkew = 2;

// One row of a sweep, branch-free: absent slots carry a zero coefficient and a clamped position, so their operand is
// loaded unconditionally and selected afterwards.  fma order = ascending column of the row.
template <int MODE>
__device__ __forceinline__ float sweep_row_step(const RowsPlanes &pl, float *zs, int q, int qe, bool edge, float nb,
                                                float prev, const float4 v, const float4 rv, const int2 fc, float e) {
    const float z0 = zs[fc.x >= 0 ? fc.x : 0], z1 = zs[fc.y >= 0 ? fc.y : 0];
    const float ze = zs[qe];
    nb = edge ? ze : nb;
    if (MODE == 0) {
        // l_ik = a_ik / u_kk, u_ii = a_ii - sum l_ik u_ki, lower entries in ascending column order
        const float p0 = fc.x >= 0 ? z0 : 1.0f, p2 = fc.y >= 0 ? z1 : 1.0f;
        const float l0 = ilu_div(v.x, p0), l1 = ilu_div(v.y, nb), l2 = ilu_div(v.z, p2), l3 = ilu_div(v.w, prev);
        float dg = fmaf(-l0, rv.x, e);
        dg = fmaf(-l1, rv.y, dg);
        dg = fmaf(-l2, rv.z, dg);
        dg = fmaf(-l3, rv.w, dg);
        pl.lval[q] = make_float4(l0, l1, l2, l3);
        pl.udiag[q] = dg;
        return dg;
    } else if (MODE == 1) {
        const float f0 = fc.x >= 0 ? z0 : 0.0f, f1 = fc.y >= 0 ? z1 : 0.0f;
        float acc = fmaf(-v.x, f0, e);
        acc = fmaf(-v.y, nb, acc);
        acc = fmaf(-v.z, f1, acc);
        return fmaf(-v.w, prev, acc);
    } else {
        const float f0 = fc.x >= 0 ? z0 : 0.0f, f1 = fc.y >= 0 ? z1 : 0.0f;
        float acc = fmaf(-v.x, prev, zs[q]);
        acc = fmaf(-v.y, f0, acc);
        acc = fmaf(-v.z, nb, acc);
        acc = fmaf(-v.w, f1, acc);
        return __fdiv_rn(acc, e);
    }
}

// Progress counters of the decoupled sweeps.  Counter and solve vector both live in SHARED memory (kFlags is only used
// with the solve vector in shared memory): the LSU performs a warp's shared-memory instructions in issue order, so
// "STS values; (warp-converged) STS counter" on the producer and "LDS counter; LDS value" on the consumer need no fence --
// st.release / ld.acquire compile to MEMBAR.ALL.CTA per level, which also waits for the cp.async prefetches in flight
// (measured: 18 % slower than a per-level barrier).  volatile + the "memory" clobbers pin the compiler's order.
__device__ __forceinline__ int ld_flag_shared(const int *p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag_shared(int *p, int v) {
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// MODE 0: ILU(0) (zs = pivots), 1: L solve (ext = right-hand side), 2: U solve.
// kFlags = false: all sweep warps meet at a named barrier after every level.
// kFlags = true:  decoupled warps.  The only producer/consumer pair that crosses a warp is (last lane of warp w) ->
//   (first lane of warp w + 1) [L / ILU; mirrored for the U solve], so each warp starts kRowsSkew steps later than its
//   predecessor: the value an edge lane needs was then produced kRowsSkew + 1 steps ago and is handed over through the
//   solve vector plus a per-warp progress counter, polled one step ahead of its use.  A warp that runs ahead of schedule
//   never waits.  Far (periodic wrap) operands come from the same thread, from the same warp (ordered by the per-step
//   __syncwarp) or from a warp whose progress has been observed transitively: they lie in the same grid row or in the
//   same grid column at least two rows away (checked on the host: `rows_ok`), which keeps them at least (warp
//   distance) + 1 steps behind.
// ring = the kernel's dynamic shared memory: [D][P] slots of 11 floats per sweep thread; lp = level_ptr copy; zs = the
// solve vector (shared or global memory).
template <int MODE, int D, bool kFlags>
__device__ __noinline__ void sweep_rows(const RowsPlanes pl, const float *ext, int dx, int dy, int P, const int *lp, float *zs,
                                        int *progress /* shared, [kBicgThreads / 32] */) {
    constexpr int K = kFlags ? kRowsSkew : 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *const r16a = (float4 *)smem_raw;                           // [D][P] values
    float4 *const r16b = r16a + D * P;                                 // [D][P] ILU: reverse values
    int2 *const r8 = (int2 *)(r16b + D * P);                           // [D][P] far positions
    float *const r4 = (float *)(r8 + D * P);                           // [D][P] right-hand side / diagonal
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int Wn = P >> 5;
    if (kFlags && t < Wn) progress[t] = 0;
    __syncthreads();
    if (t < P) {
        const bool rowok = t < dy;
        const float4 *gval = MODE == 2 ? pl.uval : (MODE == 0 ? pl.alow : pl.lval);
        const int2 *gfar = MODE == 2 ? pl.ufar : pl.lfar;
        const float *gext = MODE == 1 ? ext : (MODE == 0 ? pl.adiag : pl.udiag);
        // first step of this thread: the wavefront offset of its row plus the warp skew
        const int sig = MODE == 2 ? (dy - 1 - t) + K * (Wn - 1 - w) : t + K * w;
        const int S = dx + dy - 1 + K * (Wn - 1);                      // steps of the sweep
        auto x_of = [&](int s) { return MODE == 2 ? dx - 1 - (s - sig) : s - sig; };
        // level of step s: the same for every lane of the warp, so its level_ptr entry is one broadcast read, kept in a
        // register ring (lpv) from the step that issues the prefetch to the steps that compute on it
        const int nl = dx + dy - 1;
        const int lev0 = MODE == 2 ? nl - 1 + K * (Wn - 1 - w) : -K * w;
        auto lev_of = [&](int s) { return MODE == 2 ? lev0 - s : lev0 + s; };
        int lpv[D];
#pragma unroll
        for (int u = 0; u < D; u++) lpv[u] = 0;
        auto issue = [&](int s, int slot) {
            const int x = x_of(s), L = lev_of(s);
            lpv[slot] = lp[min(max(L, 0), nl)];
            if (rowok && (unsigned)x < (unsigned)dx) {
                const int q = lpv[slot] + t - max(0, L - dx + 1), k = slot * P + t;
                cp_async16(r16a + k, gval + q);
                if (MODE == 0) cp_async16(r16b + k, pl.arv + q);
                cp_async8(r8 + k, gfar + q);
                cp_async4(r4 + k, gext + q);
            }
            cp_async_commit();
        };
        // the one lane of the warp that consumes a value of the neighbouring warp, and that warp's counter
        const int src_w = MODE == 2 ? w + 1 : w - 1;
        const bool has_src = MODE == 2 ? (src_w < Wn) : (src_w >= 0);
        const bool poller = kFlags && has_src && lane == (MODE == 2 ? 31 : 0);
        const bool edge = MODE == 2 ? (lane == 31 && t + 1 < dy) : (lane == 0 && t > 0);
        const int te = MODE == 2 ? t + 1 : t - 1;                      // row of the y-neighbour operand
        const int *src_prog = progress + (has_src ? src_w : w);
#pragma unroll
        for (int u = 0; u < D - 1; u++) issue(u, u);
        float prev = 1.0f;                                             // (a finite non-zero stand-in for absent operands)
#pragma unroll 1
        for (int s0 = 0; s0 < S; s0 += D) {
#pragma unroll
            for (int u = 0; u < D; u++) {
                const int s = s0 + u;
                if (s < S) {                                           // uniform per CTA
                    // progress of the neighbouring warp, read ahead of its use: steps [0, s - K) must be complete
                    int seen = poller ? ld_flag_shared(src_prog) : 0x7fffffff;
                    const int lp_edge = lpv[(u + D - 1) % D];          // level_ptr of the previous step's level
                    issue(s + D - 1, (u + D - 1) % D);
                    cp_async_wait<D - 1>();
                    float nb = MODE == 2 ? __shfl_down_sync(0xffffffffu, prev, 1) : __shfl_up_sync(0xffffffffu, prev, 1);
                    const int x = x_of(s);
                    if (kFlags) {
                        const int need = s - K < S ? s - K : S;
                        while (seen < need) seen = ld_flag_shared(src_prog);
                    }
                    float res = prev;
                    if (rowok && (unsigned)x < (unsigned)dx) {
                        const int L = lev_of(s), k = u * P + t;
                        const int q = lpv[u] + t - max(0, L - dx + 1);
                        const int Le = MODE == 2 ? L + 1 : L - 1;       // level of the y-neighbour operand = previous step's
                        const int qe = edge ? lp_edge + te - max(0, Le - dx + 1) : q;
                        const float4 v = r16a[k];
                        const float4 rv = MODE == 0 ? r16b[k] : make_float4(0.f, 0.f, 0.f, 0.f);
                        const int2 fc = r8[k];
                        const float e = r4[k];
                        res = sweep_row_step<MODE>(pl, zs, q, qe, edge, nb, prev, v, rv, fc, e);
                        zs[q] = res;
                    }
                    prev = res;
                    if (kFlags) {
                        __syncwarp();                                  // the warp's stores of this step precede the counter
                        if (lane == 0) st_flag_shared(progress + w, s + 1);
                    } else {
                        named_bar(1, P);
                    }
                }
            }
        }
        cp_async_wait<0>();
    }
    __syncthreads();
}

template <bool kZsSmem, int D>
__global__ void __launch_bounds__(kBicgThreads, 1) bicgstab_rows_kernel(const BicgParams prm) {
    long long tick = clock64();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[64];
    __shared__ int s_progress[kBicgThreads / 32];
    // decoupled-warp sweeps need the solve vector in shared memory; 16: per-level barrier sweeps (A/B measurements)
    const bool skew = kZsSmem && !(prm.dbg & 16);
    const int sys = blockIdx.x;
    const int sample = sys >> 1, comp = sys & 1;
    const BicgTab &T = prm.tab[comp];
    const int n = T.n, n_max = prm.n_max, dx = T.dx, dy = T.n / T.dx, P = prm.rows_threads;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const int face_off = comp ? prm.tab[0].n : 0;
    const float *values_c = prm.values + (size_t)sample * prm.nnz_total + (comp ? prm.nnz[0] : 0);
    const int nnz_c = prm.nnz[comp];
    const float *rhs_g = prm.rhs + (size_t)sample * prm.n_face + face_off;
    const float *x0_g = prm.x0 + (size_t)sample * prm.n_face + face_off;
    float *x_g = prm.x + (size_t)sample * prm.n_face + face_off;

    float *ws = prm.workspace + (size_t)sys * prm.ws_floats;
    RowsPlanes pl;
    float *cur = ws;
    pl.alow = (float4 *)cur;  cur += 4 * (size_t)n_max;
    pl.adiag = cur;           cur += n_max;
    pl.lval = (float4 *)cur;  cur += 4 * (size_t)n_max;
    pl.uval = (float4 *)cur;  cur += 4 * (size_t)n_max;
    pl.lfar = T.m_lfar;
    pl.ufar = T.m_ufar;
    pl.udiag = cur;           cur += n_max;
    float *__restrict__ b = cur;
    float *__restrict__ x = b + n_max;
    float *__restrict__ r = x + n_max;
    float *__restrict__ rh = r + n_max;                           // rh, p, v, tt: contiguous, double as the ILU-only arv plane
    float *__restrict__ p = rh + n_max;
    float *__restrict__ v = p + n_max;
    float *__restrict__ tt = v + n_max;
    pl.arv = (float4 *)rh;
    float *const zs_g = tt + n_max;                                        // global home of the solve vector (kZsSmem = false)
    // dynamic shared memory: the ring (11 floats per slot), the level_ptr copy, the solve vector
    int *const s_lp = (int *)((float *)smem_raw + (size_t)D * P * 11);
    float *const zs = kZsSmem ? (float *)(s_lp + prm.rows_lp_cap) : zs_g;
    for (int k = tid; k < dx + dy; k += NT) s_lp[k] = T.level_ptr[k];
    __syncthreads();

    // ---- setup: canonical rows, level-major order, NaN guard (":245-256") --------------------------------------
    double nv = 0.0, nb = 0.0;
#pragma unroll 8
    for (int i = tid; i < nnz_c; i += NT) { const double a = values_c[i]; nv += a * a; }
    // rows are visited in the caller's order (coalesced reads of the CSR values and slot tables) and scattered to their
    // level-major positions
    for (int t = warp; t < dy; t += NW) {
        for (int xx = lane; xx < dx; xx += 32) {
            const int i = t * dx + xx, q = lm_pos(s_lp, dx, t, xx);
            const float bi = rhs_g[i];
            b[q] = bi; nb += (double)bi * bi;
            x[q] = x0_g[i];                                           // cublasScopy(x_old -> x) (":261")
            const float sg = prm.sign;
            auto val4 = [&](const int4 s4) {
                return make_float4(s4.x >= 0 ? sg * values_c[s4.x] : 0.0f, s4.y >= 0 ? sg * values_c[s4.y] : 0.0f,
                                   s4.z >= 0 ? sg * values_c[s4.z] : 0.0f, s4.w >= 0 ? sg * values_c[s4.w] : 0.0f);
            };
            pl.alow[q] = val4(T.c_lsrc[i]);
            pl.arv[q] = val4(T.c_lrev[i]);
            pl.uval[q] = val4(T.c_usrc[i]);
            const int ds = T.c_dsrc[i];
            pl.adiag[q] = ds >= 0 ? sg * values_c[ds] : 1.0f;
        }
    }
    block_sum2(nv, nb, red);
    int warn = (isnan((float)sqrt(nv)) || isnan((float)sqrt(nb))) ? 1 : 0;
    DPISO_TICK(0);

    // ---- ILU(0) (csrilu02, ":181-218") ---------------------------------------------------------------------
    if (prm.pivots_in && ((prm.reuse_mask >> comp) & 1)) {
        // factor reuse: the pivots d of the solve with the other orientation of this matrix are given, and
        // ILU(0)(M^T) = (U^T D^-1)(D L^T): l'_ik = m_ik / d_k, upper entries unchanged, pivots d -- a parallel pass
        // instead of the wavefront sweep (same division as the sweep, MODE 0 of sweep_rows)
        const float *d_in = prm.pivots_in + (size_t)sample * prm.n_face + face_off;      // caller's row order
        __syncthreads();
        for (int t = warp; t < dy; t += NW) {
            for (int xx = lane; xx < dx; xx += 32) {
                const int i = t * dx + xx, q = lm_pos(s_lp, dx, t, xx);
                const float4 a = pl.alow[q];
                const int2 fc = T.c_lfar[i];
                const float p0 = fc.x >= 0 ? d_in[fc.x] : 1.0f, p2 = fc.y >= 0 ? d_in[fc.y] : 1.0f;
                const float p1 = i - dx >= 0 ? d_in[i - dx] : 1.0f, p3 = i >= 1 ? d_in[i - 1] : 1.0f;
                pl.lval[q] = make_float4(ilu_div(a.x, p0), ilu_div(a.y, p1), ilu_div(a.z, p2), ilu_div(a.w, p3));
                pl.udiag[q] = d_in[i];
            }
        }
        __syncthreads();
    } else {
        if (skew) sweep_rows<0, D, true>(pl, nullptr, dx, dy, P, s_lp, zs, s_progress);
        else sweep_rows<0, D, false>(pl, nullptr, dx, dy, P, s_lp, zs, s_progress);
        if (prm.pivots_out) {
            float *d_out = prm.pivots_out + (size_t)sample * prm.n_face + face_off;
            for (int t = warp; t < dy; t += NW)
                for (int xx = lane; xx < dx; xx += 32) d_out[t * dx + xx] = pl.udiag[lm_pos(s_lp, dx, t, xx)];
        }
    }
    DPISO_TICK(1);

    auto precondition = [&](const float *src) {                      // zs = U^-1 L^-1 src   (csrsv2 x2, ":321-327")
        if (skew) {
            sweep_rows<1, D, true>(pl, src, dx, dy, P, s_lp, zs, s_progress);
            sweep_rows<2, D, true>(pl, nullptr, dx, dy, P, s_lp, zs, s_progress);
        } else {
            sweep_rows<1, D, false>(pl, src, dx, dy, P, s_lp, zs, s_progress);
            sweep_rows<2, D, false>(pl, nullptr, dx, dy, P, s_lp, zs, s_progress);
        }
    };
    // CsrmvEx row from the canonical planes: lower slots, diagonal, upper slots = ascending column order; the positions
    // of the regular neighbours come from the static table m_nbr (absent slots carry a zero coefficient and point at the
    // row itself), the periodic wrap entries from the far tables.  One row per lane: coalesced loads.
    auto spmv_row = [&](const float *vec, int q) {
        const float4 lo = pl.alow[q], up = pl.uval[q];
        const float dg = pl.adiag[q];
        const int4 nq = T.m_nbr[q];                                  // x-1, y-1, x+1, y+1
        const int2 lf = pl.lfar[q], uf = pl.ufar[q];
        const float l0 = lf.x >= 0 ? vec[lf.x] : 0.0f, l2 = lf.y >= 0 ? vec[lf.y] : 0.0f;
        const float u1 = uf.x >= 0 ? vec[uf.x] : 0.0f, u3 = uf.y >= 0 ? vec[uf.y] : 0.0f;
        float acc = fmaf(lo.x, l0, 0.0f);
        acc = fmaf(lo.y, vec[nq.y], acc);
        acc = fmaf(lo.z, l2, acc);
        acc = fmaf(lo.w, vec[nq.x], acc);
        acc = fmaf(dg, vec[q], acc);
        acc = fmaf(up.x, vec[nq.z], acc);
        acc = fmaf(up.y, u1, acc);
        acc = fmaf(up.z, vec[nq.w], acc);
        acc = fmaf(up.w, u3, acc);
        return acc;
    };

    float alpha = 1.f, rho = 1.f, rhop = 1.f, omega = 1.f, beta, nrm_r = 0.f;
    int it_count = 0, restarts = 0, exit_kind = 3;
    const float tol = prm.tol;
    const int n4 = n & ~3;                                           // float4 part of the vector phases (planes are 16-byte aligned)
    auto ld4 = [](const float *a) { return *reinterpret_cast<const float4 *>(a); };
    auto st4 = [](float *a, const float4 val) { *reinterpret_cast<float4 *>(a) = val; };

    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        __syncthreads();
        for (int q = tid; q < n; q += NT) zs[q] = x[q];               // r = b - A x  (":275-282")
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int q = tid; q < n; q += NT) {
            const float rq = __fsub_rn(b[q], spmv_row(zs, q));
            r[q] = rq; s0 += (double)rq * rq;
        }
    
        block_sum2(s0, s1, red);
        nrm_r = (float)sqrt(s0);
        if (nrm_r < tol) { exit_kind = 0; break; }                   // lucky guess (":287-289")
        for (int q = tid; q < n; q += NT) { rh[q] = r[q]; p[q] = 0.0f; v[q] = 0.0f; }
        exit_kind = 3;
        float rho_next = (float)s0;                                  // r.rh with rh = r
        for (int it = 0; it < prm.max_it; it++) {
            it_count++;
            rhop = rho;
            rho = rho_next;
            beta = __fmul_rn(__fdiv_rn(rho, rhop), __fdiv_rn(alpha, omega));
            __syncthreads();
#pragma unroll 2
            for (int q = tid * 4; q < n4; q += NT * 4) {             // p = r + beta (p - omega v)  (":315-317")
                const float4 vv = ld4(v + q), rr = ld4(r + q);
                float4 pp = ld4(p + q);
                pp.x = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.x, pp.x)), rr.x);
                pp.y = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.y, pp.y)), rr.y);
                pp.z = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.z, pp.z)), rr.z);
                pp.w = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.w, pp.w)), rr.w);
                st4(p + q, pp);
            }
            for (int q = n4 + tid; q < n; q += NT) p[q] = __fadd_rn(__fmul_rn(beta, fmaf(-omega, v[q], p[q])), r[q]);
            DPISO_TICK(4);
            precondition(p);                                         // zs = p_hat
            DPISO_TICK(2);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {
                const float vq = spmv_row(zs, q);
                v[q] = vq; s0 += (double)rh[q] * vq;
            }
        
            block_sum2(s0, s1, red);
            alpha = __fdiv_rn(rho, (float)s0);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 2
            for (int q = tid * 4; q < n4; q += NT * 4) {             // x += alpha p_hat ; r -= alpha v ; |r|
                const float4 zz = ld4(zs + q), vv = ld4(v + q);
                float4 xx = ld4(x + q), rr = ld4(r + q);
                xx.x = fmaf(alpha, zz.x, xx.x); xx.y = fmaf(alpha, zz.y, xx.y); xx.z = fmaf(alpha, zz.z, xx.z); xx.w = fmaf(alpha, zz.w, xx.w);
                rr.x = fmaf(-alpha, vv.x, rr.x); rr.y = fmaf(-alpha, vv.y, rr.y); rr.z = fmaf(-alpha, vv.z, rr.z); rr.w = fmaf(-alpha, vv.w, rr.w);
                st4(x + q, xx); st4(r + q, rr);
                s0 += (double)rr.x * rr.x; s0 += (double)rr.y * rr.y; s0 += (double)rr.z * rr.z; s0 += (double)rr.w * rr.w;
            }
            for (int q = n4 + tid; q < n; q += NT) {
                x[q] = fmaf(alpha, zs[q], x[q]);
                const float rq = fmaf(-alpha, v[q], r[q]);
                r[q] = rq; s0 += (double)rq * rq;
            }
            block_sum2(s0, s1, red);
            nrm_r = (float)sqrt(s0);
            if (nrm_r < tol) { exit_kind = 1; break; }
            DPISO_TICK(4);
            precondition(r);                                         // zs = s_hat
            DPISO_TICK(2);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = tid; q < n; q += NT) {
                const float tq = spmv_row(zs, q);
                tt[q] = tq; s0 += (double)tq * r[q]; s1 += (double)tq * tq;
            }
        
            block_sum2(s0, s1, red);
            omega = __fdiv_rn((float)s0, (float)s1);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 2
            for (int q = tid * 4; q < n4; q += NT * 4) {             // x += omega s_hat ; r -= omega t ; |r| ; r.rh
                const float4 zz = ld4(zs + q), t4 = ld4(tt + q), hh = ld4(rh + q);
                float4 xx = ld4(x + q), rr = ld4(r + q);
                xx.x = fmaf(omega, zz.x, xx.x); xx.y = fmaf(omega, zz.y, xx.y); xx.z = fmaf(omega, zz.z, xx.z); xx.w = fmaf(omega, zz.w, xx.w);
                rr.x = fmaf(-omega, t4.x, rr.x); rr.y = fmaf(-omega, t4.y, rr.y); rr.z = fmaf(-omega, t4.z, rr.z); rr.w = fmaf(-omega, t4.w, rr.w);
                st4(x + q, xx); st4(r + q, rr);
                s0 += (double)rr.x * rr.x; s0 += (double)rr.y * rr.y; s0 += (double)rr.z * rr.z; s0 += (double)rr.w * rr.w;
                s1 += (double)rr.x * hh.x; s1 += (double)rr.y * hh.y; s1 += (double)rr.z * hh.z; s1 += (double)rr.w * hh.w;
            }
            for (int q = n4 + tid; q < n; q += NT) {
                x[q] = fmaf(omega, zs[q], x[q]);
                const float rq = fmaf(-omega, tt[q], r[q]);
                r[q] = rq; s0 += (double)rq * rq; s1 += (double)rq * rh[q];
            }
            block_sum2(s0, s1, red);
            nrm_r = (float)sqrt(s0);
            rho_next = (float)s1;
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > __fmul_rn(tol, 100.0f) || isnan(nrm_r)) {        // ":392-404"
            __syncthreads();
            for (int q = tid; q < n; q += NT) x[q] = 0.0f;
            if (restart == 1) restarts = 2;
        } else break;
    }
    __syncthreads();
    DPISO_TICK(4);
    for (int t = warp; t < dy; t += NW)                              // back to the caller's row order
        for (int xx = lane; xx < dx; xx += 32) x_g[t * dx + xx] = x[lm_pos(s_lp, dx, t, xx)];
    if (tid == 0) {
        int *st = prm.stats + (size_t)sys * 4;
        st[0] = it_count; st[1] = restarts; st[2] = warn; st[3] = exit_kind;
        if (warn) *prm.warn = 1.0f;
    }
}

}  // namespace dpiso

using namespace dpiso;

static long long *g_bicg_timing = nullptr;

static void to_tab(const dpiso_bicg_tables *h, BicgTab &t) {
    t.n = h->n; t.n_levels = h->n_levels; t.wa = h->wa; t.max_level = h->max_level; t.wl = h->wl; t.wu = h->wu;
    t.dx = h->dx; t.rows_ok = h->rows_ok;
    t.level_ptr = h->level_ptr; t.perm = h->perm; t.a_col = h->a_col; t.a_src = h->a_src; t.a_rev = h->a_rev;
    t.r_col = h->r_col; t.r_src = h->r_src; t.r_rev = h->r_rev;
    t.c_lsrc = (const int4 *)h->c_lsrc; t.c_lrev = (const int4 *)h->c_lrev; t.c_usrc = (const int4 *)h->c_usrc;
    t.c_lfar = (const int2 *)h->c_lfar; t.c_ufar = (const int2 *)h->c_ufar; t.c_dsrc = h->c_dsrc;
    t.m_nbr = (const int4 *)h->m_nbr; t.m_lfar = (const int2 *)h->m_lfar; t.m_ufar = (const int2 *)h->m_ufar;
}

extern "C" {

/* profiling hook: device buffer of 8 int64 cycle counters accumulated by system 0 (0 setup, 1 ILU(0), 2 triangular
 * sweeps, 4 streaming phases); NULL disables */
int dpiso_bicgstab_set_timing(long long *dev_counters) {
    g_bicg_timing = dev_counters;
    return DPISO_OK;
}

size_t dpiso_bicgstab_workspace_floats(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v) {
    const size_t n_max = (size_t)(h_tab_u->n > h_tab_v->n ? h_tab_u->n : h_tab_v->n);
    const size_t n_pad = (n_max + 3) & ~(size_t)3;
    const size_t base = (3 * (size_t)kMaxWa + 8) * n_pad + 17 * n_pad, tile = tile_workspace_floats(h_tab_u, h_tab_v);
    return base > tile ? base : tile;
}

static int g_reuse_always = 0;
static int g_bicg_dbg = -1;          // >= 0 overrides DPISO_BICG_DBG (tests force the level-major kernel with 8)

static bool rows_kernel_applies(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int *p_rows) {
    if (!(h_tab_u->rows_ok && h_tab_v->rows_ok && h_tab_u->c_lsrc && h_tab_v->c_lsrc && h_tab_u->m_nbr && h_tab_v->m_nbr &&
          h_tab_u->dx > 0 && h_tab_v->dx > 0))
        return false;
    const int dy_u = h_tab_u->n / h_tab_u->dx, dy_v = h_tab_v->n / h_tab_v->dx;
    const int Pr = ((dy_u > dy_v ? dy_u : dy_v) + 31) & ~31;
    const size_t ring = (size_t)kRowsRing * Pr * 11 * sizeof(float);       // the shallow ring must fit
    if (p_rows) *p_rows = Pr;
    return Pr <= kBicgThreads && ring <= 200 * 1024;
}

int dpiso_bicgstab_supports_factor_reuse(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v) {
    if (!h_tab_u || !h_tab_v) return 0;
    return (rows_kernel_applies(h_tab_u, h_tab_v, nullptr) || (h_tab_u->band_ok && h_tab_v->band_ok)) ? 1 : 0;
}

int dpiso_bicgstab_set_debug(int dbg) {
    g_bicg_dbg = dbg;
    return DPISO_OK;
}

int dpiso_bicgstab_set_reuse_policy(int always) {
    g_reuse_always = always ? 1 : 0;
    return DPISO_OK;
}

int dpiso_bicgstab_ilu(int batch, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int nnz_u,
                       int nnz_v, const float *values, int negate, const float *rhs, const float *x0, float tol,
                       int max_it, float *x, int *stats, float *warn, float *pivots_out, const float *pivots_in,
                       float *workspace, void *stream) {
    DPISO_REQUIRE(batch >= 1 && h_tab_u && h_tab_v, "bad arguments");
    DPISO_REQUIRE(values && rhs && x0 && x && stats && warn && workspace, "null pointer");
    DPISO_REQUIRE(h_tab_u->wa >= 1 && h_tab_u->wa <= kMaxWa && h_tab_v->wa >= 1 && h_tab_v->wa <= kMaxWa,
                  "ELL width out of range");
    BicgParams prm;
    to_tab(h_tab_u, prm.tab[0]);
    to_tab(h_tab_v, prm.tab[1]);
    prm.nnz[0] = nnz_u; prm.nnz[1] = nnz_v; prm.nnz_total = nnz_u + nnz_v;
    prm.n_face = h_tab_u->n + h_tab_v->n;
    prm.n_max = ((h_tab_u->n > h_tab_v->n ? h_tab_u->n : h_tab_v->n) + 3) & ~3;     // plane stride, 16-byte aligned
    prm.ws_floats = dpiso_bicgstab_workspace_floats(h_tab_u, h_tab_v);
    prm.values = values; prm.rhs = rhs; prm.x0 = x0; prm.x = x; prm.stats = stats; prm.warn = warn;
    prm.sign = negate ? -1.0f : 1.0f;
    prm.pivots_out = nullptr; prm.pivots_in = nullptr; prm.reuse_mask = 0;
    prm.workspace = workspace; prm.tol = tol; prm.max_it = max_it;
    prm.timing = g_bicg_timing;
    {   // profiling experiments only; read once per process
        static const int dbg_env = [] { const char *e = getenv("DPISO_BICG_DBG"); return e ? atoi(e) : 0; }();
        prm.dbg = g_bicg_dbg >= 0 ? g_bicg_dbg : dbg_env;
    }
    DPISO_CUDA_TRY(cudaMemsetAsync(warn, 0, sizeof(float), (cudaStream_t)stream));
    // shared memory plan: level_ptr copy + two stage buffers (staged wavefront) + the solve vector zs
    const size_t kBudget = 200 * 1024;
    const int max_level = h_tab_u->max_level > h_tab_v->max_level ? h_tab_u->max_level : h_tab_v->max_level;
    const int n_levels = h_tab_u->n_levels > h_tab_v->n_levels ? h_tab_u->n_levels : h_tab_v->n_levels;
    prm.lp_cap = (n_levels + 1 + 3) & ~3;
    const size_t zs_bytes = (size_t)prm.n_max * sizeof(float);
    const size_t row_bytes = (size_t)2 * 3 * kMaxWa * sizeof(int);      // both stage buffers, per staged row
    prm.stage_rows = 0;
    prm.zs_in_smem = 0;
    size_t smem = 0;
    if (max_level <= kBicgThreads) {
        size_t avail = kBudget - prm.lp_cap * sizeof(int);
        int rows = max_level * 4 < kBicgThreads ? max_level * 4 : kBicgThreads;     // aim at 4 levels per chunk
        if (rows < max_level) rows = max_level;
        if (zs_bytes + rows * row_bytes <= avail) prm.zs_in_smem = 1;
        else if (zs_bytes + max_level * row_bytes <= avail) { prm.zs_in_smem = 1; rows = (int)((avail - zs_bytes) / row_bytes); }
        if ((size_t)rows * row_bytes <= avail) {
            prm.stage_rows = rows;
            smem = prm.lp_cap * sizeof(int) + rows * row_bytes + (prm.zs_in_smem ? zs_bytes : 0);
        }
    }
    if (!prm.stage_rows) {
        prm.lp_cap = 0;
        prm.zs_in_smem = zs_bytes <= kBudget ? 1 : 0;
        smem = prm.zs_in_smem ? zs_bytes : 0;
    }
    prm.compact = (h_tab_u->wl <= 4 && h_tab_u->wu <= 4 && h_tab_v->wl <= 4 && h_tab_v->wu <= 4) ? 1 : 0;
    {
        const size_t region_ints = (size_t)2 * prm.stage_rows * 3 * kMaxWa;
        const size_t P = (size_t)((max_level + 31) & ~31);
        prm.ring_depth = region_ints >= 16 * P * 9 ? 16 : (region_ints >= 8 * P * 9 ? 8 : 2);
        if (region_ints < 2 * P * 9) prm.compact = 0;
    }
    // row-major kernel: every row in canonical slots, one sweep thread per grid row, ring + solve vector in shared memory
    prm.rows_kernel = 0;
    // Kernel choice.  Up to 512 rows per component: bicgstab_rows_kernel (one CTA per system, one sweep thread per grid
    // row: 1.2-1.3 ms per solve of 128 systems of 128^2).  Larger grids: one cluster per system -- bicgstab_band_kernel
    // (a row per sweep thread; 1024^2 x 8: 15-20 ms) up to ~1100 rows, bicgstab_tile_kernel (4 x 4 register tiles per sweep
    // step, sweep-image storage; 2048^2 x 4: 30 ms against 44 ms) beyond.  Debug bits force a kernel: 8 level-major,
    // 64 band, 128 rows, 256 tile.
    int Pr = 0;
    const bool rows_fit = rows_kernel_applies(h_tab_u, h_tab_v, &Pr);
    const int dy_u_ = h_tab_u->dx > 0 ? h_tab_u->n / h_tab_u->dx : 0, dy_v_ = h_tab_v->dx > 0 ? h_tab_v->n / h_tab_v->dx : 0;
    const bool tile_first = (prm.dbg & 256) || (!rows_fit && !(prm.dbg & (8 | 64 | 128)) && (dy_u_ > 1100 || dy_v_ > 1100));
    if (tile_first) {
        prm.pivots_out = pivots_out; prm.pivots_in = pivots_in;
        prm.reuse_mask = g_reuse_always ? 3 : ((h_tab_u->sym ? 1 : 0) | (h_tab_v->sym ? 2 : 0));
        const int rc = launch_bicgstab_tile(prm, h_tab_u, h_tab_v, batch, stream);
        if (rc != DPISO_EUNSUPPORTED) return rc;
        prm.pivots_out = nullptr; prm.pivots_in = nullptr; prm.reuse_mask = 0;
    }
    if ((prm.dbg & 64) && !(prm.dbg & 8)) {                     // A/B: the cluster-per-system kernel on a grid that fits one CTA
        prm.pivots_out = pivots_out; prm.pivots_in = pivots_in;
        prm.reuse_mask = g_reuse_always ? 3 : ((h_tab_u->sym ? 1 : 0) | (h_tab_v->sym ? 2 : 0));
        const int rc = launch_bicgstab_band(prm, h_tab_u, h_tab_v, batch, stream);
        if (rc != DPISO_EUNSUPPORTED) return rc;
        prm.pivots_out = nullptr; prm.pivots_in = nullptr; prm.reuse_mask = 0;
    }
    if (rows_fit && !(prm.dbg & 8)) {
        // the solve vector joins the ring in shared memory if it fits
        const size_t ring8 = (size_t)kRowsRing * Pr * 11 * sizeof(float), ring16 = 2 * ring8;
        prm.rows_lp_cap = (n_levels + 1 + 3) & ~3;
        const size_t zs_bytes_r = (size_t)prm.n_max * sizeof(float) + (size_t)prm.rows_lp_cap * sizeof(int);   // + level_ptr copy
        // 8 levels in flight cover the latency; 16 measured slower with coalesced streams (0.98 M vs 1.25 M cycles)
        int depth = 8;
        const bool zs_smem = ring8 + zs_bytes_r <= kBudget;         // else the solve vector stays in global memory (L2)
        if ((prm.dbg & 32) && ring16 + zs_bytes_r <= kBudget) depth = 16;               // A/B: deep ring
        const size_t ring = depth == 16 ? ring16 : ring8;
        const size_t need_smem = ring + (zs_smem ? zs_bytes_r : (size_t)prm.rows_lp_cap * sizeof(int));
        {
            prm.pivots_out = pivots_out; prm.pivots_in = pivots_in;
            prm.reuse_mask = g_reuse_always ? 3 : ((h_tab_u->sym ? 1 : 0) | (h_tab_v->sym ? 2 : 0));
            prm.rows_kernel = 1;
            prm.rows_threads = Pr;
            static unsigned long long rows_attr_mask = 0;
            int dev = 0;
            DPISO_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 64 || !(rows_attr_mask & (1ull << dev))) {
                DPISO_CUDA_TRY(cudaFuncSetAttribute(bicgstab_rows_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
                DPISO_CUDA_TRY(cudaFuncSetAttribute(bicgstab_rows_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
                DPISO_CUDA_TRY(cudaFuncSetAttribute(bicgstab_rows_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
                DPISO_CUDA_TRY(cudaFuncSetAttribute(bicgstab_rows_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
                if (dev < 64) rows_attr_mask |= 1ull << dev;
            }
            const dim3 grid(batch * 2), block(kBicgThreads);
            cudaStream_t st = (cudaStream_t)stream;
            if (zs_smem && depth == 16) bicgstab_rows_kernel<true, 16><<<grid, block, need_smem, st>>>(prm);
            else if (zs_smem) bicgstab_rows_kernel<true, 8><<<grid, block, need_smem, st>>>(prm);
            else if (depth == 16) bicgstab_rows_kernel<false, 16><<<grid, block, need_smem, st>>>(prm);
            else bicgstab_rows_kernel<false, 8><<<grid, block, need_smem, st>>>(prm);
            DPISO_CHECK_LAUNCH();
            return DPISO_OK;
        }
    }
    if (!(prm.dbg & 8)) {   // grids beyond one CTA's reach (BASELINE config #5): one cluster per system
        prm.pivots_out = pivots_out; prm.pivots_in = pivots_in;
        prm.reuse_mask = g_reuse_always ? 3 : ((h_tab_u->sym ? 1 : 0) | (h_tab_v->sym ? 2 : 0));
        const int rc = launch_bicgstab_band(prm, h_tab_u, h_tab_v, batch, stream);
        if (rc != DPISO_EUNSUPPORTED) return rc;
        prm.pivots_out = nullptr; prm.pivots_in = nullptr; prm.reuse_mask = 0;
    }
    {   // the attribute is per device: remember which devices of this process have it
        static unsigned long long attr_set_mask = 0;
        int dev = 0;
        DPISO_CUDA_TRY(cudaGetDevice(&dev));
        if (dev >= 64 || !(attr_set_mask & (1ull << dev))) {
            DPISO_CUDA_TRY(cudaFuncSetAttribute(bicgstab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
            if (dev < 64) attr_set_mask |= 1ull << dev;
        }
    }
    bicgstab_kernel<<<batch * 2, kBicgThreads, smem, (cudaStream_t)stream>>>(prm);
    DPISO_CHECK_LAUNCH();
    return DPISO_OK;
}

}  // extern "C"
