// bicgstab_band.cu -- ILU(0)-BiCGStab with one thread-block CLUSTER per (sample, component) system.
//
// Same algorithm, control flow and per-row arithmetic as bicgstab_rows_kernel (bicgstab.cu), i.e. the sequence of
// BicgstabIluLinearSolveLauncher (CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:233-411), for grids whose systems are
// too large for one CTA (BASELINE config #5: 1024^2 and 2048^2 unknowns per component, where one CTA per system left
// 16 of 148 SMs busy for 120-440 ms per solve):
//   * the grid rows of a system are split into bands of Rc rows, one band per CTA of the cluster; in the three kinds of
//     wavefront sweeps (ILU(0), L solve, U solve) thread j of a CTA owns grid row rank*Rc + j and walks along x;
//   * NOTHING in a sweep is synchronised by a barrier.  The x-neighbour operand is the thread's own previous result, the
//     y-neighbour operand the neighbouring lane's (one shuffle); the one value per step that crosses a warp -- or a CTA --
//     boundary is handed over as an 8-byte packet {value, sweep id} stored into the CONSUMER warp's inbox (shared memory
//     of the consumer's CTA: a plain store inside the CTA, a DSMEM store across CTAs).  Value and tag travel in one
//     atomic word, so the consumer needs no fence: it polls the tag one step ahead of its use.  A warp therefore trails
//     its predecessor by the 32 steps the wavefront needs plus the hand-over latency, and the whole sweep is a pipeline
//     of dx + dy + (#warps) * latency steps whatever the number of CTAs;
//   * periodic wrap operands need no table: the in-row wrap (x = xa takes the row's own value at x = xb) is a register,
//     the in-column wrap (row ya takes row yb's value of the same column) is one more inbox, pushed by row yb's thread
//     into the CTA that owns row ya.  The four numbers per sweep direction come from the table builder, which proves
//     that they describe every far entry of the pattern (dpiso_bicg_tables::band_ok);
//   * coefficient planes and vectors are stored level by level as in bicgstab_rows_kernel, so the lanes of a warp
//     stream consecutive addresses (per-thread cp.async ring, D steps deep);
//   * SpMVs, vector updates and dot products run on all threads of all CTAs of the cluster over contiguous chunks of the
//     level-major positions; dot products are completed through DSMEM (rank-ordered, bitwise identical in every CTA);
//     barrier.cluster publishes the global-memory vectors between the phases.
#include <cooperative_groups.h>

#include <type_traits>

#include "bicgstab.cuh"

namespace cg = cooperative_groups;

namespace dpiso {

constexpr int kBandMaxCluster = 16;

struct BandFar { int xa, xb, ya, yb; };   // -1 = absent

struct BandParams {
    BicgParams p;
    int C;                 // CTAs per system
    int Rc;                // sweep rows per CTA (multiple of 32, <= kBicgThreads)
    int dxmax;             // packets per inbox (>= dx of either component)
    int lp_cap;            // ints reserved for the level_ptr copy
    BandFar far_l[2], far_u[2];   // per component: lower (ILU, L solve) and upper (U solve) far operands
};

__device__ __forceinline__ uint32_t band_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t band_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// packet store into the shared memory of any CTA of the cluster (address from mapa) / packet load from the own CTA
__device__ __forceinline__ void st_packet(uint32_t cluster_addr, float v, unsigned tag) {
    const unsigned long long pk = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(cluster_addr), "l"(pk) : "memory");
}
__device__ __forceinline__ void st_packet_local(uint32_t cta_addr, float v, unsigned tag) {
    const unsigned long long pk = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.volatile.shared.b64 [%0], %1;" ::"r"(cta_addr), "l"(pk) : "memory");
}
__device__ __forceinline__ unsigned long long ld_packet(uint32_t cta_addr) {
    unsigned long long pk;
    asm volatile("ld.relaxed.cluster.shared::cta.b64 %0, [%1];" : "=l"(pk) : "r"(cta_addr) : "memory");
    return pk;
}

// Dynamic shared memory of a CTA, all derived from (D, Rc, dxmax) inside each function so that the compiler keeps the
// accesses in the shared address space (LDS / STS / LDGSTS, no generic loads on the per-step critical path):
//   ring   [D][Rc] float4 values | [D][Rc] float4 reverse values (ILU) | [D][Rc] float | [D][Rc] float
//   boxes  [Rc / 32][dxmax] warp inboxes | [dxmax] in-column wrap inbox (lower sweeps) | [dxmax] (U solve); 8-byte packets
//   lp     level offsets: position of (row t, column x) = lp[x + t] + t
struct BandCtx {
    int dx, dy, Rc, C, rank, dxmax, dbg;
};
template <int D> __device__ __forceinline__ size_t band_off_boxes(int Rc) { return (size_t)D * Rc * 40; }
template <int D> __device__ __forceinline__ size_t band_off_lp(int Rc, int dxmax) {
    return band_off_boxes<D>(Rc) + (size_t)((Rc >> 5) + 2) * dxmax * 8;
}

// One row of a sweep; operands are passed by value.  Absent slots carry a zero coefficient; their operand is replaced by
// a neutral finite value.  fma order = ascending column of the row (lower: [column wrap, y-neighbour, row wrap,
// x-neighbour], upper: [x-neighbour, row wrap, y-neighbour, column wrap]) -- identical to sweep_row_step in bicgstab.cu.
template <int MODE>
__device__ __forceinline__ float band_row_step(const float4 v, const float4 rv, float e, float start, float nb, float prev,
                                               bool has_col, float fcol, bool has_row, float frow, float4 &l_out) {
    if (MODE == 0) {
        const float p0 = has_col ? fcol : 1.0f, p2 = has_row ? frow : 1.0f;
        const float l0 = ilu_div(v.x, p0), l1 = ilu_div(v.y, nb), l2 = ilu_div(v.z, p2), l3 = ilu_div(v.w, prev);
        float dg = fmaf(-l0, rv.x, e);
        dg = fmaf(-l1, rv.y, dg);
        dg = fmaf(-l2, rv.z, dg);
        dg = fmaf(-l3, rv.w, dg);
        l_out = make_float4(l0, l1, l2, l3);
        return dg;
    } else if (MODE == 1) {
        const float f0 = has_col ? fcol : 0.0f, f1 = has_row ? frow : 0.0f;
        float acc = fmaf(-v.x, f0, e);
        acc = fmaf(-v.y, nb, acc);
        acc = fmaf(-v.z, f1, acc);
        return fmaf(-v.w, prev, acc);
    } else {
        const float f0 = has_row ? frow : 0.0f, f1 = has_col ? fcol : 0.0f;
        float acc = fmaf(-v.x, prev, start);
        acc = fmaf(-v.y, f0, acc);
        acc = fmaf(-v.z, nb, acc);
        acc = fmaf(-v.w, f1, acc);
        return __fdiv_rn(acc, e);
    }
}

// MODE 0: ILU(0) (writes lval, udiag), 1: L solve (ext = right-hand side, writes zs), 2: U solve (zs in place).
// sid = sweep id (tag of this sweep's packets); the caller separates sweeps by barrier.cluster.
//
// A warp is alone on its scheduler for most of a sweep, so the cost of a step is the latency of its instruction stream,
// not its throughput.  The loop is therefore software-pipelined by hand: the ring slot of step u + 1 is read into registers
// and the packets of step u + 1 are polled while step u computes, the level offset (warp-uniform) is one broadcast load,
// and the full-warp steady state (all 32 lanes inside the grid) runs without per-lane activity tests.
template <int MODE, int D>
__device__ __noinline__ void band_sweep(const BandCtx c, const RowsPlanes pl, const BandFar far, const float *ext, float *zs,
                                        unsigned sid) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int Rc = c.Rc, dx = c.dx, dy = c.dy;
    const int t = c.rank * Rc + tid, t0 = t - lane;
    if (tid < Rc && t0 < dy) {                                         // warps without rows take no part
        float4 *const r16a = (float4 *)smem_raw, *const r16b = r16a + D * Rc;
        float *const r4a = (float *)(r16b + D * Rc), *const r4b = r4a + D * Rc;
        const int *const lp = (const int *)(smem_raw + band_off_lp<D>(Rc, c.dxmax));
        const uint32_t inbox0 = band_smem_u32(smem_raw + band_off_boxes<D>(Rc));
        const uint32_t wrap_l = inbox0 + (uint32_t)((Rc >> 5) * c.dxmax) * 8u, wrap_u = wrap_l + (uint32_t)c.dxmax * 8u;
        const bool rowok = t < dy;
        constexpr bool kUp = MODE == 2;
        const float4 *gval = kUp ? pl.uval : (MODE == 0 ? pl.alow : pl.lval);
        const float *gext = MODE == 1 ? ext : (MODE == 0 ? pl.adiag : pl.udiag);
        const int Wc = Rc >> 5;
        const int nsteps = dx + 31;
        const int nl1 = dx + dy - 2;                                   // last level
        // lane l works on column x(u) at local step u (the wavefront inside the warp); the level is warp-uniform
        const int xoff = kUp ? dx - 1 + (31 - lane) : -lane;
        const int loff = kUp ? dx + 30 + t0 : t0;
        auto x_of = [&](int u) { return kUp ? xoff - u : xoff + u; };
        auto q_of = [&](int u) { const int L = kUp ? loff - u : loff + u; return lp[min(max(L, 0), nl1)] + t; };
        // consumer of the neighbouring warp's edge row / producer for the other neighbour
        const bool poller = kUp ? (lane == 31 && t0 + 32 < dy) : (lane == 0 && t0 > 0);
        const bool producer = kUp ? (lane == 0 && t > 0) : (lane == 31 && t + 1 < dy);
        const uint32_t my_inbox = inbox0 + (uint32_t)(w * c.dxmax) * 8u;
        uint32_t prod_addr = 0;
        bool prod_remote = false;
        if (producer) {
            const int tw = kUp ? w - 1 : w + 1;                        // consumer warp, possibly in the neighbouring CTA
            const int trank = tw < 0 ? c.rank - 1 : (tw >= Wc ? c.rank + 1 : c.rank);
            const int twl = tw < 0 ? Wc - 1 : (tw >= Wc ? 0 : tw);
            prod_remote = trank != c.rank;
            prod_addr = inbox0 + (uint32_t)(twl * c.dxmax) * 8u;
            if (prod_remote) prod_addr = band_mapa(prod_addr, (uint32_t)trank);
        }
        const bool prod_local = producer && !prod_remote;
        // in-column wrap: row ya consumes what row yb pushes into the inbox of ya's CTA
        const bool wrapc = rowok && t == far.ya, wrapp = rowok && far.ya >= 0 && t == far.yb;
        const bool warp_wrap = __any_sync(0xffffffffu, wrapc || wrapp);
        const uint32_t wrap_box = kUp ? wrap_u : wrap_l;
        const uint32_t wrapp_addr = wrapp ? band_mapa(wrap_box, (uint32_t)(far.ya / Rc)) : 0u;
        const bool full_warp = t0 + 31 < dy && !(c.dbg & 512);
        auto issue = [&](int u) {
            const int x = x_of(u);
            if (rowok && (unsigned)x < (unsigned)dx) {
                const int q = q_of(u), k = (u & (D - 1)) * Rc + tid;
                cp_async16(r16a + k, gval + q);
                if (MODE == 0) cp_async16(r16b + k, pl.arv + q);
                cp_async4(r4a + k, gext + q);
                if (MODE == 2) cp_async4(r4b + k, zs + q);
            }
            cp_async_commit();
        };
        float4 v_cur, rv_cur = make_float4(0.f, 0.f, 0.f, 0.f), v_nxt, rv_nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        float e_cur, s_cur = 0.0f, e_nxt, s_nxt = 0.0f;
        auto load_next = [&](int u) {                                  // ring slot of step u -> registers
            const int k = (u & (D - 1)) * Rc + tid;
            v_nxt = r16a[k];
            if (MODE == 0) rv_nxt = r16b[k];
            e_nxt = r4a[k];
            if (MODE == 2) s_nxt = r4b[k];
        };
        float prev = 1.0f, keep = 1.0f;                               // finite non-zero stand-ins for absent operands
        unsigned long long pk = 0ull, wk = 0ull;
        auto fetch_packets = [&](int u) {                              // packets of step u, ahead of their use
            const uint32_t xo = (uint32_t)min(max(x_of(u), 0), dx - 1) * 8u;
            if (poller) pk = ld_packet(my_inbox + xo);
            if (warp_wrap && wrapc) wk = ld_packet(wrap_box + xo);
        };
        // one step; kFull: every lane of the warp is inside the grid (no activity test)
        auto step = [&](int u, auto full_tag) {
            constexpr bool kFull = decltype(full_tag)::value;
            issue(u + D - 2);
            cp_async_wait<D - 3>();                                    // step u + 1 has landed
            load_next(u + 1);
            const int x = x_of(u);
            const bool act = kFull || (rowok && (unsigned)x < (unsigned)dx);
            float nb = kUp ? __shfl_down_sync(0xffffffffu, prev, 1) : __shfl_up_sync(0xffffffffu, prev, 1);
            const bool late = act && ((poller && (unsigned)(pk >> 32) != sid) || (warp_wrap && wrapc && (unsigned)(wk >> 32) != sid));
            if (__any_sync(0xffffffffu, late)) {                       // a packet has not arrived yet: wait for it
                if (poller && act) while ((unsigned)(pk >> 32) != sid) pk = ld_packet(my_inbox + (uint32_t)x * 8u);
                if (wrapc && act) while ((unsigned)(wk >> 32) != sid) wk = ld_packet(wrap_box + (uint32_t)x * 8u);
            }
            if (poller) nb = __uint_as_float((unsigned)pk);
            const float fcol = __uint_as_float((unsigned)wk);
            float4 lo;
            const float res = band_row_step<MODE>(v_cur, rv_cur, e_cur, s_cur, nb, prev, wrapc, fcol, x == far.xa, keep, lo);
            if (act) {
                const int q = q_of(u);
                if (MODE == 0) { pl.lval[q] = lo; pl.udiag[q] = res; }
                else zs[q] = res;
                if (x == far.xb) keep = res;
                if (prod_local) st_packet_local(prod_addr + (uint32_t)x * 8u, res, sid);
                if (prod_remote) st_packet(prod_addr + (uint32_t)x * 8u, res, sid);
                if (warp_wrap && wrapp) st_packet(wrapp_addr + (uint32_t)x * 8u, res, sid);
                prev = res;
            }
            fetch_packets(u + 1);
            v_cur = v_nxt; rv_cur = rv_nxt; e_cur = e_nxt; s_cur = s_nxt;
        };
#pragma unroll 1
        for (int u = 0; u < D - 2; u++) issue(u);
        cp_async_wait<D - 3>();                                        // step 0 has landed
        load_next(0);
        v_cur = v_nxt; rv_cur = rv_nxt; e_cur = e_nxt; s_cur = s_nxt;
        fetch_packets(0);
        int u = 0;
        const int u_full_lo = 31, u_full_hi = full_warp ? dx : 31;     // all lanes inside the grid for u in [31, dx)
#pragma unroll 1
        for (; u < u_full_lo; u++) step(u, std::false_type());
#pragma unroll 2
        for (; u < u_full_hi; u++) step(u, std::true_type());
#pragma unroll 1
        for (; u < nsteps; u++) step(u, std::false_type());
        cp_async_wait<0>();
    }
}

#define DPISO_BAND_TICK(slot)                                                        \
    do {                                                                             \
        if (prm.timing && blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long _now = clock64();                                        \
            atomicAdd((unsigned long long *)&prm.timing[slot], (unsigned long long)(_now - tick)); \
            tick = _now;                                                             \
        }                                                                            \
    } while (0)

template <int D>
__global__ void __launch_bounds__(kBicgThreads, 1) bicgstab_band_kernel(const BandParams bp) {
    long long tick = clock64();
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[64];
    __shared__ double s_part[2][kBandMaxCluster][2];
    const BicgParams &prm = bp.p;
    const int C = bp.C, Rc = bp.Rc;
    const int rank = (int)cluster.block_rank();
    const int sys = blockIdx.x / C;
    const int sample = sys >> 1, comp = sys & 1;
    const BicgTab &T = prm.tab[comp];
    const int n = T.n, n_max = prm.n_max, dx = T.dx, dy = T.n / T.dx;
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
    float *const zs = tt + n_max;                                  // the solve vector (global memory / L2)

    // dynamic shared memory: ring | warp inboxes | wrap inboxes | level offsets (see BandCtx)
    BandCtx c;
    c.dx = dx; c.dy = dy; c.Rc = Rc; c.C = C; c.rank = rank; c.dxmax = bp.dxmax; c.dbg = prm.dbg;
    unsigned long long *const boxes = (unsigned long long *)(smem_raw + band_off_boxes<D>(Rc));
    const int n_box = (Rc >> 5) + 2;
    int *const s_lp = (int *)(smem_raw + band_off_lp<D>(Rc, bp.dxmax));
    for (int k = tid; k < n_box * bp.dxmax; k += NT) boxes[k] = 0ull;        // tag 0 = no sweep
    for (int k = tid; k < dx + dy; k += NT) s_lp[k] = T.level_ptr[k] - max(0, k - dx + 1);     // level offsets
    cluster.sync();                                               // every CTA resident, inboxes cleared

    int rb = 0;
    // cluster-wide sums of (a, b): CTA sums, all-gather through DSMEM, rank-ordered total (bitwise identical everywhere);
    // the barrier also publishes the global-memory writes of the phase that ends here
    auto cluster_sum2 = [&](double &a, double &bsum) {
        block_sum2(a, bsum, s_red);
        if (tid < C) {
            double *dst = cluster.map_shared_rank(&s_part[rb][rank][0], tid);
            dst[0] = a; dst[1] = bsum;
        }
        cluster.sync();
        double sa = 0.0, sb = 0.0;
        for (int k = 0; k < C; k++) { sa += s_part[rb][k][0]; sb += s_part[rb][k][1]; }
        a = sa; bsum = sb;
        rb ^= 1;
    };
    // this CTA's chunk of the level-major positions in the SpMV / vector phases (16-byte aligned bounds)
    const int chunk = ((n + C - 1) / C + 3) & ~3;
    const int q_lo = min(n, rank * chunk), q_hi = min(n, q_lo + chunk);
    const int q4_hi = q_lo + ((q_hi - q_lo) & ~3);

    // ---- setup: canonical rows, level-major order, NaN guard (":245-256") --------------------------------------
    double nv = 0.0, nb = 0.0;
#pragma unroll 8
    for (int i = rank * NT + tid; i < nnz_c; i += C * NT) { const double a = values_c[i]; nv += a * a; }
    for (int t = rank * NW + warp; t < dy; t += C * NW) {
        for (int xx = lane; xx < dx; xx += 32) {
            const int i = t * dx + xx, q = (s_lp[xx + t] + t);
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
    cluster_sum2(nv, nb);
    const int warn = (isnan((float)sqrt(nv)) || isnan((float)sqrt(nb))) ? 1 : 0;
    DPISO_BAND_TICK(0);

    unsigned sid = 0;
    const BandFar far_l = bp.far_l[comp], far_u = bp.far_u[comp];
    // ---- ILU(0) (csrilu02, ":181-218") ---------------------------------------------------------------------
    if (prm.pivots_in && ((prm.reuse_mask >> comp) & 1)) {
        // factor reuse (see bicgstab_rows_kernel): l'_ik = m_ik / d_k from the pivots of the other orientation
        const float *d_in = prm.pivots_in + (size_t)sample * prm.n_face + face_off;
        for (int t = rank * NW + warp; t < dy; t += C * NW) {
            for (int xx = lane; xx < dx; xx += 32) {
                const int i = t * dx + xx, q = (s_lp[xx + t] + t);
                const float4 a = pl.alow[q];
                const int2 fc = T.c_lfar[i];
                const float p0 = fc.x >= 0 ? d_in[fc.x] : 1.0f, p2 = fc.y >= 0 ? d_in[fc.y] : 1.0f;
                const float p1 = i - dx >= 0 ? d_in[i - dx] : 1.0f, p3 = i >= 1 ? d_in[i - 1] : 1.0f;
                pl.lval[q] = make_float4(ilu_div(a.x, p0), ilu_div(a.y, p1), ilu_div(a.z, p2), ilu_div(a.w, p3));
                pl.udiag[q] = d_in[i];
            }
        }
        cluster.sync();
    } else {
        band_sweep<0, D>(c, pl, far_l, nullptr, zs, ++sid);
        cluster.sync();
        if (prm.pivots_out) {
            float *d_out = prm.pivots_out + (size_t)sample * prm.n_face + face_off;
            for (int t = rank * NW + warp; t < dy; t += C * NW)
                for (int xx = lane; xx < dx; xx += 32) d_out[t * dx + xx] = pl.udiag[(s_lp[xx + t] + t)];
        }
    }
    DPISO_BAND_TICK(1);

    auto precondition = [&](const float *src) {                      // zs = U^-1 L^-1 src   (csrsv2 x2, ":321-327")
        cluster.sync();                                              // src complete in global memory
        band_sweep<1, D>(c, pl, far_l, src, zs, ++sid);
        cluster.sync();
        band_sweep<2, D>(c, pl, far_u, nullptr, zs, ++sid);
        cluster.sync();
    };
    // CsrmvEx row from the canonical planes (see bicgstab_rows_kernel)
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
    auto ld4 = [](const float *a) { return *reinterpret_cast<const float4 *>(a); };
    auto st4 = [](float *a, const float4 val) { *reinterpret_cast<float4 *>(a) = val; };

    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        cluster.sync();                                              // x complete (setup / restart reset)
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int q = q_lo + tid; q < q_hi; q += NT) {                // r = b - A x  (":275-282")
            const float rq = __fsub_rn(b[q], spmv_row(x, q));
            r[q] = rq; s0 += (double)rq * rq;
        }
        cluster_sum2(s0, s1);
        nrm_r = (float)sqrt(s0);
        if (nrm_r < tol) { exit_kind = 0; break; }                   // lucky guess (":287-289")
        for (int q = q_lo + tid; q < q_hi; q += NT) { rh[q] = r[q]; p[q] = 0.0f; v[q] = 0.0f; }
        exit_kind = 3;
        float rho_next = (float)s0;                                  // r.rh with rh = r
        for (int it = 0; it < prm.max_it; it++) {
            it_count++;
            rhop = rho;
            rho = rho_next;
            beta = __fmul_rn(__fdiv_rn(rho, rhop), __fdiv_rn(alpha, omega));
            __syncthreads();                                         // first iteration: rh, p, v were set by other threads
#pragma unroll 2
            for (int q = q_lo + tid * 4; q < q4_hi; q += NT * 4) {   // p = r + beta (p - omega v)  (":315-317")
                const float4 vv = ld4(v + q), rr = ld4(r + q);
                float4 pp = ld4(p + q);
                pp.x = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.x, pp.x)), rr.x);
                pp.y = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.y, pp.y)), rr.y);
                pp.z = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.z, pp.z)), rr.z);
                pp.w = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.w, pp.w)), rr.w);
                st4(p + q, pp);
            }
            for (int q = q4_hi + tid; q < q_hi; q += NT) p[q] = __fadd_rn(__fmul_rn(beta, fmaf(-omega, v[q], p[q])), r[q]);
            DPISO_BAND_TICK(4);
            precondition(p);                                         // zs = p_hat
            DPISO_BAND_TICK(2);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = q_lo + tid; q < q_hi; q += NT) {
                const float vq = spmv_row(zs, q);
                v[q] = vq; s0 += (double)rh[q] * vq;
            }
            cluster_sum2(s0, s1);
            alpha = __fdiv_rn(rho, (float)s0);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 2
            for (int q = q_lo + tid * 4; q < q4_hi; q += NT * 4) {   // x += alpha p_hat ; r -= alpha v ; |r|
                const float4 zz = ld4(zs + q), vv = ld4(v + q);
                float4 xx = ld4(x + q), rr = ld4(r + q);
                xx.x = fmaf(alpha, zz.x, xx.x); xx.y = fmaf(alpha, zz.y, xx.y); xx.z = fmaf(alpha, zz.z, xx.z); xx.w = fmaf(alpha, zz.w, xx.w);
                rr.x = fmaf(-alpha, vv.x, rr.x); rr.y = fmaf(-alpha, vv.y, rr.y); rr.z = fmaf(-alpha, vv.z, rr.z); rr.w = fmaf(-alpha, vv.w, rr.w);
                st4(x + q, xx); st4(r + q, rr);
                s0 += (double)rr.x * rr.x; s0 += (double)rr.y * rr.y; s0 += (double)rr.z * rr.z; s0 += (double)rr.w * rr.w;
            }
            for (int q = q4_hi + tid; q < q_hi; q += NT) {
                x[q] = fmaf(alpha, zs[q], x[q]);
                const float rq = fmaf(-alpha, v[q], r[q]);
                r[q] = rq; s0 += (double)rq * rq;
            }
            cluster_sum2(s0, s1);
            nrm_r = (float)sqrt(s0);
            if (nrm_r < tol) { exit_kind = 1; break; }
            DPISO_BAND_TICK(4);
            precondition(r);                                         // zs = s_hat
            DPISO_BAND_TICK(2);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 4
            for (int q = q_lo + tid; q < q_hi; q += NT) {
                const float tq = spmv_row(zs, q);
                tt[q] = tq; s0 += (double)tq * r[q]; s1 += (double)tq * tq;
            }
            cluster_sum2(s0, s1);
            omega = __fdiv_rn((float)s0, (float)s1);
            s0 = 0.0; s1 = 0.0;
#pragma unroll 2
            for (int q = q_lo + tid * 4; q < q4_hi; q += NT * 4) {   // x += omega s_hat ; r -= omega t ; |r| ; r.rh
                const float4 zz = ld4(zs + q), t4 = ld4(tt + q), hh = ld4(rh + q);
                float4 xx = ld4(x + q), rr = ld4(r + q);
                xx.x = fmaf(omega, zz.x, xx.x); xx.y = fmaf(omega, zz.y, xx.y); xx.z = fmaf(omega, zz.z, xx.z); xx.w = fmaf(omega, zz.w, xx.w);
                rr.x = fmaf(-omega, t4.x, rr.x); rr.y = fmaf(-omega, t4.y, rr.y); rr.z = fmaf(-omega, t4.z, rr.z); rr.w = fmaf(-omega, t4.w, rr.w);
                st4(x + q, xx); st4(r + q, rr);
                s0 += (double)rr.x * rr.x; s0 += (double)rr.y * rr.y; s0 += (double)rr.z * rr.z; s0 += (double)rr.w * rr.w;
                s1 += (double)rr.x * hh.x; s1 += (double)rr.y * hh.y; s1 += (double)rr.z * hh.z; s1 += (double)rr.w * hh.w;
            }
            for (int q = q4_hi + tid; q < q_hi; q += NT) {
                x[q] = fmaf(omega, zs[q], x[q]);
                const float rq = fmaf(-omega, tt[q], r[q]);
                r[q] = rq; s0 += (double)rq * rq; s1 += (double)rq * rh[q];
            }
            cluster_sum2(s0, s1);
            nrm_r = (float)sqrt(s0);
            rho_next = (float)s1;
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > __fmul_rn(tol, 100.0f) || isnan(nrm_r)) {        // ":392-404"
            for (int q = q_lo + tid; q < q_hi; q += NT) x[q] = 0.0f;
            if (restart == 1) restarts = 2;
        } else break;
    }
    cluster.sync();                                                  // x complete; no CTA leaves while packets may be in flight
    DPISO_BAND_TICK(4);
    for (int t = rank * NW + warp; t < dy; t += C * NW)              // back to the caller's row order
        for (int xx = lane; xx < dx; xx += 32) x_g[t * dx + xx] = x[(s_lp[xx + t] + t)];
    if (rank == 0 && tid == 0) {
        int *st = prm.stats + (size_t)sys * 4;
        st[0] = it_count; st[1] = restarts; st[2] = warn; st[3] = exit_kind;
        if (warn) *prm.warn = 1.0f;
    }
}

static int g_band_cluster = 0;       // tuning override (0 = heuristic)

size_t band_smem_bytes(int D, int Rc, int dxmax, int lp_cap) {
    return (size_t)D * Rc * 40 + (size_t)((Rc >> 5) + 2) * dxmax * 8 + (size_t)lp_cap * 4;
}

template <int D>
static int launch_band(const BandParams &bp, int batch, size_t smem, void *stream) {
    auto kernel = bicgstab_band_kernel<D>;
    DPISO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (bp.C > 8) DPISO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * 2 * bp.C));
    cfg.blockDim = dim3(kBicgThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)bp.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, bp);
    if (e != cudaSuccess) {
        set_error("BiCGStab (cluster variant, %d CTAs per system) launch failed: %s", bp.C, cudaGetErrorString(e));
        return DPISO_ECUDA;
    }
    return DPISO_OK;
}

int launch_bicgstab_band(BicgParams &prm, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int batch,
                         void *stream) {
    if (!(h_tab_u->band_ok && h_tab_v->band_ok)) return DPISO_EUNSUPPORTED;
    const int dy_u = h_tab_u->n / h_tab_u->dx, dy_v = h_tab_v->n / h_tab_v->dx;
    const int dymax = dy_u > dy_v ? dy_u : dy_v;
    BandParams bp;
    bp.dxmax = ((h_tab_u->dx > h_tab_v->dx ? h_tab_u->dx : h_tab_v->dx) + 1) & ~1;
    const int n_levels = h_tab_u->n_levels > h_tab_v->n_levels ? h_tab_u->n_levels : h_tab_v->n_levels;
    bp.lp_cap = (n_levels + 1 + 3) & ~3;
    const size_t kBudget = 220 * 1024;
    auto rows_of = [&](int C) { return (((dymax + C - 1) / C) + 31) & ~31; };
    auto feasible = [&](int C, int D) {
        const int Rc = rows_of(C);
        return Rc <= kBicgThreads && band_smem_bytes(D, Rc, bp.dxmax, bp.lp_cap) <= kBudget;
    };
    int C = 0;
    for (int cand = 1; cand <= kBandMaxCluster; cand *= 2)
        if (feasible(cand, 8)) { C = cand; break; }
    if (!C) return DPISO_EUNSUPPORTED;
    // more CTAs per system (the SpMV / vector phases scale with them) only while the batch stays within half the GPU:
    // every CTA boundary adds a DSMEM hand-over to the sweep pipeline (measured 1024^2 x 8: 4 CTAs per system 14-16 ms,
    // 8 CTAs 18-20 ms)
    while (C * 2 <= 8 && feasible(C * 2, 8) && rows_of(C * 2) >= 64 && batch * 2 * C * 2 <= 74) C *= 2;
    if (g_band_cluster && feasible(g_band_cluster, 8)) C = g_band_cluster;
    bp.C = C; bp.Rc = rows_of(C);
    bp.p = prm;
    for (int k = 0; k < 2; k++) {
        const dpiso_bicg_tables *h = k ? h_tab_v : h_tab_u;
        bp.far_l[k] = {h->far[0], h->far[1], h->far[2], h->far[3]};
        bp.far_u[k] = {h->far[4], h->far[5], h->far[6], h->far[7]};
    }
    // ring depth: 16 steps in flight where shared memory allows (a step takes ~100 cycles, an L2 / DRAM round trip ~1000)
    const bool deep = feasible(C, 16) && !(prm.dbg & 32);
    const size_t smem = band_smem_bytes(deep ? 16 : 8, bp.Rc, bp.dxmax, bp.lp_cap);
    return deep ? launch_band<16>(bp, batch, smem, stream) : launch_band<8>(bp, batch, smem, stream);
}

}  // namespace dpiso

extern "C" int dpiso_bicgstab_set_band_cluster(int cluster) {
    dpiso::g_band_cluster = cluster;
    return DPISO_OK;
}
