// bicgstab_tile.cu -- ILU(0)-BiCGStab, register-tiled wavefront sweeps, one thread-block cluster per system.
//
// Same algorithm, control flow and per-row arithmetic as the other predictor kernels (bicgstab.cu), i.e. the sequence of
// BicgstabIluLinearSolveLauncher (CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cu.cc:233-411: csrilu02 ":181-218", two
// csrsv2 per preconditioner application ":321-327", CsrmvEx, cuBLAS level 1), but the three kinds of triangular sweeps
// (ILU(0), L solve, U solve) are organised so that the dependent chain of a sweep is short:
//
//   * a sweep thread owns kTM = 4 consecutive grid rows and advances kTK = 4 columns per step, i.e. it computes a 4 x 4
//     tile of unknowns per step entirely in registers (the x-neighbour operand of a cell is the previous result of the
//     same row, the y-neighbour operand the result of the row above -- both registers).  Lane l trails lane l - 1 by one
//     step; the only values that cross lanes are the 4 results of a thread's last row (4 shuffles per step).  A warp
//     covers 128 grid rows, and a sweep over a dy x dx grid takes dx / 4 + dy / 4 steps of ~150 instructions instead of
//     dx + dy levels (bicgstab_rows_kernel: 255 levels of ~130 instructions on the 128^2 grid, now 65 steps);
//   * warps are chained without barriers: the 4 values per step that cross a warp -- or a CTA -- boundary are handed
//     over as 8-byte packets {value, sweep id} stored into the CONSUMER warp's inbox (shared memory of the consumer's
//     CTA: a plain store inside the CTA, a DSMEM store across the CTAs of the cluster).  Value and tag travel in one
//     atomic word, so no fence is needed; the consumer polls the tags one step ahead of their use;
//   * STORAGE = the image of the sweep.  Tile (tr, tc) of warp group g = tr / 32, lane l = tr % 32 is consumed at step
//     sigma = tc + l of the group's L sweep (and at step ntile + 30 - sigma of its U sweep), so every plane and every
//     vector is stored as blocks [g][sigma][item][lane] of 16-byte items: what the 32 lanes of a warp need in one step is
//     one contiguous block, every load / cp.async / store of a sweep is fully coalesced (512 bytes per instruction) and
//     lands conflict-free in the shared-memory ring, which has the same [item][lane] layout.  (A first version kept
//     row-major planes: 20 cp.async per step that each touched 32 different lines cost ~2600 cycles per step.)
//     Slots of a block whose lane is outside the grid at that step are never touched;
//   * there is no index table: the neighbours of a tile are the blocks sigma -+ 1 (same lane: x-neighbours; lane -+ 1:
//     y-neighbours), and the periodic wrap operands are four numbers per sweep direction (dpiso_bicg_tables::far, proven
//     by the table builder to describe every far entry of the pattern): the in-row wrap is a register per row, the
//     in-column wrap one more inbox pushed by the thread that owns the source row;
//   * SpMVs, vector updates and dot products run block-wise on all warps of all CTAs of the cluster (they are
//     element-wise in the image layout; the SpMV reads the neighbouring blocks); dot products are completed through DSMEM
//     (rank-ordered, bitwise identical in every CTA); barrier.cluster publishes the global-memory vectors between phases.
// Padded cells (x >= dx, rows >= dy inside a tile) carry zero off-diagonals, a unit diagonal and zero vector entries:
// they stay zero through every phase and contribute nothing to any sum.
#include <cooperative_groups.h>

#include "bicgstab.cuh"

namespace cg = cooperative_groups;

namespace dpiso {

constexpr int kTileMaxCluster = 16;
constexpr int kTM = 4, kTK = 4;                    // rows per sweep thread, columns per step
constexpr int kTileRingBytes = 48 * 1024;          // ring per sweep warp
constexpr int kTileMaxWarps = 4;                   // sweep warps per CTA
constexpr int kTileBlockFloats = 11264;            // workspace floats per block: 3 coefficient images + 10 vector images

struct TileFar { int xa, xb, ya, yb; };            // -1 = absent

struct TileParams {
    BicgParams p;
    int C, Wc;                 // CTAs per system, sweep warps per CTA
    int dxp[2], dyp[2];        // padded row length / row count (multiples of 4) per component
    int dxp_max;               // packets per inbox
    size_t blocks_max;         // blocks per image (max over the components): plane stride inside the workspace
    int use_tma;               // 1: sweeps fetch their blocks with TMA bulk copies
    TileFar far_l[2], far_u[2];
};

struct TilePlanes {
    float4 *alow, *uval, *lval, *arv;   // coefficient images [blocks][16][32]: canonical lower / upper slots of A, l_ik,
                                        // reverse entries (ILU only)
    float4 *adiag, *udiag;              // vector images [blocks][4][32]: diagonal of A, pivots
};

// geometry of one component's image
struct TileGeo {
    int dx, dy, ntile, ntr, G, S;       // faces per row / rows, tile columns / rows, warp groups, steps per sweep
};
__device__ __forceinline__ size_t vimg(int blk, int i, int lane) { return ((size_t)blk * 4 + i) * 32 + lane; }
__device__ __forceinline__ size_t cimg(int blk, int item, int lane) { return ((size_t)blk * 16 + item) * 32 + lane; }
// float index of cell (t, x) inside a vector image
__device__ __forceinline__ size_t vcell(const TileGeo &g, int t, int x) {
    const int tr = t >> 2, tc = x >> 2;
    return vimg((tr >> 5) * g.S + tc + (tr & 31), t & 3, tr & 31) * 4 + (x & 3);
}
__device__ __forceinline__ size_t ccell(const TileGeo &g, int t, int x) {
    const int tr = t >> 2, tc = x >> 2;
    return cimg((tr >> 5) * g.S + tc + (tr & 31), (t & 3) * 4 + (x & 3), tr & 31);
}

__device__ __forceinline__ uint32_t tile_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tile_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_packet_remote(uint32_t cluster_addr, float v, unsigned tag) {
    asm volatile("st.relaxed.cluster.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(cluster_addr), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_packet_cta(uint32_t cta_addr, float v, unsigned tag) {
    asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" ::"r"(cta_addr), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// two packets (16 bytes) of the own CTA's inbox
__device__ __forceinline__ uint4 ld_packets2(uint32_t cta_addr) {
    uint4 r;
    asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(cta_addr) : "memory");
    return r;
}
__device__ __forceinline__ void cp_async16_cg(uint32_t smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}

// ---- TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier).  A block of the sweep image is
// exactly what a warp consumes in one step, contiguous and 16-byte aligned: one elected lane fetches it with 2-3 bulk
// copies instead of 20-36 per-lane cp.async of the whole warp.
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct TileCtx {
    TileGeo g;
    int Wc, rank, dxp_max;
};

// One cell of a sweep; fma order = ascending column of the row (lower slots: [column wrap, y-neighbour, row wrap,
// x-neighbour], upper slots: [x-neighbour, row wrap, y-neighbour, column wrap]) -- the chain of sweep_row_step in bicgstab.cu.
// Absent slots carry a zero coefficient; their operand is any finite, non-zero value.
template <int MODE>
__device__ __forceinline__ float tile_cell(const float4 v, const float4 rv, float e, float start, float nb, float prev,
                                           float fcol, float frow, float4 &l_out) {
    if (MODE == 0) {
        const float l0 = ilu_div(v.x, fcol), l1 = ilu_div(v.y, nb), l2 = ilu_div(v.z, frow), l3 = ilu_div(v.w, prev);
        float dg = fmaf(-l0, rv.x, e);
        dg = fmaf(-l1, rv.y, dg);
        dg = fmaf(-l2, rv.z, dg);
        dg = fmaf(-l3, rv.w, dg);
        l_out = make_float4(l0, l1, l2, l3);
        return dg;
    } else if (MODE == 1) {
        float acc = fmaf(-v.x, fcol, e);
        acc = fmaf(-v.y, nb, acc);
        acc = fmaf(-v.z, frow, acc);
        return fmaf(-v.w, prev, acc);
    } else {
        float acc = fmaf(-v.x, prev, start);
        acc = fmaf(-v.y, frow, acc);
        acc = fmaf(-v.z, nb, acc);
        acc = fmaf(-v.w, fcol, acc);
        return __fdiv_rn(acc, e);
    }
}

template <int MODE> struct TileRing {
    // 16-byte items per ring slot and lane: 16 coefficient vectors, (ILU) 16 reverse vectors, per tile row one vector of
    // right-hand sides / diagonals / pivots, (U solve) one vector of the L solve's results
    static constexpr int kItems = MODE == 0 ? 36 : (MODE == 1 ? 20 : 24);
    static constexpr int kDepth = MODE == 0 ? 2 : 4;
    static_assert(kItems * 512 * kDepth <= kTileRingBytes, "ring does not fit");
};

// MODE 0: ILU(0) (writes lval, udiag), 1: L solve (ext = right-hand side image, writes zs), 2: U solve (zs in place).
// sid = sweep id (tag of this sweep's packets); the caller separates sweeps by barrier.cluster.
// kTma: the ring is filled by TMA bulk copies of whole blocks issued by lane 0 (slot k completes on mbarrier k of the warp;
// `phase` = the warp's mbarrier parity bits, carried from sweep to sweep); else by per-lane cp.async.
template <int MODE, bool kTma>
__device__ __noinline__ void tile_sweep(const TileCtx c, const TilePlanes pl, const TileFar far, const float4 *ext, float4 *zs,
                                        unsigned sid, unsigned &phase) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool kUp = MODE == 2;
    constexpr int NI = TileRing<MODE>::kItems, D = TileRing<MODE>::kDepth;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int Wc = c.Wc, ntile = c.g.ntile, ntr = c.g.ntr, G = c.g.G, S = c.g.S;
    const int g = c.rank * Wc + w;                                    // warp group of the system swept by this warp
    const int tr = g * 32 + lane;                                      // tile row of this thread
    if (w < Wc && g < G) {                                             // warps without rows take no part
        const uint32_t ring = tile_smem_u32(smem_raw) + (uint32_t)w * kTileRingBytes + (uint32_t)lane * 16u;
        const uint32_t inbox0 = tile_smem_u32(smem_raw) + (uint32_t)Wc * kTileRingBytes;
        const uint32_t wrap_l = inbox0 + (uint32_t)(Wc * c.dxp_max) * 8u, wrap_u = wrap_l + (uint32_t)c.dxp_max * 8u;
        const uint32_t bars = wrap_u + (uint32_t)c.dxp_max * 8u + (uint32_t)w * 32u;      // 4 mbarriers per sweep warp
        const bool rowok = tr < ntr;
        const float4 *gval = kUp ? pl.uval : (MODE == 0 ? pl.alow : pl.lval);
        const float4 *gext = MODE == 1 ? ext : (MODE == 0 ? pl.adiag : pl.udiag);
        const int nsteps = S;
        // block (g, sigma) is consumed at local step u = sigma (lower sweeps) / u = S - 1 - sigma (U solve)
        auto blk_of = [&](int u) { return g * S + (kUp ? S - 1 - u : u); };
        auto tc_of = [&](int u) { return (kUp ? S - 1 - u : u) - lane; };
        // consumer of the neighbouring warp's edge row / producer for the other neighbour
        const bool poller = kUp ? (lane == 31 && g + 1 < G) : (lane == 0 && g > 0);
        const bool producer = kUp ? (lane == 0 && g > 0) : (lane == 31 && tr + 1 < ntr);
        const uint32_t my_inbox = inbox0 + (uint32_t)(w * c.dxp_max) * 8u;
        uint32_t prod_addr = 0;
        bool prod_remote = false;
        if (producer) {
            const int tw = kUp ? w - 1 : w + 1;                        // consumer warp, possibly in the neighbouring CTA
            const int trank = tw < 0 ? c.rank - 1 : (tw >= Wc ? c.rank + 1 : c.rank);
            const int twl = tw < 0 ? Wc - 1 : (tw >= Wc ? 0 : tw);
            prod_remote = trank != c.rank;
            prod_addr = inbox0 + (uint32_t)(twl * c.dxp_max) * 8u;
            if (prod_remote) prod_addr = tile_mapa(prod_addr, (uint32_t)trank);
        }
        // in-column wrap: the thread that owns grid row ya consumes what the owner of row yb pushes into the inbox of ya's CTA
        const bool wrapc = rowok && far.ya >= 0 && (far.ya >> 2) == tr;
        const bool wrapp = rowok && far.ya >= 0 && (far.yb >> 2) == tr;
        const bool warp_wrap = __any_sync(0xffffffffu, wrapc || wrapp);
        const uint32_t wrap_box = kUp ? wrap_u : wrap_l;
        uint32_t wrapp_addr = 0;
        bool wrapp_remote = false;
        if (wrapp) {
            const int trank = (far.ya >> 7) / Wc;                      // group of row ya = (ya / 4) / 32
            wrapp_remote = trank != c.rank;
            wrapp_addr = wrapp_remote ? tile_mapa(wrap_box, (uint32_t)trank) : wrap_box;
        }
        const int i_yb = far.yb & 3;                                   // tile row pushed by the wrap producer
        const int xb_tc = far.xb >= 0 ? far.xb >> 2 : -1, xb_j = far.xb >= 0 ? far.xb & 3 : 0;

        auto issue = [&](int u) {
            if (kTma) {
                if (lane == 0 && u < nsteps) {                         // whole block, whatever lanes are inside the grid
                    const uint32_t slot = ring + (uint32_t)((u & (D - 1)) * NI) * 512u, bar = bars + (uint32_t)(u & (D - 1)) * 8u;   // lane 0: ring has no lane offset
                    const int blk = blk_of(u);
                    tma_mbar_expect_tx(bar, (uint32_t)NI * 512u);
                    tma_bulk_g2s(slot, gval + cimg(blk, 0, 0), 16u * 512u, bar);
                    if (MODE == 0) tma_bulk_g2s(slot + 16u * 512u, pl.arv + cimg(blk, 0, 0), 16u * 512u, bar);
                    tma_bulk_g2s(slot + (uint32_t)(MODE == 0 ? 32 : 16) * 512u, gext + vimg(blk, 0, 0), 4u * 512u, bar);
                    if (MODE == 2) tma_bulk_g2s(slot + 20u * 512u, zs + vimg(blk, 0, 0), 4u * 512u, bar);
                }
                return;
            }
            const int tc = tc_of(u);
            if (rowok && (unsigned)tc < (unsigned)ntile) {
                const uint32_t slot = ring + (uint32_t)((u & (D - 1)) * NI) * 512u;
                const int blk = blk_of(u);
                const float4 *cv = gval + cimg(blk, 0, lane), *ce = gext + vimg(blk, 0, lane);
#pragma unroll
                for (int k = 0; k < 16; k++) cp_async16_cg(slot + (uint32_t)k * 512u, cv + k * 32);
                if (MODE == 0) {
                    const float4 *cr = pl.arv + cimg(blk, 0, lane);
#pragma unroll
                    for (int k = 0; k < 16; k++) cp_async16_cg(slot + (uint32_t)(16 + k) * 512u, cr + k * 32);
                }
#pragma unroll
                for (int i = 0; i < kTM; i++) cp_async16_cg(slot + (uint32_t)((MODE == 0 ? 32 : 16) + i) * 512u, ce + i * 32);
                if (MODE == 2) {
                    const float4 *cz = zs + vimg(blk, 0, lane);
#pragma unroll
                    for (int i = 0; i < kTM; i++) cp_async16_cg(slot + (uint32_t)(20 + i) * 512u, cz + i * 32);
                }
            }
            cp_async_commit();
        };

        // persistent registers: last column of every tile row (x-neighbour operands of the next step), the row handed to
        // the neighbouring lane, the in-row wrap operands, the packets of the next step
        float side[kTM], edge[kTK], keep[kTM];
#pragma unroll
        for (int i = 0; i < kTM; i++) { side[i] = 1.0f; keep[i] = 1.0f; }   // finite non-zero stand-ins for absent operands
#pragma unroll
        for (int j = 0; j < kTK; j++) edge[j] = 1.0f;
        uint4 pk0 = make_uint4(0, 0, 0, 0), pk1 = pk0, wk0 = pk0, wk1 = pk0;
        auto fetch_packets = [&](int u) {
            const uint32_t xo = (uint32_t)min(max(tc_of(u), 0), ntile - 1) * (kTK * 8u);
            if (poller) { pk0 = ld_packets2(my_inbox + xo); pk1 = ld_packets2(my_inbox + xo + 16u); }
            if (warp_wrap && wrapc) { wk0 = ld_packets2(wrap_box + xo); wk1 = ld_packets2(wrap_box + xo + 16u); }
        };
        if (kTma && lane == 0) tma_fence_proxy_async();               // generic-proxy writes of the previous phase -> TMA reads
#pragma unroll 1
        for (int u = 0; u < D - 1; u++) issue(u);
        fetch_packets(0);
#pragma unroll 1
        for (int u = 0; u < nsteps; u++) {
            if (kTma) __syncwarp();                                    // every lane is done reading the slot refilled next
            issue(u + D - 1);
            if (kTma) {
                const int k = u & (D - 1);
                tma_mbar_wait(bars + (uint32_t)k * 8u, (phase >> k) & 1u);   // step u has landed
                phase ^= 1u << k;
            } else {
                cp_async_wait<D - 1>();                                // step u has landed
            }
            const int tc = tc_of(u);
            const bool act = rowok && (unsigned)tc < (unsigned)ntile;
            float nbv[kTK];
#pragma unroll
            for (int j = 0; j < kTK; j++)
                nbv[j] = kUp ? __shfl_down_sync(0xffffffffu, edge[j], 1) : __shfl_up_sync(0xffffffffu, edge[j], 1);
            const bool late = act && ((poller && (pk0.y != sid || pk0.w != sid || pk1.y != sid || pk1.w != sid)) ||
                                      (warp_wrap && wrapc && (wk0.y != sid || wk0.w != sid || wk1.y != sid || wk1.w != sid)));
            if (__any_sync(0xffffffffu, late)) {                       // a packet has not arrived yet: wait for it
                const uint32_t xo = (uint32_t)tc * (kTK * 8u);
                if (poller && act)
                    while (pk0.y != sid || pk0.w != sid || pk1.y != sid || pk1.w != sid) {
                        pk0 = ld_packets2(my_inbox + xo); pk1 = ld_packets2(my_inbox + xo + 16u);
                    }
                if (wrapc && act)
                    while (wk0.y != sid || wk0.w != sid || wk1.y != sid || wk1.w != sid) {
                        wk0 = ld_packets2(wrap_box + xo); wk1 = ld_packets2(wrap_box + xo + 16u);
                    }
            }
            if (poller) {
                nbv[0] = __uint_as_float(pk0.x); nbv[1] = __uint_as_float(pk0.z);
                nbv[2] = __uint_as_float(pk1.x); nbv[3] = __uint_as_float(pk1.z);
            }
            float fc[kTK] = {1.0f, 1.0f, 1.0f, 1.0f};                  // in-column wrap operands of this tile's columns
            if (warp_wrap && wrapc) {
                fc[0] = __uint_as_float(wk0.x); fc[1] = __uint_as_float(wk0.z);
                fc[2] = __uint_as_float(wk1.x); fc[3] = __uint_as_float(wk1.z);
            }
            if (act) {
                const uint32_t slot = ring + (uint32_t)((u & (D - 1)) * NI) * 512u;
                const int blk = blk_of(u);
                float res[kTM][kTK];
#pragma unroll
                for (int ii = 0; ii < kTM; ii++) {
                    const int i = kUp ? kTM - 1 - ii : ii;             // U solve: bottom row first
                    const float4 e4 = lds128(slot + (uint32_t)((MODE == 0 ? 32 : 16) + i) * 512u);
                    const float4 s4 = MODE == 2 ? lds128(slot + (uint32_t)(20 + i) * 512u) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float ev[4] = {e4.x, e4.y, e4.z, e4.w}, sv[4] = {s4.x, s4.y, s4.z, s4.w};
                    float4 lo[kTK];
#pragma unroll
                    for (int jj = 0; jj < kTK; jj++) {
                        const int j = kUp ? kTK - 1 - jj : jj;         // U solve: right column first
                        const float4 v = lds128(slot + (uint32_t)(i * kTK + j) * 512u);
                        const float4 rv = MODE == 0 ? lds128(slot + (uint32_t)(16 + i * kTK + j) * 512u) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float nb = ii == 0 ? nbv[j] : res[kUp ? i + 1 : i - 1][j];
                        const float prev = jj == 0 ? side[i] : res[i][kUp ? j + 1 : j - 1];
                        res[i][j] = tile_cell<MODE>(v, rv, ev[j], sv[j], nb, prev, fc[j], keep[i], lo[j]);
                    }
                    const float4 r4 = make_float4(res[i][0], res[i][1], res[i][2], res[i][3]);
                    if (MODE == 0) {
#pragma unroll
                        for (int j = 0; j < kTK; j++) pl.lval[cimg(blk, i * kTK + j, lane)] = lo[j];
                        pl.udiag[vimg(blk, i, lane)] = r4;
                    } else {
                        zs[vimg(blk, i, lane)] = r4;
                    }
                    side[i] = res[i][kUp ? 0 : kTK - 1];
                }
#pragma unroll
                for (int j = 0; j < kTK; j++) edge[j] = res[kUp ? 0 : kTM - 1][j];
                if (tc == xb_tc) {                                     // the column every row's in-row wrap refers to
#pragma unroll
                    for (int i = 0; i < kTM; i++)
                        keep[i] = xb_j == 0 ? res[i][0] : (xb_j == 1 ? res[i][1] : (xb_j == 2 ? res[i][2] : res[i][3]));
                }
                const uint32_t xo = (uint32_t)tc * (kTK * 8u);
                if (producer) {
#pragma unroll
                    for (int j = 0; j < kTK; j++) {
                        if (prod_remote) st_packet_remote(prod_addr + xo + 8u * j, edge[j], sid);
                        else st_packet_cta(prod_addr + xo + 8u * j, edge[j], sid);
                    }
                }
                if (warp_wrap && wrapp) {
#pragma unroll
                    for (int j = 0; j < kTK; j++) {
                        const float val = i_yb == 0 ? res[0][j] : (i_yb == 1 ? res[1][j] : (i_yb == 2 ? res[2][j] : res[3][j]));
                        if (wrapp_remote) st_packet_remote(wrapp_addr + xo + 8u * j, val, sid);
                        else st_packet_cta(wrapp_addr + xo + 8u * j, val, sid);
                    }
                }
            }
            fetch_packets(u + 1);
        }
        if (!kTma) cp_async_wait<0>();
    }
}

#define DPISO_TILE_TICK(slot)                                                        \
    do {                                                                             \
        if (prm.timing && blockIdx.x == 0 && threadIdx.x == 0) {                     \
            const long long _now = clock64();                                        \
            atomicAdd((unsigned long long *)&prm.timing[slot], (unsigned long long)(_now - tick)); \
            tick = _now;                                                             \
        }                                                                            \
    } while (0)

__device__ __forceinline__ float f4c(const float4 &a, int j) { return j == 0 ? a.x : (j == 1 ? a.y : (j == 2 ? a.z : a.w)); }

__global__ void __launch_bounds__(kBicgThreads, 1) bicgstab_tile_kernel(const TileParams tp) {
    long long tick = clock64();
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[64];
    __shared__ double s_part[2][kTileMaxCluster][2];
    const BicgParams &prm = tp.p;
    const int C = tp.C, Wc = tp.Wc;
    const int rank = (int)cluster.block_rank();
    const int sys = blockIdx.x / C;
    const int sample = sys >> 1, comp = sys & 1;
    const BicgTab &T = prm.tab[comp];
    TileGeo geo;
    geo.dx = T.dx; geo.dy = T.n / T.dx; geo.ntile = tp.dxp[comp] >> 2; geo.ntr = tp.dyp[comp] >> 2;
    geo.G = (geo.ntr + 31) >> 5; geo.S = geo.ntile + 31;
    const int dx = geo.dx, dy = geo.dy, dxp = tp.dxp[comp], dyp = tp.dyp[comp], ntile = geo.ntile, ntr = geo.ntr, S = geo.S;
    const int NB = geo.G * S;                                     // blocks of this component's images
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NT >> 5;
    const int wg = rank * NW + warp, nwg = C * NW;                // warp of the cluster: blocks wg, wg + nwg, ... are its own
    const int face_off = comp ? prm.tab[0].n : 0;
    const float *values_c = prm.values + (size_t)sample * prm.nnz_total + (comp ? prm.nnz[0] : 0);
    const int nnz_c = prm.nnz[comp];
    const float *rhs_g = prm.rhs + (size_t)sample * prm.n_face + face_off;
    const float *x0_g = prm.x0 + (size_t)sample * prm.n_face + face_off;
    float *x_g = prm.x + (size_t)sample * prm.n_face + face_off;

    const size_t cI = tp.blocks_max * 512, vI = tp.blocks_max * 128;       // float4 per coefficient / vector image
    float4 *cur = (float4 *)(prm.workspace + (size_t)sys * prm.ws_floats);
    TilePlanes pl;
    pl.alow = cur;  cur += cI;
    pl.uval = cur;  cur += cI;
    pl.lval = cur;  cur += cI;
    pl.adiag = cur; cur += vI;
    pl.udiag = cur; cur += vI;
    float4 *__restrict__ b = cur;
    float4 *__restrict__ x = b + vI;
    float4 *__restrict__ r = x + vI;
    float4 *__restrict__ rh = r + vI;                             // rh, p, v, tt: contiguous, double as the ILU-only arv image
    float4 *__restrict__ p = rh + vI;
    float4 *__restrict__ v = p + vI;
    float4 *__restrict__ tt = v + vI;
    pl.arv = rh;
    float4 *const zs = tt + vI;                                    // the solve vector

    TileCtx c;
    c.g = geo; c.Wc = Wc; c.rank = rank; c.dxp_max = tp.dxp_max;
    {   // inboxes: tag 0 = no sweep; then 4 mbarriers per sweep warp (one arrival each: the lane that issues the copies)
        unsigned long long *const boxes = (unsigned long long *)(smem_raw + (size_t)Wc * kTileRingBytes);
        for (int k = tid; k < (Wc + 2) * tp.dxp_max; k += NT) boxes[k] = 0ull;
        if (tid < Wc * 4) tma_mbar_init(tile_smem_u32(boxes + (size_t)(Wc + 2) * tp.dxp_max) + (uint32_t)tid * 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned tma_phase = 0;                                        // parity bits of this warp's mbarriers
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
    // element-wise phases: fn(idx) for every 16-byte item (4 cells of a tile row) of this warp's blocks
    auto for_items = [&](auto fn) {
        for (int blk = wg; blk < NB; blk += nwg) {
            const int g = blk / S, tc = blk - g * S - lane;
            if ((unsigned)tc < (unsigned)ntile && g * 32 + lane < ntr) {
#pragma unroll
                for (int i = 0; i < kTM; i++) fn(vimg(blk, i, lane));
            }
        }
    };
    // this CTA's share of the padded grid rows in the row-major passes (setup, pivots, result)
    const int rows_cta = (dyp + C - 1) / C;
    const int row_lo = min(dyp, rank * rows_cta), row_hi = min(dyp, row_lo + rows_cta);

    // ---- setup: canonical slots scattered into the images, NaN guard (":245-256") ---------------------------------
    double nv = 0.0, nb = 0.0;
#pragma unroll 8
    for (int i = rank * NT + tid; i < nnz_c; i += C * NT) { const double a = values_c[i]; nv += a * a; }
    const float sg = prm.sign;
    for (int t = row_lo + warp; t < row_hi; t += NW) {
        for (int xx = lane; xx < dxp; xx += 32) {
            const size_t qc = ccell(geo, t, xx), qv = vcell(geo, t, xx);
            if (t < dy && xx < dx) {
                const int i = t * dx + xx;
                const float bi = rhs_g[i];
                ((float *)b)[qv] = bi; nb += (double)bi * bi;
                ((float *)x)[qv] = x0_g[i];                           // cublasScopy(x_old -> x) (":261")
                auto val4 = [&](const int4 s4) {
                    return make_float4(s4.x >= 0 ? sg * values_c[s4.x] : 0.0f, s4.y >= 0 ? sg * values_c[s4.y] : 0.0f,
                                       s4.z >= 0 ? sg * values_c[s4.z] : 0.0f, s4.w >= 0 ? sg * values_c[s4.w] : 0.0f);
                };
                pl.alow[qc] = val4(T.c_lsrc[i]);
                pl.arv[qc] = val4(T.c_lrev[i]);
                pl.uval[qc] = val4(T.c_usrc[i]);
                const int ds = T.c_dsrc[i];
                ((float *)pl.adiag)[qv] = ds >= 0 ? sg * values_c[ds] : 1.0f;
            } else {                                                  // padding: unit diagonal, nothing else
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                ((float *)b)[qv] = 0.0f; ((float *)x)[qv] = 0.0f;
                pl.alow[qc] = z4; pl.arv[qc] = z4; pl.uval[qc] = z4;
                ((float *)pl.adiag)[qv] = 1.0f;
            }
        }
    }
    cluster_sum2(nv, nb);
    const int warn = (isnan((float)sqrt(nv)) || isnan((float)sqrt(nb))) ? 1 : 0;
    DPISO_TILE_TICK(0);

    unsigned sid = 0;
    const TileFar far_l = tp.far_l[comp], far_u = tp.far_u[comp];
    // ---- ILU(0) (csrilu02, ":181-218") ---------------------------------------------------------------------
    if (prm.pivots_in && ((prm.reuse_mask >> comp) & 1)) {
        // factor reuse: the pivots d of the solve with the other orientation of this matrix are given, and
        // ILU(0)(M^T) = (U^T D^-1)(D L^T): l'_ik = m_ik / d_k, upper entries unchanged, pivots d -- a parallel pass
        // instead of the wavefront sweep (same division as the sweep)
        const float *d_in = prm.pivots_in + (size_t)sample * prm.n_face + face_off;      // caller's row order
        for (int t = row_lo + warp; t < row_hi; t += NW) {
            for (int xx = lane; xx < dxp; xx += 32) {
                const size_t qc = ccell(geo, t, xx), qv = vcell(geo, t, xx);
                if (t < dy && xx < dx) {
                    const int i = t * dx + xx;
                    const float4 a = pl.alow[qc];
                    const float p0 = t == far_l.ya ? d_in[far_l.yb * dx + xx] : 1.0f, p2 = xx == far_l.xa ? d_in[t * dx + far_l.xb] : 1.0f;
                    const float p1 = t > 0 ? d_in[i - dx] : 1.0f, p3 = xx > 0 ? d_in[i - 1] : 1.0f;
                    pl.lval[qc] = make_float4(ilu_div(a.x, p0), ilu_div(a.y, p1), ilu_div(a.z, p2), ilu_div(a.w, p3));
                    ((float *)pl.udiag)[qv] = d_in[i];
                } else {
                    pl.lval[qc] = make_float4(0.f, 0.f, 0.f, 0.f);
                    ((float *)pl.udiag)[qv] = 1.0f;
                }
            }
        }
        cluster.sync();
    } else {
        cluster.sync();                                              // the images are complete
        if (tp.use_tma) tile_sweep<0, true>(c, pl, far_l, nullptr, zs, ++sid, tma_phase);
        else tile_sweep<0, false>(c, pl, far_l, nullptr, zs, ++sid, tma_phase);
        cluster.sync();
        if (prm.pivots_out) {
            float *d_out = prm.pivots_out + (size_t)sample * prm.n_face + face_off;
            for (int t = row_lo + warp; t < min(row_hi, dy); t += NW)
                for (int xx = lane; xx < dx; xx += 32) d_out[t * dx + xx] = ((const float *)pl.udiag)[vcell(geo, t, xx)];
        }
    }
    DPISO_TILE_TICK(1);

    auto precondition = [&](const float4 *src) {                     // zs = U^-1 L^-1 src   (csrsv2 x2, ":321-327")
        cluster.sync();                                              // src complete in global memory
        if (tp.use_tma) tile_sweep<1, true>(c, pl, far_l, src, zs, ++sid, tma_phase);
        else tile_sweep<1, false>(c, pl, far_l, src, zs, ++sid, tma_phase);
        cluster.sync();
        if (tp.use_tma) tile_sweep<2, true>(c, pl, far_u, nullptr, zs, ++sid, tma_phase);
        else tile_sweep<2, false>(c, pl, far_u, nullptr, zs, ++sid, tma_phase);
        cluster.sync();
    };
    // y = A vec tile by tile (CsrmvEx per row: lower slots, diagonal, upper slots = ascending column order).  The
    // x-neighbours of a tile are the same lane's tiles in blocks sigma -+ 1, the y-neighbours lane -+ 1 of those blocks
    // (lane 31 / 0 of the neighbouring warp group at the edges).  fn(idx, y4) consumes one tile row.
    auto spmv_tiles = [&](const float4 *vec, auto fn) {
        for (int blk = wg; blk < NB; blk += nwg) {
            const int g = blk / S, sgm = blk - g * S, tc = sgm - lane, tr = g * 32 + lane;
            if (!((unsigned)tc < (unsigned)ntile && tr < ntr)) continue;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o[kTM];
            float lw[kTM], rx[kTM], wrl[kTM], wru[kTM];
#pragma unroll
            for (int i = 0; i < kTM; i++) {
                o[i] = vec[vimg(blk, i, lane)];
                lw[i] = tc > 0 ? vec[vimg(blk - 1, i, lane)].w : 0.0f;
                rx[i] = tc < ntile - 1 ? vec[vimg(blk + 1, i, lane)].x : 0.0f;
                wrl[i] = 0.0f; wru[i] = 0.0f;
            }
            const float4 tp4 = tr > 0 ? (lane > 0 ? vec[vimg(blk - 1, 3, lane - 1)] : vec[vimg((g - 1) * S + tc + 31, 3, 31)]) : z4;
            const float4 bt4 = tr < ntr - 1 ? (lane < 31 ? vec[vimg(blk + 1, 0, lane + 1)] : vec[vimg((g + 1) * S + tc, 0, 0)]) : z4;
            if (far_l.xa >= 0 && tc == (far_l.xa >> 2)) {
#pragma unroll
                for (int i = 0; i < kTM; i++) wrl[i] = f4c(vec[vimg(g * S + (far_l.xb >> 2) + lane, i, lane)], far_l.xb & 3);
            }
            if (far_u.xa >= 0 && tc == (far_u.xa >> 2)) {
#pragma unroll
                for (int i = 0; i < kTM; i++) wru[i] = f4c(vec[vimg(g * S + (far_u.xb >> 2) + lane, i, lane)], far_u.xb & 3);
            }
            float4 cwl = z4, cwu = z4;                               // in-column wrap operands (one tile row of the system each)
            if (far_l.ya >= 0 && tr == (far_l.ya >> 2)) {
                const int ytr = far_l.yb >> 2;
                cwl = vec[vimg((ytr >> 5) * S + tc + (ytr & 31), far_l.yb & 3, ytr & 31)];
            }
            if (far_u.ya >= 0 && tr == (far_u.ya >> 2)) {
                const int ytr = far_u.yb >> 2;
                cwu = vec[vimg((ytr >> 5) * S + tc + (ytr & 31), far_u.yb & 3, ytr & 31)];
            }
#pragma unroll
            for (int i = 0; i < kTM; i++) {
                const int t = tr * kTM + i;
                const float4 dg4 = pl.adiag[vimg(blk, i, lane)];
                const float4 up_row = i > 0 ? o[i - 1] : tp4, dn_row = i < kTM - 1 ? o[i + 1] : bt4;
                float y[kTK];
#pragma unroll
                for (int j = 0; j < kTK; j++) {
                    const int xx = tc * kTK + j;
                    const float4 lo = pl.alow[cimg(blk, i * kTK + j, lane)], up = pl.uval[cimg(blk, i * kTK + j, lane)];
                    const float l0 = t == far_l.ya ? f4c(cwl, j) : 0.0f, l1 = f4c(up_row, j);
                    const float l2 = xx == far_l.xa ? wrl[i] : 0.0f, l3 = j > 0 ? f4c(o[i], j - 1) : lw[i];
                    const float u0 = j < kTK - 1 ? f4c(o[i], j + 1) : rx[i], u1 = xx == far_u.xa ? wru[i] : 0.0f;
                    const float u2 = f4c(dn_row, j), u3 = t == far_u.ya ? f4c(cwu, j) : 0.0f;
                    float acc = fmaf(lo.x, l0, 0.0f);
                    acc = fmaf(lo.y, l1, acc);
                    acc = fmaf(lo.z, l2, acc);
                    acc = fmaf(lo.w, l3, acc);
                    acc = fmaf(f4c(dg4, j), f4c(o[i], j), acc);
                    acc = fmaf(up.x, u0, acc);
                    acc = fmaf(up.y, u1, acc);
                    acc = fmaf(up.z, u2, acc);
                    acc = fmaf(up.w, u3, acc);
                    y[j] = acc;
                }
                fn(vimg(blk, i, lane), make_float4(y[0], y[1], y[2], y[3]));
            }
        }
    };
    auto dot4 = [](const float4 a, const float4 c4) {
        double s = (double)a.x * c4.x; s += (double)a.y * c4.y; s += (double)a.z * c4.z; s += (double)a.w * c4.w;
        return s;
    };

    float alpha = 1.f, rho = 1.f, rhop = 1.f, omega = 1.f, beta, nrm_r = 0.f;
    int it_count = 0, restarts = 0, exit_kind = 3;
    const float tol = prm.tol;

    for (int restart = 0; restart < 2; restart++) {
        restarts = restart;
        cluster.sync();                                              // x complete (setup / restart reset)
        double s0 = 0.0, s1 = 0.0;
        spmv_tiles(x, [&](size_t q, const float4 ax) {               // r = b - A x  (":275-282")
            const float4 bb = b[q];
            const float4 rq = make_float4(__fsub_rn(bb.x, ax.x), __fsub_rn(bb.y, ax.y), __fsub_rn(bb.z, ax.z), __fsub_rn(bb.w, ax.w));
            r[q] = rq; s0 += dot4(rq, rq);
        });
        cluster_sum2(s0, s1);
        nrm_r = (float)sqrt(s0);
        if (nrm_r < tol) { exit_kind = 0; break; }                   // lucky guess (":287-289")
        for_items([&](size_t q) { rh[q] = r[q]; p[q] = make_float4(0.f, 0.f, 0.f, 0.f); v[q] = make_float4(0.f, 0.f, 0.f, 0.f); });
        exit_kind = 3;
        float rho_next = (float)s0;                                  // r.rh with rh = r
        for (int it = 0; it < prm.max_it; it++) {
            it_count++;
            rhop = rho;
            rho = rho_next;
            beta = __fmul_rn(__fdiv_rn(rho, rhop), __fdiv_rn(alpha, omega));
            for_items([&](size_t q) {                                // p = r + beta (p - omega v)  (":315-317")
                const float4 vv = v[q], rr = r[q];
                float4 pp = p[q];
                pp.x = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.x, pp.x)), rr.x);
                pp.y = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.y, pp.y)), rr.y);
                pp.z = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.z, pp.z)), rr.z);
                pp.w = __fadd_rn(__fmul_rn(beta, fmaf(-omega, vv.w, pp.w)), rr.w);
                p[q] = pp;
            });
            DPISO_TILE_TICK(4);
            precondition(p);                                         // zs = p_hat
            DPISO_TILE_TICK(2);
            s0 = 0.0; s1 = 0.0;
            spmv_tiles(zs, [&](size_t q, const float4 vq) { v[q] = vq; s0 += dot4(rh[q], vq); });
            cluster_sum2(s0, s1);
            alpha = __fdiv_rn(rho, (float)s0);
            s0 = 0.0; s1 = 0.0;
            for_items([&](size_t q) {                                // x += alpha p_hat ; r -= alpha v ; |r|
                const float4 zz = zs[q], vv = v[q];
                float4 xx = x[q], rr = r[q];
                xx.x = fmaf(alpha, zz.x, xx.x); xx.y = fmaf(alpha, zz.y, xx.y); xx.z = fmaf(alpha, zz.z, xx.z); xx.w = fmaf(alpha, zz.w, xx.w);
                rr.x = fmaf(-alpha, vv.x, rr.x); rr.y = fmaf(-alpha, vv.y, rr.y); rr.z = fmaf(-alpha, vv.z, rr.z); rr.w = fmaf(-alpha, vv.w, rr.w);
                x[q] = xx; r[q] = rr;
                s0 += dot4(rr, rr);
            });
            cluster_sum2(s0, s1);
            nrm_r = (float)sqrt(s0);
            if (nrm_r < tol) { exit_kind = 1; break; }
            DPISO_TILE_TICK(4);
            precondition(r);                                         // zs = s_hat
            DPISO_TILE_TICK(2);
            s0 = 0.0; s1 = 0.0;
            spmv_tiles(zs, [&](size_t q, const float4 tq) { tt[q] = tq; s0 += dot4(tq, r[q]); s1 += dot4(tq, tq); });
            cluster_sum2(s0, s1);
            omega = __fdiv_rn((float)s0, (float)s1);
            s0 = 0.0; s1 = 0.0;
            for_items([&](size_t q) {                                // x += omega s_hat ; r -= omega t ; |r| ; r.rh
                const float4 zz = zs[q], t4 = tt[q], hh = rh[q];
                float4 xx = x[q], rr = r[q];
                xx.x = fmaf(omega, zz.x, xx.x); xx.y = fmaf(omega, zz.y, xx.y); xx.z = fmaf(omega, zz.z, xx.z); xx.w = fmaf(omega, zz.w, xx.w);
                rr.x = fmaf(-omega, t4.x, rr.x); rr.y = fmaf(-omega, t4.y, rr.y); rr.z = fmaf(-omega, t4.z, rr.z); rr.w = fmaf(-omega, t4.w, rr.w);
                x[q] = xx; r[q] = rr;
                s0 += dot4(rr, rr); s1 += dot4(rr, hh);
            });
            cluster_sum2(s0, s1);
            nrm_r = (float)sqrt(s0);
            rho_next = (float)s1;
            if (nrm_r < tol) { exit_kind = 2; break; }
        }
        if (nrm_r > __fmul_rn(tol, 100.0f) || isnan(nrm_r)) {        // ":392-404"
            for_items([&](size_t q) { x[q] = make_float4(0.f, 0.f, 0.f, 0.f); });
            if (restart == 1) restarts = 2;
        } else break;
    }
    cluster.sync();                                                  // x complete; no CTA leaves while packets may be in flight
    DPISO_TILE_TICK(4);
    for (int t = row_lo + warp; t < min(row_hi, dy); t += NW)        // back to the caller's row order
        for (int xx = lane; xx < dx; xx += 32) x_g[t * dx + xx] = ((const float *)x)[vcell(geo, t, xx)];
    if (rank == 0 && tid == 0) {
        int *st = prm.stats + (size_t)sys * 4;
        st[0] = it_count; st[1] = restarts; st[2] = warn; st[3] = exit_kind;
        if (warn) *prm.warn = 1.0f;
    }
}

static int g_tile_cluster = 0;       // tuning override (0 = heuristic)

static void tile_padded(const dpiso_bicg_tables *h, int *dxp, int *dyp) {
    *dxp = (h->dx + kTK - 1) / kTK * kTK;
    *dyp = (h->n / h->dx + kTM - 1) / kTM * kTM;
}
static size_t tile_blocks(const dpiso_bicg_tables *h) {
    int dxp, dyp;
    tile_padded(h, &dxp, &dyp);
    return (size_t)((dyp / kTM + 31) / 32) * (size_t)(dxp / kTK + 31);
}

size_t tile_workspace_floats(const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v) {
    const size_t bu = tile_blocks(h_tab_u), bv = tile_blocks(h_tab_v);
    return (size_t)kTileBlockFloats * (bu > bv ? bu : bv);
}

int launch_bicgstab_tile(BicgParams &prm, const dpiso_bicg_tables *h_tab_u, const dpiso_bicg_tables *h_tab_v, int batch,
                         void *stream) {
    if (!(h_tab_u->band_ok && h_tab_v->band_ok && h_tab_u->c_lsrc && h_tab_v->c_lsrc)) return DPISO_EUNSUPPORTED;
    for (int k = 0; k < 2; k++) {
        // a wrap operand must come from another tile than the one that consumes it (an earlier step of the same thread /
        // an earlier thread): true for every grid with at least 8 faces per direction
        const dpiso_bicg_tables *h = k ? h_tab_v : h_tab_u;
        if (h->dx < 2 * kTK || h->n / h->dx < 2 * kTM) return DPISO_EUNSUPPORTED;
        for (int d = 0; d < 8; d += 4) {
            if (h->far[d] >= 0 && h->far[d] / kTK == h->far[d + 1] / kTK) return DPISO_EUNSUPPORTED;
            if (h->far[d + 2] >= 0 && h->far[d + 2] / kTM == h->far[d + 3] / kTM) return DPISO_EUNSUPPORTED;
        }
    }
    TileParams tp;
    tile_padded(h_tab_u, &tp.dxp[0], &tp.dyp[0]);
    tile_padded(h_tab_v, &tp.dxp[1], &tp.dyp[1]);
    tp.dxp_max = tp.dxp[0] > tp.dxp[1] ? tp.dxp[0] : tp.dxp[1];
    tp.blocks_max = tile_workspace_floats(h_tab_u, h_tab_v) / kTileBlockFloats;
    if ((size_t)kTileBlockFloats * tp.blocks_max > prm.ws_floats) return DPISO_EUNSUPPORTED;
    const int dymax = tp.dyp[0] > tp.dyp[1] ? tp.dyp[0] : tp.dyp[1];
    const int warps = (dymax / kTM + 31) / 32;                    // sweep warps per system
    const size_t kBudget = 224 * 1024;
    auto smem_of = [&](int Wc) { return (size_t)Wc * kTileRingBytes + (size_t)(Wc + 2) * tp.dxp_max * 8 + (size_t)Wc * 32; };
    // fewest CTAs that hold the sweep warps, then more CTAs per system while the whole batch still fits one wave (the
    // SpMV / vector phases scale with the CTAs; the sweeps do not care)
    int C = 0, Wc = 0;
    for (int wc = kTileMaxWarps; wc >= 1; wc--) {
        if (smem_of(wc) > kBudget) continue;
        const int cand = (warps + wc - 1) / wc;
        if (cand <= kTileMaxCluster) { C = cand; Wc = wc; break; }
    }
    if (!C) return DPISO_EUNSUPPORTED;
    {
        const int fit = 148 / (batch * 2) < 8 ? 148 / (batch * 2) : 8;
        if (fit > C) C = fit;
    }
    if (g_tile_cluster >= C && g_tile_cluster <= kTileMaxCluster) C = g_tile_cluster;
    Wc = (warps + C - 1) / C;
    tp.C = C; tp.Wc = Wc;
    // ring refill: per-lane cp.async by default; debug bit 1024 selects the TMA bulk-copy pipeline (one elected lane fetches
    // the whole block, completion on an mbarrier).  Measured on B200 the TMA pipeline is 3-8 % slower per sweep (2048^2 x 4:
    // 24.2 M vs 22.4 M cycles; 1024^2 x 8: 12.1 M vs 11.2 M): with only 3 steps in flight the mbarrier wake-up sits on the
    // critical path of the lone sweep warp, and the per-lane copies skip the lanes that are outside the grid.
    tp.use_tma = (prm.dbg & 1024) ? 1 : 0;
    tp.p = prm;
    for (int k = 0; k < 2; k++) {
        const dpiso_bicg_tables *h = k ? h_tab_v : h_tab_u;
        tp.far_l[k] = {h->far[0], h->far[1], h->far[2], h->far[3]};
        tp.far_u[k] = {h->far[4], h->far[5], h->far[6], h->far[7]};
    }
    const size_t smem = smem_of(Wc);
    auto kernel = bicgstab_tile_kernel;
    DPISO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (C > 8) DPISO_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * 2 * C));
    cfg.blockDim = dim3(kBicgThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, tp);
    if (e != cudaSuccess) {
        set_error("BiCGStab (tile variant, %d CTAs per system) launch failed: %s", C, cudaGetErrorString(e));
        return DPISO_ECUDA;
    }
    return DPISO_OK;
}

}  // namespace dpiso

extern "C" int dpiso_bicgstab_set_tile_cluster(int cluster) {
    dpiso::g_tile_cluster = cluster;
    return DPISO_OK;
}
