// tables.cu -- host-side construction of the structure tables of the batched ILU(0)-BiCGStab kernels (bicgstab.cu).
//
// The reference launcher takes nothing but the CSR arrays and `transpose_op`
// (CUDAsrc/multi_bicgstab_ilu_linear_solve_op.cc:50-58) and lets cuSPARSE analyse the pattern on every call
// (csr2csc ":113-134", csrilu02 / csrsv2 analysis ":181-228").  Here the pattern depends only on
// (ny, nx, periodic flags, component, transpose); the analysis is done ONCE on the host, in closed form from
// structure.cuh, and uploaded:  dpiso_bicg_tables_create / _destroy.  A C or C++ caller therefore needs no Python.
//
// The builder also PROVES, for the concrete grid, the two properties the kernels rely on: (i) lx + ly is a valid
// level schedule (every lower entry points to a strictly lower level, every upper entry to a strictly higher one,
// SURVEY N6), (ii) ILU(0) on this pattern only changes the pivots and the lower entries (no fill interaction, SURVEY
// N7); otherwise it returns DPISO_EUNSUPPORTED.  diffpiso_b200/structure.py holds an independent numpy/scipy
// derivation of the same tables; tests/test_cpu_structure.py compares the two entry by entry.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "structure.cuh"

namespace dpiso {

namespace {

constexpr int kMaxWaHost = 6;

struct Entry { int col, src; };

struct HostTables {
    int n = 0, n_levels = 0, wa = 0, max_level = 0, wl = 0, wu = 0, dx = 0, rows_ok = 0, sym = 0, band_ok = 0;
    int far[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    std::vector<int> level_ptr, perm, a_col, a_src, a_rev, r_col, r_src, r_rev;
    std::vector<int> c_lsrc, c_lrev, c_usrc, c_lfar, c_ufar, c_dsrc, m_nbr, m_lfar, m_ufar;
};

// CSR value index of M(row, col), -1 if (row, col) is not an entry
int find_src(const std::vector<std::vector<Entry>> &rows, int row, int col) {
    const std::vector<Entry> &r = rows[row];
    for (const Entry &e : r) if (e.col == col) return e.src;
    return -1;
}

int build_tables(int ny, int nx, int per_x, int per_y, int comp, int transpose, HostTables &t) {
    if (ny < 3 || nx < 3) { set_error("grid too small for the 5-point pattern (need ny, nx >= 3)"); return DPISO_EINVAL; }
    const CompDims cd = comp_dims(ny, nx, comp);
    const int Dx = cd.Dx, Dy = cd.Dy, n = Dx * Dy;
    // pattern of A in the reference's CSR layout: existing neighbours + self in ascending column order
    std::vector<std::vector<Entry>> a_rows(n), m_rows(n);
    for (int ly = 0; ly < Dy; ly++)
        for (int lx = 0; lx < Dx; lx++) {
            const RowLayout L = row_layout(lx, ly, cd, per_x, per_y);
            const int row = lx + Dx * ly;
            std::vector<Entry> &r = a_rows[row];
            for (int k = 0; k < 5; k++)
                if (k == 4 || L.has[k]) r.push_back({L.col[k], L.rp + L.slot[k]});
            std::sort(r.begin(), r.end(), [](const Entry &x, const Entry &y) { return x.col < y.col; });
            for (size_t k = 1; k < r.size(); k++)
                if (r[k].col == r[k - 1].col) { set_error("degenerate grid: two neighbours of a face coincide"); return DPISO_EUNSUPPORTED; }
            for (size_t k = 0; k < r.size(); k++)
                if (r[k].src != L.rp + (int)k) { set_error("internal: CSR slot order mismatch"); return DPISO_EINVAL; }
        }
    if (transpose) {
        for (int i = 0; i < n; i++)
            for (const Entry &e : a_rows[i]) m_rows[e.col].push_back({i, e.src});   // ascending i => ascending column
    } else {
        m_rows = a_rows;
    }
    int wa = 0;
    for (int i = 0; i < n; i++) wa = std::max(wa, (int)m_rows[i].size());
    if (wa > kMaxWaHost) { set_error("row with %d entries", wa); return DPISO_EUNSUPPORTED; }
    // (i) wavefront property of the lx + ly levels
    auto level_of = [&](int row) { return row % Dx + row / Dx; };
    int wl = 0, wu = 0;
    for (int i = 0; i < n; i++) {
        int nl = 0, nu = 0;
        for (const Entry &e : m_rows[i]) {
            if (e.col < i) { nl++; if (!(level_of(e.col) < level_of(i))) { set_error("lx+ly is not a valid level schedule for this grid"); return DPISO_EUNSUPPORTED; } }
            if (e.col > i) { nu++; if (!(level_of(e.col) > level_of(i))) { set_error("lx+ly is not a valid level schedule for this grid"); return DPISO_EUNSUPPORTED; } }
        }
        wl = std::max(wl, nl); wu = std::max(wu, nu);
    }
    // (ii) ILU(0) touches only pivots and lower entries: for k in L(i) and j in row i with j > k, j != i, (k, j) must
    //      not be an entry of M
    for (int i = 0; i < n; i++) {
        const std::vector<Entry> &r = m_rows[i];
        for (size_t s1 = 0; s1 < r.size(); s1++)
            for (size_t s2 = s1 + 1; s2 < r.size(); s2++) {
                const int k = r[s1].col, j = r[s2].col;
                if (k < i && j != i && find_src(m_rows, k, j) >= 0) {
                    set_error("ILU(0) fill interaction on this grid (too small / degenerate)");
                    return DPISO_EUNSUPPORTED;
                }
            }
    }
    t.n = n; t.wa = wa; t.wl = wl; t.wu = wu; t.dx = Dx;
    // ELL in original numbering: slot k of row i = k-th entry in ascending column order
    t.r_col.assign((size_t)wa * n, 0); t.r_src.assign((size_t)wa * n, -1); t.r_rev.assign((size_t)wa * n, -1);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < wa; k++) {
            const bool on = k < (int)m_rows[i].size();
            t.r_col[(size_t)k * n + i] = on ? m_rows[i][k].col : i;
            t.r_src[(size_t)k * n + i] = on ? m_rows[i][k].src : -1;
            t.r_rev[(size_t)k * n + i] = on ? find_src(m_rows, m_rows[i][k].col, i) : -1;
        }
    t.sym = 1;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < (int)m_rows[i].size(); k++)
            if (t.r_rev[(size_t)k * n + i] < 0) t.sym = 0;
    // level-major permutation (stable by level)
    const int n_levels = Dx + Dy - 1;
    t.n_levels = n_levels;
    std::vector<int> counts(n_levels, 0);
    for (int i = 0; i < n; i++) counts[level_of(i)]++;
    t.level_ptr.assign(n_levels + 1, 0);
    for (int d = 0; d < n_levels; d++) t.level_ptr[d + 1] = t.level_ptr[d] + counts[d];
    t.max_level = *std::max_element(counts.begin(), counts.end());
    t.perm.assign(n, 0);
    std::vector<int> pos(n), fill(t.level_ptr.begin(), t.level_ptr.end() - 1);
    for (int i = 0; i < n; i++) { const int q = fill[level_of(i)]++; t.perm[q] = i; pos[i] = q; }
    t.a_col.assign((size_t)wa * n, 0); t.a_src.assign((size_t)wa * n, -1); t.a_rev.assign((size_t)wa * n, -1);
    for (int q = 0; q < n; q++) {
        const int i = t.perm[q];
        for (int k = 0; k < wa; k++) {
            const bool on = k < (int)m_rows[i].size();
            t.a_col[(size_t)k * n + q] = on ? pos[m_rows[i][k].col] : q;
            t.a_src[(size_t)k * n + q] = t.r_src[(size_t)k * n + i];
            t.a_rev[(size_t)k * n + q] = t.r_rev[(size_t)k * n + i];
            if (on) {   // after the permutation lower entries precede the row, upper entries follow it
                const int c = m_rows[i][k].col;
                if ((c < i && !(t.a_col[(size_t)k * n + q] < q)) || (c > i && !(t.a_col[(size_t)k * n + q] > q))) {
                    set_error("internal: level-major order does not preserve the triangular split");
                    return DPISO_EINVAL;
                }
            }
        }
    }
    // Row-major kernel: every row keeps its entries in canonical slots by kind -- lower: [far below the y-neighbour,
    // y-neighbour (x, y-1), far above it, x-neighbour (x-1, y)], upper: [x-neighbour, far below the y-neighbour,
    // y-neighbour, far above it] -- which preserves the ascending-column order only if each far slot is used at most
    // once per row; far operands are read one level ahead of their use, so they must be at least two levels old.
    t.c_lsrc.assign((size_t)n * 4, -1); t.c_lrev.assign((size_t)n * 4, -1); t.c_usrc.assign((size_t)n * 4, -1);
    t.c_lfar.assign((size_t)n * 2, -1); t.c_ufar.assign((size_t)n * 2, -1); t.c_dsrc.assign(n, -1);
    int rows_ok = 1;
    for (int i = 0; i < n && rows_ok; i++) {
        const int lx = i % Dx;
        int used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (const Entry &e : m_rows[i]) {
            const int c = e.col;
            if (c == i) continue;
            const bool regular = (c == i - 1 && lx > 0) || (c == i + 1 && lx < Dx - 1) || c == i - Dx || c == i + Dx;
            int kind;
            if (c < i) kind = !regular ? (c < i - Dx ? 0 : 2) : (c == i - Dx ? 1 : 3);
            else kind = !regular ? (c > i + Dx ? 7 : 5) : (c == i + Dx ? 6 : 4);
            if (used[kind]++) rows_ok = 0;
            if (!regular && std::abs(level_of(c) - level_of(i)) < 2) rows_ok = 0;
            // decoupled-warp sweeps: a far operand lies in the same grid row (same sweep thread) or in the same grid column
            if (!regular && c / Dx != i / Dx && c % Dx != i % Dx) rows_ok = 0;
            // ... and, when it comes from another row, from at least three rows away (it is fetched one step ahead of its use)
            if (!regular && c / Dx != i / Dx && std::abs(c / Dx - i / Dx) < 3) rows_ok = 0;
        }
    }
    t.rows_ok = rows_ok;
    if (rows_ok) {
        for (int i = 0; i < n; i++) {
            const int lx = i % Dx;
            for (const Entry &e : m_rows[i]) {
                const int c = e.col;
                if (c == i) { t.c_dsrc[i] = e.src; continue; }
                const bool regular = (c == i - 1 && lx > 0) || (c == i + 1 && lx < Dx - 1) || c == i - Dx || c == i + Dx;
                if (c < i) {
                    const int kind = !regular ? (c < i - Dx ? 0 : 2) : (c == i - Dx ? 1 : 3);
                    t.c_lsrc[(size_t)i * 4 + kind] = e.src;
                    t.c_lrev[(size_t)i * 4 + kind] = find_src(m_rows, c, i);
                    if (!regular) t.c_lfar[(size_t)i * 2 + kind / 2] = c;
                } else {
                    const int kind = (!regular ? (c > i + Dx ? 7 : 5) : (c == i + Dx ? 6 : 4)) - 4;
                    t.c_usrc[(size_t)i * 4 + kind] = e.src;
                    if (!regular) t.c_ufar[(size_t)i * 2 + kind / 2] = c;
                }
            }
        }
    }
    // Cluster-per-system kernel (bicgstab_band.cu): the far operands in closed form.  Lower entries: canonical slot 2 is
    // an in-row wrap (every row takes, at column xa, its own value of column xb), slot 0 an in-column wrap (grid row ya
    // takes row yb's value of the same column); upper entries: slot 1 in-row, slot 3 in-column.  band_ok = the four
    // numbers per direction describe EVERY far entry of the pattern, and nothing else.
    t.band_ok = rows_ok;
    for (int k = 0; k < 8; k++) t.far[k] = -1;
    if (rows_ok) {
        auto note = [&](int slot_a, int slot_b, int a, int b) {
            if (t.far[slot_a] < 0) { t.far[slot_a] = a; t.far[slot_b] = b; }
            else if (t.far[slot_a] != a || t.far[slot_b] != b) t.band_ok = 0;
        };
        for (int i = 0; i < n; i++) {
            const int lx = i % Dx, ly = i / Dx;
            for (int dir = 0; dir < 2; dir++) {
                const std::vector<int> &cf = dir ? t.c_ufar : t.c_lfar;
                // lower: [0] = slot 0 (in-column), [1] = slot 2 (in-row); upper: [0] = slot 1 (in-row), [1] = slot 3 (in-column)
                const int c_row = cf[(size_t)i * 2 + (dir ? 0 : 1)], c_col = cf[(size_t)i * 2 + (dir ? 1 : 0)];
                if (c_row >= 0) { if (c_row / Dx != ly) t.band_ok = 0; note(4 * dir + 0, 4 * dir + 1, lx, c_row % Dx); }
                if (c_col >= 0) { if (c_col % Dx != lx) t.band_ok = 0; note(4 * dir + 2, 4 * dir + 3, ly, c_col / Dx); }
            }
        }
        for (int i = 0; i < n && t.band_ok; i++) {                 // ... and every row the description names has the entry
            const int lx = i % Dx, ly = i / Dx;
            for (int dir = 0; dir < 2; dir++) {
                const std::vector<int> &cf = dir ? t.c_ufar : t.c_lfar;
                const bool has_row = cf[(size_t)i * 2 + (dir ? 0 : 1)] >= 0, has_col = cf[(size_t)i * 2 + (dir ? 1 : 0)] >= 0;
                if (has_row != (lx == t.far[4 * dir + 0]) || has_col != (ly == t.far[4 * dir + 2])) t.band_ok = 0;
            }
        }
        // sweep order: the lower sweeps walk x and the rows upwards, the upper sweep downwards
        if (t.far[0] >= 0 && !(t.far[1] < t.far[0])) t.band_ok = 0;
        if (t.far[2] >= 0 && !(t.far[3] < t.far[2])) t.band_ok = 0;
        if (t.far[4] >= 0 && !(t.far[5] > t.far[4])) t.band_ok = 0;
        if (t.far[6] >= 0 && !(t.far[7] > t.far[6])) t.band_ok = 0;
    }
    // level-major positions for the row-major kernel's planes / vectors
    t.m_nbr.assign((size_t)n * 4, 0); t.m_lfar.assign((size_t)n * 2, -1); t.m_ufar.assign((size_t)n * 2, -1);
    for (int q = 0; q < n; q++) {
        const int i = t.perm[q], lx = i % Dx, ly = i / Dx;
        t.m_nbr[(size_t)q * 4 + 0] = lx > 0 ? pos[i - 1] : q;
        t.m_nbr[(size_t)q * 4 + 1] = ly > 0 ? pos[i - Dx] : q;
        t.m_nbr[(size_t)q * 4 + 2] = lx < Dx - 1 ? pos[i + 1] : q;
        t.m_nbr[(size_t)q * 4 + 3] = ly < Dy - 1 ? pos[i + Dx] : q;
        if (rows_ok)
            for (int k = 0; k < 2; k++) {
                const int cl = t.c_lfar[(size_t)i * 2 + k], cu = t.c_ufar[(size_t)i * 2 + k];
                t.m_lfar[(size_t)q * 2 + k] = cl >= 0 ? pos[cl] : -1;
                t.m_ufar[(size_t)q * 2 + k] = cu >= 0 ? pos[cu] : -1;
            }
    }
    return DPISO_OK;
}

// all integer tables of one component, packed back to back (one allocation); offsets in ints
struct Packed {
    std::vector<int> data;
    size_t off[17];
};

Packed pack(const HostTables &t) {
    Packed p;
    const std::vector<int> *v[17] = {&t.level_ptr, &t.perm, &t.a_col, &t.a_src, &t.a_rev, &t.r_col, &t.r_src, &t.r_rev,
                                     &t.c_lsrc, &t.c_lrev, &t.c_usrc, &t.c_lfar, &t.c_ufar, &t.c_dsrc, &t.m_nbr, &t.m_lfar,
                                     &t.m_ufar};
    size_t total = 0;
    for (int k = 0; k < 17; k++) { p.off[k] = total; total += (v[k]->size() + 3) & ~(size_t)3; }   // 16-byte aligned pieces
    p.data.assign(total, 0);
    for (int k = 0; k < 17; k++) std::copy(v[k]->begin(), v[k]->end(), p.data.begin() + p.off[k]);
    return p;
}

void fill_struct(const HostTables &t, const Packed &p, const int *base, dpiso_bicg_tables *out) {
    out->n = t.n; out->n_levels = t.n_levels; out->wa = t.wa; out->max_level = t.max_level; out->wl = t.wl; out->wu = t.wu;
    out->dx = t.dx; out->rows_ok = t.rows_ok; out->sym = t.sym; out->band_ok = t.band_ok;
    for (int k = 0; k < 8; k++) out->far[k] = t.far[k];
    out->level_ptr = base + p.off[0]; out->perm = base + p.off[1]; out->a_col = base + p.off[2]; out->a_src = base + p.off[3];
    out->a_rev = base + p.off[4]; out->r_col = base + p.off[5]; out->r_src = base + p.off[6]; out->r_rev = base + p.off[7];
    out->c_lsrc = base + p.off[8]; out->c_lrev = base + p.off[9]; out->c_usrc = base + p.off[10];
    out->c_lfar = base + p.off[11]; out->c_ufar = base + p.off[12]; out->c_dsrc = base + p.off[13];
    out->m_nbr = base + p.off[14]; out->m_lfar = base + p.off[15]; out->m_ufar = base + p.off[16];
}

}  // namespace
}  // namespace dpiso

using namespace dpiso;

extern "C" {

int dpiso_bicg_tables_create(int ny, int nx, int per_x, int per_y, int comp, int transpose, dpiso_bicg_tables *out,
                             void *stream) {
    DPISO_REQUIRE(out && (comp == 0 || comp == 1), "bad arguments");
    HostTables t;
    const int rc = build_tables(ny, nx, per_x ? 1 : 0, per_y ? 1 : 0, comp, transpose ? 1 : 0, t);
    if (rc != DPISO_OK) return rc;
    const Packed p = pack(t);
    int *dev = nullptr;
    DPISO_CUDA_TRY(cudaMalloc((void **)&dev, p.data.size() * sizeof(int)));
    // pageable source: the call returns once the data has been staged, so `p` may go out of scope
    cudaError_t e = cudaMemcpyAsync(dev, p.data.data(), p.data.size() * sizeof(int), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { cudaFree(dev); set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(e)); return DPISO_ECUDA; }
    fill_struct(t, p, dev, out);
    out->owner = dev;
    out->owner_is_host = 0;
    return DPISO_OK;
}

int dpiso_bicg_tables_create_host(int ny, int nx, int per_x, int per_y, int comp, int transpose, dpiso_bicg_tables *out) {
    DPISO_REQUIRE(out && (comp == 0 || comp == 1), "bad arguments");
    HostTables t;
    const int rc = build_tables(ny, nx, per_x ? 1 : 0, per_y ? 1 : 0, comp, transpose ? 1 : 0, t);
    if (rc != DPISO_OK) return rc;
    const Packed p = pack(t);
    int *host = (int *)malloc(p.data.size() * sizeof(int));
    DPISO_REQUIRE(host, "out of memory");
    memcpy(host, p.data.data(), p.data.size() * sizeof(int));
    fill_struct(t, p, host, out);
    out->owner = host;
    out->owner_is_host = 1;
    return DPISO_OK;
}

int dpiso_bicg_tables_destroy(dpiso_bicg_tables *tab) {
    if (!tab || !tab->owner) return DPISO_OK;
    if (tab->owner_is_host) free(tab->owner);
    else DPISO_CUDA_TRY(cudaFree(tab->owner));
    memset(tab, 0, sizeof(*tab));
    return DPISO_OK;
}

}  // extern "C"
