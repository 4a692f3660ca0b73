"""Host-side structure tables for the batched ILU0-BiCGStab kernel (csrc/bicgstab.cu).

The sparsity pattern of the advection-diffusion matrices depends only on (ny, nx, periodic flags), so everything here
is computed once per grid with numpy/scipy, uploaded, and cached:

* `csr_pattern`   -- the reference's CSR layout (CUDAsrc/central_difference_csr_op.cu.cc:472-505, :162-231),
                     derived independently (existing neighbours + self, sorted by column).  tests/ compare it
                     bit-for-bit with the device kernel `dpiso_csr_structure` and with the oracle.
* `bicg_tables`   -- level-major ELL tables of M = A (forward) or M = A^T (adjoint; the reference transposes with
                     csr2csc and factorises again, multi_bicgstab_ilu_linear_solve_op.cu.cc:113-134): wavefront levels
                     lx+ly, permutation, per-entry column / value-source / reverse-entry indices.  The builder also
                     PROVES, for the concrete grid, the two properties the kernel relies on: (i) every lower entry
                     points to a strictly lower level and every upper entry to a strictly higher one, (ii) ILU(0) on
                     this pattern only changes the pivots and the lower entries (no fill interaction).
"""
import functools

import numpy as np
import scipy.sparse as sp

MAX_WA = 6


def comp_dims(ny, nx, comp):
    """(Dx, Dy, stag_x, stag_y); comp 0 = u (x-staggered), 1 = v (y-staggered)."""
    return (nx + 1, ny, 1, 0) if comp == 0 else (nx, ny + 1, 0, 1)


def sizes(ny, nx, per_x, per_y):
    """n_u, n_v, nnz_u, nnz_v (diffpiso/piso_tf.py:99-106)."""
    if ny < 3 or nx < 3:
        raise ValueError("grid too small for the 5-point pattern (need ny, nx >= 3)")
    out_n, out_z = [], []
    for comp in (0, 1):
        Dx, Dy, _, _ = comp_dims(ny, nx, comp)
        n = Dx * Dy
        out_n.append(n)
        out_z.append(5 * n - 2 * (n // Dx) * (1 - int(per_x)) - 2 * (n // Dy) * (1 - int(per_y)))
    return out_n[0], out_n[1], out_z[0], out_z[1]


def _component_pattern(ny, nx, per_x, per_y, comp):
    """scipy CSR matrix of one component whose data are the CSR positions + 1."""
    Dx, Dy, sx, sy = comp_dims(ny, nx, comp)
    n = Dx * Dy
    ly, lx = np.divmod(np.arange(n, dtype=np.int64), Dx)
    row = np.arange(n, dtype=np.int64)
    rows, cols = [row], [row]
    # x-, x+, y-, y+ : regular neighbour, or the periodic wrap that skips the duplicated staggered face
    for reg, col_reg, col_wrap, per in (
            (lx > 0, row - 1, row + (Dx - 1 - sx), per_x),
            (lx < Dx - 1, row + 1, row - (Dx - 1 - sx), per_x),
            (ly > 0, row - Dx, row + Dx * (Dy - 1 - sy), per_y),
            (ly < Dy - 1, row + Dx, row - Dx * (Dy - 1 - sy), per_y)):
        has = reg | bool(per)
        rows.append(row[has])
        cols.append(np.where(reg, col_reg, col_wrap)[has])
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    m = sp.csr_matrix((np.ones(rows.size, np.int64), (rows, cols)), shape=(n, n))
    m.sum_duplicates()
    m.sort_indices()
    if m.nnz != rows.size:
        raise ValueError("degenerate grid: two neighbours of a face coincide")
    m.data = np.arange(1, m.nnz + 1, dtype=np.int64)
    return m


@functools.lru_cache(maxsize=64)
def csr_pattern(ny, nx, per_x, per_y):
    """row_ptr (n_u+1 and n_v+1 back to back, each 0-based) and col_ind (0-based per component), int32."""
    mats = [_component_pattern(ny, nx, bool(per_x), bool(per_y), c) for c in (0, 1)]
    row_ptr = np.concatenate([m.indptr for m in mats]).astype(np.int32)
    col_ind = np.concatenate([m.indices for m in mats]).astype(np.int32)
    return row_ptr, col_ind


def _lookup(keys_sorted, vals, query):
    idx = np.searchsorted(keys_sorted, query)
    idx = np.minimum(idx, keys_sorted.size - 1)
    found = keys_sorted[idx] == query
    return np.where(found, vals[idx], -1), found


@functools.lru_cache(maxsize=64)
def bicg_tables(ny, nx, per_x, per_y, comp, transpose):
    """Level-major ELL tables of M = A or A^T for one component; dict of numpy arrays and ints."""
    Dx, Dy, _, _ = comp_dims(ny, nx, comp)
    n = Dx * Dy
    a = _component_pattern(ny, nx, bool(per_x), bool(per_y), comp)
    m = a.T.tocsr() if transpose else a
    m.sort_indices()
    rp, ci, src = m.indptr.astype(np.int64), m.indices.astype(np.int64), m.data.astype(np.int64) - 1
    row_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    lens = np.diff(rp)
    wa = int(lens.max())
    if wa > MAX_WA:
        raise NotImplementedError("row with %d entries" % wa)
    # reverse entries M(col, row)
    keys = row_of * n + ci                       # sorted (CSR order)
    rev, _ = _lookup(keys, src, ci * n + row_of)
    # (i) wavefront property of the lx+ly levels
    ly, lx = np.divmod(np.arange(n, dtype=np.int64), Dx)
    level = lx + ly
    lower, upper = ci < row_of, ci > row_of
    if not (np.all(level[ci[lower]] < level[row_of[lower]]) and np.all(level[ci[upper]] > level[row_of[upper]])):
        raise NotImplementedError("lx+ly is not a valid level schedule for this grid")
    # ELL in original numbering: slot k of row i = k-th entry in ascending column order
    slot = np.arange(ci.size, dtype=np.int64) - rp[row_of]
    col_e = np.full((wa, n), -1, np.int64)
    src_e = np.full((wa, n), -1, np.int64)
    rev_e = np.full((wa, n), -1, np.int64)
    col_e[slot, row_of] = ci
    src_e[slot, row_of] = src
    rev_e[slot, row_of] = rev
    # (ii) ILU(0) touches only pivots and lower entries: for k in L(i) and j in row i with j > k, j != i,
    #      (k, j) must not be an entry of M
    rows_all = np.arange(n, dtype=np.int64)
    for s1 in range(wa):
        for s2 in range(s1 + 1, wa):
            k, j = col_e[s1], col_e[s2]
            cand = (k >= 0) & (j >= 0) & (k < rows_all) & (j != rows_all)
            if cand.any():
                _, hit = _lookup(keys, src, k[cand] * n + j[cand])
                if hit.any():
                    raise NotImplementedError("ILU(0) fill interaction on this grid (too small / degenerate)")
    # level-major permutation
    perm = np.argsort(level, kind="stable")
    pos = np.empty(n, np.int64)
    pos[perm] = np.arange(n)
    counts = np.bincount(level, minlength=int(level.max()) + 1)
    level_ptr = np.concatenate([[0], np.cumsum(counts)])
    q = np.arange(n, dtype=np.int64)
    colp = col_e[:, perm]
    a_col = np.where(colp >= 0, pos[np.maximum(colp, 0)], q[None, :])
    a_src = src_e[:, perm]
    a_rev = rev_e[:, perm]
    # after the permutation lower entries precede the row, upper entries follow it
    real = a_src >= 0
    orig_row = perm[None, :].repeat(wa, 0)
    assert np.all((a_col < q[None, :])[real & (colp < orig_row)])
    assert np.all((a_col > q[None, :])[real & (colp > orig_row)])
    # Row-major variant of the solver (bicgstab_rows_kernel): every row keeps its entries in canonical slots by kind --
    # lower: [far below the y-neighbour, y-neighbour (x, y-1), far above it, x-neighbour (x-1, y)], upper mirrored -- which
    # preserves the ascending-column order only if each far slot is used at most once per row.
    regular = ((ci == row_of - 1) & (lx[row_of] > 0)) | ((ci == row_of + 1) & (lx[row_of] < Dx - 1)) | \
              (ci == row_of - Dx) | (ci == row_of + Dx) | (ci == row_of)
    far = ~regular
    slot_kind = np.where(ci < row_of,
                         np.where(far, np.where(ci < row_of - Dx, 0, 2), np.where(ci == row_of - Dx, 1, 3)),
                         np.where(far, np.where(ci > row_of + Dx, 7, 5), np.where(ci == row_of + Dx, 6, 4)))
    offdiag = ci != row_of
    rows_ok = int(np.unique(row_of[offdiag] * 8 + slot_kind[offdiag]).size == int(offdiag.sum()))
    # the sweeps read far operands one level ahead of their use: they must be at least two levels old
    if far.any() and int(np.abs(level[ci[far]] - level[row_of[far]]).min()) < 2:
        rows_ok = 0
    # decoupled-warp sweeps: a far operand lies in the same grid row (same sweep thread) or in the same grid column
    if far.any() and not np.all((ly[ci[far]] == ly[row_of[far]]) | (lx[ci[far]] == lx[row_of[far]])):
        rows_ok = 0
    # ... and, when it comes from another row, from at least three rows away (it is fetched one step ahead of its use)
    if far.any():
        drow = np.abs(ly[ci[far]] - ly[row_of[far]])
        if np.any((drow > 0) & (drow < 3)):
            rows_ok = 0
    # canonical source tables of the row-major kernel: CSR value index of every slot (-1 = absent), reverse entries of the
    # lower slots, the pivot's index, and the columns of the far slots
    c_lsrc = np.full((n, 4), -1, np.int32); c_lrev = np.full((n, 4), -1, np.int32); c_usrc = np.full((n, 4), -1, np.int32)
    c_lfar = np.full((n, 2), -1, np.int32); c_ufar = np.full((n, 2), -1, np.int32); c_dsrc = np.full(n, -1, np.int32)
    if rows_ok:
        lo_e, up_e = offdiag & (ci < row_of), offdiag & (ci > row_of)
        c_lsrc[row_of[lo_e], slot_kind[lo_e]] = src[lo_e]
        c_lrev[row_of[lo_e], slot_kind[lo_e]] = rev[lo_e]
        c_usrc[row_of[up_e], slot_kind[up_e] - 4] = src[up_e]
        fl = lo_e & far
        c_lfar[row_of[fl], slot_kind[fl] // 2] = ci[fl]
        fu = up_e & far
        c_ufar[row_of[fu], (slot_kind[fu] - 4) // 2] = ci[fu]
        dg_e = ci == row_of
        c_dsrc[row_of[dg_e]] = src[dg_e]
    # level-major positions for the row-major kernel: regular neighbours (x-1, y-1, x+1, y+1; the row itself where absent)
    # and the far slots
    orig = perm
    olx, oly = lx[orig], ly[orig]
    qq = np.arange(n, dtype=np.int64)
    m_nbr = np.stack([np.where(olx > 0, pos[np.maximum(orig - 1, 0)], qq),
                      np.where(oly > 0, pos[np.maximum(orig - Dx, 0)], qq),
                      np.where(olx < Dx - 1, pos[np.minimum(orig + 1, n - 1)], qq),
                      np.where(oly < Dy - 1, pos[np.minimum(orig + Dx, n - 1)], qq)], axis=1).astype(np.int32)
    m_lfar = np.where(c_lfar[orig] >= 0, pos[np.maximum(c_lfar[orig], 0)], -1).astype(np.int32)
    m_ufar = np.where(c_ufar[orig] >= 0, pos[np.maximum(c_ufar[orig], 0)], -1).astype(np.int32)
    wl = int(np.bincount(row_of[lower], minlength=n).max()) if lower.any() else 0
    wu = int(np.bincount(row_of[upper], minlength=n).max()) if upper.any() else 0
    return dict(n=n, n_levels=int(counts.size), wa=wa, max_level=int(counts.max()), wl=wl, wu=wu, dx=int(Dx), rows_ok=rows_ok,
                r_col=np.ascontiguousarray(np.where(col_e >= 0, col_e, rows_all[None, :]), np.int32),
                r_src=np.ascontiguousarray(src_e, np.int32), r_rev=np.ascontiguousarray(rev_e, np.int32),
                c_lsrc=c_lsrc, c_lrev=c_lrev, c_usrc=c_usrc, c_lfar=c_lfar, c_ufar=c_ufar, c_dsrc=c_dsrc,
                m_nbr=m_nbr, m_lfar=m_lfar, m_ufar=m_ufar,
                level_ptr=level_ptr.astype(np.int32), perm=perm.astype(np.int32),
                a_col=np.ascontiguousarray(a_col, np.int32), a_src=np.ascontiguousarray(a_src, np.int32),
                a_rev=np.ascontiguousarray(a_rev, np.int32), nnz=int(a.nnz))
