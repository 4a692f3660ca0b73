"""CPU suite: CSR layout and BiCGStab tables against the oracle; the kernels' per-row code (compiled for the host)
against the oracle; C-ABI surface of libdpiso.so."""
import os
import re
import subprocess

import numpy as np
import pytest

import _host as H
from oracle import oracle as O
from diffpiso_b200 import structure as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRIDS = [(3, 3), (4, 5), (5, 4), (7, 6), (8, 8), (33, 32), (16, 24)]


@pytest.mark.parametrize("ny,nx", GRIDS)
@pytest.mark.parametrize("per_x", [0, 1])
@pytest.mark.parametrize("per_y", [0, 1])
def test_csr_structure_bit_exact(ny, nx, per_x, per_y):
    """row_ptr / col_ind: host tables == kernel row code == oracle (reference slot arithmetic), and the invariants of
    SURVEY 4(i): ascending columns, row_ptr[-1] == matrix_nnz formula."""
    orp, oci = O.csr_structure(ny, nx, per_x, per_y)
    n_u, n_v, z_u, z_v = O.sizes(ny, nx, per_x, per_y)
    assert S.sizes(ny, nx, per_x, per_y) == (n_u, n_v, z_u, z_v)
    srp, sci = S.csr_pattern(ny, nx, per_x, per_y)
    hrp, hci = H.csr_structure(ny, nx, per_x, per_y, n_u, n_v, z_u + z_v)
    for rp, ci in ((srp, sci), (hrp, hci)):
        assert rp.dtype == np.int32 and ci.dtype == np.int32
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
    assert orp[n_u] == z_u and orp[-1] == z_v and orp[0] == 0 and orp[n_u + 1] == 0
    for rp, ci in ((orp[:n_u + 1], oci[:z_u]), (orp[n_u + 1:], oci[z_u:])):
        for r in range(rp.size - 1):
            row = ci[rp[r]:rp[r + 1]]
            assert np.all(np.diff(row) > 0) and r in row


def _emulate_lu(tab, vals, b):
    """numpy emulation of the kernel's table-driven ILU(0) + L/U sweeps (level-major order = ascending position)."""
    n, wa, col, src, rev = tab["n"], tab["wa"], tab["a_col"], tab["a_src"], tab["a_rev"]
    f32 = np.float32
    a = np.where(src >= 0, vals[np.maximum(src, 0)], 0).astype(f32)

    def fma(x, y, z):
        return f32(np.float64(x) * np.float64(y) + np.float64(z))
    lu, piv = a.copy(), np.zeros(n, f32)
    dslot = [[k for k in range(wa) if col[k, q] == q][0] for q in range(n)]
    for q in range(n):
        d = a[dslot[q], q]
        for k in range(wa):
            if col[k, q] < q:
                lik = f32(a[k, q] / piv[col[k, q]])
                lu[k, q] = lik
                if rev[k, q] >= 0:
                    d = fma(-lik, vals[rev[k, q]], d)
        lu[dslot[q], q] = d
        piv[q] = d
    z = np.zeros(n, f32)
    bp = b[tab["perm"]]
    for q in range(n):
        acc = bp[q]
        for k in range(wa):
            if col[k, q] < q:
                acc = fma(-lu[k, q], z[col[k, q]], acc)
        z[q] = acc
    for q in range(n - 1, -1, -1):
        acc = z[q]
        for k in range(wa):
            if col[k, q] > q:
                acc = fma(-lu[k, q], z[col[k, q]], acc)
        z[q] = f32(acc / lu[dslot[q], q])
    out = np.zeros(n, f32)
    out[tab["perm"]] = z
    return out


@pytest.mark.parametrize("ny,nx,per_x,per_y", [(5, 6, 0, 0), (6, 5, 1, 1), (7, 6, 1, 0), (6, 7, 0, 1), (8, 8, 1, 1)])
def test_bicg_tables_reproduce_generic_ilu0(ny, nx, per_x, per_y):
    """The table-driven ILU(0) + triangular sweeps (the kernel's algorithm, emulated in numpy) are BIT-identical to the
    oracle's generic IKJ ILU(0) + CSR triangular solves, for A and for A^T (csr2csc + fresh ILU0 as the reference)."""
    rng = np.random.RandomState(0)
    rp, ci = S.csr_pattern(ny, nx, per_x, per_y)
    n_u, n_v, z_u, z_v = S.sizes(ny, nx, per_x, per_y)
    for comp in (0, 1):
        n = (n_u, n_v)[comp]
        rpc = rp[:n_u + 1] if comp == 0 else rp[n_u + 1:]
        cic = ci[:z_u] if comp == 0 else ci[z_u:]
        vals = (rng.randn(rpc[-1]) * 0.1).astype(np.float32)
        vals[cic == np.repeat(np.arange(n), np.diff(rpc))] += 2.0
        b = rng.randn(n).astype(np.float32)
        for tr in (False, True):
            tab = S.bicg_tables(ny, nx, per_x, per_y, comp, tr)
            assert tab["level_ptr"][-1] == n and tab["wa"] <= S.MAX_WA
            trp, tci, tval = O.csr_transpose(rpc, cic, vals) if tr else (rpc, cic, vals)
            lu, zero_pivot = O.ilu0(trp, tci, tval)
            assert zero_pivot == -1
            assert np.array_equal(_emulate_lu(tab, vals, b), O.lu_solve(trp, tci, lu, b))


def _emulate_rows(tab, vals, b):
    """numpy emulation of bicgstab_rows_kernel's canonical-slot ILU(0) + L/U sweeps in wavefront order: thread ly walks
    along x, lower slots [far, y-neighbour, far, x-neighbour], upper slots [x-neighbour, far, y-neighbour, far]."""
    n, dx = tab["n"], tab["dx"]
    dy = n // dx
    f32 = np.float32
    lsrc, lrev, usrc, lfar, ufar, dsrc = (tab[k] for k in ("c_lsrc", "c_lrev", "c_usrc", "c_lfar", "c_ufar", "c_dsrc"))

    def val(idx):
        return np.where(idx >= 0, vals[np.maximum(idx, 0)], 0).astype(f32)

    def fma(x, y, z):
        return f32(np.float64(x) * np.float64(y) + np.float64(z))
    alow, arv, uval, dg = val(lsrc), val(lrev), val(usrc), val(dsrc)
    piv, lval = np.zeros(n, f32), np.zeros((n, 4), f32)
    order = sorted(range(n), key=lambda i: (i % dx + i // dx, i))            # level by level
    for i in order:                                                           # ILU(0)
        lx = i % dx
        ops = [piv[lfar[i, 0]] if lfar[i, 0] >= 0 else f32(1), piv[i - dx] if i >= dx else f32(1),
               piv[lfar[i, 1]] if lfar[i, 1] >= 0 else f32(1), piv[i - 1] if lx > 0 else f32(1)]
        d = dg[i]
        for m in range(4):
            lik = f32(alow[i, m] / ops[m])
            lval[i, m] = lik
            d = fma(-lik, arv[i, m], d)
        piv[i] = d
    z = np.zeros(n, f32)
    for i in order:                                                           # L solve
        lx = i % dx
        ops = [z[lfar[i, 0]] if lfar[i, 0] >= 0 else f32(0), z[i - dx] if i >= dx else f32(0),
               z[lfar[i, 1]] if lfar[i, 1] >= 0 else f32(0), z[i - 1] if lx > 0 else f32(0)]
        acc = b[i]
        for m in range(4):
            acc = fma(-lval[i, m], ops[m], acc)
        z[i] = acc
    for i in reversed(order):                                                 # U solve
        lx = i % dx
        ops = [z[i + 1] if lx < dx - 1 else f32(0), z[ufar[i, 0]] if ufar[i, 0] >= 0 else f32(0),
               z[i + dx] if i + dx < n else f32(0), z[ufar[i, 1]] if ufar[i, 1] >= 0 else f32(0)]
        acc = z[i]
        for m in range(4):
            acc = fma(-uval[i, m], ops[m], acc)
        z[i] = f32(acc / piv[i])
    return z


@pytest.mark.parametrize("ny,nx,per_x,per_y", [(5, 6, 0, 0), (6, 5, 1, 1), (7, 6, 1, 0), (6, 7, 0, 1), (8, 8, 1, 1)])
def test_row_major_canonical_slots_reproduce_generic_ilu0(ny, nx, per_x, per_y):
    """The canonical-slot tables of the row-major BiCGStab kernel: ILU(0) + triangular sweeps emulated in numpy are
    BIT-identical to the oracle's generic IKJ ILU(0) + CSR triangular solves, for A and A^T."""
    rng = np.random.RandomState(1)
    rp, ci = S.csr_pattern(ny, nx, per_x, per_y)
    n_u, n_v, z_u, z_v = S.sizes(ny, nx, per_x, per_y)
    for comp in (0, 1):
        n = (n_u, n_v)[comp]
        rpc = rp[:n_u + 1] if comp == 0 else rp[n_u + 1:]
        cic = ci[:z_u] if comp == 0 else ci[z_u:]
        vals = (rng.randn(rpc[-1]) * 0.1).astype(np.float32)
        vals[cic == np.repeat(np.arange(n), np.diff(rpc))] += 2.0
        b = rng.randn(n).astype(np.float32)
        for tr in (False, True):
            tab = S.bicg_tables(ny, nx, per_x, per_y, comp, tr)
            assert tab["rows_ok"] == 1
            trp, tci, tval = O.csr_transpose(rpc, cic, vals) if tr else (rpc, cic, vals)
            lu, zero_pivot = O.ilu0(trp, tci, tval)
            assert zero_pivot == -1
            assert np.array_equal(_emulate_rows(tab, vals, b), O.lu_solve(trp, tci, lu, b))


def test_tables_reject_degenerate_grid():
    with pytest.raises(NotImplementedError):
        S.bicg_tables(3, 3, True, True, 0, False)
    with pytest.raises(ValueError):
        S.sizes(2, 5, False, False)


@pytest.mark.parametrize("ny,nx,per_x,per_y", [(4, 5, 0, 0), (5, 4, 1, 1), (6, 7, 1, 0), (7, 6, 0, 1), (33, 32, 0, 0)])
def test_kernel_row_code_matches_oracle(ny, nx, per_x, per_y):
    """assemble / FV gradient / divergence / Laplace rows as compiled from csrc/rows.cuh == oracle, bit for bit, with
    random masks, Dirichlet rows, scalar and per-face viscosity and every ghost-cell rule."""
    rng = np.random.RandomState(1)
    n_u, n_v, z_u, z_v = O.sizes(ny, nx, per_x, per_y)
    orp, _ = O.csr_structure(ny, nx, per_x, per_y)
    u = rng.randn(ny, nx + 1).astype(np.float32)
    v = rng.randn(ny + 1, nx).astype(np.float32)
    vel = np.concatenate([u.ravel(), v.ravel()])
    dirich = (rng.rand(n_u + n_v) < 0.1).astype(np.uint8)
    nm = (ny + 2) * (nx + 2)
    active = (rng.rand(nm) < 0.8).astype(np.float32)
    access = (rng.rand(nm) < 0.8).astype(np.float32)
    noslip = (rng.rand(nm) < 0.3).astype(np.uint8)
    up, vp = O.pad_velocity(ny, nx, per_x, per_y, u, v)
    for visc in (np.float32(0.01), (rng.rand(n_u + n_v) * 0.01).astype(np.float32)):
        ov, oa = O.assemble(ny, nx, per_x, per_y, 0.1, 0.13, 2.5, up, vp, dirich, active, noslip, visc, orp)
        hv, ha = H.assemble(ny, nx, per_x, per_y, 0.1, 0.13, 2.5, vel, dirich, active, noslip, visc, z_u + z_v)
        assert np.array_equal(ov, hv) and np.array_equal(oa, ha)
    p = rng.randn(ny * nx).astype(np.float32)
    for pbc in ([0, 0, 0, 0], [1, 1, 1, 1], [2, 2, 2, 2], [0, 1, 2, 2], [1, 0, 0, 1]):
        assert np.array_equal(O.fv_gradient(ny, nx, 0.1, 0.13, pbc, access, p), H.fv_gradient(ny, nx, 0.1, 0.13, pbc, access, p))
    assert np.array_equal(O.fv_divergence(ny, nx, 0.1, 0.13, vel), H.fv_divergence(ny, nx, 0.1, 0.13, vel))
    k = rng.rand(n_u + n_v).astype(np.float32)
    for dt in (np.float64, np.float32):
        assert np.array_equal(O.laplace(ny, nx, active, access, k, dt), H.laplace(ny, nx, active, access, k, dt))


def _header_decls():
    h = open(os.path.join(ROOT, "include", "dpiso.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return {m.group(1): m.group(2).strip() for m in
            re.finditer(r"\b(?:int|size_t|const char \*)\s*(dpiso_\w+)\s*\(([^)]*)\)\s*;", h)}


def test_cabi_exports_every_declared_symbol():
    """libdpiso.so loads without a GPU and exports exactly the entry points include/dpiso.h declares; the ctypes
    signatures of the Python binding agree with the header argument by argument."""
    from diffpiso_b200 import _native as N
    decls = _header_decls()
    assert len(decls) >= 20
    nm = subprocess.run(["nm", "-D", N._LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dpiso_\w+)", nm))
    assert exported == set(decls)
    assert set(N.EXPORTS) == set(decls)
    assert N.lib.dpiso_version() == 100
    kind = {N._I: "I", N._F: "F", N._P: "P"}
    for name, args in decls.items():
        want = "" if args == "void" else "".join(
            "P" if "*" in a else ("F" if a.strip().startswith("float") else "I") for a in args.split(","))
        got = "".join(kind[t] for t in N._SIGS[name][0])
        assert got == want, name


def test_cabi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call: bad sizes return DPISO_EINVAL with a message."""
    import ctypes as C
    from diffpiso_b200 import _native as N
    n, z = (C.c_int * 2)(), (C.c_int * 2)()
    assert N.lib.dpiso_sizes(2, 2, 0, 0, n, z) == -1
    assert b"grid too small" in N.lib.dpiso_last_error()
    assert N.lib.dpiso_sizes(33, 32, 0, 0, n, z) == 0
    assert (n[0], n[1], z[0], z[1]) == O.sizes(33, 32, 0, 0)
    assert N.lib.dpiso_sizes(128, 128, 1, 1, n, z) == 0
    assert (n[0], n[1], z[0], z[1]) == (16512, 16512, 82560, 82560)
    # solver entry points: empty batch, null pointers, bad cadence -> EINVAL before anything is launched
    null = None
    assert N.lib.dpiso_pressure_cg_f64(0, 16, 16, 1, 1, null, null, 1e-8, 10, 10, 1, null, null, null, null, null) == -1
    assert b"bad sizes" in N.lib.dpiso_last_error()
    assert N.lib.dpiso_pressure_cg_f64(1, 16, 16, 1, 1, null, null, 1e-8, 10, 10, 1, null, null, null, null, null) == -1
    assert b"null pointer" in N.lib.dpiso_last_error()
    assert N.lib.dpiso_bicgstab_ilu(0, null, null, 0, 0, null, 0, null, null, 1e-8, 10, null, null, null, null, null, null, null) == -1
    assert N.lib.dpiso_assemble(0, 16, 16, 0, 0, 1.0, 1.0, 1.0, 1.0, 1.0, null, null, null, null, null, 0, null, null, null) == -1
    assert b"batch must be >= 1" in N.lib.dpiso_last_error()


def test_product_never_imports_oracle():
    """The product package must not reference the oracle or carry a CPU fallback."""
    pkg = os.path.join(ROOT, "differentiable-piso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "piso_oracle" not in txt and "from oracle" not in txt, f


@pytest.mark.parametrize("per_y,per_x", [(0, 0), (1, 1), (1, 0), (0, 1)])
def test_custom_padded_matches_oracle(per_y, per_x):
    """diffpiso_b200.custom_padded (torch indexing, runs on CPU tensors) == the oracle's restatement of
    piso_helpers.py:35-55, including the dropped duplicate face on a component's own periodic axis."""
    import torch
    import diffpiso_b200 as dp
    ny, nx = 5, 6
    rng = np.random.RandomState(0)
    u = rng.randn(ny, nx + 1).astype(np.float32)
    v = rng.randn(ny + 1, nx).astype(np.float32)
    st = dp.stack_staggered_components([torch.as_tensor(v)[None, ..., None], torch.as_tensor(u)[None, ..., None]])
    vp, up = dp.custom_padded(dp.StaggeredGrid(st), 1, (per_y, per_x))
    oup, ovp = O.pad_velocity(ny, nx, per_x, per_y, u, v)
    assert np.array_equal(up[0].numpy(), oup) and np.array_equal(vp[0].numpy(), ovp)
    flat = dp.flatten_staggered_data(st, coord_flip=True)[0].numpy()
    assert np.array_equal(flat, np.concatenate([u.ravel(), v.ravel()]))
    back = dp.stagger_flattened_data(torch.as_tensor(flat), (1, ny + 1, nx + 1, 2), coord_flip=True)
    assert torch.equal(back, st)


@pytest.mark.parametrize("ny,nx", [(4, 5), (7, 6), (8, 8), (33, 32), (16, 24), (128, 128)])
@pytest.mark.parametrize("per_x", [0, 1])
@pytest.mark.parametrize("per_y", [0, 1])
def test_native_table_builder_equals_numpy_derivation(ny, nx, per_x, per_y):
    """dpiso_bicg_tables_create_host (csrc/tables.cu, what a C caller uses) against diffpiso_b200/structure.py (an
    independent numpy / scipy derivation), entry by entry, for A and A^T of both components; `sym` is the structural
    symmetry of the pattern (lost exactly for a component that is periodic along its staggered axis)."""
    import ctypes as C
    from diffpiso_b200 import _native as N
    for comp in (0, 1):
        for transpose in (0, 1):
            st = N.BicgTables()
            try:
                want = S.bicg_tables(ny, nx, bool(per_x), bool(per_y), comp, bool(transpose))
            except NotImplementedError:
                assert N.lib.dpiso_bicg_tables_create_host(ny, nx, per_x, per_y, comp, transpose, C.byref(st)) == -3
                continue
            assert N.lib.dpiso_bicg_tables_create_host(ny, nx, per_x, per_y, comp, transpose, C.byref(st)) == 0, \
                N.lib.dpiso_last_error()
            try:
                for k in ("n", "n_levels", "wa", "max_level", "wl", "wu", "dx", "rows_ok"):
                    assert getattr(st, k) == want[k], k
                assert st.owner_is_host == 1
                assert st.sym == int(not (per_x if comp == 0 else per_y))
                # closed form of the periodic wrap operands (cluster-per-system kernel): {xa, xb, ya, yb} lower, upper
                if want["rows_ok"]:
                    Dx, Dy, sx, sy = S.comp_dims(ny, nx, comp)
                    if not transpose:
                        exp = [Dx - 1, sx, Dy - 1, sy, 0, Dx - 1 - sx, 0, Dy - 1 - sy]
                    else:
                        exp = [Dx - 1 - sx, 0, Dy - 1 - sy, 0, sx, Dx - 1, sy, Dy - 1]
                    for k in range(0, 8, 2):
                        if not (per_x if k % 4 == 0 else per_y):
                            exp[k] = exp[k + 1] = -1
                    assert st.band_ok == 1 and list(st.far) == exp, (list(st.far), exp, comp, transpose)
                    # ... and it describes every far entry of the numpy derivation, and only those
                    ly_, lx_ = np.divmod(np.arange(st.n), Dx)
                    for d, cf in enumerate((want["c_lfar"], want["c_ufar"])):
                        row_k, col_k = (1, 0) if d == 0 else (0, 1)
                        f = exp[4 * d:4 * d + 4]
                        assert np.array_equal(cf[:, row_k] >= 0, lx_ == f[0]) and np.array_equal(cf[:, col_k] >= 0, ly_ == f[2])
                        hr, hc = cf[:, row_k] >= 0, cf[:, col_k] >= 0
                        assert np.array_equal(cf[hr, row_k], ly_[hr] * Dx + f[1]) and np.array_equal(cf[hc, col_k], f[3] * Dx + lx_[hc])
                else:
                    assert st.band_ok == 0
                n, wa = st.n, st.wa
                shapes = dict(level_ptr=st.n_levels + 1, perm=n, a_col=wa * n, a_src=wa * n, a_rev=wa * n, r_col=wa * n,
                              r_src=wa * n, r_rev=wa * n, c_lsrc=4 * n, c_lrev=4 * n, c_usrc=4 * n, c_lfar=2 * n,
                              c_ufar=2 * n, c_dsrc=n, m_nbr=4 * n, m_lfar=2 * n, m_ufar=2 * n)
                for k, size in shapes.items():
                    got = np.ctypeslib.as_array(C.cast(getattr(st, k), C.POINTER(C.c_int)), shape=(size,))
                    assert np.array_equal(got, np.asarray(want[k]).ravel()), (k, comp, transpose)
            finally:
                assert N.lib.dpiso_bicg_tables_destroy(C.byref(st)) == 0
            assert not st.owner
