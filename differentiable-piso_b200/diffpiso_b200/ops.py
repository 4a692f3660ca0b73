"""Thin torch-tensor front-ends of the C ABI (one function per entry point) plus the cached per-grid state.

Everything here allocates outputs with torch and enqueues the native kernel on the current torch stream.  No
arithmetic happens in Python and there is no fallback path.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import structure as S


class Geometry(object):
    """Static per-grid state on one device: sizes, periodic flags, device copies of the BiCGStab tables."""

    _cache = {}

    def __init__(self, ny, nx, per_y, per_x, device):
        self.ny, self.nx, self.per_y, self.per_x = int(ny), int(nx), bool(per_y), bool(per_x)
        self.device = torch.device(device)
        self.n_u, self.n_v, self.nnz_u, self.nnz_v = S.sizes(self.ny, self.nx, self.per_x, self.per_y)
        self.nf, self.nc, self.nnz = self.n_u + self.n_v, self.ny * self.nx, self.nnz_u + self.nnz_v
        self._tables = {}
        self._scratch = {}
        self._csr = None

    @classmethod
    def get(cls, ny, nx, per_y, per_x, device):
        key = (int(ny), int(nx), bool(per_y), bool(per_x), str(torch.device(device)))
        g = cls._cache.get(key)
        if g is None:
            g = cls._cache[key] = cls(ny, nx, per_y, per_x, device)
        return g

    def tables(self, transpose):
        """(BicgTables_u, BicgTables_v) ctypes structs whose pointers reference cached device tensors."""
        with _DeviceGuard(self.device):
            return self._tables_impl(transpose)

    def _tables_impl(self, transpose):
        t = self._tables.get(bool(transpose))
        if t is None:
            structs = []
            for comp in (0, 1):
                st = N.BicgTables()
                # built on the host by the library itself (csrc/tables.cu) and uploaded once per grid and device
                N.check(N.lib.dpiso_bicg_tables_create(self.ny, self.nx, int(self.per_x), int(self.per_y), comp,
                                                       int(bool(transpose)), C.byref(st), N.stream()),
                        "dpiso_bicg_tables_create")
                structs.append(st)
            t = self._tables[bool(transpose)] = (structs[0], structs[1])
        return t[0], t[1]

    def scratch(self, key, nbytes):
        """Per-grid scratch buffer (uint8) reused by successive calls on the same stream; `key` names the consumer."""
        k = (key, torch.cuda.current_stream(self.device).cuda_stream)
        buf = self._scratch.get(k)
        if buf is None or buf.numel() < nbytes:
            buf = self._scratch[k] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return buf

    def csr_structure(self):
        """(row_ptr, col_ind) int32 device tensors in the reference layout, from the device kernel."""
        if self._csr is None:
            with _DeviceGuard(self.device):
                rp = torch.empty(self.nf + 2, dtype=torch.int32, device=self.device)
                ci = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
                N.check(N.lib.dpiso_csr_structure(self.ny, self.nx, int(self.per_x), int(self.per_y), N.ptr(rp),
                                                  N.ptr(ci), N.stream()), "dpiso_csr_structure")
            self._csr = (rp, ci)
        return self._csr


class _DeviceGuard(object):
    """Makes the tensor's device current for the duration of a native call: the C entry points launch on whatever
    cudaGetDevice() reports and N.stream() returns the current stream of the current device, so tensors that live on
    another device than the caller's current one would otherwise be touched by kernels of the wrong device."""
    __slots__ = ("dev", "prev")

    def __init__(self, dev):
        self.dev, self.prev = dev, None

    def __enter__(self):
        if self.dev.type != "cuda":
            raise N.DpisoError("libdpiso needs CUDA tensors (got a %s tensor): there is no CPU path" % self.dev)
        idx = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
        cur = torch.cuda.current_device()
        if idx != cur:
            self.prev = cur
            torch.cuda.set_device(idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _on_device_of(index):
    """Decorator: argument `index` (a tensor, or a Geometry for index == 0 of the pure-geometry calls) names the device."""
    def deco(fn):
        def wrapped(*args, **kwargs):
            a = args[index]
            dev = a.device if hasattr(a, "device") else torch.device("cuda", torch.cuda.current_device())
            with _DeviceGuard(torch.device(dev)):
                return fn(*args, **kwargs)
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _f32(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


def cell_areas(dy64, dx64):
    """The op's `cell_area` input (piso_tf.py:97): prod(dx) / dx[::-1].astype(float32), rounded to fp32 ->
    (area of a face normal to x, area of a face normal to y)."""
    prod = float(dy64) * float(dx64)
    return float(np.float32(prod / float(np.float32(dx64)))), float(np.float32(prod / float(np.float32(dy64))))


@_on_device_of(0)
def assemble(g, vel, dirichlet_u8, active, noslip_u8, visc, dy, dx, beta, areas=None, vel_periodic=None):
    """-> values [B, nnz], a_diag [B, nf]   (advection_matrix_cuda, diffpiso/piso_tf.py:85-137)
    vel_periodic = (y, x): whether the velocity GRID says periodic on that axis (custom_padded pads by it); defaults to the
    geometry's flags.  False on a periodic axis reproduces the replicated padding of the reference's unrolled steps."""
    vpy, vpx = (g.per_y, g.per_x) if vel_periodic is None else vel_periodic
    fx = int(g.per_x) | (2 if (g.per_x and not vpx) else 0)
    fy = int(g.per_y) | (2 if (g.per_y and not vpy) else 0)
    area_x, area_y = cell_areas(dy, dx) if areas is None else areas
    vel = _f32(vel)
    b = vel.shape[0]
    visc = _f32(visc).reshape(-1) if visc.dim() < 2 else _f32(visc)
    if visc.numel() == 1:
        mode = 0
    elif visc.dim() == 1 and visc.numel() == g.nf:
        mode = 1
    elif visc.dim() == 2 and tuple(visc.shape) == (b, g.nf):
        mode = 2
    elif visc.dim() == 2 and tuple(visc.shape) == (1, g.nf):
        mode, visc = 1, visc.reshape(-1)
    else:
        raise ValueError("viscosity must be a scalar or a flat [u, v] face field")
    values = torch.empty((b, g.nnz), dtype=torch.float32, device=vel.device)
    a_diag = torch.empty((b, g.nf), dtype=torch.float32, device=vel.device)
    N.check(N.lib.dpiso_assemble(b, g.ny, g.nx, fx, fy, dy, dx, area_x, area_y, beta, N.ptr(vel), N.ptr(dirichlet_u8),
                                 N.ptr(active), N.ptr(noslip_u8), N.ptr(visc), mode, N.ptr(values), N.ptr(a_diag),
                                 N.stream()), "dpiso_assemble")
    return values, a_diag


@_on_device_of(0)
def predictor_rhs(g, vel, pres, access, dirichlet_u8, dvals, forcing, dy, dx, beta, pbc):
    vel, pres, dvals = _f32(vel), _f32(pres), _f32(dvals)
    b = vel.shape[0]
    forcing = None if forcing is None else _f32(forcing)
    if dvals.dim() != 2 or dvals.shape[0] not in (1, b) or dvals.shape[1] != g.nf:
        raise ValueError("dirichlet values must be [1 or B, n_u+n_v]")
    rhs = torch.empty_like(vel)
    N.check(N.lib.dpiso_predictor_rhs(b, g.ny, g.nx, dy, dx, beta, N.int4(pbc), N.ptr(vel), N.ptr(pres), N.ptr(access),
                                      N.ptr(dirichlet_u8), N.ptr(dvals), int(dvals.shape[0] == b), N.ptr(forcing),
                                      N.ptr(rhs), N.stream()), "dpiso_predictor_rhs")
    return rhs


@_on_device_of(0)
def fv_gradient(g, p, access, dy, dx, pbc):
    p = _f32(p)
    b = p.shape[0]
    out = torch.empty((b, g.nf), dtype=torch.float32, device=p.device)
    N.check(N.lib.dpiso_fv_gradient(b, g.ny, g.nx, dy, dx, N.int4(pbc), N.ptr(access), N.ptr(p), N.ptr(out), N.stream()),
            "dpiso_fv_gradient")
    return out


@_on_device_of(0)
def fv_divergence(g, vel, dy, dx, a_diag=None, beta=0.0):
    vel = _f32(vel)
    b = vel.shape[0]
    out = torch.empty((b, g.nc), dtype=torch.float32, device=vel.device)
    N.check(N.lib.dpiso_fv_divergence(b, g.ny, g.nx, dy, dx, N.ptr(vel), N.ptr(a_diag), beta, N.ptr(out), N.stream()),
            "dpiso_fv_divergence")
    return out


@_on_device_of(0)
def corrector1(g, u_star, p1, a_diag, access, dy, dx, beta, pbc):
    out = torch.empty_like(u_star)
    N.check(N.lib.dpiso_corrector1(u_star.shape[0], g.ny, g.nx, dy, dx, beta, N.int4(pbc), N.ptr(access), N.ptr(u_star),
                                   N.ptr(_f32(p1)), N.ptr(a_diag), N.ptr(out), N.stream()), "dpiso_corrector1")
    return out


@_on_device_of(0)
def h_apply(g, values, a_diag, u_star, u_s2, beta):
    out = torch.empty_like(u_star)
    N.check(N.lib.dpiso_h_apply(u_star.shape[0], g.ny, g.nx, int(g.per_x), int(g.per_y), beta, N.ptr(values),
                                N.ptr(a_diag), N.ptr(u_star), N.ptr(u_s2), N.ptr(out), N.stream()), "dpiso_h_apply")
    return out


@_on_device_of(0)
def corrector2(g, u_s2, h, p2, a_diag, p, p1, access, dy, dx, beta, pbc):
    u_next = torch.empty_like(u_s2)
    p_next = torch.empty_like(p)
    N.check(N.lib.dpiso_corrector2(u_s2.shape[0], g.ny, g.nx, dy, dx, beta, N.int4(pbc), N.ptr(access), N.ptr(u_s2),
                                   N.ptr(h), N.ptr(_f32(p2)), N.ptr(a_diag), N.ptr(_f32(p)), N.ptr(_f32(p1)),
                                   N.ptr(u_next), N.ptr(p_next), N.stream()), "dpiso_corrector2")
    return u_next, p_next


@_on_device_of(0)
def fv_gradient_adj(g, gs, access, dy, dx, pbc, a_diag=None, beta=0.0, divisor=1.0, negate=False, base=None):
    gs = _f32(gs)
    b = gs.shape[0]
    out = torch.empty((b, g.nc), dtype=torch.float32, device=gs.device)
    N.check(N.lib.dpiso_fv_gradient_adj(b, g.ny, g.nx, dy, dx, N.int4(pbc), N.ptr(access), N.ptr(gs), N.ptr(a_diag), beta,
                                        divisor, int(negate), N.ptr(None if base is None else _f32(base)), N.ptr(out),
                                        N.stream()), "dpiso_fv_gradient_adj")
    return out


@_on_device_of(0)
def fv_divergence_adj(g, gc, dy, dx, base=None, a_diag=None, beta=0.0, base_sub=None, vel_periodic=None):
    """([base [- base_sub]] + D^T gc) [/ (beta - a_diag)]; vel_periodic = (y, x): which registered gradient of
    finite_volume_divergence applies (it follows the velocity grid's extrapolation, piso_helpers.py:291-305)."""
    gc = _f32(gc)
    b = gc.shape[0]
    out = torch.empty((b, g.nf), dtype=torch.float32, device=gc.device)
    vpy, vpx = (g.per_y, g.per_x) if vel_periodic is None else vel_periodic
    N.check(N.lib.dpiso_fv_divergence_adj(b, g.ny, g.nx, int(g.per_x and vpx), int(g.per_y and vpy), dy, dx, N.ptr(gc),
                                          N.ptr(None if base is None else _f32(base)),
                                          N.ptr(None if base_sub is None else _f32(base_sub)), N.ptr(a_diag), beta,
                                          N.ptr(out), N.stream()), "dpiso_fv_divergence_adj")
    return out


@_on_device_of(0)
def h_apply_adj(g, values, a_diag, gh, beta, base=None):
    """gd = M_off^T gh; with `base` also returns base + gd (one pass)."""
    gh = _f32(gh)
    tu, tv = g.tables(True)
    out = torch.empty_like(gh)
    total = None if base is None else torch.empty_like(gh)
    N.check(N.lib.dpiso_h_apply_adj(gh.shape[0], C.byref(tu), C.byref(tv), g.nnz_u, g.nnz_v, beta, N.ptr(values),
                                    N.ptr(a_diag), N.ptr(gh), N.ptr(out), N.ptr(None if base is None else _f32(base)),
                                    N.ptr(total), N.stream()), "dpiso_h_apply_adj")
    return out if base is None else (out, total)


@_on_device_of(0)
def predictor_rhs_adj(g, grhs, dirichlet_u8, dy, dx, beta, want_force, want_dvals, solve_stats=None):
    """solve_stats: int32 [B, 2, 4] of the transposed predictor solve that produced grhs -> samples whose solve raised the
    NaN warning contribute nothing (linear_solver.py:169-173, sample by sample)."""
    grhs = _f32(grhs)
    b = grhs.shape[0]
    gvel = torch.empty_like(grhs)
    gfree = torch.empty_like(grhs)
    gforce = torch.empty_like(grhs) if want_force else None
    gdvals = torch.empty_like(grhs) if want_dvals else None
    N.check(N.lib.dpiso_predictor_rhs_adj(b, g.ny, g.nx, dy, dx, beta, N.ptr(dirichlet_u8), N.ptr(grhs),
                                          N.ptr(solve_stats), N.ptr(gvel), N.ptr(gforce), N.ptr(gdvals), N.ptr(gfree),
                                          N.stream()),
            "dpiso_predictor_rhs_adj")
    return gvel, gforce, gdvals, gfree


# test hook: fill solver workspaces with NaN bit patterns before every call (the kernels must never read workspace they
# have not written in the same call)
POISON_SCRATCH = False


@_on_device_of(0)
def bicgstab_ilu(g, values, rhs, x0, tol, max_it, transpose=False, negate=False, pivots_out=None, pivots_in=None,
                 fp64=False):
    """-> x [B, nf], stats int32 [B, 2, 4] (iterations, restarts, warn, exit kind), warn float32 [1].
    Solves (values or, with negate, -values) x = rhs.  pivots_out / pivots_in [B, nf] float32: ILU(0) pivots written by /
    taken from the solve of the other orientation (factor reuse; honoured where `factor_reuse_supported`).
    fp64: the cast_to_double=True path of the reference's solver class (fp32 in, fp64 solve, fp32 out; no factor reuse)."""
    values, rhs, x0 = _f32(values), _f32(rhs), _f32(x0)
    b = rhs.shape[0]
    tu, tv = g.tables(transpose)
    if fp64:
        ws = g.scratch("bicgstab_f64", b * 2 * N.lib.dpiso_bicgstab_f64_workspace_bytes(C.byref(tu), C.byref(tv)))
        if POISON_SCRATCH:
            ws.fill_(255)
        x = torch.empty_like(rhs)
        stats = torch.empty((b, 2, 4), dtype=torch.int32, device=rhs.device)
        warn = torch.empty(1, dtype=torch.float32, device=rhs.device)
        N.check(N.lib.dpiso_bicgstab_ilu_f64(b, C.byref(tu), C.byref(tv), g.nnz_u, g.nnz_v, N.ptr(values), int(bool(negate)),
                                             N.ptr(rhs), N.ptr(x0), float(tol), int(max_it), N.ptr(x), N.ptr(stats),
                                             N.ptr(warn), N.ptr(ws), N.stream()), "dpiso_bicgstab_ilu_f64")
        return x, stats, warn
    ws_floats = N.lib.dpiso_bicgstab_workspace_floats(C.byref(tu), C.byref(tv))
    ws = g.scratch("bicgstab", b * 2 * ws_floats * 4)
    if POISON_SCRATCH:
        ws.fill_(255)                                     # every float a NaN: a read of unwritten workspace shows up in x
    x = torch.empty_like(rhs)
    stats = torch.empty((b, 2, 4), dtype=torch.int32, device=rhs.device)
    warn = torch.empty(1, dtype=torch.float32, device=rhs.device)
    N.check(N.lib.dpiso_bicgstab_ilu(b, C.byref(tu), C.byref(tv), g.nnz_u, g.nnz_v, N.ptr(values), int(bool(negate)),
                                     N.ptr(rhs), N.ptr(x0), float(tol), int(max_it), N.ptr(x), N.ptr(stats), N.ptr(warn),
                                     N.ptr(pivots_out), N.ptr(pivots_in), N.ptr(ws), N.stream()), "dpiso_bicgstab_ilu")
    return x, stats, warn


def factor_reuse_supported(g):
    """True when the predictor kernel of this grid can write / take ILU(0) pivots (the row-major kernel)."""
    tu, tv = g.tables(False)
    return bool(N.lib.dpiso_bicgstab_supports_factor_reuse(C.byref(tu), C.byref(tv)))


@_on_device_of(0)
def laplace(g, active, fluid, k_faces, mode, beta, dx_factor, fp64=True):
    """-> lap [B, nc, 5]; mode 0: k_faces = scaling field flattened [v, u]; mode 1: k_faces = a_diag [u, v]."""
    k_faces = _f32(k_faces)
    b = k_faces.shape[0]
    lap = torch.empty((b, g.nc, 5), dtype=torch.float64 if fp64 else torch.float32, device=k_faces.device)
    fn = N.lib.dpiso_laplace_f64 if fp64 else N.lib.dpiso_laplace_f32
    N.check(fn(b, g.ny, g.nx, N.ptr(active), N.ptr(fluid), N.ptr(k_faces), int(mode), beta, dx_factor, N.ptr(lap),
               N.stream()), "dpiso_laplace")
    return lap


@_on_device_of(0)
def pressure_cg(g, lap, div, accuracy, max_it, residual_reset, rank_deficient):
    """-> pressure float32 [B, nc], iterations int32 [B].  lap fp64: fp32 divergence in, fp64 solve, fp32 out
    (cast_to_double path); lap fp32: everything fp32."""
    b = div.shape[0]
    div = _f32(div).reshape(b, g.nc)
    x32 = torch.empty((b, g.nc), dtype=torch.float32, device=div.device)
    its = torch.empty(b, dtype=torch.int32, device=div.device)
    fp64 = lap.dtype == torch.float64
    ws_bytes = N.lib.dpiso_pressure_cg_workspace_bytes(b, g.ny, g.nx, 8 if fp64 else 4, 0 if fp64 else 1)
    ws = g.scratch("pressure_cg", ws_bytes) if ws_bytes else None    # only grids whose state does not fit on chip
    if fp64:
        rc = N.lib.dpiso_pressure_cg_mixed(b, g.ny, g.nx, int(g.per_x), int(g.per_y), N.ptr(lap), N.ptr(div),
                                           float(accuracy), int(max_it), int(residual_reset), int(rank_deficient),
                                           N.ptr(x32), N.ptr(its), N.ptr(ws), N.stream())
    else:
        rc = N.lib.dpiso_pressure_cg_f32(b, g.ny, g.nx, int(g.per_x), int(g.per_y), N.ptr(lap), N.ptr(div),
                                         float(accuracy), int(max_it), int(residual_reset), int(rank_deficient),
                                         N.ptr(x32), None, N.ptr(its), N.ptr(ws), N.stream())
    N.check(rc, "dpiso_pressure_cg")
    return x32, its


def pressure_cg_on_chip(g, batch):
    """True when the pressure CG of this grid keeps its solver state on chip (cluster-resident kernel, no workspace)."""
    return N.lib.dpiso_pressure_cg_workspace_bytes(int(batch), g.ny, g.nx, 8, 0) == 0


def pressure_cg_config():
    out = (C.c_int * 5)()
    N.lib.dpiso_pressure_cg_last_config(out)
    return dict(cluster=out[0], threads=out[1], cells_per_thread=out[2], smem_bytes=out[3], variant=out[4])


def to_device_masks(sim, g):
    """Device copies of the SimulationParameters masks in the layouts the kernels read (cached on `sim`)."""
    # the cache is keyed by the grid AND by the identity of the mask objects: a SimulationParameters reused on another
    # resolution, or whose masks were replaced, gets fresh (re-validated) device copies
    key = ("_dpiso_masks", str(g.device), g.ny, g.nx, g.per_y, g.per_x, id(sim.dirichlet_mask), id(sim.active_mask),
           id(sim.accessible_mask), id(sim.no_slip_mask))
    cache = getattr(sim, "_dpiso_cache", None)
    if not isinstance(cache, dict):
        cache = sim._dpiso_cache = {}
    hit = cache.get(key)
    if hit is not None:
        return hit
    from .grids import as_tensor, flatten_staggered_data
    dm = as_tensor(sim.dirichlet_mask, dtype=None)
    dm = flatten_staggered_data(dm.to(torch.float32), coord_flip=True)[0]
    if dm.numel() != g.nf:
        raise ValueError("dirichlet_mask does not match the grid")
    active = as_tensor(sim.active_mask).reshape(-1)
    access = as_tensor(sim.accessible_mask).reshape(-1)
    nm = (g.ny + 2) * (g.nx + 2)
    if active.numel() != nm or access.numel() != nm:
        raise ValueError("active/accessible masks must be [1, ny+2, nx+2, 1]")
    if sim.no_slip_mask is None:
        noslip = torch.zeros(nm, dtype=torch.uint8)
    else:
        noslip = as_tensor(sim.no_slip_mask, dtype=None).reshape(-1)
        if noslip.numel() < nm:
            raise ValueError("no_slip_mask must be indexable as the padded centred grid (ny+2)*(nx+2) "
                             "(CUDAsrc/central_difference_csr_op.cu.cc:251-253)")
        # the reference kernel reads noSlipWall[centeredNeighborIdx] and nothing else: a larger array (e.g. the
        # np.zeros_like(dirichlet_mask) of combined_training_integrated.py:536) is read through its first nm entries
        noslip = (noslip[:nm] != 0).to(torch.uint8)
    acc_np, act_np = access.cpu().numpy().reshape(g.ny + 2, g.nx + 2), active.cpu().numpy().reshape(g.ny + 2, g.nx + 2)
    prod = acc_np * act_np + (1 - acc_np) * (1 - act_np)        # piso_cuda_pressure_solver.py:84-87
    rank_def = bool(np.prod(prod[0, 1:-1]) * np.prod(prod[-1, 1:-1]) * np.prod(prod[1:-1, 0]) * np.prod(prod[1:-1, -1]))
    m = dict(dirichlet=(dm != 0).to(torch.uint8).contiguous().to(g.device), active=active.contiguous().to(g.device),
             access=access.contiguous().to(g.device), noslip=noslip.contiguous().to(g.device), rank_deficient=rank_def)
    if len(cache) > 8:
        cache.clear()
    cache[key] = m
    return m
